#!/usr/bin/env python3
"""One-off importer (run in the build container only, where /root/reference exists).

Reads the geometries that BASELINE.json's configs name from the reference's
atom_txt/*.xyz (dft.py:97-99 skips the two header lines; we do the same and
ignore the atom count on line 1, which is wrong for DHA) and the only real grid
fixture the reference ships (grid_txt/h2_grid.txt, columns `atom x y z w w`,
grid.py:11-14).  Writes

  quantum_compute_dft_b200/data/molecules.json   (element symbols + Angstrom coordinates)
  tests/golden/h2_grid.npz                       (coords [Bohr], weights)

Nothing at run time (tests, smoke, bench) reads /root/reference.
"""
import json, os, sys
import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["H2", "H2O", "Benzene", "DHA", "C33H56N7O17P3S", "CH4", "NH3", "H2S", "Ethanol"]

def read_xyz(path):
    atoms = []
    with open(path) as f:
        lines = f.readlines()[2:]
    for ln in lines:
        p = ln.split()
        if len(p) < 4:
            continue
        atoms.append([p[0], float(p[1]), float(p[2]), float(p[3])])
    return atoms

def main():
    mols = {}
    for n in NAMES:
        mols[n] = read_xyz(os.path.join(REF, "atom_txt", n + ".xyz"))
    out = os.path.join(ROOT, "quantum_compute_dft_b200", "data", "molecules.json")
    with open(out, "w") as f:
        json.dump({"unit": "angstrom", "source": "reference atom_txt/*.xyz (lines 3..)", "molecules": mols}, f, indent=0)
    print("wrote", out, {k: len(v) for k, v in mols.items()})
    g = np.loadtxt(os.path.join(REF, "grid_txt", "h2_grid.txt"))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "h2_grid.npz"),
                        atom=g[:, 0].astype(np.int32), coords=g[:, 1:4], weights=g[:, 4])
    print("h2 grid", g.shape)

if __name__ == "__main__":
    main()
