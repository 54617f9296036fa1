"""Recipes of the golden cases (inputs are rebuilt deterministically; outputs live in tests/golden/ref_*.npz,
produced by tools/make_reference_golden.py from the reference's own CUDA on a B200)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["h2_fixture", "h2o_small", "benzene_small", "h2s_small", "random_257x33"]


def build_case(name):
    from quantum_compute_dft_b200 import molgrid as M
    if name == "h2_fixture":
        g = np.load(os.path.join(ROOT, "tests", "golden", "h2_grid.npz"))
        mol = M.Molecule("H2", ["H", "H"], np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.7122 * M.BOHR_PER_ANGSTROM]]))
        basis = M.sto3g_basis(mol, renormalize=False)
        ao, grad = M.eval_ao_numpy(g["coords"], basis, deriv=1)
        return np.full((2, 2), 0.5959166139336604), ao, g["weights"].copy(), grad
    if name.endswith("_small"):
        molname, scale = {"h2o_small": ("H2O", 0.1), "benzene_small": ("Benzene", 0.02),
                          "h2s_small": ("H2S", 0.1)}[name]
        mol = M.load_molecule(molname)
        basis = M.sto3g_basis(mol)
        coords, w, _ = M.make_grid(mol, scale=scale)
        ao, grad = M.eval_ao_numpy(coords, basis, deriv=1)
        dm = M.synthetic_density(M.overlap_matrix(basis), mol.nocc, seed=0)
        return dm, ao, w, grad
    if name == "random_257x33":
        rng = np.random.default_rng(20261018)
        ngrid, nao = 257, 33
        scale = 10 ** rng.uniform(-6, 0, (ngrid, 1))
        ao = rng.standard_normal((ngrid, nao)) * scale
        grad = rng.standard_normal((3, ngrid, nao)) * scale
        C = rng.standard_normal((nao, 16)) / np.sqrt(nao)
        return 2.0 * C @ C.T, ao, rng.uniform(0.0, 1.0, ngrid), grad
    raise KeyError(name)
