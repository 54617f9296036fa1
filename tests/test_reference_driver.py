"""The reference's UNMODIFIED Python driver against this repo's library: the real drop-in test.

`/root/reference/dft.py` (staged byte-for-byte by `make -C oracle ref` into the git-ignored oracle/_ref/driver/,
because the GPU box has no /root/reference) is executed as a subprocess, `python dft.py <LDA|GGA|B3LYP> <molecule>`,
in a scratch checkout layout whose ./weights/dft.so is THIS repo's engine.  `cupy` and `pyscf` -- not installable here
-- are the stand-ins of tests/shims (device arrays on the CUDA runtime; closed-form hydrogen integrals and
McMurchie-Davidson s/p integrals for H2O, synthetic grid, oracle AO evaluation, commutator DIIS).  Everything between `ctypes.CDLL("./weights/dft.so")` and the printed
"Total Energy" is the reference's own code driving the engine through the reference's own ABI.

Asserted: the driver converges, and its printed total energy equals an independent SCF whose J, K and XC come from the
CPU oracle (reference-compatible functionals) within the north_star's 1e-7 Ha.
"""
import filecmp
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver")

# molecule name -> XYZ body (Angstrom).  H2 is the reference's own atom_txt/H2.xyz; the chain is ours (an
# asymmetric system whose density is not fixed by symmetry, so the SCF and the DIIS really iterate)
H4 = "4\nasymmetric H4 chain (tests)\n" + "".join(
    f"H 0.0 0.0 {z / 1.8897261246:.10f}\n" for z in (0.0, 1.3, 3.1, 4.6))


def _layout(tmp_path, engine_lib):
    if not os.path.exists(os.path.join(DRIVER, "dft.py")):
        pytest.skip("oracle/_ref/driver/dft.py is not staged (make -C oracle ref, where /root/reference exists)")
    w = tmp_path / "checkout"
    (w / "weights").mkdir(parents=True)
    (w / "atom_txt").mkdir()
    (w / "grid_txt").mkdir()
    for f in ("dft.py", "grid.py"):
        shutil.copy(os.path.join(DRIVER, f), w / f)
    shutil.copy(os.path.join(DRIVER, "atom_txt", "H2.xyz"), w / "atom_txt" / "H2.xyz")
    if os.path.exists(os.path.join(DRIVER, "atom_txt", "H2O.xyz")):     # the reference's own water geometry (C1 / C3)
        shutil.copy(os.path.join(DRIVER, "atom_txt", "H2O.xyz"), w / "atom_txt" / "H2O.xyz")
    (w / "atom_txt" / "H4.xyz").write_text(H4)
    shutil.copy(engine_lib, w / "weights" / "dft.so")
    return w


def test_staged_driver_is_the_reference_byte_for_byte():
    """Where the reference tree exists (the build container), the staged copy must be identical to it."""
    if not (os.path.exists("/root/reference/dft.py") and os.path.exists(os.path.join(DRIVER, "dft.py"))):
        pytest.skip("needs both /root/reference and the staged copy")
    for f in ("dft.py", "grid.py", os.path.join("atom_txt", "H2.xyz"), os.path.join("atom_txt", "H2O.xyz")):
        assert filecmp.cmp(os.path.join("/root/reference", f), os.path.join(DRIVER, f), shallow=False), f


def _oracle_scf(functional, xyz_body):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
    import scf_driver
    from oracle import oracle as O
    from pyscf import gto
    from quantum_compute_dft_b200 import molgrid as M
    from scf_backends import OracleBackend
    mol = gto.Mole(atom=xyz_body).build()
    S, H, eri, e_nuc = mol._integrals()
    coords, weights, _ = M.make_grid(mol._mol, scale=1.0)
    be = OracleBackend(O, functional, mol._basis, coords, weights, eri, mode=0)
    e, _, _, ok = scf_driver.run_scf(S, H, e_nuc, mol.nelec[1], be, functional, diis=True)
    assert ok
    return e


@pytest.mark.gpu
@pytest.mark.parametrize("functional", ["LDA", "GGA", "B3LYP"])
@pytest.mark.parametrize("molecule", ["H2", "H4", "H2O"])
def test_unmodified_reference_driver_runs_on_this_library(oracle, engine_lib, tmp_path, functional, molecule):
    """`python dft.py LDA H2O` and `python dft.py B3LYP H2O` are BASELINE.json's configs C1 and C3 verbatim (the
    reference's own atom_txt/H2O.xyz, STO-3G, level-3-sized grid; s/p integrals from tests/gauss_integrals.py)."""
    w = _layout(tmp_path, engine_lib)
    if not os.path.exists(w / "atom_txt" / f"{molecule}.xyz"):
        pytest.skip(f"atom_txt/{molecule}.xyz is not staged (make -C oracle ref)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "shims"), os.path.join(ROOT, "tests"), ROOT,
                                         env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, "dft.py", functional, molecule], cwd=w, env=env, capture_output=True, text=True,
                       timeout=900)
    out = r.stdout
    assert r.returncode == 0, (out[-3000:], r.stderr[-3000:])
    assert "Converged!" in out, out[-3000:]
    e_driver = float(re.search(r"Total Energy: (-?\d+\.\d+) Ha", out).group(1))
    xc_ms = float(re.search(r"XC\(Exc\+Vxc\) Time: (\d+\.\d+) ms", out).group(1))
    body = "".join(open(w / "atom_txt" / f"{molecule}.xyz").readlines()[2:])
    e_oracle = _oracle_scf(functional, body)
    # the driver prints 8 decimals; the criterion is the north_star's converged-energy tolerance
    assert abs(e_driver - e_oracle) <= 1e-7, (functional, molecule, e_driver, e_oracle)
    # the comparison run the driver prints at the end (exact-functional SCF standing in for PySCF): B3LYP, whose
    # reference potentials ARE the derivatives of its energies, must agree; LDA/GGA differ by the reference's D1-D3
    diff = float(re.search(r"Difference\s*:\s*([0-9.eE+-]+) Hartree", out).group(1))
    if functional == "B3LYP":
        assert diff <= 1e-6, out[-800:]
    assert xc_ms > 0.0
    print(f"{functional} {molecule}: E = {e_driver:.8f} Ha (oracle SCF {e_oracle:.10f}), XC {xc_ms:.3f} ms/iter, "
          f"|E - exact-functional SCF| = {diff:.2e}")


@pytest.mark.gpu
@pytest.mark.parametrize("functional", ["LDA", "B3LYP"])
def test_unmodified_reference_driver_fans_out(oracle, engine_lib, tmp_path, functional):
    """Single-process multi-GPU behind the unmodified driver (csrc/fanout.cu): the environment alone makes dft.py's own
    DFT_ComputeXC calls fan out -- here to three virtual devices, so that a one-GPU box runs the whole mechanism.  The
    driver converges to the same energy, and the shards were cut exactly once for the whole SCF run (the AO arrays are
    uploaded once, dft.py:155,172; only d_dm changes per iteration, dft.py:200)."""
    w = _layout(tmp_path, engine_lib)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "shims"), os.path.join(ROOT, "tests"), ROOT,
                                         env.get("PYTHONPATH", "")])
    env.update({"DFT_B200_VIRTUAL_DEVICES": "3", "DFT_B200_DEVICES_MIN_WORK": "0", "DFT_B200_VERBOSE": "1"})
    r = subprocess.run([sys.executable, "dft.py", functional, "H4"], cwd=w, env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "Converged!" in r.stdout, r.stdout[-3000:]
    cuts = re.findall(r"fan-out: shards of (\d+) x (\d+) \((\w+)\) cut for 3 devices", r.stderr)
    assert len(cuts) == 1 and cuts[0][2] == functional, r.stderr[-2000:]
    e_driver = float(re.search(r"Total Energy: (-?\d+\.\d+) Ha", r.stdout).group(1))
    body = "".join(open(w / "atom_txt" / "H4.xyz").readlines()[2:])
    assert abs(e_driver - _oracle_scf(functional, body)) <= 1e-7
