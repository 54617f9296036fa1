import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_count():
    try:
        from quantum_compute_dft_b200 import cuda_rt
        return cuda_rt.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engine_lib():
    """Path of the built product library (built in-tree; never a fallback)."""
    from quantum_compute_dft_b200 import build as B
    from quantum_compute_dft_b200.solver import DEFAULT_LIB
    if not os.path.exists(DEFAULT_LIB):
        B.build()
    return DEFAULT_LIB


@pytest.fixture(scope="session")
def h2_fixture():
    """The reference's only real grid fixture (grid_txt/h2_grid.txt) + H2/STO-3G."""
    from quantum_compute_dft_b200 import molgrid as M
    g = np.load(os.path.join(ROOT, "tests", "golden", "h2_grid.npz"))
    mol = M.Molecule("H2", ["H", "H"], np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.7122 * M.BOHR_PER_ANGSTROM]]))
    basis = M.sto3g_basis(mol, renormalize=False)
    dm = np.full((2, 2), 0.5959166139336604)
    return mol, basis, g["coords"], g["weights"], dm
