"""dft.gen_grid.Grids, dft.numint.eval_ao and dft.RKS as grid.py:6-40 and dft.py:281-291 use them."""
import numpy as np

from quantum_compute_dft_b200 import molgrid as M


class _Grids:
    def __init__(self, mol):
        self.mol, self.level, self.prune = mol, 3, "default"
        self.coords = self.weights = None

    def build(self):
        # the repo's synthetic atom-centred grid with PySCF's level-3 per-atom point counts (molgrid.make_grid)
        self.coords, self.weights, _ = M.make_grid(self.mol._mol, scale=1.0)
        return self


class gen_grid:
    Grids = _Grids


class numint:
    @staticmethod
    def eval_ao(mol, coords, deriv=0):
        """(ngrid, nao) for deriv=0; (4, ngrid, nao) = value, d/dx, d/dy, d/dz for deriv=1 (grid.py:30-31,38)."""
        from oracle import oracle as O
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if deriv == 0:
            return O.eval_ao(coords, mol._basis, deriv=0)
        ao, grad = O.eval_ao(coords, mol._basis, deriv=1)
        return np.concatenate([ao[None], grad], axis=0)


class RKS:
    """The comparison calculation dft.py:281-291 prints: an SCF on the same molecule / basis / grid with the CPU
    oracle's XC in EXACT-functional mode (potentials = derivatives of the energies, the libxc convention)."""
    _XC = {"slater,vwn5": 0, "pbe,pbe": 1, "b3lyp": 2}

    def __init__(self, mol):
        self.mol, self.xc, self.e_tot, self.converged = mol, "slater,vwn5", None, False

    def kernel(self):
        import scf_driver
        from oracle import oracle as O
        from scf_backends import OracleBackend
        xc_type = self._XC[self.xc.lower().replace(" ", "")]
        fn = ["LDA", "GGA", "B3LYP"][xc_type]
        S, H, eri, e_nuc = self.mol._integrals()
        coords, weights, _ = M.make_grid(self.mol._mol, scale=1.0)
        be = OracleBackend(O, fn, self.mol._basis, coords, weights, eri, mode=1)
        self.e_tot, _, _, self.converged = scf_driver.run_scf(S, H, e_nuc, self.mol.nelec[1], be, fn, diis=True)
        return self.e_tot
