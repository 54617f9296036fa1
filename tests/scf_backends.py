"""CPU backend for quantum_compute_dft_b200.scf.run_scf built on the oracle (test infrastructure only)."""
import numpy as np

XC = {"LDA": 0, "GGA": 1, "B3LYP": 2}


class OracleBackend:
    def __init__(self, oracle, functional, basis, coords, weights, eri, mode=0):
        self.o, self.xc_type, self.eri, self.w, self.mode = oracle, XC[functional.upper()], eri, weights, mode
        self.ao, self.grad = oracle.eval_ao(coords, basis, deriv=1)

    def coulomb_exchange(self, dm):
        return np.einsum("ijkl,kl->ij", self.eri, dm), np.einsum("ijkl,jl->ik", self.eri, dm)

    def xc(self, dm):
        return self.o.compute_xc(self.xc_type, dm, self.ao, self.w, self.grad, mode=self.mode)


def h_chain(positions_bohr):
    """Hydrogen atoms on the z axis at the given positions: an s-only STO-3G system."""
    from quantum_compute_dft_b200 import molgrid as M
    z = np.asarray(positions_bohr, dtype=np.float64)
    mol = M.Molecule(f"H{z.size}", ["H"] * z.size, np.stack([np.zeros_like(z), np.zeros_like(z), z], axis=1))
    return mol, M.sto3g_basis(mol, renormalize=False)
