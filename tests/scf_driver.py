"""TEST INFRASTRUCTURE (not product code). Host-side SCF driver: the reference's main loop (dft.py:183-266) over this engine's C ABI.

The reference takes its one- and two-electron integrals from PySCF (grid.py:61-66), which is not
installable here; for s-type basis functions (hydrogen chains in STO-3G) the integrals are closed-form,
so this module carries them itself.  That is enough to run the reference's whole per-iteration sequence
-- upload D (dft.py:200), J (dft.py:203), E_xc/V_xc (dft.py:205-208), K for B3LYP (dft.py:218), Fock
build, generalised eigenproblem, energy (dft.py:223-240), convergence test (dft.py:248) -- end to end
through `DFT_ComputeCoulomb[Exchange]` and `DFT_ComputeXC`, and to check the third parity criterion of
the north_star: the converged total energy.

`run_scf` takes a BACKEND object with `coulomb_exchange(dm) -> (J, K)` and `xc(dm) -> (E_xc, Vxc_raw)`;
`EngineBackend` below is the GPU one.  (A CPU backend built on the oracle lives in tests/.)
"""
import math

import numpy as np
from scipy.linalg import eigh
from scipy.special import erf

from quantum_compute_dft_b200.cuda_rt import DeviceArray
from quantum_compute_dft_b200.molgrid import ATOMIC_NUMBER


# --------------------------------------------------------------------------- s-type integrals
def _boys0(t):
    t = np.asarray(t, dtype=np.float64)
    small = t < 1e-12
    ts = np.where(small, 1.0, t)
    return np.where(small, 1.0 - t / 3.0, 0.5 * np.sqrt(np.pi / ts) * erf(np.sqrt(ts)))


def s_kinetic(mol, basis):
    """Kinetic-energy matrix alone (the shim's int1e_kin, grid.py:62)."""
    return s_integrals(mol, basis, kinetic_only=True)


def s_integrals(mol, basis, kinetic_only=False):
    """Overlap S, core Hamiltonian T + V_nuc, ERI (nao,nao,nao,nao) and E_nuc for a basis of contracted s
    functions (`basis.prim_coef` already contains the primitive normalisation)."""
    if np.any(basis.shell_l != 0):
        raise ValueError("s_integrals: s shells only (hydrogen / helium-like STO-nG bases)")
    n = basis.nao
    # flatten to primitives: exponent, coefficient, centre, owning AO
    ex, co, ce, ow = [], [], [], []
    for s in range(basis.nshell):
        for k in range(basis.shell_prim_off[s], basis.shell_prim_off[s] + basis.shell_nprim[s]):
            ex.append(basis.prim_exp[k]); co.append(basis.prim_coef[k]); ce.append(basis.shell_xyz[s]); ow.append(basis.shell_ao_off[s])
    ex, co, ce, ow = np.array(ex), np.array(co), np.array(ce), np.array(ow)
    npn = ex.size
    a, b = ex[:, None], ex[None, :]
    p = a + b
    ab2 = ((ce[:, None, :] - ce[None, :, :]) ** 2).sum(-1)
    kab = np.exp(-a * b / p * ab2)                      # Gaussian product prefactor
    P = (a[..., None] * ce[:, None, :] + b[..., None] * ce[None, :, :]) / p[..., None]
    cc = co[:, None] * co[None, :]
    s_pp = cc * (np.pi / p) ** 1.5 * kab
    t_pp = s_pp * (a * b / p) * (3.0 - 2.0 * a * b / p * ab2)
    v_pp = np.zeros_like(s_pp)
    for sym, R in zip(mol.symbols, mol.coords):
        pc2 = ((P - R) ** 2).sum(-1)
        v_pp -= ATOMIC_NUMBER[sym] * cc * 2.0 * np.pi / p * kab * _boys0(p * pc2)
    # contract primitives -> AOs
    C = np.zeros((npn, n))
    C[np.arange(npn), ow] = 1.0
    if kinetic_only:
        return C.T @ t_pp @ C
    S = C.T @ s_pp @ C
    H = C.T @ (t_pp + v_pp) @ C
    # (ab|cd) over primitive pairs
    pq = p.reshape(-1)
    Pq = P.reshape(-1, 3)
    pref = (cc * kab).reshape(-1)
    r2 = ((Pq[:, None, :] - Pq[None, :, :]) ** 2).sum(-1)
    rho = pq[:, None] * pq[None, :] / (pq[:, None] + pq[None, :])
    eri_pp = 2.0 * np.pi ** 2.5 / (pq[:, None] * pq[None, :] * np.sqrt(pq[:, None] + pq[None, :])) * \
        pref[:, None] * pref[None, :] * _boys0(rho * r2)
    C2 = np.einsum("pi,qj->pqij", C, C).reshape(npn * npn, n * n)
    eri = (C2.T @ eri_pp @ C2).reshape(n, n, n, n)
    e_nuc = 0.0
    for i in range(mol.natm):
        for j in range(i):
            e_nuc += ATOMIC_NUMBER[mol.symbols[i]] * ATOMIC_NUMBER[mol.symbols[j]] / \
                math.sqrt(((mol.coords[i] - mol.coords[j]) ** 2).sum())
    return S, H, np.ascontiguousarray(eri), e_nuc


# --------------------------------------------------------------------------- backends
class EngineBackend:
    """J, K and XC on the GPU through the C ABI, with the device-resident arrays of dft.py:155-176."""

    def __init__(self, solver, basis, coords, weights, eri):
        self.solver, self.nao, self.ngrid = solver, basis.nao, coords.shape[0]
        self.functional = solver.functional_type
        self.d_coords = DeviceArray.from_host(coords)
        self.d_w = DeviceArray.from_host(weights)
        self.d_ao = DeviceArray((self.ngrid, self.nao))
        self.d_grad = DeviceArray((3, self.ngrid, self.nao)) if self.functional != "LDA" else None
        solver.eval_ao(self.d_coords, basis, self.d_ao, self.d_grad)          # replaces grid.py:30,38
        self.d_eri = DeviceArray.from_host(eri.reshape(self.nao ** 2, self.nao ** 2))
        self.d_dm = DeviceArray((self.nao, self.nao), zero=True)
        self.d_J = DeviceArray((self.nao, self.nao), zero=True)
        self.d_K = DeviceArray((self.nao, self.nao), zero=True)
        self.d_v = DeviceArray((self.nao, self.nao), zero=True)

    def coulomb_exchange(self, dm):
        self.d_dm.set(dm)                                                       # dft.py:200
        self.solver.compute_coulomb_exchange(self.nao, self.d_eri, self.d_dm, self.d_J, self.d_K)
        self.solver.synchronize()
        return self.d_J.get(), self.d_K.get()

    def xc(self, dm):
        e = self.solver.compute_xc(self.ngrid, self.nao, self.d_dm, self.d_ao, self.d_w, self.d_v, self.d_grad)
        return e, self.d_v.get()

    # ---- device-resident variant: J, V_xc and K never leave the GPU
    def fock(self, dm, hcore, c_hf):
        """Upload D, run J/K and XC, assemble F on the device; returns (F on the host for eigh, E_xc)."""
        if not hasattr(self, "d_h"):
            self.d_h = DeviceArray.from_host(hcore)
            self.d_F = DeviceArray((self.nao, self.nao), zero=True)
        self.d_dm.set(dm)
        self.solver.compute_coulomb_exchange(self.nao, self.d_eri, self.d_dm, self.d_J, self.d_K)
        e_xc = self.solver.compute_xc(self.ngrid, self.nao, self.d_dm, self.d_ao, self.d_w, self.d_v, self.d_grad)
        self.solver.build_fock(self.nao, self.d_h, self.d_J, self.d_v, self.d_K if c_hf else None, c_hf, self.d_F)
        self.solver.synchronize()
        return self.d_F.get(), e_xc

    def energies(self, dm_new, c_hf):
        """(E_one, E_coul, E_hf) with the NEW density and the J, K still resident from `fock`."""
        self.d_dm.set(dm_new)
        return self.solver.scf_energies(self.nao, self.d_dm, self.d_h, self.d_J, self.d_K if c_hf else None, c_hf)


# --------------------------------------------------------------------------- the loop (dft.py:183-266)
class CDIIS:
    """Pulay DIIS on the commutator error S D F - F D S (what PySCF's scf.diis.CDIIS, dft.py:184,225, does)."""

    def __init__(self, space=8):
        self.space, self._f, self._e = space, [], []

    def update(self, s, d, f):
        sdf = s @ d @ f
        self._f.append(np.array(f, dtype=np.float64))
        self._e.append((sdf - sdf.T).ravel())
        if len(self._f) > self.space:
            self._f.pop(0)
            self._e.pop(0)
        n = len(self._f)
        if n < 2:
            return f
        B = np.zeros((n + 1, n + 1))
        for i in range(n):
            for j in range(n):
                B[i, j] = self._e[i] @ self._e[j]
        B[n, :n] = B[:n, n] = 1.0
        rhs = np.zeros(n + 1)
        rhs[n] = 1.0
        try:
            c = np.linalg.solve(B, rhs)[:n]
        except np.linalg.LinAlgError:
            c = np.linalg.lstsq(B, rhs, rcond=None)[0][:n]
        return sum(ci * fi for ci, fi in zip(c, self._f))


def run_scf(S, hcore, e_nuc, nocc, backend, functional, max_cycle=200, e_tol=1e-8, dm_tol=1e-6, verbose=False, diis=False):
    """The reference's SCF loop (dft.py:199-248).  `diis=False`: plain fixed-point iteration (both backends then
    follow the identical iteration); `diis=True`: with the commutator DIIS of dft.py:225.
    Returns (E_tot, dm, cycles, converged)."""
    c_hf = 0.2 if functional.upper() == "B3LYP" else 0.0                        # dft.py:197
    _, C = eigh(hcore, S)
    dm = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    e_old = 0.0
    adiis = CDIIS() if diis else None
    for cycle in range(max_cycle):
        J, K = backend.coulomb_exchange(dm)                                     # dft.py:203, :218
        e_xc, v_raw = backend.xc(dm)                                            # dft.py:205-208
        vxc = 0.5 * (v_raw + v_raw.T)                                           # dft.py:212
        F = hcore + J + vxc - c_hf * 0.5 * K                                    # dft.py:221-223
        if adiis is not None:
            F = adiis.update(S, dm, F)                                          # dft.py:225
        _, C = eigh(F, S)
        dm_new = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        e_one = np.sum(dm_new * hcore)                                          # dft.py:230-236
        e_coul = 0.5 * np.sum(dm_new * J)
        e_hf = -0.25 * c_hf * np.sum(dm_new * K)
        e_tot = e_one + e_coul + e_xc + e_hf + e_nuc
        d_e, d_dm = e_tot - e_old, np.linalg.norm(dm_new - dm)
        if verbose:
            print(f"{cycle + 1:4d} {e_tot:18.10f} {d_e:15.6e} {d_dm:15.6e}")
        if abs(d_e) < e_tol and d_dm < dm_tol:                                  # dft.py:248
            return e_tot, dm_new, cycle + 1, True
        dm, e_old = dm_new, e_tot
    return e_tot, dm, max_cycle, False


def run_scf_device(S, hcore, e_nuc, nocc, backend, functional, max_cycle=200, e_tol=1e-8, dm_tol=1e-6):
    """The same loop with the Fock assembly and the energy sums on the device (`backend.fock`,
    `backend.energies`): per iteration D goes up and F comes down; J, V_xc and K stay on the GPU."""
    c_hf = 0.2 if functional.upper() == "B3LYP" else 0.0
    _, C = eigh(hcore, S)
    dm = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    e_old = 0.0
    for cycle in range(max_cycle):
        F, e_xc = backend.fock(dm, hcore, c_hf)
        _, C = eigh(F, S)
        dm_new = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        e_one, e_coul, e_hf = backend.energies(dm_new, c_hf)
        e_tot = e_one + e_coul + e_xc + e_hf + e_nuc
        d_e, d_dm = e_tot - e_old, np.linalg.norm(dm_new - dm)
        if abs(d_e) < e_tol and d_dm < dm_tol:
            return e_tot, dm_new, cycle + 1, True
        dm, e_old = dm_new, e_tot
    return e_tot, dm, max_cycle, False


def hartree_fock(S, hcore, eri, e_nuc, nocc, max_cycle=100):
    """Restricted Hartree-Fock with the same integrals (pins them against literature values in tests)."""
    _, C = eigh(hcore, S)
    dm = 2.0 * C[:, :nocc] @ C[:, :nocc].T
    e_old = 0.0
    for _ in range(max_cycle):
        J = np.einsum("ijkl,kl->ij", eri, dm)
        K = np.einsum("ijkl,jl->ik", eri, dm)
        F = hcore + J - 0.5 * K
        e = 0.5 * np.sum(dm * (hcore + F)) + e_nuc
        _, C = eigh(F, S)
        dm = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        if abs(e - e_old) < 1e-12:
            break
        e_old = e
    return e
