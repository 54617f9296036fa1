"""Host-side inputs for the XC path: molecules, STO-3G shell tables, synthetic grids, densities.

The reference obtains all of this from PySCF (grid.py:42-67: gto.Mole(basis='sto-3g'),
gen_grid.Grids(level=3), numint.eval_ao).  PySCF is not installable in this
environment, so the benchmark / test harness builds equivalent
inputs itself:

* geometries of the molecules BASELINE.json's configs name (data/molecules.json,
  imported once from the reference's atom_txt/*.xyz by tools/import_reference_inputs.py);
* STO-3G s/p shell tables for H, C, N, O, P, S (the four
  self-checks listed there are unit tests in tests/test_molgrid.py);
* an atom-centred synthetic integration grid with PySCF level-3 per-atom point
  counts (Treutler-Ahlrichs M4 radial x Gauss-Legendre/uniform angular product
  rule, fuzzy-cell partition weights);
* a synthetic closed-shell density matrix D = 2 C C^T, C = S^{-1/2} Q.

This module is numpy-only host logic.  AO values on the GPU come from the
DFT_EvalAO entry point (csrc/ao_eval.cu); `eval_ao_numpy` below is the slow
host statement of the same formulas used by small CPU tests.
"""
import json
import os
from dataclasses import dataclass

import numpy as np

BOHR_PER_ANGSTROM = 1.0 / 0.52917721092  # dft.py / PySCF convention (CODATA 2010)
AO_EXP_CUTOFF = 60.0

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "molecules.json")

# ----------------------------------------------------------------------------- STO-3G
_S1 = (0.15432897, 0.53532814, 0.44463454)
_S2 = (-0.09996723, 0.39951283, 0.70011547)
_P2 = (0.15591627, 0.60768372, 0.39195739)
_S3 = (-0.21962037, 0.22559543, 0.90039843)
_P3 = (0.01058760, 0.59516701, 0.46200101)

_STO3G_EXP = {
    "H": [(3.42525091, 0.62391373, 0.16885540)],
    "C": [(71.6168370, 13.0450960, 3.5305122), (2.9412494, 0.6834831, 0.2222899)],
    "N": [(99.1061690, 18.0523120, 4.8856602), (3.7804559, 0.8784966, 0.2857144)],
    "O": [(130.7093200, 23.8088610, 6.4436083), (5.0331513, 1.1695961, 0.3803890)],
    "P": [(468.3656378, 85.3133856, 23.0891316), (28.0326396, 6.5141826, 2.1186143),
          (1.7431032, 0.4863213, 0.1903429)],
    "S": [(533.1257359, 97.1095183, 26.2816254), (33.3297517, 7.7451175, 2.5189526),
          (2.0291942, 0.5661400, 0.2215833)],
}
ATOMIC_NUMBER = {"H": 1, "C": 6, "N": 7, "O": 8, "P": 15, "S": 16}
# PySCF level-3 pruned point counts per atom and radial sizes
LEVEL3_POINTS = {"H": 10024, "C": 13902, "N": 14046, "O": 14262, "P": 18880, "S": 18880}
LEVEL3_NRAD = {"H": 50, "C": 75, "N": 75, "O": 75, "P": 80, "S": 80}


@dataclass
class Molecule:
    name: str
    symbols: list
    coords: np.ndarray  # (natm,3) Bohr

    @property
    def natm(self):
        return len(self.symbols)

    @property
    def nelectron(self):
        return int(sum(ATOMIC_NUMBER[s] for s in self.symbols))

    @property
    def nocc(self):
        return self.nelectron // 2


def load_molecule(name):
    with open(_DATA) as f:
        d = json.load(f)["molecules"]
    key = {k.lower(): k for k in d}.get(name.lower())
    if key is None:
        raise KeyError(f"unknown molecule {name!r}; have {sorted(d)}")
    atoms = d[key]
    return Molecule(key, [a[0] for a in atoms],
                    np.array([a[1:4] for a in atoms], dtype=np.float64) * BOHR_PER_ANGSTROM)


@dataclass
class Basis:
    """Flat shell tables (the layout DFT_EvalAO takes).  AO order follows PySCF: per
    atom all s shells first (1s,2s,3s), then the p shells (2p,3p), components x,y,z."""
    nshell: int
    nao: int
    shell_xyz: np.ndarray       # (nshell,3) Bohr
    shell_l: np.ndarray         # (nshell,) 0 or 1
    shell_ao_off: np.ndarray    # (nshell,)
    shell_prim_off: np.ndarray  # (nshell,)
    shell_nprim: np.ndarray     # (nshell,)
    prim_exp: np.ndarray        # (nprim_total,)
    prim_coef: np.ndarray       # contraction coefficient x primitive norm (x renormalisation)
    shell_atom: np.ndarray      # (nshell,)


def _prim_norm(l, a):
    if l == 0:
        return (2.0 * a / np.pi) ** 0.75
    return (128.0 * a ** 5 / np.pi ** 3) ** 0.25


def _contracted_self_overlap(l, exps, coefs):
    """<chi|chi> of a contracted shell component with normalised primitives."""
    s = 0.0
    for a, ca in zip(exps, coefs):
        for b, cb in zip(exps, coefs):
            p = a + b
            if l == 0:
                s += ca * cb * _prim_norm(0, a) * _prim_norm(0, b) * (np.pi / p) ** 1.5
            else:
                s += ca * cb * _prim_norm(1, a) * _prim_norm(1, b) * (np.pi / p) ** 1.5 / (2.0 * p)
    return s


def sto3g_basis(mol, renormalize=True):
    xyz, ls, ao_off, p_off, nprim, exps, coefs, satom = [], [], [], [], [], [], [], []
    nao = 0

    def add(ia, l, ex, co):
        nonlocal nao
        scale = 1.0 / np.sqrt(_contracted_self_overlap(l, ex, co)) if renormalize else 1.0
        xyz.append(mol.coords[ia]); ls.append(l); ao_off.append(nao); p_off.append(len(exps))
        nprim.append(len(ex)); satom.append(ia)
        for a, c in zip(ex, co):
            exps.append(a); coefs.append(c * _prim_norm(l, a) * scale)
        nao += 1 if l == 0 else 3

    for ia, sym in enumerate(mol.symbols):
        if sym not in _STO3G_EXP:
            raise KeyError(f"no STO-3G table for element {sym}")
        sh = _STO3G_EXP[sym]
        s_sets = [_S1, _S2, _S3][:len(sh)]
        p_sets = [None, _P2, _P3][:len(sh)]
        for ex, cs in zip(sh, s_sets):
            add(ia, 0, ex, cs)
        for ex, cp in zip(sh, p_sets):
            if cp is not None:
                add(ia, 1, ex, cp)
    return Basis(len(ls), nao, np.array(xyz, dtype=np.float64).reshape(-1, 3), np.array(ls, dtype=np.int32),
                 np.array(ao_off, dtype=np.int32), np.array(p_off, dtype=np.int32),
                 np.array(nprim, dtype=np.int32), np.array(exps), np.array(coefs),
                 np.array(satom, dtype=np.int32))


def eval_ao_numpy(coords, basis, deriv=0, exp_cutoff=AO_EXP_CUTOFF):
    """Host statement of DFT_EvalAO (same formulas and cutoff rule), vectorised per shell."""
    coords = np.asarray(coords, dtype=np.float64)
    ng = coords.shape[0]
    ao = np.zeros((ng, basis.nao))
    grad = np.zeros((3, ng, basis.nao)) if deriv else None
    for s in range(basis.nshell):
        d = coords - basis.shell_xyz[s]
        r2 = np.einsum("ij,ij->i", d, d)
        e0 = np.zeros(ng); e1 = np.zeros(ng)
        for k in range(basis.shell_prim_off[s], basis.shell_prim_off[s] + basis.shell_nprim[s]):
            a = basis.prim_exp[k]
            keep = a * r2 <= exp_cutoff
            t = np.where(keep, basis.prim_coef[k] * np.exp(-a * np.where(keep, r2, 0.0)), 0.0)
            e0 += t
            e1 += -2.0 * a * t
        o = basis.shell_ao_off[s]
        if basis.shell_l[s] == 0:
            ao[:, o] = e0
            if deriv:
                for c in range(3):
                    grad[c, :, o] = e1 * d[:, c]
        else:
            for j in range(3):
                ao[:, o + j] = d[:, j] * e0
                if deriv:
                    for c in range(3):
                        grad[c, :, o + j] = d[:, j] * d[:, c] * e1 + (e0 if j == c else 0.0)
    return (ao, grad) if deriv else ao


# ------------------------------------------------------------------- analytic overlap
def overlap_matrix(basis):
    """S_ij = <chi_i|chi_j> for contracted s/p Gaussians (closed form)."""
    n = basis.nao
    S = np.zeros((n, n))
    for s in range(basis.nshell):
        A = basis.shell_xyz[s]; la = basis.shell_l[s]; oa = basis.shell_ao_off[s]
        pa = range(basis.shell_prim_off[s], basis.shell_prim_off[s] + basis.shell_nprim[s])
        for t in range(s + 1):
            B = basis.shell_xyz[t]; lb = basis.shell_l[t]; ob = basis.shell_ao_off[t]
            pb = range(basis.shell_prim_off[t], basis.shell_prim_off[t] + basis.shell_nprim[t])
            AB2 = float(np.dot(A - B, A - B))
            blk = np.zeros((1 if la == 0 else 3, 1 if lb == 0 else 3))
            for i in pa:
                a, ca = basis.prim_exp[i], basis.prim_coef[i]
                for j in pb:
                    b, cb = basis.prim_exp[j], basis.prim_coef[j]
                    p = a + b
                    P = (a * A + b * B) / p
                    ss = ca * cb * (np.pi / p) ** 1.5 * np.exp(-a * b / p * AB2)
                    PA, PB = P - A, P - B
                    if la == 0 and lb == 0:
                        blk[0, 0] += ss
                    elif la == 1 and lb == 0:
                        blk[:, 0] += PA * ss
                    elif la == 0 and lb == 1:
                        blk[0, :] += PB * ss
                    else:
                        blk += (np.outer(PA, PB) + np.eye(3) / (2.0 * p)) * ss
            S[oa:oa + blk.shape[0], ob:ob + blk.shape[1]] = blk
            S[ob:ob + blk.shape[1], oa:oa + blk.shape[0]] = blk.T
    return S


def synthetic_density(S, nocc, seed=0, lindep=1e-8):
    """D = 2 C C^T with C = S^{-1/2} Q, Q the first nocc columns of the QR of a seeded Gaussian
    matrix.  Symmetric, PSD, rho >= 0 everywhere, tr(D S) = 2 nocc.
    S^{-1/2} is the symmetric (Loewdin) inverse square root U s^-1/2 U^T over the eigenvalues above
    lindep * s_max: unlike the canonical U s^-1/2 it does not depend on the arbitrary signs /
    rotations LAPACK gives the eigenvectors, so every host, BLAS thread count and rank builds the
    same D to rounding."""
    w, U = np.linalg.eigh(S)
    n = S.shape[0]
    keep = w > lindep * w.max()
    X = (U[:, keep] / np.sqrt(w[keep])) @ U[:, keep].T
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    Q, R = np.linalg.qr(A)
    Q = Q * np.sign(np.diag(R))          # unique QR (positive diagonal of R)
    nocc = min(nocc, n)
    C = X @ Q[:, :nocc]
    D = 2.0 * (C @ C.T)
    return 0.5 * (D + D.T)


# ------------------------------------------------------------------------ grids
def _radial_treutler_m4(n):
    i = np.arange(1, n + 1)
    th = i * np.pi / (n + 1)
    x = np.cos(th)
    ln2 = np.log(2.0)
    r = (1.0 / ln2) * (1.0 + x) ** 0.6 * np.log(2.0 / (1.0 - x))
    drdx = (1.0 / ln2) * (0.6 * (1.0 + x) ** (-0.4) * np.log(2.0 / (1.0 - x)) + (1.0 + x) ** 0.6 / (1.0 - x))
    w = (np.pi / (n + 1)) * np.sin(th) * drdx * r * r  # Chebyshev 2nd kind: int f dx = sum pi/(n+1) sin(th) f
    return r[::-1].copy(), w[::-1].copy()


def _angular_product(n_theta, n_phi):
    ct, wt = np.polynomial.legendre.leggauss(n_theta)
    st = np.sqrt(1.0 - ct * ct)
    phi = (np.arange(n_phi) + 0.5) * (2.0 * np.pi / n_phi)
    ux = np.outer(st, np.cos(phi)).ravel()
    uy = np.outer(st, np.sin(phi)).ravel()
    uz = np.outer(ct, np.ones(n_phi)).ravel()
    w = np.outer(wt, np.full(n_phi, 2.0 * np.pi / n_phi)).ravel()
    return np.stack([ux, uy, uz], axis=1), w


def _becke_partition(coords, atom_xyz, owner):
    """Becke fuzzy-cell weights (3 iterations of the cutoff polynomial, no size adjustment)."""
    natm = atom_xyz.shape[0]
    ng = coords.shape[0]
    dist = np.linalg.norm(coords[:, None, :] - atom_xyz[None, :, :], axis=2)  # (ng,natm)
    P = np.ones((ng, natm))
    for a in range(natm):
        for b in range(a):
            Rab = np.linalg.norm(atom_xyz[a] - atom_xyz[b])
            mu = (dist[:, a] - dist[:, b]) / Rab
            f = mu
            for _ in range(3):
                f = 1.5 * f - 0.5 * f ** 3
            P[:, a] *= 0.5 * (1.0 - f)
            P[:, b] *= 0.5 * (1.0 + f)
    return P[np.arange(ng), owner] / P.sum(axis=1)


def _stockholder_partition(coords, atom_xyz, owner, chunk=65536):
    """O(natm) per point smooth partition of unity (p_A ~ exp(-2 r_A)); used for the
    large benchmark molecules where the O(natm^2) Becke product is too slow on the host."""
    ng = coords.shape[0]
    out = np.empty(ng)
    for s in range(0, ng, chunk):
        c = coords[s:s + chunk]
        dist = np.linalg.norm(c[:, None, :] - atom_xyz[None, :, :], axis=2)
        p = np.exp(-2.0 * (dist - dist.min(axis=1, keepdims=True)))
        out[s:s + chunk] = p[np.arange(c.shape[0]), owner[s:s + chunk]] / p.sum(axis=1)
    return out


def atom_grid_sizes(symbol, scale=1.0):
    """(n_rad, n_theta, n_phi, n_total) for one atom; scale<1 shrinks the grid for CPU tests."""
    total = max(8, int(round(LEVEL3_POINTS[symbol] * scale)))
    n_rad = max(4, int(round(LEVEL3_NRAD[symbol] * scale ** (1.0 / 3.0))))
    n_ang = max(2, total // n_rad)
    n_theta = max(1, int(round(np.sqrt(n_ang / 2.0))))
    n_phi = max(1, n_ang // n_theta)
    return n_rad, n_theta, n_phi, total


def make_grid(mol, scale=1.0, partition="auto"):
    """Synthetic atom-centred grid.  Returns (coords (ngrid,3) Bohr, weights (ngrid,), owner).
    Per-atom point counts equal PySCF level 3 at scale=1 (zero-weight padding points make up
    the difference, as PySCF >= 2 itself pads), so ngrid matches SURVEY.md section 8:
    H2O 34 310, benzene 143 556, DHA 655 136, C33H56N7O17P3S 1 436 406."""
    pts, wts, own = [], [], []
    for ia, sym in enumerate(mol.symbols):
        n_rad, n_theta, n_phi, total = atom_grid_sizes(sym, scale)
        r, wr = _radial_treutler_m4(n_rad)
        u, wa = _angular_product(n_theta, n_phi)
        p = (r[:, None, None] * u[None, :, :]).reshape(-1, 3) + mol.coords[ia]
        w = (wr[:, None] * wa[None, :]).ravel()
        npad = total - p.shape[0]
        if npad < 0:
            p, w = p[:total], w[:total]
        elif npad > 0:
            p = np.vstack([p, np.repeat(p[-1:], npad, axis=0)])
            w = np.concatenate([w, np.zeros(npad)])
        pts.append(p); wts.append(w); own.append(np.full(p.shape[0], ia, dtype=np.int32))
    coords = np.ascontiguousarray(np.vstack(pts))
    w_atomic = np.concatenate(wts)
    owner = np.concatenate(own)
    if partition == "auto":
        partition = "becke" if mol.natm <= 16 else "stockholder"
    if mol.natm == 1:
        part = np.ones_like(w_atomic)
    elif partition == "becke":
        part = np.empty_like(w_atomic)
        for s in range(0, coords.shape[0], 32768):
            part[s:s + 32768] = _becke_partition(coords[s:s + 32768], mol.coords, owner[s:s + 32768])
    else:
        part = _stockholder_partition(coords, mol.coords, owner)
    return coords, w_atomic * part, owner


# Named workloads = BASELINE.json configs (C1..C5) -> (functional, molecule)
WORKLOADS = {
    "C1": ("LDA", "H2O"),
    "C2": ("GGA", "Benzene"),
    "C3": ("B3LYP", "H2O"),
    "C4": ("GGA", "DHA"),
    "C5": ("B3LYP", "C33H56N7O17P3S"),
}


# --------------------------------------------------------------------------- grid files
def load_grid_txt(path):
    """Read a grid in the reference's on-disk format (grid.py:11-14 `init_grid`): whitespace-separated rows
    `atom_index x y z w [w]` -- the owning atom, the point in Bohr, the quadrature weight (the reference's
    files repeat the weight in a sixth column, which `init_grid` ignores).  Returns (coords (n,3) float64,
    weights (n,) float64, atom_index (n,) int32), C-contiguous: coords feed DFT_EvalAO, weights feed
    DFT_ComputeXC, exactly as grid.py hands them to PySCF."""
    data = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if data.size == 0:
        return np.zeros((0, 3)), np.zeros((0,)), np.zeros((0,), dtype=np.int32)
    if data.shape[1] < 5:
        raise ValueError(f"{path}: expected at least 5 columns (atom x y z w), found {data.shape[1]}")
    return (np.ascontiguousarray(data[:, 1:4]), np.ascontiguousarray(data[:, 4]),
            data[:, 0].astype(np.int32))


def save_grid_txt(path, coords, weights, atom_index=None):
    """Write a grid in the same format (weight repeated in the sixth column, 20 significant digits like the
    reference's files), so that grids generated here can drive the reference's `dft.py <functional> <mol>`."""
    coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
    weights = np.asarray(weights, dtype=np.float64).reshape(-1)
    if coords.shape[0] != weights.shape[0]:
        raise ValueError("coords and weights disagree on the number of points")
    atom_index = np.zeros(weights.shape[0], dtype=np.int64) if atom_index is None else np.asarray(atom_index)
    with open(path, "w") as f:
        for a, (x, y, z), w in zip(atom_index, coords, weights):
            f.write(f"{int(a)} {x:.20e} {y:.20e} {z:.20e} {w:.20e} {w:.20e}\n")
