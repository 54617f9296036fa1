"""PySCF-numint-shaped CPU baseline (OUR restatement -- PySCF itself is not installable here).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/xc_oracle.c).  This is the shape of what the
reference's comparison run does on the CPU (dft.py:281-291 -> PySCF RKS -> numint.nr_rks):
a blocked loop over grid points with BLAS dgemm for c0 = Phi.D and V += Phi^T.aow, row-dots
for rho / grad rho, the pointwise functional (here: the oracle's C routines, OpenMP), and the
final V + V^T.  It uses all host cores numpy's BLAS and OpenMP give it.  Output convention:
the symmetric matrix 1/2 (V_ref + V_ref^T) that the reference's driver forms at dft.py:212.
"""
import numpy as np

from . import oracle as O

BLOCK = 8192  # grid points per block (PySCF's BLKSIZE-scale blocking)


def nr_rks(xc_type, dm, ao, weights, ao_grad=None, mode=O.COMPAT, block=BLOCK):
    ngrid, nao = ao.shape
    dsym = 0.5 * (dm + dm.T)
    V = np.zeros((nao, nao))
    exc_total = 0.0
    for s in range(0, ngrid, block):
        e = min(ngrid, s + block)
        phi = ao[s:e]
        w = weights[s:e]
        c0 = phi @ dsym                                   # dgemm
        rho = np.einsum("gi,gi->g", c0, phi)
        if xc_type == O.LDA:
            exc, vrho, _ = O.functional_points(xc_type, rho, None, mode=mode, gate=True)
            gate = rho >= 1e-12
            aow = (0.5 * w * vrho * gate)[:, None] * phi
        else:
            gphi = ao_grad[:, s:e]
            grad = 2.0 * np.einsum("gi,cgi->cg", c0, gphi)
            sigma = np.einsum("cg,cg->g", grad, grad)
            exc, vrho, vsig = O.functional_points(xc_type, rho, sigma, mode=mode, gate=True)
            gate = rho >= 1e-12
            aow = (0.5 * w * vrho * gate)[:, None] * phi
            wv = 2.0 * w * vsig * gate * grad             # (3, g)
            aow += np.einsum("cg,cgi->gi", wv, gphi)
        exc_total += float(np.dot(w, exc))
        V += aow.T @ phi                                  # dgemm
    return exc_total, V + V.T
