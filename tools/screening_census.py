#!/usr/bin/env python3
"""CPU census of the AO zero pattern at the granularities the TMA kernels skip at (no GPU needed).

    python tools/screening_census.py C5 [nsamples]

Samples random chunks of the workload's grid, evaluates the AOs with the numpy statement of the evaluator
(molgrid.eval_ao_numpy, same cutoff as DFT_EvalAO) and reports, in sub-problem rows (every second grid point
for odd nao):
  * density kernel: share of live 32 x 4 Phi fragments (what its k-step votes see), of k-steps live in either
    32-row half of a 64-row block, and of 64 x 16 chunks with any non-zero;
  * V kernel (8- and 16-row ring stages): share of live 16-column boxes on the M side (any of the four planes)
    and on the N side (Phi), of live (M box, N box) pairs, of live 8-column x 4-row M fragments, and -- per
    128 x 128 output tile -- of stages in which both tiles have something, the mean work when every warp skips
    its own dead M box and all skip the dead N boxes, and the same when the eight warps move in lockstep.
The numbers quoted in DESIGN.md 5.2b / 5.2c come from this script."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_compute_dft_b200 import molgrid, workload  # noqa: E402


def planes(hp, rows):
    ao = molgrid.eval_ao_numpy(hp.coords[rows], hp.basis, deriv=1)
    return ao[0], np.asarray(ao[1])


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
    nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    hp = workload.host_problem(wl)
    nao = hp.nao
    step = 2 if nao % 2 else 1
    ncg = (nao + 15) // 16
    rng = np.random.default_rng(1)
    print(f"{wl} {hp.name}: ngrid {hp.ngrid}, nao {nao}, {ncg} 16-column groups, row stride {step}")

    # ---- density kernel: 64-row blocks
    nblk = hp.ngrid // (step * 64)
    frag = np.zeros((nsamp, 2, ncg, 4), bool)
    for n, b in enumerate(rng.choice(nblk, nsamp, replace=False)):
        rows = step * (b * 64 + np.arange(64))
        phi, _ = planes(hp, rows)
        pad = np.zeros((64, ncg * 16), bool)
        pad[:, :nao] = phi != 0
        frag[n] = pad.reshape(2, 32, ncg, 4, 4).any(axis=(1, 4))
    print(f"density: live 32x4 fragments {frag.mean():.3f}; k-step live in either half {frag.any(axis=1).mean():.3f}; "
          f"live 64x16 chunks {frag.any(axis=(1, 3)).mean():.3f}")

    # ---- V kernel: VK-row stages
    for vk in (8, 16):
        nch = hp.ngrid // (step * vk)
        m16 = np.zeros((nsamp * 2, ncg), bool)
        n16 = np.zeros((nsamp * 2, ncg), bool)
        m8x4 = []
        for n, c in enumerate(rng.choice(nch, nsamp * 2, replace=False)):
            rows = step * (c * vk + np.arange(vk))
            phi, g = planes(hp, rows)
            any4 = np.zeros((vk, ncg * 16), bool)
            any4[:, :nao] = (phi != 0) | (g != 0).any(axis=0)
            p0 = np.zeros((vk, ncg * 16), bool)
            p0[:, :nao] = phi != 0
            m16[n] = any4.reshape(vk, ncg, 16).any(axis=(0, 2))
            n16[n] = p0.reshape(vk, ncg, 16).any(axis=(0, 2))
            m8x4.append(any4.reshape(vk // 4, 4, ncg * 2, 8).any(axis=(1, 3)).mean())
        print(f"V, {vk}-row stages: live M boxes {m16.mean():.3f}, live N boxes {n16.mean():.3f}, live (M, N) box pairs "
              f"{(m16[:, :, None] & n16[:, None, :]).mean():.3f}, live 8x4 M fragments {np.mean(m8x4):.3f}")
        g = 8
        nt = (ncg + g - 1) // g
        live = mean = lock = 0.0
        for ti in range(nt):
            for tj in range(nt):
                mi, nj = m16[:, g * ti:g * ti + g], n16[:, g * tj:g * tj + g]
                live += (mi.any(axis=1) & nj.any(axis=1)).mean()
                w = mi * (nj.sum(axis=1) / g)[:, None]
                mean += w.mean(axis=1).mean()
                lock += w.max(axis=1).mean()
        print(f"   128 x 128 tiles: stages with both tiles live {live / nt**2:.3f}; work with per-warp M skip + common N skip "
              f"{mean / nt**2:.3f}; same in lockstep {lock / nt**2:.3f}")


if __name__ == "__main__":
    main()
