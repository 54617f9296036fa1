#!/usr/bin/env python3
"""Where does the density kernel's time go?  Runs one workload through the DFT_PHASE_TIMING build
(python -m quantum_compute_dft_b200.build --timing) and prints per-phase cycle shares.
Usage: python tools/phase_timing.py C5 [KEY=VALUE ...]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
from quantum_compute_dft_b200 import workload
from quantum_compute_dft_b200.solver import DFTSolverWrapper, DEFAULT_LIB

lib = DEFAULT_LIB.replace("dft.so", "dft_timing.so")
wl = sys.argv[1]
hp = workload.host_problem(wl)
s = DFTSolverWrapper(lib, hp.functional)
s.set_option("timing", 1)
for kv in sys.argv[2:]:
    k, v = kv.split("="); s.set_option(k, float(v))
dp = workload.device_problem(hp, s)
for _ in range(3):
    e = s.compute_xc(dp.ngrid, dp.nao, dp.d_dm, dp.d_ao, dp.d_weights, dp.d_vxc, dp.d_ao_grad)
print(wl, "E_xc", e, "density_ms", s.stat("density_ms"), "vxc_ms", s.stat("vxc_ms"))
s.lib.DFT_DebugRead.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64]
nblk = min(148, (dp.ngrid + 127) // 128 + 1)
nv = int(os.environ.get("V_CTAS", "144"))
raw = np.zeros(65536 + 160 * 256, dtype=np.int64)
rc = s.lib.DFT_DebugRead(s.solver, b"scratch", raw.ctypes.data_as(ctypes.c_void_p), raw.nbytes)
assert rc == 0, rc
buf = raw[:nblk * 32].reshape(nblk, 8, 4)
vb = raw[8192:8192 + nv * 36].reshape(nv, 9, 4).astype(float)
# density kernel: the first 64 non-k-loop intervals (epilogue + block tail) of each consumer group
iv = raw[65536:65536 + nblk * 256].reshape(nblk, 2, 64, 2)
ov_tot = ep_tot = 0.0
for c in range(nblk):
    a, b = iv[c, 0], iv[c, 1]
    a = a[(a[:, 0] > 0) & (a[:, 1] > a[:, 0])]; b = b[(b[:, 0] > 0) & (b[:, 1] > b[:, 0])]
    if len(a) == 0 or len(b) == 0:
        continue
    t_lo, t_hi = max(a[0, 0], b[0, 0]), min(a[-1, 1], b[-1, 1])
    for s0, s1 in a:
        ep_tot += max(0, min(s1, t_hi) - max(s0, t_lo))
        for r0, r1 in b:
            ov_tot += max(0, min(s1, r1, t_hi) - max(s0, r0, t_lo))
if ep_tot > 0:
    a0 = iv[0, 0]; b0 = iv[0, 1]
    print("density kernel ping-pong: %.1f %% of group 0's epilogue time overlaps group 1's epilogues" % (100 * ov_tot / ep_tot))
    print("  CTA 0 group 0 epilogue starts (cycles, rel):", (a0[:6, 0] - a0[0, 0]).tolist(), "lengths", (a0[:6, 1] - a0[:6, 0]).tolist())
    print("  CTA 0 group 1 epilogue starts (cycles, rel):", (b0[:6, 0] - a0[0, 0]).tolist(), "lengths", (b0[:6, 1] - b0[:6, 0]).tolist())

tot = buf.sum(axis=2).astype(float)
print("cycles per warp: mean %.3e  min %.3e  max %.3e  (%.2f ms at 1.965 GHz)" % (tot.mean(), tot.min(), tot.max(), tot.mean() / 1.965e6))
for i, name in enumerate(("k-loop", "piece wait", "piece math+release", "block tail")):
    x = buf[:, :, i].astype(float)
    print(f"  {name:20s} mean {x.mean():.3e} ({100 * x.mean() / tot.mean():5.1f} %)   min {x.min():.3e} max {x.max():.3e}")
print("per-warp share of 'piece math' (CTA 0):", (buf[0, :, 2] / tot[0]).round(3))

print("V kernel, consumer warps: total %.3e cycles (%.2f ms)" % (vb[:, :8, 3].mean(), vb[:, :8, 3].mean() / 1.965e6))
for i, name in enumerate(("wait full", "fragments+DMMA", "store tile")):
    x = vb[:, :8, i]
    print(f"  {name:20s} mean {x.mean():.3e} ({100 * x.mean() / vb[:, :8, 3].mean():5.1f} %)   min {x.min():.3e} max {x.max():.3e}")
x = vb[:, 8, :]
print("V kernel, producer: wait empty %.1f %%, issue %.1f %% of %.3e cycles" % (100 * x[:, 0].mean() / x[:, 3].mean(), 100 * x[:, 1].mean() / x[:, 3].mean(), x[:, 3].mean()))
print("per-CTA consumer total min/max: %.3e %.3e" % (vb[:, :8, 3].mean(axis=1).min(), vb[:, :8, 3].mean(axis=1).max()))
# per-warp wait share by tile row (the warp that waits least is the CTA's critical path)
ntile = int(os.environ.get("V_TILES", "9"))
tn = int(round(ntile ** 0.5))
w = vb[:, :8, 0] / np.maximum(vb[:, :8, 3], 1.0)
for tm in range(tn):
    sel = [c for c in range(nv) if (c % ntile) // tn == tm]
    print(f"V kernel, tile row {tm}: per-warp 'wait full' share", w[sel].mean(axis=0).round(3), " least-waiting warp per CTA: mean %.3f" % w[sel].min(axis=1).mean())
