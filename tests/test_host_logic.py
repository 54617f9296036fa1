"""CPU checks of host-side logic that the CUDA kernels rely on (no GPU needed)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shared_memory_maps_are_bank_conflict_free():
    """The fragment-row permutations of the TMA kernels: every 64-bit fragment load of
    the k-loops, the density epilogue and the V kernel hits 16 distinct bank pairs per phase under the
    128-byte TMA swizzle, and the permutations cover every accumulator element exactly once."""
    m = _load(os.path.join(ROOT, "tools", "check_smem_maps.py"), "check_smem_maps")
    assert m.check_density()
    assert m.check_vxc(16)
    assert m.check_vxc(8)


def test_transpose_reduce_schedule():
    """The 3-shuffle transpose-reduce of the density kernel (4 lanes x 4 planes -> lane q holds plane q)."""
    import numpy as np
    rng = np.random.default_rng(0)
    v = rng.standard_normal((4, 4))          # v[lane][plane]
    s0 = np.zeros(4); s1 = np.zeros(4)
    for lane in range(4):
        b0 = lane & 1
        k0, g0 = (v[lane][1], v[lane][0]) if b0 else (v[lane][0], v[lane][1])
        k1, g1 = (v[lane][3], v[lane][2]) if b0 else (v[lane][2], v[lane][3])
        s0[lane], s1[lane] = k0, k1
        v[lane][0], v[lane][1] = g0, g1      # what this lane hands to lane ^ 1
    give = v[:, :2].copy()
    for lane in range(4):
        s0[lane] += give[lane ^ 1][0]
        s1[lane] += give[lane ^ 1][1]
    tot = np.zeros(4)
    for lane in range(4):
        b1 = (lane >> 1) & 1
        k = s1[lane] if b1 else s0[lane]
        g_partner = s0[lane ^ 2] if ((lane ^ 2) >> 1) & 1 else s1[lane ^ 2]
        tot[lane] = k + g_partner
    rng = np.random.default_rng(0)
    ref = rng.standard_normal((4, 4)).sum(axis=0)
    np.testing.assert_allclose(tot, ref, rtol=1e-14)
