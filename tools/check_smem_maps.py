#!/usr/bin/env python3
"""Host-side check of the shared-memory index maps used by the TMA kernels (no GPU needed).

A 64-bit warp load is served in two phases of 16 lanes; a phase is conflict-free when its 16
8-byte words fall into 16 distinct 8-byte bank pairs (address / 8 mod 16).  TMA writes boxes of
16 doubles per row with the 128-byte swizzle: 16-byte chunk index XOR (row mod 8).
"""
import itertools

RHO = [0, 3, 4, 7, 1, 2, 5, 6]            # A-side fragment row -> tile row (density kernel)
PERM = [2 * (j & 3) + (j >> 2) for j in range(8)]  # B-side fragment index -> D-tile row / C column


def swz(row, col):
    return row * 128 + ((((col >> 1) ^ row) & 7) << 4) + ((col & 1) << 3)


def conflict_free(addrs):
    """addrs: 32 byte addresses (lane order)."""
    for ph in range(2):
        banks = [(a // 8) % 16 for a in addrs[16 * ph:16 * ph + 16]]
        if len(set(banks)) != 16:
            return False
    return True


def check_density():
    ok = True
    for ks in range(4):
        # A fragment: lane (q, qcol) reads tile row RHO[q] (+8 mf), k column 4 ks + qcol
        a = [swz(RHO[l >> 2], 4 * ks + (l & 3)) for l in range(32)]
        b = [swz(PERM[l >> 2], 4 * ks + (l & 3)) for l in range(32)]
        ok &= conflict_free(a) and conflict_free(b)
    # epilogue piece: lane reads row RHO[q] (+8 mf), column 8 s + PERM[2 qcol + e]
    for s, e in itertools.product(range(2), range(2)):
        ad = [swz(RHO[l >> 2], 8 * s + PERM[2 * (l & 3) + e]) for l in range(32)]
        ok &= conflict_free(ad)
    # every (row, col) of an 8x8 accumulator fragment is covered exactly once
    cover = {(RHO[l >> 2], PERM[2 * (l & 3) + e]) for l in range(32) for e in range(2)}
    ok &= len(cover) == 64
    return ok


def krow(ks, qcol, vk):
    return (8 * (ks >> 1) + 2 * qcol + (ks & 1)) if vk == 16 else (2 * qcol + (ks & 1))


def check_vxc(vk):
    ok = True
    rows = set()
    for ks in range(vk // 4):
        for half in range(2):   # column group parity inside a 16-column box
            ad = [swz(krow(ks, l & 3, vk), 8 * half + (l >> 2)) for l in range(32)]
            ok &= conflict_free(ad)
        rows |= {krow(ks, qc, vk) for qc in range(4)}
    ok &= rows == set(range(vk))
    return ok


if __name__ == "__main__":
    print("density maps conflict-free:", check_density())
    print("vxc maps conflict-free (VK=16):", check_vxc(16), "(VK=8):", check_vxc(8))
