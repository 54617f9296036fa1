"""Per-kernel SASS instruction-count summary of the built library (what proves which hardware path a kernel uses):
DMMA.8x8x4 = FP64 tensor core, UTMALDG = TMA tensor load, UBLKCP = 1-D bulk copy, SYNCS = mbarrier ops,
LDGSTS = cp.async, REDUX = warp reductions into uniform registers, predicated DMMAs (should be 0 in the skipping
kernels: a predicated-off DMMA still pays its issue stall), local-memory traffic (spills).

    python tools/sass_summary.py [path/to/dft.so] > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["DMMA", "@P DMMA", "DFMA", "UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "LDL", "STL",
        "BAR", "REDUX", "SHFL", "VOTE", "BRA", "MUFU", "ATOMG", "RED"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "quantum_compute_dft_b200", "weights", "dft.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "")
            name = re.sub(r"\(.*", "", name)
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(2).split(".")[0]
            c = counts[name]
            c["total"] += 1
            if op in KEYS:
                c[op] += 1
            if op == "DMMA" and m.group(1):
                c["@P DMMA"] += 1
    print(f"# {os.path.relpath(so, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':<72} {'total':>6} " + " ".join(f"{k:>7}" for k in KEYS))
    for n, c in counts.items():
        if c["total"] == 0:
            continue
        short = n.replace("xc::tmapath::", "").replace("xc::smallpath::", "small::").replace("xc::generic::", "generic::").replace("xc::", "")
        print(f"{short[:72]:<72} {c['total']:>6} " + " ".join(f"{c[k]:>7}" for k in KEYS))


if __name__ == "__main__":
    main()
