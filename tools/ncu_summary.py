#!/usr/bin/env python3
"""Condense an `ncu --set full` report into the few numbers the roofline needs, per kernel launch.

    python tools/ncu_summary.py gpurun_out/s19/prof_C5.ncu-rep C5 profiles/r1_ncu_summary.json

Writes/updates a JSON keyed by workload -> kernel family -> metrics (DRAM bytes per launch, FP64 tensor
pipe utilisation, duration under ncu, L2 hit rate, top stall reasons).  bench.py reads the JSON for the
`roofline.traffic` field; the numbers are from a profiled run and are never used as timings."""
import csv, io, json, os, subprocess, sys

rep, workload, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}

def num(r, key):
    try:
        return float(r[ix[key]].replace(",", ""))
    except Exception:
        return None

def to_bytes(v, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return None if v is None else v * scale.get(unit, 1.0)

fam = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = ("density" if "density" in name else "vxc" if "vxc_tma" in name else "point" if "xc_point" in name
           else "ao_eval" if "eval_kernel" in name else name[:40])
    rd = to_bytes(num(r, "dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(num(r, "dram__bytes_write.sum"), units[ix["dram__bytes_write.sum"]])
    stalls = {}
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
            v = num(r, h)
            if v and v > 0.2:
                stalls[h.split("issue_stalled_")[1].split("_per")[0]] = round(v, 2)
    t = num(r, "gpu__time_duration.sum")
    tu = units[ix["gpu__time_duration.sum"]]
    t_ms = None if t is None else t * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu, 1.0)
    fam[key] = {
        "kernel": name,
        "duration_ms_under_ncu": t_ms,
        "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": (rd or 0) + (wr or 0),
        "fp64_tensor_pipe_pct": num(r, "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_hit_rate_pct": num(r, "lts__t_sector_hit_rate.pct"),
        "registers_per_thread": num(r, "launch__registers_per_thread"),
        "warp_stalls_per_issue": stalls,
        "source": os.path.basename(rep),
    }
db = {}
if os.path.exists(out):
    db = json.load(open(out))
db[workload] = fam
json.dump(db, open(out, "w"), indent=1, sort_keys=True)
for k, v in fam.items():
    print(workload, k, {a: b for a, b in v.items() if a not in ("kernel", "source")})
