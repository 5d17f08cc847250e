"""Turns ncu output brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/ncu_summary.py full  gpurun_out/prof.ncu-rep  profiles/<tag>_ncu_full.csv [--traffic profiles/traffic.json]
  python scripts/ncu_summary.py list  gpurun_out/launches.csv  profiles/<tag>_launch_shares.csv

`full`: one row per captured launch with duration, DRAM bytes read/written, DRAM / SM / FP64-pipe utilisation, registers,
achieved occupancy and cache hit rates (from `ncu --set full`).  With --traffic the per-launch DRAM traffic
(read + write, averaged per kernel) is merged into traffic.json under the kernel-class names bench.py uses.
`list`: aggregates a `--metrics gpu__time_duration.sum` launch list into time share per kernel (cold-cache, serialised).
"""
import csv
import json
import subprocess
import sys

FULL = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]
CLASS = {"assemble_kernel": "assembly", "volvars_kernel": "volvars", "bcrs_spmv_kernel": "spmv", "stencil_spmv": "spmv",
         "ilu0_lower": "ilu0_lower", "ilu0_upper": "ilu0_upper", "ilu_sweep": "ilu0_apply"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def full(rep, out, traffic_path=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), name) for m, name in FULL if m in hdr]
    per_class = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([name + ("_ms" if name == "time" else "_bytes" if name in ("dram_read", "dram_write") else "") for _, name in cols])
        for r in data:
            line = []
            rec = {}
            for i, name in cols:
                v = r[i]
                if name == "kernel":
                    v = v.split("(")[0].replace("void ", "")
                elif units[i] in SCALE:
                    v = float(v.replace(",", "")) * SCALE[units[i]]
                rec[name] = v
                line.append(f"{v:.6g}" if isinstance(v, float) else v)
            w.writerow(line)
            for key, cls in CLASS.items():
                if key in rec["kernel"]:
                    per_class.setdefault(cls, []).append(rec["dram_read"] + rec["dram_write"])
    if traffic_path:
        try:
            t = json.load(open(traffic_path))
        except Exception:
            t = {}
        for cls, v in per_class.items():
            t[cls] = sum(v) / len(v)
        if "ilu0_lower" in per_class and "ilu0_upper" in per_class:
            t["ilu0_apply"] = t["ilu0_lower"] + t["ilu0_upper"]
        if "assembly" in per_class and "volvars" in per_class:
            t["assembly_total"] = t["assembly"] + t["volvars"]
        t["_source"] = rep.split("/")[-1]
        json.dump(t, open(traffic_path, "w"), indent=1, sort_keys=True)


def launch_list(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "")
        ms = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "avg_ms", "share"])
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, f"{ms:.4f}", f"{ms / n:.4f}", f"{ms / total:.4f}"])


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "full":
        full(src, dst, sys.argv[5] if len(sys.argv) > 5 and sys.argv[4] == "--traffic" else None)
    else:
        launch_list(src, dst)
