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
CLASS = {"assemble_tile_kernel": "assembly", "tracer_assemble_kernel": "tracer_assembly", "bcrs_spmv_kernel": "spmv_bcrs",
         "stencil_spmv": "spmv", "ilu0_lower": "ilu0_lower", "ilu0_upper": "ilu0_upper", "ilu_sweep_kernel<2, 0>": "ilu_sweep_lower",
         "ilu_sweep_kernel<2, 1>": "ilu_sweep_upper", "vec_skew": "ilu_vec_skew", "vec_unskew": "ilu_vec_unskew",
         "axpy3_norm_dot": "blas1_axpy3_norm_dot", "axpy_r_norm": "blas1_axpy_r_norm", "p_update": "blas1_p_update", "dot_kernel": "blas1_dot"}
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
        parts = ("ilu_vec_skew", "ilu_sweep_lower", "ilu_sweep_upper")
        if all(p in per_class for p in parts):
            # one preconditioner application = vec_skew + lower sweep + upper sweep (+ vec_unskew in builds that have it)
            t["ilu0_apply"] = sum(t[p] for p in parts) + (t["ilu_vec_unskew"] if "ilu_vec_unskew" in per_class else 0.0)
        for stale in ("volvars", "assembly_total", "ilu0_lower", "ilu0_upper"):
            t.pop(stale, None)
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


def projected_step(list_csv, out, iterations):
    """Per-kernel share of ONE Newton step of the bench workload: ncu per-launch averages (cold cache, serialised) times the
    number of launches the step makes with `iterations` BiCGSTAB iterations (bicgstab() in csrc/linalg.cu)."""
    rows = list(csv.DictReader(open(list_csv)))
    avg = {r["kernel"]: float(r["avg_ms"]) for r in rows}
    it = iterations
    per_step = {"assemble_tile_kernel": 1, "ilu0_factor_kernel": 1, "ilu_diag_kernel": 1, "ilu_skew_kernel": 1, "vec_skew_kernel": 2 * it,
                "ilu_sweep_kernel<2, 0>": 2 * it, "ilu_sweep_kernel<2, 1>": 2 * it, "stencil_spmv_kernel": 2 * it + 1,
                "dot_kernel<1>": it + 1, "dot_kernel<2>": it, "p_update_kernel": it - 1, "axpy_r_norm_kernel": it,
                "axpy3_norm_dot_kernel": it, "residual_init_kernel": 1, "final_reduce_kernel": 4 * it + 2, "newton_update_kernel": 1}
    out_rows = []
    for name, ms in avg.items():
        n = next((c for key, c in per_step.items() if name.startswith(key) or key in name), 0)
        if n:
            out_rows.append((name, n, ms, n * ms))
    total = sum(r[3] for r in out_rows)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches_per_step", "ncu_avg_ms", "projected_ms_per_step", "share"])
        for name, n, ms, tot in sorted(out_rows, key=lambda r: -r[3]):
            w.writerow([name, n, f"{ms:.4f}", f"{tot:.2f}", f"{tot / total:.4f}"])
        w.writerow(["TOTAL", "", "", f"{total:.2f}", "1.0"])


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "step":
        projected_step(src, dst, int(sys.argv[4]))
    elif mode == "full":
        full(src, dst, sys.argv[5] if len(sys.argv) > 5 and sys.argv[4] == "--traffic" else None)
    else:
        launch_list(src, dst)
