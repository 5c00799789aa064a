"""Digest an `ncu --set full` report of the step kernels into the figures bench.py quotes.

    python tools/ncu_digest.py REPORT.ncu-rep PARTICLES CONFIG_KEY [OUT.json]

Writes / updates profiles/r02_kernel_metrics.json[CONFIG_KEY] with, per kernel of one internal step
(k_advect, k_vbuild, k_vwalk, k_finish; the first captured launch of each): duration, registers,
warp instructions, active lanes, issue and FP64-pipe utilisation, dram bytes, FP64 thread-instruction
counts; and the per-particle-step sums bench.py uses:
    dram_bytes_per_particle_step = sum(dram__bytes_read.sum + dram__bytes_write.sum) / PARTICLES
    fp64_flop_per_particle_step  = sum(2 x dfma + dmul + dadd thread instructions) / PARTICLES
Also prints a CSV summary (committed next to the json as profiles/<name>_summary.csv by the caller)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, particles, key = sys.argv[1], float(sys.argv[2]), sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=0.0):
    if name not in col:
        return default
    try:
        v = float(r[col[name]].replace(",", ""))
    except ValueError:
        return default
    u = units[col[name]].lower()
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9,
             "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * scale


kernels = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("<")[0].replace("void ", "").strip()
    if short in kernels or not short.startswith("k_"):
        continue
    # the raw page carries these as instructions per elapsed cycle summed over the sub-partitions
    cyc = val(r, "smsp__cycles_elapsed.avg")
    dfma = val(r, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cyc
    dmul = val(r, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cyc
    dadd = val(r, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cyc
    kernels[short] = {
        "duration_ms": 1e3 * val(r, "gpu__time_duration.sum"),
        "registers": val(r, "launch__registers_per_thread"),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "active_lanes_per_instruction": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_active_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
        "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
        "dfma": dfma, "dmul": dmul, "dadd": dadd, "fp64_flop": 2 * dfma + dmul + dadd,
        "stall_long_scoreboard": val(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "stall_wait": val(r, "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        "stall_short_scoreboard": val(r, "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    }
step = [k for k in ("k_advect", "k_vbuild", "k_vwalk", "k_vturb", "k_finish") if k in kernels]
entry = {
    "report": os.path.basename(rep), "particles": particles, "kernels": kernels,
    "dram_bytes_per_particle_step": sum(kernels[k]["dram_read_bytes"] + kernels[k]["dram_write_bytes"] for k in step) / particles,
    "fp64_flop_per_particle_step": sum(kernels[k]["fp64_flop"] for k in step) / particles,
    "step_kernels": step,
}
data = {}
if os.path.exists(out):
    data = json.load(open(out))
data[key] = entry
json.dump(data, open(out, "w"), indent=1)
w = csv.writer(sys.stdout)
fields = list(next(iter(kernels.values())).keys())
w.writerow(["kernel"] + fields)
for k, v in kernels.items():
    w.writerow([k] + [("%.6g" % v[f]) for f in fields])
print("# %s: dram bytes / particle-step %.1f, FP64 flop / particle-step %.0f" % (key, entry["dram_bytes_per_particle_step"], entry["fp64_flop_per_particle_step"]))
