"""Summarises .ncu-rep files (run here, no GPU needed) into profiles/<name>.json.
usage: summarize_ncu.py <name> <report.ncu-rep> [<report2.ncu-rep> ...] [--latest]
Every kernel launch found in the reports becomes one entry of "kernels"; with --latest the summary is also written to
profiles/ncu_summary_latest.json, whose "dram_bytes_per_frame" (sum over the entries: one frame's launches when the reports were
captured that way) is what bench.py reports as roofline.traffic."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
name, reps = args[0], args[1:]


def summarise(hdr, units, vals):
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def num(k):
        try:
            return float(m[k][0].replace(",", ""))
        except Exception:
            return None

    def scaled(k):
        v = num(k); u = m.get(k, ("", ""))[1]
        if v is None:
            return None
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)

    out = {
        "kernel": m.get("Kernel Name", ("?",))[0], "grid": m.get("Grid Size", ("?",))[0], "block": m.get("Block Size", ("?",))[0],
        "duration_ms": num("gpu__time_duration.sum"), "registers_per_thread": num("launch__registers_per_thread"),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "threads_per_warp_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warp_inst": num("smsp__inst_executed.sum"),
        "pipe_fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        # per-cycle-elapsed sums over all SMSPs; x cycles = thread instructions
        "ffma_per_cycle": num("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"),
        "fadd_per_cycle": num("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed"),
        "fmul_per_cycle": num("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed"),
        "sm_cycles_elapsed_max": num("sm__cycles_elapsed.max"),
        "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
        "dram_read_bytes": scaled("dram__bytes_read.sum"), "dram_write_bytes": scaled("dram__bytes_write.sum"),
        "lts_bytes": scaled("lts__t_bytes.sum"),
        "stall_long_scoreboard_per_issue": num("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "stall_no_instruction_per_issue": num("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        "stall_wait_per_issue": num("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        "stall_math_throttle_per_issue": num("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    }
    if out["dram_read_bytes"] is not None and out["dram_write_bytes"] is not None:
        out["dram_bytes_per_launch"] = out["dram_read_bytes"] + out["dram_write_bytes"]
    if out["duration_ms"]:
        per_cycle = 2 * (out["ffma_per_cycle"] or 0) + (out["fadd_per_cycle"] or 0) + (out["fmul_per_cycle"] or 0)
        out["executed_fp32_flop_per_cycle"] = per_cycle   # chip-wide; peak = 148 SM x 128 lanes x 2 = 37 888
        out["executed_fp32_tflops"] = per_cycle * (out["sm_cycles_elapsed_max"] or 0) / (out["duration_ms"] * 1e-3) / 1e12
        out["dram_gbs"] = (out.get("dram_bytes_per_launch") or 0) / (out["duration_ms"] * 1e-3) / 1e9
    return out


kernels = []
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    for vals in rows[2:]:
        if len(vals) == len(rows[0]):
            k = summarise(rows[0], rows[1], vals)
            k["report"] = os.path.basename(rep)
            kernels.append(k)
def _git_head():
    try:
        head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        dirty = subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "sol-r_b200/csrc"], capture_output=True, text=True).stdout.strip()
        return head + (" + uncommitted changes under sol-r_b200/csrc" if dirty else "")
    except Exception:
        return None


import datetime
summary = {"git_head": os.environ.get("SOLR_CAPTURE_HEAD") or _git_head(),
           "captured": datetime.datetime.utcfromtimestamp(os.path.getmtime(reps[0])).strftime("%Y-%m-%dT%H:%MZ") if reps else None,
           "kernels": kernels,
           "duration_ms_sum": sum(k["duration_ms"] or 0 for k in kernels),
           "dram_bytes_per_frame": sum(k.get("dram_bytes_per_launch") or 0 for k in kernels)}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", name + ".json"), "w") as f:
    json.dump(summary, f, indent=1)
if "--latest" in sys.argv:
    with open(os.path.join(ROOT, "profiles", "ncu_summary_latest.json"), "w") as f:
        json.dump(summary, f, indent=1)
for k in kernels:
    print({x: k[x] for x in ("kernel", "duration_ms", "registers_per_thread", "warps_active_pct", "issue_active_pct", "threads_per_warp_inst",
                             "l1_hit_pct", "l2_hit_pct", "dram_bytes_per_launch", "stall_long_scoreboard_per_issue", "executed_fp32_tflops")})
print("sum ms", summary["duration_ms_sum"], "dram bytes", summary["dram_bytes_per_frame"])
