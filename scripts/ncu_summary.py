"""Summarise an ncu report (raw + source pages) into profiles/<tag>.md and <tag>_summary.json."""
import csv
import io
import json
import subprocess
import sys

rep, tag, steps_per_launch = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def g(name, default=None):
    try:
        return float(m[name][0].replace(",", ""))
    except Exception:
        return default


keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
unit = {k: m[k][1] for k in keys if k in m}
out = {k: g(k) for k in keys if k in m}
dur_ms = out["gpu__time_duration.sum"] * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit["gpu__time_duration.sum"], 1.0)


def to_bytes(k):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit.get(k, "byte"), 1)
    return (out.get(k) or 0.0) * f


dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
inst = out.get("smsp__inst_executed.sum") or 0.0
stalls = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(v.replace(",", ""))
          for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v}
summary = {
    "kernel": m.get("Kernel Name", ("", ""))[0], "duration_ms": dur_ms, "attempted_steps_per_launch": steps_per_launch,
    "steps_per_s_under_ncu": steps_per_launch / (dur_ms * 1e-3),
    "dram_bytes_per_launch": dram, "dram_bytes_per_attempted_step": dram / steps_per_launch,
    "l2_bytes_per_attempted_step": to_bytes("lts__t_bytes.sum") / steps_per_launch,
    "warp_instructions_per_attempted_step": inst / steps_per_launch,
    "ipc_per_sm": out.get("sm__inst_executed.avg.per_cycle_elapsed"),
    "registers_per_thread": out.get("launch__registers_per_thread"),
    "grid": out.get("launch__grid_size"), "block": out.get("launch__block_size"),
    "waves_per_sm": out.get("launch__waves_per_multiprocessor"),
    "dyn_smem_per_block": to_bytes("launch__shared_mem_per_block_dynamic"),
    "warps_active_pct": out.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "l1_hit_pct": out.get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": out.get("lts__t_sector_hit_rate.pct"),
    "dram_throughput_pct": out.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "sm_throughput_pct": out.get("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    "tensor_pipe_pct": out.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    "smem_bank_conflicts": out.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "top_stalls_warps_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8]),
}
json.dump(summary, open(f"profiles/{tag}_summary.json", "w"), indent=1)
with open(f"profiles/{tag}.md", "w") as f:
    f.write(f"# ncu --set full summary: {tag}\n\nreport: `{rep}` (scratch, not committed)\n\n")
    for k, v in summary.items():
        f.write(f"* **{k}**: {v}\n")
print(json.dumps(summary, indent=1))
