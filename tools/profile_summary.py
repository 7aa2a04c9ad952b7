"""Turn the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/profile_summary.py <tag> <launches.csv> <full.ncu-rep> <kernel-substring>"""
import csv, json, os, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep, kname = sys.argv[1:5]
missions = int(sys.argv[5]) if len(sys.argv) > 5 else 1184
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list ----
rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("==")) if r]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    if len(r) > vi:
        d[r[ki]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
    f.write("# %s: every kernel launch of `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py "
            "--steps 2 --warmup 3 --missions %d --no-cpu-baseline --jacobi-missions 0`\n\n" % (tag, missions))
    f.write("Cold-cache, serialised launch times: compare SHARES, not absolutes.\n\n| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write("| `%s` | %d | %.3f | %.3f | %.4f |\n" % (k[:70], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6, sum(v) / tot))

# ---- full capture of the dominant kernel ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, vals = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"]
got = {}
for i, name in enumerate(h):
    if name in want:
        got[name] = (vals[i], units[i])
def num(x):
    try: return float(x.replace(",", ""))
    except ValueError: return None
def to_bytes(name):
    v, u = got.get(name, ("0", "byte"))
    x = num(v) or 0.0
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
json.dump({"kernel": kname, "missions": missions, "dram_bytes_per_launch": traffic, "metrics": {k: {"value": v, "unit": u} for k, (v, u) in got.items()}},
          open(os.path.join(out_dir, "%s_%s_ncu.json" % (tag, kname)), "w"), indent=1)
with open(os.path.join(out_dir, "%s_%s_ncu.md" % (tag, kname)), "w") as f:
    f.write("# %s: `ncu --set full --clock-control none --import-source on -k regex:%s` (one launch, %d missions x 64 agents)\n\n" % (tag, kname, missions))
    f.write("| metric | value | unit |\n|---|---|---|\n")
    for k in want:
        if k in got:
            f.write("| %s | %s | %s |\n" % (k, got[k][0], got[k][1]))
    f.write("\nDRAM traffic per launch (read + write): %.1f MB\n" % (traffic / 1e6))
print("wrote profiles for", tag, "traffic %.1f MB" % (traffic / 1e6))
