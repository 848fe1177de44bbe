"""Text summary of an .ncu-rep for profiles/: key metrics, stall reasons, time by function.
usage: python scripts/ncu_summary.py <report.ncu-rep> [--so <libb200poa.so of that run>] > profiles/<name>.txt"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]
so = sys.argv[sys.argv.index("--so") + 1] if "--so" in sys.argv else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, unit, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
def g(k):
    return (vals[col[k]], unit[col[k]]) if k in col else ("n/a", "")
print("report:", os.path.basename(rep))
for k in ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_tensor.sum"]:
    v, u = g(k)
    print(f"{k:70s} {v} {u}")
st = sorted(((float(vals[i]), h) for i, h in enumerate(hdr)
             if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")), reverse=True)
print("--- warps stalled per issue, by reason")
for v, h in st[:8]:
    print(f"{v:8.3f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
if so:
    env = dict(os.environ, B200POA_SO=os.path.abspath(so))
    res = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_lines.py"), rep, "--top", "0"],
                         capture_output=True, text=True, env=env).stdout
    print(res.split("--- by line")[0])
