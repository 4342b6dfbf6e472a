"""Per-source-line stall samples / instruction shares and headline metrics of one kernel from an .ncu-rep
(ncu --set full --import-source on).  usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__average_warp_latency_per_inst_issued.ratio", "launch__grid_size", "launch__block_size"]
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if h in want:
        print("%-70s %-12s %s" % (h, u, v))
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if "issue_stalled" in h and "per_issue_active" in h and float(v or 0) > 0.2:
        print("%-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = []; ts = ti = 0
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] not in ("", "Line No") and r[2] == "-":
        try: s = int(r[4]); ins = int(r[7])
        except ValueError: continue
        agg.append((cur, int(r[0]), r[1].strip()[:100], s, ins)); ts += s; ti += ins
agg.sort(key=lambda x: -x[3])
print("-- top source lines by stall samples (share of samples, share of instructions) --")
for a in agg[:top]:
    print("%-14s %5d %6.2f%% %6.2f%%  %s" % (a[0], a[1], 100 * a[3] / max(ts, 1), 100 * a[4] / max(ti, 1), a[2]))
