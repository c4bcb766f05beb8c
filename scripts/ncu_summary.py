"""Summarise an .ncu-rep (raw + source pages) -> JSON + text.  usage: ncu_summary.py rep out_prefix [games_in_profiled_launch]
(bench.py scales `dram_bytes_per_launch` by `games_in_profiled_launch` for roofline.traffic)"""
import csv, json, subprocess, sys, io, re
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = [k for k in d if re.search(r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__throughput.avg.pct|smsp__issue_active.avg.pct|smsp__inst_executed.sum$|sm__warps_active.avg.pct|launch__registers|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|thread_inst_executed_per_inst|launch__grid_size|launch__block_size|shared_mem_per_block_dynamic|gpu__dram_throughput.avg.pct|smsp__average_warps_issue_stalled.*per_issue_active|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|lts__t_bytes.sum$|l1tex__t_bytes.sum$|smsp__inst_executed_pipe_.*sum$|sm__inst_executed_pipe_.*sum$", k)]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
tob = lambda k: float(d[k][0]) * scale.get(d[k][1], 1)
summ = {"report": rep, "games_in_profiled_launch": int(sys.argv[3]) if len(sys.argv) > 3 else 500000, "dram_bytes_per_launch": tob("dram__bytes_read.sum") + tob("dram__bytes_write.sum"),
        "metrics": {k: {"value": d[k][0], "unit": d[k][1]} for k in sorted(keys)}}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
base = int(data[0][ia], 16)
segs = []; cur = None
for r in data:
    ex = int(r[iex]); a = int(r[ia], 16) - base; s = int(r[ismp])
    if cur and abs(ex - cur["ex"]) <= 0.02 * max(ex, cur["ex"], 1):
        cur["n"] += 1; cur["sum"] += ex; cur["smp"] += s; cur["end"] = a
    else:
        cur = dict(start=a, end=a, ex=ex, n=1, sum=ex, smp=s); segs.append(cur)
segs.sort(key=lambda s: -s["sum"])
summ["total_warp_instructions"] = tot
summ["hot_segments"] = [dict(start=hex(s["start"]), end=hex(s["end"]), instructions=s["n"], executions_per_instruction=s["ex"],
                             share_of_instructions=round(s["sum"] / tot, 4), share_of_samples=round(s["smp"] / max(tots, 1), 4)) for s in segs[:12]]
json.dump(summ, open(out + ".json", "w"), indent=1)
for k in sorted(keys): print(k, d[k])
for s in summ["hot_segments"]: print(s)
