"""Summarise an .ncu-rep (raw + source pages) -> JSON + text.

    python scripts/ncu_summary.py rep out_prefix [games_in_profiled_launch] [launch_index] [warp_steps]

`warp_steps`: FMC_C_WARP_STEPS / fmc_predict_stats of the same launch run without ncu (scripts/quick_bench.py prints
it): warp-level node gathers of the tree walk; with it the summary carries the kernel's L1 wavefronts per gather level.

bench.py reads the JSON: `dram_bytes_per_launch` / `games_in_profiled_launch` scale `roofline.traffic`, and
`derived.lgds_wavefronts_per_global_ld_request` (L1 data-pipe wavefronts one warp-level node gather costs) turns the
kernel's own warp-step counter into L1 wavefronts for the live roofline."""
import csv, json, subprocess, sys, io, re
rep, out = sys.argv[1], sys.argv[2]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + launch]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
SELECT = (r"^Kernel Name$|gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__throughput.avg.pct|smsp__issue_active.avg.pct|"
          r"smsp__inst_executed.sum$|sm__warps_active.avg.pct|launch__registers|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|"
          r"thread_inst_executed_per_inst|launch__grid_size|launch__block_size|shared_mem_per_block_dynamic|"
          r"gpu__dram_throughput.avg.pct|smsp__average_warps_issue_stalled.*per_issue_active|lts__t_bytes.sum$|l1tex__t_bytes.sum$|"
          r"smsp__inst_executed_pipe_.*sum$|sm__inst_executed_pipe_.*sum$|"
          # the L1 data pipe: what bounds the walk
          r"^l1tex__data_pipe_lsu_wavefronts(_mem_(shared|lgds))?(_op_(ld|st|atom))?\.(sum|avg)(\.pct_of_peak_sustained_elapsed)?$|"
          r"TriageCompute.l1tex__data_pipe_lsu_wavefronts|"
          r"^l1tex__lsu_writeback_active(_mem_lgds)?\.(avg|sum)(\.pct_of_peak_sustained_elapsed)?$|"
          r"^lts__throughput.avg.pct_of_peak_sustained_elapsed$|^l1tex__throughput.avg.pct_of_peak_sustained_elapsed$|"
          r"^l1tex__t_(requests|sectors)_pipe_lsu_mem_global_op_ld(_lookup_(hit|miss))?\.sum$|"
          r"^sass__inst_executed_register_spilling|^smsp__sass_inst_executed_op_local_(ld|st)\.sum$|"
          r"^sm__cycles_(elapsed|active)\.(avg|sum)$|^sm__cycles_elapsed.avg.per_second$|"
          r"^smsp__sass_inst_executed_op_(shared|global)_(ld|st)\.sum$|^sm__inst_executed.sum$|^sm__inst_issued.avg.pct")
keys = [k for k in d if re.search(SELECT, k)]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
tob = lambda k: float(d[k][0]) * scale.get(d[k][1], 1)
num = lambda k: float(d[k][0]) if k in d and d[k][0] not in ("", "n/a") else None
derived = {}
wf, wf_lgds, wf_sh = (num("l1tex__data_pipe_lsu_wavefronts.sum"), num("l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum"),
                      num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"))
if wf is None:      # older captures only carry the per-SM averages of the TriageCompute section
    n_sm = 148.0
    for pre in ("SM_A.TriageCompute.",):
        a = num(pre + "l1tex__data_pipe_lsu_wavefronts.avg")
        if a is not None:
            wf = a * n_sm
            wf_lgds = (num(pre + "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg") or 0) * n_sm
            wf_sh = (num(pre + "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg") or 0) * n_sm
req = num("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
sec = num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
cyc = num("sm__cycles_elapsed.sum")
if wf and cyc:
    derived["l1_data_pipe_wavefronts_per_sm_cycle"] = wf / cyc            # peak = 1 wavefront / clock / SM
if wf_lgds and req:
    derived["lgds_wavefronts_per_global_ld_request"] = wf_lgds / req
if sec and req:
    derived["sectors_per_global_ld_request"] = sec / req
if wf is not None and len(sys.argv) > 5 and float(sys.argv[5]) > 0:
    derived["warp_steps_in_profiled_launch"] = float(sys.argv[5])
    # every L1 data-pipe wavefront of the kernel (feature LDS + node gather of the walk, root / constant streams, and the
    # state machine's own shared-memory traffic) per warp-level gather level of the walk
    derived["l1_wavefronts_per_warp_step"] = wf / float(sys.argv[5])
if wf is not None:
    derived["l1_data_pipe_wavefronts"] = wf
    derived["l1_data_pipe_wavefronts_mem_lgds"] = wf_lgds
    derived["l1_data_pipe_wavefronts_mem_shared"] = wf_sh
summ = {"report": rep, "launch_index": launch, "kernel": d.get("Kernel Name", ("?",))[0],
        "games_in_profiled_launch": int(sys.argv[3]) if len(sys.argv) > 3 else 500000,
        "dram_bytes_per_launch": tob("dram__bytes_read.sum") + tob("dram__bytes_write.sum"),
        "derived": derived,
        "metrics": {k: {"value": d[k][0], "unit": d[k][1]} for k in sorted(keys)}}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
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
except Exception as exc:      # a capture without --import-source / the source page
    summ["hot_segments_error"] = repr(exc)
json.dump(summ, open(out + ".json", "w"), indent=1)
for k in sorted(keys): print(k, d[k])
print(json.dumps(derived, indent=1))
for s in summ.get("hot_segments", []): print(s)
