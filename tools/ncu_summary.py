"""Summarise an .ncu-rep (one kernel launch) into a small JSON for profiles/: selected raw metrics, warp-stall
sampling totals, instruction mix and the SASS lines with the most stall samples.  Runs without a GPU.
   python tools/ncu_summary.py <report.ncu-rep> <out.json> ["note"]"""
import collections, csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""


def page(name):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True,
                                                       text=True, check=True).stdout)))


raw = page("raw")
m = dict(zip(raw[0], zip(raw[1], raw[2])))
keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
sel = {k: {"value": m[k][1], "unit": m[k][0]} for k in keep if k in m}
stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(v[1])) for k, v in m.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k}
res = {"report": rep, "note": note, "metrics": sel, "warp_stall_samples": dict(sorted(stalls.items(), key=lambda kv: -kv[1]))}
try:
    src = page("source")
    hdr = src[1]
    ix = {h: i for i, h in enumerate(hdr)}
    rows = src[2:]
    tot_s = sum(int(r[ix["# Samples"]]) for r in rows)
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in rows)
    mix = collections.Counter()
    for r in rows:
        t = r[ix["Source"]].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        mix[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
    top = sorted(rows, key=lambda r: -int(r[ix["# Samples"]]))[:25]
    res["sass"] = {"stall_samples_total": tot_s, "warp_instructions_total": tot_i,
                   "instruction_mix_pct": {k: round(100.0 * v / tot_i, 2) for k, v in mix.most_common(20)},
                   "top_lines": [{"samples": int(r[ix["# Samples"]]), "executed": int(r[ix["Instructions Executed"]]),
                                  "sass": " ".join(r[ix["Source"]].split()),
                                  "long_sb": int(r[ix["stall_long_sb"]]), "wait": int(r[ix["stall_wait"]]),
                                  "barrier": int(r[ix["stall_barrier"]]), "short_sb": int(r[ix["stall_short_sb"]])}
                                 for r in top]}
except Exception as e:      # reports captured without --import-source still have the instruction-level page
    res["sass_error"] = str(e)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v["value"] for k, v in sel.items()}, indent=1))
print(res["warp_stall_samples"])
