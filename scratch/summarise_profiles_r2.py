"""Turn the artefacts of scratch/profile_round2.sh (gpurun_out/r2_*) into the committed summaries under profiles/:
r2_bench.json (both arms), r2_launches.txt, r2_ncu_full.txt, traffic.json (stamped with the commit it was captured at)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()

def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])

d, ref = last_json(f"{g}/r2_bench.json"), last_json(f"{g}/r2_ref.json")
json.dump({"ours": d, "reference": ref, "commit": head}, open(f"{ROOT}/profiles/{tag}_bench.json", "w"), indent=1)

def short(name):
    return name.split("(")[0].replace("ast::", "").replace("void ", "")

def canon(name):   # the FP16-split decimator is bench.py's / traffic.json's "decimate2_tc_kernel" (the library's profile span)
    return name.replace("decimate2_tc_h_kernel", "decimate2_tc_kernel")

rows = list(csv.reader(open(f"{g}/r2_launches.csv")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
L = [(short(r[ki]), r[gi], r[bi], float(r[vi].replace(",", "")) / 1000.0) for r in rows[h + 2:] if len(r) > vi]
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 80: python scratch/prof_step.py --steps 1 --warmup 1 --legs features,istft,stats",
       f"# commit {head}; per-launch device time in us (cold cache, serialised, programmatic overlap off under the profiler: compare SHARES)",
       "# B200, 64 clips x 10 s: two feature calls (prologue, decimator, STFT, CQT projection), two iSTFT calls, two statistics calls (decimator, STFT, CQT projection, finalise, accumulate)",
       "# kernel | grid | block | us"]
out += [f"{n[:50]:50s} {gr:14s} {bl:12s} {us:9.2f}" for n, gr, bl, us in L]
agg = collections.defaultdict(list)
for n, gr, bl, us in L:
    agg[canon(n)].append(us)
out.append("# ---- per kernel over all captured launches: n, mean us")
out += [f"{k[:50]:50s} n={len(v):3d} mean={sum(v) / len(v):9.2f}" for k, v in agg.items()]
step = {k: sum(v) / len(v) for k, v in agg.items() if k in ("stft_kernel<0>", "decimate2_tc_kernel", "cqt_tc_kernel<1>", "cqt_tc_kernel<1, 1>", "cqt_tc_kernel<1, 0>")}
tot = sum(step.values())
out.append("# ---- share of the feature step (serialised): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in step.items()))
kb = d["roofline"]["kernels"]
bt = sum(v["ms_per_step"] for k, v in kb.items() if k != "istft_kernel")
out.append("# ---- same shares from bench.py's live CUDA events (r2_bench.json): " +
           ", ".join(f"{k} {100 * v['ms_per_step'] / bt:.1f}%" for k, v in kb.items() if k != "istft_kernel"))
open(f"{ROOT}/profiles/{tag}_launches.txt", "w").write("\n".join(out) + "\n")

raw = subprocess.run(["ncu", "-i", f"{g}/r2_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'] + [
    f'smsp__average_warps_issue_stalled_{x}_per_issue_active.ratio' for x in
    ('long_scoreboard', 'short_scoreboard', 'barrier', 'wait', 'mio_throttle', 'lg_throttle', 'math_pipe_throttle', 'not_selected', 'no_instruction',
     'branch_resolving', 'membar', 'sleeping')]
txt = [f"# ncu --set full --clock-control none --import-source on, scratch/prof_step.py --steps 1 --warmup 1 --legs features,istft,stats (64 clips x 10 s), B200 sm_100a, commit {head}",
       "# cold-cache, serialised replays: durations are for shares, bench.py's CUDA-event times are the numbers of record.",
       "# The statistics-mode launches (stft_kernel<1>, and the decimator / CQT launches that follow it) store no features: their dram bytes are the fused pass's traffic.", ""]
traffic, seen, mode = {}, collections.Counter(), "features"
mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
for r in rows[2:]:
    dd, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    name = canon(short(dd["Kernel Name"]))
    if name == "istft_kernel":
        mode = "istft"
    elif mode == "istft" and name.startswith("decimate2"):   # the statistics calls follow the iSTFT calls
        mode = "stats"
    key = name if mode != "stats" or name == "stats_finalize_clips_kernel" else name + " [statistics call]"
    seen[key] += 1
    if seen[key] != 2 and not (seen[key] == 1 and key not in traffic and False):
        if seen[key] > 2:
            continue
        if seen[key] == 1:   # keep the FIRST capture only if no second one follows; the second (warm code) replaces it below
            pass
    txt_block = [f"== {key}   grid {dd['Grid Size']} block {dd['Block Size']}  (capture {seen[key]})"]
    txt_block += [f"   {k:95s} {dd[k]:>16s} {u[k]}" for k in keys if k in dd and dd[k] != ""]
    if seen[key] <= 2:
        traffic[key] = sum(float(dd[k]) * mul[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        if seen[key] == 2:
            txt += txt_block + [""]
open(f"{ROOT}/profiles/{tag}_ncu_full.txt", "w").write("\n".join(txt))
stats_traffic = sum(v for k, v in traffic.items() if "[statistics call]" in k or k == "stats_finalize_clips_kernel")
json.dump({"how": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, 64 clips x 10 s, cold L2 per replay",
           "commit": head, "source": f"gpurun_out/r2_full.ncu-rep summarised in profiles/{tag}_ncu_full.txt",
           "dram_bytes_per_launch": {k.split("<")[0] if "[" not in k else k: v for k, v in traffic.items()},
           "statistics_call_dram_bytes_per_clip": stats_traffic / 64.0,
           "algorithmic_bytes_per_launch": {"stft_kernel": 64 * (882000 + 4711392), "cqt_tc_kernel": 64 * (1764228 + 771616),
                                            "istft_kernel": 64 * 5591008, "decimate2_tc_kernel": 64 * 2604672,
                                            "statistics call (per clip)": 882000}},
          open(f"{ROOT}/profiles/traffic.json", "w"), indent=1)
print(json.dumps({k: v for k, v in d.items() if k in ("value", "ms_per_step")}), d["istft"]["value"], d["e2e"]["value"], ref["value"])
print({k: round(v["ms_per_step"], 4) for k, v in kb.items()}, d["roofline"]["frac"], d["roofline"]["path"]["frac"], d["roofline"]["istft"]["frac"])
print({k: round(v / 1e6, 1) for k, v in traffic.items()}, "stats MB/clip", stats_traffic / 64e6)
