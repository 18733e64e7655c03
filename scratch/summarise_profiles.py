"""Turn the artefacts of scratch/profile_round.sh (gpurun_out/) into the committed summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_v4"
for r in ("r1_v3_features", "r1_v3_istft"):
    with open(f"{g}/{r}.raw.csv", "w") as f:
        subprocess.run(["ncu", "-i", f"{g}/{r}.ncu-rep", "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
d = json.loads(open(f"{g}/bench_default.json").read().strip().splitlines()[-1])
ref = json.loads(open(f"{g}/bench_ref.json").read().strip().splitlines()[-1])
json.dump({"ours": d, "reference": ref}, open(f"{ROOT}/profiles/{tag}_bench.json", "w"), indent=1)

rows = list(csv.reader(open(f"{g}/r1_v3_launches.csv")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
L = [(r[ki].split("(")[0].replace("ast::", ""), r[gi], r[bi], float(r[vi].replace(",", "")) / 1000.0) for r in rows[h + 2:] if len(r) > vi]
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e",
       "# per-launch device time in us (cold cache, serialised, PDL overlap off under the profiler: compare SHARES, not absolutes)",
       "# B200, 64 clips x 10 s per step; kernel | grid | block | us"]
out += [f"{n[:60]:60s} {gr:14s} {bl:12s} {us:9.2f}" for n, gr, bl, us in L[:40]]
agg = collections.defaultdict(list)
for n, gr, bl, us in L:
    agg[n].append(us)
out.append("# ---- per kernel over all captured launches: n, mean us")
out += [f"{k[:60]:60s} n={len(v):3d} mean={sum(v) / len(v):9.2f}" for k, v in agg.items()]
step = {k: sum(v) / len(v) for k, v in agg.items() if k in ("stft_kernel", "decimate2_tc_kernel", "cqt_tc_kernel")}
tot = sum(step.values())
out.append("# ---- share of the feature step (stft + decimate2_tc + cqt_tc): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in step.items()))
open(f"{ROOT}/profiles/{tag}_launches.txt", "w").write("\n".join(out) + "\n")

keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'] + [
        f'smsp__average_warps_issue_stalled_{x}_per_issue_active.ratio' for x in
        ('long_scoreboard', 'short_scoreboard', 'barrier', 'wait', 'mio_throttle', 'lg_throttle', 'math_pipe_throttle', 'not_selected', 'membar', 'sleeping')]
txt = ["# ncu --set full --clock-control none --import-source on, bench.py --steps 2 --warmup 3 (64 clips x 10 s), B200 sm_100a",
       "# cold-cache, serialised replays: durations are for shares, bench.py's CUDA-event times are the numbers of record", ""]
traffic = {}
seen = set()
for f in ("r1_v3_features.raw.csv", "r1_v3_istft.raw.csv"):
    rows = list(csv.reader(open(f"{g}/{f}")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        dd, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = dd["Kernel Name"].split("(")[0].replace("ast::", "").replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        txt.append(f"== {name}   grid {dd['Grid Size']} block {dd['Block Size']}")
        txt += [f"   {k:95s} {dd[k]:>16s} {u[k]}" for k in keys if k in dd and dd[k] != ""]
        txt.append("")
        mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
        traffic[name] = sum(float(dd[k]) * mul[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
open(f"{ROOT}/profiles/{tag}_ncu_full.txt", "w").write("\n".join(txt))
json.dump({"how": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, bench.py workload "
                  "(64 clips x 10 s), cold L2 per replay", "dram_bytes_per_launch": traffic,
           "algorithmic_bytes_per_launch": {"stft_kernel": 64 * (882000 + 4711392), "cqt_tc_kernel": 64 * (1764228 + 771616),
                                            "istft_kernel": 64 * 5591008, "decimate2_tc_kernel": 64 * 2604672}},
          open(f"{ROOT}/profiles/traffic.json", "w"), indent=1)
print(json.dumps({k: v for k, v in d.items() if k in ("value", "ms_per_step")}), d["istft"]["value"], d["e2e"]["value"], ref["value"])
print({k: round(v["ms_per_step"], 4) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["frac"], d["roofline"]["path"]["frac"], d["roofline"]["istft"]["frac"])
print(traffic)
