"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native path, from the built library:

    python scratch/sass_excerpt.py > profiles/r2_sass_excerpt.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk,
UBLKPF = bulk L2 prefetch, SYNCS = mbarrier, STG.E.128 / LDG.E.128 = 16-byte global accesses (B200_PROFILING.md)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "audio-style-transfer_b200", "libast_frontend.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
keys = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FADD2", "FMUL2",
        "LDG.E.128", "STG.E.128", "LDG.E.64", "STG.E.64", "LDG.E ", "STG.E ", "LDS.128", "STS.128", "SHFL", "HMMA", "LDGSTS"]
print(f"# cuobjdump -sass audio-style-transfer_b200/libast_frontend.so (built from {head}+, nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3)")
print("# instructions per kernel, then occurrences of each mnemonic (static counts over all code paths)")
print(f"{'kernel':44s} {'instr':>6s} " + " ".join(f"{k.strip():>9s}" for k in keys))
tot = collections.Counter()
for f in re.split(r"\n\s+Function : ", sass)[1:]:
    name = f.split("\n")[0]
    m = re.match(r"_ZN3ast(\d+)([A-Za-z0-9_]+)", name)
    short = m.group(2)[: int(m.group(1))] if m else name[:40]
    if "ILi" in name or "ILb" in name:
        short += "<" + re.search(r"IL[ib](\d)", name).group(1) + ">"
    lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    c = {k: sum(1 for l in lines if (k in l)) for k in keys}
    for k in keys:
        tot[k] += c[k]
    if len(lines) > 40:
        print(f"{short:44s} {len(lines):6d} " + " ".join(f"{c[k]:9d}" for k in keys))
print(f"{'TOTAL':44s} {'':6s} " + " ".join(f"{tot[k]:9d}" for k in keys))
