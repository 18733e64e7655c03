"""Host-side model of the decimator's persistent tile pipeline, used to design the static per-CTA tile lists.
A CTA issues a tile every T_ISSUE us at most and a tile's outputs are visible to dependants LAT us after its issue;
a tile cannot issue before its dependencies are visible (the CTA's list is walked in order)."""
import sys, math
import numpy as np
T_ISSUE, LAT, START = 3.6, 11.0, 3.0
L0 = 220500
def tiles_per_clip(L=L0):
    out = []
    n = L
    for s in range(6):
        n_out = (n + 1) // 2
        out.append((n_out + 7423) // 7424)
        n = n_out
    return out
def lens(L=L0):
    out=[L]
    for s in range(6): out.append((out[-1]+1)//2)
    return out
def deps(s, k, tpc, ln):
    if s == 0: return []
    lo = max(14848 * k - 192, 0); hi = min(14848 * k + 15040, ln[s])
    return [(s - 1, kk) for kk in range(lo // 7424, (hi - 1) // 7424 + 1)]
def simulate(lists, B, tpc, ln):
    """lists[cta] = [(s, clip, k), ...]; returns per-CTA end time and per-tile finish times (event-driven, in-order per CTA)"""
    fin = {}
    pos = [0] * len(lists); t_free = [START] * len(lists); end = [START] * len(lists)
    remaining = sum(len(l) for l in lists)
    progress = True
    while remaining and progress:
        progress = False
        for c, l in enumerate(lists):
            while pos[c] < len(l):
                s, b, k = l[pos[c]]
                d = deps(s, k, tpc, ln)
                if any((ds, b, dk) not in fin for ds, dk in d): break
                ready = max([fin[(ds, b, dk)] for ds, dk in d], default=0.0)
                t = max(t_free[c], ready)
                fin[(s, b, k)] = t + LAT
                t_free[c] = t + T_ISSUE
                end[c] = t + LAT
                pos[c] += 1; remaining -= 1; progress = True
    assert remaining == 0, "deadlock"
    return np.array(end), fin
def current_scheme(B, C, tail, tpc):
    """stage-major round-robin bulk over all CTAs; chain stages (<= 2 tiles per clip) on the last `tail` CTAs"""
    first_chain = 6
    while first_chain > 0 and tpc[first_chain - 1] <= 2: first_chain -= 1
    if tail == 0: first_chain = 6
    lists = [[] for _ in range(C)]
    i = 0
    for s in range(first_chain):
        for b in range(B):
            for k in range(tpc[s]):
                lists[i % C].append((s, b, k)); i += 1
    for s in range(first_chain, 6):
        for b in range(B):
            j = C - tail + b % tail
            for k in range(tpc[s]): lists[j].append((s, b, k))
    # chain lists must be stage by stage, clip by clip: already (outer loop s)
    return lists
def report(name, end, C):
    e = np.sort(end)
    print(f"{name:40s} SM.us {end.sum():8.0f}  = {end.sum()/C:6.1f} us of the GPU   med end {np.median(end):6.1f}  p90 {np.percentile(end,90):6.1f}  max {end.max():6.1f}")
if __name__ == "__main__":
    B, C = 64, 148
    tpc, ln = tiles_per_clip(), lens()
    print("tiles per clip", tpc, "total", sum(tpc) * B)
    for tail in (0, 64, 32, 16, 8):
        lists = current_scheme(B, C, tail, tpc)
        end, fin = simulate(lists, B, tpc, ln)
        report(f"current scheme, tail {tail}", end, C)

def wavefront_scheme(B, C, Kc, g, skew, tpc, ln, lat_plan=15.0):
    """bulk stages in skewed clip-group order over the first C - Kc CTAs; chain stages on Kc dedicated CTAs whose lists
    are ordered greedily by the modelled ready times"""
    first_chain = 6
    while first_chain > 0 and tpc[first_chain - 1] <= 2: first_chain -= 1
    nb = first_chain
    G = (B + g - 1) // g
    order = []
    for w in range(G + skew * (nb - 1)):
        for s in range(nb):
            grp = w - skew * s
            if 0 <= grp < G:
                for b in range(grp * g, min(B, (grp + 1) * g)):
                    for k in range(tpc[s]): order.append((s, b, k))
    Cb = C - Kc
    lists = [[] for _ in range(C)]
    for i, t in enumerate(order): lists[i % Cb].append(t)
    # model the bulk to get ready times of the chain's inputs
    global LAT
    keep = LAT; LAT = lat_plan
    _, fin = simulate(lists[:Cb], B, tpc, ln)
    # greedy chain order per chain CTA
    for j in range(Kc):
        clips = list(range(j, B, Kc))
        nxt = {b: 0 for b in clips}            # index into the clip's chain tile sequence
        seq = {b: [(s, b, k) for s in range(nb, 6) for k in range(tpc[s])] for b in clips}
        t_free = START
        while any(nxt[b] < len(seq[b]) for b in clips):
            best = None
            for b in clips:
                if nxt[b] == len(seq[b]): continue
                s, _, k = seq[b][nxt[b]]
                d = deps(s, k, tpc, ln)
                ready = max(fin[(ds, b, dk)] for ds, dk in d)
                if best is None or ready < best[0]: best = (ready, b)
            ready, b = best
            t = max(t_free, ready)
            tile = seq[b][nxt[b]]; nxt[b] += 1
            fin[tile] = t + LAT; t_free = t + T_ISSUE
            lists[Cb + j].append(tile)
    LAT = keep
    return lists

if __name__ == "__main__":
    for lat in (11.0, 15.0, 20.0):
        LAT = lat
        print("LAT", lat)
        for tail in (0, 64):
            end, _ = simulate(current_scheme(B, C, tail, tpc), B, tpc, ln); report(f"  current, tail {tail}", end, C)
        for Kc in (8, 12, 16, 24, 32):
            for g in (4, 8, 16):
                for skew in (1, 2, 3):
                    lists = wavefront_scheme(B, C, Kc, g, skew, tpc, ln)
                    end, _ = simulate(lists, B, tpc, ln)
                    report(f"  wavefront Kc {Kc} g {g} skew {skew}", end, C)
