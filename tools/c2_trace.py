"""Per-role timeline of kb_sub_tiled_kernel (measurement aid): KB_C2_TRACE=<file> makes the library dump clock64() stamps of CTA 0 for every tick
(tile) of the last launch: row 0 = envelope warp A (start of tick, end of its work), 1 = filter warp C, 2 = a worker (start, end of B), 3 = the same
worker's end of D.  Usage: KB_C2_TRACE=/tmp/t.txt python tools/c2_probe.py sub; python tools/c2_trace.py /tmp/t.txt"""
import collections
import statistics
import sys

rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1]) if l.strip()]
by = collections.defaultdict(dict)
for r, k, a, b in rows:
    by[r][k] = (a, b)
ks = [k for k in sorted(by[0]) if 6 <= k <= 28 and k + 1 in by[0]]
tick = statistics.median(by[0][k + 1][0] - by[0][k][0] for k in ks)
print(f"tick (start to start) {tick:.0f} cycles")
for r, name in ((0, "A envelopes"), (1, "C filter")):
    print(f"{name:14s} busy {statistics.median(by[r][k][1] - by[r][k][0] for k in ks):8.0f}")
print(f"{'workers B':14s} busy {statistics.median(by[2][k][1] - by[2][k][0] for k in ks):8.0f}")
print(f"{'workers B+D':14s} busy {statistics.median(by[3][k][1] - by[2][k][0] for k in ks):8.0f}")

cta = [(a, b) for r in (4, 5, 6, 7) for k, (a, b) in sorted(by.get(r, {}).items())]
if cta:
    pro = [a for a, b in cta]; loop = [b for a, b in cta]
    print(f"per CTA ({len(cta)}): prologue median {statistics.median(pro):.0f} max {max(pro)}; tile loop median {statistics.median(loop):.0f} min {min(loop)} max {max(loop)} cycles")

if "--raw" in sys.argv:
    # every tick of CTA 0: start / end of each role relative to A's first stamp
    t0 = min(a for r in (0, 1, 2) for a, b in by[r].values() if a)
    print("tick |  A start   run-end  signalled |  C start    C end |  W iter-top  B start  B computed  B signalled   D end")
    for k in range(0, 40):
        g = lambda r, ph: (by[r][k][ph] - t0) if k in by[r] and by[r][k][ph] else -1
        print(f"{k:4d} | {g(0,0):8d} {g(8,0):9d} {g(0,1):10d} | {g(1,0):8d} {g(1,1):8d} | {g(9,0):10d} {g(2,0):8d} {g(9,1):11d} {g(2,1):12d} {g(3,1):7d}")
