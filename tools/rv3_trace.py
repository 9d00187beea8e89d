"""Per-role timeline of kb_reverb3_kernel (measurement aid): KB_RV3_TRACE=<file> makes the library dump clock64() stamps of CTA 0 (start of
work after the waits, end of work) for every role and chunk of the last launch; this prints them relative to the first stamp.
Usage: KB_RV3_TRACE=/tmp/t.txt python tools/fx_probe.py reverb 4096 [--tol]; python tools/rv3_trace.py /tmp/t.txt"""
import collections
import sys

names = {1: "F filter", 2: "E early", 4: "P premul", 5: "T taps", 6: "W fdn", 7: "S scan"}
names.update({8 + l: "S line %d" % l for l in range(8)})
rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1]) if l.strip()]
t0 = min(r[2] for r in rows)
by = collections.defaultdict(dict)
for r, k, a, b in rows:
    by[r][k] = (a - t0, b - t0)
print("role        " + "".join(f"{'k=%d' % k:>16s}" for k in (0, 1, 2, 3, 10, 11, 12, 30, 31, 50)))
for r in sorted(by):
    print(f"{names.get(r, r):12s}" + "".join((f"{by[r][k][0]:>8d}-{by[r][k][1]:<7d}" if k in by[r] else " " * 16) for k in (0, 1, 2, 3, 10, 11, 12, 30, 31, 50)))
print("\nper-chunk busy time (end - start) and period (start-to-start), cycles, median over chunks 5..45")
import statistics
for r in sorted(by):
    ks = [k for k in sorted(by[r]) if 5 <= k <= 45 and k + 1 in by[r]]
    if not ks:
        continue
    busy = statistics.median(by[r][k][1] - by[r][k][0] for k in ks)
    period = statistics.median(by[r][k + 1][0] - by[r][k][0] for k in ks)
    print(f"{names.get(r, r):12s} busy {busy:8.0f}   period {period:8.0f}")
last = max(b for r in by for (a, b) in by[r].values())
print(f"\ntotal span {last} cycles")
