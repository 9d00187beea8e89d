#!/usr/bin/env python3
"""Attribute an ncu --set full capture to SOURCE LINES: warp-state samples and executed warp instructions per line.

  python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> [min_samples]

The SASS rows of `ncu --page source --csv` carry addresses; `nvdisasm -g` of the cubin inside klang_b200/lib gives the
address -> file:line map of the same function (the library must be the build the capture ran)."""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("KB_LIB", os.path.join(ROOT, "klang_b200", "lib", "libklang_b200.so"))], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    m, infun, cur = {}, False, None
    for line in out.splitlines():
        s = re.match(r"\s*\.section\s+\.text\.(\S+),", line)
        if s:
            infun = kernel in s.group(1)
            continue
        if not infun:
            continue
        s = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if s:
            cur = (os.path.basename(s.group(1)), int(s.group(2)))
            continue
        s = re.search(r"/\*([0-9a-f]{4,6})\*/", line)
        if s and cur:
            m[int(s.group(1), 16)] = cur
    return m


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    min_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    lm = line_map(kernel)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, body = rows[1], rows[2:]
    ai, si, ii = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = [k for k, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
    base = min(int(b[ai], 16) for b in body)
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for b in body:
        key = lm.get(int(b[ai], 16) - base, ("?", 0))
        p = per[key]
        p[0] += int(b[si]); p[1] += int(b[ii])
        for k in cols:
            p[2][hdr[k][6:]] += int(b[k])
    tot_s = sum(p[0] for p in per.values()); tot_i = sum(p[1] for p in per.values())
    print(f"# {rep}: {kernel}: {tot_s} samples, {tot_i} warp instructions; lines with >= {min_samples} samples")
    print(f"{'file:line':28s} {'samples':>8s} {'%':>6s} {'warp-inst':>10s} {'%':>6s}  top stalls")
    for key, p in sorted(per.items()):
        if p[0] >= min_samples:
            top = ", ".join(f"{n} {v}" for n, v in p[2].most_common(3) if v)
            print(f"{key[0] + ':' + str(key[1]):28s} {p[0]:8d} {100 * p[0] / tot_s:6.1f} {p[1]:10d} {100 * p[1] / max(1, tot_i):6.1f}  {top}")


if __name__ == "__main__":
    main()
