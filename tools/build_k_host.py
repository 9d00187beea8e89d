#!/usr/bin/env python3
"""Builds tests/_k_bin/k_host: the reference's example `.k` programs, unmodified, compiled against this repo's own
include/compat/klang.h and linked to libklang_b200.so.  The `.k` sources are read where they lie under
/root/reference/examples (never copied); only the binary lands in tests/_k_bin/ (git-ignored, travels to the GPU box)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KLANG_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(ROOT, "tests", "_k_bin")
OUT = os.path.join(OUT_DIR, "k_host")


def build(verbose=True, overlay=None, out=None):
    """overlay: a directory searched BEFORE the reference examples (a user's edited copies of some programs); out: the binary to write."""
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import k_hash
    out = out or OUT
    if not os.path.isfile(os.path.join(REF, "examples", "PingPong.k")):
        if os.path.isfile(out):
            return out
        raise SystemExit(f"build_k_host: no reference examples under {REF} and no prebuilt {out}")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    libdir = os.path.join(ROOT, "klang_b200", "lib")
    search = ([overlay] if overlay else []) + [os.path.join(REF, "examples")]
    # the hash of every program k_host.cpp includes, taken from the file the compiler will actually find
    with open(os.path.join(ROOT, "tools", "k_host.cpp")) as f:
        rels = re.findall(r'#include "([^"]+\.k)"', f.read())
    gen = os.path.join(os.path.dirname(out), "k_hashes_" + os.path.basename(out))
    os.makedirs(gen, exist_ok=True)
    with open(os.path.join(gen, "k_hashes.h"), "w") as f:
        for rel in rels:
            found = next(os.path.join(d, rel) for d in search if os.path.isfile(os.path.join(d, rel)))
            f.write(f"#define {k_hash.macro_name(rel)} 0x{k_hash.k_hash(found):016X}ULL\n")
    cmd = ["g++", "-std=c++17", "-O1", "-w", os.path.join(ROOT, "tools", "k_host.cpp"), "-I", gen,
           "-I", os.path.join(ROOT, "include", "compat")] + [x for d in search for x in ("-I", d)] + [
           "-L", libdir, "-lklang_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,$ORIGIN/../../klang_b200/lib", "-o", out]
    if verbose:
        print("build_k_host:", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build())
