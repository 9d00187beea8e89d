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


def build(verbose=True):
    if not os.path.isfile(os.path.join(REF, "examples", "PingPong.k")):
        if os.path.isfile(OUT):
            return OUT
        raise SystemExit(f"build_k_host: no reference examples under {REF} and no prebuilt {OUT}")
    os.makedirs(OUT_DIR, exist_ok=True)
    libdir = os.path.join(ROOT, "klang_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O1", "-w", os.path.join(ROOT, "tools", "k_host.cpp"),
           "-I", os.path.join(ROOT, "include", "compat"), "-I", os.path.join(REF, "examples"),
           "-L", libdir, "-lklang_b200", "-Wl,-rpath,$ORIGIN/../../klang_b200/lib", "-o", OUT]
    if verbose:
        print("build_k_host:", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build())
