"""Builds klang_b200/lib/libklang_b200.so (the C-ABI library: CUDA kernels + host event code) in-tree with nvcc.

sm_100a only.  Floating-point flags pin the reference's arithmetic: no FMA contraction, no
flush-to-zero, IEEE division and square root; the host half is compiled without contraction as well."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OUT = os.path.join(OUT_DIR, "libklang_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "-cudart", "static",
]


def sources():
    return [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))]


def up_to_date():
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources() + [os.path.join(HERE, "..", "include", "klang_b200.h"), __file__])


def build(force=False, verbose=True):
    if not force and up_to_date():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.isfile(OUT):
            return OUT
        raise RuntimeError("nvcc not found and no prebuilt " + OUT)
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v"] * bool(os.environ.get("KB_PTXAS_V")) + [os.path.join(SRC, "kb_api.cu"), "-o", OUT]
    if verbose:
        print("klang_b200.build:", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
