#!/usr/bin/env python3
"""Generates tests/golden/klang_ref_fs{44100,48000}.npz by running the cases of tests/cases.py
through the COMPILED REFERENCE (oracle/_ref/libklang_ref.so = /root/reference/klang.h + examples,
built by oracle/build_ref.py).  Run in the build container (where /root/reference exists):

    python oracle/build_ref.py && python tests/gen_golden.py

The reference ships no tests or golden vectors of its own, so these files are what
pins the oracle: tests/test_oracle.py checks oracle/klang_port.c against them bit-for-bit."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import cases  # noqa: E402
import oracle  # noqa: E402


def main():
    os.makedirs(os.path.join(HERE, "golden"), exist_ok=True)
    for fs in (44100, 48000):
        out = {}
        out.update(cases.primitive_cases(oracle.ref, fs))
        out.update(cases.all_graph_cases(oracle.ref, fs))
        path = os.path.join(HERE, "golden", f"klang_ref_fs{fs}.npz")
        np.savez_compressed(path, **out)
        print(path, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
