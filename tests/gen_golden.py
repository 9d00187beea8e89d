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
    for fs in (44100, 48000):
        path = os.path.join(HERE, "golden", f"klang_ref_translated_fs{fs}.npz")
        out = cases.translated_cases(oracle.ref, fs)
        for name in cases.TRANSLATED_SYNTH_SCRIPTS:
            out.update(cases.translated_synth_cases(oracle.ref, name, fs))
        np.savez_compressed(path, **out)
        print(path, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")
    import json
    path = os.path.join(HERE, "golden", "presets.json")
    with open(path, "w") as f:
        json.dump(preset_tables(oracle.ref), f, indent=1)
    print(path)


def preset_tables(eng):
    """Plugin::presets of every bound program (klang.h:1940-1981): {"fx/<name>" | "synth/<name>": [[preset name, [values]], ...]}."""
    eng.set_fs(48000)
    out = {}
    for g, nm in cases.FX_NAMES.items():
        fx = eng.Fx(g)
        out[f"fx/{nm}"] = [[n, v] for n, v in fx.presets()]
        fx.close()
    for g, nm in cases.SY_NAMES.items():
        sy = eng.Synth(g, 4)
        out[f"synth/{nm}"] = [[n, v] for n, v in sy.presets()]
        sy.close()
    return out


if __name__ == "__main__":
    main()
