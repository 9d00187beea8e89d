#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — builds oracle/_ref/libklang_ref.so from the reference sources.

The reference (nashaudio/klang, /root/reference/klang.h v0.7.8 + examples/*.k) is a
header-only C++17 library.  It compiles with g++ 13 only after

  * compat flags for the missing Linux branch of its platform macros and for missing
    includes (klang.h:20-42, 375, 1399, 2378, 2386, 2546, 3203),
  * ONE mechanical patch to a build-time copy: `signals<N>` holds an anonymous struct
    with constructor-bearing members inside an anonymous union (klang.h:1205-1211),
    which g++ rejects with a hard error.  The patch replaces the union by plain
    members `l, r, _more[N-2]` plus an array-view accessor, preserving layout and every
    constructor's semantics (including the "fill only the supplied channels" behaviour
    of the variadic constructors, klang.h:1237-1241),
  * a one-token disambiguation in copies of TB303.k:106 and SynTHX.k:10, which do not
    compile with g++ as written (ambiguous operator*, Mono:: name lookup).

Patched copies live in a temporary directory that is deleted afterwards; reference
sources are never copied into the repository.  Only the shared object lands in
oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).

Flags mirror the reference's own Linux Release build (-O3 -std=c++17,
templates/juce/synth/Builds/LinuxMakefile/Makefile:103-104) plus -ffp-contract=off and
no -march=native, so the same object is both the parity oracle and the CPU baseline.
`-include math.h` pins unqualified sin/tanh/exp(float) to the float overloads, as on the
reference's native platforms (MSVC / libc++).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KLANG_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libklang_ref.so")

COMPAT = [
    "-DTHREAD_LOCAL=thread_local", "-DSQRT=::sqrt", "-DSQRTF=::sqrtf", "-DABS=::abs", "-DFABS=::fabsf",
    "-include", "math.h", "-include", "cstring", "-include", "climits", "-include", "chrono",
    "-include", "tuple", "-include", "cstdint",
    "-fpermissive", "-w",
]
CXXFLAGS = ["-std=c++17", "-O3", "-ffp-contract=off", "-fPIC", "-shared"]


def replace_once(text, old, new, what):
    if text.count(old) != 1:
        raise SystemExit(f"build_ref: expected exactly one occurrence of {what!r} in the reference, "
                         f"found {text.count(old)} — reference differs from v0.7.8 @5c5ecdd")
    return text.replace(old, new)


def patch_header(src):
    # 1. the anonymous union/struct (klang.h:1205-1211)
    old_union = (
        "\t\tunion {\n"
        "\t\t\tsignal value[CHANNELS]; ///< Array of channel values.\n"
        "\t\t\tstruct {\n"
        "\t\t\t\tsignal l; ///< Left channel\n"
        "\t\t\t\tsignal r; ///< Right channel\n"
        "\t\t\t};\n"
        "\t\t};\n")
    new_union = (
        "\t\tsignal l; signal r; signal _more[(CHANNELS > 2) ? (CHANNELS - 2) : 0];\n"
        "\t\tsignal* _value() { return reinterpret_cast<signal*>(this); }\n"
        "\t\tconst signal* _value() const { return reinterpret_cast<const signal*>(this); }\n")
    src = replace_once(src, old_union, new_union, "signals<> anonymous union")

    # 2. uses of value[...] inside struct signals only (klang.h:1213-1315)
    start = src.index(new_union)
    end = src.index("/// Return a copy of the signal with each channel offset by x.\n\ttemplate<int CHANNELS = 2> inline signals<CHANNELS> operator+(float x")
    body = src[start:end]
    # the two variadic constructors: value{ initial... } fills the supplied channels, the rest stay 0
    old_ctor = "signals(Args&... initial) : value{ initial... } {}"
    new_ctor = ("signals(Args&... initial) { const signal _tmp[] = { initial... }; "
                "for (unsigned _i = 0; _i < sizeof...(Args) && _i < (unsigned)CHANNELS; _i++) _value()[_i] = _tmp[_i]; }")
    body = replace_once(body, old_ctor, new_ctor, "signals(Args&...) ctor")
    old_ctor2 = "signals(Args... initial) : value{ initial... } {}"
    new_ctor2 = ("signals(Args... initial) { const signal _tmp[] = { initial... }; "
                 "for (unsigned _i = 0; _i < sizeof...(Args) && _i < (unsigned)CHANNELS; _i++) _value()[_i] = _tmp[_i]; }")
    body = replace_once(body, old_ctor2, new_ctor2, "signals(Args...) ctor")
    body = body.replace("value[", "_value()[")
    return src[:start] + body + src[end:]


def patch_tb303(src):
    # TB303.k:106 — ambiguous operator* between signal and Function<float>& under g++
    return replace_once(src, "* fs.nyquist) * sqr(env++);", "* fs.nyquist) * (float)sqr(env++);", "TB303.k:106")


def patch_synthx(src):
    # SynTHX.k:10 — `Mono` resolves to Stereo::Synth::Mono (klang.h:4764) inside the Synth
    return replace_once(src, "struct Partial : Mono::Oscillator {", "struct Partial : klang::Mono::Oscillator {", "SynTHX.k:10")


def build(verbose=True):
    if not os.path.isfile(os.path.join(REF, "klang.h")):
        if os.path.isfile(OUT):
            if verbose:
                print(f"build_ref: {REF} absent, keeping prebuilt {OUT}")
            return OUT
        raise SystemExit(f"build_ref: reference not found at {REF} and no prebuilt {OUT}")
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="klang_ref_build_")
    try:
        with open(os.path.join(REF, "klang.h"), encoding="utf-8", errors="surrogateescape") as f:
            hdr = f.read()
        with open(os.path.join(tmp, "klang.h"), "w", encoding="utf-8", errors="surrogateescape") as f:
            f.write(patch_header(hdr))
        for name, fn in (("TB303", patch_tb303), ("SynTHX", patch_synthx)):
            with open(os.path.join(REF, "examples", name + ".k"), encoding="utf-8", errors="surrogateescape") as f:
                k = f.read()
            with open(os.path.join(tmp, name + "_patched.k"), "w", encoding="utf-8", errors="surrogateescape") as f:
                f.write(fn(k))
        cmd = (["g++"] + CXXFLAGS + COMPAT + ["-I", tmp, "-I", os.path.join(REF, "examples"),
               os.path.join(HERE, "ref_harness.cpp"), "-o", OUT])
        if verbose:
            print("build_ref:", " ".join(cmd))
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    build()
    print("built", OUT)
