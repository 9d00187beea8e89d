"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/klang_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import klang_b200 as kb
from klang_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    build.build(verbose=False)
    return kb.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "klang_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(library):
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(library, name), f"{name} declared in include/klang_b200.h but not exported"


def test_binding_covers_header(library):
    assert sorted(n for n, _, _ in api.SYMBOLS) == declared_symbols()


def test_version_and_helpers(library):
    assert library.kb_version() == 100
    assert abs(library.kb_pitch_to_frequency(69.0) - 440.0) < 1e-4
    assert abs(library.kb_pitch_to_frequency(60.0) - 261.625549) < 1e-4


@pytest.mark.skipif(kb.device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback(library):
    """Without a device every bank constructor and primitive fails loudly (KB_ENODEV)."""
    with pytest.raises(kb.KlangB200Error, match="no CPU path"):
        kb.FxBank(kb.FX_GAIN)
    with pytest.raises(kb.KlangB200Error, match="no CPU path"):
        kb.SynthBank(kb.SY_SUBTRACTIVE)
    with pytest.raises(kb.KlangB200Error):
        kb.Engine().osc(0, 16, 441.0)
    assert library.kb_fx_bank_create(0, 1, C.c_float(48000.0), 64, 0) is None
    assert b"no such CUDA device" in library.kb_last_error()


def test_bad_arguments_are_reported(library):
    assert library.kb_fx_bank_create(99, 1, C.c_float(48000.0), 64, 0) is None
    assert b"bad argument" in library.kb_last_error()
    assert library.kb_synth_bank_create(0, 1, 1000, C.c_float(48000.0), 64, 0) is None   # > 128 voices (klang.h:4311)
    assert library.kb_fx_bank_process(None, None, 4, 0) == -1
    assert library.kb_synth_bank_note_on(None, 0, 60, C.c_float(1.0)) == -1


def test_product_does_not_import_oracle():
    """The product package must not reference the oracle (test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "klang_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "klang_port" not in text, f


def test_cpp_shim_compiles_links_and_fails_loudly_without_a_device(library, tmp_path):
    """include/klang_b200.hpp (the C++ block-driver shim of INTEGRATION.md) builds against the library."""
    import subprocess
    src = tmp_path / "shim.cpp"
    src.write_text('''
#include <cstdio>
#include <klang_b200.hpp>
int main() {
    std::vector<float> block(2 * 256, 0.25f);
    try {
        klang_b200::Effect fx(KB_FX_PINGPONG, 48000.f, 256);
        klang_b200::Synth sy(KB_SY_SUBTRACTIVE, 48000.f, 256, 32);
        fx.setControl(0, 0.7f);
        if (!fx.process(block.data(), 256)) return 3;
        sy.noteOn(60, 0.8f);
        if (!sy.process(block.data(), 256)) return 4;
        std::printf("ran on device %g\\n", block[0]);
        return 0;
    } catch (const klang_b200::Error& e) {
        std::printf("no device: %s\\n", e.what());
        return kb_device_count() == 0 ? 0 : 5;
    }
}
''')
    exe = tmp_path / "shim"
    libdir = os.path.join(ROOT, "klang_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", str(src), "-I", os.path.join(ROOT, "include"), "-L", libdir, "-lklang_b200",
                           "-Wl,-rpath," + libdir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert ("no device" in out.stdout) == (kb.device_count() == 0)


def test_preset_tables_match_the_reference_fixture(library):
    """kb_graph_num_presets / kb_graph_preset (klang_b200/csrc/kb_presets.h) against tests/golden/presets.json, the tables of the compiled
    reference (Plugin::presets, klang.h:1940-1981): same presets, same names, same float32 values for every bound program."""
    import json
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    with open(os.path.join(ROOT, "tests", "golden", "presets.json")) as f:
        want = json.load(f)
    seen = 0
    for is_synth, names in ((0, cases.FX_NAMES), (1, cases.SY_NAMES)):
        for g, nm in names.items():
            got = kb.presets(is_synth, g)
            w = want[("synth/" if is_synth else "fx/") + nm]
            assert [n for n, _ in got] == [n for n, _ in w], nm
            for (_, gv), (_, wv) in zip(got, w):
                assert np.array_equal(np.array(gv, np.float32).view(np.uint32), np.array(wv, np.float32).view(np.uint32)), nm
            seen += len(got)
    assert seen == 17
    assert library.kb_graph_preset(0, 0, 0, None, 0, None, 0) < 0            # Gain.k has none


def test_wav_decode_matches_the_reference_golden(library):
    """kb_wav_decode is host code (as File::WAV is in the reference, klang.h:5951-6085), so it is checked here, on the CPU, against the golden
    vectors the compiled reference decoded from the same file images — every encoding, bit for bit — plus the images it must refuse."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    g = np.load(os.path.join(ROOT, "tests", "golden", "klang_ref_fs48000.npz"))
    images = cases.wav_images()
    assert len(images) == 7
    for name, image in images.items():
        y, info = kb.wav_decode(image)
        assert y.shape == g[f"wav/{name}"].shape, name
        assert np.array_equal(y.view(np.uint32), g[f"wav/{name}"].view(np.uint32)), name
        assert list(info) == list(g[f"wav/{name}/info"]), name
    good = images["pcm16"]
    for bad in (good[:8], b"RIFX" + good[4:], good[:12], good[:-100], good[:36]):
        with pytest.raises(kb.KlangB200Error):
            kb.wav_decode(bad)
