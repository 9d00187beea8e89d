"""CPU tests of the oracle (test infrastructure): the plain-C restatement oracle/klang_port.c must
reproduce, bit for bit, the golden vectors recorded from the compiled reference, and — where the
compiled reference is present (oracle/_ref, built from /root/reference) — the reference itself must
still reproduce them (guards against a stale golden file)."""
import numpy as np
import pytest

import cases
import oracle


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def _assert_bit_exact(got, want, name):
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    same = _bits(got) == _bits(want)
    if not same.all():
        idx = np.argwhere(~same)[0]
        raise AssertionError(f"{name}: first mismatch at {tuple(idx)}: got {got[tuple(idx)]!r} want {want[tuple(idx)]!r} "
                             f"({(~same).sum()} of {same.size} differ)")


@pytest.mark.parametrize("fs", [44100, 48000])
def test_port_primitives_match_golden(golden, fs):
    got = cases.primitive_cases(oracle.port, fs)
    assert set(got) <= set(golden[fs])
    for name, arr in got.items():
        _assert_bit_exact(arr, golden[fs][name], name)


@pytest.mark.parametrize("fs", [44100, 48000])
def test_port_graphs_match_golden(golden, fs):
    got = cases.all_graph_cases(oracle.port, fs)
    for name, arr in got.items():
        _assert_bit_exact(arr, golden[fs][name], name)


@pytest.mark.skipif(not oracle.ref.available(), reason="compiled reference not present")
@pytest.mark.parametrize("fs", [44100, 48000])
def test_reference_reproduces_golden(golden, fs):
    got = {}
    got.update(cases.primitive_cases(oracle.ref, fs))
    got.update(cases.all_graph_cases(oracle.ref, fs))
    assert set(got) == set(golden[fs])
    for name, arr in got.items():
        _assert_bit_exact(arr, golden[fs][name], name)


def test_survey_known_answers(golden):
    """Known answers recorded independently by the survey probe (fs = 44100)."""
    g = golden[44100]
    np.testing.assert_allclose(g["osc/fast_saw/f441.0"][:6],
                               [0.989998341, 0.970000029, 0.950000048, 0.930000067, 0.910000086, 0.890000105], rtol=0, atol=1e-9)
    np.testing.assert_allclose(g["osc/fast_sine/f441.0"][:4], [0, 0.0627904683, 0.12533313, 0.187381178], rtol=0, atol=1e-9)
    np.testing.assert_allclose(g["osc/fast_saw/f441_p0_d0.05"][:6],
                               [-0.600000024, 0.199999988, 1, 0.876923084, 0.958974421, 0.938461602], rtol=0, atol=1e-9)
    np.testing.assert_allclose(g["osc/fast_triangle/f441.0"][:4], [-0.979999661, -0.940000057, -0.900000095, -0.860000134], rtol=0, atol=1e-9)
    np.testing.assert_allclose(g["osc/wt_sine/f441.0"][:4], [0.0627904534, 0.125333235, 0.187380999, 0.248689592], rtol=0, atol=1e-9)
    np.testing.assert_allclose(g["filter/biquad_lpf/coeffs"], [0.00460400945, 0.00920801889, 0.00460400945, -1.79909647, 0.817512453], rtol=1e-7)
    np.testing.assert_allclose(g["filter/biquad_lpf/impulse"][:3], [0.00460400945, 0.0174910761, 0.0323083103], rtol=1e-7)
    np.testing.assert_allclose(g["filter/biquad_hpf/coeffs_f50_q1"], [0.996438146, -1.99287629, 0.996438146, -1.99285102, 0.992901564], rtol=1e-7)
    np.testing.assert_allclose(g["filter/onepole_lpf/coeffs"][[0, 3]], [0.132791519, 0.867208481], rtol=1e-7)
    a = g["adsr/a"]
    np.testing.assert_allclose([a[0], a[1], a[441], a[442], a[1000]], [0, 0.00226757373, 1, 0.99993521, 0.963782251], rtol=1e-7)
    np.testing.assert_allclose(g["pitch/frequency"][[69, 60]], [440.0, 261.625549], rtol=1e-7)


def test_mono_synth_overwrites_and_stereo_accumulates(golden):
    """mono Synth::process output == the last active voice alone; stereo sums."""
    g = golden[48000]
    voices, mix = g["synth/subtractive/voices"], g["synth/subtractive/mix"]
    assert np.array_equal(mix[0, :512], voices[-1, 0, :512])


def _preset_fixture():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "presets.json")) as f:
        return json.load(f)


def _assert_presets(eng):
    import gen_golden
    got, want = gen_golden.preset_tables(eng), _preset_fixture()
    assert set(got) == set(want)
    for k in want:
        assert [n for n, _ in got[k]] == [n for n, _ in want[k]], k
        for (_, gv), (_, wv) in zip(got[k], want[k]):
            assert np.array_equal(np.array(gv, np.float32).view(np.uint32), np.array(wv, np.float32).view(np.uint32)), k


def test_port_presets_match_the_reference_fixture():
    """Plugin::presets (klang.h:1940-1981) of every bound program, names and values, against tests/golden/presets.json (written by
    tests/gen_golden.py from the compiled reference)."""
    _assert_presets(oracle.port)


@pytest.mark.skipif(not oracle.ref.available(), reason="compiled reference not present")
def test_reference_reproduces_the_preset_fixture():
    _assert_presets(oracle.ref)


def test_on_control_fan_out_counts_the_notes_that_are_not_off():
    """Synth::onControl (klang.h:4399-4404) reaches every note whose stage is not Off: port against the compiled reference when present."""
    for eng in ([oracle.port, oracle.ref] if oracle.ref.available() else [oracle.port]):
        eng.set_fs(48000)
        eng.srand(1)
        sy = eng.Synth(cases.SY_SUBTRACTIVE, 16)
        sy.set_control(3, 0.01)
        assert sy.on_control(0, 0.5) == 0
        for v in range(5):
            sy.voice_start(v, 50 + v, 0.8)
        assert sy.on_control(0, 0.5) == 5
        sy.voice_release(1, 0.0); sy.voice_release(3, 0.0)
        sy.process(4096)                                   # the two released notes finish inside the block and stop()
        assert sy.on_control(1, 0.2) == 3
        sy.close()
