"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through the C ABI
(klang_b200.Engine / banks -> libklang_b200.so) and compares with

  * the golden vectors recorded from the compiled reference (tests/golden), and
  * the plain-C oracle (oracle/klang_port.c) run live on the same seeded inputs.

Bar (BASELINE.json north_star): 1e-5 relative fp32.  Elementwise |g-r| <= 1e-5*|r| + 1e-6*peak(r)
(SURVEY §8d); where the arithmetic is integer or transcendental-free the assertion is bit-exact."""
import ctypes as C

import numpy as np
import pytest

import cases
import klang_b200 as kb
import oracle

pytestmark = pytest.mark.gpu

RTOL, ATOL_PEAK = 1e-5, 1e-6


@pytest.fixture(scope="module")
def eng():
    if kb.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    return kb.Engine()


def assert_parity(got, want, name, exact=False):
    """NaN / inf produced by the reference (e.g. SynTHX partials above Nyquist) must appear at the same places."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    if want.dtype != np.float32:
        assert np.array_equal(got, want), f"{name}: integer mismatch"
        return 1.0
    both_nan = np.isnan(got) & np.isnan(want)
    same = (got.view(np.uint32) == want.view(np.uint32)) | both_nan
    frac = float(same.mean()) if same.size else 1.0
    if exact:
        if not same.all():
            idx = tuple(np.argwhere(~same)[0])
            raise AssertionError(f"{name}: not bit-exact at {idx}: got {got[idx]!r} want {want[idx]!r} ({(~same).sum()} of {same.size})")
        return frac
    finite = np.isfinite(want)
    peak = float(np.max(np.abs(want[finite]))) if finite.any() else 0.0
    tol = RTOL * np.abs(want) + ATOL_PEAK * peak
    with np.errstate(invalid="ignore"):
        ok = (np.abs(got.astype(np.float64) - want.astype(np.float64)) <= tol) | same | (got == want)
    bad = ~ok
    if bad.any():
        idx = tuple(np.argwhere(bad)[0])
        raise AssertionError(f"{name}: out of tolerance at {idx}: got {got[idx]!r} want {want[idx]!r} "
                             f"({bad.sum()} of {bad.size}; bit-exact fraction {frac:.6f})")
    return frac


# ------------------------------------------------------------------------------------------------ libm
def test_device_sinf_cosf_match_host_libm(eng):
    """kb_sinf/kb_cosf restate glibc's algorithm: must be bit-identical to the host libm the oracle links."""
    libm = C.CDLL("libm.so.6")
    for fn in ("sinf", "cosf"):
        getattr(libm, fn).restype, getattr(libm, fn).argtypes = C.c_float, [C.c_float]
    x = np.concatenate([
        np.linspace(0, np.pi, 60001, dtype=np.float32),                    # w = f*2pi/fs over [0, nyquist]
        cases.noise(60000, seed=11, lo=0.0, hi=7.0),
        cases.noise(20000, seed=12, lo=-100.0, hi=119.0),
        np.float32(2.0) ** np.arange(-30, 6, dtype=np.float32),
    ]).astype(np.float32)
    for fn in ("sinf", "cosf"):
        want = np.array([getattr(libm, fn)(float(v)) for v in x], np.float32)
        assert_parity(eng.math(fn, x), want, f"device {fn}", exact=True)


def test_device_tanhf_matches_host_libm(eng):
    """kb_tanhf restates glibc's FDLIBM tanhf/expm1f in fp32: bit-identical."""
    libm = C.CDLL("libm.so.6")
    libm.tanhf.restype, libm.tanhf.argtypes = C.c_float, [C.c_float]
    x = np.concatenate([cases.noise(60000, seed=13, lo=-6.0, hi=6.0), cases.noise(20000, seed=14, lo=-30.0, hi=30.0),
                        cases.noise(20000, seed=15, lo=-0.01, hi=0.01), np.float32(2.0) ** np.arange(-40, 8, dtype=np.float32)]).astype(np.float32)
    want = np.array([libm.tanhf(float(v)) for v in x], np.float32)
    assert_parity(eng.math("tanhf", x), want, "device tanhf", exact=True)


def test_device_expf_matches_host_libm(eng):
    """kb_expf restates glibc's expf (double-precision table + cubic, FMA build): bit-identical to the box's libm — exhaustively so on the build
    host (all 2^32 floats); here a million arguments over the ranges programs use and the special cases."""
    libm = C.CDLL("libm.so.6")
    libm.expf.restype, libm.expf.argtypes = C.c_float, [C.c_float]
    x = np.concatenate([cases.noise(400000, seed=21, lo=-4.0, hi=0.5), cases.noise(300000, seed=22, lo=-104.0, hi=89.0), cases.noise(200000, seed=23, lo=-0.05, hi=0.05),
                        np.float32(2.0) ** np.arange(-60, 8, dtype=np.float32), -(np.float32(2.0) ** np.arange(-60, 8, dtype=np.float32)),
                        np.array([0.0, -0.0, 88.0, 88.72, 88.73, -103.9, -104.0, float.fromhex('0x1.04845ep+5'), float.fromhex('-0x1.f8cbb2p+5'), np.inf, -np.inf], np.float32)]).astype(np.float32)
    want = np.array([libm.expf(float(v)) for v in x], np.float32)
    assert_parity(eng.math("expf", x), want, "device expf", exact=True)


# ------------------------------------------------------------------------------------------ primitives
PRIM_EXACT = ("osc/fast_", "osc/basic_sine", "osc/wt_", "filter/biquad_lpf", "filter/biquad_hpf",
              "filter/onepole_lpf/impulse", "filter/onepole_lpf/coeffs", "filter/onepole_hpf/impulse", "filter/onepole_hpf/coeffs",
              "envelope/4pt", "envelope/3pt", "envelope/loop13", "adsr/", "pitch/")


@pytest.mark.parametrize("fs", [44100, 48000])
def test_primitives_match_reference_golden(eng, golden, fs):
    eng.set_fs(fs)
    g = golden[fs]
    n = 256
    kinds = {"fast_saw": 0, "fast_triangle": 1, "fast_square": 2, "fast_pulse": 3, "fast_sine": 4, "basic_sine": 5,
             "basic_saw": 6, "basic_triangle": 7, "basic_square": 8, "basic_pulse": 9, "wt_sine": 10, "wt_saw": 11}
    checked = 0
    for name, kind in kinds.items():
        for f in (441.0, 55.0, 3520.5, 1000.0):
            assert_parity(eng.osc(kind, n, f), g[f"osc/{name}/f{f}"], f"osc/{name}/f{f}", exact=True)
        assert_parity(eng.osc(kind, n, 441.0, 1.0), g[f"osc/{name}/f441_p1"], f"osc/{name}/f441_p1", exact=True)
        checked += 5
        if kind <= 3 or kind == 9:
            for duty in (0.05, 0.5, 0.93):
                assert_parity(eng.osc(kind, n, 441.0, 0.0, duty), g[f"osc/{name}/f441_p0_d{duty}"], f"osc/{name}/d{duty}", exact=True)
                assert_parity(eng.osc(kind, n, 2093.0, 2.0, duty), g[f"osc/{name}/f2093_p2_d{duty}"], f"osc/{name}/p2 d{duty}", exact=True)
    x = cases.noise(512, seed=7)
    imp = np.zeros(64, np.float32)
    imp[0] = 1
    sweep = (200.0 + 6000.0 * (0.5 + 0.5 * np.sin(np.arange(512) * 0.01))).astype(np.float32)
    for kind, name in ((0, "biquad_lpf"), (1, "biquad_hpf")):
        y, c = eng.filt(kind, imp, 1000.0)
        assert_parity(y, g[f"filter/{name}/impulse"], name + " impulse", exact=True)
        assert_parity(c, g[f"filter/{name}/coeffs"], name + " coeffs", exact=True)
        y, c = eng.filt(kind, x, 50.0, 1.0)
        assert_parity(y, g[f"filter/{name}/noise_f50_q1"], name + " noise", exact=True)
        y, _ = eng.filt(kind, x, sweep, 10.0, per_sample=True)        # device cosf/sinf every sample
        assert_parity(y, g[f"filter/{name}/sweep_q10"], name + " sweep", exact=True)
    for kind, name in ((2, "onepole_lpf"), (3, "onepole_hpf"), (7, "butterworth_lpf1")):
        y, c = eng.filt(kind, imp, 1000.0)
        assert_parity(y, g[f"filter/{name}/impulse"], name + " impulse", exact=True)
        assert_parity(c, g[f"filter/{name}/coeffs"], name + " coeffs", exact=True)
    for kind, name in ((4, "biquad_bpf"), (5, "biquad_brf"), (8, "butterworth_lpf2")):      # SURVEY §8f-3 primitives
        y, c = eng.filt(kind, imp, 1000.0)
        assert_parity(y, g[f"filter/{name}/impulse"], name + " impulse", exact=True)
        assert_parity(c, g[f"filter/{name}/coeffs"], name + " coeffs", exact=True)
        y, _ = eng.filt(kind, x, 50.0, 1.0)
        assert_parity(y, g[f"filter/{name}/noise_f50_q1"], name + " noise", exact=True)
        y, _ = eng.filt(kind, x, sweep, 10.0, per_sample=True)
        assert_parity(y, g[f"filter/{name}/sweep_q10"], name + " sweep", exact=True)
    y, _ = eng.filt(6, x, 1000.0, 0.5)
    assert_parity(y, g["filter/biquad_apf/noise"], "biquad_apf noise", exact=True)
    for kind, name, f, Q in ((9, "dcf", 0.995, None), (9, "dcf_r09", 0.9, None), (10, "iir1", 0.25, None), (11, "iir2", -1.6, 0.8),   # klang.h:5387-5446
                             (12, "modal", 440.0, 0.25), (12, "modal_hi", 7040.0, 0.01), (13, "follower_peak", 0.01, 0.1),       # klang.h:5817-5896
                             (14, "follower_rms", 0.002, 0.05), (13, "follower_peak_instant", 0.0, 0.02)):
        y, c = eng.filt(kind, x, f, Q)
        assert_parity(y, g[f"filter/{name}/noise"], name + " noise", exact=True)
        assert_parity(c, g[f"filter/{name}/coeffs"], name + " coeffs", exact=True)
        y, _ = eng.filt(kind, imp, f, Q)
        assert_parity(y, g[f"filter/{name}/impulse"], name + " impulse", exact=True)
    y, st = eng.envelope([(0, 0), (0.001, 1), (0.003, 0.25), (0.005, 0.5)], 400)
    assert_parity(y, g["envelope/4pt"], "envelope/4pt", exact=True)
    assert_parity(st, g["envelope/4pt_stage"], "envelope/4pt_stage")
    y, st = eng.envelope([(0, 100), (0.002, 1000), (0.004, 500)], 400, release_at=120, release_time=0.002, release_level=0.0)
    assert_parity(y, g["envelope/3pt_release"], "envelope/3pt_release", exact=True)
    y, st = eng.envelope([(0, 0), (0.001, 1), (0.002, 0.5), (0.003, 0.8)], 600, loop=(1, 3))
    assert_parity(y, g["envelope/loop13"], "envelope/loop13", exact=True)
    y, st = eng.adsr(0.01, 0.1, 0.7, 0.25, 20000, release_at=6000)
    assert_parity(y, g["adsr/a"], "adsr/a", exact=True)
    assert_parity(st, g["adsr/a_stage"], "adsr/a_stage")
    y, st = eng.adsr(0.0, 0.0, 1.0, 0.001, 600, release_at=100)
    assert_parity(y, g["adsr/zero_attack"], "adsr/zero_attack", exact=True)
    y, st = eng.adsr(0.001, 0.25, 1.0, 0.5, 2000, release_at=20)
    assert_parity(y, g["adsr/early_release"], "adsr/early_release", exact=True)
    assert_parity(np.array([eng.pitch_to_frequency(p) for p in range(128)], np.float32), g["pitch/frequency"], "pitch", exact=True)


# --------------------------------------------------------------------------------------------- effects
# bit-exact: sequential fp32 arithmetic identical to the reference; pingpong* additionally evaluates sinf on the device
@pytest.mark.parametrize("fs", [44100, 48000])
@pytest.mark.parametrize("name", list(cases.FX_SCRIPTS))
def test_effects_match_reference_golden(eng, golden, fs, name):
    got = cases.run_fx_script(eng, name, fs)
    assert_parity(got, golden[fs][f"fx/{name}"], f"fx/{name}", exact=True)


# ---------------------------------------------------------------------------------------------- synths
# every graph is sequential fp32 arithmetic identical to the reference (device sinf/cosf/tanhf restate the host libm)
EXACT_SYNTHS = tuple(cases.SYNTH_SCRIPTS)


@pytest.mark.parametrize("fs", [44100, 48000])
@pytest.mark.parametrize("name", list(cases.SYNTH_SCRIPTS))
def test_synth_voices_match_reference_golden(eng, golden, fs, name):
    r = cases.run_synth_script(eng, name, fs, per_voice=True)
    exact = name in EXACT_SYNTHS
    assert_parity(r["out"], golden[fs][f"synth/{name}/voices"], f"synth/{name}/voices", exact=exact)
    assert_parity(r["stages"], golden[fs][f"synth/{name}/stages"], f"synth/{name}/stages")


@pytest.mark.parametrize("fs", [44100, 48000])
@pytest.mark.parametrize("name", list(cases.SYNTH_SCRIPTS))
def test_synth_process_matches_reference_golden(eng, golden, fs, name):
    """Synth::process block output (mono: the last active voice overwrites, Q6; SynTHX: chained buffer + tanh)."""
    r = cases.run_synth_script(eng, name, fs, per_voice=False)
    assert_parity(r["out"], golden[fs][f"synth/{name}/mix"], f"synth/{name}/mix", exact=name in EXACT_SYNTHS)


@pytest.mark.parametrize("graph", [cases.SY_SUBTRACTIVE, cases.SY_SYNTHX])
def test_note_on_off_voice_stealing_matches_reference(eng, golden, graph):
    r = cases.run_synth_noteon_script(eng, graph, 48000)
    nm = cases.SY_NAMES[graph]
    assert_parity(r["assigned"], golden[48000][f"synth/{nm}/noteon_assigned"], nm + " assigned")
    assert_parity(r["out"], golden[48000][f"synth/{nm}/noteon_mix"], nm + " noteon mix", exact=True)


# ------------------------------------------------------------------------------- live oracle, other sizes
def _drive_bank_vs_oracle(graph, instances, voices, blocks, n, fs, exact, flags=kb.PER_VOICE):
    """Same seeded note stream into `instances` oracle synths and one CUDA bank; compare every voice stream."""
    oracle.port.set_fs(fs)
    oracle.port.srand(1)
    refs = [oracle.port.Synth(graph, voices) for _ in range(instances)]
    want = []
    for b in range(blocks):
        for i, sy in enumerate(refs):
            for v in range(voices):
                g = i * voices + v
                if b == (g % 3):
                    sy.voice_start(v, cases.voice_pitch(g), cases.voice_velocity(g))
                if b == 2 + (g % 4):
                    sy.voice_release(v)
        want.append(np.stack([sy.process_voices(n)[0] for sy in refs]))
    want = np.concatenate(want, axis=-1)
    for sy in refs:
        sy.close()

    kb.lib().kb_srand(1)
    bank = kb.SynthBank(graph, instances, voices, fs, n)
    got = []
    for b in range(blocks):
        for i in range(instances):
            for v in range(voices):
                g = i * voices + v
                if b == (g % 3):
                    bank.voice_start(v, cases.voice_pitch(g), cases.voice_velocity(g), i)
                if b == 2 + (g % 4):
                    bank.voice_release(v, 0.0, i)
        got.append(bank.process_block(n, flags))
    bank.close()
    got = np.concatenate(got, axis=-1)
    return assert_parity(got, want, f"bank graph {graph}", exact=exact)


def test_subtractive_bank_vs_live_oracle(eng):
    _drive_bank_vs_oracle(cases.SY_SUBTRACTIVE, 3, 128, 7, 1000, 48000, exact=True)   # ragged block, 384 voices


def test_supersaw_bank_vs_live_oracle(eng):
    # rand() draws happen per note-on in call order: the bank and the oracle both start voices instance-major
    _drive_bank_vs_oracle(cases.SY_SUPERSAW, 2, 32, 6, 777, 48000, exact=True)


def test_tb303_bank_vs_live_oracle(eng):
    _drive_bank_vs_oracle(cases.SY_TB303, 2, 64, 6, 512, 48000, exact=True)


def test_fm_bank_vs_live_oracle(eng):
    """FM.k (three Operator<Sine> in series, SURVEY 8f widening): 4 x 64 voices against the live oracle, bit-exact."""
    _drive_bank_vs_oracle(cases.SY_FM, 4, 64, 5, 300, 48000, exact=True)


def test_synthx_bank_vs_live_oracle(eng):
    _drive_bank_vs_oracle(cases.SY_SYNTHX, 2, 32, 4, 300, 48000, exact=True)


def test_effect_bank_instances_are_independent(eng):
    """64 instances with different inputs == 64 runs of the oracle (C4 shape, short)."""
    fs, n, blocks, inst = 48000, 512, 3, 8
    for graph in (cases.FX_PINGPONG, cases.FX_REVERB, cases.FX_DELAY_PINGPONG):
        oracle.port.set_fs(fs)
        x = np.stack([cases.fx_input(2, n * blocks, seed=10 + i) for i in range(inst)])
        want = np.empty_like(x)
        for i in range(inst):
            fx = oracle.port.Fx(graph)
            fx.set_control(0, 0.3 + 0.05 * i)
            for b in range(blocks):
                want[i, :, b * n:(b + 1) * n] = fx.process(x[i, :, b * n:(b + 1) * n])
            fx.close()
        bank = kb.FxBank(graph, inst, fs, n)
        for i in range(inst):
            bank.set_control(0, 0.3 + 0.05 * i, i)
        got = np.empty_like(x)
        for b in range(blocks):
            blk = np.ascontiguousarray(x[:, :, b * n:(b + 1) * n])
            bank.process_inplace(blk)
            got[:, :, b * n:(b + 1) * n] = blk
        bank.close()
        assert_parity(got, want, f"fx bank graph {graph}", exact=True)


# ------------------------------------------------------------------------- full-size properties (BASELINE sizes)
def test_gain_full_block_linearity_and_exactness(eng):
    """C1: 1 channel x 4096; out == in * gain exactly, in place; empty and ragged blocks."""
    bank = kb.FxBank(kb.FX_GAIN, 1, 48000, 4096)
    x = np.sin(0.01 * np.arange(4096)).astype(np.float32)
    y = x.copy().reshape(1, 1, -1)
    bank.process_inplace(y)
    assert np.array_equal(y[0, 0], x * np.float32(0.5))
    bank.set_control(0, 2.0)           # clamps to 1.0 (Dial max)
    assert bank.get_control(0) == 1.0
    y2 = x[:1001].copy().reshape(1, 1, -1)
    bank.process_inplace(y2)
    assert np.array_equal(y2[0, 0], x[:1001])
    bank.process_inplace(np.zeros((1, 1, 0), np.float32))   # empty block is a no-op
    with pytest.raises(kb.KlangB200Error):
        bank.process_inplace(np.zeros((1, 1, 4097), np.float32))
    bank.close()


def test_subtractive_1024_voices_block_split_invariance(eng):
    """C2 at full size: 8 x 128 voices.  Processing 4096 samples as one block or as 4 x 1024 is bit-identical
    (state carried in HBM between calls), and the MIX_SUM output equals the voice-order fp32 sum of the streams."""
    fs = 48000

    def run(blocks, n):
        kb.lib().kb_srand(1)
        bank = kb.SynthBank(kb.SY_SUBTRACTIVE, 8, 128, fs, 4096)
        for g in range(1024):
            bank.voice_start(g % 128, cases.voice_pitch(g), cases.voice_velocity(g), g // 128)
        outs = []
        for b in range(blocks):
            if b * n == 2048:
                for g in range(0, 1024, 2):
                    bank.voice_release(g % 128, 0.0, g // 128)
            outs.append(bank.process_block(n, kb.PER_VOICE))
        mix = bank.process_block(256, kb.MIX_SUM)
        voices = None
        bank.close()
        return np.concatenate(outs, axis=-1), mix

    a, mix_a = run(1, 4096)
    # release events are block-granular, so compare the event-free prefix of the split run as well as the rest
    b, mix_b = run(4, 1024)
    assert np.array_equal(a[..., :2048].view(np.uint32), b[..., :2048].view(np.uint32))
    assert np.isfinite(a).all() and np.abs(a).max() > 0.01

    kb.lib().kb_srand(1)
    bank = kb.SynthBank(kb.SY_SUBTRACTIVE, 8, 128, fs, 4096)
    for g in range(1024):
        bank.voice_start(g % 128, cases.voice_pitch(g), cases.voice_velocity(g), g // 128)
    v = bank.process_block(512, kb.PER_VOICE)
    bank.close()
    kb.lib().kb_srand(1)
    bank = kb.SynthBank(kb.SY_SUBTRACTIVE, 8, 128, fs, 4096)
    for g in range(1024):
        bank.voice_start(g % 128, cases.voice_pitch(g), cases.voice_velocity(g), g // 128)
    m = bank.process_block(512, kb.MIX_SUM)
    bank.close()
    acc = np.zeros((8, 1, 512), np.float32)
    for k in range(128):
        acc = acc + v[:, k]
    assert np.array_equal(acc.view(np.uint32), m.view(np.uint32))


@pytest.mark.parametrize("graph,inst,voices", [(cases.SY_SUBTRACTIVE, 8, 128), (cases.SY_SUPERSAW, 8, 32), (cases.SY_TB303, 4, 128), (cases.SY_FILTER_K, 2, 32)])
def test_tiled_schedule_equals_lane_per_voice_schedule(eng, graph, inst, voices):
    """The tiled kernels reorder independent samples only: bit-identical streams and stages at BASELINE sizes,
    ragged block lengths included (n = 4096, then 1000, then 1)."""
    outs = {}
    for flag in (0, kb.LANE_PER_VOICE):
        kb.lib().kb_srand(1)
        bank = kb.SynthBank(graph, inst, voices, 48000, 4096)
        if graph == cases.SY_TB303:
            bank.set_control(3, 1.0, instance=1)           # one instance on the square oscillator
            bank.set_control(1, 0.8, instance=2)
        for g in range(inst * bank.voices):
            if g % 5 != 4:                                  # leave some voices Off
                bank.voice_start(g % bank.voices, cases.voice_pitch(g), cases.voice_velocity(g), g // bank.voices)
        res = []
        for i, n in enumerate((4096, 1000, 1, 129)):
            if i == 1:
                for g in range(0, inst * bank.voices, 3):
                    bank.voice_release(g % bank.voices, 0.0, g // bank.voices)
            res.append(bank.process_block(n, kb.PER_VOICE | flag))
        stages = [bank.voice_stage(v, i) for i in range(inst) for v in range(bank.voices)]
        bank.close()
        outs[flag] = (np.concatenate(res, axis=-1), np.array(stages))
    a, b = outs[0], outs[kb.LANE_PER_VOICE]
    assert_parity(a[0], b[0], f"tiled vs lane-per-voice graph {graph}", exact=True)
    assert np.array_equal(a[1], b[1])
    assert np.abs(a[0]).max() > 0.01


@pytest.mark.parametrize("graph,blocks,n", [(cases.FX_PINGPONG, 30, 2048), (cases.FX_REVERB, 12, 1000), (cases.FX_DELAY_PINGPONG, 10, 3000),
                                            (cases.FX_DELAY_REVERB, 8, 5000)])
def test_chunk_parallel_effects_match_oracle_and_sequential_schedule(eng, graph, blocks, n):
    """The chunk-parallel effect kernels (engaged once control smoothers settle / delays allow) are bit-identical to the
    oracle and to the frame-sequential schedule, including the hand-over between the two schedules, ragged block lengths
    and a control change in the middle (which sends PingPong back to the sequential schedule while its smoothers move)."""
    fs, inst = 48000, 4
    oracle.port.set_fs(fs)
    lens = [n] * blocks
    lens[3] = 517
    lens[5] = 1
    total = sum(lens)
    mono = graph == cases.FX_DELAY_REVERB
    x = np.stack([cases.fx_input(1 if mono else 2, total, seed=20 + i) for i in range(inst)])
    change_at = blocks // 2

    def controls_for(b, set_control):
        if b == 0 and graph == cases.FX_PINGPONG:
            set_control(0, 0.8)
        if b == change_at:
            if graph == cases.FX_PINGPONG:
                set_control(1, 0.3)       # new delay target: smoothers move again
                set_control(5, 0.3)
            elif graph == cases.FX_REVERB:
                set_control(6, 0.5)       # room size: early taps and all 16 delay times are re-drawn
                set_control(2, 0.4)
            elif graph == cases.FX_DELAY_REVERB:
                set_control(1, 0.07)      # loop delay
                set_control(2, 900.0)     # damping cutoff: prepare() recomputes the LPF
            else:
                set_control(1, 0.31)

    want = np.empty_like(x)
    for i in range(inst):
        fx = oracle.port.Fx(graph)
        o = 0
        for b, ln in enumerate(lens):
            controls_for(b, fx.set_control)
            want[i, :, o:o + ln] = fx.process(x[i, 0, o:o + ln] if mono else x[i, :, o:o + ln])
            o += ln
        fx.close()
    for flags in (0, kb.FX_SEQUENTIAL):
        bank = kb.FxBank(graph, inst, fs, max(lens))
        got = np.empty_like(x)
        o, engaged = 0, 0
        for b, ln in enumerate(lens):
            controls_for(b, bank.set_control)
            blk = np.ascontiguousarray(x[:, :, o:o + ln])
            bank.process_inplace(blk, flags=flags)
            got[:, :, o:o + ln] = blk
            engaged += bank.parallel_instances() if flags == 0 else 0
            o += ln
        bank.close()
        assert_parity(got, want, f"fx graph {graph} flags {flags}", exact=True)
        if flags == 0:
            assert engaged > inst * blocks // 4, f"chunk-parallel schedule engaged on only {engaged} instance-blocks"


@pytest.mark.parametrize("fs", [44100, 48000, 96000, 192000])      # (at 192 kHz the live delay spans exceed shared memory: the round-1 pipeline takes those instances)
def test_reverb_pipeline_chunk_boundaries(eng, fs):
    """Reverb.k on the pipelined chunk schedule with every bus audible (direct, early, mid, late all non-zero), block
    lengths straddling 1, 2, 3 and many pipeline chunks (prologue-only, one-iteration and steady-state paths), bit-exact
    against the live oracle; KB_FX_SEQUENTIAL blocks are interleaved so the two schedules hand the state to each other."""
    oracle.port.set_fs(fs)
    inst = 3
    lens = [1, 2, 7, 8, 9, 43, 44, 45, 46, 47, 48, 49, 50, 51, 87, 88, 89, 95, 96, 97, 143, 144, 145, 60, 69, 70, 74, 75, 76, 80, 81, 91, 92, 93, 99, 100, 101, 138, 139, 149, 150, 151, 160, 161, 225, 226,
            240, 241, 1023, 4096, 333, 5000]
    total = sum(lens)
    x = np.stack([cases.fx_input(2, total, seed=40 + i) for i in range(inst)])
    settings = {0: 0.3, 1: 0.9, 2: 0.4, 3: 0.5, 4: 0.8}
    want = np.empty_like(x)
    for i in range(inst):
        fx = oracle.port.Fx(cases.FX_REVERB)
        for c, v in settings.items():
            fx.set_control(c, v)
        o = 0
        for ln in lens:
            want[i, :, o:o + ln] = fx.process(x[i, :, o:o + ln])
            o += ln
        fx.close()
    bank = kb.FxBank(kb.FX_REVERB, inst, fs, max(lens))
    for c, v in settings.items():
        bank.set_control(c, v)
    got = np.empty_like(x)
    o, engaged = 0, 0
    for b, ln in enumerate(lens):
        blk = np.ascontiguousarray(x[:, :, o:o + ln])
        seq = (b % 5 == 3)
        bank.process_inplace(blk, flags=kb.FX_SEQUENTIAL if seq else 0)
        engaged += 0 if seq else bank.parallel_instances()
        got[:, :, o:o + ln] = blk
        o += ln
    bank.close()
    assert_parity(got, want, f"reverb pipeline fs {fs}", exact=True)
    assert np.abs(want[:, 0]).max() > 0.01
    assert engaged >= inst * (len(lens) - len(lens) // 5 - 1)


# ------------------------------------------------------------------------- BASELINE-size properties, C3 / C4 / C5
def test_c4_64_instances_are_replicas_of_one_instance(eng):
    """C4 shape: 64 stereo instances fed the same input and controls produce 64 identical streams, equal to a
    single-instance bank (instances never interact), on the chunk-parallel schedule, for every delay-line graph."""
    fs, n, blocks = 48000, 4096, 3
    for graph in (kb.FX_REVERB, kb.FX_DELAY_PINGPONG, kb.FX_DELAY_REVERB, kb.FX_PINGPONG):
        ch = 1 if graph == kb.FX_DELAY_REVERB else 2
        x = cases.fx_input(ch, n * blocks, seed=5)
        one = kb.FxBank(graph, 1, fs, n)
        many = kb.FxBank(graph, 64, fs, n)
        for b in range(blocks):
            blk1 = np.ascontiguousarray(x[None, :, b * n:(b + 1) * n])
            blk64 = np.ascontiguousarray(np.broadcast_to(blk1, (64, ch, n)))
            one.process_inplace(blk1)
            many.process_inplace(blk64)
            assert np.array_equal(blk64.view(np.uint32), np.broadcast_to(blk1, (64, ch, n)).view(np.uint32)), f"graph {graph} block {b}"
        one.close()
        many.close()


def test_c3_c5_bank_mix_equals_sum_of_instance_outputs(eng):
    """KB_BANK_MIX (the multi-GPU reduce input) is the instance-order fp32 sum of the per-instance Synth::process outputs,
    at the C3 (8 x 32 SuperSaw) and C5 per-GPU (4 x 128 TB303, 4 x 128 SynTHX) sizes."""
    for graph, inst, voices, n in ((kb.SY_SUPERSAW, 8, 32, 2048), (kb.SY_TB303, 4, 128, 1024), (kb.SY_SYNTHX, 4, 128, 256)):
        outs = {}
        for flags in (kb.MIX_SUM, kb.MIX_SUM | kb.BANK_MIX):
            kb.lib().kb_srand(1)
            bank = kb.SynthBank(graph, inst, voices, 48000, n)
            for g in range(inst * voices):
                bank.voice_start(g % voices, 36 + (5 * g) % 36, cases.voice_velocity(g), g // voices)
            bank.process_block(n, flags)
            outs[flags] = bank.process_block(n, flags)
            bank.close()
        per_inst, mix = outs[kb.MIX_SUM], outs[kb.MIX_SUM | kb.BANK_MIX]
        acc = np.zeros_like(mix)
        for i in range(inst):
            acc = acc + per_inst[i]
        assert_parity(mix, acc, f"bank mix graph {graph}", exact=True)
        assert np.isfinite(mix).all() and np.abs(mix).max() > 1e-3


def test_batched_events_equal_single_calls(eng):
    """kb_synth_bank_events applies a block's events in order exactly like the individual calls (incl. rand() draws)."""
    n, voices = 512, 32
    ev = np.zeros(24, kb.EVENT_DTYPE)
    for k in range(16):
        ev[k] = (kb.EV_NOTE_ON, k % 2, 48 + k, 0.0, 0.5 + 0.03 * k)
    for k in range(16, 20):
        ev[k] = (kb.EV_NOTE_OFF, k % 2, 48 + (k - 16), 0.0, 0.0)
    ev[20] = (kb.EV_CONTROL, 0, 2, 0.0, 0.9)
    ev[21] = (kb.EV_VOICE_START, 1, 7, 61.5, 0.8)
    ev[22] = (kb.EV_VOICE_RELEASE, 1, 7, 0.0, 0.0)
    ev[23] = (kb.EV_NOTE_ON, 0, 72, 0.0, 1.0)
    outs = []
    for batched in (True, False):
        kb.lib().kb_srand(7)
        bank = kb.SynthBank(kb.SY_SUPERSAW, 2, voices, 48000, n)
        res = []
        for b in range(3):
            if batched:
                bank.events(ev if b == 0 else ev[16:20])
            else:
                for e in (ev if b == 0 else ev[16:20]):
                    t, i, key, pitch, vel = int(e["type"]), int(e["instance"]), int(e["key"]), float(e["pitch"]), float(e["velocity"])
                    if t == kb.EV_NOTE_ON:
                        bank.note_on(key, vel, i)
                    elif t == kb.EV_NOTE_OFF:
                        bank.note_off(key, vel, i)
                    elif t == kb.EV_CONTROL:
                        bank.set_control(key, vel, i)
                    elif t == kb.EV_VOICE_START:
                        bank.voice_start(key, pitch, vel, i)
                    else:
                        bank.voice_release(key, vel, i)
            res.append(bank.process_block(n, kb.PER_VOICE))
        bank.close()
        outs.append(np.concatenate(res, axis=-1))
    assert_parity(outs[0], outs[1], "batched events", exact=True)
    assert np.abs(outs[0]).max() > 0.01


def test_device_pointer_calls_match_host_pointer_calls(eng):
    """KB_DEVICE_PTR (asynchronous, caller-owned device buffers and stream) returns the same bits as the host-buffer call."""
    torch = pytest.importorskip("torch")
    n = 1024
    x = cases.fx_input(2, n, seed=9)[None]
    host = kb.FxBank(kb.FX_REVERB, 1, 48000, n)
    a = np.ascontiguousarray(x)
    host.process_inplace(a)
    host.close()
    dev = kb.FxBank(kb.FX_REVERB, 1, 48000, n)
    s = torch.cuda.Stream()
    dev.set_stream(s.cuda_stream)
    with torch.cuda.stream(s):
        t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        s.synchronize()
        dev.process_inplace(t)
        dev.sync()
    b = t.cpu().numpy()
    dev.close()
    assert_parity(b, a, "device pointer fx", exact=True)


def test_odd_bank_shapes_vs_live_oracle(eng):
    """Instance / voice counts that are not multiples of the kernels' voices-per-CTA, and a block that is not a multiple
    of the 128-sample tile."""
    _drive_bank_vs_oracle(cases.SY_SUBTRACTIVE, 5, 33, 5, 333, 48000, exact=True)
    _drive_bank_vs_oracle(cases.SY_TB303, 3, 37, 5, 200, 44100, exact=True)
    _drive_bank_vs_oracle(cases.SY_SUPERSAW, 3, 35, 4, 130, 48000, exact=True)
    # the 8-voice layout-2 Subtractive kernel with a partial last CTA (819 voices) and the 16-voice kernel (1625 voices)
    _drive_bank_vs_oracle(cases.SY_SUBTRACTIVE, 7, 117, 3, 333, 48000, exact=True)
    _drive_bank_vs_oracle(cases.SY_SUBTRACTIVE, 13, 125, 2, 200, 44100, exact=True)


def test_pingpong_long_blocks_span_sub_blocks(eng):
    """A 20000-frame process call is served as 8192-frame sub-blocks (the staged block lives in shared memory); the result
    equals the oracle fed the same frames in its own block sizes."""
    fs, inst = 48000, 2
    oracle.port.set_fs(fs)
    total = 60000
    x = np.stack([cases.fx_input(2, total, seed=40 + i) for i in range(inst)])
    want = np.empty_like(x)
    for i in range(inst):
        fx = oracle.port.Fx(cases.FX_PINGPONG)
        for o in range(0, total, 10000):
            want[i, :, o:o + 10000] = fx.process(x[i, :, o:o + 10000])
        fx.close()
    bank = kb.FxBank(cases.FX_PINGPONG, inst, fs, 20000)
    got = np.empty_like(x)
    for o in range(0, total, 20000):
        blk = np.ascontiguousarray(x[:, :, o:o + 20000])
        bank.process_inplace(blk)
        got[:, :, o:o + 20000] = blk
    assert bank.parallel_instances() == inst          # the last sub-block ran on the chunk-parallel schedule
    bank.close()
    assert_parity(got, want, "pingpong long blocks", exact=True)


def test_async_calls_with_events_every_block_match_synchronous_calls(eng):
    """A caller that runs AHEAD of the device — note events before every block, device-pointer outputs that are never joined,
    and page-locked host outputs with KB_ASYNC_HOST — gets the same bits as block-by-block synchronous calls: the packed
    event uploads go through a ring of pinned staging slots, so a queued copy is never overwritten by the next block's events."""
    torch = pytest.importorskip("torch")
    fs, n, inst, voices, blocks = 48000, 2048, 8, 128, 24
    total = inst * voices

    def schedule(b, bank):
        ev = np.zeros(len(range(b % 16, total, 16)), kb.EVENT_DTYPE)
        ids = np.arange(b % 16, total, 16)
        ev["type"] = kb.EV_VOICE_START
        ev["instance"], ev["key"] = ids // voices, ids % voices
        ev["pitch"] = [cases.voice_pitch(int(g) + b) for g in ids]
        ev["velocity"] = [cases.voice_velocity(int(g)) for g in ids]
        bank.events(ev)

    def make():
        kb.lib().kb_srand(1)
        bank = kb.SynthBank(kb.SY_SUBTRACTIVE, inst, voices, fs, n)
        for g in range(total):
            bank.voice_start(g % voices, cases.voice_pitch(g), cases.voice_velocity(g), g // voices)
        return bank

    ref = make()
    want = []
    for b in range(blocks):
        schedule(b, ref)
        want.append(ref.process_block(n, kb.MIX_SUM))
    ref.close()

    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    # (a) device-pointer outputs, nothing joined until the end
    bank = make()
    bank.set_stream(ts.cuda_stream)
    outs = [torch.empty(bank.out_shape(n), dtype=torch.float32, device="cuda") for _ in range(blocks)]
    for b in range(blocks):
        schedule(b, bank)
        bank.process_into(outs[b], n, kb.MIX_SUM)
    torch.cuda.synchronize()
    for b in range(blocks):
        assert np.array_equal(outs[b].cpu().numpy().view(np.uint32), want[b].view(np.uint32)), f"device-pointer block {b}"
    bank.close()
    # (b) page-locked host outputs with KB_ASYNC_HOST
    bank = make()
    bank.set_stream(ts.cuda_stream)
    houts = [torch.empty(bank.out_shape(n), dtype=torch.float32).pin_memory() for _ in range(blocks)]
    for b in range(blocks):
        schedule(b, bank)
        bank.process_into(houts[b].numpy(), n, kb.MIX_SUM | kb.ASYNC_HOST)
    bank.sync()
    for b in range(blocks):
        assert np.array_equal(houts[b].numpy().view(np.uint32), want[b].view(np.uint32)), f"async host block {b}"
    bank.close()
    assert np.abs(want[-1]).max() > 0.01
