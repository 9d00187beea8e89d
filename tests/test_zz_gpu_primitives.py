"""Primitives added after the round's last full GPU run (sorted last).  Device-side libc rand(): Basic::Noise / Fast::Noise ticks produced on the GPU continue the process-wide rand() stream bit for bit
and hand it back advanced (klang_b200/csrc/kb_rand.h).  The generator, the jump-ahead and the libc hand-over are
pinned on the CPU (tests/host/rand_check.cpp); this file checks the device kernel against the golden vectors of the compiled
reference and the host/device interleaving against the live oracle.  (Sorted last: written after the round's last full GPU run.)"""
import numpy as np
import pytest

import cases
import klang_b200 as kb
import oracle

pytestmark = pytest.mark.gpu


def _exact(got, want, name):
    assert got.shape == want.shape, name
    same = got.view(np.uint32) == want.view(np.uint32)
    assert same.all(), f"{name}: first mismatch at {int(np.argmin(same))}: got {got.reshape(-1)[np.argmin(same)]!r} want {want.reshape(-1)[np.argmin(same)]!r}"


@pytest.mark.parametrize("fs", [44100, 48000])
def test_device_noise_matches_reference_golden(golden, fs):
    if kb.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    eng = kb.Engine()
    eng.set_fs(fs)
    got = cases.noise_cases(eng)
    assert len(got) == 4
    for name, arr in got.items():
        _exact(arr, golden[fs][name], name)
        assert arr.min() < -0.9 and arr.max() > 0.9


def test_host_draws_after_device_noise_continue_the_stream():
    """srand(11); 300 noise ticks on the device; then a SuperSaw note-on, whose on() draws seven detune factors with libc rand() on the
    host (SuperSaw.k:17): the voice must be the one the reference gets, i.e. libc was advanced by exactly the 300 device draws."""
    fs, n = 48000.0, 128
    oracle.port.set_fs(fs)
    oracle.port.srand(11)
    wn = oracle.port.osc(cases.OSC_BASIC_NOISE, 300, 0.0)
    sy = oracle.port.Synth(cases.SY_SUPERSAW, 4)
    sy.voice_start(0, 60, 0.8)
    want = sy.process_voices(n)[0]
    wn2 = oracle.port.osc(cases.OSC_FAST_NOISE, 50, 0.0)
    sy.close()

    eng = kb.Engine()
    eng.set_fs(fs)
    eng.srand(11)
    gn = eng.osc(cases.OSC_BASIC_NOISE, 300, 0.0)
    bank = kb.SynthBank(kb.SY_SUPERSAW, 1, 4, fs, n)
    bank.voice_start(0, 60, 0.8, 0)
    got = bank.process_block(n, kb.PER_VOICE)[0]
    gn2 = eng.osc(cases.OSC_FAST_NOISE, 50, 0.0)
    bank.close()
    _exact(gn, wn, "noise before the note")
    _exact(np.ascontiguousarray(got), np.ascontiguousarray(want), "SuperSaw voice after device noise")
    _exact(gn2, wn2, "noise after the note")
    assert np.abs(want).max() > 0.01


@pytest.mark.parametrize("fs", [48000])
def test_device_delay_primitive_matches_reference_golden(golden, fs):
    """One Delay<1000> on the device, sample by sample (kb_prim_delay): tap(int), tap(float), set() + process() and the third-order
    Delay::lagrange (klang.h:3405-3489) against the compiled reference's golden vectors, bit for bit.  (The effect graphs already
    exercise write / tap / process; lagrange() is used by none of them.)"""
    eng = kb.Engine()
    eng.set_fs(fs)
    g = golden[fs]
    n = 1500
    xin = (np.arange(n) + 1).astype(np.float32)
    di = (np.arange(n) * 7 % 900).astype(np.int32)
    df = cases.noise(n, seed=3, lo=0.0, hi=998.0).astype(np.float32)
    set_at = np.full(n, -1.0, np.float32)
    set_at[10], set_at[700], set_at[1200] = 4.0, 333.25, 999.5
    oi, of, op = eng.delay1000(xin, di, df, set_at)
    _exact(oi, g["delay/tap_int"], "delay/tap_int")
    _exact(of, g["delay/tap_float"], "delay/tap_float")
    _exact(op, g["delay/process"], "delay/process")
    _exact(eng.delay_lagrange(cases.noise(n, seed=6), df), g["delay/lagrange"], "delay/lagrange")


@pytest.mark.parametrize("fs", [44100, 48000])
def test_device_window_follower_matches_reference_golden(golden, fs):
    """Envelope::Follower::Window<64> mean / rms (klang.h:5904-5948; kb_prim_filter kinds 15 / 16): the moving sum is a double on the
    device as in the reference; bit-exact against the compiled reference's golden vectors."""
    eng = kb.Engine()
    eng.set_fs(fs)
    x = cases.noise(512, seed=7)
    imp = np.zeros(64, np.float32)
    imp[0] = 1
    got = cases.window_follower_cases(eng, x, imp)
    assert len(got) == 9
    for name, arr in got.items():
        _exact(np.asarray(arr, np.float32), golden[fs][name], name)


# ---- examples/Subtractive/{Breakpoint,Ramp,Release}.k: KB_SY_BREAKPOINT / KB_SY_RAMP / KB_SY_RELEASE (a Fast::Sine times one envelope) on the
# lane-per-voice kernel; the voice functions are proven against the oracle with g++ (tests/host/senv_check.cpp).  Likewise Additive/{Saw,Square}.k
# (tests/host/add_check.cpp) and Modulation/{AM,FM,FM2}.k (tests/host/smod_check.cpp): every script of cases.SYNTH_SCRIPTS_LATE.
@pytest.mark.parametrize("fs", [44100, 48000])
@pytest.mark.parametrize("name", list(cases.SYNTH_SCRIPTS_LATE))
def test_sine_envelope_synths_match_reference_golden(golden, fs, name):
    eng = kb.Engine()
    r = cases.run_synth_script(eng, name, fs, per_voice=True)
    _exact(np.ascontiguousarray(r["out"]), golden[fs][f"synth/{name}/voices"], f"synth/{name}/voices")
    assert np.array_equal(r["stages"], golden[fs][f"synth/{name}/stages"]), f"synth/{name}/stages"
    r = cases.run_synth_script(eng, name, fs, per_voice=False)
    _exact(np.ascontiguousarray(r["out"]), golden[fs][f"synth/{name}/mix"], f"synth/{name}/mix")


@pytest.mark.parametrize("prog", ["breakpoint", "ramp", "release", "pan", "rm", "tremolo", "clipping", "echo", "feedback", "additive_saw", "additive_square", "am", "mod_fm", "mod_fm2", "functions", "mute", "additive_nyquist", "iir", "wahwah", "flanger", "moddelay",
                                  "mod_chorus"])
def test_late_k_programs_run_unmodified_on_the_device(prog, tmp_path):
    """examples/Subtractive/{Breakpoint,Ramp,Release}.k, Gain/{Pan,RM,Tremolo}.k, Distortion/Clipping.k, Delay/{Echo,Feedback}.k Additive/{Saw,Square}.k and Modulation/{AM,FM,FM2}.k, Distortion/{Functions,Mute}.k compiled UNMODIFIED against
    include/compat/klang.h (tools/k_host.cpp) and run on the device through the host program: bit-identical to the oracle run with
    the same script."""
    from test_k_programs import run_k_program_on_device
    run_k_program_on_device(prog, tmp_path)


# ---- examples/Gain/{Pan,RM,Tremolo}.k, Distortion/Clipping.k: KB_FX_PAN / RM / TREMOLO / CLIPPING on the elementwise streaming kernel
# (kb_elementwise_kernel); the per-sample function and the per-block steps are proven with g++ against the golden vectors
# (tests/host/ew_check.cpp).  examples/Delay/{Echo,Feedback}.k: KB_FX_ECHO / KB_FX_FEEDBACK on the frame-sequential kernel.
@pytest.mark.parametrize("fs", [44100, 48000])
@pytest.mark.parametrize("name", list(cases.FX_SCRIPTS_LATE))
def test_late_effects_match_reference_golden(golden, fs, name):
    eng = kb.Engine()
    got = cases.run_fx_script(eng, name, fs)
    _exact(np.ascontiguousarray(got), golden[fs][f"fx/{name}"], f"fx/{name}")


@pytest.mark.parametrize("graph", [cases.FX_PAN, cases.FX_RM, cases.FX_TREMOLO, cases.FX_CLIPPING, cases.FX_ECHO, cases.FX_FEEDBACK, cases.FX_FUNCTIONS,
                                   cases.FX_MUTE, cases.FX_IIR, cases.FX_WAHWAH, cases.FX_FLANGER, cases.FX_MODDELAY, cases.FX_MOD_CHORUS])
def test_late_effect_bank_vs_live_oracle(graph):
    """Five instances with different controls, ragged blocks (1001 frames: rows that are not 16-byte aligned take the scalar path of
    the streaming kernel, 1024 the vector path), a control change between blocks: every instance equals its own oracle object.
    Echo.k / Feedback.k (KB_FX_ECHO / KB_FX_FEEDBACK, one Delay<192000> per instance in the HBM ring arena, frame-sequential kernel;
    frame functions proven with g++, tests/host/onedelay_check.cpp): delay times of a few hundred frames so the echoes land."""
    fs, inst = 48000.0, 5
    oracle.port.set_fs(fs)
    refs = [oracle.port.Fx(graph) for _ in range(inst)]
    bank = kb.FxBank(graph, inst, fs, 1024)
    ch = bank.channels
    lo, hi = {cases.FX_PAN: (0.0, 1.0), cases.FX_RM: (1.0, 1000.0), cases.FX_TREMOLO: (1.0, 10.0), cases.FX_CLIPPING: (1.0, 11.0),
              cases.FX_ECHO: (0.0, 0.02), cases.FX_FEEDBACK: (0.0, 0.02), cases.FX_FUNCTIONS: (1.0, 25.0), cases.FX_MUTE: (0.0, 1.0),
              cases.FX_IIR: (0.0, 1.0), cases.FX_WAHWAH: (10.0, 10000.0), cases.FX_FLANGER: (0.1, 1.0), cases.FX_MODDELAY: (1.0, 10.0),
              cases.FX_MOD_CHORUS: (1.0, 10.0)}[graph]
    for i in range(inst):
        v = lo + (hi - lo) * (i + 0.5) / inst
        refs[i].set_control(0, v)
        bank.set_control(0, v, i)
    for b, n in enumerate((1001, 1024, 7, 1024)):
        if b == 2:
            refs[3].set_control(0, hi)
            bank.set_control(0, hi, 3)
            if graph in (cases.FX_RM, cases.FX_TREMOLO, cases.FX_ECHO, cases.FX_FEEDBACK):
                refs[1].set_control(1, 0.1)
                bank.set_control(1, 0.1, 1)
        x = np.stack([cases.fx_input(ch, n, seed=10 * b + i) for i in range(inst)])          # [inst, ch, n]
        want = np.stack([np.atleast_2d(refs[i].process(x[i][0] if ch == 1 else x[i])) for i in range(inst)])
        got = bank.process_inplace(np.ascontiguousarray(x, np.float32).copy())
        _exact(np.ascontiguousarray(got), np.ascontiguousarray(want, np.float32), f"graph {graph} block {b}")
    bank.close()
    for r in refs:
        r.close()


@pytest.mark.parametrize("graph", [cases.SY_ADDITIVE_SAW, cases.SY_ADDITIVE_SQUARE, cases.SY_ADDITIVE_NYQUIST])
def test_additive_time_parallel_kernel_equals_lane_per_voice_kernel(graph):
    """Additive/Saw.k / Square.k: kb_additive_kernel (thread = (voice, sample), the default) against the lane-per-voice kernel
    (KB_LANE_PER_VOICE), bit for bit over ragged blocks with re-triggers (the partials keep their phase) and cut notes; 3 x 21 voices."""
    inst, voices, fs = 3, 21, 48000.0
    outs = []
    for flag in (kb.LANE_PER_VOICE, 0):
        bank = kb.SynthBank(graph, inst, voices, fs, 1024)
        res = []
        for b, n in enumerate((1024, 117, 1, 700, 1024)):
            for g in range(inst * bank.voices):
                if b == g % 3:
                    bank.voice_start(g % bank.voices, 30 + (7 * g) % 70, 0.8, g // bank.voices)
                if b == 2 + g % 2 and g % 4 == 0:
                    bank.voice_release(g % bank.voices, 0.0, g // bank.voices)
                if b == 3 and g % 5 == 0:
                    bank.voice_start(g % bank.voices, 90 - g % 40, 0.5, g // bank.voices)
            res.append(bank.process_block(n, kb.PER_VOICE | flag))
        res.append(bank.process_block(512, kb.PER_VOICE | flag))
        bank.close()
        outs.append(np.concatenate(res, axis=-1))
    _exact(np.ascontiguousarray(outs[1]), np.ascontiguousarray(outs[0]), f"additive graph {graph}")
    assert np.abs(outs[0]).max() > 0.5


@pytest.mark.parametrize("graph", [cases.SY_BREAKPOINT, cases.SY_RAMP, cases.SY_RELEASE, cases.SY_AM])
def test_one_envelope_time_parallel_kernel_equals_lane_per_voice_kernel(graph):
    """kb_esine_tiled_kernel (one envelope lane per voice + a thread per (voice, sample); the default for Breakpoint.k / Ramp.k /
    Release.k / Modulation/AM.k) against the lane-per-voice kernel (KB_LANE_PER_VOICE): per-voice streams, the instance mix and the note
    stages, bit for bit, over ragged blocks with releases, re-triggers and control changes; 3 x 21 voices (ragged last CTA)."""
    inst, voices, fs = 3, 21, 48000.0
    outs = []
    for flag in (kb.LANE_PER_VOICE, 0):
        bank = kb.SynthBank(graph, inst, voices, fs, 4096)
        if graph == cases.SY_RELEASE:
            bank.set_control(3, 0.02)
        res = []
        for b, n in enumerate((4096, 117, 1, 128, 129, 1000, 4096, 300)):
            for g in range(inst * bank.voices):
                i, v = g // bank.voices, g % bank.voices
                if b == 0 and g % 2 == 0:
                    bank.voice_start(v, 36 + (5 * g) % 40, 0.8, i)
                if b == 2 and g % 4 == 0:
                    bank.voice_release(v, 0.0, i)
                if b == 5 and g % 6 == 1:
                    bank.voice_start(v, 50 + g % 30, 0.7, i)
            if b == 4:
                bank.set_control(0, 0.3)
                if bank.num_controls > 1:
                    bank.set_control(1, 0.25)
            res.append(bank.process_block(n, kb.PER_VOICE | flag))
        stages = np.array([bank.voice_stage(v, i) for i in range(inst) for v in range(bank.voices)])
        mix = bank.process_block(512, flag)
        bank.close()
        outs.append((np.concatenate(res, axis=-1), stages, mix))
    _exact(np.ascontiguousarray(outs[1][0]), np.ascontiguousarray(outs[0][0]), f"graph {graph} voices")
    assert np.array_equal(outs[0][1], outs[1][1]), f"graph {graph} stages"
    _exact(np.ascontiguousarray(outs[1][2]), np.ascontiguousarray(outs[0][2]), f"graph {graph} mix")
    assert np.abs(outs[0][0]).max() > 0.3


def test_echo_time_parallel_schedule_equals_sequential_schedule():
    """Echo.k: the write-sweep / read-sweep kernels (default) against the frame-sequential kernel (KB_FX_SEQUENTIAL), bit for bit over
    ragged blocks, a delay that moves between blocks, and a block with a zero delay (which the library runs sequentially either way);
    4 instances with different delays."""
    fs, inst = 48000.0, 4
    outs = []
    for flag in (kb.FX_SEQUENTIAL, 0):
        bank = kb.FxBank(kb.FX_ECHO, inst, fs, 4096)
        for i in range(inst):
            bank.set_control(0, 0.001 + 0.004 * i, i)
            bank.set_control(1, 0.3 + 0.2 * i, i)
        res = []
        for b, n in enumerate((4096, 1001, 1, 4096, 2048, 4096)):
            if b == 3:
                bank.set_control(0, 0.0123, 1)
            if b == 4:
                bank.set_control(0, 0.0, 2)
            if b == 5:
                bank.set_control(0, 0.02, 2)
            x = np.stack([cases.fx_input(1, n, seed=20 * b + i) for i in range(inst)]).astype(np.float32)
            res.append(bank.process_inplace(x.copy(), flags=flag))
        par = bank.parallel_instances()
        bank.close()
        outs.append(np.concatenate(res, axis=-1))
    _exact(np.ascontiguousarray(outs[1]), np.ascontiguousarray(outs[0]), "echo schedules")
    assert par == inst and np.abs(outs[0]).max() > 0.4


def test_feedback_chunk_parallel_schedule_equals_sequential_schedule():
    """Feedback.k: kb_feedback_par_kernel (default: chunks shorter than the delay, one CTA per instance) against the frame-sequential kernel
    (KB_FX_SEQUENTIAL), bit for bit over ragged blocks: delays of 3.5 frames to 20 ms, a delay that moves, a zero delay (frame by frame
    inside the parallel kernel); 5 instances."""
    fs, inst = 48000.0, 5
    outs = []
    for flag in (kb.FX_SEQUENTIAL, 0):
        bank = kb.FxBank(kb.FX_FEEDBACK, inst, fs, 4096)
        for i, d in enumerate((0.00008, 0.001, 0.005, 0.0123, 0.02)):
            bank.set_control(0, d, i)
            bank.set_control(1, 0.3 + 0.15 * i, i)
        res = []
        for b, n in enumerate((4096, 1001, 1, 4096, 2048, 4096)):
            if b == 3:
                bank.set_control(0, 0.0031, 1)
            if b == 4:
                bank.set_control(0, 0.0, 2)
            if b == 5:
                bank.set_control(0, 0.01, 2)
            x = np.stack([cases.fx_input(1, n, seed=30 * b + i) for i in range(inst)]).astype(np.float32)
            res.append(bank.process_inplace(x.copy(), flags=flag))
        bank.close()
        outs.append(np.concatenate(res, axis=-1))
    _exact(np.ascontiguousarray(outs[1]), np.ascontiguousarray(outs[0]), "feedback schedules")
    assert np.abs(outs[0]).max() > 0.4


def test_midi_input_equals_note_on_off_calls():
    """kb_synth_bank_midi is Synth::input(status, b1, b2) (templates/juce/synth/Source/klang.h:3921-3929): 0x90 with a velocity is
    noteOn(b1, b2 / 127.f), 0x80 or 0x90 with velocity 0 is noteOff, other messages change nothing."""
    n, voices = 384, 8
    msgs = [[(0x90, 48 + 3 * k, 20 + 9 * k) for k in range(10)] + [(0xB0, 1, 64), (0xE0, 0, 64)],    # 10 notes into 8 voices: stealing
            [(0x80, 72, 0), (0x90, 75, 0), (0x80, 60, 100), (0x91, 50, 100), (0xC0, 5, 0)],
            [(0x90, 40, 127), (0x80, 40, 0)]]
    outs = []
    for raw in (True, False):
        kb.lib().kb_srand(3)
        bank = kb.SynthBank(kb.SY_SUPERSAW, 1, voices, 48000, n)
        res = []
        for block in msgs:
            for st, b1, b2 in block:
                if raw:
                    bank.midi(st, b1, b2)
                elif st == 0x90 and b2 > 0:
                    bank.note_on(b1, float(np.float32(b2) / np.float32(127)))
                elif st == 0x80 or (st == 0x90 and b2 == 0):
                    bank.note_off(b1, float(np.float32(b2) / np.float32(127)))
            res.append(bank.process_block(n, kb.PER_VOICE))
        bank.close()
        outs.append(np.concatenate(res, axis=-1))
    _exact(np.ascontiguousarray(outs[0]), np.ascontiguousarray(outs[1]), "midi input")
    assert np.abs(outs[0]).max() > 0.01


@pytest.mark.parametrize("graph", [cases.FX_FLANGER, cases.FX_MOD_CHORUS, cases.FX_MODDELAY])
def test_modulated_line_time_parallel_schedule_equals_sequential_schedule(graph):
    """Flanger.k / Modulation/Chorus.k / ModDelay.k (its control smoother as a serial pre-pass): the kb_modline_* kernels (default: write sweep with stash, read sweep with closed-form LFOs) against
    the frame-sequential kernel (KB_FX_SEQUENTIAL), bit for bit over ragged blocks with control changes; 4 instances."""
    fs, inst = 48000.0, 4
    outs = []
    for flag in (kb.FX_SEQUENTIAL, 0):
        bank = kb.FxBank(graph, inst, fs, 4096)
        if graph == cases.FX_FLANGER:
            for i in range(inst):
                bank.set_control(0, 0.2 + 0.25 * i, i)
                bank.set_control(1, 0.1 + 1.5 * i, i)
        res = []
        for b, n in enumerate((4096, 1001, 1, 4096, 2048, 4096)):
            if b == 3 and graph == cases.FX_FLANGER:
                bank.set_control(0, 1.0, 1)
                bank.set_control(1, 5.0, 2)
            if b in (1, 3) and graph == cases.FX_MODDELAY:
                bank.set_control(1, 1.0 if b == 1 else 0.0, 1)
                bank.set_control(0, 10.0, 2)
            x = np.stack([cases.fx_input(1, n, seed=40 * b + i) for i in range(inst)]).astype(np.float32)
            res.append(bank.process_inplace(x.copy(), flags=flag))
        bank.close()
        outs.append(np.concatenate(res, axis=-1))
    _exact(np.ascontiguousarray(outs[1]), np.ascontiguousarray(outs[0]), f"graph {graph} schedules")
    assert np.abs(outs[0]).max() > 0.4


# ------------------------------------------------------------------------------------------ debug taps
@pytest.mark.parametrize("graph,ctl,flags", [
    (kb.FX_PINGPONG, {1: 0.02, 5: 0.02, 2: 0.4, 3: 0.579}, 0), (kb.FX_PINGPONG, {1: 0.02, 5: 0.02, 2: 0.4, 3: 0.579}, kb.FX_SEQUENTIAL),
    (kb.FX_RM, {0: 440.0}, 0), (kb.FX_TREMOLO, {0: 7.0, 1: 0.3}, 0), (kb.FX_MODDELAY, {0: 3.0, 1: 0.8}, 0), (kb.FX_MODDELAY, {0: 3.0, 1: 0.8}, kb.FX_SEQUENTIAL)])
def test_debug_taps_of_a_bank_match_the_reference(graph, ctl, flags):
    """`x >> debug` (klang.h:3132-3287; PingPong.k:61, RM.k:22, Tremolo.k:27, ModDelay.k:24): 6 instances with different control settings,
    5 blocks of 1000 frames; the capture of every block and instance equals the reference's Debug::buffer bit for bit, on either schedule,
    and the audio is what it is without the capture."""
    fs, n, inst, blocks = 48000, 1000, 6, 5
    chk = oracle.ref if oracle.ref.available() else oracle.port
    chk.set_fs(fs)
    bank = kb.FxBank(graph, inst, fs, n)
    bank.debug_enable(True)
    plain = kb.FxBank(graph, inst, fs, n)
    refs = [chk.Fx(graph) for _ in range(inst)]
    for i in range(inst):
        for c, v in ctl.items():
            vi = v * (1.0 + 0.1 * i)
            bank.set_control(c, vi, i); plain.set_control(c, vi, i); refs[i].set_control(c, vi)
    ch = bank.channels
    for b in range(blocks):
        x = np.stack([cases.fx_input(ch, n, seed=900 + 7 * i + b) for i in range(inst)])
        io, io2 = x.copy(), x.copy()
        bank.process_inplace(io, flags=flags)
        plain.process_inplace(io2, flags=flags)
        got = bank.debug_read(n)
        assert got is not None and got.shape == (inst, n)
        assert bank.debug_read(n) is None                                  # Buffer::get hands a capture out once
        for i in range(inst):
            want = refs[i].process(x[i][0] if ch == 1 else x[i])
            _exact(io[i].reshape(want.shape), want, f"audio, instance {i} block {b}")
            wd = refs[i].debug()
            assert wd is not None
            _exact(got[i], wd, f"debug tap, instance {i} block {b}")
        _exact(io, io2, f"block {b}: capture on / off")
    for r in refs:
        r.close()
    bank.close(); plain.close()


def test_a_program_without_a_tap_leaves_no_capture():
    bank = kb.FxBank(kb.FX_GAIN, 2, 48000.0, 256)
    bank.debug_enable(True)
    io = np.ones((2, 1, 256), np.float32)
    bank.process_inplace(io)
    assert bank.debug_read(256) is None
    bank.debug_enable(False)
    with pytest.raises(kb.KlangB200Error):
        bank.debug_read(256)
    bank.close()


# ------------------------------------------------------------------- presets and onControl
@pytest.mark.parametrize("graph,index", [(kb.FX_PINGPONG, 0), (kb.FX_PINGPONG, 5), (kb.FX_REVERB, 0)])
def test_an_effect_preset_loads_like_in_the_reference(graph, index):
    """Plugin::presets (klang.h:1940-1981): the preset goes in through Control::set and the next blocks are the reference's, bit for bit
    (Reverb.k 'Large Hall' changes every delay time: prepare() runs its controls.changed() branch on the device mirror)."""
    fs, n = 48000, 1024
    chk = oracle.ref if oracle.ref.available() else oracle.port
    chk.set_fs(fs); chk.srand(1); kb.Engine().srand(1)
    ref = chk.Fx(graph)
    bank = kb.FxBank(graph, 2, fs, n)
    x = cases.fx_input(2, 4 * n, seed=31)
    for b in range(4):
        if b == 1:
            ref.load_preset(index)
            bank.load_preset(index, instance=1)                            # instance 0 keeps its defaults
        blk = np.stack([x[:, b * n:(b + 1) * n]] * 2).copy()
        bank.process_inplace(blk)
        want = ref.process(x[:, b * n:(b + 1) * n])
        _exact(blk[1], want, f"block {b}")
        if b >= 1:
            assert not np.array_equal(blk[0], blk[1])
    for c, v in enumerate(kb.presets(0, graph)[index][1]):
        if graph != kb.FX_PINGPONG or c not in (1,):                       # PingPong.k rewrites controls[1] every sample
            assert abs(bank.get_control(c, 1) - ref.get_control(c)) == 0, c
    ref.close(); bank.close()


def test_synth_preset_and_on_control_fan_out():
    """SuperSaw.k preset 'Synth Pad' (attack 1.0, detune, mix) loaded before the notes start; Synth::onControl / onPreset (klang.h:4399-4420)
    reach the notes that are not Off — stages as the DEVICE left them.  (SuperSaw.k:17 draws libc rand() in on(): the reference runs first,
    then the product from the same seed.)"""
    fs, n = 48000, 512
    chk = oracle.ref if oracle.ref.available() else oracle.port

    def script(sy, process, srand):
        srand(7)
        outs, counts = [], []
        counts.append(sy.load_preset(2))
        for v in range(6):
            sy.voice_start(v, 48 + 3 * v, 0.7)
        counts.append(sy.on_control(0, 0.3))
        for b in range(3):
            outs.append(process())
        sy.load_preset(0)                                                  # 'Pluck': short release
        for v in (0, 2, 4):
            sy.voice_release(v, 0.0)
        for b in range(50):                                                # (the release lasts 0.5 s + 5 ms: the three notes stop() inside these blocks)
            outs.append(process())
        counts.append(sy.on_control(1, 0.5))
        return np.concatenate(outs, axis=-1), counts

    chk.set_fs(fs)
    ref = chk.Synth(kb.SY_SUPERSAW, 16)
    want, wc = script(ref, lambda: np.atleast_2d(ref.process(n)), chk.srand)
    ref.close()
    bank = kb.SynthBank(kb.SY_SUPERSAW, 1, 16, fs, n)
    got, gc = script(bank, lambda: bank.process_block(n)[0], kb.Engine().srand)
    bank.close()
    _exact(got, want, "SuperSaw.k with presets")
    assert gc[1:] == wc[1:] and wc[1] == 6 and wc[2] < 6, (gc, wc)
    assert gc[0] == 0
