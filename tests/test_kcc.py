"""Tier B, first step: `.k` Effect programs evaluated on the device from their OWN process() body (klang_b200/kcc.py,
klang_b200/csrc/kb_kdev.cuh, include/klang_b200_user.h).

  * CPU: the reference's elementwise examples and a stateful one translate and compile for sm_100a with nvcc, export the user ABI, and refuse
    to create a bank without a CUDA device; programs outside the supported subset fail loudly at translation / compile time.
  * GPU (-m gpu): every translated program is bit-identical to the reference running the same `.k` (oracle.ref), control changes included —
    and an EDITED program (Gain.k with `in * gain * 0.5`) computes what the edited text says, which no hand-bound graph id can."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import oracle
from klang_b200 import kcc

REF = "/root/reference/examples"
HAVE_REFERENCE = os.path.isfile(os.path.join(REF, "Gain", "Gain.k"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_k_bin", "kcc")          # git-ignored build output: travels to the GPU box with the snapshot
PROGRAMS = {"gain": ("Gain/Gain.k", oracle.FX_GAIN), "pan": ("Gain/Pan.k", oracle.FX_PAN), "clipping": ("Distortion/Clipping.k", oracle.FX_CLIPPING),
            "functions": ("Distortion/Functions.k", oracle.FX_FUNCTIONS), "mute": ("Distortion/Mute.k", oracle.FX_MUTE), "iir": ("Filtering/IIR.k", oracle.FX_IIR),
            # second step: programs whose members are klang objects with state — oscillators, a biquad, a delay line (kb_kdev.cuh)
            "rm": ("Gain/RM.k", oracle.FX_RM), "tremolo": ("Gain/Tremolo.k", oracle.FX_TREMOLO), "wahwah": ("Filtering/WahWah.k", oracle.FX_WAHWAH),
            "echo": ("Delay/Echo.k", oracle.FX_ECHO), "feedback": ("Delay/Feedback.k", oracle.FX_FEEDBACK), "flanger": ("Modulation/Flanger.k", oracle.FX_FLANGER),
            "moddelay": ("Modulation/ModDelay.k", oracle.FX_MODDELAY), "mod_chorus": ("Modulation/Chorus.k", oracle.FX_MOD_CHORUS),
            "delay_reverb": ("Delay/Reverb.k", oracle.FX_DELAY_REVERB),
            # a BASELINE C4 program from its own text: Stereo::Effect, two Delay<192000>, a libm Sine LFO, two DC blockers, a control it rewrites per sample
            "pingpong": ("PingPong.k", oracle.FX_PINGPONG), "delay_pingpong": ("Delay/PingPong.k", oracle.FX_DELAY_PINGPONG)}
STATELESS = ("gain", "pan", "clipping", "functions", "mute")
# control values per (block b, instance i) for the programs of the second step: {control: value}
SCHEDULES = {
    "rm": lambda b, i: {0: 200.0 + 150.0 * i + 37.0 * b},
    "tremolo": lambda b, i: {0: 2.0 + 1.5 * i + 0.5 * b, 1: 0.1 + 0.1 * ((b + i) % 4)},
    "wahwah": lambda b, i: {0: 500.0 + 900.0 * i + 300.0 * b, 1: 0.5 + 2.0 * ((b + i) % 3), 2: 4.0 + i + 0.5 * b},
    "echo": lambda b, i: {0: 0.004 + 0.0031 * i + (0.0123 if b >= 2 else 0.0), 1: 0.3 + 0.2 * i},
    "feedback": lambda b, i: {0: 0.003 + 0.0027 * i + (0.0071 if b >= 2 else 0.0), 1: 0.5 + 0.15 * i},
    "flanger": lambda b, i: {0: 0.2 + 0.3 * i + 0.1 * b, 1: 0.5 + 1.5 * i},
    "moddelay": lambda b, i: {0: 2.0 + 3.0 * i, 1: 0.2 + 0.25 * ((b + i) % 3)},
    "mod_chorus": lambda b, i: {},
    "delay_reverb": lambda b, i: {0: 0.2 + 0.3 * i, 1: 0.02 + 0.03 * i + 0.01 * b},
    "delay_pingpong": lambda b, i: {0: 0.01 + 0.004 * i, 1: 0.03 + 0.01 * b, 2: 0.02, 3: 0.6 - 0.1 * i},
    "pingpong": lambda b, i: ({0: 0.9 - 0.1 * i, 1: 0.02 + 0.01 * i, 5: 0.02 + 0.01 * i, 4: 0.3} if b == 0 else {2: 0.4, 3: 0.579} if b == 2 else {}),
}
EDITED = "gain_edited"
# third step: mono Synth programs — the note's on() / off() on the host mirror, its process() per sample on the device (lane = voice)
SYNTHS = {"filter_k": ("Subtractive/Filter.k", "filter_k"), "breakpoint": ("Subtractive/Breakpoint.k", "breakpoint"), "ramp": ("Subtractive/Ramp.k", "ramp"),
          "release": ("Subtractive/Release.k", "release"), "release_slow_attack": ("Subtractive/Release.k", "release"), "supersaw": ("SuperSaw.k", "supersaw"),
          "supersaw_wide": ("SuperSaw.k", "supersaw"), "am": ("Modulation/AM.k", "am"), "mod_fm": ("Modulation/FM.k", "mod_fm"), "mod_fm2": ("Modulation/FM2.k", "mod_fm2"),
          # TB303.k from its own text: the ladder filter's set() evaluates exp() on the device every sample (kb_expf), OnePole::HPF, tanh
          "tb303": ("TB303.k", "tb303"), "tb303_square": ("TB303.k", "tb303"),
          # FM.k from its own text: three Operator<Sine> in series; its start-up lookup tables (Table / FUNCTION) and graph plot stay on the host
          "fm": ("FM.k", "fm"), "fm_deep": ("FM.k", "fm"),
          "additive_saw": ("Additive/Saw.k", "additive_saw"), "additive_square": ("Additive/Square.k", "additive_square"), "additive_nyquist": ("Additive/Nyquist.k", "additive_nyquist")}


def so_path(name):
    return os.path.join(BIN, f"lib{name}_k.so")


def build_all():
    """(run where /root/reference exists: here, by the CPU test below and by __graft_entry__.build()); the nvcc runs go in parallel"""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(BIN, exist_ok=True)
    jobs = [(os.path.join(REF, rel), so_path(name)) for name, (rel, _) in PROGRAMS.items()]
    jobs += [(os.path.join(REF, rel), so_path("synth_" + lib)) for rel, lib in sorted(set(SYNTHS.values()))]
    jobs += [(os.path.join(REF, spec[1]), so_path(name)) for name, spec in cases.TRANSLATED_FX_SCRIPTS.items()]
    jobs += [(os.path.join(REF, spec[1]), so_path("synth_" + name)) for name, spec in cases.TRANSLATED_SYNTH_SCRIPTS.items()]
    src = open(os.path.join(REF, "Gain", "Gain.k")).read().replace("in * gain >> out;", "in * gain * 0.5 >> out;")
    edited = os.path.join(BIN, "gain_edited.k")
    with open(edited, "w") as f:
        f.write(src)
    jobs.append((edited, so_path(EDITED)))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda j: kcc.compile_k(*j), jobs))


@pytest.mark.skipif(not HAVE_REFERENCE, reason="reference examples not present")
def test_k_programs_translate_and_compile_for_the_device(tmp_path):
    build_all()
    for name in list(PROGRAMS) + [EDITED]:
        L = C.CDLL(so_path(name))
        for sym in ("kb_user_name", "kb_user_channels", "kb_user_num_controls", "kb_user_stateless", "kb_user_last_error", "kb_user_fx_create",
                    "kb_user_fx_destroy", "kb_user_fx_set_control", "kb_user_fx_get_control", "kb_user_fx_process"):
            assert hasattr(L, sym), f"{name}: {sym} not exported"
        L.kb_user_name.restype = C.c_char_p
        assert L.kb_user_num_controls() == {"wahwah": 3, "delay_reverb": 3, "pingpong": 6, "delay_pingpong": 4}.get(name, 1 if name in STATELESS + ("iir", EDITED) else 2), name
        assert L.kb_user_channels() == (2 if name in ("pan", "pingpong", "delay_pingpong") else 1)
        assert L.kb_user_stateless() == (1 if name in STATELESS + (EDITED,) else 0)   # data members (IIR.k's `signal last`, an LFO, a delay line): lane per instance
    # the translated text is the user's: only the function definitions gained a qualifier
    src, plugin, ch = kcc.translate(open(os.path.join(REF, "Distortion", "Functions.k")).read(), "Functions.k")
    assert plugin == "Functions" and ch == 1
    assert "KB_KD float hardclip(float x){" in src and "KB_KD void process() {" in src and "hardclip(in * gain) >> out;" in src
    assert "KB_KD Functions()" not in src                                        # constructors stay host code
    # outside the subset: a synth program is refused at translation, an effect that uses primitives this header lacks fails in nvcc — loudly
    with pytest.raises(kcc.KccError):
        kcc.translate(open(os.path.join(REF, "SynTHX.k")).read(), "SynTHX.k")          # a Stereo::Synth
    src, plugin, _ = kcc.translate(open(os.path.join(REF, "Additive", "Square.k")).read(), "Square.k")
    assert plugin == "Square" and "struct Additive : OscillatorT<Additive>" in src and 'KB_USER_EXPORT_SYNTH(kb_user::Square, kb_user::Square::SquareNote, "Square")' in src
    for rel, lib in set(SYNTHS.values()):
        L = C.CDLL(so_path("synth_" + lib))
        assert L.kb_user_kind() == 1 and L.kb_user_synth_voices() == 32
    with pytest.raises(kcc.KccError):
        kcc.compile_k(os.path.join(REF, "Vocoder.k"), str(tmp_path / "libvocoder_k.so"))            # (controls.add / 27 controls / Function<> locals are not in the device header)
    # the rewrites of the later steps, on the text: a Function<> member names the function it is constructed with in its type; a namespace-scope
    # constant gets a __device__ twin; a note's on() / off() stay host functions; klang's own min / max and the float libm overloads are named
    src, _, _ = kcc.translate(open(os.path.join(REF, "Distortion", "Shaping.k")).read(), "Shaping.k")
    assert "klang::FunctionT<kb_fn_f, 2> f;" in src and "struct kb_fn_f { KB_KD static float call(float x0, float x1) { return softclip(x0, x1); } };" in src
    assert "Shaping() : f() {" in src and "return kb_tanh(c * x) / kb_tanh(c);" in src and ">> kb_graph();" in src
    src, _, _ = kcc.translate(open(os.path.join(REF, "Additive", "Resynthesis.k")).read(), "Resynthesis.k")
    assert "__device__ const float FREQ_kbd[6]" in src and "#define FREQ FREQ_kbd" in src and "#define GAIN GAIN_kbh" in src
    src, _, _ = kcc.translate(open(os.path.join(REF, "FM.k")).read(), "FM.k")
    assert "KB_KD event on(" not in src and "\t\tevent on(Pitch p, Velocity v) {" in src and "kb_graph().clear();" in src and "KB_KD void process() {" in src
    src, _, _ = kcc.translate(open(os.path.join(REF, "Subtractive", "Expression.k")).read(), "Expression.k")
    assert "kb_max(1000,f0 * 1.5)" in src
    src, _, _ = kcc.translate(open(os.path.join(REF, "TB303.k")).read(), "TB303.k")
    assert "KB_KD static float shape(float x) {" in src and "kb_exp(-0.03f*resonance)" in src           # (`static inline`: KB_KD carries the inline)
    # klang::fs and the debug sink are host objects in the reference: the translation routes them through kb_fs() / a sink value
    src, _, _ = kcc.translate(open(os.path.join(REF, "Gain", "Tremolo.k")).read(), "Tremolo.k")
    assert "mod >> klang::Debug();" in src
    src, _, _ = kcc.translate(open(os.path.join(REF, "Modulation", "Chorus.k")).read(), "Chorus.k")
    assert "* kb_fs() / 1000.f;" in src and "KB_KD signal mod(int m){" in src
    import klang_b200 as kb
    if kb.device_count() == 0:
        with pytest.raises(kcc.KccError, match="no such CUDA device"):
            kcc.UserFx(so_path("gain"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(PROGRAMS))
def test_translated_k_program_matches_the_reference_on_the_device(name):
    """The program's own process() body on the B200 against the reference running the same .k: 3 instances with different inputs, four blocks
    with a control change before each, bit-exact."""
    rel, graph = PROGRAMS[name]
    assert os.path.isfile(so_path(name)), "tests/_k_bin/kcc is built where /root/reference exists and travels with the snapshot"
    chk = oracle.ref if oracle.ref.available() else oracle.port
    fs, n, inst, blocks = 48000, 2048, 3, 4
    chk.set_fs(fs)
    fx = kcc.UserFx(so_path(name), inst, fs, n)
    assert fx.stateless == (name in STATELESS)
    refs = [chk.Fx(graph) for _ in range(inst)]
    ch = fx.channels
    lo, hi = {"clipping": (1.0, 11.0), "functions": (1.0, 25.0)}.get(name, (0.0, 1.0))
    for b in range(blocks):
        x = np.stack([cases.fx_input(ch, n, seed=900 + 7 * b + i) for i in range(inst)]) * np.float32(4.0 if name in ("clipping", "functions") else 1.0)
        for i in range(inst):
            settings = SCHEDULES[name](b, i) if name in SCHEDULES else {0: lo + (hi - lo) * ((3 * b + 5 * i + 1) % 11) / 10.0}
            for c, v in settings.items():
                fx.set_control(c, v, i)
                refs[i].set_control(c, v)
                assert fx.get_control(c, i) == refs[i].get_control(c)
        want = np.stack([np.atleast_2d(refs[i].process(x[i, 0] if ch == 1 else x[i])) for i in range(inst)])
        got = x.copy()
        fx.process_inplace(got)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{name}: block {b} differs from the reference"
        assert np.abs(want).max() > 0 or name == "mute"
    fx.close()
    for r in refs:
        r.close()


@pytest.mark.gpu
def test_edited_k_program_runs_what_the_edited_text_says():
    """Gain.k with `in * gain * 0.5 >> out;`: the hand-bound graph KB_FX_GAIN refuses this source (tests/test_k_programs.py); the translated
    program evaluates it — (x * gain) * 0.5 in fp32, exactly."""
    fs, n = 48000, 4096
    fx = kcc.UserFx(so_path(EDITED), 2, fs, n)
    x = np.stack([cases.fx_input(1, n, seed=950 + i) for i in range(2)])
    fx.set_control(0, 0.8, 0)
    fx.set_control(0, 0.3, 1)
    got = x.copy()
    fx.process_inplace(got)
    for i, g in enumerate((np.float32(0.8), np.float32(0.3))):
        want = (x[i] * g) * np.float32(0.5)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32))
    fx.close()


class _UserSynthEngine:
    """tests/cases.py drives an `engine`: this one hands out translated synths (one .so per program)."""

    def __init__(self, lib):
        self.lib, self.fs = lib, 44100.0

    def set_fs(self, fs):
        self.fs = float(fs)

    def srand(self, seed):
        import klang_b200 as kb
        kb.lib().kb_srand(seed)                            # libc srand(): on() draws the process's rand() stream on the host

    def Synth(self, graph, nvoices):
        eng = self

        class S:
            def __init__(self):
                self.u = kcc.UserSynth(so_path("synth_" + eng.lib), 1, eng.fs, 16384)
                assert self.u.voices == nvoices

            def set_control(self, c, v):
                self.u.set_control(c, v)

            def voice_start(self, v, p, vel):
                self.u.voice_start(v, p, vel)

            def voice_release(self, v, vel=0.0):
                self.u.voice_release(v, vel)

            def voice_stage(self, v):
                return self.u.voice_stage(v)

            def process_voices(self, n):
                o = self.u.process_block(n, per_voice=True)
                return o.reshape(nvoices, 1, n), None

            def process(self, n):
                return self.u.process_block(n)[0]

            def close(self):
                self.u.close()

        return S()


@pytest.mark.gpu
@pytest.mark.parametrize("fs", [44100, 48000])
@pytest.mark.parametrize("name", list(SYNTHS))
def test_translated_synth_matches_the_reference_golden(golden, name, fs):
    """The scripts of tests/cases.py (voices started, controls changed, voices released over several blocks) run on the TRANSLATED program:
    per-voice streams, note stages and the Synth::process mix are the golden vectors of the compiled reference, bit for bit."""
    rel, lib = SYNTHS[name]
    assert os.path.isfile(so_path("synth_" + lib)), "tests/_k_bin/kcc is built where /root/reference exists and travels with the snapshot"
    eng = _UserSynthEngine(lib)
    r = cases.run_synth_script(eng, name, fs, per_voice=True)
    g = golden[fs]
    assert np.array_equal(r["out"].view(np.uint32), g[f"synth/{name}/voices"].view(np.uint32)), f"{name}: per-voice streams differ from the reference"
    assert np.array_equal(r["stages"], g[f"synth/{name}/stages"])
    r = cases.run_synth_script(eng, name, fs, per_voice=False)
    assert np.array_equal(np.atleast_2d(r["out"]).view(np.uint32), np.atleast_2d(g[f"synth/{name}/mix"]).view(np.uint32)), f"{name}: Synth::process output differs"


class _UserFxEngine:
    """tests/cases.py drives an `engine`: this one hands out the translated program behind a reference-only effect id."""

    def __init__(self):
        self.fs = 44100.0

    def set_fs(self, fs):
        self.fs = float(fs)

    def srand(self, seed):
        import klang_b200 as kb
        kb.lib().kb_srand(seed)

    def Fx(self, graph):
        name = next(k for k, v in cases.TRANSLATED_FX_SCRIPTS.items() if v[0] == graph)
        u = kcc.UserFx(so_path(name), 1, self.fs, 16384)

        class F:
            channels, num_controls = u.channels, u.num_controls

            def set_control(self, c, v):
                u.set_control(c, v)

            def process(self, x):
                y = np.array(x, np.float32, copy=True, order="C")
                u.process_inplace(y.reshape(1, u.channels, -1))
                return y

            def close(self):
                u.close()

        return F()


@pytest.mark.gpu
@pytest.mark.parametrize("fs", [44100, 48000])
def test_programs_without_a_bound_graph_match_the_reference(fs):
    """Filtering/Objects.k (Noise >> LPF: the device continues the process's libc rand() stream), Filtering/Bands.k (two BPF, grouped controls),
    Filtering/EQ.k (LPF / HPF set in prepare()): no KB_FX_* id exists for them — the product runs their own text — and every block is the compiled
    reference's (tests/golden/klang_ref_translated_fs*.npz, written by tests/gen_golden.py; live against oracle.ref where it is present)."""
    g = {k: v for k, v in np.load(os.path.join(ROOT, "tests", "golden", f"klang_ref_translated_fs{fs}.npz")).items() if k.startswith("fx/")}
    got = cases.translated_cases(_UserFxEngine(), fs)
    assert set(got) == set(g) and len(g) == len(cases.TRANSLATED_FX_SCRIPTS) == 6
    for k in g:
        assert np.array_equal(got[k].view(np.uint32), g[k].view(np.uint32)), f"{k}: differs from the reference"
        assert np.abs(g[k]).max() > 0
    if oracle.ref.available():
        live = cases.translated_cases(oracle.ref, fs)
        for k in g:
            assert np.array_equal(live[k].view(np.uint32), g[k].view(np.uint32)), k


@pytest.mark.gpu
@pytest.mark.parametrize("fs", [44100, 48000])
@pytest.mark.parametrize("name", list(cases.TRANSLATED_SYNTH_SCRIPTS))
def test_synth_programs_without_a_bound_graph_match_the_reference(name, fs):
    """Subtractive/Expression.k (three Saws under an enveloped vibrato LFO into an LPF; on() draws random() on the host mirror): no KB_SY_* id
    exists — the product runs the program's own text — and per-voice streams, note stages and the Synth::process mix are the compiled
    reference's (reference-only synth id, tests/golden/klang_ref_translated_fs*.npz; live against oracle.ref where it is present)."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"klang_ref_translated_fs{fs}.npz")))
    got = cases.translated_synth_cases(_UserSynthEngine(name), name, fs)
    for k, v in got.items():
        a, b = np.atleast_2d(v), np.atleast_2d(g[k])
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{k}: differs from the reference"
    assert np.abs(g[f"synth/{name}/mix"]).max() > 0
    if oracle.ref.available():
        live = cases.translated_synth_cases(oracle.ref, name, fs)
        for k in got:
            assert np.array_equal(np.atleast_2d(live[k]).view(np.uint32), np.atleast_2d(g[k]).view(np.uint32)), k


@pytest.mark.gpu
def test_noise_program_hands_the_rand_stream_back_to_the_host():
    """Objects.k draws one rand() per sample on the device; afterwards libc continues where the reference's would: a host draw after two blocks
    equals the draw after srand(5) + 2 x 1000 rand() calls."""
    import ctypes
    import klang_b200 as kb
    libc = ctypes.CDLL(None)
    kb.lib().kb_srand(5)
    for _ in range(2000):
        libc.rand()
    want = libc.rand()
    kb.lib().kb_srand(5)
    u = kcc.UserFx(so_path("k_objects"), 1, 48000.0, 1000)
    io = np.zeros((1, 1, 1000), np.float32)
    u.process_inplace(io)
    u.process_inplace(io)
    assert libc.rand() == want
    u.close()
