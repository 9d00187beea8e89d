"""TEST INFRASTRUCTURE — ctypes bindings shared by the two oracles:

  * oracle.ref  = oracle/_ref/libklang_ref.so, the reference klang.h itself compiled by
                  oracle/build_ref.py (symbols ref_*),
  * oracle.port = oracle/libklang_port.so, the plain-C restatement oracle/klang_port.c
                  (symbols kp_*, same signatures).

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py may import this package."""
import ctypes as C
import os

import numpy as np

# enums (must match oracle/ref_harness.cpp and oracle/klang_port*.c)
(OSC_FAST_SAW, OSC_FAST_TRIANGLE, OSC_FAST_SQUARE, OSC_FAST_PULSE, OSC_FAST_SINE,
 OSC_BASIC_SINE, OSC_BASIC_SAW, OSC_BASIC_TRIANGLE, OSC_BASIC_SQUARE, OSC_BASIC_PULSE,
 OSC_WT_SINE, OSC_WT_SAW, OSC_BASIC_NOISE, OSC_FAST_NOISE) = range(14)
(FLT_BIQUAD_LPF, FLT_BIQUAD_HPF, FLT_ONEPOLE_LPF, FLT_ONEPOLE_HPF,
 FLT_BIQUAD_BPF, FLT_BIQUAD_BRF, FLT_BIQUAD_APF, FLT_BUTTERWORTH_LPF1, FLT_BUTTERWORTH_LPF2,
 FLT_DCF, FLT_IIR1, FLT_IIR2, FLT_MODAL, FLT_FOLLOWER_PEAK, FLT_FOLLOWER_RMS, FLT_WINDOW_MEAN, FLT_WINDOW_RMS) = range(17)
FX_GAIN, FX_PINGPONG, FX_REVERB, FX_DELAY_PINGPONG, FX_DELAY_REVERB, FX_PAN, FX_RM, FX_TREMOLO, FX_CLIPPING, FX_ECHO, FX_FEEDBACK, FX_FUNCTIONS, FX_MUTE, FX_IIR, FX_WAHWAH, FX_FLANGER, FX_MODDELAY, FX_MOD_CHORUS = range(18)
SY_SUBTRACTIVE, SY_SUPERSAW, SY_TB303, SY_SYNTHX, SY_FILTER_K, SY_FM, SY_BREAKPOINT, SY_RAMP, SY_RELEASE, SY_ADDITIVE_SAW, SY_ADDITIVE_SQUARE, SY_AM, SY_MOD_FM, SY_MOD_FM2, SY_ADDITIVE_NYQUIST = range(15)

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Oracle:
    def __init__(self, path, prefix, hint):
        self.path, self.prefix, self.hint = path, prefix, hint
        self._lib = None

    def available(self):
        return os.path.isfile(self.path)

    def fn(self, name):
        return getattr(self.lib(), self.prefix + name)

    def lib(self):
        if self._lib is not None:
            return self._lib
        if not self.available():
            raise RuntimeError(f"{self.path} missing — {self.hint}")
        L = C.CDLL(self.path)
        p = self.prefix

        def sig(name, argtypes, restype=C.c_int):
            f = getattr(L, p + name)
            f.argtypes, f.restype = argtypes, restype

        vp, i, f32 = C.c_void_p, C.c_int, C.c_float
        sig("set_fs", [f32], None)
        sig("get_fs", [], f32)
        sig("srand", [C.c_uint], None)
        sig("pitch_to_frequency", [f32], f32)
        sig("osc", [i, i, f32, f32, f32, i, _f32p])
        sig("wavetable", [i, _f32p])
        sig("filter", [i, i, vp, vp, i, _f32p, _f32p, vp])
        sig("envelope", [i, _f32p, i, i, i, i, f32, f32, _f32p, vp])
        sig("envelope_at", [i, _f32p, i, _f32p, _f32p])
        sig("adsr", [f32, f32, f32, f32, i, i, _f32p, vp])
        sig("delay1000", [i, _f32p, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p])
        sig("stereo_delay1000", [i, _f32p, _f32p, _f32p, _f32p, _f32p])
        sig("delay1000_lagrange", [i, _f32p, _f32p, _f32p])
        sig("control_smooth", [f32, f32, f32, i, _f32p, _f32p])
        sig("sample", [_f32p, i, i, f32, f32, i, _f32p])
        sig("fx_create", [i], vp)
        sig("fx_destroy", [vp], None)
        sig("fx_channels", [vp])
        sig("fx_num_controls", [vp])
        sig("fx_set_control", [vp, i, f32], None)
        sig("fx_get_control", [vp, i], f32)
        sig("fx_process", [vp, vp, vp, i])
        sig("fx_debug", [vp, vp, i])
        sig("fx_num_presets", [vp])
        sig("fx_preset", [vp, i, C.c_char_p, i, vp, i])
        sig("fx_load_preset", [vp, i])
        sig("synth_num_presets", [vp])
        sig("synth_preset", [vp, i, C.c_char_p, i, vp, i])
        sig("synth_load_preset", [vp, i])
        sig("synth_on_control", [vp, i, f32])
        sig("synth_create", [i, i], vp)
        sig("synth_destroy", [vp], None)
        sig("synth_channels", [vp])
        sig("synth_num_voices", [vp])
        sig("synth_num_controls", [vp])
        sig("synth_set_control", [vp, i, f32], None)
        sig("synth_get_control", [vp, i], f32)
        sig("synth_note_on", [vp, i, f32])
        sig("synth_note_off", [vp, i, f32], None)
        sig("synth_voice_start", [vp, i, f32, f32], None)
        sig("synth_voice_release", [vp, i, f32], None)
        sig("synth_voice_stage", [vp, i])
        sig("synth_process", [vp, vp, vp, i])
        sig("synth_process_voices", [vp, _f32p, i, _i32p])
        self._lib = L
        return L

    # ------------------------------------------------------------------ globals
    def set_fs(self, fs):
        self.fn("set_fs")(float(fs))

    def srand(self, seed):
        self.fn("srand")(int(seed))

    def pitch_to_frequency(self, p):
        return float(self.fn("pitch_to_frequency")(float(p)))

    # --------------------------------------------------------------- primitives
    def osc(self, kind, n, f, phase=None, duty=None):
        out = np.zeros(n, np.float32)
        nargs = 1 if phase is None else (2 if duty is None else 3)
        rc = self.fn("osc")(kind, nargs, float(f), float(phase or 0.0), float(duty or 0.0), n, out)
        assert rc == 0
        return out

    def sample(self, table, n, f, phase=None):
        """klang::Sample over `table` (klang.h:3679-3720): set(f) or set(f, phase), then n ticks."""
        table = np.ascontiguousarray(table, np.float32)
        out = np.zeros(n, np.float32)
        assert self.fn("sample")(table, len(table), 1 if phase is None else 2, float(f), float(phase or 0.0), n, out) == 0
        return out

    def wav_decode(self, image):
        """File::WAV (klang.h:5951-6085): (float32 samples, (channels, samplerate, bits)).  The compiled reference loads a file
        (WAV::load(path); its load(Memory&) overload assigns a pointer to a Memory and cannot work), the port reads the image."""
        info = (C.c_int * 3)()
        f = getattr(self.lib(), self.prefix + "wav_decode")
        f.restype = C.c_int
        if self.prefix == "ref_":
            import tempfile
            with tempfile.NamedTemporaryFile(suffix=".wav") as t:
                t.write(image)
                t.flush()
                f.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p]
                n = f(t.name.encode(), None, 0, info)
                if n < 0:
                    raise ValueError(f"File::WAV refused the image ({n})")
                out = np.zeros(n, np.float32)
                f(t.name.encode(), out.ctypes.data, n, info)
        else:
            f.argtypes = [C.c_char_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]
            n = f(image, len(image), None, 0, info)
            if n < 0:
                raise ValueError(f"File::WAV refused the image ({n})")
            out = np.zeros(n, np.float32)
            f(image, len(image), out.ctypes.data, n, info)
        return out, tuple(info)

    def wavetable(self, kind):
        t = np.zeros(2048, np.float32)
        assert self.fn("wavetable")(kind, t) == 0
        return t

    def filt(self, kind, x, f, Q=None, per_sample=False):
        """f (and Q) scalars → set once; per_sample=True → set(f[s],Q[s]) before every sample."""
        x = np.ascontiguousarray(x, np.float32)
        n = len(x)
        nset = n if per_sample else 1
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(f, np.float32), (nset,)))
        qp = None
        if Q is not None:
            Q = np.ascontiguousarray(np.broadcast_to(np.asarray(Q, np.float32), (nset,)))
            qp = Q.ctypes.data
        out = np.zeros(n, np.float32)
        coeffs = np.zeros(5, np.float32)
        rc = self.fn("filter")(kind, nset, f.ctypes.data, qp, n, x, out, coeffs.ctypes.data)
        assert rc == 0
        return out, coeffs

    def envelope(self, points, n, loop=None, release_at=-1, release_time=0.0, release_level=0.0):
        xy = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1))
        out = np.zeros(n, np.float32)
        stage = np.zeros(n, np.int32)
        ls, le = loop if loop is not None else (-1, -1)
        self.fn("envelope")(len(xy) // 2, xy, ls, le, n, release_at, release_time, release_level, out, stage.ctypes.data)
        return out, stage

    def envelope_at(self, points, t):
        xy = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1))
        t = np.ascontiguousarray(t, np.float32)
        out = np.zeros(len(t), np.float32)
        self.fn("envelope_at")(len(xy) // 2, xy, len(t), t, out)
        return out

    def adsr(self, A, D, S, R, n, release_at=-1):
        out = np.zeros(n, np.float32)
        stage = np.zeros(n, np.int32)
        self.fn("adsr")(A, D, S, R, n, release_at, out, stage.ctypes.data)
        return out, stage

    def delay1000(self, x, di, df, set_at):
        x = np.ascontiguousarray(x, np.float32)
        n = len(x)
        oi, of, op = (np.zeros(n, np.float32) for _ in range(3))
        self.fn("delay1000")(n, x, np.ascontiguousarray(di, np.int32), np.ascontiguousarray(df, np.float32),
                             np.ascontiguousarray(set_at, np.float32), oi, of, op)
        return oi, of, op

    def delay_lagrange(self, x, df):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(len(x), np.float32)
        self.fn("delay1000_lagrange")(len(x), x, np.ascontiguousarray(df, np.float32), out)
        return out

    def stereo_delay1000(self, xl, xr, df):
        xl = np.ascontiguousarray(xl, np.float32)
        xr = np.ascontiguousarray(xr, np.float32)
        n = len(xl)
        ol, orr = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.fn("stereo_delay1000")(n, xl, xr, np.ascontiguousarray(df, np.float32), ol, orr)
        return ol, orr

    def control_smooth(self, lo, hi, initial, values):
        values = np.ascontiguousarray(values, np.float32)
        out = np.zeros(len(values), np.float32)
        self.fn("control_smooth")(lo, hi, initial, len(values), values, out)
        return out

    def Fx(self, graph):
        return Fx(self, graph)

    def Synth(self, graph, nvoices):
        return Synth(self, graph, nvoices)


def _presets(o, kind, h):
    out = []
    for p in range(o.fn(kind + "_num_presets")(h)):
        name, vals = C.create_string_buffer(40), (C.c_float * 16)()
        k = o.fn(kind + "_preset")(h, p, name, 40, C.addressof(vals), 16)
        out.append((name.value.decode(), [float(vals[i]) for i in range(k)]))
    return out


class Fx:
    """One oracle Effect instance (Effect::process(buffer), klang.h:4208-4216 / 4708-4716)."""

    def __init__(self, oracle, graph):
        self.o = oracle
        self.h = oracle.fn("fx_create")(graph)
        if not self.h:
            raise ValueError(f"unknown effect graph {graph}")
        self.channels = oracle.fn("fx_channels")(self.h)
        self.num_controls = oracle.fn("fx_num_controls")(self.h)

    def close(self):
        if self.h:
            self.o.fn("fx_destroy")(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_control(self, idx, v):
        self.o.fn("fx_set_control")(self.h, idx, float(v))

    def get_control(self, idx):
        return float(self.o.fn("fx_get_control")(self.h, idx))

    def process(self, x):
        """x: float32 [channels, n] (or [n] for mono effects). Returns a processed copy."""
        y = np.array(x, np.float32, copy=True, order="C")
        if self.channels == 1:
            flat = y.reshape(-1)
            rc = self.o.fn("fx_process")(self.h, flat.ctypes.data, None, len(flat))
        else:
            assert y.ndim == 2 and y.shape[0] == 2
            rc = self.o.fn("fx_process")(self.h, y[0].ctypes.data, y[1].ctypes.data, y.shape[1])
        assert rc == 0
        self._last_n = y.shape[-1]
        return y

    def presets(self):
        """Plugin::presets (klang.h:1940-1981): [(name, [values])]."""
        return _presets(self.o, "fx", self.h)

    def load_preset(self, index):
        assert self.o.fn("fx_load_preset")(self.h, index) == 0

    def debug(self):
        """The `>> debug` capture of the last block (klang.h:3132-3287): float32 [n], or None if the block wrote none."""
        d = np.zeros(self._last_n, np.float32)
        return d if self.o.fn("fx_debug")(self.h, d.ctypes.data, self._last_n) else None


class Synth:
    """One oracle Synth instance with up to 128 voices."""

    def __init__(self, oracle, graph, nvoices):
        self.o = oracle
        self.h = oracle.fn("synth_create")(graph, nvoices)
        if not self.h:
            raise ValueError(f"cannot create synth graph {graph} with {nvoices} voices")
        self.channels = oracle.fn("synth_channels")(self.h)
        self.nvoices = oracle.fn("synth_num_voices")(self.h)
        self.num_controls = oracle.fn("synth_num_controls")(self.h)

    def close(self):
        if self.h:
            self.o.fn("synth_destroy")(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def presets(self):
        return _presets(self.o, "synth", self.h)

    def load_preset(self, index):
        assert self.o.fn("synth_load_preset")(self.h, index) == 0

    def on_control(self, idx, value):
        """Synth::onControl (klang.h:4399-4404); returns how many notes (stage != Off) were notified."""
        return int(self.o.fn("synth_on_control")(self.h, idx, float(value)))

    def set_control(self, idx, v):
        self.o.fn("synth_set_control")(self.h, idx, float(v))

    def get_control(self, idx):
        return float(self.o.fn("synth_get_control")(self.h, idx))

    def note_on(self, pitch, vel):
        return self.o.fn("synth_note_on")(self.h, int(pitch), float(vel))

    def note_off(self, pitch, vel=0.0):
        self.o.fn("synth_note_off")(self.h, int(pitch), float(vel))

    def voice_start(self, voice, pitch, vel):
        self.o.fn("synth_voice_start")(self.h, voice, float(pitch), float(vel))

    def voice_release(self, voice, vel=0.0):
        self.o.fn("synth_voice_release")(self.h, voice, float(vel))

    def voice_stage(self, voice):
        return self.o.fn("synth_voice_stage")(self.h, voice)

    def process(self, n):
        """The block driver (Synth::process). Returns float32 [channels, n]."""
        out = np.zeros((self.channels, n), np.float32)
        r = out[1].ctypes.data if self.channels == 2 else None
        assert self.o.fn("synth_process")(self.h, out[0].ctypes.data, r, n) == 0
        return out

    def process_voices(self, n):
        """Every active voice rendered alone. Returns (float32 [V, channels, n], int32 active[V])."""
        out = np.zeros((self.nvoices, self.channels, n), np.float32)
        active = np.zeros(self.nvoices, np.int32)
        assert self.o.fn("synth_process_voices")(self.h, out, n, active) == 0
        return out, active


_HERE = os.path.dirname(os.path.abspath(__file__))
ref = Oracle(os.path.join(_HERE, "_ref", "libklang_ref.so"), "ref_",
             "run `python oracle/build_ref.py` where /root/reference exists")
port = Oracle(os.path.join(_HERE, "libklang_port.so"), "kp_", "run `make -C oracle libklang_port.so`")
