"""TEST INFRASTRUCTURE — ctypes bindings for oracle/_ref/libklang_ref.so (the reference
klang.h compiled by oracle/build_ref.py).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libklang_ref.so")

# enums (must match oracle/ref_harness.cpp)
(OSC_FAST_SAW, OSC_FAST_TRIANGLE, OSC_FAST_SQUARE, OSC_FAST_PULSE, OSC_FAST_SINE,
 OSC_BASIC_SINE, OSC_BASIC_SAW, OSC_BASIC_TRIANGLE, OSC_BASIC_SQUARE, OSC_BASIC_PULSE,
 OSC_WT_SINE, OSC_WT_SAW) = range(12)
(FLT_BIQUAD_LPF, FLT_BIQUAD_HPF, FLT_ONEPOLE_LPF, FLT_ONEPOLE_HPF,
 FLT_BIQUAD_BPF, FLT_BIQUAD_BRF, FLT_BIQUAD_APF, FLT_BUTTERWORTH_LPF1, FLT_BUTTERWORTH_LPF2) = range(9)
FX_GAIN, FX_PINGPONG, FX_REVERB, FX_DELAY_PINGPONG, FX_DELAY_REVERB = range(5)
SY_SUBTRACTIVE, SY_SUPERSAW, SY_TB303, SY_SYNTHX, SY_FILTER_K = range(5)

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

_lib = None


def available():
    return os.path.isfile(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing — run `python oracle/build_ref.py` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.ref_set_fs.argtypes = [C.c_float]
        L.ref_get_fs.restype = C.c_float
        L.ref_srand.argtypes = [C.c_uint]
        L.ref_pitch_to_frequency.argtypes = [C.c_float]
        L.ref_pitch_to_frequency.restype = C.c_float
        L.ref_osc.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, _f32p]
        L.ref_wavetable.argtypes = [C.c_int, _f32p]
        L.ref_filter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, _f32p, _f32p, C.c_void_p]
        L.ref_envelope.argtypes = [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f32p, C.c_void_p]
        L.ref_envelope_at.argtypes = [C.c_int, _f32p, C.c_int, _f32p, _f32p]
        L.ref_adsr.argtypes = [C.c_float] * 4 + [C.c_int, C.c_int, _f32p, C.c_void_p]
        L.ref_delay1000.argtypes = [C.c_int, _f32p, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.ref_stereo_delay1000.argtypes = [C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.ref_control_smooth.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, _f32p, _f32p]
        L.ref_fx_create.argtypes = [C.c_int]
        L.ref_fx_create.restype = C.c_void_p
        L.ref_fx_destroy.argtypes = [C.c_void_p]
        L.ref_fx_channels.argtypes = [C.c_void_p]
        L.ref_fx_num_controls.argtypes = [C.c_void_p]
        L.ref_fx_set_control.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_fx_get_control.argtypes = [C.c_void_p, C.c_int]
        L.ref_fx_get_control.restype = C.c_float
        L.ref_fx_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_synth_create.argtypes = [C.c_int, C.c_int]
        L.ref_synth_create.restype = C.c_void_p
        L.ref_synth_destroy.argtypes = [C.c_void_p]
        L.ref_synth_channels.argtypes = [C.c_void_p]
        L.ref_synth_num_voices.argtypes = [C.c_void_p]
        L.ref_synth_num_controls.argtypes = [C.c_void_p]
        L.ref_synth_set_control.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_synth_get_control.argtypes = [C.c_void_p, C.c_int]
        L.ref_synth_get_control.restype = C.c_float
        L.ref_synth_note_on.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_synth_note_off.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_synth_voice_start.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.ref_synth_voice_release.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_synth_voice_stage.argtypes = [C.c_void_p, C.c_int]
        L.ref_synth_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_synth_process_voices.argtypes = [C.c_void_p, _f32p, C.c_int, _i32p]
        _lib = L
    return _lib


def set_fs(fs):
    lib().ref_set_fs(float(fs))


def srand(seed):
    lib().ref_srand(int(seed))


def pitch_to_frequency(p):
    return float(lib().ref_pitch_to_frequency(float(p)))


def osc(kind, n, f, phase=None, duty=None):
    out = np.zeros(n, np.float32)
    nargs = 1 if phase is None else (2 if duty is None else 3)
    rc = lib().ref_osc(kind, nargs, float(f), float(phase or 0.0), float(duty or 0.0), n, out)
    assert rc == 0
    return out


def wavetable(kind):
    t = np.zeros(2048, np.float32)
    assert lib().ref_wavetable(kind, t) == 0
    return t


def filt(kind, x, f, Q=None, per_sample=False):
    """Run a filter over x.  f (and Q) scalars → set once; arrays with per_sample=True → set every sample."""
    x = np.ascontiguousarray(x, np.float32)
    n = len(x)
    f = np.ascontiguousarray(np.broadcast_to(np.asarray(f, np.float32), (n,) if per_sample else (1,)))
    nset = n if per_sample else 1
    qp = None
    if Q is not None:
        Q = np.ascontiguousarray(np.broadcast_to(np.asarray(Q, np.float32), (nset,)))
        qp = Q.ctypes.data
    out = np.zeros(n, np.float32)
    coeffs = np.zeros(5, np.float32)
    rc = lib().ref_filter(kind, nset, f.ctypes.data, qp, n, x, out, coeffs.ctypes.data)
    assert rc == 0
    return out, coeffs


def envelope(points, n, loop=None, release_at=-1, release_time=0.0, release_level=0.0):
    xy = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1))
    out = np.zeros(n, np.float32)
    stage = np.zeros(n, np.int32)
    ls, le = loop if loop is not None else (-1, -1)
    lib().ref_envelope(len(xy) // 2, xy, ls, le, n, release_at, release_time, release_level, out, stage.ctypes.data)
    return out, stage


def envelope_at(points, t):
    xy = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1))
    t = np.ascontiguousarray(t, np.float32)
    out = np.zeros(len(t), np.float32)
    lib().ref_envelope_at(len(xy) // 2, xy, len(t), t, out)
    return out


def adsr(A, D, S, R, n, release_at=-1):
    out = np.zeros(n, np.float32)
    stage = np.zeros(n, np.int32)
    lib().ref_adsr(A, D, S, R, n, release_at, out, stage.ctypes.data)
    return out, stage


def delay1000(x, di, df, set_at):
    x = np.ascontiguousarray(x, np.float32)
    n = len(x)
    oi, of, op = (np.zeros(n, np.float32) for _ in range(3))
    lib().ref_delay1000(n, x, np.ascontiguousarray(di, np.int32), np.ascontiguousarray(df, np.float32),
                        np.ascontiguousarray(set_at, np.float32), oi, of, op)
    return oi, of, op


def stereo_delay1000(xl, xr, df):
    xl = np.ascontiguousarray(xl, np.float32)
    xr = np.ascontiguousarray(xr, np.float32)
    n = len(xl)
    ol, orr = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib().ref_stereo_delay1000(n, xl, xr, np.ascontiguousarray(df, np.float32), ol, orr)
    return ol, orr


def control_smooth(lo, hi, initial, values):
    values = np.ascontiguousarray(values, np.float32)
    out = np.zeros(len(values), np.float32)
    lib().ref_control_smooth(lo, hi, initial, len(values), values, out)
    return out


class Fx:
    """One reference Effect instance (Effect::process(buffer), klang.h:4208-4216 / 4708-4716)."""

    def __init__(self, graph):
        self.h = lib().ref_fx_create(graph)
        if not self.h:
            raise ValueError(f"unknown effect graph {graph}")
        self.channels = lib().ref_fx_channels(self.h)
        self.num_controls = lib().ref_fx_num_controls(self.h)

    def close(self):
        if self.h:
            lib().ref_fx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_control(self, idx, v):
        lib().ref_fx_set_control(self.h, idx, float(v))

    def get_control(self, idx):
        return float(lib().ref_fx_get_control(self.h, idx))

    def process(self, x):
        """x: float32 [channels, n] (or [n] for mono). Returns processed copy."""
        y = np.array(x, np.float32, copy=True, order="C")
        if self.channels == 1:
            flat = y.reshape(-1)
            rc = lib().ref_fx_process(self.h, flat.ctypes.data, None, len(flat))
        else:
            assert y.ndim == 2 and y.shape[0] == 2
            rc = lib().ref_fx_process(self.h, y[0].ctypes.data, y[1].ctypes.data, y.shape[1])
        assert rc == 0
        return y


class Synth:
    """One reference Synth instance with up to 128 voices."""

    def __init__(self, graph, nvoices):
        self.h = lib().ref_synth_create(graph, nvoices)
        if not self.h:
            raise ValueError(f"cannot create synth graph {graph} with {nvoices} voices")
        self.channels = lib().ref_synth_channels(self.h)
        self.nvoices = lib().ref_synth_num_voices(self.h)
        self.num_controls = lib().ref_synth_num_controls(self.h)

    def close(self):
        if self.h:
            lib().ref_synth_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_control(self, idx, v):
        lib().ref_synth_set_control(self.h, idx, float(v))

    def get_control(self, idx):
        return float(lib().ref_synth_get_control(self.h, idx))

    def note_on(self, pitch, vel):
        return lib().ref_synth_note_on(self.h, int(pitch), float(vel))

    def note_off(self, pitch, vel=0.0):
        lib().ref_synth_note_off(self.h, int(pitch), float(vel))

    def voice_start(self, voice, pitch, vel):
        lib().ref_synth_voice_start(self.h, voice, float(pitch), float(vel))

    def voice_release(self, voice, vel=0.0):
        lib().ref_synth_voice_release(self.h, voice, float(vel))

    def voice_stage(self, voice):
        return lib().ref_synth_voice_stage(self.h, voice)

    def process(self, n):
        """The reference block driver (Synth::process). Returns float32 [channels, n]."""
        out = np.zeros((self.channels, n), np.float32)
        r = out[1].ctypes.data if self.channels == 2 else None
        assert lib().ref_synth_process(self.h, out[0].ctypes.data, r, n) == 0
        return out

    def process_voices(self, n):
        """Every active voice rendered alone. Returns (float32 [V, channels, n], int32 active[V])."""
        out = np.zeros((self.nvoices, self.channels, n), np.float32)
        active = np.zeros(self.nvoices, np.int32)
        assert lib().ref_synth_process_voices(self.h, out, n, active) == 0
        return out, active
