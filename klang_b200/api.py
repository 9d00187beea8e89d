"""ctypes binding of include/klang_b200.h plus thin bank objects.

`Engine` exposes the same surface as the parity oracles (oracle/bindings.py: set_fs / srand / osc / filt /
envelope / adsr / Fx / Synth) so the parity tests drive the CUDA path and the CPU oracles with the same
script — but it shares no code with them and never falls back to them: if the library or a CUDA device is
missing, calls raise KlangB200Error."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "lib", "libklang_b200.so")

FX_GAIN, FX_PINGPONG, FX_REVERB, FX_DELAY_PINGPONG, FX_DELAY_REVERB, FX_PAN, FX_RM, FX_TREMOLO, FX_CLIPPING, FX_ECHO, FX_FEEDBACK, FX_FUNCTIONS, FX_MUTE, FX_IIR, FX_WAHWAH, FX_FLANGER, FX_MODDELAY, FX_MOD_CHORUS = range(18)
SY_SUBTRACTIVE, SY_SUPERSAW, SY_TB303, SY_SYNTHX, SY_FILTER_K, SY_FM, SY_BREAKPOINT, SY_RAMP, SY_RELEASE, SY_ADDITIVE_SAW, SY_ADDITIVE_SQUARE, SY_AM, SY_MOD_FM, SY_MOD_FM2, SY_ADDITIVE_NYQUIST = range(15)
DEVICE_PTR, PER_VOICE, MIX_SUM, BANK_MIX, LANE_PER_VOICE, FX_SEQUENTIAL, ASYNC_HOST, FX_TOLERANCE = 1, 2, 4, 8, 16, 32, 64, 128

# every symbol include/klang_b200.h declares: (name, restype, argtypes)
_vp, _i, _f, _u, _ll, _d = C.c_void_p, C.c_int, C.c_float, C.c_uint, C.c_longlong, C.c_double
_fp, _ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
SYMBOLS = [
    ("kb_version", _i, []), ("kb_device_count", _i, []), ("kb_last_error", C.c_char_p, []),
    ("kb_srand", None, [_u]), ("kb_pitch_to_frequency", _f, [_f]),
    ("kb_graph_source_hash", C.c_ulonglong, [_i, _i]), ("kb_graph_source_path", C.c_char_p, [_i, _i]),
    ("kb_fx_bank_create", _vp, [_i, _i, _f, _i, _i]), ("kb_fx_bank_destroy", None, [_vp]),
    ("kb_fx_bank_channels", _i, [_vp]), ("kb_fx_bank_instances", _i, [_vp]), ("kb_fx_bank_num_controls", _i, [_vp]),
    ("kb_fx_bank_set_control", _i, [_vp, _i, _i, _f]), ("kb_fx_bank_get_control", _i, [_vp, _i, _i, _fp]),
    ("kb_fx_bank_process", _i, [_vp, _vp, _i, _u]), ("kb_fx_bank_sync", _i, [_vp]), ("kb_fx_bank_set_stream", _i, [_vp, _vp]),
    ("kb_fx_bank_bytes_per_frame", _d, [_vp]), ("kb_fx_bank_launches", _ll, [_vp]), ("kb_fx_bank_state_bytes", _ll, [_vp]), ("kb_fx_bank_parallel_instances", _i, [_vp]), ("kb_fx_bank_tolerance_instances", _i, [_vp]),
    ("kb_fx_bank_profile", _i, [_vp, _i]), ("kb_fx_bank_profile_read", _i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    ("kb_graph_num_presets", _i, [_i, _i]), ("kb_graph_preset", _i, [_i, _i, _i, C.c_char_p, _i, _vp, _i]),
    ("kb_fx_bank_load_preset", _i, [_vp, _i, _i]), ("kb_synth_bank_load_preset", _i, [_vp, _i, _i]),
    ("kb_synth_bank_on_control", _i, [_vp, _i, _i, _f]), ("kb_synth_bank_on_preset", _i, [_vp, _i, _i]),
    ("kb_fx_bank_debug_enable", _i, [_vp, _i]), ("kb_fx_bank_debug_read", _i, [_vp, _vp, _i, _u]),
    ("kb_synth_bank_create", _vp, [_i, _i, _i, _f, _i, _i]), ("kb_synth_bank_destroy", None, [_vp]),
    ("kb_synth_bank_channels", _i, [_vp]), ("kb_synth_bank_instances", _i, [_vp]), ("kb_synth_bank_voices", _i, [_vp]),
    ("kb_synth_bank_num_controls", _i, [_vp]),
    ("kb_synth_bank_set_control", _i, [_vp, _i, _i, _f]), ("kb_synth_bank_get_control", _i, [_vp, _i, _i, _fp]),
    ("kb_synth_bank_note_on", _i, [_vp, _i, _i, _f]), ("kb_synth_bank_note_off", _i, [_vp, _i, _i, _f]),
    ("kb_synth_bank_midi", _i, [_vp, _i, _i, _i, _i]),
    ("kb_synth_bank_voice_start", _i, [_vp, _i, _i, _f, _f]), ("kb_synth_bank_voice_release", _i, [_vp, _i, _i, _f]),
    ("kb_synth_bank_voice_stage", _i, [_vp, _i, _i]), ("kb_synth_bank_events", _i, [_vp, _i, _vp]),
    ("kb_synth_bank_process", _i, [_vp, _vp, _i, _u]), ("kb_synth_bank_sync", _i, [_vp]), ("kb_synth_bank_set_stream", _i, [_vp, _vp]),
    ("kb_synth_bank_launches", _ll, [_vp]), ("kb_synth_bank_state_bytes", _ll, [_vp]),
    ("kb_synth_bank_transfer_bytes", _i, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    ("kb_synth_bank_profile", _i, [_vp, _i]), ("kb_synth_bank_profile_read", _i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    ("kb_mixdown_create", _vp, [_i, _i, _i, _i]), ("kb_mixdown_destroy", None, [_vp]),
    ("kb_mixdown_export", _i, [_vp, _vp]), ("kb_mixdown_import", _i, [_vp, _vp]),
    ("kb_mixdown_acquire", _vp, [_vp, _vp]), ("kb_mixdown_publish", _i, [_vp, _vp]), ("kb_mixdown_put", _i, [_vp, _vp, _i, _vp]),
    ("kb_mixdown_collect", _i, [_vp, _vp, _i, _vp]),
    ("kb_mixdown_step", _i, [_vp, _vp, _i, _vp, _vp]), ("kb_mixdown_stream_wait", _i, [_vp, _vp]), ("kb_synth_bank_process_mixdown", _i, [_vp, _vp, _vp, _i, _u]),
    ("kb_synth_bank_step", _i, [_vp, _i, _vp, _vp, _i, _u]),
    ("kb_synth_bank_step_mixdown", _i, [_vp, _i, _vp, _vp, _vp, _i, _u]), ("kb_mixdown_host_wait", _i, [_vp, _i]),
    ("kb_prim_osc", _i, [_i, _i, _f, _f, _f, _f, _i, _vp]),
    ("kb_prim_filter", _i, [_i, _i, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    ("kb_prim_envelope", _i, [_i, _vp, _i, _i, _f, _i, _i, _f, _f, _vp, _vp]),
    ("kb_prim_adsr", _i, [_f, _f, _f, _f, _f, _i, _i, _vp, _vp]),
    ("kb_prim_math", _i, [_i, _i, _vp, _vp]),
    ("kb_prim_delay", _i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("kb_wav_decode", _i, [_vp, _ll, _vp, _i, _vp]), ("kb_prim_sample", _i, [_vp, _i, _i, _f, _f, _i, _vp]),
    ("kb_prim_wavetable", _i, [_i, _f, _vp]), ("kb_prim_stereo_delay", _i, [_i, _vp, _vp, _vp, _vp, _vp]),
    ("kb_prim_control_smooth", _i, [_f, _f, _f, _i, _vp, _vp]), ("kb_prim_envelope_at", _i, [_i, _vp, _i, _vp, _vp]),
]


EV_NOTE_ON, EV_NOTE_OFF, EV_VOICE_START, EV_VOICE_RELEASE, EV_CONTROL = range(5)
# numpy mirror of kb_note_event
EVENT_DTYPE = np.dtype([("type", np.int32), ("instance", np.int32), ("key", np.int32), ("pitch", np.float32), ("velocity", np.float32)])


class KlangB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded C-ABI library (loads it on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(lib_path):
            raise KlangB200Error(f"{lib_path} is missing — run `python -m klang_b200.build` (needs nvcc); there is no CPU path")
        L = C.CDLL(lib_path)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def device_count():
    return lib().kb_device_count()


def _check(rc, what):
    if rc < 0:
        raise KlangB200Error(f"{what}: error {rc}: {lib().kb_last_error().decode()}")
    return rc


def _ptr(x, numel=None, what="buffer", device=None):
    """Device pointer of a torch CUDA tensor, or host pointer of a numpy array.  The C ABI takes raw pointers and reads / writes
    `numel` float32 elements behind them, so dtype, contiguity, size (and the device of a CUDA tensor) are checked here."""
    if isinstance(x, np.ndarray):
        if x.dtype != np.float32 or not x.flags["C_CONTIGUOUS"]:
            raise KlangB200Error(f"{what}: need a C-contiguous float32 array, got dtype {x.dtype}, contiguous={x.flags['C_CONTIGUOUS']}")
        if not x.flags["WRITEABLE"]:
            raise KlangB200Error(f"{what}: array is read-only (the call writes into it)")
        if numel is not None and x.size < numel:
            raise KlangB200Error(f"{what}: {x.size} elements, the call needs {numel}")
        return x.ctypes.data, False
    if not hasattr(x, "data_ptr"):
        raise KlangB200Error(f"{what}: expected a numpy array or a torch tensor, got {type(x).__name__}")
    if str(x.dtype) != "torch.float32" or not x.is_contiguous():
        raise KlangB200Error(f"{what}: need a contiguous float32 tensor, got {x.dtype}, contiguous={x.is_contiguous()}")
    if numel is not None and x.numel() < numel:
        raise KlangB200Error(f"{what}: {x.numel()} elements, the call needs {numel}")
    if x.is_cuda and device is not None and x.device.index != device:
        raise KlangB200Error(f"{what}: tensor lives on cuda:{x.device.index}, the bank on cuda:{device}")
    return x.data_ptr(), bool(x.is_cuda)


def presets(is_synth, graph):
    """Factory presets of the program behind a graph id: [(name, [values])] (Plugin::presets, klang.h:1940-1981)."""
    out = []
    for p in range(lib().kb_graph_num_presets(1 if is_synth else 0, graph)):
        name, vals = C.create_string_buffer(40), (C.c_float * 16)()
        k = lib().kb_graph_preset(1 if is_synth else 0, graph, p, name, 40, vals, 16)
        out.append((name.value.decode(), [float(vals[i]) for i in range(k)]))
    return out


def wav_decode(image):
    """File::WAV::load + operator>> (klang.h:5997-6085) over a file image: (float32 samples, (channels, samplerate, bits)).  Host code."""
    buf = (C.c_ubyte * len(image)).from_buffer_copy(image)
    info = (C.c_int * 3)()
    n = lib().kb_wav_decode(C.addressof(buf), len(image), None, 0, info)
    if n < 0:
        _check(n, "kb_wav_decode")
    out = np.zeros(n, np.float32)
    _check(min(0, lib().kb_wav_decode(C.addressof(buf), len(image), out.ctypes.data, n, info)), "kb_wav_decode")
    return out, tuple(info)


class FxBank:
    """`instances` objects of one klang Effect evaluated together (Effect::process(buffer), klang.h:4208-4216)."""

    def __init__(self, graph, instances=1, fs=44100.0, max_block=16384, device=0):
        self.h = lib().kb_fx_bank_create(graph, instances, float(fs), max_block, device)
        if not self.h:
            raise KlangB200Error("kb_fx_bank_create: " + lib().kb_last_error().decode())
        self.graph, self.instances, self.max_block, self.device = graph, instances, max_block, device
        self.channels = lib().kb_fx_bank_channels(self.h)
        self.num_controls = lib().kb_fx_bank_num_controls(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().kb_fx_bank_destroy(self.h)
            self.h = None

    __del__ = close

    def set_control(self, idx, value, instance=None):
        for i in (range(self.instances) if instance is None else [instance]):
            _check(lib().kb_fx_bank_set_control(self.h, i, idx, float(value)), "kb_fx_bank_set_control")

    def get_control(self, idx, instance=0):
        v = C.c_float()
        _check(lib().kb_fx_bank_get_control(self.h, instance, idx, C.byref(v)), "kb_fx_bank_get_control")
        return float(v.value)

    def load_preset(self, index, instance=None):
        for i in (range(self.instances) if instance is None else [instance]):
            _check(lib().kb_fx_bank_load_preset(self.h, i, index), "kb_fx_bank_load_preset")

    def debug_enable(self, on=True):
        """`>> debug` capture (klang.h:3132-3287): while on, every process() also records the block's debug tap of every instance."""
        _check(lib().kb_fx_bank_debug_enable(self.h, 1 if on else 0), "kb_fx_bank_debug_enable")

    def debug_read(self, n):
        """The last block's capture, float32 [instances, n], or None if the program taps nothing (Debug::Buffer::get)."""
        out = np.zeros((self.instances, n), np.float32)
        rc = lib().kb_fx_bank_debug_read(self.h, out.ctypes.data, n, 0)
        if rc < 0:
            _check(rc, "kb_fx_bank_debug_read")
        return out if rc == 1 else None

    def process_inplace(self, io, n=None, flags=0):
        """io: float32 [instances, channels, n], numpy (host) or torch CUDA tensor (asynchronous)."""
        n = io.shape[-1] if n is None else n
        if n < 0:
            raise KlangB200Error("kb_fx_bank_process: negative block length")
        p, dev = _ptr(io, self.instances * self.channels * n, "kb_fx_bank_process io [instances, channels, n]", self.device)
        _check(lib().kb_fx_bank_process(self.h, p, n, (flags | DEVICE_PTR) if dev else (flags & ~DEVICE_PTR)), "kb_fx_bank_process")
        return io

    def set_stream(self, cuda_stream):
        _check(lib().kb_fx_bank_set_stream(self.h, cuda_stream), "kb_fx_bank_set_stream")

    def sync(self):
        _check(lib().kb_fx_bank_sync(self.h), "kb_fx_bank_sync")

    def bytes_per_frame(self):
        return lib().kb_fx_bank_bytes_per_frame(self.h)

    def parallel_instances(self):
        """Instances the last process() ran on the chunk-parallel schedule."""
        return _check(lib().kb_fx_bank_parallel_instances(self.h), "kb_fx_bank_parallel_instances")

    def tolerance_instances(self):
        """Instances the last process(FX_TOLERANCE) ran on a re-associating (1e-5-tolerance) schedule."""
        return _check(lib().kb_fx_bank_tolerance_instances(self.h), "kb_fx_bank_tolerance_instances")

    @property
    def state_bytes(self):
        return lib().kb_fx_bank_state_bytes(self.h)

    def profile(self, enable=True):
        _check(lib().kb_fx_bank_profile(self.h, int(enable)), "kb_fx_bank_profile")

    def profile_read(self):
        """(accumulated milliseconds of the dominant kernel, launches) since profile(True)."""
        ms, cnt = C.c_double(), C.c_longlong()
        _check(lib().kb_fx_bank_profile_read(self.h, C.byref(ms), C.byref(cnt)), "kb_fx_bank_profile_read")
        return ms.value, cnt.value

    @property
    def launches(self):
        return lib().kb_fx_bank_launches(self.h)


class SynthBank:
    """`instances` klang Synth objects with `voices` notes each (Synth::process, klang.h:4440-4466 / 4830-4858)."""

    def __init__(self, graph, instances=1, voices=32, fs=44100.0, max_block=16384, device=0):
        self.h = lib().kb_synth_bank_create(graph, instances, voices, float(fs), max_block, device)
        self._pinned_seen = set()          # pinned host buffers already accepted as `out_prev` (process_mixdown)
        if not self.h:
            raise KlangB200Error("kb_synth_bank_create: " + lib().kb_last_error().decode())
        self.graph, self.instances, self.max_block, self.device = graph, instances, max_block, device
        self.channels = lib().kb_synth_bank_channels(self.h)
        self.voices = lib().kb_synth_bank_voices(self.h)
        self.num_controls = lib().kb_synth_bank_num_controls(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().kb_synth_bank_destroy(self.h)
            self.h = None

    __del__ = close

    def set_control(self, idx, value, instance=None):
        for i in (range(self.instances) if instance is None else [instance]):
            _check(lib().kb_synth_bank_set_control(self.h, i, idx, float(value)), "kb_synth_bank_set_control")

    def get_control(self, idx, instance=0):
        v = C.c_float()
        _check(lib().kb_synth_bank_get_control(self.h, instance, idx, C.byref(v)), "kb_synth_bank_get_control")
        return float(v.value)

    def note_on(self, pitch, velocity, instance=0):
        return _check(lib().kb_synth_bank_note_on(self.h, instance, int(pitch), float(velocity)), "kb_synth_bank_note_on")

    def note_off(self, pitch, velocity=0.0, instance=0):
        _check(lib().kb_synth_bank_note_off(self.h, instance, int(pitch), float(velocity)), "kb_synth_bank_note_off")

    def midi(self, status, byte1, byte2, instance=0):
        """Synth::input(status, byte1, byte2) (templates/juce/synth/Source/klang.h:3921-3929)."""
        _check(lib().kb_synth_bank_midi(self.h, instance, int(status), int(byte1), int(byte2)), "kb_synth_bank_midi")

    def on_control(self, idx, value, instance=0):
        """Synth::onControl (klang.h:4399-4404): returns the number of notes (stage != Off) the reference notifies."""
        rc = lib().kb_synth_bank_on_control(self.h, instance, int(idx), float(value))
        if rc < 0:
            _check(rc, "kb_synth_bank_on_control")
        return rc

    def load_preset(self, index, instance=0):
        """The preset's values through Control::set, then Synth::onPreset (klang.h:4415-4420); returns the notes notified."""
        rc = lib().kb_synth_bank_load_preset(self.h, instance, int(index))
        if rc < 0:
            _check(rc, "kb_synth_bank_load_preset")
        return rc

    def voice_start(self, voice, pitch, velocity, instance=0):
        _check(lib().kb_synth_bank_voice_start(self.h, instance, voice, float(pitch), float(velocity)), "kb_synth_bank_voice_start")

    def voice_release(self, voice, velocity=0.0, instance=0):
        _check(lib().kb_synth_bank_voice_release(self.h, instance, voice, float(velocity)), "kb_synth_bank_voice_release")

    def voice_stage(self, voice, instance=0):
        return _check(lib().kb_synth_bank_voice_stage(self.h, instance, voice), "kb_synth_bank_voice_stage")

    def events(self, ev):
        """Apply a batch of events in order. ev: numpy array of EVENT_DTYPE."""
        ev = np.ascontiguousarray(ev, EVENT_DTYPE)
        _check(lib().kb_synth_bank_events(self.h, len(ev), ev.ctypes.data), "kb_synth_bank_events")

    def out_shape(self, n, flags=0):
        if flags & PER_VOICE:
            return (self.instances, self.voices, self.channels, n)
        if flags & BANK_MIX:
            return (self.channels, n)
        return (self.instances, self.channels, n)

    def process_into(self, out, n, flags=0):
        """out: float32 buffer of out_shape(n, flags), numpy (host, synchronous) or torch CUDA tensor (asynchronous)."""
        if n < 0:
            raise KlangB200Error("kb_synth_bank_process: negative block length")
        p, dev = _ptr(out, int(np.prod(self.out_shape(n, flags))), "kb_synth_bank_process out", self.device)
        _check(lib().kb_synth_bank_process(self.h, p, n, (flags | DEVICE_PTR) if dev else (flags & ~DEVICE_PTR)), "kb_synth_bank_process")
        return out

    def step_into(self, ev, out, n, flags=0):
        """events(ev) then process_into(out, n, flags) in one call into the library (kb_synth_bank_step)."""
        ev = np.ascontiguousarray(ev, EVENT_DTYPE)
        p, dev = _ptr(out, int(np.prod(self.out_shape(n, flags))), "kb_synth_bank_step out", self.device)
        _check(lib().kb_synth_bank_step(self.h, len(ev), ev.ctypes.data, p, n, (flags | DEVICE_PTR) if dev else (flags & ~DEVICE_PTR)), "kb_synth_bank_step")
        return out

    def process_mixdown(self, mixdown, out_prev, n, flags=0):
        """process(KB_BANK_MIX) whose bank-mix kernel stores into rank 0's arena (sharding.PeerMixdown); on rank 0 `out_prev` (a torch CUDA
        tensor [channels, n], or a pinned host tensor the kernel writes over PCIe) receives the rank-order sum of the PREVIOUS block."""
        p = 0
        if out_prev is not None:
            p, dev = _ptr(out_prev, self.channels * n, "kb_synth_bank_process_mixdown out_prev", self.device)
            # device memory, or page-locked host memory (device-mapped under unified addressing: the exchange kernel stores the sum into it)
            if not dev and p not in self._pinned_seen:           # (is_pinned() asks the driver: ~10 us — once per buffer, not once per block)
                if not (hasattr(out_prev, "is_pinned") and out_prev.is_pinned()):
                    raise KlangB200Error("kb_synth_bank_process_mixdown: out_prev must be device memory or a page-locked (pinned) torch tensor")
                self._pinned_seen.add(p)
        _check(lib().kb_synth_bank_process_mixdown(self.h, mixdown.h, p, n, flags), "kb_synth_bank_process_mixdown")

    def step_mixdown(self, ev, mixdown, out_prev_ptr, n, flags=0):
        """events(ev) then process_mixdown in ONE call into the library (kb_synth_bank_step_mixdown).  `out_prev_ptr`: 0 / None, or the raw address
        of a buffer process_mixdown() has accepted before (device memory or pinned host memory) — a per-block host loop checks its buffers once."""
        ev = np.ascontiguousarray(ev, EVENT_DTYPE)
        _check(lib().kb_synth_bank_step_mixdown(self.h, len(ev), ev.ctypes.data, mixdown.h, out_prev_ptr or 0, n, flags), "kb_synth_bank_step_mixdown")

    def process_into_device_ptr(self, ptr, n, flags=0):
        """out = a raw device pointer (e.g. a peer-mapped slot of sharding.PeerMixdown); asynchronous on the bank stream."""
        _check(lib().kb_synth_bank_process(self.h, ptr, n, flags | DEVICE_PTR), "kb_synth_bank_process")

    def process_block(self, n, flags=0):
        return self.process_into(np.empty(self.out_shape(n, flags), np.float32), n, flags)

    def set_stream(self, cuda_stream):
        _check(lib().kb_synth_bank_set_stream(self.h, cuda_stream), "kb_synth_bank_set_stream")

    def sync(self):
        _check(lib().kb_synth_bank_sync(self.h), "kb_synth_bank_sync")

    @property
    def state_bytes(self):
        return lib().kb_synth_bank_state_bytes(self.h)

    def transfer_bytes(self):
        """(host-to-device, device-to-host) bytes this bank has moved so far."""
        a, b = C.c_longlong(), C.c_longlong()
        _check(lib().kb_synth_bank_transfer_bytes(self.h, C.byref(a), C.byref(b)), "kb_synth_bank_transfer_bytes")
        return a.value, b.value

    def profile(self, enable=True):
        _check(lib().kb_synth_bank_profile(self.h, int(enable)), "kb_synth_bank_profile")

    def profile_read(self):
        ms, cnt = C.c_double(), C.c_longlong()
        _check(lib().kb_synth_bank_profile_read(self.h, C.byref(ms), C.byref(cnt)), "kb_synth_bank_profile_read")
        return ms.value, cnt.value

    @property
    def launches(self):
        return lib().kb_synth_bank_launches(self.h)


# ------------------------------------------------------------------------------------------------------------
# Single-object views with the call surface of the reference objects (and of the parity oracles).

class _Fx:
    def __init__(self, eng, graph):
        self.bank = FxBank(graph, 1, eng.fs, eng.max_block, eng.device)
        self.bank.debug_enable(True)
        self.channels, self.num_controls = self.bank.channels, self.bank.num_controls

    def close(self):
        self.bank.close()

    def set_control(self, idx, v):
        self.bank.set_control(idx, v, 0)

    def get_control(self, idx):
        return self.bank.get_control(idx, 0)

    def process(self, x):
        """x: float32 [channels, n] ([n] for mono effects). Returns a processed copy (the reference works in place)."""
        y = np.array(x, np.float32, copy=True, order="C")
        self.bank.process_inplace(y.reshape(1, self.channels, -1))
        self._last_n = y.shape[-1]
        return y

    def debug(self):
        """The `>> debug` capture of the last block, float32 [n], or None (same call as oracle.bindings.Fx.debug)."""
        d = self.bank.debug_read(self._last_n)
        return None if d is None else d[0]


class _Synth:
    def __init__(self, eng, graph, nvoices):
        self.bank = SynthBank(graph, 1, nvoices, eng.fs, eng.max_block, eng.device)
        self.channels, self.nvoices, self.num_controls = self.bank.channels, self.bank.voices, self.bank.num_controls

    def close(self):
        self.bank.close()

    def set_control(self, idx, v):
        self.bank.set_control(idx, v, 0)

    def get_control(self, idx):
        return self.bank.get_control(idx, 0)

    def note_on(self, pitch, vel):
        return self.bank.note_on(pitch, vel, 0)

    def note_off(self, pitch, vel=0.0):
        self.bank.note_off(pitch, vel, 0)

    def voice_start(self, voice, pitch, vel):
        self.bank.voice_start(voice, pitch, vel, 0)

    def voice_release(self, voice, vel=0.0):
        self.bank.voice_release(voice, vel, 0)

    def voice_stage(self, voice):
        return self.bank.voice_stage(voice, 0)

    def process(self, n):
        return self.bank.process_block(n)[0]

    def process_voices(self, n):
        stages = np.array([self.bank.voice_stage(v, 0) for v in range(self.nvoices)], np.int32)
        out = self.bank.process_block(n, PER_VOICE)[0]
        return out, (stages != 3).astype(np.int32)


class Engine:
    """The CUDA path behind the call surface the parity scripts use (tests/cases.py)."""

    def __init__(self, fs=44100.0, device=0, max_block=16384):
        self.fs, self.device, self.max_block = float(fs), device, max_block
        lib()

    def set_fs(self, fs):
        self.fs = float(fs)   # klang::fs is per bank here: applies to banks created afterwards

    def srand(self, seed):
        lib().kb_srand(int(seed))

    def pitch_to_frequency(self, p):
        return float(lib().kb_pitch_to_frequency(float(p)))

    def Fx(self, graph):
        return _Fx(self, graph)

    def Synth(self, graph, nvoices):
        return _Synth(self, graph, nvoices)

    # primitives on the device (one thread) -------------------------------------------------------------
    def osc(self, kind, n, f, phase=None, duty=None):
        out = np.zeros(n, np.float32)
        nargs = 1 if phase is None else (2 if duty is None else 3)
        _check(lib().kb_prim_osc(kind, nargs, float(f), float(phase or 0.0), float(duty or 0.0), self.fs, n, out.ctypes.data), "kb_prim_osc")
        return out

    def _delay_kat(self, x, di, df, set_at):
        x, di = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(di, np.int32)
        df, set_at = np.ascontiguousarray(df, np.float32), np.ascontiguousarray(set_at, np.float32)
        n = len(x)
        oi, of, op, ol = (np.zeros(n, np.float32) for _ in range(4))
        _check(lib().kb_prim_delay(n, x.ctypes.data, di.ctypes.data, df.ctypes.data, set_at.ctypes.data,
                                   oi.ctypes.data, of.ctypes.data, op.ctypes.data, ol.ctypes.data), "kb_prim_delay")
        return oi, of, op, ol

    def delay1000(self, x, di, df, set_at):
        """Delay<1000>: tap(int), tap(float), process() per sample (tests/cases.py)."""
        return self._delay_kat(x, di, df, set_at)[:3]

    def delay_lagrange(self, x, df):
        """Delay<1000>::lagrange(df[s]) after writing x[s] (klang.h:3429-3458)."""
        n = len(x)
        return self._delay_kat(x, np.zeros(n, np.int32), df, np.full(n, -1.0, np.float32))[3]

    def sample(self, table, n, f, phase=None):
        """klang::Sample over `table` (klang.h:3679-3720): set(f) or set(f, phase), then n ticks on the device."""
        table = np.ascontiguousarray(table, np.float32)
        out = np.zeros(n, np.float32)
        _check(lib().kb_prim_sample(table.ctypes.data, len(table), 1 if phase is None else 2, float(f), float(phase or 0.0), n, out.ctypes.data), "kb_prim_sample")
        return out

    def wav_decode(self, image):
        """File::WAV (klang.h:5951-6085) over a file image (bytes): (float32 samples, (channels, samplerate, bits))."""
        return wav_decode(image)

    def wavetable(self, kind):
        """Wavetables::Sine / Saw (kinds 10 / 11): the 2048-entry table as the device reads it."""
        out = np.zeros(2048, np.float32)
        _check(lib().kb_prim_wavetable(kind, self.fs, out.ctypes.data), "kb_prim_wavetable")
        return out

    def stereo_delay1000(self, xl, xr, df):
        """Stereo::Delay<1000>: write {xl[s], xr[s]}, then tap(float df[s]) (klang.h:4668-4681)."""
        xl, xr, df = (np.ascontiguousarray(a, np.float32) for a in (xl, xr, df))
        ol, orr = np.zeros(len(xl), np.float32), np.zeros(len(xl), np.float32)
        _check(lib().kb_prim_stereo_delay(len(xl), xl.ctypes.data, xr.ctypes.data, df.ctypes.data, ol.ctypes.data, orr.ctypes.data), "kb_prim_stereo_delay")
        return ol, orr

    def control_smooth(self, lo, hi, initial, values):
        values = np.ascontiguousarray(values, np.float32)
        out = np.zeros(len(values), np.float32)
        _check(lib().kb_prim_control_smooth(float(lo), float(hi), float(initial), len(values), values.ctypes.data, out.ctypes.data), "kb_prim_control_smooth")
        return out

    def envelope_at(self, points, t):
        xy = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1))
        t = np.ascontiguousarray(t, np.float32)
        out = np.zeros(len(t), np.float32)
        _check(lib().kb_prim_envelope_at(len(xy) // 2, xy.ctypes.data, len(t), t.ctypes.data, out.ctypes.data), "kb_prim_envelope_at")
        return out

    def filt(self, kind, x, f, Q=None, per_sample=False):
        x = np.ascontiguousarray(x, np.float32)
        n = len(x)
        nset = n if per_sample else 1
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(f, np.float32), (nset,)))
        qp = None
        if Q is not None:
            Q = np.ascontiguousarray(np.broadcast_to(np.asarray(Q, np.float32), (nset,)))
            qp = Q.ctypes.data
        out, coeffs = np.zeros(n, np.float32), np.zeros(5, np.float32)
        _check(lib().kb_prim_filter(kind, nset, f.ctypes.data, qp, self.fs, n, x.ctypes.data, out.ctypes.data, coeffs.ctypes.data), "kb_prim_filter")
        return out, coeffs

    def envelope(self, points, n, loop=None, release_at=-1, release_time=0.0, release_level=0.0):
        xy = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1))
        out, stage = np.zeros(n, np.float32), np.zeros(n, np.int32)
        ls, le = loop if loop is not None else (-1, -1)
        _check(lib().kb_prim_envelope(len(xy) // 2, xy.ctypes.data, ls, le, self.fs, n, release_at, release_time, release_level,
                                      out.ctypes.data, stage.ctypes.data), "kb_prim_envelope")
        return out, stage

    def adsr(self, A, D, S, R, n, release_at=-1):
        out, stage = np.zeros(n, np.float32), np.zeros(n, np.int32)
        _check(lib().kb_prim_adsr(A, D, S, R, self.fs, n, release_at, out.ctypes.data, stage.ctypes.data), "kb_prim_adsr")
        return out, stage

    def math(self, fn, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(x)
        _check(lib().kb_prim_math({"sinf": 0, "cosf": 1, "tanhf": 2, "expf": 3}[fn], len(x), x.ctypes.data, out.ctypes.data), "kb_prim_math")
        return out
