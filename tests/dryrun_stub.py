"""Dry run of the late GPU tests' PYTHON on a machine without a GPU (test infrastructure, never collected by pytest): the names
klang_b200 exports are replaced by stand-ins backed by the oracle port, so every test body executes — shapes, keys, argument
order, golden lookups — and trivially passes (oracle vs oracle).  It proves nothing about the CUDA path; it only keeps typos out
of tests whose first real run is on the GPU box.  Usage: python tests/dryrun_stub.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import oracle  # noqa: E402
import klang_b200 as real  # noqa: E402

stub = types.ModuleType("klang_b200")
for name in dir(real):
    if name.isupper():
        setattr(stub, name, getattr(real, name))
stub.device_count = lambda: 1
stub.presets = real.presets
stub.wav_decode = lambda image: oracle.port.wav_decode(image)


class Engine:
    def __init__(self):
        self.fs = 44100.0

    def set_fs(self, fs):
        self.fs = float(fs)
        oracle.port.set_fs(fs)

    def __getattr__(self, name):                      # osc, filt, srand, delay1000, delay_lagrange, Synth, Fx, ...
        return getattr(oracle.port, name)


class SynthBank:
    def __init__(self, graph, instances, voices, fs, max_block, device=0):
        oracle.port.set_fs(fs)
        self.s = [oracle.port.Synth(graph, voices) for _ in range(instances)]
        self.instances, self.voices, self.channels, self.num_controls = instances, self.s[0].nvoices, self.s[0].channels, self.s[0].num_controls

    def set_control(self, idx, v, instance=None):
        for i in (range(self.instances) if instance is None else [instance]):
            self.s[i].set_control(idx, v)

    def voice_start(self, voice, pitch, vel, instance=0):
        self.s[instance].voice_start(voice, pitch, vel)

    def voice_release(self, voice, vel=0.0, instance=0):
        self.s[instance].voice_release(voice, vel)

    def voice_stage(self, voice, instance=0):
        return self.s[instance].voice_stage(voice)

    def note_on(self, pitch, vel, instance=0):
        return self.s[instance].note_on(pitch, vel)

    def note_off(self, pitch, vel=0.0, instance=0):
        self.s[instance].note_off(pitch, vel)

    def on_control(self, idx, value, instance=0):
        return self.s[instance].on_control(idx, value)

    def load_preset(self, index, instance=0):
        self.s[instance].load_preset(index)
        return self.s[instance].on_control(0, 0.0)

    def midi(self, st, b1, b2, instance=0):
        if st == 0x90 and b2 > 0:
            self.note_on(b1, float(np.float32(b2) / np.float32(127)), instance)
        elif st == 0x80 or (st == 0x90 and b2 == 0):
            self.note_off(b1, 0.0, instance)

    def events(self, ev):
        for e in ev:
            t, i, key = int(e["type"]), int(e["instance"]), int(e["key"])
            if t == real.EV_NOTE_ON:
                self.note_on(key, float(e["velocity"]), i)
            elif t == real.EV_NOTE_OFF:
                self.note_off(key, float(e["velocity"]), i)
            elif t == real.EV_VOICE_START:
                self.voice_start(key, float(e["pitch"]), float(e["velocity"]), i)
            elif t == real.EV_VOICE_RELEASE:
                self.voice_release(key, float(e["velocity"]), i)
            else:
                self.set_control(key, float(e["velocity"]), i)

    def process_block(self, n, flags=0):
        if flags & real.PER_VOICE:
            return np.stack([s.process_voices(n)[0] for s in self.s])
        return np.stack([np.atleast_2d(s.process(n)) for s in self.s])

    def close(self):
        for s in self.s:
            s.close()


class FxBank:
    def __init__(self, graph, instances=1, fs=44100.0, max_block=16384, device=0):
        oracle.port.set_fs(fs)
        self.f = [oracle.port.Fx(graph) for _ in range(instances)]
        self.instances, self.channels, self.num_controls = instances, self.f[0].channels, self.f[0].num_controls

    def set_control(self, idx, v, instance=None):
        for i in (range(self.instances) if instance is None else [instance]):
            self.f[i].set_control(idx, v)

    def process_inplace(self, io, n=None, flags=0):
        for i in range(self.instances):
            y = self.f[i].process(io[i][0] if self.channels == 1 else io[i])
            io[i] = np.atleast_2d(y)
        return io

    def parallel_instances(self):
        return self.instances

    def load_preset(self, index, instance=None):
        for i in (range(self.instances) if instance is None else [instance]):
            self.f[i].load_preset(index)

    def get_control(self, idx, instance=0):
        return self.f[instance].get_control(idx)

    def debug_enable(self, on=True):
        self.dbg = on

    def debug_read(self, n):
        if not getattr(self, "dbg", False):
            raise real.KlangB200Error("debug capture is not enabled")
        d = [f.debug() for f in self.f]
        return None if d[0] is None else np.stack(d)

    def close(self):
        for f in self.f:
            f.close()


stub.Engine, stub.SynthBank, stub.FxBank, stub.KlangB200Error = Engine, SynthBank, FxBank, real.KlangB200Error
stub.lib = real.lib
sys.modules["klang_b200"] = stub

import pytest  # noqa: E402

if __name__ == "__main__":
    # subprocess-based tests (k_host, probes) cannot be dry-run: deselect them
    # (the far-end and validation tests of test_gpu_baseline_shapes.py assert on the schedule the CUDA library picked / on its argument checks)
    sys.exit(pytest.main([os.path.join(HERE, "test_zz_gpu_primitives.py"), os.path.join(HERE, "test_gpu_baseline_shapes.py"), "-q", "-x", "-p", "no:cacheprovider",
                          "-k", "not k_programs and not far_end and not validation and not tolerance_mode and not pinned_host_buffer"]))
