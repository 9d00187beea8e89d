"""Shared parity cases.

Every case is a function of an "engine": an object with the interface of oracle.bindings.Oracle
(osc / filt / envelope / adsr / ... primitives, `Fx(graph)` and `Synth(graph, nvoices)` factories).
The same cases are run

  * by tests/gen_golden.py against the compiled reference (oracle.ref) to produce tests/golden/*.npz,
  * by tests/test_oracle.py against the plain-C restatement (oracle.port) — bit-exact vs golden,
  * by tests/test_gpu_parity.py against the CUDA path through the C ABI (klang_b200.Engine).

Inputs are generated from fixed integer hashes (no dependence on numpy's RNG version).
"""
import numpy as np

# graph / primitive ids — identical in oracle/bindings.py and include/klang_b200.h
(OSC_FAST_SAW, OSC_FAST_TRIANGLE, OSC_FAST_SQUARE, OSC_FAST_PULSE, OSC_FAST_SINE,
 OSC_BASIC_SINE, OSC_BASIC_SAW, OSC_BASIC_TRIANGLE, OSC_BASIC_SQUARE, OSC_BASIC_PULSE,
 OSC_WT_SINE, OSC_WT_SAW, OSC_BASIC_NOISE, OSC_FAST_NOISE) = range(14)
(FLT_BIQUAD_LPF, FLT_BIQUAD_HPF, FLT_ONEPOLE_LPF, FLT_ONEPOLE_HPF,
 FLT_BIQUAD_BPF, FLT_BIQUAD_BRF, FLT_BIQUAD_APF, FLT_BUTTERWORTH_LPF1, FLT_BUTTERWORTH_LPF2,
 FLT_DCF, FLT_IIR1, FLT_IIR2, FLT_MODAL, FLT_FOLLOWER_PEAK, FLT_FOLLOWER_RMS, FLT_WINDOW_MEAN, FLT_WINDOW_RMS) = range(17)
FX_GAIN, FX_PINGPONG, FX_REVERB, FX_DELAY_PINGPONG, FX_DELAY_REVERB, FX_PAN, FX_RM, FX_TREMOLO, FX_CLIPPING, FX_ECHO, FX_FEEDBACK, FX_FUNCTIONS, FX_MUTE, FX_IIR, FX_WAHWAH, FX_FLANGER, FX_MODDELAY, FX_MOD_CHORUS = range(18)
SY_SUBTRACTIVE, SY_SUPERSAW, SY_TB303, SY_SYNTHX, SY_FILTER_K, SY_FM, SY_BREAKPOINT, SY_RAMP, SY_RELEASE, SY_ADDITIVE_SAW, SY_ADDITIVE_SQUARE, SY_AM, SY_MOD_FM, SY_MOD_FM2, SY_ADDITIVE_NYQUIST = range(15)

FX_NAMES = {FX_GAIN: "gain", FX_PINGPONG: "pingpong", FX_REVERB: "reverb",
            FX_DELAY_PINGPONG: "delay_pingpong", FX_DELAY_REVERB: "delay_reverb", FX_PAN: "pan", FX_RM: "rm", FX_TREMOLO: "tremolo",
            FX_CLIPPING: "clipping", FX_ECHO: "echo", FX_FEEDBACK: "feedback", FX_FUNCTIONS: "functions", FX_MUTE: "mute", FX_IIR: "iir", FX_WAHWAH: "wahwah", FX_FLANGER: "flanger",
            FX_MODDELAY: "moddelay", FX_MOD_CHORUS: "mod_chorus"}
SY_NAMES = {SY_SUBTRACTIVE: "subtractive", SY_SUPERSAW: "supersaw", SY_TB303: "tb303",
            SY_SYNTHX: "synthx", SY_FILTER_K: "filter_k", SY_FM: "fm", SY_BREAKPOINT: "breakpoint", SY_RAMP: "ramp",
            SY_RELEASE: "release", SY_ADDITIVE_SAW: "additive_saw", SY_ADDITIVE_SQUARE: "additive_square",
            SY_AM: "am", SY_MOD_FM: "mod_fm", SY_MOD_FM2: "mod_fm2", SY_ADDITIVE_NYQUIST: "additive_nyquist"}


def noise(n, seed=1, lo=-1.0, hi=1.0):
    """PCG-style hash noise, uniform [lo, hi) in float32 — reproducible without numpy.random."""
    i = np.arange(n, dtype=np.uint64) + np.uint64((seed * 0x9E3779B97F4A7C15) % (1 << 64))
    x = i * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
    x ^= x >> np.uint64(33)
    x = x * np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33)
    u = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(np.float32)


def voice_pitch(v):
    """pitch 36 + (7 v mod 61)."""
    return 36 + (7 * v) % 61


def voice_velocity(v):
    """0.25 + 0.75 * ((v * 2654435761 mod 2^32) >> 16) / 65535."""
    return 0.25 + 0.75 * (((v * 2654435761) % (1 << 32)) >> 16) / 65535.0


# ----------------------------------------------------------------------------- primitives

def window_follower_cases(eng, x, imp):
    """Envelope::Follower::Window<64> mean / rms (klang.h:5904-5948): moving sum over 64 samples kept in a double, then the AR smoother."""
    out = {}
    for kind, name, f, Q in ((FLT_WINDOW_MEAN, "window_mean", 0.01, 0.1), (FLT_WINDOW_RMS, "window_rms", 0.002, 0.05),
                             (FLT_WINDOW_MEAN, "window_mean_instant", 0.0, 0.02)):
        y, c = eng.filt(kind, x, f, Q)
        out[f"filter/{name}/noise"] = y
        out[f"filter/{name}/coeffs"] = c
        y, c = eng.filt(kind, imp, f, Q)
        out[f"filter/{name}/impulse"] = y
    return out


def noise_cases(eng):
    """Basic::Noise / Fast::Noise (klang.h:4947-4951, 5357-5366): one libc rand() per tick from the process-wide stream —
    seeded runs, and runs that continue the stream where the previous call left it."""
    out = {}
    eng.srand(5)
    out["osc/basic_noise/seed5"] = eng.osc(OSC_BASIC_NOISE, 256, 0.0)
    out["osc/fast_noise/continued"] = eng.osc(OSC_FAST_NOISE, 256, 0.0)
    eng.srand(272839)
    out["osc/fast_noise/seed272839"] = eng.osc(OSC_FAST_NOISE, 1000, 0.0)
    out["osc/basic_noise/continued"] = eng.osc(OSC_BASIC_NOISE, 777, 0.0)
    return out



def wav_image(fmt, channels, bits, payload, rate=44100, extra=b""):
    """A RIFF/WAVE file image: fmt chunk (16 bytes), optional extra chunks, data chunk."""
    import struct
    block = channels * bits // 8
    f = struct.pack("<4sIHHIIHH", b"fmt ", 16, fmt, channels, rate, rate * block, block, bits)
    body = b"WAVE" + f + extra + struct.pack("<4sI", b"data", len(payload)) + payload
    return struct.pack("<4sI", b"RIFF", len(body)) + body


def wav_images():
    """name -> file image: the four encodings File::WAV decodes (klang.h:6063-6083), a chunk to skip, a stereo file, an encoding it leaves as zeros."""
    import struct
    i = np.arange(1500, dtype=np.int64)
    s16 = ((i * 7919) % 65536 - 32768).astype(np.int16)
    s16[:4] = [-32768, 32767, 0, -1]
    u8 = ((i * 37) % 256).astype(np.uint8)
    s32 = ((i * 2654435761) % (1 << 32) - (1 << 31)).astype(np.int32)
    s32[:3] = [-(1 << 31), (1 << 31) - 1, 1]
    f32 = noise(1500, seed=77)
    return {
        "pcm16": wav_image(1, 1, 16, s16.tobytes()),
        "pcm8": wav_image(1, 1, 8, u8.tobytes(), rate=22050),
        "pcm32": wav_image(1, 1, 32, s32.tobytes(), rate=48000),
        "float32": wav_image(3, 1, 32, f32.tobytes(), rate=96000),
        "pcm16_list_chunk": wav_image(1, 1, 16, s16.tobytes(), extra=struct.pack("<4sI", b"LIST", 6) + b"abcdef"),
        "pcm16_stereo": wav_image(1, 2, 16, s16.tobytes()),                  # BlockAlign 4: the first half of the interleaved data
        "pcm24_unsupported": wav_image(1, 1, 24, bytes(300)),                # decoded as zeros (no 24-bit branch)
    }


def sample_wav_cases(eng):
    """File::WAV decode (klang.h:5951-6085) and klang::Sample playback (klang.h:3679-3720) of the decoded table: from the start, and from a
    fractional phase (interpolated reads).  Runs stop before the position reaches the table size (the reference reads samples[size] there)."""
    out = {}
    for name, image in wav_images().items():
        y, info = eng.wav_decode(image)
        out[f"wav/{name}"] = y
        out[f"wav/{name}/info"] = np.array(info, np.int32)
    table = out["wav/float32"]
    out["sample/from_start"] = eng.sample(table, 1400, 440.0)
    out["sample/from_phase"] = eng.sample(table, 1000, 220.0, 0.0101)         # position 445.41: every read interpolates
    out["sample/pcm16_table"] = eng.sample(out["wav/pcm16"], 700, 1000.0, 0.005)
    return out


def primitive_cases(eng, fs):
    """dict name -> float32/int32 array for every primitive on the hot path."""
    eng.set_fs(fs)
    out = {}
    n = 256
    for kind, name in ((OSC_FAST_SAW, "fast_saw"), (OSC_FAST_TRIANGLE, "fast_triangle"), (OSC_FAST_SQUARE, "fast_square"),
                       (OSC_FAST_PULSE, "fast_pulse"), (OSC_FAST_SINE, "fast_sine"), (OSC_BASIC_SINE, "basic_sine"),
                       (OSC_BASIC_SAW, "basic_saw"), (OSC_BASIC_TRIANGLE, "basic_triangle"), (OSC_BASIC_SQUARE, "basic_square"),
                       (OSC_BASIC_PULSE, "basic_pulse"), (OSC_WT_SINE, "wt_sine"), (OSC_WT_SAW, "wt_saw")):
        for f in (441.0, 55.0, 3520.5, 1000.0):
            out[f"osc/{name}/f{f}"] = eng.osc(kind, n, f)
        out[f"osc/{name}/f441_p1"] = eng.osc(kind, n, 441.0, 1.0)
    for kind, name in ((OSC_FAST_SAW, "fast_saw"), (OSC_FAST_TRIANGLE, "fast_triangle"), (OSC_FAST_SQUARE, "fast_square"),
                       (OSC_FAST_PULSE, "fast_pulse"), (OSC_BASIC_PULSE, "basic_pulse")):
        for duty in (0.05, 0.5, 0.93):
            out[f"osc/{name}/f441_p0_d{duty}"] = eng.osc(kind, n, 441.0, 0.0, duty)
            out[f"osc/{name}/f2093_p2_d{duty}"] = eng.osc(kind, n, 2093.0, 2.0, duty)
    out["wavetable/sine"] = eng.wavetable(OSC_WT_SINE)
    out["wavetable/saw"] = eng.wavetable(OSC_WT_SAW)
    out.update(noise_cases(eng))
    out.update(sample_wav_cases(eng))

    x = noise(512, seed=7)
    imp = np.zeros(64, np.float32)
    imp[0] = 1
    sweep = (200.0 + 6000.0 * (0.5 + 0.5 * np.sin(np.arange(512) * 0.01))).astype(np.float32)
    for kind, name in ((FLT_BIQUAD_LPF, "biquad_lpf"), (FLT_BIQUAD_HPF, "biquad_hpf"), (FLT_BIQUAD_BPF, "biquad_bpf"),
                       (FLT_BIQUAD_BRF, "biquad_brf"), (FLT_BUTTERWORTH_LPF2, "butterworth_lpf2")):
        y, c = eng.filt(kind, imp, 1000.0)
        out[f"filter/{name}/impulse"] = y
        out[f"filter/{name}/coeffs"] = c
        y, c = eng.filt(kind, x, 50.0, 1.0)
        out[f"filter/{name}/noise_f50_q1"] = y
        out[f"filter/{name}/coeffs_f50_q1"] = c
        y, _ = eng.filt(kind, x, sweep, 10.0, per_sample=True)
        out[f"filter/{name}/sweep_q10"] = y
    y, c = eng.filt(FLT_BIQUAD_APF, x, 1000.0, 0.5)
    out["filter/biquad_apf/noise"] = y
    for kind, name in ((FLT_ONEPOLE_LPF, "onepole_lpf"), (FLT_ONEPOLE_HPF, "onepole_hpf"), (FLT_BUTTERWORTH_LPF1, "butterworth_lpf1")):
        y, c = eng.filt(kind, imp, 1000.0)
        out[f"filter/{name}/impulse"] = y
        out[f"filter/{name}/coeffs"] = c
        y, _ = eng.filt(kind, x, sweep, None, per_sample=True)
        out[f"filter/{name}/sweep"] = y

    # Filters::DCF / IIR<1> / IIR<2> (klang.h:5387-5464): f = r / coefficient / a1, Q = a2
    for kind, name, f, Q in ((FLT_DCF, "dcf", 0.995, None), (FLT_DCF, "dcf_r09", 0.9, None), (FLT_IIR1, "iir1", 0.25, None),
                             (FLT_IIR2, "iir2", -1.6, 0.8)):
        y, c = eng.filt(kind, x, f, Q)
        out[f"filter/{name}/noise"] = y
        out[f"filter/{name}/coeffs"] = c
        y, c = eng.filt(kind, imp, f, Q)
        out[f"filter/{name}/impulse"] = y

    # Modifiers::Modal (f, decay) and Envelope::Follower peak / rms (attack, release)   klang.h:5817-5896
    for kind, name, f, Q in ((FLT_MODAL, "modal", 440.0, 0.25), (FLT_MODAL, "modal_hi", 7040.0, 0.01),
                             (FLT_FOLLOWER_PEAK, "follower_peak", 0.01, 0.1), (FLT_FOLLOWER_RMS, "follower_rms", 0.002, 0.05),
                             (FLT_FOLLOWER_PEAK, "follower_peak_instant", 0.0, 0.02)):
        y, c = eng.filt(kind, x, f, Q)
        out[f"filter/{name}/noise"] = y
        out[f"filter/{name}/coeffs"] = c
        y, c = eng.filt(kind, imp, f, Q)
        out[f"filter/{name}/impulse"] = y

    out.update(window_follower_cases(eng, x, imp))

    y, st = eng.envelope([(0, 0), (0.001, 1), (0.003, 0.25), (0.005, 0.5)], 400)
    out["envelope/4pt"] = y
    out["envelope/4pt_stage"] = st
    y, st = eng.envelope([(0, 100), (0.002, 1000), (0.004, 500)], 400, release_at=120, release_time=0.002, release_level=0.0)
    out["envelope/3pt_release"] = y
    out["envelope/3pt_release_stage"] = st
    y, st = eng.envelope([(0, 0), (0.001, 1), (0.002, 0.5), (0.003, 0.8)], 600, loop=(1, 3))
    out["envelope/loop13"] = y
    out["envelope/at"] = eng.envelope_at([(0, 0.0625), (0.25, 0.125), (0.5, 0.25), (1.0, 1.0)],
                                         np.linspace(-0.1, 1.2, 53).astype(np.float32))
    y, st = eng.adsr(0.01, 0.1, 0.7, 0.25, 20000, release_at=6000)
    out["adsr/a"] = y
    out["adsr/a_stage"] = st
    y, st = eng.adsr(0.0, 0.0, 1.0, 0.001, 600, release_at=100)
    out["adsr/zero_attack"] = y
    out["adsr/zero_attack_stage"] = st
    y, st = eng.adsr(0.001, 0.25, 1.0, 0.5, 2000, release_at=20)   # release during attack
    out["adsr/early_release"] = y

    n = 1500
    xin = (np.arange(n) + 1).astype(np.float32)
    di = (np.arange(n) * 7 % 900).astype(np.int32)
    df = (noise(n, seed=3, lo=0.0, hi=998.0)).astype(np.float32)
    set_at = np.full(n, -1.0, np.float32)
    set_at[10] = 4.0
    set_at[700] = 333.25
    set_at[1200] = 999.5
    oi, of, op = eng.delay1000(xin, di, df, set_at)
    out["delay/tap_int"], out["delay/tap_float"], out["delay/process"] = oi, of, op
    out["delay/lagrange"] = eng.delay_lagrange(noise(n, seed=6), df)      # Delay::lagrange, third order (klang.h:3429-3458)
    ol, orr = eng.stereo_delay1000(noise(n, seed=4), noise(n, seed=5), df)
    out["delay/stereo_l"], out["delay/stereo_r"] = ol, orr
    vals = np.where(np.arange(2000) < 1000, 0.8, 0.1).astype(np.float32)
    out["control/smooth"] = eng.control_smooth(0.0, 1.0, 0.5, vals)
    out["pitch/frequency"] = np.array([eng.pitch_to_frequency(p) for p in range(128)], np.float32)
    return out


# -------------------------------------------------------------------------------- effects

def fx_input(channels, n, seed, burst=None):
    """PCG noise in [-0.5, 0.5); if `burst` only the first `burst` frames are non-zero."""
    x = np.stack([noise(n, seed=seed * 2 + c, lo=-0.5, hi=0.5) for c in range(channels)])
    if burst is not None:
        x[:, burst:] = 0
    return x


FX_SCRIPTS = {
    # name: (graph, total frames, block, [(block_index, control, value)], burst)
    "gain": (FX_GAIN, 4096, 1024, [(2, 0, 0.25)], None),
    "pingpong_default": (FX_PINGPONG, 6144, 1024, [], None),
    # short delay so the feedback path and the moving read head are exercised, plus scratch LFO
    "pingpong_short": (FX_PINGPONG, 8192, 1024, [(0, 1, 0.02), (0, 5, 0.02), (0, 0, 0.9), (0, 4, 0.3), (3, 2, 0.4), (3, 3, 0.579)], None),
    "reverb_default": (FX_REVERB, 4096, 1024, [], 1024),
    "reverb_hall": (FX_REVERB, 6144, 2048, [(0, 0, 1.0), (0, 2, 0.419), (0, 3, 0.329), (0, 7, 0.5), (0, 8, 0.5), (1, 6, 0.4)], None),
    "delay_pingpong": (FX_DELAY_PINGPONG, 6144, 1024, [(0, 0, 0.01), (0, 1, 0.7), (0, 2, 0.02), (0, 3, 0.6)], None),
    "delay_reverb": (FX_DELAY_REVERB, 8192, 4096, [(1, 1, 0.05)], None),
}


# examples/Gain/{Pan,RM,Tremolo}.k and Distortion/Clipping.k: elementwise effects (RM / Tremolo with a Fast::Sine LFO).  Kept apart from
# FX_SCRIPTS: their device tests live in tests/test_zz_gpu_primitives.py.
FX_SCRIPTS_LATE = {
    "pan": (FX_PAN, 3000, 1000, [(1, 0, 0.2), (2, 0, 1.0)], None),
    "rm": (FX_RM, 4096, 1024, [(2, 0, 440.0)], None),
    "rm_at_cached_rate": (FX_RM, 3072, 1024, [(0, 0, 1000.0), (2, 0, 999.0)], None),      # set(1000) on a fresh Sine is a no-op: silent LFO (Q3)
    "tremolo": (FX_TREMOLO, 4500, 1500, [(1, 1, 0.2), (2, 0, 10.0)], None),
    "clipping": (FX_CLIPPING, 3072, 1024, [(1, 0, 3.0), (2, 0, 11.0)], None),
    # Delay/Echo.k (feed-forward tap) and Delay/Feedback.k (the line is fed the output): short delays so echoes land inside the run,
    # a delay time that moves between blocks, and a fractional delay (0.0123 s x fs)
    "echo": (FX_ECHO, 6144, 1024, [(0, 0, 0.01), (0, 1, 0.7), (3, 0, 0.0123)], 4096),
    "feedback": (FX_FEEDBACK, 8192, 1024, [(0, 0, 0.005), (0, 1, 0.8), (4, 0, 0.0123), (6, 1, 0.3)], 1500),
    "feedback_zero_delay": (FX_FEEDBACK, 2048, 1024, [(0, 0, 0.0), (0, 1, 0.5)], None),
    # Distortion/Functions.k (hardclip(in * gain), gain up to 25) and Distortion/Mute.k (a Toggle: in * 0 keeps the sign of zero)
    "functions": (FX_FUNCTIONS, 3072, 1024, [(1, 0, 4.0), (2, 0, 25.0)], None),
    "mute": (FX_MUTE, 3072, 1024, [(1, 0, 1.0), (2, 0, 0.0)], None),
    # Filtering/IIR.k: one-pole smoother with coefficient cube(control), state carried across blocks
    "iir": (FX_IIR, 4096, 1024, [(1, 0, 0.2), (2, 0, 0.9), (3, 0, 0.0)], 3000),
    # Filtering/WahWah.k: Biquad::LPF whose cutoff is set every sample from a squared sine LFO (cosf / sinf per sample)
    "wahwah": (FX_WAHWAH, 6144, 1024, [(2, 0, 4000.0), (2, 1, 8.0), (4, 2, 10.0)], None),
    # Modulation/Flanger.k (triangle LFO), ModDelay.k (sine LFO, smoothed depth control), Chorus.k (three sine LFOs): one delay line tapped
    # at a delay that moves every sample
    "flanger": (FX_FLANGER, 8192, 1024, [(3, 0, 1.0), (3, 1, 5.0)], None),
    "moddelay": (FX_MODDELAY, 8192, 1024, [(2, 1, 1.0), (5, 0, 10.0)], None),
    "mod_chorus": (FX_MOD_CHORUS, 8192, 1024, [], None),
}


# scripts whose program taps a signal with `>> debug` (PingPong.k:61, RM.k:22, Tremolo.k:27, ModDelay.k:24): the capture of every block is a case too
FX_DEBUG_TAPS = ("pingpong_default", "pingpong_short", "rm", "rm_at_cached_rate", "tremolo", "moddelay")


# Programs the product has no hand-written graph for and runs from their own source (klang_b200/kcc.py): reference-only effect ids from 100
# (oracle/ref_harness.cpp); their golden vectors live in tests/golden/klang_ref_translated_fs*.npz.
#   name: (reference id, path under examples/, total frames, block, [(block_index, control, value)], burst)
TRANSLATED_FX_SCRIPTS = {
    "k_objects": (100, "Filtering/Objects.k", 4096, 1024, [(2, 0, 300.0), (2, 1, 4.0)], None),          # Noise >> LPF: one libc rand() per sample
    "k_bands": (101, "Filtering/Bands.k", 4096, 1024, [(1, 0, 250.0), (1, 3, 5.0), (3, 2, 900.0)], None),
    "k_eq": (102, "Filtering/EQ.k", 4096, 1024, [(1, 0, 0.9), (2, 1, 0.1), (3, 2, 1.0)], None),
    "k_patterns": (103, "Delay/Patterns.k", 131072, 16384, [(3, 0, 1.0), (6, 0, 2.0)], 3000),              # a Menu control picks one of three tap patterns (taps up to 1.5 s)
    "k_shaping": (105, "Distortion/Shaping.k", 4096, 1024, [(1, 0, 2.5), (2, 0, 5.6), (3, 0, 0.4)], None),      # Function<float, float>: `in >> f(distort) >> out` = tanh(c x) / tanh(c)
    "k_reverb2": (104, "Delay/Reverb2.k", 16384, 4096, [(1, 0, 0.45), (2, 1, 0.05), (2, 2, 4000.0)], None),  # Stereo::Effect: in[c], out.l >> feedback[0]
}

# The same for synths: reference-only synth ids from 100.
#   name: (reference id, path under examples/, nvoices, started voices, blocks, block size, release block base, [(block, control, value)])
TRANSLATED_SYNTH_SCRIPTS = {
    "k_expression": (100, "Subtractive/Expression.k", 32, 8, 10, 4096, 4, []),   # three Saws with an enveloped vibrato LFO, LPF; on() draws random() four times per note
    # six Sine partials of the program's own Oscillator, each scaled by `GAIN[o] -> Amplitude` (dB -> expf) per sample; namespace-scope constants
    "k_resynthesis": (101, "Additive/Resynthesis.k", 32, 8, 6, 2048, 2, []),
    # three Operator<Sine> in series (`(op1 * c1) >> (op2 * c3) >> op3) * adsr >> dcfilter`: the ADSR becomes op3's amplitude), Biquad HPF; controls read in on()
    "k_operators": (102, "Modulation/Operators.k", 32, 8, 6, 2048, 2, [(0, 0, 1.0), (0, 1, 0.5), (0, 2, 1.0), (0, 3, 4.296), (0, 4, 2.0), (3, 1, 2.5)]),
}


def run_fx_script(eng, name, fs, seed=1, debug=False):
    """debug=True returns (output, the concatenated `>> debug` captures of the blocks); a block without a capture raises."""
    if name in TRANSLATED_FX_SCRIPTS:
        graph, _, total, block, events, burst = TRANSLATED_FX_SCRIPTS[name]
    else:
        graph, total, block, events, burst = (FX_SCRIPTS.get(name) or FX_SCRIPTS_LATE[name])
    eng.set_fs(fs)
    eng.srand(1)
    fx = eng.Fx(graph)
    x = fx_input(fx.channels, total, seed, burst)
    if fx.channels == 1:
        x = x[0]
    ys, ds = [], []
    for b in range(total // block):
        for (bi, c, v) in events:
            if bi == b:
                fx.set_control(c, v)
        ys.append(fx.process(x[..., b * block:(b + 1) * block]))
        if debug:
            d = fx.debug()
            assert d is not None and d.shape == (block,), f"{name}: block {b} left no debug capture"
            ds.append(d)
    fx.close()
    return (np.concatenate(ys, axis=-1), np.concatenate(ds)) if debug else np.concatenate(ys, axis=-1)


# --------------------------------------------------------------------------------- synths

SYNTH_SCRIPTS = {
    # name: (graph, nvoices, started voices, blocks, block size, release block base, [(block, control, value)])
    "subtractive": (SY_SUBTRACTIVE, 32, 12, 8, 512, 2, []),
    "subtractive_fast_release": (SY_SUBTRACTIVE, 32, 8, 6, 1024, 1, [(0, 0, 0.001), (0, 1, 0.01), (0, 3, 0.01)]),
    "filter_k": (SY_FILTER_K, 32, 6, 5, 512, 2, []),
    "supersaw": (SY_SUPERSAW, 32, 10, 6, 512, 2, []),
    "supersaw_wide": (SY_SUPERSAW, 32, 6, 4, 512, 1, [(0, 1, 0.5), (0, 2, 1.0), (0, 0, 0.01)]),
    "tb303": (SY_TB303, 32, 8, 6, 512, 2, []),
    "tb303_square": (SY_TB303, 32, 6, 6, 512, 2, [(0, 3, 1.0), (0, 1, 0.9), (0, 4, 3.0), (2, 0, 0.4)]),
    "synthx": (SY_SYNTHX, 32, 4, 4, 256, 1, [(0, 0, 0.01)]),
    "fm": (SY_FM, 32, 10, 6, 512, 2, [(0, 3, 0.01)]),
    "fm_deep": (SY_FM, 32, 6, 5, 512, 1, [(0, 0, 2.5), (0, 1, 3.0), (0, 2, 7.5), (0, 3, 0.002), (2, 1, 0.5)]),
}


# examples/Subtractive/{Breakpoint,Ramp,Release}.k (a Fast::Sine times one breakpoint envelope).  Kept apart from SYNTH_SCRIPTS:
# their device tests live in tests/test_zz_gpu_primitives.py.
SYNTH_SCRIPTS_LATE = {
    "breakpoint": (SY_BREAKPOINT, 32, 8, 12, 512, 4, []),                  # noteOff cuts the note (NoteBase::off default, klang.h:4237)
    "ramp": (SY_RAMP, 32, 6, 12, 512, 8, [(0, 0, 0.1)]),
    "release": (SY_RELEASE, 32, 10, 10, 512, 2, [(0, 3, 0.03)]),
    "release_slow_attack": (SY_RELEASE, 32, 5, 8, 512, 1, [(0, 0, 0.02), (0, 1, 0.01), (0, 2, 0.6), (0, 3, 0.01)]),
    # Additive/Saw.k, Additive/Square.k: 32 Fast::Sine partials summed in order (Square.k: odd harmonics below Nyquist only)
    "additive_saw": (SY_ADDITIVE_SAW, 32, 10, 5, 512, 2, []),
    "additive_square": (SY_ADDITIVE_SQUARE, 32, 10, 5, 512, 2, []),
    "additive_nyquist": (SY_ADDITIVE_NYQUIST, 32, 10, 5, 512, 2, []),            # Nyquist.k: every partial below Nyquist
    # Modulation/AM.k, FM.k, FM2.k: sine carrier with one or two sine modulators whose frequency is set every sample
    "am": (SY_AM, 32, 8, 6, 512, 2, [(3, 0, 2.2), (3, 1, 0.9)]),
    "mod_fm": (SY_MOD_FM, 32, 8, 6, 512, 2, [(3, 0, 1.5), (3, 1, 7.0)]),
    "mod_fm2": (SY_MOD_FM2, 32, 8, 6, 512, 2, [(0, 0, 3.0), (0, 1, 10.0), (0, 2, 6.791), (4, 2, 1.0)]),
}


def run_synth_script(eng, name, fs, per_voice=True):
    """Drive voices through start/release/process like Synth::process does (klang.h:4440-4466).

    Returns dict with 'voices' [blocks][V, C, n] concatenated over time (per-voice streams, each voice
    rendered alone — SURVEY Q6) or 'mix' (the Synth::process block output), plus 'stages'."""
    if name in TRANSLATED_SYNTH_SCRIPTS:
        graph, _, nvoices, started, blocks, n, rel0, events = TRANSLATED_SYNTH_SCRIPTS[name]
    else:
        graph, nvoices, started, blocks, n, rel0, events = (SYNTH_SCRIPTS.get(name) or SYNTH_SCRIPTS_LATE[name])
    eng.set_fs(fs)
    eng.srand(1)
    sy = eng.Synth(graph, nvoices)
    outs, stages = [], []
    for b in range(blocks):
        for (bi, c, v) in events:
            if bi == b:
                sy.set_control(c, v)
        if b == 0:
            for v in range(started):
                sy.voice_start(v, voice_pitch(v), voice_velocity(v))
        for v in range(started):
            if b == rel0 + (v % 3):
                sy.voice_release(v, 0.0)
        if per_voice:
            o, _ = sy.process_voices(n)
            outs.append(o[:started])
        else:
            outs.append(sy.process(n))
        stages.append([sy.voice_stage(v) for v in range(started)])
    sy.close()
    return {"out": np.concatenate(outs, axis=-1), "stages": np.array(stages, np.int32)}


def run_synth_noteon_script(eng, graph, fs, nvoices=32, notes=40, blocks=4, n=256):
    """Synth::noteOn / noteOff with voice stealing (Notes::assign, klang.h:4336-4372)."""
    eng.set_fs(fs)
    eng.srand(1)
    sy = eng.Synth(graph, nvoices)
    assigned, outs = [], []
    k = 0
    for b in range(blocks):
        for _ in range(notes // blocks):
            assigned.append(sy.note_on(voice_pitch(k), voice_velocity(k)))
            k += 1
        if b >= 1:
            for j in range(5):
                sy.note_off(voice_pitch(j + 5 * (b - 1)), 0.0)
        outs.append(sy.process(n))
    sy.close()
    return {"out": np.concatenate(outs, axis=-1), "assigned": np.array(assigned, np.int32)}


def all_graph_cases(eng, fs):
    out = {}
    for name in list(FX_SCRIPTS) + list(FX_SCRIPTS_LATE):
        if name in FX_DEBUG_TAPS:
            out[f"fx/{name}"], out[f"fx/{name}/debug"] = run_fx_script(eng, name, fs, debug=True)
        else:
            out[f"fx/{name}"] = run_fx_script(eng, name, fs)
    for name in list(SYNTH_SCRIPTS) + list(SYNTH_SCRIPTS_LATE):
        r = run_synth_script(eng, name, fs, per_voice=True)
        out[f"synth/{name}/voices"] = r["out"]
        out[f"synth/{name}/stages"] = r["stages"]
        r = run_synth_script(eng, name, fs, per_voice=False)
        out[f"synth/{name}/mix"] = r["out"]
    for graph in (SY_SUBTRACTIVE, SY_SYNTHX):
        r = run_synth_noteon_script(eng, graph, fs)
        out[f"synth/{SY_NAMES[graph]}/noteon_mix"] = r["out"]
        out[f"synth/{SY_NAMES[graph]}/noteon_assigned"] = r["assigned"]
    return out


def translated_cases(eng, fs):
    """The scripts of TRANSLATED_FX_SCRIPTS (reference side: eng = oracle.ref; product side: an engine that hands out translated programs)."""
    return {f"fx/{name}": run_fx_script(eng, name, fs) for name in TRANSLATED_FX_SCRIPTS}


def translated_synth_cases(eng, name, fs):
    """One script of TRANSLATED_SYNTH_SCRIPTS: per-voice streams, note stages and the Synth::process mix."""
    r = run_synth_script(eng, name, fs, per_voice=True)
    out = {f"synth/{name}/voices": r["out"], f"synth/{name}/stages": r["stages"]}
    out[f"synth/{name}/mix"] = run_synth_script(eng, name, fs, per_voice=False)["out"]
    return out
