"""CPU tests of the host-visible logic of the product sources: the run-length envelope and the closed-form
oscillator (klang_b200/csrc/kb_prims.cuh, compiled here as plain C++) are bit-identical to the per-tick forms."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_run_length_envelope_and_closed_form_oscillator():
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "env_run_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "env_run_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's keys;
    it needs no GPU: the compiled reference (or the C port) on the host cores."""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "voice-samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["config"]["block"] == 4096 and d["config"]["voices_per_instance"] == 128
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_clock_sampler_without_nvml_reports_nothing():
    """The bench's clock sampler never raises and reports `sm_mhz: None` when neither NVML nor nvidia-smi can see a GPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("kb_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0).start()
    s.sample()
    r = s.stop()
    assert "sm_mhz" in r and "reasons" in r


def test_fm_voice_product_functions_match_oracle_on_host():
    """The product's FM.k voice (kb_fm_on / kb_fm_tick, the functions kb_voice_kernel<KB_SY_FM> runs per lane) compiled with
    g++ equals the oracle port bit for bit over two scenarios (release included)."""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    import oracle
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "fm_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "fm_check.cpp"), "-o", exe])
    got = np.frombuffer(subprocess.run([exe], capture_output=True).stdout, np.float32)

    def port(fs, pitch, ctl, n, rel):
        oracle.port.set_fs(fs)
        oracle.port.srand(1)
        sy = oracle.port.Synth(oracle.SY_FM, 32)
        for i, v in enumerate(ctl):
            sy.set_control(i, v)
        sy.voice_start(0, pitch, 0.8)
        a, _ = sy.process_voices(rel)
        sy.voice_release(0, 0.0)
        b, _ = sy.process_voices(n - rel)
        sy.close()
        return np.concatenate([a[0, 0], b[0, 0]])

    want = np.concatenate([port(48000.0, 60, (1.0, 0.37, 0.37, 0.5), 3000, 1500), port(44100.0, 72, (2.5, 3.0, 7.5, 0.002), 2000, 700)])
    assert len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(want).max() > 0.05


def test_fm_time_parallel_form_equals_per_tick_form():
    """kb_fm_at / kb_fm_block_end with kb_envr_run rows (what the opt-in kb_fm_tiled_kernel runs) equal kb_fm_tick bit for bit,
    samples and the state left behind, over ragged blocks, mid-tile releases, notes running into Off and re-triggers."""
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "fm_tiled_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "fm_tiled_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches, 0 state mismatches" in out.stdout


def test_device_rand_restates_glibc_and_attaches_to_the_live_stream():
    """klang_b200/csrc/kb_rand.h against this box's libc: srand()/rand() draw for draw over 12 seeds, jump-ahead == sequential stepping,
    capture / commit read and advance the LIVE libc stream, and the two Noise maps equal the reference's expressions."""
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "rand_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "rand_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches, 0 jump mismatches, 0 live-stream mismatches, 0 noise mismatches" in out.stdout


def test_delay_known_answer_loop_matches_reference_golden_on_host():
    """kb_delay_kat (write / tap(int) / tap(float) / lagrange / set / process of one Delay<1000>, the loop kb_prim_delay_kernel runs)
    compiled with g++ reproduces the compiled reference's golden vectors bit for bit, Delay::lagrange included."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    tmp = tempfile.mkdtemp(prefix="kb_host_")
    exe = os.path.join(tmp, "delay_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "delay_check.cpp"), "-o", exe])

    class HostDelay:                         # the delay entries of the engine interface, backed by the host build of the product loop
        def run(self, x, di, df, set_at):
            n = len(x)
            path = os.path.join(tmp, "in.bin")
            with open(path, "wb") as f:
                f.write(np.int32(n).tobytes())
                for a, t in ((x, np.float32), (di, np.int32), (df, np.float32), (set_at, np.float32)):
                    f.write(np.ascontiguousarray(a, t).tobytes())
            out = subprocess.run([exe, path], capture_output=True)
            assert out.returncode == 0
            return np.frombuffer(out.stdout, np.float32).reshape(4, n)

    g = np.load(os.path.join(ROOT, "tests", "golden", "klang_ref_fs48000.npz"))
    n = 1500                                 # the inputs of cases.primitive_cases
    xin = (np.arange(n) + 1).astype(np.float32)
    di = (np.arange(n) * 7 % 900).astype(np.int32)
    df = cases.noise(n, seed=3, lo=0.0, hi=998.0).astype(np.float32)
    set_at = np.full(n, -1.0, np.float32)
    set_at[10], set_at[700], set_at[1200] = 4.0, 333.25, 999.5
    h = HostDelay()
    oi, of, op, _ = h.run(xin, di, df, set_at)
    lag = h.run(cases.noise(n, seed=6), np.zeros(n, np.int32), df, np.full(n, -1.0, np.float32))[3]
    for got, key in ((oi, "delay/tap_int"), (of, "delay/tap_float"), (op, "delay/process"), (lag, "delay/lagrange")):
        assert np.array_equal(got.view(np.uint32), g[key].view(np.uint32)), key
    assert np.abs(lag).max() > 0.1


def test_window_follower_loop_matches_reference_golden_on_host():
    """kb_window_follower_run (Envelope::Follower::Window<64> mean / rms, the loop kb_prim_filter_kernel runs for kinds 15 / 16)
    compiled with g++ reproduces the compiled reference's golden vectors bit for bit (double moving sum included)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    tmp = tempfile.mkdtemp(prefix="kb_host_")
    exe = os.path.join(tmp, "window_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "window_check.cpp"), "-o", exe])
    x = cases.noise(512, seed=7)
    imp = np.zeros(64, np.float32)
    imp[0] = 1
    for fs in (44100, 48000):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"klang_ref_fs{fs}.npz"))
        for name, rms in (("window_mean", 0), ("window_rms", 1), ("window_mean_instant", 0)):
            c = g[f"filter/{name}/coeffs"]
            for sig, key in ((x, "noise"), (imp, "impulse")):
                path = os.path.join(tmp, "in.bin")
                with open(path, "wb") as f:
                    f.write(np.int32(rms).tobytes() + c[0:1].tobytes() + c[1:2].tobytes() + np.int32(len(sig)).tobytes() + sig.tobytes())
                out = subprocess.run([exe, path], capture_output=True)
                assert out.returncode == 0
                got = np.frombuffer(out.stdout, np.float32)
                assert np.array_equal(got[:len(sig)].view(np.uint32), g[f"filter/{name}/{key}"].view(np.uint32)), (fs, name, key)
                if key == "noise":
                    assert np.array_equal(got[len(sig):].view(np.uint32), c.view(np.uint32)), (fs, name, "coeffs")


def test_sine_envelope_voices_match_oracle_on_host():
    """Breakpoint.k / Ramp.k / Release.k: the product's voice functions (kb_senv_on / kb_senv_tick) compiled with g++ equal the oracle
    port bit for bit, the cut on noteOff (NoteBase::off default) and Release.k's release() included."""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    import oracle
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "senv_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "senv_check.cpp"), "-o", exe])
    got = np.frombuffer(subprocess.run([exe], capture_output=True).stdout, np.float32)

    def port(graph, fs, pitch, ctl, n, rel):
        oracle.port.set_fs(fs)
        oracle.port.srand(1)
        sy = oracle.port.Synth(graph, 32)
        for i, v in ctl:
            sy.set_control(i, v)
        sy.voice_start(0, pitch, 0.8)
        if rel < 0:
            out = sy.process_voices(n)[0][0, 0]
        else:
            a, _ = sy.process_voices(rel)
            sy.voice_release(0, 0.0)
            b, _ = sy.process_voices(n - rel)
            out = np.concatenate([a[0, 0], b[0, 0]])
        sy.close()
        return out

    want = np.concatenate([port(oracle.SY_BREAKPOINT, 48000.0, 60, (), 6000, 5500), port(oracle.SY_RAMP, 44100.0, 72, (), 6000, -1),
                           port(oracle.SY_RELEASE, 48000.0, 45, ((1, 0.01), (2, 0.5), (3, 0.02)), 4000, 1500)])
    assert len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(want[:6000]).max() > 0.5 and np.abs(want[12000:]).max() > 0.3


def test_elementwise_effects_match_reference_golden_on_host():
    """Pan.k / RM.k / Tremolo.k / Clipping.k: the product's per-sample function (kb_ew_sample, closed-form LFO phase) with the library's
    per-block steps, compiled with g++, reproduces the compiled reference's golden vectors bit for bit — the silent LFO of a fresh
    Sine set to its cached 1000 Hz (Q3) included."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    tmp = tempfile.mkdtemp(prefix="kb_host_")
    exe = os.path.join(tmp, "ew_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "ew_check.cpp"), "-o", exe])
    for fs in (44100, 48000):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"klang_ref_fs{fs}.npz"))
        for name, (graph, total, block, events, burst) in cases.FX_SCRIPTS_LATE.items():
            if graph in (cases.FX_ECHO, cases.FX_FEEDBACK, cases.FX_IIR, cases.FX_WAHWAH, cases.FX_FLANGER, cases.FX_MODDELAY, cases.FX_MOD_CHORUS):
                continue                                  # test_one_delay_effects_match_reference_golden_on_host
            channels = 2 if graph == cases.FX_PAN else 1
            x = cases.fx_input(channels, total, 1, burst)
            path = os.path.join(tmp, "in.f32")
            np.ascontiguousarray(x, np.float32).tofile(path)
            args = [exe, str(graph), str(fs), str(channels), str(total), str(block), path]
            for (bi, c, v) in events:
                args += [str(bi), str(c), repr(float(v))]
            out = subprocess.run(args, capture_output=True)
            assert out.returncode == 0, name
            got = np.frombuffer(out.stdout, np.float32).reshape(channels, total)
            want = g[f"fx/{name}"].reshape(channels, total)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (fs, name)


def test_one_delay_effects_match_reference_golden_on_host():
    """Delay/Echo.k and Delay/Feedback.k: the product's frame functions (kb_echo_frame / kb_feedback_frame) compiled with g++ reproduce
    the compiled reference's golden vectors bit for bit (moving and fractional delay times, zero delay)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    tmp = tempfile.mkdtemp(prefix="kb_host_")
    exe = os.path.join(tmp, "onedelay_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "onedelay_check.cpp"), "-o", exe])
    checked = 0
    for fs in (44100, 48000):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"klang_ref_fs{fs}.npz"))
        for name, (graph, total, block, events, burst) in cases.FX_SCRIPTS_LATE.items():
            if graph not in (cases.FX_ECHO, cases.FX_FEEDBACK, cases.FX_IIR, cases.FX_WAHWAH, cases.FX_FLANGER, cases.FX_MODDELAY, cases.FX_MOD_CHORUS):
                continue
            path = os.path.join(tmp, "in.f32")
            np.ascontiguousarray(cases.fx_input(1, total, 1, burst)[0], np.float32).tofile(path)
            args = [exe, str(graph), str(fs), str(total), str(block), path]
            for (bi, c, v) in events:
                args += [str(bi), str(c), repr(float(v))]
            out = subprocess.run(args, capture_output=True)
            assert out.returncode == 0, name
            got = np.frombuffer(out.stdout, np.float32)
            assert np.array_equal(got.view(np.uint32), g[f"fx/{name}"].view(np.uint32)), (fs, name)
            checked += 1
    assert checked == 16                                  # echo, feedback, feedback_zero_delay, iir, wahwah, flanger, moddelay, mod_chorus at both rates


def test_additive_voices_match_oracle_and_time_parallel_form_on_host():
    """Additive/Saw.k / Square.k: per-tick form == time-parallel form (samples and state, checked inside the program) == the oracle port,
    bit for bit, across a re-trigger that keeps the partial phases and a pitch whose upper partials cross Nyquist."""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    import oracle
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "add_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "add_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True)
    assert out.returncode == 0, "per-tick and time-parallel forms differ"
    got = np.frombuffer(out.stdout, np.float32)

    def port(graph, fs):
        oracle.port.set_fs(fs)
        oracle.port.srand(1)
        sy = oracle.port.Synth(graph, 32)
        sy.voice_start(0, 57, 0.8)
        a = sy.process_voices(830)[0][0, 0]
        sy.voice_start(0, 88, 0.8)
        b = sy.process_voices(2170)[0][0, 0]
        sy.close()
        return np.concatenate([a, b])

    want = np.concatenate([port(oracle.SY_ADDITIVE_SAW, 48000.0), port(oracle.SY_ADDITIVE_SQUARE, 44100.0), port(oracle.SY_ADDITIVE_NYQUIST, 44100.0)])
    assert len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(want).max() > 0.5


def test_sine_modulation_voices_match_oracle_on_host():
    """Modulation/AM.k, FM.k, FM2.k: the product's voice functions (kb_smod_on / kb_smod_tick) compiled with g++ equal the oracle port bit
    for bit — frequencies set every sample (negative ones included), release, and a second note on the same voice."""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    import oracle
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "smod_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "smod_check.cpp"), "-o", exe])
    got = np.frombuffer(subprocess.run([exe], capture_output=True).stdout, np.float32)

    def port(graph, fs, ctl):
        oracle.port.set_fs(fs)
        oracle.port.srand(1)
        sy = oracle.port.Synth(graph, 32)
        for i, v in enumerate(ctl):
            sy.set_control(i, v)
        sy.voice_start(0, 60, 0.8)
        a = sy.process_voices(1500)[0][0, 0]
        sy.voice_release(0, 0.0)
        b = sy.process_voices(600)[0][0, 0]
        sy.voice_start(0, 67, 0.8)
        c = sy.process_voices(1900)[0][0, 0]
        sy.close()
        return np.concatenate([a, b, c])

    want = np.concatenate([port(oracle.SY_AM, 48000.0, (2.2, 0.9)), port(oracle.SY_MOD_FM, 44100.0, (1.5, 7.0)),
                           port(oracle.SY_MOD_FM2, 48000.0, (3.0, 10.0, 6.791))])
    assert len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(want[:4000]).max() > 0.3 and np.abs(want[4000:8000]).max() > 0.3 and np.abs(want[8000:]).max() > 0.01


def test_one_envelope_time_parallel_form_equals_per_tick_form():
    """kb_es_begin / kb_es_at / kb_es_end with kb_envr_run rows (what kb_esine_tiled_kernel runs for Breakpoint.k / Ramp.k / Release.k /
    Modulation/AM.k) equal kb_senv_tick / kb_smod_tick bit for bit, samples and state, over ragged blocks, releases, notes running into
    Off, re-triggers and control changes."""
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "esine_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "esine_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches, 0 state mismatches" in out.stdout


def test_echo_time_parallel_form_equals_frame_sequential_form():
    """Echo.k: write sweep + read sweep + position advance (kb_echo_write_at / kb_echo_read_at, what the time-parallel kernels run) equal
    the frame-sequential kb_echo_frame bit for bit — samples, ring contents, position — with the ring wrapping, moving / fractional
    delays, and the sequential fallback for delays below one frame."""
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "echo_par_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "echo_par_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 block mismatches, 0 ring / position mismatches" in out.stdout


def test_feedback_chunk_parallel_form_equals_frame_sequential_form():
    """Feedback.k: chunks shorter than the delay, frames of a chunk in any order (kb_feedback_chunk / kb_feedback_at, what
    kb_feedback_par_kernel runs) equal the frame-sequential kb_feedback_frame bit for bit — samples, ring contents, position — down to
    the 3-frame delay limit, with the frame-by-frame path below it."""
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "feedback_par_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "feedback_par_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 block mismatches, 0 ring / position mismatches" in out.stdout


def test_late_gpu_tests_execute_against_the_oracle_backed_stub():
    """The GPU tests written after the round's last GPU run (tests/test_zz_gpu_primitives.py, the MIDI test) are executed here with the
    klang_b200 names replaced by oracle-backed stand-ins (tests/dryrun_stub.py): every test body runs — shapes, keys, argument order,
    golden lookups.  Says nothing about the CUDA path; it keeps typos out of tests whose first real run is on the GPU box."""
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_stub.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    assert " passed" in out.stdout and "failed" not in out.stdout


def test_modulated_line_time_parallel_form_equals_frame_sequential_form():
    """Flanger.k / Modulation/Chorus.k: begin + write sweep with stash + read sweep (closed-form LFOs, stash-aware taps) + end — what the
    kb_modline_* kernels run — equal the frame-sequential kb_moddelay_frame bit for bit (samples, ring, state), zero-delay frames and
    negative-zero inputs included."""
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "modline_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "modline_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 block mismatches, 0 state / ring mismatches" in out.stdout


def test_reverb_tolerance_scan_stays_inside_the_parity_bar_on_host(tmp_path):
    """tests/host/reverb_scan_check.cpp: the arithmetic of the KB_FX_TOLERANCE filter scan (kb_scan.cuh), run lane by lane with g++, keeps
    Reverb.k inside 1e-5 |r| + 1e-6 peak for every configuration the plan admits, and the chunked read-ahead evaluation with the sequential
    filter is bit-identical to the frame-by-frame one.  The frame-by-frame output it dumps is compared with the compiled reference (or the
    port) fed the same input, bit for bit: the yardstick of the check is the reference's own result."""
    import numpy as np
    import oracle
    exe = str(tmp_path / "reverb_scan_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "reverb_scan_check.cpp"), "-o", exe])
    dump = str(tmp_path / "a.bin")
    out = subprocess.run([exe, dump], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "DIFFERENT" not in out.stdout and out.stdout.strip().endswith("ok")
    d = np.fromfile(dump, np.float32).reshape(8, 4, 4096)             # per block: in l, in r, out l, out r
    chk = oracle.ref if oracle.ref.available() else oracle.port
    chk.set_fs(48000)
    fx = chk.Fx(oracle.FX_REVERB)
    for c, v in enumerate((0.3, 0.9, 0.4, 0.5, 0.8)):
        fx.set_control(c, v)
    for b in range(8):
        want = fx.process(np.ascontiguousarray(d[b, :2]))
        assert np.array_equal(want.view(np.uint32), d[b, 2:].view(np.uint32)), f"block {b}: the host check's exact path differs from the reference"
    fx.close()
    chk.set_fs(44100)
