#!/usr/bin/env python3
"""klang-b200 benchmark (driver contract: python bench.py --gpus N --steps K --warmup W [--impl reference]).

Headline workload (BASELINE.json configs[1]): Subtractive synth (Saw >> LPF(env) >> ADSR), 1024 voices = 8 Synth
instances x 128 voices per GPU, fs = 48 kHz, 4096-sample blocks.  A step = one Synth::process pass over one block for
every voice, preceded by that block's note events (1/16 of the voices are re-triggered each block so envelopes and
filter coefficients keep moving — SURVEY §8d "steady-state variant").  metric = voice-samples/s.

  value      events + process with the output left in HBM (KB_DEVICE_PTR), CUDA-event timed per step, L2 flushed
             between steps (flush outside the per-step events).
  e2e        the same through the public host-buffer call: events, state upload, kernels, output D2H into pinned memory.
  roofline   dominant kernel (the voice kernel) bracketed by CUDA events inside the library (kb_*_bank_profile).
  cpu_baseline / --impl reference: the compiled reference klang.h (oracle/_ref) — or the plain-C port when the reference
             library is absent — on all host cores as fork()ed workers (the reference is not thread-safe).
N > 1: one rank per GPU, every rank renders its own 1024 voices (weak scaling) and the per-rank bank mixes
[channels][n] are reduced to rank 0 over NCCL each step (the polyphonic mix-down is the path's only exchange).
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000.0
BLOCK = 4096
INSTANCES, VOICES = 8, 128           # 1024 voices per GPU
RETRIGGER_GROUPS = 16


def voice_pitch(v):
    return 36 + (7 * v) % 61


def voice_velocity(v):
    return 0.25 + 0.75 * (((v * 2654435761) % (1 << 32)) >> 16) / 65535.0


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  In-process NVML (nvidia_ml_py) polled by a thread every
    10 ms (sparse on purpose: NVML queries can serialise with kernel launches in the driver), plus one explicit sample() right
    after the last timed step has been queued, while the GPU is still working through the region.  No NVML call sits between two timed
    steps: with N > 1 a rank that is late to queue a step would be waited for by the others inside their timed intervals.
    Falls back to one `nvidia-smi --query-gpu` per sample() without NVML."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        import threading
        self.gpu_index = gpu_index
        self.sm, self.reasons, self.mx = [], set(), None
        self.nvml = self.handle = None
        self.lock = threading.Lock()
        self.running = False
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if visible:
                try:
                    idx = int(visible.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.thread = None

    def start(self):
        import threading
        self.running = True
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        return self

    def _poll(self):
        while self.running:
            self.sample()
            time.sleep(0.010)

    def sample(self):
        if self.nvml is not None:
            try:
                n = self.nvml
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                names = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))
                with self.lock:
                    self.sm.append(mhz)
                    for name, bit in names:
                        if mask & bit:
                            self.reasons.add(name)
                return
            except Exception:
                pass
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout
            for line in out.splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                with self.lock:
                    self.sm.append(float(f[1]))
                    self.mx = max(self.mx or 0.0, float(f[2]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if val.lower().startswith("active"):
                            self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        self.running = False
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        with self.lock:
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0}
            return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "how": "NVML polled every 10 ms during the timed region plus once after the last step was queued" if self.nvml is not None
                    else "nvidia-smi --query-gpu once after the last timed step was queued"}


# ------------------------------------------------------------------------------------- reference (CPU) arm
def _cpu_worker(args):
    """One fork()ed worker: its share of the voices as one reference Synth, W+K steps of the steady-state schedule."""
    kind, voice_ids, steps, warmup, block = args
    import oracle
    eng = oracle.ref if kind == "reference" else oracle.port
    eng.set_fs(FS)
    eng.srand(1)
    sy = eng.Synth(oracle.SY_SUBTRACTIVE, max(1, len(voice_ids)))
    for k, g in enumerate(voice_ids):
        sy.voice_start(k, voice_pitch(g), voice_velocity(g))
    t0 = 0.0
    for s in range(warmup + steps):
        if s == warmup:
            t0 = time.perf_counter()
        for k, g in enumerate(voice_ids):
            if g % RETRIGGER_GROUPS == s % RETRIGGER_GROUPS:
                sy.voice_start(k, voice_pitch(g), voice_velocity(g))
        sy.process(block)
    dt = time.perf_counter() - t0
    sy.close()
    return dt


def cpu_reference_run(steps, warmup, total_voices=INSTANCES * VOICES, block=BLOCK, cores=None, passes=5, one_thread=True):
    """Times the reference CPU implementation of the C2 step on all host cores: best of `passes` runs of the whole (warmup + steps) schedule
    (best of 5; wall = slowest fork()ed worker), plus a 1-thread figure on a 1/16 share of the voices.  Returns (value, wall, dict)."""
    import oracle
    kind = "reference" if oracle.ref.available() else "port"
    (oracle.ref if kind == "reference" else oracle.port).lib()
    cores = cores or len(os.sched_getaffinity(0))
    cores = max(1, min(cores, total_voices))
    shares = [list(range(w, total_voices, cores)) for w in range(cores)]
    ctx = mp.get_context("fork")
    walls = []
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, passes)):
            walls.append(max(pool.map(_cpu_worker, [(kind, sh, steps, warmup, block) for sh in shares])))
    wall = min(walls)
    value = total_voices * block * steps / wall
    info = {"value": value, "unit": "voice-samples/s", "cores": cores, "kind": kind,
            "sample": f"best of {len(walls)} passes of {steps} steps x {total_voices} voices x {block} samples of the C2 schedule, {cores} fork()ed workers "
                      f"({'oracle/_ref = compiled klang.h' if kind == 'reference' else 'oracle/klang_port.c'}, g++/gcc -O3/-O2 -ffp-contract=off), "
                      f"wall = slowest worker, best pass {wall:.3f} s (passes: {', '.join('%.3f' % w for w in walls)})",
            "cpu_model": _cpu_model()}
    if one_thread:
        share = list(range(0, total_voices, 16))                   # 64 voices on one thread, same schedule
        with ctx.Pool(1) as pool:
            t1 = min(pool.map(_cpu_worker, [(kind, share, max(2, steps // 4), 1, block)])[0] for _ in range(2))
        info["one_thread_value"] = len(share) * block * max(2, steps // 4) / t1
    return value, wall, info


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _cpu_extra_worker(args):
    kind, what, graph, units, n, blocks = args
    import numpy as np
    import oracle
    eng = oracle.ref if kind == "reference" else oracle.port
    eng.set_fs(FS)
    eng.srand(1)
    if what == "fx":
        objs = [eng.Fx(graph) for _ in range(units)]
        x = (np.random.default_rng(1).random((objs[0].channels, n), dtype=np.float32) - 0.5)
        x = x[0] if objs[0].channels == 1 else x
        for o in objs:
            o.process(x)
        t0 = time.perf_counter()
        for _ in range(blocks):
            for o in objs:
                o.process(x)
        return time.perf_counter() - t0
    sy = eng.Synth(graph, max(1, units))
    for k in range(units):
        sy.voice_start(k, voice_pitch(k), voice_velocity(k))
    sy.process(n)
    t0 = time.perf_counter()
    for _ in range(blocks):
        sy.process(n)
    return time.perf_counter() - t0


def cpu_extra(what, graph, units_per_core, n, blocks):
    """Reference CPU throughput (units x samples per second over all cores) for a secondary workload."""
    import oracle
    kind = "reference" if oracle.ref.available() else "port"
    cores = len(os.sched_getaffinity(0))
    with mp.get_context("fork").Pool(cores) as pool:
        times = pool.map(_cpu_extra_worker, [(kind, what, graph, units_per_core, n, blocks)] * cores)
    return {"value": cores * units_per_core * n * blocks / max(times), "cores": cores, "kind": kind,
            "sample": f"{cores} workers x {units_per_core} {'instances' if what == 'fx' else 'voices'} x {blocks} blocks of {n}"}


def run_reference_arm(args, rank, world, emit):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    value, wall, info = cpu_reference_run(steps, warmup)
    line = {
        "impl": "reference", "metric": "voice-samples/sec (48 kHz equiv) at 1024 voices", "value": value, "unit": "voice-samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": wall / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "realtime_voices_48k": value / 48000.0,
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def make_mixdown(sharding, dist, local_rank, max_floats):
    """N > 1: the mix-down of the per-rank bank mixes.  Default = the fused peer-memory step (kb_synth_bank_process_mixdown: the bank-mix
    kernel stores into rank 0's arena over NVLink and rank 0's next kernel sums the previous block in rank order — no extra launches, no
    collective library on the data path); KB_MIXDOWN=nccl selects one ncclReduce per block beside the next block's kernels."""
    if dist is None:
        return None, "none (one GPU)"
    if os.environ.get("KB_MIXDOWN", "peer") == "nccl":
        return None, "ncclReduce of the [channels][n] mix per block, overlapped with the next block's kernels (KB_MIXDOWN=nccl)"
    try:
        return sharding.PeerMixdown(local_rank, max_floats), ("NVLink peer memory, fused: every rank's bank-mix kernel stores its [channels][n] mix into rank 0's arena and raises a flag; "
                                                              "rank 0's next bank-mix kernel sums the previous block in rank order (kb_synth_bank_process_mixdown)")
    except Exception as e:
        return None, f"ncclReduce of the [channels][n] mix per block (peer mapping unavailable: {e})"


def workload_config():
    return {"workload": "C2 Subtractive synth (Saw>>LPF(env,Q=10)>>ADSR), 8 Synth instances x 128 voices = 1024 voices per GPU, "
                        "fs 48 kHz, block 4096, 1/16 of the voices re-triggered per block",
            "instances": INSTANCES, "voices_per_instance": VOICES, "block": BLOCK, "fs": FS,
            "l2": "flushed between timed steps (256 MiB write, outside the per-step events)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (C1/C3/C4/C5 numbers inside the JSON)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 = headline (Subtractive, 1024 voices/GPU; carries the C4 / C5 numbers inside `roofline`); c5 = only TB303.k + SynTHX.k")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: everything native libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference_arm(args, rank, world, emit)
        return

    import numpy as np
    import torch
    import klang_b200 as kb

    if not torch.cuda.is_available() or kb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — klang_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # a real (non-legacy) stream: the library launches on it, torch events time it, NCCL orders against it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    hbm_peak, peak_src, sm_max = load_peaks()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2(first=False):
        """a 256 MiB write between timed steps.  Before the FIRST timed step (the GPU is idle behind a join) the write is queued four times: while
        the GPU works through them the host queues the step, so the first event interval is device time like every later one, not the host's
        submission latency on an idle stream (measured: 0.15-0.19 ms for step 0 against 0.073 ms for steps 1..K-1 without this)"""
        for _ in range(4 if first else 1):
            flush_buf.fill_(1)

    ctx = {"args": args, "rank": rank, "world": world, "local_rank": local_rank, "dist": dist, "dev": dev, "stream": stream,
           "flush_l2": flush_l2, "barrier": barrier, "hbm_peak": hbm_peak, "peak_src": peak_src, "sm_max": sm_max}
    if args.workload == "c5":
        line = c5_measure(kb, torch, ctx, args.steps, max(3, args.warmup), full_line=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            emit(line)
        return

    # ---- C2 bank
    kb.lib().kb_srand(1)
    bank = kb.SynthBank(kb.SY_SUBTRACTIVE, INSTANCES, VOICES, FS, BLOCK, local_rank)
    bank.set_stream(stream.cuda_stream)
    from klang_b200 import sharding
    total = INSTANCES * VOICES
    lo, hi = sharding.shard_instances(INSTANCES * world, rank, world)   # weak scaling: 8 Synth instances per rank
    assert hi - lo == INSTANCES
    gid0 = lo * VOICES                                    # global voice ids of this rank
    for g in range(total):
        bank.voice_start(g % VOICES, voice_pitch(gid0 + g), voice_velocity(gid0 + g), g // VOICES)
    # the SAME work at every world size: all voices rendered, summed per instance in voice order (the Stereo::Note rule, klang.h:4731,
    # SURVEY Q6 / 8e) and over the bank's instances into one [1][n] mix; N > 1 adds only the exchange of that mix
    flags = kb.BANK_MIX | kb.MIX_SUM
    out_dev = torch.empty(bank.out_shape(BLOCK, flags), dtype=torch.float32, device=dev)
    out_host = torch.empty(bank.out_shape(BLOCK, flags), dtype=torch.float32).pin_memory()
    step_counter = [0]

    # the 16 event batches of the steady-state schedule (one kb_note_event array per retrigger group)
    batches = []
    for grp in range(RETRIGGER_GROUPS):
        ids = list(range(grp, total, RETRIGGER_GROUPS))
        ev = np.zeros(len(ids), kb.EVENT_DTYPE)
        ev["type"] = kb.EV_VOICE_START
        ev["instance"] = [g // VOICES for g in ids]
        ev["key"] = [g % VOICES for g in ids]
        ev["pitch"] = [voice_pitch(gid0 + g) for g in ids]
        ev["velocity"] = [voice_velocity(gid0 + g) for g in ids]
        batches.append(ev)

    def next_events():
        s = step_counter[0]
        step_counter[0] += 1
        return batches[s % RETRIGGER_GROUPS]

    mixdown, mix_kind = make_mixdown(sharding, dist, local_rank, out_dev.numel())
    pipe = {"k": 0, "work": None, "last": out_dev}
    mix_bufs = [out_dev, torch.empty_like(out_dev)]

    def step_device():
        """the block's note events, one process() of this rank's bank and (N > 1) the cross-GPU mix-down towards out_dev on rank 0"""
        if dist is None:
            bank.step_into(next_events(), out_dev, BLOCK, flags)           # kb_synth_bank_step: events + process in one call
        elif mixdown is not None:
            bank.events(next_events())
            bank.process_mixdown(mixdown, out_dev if rank == 0 else None, BLOCK, kb.MIX_SUM)
        else:
            # the reduce of block k runs on NCCL's stream beside the voice kernels of block k+1 (two mix buffers); it is
            # joined after the next block's kernels have been queued, and the last one by drain() inside the timed region
            buf = mix_bufs[pipe["k"] & 1]
            pipe["k"] += 1
            bank.step_into(next_events(), buf, BLOCK, flags)
            if pipe["work"] is not None:
                pipe["work"].wait()
            pipe["work"] = dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM, async_op=True)
            pipe["last"] = buf

    def drain():
        """join the exchange still in flight (the timed loops call it before their closing event / barrier)"""
        if mixdown is not None:
            if rank == 0:
                mixdown.collect(out_dev, out_dev.numel(), stream.cuda_stream)
        elif pipe["work"] is not None:
            pipe["work"].wait()
            pipe["work"] = None

    host_bufs = [out_host, torch.empty_like(out_host).pin_memory()]
    host_done = [torch.cuda.Event(), torch.cuda.Event()]
    assert all(hb.is_pinned() and hb.is_contiguous() and hb.dtype == torch.float32 for hb in host_bufs)
    host_ptrs = [hb.data_ptr() for hb in host_bufs]
    copy_stream = torch.cuda.Stream(device=dev)
    e2e_k = [0]

    def step_e2e():
        if dist is None:
            # host-buffer call with KB_ASYNC_HOST: state upload, kernels and the D2H copy into pinned memory are queued; the host
            # prepares the next block's events meanwhile and joins a buffer only before reusing it (two buffers, as a streaming
            # host would).  Every block's result still crosses PCIe inside the timed region.
            i = e2e_k[0] & 1
            if e2e_k[0] >= 2:
                host_done[i].synchronize()
            bank.step_into(next_events(), host_bufs[i].numpy(), BLOCK, flags | kb.ASYNC_HOST)
            host_done[i].record(stream)
            e2e_k[0] += 1
        elif mixdown is not None:
            # fused peer mix-down: the exchange kernel of block k delivers the rank-order sum of block k-1 straight into pinned host memory on rank 0
            # (device-mapped: 16 KiB of stores over PCIe inside the kernel; two host buffers, each joined through an event behind that kernel
            # before reuse, the last block by drain_e2e() inside the timed region).  Every block's reduced mix crosses PCIe.
            k = e2e_k[0]
            i = k & 1
            if rank == 0 and k >= 2:
                mixdown.host_wait(1)                                   # the exchange that wrote host_bufs[i] two blocks ago (one before the last) has finished
            # one call per block: the events, the kernels, the exchange; pinned out_prev: the exchange kernel stores the sum over PCIe
            bank.step_mixdown(next_events(), mixdown, host_ptrs[i] if rank == 0 else 0, BLOCK, kb.MIX_SUM)
            e2e_k[0] += 1
        else:
            step_device()
            drain()                                              # NCCL transport: every block's reduced mix is joined and read back
            if rank == 0:
                out_host.copy_(pipe["last"], non_blocking=True)
            torch.cuda.synchronize()

    def drain_e2e():
        """the last block's sum (fused transport) and every copy still in flight"""
        if dist is not None and mixdown is not None:
            drain()
            if rank == 0:
                out_host.copy_(out_dev, non_blocking=True)
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, clocks=None):
        for _ in range(warmup):
            step_fn()
        drain()
        barrier()
        if clocks:
            clocks.start()
        evs = []
        launches0 = bank.launches
        t_wall0 = time.perf_counter()
        for i in range(steps):
            flush_l2(first=(i == 0))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step_fn()
            b.record(stream)
            evs.append((a, b))
        if dist is not None:                          # the last block's exchange, still in flight, belongs to the timed region
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            drain()
            b.record(stream)
            evs.append((a, b))
        launches = bank.launches - launches0 + (1 if (mixdown is not None and rank == 0) else 0)
        if clocks:
            clocks.sample()                          # everything is queued, the GPU is still inside the timed region
        barrier()
        wall = time.perf_counter() - t_wall0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if os.environ.get("KB_BENCH_DEBUG"):
            print("per-step ms:", " ".join(f"{a.elapsed_time(b):.4f}" for a, b in evs), file=sys.stderr)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, launches

    clocks = ClockSampler(local_rank) if rank == 0 else None
    ms_total, _, launches = timed(step_device, args.steps, args.warmup, clocks)
    gpu_launches = launches / args.steps                 # this library's kernels per step (the L2 flush fill between steps is torch's)
    clk = clocks.stop() if clocks else None
    ms_per_step = ms_total / args.steps
    value = world * total * BLOCK / (ms_per_step * 1e-3)

    # ---- e2e through host buffers (host wall clock is part of it: events run on the host)
    drain()
    for _ in range(2):
        step_e2e()
    drain_e2e()
    e2e_k[0] = 0
    barrier()
    h2d0, d2h0 = bank.transfer_bytes()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    drain_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * total * BLOCK * args.steps / float(t.item())
    h2d1, d2h1 = bank.transfer_bytes()
    out_bytes = out_host.numel() * 4 if dist is not None else 0      # (N=1: the library's host-buffer call counts its own D2H)
    e2e = {"value": e2e_value, "unit": "voice-samples/s",
           "h2d_bytes_per_step": int((h2d1 - h2d0) / args.steps), "d2h_bytes_per_step": int((d2h1 - d2h0) / args.steps + out_bytes),
           "note": "per step: the host applies 64 note events on its state mirror (packed dirty-voice upload H2D), kernels, output D2H "
                   "into pinned memory (N=1: KB_ASYNC_HOST, two host buffers, block k+1's events are prepared while block k renders; "
                   "every buffer is joined before reuse and at the end of the timed region; N>1, fused peer mix-down: the same with the reduced mix of block k-1 stored into pinned host memory by the exchange kernel while block k renders; N>1 over NCCL: joined and read back every block); "
                   "bytes counted by the library"}

    # ---- dominant kernel, CUDA events inside the library
    bank.profile(True)
    for _ in range(args.steps):
        flush_l2()
        step_device()
    drain()
    k_ms, k_n = bank.profile_read()
    bank.profile(False)
    # A/B: the plain lane-per-voice schedule of the same arithmetic (same results, see tests)
    lane_ms = None
    if dist is None:
        bank.profile(True)
        for _ in range(3):
            bank.process_into(out_dev, BLOCK, flags | kb.LANE_PER_VOICE)
        l_ms, l_n = bank.profile_read()
        bank.profile(False)
        lane_ms = l_ms / max(1, l_n)
    k_ms_avg = k_ms / max(1, k_n)
    alg_bytes = total * BLOCK * 4.0      # 4 B per voice-sample (the per-voice stream this kernel writes)
    hbm_achieved = alg_bytes / (k_ms_avg * 1e-3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = (lambda t: t.get("kb_sub_mbar_kernel", t.get("kb_sub_flow_kernel", t.get("kb_sub_tiled_kernel"))))(json.load(f))
    except Exception:
        pass
    # What bounds this kernel: one fp32 recurrence per voice (the TDF-II biquad, klang.h:5605-5612) that parity forbids re-ordering:
    # 4 dependent fp32 operations per sample.  Its floor was measured alone on one warp of this chip (tools/micro/serial_floor.cu,
    # profiles/r01_probes.txt): 16.9 cycles per sample.  Every voice's chain advances one sample per tick of its CTA and all CTAs tick in
    # parallel, so the floor of a block is BLOCK x 16.9 cycles whatever the voice count (up to one wave of CTAs).
    floor_cycles = 16.9
    kernel_vs = total * BLOCK / (k_ms_avg * 1e-3)
    floor_vs = total * sm_max * 1e6 / floor_cycles
    roofline = {"bound": "latency", "kernel": "kb_sub_mbar_kernel", "achieved": kernel_vs, "peak": floor_vs, "unit": "voice-samples/s",
                "frac": kernel_vs / floor_vs, "traffic": traffic,
                "peak_source": f"serial-chain floor: {floor_cycles} cycles per sample of the TDF-II biquad recurrence (measured alone, profiles/r01_probes.txt) at the {sm_max:.0f} MHz of MEASURED_PEAKS.json, x {total} voices in flight",
                "latency_frac": kernel_vs / floor_vs,
                "achieved_cycles_per_sample": k_ms_avg * 1e-3 * sm_max * 1e6 / BLOCK, "floor_cycles_per_sample": floor_cycles,
                "kernel_ms": k_ms_avg, "kernel_share_of_step": k_ms_avg / ms_per_step,
                "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg_bytes,
                        "note": "4 B per voice-sample = the per-voice stream the kernel writes (SURVEY 8d, parity-dump figure); the streams stay in L2 "
                                "(traffic = DRAM bytes of one launch from the ncu capture), so the HBM fraction of C2 is a few % by nature (SURVEY H6)"},
                "lane_per_voice_kernel_ms": lane_ms}
    bank.close()

    line = {
        "metric": "voice-samples/sec (48 kHz equiv) at 1024 voices", "value": value, "unit": "voice-samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(), "mixdown": mix_kind, "realtime_voices_48k": value / 48000.0,
        "clocks": clk, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline,
    }
    if mixdown is not None:
        torch.cuda.synchronize()
        dist.barrier()
        mixdown.close()

    # ---- BASELINE configs C4 (one GPU: effect instances are replicas) and C5 (all ranks), inside keys the driver keeps
    if not args.no_extras:
        try:
            c5 = c5_measure(kb, torch, ctx, steps=5, warmup=3, full_line=False)
            if rank == 0:
                roofline["c5"] = c5
        except Exception as e:
            roofline["c5"] = {"error": repr(e)}
    if rank == 0 and not args.no_extras:
        try:
            other = extras(kb, torch, dev, stream, flush_l2, hbm_peak, peak_src, local_rank, with_cpu=(world == 1 and not args.no_cpu))
            line["other_workloads"] = other
            for key, name in (("c4_reverb_64", "c4_reverb"), ("c4_reverb_64_tolerance", "c4_reverb_tolerance"), ("c4_pingpong_64", "c4_pingpong"),
                              ("c4_delay_pingpong_64", "c4_delay_pingpong"), ("c4_delay_pingpong_64_x65536", "c4_delay_pingpong_65536"), ("c4_delay_reverb_64", "c4_delay_reverb")):
                w = other.get(key)
                if isinstance(w, dict) and "roofline" in w:
                    roofline[name + "_frac"] = w["roofline"]["frac"]
                    roofline[name + "_ms"] = w["ms_per_step"]
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    tj = json.load(f)
                roofline["c4_reverb_traffic"] = tj.get("kb_reverb3_kernel<0>")
                roofline["c4_reverb_tolerance_traffic"] = tj.get("kb_reverb3_kernel<1>")
                roofline["c4_pingpong_traffic"] = tj.get("kb_pingpong3_kernel")
                roofline["c4_reverb_algorithmic_bytes"] = 64 * 4096 * 664
            except Exception:
                pass
            roofline["c4_note"] = ("C4 = 64 stereo instances x 4096-frame blocks, distinct input per step, L2 flushed before every timed step; frac = algorithmic bytes per frame "
                                   "(SURVEY 8d: Reverb.k 664, PingPong.k 48, Delay/PingPong.k 40, Delay/Reverb.k 88) x frames / time / measured copy peak; "
                                   "*_tolerance = the 1e-5-tolerance schedule (KB_FX_TOLERANCE), everything else bit-exact")
        except Exception as e:   # secondary numbers never take the headline down
            line["other_workloads"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            _, _, info = cpu_reference_run(steps=100, warmup=2, passes=3)      # ~0.5 s wall per pass on every host core: 10-30 s of CPU work
            line["cpu_baseline"] = info
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def c5_measure(kb, torch, ctx, steps, warmup, full_line):
    """BASELINE config C5: TB303.k + SynTHX.k mixed bank, 8192 voices over 8 GPUs = per GPU 4 x 128 TB303 voices (mono)
    + 4 x 128 SynTHX voices (stereo), block 4096; the two banks render on two streams, every rank mixes its instances into one stereo
    bus (KB_BANK_MIX, mono voices to both channels) and the buses are summed on rank 0."""
    from klang_b200 import sharding
    rank, world, local_rank, dist, dev, stream = ctx["rank"], ctx["world"], ctx["local_rank"], ctx["dist"], ctx["dev"], ctx["stream"]
    hbm_peak, peak_src = ctx["hbm_peak"], ctx["peak_src"]
    inst, voices, n = 4, 128, BLOCK
    kb.lib().kb_srand(1 + rank)
    tb = kb.SynthBank(kb.SY_TB303, inst, voices, FS, n, local_rank)
    sx = kb.SynthBank(kb.SY_SYNTHX, inst, voices, FS, n, local_rank)
    side = torch.cuda.Stream(device=dev)                   # TB303 renders beside SynTHX: neither fills the chip alone
    tb.set_stream(side.cuda_stream)
    sx.set_stream(stream.cuda_stream)
    gid0 = rank * 2 * inst * voices
    for g in range(inst * voices):
        tb.voice_start(g % voices, 36 + (5 * (gid0 + g)) % 36, voice_velocity(gid0 + g), g // voices)
        sx.voice_start(g % voices, 36 + (7 * (gid0 + g)) % 30, voice_velocity(gid0 + g), g // voices)
    tb_mix = torch.empty(1, n, dtype=torch.float32, device=dev)
    bus = torch.empty(2, n, dtype=torch.float32, device=dev)
    bus_out = torch.empty(2, n, dtype=torch.float32, device=dev)
    bus_host = torch.empty(2, n, dtype=torch.float32).pin_memory()
    mixdown, mix_kind = make_mixdown(sharding, dist, local_rank, 2 * n)
    fork, join = torch.cuda.Event(), torch.cuda.Event()

    def drain():
        if mixdown is not None and rank == 0:
            mixdown.collect(bus_out, 2 * n, stream.cuda_stream)

    def step(e2e=False):
        fork.record(stream)
        side.wait_event(fork)
        tb.process_into(tb_mix, n, kb.BANK_MIX | kb.MIX_SUM)       # side stream
        join.record(side)
        sx.process_into(bus, n, kb.BANK_MIX)                       # main stream
        stream.wait_event(join)
        bus.add_(tb_mix)                                   # mono voices feed both channels
        if mixdown is not None:
            mixdown.step(bus, 2 * n, bus_out if rank == 0 else None, stream.cuda_stream)   # fused: store into rank 0's arena + flag; rank 0 sums the previous block
        elif dist is not None:
            sharding.reduce_mix(bus, dst=0)
        if e2e:
            drain()
            if rank == 0:
                bus_host.copy_(bus_out if mixdown is not None else bus, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    drain()
    ctx["barrier"]()
    clocks = ClockSampler(local_rank).start() if (rank == 0 and full_line) else None
    evs = []
    for i in range(steps):
        ctx["flush_l2"](first=(i == 0))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        evs.append((a, b))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    drain()
    b.record(stream)
    evs.append((a, b))
    if clocks:
        clocks.sample()
    ctx["barrier"]()
    clk = clocks.stop() if clocks else None
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    total = 2 * inst * voices
    value = world * total * n / (ms_per_step * 1e-3)
    ctx["barrier"]()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(e2e=True)
    ctx["barrier"]()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * total * n * steps / float(te.item())
    sx.profile(True)
    tb.profile(True)
    l0 = tb.launches + sx.launches
    for _ in range(5):
        step()
    drain()
    sx_ms, sx_n = sx.profile_read()
    tb_ms, tb_n = tb.profile_read()
    launches = (tb.launches + sx.launches - l0) / 5 + (1 if mixdown is not None else 0)
    k_ms = sx_ms / max(1, sx_n)
    tb.close()
    sx.close()
    if mixdown is not None:
        torch.cuda.synchronize()
        dist.barrier()
        mixdown.close()
    workload = ("C5 TB303.k + SynTHX.k mixed bank: per GPU 4 x 128 TB303 voices + 4 x 128 SynTHX voices, fs 48 kHz, block 4096, the two banks on two streams, "
                "stereo bus summed on rank 0")
    if not full_line:
        return {"workload": workload, "value": value, "unit": "voice-samples/s", "n_gpus": world, "voices_total": world * total, "ms_per_step": ms_per_step,
                "e2e_value": e2e_value, "steps": steps, "synthx_kernel_ms": k_ms, "tb303_kernel_ms": tb_ms / max(1, tb_n), "mixdown": mix_kind, "gpu_launches": launches}
    alg = inst * voices * n * 8.0                         # SynTHX per-voice streams are never materialised: 8 B per voice-sample of ADSR + bus traffic
    return {
        "metric": "voice-samples/sec (48 kHz equiv) at 1024 voices", "value": value, "unit": "voice-samples/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "block": n, "fs": FS, "l2": "flushed between timed steps (256 MiB write, outside the per-step events)"},
        "mixdown": mix_kind, "realtime_voices_48k": value / 48000.0, "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 2 * n * 4,
                "note": "no events in this schedule (all voices held); per step the reduced stereo bus is copied to pinned host memory"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "kb_sx_render_kernel", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": alg / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src, "kernel_ms": k_ms,
                     "kernel_share_of_step": k_ms / ms_per_step, "tb303_kernel_ms": tb_ms / max(1, tb_n),
                     "note": "bound by the ordered fp32 accumulation chain per output sample (DESIGN.md 4.2), not HBM"},
    }


def extras(kb, torch, dev, stream, flush_l2, hbm_peak, peak_src, device_index, with_cpu=True):
    """Secondary BASELINE configs, short runs: C1 Gain, C3 SuperSaw, C4 PingPong / Reverb / Delay-PingPong / Delay-Reverb, streaming effects."""
    import oracle
    res = {}

    def time_steps(fn, steps, warmup=3, pre=None):
        """CUDA-event time of fn() per step; `pre` (untimed) runs before the L2 flush of every step, so what it writes is not L2-resident"""
        for _ in range(warmup):
            if pre:
                pre()
            fn()
        torch.cuda.synchronize()
        evs = []
        for i in range(steps):
            if pre:
                pre()
            flush_l2(first=(i == 0))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            evs.append((a, b))
        # (no host join between steps: while the GPU writes the 256 MiB flush the host has already queued the step behind it, so the event
        # interval is the device time of the call's kernels, not the host's launch latency on an idle stream)
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / steps

    def roof(gbs):
        return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": peak_src}

    # C1: Gain, 1 channel x 4096 (launch-bound) and batched 64 Mi samples (HBM-bound)
    def streaming_line(name, graph, inst, n):
        fx = kb.FxBank(graph, inst, FS, n, device_index)
        fx.set_stream(stream.cuda_stream)
        io = torch.rand(inst, fx.channels, n, device=dev) - 0.5
        ms = time_steps(lambda: fx.process_inplace(io), 10)
        res[name] = {"samples_per_s": inst * n / (ms * 1e-3), "ms_per_step": ms, "roofline": roof(inst * fx.channels * n * 8 / (ms * 1e-3) / 1e9)}
        fx.close()

    streaming_line("c1_gain_1x4096", kb.FX_GAIN, 1, 4096)
    streaming_line("c1_gain_batched_64x1Mi", kb.FX_GAIN, 64, 1 << 20)
    streaming_line("c1_gain_batched_256x1Mi", kb.FX_GAIN, 256, 1 << 20)      # 2 GiB of traffic per call: the size class the copy peak was measured on
    if with_cpu:
        res["c1_gain_1x4096"]["cpu_reference"] = cpu_extra("fx", oracle.FX_GAIN, 1, 4096, 2000)

    # C4: delay-line effects, 64 stereo instances x 4096-frame blocks (the BASELINE shape); every timed step gets fresh input copied from a
    # master buffer BEFORE the L2 flush (effects work in place), so neither io nor rings are L2-warm from the previous step's input
    def c4_line(name, graph, n, steps, flags=0, settle=True):
        fx = kb.FxBank(graph, 64, FS, n, device_index)
        fx.set_stream(stream.cuda_stream)
        master = torch.rand(4, 64, fx.channels, n, device=dev) - 0.5
        io = torch.empty(64, fx.channels, n, device=dev)
        k = [0]

        def refill():
            io.copy_(master[k[0] & 3])
            k[0] += 1

        if settle:
            for _ in range(40000 // n + 2):                     # PingPong.k: let the control smoothers reach their fixed point
                refill()
                fx.process_inplace(io, flags=flags)
        ms = time_steps(lambda: fx.process_inplace(io, flags=flags), steps, warmup=2, pre=refill)
        par = fx.parallel_instances()
        bpf = fx.bytes_per_frame()
        res[name] = {"frames_per_s": 64 * n / (ms * 1e-3), "ms_per_step": ms, "block": n, "bytes_per_frame": bpf,
                     "instances_on_parallel_schedule": par, "roofline": roof(64 * n * bpf / (ms * 1e-3) / 1e9)}
        if n <= 4096 and flags == 0:
            res[name]["ms_per_step_sequential_schedule"] = time_steps(lambda: fx.process_inplace(io, flags=kb.FX_SEQUENTIAL), 1, warmup=0, pre=refill)
        fx.close()

    c4_line("c4_reverb_64", kb.FX_REVERB, 4096, 5)
    if hasattr(kb, "FX_TOLERANCE"):
        c4_line("c4_reverb_64_tolerance", kb.FX_REVERB, 4096, 5, flags=kb.FX_TOLERANCE)
    c4_line("c4_pingpong_64", kb.FX_PINGPONG, 4096, 5)
    c4_line("c4_delay_pingpong_64", kb.FX_DELAY_PINGPONG, 4096, 5)
    c4_line("c4_delay_pingpong_64_x65536", kb.FX_DELAY_PINGPONG, 65536, 5)
    c4_line("c4_delay_reverb_64", kb.FX_DELAY_REVERB, 4096, 5)
    if with_cpu:
        res["c4_pingpong_64"]["cpu_reference"] = cpu_extra("fx", oracle.FX_PINGPONG, 1, 4096, 40)
        res["c4_reverb_64"]["cpu_reference"] = cpu_extra("fx", oracle.FX_REVERB, 1, 4096, 8)
        res["c4_delay_pingpong_64"]["cpu_reference"] = cpu_extra("fx", oracle.FX_DELAY_PINGPONG, 1, 4096, 40)
        res["c4_delay_reverb_64"]["cpu_reference"] = cpu_extra("fx", oracle.FX_DELAY_REVERB, 1, 4096, 40)

    # C3 SuperSaw 8 x 32 voices; C5 per-GPU share: 4 x 128 TB303 + 4 x 128 SynTHX voices
    for name, graph, inst, voices, n, steps in (("c3_supersaw_256", kb.SY_SUPERSAW, 8, 32, 4096, 5),
                                                ("c5_tb303_512", kb.SY_TB303, 4, 128, 4096, 3),
                                                ("c5_synthx_512", kb.SY_SYNTHX, 4, 128, 1024, 2),
                                                ("fm_k_1024", kb.SY_FM, 8, 128, 4096, 3)):      # examples/FM.k (time-parallel kernel)
        kb.lib().kb_srand(1)
        sb = kb.SynthBank(graph, inst, voices, FS, n, device_index)
        sb.set_stream(stream.cuda_stream)
        for g in range(inst * voices):
            sb.voice_start(g % voices, 36 + (5 * g) % 36, voice_velocity(g), g // voices)
        out = torch.empty(sb.out_shape(n), dtype=torch.float32, device=dev)
        ms = time_steps(lambda: sb.process_into(out, n), steps, warmup=3)
        res[name] = {"voice_samples_per_s": inst * voices * n / (ms * 1e-3), "ms_per_step": ms, "block": n}
        if graph == kb.SY_FM:                              # A/B against the plain lane-per-voice schedule (same results)
            res[name]["ms_per_step_lane_per_voice_schedule"] = time_steps(lambda: sb.process_into(out, n, kb.LANE_PER_VOICE), 1, warmup=1)
        sb.close()
    if with_cpu:
        res["c3_supersaw_256"]["cpu_reference"] = cpu_extra("synth", oracle.SY_SUPERSAW, 16, 4096, 8)
        res["c5_tb303_512"]["cpu_reference"] = cpu_extra("synth", oracle.SY_TB303, 16, 4096, 4)
        res["c5_synthx_512"]["cpu_reference"] = cpu_extra("synth", oracle.SY_SYNTHX, 4, 1024, 2)
        res["fm_k_1024"]["cpu_reference"] = cpu_extra("synth", oracle.SY_FM, 16, 4096, 8)
    # the other elementwise effects on the streaming schedule (Tremolo.k with its closed-form LFO, Pan.k stereo)
    for name, graph, inst in (("tremolo_k_batched_64x1Mi", kb.FX_TREMOLO, 64), ("pan_k_batched_32x1Mi", kb.FX_PAN, 32)):
        try:
            streaming_line(name, graph, inst, 1 << 20)
        except Exception as e:
            res[name] = {"error": repr(e)}
    # Delay/Echo.k and Modulation/Flanger.k, 64 instances x 65536 frames (20 / 24 algorithmic bytes per frame)
    for name, graph, bpf in (("echo_k_64", kb.FX_ECHO, 20), ("flanger_k_64", kb.FX_FLANGER, 24)):
        try:
            fx = kb.FxBank(graph, 64, FS, 65536, device_index)
            fx.set_stream(stream.cuda_stream)
            if graph == kb.FX_ECHO:
                fx.set_control(0, 0.25)
            master = torch.rand(64, 1, 65536, device=dev) - 0.5
            io = torch.empty_like(master)
            ms = time_steps(lambda: fx.process_inplace(io), 5, warmup=2, pre=lambda: io.copy_(master))
            res[name] = {"frames_per_s": 64 * 65536 / (ms * 1e-3), "ms_per_step": ms, "block": 65536, "bytes_per_frame": bpf,
                         "roofline": roof(64 * 65536 * bpf / (ms * 1e-3) / 1e9)}
            fx.close()
        except Exception as e:
            res[name] = {"error": repr(e)}
    # Additive/Saw.k (32 sine partials per voice, no recurrence at all), Subtractive/Release.k and Modulation/AM.k (one envelope x closed-form
    # sines): time-parallel kernels, each with the lane-per-voice A/B
    for name, graph in (("additive_saw_k_1024", kb.SY_ADDITIVE_SAW), ("release_k_1024", kb.SY_RELEASE), ("am_k_1024", kb.SY_AM)):
        try:
            sb = kb.SynthBank(graph, 8, 128, FS, 4096, device_index)
            sb.set_stream(stream.cuda_stream)
            for g in range(1024):
                sb.voice_start(g % 128, 36 + (5 * g) % 36, voice_velocity(g), g // 128)
            out = torch.empty(sb.out_shape(4096), dtype=torch.float32, device=dev)
            ms = time_steps(lambda: sb.process_into(out, 4096), 3, warmup=3)
            res[name] = {"voice_samples_per_s": 1024 * 4096 / (ms * 1e-3), "ms_per_step": ms, "block": 4096,
                         "ms_per_step_lane_per_voice_schedule": time_steps(lambda: sb.process_into(out, 4096, kb.LANE_PER_VOICE), 1, warmup=1)}
            sb.close()
        except Exception as e:
            res[name] = {"error": repr(e)}
    return res


if __name__ == "__main__":
    main()
