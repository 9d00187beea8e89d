"""Quick device probe of the streaming effect kernels (measurement aid, same method as bench.py's extras: L2 flushed before every timed call,
CUDA events around kb_fx_bank_process, algorithmic bytes / time against the measured copy peak).  Usage: python tools/stream_probe.py [inst_scale]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import klang_b200 as kb

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
CASES = [("gain", kb.FX_GAIN, 64, 1 << 20, 8), ("pan", kb.FX_PAN, 32, 1 << 20, 8), ("tremolo", kb.FX_TREMOLO, 64, 1 << 20, 8), ("clipping", kb.FX_CLIPPING, 64, 1 << 20, 8),
         ("echo", kb.FX_ECHO, 64, 65536, 20), ("flanger", kb.FX_FLANGER, 64, 65536, 24), ("chorus", kb.FX_MOD_CHORUS, 64, 65536, 40)]
with torch.cuda.stream(stream):
    for name, graph, inst, n, bpf in CASES:
        inst *= scale
        fx = kb.FxBank(graph, inst, 48000.0, n, 0)
        fx.set_stream(stream.cuda_stream)
        io = torch.rand(inst, fx.channels, n, device=dev) - 0.5
        for _ in range(3):
            fx.process_inplace(io)
        torch.cuda.synchronize()
        evs = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fx.process_inplace(io); b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        med = ms[len(ms) // 2]
        gbs = inst * fx.channels * n * bpf / (med * 1e-3) / 1e9
        print(f"{name:9s} {inst:4d} x {n:8d}: {med * 1e3:7.1f} us (min {ms[0] * 1e3:.1f})  {gbs:7.0f} GB/s  {gbs / peak:.3f} of the copy peak", flush=True)
        fx.close()
