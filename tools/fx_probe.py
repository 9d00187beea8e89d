"""Quick device probe of the effect kernels (not part of the bench): steady-state kernel time via the library's event profile."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import klang_b200 as kb

name = sys.argv[1] if len(sys.argv) > 1 else "reverb"
graph = {"pingpong": kb.FX_PINGPONG, "reverb": kb.FX_REVERB, "dpingpong": kb.FX_DELAY_PINGPONG, "gain": kb.FX_GAIN, "dreverb": kb.FX_DELAY_REVERB}[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4096
inst = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 64
flags = kb.FX_SEQUENTIAL if "--seq" in sys.argv else (kb.FX_TOLERANCE if "--tol" in sys.argv else 0)
fx = kb.FxBank(graph, inst, 48000.0, n)
ch = fx.channels
io = torch.rand(inst, ch, n, device="cuda") - 0.5
warm = max(3, 40000 // n + 2)          # let PingPong's control smoothers settle
for _ in range(warm):
    fx.process_inplace(io, flags=flags)
if "--allbus" in sys.argv and name == "reverb":
    for c, v in ((0, 0.3), (1, 0.9), (2, 0.4), (3, 0.5), (4, 0.8)):
        fx.set_control(c, v)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fx.profile(True)
steps = 5
for _ in range(steps):
    io.uniform_(-0.5, 0.5)
    flush.fill_(1)                       # L2 cold
    fx.process_inplace(io, flags=flags)
ms, cnt = fx.profile_read()
bpf = fx.bytes_per_frame()
t = ms / steps * 1e-3
print(f"{name} n={n} inst={inst} flags={flags} parallel_instances={fx.parallel_instances()} tolerance_instances={fx.tolerance_instances()} step {t * 1e6:.1f} us -> {inst * n / t:.3e} frames/s, {inst * n * bpf / t / 1e9:.1f} GB/s algorithmic ({bpf} B/frame)")
