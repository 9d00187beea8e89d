"""Quick device probe of the C2 voice kernel (not part of the bench): kernel milliseconds via the library's event profile."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import klang_b200 as kb

graph = {"sub": kb.SY_SUBTRACTIVE, "ssaw": kb.SY_SUPERSAW, "tb": kb.SY_TB303, "sx": kb.SY_SYNTHX, "fm": kb.SY_FM, "add": kb.SY_ADDITIVE_SAW, "senv": kb.SY_RELEASE}[sys.argv[1] if len(sys.argv) > 1 else "sub"]
inst, voices = (8, 128) if graph != kb.SY_SUPERSAW else (8, 32)
if len(sys.argv) > 3:
    inst, voices = int(sys.argv[2]), int(sys.argv[3])
N = 1024 if graph == kb.SY_SYNTHX else 4096
bank = kb.SynthBank(graph, inst, voices, 48000.0, N)
for g in range(inst * voices):
    bank.voice_start(g % voices, 36 + (5 * g) % 36, 0.8, g // voices)
out = torch.empty(bank.out_shape(N), dtype=torch.float32, device="cuda")
for _ in range(3):
    bank.process_into(out, N)
bank.profile(True)
for _ in range(10):
    bank.process_into(out, N)
ms, n = bank.profile_read()
print(f"graph {sys.argv[1] if len(sys.argv) > 1 else 'sub'} KB_TILE_G={os.environ.get('KB_TILE_G', 'auto')} voices {inst * voices}: kernel {ms / n * 1e3:.1f} us -> {inst * voices * N / (ms / n * 1e-3):.3e} voice-samples/s")
