"""A/B of the C2 voice-kernel layouts (measurement aid): each variant runs in its own process (KB_TILE_LAYOUT / KB_TILE_G / KB_TILE_ASP0 are read once),
renders the same 6 blocks of per-voice streams with re-triggers, and reports the kernel time; the parent compares the streams bit for bit.
Usage: python tools/c2_ab.py            (parent)      python tools/c2_ab.py child out.npy   (one variant)"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

VARIANTS = [("layout2 G8", dict(KB_TILE_LAYOUT="2")), ("flow G8", dict(KB_TILE_LAYOUT="3", KB_TILE_G="8")),
            ("flow G7", dict(KB_TILE_LAYOUT="3", KB_TILE_G="7")), ("flow G7 envr_run", dict(KB_TILE_LAYOUT="3", KB_TILE_G="7", KB_C2_VARIANT="32")),
            ("mbar G7", dict(KB_TILE_LAYOUT="4", KB_TILE_G="7")), ("mbar G8", dict(KB_TILE_LAYOUT="4", KB_TILE_G="8"))]

def child(path):
    import torch
    import klang_b200 as kb
    inst, voices, N = 8, 128, 4096
    bank = kb.SynthBank(kb.SY_SUBTRACTIVE, inst, voices, 48000.0, N)
    for g in range(inst * voices):
        bank.voice_start(g % voices, 36 + (5 * g) % 36, 0.8, g // voices)
    outs = []
    rng = np.random.default_rng(5)
    for blk in range(6):
        for g in rng.choice(inst * voices, 64, replace=False):
            if blk % 2: bank.voice_release(int(g) % voices, 0.0, int(g) // voices)
            else: bank.voice_start(int(g) % voices, 40 + int(g) % 30, 0.7, int(g) // voices)
        outs.append(bank.process_block(N if blk != 3 else 1000, flags=kb.PER_VOICE))
    np.save(path, np.concatenate([o.reshape(inst * voices, -1) for o in outs], axis=1))
    out = torch.empty(bank.out_shape(N), dtype=torch.float32, device="cuda")
    for _ in range(3): bank.process_into(out, N)
    bank.profile(True)
    for _ in range(20): bank.process_into(out, N)
    ms, n = bank.profile_read()
    print(f"kernel {ms / n * 1e3:.1f} us", flush=True)

if len(sys.argv) > 2 and sys.argv[1] == "child":
    child(sys.argv[2])
else:
    os.makedirs("gpurun_out/c2ab", exist_ok=True)
    ref = None
    for i, (name, env) in enumerate(VARIANTS):
        p = f"/tmp/c2ab_{i}.npy"
        r = subprocess.run([sys.executable, __file__, "child", p], env={**os.environ, **env}, capture_output=True, text=True, timeout=120)
        if r.returncode != 0:
            print(name, "FAILED", r.stderr[-800:]); continue
        x = np.load(p)
        if ref is None: ref = x
        same = x.shape == ref.shape and np.array_equal(x.view(np.uint32), ref.view(np.uint32))
        print(f"{name:16s} {r.stdout.strip():20s} bit-identical to layout 2: {same}  (nonzero {np.count_nonzero(x)})", flush=True)
