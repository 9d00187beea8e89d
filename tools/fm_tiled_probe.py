"""Device check of the time-parallel FM.k kernel: same bank rendered by the lane-per-voice kernel (KB_LANE_PER_VOICE) and by
kb_fm_tiled_kernel (the default), over ragged blocks with releases and a re-trigger; prints one JSON line
{"equal": bool, "first_bad": ..., "tiled_us": ..., "lane_us": ...} and exits 0 when the two are bit-identical."""
import json, os, sys
os.environ["KB_FM_TILED"] = "1"          # read once by the library at the first synth block
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import klang_b200 as kb

inst, voices, fs = 3, 21, 48000.0        # 63 voices: a ragged last CTA
blocks = [4096, 117, 1, 128, 129, 1000, 4096, 4096, 300, 4096] + [4096] * 12
outs = {}
for name, flag in (("lane", kb.LANE_PER_VOICE), ("tiled", 0)):
    bank = kb.SynthBank(kb.SY_FM, inst, voices, fs, 4096)
    bank.set_control(3, 0.02)
    for g in range(0, inst * voices, 2):
        bank.voice_start(g % voices, 36 + (5 * g) % 40, 0.8, g // voices)
    res = []
    for k, n in enumerate(blocks):
        if k == 2:
            for g in range(0, inst * voices, 4):
                bank.voice_release(g % voices, 0.0, g // voices)
        if k == 5:
            bank.set_control(1, 2.5); bank.set_control(2, 0.9)
            bank.voice_start(1, 50, 0.7, 0); bank.voice_start(0, 62, 0.7, 1)
        res.append(bank.process_block(n, kb.PER_VOICE | flag))
    stages = [bank.voice_stage(v, i) for i in range(inst) for v in range(voices)]
    mix = bank.process_block(512, flag)
    bank.close()
    outs[name] = (np.concatenate(res, axis=-1), np.asarray(stages), mix)
a, b = outs["lane"], outs["tiled"]
diff = np.flatnonzero(a[0].view(np.uint32).ravel() != b[0].view(np.uint32).ravel())
equal = diff.size == 0 and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
line = {"equal": bool(equal), "mismatches": int(diff.size), "first_bad": int(diff[0]) if diff.size else None,
        "stages_equal": bool(np.array_equal(a[1], b[1])), "ended": int((a[1] == 3).sum()), "peak": float(np.abs(a[0]).max())}
if equal and "--time" in sys.argv:
    import torch
    for name, flag in (("lane_us", kb.LANE_PER_VOICE), ("tiled_us", 0)):
        bank = kb.SynthBank(kb.SY_FM, 8, 128, fs, 4096)
        for g in range(1024):
            bank.voice_start(g % 128, 36 + (5 * g) % 36, 0.8, g // 128)
        out = torch.empty(bank.out_shape(4096), dtype=torch.float32, device="cuda")
        for _ in range(3):
            bank.process_into(out, 4096, flag)
        bank.profile(True)
        for _ in range(10):
            bank.process_into(out, 4096, flag)
        ms, cnt = bank.profile_read()
        line[name] = ms / cnt * 1e3
        bank.close()
print(json.dumps(line))
sys.exit(0 if equal else 1)
