"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck): short blocks, few voices."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import klang_b200 as kb

fs = 48000.0
for graph, inst, voices, n in ((kb.SY_SUBTRACTIVE, 2, 16, 300), (kb.SY_SUBTRACTIVE, 8, 100, 260), (kb.SY_SUPERSAW, 2, 32, 300), (kb.SY_TB303, 1, 32, 300), (kb.SY_SYNTHX, 1, 32, 70), (kb.SY_FILTER_K, 1, 32, 129), (kb.SY_FM, 2, 21, 300),
                                (kb.SY_RELEASE, 2, 21, 300), (kb.SY_AM, 2, 21, 300), (kb.SY_MOD_FM2, 1, 32, 200), (kb.SY_ADDITIVE_SQUARE, 2, 21, 300), (kb.SY_BREAKPOINT, 1, 32, 130)):
    b = kb.SynthBank(graph, inst, voices, fs, n)
    for g in range(0, inst * b.voices, 2):
        b.voice_start(g % b.voices, 40 + g % 30, 0.7, g // b.voices)
    for k in range(3):
        if k == 1:
            b.voice_release(0, 0.0, 0)
        b.process_block(n)
        b.process_block(n // 2, kb.PER_VOICE)
        b.process_block(n, kb.BANK_MIX | kb.MIX_SUM)
    b.process_block(n, kb.LANE_PER_VOICE)
    b.close()
for graph, n, blocks in ((kb.FX_GAIN, 1001, 2), (kb.FX_PINGPONG, 2048, 24), (kb.FX_REVERB, 700, 3), (kb.FX_DELAY_PINGPONG, 3000, 3), (kb.FX_DELAY_REVERB, 500, 2),
                         (kb.FX_PAN, 1001, 2), (kb.FX_TREMOLO, 1024, 2), (kb.FX_CLIPPING, 7, 2), (kb.FX_ECHO, 3000, 3), (kb.FX_FEEDBACK, 3000, 3),
                         (kb.FX_FLANGER, 3000, 3), (kb.FX_MOD_CHORUS, 3000, 3), (kb.FX_MODDELAY, 500, 2), (kb.FX_WAHWAH, 500, 2), (kb.FX_IIR, 500, 2), (kb.FX_MUTE, 9, 2)):
    fx = kb.FxBank(graph, 3, fs, n)
    if graph in (kb.FX_ECHO, kb.FX_FEEDBACK):
        fx.set_control(0, 0.004)                      # a 192-frame delay: several chunks per block
    x = (np.random.default_rng(1).random((3, fx.channels, n), dtype=np.float32) - 0.5)
    for k in range(blocks):
        fx.process_inplace(x.copy())
    if graph == kb.FX_REVERB:                          # the tolerance schedule (parallel-scan line filters) and a ragged block (round-1 pipeline)
        for c, v in ((0, 0.3), (2, 0.4), (3, 0.5)):
            fx.set_control(c, v)
        fx.process_inplace(x.copy(), flags=kb.FX_TOLERANCE)
        fx.process_inplace(x.copy(), flags=kb.FX_TOLERANCE)
        print(graph, "tolerance instances", fx.tolerance_instances())
        fx.process_inplace(x[:, :, :333].copy())
    print(graph, "parallel instances", fx.parallel_instances())
    fx.close()
# the mix-down protocol on one rank (put / acquire + publish / collect over more steps than slot parities)
import ctypes as C
import torch
L = kb.lib()
h = L.kb_mixdown_create(0, 1, 0, 512)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
src = torch.rand(512, device="cuda"); dst = torch.empty(512, device="cuda")
for k in range(5):
    L.kb_mixdown_put(h, src.data_ptr(), 512, ts.cuda_stream)
    L.kb_mixdown_collect(h, dst.data_ptr(), 512, ts.cuda_stream)
torch.cuda.synchronize()
assert torch.equal(src, dst)
prev = torch.zeros(512, device="cuda")
for k in range(5):                                      # the fused step: store + flag + rank 0 sums the previous step
    L.kb_mixdown_step(h, src.data_ptr(), 512, prev.data_ptr(), ts.cuda_stream)
L.kb_mixdown_collect(h, dst.data_ptr(), 512, ts.cuda_stream)
torch.cuda.synchronize()
assert torch.equal(src, dst) and torch.equal(src, prev)
L.kb_mixdown_destroy(h)
# round 2, second half: the decoupled C2 kernel with the staged voice upload (8 x 100 = 800 voices -> kb_sub_mbar_kernel<7>) and re-triggers in every
# block, the fused voice-sum / bank-mix launch, debug taps, klang::Sample, and translated programs (an effect with a delay line, a synth, noise)
b = kb.SynthBank(kb.SY_SUBTRACTIVE, 8, 100, fs, 260)
for g in range(800):
    b.voice_start(g % 100, 40 + g % 30, 0.7, g // 100)
for k in range(4):
    for g in range(k, 800, 17):
        b.voice_start(g % 100, 45 + g % 20, 0.6, g // 100)
    b.process_block(260 if k != 2 else 131, kb.BANK_MIX | kb.MIX_SUM)
b.close()
for graph in (kb.FX_PINGPONG, kb.FX_RM, kb.FX_MODDELAY):
    fx = kb.FxBank(graph, 3, fs, 500)
    fx.debug_enable(True)
    fx.process_inplace(np.zeros((3, fx.channels, 500), np.float32))
    assert fx.debug_read(500) is not None
    fx.close()
kb.Engine().sample(np.linspace(-1, 1, 300, dtype=np.float32), 250, 440.0, 0.001)
from klang_b200 import kcc
kdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "_k_bin", "kcc")
if os.path.isfile(os.path.join(kdir, "libecho_k.so")):
    u = kcc.UserFx(os.path.join(kdir, "libecho_k.so"), 2, fs, 400)
    u.set_control(0, 0.002)
    for k in range(3):
        u.process_inplace(np.ones((2, 1, 400), np.float32))
    u.close()
    u = kcc.UserFx(os.path.join(kdir, "libk_objects_k.so"), 2, fs, 400)
    u.process_inplace(np.zeros((2, 1, 400), np.float32))
    u.close()
    sy = kcc.UserSynth(os.path.join(kdir, "libsynth_supersaw_k.so"), 2, fs, 300)
    for v in range(5):
        sy.note_on(50 + v, 0.7, v % 2)
    sy.process_block(300)
    sy.note_off(50, 0.0, 0)
    sy.process_block(300, per_voice=True)
    sy.close()
print("sanitize run done")
