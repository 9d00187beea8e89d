"""GPU tests of the multi-GPU mix-down (kb_mixdown_*): the single-rank round trip on one device, and — when the
box has two devices — two processes whose bank mixes meet in rank 0's arena over peer memory, compared bit for bit with the
rank-order fp32 sum of the same banks rendered on one device."""
import os
import socket

import numpy as np
import pytest

import klang_b200 as kb

pytestmark = pytest.mark.gpu


def _bank(device, first_voice, fs, n, inst=2, voices=32):
    b = kb.SynthBank(kb.SY_SUBTRACTIVE, inst, voices, fs, n, device)
    for g in range(inst * voices):
        b.voice_start(g % voices, 40 + (first_voice + g) % 40, 0.5 + 0.4 * ((first_voice + g) % 3) / 2.0, g // voices)
    return b


def test_mixdown_single_rank_round_trip():
    import torch
    if kb.device_count() < 1:
        pytest.skip("needs a CUDA device")
    L = kb.lib()
    n, fs = 1024, 48000.0
    h = L.kb_mixdown_create(0, 1, 0, n)
    assert h, L.kb_last_error()
    ts = torch.cuda.Stream()            # (a NULL stream handle means "the bank's own stream" in the ABI: use a real one)
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    # (a) put / collect of arbitrary data, more steps than the two slot parities
    x = torch.rand(5, n, device="cuda")
    out = torch.empty(n, device="cuda")
    for k in range(5):
        assert L.kb_mixdown_put(h, x[k].data_ptr(), n, stream) == 0
        assert L.kb_mixdown_collect(h, out.data_ptr(), n, stream) == 0
        torch.cuda.synchronize()
        assert torch.equal(out, x[k])
    # (b) the bank-mix kernel writing the slot itself (KB_DEVICE_PTR | KB_BANK_MIX with the slot as `out`)
    a, b = _bank(0, 0, fs, n), _bank(0, 0, fs, n)
    want = torch.empty(1, n, device="cuda")
    for blk in range(3):
        a.process_into(want, n, kb.BANK_MIX | kb.MIX_SUM)
        b.set_stream(stream)
        p = L.kb_mixdown_acquire(h, stream)
        assert p
        b.process_into_device_ptr(p, n, kb.BANK_MIX | kb.MIX_SUM)
        assert L.kb_mixdown_publish(h, stream) == 0
        assert L.kb_mixdown_collect(h, out.data_ptr(), n, stream) == 0
        a.sync()
        torch.cuda.synchronize()
        assert torch.equal(out, want[0]) and float(out.abs().max()) > 1e-3
    a.close(); b.close()
    L.kb_mixdown_destroy(h)
    # (c) bad arguments
    assert not L.kb_mixdown_create(0, 2, 2, n)
    assert b"bad argument" in L.kb_last_error()


def test_mixdown_fused_step_single_rank():
    """kb_mixdown_step / kb_synth_bank_process_mixdown on one rank: every step's sum arrives one step later through out_prev, the last
    one through collect(); more steps than slot parities, so the in-kernel back-pressure and the consumed counter are exercised."""
    import torch
    if kb.device_count() < 1:
        pytest.skip("needs a CUDA device")
    L = kb.lib()
    n, fs = 1024, 48000.0
    h = L.kb_mixdown_create(0, 1, 0, n)
    assert h, L.kb_last_error()
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    x = torch.rand(7, n, device="cuda")
    prev = torch.zeros(7, n, device="cuda")
    torch.cuda.synchronize()
    for k in range(7):
        assert L.kb_mixdown_step(h, x[k].data_ptr(), n, prev[k].data_ptr(), stream) == 0, L.kb_last_error()
    last = torch.empty(n, device="cuda")
    assert L.kb_mixdown_collect(h, last.data_ptr(), n, stream) == 0
    torch.cuda.synchronize()
    for k in range(1, 7):
        assert torch.equal(prev[k], x[k - 1]), f"step {k}"
    assert torch.equal(last, x[6])
    # a bank whose bank-mix kernel is the fused step, against the plain KB_BANK_MIX call
    a, b = _bank(0, 0, fs, n), _bank(0, 0, fs, n)
    b.set_stream(stream)
    want = torch.empty(4, 1, n, device="cuda")
    got = torch.zeros(5, 1, n, device="cuda")
    for blk in range(4):
        a.process_into(want[blk], n, kb.BANK_MIX | kb.MIX_SUM)
        assert L.kb_synth_bank_process_mixdown(b.h, h, got[blk].data_ptr(), n, kb.MIX_SUM) == 0, L.kb_last_error()
    assert L.kb_mixdown_collect(h, got[4].data_ptr(), n, stream) == 0
    a.sync()
    torch.cuda.synchronize()
    assert float(got[0].abs().max()) == 0.0                    # (the step before the bank's first block was already collected above)
    for blk in range(4):
        assert torch.equal(got[blk + 1], want[blk]) and float(want[blk].abs().max()) > 1e-3, f"block {blk}"
    a.close(); b.close()
    L.kb_mixdown_destroy(h)
    # the same with out_prev in page-locked host memory (the exchange kernel stores the sum over PCIe: bench.py's end-to-end leg at N > 1)
    h = L.kb_mixdown_create(0, 1, 0, n)
    a, b = _bank(0, 0, fs, n), _bank(0, 0, fs, n)
    b.set_stream(stream)
    want = torch.empty(6, 1, n, device="cuda")
    host = [torch.zeros(1, n).pin_memory() for _ in range(7)]
    for blk in range(6):
        a.process_into(want[blk], n, kb.BANK_MIX | kb.MIX_SUM)
        b.process_mixdown(_Mix(h), host[blk], n, kb.MIX_SUM)
    last = torch.empty(1, n, device="cuda")
    assert L.kb_mixdown_collect(h, last.data_ptr(), n, stream) == 0
    a.sync()
    torch.cuda.synchronize()
    for blk in range(5):
        assert torch.equal(host[blk + 1], want[blk].cpu()), f"pinned out_prev, block {blk}"
    assert torch.equal(last, want[5])
    a.close(); b.close()
    L.kb_mixdown_destroy(h)


class _Mix:                                                        # (process_mixdown reads `.h` of a sharding.PeerMixdown)
    def __init__(self, h):
        self.h = h


def _worker(rank, world, port, n, fs, out_path):
    import torch
    import torch.distributed as dist
    from klang_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    mix = sharding.PeerMixdown(rank, n)
    bank = _bank(rank, rank * 64, fs, n)
    bank.set_stream(stream)
    got = torch.empty(4, n, device="cuda")
    for blk in range(4):
        bank.process_into_device_ptr(mix.acquire(stream), n, kb.BANK_MIX | kb.MIX_SUM)
        mix.publish(stream)
        if rank == 0:
            mix.collect(got[blk], n, stream)
    # the same four blocks again through the fused step (one kernel per block: store + flag, rank 0 sums the previous block)
    fused = torch.zeros(5, 1, n, device="cuda")
    for blk in range(4):
        bank.process_mixdown(mix, fused[blk] if rank == 0 else None, n, kb.MIX_SUM)
    if rank == 0:
        mix.collect(fused[4], n, stream)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        np.save(out_path, got.cpu().numpy())
        np.save(out_path + ".fused.npy", fused[1:, 0].cpu().numpy())
    mix.close()
    bank.close()
    dist.destroy_process_group()


def test_mixdown_two_gpus_equals_rank_order_sum(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if kb.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n, fs = 2048, 48000.0
    out_path = str(tmp_path / "mix.npy")
    mp.spawn(_worker, args=(2, port, n, fs, out_path), nprocs=2, join=True)
    got = np.concatenate([np.load(out_path), np.load(out_path + ".fused.npy")])
    banks = [_bank(0, r * 64, fs, n) for r in range(2)]
    for blk in range(8):
        parts = [b.process_block(n, kb.BANK_MIX | kb.MIX_SUM)[0] for b in banks]
        want = parts[0] + parts[1]                                   # rank order, fp32
        assert np.array_equal(got[blk].view(np.uint32), want.view(np.uint32)), f"block {blk}"
        assert np.abs(want).max() > 1e-3
    for b in banks:
        b.close()
