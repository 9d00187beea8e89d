"""Multi-GPU host logic of the path: banks shard by *instance* (a Synth / Effect object is never split: its voices share
synth-level post-fx, klang.h:4851) and the ranks' [channels][n] bank mixes are combined with one sum-reduce per block —
the path's only exchange.  Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""


def shard_instances(total_instances, rank, world):
    """Contiguous instance range [lo, hi) of `rank`; sizes differ by at most one, lower ranks take the remainder."""
    if not (0 <= rank < world) or total_instances < 0:
        raise ValueError("bad rank / world / total")
    base, rem = divmod(total_instances, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def voice_ids(total_instances, voices_per_instance, rank, world):
    """Global voice ids (instance-major) owned by `rank` — the ids the note schedules are keyed on."""
    lo, hi = shard_instances(total_instances, rank, world)
    return range(lo * voices_per_instance, hi * voices_per_instance)


def reduce_mix(mix, dst=0, group=None):
    """Sum the per-rank bank mixes into rank `dst` (in place). `mix` is a torch tensor [channels, n] on the rank's device."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(mix, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return mix


class PeerMixdown:
    """The mix-down over NVLink peer memory (include/klang_b200.h: kb_mixdown_*): every rank's bank-mix kernel stores its
    [channels][n] mix straight into rank 0's arena; rank 0 sums the slots in rank order.  One process per GPU of one node;
    the 64-byte CUDA IPC handle of the arena travels once, at construction, through torch.distributed.

    per step:   ptr = mix.acquire(stream); bank.process_into_device_ptr(ptr, n, BANK_MIX | ...); mix.publish(stream)
                rank 0 only: mix.collect(dst_tensor, count, stream)
    """

    def __init__(self, device_index, max_floats, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import api
        self._api = api
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = api.lib()
        self.h = L.kb_mixdown_create(device_index, self.world, self.rank, int(max_floats))
        if not self.h:
            raise api.KlangB200Error("kb_mixdown_create: " + L.kb_last_error().decode())
        box = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(64)
            api._check(L.kb_mixdown_export(self.h, buf), "kb_mixdown_export")
            box[0] = bytes(buf.raw)
        dist.broadcast_object_list(box, src=0, group=group)
        ok = 1
        if self.rank != 0:
            rc = L.kb_mixdown_import(self.h, C.create_string_buffer(box[0], 64))
            ok = 1 if rc >= 0 else 0
            self.error = None if ok else L.kb_last_error().decode()
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=group)          # every rank learns whether the peer mapping exists everywhere
        if not all(flags):
            self.close()
            raise api.KlangB200Error("kb_mixdown_import failed on a rank (CUDA IPC / peer access unavailable)")

    def acquire(self, cuda_stream):
        p = self._api.lib().kb_mixdown_acquire(self.h, cuda_stream)
        if not p:
            raise self._api.KlangB200Error("kb_mixdown_acquire: " + self._api.lib().kb_last_error().decode())
        return p

    def publish(self, cuda_stream):
        self._api._check(self._api.lib().kb_mixdown_publish(self.h, cuda_stream), "kb_mixdown_publish")

    def put(self, src, count, cuda_stream):
        """acquire + device copy of a finished local mix (torch CUDA tensor) into the slot + publish."""
        self._api._check(self._api.lib().kb_mixdown_put(self.h, src.data_ptr(), int(count), cuda_stream), "kb_mixdown_put")

    def step(self, src, count, out_prev, cuda_stream):
        """The fused form (kb_mixdown_step): one kernel stores `src` into this rank's slot and raises its flag; on rank 0 it also sums the
        PREVIOUS step's slots into `out_prev`."""
        self._api._check(self._api.lib().kb_mixdown_step(self.h, src.data_ptr(), int(count), out_prev.data_ptr() if out_prev is not None else 0, cuda_stream), "kb_mixdown_step")

    def stream_wait(self, cuda_stream):
        """`cuda_stream` waits for the exchange kernel of the last fused step (before it reads that step's out_prev)."""
        self._api._check(self._api.lib().kb_mixdown_stream_wait(self.h, cuda_stream), "kb_mixdown_stream_wait")

    def host_wait(self, back=0):
        """the host blocks until the exchange kernel of the fused step `back` steps before the last one has finished (its out_prev is valid)"""
        self._api._check(self._api.lib().kb_mixdown_host_wait(self.h, int(back)), "kb_mixdown_host_wait")

    def collect(self, dst, count, cuda_stream):
        self._api._check(self._api.lib().kb_mixdown_collect(self.h, dst.data_ptr(), int(count), cuda_stream), "kb_mixdown_collect")

    def close(self):
        if getattr(self, "h", None):
            self._api.lib().kb_mixdown_destroy(self.h)
            self.h = None
