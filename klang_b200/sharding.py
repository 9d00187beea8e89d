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
