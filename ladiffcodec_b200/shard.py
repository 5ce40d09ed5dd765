"""Clip sharding across the GPUs of one box.

Clips are independent, so the path shards with no data-path collective: rank r decodes the contiguous range
``clip_range(n, world, r)`` with replicated weights.  When the clips start on one rank (a serving front-end),
``scatter_clips`` / ``gather_clips`` move ``[B,1,T]`` fp32 waveforms with one ``torch.distributed`` scatter / gather each
(NCCL over NVLink on GPUs, gloo in the CPU tests) — 154 KB per clip per direction, outside the step loop.
"""
import torch
import torch.distributed as dist


def clip_range(n_clips, world, rank):
    """Balanced contiguous partition: the first ``n % world`` ranks get one extra clip."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def scatter_clips(wav, n_clips, T, src=0, device=None, group=None):
    """``wav`` [n_clips,1,T] on rank ``src`` (ignored elsewhere) → this rank's ``[n_local,1,T]`` slice."""
    world, rank = _world(group)
    lo, hi = clip_range(n_clips, world, rank)
    if world == 1:
        return wav[lo:hi]
    device = device or (wav.device if wav is not None else torch.device("cpu"))
    per = -(-n_clips // world)                       # equal-size chunks (padded) keep it a single collective
    out = torch.empty(per, 1, T, dtype=torch.float32, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = clip_range(n_clips, world, r)
            c = torch.zeros(per, 1, T, dtype=torch.float32, device=device)
            c[: b - a] = wav[a:b].to(device=device, dtype=torch.float32)
            chunks.append(c)
    dist.scatter(out, chunks, src=src, group=group)
    return out[: hi - lo]


def gather_clips(local, n_clips, dst=0, group=None):
    """Inverse of ``scatter_clips``: returns ``[n_clips,1,T]`` on rank ``dst`` (None elsewhere)."""
    world, rank = _world(group)
    if world == 1:
        return local
    per = -(-n_clips // world)
    T = local.shape[-1]
    buf = torch.zeros(per, 1, T, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        a, b = clip_range(n_clips, world, r)
        out.append(parts[r][: b - a])
    return torch.cat(out, 0)


def synthesize_sharded(decode_fn, wav, n_clips, T, src=0, device=None, group=None):
    """scatter → ``decode_fn(local_wav) -> local_out`` on every rank → gather.  ``decode_fn`` is
    ``lambda w: sample.synthesize(model, cmodel, w, ...)`` in production."""
    local = scatter_clips(wav, n_clips, T, src=src, device=device, group=group)
    out = decode_fn(local) if local.shape[0] else local
    return gather_clips(out, n_clips, dst=src, group=group)
