"""N > 1 host logic on CPU: world_size-2 gloo run of the clip scatter → per-rank decode → gather plumbing
(ladiffcodec_b200/shard.py).  The per-rank decode is a stand-in per-clip function (the CUDA path needs a GPU); what is
checked is that every clip is decoded exactly once, by the rank the partition names, and comes back in order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ladiffcodec_b200.shard import clip_range, synthesize_sharded


def test_clip_range_partitions():
    for n in (0, 1, 5, 32, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [clip_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        clip_range(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_clips, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wav = None
        if rank == 0:
            wav = torch.arange(n_clips, dtype=torch.float32).reshape(n_clips, 1, 1).expand(n_clips, 1, T).contiguous()
        seen = []

        def decode(w):                                     # per-clip, rank-tagged stand-in for sample.synthesize
            seen.append(w[:, 0, 0].clone())
            return w * 2.0 + 1000.0 * rank

        out = synthesize_sharded(decode, wav, n_clips, T, src=0)
        lo, hi = clip_range(n_clips, world, rank)
        got = torch.cat(seen) if seen else torch.empty(0)
        assert torch.equal(got, torch.arange(lo, hi, dtype=torch.float32)), (rank, got)
        if rank == 0:
            exp = torch.arange(n_clips, dtype=torch.float32) * 2.0
            for r in range(world):
                a, b = clip_range(n_clips, world, r)
                exp[a:b] += 1000.0 * r
            assert out.shape == (n_clips, 1, T)
            assert torch.equal(out[:, 0, 0], exp) and torch.equal(out[:, 0, -1], exp)
        else:
            assert out is None
        q.put((rank, "ok"))
    except Exception as e:                                 # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 8])
def test_scatter_decode_gather_world2(n_clips):
    world, T = 2, 640
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
