"""World-size-2 gloo tests of the clip-sharded data-parallel path: contiguous partition, zero
data-path collectives, one all-gather of the waveforms.  The sampler is replaced by a deterministic per-clip
function, so this covers the host logic on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from syncfusion_b200.parallel import gather_waveforms, sample_sharded, shard_bounds


def test_shard_bounds_cover_batch():
    for batch in (1, 5, 16, 17, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_sample(x_noisy, num_steps, channels, embedding, embedding_scale):
    # per-clip function of every input, so a mis-sharded tensor changes the result
    return x_noisy * num_steps + channels[0].mean(dim=(1, 2), keepdim=True) + embedding.sum(dim=(1, 2)).reshape(-1, 1, 1) * embedding_scale


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 1, 64, generator=g)
    ch = [torch.randn(batch, 2, 64, generator=g), torch.randn(batch, 8, 16, generator=g)]
    e = torch.randn(batch, 1, 512, generator=g)
    out = sample_sharded(_fake_sample, x, 5, ch, e, 2.0)
    ref = _fake_sample(x, 5, ch, e, 2.0)
    ok = out.shape == ref.shape and torch.allclose(out, ref)
    local = sample_sharded(_fake_sample, x, 5, ch, e, 2.0, gather=False)
    lo, hi = shard_bounds(batch, world, rank)
    ok = ok and torch.allclose(local, ref[lo:hi])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_sample_sharded_world2_even():
    _run(8)


def test_sample_sharded_world2_ragged():
    _run(5)


def test_gather_is_identity_without_process_group():
    x = torch.randn(3, 1, 16)
    assert gather_waveforms(x, 3) is x
