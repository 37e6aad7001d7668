"""N>1 path on CPU: world_size-2 gloo processes shard a batch of patches, 'sample' their shard with a deterministic
per-patch stand-in (the CUDA kernels need a GPU) and all_gather the result; the gathered batch must equal the
single-process result for even and ragged batch sizes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dif_pan_b200.sharding import gather_patches, sample_sharded, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_sampler(cond, noise=None):
    out = cond[:, :8] * 2.0 + cond[:, 8:9]
    if noise is not None:
        out = out + 0.1 * noise[0] - 0.01 * noise[1]
    return out


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    cond = torch.randn(total, 20, 8, 8, generator=g)
    noise = [torch.randn(total, 8, 8, 8, generator=g) for _ in range(2)]
    full = sample_sharded(_fake_sampler, cond, noise)
    ref = _fake_sampler(cond, noise)
    ok = torch.equal(full, ref)
    lo, hi = shard_range(total, rank, world)
    ok = ok and torch.equal(gather_patches(ref[lo:hi], total), ref)

    # without injected noise the in-kernel generator is indexed by the GLOBAL patch index: sample_sharded tells the diffusion object
    # which slice of the batch this rank holds (noise_shard = (lo, total)) and resets it afterwards
    class FakeDiffusion:
        noise_shard = None

    dif = FakeDiffusion()

    def sampler_with_rng(c):
        l0, tot = dif.noise_shard
        idx = torch.arange(l0, l0 + c.shape[0], dtype=torch.float32).view(-1, 1, 1, 1)  # stand-in for Philox(seed, global index)
        return _fake_sampler(c) + idx / tot

    full2 = sample_sharded(sampler_with_rng, cond, diffusion=dif)
    ref2 = _fake_sampler(cond) + torch.arange(total, dtype=torch.float32).view(-1, 1, 1, 1) / total
    ok = ok and torch.equal(full2, ref2) and dif.noise_shard is None
    q.put((rank, bool(ok), tuple(full.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 2])
def test_two_rank_gloo_shard_and_gather(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (total, 8, 8, 8) for _, _, shape in res)


def test_shard_range_partition():
    for n in (0, 1, 7, 8, 256, 257):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
