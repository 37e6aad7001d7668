"""Training all-reduce on CPU: two gloo ranks back-propagate different halves of a batch through the same small network; after
GradientAllReducer.finish() every rank holds the gradient of the MEAN loss over the whole batch (= what one process would compute)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dif_pan_b200.ddp import GradientAllReducer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GroupNorm(1, 8), torch.nn.SiLU(), torch.nn.Conv2d(8, 8, 3, padding=1),
                               torch.nn.Conv2d(8, 3, 1))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_grad_enabled(True)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 3, 8, 8, generator=g), torch.randn(8, 3, 8, 8, generator=g)
    ref = _model()
    (ref(x) - y).abs().mean().backward()
    net = _model()
    red = GradientAllReducer(net.parameters(), bucket_bytes=1024)       # several buckets
    ok = len(red.buckets) > 1
    for it in range(2):                                                   # second iteration: buffers are reused after zero_grad()
        red.zero_grad()
        lo, hi = rank * 4, rank * 4 + 4
        (net(x[lo:hi]) - y[lo:hi]).abs().mean().backward()
        red.finish()
        for p, pr in zip(net.parameters(), ref.parameters()):
            ok = ok and torch.allclose(p.grad, pr.grad, rtol=1e-5, atol=1e-7)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_all_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_single_process_is_a_no_op():
    torch.set_grad_enabled(True)   # sibling test modules switch autograd off at import
    net = _model()
    red = GradientAllReducer(net.parameters())
    red.zero_grad()
    net(torch.randn(2, 3, 8, 8)).sum().backward()
    g0 = [p.grad.clone() for p in net.parameters()]
    red.finish()
    assert all(torch.equal(a, p.grad) for a, p in zip(g0, net.parameters()))
    assert red.nbytes == sum(p.numel() * 4 for p in net.parameters())
