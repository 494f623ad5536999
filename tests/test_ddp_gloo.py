"""Host-side data-parallel logic on CPU: world_size 2 over gloo (127.0.0.1).  Checks that the flat-arena gradient
all-reduce reproduces the full-batch gradient, that unused parameters keep .grad = None, that duplicated parameters
are reduced once, and the strided batch shard of the reference's sampler."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Conv2d(3, 4, 3, padding=1)
        self.b = nn.Linear(4, 2)
        self.unused = nn.Linear(5, 5)   # like torchvision's fc: registered, never in the graph

    def forward(self, x):
        return self.b(torch.relu(self.a(x)).mean((2, 3)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from mono_vifi_b200 import ddp
    r, l, w = ddp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = Tiny()
    ddp.broadcast_parameters(list(model.parameters()))
    params = list(model.parameters()) + list(model.a.parameters())   # duplicates, as train.py:198-200 produces
    red = ddp.FlatGradAllReduce(params)
    g = torch.Generator().manual_seed(5)
    x, y = torch.rand(8, 3, 6, 6, generator=g), torch.rand(8, 2, generator=g)
    idx = ddp.shard_batch(list(range(8)), rank, world)
    out = []
    for it in range(2):   # second iteration exercises attach() after `used` was learnt
        red.attach()
        loss = ((model(x[idx]) - y[idx]) ** 2).mean()
        loss.backward()
        red.allreduce_mean()
        out.append([None if p.grad is None else p.grad.clone() for p in model.parameters()])
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = Tiny()
    g = torch.Generator().manual_seed(5)
    x, y = torch.rand(8, 3, 6, 6, generator=g), torch.rand(8, 2, generator=g)
    ((model(x) - y) ** 2).mean().backward()
    for it in range(2):
        for got, p in zip(out[it], model.parameters()):
            if p.grad is None:
                assert got is None       # unused parameter: stays None, the optimizer skips it
            else:
                torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-7)


def test_shard_batch_is_strided():
    from mono_vifi_b200 import ddp
    assert ddp.shard_batch(list(range(10)), 1, 4) == [1, 5, 9]
    parts = [ddp.shard_batch(list(range(24)), r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(24)) and all(len(p) == 3 for p in parts)


def test_single_process_is_identity():
    from mono_vifi_b200 import ddp
    m = Tiny()
    red = ddp.FlatGradAllReduce(list(m.parameters()))
    red.attach()
    m(torch.rand(2, 3, 6, 6)).sum().backward()
    arena = red.allreduce_mean()
    assert m.unused.weight.grad is None and m.a.weight.grad is not None
    assert m.a.weight.grad.data_ptr() >= arena.data_ptr()   # gradients live in the flat arena
