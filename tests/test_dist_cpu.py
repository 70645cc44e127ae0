"""CPU, world_size 2 over gloo: the host-side data-parallel logic (hgnn-2_b200/dist.py) - contiguous
graph sharding, flat parameter/gradient buffers, broadcast and the single flat-gradient all-reduce.
The CUDA kernels are not involved (a tiny torch module stands in for the model)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hgnn_b200  # noqa: F401
from hgnn_b200.dist import FlatParams, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                    # different init per rank on purpose
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
    fp = FlatParams(model)
    fp.broadcast(0)                                   # now identical everywhere
    gen = torch.Generator().manual_seed(7)
    X, y = torch.randn(16, 5, generator=gen), torch.randint(0, 2, (16,), generator=gen)
    lo, hi = shard_range(16, rank, world)
    fp.zero_grad()
    loss = torch.nn.functional.cross_entropy(model(X[lo:hi]), y[lo:hi], reduction="sum") / 16
    loss.backward()
    assert model[0].weight.data_ptr() == fp.flat.data_ptr()          # parameters are views of the flat buffer
    fp.all_reduce_grad()
    out[rank] = (fp.flat.clone(), fp.grad.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        (p0, g0), (p1, g1) = out[0], out[1]
    assert torch.equal(p0, p1) and torch.equal(g0, g1)
    # single-process reference: full batch, same parameters
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
    gen = torch.Generator().manual_seed(7)
    X, y = torch.randn(16, 5, generator=gen), torch.randint(0, 2, (16,), generator=gen)
    torch.nn.functional.cross_entropy(model(X), y).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(g0, ref, atol=1e-6)
