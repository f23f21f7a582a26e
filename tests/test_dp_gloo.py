"""CPU (-m "not gpu"): the N>1 host logic of the data-parallel step on world_size=2 / gloo — flat parameter+gradient
arenas, ONE all-reduce per optimizer step, 1/world_size folded into the update, sample sharding.  The CUDA update
kernels cannot run here, so the test injects the oracle's torch implementation of the same update formula in their place
(the kernels themselves are checked against torch.optim on the GPU: tests/test_parity_gpu.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torch_rmsprop(p, g, sq, lr, alpha=0.99, eps=1e-8, grad_scale=1.0, clip=0.0):
    gr = g * grad_scale
    sq.mul_(alpha).addcmul_(gr, gr, value=1 - alpha)
    p.addcdiv_(gr, sq.sqrt().add_(eps), value=-lr)
    if clip > 0:
        p.clamp_(-clip, clip)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from shapegan_b200 import raw, train
    raw.rmsprop = _torch_rmsprop                       # checker stands in for sg_rmsprop on this CPU-only host
    torch.manual_seed(0)                               # identical initial parameters on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 1))
    opt = train.FlatOptimizer(net.parameters(), 'rmsprop', lr=1e-2, clip=0.05, world_size=world)
    assert all(p.data.data_ptr() >= opt.flat.data_ptr() for p in net.parameters())      # parameters are arena views
    g = torch.Generator().manual_seed(100)
    data = torch.randn((8, 6), generator=g)            # global batch of 8 samples; rank r owns rows [4r, 4r+4)
    for _ in range(3):
        opt.zero_grad()
        net(data[4 * rank:4 * rank + 4]).pow(2).mean().backward()
        assert net[0].weight.grad.data_ptr() >= opt.grad.data_ptr()                     # autograd wrote into the arena
        opt.step()
    out[rank] = opt.flat.clone()
    dist.destroy_process_group()


def test_flat_optimizer_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert torch.equal(a, b), 'ranks diverged'
    # single-process reference: same model, torch.optim.RMSprop on the mean of the two shard losses, then clip
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 1))
    opt = torch.optim.RMSprop(net.parameters(), lr=1e-2)
    data = torch.randn((8, 6), generator=torch.Generator().manual_seed(100))
    for _ in range(3):
        opt.zero_grad()
        (0.5 * (net(data[:4]).pow(2).mean() + net(data[4:]).pow(2).mean())).backward()
        opt.step()
        with torch.no_grad():
            for p in net.parameters():
                p.clamp_(-0.05, 0.05)
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert torch.allclose(a, ref, rtol=1e-5, atol=1e-6)
