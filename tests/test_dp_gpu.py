"""GPU (-m gpu, needs >= 2 GPUs; skipped on the single-GPU boxes): the fused data-parallel optimizer step over NVLink peer memory
(csrc/sg_dp.cu through train.FlatOptimizer) against torch.optim on the rank-averaged gradient, and against the NCCL path."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, kind, mode, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), SG_B200_DP=mode)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
    from shapegan_b200 import train
    g = torch.Generator().manual_seed(5)
    sizes = [(257, 33), (1001,), (64, 3, 4, 4, 4), (7,)]                     # total not a multiple of 4 * world
    params = [torch.nn.Parameter((torch.randn(s, generator=g) * 0.02).cuda()) for s in sizes]
    opt = train.FlatOptimizer(params, kind, 5e-5 if kind == 'rmsprop' else 1e-3, clip=0.01 if kind == 'rmsprop' else 0.0, world_size=world)
    assert (mode == 'fused') == (opt.peers is not None), opt.dp_note
    ref = [torch.nn.Parameter(p.detach().clone()) for p in params]
    ropt = (torch.optim.RMSprop(ref, lr=5e-5) if kind == 'rmsprop' else torch.optim.Adam(ref, lr=1e-3))
    for step in range(3):
        opt.zero_grad()
        grads = []
        for r in range(world):
            gg = torch.Generator().manual_seed(100 * step + r)
            grads.append([torch.randn(s, generator=gg).cuda() for s in sizes])
        for p, gr in zip(params, grads[rank]):
            p.grad.add_(gr)
        opt.step()
        for i, p in enumerate(ref):
            p.grad = sum(grads[r][i] for r in range(world)) / world
        ropt.step()
        if kind == 'rmsprop':
            with torch.no_grad():
                for p in ref:
                    p.clamp_(-0.01, 0.01)
    torch.cuda.synchronize()
    worst = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(params, ref))
    flat = opt.flat.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    from shapegan_b200 import _lib as L
    err = L.lib().sg_check_device_error()
    if rank == 0:
        torch.save({'worst': worst, 'same': same, 'err': err, 'note': opt.dp_note}, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('kind', ['rmsprop', 'adam'])
@pytest.mark.parametrize('mode', ['fused', 'nccl'])
def test_data_parallel_step(kind, mode, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, kind, mode, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r['err'] == 0
    assert r['same'], 'replicas diverged: ' + r['note']
    assert r['worst'] < 2e-6, (r['worst'], r['note'])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_modules_survive_nn_dataparallel():
    """The reference wraps its models in nn.DataParallel when it sees several GPUs (train_hybrid_progressive_gan.py:62-71).
    Our modules must survive that wrapper unchanged: replicate() shallow-copies the module (the LinOp objects are shared, the parameters
    of a replica are plain broadcast tensors without an optimizer arena), the scattered batch runs on both devices, the gradients are
    reduced onto the original parameters.  Compared with the same modules run un-wrapped on the whole batch."""
    from shapegan_b200 import config
    from model.gan import Discriminator
    from model.progressive_gan import Discriminator as ProgressiveDiscriminator
    old = config.precision()
    config.set_precision('fp32x')
    try:
        torch.manual_seed(3)
        cases = [(Discriminator().cuda(0), torch.randn(4, 32, 32, 32).cuda(0) * 0.1)]
        pd = ProgressiveDiscriminator().cuda(0)
        pd.set_iteration(1)
        cases.append((pd, torch.randn(4, 16, 16, 16).cuda(0) * 0.1))
        for module, x in cases:
            ref = module(x)
            ref.sum().backward()
            gref = [p.grad.detach().clone() for p in module.parameters() if p.grad is not None]
            module.zero_grad(set_to_none=True)
            par = torch.nn.DataParallel(module, device_ids=[0, 1])
            if hasattr(module, 'iteration'):
                par.module.set_iteration(1)
            out = par(x)
            assert out.shape == ref.shape
            assert torch.allclose(out, ref, rtol=2e-3, atol=2e-4), (out - ref).abs().max()
            out.sum().backward()
            got = [p.grad.detach().clone() for p in module.parameters() if p.grad is not None]
            assert len(got) == len(gref)
            for a, b in zip(got, gref):
                assert torch.allclose(a, b, rtol=5e-3, atol=5e-4 * float(b.abs().max() + 1e-6)), (a - b).abs().max()
    finally:
        config.set_precision(old)
