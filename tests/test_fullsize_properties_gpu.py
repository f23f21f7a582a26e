"""Size-independent properties at BASELINE.json's full sizes (configs[1]: 32^3, B=64; configs[2]: 16384 points x many
shapes; bf16 = the benchmarked mode), where the CPU oracle would take minutes:

  * samples / points are independent units: the result for a batch is BIT-EXACT the concatenation of the
    results of its parts -- this is also what lets WGANStep run critic(fake) and critic(real) as one batch of 2B;
  * weight gradients are sums over samples: grad(whole batch) == grad(first half) + grad(second half) up to fp32
    summation order;
  * the fused kernels are deterministic (no atomics on the value path): two runs give identical bits;
  * the 1-bit ReLU masks the fused forward emits are exactly `stash > 0`.
"""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture()
def bf16_mode():
    from shapegan_b200 import config
    old = config.precision()
    config.set_precision('bf16')
    yield
    config.set_precision(old)


def synth_voxels(b, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.clamp(torch.randn((b, 32, 32, 32), generator=g) * 0.05, -0.1, 0.1) / 0.1).cuda()


def test_critic_batch_independence_and_gradient_additivity(bf16_mode):
    from model.gan import Discriminator
    torch.manual_seed(3)
    dis = Discriminator()
    dis.use_sigmoid = False
    x = synth_voxels(128, 1)
    with torch.no_grad():
        whole = dis(x)
        parts = torch.cat((dis(x[:64]), dis(x[64:])))
        again = dis(x)
    # rows of the implicit GEMMs never mix; the K SPLIT of Conv3d(128->256) on the 4^3 grid depends on the batch (4 splits at 64
    # samples, 2 at 128), so the fp32 summation order -- and nothing else -- differs between the whole batch and its halves
    # (bf16 mode: a different fp32 order flips the bf16 rounding of a few activations of that layer, ~1e-4 of the score)
    assert rel_l2(whole, parts) < 1e-3
    assert torch.equal(whole, again)                 # deterministic
    grads = []
    for sl in (slice(0, 128), slice(0, 64), slice(64, 128)):
        dis.zero_grad()
        dis(x[sl]).sum().backward()
        grads.append([p.grad.clone() for p in dis.parameters()])
    for g_all, g_a, g_b in zip(*grads):
        assert rel_l2(g_all, g_a + g_b) < 2e-3       # fp32 accumulation order / split-K partials differ, values do not


def test_generator_eval_batch_independence(bf16_mode):
    from model.gan import Generator
    torch.manual_seed(4)
    gen = Generator()
    gen.eval()                                       # train-mode BatchNorm couples the samples (model/gan.py:10); eval does not
    z = torch.randn((64, 128), generator=torch.Generator().manual_seed(5)).cuda()
    with torch.no_grad():
        whole = gen(z)
        parts = torch.cat((gen(z[:32]), gen(z[32:])))
    assert list(whole.shape) == [64, 1, 32, 32, 32]
    assert torch.equal(whole, parts)


def test_sdfnet_chunk_invariance_masks_and_determinism(bf16_mode):
    from model.sdf_net import SDFNet
    from shapegan_b200 import raw, sdf_ops
    torch.manual_seed(6)
    net = SDFNet()
    shapes, per = 64, 16384                          # 1 M points: every CTA of the persistent kernel loops over many tile pairs
    n = shapes * per
    g = torch.Generator().manual_seed(7)
    pts = (torch.rand((n, 3), generator=g) * 2 - 1).cuda()
    table = (torch.randn((shapes, 128), generator=g) * 0.1).cuda()
    idx = (torch.arange(n, device='cuda') // per).to(torch.int32)
    with torch.no_grad():
        whole = net(pts, table, idx)
        again = net(pts, table, idx)
        cut = 256 * 371                              # a tile-PAIR boundary inside shape 5: bit-exact
        parts = torch.cat((net(pts[:cut], table, idx[:cut]), net(pts[cut:], table, idx[cut:])))
        odd_cut = 5 * per + 12345                    # rows change between the two tiles of a pair: layers2.0 adds its latent part
        ragged = torch.cat((net(pts[:odd_cut], table, idx[:odd_cut]), net(pts[odd_cut:], table, idx[odd_cut:])))   # before (tile A) or
        mat = net(pts[:per], table[idx[:per].long()])            # after (tile B) the hidden part: fp32 summation order only.
        # materialised [N,128] latents == indexed table
    assert torch.equal(whole, again)
    assert torch.equal(whole, parts)
    assert torch.equal(whole[:odd_cut], ragged[:odd_cut]) and rel_l2(ragged, whole) < 1e-4
    assert torch.equal(whole[:per], mat)
    assert whole.abs().max().item() <= 1.0           # tanh range (model/sdf_net.py:51)
    # stash / mask consistency of the training-mode forward
    w = [p for i, p in enumerate(net._params()) if i % 2 == 0]
    b = [p for i, p in enumerate(net._params()) if i % 2 == 1]
    img, aux = sdf_ops._fused_pack(w, b)
    m = 128 * 301 + 5
    stash = torch.empty((7, m, 256), dtype=torch.bfloat16, device='cuda')
    mstash = torch.zeros((7, m, 8), dtype=torch.int32, device='cuda')
    out = raw.sdfnet_fwd(pts[:m].contiguous(), table, idx[:m].contiguous(), img, aux, stash, mstash)
    assert torch.equal(out, whole[:m])
    assert (stash >= 0).all()
    k = torch.arange(16, device='cuda')
    for word in range(8):                            # bit k = column 32*word + 2k, bit 16+k = column 32*word + 2k + 1
        bits = mstash[:, :, word].unsqueeze(-1)
        even = ((bits >> k) & 1).bool()
        odd = ((bits >> (16 + k)) & 1).bool()
        cols = stash[:, :, 32 * word:32 * word + 32] > 0
        assert torch.equal(even, cols[:, :, 0::2]) and torch.equal(odd, cols[:, :, 1::2])


def test_autodecoder_gradient_additivity_over_shapes(bf16_mode):
    """dL/dW of a sum over points == sum of the per-half gradients; the latent gradient of a shape only depends on its own points."""
    from model.sdf_net import SDFNet
    torch.manual_seed(8)
    net = SDFNet()
    shapes, per = 16, 16384
    n = shapes * per
    g = torch.Generator().manual_seed(9)
    pts = (torch.rand((n, 3), generator=g) * 2 - 1).cuda()
    table = (torch.randn((shapes, 128), generator=g) * 0.1).cuda()
    idx = (torch.arange(n, device='cuda') // per).to(torch.int32)
    half = n // 2
    res = []
    for sl in (slice(0, n), slice(0, half), slice(half, n)):
        net.zero_grad()
        t = table.clone().requires_grad_(True)
        net(pts[sl], t, idx[sl]).sum().backward()
        res.append(([p.grad.clone() for p in net.parameters()], t.grad.clone()))
    for g_all, g_a, g_b in zip(res[0][0], res[1][0], res[2][0]):
        assert rel_l2(g_all, g_a + g_b) < 2e-3
    # first 8 shapes: only touched by the first half of the points (run-aggregated fp32 atomics: order may differ, values not)
    assert rel_l2(res[0][1][:shapes // 2], res[1][1][:shapes // 2]) < 1e-5
    assert rel_l2(res[0][1][shapes // 2:], res[2][1][shapes // 2:]) < 1e-5
    assert res[1][1][shapes // 2:].abs().max().item() == 0.0
