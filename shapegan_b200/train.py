"""Training-step bodies of the reference's scripts on the B200-native modules, with the optimizer and the
data-parallel gradient exchange fused the B200 way:

  * all parameters of a module live in ONE flat fp32 arena, all gradients in another (autograd accumulates straight
    into views of it), so a step is: backward -> ONE NCCL all-reduce of the flat gradient arena over NVLink (only
    when world_size > 1) -> ONE fused RMSprop/Adam kernel (sg_rmsprop / sg_adam) that also applies 1/world_size and,
    for the critic, the weight clip of model/gan.py:67-69.
  * batches shard over samples: every rank owns `batch` samples (weak scaling); there is no other exchange.
    BatchNorm statistics are per replica (like nn.DataParallel in the reference, train_hybrid_progressive_gan.py:62-68).

Step definitions:
  WGANStep            train_wgan.py:62-71 (critic update, clip) + :75-84 (generator update); `gp=True` adds the gradient
                      penalty of train_hybrid_progressive_gan.py:102-111 to the critic loss instead of clipping (D1).
  GANStep             train_gan.py:58-86 (G step, D-fake step, D-real step; Adam 1e-3 / 1e-5).
  AutodecoderStep     train_sdf_autodecoder.py:77-91 with `//` at :78 (two Adams, lr 1e-5, L1 + 0.01*mean(z^2)).
  VAEStep             train_autoencoder.py:98-117 (Adam 5e-5).
"""
import torch
import torch.distributed as dist

import os

from . import raw
from .critic import CriticUpdate
from .ops import arena_backward


def _fused_critic_enabled():
    return os.environ.get('SG_B200_NO_FUSED_CRITIC') != '1'


def _fused_dp_enabled():
    return os.environ.get('SG_B200_DP', 'fused') == 'fused'


class _PeerArenas:
    """Symmetric (peer-mapped) gradient / parameter arenas + signal pads of one FlatOptimizer for the fused data-parallel step
    (csrc/sg_dp.cu): every rank can load every peer's gradients and store into every peer's parameters over NVLink."""

    def __init__(self, n, dev, world, rank):
        import ctypes

        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD
        n_pad = raw.round_up(n, 4)
        self.world, self.rank = world, rank
        self.chunk = raw.round_up((n_pad + world - 1) // world, 4)
        self.flat_full = symm.empty(n_pad, dtype=torch.float32, device=dev)
        self.grad_full = symm.empty(n_pad, dtype=torch.float32, device=dev)
        self.pads = symm.empty(max(2 * world, 32), dtype=torch.int32, device=dev)
        self.flat_full.zero_(); self.grad_full.zero_(); self.pads.zero_()
        hs = [symm.rendezvous(t, group) for t in (self.grad_full, self.flat_full, self.pads)]
        self._handles = hs
        mk = lambda h: (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(h.buffer_ptrs[r])) for r in range(world)])     # noqa: E731
        self.peer_grad, self.peer_param, self.peer_pad = mk(hs[0]), mk(hs[1]), mk(hs[2])
        self.sync = torch.zeros(4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier()                                   # every rank's pads are zero before anybody signals


class FlatOptimizer:
    """Flat-arena RMSprop / Adam with torch.optim default hyper-parameters (train_wgan.py:45-46, train_gan.py:28-31).

    world_size > 1: the step is ONE kernel over NVLink peer memory (sg_dp_step: reduce-scatter of the gradient arenas by peer loads,
    the update of the rank's own shard -- optimizer state is sharded --, all-gather of the new parameters by peer stores) when the
    arenas can live in symmetric memory; otherwise (SG_B200_DP=nccl, or no symmetric memory) one NCCL all-reduce + the fused update."""

    def __init__(self, params, kind, lr, clip=0.0, world_size=1):
        self.params = [p for p in params]
        seen, uniq = set(), []
        for p in self.params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        self.kind, self.lr, self.clip, self.world = kind, lr, clip, world_size
        n = sum(p.numel() for p in self.params)
        self.n = n
        dev = self.params[0].device
        self.peers, self.dp_note = None, 'single'
        if world_size > 1:
            self.dp_note = 'nccl all-reduce + fused update'
            if _fused_dp_enabled() and dev.type == 'cuda':
                try:
                    self.peers = _PeerArenas(n, dev, world_size, dist.get_rank())
                    self.dp_note = 'sg_dp_step (peer-memory reduce-scatter + sharded update + all-gather, one kernel)'
                except Exception as e:                    # no symmetric memory on this platform / build: NCCL path
                    self.peers = None
                    self.dp_note = 'nccl all-reduce + fused update (symmetric memory unavailable: %s)' % str(e).split('\n')[0][:80]
        if self.peers is not None:
            self.flat, self.grad = self.peers.flat_full[:n], self.peers.grad_full[:n]
        else:
            self.flat = torch.empty(n, dtype=torch.float32, device=dev)
            self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)           # parameters become views of the arena
            p.grad = self.grad[off:off + k].view_as(p.data)           # autograd accumulates in place into the arena
            off += k
        if world_size > 1 and dev.type == 'cuda':
            dist.broadcast(self.flat, 0)                              # replicas start from rank 0's parameters (what nn.DataParallel does every forward)
        n_state = self.peers.chunk if self.peers is not None else n   # fused path: state of the local shard only
        self.s1 = torch.zeros(n_state, dtype=torch.float32, device=dev)     # square_avg / exp_avg
        self.s2 = torch.zeros(n_state, dtype=torch.float32, device=dev) if kind == 'adam' else None
        self.steps = 0
        from .ops import invalidate_weight_cache
        invalidate_weight_cache()

    def zero_grad(self):
        self.grad.zero_()

    def step(self):
        self.steps += 1
        scale = 1.0 / self.world
        if self.peers is not None:
            raw.dp_step(self.peers, self.n, self.s1, self.s2, self.kind, self.lr, self.steps, grad_scale=scale, clip=self.clip)
            _bump_versions(self.params)
            return
        if self.world > 1:
            dist.all_reduce(self.grad)                                 # NCCL ring/NVLS over NVLink 5
        if self.kind == 'rmsprop':
            raw.rmsprop(self.flat, self.grad, self.s1, self.lr, grad_scale=scale, clip=self.clip)
        else:
            raw.adam(self.flat, self.grad, self.s1, self.s2, self.lr, self.steps, grad_scale=scale)
        _bump_versions(self.params)


def _bump_versions(params):
    """The fused kernels write parameter memory behind autograd's back: bump the version counters (host-only) so the
    packed-weight cache and autograd's saved-tensor checks see the mutation."""
    try:
        for p in params:
            torch.autograd.graph.increment_version(p)
    except Exception:
        from .ops import invalidate_weight_cache
        invalidate_weight_cache()


def gradient_penalty(critic, real, fake, alpha, weight=10.0):
    """train_hybrid_progressive_gan.py:102-111 with alpha [B,1,1,1] injected (device RNG in the reference)."""
    a = alpha.expand(real.shape)
    xi = (a * real + (1 - a) * fake).detach().requires_grad_(True)
    out = critic(xi)
    grads = torch.autograd.grad(outputs=out, inputs=xi, grad_outputs=torch.ones_like(out), create_graph=True,
                                retain_graph=True, only_inputs=True)[0]
    return ((grads.norm(2, dim=(1, 2, 3)) - 1) ** 2).mean() * weight


class WGANStep:
    def __init__(self, generator, critic, lr=0.00005, clip=0.01, gp=False, gp_weight=10.0, world_size=1):
        self.gen, self.cri, self.gp, self.gp_weight = generator, critic, gp, gp_weight
        critic.use_sigmoid = False                                     # train_wgan.py:31
        self.gopt = FlatOptimizer(generator.parameters(), 'rmsprop', lr, world_size=world_size)
        self.copt = FlatOptimizer(critic.parameters(), 'rmsprop', lr, clip=0.0 if gp else clip, world_size=world_size)
        self.critic_update = CriticUpdate(critic)

    def __call__(self, real, z_critic, z_gen, alpha=None):
        """real [B,32,32,32] fp32, z_* [B,128] fp32, alpha [B,1,1,1] (GP variant) — all on the device.
        Returns (critic_loss, generator_loss) as 0-d device tensors (no host sync)."""
        gen, cri = self.gen, self.cri
        self.gopt.zero_grad(); self.copt.zero_grad()                    # train_wgan.py:62-63
        with torch.no_grad():
            fake = gen(z_critic)                                       # .detach() at :65: no graph is needed
        # :66-68  critic(fake) and critic(valid) as ONE batch of 2B samples: the critic has no BatchNorm (model/gan.py:48-57), so the
        # samples are independent and mean(D(fake)) - mean(D(real)) and every gradient are the same sums, in half the launches
        b = real.shape[0]
        if _fused_critic_enabled() and self.critic_update.supported():
            # :66-69 (+ the gradient penalty) as the hand-scheduled sweep of shapegan_b200/critic.py: no autograd graph
            with torch.no_grad():
                closs = self.critic_update(fake.reshape(real.shape), real, alpha if self.gp else None, self.gp_weight)[0]
        else:
            score = cri(torch.cat((fake.reshape(real.shape), real), 0))
            closs = torch.mean(score[:b]) - torch.mean(score[b:])
            if self.gp:
                closs = closs + gradient_penalty(cri, real, fake.squeeze(1), alpha, self.gp_weight)
            with arena_backward():                                           # :69
                closs.backward()
        self.copt.step()                                               # :70-71 (clip fused)
        self.gopt.zero_grad(); self.copt.zero_grad()                    # :75-76
        # The reference's backward here also fills the critic's weight gradients, which :62-63 of the next batch zeroes
        # unread.  Same update without that dead work: the critic parameters do not require grad during this backward.
        for q in self.copt.params:
            q.requires_grad_(False)
        try:
            gloss = -torch.mean(cri(gen(z_gen)))                       # :78-82
            with arena_backward():                                           # :83
                gloss.backward()
        finally:
            for q in self.copt.params:
                q.requires_grad_(True)
        self.gopt.step()                                               # :84
        return closs.detach(), gloss.detach()


class GANStep:
    """train_gan.py:58-86."""

    def __init__(self, generator, discriminator, world_size=1):
        self.gen, self.dis = generator, discriminator
        self.gopt = FlatOptimizer(generator.parameters(), 'adam', 0.001, world_size=world_size)
        self.dopt = FlatOptimizer(discriminator.parameters(), 'adam', 0.00001, world_size=world_size)

    def __call__(self, real, z_gen, z_dis):
        bce = torch.nn.functional.binary_cross_entropy
        gen, dis = self.gen, self.dis
        b = real.shape[0]
        self.gopt.zero_grad(); self.dopt.zero_grad()
        # :61-65.  The reference's backward here also fills the discriminator's weight gradients, which :74 zeroes unread: same
        # update without that dead work (the discriminator parameters do not require grad during this backward).
        for q in self.dopt.params:
            q.requires_grad_(False)
        try:
            gloss = -torch.mean(torch.log(dis(gen(z_gen))))
            with arena_backward():
                gloss.backward()
        finally:
            for q in self.dopt.params:
                q.requires_grad_(True)
        self.gopt.step()
        self.dopt.zero_grad()
        with torch.no_grad():
            fake = gen(z_dis)
        out_fake = dis(fake)
        floss = bce(out_fake, torch.zeros(b, device=real.device))      # :76-79
        with arena_backward():
            floss.backward()
        self.dopt.step()
        self.dopt.zero_grad()
        vloss = bce(dis(real), torch.ones(b, device=real.device))      # :82-85
        with arena_backward():
            vloss.backward()
        self.dopt.step()
        return gloss.detach(), floss.detach(), vloss.detach()


class AutodecoderStep:
    """train_sdf_autodecoder.py:77-91: the latent table is a raw leaf tensor with its own Adam."""

    def __init__(self, sdf_net, latent_table, lr=1e-5, sigma=0.01, world_size=1, points_per_shape=0):
        """points_per_shape = P > 0: every batch is laid out shape by shape, P consecutive points each (shape_index[i] == i // P)"""
        self.net, self.sigma, self.pps = sdf_net, sigma, points_per_shape
        self.table = latent_table.detach().clone().requires_grad_(True)
        self.nopt = FlatOptimizer(sdf_net.parameters(), 'adam', lr, world_size=world_size)
        self.lopt = FlatOptimizer([self.table], 'adam', lr, world_size=world_size)

    def __call__(self, points, sdf, shape_index):
        """points [N,3], sdf [N], shape_index int32 [N] (= point_index // POINTCLOUD_SIZE, :78 with the D6 fix)."""
        self.nopt.zero_grad(); self.lopt.zero_grad()                    # :84-86
        out = self.net(points, self.table, shape_index, points_per_shape=self.pps)      # :80,87 without materialising table[index]
        # mean(z_batch^2) over the gathered rows == sum_s count_s*|table_s|^2 / (N*L): [S,L] math instead of [N,L].  The counts are
        # recomputed from the index tensor on every call (no host sync, graph-capturable): a (data_ptr, _version, numel) key is not
        # a tensor identity -- the caching allocator recycles addresses across the fresh index tensors of a data loader.
        if self.pps:
            counts = torch.full((self.table.shape[0],), float(self.pps), dtype=torch.float32, device=points.device)
        else:
            counts = torch.zeros(self.table.shape[0], dtype=torch.float32, device=points.device)
            counts.index_add_(0, shape_index.long(), torch.ones((), dtype=torch.float32, device=points.device).expand(shape_index.shape[0]))
        reg = (counts.unsqueeze(1) * torch.pow(self.table, 2)).sum() / (points.shape[0] * self.table.shape[1])
        loss = torch.mean(torch.abs(out - sdf)) + self.sigma * reg     # :88
        with arena_backward():                                                # :89
            loss.backward()
        self.nopt.step(); self.lopt.step()                              # :90-91
        return loss.detach()


class HybridProgressiveStep:
    """train_hybrid_progressive_gan.py:134-166: SDFNet generator evaluated on the R^3 grid of every sample (latents broadcast by
    shape index instead of `.repeat`, :92), progressive discriminator, WGAN-GP, RMSprop 1e-4 for both (:81-82).
    `generator_update` is the every-5th-batch branch (:136-146); `discriminator_update` is :153-166."""

    def __init__(self, generator, discriminator, iteration, lr=0.0001, gp_weight=10.0, world_size=1):
        from .nn.progressive_gan import RESOLUTIONS
        from .nn.sdf_net import get_voxel_coordinates
        self.gen, self.dis, self.gp_weight = generator, discriminator, gp_weight
        discriminator.set_iteration(iteration)
        self.r = RESOLUTIONS[iteration]
        dev = next(generator.parameters()).device
        self.grid = get_voxel_coordinates(self.r, return_torch_tensor=True, device=dev)     # :95
        self.gopt = FlatOptimizer(generator.parameters(), 'rmsprop', lr, world_size=world_size)
        self.dopt = FlatOptimizer(discriminator.parameters(), 'rmsprop', lr, world_size=world_size)
        self.critic_update = CriticUpdate(discriminator)
        self._cache = {}

    def generate(self, z):
        b, g = z.shape[0], self.grid.shape[0]
        if b not in self._cache:
            self._cache[b] = (self.grid.repeat((b, 1)), torch.arange(b, device=z.device, dtype=torch.int32).repeat_interleave(g))
        pts, idx = self._cache[b]
        return self.gen(pts, z, idx, points_per_shape=g).reshape(-1, self.r, self.r, self.r)      # :139-140

    def generator_update(self, z):
        self.gopt.zero_grad(); self.dopt.zero_grad()
        for q in self.dopt.params:                       # the critic's weight gradients are discarded by :158 (zero_grad)
            q.requires_grad_(False)
        try:
            loss = -self.dis(self.generate(z)).mean()                                  # :143-144
            with arena_backward():
                loss.backward()
        finally:
            for q in self.dopt.params:
                q.requires_grad_(True)
        self.gopt.step()                                                               # :146
        return loss.detach()

    def discriminator_update(self, valid, z, alpha):
        self.gopt.zero_grad(); self.dopt.zero_grad()                                    # :153
        with torch.no_grad():                            # the reference back-propagates into G here and discards it (:136)
            fake = self.generate(z)
        b = valid.shape[0]
        if _fused_critic_enabled() and self.critic_update.supported():
            with torch.no_grad():                        # :157-164 as the hand-scheduled sweep of shapegan_b200/critic.py
                out4 = self.critic_update(fake, valid, alpha, self.gp_weight)
            self.dopt.step()                                                           # :166
            return out4[0], out4[1]
        # one batch of 2B samples (no BatchNorm in the critic: same sums, half the launches)
        out = self.dis(torch.cat((fake, valid), 0))                                    # :157,160
        out_fake, out_valid = out[:b], out[b:]
        gp = gradient_penalty(self.dis, valid, fake, alpha, self.gp_weight)            # :162
        loss = out_fake.mean() - out_valid.mean() + gp                                 # :163
        with arena_backward():
            loss.backward()
        self.dopt.step()                                                               # :166
        return loss.detach(), gp.detach()


class VAEStep:
    """train_autoencoder.py:98-117."""

    def __init__(self, autoencoder, world_size=1):
        self.m = autoencoder
        self.opt = FlatOptimizer(autoencoder.parameters(), 'adam', 0.00005, world_size=world_size)

    def __call__(self, batch):
        m = self.m
        self.opt.zero_grad()
        m.train()
        if m.is_variational:
            out, mean, logvar = m(batch)
            kld = -0.5 * torch.sum(1 + logvar - mean.pow(2) - logvar.exp()) / mean.nelement()      # :54-55
        else:
            out, kld = m(batch), 0
        diff = out - batch
        diff = torch.where(batch < 0, diff * 32, diff)                  # :57-62
        loss = torch.mean(torch.abs(diff)) + kld
        with arena_backward():
            loss.backward()
        self.opt.step()
        return loss.detach()
