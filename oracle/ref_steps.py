"""ORACLE — test/benchmark infrastructure, NOT product code.

CPU fp32 restatement of the reference's training-step bodies on top of oracle/ref_torch.py (plain torch CPU ops and
torch.optim, exactly what the reference's scripts execute on a CPU host).  Used as the checker in tests and as the
timed `cpu_baseline` / `--impl reference` arm of bench.py (kind = "port": /root/reference itself cannot travel to the
GPU box).  Every function cites the script lines it follows."""
import torch

from . import ref_torch as R


def make_params(shapes, seed):
    sd = R.seeded_state_dict(shapes, seed)
    for k, v in sd.items():
        if v.dtype.is_floating_point and 'running' not in k:
            v.requires_grad_(True)
    return sd


def trainable(sd):
    return [v for k, v in sd.items() if v.requires_grad]


class WGANStepCPU:
    """train_wgan.py:62-71 + :75-84 (RMSprop 5e-5, clip 0.01); gp=True adds train_hybrid_progressive_gan.py:102-111."""

    def __init__(self, gen_sd, cri_sd, gp=False):
        self.g, self.c, self.gp = gen_sd, cri_sd, gp
        self.gopt = torch.optim.RMSprop(trainable(gen_sd), lr=0.00005)
        self.copt = torch.optim.RMSprop(trainable(cri_sd), lr=0.00005)

    def _zero(self):
        for p in trainable(self.g) + trainable(self.c):
            p.grad = None

    def __call__(self, real, z_critic, z_gen, alpha=None):
        stats = {}
        self._zero()
        fake = R.generator_forward(self.g, z_critic, True, stats).detach()
        for k, v in stats.items():
            self.g[k] = v
        closs = torch.mean(R.discriminator_forward(self.c, fake, False)) - torch.mean(R.discriminator_forward(self.c, real, False))
        if self.gp:
            closs = closs + R.gradient_penalty(lambda x: R.discriminator_forward(self.c, x, False), real, fake.squeeze(1), alpha)
        closs.backward()
        self.copt.step()
        if not self.gp:
            with torch.no_grad():
                for p in trainable(self.c):
                    p.clamp_(-0.01, 0.01)
        self._zero()
        stats = {}
        gloss = -torch.mean(R.discriminator_forward(self.c, R.generator_forward(self.g, z_gen, True, stats), False))
        for k, v in stats.items():
            self.g[k] = v
        gloss.backward()
        self.gopt.step()
        return closs.detach(), gloss.detach()


class AutodecoderStepCPU:
    """train_sdf_autodecoder.py:77-91 (two Adams lr 1e-5)."""

    def __init__(self, sd, table):
        self.sd = sd
        self.table = table.clone().requires_grad_(True)
        self.nopt = torch.optim.Adam(trainable(sd), lr=1e-5)
        self.lopt = torch.optim.Adam([self.table], lr=1e-5)

    def __call__(self, points, sdf, index):
        for p in trainable(self.sd) + [self.table]:
            p.grad = None
        loss = R.sdfnet_autodecoder_loss(self.sd, points, self.table, index.long(), sdf)
        loss.backward()
        self.nopt.step()
        self.lopt.step()
        return loss.detach()


class GANStepCPU:
    """train_gan.py:58-86 (Adam 1e-3 / 1e-5, BCE)."""

    def __init__(self, gen_sd, dis_sd):
        self.g, self.d = gen_sd, dis_sd
        self.gopt = torch.optim.Adam(trainable(gen_sd), lr=0.001)
        self.dopt = torch.optim.Adam(trainable(dis_sd), lr=0.00001)

    def _zero(self, sd):
        for p in trainable(sd):
            p.grad = None

    def __call__(self, real, z_gen, z_dis):
        bce = torch.nn.functional.binary_cross_entropy
        b = real.shape[0]
        self._zero(self.g); self._zero(self.d)
        stats = {}
        gloss = -torch.mean(torch.log(R.discriminator_forward(self.d, R.generator_forward(self.g, z_gen, True, stats), True)))
        self.g.update(stats)
        gloss.backward(); self.gopt.step()
        self._zero(self.d)
        stats = {}
        fake = R.generator_forward(self.g, z_dis, True, stats).detach()
        self.g.update(stats)
        floss = bce(R.discriminator_forward(self.d, fake, True), torch.zeros(b))
        floss.backward(); self.dopt.step()
        self._zero(self.d)
        vloss = bce(R.discriminator_forward(self.d, real, True), torch.ones(b))
        vloss.backward(); self.dopt.step()
        return gloss.detach(), floss.detach(), vloss.detach()


class VAEStepCPU:
    """train_autoencoder.py:98-117 (Adam 5e-5)."""

    def __init__(self, sd, variational=True):
        self.sd, self.variational = sd, variational
        self.opt = torch.optim.Adam(trainable(sd), lr=0.00005)

    def __call__(self, batch, eps):
        for p in trainable(self.sd):
            p.grad = None
        stats = {}
        if self.variational:
            out, mean, logvar = R.autoencoder_forward(self.sd, batch, True, True, eps, stats)
            loss = R.reconstruction_loss(out, batch) + R.kld_loss(mean, logvar)
        else:
            loss = R.reconstruction_loss(R.autoencoder_forward(self.sd, batch, False, True, None, stats), batch)
        self.sd.update(stats)
        loss.backward()
        self.opt.step()
        return loss.detach()


class HybridProgressiveStepCPU:
    """train_hybrid_progressive_gan.py:134-166: SDFNet generator on the R^3 grid, progressive discriminator, WGAN-GP,
    RMSprop 1e-4 for both."""

    def __init__(self, gen_sd, dis_sd, iteration, fade=1.0):
        self.g, self.d, self.it, self.fade = gen_sd, dis_sd, iteration, fade
        self.r = R.RESOLUTIONS[iteration]
        self.grid = R.voxel_coordinates(self.r)                                       # :95
        keys = ['head.1.weight', 'head.1.bias', 'head.3.weight', 'head.3.bias'] + \
               ['optional_layers.%d.0.%s' % (i, n) for i in range(4) for n in ('weight', 'bias')]
        self.dparams = [dis_sd[k] for k in keys]
        self.gopt = torch.optim.RMSprop(trainable(gen_sd), lr=0.0001)                 # :81
        self.dopt = torch.optim.RMSprop(self.dparams, lr=0.0001)                      # :82

    def _disc(self, x):
        return R.progressive_discriminator_forward(self.d, x, self.it, self.fade)

    def _generate(self, z):
        b, g = z.shape[0], self.grid.shape[0]
        latent = z.repeat((1, 1, g)).reshape(-1, 128)                                  # :92
        return R.sdfnet_forward(self.g, self.grid.repeat((b, 1)), latent).reshape(-1, self.r, self.r, self.r)   # :139-140

    def generator_update(self, z):
        for p in trainable(self.g) + self.dparams:
            p.grad = None
        loss = -self._disc(self._generate(z)).mean()                                   # :143-144
        loss.backward()
        self.gopt.step()                                                               # :146
        return loss.detach()

    def discriminator_update(self, valid, z, alpha):
        for p in trainable(self.g) + self.dparams:
            p.grad = None
        fake = self._generate(z)                                                       # :155-156 (not detached in the reference)
        out_fake, out_valid = self._disc(fake), self._disc(valid)                      # :157,160
        gp = R.gradient_penalty(self._disc, valid.detach(), fake.detach(), alpha)      # :162
        loss = out_fake.mean() - out_valid.mean() + gp                                 # :163
        loss.backward()
        self.dopt.step()                                                               # :166
        return loss.detach(), gp.detach()
