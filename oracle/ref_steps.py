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
