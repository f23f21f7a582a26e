"""ORACLE — test infrastructure, NOT product code.

Functional fp32 restatement of the reference's hot path (marian42/shapegan) on
top of plain ``torch.nn.functional`` CPU ops.  The reference's arithmetic lives
in third-party torch (version unpinned by the reference); this file restates
what the reference's ``nn.Sequential`` stacks compute, operating directly on
``state_dict`` tensors so that it is independent of the product's module
classes.  It is pinned against golden vectors produced by importing the real
reference modules from /root/reference (see ``oracle/gen_golden.py``; fixtures
under ``tests/golden``).  The reference has no tests of its own, so those
fixtures are the only pin there is: "reference modules x torch 2.11 CPU fp32".

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.

Every function cites the reference file:line it follows.
"""
import math

import torch
import torch.nn.functional as F

LATENT_CODE_SIZE = 128          # model/__init__.py:10
LRELU_SLOPE = 0.2               # model/gan.py:11 etc.
BN_EPS = 1e-5                   # torch.nn.BatchNorm default used at model/gan.py:10
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------
# DeepSDF MLP   (model/sdf_net.py:23-61)
# ----------------------------------------------------------------------------
def sdfnet_forward(sd, points, latent_codes):
    """model/sdf_net.py:56-61.  points [N,3], latent_codes [N,L] -> [N] (squeezed)."""
    x_in = torch.cat((points, latent_codes), dim=1)                    # :57
    x = x_in
    for i in (0, 2, 4, 6):                                             # layers1, :26-38
        x = F.relu(F.linear(x, sd['layers1.%d.weight' % i], sd['layers1.%d.bias' % i]))
    x = torch.cat((x, x_in), dim=1)                                    # :59
    for i in (0, 2, 4):                                                # layers2, :40-48
        x = F.relu(F.linear(x, sd['layers2.%d.weight' % i], sd['layers2.%d.bias' % i]))
    x = torch.tanh(F.linear(x, sd['layers2.6.weight'], sd['layers2.6.bias']))   # :50-51
    return x.squeeze()                                                 # :61


def sdfnet_autodecoder_loss(sd, points, latent_table, shape_index, target_sdf, sigma=0.01):
    """train_sdf_autodecoder.py:80-88 (with `//` at :78)."""
    z = latent_table[shape_index, :]                                   # :80
    out = sdfnet_forward(sd, points, z)                                # :87
    return torch.mean(torch.abs(out - target_sdf)) + sigma * torch.mean(torch.pow(z, 2))   # :88


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def _bn(x, sd, prefix, training, stats_out=None):
    """torch BatchNorm{1,3}d semantics used at model/gan.py:10,14,18 and
    model/autoencoder.py:17..60: biased variance to normalise, unbiased for
    running_var, momentum 0.1, eps 1e-5.  Functional: returns y and (optionally)
    the updated running stats instead of mutating sd."""
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    if training:
        dims = [0] + list(range(2, x.dim()))
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        if stats_out is not None:
            unbiased = var * (n / max(n - 1, 1))
            stats_out[prefix + '.running_mean'] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
            stats_out[prefix + '.running_var'] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * unbiased.detach()
            stats_out[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
    else:
        mean, var = rm, rv
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - mean.reshape(shape)) / torch.sqrt(var.reshape(shape) + BN_EPS) * w.reshape(shape) + b.reshape(shape)


def _lrelu(x):
    return F.leaky_relu(x, LRELU_SLOPE)


# ----------------------------------------------------------------------------
# Voxel GAN   (model/gan.py)
# ----------------------------------------------------------------------------
def generator_forward(sd, z, training=True, stats_out=None):
    """model/gan.py:8-29.  z [B,128] -> [B,1,32,32,32]."""
    x = z.reshape((-1, LATENT_CODE_SIZE, 1, 1, 1))                                        # :28
    x = F.conv_transpose3d(x, sd['layers.0.weight'], sd['layers.0.bias'], stride=1)      # :9
    x = _lrelu(_bn(x, sd, 'layers.1', training, stats_out))                              # :10-11
    x = F.conv_transpose3d(x, sd['layers.3.weight'], sd['layers.3.bias'], stride=2, padding=1)   # :13
    x = _lrelu(_bn(x, sd, 'layers.4', training, stats_out))
    x = F.conv_transpose3d(x, sd['layers.6.weight'], sd['layers.6.bias'], stride=2, padding=1)   # :17
    x = _lrelu(_bn(x, sd, 'layers.7', training, stats_out))
    x = F.conv_transpose3d(x, sd['layers.9.weight'], sd['layers.9.bias'], stride=2, padding=1)   # :21
    return torch.tanh(x)                                                                  # :22


def discriminator_forward(sd, x, use_sigmoid=True):
    """model/gan.py:48-65.  x [B,32,32,32] or [B,1,32,32,32] -> [B] (squeezed)."""
    if x.dim() < 5:
        x = x.unsqueeze(1)                                                                # :62-63
    x = _lrelu(F.conv3d(x, sd['layers.0.weight'], sd['layers.0.bias'], stride=2, padding=1))
    x = _lrelu(F.conv3d(x, sd['layers.2.weight'], sd['layers.2.bias'], stride=2, padding=1))
    x = _lrelu(F.conv3d(x, sd['layers.4.weight'], sd['layers.4.bias'], stride=2, padding=1))
    x = F.conv3d(x, sd['layers.6.weight'], sd['layers.6.bias'], stride=1)                # :55
    if use_sigmoid:
        x = torch.sigmoid(x)                                                              # :56
    return x.squeeze()                                                                    # :65


def clip_weights(sd, value):
    """model/gan.py:67-69 (functional)."""
    return {k: v.clamp(-value, value) for k, v in sd.items()}


# ----------------------------------------------------------------------------
# Progressive discriminator   (model/progressive_gan.py)
# ----------------------------------------------------------------------------
RESOLUTIONS = [8, 16, 32, 64]                 # progressive_gan.py:4
FEATURE_COUNTS = [128, 64, 32, 1]             # :5
FINAL_LAYER_FEATURES = 256                    # :6


def from_sdf(x, iteration):
    """model/progressive_gan.py:9-16: zero-pad the single SDF channel."""
    r = RESOLUTIONS[iteration]
    c = FEATURE_COUNTS[iteration]
    x = x.reshape((-1, 1, r, r, r))
    pad = torch.zeros((x.shape[0], c - 1, r, r, r), dtype=x.dtype)
    return torch.cat((x, pad), dim=1)


def progressive_discriminator_forward(sd, x, iteration, fade_in_progress=1.0):
    """model/progressive_gan.py:44-57."""
    def block(i, t):
        return _lrelu(F.conv3d(t, sd['optional_layers.%d.0.weight' % i], sd['optional_layers.%d.0.bias' % i],
                               stride=2, padding=1))                                      # :37-39
    x_in = x
    x = block(iteration, from_sdf(x, iteration))                                          # :46-47
    if fade_in_progress < 1.0 and iteration > 0:                                          # :48
        x2 = from_sdf(x_in[:, ::2, ::2, ::2], iteration - 1)                              # :49
        x = fade_in_progress * x + (1.0 - fade_in_progress) * x2                          # :50
    i = iteration - 1
    while i >= 0:                                                                         # :52-55
        x = block(i, x)
        i -= 1
    x = x.reshape(-1, 64 * FINAL_LAYER_FEATURES)                                          # :27
    x = _lrelu(F.linear(x, sd['head.1.weight'], sd['head.1.bias']))                       # :28-29
    x = F.linear(x, sd['head.3.weight'], sd['head.3.bias'])                               # :30
    return x.squeeze()                                                                    # :57


def gradient_penalty(disc_fn, real, fake, alpha, weight=10.0):
    """train_hybrid_progressive_gan.py:102-111 with alpha injected ([B,1,1,1])."""
    a = alpha.expand(real.shape)                                                          # :103
    xi = (a * real + (1 - a) * fake).detach().requires_grad_(True)                       # :105-106
    out = disc_fn(xi)                                                                     # :108
    grads = torch.autograd.grad(outputs=out, inputs=xi, grad_outputs=torch.ones(out.shape),
                                create_graph=True, retain_graph=True, only_inputs=True)[0]   # :110
    return ((grads.norm(2, dim=(1, 2, 3)) - 1) ** 2).mean() * weight                      # :111


# ----------------------------------------------------------------------------
# Voxel (V)AE   (model/autoencoder.py)
# ----------------------------------------------------------------------------
def autoencoder_encode(sd, x, is_variational=True, training=True, eps=None, stats_out=None):
    """model/autoencoder.py:67-89.  Returns (z, mean, log_variance) for the VAE, z otherwise."""
    x = x.reshape((-1, 1, 32, 32, 32))                                                    # :68
    for conv, bn, stride, pad in ((0, 1, 2, 1), (3, 4, 2, 1), (6, 7, 2, 1), (9, 10, 1, 0)):   # :16-30
        x = F.conv3d(x, sd['encoder.%d.weight' % conv], sd['encoder.%d.bias' % conv], stride=stride, padding=pad)
        x = _lrelu(_bn(x, sd, 'encoder.%d' % bn, training, stats_out))
    x = x.reshape(x.shape[0], -1)                                                         # :32
    x = F.linear(x, sd['encoder.13.weight'], sd['encoder.13.bias'])                       # :34
    if not is_variational:
        return x
    x = _lrelu(_bn(x, sd, 'encoder.vae-bn', training, stats_out))                         # :38-39
    mean = F.linear(x, sd['encode_mean.weight'], sd['encode_mean.bias']).squeeze()        # :74
    log_variance = F.linear(x, sd['encode_log_variance.weight'], sd['encode_log_variance.bias']).squeeze()   # :77
    if training:
        z = mean + torch.exp(log_variance * 0.5) * eps                                    # :78-82
    else:
        z = mean
    return z, mean, log_variance


def autoencoder_decode(sd, z, training=True, stats_out=None):
    """model/autoencoder.py:91-95."""
    if z.dim() == 1:
        z = z.unsqueeze(0)
    x = F.linear(z, sd['decoder.0.weight'], sd['decoder.0.bias'])                         # :45
    x = _lrelu(_bn(x, sd, 'decoder.1', training, stats_out))                              # :46-47
    x = x.reshape(-1, LATENT_CODE_SIZE * 2, 1, 1, 1)                                      # :49
    x = F.conv_transpose3d(x, sd['decoder.4.weight'], sd['decoder.4.bias'], stride=1)    # :51
    x = _lrelu(_bn(x, sd, 'decoder.5', training, stats_out))
    x = F.conv_transpose3d(x, sd['decoder.7.weight'], sd['decoder.7.bias'], stride=2, padding=1)
    x = _lrelu(_bn(x, sd, 'decoder.8', training, stats_out))
    x = F.conv_transpose3d(x, sd['decoder.10.weight'], sd['decoder.10.bias'], stride=2, padding=1)
    x = _lrelu(_bn(x, sd, 'decoder.11', training, stats_out))
    x = F.conv_transpose3d(x, sd['decoder.13.weight'], sd['decoder.13.bias'], stride=2, padding=1)   # :63
    return x.squeeze()                                                                    # :95


def autoencoder_forward(sd, x, is_variational=True, training=True, eps=None, stats_out=None):
    """model/autoencoder.py:97-104."""
    if not is_variational:
        return autoencoder_decode(sd, autoencoder_encode(sd, x, False, training, None, stats_out), training, stats_out)
    z, mean, log_variance = autoencoder_encode(sd, x, True, training, eps, stats_out)
    return autoencoder_decode(sd, z, training, stats_out), mean, log_variance


def kld_loss(mean, log_variance):
    """train_autoencoder.py:54-55."""
    return -0.5 * torch.sum(1 + log_variance - mean.pow(2) - log_variance.exp()) / mean.nelement()


def reconstruction_loss(output, target):
    """train_autoencoder.py:57-62 (sign-weighted L1; functional form of the in-place multiply)."""
    difference = output - target
    difference = torch.where(target < 0, difference * 32, difference)
    return torch.mean(torch.abs(difference))


# ----------------------------------------------------------------------------
# grid indexing  (util.py:60-74) -- bit-exact spot
# ----------------------------------------------------------------------------
def voxel_coordinates(resolution, size=1.0):
    """util.py:60-74: row index i*R^2 + j*R + k <-> (x_i, y_j, z_k), z fastest;
    inclusive linspace computed in float64 then cast to float32."""
    import numpy as np
    lin = np.linspace(-size, size, resolution)
    pts = np.stack(np.meshgrid(lin, lin, lin))
    pts = np.swapaxes(pts, 1, 2).reshape(3, -1).transpose()
    return torch.from_numpy(pts.astype(np.float32))


# ----------------------------------------------------------------------------
# deterministic synthetic weights shared by gen_golden.py, tests and bench
# ----------------------------------------------------------------------------
def seeded_state_dict(shapes, seed, scale=None):
    """Build a state_dict from {name: shape} with a CPU generator.  Weight tensors ~ U(-b, b)
    with b = 1/sqrt(fan_in) (same bound torch's default init uses), BN weights ~ U(0.5,1.5),
    running_var ~ U(0.5,1.5), everything reproducible across machines for a fixed torch build."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        if name.endswith('num_batches_tracked'):
            sd[name] = torch.tensor(3, dtype=torch.int64)
            continue
        shape = tuple(shape)
        if name.endswith('running_var') or (len(shape) == 1 and name.endswith('.weight')
                                            and scale is None and _is_bn(name, shapes)):
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            b = (scale if scale is not None else 1.0) / math.sqrt(fan_in)
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * b
    return sd


def _is_bn(name, shapes):
    return name[:-len('weight')] + 'running_mean' in shapes
