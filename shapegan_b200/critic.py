"""Hand-scheduled critic update: the gradient of

        mean D(fake) - mean D(real)  [ + weight * mean_b (|d D(x^_b) / d x^_b|_2 - 1)^2 ]         x^ = alpha real + (1 - alpha) fake

w.r.t. every critic parameter (train_wgan.py:66-69; train_hybrid_progressive_gan.py:102-111, :157-164), without autograd.

Both discriminators of the reference are piecewise linear -- Conv3d / Linear + LeakyReLU(0.2), a linear head, no BatchNorm
(model/gan.py:48-57, model/progressive_gan.py:26-39) -- so the whole double backward of the gradient penalty is four
sweeps over the SAME layers, and the sweeps batch:

  1 forward   ONE pass over [fake; real; x^] (3B samples): H_l = lrelu(L_l H_{l-1} + b_l), scores s
  2 backward  ONE pass over the same 3B rows, seeds (+1/B, -1/B, 1): U_{l-1} = (L_l^T U_l) * lrelu'(H_{l-1}), the activation
              backward fused into the GEMM epilogue.  The x^ rows of U_0 are g = d D(x^) / d x^.
  3 penalty   n_b = |g_b|, gp = weight mean (n_b - 1)^2, v_b = d gp / d g_b -- one kernel; v overwrites x^ in the input buffer.
  4 adjoint   over the B penalty rows only: C_l = (L_l C_{l-1}) * lrelu'(H_l), C_0 = v, written IN PLACE over the x^ rows of H_l
              (the mask is read by the same thread that overwrites it).
  5 weights   ONE weight-gradient GEMM per layer over all 3B rows: rows 0..2B pair (H_{l-1}, U_l) = the Wasserstein term, rows
              2B..3B pair (C_{l-1}, U_l) = d gp / d W_l = wgrad(C_{l-1}, U_l) -- the same contraction, so they share the launch.

12 GEMM launches for the whole critic update of gan.Discriminator (3 + 3 + 3 + 3) instead of 5 autograd passes of separate
launches; the non-GEMM work is three tiny kernels (csrc/sg_critic.cu).  Gradients accumulate into the parameters' `.grad`
(the flat gradient arena of train.FlatOptimizer), biases get exactly zero from the penalty."""
import ctypes

import torch

from . import _lib as L
from . import ops, raw

LRELU = L.ACT_LRELU


def _call(name, *args):
    L.check(getattr(L.lib(), name)(*args, raw.stream()), name)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Layer:
    """one LinOp + LeakyReLU stage: op, weight, bias, and how its input/output tensors are shaped"""

    def __init__(self, op, weight, bias, kind, out_shape):
        self.op, self.w, self.b, self.kind, self.out_shape = op, weight, bias, kind, out_shape


def plan_for(critic):
    """(resolution, layers, head) of a supported critic, or None (fade-in blend of the progressive critic: autograd path)"""
    from .nn import gan as G
    from .nn import progressive_gan as PG
    if isinstance(critic, G.Discriminator):
        if critic.use_sigmoid:
            return None
        l = critic.layers
        c = G._D_CHANNELS
        layers = [_Layer(critic._op0, l[0].weight, l[0].bias, 'patch', lambda n: (n, 16, 16, 16, c[1])),
                  _Layer(critic._op1, l[2].weight, l[2].bias, 'conv', lambda n: (n, 8, 8, 8, c[2])),
                  _Layer(critic._op2, l[4].weight, l[4].bias, 'conv', lambda n: (n, 4, 4, 4, c[3]))]
        head = dict(w=l[6].weight, b=l[6].bias, c=64 * c[3], wc=c[3], s_t=1, s_c=64)
        return 32, layers, head
    if isinstance(critic, PG.Discriminator):
        it = critic.iteration
        if critic.fade_in_progress < 1.0 and it > 0:
            return None
        r = PG.RESOLUTIONS[it]
        layers = []
        conv = critic.optional_layers[it][0]
        cout = conv.weight.shape[0]
        res = r // 2
        layers.append(_Layer(critic._ops_first[it], conv.weight, conv.bias, 'patch', (lambda rr, cc: lambda n: (n, rr, rr, rr, cc))(res, cout)))
        for i in range(it - 1, -1, -1):
            conv = critic.optional_layers[i][0]
            res //= 2
            layers.append(_Layer(critic._ops_inner[i], conv.weight, conv.bias, 'conv', (lambda rr, cc: lambda n: (n, rr, rr, rr, cc))(res, conv.weight.shape[0])))
        layers.append(_Layer(critic._op_head, critic.head[1].weight, critic.head[1].bias, 'dense', lambda n: (n, 128)))
        head = dict(w=critic.head[3].weight, b=critic.head[3].bias, c=128, wc=0, s_t=0, s_c=1)
        return r, layers, head
    return None


class CriticUpdate:
    """`CriticUpdate(critic)(fake, real, alpha=None, gp_weight=10.0)` accumulates the critic-loss gradient into the parameters'
    `.grad` and returns a device tensor [4] = (loss, gp, mean D(fake), mean D(real)).  alpha None = no gradient penalty."""

    def __init__(self, critic):
        self.critic = critic
        self._seeds = {}

    def supported(self):
        return plan_for(self.critic) is not None

    def _seed_vector(self, b, gp, dev):
        key = (b, gp, str(dev))
        if key not in self._seeds:
            s = [1.0 / b] * b + [-1.0 / b] * b + ([1.0] * b if gp else [])
            self._seeds[key] = torch.tensor(s, dtype=torch.float32, device=dev)
        return self._seeds[key]

    def __call__(self, fake, real, alpha=None, gp_weight=10.0):
        plan = plan_for(self.critic)
        assert plan is not None, 'CriticUpdate: unsupported critic configuration'
        r, layers, head = plan
        gp = alpha is not None
        b = real.shape[0]
        dev = real.device
        planes = ops._planes()
        m = r * r * r
        n = (3 if gp else 2) * b
        # ---- input buffer [fake; real; x^]
        x0 = torch.empty((n, r, r, r), dtype=torch.float32, device=dev)
        x0[:b].copy_(fake.reshape(b, r, r, r))
        x0[b:2 * b].copy_(real.reshape(b, r, r, r))
        if gp:
            _call('sg_gp_interp', _ptr(x0[b:2 * b]), _ptr(x0[:b]), _ptr(alpha.reshape(-1).contiguous()), _ptr(x0[2 * b:]), b, m)
        # ---- 1 forward over all rows
        hs, h = [], x0
        for ly in layers:
            if ly.kind == 'dense':
                h = h.reshape(h.shape[0], n, -1)
            h = ly.op.fwd(h, ly.w, ly.b, LRELU)
            hs.append(h)
        hf = h.reshape(h.shape[0], n, head['c'])
        scores = raw.rowdot_fwd(hf, head['c'], head['w'].detach(), head['b'].detach(), L.ACT_NONE, head['wc'], head['s_t'], head['s_c'])
        # ---- 2 backward over all rows; head first (seed * w, masked by the last activation)
        seeds = self._seed_vector(b, gp, dev)
        u, _ = raw.rowdot_bwd(seeds, scores, L.ACT_NONE, hf, head['c'], head['w'].detach(), True, False, planes, n, head['wc'], head['s_t'], head['s_c'],
                              x_mask_act=LRELU)
        us = [None] * len(layers)
        for i in range(len(layers) - 1, -1, -1):
            ly = layers[i]
            us[i] = u.reshape(hs[i].shape)
            if i == 0:
                break
            g_in = us[i].reshape(us[i].shape[0], n, -1) if ly.kind == 'dense' else us[i]
            u = ly.op.tr(g_in, ly.w, mask=hs[i - 1].reshape(hs[i - 1].shape[0], n, -1) if ly.kind == 'dense' else hs[i - 1], mask_act=LRELU)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        gp_sum = None
        if gp:
            # ---- 3 penalty: g = input gradient of the x^ rows only (first-layer transpose + col2im), v = d gp / d g into x0[x^ rows]
            l0 = layers[0]
            g = l0.op.tr(us[0][:, 2 * b:], l0.w)          # a batch slice of a plane tensor: the kernels address planes by stride
            gp_sum = raw.dsums(1, dev)
            _call('sg_gp_seed', _ptr(g), _ptr(x0[2 * b:]), b, m, float(gp_weight), _ptr(gp_sum))
            # ---- 4 adjoint sweep over the penalty rows, in place over the x^ rows of H_l
            c = x0[2 * b:]
            for ly, hfull in zip(layers, hs):
                if ly.kind == 'dense':
                    c = c.reshape(c.shape[0], b, -1)
                    tgt = hfull[:, 2 * b:]
                else:
                    tgt = hfull[:, 2 * b:]
                c = ly.op.fwd(c, ly.w, None, L.ACT_NONE, out=tgt, mask=tgt, mask_act=LRELU)
        _call('sg_critic_loss', _ptr(scores), b, _ptr(gp_sum), _ptr(out4))
        # ---- 5 weight gradients: one GEMM per layer over all rows; bias sums over the Wasserstein rows only
        x_in = x0
        for i, ly in enumerate(layers):
            gi = us[i]
            xi = x_in
            if ly.kind == 'dense':
                gi = gi.reshape(gi.shape[0], n, -1)
                xi = xi.reshape(xi.shape[0], n, -1)
            if ly.w.grad is None:
                ly.w.grad = torch.zeros_like(ly.w)
            ly.op.wgrad(xi, gi, None, into=ly.w.grad)
            if ly.b is not None:
                if ly.b.grad is None:
                    ly.b.grad = torch.zeros_like(ly.b)
                cch = ly.b.shape[0]
                wrows = us[i][:, :2 * b]
                _, sums = raw.act_bwd(wrows, None, L.ACT_NONE, cch, want_sums=True, want_g=False)
                raw.emit_sums(sums, ly.b.grad, cch, accumulate=True)
            x_in = hs[i]
        # head: dW = sum_r seed'_r * H'_r (the x^ rows of H now hold the adjoint C, their seed is 1); its bias gets sum(seeds) = 0
        hw = head['w']
        if hw.grad is None:
            hw.grad = torch.zeros_like(hw)
        if head['b'] is not None and head['b'].grad is None:
            head['b'].grad = torch.zeros_like(head['b'])          # d loss / d b_head = sum of the seeds of the Wasserstein rows = 0
        _, sums = raw.rowdot_bwd(seeds, scores, L.ACT_NONE, hf, head['c'], hw.detach(), False, True, planes, n, head['wc'], head['s_t'], head['s_c'])
        raw.emit_sums(sums, hw.grad, head['c'], accumulate=True, wc=head['wc'] if head['wc'] > 0 else 0, s_t=head['s_t'], s_c=head['s_c'])
        return out4
