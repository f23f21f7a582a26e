"""ORACLE — test infrastructure.  state_dict key/shape tables of the reference's modules, in the
reference's state_dict() ORDER (seeded_state_dict draws in this order).  Pinned by tests/test_oracle_golden.py against the
golden vectors generated from the unmodified reference."""
from . import ref_torch as R

SDF_SHAPES = {}
for _p, _dims in (('layers1.0', (256, 131)), ('layers1.2', (256, 256)), ('layers1.4', (256, 256)), ('layers1.6', (256, 256)),
                  ('layers2.0', (256, 387)), ('layers2.2', (256, 256)), ('layers2.4', (256, 256)), ('layers2.6', (1, 256))):
    SDF_SHAPES[_p + '.weight'] = _dims
    SDF_SHAPES[_p + '.bias'] = (_dims[0],)


def sdf_shapes(latent=128):
    s = dict(SDF_SHAPES)
    s['layers1.0.weight'] = (256, 3 + latent)
    s['layers2.0.weight'] = (256, 256 + latent + 3)
    return s


def gen_shapes():
    s = {}
    for i, (cin, cout) in zip((0, 3, 6, 9), ((128, 256), (256, 128), (128, 64), (64, 1))):
        s['layers.%d.weight' % i] = (cin, cout, 4, 4, 4)
        s['layers.%d.bias' % i] = (cout,)
        if i != 9:
            for n in ('weight', 'bias', 'running_mean', 'running_var'):
                s['layers.%d.%s' % (i + 1, n)] = (cout,)
            s['layers.%d.num_batches_tracked' % (i + 1)] = ()
    return s


def disc_shapes():
    s = {}
    for i, (cin, cout) in zip((0, 2, 4, 6), ((1, 64), (64, 128), (128, 256), (256, 1))):
        s['layers.%d.weight' % i] = (cout, cin, 4, 4, 4)
        s['layers.%d.bias' % i] = (cout,)
    return s


def prog_shapes():
    s = {'head.1.weight': (128, 16384), 'head.1.bias': (128,), 'head.3.weight': (1, 128), 'head.3.bias': (1,)}
    fc = R.FEATURE_COUNTS
    for prefix in ('optional_layers.%d.0', 'optional_layer_%d.0'):
        for i in range(4):
            cout = fc[i - 1] if i > 0 else 256
            s[(prefix % i) + '.weight'] = (cout, fc[i], 4, 4, 4)
            s[(prefix % i) + '.bias'] = (cout,)
    return s


def ae_shapes(variational):
    s = {}

    def bn(prefix, c):
        for n in ('weight', 'bias', 'running_mean', 'running_var'):
            s['%s.%s' % (prefix, n)] = (c,)
        s[prefix + '.num_batches_tracked'] = ()
    for i, (cin, cout) in zip((0, 3, 6, 9), ((1, 24), (24, 48), (48, 96), (96, 256))):
        s['encoder.%d.weight' % i] = (cout, cin, 4, 4, 4)
        s['encoder.%d.bias' % i] = (cout,)
        bn('encoder.%d' % (i + 1), cout)
    s['encoder.13.weight'] = (128, 256)
    s['encoder.13.bias'] = (128,)
    if variational:
        bn('encoder.vae-bn', 128)
        for n in ('encode_mean', 'encode_log_variance'):
            s[n + '.weight'] = (128, 128)
            s[n + '.bias'] = (128,)
    s['decoder.0.weight'] = (256, 128)
    s['decoder.0.bias'] = (256,)
    bn('decoder.1', 256)
    for i, (cin, cout) in zip((4, 7, 10, 13), ((256, 96), (96, 48), (48, 24), (24, 1))):
        s['decoder.%d.weight' % i] = (cin, cout, 4, 4, 4)
        s['decoder.%d.bias' % i] = (cout,)
        if i != 13:
            bn('decoder.%d' % (i + 1), cout)
    return s
