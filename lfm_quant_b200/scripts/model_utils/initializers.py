"""Initializer factory (reference: scripts/model_utils/initializers.py:14-24), NumPy draws on the host.

RandomUniform(-init_scale, init_scale, seed) when use_custom_init (the default), else Glorot{Normal,Uniform}(seed)
for the LSTM ``kernel``; Keras defaults for everything else (orthogonal recurrent kernel, unit forget bias,
gamma=1 / beta=0, Glorot-uniform Dense, zero Dense bias).  TensorFlow's random streams cannot be reproduced
(SURVEY App. B #7); the distribution families are.
"""
from __future__ import absolute_import, division, print_function

import numpy as np


class Initializer(object):

    def __init__(self, config):
        self.config = config

    def get_initializer(self):
        """Returns ``f(shape, rng) -> ndarray`` for the LSTM kernel."""
        cfg = self.config
        if cfg.use_custom_init:
            s = cfg.init_scale
            return lambda shape, rng: rng.uniform(-s, s, size=shape)
        if cfg.initializer == 'GlorotNormal':
            def glorot_normal(shape, rng):
                std = np.sqrt(2.0 / (shape[0] + shape[1])) / .87962566103423978   # truncated normal, keras scaling
                v = rng.normal(size=shape)
                bad = np.abs(v) > 2
                while bad.any():
                    v[bad] = rng.normal(size=int(bad.sum()))
                    bad = np.abs(v) > 2
                return v * std
            return glorot_normal
        if cfg.initializer == 'GlorotUniform':
            return lambda shape, rng: rng.uniform(-np.sqrt(6.0 / sum(shape)), np.sqrt(6.0 / sum(shape)), size=shape)
        raise NotImplementedError

    def initial_weights(self, specs):
        """specs: [(name, shape)] of the trainable variables in Keras order -> list of fp32 arrays."""
        rng = np.random.RandomState(self.config.seed)
        kernel_init = self.get_initializer()
        out = []
        for name, shape in specs:
            leaf = name.split('/')[-1]
            if name.startswith(('lstm', 'gru')) and leaf == 'kernel':
                w = kernel_init(shape, rng)
            elif leaf == 'recurrent_kernel':            # keras Orthogonal on [H, 4H] (GRU: [H, 3H])
                a = rng.normal(size=(shape[1], shape[0]))
                q, r = np.linalg.qr(a)
                w = (q * np.sign(np.diag(r))).T
            elif name.startswith('lstm') and leaf == 'bias':
                H = shape[0] // 4
                w = np.zeros(shape)
                w[H:2 * H] = 1.0                        # unit_forget_bias
            elif leaf == 'gamma':
                w = np.ones(shape)
            elif leaf in ('beta', 'bias'):
                w = np.zeros(shape)
            elif leaf == 'kernel':                      # Dense: glorot_uniform
                lim = np.sqrt(6.0 / (shape[0] + shape[1]))
                w = rng.uniform(-lim, lim, size=shape)
            else:
                raise KeyError(name)
            out.append(np.asarray(w, dtype=np.float32))
        return out
