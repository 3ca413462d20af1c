"""Optimizer factory (reference: scripts/model_utils/optimizers.py:15-54).

The update rules themselves (Adadelta / Adam / RMSprop / SGD with Keras defaults) run inside the fused native
step (lfmq_apply); this object carries the choice, the learning-rate schedule evaluated on the host, and the
iteration counter -- what ``keras.optimizers.*`` exposes to scripts/train.py:35,198.
"""
from __future__ import absolute_import, division, print_function

import math

SUPPORTED = ('Adam', 'Adadelta', 'RMSprop', 'SGD')


class LearningRateSchedule(object):
    """ExponentialDecay(staircase) / PolynomialDecay / PiecewiseConstantDecay (optimizers.py:37-49)."""

    def __init__(self, config):
        self.c = config
        if config.lr_schedule not in ('ExponentialDecay', 'PolynomialDecay', 'PiecewiseConstantDecay'):
            print("Invalid learning rate scheduler specified")
            raise ValueError

    def __call__(self, step):
        c = self.c
        if c.lr_schedule == 'ExponentialDecay':
            return c.learning_rate * c.lr_decay ** math.floor(step / c.decay_steps)
        if c.lr_schedule == 'PolynomialDecay':
            s = min(step, c.decay_steps)
            return (c.learning_rate - c.end_learning_rate) * (1 - s / c.decay_steps) ** c.decay_power + \
                c.end_learning_rate
        for bnd, v in zip(c.piecewise_lr_boundaries, c.piecewise_lr_values):
            if step <= bnd:
                return v
        return c.piecewise_lr_values[len(c.piecewise_lr_boundaries)]


class NativeOptimizer(object):
    """What train.py needs from a keras optimizer: ``iterations``, ``learning_rate`` and ``apply_gradients``."""

    def __init__(self, name, schedule, sgd_momentum=0.0):
        self.name, self.learning_rate, self.momentum = name, schedule, sgd_momentum
        self.iterations = 0
        self._model = None

    def bind(self, model):
        self._model = model

    def current_lr(self):
        return float(self.learning_rate(self.iterations))

    def apply_gradients(self, grads_and_vars=None):
        """Clip + update + MaxNorm on the gradients lfmq_backward left on the device (train.py:195-198)."""
        assert self._model is not None, 'optimizer is not bound to a model'
        self._model.engine.apply(self.current_lr(), self.iterations)
        self.iterations += 1


class Optimizers(object):

    def __init__(self, config):
        self.config = config
        self.optimizer = self.config.optimizer
        self.lr_decay = self.config.lr_decay
        self.learning_rate = self.get_learning_rate()

    def get_optimizer(self):
        if self.optimizer not in SUPPORTED:
            raise ValueError("%s optimizer not found in tf.keras.optimizers" % self.optimizer)
        return NativeOptimizer(self.optimizer, self.learning_rate, self.config.sgd_momentum)

    def get_learning_rate(self):
        return LearningRateSchedule(self.config)
