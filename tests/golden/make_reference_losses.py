"""Generates tests/golden/reference_losses.npz by running the UNMODIFIED reference loss code
(/root/reference/scripts/model_utils/losses.py: Losses.weight_adjusted_mse / weight_adjusted_uq_loss) on fixed inputs.

That module is plain tensor algebra over 14 TensorFlow primitives (expand_dims, squared_difference, reduce_mean,
reduce_sum, reduce_all, equal, cast, multiply, divide, constant, convert_to_tensor, math.log, float32, Huber [unused on the
RNN branch]).  TensorFlow is not installable here, so an import shim maps each of them to its NumPy equivalent (float32
arithmetic, like the reference's tensors); the masking, slicing, weighting and denominators that run are the reference's
own lines.  What this pins is therefore the reference's loss LOGIC; the elementwise / reduction primitives are NumPy's.
Only runs where /root/reference exists; the .npz is what travels.

usage: python tests/golden/make_reference_losses.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

CHILD = r'''
import sys, types
import numpy as np

tf = types.ModuleType('tensorflow')
tf.float32 = np.float32
tf.constant = lambda v, dtype=None: np.asarray(v, dtype=dtype)
tf.equal = np.equal
tf.reduce_all = lambda a, axis=None: np.all(a, axis=axis)
tf.cast = lambda a, dtype=None: np.asarray(a).astype(dtype)
tf.multiply = np.multiply
tf.divide = np.divide
tf.expand_dims = np.expand_dims
tf.reduce_mean = lambda a, axis=None: np.mean(a, axis=axis, dtype=np.asarray(a).dtype)
tf.reduce_sum = lambda a, axis=None: np.sum(a, axis=axis, dtype=np.asarray(a).dtype)
tf.math = types.SimpleNamespace(log=np.log)
py = types.ModuleType('tensorflow.python'); fw = types.ModuleType('tensorflow.python.framework')
opsm = types.ModuleType('tensorflow.python.framework.ops'); opsm.convert_to_tensor = lambda a: np.asarray(a)
pops = types.ModuleType('tensorflow.python.ops'); mo = types.ModuleType('tensorflow.python.ops.math_ops')
mo.squared_difference = lambda a, b: np.square(np.subtract(a, b))
mo.cast = lambda a, dtype: np.asarray(a).astype(dtype)
keras = types.ModuleType('tensorflow.keras'); kl = types.ModuleType('tensorflow.keras.losses')
kl.Huber = type('Huber', (), {})
fw.ops = opsm; pops.math_ops = mo; py.framework = fw; py.ops = pops; keras.losses = kl; tf.python = py; tf.keras = keras
for name, mod in (('tensorflow', tf), ('tensorflow.python', py), ('tensorflow.python.framework', fw),
                  ('tensorflow.python.framework.ops', opsm), ('tensorflow.python.ops', pops),
                  ('tensorflow.python.ops.math_ops', mo), ('tensorflow.keras', keras), ('tensorflow.keras.losses', kl)):
    sys.modules[name] = mod

OUT = sys.argv[1]
sys.path.insert(0, '/root/reference/scripts')
sys.path.insert(0, '/root/reference/scripts/model_utils')
sys.argv = ['x']
import losses as ref
import base_config
c = base_config.get_configs()
c.nn_type = 'RNNPointEstimate'
c.forecast_steps = 1
c.forecast_steps_weights = [1.0]
out = {}
np.seterr(divide='ignore', invalid='ignore')

# (1) the fixture in the reference file itself (losses.py:287-310)
yt = np.array([[[0, 0, 0], [0, 0, 0], [4, 5, 6], [7, 8, 9], [1, 2, 3]],
               [[0, 0, 0], [0, 0, 0], [1, 2, 3], [4, 5, 6], [7, 8, 9]]], dtype=np.float32)
yp = np.ones_like(yt)
for tag, p1, p2 in (('a', 1.0, 0.0), ('b', 0.5, 0.7)):
    c.target_lambda, c.rnn_lambda = p1, p2
    l, m = ref.Losses(c, 2).weight_adjusted_mse([yt], [yp])
    out['fix_%s' % tag] = np.array([p1, p2, float(l), float(m)])
out['fix_y_true'], out['fix_y_pred'] = yt, yp

# (2) random inputs with zero-padded steps
rng = np.random.RandomState(7)
B, T, O, tidx = 6, 5, 4, 2
y = rng.normal(size=(B, T, O)).astype(np.float32)
y[0, :2] = 0.0
y[3, 0] = 0.0
p = rng.normal(size=(B, T, O)).astype(np.float32)
v = (np.abs(rng.normal(size=(B, T, O))) + 0.1).astype(np.float32)
c.target_lambda, c.rnn_lambda = 0.5, 0.7
l, m = ref.Losses(c, tidx).weight_adjusted_mse([y], [p])
lv, mv = ref.Losses(c, tidx).weight_adjusted_mse([y], [p], True)
out['pt_y'], out['pt_p'], out['pt_tidx'] = y, p, np.int64(tidx)
out['pt_out'] = np.array([float(l), float(m), float(lv), float(mv)])

# (3) UQ loss: without and with a zero-padded step
c.nn_type, c.UQ = 'RNNUqRangeEstimate', True
y2 = rng.normal(size=(B, T, O)).astype(np.float32)
u, u0, um = ref.Losses(c, tidx).weight_adjusted_uq_loss([y2], [p], [v])
out['uq_y'], out['uq_p'], out['uq_v'] = y2, p, v
out['uq_out'] = np.array([float(u), float(u0), float(um)])
u, u0, um = ref.Losses(c, tidx).weight_adjusted_uq_loss([y], [p], [v])          # y has padded steps
out['uq_pad_y'] = y
out['uq_pad_out'] = np.array([float(u), float(u0), float(um)])
np.savez_compressed(OUT, **out)
print('ok', {k: (v.tolist() if v.size <= 4 else v.shape) for k, v in out.items()})
'''


def main():
    dst = os.path.join(HERE, 'reference_losses.npz')
    r = subprocess.run([sys.executable, '-c', CHILD, dst], capture_output=True, text=True)
    print(r.returncode)
    print((r.stdout.strip().splitlines() or [''])[-1][:1500])
    if r.returncode != 0:
        print('\n'.join(r.stderr.strip().splitlines()[-12:]))


if __name__ == '__main__':
    main()
