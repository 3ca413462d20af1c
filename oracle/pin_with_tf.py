"""Pin the NumPy oracle against real TensorFlow/Keras -- TEST INFRASTRUCTURE ONLY, never imported by the product.

STATUS: the comparison branch of this script HAS NEVER BEEN EXECUTED.  TensorFlow is not installed in the build image
or on the GPU boxes, so only the "TF oracle unavailable" exit is exercised (tests/test_oracle.py).
Until someone runs it where TensorFlow exists and it prints PINNED, lfm_oracle.py stays "parity unpinned".

What it does when `import tensorflow` works: builds the layer stack the reference builds for one LSTM layer
(scripts/models/point_estimate/rnn_point_estimate.py:80-105: LSTM(H, return_sequences=True) -> BatchNormalization ->
Dropout -> Dense(O)), injects the oracle's seeded weights with set_weights, and compares at 1e-4 relative (fp32):
  1. model(x, training=False)                 vs  lfm_oracle.forward
  2. the weighted-MSE loss of scripts/model_utils/losses.py:55-98,121-135 written with tf ops here vs
     lfm_oracle.loss_point_estimate
  3. tape.gradient(loss, trainable_variables) vs  lfm_oracle.backward   (the path of scripts/train.py:181-192)
If the reference tree is importable (LFM_QUANT_REF=/path/to/lfm_quant) its own Losses class is used for step 2 instead
of the local restatement.  Dropout rate is 0 here: Keras' RNG stream cannot be matched, masks are pinned by
test_philox_kat instead.

usage: python oracle/pin_with_tf.py        (exit 0 on PINNED or on unavailable, 1 on mismatch)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lfm_oracle as orc  # noqa: E402

TOL = 1e-4


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def main():
    try:
        import tensorflow as tf
    except Exception as e:  # ModuleNotFoundError here; any import-time failure counts as unavailable
        print('TF oracle unavailable: %s: %s' % (type(e).__name__, e))
        return 0
    from tensorflow.keras import layers

    B, T, F, O, H, L = 16, 7, 6, 3, 8, 1
    tidx, p1, p2 = 1, 0.5, 0.7
    rng = np.random.RandomState(11)
    params = orc.init_params(L, F, O, H, init_scale=1.0, seed=521, dtype=np.float32)
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    y = rng.standard_normal((B, T, O)).astype(np.float32)
    y[0, :2] = 0.0                                   # a masked prefix, as padded windows have

    inp = layers.Input(shape=(T, F))
    h = layers.LSTM(H, return_sequences=True, name='lstm_1')(inp, training=False)
    h = layers.BatchNormalization(name='batch_normalization')(h, training=False)
    out = layers.Dense(O, name='OUTPUT_1')(h)
    model = tf.keras.Model(inp, out)
    W, U, b, gamma, beta, Wo, bo = params
    model.get_layer('lstm_1').set_weights([W, U, b])
    model.get_layer('batch_normalization').set_weights([gamma, beta, np.zeros(H, np.float32), np.ones(H, np.float32)])
    model.get_layer('OUTPUT_1').set_weights([Wo, bo])

    p64 = [q.astype(np.float64) for q in params]
    ref_pred, fcache = orc.forward(p64, x.astype(np.float64), num_layers=L)
    ref_loss, ref_mse, dpred, _ = orc.loss_point_estimate(y.astype(np.float64), ref_pred, target_idx=tidx,
                                                          target_lambda=p1, rnn_lambda=p2)
    ref_grads = orc.backward(dpred, fcache, num_layers=L)

    def tf_loss(y_true, y_pred):
        m = tf.cast(tf.reduce_any(tf.not_equal(y_true, 0.0), axis=-1, keepdims=True), y_pred.dtype)
        yp = y_pred * m
        mse0 = tf.reduce_mean(tf.square(y_true[:, -1, tidx] - yp[:, -1, tidx]))
        mse1 = tf.reduce_mean(tf.square(y_true[:, -1, :] - yp[:, -1, :]))
        mse2 = tf.reduce_sum(tf.square(yp - y_true)) / (tf.reduce_sum(m) * O)
        return p1 * mse0 + (1 - p1) * (p2 * mse1 + (1 - p2) * mse2), mse0

    tv = [model.get_layer(n).trainable_variables for n in ('lstm_1', 'batch_normalization', 'OUTPUT_1')]
    tv = [v for grp in tv for v in grp]
    with tf.GradientTape() as tape:
        pred = model(tf.constant(x), training=False)
        loss, mse0 = tf_loss(tf.constant(y), pred)
    grads = tape.gradient(loss, tv)

    bad = 0
    checks = [('forward', rel(pred.numpy(), ref_pred)), ('loss', abs(float(loss) - ref_loss) / abs(ref_loss)),
              ('mse_0', abs(float(mse0) - ref_mse) / abs(ref_mse))]
    checks += [('grad ' + n, rel(g.numpy(), r)) for n, g, r in zip(orc.param_names(L), grads, ref_grads)]
    for name, e in checks:
        ok = e < TOL
        bad += not ok
        print('%-40s rel err %.3e  %s' % (name, e, 'ok' if ok else 'MISMATCH'))
    print('PINNED against tensorflow %s' % tf.__version__ if not bad else 'NOT PINNED: %d mismatches' % bad)
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
