"""Pins the loss rows (a7, and the UQ loss of 8f-2) to the reference's own code.

tests/golden/reference_losses.npz holds what the unmodified scripts/model_utils/losses.py returned in the build container
with its 14 TensorFlow primitives mapped to NumPy (generator: tests/golden/make_reference_losses.py): the masking,
slicing, weights and denominators are the reference's lines, the elementwise / reduction arithmetic is NumPy float32.
The CUDA loss kernels are held to the oracle in tests/test_gpu_parity.py (lfmq_loss, lfmq_loss_uq)."""
import os

import numpy as np
import pytest

import lfm_oracle as orc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_losses.npz'))
RTOL = 2e-6          # the reference side accumulated in float32


def test_reference_file_fixture():
    """losses.py:287-310, the only fixture the reference ships: 34.0 / 29.383333 were derived by hand in
    tests/test_oracle.py -- this is the reference's own answer."""
    for tag in ('a', 'b'):
        p1, p2, loss, mse = G['fix_' + tag]
        l, m, _, _ = orc.loss_point_estimate(G['fix_y_true'].astype(np.float64), G['fix_y_pred'].astype(np.float64),
                                             target_idx=2, target_lambda=p1, rnn_lambda=p2)
        assert l == pytest.approx(loss, rel=RTOL) and m == pytest.approx(mse, rel=RTOL)
    assert G['fix_a'][2] == pytest.approx(34.0) and G['fix_b'][2] == pytest.approx(29.383333, rel=1e-6)


def test_point_estimate_loss_with_padded_steps():
    loss, mse, loss_v, mse_v = G['pt_out']
    l, m, _, _ = orc.loss_point_estimate(G['pt_y'].astype(np.float64), G['pt_p'].astype(np.float64),
                                         target_idx=int(G['pt_tidx']), target_lambda=0.5, rnn_lambda=0.7)
    assert l == pytest.approx(loss, rel=RTOL) and m == pytest.approx(mse, rel=RTOL)
    assert (loss_v, mse_v) == (loss, mse)            # is_validation only matters on the MLP / Huber branch (:104-107)


def test_uq_loss_and_its_nan_on_padded_steps():
    kw = dict(target_idx=int(G['pt_tidx']), target_lambda=0.5, rnn_lambda=0.7)
    u, u0, um = G['uq_out']
    l, l0, m, _, _ = orc.loss_uq_estimate(G['uq_y'].astype(np.float64), G['uq_p'].astype(np.float64),
                                          G['uq_v'].astype(np.float64), **kw)
    assert l == pytest.approx(u, rel=RTOL) and l0 == pytest.approx(u0, rel=RTOL) and m == pytest.approx(um, rel=RTOL)
    # a zero-padded step: the reference multiplies the variance by the mask, then divides by it and takes its log
    u, u0, um = G['uq_pad_out']
    l, l0, m, _, _ = orc.loss_uq_estimate(G['uq_pad_y'].astype(np.float64), G['uq_p'].astype(np.float64),
                                          G['uq_v'].astype(np.float64), **kw)
    assert np.isnan(u) and np.isnan(l)
    assert l0 == pytest.approx(u0, rel=RTOL) and m == pytest.approx(um, rel=RTOL)
