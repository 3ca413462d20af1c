"""GPU tests of the general tensor-core path (csrc/rnn_tc.cu): LFMQ_PREC_BF16 for every LSTM point-estimate shape outside
the H=256 / L=1 cluster kernels (H in {64, 128, 256, 512}, stacked layers, dropout, recurrent dropout) and
LFMQ_PREC_BF16X3, the fp32-tolerance forward (three bf16 products per GEMM, accurate gate nonlinearities).

Checker = the fp64 oracle (oracle/lfm_oracle.py).  Tolerances: bf16 3e-2 on outputs / loss, gradients per tensor
max-norm relative 6e-2 and cosine 0.999 (T <= 8 here; T = 48 in test_gpu_baseline_shapes.py); bf16x3 1e-4 on outputs
(the north-star tolerance of predict.py:129).
"""
import os

import numpy as np
import pytest
import torch

import lfm_oracle as orc
from util import make_engine, make_problem, rel_err

pytestmark = pytest.mark.gpu

BF16_TOL = 3e-2
X3_TOL = 1e-4


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _oracle_grads(params, x, y, target_idx, L, **kw):
    preds, fc = orc.forward(params, x.astype(np.float64), num_layers=L, **kw)
    loss, mse, dpred, _ = orc.loss_point_estimate(y.astype(np.float64), preds, target_idx=target_idx,
                                                  target_lambda=0.5, rnn_lambda=0.7)
    return loss, mse, orc.backward(dpred, fc, num_layers=L)


# (B, T, F, O, H, L): ragged batches, one and two layers, narrow inputs / outputs, every hidden size of the path
FWD_SHAPES = [(200, 5, 32, 16, 64, 1), (130, 4, 20, 7, 128, 2), (300, 6, 32, 16, 512, 2), (129, 3, 9, 3, 192, 3),
              (64, 8, 70, 16, 320, 1)]


@pytest.mark.parametrize('shape', FWD_SHAPES)
def test_generic_forward_matches_oracle(shape):
    B, T, F, O, H, L = shape
    params, x, y = make_problem(B, T, F, O, H, L, seed=3, init_scale=0.3)
    eng = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='bf16', target_idx=O - 1)
    eng.set_weights(params)
    preds = eng.forward(_cuda(x)).cpu().numpy()
    ref, _ = orc.forward(params, x.astype(np.float64), num_layers=L)
    assert np.isfinite(preds).all()
    assert rel_err(preds, ref) < BF16_TOL
    p2 = eng.forward(_cuda(x[:B - 37])).cpu().numpy()          # ragged call on the same handle
    assert rel_err(p2, ref[:B - 37]) < BF16_TOL
    eng.close()


def test_generic_path_at_h256_matches_cluster_path_and_oracle():
    """LFMQ_FORCE_GENERIC=1 sends the H=256 / L=1 shape through the general path as well."""
    B, T, F, O, H, L = 300, 6, 32, 16, 256, 1
    params, x, y = make_problem(B, T, F, O, H, L, seed=4, init_scale=0.3)
    ref, _ = orc.forward(params, x.astype(np.float64), num_layers=L)
    outs = {}
    for force in ('0', '1'):
        os.environ['LFMQ_FORCE_GENERIC'] = force
        try:
            eng = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='bf16', target_idx=3)
        finally:
            os.environ.pop('LFMQ_FORCE_GENERIC', None)
        eng.set_weights(params)
        outs[force] = eng.forward(_cuda(x)).cpu().numpy()
        eng.close()
    assert rel_err(outs['1'], ref) < BF16_TOL and rel_err(outs['0'], ref) < BF16_TOL


@pytest.mark.parametrize('shape,kw', [
    ((200, 5, 32, 16, 64, 1), dict()),
    ((130, 4, 20, 7, 128, 2), dict(dropout=0.25, recurrent_dropout=0.2, seed=99)),
    ((300, 6, 32, 16, 512, 2), dict(dropout=0.2, seed=521)),
    ((256, 8, 32, 16, 256, 2), dict(recurrent_dropout=0.3, seed=7)),
    ((140, 1, 32, 16, 128, 1), dict()),
])
def test_generic_gradients_match_oracle(shape, kw):
    B, T, F, O, H, L = shape
    params, x, y = make_problem(B, T, F, O, H, L, seed=21, init_scale=0.3)
    eng = make_engine(B, T, F, O, H, L, target_idx=O - 1, precision='bf16', train=True, **kw)
    eng.set_weights(params)
    eng.backward(_cuda(x), _cuda(y), step=5, row0=512)
    tail = eng.grads[eng.n_trainable:eng.n_trainable + 2].cpu().numpy()
    loss, mse, ref = _oracle_grads(params, x, y, O - 1, L, training=True, step=5, row0=512, **kw)
    assert tail[0] == pytest.approx(loss, rel=BF16_TOL)
    assert tail[1] == pytest.approx(mse, rel=BF16_TOL)
    for (name, _, _, _), g, r in zip(eng.trainable_specs, eng.grads_list(), ref):
        assert np.isfinite(g).all(), name
        if np.abs(r).max() == 0.0:            # dU at T=1 (h_prev = 0)
            assert np.abs(g).max() < 1e-6, name
            continue
        cos = float(np.sum(g * r) / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
        assert cos > 0.999, (name, cos)
        assert rel_err(g, r) < 2 * BF16_TOL, (name, rel_err(g, r))
    eng.close()


def test_generic_smaller_batch_after_larger_on_one_handle():
    """The weight-gradient GEMMs sum over all time-major rows of the workspace: rows of row tiles a later, smaller batch
    does not launch must not leak an earlier call's values into its gradients."""
    B, T, F, O, H, L = 400, 4, 32, 16, 128, 2
    params, x, y = make_problem(B, T, F, O, H, L, seed=31, init_scale=0.3)
    eng = make_engine(B, T, F, O, H, L, target_idx=3, precision='bf16')
    eng.set_weights(params)
    eng.backward(_cuda(x), _cuda(y))
    n = 100
    eng.backward(_cuda(x[:n]), _cuda(y[:n]))
    loss, mse, ref = _oracle_grads(params, x[:n], y[:n], 3, L)
    for (name, _, _, _), g, r in zip(eng.trainable_specs, eng.grads_list(), ref):
        assert rel_err(g, r) < 2 * BF16_TOL, name
    eng.close()


def test_generic_train_steps_track_oracle():
    B, T, F, O, H, L = 256, 6, 32, 16, 128, 2
    params, x, y = make_problem(B, T, F, O, H, L, seed=23, init_scale=0.3)
    cfg = dict(num_layers=L, target_idx=3, target_lambda=0.5, rnn_lambda=0.7, max_grad_norm=50.0, optimizer='Adadelta',
               max_norm=3.0, train=True, dropout=0.0, recurrent_dropout=0.0)
    eng = make_engine(B, T, F, O, H, L, target_idx=3, optimizer='Adadelta', precision='bf16')
    eng.set_weights(params)
    p = [q.copy() for q in params]
    slots = orc.zero_slots('Adadelta', p)
    xc, yc = _cuda(x), _cuda(y)
    for it in range(3):
        out = eng.train_step(xc, yc, it, 0.6).cpu().numpy()
        p, mse, loss, raw, gn = orc.train_step(p, slots, x.astype(np.float64), y.astype(np.float64), it, cfg, lr=0.6)
        assert out[1] == pytest.approx(mse, rel=BF16_TOL), it
        assert out[0] == pytest.approx(loss, rel=BF16_TOL), it
    for (name, _, _, _), w, r in zip(eng.trainable_specs, eng.get_weights(), p):
        assert rel_err(w, r) < BF16_TOL, name
    eng.close()


@pytest.mark.parametrize('shape', [(300, 6, 32, 16, 256, 1), (130, 5, 20, 7, 128, 2), (64, 48, 32, 16, 256, 1)])
def test_bf16x3_forward_meets_the_fp32_tolerance(shape):
    """LFMQ_PREC_BF16X3: forward / predict.py outputs within 1e-4 relative of the oracle (north_star), on tensor cores."""
    B, T, F, O, H, L = shape
    params, x, y = make_problem(B, T, F, O, H, L, seed=8, init_scale=0.5)
    eng = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='bf16x3', target_idx=O - 1)
    eng.set_weights(params)
    preds = eng.forward(_cuda(x)).cpu().numpy()
    ref, _ = orc.forward(params, x.astype(np.float64), num_layers=L)
    err = rel_err(preds, ref)
    print('bf16x3 %s: max-norm rel err %.3e' % (shape, err))
    assert err < X3_TOL
    eng.close()


def test_bf16x3_refuses_training_handles_loudly():
    from lfm_quant_b200._native import LfmqError
    with pytest.raises(LfmqError):
        make_engine(8, 4, 32, 16, 256, 1, precision='bf16x3')


def test_tensor_core_paths_refuse_unsupported_shapes_loudly():
    from lfm_quant_b200._native import LfmqError
    with pytest.raises(LfmqError):
        make_engine(8, 4, 32, 16, 96, 1, precision='bf16')       # H not a multiple of 64
    with pytest.raises(LfmqError):
        make_engine(8, 4, 32, 17, 128, 1, precision='bf16')      # n_outputs > 16


def test_persistent_steps_equal_one_launch_per_step():
    """The persistent step launches (GArgs::n_steps, one barrier per row-tile group) run the same arithmetic in the same
    order as one launch per time step: the gradients must agree to fp32 rounding (the head's loss / bias-gradient sums
    go through shared-memory atomics, so a last-bit difference between two runs is possible whatever the mode).
    LFMQ_GEN_PERSIST is read once per process, hence the two subprocesses."""
    import os
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/oracle'); sys.path.insert(0, %r + '/tests')
from util import make_engine, make_problem
B, T, F, O, H, L = 200, 12, 16, 8, 128, 2
params, x, y = make_problem(B, T, F, O, H, L, seed=4)
eng = make_engine(B, T, F, O, H, L, target_idx=1, precision='bf16', train=True, dropout=0.2, recurrent_dropout=0.1, seed=9)
eng.set_weights(params)
xc, yc = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
eng.backward(xc, yc, step=3, row0=7)
torch.cuda.synchronize()
np.save(sys.argv[1], eng.grads.detach().cpu().numpy())
''' % (root, root, root)
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for mode in ('1', '0'):
            path = os.path.join(d, 'g%s.npy' % mode)
            env = dict(os.environ, LFMQ_GEN_PERSIST=mode)
            subprocess.run([sys.executable, '-c', code, path], check=True, env=env, timeout=120)
            outs.append(np.load(path))
    assert np.isfinite(outs[0]).all() and np.abs(outs[0]).max() > 0
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-6, atol=1e-9)
