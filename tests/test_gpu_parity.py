"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerance: north_star asks for <= 1e-4 relative on fp32 for forward / predict outputs; gradients and
updated weights are held to the same bar (relative to the tensor's max magnitude).
"""
import numpy as np
import pytest
import torch

import lfm_oracle as orc
from util import make_engine, make_problem, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


SHAPES = [
    # B,  T,  F,  O,  H,  L
    (32, 20, 32, 16, 64, 1),     # BASELINE cfg1: system-test.conf shape
    (5, 6, 7, 3, 8, 2),          # ragged everything
    (1, 1, 4, 2, 4, 1),          # degenerate: one window, one step
    (130, 9, 33, 17, 132, 2),    # tile-boundary crossers
]


@pytest.mark.parametrize('shape', SHAPES)
def test_forward_matches_oracle(shape):
    B, T, F, O, H, L = shape
    params, x, y = make_problem(B, T, F, O, H, L, seed=3)
    eng = make_engine(B, T, F, O, H, L, train=False, forward_only=True)
    eng.set_weights(params)
    preds = eng.forward(_cuda(x)).cpu().numpy()
    ref, _ = orc.forward(params, x.astype(np.float64), num_layers=L)
    assert rel_err(preds, ref) < TOL
    # fewer rows than max_batch
    if B > 2:
        p2 = eng.forward(_cuda(x[:B - 2])).cpu().numpy()
        assert rel_err(p2, ref[:B - 2]) < TOL


@pytest.mark.parametrize('shape', SHAPES)
def test_gradients_match_oracle(shape):
    B, T, F, O, H, L = shape
    params, x, y = make_problem(B, T, F, O, H, L, seed=4)
    eng = make_engine(B, T, F, O, H, L, target_idx=O - 1, target_lambda=0.5, rnn_lambda=0.7)
    eng.set_weights(params)
    eng.backward(_cuda(x), _cuda(y))
    got = eng.grads_list()
    tail = eng.grads[eng.n_trainable:eng.n_trainable + 2].cpu().numpy()
    preds, fc = orc.forward(params, x.astype(np.float64), num_layers=L)
    loss, mse, dpred, _ = orc.loss_point_estimate(y.astype(np.float64), preds, target_idx=O - 1, target_lambda=0.5,
                                                  rnn_lambda=0.7)
    ref = orc.backward(dpred, fc, num_layers=L)
    assert tail[0] == pytest.approx(loss, rel=TOL)
    assert tail[1] == pytest.approx(mse, rel=TOL)
    for (name, _, _, _), g, r in zip(eng.trainable_specs, got, ref):
        assert rel_err(g, r) < TOL, name


def test_dropout_masks_match_philox_oracle():
    B, T, F, O, H, L = 12, 5, 8, 4, 16, 2
    params, x, y = make_problem(B, T, F, O, H, L, seed=5)
    kw = dict(dropout=0.3, recurrent_dropout=0.2, seed=521)
    eng = make_engine(B, T, F, O, H, L, train=True, **kw)
    eng.set_weights(params)
    for step, row0 in ((0, 0), (7, 4096)):
        preds = eng.forward(_cuda(x), step=step, row0=row0).cpu().numpy()
        ref, _ = orc.forward(params, x.astype(np.float64), num_layers=L, training=True, step=step, row0=row0, **kw)
        assert rel_err(preds, ref) < TOL
    # gradients with dropout active
    eng.backward(_cuda(x), _cuda(y), step=3, row0=24)
    preds, fc = orc.forward(params, x.astype(np.float64), num_layers=L, training=True, step=3, row0=24, **kw)
    _, _, dpred, _ = orc.loss_point_estimate(y.astype(np.float64), preds, target_idx=0, target_lambda=0.5,
                                             rnn_lambda=0.7)
    ref = orc.backward(dpred, fc, num_layers=L)
    for (name, _, _, _), g, r in zip(eng.trainable_specs, eng.grads_list(), ref):
        assert rel_err(g, r) < TOL, name


@pytest.mark.parametrize('opt', ['Adadelta', 'Adam', 'RMSprop', 'SGD'])
def test_train_steps_match_oracle(opt):
    B, T, F, O, H, L = 16, 6, 8, 4, 16, 2
    params, x, y = make_problem(B, T, F, O, H, L, seed=6)
    cfg = dict(num_layers=L, target_idx=1, target_lambda=0.5, rnn_lambda=0.7, max_grad_norm=0.05, optimizer=opt,
               max_norm=0.8, sgd_momentum=0.9, train=True, dropout=0.0, recurrent_dropout=0.0)
    lr = {'Adadelta': 0.6, 'Adam': 0.01, 'RMSprop': 0.01, 'SGD': 0.05}[opt]
    eng = make_engine(B, T, F, O, H, L, target_idx=1, optimizer=opt, max_grad_norm=0.05, max_norm=0.8,
                      sgd_momentum=0.9)
    eng.set_weights(params)
    p = [q.copy() for q in params]
    slots = orc.zero_slots(opt, p)
    xc, yc = _cuda(x), _cuda(y)
    for it in range(4):
        out = eng.train_step(xc, yc, it, lr).cpu().numpy()
        p, mse, loss, raw, gn = orc.train_step(p, slots, x.astype(np.float64), y.astype(np.float64), it, cfg, lr=lr)
        assert out[1] == pytest.approx(mse, rel=TOL), it
        assert out[0] == pytest.approx(loss, rel=TOL), it
        assert float(eng.grads[eng.n_trainable + 2]) == pytest.approx(gn, rel=TOL)
    for (name, _, _, _), w, r in zip(eng.trainable_specs, eng.get_weights(), p):
        assert rel_err(w, r) < 5 * TOL, name
    # MaxNorm actually bit: every LSTM kernel column norm <= max_norm
    w0 = eng.get_weights()[0]
    assert np.linalg.norm(w0, axis=0).max() <= 0.8 * (1 + 1e-5)


def test_loss_kat_through_c_abi():
    """The reference's only fixture (scripts/model_utils/losses.py:296-301); values derived in test_oracle."""
    y_true = np.array([[[0, 0, 0], [0, 0, 0], [4, 5, 6], [7, 8, 9], [1, 2, 3]],
                       [[0, 0, 0], [0, 0, 0], [1, 2, 3], [4, 5, 6], [7, 8, 9]]], dtype=np.float32)
    y_pred = np.ones_like(y_true)
    eng = make_engine(2, 5, 4, 3, 4, 1, target_idx=2, target_lambda=1.0, rnn_lambda=0.0)
    out = eng.loss(_cuda(y_pred), _cuda(y_true)).cpu().numpy()
    assert out[0] == pytest.approx(34.0) and out[1] == pytest.approx(34.0)
    eng2 = make_engine(2, 5, 4, 3, 4, 1, target_idx=2, target_lambda=0.5, rnn_lambda=0.7)
    out = eng2.loss(_cuda(y_pred), _cuda(y_true)).cpu().numpy()
    assert out[0] == pytest.approx(29.383333, rel=1e-6)
    mc = eng2.mask_count(_cuda(y_true)).cpu().numpy()
    assert mc[0] == 2 and mc[1] == 6


def test_global_denominators_sum_to_single_process_gradient():
    """two shards with global denominators == the unsharded gradient (no NCCL needed to check)."""
    B, T, F, O, H, L = 8, 5, 6, 3, 8, 1
    params, x, y = make_problem(B, T, F, O, H, L, seed=8)
    eng = make_engine(B, T, F, O, H, L, target_idx=1)
    eng.set_weights(params)
    eng.backward(_cuda(x), _cuda(y))
    full = eng.grads[:eng.n_trainable + 2].clone()
    denom = eng.mask_count(_cuda(y))
    acc = torch.zeros_like(full)
    for lo, hi in ((0, 3), (3, 8)):
        eng.backward(_cuda(x[lo:hi]), _cuda(y[lo:hi]), row0=lo, denom=denom)
        acc += eng.grads[:eng.n_trainable + 2]
    assert rel_err(acc.cpu().numpy(), full.cpu().numpy()) < 1e-5


def _make_table(n_keys=6, n_months=40, n_fin=5, n_aux=3, seed=0):
    rng = np.random.RandomState(seed)
    rows = []
    for k in range(n_keys):
        mc = np.exp(rng.normal(5, 2, size=n_months))
        mc[rng.randint(n_months)] = 3.0          # below _MIN_SEQ_NORM
        fin = rng.normal(size=(n_months, n_fin)) * mc[:, None]
        aux = rng.normal(size=(n_months, n_aux))
        rows.append(np.concatenate([np.zeros((n_months, 3)), fin, aux, mc[:, None]], axis=1))
    return np.concatenate(rows, axis=0)


# (5, 3): scalar kernel; (8, 4) and (16, 16): the vectorised kernel (F % 4 == O % 4 == 0) with 16-byte aligned and
# misaligned column groups (the first data column is table column 3)
@pytest.mark.parametrize('n_fin,n_aux', [(5, 3), (8, 4), (16, 16)])
@pytest.mark.parametrize('train', [True, False])
def test_gather_batch_matches_oracle(train, n_fin, n_aux):
    from lfm_quant_b200.engine import gather_batch
    T, stride, fn = 4, 3, 3
    table = _make_table(n_fin=n_fin, n_aux=n_aux)
    n_rows, n_cols = table.shape
    inp_cols = list(range(3, 3 + n_fin + n_aux))
    fin_cols = list(range(3, 3 + n_fin))
    seq_norm_col = n_cols - 1
    rng = np.random.RandomState(1)
    center = rng.normal(size=n_fin + n_aux)
    scale = np.abs(rng.normal(size=n_fin + n_aux)) + 0.5
    scale_ids = [i for i in range(n_fin + n_aux) if i not in (n_fin, n_fin + 2)]   # two aux columns in dont_scale_fields
    inp_idx, tar_idx, valid = [], [], []
    for i in range(12, n_rows - fn, 7):
        for pad in (0, 1, 2):
            seq_len = (T - pad - 1) * stride + 1
            inp_idx.append([i - seq_len + 1, i, pad])
            tar_idx.append([i - seq_len + 1 + fn, i + fn, pad])
            valid.append(True)
    if not train:                                  # windows whose target is missing (data_processing.py:275-279)
        for i in (n_rows - 1, n_rows - 2, 50):
            seq_len = (T - 1) * stride + 1
            inp_idx.append([i - seq_len + 1, i, 0])
            tar_idx.append([i - seq_len + 1 + fn, i, 0])
            valid.append(False)
    inp_idx = np.array(inp_idx, dtype=np.int32)
    tar_idx = np.array(tar_idx, dtype=np.int32)
    for aux_masking in (False, True):
        ref = orc.gather_batch(table, inp_idx, tar_idx, np.array(valid), seq_len=T, stride=stride,
                               inp_cols=inp_cols, fin_cols=fin_cols, seq_norm_col=seq_norm_col, center=center,
                               scale=scale, scale_inp_ids=scale_ids, aux_inp_ids=list(range(n_fin, n_fin + n_aux)),
                               log_squash=True, aux_masking=aux_masking, train=train)
        sflag = np.zeros(n_fin + n_aux, dtype=np.uint8)
        sflag[scale_ids] = 1
        aflag = np.zeros(n_fin + n_aux, dtype=np.uint8)
        aflag[n_fin:] = 1
        x, y, sn = gather_batch(_cuda(table), _cuda(inp_idx), _cuda(tar_idx), seq_len=T, stride=stride,
                                inp_cols=_cuda(np.array(inp_cols, dtype=np.int32)),
                                fin_cols=_cuda(np.array(fin_cols, dtype=np.int32)), seq_norm_col=seq_norm_col,
                                center=_cuda(center), scale=_cuda(scale), scale_flag=_cuda(sflag),
                                aux_flag=_cuda(aflag), log_squasher=True, aux_masking=aux_masking)
        np.testing.assert_array_equal(sn.cpu().numpy(), ref[2])
        # fp64 arithmetic then cast: equal up to 1 fp32 ulp (device log1p vs libm log1p), NaNs in the same places
        np.testing.assert_array_equal(np.isnan(y.cpu().numpy()), np.isnan(ref[1]))
        if not train:
            assert np.isnan(ref[1]).any()
        np.testing.assert_array_max_ulp(x.cpu().numpy(), ref[0], maxulp=1)
        np.testing.assert_array_max_ulp(np.nan_to_num(y.cpu().numpy()), np.nan_to_num(ref[1]), maxulp=1)
        assert ref[2].min() == 10.0                # the _MIN_SEQ_NORM floor was exercised


def test_bad_arguments_raise():
    from lfm_quant_b200._native import LfmqError
    eng = make_engine(4, 3, 4, 2, 8, 1, forward_only=True)
    x = torch.zeros(5, 3, 4, device='cuda')
    with pytest.raises(LfmqError):
        eng.forward(x)                              # B > max_batch
    with pytest.raises(LfmqError):
        eng.backward(torch.zeros(4, 3, 4, device='cuda'), torch.zeros(4, 3, 2, device='cuda'))  # forward_only


def test_predict_full_size_rows_match_oracle():
    """BASELINE configs[4]: predict.py path at B=65536, T=48, H=256 in the fp32 parity mode.  Windows are independent,
    so a random subset of rows is checked against the oracle (<= 1e-4 relative) and the kept scalar
    pred[:, -1, target_idx] (predict.py:284-291) is what is compared as well."""
    B, T, F, O, H, L = 65536, 48, 32, 16, 256, 1
    rng = np.random.RandomState(17)
    params = [p.astype(np.float32).astype(np.float64) for p in orc.init_params(L, F, O, H, init_scale=1.0, seed=521)]
    x = rng.standard_normal((B, T, F)).astype(np.float32)
    eng = make_engine(B, T, F, O, H, L, train=False, forward_only=True)
    eng.set_weights(params)
    preds = eng.forward(_cuda(x)).cpu().numpy()
    rows = rng.choice(B, size=96, replace=False)
    ref, _ = orc.forward(params, x[rows].astype(np.float64), num_layers=L)
    assert rel_err(preds[rows], ref) < TOL
    assert rel_err(preds[rows, -1, 3], ref[:, -1, 3]) < TOL
    eng.close()
    # the tensor-core path on the same inputs, bf16 tolerance
    e16 = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='bf16')
    e16.set_weights(params)
    p16 = e16.forward(_cuda(x)).cpu().numpy()
    assert np.isfinite(p16).all()
    assert np.abs(p16[rows] - ref).max() / np.abs(ref).max() < 6e-2


def test_cfg3_full_size_dropout_block_and_shard_sum():
    """BASELINE configs[2]: B=4096, T=48, F=32, H=512, L=2 with dropout (rate 0.2) on the fp32 path.
    The oracle cannot run 4096 windows of this size in seconds, so: (1) a contiguous block of rows of the training-mode
    forward is checked against the oracle with the same global row offset (dropout masks are keyed by global row);
    (2) size-independent property: gradients of two row shards taken with the global loss denominators sum to the
    unsharded gradient (the data-parallel contract)."""
    B, T, F, O, H, L = 4096, 48, 32, 16, 512, 2
    kw = dict(dropout=0.2, recurrent_dropout=0.0, seed=521)
    params, x, y = make_problem(B, T, F, O, H, L, seed=9, init_scale=0.1)
    eng = make_engine(B, T, F, O, H, L, train=True, target_idx=3, **kw)
    eng.set_weights(params)
    xc, yc = _cuda(x), _cuda(y)
    preds = eng.forward(xc, step=5, row0=0).cpu().numpy()
    r0, n = 1000, 24
    ref, _ = orc.forward(params, x[r0:r0 + n].astype(np.float64), num_layers=L, training=True, step=5, row0=r0, **kw)
    assert rel_err(preds[r0:r0 + n], ref) < TOL
    eng.backward(xc, yc, step=5)
    nt = eng.n_trainable
    full = eng.grads[:nt + 2].clone()
    denom = eng.mask_count(yc)
    acc = torch.zeros_like(full)
    for lo, hi in ((0, 1500), (1500, B)):
        eng.backward(xc[lo:hi].contiguous(), yc[lo:hi].contiguous(), step=5, row0=lo, denom=denom)
        acc += eng.grads[:nt + 2]
    assert torch.isfinite(full).all() and float(full[:nt].abs().max()) > 0
    assert rel_err(acc.cpu().numpy(), full.cpu().numpy()) < TOL
    eng.close()


# ---- GRU cell through the C ABI (LFMQ_CELL_GRU; rnn_point_estimate.py:89-98) ----
GRU_SHAPES = [(32, 20, 32, 16, 64, 1), (5, 6, 7, 3, 8, 2), (1, 1, 4, 2, 4, 1), (130, 9, 33, 17, 132, 2)]


@pytest.mark.parametrize('shape', GRU_SHAPES)
def test_gru_forward_and_gradients_match_oracle(shape):
    B, T, F, O, H, L = shape
    params, x, y = make_problem(B, T, F, O, H, L, seed=13, rnn_cell='gru')
    eng = make_engine(B, T, F, O, H, L, target_idx=O - 1, rnn_cell='gru')
    assert [n for n, _, _, _ in eng.trainable_specs][:3] == ['gru_1/kernel', 'gru_1/recurrent_kernel', 'gru_1/bias']
    eng.set_weights(params)
    preds = eng.forward(_cuda(x)).cpu().numpy()
    ref, fc = orc.forward(params, x.astype(np.float64), num_layers=L, rnn_cell='gru')
    assert rel_err(preds, ref) < TOL
    eng.backward(_cuda(x), _cuda(y))
    loss, mse, dpred, _ = orc.loss_point_estimate(y.astype(np.float64), ref, target_idx=O - 1, target_lambda=0.5,
                                                  rnn_lambda=0.7)
    gref = orc.backward(dpred, fc, num_layers=L, rnn_cell='gru')
    tail = eng.grads[eng.n_trainable:eng.n_trainable + 2].cpu().numpy()
    assert tail[0] == pytest.approx(loss, rel=TOL) and tail[1] == pytest.approx(mse, rel=TOL)
    for (name, _, _, _), g, r in zip(eng.trainable_specs, eng.grads_list(), gref):
        assert g.shape == r.shape and rel_err(g, r) < TOL, name


def test_gru_dropout_and_train_steps_match_oracle():
    B, T, F, O, H, L = 16, 6, 8, 4, 16, 2
    params, x, y = make_problem(B, T, F, O, H, L, seed=14, rnn_cell='gru')
    kw = dict(dropout=0.3, recurrent_dropout=0.2, seed=521)
    cfg = dict(num_layers=L, target_idx=1, target_lambda=0.5, rnn_lambda=0.7, max_grad_norm=0.05, optimizer='Adam',
               max_norm=0.8, sgd_momentum=0.0, train=True, rnn_cell='gru', **kw)
    eng = make_engine(B, T, F, O, H, L, target_idx=1, optimizer='Adam', max_grad_norm=0.05, max_norm=0.8,
                      rnn_cell='gru', **kw)
    eng.set_weights(params)
    p = [q.copy() for q in params]
    slots = orc.zero_slots('Adam', p)
    xc, yc = _cuda(x), _cuda(y)
    for it in range(3):
        out = eng.train_step(xc, yc, it, 0.01).cpu().numpy()
        p, mse, loss, raw, gn = orc.train_step(p, slots, x.astype(np.float64), y.astype(np.float64), it, cfg, lr=0.01)
        assert out[0] == pytest.approx(loss, rel=TOL) and out[1] == pytest.approx(mse, rel=TOL), it
    for (name, _, _, _), w, r in zip(eng.trainable_specs, eng.get_weights(), p):
        assert rel_err(w, r) < 5 * TOL, name
    assert np.linalg.norm(eng.get_weights()[0], axis=0).max() <= 0.8 * (1 + 1e-5)       # MaxNorm on gru_1/kernel


def test_gru_is_refused_loudly_on_the_bf16_path():
    from lfm_quant_b200 import _native as N
    with pytest.raises(N.LfmqError, match='LSTM cell only'):
        make_engine(256, 8, 32, 16, 256, 1, precision='bf16', rnn_cell='gru')


# ---- RNNUqRangeEstimate through the C ABI (cfg.uq = 1) ----
def _uq_problem(B, T, F, O, H, L, seed, rnn_cell='lstm'):
    rng = np.random.RandomState(seed)
    params = orc.init_params(L, F, O, H, init_scale=0.5, seed=seed + 1, dtype=np.float64, rnn_cell=rnn_cell, uq=True)
    params[-1] = rng.normal(size=O) * 0.5
    params[-3] = rng.normal(size=O) * 0.5
    params = [p.astype(np.float32).astype(np.float64) for p in params]
    x = rng.normal(size=(B, T, F)).astype(np.float32)
    y = rng.normal(size=(B, T, O)).astype(np.float32)          # no zero-padded steps: those make the loss NaN (below)
    return params, x, y


@pytest.mark.parametrize('shape,cell', [((32, 20, 32, 16, 64, 1), 'lstm'), ((5, 6, 7, 3, 8, 2), 'lstm'),
                                        ((130, 9, 33, 17, 132, 2), 'gru')])
def test_uq_forward_loss_and_gradients_match_oracle(shape, cell):
    B, T, F, O, H, L = shape
    params, x, y = _uq_problem(B, T, F, O, H, L, seed=21, rnn_cell=cell)
    kw = dict(dropout=0.2, recurrent_dropout=0.1, seed=521)
    eng = make_engine(B, T, F, O, H, L, target_idx=O - 1, uq=True, rnn_cell=cell, train=False, **kw)   # train=False:
    assert [n for n, _, _, _ in eng.trainable_specs][-4:] == orc.param_names(L, cell, uq=True)[-4:]    # dropout still on
    eng.set_weights(params)
    pred, var = eng.forward(_cuda(x), step=4, row0=64)
    okw = dict(num_layers=L, rnn_cell=cell, step=4, row0=64, **kw)
    rp, rv, fc = orc.forward_uq(params, x.astype(np.float64), **okw)
    assert rel_err(pred.cpu().numpy(), rp) < TOL and rel_err(var.cpu().numpy(), rv) < TOL
    assert float(var.min()) >= 1e-6
    out = eng.loss_uq(pred, var, _cuda(y)).cpu().numpy()
    loss, uq0, mse0, dp, dv = orc.loss_uq_estimate(y.astype(np.float64), rp, rv, target_idx=O - 1, target_lambda=0.5,
                                                   rnn_lambda=0.7)
    assert out[0] == pytest.approx(loss, rel=TOL) and out[1] == pytest.approx(uq0, rel=TOL)
    assert out[2] == pytest.approx(mse0, rel=TOL)
    eng.backward(_cuda(x), _cuda(y), step=4, row0=64)
    gref = orc.backward_uq(dp, dv, fc, num_layers=L, rnn_cell=cell)
    nt = eng.n_trainable
    tail = eng.grads[nt:nt + 5].cpu().numpy()
    assert tail[0] == pytest.approx(loss, rel=TOL) and tail[1] == pytest.approx(mse0, rel=TOL)
    assert tail[4] == pytest.approx(uq0, rel=TOL)
    for (name, _, _, _), g, r in zip(eng.trainable_specs, eng.grads_list(), gref):
        assert rel_err(g, r) < TOL, name


def test_uq_train_steps_match_oracle():
    B, T, F, O, H, L = 16, 6, 8, 4, 16, 2
    params, x, y = _uq_problem(B, T, F, O, H, L, seed=22)
    kw = dict(dropout=0.3, recurrent_dropout=0.0, seed=521)
    cfg = dict(num_layers=L, target_idx=1, target_lambda=0.5, rnn_lambda=0.7, max_grad_norm=0.5, optimizer='Adadelta',
               max_norm=0.8, uq=True, **kw)
    eng = make_engine(B, T, F, O, H, L, target_idx=1, optimizer='Adadelta', max_grad_norm=0.5, max_norm=0.8, uq=True, **kw)
    eng.set_weights(params)
    p = [q.copy() for q in params]
    slots = orc.zero_slots('Adadelta', p)
    xc, yc = _cuda(x), _cuda(y)
    for it in range(3):
        out = eng.train_step(xc, yc, it, 0.6).cpu().numpy()
        p, mse, loss, raw, gn, uq0 = orc.train_step(p, slots, x.astype(np.float64), y.astype(np.float64), it, cfg, lr=0.6)
        assert out[0] == pytest.approx(uq0, rel=TOL) and out[1] == pytest.approx(mse, rel=TOL), it
        assert float(eng.grads[eng.n_trainable]) == pytest.approx(loss, rel=TOL)
    for (name, _, _, _), w, r in zip(eng.trainable_specs, eng.get_weights(), p):
        assert rel_err(w, r) < 5 * TOL, name


def test_uq_loss_is_nan_with_a_padded_step_as_in_the_reference():
    B, T, F, O, H, L = 8, 5, 6, 3, 8, 1
    params, x, y = _uq_problem(B, T, F, O, H, L, seed=23)
    y[0, :2] = 0.0
    eng = make_engine(B, T, F, O, H, L, target_idx=1, uq=True)
    eng.set_weights(params)
    pred, var = eng.forward(_cuda(x))
    out = eng.loss_uq(pred, var, _cuda(y)).cpu().numpy()
    ref = orc.loss_uq_estimate(y.astype(np.float64), pred.cpu().numpy().astype(np.float64),
                               var.cpu().numpy().astype(np.float64), target_idx=1, target_lambda=0.5, rnn_lambda=0.7)
    assert np.isnan(out[0]) and np.isnan(ref[0])
    assert out[1] == pytest.approx(ref[1], rel=TOL) and out[2] == pytest.approx(ref[2], rel=TOL)


def test_uq_handles_refuse_the_point_estimate_entry_points_and_bf16():
    from lfm_quant_b200 import _native as N
    eng = make_engine(4, 3, 4, 2, 4, 1, uq=True)
    out = torch.empty(4, 3, 2, device='cuda')
    x = torch.zeros(4, 3, 4, device='cuda')
    rc = eng.lib.lfmq_forward(eng.handle, x.data_ptr(), 4, 0, 0, out.data_ptr(), None)
    assert rc != 0 and b'lfmq_forward_uq' in eng.lib.lfmq_last_error()
    with pytest.raises(N.LfmqError, match='point-estimate head only'):
        make_engine(256, 8, 32, 16, 256, 1, precision='bf16', uq=True)
