"""Pins the CPU oracle: reference loss fixture (KAT), torch autograd second opinion, fp64 finite
differences, Philox known-answer vectors.  CPU only."""
import numpy as np
import pytest
import torch

import lfm_oracle as orc


def test_philox_kat():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        got = orc.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == exp


def _loss_fixture():
    """Inputs of the reference's only fixture, scripts/model_utils/losses.py:296-301."""
    y_true = np.array([[[0, 0, 0], [0, 0, 0], [4, 5, 6], [7, 8, 9], [1, 2, 3]],
                       [[0, 0, 0], [0, 0, 0], [1, 2, 3], [4, 5, 6], [7, 8, 9]]], dtype=np.float64)
    y_pred = np.ones_like(y_true)
    return y_true, y_pred


def test_loss_kat_reference_fixture():
    """Expected values derived by hand from losses.py:72-98,131-135 (the reference only prints them)."""
    y_true, y_pred = _loss_fixture()
    # settings of the reference __main__ block (losses.py:291-294): Losses(config, 2), lambdas (1.0, 0.0)
    loss, mse, _, sums = orc.loss_point_estimate(y_true, y_pred, target_idx=2, target_lambda=1.0, rnn_lambda=0.0)
    assert mse == pytest.approx(34.0)          # ((3-1)^2 + (9-1)^2) / 2
    assert loss == pytest.approx(34.0)
    # default lambdas (lfm_quant.py:65-66): 0.5, 0.7
    loss, mse, _, (s0, s1, s2, m) = orc.loss_point_estimate(y_true, y_pred, target_idx=2, target_lambda=0.5,
                                                             rnn_lambda=0.7)
    assert m == 6
    assert s0 / 2 == pytest.approx(34.0)
    assert s1 / 6 == pytest.approx(154.0 / 6)   # mse_1 = 25.6667
    assert s2 / 18 == pytest.approx(408.0 / 18)  # mse_2 = 22.6667
    assert loss == pytest.approx(0.5 * 34.0 + 0.5 * (0.7 * 154.0 / 6 + 0.3 * 408.0 / 18))
    assert loss == pytest.approx(29.383333, rel=1e-6)


def _rand_problem(B=5, T=6, F=7, H=8, O=3, L=2, seed=0, dt=np.float64):
    rng = np.random.RandomState(seed)
    params = orc.init_params(L, F, O, H, init_scale=0.5, seed=seed + 1, dtype=dt)
    # perturb gamma/beta/biases so their grads are non-trivial
    for l in range(L):
        params[5 * l + 2] = params[5 * l + 2] + rng.normal(size=4 * H).astype(dt) * 0.1
        params[5 * l + 3] = params[5 * l + 3] + rng.normal(size=H).astype(dt) * 0.1
        params[5 * l + 4] = params[5 * l + 4] + rng.normal(size=H).astype(dt) * 0.1
    x = rng.normal(size=(B, T, F)).astype(dt)
    y = rng.normal(size=(B, T, O)).astype(dt)
    y[0, :2, :] = 0.0   # padded steps -> mask 0
    return params, x, y


def _torch_forward(params, x, L):
    cur = torch.from_numpy(x)
    tp = [torch.from_numpy(p.copy()).requires_grad_(True) for p in params]
    for l in range(L):
        W, U, b, gamma, beta = tp[5 * l:5 * l + 5]
        H = U.shape[0]
        B, T, _ = cur.shape
        h = torch.zeros(B, H, dtype=cur.dtype)
        c = torch.zeros(B, H, dtype=cur.dtype)
        outs = []
        for t in range(T):
            z = cur[:, t] @ W + h @ U + b
            i, f, g, o = z[:, :H].sigmoid(), z[:, H:2 * H].sigmoid(), z[:, 2 * H:3 * H].tanh(), z[:, 3 * H:].sigmoid()
            c = f * c + i * g
            h = o * c.tanh()
            outs.append(h)
        hs = torch.stack(outs, 1)
        cur = gamma * hs / np.sqrt(1.0 + orc.BN_EPS) + beta
    return cur @ tp[5 * L] + tp[5 * L + 1], tp


def test_lstm_matches_torch_nn_lstm():
    rng = np.random.RandomState(3)
    B, T, F, H = 4, 9, 6, 5
    x = rng.normal(size=(B, T, F))
    W = rng.normal(size=(F, 4 * H)) * 0.3
    U = rng.normal(size=(H, 4 * H)) * 0.3
    b = rng.normal(size=4 * H) * 0.1
    hs, _ = orc.lstm_forward(x, W, U, b)
    m = torch.nn.LSTM(F, H, batch_first=True).double()
    with torch.no_grad():
        m.weight_ih_l0.copy_(torch.from_numpy(W.T))
        m.weight_hh_l0.copy_(torch.from_numpy(U.T))
        m.bias_ih_l0.copy_(torch.from_numpy(b))
        m.bias_hh_l0.zero_()
        ref = m(torch.from_numpy(x))[0].numpy()
    np.testing.assert_allclose(hs, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('L', [1, 2])
def test_backward_matches_torch_autograd(L):
    params, x, y = _rand_problem(L=L)
    preds, fc = orc.forward(params, x, num_layers=L)
    loss, mse, dpred, _ = orc.loss_point_estimate(y, preds, target_idx=1, target_lambda=0.5, rnn_lambda=0.7)
    grads = orc.backward(dpred, fc, num_layers=L)
    tpred, tp = _torch_forward(params, x, L)
    np.testing.assert_allclose(preds, tpred.detach().numpy(), rtol=1e-12, atol=1e-12)
    ty = torch.from_numpy(y)
    mask = (~(ty == 0).all(-1)).double()
    yp = tpred * mask[..., None]
    mse0 = ((ty[:, -1, 1] - yp[:, -1, 1]) ** 2).mean()
    mse1 = ((ty[:, -1] - yp[:, -1]) ** 2).mean()
    mse2 = ((yp - ty) ** 2).sum() / (mask.sum() * y.shape[-1])
    tloss = 0.5 * mse0 + 0.5 * (0.7 * mse1 + 0.3 * mse2)
    assert float(tloss) == pytest.approx(float(loss), rel=1e-12)
    tloss.backward()
    for g, t in zip(grads, tp):
        np.testing.assert_allclose(g, t.grad.numpy(), rtol=1e-9, atol=1e-11)


def test_backward_finite_differences_with_dropout():
    L = 2
    params, x, y = _rand_problem(L=L, B=3, T=4, F=4, H=4, O=2)
    kw = dict(num_layers=L, dropout=0.25, recurrent_dropout=0.25, training=True, seed=77, step=5)
    lk = dict(target_idx=0, target_lambda=0.4, rnn_lambda=0.6)

    def f(ps):
        p, _ = orc.forward(ps, x, **kw)
        return orc.loss_point_estimate(y, p, **lk)[0]

    preds, fc = orc.forward(params, x, **kw)
    grads = orc.backward(orc.loss_point_estimate(y, preds, **lk)[2], fc, num_layers=L)
    rng = np.random.RandomState(0)
    for j, p in enumerate(params):
        for _ in range(4):
            idx = tuple(rng.randint(s) for s in p.shape)
            eps = 1e-6
            pp = [q.copy() for q in params]
            pp[j][idx] += eps
            up = f(pp)
            pp[j][idx] -= 2 * eps
            dn = f(pp)
            assert grads[j][idx] == pytest.approx((up - dn) / (2 * eps), rel=2e-5, abs=1e-8)


def test_optimizers_match_torch():
    rng = np.random.RandomState(1)
    p0 = [rng.normal(size=(4, 3)), rng.normal(size=(5,))]
    gs = [[rng.normal(size=q.shape) for q in p0] for _ in range(4)]
    # Adadelta: torch uses the same recurrences with eps inside the sqrt
    p = [q.copy() for q in p0]
    slots = orc.zero_slots('Adadelta', p)
    tp = [torch.from_numpy(q.copy()).requires_grad_(True) for q in p0]
    opt = torch.optim.Adadelta(tp, lr=0.6, rho=0.95, eps=1e-7)
    for it, g in enumerate(gs):
        p = orc.optimizer_update('Adadelta', p, g, slots, 0.6, it)
        for t, gg in zip(tp, g):
            t.grad = torch.from_numpy(gg.copy())
        opt.step()
    for a, t in zip(p, tp):
        np.testing.assert_allclose(a, t.detach().numpy(), rtol=1e-12)
    # SGD + momentum (keras: m = mom*m - lr*g ; p += m  ==  torch with dampening 0, same lr every step)
    p = [q.copy() for q in p0]
    slots = orc.zero_slots('SGD', p)
    tp = [torch.from_numpy(q.copy()).requires_grad_(True) for q in p0]
    opt = torch.optim.SGD(tp, lr=0.1, momentum=0.9)
    for it, g in enumerate(gs):
        p = orc.optimizer_update('SGD', p, g, slots, 0.1, it, sgd_momentum=0.9)
        for t, gg in zip(tp, g):
            t.grad = torch.from_numpy(gg.copy())
        opt.step()
    for a, t in zip(p, tp):
        np.testing.assert_allclose(a, t.detach().numpy(), rtol=1e-12)


def test_adam_rmsprop_closed_form_first_step():
    g = [np.array([0.5, -2.0])]
    p = [np.array([1.0, 1.0])]
    s = orc.zero_slots('Adam', p)
    new = orc.optimizer_update('Adam', p, g, s, 0.1, 0)
    # step 1: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1) -> p - lr * g/(|g| + eps*sqrt(1-b2)...)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = p[0] - lr_t * (0.1 * g[0]) / (np.sqrt(0.001 * g[0] ** 2) + 1e-7)
    np.testing.assert_allclose(new[0], exp, rtol=1e-12)
    s = orc.zero_slots('RMSprop', p)
    new = orc.optimizer_update('RMSprop', p, g, s, 0.1, 0)
    exp = p[0] - 0.1 * g[0] / (np.sqrt(0.1 * g[0] ** 2) + 1e-7)
    np.testing.assert_allclose(new[0], exp, rtol=1e-12)


def test_clip_and_maxnorm_and_lr():
    g = [np.array([3.0, 0.0]), np.array([[4.0]])]
    c, gn = orc.clip_by_global_norm(g, 2.5)
    assert gn == pytest.approx(5.0)
    np.testing.assert_allclose(c[0], [1.5, 0.0])
    c, _ = orc.clip_by_global_norm(g, 50.0)
    np.testing.assert_allclose(c[1], [[4.0]])
    w = np.array([[3.0, 0.3], [4.0, 0.4]])
    out = orc.max_norm_constraint(w, 3)
    np.testing.assert_allclose(np.linalg.norm(out[:, 0]), 3.0, rtol=1e-6)
    np.testing.assert_allclose(out[:, 1], w[:, 1] * (0.5 / (1e-7 + 0.5)))
    assert orc.learning_rate(1499, learning_rate=0.6, lr_decay=0.5, decay_steps=1500) == pytest.approx(0.6)
    assert orc.learning_rate(1500, learning_rate=0.6, lr_decay=0.5, decay_steps=1500) == pytest.approx(0.3)
    assert orc.learning_rate(750, lr_schedule='PolynomialDecay', learning_rate=0.6, decay_steps=1500,
                             end_learning_rate=0.01, decay_power=0.5) == pytest.approx(0.59 * np.sqrt(0.5) + 0.01)
    kw = dict(lr_schedule='PiecewiseConstantDecay', piecewise_lr_boundaries=[4000, 5500, 5500],
              piecewise_lr_values=[0.5, 0.1, 0.05, 0.1])
    assert orc.learning_rate(4000, **kw) == 0.5 and orc.learning_rate(4001, **kw) == 0.1
    assert orc.learning_rate(6000, **kw) == 0.1


def test_dropout_mask_shard_invariant():
    full = orc.dropout_mask(521, 7, 2, 0, 16, 8 * 12, 0.3)
    lo = orc.dropout_mask(521, 7, 2, 0, 8, 8 * 12, 0.3)
    hi = orc.dropout_mask(521, 7, 2, 8, 8, 8 * 12, 0.3)
    np.testing.assert_array_equal(full, np.concatenate([lo, hi]))
    assert set(np.unique(full)) <= {0.0, np.float64(np.float32(1.0) / (np.float32(1.0) - np.float32(0.3)))}


def test_train_step_dp_equivalence():
    """Two shards with global denominators sum to the single-process gradient."""
    params, x, y = _rand_problem(B=6, L=1)
    lk = dict(target_idx=1, target_lambda=0.5, rnn_lambda=0.7)
    preds, fc = orc.forward(params, x, num_layers=1)
    _, _, dpred, (_, _, _, m) = orc.loss_point_estimate(y, preds, **lk)
    full = orc.backward(dpred, fc, num_layers=1)
    acc = None
    for lo, hi in ((0, 3), (3, 6)):
        p, c = orc.forward(params, x[lo:hi], num_layers=1)
        _, _, dp, _ = orc.loss_point_estimate(y[lo:hi], p, batch_global=6, mask_count_global=m, **lk)
        g = orc.backward(dp, c, num_layers=1)
        acc = g if acc is None else [a + b for a, b in zip(acc, g)]
    for a, b in zip(acc, full):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)


def test_tf_pin_script_degrades_when_tensorflow_is_absent():
    """oracle/pin_with_tf.py is the only route to a pinned oracle.  Without TensorFlow it must say so and
    exit 0; its comparison branch is NOT covered here (never executed: no TensorFlow in this image)."""
    import importlib.util
    import os
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'pin_with_tf.py')
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600)
    if importlib.util.find_spec('tensorflow') is None:
        assert r.returncode == 0 and 'TF oracle unavailable' in r.stdout
    else:
        assert r.returncode == 0 and 'PINNED against tensorflow' in r.stdout, r.stdout + r.stderr


# ---- GRU cell (rnn_point_estimate.py:89-98).  torch.nn.GRU implements the same reset-after equations
# with gate order r|z|n, so it pins the oracle's z|r|h restatement forward and backward.
def _torch_gru(W, U, b2):
    H = U.shape[0]
    perm = np.concatenate([np.arange(H, 2 * H), np.arange(0, H), np.arange(2 * H, 3 * H)])
    m = torch.nn.GRU(W.shape[0], H, batch_first=True).double()
    with torch.no_grad():
        m.weight_ih_l0.copy_(torch.from_numpy(W[:, perm].T.copy()))
        m.weight_hh_l0.copy_(torch.from_numpy(U[:, perm].T.copy()))
        m.bias_ih_l0.copy_(torch.from_numpy(b2[0][perm].copy()))
        m.bias_hh_l0.copy_(torch.from_numpy(b2[1][perm].copy()))
    return m, perm


def test_gru_matches_torch_nn_gru_forward_and_backward():
    B, T, F, H = 6, 9, 5, 12
    rng = np.random.RandomState(2)
    W, U = rng.normal(size=(F, 3 * H)) * 0.4, rng.normal(size=(H, 3 * H)) * 0.4
    b2 = rng.normal(size=(2, 3 * H)) * 0.3
    x = rng.normal(size=(B, T, F))
    dh = rng.normal(size=(B, T, H))
    hs, cache = orc.gru_forward(x, W, U, b2)
    m, perm = _torch_gru(W, U, b2)
    xt = torch.from_numpy(x).requires_grad_(True)
    out = m(xt)[0]
    assert np.abs(hs - out.detach().numpy()).max() < 1e-12
    out.backward(torch.from_numpy(dh))
    dW, dU, db, dx = orc.gru_backward(dh, cache, need_dx=True)
    inv = np.argsort(perm)
    assert np.abs(dW - m.weight_ih_l0.grad.numpy().T[:, inv]).max() < 1e-10
    assert np.abs(dU - m.weight_hh_l0.grad.numpy().T[:, inv]).max() < 1e-10
    assert np.abs(db[0] - m.bias_ih_l0.grad.numpy()[inv]).max() < 1e-10
    assert np.abs(db[1] - m.bias_hh_l0.grad.numpy()[inv]).max() < 1e-10
    assert np.abs(dx - xt.grad.numpy()).max() < 1e-10


def test_gru_stack_finite_differences_with_dropout():
    B, T, F, O, H, L = 4, 5, 6, 3, 8, 2
    rng = np.random.RandomState(4)
    P = orc.init_params(L, F, O, H, seed=7, dtype=np.float64, rnn_cell='gru')
    for l in range(L):
        P[5 * l + 2] = rng.normal(size=(2, 3 * H)) * 0.3
    assert orc.param_names(L, 'gru')[:3] == ['gru_1/kernel', 'gru_1/recurrent_kernel', 'gru_1/bias']
    assert [p.shape for p in P] == [tuple(s) for s in orc.param_shapes(L, F, O, H, 'gru')]
    x = rng.normal(size=(B, T, F))
    y = rng.normal(size=(B, T, O))
    y[0, :2] = 0.0
    kw = dict(num_layers=L, rnn_cell='gru', training=True, dropout=0.25, recurrent_dropout=0.2, seed=9, step=3, row0=40)

    def loss_of(Q):
        pr, fc = orc.forward(Q, x, **kw)
        return orc.loss_point_estimate(y, pr, target_idx=1, target_lambda=0.5, rnn_lambda=0.7), fc

    (l0, _, dp, _), fc = loss_of(P)
    g = orc.backward(dp, fc, num_layers=L, rnn_cell='gru')
    for k in range(len(P)):
        for _ in range(4):
            idx = tuple(rng.randint(s) for s in P[k].shape)
            Qp, Qm = [q.copy() for q in P], [q.copy() for q in P]
            Qp[k][idx] += 1e-6
            Qm[k][idx] -= 1e-6
            fd = (loss_of(Qp)[0][0] - loss_of(Qm)[0][0]) / 2e-6
            assert abs(fd - g[k][idx]) < 1e-6 * max(1.0, abs(fd)), (k, idx)


# ---- RNNUqRangeEstimate: rnn_uq_range_estimate.py:66-110, losses.py:180-284, train.py:201-225 ----
def _uq_problem(seed=0, L=2):
    B, T, F, O, H = 5, 6, 7, 3, 8
    rng = np.random.RandomState(seed)
    P = orc.init_params(L, F, O, H, seed=3, dtype=np.float64, uq=True)
    P[-1] = rng.normal(size=O) * 0.5
    P[-3] = rng.normal(size=O) * 0.5
    return P, rng.normal(size=(B, T, F)), rng.normal(size=(B, T, O)), L, O


def test_uq_loss_and_gradients_match_torch_autograd():
    P, x, y, L, O = _uq_problem()
    assert orc.param_names(L, uq=True)[-4:] == ['OUTPUT_TARGET_1/kernel', 'OUTPUT_TARGET_1/bias',
                                                'OUTPUT_VARIANCE_1/kernel', 'OUTPUT_VARIANCE_1/bias']
    p, v, _ = orc.forward_uq(P, x, num_layers=L, dropout=0.2, recurrent_dropout=0.1, seed=5, step=2, row0=10)
    assert v.min() >= orc.VAR_FLOOR
    loss, uq0, mse0, dp, dv = orc.loss_uq_estimate(y, p, v, target_idx=1, target_lambda=0.5, rnn_lambda=0.7)
    tp, tv, ty = torch.tensor(p, requires_grad=True), torch.tensor(v, requires_grad=True), torch.tensor(y)
    m = (~(ty == 0).all(-1)).double()
    term = (tp * m[..., None] - ty) ** 2 / (tv * m[..., None]) + torch.log(tv * m[..., None])
    t0 = term[:, -1, 1].sum() / m[:, -1].sum()
    t1 = term[:, -1, :].sum() / (m[:, -1].sum() * O)
    t2 = term.sum() / (m.sum() * O)
    lt = 0.5 * t0 + 0.5 * (0.7 * t1 + 0.3 * t2)
    lt.backward()
    assert abs(lt.item() - loss) < 1e-12 and abs(t0.item() - uq0) < 1e-12
    assert mse0 == pytest.approx(np.mean((y[:, -1, 1] - p[:, -1, 1]) ** 2), rel=1e-12)
    assert np.abs(tp.grad.numpy() - dp).max() < 1e-12 and np.abs(tv.grad.numpy() - dv).max() < 1e-12


def test_uq_stack_finite_differences_and_dropout_is_always_on():
    P, x, y, L, O = _uq_problem(seed=1)
    kw = dict(num_layers=L, dropout=0.3, recurrent_dropout=0.2, seed=5, step=2, row0=10)

    def loss_of(Q):
        p, v, fc = orc.forward_uq(Q, x, **kw)
        return orc.loss_uq_estimate(y, p, v, target_idx=1, target_lambda=0.5, rnn_lambda=0.7), fc

    (l0, _, _, dp, dv), fc = loss_of(P)
    g = orc.backward_uq(dp, dv, fc, num_layers=L)
    rng = np.random.RandomState(3)
    for k in range(len(P)):
        for _ in range(4):
            idx = tuple(rng.randint(s) for s in P[k].shape)
            Qp, Qm = [q.copy() for q in P], [q.copy() for q in P]
            Qp[k][idx] += 1e-6
            Qm[k][idx] -= 1e-6
            fd = (loss_of(Qp)[0][0] - loss_of(Qm)[0][0]) / 2e-6
            assert abs(fd - g[k][idx]) < 1e-6 * max(1.0, abs(fd)), (k, idx)
    # training=True is a literal in the reference's UQ model: a different step draws different masks
    p_a = orc.forward_uq(P, x, **kw)[0]
    p_b = orc.forward_uq(P, x, **dict(kw, step=3))[0]
    assert np.abs(p_a - p_b).max() > 1e-3


def test_uq_loss_is_nan_with_a_padded_step_like_the_reference_formula():
    """losses.py:200-201 multiplies the variance by the mask, :272 divides by it and takes its log."""
    P, x, y, L, O = _uq_problem(seed=2)
    p, v, _ = orc.forward_uq(P, x, num_layers=L)
    y[0, :2] = 0.0
    loss, uq0, mse0, dp, dv = orc.loss_uq_estimate(y, p, v, target_idx=1, target_lambda=0.5, rnn_lambda=0.7)
    assert np.isnan(loss) and np.isfinite(uq0) and np.isfinite(mse0)
    assert np.isnan(dp[0, 0]).all() and np.isnan(dv[0, 0]).all()


# ---- forecast_steps > 1 (second half): rnn_point_estimate.py:109-150 -- oracle only so far ----
def test_forecast_chain_window_construction_and_names():
    B, T, F, O, H, L, S = 3, 4, 6, 2, 5, 1, 3
    rng = np.random.RandomState(1)
    P = orc.init_forecast_params(L, F, O, H, S, seed=5, dtype=np.float64)
    names = orc.forecast_param_names(L, S)
    assert len(P) == len(names) == 5 * L + 2 + 7 * (S - 1)
    assert names[5 * L:5 * L + 2] == ['OUTPUT_1/kernel', 'OUTPUT_1/bias']
    assert names[-7:] == ['lstm_3/kernel', 'lstm_3/recurrent_kernel', 'lstm_3/bias', 'batch_normalization_2/gamma',
                          'batch_normalization_2/beta', 'OUTPUT_3/kernel', 'OUTPUT_3/bias']
    x = rng.normal(size=(B, T, F))
    preds, fc = orc.forward_forecast(P, x, num_layers=L, forecast_steps=S)
    assert len(preds) == S and all(p.shape == (B, T, O) for p in preds)
    # the window of step 2: first step dropped, [pred_1[:, -1], aux of the last ORIGINAL step] appended
    win2 = np.concatenate([x[:, 1:], np.concatenate([preds[0][:, -1:, :], x[:, -1:, O:]], axis=2)], axis=1)
    ref2, _ = orc.forward(P[5 * L + 2:5 * L + 2 + 7], win2, num_layers=1)
    assert np.abs(ref2 - preds[1]).max() < 1e-12
    win3 = np.concatenate([win2[:, 1:], np.concatenate([preds[1][:, -1:, :], x[:, -1:, O:]], axis=2)], axis=1)
    ref3, _ = orc.forward(P[5 * L + 2 + 7:], win3, num_layers=1)
    assert np.abs(ref3 - preds[2]).max() < 1e-12
    # forecast_steps = 1 degenerates to the plain model
    p1, _ = orc.forward_forecast(P[:5 * L + 2], x, num_layers=L, forecast_steps=1)
    assert np.abs(p1[0] - preds[0]).max() == 0.0


@pytest.mark.parametrize('cell', ['lstm', 'gru'])
def test_forecast_chain_backward_finite_differences(cell):
    B, T, F, O, H, L, S = 4, 5, 7, 3, 8, 2, 3            # H % 4 == 0: one Philox call yields 4 mask elements
    rng = np.random.RandomState(2)
    P = orc.init_forecast_params(L, F, O, H, S, seed=5, dtype=np.float64, rnn_cell=cell)
    x = rng.normal(size=(B, T, F))
    ys = [rng.normal(size=(B, T, O)) for _ in range(S)]
    ys[1][0, :2] = 0.0
    w = [1.0, 0.6, 0.3]
    kw = dict(target_idx=1, target_lambda=0.5, rnn_lambda=0.7)
    fkw = dict(num_layers=L, forecast_steps=S, rnn_cell=cell, training=True, dropout=0.2, recurrent_dropout=0.1, seed=3,
               step=4, row0=8)

    def loss_of(Q):
        pr, fc = orc.forward_forecast(Q, x, **fkw)
        l, m, dp = orc.loss_forecast(ys, pr, w, **kw)
        return l, dp, fc

    l0, dp, fc = loss_of(P)
    g = orc.backward_forecast(dp, fc, rnn_cell=cell)
    assert len(g) == len(P)
    for k in range(len(P)):
        assert g[k].shape == P[k].shape
        for _ in range(3):
            idx = tuple(rng.randint(s) for s in P[k].shape)
            Qp, Qm = [q.copy() for q in P], [q.copy() for q in P]
            Qp[k][idx] += 1e-6
            Qm[k][idx] -= 1e-6
            fd = (loss_of(Qp)[0] - loss_of(Qm)[0]) / 2e-6
            assert abs(fd - g[k][idx]) < 1e-6 * max(1.0, abs(fd)), (k, idx)
