"""GPU parity at the shapes the benchmark runs (VERDICT r1 "weak" #1): T = 48 everywhere.

* BASELINE configs[1] family (H=256, F=32, O=16, L=1): gradients and a 3-step Adadelta trajectory against the fp64
  oracle at B=256 and a ragged B=300, on the fp32 path (1e-4) and on the bf16 tensor-core path (tolerances below,
  per tensor);
* BASELINE configs[2] family (H=512, L=2, dropout 0.2 + recurrent dropout 0.1): gradients against the oracle at
  B=128, on the fp32 path (1e-4) and on the tensor-core path (bf16 tolerances);
* a 50-step loss trajectory of the bf16 path against the fp32 path on the bench's own synthetic batches, with the bound
  the bench line's `loss_check` is held to.

Tolerances (stated, asserted):
  fp32 path:  every tensor max-norm relative error <= 1e-4 (5e-4 on updated weights after 3 steps)
  bf16 path:  loss / mse_0 relative 3e-2; every gradient tensor cosine >= 0.9995 and max-norm relative error <= 5e-2
              at T=48 (measured on B200, profiles/r02_summary.md: worst tensor 1.6e-2 / cosine 0.99994 on the cluster
              kernels, 6.8e-3 / 0.99998 on the general path at H=512, L=2 with both dropouts)
"""
import numpy as np
import pytest
import torch

import lfm_oracle as orc
from util import make_engine, make_problem, rel_err

pytestmark = pytest.mark.gpu

TOL32 = 1e-4
BF16_LOSS, BF16_COS, BF16_REL = 3e-2, 0.9995, 5e-2


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _oracle_grads(params, x, y, target_idx, L, **kw):
    preds, fc = orc.forward(params, x.astype(np.float64), num_layers=L, **kw)
    loss, mse, dpred, _ = orc.loss_point_estimate(y.astype(np.float64), preds, target_idx=target_idx,
                                                  target_lambda=0.5, rnn_lambda=0.7)
    return loss, mse, orc.backward(dpred, fc, num_layers=L)


def _check_grads(eng, ref, prec, what):
    worst = (0.0, 1.0, '')
    for (name, _, _, _), g, r in zip(eng.trainable_specs, eng.grads_list(), ref):
        assert np.isfinite(g).all(), name
        e = rel_err(g, r)
        cos = float(np.sum(g * r) / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
        if e > worst[0]:
            worst = (e, cos, name)
        if prec == 'fp32':
            assert e < TOL32, (what, name, e)
        else:
            assert cos > BF16_COS, (what, name, cos)
            assert e < BF16_REL, (what, name, e)
    print('%s [%s]: worst gradient tensor %s rel %.3e cos %.6f' % (what, prec, worst[2], worst[0], worst[1]))


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('B', [256, 300])
def test_t48_gradients_and_adadelta_steps_match_oracle(prec, B):
    T, F, O, H, L = 48, 32, 16, 256, 1
    params, x, y = make_problem(B, T, F, O, H, L, seed=41 + B, init_scale=1.0)     # bench init: W ~ U(-1, 1)
    eng = make_engine(B, T, F, O, H, L, target_idx=3, precision=prec, optimizer='Adadelta')
    eng.set_weights(params)
    xc, yc = _cuda(x), _cuda(y)
    eng.backward(xc, yc)
    tail = eng.grads[eng.n_trainable:eng.n_trainable + 2].cpu().numpy()
    loss, mse, ref = _oracle_grads(params, x, y, 3, L)
    ltol = TOL32 if prec == 'fp32' else BF16_LOSS
    assert tail[0] == pytest.approx(loss, rel=ltol)
    assert tail[1] == pytest.approx(mse, rel=ltol)
    _check_grads(eng, ref, prec, 'cfg2-family B=%d T=48' % B)
    # three full train steps (clip 50, Adadelta lr 0.6, MaxNorm 3 -- the bench's step)
    cfg = dict(num_layers=L, target_idx=3, target_lambda=0.5, rnn_lambda=0.7, max_grad_norm=50.0, optimizer='Adadelta',
               max_norm=3.0, train=True, dropout=0.0, recurrent_dropout=0.0)
    p = [q.copy() for q in params]
    slots = orc.zero_slots('Adadelta', p)
    for it in range(3):
        out = eng.train_step(xc, yc, it, 0.6).cpu().numpy()
        p, mse_o, loss_o, _, _ = orc.train_step(p, slots, x.astype(np.float64), y.astype(np.float64), it, cfg, lr=0.6)
        assert out[0] == pytest.approx(loss_o, rel=ltol), it
        assert out[1] == pytest.approx(mse_o, rel=ltol), it
    wtol = 5 * TOL32 if prec == 'fp32' else BF16_LOSS
    for (name, _, _, _), w, r in zip(eng.trainable_specs, eng.get_weights(), p):
        assert rel_err(w, r) < wtol, name
    eng.close()


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_cfg3_family_gradients_match_oracle(prec):
    """H=512, L=2, dropout 0.2 + recurrent dropout 0.1, T=48 (BASELINE configs[2] at a batch the oracle can run)."""
    B, T, F, O, H, L = 128, 48, 32, 16, 512, 2
    kw = dict(dropout=0.2, recurrent_dropout=0.1, seed=521)
    params, x, y = make_problem(B, T, F, O, H, L, seed=77, init_scale=0.1)
    eng = make_engine(B, T, F, O, H, L, target_idx=3, precision=prec, train=True, **kw)
    eng.set_weights(params)
    eng.backward(_cuda(x), _cuda(y), step=7, row0=4096)
    tail = eng.grads[eng.n_trainable:eng.n_trainable + 2].cpu().numpy()
    loss, mse, ref = _oracle_grads(params, x, y, 3, L, training=True, step=7, row0=4096, **kw)
    ltol = TOL32 if prec == 'fp32' else BF16_LOSS
    assert tail[0] == pytest.approx(loss, rel=ltol)
    assert tail[1] == pytest.approx(mse, rel=ltol)
    _check_grads(eng, ref, prec, 'cfg3-family B=128 T=48')
    eng.close()


TRAJ_STEPS, TRAJ_BOUND = 50, 2e-3      # measured: 2.2e-4 (profiles/r02_summary.md)


def test_bf16_loss_trajectory_tracks_fp32_on_bench_batches():
    """50 train steps on the bench's synthetic batches (B=4096, T=48, 4 rotating batches, Adadelta lr 0.6): the bf16
    tensor-core path's {loss, mse_0} stays within TRAJ_BOUND (relative) of the fp32 path at every step.  bench.py holds
    its own final_loss_mse to the same bound (`loss_check`, 2e-3)."""
    import bench
    rng = np.random.default_rng(bench.SEED)
    host = [bench.synthetic(4096, rng) for _ in range(4)]
    traj = {}
    for prec in ('fp32', 'bf16'):
        eng = make_engine(4096, bench.T, bench.F, bench.O, bench.H, bench.L, target_idx=bench.TARGET_IDX,
                          precision=prec, optimizer='Adadelta', seed=bench.SEED)
        eng.set_weights(bench.initial_weights())
        dev = [(_cuda(a), _cuda(b)) for a, b in host]
        outs = [eng.train_step(dev[i % 4][0], dev[i % 4][1], i, 0.6) for i in range(TRAJ_STEPS)]
        traj[prec] = torch.stack(outs).cpu().numpy()
        eng.close()
        del eng, dev
        torch.cuda.empty_cache()
    d = np.abs(traj['bf16'] - traj['fp32']) / np.abs(traj['fp32'])
    print('bf16 vs fp32 trajectory: max rel diff loss %.3e mse %.3e (step %d)' % (d[:, 0].max(), d[:, 1].max(),
                                                                                  int(d.max(axis=1).argmax())))
    assert np.isfinite(traj['bf16']).all()
    assert d.max() < TRAJ_BOUND
    assert traj['fp32'][-1, 0] < traj['fp32'][0, 0]          # and it actually trains
