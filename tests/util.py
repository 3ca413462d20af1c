"""Shared helpers for the parity tests (oracle = checker only)."""
import numpy as np

import lfm_oracle as orc


def make_problem(B, T, F, O, H, L, seed=0, zero_rows=True, init_scale=0.5, rnn_cell='lstm'):
    rng = np.random.RandomState(seed)
    params = orc.init_params(L, F, O, H, init_scale=init_scale, seed=seed + 1, dtype=np.float64, rnn_cell=rnn_cell)
    for l in range(L):
        params[5 * l + 2] = params[5 * l + 2] + rng.normal(size=params[5 * l + 2].shape) * 0.1
        params[5 * l + 3] = params[5 * l + 3] + rng.normal(size=H) * 0.1
        params[5 * l + 4] = params[5 * l + 4] + rng.normal(size=H) * 0.1
    params[5 * L + 1] = params[5 * L + 1] + rng.normal(size=O) * 0.1
    params = [p.astype(np.float32).astype(np.float64) for p in params]   # exactly representable in fp32
    x = rng.normal(size=(B, T, F)).astype(np.float32)
    y = rng.normal(size=(B, T, O)).astype(np.float32)
    if zero_rows and B > 1 and T > 2:
        y[0, :2, :] = 0.0       # zero-padded steps -> loss mask 0 (losses.py:72)
        y[B - 1, 0, :] = 0.0
    return params, x, y


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def make_engine(B, T, F, O, H, L, **kw):
    from lfm_quant_b200.engine import ForecasterEngine
    return ForecasterEngine(max_batch=B, seq_len=T, n_inputs=F, n_outputs=O, num_hidden=H, num_layers=L, **kw)
