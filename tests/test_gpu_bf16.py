"""GPU tests of the bf16 tensor-core path (LFMQ_PREC_BF16): tcgen05 gate GEMMs with bf16 operands, fp32 accumulate.

bf16 operands carry 8 mantissa bits, so these tests use a bf16-level tolerance against the fp64 oracle (the fp32
parity mode is held to 1e-4 in test_gpu_parity.py) and additionally check the tensor-core path against the fp32
CUDA path at the BASELINE sizes.
"""
import numpy as np
import pytest
import torch

import lfm_oracle as orc
from util import make_engine, make_problem, rel_err

pytestmark = pytest.mark.gpu

BF16_TOL = 3e-2


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('B,T', [(300, 6), (128, 1), (700, 3)])
def test_bf16_forward_matches_oracle(B, T):
    F, O, H, L = 32, 16, 256, 1
    params, x, y = make_problem(B, T, F, O, H, L, seed=11, init_scale=0.3)
    eng = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='bf16')
    eng.set_weights(params)
    preds = eng.forward(_cuda(x)).cpu().numpy()
    ref, _ = orc.forward(params, x.astype(np.float64), num_layers=L)
    assert np.isfinite(preds).all()
    assert rel_err(preds, ref) < BF16_TOL
    # ragged call on the same handle
    p2 = eng.forward(_cuda(x[:B - 37])).cpu().numpy()
    assert rel_err(p2, ref[:B - 37]) < BF16_TOL


def test_bf16_forward_small_inputs_and_outputs():
    B, T, F, O, H, L = 130, 4, 20, 7, 256, 1
    params, x, y = make_problem(B, T, F, O, H, L, seed=12, init_scale=0.3)
    eng = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='bf16')
    eng.set_weights(params)
    preds = eng.forward(_cuda(x)).cpu().numpy()
    ref, _ = orc.forward(params, x.astype(np.float64), num_layers=L)
    assert rel_err(preds, ref) < BF16_TOL


def test_bf16_unsupported_shapes_fail_loudly():
    from lfm_quant_b200._native import LfmqError
    with pytest.raises(LfmqError):
        make_engine(8, 4, 32, 16, 64, 1, precision='bf16')      # H != 256
    with pytest.raises(LfmqError):
        make_engine(8, 4, 32, 16, 256, 2, precision='bf16')     # L != 1


def test_bf16_matches_fp32_path_at_baseline_shape():
    """BASELINE cfg2 shape (B=4096, T=48, F=32, H=256): tensor-core forward vs the fp32 CUDA path."""
    B, T, F, O, H, L = 4096, 48, 32, 16, 256, 1
    params, x, y = make_problem(B, T, F, O, H, L, seed=13, init_scale=1.0, zero_rows=False)
    e32 = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='fp32')
    e16 = make_engine(B, T, F, O, H, L, train=False, forward_only=True, precision='bf16')
    e32.set_weights(params)
    e16.set_weights(params)
    xc = _cuda(x)
    p32 = e32.forward(xc).cpu().numpy()
    p16 = e16.forward(xc).cpu().numpy()
    assert np.isfinite(p16).all()
    # error relative to the output scale, and its RMS
    scale = np.abs(p32).max()
    assert np.abs(p16 - p32).max() / scale < 6e-2
    assert np.sqrt(np.mean((p16 - p32) ** 2)) / np.sqrt(np.mean(p32 ** 2)) < 2e-2
