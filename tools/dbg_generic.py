"""Minimal reproducer for the general path: one backward at a small shape (run with LFMQ_DEBUG_SYNC=1)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from util import make_engine, make_problem  # noqa: E402

B, T, F, O, H, L = (int(v) for v in (sys.argv[1:7] if len(sys.argv) > 6 else (200, 5, 32, 16, 64, 1)))
params, x, y = make_problem(B, T, F, O, H, L, seed=21, init_scale=0.3)
print('creating engine', flush=True)
eng = make_engine(B, T, F, O, H, L, target_idx=O - 1, precision='bf16', train=True)
eng.set_weights(params)
print('backward', flush=True)
eng.backward(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), step=5, row0=512)
torch.cuda.synchronize()
print('done; grad norm', float(eng.grads[:eng.n_trainable].norm()), flush=True)
