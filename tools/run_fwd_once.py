"""One bf16 forward (and optionally one train step) at B=4096, T=48 -- the ncu target."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import lfm_oracle as orc  # noqa: E402
from lfm_quant_b200.engine import ForecasterEngine  # noqa: E402

train = len(sys.argv) > 1 and sys.argv[1] == 'train'
T, F, O, H, B = 48, 32, 16, 256, 4096
rng = np.random.default_rng(0)
eng = ForecasterEngine(max_batch=B, seq_len=T, n_inputs=F, n_outputs=O, num_hidden=H, target_idx=3, precision='bf16',
                       train=train, forward_only=not train)
eng.set_weights(orc.init_params(1, F, O, H, init_scale=1.0, seed=521, dtype=np.float32))
x = torch.from_numpy(rng.standard_normal((B, T, F), dtype=np.float32)).cuda()
y = torch.from_numpy(rng.standard_normal((B, T, O), dtype=np.float32)).cuda()
for i in range(3):
    if train:
        eng.train_step(x, y, i, 0.6)
    else:
        eng.forward(x)
torch.cuda.synchronize()
print('done')
