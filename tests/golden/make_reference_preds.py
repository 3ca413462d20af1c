"""Generates tests/golden/reference_preds.dat by running the UNMODIFIED reference prediction driver
(/root/reference/scripts/predict.py: Predict.__init__ / predict, on the reference Dataset) with a STUB MODEL.

Only the Keras model is replaced: `model_utils.model.Model(config, dataset).get_model()` returns an object whose
predict(inp) is a fixed function of the input batch (below) and whose load_weights is a no-op.  Everything else that runs
is the reference's: batching, extraction of the last step / target field, un-scaling, reverse log-squash, seq-norm, the
error columns and the file format of preds.dat (SURVEY 8b "Files").  TensorFlow is replaced by the array-wrapping import
shim of make_reference_batcher.py.  Only runs where /root/reference exists; the .dat file is what travels.

usage: python tests/golden/make_reference_preds.py      (after make_reference_batcher.py: same synthetic table)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CHILD = r'''
import os, sys, types, json, random
import numpy as np

tf = types.ModuleType('tensorflow')
tf.float32 = np.float32
class _T(np.ndarray):                      # what get_batch returns: callers use it as an array and call .numpy()
    def numpy(self): return np.asarray(self)
def _bytes(a):                             # TensorFlow string tensors come back from .numpy() as bytes
    a = np.asarray(a)
    return np.char.encode(a.astype('U'), 'utf-8') if a.dtype.kind in 'UO' else a
tf.convert_to_tensor = lambda a, dtype=None: np.asarray(_bytes(a), dtype=dtype).view(_T)
class _DS(object):
    def __init__(self, items): self.items = list(items)
    def __iter__(self): return iter(self.items)
    def __len__(self): return len(self.items)
    @staticmethod
    def from_tensor_slices(a): return _DS([_Eager(np.asarray(x)) for x in _bytes(a)])
    @staticmethod
    def zip(parts): return _DS(list(zip(*[list(p) for p in parts])))
    def batch(self, batch_size):          # tf.data batching of (X, Y, meta) triples
        out = []
        for s in range(0, len(self.items), batch_size):
            chunk = self.items[s:s + batch_size]
            out.append(tuple(_Eager(np.stack([np.asarray(c[k].numpy()) for c in chunk])) for k in range(3)))
        return _DS(out)
class _Eager(object):
    def __init__(self, a): self.a = a
    def numpy(self): return self.a
tf.data = types.SimpleNamespace(Dataset=_DS)
tf.config = types.SimpleNamespace(experimental=types.SimpleNamespace(list_physical_devices=lambda *_: [],
                                                                    set_memory_growth=lambda *_: None))
sys.modules['tensorflow'] = tf

def stub_predict(inp):
    """Fixed function of the input batch [B,T,F] -> [B,T,O]; O = 16 financial fields."""
    x = np.asarray(inp, dtype=np.float64)
    base = np.tanh(x.mean(axis=2, keepdims=True))
    return (base + 0.1 * np.arange(16)[None, None, :] * x[:, :, :1]).astype(np.float32)

class _StubModel(object):
    def predict(self, inp): return stub_predict(inp.numpy() if hasattr(inp, 'numpy') else inp)
    def load_weights(self, path): return None
    def summary(self): return 'stub'
mu = types.ModuleType('model_utils'); mm = types.ModuleType('model_utils.model')
class Model(object):
    def __init__(self, config, dataset): pass
    def get_model(self): return _StubModel()
mm.Model = Model; mu.model = mm
sys.modules['model_utils'] = mu; sys.modules['model_utils.model'] = mm

ROOT, OUT, STAGE = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
from lfm_quant_b200.scripts.synthetic import write_open_dataset
work = os.path.join(os.path.dirname(OUT), 'preds.work')
os.makedirs(os.path.join(work, 'datasets'), exist_ok=True)
write_open_dataset(os.path.join(work, 'datasets', 'open-dataset.dat'), n_keys=8, n_months=120, seed=11)
sys.path.insert(0, '/root/reference/scripts')
base = ['x', '--datafile', 'open-dataset.dat', '--data_dir', os.path.join(work, 'datasets'),
        '--experiments_dir', os.path.join(work, 'experiments'), '--model_dir', 'm',
        '--financial_fields', 'saleq_ttm-ltq_mrq', '--aux_fields', 'rel_mom1m-rel_mom9m',
        '--target_field', 'oiadpq_ttm', '--scale_field', 'mrkcap', '--stride', '12', '--forecast_n', '12',
        '--min_unrollings', '3', '--max_unrollings', '5', '--start_date', '197001', '--end_date', '209912',
        '--validation_size', '0.3', '--seed', '521', '--batch_size', '64']
os.makedirs(os.path.join(work, 'experiments', 'm', 'chkpts'), exist_ok=True)
import data_processing as refdp
if STAGE == 'train':                      # writes scales.dat (the reference's flag parser is global: one stage per process)
    sys.argv = base + ['--train=True']
    random.seed(20260921)                 # same harness seed as make_reference_batcher.py -> same scales.dat
    D = refdp.Dataset(refdp.get_configs()); D.generate_dataset()
    print('ok train')
    sys.exit(0)
sys.argv = base + ['--train=False']
import predict as refpredict
c = refdp.get_configs()
P = refpredict.Predict(c, refdp.Dataset(c))
P.predict()
src = os.path.join(work, 'experiments', 'm', 'pred', c.preds_fname)
import shutil
shutil.copyfile(src, OUT)
print('ok', open(OUT).readline().strip()[:300], sum(1 for _ in open(OUT)))
'''


def main():
    dst = os.path.join(HERE, 'reference_preds.dat')
    r = subprocess.run([sys.executable, '-c', CHILD, ROOT, dst, 'train'], capture_output=True, text=True)
    print('train stage', r.returncode, (r.stderr.strip().splitlines() or [''])[-1][:300] if r.returncode else '')
    r = subprocess.run([sys.executable, '-c', CHILD, ROOT, dst, 'predict'], capture_output=True, text=True)
    print(r.returncode)
    print('\n'.join((r.stdout.strip().splitlines() or [''])[-3:])[:900])
    if r.returncode != 0:
        print('\n'.join(r.stderr.strip().splitlines()[-14:]))
    shutil.rmtree(os.path.join(HERE, 'preds.work'), ignore_errors=True)


if __name__ == '__main__':
    main()
