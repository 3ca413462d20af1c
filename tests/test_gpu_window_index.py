"""Window index on the device (reference scripts/data_processing.py:170-305) through lfmq_window_index:
bit-exact against the oracle's row-by-row restatement, and -- through Dataset -- against the vectors the unmodified
reference produced (tests/golden/reference_batcher_train.npz)."""
import os

import numpy as np
import pytest
import torch

import lfm_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _table(n_keys, seed, min_len=1, max_len=90):
    rng = np.random.RandomState(seed)
    keys, dates, active = [], [], []
    for k in range(n_keys):
        L = int(rng.randint(min_len, max_len + 1))
        y0, m0 = 1975 + int(rng.randint(0, 30)), int(rng.randint(1, 13))
        for j in range(L):
            mm = m0 - 1 + j
            keys.append(k * 7 + 3)
            dates.append((y0 + mm // 12) * 10000 + (mm % 12 + 1) * 100 + 28)
            active.append(int(rng.rand() > 0.15))
    return np.asarray(keys, np.int32), np.asarray(active, np.uint8), np.asarray(dates, np.int32)


def _device(keys, active, dates, **kw):
    from lfm_quant_b200.engine import window_index
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    inp, tar, rows = window_index(cu(keys), cu(active), cu(dates), **kw)
    return inp.cpu().numpy(), tar.cpu().numpy(), rows.cpu().numpy()


@pytest.mark.parametrize('train', [True, False])
@pytest.mark.parametrize('n_keys,stride,fn,mn,mx', [(40, 1, 1, 1, 1), (300, 3, 3, 2, 5), (900, 12, 12, 3, 5),
                                                    (25000, 3, 12, 4, 20)])
def test_window_index_matches_oracle(train, n_keys, stride, fn, mn, mx):
    keys, active, dates = _table(n_keys, seed=n_keys + stride)
    kw = dict(train=train, stride=stride, forecast_n=fn, min_unrollings=mn, max_unrollings=mx,
              start_date=19800101, end_date=20051231, last_train_date=20041231)
    ri, rt, rr = orc.create_window_index(keys.astype(str), active, dates, **kw)
    di, dt, dr = _device(keys, active, dates, **kw)
    assert len(rr) > 0 or n_keys < 100
    np.testing.assert_array_equal(dr, rr)
    np.testing.assert_array_equal(di, ri)
    np.testing.assert_array_equal(dt, rt)


def test_window_index_edge_cases():
    kw = dict(train=True, stride=2, forecast_n=2, min_unrollings=2, max_unrollings=3, start_date=0, end_date=99999999,
              last_train_date=99999999)
    # one row; one long run spanning several 1024-row blocks; a run that starts exactly on a block boundary
    for keys in (np.zeros(1, np.int32), np.zeros(5000, np.int32),
                 np.concatenate([np.zeros(1024, np.int32), np.ones(1024, np.int32), np.full(7, 2, np.int32)])):
        n = keys.size
        active = np.ones(n, np.uint8)
        dates = np.arange(n, dtype=np.int32) + 1
        ri, rt, rr = orc.create_window_index(keys.astype(str), active, dates, **kw)
        di, dt, dr = _device(keys, active, dates, **kw)
        np.testing.assert_array_equal(dr, rr)
        np.testing.assert_array_equal(di, ri)
        np.testing.assert_array_equal(dt, rt)
    # nothing qualifies -> empty index
    di, dt, dr = _device(np.zeros(10, np.int32), np.zeros(10, np.uint8), np.arange(10, dtype=np.int32), **kw)
    assert di.shape == (0, 3) and dt.shape == (0, 3) and dr.shape == (0,)


def test_window_index_more_than_1024_blocks():
    """> 1024 blocks of 1024 rows: the scan over the block aggregates takes more than one chunk."""
    keys, active, dates = _table(26000, seed=5, min_len=20, max_len=70)
    assert keys.size > 1024 * 1024
    kw = dict(train=True, stride=3, forecast_n=3, min_unrollings=3, max_unrollings=8, start_date=19800101,
              end_date=20051231, last_train_date=20041231)
    ri, rt, rr = orc.create_window_index(keys.astype(str), active, dates, **kw)
    di, dt, dr = _device(keys, active, dates, **kw)
    np.testing.assert_array_equal(dr, rr)
    np.testing.assert_array_equal(di, ri)
    np.testing.assert_array_equal(dt, rt)


def test_dataset_builds_its_index_on_the_device_and_matches_the_reference_golden(tmp_path, monkeypatch):
    from test_golden_batcher import _mirror
    from lfm_quant_b200.scripts import configs
    from lfm_quant_b200.scripts.data_processing import Dataset
    calls = []
    real = Dataset._create_index_device
    monkeypatch.setattr(Dataset, '_create_index_device', lambda self, *a: calls.append(1) or real(self, *a))
    g = np.load(os.path.join(GOLD, 'reference_batcher_train.npz'))
    try:
        c, ds = _mirror(tmp_path, True)
        assert calls, 'the device path did not run'
        for k in ('train_X', 'train_Y', 'valid_X', 'valid_Y'):
            np.testing.assert_array_equal(np.asarray(ds._dataset[k]), g['ds_' + k], err_msg=k)
        monkeypatch.setenv('LFMQ_HOST_INDEX', '1')                 # and the host form gives the same triples
        c2, ds2 = _mirror(tmp_path, True)
        for k in ('train_X', 'train_Y', 'valid_X', 'valid_Y'):
            np.testing.assert_array_equal(np.asarray(ds2._dataset[k]), np.asarray(ds._dataset[k]), err_msg=k)
    finally:
        configs.reset()
