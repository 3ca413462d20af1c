"""Synthetic stand-in for the reference's unshipped ``open-dataset.dat`` + ``config/system-test.conf``
(README.md:33-36 names them; neither is in the repository -- SURVEY section 4).

File contract (scripts/dataset_api.md:49-56, data_processing.py:40,90): space-separated, header
``date gvkey active <fin fields> mrkcap <aux fields>``, one row per (company, month), rows of a company
contiguous and date-ordered, dates as YYYYMM.
"""
from __future__ import absolute_import, division, print_function

import os

import numpy as np

FIN_FIELDS = ['saleq_ttm', 'cogsq_ttm', 'xsgaq_ttm', 'oiadpq_ttm', 'mkvaltq_ttm', 'niq_ttm', 'cheq_mrq', 'rectq_mrq',
              'invtq_mrq', 'acoq_mrq', 'ppentq_mrq', 'aoq_mrq', 'dlcq_mrq', 'apq_mrq', 'txpq_mrq', 'ltq_mrq']
AUX_FIELDS = ['rel_mom1m', 'rel_mom3m', 'rel_mom6m', 'mom1m', 'mom3m', 'mom6m', 'mom9m', 'vol1m', 'vol3m', 'vol6m',
              'vol9m', 'beta1y', 'ep_rank', 'bp_rank', 'sp_rank', 'rel_mom9m']


def write_open_dataset(path, n_keys=40, start=(1970, 1), n_months=420, seed=521):
    """Writes a dataset of ``n_keys`` companies x up to ``n_months`` months; returns the number of rows."""
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    rows = 0
    with open(path, 'w') as fh:
        fh.write(' '.join(['date', 'gvkey', 'active'] + FIN_FIELDS + ['mrkcap'] + AUX_FIELDS) + '\n')
        for k in range(n_keys):
            first = rng.randint(0, n_months // 4)
            length = rng.randint(n_months // 2, n_months - first + 1)
            mc = np.exp(rng.normal(6.0, 1.5)) * np.exp(np.cumsum(rng.normal(0.004, 0.05, size=length)))
            base = rng.normal(0.0, 0.4, size=len(FIN_FIELDS))
            fin = (base + rng.normal(0, 0.05, size=(length, len(FIN_FIELDS))).cumsum(axis=0) * 0.2) * mc[:, None]
            aux = rng.normal(0, 1, size=(length, len(AUX_FIELDS)))
            active = (rng.uniform(size=length) > 0.03).astype(int)
            for i in range(length):
                m = start[1] - 1 + first + i
                date = (start[0] + m // 12) * 100 + (m % 12) + 1
                vals = ['%d' % date, '%06d' % (1000 + k), '%d' % active[i]]
                vals += ['%.6g' % v for v in fin[i]] + ['%.6g' % mc[i]] + ['%.6g' % v for v in aux[i]]
                fh.write(' '.join(vals) + '\n')
                rows += 1
    return rows


SYSTEM_TEST_CONF = """--name system-test
--datafile open-dataset.dat
--data_dir {data_dir}
--experiments_dir {experiments_dir}
--model_dir system-test-model
--nn_type RNNPointEstimate
--rnn_cell lstm
--financial_fields saleq_ttm-ltq_mrq
--aux_fields rel_mom1m-rel_mom9m
--target_field oiadpq_ttm
--scale_field mrkcap
--num_layers 1
--num_hidden 64
--batch_size 32
--min_unrollings 20
--max_unrollings 20
--stride 12
--forecast_n 12
--start_date 197501
--end_date 200412
--max_epoch 2
--early_stop 5
--optimizer Adadelta
--learning_rate 0.6
--logging_interval 10
--seed 521
"""


def write_system_test_conf(path, data_dir, experiments_dir):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'w') as fh:
        fh.write(SYSTEM_TEST_CONF.format(data_dir=data_dir, experiments_dir=experiments_dir))
