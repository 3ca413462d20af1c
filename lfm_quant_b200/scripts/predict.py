"""Predict driver (reference: scripts/predict.py:27-300): load weights, forward every test batch, keep
``pred[:, -1, target_index]``, undo the scaling, write ``<model_dir>/pred/<preds_fname>`` with the reference's
column order (predict.py:113-126,148-157,174-182,214-259; SURVEY section 8b "Files")."""
from __future__ import absolute_import, division, print_function

import os
import pathlib
from collections import defaultdict

import numpy as np
import pandas as pd

from .model_utils.model import Model


class Predict(object):

    def __init__(self, config, dataset):
        self.config = config
        assert not self.config.train, 'Predict can only be instantiated when config.train is False'
        self.dataset = dataset
        self.dataset.generate_dataset()
        self.model = Model(self.config, self.dataset).get_model()
        self.target_index = self.dataset.target_index
        self.seq_len = self.dataset.seq_len
        self.n_inputs = self.dataset.n_inputs
        self.test_set = self.dataset.test_set.batch(batch_size=self.config.batch_size)
        self.model_dir = os.path.join(self.config.experiments_dir, self.config.model_dir)
        self.chkpts_dir = os.path.join(self.model_dir, 'chkpts')
        if not os.path.isdir(self.model_dir):
            os.mkdir(self.model_dir)
        print('creating preds directory')
        self.pred_dir = os.path.join(self.model_dir, 'pred')
        if 'Naive' not in self.config.nn_type:
            assert os.path.isdir(self.chkpts_dir), 'No checkpoint dir found to load the model'
        if not os.path.isdir(self.pred_dir):
            pathlib.Path(self.pred_dir).mkdir(parents=True, exist_ok=True)
        print("Creating batches ...")
        self._batches = [self.dataset.get_batch(*items) for items in self.test_set]

    def predict(self):
        self.model.load_weights(os.path.join(self.chkpts_dir, "chkpt"))
        outputs = defaultdict(list)
        for (inp_dev, target_dev, metadata) in self._batches:
            inp, targets = inp_dev.cpu().numpy(), [target_dev.cpu().numpy()]
            outputs['date'].append(np.expand_dims(metadata[:, 0].astype('int32'), -1))
            outputs['gvkey'].append(np.expand_dims(metadata[:, 1], -1))
            outputs['seq_norm'].append(np.expand_dims(metadata[:, 2].astype('float32'), -1))
            if self.config.write_inp_to_out_file:
                for i, t_step in enumerate(range(1 - self.seq_len, 1)):
                    outputs['inp_t' + str(t_step)].append(self._extract_inputs(inp, i))
            outputs['targets'].append(self._extract_targets(targets))
            if not self.config.UQ:
                preds = [self.model.predict(inp_dev)]
                variance = [np.zeros(x.shape) for x in preds]   # predict.py:133
            else:                                               # predict.py:135-138
                model_preds = self.model.predict(inp_dev)
                preds, variance = model_preds[0::2], model_preds[1::2]
            outputs['norm_preds'].append(self._extract_preds(preds))
            outputs['norm_variance'].append(self._extract_preds(variance))

        single_outputs = ['date', 'gvkey', 'seq_norm'] + [x for x in outputs.keys() if 'inp_t' in x]
        for key in single_outputs:
            outputs[key] = np.vstack(outputs[key])
        for key in ['targets', 'norm_preds', 'norm_variance']:
            for i in range(self.config.forecast_steps):
                outputs[key + '_' + str(i + 1)] = np.vstack([x[i] for x in outputs[key]])
            outputs.pop(key)
        outputs = {k: v.flatten() for k, v in outputs.items()}

        df = pd.DataFrame.from_dict(outputs)
        try:
            df['date'] = pd.to_datetime(df['date'].astype(str), format="%Y%m%d")
        except ValueError:
            print("Input date is not in the '%Y%m%d' format")
            raise
        scale = self.dataset.scaling_params['scale'][self.target_index]
        center = self.dataset.scaling_params['center'][self.target_index]

        def unscale(col):
            """reverse centre/scale, log squash and seq_norm (predict.py:184-247)."""
            v = np.multiply(np.asarray(col, dtype=np.float64), scale) + center
            if self.config.log_squasher:
                v = self.dataset.reverse_log_squasher(v)
            return v * df['seq_norm'].values.astype(np.float64)

        tar_cols = [x for x in df.columns if 'targets' in x]
        for tar_c in tar_cols:
            step = tar_c.split('_')[-1]
            df['norm_' + tar_c] = df[tar_c]
            df['norm_squared_diff_' + step] = np.square(df['norm_' + tar_c] - df['norm_preds_' + step])
        for tar_c in tar_cols:
            df[tar_c] = unscale(df[tar_c])
        if self.config.write_inp_to_out_file:
            for c in [x for x in df.columns if 'inp_t' in x]:
                df[c] = unscale(df[c])
        preds_keys = [x for x in df.columns if 'norm_preds_' in x]
        for pred_key in preds_keys:
            df['preds_' + pred_key.split('_')[-1]] = unscale(df[pred_key])
        for var_key in [x for x in df.columns if 'norm_variance_' in x]:
            df['variance_' + var_key.split('_')[-1]] = unscale(df[var_key])     # not zero: SURVEY App. B #8
        for o_key in [x.split('_')[-1] for x in preds_keys]:
            df['fcst_err_' + o_key] = ((df['targets_' + o_key] - df['preds_' + o_key]) / df['targets_' + o_key]).abs()
            df['abs_err_' + o_key] = (df['targets_' + o_key] - df['preds_' + o_key]).abs()
            df['unscaled_squared_err_' + o_key] = np.square(df['abs_err_' + o_key] / df['seq_norm'])

        df['gvkey'] = df['gvkey'].apply(lambda x: x.decode('utf-8') if isinstance(x, bytes) else str(x))
        df.to_csv(os.path.join(self.pred_dir, self.config.preds_fname), sep=' ', index=False, date_format="%Y%m%d")
        print("Unscaled MSE normalized by Seq_Norm (%s)" % self.config.model_dir.split('/')[-1])
        print(df[[x for x in df.columns if 'unscaled' in x]].mean())
        for i in range(self.config.forecast_steps):
            print("Scaled MSE for step %i: %1.4f" % (i + 1, df['norm_squared_diff_' + str(i + 1)].mean()))
        return df

    def _extract_targets(self, targets):
        return [np.expand_dims(x[:, -1, self.target_index], -1) for x in targets]

    def _extract_preds(self, preds):
        return [np.expand_dims(x[:, -1, self.target_index], -1) for x in preds]

    def _extract_inputs(self, inp, i):
        return np.expand_dims(inp[:, i, self.target_index], -1)
