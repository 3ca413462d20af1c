"""ModelExecution (reference: scripts/runtime/model_execution.py:19-157): ``fixed_dates`` training / prediction of
one config in-process.  The ensemble / multi-process fan-out of the reference (:113-146,159-199) spawns whole
independent trainings through files and is outside the data-parallel hot path; ``num_procs > 1`` is rejected."""
from __future__ import absolute_import, division, print_function

import os

import numpy as np
import pandas as pd


class ModelExecution(object):

    def __init__(self, config):
        self.config = config

    def __call__(self):
        if self.config.training_type == 'iterative':
            raise NotImplementedError
        return self.fixed_dates_execution()

    @staticmethod
    def read_results(config):
        """Validation metrics after training, predictions after inference (model_execution.py:31-78)."""
        try:
            if config.train:
                df = pd.read_csv(os.path.join(config.experiments_dir, config.model_dir, 'train_log',
                                              config.name + '-train-logs-epoch.csv'), sep=',')
                df = df.sort_values(by='valid_mse').reset_index()
                return df.iloc[0]['valid_mse'], None, None
            df = pd.read_csv(os.path.join(config.experiments_dir, config.model_dir, 'pred', config.preds_fname),
                             sep=' ', dtype={'gvkey': str})
            valid_loss = 0
            for i in range(config.forecast_steps):
                valid_loss += df['norm_squared_diff_' + str(i + 1)].mean() * config.forecast_steps_weights[i]
            return valid_loss, None, df
        except FileNotFoundError:
            print("Output file not found")
            return np.inf, np.inf, None

    @staticmethod
    def single_execution(config):
        """model_execution.py:81-111; the GPU is the last character of ``default_gpu`` (:88)."""
        import torch
        gpu = str(config.default_gpu)[-1]
        if gpu.isdigit() and torch.cuda.is_available() and int(gpu) < torch.cuda.device_count():
            torch.cuda.set_device(int(gpu))
        from ..data_processing import CDRSInferenceData, Dataset
        from ..predict import Predict
        from ..train import Train
        dataset = CDRSInferenceData(config) if config.cdrs_inference else Dataset(config)
        if config.train:
            print("Training")
            return Train(config, dataset).train()
        print("Prediction")
        return Predict(config, dataset).predict()

    def uq_estimate_execution(self):
        """runtime/model_execution.py:201-241: training, or prediction with one process, is a single execution; the
        multi-process MC-dropout ensemble (num_procs > 1 at predict time) fans out through files like the other
        ensembles and is not built."""
        if self.config.train or self.config.num_procs == 1:
            return self.single_execution(self.config)
        raise NotImplementedError('UQ ensemble prediction (num_procs > 1) fans out whole predictions through files; run '
                                  'one lfm_quant.py per member')

    def fixed_dates_execution(self):
        if self.config.UQ:
            return self.uq_estimate_execution()
        if self.config.num_procs != 1:
            raise NotImplementedError('ensembles (num_procs > 1) fan out whole trainings through files; run one '
                                      'lfm_quant.py per member')
        return self.single_execution(self.config)
