"""Train driver (reference: scripts/train.py:27-432).

Same life cycle -- build model/optimizer/losses, materialise every batch up front (train.py:64-80), epoch loop
with shuffled batch order, CSV logs (train.py:133-147,156-167,227-238), save-best / early-stop
(train.py:240-266) -- but the step body ``_train_step_point`` (train.py:178-199) is ONE native call
(fwd + loss + BPTT + clip + optimizer + MaxNorm on the GPU) and the per-step ``mse.numpy()`` host sync of the
reference (train.py:199) is deferred: device scalars are read back only when a log line is due.
"""
from __future__ import absolute_import, division, print_function

import copy
import os
import pathlib
import random
import time
from collections import defaultdict

import numpy as np
import pandas as pd

from .model_utils.losses import Losses
from .model_utils.model import Model
from .model_utils.optimizers import Optimizers


class Train(object):

    def __init__(self, config, dataset):
        self.config = config
        self.dataset = dataset
        self.dataset.generate_dataset()
        # Data parallelism over the company-batch axis: under torchrun every rank walks the same seeded
        # batch sequence, gathers only its rows of each global batch and all-reduces the flat gradient once per step.
        # The process group / device is chosen before the model (and its workspace) is created on that device.
        from .. import dp as _dp
        self.rank, self.world = _dp.init_from_env()
        self.model = Model(self.config, self.dataset).get_model()
        if self.world > 1 and self.config.forecast_steps > 1:
            raise NotImplementedError('data parallelism is built for forecast_steps = 1')
        if self.world > 1:                       # one set of initial weights: rank 0's
            import torch
            import torch.distributed as dist
            flat = torch.from_numpy(self.model.engine.get_flat()).to(self.model.engine.device)
            dist.broadcast(flat, 0)
            self.model.engine.set_flat(flat.cpu().numpy())
        self.target_index = self.dataset.target_index
        self.optimizer = Optimizers(self.config).get_optimizer()
        self.optimizer.bind(self.model)
        self.losses = Losses(self.config, self.target_index, engine=self.model.engine)

        self.train_set = self.dataset.train_set
        self.valid_set = self.dataset.valid_set
        self.train_set = self.train_set.shuffle(buffer_size=10000, seed=self.config.seed)
        self.train_set = self.train_set.batch(batch_size=self.config.batch_size)
        self.valid_set = self.valid_set.batch(batch_size=self.config.batch_size)

        self.model_dir = os.path.join(self.config.experiments_dir, self.config.model_dir)
        self.chkpts_dir = os.path.join(self.model_dir, 'chkpts')
        self.train_log_dir = os.path.join(self.model_dir, 'train_log')
        train_dirs = [self.model_dir, self.chkpts_dir, self.train_log_dir]
        print(train_dirs)
        for dirname in train_dirs:
            if not os.path.isdir(dirname):
                print("Creating dir %s" % dirname)
                pathlib.Path(dirname).mkdir(parents=True, exist_ok=True)

        self.min_valid_mse = np.inf
        self.min_valid_uq_loss = np.inf
        self._grad_norm = 1.0

        print("Creating batches ...")
        self._batches = [self._build_train_batch(items) for items in self.train_set]
        self._valid_batches = [self.dataset.get_batch(*items) for items in self.valid_set]

    def _build_train_batch(self, items):
        """(inp, target, metadata[, row0, global denominators]) of one global batch; the shard of this rank under DP."""
        if self.world == 1:
            return self.dataset.get_batch(*items)
        import torch.distributed as dist
        from .. import dp as _dp
        inp_idx, tar_idx, meta, row0 = _dp.shard_batch_indices(items[0], items[1], items[2], self.rank, self.world)
        eng = self.model.engine
        if len(inp_idx) == 0:                    # fewer windows than ranks: this rank only joins the collectives
            import torch
            denom = _dp.global_denominators(torch.zeros(2, dtype=torch.float32, device=eng.device), dist)
            return (None, None, None, row0, denom)
        x, y, md = self.dataset.get_batch(inp_idx, tar_idx, meta)
        denom = _dp.global_denominators(eng.mask_count(y), dist)     # loss denominators are global (losses.py:87,90,131-135)
        return (x, y, md, row0, denom)

    def train(self):
        if self.config.load_saved_weights:
            self.model.load_weights(os.path.join(self.chkpts_dir, "chkpt"))
        print("Training in progress ...")
        epochs = self.config.max_epoch
        checkpoint_prefix = os.path.join(self.chkpts_dir, "chkpt")
        train_logs_batch = defaultdict(list)
        train_logs_epoch = defaultdict(list)
        if self.rank == 0:
            self.model.save_weights(checkpoint_prefix)
        valid_mse = None
        start = time.time()
        for epoch in range(epochs):
            self.model.reset_states()
            mse_steps, uq_loss_steps = [], []
            if self.world == 1 and not os.environ.get('LFMQ_SEEDED_SHUFFLE'):
                random.shuffle(self._batches)                   # train.py:115 (unseeded in the reference)
            else:                                               # every rank must walk the same batch order
                random.Random(self.config.seed + epoch).shuffle(self._batches)
            for (batch_n, cur_batch) in enumerate(self._batches):
                inp, target = cur_batch[0], cur_batch[1]
                if self.world > 1:
                    mse = self._train_step_point_dp(cur_batch)
                    uq_loss = None
                elif self.config.UQ:
                    uq_loss, mse = self._train_step_uq_range(inp, target)
                else:
                    mse = self._train_step_point(inp, target)   # device scalar, no host sync here
                    uq_loss = None
                mse_steps.append(mse)
                uq_loss_steps.append(uq_loss)
                if batch_n % self.config.logging_interval == 0:
                    train_logs_batch['batch_n'].append(batch_n)
                    train_logs_batch['time'].append(time.time() - start)
                    train_logs_batch['mse'].append(self._mean(mse_steps))
                    train_logs_batch['uq_loss'].append(self._mean(uq_loss_steps) if self.config.UQ else None)
                    train_logs_batch['valid_mse'].append(None)
                    train_logs_batch['valid_uq_loss'].append(None)
                    self._write_train_logs(train_logs_batch, 'train-logs-batch')

            if epoch % self.config.epoch_logging_interval == 0:
                if self.config.UQ:
                    valid_uq_loss, valid_mse, valid_mse_fcst = self._validation_metrics_uq_range_estimate()
                else:
                    valid_uq_loss, valid_mse, valid_mse_fcst = self._validation_metrics_point_estimate()
                train_logs_epoch['epoch'].append(epoch)
                train_logs_epoch['time'].append(time.time() - start)
                train_logs_epoch['mse'].append(self._mean(mse_steps))
                train_logs_epoch['uq_loss'].append(self._mean(uq_loss_steps) if self.config.UQ else None)
                train_logs_epoch['valid_mse'].append(valid_mse)
                train_logs_epoch['valid_uq_loss'].append(valid_uq_loss)
                train_logs_epoch['valid_mse_fcst'].append(valid_mse_fcst)
                self._write_train_logs(train_logs_epoch, 'train-logs-epoch')
                if self._save_criteria(train_logs_epoch) and self.rank == 0:
                    self.model.save_weights(checkpoint_prefix)
                if self._stop_criteria(train_logs_epoch):
                    break
        return valid_mse

    @staticmethod
    def _mean(device_scalars):
        import torch
        return float(torch.stack(list(device_scalars)).mean().item())

    def _train_step_point(self, inp, targets):
        """train.py:178-199 as one fused native step; returns the step's mse_0 as a device scalar."""
        assert not self.config.UQ
        lr = self.optimizer.current_lr()
        out = self.model.train_step(inp, targets, lr, self.optimizer.iterations)
        self.optimizer.iterations += 1
        return out[1]

    def _train_step_point_dp(self, cur_batch):
        """train.py:178-199 under data parallelism: local BPTT with global denominators, ONE all-reduce of the flat
        gradient (+ loss / mse tail), replicated clip + optimizer + MaxNorm.  Same trajectory as one GPU."""
        assert not self.config.UQ, 'data parallelism is built for the point-estimate model'
        import torch.distributed as dist
        from ..dp import allreduce_flat_gradient
        eng = self.model.engine
        inp, targets, _, row0, denom = cur_batch
        lr = self.optimizer.current_lr()
        it = self.optimizer.iterations
        if inp is None:
            eng.grads[:eng.n_trainable + 2].zero_()
            allreduce_flat_gradient(eng.grads, eng.n_trainable, dist)
            eng.apply(lr, it)
            out = eng.grads[eng.n_trainable:eng.n_trainable + 2].clone()
        else:
            out = eng.train_step_dp(inp, targets, it, lr, row0, denom)
        self.optimizer.iterations += 1
        return out[1]

    def _train_step_uq_range(self, inp, targets):
        """train.py:201-225 as one fused native step; returns device scalars (uq_loss_last_tar, mse_0)."""
        assert self.config.UQ
        lr = self.optimizer.current_lr()
        out = self.model.train_step(inp, targets, lr, self.optimizer.iterations)
        self.optimizer.iterations += 1
        return out[0], out[1]

    def _write_train_logs(self, train_logs, name):
        if self.rank != 0:
            return
        df = pd.DataFrame.from_dict(train_logs)
        fname = os.path.join(self.train_log_dir, self.config.name + '-' + name + '.csv')
        df.to_csv(fname, sep=',', index=False)

    def _save_criteria(self, train_logs_epoch):
        assert len(train_logs_epoch['valid_mse']) > 0, 'Error in computing valid_mse or incorrect train log dict passed'
        if train_logs_epoch['valid_mse'][-1] < self.min_valid_mse:
            self.min_valid_mse = train_logs_epoch['valid_mse'][-1]
            return True
        return False

    def _stop_criteria(self, train_logs_epoch):
        """train.py:253-266, including its off-by-one."""
        assert len(train_logs_epoch['valid_mse']) > 0, \
            'Error in computing valid_mse or incorrect train log dict passed'
        valid_mse = np.array(train_logs_epoch['valid_mse'])
        return bool(valid_mse.shape[0] - np.argmin(valid_mse) + 1 >= self.config.early_stop)

    def _validation_metrics_point_estimate(self):
        """train.py:268-336: predict every validation batch, loss over the stacked arrays, scaled and un-scaled -- all on
        the device: the validation batches are resident (built once in __init__), predictions stay in HBM
        (predict_device), un-scaling is lfmq_unscale, both losses are lfmq_loss; the only host traffic per epoch is the
        two result scalars."""
        import torch
        if not self._valid_batches:
            return None, float('nan'), float('nan')
        preds = [self.model.predict_device(cur_batch[0]) for cur_batch in self._valid_batches]
        targets = [cur_batch[1] for cur_batch in self._valid_batches]
        if self.config.forecast_steps == 1:              # preds, targets need to be lists (train.py:291-294)
            preds = [[p] for p in preds]
            targets = [[t] for t in targets]
        assert all(isinstance(t, (list, tuple)) for t in targets), \
            'forecast_steps > 1 needs a list of forecast_steps targets per batch (train.py:296-299)'
        scale, center = self._scaling_dev()
        eng = self.model.engine
        S = self.config.forecast_steps
        pred_all = [torch.cat([p[i] for p in preds], dim=0) for i in range(S)]            # train.py:316-320
        target_all = [torch.cat([t[i] for t in targets], dim=0) for i in range(S)]
        pred_unscaled = [eng.unscale(p, scale, center, self.config.log_squasher) for p in pred_all]
        target_unscaled = [eng.unscale(t, scale, center, self.config.log_squasher) for t in target_all]
        _, valid_mse = self.losses.weight_adjusted_mse(target_all, pred_all, True)
        _, valid_mse_fcst = self.losses.weight_adjusted_mse(target_unscaled, pred_unscaled, True)
        both = torch.stack([valid_mse._t, valid_mse_fcst._t]).cpu().numpy()      # the epoch's single D2H
        return None, float(both[0]), float(both[1])

    def _scaling_dev(self):
        import torch
        if getattr(self, '_scale_dev', None) is None:
            n = self.dataset.n_outputs
            dev = self.model.engine.device
            self._scale_dev = torch.from_numpy(np.ascontiguousarray(self.dataset.scaling_params['scale'][:n],
                                                                    dtype=np.float64)).to(dev)
            self._center_dev = torch.from_numpy(np.ascontiguousarray(self.dataset.scaling_params['center'][:n],
                                                                     dtype=np.float64)).to(dev)
        return self._scale_dev, self._center_dev

    def _validation_metrics_uq_range_estimate(self):
        """train.py:338-416: predict every validation batch (dropout stays on), uq loss over the stacked arrays, MSE of
        the un-scaled target predictions."""
        preds, var, targets = [], [], []
        for cur_batch in self._valid_batches:
            batch_pred = self.model.predict(cur_batch[0])
            preds.append(batch_pred[0::2][0])
            var.append(batch_pred[1::2][0])
            targets.append(cur_batch[1].cpu().numpy())
        if not preds:
            return float('nan'), float('nan'), float('nan')
        pred_all, var_all, target_all = (np.vstack(a).astype('float32') for a in (preds, var, targets))
        pred_unscaled = self._unscale_preds(copy.deepcopy(pred_all))
        target_unscaled = self._unscale_preds(copy.deepcopy(target_all))
        _, valid_uq_loss, valid_mse = self.losses.weight_adjusted_uq_loss([target_all], [pred_all], [var_all])
        _, valid_mse_fcst = self.losses.weight_adjusted_mse([target_unscaled], [pred_unscaled])
        return valid_uq_loss.numpy(), valid_mse.numpy(), valid_mse_fcst.numpy()

    def _unscale_preds(self, arr):
        """train.py:420-432."""
        arr = np.multiply(arr, self.dataset.scaling_params['scale'][:self.dataset.n_outputs]) + \
            self.dataset.scaling_params['center'][:self.dataset.n_outputs]
        if self.config.log_squasher:
            arr = self.dataset.reverse_log_squasher(arr)
        return arr
