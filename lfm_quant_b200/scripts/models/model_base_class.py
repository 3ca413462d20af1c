"""BaseModelClass (reference: scripts/models/model_base_class.py:11-51): shape bookkeeping shared by model classes.
The Cropping1D helpers there only serve forecast_steps > 1, which is a 'next' row of the scope table."""
from __future__ import absolute_import, division, print_function


class BaseModelClass(object):

    def __init__(self, seq_len, n_inputs, n_outputs):
        self.seq_len = seq_len
        self.n_inputs = n_inputs
        self.n_outputs = n_outputs

    def get_last_time_step_aux(self, sequence):
        """inp[:, -1:, n_outputs:] (model_base_class.py:18-38)."""
        return sequence[:, self.seq_len - 1:, self.n_outputs:]

    def get_last_time_step(self, sequence, count=None):
        """sequence[:, -1:, :] (model_base_class.py:40-51)."""
        return sequence[:, self.seq_len - 1:, :]
