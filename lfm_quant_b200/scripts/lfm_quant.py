#!/usr/bin/env python3
"""CLI entry (reference: scripts/lfm_quant.py:1-150):  lfm_quant.py --config=FILE [--train=True|False] [flags]

Every flag of the reference is accepted (base_config.SCHEMA); ``--precision bf16`` selects the tensor-core path.
LFM_QUANT_ROOT must be set, as in the reference (lfm_quant.py:15-16).
"""
from __future__ import absolute_import, division, print_function

import os
import sys

if __package__ in (None, ''):       # executed as a script: make the package importable
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from lfm_quant_b200.scripts import base_config
    from lfm_quant_b200.scripts.runtime.model_execution import ModelExecution
else:
    from . import base_config
    from .runtime.model_execution import ModelExecution


def get_configs(argv=None):
    return base_config.get_configs(argv, list_sep='-')


def main(argv=None):
    assert os.environ.get('LFM_QUANT_ROOT') is not None, "Environment Variable LFM_QUANT_ROOT not set"
    config = get_configs(argv)
    if config.cdrs_inference:
        raise NotImplementedError('CDRS inference needs the proprietary cdrs package (runtime/cdrs_data.py:6)')
    return ModelExecution(config)()


if __name__ == "__main__":
    main()
