"""Generates tests/golden/reference_flags.json by running the UNMODIFIED reference flag parser
(/root/reference/scripts/configs.py + base_config.py -- pure Python, no TensorFlow) in a subprocess per argv case.

Only runs where /root/reference exists (the build container); the JSON it writes is what travels.  The numeric code of
the reference cannot be run this way (it imports TensorFlow at module level), so this fixture pins the flag / config
boundary only (the --config CLI and its flag schema).

usage: python tests/golden/make_reference_flags.py
"""
import json
import os
import subprocess
import sys

REF = '/root/reference/scripts'
CASES = {
    'defaults': [],
    'overrides': ['--nn_type', 'RNNUqRangeEstimate', '--UQ=True', '--rnn_cell', 'gru', '--num_hidden', '128',
                  '--learning_rate', '0.01', '--train=False', '--forecast_steps_weights', '1.0,0.5'],
    'bool_forms': ['--UQ', '--notrain', '--log_squasher=false'],
    'years': ['--min_years', '3', '--pls_years', '2', '--stride', '3'],
}
CHILD = r'''
import json, sys
sys.path.insert(0, %r)
sys.argv = ['lfm_quant.py'] + json.loads(%r)
import base_config
c = base_config.get_configs()
print(json.dumps(c.__dict__['__configs'], sort_keys=True, default=str))
'''


def main():
    out = {}
    for name, argv in CASES.items():
        r = subprocess.run([sys.executable, '-c', CHILD % (REF, json.dumps(argv))], capture_output=True, text=True)
        if r.returncode != 0:
            out[name] = {'argv': argv, 'error': r.stderr.strip().splitlines()[-1] if r.stderr.strip() else 'failed'}
            continue
        out[name] = {'argv': argv, 'values': json.loads(r.stdout.strip().splitlines()[-1])}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_flags.json')
    json.dump(out, open(dst, 'w'), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, 'error: ' + v['error'] if 'error' in v else '%d flags' % len(v['values']))


if __name__ == '__main__':
    main()
