"""TF-format checkpoint container (lfm_quant_b200/tf_checkpoint.py) -- CPU tests: CRC-32C known answers, the
LevelDB-format index (prefix compression, restarts, several blocks), the tensor round trip, Keras key naming for the
reference's layer stack, corruption detection.  (Not exercised against a real TensorFlow: none is installable here.)"""
import os

import numpy as np
import pytest

from lfm_quant_b200 import tf_checkpoint as tfc

import lfm_oracle as orc


def test_crc32c_known_answers_and_combine():
    assert tfc.crc32c(b'123456789') == 0xE3069283               # the standard CRC-32C check value
    assert tfc.crc32c(b'\x00' * 32) == 0x8A9136AA               # RFC 3720 B.4
    assert tfc.crc32c(b'\xff' * 32) == 0x62A8AB43
    a, b = os.urandom(1000), os.urandom(777)
    assert tfc.crc32c_combine(tfc.crc32c(a), tfc.crc32c(b), len(b)) == tfc.crc32c(a + b)
    big = np.frombuffer(os.urandom(300001), dtype=np.uint8)
    assert tfc.crc32c_array(big) == tfc.crc32c(big.tobytes())
    for c in (0, 1, 0xdeadbeef, 0xffffffff):
        assert tfc.unmask_crc(tfc.mask_crc(c)) == c


def test_bundle_round_trip_many_keys_and_blocks(tmp_path):
    rng = np.random.RandomState(0)
    tensors = {'layer_with_weights-%d/v%03d/.ATTRIBUTES/VARIABLE_VALUE' % (i % 7, i): rng.normal(size=(i % 5 + 1, 3)).astype(np.float32)
               for i in range(300)}                              # > one 4 KB index block, long shared prefixes
    tensors['scalar'] = np.array(3.5, dtype=np.float32)
    tensors['ints'] = np.arange(6, dtype=np.int64).reshape(2, 3)
    prefix = str(tmp_path / 'chkpts' / 'chkpt')
    tfc.write_bundle(prefix, tensors, {tfc.OBJECT_GRAPH_KEY: b'graph-bytes'})
    assert os.path.isfile(prefix + '.index') and os.path.isfile(prefix + '.data-00000-of-00001')
    assert 'model_checkpoint_path: "chkpt"' in open(str(tmp_path / 'chkpts' / 'checkpoint')).read()
    back = tfc.read_bundle(prefix)
    assert set(back) == set(tensors) | {tfc.OBJECT_GRAPH_KEY}
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape
        np.testing.assert_array_equal(back[k], v)
    raw = back[tfc.OBJECT_GRAPH_KEY]
    assert raw.endswith(b'graph-bytes') and raw[0] == len(b'graph-bytes')


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'c')
    tfc.write_bundle(prefix, {'a': np.arange(8, dtype=np.float32)})
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[5] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    with pytest.raises(ValueError, match='checksum'):
        tfc.read_bundle(prefix)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[3] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError):
        tfc.read_bundle(prefix)


@pytest.mark.parametrize('L,cell,uq', [(1, 'lstm', False), (2, 'lstm', False), (2, 'gru', False), (1, 'lstm', True)])
def test_keras_keys_follow_the_reference_layer_stack(tmp_path, L, cell, uq):
    names = orc.param_names(L, cell, uq=uq)
    for l in range(L):
        bn = 'batch_normalization' if l == 0 else 'batch_normalization_%d' % l
        names += [bn + '/moving_mean', bn + '/moving_variance']
    keys = tfc.keras_keys(names)
    c = 'lstm' if cell == 'lstm' else 'gru'
    assert keys['%s_1/kernel' % c] == 'layer_with_weights-0/cell/kernel/.ATTRIBUTES/VARIABLE_VALUE'
    assert keys['%s_1/recurrent_kernel' % c] == 'layer_with_weights-0/cell/recurrent_kernel/.ATTRIBUTES/VARIABLE_VALUE'
    assert keys['batch_normalization/gamma'] == 'layer_with_weights-1/gamma/.ATTRIBUTES/VARIABLE_VALUE'
    assert keys['batch_normalization/moving_variance'] == 'layer_with_weights-1/moving_variance/.ATTRIBUTES/VARIABLE_VALUE'
    if L == 2:
        assert keys['%s_2/bias' % c] == 'layer_with_weights-2/cell/bias/.ATTRIBUTES/VARIABLE_VALUE'
        assert keys['batch_normalization_1/beta'] == 'layer_with_weights-3/beta/.ATTRIBUTES/VARIABLE_VALUE'
    head = 'OUTPUT_TARGET_1' if uq else 'OUTPUT_1'
    assert keys[head + '/kernel'] == 'layer_with_weights-%d/kernel/.ATTRIBUTES/VARIABLE_VALUE' % (2 * L)
    if uq:
        assert keys['OUTPUT_VARIANCE_1/bias'] == 'layer_with_weights-%d/bias/.ATTRIBUTES/VARIABLE_VALUE' % (2 * L + 1)
    # round trip through the Keras-named container
    rng = np.random.RandomState(1)
    arrays = {n: rng.normal(size=(3, 4)).astype(np.float32) for n in names}
    prefix = str(tmp_path / 'chkpt')
    tfc.write_keras_checkpoint(prefix, arrays)
    back = tfc.read_keras_checkpoint(prefix, names, {n: (3, 4) for n in names})
    for n in names:
        np.testing.assert_array_equal(back[n], arrays[n])
    graph = tfc.read_bundle(prefix)[tfc.OBJECT_GRAPH_KEY]
    assert b'layer_with_weights-0' in graph and b'VARIABLE_VALUE' in graph and b'cell' in graph
    with pytest.raises(KeyError, match='layer_with_weights'):
        tfc.read_keras_checkpoint(prefix, names + ['lstm_9/kernel'])


def test_construction_order_keys_for_the_multi_step_graph(tmp_path):
    """forecast_steps > 1 (rnn_point_estimate.py:109-150): heads and extra recurrent layers interleave, so the
    layer_with_weights numbering follows construction order; round trip through the TF-format container."""
    from lfm_quant_b200 import tf_checkpoint as tfc
    names = ['lstm_1/kernel', 'lstm_1/recurrent_kernel', 'lstm_1/bias', 'batch_normalization/gamma',
             'batch_normalization/beta', 'OUTPUT_1/kernel', 'OUTPUT_1/bias', 'batch_normalization/moving_mean',
             'batch_normalization/moving_variance', 'lstm_2/kernel', 'lstm_2/recurrent_kernel', 'lstm_2/bias',
             'batch_normalization_1/gamma', 'batch_normalization_1/beta', 'OUTPUT_2/kernel', 'OUTPUT_2/bias',
             'batch_normalization_1/moving_mean', 'batch_normalization_1/moving_variance']
    keys = tfc.keras_keys(names, construction_order=True)
    assert keys['lstm_1/kernel'].startswith('layer_with_weights-0/cell/kernel')
    assert keys['batch_normalization/moving_mean'].startswith('layer_with_weights-1/moving_mean')
    assert keys['OUTPUT_1/bias'].startswith('layer_with_weights-2/bias')
    assert keys['lstm_2/recurrent_kernel'].startswith('layer_with_weights-3/cell/recurrent_kernel')
    assert keys['batch_normalization_1/gamma'].startswith('layer_with_weights-4/gamma')
    assert keys['OUTPUT_2/kernel'].startswith('layer_with_weights-5/kernel')
    # the default (forecast_steps = 1) numbering puts every recurrent layer before the heads
    assert tfc.keras_keys(names)['OUTPUT_1/bias'].startswith('layer_with_weights-4/')
    rng = np.random.RandomState(0)
    arrs = {n: rng.normal(size=(3, 4) if n.endswith('kernel') else (4,)).astype(np.float32) for n in names}
    prefix = str(tmp_path / 'chkpt')
    tfc.write_keras_checkpoint(prefix, arrs, construction_order=True)
    back = tfc.read_keras_checkpoint(prefix, names, {n: a.shape for n, a in arrs.items()}, construction_order=True)
    for n in names:
        np.testing.assert_array_equal(back[n], arrs[n])
