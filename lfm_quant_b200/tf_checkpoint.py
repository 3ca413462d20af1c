"""TensorFlow "tensor bundle" checkpoints without TensorFlow.

The reference saves and restores weights with ``model.save_weights(<model_dir>/chkpts/chkpt)`` /
``model.load_weights(...)`` (scripts/train.py:87,99,171; scripts/predict.py:93): Keras' TF-format checkpoint, i.e. the
files ``chkpt.index`` + ``chkpt.data-00000-of-00001`` (+ the ``checkpoint`` state file).  This module reads and writes
that container in pure Python/NumPy so a model trained by either side can be served by the other:

* ``chkpt.index`` is a LevelDB-format sorted string table (blocks of prefix-compressed key/value entries with restart
  points, 5-byte block trailers = compression type + masked CRC-32C, metaindex + index blocks, 48-byte footer with the
  magic 0xdb4775248b80fb57).  Key "" holds a BundleHeaderProto, every other key a BundleEntryProto (dtype, shape,
  shard_id, offset, size, masked crc32c) pointing into the data shard.  TensorFlow writes it uncompressed
  (tensor_bundle.cc sets ``options.compression = kNoCompression``); a snappy block raises.
* ``chkpt.data-00000-of-00001`` is the raw little-endian tensor bytes, back to back.
* Keras object-based checkpoints name variables by their path in the object graph,
  ``layer_with_weights-<k>/<attr>[/<attr>]/.ATTRIBUTES/VARIABLE_VALUE``, and store that graph as a serialized
  ``TrackableObjectGraph`` proto under ``_CHECKPOINTABLE_OBJECT_GRAPH`` (a DT_STRING tensor).  ``keras_keys`` maps the
  reference's layer stack (rnn_point_estimate.py:76-107: per layer LSTM|GRU -> BatchNormalization, then Dense) to these
  keys; ``write_keras_checkpoint`` also emits the object graph so that ``load_weights`` can match by structure.

STATUS: written from the published formats; the round trip (write -> read) and every checksum are tested here
(tests/test_tf_checkpoint.py), but TensorFlow is not installable in this image, so neither direction has been exercised
against a real TensorFlow yet -- the same "unpinned" status as oracle/pin_with_tf.py.
"""
from __future__ import absolute_import, division, print_function

import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DT = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 9: np.dtype('<i8'), 10: np.dtype('bool')}
_DT_OF = {np.dtype('float32'): 1, np.dtype('float64'): 2, np.dtype('int32'): 3, np.dtype('int64'): 9,
          np.dtype('bool'): 10}
DT_STRING = 7
OBJECT_GRAPH_KEY = '_CHECKPOINTABLE_OBJECT_GRAPH'
_SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'

# ---- CRC-32C (Castagnoli), table driven; TensorFlow / LevelDB store it masked -------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = tab
    return _CRC_TABLE


def crc32c(data, crc=0):
    tab = _crc_table()
    c = (~crc) & 0xFFFFFFFF
    for b in bytes(data):
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return (~c) & 0xFFFFFFFF


def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def crc32c_combine(crc1, crc2, len2):
    """CRC-32C of A+B from crc(A), crc(B), len(B) (zlib's crc32_combine with the Castagnoli polynomial)."""
    if len2 <= 0:
        return crc1
    odd = [0x82F63B78] + [1 << n for n in range(31)]          # operator for one zero bit
    even = _gf2_square(odd)                                    # two zero bits
    odd = _gf2_square(even)                                    # four
    while True:
        even = _gf2_square(odd)
        if len2 & 1:
            crc1 = _gf2_times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = _gf2_square(even)
        if len2 & 1:
            crc1 = _gf2_times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


def crc32c_array(a):
    """CRC-32C of a large buffer: the bytes are cut into equal lanes whose table walks run side by side in NumPy (one
    vector step per byte position), and the lane CRCs are folded together with crc32c_combine."""
    buf = np.frombuffer(memoryview(np.ascontiguousarray(a)).cast('B'), dtype=np.uint8)
    n = buf.size
    if n < (1 << 16):
        return crc32c(buf.tobytes())
    lanes = 2048
    ln = n // lanes
    body = buf[:lanes * ln].reshape(lanes, ln)
    tab = _crc_table()
    c = np.full(lanes, 0xFFFFFFFF, dtype=np.uint32)
    for i in range(ln):
        c = tab[(c ^ body[:, i]) & 0xFF] ^ (c >> np.uint32(8))
    c = ~c
    total = int(c[0])
    for k in range(1, lanes):
        total = crc32c_combine(total, int(c[k]), ln)
    tail = buf[lanes * ln:].tobytes()
    if tail:
        total = crc32c_combine(total, crc32c(tail), len(tail))
    return total


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - 0xa282ead8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf -------------------------------------------------------------------------------------
def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _pb_fields(buf):
    """Yields (field_number, wire_type, value) of one serialized message (varint, 64-bit, length-delimited, 32-bit)."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield f, wt, v


def _pb_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v)


def _pb_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + bytes(b)


def _pb_fixed32(field, v):
    return _put_varint((field << 3) | 5) + struct.pack('<I', v)


# ---- LevelDB-format table ---------------------------------------------------------------------------------------------
def _read_block(f, offset, size):
    raw = f[offset:offset + size + 5]
    body, ctype, stored = raw[:size], raw[size], struct.unpack_from('<I', raw, size + 1)[0]
    if unmask_crc(stored) != crc32c(raw[:size + 1]):
        raise ValueError('table block at %d: checksum mismatch' % offset)
    if ctype != 0:
        raise ValueError('table block at %d is compressed (type %d); TensorFlow writes bundle indexes uncompressed' %
                         (offset, ctype))
    n_restarts = struct.unpack_from('<I', body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * n_restarts
    entries, pos, key = [], 0, b''
    while pos < end:
        shared, pos = _get_varint(body, pos)
        non_shared, pos = _get_varint(body, pos)
        vlen, pos = _get_varint(body, pos)
        key = key[:shared] + bytes(body[pos:pos + non_shared])
        pos += non_shared
        entries.append((key, bytes(body[pos:pos + vlen])))
        pos += vlen
    return entries


def _read_table(path):
    f = open(path, 'rb').read()
    if len(f) < 48 or struct.unpack_from('<Q', f, len(f) - 8)[0] != _MAGIC:
        raise ValueError('%s is not a TensorFlow checkpoint index (bad table magic)' % path)
    foot = f[len(f) - 48:]
    pos = 0
    _, pos = _get_varint(foot, pos)       # metaindex offset
    _, pos = _get_varint(foot, pos)       # metaindex size
    ioff, pos = _get_varint(foot, pos)
    isz, pos = _get_varint(foot, pos)
    out = []
    for _, handle in _read_block(f, ioff, isz):
        boff, p = _get_varint(handle, 0)
        bsz, p = _get_varint(handle, p)
        out += _read_block(f, boff, bsz)
    return out


def _build_block(entries, restart_interval=16):
    body, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(body))
            shared = 0
        else:
            shared = 0
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack('<I', r)
    body += struct.pack('<I', len(restarts))
    return bytes(body)


def _write_table(path, items, block_size=4096):
    """items: sorted list of (key bytes, value bytes)."""
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)                                           # kNoCompression
        out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
        return _put_varint(off) + _put_varint(len(block))

    index, cur, cur_bytes = [], [], 0
    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 8
        if cur_bytes >= block_size:
            index.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_bytes = [], 0
    if cur:
        index.append((cur[-1][0], emit(_build_block(cur))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    foot = meta + idx
    out.extend(foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', _MAGIC))
    with open(path, 'wb') as fh:
        fh.write(bytes(out))


# ---- tensor bundle ----------------------------------------------------------------------------------------------------
def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None)
    for f, wt, v in _pb_fields(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:                                     # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3
                    e['shape'].append(size)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = v
    return e


def read_bundle(prefix, verify=True):
    """{key: ndarray} of every numeric tensor of the checkpoint ``prefix`` (string tensors such as the object graph are
    returned as raw bytes under their key)."""
    entries = _read_table(prefix + '.index')
    if not entries or entries[0][0] != b'':
        raise ValueError('%s.index has no bundle header' % prefix)
    num_shards = 1
    for f, _, v in _pb_fields(entries[0][1]):
        if f == 1:
            num_shards = v
        if f == 2 and v != 0:
            raise ValueError('big-endian checkpoints are not supported')
    shards = {}
    out = {}
    for key, val in entries[1:]:
        e = _parse_entry(val)
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = open('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), 'rb').read()
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if len(raw) != e['size']:
            raise ValueError('checkpoint %s: tensor %r is truncated' % (prefix, key))
        name = key.decode('utf-8')
        if e['dtype'] == DT_STRING:
            out[name] = raw
            continue
        if e['dtype'] not in _DT:
            raise ValueError('checkpoint %s: tensor %r has unsupported dtype %d' % (prefix, key, e['dtype']))
        if verify and e['crc32c'] is not None and unmask_crc(e['crc32c']) != crc32c_array(np.frombuffer(raw, dtype=np.uint8)):
            raise ValueError('checkpoint %s: tensor %r fails its checksum' % (prefix, key))
        out[name] = np.frombuffer(raw, dtype=_DT[e['dtype']]).reshape(e['shape']).copy()
    return out


def _entry_proto(dtype, shape, offset, size, crc):
    shp = b''.join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)
    msg = _pb_varint(1, dtype)
    msg += _pb_bytes(2, shp)
    if offset:
        msg += _pb_varint(4, offset)
    msg += _pb_varint(5, size)
    msg += _pb_fixed32(6, crc)
    return msg


def _string_tensor_bytes(s):
    """Scalar DT_STRING tensor as tensor_bundle.cc:WriteStringTensor lays it out: [varint64 length][4-byte masked
    CRC-32C of the lengths, each taken as a uint32][the bytes]; the entry checksum runs over the length (as uint32), the
    length checksum and the bytes."""
    ln = struct.pack('<I', len(s))
    c = crc32c(ln)
    length_ck = struct.pack('<I', mask_crc(c))
    c = crc32c(length_ck, c)
    c = crc32c(s, c)
    return _put_varint(len(s)) + length_ck + bytes(s), mask_crc(c)


def write_bundle(prefix, tensors, string_tensors=None):
    """Writes ``prefix.index`` + ``prefix.data-00000-of-00001`` (+ the ``checkpoint`` state file beside them).
    tensors: {key: ndarray}; string_tensors: {key: bytes} scalar DT_STRING entries (the object graph)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data = bytearray()
    items = []
    allk = sorted(list(tensors) + list(string_tensors or {}))
    for key in allk:
        off = len(data)
        if string_tensors and key in string_tensors:
            raw, crc = _string_tensor_bytes(string_tensors[key])
            data += raw
            items.append((key.encode('utf-8'), _entry_proto(DT_STRING, [], off, len(raw), crc)))
            continue
        a = np.asarray(tensors[key])                 # (np.ascontiguousarray would turn a scalar into shape (1,))
        if a.dtype not in _DT_OF:
            raise ValueError('write_bundle: dtype %s of %r is not supported' % (a.dtype, key))
        raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes(order='C')
        data += raw
        items.append((key.encode('utf-8'), _entry_proto(_DT_OF[a.dtype], a.shape, off, len(raw),
                                                       mask_crc(crc32c_array(np.frombuffer(raw, dtype=np.uint8))))))
    header = _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))     # num_shards = 1, little endian, version.producer = 1
    with open(prefix + '.data-00000-of-00001', 'wb') as fh:
        fh.write(bytes(data))
    _write_table(prefix + '.index', [(b'', header)] + items)
    base = os.path.basename(prefix)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), 'checkpoint'), 'w') as fh:
        fh.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


# ---- Keras naming for the reference's layer stack -------------------------------------------------------------------------
def keras_keys(spec_names, construction_order=False):
    """Maps this package's variable names (lfmq_param_spec: ``lstm_1/kernel``, ``batch_normalization/gamma``,
    ``OUTPUT_1/kernel`` ...) to the keys Keras' object-based save_weights uses.  Layers with weights are numbered in
    construction order (rnn_point_estimate.py:76-107): recurrent layer, its BatchNormalization, the next recurrent layer
    ..., then the output Dense layer(s).  Recurrent variables live on the layer's ``cell``."""
    order = []
    for n in spec_names:
        layer = n.split('/')[0]
        if layer not in order:
            order.append(layer)

    def rank(layer):        # lstm_1, batch_normalization, lstm_2, batch_normalization_1, ..., OUTPUT_*
        if layer.startswith(('lstm_', 'gru_')):
            return (0, 2 * (int(layer.split('_')[-1]) - 1))
        if layer.startswith('batch_normalization'):
            k = 0 if layer == 'batch_normalization' else int(layer.split('_')[-1])
            return (0, 2 * k + 1)
        return (1, order.index(layer))
    # forecast_steps > 1 (rnn_point_estimate.py:109-150) interleaves heads and extra recurrent layers: the names then
    # arrive in construction order already (trunk, OUTPUT_1, lstm_{L+1}, its BatchNormalization, OUTPUT_2, ...).
    # NOTE: that numbering is not pinned against a TensorFlow-written multi-step checkpoint (none exists upstream).
    layers = list(order) if construction_order else sorted(order, key=rank)
    out = {}
    for n in spec_names:
        layer, var = n.split('/')
        k = layers.index(layer)
        mid = 'cell/' if layer.startswith(('lstm_', 'gru_')) else ''
        out[n] = 'layer_with_weights-%d/%s%s%s' % (k, mid, var, _SUFFIX)
    return out


def _object_graph(keys_by_name, keras_var_names):
    """Serialized TrackableObjectGraph: root -> layer_with_weights-k -> [cell ->] variable; every variable node carries
    one SerializedTensor {name 'VARIABLE_VALUE', full_name, checkpoint_key}."""
    nodes = [dict(children=[], attrs=[])]          # node 0 = the Model

    def child(parent, local):
        for cid, name in nodes[parent]['children']:
            if name == local:
                return cid
        nodes.append(dict(children=[], attrs=[]))
        nodes[parent]['children'].append((len(nodes) - 1, local))
        return len(nodes) - 1
    for name in sorted(keys_by_name, key=lambda n: keys_by_name[n]):
        key = keys_by_name[name]
        path = key[:-len(_SUFFIX)].split('/')
        cur = 0
        for part in path:
            cur = child(cur, part)
        nodes[cur]['attrs'].append(('VARIABLE_VALUE', keras_var_names.get(name, name), key))
    msg = b''
    for nd in nodes:
        body = b''
        for cid, local in nd['children']:
            body += _pb_bytes(1, _pb_varint(1, cid) + _pb_bytes(2, local.encode()))
        for nm, full, ck in nd['attrs']:
            body += _pb_bytes(2, _pb_bytes(1, nm.encode()) + _pb_bytes(2, full.encode()) + _pb_bytes(3, ck.encode()))
        msg += _pb_bytes(1, body)
    return msg


def write_keras_checkpoint(prefix, arrays, construction_order=False):
    """arrays: {this package's variable name: ndarray} -> TF-format checkpoint the reference's load_weights addresses."""
    keys = keras_keys(list(arrays), construction_order)
    full = {}
    for n in arrays:                                # Keras variable names: lstm_1/lstm_cell/kernel:0 etc.
        layer, var = n.split('/')
        cell = {'lstm': 'lstm_cell', 'gru': 'gru_cell'}.get(layer.split('_')[0])
        full[n] = ('%s/%s/%s' % (layer, cell, var)) if cell and layer.startswith(('lstm_', 'gru_')) else n
    write_bundle(prefix, {keys[n]: np.asarray(a, dtype=np.float32) for n, a in arrays.items()},
                 {OBJECT_GRAPH_KEY: _object_graph(keys, full)})


def read_keras_checkpoint(prefix, spec_names, shapes=None, construction_order=False):
    """{this package's variable name: ndarray} from a TF-format checkpoint written by the reference (or by
    write_keras_checkpoint).  Raises KeyError naming the first variable the checkpoint does not hold."""
    bundle = read_bundle(prefix)
    keys = keras_keys(list(spec_names), construction_order)
    out = {}
    for n in spec_names:
        if keys[n] not in bundle:
            have = sorted(k for k in bundle if k.endswith(_SUFFIX))
            raise KeyError('checkpoint %s has no %s (for %s); it holds: %s' % (prefix, keys[n], n, ', '.join(have)))
        out[n] = bundle[keys[n]]
        if shapes is not None and tuple(out[n].shape) != tuple(shapes[n]):
            raise ValueError('checkpoint %s: %s has shape %s, the model wants %s' % (prefix, keys[n], out[n].shape, shapes[n]))
    return out
