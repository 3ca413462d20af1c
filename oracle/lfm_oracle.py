"""CPU oracle for the lfm_quant recurrent-forecaster hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
path (``lfm_quant_b200``) fails loudly when its CUDA extension is missing.

PINNING STATUS, by row of SURVEY section 8a.
  * Batcher rows (a1 window index, a2 get_batch, scaler): PINNED to the reference itself.  Its data_processing.py uses
    TensorFlow only to wrap arrays, so the unmodified file runs in the build container behind an import shim;
    tests/golden/make_reference_batcher.py recorded what Dataset.generate_dataset / get_batch return on a synthetic
    table (tests/golden/reference_batcher_*.npz) and tests/test_golden_batcher.py holds gather_batch, the window index,
    the split and the scaler procedure to it.  The flag parser is pinned the same way (reference_flags.json), and so is
    the prediction driver: reference_preds.dat is what the unmodified predict.py writes with a stub model, and
    tests/test_gpu_cli.py holds the package's Predict + CUDA batcher to it (oracle not involved).
  * Loss rows (a7 weight_adjusted_mse, and weight_adjusted_uq_loss): LOGIC PINNED to the reference's own code.  losses.py
    is tensor algebra over 14 TensorFlow primitives; with each mapped to its NumPy equivalent the unmodified file runs
    (tests/golden/make_reference_losses.py -> reference_losses.npz) and tests/test_golden_losses.py holds
    loss_point_estimate / loss_uq_estimate to its answers, incl. the NaN of the UQ loss on a zero-padded step.  What is
    pinned is the masking / slicing / weighting / denominators; the primitives' arithmetic is NumPy float32, not TF's.
  * Model / backward / optimizer rows (a3-a6, a8-a11): PARITY UNPINNED.  The reference (lakshaykc/lfm_quant @ ac6f47c) keeps that
    arithmetic inside TensorFlow 2.x / Keras, which is neither vendored under /root/reference nor pinned (no
    requirements.txt / lock file; API usage dates it to TF 2.0-2.3) and cannot be installed here, and it ships no tests or
    golden vectors for it.  Those functions are a NumPy restatement of the reference's own call sites plus the published
    Keras layer algorithms, checked against
      (1) the only fixture in the reference, the loss example in scripts/model_utils/losses.py:287-310 (expected values
          derived by hand in tests/test_oracle.py, confirmed by the reference's own code in reference_losses.npz),
      (2) torch.nn.LSTM / torch.nn.GRU / torch autograd / torch.optim on CPU as independent second opinions,
      (3) fp64 finite differences of the full step,
    and oracle/pin_with_tf.py is the (never yet executed) route to pin them where TensorFlow exists.

Every function cites the reference file:line (relative to /root/reference/scripts) it follows.
"""
from __future__ import annotations

import numpy as np

_MIN_SEQ_NORM = 10  # data_processing.py:20
BN_EPS = 1e-3       # keras.layers.BatchNormalization default epsilon (rnn_point_estimate.py:88)


# --------------------------------------------------------------------------------------
# Counter-based RNG shared by oracle and kernels (dropout masks must be reproducible on any
# GPU count).  Philox4x32-10, Salmon et al. 2011.
# --------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask32 = np.uint64(0xFFFFFFFF)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(_PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_threshold(rate):
    """Integer threshold on the top 24 random bits: keep iff (r >> 8) >= thr."""
    return int(np.float64(rate) * 16777216.0)


def dropout_mask(seed, step, stream, row0, n_rows, inner, rate, dtype=np.float64):
    """Inverted-dropout scale mask of shape [n_rows, inner], value in {0, 1/(1-rate)}.

    Element (r, j) is keyed by the *global* row index ``row0 + r`` so the mask does not depend
    on how the batch is sharded over GPUs.  ``inner`` must be a multiple of 4; one Philox call
    yields 4 consecutive elements.  Counter = (q_lo, q_hi, stream, step), q = (row*inner + j)//4.
      stream = 2*layer   : Dropout after BN   (rnn_point_estimate.py:89), inner = T*H
      stream = 2*layer+1 : recurrent dropout  (rnn_point_estimate.py:86), inner = H
    """
    assert inner % 4 == 0
    if rate <= 0.0:
        return np.ones((n_rows, inner), dtype=dtype)
    rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(row0))[:, None]
    q = rows * np.uint64(inner // 4) + np.arange(inner // 4, dtype=np.uint64)[None, :]
    c0 = (q & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (q >> np.uint64(32)).astype(np.uint32)
    r = philox4x32_10(c0, c1, np.uint32(stream), np.uint32(step & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = np.stack(r, axis=-1).reshape(n_rows, inner)
    keep = (bits >> np.uint32(8)) >= np.uint32(dropout_threshold(rate))
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(rate))   # fp32, as the kernels do
    return np.where(keep, dtype(scale), dtype(0.0)).astype(dtype)


# --------------------------------------------------------------------------------------
# Batcher  (data_processing.py:307-449, 600-619)
# --------------------------------------------------------------------------------------
def log_squasher(x):
    """data_processing.py:600-609: sign(x) * log1p(|x|)."""
    return np.sign(x) * np.log1p(np.abs(x))


def reverse_log_squasher(x):
    """data_processing.py:611-619."""
    return np.sign(x) * np.expm1(np.fabs(x))


def _get_seq(table, start, end, pad, stride, seq_len, nan_fill):
    """data_processing.py:370-398 (_get_train_seq) and :400-449 (_get_pred_seq, nan_fill=True).

    Returns the [seq_len, n_cols] float64 slab (first ``pad`` rows zero).  Training windows always
    have exactly seq_len - pad rows in start..end.  In the prediction variant with a missing target
    (tar_key != inp_key) the reference fills the slab with NaN and copies only the rows
    range(start, end + stride, stride) that exist (:428-435); ``end`` was not advanced by
    forecast_n for such windows (:276-279), so the tail rows stay NaN.
    """
    n_cols = table.shape[1]
    seq = np.zeros((seq_len, n_cols), dtype=np.float64)
    rows = start + stride * np.arange(seq_len - pad)
    ok = (rows <= end) & (rows < table.shape[0])
    if not nan_fill:
        assert ok.all(), (start, end, pad, stride, seq_len)
    body = np.full((seq_len - pad, n_cols), np.nan)
    body[ok] = table[rows[ok]]
    seq[pad:] = body
    return seq


def gather_batch(table, inp_idx, tar_idx, tar_valid, *, seq_len, stride, inp_cols, fin_cols,
                 seq_norm_col, center, scale, scale_inp_ids, aux_inp_ids=(), log_squash=True,
                 aux_masking=False, train=True):
    """data_processing.py:307-368 (get_batch) restated over a numeric table.

    table        float64 [n_rows, n_cols]  numeric view of Dataset.data_values
    inp_idx      int [B,3] (start, end, pad)      data_processing.py:267
    tar_idx      int [B,3]                        data_processing.py:271-279
    tar_valid    bool [B]  tar_key == inp_key     (only consulted when train=False, :417-435)
    seq_norm_col column of the scale field or None/0 (``if self._seq_norm_idx`` is falsy for 0, :393)
    Returns (inp f32 [B,T,F], tar f32 [B,T,O], seq_norm f64 [B]).
    """
    B = inp_idx.shape[0]
    n_fin = len(fin_cols)
    inp = np.empty((B, seq_len, len(inp_cols)))
    tar = np.empty((B, seq_len, n_fin))
    norms = np.empty(B)
    for i in range(B):
        s, e, p = (int(v) for v in inp_idx[i])
        seq = _get_seq(table, s, e, p, stride, seq_len, nan_fill=False)
        # data_processing.py:393-396 / :444-447
        seq_norm = max(seq[-1, seq_norm_col], _MIN_SEQ_NORM) if seq_norm_col else 1.0
        inp[i] = seq[:, inp_cols]
        s, e, p = (int(v) for v in tar_idx[i])
        if train or tar_valid[i]:
            tseq = _get_seq(table, s, e, p, stride, seq_len, nan_fill=False)
        else:
            tseq = _get_seq(table, s, e, p, stride, seq_len, nan_fill=True)
        tar[i] = tseq[:, fin_cols]
        # :341-346
        inp[i, :, :n_fin] /= seq_norm
        tar[i, :, :n_fin] /= seq_norm
        if log_squash:
            inp[i, :, :n_fin] = log_squasher(inp[i, :, :n_fin])
            tar[i, :, :n_fin] = log_squasher(tar[i, :, :n_fin])
        norms[i] = seq_norm
    # :352-357
    sid = list(scale_inp_ids)
    inp[:, :, sid] = (inp[:, :, sid] - center[sid]) / scale[sid]
    tar = (tar - center[:n_fin]) / scale[:n_fin]
    if aux_masking:  # :359-361
        inp[:, :seq_len - 1, list(aux_inp_ids)] = 0.0
    return inp.astype(np.float32), tar.astype(np.float32), norms


def create_window_index(keys, active, dates, *, train, stride, forecast_n, min_unrollings, max_unrollings,
                        start_date, end_date, last_train_date):
    """data_processing.py:203-305 as the reference writes it: one Python iteration per table row.

    keys [n] str, active [n] bool/int, dates [n] comparable (e.g. np.datetime64).  Returns
    (inp [N,3] int, tar [N,3] int, row_index [N]) for the rows that yield a window.
    """
    n = len(keys)
    min_steps = stride * (min_unrollings - 1) + 1
    max_steps = stride * (max_unrollings - 1) + 1
    last_key, cur_len = '', 1
    inp, tar, rows = [], [], []
    for i in range(n):
        key = keys[i]
        act = bool(int(active[i]))
        date = dates[i]
        tar_key = keys[i + forecast_n] if i + forecast_n <= n - 1 else ''
        if key != last_key:
            cur_len = 1
        if train:
            ok = cur_len >= min_steps and act and start_date <= date <= last_train_date and tar_key == key
        else:
            ok = cur_len >= min_steps and act and start_date <= date <= end_date
        if ok:
            seq_len = min(cur_len - (cur_len - 1) % stride, max_steps)        # :263
            pad = (max_steps - seq_len) // stride                              # :264
            inp.append([i - seq_len + 1, i, pad])
            if key == tar_key:
                tar.append([i - seq_len + 1 + forecast_n, i + forecast_n, pad])
            else:
                tar.append([i - seq_len + 1 + forecast_n, i, pad])
            rows.append(i)
        cur_len += 1
        last_key = key
    return np.array(inp, dtype=np.int64).reshape(-1, 3), np.array(tar, dtype=np.int64).reshape(-1, 3), \
        np.array(rows, dtype=np.int64)


# --------------------------------------------------------------------------------------
# Model forward  (rnn_point_estimate.py:76-107; Keras layer algorithms)
# --------------------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


N_GATES = {'lstm': 4, 'gru': 3}


def param_names(num_layers, rnn_cell='lstm', uq=False):
    """Keras ``model.trainable_variables`` order for RNNPointEstimate (rnn_point_estimate.py:76-105);
    layers are named lstm_k or gru_k after config.rnn_cell (:80-102)."""
    names = []
    for l in range(1, num_layers + 1):
        bn = 'batch_normalization' if l == 1 else 'batch_normalization_%d' % (l - 1)
        names += ['%s_%d/kernel' % (rnn_cell, l), '%s_%d/recurrent_kernel' % (rnn_cell, l),
                  '%s_%d/bias' % (rnn_cell, l), bn + '/gamma', bn + '/beta']
    if uq:      # RNNUqRangeEstimate (rnn_uq_range_estimate.py:104-108): target head, then variance head
        names += ['OUTPUT_TARGET_1/kernel', 'OUTPUT_TARGET_1/bias', 'OUTPUT_VARIANCE_1/kernel', 'OUTPUT_VARIANCE_1/bias']
    else:
        names += ['OUTPUT_1/kernel', 'OUTPUT_1/bias']
    return names


def param_shapes(num_layers, n_inputs, n_outputs, num_hidden, rnn_cell='lstm', uq=False):
    """GRU (Keras default reset_after=True): 3 gate blocks z|r|h and a [2, 3H] bias (input row, recurrent row)."""
    H = num_hidden
    G = N_GATES[rnn_cell]
    shapes = []
    for l in range(num_layers):
        I = n_inputs if l == 0 else H
        shapes += [(I, G * H), (H, G * H), (4 * H,) if rnn_cell == 'lstm' else (2, 3 * H), (H,), (H,)]
    shapes += [(H, n_outputs), (n_outputs,)] * (2 if uq else 1)
    return shapes


def init_params(num_layers, n_inputs, n_outputs, num_hidden, init_scale=1.0, seed=521,
                dtype=np.float32, rnn_cell='lstm', uq=False):
    """Initial weights in the reference's distribution families (initializers.py:14-24 for the
    LSTM kernel; Keras defaults for the rest: orthogonal recurrent kernel, unit forget bias,
    gamma=1, beta=0, Glorot-uniform Dense).  TF's RNG streams cannot be reproduced (SURVEY
    App. B #7) so parity tests always *inject* these numpy-drawn weights on both sides."""
    rng = np.random.RandomState(seed)
    H = num_hidden
    G = N_GATES[rnn_cell]
    out = []
    for l in range(num_layers):
        I = n_inputs if l == 0 else H
        out.append(rng.uniform(-init_scale, init_scale, size=(I, G * H)))
        a = rng.normal(size=(G * H, H))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        out.append(q.T.copy())                       # [H, G*H], orthonormal rows
        if rnn_cell == 'lstm':
            b = np.zeros(4 * H)
            b[H:2 * H] = 1.0                         # unit_forget_bias
        else:
            b = np.zeros((2, 3 * H))                 # Keras GRU: bias_initializer='zeros'
        out.append(b)
        out.append(np.ones(H))
        out.append(np.zeros(H))
    lim = np.sqrt(6.0 / (H + n_outputs))
    for _ in range(2 if uq else 1):
        out.append(rng.uniform(-lim, lim, size=(H, n_outputs)))
        out.append(np.zeros(n_outputs))
    return [p.astype(dtype) for p in out]


def lstm_forward(x, W, U, b, rec_mask=None):
    """Keras LSTM(return_sequences=True), implementation=2 (rnn_point_estimate.py:80-87).

    x [B,T,I]; W [I,4H]; U [H,4H]; b [4H], gate blocks i|f|c|o; h0=c0=0.
    rec_mask [B,H] multiplies h_{t-1} before the recurrent matmul (recurrent_dropout), or None.
    Returns h [B,T,H] and a cache for ``lstm_backward``.
    """
    B, T, I = x.shape
    H = U.shape[0]
    dt = x.dtype
    h = np.zeros((B, H), dtype=dt)
    c = np.zeros((B, H), dtype=dt)
    hs = np.empty((B, T, H), dtype=dt)
    cs = np.empty((B, T, H), dtype=dt)
    gates = np.empty((B, T, 4 * H), dtype=dt)
    for t in range(T):
        hm = h if rec_mask is None else h * rec_mask
        z = x[:, t, :] @ W + hm @ U + b
        i = sigmoid(z[:, :H])
        f = sigmoid(z[:, H:2 * H])
        g = np.tanh(z[:, 2 * H:3 * H])
        o = sigmoid(z[:, 3 * H:])
        c = f * c + i * g
        h = o * np.tanh(c)
        hs[:, t] = h
        cs[:, t] = c
        gates[:, t, :H] = i
        gates[:, t, H:2 * H] = f
        gates[:, t, 2 * H:3 * H] = g
        gates[:, t, 3 * H:] = o
    return hs, (x, W, U, hs, cs, gates, rec_mask)


def lstm_backward(dh_out, cache, need_dx):
    """BPTT through ``lstm_forward`` (what tape.gradient does at train.py:192)."""
    x, W, U, hs, cs, gates, rec_mask = cache
    B, T, I = x.shape
    H = U.shape[0]
    dt = x.dtype
    dz_all = np.empty((B, T, 4 * H), dtype=dt)
    dh_next = np.zeros((B, H), dtype=dt)
    dc_next = np.zeros((B, H), dtype=dt)
    for t in range(T - 1, -1, -1):
        i = gates[:, t, :H]
        f = gates[:, t, H:2 * H]
        g = gates[:, t, 2 * H:3 * H]
        o = gates[:, t, 3 * H:]
        c_prev = cs[:, t - 1] if t > 0 else np.zeros((B, H), dtype=dt)
        tc = np.tanh(cs[:, t])
        dh = dh_out[:, t] + dh_next
        do = dh * tc
        dc = dc_next + dh * o * (1.0 - tc * tc)
        di = dc * g
        dg = dc * i
        df = dc * c_prev
        dc_next = dc * f
        dz = np.concatenate([di * i * (1 - i), df * f * (1 - f), dg * (1 - g * g), do * o * (1 - o)], axis=1)
        dz_all[:, t] = dz
        dh_next = dz @ U.T
        if rec_mask is not None:
            dh_next = dh_next * rec_mask
    h_prev = np.concatenate([np.zeros((B, 1, H), dtype=dt), hs[:, :-1]], axis=1)
    if rec_mask is not None:
        h_prev = h_prev * rec_mask[:, None, :]
    dz2 = dz_all.reshape(B * T, 4 * H)
    dW = x.reshape(B * T, I).T @ dz2
    dU = h_prev.reshape(B * T, H).T @ dz2
    db = dz2.sum(axis=0)
    dx = (dz2 @ W.T).reshape(B, T, I) if need_dx else None
    return dW, dU, db, dx


def gru_forward(x, W, U, b2, rec_mask=None):
    """Keras GRU(return_sequences=True), reset_after=True (the TF 2.x default), implementation=2
    (rnn_point_estimate.py:89-98).  Restated from the published cell equations [EXT, SURVEY 8f-1]:

        xz = x_t W + b2[0];  hz = (h_{t-1} * mask) U + b2[1]          gate blocks z|r|h
        z = sigmoid(xz_z + hz_z);  r = sigmoid(xz_r + hz_r);  hh = tanh(xz_h + r * hz_h)
        h_t = z * h_{t-1} + (1 - z) * hh                              (the carry uses the unmasked h_{t-1})

    Pinned against torch.nn.GRU (same equations, gate order r|z|n) in tests/test_oracle.py.
    """
    B, T, I = x.shape
    H = U.shape[0]
    dt = x.dtype
    h = np.zeros((B, H), dtype=dt)
    hs = np.empty((B, T, H), dtype=dt)
    gates = np.empty((B, T, 4 * H), dtype=dt)         # z | r | hh | q = hz_h (kept for the backward pass)
    for t in range(T):
        hm = h if rec_mask is None else h * rec_mask
        xz = x[:, t, :] @ W + b2[0]
        hz = hm @ U + b2[1]
        z = sigmoid(xz[:, :H] + hz[:, :H])
        r = sigmoid(xz[:, H:2 * H] + hz[:, H:2 * H])
        q = hz[:, 2 * H:]
        hh = np.tanh(xz[:, 2 * H:] + r * q)
        h = z * h + (1.0 - z) * hh
        hs[:, t] = h
        gates[:, t, :H] = z
        gates[:, t, H:2 * H] = r
        gates[:, t, 2 * H:3 * H] = hh
        gates[:, t, 3 * H:] = q
    return hs, (x, W, U, hs, gates, rec_mask)


def gru_backward(dh_out, cache, need_dx):
    """BPTT through ``gru_forward``.  Returns dW, dU, db [2,3H], dx."""
    x, W, U, hs, gates, rec_mask = cache
    B, T, I = x.shape
    H = U.shape[0]
    dt = x.dtype
    dxz_all = np.empty((B, T, 3 * H), dtype=dt)
    dhz_all = np.empty((B, T, 3 * H), dtype=dt)
    dh_carry = np.zeros((B, H), dtype=dt)             # through z * h_{t-1}
    dh_rec = np.zeros((B, H), dtype=dt)               # through the recurrent matmul
    for t in range(T - 1, -1, -1):
        z = gates[:, t, :H]
        r = gates[:, t, H:2 * H]
        hh = gates[:, t, 2 * H:3 * H]
        q = gates[:, t, 3 * H:]
        h_prev = hs[:, t - 1] if t > 0 else np.zeros((B, H), dtype=dt)
        dh = dh_out[:, t] + dh_carry + dh_rec
        da_h = dh * (1.0 - z) * (1.0 - hh * hh)
        da_z = dh * (h_prev - hh) * z * (1.0 - z)
        da_r = da_h * q * r * (1.0 - r)
        dh_carry = dh * z
        dxz_all[:, t] = np.concatenate([da_z, da_r, da_h], axis=1)
        dhz = np.concatenate([da_z, da_r, da_h * r], axis=1)
        dhz_all[:, t] = dhz
        dh_rec = dhz @ U.T
        if rec_mask is not None:
            dh_rec = dh_rec * rec_mask
    h_prev = np.concatenate([np.zeros((B, 1, H), dtype=dt), hs[:, :-1]], axis=1)
    if rec_mask is not None:
        h_prev = h_prev * rec_mask[:, None, :]
    dxz2 = dxz_all.reshape(B * T, 3 * H)
    dhz2 = dhz_all.reshape(B * T, 3 * H)
    dW = x.reshape(B * T, I).T @ dxz2
    dU = h_prev.reshape(B * T, H).T @ dhz2
    db = np.stack([dxz2.sum(axis=0), dhz2.sum(axis=0)])
    dx = (dxz2 @ W.T).reshape(B, T, I) if need_dx else None
    return dW, dU, db, dx


def _trunk(params, x, *, num_layers, dropout, recurrent_dropout, training, seed, step, row0, bn_mean, bn_var,
           rnn_cell, layer0=0):
    """The recurrent stack shared by RNNPointEstimate and RNNUqRangeEstimate: per layer LSTM|GRU ->
    BatchNormalization (inference-mode affine) -> Dropout."""
    B, T, _ = x.shape
    dt = x.dtype
    cur = x
    caches = []
    for l in range(num_layers):
        W, U, b, gamma, beta = params[5 * l:5 * l + 5]
        H = U.shape[0]
        mean = np.zeros(H, dtype=dt) if bn_mean is None else bn_mean[l].astype(dt)
        var = np.ones(H, dtype=dt) if bn_var is None else bn_var[l].astype(dt)
        rmask = None
        if training and recurrent_dropout > 0.0:
            rmask = dropout_mask(seed, step, 2 * (layer0 + l) + 1, row0, B, H, recurrent_dropout, dtype=dt.type)
        hs, lc = (lstm_forward if rnn_cell == 'lstm' else gru_forward)(cur, W, U, b, rmask)
        inv = (1.0 / np.sqrt(var + dt.type(BN_EPS))).astype(dt)
        y = gamma * (hs - mean) * inv + beta
        dmask = None
        if training and dropout > 0.0:
            dmask = dropout_mask(seed, step, 2 * (layer0 + l), row0, B, T * H, dropout, dtype=dt.type).reshape(B, T, H)
            y = y * dmask
        caches.append((lc, hs, mean, inv, gamma, dmask))
        cur = y
    return cur, caches


def forward(params, x, *, num_layers, dropout=0.0, recurrent_dropout=0.0, training=False,
            seed=0, step=0, row0=0, bn_mean=None, bn_var=None, rnn_cell='lstm'):
    """model(inp) for RNNPointEstimate, forecast_steps=1 (rnn_point_estimate.py:66-107).

    BatchNormalization runs in inference mode in *both* train and predict;
    Dropout / recurrent dropout are active iff ``training`` (= config.train, :87,89).
    Returns preds [B,T,O] and the cache for ``backward``.
    """
    cur, caches = _trunk(params, x, num_layers=num_layers, dropout=dropout, recurrent_dropout=recurrent_dropout,
                         training=training, seed=seed, step=step, row0=row0, bn_mean=bn_mean, bn_var=bn_var,
                         rnn_cell=rnn_cell)
    Wo, bo = params[5 * num_layers], params[5 * num_layers + 1]
    preds = cur @ Wo + bo
    return preds, (caches, cur, Wo)


VAR_FLOOR = 1e-6


def softplus(a):
    return np.logaddexp(a, 0.0)


def forward_uq(params, x, *, num_layers, dropout=0.0, recurrent_dropout=0.0, seed=0, step=0, row0=0,
               bn_mean=None, bn_var=None, rnn_cell='lstm'):
    """model(inp) for RNNUqRangeEstimate, forecast_steps=1 (rnn_uq_range_estimate.py:66-110): the same trunk with
    Dropout / recurrent dropout ALWAYS active (``training=True`` is a literal there, :86,88 -- MC dropout, also in
    predict), a target head and a variance head ``max(softplus(.), 1e-6)`` (model_utils/custom_layers.py:12-13).
    Returns (pred [B,T,O], var [B,T,O]) and the cache for ``backward_uq``."""
    cur, caches = _trunk(params, x, num_layers=num_layers, dropout=dropout, recurrent_dropout=recurrent_dropout,
                         training=True, seed=seed, step=step, row0=row0, bn_mean=bn_mean, bn_var=bn_var,
                         rnn_cell=rnn_cell)
    Wt, bt, Wv, bv = params[5 * num_layers:5 * num_layers + 4]
    pred = cur @ Wt + bt
    a = cur @ Wv + bv
    var = np.maximum(softplus(a), x.dtype.type(VAR_FLOOR))
    return pred, var, (caches, cur, Wt, Wv, a)


def loss_uq_estimate(y_true, y_pred, y_var, *, target_idx, target_lambda, rnn_lambda):
    """losses.py:180-247 + :249-284, 'RNN' branch.  Returns (uq_loss, uq_loss_last_tar, mse_0, dLoss/dpred, dLoss/dvar).

    Per element the loss term is (p*m - y)^2 / (v*m) + log(v*m) with m the [b,t] mask (:197-201, :272): a zero-padded
    step therefore contributes 0 * inf + log 0 = NaN and the reference's loss (and every gradient) is NaN as soon as a
    batch holds one -- reproduced here by plain IEEE arithmetic, not special-cased.  Denominators: number of unmasked
    rows in the slice times the slice's last dimension (:270-271), i.e. 1 for the target-field slice, O otherwise."""
    B, T, K = y_true.shape
    dt = y_pred.dtype
    mask = (~np.all(y_true == 0.0, axis=-1)).astype(dt)
    m3 = mask[..., None]
    pm, vm = y_pred * m3, y_var * m3
    with np.errstate(divide='ignore', invalid='ignore'):
        diff = (pm - y_true) ** 2
        term = diff * (1.0 / vm) + np.log(vm)
        ms_last = mask[:, -1].sum()
        ms_all = mask.sum()
        uq0 = np.sum(term[:, -1, target_idx]) / (ms_last * 1)
        uq1 = np.sum(term[:, -1, :]) / (ms_last * K)
        uq2 = np.sum(term) / (ms_all * K)
        mse_0 = np.mean((y_true[:, -1, target_idx] - pm[:, -1, target_idx]) ** 2)
        p1, p2 = dt.type(target_lambda), dt.type(rnn_lambda)
        loss = p1 * uq0 + (1 - p1) * (p2 * uq1 + (1 - p2) * uq2)
        coef = np.full((B, T, K), (1 - p1) * (1 - p2) / (ms_all * K), dtype=dt)
        coef[:, -1, :] += (1 - p1) * p2 / (ms_last * K)
        coef[:, -1, target_idx] += p1 / ms_last
        dpred = coef * (2.0 * (pm - y_true) / vm) * m3
        dvar = coef * (-diff / (vm * vm) + 1.0 / vm) * m3
    return loss, uq0, mse_0, dpred.astype(dt), dvar.astype(dt)


def loss_point_estimate(y_true, y_pred, *, target_idx, target_lambda, rnn_lambda,
                        batch_global=None, mask_count_global=None):
    """losses.py:55-98 + :121-135 for the 'RNN' branch.  Returns (loss, mse_0, dloss/dpred, sums).

    ``batch_global`` / ``mask_count_global`` replace the local denominators under data
    parallelism; the returned loss/mse are then this shard's *contribution*.
    """
    B, T, K = y_true.shape
    dt = y_pred.dtype
    mask = (~np.all(y_true == 0.0, axis=-1)).astype(dt)            # losses.py:72-73
    yp = y_pred * mask[..., None]                                   # :75
    Bg = dt.type(B if batch_global is None else batch_global)
    Mg = dt.type(mask.sum() if mask_count_global is None else mask_count_global)
    d_last = yp[:, -1, :] - y_true[:, -1, :]
    s0 = np.sum(d_last[:, target_idx] ** 2)
    s1 = np.sum(d_last ** 2)
    d_all = yp - y_true
    s2 = np.sum(d_all ** 2)
    mse_0 = s0 / Bg                                                 # :87
    mse_1 = s1 / (Bg * K)                                           # :90
    mse_2 = s2 / (Mg * K)                                           # :131-135
    p1, p2 = dt.type(target_lambda), dt.type(rnn_lambda)
    loss = p1 * mse_0 + (1 - p1) * (p2 * mse_1 + (1 - p2) * mse_2)  # :98
    c_all = (1 - p1) * (1 - p2) / (K * Mg)
    c_last = (1 - p1) * p2 / (Bg * K)
    c_tar = p1 / Bg
    coef = np.full((B, T, K), c_all, dtype=dt)
    coef[:, -1, :] += c_last
    coef[:, -1, target_idx] += c_tar
    dpred = (2.0 * d_all * coef * mask[..., None]).astype(dt)
    return loss, mse_0, dpred, (s0, s1, s2, mask.sum())


def _trunk_backward(dy, caches, grads, *, num_layers, rnn_cell):
    for l in range(num_layers - 1, -1, -1):
        lc, hs, mean, inv, gamma, dmask = caches[l]
        if dmask is not None:
            dy = dy * dmask
        grads[5 * l + 3] = np.sum(dy * (hs - mean) * inv, axis=(0, 1))
        grads[5 * l + 4] = np.sum(dy, axis=(0, 1))
        dh_out = dy * gamma * inv
        dW, dU, db, dx = (lstm_backward if rnn_cell == 'lstm' else gru_backward)(dh_out, lc, need_dx=(l > 0))
        grads[5 * l], grads[5 * l + 1], grads[5 * l + 2] = dW, dU, db
        dy = dx
    return grads


def backward(dpred, fcache, *, num_layers, rnn_cell='lstm'):
    """Gradients for every trainable variable, Keras order (train.py:192)."""
    caches, y_last, Wo = fcache
    B, T, O = dpred.shape
    H = Wo.shape[0]
    grads = [None] * (5 * num_layers + 2)
    grads[5 * num_layers] = y_last.reshape(B * T, H).T @ dpred.reshape(B * T, O)
    grads[5 * num_layers + 1] = dpred.reshape(B * T, O).sum(axis=0)
    dy = dpred @ Wo.T
    return _trunk_backward(dy, caches, grads, num_layers=num_layers, rnn_cell=rnn_cell)


def backward_uq(dpred, dvar, fcache, *, num_layers, rnn_cell='lstm'):
    """tape.gradient(uq_loss, trainable_variables) (train.py:218).  The variance head's activation passes the gradient
    through softplus where it is above the 1e-6 floor and blocks it below (tf.maximum)."""
    caches, y_last, Wt, Wv, a = fcache
    B, T, O = dpred.shape
    H = Wt.shape[0]
    da = dvar * np.where(softplus(a) > VAR_FLOOR, sigmoid(a), 0.0)
    grads = [None] * (5 * num_layers + 4)
    y2 = y_last.reshape(B * T, H)
    grads[5 * num_layers] = y2.T @ dpred.reshape(B * T, O)
    grads[5 * num_layers + 1] = dpred.reshape(B * T, O).sum(axis=0)
    grads[5 * num_layers + 2] = y2.T @ da.reshape(B * T, O)
    grads[5 * num_layers + 3] = da.reshape(B * T, O).sum(axis=0)
    dy = dpred @ Wt.T + da @ Wv.T
    return _trunk_backward(dy, caches, grads, num_layers=num_layers, rnn_cell=rnn_cell)


# --------------------------------------------------------------------------------------
# forecast_steps > 1 (rnn_point_estimate.py:109-150, model_base_class.py:18-51).  Oracle only so far: the CUDA path
# refuses forecast_steps != 1; this is the checker the next build step is written against.
# --------------------------------------------------------------------------------------
def forecast_param_names(num_layers, forecast_steps, rnn_cell='lstm'):
    """trainable_variables order: the trunk and OUTPUT_1 as for forecast_steps = 1, then per extra step s one
    recurrent layer (numbered on from the trunk, :129), its BatchNormalization and the Dense OUTPUT_{s+1} (:150)."""
    names = param_names(num_layers, rnn_cell)
    for s in range(1, forecast_steps):
        l = num_layers + s
        names += ['%s_%d/kernel' % (rnn_cell, l), '%s_%d/recurrent_kernel' % (rnn_cell, l), '%s_%d/bias' % (rnn_cell, l),
                  'batch_normalization_%d/gamma' % (l - 1), 'batch_normalization_%d/beta' % (l - 1),
                  'OUTPUT_%d/kernel' % (s + 1), 'OUTPUT_%d/bias' % (s + 1)]
    return names


def init_forecast_params(num_layers, n_inputs, n_outputs, num_hidden, forecast_steps, init_scale=1.0, seed=521,
                         dtype=np.float32, rnn_cell='lstm'):
    out = init_params(num_layers, n_inputs, n_outputs, num_hidden, init_scale, seed, dtype, rnn_cell)
    for s in range(1, forecast_steps):
        extra = init_params(1, n_inputs, n_outputs, num_hidden, init_scale, seed + 7 * s, dtype, rnn_cell)
        out += extra                      # one recurrent layer on the raw feature width + its head
    return out


def forward_forecast(params, x, *, num_layers, forecast_steps, dropout=0.0, recurrent_dropout=0.0, training=False,
                     seed=0, step=0, row0=0, rnn_cell='lstm'):
    """model(inp) -> [pred_1, ..., pred_S].  Step s >= 2 runs ONE new recurrent layer over the input window shifted by
    one: the first time step is dropped and [pred_{s-1}[:, -1, :], aux features of the LAST ORIGINAL time step]
    (model_base_class.py:18-41: the last n_inputs - n_outputs columns) is appended (rnn_point_estimate.py:113-124)."""
    B, T, F = x.shape
    O = params[5 * num_layers].shape[1]
    n0 = 5 * num_layers + 2
    preds, caches = [], []
    p0, fc0 = forward(params[:n0], x, num_layers=num_layers, dropout=dropout, recurrent_dropout=recurrent_dropout,
                      training=training, seed=seed, step=step, row0=row0, rnn_cell=rnn_cell)
    preds.append(p0)
    caches.append(fc0)
    aux = x[:, -1:, O:]
    prev_input = x
    for s in range(1, forecast_steps):
        new_step = np.concatenate([preds[-1][:, -1:, :], aux], axis=2)
        cur_input = np.concatenate([prev_input, new_step], axis=1)[:, 1:, :]
        prev_input = cur_input
        ps = params[n0 + 7 * (s - 1):n0 + 7 * s]
        # the Philox stream index continues after the trunk's layers: layer num_layers + s - 1
        cur, lc = _trunk(ps[:5], cur_input, num_layers=1, dropout=dropout, recurrent_dropout=recurrent_dropout,
                         training=training, seed=seed, step=step, row0=row0, bn_mean=None, bn_var=None,
                         rnn_cell=rnn_cell, layer0=num_layers + s - 1)
        preds.append(cur @ ps[5] + ps[6])
        caches.append((lc, cur, ps[5]))
    return preds, (caches, num_layers, forecast_steps, O)


def loss_forecast(y_trues, preds, weights, **kw):
    """Losses.weight_adjusted_mse (losses.py:19-53): sum_s w_s * loss_s, sum_s w_s * mse_s."""
    loss = mse = 0.0
    dpreds = []
    for y, p, w in zip(y_trues, preds, weights):
        l, m, dp, _ = loss_point_estimate(y.astype(p.dtype), p, **kw)
        loss += w * l
        mse += w * m
        dpreds.append(w * dp)
    return loss, mse, dpreds


def backward_forecast(dpreds, fcache, *, rnn_cell='lstm'):
    """Gradients in forecast_param_names order.  The window of step s holds, at position T-1-j, the appended step of
    forecast s-j (j = 0..s-1), whose first O columns are pred_{s-j}[:, -1, :]: the input gradient of the extra layer
    flows back into those earlier predictions' last time step."""
    caches, num_layers, S, O = fcache
    dpreds = [d.copy() for d in dpreds]
    B, T, _ = dpreds[0].shape
    extra = []
    for s in range(S - 1, 0, -1):
        lc, y_last, Wo = caches[s]
        dp = dpreds[s]
        H = Wo.shape[0]
        g = [None] * 7
        g[5] = y_last.reshape(B * T, H).T @ dp.reshape(B * T, O)
        g[6] = dp.reshape(B * T, O).sum(axis=0)
        dy = dp @ Wo.T
        (lcache, hs, mean, inv, gamma, dmask) = lc[0]
        if dmask is not None:
            dy = dy * dmask
        g[3] = np.sum(dy * (hs - mean) * inv, axis=(0, 1))
        g[4] = np.sum(dy, axis=(0, 1))
        dW, dU, db, dx = (lstm_backward if rnn_cell == 'lstm' else gru_backward)(dy * gamma * inv, lcache, need_dx=True)
        g[0], g[1], g[2] = dW, dU, db
        for j in range(s):                       # appended steps inside this window
            dpreds[s - 1 - j][:, -1, :] += dx[:, T - 1 - j, :O]
        extra = g + extra
    g0 = backward(dpreds[0], caches[0], num_layers=num_layers, rnn_cell=rnn_cell)
    return g0 + extra


# --------------------------------------------------------------------------------------
# Step tail: clip, LR schedule, optimizers, MaxNorm (train.py:195-198, optimizers.py:15-54,
# rnn_point_estimate.py:85).  Keras/TF 2.x update rules, SURVEY App. A.5.
# --------------------------------------------------------------------------------------
def clip_by_global_norm(grads, clip_norm):
    """tf.clip_by_global_norm (train.py:196): g * clip / max(||g||, clip)."""
    dt = grads[0].dtype
    gn = np.sqrt(sum(np.sum(g.astype(dt) ** 2) for g in grads))
    s = dt.type(clip_norm) / max(gn, dt.type(clip_norm))
    return [g * s for g in grads], gn


def learning_rate(it, *, lr_schedule='ExponentialDecay', learning_rate=0.6, lr_decay=1.0,
                  decay_steps=1500, end_learning_rate=0.01, decay_power=0.5,
                  piecewise_lr_boundaries=(), piecewise_lr_values=()):
    """optimizers.py:31-54.  ``it`` = optimizer.iterations before this update."""
    if lr_schedule == 'ExponentialDecay':          # staircase=True, :37-41
        return learning_rate * lr_decay ** np.floor(it / decay_steps)
    if lr_schedule == 'PolynomialDecay':           # :42-46 (cycle=False)
        s = min(it, decay_steps)
        return (learning_rate - end_learning_rate) * (1 - s / decay_steps) ** decay_power + end_learning_rate
    if lr_schedule == 'PiecewiseConstantDecay':    # :47-49: values[i] for boundaries[i-1] < it <= boundaries[i]
        for bnd, v in zip(piecewise_lr_boundaries, piecewise_lr_values):
            if it <= bnd:
                return v
        return piecewise_lr_values[len(piecewise_lr_boundaries)]
    raise ValueError('Invalid learning rate scheduler specified')


OPT_SLOTS = {'Adadelta': 2, 'Adam': 2, 'RMSprop': 1, 'SGD': 1}


def optimizer_update(name, params, grads, slots, lr, it, *, sgd_momentum=0.0):
    """One apply_gradients (train.py:198) with Keras defaults (optimizers.py:21-27).

    ``slots`` is a list (per slot) of lists (per variable) of arrays, updated in place; ``it`` is
    the iteration count before this update.  Returns the new params.
    """
    dt = params[0].dtype
    lr = dt.type(lr)
    new = []
    for j, (p, g) in enumerate(zip(params, grads)):
        g = g.astype(dt)
        if name == 'Adadelta':                     # rho=0.95, eps=1e-7
            rho, eps = dt.type(0.95), dt.type(1e-7)
            a, d = slots[0][j], slots[1][j]
            a[...] = rho * a + (1 - rho) * g * g
            upd = np.sqrt(d + eps) / np.sqrt(a + eps) * g
            d[...] = rho * d + (1 - rho) * upd * upd
            new.append(p - lr * upd)
        elif name == 'Adam':                       # b1=.9, b2=.999, eps=1e-7
            b1, b2, eps = dt.type(0.9), dt.type(0.999), dt.type(1e-7)
            m, v = slots[0][j], slots[1][j]
            t = it + 1
            lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
            m[...] = b1 * m + (1 - b1) * g
            v[...] = b2 * v + (1 - b2) * g * g
            new.append(p - dt.type(lr_t) * m / (np.sqrt(v) + eps))
        elif name == 'RMSprop':                    # rho=.9, eps=1e-7, momentum=0, centered=False
            rho, eps = dt.type(0.9), dt.type(1e-7)
            v = slots[0][j]
            v[...] = rho * v + (1 - rho) * g * g
            new.append(p - lr * g / (np.sqrt(v) + eps))
        elif name == 'SGD':                        # optimizers.py:27
            mom = dt.type(sgd_momentum)
            if sgd_momentum > 0.0:
                m = slots[0][j]
                m[...] = mom * m - lr * g
                new.append(p + m)
            else:
                new.append(p - lr * g)
        else:
            raise ValueError('%s optimizer not found in tf.keras.optimizers' % name)
    return new


def max_norm_constraint(w, max_norm):
    """keras.constraints.MaxNorm(max_value, axis=0) (rnn_point_estimate.py:85)."""
    dt = w.dtype
    norms = np.sqrt(np.sum(w * w, axis=0, keepdims=True))
    desired = np.clip(norms, 0, dt.type(max_norm))
    return w * (desired / (dt.type(1e-7) + norms))


def zero_slots(name, params):
    return [[np.zeros_like(p) for p in params] for _ in range(OPT_SLOTS[name])]


def train_step(params, slots, x, y, it, cfg, *, row0=0, batch_global=None, mask_count_global=None,
               lr=None):
    """Train._train_step_point (train.py:178-199): fwd -> loss -> grads -> clip -> apply -> MaxNorm.

    cfg: dict with num_layers, dropout, recurrent_dropout, train, seed, target_idx, target_lambda,
    rnn_lambda, max_grad_norm, optimizer, max_norm, sgd_momentum and the LR-schedule keys.
    Returns (new_params, mse_0, loss, grads_before_clip, grad_norm).
    """
    L = cfg['num_layers']
    if cfg.get('uq', False):
        return _train_step_uq(params, slots, x, y, it, cfg, row0=row0, lr=lr)
    preds, fc = forward(params, x, num_layers=L, dropout=cfg.get('dropout', 0.0),
                        recurrent_dropout=cfg.get('recurrent_dropout', 0.0),
                        training=cfg.get('train', True), seed=cfg.get('seed', 0), step=it, row0=row0,
                        rnn_cell=cfg.get('rnn_cell', 'lstm'))
    loss, mse, dpred, _ = loss_point_estimate(y.astype(preds.dtype), preds, target_idx=cfg['target_idx'],
                                              target_lambda=cfg['target_lambda'], rnn_lambda=cfg['rnn_lambda'],
                                              batch_global=batch_global, mask_count_global=mask_count_global)
    grads = backward(dpred, fc, num_layers=L, rnn_cell=cfg.get('rnn_cell', 'lstm'))
    raw = [g.copy() for g in grads]
    gn = None
    if cfg.get('max_grad_norm', 0.0) > 0:
        grads, gn = clip_by_global_norm(grads, cfg['max_grad_norm'])
    if lr is None:
        keys = ('lr_schedule', 'learning_rate', 'lr_decay', 'decay_steps', 'end_learning_rate', 'decay_power',
                'piecewise_lr_boundaries', 'piecewise_lr_values')
        lr = learning_rate(it, **{k: cfg[k] for k in keys if k in cfg})
    new = optimizer_update(cfg.get('optimizer', 'Adadelta'), params, grads, slots, lr, it,
                           sgd_momentum=cfg.get('sgd_momentum', 0.0))
    for l in range(L):
        new[5 * l] = max_norm_constraint(new[5 * l], cfg.get('max_norm', 3))
    return new, mse, loss, raw, gn



def _train_step_uq(params, slots, x, y, it, cfg, *, row0=0, lr=None):
    """Train._train_step_uq_range (train.py:201-225).  Returns (new_params, mse_0, uq_loss, grads_before_clip,
    grad_norm, uq_loss_last_tar); the reference's step returns (uq_loss_last_tar, mse)."""
    L = cfg['num_layers']
    cell = cfg.get('rnn_cell', 'lstm')
    pred, var, fc = forward_uq(params, x, num_layers=L, dropout=cfg.get('dropout', 0.0),
                               recurrent_dropout=cfg.get('recurrent_dropout', 0.0), seed=cfg.get('seed', 0), step=it,
                               row0=row0, rnn_cell=cell)
    loss, uq0, mse, dpred, dvar = loss_uq_estimate(y.astype(pred.dtype), pred, var, target_idx=cfg['target_idx'],
                                                   target_lambda=cfg['target_lambda'], rnn_lambda=cfg['rnn_lambda'])
    grads = backward_uq(dpred, dvar, fc, num_layers=L, rnn_cell=cell)
    raw = [g.copy() for g in grads]
    gn = None
    if cfg.get('max_grad_norm', 0.0) > 0:
        grads, gn = clip_by_global_norm(grads, cfg['max_grad_norm'])
    if lr is None:
        keys = ('lr_schedule', 'learning_rate', 'lr_decay', 'decay_steps', 'end_learning_rate', 'decay_power',
                'piecewise_lr_boundaries', 'piecewise_lr_values')
        lr = learning_rate(it, **{k: cfg[k] for k in keys if k in cfg})
    new = optimizer_update(cfg.get('optimizer', 'Adadelta'), params, grads, slots, lr, it,
                           sgd_momentum=cfg.get('sgd_momentum', 0.0))
    for l in range(L):
        new[5 * l] = max_norm_constraint(new[5 * l], cfg.get('max_norm', 3))
    return new, mse, loss, raw, gn, uq0
