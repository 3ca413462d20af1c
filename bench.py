#!/usr/bin/env python
"""Headline benchmark: train-step sequences/sec of the recurrent forecaster (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision bf16|fp32]

N=1 workload = BASELINE.json configs[1]: synthetic B=4096, T=48, F=32, O=16, 1-layer LSTM H=256, one full
Train._train_step_point (fwd, weighted-MSE loss, BPTT, clip, Adadelta, MaxNorm).  N>1 = configs[3]: the
company-batch axis sharded, 4096 windows per rank (weak scaling), ONE NCCL all-reduce of the flat gradient.
Prints ONE JSON line on rank 0 (contract in the task statement: value, e2e, roofline, cpu_baseline, clocks ...).

`--impl reference` times the reference's CPU path.  The reference is TensorFlow/Keras Python and TensorFlow is
not installable here (no wheel in /opt/wheelhouse, no network), so that arm runs the NumPy restatement of the
same step (oracle/, kind "port") on all host cores -- labelled as such.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'oracle')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

T, F, O, H, L = 48, 32, 16, 256, 1
TARGET_IDX = 3
SEED = 521                       # reference default, scripts/lfm_quant.py:72
METRIC = 'train-step sequences/sec at B=4096,T=48,H=256'
# algorithmic gate-GEMM FLOPs per window for this config (fwd 2*T*(F+H)*4H, bwd-weight the same,
# bwd-data 2*T*4H*H)
FLOP_FWD_PER_SEQ = 2.0 * T * (F + H) * 4 * H
FLOP_TRAIN_PER_SEQ = 2 * FLOP_FWD_PER_SEQ + 2.0 * T * 4 * H * H


def synthetic(batch, rng):
    x = rng.standard_normal((batch, T, F), dtype=np.float32)
    y = rng.standard_normal((batch, T, O), dtype=np.float32)
    return x, y


def initial_weights():
    import lfm_oracle as orc
    # W~U(-1,1) (init_scale=1.0, lfm_quant.py:54), U orthogonal, b=[0,1,0,0], gamma=1, beta=0, Glorot head
    return orc.init_params(L, F, O, H, init_scale=1.0, seed=SEED, dtype=np.float32)


def oracle_cfg():
    return dict(num_layers=L, target_idx=TARGET_IDX, target_lambda=0.5, rnn_lambda=0.7, max_grad_norm=50.0,
                optimizer='Adadelta', max_norm=3.0, train=True, dropout=0.0, recurrent_dropout=0.0)


def cpu_port_seq_per_s(sample_rows, steps, threads):
    """Times oracle.train_step (fp32 NumPy) on a bounded sample of the workload.  The BLAS thread count is scanned
    over a few values up to the host's core count and the best one is kept (small per-step matmuls do not scale to
    128 threads); returns (windows/s, seconds/step, threads used)."""
    import lfm_oracle as orc
    from threadpoolctl import threadpool_limits
    rng = np.random.default_rng(SEED)
    x, y = synthetic(sample_rows, rng)
    cfg = oracle_cfg()

    def run(n_threads, n_steps):
        with threadpool_limits(limits=n_threads):
            params = initial_weights()
            slots = orc.zero_slots('Adadelta', params)
            params, *_ = orc.train_step(params, slots, x, y, 0, cfg, lr=0.6)       # warm-up
            t0 = time.perf_counter()
            for it in range(n_steps):
                params, *_ = orc.train_step(params, slots, x, y, it + 1, cfg, lr=0.6)
            return (time.perf_counter() - t0) / n_steps

    cands = sorted({min(threads, c) for c in (4, 8, 16, 32, threads)})
    best = min(cands, key=lambda c: run(c, 1))
    per_step = run(best, steps)
    return sample_rows / per_step, per_step, best


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.2] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def run_reference(args, rank, world):
    """Reference arm: CPU path of the same step on the host cores, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_rows
    v, s_per_step, used = cpu_port_seq_per_s(sample, max(1, args.steps), cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'sequences/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': s_per_step * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1]: B=4096,T=48,F=32,O=16,L=1,H=256 train step '
                               '(bounded sample of %d windows per step)' % sample},
        'cpu_baseline': {'value': v, 'unit': 'sequences/s', 'cores': used, 'kind': 'port',
                         'sample': '%d of 4096 windows per step, %d steps; NumPy fp32 restatement (oracle/) -- '
                                   'TensorFlow (the reference runtime) is not installable in this image; %d host cores, best BLAS thread '
                                   'count used' % (sample, args.steps, cores)},
        'e2e': {'value': v, 'unit': 'sequences/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('LFMQ_BENCH_PRECISION', 'bf16'), choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=4096, help='windows per GPU')
    ap.add_argument('--cpu-rows', type=int, default=256, help='windows per step of the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from lfm_quant_b200.engine import ForecasterEngine

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    eng = ForecasterEngine(max_batch=B, seq_len=T, n_inputs=F, n_outputs=O, num_hidden=H, num_layers=L,
                           target_idx=TARGET_IDX, train=True, precision=args.precision, optimizer='Adadelta',
                           target_lambda=0.5, rnn_lambda=0.7, max_grad_norm=50.0, max_norm=3.0, seed=SEED)
    eng.set_weights(initial_weights())
    lr = 0.6
    rng = np.random.default_rng(SEED + 1000 * rank)
    NB = 4                                              # rotating resident batches
    host = [synthetic(B, rng) for _ in range(NB)]
    pinned = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()) for x, y in host]
    resident = [(px.to(dev), py.to(dev)) for px, py in pinned]
    denoms = None
    if world > 1:                                       # global loss denominators, once per batch
        denoms = []
        for _, y in resident:
            d = eng.mask_count(y)
            dist.all_reduce(d)
            denoms.append(d)

    def step_resident(i):
        x, y = resident[i % NB]
        if world > 1:
            return eng.train_step_dp(x, y, i, lr, rank * B, denoms[i % NB])
        return eng.train_step(x, y, i, lr)

    # e2e: the public host-batch API (lfm_quant_b200.engine.HostBatchPipeline): every step copies its own inputs
    # from pinned host memory (H2D on a copy stream, overlapped with the previous step's compute) and reads its
    # {loss, mse_0} back to the host.
    from lfm_quant_b200.engine import HostBatchPipeline
    if world > 1:
        step_fn = lambda x, y, i, lr_: eng.train_step_dp(x, y, i, lr_, rank * B, denoms[i % NB])
    else:
        step_fn = None
    pipe = HostBatchPipeline(eng, B, step_fn=step_fn)

    def step_e2e(i):
        return pipe.step(pinned[i % NB][0], pinned[i % NB][1], i, lr, next_batch=pinned[(i + 1) % NB])

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, first):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(first + i)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    it = 0
    for _ in range(args.warmup):
        step_resident(it)
        it += 1
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    eng.profile(True)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    ms = timed(step_resident, args.steps, it)
    t1 = time.perf_counter()
    launches = eng.launch_count - l0
    regions = eng.profile_read()
    eng.profile(False)
    it += args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for _ in range(2):
        step_e2e(it)
        it += 1
    ms_e2e = timed(step_e2e, args.steps, it)
    pipe.finish()
    it += args.steps
    final = step_resident(it + 1).float().cpu().numpy()

    if rank == 0:
        value = world * B * args.steps / (ms * 1e-3)
        e2e = world * B * args.steps / (ms_e2e * 1e-3)
        peaks = {}
        pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.isfile(pk):
            peaks = json.load(open(pk))
        peak_tf = peaks.get('bf16_tflops_sustained') or 1400.0
        peak_bw = peaks.get('hbm_gbs') or 6650.0
        peak_src = 'MEASURED_PEAKS.json (bf16_tflops_sustained for kernels timed inside a long step; hbm_gbs)' if peaks \
            else 'fallback 1.4 PFLOP/s sustained / 6.65 TB/s (B200_PROFILING.md)'
        # Gate GEMMs = forward recurrence + backward recurrence + weight-gradient GEMM.  Algorithmic FLOPs per launch =
        # per-window figure x the windows one launch processes; durations = CUDA events recorded by the
        # library around each kernel inside the timed region (lfmq_profile_*), on the stream the kernels run on.
        k = max(args.steps, 1)
        fwd_ms, bwd_ms, wg_ms, head_ms = (regions[n][0] / k for n in ('fwd', 'bwd', 'wgrad', 'head'))
        flop_fwd = FLOP_FWD_PER_SEQ * B
        flop_bwd = 2.0 * T * 4 * H * H * B
        flop_wg = FLOP_FWD_PER_SEQ * B

        def tf(flop, ms_):
            return flop / (ms_ * 1e-3) / 1e12 if ms_ > 0 else None

        gate_ms = fwd_ms + bwd_ms + wg_ms
        achieved = tf(flop_fwd + flop_bwd + flop_wg, gate_ms)
        # algorithmic HBM bytes of the backward recurrence (the longest kernel): saved gates + 2x cell states + dpred
        # tile read, dz written (bf16)
        # (bf16, no dropout: dLoss/dh is not materialised -- the head leaves 64-byte bf16 dpred rows and the backward
        # kernel expands them on its tensor cores)
        bwd_bytes = B * T * (4 * H * 2 + 2 * H * 2 + 64) + B * T * 4 * H * 2
        head_bytes = B * T * (H * 2 + O * 4) + B * T * 64
        # DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) comes ONLY from a committed
        # `ncu --set full` capture of these kernels: profiles/r01_ncu_traffic.json is written by tools/ncu_traffic.py from
        # the .ncu-rep and names the capture it was parsed from.  No file (or a kernel missing from it) -> null.
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')
        gate_kernels = ('lstm_fwd_tc_kernel', 'lstm_bwd_tc_kernel', 'wgrad_tc_kernel')
        if args.precision == 'bf16' and os.path.isfile(tp):
            tj = json.load(open(tp))
            per = tj.get('dram_bytes_per_launch', {})
            if all(kn in per for kn in gate_kernels):
                traffic = float(sum(per[kn] for kn in gate_kernels))
                traffic_src = tj.get('source')
        roofline = {
            'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
            'frac': (achieved / peak_tf) if achieved else None,
            'traffic': traffic, 'traffic_source': traffic_src,
            'peak_source': peak_src,
            'kernel': 'gate GEMMs: lstm_fwd_tc_kernel + lstm_bwd_tc_kernel + wgrad_tc_kernel (one launch each per step)',
            'per_kernel': {
                'lstm_fwd_tc_kernel': {'ms': fwd_ms, 'tflops': tf(flop_fwd, fwd_ms), 'frac': (tf(flop_fwd, fwd_ms) or 0) / peak_tf},
                'lstm_bwd_tc_kernel': {'ms': bwd_ms, 'tflops': tf(flop_bwd, bwd_ms), 'frac': (tf(flop_bwd, bwd_ms) or 0) / peak_tf,
                                       'hbm_gbs': bwd_bytes / (bwd_ms * 1e-3) / 1e9 if bwd_ms > 0 else None,
                                       'hbm_frac': bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peak_bw if bwd_ms > 0 else None},
                'wgrad_tc_kernel': {'ms': wg_ms, 'tflops': tf(flop_wg, wg_ms), 'frac': (tf(flop_wg, wg_ms) or 0) / peak_tf},
                'head (fused pointwise tail, HBM-bound)': {'ms': head_ms,
                                                           'hbm_gbs': head_bytes / (head_ms * 1e-3) / 1e9 if head_ms > 0 else None,
                                                           'hbm_frac': head_bytes / (head_ms * 1e-3) / 1e9 / peak_bw if head_ms > 0 else None},
            },
            'regions_ms_per_step': {n: v[0] / k for n, v in regions.items()},
        }
        line = {
            'metric': METRIC, 'value': value, 'unit': 'sequences/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': 'BASELINE configs[%d]: %s' % (1 if world == 1 else 3,
                       'B=%d per GPU (global %d), T=48, F=32, O=16, 1-layer LSTM H=256, Adadelta lr 0.6, clip 50, '
                       'MaxNorm 3; full train step' % (B, B * world)),
                       'parallelism': 'dp%d' % world,
                       'l2': 'per-step working set (saved activations, >0.5 GB) exceeds the 126 MB L2; %d rotating '
                             'input batches' % NB,
                       'precision': args.precision},
            'e2e': {'value': e2e, 'unit': 'sequences/s', 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': int(world * B * T * (F + O) * 4), 'd2h_bytes_per_step': 8 * world,
                    'h2d_bytes_per_step_per_gpu': int(B * T * (F + O) * 4)},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'clocks': clocks,
            'final_loss_mse': [float(final[0]), float(final[1])],
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, s_step, used = cpu_port_seq_per_s(args.cpu_rows, 3, cores)
            line['cpu_baseline'] = {'value': v, 'unit': 'sequences/s', 'cores': used, 'kind': 'port',
                                    'sample': '%d of %d windows per step, 3 steps (%.1f s); NumPy fp32 restatement of '
                                              'the reference step (TensorFlow unavailable)' %
                                              (args.cpu_rows, B, 4 * s_step)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
