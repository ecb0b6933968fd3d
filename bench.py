#!/usr/bin/env python
"""Benchmark of the per-driving-frame talking-head path (BASELINE.json metric: 256x256 frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames T] [--batch B]

One *step* = animate one clip: per-clip constants (source key-points, initial driving key-points, source encoder
features) + T driving frames through KP detector -> dense motion -> multi-scale compensation/warp -> decoder -> uint8,
in micro-batches of B frames.  Default workload = BASELINE.json configs[1] (256x256 source, 64 synthetic driving
frames, full generator forward on one B200).  With N>1 (torchrun, one rank per GPU) every rank animates its own
T-frame block of an N*T-frame clip (weak scaling) and one NCCL all-gather of the uint8 block per step
reassembles the clip on every rank.

Printed JSON line (rank 0): `value` = frames/s with inputs resident in HBM; `e2e` = the same clip through the public
`make_animation` call with HOST tensors (pinned H2D of every driving frame, D2H of the uint8 frames inside the timed
region); `roofline` = the dominant kernel (implicit-GEMM convolution) timed per launch with CUDA events on the
launching stream in a separate pass (algorithmic fp32 flops / time against the measured bf16 dense peak; the kernels spend three
fp16 tensor-core MACs per fp32 MAC, so the ceiling of this fraction is 1/3); `cpu_baseline` = the CPU oracle port (oracle/sma_oracle.py, the reference's
algorithm in plain fp32 torch ops) on the host cores over a bounded sample of the same clip.

`--impl reference` times the reference's own algorithm on the host CPU (the oracle port: the Python reference cannot
travel to the GPU box) with all host threads, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 256
CFG_KEYS = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'state_keys.json')))


def net_cfg():
    import yaml
    return yaml.safe_load(open(os.path.join(ROOT, 'options', 'test.yml')))


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops_burst': d['bf16_tflops'], 'tflops_sustained': d['bf16_tflops_sustained'],
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(self.rows), 'reasons': reasons}


def cpu_oracle_fps(n_frames, threads):
    """The oracle port on the host cores: per-clip constants + n_frames driving frames, batch 1 (as the reference's loop)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import sma_oracle as O
    torch.set_num_threads(threads)
    P_g, P_me = O.synthetic_state_dict(CFG_KEYS['net_g'], 0), O.synthetic_state_dict(CFG_KEYS['motion_estimator'], 1)
    src, drv = O.synthetic_frames(n_frames + 1, seed=1234)
    O.make_animation(P_g, P_me, src, drv[:1], True, True)            # warm-up frame
    t0 = time.perf_counter()
    O.make_animation(P_g, P_me, src, drv[1:], True, True)
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(2, min(args.frames, args.ref_frames))
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        fps, dt = cpu_oracle_fps(sample, threads)
        vals.append((fps, dt))
    fps = sum(v[0] for v in vals) / len(vals)
    line = {'impl': 'reference', 'metric': '256x256 frames/sec', 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': len(vals), 'warmup': 1, 'ms_per_step': 1e3 * sum(v[1] for v in vals) / len(vals), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'256x256 source + {args.frames} synthetic driving frames, full generator forward (configs[1])',
                       'frames_per_step': sample, 'micro_batch': 1},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{sample} driving frames of the clip per step (+1 warm-up frame), oracle/sma_oracle.py '
                                       f'make_animation, torch fp32, {threads} threads'},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--frames', type=int, default=64, help='driving frames per GPU per step (configs[1]: 64)')
    ap.add_argument('--batch', type=int, default=64, help='driving frames per micro-batch')
    ap.add_argument('--ref-frames', type=int, default=12, help='frames per step of the CPU reference / cpu_baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--detail', action='store_true', help='print the per-shape time table of the roofline pass to stderr')
    ap.add_argument('--ncu', action='store_true', help='profiling aid: one warm-up step and one step, nothing else (run under ncu)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import sma_b200 as S
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import sma_oracle as O          # synthetic weights / frames + the cpu_baseline leg only; never on the product path

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    # host-side frame gathering of the e2e arm (torch.stack into pinned memory): torchrun pins OMP_NUM_THREADS to 1, give every rank its share of the cores
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    T, Bm, K, Wm = args.frames, args.batch, args.steps, max(3, args.warmup)

    CFG = net_cfg()
    g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
    g.load_state_dict(O.synthetic_state_dict(CFG_KEYS['net_g'], 0)); me.load_state_dict(O.synthetic_state_dict(CFG_KEYS['motion_estimator'], 1))
    g, me = g.eval().to(dev), me.eval().to(dev)
    # rank r owns frames [r*T, (r+1)*T) of the N*T-frame clip (same source on every rank)
    src, drv_all = O.synthetic_frames(world * T, seed=1234)
    drv = drv_all[rank * T:(rank + 1) * T]
    first = drv_all[0]
    src_d = src.unsqueeze(0).to(dev)
    first_d = first.unsqueeze(0).to(dev)
    drv_d = torch.stack(drv).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    clip_u8 = torch.empty((T, H, W, 3), dtype=torch.uint8, device=dev)
    gathered = torch.empty((world * T, H, W, 3), dtype=torch.uint8, device=dev) if world > 1 else None

    def step_device():
        g._src_cache = None; me.dense_motion_network._src_cache = None           # per-clip work is redone every step
        anim = S.ClipAnimator(g, me, src_d, first_d, True, True, 1.0)
        for i0 in range(0, T, Bm):
            clip_u8[i0:i0 + Bm] = anim.step(drv_d[i0:i0 + Bm])
        if world > 1:
            dist.all_gather_into_tensor(gathered, clip_u8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        """k steps, L2 flushed before each, CUDA events on the launching stream; returns total ms (max over ranks)."""
        evs = []
        barrier()
        for _ in range(k):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if args.detail:           # is the step bound by the host's launch rate?  enqueue time (no sync) vs device time
        step_device(); torch.cuda.synchronize()
        t0 = time.perf_counter(); step_device(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f'host enqueue {1e3 * (t1 - t0):.1f} ms, until device idle {1e3 * (t2 - t0):.1f} ms', file=sys.stderr)
    if args.ncu:
        step_device(); torch.cuda.synchronize()
        torch.cuda.profiler.start()          # ncu --profile-from-start off: exactly one step is profiled
        step_device(); torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for _ in range(Wm):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = S.ops.launch_count()
    ms = timed(step_device, K)
    launches = S.ops.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    fps = world * T * K / (ms * 1e-3)

    # ---- end to end through the public API, host tensors -------------------------------------------------
    src_h, drv_h = src, list(drv)

    def step_e2e():
        g._src_cache = None; me.dense_motion_network._src_cache = None
        preds, _ = S.make_animation(src_h, drv_h, g, me, relative=True, adapt_movement_scale=True, batch=Bm)
        if world > 1:
            dist.all_gather_into_tensor(gathered, clip_u8)     # same collective as the device arm
        return preds

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_fps = world * T * K / float(e2e_s.item())
    h2d = (T + 2) * 3 * H * W * 4
    d2h = 2 * T * H * W * 3

    # ---- roofline of the dominant kernel (rank 0, N=1 accounting; separate pass, per-launch CUDA events) ----
    roof, stage_table = None, None
    if rank == 0 and not args.no_roofline:
        S.ops.PROFILE = []
        step_device()
        torch.cuda.synchronize()
        prof, S.ops.PROFILE = S.ops.PROFILE, None
        agg = {}
        detail = {}
        for kind, fl, nb, e0, e1, label in prof:
            t = e0.elapsed_time(e1)
            a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += fl; a[2] += nb; a[3] += t
            dd = detail.setdefault(label, [0, 0.0, 0.0])
            dd[0] += 1; dd[1] += fl; dd[2] += t
        if args.detail:
            for label, (n, fl, t) in sorted(detail.items(), key=lambda kv: -kv[1][2])[:45]:
                print(f'{t:8.2f} ms  {n:4d}x  {fl / max(t, 1e-9) / 1e9:7.1f} TF  {label}', file=sys.stderr)
        pk = peaks()
        stage_table = {k: {'launches': v[0], 'gflop': v[1] / 1e9, 'mbytes': v[2] / 1e6, 'ms': v[3],
                           'tflops': v[1] / (v[3] * 1e-3) / 1e12 if v[3] else None,
                           'gbs': v[2] / (v[3] * 1e-3) / 1e9 if v[3] else None} for k, v in agg.items()}
        c = agg.get('conv')
        if c:
            ach = c[1] / (c[3] * 1e-3) / 1e12
            traffic = None                      # dram bytes (read + write) per conv launch from the committed ncu pass of this command
            tp = os.path.join(ROOT, 'profiles', 'r1_conv_traffic.json')
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get('dram_bytes_per_conv_launch')
            roof = {'kernel': 'implicit-GEMM conv (sma_conv2d_fwd: conv_tc2_kernel and friends)', 'bound': 'tensor', 'achieved': ach,
                    'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s', 'frac': ach / pk['tflops_sustained'], 'traffic': traffic,
                    'peak_source': pk['source'] + ' bf16 dense, sustained',
                    'launches_per_step': c[0], 'avg_launch_ms': c[3] / c[0], 'gflop_per_launch': c[1] / c[0] / 1e9,
                    'share_of_step_ms': c[3] / (ms / K),
                    'note': 'algorithmic flops = 2*M*N*K of the fp32 convolution; the kernels spend 3 fp16 tensor-core MACs per fp32 MAC '
                            '(fp16 hi/lo split, fp32 accumulate) to stay within 1e-3 of the fp32 reference, so frac <= 1/3 by construction'}
        w_ = agg.get('warp')
        if w_:
            stage_table['warp']['hbm_frac'] = w_[2] / (w_[3] * 1e-3) / 1e9 / pk['hbm_gbs']

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = max(2, min(T, args.ref_frames))
        v, dt = cpu_oracle_fps(n, threads)
        cpu = {'value': v, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n} driving frames of the same clip (+1 warm-up), oracle/sma_oracle.py make_animation, torch fp32, '
                         f'{threads} threads, {dt:.1f} s'}

    if rank == 0:
        line = {'metric': '256x256 frames/sec', 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
                'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic',
                'config': {'workload': f'256x256 source + {T} synthetic driving frames per GPU, full generator forward '
                                       f'(BASELINE configs[1]{" x N, frame-sharded, NCCL all-gather of the uint8 clip" if world > 1 else ""})',
                           'frames_per_step_per_gpu': T, 'micro_batch': Bm, 'l2': 'flushed (256 MiB write) before every timed step',
                           'weights': 'seeded synthetic (no pretrained checkpoint offline)', 'parallelism': f'frames-dp{world}'},
                'e2e': {'value': e2e_fps, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'stages': stage_table}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
