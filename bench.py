#!/usr/bin/env python
"""Benchmark of the per-driving-frame talking-head path (BASELINE.json metric: 256x256 frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames T] [--batch B]
                    [--clip-frames F] [--sources S] [--device cpu|cuda] [--tf32]

One *step* = animate one clip: per-clip constants (source key-points, initial driving key-points, source encoder
features) + T driving frames through KP detector -> dense motion -> multi-scale compensation/warp -> decoder -> uint8,
in micro-batches of B frames.  Default workload = BASELINE.json configs[1] (256x256 source, 64 synthetic driving
frames, full generator forward on one B200).  With N>1 (torchrun, one rank per GPU) every rank animates its own
T-frame block of an N*T-frame clip (weak scaling) and ONE NCCL all-gather of the uint8 block per step
reassembles the clip on every rank.  `--clip-frames 1024` instead fixes the TOTAL clip length (configs[2]: a 1024-frame
clip frame-sharded over the ranks); `--sources 16` runs configs[4] (S identities x T shared driving frames, identities
partitioned over the ranks).

Printed JSON line (rank 0): `value` = frames/s with inputs resident in HBM (CUDA events, L2 flushed before every step, max
over ranks); `e2e` = the same clip through the public API (`make_animation` / `make_animation_sharded` / `make_animation_multi`)
from HOST uint8 frames - what a video reader hands demo.py:166-185 - with the page-locked H2D of every driving frame, the
all-gather and the D2H of the finished uint8 clip inside the timed region; `roofline` = the dominant kernel family
(implicit-GEMM convolution) timed per launch with CUDA events on the launching stream in a separate, collective-free pass on
rank 0 (algorithmic fp32 flops / time against the measured bf16 dense peak; the kernels spend three fp16 tensor-core MACs per
fp32 MAC, so the ceiling of this fraction is 1/3); `cpu_baseline` = the CPU oracle port (oracle/sma_oracle.py, the reference's
algorithm in plain fp32 torch ops) on the host cores over a bounded sample of the same clip (N=1 only).

`--impl reference` times the reference's own algorithm on the host CPU (the oracle port: the Python reference cannot travel
to the GPU box) with all host threads, same metric/config, honouring --steps/--warmup on a bounded sample per step.  With
`--device cuda` the same eager fp32 torch code runs on the GPU (TF32 off; `--tf32` = PyTorch's default conv TF32): the
"reference algorithm on the same GPU" number of SURVEY.md 8d.
"""
import argparse
import datetime
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
DTYPE = 'f32 (fp32 storage; contractions as error-compensated fp16x3 splits on tcgen05 kind::f16, fp32 accumulate in TMEM; dense-motion (S1) and motion-codebook (S3m) convs single-pass, key-point detector and generator three-pass)'


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


def workload_text(args, world):
    if args.size == 512:
        return (f'512x512 variant (token grid 64x64; no reference behaviour at 512: parity unpinned by the reference), {args.frames} synthetic driving '
                f'frames per GPU on {world} GPU(s) (BASELINE configs[3])')
    if args.sources:
        return (f'{args.sources} source identities x {args.frames} shared synthetic driving frames, cross-reenactment batch (BASELINE configs[4]), '
                f'identities partitioned over {world} GPU(s)')
    if args.clip_frames:
        return f'256x256 source, {args.clip_frames}-frame synthetic driving clip frame-sharded over {world} GPU(s), NCCL all-gather of the uint8 clip (BASELINE configs[2])'
    return (f'256x256 source + {args.frames} synthetic driving frames per GPU, full generator forward '
            f'(BASELINE configs[1]{" x N, frame-sharded, NCCL all-gather of the uint8 clip" if world > 1 else ""})')


# -----------------------------------------------------------------------------------------------------------------------
# reference arm: the oracle port (the reference's algorithm, plain fp32 torch) on the host cores - or, on request, on the GPU
# -----------------------------------------------------------------------------------------------------------------------
def oracle_clip(O, P_g, P_me, src, drv, device):
    """per-clip constants + len(drv) driving frames, batch 1 (as the reference's loop); returns seconds (synchronised)."""
    import torch
    if device != 'cpu':
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    O.make_animation(P_g, P_me, src, drv, True, True)
    if device != 'cpu':
        torch.cuda.synchronize()
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import sma_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    dev = args.device
    K, Wm = max(1, args.steps), max(0, args.warmup)
    P_g, P_me = O.synthetic_state_dict(CFG_KEYS['net_g'], 0), O.synthetic_state_dict(CFG_KEYS['motion_estimator'], 1)
    if dev == 'cuda':
        torch.backends.cudnn.allow_tf32 = bool(args.tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True            # the reference sets it (basicsr/animate.py:53)
        P_g = {k: v.cuda() for k, v in P_g.items()}; P_me = {k: v.cuda() for k, v in P_me.items()}
        sample = args.frames
    else:
        # bounded sample: ~0.45 s per frame on 16 cores; keep the whole --steps/--warmup run within a few minutes
        sample = max(2, min(args.frames, args.ref_frames, int(150.0 / (0.45 * (K + Wm + 1)))))
    src, drv = O.synthetic_frames(sample, seed=1234)
    if dev == 'cuda':
        O_dev = torch.device('cuda')
        src, drv = src.to(O_dev), [f.to(O_dev) for f in drv]
        if hasattr(O, 'set_device'):
            O.set_device(O_dev)
    for _ in range(Wm):
        oracle_clip(O, P_g, P_me, src, drv[:max(1, min(2, sample))], dev)
    times = [oracle_clip(O, P_g, P_me, src, drv, dev) for _ in range(K)]
    dt = sum(times) / len(times)
    fps = sample / dt
    what = ('eager fp32 torch on cuda:0, cudnn.allow_tf32=%s, matmul.allow_tf32=False, batch 1' % bool(args.tf32)) if dev == 'cuda' else \
           f'torch fp32 on the host CPU, {threads} threads, batch 1'
    line = {'impl': 'reference', 'metric': '256x256 frames/sec', 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': K, 'warmup': Wm, 'ms_per_step': 1e3 * dt, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_text(args, 1), 'frames_per_step': sample, 'micro_batch': 1,
                       'runs': f'oracle/sma_oracle.py make_animation (the reference algorithm restated; pinned to the live reference by oracle/make_golden.py; '
                               f'baseline/_ref cannot travel to this box without its 7 absent third-party imports): {what}',
                       'sample': f'each step animates the first {sample} driving frames of the clip (per-clip constants included)'},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads if dev == 'cpu' else 0, 'kind': 'port',
                             'sample': f'{sample} driving frames of the clip per step, {what}'},
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
    ap.add_argument('--clip-frames', type=int, default=0, help='TOTAL driving frames of the clip, sharded over the ranks (configs[2]: 1024); overrides --frames')
    ap.add_argument('--sources', type=int, default=0, help='configs[4]: this many source identities share the --frames driving frames; identities are partitioned over the ranks')
    ap.add_argument('--batch', type=int, default=0, help='driving frames per micro-batch (default 64; 32 at --size 512)')
    ap.add_argument('--size', type=int, default=256, choices=[256, 512], help='image size: 512 = the 512x512 variant of BASELINE configs[3] (use with --frames 256)')
    ap.add_argument('--ref-frames', type=int, default=12, help='frames per step of the CPU reference / cpu_baseline sample')
    ap.add_argument('--device', default='cpu', choices=['cpu', 'cuda'], help='--impl reference: where the oracle port runs')
    ap.add_argument('--tf32', action='store_true', help='--impl reference --device cuda: leave cudnn.allow_tf32 at the PyTorch default (True)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--detail', action='store_true', help='print the per-shape time table of the roofline pass to stderr')
    ap.add_argument('--ncu', action='store_true', help='profiling aid: one warm-up step and one step, nothing else (run under ncu)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
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
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    # host-side frame gathering of the e2e arm: torchrun pins OMP_NUM_THREADS to 1, give every rank its share of the cores
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    if args.clip_frames:
        if args.clip_frames % world:
            raise SystemExit('--clip-frames must be a multiple of the number of ranks')
        args.frames = args.clip_frames // world
    global H, W
    H = W = args.size
    if not args.batch:
        args.batch = 64 if args.size == 256 else 32
    T, Bm, K, Wm = args.frames, args.batch, args.steps, max(3, args.warmup)
    nsrc_total = args.sources
    if nsrc_total and nsrc_total % world:
        raise SystemExit('--sources must be a multiple of the number of ranks')
    nsrc = nsrc_total // world if nsrc_total else 1

    CFG = net_cfg()
    CFG['network_g']['img_size'] = args.size
    g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
    g.load_state_dict(O.synthetic_state_dict(O.variant_shapes(CFG_KEYS['net_g'], args.size), 0)); me.load_state_dict(O.synthetic_state_dict(CFG_KEYS['motion_estimator'], 1))
    g, me = g.eval().to(dev), me.eval().to(dev)
    if nsrc_total:
        # configs[4]: every rank animates its own identities over the SAME T driving frames
        src0, drv_all = O.synthetic_frames(T, seed=1234, size=args.size)
        my_ids = S.dist.shard_sources(nsrc_total, rank, world)
        srcs = [src0 if i == 0 else O.synthetic_frames(0, seed=1234 + i, size=args.size)[0] for i in my_ids]
        drv, first = drv_all, drv_all[0]
        n_clip = T
    else:
        # rank r owns frames [r*T, (r+1)*T) of the N*T-frame clip (same source on every rank)
        src0, drv_all = O.synthetic_frames(world * T, seed=1234, size=args.size)
        srcs = [src0]
        drv, first = drv_all[rank * T:(rank + 1) * T], drv_all[0]
        n_clip = world * T
    srcs_d = [s.unsqueeze(0).to(dev) for s in srcs]
    first_d = first.unsqueeze(0).to(dev)
    drv_d = torch.stack(drv).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    clip_u8 = torch.empty((nsrc * T, H, W, 3), dtype=torch.uint8, device=dev)
    gathered = torch.empty((world * nsrc * T, H, W, 3), dtype=torch.uint8, device=dev) if world > 1 else None

    def clear_caches():
        g.clear_source_cache(); me.dense_motion_network.clear_source_cache()           # per-clip work is redone every step

    def step_device(gather=True):
        """Inputs resident in HBM.  `gather=False`: the collective-free variant used by the rank-0-only roofline pass."""
        clear_caches()
        for si, s_d in enumerate(srcs_d):
            # the clip's first driving frame fixes kp_driving_initial: on rank 0 of a single-identity clip it is frame 0 of the first micro-batch
            anim = S.ClipAnimator(g, me, s_d, first_d if (rank > 0 or nsrc_total) else None, True, True, 1.0)
            for i0 in range(0, T, Bm):
                clip_u8[si * T + i0:si * T + i0 + Bm] = anim.step(drv_d[i0:i0 + Bm])
        if world > 1 and gather:
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
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
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
    fps = world * nsrc * T * K / (ms * 1e-3)

    # ---- end to end through the public API: HOST uint8 frames in, uint8 clip on the host out ------------------------------
    to_u8 = lambda f: np.ascontiguousarray(O.to_uint8(f))          # what a video reader yields (demo.py:166-178): HWC uint8 RGB
    srcs_h = [to_u8(s) for s in srcs]
    drv_h_all = [to_u8(f) for f in drv_all]
    e2e_check = {}

    def step_e2e():
        clear_caches()
        if nsrc_total:
            preds, _ = S.make_animation_multi(srcs_h, drv_h_all, g, me, relative=True, adapt_movement_scale=True, batch=Bm)
            out = preds
            if world > 1:          # reassemble the identities of all ranks (device all-gather of this rank's frames)
                loc = torch.from_numpy(np.stack([np.stack(p) for p in preds])).to(dev)
                dist.all_gather_into_tensor(gathered, loc.view(-1, H, W, 3))
            return out
        if world > 1:
            # every rank renders its block and ONE all-gather reassembles the clip on every rank; rank 0 reads the whole clip back
            clip = S.make_animation_sharded(srcs_h[0], drv_h_all, g, me, relative=True, adapt_movement_scale=True, batch=Bm)
            if rank == 0:
                host = S.animate._io(dev).result(tuple(clip.shape), torch.uint8)
                host.copy_(clip, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                e2e_check['n'] = int(clip.shape[0])
                return S.animate._io(dev).hand_out(host)
            torch.cuda.current_stream().synchronize()
            return None
        preds, _ = S.make_animation(srcs_h[0], drv_h_all, g, me, relative=True, adapt_movement_scale=True, batch=Bm)
        return preds

    out = step_e2e()
    del out
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        out = step_e2e()
        del out
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_fps = world * nsrc * T * K / float(e2e_s.item())
    # bytes per step on this rank (counted from the tensors copied): uint8 source(s) + first frame + this rank's driving frames in;
    # the finished uint8 frames out (rank 0 of a sharded clip reads the whole gathered clip)
    h2d = (nsrc + (1 if (world > 1 and not nsrc_total) else 0) + T) * 3 * H * W
    d2h = (world * T if (world > 1 and not nsrc_total) else nsrc * T) * H * W * 3

    # ---- roofline of the dominant kernel (rank 0 alone: a collective-free step, per-launch CUDA events) --------------------
    roof, stage_table = None, None
    if rank == 0 and not args.no_roofline:
        S.ops.PROFILE = []
        step_device(gather=False)
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
            for label, (n, fl, t) in sorted(detail.items(), key=lambda kv: -kv[1][2])[:60]:
                print(f'{t:8.2f} ms  {n:4d}x  {fl / max(t, 1e-9) / 1e9:7.1f} TF  {label}', file=sys.stderr)
        pk = peaks()
        stage_table = {k: {'launches': v[0], 'gflop': v[1] / 1e9, 'mbytes': v[2] / 1e6, 'ms': v[3],
                           'tflops': v[1] / (v[3] * 1e-3) / 1e12 if v[3] else None,
                           'gbs': v[2] / (v[3] * 1e-3) / 1e9 if v[3] else None} for k, v in agg.items()}
        c = agg.get('conv')
        if c:
            ach = c[1] / (c[3] * 1e-3) / 1e12
            traffic = None                      # dram bytes (read + write) per conv launch from the committed ncu pass of this command
            for name in ('r2_conv_traffic.json', 'r1_conv_traffic.json'):
                tp = os.path.join(ROOT, 'profiles', name)
                if os.path.exists(tp):
                    traffic = json.load(open(tp)).get('dram_bytes_per_conv_launch')
                    break
            roof = {'kernel': 'implicit-GEMM conv (sma_conv2d_fwd: conv_tc2_kernel and friends)', 'bound': 'tensor', 'achieved': ach,
                    'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s', 'frac': ach / pk['tflops_sustained'], 'traffic': traffic,
                    'peak_source': pk['source'] + ' bf16 dense, sustained',
                    'launches_per_step': c[0], 'avg_launch_ms': c[3] / c[0], 'gflop_per_launch': c[1] / c[0] / 1e9,
                    'share_of_step_ms': c[3] / (ms / K),
                    'note': 'algorithmic flops = 2*M*N*K of the fp32 convolution (zero-padded input channels included); the kernels spend 3 fp16 '
                            'tensor-core MACs per fp32 MAC (fp16 hi/lo split, fp32 accumulate) to stay within 1e-3 of the fp32 reference, so '
                            'frac <= 1/3 by construction; since round 2b these launches also produce the GroupNorm statistics of their outputs in the epilogue '
                            '(the standalone statistics passes, 2.9 ms per step, are gone; the conv launches themselves got 1.7 ms longer) and stage '
                            'their input tiles with TMA tensor maps; measured on rank 0 in a separate collective-free step'}
        w_ = agg.get('warp')
        if w_:
            stage_table['warp']['hbm_frac'] = w_[2] / (w_[3] * 1e-3) / 1e9 / pk['hbm_gbs']

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        n = max(2, min(T, args.ref_frames if args.size == 256 else 3))
        P_g, P_me = O.synthetic_state_dict(O.variant_shapes(CFG_KEYS['net_g'], args.size), 0), O.synthetic_state_dict(CFG_KEYS['motion_estimator'], 1)
        s_, d_ = O.synthetic_frames(n + 1, seed=1234, size=args.size)
        oracle_clip(O, P_g, P_me, s_, d_[:1], 'cpu')            # warm-up frame
        dt = oracle_clip(O, P_g, P_me, s_, d_[1:], 'cpu')
        cpu = {'value': n / dt, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
               'sample': f'{n} driving frames of the same clip (+1 warm-up), oracle/sma_oracle.py make_animation, torch fp32, '
                         f'{threads} threads, {dt:.1f} s'}

    if rank == 0:
        line = {'metric': f'{args.size}x{args.size} frames/sec', 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
                'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong' if args.clip_frames or nsrc_total else 'weak', 'vs_baseline': None,
                'dtype': DTYPE, 'data': 'synthetic',
                'config': {'workload': workload_text(args, world),
                           'frames_per_step_per_gpu': nsrc * T, 'clip_frames_total': n_clip, 'sources_per_gpu': nsrc, 'micro_batch': Bm,
                           'l2': 'flushed (256 MiB write) before every timed step',
                           'weights': 'seeded synthetic (no pretrained checkpoint offline)', 'parallelism': f'frames-dp{world}',
                           'e2e_input': 'host uint8 HWC frames (what demo.py:166-178 reads from the video), converted on the device'},
                'e2e': {'value': e2e_fps, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'stages': stage_table}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
