"""Benchmark of the B200-native multi_part_assembly hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): shapes/sec of pn_transformer forward + loss on synthetic
[B=32, P=20 valid parts, N=1000 points] Breaking-Bad-like shapes
(configs/pn_transformer ... everyday), bf16 autocast for the encoder /
attention GEMMs, Chamfer and SE(3) always fp32 -- BASELINE config C.

One process per GPU (torchrun for N > 1): every rank steps its own batch of 32
shapes (weak scaling, no data-path collective -- fwd+loss has none, SURVEY.md
8e); the timed region is bracketed by barrier + synchronize, timed with CUDA
events, max over ranks.

Keys beyond the base contract:
  roofline     dominant kernel (Chamfer grid search of shape_cd_loss): achieved
               algorithmic GB/s from CUDA events inside the library
  cpu_baseline the CPU oracle (port of the reference path) on the host cores
  e2e          same metric through BaseModel.forward_pass with the batch in
               pinned HOST memory: H2D of the batch + D2H of the loss per step
  train_step   secondary figure: forward + loss + backward + gradient
               all-reduce over the ranks + Adam as one CUDA graph (runtime.GraphedTrainStep),
               same batch, timed like the headline; `--no-train` skips it
`--impl reference` times the CPU restatement of the reference path (the
reference has no CPU Chamfer of its own: chamfer.py:18 asserts CUDA) with all
host threads on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU, P, N_PTS = 32, 20, 1000
WORKLOAD = 'configs[2]: configs/pn_transformer everyday, B=32 x P=20 valid parts x N=1000 pts, fwd+loss'


# --------------------------------------------------------------------------
def cpu_step(B, seed, threads):
    """One forward+loss of the CPU oracle on B shapes; returns seconds."""
    import torch
    from oracle import torch_ref, cpu as ocpu
    from oracle.params import fill_params_
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    torch.set_num_threads(threads)
    ocpu.set_num_threads(threads)
    if not hasattr(cpu_step, 'sd'):
        model = fill_params_(build_model(get_cfg('pn_transformer')), 0)
        cpu_step.sd = {k: v.detach() for k, v in model.state_dict().items()}
    batch = make_batch(B, P=P, N=N_PTS, num_valid=P, seed=seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        rot, trans = torch_ref.pn_transformer_forward(batch, cpu_step.sd, training=True)
        out, _ = torch_ref.geometric_losses(batch, rot, trans, training=True)
    float(out['loss'])
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: CPU restatement of the reference path, rank 0 only."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    Bs = 4  # bounded sample per step
    for i in range(args.warmup):
        cpu_step(Bs, i, threads)
    t = [cpu_step(Bs, 100 + i, threads) for i in range(args.steps)]
    total = sum(t)
    value = Bs * args.steps / total
    line = {
        'impl': 'reference', 'metric': 'shapes_per_sec_pn_transformer_fwd_loss', 'value': value,
        'unit': 'shapes/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': f'{Bs} shapes per step (same P, N)'},
        'cpu_baseline': {'value': value, 'unit': 'shapes/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{args.steps} steps x {Bs} shapes, oracle/ C+torch port, '
                                   f'{threads} threads'},
        'e2e': {'value': value, 'unit': 'shapes/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    args.emit(line)


# --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def run_native(args):
    import torch
    import torch.distributed as dist
    from multi_part_assembly_b200 import _lib
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    from multi_part_assembly_b200.compat.lightning import Trainer
    from multi_part_assembly_b200 import profiler

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the hot path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')  # stdout carries the one JSON line only
        dist.init_process_group('nccl', device_id=dev)

    torch.manual_seed(rank)
    cfg = get_cfg('pn_transformer', 'everyday')
    model = build_model(cfg).to(dev).train()
    model.trainer = Trainer()
    for m in model.modules():  # dropout off on both arms for comparability
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, 'dropout') and isinstance(m.dropout, float):
            m.dropout = 0.0
    host = make_batch(B_PER_GPU, P=P, N=N_PTS, num_valid=P, seed=rank, pin_memory=True)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def eager_step(batch):
        with torch.no_grad():
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=args.dtype == 'bf16'):
                return model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']

    graphed = None
    if not args.no_graph:
        from multi_part_assembly_b200.runtime import GraphedStep
        graphed = GraphedStep(model, resident, mode='train',
                              autocast_dtype=torch.bfloat16 if args.dtype == 'bf16' else None)

    def step(batch):
        if graphed is not None:
            return graphed(batch)['loss']
        return eager_step(batch)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps, L2 flushed before each (events exclude the flush)."""
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            fn()
            stops[i].record()
        barrier()
        return sum(s.elapsed_time(e) for s, e in zip(starts, stops))

    for _ in range(args.warmup):
        step(None if graphed is not None else resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    eager_step(resident)
    launches_per_step = _lib.launch_count() - l0  # a graph replay issues the same kernel nodes
    torch.cuda.synchronize()
    # inputs already resident in HBM (the graph's static input buffers / `resident`)
    ms_total = timed(lambda: step(None if graphed is not None else resident), args.steps)
    launches = launches_per_step * args.steps

    # e2e: pinned host batch -> device, forward_pass, loss back to the host
    # e2e: every step copies ITS batch from pinned host memory and reads its loss back.
    # With the graph runtime the copy of batch k+1 is issued before step k runs (input
    # prefetch on a side stream, as a data loader does), so it overlaps the compute.
    host2 = make_batch(B_PER_GPU, P=P, N=N_PTS, num_valid=P, seed=rank + 1000, pin_memory=True)
    host_batches = [host, host2]

    def e2e_loop(steps):
        """K end-to-end steps; returns the summed CUDA-event time of the steps."""
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        if graphed is not None:
            flush.zero_()
            starts[0].record()
            graphed.prefetch(host_batches[0])
            for i in range(steps):
                out = graphed.run_prefetched()
                if i + 1 < steps:
                    graphed.prefetch(host_batches[(i + 1) % 2])  # next batch's H2D overlaps this step
                loss = float(out['loss'])                          # D2H of this step's result
                stops[i].record()
                if i + 1 < steps:
                    flush.zero_()
                    starts[i + 1].record()
        else:
            for i in range(steps):
                flush.zero_()
                starts[i].record()
                batch = {k: v.to(dev, non_blocking=True) for k, v in host_batches[i % 2].items()}
                loss = float(eager_step(batch))
                stops[i].record()
        barrier()
        return sum(s.elapsed_time(e) for s, e in zip(starts, stops))

    e2e_loop(max(2, args.warmup // 2))
    ms_e2e = e2e_loop(args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel: same steps with the library's per-kernel events on
    profiler.enable(True)
    for _ in range(min(args.steps, 10)):
        flush.zero_()
        eager_step(resident)  # eager: the library's event pairs cannot be recorded inside a graph
    torch.cuda.synchronize()
    prof = profiler.report()
    profiler.enable(False)

    # secondary figure: the whole training step -- forward, loss, backward,
    # gradient all-reduce over the ranks (the path's one collective), Adam -- as one CUDA graph
    train = None
    ms_train = 0.0
    if not args.no_train:
        try:
            from multi_part_assembly_b200.runtime import GraphedTrainStep
            opt = model.configure_optimizers()
            if isinstance(opt, tuple):
                opt = opt[0][0]
            gts = GraphedTrainStep(model, opt, resident,
                                   autocast_dtype=torch.bfloat16 if args.dtype == 'bf16' else None)
            k_train = max(3, min(args.steps, 50))
            for _ in range(3):
                gts()
            ms_train = timed(lambda: gts(), k_train)
            train = {'steps': k_train, 'cuda_graph': True,
                     'includes': 'forward + loss + backward + gradient all-reduce (N > 1) + Adam'}
        except Exception as e:  # keep the headline line even if the capture fails on this box
            train = {'error': repr(e)[:300]}
            torch.cuda.synchronize()

    t = torch.tensor([ms_total, ms_e2e, ms_train], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_train = t.tolist()
    if train is not None and 'error' not in train:
        train['ms_per_step'] = ms_train / train['steps']
        train['shapes_per_s'] = B_PER_GPU * world * train['steps'] / (ms_train / 1e3)
    if world > 1:
        # last collective done.  The captured graphs hold NCCL resources and tearing the
        # process group down around them can block: ranks leave without the teardown.
        torch.cuda.synchronize()
        if rank != 0:
            sys.stdout.flush()
            os._exit(0)
    if rank != 0:
        return

    shapes = B_PER_GPU * world * args.steps
    value = shapes / (ms_total / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    dom = max(prof.items(), key=lambda kv: kv[1]['ms_total'])[0] if prof else None
    roof = None
    if dom is not None:
        d = prof[dom]
        avg_ms = d['ms_total'] / d['launches']
        # algorithmic bytes of one shape-level Chamfer launch: 24 B per point of both
        # clouds (12 read + 4 dist + 8 idx as the reference writes them), SURVEY.md 8d
        alg_bytes = d.get('alg_bytes_per_launch', 24.0 * B_PER_GPU * 2 * P * N_PTS)
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        # DRAM bytes of one launch of that kernel from the committed `ncu --set full` capture
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json'))).get(dom)
            if tr is not None:
                traffic = tr['dram_bytes_read'] + tr['dram_bytes_write']
        except (OSError, ValueError, KeyError):
            pass
        roof = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': hbm_peak,
                'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': traffic,
                'avg_launch_ms': avg_ms, 'share_of_step': d['ms_total'] / sum(
                    v['ms_total'] for v in prof.values()),
                'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback 6650 GB/s',
                'note': 'the search is paced by lane divergence and L2/L1 load latency, not by HBM '
                        '(DESIGN.md 4a); HBM-bound kernels of the path: the BatchNorm passes of the '
                        'PointNet backward (35-60 % of peak, DESIGN.md 4)',
                'kernels_ms_per_step': {k: v['ms_total'] / min(args.steps, 10) for k, v in prof.items()}}

    threads = os.cpu_count() or 1
    cpu_B = 8
    cpu_step(2, 0, threads)
    cpu_t = cpu_step(cpu_B, 1, threads)
    line = {
        'metric': 'shapes_per_sec_pn_transformer_fwd_loss', 'value': value, 'unit': 'shapes/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': B_PER_GPU, 'parts': P, 'points': N_PTS,
                   'l2': 'flushed (192 MiB memset) before every timed step',
                   'mode': 'training-mode forward (BatchNorm batch statistics) + all loss terms, '
                           'no autograd recording, dropout 0',
                   'cuda_graph': graphed is not None},
        'clocks': clocks,
        'e2e': {'value': shapes / (ms_e2e / 1e3), 'unit': 'shapes/s',
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4},
        'gpu_launches': launches,
        'train_step': train,
        'roofline': roof,
        'cpu_baseline': {'value': cpu_B / cpu_t, 'unit': 'shapes/s', 'cores': threads,
                         'kind': 'port',
                         'sample': f'1 step of {cpu_B} shapes (same P, N), oracle/ C+torch port'},
    }
    args.emit(line)
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of a CUDA graph')
    ap.add_argument('--no-train', action='store_true', help='skip the secondary training-step figure')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'native' else args.warmup
    # stdout carries exactly one JSON line: while the benchmark runs, file descriptor 1 points
    # at stderr, so that banners printed by libraries (NCCL prints its version to stdout on
    # some boxes) cannot get in front of it; `emit` restores it for the final line.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        sys.stdout.write(json.dumps(line) + '\n')
        sys.stdout.flush()

    args.emit = emit
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()
