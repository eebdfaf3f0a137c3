#!/usr/bin/env python
"""bench.py -- keypoint-pairs/s of the MDGAT forward hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                      (CPU arm: the unmodified reference under torch on the host cores)

One "step" = one forward of MDGAT over one batch of synthetic pairs: encoder -> 18 GNN layers
(full / top-k attention) -> score matrix -> 100 log-Sinkhorn iterations -> match extraction +
triplet loss. Workload at every N: BASELINE.json configs[1] per GPU (batch 32, 2x512 keypoints,
33-dim descriptors, L=9, T=100); ranks process independent batches (weak scaling); the per-rank
match results are all-gathered once per step INSIDE both timed regions (one packed NCCL
all_gather_into_tensor, no host synchronisation).

Prints ONE JSON line (rank 0). `value` = pairs/s with inputs resident in HBM; `e2e` = the same
through MDGAT.forward() with pinned HOST inputs (H2D + D2H inside the timed region).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

DEFAULT_K = [128, None, 128, None, 64, None, 64, None]     # /root/reference/test.py:83
PRETRAINED = os.path.join(ROOT, 'oracle', '_ref', 'best_model_fp32.npz')   # data fixture, built by build()


def net_config(L, T):
    return {'sinkhorn_iterations': T, 'match_threshold': 0.2, 'lr': 1e-4, 'loss_method': 'triplet_loss',
            'k': list(DEFAULT_K), 'descriptor': 'FPFH', 'mutual_check': False, 'triplet_loss_gamma': 0.5,
            'train_step': 3, 'L': L}


def workload_name(B, N, L, T):
    """Names the BASELINE.json config the arguments correspond to."""
    shape = 'batch %d per GPU, 2x%d keypoints, 33-dim desc, %d MDGAT layers (L=%d), %d Sinkhorn iters' % (B, N, L, L, T)
    if (B, N, L, T) == (32, 512, 9, 100):
        return 'cfg2: ' + shape
    if (B, N, L, T) == (32, 2048, 9, 100):
        return 'cfg4: ' + shape
    if (B, N, L, T) == (1, 128, 4, 20):
        return 'cfg1: ' + shape
    return 'custom: ' + shape


def workload_config(B, N, L, T, world):
    """The `config` object of the JSON line; both arms (b200 and reference) print the same one."""
    return {'workload': workload_name(B, N, L, T), 'batch_per_gpu': B, 'N': N, 'M': N, 'L': L, 'sinkhorn_iterations': T,
            'k': [k or 0 for k in DEFAULT_K], 'loss_method': 'triplet_loss',
            'parallelism': 'batch sharded over %d rank(s); one all_gather_into_tensor of the match results per step when ranks > 1' % world,
            'l2': 'per-step working set (activations + q/k/v digit planes + logits scratch, ~0.6 GB at cfg2) exceeds the 126 MB L2'}


def load_weights(L):
    """(state dict of torch tensors, description). Pre-trained weights when the fixture travelled
    with the snapshot (L=9 only), else seeded random weights of the same architecture."""
    from mdgat_matcher_b200 import synth
    if L == 9 and os.path.isfile(PRETRAINED):
        with np.load(PRETRAINED) as z:
            sd = {k: torch.from_numpy(z[k].astype(np.float64) if z[k].dtype.kind == 'f' else z[k]) for k in z.files}
        return sd, 'pre-trained checkpoint weights (fp64(fp32), test.py load order)'
    return synth.seeded_state_dict(L, 0), 'seeded random-init weights'


def flops_per_pair(N, M, L):
    """Dense FLOPs EXECUTED for one pair: (linear layers incl. encoders/final/score, attention).
    SURVEY.md 8d counts 20*D^2 per point and layer for the reference; the merge conv (2*D^2 of
    those) is folded into the MLP weights by the packer and never runs, so 18*D^2 are executed."""
    D = 128
    enc = 106880 * (N + M)
    lin = 2 * L * (18 * D * D) * (N + M)
    attn_self = 4 * D * (N * N + M * M)
    attn_cross = 4 * D * (2 * N * M)
    attn = L * (attn_self + attn_cross)
    final = 2 * D * D * (N + M) + 2 * D * N * M
    return enc + lin + final, attn


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:       # NVML missing: clocks are reported as unavailable, never invented
            self.nv, self.err = None, repr(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake',
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if self.nv is None or not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'note': 'NVML unavailable'}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


def reference_cpu_net(L, T):
    """(net, description): the UNMODIFIED reference module (oracle/_ref/reference/models/mdgat.py, a byte-identical copy
    of /root/reference/models/mdgat.py placed there by oracle/build_ref.py; or /root/reference itself in the build
    container), fp64, test.py load order, run on the host through oracle/ref_loader.py's torch-name shim for the
    hard-coded torch.device('cuda') of mdgat.py:200. None when neither tree exists."""
    from oracle import ref_loader as RL
    if not RL.reference_available():
        return None, 'reference tree not available'
    cfg = RL.net_config(L=L, sinkhorn_iterations=T)
    if L == 9 and (os.path.isfile(PRETRAINED) or os.path.isfile(RL.CHECKPOINT)):
        net, mod, zcap = RL.build_reference_net(cfg, 'checkpoint')
        return net, 'pre-trained checkpoint weights (fp64(fp32), test.py load order)'
    from mdgat_matcher_b200 import synth
    sd = synth.seeded_state_dict(L, 0)
    net, mod, zcap = RL.build_reference_net(cfg, {'module.' + k: v for k, v in sd.items()})
    return net, 'seeded random-init weights'


def reference_cpu_time(net, data, pairs):
    """Seconds of one reference forward over the first `pairs` pairs of `data` (host tensors)."""
    from oracle import ref_loader as RL
    d = {k: v[:pairs] for k, v in data.items()}
    t0 = time.perf_counter()
    RL.run_reference(net, d)
    return time.perf_counter() - t0


def pick_sample_pairs(per_pair_s, batch, steps, budget_s):
    """Largest of batch, batch/2, batch/4, .. whose `steps` forwards fit the time budget (at least 1)."""
    pairs = batch
    while pairs > 1 and per_pair_s * pairs * steps > budget_s:
        pairs //= 2
    return max(pairs, 1)


def run_reference_arm(args, rank):
    """--impl reference: the reference's own implementation of the path on the host cores -- the unmodified
    models/mdgat.py under torch with every host thread. Each step is one forward over a
    bounded sample of the batch: the full batch when K steps of it fit ~4 minutes, else a power-of-two fraction."""
    if rank != 0:
        return
    from mdgat_matcher_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B, N, L, T = args.batch, args.n, args.layers, args.sinkhorn
    net, wdesc = reference_cpu_net(L, T)
    kind = 'reference'
    if net is None:
        raise SystemExit('bench.py --impl reference: oracle/_ref/reference is missing; run `python __graft_entry__.py build` '
                         'where /root/reference exists (the snapshot carries oracle/_ref to the GPU box)')
    data = synth.make_batch(1000, B, N)
    with torch.no_grad():
        probe = min(4, B)
        reference_cpu_time(net, data, probe)                       # first call: thread pool, allocator
        per_pair = reference_cpu_time(net, data, probe) / probe
        pairs = args.cpu_pairs if args.cpu_pairs > 0 else pick_sample_pairs(per_pair, B, args.steps + 1, args.cpu_budget)
        for _ in range(args.warmup):
            reference_cpu_time(net, data, pairs if _ == 0 else probe)
        t0 = time.perf_counter()
        for s in range(args.steps):
            reference_cpu_time(net, data, pairs)
        dt = time.perf_counter() - t0
    value = pairs * args.steps / dt
    sample = ('%d steps x one forward over %d of the %d pairs (N=M=%d, L=%d, T=%d), unmodified reference models/mdgat.py, torch %s, '
              '%d threads' % (args.steps, pairs, B, N, L, T, torch.__version__, threads))
    line = {
        'impl': 'reference', 'metric': 'keypoint-pairs/sec', 'value': value, 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic in-distribution keypoint pairs (SURVEY 8d generator); ' + wdesc,
        'config': workload_config(B, N, L, T, 1),
        'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def digit_products(S, SP):
    """int8 plane products per float64 contraction: GEMM / Q K^T keep the pairs s + t <= S - 1; pass 1 of the attention
    kernel multiplies the 3 leading pairs; P V keeps a + t <= S - 1 for the SP planes of P."""
    return S * (S + 1) // 2, 3, SP * S - SP * (SP - 1) // 2


def sweep_parity(out, seed, B, N, L, T):
    """Compares this run's outputs with the stored outputs of the UNMODIFIED reference for the same seeded batch
    (tests/golden/sweep, oracle/gen_sweep.py), when that batch is the one being timed."""
    path = os.path.join(ROOT, 'tests', 'golden', 'sweep', 'cfg2_s%d_b%d.npz' % (seed, B))
    if (N, L, T) != (512, 9, 100) or not os.path.isfile(path):
        return None
    ref = np.load(path)
    flips, errs = 0, []
    for side in ('0', '1'):
        flips += int((out['matches' + side].cpu().numpy() != ref['matches' + side].astype(np.int64)).sum())
        errs.append(np.abs(out['matching_scores' + side].cpu().numpy() - ref['matching_scores' + side]).ravel())
    e = np.concatenate(errs)
    return {'against': 'unmodified reference outputs for the timed batch (tests/golden/sweep/%s)' % os.path.basename(path),
            'rows': int(e.size), 'index_flips': flips, 'max_score_err': float(e.max()), 'p99_score_err': float(np.quantile(e, 0.99)),
            'loss_err': abs(float(out['loss']) - float(ref['loss'])), 'bar': 'indices exact, scores <= 1e-4'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=32, help='pairs per GPU per step')
    ap.add_argument('--n', type=int, default=512, help='keypoints per set')
    ap.add_argument('--layers', type=int, default=9, help='L (2L GNN layers)')
    ap.add_argument('--sinkhorn', type=int, default=100)
    ap.add_argument('--cpu-pairs', type=int, default=0, help='pairs per step of the CPU reference (0 = auto)')
    ap.add_argument('--cpu-budget', type=float, default=240.0, help='seconds the --impl reference run may take')
    ap.add_argument('--gemm', default='tcgen05_i8', choices=['tcgen05_i8', 'dmma'], help='engine of the per-layer projections')
    ap.add_argument('--attention', default='tcgen05_i8', choices=['tcgen05_i8', 'tcgen05_i8_all', 'dmma'], help='engine of Q K^T / P V')
    ap.add_argument('--precision', default=None, choices=['sweep', 'exact'],
                    help="digit planes: 'sweep' = the smallest counts that pass the 131 k-row parity sweep (module default), 'exact' = 7/7/6")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eager', action='store_true', help='skip the eager-GPU timing of the unmodified reference')
    ap.add_argument('--no-latency', action='store_true', help='skip the batch-1 latency measurement')
    ap.add_argument('--cuda-graph', action='store_true', help="(default) run the timed forwards with config['cuda_graph']: the launch sequence of one forward captured once, replayed per step")
    ap.add_argument('--no-cuda-graph', action='store_true', help='plain kernel launches (programmatic dependent launch) instead of the captured graph')
    args = ap.parse_args()

    # NCCL prints its version banner on STDOUT at level VERSION; this script owes the driver ONE json line
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py (impl b200) needs a CUDA device; there is no CPU fallback')
    from mdgat_matcher_b200 import synth, _capi, ops, dist as mdist
    from mdgat_matcher_b200.models.mdgat import MDGAT

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        mdist.init_process_group('nccl', device_id=dev)

    B, N, L, T = args.batch, args.n, args.layers, args.sinkhorn
    cfg = net_config(L, T)
    cfg['gemm'], cfg['attention'] = args.gemm, args.attention
    if args.precision:
        cfg['precision'] = args.precision
    cfg['cuda_graph'] = not args.no_cuda_graph
    sd, wdesc = load_weights(L)
    net = MDGAT(cfg)
    net.load_state_dict(sd)
    net = net.double().eval().to(dev)
    planes = net.digit_planes()                                   # (gemm S, attention S, attention SP)

    seed = 1000 + rank
    host = synth.make_batch(seed, B, N)                           # every rank its own pairs
    host = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    out_keys = ('matches0', 'matches1', 'matching_scores0', 'matching_scores1', 'loss')

    @torch.no_grad()
    def step_resident():
        d = dict(resident)
        d['gt_matches0'] = resident['gt_matches0'].clone()      # forward rewrites gt in place (mdgat.py:519)
        d['gt_matches1'] = resident['gt_matches1'].clone()
        o = net(d)
        if world > 1:
            # the one collective of the path, every step: all ranks' match results
            o['gathered'] = mdist.all_gather_outputs(o)
        return o

    # End-to-end step as a pipelined caller drives it (a loader that prefetches, a consumer that reads the results of
    # step i-1 while step i runs): the pinned host inputs of step i+1 travel on a copy stream during step i, the results
    # of every step are copied to pinned host memory and waited for one step later. Every step still moves all its
    # inputs H2D and all its results D2H inside the timed region; nothing is skipped or cached.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    slot_ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    host_out = [{}, {}]
    keep = [None, None]
    e2e_state = {'i': 0, 'primed': False, 'consumed': 0.0}

    def h2d_async(slot):
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                slots[slot][k].copy_(v, non_blocking=True)
            slot_ready[slot].record(copy_stream)

    def e2e_drain():
        i = e2e_state['i']
        if i > 0:
            done[(i - 1) % 2].synchronize()
            e2e_state['consumed'] += float(host_out[(i - 1) % 2]['loss'])

    @torch.no_grad()
    def step_e2e():
        i = e2e_state['i']
        if not e2e_state['primed']:
            h2d_async(i % 2)
            e2e_state['primed'] = True
        torch.cuda.current_stream().wait_event(slot_ready[i % 2])
        o = net(slots[i % 2])
        if world > 1:
            o['gathered'] = mdist.all_gather_outputs(o)
        for k in out_keys:
            if k not in host_out[i % 2]:
                host_out[i % 2][k] = torch.empty(o[k].shape, dtype=o[k].dtype).pin_memory()
            host_out[i % 2][k].copy_(o[k], non_blocking=True)
        done[i % 2].record()
        keep[i % 2] = o
        e2e_drain()                                             # results of step i-1 are on the host: the caller reads them
        h2d_async((i + 1) % 2)                                  # inputs of step i+1 (that slot's last reader, step i-1, is done)
        e2e_state['i'] = i + 1
        return o

    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _capi.lib.mdgat_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_resident()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = _capi.lib.mdgat_launch_count() - l0

    # ---------------- per-stage device times: a separate pass of the same K steps with the stage events switched on
    # (~230 cudaEventRecord calls per forward cost ~0.5 ms per step, so they stay out of the timed region above)
    _capi.lib.mdgat_profile_enable(1)
    graph_on = bool(net.config.get('cuda_graph', False))
    net.config['cuda_graph'] = False                           # stage events are host-side records: plain launches for this pass
    for _ in range(args.steps):
        step_resident()
    torch.cuda.synchronize()
    net.config['cuda_graph'] = graph_on
    stages = _capi.profile_collect()
    _capi.lib.mdgat_profile_enable(0)

    # ---------------- timed region 2: end to end through MDGAT.forward with host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_drain()                                                 # results of the last step
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        assert out['gathered']['matches0'].shape[0] == B * world
    ms_total, e2e_ms = float(t[0]), float(t[1])

    if rank != 0:
        torch.distributed.destroy_process_group()
        return

    pairs_total = B * world * args.steps
    value = pairs_total / (ms_total * 1e-3)
    e2e_value = pairs_total / (e2e_ms * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = sum(v.numel() * v.element_size() for v in host_out[0].values())
    config = workload_config(B, N, L, T, world)
    late = net.digit_planes_late()
    config['digit_planes'] = {'gemm': planes[0], 'attention_qkv': planes[1], 'attention_p': planes[2],
                              'late_layers': ({'from_layer': late[0], 'gemm': late[1], 'attention_qkv': late[2], 'attention_p': late[3]} if late else None),
                              'sinkhorn_kernel_matrix': 'float32 storage, float64 arithmetic' if net.sinkhorn_k32() else 'float64',
                              'note': 'int8 planes per float64 operand; chosen on the 131 k-row sweep against the unmodified reference (DESIGN.md 2)'}
    config['launch'] = ("config['cuda_graph']: one captured CUDA graph of the forward's kernel sequence replayed per step (inputs copied into its "
                        "static buffers, results out of them, inside the timed regions)") if cfg['cuda_graph'] else 'plain launches with programmatic dependent launch'
    try:
        sk = net.sinkhorn_status()
        its = sk['iterations']
        config['sinkhorn'] = {'iterations_requested': T,
                              'iterations_run_per_pair': {'min': int(min(its)), 'mean': float(sum(its)) / len(its), 'max': int(max(its))},
                              'fallback_pairs': int(sum(sk['fallback'])),
                              'exit_rule': ('no column scaling moved by more than %s relative in one iteration (skipped iterations move a '
                                            'log-potential by <= (T - t) tol; DESIGN.md 4.6)' % os.environ.get('MDGAT_SK_TOL', '2^-35'))
                                           if net.sinkhorn_k32() else 'iterate repeats bit for bit'}
    except Exception as exc:                                  # diagnostic only
        config['sinkhorn'] = {'status_error': str(exc)}
    if world > 1:
        config['collective'] = {'op': 'all_gather_into_tensor (NCCL)', 'per_step': 1, 'inside_timed_region': True,
                                'bytes_per_rank_per_step': int(B * 2 * (N + N) * 8)}

    # ---------------- roofline of the dominant stage (device time from CUDA events on the launch stream)
    lin_f, attn_f = flops_per_pair(N, N, L)
    dmma_peak, dfma_peak = ops.measure_fp64_peak()
    peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    measured = json.load(open(peaks_file)) if os.path.isfile(peaks_file) else None
    k_sched = [0 if k is None else k for k in ([None] * (2 * L - len(DEFAULT_K)) + DEFAULT_K)][-2 * L:]
    n_topk = sum(1 for k in k_sched if k)
    attn_full_f = attn_f * (2 * L - n_topk) / (2 * L)
    stage_flops = {'gemm': lin_f * B, 'attn_full': attn_full_f * B}
    per_step = {s: v['ms'] / args.steps for s, v in stages.items()}
    dominant = max(('gemm', 'attn_full'), key=lambda s: per_step[s])
    seg = stages[dominant]
    i8 = {'gemm': args.gemm == 'tcgen05_i8', 'attn_full': args.attention != 'dmma'}[dominant]
    fp64_equiv = stage_flops[dominant] / (per_step[dominant] * 1e-3) / 1e12
    if i8:
        # The dominant kernel runs on the int8 tensor pipe: its algorithmic work is the exact digit-plane products
        # (DESIGN.md 4.1 / 4.3), 2 operations per int8 multiply-add; the ceiling is the kind::i8 issue rate measured
        # in this process. The float64 FLOPs those products stand for are reported next to it.
        R = B * 2 * N
        gq, _, _ = digit_products(planes[0], planes[2])
        qk, p1, pv = digit_products(planes[1], planes[2])
        # digit products summed over the layers (a second plane setting applies from layer late[0] on)
        gsum = asum = 0
        for l, kk in enumerate(k_sched):
            pl = (late[1], late[2], late[3]) if (late and l >= late[0]) else planes
            gsum += digit_products(pl[0], pl[2])[0]
            if not kk:
                asum += sum(digit_products(pl[1], pl[2]))
        i8_ops = {'gemm': gsum * 2.0 * R * (384 * 128 + 256 * 256 + 128 * 256),
                  'attn_full': asum * 2.0 * N * N * 32 * 4 * 2 * B}[dominant]
        i8_peak = ops.measure_i8_peak()
        achieved = i8_ops / (per_step[dominant] * 1e-3) / 1e12
        roofline = {
            'bound': 'tensor',
            'kernel': {'gemm': 'ozaki_gemm_kernel (tcgen05.mma.kind::i8, %d exact digit-plane products per float64 GEMM)' % gq,
                       'attn_full': 'attn_i8_kernel (tcgen05.mma.kind::i8, %d exact digit-plane products per float64 Q K^T + P V)' % (qk + p1 + pv)}[dominant],
            'achieved': achieved, 'peak': i8_peak, 'unit': 'TFLOP/s', 'frac': achieved / i8_peak, 'traffic': None,
            'op_kind': 'int8 tensor operations (2 per multiply-add) of the digit-plane products issued per step',
            'peak_source': 'tcgen05.mma kind::i8 issue-rate microbenchmark run in this process (128x256x32 MMAs from shared memory); '
                           'MEASURED_PEAKS.json carries HBM and bf16 figures only',
            'fp64_equivalent': {'achieved_tflops': fp64_equiv, 'fp64_dmma_peak_tflops': dmma_peak, 'ratio_to_fp64_pipe': fp64_equiv / dmma_peak,
                                'note': 'algorithmic float64 FLOPs of the stage (SURVEY 8d) over its device time, against the DMMA ceiling '
                                        'the reference arithmetic would be bound by'},
            'algorithmic_ops_per_step': i8_ops,
            'frac_note': 'frac counts the int8 digit-plane operations the kernel issues, so it falls when a precision setting needs fewer '
                         'planes: round 1 ran 58 digit products per attention contraction (frac 0.19 at 4.02 ms per forward), this build '
                         '%d (the stage takes %.2f ms). The kernel is bound by its SIMT epilogue (ncu: issue slots 55 %%, FP64 pipe 20 %%, '
                         'tensor pipe 14 %%), not by the tensor pipe; fp64_equivalent is the algorithmic float64 rate of SURVEY 8d.'
                         % (qk + p1 + pv, per_step[dominant]) if dominant == 'attn_full' else
                         'frac counts the int8 digit-plane operations the kernel issues (15 products per float64 GEMM at 5 planes)',
        }
    else:
        roofline = {
            'bound': 'tensor', 'kernel': {'gemm': 'gemm_f64_kernel (DMMA.8x8x4)', 'attn_full': 'attn_full_kernel (DMMA.8x8x4)'}[dominant],
            'achieved': fp64_equiv, 'peak': dmma_peak, 'unit': 'TFLOP/s', 'frac': fp64_equiv / dmma_peak, 'traffic': None,
            'peak_source': 'fp64 DMMA issue-loop microbenchmark run in this process (MEASURED_PEAKS.json has no fp64 entry; '
                           'the path computes in float64 for parity, tcgen05 has no f64 kind)',
        }
    # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture of this workload
    # (profiles/r*_traffic.json, written by tools/summarize_profiles.py); next to it the bytes the kernel has to move at least
    if (B, N, L, T) == (32, 512, 9, 100):
        import glob
        tfiles = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_traffic.json')))
        if tfiles:
            tj = json.load(open(tfiles[-1]))
            want = {'gemm': 'ozaki_gemm_kernel' if i8 else 'gemm_f64_kernel', 'attn_full': 'attn_i8_kernel' if i8 else 'attn_full_kernel'}[dominant]
            hits = [v for k, v in tj['kernels'].items() if want in k]
            if hits:
                roofline['traffic'] = sum(h['dram_bytes_per_launch'] * h['launches'] for h in hits) / sum(h['launches'] for h in hits)
                roofline['traffic_source'] = '%s: %s' % (os.path.basename(tfiles[-1]), tj['source'])
                if dominant == 'attn_full' and i8:
                    # digit planes of q, k, v read once and the message planes written once (scales are < 1 %)
                    roofline['compulsory_bytes_per_launch'] = float(B * 2 * N * 128 * (3 * planes[1] + planes[0]))
    roofline.update({
        'dfma_peak_tflops': dfma_peak,
        'launches_per_step': seg['launches'] / args.steps,
        'avg_launch_ms': seg['ms'] / max(seg['launches'], 1),
        'algorithmic_flops_per_step': stage_flops[dominant],
        'measured_peaks_file': measured,
        'stage_ms_per_step': per_step,
        'stage_share': {s: per_step[s] / max(sum(per_step.values()), 1e-9) for s in per_step},
        'all_fp64_tflops': (lin_f + attn_f) * B / ((per_step['gemm'] + per_step['attn_full'] + per_step['attn_topk'] + per_step.get('slice', 0.0)) * 1e-3) / 1e12,
    })

    parity = sweep_parity(out, seed, B, N, L, T)

    # ---------------- the reference's own CPU path, bounded sample (the unmodified models/mdgat.py under torch)
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        ref_net, _ = reference_cpu_net(L, T)
        if ref_net is not None:
            cpu_data = {k: v.clone() for k, v in host.items()}
            with torch.no_grad():
                probe = min(2, B)
                reference_cpu_time(ref_net, cpu_data, probe)
                per_pair = reference_cpu_time(ref_net, cpu_data, probe) / probe
                pairs = args.cpu_pairs if args.cpu_pairs > 0 else pick_sample_pairs(per_pair, B, 1, 20.0)
                dt = reference_cpu_time(ref_net, cpu_data, pairs)
            cpu_baseline = {'value': pairs / dt, 'unit': 'pairs/s', 'cores': threads, 'kind': 'reference',
                            'sample': 'one forward over %d of the %d timed pairs (N=M=%d, L=%d, T=%d; %.1f s), unmodified reference '
                                      'models/mdgat.py under torch %s on %d host threads' % (pairs, B, N, L, T, dt, torch.__version__, threads)}
            del ref_net
        else:
            cpu_baseline = {'value': None, 'unit': 'pairs/s', 'cores': threads, 'kind': 'reference',
                            'sample': 'unavailable: oracle/_ref/reference missing (run __graft_entry__.build() where /root/reference exists)'}

    # ---------------- the reference's own eager-GPU path: the UNMODIFIED models/mdgat.py on this GPU, fp64, same inputs
    # (the denominator of the >= 10x north-star target). No shim is needed on a CUDA box.
    eager = None
    if not args.no_eager and world == 1:
        try:
            from oracle import ref_loader as RL
            if not RL.reference_available():
                raise FileNotFoundError('oracle/_ref/reference missing')
            rcfg = RL.net_config(L=L, sinkhorn_iterations=T)
            if L == 9:
                rnet, rmod, zcap = RL.build_reference_net(rcfg, 'checkpoint', target=dev)
            else:
                rnet, rmod, zcap = RL.build_reference_net(rcfg, {'module.' + k: v for k, v in sd.items()}, target=dev)
            rmodule = rnet.module                                    # the MDGAT module itself (DataParallel would replicate over every visible GPU)
            with torch.no_grad():
                def step_eager():
                    d = dict(resident)
                    d['gt_matches0'] = resident['gt_matches0'].clone()
                    d['gt_matches1'] = resident['gt_matches1'].clone()
                    return rmodule(d)
                for _ in range(3):
                    oe = step_eager()
                torch.cuda.synchronize()
                times = []
                for _ in range(10):
                    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    g0.record()
                    oe = step_eager()
                    g1.record()
                    torch.cuda.synchronize()
                    times.append(g0.elapsed_time(g1))
            ems = float(np.median(times))
            eager = {'ms_per_step': ems, 'min_ms_per_step': float(min(times)), 'pairs_per_s': B / (ems * 1e-3),
                     'speedup_of_value': value / (B / (ems * 1e-3)), 'speedup_of_e2e': e2e_value / (B / (ems * 1e-3)),
                     'matches_equal': bool(torch.equal(oe['matches0'], out['matches0']) and torch.equal(oe['matches1'], out['matches1'])),
                     'max_score_diff': float(max((oe['matching_scores0'] - out['matching_scores0']).abs().max(),
                                                 (oe['matching_scores1'] - out['matching_scores1']).abs().max())),
                     'peak_memory_gb': torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                     'what': 'UNMODIFIED reference models/mdgat.py (MDGAT.forward, eval, fp64, checkpoint weights in test.py load order) as eager '
                             'PyTorch %s on this GPU, same resident inputs, 3 warm-up + 10 timed forwards, median (CUDA events)' % torch.__version__}
            del oe, rnet, rmodule
            torch.cuda.empty_cache()
        except Exception as e:          # e.g. out of memory on the retained attn.prob tensors at N = 2048
            eager = {'error': repr(e)[:300]}

    # ---------------- single-call latency at the KITTI-like shape (batch 1, 2 x 256 keypoints, T = 20: cfg5's step)
    latency = None
    if not args.no_latency and world == 1:
        latency = batch1_latency(dev, sd, L)

    line = {
        'metric': 'keypoint-pairs/sec', 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic in-distribution keypoint pairs (SURVEY 8d generator); ' + wdesc,
        'config': config,
        'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': int(launches),
        'clocks': sampler.summary(),
        'roofline': roofline,
        'cpu_baseline': cpu_baseline,
        'parity': parity,
        'gpu_eager_reference': eager,
        'latency_batch1': latency,
        'candidate_correspondences_per_s': value * N * N,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def batch1_latency(dev, sd, L):
    """Milliseconds of ONE MDGAT.forward call at batch 1, 2 x 256 keypoints, T = 20 (the per-step shape of
    test_registration_metric.py, BASELINE config 5): plain launches, and replayed from a captured CUDA graph."""
    from mdgat_matcher_b200 import synth
    from mdgat_matcher_b200.models.mdgat import MDGAT
    try:
        res = {}
        data = {k: v.to(dev) for k, v in synth.make_batch(77, 1, 256).items()}
        for graph in (False, True):
            cfg = net_config(L, 20)
            cfg['cuda_graph'] = graph
            net = MDGAT(cfg)
            net.load_state_dict(sd)
            net = net.double().eval().to(dev)
            with torch.no_grad():
                def call():
                    d = dict(data)
                    d['gt_matches0'], d['gt_matches1'] = data['gt_matches0'].clone(), data['gt_matches1'].clone()
                    return net(d)
                for _ in range(5):
                    o = call()
                torch.cuda.synchronize()
                times = []
                for _ in range(30):
                    t0 = time.perf_counter()
                    o = call()
                    float(o['loss'])                                  # the caller reads a result: device -> host, synchronises
                    times.append((time.perf_counter() - t0) * 1e3)
            res['graph' if graph else 'launches'] = {'median_ms': float(np.median(times)), 'min_ms': float(min(times))}
        res['shape'] = 'batch 1, 2x256 keypoints, L=%d, T=20; wall clock of forward() + reading the loss on the host' % L
        # the unmodified reference, eager fp64 on the same GPU, same call
        try:
            from oracle import ref_loader as RL
            if RL.reference_available():
                rcfg = RL.net_config(L=L, sinkhorn_iterations=20)
                weights = 'checkpoint' if (L == 9 and (os.path.isfile(PRETRAINED) or os.path.isfile(RL.CHECKPOINT))) else {'module.' + k: v for k, v in sd.items()}
                rnet, _mod, _z = RL.build_reference_net(rcfg, weights, target=str(dev))
                with torch.no_grad():
                    def rcall():
                        return rnet.module({k: v.clone() for k, v in data.items()})
                    for _ in range(3):
                        o = rcall()
                    torch.cuda.synchronize()
                    times = []
                    for _ in range(10):
                        t0 = time.perf_counter()
                        o = rcall()
                        float(o['loss'])
                        times.append((time.perf_counter() - t0) * 1e3)
                res['reference_eager'] = {'median_ms': float(np.median(times)), 'min_ms': float(min(times)),
                                          'what': 'unmodified reference models/mdgat.py, eager PyTorch fp64 on this GPU, same call'}
                res['speedup_vs_reference_eager'] = res['reference_eager']['median_ms'] / res['launches']['median_ms']
                del rnet
        except Exception as e:
            res['reference_eager'] = {'error': repr(e)[:200]}
        return res
    except Exception as e:
        return {'error': repr(e)[:300]}


if __name__ == '__main__':
    main()
