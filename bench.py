#!/usr/bin/env python
"""bench.py — RefineDet512 detect-stage throughput (decode + NMS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sparse|dense] [--impl reference]

A *step* is one pass of the detect stage (ARM filter, two-stage decode, per-class threshold,
top-k 1000, pixel NMS 0.45, keep 500/class) over a batch of 32 synthetic images of
BASELINE.json config 3 (P = 16,320 anchors, C = 81 classes); every rank owns its own batch
(weak scaling, no data-path collective).  One JSON line on stdout (rank 0):

  value        images/s, device time (CUDA events), inputs resident in HBM, L2 flushed and
               input buffers rotated between steps, max over ranks
  e2e          images/s through the public API with HOST (pinned) inputs: H2D copies, the stage's
               kernels, packing, D2H of counts + packed rows, all inside the timed region
  roofline     algorithmic bytes of the stage (5,940,480 B/image) / device time
               of the stage's kernels, against MEASURED_PEAKS.json hbm_gbs; per-kernel shares
  cpu_baseline the numpy oracle (port of the reference's CPU path) on a bounded sample
  clocks       SM clock / throttle reasons sampled through NVML during the timed regions

``--impl reference`` times the CPU port of the reference path (oracle/) on the box's host
cores and prints the same line with "impl": "reference".
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'RefineDet512 detect-stage images/s (decode+NMS)'
UNIT = 'images/s'
SIZE, P, C, BATCH = '512', 16320, 81, 32
TOP_K, KEEP_TOP_K, CONF_THR, NMS_THR, OBJ_THR = 1000, 500, 0.01, 0.45, 0.01
BYTES_PER_IMAGE = 4 * P * (4 + 2 + 4 + C)           # SURVEY.md §8d, + 20 B per kept row (added at run time)
NBUF = 8                                             # rotated device input sets (8 x 190 MB >> 126 MB L2)
NBUF_HOST = 4                                        # rotated pinned host input sets (e2e)
TRAFFIC_FILE = 'r01_traffic.json'


def seed_for(rank, buf=0):
    return 1234 + 1000 * 3 + rank + 100 * buf        # 1234 + 1000*config + rank


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """Polls NVML (SM clock, throttle reasons) from a thread while a timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self._err = repr(e)
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        if self._h is None:
            return
        nv = self._nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
            getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake_slowdown',
        }
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def __enter__(self):
        self._active.set()
        return self

    def __exit__(self, *a):
        self._active.clear()

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': (s[len(s) // 2] if s else None), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


# ---------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle) — the cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_one_image(args):
    """One image through the reference's CPU path: Detect_RefineDet.forward +
    eval_refinedet_coco.py:205-232 (numpy port in oracle/box_oracle.py)."""
    seed, kind = args
    from oracle import box_oracle as bo
    from refinedet.pytorch_b200 import synthetic
    key = (seed, kind)
    if key not in _CPU_CACHE:
        torch.set_num_threads(1)
        a = [t.numpy() for t in synthetic.detect_inputs(seed, 1, P, C, kind)]
        _CPU_CACHE.clear()
        _CPU_CACHE[key] = a + [bo.prior_box(bo.REFINEDET_CFG[SIZE])]
        return 0.0                                            # warm-up call: inputs only
    arm_loc, arm_conf, odm_loc, odm_conf, priors = _CPU_CACHE[key]
    t0 = time.perf_counter()
    boxes, scores = bo.detect_forward(arm_loc, arm_conf, odm_loc, odm_conf.copy(), priors, OBJ_THR)
    bo.detect_stage_eval(boxes[0], scores[0], np.array([512.0] * 4, np.float32), CONF_THR, TOP_K, NMS_THR,
                         KEEP_TOP_K)
    return time.perf_counter() - t0


def cpu_reference_run(kind, steps, warmup, cores, budget_s=None):
    """Each step = `cores` images, one per worker process.  Returns (images/s, steps done, sample)."""
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        jobs = [(seed_for(0) + 7 * w, kind) for w in range(cores)]
        pool.map(_cpu_one_image, jobs, chunksize=1)           # generate inputs in the workers
        for _ in range(warmup):
            pool.map(_cpu_one_image, jobs, chunksize=1)
        t_begin = time.perf_counter()
        done = 0
        for _ in range(steps):
            pool.map(_cpu_one_image, jobs, chunksize=1)
            done += 1
            if budget_s is not None and time.perf_counter() - t_begin > budget_s:
                break
        elapsed = time.perf_counter() - t_begin
    return cores * done / elapsed, done, elapsed


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='sparse', choices=['sparse', 'dense'])
    ap.add_argument('--streams', type=int, default=4, help='batches in flight (lanes: stream + workspace + outputs)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-secondary', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    workload = 'RefineDet512 COCO detect stage: B=%d/GPU, P=%d, C=%d, top_k=%d, keep_top_k=%d, %s generator' % (
        BATCH, P, C, TOP_K, KEEP_TOP_K, args.workload)
    config = {'workload': workload, 'batch_per_gpu': BATCH, 'anchors': P, 'classes': C, 'generator': args.workload,
              'sharding': 'images sharded by rank, no data-path collective'}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        value, done, elapsed = cpu_reference_run(args.workload, steps, max(0, args.warmup), cores, budget_s=150.0)
        sample = '%d steps x %d images (1 per worker process), numpy port of Detect_RefineDet.forward + ' \
                 'eval_refinedet_coco.py:205-232 with py_cpu_nms semantics' % (done, cores)
        line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': done, 'warmup': args.warmup, 'ms_per_step': 1e3 * elapsed / done,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
                'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return 0

    # ---- native arm -------------------------------------------------------------------------
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (native arm) needs a CUDA device: there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    import refinedet.pytorch_b200 as rd
    from refinedet.pytorch_b200 import _ffi, synthetic
    _ffi.lib()

    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[SIZE]).forward().to(dev)
    S = max(1, args.streams)
    host_sets = [[t.pin_memory() for t in synthetic.detect_inputs(seed_for(rank, i), BATCH, P, C, args.workload)]
                 for i in range(NBUF_HOST)]
    dev_sets = [[t.to(dev) for t in hs] for hs in host_sets]
    dev_sets += [[t.to(dev) for t in synthetic.detect_inputs(seed_for(rank, i), BATCH, P, C, args.workload)]
                 for i in range(NBUF_HOST, NBUF)]
    arm_pass = float((host_sets[0][1][..., 1] > OBJ_THR).float().mean())
    scale = torch.tensor([512.0] * 4, device=dev).reshape(1, 4).expand(BATCH, 4).contiguous()
    det = rd.Detect_RefineDet(C, 512, 0, TOP_K, CONF_THR, NMS_THR, OBJ_THR, KEEP_TOP_K)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    clocks = ClockSampler(local_rank)

    # S lanes = batches in flight: each lane owns a stream, a workspace and output slots; one plan
    # (CUDA graph of the launch chain, rd_detect_plan_*) per (lane, input set)
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    lanes = [(det.new_workspace(BATCH, P, dev), det.new_outputs(BATCH, dev)) for _ in range(S)]
    plans = [[det.plan(a[0], a[1], a[2], a[3], priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1])
              for a in dev_sets] for l in range(S)]

    def step(i):
        return plans[i % S][i % NBUF].launch(streams[i % S])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(3, args.warmup)
    for i in range(max(W, S * NBUF)):
        res = step(i)
    torch.cuda.synchronize()
    kept_rows = int(plans[0][0].launch(streams[0]).counts.sum())
    # the replayed plan and the direct launch chain must agree (same kernels, same buffers)
    chk = det.detect(dev_sets[0][0], dev_sets[0][1], dev_sets[0][2], dev_sets[0][3], priors, scale=scale)
    torch.cuda.synchronize()
    if not torch.equal(chk.counts, plans[0][0].result.counts):
        raise RuntimeError('plan replay and direct launch disagree')

    # ---- timed region: exactly K steps, S batches in flight, one event pair around all of them --------
    K = max(1, args.steps)
    main = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = _ffi.launch_count()
    with clocks:
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        for i in range(K):
            step(i)
        for st in streams:
            main.wait_stream(st)
        e1.record(main)
        barrier()
    launches = _ffi.launch_count() - launches0
    local_total_ms = float(e0.elapsed_time(e1))
    total_ms = local_total_ms
    if dist is not None:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = world * BATCH * K / (total_ms * 1e-3)

    # ---- latency of ONE batch: single stream, L2 flushed (untimed) before every step -------------------
    KL = min(K, 50)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(KL)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(KL)]
    with clocks:
        for i in range(KL):
            flush_buf.zero_()
            starts[i].record(main)
            plans[0][i % NBUF].launch(main)
            stops[i].record(main)
        torch.cuda.synchronize()
    lat = sorted(s_.elapsed_time(e_) for s_, e_ in zip(starts, stops))
    latency_ms = float(sum(lat)) / KL

    class _Flush(object):
        def zero_(self):
            flush_buf.zero_()
    # per-kernel breakdown (separate pass: stage events recorded between the kernels)
    kern = det.profile_stage(dev_sets, priors, scale, _Flush(), steps=min(K, 20))
    stage_ms_serialised = sum(kern.values())
    bytes_per_launch = BATCH * BYTES_PER_IMAGE + 20 * kept_rows
    peak, peak_src = measured_peak()
    # The stage is one launch chain (collect -> graph || nms_small -> nms_large), replayed as one CUDA
    # graph; its kernels overlap (programmatic dependent launch) and so do the chains of the S batches in
    # flight, so the duration that counts is the event-timed region / K.
    achieved = bytes_per_launch / (local_total_ms / K * 1e-3) / 1e9
    achieved_single = bytes_per_launch / (latency_ms * 1e-3) / 1e9
    dominant = max(kern, key=kern.get)
    groups = {'collect_kernel': ['collect_kernel'], 'graph_kernel': ['graph_kernel'],
              'nms_small_large_kernels': ['nms_small_kernel', 'sort_kernel', 'resolve_kernel', 'nms_large_kernel']}
    traffic, traffic_src = None, None
    try:                                   # DRAM bytes per launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, 'profiles', TRAFFIC_FILE)) as f:
            tj = json.load(f)
        per_group = {g: sum(tj['kernels'][k]['traffic_bytes'] for k in ks if k in tj['kernels'])
                     for g, ks in groups.items()}
        traffic = tj['stage_traffic_bytes']
        traffic_src = {'file': 'profiles/' + TRAFFIC_FILE, 'what': 'dram__bytes_read.sum + dram__bytes_write.sum of '
                       'the whole launch chain, one ncu --set full capture', 'per_kernel_group': per_group}
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                'kernel': 'detect stage launch chain (dominant group: %s)' % dominant,
                'algorithmic_bytes_per_launch': bytes_per_launch,
                'stage_ms': local_total_ms / K, 'batches_in_flight': S,
                'single_batch': {'ms': latency_ms, 'median_ms': lat[KL // 2], 'achieved': achieved_single,
                                 'frac': achieved_single / peak,
                                 'how': 'one stream, 512 MiB memset (L2 flush, untimed) before every step, %d steps' % KL},
                'stage_ms_serialised': stage_ms_serialised,
                'kernels_ms': kern, 'kernel_share': {k: v / stage_ms_serialised for k, v in kern.items()},
                'note': 'achieved = stage bytes (SURVEY 8d: 4*P*(10+C) B/image + 20 B/kept row) / (event-timed region '
                        '/ K steps), %d batches in flight; ARM-filtered anchors (%.1f%% here) are skipped, so DRAM '
                        'traffic is far below the algorithmic bytes: the stage is issue/latency-bound, not '
                        'bandwidth-bound' % (S, 100 * (1 - arm_pass))}

    # e2e: host (pinned) inputs -> kernels read them over PCIe -> pack -> rows land in pinned host memory
    e2e = None
    if not args.no_e2e:
        e2e_steps = min(K, 40)
        full_bytes = sum(t.numel() * t.element_size() for t in host_sets[0])
        pipe = rd.DetectHostPipeline(det, priors, scale, BATCH, lanes=S)

        def run_pipe():
            d2h = 0
            for i in range(S * NBUF_HOST):                        # plans of every (lane, host set)
                pipe.result(pipe.submit(host_sets[i % NBUF_HOST]))
            barrier()
            with clocks:
                t0 = time.perf_counter()
                pending = []
                for i in range(e2e_steps):
                    pending.append(pipe.submit(host_sets[i % NBUF_HOST]))
                    if len(pending) == S:
                        counts_h, rows_h = pipe.result(pending.pop(0))
                        d2h = counts_h.numel() * 4 + rows_h.numel() * 4 + 4
                while pending:
                    counts_h, rows_h = pipe.result(pending.pop(0))
                barrier()
                sec = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([sec], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
            return sec, d2h

        def run_serial(zero_copy):
            stage = [torch.empty_like(t, device=dev) for t in host_sets[0]]
            for i in range(2):
                det.detect_host(host_sets[i % NBUF_HOST], priors, scale, stage, zero_copy=zero_copy)
            barrier()
            n = min(e2e_steps, 10)
            t0 = time.perf_counter()
            for i in range(n):
                det.detect_host(host_sets[i % NBUF_HOST], priors, scale, stage, zero_copy=zero_copy)
            barrier()
            return (time.perf_counter() - t0) / n

        pipe_s, d2h = run_pipe()
        serial_zc = run_serial(True)
        serial_copy = run_serial(False)
        # bytes that cross PCIe host->device per step: arm_conf in full (DMA) + per ARM-passing anchor the whole
        # 128-byte lines covering its odm_conf row (collect_kernel's line-granular fetch) and one 32-byte sector
        # for each of its two loc vectors
        passing = (host_sets[0][1][..., 1] > OBJ_THR).reshape(-1).numpy()
        first = np.flatnonzero(passing).astype(np.int64) * (C * 4)
        lines = (first + C * 4 - 1) // 128 - first // 128 + 1
        zc_bytes = int(host_sets[0][1].numel() * 4 + lines.sum() * 128 + passing.sum() * 2 * 32)
        e2e = {'value': world * BATCH * e2e_steps / pipe_s, 'unit': UNIT, 'h2d_bytes_per_step': zc_bytes,
               'd2h_bytes_per_step': d2h, 'steps': e2e_steps, 'ms_per_step': 1e3 * pipe_s / e2e_steps,
               'batches_in_flight': S,
               'mode': 'DetectHostPipeline: arm_conf by DMA, the other pinned host inputs read by the kernels over PCIe '
                       '(only rows of ARM-passing anchors cross the bus: h2d_bytes_per_step; the tensors hold %d B), '
                       'rows packed on the device and copied back by one DMA of exactly the kept rows, every result '
                       'read on the host' % full_bytes,
               'serial_zero_copy': {'value': world * BATCH / serial_zc, 'ms_per_step': 1e3 * serial_zc},
               'serial_staged_copy': {'value': world * BATCH / serial_copy, 'ms_per_step': 1e3 * serial_copy,
                                      'h2d_bytes_per_step': full_bytes}}
        del pipe

    # secondary (reported, not the headline): the other generator of SURVEY.md §8d on the same config
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        other = 'dense' if args.workload == 'sparse' else 'sparse'
        o_dev = [t.to(dev) for t in synthetic.detect_inputs(seed_for(rank, 9), BATCH, P, C, other)]
        for _ in range(3):
            det.detect(o_dev[0], o_dev[1], o_dev[2], o_dev[3], priors, scale=scale)
        torch.cuda.synchronize()
        n_sec = 10
        s_ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_sec)]
        e_ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_sec)]
        for i in range(n_sec):
            flush_buf.zero_()
            s_ev[i].record()
            r2 = det.detect(o_dev[0], o_dev[1], o_dev[2], o_dev[3], priors, scale=scale)
            e_ev[i].record()
        torch.cuda.synchronize()
        ms2 = sum(a.elapsed_time(b) for a, b in zip(s_ev, e_ev)) / n_sec
        secondary = {'generator': other, 'value': BATCH / (ms2 * 1e-3), 'unit': UNIT, 'ms_per_step': ms2,
                     'arm_pass_fraction': float((o_dev[1][..., 1] > OBJ_THR).float().mean()),
                     'kept_rows_per_step': int(r2.counts.sum()),
                     'note': 'dense = stress case: 86 % of the anchors pass the ARM filter, every class saturates '
                             'top_k = 1000 and goes through the large-problem kernel (radix select + own bins)'}
        del o_dev

    # BASELINE.json configs 2 and 5 (reported, not the headline): RefineDet320 VOC and the 2-class SAR-ship
    # RefineDet512 detect stage, batch 32, both generators, one batch at a time with the L2 flushed before it
    other_configs = None
    if rank == 0 and world == 1 and not args.no_secondary:
        other_configs = {}
        for name, size, dim, C2, nms_thr in (('cfg2_refinedet320_voc', '320', 320.0, 21, 0.45),
                                             ('cfg5_sarship_2class', '512', 512.0, 2, 0.49)):
            pri2 = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().to(dev)
            P2 = pri2.shape[0]
            det2 = rd.Detect_RefineDet(C2, int(dim), 0, TOP_K, CONF_THR, nms_thr, OBJ_THR, KEEP_TOP_K)
            scale2 = torch.tensor([dim] * 4, device=dev).reshape(1, 4).expand(BATCH, 4).contiguous()
            for gen in ('sparse', 'dense'):
                x = [t.to(dev) for t in synthetic.detect_inputs(seed_for(rank, 20), BATCH, P2, C2, gen)]
                for _ in range(3):
                    r2 = det2.detect(x[0], x[1], x[2], x[3], pri2, scale=scale2)
                torch.cuda.synchronize()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
                for a, b in evs:
                    flush_buf.zero_()
                    a.record()
                    det2.detect(x[0], x[1], x[2], x[3], pri2, scale=scale2)
                    b.record()
                torch.cuda.synchronize()
                ms_o = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
                bytes_o = BATCH * 4 * P2 * (10 + C2) + 20 * int(r2.counts.sum())
                other_configs['%s_%s' % (name, gen)] = {
                    'ms_per_step': ms_o, 'value': BATCH / (ms_o * 1e-3), 'unit': UNIT, 'anchors': P2, 'classes': C2,
                    'kept_rows_per_step': int(r2.counts.sum()), 'algorithmic_bytes_per_launch': bytes_o,
                    'frac_of_hbm_peak': bytes_o / (ms_o * 1e-3) / 1e9 / peak}
                del x

    # f-1: logits in (softmax folded into the stage) against torch.softmax + the stage, one stream, L2 flushed
    logits_in = None
    if rank == 0 and world == 1 and not args.no_secondary:
        lg = [t.to(dev) for t in synthetic.detect_logits(seed_for(rank, 0), BATCH, P, C, args.workload)]
        plan_l = det.plan(lg[0], lg[1], lg[2], lg[3], priors, scale=scale, workspace=lanes[0][0], out=lanes[0][1],
                          logits=True)

        def fused():
            plan_l.launch(main)

        def unfused():
            det.detect(lg[0], torch.softmax(lg[1], -1), lg[2], torch.softmax(lg[3], -1), priors, scale=scale)

        def timed(fn, n=20):
            for _ in range(3):
                fn()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for a, b in evs:
                flush_buf.zero_()
                a.record(main)
                fn()
                b.record(main)
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs) / n
        ms_f, ms_u = timed(fused), timed(unfused)
        same = bool(torch.equal(plan_l.launch(main).counts, plans[0][0].launch(main).counts))
        logits_in = {'ms_per_step': ms_f, 'value': BATCH / (ms_f * 1e-3), 'unit': UNIT,
                     'torch_softmax_then_stage_ms': ms_u, 'counts_equal_to_probability_input': same,
                     'note': 'RD_INPUT_LOGITS: the softmax of models/refinedet.py:143-147 inside collect_kernel (rows of '
                             'ARM-passing anchors only) instead of a separate read+write pass over odm_conf'}
        del plan_l

    # a3: Detect_RefineDet.forward as models/refinedet.py:141 calls it (dense boxes + scores out, in-place zeroing)
    a3 = None
    if rank == 0 and world == 1 and not args.no_secondary:
        a = dev_sets[0]
        n3 = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n3 + 2)]
        for s_ev3, e_ev3 in evs:
            conf = a[3].clone()
            flush_buf.zero_()
            s_ev3.record(main)
            det.forward(a[0], a[1], a[2], conf, priors)
            e_ev3.record(main)
        torch.cuda.synchronize()
        ms3 = sum(x.elapsed_time(y) for x, y in evs[2:]) / n3
        bytes3 = BATCH * (BYTES_PER_IMAGE + 4 * P * (4 + C))             # SURVEY 8d "a3 contract only"
        a3 = {'ms_per_step': ms3, 'value': BATCH / (ms3 * 1e-3), 'unit': UNIT,
              'achieved_GBs': bytes3 / (ms3 * 1e-3) / 1e9, 'frac_of_hbm_peak': bytes3 / (ms3 * 1e-3) / 1e9 / peak,
              'algorithmic_bytes_per_launch': bytes3}

    # BASELINE.json config 4 (reported, not the headline): RefineDetMultiBoxLoss training step — ARM + ODM criteria,
    # forward + backward, 50 ground-truth boxes per image, batch 32 (match, conf loss, HNM, reduce, backward kernels)
    train_step = None
    if rank == 0 and world == 1 and not args.no_secondary:
        tp = [t.to(dev) for t in synthetic.train_predictions(seed_for(rank, 40), BATCH, P, C)]
        tg = [t.to(dev) for t in synthetic.targets(seed_for(rank, 41), BATCH, 50, C)]
        arm_crit = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
        odm_crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True)
        leaves = [t.clone().requires_grad_(True) for t in tp]

        def one_step():
            rd.box_utils._PAD_CACHE = None                      # a real step brings new targets: pad them once, not zero times
            preds = (leaves[0], leaves[1], leaves[2], leaves[3], priors)
            al, ac = arm_crit(preds, tg)
            ol, oc = odm_crit(preds, tg)
            (al + ac + ol + oc).backward()                      # train_refinedet.py:252-256
            for t in leaves:
                t.grad = None
        for _ in range(3):
            one_step()
        torch.cuda.synchronize()
        n_t = 10
        t0 = time.perf_counter()
        for _ in range(n_t):
            one_step()
        torch.cuda.synchronize()
        ms_t = 1e3 * (time.perf_counter() - t0) / n_t
        # the same step with the criteria's sync_free extension (no host read of N inside the criteria; the
        # losses are read once after backward, as train_refinedet.py:258-261 does with .item())
        arm_sf = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True, sync_free=True)
        odm_sf = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True, sync_free=True)

        def one_step_sf():
            rd.box_utils._PAD_CACHE = None
            preds = (leaves[0], leaves[1], leaves[2], leaves[3], priors)
            al, ac = arm_sf(preds, tg)
            ol, oc = odm_sf(preds, tg)
            (al + ac + ol + oc).backward()
            vals = torch.stack([al.detach(), ac.detach(), ol.detach(), oc.detach()]).tolist()   # one read per step
            for t in leaves:
                t.grad = None
            return vals
        for _ in range(3):
            one_step_sf()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_t):
            one_step_sf()
        torch.cuda.synchronize()
        ms_sf = 1e3 * (time.perf_counter() - t0) / n_t
        train_step = {'ms_per_step': ms_t, 'value': BATCH / (ms_t * 1e-3), 'unit': UNIT,
                      'sync_free': {'ms_per_step': ms_sf, 'value': BATCH / (ms_sf * 1e-3),
                                    'what': 'criteria built with sync_free=True: no host read of N inside the criteria, '
                                            'the four losses read once after backward'},
                      'what': 'ARM + ODM RefineDetMultiBoxLoss forward + backward (B=32, P=16320, C=81, 50 GT/image), '
                              'wall clock including the host glue and the two N < 1 host checks'}
        del leaves, tp, tg

    # the one exchange step of the multi-GPU path, timed separately from the stage (N > 1 only):
    # the gather of every rank's compact detections, through NCCL and through the fused pack + P2P scatter
    gather = None
    if dist is not None:
        from refinedet.pytorch_b200 import dist as rdist
        res_g = plans[0][0].launch(main)
        torch.cuda.synchronize()

        def timed_gather(fn, n=20):
            for _ in range(3):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / n], device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt) * 1e3
        counts_n, rows_n = rdist.gather_detections(res_g)
        gather = {'rows_per_rank': int(rows_n[rank].shape[0]), 'bytes_per_rank': int(rows_n[rank].numel() * 4),
                  'nccl_ms': timed_gather(lambda: rdist.gather_detections(res_g)),
                  'nccl_what': 'pack kernel, header all_gather, host read, padded all_gather_into_tensor '
                               '(dist.gather_packed), wall clock, max over ranks'}
        try:
            ex = rdist.PeerExchange(BATCH, C, res_g.dets.shape[2], dev)
            ex.exchange(res_g)
            counts_p, rows_p = ex.result()
            same = all(torch.equal(counts_p[r], counts_n[r]) and torch.equal(rows_p[r], rows_n[r])
                       for r in range(world))
            gather.update({'peer_equals_nccl': bool(same),
                           'peer_ms': timed_gather(lambda: (ex.exchange(res_g), ex.result())),
                           'peer_device_ms': timed_gather(lambda: ex.exchange(res_g)),
                           'peer_what': 'rd_pack_scatter: the pack kernel stores counts + rows into every peer\'s '
                                        'buffer over NVLink (symmetric memory), two device barriers; peer_ms adds '
                                        'the host read of the headers'})
            del ex
        except Exception as e:                                  # symmetric memory unavailable: NCCL number only
            gather['peer_error'] = repr(e)[:200]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, done, elapsed = cpu_reference_run(args.workload, 1000, 1, cores, budget_s=12.0)
        cpu_baseline = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': '%d steps x %d images (1 per worker process) of the same workload, %.1f s'
                                  % (done, cores, elapsed)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K,
                'warmup': W, 'ms_per_step': ms_per_step, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': dict(config, l2='inputs larger than L2: every step reads its own 190 MB input set, %d sets '
                                          '(1.5 GB) rotated; no flush inside the timed region' % NBUF,
                               batches_in_flight=S, launch='one CUDA-graph replay per step (rd_detect_plan_launch)',
                               arm_pass_fraction=arm_pass, kept_rows_per_step=kept_rows),
                'latency_ms_per_batch': latency_ms,
                'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'secondary': secondary,
                'other_configs': other_configs, 'gather': gather, 'logits_in': logits_in, 'a3_forward': a3, 'train_step': train_step,
                'gpu_launches': int(launches),
                'clocks': clocks.summary()}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
