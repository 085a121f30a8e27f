#!/usr/bin/env python
"""bench.py — RefineDet512 detect-stage throughput (decode + NMS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sparse|dense] [--scaling weak|strong]
                    [--impl reference]

A *step* is one pass of the detect stage (ARM filter, two-stage decode, per-class threshold,
top-k 1000, pixel NMS 0.45, keep 500/class) over a batch of synthetic images of BASELINE.json
config 3 (P = 16,320 anchors, C = 81 classes): 32 images per rank (``--scaling weak``, default) or
32 images in all, sharded by rank (``--scaling strong``).  With more than one rank the line also
carries ``with_exchange``: the same steps with the path's one exchange inside every step -- the
all-gather of the compact detections (pack + P2P copy over NVLink + rendezvous, ``dist.PeerExchange``),
captured into the step's plan and overlapped with the stage of the next batches on the other lanes.

ONE compact JSON line on stdout (rank 0); the long-form record (notes, per-kernel times, every
secondary measurement) goes to stderr as a second JSON line prefixed ``BENCH_DETAIL``.

  value        images/s: exactly K steps between barrier + synchronize, CUDA events, max over ranks;
               the region is repeated ``--regions`` times and the MEDIAN region is reported
               (``regions`` carries min / median / max).  Inputs resident in HBM; every step reads
               its own 190 MB input set (8 sets rotated, 1.5 GB >> 126 MB L2), no flush inside the
               timed region
  e2e          images/s through the public API with HOST (pinned) inputs: H2D traffic, the stage's
               kernels, packing, D2H of counts + packed rows, all inside the timed region
  roofline     algorithmic bytes of the stage (5,940,480 B/image + 20 B/kept row) /
               event-timed step, against MEASURED_PEAKS.json hbm_gbs; dram_frac = ncu DRAM bytes /
               the same time; single_batch_frac = one batch alone, L2 flushed; dense_frac = the
               stress generator
  cpu_baseline the UNMODIFIED reference (baseline/_ref: Detect_RefineDet.forward + the
               eval_refinedet_coco.py:213-232 loop with py_cpu_nms) on a bounded sample, all host cores
  clocks       SM clock / throttle reasons sampled through NVML during the timed regions

``--impl reference`` times that reference path alone and prints the same line with "impl": "reference".
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


def seed_for(rank, buf=0):
    return 1234 + 1000 * 3 + rank + 100 * buf        # 1234 + 1000*config + rank


def make_config(workload, scaling):
    """The workload description: identical in the native and the reference arm."""
    if scaling == 'strong':
        batch, per_gpu = 'B=%d in all, sharded over the ranks' % BATCH, None
    else:
        batch, per_gpu = 'B=%d/GPU' % BATCH, BATCH
    sharding = 'images sharded by rank; value = the detect stage of every rank (no data-path collective inside it); ' \
               'with_exchange = the same steps with the all-gather of the compact detections over NVLink inside every step'
    return {'workload': 'RefineDet512 COCO detect stage: %s, P=%d, C=%d, top_k=%d, keep_top_k=%d, %s generator'
                        % (batch, P, C, TOP_K, KEEP_TOP_K, workload),
            'batch_per_gpu': per_gpu, 'anchors': P, 'classes': C, 'generator': workload, 'sharding': sharding}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def latest_traffic_file():
    d = os.path.join(ROOT, 'profiles')
    try:
        names = sorted(n for n in os.listdir(d) if n.endswith('_traffic.json'))
    except OSError:
        names = []
    return names[-1] if names else None


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
# the reference's CPU path — cpu_baseline leg and --impl reference (the only users of baseline/ and oracle/ here)
# ---------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _port_one_image(args):
    """Fallback when baseline/_ref is absent: one image through the numpy PORT of the reference's CPU path
    (oracle/box_oracle.py: Detect_RefineDet.forward + eval_refinedet_coco.py:205-232)."""
    seed, kind = args
    from oracle import box_oracle as bo
    from refinedet.pytorch_b200 import synthetic
    key = (seed, kind)
    if key not in _CPU_CACHE:
        torch.set_num_threads(1)
        a = [t.numpy() for t in synthetic.detect_inputs(seed, 1, P, C, kind)]
        _CPU_CACHE.clear()
        _CPU_CACHE[key] = a + [bo.prior_box(bo.REFINEDET_CFG[SIZE])]
        return 0.0
    arm_loc, arm_conf, odm_loc, odm_conf, priors = _CPU_CACHE[key]
    t0 = time.perf_counter()
    boxes, scores = bo.detect_forward(arm_loc, arm_conf, odm_loc, odm_conf.copy(), priors, OBJ_THR)
    bo.detect_stage_eval(boxes[0], scores[0], np.array([512.0] * 4, np.float32), CONF_THR, TOP_K, NMS_THR, KEEP_TOP_K)
    return time.perf_counter() - t0


def cpu_reference_run(kind, steps, warmup, cores, budget_s=None):
    """Each step = ``cores`` images, one per worker process.  Returns (images/s, steps done, seconds, kind, what)."""
    from baseline import reference_arm as ra
    if ra.available():
        v, done, elapsed = ra.run_detect(seed_for(0), kind, P, C, SIZE, (CONF_THR, TOP_K, NMS_THR, KEEP_TOP_K, OBJ_THR),
                                         steps, warmup, cores, budget_s)
        return v, done, elapsed, 'reference', 'unmodified reference (baseline/_ref): Detect_RefineDet.forward + the ' \
            'eval_refinedet_coco.py:213-232 class loop around its py_cpu_nms'
    import multiprocessing as mp
    with mp.get_context('fork').Pool(cores) as pool:
        jobs = [(seed_for(0) + 7 * w, kind) for w in range(cores)]
        pool.map(_port_one_image, jobs, chunksize=1)
        for _ in range(warmup):
            pool.map(_port_one_image, jobs, chunksize=1)
        t_begin = time.perf_counter()
        done = 0
        for _ in range(steps):
            pool.map(_port_one_image, jobs, chunksize=1)
            done += 1
            if budget_s is not None and time.perf_counter() - t_begin > budget_s:
                break
        elapsed = time.perf_counter() - t_begin
    return cores * done / elapsed, done, elapsed, 'port', 'numpy port (oracle/box_oracle.py) of Detect_RefineDet.forward + ' \
        'eval_refinedet_coco.py:205-232 with py_cpu_nms semantics (baseline/_ref absent)'


def bind_to_gpu_numa_node(index):
    """Pin this process (and, by first touch, the pinned host buffers it allocates afterwards) to the CPU cores NVML
    reports as local to GPU ``index``: with one process per GPU the zero-copy reads of the e2e path then cross the
    GPU's own PCIe root instead of the socket interconnect.  Returns the number of cores, or None when unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def median(xs):
    s = sorted(xs)
    return s[len(s) // 2]


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='sparse', choices=['sparse', 'dense'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--regions', type=int, default=31, help='repetitions of the K-step timed region (median reported)')
    ap.add_argument('--streams', type=int, default=4, help='batches in flight (lanes: stream + workspace + outputs)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--no-exchange', action='store_true', help='N > 1: leave the gather out of the step')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    config = make_config(args.workload, args.scaling)

    if args.impl == 'reference':
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        value, done, elapsed, kind, what = cpu_reference_run(args.workload, steps, max(0, args.warmup), cores, budget_s=150.0)
        sample = '%d steps x %d images (1 per worker process); %s' % (done, cores, what)
        line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': done, 'warmup': args.warmup, 'ms_per_step': 1e3 * elapsed / done,
                'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
                'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return 0

    # ---- native arm -------------------------------------------------------------------------
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (native arm) needs a CUDA device: there is no CPU fallback')
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    import refinedet.pytorch_b200 as rd
    from refinedet.pytorch_b200 import _ffi, synthetic
    from refinedet.pytorch_b200 import dist as rdist
    _ffi.lib()
    detail = {}

    if args.scaling == 'strong':
        lo, hi = rdist.shard_range(BATCH, rank, world)
        B_loc, B_total = hi - lo, BATCH
        B_max = (BATCH + world - 1) // world
    else:
        B_loc, B_total, B_max = BATCH, BATCH * world, BATCH

    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[SIZE]).forward().to(dev)
    S = max(1, args.streams)
    host_sets = [[t.pin_memory() for t in synthetic.detect_inputs(seed_for(rank, i), B_loc, P, C, args.workload)]
                 for i in range(NBUF_HOST)]
    dev_sets = [[t.to(dev) for t in hs] for hs in host_sets]
    dev_sets += [[t.to(dev) for t in synthetic.detect_inputs(seed_for(rank, i), B_loc, P, C, args.workload)]
                 for i in range(NBUF_HOST, NBUF)]
    arm_pass = float((host_sets[0][1][..., 1] > OBJ_THR).float().mean())
    scale = torch.tensor([512.0] * 4, device=dev).reshape(1, 4).expand(B_loc, 4).contiguous()
    det = rd.Detect_RefineDet(C, 512, 0, TOP_K, CONF_THR, NMS_THR, OBJ_THR, KEEP_TOP_K)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    clocks = ClockSampler(local_rank)

    # S lanes = batches in flight: each lane owns a stream, a workspace, output slots and (N > 1) an exchange
    # buffer; one plan (CUDA graph of the launch chain, rd_detect_plan_*) per (lane, input set)
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    lanes = [(det.new_workspace(B_loc, P, dev), det.new_outputs(B_loc, dev)) for _ in range(S)]
    plans = [[det.plan(a[0], a[1], a[2], a[3], priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1])
              for a in dev_sets] for l in range(S)]
    exchanges, exchange_err, xplans = None, None, None
    if dist is not None and not args.no_exchange:
        try:
            exchanges = [rdist.PeerExchange(B_max, C, lanes[l][1].dets.shape[2], dev) for l in range(S)]
            # stage + exchange of its result as ONE plan per (lane, input set): a step stays one driver call
            xplans = [[det.plan(a[0], a[1], a[2], a[3], priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1],
                                then=exchanges[l].exchange) for a in dev_sets] for l in range(S)]
        except Exception as e:                                              # symmetric memory unavailable
            exchanges, xplans, exchange_err = None, None, repr(e)[:200]

    def step(i, with_exchange=True):
        l = i % S
        if xplans is not None and with_exchange:
            return xplans[l][i % NBUF].launch(streams[l])
        return plans[l][i % NBUF].launch(streams[l])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    main_st = torch.cuda.current_stream(dev)

    def timed_regions(K, R, fn):
        """R repetitions of: barrier + synchronize, K steps (S lanes in flight), join, barrier + synchronize.
        Returns the per-region milliseconds (max over ranks) and this rank's own."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(R)]
        n = 0
        for e0, e1 in evs:
            barrier()
            e0.record(main_st)
            for st in streams:
                st.wait_event(e0)
            for _ in range(K):
                fn(n)
                n += 1
            for st in streams:
                main_st.wait_stream(st)
            e1.record(main_st)
            barrier()
        t = torch.tensor([a.elapsed_time(b) for a, b in evs], device=dev, dtype=torch.float64)
        local = t.clone()
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist(), local.tolist()

    W = max(3, args.warmup)
    for i in range(max(W, S * NBUF)):
        step(i)
    torch.cuda.synchronize()
    kept_rows = int(plans[0][0].launch(streams[0]).counts.sum())
    # the replayed plan and the direct launch chain must agree (same kernels, same buffers)
    chk = det.detect(dev_sets[0][0], dev_sets[0][1], dev_sets[0][2], dev_sets[0][3], priors, scale=scale)
    torch.cuda.synchronize()
    if not torch.equal(chk.counts, plans[0][0].result.counts):
        raise RuntimeError('plan replay and direct launch disagree')

    # ---- timed regions: exactly K steps each --------------------------------------------------------
    K = max(1, args.steps)
    R = max(1, args.regions)
    # value = the detect stage (decode + NMS: BASELINE.json's metric), every rank on its own images
    launches0 = _ffi.launch_count()
    with clocks:
        reg_ms, reg_local = timed_regions(K, R, lambda i: step(i, False))
    launches = (_ffi.launch_count() - launches0) // R
    total_ms = median(reg_ms)
    ms_per_step = total_ms / K
    value = B_total * K / (total_ms * 1e-3)
    regions = {'n': R, 'ms_per_step_min': min(reg_ms) / K, 'ms_per_step_median': ms_per_step,
               'ms_per_step_max': max(reg_ms) / K}
    # with_exchange = the same K-step regions with the path's one exchange inside EVERY step: stage + all-gather of the
    # compact detections as one captured plan per lane, the exchange of a batch overlapping the stage of the next
    # batches on the other lanes (N > 1 only)
    with_exchange = None
    if exchanges is not None:
        with clocks:
            x_ms, _ = timed_regions(K, R, lambda i: step(i, True))
        with_exchange = {'value': B_total * K / (median(x_ms) * 1e-3), 'unit': UNIT, 'ms_per_step': median(x_ms) / K,
                         'ms_per_step_min': min(x_ms) / K, 'ms_per_step_max': max(x_ms) / K, 'regions': R,
                         'mode': exchanges[0].mode,
                         'bytes_received_per_gpu_per_step': 20 * kept_rows * (world - 1)}
    elif dist is not None:
        with_exchange = {'error': exchange_err}

    # ---- latency of ONE batch: single stream, L2 flushed (untimed) before every step -------------------
    KL = min(max(K, 20), 50)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(KL)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(KL)]
    with clocks:
        for i in range(KL):
            flush_buf.zero_()
            starts[i].record(main_st)
            plans[0][i % NBUF].launch(main_st)
            stops[i].record(main_st)
        torch.cuda.synchronize()
    lat = sorted(s_.elapsed_time(e_) for s_, e_ in zip(starts, stops))
    latency_ms = lat[KL // 2]

    class _Flush(object):
        def zero_(self):
            flush_buf.zero_()
    # per-kernel breakdown (separate pass: stage events recorded between the kernels)
    kern = det.profile_stage(dev_sets, priors, scale, _Flush(), steps=20)
    stage_ms_serialised = sum(kern.values())
    bytes_per_launch = B_loc * BYTES_PER_IMAGE + 20 * kept_rows
    peak, peak_src = measured_peak()
    # The stage is one launch chain (collect -> graph || nms_small -> nms_large), replayed as one CUDA graph; its
    # kernels overlap (programmatic dependent launch) and so do the chains of the S batches in flight, so the
    # duration that counts is the event-timed region / K.
    local_ms_per_step = median(reg_local) / K
    achieved = bytes_per_launch / (local_ms_per_step * 1e-3) / 1e9
    achieved_single = bytes_per_launch / (latency_ms * 1e-3) / 1e9
    dominant = max(kern, key=kern.get)
    traffic, traffic_file = None, latest_traffic_file()
    if args.workload == 'sparse' and args.scaling == 'weak' and traffic_file:
        try:                               # DRAM bytes per launch from the committed ncu --set full capture
            with open(os.path.join(ROOT, 'profiles', traffic_file)) as f:
                traffic = float(json.load(f)['stage_traffic_bytes'])
        except Exception:
            traffic = None
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic,
                'dram_frac': (traffic / (local_ms_per_step * 1e-3) / 1e9 / peak) if traffic else None,
                'single_batch_frac': achieved_single / peak, 'single_batch_ms': latency_ms, 'dense_frac': None,
                'kernel': 'detect stage launch chain (largest share: %s)' % dominant,
                'bytes_per_launch': bytes_per_launch}
    detail['roofline'] = {
        'traffic_source': 'profiles/%s: dram__bytes_read.sum + dram__bytes_write.sum of the whole launch chain, one ncu '
                          '--set full capture' % traffic_file, 'peak_source': peak_src,
        'stage_ms_serialised': stage_ms_serialised, 'kernels_ms': kern,
        'kernel_share': {k: v / stage_ms_serialised for k, v in kern.items()},
        'single_batch': {'median_ms': latency_ms, 'mean_ms': float(sum(lat)) / KL, 'min_ms': lat[0],
                         'how': 'one stream, 512 MiB memset (L2 flush, untimed) before every step, %d steps' % KL},
        'note': 'achieved = stage bytes (SURVEY 8d: 4*P*(10+C) B/image + 20 B/kept row) / (median event-timed region / K '
                'steps), %d batches in flight; ARM-filtered anchors (%.1f%% here) are never fetched, so DRAM traffic '
                '(dram_frac) is far below the algorithmic bytes: the stage is issue/latency-bound' % (S, 100 * (1 - arm_pass))}

    # e2e: host (pinned) inputs -> kernels read them over PCIe -> pack -> rows land in pinned host memory
    e2e = None
    if not args.no_e2e:
        e2e_steps = min(K, 40)
        full_bytes = sum(t.numel() * t.element_size() for t in host_sets[0])
        pipe = rd.DetectHostPipeline(det, priors, scale, B_loc, lanes=S)

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
        e2e = {'value': B_total * e2e_steps / pipe_s, 'unit': UNIT, 'h2d_bytes_per_step': zc_bytes,
               'd2h_bytes_per_step': d2h, 'steps': e2e_steps, 'ms_per_step': 1e3 * pipe_s / e2e_steps,
               'staged_copy_value': B_total / serial_copy, 'staged_h2d_bytes_per_step': full_bytes}
        detail['e2e'] = {
            'batches_in_flight': S,
            'mode': 'DetectHostPipeline: arm_conf by DMA, the other pinned host inputs read by the kernels over PCIe '
                    '(only rows of ARM-passing anchors cross the bus: h2d_bytes_per_step; the tensors hold %d B), rows '
                    'packed on the device and copied back by one DMA of exactly the kept rows, every result read on the '
                    'host' % full_bytes,
            'serial_zero_copy': {'value': B_total / serial_zc, 'ms_per_step': 1e3 * serial_zc},
            'serial_staged_copy': {'value': B_total / serial_copy, 'ms_per_step': 1e3 * serial_copy,
                                   'h2d_bytes_per_step': full_bytes}}
        del pipe

    def flushed_ms(fn, n=10, warm=3, flush=True):
        """Median device time of fn() with a 512 MiB memset (L2 flush, untimed) before every call.  The memset also keeps
        the GPU busy while the host prepares the launch, so the event pair brackets device time only (measured: with
        rotated inputs and NO memset the same kernels read 12 - 14 us longer -- the host's launch latency)."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evs:
            if flush:
                flush_buf.zero_()
            a.record(main_st)
            fn()
            b.record(main_st)
        torch.cuda.synchronize()
        return median([a.elapsed_time(b) for a, b in evs])

    solo = rank == 0 and world == 1 and not args.no_secondary
    # secondary (reported, not the headline): the other generator of SURVEY.md §8d on the same config
    secondary = None
    if solo:
        other = 'dense' if args.workload == 'sparse' else 'sparse'
        o_dev = [t.to(dev) for t in synthetic.detect_inputs(seed_for(rank, 9), BATCH, P, C, other)]
        holder = {}

        def run_other():
            holder['r'] = det.detect(o_dev[0], o_dev[1], o_dev[2], o_dev[3], priors, scale=scale)
        ms2 = flushed_ms(run_other)
        kept2 = int(holder['r'].counts.sum())
        bytes2 = BATCH * BYTES_PER_IMAGE + 20 * kept2
        frac2 = bytes2 / (ms2 * 1e-3) / 1e9 / peak
        secondary = {'generator': other, 'value': BATCH / (ms2 * 1e-3), 'unit': UNIT, 'ms_per_step': ms2,
                     'roofline_frac': frac2, 'arm_pass_fraction': float((o_dev[1][..., 1] > OBJ_THR).float().mean()),
                     'kept_rows_per_step': kept2}
        if other == 'dense':
            roofline['dense_frac'] = frac2
        detail['secondary_note'] = 'dense = stress case: 86 % of the anchors pass the ARM filter, every class saturates ' \
                                   'top_k = 1000 and goes through the large-problem kernel; one batch, L2 flushed'
        del o_dev, holder

    # BASELINE.json configs 2 and 5 (reported, not the headline): RefineDet320 VOC and the 2-class SAR-ship
    # RefineDet512 detect stage, batch 32, both generators, one batch at a time with the L2 flushed before it
    other_configs = None
    if solo:
        other_configs = {}
        for name, size, dim, C2, nms_thr in (('cfg2_refinedet320_voc', '320', 320.0, 21, 0.45),
                                             ('cfg5_sarship_2class', '512', 512.0, 2, 0.49)):
            pri2 = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().to(dev)
            P2 = pri2.shape[0]
            det2 = rd.Detect_RefineDet(C2, int(dim), 0, TOP_K, CONF_THR, nms_thr, OBJ_THR, KEEP_TOP_K)
            scale2 = torch.tensor([dim] * 4, device=dev).reshape(1, 4).expand(BATCH, 4).contiguous()
            for gen in ('sparse', 'dense'):
                x = [t.to(dev) for t in synthetic.detect_inputs(seed_for(rank, 20), BATCH, P2, C2, gen)]
                holder = {}

                def run2():
                    holder['r'] = det2.detect(x[0], x[1], x[2], x[3], pri2, scale=scale2)
                ms_o = flushed_ms(run2)
                bytes_o = BATCH * 4 * P2 * (10 + C2) + 20 * int(holder['r'].counts.sum())
                other_configs['%s_%s' % (name, gen)] = {
                    'ms_per_step': ms_o, 'value': BATCH / (ms_o * 1e-3),
                    'roofline_frac': bytes_o / (ms_o * 1e-3) / 1e9 / peak}
                del x, holder

    # f-1: logits in (softmax folded into the stage) against torch.softmax + the stage, one stream, L2 flushed
    logits_in = None
    if solo:
        lg = [t.to(dev) for t in synthetic.detect_logits(seed_for(rank, 0), BATCH, P, C, args.workload)]
        plan_l = det.plan(lg[0], lg[1], lg[2], lg[3], priors, scale=scale, workspace=lanes[0][0], out=lanes[0][1],
                          logits=True)
        ms_f = flushed_ms(lambda: plan_l.launch(main_st), n=20)
        ms_u = flushed_ms(lambda: det.detect(lg[0], torch.softmax(lg[1], -1), lg[2], torch.softmax(lg[3], -1), priors,
                                             scale=scale), n=20)
        same = bool(torch.equal(plan_l.launch(main_st).counts, plans[0][0].launch(main_st).counts))
        logits_in = {'ms_per_step': ms_f, 'value': BATCH / (ms_f * 1e-3), 'torch_softmax_then_stage_ms': ms_u,
                     'counts_equal_to_probability_input': same}
        del plan_l, lg

    # a3: Detect_RefineDet.forward as models/refinedet.py:141 calls it (dense boxes + scores out, in-place zeroing)
    a3 = None
    if solo:
        a = dev_sets[0]
        confs = [a[3].clone() for _ in range(6)]
        it = {'i': 0}

        def run3():
            it['i'] += 1
            det.forward(a[0], a[1], a[2], confs[it['i'] % 6], priors)
        run3()
        ms3 = flushed_ms(run3, n=5, warm=0)            # every conf copy is used once: the in-place zeroing is real work
        bytes3 = BATCH * (BYTES_PER_IMAGE + 4 * P * (4 + C))             # SURVEY 8d "a3 contract only"
        a3 = {'ms_per_step': ms3, 'value': BATCH / (ms3 * 1e-3), 'roofline_frac': bytes3 / (ms3 * 1e-3) / 1e9 / peak}
        del confs

    # BASELINE.json config 4 (reported, not the headline): RefineDetMultiBoxLoss training step — ARM + ODM criteria,
    # forward + backward, 50 ground-truth boxes per image, batch 32 (match, conf loss, HNM, reduce, backward kernels)
    train_step = None
    if solo:
        tp = [t.to(dev) for t in synthetic.train_predictions(seed_for(rank, 40), BATCH, P, C)]
        tg = [t.to(dev) for t in synthetic.targets(seed_for(rank, 41), BATCH, 50, C)]
        leaves = [t.clone().requires_grad_(True) for t in tp]

        def make_step(sync_free, paired=True, read=True):
            arm_crit = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True, sync_free=sync_free)
            odm_crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True, sync_free=sync_free)
            pair = rd.RefineDetCriterionPair(arm_crit, odm_crit)

            def one_step():
                rd.box_utils.clear_pad_cache()                  # a real step brings new targets: pad them once, not zero times
                preds = (leaves[0], leaves[1], leaves[2], leaves[3], priors)
                if paired:                                      # one call for both criteria (the two chains on two streams)
                    al, ac, ol, oc = pair(preds, tg)
                else:                                           # the reference's call sequence, train_refinedet.py:252-253
                    al, ac = arm_crit(preds, tg)
                    ol, oc = odm_crit(preds, tg)
                (al + ac + ol + oc).backward()                  # train_refinedet.py:254-256
                vals = torch.stack([al.detach().reshape(()), ac.detach().reshape(()), ol.detach().reshape(()),
                                    oc.detach().reshape(())]).tolist() if (sync_free and read) else None   # :258-261 .item()
                for t in leaves:
                    t.grad = None
                return vals
            return one_step

        def wall_ms(fn, n=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return 1e3 * (time.perf_counter() - t0) / n
        ms_t = wall_ms(make_step(False))
        ms_sf = wall_ms(make_step(True))
        ms_two = wall_ms(make_step(False, paired=False))
        ms_two_sf = wall_ms(make_step(True, paired=False))
        # the kernels of the ODM criterion one by one (device time, L2 flushed) against their algorithmic bytes:
        # refine_match 40 P + 20 G per image, conf loss 4 P C + 17 P, mining 6 P, backward 4 P C + 18 P
        bu = rd.box_utils
        truths_k, labels_k, cnt_k = bu.pad_targets(tg, dev)
        lt_k, ct_k = bu.match_batch(0.5, truths_k, labels_k, cnt_k, priors, [0.1, 0.2], tp[0], bu.LABEL_ODM)
        ce_k, lse_k, pos_k = bu.conf_loss(tp[3], ct_k, tp[1], 0.01)
        neg_k, npos_k = bu.hnm_select(ce_k, pos_k, 3)
        one_k, n_k = torch.ones((), device=dev), pos_k.sum().float()
        kernels = {}
        conf_rot = [tp[3], tp[3].clone()]                       # 2 x 169 MB, alternated
        gconf_rot = [torch.empty_like(tp[3]) for _ in range(2)]
        gloc_k = torch.empty_like(tp[2])
        rot = {'i': 0}

        def conf_loss_rot():
            rot['i'] += 1
            bu.conf_loss(conf_rot[rot['i'] & 1], ct_k, tp[1], 0.01)

        def backward_rot():
            rot['i'] += 1
            with torch.cuda.device(dev):
                _ffi.check(_ffi.lib().rd_multibox_loss_backward(
                    _ffi.ptr(tp[2]), _ffi.ptr(lt_k), _ffi.ptr(conf_rot[rot['i'] & 1]), _ffi.ptr(ct_k), _ffi.ptr(lse_k),
                    _ffi.ptr(pos_k), _ffi.ptr(neg_k), _ffi.ptr(one_k), _ffi.ptr(one_k), _ffi.ptr(n_k), BATCH * P, C,
                    _ffi.ptr(gloc_k), _ffi.ptr(gconf_rot[rot['i'] & 1]), _ffi.stream_ptr()), 'rd_multibox_loss_backward')
        for name, byts, fn, flush in (
                ('refine_match', BATCH * (40 * P + 20 * 50),
                 lambda: bu.match_batch(0.5, truths_k, labels_k, cnt_k, priors, [0.1, 0.2], tp[0], bu.LABEL_ODM), True),
                ('conf_loss', BATCH * P * (4 * C + 17), conf_loss_rot, True),
                ('hnm_select', BATCH * 6 * P, lambda: bu.hnm_select(ce_k, pos_k, 3), True),
                ('loss_reduce', BATCH * P * 6, lambda: bu.multibox_loss_reduce(tp[2], lt_k, ce_k, pos_k, neg_k, npos_k), True),
                ('loss_backward', BATCH * P * (4 * C + 18), backward_rot, True)):
            ms_k = flushed_ms(fn, n=10, flush=flush)
            kernels[name] = {'ms': round(ms_k, 5), 'frac': round(byts / (ms_k * 1e-3) / 1e9 / peak, 4)}
        del conf_rot, gconf_rot, gloc_k
        # device time of the whole step with the host AHEAD of the GPU, as it is inside a training loop (the criterion is
        # queued while the network's forward pass still runs): a spin kernel holds the stream while the steps are queued,
        # nothing is read back in between
        def device_ms(step, k=6):
            step()
            torch.cuda.synchronize()
            ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(int(12e-3 * 1.9e9))              # ~12 ms; the host queues k steps in 3-4 ms
            ev_a.record(main_st)
            for _ in range(k):
                step()
            ev_b.record(main_st)
            torch.cuda.synchronize()
            return ev_a.elapsed_time(ev_b) / k
        dev_pair = device_ms(make_step(True, read=False))
        dev_two = device_ms(make_step(True, paired=False, read=False))
        train_step = {'ms_per_step': ms_t, 'value': BATCH / (ms_t * 1e-3), 'sync_free_ms_per_step': ms_sf,
                      'two_call_ms_per_step': ms_two, 'two_call_sync_free_ms_per_step': ms_two_sf,
                      'device_ms_per_step': dev_pair, 'two_call_device_ms_per_step': dev_two, 'kernels': kernels,
                      'cpu_baseline': None}
        del truths_k, labels_k, cnt_k, lt_k, ct_k, ce_k, lse_k, pos_k, neg_k
        detail['train_step'] = 'ARM + ODM RefineDetMultiBoxLoss forward + backward (B=32, P=16320, C=81, 50 GT/image), wall ' \
                               'clock incl. host glue, through RefineDetCriterionPair (one call for both criteria, N read ' \
                               'once per step); two_call: the two modules one after the other as train_refinedet.py:252-253; ' \
                               'sync_free: no host read of N inside the criteria, losses read once; device_ms: GPU time per step when ' \
                               'the host runs ahead (as inside a training loop), no read-back'
        if not args.no_cpu_baseline:
            try:
                from baseline import reference_arm as ra
                if ra.available():
                    sec, _ = ra.run_train_step(seed_for(rank, 40), seed_for(rank, 41), BATCH, P, C, 50, steps=2, threads=cores)
                    train_step['cpu_baseline'] = {'value': BATCH / sec, 'unit': UNIT, 'ms_per_step': 1e3 * sec, 'cores': cores,
                                                  'kind': 'reference',
                                                  'sample': '2 steps of the same batch: unmodified reference criteria '
                                                            '(refinedet_multibox_loss.py:50-139) forward + backward on the CPU'}
            except Exception as e:
                train_step['cpu_baseline'] = {'error': repr(e)[:160]}
        del leaves, tp, tg

    # the exchange step on its own (N > 1): NCCL form against the fused pack + P2P scatter
    gather = None
    if dist is not None:
        res_g = plans[0][0].launch(main_st)
        torch.cuda.synchronize()

        def timed_gather(fn, n=20):
            for _ in range(3):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(main_st)
            for _ in range(n):
                fn()
            e1.record(main_st)
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / n * 1e3, e0.elapsed_time(e1) / n], device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return [float(v) for v in dt]
        counts_n, rows_n = rdist.gather_detections(res_g)
        gather = {'rows_per_rank': int(rows_n[rank].shape[0]), 'bytes_per_rank': int(rows_n[rank].numel() * 4),
                  'nccl_ms': timed_gather(lambda: rdist.gather_detections(res_g))[0]}
        if exchanges is not None:
            ex = exchanges[0]
            ex.exchange(res_g)
            counts_p, rows_p = ex.result()
            same = all(torch.equal(counts_p[r], counts_n[r]) and torch.equal(rows_p[r], rows_n[r])
                       for r in range(world))
            wall, devms = timed_gather(lambda: ex.exchange(res_g))
            gather.update({'peer_equals_nccl': bool(same), 'peer_device_ms': devms, 'peer_wall_ms': wall,
                           'peer_ms_with_host_read': timed_gather(lambda: (ex.exchange(res_g), ex.result()))[0],
                           'mode': getattr(ex, 'mode', 'p2p')})
        else:
            gather['peer_error'] = exchange_err
        detail['gather'] = 'nccl_ms: pack kernel, header all_gather, host read, padded all_gather_into_tensor ' \
                           '(dist.gather_packed), wall clock; peer_device_ms: rd_pack_scatter (pack + stores into every ' \
                           'peer\'s buffer over NVLink, symmetric memory) between two device barriers, CUDA events; max over ranks'

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, done, elapsed, kind, what = cpu_reference_run(args.workload, 1000, 1, cores, budget_s=12.0)
        cpu_baseline = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
                        'sample': '%d steps x %d images (1 per worker process) of the same workload, %.1f s; %s'
                                  % (done, cores, elapsed, what)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K,
                'warmup': W, 'ms_per_step': ms_per_step, 'higher_is_better': True,
                'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'with_exchange': with_exchange, 'config': config, 'roofline': roofline, 'regions': regions,
                'gpu_launches': int(launches), 'clocks': clocks.summary(),
                'cpu_baseline': cpu_baseline, 'e2e': e2e, 'gather': gather, 'secondary': secondary,
                'train_step': train_step, 'a3_forward': a3, 'logits_in': logits_in, 'other_configs': other_configs,
                'run_info': {'l2': 'inputs larger than L2: %d input sets (1.5 GB) rotated, no flush in the timed region' % NBUF,
                             'batches_in_flight': S, 'launch': 'one CUDA-graph replay per step',
                             'arm_pass_fraction': arm_pass,
                             'kept_rows_per_step': kept_rows, 'batch_this_rank': B_loc, 'numa_bound_cores': numa}}
        print(json.dumps(line))
        sys.stderr.write('BENCH_DETAIL ' + json.dumps(detail) + '\n')
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
