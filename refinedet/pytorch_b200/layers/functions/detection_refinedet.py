"""Drop-in for ``layers/functions/detection_refinedet.py`` (reference :7-113).

``Detect_RefineDet`` keeps the reference's constructor and the two methods
``models/refinedet.py:141`` and the eval scripts call (``forward``,
``forward_python_nms``) and adds the fused detect stage the metric is quoted on
(``detect``: Detect_RefineDet.forward + the per-class loop of
``eval_refinedet_coco.py:205-232`` in four kernel launches, inputs read once).
"""
import torch

from .. import box_utils  # noqa: F401  (kept importable like the reference module)
from ... import _ffi
from ..._ffi import check, lib, on_device, ptr, require_cuda_f32, stream_ptr

# data/config.py:103,115 — the only field of the config this layer reads (:25)
_VARIANCE = {'320': [0.1, 0.2], '512': [0.1, 0.2]}


class Detections(object):
    """Result of the fused detect stage.

    ``counts[B,C]`` int32, ``dets[B,C,max_out,5]`` f32 (only the first ``counts[b,c]``
    rows of a slot are meaningful), ``anchors[B,C,max_out]`` int32 anchor index per row."""

    def __init__(self, counts, dets, anchors, row_layout):
        self.counts, self.dets, self.anchors, self.row_layout = counts, dets, anchors, row_layout

    def packed(self):
        """``(offsets[B*C+1] int32, rows[total,5])`` — one device pass, one host sync (the row count is read first, so
        the packed buffer holds exactly ``total`` rows instead of the ``B*C*max_out`` worst case: 7.7 MB, not 26 MB, for a
        config-3 batch)."""
        B, C, max_out, _ = self.dets.shape
        dev = self.dets.device
        total = int(self.counts.sum().item())
        offsets = torch.empty(B * C + 1, dtype=torch.int32, device=dev)
        rows = torch.empty(max(total, 1), 5, dtype=torch.float32, device=dev)
        with on_device(dev):
            check(lib().rd_pack_detections(ptr(self.counts), ptr(self.dets), B, C, max_out, ptr(offsets),
                                           ptr(rows), rows.shape[0], stream_ptr()), 'rd_pack_detections')
        return offsets, rows[:total]

    def to_coco_arrays(self, class_to_cat_id=None):
        """The records of :meth:`to_coco_results` as two numpy arrays, built on the device (``rd_coco_records``) and
        copied back once: ``ids[n,2]`` int32 = (image index b, class c), ``vals[n,5]`` float64 = (x, y, w, h, score),
        in the reference's order (classes ascending, images ascending inside a class, rows score-descending)."""
        import numpy as np
        B, C, max_out, _ = self.dets.shape
        dev = self.dets.device
        cats = None
        if class_to_cat_id is not None:
            cats = torch.tensor([-1 if (c == 0 or class_to_cat_id[c] is None) else 1 for c in range(C)],
                                dtype=torch.int32).to(dev)
        cap = B * C * max_out
        total = torch.empty(1, dtype=torch.int32, device=dev)
        ids = torch.empty(cap, 2, dtype=torch.int32, device=dev)
        vals = torch.empty(cap, 5, dtype=torch.float64, device=dev)
        with on_device(dev):
            check(lib().rd_coco_records(ptr(self.counts), ptr(self.dets), B, C, max_out, None, ptr(cats), ptr(ids),
                                        ptr(vals), cap, ptr(total), stream_ptr()), 'rd_coco_records')
        n = int(total.item())
        return ids[:n].cpu().numpy(), vals[:n].cpu().numpy()

    def to_coco_numpy(self, image_ids, class_to_cat_id):
        """The results as ONE ``[n,7]`` float64 array ``(image_id, x, y, w, h, score, category_id)`` — the layout
        ``pycocotools``' ``COCO.loadRes`` accepts directly (``loadNumpyAnnotations``), so the evaluation needs neither
        the list of dicts nor the json file of data/sarship_coco.py:293-336.  No Python loop over rows.
        (``image_ids`` / category ids must be numeric here.)"""
        import numpy as np
        ids, vals = self.to_coco_arrays(class_to_cat_id)
        out = np.empty((ids.shape[0], 7), np.float64)
        out[:, 0] = np.asarray(image_ids, np.float64)[ids[:, 0]]
        out[:, 1:6] = vals
        cats = np.asarray([-1 if c is None else c for c in class_to_cat_id], np.float64)
        out[:, 6] = cats[ids[:, 1]]
        return out

    def to_coco_results(self, image_ids, class_to_cat_id):
        """The result wire format of the reference's eval (``data/sarship_coco.py:293-336``): a list of
        ``{'image_id', 'category_id', 'bbox': [x, y, w, h], 'score'}`` with ``w = x2 - x1 + 1``,
        ``h = y2 - y1 + 1`` computed in float64 (the reference's ``astype(np.float)``), classes ascending
        (background skipped), images ascending inside a class, rows score-descending.
        ``image_ids[b]`` is the dataset index of image b, ``class_to_cat_id[c]`` the COCO category of
        class c (``None`` skips the class).  The records come from the device (:meth:`to_coco_arrays`); the host
        maps the two indices and wraps the rows — the dicts themselves are what takes the time at 4 x 10^5 rows;
        :meth:`to_coco_numpy` avoids them."""
        ids, vals = self.to_coco_arrays(class_to_cat_id)
        img = [image_ids[b] for b in ids[:, 0].tolist()]
        cat = [class_to_cat_id[c] for c in ids[:, 1].tolist()]
        return [{'image_id': i, 'category_id': c, 'bbox': bb, 'score': sc}
                for i, c, bb, sc in zip(img, cat, vals[:, :4].tolist(), vals[:, 4].tolist())]

    def to_all_boxes(self):
        """``all_boxes[c][b]`` numpy arrays ``[n,5]`` as built by eval_refinedet_coco.py:214-232."""
        counts = self.counts.cpu().numpy()
        dets = self.dets.cpu().numpy()
        B, C = counts.shape
        return [[dets[b, c, :counts[b, c]].copy() for b in range(B)] for c in range(C)]


class DetectPlan(object):
    """The fused detect stage for fixed buffers as one CUDA-graph replay (``rd_detect_plan_*``).

    ``launch(stream=None)`` replays asynchronously on ``stream`` (default: the current stream) and
    returns the plan's :class:`Detections` (persistent buffers, overwritten by every replay)."""

    def __init__(self, args, result, device, keep_alive):
        import ctypes
        self.result, self.device, self._keep = result, device, keep_alive
        handle = ctypes.c_void_p(0)
        with on_device(device):
            check(lib().rd_detect_plan_create(*args, ctypes.byref(handle)), 'rd_detect_plan_create')
        self._handle = handle

    @classmethod
    def capture(cls, device, body, result=None, keep_alive=None):
        """A plan from whatever ``body()`` enqueues on the current stream (``rd_plan_capture_begin/end``): the
        stage's launch chain followed by the exchange of its result, for example."""
        import ctypes
        self = cls.__new__(cls)
        self.result, self.device, self._keep = result, device, keep_alive
        handle = ctypes.c_void_p(0)
        st = torch.cuda.Stream(device)
        torch.cuda.synchronize(device)
        with torch.cuda.stream(st):
            check(lib().rd_plan_capture_begin(st.cuda_stream), 'rd_plan_capture_begin')
            try:
                body()
            finally:
                rc = lib().rd_plan_capture_end(st.cuda_stream, ctypes.byref(handle))
            check(rc, 'rd_plan_capture_end')
        self._handle = handle
        return self

    def launch(self, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(lib().rd_detect_plan_launch(self._handle, s.cuda_stream), 'rd_detect_plan_launch')
        return self.result

    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle.value:
            lib().rd_detect_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DetectHostPipeline(object):
    """End-to-end detect stage for HOST inputs with several batches in flight.

    Each lane owns a stream, a workspace, device output slots and pinned host result buffers.
    ``submit(host_inputs)`` (pinned ``arm_loc, arm_conf, odm_loc, odm_conf``) copies ``arm_conf`` by DMA,
    replays the lane's plan — the kernels read the other pinned tensors over PCIe, so only the rows of
    ARM-passing anchors cross the bus — packs the rows and returns a ticket at once.
    ``result(ticket)`` waits for that batch only and returns CPU tensors ``(counts[B,C], rows[total,5])``
    (views of the lane's buffers, valid until the lane is reused ``lanes`` submits later).
    ``dma_rows=True`` (default): rows are packed on the device and ``result`` issues one DMA of exactly
    ``total`` rows; ``False``: the pack kernel stores them straight into pinned host memory (one wait less,
    measured ~12 % slower at four lanes).  With several lanes the PCIe reads of batch i+1 overlap the
    kernels and the write-back of batch i."""

    def __init__(self, det, prior_data, scale, B, lanes=2, dma_rows=True):
        self.det, self.priors, self.B = det, require_cuda_f32(prior_data, 'prior_data'), B
        self.dma_rows = dma_rows      # False: pack kernel stores rows into pinned host memory; True: pack on device + DMA
        dev = self.priors.device
        self.device = dev
        C = det.num_classes
        max_out = max(1, min(int(det.keep_top_k), int(det.top_k)))
        self.scale = None
        if scale is not None:
            sc = torch.as_tensor(scale, dtype=torch.float32).to(dev)
            self.scale = sc.reshape(1, 4).expand(B, 4).contiguous() if sc.numel() == 4 else sc.reshape(B, 4).contiguous()
        self.lanes = []
        for _ in range(lanes):
            self.lanes.append({
                'stream': torch.cuda.Stream(dev),
                'ws': det.new_workspace(B, self.priors.shape[0], dev),
                'out': det.new_outputs(B, dev),
                'dev_offsets': torch.empty(B * C + 1, dtype=torch.int32, device=dev),
                'host_offsets': torch.empty(B * C + 1, dtype=torch.int32).pin_memory(),
                'host_counts': torch.empty(B, C, dtype=torch.int32).pin_memory(),
                'host_rows': torch.empty(B * C * max_out, 5, dtype=torch.float32).pin_memory(),
                'dev_rows': torch.empty(B * C * max_out, 5, dtype=torch.float32, device=dev) if dma_rows else None,
                'done': torch.cuda.Event(),
                'arm_conf': torch.empty(B, self.priors.shape[0], 2, dtype=torch.float32, device=dev),
                'plans': {},
            })
        torch.cuda.synchronize(dev)
        self._next = 0

    def submit(self, host_inputs):
        lane = self.lanes[self._next % len(self.lanes)]
        self._next += 1
        key = tuple(t.data_ptr() for t in host_inputs)
        plan = lane['plans'].get(key)
        if plan is None:
            for t in host_inputs:
                if t.is_cuda or not t.is_pinned() or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError('DetectHostPipeline needs contiguous pinned float32 host tensors')
            args, res, dev, keep = self.det._prepare(
                host_inputs[0], lane['arm_conf'], host_inputs[2], host_inputs[3], self.priors, self.scale,
                _ffi.RD_NMS_PIXEL_PLUS1, _ffi.RD_ROW_BOX_SCORE, self.det.keep_top_k, host_mapped=True,
                workspace=lane['ws'], out=lane['out'])
            torch.cuda.synchronize(self.device)
            plan = lane['plans'][key] = DetectPlan(args, res, dev, keep)
        st = lane['stream']
        with torch.cuda.stream(st):      # arm_conf (8 B / anchor, read by every CTA of an image) goes by DMA
            lane['arm_conf'].copy_(host_inputs[1].view_as(lane['arm_conf']), non_blocking=True)
        res = plan.launch(st)
        B, C, max_out, _ = res.dets.shape
        with torch.cuda.device(self.device), torch.cuda.stream(st):
            rows_dst = lane['dev_rows'] if self.dma_rows else lane['host_rows']
            check(lib().rd_pack_detections(ptr(res.counts), ptr(res.dets), B, C, max_out, ptr(lane['dev_offsets']),
                                           ptr(rows_dst), rows_dst.shape[0], st.cuda_stream),
                  'rd_pack_detections')
            lane['host_counts'].copy_(res.counts, non_blocking=True)
            lane['host_offsets'].copy_(lane['dev_offsets'], non_blocking=True)
            lane['done'].record(st)
        return lane

    def result(self, ticket):
        ticket['done'].synchronize()
        total = int(ticket['host_offsets'][-1])
        if self.dma_rows:       # rows packed on the device: one DMA of exactly `total` rows, then a second wait
            st = ticket['stream']
            with torch.cuda.stream(st):
                ticket['host_rows'][:total].copy_(ticket['dev_rows'][:total], non_blocking=True)
            st.synchronize()
        return ticket['host_counts'], ticket['host_rows'][:total]


class Detect_RefineDet(object):
    """At test time, the final layer of RefineDet: ARM-objectness filter, two-stage decode
    and (in ``detect`` / ``forward_python_nms``) per-class threshold, top-k and NMS.

    Constructor arguments are the reference's (detection_refinedet.py:13-25)."""

    def __init__(self, num_classes, size, bkg_label, top_k, conf_thresh, nms_thresh,
                 objectness_thre, keep_top_k):
        self.num_classes = num_classes
        self.background_label = bkg_label
        self.top_k = top_k
        self.keep_top_k = keep_top_k
        self.nms_thresh = nms_thresh
        if nms_thresh <= 0:
            raise ValueError('nms_threshold must be non negative.')
        self.conf_thresh = conf_thresh
        self.objectness_thre = objectness_thre
        self.variance = _VARIANCE[str(size)]
        self._ws = None
        self._ws_key = None

    # -- helpers ---------------------------------------------------------------------------
    def _inputs(self, arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data):
        arm_loc = require_cuda_f32(arm_loc_data, 'arm_loc_data')
        arm_conf = require_cuda_f32(arm_conf_data, 'arm_conf_data', align=8)
        odm_loc = require_cuda_f32(odm_loc_data, 'odm_loc_data')
        odm_conf = require_cuda_f32(odm_conf_data, 'odm_conf_data')
        priors = require_cuda_f32(prior_data, 'prior_data')
        B = odm_loc.shape[0]
        P = priors.shape[0]
        C = self.num_classes
        if tuple(arm_loc.shape) != (B, P, 4) or tuple(odm_loc.shape) != (B, P, 4):
            raise ValueError('loc tensors must be [B,P,4] with P = prior_data.size(0)')
        if arm_conf.numel() != B * P * 2 or odm_conf.numel() != B * P * C:
            raise ValueError('conf tensors must hold [B,P,2] and [B,P,num_classes] values')
        return arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C

    def _workspace(self, B, P, C, device):
        # keyed by the stream too: the control block must be zero at call start and is re-zeroed by the call's
        # own kernels, which orders calls on ONE stream only.  A call on another stream gets a fresh workspace
        # (allocated and reset on that stream) instead of racing on this one.
        key = (B, P, C, device, stream_ptr(device).value)
        if self._ws_key != key:
            L = lib()
            nbytes = int(L.rd_detect_workspace_bytes(B, P, C))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            with on_device(device):
                check(L.rd_detect_workspace_reset(ptr(ws), nbytes, stream_ptr()), 'rd_detect_workspace_reset')
            self._ws, self._ws_key = ws, key
        return self._ws

    # -- reference API ---------------------------------------------------------------------
    def forward(self, arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data):
        """detection_refinedet.py:27-65.  Returns ``(boxes[B,P,4], scores[B,P,C])`` and, like
        the reference (:40-42), zeroes the ARM-filtered rows of the caller's ``odm_conf_data``
        in place."""
        arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C = self._inputs(
            arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data)
        dev = odm_loc.device
        boxes = torch.empty(B, P, 4, dtype=torch.float32, device=dev)
        scores = torch.empty(B, P, C, dtype=torch.float32, device=dev)
        with on_device(dev):
            check(lib().rd_detect_forward(ptr(arm_loc), ptr(arm_conf), ptr(odm_loc), ptr(odm_conf), ptr(priors),
                                          B, P, C, float(self.objectness_thre), float(self.variance[0]),
                                          float(self.variance[1]), ptr(boxes), ptr(scores), stream_ptr()),
                  'rd_detect_forward')
        if odm_conf.data_ptr() != odm_conf_data.data_ptr():      # a copy was made: mirror the in-place write
            odm_conf_data.detach().copy_(odm_conf.view(odm_conf_data.shape))
        self.boxes, self.scores = boxes, scores                  # the reference keeps them on the instance
        return self.boxes, self.scores

    _INSTANCES = {None: 0, 'auto': 0, 256: _ffi.RD_DEBUG_INSTANCE_256, 1024: _ffi.RD_DEBUG_INSTANCE_1024}

    def detect(self, arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data, scale=None,
               force_cpu_semantics=False, logits=False, instance=None):
        """Fused detect stage as evaluated (eval_refinedet_coco.py:205-232), whole batch:
        ARM filter, two-stage decode, ``boxes *= scale``, per class ``score > conf_thresh``,
        top ``top_k``, pixel(+1) NMS at ``nms_thresh``, first ``keep_top_k`` rows per class.

        ``scale``: None, a 4-vector, or ``[B,4]`` (x,y,x,y image size).  The inputs are NOT
        modified.  ``logits=True``: ``arm_conf_data`` / ``odm_conf_data`` are the head outputs BEFORE
        ``models/refinedet.py:143-147``'s softmax, which is folded into the stage (the extra read+write
        pass over ``odm_conf`` disappears).  Returns a :class:`Detections` with rows ``x1,y1,x2,y2,score``."""
        flags = _ffi.RD_NMS_PIXEL_PLUS1 | (_ffi.RD_NMS_SUPPRESS_EQ if force_cpu_semantics else 0) | \
            (_ffi.RD_INPUT_LOGITS if logits else 0)
        return self._fused(arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data, scale, flags,
                           _ffi.RD_ROW_BOX_SCORE, self.keep_top_k)

    def detect_host(self, host_inputs, prior_data, scale=None, staging=None, zero_copy=True):
        """End-to-end form of :meth:`detect` for HOST inputs (pinned tensors
        ``arm_loc, arm_conf, odm_loc, odm_conf``).  Returns CPU tensors ``(counts[B,C], rows[total,5])``.

        ``zero_copy=True`` (default): the kernels read the pinned host buffers directly over PCIe
        (unified addressing) — ``arm_conf`` is streamed once and only the loc / odm_conf rows of
        ARM-passing anchors ever cross the bus, instead of copying every tensor in full.
        ``zero_copy=False``: asynchronous H2D copies into ``staging`` device buffers first.
        Either way the stage, device-side packing and the D2H of counts + packed rows follow."""
        dev = prior_data.device
        if zero_copy:
            for t in host_inputs:
                if t.is_cuda or not t.is_pinned() or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError('detect_host(zero_copy=True) needs contiguous pinned float32 host tensors')
            # arm_conf (8 B / anchor) is read by every CTA of an image: DMA it, the rest stays host-mapped
            if getattr(self, '_arm_stage', None) is None or self._arm_stage.shape != host_inputs[1].shape:
                self._arm_stage = torch.empty_like(host_inputs[1], device=dev)
            self._arm_stage.copy_(host_inputs[1], non_blocking=True)
            res = self._fused(host_inputs[0], self._arm_stage, host_inputs[2], host_inputs[3], prior_data, scale,
                              _ffi.RD_NMS_PIXEL_PLUS1, _ffi.RD_ROW_BOX_SCORE, self.keep_top_k, host_mapped=True)
        else:
            if staging is None:
                staging = [torch.empty_like(t, device=dev) for t in host_inputs]
            for d, h in zip(staging, host_inputs):
                d.copy_(h, non_blocking=True)
            res = self.detect(staging[0], staging[1], staging[2], staging[3], prior_data, scale=scale)
        return self._to_host(res)

    def _to_host(self, res):
        """counts + packed rows of ``res`` as CPU tensors with ONE stream synchronisation: the pack kernels
        write offsets and rows straight into persistent pinned host buffers (zero-copy stores over PCIe),
        counts follow with an async copy.  The returned tensors are views of those buffers, valid until
        the next call."""
        B, C, max_out, _ = res.dets.shape
        dev = res.dets.device
        key = (B, C, max_out)
        if getattr(self, '_host_key', None) != key:
            self._host_offsets = torch.empty(B * C + 1, dtype=torch.int32).pin_memory()
            self._host_rows = torch.empty(B * C * max_out, 5, dtype=torch.float32).pin_memory()
            self._host_counts = torch.empty(B, C, dtype=torch.int32).pin_memory()
            self._dev_offsets = torch.empty(B * C + 1, dtype=torch.int32, device=dev)
            self._host_key = key
        with on_device(dev):
            check(lib().rd_pack_detections(ptr(res.counts), ptr(res.dets), B, C, max_out, ptr(self._dev_offsets),
                                           ptr(self._host_rows), self._host_rows.shape[0], stream_ptr()),
                  'rd_pack_detections')
            self._host_counts.copy_(res.counts, non_blocking=True)
            self._host_offsets.copy_(self._dev_offsets, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        total = int(self._host_offsets[-1])
        return self._host_counts, self._host_rows[:total]

    def profile_stage(self, input_sets, prior_data, scale, flush=None, steps=10):
        """Device time (ms, mean over ``steps``) of each kernel of the fused stage, measured with
        CUDA events recorded between the launches (``rd_detect_fused_timed``)."""
        import ctypes
        names = ('collect_kernel', 'graph_kernel', 'nms_small_large_kernels')
        acc = [0.0, 0.0, 0.0]
        ms = (ctypes.c_float * 4)()
        for i in range(steps):
            if flush is not None:
                flush.zero_()
            a = input_sets[i % len(input_sets)]
            self._fused(a[0], a[1], a[2], a[3], prior_data, scale, _ffi.RD_NMS_PIXEL_PLUS1, _ffi.RD_ROW_BOX_SCORE,
                        self.keep_top_k, timed=ms)
            for k in range(3):
                acc[k] += float(ms[k])
        return {n: v / steps for n, v in zip(names, acc)}

    def _prepare(self, arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data, scale, flags,
                 row_layout, max_out, dets=None, host_mapped=False, workspace=None, out=None):
        """Validated argument tuple of ``rd_detect_fused`` (everything but the stream) + the result object."""
        if host_mapped:      # pinned host tensors, dereferenced by the kernels through unified addressing
            arm_loc, arm_conf, odm_loc, odm_conf = arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data
            priors = require_cuda_f32(prior_data, 'prior_data')
            B, P, C = odm_loc.shape[0], priors.shape[0], self.num_classes
            if tuple(arm_loc.shape) != (B, P, 4) or arm_conf.numel() != B * P * 2 or odm_conf.numel() != B * P * C:
                raise ValueError('host tensors must be [B,P,4], [B,P,2], [B,P,4], [B,P,C]')
            dev = priors.device
        else:
            arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C = self._inputs(
                arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data)
            dev = odm_loc.device
        if self.top_k > _ffi.RD_MAX_NMS_BOXES:
            raise RuntimeError('top_k = %d exceeds the supported %d' % (self.top_k, _ffi.RD_MAX_NMS_BOXES))
        max_out = max(1, min(int(max_out), int(self.top_k)))
        if scale is not None:
            scale = torch.as_tensor(scale, dtype=torch.float32).to(dev)
            scale = scale.reshape(1, 4).expand(B, 4).contiguous() if scale.numel() == 4 else scale.reshape(B, 4).contiguous()
        ws = workspace if workspace is not None else self._workspace(B, P, C, dev)
        if out is not None:
            counts, dets, anchors = out.counts, out.dets, out.anchors
            if tuple(counts.shape) != (B, C) or tuple(dets.shape) != (B, C, max_out, 5):
                raise ValueError('out= buffers must be counts[B,C], dets[B,C,max_out,5], anchors[B,C,max_out]')
        else:
            counts = torch.empty(B, C, dtype=torch.int32, device=dev)
            if dets is None:
                dets = torch.empty(B, C, max_out, 5, dtype=torch.float32, device=dev)
            anchors = torch.empty(B, C, max_out, dtype=torch.int32, device=dev)
        args = (ptr(arm_loc), ptr(arm_conf), ptr(odm_loc), ptr(odm_conf), ptr(priors),
                B, P, C, float(self.objectness_thre), float(self.conf_thresh),
                float(self.nms_thresh), int(self.top_k), max_out, ptr(scale), int(flags),
                int(row_layout), float(self.variance[0]), float(self.variance[1]),
                ptr(ws), ws.numel(), ptr(counts), ptr(dets), ptr(anchors))
        keep_alive = (arm_loc, arm_conf, odm_loc, odm_conf, priors, scale, ws)
        return args, Detections(counts, dets, anchors, row_layout), dev, keep_alive

    def _fused(self, arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data, scale, flags,
               row_layout, max_out, dets=None, timed=None, host_mapped=False):
        args, res, dev, _ = self._prepare(arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data,
                                          scale, flags, row_layout, max_out, dets=dets, host_mapped=host_mapped)
        try:
            with on_device(dev):
                if timed is None:
                    check(lib().rd_detect_fused(*args, stream_ptr()), 'rd_detect_fused')
                else:
                    import ctypes
                    check(lib().rd_detect_fused_timed(*args, stream_ptr(), ctypes.cast(timed, ctypes.c_void_p)),
                          'rd_detect_fused_timed')
        except RuntimeError:
            # a launch of the chain failed after collect_kernel may have marked the control block: the cached
            # workspace is no longer known to be clean, so it is dropped (the next call allocates + resets one)
            self._ws, self._ws_key = None, None
            raise
        return res

    # -- plans: the stage for fixed buffers as one CUDA-graph replay ---------------------------
    def new_workspace(self, B, P, device):
        """A private workspace (one per batch in flight; :meth:`plan` ``workspace=``)."""
        L = lib()
        nbytes = int(L.rd_detect_workspace_bytes(B, P, self.num_classes))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with on_device(device):
            check(L.rd_detect_workspace_reset(ptr(ws), nbytes, stream_ptr()), 'rd_detect_workspace_reset')
        return ws

    def new_outputs(self, B, device, max_out=None, row_layout=_ffi.RD_ROW_BOX_SCORE):
        """Persistent output buffers for :meth:`plan` ``out=``."""
        max_out = max(1, min(int(self.keep_top_k if max_out is None else max_out), int(self.top_k)))
        C = self.num_classes
        return Detections(torch.empty(B, C, dtype=torch.int32, device=device),
                          torch.empty(B, C, max_out, 5, dtype=torch.float32, device=device),
                          torch.empty(B, C, max_out, dtype=torch.int32, device=device), row_layout)

    def plan(self, arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data, scale=None,
             workspace=None, out=None, force_cpu_semantics=False, logits=False, instance=None, then=None):
        """:meth:`detect` for FIXED input buffers, captured once (``rd_detect_plan_create``): the returned
        :class:`DetectPlan` replays the whole launch chain with one driver call per batch.  The tensors are
        referenced, not copied — refill them in place between replays.  Plans that share ``workspace`` /
        ``out`` must be replayed on the same stream; give every batch in flight its own pair
        (:meth:`new_workspace`, :meth:`new_outputs`).  ``then(result)``: a callable that enqueues the consumer of
        the result on the current stream (e.g. ``PeerExchange.exchange``) — captured into the same plan, so that a
        step of stage + exchange is still one driver call."""
        flags = _ffi.RD_NMS_PIXEL_PLUS1 | (_ffi.RD_NMS_SUPPRESS_EQ if force_cpu_semantics else 0) | \
            (_ffi.RD_INPUT_LOGITS if logits else 0) | self._INSTANCES[instance]
        for name, t in (('arm_loc_data', arm_loc_data), ('arm_conf_data', arm_conf_data),
                        ('odm_loc_data', odm_loc_data), ('odm_conf_data', odm_conf_data)):
            if not t.is_contiguous() or t.data_ptr() % 16:
                raise ValueError('plan(): %s must be contiguous and 16-byte aligned (it is referenced, not copied)' % name)
        args, res, dev, keep = self._prepare(arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data,
                                             scale, flags, _ffi.RD_ROW_BOX_SCORE, self.keep_top_k,
                                             workspace=workspace, out=out)
        torch.cuda.synchronize(dev)      # the workspace reset / pending writers are done before the capture
        if then is None:
            return DetectPlan(args, res, dev, keep)

        def body():                      # the launch chain, then whatever consumes the result on the same stream
            check(lib().rd_detect_fused(*args, stream_ptr()), 'rd_detect_fused')
            then(res)
        return DetectPlan.capture(dev, body, res, (keep, then))

    def forward_python_nms(self, arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data):
        """detection_refinedet.py:67-113.  Returns ``output[B,C,top_k,5]`` rows
        ``(score,x1,y1,x2,y2)``, zero padded, class 0 empty; NMS on normalised boxes without
        the +1 convention (box_utils.nms).  Like the reference it zeroes the ARM-filtered rows
        of ``odm_conf_data`` in place; the reference's cross-class ``keep_top_k`` step
        (:109-112) fills a temporary and has no effect, so it is not performed."""
        B = odm_loc_data.shape[0]
        output = torch.zeros(B, self.num_classes, self.top_k, 5, dtype=torch.float32,
                             device=odm_loc_data.device)
        res = self._fused(arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, prior_data, None,
                          _ffi.RD_NMS_NORMALISED, _ffi.RD_ROW_SCORE_BOX, self.top_k, dets=output)
        self.last_detections = res
        # in-place ARM zeroing of the caller's tensor (:79-81)
        arm_c = require_cuda_f32(arm_conf_data, 'arm_conf_data', align=8)
        conf = odm_conf_data.detach()
        conf_c = conf if conf.is_contiguous() else conf.contiguous()
        with on_device(conf_c.device):
            check(lib().rd_arm_zero_rows(ptr(arm_c), ptr(conf_c), arm_c.numel() // 2, self.num_classes,
                                         float(self.objectness_thre), stream_ptr()), 'rd_arm_zero_rows')
        if conf_c.data_ptr() != conf.data_ptr():                 # a copy was made: mirror the in-place write
            conf.copy_(conf_c.view(conf.shape))
        return output

    __call__ = forward
