"""Multi-GPU plumbing for the detect stage: one process per GPU, images sharded by rank.

The reference has no distributed code beyond ``nn.DataParallel`` (``train_refinedet.py:138-139``),
whose gather funnels every head output to GPU 0.  Here every rank post-processes its own images
and the only exchange is ONE gather of the compact detection buffer:
``counts[B_loc, C]`` int32 plus the packed rows ``[total, 5]`` padded to the largest per-rank
total — never the dense ``[B_loc, C, keep_top_k, 5]`` layout (26 MB/rank at config 3).

The functions are device agnostic (NCCL on CUDA tensors, gloo on CPU tensors), so the host logic
is covered by world_size-2 gloo tests on a machine without a GPU.
"""
import torch
import torch.distributed as dist


def shard_range(num_images, rank, world_size):
    """Images ``[lo, hi)`` owned by ``rank``: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(num_images), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_packed(counts, rows, group=None):
    """All-gather the compact detections of every rank.

    ``counts[B_loc, C]`` int32 and ``rows[total, 5]`` float32 of this rank (``B_loc`` may differ
    between ranks by one).  Returns ``(counts_all, rows_all)``: lists with one tensor per rank,
    rank-major = image-major order, rows trimmed to each rank's true total.  Two collectives:
    the per-rank (B_loc, total) header, then one padded all_gather of counts + rows.
    """
    if not dist.is_available() or not dist.is_initialized():
        return [counts], [rows]
    world = dist.get_world_size(group)
    dev = counts.device
    C = counts.shape[1]
    header = torch.tensor([counts.shape[0], rows.shape[0]], dtype=torch.int64, device=dev)
    headers = [torch.empty_like(header) for _ in range(world)]
    dist.all_gather(headers, header, group=group)
    headers = torch.stack(headers).cpu()
    max_b, max_rows = int(headers[:, 0].max()), int(headers[:, 1].max())
    # one buffer per rank: [max_b * C] counts (as float32 bit patterns) followed by [max_rows * 5] rows
    buf = torch.zeros(max_b * C + max_rows * 5, dtype=torch.float32, device=dev)
    buf[:counts.numel()] = counts.reshape(-1).contiguous().view(torch.float32)
    buf[max_b * C:max_b * C + rows.numel()] = rows.reshape(-1)
    gathered = torch.empty(world * buf.numel(), dtype=torch.float32, device=dev)
    if hasattr(dist, 'all_gather_into_tensor') and dev.type == 'cuda':
        dist.all_gather_into_tensor(gathered, buf, group=group)
        parts = gathered.view(world, -1)
    else:
        lst = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(lst, buf, group=group)
        parts = torch.stack(lst)
    counts_all, rows_all = [], []
    for r in range(world):
        b_r, n_r = int(headers[r, 0]), int(headers[r, 1])
        counts_all.append(parts[r, :b_r * C].view(torch.int32).view(b_r, C))
        rows_all.append(parts[r, max_b * C:max_b * C + n_r * 5].view(n_r, 5))
    return counts_all, rows_all


def gather_detections(detections, group=None):
    """``Detections`` of this rank (CUDA) -> ``(counts_all, rows_all)`` of every rank."""
    _, rows = detections.packed()
    return gather_packed(detections.counts, rows, group)


def gathered_to_coco_arrays(counts_all, rows_all, class_to_cat_id=None):
    """COCO result records (``rd_coco_records``, see ``Detections.to_coco_arrays``) of the GATHERED detections of
    every rank — ``(counts_all, rows_all)`` as returned by :func:`gather_packed` / ``PeerExchange.result`` — built on
    the device from the packed rows: ``ids[n,2]`` int32 = (global image index, class), ``vals[n,5]`` float64."""
    from ._ffi import check, lib, on_device, ptr, stream_ptr
    counts = torch.cat([c.reshape(-1, c.shape[-1]) for c in counts_all]).contiguous()
    rows = torch.cat([r.reshape(-1, 5) for r in rows_all]).contiguous()
    if not counts.is_cuda:
        raise RuntimeError('gathered_to_coco_arrays needs CUDA tensors (refinedet.pytorch_b200 has no CPU fallback)')
    B, C = counts.shape
    dev = counts.device
    offsets = torch.zeros(B * C + 1, dtype=torch.int32, device=dev)
    offsets[1:] = counts.reshape(-1).cumsum(0)
    cats = None
    if class_to_cat_id is not None:
        cats = torch.tensor([-1 if (c == 0 or class_to_cat_id[c] is None) else 1 for c in range(C)],
                            dtype=torch.int32).to(dev)
    cap = rows.shape[0]
    total = torch.empty(1, dtype=torch.int32, device=dev)
    ids = torch.empty(max(cap, 1), 2, dtype=torch.int32, device=dev)
    vals = torch.empty(max(cap, 1), 5, dtype=torch.float64, device=dev)
    with on_device(dev):
        check(lib().rd_coco_records(ptr(counts), ptr(rows), B, C, 0, ptr(offsets), ptr(cats), ptr(ids), ptr(vals), cap,
                                    ptr(total), stream_ptr()), 'rd_coco_records')
    n = int(total.item())
    return ids[:n].cpu().numpy(), vals[:n].cpu().numpy()


# ---------------------------------------------------------------------------------------------
# packing fused with the gather: P2P stores into every peer's exchange buffer (rd_pack_scatter)
# ---------------------------------------------------------------------------------------------
def decode_slots(buf, world, slot_bytes, nbc_capacity):
    """Views into an exchange buffer (``world`` slots of ``slot_bytes`` bytes, the layout of
    ``rd_pack_scatter``): ``(counts_all, rows_all)``, one tensor per source rank.  One host sync (the
    headers)."""
    slots = buf.view(world, slot_bytes)
    headers = slots[:, :16].contiguous().view(torch.int32).cpu()          # rows stored, B, C, rows total
    rows_off = 256 + (nbc_capacity * 4 + 255) // 256 * 256
    counts_all, rows_all = [], []
    for r in range(world):
        stored, b_r, c_r, total = (int(v) for v in headers[r])
        if stored != total:
            raise RuntimeError('exchange slot of rank %d holds %d of %d rows: capacity too small' % (r, stored, total))
        counts_all.append(slots[r, 256:256 + b_r * c_r * 4].view(torch.int32).view(b_r, c_r))
        rows_all.append(slots[r, rows_off:rows_off + stored * 20].view(torch.float32).view(stored, 5))
    return counts_all, rows_all


class PeerExchange(object):
    """All-gather of the compact detections WITHOUT a collective call and without barrier kernels
    (``rd_exchange_round``): every rank packs its counts and rows into its own slot of its own buffer, streams the
    slot's used prefix into the same slot of every peer's buffer — symmetric memory: each buffer is mapped into every
    process of the node, the stores travel over NVLink / NVSwitch — and the copy kernel itself publishes the round
    number on every rank and waits for everybody's: ONE cross-GPU rendezvous per round, two launches, replayable
    from a CUDA graph.  With NVSwitch multicast (``mode == 'multicast'``) one ``multimem.st`` is replicated to all
    ranks by the switch, so the rows leave the GPU once instead of ``world - 1`` times; otherwise (``mode == 'p2p'``)
    one group of CTAs stores to each peer.  Slots are double-buffered by the parity of the round.
    ``B`` = the largest per-rank batch; all ranks construct it collectively and call :meth:`exchange` in lockstep.

    Needs ``torch.distributed._symmetric_memory`` (one node, NVLink/PCIe P2P); raises otherwise — callers
    that want a portable path use :func:`gather_packed` (NCCL / gloo)."""

    def __init__(self, B, C, max_out, device, group=None, capacity_rows=None, mode=None, copy_ctas=0, timeout_ms=10000):
        import ctypes
        import os
        import torch.distributed._symmetric_memory as symm_mem
        from ._ffi import lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.B, self.C, self.max_out = int(B), int(C), int(max_out)
        self.capacity = int(capacity_rows) if capacity_rows is not None else self.B * self.C * self.max_out
        self.slot_bytes = int(lib().rd_exchange_slot_bytes(self.B, self.C, self.capacity))
        self.ctrl_bytes = int(lib().rd_exchange_ctrl_bytes())
        self.buf = symm_mem.empty(self.ctrl_bytes + 2 * self.world * self.slot_bytes, dtype=torch.uint8, device=device)
        self.buf.zero_()                                         # control block: epoch 0, flags 0
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        off = int(getattr(self.hdl, 'offset', 0) or 0)
        self._bases = (ctypes.c_void_p * self.world)(*[int(p) + off for p in self.hdl.buffer_ptrs])
        # 'p2p' (default): unicast, the rows are moved by the TMA engines; 'multicast': one multimem.st per 16 bytes,
        # replicated by the NVSwitch.  Measured on 8 B200: both are bound by the 54 MB every GPU RECEIVES per round
        # (0.098 - 0.104 ms, ~550 GB/s of ingress) -- multicast saves egress, which is not the limit -- and at 2 GPUs
        # unicast is faster (0.027 against 0.038 ms), so multicast is opt-in.
        mode = mode or os.environ.get('RD_EXCHANGE_MODE') or 'p2p'
        mc = 0
        if mode == 'multicast' and self.world > 1:
            try:
                mc = int(self.hdl.multicast_ptr or 0)
            except Exception:
                mc = 0
            if not mc:
                raise RuntimeError('PeerExchange(mode="multicast"): the symmetric buffer has no multicast mapping')
        self._mc_base = ctypes.c_void_p(mc + off) if mc else ctypes.c_void_p(0)
        self.mode = 'multicast' if mc else 'p2p'
        self.copy_ctas = int(copy_ctas or os.environ.get('RD_EXCHANGE_CTAS') or 0)
        self.timeout_ms = int(timeout_ms)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                                 # every rank's control block is zero before the first round

    def exchange(self, detections, stream=None):
        """Asynchronous on ``stream`` (default: the current stream): pack + copy + rendezvous.  When the work
        completes on the stream, the rows of every rank are in the local buffer (:meth:`result`).  Read them on that
        stream (or after synchronising with it) before the next :meth:`exchange` is enqueued."""
        from ._ffi import check, lib, ptr
        B, C, max_out, _ = detections.dets.shape
        if B > self.B or C != self.C or max_out != self.max_out:
            raise ValueError('detections [%d,%d,%d] do not fit the exchange [%d,%d,%d]'
                             % (B, C, max_out, self.B, self.C, self.max_out))
        dev = self.buf.device
        ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.device(dev)
        with ctx:
            st = torch.cuda.current_stream(dev).cuda_stream
            check(lib().rd_exchange_round(ptr(detections.counts), ptr(detections.dets), B, C, max_out, self._bases,
                                          self._mc_base, self.world, self.rank, self.B, self.capacity, self.copy_ctas,
                                          self.timeout_ms, st), 'rd_exchange_round')

    def result(self):
        """``(counts_all, rows_all)`` of the latest completed round (views into the local buffer; one host sync)."""
        ctrl = self.buf[:self.ctrl_bytes].view(torch.int32).cpu()
        epoch, error = int(ctrl[64]), int(ctrl[66])
        if error:
            raise RuntimeError('PeerExchange: a rank did not arrive within %d ms (control block error flag)' % self.timeout_ms)
        half = self.buf[self.ctrl_bytes + (epoch & 1) * self.world * self.slot_bytes:
                        self.ctrl_bytes + ((epoch & 1) + 1) * self.world * self.slot_bytes]
        return decode_slots(half, self.world, self.slot_bytes, self.B * self.C)
