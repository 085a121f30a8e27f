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
