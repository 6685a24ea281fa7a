"""ROI sharding across the GPUs of one box + the single pose all-gather.

Mirrors the reference's inference sharding: contiguous shards per rank
(core/utils/my_distributed_sampler.py:172-200, InferenceSampler) and one gather of the predictions at the end
(core/gdrn_modeling/engine/gdrn_evaluator.py:575-583), which there is a pickled-object all_gather
(core/utils/my_comm.py:70-171) and here is one all_gather of a [n_local,12] float tensor (R row-major 9 + t 3).
Works with the NCCL backend on GPUs and with gloo on CPU tensors (tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [begin, end) of `total` items for `rank` (InferenceSampler semantics: ceil-sized shards)."""
    shard = (total - 1) // world + 1 if total > 0 else 0
    begin = min(shard * rank, total)
    end = min(shard * (rank + 1), total)
    return begin, end


def pack_poses(rot, trans):
    """[n,3,3], [n,3] -> [n,12]"""
    return torch.cat([rot.reshape(-1, 9), trans.reshape(-1, 3)], dim=1).contiguous()


def unpack_poses(p):
    return p[:, :9].reshape(-1, 3, 3), p[:, 9:12]


def all_gather_poses(rot, trans, total=None):
    """Gather every rank's poses -> ([N,3,3], [N,3]) on all ranks, in rank order.

    Shards may be ragged (last ranks shorter or empty): they are padded to the largest shard for the
    collective and trimmed afterwards.  `total` (optional) = global ROI count, used to derive shard sizes
    without an extra size exchange."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rot, trans
    world = dist.get_world_size()
    local = pack_poses(rot, trans)
    if total is not None:
        sizes = [shard_range(total, r, world) for r in range(world)]
        sizes = [e - b for b, e in sizes]
    else:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(all_n, n)
        sizes = [int(t.item()) for t in all_n]
    mx = max(sizes) if sizes else 0
    if mx == 0:
        return rot, trans
    buf = torch.zeros((mx, 12), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * mx, 12), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf) if hasattr(dist, "all_gather_into_tensor") and local.is_cuda else \
        _all_gather_list(out, buf, world, mx)
    parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
    return unpack_poses(torch.cat(parts, dim=0))


def _all_gather_list(out, buf, world, mx):
    chunks = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(chunks, buf)
    for r in range(world):
        out[r * mx:(r + 1) * mx] = chunks[r]
