"""Frame-batch sharding of a stream across the GPUs of one box.

The chain has no cross-frame state (the reference is one process per image), so a stream of `total`
frames is cut into contiguous ranges, one per rank, and the data path needs no collective.  The only
cross-rank step is bookkeeping: frame counts and per-frame checksums are gathered on the host side with
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""


def frame_range(rank, world, total):
    """Contiguous range [first, first + count) of rank `rank`: sizes differ by at most one frame."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank / world / total")
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def batches(first, count, batch):
    """Cut one rank's range into consecutive batches of at most `batch` frames."""
    if batch < 1:
        raise ValueError("batch must be >= 1")
    out = []
    f = first
    while f < first + count:
        n = min(batch, first + count - f)
        out.append((f, n))
        f += n
    return out


def gather_checksums(local_sums, rank, world, total, dist=None, device="cpu"):
    """All ranks' per-frame uint64 checksums in stream order (every rank gets the full vector).

    `local_sums` are the checksums of this rank's frame_range, in order.  Ranges may differ by one frame,
    so every rank pads to the longest range before the all_gather."""
    import numpy as np
    import torch

    first, count = frame_range(rank, world, total)
    local = np.asarray(local_sums, dtype=np.uint64)
    if local.shape != (count,):
        raise ValueError("rank %d holds %d checksums for a range of %d frames" % (rank, local.size, count))
    if world == 1 or dist is None:
        return local
    longest = frame_range(0, world, total)[1]
    pad = np.zeros(longest, dtype=np.int64)
    pad[:count] = local.view(np.int64)
    mine = torch.from_numpy(pad).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.empty(total, dtype=np.uint64)
    for r in range(world):
        f, c = frame_range(r, world, total)
        out[f:f + c] = parts[r].cpu().numpy()[:c].view(np.uint64)
    return out
