"""Multi-GPU plumbing for the IBF classifier: one process per GPU, torch.distributed underneath.

Two ways the path shards:
  * read-sharded, IBF replicated: reads are independent, rank r classifies reads [lo, hi); there is
    no data-path collective.  `gather_results` is only for a caller that wants everything on rank 0.
  * bin-sharded: rank r holds row words [W*r/N, W*(r+1)/N) of every row (rb_ibf_load_shard); every
    rank classifies ALL reads against its bins; the per-read packed summary keys (rb_ibf.h RB_KEY_*)
    are combined with ONE all-reduce(MAX) -- 8 bytes per read and threshold table, over NVLink with
    NCCL on GPUs (gloo on CPU in the tests).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Balanced contiguous split of n reads: rank r gets [lo, hi)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bin_shard_columns(bin_width, rank, world):
    """(col_begin, col_words) of a rank's column slice; same arithmetic as derive_geometry() in the C ABI."""
    b = bin_width * rank // world
    e = bin_width * (rank + 1) // world
    return b, e - b


def per_rank_bin_ranges(bins_per_rank, world):
    """Bin ranges [lo, hi) of the column slices when every rank contributes `bins_per_rank` bins (a multiple of 64) to ONE
    filter that is built slice by slice (rb_ibf_create_shard): the filter has world * bins_per_rank bins in
    ceil(bins / 64) = world * bins_per_rank / 64 row words (derive_geometry in the C ABI), split evenly, so slice r is
    exactly the bins [r * bins_per_rank, (r + 1) * bins_per_rank)."""
    assert bins_per_rank % 64 == 0 and bins_per_rank > 0
    n_bins = bins_per_rank * world
    bin_width = (n_bins + 63) // 64
    out = []
    for r in range(world):
        b, w = bin_shard_columns(bin_width, r, world)
        out.append((64 * b, min(n_bins, 64 * (b + w))))
    return out


def combine_keys(keys, group=None):
    """In-place elementwise MAX of packed summary keys across ranks (int64 tensor; keys are < 2^49).  The C-ABI twins for
    a C++ host: rb_keys_combine_nccl (one process per GPU) and rb_ibf_count_batch_sharded (one process, keys folded over
    NVLink peer memory by the count kernels themselves)."""
    assert keys.dtype == torch.int64
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys


def gather_results(local, n_total, group=None):
    """Read-sharded mode: concatenate per-rank 1-D result tensors (shard_range order) on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_range(n_total, r, world) for r in range(world))]
    width = max(sizes)                      # all_gather wants equal shapes: pad, then trim
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    bufs = [torch.empty(width, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])
