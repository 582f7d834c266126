"""Multi-GPU partitioning of the ACF hot path inside one box (one process per GPU, torch.distributed).

Two cases; the reference itself is single-process:
  * independent series -- umbrella windows (setup.sh:6-13), the six tensor components of viscosity.cpp:197,
    PACF/VACF of one window -- are dealt out round-robin, item i on rank i mod G.  No data-path collective;
    results are gathered to rank 0 as Python objects.
  * one long series is cut into G contiguous time blocks.  Rank g receives x[lo_g, min(n, hi_g + m - 1)) (its
    block plus an (m-1)-sample halo), computes the raw lag sums of the pairs whose first element lies in its
    block, and ONE all-reduce of m doubles per series (NCCL over NVLink on GPUs) combines them; the divide by
    (n-k) (ACFcalculator.cpp:139-143) happens after the sum.

torch is plumbing here (process group, tensors); the per-block arithmetic is the CUDA library's
acfb200_lag_sums_device.  `lag_sums_fn` is injectable so the partitioning/collective logic is also exercised
on CPU with the gloo backend (tests/test_parallel_cpu.py).
"""
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_items(nitems, world, rank):
    """Indices of the independent series owned by `rank` (round-robin, item i -> rank i mod world)."""
    return list(range(rank, nitems, world))


def time_block(n, m, world, rank):
    """(lo, hi, halo_end): rank owns first-elements i in [lo, hi) and needs samples [lo, halo_end)."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi, min(n, hi + max(m - 1, 0))


def _gpu_lag_sums(block, i_count, m):
    from . import api
    return api.lag_sums_device(block, m, i_end=i_count)


def _gpu_normalise(sums, n_total):
    from . import api
    return api.normalise_device(sums, n_total)


def _host_normalise(sums, n_total):
    k = torch.arange(sums.numel(), dtype=torch.float64, device=sums.device)
    return sums / (float(n_total) - k)


def time_block_correlation(block, i_count, n_total, m, lag_sums_fn=None, normalise_fn=None, group=None):
    """calcCorrelation of a series of n_total samples of which this rank holds `block` = its time block plus halo.

    block      1-D float64 tensor: x[lo, halo_end) of this rank
    i_count    hi - lo, the number of first-elements this rank owns
    returns    corr[m] (identical on every rank)
    """
    lag_sums_fn = lag_sums_fn or _gpu_lag_sums
    sums = lag_sums_fn(block, i_count, m)
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    if normalise_fn is None:
        normalise_fn = _gpu_normalise if sums.is_cuda else _host_normalise
    return normalise_fn(sums, n_total)


def time_block_correlation_batch(blocks, i_count, n_total, m, batch_fn, normalise_fn, group=None):
    """Several series (e.g. the Green-Kubo components) split at the same block edges: one launch for all
    local blocks, ONE all-reduce of [len(blocks), m] doubles, then the normalisation."""
    sums = batch_fn(blocks, i_count, m)  # [C, m]
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return torch.stack([normalise_fn(sums[c], n_total) for c in range(sums.shape[0])])


def time_block_correlation_batch_host(h_blocks, i_count, n_total, m, stream=None, group=None):
    """Host-buffer form of time_block_correlation_batch for one process per GPU: `h_blocks` are this rank's time blocks
    (block + halo) of several series in host memory (numpy arrays; pinned memory makes the copies asynchronous).  The
    library uploads every block in pieces under its own lag kernels (acfb200_lag_sums_blocks), then ONE all-reduce of
    [len(h_blocks), m] doubles and the divide by (n_total - k).  Returns corr [C, m] on the device (same on every rank);
    the caller's .cpu() is the device-to-host read."""
    from . import api
    st = torch.cuda.current_stream() if stream is None else stream
    sums = api.lag_sums_blocks(h_blocks, [int(i_count)] * len(h_blocks), int(m))   # complete on return
    _, world = world_info(group)
    if world > 1:
        with torch.cuda.stream(st):
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return torch.stack([api.normalise_device(sums[c], n_total, stream=st) for c in range(len(h_blocks))])


def gather_results(local, nitems, group=None):
    """Collect {item index: result} dictionaries from all ranks on rank 0 (None elsewhere)."""
    rank, world = world_info(group)
    if world == 1:
        return [local[i] for i in range(nitems)]
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for part in bucket:
        merged.update(part)
    return [merged[i] for i in range(nitems)]
