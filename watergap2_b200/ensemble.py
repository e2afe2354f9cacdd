"""Member sharding and ensemble statistics across GPUs (plumbing over torch.distributed).

The daily hot path shards by member (calibration parameter sets, EnKF ensemble members —
calibration.cpp / enKF2wghmState.cpp run them as separate processes); members never exchange
data inside a simulated day, so there is NO data-path collective.  The only exchange is the
statistic an assimilation cycle needs once per cycle: all-reduce of the sum and
the sum of squares of a state field over all members of all ranks (NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
import numpy as np


def shard_members(nmember_total, world_size, rank):
    """contiguous block of members owned by `rank`: (first, count); blocks differ by at most one"""
    if not (0 <= rank < world_size) or nmember_total < 0:
        raise ValueError("bad rank / world size")
    base, extra = divmod(nmember_total, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def shard_by_basin(basin_of_cell, world_size):
    """whole drainage basins to ranks, longest-processing-time first by cell count (config 5:
    high-resolution grids are split by basin; basins exchange no water without water use).
    -> rank_of_cell (int32).  Basin id 0 (single-cell basins of G_BASINS_2) is spread cell-wise."""
    b = np.asarray(basin_of_cell).astype(np.int64)
    ids, counts = np.unique(b[b > 0], return_counts=True)
    load = np.zeros(world_size, np.int64)
    rank_of_basin = {}
    for i in np.argsort(-counts, kind="stable"):
        r = int(np.argmin(load))
        rank_of_basin[int(ids[i])] = r
        load[r] += counts[i]
    out = np.empty(b.size, np.int32)
    for n in range(b.size):
        if b[n] > 0:
            out[n] = rank_of_basin[int(b[n])]
        else:
            r = int(np.argmin(load))
            out[n] = r
            load[r] += 1
    return out


def subgrid_inputs(fields, rout_order, downstream_cell, cells, ncell_table=18):
    """Inputs of a model context that holds only `cells` (0-based indices of WHOLE drainage basins, e.g. the
    cells shard_by_basin gave to one rank): per-cell fields are cut, the routing ranks are renumbered
    1..len(cells) in unchanged relative order (so the dependency levels and the order of every upstream sum
    stay those of the full grid and the shard's results are bit-identical to the full run's), downstream
    cell numbers are translated.  -> (fields, rout_order, downstream_cell) of the shard."""
    cells = np.asarray(cells, np.int64)
    ro = np.asarray(rout_order)
    dc = np.asarray(downstream_cell)
    ng = ro.size
    new_of_old = np.zeros(ng + 1, np.int64)  # 1-based old cell number -> 1-based new, 0 = outside / none
    new_of_old[cells + 1] = np.arange(1, cells.size + 1)
    sub_dc = new_of_old[dc[cells]]
    if ((dc[cells] > 0) & (sub_dc == 0)).any():
        raise ValueError("cells do not form whole drainage basins: a downstream cell lies outside the shard")
    sub_ro = np.empty(cells.size, np.int32)
    sub_ro[np.argsort(ro[cells], kind="stable")] = np.arange(1, cells.size + 1)
    out = {}
    for k, v in fields.items():
        if k.startswith("_"):
            continue
        a = np.asarray(v)
        if a.ndim >= 1 and a.shape[0] == ng:
            out[k] = a[cells]
        elif a.ndim == 2 and a.shape[1] == ng:  # params [26][ncell]
            out[k] = a[:, cells]
        elif a.ndim == 1 and a.size % ng == 0 and a.size != ncell_table and a.size > ng:  # flattened [ncell][bands]
            out[k] = a.reshape(ng, -1)[cells].ravel()
        else:
            out[k] = a  # land-cover tables, scalars
    return out, sub_ro, sub_dc.astype(np.int32)


def routing_levels(rout_order, downstream_cell):
    """0-based dependency level of every cell (longest path from a headwater)"""
    ro = np.asarray(rout_order)
    dc = np.asarray(downstream_cell)
    lvl = np.zeros(ro.size, np.int32)
    for n in np.argsort(ro, kind="stable"):
        d = dc[n]
        if d > 0 and lvl[d - 1] < lvl[n] + 1:
            lvl[d - 1] = lvl[n] + 1
    return lvl


def tile_inputs(fields, rout_order, downstream_cell, ntiles):
    """A grid of `ntiles` disjoint copies of a world (cell t*ncell + n = cell n of copy t): the way the test-suite
    and bench.py build a grid with the cell count of a 5-arcmin world (2.2 M cells) from the 0.5 degree
    generator, whose raster is fixed.  Ranks are renumbered level-major (level, copy, original rank), as
    rout_order's sweeps would number them; every copy keeps its own basins.
    -> (fields, rout_order, downstream_cell) of the tiled grid."""
    ro = np.asarray(rout_order)
    dc = np.asarray(downstream_cell)
    ng, T = ro.size, int(ntiles)
    lvl = routing_levels(ro, dc)
    order = np.lexsort((np.tile(ro, T), np.repeat(np.arange(T), ng), np.tile(lvl, T)))
    new_ro = np.empty(ng * T, np.int32)
    new_ro[order] = np.arange(1, ng * T + 1, dtype=np.int32)
    off = np.repeat(np.arange(T, dtype=np.int64) * ng, ng)
    tdc = np.tile(dc.astype(np.int64), T)
    new_dc = np.where(tdc > 0, tdc + off, 0).astype(np.int32)
    out = {}
    for k, v in fields.items():
        if k.startswith("_"):
            continue
        a = np.asarray(v)
        if a.ndim >= 1 and a.shape[0] == ng:
            out[k] = np.concatenate([a] * T, axis=0)
        elif a.ndim == 2 and a.shape[1] == ng:  # params [26][ncell]
            out[k] = np.concatenate([a] * T, axis=1)
        elif a.ndim == 1 and a.size > ng and a.size % ng == 0:  # flattened [ncell][bands]
            out[k] = np.tile(a, T)
        else:
            out[k] = a
    return out, new_ro, new_dc


class _CudaArray:
    """zero-copy view of a wgk device buffer for torch.as_tensor (CUDA array interface v2)"""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, shape, device, typestr="<f8"):
    """zero-copy torch view of a device buffer owned by a wgk context"""
    import torch
    return torch.as_tensor(_CudaArray(ptr, shape, typestr), device=f"cuda:{device}")


def ensemble_state_moments(model, nmember_total, kind="month", cells=None, group=None, timing=None):
    """Mean and population variance, over ALL members of all ranks, of the extract_sub_ state vector [ncells, 10]
    (what an assimilation cycle needs once per month).  Product path, nothing of it in eager PyTorch:
    k_ensemble_moments reads the rank's member state once and leaves sum | sumsq in one device buffer, ONE in-place
    all-reduce of 2 * ncells * 10 doubles (NCCL over NVLink, issued on the context's stream) adds the ranks, and
    k_moments_finish turns the sums into the statistics.  `timing` (a dict) receives CUDA-event times in ms of the
    moments kernel and of the collective.  -> (mean, var) as host arrays [ncells, 10]."""
    import torch
    import torch.distributed as dist
    ps, _, n = model.ensemble_moments(kind, cells)
    buf = device_tensor(ps, (2, n, 10), model.device)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    stream = torch.cuda.ExternalStream(model.stream, device=model.device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timing is not None else None
    with torch.cuda.stream(stream):
        if ev:
            ev[0].record(stream)
        if multi:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if ev:
            ev[1].record(stream)
    mean, var = model.moments_finish(nmember_total, n)
    if ev:
        timing["collective_ms"] = ev[0].elapsed_time(ev[1])
        timing["allreduce_bytes"] = int(buf.numel() * 8)
    return mean, var


def ensemble_mean_var(local, nmember_total, group=None):
    """mean and (population) variance over ALL members of all ranks of a generic tensor `local`
    [members_on_this_rank, n]: the host-logic form of the exchange (member sharding + one all-reduce of the stacked
    sums), used by the gloo CPU tests and for fields outside the state vector.  The model state goes through
    ensemble_state_moments (CUDA reduction kernel + NCCL)."""
    import torch
    import torch.distributed as dist
    sums = torch.stack([local.sum(0), torch.einsum("mn,mn->n", local, local)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    mean = sums[0] / nmember_total
    var = torch.clamp(sums[1] / nmember_total - mean * mean, min=0.0)
    return mean, var
