"""Multi-GPU sharding of the path: A is replicated on every GPU, the independent units
(probe columns, candidate edges, distinct row indices) are split across ranks, and ONE tiny collective
per outer step brings the scalars together (all-reduce of the partial trace, all-gather of the
candidate scores).  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Nothing here computes: the per-rank work is a callable (the device entry points in production, the
oracle in the world_size-2 gloo tests), so the sharding / gathering / tie-breaking logic is testable
without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _device():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def shard_bounds(n_items, rank=None, world=None):
    """Contiguous, balanced split: items [lo, hi) belong to `rank`."""
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum(value):
    """Sum of a Python float over ranks (the single all-reduce of the trace / objective scalars)."""
    _, world = _world()
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allgather_concat(local, n_total):
    """Concatenate per-rank float64 vectors (shards of shard_bounds order) into the full vector."""
    _, world = _world()
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local
    cap = -(-int(n_total) // world)
    buf = torch.zeros(cap, dtype=torch.float64, device=_device())
    buf[:local.size] = torch.from_numpy(local).to(buf.device)
    out = [torch.zeros(cap, dtype=torch.float64, device=buf.device) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r][:hi - lo].cpu().numpy())
    return np.concatenate(parts)


def sharded_edge_scores(score_fn, E):
    """Scores for ALL candidate edges E (k x 2), every rank scoring only its contiguous shard with
    `score_fn(E_shard) -> vector`, followed by one all-gather.  Every rank returns the identical full
    vector, so the reference's first-wins arg-min / arg-max (functions/krylov_miobi.m:112-124) picks
    the same edge everywhere."""
    E = np.atleast_2d(np.asarray(E))
    lo, hi = shard_bounds(E.shape[0])
    local = np.asarray(score_fn(E[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    return allgather_concat(local, E.shape[0])


def make_sharded_scorer(fun="exp"):
    """scorer(M, E, b_offdiag, tol, it) for functions.greedy_krylov / krylov_miobi that splits the
    candidates across ranks (device path on each rank)."""
    from . import functions as F

    def scorer(M, E, b_offdiag, tol, it):
        return sharded_edge_scores(lambda Es: F.trace_fun_update_edges(M, Es, b_offdiag, tol, it, fun)[0], E)
    return scorer


def sharded_probe_trace(local_sum_fn, k_total):
    """Hutchinson / SLQ estimate over k_total probe columns split across ranks:
    `local_sum_fn(lo, hi)` returns the SUM of the per-probe values of columns [lo, hi); the result is
    the global mean (one all-reduce of one double)."""
    lo, hi = shard_bounds(k_total)
    s = float(local_sum_fn(lo, hi)) if hi > lo else 0.0
    return allreduce_sum(s) / float(k_total)


def sharded_entries(entries_fn, omega):
    """f(A)(i, j) for the index pairs omega (k x 2, 1-based) with the DISTINCT FIRST INDICES split across the ranks
    (one single-vector Krylov space per distinct row, functions/function_multiple_entries.m:42,
    86-110 - the spaces are independent, so a rank advances only its share of them).  ``entries_fn(omega_part) ->
    (values, iter)`` is the per-rank evaluator (functions.function_multiple_entries on the device).  Returns the full
    value vector in the order of omega on every rank and the largest step count (one all-gather of the disjoint
    values + one all-reduce of an int)."""
    rank, world = _world()
    omega = np.atleast_2d(np.asarray(omega)).astype(np.int64)
    k = omega.shape[0]
    first = omega[:, 0]
    uniq, inv = np.unique(first, return_inverse=True)          # the split must be the same on every rank
    lo, hi = shard_bounds(uniq.size)
    mine = np.nonzero((inv >= lo) & (inv < hi))[0]
    if mine.size:
        vals, it = entries_fn(omega[mine])
        vals, it = np.asarray(vals, dtype=np.float64), int(it)
    else:
        vals, it = np.zeros(0), 0
    if world == 1:
        return vals, it
    # disjoint contributions: scatter into a zero vector and sum (every pair is owned by exactly one rank)
    full = torch.zeros(k, dtype=torch.float64, device=_device())
    if mine.size:
        full[torch.from_numpy(mine).to(full.device)] = torch.from_numpy(vals).to(full.device)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    itmax = torch.tensor([it], dtype=torch.int64, device=_device())
    dist.all_reduce(itmax, op=dist.ReduceOp.MAX)
    return full.cpu().numpy(), int(itmax.item())


def sharded_expmv(expmv_fn, b):
    """e^{tA} b with the COLUMNS of b (n x q) split across the ranks.  The Taylor recurrences
    of the columns are independent once the degree m and the scaling s are fixed, but the early-termination test of
    functions/expmv.m:83 uses matrix inf-norms over ALL columns - so the per-rank evaluator must run with
    ``full_term=True`` (every rank then computes exactly the terms a single-GPU full_term run computes for its
    columns; SURVEY.md 8e names this option) and must be given the degree table M of the FULL block (select_taylor_degree
    looks at size(b, 2), functions/select_taylor_degree.m:44).  ``expmv_fn(b_part) -> f_part``.  Returns the full n x q
    result on every rank (one all-gather)."""
    rank, world = _world()
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = b[:, None]
    n, q = b.shape
    lo, hi = shard_bounds(q)
    fpart = np.asarray(expmv_fn(b[:, lo:hi]), dtype=np.float64).reshape(n, hi - lo) if hi > lo else np.zeros((n, 0))
    if world == 1:
        return fpart
    cap = -(-q // world)
    buf = torch.zeros((cap, n), dtype=torch.float64, device=_device())
    if hi > lo:
        buf[:hi - lo] = torch.from_numpy(np.ascontiguousarray(fpart.T)).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    cols = []
    for r in range(world):
        l, h = shard_bounds(q, r, world)
        cols.append(out[r][:h - l].cpu().numpy().T)
    return np.concatenate(cols, axis=1)


def bind_to_gpu_numa_node(device_index):
    """Pin this process (CPU affinity, hence first-touch page placement of everything allocated afterwards,
    pinned staging buffers included) to the NUMA node the GPU hangs off.  On a two-socket host a rank whose
    pinned probe block lives on the far socket uploads it several times slower.  Host plumbing only; returns a
    small dict for logging, or None when the topology is not exposed (single node, container without sysfs)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"gpu": bdf, "node": node, "cpus": cpulist, "bound_cpus": len(allowed)}
    except Exception:
        return None

