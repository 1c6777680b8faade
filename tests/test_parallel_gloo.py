"""world_size-2 gloo test of the multi-rank host logic (sharding, all-gather, first-wins selection,
trace all-reduce) with the oracle standing in for the per-rank device work."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import warnings
    warnings.simplefilter("ignore")
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from conftest import load_graph, edge_UB
    from krylov_robustness_b200 import parallel as P
    from krylov_robustness_b200.functions import select_candidate
    A = load_graph("oregon_A0")
    n = A.shape[0]
    nrm, _ = O.normest(A, 1e-2)
    tol = 1e-6 * np.exp(nrm)
    c = O.compute_centrality(A, "eig")
    E = O.find_top_edges(A, c, 11, "min")            # odd count: ragged shards

    def score(Es):
        out = []
        for i, j in Es:
            U, B = edge_UB(n, int(i), int(j), -1.0)
            out.append(O.trace_fun_update(A, U, B, tol, 100)[0])
        return out

    vals = P.sharded_edge_scores(score, E)
    best, bestval = select_candidate(vals, "break")
    Z = np.sign(np.random.default_rng(3).standard_normal((n, 7)))
    As = A / 8.0
    tr = P.sharded_probe_trace(lambda lo, hi: O.slq_trace(As, Z[:, lo:hi], 12, "exp")[1].sum(), 7)
    lo, hi = P.shard_bounds(11)
    # C2 sharding: distinct first indices of omega split across the ranks (pairs repeat first indices on purpose)
    Om = np.array([[5, 9], [17, 3], [5, 40], [201, 7], [17, 17], [88, 2], [201, 5]], dtype=np.int64)
    ent, ent_it = P.sharded_entries(lambda om: O.function_multiple_entries(As, om, "exp", 1e-12, 60), Om)
    # C4 sharding: right-hand-side columns split, full_term evaluators, degree table of the FULL block
    Bm = np.random.default_rng(5).standard_normal((n, 5))
    Mtab = O.select_taylor_degree(As, Bm)[0]
    fexp = P.sharded_expmv(lambda bp: O.expmv(1.0, As, bp, Mtab, "double", True, False, True)[0], Bm)
    q.put((rank, vals.tolist(), best, bestval, tr, (lo, hi), ent.tolist(), ent_it, fexp.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_serial(graphs):
    import oracle as O
    from conftest import edge_UB
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A = graphs("oregon_A0")
    n = A.shape[0]
    nrm, _ = O.normest(A, 1e-2)
    tol = 1e-6 * np.exp(nrm)
    c = O.compute_centrality(A, "eig")
    E = O.find_top_edges(A, c, 11, "min")
    serial = []
    for i, j in E:
        U, B = edge_UB(n, int(i), int(j), -1.0)
        serial.append(O.trace_fun_update(A, U, B, tol, 100)[0])
    Z = np.sign(np.random.default_rng(3).standard_normal((n, 7)))
    tr_serial = O.slq_trace(A / 8.0, Z, 12, "exp")[0]
    assert res[0][5] == (0, 6) and res[1][5] == (6, 11)
    Om = np.array([[5, 9], [17, 3], [5, 40], [201, 7], [17, 17], [88, 2], [201, 5]], dtype=np.int64)
    ent_serial, it_serial = O.function_multiple_entries(A / 8.0, Om, "exp", 1e-12, 60)
    Bm = np.random.default_rng(5).standard_normal((n, 5))
    Mtab = O.select_taylor_degree(A / 8.0, Bm)[0]
    f_serial = O.expmv(1.0, A / 8.0, Bm, Mtab, "double", True, False, True)[0]
    for rank, vals, best, bestval, tr, _, ent, ent_it, fexp in res:
        assert vals == serial                       # identical full vector on every rank
        assert best == int(np.argmin(serial)) and bestval == min(serial)
        assert abs(tr - tr_serial) <= 1e-12 * abs(tr_serial)
        assert np.array_equal(np.array(ent), ent_serial) and ent_it == it_serial      # spaces are independent
        assert np.array_equal(np.array(fexp), f_serial)                               # columns are independent (full_term)


def test_shard_bounds_cover_everything():
    from krylov_robustness_b200.parallel import shard_bounds
    for n in (0, 1, 7, 512, 100001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_numa_binding_is_a_noop_without_topology():
    """bind_to_gpu_numa_node must never raise: no GPU / no sysfs topology -> None and the affinity is untouched."""
    import os
    from krylov_robustness_b200.parallel import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), dict)
    if not os.path.exists("/dev/nvidia0"):
        assert os.sched_getaffinity(0) == before
