"""SpMM tuning harness: time kr_spmm_dev (k = 512, C3 graph) for several library builds / env settings.

usage: python scripts/exp_spmm.py name=lib.so[,ENV=VAL,...] ...
The graph is generated once and cached in /tmp; every variant runs in its own process (the library is
loaded once per process) and is checked against SciPy on a 16-column block before it is timed.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/kr_exp_graph.npz"
N = int(os.environ.get("KR_BENCH_N", 1_000_000))
NNZ = int(os.environ.get("KR_BENCH_NNZ", 20_000_000))
K = int(os.environ.get("KR_BENCH_K", 512))


def graph():
    if not os.path.exists(CACHE):
        from krylov_robustness_b200.graphs import power_law_graph
        A = power_law_graph(N, NNZ, 2.2, 20260310)
        np.savez(CACHE, indptr=A.indptr, indices=A.indices)
    z = np.load(CACHE)
    return sp.csr_matrix((np.full(z["indices"].size, 1.0 / 64.0), z["indices"], z["indptr"]), shape=(N, N))


def child(name):
    import ctypes as C
    from krylov_robustness_b200.engine import Context, Dense, Matrix, rademacher_host
    A = graph()
    ctx = Context.default()
    t0 = time.time()
    M = Matrix(A, ctx)
    t_up = time.time() - t0
    xs = rademacher_host(N, 16, 7)
    err = np.abs(M.matmul(xs) - A @ xs).max()
    X, Y = Dense(N, K, ctx).fill_rademacher(1), Dense(N, K, ctx)
    lib = ctx.lib
    for _ in range(3):
        lib.kr_spmm_dev(ctx.h, M.h, X.h, Y.h)
    ctx.sync()
    ctx.set_timing(True)
    ctx.spmm_time(True)
    for _ in range(8):
        lib.kr_spmm_dev(ctx.h, M.h, X.h, Y.h)
    ctx.sync()
    ms, cnt = ctx.spmm_time(True)
    # sustained: ~4 s of back-to-back launches (the 30-step SLQ pass of bench.py is power-capped); last 40 launches
    sustained = None
    if os.environ.get("KR_EXP_SUSTAINED", "1") != "0":
        ctx.set_timing(False)
        for _ in range(int(os.environ.get("KR_EXP_SUSTAINED_LAUNCHES", 500))):
            lib.kr_spmm_dev(ctx.h, M.h, X.h, Y.h)
        ctx.set_timing(True)
        ctx.spmm_time(True)
        for _ in range(40):
            lib.kr_spmm_dev(ctx.h, M.h, X.h, Y.h)
        ctx.sync()
        ms_s, cnt_s = ctx.spmm_time(True)
        sustained = ms_s / cnt_s
    # the Lanczos flavour (EpiDot epilogue: every row also reads its own row of X), inside a short SLQ run
    import krylov_robustness_b200 as kr
    kr.slq_trace(M, X, 3, "exp")
    ctx.sync()
    ctx.spmm_time(True)
    t0 = time.time()
    kr.slq_trace(M, X, 6, "exp")
    ctx.sync()
    t_slq = time.time() - t0
    ms_dot, cnt_dot = ctx.spmm_time(True)
    print(json.dumps({"variant": name, "ms_per_spmm": ms / cnt, "launches": cnt, "max_abs_err_k16": err, "ms_per_spmm_sustained": sustained, "ms_per_spmm_dot": ms_dot / max(cnt_dot, 1),
                      "slq6_ms": t_slq * 1e3,
                      "upload_s": t_up}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    graph()
    for spec in sys.argv[1:]:
        name, rest = spec.split("=", 1)
        parts = rest.split(",")
        env = dict(os.environ)
        env["KR_B200_LIB"] = os.path.join(ROOT, "krylov_robustness_b200", parts[0])
        for kv in parts[1:]:
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], env=env,
                           capture_output=True, text=True, timeout=600)
        out = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(out[-1] if out else json.dumps({"variant": name, "error": (r.stderr or r.stdout)[-400:]}), flush=True)
