#!/usr/bin/env python
"""Pin parity: run the REFERENCE's own source files on the committed inputs and write the golden file.

    python scripts/run_reference_goldens.py [--refdir /root/reference] [--check]

Executes scripts/make_reference_goldens.m - the script a maintainer with Octave / MATLAB would run - with the
MATLAB-subset interpreter of oracle/mlab.  Every function it calls (trace_fun_update, fun_update, lanczos_krylov,
arnoldi_krylov, function_multiple_entries, expmv, select_taylor_degree, normAm, mc_trace, krylov_miobi, greedy_krylov,
find_top_edges, find_top_missing_edges, multiple_frechet_eval, hessianfcn_*, fun_and_grad_krylov_*) is read from
<refdir>/functions/*.m at run time, unmodified; nothing of the reference is copied into this repository.  Output:
tests/golden/reference_golden.json (the values) and tests/golden/reference_golden.provenance.json (which engine
produced them, how often each reference function ran, SHA-256 of each source file that was executed).

--check re-runs and compares with the committed file instead of writing (used by tests/test_mlab_reference.py when
/root/reference is present)."""
import argparse
import hashlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


SCRIPTS = {"main": ("make_reference_goldens.m", "reference_golden"),
           "c1": ("make_reference_goldens_c1.m", "reference_golden_c1"),                 # ~10 min
           "c1trace": ("make_reference_goldens_c1_trace.m", "reference_golden_c1_trace")}  # ~5 min


def run(refdir, out_json=None, which="main"):
    """-> (golden dict, provenance dict).  The .m script writes the JSON itself; `out_json` redirects it."""
    script, stem = SCRIPTS[which]
    from oracle.mlab import Interpreter
    interp = Interpreter(stdout=io.StringIO())
    ws = {"refdir": refdir}
    if out_json is not None:
        ws["golden_path"] = out_json
    t0 = time.time()
    interp.run_script(os.path.join(ROOT, "scripts", script), ws)
    path = out_json or os.path.join(ROOT, "tests", "golden", stem + ".json")
    golden = json.load(open(path))
    executed = sorted(f for f in interp.cache if os.path.abspath(f).startswith(os.path.abspath(refdir)))
    prov = {
        "engine": "oracle/mlab (MATLAB-subset interpreter of this repository) executing the reference's unmodified "
                  "functions/*.m; built-ins are NumPy/SciPy (LAPACK) stand-ins for MATLAB's closed-source ones",
        "script": "scripts/" + script,
        "reference_files_executed": {os.path.relpath(f, refdir): hashlib.sha256(open(f, "rb").read()).hexdigest()
                                     for f in executed},
        "reference_function_calls": dict(sorted(interp.calls.items())),
        "warnings_raised_by_the_reference": sorted(set(interp.warnings)),
        "seconds": round(time.time() - t0, 1),
    }
    return golden, prov


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refdir", default="/root/reference")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--which", default="main", choices=sorted(SCRIPTS), help="main (25 s); c1: the BASELINE-C1 greedy runs; "
                    "c1trace: trace_exp on four Oregon graphs")
    ap.add_argument("--c1", action="store_true", help="same as --which c1")
    a = ap.parse_args()
    if a.c1:
        a.which = "c1"
    stem = SCRIPTS[a.which][1]
    if not os.path.isdir(os.path.join(a.refdir, "functions")):
        sys.exit("no reference checkout at %s" % a.refdir)
    gpath = os.path.join(ROOT, "tests", "golden", stem + ".json")
    if a.check:
        import tempfile
        import numpy as np
        tmp = os.path.join(tempfile.mkdtemp(), "g.json")
        fresh, _ = run(a.refdir, tmp, a.which)
        stored = json.load(open(gpath))
        assert set(fresh) == set(stored), sorted(set(fresh) ^ set(stored))
        for k in fresh:
            assert np.array_equal(np.asarray(fresh[k]), np.asarray(stored[k])), k
        print("%s.json reproduced bit for bit (%d entries)" % (stem, len(fresh)))
        return
    golden, prov = run(a.refdir, None, a.which)
    with open(os.path.join(ROOT, "tests", "golden", stem + ".provenance.json"), "w") as fh:
        json.dump(prov, fh, indent=1)
        fh.write("\n")
    print("wrote %s (%d entries) in %.1f s; reference functions executed: %s"
          % (os.path.relpath(gpath, ROOT), len(golden), prov["seconds"], ", ".join(prov["reference_function_calls"])))


if __name__ == "__main__":
    main()
