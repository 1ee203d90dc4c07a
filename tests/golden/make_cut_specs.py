"""Writes tests/golden/cut_specs.json: the cutter's optimum for the BASELINE.json configs (seed 0, the limits of
benchmarks/benchmark.py:41).  Run from the repo root: ``python tests/golden/make_cut_specs.py``.  Needs z3 only."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hardwareawareoptimalquantumcircuitcuttingandknitting_b200 import cutter, cutting, generators  # noqa: E402

out = {}
for config, (name, n, depth, P, q) in cutting.BASELINE_CONFIGS.items():
    circ = generators.gen_circ(name, n, depth, seed=0)
    t = time.time()
    cu = cutter.Cutter(circ, P, q, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
    ok = cu.solve()
    entry = {"generator": [name, n, depth], "p": P, "q": q, "seed": 0, "feasible": ok,
             "solve_seconds": round(time.time() - t, 2)}
    if ok:
        entry["key_results"] = list(cu.getModelKeyResults())
        entry["spec"] = json.loads(cutter.cut_spec_to_json(cu.cut_spec()))
    out[config] = entry
    print(config, entry)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cut_specs.json"), "w") as fh:
    json.dump(out, fh, indent=1)
