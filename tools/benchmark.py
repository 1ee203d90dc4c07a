"""Same command line as the reference's driver (benchmarks/benchmark.py:22-29):

    python tools/benchmark.py -p 2 -q 10 [syc|hwe|bv|qft|aqft|add] <nQubits> <depth> [--cut-only] [--seed S]

generate the circuit, solve the cut model with the limits of benchmarks/benchmark.py:41, log the key results and -
unless --cut-only (the reference ships with CUT_ONLY = True, benchmarks/benchmark.py:20) - run the uncut and the cut
circuit on the GPU and log the three fidelities of Utilities.compareOriginalCircWithCutCirc.  The cut spec is
written next to the log as JSON instead of the reference's circuit drawings (no matplotlib here)."""
import argparse
import datetime
import logging
import os
import pathlib
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("-p", type=int, default=2, help="maximum number of partitions")
    ap.add_argument("-q", type=int, default=10, help="maximum number of qubits per partition")
    ap.add_argument("name")
    ap.add_argument("n_qubits", type=int)
    ap.add_argument("depth", type=int)
    ap.add_argument("--cut-only", action="store_true")
    ap.add_argument("--seed", type=int, default=0, help="syc / sup circuits are random (the reference does not seed)")
    ap.add_argument("--out", default="./benchmark_results")
    args = ap.parse_args(argv)
    from importlib import import_module
    generators = import_module(f"{PKG}.generators")
    cutter_mod = import_module(f"{PKG}.cutter")

    out_dir = pathlib.Path(args.out) / (f"{args.name}_{args.n_qubits}_{args.depth}_{args.p}_{args.q}_"
                                        f"{datetime.datetime.now():%Y%m%d_%H%M%S}")
    out_dir.mkdir(parents=True, exist_ok=True)
    log = logging.getLogger("benchmark")            # stream INFO + run.log, as the reference's Logger.py:24-59
    log.setLevel(logging.INFO)
    log.propagate = False
    for h in list(log.handlers):
        log.removeHandler(h)
        h.close()
    for h in (logging.StreamHandler(), logging.FileHandler(out_dir / "run.log")):
        h.setFormatter(logging.Formatter("%(asctime)s %(message)s"))
        log.addHandler(h)

    circ = generators.gen_circ(args.name.lower(), args.n_qubits, args.depth, seed=args.seed)
    cutter = cutter_mod.Cutter(circ, args.p, args.q, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
    start = datetime.datetime.now()
    log.info("solving STARTED")
    success = cutter.solve()
    log.info("solving DONE")
    log.info(f"solving time elapsed: {datetime.datetime.now() - start}")
    log.info(f"success => {success}")
    if not success:
        return 0
    decomposed, _marked, _with_moves, cut, _inst = cutter.getResultCircs(getInstantiations=False)
    S, A, L, n_wire, n_gate, Q, Q_p, C, C_p = cutter.getModelKeyResults()
    for key, val in (("S", S), ("A", A), ("L", L), ("Q", Q), ("C", C), ("nWireCuts", n_wire), ("nGateCuts", n_gate)):
        log.info(f"{key}: {val}")
    for i, (q_p, c_p) in enumerate(zip(Q_p, C_p)):
        log.info(f"  Q_p{i}: {q_p}  C_p{i}: {c_p}")
    (out_dir / "cut_spec.json").write_text(cutter_mod.cut_spec_to_json(cutter.cut_spec()))
    log.info(f"fragments: {[len(r) for r in cut.qregs]} qubits; cut spec -> {out_dir / 'cut_spec.json'}")
    if args.cut_only:
        log.info("--cut-only => Simulation will not run.")
        return 0
    utilities = import_module(f"{PKG}.utilities")
    n_shots = 1000
    log.info("Circuits will be run (exact distributions) to calculate fidelity...")
    f_in, f_cut, f_cross = utilities.compareOriginalCircWithCutCirc(decomposed, cut, None, n_shots)
    log.info(f"inputCircFidelity: {f_in}")
    log.info(f"cutCircFidelity: {f_cut}")
    log.info(f"cutVsUncutFidelity: {f_cross}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
