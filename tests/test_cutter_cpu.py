"""Qiskit-free cutter - host side, CPU only.

The reference has no tests for ``Cutter.py``; what pins the restatement here is (1) an exhaustive search over every
vertex assignment / teleport choice of small circuits that evaluates the reference's constraints and objectives
(``Cutter.py:383-567``) directly, (2) the cut shapes SURVEY Appendix B / C.4 records for the BASELINE configs, and
(3) the end-to-end property that the cut circuit knits back to the uncut distribution (oracle on both sides).
"""
import itertools
from importlib import import_module

import numpy as np
import pytest

from conftest import oracle_knit
from oracle import statevector as sv

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutter_mod = import_module(f"{PKG}.cutter")
cutting = import_module(f"{PKG}.cutting")
circuit = import_module(f"{PKG}.circuit")
generators = import_module(f"{PKG}.generators")
z3 = pytest.importorskip("z3")

Cutter = cutter_mod.Cutter


def _brute_force(cu, max_cuts=None, max_qpd=None, max_cuts_per_part=None):
    """Lexicographic optimum (soft, Q, S, A, L, C) by enumeration; None if infeasible."""
    P, nv = cu.maxNPartitions, len(cu.V)
    best = None
    for assign in itertools.product(range(P), repeat=nv):
        cut_edges = [i for i, (u, v, _t) in enumerate(cu.edges) if assign[u] != assign[v]]
        if max_cuts is not None and len(cut_edges) > max_cuts:
            continue
        for tele in itertools.product((False, True), repeat=len(cut_edges)):
            tele_set = {e for e, b in zip(cut_edges, tele) if b}
            qpd_set = [e for e in cut_edges if e not in tele_set]
            if max_qpd is not None:
                if len(qpd_set) > max_qpd or (tele_set and len(qpd_set) != max_qpd):
                    continue
            Qp, Cp = [], []
            for p in range(P):
                q = sum(1 for v in cu.I if assign[v.idx] == p)
                q += sum(1 for e in cut_edges if cu.edges[e][2] == "W" and assign[cu.edges[e][1]] == p)
                q += sum(1 for e in tele_set if p in (assign[cu.edges[e][0]], assign[cu.edges[e][1]]))
                Qp.append(q)
                Cp.append(sum(1 for e in qpd_set if p in (assign[cu.edges[e][0]], assign[cu.edges[e][1]])))
            if any(q > m for q, m in zip(Qp, cu.maxNQubitsPerPartition)):
                continue
            if max_cuts_per_part is not None and max(Cp) > max_cuts_per_part:
                continue
            S, A, L = 1, 0, 0
            for e in cut_edges:
                gate = cu.edges[e][2] == "G"
                if e in tele_set:
                    A += 2; L += 10
                else:
                    S *= 6 if gate else 8
                    A += 0 if gate else 1
            A *= S
            mx = max([cu.edges[e][1] for e in qpd_set], default=-1)
            mn = min([cu.edges[e][0] for e in tele_set], default=nv)
            key = (0 if mx < mn else 1, max(Qp), S, A, L, max(Cp))
            if best is None or key < best:
                best = key
    return best


def _small_circuits():
    QC, QR = circuit.QuantumCircuit, circuit.QuantumRegister
    out = {}
    c = QC(QR(4, "q"))
    c.h(0); c.cx(0, 1); c.cx(1, 2); c.cx(2, 3); c.cx(0, 1)
    out["chain4"] = c
    c = QC(QR(4, "q"))
    c.cx(0, 1); c.cz(1, 2); c.cx(2, 3); c.cz(3, 0)
    out["ring4"] = c
    c = QC(QR(3, "q"))
    c.cx(0, 2); c.cx(1, 2); c.cx(0, 2); c.cx(1, 2)
    out["star3"] = c
    c = QC(QR(5, "q"))
    for i in range(4):
        c.cx(i, 4)
    out["bv5like"] = c
    return out


@pytest.mark.parametrize("name,P,q,limits", [
    ("chain4", 2, 2, dict(maxNCuts=3, maxNQpdCuts=3)),
    ("chain4", 2, 3, dict()),
    ("ring4", 2, 2, dict(maxNCuts=4, maxNQpdCuts=4)),
    ("ring4", 2, 3, dict(maxNCuts=2, maxNQpdCuts=1)),          # teleport cuts become possible
    ("star3", 2, 2, dict(maxNCuts=4, maxNQpdCuts=4, maxCutsPerPartitions=4)),
    ("bv5like", 2, 3, dict(maxNCuts=5, maxNQpdCuts=5, maxCutsPerPartitions=5)),
    ("chain4", 3, 2, dict(maxNCuts=3)),
])
def test_optimum_equals_exhaustive_search(name, P, q, limits):
    c = _small_circuits()[name]
    cu = Cutter(c, P, q, **limits)
    want = _brute_force(cu, limits.get("maxNCuts"), limits.get("maxNQpdCuts"), limits.get("maxCutsPerPartitions"))
    ok = cu.solve()
    if want is None:
        assert not ok
        return
    assert ok
    S, A, L, n_w, n_g, Q, Q_p, C, C_p = cu.getModelKeyResults()
    assert (Q, S, A, L, C) == want[1:]
    assert Q == max(Q_p) and C == max(C_p)
    assert all(qp <= m for qp, m in zip(Q_p, cu.maxNQubitsPerPartition))


def test_infeasible_returns_false():
    c = _small_circuits()["ring4"]
    cu = Cutter(c, 2, 2, maxNCuts=1, maxNQpdCuts=1)
    assert _brute_force(cu, 1, 1) is None
    assert cu.solve() is False
    with pytest.raises(RuntimeError):
        cu.getModelKeyResults()


def test_graph_of_bv5():
    # Cutter.py:212-275: two vertices per two-qubit gate, wire edges between consecutive vertices of a qubit
    c = generators.gen_circ("bv", 5, 1)
    cu = Cutter(c, 2, 10)
    assert len(cu.V) == 8 and len(cu.G) == 4 and len(cu.W) == 3 and len(cu.I) == 5
    assert [v.qubit for v in cu.I] == [0, 4, 1, 2, 3]


def _knits_back(cut_circ, uncut, tol=1e-12):
    res, ov = oracle_knit(cut_circ)
    want = sv.exact_distribution(uncut)
    keys = set(res) | set(want)
    assert max(abs(res.get(k, 0.0) - want.get(k, 0.0)) for k in keys) < tol


def test_bv5_wire_cut_and_round_trip():
    # README.md:27-28 / SURVEY C.4: BV(5) results in one wire cut, Q = 3, S = 8
    c = generators.gen_circ("bv", 5, 1)
    cu = Cutter(c, 2, 10, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
    assert cu.solve()
    S, A, L, n_w, n_g, Q, Q_p, C, C_p = cu.getModelKeyResults()
    assert (S, A, L, n_w, n_g, Q, sorted(Q_p), C) == (8, 8, 0, 1, 0, 3, [3, 3], 1)
    spec = cu.cut_spec()
    assert spec.wire_cuts and spec.wire_cuts[0][0] == 4 and not spec.gate_cuts
    again = cutter_mod.cut_spec_from_json(cutter_mod.cut_spec_to_json(spec))
    assert again == spec
    cut = cu.getCutCirc()
    assert sorted(len(r) for r in cut.qregs) == [3, 3]
    _knits_back(cut, cu.decomposedCirc)


def test_gate_cut_circuit_knits_back():
    QC, QR = circuit.QuantumCircuit, circuit.QuantumRegister
    c = QC(QR(4, "q"))
    for q in range(4):
        c.ry(0.3 + 0.2 * q, q)
    c.cx(0, 1); c.cx(2, 3); c.cz(1, 2); c.rx(0.4, 1); c.ry(0.9, 2); c.cx(0, 1); c.cx(2, 3)
    c.measure_all()
    cu = Cutter(c, 2, 2, maxNCuts=2, maxNQpdCuts=2)
    assert cu.solve()
    S, A, L, n_w, n_g, Q, Q_p, C, C_p = cu.getModelKeyResults()
    assert (S, n_w, n_g, Q) == (6, 0, 1, 2)
    _knits_back(cu.getCutCirc(), cu.decomposedCirc)


def test_bv16_matches_survey_shape():
    # Q = 9, one wire cut on q15, S = 8, fragments of 9 and 8 qubits
    c = generators.gen_circ("bv", 16, 1)
    cu = Cutter(c, 2, 10, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
    assert cu.solve()
    S, A, L, n_w, n_g, Q, Q_p, C, C_p = cu.getModelKeyResults()
    assert (S, n_w, n_g, Q, sorted(Q_p)) == (8, 1, 0, 9, [8, 9])
    spec = cu.cut_spec()
    assert [w[0] for w in spec.wire_cuts] == [15]
    cut = cu.getCutCirc()
    assert sorted(len(r) for r in cut.qregs) == [8, 9]


def test_syc32d1_needs_no_cut():
    # already a tensor product - Q = 14, S = 1, leftovers go to the first partition: 18 | 14
    c = generators.gen_circ("syc", 32, 1, seed=0)
    cu = Cutter(c, 2, 50, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
    assert cu.solve()
    S, A, L, n_w, n_g, Q, Q_p, C, C_p = cu.getModelKeyResults()
    assert (S, A, L, n_w, n_g, Q, Q_p, C) == (1, 0, 0, 0, 0, 14, [14, 14], 0)
    spec = cu.cut_spec()
    assert [len(p) for p in spec.partitions] == [18, 14]
    assert sorted(q for p in spec.partitions for q in p) == list(range(32))
    cut = cu.getCutCirc()
    assert [len(r) for r in cut.qregs] == [18, 14]


def test_constructor_checks_follow_the_reference():
    c = _small_circuits()["chain4"]
    with pytest.raises(AssertionError):
        Cutter(c, 2, 1)                          # 4 qubits do not fit 2 x 1 (Cutter.py:62)
    with pytest.raises(RuntimeError):
        Cutter(c, 2, "10")
    with pytest.raises(AssertionError):
        Cutter(c, 2, [3, 3, 3])


# ---------------------------------------------------------------------- BASELINE.json configs
def _solve_config(config):
    name, n, depth, P, q = cutting.BASELINE_CONFIGS[config]
    circ = generators.gen_circ(name, n, depth, seed=0)
    cu = Cutter(circ, P, q, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)      # benchmarks/benchmark.py:41
    return cu, cu.solve()


@pytest.mark.parametrize("config,S,n_wire,n_gate,Q", [
    ("bv16", 8, 1, 0, 9), ("hwe16d5", 7776, 0, 5, 8), ("syc16d5", 1296, 0, 4, 8)])
def test_solver_reproduces_the_recorded_cut_shapes(config, S, n_wire, n_gate, Q, golden):
    cu, ok = _solve_config(config)
    assert ok
    res = cu.getModelKeyResults()
    assert (res[0], res[3], res[4], res[5]) == (S, n_wire, n_gate, Q)
    spec = cu.cut_spec()
    want = cutting.baseline_cut_spec(config, cu.decomposedCirc)       # the shapes bench.py and the GPU tests use
    assert sorted(spec.gate_cuts) == sorted(want.gate_cuts)
    assert [q for q, _ in spec.wire_cuts] == [q for q, _ in want.wire_cuts]
    # the cut circuit has the fragment sizes of Appendix B, whichever of the equivalent optima the solver returns
    sizes = sorted(len(r) for r in cu.getCutCirc().qregs)
    assert sizes == sorted(len(r) for r in cutting.apply_cuts(cu.decomposedCirc, want).qregs)
    # fixture written by tests/golden/make_cut_specs.py from this solver (the z3 search is cached)
    fixture = golden("cut_specs.json")[config]
    assert fixture["key_results"][:6] == list(res[:6])
    assert sorted(fixture["spec"]["gate_cuts"]) == sorted(spec.gate_cuts)


@pytest.mark.parametrize("config", ["qft16", "aqft16"])
def test_qft_family_is_infeasible_at_q10(config):
    # the reference itself exits at benchmarks/benchmark.py:53-54 for -q 10
    name, n, depth, P, _q = cutting.BASELINE_CONFIGS[config]
    cu = Cutter(generators.gen_circ(name, n, depth, seed=0), P, 10, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
    assert cu.solve() is False


def test_make_baseline_with_the_solver():
    circ, cut = cutting.make_baseline("hwe16d5", cut="solver")
    _circ, cut_table = cutting.make_baseline("hwe16d5")
    assert sorted(len(r) for r in cut.qregs) == sorted(len(r) for r in cut_table.qregs) == [8, 8]
    names = lambda c: sorted(type(i.operation).__name__ for i in c.data if len(i.qubits) == 2)
    assert names(cut) == names(cut_table)


def test_add6_solver_cut_knits_back():
    # Q is minimised before S (Cutter.py:567-568): even at -q 50 the 6-qubit adder is cut (two wire cuts on q2, 3 | 3
    # original qubits) - the resulting circuit must still reproduce the uncut distribution
    cu, ok = _solve_config("add6")
    assert ok
    S, A, L, n_w, n_g, Q, Q_p, C, C_p = cu.getModelKeyResults()
    assert (S, n_w, n_g, Q) == (64, 2, 0, 4)
    _knits_back(cu.getCutCirc(), cu.decomposedCirc, tol=1e-10)


@pytest.mark.parametrize("seed", range(16))
def test_random_small_circuits_against_exhaustive_search(seed):
    """Random circuits of 3-5 qubits (cx / cz / rzz; the cutter decomposes them to cx first, Cutter.py:84), random
    limits, 2 or 3 partitions: the optimum equals the exhaustive one."""
    import random
    rnd = random.Random(1000 + seed)
    n = rnd.randint(3, 5)
    P = rnd.choice([2, 2, 3])
    budget = 10 if P == 2 else 8                  # vertices the exhaustive search can afford
    c = circuit.QuantumCircuit(circuit.QuantumRegister(n, "q"))
    used = 0
    while True:
        kind = rnd.choice(["cx", "cx", "cz", "rzz"])
        cost = 4 if kind == "rzz" else 2          # rzz -> two cx
        if used + cost > budget:
            break
        a, b = rnd.sample(range(n), 2)
        if kind == "rzz":
            c.rzz(0.3, a, b)
        else:
            getattr(c, kind)(a, b)
        c.h(rnd.randrange(n))
        used += cost
    q = max(rnd.randint(2, n), -(-n // P))
    limits = {}
    if rnd.random() < 0.7:
        limits["maxNCuts"] = rnd.randint(1, 4)
        if rnd.random() < 0.7:
            limits["maxNQpdCuts"] = rnd.randint(0, limits["maxNCuts"])
    if rnd.random() < 0.4:
        limits["maxCutsPerPartitions"] = rnd.randint(1, 3)
    cu = Cutter(c, P, q, **limits)
    assert len(cu.V) <= budget
    want = _brute_force(cu, limits.get("maxNCuts"), limits.get("maxNQpdCuts"), limits.get("maxCutsPerPartitions"))
    ok = cu.solve()
    assert ok == (want is not None)
    if ok:
        S, A, L, n_w, n_g, Q, Q_p, C, C_p = cu.getModelKeyResults()
        assert (Q, S, A, L, C) == want[1:], (limits, P, q)


def test_get_result_circs_shapes():
    # Cutter.getResultCircs (Cutter.py:128-160): decomposed, marked, marked with moves, cut, instantiations
    c = generators.gen_circ("bv", 5, 1)
    cu = Cutter(c, 2, 10, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
    with pytest.raises(RuntimeError):
        cu.getResultCircs()
    assert cu.solve()
    dec, marked, with_moves, cut, inst = cu.getResultCircs(True)
    assert dec is cu.decomposedCirc
    names = lambda circ: [type(i.operation).__name__ for i in circ.data]
    assert names(marked).count("WireCut") == 1 and len(marked.data) == len(dec.data) + 1
    assert names(with_moves).count("VirtualMove") == 1 and with_moves.num_qubits == dec.num_qubits + 1
    assert names(cut) == names(with_moves)                      # same ops, qubits renamed into frag registers
    assert sorted(len(r) for r in cut.qregs) == [3, 3]
    assert sorted(len(x) for x in inst) == [8, 8]               # one VirtualMove: 8 instantiations on either side
    assert cu.getResultCircs()[4] == []
    # a gate cut: the virtual gate sits where the cx was
    QC, QR = circuit.QuantumCircuit, circuit.QuantumRegister
    c = QC(QR(4, "q"))
    c.cx(0, 1); c.cx(2, 3); c.cz(1, 2); c.cx(0, 1); c.cx(2, 3)
    c.measure_all()
    cu = Cutter(c, 2, 2, maxNCuts=2, maxNQpdCuts=2)
    assert cu.solve()
    dec, marked, with_moves, cut, inst = cu.getResultCircs(True)
    assert [n for n in names(marked) if n.startswith("Virtual")] == ["VirtualCX"]      # cz decomposes to h cx h
    assert sorted(len(x) for x in inst) == [6, 6]


def test_benchmark_driver_cut_only(tmp_path):
    # tools/benchmark.py keeps the reference's command line (benchmarks/benchmark.py:22-29)
    import importlib.util, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("qck_benchmark_driver", os.path.join(root, "tools", "benchmark.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(["-p", "2", "-q", "10", "bv", "5", "1", "--cut-only", "--out", str(tmp_path)]) == 0
    (run_dir,) = list(tmp_path.iterdir())
    cut = json.loads((run_dir / "cut_spec.json").read_text())
    assert cut["gate_cuts"] == [] and [w[0] for w in cut["wire_cuts"]] == [4]
    log = (run_dir / "run.log").read_text()
    assert "S: 8" in log and "nWireCuts: 1" in log and "success => True" in log
    # an infeasible request ends like the reference (benchmarks/benchmark.py:53-54): logged, exit code 0
    assert mod.main(["-p", "2", "-q", "10", "qft", "16", "1", "--cut-only", "--out", str(tmp_path)]) == 0
