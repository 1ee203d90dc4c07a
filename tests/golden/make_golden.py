#!/usr/bin/env python
"""Generate tests/golden/*.json by RUNNING THE REFERENCE'S OWN CODE.

Build-container only (needs /root/reference).  ``oracle/ref_loader.py`` imports
``third_party/qvm/qvm/virtual_gates.py`` and ``quasi_distr.py`` unmodified under a
qiskit stub; everything written here is an output of those modules:

instantiation_tables.json  ``_instantiations()`` of every virtual gate class
                           (name, qubit, measures?, params) + ``num_instantiations``
knit_cases.json            for seeded random sparse inputs: ``QuasiDistr.split / merge /
                           + / - / *``, every ``Virtual*.knit`` and
                           ``nearest_probability_distribution``, at ACCURACY = 1e-5
                           (the reference's value) and 0 (exact mode), including
                           values straddling the 1e-5 threshold and degenerate RZZ angles
semcheck.json              4-qubit circuits (one wire cut + one gate cut): per-instance
                           exact distributions (oracle simulator), then the REFERENCE's
                           merge + level-by-level knit; stored with the uncut distribution

Usage:  python tests/golden/make_golden.py
"""
import itertools
import json
import math
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import instantiate as oi  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402
from oracle import statevector as sv  # noqa: E402

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
ANGLES = [0.83, -1.3, 2.5, 0.0, math.pi, 2 * math.pi, 3e-6, math.pi - 4e-6]


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as fh:
        json.dump(obj, fh, indent=0, separators=(",", ":"))
    print("wrote", name, os.path.getsize(os.path.join(HERE, name)), "bytes")


def jd(d):
    """dict[int, float] -> JSON-able list of [key, value] in insertion order."""
    return [[int(k), float(v)] for k, v in d.items()]


def tables(vg):
    out = []
    for kind in ("move", "cz", "cx", "cy"):
        g = rl.make_vgate(vg, kind)
        out.append({"kind": kind, "theta": None, "n": g.num_instantiations, "table": rl.dump_table(g)})
    for kind in ("rzz", "cp"):
        for th in ANGLES:
            g = rl.make_vgate(vg, kind, th)
            out.append({"kind": kind, "theta": th, "n": g.num_instantiations, "table": rl.dump_table(g),
                        "params_after_init": [float(p) for p in g._params]})
    return out


def rand_distr(rng, nbits, n, scale=1.0, signed=True, near_threshold=False):
    d = {}
    for _ in range(n):
        k = rng.randrange(1 << nbits)
        v = rng.uniform(-1, 1) if signed else rng.random()
        v *= scale
        if near_threshold and rng.random() < 0.4:
            v = math.copysign(1e-5 * rng.choice([0.5, 0.9, 0.999, 1.0, 1.001, 1.1, 2.0]), v)
        d[k] = v
    return d


def knit_cases(vg, qd):
    rng = random.Random(1234)
    cases = {"ops": [], "knit": [], "npd": []}
    for acc in (1e-5, 0.0):
        qd.ACCURACY = acc
        for trial in range(12):
            nb = rng.choice([3, 5, 8])
            a_raw = rand_distr(rng, nb, 12, near_threshold=(trial % 2 == 0))
            b_raw = rand_distr(rng, nb, 12, scale=rng.choice([1.0, 1e-2, 3e-3]), near_threshold=(trial % 3 == 0))
            a, b = qd.QuasiDistr(a_raw), qd.QuasiDistr(b_raw)
            bit = rng.randrange(nb)
            lo, hi = a.split(bit)
            # merge needs disjoint supports: shift b above a
            b_shift = qd.QuasiDistr({k << nb: v for k, v in b_raw.items()})
            s = rng.uniform(-2, 2)
            cases["ops"].append({
                "acc": acc, "nbits": nb, "a_raw": jd(a_raw), "b_raw": jd(b_raw), "a": jd(a), "b": jd(b),
                "bit": bit, "split_lo": jd(lo), "split_hi": jd(hi), "add": jd(a + b), "sub": jd(a - b),
                "scale": s, "mul": jd(a * s), "rmul": jd(s * a), "merge": jd(a.merge(b_shift)),
            })
        # per-gate knit on random inputs; config bit = top bit
        specs = [("move", None), ("cz", None), ("cx", None), ("cy", None)] + \
                [(k, th) for k in ("rzz", "cp") for th in ANGLES]
        for kind, th in specs:
            for trial in range(3):
                g = rl.make_vgate(vg, kind, th)
                nb = 4
                clbit = nb
                results_raw = [rand_distr(rng, nb + 1, 14, scale=rng.choice([1.0, 1e-3]),
                                          near_threshold=(trial == 0)) for _ in range(g.num_instantiations)]
                results = [qd.QuasiDistr(r) for r in results_raw]
                out = g.knit(results, clbit)
                cases["knit"].append({"acc": acc, "kind": kind, "theta": th, "nbits": nb + 1, "clbit": clbit,
                                      "results_raw": [jd(r) for r in results_raw], "out": jd(out)})
        for trial in range(10):
            raw = rand_distr(rng, 6, rng.choice([1, 5, 20, 40]), signed=True)
            if trial % 3 == 0:
                raw = {k: abs(v) for k, v in raw.items()}
            if trial % 3 == 1:      # mostly positive with a few small negatives (the realistic case)
                raw = {k: (abs(v) if i % 4 else -abs(v) * 1e-2) for i, (k, v) in enumerate(raw.items())}
            q = qd.QuasiDistr(raw)
            if sum(q.values()) <= 0:
                continue
            try:
                out = q.nearest_probability_distribution()
            except ZeroDivisionError:
                continue
            cases["npd"].append({"acc": acc, "nbits": 6, "raw": jd(raw), "out": jd(out)})
    qd.ACCURACY = 1e-5
    return cases


def semcheck(vg, qd):
    from importlib import import_module
    circuit = import_module(f"{PKG}.circuit")
    cutting = import_module(f"{PKG}.cutting")
    out = []
    for gname, theta in [("cx", None), ("cz", None), ("cy", None), ("rzz", 0.83), ("cp", 0.83)]:
        qc = circuit.QuantumCircuit(circuit.QuantumRegister(4, "q"))
        for q in range(4):
            qc.ry(0.3 + 0.4 * q, q)
            qc.rz(0.2 * q + 0.1, q)
        qc.cx(0, 1); qc.h(1); qc.cx(1, 2); qc.rx(0.7, 1); qc.cx(2, 3)
        if theta is None:
            getattr(qc, gname)(0, 3)
        else:
            getattr(qc, gname)(theta, 0, 3)
        qc.ry(0.5, 0); qc.rx(0.4, 3); qc.h(2)
        qc.measure_all()
        gidx = [i for i, ins in enumerate(qc.data) if ins.operation.name == gname][-1]
        cut = cutting.apply_cuts(qc, cutting.CutSpec(gate_cuts=[gidx], wire_cuts=[(1, 8)]))
        ov = oi.OracleVirtualCircuit(cut)
        for acc in (0.0, 1e-5):
            qd.ACCURACY = acc
            frag_results = []
            per_instance = []
            for frag in ov.fragments:
                labels = ov.instance_labels(frag)
                dists = [sv.exact_distribution(ov.instance(frag, l)) for l in labels]
                per_instance.append({"labels": [list(l) for l in labels], "dists": [jd(d) for d in dists]})
                # what run.py:56 does with the backend's counts (exact probabilities as "counts")
                by_label = dict(zip(labels, [qd.QuasiDistr(d) for d in dists]))
                touch = ov.touches(frag)
                frag_results.append([by_label[tuple(g[k] if touch[k] else -1 for k in range(len(g)))]
                                     for g in ov.global_labels()])
            # virtual_circuit.py:216-228 / 50-68 driven by hand, all arithmetic by the reference classes
            merged = []
            for group in zip(*frag_results):
                m = group[0]
                for other in group[1:]:
                    m = m.merge(other)
                merged.append(m)
            ref_gates = [rl.make_vgate(vg, kind, th) for kind, th, _ in ov.vgates]
            clbit = ov.n_clbits + len(ref_gates) - 1
            for g in reversed(ref_gates):
                n = g.num_instantiations
                merged = [g.knit(merged[i:i + n], clbit) for i in range(0, len(merged), n)]
                clbit -= 1
            knitted = merged[0]
            out.append({
                "gate": gname, "theta": theta, "acc": acc, "n_clbits": ov.n_clbits,
                "vgates": [[k, th] for k, th, _ in ov.vgates],
                "fragments": per_instance,
                "knit": jd(knitted),
                "npd": jd(knitted.nearest_probability_distribution()),
                "uncut": jd(sv.exact_distribution(qc)),
            })
    qd.ACCURACY = 1e-5
    return out


def main():
    if not rl.available():
        raise SystemExit("needs /root/reference")
    vg, qd = rl.load()
    dump("instantiation_tables.json", tables(vg))
    dump("knit_cases.json", knit_cases(vg, qd))
    dump("semcheck.json", semcheck(vg, qd))
    labels = {}
    for radices in ([8], [6, 6], [8, 6], [6] * 5, [1, 6, 8]):
        labels["x".join(map(str, radices))] = [list(t) for t in itertools.product(*[range(r) for r in radices])][:4000]
    dump("label_enumeration.json", labels)


if __name__ == "__main__":
    main()
