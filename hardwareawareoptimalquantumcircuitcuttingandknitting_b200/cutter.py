"""Qiskit-free restatement of the hardware-aware cutter (host side).

``src/HwAwareCutter/Cutter.py:38-571`` chooses gate cuts and wire cuts with a z3 model so that the circuit
splits into at most ``maxNPartitions`` partitions of at most ``maxNQubitsPerPartition`` qubits.  This module
restates that model on this package's circuit IR so that cut specs for arbitrary circuits can be produced on an
image without qiskit, and hands the result to ``cutting.apply_cuts`` (which builds what
``Cutter.getResultCircs`` gives qvm).  It is host-only input preparation: nothing here runs on the hot path.

Differences from the reference, both forced by the environment:

* the reference solves with ``z3.Optimize`` and five ``minimize`` objectives; z3 4.15.4 (this image) returns
  constraint-violating models for exactly this pattern, so the same lexicographic optimum (soft constraint
  first, then Q, S, A, L, C - the order the objectives are added, ``Cutter.py:553-567``) is found by iterative
  tightening on z3's finite-domain solver.  Every integer of the reference's model is a sum of 0/1 terms, so the
  model is stated over Booleans with pseudo-Boolean cardinality constraints; S = 6^a 8^b is never multiplied out
  but searched over the (a, b) pairs in ascending order of their product.  Solve times (this container, one
  core): bv-16 0.5 s, hwe-16 d5 1.4 s, syc-16 d5 0.8 s, qft-16 -q 10 proved infeasible in 3.3 s (the same
  search over ``Int`` sums took 16 min for hwe-16 d5; the reference reports up to 12 min);
* teleportation cuts (``b_e``) are part of the model as in the reference, but - as there (``Cutter.py:574``
  FIXME) - cannot be turned into a circuit: ``cut_spec()`` raises if the optimum uses one.

Graph (``Cutter.py:212-275``): every two-qubit gate contributes two vertices (one per qubit); the pair is a
gate-cut edge, consecutive vertices on a wire form a wire-cut edge; I = first vertex of every wire.
"""
from __future__ import annotations

import json
from dataclasses import dataclass

from .circuit import Barrier, Gate, QuantumCircuit
from .cutting import CutSpec, apply_cuts
from .virtual_gates import VIRTUAL_GATE_TYPES

__all__ = ["Cutter", "cut_spec_to_json", "cut_spec_from_json"]

# Cutter.py:453-472
_GATE_QPD = (6, 0, 0)       # (overhead sampling, ancilla, teleport latency)
_WIRE_QPD = (8, 1, 0)
_GATE_TELE = (1, 2, 10)
_WIRE_TELE = (1, 2, 10)


@dataclass
class _Vertex:
    idx: int
    qubit: int           # index into circuit.qubits
    data_idx: int        # instruction the vertex belongs to


class Cutter:
    """Same constructor arguments, ``solve()`` and ``getModelKeyResults()`` as the reference class."""

    def __init__(self, inputCirc: QuantumCircuit, maxNPartitions: int = 2, maxNQubitsPerPartition=10,
                 forceNWireCuts=None, forceNGateCuts=None, maxNQpdCuts=None, maxNCuts=None,
                 maxCutsPerPartitions=None) -> None:
        import z3
        self.z3 = z3
        self.maxNPartitions = int(maxNPartitions)
        if isinstance(maxNQubitsPerPartition, int):
            self.maxNQubitsPerPartition = [maxNQubitsPerPartition] * self.maxNPartitions
        elif isinstance(maxNQubitsPerPartition, list):
            self.maxNQubitsPerPartition = list(maxNQubitsPerPartition)
        else:
            raise RuntimeError("Invalid type")
        assert len(self.maxNQubitsPerPartition) == self.maxNPartitions
        assert len(inputCirc.qubits) <= sum(self.maxNQubitsPerPartition)
        assert forceNWireCuts is None or forceNWireCuts >= 0
        assert forceNGateCuts is None or forceNGateCuts >= 0
        if maxNCuts is not None:
            assert maxNCuts > 0 and maxNCuts >= (forceNWireCuts or 0) + (forceNGateCuts or 0)
        if maxNQpdCuts is not None:
            assert maxNQpdCuts >= 0 and (maxNCuts is None or maxNQpdCuts <= maxNCuts)
        assert maxCutsPerPartitions is None or maxCutsPerPartitions > 0
        self.forceNWireCuts, self.forceNGateCuts = forceNWireCuts, forceNGateCuts
        self.maxNCuts, self.maxNQpdCuts, self.maxCutsPerPartitions = maxNCuts, maxNQpdCuts, maxCutsPerPartitions
        # Cutter.py:84 decomposes once: afterwards cx is the only two-qubit gate
        self.decomposedCirc = inputCirc.decompose_two_qubit()
        self.V, self.W, self.G, self.I = self._read_circ(self.decomposedCirc)
        self.model = None
        self.nWireCuts = self.nGateCuts = 0
        self.s = z3.SolverFor("QF_FD")
        self._build_model()

    # ------------------------------------------------------------------ graph (Cutter.py:212-294)
    def _read_circ(self, circ: QuantumCircuit):
        qpos = {q: i for i, q in enumerate(circ.qubits)}
        V: list[_Vertex] = []
        W: list[tuple[int, int]] = []
        G: list[tuple[int, int]] = []
        I: list[_Vertex] = []
        prev: dict[int, int] = {}
        for di, ins in enumerate(circ.data):
            op = ins.operation
            if len(ins.qubits) != 2 or isinstance(op, Barrier):       # barriers, virtual gates, VirtualMove
                continue
            q0, q1 = (qpos[q] for q in ins.qubits)
            v0, v1 = len(V), len(V) + 1
            V.append(_Vertex(v0, q0, di))
            V.append(_Vertex(v1, q1, di))
            G.append((v0, v1))
            for q, v in ((q0, v0), (q1, v1)):
                if q in prev:
                    W.append((prev[q], v))
                else:
                    I.append(V[v])
                prev[q] = v
        for u, v in W + G:
            assert u < v < len(V)
        return V, W, G, I

    # ------------------------------------------------------------------ model (Cutter.py:297-567)
    def _build_model(self) -> None:
        """The reference's constraints over Booleans only: every integer of its model (Q_p, C_p, the cut counts, A, L)
        is a sum of 0/1 terms, so ``Q_p <= cap`` etc. become pseudo-Boolean constraints (``PbLe`` / ``PbEq``) and the
        objectives are minimised by tightening such bounds.  (With ``Int`` sums the same search took 16 min for
        hwe-16 d5; the finite-domain solver needs seconds.)"""
        z3, s, P = self.z3, self.s, self.maxNPartitions
        assert P <= max(len(self.V), 1)
        self.o = [[z3.Bool(f"o_{v}_{p}") for p in range(P)] for v in range(len(self.V))]
        self.c_e, self.b_e, self.edges = [], [], []      # edges: (u, v, "W" | "G")
        for e, (u, v) in enumerate(self.W):
            self.c_e.append(z3.Bool(f"c_{e}[W]_{u}_{v}"))
            self.b_e.append(z3.Bool(f"b_{e}[W]_{u}_{v}"))
            self.edges.append((u, v, "W"))
        for e, (u, v) in enumerate(self.G):
            op = self.decomposedCirc.data[self.V[u].data_idx].operation
            if not isinstance(op, Gate) or op.name not in VIRTUAL_GATE_TYPES:
                continue                                  # not virtualisable: never cut (Cutter.py:353-356)
            self.c_e.append(z3.Bool(f"c_{e}[G]_{u}_{v}"))
            self.b_e.append(z3.Bool(f"b_{e}[G]_{u}_{v}"))
            self.edges.append((u, v, "G"))
        o, c_e, b_e, edges = self.o, self.c_e, self.b_e, self.edges
        n_aux = [0]

        def lit(expr):
            """A fresh Boolean equal to ``expr`` (pseudo-Boolean constraints count literals)."""
            n_aux[0] += 1
            x = z3.Bool(f"aux_{n_aux[0]}")
            s.add(x == expr)
            return x

        for i, (u, v, _t) in enumerate(edges):
            s.add(c_e[i] == z3.Or([o[u][p] != o[v][p] for p in range(P)]))
            s.add(z3.Implies(b_e[i], c_e[i]))
        for v in range(len(self.V)):                      # exactly one partition per vertex (Cutter.py:396-410)
            s.add(z3.PbEq([(x, 1) for x in o[v]], 1))
        # partitions of equal capacity are interchangeable: pin the first vertex to partition 0, and let vertex v
        # use partition p only if an earlier vertex uses p - 1 (halves the search at P = 2; the optimum is unchanged)
        if self.V and len(set(self.maxNQubitsPerPartition)) == 1:
            s.add(o[0][0])
            for v in range(1, len(self.V)):
                for p in range(1, P):
                    if p > v:
                        s.add(z3.Not(o[v][p]))
                    elif p >= 2:
                        s.add(z3.Implies(o[v][p], z3.Or([o[u][p - 1] for u in range(v)])))
        # teleportation needs exactly maxNQpdCuts QPD cuts besides itself (Cutter.py:536-540): impossible when that
        # already exhausts maxNCuts (the limits of benchmarks/benchmark.py:41)
        self.teleport_possible = not (self.maxNQpdCuts is not None and self.maxNCuts is not None
                                      and self.maxNQpdCuts >= self.maxNCuts)
        if not self.teleport_possible:
            s.add([z3.Not(b) for b in b_e])
        self.qpd = [lit(z3.And(c, z3.Not(b))) if self.teleport_possible else c for c, b in zip(c_e, b_e)]
        self.gate_qpd = [x for x, e in zip(self.qpd, edges) if e[2] == "G"]
        self.wire_qpd = [x for x, e in zip(self.qpd, edges) if e[2] == "W"]
        # Q_p (Cutter.py:412-439) and C_p (:441-451) as lists of literals
        self.Qp_terms, self.Cp_terms = [], []
        for p in range(P):
            terms = [o[v.idx][p] for v in self.I]
            terms += [lit(z3.And(c_e[i], o[v][p])) for i, (u, v, t) in enumerate(edges) if t == "W"]
            if self.teleport_possible:
                terms += [lit(z3.And(b_e[i], z3.Or(o[u][p], o[v][p]))) for i, (u, v, t) in enumerate(edges)]
            self.Qp_terms.append(terms)
            self.Cp_terms.append([lit(z3.And(self.qpd[i], z3.Or(o[u][p], o[v][p]))) for i, (u, v, t) in enumerate(edges)])
            if terms:
                s.add(z3.PbLe([(x, 1) for x in terms], self.maxNQubitsPerPartition[p]))
            if self.maxCutsPerPartitions is not None and self.Cp_terms[p]:
                s.add(z3.PbLe([(x, 1) for x in self.Cp_terms[p]], self.maxCutsPerPartitions))
        wire = [(c, 1) for c, e in zip(c_e, edges) if e[2] == "W"]
        gate = [(c, 1) for c, e in zip(c_e, edges) if e[2] == "G"]
        if self.forceNWireCuts is not None:
            s.add(z3.PbEq(wire, self.forceNWireCuts) if wire else z3.BoolVal(self.forceNWireCuts == 0))
        if self.forceNGateCuts is not None:
            s.add(z3.PbEq(gate, self.forceNGateCuts) if gate else z3.BoolVal(self.forceNGateCuts == 0))
        if self.maxNCuts is not None and wire + gate:
            s.add(z3.PbLe(wire + gate, self.maxNCuts))
        if self.maxNQpdCuts is not None and self.qpd:
            s.add(z3.PbLe([(x, 1) for x in self.qpd], self.maxNQpdCuts))
            if self.teleport_possible:
                full = lit(z3.PbEq([(x, 1) for x in self.qpd], self.maxNQpdCuts))
                s.add([z3.Implies(b, full) for b in b_e])
        # ancilla and latency sums (Cutter.py:453-510): A = S * (2 per teleport cut + 1 per QPD wire cut), L = 10 per
        # teleport cut
        self.A_terms = [(x, _WIRE_QPD[1]) for x in self.wire_qpd]
        self.L_terms = []
        if self.teleport_possible:
            self.A_terms += [(b, _GATE_TELE[1]) for b in b_e]
            self.L_terms = [(b, _GATE_TELE[2]) for b in b_e]
        # soft constraint (Cutter.py:545-551): every QPD cut lies before every teleport cut, i.e. no pair
        # (QPD cut i, teleport cut j) with v_i >= u_j
        if self.teleport_possible and edges:
            self.soft = z3.And([z3.Not(z3.And(self.qpd[i], b_e[j]))
                                for i in range(len(edges)) for j in range(len(edges)) if edges[i][1] >= edges[j][0]]
                               or [z3.BoolVal(True)])
        else:
            self.soft = z3.BoolVal(True)

    # ------------------------------------------------------------------ solving
    def _true(self, x) -> bool:
        return self.z3.is_true(self._m.eval(x, model_completion=True))

    def _count(self, terms) -> int:
        return sum(1 for x in terms if self._true(x))

    def _weight(self, terms) -> int:
        return sum(w for x, w in terms if self._true(x))

    def _tighten(self, value, bound, lower: int = 0) -> int:
        """Smallest value of an objective: ``value()`` reads it off the current model ``self._m``, ``bound(k)`` is the
        constraint "objective <= k".  The optimum is asserted and returned; ``self._m`` stays a model of everything
        asserted so far, so no re-check is needed."""
        z3, s = self.z3, self.s
        best = value()
        while best > lower:                               # ``lower``: a bound known without search
            s.push()
            s.add(bound(best - 1))
            sat = s.check() == z3.sat
            if sat:
                self._m = s.model()
                best = value()
            s.pop()
            if not sat:
                break
        s.add(bound(best))
        return best

    def solve(self) -> bool:
        z3, s = self.z3, self.s
        self.model = None
        self.nWireCuts = self.nGateCuts = 0
        if s.check() != z3.sat:
            return False
        self._m = s.model()
        if self.teleport_possible:                        # the soft constraint comes first in the lexicographic order
            s.push()
            s.add(self.soft)
            if s.check() == z3.sat:
                self._m = s.model()
            else:
                s.pop()
        P = self.maxNPartitions
        # Q = max_p Q_p (Cutter.py:512-514, objective :567)
        # (every wire starts in exactly one partition: Q >= ceil(|I| / P) needs no proof by search)
        self._Q = self._tighten(lambda: max(self._count(t) for t in self.Qp_terms),
                                lambda k: z3.And([z3.PbLe([(x, 1) for x in t], k) for t in self.Qp_terms if t]),
                                lower=-(-len(self.I) // P))
        # S (objective :568)
        self._S = self._solve_S()
        # A = S * ancillas (:569), L (:570), C = max_p C_p (:571)
        if self.A_terms:
            self._A = self._S * self._tighten(lambda: self._weight(self.A_terms), lambda k: z3.PbLe(self.A_terms, k))
        else:
            self._A = 0
        self._L = self._tighten(lambda: self._weight(self.L_terms), lambda k: z3.PbLe(self.L_terms, k)) \
            if self.L_terms else 0
        self._C = self._tighten(lambda: max(self._count(t) for t in self.Cp_terms),
                                lambda k: z3.And([z3.PbLe([(x, 1) for x in t], k) for t in self.Cp_terms if t])) \
            if any(self.Cp_terms) else 0
        self.model = self._m
        for c, (_u, _v, t) in zip(self.c_e, self.edges):
            if self._true(c):
                if t == "W":
                    self.nWireCuts += 1
                else:
                    self.nGateCuts += 1
        return True

    def _solve_S(self) -> int:
        z3, s = self.z3, self.s
        if not self.qpd:
            return 1
        ones = [(x, 1) for x in self.qpd]
        # the lower bound on the number of cuts is a consequence, not an objective: prove it, keep it, but do not
        # fix the count (a larger count can have a smaller S: one wire cut = 8 < two gate cuts = 36)
        t_min = self._count(self.qpd)
        model = self._m
        while t_min > 0:
            s.push()
            s.add(z3.PbLe(ones, t_min - 1))
            sat = s.check() == z3.sat
            if sat:
                model = s.model()
                t_min = sum(1 for x in self.qpd if z3.is_true(model.eval(x, model_completion=True)))
            s.pop()
            if not sat:
                break
        s.add(z3.PbGe(ones, t_min))
        n_g, n_w = len(self.gate_qpd), len(self.wire_qpd)
        bound = min(x for x in (self.maxNQpdCuts, self.maxNCuts, n_g + n_w) if x is not None)
        all_pairs = [(a, b) for a in range(min(n_g, bound) + 1) for b in range(min(n_w, bound) + 1)
                     if t_min <= a + b <= bound]

        def exactly(a, b):
            cs = []
            if self.gate_qpd:
                cs.append(z3.PbEq([(x, 1) for x in self.gate_qpd], a))
            if self.wire_qpd:
                cs.append(z3.PbEq([(x, 1) for x in self.wire_qpd], b))
            return z3.And(cs)

        for v in sorted({6 ** a * 8 ** b for a, b in all_pairs}):
            choice = z3.Or([exactly(a, b) for a, b in all_pairs if 6 ** a * 8 ** b == v])
            s.push()
            s.add(choice)
            sat = s.check() == z3.sat
            if sat:
                self._m = s.model()
            s.pop()
            if sat:
                s.add(choice)
                return v
        raise RuntimeError("no feasible sampling overhead")   # cannot happen: the constraints were satisfiable

    def getModelKeyResults(self):
        """-> S, A, L, nWireCuts, nGateCuts, Q, [Q_p], C, [C_p] (Cutter.py:162-178)."""
        if self.model is None:
            raise RuntimeError("no model exists")
        return (self._S, self._A, self._L, self.nWireCuts, self.nGateCuts, self._Q,
                [self._count(t) for t in self.Qp_terms], self._C, [self._count(t) for t in self.Cp_terms])

    # ------------------------------------------------------------------ results
    def cut_spec(self) -> CutSpec:
        """The chosen cuts for ``cutting.apply_cuts(self.decomposedCirc, spec)``."""
        if self.model is None:
            raise RuntimeError("no model exists")
        z3 = self.z3
        true = lambda x: z3.is_true(self.model.eval(x, model_completion=True))
        gate_cuts, wire_cuts = [], []
        for c, b, (u, v, t) in zip(self.c_e, self.b_e, self.edges):
            if not true(c):
                continue
            if true(b):
                raise NotImplementedError("the optimum uses a teleportation cut, which the reference cannot turn into a "
                                          "circuit either (Cutter.py:574)")
            if t == "G":
                gate_cuts.append(self.V[u].data_idx)
            else:                                         # cut the wire right after u's gate
                wire_cuts.append((self.V[u].qubit, self.V[u].data_idx))
        part_of_vertex = {}
        for v in range(len(self.V)):
            for p in range(self.maxNPartitions):
                if true(self.o[v][p]):
                    part_of_vertex[v] = p
        n_q = len(self.decomposedCirc.qubits)
        parts: list[list[int]] = [[] for _ in range(self.maxNPartitions)]
        placed = set()
        for v in self.I:                                  # a wire starts in the partition of its first vertex
            parts[part_of_vertex[v.idx]].append(v.qubit)
            placed.add(v.qubit)
        # wires that move after a cut occupy a slot in the partition they move to (the vmove qubit)
        used = [len(p) for p in parts]
        for c, (u, v, t) in zip(self.c_e, self.edges):
            if t == "W" and true(c):
                used[part_of_vertex[v]] += 1
        left = [q for q in range(n_q) if q not in placed]  # qubits without two-qubit gates (Cutter.py:672-700)
        if sum(self.maxNQubitsPerPartition) - sum(used) < len(left):
            raise RuntimeError("not enough available spots")
        for p in range(self.maxNPartitions):
            while left and used[p] < self.maxNQubitsPerPartition[p]:
                parts[p].append(left.pop(0))
                used[p] += 1
        return CutSpec(gate_cuts=sorted(gate_cuts), wire_cuts=sorted(wire_cuts, key=lambda t: (t[1], t[0])),
                       partitions=[sorted(p) for p in parts if p])

    def getResultCircs(self, getInstantiations: bool = False):
        """-> (decomposed circuit, cut-marked circuit, cut-marked circuit with VirtualMoves, cut circuit,
        instantiated circuits) as ``Cutter.getResultCircs`` (``Cutter.py:128-160``).  The two marked circuits are for
        display there: here the first carries the virtual gates and a ``WireCut`` marker after every cut position
        on the original register, the second is the cut circuit before its qubits are renamed into ``frag*``
        registers (one register ``q`` plus ``vmove``).  Instantiations: one list per fragment, in
        ``get_instance_labels`` order (``Cutter.py:702-707``)."""
        from .circuit import CircuitInstruction, QuantumRegister
        from .virtual_gates import VirtualMove, WireCut
        spec = self.cut_spec()
        circ = self.decomposedCirc
        gate_cuts = set(spec.gate_cuts)
        wire_after: dict[int, list[int]] = {}
        for q, idx in spec.wire_cuts:
            wire_after.setdefault(idx, []).append(q)
        marked = QuantumCircuit(*circ.qregs, *circ.cregs, name=circ.name)
        moves_reg = QuantumRegister(len(spec.wire_cuts), "vmove")
        with_moves = QuantumCircuit(*circ.qregs, moves_reg, *circ.cregs, name=circ.name)
        current = {q: q for q in circ.qubits}
        n_moves = 0
        for idx, ins in enumerate(circ.data):
            op = ins.operation
            if idx in gate_cuts:
                op = VIRTUAL_GATE_TYPES[op.name](op, f"{op.name} {getattr(op, 'label', None)}")
            marked.data.append(CircuitInstruction(op, ins.qubits, ins.clbits))
            with_moves.data.append(CircuitInstruction(op, tuple(current[q] for q in ins.qubits), ins.clbits))
            for qi in wire_after.get(idx, ()):
                q = circ.qubits[qi]
                marked.data.append(CircuitInstruction(WireCut(1, f"{idx}_{qi}"), (q,), ()))
                target = moves_reg[n_moves]
                with_moves.data.append(CircuitInstruction(VirtualMove(Gate("swap", 2, (), label=f"{idx}_{qi}")),
                                                          (current[q], target), ()))
                current[q] = target
                n_moves += 1
        cut = apply_cuts(circ, spec)
        instantiations = []
        if getInstantiations:
            from .virtual_circuit import VirtualCircuit, generate_instantiations
            virt = VirtualCircuit(cut.copy())
            for frag, frag_circ in virt.fragment_circuits.items():
                instantiations.append(generate_instantiations(frag_circ, virt.get_instance_labels(frag)))
        return circ, marked, with_moves, cut, instantiations

    def getCutCirc(self) -> QuantumCircuit:
        """The circuit ``Cutter.getResultCircs`` hands to qvm: ``frag*`` registers, virtual gates, VirtualMoves."""
        return apply_cuts(self.decomposedCirc, self.cut_spec())


# ---------------------------------------------------------------------- cut-spec wire format
def cut_spec_to_json(spec: CutSpec) -> str:
    return json.dumps({"gate_cuts": list(spec.gate_cuts), "wire_cuts": [list(w) for w in spec.wire_cuts],
                       "partitions": spec.partitions})


def cut_spec_from_json(text: str) -> CutSpec:
    d = json.loads(text)
    return CutSpec(gate_cuts=[int(x) for x in d.get("gate_cuts", [])],
                   wire_cuts=[(int(q), int(i)) for q, i in d.get("wire_cuts", [])],
                   partitions=None if d.get("partitions") is None else [[int(q) for q in p] for p in d["partitions"]])
