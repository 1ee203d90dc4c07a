"""Qiskit-free restatement of the hardware-aware cutter (host side).

``src/HwAwareCutter/Cutter.py:38-571`` chooses gate cuts and wire cuts with a z3 model so that the circuit
splits into at most ``maxNPartitions`` partitions of at most ``maxNQubitsPerPartition`` qubits.  This module
restates that model on this package's circuit IR so that cut specs for arbitrary circuits can be produced on an
image without qiskit, and hands the result to ``cutting.apply_cuts`` (which builds what
``Cutter.getResultCircs`` gives qvm).  It is host-only input preparation: nothing here runs on the hot path.

Differences from the reference, both forced by the environment:

* the reference solves with ``z3.Optimize`` and five ``minimize`` objectives; z3 4.15.4 (this image) returns
  constraint-violating models for exactly this pattern, so the same lexicographic optimum (soft constraint
  first, then Q, S, A, L, C - the order the objectives are added, ``Cutter.py:553-567``) is found with a plain
  ``z3.Solver`` and iterative tightening;
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
        self.s = z3.Solver()
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
        self.Q_p = [z3.Int(f"Q_p{p}") for p in range(P)]
        self.C_p = [z3.Int(f"C_p{p}") for p in range(P)]
        self.Q, self.S, self.A, self.L, self.C = (z3.Int(n) for n in "QSALC")
        o, c_e, b_e = self.o, self.c_e, self.b_e
        for i, (u, v, _t) in enumerate(self.edges):
            s.add(c_e[i] == z3.Or([o[u][p] != o[v][p] for p in range(P)]))
            s.add(z3.Implies(b_e[i], c_e[i]))
        for v in range(len(self.V)):                      # exactly one partition per vertex
            for i in range(P):
                for j in range(P):
                    if i != j:
                        s.add(z3.Implies(o[v][i], z3.Not(o[v][j])))
            s.add(z3.Or(o[v]))
            # redundant, for the arithmetic solver only: it cannot see "exactly one" through the implications
            s.add(z3.Sum([z3.If(o[v][p], 1, 0) for p in range(P)]) == 1)
        if self.I:                                        # every wire starts in exactly one partition (pigeonhole bound)
            s.add(self.Q >= -(-len(self.I) // P))
        for p in range(P):
            terms = [z3.If(o[v.idx][p], 1, 0) for v in self.I]
            terms += [z3.If(z3.And(c_e[i], o[v][p]), 1, 0) for i, (u, v, t) in enumerate(self.edges) if t == "W"]
            terms += [z3.If(z3.And(b_e[i], z3.Or(o[u][p], o[v][p])), 1, 0) for i, (u, v, t) in enumerate(self.edges)]
            s.add(self.Q_p[p] == z3.Sum(terms) if terms else self.Q_p[p] == 0)
            cuts = [z3.If(z3.And(c_e[i], z3.Or(o[u][p], o[v][p]), z3.Not(b_e[i])), 1, 0)
                    for i, (u, v, t) in enumerate(self.edges)]
            s.add(self.C_p[p] == z3.Sum(cuts) if cuts else self.C_p[p] == 0)
        total_s, total_a, total_l = z3.IntVal(1), z3.IntVal(0), z3.IntVal(0)
        for i, (_u, _v, t) in enumerate(self.edges):
            qpd, tele = (_GATE_QPD, _GATE_TELE) if t == "G" else (_WIRE_QPD, _WIRE_TELE)
            total_s = total_s * z3.If(c_e[i], z3.If(b_e[i], tele[0], qpd[0]), 1)
            total_a = total_a + z3.If(c_e[i], z3.If(b_e[i], tele[1], qpd[1]), 0)
            total_l = total_l + z3.If(c_e[i], z3.If(b_e[i], tele[2], qpd[2]), 0)
        # S (a product) is NOT asserted here: the definitions of S, A and L do not restrict the cuts, and the
        # non-linear product makes every check slow.  solve() minimises Q on the linear part first and then
        # walks the possible values of S = 6^a 8^b in ascending order (see _minimise_S).
        self._total_s, self._total_a, self._total_l = total_s, total_a, total_l
        self._n_gate_qpd = z3.Sum([z3.If(z3.And(c, z3.Not(b)), 1, 0) for c, b, (_u, _v, t) in zip(c_e, b_e, self.edges)
                                   if t == "G"] or [z3.IntVal(0)])
        self._n_wire_qpd = z3.Sum([z3.If(z3.And(c, z3.Not(b)), 1, 0) for c, b, (_u, _v, t) in zip(c_e, b_e, self.edges)
                                   if t == "W"] or [z3.IntVal(0)])
        for p in range(P):
            s.add(self.Q >= self.Q_p[p], self.Q_p[p] <= self.maxNQubitsPerPartition[p], self.C >= self.C_p[p])
            if self.maxCutsPerPartitions is not None:
                s.add(self.C_p[p] <= self.maxCutsPerPartitions)
        wire = [z3.If(c, 1, 0) for c, (_u, _v, t) in zip(c_e, self.edges) if t == "W"]
        gate = [z3.If(c, 1, 0) for c, (_u, _v, t) in zip(c_e, self.edges) if t == "G"]
        n_wire = z3.Sum(wire) if wire else z3.IntVal(0)
        n_gate = z3.Sum(gate) if gate else z3.IntVal(0)
        if self.forceNWireCuts is not None:
            s.add(n_wire == self.forceNWireCuts)
        if self.forceNGateCuts is not None:
            s.add(n_gate == self.forceNGateCuts)
        if self.maxNCuts is not None:
            s.add(n_wire + n_gate <= self.maxNCuts)
        if self.maxNQpdCuts is not None:
            qpd = [z3.If(z3.And(c, z3.Not(b)), 1, 0) for c, b in zip(c_e, b_e)]
            n_qpd = z3.Sum(qpd) if qpd else z3.IntVal(0)
            s.add([z3.Implies(b, n_qpd == self.maxNQpdCuts) for b in b_e])
            s.add(n_qpd <= self.maxNQpdCuts)
        # soft constraint (Cutter.py:545-551): every QPD cut lies before every teleport cut
        n_v = len(self.V)
        if self.edges:
            qpd_idx = [z3.If(z3.And(c, z3.Not(b)), v, -1) for c, b, (_u, v, _t) in zip(c_e, b_e, self.edges)]
            tele_idx = [z3.If(b, u, n_v) for b, (u, _v, _t) in zip(b_e, self.edges)]
            mx, mn = qpd_idx[0], tele_idx[0]
            for x in qpd_idx[1:]:
                mx = z3.If(x > mx, x, mx)
            for x in tele_idx[1:]:
                mn = z3.If(x < mn, x, mn)
            self.soft = mx < mn
        else:
            self.soft = z3.BoolVal(True)

    # ------------------------------------------------------------------ solving
    def _minimise(self, expr) -> None:
        """Fix ``expr`` to its minimum under the current constraints (which must be satisfiable)."""
        z3, s = self.z3, self.s
        best = s.model().eval(expr, model_completion=True).as_long()
        while True:
            s.push()
            s.add(expr < best)
            if s.check() == z3.sat:
                best = s.model().eval(expr, model_completion=True).as_long()
                s.pop()
            else:
                s.pop()
                break
        s.add(expr == best)
        assert s.check() == z3.sat

    def _minimise_S(self) -> None:
        """S = 6^(gate QPD cuts) * 8^(wire QPD cuts) (teleport cuts cost 1): the smallest feasible value is found by
        trying the (a, b) pairs in ascending order of 6^a 8^b - linear constraints only."""
        z3, s = self.z3, self.s
        n_g = sum(1 for e in self.edges if e[2] == "G")
        n_w = len(self.edges) - n_g
        bound = min(x for x in (self.maxNQpdCuts, self.maxNCuts, n_g + n_w) if x is not None)
        values = sorted({6 ** a * 8 ** b for a in range(min(n_g, bound) + 1) for b in range(min(n_w, bound) + 1)
                         if a + b <= bound})
        for v in values:
            pairs = [(a, b) for a in range(min(n_g, bound) + 1) for b in range(min(n_w, bound) + 1)
                     if a + b <= bound and 6 ** a * 8 ** b == v]
            s.push()
            s.add(z3.Or([z3.And(self._n_gate_qpd == a, self._n_wire_qpd == b) for a, b in pairs]))
            if s.check() == z3.sat:
                s.pop()
                s.add(z3.Or([z3.And(self._n_gate_qpd == a, self._n_wire_qpd == b) for a, b in pairs]))
                s.add(self.S == v)
                assert s.check() == z3.sat
                return
            s.pop()
        raise RuntimeError("no feasible sampling overhead")   # cannot happen: the constraints were satisfiable

    def solve(self) -> bool:
        z3, s = self.z3, self.s
        self.model = None
        self.nWireCuts = self.nGateCuts = 0
        if s.check() != z3.sat:
            return False
        s.push()                                          # the soft constraint comes first in the lexicographic order
        s.add(self.soft)
        if s.check() != z3.sat:
            s.pop()
            s.check()
        self._minimise(self.Q)
        self._minimise_S()
        s_val = s.model().eval(self.S, model_completion=True).as_long()
        s.add(self.A == self._total_a * s_val, self.L == self._total_l)     # linear now that S is a constant
        assert s.check() == z3.sat
        for obj in (self.A, self.L, self.C):
            self._minimise(obj)
        self.model = s.model()
        for c, (_u, _v, t) in zip(self.c_e, self.edges):
            if z3.is_true(self.model.eval(c, model_completion=True)):
                if t == "W":
                    self.nWireCuts += 1
                else:
                    self.nGateCuts += 1
        return True

    def getModelKeyResults(self):
        """-> S, A, L, nWireCuts, nGateCuts, Q, [Q_p], C, [C_p] (Cutter.py:162-178)."""
        if self.model is None:
            raise RuntimeError("no model exists")
        val = lambda x: self.model.eval(x, model_completion=True).as_long()
        return (val(self.S), val(self.A), val(self.L), self.nWireCuts, self.nGateCuts, val(self.Q),
                [val(q) for q in self.Q_p], val(self.C), [val(c) for c in self.C_p])

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
