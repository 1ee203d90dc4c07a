"""Apply a *given* set of cuts to a circuit (host front end, inputs only).

The z3 model that *chooses* cuts (``src/HwAwareCutter/Cutter.py:38-571``) stays
on the host and is out of scope (BASELINE.json north_star).  What the hot path
consumes is the circuit ``Cutter.getResultCircs`` hands to qvm
(``Cutter.py:128-160``): registers ``frag0, frag1, ...``, cut two-qubit gates
replaced by ``VIRTUAL_GATE_TYPES[name](gate, label)`` (``Cutter.py:575-590``),
and each wire cut turned into a ``VirtualMove(SwapGate)`` onto a fresh ``vmove``
qubit to which every later op of that wire is re-targeted
(``Cutter.py:614-643``).  ``apply_cuts`` builds exactly that from an explicit
``CutSpec`` so the benchmark configs can be fed to the hot path without qiskit
or z3; ``BASELINE_CUTS`` records the cut shapes of the BASELINE.json configs
(found there with a sound z3 search).

Vgate index k = circuit order of the virtual gates, as in
``virtual_circuit.py:22-27``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .circuit import Barrier, CircuitInstruction, Gate, Measure, QuantumCircuit, QuantumRegister
from .virtual_gates import VIRTUAL_GATE_TYPES, VirtualMove

__all__ = ["CutSpec", "apply_cuts", "baseline_cut_spec", "BASELINE_CONFIGS"]


@dataclass
class CutSpec:
    """``gate_cuts``: indices into ``circuit.data`` of two-qubit gates to virtualise.
    ``wire_cuts``: ``(qubit index, data index)`` - cut that qubit's wire right after the
    instruction at ``data index`` (which must act on it).
    ``partitions``: optional grouping of the *original* qubit indices into fragments;
    ``None`` = connected components of what is left after cutting."""
    gate_cuts: list[int] = field(default_factory=list)
    wire_cuts: list[tuple[int, int]] = field(default_factory=list)
    partitions: list[list[int]] | None = None


def apply_cuts(circuit: QuantumCircuit, spec: CutSpec) -> QuantumCircuit:
    qubits = circuit.qubits
    nq = len(qubits)
    qpos = {q: i for i, q in enumerate(qubits)}
    gate_cuts = set(spec.gate_cuts)
    wire_after: dict[int, list[int]] = {}
    for q, idx in spec.wire_cuts:
        if qubits[q] not in circuit.data[idx].qubits:
            raise ValueError(f"wire cut ({q}, {idx}): instruction {idx} does not act on qubit {q}")
        wire_after.setdefault(idx, []).append(q)
    n_move = len(spec.wire_cuts)

    # pass 1: rewrite ops over integer wires 0..nq-1 (original) and nq.. (vmove)
    current = list(range(nq))                     # original qubit -> wire it currently lives on
    ops: list[tuple[object, tuple[int, ...], tuple]] = []
    cut_ctr = 0
    move_src: dict[int, int] = {}
    for idx, ins in enumerate(circuit.data):
        op = ins.operation
        wires = tuple(current[qpos[q]] for q in ins.qubits)
        if idx in gate_cuts:
            if not isinstance(op, Gate) or op.num_qubits != 2 or op.name not in VIRTUAL_GATE_TYPES:
                raise ValueError(f"gate cut at {idx}: cannot virtualise {op!r}")
            op = VIRTUAL_GATE_TYPES[op.name](op, f"{op.name} {getattr(op, 'label', None)}")
        ops.append((op, wires, ins.clbits))
        for q in wire_after.get(idx, ()):
            new_wire = nq + cut_ctr
            ops.append((VirtualMove(Gate("swap", 2, (), label=f"{idx}_{q}")), (current[q], new_wire), ()))
            move_src[new_wire] = current[q]
            current[q] = new_wire
            cut_ctr += 1
    n_wires = nq + n_move

    # pass 2: connected components over uncut multi-qubit ops
    parent = list(range(n_wires))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for op, wires, _ in ops:
        if isinstance(op, Barrier):               # includes virtual gates and wire-cut markers
            continue
        for w in wires[1:]:
            parent[find(w)] = find(wires[0])
    comps: dict[int, list[int]] = {}
    for w in range(n_wires):
        comps.setdefault(find(w), []).append(w)

    if spec.partitions is None:
        groups = sorted(comps.values(), key=min)
    else:
        owner = {}
        for f, part in enumerate(spec.partitions):
            for q in part:
                if q in owner:
                    raise ValueError(f"qubit {q} listed in two partitions")
                owner[q] = f
        if set(owner) != set(range(nq)):
            raise ValueError("partitions must list every original qubit exactly once")
        groups = [[] for _ in spec.partitions]
        for comp in comps.values():
            fs = {owner[w] for w in comp if w < nq}
            if len(fs) > 1:
                raise ValueError("Circuit contains gates that act on multiple fragments.")
            if not fs:      # component made only of vmove wires: follow the wire it was cut from
                src = comp[0]
                while src >= nq:
                    src = move_src[src]
                fs = {owner[src]}
            groups[fs.pop()].extend(comp)
        groups = [sorted(g) for g in groups]

    regs = [QuantumRegister(len(g), f"frag{i}") for i, g in enumerate(groups)]
    wire_to_qubit = {}
    for reg, g in zip(regs, groups):
        for j, w in enumerate(g):
            wire_to_qubit[w] = reg[j]
    out = QuantumCircuit(*regs, *circuit.cregs, name=f"{circuit.name}_cut")
    for op, wires, clbits in ops:
        out.data.append(CircuitInstruction(op, tuple(wire_to_qubit[w] for w in wires), clbits))
    return out


# --------------------------------------------------------------------------- BASELINE.json configs
BASELINE_CONFIGS = {
    # name: (generator name, qubits, depth, benchmark.py -p, -q)
    "bv16": ("bv", 16, 1, 2, 10),
    "hwe16d5": ("hwe", 16, 5, 2, 10),
    "syc16d5": ("syc", 16, 5, 2, 10),
    "syc32d1": ("syc", 32, 1, 2, 50),
    "qft16": ("qft", 16, 1, 2, 50),        # -q 10 is UNSAT in the reference: uncut
    "aqft16": ("aqft", 16, 1, 2, 50),      # likewise
    "add6": ("add", 6, 1, 2, 50),
}


def _two_qubit_indices(circuit: QuantumCircuit, a: int, b: int) -> list[int]:
    qa, qb = circuit.qubits[a], circuit.qubits[b]
    return [i for i, ins in enumerate(circuit.data)
            if isinstance(ins.operation, Gate) and ins.operation.num_qubits == 2
            and set(ins.qubits) == {qa, qb}]


def baseline_cut_spec(config: str, circuit: QuantumCircuit) -> CutSpec:
    """Cut shapes of SURVEY.md Appendix B for the *decomposed* circuit
    (``circuit.decompose_two_qubit()``: only ``cx`` left)."""
    if config == "bv16":
        # one wire cut on q15 between the 8th and the 9th cx -> 9 | 8 qubits, L = 8
        return CutSpec(wire_cuts=[(15, _two_qubit_indices(circuit, 7, 15)[0])])
    if config == "hwe16d5":
        # cx(7, 8) of every layer -> 8 | 8 qubits, L = 6^5
        return CutSpec(gate_cuts=_two_qubit_indices(circuit, 7, 8))
    if config == "syc16d5":
        # (5,6),(13,14) in layer A and (1,2),(9,10) in layer B -> columns {0,1} | {2,3}, L = 6^4
        cuts = []
        for a, b in ((5, 6), (13, 14), (1, 2), (9, 10)):
            cuts += _two_qubit_indices(circuit, a, b)
        return CutSpec(gate_cuts=sorted(cuts))
    if config == "syc32d1":
        # already a tensor product: no cut; rows {0,1} + the four gate-less qubits | rows {2,3}
        p0 = [q for q in range(16)] + [24, 31]
        p1 = [q for q in range(16, 32) if q not in (24, 31)]
        return CutSpec(partitions=[p0, p1])
    if config in ("qft16", "aqft16", "add6"):
        return CutSpec()
    raise KeyError(config)


def make_baseline(config: str, seed: int = 0, cut: str = "table") -> tuple[QuantumCircuit, QuantumCircuit]:
    """-> (decomposed input circuit, cut circuit) of a BASELINE.json config, i.e. the two
    arguments of ``compareOriginalCircWithCutCirc`` (``benchmarks/benchmark.py:99``).
    ``cut="table"`` applies the recorded cut shape (``baseline_cut_spec``); ``cut="solver"`` runs the
    cutter (``cutter.Cutter``, needs z3) with the limits of ``benchmarks/benchmark.py:41``."""
    from .generators import gen_circ
    if config.endswith(":solver"):                     # "aqft16:solver" == make_baseline("aqft16", cut="solver")
        config, cut = config[:-len(":solver")], "solver"
    name, n, depth, p, q = BASELINE_CONFIGS[config]
    circ = gen_circ(name, n, depth, seed=seed).decompose_two_qubit()
    if cut == "solver":
        from .cutter import Cutter
        cutter = Cutter(circ, p, q, maxNQpdCuts=5, maxNCuts=5, maxCutsPerPartitions=5)
        if not cutter.solve():
            raise ValueError(f"{config}: no cut satisfies -p {p} -q {q}")
        return cutter.decomposedCirc, cutter.getCutCirc()
    if cut != "table":
        raise ValueError(f"unknown cut source {cut!r}")
    return circ, apply_cuts(circ, baseline_cut_spec(config, circ))
