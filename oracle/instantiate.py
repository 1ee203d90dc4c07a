"""Fragment extraction, label enumeration and instantiation (TEST INFRASTRUCTURE).

Restates ``third_party/qvm/qvm/virtual_circuit.py`` for the oracle:

* ``:21-27``   circuit-order list of virtual gates defines the vgate index ``k``;
* ``:97-113``  every virtual gate becomes two one-qubit endpoints ``(k, side)``;
* ``:115-131`` a fragment keeps the ops whose qubits all lie in its register;
  every classical register is kept on every fragment; barriers spanning
  fragments are dropped; any other straddling op is an error;
* ``:39-48``   labels = ``itertools.product`` of ``range(n_k)`` (vgate touches the
  fragment) or ``(-1,)``; ``[()]`` without virtual gates;
* ``:183-213`` an instance = fragment circuit + classical register ``vgate_c[K]``
  appended after the original cregs, endpoint ``k`` replaced by its share of
  instantiation ``label[k]`` with the measurement re-targeted to ``vgate_c[k]``;
  without virtual gates the fragment circuit is used as is.

The cut circuit is read by duck typing (virtual gates are recognised by class
name and carry ``original_gate`` / ``params``); tables come from
``oracle.qpd_tables``, matrices from ``oracle.gates`` - none of the product's
compiler, tables or kernels is used.
"""
import itertools
from types import SimpleNamespace as NS

from . import qpd_tables

_KINDS = {"VirtualMove": "move", "VirtualCX": "cx", "VirtualCZ": "cz", "VirtualCY": "cy",
          "VirtualRZZ": "rzz", "VirtualCPhase": "cp"}


def vgate_kind(op):
    return _KINDS.get(type(op).__name__)


def vgate_theta(op):
    """The user-facing angle theta of an rzz/cp virtual gate.  For cp the product
    mirrors the reference and has already overwritten params[0] with -theta/2."""
    kind = vgate_kind(op)
    if kind == "rzz":
        return op.params[0]
    if kind == "cp":
        return -2.0 * op.params[0]
    return None


class OracleVirtualCircuit:
    def __init__(self, circuit):
        self.circuit = circuit
        self.vgates = []                     # [(kind, theta, (qubit0, qubit1))]
        self.n_clbits = len(circuit.clbits)
        ops = []                             # flattened with endpoints
        for ins in circuit.data:
            kind = vgate_kind(ins.operation)
            if kind is not None:
                k = len(self.vgates)
                self.vgates.append((kind, vgate_theta(ins.operation), tuple(ins.qubits)))
                for side in range(2):
                    ops.append(NS(endpoint=(k, side), qubits=(ins.qubits[side],), clbits=(), operation=None))
                continue
            ops.append(NS(endpoint=None, qubits=tuple(ins.qubits), clbits=tuple(ins.clbits),
                          operation=ins.operation))
        self.tables = [qpd_tables.table(kind, th) for kind, th, _ in self.vgates]
        self.radices = [len(t) for t in self.tables]
        self.fragments = list(circuit.qregs)
        self.frag_ops = {}
        for frag in self.fragments:
            fq = set(frag)
            keep = []
            for op in ops:
                qs = set(op.qubits)
                if qs <= fq:
                    keep.append(op)
                elif op.operation is not None and type(op.operation).__name__ == "Barrier":
                    continue
                elif qs & fq:
                    raise ValueError("Circuit contains gates that act on multiple fragments.")
            self.frag_ops[frag] = keep

    # virtual_circuit.py:39-48
    def instance_labels(self, frag):
        if not self.vgates:
            return [()]
        fq = set(frag)
        per = [tuple(range(self.radices[k])) if set(qs) & fq else (-1,)
               for k, (_, _, qs) in enumerate(self.vgates)]
        return list(itertools.product(*per))

    # virtual_circuit.py:133-137
    def global_labels(self):
        return list(itertools.product(*[range(r) for r in self.radices]))

    def touches(self, frag):
        fq = set(frag)
        return [bool(set(qs) & fq) for _, _, qs in self.vgates]

    # virtual_circuit.py:197-213
    def instance(self, frag, label):
        """A flat circuit namespace (qubits, clbits, data) for statevector.exact_distribution."""
        qubits = list(frag)
        K = len(label)
        clbits = list(self.circuit.clbits) + [("vgate_c", k) for k in range(K)]
        data = []
        for op in self.frag_ops[frag]:
            if op.endpoint is None:
                data.append(NS(operation=op.operation, qubits=op.qubits, clbits=op.clbits))
                continue
            k, side = op.endpoint
            for entry in self.tables[k][label[k]][side]:
                if entry == qpd_tables.M:
                    data.append(NS(operation=NS(name="measure", params=[]), qubits=op.qubits,
                                   clbits=(("vgate_c", k),)))
                else:
                    data.append(NS(operation=NS(name=entry[0], params=list(entry[1]), _matrix=None),
                                   qubits=op.qubits, clbits=()))
        return NS(qubits=qubits, clbits=clbits, data=data)

    def has_measurement(self, frag, label):
        return any(getattr(d.operation, "name", "") == "measure" for d in self.instance(frag, label).data)
