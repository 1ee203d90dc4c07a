"""Oracle-side fragment tables (TEST INFRASTRUCTURE): what the knit consumes, computed from the oracle's
own simulators only.  Restates the data flow of ``third_party/qvm/qvm/run.py:36-58``
(instantiate -> run -> ``from_counts``) for fragments without virtual gates, with exact
probabilities in place of sampled counts."""
import numpy as np

from . import cport


def fragment_out_mask(ov, frag) -> int:
    """clbits the fragment's own (final) measurements write (bit i = clbit i)."""
    cidx = {c: i for i, c in enumerate(ov.circuit.clbits)}
    mask = 0
    for op in ov.frag_ops[frag]:
        if op.operation is not None and getattr(op.operation, "name", "") == "measure":
            mask |= 1 << cidx[op.clbits[0]]
    return mask


def fragment_table_k0(ov, frag):
    """Exact table of a fragment without virtual gates: C oracle statevector, |amp|^2 compacted
    to the measured clbits in ascending order.  -> (table, clbit mask)"""
    inst = ov.instance(frag, ())
    qidx = {q: i for i, q in enumerate(inst.qubits)}
    cidx = {c: i for i, c in enumerate(inst.clbits)}
    last = {}
    for i, d in enumerate(inst.data):
        for q in d.qubits:
            last[q] = i
    pairs = []
    for i, d in enumerate(inst.data):
        if getattr(d.operation, "name", "") == "measure":
            assert last[d.qubits[0]] == i, "C-oracle fast path needs terminal measurements"
            pairs.append((cidx[d.clbits[0]], qidx[d.qubits[0]]))
    pairs.sort()
    mask = 0
    for c, _ in pairs:
        mask |= 1 << c
    prob = cport.simulate_probabilities(inst)
    if [q for _, q in pairs] == list(range(len(inst.qubits))):
        return prob, mask
    idx = np.arange(1 << len(inst.qubits), dtype=np.uint64)
    comp = np.zeros_like(idx)
    for j, (_, q) in enumerate(pairs):
        comp |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(j)
    return np.bincount(comp.astype(np.int64), weights=prob, minlength=1 << len(pairs)), mask


def all_tables_k0(cut):
    """-> (tables, masks) of every measuring fragment of a cut circuit without virtual gates."""
    from . import instantiate as oi
    ov = oi.OracleVirtualCircuit(cut)
    assert not ov.vgates
    pairs = [fragment_table_k0(ov, f) for f in ov.fragments if ov.has_measurement(f, ())]
    return [p[0] for p in pairs], [p[1] for p in pairs]
