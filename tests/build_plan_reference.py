"""The original per-pattern emit loop of ``FragmentProgram._build_plan`` (one Python call per op and pattern), kept as
the reference of the template-based form in compiler.py (which builds the rows once per program and selects /
renumbers them per pattern with numpy).  Returns what the loop produced before any scheduling."""
import numpy as np

from hardwareawareoptimalquantumcircuitcuttingandknitting_b200 import _lib


def emit_ops(prog, pattern: int, fold: bool):
    n = prog.n_qubits
    n_anc = 0
    ops = []
    cfg_pos = {}
    extra_out = []

    def emit(kind, q0, q1=0, mat=0, sel=-1, stride=0):
        ops.append([kind, q0, q1, mat, sel, stride, n + n_anc, 0])

    for top in prog.tops:
        if top[0] == "u1":
            emit(_lib.OP_U1, top[1], 0, top[2])
        elif top[0] == "cx":
            emit(_lib.OP_CX, top[1], top[2])
        elif top[0] == "cz":
            emit(_lib.OP_CZ, top[1], top[2])
        elif top[0] == "u2":
            emit(_lib.OP_U2, top[1], top[2], top[3])
        elif top[0] == "mmeas":
            anc = n + n_anc
            n_anc += 1
            emit(_lib.OP_CX, top[1], anc)
            extra_out.append((top[2], anc))
        else:
            s = top[1]
            slot = prog.slots[s]
            measures = bool((pattern >> s) & 1)
            if slot.pre_off >= 0:
                emit(_lib.OP_U1, slot.qubit, 0, slot.pre_off, slot.digit, 8)
            if measures:
                if slot.digit in cfg_pos:
                    raise NotImplementedError("both ends of a virtual gate measure inside one fragment")
                if slot.terminal:
                    cfg_pos[slot.digit] = slot.qubit
                else:
                    anc = n + n_anc
                    n_anc += 1
                    emit(_lib.OP_CX, slot.qubit, anc)
                    cfg_pos[slot.digit] = anc
            if slot.post_off >= 0:
                emit(_lib.OP_U1, slot.qubit, 0, slot.post_off, slot.digit, 8)
    n_state = n + n_anc
    bits = sorted(prog.out_bits + extra_out)
    out_pos = [p for _, p in bits]
    if not fold:
        out_pos += [cfg_pos.get(d, -1) for d in range(len(prog.radix))]
    used = {p for p in out_pos if p >= 0}
    sum_mask = sum(1 << b for b in range(n_state) if b not in used)
    sign_mask = sum(1 << p for p in cfg_pos.values()) if fold else 0
    return np.asarray(ops, dtype=np.int32).reshape(-1, 8), n_state, out_pos, sum_mask, sign_mask
