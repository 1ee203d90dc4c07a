"""Python original of the cluster list scheduler (was compiler._cluster_sweeps): the reference the C port
``qck_host_cluster_ops`` (csrc/host_compile.cu) is tested against.  Test infrastructure only."""
import numpy as np

from hardwareawareoptimalquantumcircuitcuttingandknitting_b200 import _lib

MAX_CLUSTER_OPS = 32


def cluster_sweeps_reference(ops: np.ndarray, sweeps: list):
    """Group the ops of every sweep into clusters acting on <= 3 tile qubits (list scheduling as
    in _schedule_sweeps, qubit sets as bit masks).  A cluster is emitted as a QCK_OP_CLUSTER
    header followed by its member ops, whose qubits are re-expressed as ranks among the cluster's
    three ascending positions."""
    R = _lib.CLUSTER_QUBITS
    rows = ops.tolist()
    out, new_sweeps = [], []
    for positions, b, e in sweeps:
        T = len(positions)
        begin = len(out)
        seg = rows[b:e]
        if T < R:
            out.extend(seg)
            new_sweeps.append((positions, begin, len(out)))
            continue
        masks = [(1 << r[1]) if r[0] == _lib.OP_U1 else ((1 << r[1]) | (1 << r[2])) for r in seg]
        remaining = list(range(len(seg)))
        nl_run = 0      # n_live must never shrink in EXECUTION order: a reordered ancilla CX may
        #                 already have populated a higher bit when an "earlier" op finally runs
        while remaining:
            cset = blocked = 0
            taken, rest = [], []
            for i in remaining:
                m = masks[i]
                if not (m & blocked) and bin(cset | m).count("1") <= R and len(taken) < MAX_CLUSTER_OPS:
                    cset |= m
                    taken.append(i)
                else:
                    blocked |= m
                    rest.append(i)
            nl = max(seg[i][6] for i in taken)
            nl = 0 if (nl <= 0 or nl_run < 0) else max(nl, nl_run)
            nl_run = nl if nl > 0 else -1      # 0 / -1: whole tile from here on
            live = nl if 0 < nl <= T else T
            if live < R:                       # fewer live bits than register qubits: plain ops
                for i in taken:
                    r = list(seg[i])
                    r[6] = nl
                    out.append(r)
                remaining = rest
                continue
            if len(taken) == 1 and seg[taken[0]][0] in (_lib.OP_U2, _lib.OP_CX, _lib.OP_CZ):
                # a lone two-qubit op: the plain pass (matrix in registers, every thread a few quads)
                # beats a one-member cluster
                r = list(seg[taken[0]])
                r[6] = nl
                out.append(r)
                remaining = rest
                continue
            p = 0
            while bin(cset).count("1") < R:    # pad with the lowest free live positions
                if not (cset >> p) & 1:
                    cset |= 1 << p
                p += 1
            pos = [q for q in range(T) if (cset >> q) & 1]
            rank = {q: j for j, q in enumerate(pos)}
            out.append([_lib.OP_CLUSTER, len(taken), R, pos[0], pos[1], pos[2], nl, 0])
            for i in taken:
                r = list(seg[i])
                r[1] = rank[r[1]]
                if r[0] != _lib.OP_U1:
                    r[2] = rank[r[2]]
                r[7] = 1                       # cluster member: qubits are ranks, not tile positions
                out.append(r)
            remaining = rest
        new_sweeps.append((positions, begin, len(out)))
    return np.asarray(out, dtype=np.int32).reshape(-1, 8), new_sweeps


