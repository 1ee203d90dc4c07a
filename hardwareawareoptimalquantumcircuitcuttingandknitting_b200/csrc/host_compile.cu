// Host-side pieces of the program compiler (no CUDA calls): the parts of compiler.py whose Python run time
// dominated the cold end-to-end path of the 16-qubit configs (one plan per measurement pattern, 64 plans
// for hwe-16 d5).
#include "qck_common.cuh"

#include <vector>

// List scheduling of the ops of ONE on-chip sweep into register clusters of QCK_CLUSTER_QUBITS tile qubits
// (see qck.h: QCK_OP_CLUSTER).  Each round takes, in program order, every op that is not blocked by an
// earlier untaken op and whose qubits still fit the cluster; a cluster is emitted as a header followed by
// its members with their qubits renamed to ranks among the cluster's ascending positions.
//   ops      [n_ops][8] int32 records on tile-local qubits (qck_op layout)
//   out      room for 2 * n_ops records; *n_out receives the number written
//   n_live of an emitted record / header never shrinks in execution order (a reordered ancilla CX may already
//   have populated a higher bit when an "earlier" op finally runs); 0 = the whole tile.
extern "C" int qck_host_cluster_ops(const int32_t* ops, int n_ops, int n_tile, int max_cluster_ops, int32_t* out,
                                    int* n_out) {
    if (!ops || !out || !n_out || n_ops < 0 || n_tile < 1 || n_tile > 32 || max_cluster_ops < 1) return QCK_ERR_INVALID_ARG;
    const int R = QCK_CLUSTER_QUBITS;
    int w = 0;
    auto emit = [&](const int32_t* r) { for (int j = 0; j < 8; ++j) out[8 * w + j] = r[j]; return out + 8 * w++; };
    if (n_tile < R) {
        for (int i = 0; i < n_ops; ++i) emit(ops + 8 * i);
        *n_out = w;
        return QCK_OK;
    }
    std::vector<uint32_t> mask(n_ops);
    for (int i = 0; i < n_ops; ++i) {
        const int32_t* r = ops + 8 * i;
        for (int k = 1; k <= (r[0] == QCK_OP_U1 ? 1 : 2); ++k)
            if (r[k] < 0 || r[k] >= n_tile) return QCK_ERR_INVALID_ARG;
        mask[i] = r[0] == QCK_OP_U1 ? (1u << r[1]) : ((1u << r[1]) | (1u << r[2]));
    }
    std::vector<int> remaining(n_ops), taken, rest;
    for (int i = 0; i < n_ops; ++i) remaining[i] = i;
    int nl_run = 0;
    while (!remaining.empty()) {
        uint32_t cset = 0, blocked = 0;
        taken.clear();
        rest.clear();
        for (int i : remaining) {
            const uint32_t m = mask[i];
            if (!(m & blocked) && __builtin_popcount(cset | m) <= R && (int)taken.size() < max_cluster_ops) {
                cset |= m;
                taken.push_back(i);
            } else {
                blocked |= m;
                rest.push_back(i);
            }
        }
        int nl = 0;
        for (int i : taken) nl = ops[8 * i + 6] > nl ? ops[8 * i + 6] : nl;
        nl = (nl <= 0 || nl_run < 0) ? 0 : (nl > nl_run ? nl : nl_run);
        nl_run = nl > 0 ? nl : -1;
        const int live = (nl > 0 && nl <= n_tile) ? nl : n_tile;
        const int k0 = ops[8 * taken[0]];
        if (live < R || (taken.size() == 1 && (k0 == QCK_OP_U2 || k0 == QCK_OP_CX || k0 == QCK_OP_CZ))) {
            // fewer live bits than register qubits, or a lone two-qubit op (the plain pass beats a one-member cluster)
            for (int i : taken) emit(ops + 8 * i)[6] = nl;
        } else {
            for (int p = 0; __builtin_popcount(cset) < R; ++p)  // pad with the lowest free positions
                if (!((cset >> p) & 1u)) cset |= 1u << p;
            int pos[3], rank[32], np = 0;
            for (int q = 0; q < 32; ++q)
                if ((cset >> q) & 1u) {
                    rank[q] = np;
                    if (np < 3) pos[np] = q;
                    ++np;
                }
            const int32_t header[8] = {QCK_OP_CLUSTER, (int32_t)taken.size(), R, pos[0], pos[1], pos[2], nl, 0};
            emit(header);
            for (int i : taken) {
                int32_t* r = emit(ops + 8 * i);
                r[1] = rank[r[1]];
                if (r[0] != QCK_OP_U1) r[2] = rank[r[2]];
                r[7] = 1;  // cluster member: qubits are ranks, not tile positions
            }
        }
        remaining.swap(rest);
    }
    *n_out = w;
    return QCK_OK;
}
