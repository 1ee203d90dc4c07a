// Host-side program compiler (no CUDA calls): a flattened fragment circuit -> the op list the device programs
// are made of.  The reference has no counterpart for this step: it materialises every instance as a Qiskit
// circuit (third_party/qvm/qvm/virtual_circuit.py:183-213, `.decompose()` included) and hands the list to Aer
// (run.py:42).  Here a fragment is compiled ONCE; this file is the part of compiler.py whose Python run time
// was the cold end-to-end cost of every configuration (lowering + products of consecutive one-qubit gates,
// pair fusion).  The Python originals live on under tests/ as the reference these functions are checked
// against (same op lists, matrices to the last few ulp).
#include "qck_common.cuh"

#include <algorithm>
#include <complex>
#include <vector>

namespace {

typedef std::complex<double> cplx;

enum { IN_G1 = 1, IN_G2 = 2, IN_CX = 3, IN_CZ = 4, IN_MEASURE = 5, IN_ENDPOINT = 6 };
enum { T_U1 = 0, T_CX = 1, T_CZ = 2, T_U2 = 3, T_SLOT = 4, T_MMEAS = 5 };

struct Top {
    int kind, a, b, c;
};

struct M2 {
    cplx m[4];
};
struct M4 {
    cplx m[16];
};

static inline cplx cmul(const cplx& x, const cplx& y) {  // plain product (operator* goes through __muldc3)
    return cplx(x.real() * y.real() - x.imag() * y.imag(), x.real() * y.imag() + x.imag() * y.real());
}
static inline M2 mul2(const M2& a, const M2& b) {  // a @ b
    M2 r;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) r.m[2 * i + j] = cmul(a.m[2 * i], b.m[j]) + cmul(a.m[2 * i + 1], b.m[2 + j]);
    return r;
}
static inline M4 mul4(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            cplx s = 0;
            for (int k = 0; k < 4; ++k) s += cmul(a.m[4 * i + k], b.m[4 * k + j]);
            r.m[4 * i + j] = s;
        }
    return r;
}
static inline bool is_identity(const M2& x) {  // exact comparison, as compiler.py's _is_identity
    return x.m[0] == cplx(1, 0) && x.m[3] == cplx(1, 0) && x.m[1] == cplx(0, 0) && x.m[2] == cplx(0, 0);
}

struct Slot {
    int digit, vgate_idx, side, qubit, terminal, n_var, meas_mask, pre_off, post_off;
    M2 pre[QCK_MAX_VARIANTS], post[QCK_MAX_VARIANTS];
};

struct SweepH {
    std::vector<int> pos;
    int begin, end;
};
struct PlanH {
    int64_t pattern = 0;
    std::vector<int32_t> labels;
    int n_state = 0;
    std::vector<int32_t> ops;  // 8 per record
    std::vector<SweepH> sweeps;
    std::vector<int> out_pos;
    uint64_t sum_mask = 0, sign_mask = 0;
    int shared = 0, warp_base = 0, op_base = 0;
};
struct TreeLevelH {
    int kind, qubit, digit, pre_off, post_off, col_bit, seg_begin, seg_end;
    std::vector<std::pair<int, int>> choices;  // (representative variant, outcome or -1)
    std::vector<int> canon, meas;
};
struct TreeH {
    int n_base = 0, n_out_bits = 0, seg0_begin = 0, seg0_end = 0;
    std::vector<int32_t> ops;
    std::vector<TreeLevelH> levels;
    std::vector<std::pair<int, int>> free_bits;  // (row bit, state position)
    uint64_t base_sum = 0;
    std::vector<int64_t> counts;
    qck_sim_tree_plan st;
};
struct ImageH {
    std::vector<uint8_t> blob;
    int64_t off_ops = 0, off_labels = 0, off_extra = 0, off_src = 0;
    int dedupe = 0;
    std::vector<int32_t> labels;                  // all plans' labels, concatenated
    std::vector<int64_t> rep_ranges;              // (offset, count) per plan into the representatives
    std::vector<qck_sweep> sweeps;                // all plans' sweeps
    std::vector<qck_sim_plan> structs;            // d_ops / d_mats left NULL
    std::vector<int64_t> label_ranges;            // (offset, count) per plan into `labels`
};

}  // namespace

struct qck_host_program {
    int n_qubits = 0, n_clbits = 0;
    int warp = 0, mid_measures = 0, measures_anything = 0;
    double work = 0.0;  // sum over the labels of 2^(state bits) x op records: dist.py's sharding threshold
    std::vector<double> pool;
    std::vector<Top> tops;
    std::vector<Slot> slots;
    std::vector<int> order;      // state position -> fragment qubit
    std::vector<int> out_bits;   // (clbit, position) pairs of the terminal measurements, ascending clbit
    std::vector<int> touched;    // virtual gates with an endpoint here, ascending
    std::vector<int> radix;
    std::vector<int> out_clbits;
    char err[256] = {0};
    // second stage: per-pattern plans (index: fold) or the tree program, and their host images
    struct Config {
        int onchip_max = 13, stream_tile = 12, cluster = 1, share_prefix = 2, tree = 1, dedupe = 2;
        uint64_t early_bits = 0;
    } cfg;
    std::vector<PlanH> plans[2];
    bool have_plans[2] = {false, false};
    int tree_state = 0;  // 0: not tried, 1: built, -1: not eligible
    TreeH tree;
    std::vector<int32_t> canon;  // canonical label of every label
    ImageH image[2];
    bool have_image[2] = {false, false};
    std::vector<uint8_t> tree_blob;
    int64_t tree_off_ops = 0;

    int add(const cplx* m, int n) {
        const int off = (int)pool.size();
        for (int i = 0; i < n; ++i) {
            pool.push_back(m[i].real());
            pool.push_back(m[i].imag());
        }
        return off;
    }
    M2 mat2(int off) const {
        M2 r;
        for (int i = 0; i < 4; ++i) r.m[i] = cplx(pool[off + 2 * i], pool[off + 2 * i + 1]);
        return r;
    }
    M4 mat4(int off) const {
        M4 r;
        for (int i = 0; i < 16; ++i) r.m[i] = cplx(pool[off + 2 * i], pool[off + 2 * i + 1]);
        return r;
    }
    void fuse_pairs();
};

// ---------------------------------------------------------------------------------------------- pair fusion
// Runs of gates that act inside one qubit pair become a single 4x4 unitary when a cost model (FP64 work per
// amplitude + per-op overhead) says the dense product is cheaper than the separate gates.  Slots and mid-circuit
// measurements are barriers on their qubit; only gates on disjoint qubits are reordered.
void qck_host_program::fuse_pairs() {
    const double OVERHEAD = 6.0;
    struct Group {
        int a, b;
        std::vector<Top> ops;
        bool open;
    };
    std::vector<Group> groups;                 // in creation order (= the order Python's dict yields them)
    std::vector<int> group_of(n_qubits + 64, -1);
    std::vector<std::vector<Top>> pending(n_qubits + 64);
    std::vector<long> stamp(n_qubits + 64, -1);  // insertion order of the pending lists
    long clock = 0;
    std::vector<Top> out;
    out.reserve(tops.size());

    auto op_cost = [&](const Top& t) -> double {
        if (t.kind == T_U1) {
            const M2 m = mat2(t.b);
            return (m.m[1] == cplx(0, 0) && m.m[2] == cplx(0, 0)) ? 2.0 : 8.0;
        }
        if (t.kind == T_CX || t.kind == T_CZ) return 1.0;
        return 16.0;
    };
    auto embed = [&](const Top& t, int a, int b) -> M4 {
        M4 r;
        for (auto& x : r.m) x = 0;
        if (t.kind == T_U1) {
            const M2 u = mat2(t.b);
            if (t.a == a) {  // kron(I, u): index bit 0
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 2; ++j) r.m[4 * i + j] = r.m[4 * (i + 2) + j + 2] = u.m[2 * i + j];
            } else {         // kron(u, I): index bit 1
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 2; ++j) r.m[4 * (2 * i) + 2 * j] = r.m[4 * (2 * i + 1) + 2 * j + 1] = u.m[2 * i + j];
            }
            return r;
        }
        if (t.kind == T_CX) {
            r.m[0] = r.m[4 * 1 + 3] = r.m[4 * 2 + 2] = r.m[4 * 3 + 1] = 1;
        } else if (t.kind == T_CZ) {
            r.m[0] = r.m[5] = r.m[10] = 1;
            r.m[15] = -1;
        } else {
            r = mat4(t.c);
        }
        if (t.a == a && t.b == b) return r;
        static const int sw[4] = {0, 2, 1, 3};   // SWAP @ m @ SWAP
        M4 s;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) s.m[4 * i + j] = r.m[4 * sw[i] + sw[j]];
        return s;
    };
    auto close = [&](int gi) {
        Group& g = groups[gi];
        if (group_of[g.a] == gi) group_of[g.a] = -1;
        if (group_of[g.b] == gi) group_of[g.b] = -1;
        g.open = false;
        double separate = 0;
        for (const Top& t : g.ops) separate += op_cost(t) + OVERHEAD;
        if (g.ops.size() > 1 && separate > 16.0 + OVERHEAD) {
            M4 m;
            for (int i = 0; i < 16; ++i) m.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
            for (const Top& t : g.ops) m = mul4(embed(t, g.a, g.b), m);
            out.push_back({T_U2, g.a, g.b, add(m.m, 16)});
        } else {
            out.insert(out.end(), g.ops.begin(), g.ops.end());
        }
    };
    auto flush_pending = [&](int q) {
        out.insert(out.end(), pending[q].begin(), pending[q].end());
        pending[q].clear();
        stamp[q] = -1;
    };
    for (const Top& t : tops) {
        if (t.kind == T_U1) {
            const int q = t.a;
            if (group_of[q] >= 0) {
                groups[group_of[q]].ops.push_back(t);
            } else {
                if (stamp[q] < 0) stamp[q] = clock++;
                pending[q].push_back(t);
            }
        } else if (t.kind == T_CX || t.kind == T_CZ || t.kind == T_U2) {
            const int a = t.a, b = t.b;
            if (group_of[a] >= 0 && group_of[a] == group_of[b]) {
                groups[group_of[a]].ops.push_back(t);
                continue;
            }
            if (group_of[a] >= 0) close(group_of[a]);
            if (group_of[b] >= 0) close(group_of[b]);
            Group g{a, b, {}, true};
            g.ops.insert(g.ops.end(), pending[a].begin(), pending[a].end());
            g.ops.insert(g.ops.end(), pending[b].begin(), pending[b].end());
            pending[a].clear();
            pending[b].clear();
            stamp[a] = stamp[b] = -1;
            g.ops.push_back(t);
            groups.push_back(std::move(g));
            group_of[a] = group_of[b] = (int)groups.size() - 1;
        } else {  // slot / mid-circuit measurement: barrier on its qubit
            const int q = t.kind == T_SLOT ? slots[t.a].qubit : t.a;
            if (group_of[q] >= 0) close(group_of[q]);
            flush_pending(q);
            out.push_back(t);
        }
    }
    for (int gi = 0; gi < (int)groups.size(); ++gi)
        if (groups[gi].open) close(gi);
    std::vector<std::pair<long, int>> left;
    for (int q = 0; q < (int)pending.size(); ++q)
        if (stamp[q] >= 0) left.push_back({stamp[q], q});
    std::sort(left.begin(), left.end());
    for (auto& p : left) flush_pending(p.second);
    tops.swap(out);
}

// ---------------------------------------------------------------------------------------------- lowering
extern "C" int qck_host_lower(const int32_t* instr, int n_instr, const int32_t* endpoints, int n_endpoints,
                              const double* in_pool, int64_t in_pool_len, int n_qubits, int n_clbits, int flags,
                              int warp_max_qubits, int warp_max_depth, qck_host_program** out_prog) {
    if (!out_prog || n_instr < 0 || n_qubits < 1 || n_qubits > 62 || (n_instr && !instr)) return QCK_ERR_INVALID_ARG;
    *out_prog = nullptr;
    qck_host_program* P = new qck_host_program();
    P->n_qubits = n_qubits;
    P->n_clbits = n_clbits;
    auto fail = [&](int code) {
        delete P;
        return code;
    };
    auto in2 = [&](int off) {
        M2 r;
        for (int i = 0; i < 4; ++i) r.m[i] = cplx(in_pool[off + 2 * i], in_pool[off + 2 * i + 1]);
        return r;
    };
    // validate
    for (int i = 0; i < n_instr; ++i) {
        const int32_t* r = instr + 6 * i;
        const int two = r[0] == IN_G2 || r[0] == IN_CX || r[0] == IN_CZ;
        if (r[0] < IN_G1 || r[0] > IN_ENDPOINT || r[1] < 0 || r[1] >= n_qubits) return fail(QCK_ERR_INVALID_ARG);
        if (two && (r[2] < 0 || r[2] >= n_qubits || r[2] == r[1])) return fail(QCK_ERR_INVALID_ARG);
        if ((r[0] == IN_G1 && (r[4] < 0 || r[4] + 8 > in_pool_len)) || (r[0] == IN_G2 && (r[4] < 0 || r[4] + 32 > in_pool_len)))
            return fail(QCK_ERR_INVALID_ARG);
        if (r[0] == IN_ENDPOINT) {
            if (r[5] < 0 || r[5] >= n_endpoints) return fail(QCK_ERR_INVALID_ARG);
            const int32_t* e = endpoints + 6 * r[5];
            if (e[2] < 1 || e[2] > QCK_MAX_VARIANTS || e[4] < 0 || e[5] < 0 || e[4] + 8 * e[2] > in_pool_len ||
                e[5] + 8 * e[2] > in_pool_len || e[0] < 0)
                return fail(QCK_ERR_INVALID_ARG);
        }
    }
    std::vector<int> last_use(n_qubits, -1);
    for (int i = 0; i < n_instr; ++i) {
        const int32_t* r = instr + 6 * i;
        last_use[r[1]] = i;
        if (r[0] == IN_G2 || r[0] == IN_CX || r[0] == IN_CZ) last_use[r[2]] = i;
    }
    // finally-measured clbits define the output row
    std::vector<std::pair<int, int>> terminal;  // (clbit, qubit)
    for (int i = 0; i < n_instr; ++i) {
        const int32_t* r = instr + 6 * i;
        if (r[0] == IN_MEASURE && last_use[r[1]] == i) terminal.push_back({r[3], r[1]});
    }
    std::stable_sort(terminal.begin(), terminal.end(), [](auto& x, auto& y) { return x.first < y.first; });
    for (size_t i = 1; i < terminal.size(); ++i)
        if (terminal[i].first == terminal[i - 1].first) return fail(-1);  // two terminal measurements, one clbit
    std::vector<int> pos(n_qubits, -1);
    for (auto& t : terminal) {
        pos[t.second] = (int)P->order.size();
        P->order.push_back(t.second);
    }
    for (int q = 0; q < n_qubits; ++q)
        if (pos[q] < 0) {
            pos[q] = (int)P->order.size();
            P->order.push_back(q);
        }
    for (int i = 0; i < n_instr; ++i)
        if (instr[6 * i] == IN_ENDPOINT) P->touched.push_back(endpoints[6 * instr[6 * i + 5]]);
    std::sort(P->touched.begin(), P->touched.end());
    P->touched.erase(std::unique(P->touched.begin(), P->touched.end()), P->touched.end());
    P->radix.assign(P->touched.size(), 0);
    for (auto& t : terminal) {
        P->out_bits.push_back(t.first);
        P->out_bits.push_back(pos[t.second]);
    }
    std::vector<M2> pending(n_qubits);
    std::vector<char> has(n_qubits, 0);
    auto flush = [&](int q) {
        if (has[q]) {
            has[q] = 0;
            if (!is_identity(pending[q])) P->tops.push_back({T_U1, q, P->add(pending[q].m, 4), 0});
        }
    };
    for (int i = 0; i < n_instr; ++i) {
        const int32_t* r = instr + 6 * i;
        const int q0 = pos[r[1]];
        if (r[0] == IN_ENDPOINT) {
            const int32_t* e = endpoints + 6 * r[5];
            Slot s;
            s.vgate_idx = e[0];
            s.side = e[1];
            s.n_var = e[2];
            s.meas_mask = e[3];
            s.digit = (int)(std::lower_bound(P->touched.begin(), P->touched.end(), e[0]) - P->touched.begin());
            s.qubit = q0;
            s.terminal = last_use[r[1]] == i;
            s.pre_off = s.post_off = -1;
            P->radix[s.digit] = s.n_var;
            bool pre_id = true, post_id = true;
            for (int v = 0; v < s.n_var; ++v) {
                s.pre[v] = in2(e[4] + 8 * v);
                s.post[v] = in2(e[5] + 8 * v);
                if (has[q0]) s.pre[v] = mul2(s.pre[v], pending[q0]);
                pre_id = pre_id && is_identity(s.pre[v]);
                post_id = post_id && is_identity(s.post[v]);
            }
            has[q0] = 0;
            if (!pre_id) {
                s.pre_off = (int)P->pool.size();
                for (int v = 0; v < s.n_var; ++v) P->add(s.pre[v].m, 4);
            }
            if (!s.terminal && !post_id) {
                s.post_off = (int)P->pool.size();
                for (int v = 0; v < s.n_var; ++v) P->add(s.post[v].m, 4);
            }
            P->tops.push_back({T_SLOT, (int)P->slots.size(), 0, 0});
            P->slots.push_back(s);
        } else if (r[0] == IN_MEASURE) {
            flush(q0);
            if (last_use[r[1]] != i) {  // mid-circuit measurement of the input circuit
                P->tops.push_back({T_MMEAS, q0, r[3], 0});
                P->mid_measures++;
            }
        } else if (r[0] == IN_G1) {
            const M2 m = in2(r[4]);
            pending[q0] = has[q0] ? mul2(m, pending[q0]) : m;
            has[q0] = 1;
        } else {
            const int q1 = pos[r[2]];
            flush(q0);
            flush(q1);
            if (r[0] == IN_CX)
                P->tops.push_back({T_CX, q0, q1, 0});
            else if (r[0] == IN_CZ)
                P->tops.push_back({T_CZ, q0, q1, 0});
            else {
                M4 m;
                for (int k = 0; k < 16; ++k) m.m[k] = cplx(in_pool[r[4] + 2 * k], in_pool[r[4] + 2 * k + 1]);
                P->tops.push_back({T_U2, q0, q1, P->add(m.m, 16)});
            }
        }
    }
    // one-qubit gates still pending act on wires that are never measured afterwards: dropped
    int branch_points = P->mid_measures;
    bool any_u2 = false, any_slot_meas = false;
    for (const Slot& s : P->slots) {
        if (!s.terminal && s.meas_mask) branch_points++;
        any_slot_meas = any_slot_meas || s.meas_mask;
    }
    for (const Top& t : P->tops) any_u2 = any_u2 || t.kind == T_U2;
    const bool warp_wanted = flags & 1, fuse = flags & 2;
    P->warp = warp_wanted && n_qubits <= warp_max_qubits && branch_points <= warp_max_depth && !any_u2 &&
              (long)P->tops.size() + 2 * (long)P->slots.size() < 60000;
    if (!P->warp && fuse) P->fuse_pairs();
    for (size_t i = 0; i < P->out_bits.size(); i += 2) P->out_clbits.push_back(P->out_bits[i]);
    for (const Top& t : P->tops)
        if (t.kind == T_MMEAS) P->out_clbits.push_back(t.b);
    std::sort(P->out_clbits.begin(), P->out_clbits.end());
    P->measures_anything = !terminal.empty() || P->mid_measures > 0 || any_slot_meas;
    {   // work = sum_l 2^(n + a(l)) (base + a(l)), a(l) = ancillas of label l's pattern: a product over the digits
        double base = 0.0;
        for (const Top& t : P->tops) {
            if (t.kind == T_SLOT) base += (P->slots[t.a].pre_off >= 0) + (P->slots[t.a].post_off >= 0);
            else base += 1.0;
        }
        const int n_dig = (int)P->radix.size();
        std::vector<double> S(n_dig, 0.0), T(n_dig, 0.0);
        for (int d = 0; d < n_dig; ++d)
            for (int v = 0; v < P->radix[d]; ++v) {
                int a = 0;
                for (const Slot& sl : P->slots)
                    if (sl.digit == d && !sl.terminal && ((sl.meas_mask >> v) & 1)) ++a;
                S[d] += ldexp(1.0, a);
                T[d] += a * ldexp(1.0, a);
            }
        double prod = 1.0, extra = 0.0;
        for (int d = 0; d < n_dig; ++d) prod *= S[d];
        for (int d = 0; d < n_dig; ++d)
            if (S[d] > 0) extra += T[d] * prod / S[d];
        P->work = ldexp(1.0, n_qubits + P->mid_measures) * ((base > 1.0 ? base : 1.0) * prod + extra);
    }
    *out_prog = P;
    return QCK_OK;
}

extern "C" void qck_host_program_free(qck_host_program* p) { delete p; }

// ============================================================================================== second stage
// Per-pattern programs (compiler.py: plans / _plan_template / _build_plan / _schedule_sweeps), the tree program
// (tree), identical instances (canonical_labels) and the host images the executor uploads.
namespace {

const int LOW_RUN = 5, MAX_TERMS = 30, MAX_CLUSTER_OPS = 32;
const double DIAG_TOL = 1e-13;
const int64_t TREE_MAX_WORK_BYTES = (int64_t)2 << 30;
const int SHARE_PREFIX_MIN_INSTANCES = 1024, SHARE_PREFIX_MIN_PLAN = 64, DEDUPE_MIN_SAVED = 2048;
const double SHARE_PREFIX_MIN_FRACTION = 0.5;

struct TRow {
    int32_t r[8];
    int opt, alloc, clbit;
};

static std::vector<TRow> plan_template(const qck_host_program& P) {
    std::vector<TRow> rows;
    auto emit = [&](int kind, int q0, int q1, int mat, int sel, int stride, int o, int a, int c) {
        rows.push_back({{kind, q0, q1, mat, sel, stride, 0, 0}, o, a, c});
    };
    for (const Top& t : P.tops) {
        if (t.kind == T_U1) emit(QCK_OP_U1, t.a, 0, t.b, -1, 0, -1, 0, -1);
        else if (t.kind == T_CX) emit(QCK_OP_CX, t.a, t.b, 0, -1, 0, -1, 0, -1);
        else if (t.kind == T_CZ) emit(QCK_OP_CZ, t.a, t.b, 0, -1, 0, -1, 0, -1);
        else if (t.kind == T_U2) emit(QCK_OP_U2, t.a, t.b, t.c, -1, 0, -1, 0, -1);
        else if (t.kind == T_MMEAS) emit(QCK_OP_CX, t.a, 0, 0, -1, 0, -1, 1, t.b);
        else {
            const Slot& sl = P.slots[t.a];
            if (sl.pre_off >= 0) emit(QCK_OP_U1, sl.qubit, 0, sl.pre_off, sl.digit, 8, -1, 0, -1);
            if (!sl.terminal) emit(QCK_OP_CX, sl.qubit, 0, 0, -1, 0, t.a, 1, -1);
            if (sl.post_off >= 0) emit(QCK_OP_U1, sl.qubit, 0, sl.post_off, sl.digit, 8, -1, 0, -1);
        }
    }
    return rows;
}

// -> (need mask, diag mask) of one op record: a qubit is diag when the op never mixes amplitudes that differ in it
static void op_roles(const qck_host_program& P, const int32_t* r, uint64_t* need, uint64_t* diag) {
    const int kind = r[0], q0 = r[1], q1 = r[2], mat = r[3], sel = r[4];
    if (kind == QCK_OP_U1) {
        if (sel < 0) {
            const M2 m = P.mat2(mat);
            if (std::abs(m.m[1]) < DIAG_TOL && std::abs(m.m[2]) < DIAG_TOL) {
                *need = 0, *diag = 1ull << q0;
                return;
            }
        }
        *need = 1ull << q0, *diag = 0;
    } else if (kind == QCK_OP_CX) {
        *need = 1ull << q1, *diag = 1ull << q0;
    } else if (kind == QCK_OP_CZ) {
        *need = 0, *diag = (1ull << q0) | (1ull << q1);
    } else {
        const M4 m = P.mat4(mat);
        bool d0 = true, d1 = true;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                if (std::abs(m.m[4 * i + j]) >= DIAG_TOL) {
                    if ((i & 1) != (j & 1)) d0 = false;
                    if ((i >> 1) != (j >> 1)) d1 = false;
                }
        *need = (d0 ? 0 : 1ull << q0) | (d1 ? 0 : 1ull << q1);
        *diag = (d0 ? 1ull << q0 : 0) | (d1 ? 1ull << q1 : 0);
    }
}

static int schedule_sweeps(qck_host_program& P, PlanH& plan, int tile) {
    const int n_state = plan.n_state;
    tile = std::min(tile, n_state);
    const int low = std::max(0, std::min(LOW_RUN, tile - 2));
    const int n = (int)plan.ops.size() / 8;
    std::vector<int32_t> rows = plan.ops;
    std::vector<uint64_t> need(n), diag(n);
    for (int i = 0; i < n; ++i) op_roles(P, &rows[8 * i], &need[i], &diag[i]);
    std::vector<int> remaining(n), taken, rest;
    for (int i = 0; i < n; ++i) remaining[i] = i;
    std::vector<int32_t> new_ops;
    std::vector<SweepH> sweeps;
    struct Item {
        bool chain;
        int32_t row[8];
        int q;
        std::vector<int32_t> terms;
    };
    while (!remaining.empty()) {
        uint64_t tile_set = (1ull << low) - 1;
        if (sweeps.empty() && __builtin_popcountll(tile_set | P.cfg.early_bits) <= tile - 2)
            tile_set |= P.cfg.early_bits & ((1ull << n_state) - 1);
        uint64_t blocked_full = 0, blocked_diag = 0;
        taken.clear();
        rest.clear();
        for (int i : remaining) {
            const bool conflict = (need[i] & (blocked_full | blocked_diag)) || (diag[i] & blocked_full);
            if (!conflict && __builtin_popcountll(tile_set | need[i]) <= tile) {
                tile_set |= need[i];
                taken.push_back(i);
            } else {
                blocked_full |= need[i];
                blocked_diag |= diag[i];
                rest.push_back(i);
            }
        }
        if (taken.empty()) return QCK_ERR_UNSUPPORTED;  // an op does not fit into a tile
        for (int b = 0; __builtin_popcountll(tile_set) < tile; ++b) tile_set |= 1ull << b;
        SweepH sw;
        int local[64];
        for (int q = 0; q < 64; ++q) local[q] = -1;
        for (int q = 0; q < n_state; ++q)
            if ((tile_set >> q) & 1) {
                local[q] = (int)sw.pos.size();
                sw.pos.push_back(q);
            }
        sw.begin = (int)new_ops.size() / 8;
        std::vector<Item> items;
        int chain_of[64];
        for (int q = 0; q < 64; ++q) chain_of[q] = -1;
        std::vector<int32_t> phase_terms;
        for (int i : taken) {
            int32_t r[8];
            memcpy(r, &rows[8 * i], sizeof(r));
            const int nq = r[0] == QCK_OP_U1 ? 1 : 2;
            const int qs[2] = {r[1], r[2]};
            int outside[2], n_out = 0;
            for (int k = 0; k < nq; ++k)
                if (local[qs[k]] < 0) outside[n_out++] = qs[k];
            r[6] = 0;  // n_live is meaningless inside a tile
            if (n_out == 0) {
                for (int k = 0; k < nq; ++k) chain_of[qs[k]] = -1;  // anything else on the qubit closes its chain
                r[1] = local[r[1]];
                if (r[0] != QCK_OP_U1) r[2] = local[r[2]];
                Item it;
                it.chain = false;
                memcpy(it.row, r, sizeof(r));
                items.push_back(it);
            } else if (n_out == nq) {  // nothing in the tile: a scalar per tile
                cplx sc[4];
                int len;
                if (r[0] == QCK_OP_U1) {
                    const M2 m = P.mat2(r[3]);
                    sc[0] = m.m[0], sc[1] = m.m[3], len = 2;
                } else if (r[0] == QCK_OP_CZ) {
                    sc[0] = sc[1] = sc[2] = 1, sc[3] = -1, len = 4;
                } else {
                    const M4 m = P.mat4(r[3]);
                    for (int k = 0; k < 4; ++k) sc[k] = m.m[5 * k];
                    len = 4;
                }
                const int off = P.add(sc, len);
                const int32_t term[8] = {QCK_OP_TERM, outside[0], n_out > 1 ? outside[1] : -1, off, -1, len, 0, 1};
                phase_terms.insert(phase_terms.end(), term, term + 8);
            } else {  // one diag qubit outside: conditional op on the other one
                const int ext = outside[0];
                const int a = qs[1] == ext ? qs[0] : qs[1];
                cplx mm[8];
                for (auto& x : mm) x = 0;
                if (r[0] == QCK_OP_CX) {
                    mm[0] = mm[3] = 1, mm[5] = mm[6] = 1;
                } else if (r[0] == QCK_OP_CZ) {
                    mm[0] = mm[3] = 1, mm[4] = 1, mm[7] = -1;
                } else {
                    const M4 m = P.mat4(r[3]);
                    for (int i2 = 0; i2 < 2; ++i2)
                        for (int j2 = 0; j2 < 2; ++j2) {
                            if (ext == r[2]) {  // blocks of index bit 1
                                mm[2 * i2 + j2] = m.m[4 * i2 + j2];
                                mm[4 + 2 * i2 + j2] = m.m[4 * (i2 + 2) + j2 + 2];
                            } else {            // blocks of index bit 0
                                mm[2 * i2 + j2] = m.m[4 * (2 * i2) + 2 * j2];
                                mm[4 + 2 * i2 + j2] = m.m[4 * (2 * i2 + 1) + 2 * j2 + 1];
                            }
                        }
                }
                const int off = P.add(mm, 8);
                int ci = chain_of[a];
                if (ci < 0 || (int)items[ci].terms.size() / 8 >= MAX_TERMS) {
                    Item it;
                    it.chain = true;
                    it.q = local[a];
                    items.push_back(it);
                    ci = chain_of[a] = (int)items.size() - 1;
                }
                const int32_t term[8] = {QCK_OP_TERM, ext, -1, off, -1, 8, 0, 1};
                items[ci].terms.insert(items[ci].terms.end(), term, term + 8);
            }
        }
        const int n_phase = (int)phase_terms.size() / 8;
        for (int c0 = 0; c0 < n_phase; c0 += MAX_TERMS) {
            const int cnt = std::min(MAX_TERMS, n_phase - c0);
            const int32_t hdr[8] = {QCK_OP_PHASE, 0, cnt, 0, -1, 0, 0, 1};
            new_ops.insert(new_ops.end(), hdr, hdr + 8);
            new_ops.insert(new_ops.end(), phase_terms.begin() + 8 * c0, phase_terms.begin() + 8 * (c0 + cnt));
        }
        for (const Item& it : items) {
            if (it.chain) {
                const int32_t hdr[8] = {QCK_OP_U1X, it.q, (int32_t)it.terms.size() / 8, 0, -1, 0, 0, 0};
                new_ops.insert(new_ops.end(), hdr, hdr + 8);
                new_ops.insert(new_ops.end(), it.terms.begin(), it.terms.end());
            } else {
                new_ops.insert(new_ops.end(), it.row, it.row + 8);
            }
        }
        sw.end = (int)new_ops.size() / 8;
        sweeps.push_back(sw);
        remaining.swap(rest);
    }
    plan.ops.swap(new_ops);
    plan.sweeps.swap(sweeps);
    return QCK_OK;
}

static int cluster_sweeps(PlanH& plan) {
    const int n = (int)plan.ops.size() / 8;
    std::vector<int32_t> out((size_t)(2 * n + 1) * 8);
    int w = 0;
    for (SweepH& sw : plan.sweeps) {
        int n_out = 0;
        if (sw.end == sw.begin) {
            sw.begin = sw.end = w;
            continue;
        }
        const int rc = qck_host_cluster_ops(plan.ops.data() + 8 * sw.begin, sw.end - sw.begin, (int)sw.pos.size(),
                                            MAX_CLUSTER_OPS, out.data() + 8 * w, &n_out);
        if (rc != QCK_OK) return rc;
        sw.begin = w;
        sw.end = w + n_out;
        w += n_out;
    }
    out.resize((size_t)w * 8);
    plan.ops.swap(out);
    return QCK_OK;
}

static int build_plan(qck_host_program& P, const std::vector<TRow>& tmpl, const std::vector<int>& slot_order,
                      PlanH& plan, bool fold) {
    const int n = P.n_qubits;
    const int64_t pattern = plan.pattern;
    std::vector<int> anc_of_slot(P.slots.size(), -1);
    std::vector<std::pair<int, int>> bits;  // (clbit, position)
    for (size_t i = 0; i < P.out_bits.size(); i += 2) bits.push_back({P.out_bits[i], P.out_bits[i + 1]});
    int cum = 0;
    for (const TRow& t : tmpl) {
        if (t.opt >= 0 && !((pattern >> t.opt) & 1)) continue;
        int32_t r[8];
        memcpy(r, t.r, sizeof(r));
        if (t.alloc) {
            ++cum;
            r[2] = n + cum - 1;  // the CX target: the ancilla the row allocates
            if (t.clbit >= 0) bits.push_back({t.clbit, r[2]});
            else anc_of_slot[t.opt] = r[2];
        }
        r[6] = n + cum;          // n_live = fragment qubits + ancillas so far (its own included)
        plan.ops.insert(plan.ops.end(), r, r + 8);
    }
    const int n_anc = cum, n_state = n + n_anc, n_dig = (int)P.radix.size();
    std::vector<int> cfg_pos(n_dig, -1);
    for (int s : slot_order) {
        if (!((pattern >> s) & 1)) continue;
        const Slot& sl = P.slots[s];
        if (cfg_pos[sl.digit] >= 0) return QCK_ERR_UNSUPPORTED;  // both ends of a virtual gate measure in one fragment
        cfg_pos[sl.digit] = sl.terminal ? sl.qubit : anc_of_slot[s];
    }
    if (n_state > 40) return QCK_ERR_UNSUPPORTED;
    plan.n_state = n_state;
    std::sort(bits.begin(), bits.end());
    for (size_t i = 1; i < bits.size(); ++i)
        if (bits[i].first == bits[i - 1].first) return -2;  // a clbit is written by more than one measurement
    for (auto& b : bits) plan.out_pos.push_back(b.second);
    if (!fold)
        for (int d = 0; d < n_dig; ++d) plan.out_pos.push_back(cfg_pos[d]);
    if ((int)plan.out_pos.size() > QCK_MAX_OUT_BITS) return QCK_ERR_UNSUPPORTED;
    uint64_t used = 0;
    for (int p : plan.out_pos)
        if (p >= 0) used |= 1ull << p;
    plan.sum_mask = ((n_state >= 64 ? ~0ull : (1ull << n_state) - 1)) & ~used;
    plan.sign_mask = 0;
    if (fold)
        for (int d = 0; d < n_dig; ++d)
            if (cfg_pos[d] >= 0) plan.sign_mask |= 1ull << cfg_pos[d];
    const int n_ops = (int)plan.ops.size() / 8;
    auto whole = [&](int b, int e) {
        SweepH sw;
        for (int q = 0; q < n_state; ++q) sw.pos.push_back(q);
        sw.begin = b, sw.end = e;
        return sw;
    };
    if (P.warp) {
        plan.sweeps.push_back(whole(0, n_ops));
        plan.warp_base = n;
        return QCK_OK;
    }
    if (n_state <= P.cfg.onchip_max) {
        int split = 0;
        for (int i = 0; i < n_ops; ++i)
            if (plan.ops[8 * i + 4] >= 0) {
                split = i;
                break;
            }
        bool any_dep = false;
        for (int i = 0; i < n_ops; ++i) any_dep = any_dep || plan.ops[8 * i + 4] >= 0;
        if (!any_dep) split = 0;
        int64_t num_labels = 1;
        for (int r : P.radix) num_labels *= r;
        const bool share = split > 0 && (P.cfg.share_prefix == 1 ||
                                         (P.cfg.share_prefix == 2 && num_labels >= SHARE_PREFIX_MIN_INSTANCES &&
                                          (int)plan.labels.size() >= SHARE_PREFIX_MIN_PLAN &&
                                          split >= SHARE_PREFIX_MIN_FRACTION * n_ops));
        if (share) {
            plan.shared = 1;
            plan.sweeps.push_back(whole(0, split));
            plan.sweeps.push_back(whole(split, n_ops));
        } else {
            plan.sweeps.push_back(whole(0, n_ops));
        }
    } else {
        const int rc = schedule_sweeps(P, plan, P.cfg.stream_tile);
        if (rc != QCK_OK) return rc;
    }
    if (P.cfg.cluster && n_state <= P.cfg.onchip_max) return cluster_sweeps(plan);
    return QCK_OK;
}

static bool same_variant(const Slot& s, int v, int w) {
    for (int k = 0; k < 4; ++k)
        if (s.pre[v].m[k] != s.pre[w].m[k] || s.post[v].m[k] != s.post[w].m[k]) return false;
    return ((s.meas_mask >> v) & 1) == ((s.meas_mask >> w) & 1);
}

static void canonical_labels(qck_host_program& P) {
    if (!P.canon.empty()) return;
    const int n_dig = (int)P.radix.size();
    int64_t num_labels = 1;
    for (int r : P.radix) num_labels *= r;
    std::vector<std::vector<int>> maps(n_dig);
    for (int d = 0; d < n_dig; ++d) {
        maps[d].resize(P.radix[d]);
        for (int v = 0; v < P.radix[d]; ++v) {
            int first = v;
            for (int w = 0; w < v; ++w) {
                bool same = true;
                for (const Slot& s : P.slots)
                    if (s.digit == d && !same_variant(s, v, w)) same = false;
                if (same) {
                    first = w;
                    break;
                }
            }
            maps[d][v] = first;
        }
    }
    P.canon.resize(num_labels);
    std::vector<int> dig(n_dig, 0);
    for (int64_t l = 0; l < num_labels; ++l) {
        int64_t c = 0;
        for (int d = 0; d < n_dig; ++d) c = c * P.radix[d] + maps[d][dig[d]];
        P.canon[l] = (int32_t)c;
        for (int d = n_dig - 1; d >= 0; --d) {
            if (++dig[d] < P.radix[d]) break;
            dig[d] = 0;
        }
    }
}

static int build_plans(qck_host_program& P, bool fold) {
    if (P.have_plans[fold]) return QCK_OK;
    const int n_dig = (int)P.radix.size();
    int64_t num_labels = 1;
    for (int r : P.radix) num_labels *= r;
    if (num_labels >= ((int64_t)1 << 31) || P.slots.size() > 62) return QCK_ERR_UNSUPPORTED;
    // measurement pattern of every label
    std::vector<int64_t> pat(num_labels, 0);
    {   // contribution of (digit, variant) to the pattern, then an odometer over the labels (no divisions)
        std::vector<int64_t> contrib((size_t)n_dig * QCK_MAX_VARIANTS, 0);
        for (size_t s = 0; s < P.slots.size(); ++s) {
            const Slot& sl = P.slots[s];
            for (int v = 0; v < sl.n_var; ++v)
                if ((sl.meas_mask >> v) & 1) contrib[(size_t)sl.digit * QCK_MAX_VARIANTS + v] |= (int64_t)1 << s;
        }
        std::vector<int> dig(n_dig, 0);
        for (int64_t l = 0; l < num_labels; ++l) {
            int64_t p = 0;
            for (int d = 0; d < n_dig; ++d) p |= contrib[(size_t)d * QCK_MAX_VARIANTS + dig[d]];
            pat[l] = p;
            for (int d = n_dig - 1; d >= 0; --d) {
                if (++dig[d] < P.radix[d]) break;
                dig[d] = 0;
            }
        }
    }
    std::vector<int64_t> uniq(pat);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    const std::vector<TRow> tmpl = plan_template(P);
    std::vector<int> slot_order;
    for (const Top& t : P.tops)
        if (t.kind == T_SLOT) slot_order.push_back(t.a);
    std::vector<PlanH> plans(uniq.size());
    for (size_t i = 0; i < uniq.size(); ++i) plans[i].pattern = uniq[i];
    for (int64_t l = 0; l < num_labels; ++l) {
        const size_t i = std::lower_bound(uniq.begin(), uniq.end(), pat[l]) - uniq.begin();
        plans[i].labels.push_back((int32_t)l);
    }
    int op_base = 0;
    for (PlanH& plan : plans) {
        const int rc = build_plan(P, tmpl, slot_order, plan, fold);
        if (rc != QCK_OK) return rc;
        plan.op_base = op_base;
        op_base += (int)plan.ops.size() / 8;
    }
    P.plans[fold].swap(plans);
    P.have_plans[fold] = true;
    return QCK_OK;
}

// The fragment as ONE tree program, when eligible: register regime, at least one branching op, every virtual gate
// with exactly one endpoint here.
static void build_tree(qck_host_program& P) {
    if (P.tree_state) return;
    P.tree_state = -1;
    if (!(P.warp && P.cfg.tree)) return;
    TreeH T;
    std::vector<std::pair<int, int>> bits;  // (clbit, position or -1 for a mid-circuit measurement)
    for (size_t i = 0; i < P.out_bits.size(); i += 2) bits.push_back({P.out_bits[i], P.out_bits[i + 1]});
    for (const Top& t : P.tops)
        if (t.kind == T_MMEAS) bits.push_back({t.b, -1});
    std::stable_sort(bits.begin(), bits.end(), [](auto& x, auto& y) { return x.first < y.first; });
    auto row_bit_of_clbit = [&](int c) {
        for (size_t j = 0; j < bits.size(); ++j)
            if (bits[j].first == c) return (int)j;
        return -1;
    };
    for (size_t j = 0; j < bits.size(); ++j)
        if (bits[j].second >= 0) T.free_bits.push_back({(int)j, bits[j].second});
    int seg_begin = 0;
    std::vector<char> seen(P.radix.size(), 0);
    auto n_rows = [&]() { return (int)T.ops.size() / 8; };
    auto close_segment = [&]() {
        if (!T.levels.empty()) T.levels.back().seg_begin = seg_begin, T.levels.back().seg_end = n_rows();
        else T.seg0_begin = seg_begin, T.seg0_end = n_rows();
        seg_begin = n_rows();
    };
    for (const Top& t : P.tops) {
        if (t.kind == T_U1 || t.kind == T_CX || t.kind == T_CZ) {
            const int32_t r[8] = {t.kind == T_U1 ? QCK_OP_U1 : t.kind == T_CX ? QCK_OP_CX : QCK_OP_CZ, t.a,
                                  t.kind == T_U1 ? 0 : t.b, t.kind == T_U1 ? t.b : 0, -1, 0, 0, 0};
            T.ops.insert(T.ops.end(), r, r + 8);
        } else if (t.kind == T_MMEAS) {
            close_segment();
            TreeLevelH L{QCK_TREE_MMEAS, t.a, -1, -1, -1, row_bit_of_clbit(t.b), 0, 0, {{0, 0}, {0, 1}}, {0}, {1}};
            T.levels.push_back(L);
        } else if (t.kind == T_SLOT) {
            const Slot& sl = P.slots[t.a];
            if (seen[sl.digit]) return;  // both ends of a virtual gate in one fragment
            seen[sl.digit] = 1;
            close_segment();
            TreeLevelH L{sl.terminal ? QCK_TREE_TERMINAL : QCK_TREE_SLOT, sl.qubit, sl.digit, sl.pre_off,
                         sl.terminal ? -1 : sl.post_off, -1, 0, 0, {}, {}, {}};
            for (int v = 0; v < sl.n_var; ++v) {
                int first = v;
                for (int w = 0; w < v; ++w)
                    if (same_variant(sl, v, w)) {
                        first = w;
                        break;
                    }
                L.canon.push_back(first);
                L.meas.push_back((sl.meas_mask >> v) & 1);
            }
            for (int v = 0; v < sl.n_var; ++v) {
                if (L.canon[v] != v) continue;
                if (L.meas[v] && !sl.terminal) {
                    L.choices.push_back({v, 0});
                    L.choices.push_back({v, 1});
                } else {
                    L.choices.push_back({v, -1});
                }
            }
            T.levels.push_back(L);
        } else {
            return;
        }
    }
    close_segment();
    if (T.levels.empty() || (int)T.levels.size() > QCK_TREE_MAX_LEVELS) return;
    for (auto& L : T.levels)
        if ((int)L.choices.size() > QCK_TREE_MAX_CHOICES) return;
    for (char c : seen)
        if (!c) return;
    T.counts.push_back(1);
    for (auto& L : T.levels) {
        if (T.counts.back() > ((int64_t)1 << 40)) return;
        T.counts.push_back(T.counts.back() * (int64_t)L.choices.size());
    }
    const int64_t state_bytes = (int64_t)16 << std::max(P.n_qubits, 5);
    int64_t inner = 0;
    for (size_t i = 1; i + 1 < T.counts.size(); ++i) inner = std::max(inner, T.counts[i]);
    if ((double)2 * inner * state_bytes + (double)T.counts.back() * (double)((int64_t)8 << T.free_bits.size()) >
            (double)TREE_MAX_WORK_BYTES || T.counts.back() >= ((int64_t)1 << 31))
        return;
    uint64_t used = 0;
    for (auto& f : T.free_bits) used |= 1ull << f.second;
    T.base_sum = ((1ull << P.n_qubits) - 1) & ~used;
    T.n_base = P.n_qubits;
    T.n_out_bits = (int)bits.size();
    if (T.free_bits.size() > 16 || P.radix.size() > QCK_MAX_DIGITS) return;
    // the C ABI's struct with the host fields filled
    qck_sim_tree_plan& st = T.st;
    memset(&st, 0, sizeof(st));
    st.n_base = T.n_base, st.n_levels = (int)T.levels.size(), st.n_digits = (int)P.radix.size(), st.n_out_bits = T.n_out_bits;
    st.seg0_begin = T.seg0_begin, st.seg0_end = T.seg0_end;
    st.n_free = (int)T.free_bits.size();
    for (size_t r = 0; r < T.free_bits.size(); ++r) st.free_bit[r] = (int8_t)T.free_bits[r].first, st.free_pos[r] = (int8_t)T.free_bits[r].second;
    st.base_sum = T.base_sum;
    for (size_t k = 0; k < P.radix.size(); ++k) st.radix[k] = P.radix[k];
    for (size_t i = 0; i < T.levels.size(); ++i) {
        const TreeLevelH& lv = T.levels[i];
        qck_tree_level& L = st.level[i];
        L.seg_begin = lv.seg_begin, L.seg_end = lv.seg_end;
        L.kind = lv.kind, L.qubit = lv.qubit, L.digit = lv.digit, L.pre_off = lv.pre_off, L.post_off = lv.post_off;
        L.n_choices = (int)lv.choices.size(), L.col_bit = lv.col_bit;
        for (size_t v = 0; v < lv.meas.size(); ++v)
            if (lv.meas[v]) L.meas_mask |= 1u << v;
        for (size_t v = 0; v < lv.canon.size(); ++v) L.canon |= (uint32_t)lv.canon[v] << (4 * v);
        for (int v = 0; v < 8; ++v) L.first_choice[v] = -1;
        for (size_t c = 0; c < lv.choices.size(); ++c) {
            L.choice_variant[c] = (uint8_t)lv.choices[c].first;
            L.choice_outcome[c] = (int8_t)lv.choices[c].second;
            if (L.first_choice[lv.choices[c].first] < 0) L.first_choice[lv.choices[c].first] = (int8_t)c;
        }
    }
    // blob [mats f64 | ops i32]
    const int64_t mats_bytes = (int64_t)std::max<size_t>(P.pool.size(), 8) * 8;
    const int64_t n_ops_bytes = std::max<int64_t>((int64_t)T.ops.size() * 4, 32);
    P.tree_off_ops = (mats_bytes + 255) & ~(int64_t)255;
    P.tree_blob.assign(P.tree_off_ops + n_ops_bytes, 0);
    if (!P.pool.empty()) memcpy(P.tree_blob.data(), P.pool.data(), P.pool.size() * 8);
    if (!T.ops.empty()) memcpy(P.tree_blob.data() + P.tree_off_ops, T.ops.data(), T.ops.size() * 4);
    P.tree = T;
    P.tree_state = 1;
}

static int build_image(qck_host_program& P, bool fold) {
    if (P.have_image[fold]) return QCK_OK;
    int rc = build_plans(P, fold);
    if (rc != QCK_OK) return rc;
    std::vector<PlanH>& plans = P.plans[fold];
    ImageH& I = P.image[fold];
    canonical_labels(P);
    const int64_t num_labels = (int64_t)P.canon.size();
    int64_t saved = 0;
    for (int64_t l = 0; l < num_labels; ++l) saved += P.canon[l] != l;
    I.dedupe = saved > 0 && (P.cfg.dedupe == 1 || (P.cfg.dedupe == 2 && saved >= DEDUPE_MIN_SAVED));
    std::vector<int32_t> ops, extra;
    for (PlanH& p : plans) {
        ops.insert(ops.end(), p.ops.begin(), p.ops.end());
        I.label_ranges.push_back((int64_t)I.labels.size());
        I.label_ranges.push_back((int64_t)p.labels.size());
        I.labels.insert(I.labels.end(), p.labels.begin(), p.labels.end());
    }
    if (ops.empty()) ops.assign(8, 0);
    int64_t n_reps = 0;
    if (I.dedupe) {
        for (PlanH& p : plans) {
            int64_t cnt = 0;
            for (int32_t l : p.labels)
                if (P.canon[l] == l) {
                    extra.push_back(l);
                    ++cnt;
                }
            I.rep_ranges.push_back(n_reps);
            I.rep_ranges.push_back(cnt);
            n_reps += cnt;
        }
        extra.insert(extra.end(), P.canon.begin(), P.canon.end());
    }
    if ((int)P.radix.size() > QCK_MAX_DIGITS) return QCK_ERR_UNSUPPORTED;
    const int64_t mats_bytes = (int64_t)std::max<size_t>(P.pool.size(), 8) * 8;
    I.off_ops = (mats_bytes + 255) & ~(int64_t)255;
    I.off_labels = (I.off_ops + (int64_t)ops.size() * 4 + 255) & ~(int64_t)255;
    I.off_extra = (I.off_labels + (int64_t)I.labels.size() * 4 + 255) & ~(int64_t)255;
    I.off_src = I.off_extra + 4 * n_reps;
    I.blob.assign(I.off_extra + (int64_t)extra.size() * 4, 0);
    if (!P.pool.empty()) memcpy(I.blob.data(), P.pool.data(), P.pool.size() * 8);
    memcpy(I.blob.data() + I.off_ops, ops.data(), ops.size() * 4);
    if (!I.labels.empty()) memcpy(I.blob.data() + I.off_labels, I.labels.data(), I.labels.size() * 4);
    if (!extra.empty()) memcpy(I.blob.data() + I.off_extra, extra.data(), extra.size() * 4);
    size_t n_sweeps = 0;
    for (PlanH& p : plans) n_sweeps += p.sweeps.size();
    I.sweeps.resize(n_sweeps);
    I.structs.resize(plans.size());
    size_t w = 0;
    for (size_t pi = 0; pi < plans.size(); ++pi) {
        PlanH& p = plans[pi];
        qck_sim_plan& st = I.structs[pi];
        memset(&st, 0, sizeof(st));
        st.n_state_qubits = p.n_state;
        st.n_sweeps = (int)p.sweeps.size();
        st.sweeps = I.sweeps.data() + w;
        for (size_t i = 0; i < p.sweeps.size(); ++i) {
            const SweepH& sw = p.sweeps[i];
            qck_sweep& o = I.sweeps[w++];
            memset(&o, 0, sizeof(o));
            const int n_pos = p.warp_base ? p.warp_base : (int)sw.pos.size();
            if (n_pos > QCK_MAX_TILE_QUBITS + 2) return QCK_ERR_UNSUPPORTED;
            o.n_tile = n_pos;
            o.op_begin = p.op_base + sw.begin;
            o.op_end = p.op_base + sw.end;
            int f = 0;
            for (int k = sw.begin; k < sw.end; ++k) {
                const int kind = p.ops[8 * k];
                if (kind == QCK_OP_U1X || kind == QCK_OP_PHASE) f |= 1;
                if (kind == QCK_OP_CLUSTER) f |= 2;
            }
            if (p.shared && i == 0) f |= QCK_SWEEP_SHARED;
            if (p.warp_base) f |= QCK_SWEEP_WARP | (p.warp_base << 8);
            o.flags = f;
            for (int j = 0; j < n_pos; ++j) o.pos[j] = sw.pos[j];
        }
        st.n_digits = (int)P.radix.size();
        for (size_t k = 0; k < P.radix.size(); ++k) st.radix[k] = P.radix[k];
        st.n_out_bits = (int)p.out_pos.size();
        for (size_t j = 0; j < p.out_pos.size(); ++j) st.out_pos[j] = p.out_pos[j];
        st.sum_mask = p.sum_mask;
        st.sign_mask = p.sign_mask;
    }
    P.have_image[fold] = true;
    return QCK_OK;
}

}  // namespace

// Planning knobs (compiler.py's module constants / constructor arguments); call before the first query.
extern "C" int qck_host_program_configure(qck_host_program* p, int onchip_max, int stream_tile, int cluster,
                                          int share_prefix, int tree, int dedupe, uint64_t early_bits) {
    if (!p || onchip_max < 1 || stream_tile < 2 || stream_tile > QCK_MAX_TILE_QUBITS) return QCK_ERR_INVALID_ARG;
    p->cfg.onchip_max = onchip_max, p->cfg.stream_tile = stream_tile, p->cfg.cluster = cluster;
    p->cfg.share_prefix = share_prefix, p->cfg.tree = tree, p->cfg.dedupe = dedupe, p->cfg.early_bits = early_bits;
    return QCK_OK;
}

// stage: 0 = tree program (returns 1 when the fragment is eligible, 0 when not), 1 = plans(fold), 2 = host image
// of plans(fold), 3 = canonical labels.  Errors: QCK_ERR_UNSUPPORTED, -2 (a clbit written twice).
extern "C" int qck_host_program_build(qck_host_program* p, int stage, int fold) {
    if (!p || stage < 0 || stage > 3) return QCK_ERR_INVALID_ARG;
    fold = fold ? 1 : 0;
    if (stage == 0) {
        build_tree(*p);
        return p->tree_state == 1 ? 1 : 0;
    }
    if (stage == 1) return build_plans(*p, fold);
    if (stage == 2) return build_image(*p, fold);
    canonical_labels(*p);
    return QCK_OK;
}

// What a caller reads back (sizes first with buf == NULL; the stage must have been built):
//   0 meta   int32 {n_qubits, warp, mid_measures, measures_anything, n_tops, n_slots, n_digits, pool doubles}
//   1 tops   int32 [n_tops][4]          2 slots int32 [n_slots][9]     3 slot pre  f64 [n_slots][8][8]
//   4 order  int32 [n_qubits]           5 out_bits int32 [n][2]        6 touched int32   7 radix int32
//   8 out_clbits int32                  9 pool f64 (grows while plans are built)        10 slot post f64
//  11 summary int32: item 0 + {n_touched, n_out_clbits}, then touched, radix, out_clbits      12 work estimate f64
//  50 tree meta int64 {n_base, n_out_bits, seg0_begin, seg0_end, base_sum, n_levels, n_ops, blob offset of the ops}
//  51 tree ops int32 [n][8]            52 tree levels int32 [n][58]   53 free bits int32 [n][2]
//  54 node counts int64                55 qck_sim_tree_plan bytes     56 tree blob bytes
//  60 canonical labels int32
//  100 + 20 * fold + k:  0 plan meta int64 [n_plans][12]   1 labels int32   2 ops int32   3 sweeps int32 [n][43]
//      4 out_pos int32   5 image meta int64 {off_ops, off_labels, off_extra, off_src, dedupe, n_plans, blob bytes}
//      6 image blob bytes   7 qck_sim_plan structs (sweeps -> memory owned by the program)   8 representative
//      ranges int64 [n_plans][2]   9 label ranges int64 [n_plans][2]
extern "C" int64_t qck_host_program_get(const qck_host_program* p, int what, void* buf, int64_t cap_bytes) {
    if (!p) return -1;
    std::vector<int32_t> iv;
    std::vector<int64_t> lv;
    std::vector<double> dv;
    const void* raw = nullptr;
    int64_t bytes = -1;
    int type = 0;  // 0 int32, 1 double, 2 int64, 3 raw
    if (what >= 100 && what < 140) {
        const int fold = (what - 100) / 20, k = (what - 100) % 20;
        if (k <= 4 && !p->have_plans[fold]) return -1;
        if (k >= 5 && !p->have_image[fold]) return -1;
        const std::vector<PlanH>& plans = p->plans[fold];
        const ImageH& I = p->image[fold];
        switch (k) {
            case 0:
                type = 2;
                for (const PlanH& q : plans)
                    lv.insert(lv.end(), {q.pattern, (int64_t)q.labels.size(), q.n_state, (int64_t)q.ops.size() / 8,
                                         (int64_t)q.sweeps.size(), (int64_t)q.out_pos.size(), q.shared, q.warp_base,
                                         q.op_base, (int64_t)q.sum_mask, (int64_t)q.sign_mask, 0});
                break;
            case 1:
                for (const PlanH& q : plans) iv.insert(iv.end(), q.labels.begin(), q.labels.end());
                break;
            case 2:
                for (const PlanH& q : plans) iv.insert(iv.end(), q.ops.begin(), q.ops.end());
                break;
            case 3:
                for (const PlanH& q : plans)
                    for (const SweepH& sw : q.sweeps) {
                        int32_t rec[43];
                        for (int& x : rec) x = -1;
                        rec[0] = (int)sw.pos.size(), rec[1] = sw.begin, rec[2] = sw.end;
                        for (size_t t = 0; t < sw.pos.size() && t < 40; ++t) rec[3 + t] = sw.pos[t];
                        iv.insert(iv.end(), rec, rec + 43);
                    }
                break;
            case 4:
                for (const PlanH& q : plans) iv.insert(iv.end(), q.out_pos.begin(), q.out_pos.end());
                break;
            case 5:
                type = 2;
                lv = {I.off_ops, I.off_labels, I.off_extra, I.off_src, I.dedupe, (int64_t)I.structs.size(), (int64_t)I.blob.size()};
                break;
            case 6: type = 3, raw = I.blob.data(), bytes = (int64_t)I.blob.size(); break;
            case 7: type = 3, raw = I.structs.data(), bytes = (int64_t)(I.structs.size() * sizeof(qck_sim_plan)); break;
            case 8: type = 2, lv = I.rep_ranges; break;
            case 9: type = 2, lv = I.label_ranges; break;
            default: return -1;
        }
    } else if (what >= 50 && what <= 56) {
        if (p->tree_state != 1) return -1;
        const TreeH& T = p->tree;
        switch (what) {
            case 50:
                type = 2;
                lv = {T.n_base, T.n_out_bits, T.seg0_begin, T.seg0_end, (int64_t)T.base_sum, (int64_t)T.levels.size(),
                      (int64_t)T.ops.size() / 8, p->tree_off_ops};
                break;
            case 51: iv = T.ops; break;
            case 52:
                for (const TreeLevelH& L : T.levels) {
                    int32_t rec[58];
                    for (int& x : rec) x = -1;
                    const int32_t head[10] = {L.kind, L.qubit, L.digit, L.pre_off, L.post_off, L.col_bit, L.seg_begin,
                                              L.seg_end, (int32_t)L.choices.size(), (int32_t)L.canon.size()};
                    memcpy(rec, head, sizeof(head));
                    for (size_t c = 0; c < L.choices.size(); ++c) rec[10 + 2 * c] = L.choices[c].first, rec[11 + 2 * c] = L.choices[c].second;
                    for (size_t v = 0; v < L.canon.size(); ++v) rec[42 + v] = L.canon[v], rec[50 + v] = L.meas[v];
                    iv.insert(iv.end(), rec, rec + 58);
                }
                break;
            case 53:
                for (auto& f : T.free_bits) iv.insert(iv.end(), {f.first, f.second});
                break;
            case 54: type = 2, lv = T.counts; break;
            case 55: type = 3, raw = &T.st, bytes = (int64_t)sizeof(T.st); break;
            case 56: type = 3, raw = p->tree_blob.data(), bytes = (int64_t)p->tree_blob.size(); break;
        }
    } else {
        switch (what) {
            case 0:
                iv = {p->n_qubits, p->warp, p->mid_measures, p->measures_anything, (int)p->tops.size(), (int)p->slots.size(),
                      (int)p->radix.size(), (int)p->pool.size()};
                break;
            case 1:
                for (const Top& t : p->tops) iv.insert(iv.end(), {t.kind, t.a, t.b, t.c});
                break;
            case 2:
                for (const Slot& s : p->slots)
                    iv.insert(iv.end(), {s.digit, s.vgate_idx, s.side, s.qubit, s.terminal, s.n_var, s.meas_mask, s.pre_off, s.post_off});
                break;
            case 3:
            case 10:
                type = 1;
                for (const Slot& s : p->slots)
                    for (int v = 0; v < QCK_MAX_VARIANTS; ++v)
                        for (int k = 0; k < 4; ++k) {
                            const cplx x = v < s.n_var ? (what == 3 ? s.pre[v].m[k] : s.post[v].m[k]) : cplx(0, 0);
                            dv.push_back(x.real());
                            dv.push_back(x.imag());
                        }
                break;
            case 4: iv.assign(p->order.begin(), p->order.end()); break;
            case 5: iv.assign(p->out_bits.begin(), p->out_bits.end()); break;
            case 6: iv.assign(p->touched.begin(), p->touched.end()); break;
            case 7: iv.assign(p->radix.begin(), p->radix.end()); break;
            case 8: iv.assign(p->out_clbits.begin(), p->out_clbits.end()); break;
            case 9: type = 3, raw = p->pool.data(), bytes = (int64_t)p->pool.size() * 8; break;
            case 11:  // everything the hot path reads, in one call: item 0, then touched, radix, out_clbits
                iv = {p->n_qubits, p->warp, p->mid_measures, p->measures_anything, (int)p->tops.size(), (int)p->slots.size(),
                      (int)p->radix.size(), (int)p->pool.size(), (int)p->touched.size(), (int)p->out_clbits.size()};
                iv.insert(iv.end(), p->touched.begin(), p->touched.end());
                iv.insert(iv.end(), p->radix.begin(), p->radix.end());
                iv.insert(iv.end(), p->out_clbits.begin(), p->out_clbits.end());
                break;
            case 12: type = 1, dv = {p->work}; break;
            case 60: type = 3, raw = p->canon.data(), bytes = (int64_t)p->canon.size() * 4; break;
            default: return -1;
        }
    }
    if (type == 0) raw = iv.data(), bytes = (int64_t)iv.size() * 4;
    if (type == 1) raw = dv.data(), bytes = (int64_t)dv.size() * 8;
    if (type == 2) raw = lv.data(), bytes = (int64_t)lv.size() * 8;
    if (buf) {
        if (cap_bytes < bytes) return -1;
        if (bytes) memcpy(buf, raw, bytes);
    }
    return bytes;
}
