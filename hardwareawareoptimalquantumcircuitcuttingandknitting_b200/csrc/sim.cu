// Batched statevector simulation of fragment instances (sm_100a).
//
// Replaces the reference's per-instance circuit construction and Aer run
// (third_party/qvm/qvm/virtual_circuit.py:183-213, run.py:36-58).  One CTA owns one
// tile of one instance: the tile (2^n_tile complex128 amplitudes) lives in shared
// memory while a run of gates that only touch tile-local qubits is applied.
//   * on-chip regime  (n_state <= 13): one sweep, the whole state is the tile, HBM is
//     touched only by the output row;
//   * streaming regime (n_state  > 13): the state lives in HBM/L2, every sweep is one
//     coalesced read + one coalesced write of the state (tiles always contain the low
//     qubits so that a tile is made of >= 256-byte contiguous runs).
// Measurements whose qubit lives on are CX gates onto ancilla bits (deferred
// measurement); the epilogue folds |amp|^2 into the instance's output row with the
// (-1)^bit weights of the knit rules.
#include "qck_common.cuh"

#include <stdlib.h>

struct PlanDev {
    int n_state;
    const qck_op* ops;
    const double* mats;
    int n_digits;
    int radix[QCK_MAX_DIGITS];
    int n_out_bits;
    int out_pos[QCK_MAX_OUT_BITS];
    int out_ident;  // out_pos[j] == j for j < out_ident: those row bits map straight to state bits
    int n_stage;    // capacity (records) of the shared-memory program stage of this launch
    unsigned long long sum_mask, sign_mask;
};

struct SweepDev {
    int n_tile;
    int op_begin, op_end;
    int n_low;  // pos[i] == i for i < n_low (contiguous low run)
    int init;   // 1: synthesise |0..0> instead of reading
    int pos[QCK_MAX_TILE_QUBITS + 2];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

__device__ __forceinline__ void decode_digits(const PlanDev& plan, int label, int* digits) {
    int rem = label;
    for (int k = plan.n_digits - 1; k >= 0; --k) {
        int r = plan.radix[k];
        digits[k] = rem % r;
        rem /= r;
    }
}

// Shared-memory layout: amplitude i lives at slot swz(i).  The XOR folds index bits 3-5 into
// bits 0-2 so that eight lanes that differ in *any* three of those six bits hit eight different
// 16-byte bank groups: cluster passes whose register qubits are the low qubits (lane stride
// 128 B) stay conflict-free.  The permutation stays inside 128-byte lines.
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u); }

// ---- register-tiled gate application: a[k], k in [0, 8), bit j of k = cluster qubit j
template <int J>
__device__ __forceinline__ void reg_u1(double2 (&a)[8], double2 m00, double2 m01, double2 m10, double2 m11) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (!(k & (1 << J))) {
            const double2 x = a[k], y = a[k | (1 << J)];
            a[k] = cfma(m01, y, cmul(m00, x));
            a[k | (1 << J)] = cfma(m11, y, cmul(m10, x));
        }
}
template <int J>
__device__ __forceinline__ void reg_diag(double2 (&a)[8], double2 d0, double2 d1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = cmul((k & (1 << J)) ? d1 : d0, a[k]);
}
template <int J0, int J1>
__device__ __forceinline__ void reg_cx(double2 (&a)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if ((k & (1 << J0)) && !(k & (1 << J1))) {
            const double2 t = a[k];
            a[k] = a[k | (1 << J1)];
            a[k | (1 << J1)] = t;
        }
}
template <int J0, int J1>
__device__ __forceinline__ void reg_cz(double2 (&a)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if ((k & (1 << J0)) && (k & (1 << J1))) a[k] = make_double2(-a[k].x, -a[k].y);
}
template <int J0, int J1>
__device__ __forceinline__ void reg_u2(double2 (&a)[8], const double2* m) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (!(k & ((1 << J0) | (1 << J1)))) {
            const double2 x[4] = {a[k], a[k | (1 << J0)], a[k | (1 << J1)], a[k | (1 << J0) | (1 << J1)]};
            double2 r[4];
#pragma unroll
            for (int row = 0; row < 4; ++row) {
                double2 acc = cmul(m[row * 4], x[0]);
#pragma unroll
                for (int c = 1; c < 4; ++c) acc = cfma(m[row * 4 + c], x[c], acc);
                r[row] = acc;
            }
            a[k] = r[0];
            a[k | (1 << J0)] = r[1];
            a[k | (1 << J1)] = r[2];
            a[k | (1 << J0) | (1 << J1)] = r[3];
        }
}

#define QCK_PAIR_DISPATCH(FN, ...)                                     \
    switch (q0 * 3 + q1) {                                             \
        case 1: FN<0, 1>(__VA_ARGS__); break;                          \
        case 2: FN<0, 2>(__VA_ARGS__); break;                          \
        case 3: FN<1, 0>(__VA_ARGS__); break;                          \
        case 5: FN<1, 2>(__VA_ARGS__); break;                          \
        case 6: FN<2, 0>(__VA_ARGS__); break;                          \
        default: FN<2, 1>(__VA_ARGS__); break;                         \
    }

// ---- program staging --------------------------------------------------------------------------
// Op records and the (label-digit-resolved) 2x2 matrices of a chunk of the program are staged in
// shared memory once per CTA, so that per-op fetches are broadcast LDS instead of two dependent
// global loads (op record -> matrix), which dominated the run time of small instances.
struct StagedOp {
    int4 w0, w1;     // the qck_op words; w0.w = resolved matrix offset
    double2 m[16];   // U1: m[0..3]; U2: the 4x4 matrix, row major
};
#define QCK_STAGE_OPS 96
#define QCK_MAX_CLUSTER_OPS 32

__device__ __forceinline__ void stage_ops(StagedOp* so, const qck_op* __restrict__ ops, int c0, int n,
                                          const double* __restrict__ mats, const int* digits) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int4 w0 = __ldg(reinterpret_cast<const int4*>(ops + c0 + i));
        const int4 w1 = __ldg(reinterpret_cast<const int4*>(ops + c0 + i) + 1);
        if (w0.x == QCK_OP_U1 || w0.x == QCK_OP_U2) {
            if (w1.x >= 0) w0.w += digits[w1.x] * w1.y;
            const double2* m = reinterpret_cast<const double2*>(mats + w0.w);
            const int n_m = w0.x == QCK_OP_U1 ? 4 : 16;
            for (int e = 0; e < n_m; ++e) so[i].m[e] = __ldg(m + e);
        }
        so[i].w0 = w0;
        so[i].w1 = w1;
    }
}

// One cluster: header so[h], members so[h+1 .. h+n].  Every thread owns groups of 8 amplitudes
// (the 3 cluster bits enumerated) and applies all member ops in registers.
__device__ void run_cluster(double2* s, int T, const StagedOp* so, int h, const double* __restrict__ mats) {
    const int4 h0 = so[h].w0, h1 = so[h].w1;
    const int n_ops = h0.y, p0 = h0.w, p1 = h1.x, p2 = h1.y;
    int nl = h1.z;
    if (nl <= 0 || nl > T) nl = T;
    const uint32_t n_groups = 1u << (nl - 3);
    const uint32_t b0 = 1u << p0, b1 = 1u << p1, b2 = 1u << p2;
    for (uint32_t g = threadIdx.x; g < n_groups; g += blockDim.x) {
        const uint32_t base = insert_zero(insert_zero(insert_zero(g, p0), p1), p2);
        double2 a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            a[k] = s[swz(base | ((k & 1) ? b0 : 0u) | ((k & 2) ? b1 : 0u) | ((k & 4) ? b2 : 0u))];
        for (int i = h + 1; i <= h + n_ops; ++i) {
            const int4 w0 = so[i].w0;
            const int kind = w0.x, q0 = w0.y, q1 = w0.z;
            if (kind == QCK_OP_U1) {
                const double2 m00 = so[i].m[0], m01 = so[i].m[1], m10 = so[i].m[2], m11 = so[i].m[3];
                if (m01.x == 0.0 && m01.y == 0.0 && m10.x == 0.0 && m10.y == 0.0) {
                    if (m00.x == 1.0 && m00.y == 0.0 && m11.x == 1.0 && m11.y == 0.0) continue;
                    if (q0 == 0) reg_diag<0>(a, m00, m11);
                    else if (q0 == 1) reg_diag<1>(a, m00, m11);
                    else reg_diag<2>(a, m00, m11);
                } else {
                    if (q0 == 0) reg_u1<0>(a, m00, m01, m10, m11);
                    else if (q0 == 1) reg_u1<1>(a, m00, m01, m10, m11);
                    else reg_u1<2>(a, m00, m01, m10, m11);
                }
            } else if (kind == QCK_OP_CX) {
                QCK_PAIR_DISPATCH(reg_cx, a)
            } else if (kind == QCK_OP_CZ) {
                QCK_PAIR_DISPATCH(reg_cz, a)
            } else {
                const double2* m = so[i].m;
                QCK_PAIR_DISPATCH(reg_u2, a, m)
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            s[swz(base | ((k & 1) ? b0 : 0u) | ((k & 2) ? b1 : 0u) | ((k & 4) ? b2 : 0u))] = a[k];
    }
    __syncthreads();
}

// One un-clustered op (states with fewer than 3 live bits, or clustering disabled).
__device__ void run_single(double2* s, int T, const StagedOp& op, const double* __restrict__ mats) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const int kind = op.w0.x, q0 = op.w0.y, q1 = op.w0.z;
    int nl = op.w1.z;
    if (nl <= 0 || nl > T) nl = T;
    if (kind == QCK_OP_U1) {
        const double2 m00 = op.m[0], m01 = op.m[1], m10 = op.m[2], m11 = op.m[3];
        const uint32_t n = 1u << (nl - 1);
        if (m01.x == 0.0 && m01.y == 0.0 && m10.x == 0.0 && m10.y == 0.0) {
            if (m00.x == 1.0 && m00.y == 0.0 && m11.x == 1.0 && m11.y == 0.0) return;  // uniform
            for (uint32_t p = tid; p < n; p += nth) {
                const uint32_t i0 = insert_zero(p, q0), i1 = i0 | (1u << q0);
                s[swz(i0)] = cmul(m00, s[swz(i0)]);
                s[swz(i1)] = cmul(m11, s[swz(i1)]);
            }
        } else {
            for (uint32_t p = tid; p < n; p += nth) {
                const uint32_t i0 = insert_zero(p, q0), i1 = i0 | (1u << q0);
                const double2 a0 = s[swz(i0)], a1 = s[swz(i1)];
                s[swz(i0)] = cfma(m01, a1, cmul(m00, a0));
                s[swz(i1)] = cfma(m11, a1, cmul(m10, a0));
            }
        }
    } else {
        const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
        const uint32_t n = 1u << (nl - 2);
        const uint32_t b0 = 1u << q0, b1 = 1u << q1;
        if (kind == QCK_OP_CX) {
            for (uint32_t p = tid; p < n; p += nth) {
                const uint32_t base = insert_zero(insert_zero(p, lo), hi) | b0;  // control set
                const double2 a = s[swz(base)], b = s[swz(base | b1)];
                s[swz(base)] = b;
                s[swz(base | b1)] = a;
            }
        } else if (kind == QCK_OP_CZ) {
            for (uint32_t p = tid; p < n; p += nth) {
                const uint32_t idx = insert_zero(insert_zero(p, lo), hi) | b0 | b1;
                const double2 a = s[swz(idx)];
                s[swz(idx)] = make_double2(-a.x, -a.y);
            }
        } else {  // QCK_OP_U2: generic 4x4, row/col index = bit(q0) + 2 bit(q1)
            // The matrix lives in registers for all of this thread's quads: 64 FP64 instructions per
            // 4 amplitudes plus 4 LDS + 4 STS.  Index math: insert_zero and swz are bitwise-linear, so
            // for p = tid + k * nth (nth a power of two) slot(p) = slot(tid) ^ slot(k * nth): the
            // per-thread part is computed once per op, the per-k part is warp-uniform.
            double2 m[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) m[e] = op.m[e];
            const uint32_t x1 = swz(b0), x2 = swz(b1), x3 = x1 ^ x2;
            if ((nth & (nth - 1)) == 0) {
                const uint32_t st = swz(insert_zero(insert_zero((uint32_t)tid, lo), hi));
                if ((uint32_t)tid < n) {
#pragma unroll 2
                    for (uint32_t pk = 0; pk < n; pk += nth) {
                        const uint32_t i0 = st ^ swz(insert_zero(insert_zero(pk, lo), hi));
                        const uint32_t i1 = i0 ^ x1, i2 = i0 ^ x2, i3 = i0 ^ x3;
                        const double2 a0 = s[i0], a1 = s[i1], a2 = s[i2], a3 = s[i3];
                        s[i0] = cfma(m[3], a3, cfma(m[2], a2, cfma(m[1], a1, cmul(m[0], a0))));
                        s[i1] = cfma(m[7], a3, cfma(m[6], a2, cfma(m[5], a1, cmul(m[4], a0))));
                        s[i2] = cfma(m[11], a3, cfma(m[10], a2, cfma(m[9], a1, cmul(m[8], a0))));
                        s[i3] = cfma(m[15], a3, cfma(m[14], a2, cfma(m[13], a1, cmul(m[12], a0))));
                    }
                }
            } else {
                for (uint32_t p = tid; p < n; p += nth) {
                    const uint32_t i0 = swz(insert_zero(insert_zero(p, lo), hi));
                    const uint32_t i1 = i0 ^ x1, i2 = i0 ^ x2, i3 = i0 ^ x3;
                    const double2 a0 = s[i0], a1 = s[i1], a2 = s[i2], a3 = s[i3];
                    s[i0] = cfma(m[3], a3, cfma(m[2], a2, cfma(m[1], a1, cmul(m[0], a0))));
                    s[i1] = cfma(m[7], a3, cfma(m[6], a2, cfma(m[5], a1, cmul(m[4], a0))));
                    s[i2] = cfma(m[11], a3, cfma(m[10], a2, cfma(m[9], a1, cmul(m[8], a0))));
                    s[i3] = cfma(m[15], a3, cfma(m[14], a2, cfma(m[13], a1, cmul(m[12], a0))));
                }
            }
        }
    }
    __syncthreads();
}

// Apply ops[begin, end) to the tile `s` (2^T amplitudes in shared memory).  All threads of the
// CTA call this with identical arguments; the state is synchronised on return.
__device__ void apply_ops(double2* s, int T, StagedOp* so, int n_stage, const qck_op* __restrict__ ops, int begin,
                          int end, const double* __restrict__ mats, const int* digits, bool prestaged) {
    int c0 = begin;
    while (c0 < end) {
        const int n = (end - c0) < n_stage ? (end - c0) : n_stage;
        if (!(prestaged && c0 == begin)) {  // the caller may have staged the first chunk already
            stage_ops(so, ops, c0, n, mats, digits);
            __syncthreads();
        }
        int i = 0;
        while (i < n) {
            if (so[i].w0.x == QCK_OP_CLUSTER) {
                const int members = so[i].w0.y;
                if (i + members >= n && c0 + n < end) break;  // cluster continues past the staged chunk
                run_cluster(s, T, so, i, mats);
                i += 1 + members;
            } else {
                run_single(s, T, so[i], mats);
                ++i;
            }
        }
        c0 += i;
        __syncthreads();  // everyone is done with so[] before it is restaged
    }
}

// |amp|^2 folded into the output row (deterministic order, no atomics): one thread per
// output entry walks the summed-out bits with the masked-increment trick.
template <typename LoadAmp>
__device__ __forceinline__ double fold_entry(const PlanDev& plan, uint64_t o, LoadAmp load) {
    uint64_t base = o & ((1ull << plan.out_ident) - 1ull);
    for (int j = plan.out_ident; j < plan.n_out_bits; ++j) {
        if ((o >> j) & 1ull) {
            if (plan.out_pos[j] < 0) return 0.0;  // that bit is never written in this pattern
            base |= 1ull << plan.out_pos[j];
        }
    }
    const uint64_t sm = plan.sum_mask, sg = plan.sign_mask;
    double acc = 0.0;
    uint64_t sub = 0;
    do {
        double2 a = load(base | sub);
        double p = fma(a.x, a.x, a.y * a.y);
        acc += (__popcll(sub & sg) & 1) ? -p : p;
        sub = (sub - sm) & sm;
    } while (sub != 0);
    return acc;
}

// ------------------------------------------------------------------ on-chip regime
extern __shared__ __align__(16) unsigned char smem_raw[];

__global__ void __launch_bounds__(256) sim_onchip_kernel(PlanDev plan, int op_begin, int op_end,
                                                         const int32_t* __restrict__ labels,
                                                         double* __restrict__ out, long long row_stride) {
    double2* s = reinterpret_cast<double2*>(smem_raw);
    StagedOp* so = reinterpret_cast<StagedOp*>(smem_raw + ((size_t)16 << plan.n_state));
    __shared__ int digits[QCK_MAX_DIGITS];
    const int label = labels[blockIdx.x];
    if (threadIdx.x == 0) decode_digits(plan, label, digits);
    const uint32_t n_amp = 1u << plan.n_state;
    __syncthreads();  // digits visible
    {
        const int n0 = (op_end - op_begin) < plan.n_stage ? (op_end - op_begin) : plan.n_stage;
        stage_ops(so, plan.ops, op_begin, n0, plan.mats, digits);  // global loads overlap the state init
    }
    for (uint32_t i = threadIdx.x; i < n_amp; i += blockDim.x) s[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);  // swz(0) == 0
    __syncthreads();
    apply_ops(s, plan.n_state, so, plan.n_stage, plan.ops, op_begin, op_end, plan.mats, digits, true);
    const uint64_t n_out = 1ull << plan.n_out_bits;
    double* row = out + (long long)label * row_stride;
    for (uint64_t o = threadIdx.x; o < n_out; o += blockDim.x)
        row[o] = fold_entry(plan, o, [&](uint64_t idx) { return s[swz((uint32_t)idx)]; });
}

// Grouped form: one launch covers several programs (measurement patterns) of the same state size.
// The per-program parts travel in the kernel parameters (CUDA >= 12.1 allows 32 KB of them), the CTA
// looks up its program from its block index.  Replaces one launch per pattern (64 per step at
// hwe-16-d5) by one per state size.
#define QCK_GROUP_MAX 24
struct PlanVarDev {
    int op_begin, op_end;
    int n_out_bits, out_ident;
    signed char out_pos[QCK_MAX_OUT_BITS];
    unsigned long long sum_mask, sign_mask;
    const int32_t* labels;
    int cta_begin, n_stage_unused;
};
struct GroupDev {
    int n_state, n_vars, n_stage, n_digits;
    const qck_op* ops;
    const double* mats;
    int radix[QCK_MAX_DIGITS];
    PlanVarDev var[QCK_GROUP_MAX];
};

__global__ void __launch_bounds__(256) sim_onchip_group_kernel(const __grid_constant__ GroupDev G,
                                                               double* __restrict__ out, long long row_stride) {
    double2* s = reinterpret_cast<double2*>(smem_raw);
    StagedOp* so = reinterpret_cast<StagedOp*>(smem_raw + ((size_t)16 << G.n_state));
    __shared__ int digits[QCK_MAX_DIGITS];
    __shared__ PlanDev plan;
    int v = 0;
    while (v + 1 < G.n_vars && (int)blockIdx.x >= G.var[v + 1].cta_begin) ++v;  // uniform
    const PlanVarDev& pv = G.var[v];
    const int label = __ldg(pv.labels + ((int)blockIdx.x - pv.cta_begin));
    if (threadIdx.x == 0) {
        plan.n_state = G.n_state;
        plan.ops = G.ops;
        plan.mats = G.mats;
        plan.n_digits = G.n_digits;
        for (int k = 0; k < QCK_MAX_DIGITS; ++k) plan.radix[k] = G.radix[k];
        plan.n_out_bits = pv.n_out_bits;
        plan.out_ident = pv.out_ident;
        for (int j = 0; j < QCK_MAX_OUT_BITS; ++j) plan.out_pos[j] = pv.out_pos[j];
        plan.sum_mask = pv.sum_mask;
        plan.sign_mask = pv.sign_mask;
        plan.n_stage = G.n_stage;
        decode_digits(plan, label, digits);
    }
    __syncthreads();
    {
        const int n0 = (pv.op_end - pv.op_begin) < G.n_stage ? (pv.op_end - pv.op_begin) : G.n_stage;
        stage_ops(so, G.ops, pv.op_begin, n0, G.mats, digits);
    }
    const uint32_t n_amp = 1u << G.n_state;
    for (uint32_t i = threadIdx.x; i < n_amp; i += blockDim.x) s[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);  // swz(0) == 0
    __syncthreads();
    apply_ops(s, G.n_state, so, G.n_stage, G.ops, pv.op_begin, pv.op_end, G.mats, digits, true);
    const uint64_t n_out = 1ull << plan.n_out_bits;
    double* row = out + (long long)label * row_stride;
    for (uint64_t o = threadIdx.x; o < n_out; o += blockDim.x)
        row[o] = fold_entry(plan, o, [&](uint64_t idx) { return s[swz((uint32_t)idx)]; });
}

// ------------------------------------------------------------------ streaming regime
__global__ void __launch_bounds__(256) sim_sweep_kernel(PlanDev plan, SweepDev sw,
                                                        const int32_t* __restrict__ labels, int inst_base,
                                                        double2* __restrict__ work,
                                                        unsigned long long state_stride) {
    const int T = sw.n_tile, c = sw.n_low, n_hi = T - c;
    double2* s = reinterpret_cast<double2*>(smem_raw);
    unsigned long long* hi_off = reinterpret_cast<unsigned long long*>(smem_raw + ((size_t)16 << T));
    StagedOp* so = reinterpret_cast<StagedOp*>(smem_raw + ((size_t)16 << T) + ((((size_t)8 << n_hi) + 15) & ~(size_t)15));
    __shared__ int digits[QCK_MAX_DIGITS];
    const int label = labels[inst_base + blockIdx.y];
    if (threadIdx.x == 0) decode_digits(plan, label, digits);
    // offsets of the non-contiguous tile bits
    for (uint32_t j = threadIdx.x; j < (1u << n_hi); j += blockDim.x) {
        unsigned long long off = 0;
        for (int b = 0; b < n_hi; ++b)
            if ((j >> b) & 1u) off |= 1ull << sw.pos[c + b];
        hi_off[j] = off;
    }
    // base address of this tile: spread the tile index over the non-tile bit positions
    unsigned long long base = blockIdx.x;
    for (int j = 0; j < T; ++j) base = insert_zero64(base, sw.pos[j]);
    double2* st = work + (unsigned long long)blockIdx.y * state_stride;
    __syncthreads();
    const uint32_t n_amp = 1u << T, low_mask = (1u << c) - 1u;
    if (sw.init) {
        for (uint32_t j = threadIdx.x; j < n_amp; j += blockDim.x)
            s[j] = make_double2((j == 0 && base == 0) ? 1.0 : 0.0, 0.0);  // swz(0) == 0
    } else {
        // cp.async (LDGSTS): every thread has all of its 16-byte loads in flight at once, the
        // data goes straight to shared memory (L1 bypassed)
        // j = tid + k * 256: the low 8 bits (hence the swizzle term and, for c <= 8, the low-run
        // offset) are per-thread constants; only the hi_off index advances with k
        const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s);
        const uint32_t xr = (threadIdx.x >> 3) & 7u;
        for (uint32_t j = threadIdx.x; j < n_amp; j += blockDim.x) {
            const double2* src = st + (base | hi_off[j >> c] | (j & low_mask));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_base + ((j ^ xr) << 4)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    {   // stage the program while the tile is in flight
        const int n0 = (sw.op_end - sw.op_begin) < plan.n_stage ? (sw.op_end - sw.op_begin) : plan.n_stage;
        stage_ops(so, plan.ops, sw.op_begin, n0, plan.mats, digits);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    apply_ops(s, T, so, plan.n_stage, plan.ops, sw.op_begin, sw.op_end, plan.mats, digits, true);
    {
        const uint32_t xr = (threadIdx.x >> 3) & 7u;  // blockDim.x == 256: bits 3-5 of j never change
#pragma unroll 4
        for (uint32_t j = threadIdx.x; j < n_amp; j += blockDim.x)
            __stcs(st + (base | hi_off[j >> c] | (j & low_mask)), s[j ^ xr]);
    }
}

// Pipelined, persistent form of the sweep for states that do not fit L2: one CTA per SM keeps a
// ring of PIPE_STAGES tiles in shared memory.  While tile i is transformed in place, the loads of
// tiles i+1 .. i+PIPE_STAGES-1 are in flight (cp.async groups) and the stores of tile i-1 drain,
// so HBM stays busy during the gate passes.  Tiles are claimed from an atomic counter (a static
// split of a streaming kernel leaves ~15 % on the table on B200, see knit.cu).
#define PIPE_STAGES 3
#define PIPE_THREADS 512
__global__ void __launch_bounds__(PIPE_THREADS, 1) sim_sweep_pipe_kernel(PlanDev plan, SweepDev sw,
                                                                const int32_t* __restrict__ labels,
                                                                int inst_base, double2* __restrict__ work,
                                                                unsigned long long state_stride,
                                                                unsigned long long tiles_per_inst,
                                                                unsigned long long n_work,
                                                                unsigned long long* __restrict__ counter) {
    const int T = sw.n_tile, c = sw.n_low, n_hi = T - c;
    const size_t stage_bytes = (size_t)16 << T;
    unsigned long long* hi_off = reinterpret_cast<unsigned long long*>(smem_raw + PIPE_STAGES * stage_bytes);
    StagedOp* so =
        reinterpret_cast<StagedOp*>(smem_raw + PIPE_STAGES * stage_bytes + ((((size_t)8 << n_hi) + 15) & ~(size_t)15));
    __shared__ int digits[QCK_MAX_DIGITS];
    __shared__ unsigned long long claimed[PIPE_STAGES];
    const uint32_t n_amp = 1u << T, low_mask = (1u << c) - 1u;
    const unsigned long long NONE = ~0ull;

    for (uint32_t j = threadIdx.x; j < (1u << n_hi); j += blockDim.x) {
        unsigned long long off = 0;
        for (int b = 0; b < n_hi; ++b)
            if ((j >> b) & 1u) off |= 1ull << sw.pos[c + b];
        hi_off[j] = off;
    }
    auto tile_base = [&](unsigned long long w, double2*& st) {
        const unsigned long long inst = w / tiles_per_inst;
        unsigned long long base = w - inst * tiles_per_inst;
        for (int j = 0; j < T; ++j) base = insert_zero64(base, sw.pos[j]);
        st = work + inst * state_stride;
        return base;
    };
    auto issue_load = [&](int stage) {  // all threads; claimed[stage] already visible
        const unsigned long long w = claimed[stage];
        if (w != NONE && !sw.init) {
            double2* st;
            const unsigned long long base = tile_base(w, st);
            const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem_raw + stage * stage_bytes);
            for (uint32_t j = threadIdx.x; j < n_amp; j += blockDim.x) {
                const double2* src = st + (base | hi_off[j >> c] | (j & low_mask));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_base + swz(j) * 16u), "l"(src)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // always: keeps the group count uniform
    };
    // prologue: claim and start loading the first PIPE_STAGES - 1 tiles
    if (threadIdx.x == 0) {  // one thread: claims must be ordered (NONE only ever follows valid tiles)
        for (int st = 0; st < PIPE_STAGES - 1; ++st) {
            const unsigned long long w = atomicAdd(counter, 1ull);
            claimed[st] = w < n_work ? w : NONE;
        }
    }
    __syncthreads();
#pragma unroll
    for (int st = 0; st < PIPE_STAGES - 1; ++st) issue_load(st);
    long long staged_inst = -1;
    for (int k = 0;; ++k) {
        const int cur = k % PIPE_STAGES, nxt = (k + PIPE_STAGES - 1) % PIPE_STAGES;
        if (threadIdx.x == 0) {
            const unsigned long long w = atomicAdd(counter, 1ull);
            claimed[nxt] = w < n_work ? w : NONE;
        }
        __syncthreads();  // claimed[nxt] visible; everyone finished storing the tile that lived in nxt
        issue_load(nxt);
        asm volatile("cp.async.wait_group %0;" ::"n"(PIPE_STAGES - 1) : "memory");
        __syncthreads();  // tile `cur` resident for all threads
        const unsigned long long w = claimed[cur];
        if (w == NONE) break;  // tiles are claimed in order: nothing later is pending either
        double2* s = reinterpret_cast<double2*>(smem_raw + cur * stage_bytes);
        double2* st;
        const unsigned long long base = tile_base(w, st);
        const long long inst = (long long)(w / tiles_per_inst);
        if (sw.init) {
            for (uint32_t j = threadIdx.x; j < n_amp; j += blockDim.x)
                s[j] = make_double2((j == 0 && base == 0) ? 1.0 : 0.0, 0.0);  // swz(0) == 0
            __syncthreads();
        }
        if (inst != staged_inst) {  // uniform; the matrices depend on the instance's label digits
            if (threadIdx.x == 0) decode_digits(plan, labels[inst_base + inst], digits);
            __syncthreads();
            stage_ops(so, plan.ops, sw.op_begin, sw.op_end - sw.op_begin, plan.mats, digits);
            __syncthreads();
            staged_inst = inst;
        }
        for (int i = 0; i < sw.op_end - sw.op_begin;) {
            if (so[i].w0.x == QCK_OP_CLUSTER) {
                run_cluster(s, T, so, i, plan.mats);
                i += 1 + so[i].w0.y;
            } else {
                run_single(s, T, so[i], plan.mats);
                ++i;
            }
        }
        for (uint32_t j = threadIdx.x; j < n_amp; j += blockDim.x)
            __stcs(st + (base | hi_off[j >> c] | (j & low_mask)), s[swz(j)]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256) fold_probs_kernel(PlanDev plan, const int32_t* __restrict__ labels,
                                                         int inst_base, const double2* __restrict__ work,
                                                         unsigned long long state_stride,
                                                         double* __restrict__ out, long long row_stride) {
    const int label = labels[inst_base + blockIdx.y];
    const double2* st = work + (unsigned long long)blockIdx.y * state_stride;
    double* row = out + (long long)label * row_stride;
    const uint64_t n_out = 1ull << plan.n_out_bits;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out;
         o += (uint64_t)gridDim.x * blockDim.x)
        row[o] = fold_entry(plan, o, [&](uint64_t idx) { return __ldcs(st + idx); });
}

// ------------------------------------------------------------------ host side
static int validate_plan(qck_handle* h, const qck_sim_plan* plan) {
    if (!plan) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "plan is NULL");
    if (plan->n_state_qubits < 1 || plan->n_state_qubits > 40)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "n_state_qubits=%d out of range [1,40]", plan->n_state_qubits);
    if (plan->n_sweeps < 1 || !plan->sweeps) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "plan has no sweeps");
    if (plan->n_digits < 0 || plan->n_digits > QCK_MAX_DIGITS)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "n_digits=%d out of range", plan->n_digits);
    if (plan->n_out_bits < 0 || plan->n_out_bits > QCK_MAX_OUT_BITS)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "n_out_bits=%d out of range", plan->n_out_bits);
    for (int k = 0; k < plan->n_digits; ++k)
        if (plan->radix[k] < 1) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "radix[%d]=%d", k, plan->radix[k]);
    for (int j = 0; j < plan->n_out_bits; ++j)
        if (plan->out_pos[j] >= plan->n_state_qubits)
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "out_pos[%d]=%d >= n_state_qubits", j, plan->out_pos[j]);
    for (int i = 0; i < plan->n_sweeps; ++i) {
        const qck_sweep& sw = plan->sweeps[i];
        if (sw.n_tile < 1 || sw.n_tile > QCK_MAX_TILE_QUBITS || sw.n_tile > plan->n_state_qubits)
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "sweep %d: n_tile=%d invalid", i, sw.n_tile);
        for (int j = 0; j < sw.n_tile; ++j) {
            if (sw.pos[j] < 0 || sw.pos[j] >= plan->n_state_qubits || (j > 0 && sw.pos[j] <= sw.pos[j - 1]))
                QCK_FAIL(h, QCK_ERR_INVALID_ARG, "sweep %d: tile positions must be ascending and < n_state", i);
        }
        if (sw.op_begin < 0 || sw.op_end < sw.op_begin)
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "sweep %d: bad op range", i);
    }
    if ((plan->n_sweeps > 0 && plan->sweeps[0].op_end > plan->sweeps[0].op_begin) && (!plan->d_ops || !plan->d_mats))
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "plan has ops but d_ops/d_mats is NULL");
    return QCK_OK;
}

static PlanDev to_dev(const qck_sim_plan* plan) {
    PlanDev d;
    memset(&d, 0, sizeof(d));
    d.n_state = plan->n_state_qubits;
    d.ops = plan->d_ops;
    d.mats = plan->d_mats;
    d.n_digits = plan->n_digits;
    for (int k = 0; k < QCK_MAX_DIGITS; ++k) d.radix[k] = k < plan->n_digits ? plan->radix[k] : 1;
    d.n_out_bits = plan->n_out_bits;
    for (int j = 0; j < QCK_MAX_OUT_BITS; ++j) d.out_pos[j] = j < plan->n_out_bits ? plan->out_pos[j] : -1;
    d.out_ident = 0;
    while (d.out_ident < plan->n_out_bits && plan->out_pos[d.out_ident] == d.out_ident) ++d.out_ident;
    d.sum_mask = plan->sum_mask;
    d.sign_mask = plan->sign_mask;
    return d;
}

static SweepDev sweep_dev(const qck_sweep& sw, bool init) {
    SweepDev s;
    memset(&s, 0, sizeof(s));
    s.n_tile = sw.n_tile;
    s.op_begin = sw.op_begin;
    s.op_end = sw.op_end;
    s.init = init ? 1 : 0;
    int c = 0;
    while (c < sw.n_tile && sw.pos[c] == c) ++c;
    s.n_low = c;
    for (int j = 0; j < sw.n_tile; ++j) s.pos[j] = sw.pos[j];
    return s;
}

// records to stage for an op range: everything when it fits, else the largest chunk (a cluster of
// up to QCK_MAX_CLUSTER_OPS members + header must fit)
static int stage_records(int n_ops) {
    int n = n_ops < QCK_STAGE_OPS ? n_ops : QCK_STAGE_OPS;
    return n < 1 ? 1 : n;
}

static bool is_onchip(const qck_sim_plan* plan) {
    return plan->n_sweeps == 1 && plan->sweeps[0].n_tile == plan->n_state_qubits;
}

int qck_ensure_partials(qck_handle* h, size_t count);  // api.cu

static int run_sweeps(qck_handle* h, const qck_sim_plan* plan, const PlanDev& pd, const int32_t* d_labels,
                      int inst_base, int batch, double2* work, unsigned long long state_stride,
                      cudaStream_t st) {
    for (int i = 0; i < plan->n_sweeps; ++i) {
        SweepDev sd = sweep_dev(plan->sweeps[i], i == 0);
        if (sd.n_tile - sd.n_low > 10)
            QCK_FAIL(h, QCK_ERR_UNSUPPORTED, "sweep %d: more than 10 non-contiguous tile bits", i);
        PlanDev pdl = pd;
        pdl.n_stage = stage_records(sd.op_end - sd.op_begin);
        const size_t aux = ((((size_t)8 << (sd.n_tile - sd.n_low)) + 15) & ~(size_t)15) + sizeof(StagedOp) * pdl.n_stage;
        size_t smem = ((size_t)16 << sd.n_tile) + aux;
        if ((int)smem + 1024 > h->max_smem_optin)
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "sweep %d: tile of 2^%d amplitudes does not fit shared memory", i,
                     sd.n_tile);
        unsigned long long tiles = 1ull << (plan->n_state_qubits - sd.n_tile);
        if (tiles > 0x7fffffffull) QCK_FAIL(h, QCK_ERR_UNSUPPORTED, "too many tiles");
        // pipelined persistent kernel when there is enough work to fill the machine several times
        const size_t pipe_smem = PIPE_STAGES * ((size_t)16 << sd.n_tile) + aux;
        int want_pipe = 0;  // measured slower than the plain kernel on B200 (305 vs 229 ms at 32 qubits);
        //                     kept behind QCK_SIM_PIPE=1 for tuning
        if (const char* env = getenv("QCK_SIM_PIPE")) want_pipe = atoi(env);
        const bool pipe = want_pipe && tiles * (unsigned long long)batch >= 8ull * h->sm_count &&
                          (sd.op_end - sd.op_begin) <= pdl.n_stage && (int)pipe_smem + 1024 <= h->max_smem_optin;
        if (pipe) {
            int rc = qck_ensure_partials(h, 8);
            if (rc) return rc;
            unsigned long long* counter = reinterpret_cast<unsigned long long*>(h->d_partials);
            QCK_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
            sim_sweep_pipe_kernel<<<h->sm_count, PIPE_THREADS, pipe_smem, st>>>(pdl, sd, d_labels, inst_base, work, state_stride,
                                                                        tiles, tiles * (unsigned long long)batch, counter);
            QCK_CHECK_LAUNCH(h);
            continue;
        }
        dim3 grid((unsigned)tiles, (unsigned)batch);
        // few tiles (L2-resident states): latency bound, use all 256 threads per tile
        int sweep_threads = tiles * (unsigned long long)batch >= 4ull * h->sm_count ? 128 : 256;  // 3 CTAs per SM de-phase load / compute / store (256 threads: 2 phase-locked CTAs);
        //                           QCK_SWEEP_THREADS overrides (tuning knob; multiple of 64)
        if (const char* env = getenv("QCK_SWEEP_THREADS")) {
            int v = atoi(env);
            if (v >= 64 && v <= 256 && v % 64 == 0) sweep_threads = v;
        }
        sim_sweep_kernel<<<grid, sweep_threads, smem, st>>>(pdl, sd, d_labels, inst_base, work, state_stride);
        QCK_CHECK_LAUNCH(h);
    }
    return QCK_OK;
}

extern "C" int qck_sim_fragments(qck_handle* h, const qck_sim_plan* plan, const int32_t* d_labels,
                                 int64_t n_instances, double* d_out, int64_t out_row_stride, void* d_work,
                                 size_t work_bytes, qck_stream stream) {
    if (!h) return QCK_ERR_INVALID_ARG;
    int rc = validate_plan(h, plan);
    if (rc) return rc;
    if (n_instances == 0) return QCK_OK;
    if (n_instances < 0 || !d_labels || !d_out) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "bad instance list / output");
    if (out_row_stride < (1ll << plan->n_out_bits))
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "out_row_stride smaller than the row (2^%d)", plan->n_out_bits);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    PlanDev pd = to_dev(plan);
    if (is_onchip(plan)) {
        const int N = plan->n_state_qubits;
        const qck_sweep& sw0 = plan->sweeps[0];
        pd.n_stage = stage_records(sw0.op_end - sw0.op_begin);
        size_t smem = ((size_t)16 << N) + sizeof(StagedOp) * pd.n_stage;
        if ((int)smem + 1024 > h->max_smem_optin)
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "on-chip plan with %d qubits does not fit shared memory", N);
        int threads = 1 << (N > 3 ? N - 3 : 0);  // one thread per group of 8 amplitudes
        if (threads < 32) threads = 32;
        if (threads > 256) threads = 256;
        const qck_sweep& sw = plan->sweeps[0];
        for (int64_t done = 0; done < n_instances;) {
            int64_t batch = n_instances - done;
            if (batch > (1ll << 30)) batch = 1ll << 30;
            sim_onchip_kernel<<<(unsigned)batch, threads, smem, st>>>(pd, sw.op_begin, sw.op_end, d_labels + done,
                                                                       d_out, (long long)out_row_stride);
            QCK_CHECK_LAUNCH(h);
            done += batch;
        }
        return QCK_OK;
    }
    // streaming regime
    const unsigned long long state_amps = 1ull << plan->n_state_qubits;
    const size_t state_bytes = (size_t)state_amps * 16;
    if (!d_work || work_bytes < state_bytes)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "streaming regime needs >= %zu bytes of work space, got %zu", state_bytes,
                 work_bytes);
    int64_t cap = (int64_t)(work_bytes / state_bytes);
    if (cap > 65535) cap = 65535;
    for (int64_t done = 0; done < n_instances;) {
        int batch = (int)((n_instances - done) < cap ? (n_instances - done) : cap);
        rc = run_sweeps(h, plan, pd, d_labels, (int)done, batch, (double2*)d_work, state_amps, st);
        if (rc) return rc;
        unsigned long long n_out = 1ull << plan->n_out_bits;
        unsigned gx = (unsigned)((n_out + 255) / 256 < 148ull * 16 ? (n_out + 255) / 256 : 148ull * 16);
        fold_probs_kernel<<<dim3(gx, batch), 256, 0, st>>>(pd, d_labels, (int)done, (const double2*)d_work,
                                                             state_amps, d_out, (long long)out_row_stride);
        QCK_CHECK_LAUNCH(h);
        done += batch;
    }
    return QCK_OK;
}

// Opt-in shared-memory limits are per-function process state: set them ONCE to the device maximum
// (at handle creation) - setting them per launch races between host threads.
int qck_sim_init(qck_handle* h) {
    QCK_CUDA(h, qck_allow_max_smem(sim_onchip_kernel, h->max_smem_optin));
    QCK_CUDA(h, qck_allow_max_smem(sim_onchip_group_kernel, h->max_smem_optin));
    QCK_CUDA(h, qck_allow_max_smem(sim_sweep_kernel, h->max_smem_optin));
    QCK_CUDA(h, qck_allow_max_smem(sim_sweep_pipe_kernel, h->max_smem_optin));
    return QCK_OK;
}

static int ensure_side_streams(qck_handle* h) {
    if (h->side_ready) return QCK_OK;
    for (int i = 0; i < QCK_SIDE_STREAMS; ++i) {
        QCK_CUDA(h, cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
        QCK_CUDA(h, cudaEventCreateWithFlags(&h->side_done[i], cudaEventDisableTiming));
    }
    QCK_CUDA(h, cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
    h->side_ready = 1;
    return QCK_OK;
}

static int launch_group(qck_handle* h, const qck_sim_plan* plans, const int* idx, int n, const int32_t* const* d_labels,
                        const int64_t* n_instances, double* d_out, int64_t out_row_stride, cudaStream_t st) {
    GroupDev G;
    memset(&G, 0, sizeof(G));
    const qck_sim_plan& p0 = plans[idx[0]];
    G.n_state = p0.n_state_qubits;
    G.n_vars = n;
    G.n_digits = p0.n_digits;
    G.ops = p0.d_ops;
    G.mats = p0.d_mats;
    for (int k = 0; k < QCK_MAX_DIGITS; ++k) G.radix[k] = k < p0.n_digits ? p0.radix[k] : 1;
    long long ctas = 0;
    int max_ops = 1;
    for (int i = 0; i < n; ++i) {
        const qck_sim_plan& p = plans[idx[i]];
        if (p.d_ops != p0.d_ops || p.d_mats != p0.d_mats || p.n_digits != p0.n_digits)
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "plans of one batch must share the program blob and the label radices");
        if (out_row_stride < (1ll << p.n_out_bits))
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "out_row_stride smaller than the row (2^%d)", p.n_out_bits);
        PlanDev pd = to_dev(&p);
        PlanVarDev& v = G.var[i];
        v.op_begin = p.sweeps[0].op_begin;
        v.op_end = p.sweeps[0].op_end;
        v.n_out_bits = pd.n_out_bits;
        v.out_ident = pd.out_ident;
        for (int j = 0; j < QCK_MAX_OUT_BITS; ++j) v.out_pos[j] = (signed char)pd.out_pos[j];
        v.sum_mask = pd.sum_mask;
        v.sign_mask = pd.sign_mask;
        v.labels = d_labels[idx[i]];
        v.cta_begin = (int)ctas;
        ctas += n_instances[idx[i]];
        if (v.op_end - v.op_begin > max_ops) max_ops = v.op_end - v.op_begin;
    }
    if (ctas > 0x7fffffffll) QCK_FAIL(h, QCK_ERR_UNSUPPORTED, "too many instances in one group");
    G.n_stage = stage_records(max_ops);
    const int N = G.n_state;
    size_t smem = ((size_t)16 << N) + sizeof(StagedOp) * G.n_stage;
    if ((int)smem + 2048 > h->max_smem_optin)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "on-chip plan with %d qubits does not fit shared memory", N);
    int threads = 1 << (N > 3 ? N - 3 : 0);
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    sim_onchip_group_kernel<<<(unsigned)ctas, threads, smem, st>>>(G, d_out, (long long)out_row_stride);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

extern "C" int qck_sim_fragments_batch(qck_handle* h, int n_plans, const qck_sim_plan* plans,
                                       const int32_t* const* d_labels, const int64_t* n_instances, double* d_out,
                                       int64_t out_row_stride, void* d_work, size_t work_bytes, qck_stream stream) {
    if (!h) return QCK_ERR_INVALID_ARG;
    if (n_plans < 0 || (n_plans > 0 && (!plans || !d_labels || !n_instances || !d_out)))
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "bad plan list");
    for (int i = 0; i < n_plans; ++i) {
        int rc = validate_plan(h, &plans[i]);
        if (rc) return rc;
        if (n_instances[i] < 0 || (n_instances[i] > 0 && !d_labels[i])) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "bad instance list");
    }
    DeviceGuard guard(h->device);
    cudaStream_t main_st = (cudaStream_t)stream;
    // on-chip plans grouped by state size -> one launch per group (chunks of QCK_GROUP_MAX)
    int n_groups = 0;
    for (int N = 1; N <= QCK_MAX_TILE_QUBITS; ++N) {
        int cnt = 0;
        for (int i = 0; i < n_plans; ++i)
            if (n_instances[i] > 0 && is_onchip(&plans[i]) && plans[i].n_state_qubits == N) ++cnt;
        n_groups += (cnt + QCK_GROUP_MAX - 1) / QCK_GROUP_MAX;
    }
    const bool fan = n_groups >= 2;
    int used = 0, k = 0;
    if (fan) {
        int rc = ensure_side_streams(h);
        if (rc) return rc;
        QCK_CUDA(h, cudaEventRecord(h->fork, main_st));
    }
    for (int N = QCK_MAX_TILE_QUBITS; N >= 1; --N) {  // largest states first: they run longest
        int idx[QCK_GROUP_MAX], cnt = 0;
        for (int i = 0; i <= n_plans; ++i) {
            const bool take = i < n_plans && n_instances[i] > 0 && is_onchip(&plans[i]) && plans[i].n_state_qubits == N;
            if (take) idx[cnt++] = i;
            if (cnt == QCK_GROUP_MAX || (i == n_plans && cnt > 0)) {
                cudaStream_t st = main_st;
                if (fan) {
                    const int slot = k++ % QCK_SIDE_STREAMS;
                    st = h->side[slot];
                    if (slot >= used) {  // first use in this call: order after the fork point
                        QCK_CUDA(h, cudaStreamWaitEvent(st, h->fork, 0));
                        used = slot + 1;
                    }
                }
                int rc = launch_group(h, plans, idx, cnt, d_labels, n_instances, d_out, out_row_stride, st);
                if (rc) return rc;
                cnt = 0;
            }
        }
    }
    for (int s = 0; s < used; ++s) {  // join
        QCK_CUDA(h, cudaEventRecord(h->side_done[s], h->side[s]));
        QCK_CUDA(h, cudaStreamWaitEvent(main_st, h->side_done[s], 0));
    }
    // streaming plans: one after the other on the caller's stream, sharing d_work
    for (int i = 0; i < n_plans; ++i) {
        if (n_instances[i] <= 0 || is_onchip(&plans[i])) continue;
        int rc = qck_sim_fragments(h, &plans[i], d_labels[i], n_instances[i], d_out, out_row_stride, d_work, work_bytes,
                                   (qck_stream)main_st);
        if (rc) return rc;
    }
    return QCK_OK;
}

extern "C" int qck_sim_statevector(qck_handle* h, const qck_sim_plan* plan, int32_t label, void* d_state,
                                   size_t state_bytes, qck_stream stream) {
    if (!h) return QCK_ERR_INVALID_ARG;
    int rc = validate_plan(h, plan);
    if (rc) return rc;
    if (is_onchip(plan) && plan->n_state_qubits > 0) {
        // an on-chip plan is also a valid one-sweep streaming plan
    }
    const unsigned long long state_amps = 1ull << plan->n_state_qubits;
    if (!d_state || state_bytes < state_amps * 16)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "state buffer too small: need %llu bytes", state_amps * 16);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    // the label travels through a one-element device list kept in the handle scratch
    int32_t* d_label = reinterpret_cast<int32_t*>(h->d_partials);
    QCK_CUDA(h, cudaMemcpyAsync(d_label, &label, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PlanDev pd = to_dev(plan);
    return run_sweeps(h, plan, pd, d_label, 0, 1, (double2*)d_state, state_amps, st);
}
