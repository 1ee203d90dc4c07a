// Reference-faithful knit (ACCURACY > 0) as one fused evaluation per output bitstring.
//
// The reference prunes |v| <= ACCURACY after EVERY QuasiDistr operation
// (third_party/qvm/qvm/quasi_distr.py:3,7-10): from_counts, each merge fold
// (virtual_circuit.py:216-228), each half of split, and every + / - / scalar * of the per-gate
// knit formulas (virtual_gates.py:105-124,179-194,262-286), which are evaluated left to right, one
// virtual gate per level from the last to the first (virtual_circuit.py:59-68).  All of it is
// elementwise in the output bitstring x: the value of level k at (x, config bits c_0..c_{k-1}) is a
// fixed expression of the level-(k+1) values of its n_k children at c_k = 0 and 1.  So instead of
// materialising 2^(n_cl+K)-entry dictionaries per label, every thread evaluates that expression
// tree for its own x by recursion over the label digits, reading the UNFOLDED fragment tables
// (config bits kept as extra row bits).  Subtrees with c_k = 1 under a variant that measures
// nothing are identically zero and are skipped.
//
// Parallelism: level 0 is split off - one thread per (x, variant of gate 0, c_0) evaluates the
// subtree below it, a second small kernel applies gate 0's formula.
#include "qck_common.cuh"

#define FA_MAXF QCK_MAX_FRAGMENTS
#define FA_MAXK QCK_MAX_DIGITS
#define FA_MAXV QCK_MAX_VARIANTS

struct FaithParams {
    int n_frag, K, n_out_bits;
    const double* table[FA_MAXF];
    unsigned long long mask[FA_MAXF];
    long long row_stride[FA_MAXF];
    int m_bits[FA_MAXF];
    int radix[FA_MAXK];
    int form[FA_MAXK];                  // 0: signed chain (move / cz / cx / cy), 1: rzz family
    int degenerate[FA_MAXK];            // rzz: 0 full, 1: |cos| < eps -> r * sin^2, 2: |sin| < eps -> r * cos^2
    double sign[FA_MAXK][FA_MAXV];      // chain: +1 / -1 per variant
    double c1[FA_MAXK], s1[FA_MAXK];    // rzz: cos(m/2), sin(m/2)
    double c2[FA_MAXK], s2[FA_MAXK];    // rzz: cos^2, sin^2 (as Python's ** 2)
    int frag_stride[FA_MAXF][FA_MAXK];  // fragment label = sum_k digit_k * stride
    int cfg_bit[FA_MAXF][FA_MAXK];      // position of gate k's config bit among f's extra row bits, -1: untouched
    unsigned char measures[FA_MAXF][FA_MAXK][FA_MAXV];  // does f's instance measure gate k under variant v
    unsigned char any_measure[FA_MAXK][FA_MAXV];
    double acc;
};

__device__ __forceinline__ double prn(double v, double acc) { return fabs(v) > acc ? v : 0.0; }

struct EvalCtx {
    const FaithParams* P;
    unsigned long long xf[FA_MAXF];  // x restricted to each fragment's output bits (compact)
    int d[FA_MAXK];
};

// merged_l(x, c): left fold over the fragments of prune(p_f) with a prune after each product
__device__ double leaf(EvalCtx& C, unsigned c) {
    const FaithParams& P = *C.P;
    // every set config bit needs a fragment that measures it in this label
    for (int k = 0; k < P.K; ++k)
        if (((c >> k) & 1u) && !P.any_measure[k][C.d[k]]) return 0.0;
    double v = 0.0;
    for (int f = 0; f < P.n_frag; ++f) {
        long long row = 0;
        unsigned long long cf = 0;
        for (int k = 0; k < P.K; ++k) {
            row += (long long)C.d[k] * P.frag_stride[f][k];
            if (((c >> k) & 1u) && P.cfg_bit[f][k] >= 0 && P.measures[f][k][C.d[k]]) cf |= 1ull << P.cfg_bit[f][k];
        }
        const double pf = prn(__ldg(P.table[f] + row * P.row_stride[f] + (C.xf[f] | (cf << P.m_bits[f]))), P.acc);
        v = f == 0 ? pf : prn(v * pf, P.acc);
    }
    return v;
}

__device__ double eval(EvalCtx& C, int k, unsigned c);

// apply gate k's knit formula given a functor child(i, bit) -> level-(k+1) value
template <typename Child>
__device__ __forceinline__ double knit_formula(const FaithParams& P, int k, Child child) {
    const double acc = P.acc;
    if (P.form[k] == 0) {
        double total = 0.0;
        for (int i = 0; i < P.radix[k]; ++i) {
            const double r0 = prn(child(i, 0), acc);
            const double r1 = P.any_measure[k][i] ? prn(child(i, 1), acc) : 0.0;
            const double diff = prn(r0 - r1, acc);
            total = i == 0 ? diff : prn(P.sign[k][i] > 0.0 ? total + diff : total - diff, acc);
        }
        return prn(0.5 * total, acc);
    }
    if (P.degenerate[k] == 1) return prn(prn(child(0, 0), acc) * P.s2[k], acc);
    if (P.degenerate[k] == 2) return prn(prn(child(0, 0), acc) * P.c2[k], acc);
    const double r0 = prn(child(0, 0), acc), r1 = prn(child(1, 0), acc);
    const double r23_0 = prn(child(2, 0) + child(3, 0), acc);
    const double r23_1 = prn((P.any_measure[k][2] ? child(2, 1) : 0.0) + (P.any_measure[k][3] ? child(3, 1) : 0.0), acc);
    const double r45_0 = prn(child(4, 0) + child(5, 0), acc);
    const double r45_1 = prn((P.any_measure[k][4] ? child(4, 1) : 0.0) + (P.any_measure[k][5] ? child(5, 1) : 0.0), acc);
    const double mixed = prn(prn(prn(r23_0 - r23_1, acc) - r45_0, acc) + r45_1, acc);
    const double head = prn(prn(r0 * P.c2[k], acc) + prn(r1 * P.s2[k], acc), acc);
    return prn(head + prn(prn(mixed * P.c1[k], acc) * P.s1[k], acc), acc);
}

__device__ double eval(EvalCtx& C, int k, unsigned c) {
    const FaithParams& P = *C.P;
    if (k == P.K) return leaf(C, c);
    return knit_formula(P, k, [&](int i, int bit) {
        C.d[k] = i;
        return eval(C, k + 1, bit ? (c | (1u << k)) : c);
    });
}

__device__ __forceinline__ void init_ctx(EvalCtx& C, const FaithParams& P, unsigned long long x) {
    C.P = &P;
    for (int f = 0; f < P.n_frag; ++f) C.xf[f] = soft_pext(x, P.mask[f]);
    for (int k = 0; k < FA_MAXK; ++k) C.d[k] = 0;
}

// stage 1: scratch[(i * 2 + bit) << n_out | x] = level-1 value under gate 0's variant i, c_0 = bit
__global__ void __launch_bounds__(128) faith_subtree_kernel(const __grid_constant__ FaithParams P,
                                                            double* __restrict__ scratch) {
    const unsigned long long n = 1ull << P.n_out_bits;
    const unsigned long long total = n * (unsigned long long)(2 * P.radix[0]);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long x = t & (n - 1);
        const int j = (int)(t >> P.n_out_bits), i = j >> 1, bit = j & 1;
        double v = 0.0;
        if (!bit || P.any_measure[0][i]) {
            EvalCtx C;
            init_ctx(C, P, x);
            C.d[0] = i;
            v = eval(C, 1, bit ? 1u : 0u);
        }
        scratch[t] = v;
    }
}

// stage 2: gate 0's formula over the stored subtrees
__global__ void __launch_bounds__(256) faith_top_kernel(const __grid_constant__ FaithParams P,
                                                        const double* __restrict__ scratch, double* __restrict__ out) {
    const unsigned long long n = 1ull << P.n_out_bits;
    for (unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; x < n;
         x += (unsigned long long)gridDim.x * blockDim.x)
        out[x] = knit_formula(P, 0, [&](int i, int bit) { return scratch[((unsigned long long)(2 * i + bit) << P.n_out_bits) | x]; });
}

// no virtual gates: the pruned merge only
__global__ void __launch_bounds__(256) faith_leaf_kernel(const __grid_constant__ FaithParams P, double* __restrict__ out) {
    const unsigned long long n = 1ull << P.n_out_bits;
    for (unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; x < n;
         x += (unsigned long long)gridDim.x * blockDim.x) {
        EvalCtx C;
        init_ctx(C, P, x);
        out[x] = leaf(C, 0u);
    }
}

int qck_ensure_scratch(qck_handle* h, size_t bytes, void** out);  // api.cu

extern "C" int qck_knit_faithful(qck_handle* h, int n_frag, const double* const* d_tables, const uint64_t* masks,
                                 const int64_t* row_strides, int n_out_bits, int n_gates, const qck_faithful_gate* gates,
                                 const int32_t* frag_stride, const int32_t* cfg_bit, const uint8_t* measures,
                                 double accuracy, double* d_out, qck_stream stream) {
    if (!h) return QCK_ERR_INVALID_ARG;
    if (n_frag < 1 || n_frag > FA_MAXF) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "n_frag=%d out of range", n_frag);
    if (n_gates < 0 || n_gates > FA_MAXK) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "n_gates=%d out of range", n_gates);
    if (!d_tables || !masks || !row_strides || !d_out || (n_gates > 0 && (!gates || !frag_stride || !cfg_bit || !measures)))
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    if (n_out_bits < 0 || n_out_bits > 30) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "n_out_bits=%d out of range", n_out_bits);
    if (!(accuracy >= 0.0)) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "accuracy must be >= 0");
    FaithParams P;
    memset(&P, 0, sizeof(P));
    P.n_frag = n_frag;
    P.K = n_gates;
    P.n_out_bits = n_out_bits;
    P.acc = accuracy;
    uint64_t seen = 0;
    for (int f = 0; f < n_frag; ++f) {
        if (!d_tables[f]) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "table %d is NULL", f);
        if (masks[f] & seen) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "fragment masks overlap");
        seen |= masks[f];
        P.table[f] = d_tables[f];
        P.mask[f] = masks[f];
        P.row_stride[f] = row_strides[f];
        P.m_bits[f] = __builtin_popcountll(masks[f]);
    }
    for (int k = 0; k < n_gates; ++k) {
        const qck_faithful_gate& g = gates[k];
        if (g.n_variants < 1 || g.n_variants > FA_MAXV) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "gate %d: n_variants=%d", k, g.n_variants);
        if (g.form != 0 && g.form != 1) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "gate %d: unknown knit form %d", k, g.form);
        if (g.form == 1 && g.degenerate == 0 && g.n_variants != 6)
            QCK_FAIL(h, QCK_ERR_INVALID_ARG, "gate %d: the rzz form needs 6 variants", k);
        P.radix[k] = g.n_variants;
        P.form[k] = g.form;
        P.degenerate[k] = g.degenerate;
        for (int v = 0; v < FA_MAXV; ++v) P.sign[k][v] = g.sign[v];
        P.c1[k] = g.cos_half;
        P.s1[k] = g.sin_half;
        P.c2[k] = g.cos_half_sq;
        P.s2[k] = g.sin_half_sq;
        for (int f = 0; f < n_frag; ++f) {
            P.frag_stride[f][k] = frag_stride[f * FA_MAXK + k];
            P.cfg_bit[f][k] = cfg_bit[f * FA_MAXK + k];
            for (int v = 0; v < FA_MAXV; ++v) {
                P.measures[f][k][v] = measures[(f * FA_MAXK + k) * FA_MAXV + v];
                P.any_measure[k][v] |= P.measures[f][k][v];
            }
        }
    }
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long n = 1ull << n_out_bits;
    int grid = (int)((n + 255) / 256 < (unsigned long long)h->sm_count * 8 ? (n + 255) / 256 : (unsigned long long)h->sm_count * 8);
    if (n_gates == 0) {
        faith_leaf_kernel<<<grid, 256, 0, st>>>(P, d_out);
        QCK_CHECK_LAUNCH(h);
        return QCK_OK;
    }
    void* scratch = nullptr;
    const unsigned long long total = n * (unsigned long long)(2 * P.radix[0]);
    int rc = qck_ensure_scratch(h, total * sizeof(double), &scratch);
    if (rc) return rc;
    // recursion depth <= K: make sure the per-thread stack can hold it
    size_t stack = 0;
    cudaDeviceGetLimit(&stack, cudaLimitStackSize);
    if (stack < 4096) QCK_CUDA(h, cudaDeviceSetLimit(cudaLimitStackSize, 4096));
    unsigned long long want = (total + 127) / 128;
    int sgrid = (int)(want < (unsigned long long)h->sm_count * 32 ? want : (unsigned long long)h->sm_count * 32);
    faith_subtree_kernel<<<sgrid, 128, 0, st>>>(P, (double*)scratch);
    QCK_CHECK_LAUNCH(h);
    faith_top_kernel<<<grid, 256, 0, st>>>(P, (const double*)scratch, d_out);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}
