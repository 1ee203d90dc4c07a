// Dense device forms of the QuasiDistr algebra (third_party/qvm/qvm/quasi_distr.py:45-86)
// with the reference's pruning rule applied after every operation:
//     v = (|v| > acc) ? v : 0          (quasi_distr.py:7-10, ACCURACY)
// They back the per-gate operator API VirtualBinaryGate.knit(results, clbit_idx)
// (virtual_gates.py:39) and the reference-faithful (acc = 1e-5) knit mode, where the
// order of operations and of pruning is part of the result.  All are HBM-bound
// elementwise streams; grids are sized in multiples of the SM count.
#include "qck_common.cuh"

__device__ __forceinline__ double prune1(double v, double acc) { return fabs(v) > acc ? v : 0.0; }

static int ew_grid(qck_handle* h, unsigned long long n) {
    unsigned long long want = (n + 255) / 256;
    unsigned long long cap = (unsigned long long)h->sm_count * 16;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

#define EW_LOOP(i, n)                                                                          \
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); \
         i += (unsigned long long)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(256) qd_prune_kernel(double* v, unsigned long long n, double acc) {
    EW_LOOP(i, n) v[i] = prune1(v[i], acc);
}

__global__ void __launch_bounds__(256) qd_sqrt_kernel(const double* __restrict__ v, double* __restrict__ out,
                                                      unsigned long long n) {
    EW_LOOP(i, n) {
        double x = v[i];
        out[i] = x > 0.0 ? sqrt(x) : 0.0;
    }
}

__global__ void __launch_bounds__(256) qd_axpby_kernel(double a, const double* __restrict__ x, double b,
                                                       const double* __restrict__ y, double* __restrict__ out,
                                                       unsigned long long n, double acc) {
    // a, b in {+1, -1} reproduce the reference's add / sub exactly; a scalar multiply is b == 0
    if (y) {
        EW_LOOP(i, n) out[i] = prune1(a * x[i] + b * y[i], acc);
    } else {
        EW_LOOP(i, n) out[i] = prune1(a * x[i], acc);
    }
}

__global__ void __launch_bounds__(256) qd_split_kernel(const double* __restrict__ v, unsigned long long half, int bit,
                                                       double* __restrict__ lo, double* __restrict__ hi, double acc) {
    EW_LOOP(i, half) {
        unsigned long long i0 = insert_zero64(i, bit);
        if (lo) lo[i] = prune1(v[i0], acc);
        if (hi) hi[i] = prune1(v[i0 | (1ull << bit)], acc);
    }
}

__global__ void __launch_bounds__(256) qd_merge_kernel(const double* __restrict__ a, unsigned long long mask_a,
                                                       const double* __restrict__ b, unsigned long long mask_b,
                                                       double* __restrict__ out, unsigned long long n, double acc) {
    // a key with a bit outside both supports does not exist in the sparse reference
    const unsigned long long outside = ~(mask_a | mask_b);
    EW_LOOP(k, n) out[k] = (k & outside) ? 0.0 : prune1(a[k & mask_a] * b[k & mask_b], acc);
}

struct LevelParams {
    int n_results;
    const double* r[QCK_MAX_VARIANTS];
    double c0[QCK_MAX_VARIANTS], c1[QCK_MAX_VARIANTS];
};

__global__ void __launch_bounds__(256) qd_knit_level_kernel(const __grid_constant__ LevelParams P,
                                                            unsigned long long half, int bit,
                                                            double* __restrict__ out) {
    EW_LOOP(i, half) {
        unsigned long long i0 = insert_zero64(i, bit), i1 = i0 | (1ull << bit);
        double acc = 0.0;
        for (int r = 0; r < P.n_results; ++r) acc += P.c0[r] * P.r[r][i0] + P.c1[r] * P.r[r][i1];
        out[i] = acc;
    }
}

#define QD_PROLOGUE()                      \
    if (!h) return QCK_ERR_INVALID_ARG;    \
    DeviceGuard guard(h->device);          \
    cudaStream_t st = (cudaStream_t)stream

extern "C" int qck_qd_prune(qck_handle* h, double* d_v, uint64_t n, double acc, qck_stream stream) {
    QD_PROLOGUE();
    if (!d_v) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    if (n == 0) return QCK_OK;
    qd_prune_kernel<<<ew_grid(h, n), 256, 0, st>>>(d_v, n, acc);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

extern "C" int qck_qd_sqrt(qck_handle* h, const double* d_v, double* d_out, uint64_t n, qck_stream stream) {
    QD_PROLOGUE();
    if (!d_v || !d_out) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    if (n == 0) return QCK_OK;
    qd_sqrt_kernel<<<ew_grid(h, n), 256, 0, st>>>(d_v, d_out, n);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

extern "C" int qck_qd_axpby(qck_handle* h, double a, const double* d_x, double b, const double* d_y, double* d_out,
                            uint64_t n, double acc, qck_stream stream) {
    QD_PROLOGUE();
    if (!d_x || !d_out) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    if (n == 0) return QCK_OK;
    qd_axpby_kernel<<<ew_grid(h, n), 256, 0, st>>>(a, d_x, b, d_y, d_out, n, acc);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

extern "C" int qck_qd_split(qck_handle* h, const double* d_v, uint64_t n, int bit, double* d_lo, double* d_hi,
                            double acc, qck_stream stream) {
    QD_PROLOGUE();
    if (!d_v || (!d_lo && !d_hi)) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    if (n < 2 || (n & (n - 1))) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "split needs a power-of-two length >= 2");
    if (bit < 0 || (1ull << bit) >= n) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "split bit %d outside the key width", bit);
    qd_split_kernel<<<ew_grid(h, n / 2), 256, 0, st>>>(d_v, n / 2, bit, d_lo, d_hi, acc);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

extern "C" int qck_qd_merge(qck_handle* h, const double* d_a, uint64_t mask_a, const double* d_b, uint64_t mask_b,
                            double* d_out, uint64_t n, double acc, qck_stream stream) {
    QD_PROLOGUE();
    if (!d_a || !d_b || !d_out) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    if (mask_a & mask_b) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "merge: supports overlap (mask_a & mask_b != 0)");
    if (n == 0) return QCK_OK;
    qd_merge_kernel<<<ew_grid(h, n), 256, 0, st>>>(d_a, mask_a, d_b, mask_b, d_out, n, acc);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

extern "C" int qck_qd_knit_level(qck_handle* h, int n_results, const double* const* d_results, uint64_t n, int bit,
                                 const double* coef0, const double* coef1, double* d_out, qck_stream stream) {
    QD_PROLOGUE();
    if (n_results < 1 || n_results > QCK_MAX_VARIANTS)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "n_results=%d out of range [1,%d]", n_results, QCK_MAX_VARIANTS);
    if (!d_results || !coef0 || !coef1 || !d_out) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    if (n < 2 || (n & (n - 1))) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "knit level needs a power-of-two length >= 2");
    if (bit < 0 || (1ull << bit) >= n) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "bit %d outside the key width", bit);
    LevelParams P;
    memset(&P, 0, sizeof(P));
    P.n_results = n_results;
    for (int r = 0; r < n_results; ++r) {
        if (!d_results[r]) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "result %d is NULL", r);
        P.r[r] = d_results[r];
        P.c0[r] = coef0[r];
        P.c1[r] = coef1[r];
    }
    qd_knit_level_kernel<<<ew_grid(h, n / 2), 256, 0, st>>>(P, n / 2, bit, d_out);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}
