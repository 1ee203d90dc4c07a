// Reductions over dense distributions (sm_100a): statistics and Hellinger sums
// (nearest_probability_distribution lives in npd.cu).
#include "qck_common.cuh"

int qck_ensure_partials(qck_handle* h, size_t count);

// ---- statistics: sum, min, sum sqrt(max(v,0)), nnz over entries with |v| > acc
__global__ void __launch_bounds__(256) stats_kernel(const double* __restrict__ p, unsigned long long n, double acc,
                                                    double* __restrict__ partials) {
    double s = 0.0, m = INFINITY, q = 0.0, z = 0.0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        double v = p[i];
        if (fabs(v) > acc) {
            s += v;
            m = fmin(m, v);
            q += v > 0.0 ? sqrt(v) : 0.0;
            z += 1.0;
        }
    }
    __shared__ double red[4][8];
    s = warp_sum(s);
    m = warp_min(m);
    q = warp_sum(q);
    z = warp_sum(z);
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        red[0][w] = s;
        red[1][w] = m;
        red[2][w] = q;
        red[3][w] = z;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ss = 0.0, mm = INFINITY, qq = 0.0, zz = 0.0;
        for (int w = 0; w < 8; ++w) {
            ss += red[0][w];
            mm = fmin(mm, red[1][w]);
            qq += red[2][w];
            zz += red[3][w];
        }
        partials[4 * blockIdx.x + 0] = ss;
        partials[4 * blockIdx.x + 1] = mm;
        partials[4 * blockIdx.x + 2] = qq;
        partials[4 * blockIdx.x + 3] = zz;
    }
}

__global__ void stats_final_kernel(const double* __restrict__ partials, int n, qck_stats* out) {
    double s = 0.0, m = INFINITY, q = 0.0, z = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) {
        s += partials[4 * i + 0];
        m = fmin(m, partials[4 * i + 1]);
        q += partials[4 * i + 2];
        z += partials[4 * i + 3];
    }
    s = warp_sum(s);
    m = warp_min(m);
    q = warp_sum(q);
    z = warp_sum(z);
    if (threadIdx.x == 0) {
        out->sum = s;
        out->min = m;
        out->sum_sqrt = q;
        out->nnz = z;
    }
}

static int reduce_grid(qck_handle* h, unsigned long long n) {
    unsigned long long want = (n + 255) / 256;
    unsigned long long cap = (unsigned long long)h->sm_count * 8;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

extern "C" int qck_stats_dense(qck_handle* h, const double* d_p, uint64_t n, double acc, qck_stats* d_stats,
                               qck_stream stream) {
    if (!h) return QCK_ERR_INVALID_ARG;
    if (!d_p || !d_stats) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    int grid = reduce_grid(h, n);
    int rc = qck_ensure_partials(h, (size_t)grid * 4);
    if (rc) return rc;
    stats_kernel<<<grid, 256, 0, st>>>(d_p, n, acc, h->d_partials);
    QCK_CHECK_LAUNCH(h);
    stats_final_kernel<<<1, 32, 0, st>>>(h->d_partials, grid, d_stats);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

// ---- Hellinger building blocks: sum p, sum q, sum sqrt(p q) over the non-negative parts
__global__ void __launch_bounds__(256) hellinger_kernel(const double* __restrict__ p, const double* __restrict__ q,
                                                        unsigned long long n, double* __restrict__ partials) {
    double sp = 0.0, sq = 0.0, bc = 0.0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        double a = fmax(p[i], 0.0), b = fmax(q[i], 0.0);
        sp += a;
        sq += b;
        bc += sqrt(a * b);
    }
    __shared__ double red[3][8];
    sp = warp_sum(sp);
    sq = warp_sum(sq);
    bc = warp_sum(bc);
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        red[0][w] = sp;
        red[1][w] = sq;
        red[2][w] = bc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int w = 0; w < 8; ++w) {
            a += red[0][w];
            b += red[1][w];
            c += red[2][w];
        }
        partials[3 * blockIdx.x + 0] = a;
        partials[3 * blockIdx.x + 1] = b;
        partials[3 * blockIdx.x + 2] = c;
    }
}

__global__ void sum3_final_kernel(const double* __restrict__ partials, int n, double* out3) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) {
        a += partials[3 * i + 0];
        b += partials[3 * i + 1];
        c += partials[3 * i + 2];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    if (threadIdx.x == 0) {
        out3[0] = a;
        out3[1] = b;
        out3[2] = c;
    }
}

extern "C" int qck_hellinger(qck_handle* h, const double* d_p, const double* d_q, uint64_t n, double* d_result3,
                             qck_stream stream) {
    if (!h) return QCK_ERR_INVALID_ARG;
    if (!d_p || !d_q || !d_result3) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    int grid = reduce_grid(h, n);
    int rc = qck_ensure_partials(h, (size_t)grid * 3);
    if (rc) return rc;
    hellinger_kernel<<<grid, 256, 0, st>>>(d_p, d_q, n, h->d_partials);
    QCK_CHECK_LAUNCH(h);
    sum3_final_kernel<<<1, 32, 0, st>>>(h->d_partials, grid, d_result3);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}
