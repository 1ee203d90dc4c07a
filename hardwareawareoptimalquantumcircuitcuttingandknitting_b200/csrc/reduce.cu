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

// ---- cross-rank reduction of a qck_stats over peer-mapped mailboxes
// A sharded result needs one exchange per step: (sum, min, ...) of every rank's slice.  Through NCCL that is a
// collective launch plus the small kernels that combine the gathered values (measured: 46 us of a 2.4 ms step
// on two GPUs, of a 0.64 ms step on eight).  Here every rank writes its four doubles straight into a slot of
// every peer's mailbox (NVLink stores into cudaIpc-mapped memory), publishes a sequence number, waits until its
// own mailbox holds the current sequence number from every rank and adds the slots in rank order (the same
// bits on every rank).  One launch of one warp; the sequence counter lives on the device, so the launch can be
// captured in a CUDA graph and replayed.  Slots are double buffered by the parity of the sequence number: a
// rank can only be two exchanges ahead of another one after that one has finished reading.
// (StatsSlot, ExchangeParams and the warp-level protocol live in qck_common.cuh: knit_outer runs the same exchange
// in its tail)
__global__ void __launch_bounds__(32) stats_exchange_kernel(const __grid_constant__ ExchangeParams P, qck_stats* stats,
                                                            unsigned long long* seq_counter) {
    stats_exchange_warp(P, stats, seq_counter);
}

extern "C" size_t qck_stats_exchange_mailbox_bytes(int world) { return sizeof(StatsSlot) * 2 * (size_t)world + 64; }

extern "C" int qck_stats_exchange(qck_handle* h, qck_stats* d_stats, int rank, int world, void* const* d_mailboxes,
                                  qck_stream stream) {
    if (!h) return QCK_ERR_INVALID_ARG;
    if (!d_stats || !d_mailboxes || world < 1 || world > QCK_MAX_RANKS || rank < 0 || rank >= world)
        QCK_FAIL(h, QCK_ERR_INVALID_ARG, "bad stats exchange arguments");
    DeviceGuard guard(h->device);
    ExchangeParams P;
    memset(&P, 0, sizeof(P));
    for (int r = 0; r < world; ++r) {
        if (!d_mailboxes[r]) QCK_FAIL(h, QCK_ERR_INVALID_ARG, "mailbox of rank %d is NULL", r);
        P.box[r] = reinterpret_cast<StatsSlot*>(reinterpret_cast<char*>(d_mailboxes[r]) + 64);
    }
    P.rank = rank;
    P.world = world;
    // the sequence counter is the first 8 bytes of this rank's own mailbox allocation
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(d_mailboxes[rank]);
    stats_exchange_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, d_stats, counter);
    QCK_CHECK_LAUNCH(h);
    return QCK_OK;
}

extern "C" int qck_mem_zero(qck_handle* h, void* d_ptr, size_t bytes, qck_stream stream) {
    if (!h || !d_ptr) return QCK_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    QCK_CUDA(h, cudaMemsetAsync(d_ptr, 0, bytes, (cudaStream_t)stream));
    return QCK_OK;
}
