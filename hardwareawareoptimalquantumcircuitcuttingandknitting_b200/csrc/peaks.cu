// Roofline denominators that MEASURED_PEAKS.json does not hold ("FP64 / smem peaks ... the
// builder measures them on the box and records them beside each claim"): FP64 FMA issue rate, FP64
// tensor-core (DMMA, mma.sync.m8n8k4.f64) rate and shared-memory load bandwidth of THIS device.  The
// on-chip simulator (state in shared memory / registers) and the label contraction are bound by these,
// not by HBM.  Measurement utilities only - nothing on the hot path calls them.
#include "qck_common.cuh"

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-9;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;  // never true: keeps the chain alive
}

__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double seed) {
    double c0[4] = {0.0, 0.0, 0.0, 0.0}, c1[4] = {0.0, 0.0, 0.0, 0.0};  // two independent accumulator tiles
    double a = seed + threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0[(u & 1) * 2]), "+d"(c0[(u & 1) * 2 + 1]) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c1[(u & 1) * 2]), "+d"(c1[(u & 1) * 2 + 1]) : "d"(a), "d"(b));
        }
    }
    const double s = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_smem_kernel(double* out, int iters) {
    extern __shared__ __align__(16) unsigned char peak_smem[];
    double2* s = reinterpret_cast<double2*>(peak_smem);
    const int n = 8192;  // 128 KiB of double2
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0.0, 0.0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // 16-byte loads, consecutive lanes -> consecutive 16-byte words
            double2 v;  // volatile: the loads must not be hoisted out of the loop (the buffer is never written)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                         : "=d"(v.x), "=d"(v.y)
                         : "r"((unsigned)__cvta_generic_to_shared(s + ((idx + u * 256) & (n - 1)))));
            acc.x += v.x;
            acc.y += v.y;
        }
        idx = (idx + 2048) & (n - 1);
    }
    if (acc.x == 12345.678) out[0] = acc.x + acc.y;
}

// out[0] = FP64 FMA TFLOP/s (2 flop per DFMA), out[1] = DMMA TFLOP/s (512 flop per warp-level m8n8k4),
// out[2] = shared-memory load TB/s (LDS.128, whole device), out[3] = SM count.  Synchronises.
extern "C" int qck_measure_peaks(qck_handle* h, double* out4) {
    if (!h || !out4) return QCK_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    QCK_CUDA(h, cudaFuncSetAttribute(peak_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 << 10));
    cudaEvent_t e0, e1;
    QCK_CUDA(h, cudaEventCreate(&e0));
    QCK_CUDA(h, cudaEventCreate(&e1));
    const int grid = h->sm_count * 8;
    float ms = 0.f;
    double best[3] = {0.0, 0.0, 0.0};
    for (int rep = 0; rep < 4; ++rep) {
        const int it_f = 1 << 14;
        QCK_CUDA(h, cudaEventRecord(e0, 0));
        peak_dfma_kernel<<<grid, 256>>>(h->d_partials, it_f, 0.5);
        QCK_CUDA(h, cudaEventRecord(e1, 0));
        QCK_CUDA(h, cudaEventSynchronize(e1));
        QCK_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        double v = 2.0 * 8.0 * it_f * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep && v > best[0]) best[0] = v;
        const int it_m = 1 << 13;
        QCK_CUDA(h, cudaEventRecord(e0, 0));
        peak_dmma_kernel<<<grid, 256>>>(h->d_partials, it_m, 0.5);
        QCK_CUDA(h, cudaEventRecord(e1, 0));
        QCK_CUDA(h, cudaEventSynchronize(e1));
        QCK_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        v = 512.0 * 8.0 * it_m * 8.0 * grid / (ms * 1e-3) / 1e12;  // 8 mma per iteration, 8 warps per CTA
        if (rep && v > best[1]) best[1] = v;
        const int it_s = 1 << 12;
        QCK_CUDA(h, cudaEventRecord(e0, 0));
        peak_smem_kernel<<<h->sm_count, 256, 128 << 10>>>(h->d_partials, it_s);
        QCK_CUDA(h, cudaEventRecord(e1, 0));
        QCK_CUDA(h, cudaEventSynchronize(e1));
        QCK_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        v = 16.0 * 8.0 * it_s * 256.0 * h->sm_count / (ms * 1e-3) / 1e12;
        if (rep && v > best[2]) best[2] = v;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    QCK_CHECK_LAUNCH(h);
    out4[0] = best[0];
    out4[1] = best[1];
    out4[2] = best[2];
    out4[3] = (double)h->sm_count;
    return QCK_OK;
}
