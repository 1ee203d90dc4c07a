// Scratch microbenchmark (not part of the product): how fast can 32 GiB be written on a B200
// as a function of store flavour and address order?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ void st4(double* p, double a) {
    if (MODE == 0) asm volatile("st.global.cs.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(p), "d"(a) : "memory");
    if (MODE == 1) asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(p), "d"(a) : "memory");
    if (MODE == 2) { asm volatile("st.global.cs.v2.f64 [%0], {%1,%1};" ::"l"(p), "d"(a) : "memory");
                     asm volatile("st.global.cs.v2.f64 [%0], {%1,%1};" ::"l"(p + 2), "d"(a) : "memory"); }
    if (MODE == 3) asm volatile("st.global.wt.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(p), "d"(a) : "memory");
    if (MODE == 4) asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(p), "d"(a) : "memory");
}

// linear: CTA b writes chunks b, b+grid, ...  (chunk = 4096 doubles)
template <int MODE> __global__ void __launch_bounds__(256) k_linear(double* out, unsigned long long n_chunks, double a) {
    for (unsigned long long g = blockIdx.x; g < n_chunks; g += gridDim.x) {
        double* dst = out + (g << 12) + 4 * threadIdx.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) st4<MODE>(dst + 1024 * e, a);
    }
}
// blocked: CTA b writes a contiguous range of chunks
template <int MODE> __global__ void __launch_bounds__(256) k_blocked(double* out, unsigned long long n_chunks, double a) {
    unsigned long long g0 = n_chunks * blockIdx.x / gridDim.x, g1 = n_chunks * (blockIdx.x + 1ull) / gridDim.x;
    for (unsigned long long g = g0; g < g1; ++g) {
        double* dst = out + (g << 12) + 4 * threadIdx.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) st4<MODE>(dst + 1024 * e, a);
    }
}
// syc32 order: chunk number g (blocked per CTA) -> y_hi with B bits fastest
__device__ __forceinline__ unsigned long long syc_addr(unsigned long long g) {
    // order of free hi bits (relative to bit 12): B: 4..11, 13..18 ; A: 0..3, 12, 19
    const int order[20] = {4,5,6,7,8,9,10,11,13,14,15,16,17,18,0,1,2,3,12,19};
    unsigned long long y = 0;
#pragma unroll
    for (int j = 0; j < 20; ++j) y |= ((g >> j) & 1ull) << order[j];
    return y;
}
template <int MODE> __global__ void __launch_bounds__(256) k_syc(double* out, unsigned long long n_chunks, double a) {
    unsigned long long g0 = n_chunks * blockIdx.x / gridDim.x, g1 = n_chunks * (blockIdx.x + 1ull) / gridDim.x;
    for (unsigned long long g = g0; g < g1; ++g) {
        double* dst = out + (syc_addr(g) << 12) + 4 * threadIdx.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) st4<MODE>(dst + 1024 * e, a);
    }
}
// syc32 order, interleaved: CTA b takes chunks b, b+grid, ... of the permuted sequence
template <int MODE> __global__ void __launch_bounds__(256) k_syc_rr(double* out, unsigned long long n_chunks, double a) {
    for (unsigned long long g = blockIdx.x; g < n_chunks; g += gridDim.x) {
        double* dst = out + (syc_addr(g) << 12) + 4 * threadIdx.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) st4<MODE>(dst + 1024 * e, a);
    }
}


// 128-bit fully coalesced: lane i writes 16 B at base + 16 i (512 B per warp instruction)
template <int MODE> __global__ void __launch_bounds__(256) k_lin128(double* out, unsigned long long n_chunks, double a) {
    for (unsigned long long g = blockIdx.x; g < n_chunks; g += gridDim.x) {
        double* dst = out + (g << 12) + 2 * threadIdx.x;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (MODE == 0) asm volatile("st.global.cs.v2.f64 [%0], {%1,%1};" ::"l"(dst + 512 * e), "d"(a) : "memory");
            else asm volatile("st.global.v2.f64 [%0], {%1,%1};" ::"l"(dst + 512 * e), "d"(a) : "memory");
        }
    }
}
// non-persistent: one CTA per chunk (hardware CTA scheduler walks memory linearly)
template <int MODE> __global__ void __launch_bounds__(256) k_np256(double* out, double a) {
    double* dst = out + ((unsigned long long)blockIdx.x << 12) + 4 * threadIdx.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) st4<MODE>(dst + 1024 * e, a);
}
template <int MODE> __global__ void __launch_bounds__(256) k_np128(double* out, double a) {
    double* dst = out + ((unsigned long long)blockIdx.x << 12) + 2 * threadIdx.x;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (MODE == 0) asm volatile("st.global.cs.v2.f64 [%0], {%1,%1};" ::"l"(dst + 512 * e), "d"(a) : "memory");
        else asm volatile("st.global.v2.f64 [%0], {%1,%1};" ::"l"(dst + 512 * e), "d"(a) : "memory");
    }
}
// non-persistent, small CTAs like at::fill (128 threads x 4 x 16 B = 8 KiB per CTA)
__global__ void __launch_bounds__(128) k_np_small(double* out, double a) {
    double* dst = out + ((unsigned long long)blockIdx.x << 10) + 2 * threadIdx.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) asm volatile("st.global.v2.f64 [%0], {%1,%1};" ::"l"(dst + 256 * e), "d"(a) : "memory");
}
// persistent with 1024-thread CTAs
template <int MODE> __global__ void __launch_bounds__(1024) k_lin1024(double* out, unsigned long long n_chunks, double a) {
    for (unsigned long long g = blockIdx.x; g < n_chunks; g += gridDim.x) {
        double* dst = out + (g << 12) + 4 * threadIdx.x;
        st4<MODE>(dst, a);
    }
}

// persistent + dynamic: CTAs fetch batches of chunks from an atomic counter
template <int BATCH> __global__ void __launch_bounds__(256) k_dyn(double* out, unsigned long long n_chunks, double a,
                                                                  unsigned long long* counter) {
    __shared__ unsigned long long s_g;
    while (true) {
        if (threadIdx.x == 0) s_g = atomicAdd(counter, (unsigned long long)BATCH);
        __syncthreads();
        const unsigned long long g0 = s_g;
        __syncthreads();
        if (g0 >= n_chunks) break;
        for (int b = 0; b < BATCH; ++b) {
            double* dst = out + (syc_addr(g0 + b) << 12) + 4 * threadIdx.x;
#pragma unroll
            for (int e = 0; e < 4; ++e) st4<0>(dst + 1024 * e, a);
        }
    }
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main() {
    const unsigned long long n = 1ull << 32, n_chunks = n >> 12;
    double* out; CK(cudaMalloc(&out, n * 8));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("cudaMemset: %.3f ms\n", timeit([&] { cudaMemsetAsync(out, 0, n * 8); }));
    for (int per = 4; per <= 4; per *= 2) {
        int grid = sms * per;
#define RUN(K, M) do { t = timeit([&] { K<M><<<grid, 256>>>(out, n_chunks, 1.5); }); printf(#K " mode %d ctas/sm %d: %.3f ms  %.0f GB/s\n", M, per, t, 8.0 * n / t / 1e6); } while (0)
        float t;
        RUN(k_linear, 0); RUN(k_linear, 1); RUN(k_linear, 2); RUN(k_linear, 3); RUN(k_linear, 4);
        RUN(k_blocked, 0); RUN(k_blocked, 1);
        RUN(k_syc, 0); RUN(k_syc, 1);
        RUN(k_syc_rr, 0); RUN(k_syc_rr, 1);
    }
    { float t;
      t = timeit([&] { k_np256<0><<<(unsigned)n_chunks, 256>>>(out, 1.5); }); printf("k_np256 cs: %.3f ms %.0f GB/s\n", t, 8.0*n/t/1e6);
      t = timeit([&] { k_np256<1><<<(unsigned)n_chunks, 256>>>(out, 1.5); }); printf("k_np256 wb: %.3f ms %.0f GB/s\n", t, 8.0*n/t/1e6);
      t = timeit([&] { k_np128<0><<<(unsigned)n_chunks, 256>>>(out, 1.5); }); printf("k_np128 cs: %.3f ms %.0f GB/s\n", t, 8.0*n/t/1e6);
      t = timeit([&] { k_np128<1><<<(unsigned)n_chunks, 256>>>(out, 1.5); }); printf("k_np128 wb: %.3f ms %.0f GB/s\n", t, 8.0*n/t/1e6);
      t = timeit([&] { k_np_small<<<(unsigned)(n >> 10), 128>>>(out, 1.5); }); printf("k_np_small: %.3f ms %.0f GB/s\n", t, 8.0*n/t/1e6);
      for (int per = 1; per <= 8; per *= 2) {
        t = timeit([&] { k_lin128<0><<<sms*per, 256>>>(out, n_chunks, 1.5); }); printf("k_lin128 cs ctas/sm %d: %.3f ms %.0f GB/s\n", per, t, 8.0*n/t/1e6);
        t = timeit([&] { k_lin128<1><<<sms*per, 256>>>(out, n_chunks, 1.5); }); printf("k_lin128 wb ctas/sm %d: %.3f ms %.0f GB/s\n", per, t, 8.0*n/t/1e6);
      }
      for (int per = 1; per <= 2; per *= 2) {
        t = timeit([&] { k_lin1024<0><<<sms*per, 1024>>>(out, n_chunks, 1.5); }); printf("k_lin1024 cs ctas/sm %d: %.3f ms %.0f GB/s\n", per, t, 8.0*n/t/1e6);
      }
    }
    { float t; unsigned long long* ctr; CK(cudaMalloc(&ctr, 8));
      for (int per = 2; per <= 8; per *= 2) {
        t = timeit([&] { cudaMemsetAsync(ctr, 0, 8); k_dyn<1><<<sms*per, 256>>>(out, n_chunks, 1.5, ctr); }); printf("k_dyn<1> ctas/sm %d: %.3f ms %.0f GB/s\n", per, t, 8.0*n/t/1e6);
        t = timeit([&] { cudaMemsetAsync(ctr, 0, 8); k_dyn<4><<<sms*per, 256>>>(out, n_chunks, 1.5, ctr); }); printf("k_dyn<4> ctas/sm %d: %.3f ms %.0f GB/s\n", per, t, 8.0*n/t/1e6);
        t = timeit([&] { cudaMemsetAsync(ctr, 0, 8); k_dyn<32><<<sms*per, 256>>>(out, n_chunks, 1.5, ctr); }); printf("k_dyn<32> ctas/sm %d: %.3f ms %.0f GB/s\n", per, t, 8.0*n/t/1e6);
        t = timeit([&] { cudaMemsetAsync(ctr, 0, 8); k_dyn<128><<<sms*per, 256>>>(out, n_chunks, 1.5, ctr); }); printf("k_dyn<128> ctas/sm %d: %.3f ms %.0f GB/s\n", per, t, 8.0*n/t/1e6);
      }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
