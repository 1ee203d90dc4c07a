// Handle lifetime, error strings and scratch management of libqck.so.
#include "qck_common.cuh"

#include <new>

int qck_sim_init(qck_handle* h);   // sim.cu
int qck_knit_init(qck_handle* h);  // knit.cu
int qck_npd_init(qck_handle* h);   // npd.cu

extern "C" int qck_abi_version(void) { return QCK_ABI_VERSION; }

extern "C" const char* qck_status_string(int status) {
    switch (status) {
        case QCK_OK: return "ok";
        case QCK_ERR_INVALID_ARG: return "invalid argument";
        case QCK_ERR_CUDA: return "CUDA error";
        case QCK_ERR_UNSUPPORTED: return "unsupported";
        case QCK_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

extern "C" int qck_create(int device, qck_handle** out) {
    if (!out) return QCK_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return QCK_ERR_CUDA;
    if (device < 0 || device >= count) return QCK_ERR_INVALID_ARG;
    qck_handle* full = new (std::nothrow) qck_handle;
    if (!full) return QCK_ERR_NOMEM;
    memset(full, 0, sizeof(*full));
    qck_handle* h = full;
    h->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete full;
        return QCK_ERR_CUDA;
    }
    if (prop.major < 10) {  // sm_100a only: no fallback code path exists
        delete full;
        return QCK_ERR_UNSUPPORTED;
    }
    h->sm_count = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->partials_count = 1 << 16;
    if (cudaMalloc(&h->d_partials, h->partials_count * sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&h->h_pinned, 64 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->knit_ctr, (1024 + 8) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(h->knit_ctr, 0, (1024 + 8) * sizeof(unsigned long long)) != cudaSuccess) {
        if (h->d_partials) cudaFree(h->d_partials);
        if (h->h_pinned) cudaFreeHost(h->h_pinned);
        if (h->knit_ctr) cudaFree(h->knit_ctr);
        delete full;
        return QCK_ERR_NOMEM;
    }
    if (qck_sim_init(h) != QCK_OK || qck_knit_init(h) != QCK_OK || qck_npd_init(h) != QCK_OK) {
        if (h->npd_ws) cudaFree(h->npd_ws);
        cudaFree(h->d_partials);
        cudaFreeHost(h->h_pinned);
        delete full;
        return QCK_ERR_CUDA;
    }
    *out = h;
    return QCK_OK;
}

extern "C" int qck_destroy(qck_handle* h) {
    if (!h) return QCK_OK;
    DeviceGuard guard(h->device);
    if (h->d_partials) cudaFree(h->d_partials);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->scratch) cudaFree(h->scratch);
    if (h->npd_ws) cudaFree(h->npd_ws);
    if (h->knit_ctr) cudaFree(h->knit_ctr);
    for (int i = 0; i < QCK_SIDE_STREAMS; ++i)
        if (h->warp_stash[i]) cudaFree(h->warp_stash[i]);
    if (h->side_ready) {
        for (int i = 0; i < QCK_SIDE_STREAMS; ++i) {
            cudaStreamDestroy(h->side[i]);
            cudaEventDestroy(h->side_done[i]);
        }
        cudaEventDestroy(h->fork);
    }
    delete h;
    return QCK_OK;
}

extern "C" const char* qck_last_error_string(const qck_handle* h) { return h ? h->err : "null handle"; }

extern "C" int64_t qck_launch_count(const qck_handle* h) { return h ? h->launches : 0; }

int qck_ensure_partials(qck_handle* h, size_t count) {
    if (count + 8 <= h->partials_count) return QCK_OK;
    // partials are consumed by the kernel enqueued right after they are written, on the same
    // stream; growing here would race with in-flight work, so the size is fixed at create time
    QCK_FAIL(h, QCK_ERR_UNSUPPORTED, "reduction scratch too small (%zu > %zu)", count, h->partials_count);
}

int qck_ensure_scratch(qck_handle* h, size_t bytes, void** out) {
    if (bytes > h->scratch_bytes) {
        // synchronous by design: cudaFree waits for in-flight users of the old block
        if (h->scratch) {
            cudaError_t e = cudaFree(h->scratch);
            h->scratch = nullptr;
            h->scratch_bytes = 0;
            if (e != cudaSuccess) QCK_FAIL(h, QCK_ERR_CUDA, "cudaFree(scratch): %s", cudaGetErrorString(e));
        }
        size_t want = bytes + (bytes >> 2);
        cudaError_t e = cudaMalloc(&h->scratch, want);
        if (e != cudaSuccess) {
            h->scratch = nullptr;
            QCK_FAIL(h, QCK_ERR_NOMEM, "cudaMalloc(%zu bytes of scratch): %s", want, cudaGetErrorString(e));
        }
        h->scratch_bytes = want;
    }
    *out = h->scratch;
    return QCK_OK;
}
