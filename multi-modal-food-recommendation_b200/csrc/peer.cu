// Peer-mapped tables for the row-partitioned multi-GPU propagation (one process per GPU, one node).
//
// Every rank owns full-size [n_padded, d] fp32 tables allocated here with cudaMalloc and exported as CUDA IPC
// handles; the other ranks of the node map them (NVLink / NVSwitch peer access) and the propagation kernel's
// push epilogue (`fr_spmm_csr_f32_push`, spmm.cu) stores each finished output row straight into all of them.
// `fr_push_rows` is the same exchange for rows that no SpMM produced (the layer-0 input).
// Ordering between ranks is the caller's stream-ordered barrier (a 4-byte NCCL all-reduce); nothing here
// waits on another rank.
#include "common.cuh"

namespace {

struct PushParams {
    float *peer[8];
    int n_peers;
};

__global__ void __launch_bounds__(256) push_rows_kernel(const float4 *__restrict__ src, long long n_vec,
                                                        long long off_vec, PushParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const float4 v = __ldg(src + i);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < p.n_peers) reinterpret_cast<float4 *>(p.peer[q])[off_vec + i] = v;
    }
}

}  // namespace

extern "C" int fr_peer_alloc(int64_t bytes, void **ptr, void *handle64) {
    FR_REQUIRE(bytes > 0 && ptr && handle64, "fr_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), *ptr);
    if (e != cudaSuccess) {
        fr::set_error("fr_peer_alloc(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
        return FR_ECUDA;
    }
    return FR_OK;
}

extern "C" int fr_peer_open(const void *handle64, void **ptr) {
    FR_REQUIRE(handle64 && ptr, "fr_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        fr::set_error("fr_peer_open: %s", cudaGetErrorString(e));
        return FR_ECUDA;
    }
    return FR_OK;
}

extern "C" int fr_peer_close(void *ptr) {
    cudaError_t e = ptr ? cudaIpcCloseMemHandle(ptr) : cudaSuccess;
    if (e != cudaSuccess) {
        fr::set_error("fr_peer_close: %s", cudaGetErrorString(e));
        return FR_ECUDA;
    }
    return FR_OK;
}

extern "C" int fr_peer_free(void *ptr) {
    cudaError_t e = ptr ? cudaFree(ptr) : cudaSuccess;
    if (e != cudaSuccess) {
        fr::set_error("fr_peer_free: %s", cudaGetErrorString(e));
        return FR_ECUDA;
    }
    return FR_OK;
}

extern "C" int fr_push_rows(const float *src, int64_t rows, int32_t d, float *const *peers_host, int32_t n_peers,
                            int64_t row_off, void *stream) {
    FR_REQUIRE(rows >= 0 && d > 0 && d % 4 == 0 && n_peers >= 1 && n_peers <= 8 && peers_host && row_off >= 0,
               "fr_push_rows: bad arguments (d must be a multiple of 4, 1..8 peer tables)");
    if (rows == 0) return FR_OK;
    FR_REQUIRE(src && ((uintptr_t)src & 15) == 0, "fr_push_rows: src must be 16-byte aligned");
    PushParams p;
    p.n_peers = n_peers;
    for (int q = 0; q < 8; ++q) {
        p.peer[q] = q < n_peers ? peers_host[q] : nullptr;
        FR_REQUIRE(q >= n_peers || (p.peer[q] && ((uintptr_t)p.peer[q] & 15) == 0), "fr_push_rows: bad peer table");
    }
    const long long n_vec = rows * (d / 4), off_vec = row_off * (d / 4);
    const int blocks = (int)std::min<long long>((n_vec + 255) / 256, (long long)fr::num_sms() * 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fr::LaunchTimer timer("push_rows", st);
    push_rows_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(src), n_vec, off_vec, p);
    return fr::check_launch("fr_push_rows");
}
