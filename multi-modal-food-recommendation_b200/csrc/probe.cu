// Measurement probe: the ceiling of a row gather on this GPU.
//
// The propagation kernel is bound by how fast 256-byte embedding rows can be gathered (from L2 when the
// tables fit its 126 MB, from HBM otherwise), not by the algorithmic "every byte once" HBM traffic.  This kernel
// does nothing but that gather -- an 8-lane group per index, two 128-bit loads per lane per row, `inflight` rows
// in flight per group, a register sum so the loads cannot be dropped -- so its GB/s is the roofline any gather-
// based SpMM can be held against (`scripts/microbench_gather_roofline.py`).  Not on the product path.
#include "common.cuh"

namespace {

template <int U>
__global__ void __launch_bounds__(256) probe_gather_kernel(const float *__restrict__ tab, const int *__restrict__ idx,
                                                           long long n_idx, float *__restrict__ out) {
    constexpr int D = 64, LPR = 8;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const long long n_groups = (long long)gridDim.x * blockDim.x / LPR;
    const int lg = threadIdx.x % LPR;
    const float *base = tab + lg * 4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const long long per = (n_idx + n_groups - 1) / n_groups;
    const long long lo = group * per, hi = min(n_idx, lo + per);
    for (long long i = lo; i < hi; i += U) {
        float4 x0[U], x1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = __ldg(idx + min(i + u, hi - 1));
            x0[u] = fr::ldg_f4(base + (size_t)r * D);
            x1[u] = fr::ldg_f4(base + (size_t)r * D + 32);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            fr::add4(a0, x0[u]);
            fr::add4(a1, x1[u]);
        }
    }
    fr::add4(a0, a1);
    if (a0.x + a0.y + a0.z + a0.w == 123.456f) out[group] = a0.x;   // keeps the loads alive, practically never taken
}

}  // namespace

extern "C" int fr_probe_gather(const float *tab, int32_t d, const int32_t *idx, int64_t n_idx, int32_t inflight,
                               int32_t blocks, float *out, void *stream) {
    FR_REQUIRE(d == 64 && tab && idx && out && n_idx > 0 && blocks > 0, "fr_probe_gather: d = 64 tables only");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (inflight >= 8)
        probe_gather_kernel<8><<<blocks, 256, 0, st>>>(tab, idx, n_idx, out);
    else
        probe_gather_kernel<4><<<blocks, 256, 0, st>>>(tab, idx, n_idx, out);
    return fr::check_launch("fr_probe_gather");
}
