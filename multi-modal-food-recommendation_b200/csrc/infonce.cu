// SimCLR NT-Xent ("InfoNCE") over the two halves of a [2b, d] batch, forward + backward, sm_100a.
//
// Reference: PRICAI_ModelX.CL_loss (FoodRec/models/pricai_modelx.py:354-378; dormant there -- its call is
// commented out at :259 -- but part of the model's API).  The reference builds four b x b logit blocks
// (aa, bb, ab, ba), subtracts 1e9 on the diagonals of aa / bb, concatenates and calls cross_entropy twice.
// With H = row-normalised hidden and G = H H^T / tau (one symmetric 2b x 2b Gram matrix) this is
//     loss = sum_r [ logsumexp_{c != r} G[r, c] - G[r, pair(r)] ] / b^2,      pair(r) = r +- b
// (the -1e9 entries vanish exactly in fp32), and
//     dL/dH = (W + W^T) H / tau,  W[r, c] = (softmax_r[c] [c != r] - [c = pair(r)]) / b^2,
// followed by the backward of the row normalisation.  Three launches forward (normalise, Gram tiles +
// per-tile online max/sum, fold + loss), one backward.  n = 2b <= a few thousand: latency bound.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kT = 64;

__global__ void nce_normalise_kernel(const float *__restrict__ h, int n, int d, int normalise, float *__restrict__ hn,
                                     float *__restrict__ norm) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) { const float v = h[(size_t)r * d + k]; s = fmaf(v, v, s); }
    s = fr::warp_sum(s);
    const float nr = normalise ? fmaxf(sqrtf(s), 1e-12f) : 1.f;   // F.normalize(p=2, eps=1e-12)
    for (int k = lane; k < d; k += 32) hn[(size_t)r * d + k] = h[(size_t)r * d + k] / nr;
    if (lane == 0) norm[r] = nr;
}

// 64 x 64 tile of G; per (row, column tile): max and sum of exp over the tile's off-diagonal entries.
template <int D>
__global__ void __launch_bounds__(kThreads)
nce_gram_kernel(const float *__restrict__ hn, int n, float inv_tau, float *__restrict__ G, float *__restrict__ pmax,
                float *__restrict__ psum, int n_tiles) {
    __shared__ __align__(16) float Xi[D][kT + 4];
    __shared__ __align__(16) float Xj[D][kT + 4];
    const int i0 = blockIdx.y * kT, j0 = blockIdx.x * kT;
    for (int t = threadIdx.x; t < kT * D; t += kThreads) {
        const int r = t / D, k = t - r * D;
        Xi[k][r] = (i0 + r < n) ? hn[(size_t)(i0 + r) * D + k] : 0.f;
        Xj[k][r] = (j0 + r < n) ? hn[(size_t)(j0 + r) * D + k] : 0.f;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
    for (int k = 0; k < D; ++k) {
        const float4 av = *reinterpret_cast<const float4 *>(&Xi[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4 *>(&Xj[k][tx * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        float m = -INFINITY;
        float g4[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            g4[b] = acc[a][b] * inv_tau;
            if (i < n && j < n) {
                G[(size_t)i * n + j] = g4[b];
                if (j != i) m = fmaxf(m, g4[b]);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            if (i < n && j < n && j != i) s += expf(g4[b] - m);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tx == 0 && i < n) {
            pmax[(size_t)blockIdx.x * n + i] = m;
            psum[(size_t)blockIdx.x * n + i] = s;
        }
    }
}

// lse[r] from the per-tile (max, sum) pairs, then loss = sum_r (lse[r] - G[r, pair(r)]) / b^2 (one block,
// fixed order: deterministic).
__global__ void __launch_bounds__(kThreads)
nce_loss_kernel(const float *__restrict__ G, const float *__restrict__ pmax, const float *__restrict__ psum, int n, int b,
                int n_tiles, float *__restrict__ lse, float *__restrict__ out) {
    __shared__ float red[kThreads / 32];
    float part = 0.f;
    for (int r = threadIdx.x; r < 2 * b; r += kThreads) {
        float m = -INFINITY;
        for (int t = 0; t < n_tiles; ++t) m = fmaxf(m, pmax[(size_t)t * n + r]);
        float s = 0.f;
        for (int t = 0; t < n_tiles; ++t) {
            const float pm = pmax[(size_t)t * n + r];
            if (pm > -INFINITY) s += psum[(size_t)t * n + r] * expf(pm - m);
        }
        const float l = m + logf(s);
        lse[r] = l;
        const int pr = r < b ? r + b : r - b;
        part += l - G[(size_t)r * n + pr];
    }
    part = fr::warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kThreads / 32; ++i) t += red[i];
        out[0] = t / ((float)b * (float)b);
    }
}

// dH_r = sum_c Wsym[r, c] Hn_c / tau, Wsym[r, c] = (e^{G - lse_r} + e^{G - lse_c}) [c != r] / b^2 - 2 [c = pair(r)] / b^2,
// then d h_r = g * (dHn_r - Hn_r (Hn_r . dHn_r)) / norm_r.
template <int D>
__global__ void __launch_bounds__(kThreads)
nce_bwd_kernel(const float *__restrict__ hn, const float *__restrict__ norm, const float *__restrict__ G,
               const float *__restrict__ lse, int n, int b, float inv_tau, int normalise, const float *__restrict__ g_out,
               float *__restrict__ dh) {
    constexpr int RB = 16, JT = 64, KQ = D / 4;
    static_assert(RB * KQ <= kThreads, "tile/threads mismatch");
    __shared__ float Ws[RB][JT + 1];
    __shared__ __align__(16) float Xs[JT][D];
    const int i0 = blockIdx.x * RB;
    const int orow = threadIdx.x / KQ, okq = threadIdx.x % KQ;
    const bool owner = threadIdx.x < RB * KQ;
    const float inv_b2 = 1.f / ((float)b * (float)b);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < n; j0 += JT) {
        for (int t = threadIdx.x; t < JT * KQ; t += kThreads) {
            const int r = t / KQ, q = t - r * KQ;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + r < n) x = fr::ldg_f4(hn + (size_t)(j0 + r) * D + 4 * q);
            *reinterpret_cast<float4 *>(&Xs[r][4 * q]) = x;
        }
        for (int t = threadIdx.x; t < RB * JT; t += kThreads) {
            const int r = t / JT, c = t - r * JT;
            const int i = i0 + r, j = j0 + c;
            float w = 0.f;
            if (i < n && j < n && i != j) {
                const float gij = G[(size_t)i * n + j];
                w = (expf(gij - lse[i]) + expf(gij - lse[j])) * inv_b2;
                const int pr = i < b ? i + b : i - b;
                if (j == pr) w -= 2.f * inv_b2;
            }
            Ws[r][c] = w;
        }
        __syncthreads();
        if (owner) {
#pragma unroll 8
            for (int c = 0; c < JT; ++c) fr::fma4(acc, Ws[orow][c], *reinterpret_cast<const float4 *>(&Xs[c][4 * okq]));
        }
        __syncthreads();
    }
    // normalisation backward needs Hn_r . dHn_r over the KQ threads of a row (consecutive lanes of one warp)
    const int i = i0 + orow;
    const bool live = owner && i < n;
    const float g = __ldg(g_out);
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) hv = fr::ldg_f4(hn + (size_t)i * D + 4 * okq);
    acc.x *= inv_tau; acc.y *= inv_tau; acc.z *= inv_tau; acc.w *= inv_tau;
    float dot = hv.x * acc.x + hv.y * acc.y + hv.z * acc.z + hv.w * acc.w;
#pragma unroll
    for (int o = KQ / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (live) {
        float4 o4;
        if (normalise) {
            const float inv_n = 1.f / norm[i];
            o4 = make_float4(g * (acc.x - hv.x * dot) * inv_n, g * (acc.y - hv.y * dot) * inv_n,
                             g * (acc.z - hv.z * dot) * inv_n, g * (acc.w - hv.w * dot) * inv_n);
        } else {
            o4 = make_float4(g * acc.x, g * acc.y, g * acc.z, g * acc.w);
        }
        *reinterpret_cast<float4 *>(dh + (size_t)i * D + 4 * okq) = o4;
    }
}

}  // namespace

extern "C" int64_t fr_infonce_ws_floats(int32_t n) { return 2 * (int64_t)((n + kT - 1) / kT) * n; }

extern "C" int fr_infonce_fwd(const float *hidden, int32_t b, int32_t d, float temperature, int32_t normalise, float *hn,
                              float *norm, float *G, float *lse, float *out, float *ws, void *stream) {
    const int n = 2 * b;
    FR_REQUIRE(hidden && hn && norm && G && lse && out && ws && b > 0 && temperature > 0.f, "fr_infonce_fwd: bad argument");
    FR_REQUIRE(d == 32 || d == 64, "fr_infonce_fwd: d=%d unsupported (32, 64)", d);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (n + kT - 1) / kT;
    float *pmax = ws, *psum = ws + (size_t)n_tiles * n;
    {
        fr::LaunchTimer _lt("nce_normalise_kernel", st);
        nce_normalise_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(hidden, n, d, normalise, hn, norm);
    }
    if (int rc = fr::check_launch("fr_infonce_fwd/normalise")) return rc;
    {
        fr::LaunchTimer _lt("nce_gram_kernel", st);
        dim3 grid(n_tiles, n_tiles);
        if (d == 64) nce_gram_kernel<64><<<grid, kThreads, 0, st>>>(hn, n, 1.f / temperature, G, pmax, psum, n_tiles);
        else nce_gram_kernel<32><<<grid, kThreads, 0, st>>>(hn, n, 1.f / temperature, G, pmax, psum, n_tiles);
    }
    if (int rc = fr::check_launch("fr_infonce_fwd/gram")) return rc;
    fr::LaunchTimer _lt("nce_loss_kernel", st);
    nce_loss_kernel<<<1, kThreads, 0, st>>>(G, pmax, psum, n, b, n_tiles, lse, out);
    return fr::check_launch("fr_infonce_fwd/loss");
}

extern "C" int fr_infonce_bwd(const float *hn, const float *norm, const float *G, const float *lse, int32_t b, int32_t d,
                              float temperature, int32_t normalise, const float *g_out, float *d_hidden, void *stream) {
    const int n = 2 * b;
    FR_REQUIRE(hn && norm && G && lse && g_out && d_hidden && b > 0 && temperature > 0.f, "fr_infonce_bwd: bad argument");
    FR_REQUIRE(d == 32 || d == 64, "fr_infonce_bwd: d=%d unsupported (32, 64)", d);
    cudaStream_t st = (cudaStream_t)stream;
    fr::LaunchTimer _lt("nce_bwd_kernel", st);
    if (d == 64) nce_bwd_kernel<64><<<(n + 15) / 16, kThreads, 0, st>>>(hn, norm, G, lse, n, b, 1.f / temperature, normalise, g_out, d_hidden);
    else nce_bwd_kernel<32><<<(n + 15) / 16, kThreads, 0, st>>>(hn, norm, G, lse, n, b, 1.f / temperature, normalise, g_out, d_hidden);
    return fr::check_launch("fr_infonce_bwd");
}
