// SCHGN's per-pair scorer for full-sort evaluation, fused (sm_100a).
//
// Reference: FoodRec/models/schgn.py:159-206 (ingredient-level and component-level attention),
// :233-268 (compute_score), :318-345 (full_sort_predict: one user against every item, with python loops,
// an [I, Dv] upload and a full GCN per user).  Everything that does not depend on the user is computed
// once by the host side (models/schgn.py `_item_side`): final ingredient / item / image / health rows and
// their images under the attention weights.  What is left per (user, item) pair is
//
//   a_j    = h_i . tanh(ingre_key[code_j] + img_key[item] + user_key[u])          j < n_item
//   A      = softmax_j(a)                 att = sum_j A_j ingre_final[code_j]
//   l_1    = h_c . tanh(user_comp[u] + sum_j A_j ingre_comp[code_j])              (linear in att)
//   l_k    = h_c . tanh(user_comp[u] + comp_key_k[item])                          k = item, image, health
//   B      = softmax over the reference's `.view(b, -1)` grouping of the [4, I] logits (schgn.py:198)
//   x      = B_0 item + B_1 att + B_2 image + B_3 health
//   score  = w_out . relu(user_hidden[u] + W_item x + W_prod (u_final * x))
//
// Two launches per block of users because B mixes logits of different items:
//   schgn_attend_kernel: one warp per item, lanes over the 64 features, 16 users per pass so every
//     gathered ingredient row is reused 16 times from registers; writes att [nu, I, 64] and logits [nu, 4, I].
//   schgn_score_kernel: one thread per (user, 2 items); the user's 64x64 matrix
//     M_u = W_item + W_prod diag(u_final) sits in shared memory (broadcast reads), x in registers.
// Arithmetic is fp32 throughout (scores feed a top-K that must match the reference's).
#include "common.cuh"

namespace {

constexpr int D = 64;           // embedding width of the model (schgn.py:69-70 hard-codes 64)
constexpr int MAX_SLOTS = 32;   // ingredient slots per recipe (reference: 20)
constexpr int UB = 16;          // users per attend pass
constexpr int ATT_WARPS = 8;

struct AttendParams {
    const float *user_key, *user_comp;
    const int32_t *codes, *nums;
    const float *ingre_key, *ingre_final, *ingre_comp, *img_key, *comp_keys, *h_ingre, *h_comp;
    float *att, *logits;
    int32_t nu, n_items, slots;
};

template <bool FAST>
__device__ __forceinline__ float tanh_f(float x) {
    if (FAST) return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
    return tanhf(x);
}

// Sum each of v[0..15] over the 32 lanes with 16 shuffles; lane L ends up holding the total of
// v[L >> 1] (both lanes of a pair hold the same value).
__device__ __forceinline__ float reduce16(float (&v)[UB], int lane) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = lane & bit;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <bool FAST>
__global__ void __launch_bounds__(ATT_WARPS * 32) schgn_attend_kernel(const AttendParams p) {
    __shared__ float s_ukey[UB][D];
    __shared__ float s_ucomp[UB][D];
    __shared__ float s_logit[ATT_WARPS][UB][MAX_SLOTS];
    const int u0 = blockIdx.y * UB;
    for (int t = threadIdx.x; t < UB * D; t += blockDim.x) {
        const int u = min(u0 + t / D, p.nu - 1);
        s_ukey[t / D][t % D] = p.user_key[(size_t)u * D + t % D];
        s_ucomp[t / D][t % D] = p.user_comp[(size_t)u * D + t % D];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * ATT_WARPS + warp;
    if (item >= p.n_items) return;
    const int n = max(0, min(p.nums[item], p.slots));
    const int code = lane < p.slots ? p.codes[(size_t)item * p.slots + lane] : 0;
    const int c2 = 2 * lane;
    const float2 pi = *reinterpret_cast<const float2 *>(p.img_key + (size_t)item * D + c2);
    const float2 hi = *reinterpret_cast<const float2 *>(p.h_ingre + c2);
    const float2 hc = *reinterpret_cast<const float2 *>(p.h_comp + c2);

    // pass 1: attention logits a[u][j]
    for (int j = 0; j < n; ++j) {
        const int c = __shfl_sync(0xffffffffu, code, j);
        float2 pe = __ldg(reinterpret_cast<const float2 *>(p.ingre_key + (size_t)c * D + c2));
        pe.x += pi.x;
        pe.y += pi.y;
        float v[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const float2 uk = *reinterpret_cast<const float2 *>(&s_ukey[u][c2]);
            v[u] = hi.x * tanh_f<FAST>(pe.x + uk.x) + hi.y * tanh_f<FAST>(pe.y + uk.y);
        }
        const float tot = reduce16(v, lane);
        if (!(lane & 1)) s_logit[warp][lane >> 1][j] = tot;
    }
    __syncwarp();

    // softmax over the recipe's real ingredients (masked slots carry weight exactly 0 in the reference:
    // exp(-1e12 - max) underflows); lane j keeps A[u][j]
    float A[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        const float x = lane < n ? s_logit[warp][u][lane] : -INFINITY;
        float m = x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float e = lane < n ? expf(x - m) : 0.0f;
        const float s = fr::warp_sum(e);
        A[u] = n > 0 ? e / s : 0.0f;
    }

    // pass 2: attended ingredient row and its image under W_att_comp's component half
    float2 att[UB], q[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) att[u] = q[u] = make_float2(0.0f, 0.0f);
    for (int j = 0; j < n; ++j) {
        const int c = __shfl_sync(0xffffffffu, code, j);
        const float2 ef = __ldg(reinterpret_cast<const float2 *>(p.ingre_final + (size_t)c * D + c2));
        const float2 qe = __ldg(reinterpret_cast<const float2 *>(p.ingre_comp + (size_t)c * D + c2));
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const float w = __shfl_sync(0xffffffffu, A[u], j);
            att[u].x = fmaf(w, ef.x, att[u].x);
            att[u].y = fmaf(w, ef.y, att[u].y);
            q[u].x = fmaf(w, qe.x, q[u].x);
            q[u].y = fmaf(w, qe.y, q[u].y);
        }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u)
        if (u0 + u < p.nu)
            *reinterpret_cast<float2 *>(p.att + ((size_t)(u0 + u) * p.n_items + item) * D + c2) = att[u];

    // component logits, reference order: item id, attended ingredients, image, health
    const float *ck = p.comp_keys + (size_t)item * 3 * D + c2;
    const float2 k_item = *reinterpret_cast<const float2 *>(ck);
    const float2 k_img = *reinterpret_cast<const float2 *>(ck + D);
    const float2 k_hl = *reinterpret_cast<const float2 *>(ck + 2 * D);
    const int my_u = u0 + (lane >> 1);
#pragma unroll
    for (int comp = 0; comp < 4; ++comp) {
        float v[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const float2 uc = *reinterpret_cast<const float2 *>(&s_ucomp[u][c2]);
            const float2 k = comp == 0 ? k_item : comp == 1 ? q[u] : comp == 2 ? k_img : k_hl;
            v[u] = hc.x * tanh_f<FAST>(uc.x + k.x) + hc.y * tanh_f<FAST>(uc.y + k.y);
        }
        const float tot = reduce16(v, lane);
        if (!(lane & 1) && my_u < p.nu) p.logits[((size_t)my_u * 4 + comp) * p.n_items + item] = tot;
    }
}

struct ScoreParams {
    const float *user_final, *user_hidden, *W_item, *W_prod, *w_out, *comps, *att, *logits;
    float *scores;
    int32_t nu, n_items;
};

constexpr int SC_THREADS = 128;
constexpr int IPT = 2;  // items per thread

__global__ void __launch_bounds__(SC_THREADS) schgn_score_kernel(const ScoreParams p) {
    __shared__ __align__(16) float sM[D][D];
    __shared__ float sH[D], sW[D];
    const int u = blockIdx.y;
    for (int t = threadIdx.x; t < D * D; t += SC_THREADS)
        sM[t / D][t % D] = fmaf(p.W_prod[t], p.user_final[(size_t)u * D + t % D], p.W_item[t]);
    if (threadIdx.x < D) {
        sH[threadIdx.x] = p.user_hidden[(size_t)u * D + threadIdx.x];
        sW[threadIdx.x] = p.w_out[threadIdx.x];
    }
    __syncthreads();
    const int base = (blockIdx.x * SC_THREADS + threadIdx.x) * IPT;
    if (base >= p.n_items) return;
    const float *lg = p.logits + (size_t)u * 4 * p.n_items;  // [4, I] read as [I, 4]: schgn.py:198
    float x[IPT][D];
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
        const int r = min(base + t, p.n_items - 1);
        const float4 l = *reinterpret_cast<const float4 *>(lg + (size_t)4 * r);
        const float m = fmaxf(fmaxf(l.x, l.y), fmaxf(l.z, l.w));
        const float e0 = expf(l.x - m), e1 = expf(l.y - m), e2 = expf(l.z - m), e3 = expf(l.w - m);
        const float inv = 1.0f / (e0 + e1 + e2 + e3);
        const float b0 = e0 * inv, b1 = e1 * inv, b2 = e2 * inv, b3 = e3 * inv;
        const float *cr = p.comps + (size_t)r * 3 * D;
        const float *ar = p.att + ((size_t)u * p.n_items + r) * D;
#pragma unroll
        for (int d = 0; d < D; d += 4) {
            const float4 ci = fr::ldg_f4(cr + d), cm = fr::ldg_f4(cr + D + d), ch = fr::ldg_f4(cr + 2 * D + d);
            const float4 ca = fr::ldg_f4(ar + d);
            x[t][d + 0] = b0 * ci.x + b1 * ca.x + b2 * cm.x + b3 * ch.x;
            x[t][d + 1] = b0 * ci.y + b1 * ca.y + b2 * cm.y + b3 * ch.y;
            x[t][d + 2] = b0 * ci.z + b1 * ca.z + b2 * cm.z + b3 * ch.z;
            x[t][d + 3] = b0 * ci.w + b1 * ca.w + b2 * cm.w + b3 * ch.w;
        }
    }
    float score[IPT];
#pragma unroll
    for (int t = 0; t < IPT; ++t) score[t] = 0.0f;
#pragma unroll 2
    for (int e = 0; e < D; ++e) {
        float h[IPT][4];
#pragma unroll
        for (int t = 0; t < IPT; ++t) h[t][0] = sH[e], h[t][1] = h[t][2] = h[t][3] = 0.0f;
#pragma unroll
        for (int d = 0; d < D; d += 4) {
            const float4 m = *reinterpret_cast<const float4 *>(&sM[e][d]);
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                h[t][0] = fmaf(m.x, x[t][d + 0], h[t][0]);
                h[t][1] = fmaf(m.y, x[t][d + 1], h[t][1]);
                h[t][2] = fmaf(m.z, x[t][d + 2], h[t][2]);
                h[t][3] = fmaf(m.w, x[t][d + 3], h[t][3]);
            }
        }
        const float w = sW[e];
#pragma unroll
        for (int t = 0; t < IPT; ++t)
            score[t] = fmaf(w, fmaxf((h[t][0] + h[t][1]) + (h[t][2] + h[t][3]), 0.0f), score[t]);
    }
#pragma unroll
    for (int t = 0; t < IPT; ++t)
        if (base + t < p.n_items) p.scores[(size_t)u * p.n_items + base + t] = score[t];
}

}  // namespace

extern "C" int fr_schgn_attend(const float *user_key, const float *user_comp, int32_t nu, const int32_t *codes,
                               int32_t slots, const int32_t *nums, int32_t n_items, const float *ingre_key,
                               const float *ingre_final, const float *ingre_comp, const float *img_key,
                               const float *comp_keys, const float *h_ingre, const float *h_comp, int32_t d,
                               int32_t fast_tanh, float *att, float *logits, void *stream) {
    FR_REQUIRE(d == D, "fr_schgn_attend: embedding width %d unsupported (the model fixes 64)", d);
    FR_REQUIRE(nu >= 0 && n_items >= 0 && slots >= 1 && slots <= MAX_SLOTS,
               "fr_schgn_attend: bad extents (nu=%d, n_items=%d, slots=%d; at most %d slots)", nu, n_items, slots,
               MAX_SLOTS);
    if (nu == 0 || n_items == 0) return FR_OK;
    FR_REQUIRE(user_key && user_comp && codes && nums && ingre_key && ingre_final && ingre_comp && img_key &&
                   comp_keys && h_ingre && h_comp && att && logits,
               "fr_schgn_attend: null pointer");
    FR_REQUIRE((nu + UB - 1) / UB <= 65535, "fr_schgn_attend: at most %d users per call", 65535 * UB);
    AttendParams p{user_key, user_comp, codes,   nums, ingre_key, ingre_final, ingre_comp, img_key,
                   comp_keys, h_ingre,  h_comp, att,  logits,    nu,          n_items,    slots};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fr::LaunchTimer timer("schgn_attend", st);
    dim3 grid((n_items + ATT_WARPS - 1) / ATT_WARPS, (nu + UB - 1) / UB);
    if (fast_tanh)
        schgn_attend_kernel<true><<<grid, ATT_WARPS * 32, 0, st>>>(p);
    else
        schgn_attend_kernel<false><<<grid, ATT_WARPS * 32, 0, st>>>(p);
    return fr::check_launch("fr_schgn_attend");
}

extern "C" int fr_schgn_score(const float *user_final, const float *user_hidden, int32_t nu, const float *W_item,
                              const float *W_prod, const float *w_out, const float *comps, const float *att,
                              const float *logits, int32_t n_items, int32_t d, float *scores, void *stream) {
    FR_REQUIRE(d == D, "fr_schgn_score: embedding width %d unsupported (the model fixes 64)", d);
    FR_REQUIRE(nu >= 0 && nu <= 65535 && n_items >= 0, "fr_schgn_score: bad extents (nu=%d, n_items=%d)", nu, n_items);
    if (nu == 0 || n_items == 0) return FR_OK;
    FR_REQUIRE(user_final && user_hidden && W_item && W_prod && w_out && comps && att && logits && scores,
               "fr_schgn_score: null pointer");
    ScoreParams p{user_final, user_hidden, W_item, W_prod, w_out, comps, att, logits, scores, nu, n_items};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fr::LaunchTimer timer("schgn_score", st);
    dim3 grid((n_items + SC_THREADS * IPT - 1) / (SC_THREADS * IPT), nu);
    schgn_score_kernel<<<grid, SC_THREADS, 0, st>>>(p);
    return fr::check_launch("fr_schgn_score");
}
