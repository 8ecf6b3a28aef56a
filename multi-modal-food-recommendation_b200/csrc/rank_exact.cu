// Exactness layer of the ranking path: fp32 re-scoring of the bf16 candidates with a per-row CERTIFICATE, and an
// exact fp32 row kernel for the rows the certificate rejects.
//
// The tensor-core pass ranks the bf16-rounded operands a' = bf16(a), b' = bf16(b).  Exactly,
//     a'.b' - a.b = (a' - a).b' + a.(b' - b),
// so the bf16 score s~ of any column differs from its fp32 score s by at most
//     E_row = |scale| ( |a' - a| max_b |b'|  +  |a| max_b |b' - b|  +  K 2^-23 |a| max_b |b'| )
// (Cauchy-Schwarz on both products; the last term covers the fp32 accumulation of K terms).  |a' - a| and |a| are
// computed per row from the fp32 row itself, the two column maxima once per table (`fr_max_row_norm`) -- about
// 2^-9 |a| max|b|, 2.5x tighter than the worst-case relative bound 2^-8 |a||b|.  The bf16 pass keeps the kc best
// columns; every column it dropped has s~ <= s~_kc (the smallest kept bf16 score), hence s <= s~_kc + E_row.  If that
// is below the k-th best fp32 re-score v_k, no dropped column can belong to the fp32 top-k:
//     certified  <=>  fewer than kc eligible columns exist   or   s~_kc + E_row < v_k
// Rows that fail are re-ranked with a wider candidate set and, if they still fail, by `fr_exact_topk_f32`, which
// scores the row against every column in fp32 on the CUDA cores and selects by radix select.  The result is therefore
// the fp32 top-k (ties: lower column first) for every row, not "unless rounding displaced an item by > slack ranks".
#include <algorithm>

#include "rank_common.cuh"

namespace {
using rk::order_key;

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }   // as fr_f32_to_bf16

// adds |bf16(x)|^2 to n2 and |bf16(x) - x|^2 to e2 for the four components
__device__ __forceinline__ void bf16_norms(const float4 x, float &n2, float &e2) {
    const float r0 = bf16_round(x.x), r1 = bf16_round(x.y), r2 = bf16_round(x.z), r3 = bf16_round(x.w);
    n2 = fmaf(r0, r0, n2); n2 = fmaf(r1, r1, n2); n2 = fmaf(r2, r2, n2); n2 = fmaf(r3, r3, n2);
    const float d0 = r0 - x.x, d1 = r1 - x.y, d2 = r2 - x.z, d3 = r3 - x.w;
    e2 = fmaf(d0, d0, e2); e2 = fmaf(d1, d1, e2); e2 = fmaf(d2, d2, e2); e2 = fmaf(d3, d3, e2);
}

// out[0] = max_r |bf16(B_r)|, out[1] = max_r |bf16(B_r) - B_r|   (non-negative floats order like their bit patterns)
__global__ void max_row_norm_kernel(const float *__restrict__ B, long long n, int d, unsigned *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    float best_n = 0.f, best_e = 0.f;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += ((long long)gridDim.x * blockDim.x) >> 5) {
        float n2 = 0.f, e2 = 0.f;
        for (int q = lane * 4; q < d; q += 128) bf16_norms(fr::ldg_f4(B + r * d + q), n2, e2);
        best_n = fmaxf(best_n, fr::warp_sum(n2));
        best_e = fmaxf(best_e, fr::warp_sum(e2));
    }
    if (lane == 0) {
        atomicMax(out, __float_as_uint(sqrtf(best_n) * 1.000001f));
        atomicMax(out + 1, __float_as_uint(sqrtf(best_e) * 1.000001f));
    }
}

// One warp per row.  cand: [M, kc] columns from the bf16 pass (descending bf16 score, -1 padding), cand_val their
// bf16-pass scores.  Writes the k best by exact fp32 score and (optionally) the row's certificate.
template <typename IdxT>
__global__ void __launch_bounds__(256)
rescore_topk_kernel(const float *__restrict__ A, const int64_t *__restrict__ a_rows, const float *__restrict__ B,
                    int d, float scale, const float *__restrict__ bias, int metric,
                    const int32_t *__restrict__ cand, const float *__restrict__ cand_val, int kc, int M, int k,
                    float *__restrict__ out_val, IdxT *__restrict__ out_idx, const float *__restrict__ bmax,
                    uint8_t *__restrict__ cert) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= M) return;
    const float *a = A + (size_t)(a_rows ? a_rows[row] : row) * d;
    float v[2];
    int ci[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int c = lane + 32 * t;
        ci[t] = c < kc ? cand[(size_t)row * kc + c] : -1;
        v[t] = -INFINITY;
    }
    const int last_cand = cand[(size_t)row * kc + kc - 1];        // -1: the bf16 pass ran out of eligible columns
    if (d >= 256) {
        // wide rows (kNN / centroid features): the whole warp walks one candidate row at a time with coalesced
        // 128-bit loads and folds by shuffle; lane (c % 32) keeps candidate c's score
        for (int c = 0; c < kc; ++c) {
            const int cc = __shfl_sync(0xffffffffu, ci[c >> 5], c & 31);
            if (cc < 0) continue;
            const float *b = B + (size_t)cc * d;
            float s = 0.f;
            for (int q = lane * 4; q < d; q += 128) {
                const float4 x = fr::ldg_f4(a + q), y = fr::ldg_f4(b + q);
                if (metric == 0) {
                    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
                } else {
                    const float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
                    s = fmaf(e0, e0, s); s = fmaf(e1, e1, s); s = fmaf(e2, e2, s); s = fmaf(e3, e3, s);
                }
            }
            s = fr::warp_sum(s);
            const float sc = metric == 0 ? s * scale + (bias ? __ldg(bias + cc) : 0.f) : -s;
            if (lane == (c & 31)) v[c >> 5] = sc;
        }
    } else {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (ci[t] < 0) continue;
            const float *b = B + (size_t)ci[t] * d;
            float s = 0.f;
            if (metric == 0) {
                for (int q = 0; q < d; q += 4) {
                    const float4 x = fr::ldg_f4(a + q), y = fr::ldg_f4(b + q);
                    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
                }
                v[t] = s * scale + (bias ? __ldg(bias + ci[t]) : 0.f);
            } else {  // negative squared Euclidean distance: no cancellation, unlike x.c - |c|^2/2
                for (int q = 0; q < d; q += 4) {
                    const float4 x = fr::ldg_f4(a + q), y = fr::ldg_f4(b + q);
                    const float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
                    s = fmaf(e0, e0, s); s = fmaf(e1, e1, s); s = fmaf(e2, e2, s); s = fmaf(e3, e3, s);
                }
                v[t] = -s;
            }
        }
    }
    float vk = -INFINITY;                     // k-th best exact score among the candidates
    int emitted = 0;
    for (int o = 0; o < k; ++o) {
        // lane-local best, then warp arg-max (value desc, column asc)
        int t = (ci[1] >= 0 && (ci[0] < 0 || v[1] > v[0] || (v[1] == v[0] && ci[1] < ci[0]))) ? 1 : 0;
        float bv = ci[t] >= 0 ? v[t] : -INFINITY;
        int bi = ci[t] >= 0 ? ci[t] : 0x7fffffff;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            out_val[(size_t)row * k + o] = bi == 0x7fffffff ? -INFINITY : bv;
            out_idx[(size_t)row * k + o] = bi == 0x7fffffff ? (IdxT)-1 : (IdxT)bi;
        }
        if (bi != 0x7fffffff) { vk = bv; ++emitted; }
        if (ci[0] == bi) ci[0] = -1;
        if (ci[1] == bi) ci[1] = -1;
    }
    if (cert == nullptr) return;
    bool ok = last_cand < 0;                  // every eligible column was a candidate
    if (!ok && emitted == k) {
        float an = 0.f, rn = 0.f, en = 0.f;              // |a|^2, |a'|^2, |a' - a|^2
        for (int q = lane * 4; q < d; q += 128) {
            const float4 x = fr::ldg_f4(a + q);
            an = fmaf(x.x, x.x, an); an = fmaf(x.y, x.y, an); an = fmaf(x.z, x.z, an); an = fmaf(x.w, x.w, an);
            bf16_norms(x, rn, en);
        }
        an = fr::warp_sum(an);
        rn = fr::warp_sum(rn);
        en = fr::warp_sum(en);
        // exact score in the units the bf16 pass ranked by: metric 1 ranks x.c - |c|^2/2 = (|x|^2 - |x - c|^2) / 2
        const float vk_key = metric == 0 ? vk : 0.5f * (an + vk);
        const float bn = __ldg(bmax), be = __ldg(bmax + 1);
        const float E = 1.0001f * fabsf(metric == 0 ? scale : 1.f) *
                            (sqrtf(en) * bn + sqrtf(an) * be + (float)d * 1.1920929e-07f * sqrtf(rn) * bn)
                        + 1e-6f * fabsf(vk_key);            // fp32 rounding of the re-score and of the bias add
        ok = cand_val[(size_t)row * kc + kc - 1] + E < vk_key;
    }
    if (lane == 0) cert[row] = ok ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- exact fp32 rows
// scores[r, c] = scale * <A[a_rows[r]], B[c]> + bias[c]   (metric 1: -|A[a_rows[r]] - B[c]|^2), 64 x 64 tile per CTA,
// 4 x 4 outputs per thread, k-slices of 16 staged in shared memory (SIMT fp32: this path serves the rare rows whose
// certificate failed, exactness matters, not throughput).
constexpr int XT = 64, XK = 16;
constexpr int LCAP = 512;          // slots of a row's above-threshold list (filtered variant)
// FILTER: nothing dense is written.  A score goes to the row's list (score, column) only if it reaches the row's
// threshold `thr[r]` -- a lower bound of the row's k-th best fp32 score, known from the re-scored candidates -- and the
// column is not in the row's history.  Every member of the fp32 top-k passes; about k + a few columns per row do.
struct ExactFilter {
    const float *thr;              // [Mf]
    float2 *list;                  // [Mf, LCAP] (score, column bits)
    int *count;                    // [Mf] zeroed; may exceed LCAP (overflow: the caller re-runs the row densely)
    const int64_t *hist_rows, *hist_ptr;
    const int32_t *hist_idx;
};
template <bool FILTER>
__global__ void __launch_bounds__(256)
exact_scores_kernel(const float *__restrict__ A, const int64_t *__restrict__ a_rows, int Mf, const float *__restrict__ B,
                    int N, int d, float scale, const float *__restrict__ bias, int metric, float *__restrict__ S,
                    const ExactFilter F) {
    __shared__ float As[XK][XT + 4], Bs[XK][XT + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r0 = blockIdx.y * XT, c0 = blockIdx.x * XT;
    float acc[4][4] = {};
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;          // loader: row lr, k offset lk..lk+3
    const int ar = r0 + lr, bc = c0 + lr;
    const float *ap = ar < Mf ? A + (size_t)a_rows[ar] * d : nullptr;
    const float *bp = bc < N ? B + (size_t)bc * d : nullptr;
    for (int k0 = 0; k0 < d; k0 += XK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (ap && k0 + lk < d) av = fr::ldg_f4(ap + k0 + lk);
        if (bp && k0 + lk < d) bv = fr::ldg_f4(bp + k0 + lk);
        As[lk][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
        Bs[lk][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < XK; ++kk) {
            float a4[4], b4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a4[i] = As[kk][ty * 4 + i]; b4[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (metric == 0) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
                    else { const float e = a4[i] - b4[j]; acc[i][j] = fmaf(e, e, acc[i][j]); }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r >= Mf) continue;
        const float thr = FILTER ? __ldg(F.thr + r) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c >= N) continue;
            const float sc = metric == 0 ? acc[i][j] * scale + (bias ? __ldg(bias + c) : 0.f) : -acc[i][j];
            if (!FILTER) {
                S[(size_t)r * N + c] = sc;
            } else if (sc >= thr) {
                bool masked = false;
                if (F.hist_rows != nullptr) {           // binary search in the row's sorted history
                    const long long id = F.hist_rows[r];
                    long long lo = F.hist_ptr[id], hi = F.hist_ptr[id + 1];
                    while (lo < hi) {
                        const long long mid = (lo + hi) >> 1;
                        if (F.hist_idx[mid] < c) lo = mid + 1; else hi = mid;
                    }
                    masked = lo < F.hist_ptr[id + 1] && F.hist_idx[lo] == c;
                }
                if (!masked) {
                    const int slot = atomicAdd(F.count + r, 1);
                    if (slot < LCAP) F.list[(size_t)r * LCAP + slot] = make_float2(sc, __int_as_float(c));
                }
            }
        }
    }
}

// One CTA per row of the filtered variant: bitonic sort of the row's list by (score desc, column asc) in shared
// memory, first k out.  Rows whose list overflowed are left untouched and counted in *overflow.
__global__ void __launch_bounds__(256)
exact_list_select_kernel(const float2 *__restrict__ list, const int *__restrict__ count, int k,
                         float *__restrict__ out_val, int64_t *__restrict__ out_idx, int *__restrict__ overflow) {
    __shared__ unsigned long long key[LCAP];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int n = count[r];
    if (n > LCAP) {
        if (tid == 0) atomicAdd(overflow, 1);
        return;
    }
    for (int i = tid; i < LCAP; i += 256) {
        unsigned long long kv = 0ull;                    // padding sorts last
        if (i < n) {
            const float2 e = list[(size_t)r * LCAP + i];
            // larger score first, then smaller column first: (order key, ~column) descending
            kv = ((unsigned long long)order_key(e.x) << 32) | (unsigned)(~__float_as_int(e.y));
        }
        key[i] = kv;
    }
    __syncthreads();
    for (int size = 2; size <= LCAP; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < LCAP / 2; i += 256) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = key[lo], b = key[hi];
                if ((a < b) == desc) { key[lo] = b; key[hi] = a; }
            }
            __syncthreads();
        }
    if (tid < k) {
        const bool have = tid < n;
        const unsigned long long kv = key[tid];
        const float v = have ? rk::key_value((unsigned)(kv >> 32)) : -INFINITY;
        out_val[(size_t)r * k + tid] = v;
        out_idx[(size_t)r * k + tid] = (have && v != -INFINITY) ? (int64_t)(int)(~(unsigned)(kv & 0xffffffffull)) : -1;
    }
}

// history columns of every row -> -inf (one warp per row)
__global__ void exact_mask_kernel(float *__restrict__ S, int Mf, int N, const int64_t *__restrict__ hist_rows,
                                  const int64_t *__restrict__ hist_ptr, const int32_t *__restrict__ hist_idx) {
    const int lane = threadIdx.x & 31;
    const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= Mf) return;
    const long long id = hist_rows[r];
    for (long long h = hist_ptr[id] + lane; h < hist_ptr[id + 1]; h += 32) {
        const int c = hist_idx[h];
        if (c >= 0 && c < N) S[(size_t)r * N + c] = -INFINITY;
    }
}

// One CTA per row: exact k-th largest key by a 4 x 8-bit radix select over the N scores, then the winners
// (strictly larger first; equal keys by ascending column), ordered by (value desc, column asc).
__global__ void __launch_bounds__(256)
exact_select_kernel(const float *__restrict__ S, int N, int k, float *__restrict__ out_val, int64_t *__restrict__ out_idx) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_need, s_cnt_gt, s_cnt_eq, s_warp_eq[8];
    __shared__ float win_v[rk::MAXK];
    __shared__ int win_i[rk::MAXK];
    const float *s = S + (size_t)blockIdx.x * N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int kk = min(k, N);
    if (tid == 0) { s_prefix = 0u; s_need = (unsigned)kk; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix, pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int c = tid; c < N; c += 256) {
            const unsigned key = order_key(s[c]);
            if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {       // walk the bins from the top: the bin holding the `need`-th largest key
            unsigned need = s_need, b = 255;
            for (;; --b) {
                if (hist[b] >= need || b == 0) break;
                need -= hist[b];
            }
            s_prefix = prefix | (b << shift);
            s_need = need;
        }
        __syncthreads();
    }
    const unsigned Tk = s_prefix;                // the kk-th largest key; s_need of its copies are wanted
    const unsigned need_eq = s_need;
    if (tid == 0) { s_cnt_gt = 0u; s_cnt_eq = 0u; }
    __syncthreads();
    for (int c0 = 0; c0 < N; c0 += 256) {
        const int c = c0 + tid;
        const unsigned key = c < N ? order_key(s[c]) : 0u;
        const bool gt = c < N && key > Tk, eq = c < N && key == Tk;
        if (gt) {
            const unsigned slot = atomicAdd(&s_cnt_gt, 1u);      // < kk - need_eq by construction
            win_v[slot] = s[c];
            win_i[slot] = c;
        }
        // equal keys are taken in ascending column order: ordered block scan of the `eq` flags
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp_eq[wid] = __popc(bal);
        __syncthreads();
        if (eq) {
            unsigned before = s_cnt_eq + __popc(bal & ((1u << lane) - 1u));
            for (int w = 0; w < wid; ++w) before += s_warp_eq[w];
            if (before < need_eq) {
                const int slot = kk - (int)need_eq + (int)before;
                win_v[slot] = s[c];
                win_i[slot] = c;
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0;
            for (int w = 0; w < 8; ++w) t += s_warp_eq[w];
            s_cnt_eq += t;
        }
        __syncthreads();
    }
    if (tid < k) {
        if (tid < kk) {
            const float v = win_v[tid];
            const int ix = win_i[tid];
            int rank = 0;
            for (int j = 0; j < kk; ++j) rank += (win_v[j] > v || (win_v[j] == v && win_i[j] < ix)) ? 1 : 0;
            const bool dead = v == -INFINITY;                    // masked / padded column: report as absent
            out_val[(size_t)blockIdx.x * k + rank] = v;
            out_idx[(size_t)blockIdx.x * k + rank] = dead ? -1 : ix;
        } else {
            out_val[(size_t)blockIdx.x * k + tid] = -INFINITY;
            out_idx[(size_t)blockIdx.x * k + tid] = -1;
        }
    }
}
}  // namespace

extern "C" int fr_max_row_norm(const float *B, int64_t rows, int32_t d, float *out, void *stream) {
    FR_REQUIRE(B && out && rows >= 0 && d > 0 && d % 4 == 0, "fr_max_row_norm: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, 2 * sizeof(float), st);
    if (e != cudaSuccess) {
        fr::set_error("fr_max_row_norm: %s", cudaGetErrorString(e));
        return FR_ECUDA;
    }
    if (rows == 0) return FR_OK;
    const int blocks = (int)std::min<int64_t>((rows * 32 + 255) / 256, (int64_t)fr::num_sms() * 8);
    fr::LaunchTimer _lt("max_row_norm_kernel", st);
    max_row_norm_kernel<<<blocks, 256, 0, st>>>(B, rows, d, reinterpret_cast<unsigned *>(out));
    return fr::check_launch("fr_max_row_norm");
}

extern "C" int fr_rescore_topk_f32(const float *A, const int64_t *a_rows, const float *B, int32_t d, float scale,
                                   const float *bias, int32_t metric, const int32_t *cand, const float *cand_val,
                                   int32_t kc, int32_t M, int32_t k, float *out_val, void *out_idx, int32_t idx64,
                                   const float *bmax, uint8_t *cert, void *stream) {
    FR_REQUIRE(M >= 0 && d > 0 && d % 4 == 0, "fr_rescore_topk_f32: M=%d d=%d", M, d);
    if (M == 0) return FR_OK;
    FR_REQUIRE(A && B && cand && out_val && out_idx, "fr_rescore_topk_f32: null pointer");
    FR_REQUIRE(kc >= 1 && kc <= 64 && k >= 1 && k <= kc, "fr_rescore_topk_f32: k=%d kc=%d", k, kc);
    FR_REQUIRE(metric == 0 || metric == 1, "fr_rescore_topk_f32: metric=%d", metric);
    FR_REQUIRE(cert == nullptr || (cand_val != nullptr && bmax != nullptr),
               "fr_rescore_topk_f32: the certificate needs the bf16 candidate scores and max |B row|");
    const long long blocks = ((long long)M * 32 + 255) / 256;
    cudaStream_t st = (cudaStream_t)stream;
    fr::LaunchTimer _lt("rescore_topk_kernel", st);
    if (idx64)
        rescore_topk_kernel<int64_t><<<(unsigned)blocks, 256, 0, st>>>(A, a_rows, B, d, scale, bias, metric, cand, cand_val, kc, M, k,
                                                                     out_val, reinterpret_cast<int64_t *>(out_idx), bmax, cert);
    else
        rescore_topk_kernel<int32_t><<<(unsigned)blocks, 256, 0, st>>>(A, a_rows, B, d, scale, bias, metric, cand, cand_val, kc, M, k,
                                                                     out_val, reinterpret_cast<int32_t *>(out_idx), bmax, cert);
    return fr::check_launch("fr_rescore_topk_f32");
}

extern "C" int fr_exact_topk_f32(const float *A, const int64_t *a_rows, int32_t Mf, const float *B, int32_t N, int32_t d,
                                 float scale, const float *bias, int32_t metric, const int64_t *hist_rows,
                                 const int64_t *hist_ptr, const int32_t *hist_idx, int32_t k, float *scores_ws,
                                 float *out_val, int64_t *out_idx, void *stream) {
    FR_REQUIRE(Mf >= 0 && N > 0 && d > 0 && d % 4 == 0, "fr_exact_topk_f32: Mf=%d N=%d d=%d", Mf, N, d);
    if (Mf == 0) return FR_OK;
    FR_REQUIRE(A && a_rows && B && scores_ws && out_val && out_idx, "fr_exact_topk_f32: null pointer");
    FR_REQUIRE(k >= 1 && k <= rk::MAXK, "fr_exact_topk_f32: k=%d out of [1, %d]", k, rk::MAXK);
    FR_REQUIRE(metric == 0 || metric == 1, "fr_exact_topk_f32: metric=%d", metric);
    FR_REQUIRE((hist_rows == nullptr) == (hist_ptr == nullptr) && (hist_rows == nullptr) == (hist_idx == nullptr),
               "fr_exact_topk_f32: hist_rows / hist_ptr / hist_idx go together");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((N + XT - 1) / XT, (Mf + XT - 1) / XT);
    {
        fr::LaunchTimer _lt("exact_scores_kernel", st);
        exact_scores_kernel<false><<<grid, 256, 0, st>>>(A, a_rows, Mf, B, N, d, scale, bias, metric, scores_ws, ExactFilter{});
        if (int rc = fr::check_launch("fr_exact_topk_f32(scores)")) return rc;
    }
    if (hist_rows != nullptr) {
        fr::LaunchTimer _lt("exact_mask_kernel", st);
        exact_mask_kernel<<<(Mf * 32 + 255) / 256, 256, 0, st>>>(scores_ws, Mf, N, hist_rows, hist_ptr, hist_idx);
        if (int rc = fr::check_launch("fr_exact_topk_f32(mask)")) return rc;
    }
    fr::LaunchTimer _lt("exact_select_kernel", st);
    exact_select_kernel<<<Mf, 256, 0, st>>>(scores_ws, N, k, out_val, out_idx);
    return fr::check_launch("fr_exact_topk_f32(select)");
}

extern "C" int64_t fr_exact_topk_thr_ws_bytes(int32_t Mf) { return (int64_t)Mf * (LCAP * 8 + 4); }

extern "C" int fr_exact_topk_thr_f32(const float *A, const int64_t *a_rows, int32_t Mf, const float *B, int32_t N, int32_t d,
                                     float scale, const float *bias, int32_t metric, const int64_t *hist_rows,
                                     const int64_t *hist_ptr, const int32_t *hist_idx, int32_t k, const float *thr, void *ws,
                                     float *out_val, int64_t *out_idx, int32_t *overflow, void *stream) {
    FR_REQUIRE(Mf >= 0 && N > 0 && d > 0 && d % 4 == 0, "fr_exact_topk_thr_f32: Mf=%d N=%d d=%d", Mf, N, d);
    if (Mf == 0) return FR_OK;
    FR_REQUIRE(A && a_rows && B && thr && ws && out_val && out_idx && overflow, "fr_exact_topk_thr_f32: null pointer");
    FR_REQUIRE(k >= 1 && k <= rk::MAXK, "fr_exact_topk_thr_f32: k=%d out of [1, %d]", k, rk::MAXK);
    FR_REQUIRE(metric == 0 || metric == 1, "fr_exact_topk_thr_f32: metric=%d", metric);
    FR_REQUIRE((hist_rows == nullptr) == (hist_ptr == nullptr) && (hist_rows == nullptr) == (hist_idx == nullptr),
               "fr_exact_topk_thr_f32: hist_rows / hist_ptr / hist_idx go together");
    FR_REQUIRE(((uintptr_t)ws & 7) == 0, "fr_exact_topk_thr_f32: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    ExactFilter F;
    F.thr = thr;
    F.list = reinterpret_cast<float2 *>(ws);
    F.count = reinterpret_cast<int *>(F.list + (size_t)Mf * LCAP);
    F.hist_rows = hist_rows;
    F.hist_ptr = hist_ptr;
    F.hist_idx = hist_idx;
    cudaError_t e = cudaMemsetAsync(F.count, 0, (size_t)Mf * sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(overflow, 0, sizeof(int), st);
    if (e != cudaSuccess) {
        fr::set_error("fr_exact_topk_thr_f32: %s", cudaGetErrorString(e));
        return FR_ECUDA;
    }
    dim3 grid((N + XT - 1) / XT, (Mf + XT - 1) / XT);
    {
        fr::LaunchTimer _lt("exact_scores_kernel<filter>", st);
        exact_scores_kernel<true><<<grid, 256, 0, st>>>(A, a_rows, Mf, B, N, d, scale, bias, metric, nullptr, F);
        if (int rc = fr::check_launch("fr_exact_topk_thr_f32(scores)")) return rc;
    }
    fr::LaunchTimer _lt("exact_list_select_kernel", st);
    exact_list_select_kernel<<<Mf, 256, 0, st>>>(F.list, F.count, k, out_val, out_idx, overflow);
    return fr::check_launch("fr_exact_topk_thr_f32(select)");
}
