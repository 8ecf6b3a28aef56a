// Previous generation of the fused score + top-K kernel (cta_group::1, one CTA per 128-row block, every CTA streams
// all of B).  Kept only as the A/B baseline of `rank_topk_pair_kernel` (gemm_topk.cu) while that kernel is being
// brought up on hardware: selected with FR_TOPK_IMPL=2.
#include <algorithm>

#include "rank_common.cuh"

namespace rk {

// ================================================================================================
// v2: the same MMA / TMA pipeline with SIXTEEN epilogue warps (four per SM sub-partition).
// Profiling v1 showed the tensor pipe 6 % busy and the four epilogue warps at 0.1-0.2 IPC: with one warp
// per scheduler every dependent instruction and branch is exposed.  Here the four warps that may read a
// TMEM lane group split each 256-column tile into 64-column quarters; a thread owns (row, quarter) and
// keeps its candidate list in an L2-resident global workspace (CAPG slots), which frees the shared
// memory that capped the warp count (and pays for a 4-stage operand ring).  The four quarter threads of
// a row share the best known lower bound of the row's k-th score through a shared-memory key
// (atomicMax at prune time), and their lists are merged, selected and sorted per row at the end of the
// row block.
// Two sweeps per row block (the MMA pipe is idle > 90 % of the time, so recomputing the tiles is free):
//   pass 0 (bounding): branch-free -- per (row, quarter) the running maximum of NG column groups goes to
//          shared memory (atomicMax on order-preserving keys).  The (k + h)-th largest group maximum T
//          (h = the row's history length, because a masked column may hold a group's maximum) is a lower
//          bound of the k-th eligible score: at least k eligible scores are >= T.
//   pass 1 (collection): the streaming top-k above, started from threshold T instead of -inf, so only
//          ~k scores per row ever take the candidate path.
// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr int EW2 = 16;
constexpr int THREADS2 = 64 + EW2 * 32;
constexpr int STAGES2 = 3;
constexpr int NG = 128;                        // column groups per row for the bounding pass
constexpr int CAPG = 96;                       // slots per (row, quarter) list; prune when > CAPG - 32
constexpr int SMEM2 = 1024 + STAGES2 * STAGE_BYTES + BM * 4 * 4 + BM * 4 + NG * BM * 4 + 256;

__device__ __noinline__ float warp_prune_g(float *bv, int *bi, int cnt, int k, int lane) {
    constexpr int T = CAPG / 32;
    float v[T];
    int ix[T];
    uint32_t key[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int s = lane + 32 * t;
        const bool ok = s < cnt;
        v[t] = ok ? __ldcg(bv + s) : 0.f;
        ix[t] = ok ? __ldcg(bi + s) : -1;
        key[t] = ok ? order_key(v[t]) : 0u;
    }
    uint32_t Tk = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c = Tk | (1u << bit);
        int n = 0;
#pragma unroll
        for (int t = 0; t < T; ++t) n += __popc(__ballot_sync(0xffffffffu, key[t] >= c));
        if (n >= k) Tk = c;
    }
    int g = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) g += __popc(__ballot_sync(0xffffffffu, key[t] > Tk));
    int need_eq = k - g, base = 0;
    const uint32_t below = (1u << lane) - 1u;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const uint32_t mg = __ballot_sync(0xffffffffu, key[t] > Tk);
        const bool is_eq = key[t] == Tk && ix[t] >= 0;
        const uint32_t me = __ballot_sync(0xffffffffu, is_eq);
        const bool keep_eq = is_eq && __popc(me & below) < need_eq;
        const uint32_t mk = mg | __ballot_sync(0xffffffffu, keep_eq);
        if ((mk >> lane) & 1u) {
            const int dst = base + __popc(mk & below);
            __stcg(bv + dst, v[t]);
            __stcg(bi + dst, ix[t]);
        }
        base += __popc(mk);
        need_eq -= min(need_eq, __popc(me));
    }
    __threadfence_block();
    __syncwarp();
    return key_value(Tk);
}

struct Params2 {
    Params p;
    float *lv;   // [grid][BM][4][CAPG]
    int *li;
    int two_pass;
    int bstride;   // the bounding sweep visits every bstride-th column tile (a subset still bounds from below)
};

template <bool AFFINE>   // AFFINE: scores are scale * acc + bias[col]; the plain instantiation carries none of that code
__global__ void __launch_bounds__(THREADS2, 1)
gemm_topk_kernel_v2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params2 PP) {
    const Params &P = PP.p;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles = smem;
    int *cnt_s = reinterpret_cast<int *>(smem + STAGES2 * STAGE_BYTES);   // [BM][4]
    uint32_t *thr_key = reinterpret_cast<uint32_t *>(cnt_s + BM * 4);      // [BM] best known k-th key per row
    uint32_t *gkey = thr_key + BM;                                         // [NG][BM] group maxima (pass 0)
    uint64_t *bars = reinterpret_cast<uint64_t *>(((uintptr_t)(gkey + NG * BM) + 7) & ~(uintptr_t)7);
    uint64_t *full = bars, *empty = bars + STAGES2, *tfull = bars + 2 * STAGES2, *tempty = bars + 2 * STAGES2 + ACC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES2 + 2 * ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mblk = (P.M + BM - 1) / BM, n_nblk = (P.N + BN - 1) / BN, n_kblk = (P.K + BK - 1) / BK;
    const int n_pass = PP.two_pass ? 2 : 1;
    const int bstride = PP.bstride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, EW2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < BM) thr_key[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < NG * BM; i += THREADS2) gkey[i] = 0u;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mb = blockIdx.x; mb < n_mblk; mb += gridDim.x)
              for (int pass = 0; pass < n_pass; ++pass)
                for (int nb = 0; nb < n_nblk; nb += (n_pass == 2 && pass == 0) ? bstride : 1)
                    for (int kb = 0; kb < n_kblk; ++kb) {
                        mbar_wait(empty + stage, phase ^ 1);
                        uint8_t *a = tiles + stage * STAGE_BYTES, *b = a + A_BYTES;
                        mbar_expect_tx(full + stage, STAGE_BYTES);
                        tma_load_2d(&tmA, full + stage, a, kb * BK, mb * BM);
                        tma_load_2d(&tmB, full + stage, b, kb * BK, nb * BN);
                        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        for (int mb = blockIdx.x; mb < n_mblk; mb += gridDim.x)
          for (int pass = 0; pass < n_pass; ++pass)
            for (int nb = 0; nb < n_nblk; nb += (n_pass == 2 && pass == 0) ? bstride : 1) {
                if (lane == 0) mbar_wait(tempty + as, aphase ^ 1);
                __syncwarp();
                tc_fence_after();
                for (int kb = 0; kb < n_kblk; ++kb) {
                    if (lane == 0) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t a = smem_u32(tiles + stage * STAGE_BYTES);
                        const uint64_t ad = make_desc(a), bd = make_desc(a + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_f16(tmem_base + as * BN, ad + 2 * k, bd + 2 * k, kIdesc, (kb | k) != 0);
                        umma_commit(empty + stage);
                        if (kb == n_kblk - 1) umma_commit(tfull + as);
                    }
                    __syncwarp();
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
            }
    } else {
        const int ew = warp - 2;
        const int lg = warp & 3;                 // TMEM lane group this warp may read
        const int q = ew >> 2;                   // column quarter of every tile handled by this warp
        const int r_in_blk = lg * 32 + lane;
        const int kk = P.topk;
        const size_t list0 = (((size_t)blockIdx.x * BM + lg * 32) * 4 + q) * CAPG;   // lane 0's list of this warp
        float *lv = PP.lv + list0 + (size_t)lane * 4 * CAPG;
        int *li = PP.li + list0 + (size_t)lane * 4 * CAPG;
        int as = 0;
        uint32_t aphase = 0;
        for (int mb = blockIdx.x; mb < n_mblk; mb += gridDim.x) {
            const int row = mb * BM + r_in_blk;
            float thr = -INFINITY;
            int cnt = 0;
            long long hlo = 0, hhi = 0;
            if (P.row_ids != nullptr && row < P.M) {
                const long long id = P.row_ids[row];
                hlo = P.hist_ptr[id];
                hhi = P.hist_ptr[id + 1];
            }
            if (n_pass == 2) {
                // ---------------- pass 0: group maxima -> lower bound of the k-th eligible score
                float gmax = -INFINITY;
                // Only every bstride-th column tile is visited: the (k + h)-th largest group maximum of a SUBSET of
                // the columns is still a lower bound of the k-th eligible score over all of them -- a looser one
                // (about bstride x more scores reach the candidate path of the collection sweep, still a vanishing
                // fraction), for 1 / bstride of the bounding work.
                // group id of (visited tile i, quarter q) = floor((4 i + q) NG / (4 n_vis)), advanced incrementally:
                // num = (4 i + q) NG - gid * den stays in [0, den) (no per-tile 64-bit division)
                const long long den = 4LL * ((n_nblk + bstride - 1) / bstride);
                int gid = (int)(((long long)q * NG) / den);
                long long num = (long long)q * NG - (long long)gid * den;
                int gcur = gid;
                for (int nb = 0; nb < n_nblk; nb += bstride) {
                    if (lane == 0) mbar_wait(tfull + as, aphase);
                    __syncwarp();
                    tc_fence_after();
                    if (gid != gcur) {
                        if (gmax > -INFINITY) atomicMax(gkey + gcur * BM + r_in_blk, order_key(gmax));
                        gmax = -INFINITY;
                        gcur = gid;
                    }
                    const uint32_t tbase = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN;
#pragma unroll 1
                    for (int c = 2 * q; c < 2 * q + 2; ++c) {
                        float r[32];
                        tmem_ld32(tbase + c * 32, r);
                        const int col0 = nb * BN + c * 32;
                        if (col0 >= P.N) break;
                        if (AFFINE && P.bias != nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int col = col0 + j;
                                r[j] = fmaf(r[j], P.scale, col < P.N ? __ldg(P.bias + col) : 0.f);
                            }
                        } else if (AFFINE && P.scale != 1.f) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) r[j] *= P.scale;
                        }
                        if (col0 + 32 > P.N) {          // zero-filled columns past N must not bound anything
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j >= P.N) r[j] = -INFINITY;
                        }
                        float m0 = max3(r[0], r[1], r[2]), m1 = max3(r[3], r[4], r[5]);
#pragma unroll
                        for (int j = 6; j < 30; j += 6) {
                            m0 = max3(m0, r[j], r[j + 1]);
                            m1 = max3(m1, r[j + 2], r[j + 3]);
                            m0 = fmaxf(m0, r[j + 4]);
                            m1 = fmaxf(m1, r[j + 5]);
                        }
                        gmax = fmaxf(gmax, max3(m0, m1, fmaxf(r[30], r[31])));
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + as);
                    if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
                    num += 4LL * NG;                      // next visited tile: (4 (i + 1) + q) NG
                    while (num >= den) { num -= den; ++gid; }
                }
                if (gmax > -INFINITY) atomicMax(gkey + gcur * BM + r_in_blk, order_key(gmax));
                __threadfence_block();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "r"(128) : "memory");
                // warp q of the lane group bounds rows q*8 .. q*8+7: the (k + h)-th largest of NG group maxima
                for (int rr = 0; rr < 8; ++rr) {
                    const int rib = lg * 32 + q * 8 + rr;
                    uint32_t key[NG / 32];
#pragma unroll
                    for (int t = 0; t < NG / 32; ++t) {
                        key[t] = gkey[(lane + 32 * t) * BM + rib];
                        gkey[(lane + 32 * t) * BM + rib] = 0u;       // ready for the next row block
                    }
                    int h = 0;
                    const int orow = mb * BM + rib;
                    if (P.row_ids != nullptr && orow < P.M) {
                        const long long id = P.row_ids[orow];
                        { const long long hl = P.hist_ptr[id + 1] - P.hist_ptr[id]; h = hl > NG ? NG : (int)hl; }
                    }
                    const int want = kk + h;
                    uint32_t Tk = 0;
                    if (want <= NG) {
                        for (int bit = 31; bit >= 0; --bit) {
                            const uint32_t c = Tk | (1u << bit);
                            int n = 0;
#pragma unroll
                            for (int t = 0; t < NG / 32; ++t) n += __popc(__ballot_sync(0xffffffffu, key[t] >= c));
                            if (n >= want) Tk = c;
                        }
                    }
                    // collection accepts `score > thr`: publish the key just below T so that ties with T pass
                    if (lane == 0) thr_key[rib] = Tk > 1u ? Tk - 1u : 0u;
                }
                __threadfence_block();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "r"(128) : "memory");
            }
            for (int nb = 0; nb < n_nblk; ++nb) {
                if (lane == 0) mbar_wait(tfull + as, aphase);   // one poller per warp; the rest park at the syncwarp
                __syncwarp();
                tc_fence_after();
                const uint32_t shared_key = thr_key[r_in_blk];
                if (shared_key != 0u) thr = fmaxf(thr, key_value(shared_key));
                const uint32_t tbase = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN;
#pragma unroll 1
                for (int c = 2 * q; c < 2 * q + 2; ++c) {
                    float r[32];
                    tmem_ld32(tbase + c * 32, r);
                    const int col0 = nb * BN + c * 32;
                    if (col0 >= P.N) break;
                    if (AFFINE && P.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = col0 + j;
                            r[j] = fmaf(r[j], P.scale, col < P.N ? __ldg(P.bias + col) : 0.f);
                        }
                    } else if (AFFINE && P.scale != 1.f) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] *= P.scale;
                    }
                    // hot path: 3-input max tree over four 8-column groups, one compare per 32 scores
                    float g[4];
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq)
                        g[gq] = max3(max3(r[8 * gq], r[8 * gq + 1], r[8 * gq + 2]),
                                     max3(r[8 * gq + 3], r[8 * gq + 4], r[8 * gq + 5]), fmaxf(r[8 * gq + 6], r[8 * gq + 7]));
                    const float mx = fmaxf(max3(g[0], g[1], g[2]), g[3]);
                    if (mx > thr) {
                        // only the lanes (rows) that have a candidate come here, and each walks only the
                        // 8-column groups that hold one: the cost follows the number of candidates
                        const int valid = min(32, P.N - col0);
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            if (g[gq] > thr) {
#pragma unroll
                                for (int j = 8 * gq; j < 8 * gq + 8; ++j) {
                                    if (j < valid && r[j] > thr &&
                                        (hlo == hhi || !in_history(P.hist_idx, hlo, hhi, col0 + j))) {
                                        __stcg(lv + cnt, r[j]);
                                        __stcg(li + cnt, col0 + j);
                                        ++cnt;
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    uint32_t need = __ballot_sync(0xffffffffu, cnt > CAPG - 32);
                    if (need) { __threadfence_block(); __syncwarp(); }
                    while (need) {
                        const int src = __ffs(need) - 1;
                        need &= need - 1;
                        const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                        const float t_new = warp_prune_g(PP.lv + list0 + (size_t)src * 4 * CAPG,
                                                         PP.li + list0 + (size_t)src * 4 * CAPG, c_src, kk, lane);
                        if (lane == src) {
                            thr = fmaxf(thr, t_new);
                            cnt = kk;
                            atomicMax(thr_key + r_in_blk, order_key(t_new));   // a bound every quarter of the row may use
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + as);
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
            }
            // ---- end of the row block: trim own lists to k, publish counts, merge the four quarters per row
            {
                uint32_t need = __ballot_sync(0xffffffffu, cnt > kk);
                if (need) { __threadfence_block(); __syncwarp(); }
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                    warp_prune_g(PP.lv + list0 + (size_t)src * 4 * CAPG, PP.li + list0 + (size_t)src * 4 * CAPG, c_src, kk, lane);
                    if (lane == src) cnt = kk;
                }
            }
            cnt_s[r_in_blk * 4 + q] = cnt;
            __threadfence_block();
            asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "r"(128) : "memory");
            for (int rr = 0; rr < 8; ++rr) {
                const int rib = lg * 32 + q * 8 + rr;
                const int orow = mb * BM + rib;
                if (orow >= P.M) break;
                const int c0 = cnt_s[rib * 4], c1 = cnt_s[rib * 4 + 1], c2 = cnt_s[rib * 4 + 2], c3 = cnt_s[rib * 4 + 3];
                const int total = c0 + c1 + c2 + c3;       // <= 4 kk <= 256
                const size_t rbase = ((size_t)blockIdx.x * BM + rib) * 4 * CAPG;
                float v[8];
                int ix[8];
                uint32_t key[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    int e = lane + 32 * t;
                    const bool ok = e < total;
                    int qq = 0;
                    if (e >= c0) { e -= c0; qq = 1; if (e >= c1) { e -= c1; qq = 2; if (e >= c2) { e -= c2; qq = 3; } } }
                    v[t] = ok ? __ldcg(PP.lv + rbase + (size_t)qq * CAPG + e) : 0.f;
                    ix[t] = ok ? __ldcg(PP.li + rbase + (size_t)qq * CAPG + e) : -1;
                    key[t] = ok ? order_key(v[t]) : 0u;
                }
                if (total > kk) {          // keep exactly the kk best (radix select on the keys)
                    uint32_t Tk = 0;
                    for (int bit = 31; bit >= 0; --bit) {
                        const uint32_t c = Tk | (1u << bit);
                        int n = 0;
#pragma unroll
                        for (int t = 0; t < 8; ++t) n += __popc(__ballot_sync(0xffffffffu, key[t] >= c));
                        if (n >= kk) Tk = c;
                    }
                    int g = 0;
#pragma unroll
                    for (int t = 0; t < 8; ++t) g += __popc(__ballot_sync(0xffffffffu, key[t] > Tk));
                    int need_eq = kk - g;
                    const uint32_t below = (1u << lane) - 1u;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const bool is_eq = key[t] == Tk && ix[t] >= 0;
                        const uint32_t me = __ballot_sync(0xffffffffu, is_eq);
                        const bool keep = key[t] > Tk || (is_eq && __popc(me & below) < need_eq);
                        need_eq -= min(need_eq, __popc(me));
                        if (!keep) ix[t] = -1;
                    }
                }
                for (int o = 0; o < kk; ++o) {
                    float bv = -INFINITY;
                    int bi = 0x7fffffff;
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (ix[t] >= 0 && (bi == 0x7fffffff || v[t] > bv || (v[t] == bv && ix[t] < bi))) { bv = v[t]; bi = ix[t]; }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
                    }
                    if (lane == 0) {
                        P.out_val[(size_t)orow * kk + o] = bi == 0x7fffffff ? -INFINITY : bv;
                        P.out_idx[(size_t)orow * kk + o] = bi == 0x7fffffff ? -1 : bi;
                    }
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (ix[t] == bi) ix[t] = -1;
                }
            }
            if (q == 0) thr_key[r_in_blk] = 0u;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "r"(128) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

int launch_topk_v2(const CUtensorMap &ma, const CUtensorMap &mb, const Params &P, void *ws, int64_t ws_bytes, int two_pass,
                   int bstride, cudaStream_t stream) {
    const int n_mblk = (P.M + BM - 1) / BM;
    const int grid = std::min(n_mblk, fr::num_sms());
    const int64_t need = (int64_t)grid * BM * 4 * CAPG * 8;
    if (ws == nullptr || ws_bytes < need) {
        fr::set_error("fr_gemm_topk_bf16: workspace of %lld bytes required (got %lld)", (long long)need, (long long)ws_bytes);
        return FR_EINVAL;
    }
    static bool attr2 = false;
    if (!attr2) {
        cudaError_t e = cudaFuncSetAttribute(gemm_topk_kernel_v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_topk_kernel_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2);
        if (e != cudaSuccess) {
            fr::set_error("fr_gemm_topk_bf16: cannot reserve %d bytes of shared memory: %s", SMEM2, cudaGetErrorString(e));
            return FR_ECUDA;
        }
        attr2 = true;
    }
    Params2 P2{P, reinterpret_cast<float *>(ws),
               reinterpret_cast<int *>(reinterpret_cast<float *>(ws) + (size_t)grid * BM * 4 * CAPG), two_pass, bstride};
    fr::LaunchTimer _lt2("gemm_topk_kernel_v2", stream);
    if (P.bias != nullptr || P.scale != 1.f)
        gemm_topk_kernel_v2<true><<<grid, THREADS2, SMEM2, stream>>>(ma, mb, P2);
    else
        gemm_topk_kernel_v2<false><<<grid, THREADS2, SMEM2, stream>>>(ma, mb, P2);
    return fr::check_launch("fr_gemm_topk_bf16(v2)");
}

int64_t topk_v2_ws_bytes(int32_t M) {
    const int n_mblk = (M + BM - 1) / BM;
    return (int64_t)std::min(n_mblk, fr::num_sms()) * BM * 4 * CAPG * 8;
}

}  // namespace rk
