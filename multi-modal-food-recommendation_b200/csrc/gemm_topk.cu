// Fused  scores = A . B^T (bf16 in, fp32 accumulate)  ->  per-row top-k,  tcgen05 / TMEM / TMA, sm_100a.
//
// One kernel family for the three ranking shapes of the path:
//   full-sort      A = user_all[users] [M, 64],  B = item_all [N, 64], optional history mask, k <= 64
//   cosine kNN     A = B = row-normalised features [N, D], k = knn_k (self kept, utils.py:119)
//   centroids      A = features [I, D], B = centres [C, D], bias = -|c|^2 / 2, scale 1  (argmin |x - c|)
// The [M, N] score matrix never exists in HBM: a 128 x 256 fp32 tile lives in TMEM (two stages), the
// four epilogue warps read it back with tcgen05.ld (one accumulator row per thread) and keep a
// running top-k per row in shared memory.  Hot path per score: one predicated compare against the
// row's current k-th value; inserts are rare (~ k ln(N/k) per row).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane issues tcgen05.mma, M = 128, N = 256, K = 16 per instruction, SWIZZLE_128B K-major operands),
// warps 2..5 = epilogue (TMEM lane group = warp_id % 4).  Pipelines: 3-stage smem ring (full/empty
// mbarriers, TMA complete_tx / tcgen05.commit) and a 2-stage TMEM ring (tmem_full / tmem_empty).
// Persistent over 128-row blocks of A; each CTA sweeps every 256-column block of B for its rows.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64;  // tile; BK bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int MAXK = 64;
constexpr int CAP = 128;          // candidate slots per row (append-only between prunes)
constexpr int CSTRIDE = CAP + 1;  // padded row stride: appends and warp-wide row reads both conflict-free
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + BM * CSTRIDE * 8 + 256;

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    // (an explicit suspend-time hint of 1 us was measured: 40.0 -> 47.6 ms on 200k x 500k; the default wins)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must fault the launch, never hang the GPU.  The clock is consulted only
// every 256 failed polls (each poll already sleeps in hardware), so the common path is poll + branch.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try(bar, parity)) return;
        if ((spins & 255u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) {
                printf("foodrec_b200 gemm_topk: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
    uint32_t *u = reinterpret_cast<uint32_t *>(r);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in
// [0,14), leading byte offset >> 4 in [16,30) (= 1 for swizzled K-major), stride byte offset >> 4 in [32,46)
// (8 rows x 128 B = 1024 B between row groups), descriptor version 1 in [46,48), layout type 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Params {
    int M, N, K, topk;
    float scale;
    const float *bias;          // [N] or null, added after scaling
    const int64_t *row_ids;     // [M] id of each A row in the history CSR, null = no mask
    const int64_t *hist_ptr;    // CSR over ids
    const int32_t *hist_idx;    // sorted ascending within a row
    float *out_val;             // [M, topk]
    int32_t *out_idx;           // [M, topk], -1 where fewer than topk columns were eligible
};

// Running top-k of one accumulator row = an append-only candidate list in shared memory plus a
// threshold in a register.  A score enters the list iff it beats the threshold (one compare on the hot
// path, two stores when it does); when a list passes CAP - 32 entries the WARP prunes it together:
// radix-select of the k-th largest key by ballots, keep the k best, raise the threshold to the k-th.
__device__ __forceinline__ uint32_t order_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// All 32 lanes call this for the same row.  Returns the new threshold (k-th largest value); the
// list is compacted to exactly k entries (requires cnt > k).
__device__ __noinline__ float warp_prune(uint32_t bv, uint32_t bi, int cnt, int k, int lane) {  // shared addresses
    float v[CAP / 32];
    int ix[CAP / 32];
    uint32_t key[CAP / 32];
#pragma unroll
    for (int t = 0; t < CAP / 32; ++t) {
        const int s = lane + 32 * t;
        const bool ok = s < cnt;
        v[t] = ok ? __uint_as_float(lds32(bv + 4 * s)) : 0.f;
        ix[t] = ok ? (int)lds32(bi + 4 * s) : -1;
        key[t] = ok ? order_key(v[t]) : 0u;
    }
    uint32_t T = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c = T | (1u << bit);
        int n = 0;
#pragma unroll
        for (int t = 0; t < CAP / 32; ++t) n += __popc(__ballot_sync(0xffffffffu, key[t] >= c));
        if (n >= k) T = c;
    }
    int g = 0;
#pragma unroll
    for (int t = 0; t < CAP / 32; ++t) g += __popc(__ballot_sync(0xffffffffu, key[t] > T));
    int need_eq = k - g, base = 0;
    __syncwarp();
    const uint32_t below = (1u << lane) - 1u;
#pragma unroll
    for (int t = 0; t < CAP / 32; ++t) {
        const uint32_t mg = __ballot_sync(0xffffffffu, key[t] > T);
        const uint32_t me = __ballot_sync(0xffffffffu, key[t] == T && ix[t] >= 0);
        const int eq_rank = __popc(me & below);
        const bool keep_eq = (key[t] == T && ix[t] >= 0) && eq_rank < need_eq;
        const uint32_t mk = mg | __ballot_sync(0xffffffffu, keep_eq);
        if ((mk >> lane) & 1u) {
            const int dst = base + __popc(mk & below);
            sts32(bv + 4 * dst, __float_as_uint(v[t]));
            sts32(bi + 4 * dst, (uint32_t)ix[t]);
        }
        base += __popc(mk);
        need_eq -= min(need_eq, __popc(me));
    }
    __syncwarp();
    return key_value(T);
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ bool in_history(const int32_t *h, long long lo, long long hi, int col) {
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const int x = __ldg(h + mid);
        if (x == col) return true;
        if (x < col) lo = mid + 1; else hi = mid;
    }
    return false;
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles = smem;                                              // STAGES x (A | B), 1024-aligned
    float *topv = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);  // [BM][CSTRIDE] candidate values
    int *topi = reinterpret_cast<int *>(topv + BM * CSTRIDE);              // [BM][CSTRIDE] candidate columns
    uint64_t *bars = reinterpret_cast<uint64_t *>(((uintptr_t)(topi + BM * CSTRIDE) + 7) & ~(uintptr_t)7);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + ACC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 2 * ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mblk = (P.M + BM - 1) / BM, n_nblk = (P.N + BN - 1) / BN, n_kblk = (P.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 accumulator stages x 256 fp32 columns = all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mb = blockIdx.x; mb < n_mblk; mb += gridDim.x)
                for (int nb = 0; nb < n_nblk; ++nb)
                    for (int kb = 0; kb < n_kblk; ++kb) {
                        mbar_wait(empty + stage, phase ^ 1);
                        uint8_t *a = tiles + stage * STAGE_BYTES, *b = a + A_BYTES;
                        mbar_expect_tx(full + stage, STAGE_BYTES);
                        tma_load_2d(&tmA, full + stage, a, kb * BK, mb * BM);
                        tma_load_2d(&tmB, full + stage, b, kb * BK, nb * BN);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        for (int mb = blockIdx.x; mb < n_mblk; mb += gridDim.x)
            for (int nb = 0; nb < n_nblk; ++nb) {
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
                        for (int k = 0; k < BK / UMMA_K; ++k)  // +32 B per K step inside the swizzle atom
                            umma_f16(tmem_base + as * BN, ad + 2 * k, bd + 2 * k, kIdesc, (kb | k) != 0);
                        umma_commit(empty + stage);                       // frees the smem slot when the MMAs retire
                        if (kb == n_kblk - 1) umma_commit(tfull + as);    // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
            }
    } else {
        // ================================ epilogue: running top-k =====================
        const int lg = warp & 3;                 // TMEM lane group this warp may read
        const int r_in_blk = lg * 32 + lane;     // accumulator row owned by this thread
        const uint32_t topv_a = smem_u32(topv), topi_a = smem_u32(topi);   // explicit shared-space addressing
        const uint32_t tv = topv_a + 4u * r_in_blk * CSTRIDE, ti = topi_a + 4u * r_in_blk * CSTRIDE;
        const int kk = P.topk;
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
            for (int nb = 0; nb < n_nblk; ++nb) {
                mbar_wait(tfull + as, aphase);
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN;
                const int n0 = nb * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    float r[32];
                    tmem_ld32(tbase + c * 32, r);
                    const int col0 = n0 + c * 32;
                    if (col0 >= P.N) break;
                    if (P.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = col0 + j;
                            r[j] = fmaf(r[j], P.scale, col < P.N ? __ldg(P.bias + col) : 0.f);
                        }
                    } else if (P.scale != 1.f) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] *= P.scale;
                    }
                    // hot path: one 3-input max per two scores, one compare per 32
                    float mx = fmaxf(r[0], r[1]);
#pragma unroll
                    for (int j = 2; j < 32; j += 2) mx = fmaxf(fmaxf(r[j], r[j + 1]), mx);
                    if (!__any_sync(0xffffffffu, mx > thr)) continue;
                    // some row of this warp has a candidate among these 32 columns: walk the columns with
                    // warp-uniform branches so the cost follows the number of candidates, not 32 x divergence
                    const int valid = min(32, P.N - col0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const bool p = (j < valid) && (r[j] > thr);
                        if (__any_sync(0xffffffffu, p)) {
                            if (p && (hlo == hhi || !in_history(P.hist_idx, hlo, hhi, col0 + j))) {
                                sts32(tv + 4 * cnt, __float_as_uint(r[j]));
                                sts32(ti + 4 * cnt, (uint32_t)(col0 + j));
                                ++cnt;
                            }
                        }
                    }
                    // lists that could overflow on the next 32 columns are pruned by the whole warp
                    uint32_t need = __ballot_sync(0xffffffffu, cnt > CAP - 32);
                    if (need) __syncwarp();  // owners' appends visible to the lanes that help prune
                    while (need) {
                        const int src = __ffs(need) - 1;
                        need &= need - 1;
                        const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                        const uint32_t ro = 4u * (lg * 32 + src) * CSTRIDE;
                        const float t_new = warp_prune(topv_a + ro, topi_a + ro, c_src, kk, lane);
                        if (lane == src) { thr = t_new; cnt = kk; }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + as);
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
            }
            // ---- emit the rows of this warp: final prune to k, then k rounds of warp arg-max
            //      (descending by value, ties to the lower column)
            __syncwarp();
            for (int src = 0; src < 32; ++src) {
                int c_src = __shfl_sync(0xffffffffu, cnt, src);
                const int orow = mb * BM + lg * 32 + src;
                if (orow >= P.M) break;
                const uint32_t rv = topv_a + 4u * (lg * 32 + src) * CSTRIDE, ri = topi_a + 4u * (lg * 32 + src) * CSTRIDE;
                if (c_src > kk) { warp_prune(rv, ri, c_src, kk, lane); c_src = kk; }
                float v0 = lane < c_src ? __uint_as_float(lds32(rv + 4 * lane)) : -INFINITY;
                float v1 = lane + 32 < c_src ? __uint_as_float(lds32(rv + 4 * (lane + 32))) : -INFINITY;
                int i0 = lane < c_src ? (int)lds32(ri + 4 * lane) : -1, i1 = lane + 32 < c_src ? (int)lds32(ri + 4 * (lane + 32)) : -1;
                for (int o = 0; o < kk; ++o) {
                    const bool second = i1 >= 0 && (i0 < 0 || v1 > v0 || (v1 == v0 && i1 < i0));
                    float bv = second ? v1 : v0;
                    int bi = second ? i1 : i0;
                    if (bi < 0) bi = 0x7fffffff;
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
                    if (i0 == bi) i0 = -1;
                    if (i1 == bi) i1 = -1;
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}


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

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap *m, const void *base, int rows, int K, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        fr::set_error("gemm_topk: cuTensorMapEncodeTiled unavailable");
        return FR_ECUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fr::set_error("gemm_topk: cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
        return FR_ECUDA;
    }
    return FR_OK;
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, long long n,
                                   int d, int normalise) {
    // one warp per row when normalising (cosine), otherwise flat
    if (!normalise) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * d; i += (long long)gridDim.x * blockDim.x)
            y[i] = __float2bfloat16(x[i]);
        return;
    }
    const int lane = threadIdx.x & 31;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n;
         r += ((long long)gridDim.x * blockDim.x) >> 5) {
        float s = 0.f;
        for (int k = lane; k < d; k += 32) { const float v = x[r * d + k]; s = fmaf(v, v, s); }
        s = fr::warp_sum(s);
        const float nrm = sqrtf(s);
        for (int k = lane; k < d; k += 32) y[r * d + k] = __float2bfloat16(x[r * d + k] / nrm);
    }
}

}  // namespace

extern "C" int fr_f32_to_bf16(const float *x, void *y, int64_t rows, int32_t d, int32_t l2_normalise, void *stream) {
    FR_REQUIRE(rows >= 0 && d > 0, "fr_f32_to_bf16: rows=%lld d=%d", (long long)rows, d);
    if (rows == 0) return FR_OK;
    FR_REQUIRE(x && y, "fr_f32_to_bf16: null pointer");
    const int grid = fr::num_sms() * 8;
    fr::LaunchTimer _lt("f32_to_bf16_kernel", (cudaStream_t)stream);
    f32_to_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16 *>(y), rows, d,
                                                              l2_normalise);
    return fr::check_launch("fr_f32_to_bf16");
}

extern "C" int64_t fr_gemm_topk_ws_bytes(int32_t M) {
    const int n_mblk = (M + BM - 1) / BM;
    return (int64_t)std::min(n_mblk, fr::num_sms()) * BM * 4 * CAPG * 8;
}

extern "C" int fr_gemm_topk_bf16(const void *A, int32_t M, const void *B, int32_t N, int32_t K, float scale,
                                 const float *bias, const int64_t *row_ids, const int64_t *hist_ptr,
                                 const int32_t *hist_idx, int32_t topk, float *out_val, int32_t *out_idx,
                                 void *ws, int64_t ws_bytes, void *stream) {
    FR_REQUIRE(M >= 0 && N > 0 && K > 0, "fr_gemm_topk_bf16: M=%d N=%d K=%d", M, N, K);
    if (M == 0) return FR_OK;
    FR_REQUIRE(A && B && out_val && out_idx, "fr_gemm_topk_bf16: null pointer");
    FR_REQUIRE(topk >= 1 && topk <= MAXK, "fr_gemm_topk_bf16: topk=%d out of [1, %d]", topk, MAXK);
    FR_REQUIRE(K % 8 == 0, "fr_gemm_topk_bf16: K=%d must be a multiple of 8 (16-byte rows for TMA)", K);
    FR_REQUIRE((((uintptr_t)A | (uintptr_t)B) & 15) == 0, "fr_gemm_topk_bf16: operands must be 16-byte aligned");
    FR_REQUIRE((row_ids == nullptr) == (hist_ptr == nullptr) && (row_ids == nullptr) == (hist_idx == nullptr),
               "fr_gemm_topk_bf16: row_ids / hist_ptr / hist_idx go together");
    CUtensorMap ma, mb;
    if (int rc = make_map(&ma, A, M, K, BM)) return rc;
    if (int rc = make_map(&mb, B, N, K, BN)) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) {
            fr::set_error("fr_gemm_topk_bf16: cannot reserve %d bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
            return FR_ECUDA;
        }
        attr_set = true;
    }
    Params P{M, N, K, topk, scale, bias, row_ids, hist_ptr, hist_idx, out_val, out_idx};
    const int n_mblk = (M + BM - 1) / BM;
    const int grid = std::min(n_mblk, fr::num_sms());
    static int impl = -1;
    if (impl < 0) {
        const char *e = getenv("FR_TOPK_IMPL");
        impl = e ? atoi(e) : 2;
    }
    if (impl == 2) {
        const int64_t need = (int64_t)grid * BM * 4 * CAPG * 8;
        FR_REQUIRE(ws != nullptr && ws_bytes >= need, "fr_gemm_topk_bf16: workspace of %lld bytes required (got %lld)",
                   (long long)need, (long long)ws_bytes);
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
        // The bounding sweep doubles the MMA work: worth it while the epilogue is the bottleneck (short inner
        // dimension: full-sort scoring, K = 64), not when a tile already carries many k-blocks (kNN, K = 384..4096).
        static int two_pass_env = -2;
        if (two_pass_env == -2) {
            const char *e = getenv("FR_TOPK_TWO_PASS");
            two_pass_env = e ? atoi(e) : -1;
        }
        const int two_pass = two_pass_env >= 0 ? two_pass_env : (K <= 256 ? 1 : 0);
        // bounding sweep on a subset of the column tiles, keeping >= 32 tiles (4 x 32 = NG distinct groups)
        static int bstride_env = -2;
        if (bstride_env == -2) {
            const char *e = getenv("FR_TOPK_BOUND_STRIDE");
            bstride_env = e ? atoi(e) : -1;
        }
        const int n_nblk_h = (int)((N + BN - 1) / BN);
        // measured at 200 000 x 500 000, k = 32: stride 1 / 2 / 4 / 8 = 36.8 / 32.8 / 35.7 / 45.7 ms (a looser bound sends
        // more 32-score chunks down the divergent candidate path); at 45 000 columns stride 1 is best
        const int bstride = std::max(1, std::min(bstride_env > 0 ? bstride_env : (n_nblk_h >= 1024 ? 2 : 1), n_nblk_h / 32));
        Params2 P2{P, reinterpret_cast<float *>(ws),
                   reinterpret_cast<int *>(reinterpret_cast<float *>(ws) + (size_t)grid * BM * 4 * CAPG), two_pass, bstride};
        fr::LaunchTimer _lt2("gemm_topk_kernel_v2", (cudaStream_t)stream);
        if (bias != nullptr || scale != 1.f)
            gemm_topk_kernel_v2<true><<<grid, THREADS2, SMEM2, (cudaStream_t)stream>>>(ma, mb, P2);
        else
            gemm_topk_kernel_v2<false><<<grid, THREADS2, SMEM2, (cudaStream_t)stream>>>(ma, mb, P2);
        return fr::check_launch("fr_gemm_topk_bf16(v2)");
    }
    fr::LaunchTimer _lt("gemm_topk_kernel", (cudaStream_t)stream);
    gemm_topk_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, P);
    return fr::check_launch("fr_gemm_topk_bf16");
}

// ------------------------------------------------------------------------------- fp32 re-scoring
// The bf16 pass selects kc >= k candidates per row; this pass re-scores them exactly in fp32 from the
// fp32 tables and keeps the best k (descending, ties to the lower column) -- the order an fp32
// `scores.topk(k)` gives except where two fp32 scores tie.  One warp per row.
namespace {
__global__ void __launch_bounds__(256)
rescore_topk_kernel(const float *__restrict__ A, const int64_t *__restrict__ a_rows, const float *__restrict__ B,
                    int d, float scale, const float *__restrict__ bias, int metric,
                    const int32_t *__restrict__ cand, int kc, int M, int k, float *__restrict__ out_val,
                    int64_t *__restrict__ out_idx) {
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

        v[t] = -INFINITY;
        if (ci[t] >= 0) {
            const float *b = B + (size_t)ci[t] * d;
            float s = 0.f;
            if (metric == 0) {
                for (int q = 0; q < d; q += 4) {
                    const float4 x = fr::ldg_f4(a + q), y = fr::ldg_f4(b + q);
                    s = fmaf(x.x, y.x, s);
                    s = fmaf(x.y, y.y, s);
                    s = fmaf(x.z, y.z, s);
                    s = fmaf(x.w, y.w, s);
                }
                v[t] = s * scale + (bias ? __ldg(bias + ci[t]) : 0.f);
            } else {  // negative squared Euclidean distance: no cancellation, unlike x.c - |c|^2/2
                for (int q = 0; q < d; q += 4) {
                    const float4 x = fr::ldg_f4(a + q), y = fr::ldg_f4(b + q);
                    const float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
                    s = fmaf(e0, e0, s);
                    s = fmaf(e1, e1, s);
                    s = fmaf(e2, e2, s);
                    s = fmaf(e3, e3, s);
                }
                v[t] = -s;
            }
        }
    }
    }
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
            out_idx[(size_t)row * k + o] = bi == 0x7fffffff ? -1 : bi;
        }
        if (ci[0] == bi) ci[0] = -1;
        if (ci[1] == bi) ci[1] = -1;
    }
}
}  // namespace

extern "C" int fr_rescore_topk_f32(const float *A, const int64_t *a_rows, const float *B, int32_t d, float scale,
                                   const float *bias, int32_t metric, const int32_t *cand, int32_t kc, int32_t M,
                                   int32_t k, float *out_val, int64_t *out_idx, void *stream) {
    FR_REQUIRE(M >= 0 && d > 0 && d % 4 == 0, "fr_rescore_topk_f32: M=%d d=%d", M, d);
    if (M == 0) return FR_OK;
    FR_REQUIRE(A && B && cand && out_val && out_idx, "fr_rescore_topk_f32: null pointer");
    FR_REQUIRE(kc >= 1 && kc <= 64 && k >= 1 && k <= kc, "fr_rescore_topk_f32: k=%d kc=%d", k, kc);
    FR_REQUIRE(metric == 0 || metric == 1, "fr_rescore_topk_f32: metric=%d", metric);
    const long long blocks = ((long long)M * 32 + 255) / 256;
    fr::LaunchTimer _lt("rescore_topk_kernel", (cudaStream_t)stream);
    rescore_topk_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, a_rows, B, d, scale, bias, metric, cand,
                                                                           kc, M, k, out_val, out_idx);
    return fr::check_launch("fr_rescore_topk_f32");
}
