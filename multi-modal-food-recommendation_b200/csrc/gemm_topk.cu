// Fused  scores = A . B^T (bf16 in, fp32 accumulate)  ->  per-row top-k,  tcgen05 / TMEM / TMA, sm_100a.
//
// One kernel family for the three ranking shapes of the path:
//   full-sort      A = user_all[users] [M, 64],  B = item_all [N, 64], optional history mask, k <= 64
//   cosine kNN     A = B = row-normalised features [N, D], k = knn_k (self kept, utils.py:119)
//   centroids      A = features [I, D], B = centres [C, D], bias = -|c|^2 / 2, scale 1  (argmin |x - c|)
// The [M, N] score matrix never exists in HBM.
//
// `rank_topk_pair_kernel`: a CLUSTER OF TWO CTAs (one per SM of a TPC) owns a 256-row block of A and issues
// `tcgen05.mma.cta_group::2` (M = 256, N = 256, K = 16 per instruction): each CTA supplies its own 128 rows of A and
// HALF of the B tile (128 item rows), the pair's tensor cores read both halves, and each CTA receives the
// 128 x 256 fp32 accumulator of its rows in its own TMEM (two stages = all 512 columns).  Against one CTA per row
// block this halves the B bytes every SM pulls from L2 and the B bytes its MMAs read from shared memory -- the two
// feeds that capped the single-CTA kernel (every CTA re-streamed all of B; 128 x 256 x 16 MMAs read 96 B/clk of
// operands while TMA wrote another 96 B/clk into the same shared memory).
//   warp 0      TMA producer (both CTAs): SWIZZLE_128B K-major boxes, `cp.async.bulk.tensor.2d.cta_group::2`
//               completing on the LEADER's mbarrier; for K <= 128 the A block is loaded once per row block
//               (double-buffered) and only B is streamed
//   warp 1      TMEM allocator (both CTAs) + single-lane MMA issuer (leader CTA only); `tcgen05.commit ...
//               multicast::cluster` releases the operand slots / publishes the accumulator in BOTH CTAs
//   warps 2-17  epilogue: the four warps that may read a TMEM lane group split each 256-column tile into 64-column
//               quarters; a thread owns (row, quarter): one `tcgen05.ld.32x32b.x64`, a 3-input max tree, one
//               compare per 64 scores; candidate lists live in an L2-resident workspace
// Two sweeps per row block when the inner dimension is short (K <= 256: the tensor pipe outruns the epilogue):
//   pass 0 (bounding, branch-free): per (row, quarter) the running maximum of NG column groups -> shared memory;
//          the (k + h)-th largest group maximum T (h = the row's history length: a masked column may hold a group's
//          maximum) is a lower bound of the k-th eligible score.  May visit only every `bstride`-th tile.
//   pass 1 (collection): streaming top-k started from T instead of -inf, so ~k scores per row take the candidate
//          path; a list that could overflow is pruned by its whole warp (radix select of the k-th key by ballots).
// At the end of a row block the four quarter lists of a row are merged, selected and emitted in descending order
// (ties -> lower column).
#include <algorithm>

#include "rank_common.cuh"

namespace {
using namespace rk;

constexpr int EW = 16;                       // epilogue warps (four per TMEM lane group)
constexpr int THREADS = 64 + EW * 32;
constexpr int NG = 128;                      // column groups per row for the bounding pass
constexpr int CAP = 128;                     // slots per (row, quarter) list; one visit appends <= 64, prune when > CAP - 64
constexpr int BH_BYTES = (BN / 2) * BK * 2;  // this CTA's half of a B tile (128 item rows x 64 k)
constexpr int OWN_CAP = 64;                   // per (row, quarter): history columns that fall into the quarter (scratch list)
constexpr int MAX_STAGES = 8;
constexpr int SN = 128;                      // accumulator stage width: a 256-column tile is two N = 128 MMAs
constexpr int NACC = 4;                      // accumulator stages (4 x 128 = all 512 TMEM columns)
constexpr int SMEM_LIMIT = 227 * 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// one lane of a converged warp (the role loops stay warp-uniform, so tile counters, shared-memory addresses and
// descriptors live in uniform registers and feed UTMALDG / UTCHMMA without a per-instruction R2UR + BRA.U.ANY loop:
// with a single-thread issuer the loop was ~160 dependent SASS instructions per tile and paced the whole kernel)
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on an mbarrier of (possibly) the other CTA of the pair.  Default semantics (release at CTA scope): the TMEM
// reads it orders are already complete (`tcgen05.wait::ld` + `tcgen05.fence::before_thread_sync`); a
// `.release.cluster` arrive compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive, which was a third of all
// stall samples of the epilogue warps (profiles/r2_rank_pair_probe_*).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes complete on an mbarrier of (possibly) the other CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *map, uint32_t bar_cluster_addr, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrives (once) on the barrier at this shared-memory offset in BOTH CTAs when the pair's earlier MMAs have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// asynchronous 32-lane x 32-column TMEM load: the registers are valid only after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, float (&r)[32]) {
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
}
// The wait names the loaded registers as in/out operands, so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_wait_ld(float (&a)[32], float (&b)[32]) {
    uint32_t *u = reinterpret_cast<uint32_t *>(a), *w = reinterpret_cast<uint32_t *>(b);
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]), "+r"(u[8]),
                   "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]), "+r"(u[16]),
                   "+r"(u[17]), "+r"(u[18]), "+r"(u[19]), "+r"(u[20]), "+r"(u[21]), "+r"(u[22]), "+r"(u[23]), "+r"(u[24]),
                   "+r"(u[25]), "+r"(u[26]), "+r"(u[27]), "+r"(u[28]), "+r"(u[29]), "+r"(u[30]), "+r"(u[31]),
                   "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]),
                   "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]),
                   "+r"(w[17]), "+r"(w[18]), "+r"(w[19]), "+r"(w[20]), "+r"(w[21]), "+r"(w[22]), "+r"(w[23]), "+r"(w[24]),
                   "+r"(w[25]), "+r"(w[26]), "+r"(w[27]), "+r"(w[28]), "+r"(w[29]), "+r"(w[30]), "+r"(w[31])
                 :
                 : "memory");
}

// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t kIdescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

// All 32 lanes call this for the same list.  Keeps the k best of `cnt` (> k) entries, returns the k-th value.
__device__ __noinline__ float warp_prune(float *bv, int *bi, int cnt, int k, int lane) {
    constexpr int T = CAP / 32;
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

struct PairParams {
    Params p;
    float *lv;   // [n_cta][BM][4][CAP] candidate values
    int *li;     //                     candidate columns
    int *own;    // [n_cta][BM][4][OWN_CAP] each thread's own history columns (the ones inside its quarter of the tiles)
    int two_pass;
    int bstride;   // the bounding sweep visits every bstride-th column tile (a subset still bounds from below)
    int n_stages;  // depth of the operand ring
    float thr0;    // initial threshold: -inf (debug: +inf measures the pipeline with no candidate ever taken)
    uint32_t off_ring, off_cnt, off_thr, off_gkey, off_bars;   // byte offsets inside the 1024-aligned dynamic shared memory
};

// ---- epilogue helpers.  A thread owns (row, 64-column quarter of every tile) as two 32-column register halves
// (576 threads put five warps on one SM sub-partition: <= 96 registers per thread).

// optional affine map and tail handling of one half: columns >= N forced to -inf (they are zero-filled by TMA)
template <bool AFFINE>
__device__ __forceinline__ void fix_scores(float (&r)[32], const Params &P, int col0, int valid) {
    if (AFFINE) {
        if (P.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = fmaf(r[j], P.scale, j < valid ? __ldg(P.bias + col0 + j) : 0.f);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] *= P.scale;
        }
    }
    if (valid < 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j >= valid) r[j] = -INFINITY;
    }
}

__device__ __forceinline__ float max8(const float (&r)[32], int g) {
    return max3(max3(r[8 * g], r[8 * g + 1], r[8 * g + 2]), max3(r[8 * g + 3], r[8 * g + 4], r[8 * g + 5]),
                fmaxf(r[8 * g + 6], r[8 * g + 7]));
}
__device__ __forceinline__ float max32(const float (&r)[32]) {
    return fmaxf(max3(max8(r, 0), max8(r, 1), max8(r, 2)), max8(r, 3));
}

// History mask as a CURSOR over the row's sorted history slice: `h0` is the next history column not yet passed,
// `h1` the one after it (loaded one event ahead, so the global-load latency is never waited for).  The common case is
// one compare per 64 scores; a history column inside the current 64 columns poisons its register (-inf), so masked
// scores never reach the bounds, the thresholds or the candidate lists (and no candidate needs a search).
struct HistCursor {
    const int32_t *base;
    int n, pos, h0, h1;
    __device__ __forceinline__ void reset() {
        pos = 0;
        h0 = n > 0 ? __ldcg(base) : 0x7fffffff;      // (L2-only loads: the list may have been written by this kernel)
        h1 = n > 1 ? __ldcg(base + 1) : 0x7fffffff;
    }
    __device__ __forceinline__ void advance() {
        h0 = h1;
        ++pos;
        h1 = pos + 1 < n ? __ldcg(base + pos + 1) : 0x7fffffff;
    }
};
// r[d] = -inf for a runtime d as a compare-and-select chain (a `switch` compiles to a chain of BRANCHES here, not to a
// jump table: measured 8.1 -> 9.8 ms on the masked C4 slice)
__device__ __forceinline__ void poison(float (&r)[32], int d) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = j == d ? -INFINITY : r[j];
}
// all history columns < col0 + 64 are consumed; those inside [col0, col0 + 64) are poisoned
__device__ __forceinline__ void apply_history(HistCursor &hc, float (&r0)[32], float (&r1)[32], int col0) {
    while (hc.h0 < col0 + 64) {
        const int d = hc.h0 - col0;
        if (d >= 32) poison(r1, d - 32);
        else if (d >= 0) poison(r0, d);
        hc.advance();
    }
}

// rare path: append every score above the threshold (8-column groups without one are skipped)
// (`li_off`: distance in words from the value lists to the column lists -- one live pointer instead of two)
__device__ __forceinline__ void collect(const float (&r)[32], int col0, float thr, float *lv, long long li_off, int &cnt) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (max8(r, g) > thr) {
#pragma unroll
            for (int j = 8 * g; j < 8 * g + 8; ++j) {
                if (r[j] > thr) {
                    __stcg(lv + cnt, r[j]);
                    __stcg(reinterpret_cast<int *>(lv) + li_off + cnt, col0 + j);
                    ++cnt;
                }
            }
        }
    }
}

// mbarrier wait on a precomputed shared-memory address: the fast path is one try_wait + branch; the bounded spin
// (rank_common.cuh: a protocol bug must trap, not hang) stays out of line
__device__ __forceinline__ bool mbar_try_a(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_spin_a(uint32_t addr, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try_a(addr, parity)) return;
        if ((spins & 255u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) {
                printf("foodrec_b200 rank_topk_pair: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    if (!mbar_try_a(addr, parity)) mbar_spin_a(addr, parity);
}

// One visit of a tile by an epilogue warp: wait for the MMAs of its accumulator stage, load both 32-column halves of its
// quarter with one wait, and hand the stage back to the MMA issuer at once (the scores are in registers).  Tiles are
// visited in (even, odd) pairs so the stage, its barriers and its TMEM address are compile-time offsets of
// `tf` (shared address of tfull[q & 1]), `te` (cluster address of the leader's tempty[q & 1]) and `tl`.
template <int ODD>
__device__ __forceinline__ void visit_tile(uint32_t tf, uint32_t te, uint32_t tl, uint32_t ph, float (&r0)[32], float (&r1)[32]) {
    mbar_wait_a(tf + 16 * ODD, ph);
    tc_fence_after();
    tmem_ld32_async(tl + 2 * SN * ODD, r0);
    tmem_ld32_async(tl + 2 * SN * ODD + 32, r1);
    tmem_wait_ld(r0, r1);
    tc_fence_before();
    __syncwarp();
    if (elect_one()) mbar_arrive_cluster(te + 16 * ODD);
}

template <bool AFFINE, bool ARES>   // AFFINE: scores are scale * acc + bias[col];  ARES: the A block stays resident (K <= 128)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
rank_topk_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, PairParams PP) {
    const Params &P = PP.p;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // same padding in both CTAs of the pair
    constexpr int STG_A = ARES ? 0 : A_BYTES;
    constexpr int STG_BYTES = STG_A + BH_BYTES;
    uint8_t *ares = smem;                                                    // [2][n_kblk] A blocks (ARES)
    uint8_t *ring = smem + PP.off_ring;                                      // n_stages x ([A] | B half)
    int *cnt_s = reinterpret_cast<int *>(smem + PP.off_cnt);                 // [BM][4]
    uint32_t *thr_key = reinterpret_cast<uint32_t *>(smem + PP.off_thr);     // [BM] best known k-th key per row
    uint32_t *gkey = reinterpret_cast<uint32_t *>(smem + PP.off_gkey);       // [NG][BM] group maxima (pass 0)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + PP.off_bars);
    uint64_t *full = bars, *empty = bars + MAX_STAGES, *tfull = bars + 2 * MAX_STAGES, *tempty = tfull + NACC;
    uint64_t *afull = tempty + NACC, *aempty = afull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(aempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_sblk = (P.M + 2 * BM - 1) / (2 * BM), n_nblk = (P.N + BN - 1) / BN, n_kblk = (P.K + BK - 1) / BK;
    const int n_pairs = (int)(gridDim.x >> 1), pair = (int)(blockIdx.x >> 1);
    const int n_pass = PP.two_pass ? 2 : 1;
    const int bstride = PP.bstride, S = PP.n_stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < NACC; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, EW); }   // 8 warps x 2 CTAs per stage
        for (int s = 0; s < 2; ++s) { mbar_init(afull + s, 1); mbar_init(aempty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < BM) thr_key[threadIdx.x] = 0u;
    if (PP.two_pass)
        for (int i = threadIdx.x; i < NG * BM; i += THREADS) gkey[i] = 0u;
    if (warp == 1) {   // both CTAs of the pair allocate all 512 TMEM columns (2 accumulator stages x 256)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();        // the peer's barriers are initialised before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) ================================
        {
            const bool elected = elect_one();
            int stage = 0, ab = 0;
            uint32_t phase = 0, abphase = 0;
            const uint32_t full0 = mapa(smem_u32(full), 0), afull0 = mapa(smem_u32(afull), 0);   // the leader's barriers
            for (int sb = pair; sb < n_sblk; sb += n_pairs) {
                const int row0 = sb * 2 * BM + (int)rank * BM;
                if (ARES) {
                    mbar_wait(aempty + ab, abphase ^ 1);
                    if (elected) {
                        if (leader) mbar_expect_tx(afull + ab, 2u * (uint32_t)n_kblk * A_BYTES);
                        for (int kb = 0; kb < n_kblk; ++kb)
                            tma_load_2d_pair(&tmA, afull0 + 8u * ab, ares + (size_t)(ab * n_kblk + kb) * A_BYTES, kb * BK, row0);
                    }
                    __syncwarp();
                    if (++ab == 2) { ab = 0; abphase ^= 1; }
                }
                for (int pass = 0; pass < n_pass; ++pass) {
                    // every sweep visits an EVEN number of tiles (an odd sweep gets one dummy visit past the last tile,
                    // which the epilogue reads as all -inf), so a sweep always starts on accumulator stages 0 / 1
                    const int step = (n_pass == 2 && pass == 0) ? bstride : 1;
                    const int n_vis = (n_nblk + step - 1) / step, n_vis2 = n_vis + (n_vis & 1);
                    for (int v = 0; v < n_vis2; ++v) {
                        const int nb = min(v * step, n_nblk - 1);
                        for (int kb = 0; kb < n_kblk; ++kb) {
                            mbar_wait(empty + stage, phase ^ 1);
                            if (elected) {
                                uint8_t *st = ring + (size_t)stage * STG_BYTES;
                                if (leader) mbar_expect_tx(full + stage, 2u * STG_BYTES);
                                if (!ARES) tma_load_2d_pair(&tmA, full0 + 8u * stage, st, kb * BK, row0);
                                tma_load_2d_pair(&tmB, full0 + 8u * stage, st + STG_A, kb * BK, nb * BN + (int)rank * (BN / 2));
                            }
                            __syncwarp();
                            if (++stage == S) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (one lane of the leader CTA) ==================
        // A 256-column tile is issued as two N = 128 MMAs per K step (B rows 0-63 / 64-127 of each CTA's half tile)
        // into two of the FOUR accumulator stages, so the epilogue warps that own one half start a quarter of a tile
        // time before the others and a stage is back 128 columns at a time.  Stage s of tile t: 2 (t & 1) + s.
        if (leader) {
            const bool elected = elect_one();
            int stage = 0, ab = 0;
            uint32_t phase = 0, abphase = 0, ti = 0;
            for (int sb = pair; sb < n_sblk; sb += n_pairs) {
                if (ARES) {
                    mbar_wait(afull + ab, abphase);
                    tc_fence_after();
                }
                for (int pass = 0; pass < n_pass; ++pass) {
                    const int step = (n_pass == 2 && pass == 0) ? bstride : 1;
                    const int n_vis = (n_nblk + step - 1) / step, n_vis2 = n_vis + (n_vis & 1);
                    for (int v = 0; v < n_vis2; ++v, ++ti) {
                        const int s0 = 2 * (int)(ti & 1u);
                        const uint32_t tph = (ti >> 1) & 1u;
                        if (n_kblk == 1) {
                            mbar_wait(full + stage, phase);
                            const uint32_t st = smem_u32(ring + (size_t)stage * STG_BYTES);
                            const uint32_t a = ARES ? smem_u32(ares + (size_t)(ab * n_kblk) * A_BYTES) : st;
                            const uint64_t ad = make_desc(a), bd = make_desc(st + STG_A);
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                mbar_wait(tempty + s0 + s, tph ^ 1);     // both CTAs' epilogues have drained this stage
                                tc_fence_after();
                                if (elected) {
#pragma unroll
                                    for (int k = 0; k < BK / UMMA_K; ++k)    // +32 B per K step inside the swizzle atom
                                        umma_f16_pair(tmem_base + (s0 + s) * SN, ad + 2 * k, bd + s * (BH_BYTES / 32) + 2 * k, kIdescPair, k != 0);
                                    umma_commit_pair(tfull + s0 + s);         // this half of the tile is complete in both CTAs
                                    if (s == 1) umma_commit_pair(empty + stage);   // frees the slot in both CTAs
                                }
                                __syncwarp();
                            }
                            if (++stage == S) { stage = 0; phase ^= 1; }
                        } else {
                            mbar_wait(tempty + s0, tph ^ 1);
                            mbar_wait(tempty + s0 + 1, tph ^ 1);
                            tc_fence_after();
                            for (int kb = 0; kb < n_kblk; ++kb) {
                                mbar_wait(full + stage, phase);
                                tc_fence_after();
                                const uint32_t st = smem_u32(ring + (size_t)stage * STG_BYTES);
                                const uint32_t a = ARES ? smem_u32(ares + (size_t)(ab * n_kblk + kb) * A_BYTES) : st;
                                const uint64_t ad = make_desc(a), bd = make_desc(st + STG_A);
                                if (elected) {
#pragma unroll
                                    for (int s = 0; s < 2; ++s)
#pragma unroll
                                        for (int k = 0; k < BK / UMMA_K; ++k)
                                            umma_f16_pair(tmem_base + (s0 + s) * SN, ad + 2 * k, bd + s * (BH_BYTES / 32) + 2 * k, kIdescPair,
                                                          (kb | k) != 0);
                                    umma_commit_pair(empty + stage);
                                    if (kb == n_kblk - 1) {
                                        umma_commit_pair(tfull + s0);
                                        umma_commit_pair(tfull + s0 + 1);
                                    }
                                }
                                __syncwarp();
                                if (++stage == S) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                }
                if (ARES) {
                    if (elected) umma_commit_pair(aempty + ab);          // the A block may be overwritten once these MMAs retire
                    __syncwarp();
                    if (++ab == 2) { ab = 0; abphase ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue (both CTAs) =====================================
        const int ew = warp - 2;
        const int lg = warp & 3;                 // TMEM lane group this warp may read
        const int q = ew >> 2;                   // column quarter of every tile handled by this warp
        const int r_in_blk = lg * 32 + lane;
        const int kk = P.topk;
        float *const lv = PP.lv + ((((size_t)blockIdx.x * BM + r_in_blk) * 4 + q) * CAP);   // this thread's candidate list
        const long long li_off = reinterpret_cast<const int *>(PP.li) - reinterpret_cast<const int *>(PP.lv);
        // quarter q = items [64 q, 64 q + 64) of a tile = columns 64 (q >> 1) .. of accumulator stage 2 (t & 1) + (q & 1);
        // the MMA issuer waits on the LEADER's `tempty` barriers
        const uint32_t tf = smem_u32(tfull + (q & 1));
        const uint32_t te = mapa(smem_u32(tempty + (q & 1)), 0);
        const uint32_t tl = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(q & 1) * SN + (uint32_t)(q >> 1) * 64;
        const uint32_t thr_a = smem_u32(thr_key + r_in_blk);
        uint32_t ph = 0;                         // phase of the stage barriers: flips after every (even, odd) pair of visits
        for (int sb = pair; sb < n_sblk; sb += n_pairs) {
            const int row_blk0 = sb * 2 * BM + (int)rank * BM;
            const int row = row_blk0 + r_in_blk;
            float thr = PP.thr0;
            int cnt = 0;
            HistCursor hc{P.hist_idx, 0, 0, 0x7fffffff, 0x7fffffff};
            if (P.row_ids != nullptr && row < P.M) {
                const long long id = P.row_ids[row];
                const long long hlo = P.hist_ptr[id];
                hc.base = P.hist_idx + hlo;
                hc.n = (int)(P.hist_ptr[id + 1] - hlo);
                // Each of the row's four quarter threads would walk the WHOLE history in every sweep, leaving the hot
                // path once per entry (~100 issue slots each); a thread first copies the columns inside its own quarter
                // ((col / 64) % 4 == q) to a scratch list and walks only those.  A list that does not fit keeps the
                // full walk.
                if (hc.n > 0) {
                    int *own = PP.own + ((((size_t)blockIdx.x * BM + r_in_blk) * 4 + q) * OWN_CAP);
                    int n_own = 0;
                    for (int t = 0; t < hc.n; ++t) {
                        const int c = __ldg(hc.base + t);
                        if (((c >> 6) & 3) == q) {
                            if (n_own < OWN_CAP) __stcg(own + n_own, c);
                            ++n_own;
                        }
                    }
                    if (n_own <= OWN_CAP) {
                        hc.base = own;
                        hc.n = n_own;
                    }
                }
            }
            __syncwarp();
            // tails / affine map / history of one visited quarter (`checked`: the last pair of a sweep, which may hold
            // the ragged tile and the dummy visit)
            auto fix = [&](float (&r0)[32], float (&r1)[32], int col0, bool checked) {
                if (AFFINE || checked) {
                    const int valid = P.N - col0;
                    if (AFFINE || valid < 64) {
                        fix_scores<AFFINE>(r0, P, col0, min(32, valid));
                        fix_scores<AFFINE>(r1, P, col0 + 32, min(32, valid - 32));
                    }
                }
                if (hc.h0 < col0 + 64) apply_history(hc, r0, r1, col0);
            };
            if (n_pass == 2) {
                // ---------------- pass 0: group maxima -> lower bound of the k-th eligible score (masked columns are
                // poisoned before the maxima, so the k-th largest group maximum bounds the k-th ELIGIBLE score)
                // group id of (visited tile i, quarter q) = floor((4 i + q) NG / (4 n_vis)), advanced incrementally:
                // num = (4 i + q) NG - gid * den stays in [0, den)
                hc.reset();
                float gmax = -INFINITY;
                const int n_vis = (n_nblk + bstride - 1) / bstride;
                const long long den = 4LL * n_vis;
                int gid = (int)(((long long)q * NG) / den);
                long long num = (long long)q * NG - (long long)gid * den;
                int gcur = gid;
                auto bound_tile = [&](float (&r0)[32], float (&r1)[32], int col0, bool checked) {
                    fix(r0, r1, col0, checked);
                    if (gid != gcur) {
                        if (gmax > -INFINITY) atomicMax(gkey + gcur * BM + r_in_blk, order_key(gmax));
                        gmax = -INFINITY;
                        gcur = gid;
                    }
                    gmax = max3(gmax, max32(r0), max32(r1));
                    num += 4LL * NG;                      // next visited tile: (4 (i + 1) + q) NG
                    while (num >= den) { num -= den; ++gid; }
                };
                for (int v = 0; v < n_vis; v += 2) {
                    const bool checked = v + 2 >= n_vis;
                    float r0[32], r1[32];
                    visit_tile<0>(tf, te, tl, ph, r0, r1);
                    bound_tile(r0, r1, v * bstride * BN + q * 64, checked);
                    visit_tile<1>(tf, te, tl, ph, r0, r1);
                    bound_tile(r0, r1, (v + 1) * bstride * BN + q * 64, checked);
                    ph ^= 1u;
                }
                if (gmax > -INFINITY && gcur < NG) atomicMax(gkey + gcur * BM + r_in_blk, order_key(gmax));
                __threadfence_block();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "r"(128) : "memory");
                // warp q of the lane group bounds rows q*8 .. q*8+7: the k-th largest of NG group maxima
                for (int rr = 0; rr < 8; ++rr) {
                    const int rib = lg * 32 + q * 8 + rr;
                    uint32_t key[NG / 32];
#pragma unroll
                    for (int t = 0; t < NG / 32; ++t) {
                        key[t] = gkey[(lane + 32 * t) * BM + rib];
                        gkey[(lane + 32 * t) * BM + rib] = 0u;       // ready for the next row block
                    }
                    uint32_t Tk = 0;
                    for (int bit = 31; bit >= 0; --bit) {
                        const uint32_t c = Tk | (1u << bit);
                        int n = 0;
#pragma unroll
                        for (int t = 0; t < NG / 32; ++t) n += __popc(__ballot_sync(0xffffffffu, key[t] >= c));
                        if (n >= kk) Tk = c;
                    }
                    // collection accepts `score > thr`: publish the key just below T so that ties with T pass
                    // (key 0 = "no group", order_key(-inf) = 0x007fffff: fewer than k finite maxima give thr = -inf)
                    if (lane == 0) thr_key[rib] = Tk > order_key(-INFINITY) ? Tk - 1u : 0u;
                }
                __threadfence_block();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "r"(128) : "memory");
            }
            // -------------------- collection sweep
            hc.reset();
            // hot path per tile: two 3-input max trees, one compare, one vote.  A warp leaves it only when one of its
            // rows holds a candidate; then just those lanes walk the 8-column groups that hold one (-inf never passes)
            auto collect_tile = [&](float (&r0)[32], float (&r1)[32], int col0, bool checked) {
                fix(r0, r1, col0, checked);
                const float m0 = max32(r0), m1 = max32(r1);
                if (__any_sync(0xffffffffu, fmaxf(m0, m1) > thr)) {
                    if (m0 > thr) collect(r0, col0, thr, lv, li_off, cnt);
                    if (m1 > thr) collect(r1, col0 + 32, thr, lv, li_off, cnt);
                    // lists that could overflow on the next visit are pruned by the whole warp
                    uint32_t need = __ballot_sync(0xffffffffu, cnt > CAP - 64);
                    if (need) {
                        __threadfence_block();
                        __syncwarp();
                        float *const l0 = PP.lv + ((((size_t)blockIdx.x * BM + lg * 32) * 4 + q) * CAP);   // lane 0's list
                        do {
                            const int src = __ffs(need) - 1;
                            need &= need - 1;
                            const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                            const float t_new = warp_prune(l0 + (size_t)src * 4 * CAP, reinterpret_cast<int *>(l0) + li_off + (size_t)src * 4 * CAP,
                                                           c_src, kk, lane);
                            if (lane == src) {
                                thr = fmaxf(thr, t_new);
                                cnt = kk;
                                atomicMax(thr_key + r_in_blk, order_key(t_new));   // a bound every quarter of the row may use
                            }
                        } while (need);
                    }
                }
            };
            const int col_last = ((n_nblk - 1) & ~1) * BN + q * 64;      // first tile of the last (checked) pair
            for (int col0 = q * 64; col0 <= col_last; col0 += 2 * BN) {
                const uint32_t shared_key = lds32(thr_a);
                if (shared_key != 0u) thr = fmaxf(thr, key_value(shared_key));
                float r0[32], r1[32];
                visit_tile<0>(tf, te, tl, ph, r0, r1);
                collect_tile(r0, r1, col0, col0 == col_last);
                visit_tile<1>(tf, te, tl, ph, r0, r1);
                collect_tile(r0, r1, col0 + BN, col0 == col_last);
                ph ^= 1u;
            }
            // ---- end of the row block: trim own lists to k, publish counts, merge the four quarters per row
            {
                uint32_t need = __ballot_sync(0xffffffffu, cnt > kk);
                if (need) { __threadfence_block(); __syncwarp(); }
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                    const size_t l_src = (((size_t)blockIdx.x * BM + lg * 32 + src) * 4 + q) * CAP;
                    warp_prune(PP.lv + l_src, PP.li + l_src, c_src, kk, lane);
                    if (lane == src) cnt = kk;
                }
            }
            cnt_s[r_in_blk * 4 + q] = cnt;
            __threadfence_block();
            asm volatile("bar.sync %0, %1;" ::"r"(1 + lg), "r"(128) : "memory");
            for (int rr = 0; rr < 8; ++rr) {
                const int rib = lg * 32 + q * 8 + rr;
                const int orow = row_blk0 + rib;
                if (orow >= P.M) break;
                const int c0 = cnt_s[rib * 4], c1 = cnt_s[rib * 4 + 1], c2 = cnt_s[rib * 4 + 2], c3 = cnt_s[rib * 4 + 3];
                const int total = c0 + c1 + c2 + c3;       // <= 4 kk <= 256
                const size_t rbase = ((size_t)blockIdx.x * BM + rib) * 4 * CAP;
                float v[8];
                int ix[8];
                uint32_t key[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    int e = lane + 32 * t;
                    const bool ok = e < total;
                    int qq = 0;
                    if (e >= c0) { e -= c0; qq = 1; if (e >= c1) { e -= c1; qq = 2; if (e >= c2) { e -= c2; qq = 3; } } }
                    v[t] = ok ? __ldcg(PP.lv + rbase + (size_t)qq * CAP + e) : 0.f;
                    ix[t] = ok ? __ldcg(PP.li + rbase + (size_t)qq * CAP + e) : -1;
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
    cluster_sync();        // neither CTA frees its TMEM while the pair's MMAs or the peer's arrivals may still touch it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------ host
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

struct PairLaunch {
    int grid, stages, smem;
    bool ares;
    uint32_t off_ring, off_cnt, off_thr, off_gkey, off_bars;
};

PairLaunch plan_pair(int M, int K, bool two_pass) {
    PairLaunch L;
    const int n_sblk = (M + 2 * BM - 1) / (2 * BM), n_kblk = (K + BK - 1) / BK;
    L.grid = 2 * std::max(1, std::min(n_sblk, fr::num_sms() / 2));
    L.ares = n_kblk <= 2;
    const int stage_bytes = (L.ares ? 0 : A_BYTES) + BH_BYTES;
    const uint32_t ares_bytes = L.ares ? 2u * n_kblk * A_BYTES : 0u;
    const uint32_t tail = BM * 4 * 4 + BM * 4 + (two_pass ? NG * BM * 4 : 0) + 256;
    L.stages = std::max(2, std::min<int>(MAX_STAGES, (SMEM_LIMIT - 1024 - (int)ares_bytes - (int)tail) / stage_bytes));
    L.off_ring = ares_bytes;
    L.off_cnt = L.off_ring + (uint32_t)L.stages * stage_bytes;
    L.off_thr = L.off_cnt + BM * 4 * 4;
    L.off_gkey = L.off_thr + BM * 4;
    L.off_bars = L.off_gkey + (two_pass ? NG * BM * 4 : 0);
    L.smem = 1024 + (int)L.off_bars + 256;
    return L;
}

template <bool AFFINE, bool ARES>
int launch_pair(const CUtensorMap &ma, const CUtensorMap &mb, const PairParams &PP, const PairLaunch &L, cudaStream_t st) {
    static int attr_smem = 0;
    if (attr_smem < L.smem) {
        cudaError_t e = cudaFuncSetAttribute(rank_topk_pair_kernel<AFFINE, ARES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem);
        if (e != cudaSuccess) {
            fr::set_error("fr_gemm_topk_bf16: cannot reserve %d bytes of shared memory: %s", L.smem, cudaGetErrorString(e));
            return FR_ECUDA;
        }
        attr_smem = L.smem;
    }
    fr::LaunchTimer _lt("rank_topk_pair_kernel", st);
    rank_topk_pair_kernel<AFFINE, ARES><<<L.grid, THREADS, L.smem, st>>>(ma, mb, PP);
    return fr::check_launch("fr_gemm_topk_bf16");
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
    return (int64_t)plan_pair(M, 64, true).grid * BM * 4 * (CAP * 8 + OWN_CAP * 4);
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
    if (int rc = make_map(&mb, B, N, K, BN / 2)) return rc;       // each CTA of the pair loads half of a B tile
    Params P{M, N, K, topk, scale, bias, row_ids, hist_ptr, hist_idx, out_val, out_idx};
    // The bounding sweep doubles the MMA work: worth it while the epilogue is the bottleneck (short inner
    // dimension: full-sort scoring, K = 64), not when a tile already carries many k-blocks (kNN, K = 384..4096).
    static int two_pass_env = -2, bstride_env = -2;
    if (two_pass_env == -2) {
        const char *e = getenv("FR_TOPK_TWO_PASS");
        two_pass_env = e ? atoi(e) : -1;
        e = getenv("FR_TOPK_BOUND_STRIDE");
        bstride_env = e ? atoi(e) : -1;
    }
    const int two_pass = two_pass_env >= 0 ? two_pass_env : (K <= 256 ? 1 : 0);
    // bounding sweep on a subset of the column tiles, keeping >= 32 tiles (4 x 32 = NG distinct groups)
    const int n_nblk_h = (int)((N + BN - 1) / BN);
    const int bstride = std::max(1, std::min(bstride_env > 0 ? bstride_env : (n_nblk_h >= 1024 ? 2 : 1), n_nblk_h / 32));
    cudaStream_t st = (cudaStream_t)stream;
    const PairLaunch L = plan_pair(M, K, two_pass != 0);
    const int64_t need = (int64_t)L.grid * BM * 4 * (CAP * 8 + OWN_CAP * 4);
    FR_REQUIRE(ws != nullptr && ws_bytes >= need, "fr_gemm_topk_bf16: workspace of %lld bytes required (got %lld)",
               (long long)need, (long long)ws_bytes);
    PairParams PP{P, reinterpret_cast<float *>(ws), reinterpret_cast<int *>(reinterpret_cast<float *>(ws) + (size_t)L.grid * BM * 4 * CAP),
                  reinterpret_cast<int *>(ws) + (size_t)L.grid * BM * 4 * CAP * 2,
                  two_pass, bstride, L.stages, getenv("FR_TOPK_PROBE") ? INFINITY : -INFINITY, L.off_ring, L.off_cnt, L.off_thr, L.off_gkey, L.off_bars};
    const bool affine = bias != nullptr || scale != 1.f;
    if (affine) return L.ares ? launch_pair<true, true>(ma, mb, PP, L, st) : launch_pair<true, false>(ma, mb, PP, L, st);
    return L.ares ? launch_pair<false, true>(ma, mb, PP, L, st) : launch_pair<false, false>(ma, mb, PP, L, st);
}
