// Normalised-adjacency propagation  Y = alpha * (S . X) + beta * Z  (+ bias, tanh)  on sm_100a.
//
// Gather-reduce bound by the row gathers (L1/L2 -> SM while the tables fit the 126 MB L2, HBM beyond).  Layout: S in CSR
// (int32 col, fp32 val), X/Z/Y fp32 row-major [N, D].  Work unit = a SEGMENT (<= seg_len nonzeros of one row).  A GROUP
// of 8 lanes owns a segment and every lane keeps D / 32 128-bit column slices of the output row for the whole segment
// (D = 64: 8 lanes x 2 float4, four rows per warp): no cross-lane reduction; col / val are fetched 8 at a time per
// group with one coalesced load and broadcast by width-8 shuffles; eight 128-bit row gathers in flight per lane.
// Rows longer than a segment write per-segment partials; the last segment to arrive (atomic ticket) folds them in
// fixed order, so the result is bit-reproducible run to run.  Algorithmic bytes per launch:
//   8*nnz + 4*(R+1) + 4*D*C + 4*D*R   (+ 4*D*R when the fused Z stream is used).
// (Two earlier variants -- a warp per segment, and a cp.async.bulk row gather into a shared-memory ring -- were measured
// at 75 and 143 us against this kernel's 44 us on the C2 user-item graph and have been removed)
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
#ifndef FR_GROUP_MIN_BLOCKS
#define FR_GROUP_MIN_BLOCKS 4
#endif

// Optional two-segment operands: rows [0, split) come from the primary pointer, rows >= split from a second
// table (`*_adj` = second pointer - split * D, so both are indexed by the global row).  This is how the
// reference's `torch.cat((user_w, item_emb))` / `cat((item_w, side_w))` ego tables are consumed without
// materialising the concatenation.  split = INT_MAX: single table.
struct Split {
    const float *X1_adj;
    int x_split;
    const float *Z1_adj;
    int z_split;
    // Optional row-activity masks (backward of a mini-batch loss: most rows of the upstream gradient are
    // exactly zero).  x_mask[c] == 0 promises that row c of X is all zeros, so its gather is skipped;
    // y_mask[r] receives whether output row r has any nonzero, for the next layer.
    const unsigned char *x_mask;
    unsigned char *y_mask;
    // Optional push epilogue (row-partitioned multi-GPU propagation): every finished output row is also
    // stored at row `row_off + r` of up to 8 full-size tables, one per rank of the node (peer memory over
    // NVLink, this rank's own copy included) -- the all-gather of the next layer's input, fused into the
    // producing kernel so the transfer overlaps the gathers tile by tile.
    float *peer[8];
    int n_peers;
    long long row_off;
};

template <int D, int ACT, bool PUSH = false>
__device__ __forceinline__ bool epilogue_store(float4 acc, int row, int off, const float *__restrict__ Z,
                                               float alpha, float beta, const float *__restrict__ bias,
                                               float *__restrict__ Y, const Split *sp = nullptr) {
    float4 y = make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w);
    const size_t o = (size_t)row * D + off;
    if (Z != nullptr) {
        const float4 z = fr::ldg_f4(Z + o);
        y.x = fmaf(beta, z.x, y.x);
        y.y = fmaf(beta, z.y, y.y);
        y.z = fmaf(beta, z.z, y.z);
        y.w = fmaf(beta, z.w, y.w);
    }
    if (bias != nullptr) {
        const float4 b = fr::ldg_f4(bias + off);
        fr::add4(y, b);
    }
    if (ACT == 1) {
        y.x = tanhf(y.x);
        y.y = tanhf(y.y);
        y.z = tanhf(y.z);
        y.w = tanhf(y.w);
    }
    if (!PUSH || Y != nullptr) *reinterpret_cast<float4 *>(Y + o) = y;
    if (PUSH) {
        const size_t po = (size_t)(sp->row_off + row) * D + off;
#pragma unroll
        for (int q = 0; q < 8; ++q)     // static indices: the table pointers stay in the kernel-parameter bank
            if (q < sp->n_peers) *reinterpret_cast<float4 *>(sp->peer[q] + po) = y;
    }
    return (y.x != 0.f) | (y.y != 0.f) | (y.z != 0.f) | (y.w != 0.f);
}

// One GROUP of LPR lanes owns one segment and every lane keeps VPL = D/(4 LPR) float4
// column slices of the output row for the whole segment: no cross-lane reduction, a quarter of the
// per-row bookkeeping of the warp-per-segment kernel (D = 64: 8 lanes x 2 float4, four rows per warp),
// and the shuffle / address / predicate work of a step is shared by 4 nonzeros.  Groups of one warp take
// adjacent segments of the length-sorted plan, so their trip counts match.
// `w` = index of this warp among the warps of ONE propagation (the plain launch derives it from blockIdx; the grouped
// launch, which runs several independent propagations in one grid, from its block map).
template <int D, int LPR, int U, int ACT, bool SPLIT, bool MASKED, bool PUSH = false>
__device__ __forceinline__ void
spmm_group_body(const long long w, const int4 *__restrict__ seg, long long n_seg, const int4 *__restrict__ long_rows,
                const int *__restrict__ col, const float *__restrict__ val, const float *__restrict__ X,
                const float *__restrict__ Z, float alpha, float beta, const float *__restrict__ bias,
                float *__restrict__ Y, float *__restrict__ partial, int *__restrict__ counters, const Split &sp) {
    constexpr int G = 32 / LPR, VPL = D / (4 * LPR);
    static_assert(VPL >= 1 && LPR % U == 0, "bad group shape");
    if (w * G >= n_seg) return;
    const int lane = threadIdx.x & 31;
    const int g = lane / LPR, lg = lane % LPR;
    const long long sidx = w * G + g;
    int4 s = make_int4(0, 0, 0, -2);                     // .w == -2: no segment for this group
    if (sidx < n_seg) s = __ldg(seg + sidx);
    int maxlen = s.z;
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    const int *__restrict__ cp = col + s.y;
    const float *__restrict__ vp = val + s.y;
    const float *__restrict__ Xo = X + lg * 4;           // slice t of this lane starts at column 4 (lg + LPR t)
    const float *__restrict__ X1o = sp.X1_adj + lg * 4;

    float4 acc[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = 0; base < maxlen; base += LPR) {
        const int cnt = max(0, min(LPR, s.z - base));
        int c = 0;
        float v = 0.f;
        if (lg < cnt) {
            c = __ldg(cp + base + lg);
            v = __ldg(vp + base + lg);
            if (MASKED && sp.x_mask[c] == 0) c = -1;      // source row is all zeros: nothing to gather
        }
        const int lim = min(LPR, maxlen - base);
        for (int j = 0; j < lim; j += U) {
            float4 x[U][VPL];
            float vv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = j + u;
                const int cj = __shfl_sync(0xffffffffu, c, e, LPR);
                vv[u] = __shfl_sync(0xffffffffu, v, e, LPR);
                const float *xr = ((!SPLIT || cj < sp.x_split) ? Xo : X1o) + (size_t)cj * D;
                if (e < cnt && (!MASKED || cj >= 0)) {
#pragma unroll
                    for (int t = 0; t < VPL; ++t) x[u][t] = fr::ldg_f4(xr + 4 * LPR * t);
                } else {
                    vv[u] = 0.f;
#pragma unroll
                    for (int t = 0; t < VPL; ++t) x[u][t] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int t = 0; t < VPL; ++t) fr::fma4(acc[t], vv[u], x[u][t]);
        }
    }
    if (s.w == -2) return;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
    if (s.w < 0) {
        bool nz = false;
#pragma unroll
        for (int t = 0; t < VPL; ++t)
            nz |= epilogue_store<D, ACT, PUSH>(acc[t], s.x, 4 * (lg + LPR * t),
                                               (Z != nullptr && s.x >= sp.z_split) ? sp.Z1_adj : Z, alpha, beta, bias, Y, &sp);
        if (sp.y_mask != nullptr) {
            nz = __any_sync(gmask, nz);
            if (lg == 0) sp.y_mask[s.x] = nz ? 1 : 0;
        }
        return;
    }
    // ---- long row: publish this segment's partial; the last one to arrive (atomic ticket) folds the entry's partials in
    //      fixed order.  Rows of many segments fold in TWO levels (plan: child entries of ~sqrt(k) segments each, whose
    //      folded sums become the partials of a parent entry), so the serial chain of the last arriver is ~2 sqrt(k)
    //      loads instead of k: the tail of the launch no longer waits for one group walking thousands of partial rows.
    int id = s.w;
    int4 lr = __ldg(long_rows + id);                     // first_seg | first_child, n_parts, part_base, row | -(parent + 1)
    int slot = s.x;                                      // long-row segments carry their slot, not the row (the entry has it)
    for (;;) {
#pragma unroll
        for (int t = 0; t < VPL; ++t)
            __stcg(reinterpret_cast<float4 *>(partial + ((size_t)lr.z + slot) * D + 4 * (lg + LPR * t)), acc[t]);
        __threadfence();
        __syncwarp(gmask);
        int ticket = 0;
        if (lg == 0) ticket = atomicAdd(counters + id, 1);
        ticket = __shfl_sync(gmask, ticket, g * LPR);
        if (ticket != lr.y - 1) return;
        __threadfence();
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
            float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
            const float *pp = partial + (size_t)lr.z * D + 4 * (lg + LPR * t);
            int k = 0;
            for (; k + 8 <= lr.y; k += 8) {      // eight partial rows in flight; the fold order stays k = 0, 1, 2, ...
                float4 p8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) p8[u] = fr::ldcg_f4(pp + (size_t)(k + u) * D);
#pragma unroll
                for (int u = 0; u < 8; ++u) fr::add4(tot, p8[u]);
            }
            for (; k < lr.y; ++k) fr::add4(tot, fr::ldcg_f4(pp + (size_t)k * D));
            acc[t] = tot;
        }
        if (lg == 0) counters[id] = 0;
        if (lr.w >= 0) break;                            // a row's own (top-level) entry: finish below
        const int parent = -lr.w - 1;                    // child entry: its sum is partial `id - first_child` of the parent
        const int4 pr = __ldg(long_rows + parent);
        slot = id - pr.x;
        id = parent;
        lr = pr;
    }
    bool nz_long = false;
#pragma unroll
    for (int t = 0; t < VPL; ++t)
        nz_long |= epilogue_store<D, ACT, PUSH>(acc[t], lr.w, 4 * (lg + LPR * t),
                                                (Z != nullptr && lr.w >= sp.z_split) ? sp.Z1_adj : Z, alpha, beta, bias, Y, &sp);
    if (sp.y_mask != nullptr) {
        nz_long = __any_sync(gmask, nz_long);
        if (lg == 0) sp.y_mask[lr.w] = nz_long ? 1 : 0;
    }
}

template <int D, int LPR, int U, int ACT, bool SPLIT, bool MASKED, bool PUSH = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, FR_GROUP_MIN_BLOCKS)
spmm_group_kernel(const int4 *__restrict__ seg, long long n_seg, const int4 *__restrict__ long_rows,
                  const int *__restrict__ col, const float *__restrict__ val, const float *__restrict__ X,
                  const float *__restrict__ Z, float alpha, float beta, const float *__restrict__ bias,
                  float *__restrict__ Y, float *__restrict__ partial, int *__restrict__ counters, Split sp) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    spmm_group_body<D, LPR, U, ACT, SPLIT, MASKED, PUSH>(w, seg, n_seg, long_rows, col, val, X, Z, alpha, beta, bias, Y, partial,
                                                         counters, sp);
}

// Grouped launch: up to FR_SPMM_MAX_TASKS independent propagations (different graphs, different operands) in ONE grid.
// CLUSSL's three item-side graphs are each too small to fill 148 SMs and end in a tail of a few long rows; launched as
// one grid their tails overlap and the step has 6 propagation launches instead of 14.  `blk_map[b] = (task, block of
// that task)` interleaves the tasks' blocks in proportion, so every task's long-row segments (first in its plan) are
// scheduled early.  Two-segment operands are always compiled in (x_split = INT_MAX: single table).
struct GroupedTask {
    const int4 *seg;
    long long n_seg;
    const int4 *long_rows;
    const int *col;
    const float *val;
    const float *X, *X1_adj, *Z, *Z1_adj;
    float *Y, *partial;
    int *counters;
    int x_split, z_split;
    float alpha, beta;
};
struct GroupedParams {
    GroupedTask t[FR_SPMM_MAX_TASKS];
};

template <int D>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, FR_GROUP_MIN_BLOCKS)
spmm_grouped_kernel(const __grid_constant__ GroupedParams P, const int2 *__restrict__ blk_map) {
    const int2 m = __ldg(blk_map + blockIdx.x);
    const GroupedTask &T = P.t[m.x];
    Split sp;
    sp.X1_adj = T.X1_adj;
    sp.x_split = T.x_split;
    sp.Z1_adj = T.Z1_adj;
    sp.z_split = T.z_split;
    sp.x_mask = nullptr;
    sp.y_mask = nullptr;
    sp.n_peers = 0;
    sp.row_off = 0;
    const long long w = (long long)m.y * kWarpsPerBlock + (threadIdx.x >> 5);
    spmm_group_body<D, 8, 4, 0, true, false, false>(w, T.seg, T.n_seg, T.long_rows, T.col, T.val, T.X, T.Z, T.alpha, T.beta,
                                                    nullptr, T.Y, T.partial, T.counters, sp);
}

template <int D, int LPR, int U, int ACT>
int launch_group_shape(const int4 *seg, int64_t n_seg, const int4 *lrows, const int *col, const float *val,
                       const float *X, const float *Z, float alpha, float beta, const float *bias, float *Y,
                       float *partial, int *counters, cudaStream_t st, Split sp) {
    constexpr int G = 32 / LPR;
    const long long warps = (n_seg + G - 1) / G;
    const long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > 0x7fffffffLL) {
        fr::set_error("fr_spmm_csr_f32: too many segments (%lld)", (long long)n_seg);
        return FR_EUNSUPPORTED;
    }
    fr::LaunchTimer _lt("spmm_group_kernel", st);
    if (sp.n_peers > 0) {           // push epilogue: default shape, no activation, single-table operands
        if constexpr (ACT == 0 && LPR == 8 && U == 4) {
            spmm_group_kernel<D, LPR, U, 0, false, false, true><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
                seg, n_seg, lrows, col, val, X, Z, alpha, beta, bias, Y, partial, counters, sp);
            return fr::check_launch("fr_spmm_csr_f32_push");
        } else {
            fr::set_error("fr_spmm_csr_f32_push: needs act = 0 and the default kernel shape");
            return FR_EUNSUPPORTED;
        }
    }
    if (sp.x_split != 0x7fffffff)   // two-segment gather only where it is used (first layer of a forward)
        spmm_group_kernel<D, LPR, U, ACT, true, false><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
            seg, n_seg, lrows, col, val, X, Z, alpha, beta, bias, Y, partial, counters, sp);
    else if (sp.x_mask != nullptr)  // masked gather only in the backward launches that carry a row mask
        spmm_group_kernel<D, LPR, U, ACT, false, true><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
            seg, n_seg, lrows, col, val, X, Z, alpha, beta, bias, Y, partial, counters, sp);
    else
        spmm_group_kernel<D, LPR, U, ACT, false, false><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
            seg, n_seg, lrows, col, val, X, Z, alpha, beta, bias, Y, partial, counters, sp);
    return fr::check_launch("fr_spmm_csr_f32(group)");
}

template <int D>
int launch(const int4 *seg, int64_t n_seg, const int4 *lrows, const int *col, const float *val, const float *X,
           const float *Z, float alpha, float beta, const float *bias, int act, float *Y, float *partial,
           int *counters, cudaStream_t st, Split sp) {
    // 8 lanes per row, 4 gathers per lane per step: shapes 4x2, 4x4, 8x8, 16x4 measured 88, 64, 66, 64 us against 52
    if (act == 0) return launch_group_shape<D, 8, 4, 0>(seg, n_seg, lrows, col, val, X, Z, alpha, beta, bias, Y, partial, counters, st, sp);
    return launch_group_shape<D, 8, 4, 1>(seg, n_seg, lrows, col, val, X, Z, alpha, beta, bias, Y, partial, counters, st, sp);
}

struct PlanCounts {
    int64_t n_seg = 0, n_long = 0, n_part = 0;
};

// Fold shape of a row of k > 1 segments: up to kFlatFold segments fold in one level (one entry); beyond that the segments
// are dealt to `children` child entries of `per` consecutive segments (per ~ sqrt(k)) and a parent entry folds the
// children's sums.  The shape depends on k only, so a row folds in the same order in every graph that contains it.
constexpr int64_t kFlatFold = 16;
struct FoldShape {
    int64_t per, children;      // children == 0: single level
};
FoldShape fold_shape(int64_t k) {
    if (k <= kFlatFold) return {k, 0};
    int64_t per = 1;
    while (per * per < k) ++per;
    return {per, (k + per - 1) / per};
}

PlanCounts count_plan(const int32_t *rp, int32_t n_rows, int64_t SEG) {
    PlanCounts c;
    for (int32_t r = 0; r < n_rows; ++r) {
        const int64_t deg = (int64_t)rp[r + 1] - rp[r];
        const int64_t k = deg <= SEG ? 1 : (deg + SEG - 1) / SEG;
        c.n_seg += k;
        if (k > 1) {
            const FoldShape f = fold_shape(k);
            c.n_long += f.children ? f.children + 1 : 1;
            c.n_part += k + f.children;
        }
    }
    return c;
}

}  // namespace

extern "C" int fr_spmm_plan_sizes(const int32_t *row_ptr_host, int32_t n_rows, int32_t seg_len, int64_t *n_seg,
                                  int64_t *n_long, int64_t *n_part) {
    FR_REQUIRE(row_ptr_host && n_seg && n_long && n_part && n_rows >= 0, "fr_spmm_plan_sizes: bad argument");
    FR_REQUIRE(seg_len >= 8 && seg_len <= FR_SPMM_SEG, "fr_spmm_plan_sizes: seg_len=%d outside [8, %d]", seg_len, FR_SPMM_SEG);
    for (int32_t r = 0; r < n_rows; ++r)
        FR_REQUIRE(row_ptr_host[r + 1] >= row_ptr_host[r], "fr_spmm_plan_sizes: row_ptr not monotone at %d", r);
    const PlanCounts c = count_plan(row_ptr_host, n_rows, seg_len);
    FR_REQUIRE(c.n_long <= 0x7fffffffLL && c.n_part <= 0x7fffffffLL && c.n_seg <= 0x7fffffffLL,
               "fr_spmm_plan_sizes: plan too large for 32-bit ids");
    *n_seg = c.n_seg;
    *n_long = c.n_long;
    *n_part = c.n_part;
    return FR_OK;
}

// Segment order: every long-row segment first (they are the critical path: the last one also
// performs the reduction), then whole-row segments by descending length (longest-first keeps the
// tail of the launch short); ties keep row order so neighbouring rows stay neighbours.
// seg entries (int4): (row, start, len, -1) for a whole row; (slot, start, len, entry) for a segment of a long row, whose
// partial sum is slot `slot` of fold entry `entry`.
// long_rows entries (int4): a row's own entry (first segment within the row | first_child, n_parts, part_base, row >= 0); a
// child entry of a two-level row (first segment within the row, n_parts, part_base, -(parent + 1)), children contiguous
// and directly followed by their parent.
extern "C" int fr_spmm_plan_fill(const int32_t *row_ptr_host, int32_t n_rows, int32_t seg_len, int32_t *seg_host,
                                 int32_t *long_rows_host) {
    FR_REQUIRE(row_ptr_host && seg_host && n_rows >= 0, "fr_spmm_plan_fill: bad argument");
    FR_REQUIRE(seg_len >= 8 && seg_len <= FR_SPMM_SEG, "fr_spmm_plan_fill: seg_len=%d outside [8, %d]", seg_len, FR_SPMM_SEG);
    const int64_t SEG = seg_len;
    const int32_t *rp = row_ptr_host;
    int64_t s = 0, nl = 0, pb = 0;
    struct LongSeg { int32_t key, q[4]; };
    std::vector<LongSeg> lsegs;
    // Graphs whose table cannot live in the 126 MB L2 (>= 262 144 rows of 256 bytes) schedule their long-row segments
    // by RELATIVE POSITION inside the row instead of row after row (below); smaller graphs keep the row order.
    const bool by_position = n_rows >= 262144;
    for (int32_t r = 0; r < n_rows; ++r) {
        const int64_t deg = (int64_t)rp[r + 1] - rp[r];
        if (deg <= SEG) continue;
        FR_REQUIRE(long_rows_host, "fr_spmm_plan_fill: long_rows_host is null but row %d is long", r);
        const int64_t k = (deg + SEG - 1) / SEG;
        const FoldShape f = fold_shape(k);
        const int64_t entries = f.children ? f.children : 1;
        const int64_t parent = nl + entries;             // id of the parent entry (two-level rows only)
        for (int64_t c = 0; c < entries; ++c) {
            const int64_t first = c * f.per, cnt = std::min<int64_t>(f.per, k - first);
            int32_t *lr = long_rows_host + 4 * (nl + c);
            lr[0] = (int32_t)first;                      // index of the entry's first segment within its row (informative)
            lr[1] = (int32_t)cnt;
            lr[2] = (int32_t)pb;
            lr[3] = f.children ? (int32_t)(-(parent + 1)) : r;
            for (int64_t i = first; i < first + cnt; ++i, ++s) {
                LongSeg ls;
                ls.key = by_position ? (int32_t)((i << 16) / k) : 0;
                ls.q[0] = (int32_t)(i - first);          // slot of the segment's partial inside its entry
                ls.q[1] = rp[r] + (int32_t)(i * SEG);
                ls.q[2] = (int32_t)std::min<int64_t>(SEG, deg - i * SEG);
                ls.q[3] = (int32_t)(nl + c);
                lsegs.push_back(ls);
            }
            pb += cnt;
        }
        if (f.children) {
            int32_t *lr = long_rows_host + 4 * parent;
            lr[0] = (int32_t)nl;                         // first child entry
            lr[1] = (int32_t)f.children;
            lr[2] = (int32_t)pb;
            lr[3] = r;
            pb += f.children;
        }
        nl += entries + (f.children ? 1 : 0);
    }
    // The long-row segments run first.  In the HBM regime they are ordered by their relative position inside their row
    // (ties keep row order): the columns of a row are sorted, so segments at the same relative position of different
    // rows gather from the same band of X at the same time -- a row of X that several popular rows reference is fetched
    // from DRAM once and served from the L2 to the others (item rows of the C5-shaped graph: 1.395 -> 1.261 ms, A/B of
    // two builds on one box; neutral for the L2-resident C2 graphs, which keep the row order).  The fold is by slot, not
    // by arrival, so the order of execution does not change a single bit of the result.
    std::stable_sort(lsegs.begin(), lsegs.end(), [](const LongSeg &a, const LongSeg &b) { return a.key < b.key; });
    for (size_t i = 0; i < lsegs.size(); ++i)
        for (int j = 0; j < 4; ++j) seg_host[4 * i + j] = lsegs[i].q[j];
    // counting sort of the remaining rows by descending degree
    std::vector<int64_t> start((size_t)SEG + 2, 0);
    for (int32_t r = 0; r < n_rows; ++r) {
        const int64_t deg = (int64_t)rp[r + 1] - rp[r];
        if (deg <= SEG) start[SEG - deg + 1] += 1;
    }
    for (int i = 1; i <= SEG + 1; ++i) start[i] += start[i - 1];
    for (int32_t r = 0; r < n_rows; ++r) {
        const int64_t deg = (int64_t)rp[r + 1] - rp[r];
        if (deg > SEG) continue;
        int32_t *q = seg_host + 4 * (s + start[SEG - deg]++);
        q[0] = r;
        q[1] = rp[r];
        q[2] = (int32_t)deg;
        q[3] = -1;
    }
    return FR_OK;
}

static int spmm_entry(const int32_t *seg, int64_t n_seg, const int32_t *long_rows, int64_t n_long,
                      const int32_t *col_idx, const float *val, int32_t d, const float *X0, const float *X1,
                      int32_t x_split, const float *Z0, const float *Z1, int32_t z_split, float alpha,
                      float beta, const float *bias, int32_t act, float *Y, float *partial,
                      int32_t *counters, const uint8_t *x_mask, uint8_t *y_mask, void *stream,
                      float *const *peers_host = nullptr, int32_t n_peers = 0, int64_t row_off = 0) {
    FR_REQUIRE(n_seg >= 0 && n_long >= 0, "fr_spmm_csr_f32: negative extent");
    if (n_seg == 0) return FR_OK;
    FR_REQUIRE(seg && X0 && (Y || n_peers > 0), "fr_spmm_csr_f32: null seg/X/Y");
    FR_REQUIRE(n_long == 0 || (long_rows && partial && counters), "fr_spmm_csr_f32: long rows need workspace");
    FR_REQUIRE(act == 0 || act == 1, "fr_spmm_csr_f32: act must be 0 or 1");
    FR_REQUIRE((((uintptr_t)X0 | (uintptr_t)X1 | (uintptr_t)Y | (uintptr_t)Z0 | (uintptr_t)Z1 | (uintptr_t)bias |
                 (uintptr_t)partial | (uintptr_t)seg | (uintptr_t)long_rows) & 15) == 0,
               "fr_spmm_csr_f32: pointers must be 16-byte aligned");
    FR_REQUIRE(Y == nullptr || (X0 != Y && X1 != Y), "fr_spmm_csr_f32: in-place propagation is not supported");
    FR_REQUIRE(x_split >= 0 && z_split >= 0 && (X1 != nullptr || x_split == 0) && (Z1 == nullptr || Z0 != nullptr),
               "fr_spmm_csr_f32_split: bad split arguments");
    Split sp;
    sp.x_split = X1 ? x_split : 0x7fffffff;
    sp.X1_adj = X1 ? X1 - (size_t)x_split * d : X0;
    sp.z_split = Z1 ? z_split : 0x7fffffff;
    sp.Z1_adj = Z1 ? Z1 - (size_t)z_split * d : Z0;
    FR_REQUIRE(x_mask == nullptr || X1 == nullptr, "fr_spmm_csr_f32_masked: a row mask needs a single-table X");
    sp.x_mask = x_mask;
    sp.y_mask = y_mask;
    FR_REQUIRE(n_peers >= 0 && n_peers <= 8 && row_off >= 0 && (n_peers == 0 || peers_host != nullptr),
               "fr_spmm_csr_f32_push: 1..8 peer tables");
    FR_REQUIRE(n_peers == 0 || (X1 == nullptr && Z1 == nullptr && x_mask == nullptr && y_mask == nullptr && act == 0),
               "fr_spmm_csr_f32_push: single-table operands, no masks, no activation");
    sp.n_peers = n_peers;
    sp.row_off = row_off;
    for (int q = 0; q < 8; ++q) {
        sp.peer[q] = q < n_peers ? peers_host[q] : nullptr;
        FR_REQUIRE(q >= n_peers || (sp.peer[q] != nullptr && ((uintptr_t)sp.peer[q] & 15) == 0 && sp.peer[q] != X0),
                   "fr_spmm_csr_f32_push: peer tables must be non-null, 16-byte aligned and distinct from X");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int4 *sg = reinterpret_cast<const int4 *>(seg);
    const int4 *lr = reinterpret_cast<const int4 *>(long_rows);
    switch (d) {
        case 32: return launch<32>(sg, n_seg, lr, col_idx, val, X0, Z0, alpha, beta, bias, act, Y, partial, counters, st, sp);
        case 64: return launch<64>(sg, n_seg, lr, col_idx, val, X0, Z0, alpha, beta, bias, act, Y, partial, counters, st, sp);
        case 128: return launch<128>(sg, n_seg, lr, col_idx, val, X0, Z0, alpha, beta, bias, act, Y, partial, counters, st, sp);
        default:
            fr::set_error("fr_spmm_csr_f32: d=%d unsupported (32, 64, 128)", d);
            return FR_EUNSUPPORTED;
    }
}

extern "C" int fr_spmm_csr_f32_split(const int32_t *seg, int64_t n_seg, const int32_t *long_rows, int64_t n_long,
                                     const int32_t *col_idx, const float *val, int32_t d, const float *X0, const float *X1,
                                     int32_t x_split, const float *Z0, const float *Z1, int32_t z_split, float alpha,
                                     float beta, const float *bias, int32_t act, float *Y, float *partial,
                                     int32_t *counters, void *stream) {
    return spmm_entry(seg, n_seg, long_rows, n_long, col_idx, val, d, X0, X1, x_split, Z0, Z1, z_split, alpha, beta, bias, act,
                      Y, partial, counters, nullptr, nullptr, stream);
}

extern "C" int fr_spmm_csr_f32_push(const int32_t *seg, int64_t n_seg, const int32_t *long_rows, int64_t n_long,
                                    const int32_t *col_idx, const float *val, int32_t d, const float *X, const float *Z,
                                    float alpha, float beta, float *Y, float *partial, int32_t *counters,
                                    float *const *peers_host, int32_t n_peers, int64_t row_off, void *stream) {
    FR_REQUIRE(n_peers >= 1, "fr_spmm_csr_f32_push: at least one destination table");
    return spmm_entry(seg, n_seg, long_rows, n_long, col_idx, val, d, X, nullptr, 0, Z, nullptr, 0, alpha, beta, nullptr, 0, Y,
                      partial, counters, nullptr, nullptr, stream, peers_host, n_peers, row_off);
}

extern "C" int fr_spmm_csr_f32_masked(const int32_t *seg, int64_t n_seg, const int32_t *long_rows, int64_t n_long,
                                      const int32_t *col_idx, const float *val, int32_t d, const float *X, const float *Z,
                                      float alpha, float beta, float *Y, float *partial, int32_t *counters,
                                      const uint8_t *x_mask, uint8_t *y_mask, void *stream) {
    return spmm_entry(seg, n_seg, long_rows, n_long, col_idx, val, d, X, nullptr, 0, Z, nullptr, 0, alpha, beta, nullptr, 0, Y,
                      partial, counters, x_mask, y_mask, stream);
}

extern "C" int fr_spmm_csr_f32(const int32_t *seg, int64_t n_seg, const int32_t *long_rows, int64_t n_long,
                               const int32_t *col_idx, const float *val, int32_t d, const float *X, const float *Z,
                               float alpha, float beta, const float *bias, int32_t act, float *Y, float *partial,
                               int32_t *counters, void *stream) {
    return fr_spmm_csr_f32_split(seg, n_seg, long_rows, n_long, col_idx, val, d, X, nullptr, 0, Z, nullptr, 0, alpha, beta,
                                 bias, act, Y, partial, counters, stream);
}

extern "C" int64_t fr_spmm_task_blocks(int64_t n_seg) {
    return (n_seg + 4 * kWarpsPerBlock - 1) / (4 * kWarpsPerBlock);     // 4 segments per warp, kWarpsPerBlock warps per block
}

extern "C" int fr_spmm_csr_f32_grouped(const fr_spmm_task *tasks_host, int32_t n_tasks, int32_t d, const int32_t *blk_map,
                                       int64_t n_blocks, void *stream) {
    FR_REQUIRE(tasks_host && n_tasks >= 1 && n_tasks <= FR_SPMM_MAX_TASKS, "fr_spmm_csr_f32_grouped: 1..%d tasks", FR_SPMM_MAX_TASKS);
    FR_REQUIRE(blk_map && n_blocks >= 0 && n_blocks <= 0x7fffffffLL, "fr_spmm_csr_f32_grouped: bad block map");
    if (n_blocks == 0) return FR_OK;
    GroupedParams P;
    int64_t blocks = 0;
    for (int t = 0; t < FR_SPMM_MAX_TASKS; ++t) {
        GroupedTask &T = P.t[t];
        if (t >= n_tasks) {
            T = P.t[0];
            continue;
        }
        const fr_spmm_task &h = tasks_host[t];
        FR_REQUIRE(h.n_seg >= 0 && h.n_long >= 0 && h.seg && h.X0 && h.Y, "fr_spmm_csr_f32_grouped: task %d: null pointer", t);
        FR_REQUIRE(h.n_long == 0 || (h.long_rows && h.partial && h.counters), "fr_spmm_csr_f32_grouped: task %d: long rows need workspace", t);
        FR_REQUIRE((((uintptr_t)h.X0 | (uintptr_t)h.X1 | (uintptr_t)h.Y | (uintptr_t)h.Z0 | (uintptr_t)h.Z1 | (uintptr_t)h.partial |
                     (uintptr_t)h.seg | (uintptr_t)h.long_rows) & 15) == 0, "fr_spmm_csr_f32_grouped: task %d: pointers must be 16-byte aligned", t);
        FR_REQUIRE(h.X0 != h.Y && h.X1 != h.Y, "fr_spmm_csr_f32_grouped: task %d: in-place propagation is not supported", t);
        FR_REQUIRE(h.x_split >= 0 && h.z_split >= 0 && (h.X1 != nullptr || h.x_split == 0) && (h.Z1 == nullptr || h.Z0 != nullptr),
                   "fr_spmm_csr_f32_grouped: task %d: bad split arguments", t);
        T.seg = reinterpret_cast<const int4 *>(h.seg);
        T.n_seg = h.n_seg;
        T.long_rows = reinterpret_cast<const int4 *>(h.long_rows);
        T.col = h.col_idx;
        T.val = h.val;
        T.X = h.X0;
        T.x_split = h.X1 ? h.x_split : 0x7fffffff;
        T.X1_adj = h.X1 ? h.X1 - (size_t)h.x_split * d : h.X0;
        T.Z = h.Z0;
        T.z_split = h.Z1 ? h.z_split : 0x7fffffff;
        T.Z1_adj = h.Z1 ? h.Z1 - (size_t)h.z_split * d : h.Z0;
        T.Y = h.Y;
        T.partial = h.partial;
        T.counters = h.counters;
        T.alpha = h.alpha;
        T.beta = h.beta;
        blocks += fr_spmm_task_blocks(h.n_seg);
    }
    FR_REQUIRE(blocks == n_blocks, "fr_spmm_csr_f32_grouped: block map has %lld entries, the tasks need %lld", (long long)n_blocks,
               (long long)blocks);
    cudaStream_t st = (cudaStream_t)stream;
    const int2 *bm = reinterpret_cast<const int2 *>(blk_map);
    fr::LaunchTimer _lt("spmm_grouped_kernel", st);
    switch (d) {
        case 32: spmm_grouped_kernel<32><<<(unsigned)n_blocks, kWarpsPerBlock * 32, 0, st>>>(P, bm); break;
        case 64: spmm_grouped_kernel<64><<<(unsigned)n_blocks, kWarpsPerBlock * 32, 0, st>>>(P, bm); break;
        case 128: spmm_grouped_kernel<128><<<(unsigned)n_blocks, kWarpsPerBlock * 32, 0, st>>>(P, bm); break;
        default:
            fr::set_error("fr_spmm_csr_f32_grouped: d=%d unsupported (32, 64, 128)", d);
            return FR_EUNSUPPORTED;
    }
    return fr::check_launch("fr_spmm_csr_f32_grouped");
}
