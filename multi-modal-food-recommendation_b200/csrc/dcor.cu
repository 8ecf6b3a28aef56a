// Distance-correlation contrastive term over V <= 3 gathered views (forward + backward), sm_100a.
//
// The reference evaluates, per pair of [n, d] views, two n x n distance matrices, their double
// centring and three inner products with ~40 elementwise launches and as many again in autograd
// (FoodRec/models/pricai_modelx.py:409-437, called three times at :263).  Here the whole term is
// three launches forward and one backward:
//   dcor_dist_kernel    D_v[i][j] = sqrt(max(r_i - 2 x_i.x_j + r_j, 0) + 1e-8), row means (rows are
//                       gathered from the propagated tables on the fly, no [n, d] copies)
//   dcor_dot_kernel     s_vw = sum_ij A_v A_w / n^2 for all view pairs with A = double-centred D
//                       (never stored), block partials folded in fixed order by the last block,
//                       which also emits dcor_p and the partial derivatives d dcor_p / d s_xy
//   dcor_w_kernel       W = (sum_w C_vw A_w) / (2 D_v) off-diagonal and its row sums
//   dcor_wx_kernel      dX_v = 4 (diag(W 1) X - W X), scatter-added straight into the dense table gradients.
// The double-centred matrices are projections (P A P = A), so the gradient w.r.t. D_v needs no
// re-centring; the diagonal of D (mathematically constant) is dropped analytically instead of being
// left to cancel numerically as autograd does.  n = 2B = 1024, d = 64: latency-bound, all in L2.
#include "common.cuh"

namespace {

constexpr int kMaxV = 3;
constexpr int kMaxP = 3;
constexpr int kThreads = 256;
constexpr int kRB2 = 4;     // rows per CTA in the dot kernel
constexpr int kMaxD = 128;  // max embedding width
constexpr float kEpsD = 1e-8f;

struct Views {
    const float *tab[kMaxV];
    float *dtab[kMaxV];
    unsigned char *mask[kMaxV];   // optional row-activity masks of the dense gradients
    int V;
};
struct Pairs {
    int a[kMaxP], b[kMaxP];
    int P;
};

__device__ __forceinline__ float block_sum(float v, float *sh) {
    v = fr::warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
    return t;
}

__device__ __forceinline__ double block_sum_d(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
    return t;
}

// ---------------------------------------------------------------------------------------- K1
// 64 x 64 tile of D_v per CTA (256 threads, 4 x 4 outputs each) from transposed shared tiles of the
// gathered rows; squared norms ride along in the k loop in the same summation order as the dot product,
// so the diagonal is exactly sqrt(1e-8).  Row sums go out as one partial per (row, column tile) and are
// folded in fixed order by dcor_rowmean_kernel.
constexpr int kT = 64;  // tile edge
template <int D>
__global__ void __launch_bounds__(kThreads)
dcor_dist_kernel(Views vw, const int64_t *__restrict__ idx, int n, float *__restrict__ Dm,
                 float *__restrict__ rowpart, int n_tiles) {
    __shared__ __align__(16) float Xi[D][kT + 4];
    __shared__ __align__(16) float Xj[D][kT + 4];
    const int v = blockIdx.z;
    const float *__restrict__ tab = vw.tab[v];
    const int i0 = blockIdx.y * kT, j0 = blockIdx.x * kT;
    for (int t = threadIdx.x; t < kT * D; t += kThreads) {
        const int r = t / D, k = t - r * D;
        Xi[k][r] = (i0 + r < n) ? __ldg(tab + (size_t)idx[i0 + r] * D + k) : 0.f;
        Xj[k][r] = (j0 + r < n) ? __ldg(tab + (size_t)idx[j0 + r] * D + k) : 0.f;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[4][4], ri[4] = {0.f, 0.f, 0.f, 0.f}, rj[4] = {0.f, 0.f, 0.f, 0.f};
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
        for (int a = 0; a < 4; ++a) {
            ri[a] = fmaf(a4[a], a4[a], ri[a]);
            rj[a] = fmaf(b4[a], b4[a], rj[a]);
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
        }
    }
    float *__restrict__ Dv = Dm + (size_t)v * n * n;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        float d4[4], rs = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float m = (ri[a] - 2.f * acc[a][b]) + rj[b];
            d4[b] = sqrtf(fmaxf(m, 0.f) + kEpsD);
            if (j0 + tx * 4 + b < n) rs += d4[b];
        }
        if (i < n) {
            if (j0 + tx * 4 + 3 < n && (n & 3) == 0)
                *reinterpret_cast<float4 *>(Dv + (size_t)i * n + j0 + tx * 4) = make_float4(d4[0], d4[1], d4[2], d4[3]);
            else
                for (int b = 0; b < 4; ++b)
                    if (j0 + tx * 4 + b < n) Dv[(size_t)i * n + j0 + tx * 4 + b] = d4[b];
        }
        // fold the 16 column-threads of this row (a half-warp), one partial per (row, column tile)
        rs += __shfl_xor_sync(0xffffffffu, rs, 8);
        rs += __shfl_xor_sync(0xffffffffu, rs, 4);
        rs += __shfl_xor_sync(0xffffffffu, rs, 2);
        rs += __shfl_xor_sync(0xffffffffu, rs, 1);
        if (tx == 0 && i < n) rowpart[((size_t)v * n_tiles + blockIdx.x) * n + i] = rs;
    }
}

// The centring and the six inner products run in DOUBLE (B200's fp64 pipe makes that free at n = 1024): the cross
// products sum_ij A_v A_w of nearly independent views cancel to ~1e-3 of their absolute mass, and with fp32 centring
// their rounding error reached the gradients through 1 / (2 sqrt(s_vw)) at ~3e-5 relative (the reference's own fp32
// autograd sits at 3e-4..7e-4 from the fp64 result; with double centring this path is at ~7e-7).
__global__ void dcor_rowmean_kernel(const float *__restrict__ rowpart, int n, int n_tiles, int V,
                                    float *__restrict__ rowmean, double *__restrict__ rowmean64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (i >= n) return;
    double s = 0.0;
    for (int t = 0; t < n_tiles; ++t) s += (double)rowpart[((size_t)v * n_tiles + t) * n + i];
    s /= (double)n;
    rowmean64[(size_t)v * n + i] = s;
    rowmean[(size_t)v * n + i] = (float)s;
}

// ---------------------------------------------------------------------------------------- K2
// out layout: out[p] = dcor_p; dfds[p*3 + {0,1,2}] = d dcor_p / d {s_ab, s_aa, s_bb}; gm[v] grand means.
__global__ void __launch_bounds__(kThreads)
dcor_dot_kernel(int V, Pairs pr, int n, const float *__restrict__ Dm, const double *__restrict__ rowmean,
                float *__restrict__ out, float *__restrict__ dfds, float *__restrict__ gm_out,
                float *__restrict__ ws, float scale) {
    __shared__ double red[kThreads / 32];
    __shared__ double gm[kMaxV];
    __shared__ double fin[6];
    __shared__ int last;
    double *__restrict__ part = reinterpret_cast<double *>(ws + 8);     // [grid][8] block partials
    for (int v = 0; v < V; ++v) {
        double s = 0.0;
        for (int j = threadIdx.x; j < n; j += kThreads) s += rowmean[(size_t)v * n + j];
        s = block_sum_d(s, red);
        if (threadIdx.x == 0) gm[v] = s / (double)n;
    }
    __syncthreads();
    // accumulate s[v][w], v <= w, slot = v*3 + w - (v*(v+1))/2 -> (00,01,02,11,12,22)
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int i0 = blockIdx.x * kRB2;
    for (int r = 0; r < kRB2 && i0 + r < n; ++r) {
        const int i = i0 + r;
        double rmi[kMaxV];
        for (int v = 0; v < kMaxV; ++v) rmi[v] = v < V ? rowmean[(size_t)v * n + i] : 0.0;
#pragma unroll 4
        for (int j = threadIdx.x; j < n; j += kThreads) {
            double A[kMaxV];
#pragma unroll
            for (int v = 0; v < kMaxV; ++v)
                A[v] = v < V ? ((((double)__ldg(Dm + ((size_t)v * n + i) * n + j) - rowmean[(size_t)v * n + j]) - rmi[v]) + gm[v])
                             : 0.0;
            acc[0] = fma(A[0], A[0], acc[0]);
            acc[1] = fma(A[0], A[1], acc[1]);
            acc[2] = fma(A[0], A[2], acc[2]);
            acc[3] = fma(A[1], A[1], acc[3]);
            acc[4] = fma(A[1], A[2], acc[4]);
            acc[5] = fma(A[2], A[2], acc[5]);
        }
    }
    for (int q = 0; q < 6; ++q) {
        const double s = block_sum_d(acc[q], red);
        if (threadIdx.x == 0) __stcg(part + (size_t)blockIdx.x * 8 + q, s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(reinterpret_cast<int *>(ws), 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 6) {
        double s = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(part + (size_t)b * 8 + threadIdx.x);
        fin[threadIdx.x] = s / ((double)n * (double)n);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        auto S = [&](int a, int b) -> double {
            if (a > b) { const int t = a; a = b; b = t; }
            return fin[a * 3 + b - (a * (a + 1)) / 2];
        };
        double total = 0.0;
        for (int p = 0; p < pr.P; ++p) {
            const int a = pr.a[p], b = pr.b[p];
            const double sab = S(a, b), saa = S(a, a), sbb = S(b, b);
            const double cab = sqrt(fmax(sab, 0.0) + 1e-8);
            const double caa = sqrt(fmax(saa, 0.0) + 1e-8);
            const double cbb = sqrt(fmax(sbb, 0.0) + 1e-8);
            const double q = sqrt(fmax(caa * cbb, 0.0) + 1e-10);
            const double o = (double)scale * (cab / q);
            out[p] = (float)o;
            total += o;
            const double dq = -cab / (2.0 * q * q * q);  // d f / d (caa*cbb)
            dfds[p * 3 + 0] = sab > 0.0 ? (float)((double)scale / (2.0 * cab * q)) : 0.f;
            dfds[p * 3 + 1] = saa > 0.0 ? (float)((double)scale * dq * cbb / (2.0 * caa)) : 0.f;
            dfds[p * 3 + 2] = sbb > 0.0 ? (float)((double)scale * dq * caa / (2.0 * cbb)) : 0.f;
        }
        out[pr.P] = (float)total;      // sum of the (scaled) terms, the value CLUSSL's loss uses
        for (int v = 0; v < V; ++v) gm_out[v] = (float)gm[v];
        *reinterpret_cast<int *>(ws) = 0;
    }
}

// ---------------------------------------------------------------------------------------- K3
// Backward in two launches (one kernel that interleaved the W computation with the product spent its time in chains of
// dependent gathers at two CTAs per SM: 72 us at n = 1024; profiles/r2_dcor_bwd_*):
//   dcor_w_kernel     W_v[i][j] = (sum_u C_vu A_u[i][j]) / (2 D_v[i][j]) off the diagonal (0 on it and where D_v == sqrt(eps),
//                     i.e. duplicate rows: the analytic value), and the row sums  wsum_v[i]; one pass over the three distance
//                     matrices, fully coalesced, 4 rows per CTA
//   dcor_wx_kernel    dX_v[idx[i]] += 4 (wsum_v[i] x_i - sum_j W_v[i][j] x_j): a batched 32 x D x n product per CTA with the
//                     column range split over blockIdx.z, register-prefetched operand tiles, float4 atomics into the dense
//                     table gradients
__global__ void __launch_bounds__(kThreads)
dcor_w_kernel(int V, Pairs pr, int n, const float *__restrict__ Dm, const float *__restrict__ rowmean,
              const float *__restrict__ gm, const float *__restrict__ dfds, const float *__restrict__ g_out,
              float *__restrict__ W, float *__restrict__ wsum) {
    __shared__ float Cm[kMaxV][kMaxV];
    __shared__ float gms[kMaxV];
    __shared__ float red[kThreads / 32];
    if (threadIdx.x == 0) {
        // C_vw: coefficient of A_w in d(sum_p g_p dcor_p) / d D_v
        float c[kMaxV][kMaxV] = {};
        const float inv = 1.f / ((float)n * (float)n);
        for (int p = 0; p < pr.P; ++p) {
            const float g = __ldg(g_out + p);
            const int a = pr.a[p], b = pr.b[p];
            c[a][b] += g * dfds[p * 3 + 0] * inv;
            c[a][a] += 2.f * g * dfds[p * 3 + 1] * inv;
            c[b][a] += g * dfds[p * 3 + 0] * inv;
            c[b][b] += 2.f * g * dfds[p * 3 + 2] * inv;
        }
        for (int v = 0; v < kMaxV; ++v) {
            gms[v] = v < V ? gm[v] : 0.f;
            for (int w = 0; w < kMaxV; ++w) Cm[v][w] = c[v][w];
        }
    }
    __syncthreads();
    const float eps_d = sqrtf(kEpsD);
    const int i0 = blockIdx.x * kRB2;
    for (int r = 0; r < kRB2 && i0 + r < n; ++r) {
        const int i = i0 + r;
        float rmi[kMaxV], ws[kMaxV] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < kMaxV; ++u) rmi[u] = u < V ? __ldg(rowmean + (size_t)u * n + i) : 0.f;
#pragma unroll 2
        for (int j = threadIdx.x; j < n; j += kThreads) {
            float d[kMaxV], A[kMaxV];
#pragma unroll
            for (int u = 0; u < kMaxV; ++u) {
                d[u] = u < V ? __ldg(Dm + ((size_t)u * n + i) * n + j) : 1.f;
                A[u] = u < V ? ((d[u] - __ldg(rowmean + (size_t)u * n + j)) - rmi[u]) + gms[u] : 0.f;
            }
#pragma unroll
            for (int v = 0; v < kMaxV; ++v) {
                if (v < V) {
                    float g = 0.f;
#pragma unroll
                    for (int u = 0; u < kMaxV; ++u) g = fmaf(Cm[v][u], A[u], g);
                    const float w = (i != j && d[v] > eps_d) ? g / (2.f * d[v]) : 0.f;
                    W[((size_t)v * n + i) * n + j] = w;
                    ws[v] += w;
                }
            }
        }
        for (int v = 0; v < V; ++v) {
            const float t = block_sum(ws[v], red);
            if (threadIdx.x == 0) wsum[(size_t)v * n + i] = t;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(kThreads)
dcor_wx_kernel(Views vw, const int64_t *__restrict__ idx, int n, const float *__restrict__ W,
               const float *__restrict__ wsum) {
    constexpr int RB = 32, JT = 32, KQ = D / 4;           // 32 rows x (D/4) float4 columns per CTA, 32 columns of W per step
    constexpr int RPT = RB * KQ / kThreads;               // output rows per thread (D = 64: 2, D = 32: 1)
    static_assert(RB * KQ % kThreads == 0 && RB * JT == 4 * kThreads && JT * KQ <= 2 * kThreads, "tile/threads mismatch");
    __shared__ __align__(16) float Ws[2][RB][JT + 4];
    __shared__ __align__(16) float Xs[2][JT][D];
    const int v = blockIdx.y;
    float *__restrict__ dtab = vw.dtab[v];
    if (dtab == nullptr) return;
    const float *__restrict__ tab = vw.tab[v];
    const float *__restrict__ Wv = W + (size_t)v * n * n;
    const int i0 = blockIdx.x * RB;
    const int jt_total = (n + JT - 1) / JT, jt_per = (jt_total + gridDim.z - 1) / gridDim.z;
    const int jt_lo = blockIdx.z * jt_per, jt_hi = min(jt_total, jt_lo + jt_per);
    const int okq = threadIdx.x % KQ, orow = (threadIdx.x / KQ) * RPT;      // this thread's float4 column and first row
    // operand loads of one step: W tile RB x JT = 1024 floats (one float4 per thread), X tile JT x D (<= 2 float4 per thread)
    const int wr = threadIdx.x / (JT / 4), wc = (threadIdx.x % (JT / 4)) * 4;
    float4 wreg, xreg[2];
    auto fetch = [&](int jt) {
        const int j0 = jt * JT;
        const int i = i0 + wr;
        wreg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            if (j0 + wc + 3 < n && (n & 3) == 0) wreg = fr::ldg_f4(Wv + (size_t)i * n + j0 + wc);
            else {
                float t[4] = {0.f, 0.f, 0.f, 0.f};
                for (int e = 0; e < 4; ++e)
                    if (j0 + wc + e < n) t[e] = __ldg(Wv + (size_t)i * n + j0 + wc + e);
                wreg = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int t = threadIdx.x + e * kThreads;
            xreg[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < JT * KQ) {
                const int r = t / KQ, q = t % KQ;
                if (j0 + r < n) xreg[e] = fr::ldg_f4(tab + (size_t)__ldg(idx + j0 + r) * D + 4 * q);
            }
        }
    };
    auto stash = [&](int buf) {
        *reinterpret_cast<float4 *>(&Ws[buf][wr][wc]) = wreg;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int t = threadIdx.x + e * kThreads;
            if (t < JT * KQ) *reinterpret_cast<float4 *>(&Xs[buf][t / KQ][4 * (t % KQ)]) = xreg[e];
        }
    };
    float4 acc[RPT];
#pragma unroll
    for (int a = 0; a < RPT; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (jt_lo < jt_hi) {
        fetch(jt_lo);
        stash(0);
        __syncthreads();
        for (int jt = jt_lo; jt < jt_hi; ++jt) {
            const int buf = (jt - jt_lo) & 1;
            if (jt + 1 < jt_hi) fetch(jt + 1);              // next step's operands travel while this one is multiplied
#pragma unroll 8
            for (int c = 0; c < JT; ++c) {
                const float4 x = *reinterpret_cast<const float4 *>(&Xs[buf][c][4 * okq]);
#pragma unroll
                for (int a = 0; a < RPT; ++a) fr::fma4(acc[a], Ws[buf][orow + a][c], x);
            }
            if (jt + 1 < jt_hi) stash(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int a = 0; a < RPT; ++a) {
        const int i = i0 + orow + a;
        if (i >= n) continue;
        const size_t row = (size_t)__ldg(idx + i);
        float4 o = make_float4(-4.f * acc[a].x, -4.f * acc[a].y, -4.f * acc[a].z, -4.f * acc[a].w);
        if (blockIdx.z == 0) {       // the diag(W 1) X term once per row
            const float w4 = 4.f * __ldg(wsum + (size_t)v * n + i);
            const float4 xi = fr::ldg_f4(tab + row * D + 4 * okq);
            o.x = fmaf(w4, xi.x, o.x);
            o.y = fmaf(w4, xi.y, o.y);
            o.z = fmaf(w4, xi.z, o.z);
            o.w = fmaf(w4, xi.w, o.w);
            if (vw.mask[v] != nullptr && okq == 0) vw.mask[v][row] = 1;
        }
        atomicAdd(reinterpret_cast<float4 *>(dtab + row * D + 4 * okq), o);
    }
}

int fill(Views &vw, Pairs &pr, int V, const float *const *tab, float *const *dtab, int P, const int32_t *pairs) {
    FR_REQUIRE(V >= 1 && V <= kMaxV && P >= 1 && P <= kMaxP && tab && pairs, "dcor: V=%d P=%d out of range", V, P);
    vw.V = V;
    for (int v = 0; v < kMaxV; ++v) {
        vw.tab[v] = v < V ? tab[v] : nullptr;
        vw.dtab[v] = (v < V && dtab) ? dtab[v] : nullptr;
        vw.mask[v] = nullptr;
        FR_REQUIRE(v >= V || tab[v], "dcor: null view table %d", v);
    }
    pr.P = P;
    for (int p = 0; p < kMaxP; ++p) {
        pr.a[p] = p < P ? pairs[2 * p] : 0;
        pr.b[p] = p < P ? pairs[2 * p + 1] : 0;
        FR_REQUIRE(p >= P || (pr.a[p] >= 0 && pr.a[p] < V && pr.b[p] >= 0 && pr.b[p] < V && pr.a[p] != pr.b[p]),
                   "dcor: pair %d out of range", p);
    }
    return FR_OK;
}

}  // namespace

// ws layout (floats): [0] ticket, [8 ...) dot-kernel block partials (8 doubles per block), the double row means
// [V][n], then the row-sum partials [V][n_tiles][n]
static int64_t dot_blocks(int n) { return (n + kRB2 - 1) / kRB2; }
static int64_t rm64_off(int n) { return 8 + 16 * dot_blocks(n); }
static int64_t rowpart_off(int n) { return rm64_off(n) + 2 * (int64_t)kMaxV * n; }
extern "C" int64_t fr_dcor_ws_floats(int32_t n) {
    return rowpart_off(n) + (int64_t)kMaxV * ((n + kT - 1) / kT) * n;
}

extern "C" int fr_dcor_fwd(const float *const *tab_host, int32_t V, int32_t d, const int64_t *idx, int32_t n,
                           const int32_t *pairs_host, int32_t P, float scale, float *Dm, float *rowmean, float *out,
                           float *dfds, float *gm, float *ws, void *stream) {
    FR_REQUIRE(idx && Dm && rowmean && out && dfds && gm && ws && n > 0, "fr_dcor_fwd: bad argument");
    Views vw;
    Pairs pr;
    if (int rc = fill(vw, pr, V, tab_host, nullptr, P, pairs_host)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (n + kT - 1) / kT;
    FR_REQUIRE(((uintptr_t)ws & 7) == 0, "fr_dcor_fwd: workspace must be 8-byte aligned");
    float *rowpart = ws + rowpart_off(n);
    double *rowmean64 = reinterpret_cast<double *>(ws + rm64_off(n));
    dim3 g1(n_tiles, n_tiles, V);
    {
    fr::LaunchTimer _lt("dcor_dist_kernel", st);
    if (d == 64) dcor_dist_kernel<64><<<g1, kThreads, 0, st>>>(vw, idx, n, Dm, rowpart, n_tiles);
    else if (d == 32) dcor_dist_kernel<32><<<g1, kThreads, 0, st>>>(vw, idx, n, Dm, rowpart, n_tiles);
    else {
        fr::set_error("fr_dcor_fwd: d=%d unsupported (32, 64)", d);
        return FR_EUNSUPPORTED;
    }
    }
    if (int rc = fr::check_launch("fr_dcor_fwd/dist")) return rc;
    dcor_rowmean_kernel<<<dim3((n + 255) / 256, V), 256, 0, st>>>(rowpart, n, n_tiles, V, rowmean, rowmean64);
    if (int rc = fr::check_launch("fr_dcor_fwd/rowmean")) return rc;
    fr::LaunchTimer _lt2("dcor_dot_kernel", st);
    dcor_dot_kernel<<<(n + kRB2 - 1) / kRB2, kThreads, 0, st>>>(V, pr, n, Dm, rowmean64, out, dfds, gm, ws, scale);
    return fr::check_launch("fr_dcor_fwd/dot");
}

extern "C" int64_t fr_dcor_bwd_ws_floats(int32_t n) { return (int64_t)kMaxV * n * n + (int64_t)kMaxV * n; }

extern "C" int fr_dcor_bwd(const float *const *tab_host, int32_t V, int32_t d, const int64_t *idx, int32_t n,
                           const int32_t *pairs_host, int32_t P, const float *Dm, const float *rowmean,
                           const float *dfds, const float *gm, const float *g_out, float *const *d_tab_host,
                           uint8_t *const *mask_host, float *ws, void *stream) {
    FR_REQUIRE(idx && Dm && rowmean && dfds && gm && g_out && d_tab_host && ws && n > 0, "fr_dcor_bwd: bad argument");
    FR_REQUIRE(((uintptr_t)ws & 15) == 0, "fr_dcor_bwd: workspace must be 16-byte aligned");
    Views vw;
    Pairs pr;
    if (int rc = fill(vw, pr, V, tab_host, d_tab_host, P, pairs_host)) return rc;
    for (int v = 0; v < V; ++v)
        FR_REQUIRE(vw.dtab[v] == nullptr || ((uintptr_t)vw.dtab[v] & 15) == 0, "fr_dcor_bwd: gradient tables must be 16-byte aligned");
    if (mask_host)
        for (int v = 0; v < V; ++v) vw.mask[v] = mask_host[v];
    cudaStream_t st = (cudaStream_t)stream;
    float *W = ws, *wsum = ws + (size_t)kMaxV * n * n;
    {
        fr::LaunchTimer _lt("dcor_w_kernel", st);
        dcor_w_kernel<<<(n + kRB2 - 1) / kRB2, kThreads, 0, st>>>(V, pr, n, Dm, rowmean, gm, dfds, g_out, W, wsum);
        if (int rc = fr::check_launch("fr_dcor_bwd/w")) return rc;
    }
    dim3 g((n + 31) / 32, V, 8);
    fr::LaunchTimer _lt("dcor_wx_kernel", st);
    if (d == 64) dcor_wx_kernel<64><<<g, kThreads, 0, st>>>(vw, idx, n, W, wsum);
    else if (d == 32) dcor_wx_kernel<32><<<g, kThreads, 0, st>>>(vw, idx, n, W, wsum);
    else {
        fr::set_error("fr_dcor_bwd: d=%d unsupported (32, 64)", d);
        return FR_EUNSUPPORTED;
    }
    return fr::check_launch("fr_dcor_bwd");
}
