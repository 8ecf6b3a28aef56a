// Row gather / scatter-add / paired dot products on fp32 embedding tables (sm_100a).
// HBM-bound byte movers: 128-bit accesses, one (sub-)warp per 256-byte row.
#include "common.cuh"

namespace {

__global__ void gather_rows_kernel(const float *__restrict__ tab, int d4, const int64_t *__restrict__ idx,
                                   long long n, float *__restrict__ out) {
    const long long total = n * d4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d4;
        const int c = (int)(i - r * d4);
        reinterpret_cast<float4 *>(out)[i] = fr::ldg_f4(tab + ((size_t)idx[r] * d4 + c) * 4);
    }
}

__global__ void scatter_add_rows_kernel(const float *__restrict__ g, int d, const int64_t *__restrict__ idx,
                                        long long n, float *__restrict__ d_tab) {
    const long long total = n * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d;
        const int c = (int)(i - r * d);
        atomicAdd(d_tab + (size_t)idx[r] * d + c, __ldg(g + i));
    }
}

__global__ void pair_scores_kernel(const float *__restrict__ ut, const float *__restrict__ it, int d,
                                   const int64_t *__restrict__ user, const int64_t *__restrict__ item, long long n,
                                   float *__restrict__ scores) {
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const float *a = ut + (size_t)user[w] * d, *b = it + (size_t)item[w] * d;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) s = fmaf(__ldg(a + k), __ldg(b + k), s);
    s = fr::warp_sum(s);
    if (lane == 0) scores[w] = s;
}

int grid1d(long long work, int block) {
    long long b = (work + block - 1) / block;
    return (int)std::max<long long>(1, std::min<long long>(b, (long long)fr::num_sms() * 16));
}

}  // namespace

extern "C" int fr_gather_rows(const float *tab, int32_t d, const int64_t *idx, int64_t n, float *out, void *stream) {
    FR_REQUIRE(n >= 0 && d > 0 && d % 4 == 0, "fr_gather_rows: n=%lld d=%d (d must be a multiple of 4)", (long long)n, d);
    if (n == 0) return FR_OK;
    FR_REQUIRE(tab && idx && out, "fr_gather_rows: null pointer");
    fr::LaunchTimer _lt("gather_rows_kernel", (cudaStream_t)stream);
    gather_rows_kernel<<<grid1d(n * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(tab, d / 4, idx, n, out);
    return fr::check_launch("fr_gather_rows");
}

extern "C" int fr_scatter_add_rows(const float *g, int32_t d, const int64_t *idx, int64_t n, float *d_tab,
                                   void *stream) {
    FR_REQUIRE(n >= 0 && d > 0, "fr_scatter_add_rows: n=%lld d=%d", (long long)n, d);
    if (n == 0) return FR_OK;
    FR_REQUIRE(g && idx && d_tab, "fr_scatter_add_rows: null pointer");
    fr::LaunchTimer _lt("scatter_add_rows_kernel", (cudaStream_t)stream);
    scatter_add_rows_kernel<<<grid1d(n * d, 256), 256, 0, (cudaStream_t)stream>>>(g, d, idx, n, d_tab);
    return fr::check_launch("fr_scatter_add_rows");
}

extern "C" int fr_pair_scores(const float *user_tab, const float *item_tab, int32_t d, const int64_t *user,
                              const int64_t *item, int64_t n, float *scores, void *stream) {
    FR_REQUIRE(n >= 0 && d > 0, "fr_pair_scores: n=%lld d=%d", (long long)n, d);
    if (n == 0) return FR_OK;
    FR_REQUIRE(user_tab && item_tab && user && item && scores, "fr_pair_scores: null pointer");
    const long long blocks = (n * 32 + 255) / 256;
    fr::LaunchTimer _lt("pair_scores_kernel", (cudaStream_t)stream);
    pair_scores_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(user_tab, item_tab, d, user, item, n, scores);
    return fr::check_launch("fr_pair_scores");
}

// ------------------------------------------------------------------------------------------------
// Row-sorted COO -> CSR row pointers on the device (histogram + single-block scan).  The reference hands
// its adjacency around as a row-major-sorted COO (`coo_matrix(L)` -> `torch.sparse.FloatTensor`,
// FoodRec/models/cikm_model.py:174-180) and lets `torch.sparse.mm` rebuild CSR on every call.
namespace {
__global__ void coo_count_kernel(const int64_t *__restrict__ rows, long long nnz, int *__restrict__ cnt) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(cnt + rows[i], 1);
}
__global__ void scan_rows_kernel(const int *__restrict__ cnt, int n, int *__restrict__ row_ptr) {
    // one block, chunked inclusive scan; n + 1 outputs
    __shared__ int carry;
    __shared__ int buf[1024];
    if (threadIdx.x == 0) { carry = 0; row_ptr[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        buf[threadIdx.x] = i < n ? cnt[i] : 0;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int v = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += v;
            __syncthreads();
        }
        if (i < n) row_ptr[i + 1] = carry + buf[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
}
}  // namespace

extern "C" int fr_csr_from_coo(const int64_t *coo_rows, int64_t nnz, int32_t n_rows, int32_t *row_ptr, int32_t *scratch,
                               void *stream) {
    FR_REQUIRE(nnz >= 0 && n_rows >= 0, "fr_csr_from_coo: nnz=%lld n_rows=%d", (long long)nnz, n_rows);
    FR_REQUIRE(row_ptr && scratch && (nnz == 0 || coo_rows), "fr_csr_from_coo: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(scratch, 0, sizeof(int32_t) * (size_t)std::max(n_rows, 1), st);
    if (nnz > 0) {
        coo_count_kernel<<<grid1d(nnz, 256), 256, 0, st>>>(coo_rows, nnz, scratch);
        if (int rc = fr::check_launch("fr_csr_from_coo/count")) return rc;
    }
    scan_rows_kernel<<<1, 1024, 0, st>>>(scratch, n_rows, row_ptr);
    return fr::check_launch("fr_csr_from_coo/scan");
}


// ------------------------------------------------------------------------------------------------
// out[r] = sum_v tab_v[r] for r < rows (CLUSSL: item_emb = ingre[:I] + image[:I] + text[:I],
// FoodRec/models/pricai_modelx.py:219) and its adjoint: d_tab_v[r] = r < rows ? g[r] : 0 over each table's
// full height, i.e. the zero-fill and the slice-gradient of all three tables in one launch.
namespace {
struct Tabs4 {
    const float *in[4];
    float *out[4];
    unsigned char *mask[4];
    long long rows_total[4];
    int n;
};
__global__ void sum_rows_kernel(Tabs4 t, long long n4, float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = fr::ldg_f4(t.in[0] + 4 * i);
        for (int v = 1; v < t.n; ++v) fr::add4(a, fr::ldg_f4(t.in[v] + 4 * i));
        reinterpret_cast<float4 *>(out)[i] = a;
    }
}
__global__ void spread_rows_kernel(Tabs4 t, long long n4_src, const float *__restrict__ g, int accumulate,
                                   const unsigned char *__restrict__ src_mask, int d4) {
    const int v = blockIdx.y;
    float4 *dst = reinterpret_cast<float4 *>(t.out[v]);
    if (dst == nullptr) return;
    if (t.mask[v] != nullptr) {     // row-activity mask of the result (one byte per row)
        const long long rows_src = n4_src / d4, rows_tot = t.rows_total[v] / d4;
        for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < (accumulate ? rows_src : rows_tot);
             r += (long long)gridDim.x * blockDim.x) {
            const unsigned char m = r < rows_src ? (src_mask ? src_mask[r] : (unsigned char)1) : (unsigned char)0;
            if (accumulate) { if (m) t.mask[v][r] = 1; }
            else t.mask[v][r] = m;
        }
    }
    if (accumulate) {            // d_tab_v[r] += g[r] for r < rows; the rest already holds its value
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_src; i += (long long)gridDim.x * blockDim.x) {
            float4 a = dst[i];
            fr::add4(a, fr::ldg_f4(g + 4 * i));
            dst[i] = a;
        }
        return;
    }
    const long long tot4 = t.rows_total[v];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot4; i += (long long)gridDim.x * blockDim.x)
        dst[i] = i < n4_src ? fr::ldg_f4(g + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
}
}  // namespace

extern "C" int fr_sum_rows(const float *const *tab_host, int32_t n_tabs, int32_t d, int64_t rows, float *out, void *stream) {
    FR_REQUIRE(n_tabs >= 1 && n_tabs <= 4 && d > 0 && d % 4 == 0 && rows >= 0 && tab_host && out, "fr_sum_rows: bad argument");
    if (rows == 0) return FR_OK;
    Tabs4 t{};
    t.n = n_tabs;
    for (int v = 0; v < n_tabs; ++v) {
        FR_REQUIRE(tab_host[v], "fr_sum_rows: null table %d", v);
        t.in[v] = tab_host[v];
    }
    const long long n4 = rows * (d / 4);
    fr::LaunchTimer _lt("sum_rows_kernel", (cudaStream_t)stream);
    sum_rows_kernel<<<grid1d(n4, 256), 256, 0, (cudaStream_t)stream>>>(t, n4, out);
    return fr::check_launch("fr_sum_rows");
}

extern "C" int fr_spread_rows(const float *g, int32_t d, int64_t rows, float *const *d_tab_host, const int64_t *rows_total_host,
                              int32_t n_tabs, int32_t accumulate, const uint8_t *src_mask, uint8_t *const *mask_host,
                              void *stream) {
    FR_REQUIRE(n_tabs >= 1 && n_tabs <= 4 && d > 0 && d % 4 == 0 && rows >= 0 && g && d_tab_host && rows_total_host,
               "fr_spread_rows: bad argument");
    Tabs4 t{};
    t.n = n_tabs;
    long long mx = 0;
    for (int v = 0; v < n_tabs; ++v) {
        FR_REQUIRE(rows_total_host[v] >= rows, "fr_spread_rows: table %d shorter than the source", v);
        t.out[v] = d_tab_host[v];
        t.mask[v] = mask_host ? mask_host[v] : nullptr;
        t.rows_total[v] = rows_total_host[v] * (d / 4);
        mx = std::max(mx, t.rows_total[v]);
    }
    if (mx == 0) return FR_OK;
    fr::LaunchTimer _lt("spread_rows_kernel", (cudaStream_t)stream);
    spread_rows_kernel<<<dim3(grid1d(mx, 256), n_tabs), 256, 0, (cudaStream_t)stream>>>(t, rows * (d / 4), g, accumulate, src_mask, d / 4);
    return fr::check_launch("fr_spread_rows");
}

// ------------------------------------------------------------------------------------------------
// Mean cosine similarity between dense rows A[i] and gathered rows T[idx[i]] -- HealthRec's knowledge-
// distillation term `1 - cosine_similarity(item_know, cat(pos_e, neg_e)).mean()`
// (FoodRec/models/cikm_model.py:263-264,304-308) without materialising the gathered rows.
// cos = a.t / (max(|a|, eps) max(|t|, eps)), eps = 1e-8 (torch.nn.functional.cosine_similarity).
namespace {
__global__ void cosine_mean_fwd_kernel(const float *__restrict__ A, const float *__restrict__ T, const int64_t *__restrict__ idx,
                                       long long n, int d, float *__restrict__ cosv, float *__restrict__ na, float *__restrict__ nt) {
    const int lane = threadIdx.x & 31;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const float *a = A + (size_t)i * d, *t = T + (size_t)idx[i] * d;
    float dot = 0.f, sa = 0.f, st = 0.f;
    for (int k = lane; k < d; k += 32) {
        const float x = a[k], y = __ldg(t + k);
        dot = fmaf(x, y, dot);
        sa = fmaf(x, x, sa);
        st = fmaf(y, y, st);
    }
    dot = fr::warp_sum(dot);
    sa = fr::warp_sum(sa);
    st = fr::warp_sum(st);
    if (lane == 0) {
        const float x = fmaxf(sqrtf(sa), 1e-8f), y = fmaxf(sqrtf(st), 1e-8f);
        na[i] = x;
        nt[i] = y;
        cosv[i] = dot / (x * y);
    }
}
__global__ void mean_kernel(const float *__restrict__ v, long long n, float *__restrict__ out) {   // one block, fixed order
    __shared__ float red[32];
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    s = fr::warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        out[0] = t / (float)n;
    }
}
__global__ void cosine_mean_bwd_kernel(const float *__restrict__ A, const float *__restrict__ T, const int64_t *__restrict__ idx,
                                       long long n, int d, const float *__restrict__ cosv, const float *__restrict__ na,
                                       const float *__restrict__ nt, const float *__restrict__ g_out, float *__restrict__ dA,
                                       float *__restrict__ dT) {
    const int lane = threadIdx.x & 31;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const size_t ao = (size_t)i * d, to = (size_t)idx[i] * d;
    const float g = __ldg(g_out) / (float)n, c = cosv[i], x = na[i], y = nt[i];
    const bool ax = x > 1e-8f, ty = y > 1e-8f;      // clamped norms carry no gradient through the norm
    for (int k = lane; k < d; k += 32) {
        const float a = A[ao + k], t = __ldg(T + to + k);
        if (dA != nullptr) dA[ao + k] = g * (t / (x * y) - (ax ? c * a / (x * x) : 0.f));
        if (dT != nullptr) atomicAdd(dT + to + k, g * (a / (x * y) - (ty ? c * t / (y * y) : 0.f)));
    }
}
}  // namespace

extern "C" int fr_cosine_mean_fwd(const float *A, const float *T, const int64_t *idx, int64_t n, int32_t d, float *out,
                                  float *cosv, float *na, float *nt, void *stream) {
    FR_REQUIRE(A && T && idx && out && cosv && na && nt && n > 0 && d > 0, "fr_cosine_mean_fwd: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    {
        fr::LaunchTimer _lt("cosine_mean_fwd_kernel", st);
        cosine_mean_fwd_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(A, T, idx, n, d, cosv, na, nt);
    }
    if (int rc = fr::check_launch("fr_cosine_mean_fwd")) return rc;
    mean_kernel<<<1, 1024, 0, st>>>(cosv, n, out);
    return fr::check_launch("fr_cosine_mean_fwd/mean");
}

extern "C" int fr_cosine_mean_bwd(const float *A, const float *T, const int64_t *idx, int64_t n, int32_t d, const float *cosv,
                                  const float *na, const float *nt, const float *g_out, float *dA, float *dT, void *stream) {
    FR_REQUIRE(A && T && idx && cosv && na && nt && g_out && n > 0 && d > 0, "fr_cosine_mean_bwd: bad argument");
    fr::LaunchTimer _lt("cosine_mean_bwd_kernel", (cudaStream_t)stream);
    cosine_mean_bwd_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, T, idx, n, d, cosv, na, nt,
                                                                                              g_out, dA, dT);
    return fr::check_launch("fr_cosine_mean_bwd");
}
