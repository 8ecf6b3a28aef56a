// Negative sampling on the device.
//
// Reference: `TrainDataLoader.get_random_neg` (FoodRec/utils/dataloader.py:145-151) draws
// `np.random.randint(num_items)` in a python loop until the item is in neither the user's training list nor
// their validation/test set -- once per training sample, on the host.  Here one thread per sample draws from a
// counter-based generator (no state: the stream is a pure function of (seed, epoch step, sample, attempt), so a
// batch is reproducible and independent of the launch shape) and tests membership with a binary search in the
// user's sorted exclusion list (CSR over users).  The distribution is the reference's -- uniform over the items
// the user has not interacted with; the random stream itself is of course not numpy's (statistical parity).
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) sample_negatives_kernel(const long long *__restrict__ ptr,
                                                               const int *__restrict__ idx,
                                                               const long long *__restrict__ users, long long n,
                                                               int n_items, uint64_t seed, uint64_t step,
                                                               long long *__restrict__ out, int *__restrict__ fail) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long u = users[i];
    const long long lo0 = ptr[u], hi0 = ptr[u + 1];
    const uint64_t key = mix64(seed ^ mix64(step * 0x9E3779B97F4A7C15ULL + (uint64_t)i));
    int item = -1;
    for (int attempt = 0; attempt < 4096; ++attempt) {
        const uint64_t r = mix64(key + (uint64_t)attempt * 0xD1B54A32D192ED03ULL);
        const int cand = (int)(((r >> 32) * (uint64_t)n_items) >> 32);   // uniform in [0, n_items), bias < 2^-32 n_items
        long long lo = lo0, hi = hi0;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(idx + mid) < cand) lo = mid + 1; else hi = mid;
        }
        if (lo >= hi0 || __ldg(idx + lo) != cand) { item = cand; break; }
    }
    if (item < 0) { atomicAdd(fail, 1); item = 0; }   // user excludes (almost) every item: reported, never silent
    out[i] = item;
}

}  // namespace

extern "C" int fr_sample_negatives(const int64_t *excl_ptr, const int32_t *excl_idx, const int64_t *users, int64_t n,
                                   int32_t n_items, uint64_t seed, uint64_t step, int64_t *out_neg, int32_t *fail_count,
                                   void *stream) {
    FR_REQUIRE(n >= 0 && n_items > 0, "fr_sample_negatives: bad extents");
    if (n == 0) return FR_OK;
    FR_REQUIRE(excl_ptr && excl_idx && users && out_neg && fail_count, "fr_sample_negatives: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fr::LaunchTimer timer("sample_negatives", st);
    sample_negatives_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const long long *>(excl_ptr), excl_idx, reinterpret_cast<const long long *>(users), n, n_items,
        seed, step, reinterpret_cast<long long *>(out_neg), fail_count);
    return fr::check_launch("fr_sample_negatives");
}
