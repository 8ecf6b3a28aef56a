// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM on sm_100a, for the shapes the ranking epilogue
// can use.  One CTA per SM allocates all 512 columns; W warps (W % 4 == 0, warp w reads lane group w % 4) loop over
// loads with the wait either after every load or after a batch.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tmem_ld_bw scripts/probes/tmem_ld_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(addr, r)                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, " \
                 "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"     \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),            \
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),            \
                   "=r"(r[30]), "=r"(r[31])                                                                              \
                 : "r"(addr)                                                                                            \
                 : "memory")
#define LD16(addr, r)                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, " \
                 "%14, %15}, [%16];"                                                                                    \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])  \
                 : "r"(addr)                                                                                            \
                 : "memory")
// 16x256b.x8: 16 lanes x 256 bit x 8 = each thread gets 32 registers (lanes 0-15 of the lane group's first half)
#define LD16x256(addr, r)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, " \
                 "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"     \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),            \
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),            \
                   "=r"(r[30]), "=r"(r[31])                                                                              \
                 : "r"(addr)                                                                                            \
                 : "memory")
#define WAITLD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// MODE 0: x32 + wait each; 1: two x32 then one wait; 2: x16 + wait each; 3: 16x256b.x8 + wait each; 4: four x32 (128 regs) then wait
template <int MODE>
__global__ void __launch_bounds__(512, 1) ld_kernel(int iters, long long *clk_out, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (uint32_t)((it * 64 + (warp >> 2) * 32) & 511);
        if (MODE == 0) {
            uint32_t r[32];
            LD32(base + (col & 480), r); WAITLD();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        } else if (MODE == 1) {
            uint32_t r[32], q[32];
            LD32(base + (col & 448), r); LD32(base + (col & 448) + 32, q); WAITLD();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j] ^ q[j];
        } else if (MODE == 2) {
            uint32_t r[16];
            LD16(base + (col & 496), r); WAITLD();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= r[j];
        } else if (MODE == 3) {
            uint32_t r[32];
            LD16x256(base + (col & 448), r); WAITLD();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        } else {
            uint32_t r[32], q[32], s[32];
            const uint32_t c0 = col & 384;
            LD32(base + c0, r); LD32(base + c0 + 32, q); LD32(base + c0 + 64, s); WAITLD();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j] ^ q[j] ^ s[j];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int MODE>
void run(const char *name, int warps, double bytes_per_iter_per_warp) {
    long long *clk;
    uint32_t *sink;
    cudaMalloc(&clk, 148 * 8);
    cudaMalloc(&sink, 4);
    const int iters = 20000;
    ld_kernel<MODE><<<148, warps * 32>>>(100, clk, sink);
    cudaDeviceSynchronize();
    ld_kernel<MODE><<<148, warps * 32>>>(iters, clk, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0];
    printf("%-28s warps=%2d  %s  clk/iter/warp=%.1f  B/clk/SM=%.1f\n", name, warps, cudaGetErrorString(e), c / iters,
           bytes_per_iter_per_warp * warps * iters / c);
    cudaFree(clk);
    cudaFree(sink);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("32x32b.x32 wait-each", w, 4096);
        run<1>("32x32b.x32 x2 then wait", w, 8192);
        run<2>("32x32b.x16 wait-each", w, 2048);
        run<3>("16x256b.x8 wait-each", w, 4096);
        run<4>("32x32b.x32 x3 then wait", w, 12288);
    }
    return 0;
}
