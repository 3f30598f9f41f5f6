// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM with 1 / 4 / 8 warps, to decide whether a tensor-core
// dictionary-matching kernel (2 accumulator reads per (pixel, atom) score) can beat the FP32-FMA kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench ldtm_bench.cu && ./ldtm_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: loads only; mode 1: loads + the matching epilogue arithmetic (s = re*re + im*im, running max) on the loaded values
__global__ void __launch_bounds__(256, 1) ldtm_bench(int iters, int mode, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    float best = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t a[32], b[32];
        tc_ld32(taddr + (uint32_t)((i & 1) * 64), a);
        tc_ld32(taddr + (uint32_t)((i & 1) * 64 + 32), b);
        tc_wait_ld();
        if (mode == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float re = __uint_as_float(a[j]), im = __uint_as_float(b[j]);
                best = fmaxf(best, fmaf(im, im, re * re));
            }
        } else {
            best += __uint_as_float(a[0] ^ b[31]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = best;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}

int main() {
    long long* d_cyc;
    float* d_sink;
    cudaMalloc(&d_cyc, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 148 * 256 * sizeof(float));
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 4, 8}) {
            for (int rep = 0; rep < 2; ++rep) {
                ldtm_bench<<<148, warps * 32, 0>>>(iters, mode, d_cyc, d_sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long c;
            cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * 2 * 32 * 32 * 4 * warps;   // per SM
            printf("mode %d warps %d: %lld cycles, %.1f B/clk/SM TMEM read, %.2f score elements/clk/SM\n", mode, warps, c, bytes / c,
                   (double)iters * 32 * 32 * warps / c);
        }
    return 0;
}
