// Microbenchmark / semantics probe: tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st) against the usual
// shared-memory descriptor form.  Two questions: (1) correctness - which TMEM layout does the A operand take for kind::f16 (bf16)
// and kind::tf32 (checked against a host GEMM); (2) rate - cycles per MMA when only B crosses the shared-memory port.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_a_bench tmem_a_bench.cu && ./tmem_a_bench
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16);
    const uint64_t hi = (uint64_t)64 | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
    return lo | (hi << 32);
}
template <int KIND>  // 0: kind::f16 (bf16), 1: kind::tf32
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
          "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

constexpr int M = 128;
constexpr int KB = 128;  // bytes of K per operand row (one swizzle row): 64 bf16 or 32 tf32

// integer-valued test operands (exact in bf16 / tf32 and in the fp32 accumulation)
__host__ __device__ inline float a_val(int m, int k) { return (float)(((m * 5 + k * 3) % 13) - 6); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n * 7 + k * 11) % 9) - 4); }
__device__ inline uint32_t bf16_bits(float f) { return __float_as_uint(f) >> 16; }  // exact for small integers

// mode 0: correctness (TS form), D -> out[M][N];  mode 1: time `iters` SS MMAs;  mode 2: time `iters` TS MMAs
template <int KIND, int N>
__global__ void __launch_bounds__(128, 1) probe(int mode, int iters, float* out, long long* cycles) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* smA = smem;                  // [128 rows][128 B] swizzled (SS form only)
    unsigned char* smB = smem + M * KB;         // [N rows][128 B] swizzled
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int KE = KIND == 0 ? 64 : 32;     // K elements per row
    constexpr int ES = KIND == 0 ? 2 : 4;       // element size
    constexpr int KSTEP = KIND == 0 ? 16 : 8;   // K per MMA
    // operands into swizzled shared memory: byte (row r, k) at r * 128 + ((k * ES / 16) ^ (r & 7)) * 16 + (k * ES) % 16
    for (int i = tid; i < (M + N) * KE; i += blockDim.x) {
        const int r = i / KE, k = i % KE;
        const bool isA = r < M;
        const int rr = isA ? r : r - M;
        const float v = isA ? a_val(rr, k) : b_val(rr, k);
        unsigned char* base = isA ? smA : smB;
        const int byte = rr * KB + ((((k * ES) >> 4) ^ (rr & 7)) << 4) + ((k * ES) & 15);
        if (KIND == 0) *reinterpret_cast<uint16_t*>(base + byte) = (uint16_t)bf16_bits(v);
        else *reinterpret_cast<float*>(base + byte) = v;
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    const uint32_t tD = tbase, tA = tbase + 480;   // D: columns [0, 2N) (two accumulators when they fit), A: columns [480, 512)
    // A into TMEM: lane = row m, 32 columns: bf16 pairs (k = 2 j low half, 2 j + 1 high half) or tf32 (k = j)
    {
        uint32_t r[32];
        const int m = tid;  // warp w owns lanes 32 w .. 32 w + 31
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (KIND == 0) r[j] = bf16_bits(a_val(m, 2 * j)) | (bf16_bits(a_val(m, 2 * j + 1)) << 16);
            else r[j] = __float_as_uint(a_val(m, j));
        }
        tc_st32(tA + ((uint32_t)(warp * 32) << 16), r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = KIND == 0 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24))
                                     : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24));
    const uint64_t dA = umma_desc_sw128(smem_u32(smA)), dB = umma_desc_sw128(smem_u32(smB));
    constexpr int NK = KE / KSTEP;                       // 4 MMAs cover the row
    constexpr uint32_t ACOLS = KIND == 0 ? 8 : 8;        // TMEM columns per K step (16 bf16 = 8 columns; 8 tf32 = 8 columns)
    if (mode == 0) {
        if (tid == 0) {
            for (int k = 0; k < NK; ++k) mma_ts<KIND>(tD, tA + ACOLS * k, dB + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
            tc_commit(&bar);
        }
        mbar_wait(&bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tc_ld32(tD + ((uint32_t)(warp * 32) << 16) + c0, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32 && c0 + j < N; ++j) out[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
        }
    } else {
        long long t0 = 0, t1 = 0;
        if (tid == 0) {
            t0 = clock64();
            const int my_iters = mode >= 5 ? iters / 2 : iters;
            for (int i = 0; i < my_iters; ++i) {
                const int k = i & (NK - 1);
                // modes 1 / 2: two accumulators alternating every 4 MMAs (a K loop); modes 3 / 4: `nacc` accumulators in rotation, a new one
                // every MMA (interleaved K loops of independent tiles)
                constexpr int NACC = 480 / N > 4 ? 4 : 480 / N;
                uint32_t d;
                if (mode <= 2) d = tD + (2 * N <= 480 ? (uint32_t)((i >> 2) & 1) * N : 0u);
                else if (mode >= 5) d = tD;
                else d = tD + (uint32_t)(i % NACC) * N;
                if (mode == 1 || mode == 3 || mode == 5) mma_ss<KIND>(d, dA + (uint64_t)(2 * k), dB + (uint64_t)(2 * k), idesc, 1u);
                else mma_ts<KIND>(d, tA + ACOLS * k, dB + (uint64_t)(2 * k), idesc, 1u);
            }
            tc_commit(&bar);
        }
        if (mode >= 5 && tid == 32) {  // second issuing thread (another warp): the other half of the MMAs, its own accumulator
            for (int i = 0; i < iters / 2; ++i) {
                const int k = i & (NK - 1);
                const uint32_t d = tD + (2 * N <= 480 ? (uint32_t)N : 0u);
                if (mode == 5) mma_ss<KIND>(d, dA + (uint64_t)(2 * k), dB + (uint64_t)(2 * k), idesc, 1u);
                else mma_ts<KIND>(d, tA + ACOLS * k, dB + (uint64_t)(2 * k), idesc, 1u);
            }
            tc_commit(&bar2);
        }
        mbar_wait(&bar, 0);
        if (mode >= 5) mbar_wait(&bar2, 0);
        if (tid == 0) {
            t1 = clock64();
            cycles[blockIdx.x] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int KIND, int N>
static int run(const char* name) {
    constexpr int KE = KIND == 0 ? 64 : 32;
    float* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, sizeof(float) * M * N);
    cudaMalloc(&d_cyc, sizeof(long long) * 148);
    const size_t smem = (size_t)(M + N) * KB + 2048;
    cudaFuncSetAttribute(probe<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<KIND, N><<<1, 128, smem>>>(0, 0, d_out, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: correctness launch failed: %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> out((size_t)M * N);
    cudaMemcpy(out.data(), d_out, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < KE; ++k) ref += a_val(m, k) * b_val(n, k);
            const double err = fabs((double)out[(size_t)m * N + n] - ref);
            if (err > 1e-3) { if (bad < 5) printf("  mismatch m=%d n=%d got %g ref %g\n", m, n, out[(size_t)m * N + n], ref); ++bad; }
            if (err > maxerr) maxerr = err;
        }
    printf("%s N=%d: A-in-TMEM GEMM vs host: %d mismatches of %d, max |err| %.3g\n", name, N, bad, M * N, maxerr);
    const int iters = 4096;
    for (int mode = 1; mode <= 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            probe<KIND, N><<<148, 128, smem>>>(mode, iters, d_out, d_cyc);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: timing launch failed: %s\n", name, cudaGetErrorString(e)); return 1; }
        }
        long long c;
        cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double per = (double)c / iters;
        const double ideal = (double)M * N * (KIND == 0 ? 16 : 8) / (KIND == 0 ? 4096.0 : 2048.0);
        printf("%s N=%d %s: %.1f cycles per MMA (tensor-pipe time at the nominal rate: %.0f), smem operand bytes per MMA %d -> %.1f B/clk\n", name, N,
               mode == 1 ? "A and B from shared memory, same accumulator 4 x" : mode == 2 ? "A from TMEM, B from shared memory, same accumulator 4 x"
               : mode == 3 ? "A and B from shared memory, accumulators in rotation" : mode == 4 ? "A from TMEM, B from shared memory, accumulators in rotation"
               : mode == 5 ? "A and B from shared memory, TWO issuing threads" : "A from TMEM, B from shared memory, TWO issuing threads",
               per, ideal, (mode & 1) ? (M + N) * 32 : N * 32, ((mode & 1) ? (M + N) * 32 : N * 32) / per);
    }
    cudaFree(d_out);
    cudaFree(d_cyc);
    return bad != 0;
}

int main() {
    int rc = 0;
    rc |= run<0, 128>("bf16");
    rc |= run<0, 112>("bf16");
    rc |= run<0, 64>("bf16");
    rc |= run<0, 256>("bf16");
    rc |= run<1, 128>("tf32");
    rc |= run<1, 96>("tf32");
    return rc;
}
