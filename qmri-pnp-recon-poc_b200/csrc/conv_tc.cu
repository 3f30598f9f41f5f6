// K3 (tensor mode) - DRUNet 3x3 convolutions as a tcgen05 implicit GEMM with split-bf16 operands.
//
// Reference being replaced: the 58 bias-free 3x3 convolutions of UNetRes
//   PyTorch_Denoiser/zhang_dpir_testing_code/network_unet.py:106-117, basicblock.py:61-98,211-223.
//
// Precision: the parity bar is 1e-4 relative L2 against the fp32 CPU forward.  Single-pass TF32
// (5e-4) and bf16 (4e-3) operands fail it, so every fp32 value a is carried as
// a_hi = bf16(a), a_lo = bf16(a - a_hi) (16 significant bits) and a*w is evaluated as
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation in TMEM (three kind::f16 MMAs per
// K step; the dropped a_lo*w_lo term is 2^-18 relative).
//
// Data layout: activations are two bf16 planes (hi, lo), each [S][Y][X][C] (channels innermost,
// 2 B): same bytes as one fp32 tensor.  Weights are [Cout][9*Cin] K-major, tap-major K, hi / lo.
//
// Kernel: persistent, warp-specialised.  GEMM tile = 128 output pixels (a BH x BW patch of one
// slice) x BN output channels; K loop over 9 taps x Cin/64.  For each k-block the producer thread
// issues 4 TMA tile loads (A_hi, A_lo as 4-D boxes of the shifted patch - out-of-image pixels are
// zero-filled by TMA, which is the conv's zero padding; B_hi, B_lo as 2-D boxes) into a 128B-swizzled
// shared-memory ring; one MMA thread issues 12 tcgen05.mma (M128 x BN x K16) per k-block into one of
// two TMEM accumulators; four epilogue warps drain the other accumulator with tcgen05.ld, add the
// residuals, apply ReLU, re-split to (hi, lo) and store.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "conv_kernels.h"

namespace {

constexpr int TC_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;        // one 128-byte swizzle row of bf16
constexpr uint32_t A_TILE_BYTES = TC_BM * TC_BK * 2;  // 16 KB

template <int BN>
struct TcCfg {
    static constexpr uint32_t B_TILE_BYTES = BN * TC_BK * 2;
    static constexpr uint32_t STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = (BN == 64) ? 4 : 3;
    static constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulators
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must abort the kernel, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16);
    const uint64_t hi = (uint64_t)64 | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
    return lo | (hi << 32);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack_bf16(uint32_t u, float& a, float& b) {
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xffff0000u);
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, TcConvParams p) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ unsigned char tc_smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t raw = smem_u32(tc_smem_raw);
    unsigned char* smem = tc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;                  // [STAGES]
    uint64_t* empty = bars + STAGES;        // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;    // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = 9 * (p.Cin / TC_BK);
    const int cblocks = p.Cin / TC_BK;
    const int NT = p.Cout / BN;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_xy * NT * p.S;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmA_lo);
        tma_prefetch_desc(&tmB_hi);
        tma_prefetch_desc(&tmB_lo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = t % NT;
                int r = t / NT;
                const int txy = r % tiles_xy;
                const int s = r / tiles_xy;
                const int x0 = (txy % p.tiles_x) * p.BW, y0 = (txy / p.tiles_x) * p.BH;
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    for (int cb = 0; cb < cblocks; ++cb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        unsigned char* st = smem + (size_t)stage * Cfg::STAGE_BYTES;
                        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        tma_load_4d(st, &tmA_hi, &full[stage], cb * TC_BK, x0 + dx, y0 + dy, s);
                        tma_load_4d(st + A_TILE_BYTES, &tmA_lo, &full[stage], cb * TC_BK, x0 + dx, y0 + dy, s);
                        tma_load_2d(st + 2 * A_TILE_BYTES, &tmB_hi, &full[stage], tap * p.Cin + cb * TC_BK, nt * BN);
                        tma_load_2d(st + 2 * A_TILE_BYTES + Cfg::B_TILE_BYTES, &tmB_lo, &full[stage], tap * p.Cin + cb * TC_BK, nt * BN);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int ab = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[ab], aphase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::STAGE_BYTES);
                    const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + A_TILE_BYTES);
                    const uint64_t b_hi = umma_desc(sa + 2 * A_TILE_BYTES), b_lo = umma_desc(sa + 2 * A_TILE_BYTES + Cfg::B_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);  // 32 bytes along K inside the swizzle atom (16-byte units)
                        tc_mma(tmem_d, a_hi + ko, b_hi + ko, idesc, (kb | k) ? 1u : 0u);
                        tc_mma(tmem_d, a_hi + ko, b_lo + ko, idesc, 1u);
                        tc_mma(tmem_d, a_lo + ko, b_hi + ko, idesc, 1u);
                    }
                    tc_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&tfull[ab]);  // accumulator complete
            }
        }
    } else {
        // ================= epilogue warps (TMEM lanes (warp % 4) * 32 ..) =================
        const int lg = warp & 3;
        const int row = lg * 32 + lane;  // row of the 128-pixel tile
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int ab = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int nt = t % NT;
            int r = t / NT;
            const int txy = r % tiles_xy;
            const int s = r / tiles_xy;
            const int x = (txy % p.tiles_x) * p.BW + row % p.BW;
            const int y = (txy / p.tiles_x) * p.BH + row / p.BW;
            const bool ok = (x < p.W) && (y < p.H);
            const size_t o = (((size_t)s * p.H + y) * p.W + x) * p.Cout + (size_t)nt * BN;
            mbar_wait(&tfull[ab], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ab * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t acc[16];
                tc_ld16(taddr + c0, acc);
                tc_wait_ld();
                if (ok) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
                    if (p.res1_hi) {
                        const uint4* rh = reinterpret_cast<const uint4*>(p.res1_hi + o + c0);
                        const uint4* rl = reinterpret_cast<const uint4*>(p.res1_lo + o + c0);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint4 a = rh[h], b = rl[h];
                            uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float h0, h1, l0, l1;
                                unpack_bf16(ua[j], h0, h1);
                                unpack_bf16(ub[j], l0, l1);
                                v[8 * h + 2 * j] += h0 + l0;
                                v[8 * h + 2 * j + 1] += h1 + l1;
                            }
                        }
                    }
                    if (p.res2_hi) {
                        const uint4* rh = reinterpret_cast<const uint4*>(p.res2_hi + o + c0);
                        const uint4* rl = reinterpret_cast<const uint4*>(p.res2_lo + o + c0);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint4 a = rh[h], b = rl[h];
                            uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float h0, h1, l0, l1;
                                unpack_bf16(ua[j], h0, h1);
                                unpack_bf16(ub[j], l0, l1);
                                v[8 * h + 2 * j] += h0 + l0;
                                v[8 * h + 2 * j + 1] += h1 + l1;
                            }
                        }
                    }
                    uint32_t ph[8], pl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float a = v[2 * j], b = v[2 * j + 1];
                        if (p.relu) {
                            a = fmaxf(a, 0.f);
                            b = fmaxf(b, 0.f);
                        }
                        ph[j] = pack_bf16(a, b);
                        float ha, hb;
                        unpack_bf16(ph[j], ha, hb);
                        pl[j] = pack_bf16(a - ha, b - hb);
                    }
                    uint4* oh = reinterpret_cast<uint4*>(p.out_hi + o + c0);
                    uint4* ol = reinterpret_cast<uint4*>(p.out_lo + o + c0);
                    oh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    oh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                    ol[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    ol[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[ab]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

}  // namespace

int tc_make_act_map(void* out_map, const void* base, int S, int H, int W, int C, int BW, int BH) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", S, H, W, C, (int)r);
    return QMRI_OK;
}

int tc_make_weight_map(void* out_map, const void* base, int K, int Cout, int BN) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled(weights %d x %d) failed: %d", Cout, K, (int)r);
    return QMRI_OK;
}

int tc_tile_shape(int W, int H, int* BW, int* BH) {
    // 128-pixel patches; pick the widest box that divides the row (else the one wasting least)
    const int cand[4][2] = {{16, 8}, {8, 16}, {32, 4}, {4, 32}};
    int best = -1;
    double best_eff = -1;
    for (int i = 0; i < 4; ++i) {
        int bw = cand[i][0], bh = cand[i][1];
        double eff = ((double)W * H) / ((double)((W + bw - 1) / bw * bw) * ((H + bh - 1) / bh * bh));
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = i;
        }
    }
    *BW = cand[best][0];
    *BH = cand[best][1];
    return QMRI_OK;
}

int tc_block_n(int Cout) { return Cout == 64 ? 64 : 128; }

int conv3x3_tc(qmri_ctx* ctx, const void* mapA_hi, const void* mapA_lo, const void* mapB_hi, const void* mapB_lo,
               const TcConvParams& p) {
    if (p.Cin % TC_BK) return qmri_fail(QMRI_EINVAL, "conv3x3_tc: Cin %% 64");
    const int BN = tc_block_n(p.Cout);
    if (p.Cout % BN) return qmri_fail(QMRI_EINVAL, "conv3x3_tc: Cout %% %d", BN);
    const int total = p.tiles_x * p.tiles_y * (p.Cout / BN) * p.S;
    const int grid = total < ctx->sm_count ? total : ctx->sm_count;
    const CUtensorMap& a_hi = *(const CUtensorMap*)mapA_hi;
    const CUtensorMap& a_lo = *(const CUtensorMap*)mapA_lo;
    const CUtensorMap& b_hi = *(const CUtensorMap*)mapB_hi;
    const CUtensorMap& b_lo = *(const CUtensorMap*)mapB_lo;
    static bool cfg64 = false, cfg128 = false;
    if (BN == 64) {
        if (!cfg64) {
            QCUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<64>::SMEM));
            cfg64 = true;
        }
        conv3x3_tc_kernel<64><<<grid, TC_THREADS, TcCfg<64>::SMEM, ctx->stream>>>(a_hi, a_lo, b_hi, b_lo, p);
    } else {
        if (!cfg128) {
            QCUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<128>::SMEM));
            cfg128 = true;
        }
        conv3x3_tc_kernel<128><<<grid, TC_THREADS, TcCfg<128>::SMEM, ctx->stream>>>(a_hi, a_lo, b_hi, b_lo, p);
    }
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
