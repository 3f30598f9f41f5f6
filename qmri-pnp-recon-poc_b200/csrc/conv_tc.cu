// K3 (tensor mode) - the DRUNet convolutions as tcgen05 implicit GEMMs with split-bf16 operands.
//
// Reference being replaced: the bias-free convolutions of UNetRes
//   PyTorch_Denoiser/zhang_dpir_testing_code/network_unet.py:106-117, basicblock.py:61-98,211-223
//   (58 x 3x3 s1 p1), :437-443 (3 x Conv2d 2x2 s2) and :413-419 (3 x ConvTranspose2d 2x2 s2).
//
// Precision: the parity bar is 1e-4 relative L2 against the fp32 CPU forward.  Single-pass TF32
// (5e-4) and bf16 (4e-3) operands fail it, so every fp32 value a is carried as
// a_hi = bf16(a), a_lo = bf16(a - a_hi) (16 significant bits) and a*w is evaluated as
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation in TMEM (three kind::f16 MMAs per
// K step; the dropped a_lo*w_lo term is 2^-18 relative).
//
// Data layout: activations are two bf16 planes (hi, lo), each [S][Y][X][C] (channels innermost,
// 2 B): same bytes as one fp32 tensor.  Weights are K-major rows [N][K], hi / lo planes.
//
// One kernel, three GEMM views (MODE):
//   0  3x3 s1 p1     M = pixels, N = Cout, K = 9 taps x Cin.  The A tile of tap (dy, dx) is the
//                    4-D TMA box of the pixel patch shifted by (dx, dy); out-of-image pixels are
//                    zero-filled by TMA, which is the conv's zero padding.
//   1  2x2 s2 conv   M = output pixels, N = Cout, K = 4 taps x Cin.  Tap (dy, dx) has its own
//                    tensor map: the input viewed with pixel strides 2 and base offset (dy, dx).
//   2  2x2 s2 convT  M = input pixels, N = 4 phases x Cout, K = Cin.  The epilogue scatters the
//                    N block of phase (dy, dx) to output pixel (2y + dy, 2x + dx).
//
// Kernel: persistent, warp-specialised.  GEMM tile = 128 pixels (a BH x BW patch of one slice) x BN
// columns.  Per k-block (one tap, 64 channels) the producer thread issues 4 TMA tile loads (A_hi,
// A_lo, B_hi, B_lo) into a 128B-swizzled shared-memory ring; one MMA thread issues 12 tcgen05.mma
// (M128 x BN x K16) per k-block into one of two TMEM accumulators; four epilogue warps drain the
// other accumulator with tcgen05.ld, add the residuals, apply ReLU, re-split to (hi, lo) and store.
//
// Split-K: when a layer has fewer tiles than SMs (deep U-Net levels at small slice batches) the K loop
// is divided over `nsplit` CTAs per tile.  Every CTA writes its fp32 partial tile to an L2-resident
// workspace; the last epilogue warp to arrive for a 32-row quarter of the tile (atomic ticket) sums
// the partials in split order (deterministic) and runs the normal epilogue.
//
// Measured on B200 (profiles/): the kernel is bound by L2 -> SM operand traffic (~12 TB/s chip-wide),
// not by the tensor pipe.  An A-tile "halo" variant that reused one shared-memory patch for all 9
// taps through 128-byte-offset (not 1024-byte aligned) UMMA descriptors gave correct results but ran
// its MMAs ~2x slower than aligned descriptors (with all loads disabled), and was removed.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "common.cuh"
#include "conv_kernels.h"

namespace {

constexpr int TC_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int PAIR_THREADS = 224;  // the CTA-pair kernel adds warp 6: a second MMA issuer (see tc_conv3x3_pair_kernel)
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
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// Programmatic dependent launch: the next conv kernel of the stream may be scheduled - and run its prologue (barrier init, TMEM
// allocation, descriptor prefetch) on SMs that are free - while this one is still running; it must not touch memory written by its
// predecessors before pdl_wait() returns (= every earlier kernel has completed and flushed).  QMRI_NO_PDL=1 launches without it.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// warp-converged forms (whole warp executes, one elected lane issues): see tc2_mma_elect below for why
__device__ __forceinline__ void tc_mma_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}

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


struct TcMaps {
    CUtensorMap a_hi[4], a_lo[4];  // MODE 1 uses one pair per tap, the others only [0]
    CUtensorMap b_hi, b_lo;
};

// device-side copy of TcConvParams (without the host tensor-map pointers)
struct TcK {
    uint16_t *out_hi, *out_lo;
    const uint16_t *res1_hi, *res1_lo, *res2_hi, *res2_lo;
    float* partial;
    int* tickets;
    int* trace;  // QMRI_TC_TRACE: host-mapped progress words, 8 per CTA (debugging hangs of the warp-specialised roles)
    int S, H, W, Cin, Cout;
    int BW, BH, tiles_x, tiles_y;
    int relu, nsplit;
    int cluster_splitk;  // split-K partials meet in the shared memory of a thread-block cluster of nsplit CTAs (one work item per CTA)
    int dual;            // CTA-pair kernel, Cout = 64: a second warp issues the a_lo w_hi MMAs into their own accumulator columns
    int resw;            // CTA-pair kernel, Cin = Cout = 64: the layer's weights stay resident in shared memory
    int tma_out;         // CTA-pair kernel, resw layers: the output tile leaves through shared memory + bulk tensor stores
    int l2pf;            // CTA-pair kernel: the producer prefetches the next work item's slabs into L2
    int res_tma;         // ... and the ResBlock residual tile arrives in the same staging tile through a bulk tensor load
    unsigned long long* prof;  // QMRI_TC_PROF: per-CTA cycle counters of the pair kernel's roles (8 per CTA), else null
};

// role-level attribution (QMRI_TC_PROF=1): cycles a role spends inside a wait, accumulated per CTA
#define TC_PROF_T0() const long long prof_t0 = p.prof ? clock64() : 0
#define TC_PROF_ADD(var) do { if (p.prof) (var) += (unsigned long long)(clock64() - prof_t0); } while (0)
#define TC_PROF_PUT(slot, val) do { if (p.prof) p.prof[blockIdx.x * 16 + (slot)] = (unsigned long long)(val); } while (0)

#define TC_TRACE(slot, val)                                                   \
    do {                                                                      \
        if (p.trace) {                                                        \
            ((volatile int*)p.trace)[blockIdx.x * 8 + (slot)] = (int)(val);   \
            __threadfence_system();                                           \
        }                                                                     \
    } while (0)

// v[0..8) += hi + lo (8 bf16 pairs)
__device__ __forceinline__ void add_split8(float* v, const uint4 a, const uint4 b) {
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float h0, h1, l0, l1;
        unpack_bf16(ua[j], h0, h1);
        unpack_bf16(ub[j], l0, l1);
        v[2 * j] += h0 + l0;
        v[2 * j + 1] += h1 + l1;
    }
}

// ReLU, split to (hi, lo) and store 16 consecutive channels
__device__ __forceinline__ void store_split16(uint16_t* oh, uint16_t* ol, const float* v, int relu) {
    uint32_t ph[8], pl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a = v[2 * j], b = v[2 * j + 1];
        if (relu) {
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
        }
        ph[j] = pack_bf16(a, b);
        float ha, hb;
        unpack_bf16(ph[j], ha, hb);
        pl[j] = pack_bf16(a - ha, b - hb);
    }
    uint4* o4h = reinterpret_cast<uint4*>(oh);
    uint4* o4l = reinterpret_cast<uint4*>(ol);
    o4h[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    o4h[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
    o4l[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    o4l[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
}

// ReLU, split to (hi, lo) and park 16 consecutive channels (two 16-byte chunks per plane) of tile row `r` in a shared-memory
// tile of 128-byte rows laid out as TMA SWIZZLE_128B expects (chunk j of row r at position j ^ (r & 7)): conflict-free for a
// warp (8 consecutive rows hit 8 different chunk columns), and one bulk tensor store then writes whole 128-byte lines.
__device__ __forceinline__ void store_split16_smem(uint32_t hi_row, uint32_t lo_row, int chunk0, int r7, const float* v, int relu) {
    uint32_t ph[8], pl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a = v[2 * j], b = v[2 * j + 1];
        if (relu) {
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
        }
        ph[j] = pack_bf16(a, b);
        float ha, hb;
        unpack_bf16(ph[j], ha, hb);
        pl[j] = pack_bf16(a - ha, b - hb);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t off = (uint32_t)(((chunk0 + h) ^ r7) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_row + off), "r"(ph[4 * h]), "r"(ph[4 * h + 1]), "r"(ph[4 * h + 2]), "r"(ph[4 * h + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_row + off), "r"(pl[4 * h]), "r"(pl[4 * h + 1]), "r"(pl[4 * h + 2]), "r"(pl[4 * h + 3]) : "memory");
    }
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// Epilogue of one 128 x BN tile (no split-K).  This thread owns one output pixel (TMEM lane) and BN
// channels.  The ResBlock residual of the pixel (BN hi + BN lo bf16 = BN/4 uint4) is fetched into
// registers BEFORE waiting for the accumulator, so its global-memory latency hides behind the tile's
// MMA main loop; the accumulator is then drained 32 columns at a time (two tcgen05.ld in flight).
template <int BN>
__device__ __forceinline__ void tc_epilogue_tile(const TcK& p, uint32_t taddr, size_t o, bool ok, uint64_t* tfull_bar, uint32_t parity) {
    uint4 rh[BN / 8], rl[BN / 8];
    const bool has_r1 = ok && (p.res1_hi != nullptr);
    if (has_r1) {
        const uint4* gh = reinterpret_cast<const uint4*>(p.res1_hi + o);
        const uint4* gl = reinterpret_cast<const uint4*>(p.res1_lo + o);
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
            rh[i] = gh[i];
            rl[i] = gl[i];
        }
    }
    mbar_wait(tfull_bar, parity);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t acc[2][16];
        tc_ld16(taddr + c0, acc[0]);
        tc_ld16(taddr + c0 + 16, acc[1]);
        tc_wait_ld();
        if (ok) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int cc = c0 + 16 * half;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[half][j]);
                if (has_r1) {
                    add_split8(v, rh[cc / 8], rl[cc / 8]);
                    add_split8(v + 8, rh[cc / 8 + 1], rl[cc / 8 + 1]);
                }
                if (p.res2_hi) {  // U-skip: only the last conv of a level
                    const uint4* gh = reinterpret_cast<const uint4*>(p.res2_hi + o + cc);
                    const uint4* gl = reinterpret_cast<const uint4*>(p.res2_lo + o + cc);
                    add_split8(v, gh[0], gl[0]);
                    add_split8(v + 8, gh[1], gl[1]);
                }
                store_split16(p.out_hi + o + cc, p.out_lo + o + cc, v, p.relu);
            }
        }
    }
}

// Split-K epilogue, phase 1: park this CTA's partial accumulator row in the workspace.
template <int BN>
__device__ __forceinline__ void tc_epilogue_park(uint32_t taddr, float* prow, uint64_t* tfull_bar, uint32_t parity) {
    mbar_wait(tfull_bar, parity);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t acc[2][16];
        tc_ld16(taddr + c0, acc[0]);
        tc_ld16(taddr + c0 + 16, acc[1]);
        tc_wait_ld();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4* d = reinterpret_cast<float4*>(prow + c0 + 16 * half);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                d[q] = make_float4(__uint_as_float(acc[half][4 * q]), __uint_as_float(acc[half][4 * q + 1]),
                                   __uint_as_float(acc[half][4 * q + 2]), __uint_as_float(acc[half][4 * q + 3]));
        }
    }
}

// Split-K epilogue, phase 2 (last arriver of the row quarter): sum the partials in split order and finish.
// The loop is latency-bound (L2 round trips), so one split's partials for 64 columns are fetched as 16 independent
// 128-bit loads before any is consumed; a per-chunk loop (4 loads in flight) was measured ~10 us per tile.
template <int BN>
__device__ __forceinline__ void tc_epilogue_reduce(const TcK& p, const float* prow0, size_t o) {
    const size_t split_stride = (size_t)TC_BM * BN;
    constexpr int CW = BN < 64 ? BN : 64;  // columns per pass
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CW) {
        float v[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = 0.f;
#pragma unroll 1
        for (int s = 0; s < p.nsplit; ++s) {
            const float4* src = reinterpret_cast<const float4*>(prow0 + (size_t)s * split_stride + c0);
            float4 t[CW / 4];
#pragma unroll
            for (int q = 0; q < CW / 4; ++q) t[q] = __ldcg(src + q);  // written by other SMs: read through L2
#pragma unroll
            for (int q = 0; q < CW / 4; ++q) {
                v[4 * q] += t[q].x;
                v[4 * q + 1] += t[q].y;
                v[4 * q + 2] += t[q].z;
                v[4 * q + 3] += t[q].w;
            }
        }
#pragma unroll
        for (int cc = 0; cc < CW; cc += 16) {
            if (p.res1_hi) {
                const uint4* gh = reinterpret_cast<const uint4*>(p.res1_hi + o + c0 + cc);
                const uint4* gl = reinterpret_cast<const uint4*>(p.res1_lo + o + c0 + cc);
                add_split8(v + cc, gh[0], gl[0]);
                add_split8(v + cc + 8, gh[1], gl[1]);
            }
            if (p.res2_hi) {
                const uint4* gh = reinterpret_cast<const uint4*>(p.res2_hi + o + c0 + cc);
                const uint4* gl = reinterpret_cast<const uint4*>(p.res2_lo + o + c0 + cc);
                add_split8(v + cc, gh[0], gl[0]);
                add_split8(v + cc + 8, gh[1], gl[1]);
            }
            store_split16(p.out_hi + o + c0 + cc, p.out_lo + o + c0 + cc, v + cc, p.relu);
        }
    }
}

// ---- split-K through distributed shared memory (cluster of nsplit CTAs = the K splits of one tile) ----------------------
// Phase 1: this CTA's partial accumulator -> its own shared memory, column-major [BN][128] floats (lanes = rows: conflict-free).
template <int BN>
__device__ __forceinline__ void tc_epilogue_park_smem(uint32_t taddr, float* red, int row, uint64_t* tfull_bar, uint32_t parity) {
    mbar_wait(tfull_bar, parity);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t acc[2][16];
        tc_ld16(taddr + c0, acc[0]);
        tc_ld16(taddr + c0 + 16, acc[1]);
        tc_wait_ld();
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int j = 0; j < 16; ++j) red[(c0 + 16 * half + j) * TC_BM + row] = __uint_as_float(acc[half][j]);
    }
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// Phase 2 (after a cluster barrier): CTA `rank` finishes the 16-column chunks rank, rank + nsplit, ...: partials are added in
// split order (the same order as the global-memory path: deterministic), then residuals / ReLU / split / store as usual.
template <int BN>
__device__ __forceinline__ void tc_epilogue_reduce_dsmem(const TcK& p, const float* red, int row, int rank, size_t o, bool ok) {
    const uint32_t mine = smem_u32(red);
#pragma unroll 1
    for (int c0 = 16 * rank; c0 < BN; c0 += 16 * p.nsplit) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll 1
        for (int s = 0; s < p.nsplit; ++s) {
            const uint32_t base = mapa_cluster(mine, (uint32_t)s) + (uint32_t)((c0 * TC_BM + row) * 4);
            float t[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) t[j] = ld_dsmem_f32(base + (uint32_t)(j * TC_BM * 4));
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += t[j];
        }
        if (ok) {
            if (p.res1_hi) {
                const uint4* gh = reinterpret_cast<const uint4*>(p.res1_hi + o + c0);
                const uint4* gl = reinterpret_cast<const uint4*>(p.res1_lo + o + c0);
                add_split8(v, gh[0], gl[0]);
                add_split8(v + 8, gh[1], gl[1]);
            }
            if (p.res2_hi) {
                const uint4* gh = reinterpret_cast<const uint4*>(p.res2_hi + o + c0);
                const uint4* gl = reinterpret_cast<const uint4*>(p.res2_lo + o + c0);
                add_split8(v, gh[0], gl[0]);
                add_split8(v + 8, gh[1], gl[1]);
            }
            store_split16(p.out_hi + o + c0, p.out_lo + o + c0, v, p.relu);
        }
    }
}

template <int BN, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(const __grid_constant__ TcMaps maps, const TcK p) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int TAPS = (MODE == 0) ? 9 : (MODE == 1) ? 4 : 1;
    extern __shared__ unsigned char tc_smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t raw = smem_u32(tc_smem_raw);
    unsigned char* smem = tc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;                     // [STAGES]
    uint64_t* empty = bars + STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;       // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // shuffle: tells ptxas the warp index is warp-uniform
    const int cblocks = p.Cin / TC_BK;
    const int KB = TAPS * cblocks;
    const int NT = ((MODE == 2) ? 4 * p.Cout : p.Cout) / BN;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int total_work = tiles_xy * NT * p.S * p.nsplit;  // work item = (tile, K split)

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
        tma_prefetch_desc(&maps.a_hi[0]);
        tma_prefetch_desc(&maps.a_lo[0]);
        tma_prefetch_desc(&maps.b_hi);
        tma_prefetch_desc(&maps.b_lo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();  // everything below reads activations / partials / tickets produced by earlier kernels
    size_t cs_o = 0;      // cluster split-K: this thread's output offset / validity / tile row (epilogue warps)
    bool cs_ok = false;
    int cs_row = 0;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int t = w / p.nsplit, split = w - t * p.nsplit;
                const int nt = t % NT;
                const int r = t / NT;
                const int txy = r % tiles_xy;
                const int s = r / tiles_xy;
                const int x0 = (txy % p.tiles_x) * p.BW, y0 = (txy / p.tiles_x) * p.BH;
                const int kb0 = (split * KB) / p.nsplit, kb1 = ((split + 1) * KB) / p.nsplit;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int tap = kb / cblocks, cb = kb - tap * cblocks;
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* st = smem + (size_t)stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    if (MODE == 0) {
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        tma_load_4d(st, &maps.a_hi[0], &full[stage], cb * TC_BK, x0 + dx, y0 + dy, s);
                        tma_load_4d(st + A_TILE_BYTES, &maps.a_lo[0], &full[stage], cb * TC_BK, x0 + dx, y0 + dy, s);
                    } else if (MODE == 1) {
                        tma_load_4d(st, &maps.a_hi[tap], &full[stage], cb * TC_BK, x0, y0, s);
                        tma_load_4d(st + A_TILE_BYTES, &maps.a_lo[tap], &full[stage], cb * TC_BK, x0, y0, s);
                    } else {
                        tma_load_4d(st, &maps.a_hi[0], &full[stage], cb * TC_BK, x0, y0, s);
                        tma_load_4d(st + A_TILE_BYTES, &maps.a_lo[0], &full[stage], cb * TC_BK, x0, y0, s);
                    }
                    tma_load_2d(st + 2 * A_TILE_BYTES, &maps.b_hi, &full[stage], kb * TC_BK, nt * BN);
                    tma_load_2d(st + 2 * A_TILE_BYTES + Cfg::B_TILE_BYTES, &maps.b_lo, &full[stage], kb * TC_BK, nt * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (the whole warp runs the loop, one elected lane issues: see tc2_mma_elect) =================
        {
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
                const int t = w / p.nsplit, split = w - t * p.nsplit;
                const int kb0 = (split * KB) / p.nsplit, kb1 = ((split + 1) * KB) / p.nsplit;
                const int ab = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[ab], aphase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::STAGE_BYTES);
                    const uint64_t a_hi = umma_desc(sa), a_lo = umma_desc(sa + A_TILE_BYTES);
                    const uint64_t b_hi = umma_desc(sa + 2 * A_TILE_BYTES), b_lo = umma_desc(sa + 2 * A_TILE_BYTES + Cfg::B_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);  // 32 bytes along K inside the swizzle atom (16-byte units)
                        tc_mma_elect(tmem_d, a_hi + ko, b_hi + ko, idesc, ((kb - kb0) | k) ? 1u : 0u);
                        tc_mma_elect(tmem_d, a_hi + ko, b_lo + ko, idesc, 1u);
                        tc_mma_elect(tmem_d, a_lo + ko, b_hi + ko, idesc, 1u);
                    }
                    tc_commit_elect(&empty[stage]);  // frees the smem slot once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit_elect(&tfull[ab]);  // accumulator complete
            }
        }
    } else {
        // ================= epilogue warps (TMEM lanes (warp % 4) * 32 ..) =================
        const int lg = warp & 3;
        const int row = lg * 32 + lane;  // row of the 128-pixel tile
        cs_row = row;
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
            const int t = w / p.nsplit, split = w - t * p.nsplit;
            const int ab = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int nt = t % NT;
            const int r = t / NT;
            const int txy = r % tiles_xy;
            const int s = r / tiles_xy;
            const int x = (txy % p.tiles_x) * p.BW + row % p.BW;
            const int y = (txy / p.tiles_x) * p.BH + row / p.BW;
            const bool ok = (x < p.W) && (y < p.H);
            size_t o;
            if (MODE == 2) {  // N block nt belongs to output phase (dy, dx); scatter to (2y + dy, 2x + dx)
                const int n0 = nt * BN;
                const int ph = n0 / p.Cout, co0 = n0 - ph * p.Cout;
                o = (((size_t)s * 2 * p.H + 2 * y + (ph >> 1)) * (2 * p.W) + 2 * x + (ph & 1)) * p.Cout + co0;
            } else {
                o = (((size_t)s * p.H + y) * p.W + x) * p.Cout + (size_t)nt * BN;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ab * BN);
            if (p.nsplit == 1) {
                tc_epilogue_tile<BN>(p, taddr, o, ok, &tfull[ab], aphase);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[ab]);
            } else if (p.cluster_splitk) {
                // one work item per CTA: the pipeline stages are idle once the accumulator is complete and stage 0 (64 KB) takes
                // the partial tile; the sum happens after the cluster barrier below
                tc_epilogue_park_smem<BN>(taddr, reinterpret_cast<float*>(smem), row, &tfull[ab], aphase);
                cs_o = o;
                cs_ok = ok;
            } else {
                float* prow0 = p.partial + ((size_t)t * p.nsplit * TC_BM + row) * BN;  // split 0's row of this tile
                tc_epilogue_park<BN>(taddr, prow0 + (size_t)split * TC_BM * BN, &tfull[ab], aphase);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[ab]);
                __threadfence();
                __syncwarp();
                int old = 0;
                if (lane == 0) old = atomicAdd(&p.tickets[t * 4 + lg], 1);
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old == p.nsplit - 1) {  // every other split of this row quarter is parked
                    __threadfence();
                    if (lane == 0) p.tickets[t * 4 + lg] = 0;  // ready for the next launch
                    if (ok) tc_epilogue_reduce<BN>(p, prow0, o);
                }
            }
        }
    }
    if (p.cluster_splitk) {
        // every thread of every CTA of the cluster: partial tiles are in shared memory -> sum -> nobody leaves before all reads are done
        __syncwarp();
        cluster_sync_all();
        if (warp >= 2) tc_epilogue_reduce_dsmem<BN>(p, reinterpret_cast<const float*>(smem), cs_row, (int)cluster_ctarank(), cs_o, cs_ok);
        __syncwarp();
        cluster_sync_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
    }
}


// ------------------------------------------------------------------------------------------------
// 3x3 conv, CTA-pair variant (tcgen05 cta_group::2) - the production path for the 58 3x3 convs.
//
// Measured with the single-CTA kernel above (profiles/): every MMA re-reads its A (4 KB) and B
// (N x 32 B) operands from shared memory, and one SM feeds its tensor core at ~64 B/clk, so
// M128 x N128 x K16 MMAs run at ~132 clk instead of 64 - half rate, whatever the L2 traffic (a
// multicast / A-reuse variant with half the L2 traffic ran at exactly the same speed).  The fix
// is the CTA pair: two CTAs (two SMs of a TPC) execute one M256 MMA; each supplies its own 128
// pixel rows of A and HALF of the B columns, so per-SM operand traffic per flop halves.
//
// Work per pair: two pixel tiles (one per CTA) x one block of output channels.
//   STACK = 0 (Cout >= 256): N = 256 output channels; three MMAs per K step
//             (a_hi b_hi, a_hi b_lo, a_lo b_hi), CTA r holds weight rows [128 r, 128 r + 128) of both planes.
//   STACK = 1 (Cout = NA/2 = 128 or 64): the hi and lo weight planes are stacked along N, B1 = [w_hi ; w_lo]:
//             MMA 1 = a_hi x B1 (N = NA; CTA 0 holds w_hi, CTA 1 holds w_lo) leaves a_hi w_hi in
//             accumulator columns [0, Cout) and a_hi w_lo in [Cout, 2 Cout); MMA 2 = a_lo x w_hi (N = Cout)
//             adds into [0, Cout); the epilogue sums the two column blocks.  Two MMAs instead of three.
// A operand: "slab" reuse - for every (64-channel block, dx) the (BH + 2) x BW pixel slab shifted by dx is
// loaded once; BW is a multiple of 8 pixels (8 x 128 B = one swizzle atom), so the A operand of tap dy is
// the same slab (dy + 1) * BW rows further, still 1024-byte aligned (descriptors that start inside a
// swizzle atom were measured ~2x slower).
// Barriers: full (TMA -> MMA) and tempty (epilogue -> MMA) live in the leader CTA, the peer's TMA and
// epilogue warps signal them remotely; empty / tfull are per CTA and are signalled by multicast commits.
// ------------------------------------------------------------------------------------------------
template <int NA, int STACK>
struct PairCfg {
    static constexpr int NOUT = STACK ? NA / 2 : NA;             // output channels per tile
    static constexpr uint32_t A_PLANE = 20480;                   // (8 + 2) x 16 rows of 128 B: the larger of the two tile shapes
    static constexpr uint32_t A_SLOT = 2 * A_PLANE;              // hi + lo
    static constexpr uint32_t B1_BYTES = (NA / 2) * 128;         // this CTA's half of the N = NA operand
    static constexpr uint32_t B2_BYTES = STACK ? (NA / 4) * 128 : B1_BYTES;  // STACK: half of the N = NA/2 operand; else the lo plane
    static constexpr uint32_t B_STAGE = B1_BYTES + B2_BYTES;
    static constexpr int AS = (NA == 128) ? 3 : 2;
    static constexpr int ASB = 4;                                // slab-stage barrier slots (resident-weight mode runs up to four stages)
    static constexpr int BS = (NA == 128) ? 8 : (STACK ? 5 : 4);
    static constexpr bool DUAL_OK = STACK && NA == 128;          // the 64 -> 64 instance (lean epilogue, staging tile, store thread)
    static constexpr int EPW = DUAL_OK ? 8 : 4;                  // epilogue warps: 2 .. 5 (+ 7 .. 10: the second 32 channels of each lane quarter)
    static constexpr int THREADS = DUAL_OK ? 352 : PAIR_THREADS;
    static constexpr uint32_t TMEM_COLS = DUAL_OK ? 512 : 2 * NA; // two accumulators (+ 2 x NA/2 columns for the second issuer's)
    static constexpr uint32_t RESW_BYTES = 9 * B1_BYTES;         // resident-weight mode: nine taps of this CTA's 64 stacked weight rows
    static constexpr size_t STAGE_AREA = (size_t)AS * A_SLOT + (size_t)BS * B_STAGE;
    static constexpr size_t SMEM = STAGE_AREA + 1024 + 256;
};

__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// accumulator hand-back: nothing this thread WROTE has to become visible with the arrival (the TMEM reads are ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync), so the arrival is relaxed - the release form cost ~1000 cycles per tile
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile loads of a CTA pair: data lands in this CTA, completion is signalled on `bar_cluster_addr` (the leader's barrier)
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// bring a tile into L2 ahead of its shared-memory load (no shared memory, no barrier: pipeline depth that costs nothing on chip)
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc2_commit(uint64_t* bar) {  // arrive on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Warp-converged forms: the WHOLE warp runs the issue loop and one elected lane executes the instruction.  With the loop inside
// `if (lane == 0)` ptxas cannot keep descriptors / addresses in uniform registers and wraps every UTCHMMA in an ELECT + 5 x
// R2UR.BROADCAST + BRA.U.ANY "waterfall" loop - that loop, not the hardware, was the ~103-cycle-per-MMA "issue floor" of
// profiles/r02_tmem_a_bench.txt (profiles/r02u_pair_roles.md).
__device__ __forceinline__ void tc2_mma_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc2_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}

// 16 consecutive output channels of this thread's pixel from the accumulator (STACK: sum of the two column blocks)
// resw (resident-weight layout of the 64 -> 64 layers): the accumulator columns are [hi(0-31) | lo(32-63) | hi(32-63) | lo(0-31)]
// (+ a_lo w_hi(c) added onto column c), so the second term of channel c sits at column c + 96 (c < 32) or c + 32 (c >= 32)
template <int NA, int STACK>
__device__ __forceinline__ void pair_acc16(uint32_t taddr, uint32_t taddr2, int c0, float* v, bool resw = false) {
    uint32_t a[16];
    tc_ld16(taddr + c0, a);
    if (STACK) {
        uint32_t b[16];
        tc_ld16(taddr + (resw ? (c0 < NA / 4 ? 3 * NA / 4 : NA / 4) : NA / 2) + c0, b);
        if (taddr2) {  // dual-issuer mode: a_lo w_hi was accumulated in its own columns
            uint32_t c[16];
            tc_ld16(taddr2 + c0, c);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (__uint_as_float(a[j]) + __uint_as_float(c[j])) + __uint_as_float(b[j]);
        } else {
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]) + __uint_as_float(b[j]);
        }
    } else {
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]);
    }
}

template <int NA, int STACK>
__global__ void __launch_bounds__(PairCfg<NA, STACK>::THREADS, 1)
tc_conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                       const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                       const __grid_constant__ CUtensorMap tmB_h2, const __grid_constant__ CUtensorMap tmB_l2,
                       const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                       const __grid_constant__ CUtensorMap tmR_hi, const __grid_constant__ CUtensorMap tmR_lo, const TcK p) {
    using Cfg = PairCfg<NA, STACK>;
    constexpr int AS = Cfg::AS, BS = Cfg::BS, NOUT = Cfg::NOUT;
    extern __shared__ unsigned char tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    unsigned char* smem = tc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* smemB = smem + (size_t)AS * Cfg::A_SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + (size_t)BS * Cfg::B_STAGE);
    constexpr int ASB = Cfg::ASB;
    uint64_t* fullA = bars;               // leader's are used
    uint64_t* emptyA = bars + ASB;        // per CTA
    uint64_t* fullB = bars + 2 * ASB;     // leader's
    uint64_t* emptyB = bars + 2 * ASB + BS;
    uint64_t* tfull = bars + 2 * ASB + 2 * BS;  // per CTA
    uint64_t* tempty = tfull + 2;              // leader's
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // shuffle: tells ptxas the warp index is warp-uniform
    const int cblocks = p.Cin / TC_BK;
    const int U = 3 * cblocks;                                // K units: (channel block, dx); each = 1 slab + 3 weight tiles
    const int NT = p.Cout / NOUT;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int ptiles = tiles_xy * p.S;                        // pixel tiles
    const int pgroups = (ptiles + 1) / 2;
    const int total_work = pgroups * NT * p.nsplit;           // per pair: (pixel-tile pair, channel block, K split)
    const int crank = (int)cluster_ctarank();
    const int cidx = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const uint32_t slab_bytes = (uint32_t)((p.BH + 2) * p.BW * 128);
    const uint32_t dy_bytes = (uint32_t)(p.BW * 128);
    // Dual-issuer mode (Cout = 64 layers) - RETIRED (launch_pair never sets it): a second warp issued a_lo x w_hi into its own accumulator
    // columns to get around a ~103-cycle "issue floor" that was ptxas' register-to-uniform loop (see tc2_mma_elect); the lean epilogue of
    // the 64 -> 64 instance no longer reads a third accumulator block.
    const bool dual = Cfg::DUAL_OK && (p.dual & 1);
    // Resident-weight mode (Cin = Cout = 64 layers).  With Cout = 64 a 128-pixel tile re-streamed the whole layer's weights (108 KB per
    // CTA: nine taps of [w_hi ; w_lo] and w_hi) next to 120 KB of activation slabs - 66 B/clk per SM at the tensor rate against the
    // ~43 B/clk an SM gets from L2 through TMA.  Here the nine weight tiles are loaded ONCE per CTA and the main loop streams activation
    // slabs only.  Per tap a CTA keeps 64 rows: CTA 0 [w_hi(0-31) ; w_lo(32-63)], CTA 1 [w_hi(32-63) ; w_lo(0-31)].  The N = 128 MMA
    // a_hi x B then leaves [hi(0-31) | lo(32-63) | hi(32-63) | lo(0-31)] in the accumulator columns, and the N = 64 MMA a_lo x (the
    // FIRST 32 rows of each CTA, same descriptor) adds a_lo w_hi(c) onto column c for all 64 channels - no second copy of w_hi, 72 KB
    // instead of 108 KB, which leaves room for three or four slab stages (two were measured to starve the MMA warp: a 40 KB stage is
    // consumed in ~1400 cycles, less than its fetch takes under load).
    const bool resw = Cfg::DUAL_OK && p.resw && p.Cin == TC_BK;
    const uint32_t a_stage = resw ? 2 * slab_bytes : Cfg::A_SLOT;     // bytes per slab stage (hi + lo)
    const uint32_t a_plane = resw ? slab_bytes : Cfg::A_PLANE;        // offset of the lo plane inside a stage
    // resw + p.tma_out: the finished tile (128 pixel rows x 64 channels, hi and lo planes, 32 KB) is parked in shared memory and
    // written with two bulk tensor stores - 16-byte stores of 32 lanes to 32 different lines made the epilogue the bound of the
    // 64 -> 64 layers (one L2 request per lane and instruction: ~6400 cycles per tile, profiles/r02u_pair_roles.md)
    // The staging tile is double-buffered, and for a layer with a ResBlock residual the producer warp first fills it with the
    // residual tile (bulk tensor load): the epilogue reads its own row, adds, and writes the result back in place.
    const bool tma_out = resw && p.tma_out;
    const bool res_tma = tma_out && p.res_tma;
    constexpr uint32_t OUT_BYTES = 2 * TC_BM * 128;
    const int obufs = p.tma_out == 1 ? 1 : 2;   // staging buffers: one leaves room for a third slab stage (host's choice per layer)
    const int as_fit = (int)((Cfg::STAGE_AREA - Cfg::RESW_BYTES - (tma_out ? obufs * OUT_BYTES : 0u)) / (2 * slab_bytes));
    uint64_t* fullR = fullB + 2;   // [2]; resw mode uses fullB[0] only, so the weight-ring barriers (count 1) are free
    uint64_t* emptyR = fullB + 4;  // [2]
    uint64_t* outFull = fullB + 6; // [2], count 4: the epilogue warps have written the finished tile into the staging buffer
    const int as_cap = (p.resw >= 2 && p.resw < Cfg::ASB) ? p.resw : Cfg::ASB;   // QMRI_TC_RESW=2 / 3: fewer stages (A/B measurements)
    const int AS_EFF = resw ? (as_fit < as_cap ? as_fit : as_cap) : AS;
    unsigned char* smemW = smem;                                      // resw: nine taps x 64 rows x 128 B
    unsigned char* smemO = smem + Cfg::RESW_BYTES;                    // tma_out: two staging tiles, each hi plane then lo plane
    unsigned char* smemA = resw ? smem + Cfg::RESW_BYTES + (tma_out ? obufs * OUT_BYTES : 0u) : smem;      // slab stages

    if (threadIdx.x == 0) {
        const uint32_t nissue = dual ? 2u : 1u;  // MMA issuers that commit on the stage / accumulator barriers
        for (int s = 0; s < ASB; ++s) {
            mbar_init(&fullA[s], 1);
            mbar_init(&emptyA[s], nissue);
        }
        for (int s = 0; s < BS; ++s) {
            mbar_init(&fullB[s], 1);
            mbar_init(&emptyB[s], nissue);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], nissue);
            mbar_init(&tempty[a], 2 * Cfg::EPW);  // the epilogue warps of both CTAs
            if (tma_out) mbar_init(&outFull[a], Cfg::EPW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmA_lo);
        tma_prefetch_desc(&tmB_hi);
        tma_prefetch_desc(&tmB_lo);
        tma_prefetch_desc(&tmB_h2);
        tma_prefetch_desc(&tmB_l2);
        tma_prefetch_desc(&tmO_hi);
        tma_prefetch_desc(&tmO_lo);
        tma_prefetch_desc(&tmR_hi);
        tma_prefetch_desc(&tmR_lo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TC_TRACE(7, -1);
    cluster_sync_all();  // both CTAs' barriers are initialised and TMEM is allocated before any cross-CTA signal
    if (threadIdx.x == 0) TC_TRACE(7, -2);
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();  // everything below reads activations / partials / tickets produced by earlier kernels

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            unsigned long long prof_w5 = 0;
            if (resw && cidx < total_work) {  // the layer's weights, once: completion on the leader's fullB[0]
                if (crank == 0) mbar_expect_tx(&fullB[0], 2 * Cfg::RESW_BYTES);
                const uint32_t barW = mapa_rank(smem_u32(&fullB[0]), 0);
                for (int tap = 0; tap < 9; ++tap) {
                    unsigned char* wp = smemW + (size_t)tap * Cfg::B1_BYTES;
                    tma2_load_2d(wp, &tmB_h2, barW, tap * p.Cin, crank * (NOUT / 2));
                    tma2_load_2d(wp + Cfg::B1_BYTES / 2, &tmB_l2, barW, tap * p.Cin, (1 - crank) * (NOUT / 2));
                }
            }
            for (int w = cidx; w < total_work; w += nclusters) {
                const int split = w % p.nsplit;
                const int r = w / p.nsplit;
                const int nt = r % NT;
                const int pt = (r / NT) * 2 + crank;
                // a rank past the last pixel tile still supplies its half of the weights; its slab is all out of bounds (zeros)
                const int txy = pt % tiles_xy;
                const int s = (pt < ptiles) ? pt / tiles_xy : p.S;
                const int x0 = (txy % p.tiles_x) * p.BW, y0 = (txy / p.tiles_x) * p.BH;
                const int u0 = (split * U) / p.nsplit, u1 = ((split + 1) * U) / p.nsplit;
                if (p.l2pf && w + nclusters < total_work) {  // the NEXT work item's slabs into L2 (the layer input streams from DRAM)
                    const int w2 = w + nclusters;
                    const int r2 = w2 / p.nsplit, split2 = w2 % p.nsplit;
                    const int pt2 = (r2 / NT) * 2 + crank;
                    if (pt2 < ptiles) {
                        const int txy2 = pt2 % tiles_xy, s2 = pt2 / tiles_xy;
                        const int px0 = (txy2 % p.tiles_x) * p.BW, py0 = (txy2 / p.tiles_x) * p.BH;
                        const int v0 = (split2 * U) / p.nsplit, v1 = ((split2 + 1) * U) / p.nsplit;
                        for (int u = v0; u < v1; ++u) {
                            const int cb = u / 3, dxi = u - 3 * cb;
                            tma_prefetch_l2_4d(&tmA_hi, cb * TC_BK, px0 + dxi - 1, py0 - 1, s2);
                            tma_prefetch_l2_4d(&tmA_lo, cb * TC_BK, px0 + dxi - 1, py0 - 1, s2);
                        }
                    }
                }
                for (int u = u0; u < u1; ++u) {
                    const int cb = u / 3, dxi = u - 3 * cb;
                    TC_TRACE(0, (w << 12) | (u << 4) | 1);
                    {
                        TC_PROF_T0();
                        mbar_wait(&emptyA[sa], pa ^ 1);
                        TC_PROF_ADD(prof_w5);
                    }
                    unsigned char* st = smemA + (size_t)sa * a_stage;
                    if (p.dual & 256) {  // timing experiment: no slab traffic, the stage is declared full at once
                        if (crank == 0) mbar_arrive(&fullA[sa]);
                    } else {
                    if (crank == 0) mbar_expect_tx(&fullA[sa], 4 * slab_bytes);  // both CTAs' hi + lo slabs
                    const uint32_t barA = mapa_rank(smem_u32(&fullA[sa]), 0);
                    tma2_load_4d(st, &tmA_hi, barA, cb * TC_BK, x0 + dxi - 1, y0 - 1, s);
                    tma2_load_4d(st + a_plane, &tmA_lo, barA, cb * TC_BK, x0 + dxi - 1, y0 - 1, s);
                    }
                    if (++sa == AS_EFF) {
                        sa = 0;
                        pa ^= 1;
                    }
                    for (int dyi = 0; dyi < 3 && !resw; ++dyi) {
                        const int kcol = (dyi * 3 + dxi) * p.Cin + cb * TC_BK;
                        TC_TRACE(0, (w << 12) | (u << 4) | (2 + dyi));
                        mbar_wait(&emptyB[sb], pb ^ 1);
                        unsigned char* sbp = smemB + (size_t)sb * Cfg::B_STAGE;
                        if (crank == 0) mbar_expect_tx(&fullB[sb], 2 * Cfg::B_STAGE);
                        const uint32_t barB = mapa_rank(smem_u32(&fullB[sb]), 0);
                        if (STACK) {
                            // B1 = [w_hi ; w_lo] (N = NA): CTA 0 holds the hi plane, CTA 1 the lo plane; B2 = w_hi (N = NOUT), split by rows
                            tma2_load_2d(sbp, crank == 0 ? &tmB_hi : &tmB_lo, barB, kcol, 0);
                            tma2_load_2d(sbp + Cfg::B1_BYTES, &tmB_h2, barB, kcol, crank * (NOUT / 2));
                        } else {
                            tma2_load_2d(sbp, &tmB_hi, barB, kcol, nt * NOUT + crank * (NA / 2));
                            tma2_load_2d(sbp + Cfg::B1_BYTES, &tmB_lo, barB, kcol, nt * NOUT + crank * (NA / 2));
                        }
                        if (++sb == BS) {
                            sb = 0;
                            pb ^= 1;
                        }
                    }
                }
            }
            TC_PROF_PUT(5, prof_w5);
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only; the whole warp runs the loop, one elected lane issues) =================
        if (crank == 0) {
            // instruction descriptors: D = f32, A = B = bf16, K-major, M = 256
            const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t idesc_full = idesc_base | ((uint32_t)(NA >> 3) << 17);
            const uint32_t idesc_half = idesc_base | ((uint32_t)((NA / 2) >> 3) << 17);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            const long long role_t0 = p.prof ? clock64() : 0;
            unsigned long long role_ns0 = 0;
            if (p.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(role_ns0));
            unsigned long long prof_w0 = 0, prof_w1 = 0, prof_n = 0, prof_w8 = 0, prof_w9 = 0;
            if (resw && cidx < total_work) {
                mbar_wait(&fullB[0], 0);  // the resident weights have landed in both CTAs
                tc_fence_after();
            }
            for (int w = cidx; w < total_work; w += nclusters, ++it) {
                const int split = w % p.nsplit;
                const int u0 = (split * U) / p.nsplit, u1 = ((split + 1) * U) / p.nsplit;
                const int ab = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                if (lane == 0) TC_TRACE(1, (w << 12) | 0xF00);
                {
                    TC_PROF_T0();
                    mbar_wait(&tempty[ab], aphase ^ 1);
                    TC_PROF_ADD(prof_w1);
                    ++prof_n;
                }
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * NA);
                for (int u = u0; u < u1; ++u) {
                    if (lane == 0) TC_TRACE(1, (w << 12) | (u << 4) | 1);
                    {
                        TC_PROF_T0();
                        mbar_wait(&fullA[sa], pa);
                        TC_PROF_ADD(prof_w0);
                    }
                    tc_fence_after();
                    const uint32_t ha = smem_u32(smemA + (size_t)sa * a_stage);
                    const long long issue_t0 = p.prof ? clock64() : 0;
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        if (lane == 0) TC_TRACE(1, (w << 12) | (u << 4) | (2 + dyi));
                        if (!resw) {
                            mbar_wait(&fullB[sb], pb);
                            tc_fence_after();
                        }
                        if (p.dual & 16) continue;  // timing experiment (QMRI_TC_DUAL=16): no MMAs at all
                        const uint32_t sbase = resw ? smem_u32(smemW + (size_t)(dyi * 3 + (u - 3 * (u / 3))) * Cfg::B1_BYTES)
                                                    : smem_u32(smemB + (size_t)sb * Cfg::B_STAGE);
                        const uint64_t a_hi = umma_desc(ha + dyi * dy_bytes), a_lo = umma_desc(ha + a_plane + dyi * dy_bytes);
                        const uint64_t b_1 = umma_desc(sbase), b_2 = umma_desc(resw ? sbase : sbase + Cfg::B1_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {
                            const uint64_t ko = (uint64_t)(k * 2);
                            const uint32_t acc = ((u - u0) | dyi | k) ? 1u : 0u;
                            if (STACK) {
                                tc2_mma_elect(tmem_d, a_hi + ko, b_1 + ko, idesc_full, acc);   // [a_hi w_hi | a_hi w_lo]
                                if (!dual) tc2_mma_elect(tmem_d, a_lo + ko, b_2 + ko, idesc_half, 1u);    // += a_lo w_hi into the first block
                            } else {
                                tc2_mma_elect(tmem_d, a_hi + ko, b_1 + ko, idesc_full, acc);
                                tc2_mma_elect(tmem_d, a_hi + ko, b_2 + ko, idesc_full, 1u);
                                tc2_mma_elect(tmem_d, a_lo + ko, b_1 + ko, idesc_full, 1u);
                            }
                        }
                        if (!resw) {
                            tc2_commit_elect(&emptyB[sb]);  // frees the stage in both CTAs
                            if (++sb == BS) {
                                sb = 0;
                                pb ^= 1;
                            }
                        }
                    }
                    if (p.prof) prof_w8 += (unsigned long long)(clock64() - issue_t0);
                    {
                        TC_PROF_T0();
                        tc2_commit_elect(&emptyA[sa]);
                        TC_PROF_ADD(prof_w9);
                    }
                    if (++sa == AS_EFF) {
                        sa = 0;
                        pa ^= 1;
                    }
                }
                {
                    TC_PROF_T0();
                    tc2_commit_elect(&tfull[ab]);  // accumulator complete: wakes the epilogue warps of both CTAs
                    TC_PROF_ADD(prof_w9);
                }
            }
            if (lane == 0) {
            TC_PROF_PUT(8, prof_w8);
            TC_PROF_PUT(9, prof_w9);
            TC_PROF_PUT(2, clock64() - role_t0);
            }
            if (p.prof && lane == 0) {
                unsigned long long role_ns1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(role_ns1));
                TC_PROF_PUT(6, role_ns1 - role_ns0);
            }
            if (lane == 0) {
            TC_PROF_PUT(0, prof_w0);
            TC_PROF_PUT(1, prof_w1);
            TC_PROF_PUT(7, prof_n);
            }
        }
    } else if (warp == 6) {
        // ================= store thread (64 -> 64 layers): finished tiles leave through bulk tensor stores =================
        if (lane == 0 && tma_out) {
            // residual tile of work item w2 into staging buffer b (bulk tensor load; the epilogue warps wait on fullR[b]).  Issued by
            // THIS thread because it knows when a buffer's previous store has been read out - a whole tile time before the
            // epilogue needs the residual - and the slab producer never waits for a staging buffer.
            auto load_residual = [&](int w2, int b) {
                if (!res_tma || w2 >= total_work) return;
                const int pt2 = ((w2 / p.nsplit) / NT) * 2 + crank;
                const int txy2 = pt2 % tiles_xy, s2 = (pt2 < ptiles) ? pt2 / tiles_xy : p.S;  // past the last tile: all out of bounds (zeros)
                const int rx0 = (txy2 % p.tiles_x) * p.BW, ry0 = (txy2 / p.tiles_x) * p.BH;
                mbar_expect_tx(&fullR[b], OUT_BYTES);
                unsigned char* rp = smemO + (size_t)b * OUT_BYTES;
                tma_load_4d(rp, &tmR_hi, &fullR[b], 0, rx0, ry0, s2);
                tma_load_4d(rp + TC_BM * 128, &tmR_lo, &fullR[b], 0, rx0, ry0, s2);
            };
            for (int b = 0; b < obufs; ++b) load_residual(cidx + b * nclusters, b);
            int it = 0;
            for (int w = cidx; w < total_work; w += nclusters, ++it) {
                const int pt = ((w / p.nsplit) / NT) * 2 + crank;
                const int txy = pt % tiles_xy, s = pt / tiles_xy;
                const int sbuf = obufs == 2 ? (it & 1) : 0;
                mbar_wait(&outFull[sbuf], (uint32_t)(obufs == 2 ? it >> 1 : it) & 1u);
                if (pt < ptiles && !(p.dual & 32)) {
                    const int tx0 = (txy % p.tiles_x) * p.BW, ty0 = (txy / p.tiles_x) * p.BH;
                    tma_store_4d(&tmO_hi, smem_u32(smemO) + (uint32_t)sbuf * OUT_BYTES, 0, tx0, ty0, s);
                    tma_store_4d(&tmO_lo, smem_u32(smemO) + (uint32_t)sbuf * OUT_BYTES + TC_BM * 128u, 0, tx0, ty0, s);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging buffer has been read out: hand it back
                if (res_tma) load_residual(w + obufs * nclusters, sbuf);   // ... to the residual tile of its next user
                else mbar_arrive(&emptyR[sbuf]);                          // ... or straight to the epilogue warps
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every bulk store of this CTA has completed
        }
        // ================= second MMA issuer (leader CTA only, dual-issuer mode): a_lo x w_hi into its own accumulator =================
        if (lane == 0 && crank == 0 && dual) {
            const uint32_t idesc_half = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 4) << 24) | ((uint32_t)((NA / 2) >> 3) << 17);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int w = cidx; w < total_work; w += nclusters, ++it) {
                const int split = w % p.nsplit;
                const int u0 = (split * U) / p.nsplit, u1 = ((split + 1) * U) / p.nsplit;
                const int ab = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                if (resw && it == 0) {
                    mbar_wait(&fullB[0], 0);
                    tc_fence_after();
                }
                mbar_wait(&tempty[ab], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d2 = tmem_base + (uint32_t)(2 * NA + ab * (NA / 2));
                for (int u = u0; u < u1; ++u) {
                    mbar_wait(&fullA[sa], pa);
                    tc_fence_after();
                    const uint32_t ha = smem_u32(smemA + (size_t)sa * a_stage);
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        if (!resw) {
                            mbar_wait(&fullB[sb], pb);
                            tc_fence_after();
                        }
                        const uint32_t sbase = resw ? smem_u32(smemW + (size_t)(dyi * 3 + (u - 3 * (u / 3))) * Cfg::B1_BYTES)
                                                    : smem_u32(smemB + (size_t)sb * Cfg::B_STAGE);
                        const uint64_t a_lo = umma_desc(ha + a_plane + dyi * dy_bytes);
                        const uint64_t b_2 = umma_desc(resw ? sbase : sbase + Cfg::B1_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {
                            const uint64_t ko = (uint64_t)(k * 2);
                            tc2_mma(tmem_d2, a_lo + ko, b_2 + ko, idesc_half, ((u - u0) | dyi | k) ? 1u : 0u);
                        }
                        if (!resw) {
                            tc2_commit(&emptyB[sb]);
                            if (++sb == BS) {
                                sb = 0;
                                pb ^= 1;
                            }
                        }
                    }
                    tc2_commit(&emptyA[sa]);
                    if (++sa == AS_EFF) {
                        sa = 0;
                        pa ^= 1;
                    }
                }
                tc2_commit(&tfull[ab]);
            }
        }
    } else {
        // ================= epilogue warps (both CTAs; each drains its own 128 accumulator rows) =================
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        int it = 0;
        const long long role_t0 = p.prof ? clock64() : 0;
        unsigned long long prof_w3 = 0, prof_w10 = 0, prof_w11 = 0, prof_w12 = 0, prof_w13 = 0;
        for (int w = cidx; w < total_work; w += nclusters, ++it) {
            const int split = w % p.nsplit;
            const int r = w / p.nsplit;
            const int ab = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int nt = r % NT;
            const int pt = (r / NT) * 2 + crank;
            const bool live = pt < ptiles;
            const int txy = pt % tiles_xy;
            const int s = pt / tiles_xy;
            const int x = (txy % p.tiles_x) * p.BW + row % p.BW;
            const int y = (txy / p.tiles_x) * p.BH + row / p.BW;
            const bool ok = live && (x < p.W) && (y < p.H);
            const size_t o = (((size_t)s * p.H + y) * p.W + x) * p.Cout + (size_t)nt * NOUT;
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ab * NA);
            const uint32_t taddr2 = dual ? tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(2 * NA + ab * (NA / 2)) : 0u;
            const uint32_t tempty_leader = mapa_rank(smem_u32(&tempty[ab]), 0);
            if (lane == 0) TC_TRACE(2 + lg, (w << 12) | 1);
            if (p.nsplit == 1) {
                if constexpr (Cfg::DUAL_OK) {
                    // ---- 64 -> 64 layers: lean epilogue.  The accumulator is drained 16 channels at a time with the tcgen05.ld of the
                    // next chunk in flight; the ResBlock residual comes from the staging tile (filled by the producer's bulk tensor
                    // load), the result goes back into the same row of that tile and leaves with two bulk tensor stores.  Without the
                    // tile maps (tma_out off) residuals and outputs use 16-byte global accesses (slow, kept for A/B runs).
                    const int sbuf = obufs == 2 ? (it & 1) : 0;
                    const uint32_t suse = (uint32_t)(obufs == 2 ? it >> 1 : it);
                    const uint32_t orow_hi = smem_u32(smemO) + (uint32_t)sbuf * OUT_BYTES + (uint32_t)row * 128u, orow_lo = orow_hi + TC_BM * 128u;
                    const int r7 = row & 7;
                    const bool g_r1 = ok && !res_tma && (p.res1_hi != nullptr), g_r2 = ok && (p.res2_hi != nullptr);
                    uint4 q2[4];  // second residual (U-skip; last conv of a level only): global, fetched one chunk ahead
                    if (g_r2) {
                        const int c0 = (Cfg::EPW == 8 && warp >= 7) ? NOUT / 2 : 0;
                        q2[0] = *reinterpret_cast<const uint4*>(p.res2_hi + o + c0);
                        q2[1] = *reinterpret_cast<const uint4*>(p.res2_hi + o + c0 + 8);
                        q2[2] = *reinterpret_cast<const uint4*>(p.res2_lo + o + c0);
                        q2[3] = *reinterpret_cast<const uint4*>(p.res2_lo + o + c0 + 8);
                    }
                    if (tma_out) {  // staging buffer: residual tile landed / previous bulk store of this buffer read out
                        TC_PROF_T0();
                        if (res_tma) mbar_wait(&fullR[sbuf], suse & 1u);
                        else mbar_wait(&emptyR[sbuf], (suse & 1u) ^ 1u);
                        TC_PROF_ADD(prof_w11);
                    }
                    {
                        TC_PROF_T0();
                        mbar_wait(&tfull[ab], aphase);
                        TC_PROF_ADD(prof_w3);
                    }
                    tc_fence_after();
                    const long long drain_t0 = p.prof ? clock64() : 0;
                    if (!(p.dual & 32)) {  // (QMRI_TC_DUAL=32, timing experiment: the epilogue only hands the accumulator back)
                        // eight epilogue warps: warps 2 - 5 take channels [0, 32) of their lane quarter, warps 7 - 10 channels [32, 64)
                        // (one warp per scheduler cannot hide its own ALU / LDTM latencies: ~5 cycles per instruction were measured)
                        constexpr int CPW = NOUT / 16 / (Cfg::EPW / 4);   // 16-channel chunks per warp
                        const int ch0 = (Cfg::EPW == 8 && warp >= 7) ? CPW : 0;
                        uint32_t ta[2][16], tb[2][16];
                        {
                            const int c0 = 16 * ch0;
                            tc_ld16(taddr + c0, ta[0]);
                            tc_ld16(taddr + (resw ? (c0 < NA / 4 ? 3 * NA / 4 : NA / 4) : NA / 2) + c0, tb[0]);
                        }
#pragma unroll
                        for (int ci = 0; ci < CPW; ++ci) {
                            const int ch = ch0 + ci;
                            const int cc = 16 * ch, cur = ci & 1;
                            {
                                TC_PROF_T0();
                                tc_wait_ld();
                                TC_PROF_ADD(prof_w10);
                            }
                            if (ci + 1 < CPW) {  // second term of channel c: column c + 96 (c < 32) or c + 32 (c >= 32), see pair_acc16
                                const int cn = cc + 16;
                                tc_ld16(taddr + cn, ta[cur ^ 1]);
                                tc_ld16(taddr + (resw ? (cn < NA / 4 ? 3 * NA / 4 : NA / 4) : NA / 2) + cn, tb[cur ^ 1]);
                            }
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(ta[cur][j]) + __uint_as_float(tb[cur][j]);
                            if (res_tma) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const uint32_t off = (uint32_t)((((cc >> 3) + h) ^ r7) << 4);
                                    uint4 rh, rl;
                                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rh.x), "=r"(rh.y), "=r"(rh.z), "=r"(rh.w) : "r"(orow_hi + off));
                                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rl.x), "=r"(rl.y), "=r"(rl.z), "=r"(rl.w) : "r"(orow_lo + off));
                                    add_split8(v + 8 * h, rh, rl);
                                }
                            } else if (g_r1) {
                                const uint4* gh = reinterpret_cast<const uint4*>(p.res1_hi + o + cc);
                                const uint4* gl = reinterpret_cast<const uint4*>(p.res1_lo + o + cc);
                                add_split8(v, gh[0], gl[0]);
                                add_split8(v + 8, gh[1], gl[1]);
                            }
                            if (g_r2) {
                                add_split8(v, q2[0], q2[2]);
                                add_split8(v + 8, q2[1], q2[3]);
                                if (ci + 1 < CPW) {
                                    q2[0] = *reinterpret_cast<const uint4*>(p.res2_hi + o + cc + 16);
                                    q2[1] = *reinterpret_cast<const uint4*>(p.res2_hi + o + cc + 24);
                                    q2[2] = *reinterpret_cast<const uint4*>(p.res2_lo + o + cc + 16);
                                    q2[3] = *reinterpret_cast<const uint4*>(p.res2_lo + o + cc + 24);
                                }
                            }
                            if (tma_out) store_split16_smem(orow_hi, orow_lo, cc >> 3, r7, v, p.relu);
                            else if (ok) store_split16(p.out_hi + o + cc, p.out_lo + o + cc, v, p.relu);
                        }
                    }
                    if (p.prof) prof_w12 += (unsigned long long)(clock64() - drain_t0);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader);
                    const long long post_t0 = p.prof ? clock64() : 0;
                    if (tma_out) {  // hand the finished tile to the store thread (warp 6)
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's tile rows -> visible to the bulk copy engine
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&outFull[sbuf]);
                    }
                    if (p.prof) prof_w13 += (unsigned long long)(clock64() - post_t0);
                } else {
                    // The ResBlock residual of this pixel is fetched into registers BEFORE waiting for the accumulator (up to 128
                    // channels at a time), so its latency hides behind the tile's MMA main loop.  tcgen05.ld is .sync.aligned:
                    // every lane of the warp executes it, only the memory accesses depend on `ok`.
                    // Both residuals (ResBlock input, U-skip) are summed into fp32 registers here: the last conv of an up level used to fetch
                    // its second residual inside the drain loop, one exposed global-memory latency per 16 channels (profiles/r02u_pair_roles.md).
                    constexpr int PRE = NOUT < 128 ? NOUT : 128;
                    const bool has_r1 = ok && (p.res1_hi != nullptr), has_r2 = ok && (p.res2_hi != nullptr);
                    float rs[PRE];
    #pragma unroll
                    for (int h0 = 0; h0 < NOUT; h0 += PRE) {
                        if (has_r1 || has_r2) {
    #pragma unroll
                            for (int i = 0; i < PRE; ++i) rs[i] = 0.f;
                        }
                        if (has_r1) {
                            const uint4* gh = reinterpret_cast<const uint4*>(p.res1_hi + o + h0);
                            const uint4* gl = reinterpret_cast<const uint4*>(p.res1_lo + o + h0);
    #pragma unroll
                            for (int i = 0; i < PRE / 8; ++i) add_split8(rs + 8 * i, gh[i], gl[i]);
                        }
                        if (has_r2) {
                            const uint4* gh = reinterpret_cast<const uint4*>(p.res2_hi + o + h0);
                            const uint4* gl = reinterpret_cast<const uint4*>(p.res2_lo + o + h0);
    #pragma unroll
                            for (int i = 0; i < PRE / 8; ++i) add_split8(rs + 8 * i, gh[i], gl[i]);
                        }
                        if (h0 == 0) {
                            TC_PROF_T0();
                            mbar_wait(&tfull[ab], aphase);
                            TC_PROF_ADD(prof_w3);
                            tc_fence_after();
                        }
                        if (p.dual & 32) continue;  // timing experiment: the epilogue only hands the accumulator back
    #pragma unroll
                        for (int c1 = 0; c1 < PRE; c1 += 16) {
                            const int cc = h0 + c1;
                            float v[16];
                            pair_acc16<NA, STACK>(taddr, taddr2, cc, v, resw);
                            if (has_r1 || has_r2) {
    #pragma unroll
                                for (int j = 0; j < 16; ++j) v[j] += rs[c1 + j];
                            }
                            if (ok) store_split16(p.out_hi + o + cc, p.out_lo + o + cc, v, p.relu);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_leader);
                }
            } else {
                mbar_wait(&tfull[ab], aphase);
                tc_fence_after();
                const int t = pt * NT + nt;
                float* prow0 = p.partial + ((size_t)t * p.nsplit * TC_BM + row) * NOUT;
                if (live) {
                    float* prow = prow0 + (size_t)split * TC_BM * NOUT;
#pragma unroll 1
                    for (int cc = 0; cc < NOUT; cc += 16) {
                        float v[16];
                        pair_acc16<NA, STACK>(taddr, taddr2, cc, v, resw);
                        float4* d = reinterpret_cast<float4*>(prow + cc);
#pragma unroll
                        for (int q = 0; q < 4; ++q) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader);
                if (live) {
                    __threadfence();
                    __syncwarp();
                    int old = 0;
                    if (lane == 0) old = atomicAdd(&p.tickets[t * 4 + lg], 1);
                    old = __shfl_sync(0xffffffffu, old, 0);
                    if (old == p.nsplit - 1) {
                        __threadfence();
                        if (lane == 0) p.tickets[t * 4 + lg] = 0;
                        if (ok) tc_epilogue_reduce<NOUT>(p, prow0, o);
                    }
                }
            }
        }
        if (threadIdx.x == 64) {
            TC_PROF_PUT(4, clock64() - role_t0);
            TC_PROF_PUT(3, prof_w3);
            TC_PROF_PUT(10, prof_w10);
            TC_PROF_PUT(11, prof_w11);
            TC_PROF_PUT(12, prof_w12);
            TC_PROF_PUT(13, prof_w13);
        }
    }
    if (lane == 0) TC_TRACE(6, 0x1000 + warp);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TC_TRACE(7, 1);
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal its barriers or run MMAs on it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
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

// 4-D activation map over pixels (x, y) with pixel strides (sx, sy) elements of C channels
int make_act_map(void* out_map, const void* base, int S, int H, int W, int C, size_t row_pitch_px, size_t slice_pitch_px, int px_stride,
                 int BW, int BH) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S};
    cuuint64_t gstr[3] = {(cuuint64_t)px_stride * C * 2, (cuuint64_t)row_pitch_px * C * 2, (cuuint64_t)slice_pitch_px * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", S, H, W, C, (int)r);
    return QMRI_OK;
}

}  // namespace

int tc_make_act_map(void* out_map, const void* base, int S, int H, int W, int C, int BW, int BH) {
    return make_act_map(out_map, base, S, H, W, C, (size_t)W, (size_t)H * W, 1, BW, BH);
}

// tap (dy, dx) of a 2x2 stride-2 conv: the [S][H][W][C] input seen as an (H/2) x (W/2) image of the pixels (2y + dy, 2x + dx)
int tc_make_down_map(void* out_map, const void* base, int S, int H, int W, int C, int dy, int dx, int BW, int BH) {
    const uint16_t* b = reinterpret_cast<const uint16_t*>(base) + ((size_t)dy * W + dx) * C;
    return make_act_map(out_map, b, S, H / 2, W / 2, C, (size_t)2 * W, (size_t)H * W, 2, BW, BH);
}

int tc_make_weight_map(void* out_map, const void* base, int K, int N, int BN) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled(weights %d x %d) failed: %d", N, K, (int)r);
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

size_t tc_partial_elems(int sm_count) { return (size_t)sm_count * TC_BM * 256; }
size_t tc_ticket_count(int sm_count) { return (size_t)sm_count * 4; }

static bool tc_pdl_enabled() { return getenv("QMRI_NO_PDL") == nullptr; }  // read per call: tests toggle it

static bool tc_cluster_splitk_enabled() { return getenv("QMRI_NO_CLUSTER_SPLITK") == nullptr; }  // read per call: tests toggle it

template <int BN, int MODE>
static int tc_configure(qmri_ctx* ctx) {
    static bool configured_dev[QMRI_MAX_DEV] = {};
    bool& configured = configured_dev[qmri_dev_slot(ctx)];
    if (!configured) {
        QCUDA(cudaFuncSetAttribute(tc_conv_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<BN>::SMEM));
        configured = true;
    }
    return QMRI_OK;
}

// clusters of `cs` CTAs of this kernel that can be resident at once (one CTA per SM, a cluster inside one GPC); cached
template <int BN, int MODE>
static int tc_max_active_clusters(qmri_ctx* ctx, int cs) {
    static int cache_dev[QMRI_MAX_DEV][9] = {};
    int* cache = cache_dev[qmri_dev_slot(ctx)];
    if (cs < 2 || cs > 8) return 0;
    if (cache[cs]) return cache[cs] < 0 ? 0 : cache[cs];
    if (tc_configure<BN, MODE>(ctx) != QMRI_OK) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * ctx->sm_count);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TcCfg<BN>::SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tc_conv_kernel<BN, MODE>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    cache[cs] = n > 0 ? n : -1;
    return n;
}

template <int BN, int MODE>
static int launch_tc(qmri_ctx* ctx, const TcMaps& maps, const TcK& k, int grid) {
    QCHECK((tc_configure<BN, MODE>(ctx)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TcCfg<BN>::SMEM;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tc_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (k.cluster_splitk) {
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = k.nsplit;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    QCUDA(cudaLaunchKernelEx(&cfg, tc_conv_kernel<BN, MODE>, maps, k));
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

// split-K through a cluster: the largest split count whose clusters all fit the machine in one wave (0 = use the global workspace)
template <int BN, int MODE>
static int tc_cluster_split(qmri_ctx* ctx, int total_tiles, int KB) {
    if (!tc_cluster_splitk_enabled()) return 0;
    for (int ns = std::min(8, KB / 4); ns >= 2; --ns) {
        if (total_tiles * ns > ctx->sm_count) continue;
        if (total_tiles > tc_max_active_clusters<BN, MODE>(ctx, ns)) continue;
        return ns;
    }
    return 0;
}

int conv_tc(qmri_ctx* ctx, const TcConvParams& p) {
    if (p.Cin % TC_BK) return qmri_fail(QMRI_EINVAL, "conv_tc: Cin %% 64");
    if (p.mode < 0 || p.mode > 2) return qmri_fail(QMRI_EINVAL, "conv_tc: mode %d", p.mode);
    const int BN = tc_block_n(p.Cout);
    const int ncols = (p.mode == TC_UP2X2) ? 4 * p.Cout : p.Cout;
    if (ncols % BN) return qmri_fail(QMRI_EINVAL, "conv_tc: N %% %d", BN);
    const int taps = (p.mode == TC_CONV3X3) ? 9 : (p.mode == TC_DOWN2X2) ? 4 : 1;
    const int KB = taps * (p.Cin / TC_BK);
    const int total_tiles = p.tiles_x * p.tiles_y * (ncols / BN) * p.S;
    // split K when the layer cannot fill the SMs (>= 4 k-blocks per split, workspace sized for sm_count work items)
    int nsplit = 1;
    if (p.partial && p.tickets && 2 * total_tiles <= ctx->sm_count) {
        nsplit = ctx->sm_count / total_tiles;
        if (nsplit > KB / 4) nsplit = KB / 4;
        if (nsplit > 8) nsplit = 8;
        if (nsplit < 1) nsplit = 1;
    }
    int cluster_splitk = 0;
    if (nsplit > 1) {  // prefer the cluster variant: partial tiles meet in distributed shared memory instead of L2 + tickets
        int cs = 0;
        if (BN == 64) cs = p.mode == TC_CONV3X3 ? tc_cluster_split<64, 0>(ctx, total_tiles, KB) : p.mode == TC_DOWN2X2 ? tc_cluster_split<64, 1>(ctx, total_tiles, KB) : tc_cluster_split<64, 2>(ctx, total_tiles, KB);
        else cs = p.mode == TC_CONV3X3 ? tc_cluster_split<128, 0>(ctx, total_tiles, KB) : p.mode == TC_DOWN2X2 ? tc_cluster_split<128, 1>(ctx, total_tiles, KB) : tc_cluster_split<128, 2>(ctx, total_tiles, KB);
        if (cs >= 2) {
            nsplit = cs;
            cluster_splitk = 1;
        }
    }
    const int work = total_tiles * nsplit;
    const int grid = work < ctx->sm_count ? work : ctx->sm_count;
    TcMaps maps;
    const int na = (p.mode == TC_DOWN2X2) ? 4 : 1;
    for (int i = 0; i < 4; ++i) {
        const int j = i < na ? i : 0;
        if (!p.mapA_hi[j] || !p.mapA_lo[j]) return qmri_fail(QMRI_EINVAL, "conv_tc: missing activation tensor map");
        maps.a_hi[i] = *(const CUtensorMap*)p.mapA_hi[j];
        maps.a_lo[i] = *(const CUtensorMap*)p.mapA_lo[j];
    }
    maps.b_hi = *(const CUtensorMap*)p.mapB_hi;
    maps.b_lo = *(const CUtensorMap*)p.mapB_lo;
    TcK k;
    k.out_hi = p.out_hi; k.out_lo = p.out_lo;
    k.res1_hi = p.res1_hi; k.res1_lo = p.res1_lo;
    k.res2_hi = p.res2_hi; k.res2_lo = p.res2_lo;
    k.partial = p.partial; k.tickets = p.tickets;
    k.S = p.S; k.H = p.H; k.W = p.W; k.Cin = p.Cin; k.Cout = p.Cout;
    k.BW = p.BW; k.BH = p.BH; k.tiles_x = p.tiles_x; k.tiles_y = p.tiles_y;
    k.relu = p.relu; k.nsplit = nsplit; k.trace = nullptr;
    k.cluster_splitk = cluster_splitk;
    k.dual = 0;
    k.resw = 0;
    k.tma_out = 0;
    k.res_tma = 0;
    k.l2pf = 0;
    k.prof = nullptr;
    if (BN == 64) {
        if (p.mode == TC_CONV3X3) return launch_tc<64, 0>(ctx, maps, k, grid);
        if (p.mode == TC_DOWN2X2) return launch_tc<64, 1>(ctx, maps, k, grid);
        return launch_tc<64, 2>(ctx, maps, k, grid);
    }
    if (p.mode == TC_CONV3X3) return launch_tc<128, 0>(ctx, maps, k, grid);
    if (p.mode == TC_DOWN2X2) return launch_tc<128, 1>(ctx, maps, k, grid);
    return launch_tc<128, 2>(ctx, maps, k, grid);
}

// pair kernel: BW must be a multiple of 8 pixels (dy shifts stay 1024-byte aligned)
int tc_slab_tile_shape(int W, int H, int* BW, int* BH) {
    const int cand[2][2] = {{16, 8}, {8, 16}};
    int best = 0;
    double best_eff = -1;
    for (int i = 0; i < 2; ++i) {
        int bw = cand[i][0], bh = cand[i][1];
        double eff = ((double)W * H) / ((double)((W + bw - 1) / bw * bw) * ((H + bh - 1) / bh * bh));
        if (i == 1) eff *= 1.04;  // the 8 x 16 slab is 10 % smaller; prefer it on ties
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = i;
        }
    }
    *BW = cand[best][0];
    *BH = cand[best][1];
    return QMRI_OK;
}

// weight-map box rows the pair kernel expects for a 3x3 conv with Cout output channels:
// rows[0] for mapB_hi / mapB_lo, rows[1] for the third map (second half-width copy of w_hi; stacked mode only)
void tc_pair_weight_boxes(int Cout, int* rows_main, int* rows_h2) {
    if (Cout >= 256) {
        *rows_main = 128;
        *rows_h2 = 128;  // unused
    } else {
        *rows_main = Cout;
        *rows_h2 = Cout / 2;
    }
}

template <int NA, int STACK>
static int launch_pair(qmri_ctx* ctx, const TcConvParams& p, TcK k) {
    using Cfg = PairCfg<NA, STACK>;
    static bool configured_dev[QMRI_MAX_DEV] = {};
    bool& configured = configured_dev[qmri_dev_slot(ctx)];
    if (!configured) {
        QCUDA(cudaFuncSetAttribute(tc_conv3x3_pair_kernel<NA, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    const int ptiles = p.tiles_x * p.tiles_y * p.S;
    const int NT = p.Cout / Cfg::NOUT;
    const int groups = ((ptiles + 1) / 2) * NT;
    const int max_clusters = ctx->sm_count / 2;
    const int U = 3 * (p.Cin / TC_BK);
    int nsplit = 1;
    if (p.partial && p.tickets && 2 * groups <= max_clusters) {
        nsplit = max_clusters / groups;
        if (nsplit > U / 2) nsplit = U / 2;
        if (nsplit > 8) nsplit = 8;
        if (nsplit < 1) nsplit = 1;
    }
    k.nsplit = nsplit;
    const int work = groups * nsplit;
    const int nclusters = work < max_clusters ? work : max_clusters;
    cudaLaunchConfig_t cfg = {};
    // A/B switches (tests, profiling): QMRI_TC_DUAL=1 turns the second MMA issuer on (measured 3 % slower: off by default),
    // QMRI_TC_RESW=0 turns the resident weights of the 64 -> 64 layers off
    static const int dual_on = getenv("QMRI_TC_DUAL") ? atoi(getenv("QMRI_TC_DUAL")) : 0;   // 1: second issuer; 4 / 8 / 12: timing experiments
    static const bool resw_off = getenv("QMRI_TC_RESW") && !strcmp(getenv("QMRI_TC_RESW"), "0");
    static const int resw_stages = getenv("QMRI_TC_RESW") ? atoi(getenv("QMRI_TC_RESW")) : 1;   // 2 / 3: cap on the slab stages
    k.dual = Cfg::DUAL_OK ? (dual_on & ~1) : 0;  // (the second-issuer mode, bit 0, is retired: its epilogue is gone)
    // QMRI_TC_DBG_LAYER=n: the experiment bits (and the profile line) apply to every 16th launch of this kernel only, starting with
    // the n-th - one layer of a forward runs the experiment on real data while the others run normally
    static const int dbg_layer = getenv("QMRI_TC_DBG_LAYER") ? atoi(getenv("QMRI_TC_DBG_LAYER")) : -1;
    static int launch_no = 0;
    bool dbg_this = true;
    if (Cfg::DUAL_OK && dbg_layer >= 0) {
        dbg_this = (launch_no++ % 16) == dbg_layer;
        if (!dbg_this) k.dual = 0;
    }
    k.resw = (Cfg::DUAL_OK && !resw_off && p.Cin == TC_BK) ? (resw_stages >= 2 ? resw_stages : 1) : 0;
    const char* tmaout_s = getenv("QMRI_TC_TMAOUT");   // A/B switches, read per call: tests toggle them
    const bool tma_out_off = tmaout_s && !strcmp(tmaout_s, "0");
    // tma_out: 2 = two staging buffers + two slab stages, 1 = one staging buffer + three slab stages (QMRI_TC_OUTBUF=1 / 2 forces one)
    const char* outbuf_s = getenv("QMRI_TC_OUTBUF");
    const int outbuf_env = outbuf_s ? atoi(outbuf_s) : 0;
    k.tma_out = (k.resw && p.mapO_hi && p.mapO_lo && !tma_out_off) ? (outbuf_env == 1 || outbuf_env == 2 ? outbuf_env : (p.res1_hi ? 2 : 1)) : 0;
    k.res_tma = (k.tma_out && p.res1_hi && p.mapR_hi && p.mapR_lo) ? 1 : 0;
    static const int l2pf_env = getenv("QMRI_TC_L2PF") ? atoi(getenv("QMRI_TC_L2PF")) : 0;   // A/B switch: 1 = on for all pair layers
    k.l2pf = l2pf_env > 0 ? 1 : 0;   // measured: no gain (the slab stream is throughput-, not latency-bound), off by default
    cfg.gridDim = dim3(nclusters * 2);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = tc_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    static int* trace_host = nullptr;
    static int* trace_dev = nullptr;
    static const bool trace_on = getenv("QMRI_TC_TRACE") != nullptr;
    if (trace_on && !trace_host) {
        QCUDA(cudaHostAlloc((void**)&trace_host, 256 * 8 * sizeof(int), cudaHostAllocMapped));
        QCUDA(cudaHostGetDevicePointer((void**)&trace_dev, trace_host, 0));
    }
    if (trace_on) {
        memset(trace_host, 0, 256 * 8 * sizeof(int));
        k.trace = trace_dev;
    }
    // QMRI_TC_PROF=1: role-level attribution - cycles each role of each CTA spends waiting (serialises the launches)
    static unsigned long long* prof_dev = nullptr;
    static const bool prof_on = getenv("QMRI_TC_PROF") != nullptr;
    k.prof = nullptr;
    if (prof_on) {
        if (!prof_dev) QCUDA(cudaMalloc((void**)&prof_dev, 256 * 16 * sizeof(unsigned long long)));
        QCUDA(cudaMemsetAsync(prof_dev, 0, 256 * 16 * sizeof(unsigned long long), ctx->stream));
        k.prof = prof_dev;
    }
    QCUDA(cudaLaunchKernelEx(&cfg, tc_conv3x3_pair_kernel<NA, STACK>, *(const CUtensorMap*)p.mapA_hi[0], *(const CUtensorMap*)p.mapA_lo[0],
                             *(const CUtensorMap*)p.mapB_hi, *(const CUtensorMap*)p.mapB_lo, *(const CUtensorMap*)p.mapB_h2,
                             *(const CUtensorMap*)(p.mapB_l2 ? p.mapB_l2 : p.mapB_h2),
                             *(const CUtensorMap*)(p.mapO_hi ? p.mapO_hi : p.mapA_hi[0]), *(const CUtensorMap*)(p.mapO_lo ? p.mapO_lo : p.mapA_lo[0]),
                             *(const CUtensorMap*)(p.mapR_hi ? p.mapR_hi : p.mapA_hi[0]), *(const CUtensorMap*)(p.mapR_lo ? p.mapR_lo : p.mapA_lo[0]),
                             (const TcK)k));
    QLAUNCH_CHECK(ctx);
    if (prof_on && dbg_this) {
        std::vector<unsigned long long> h(256 * 16);
        QCUDA(cudaStreamSynchronize(ctx->stream));
        QCUDA(cudaMemcpy(h.data(), prof_dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double a[16] = {};
        int nl = 0;
        for (int c = 0; c < nclusters * 2; c += 2, ++nl)  // leader CTAs
            for (int j = 0; j < 16; ++j) a[j] += (double)h[c * 16 + j];
        for (int j = 0; j < 16; ++j) a[j] /= nl;
        fprintf(stderr, "[qmri prof] pair<%d,%d> S=%d %dx%d Cin=%d Cout=%d resw=%d: per leader CTA: tiles %.1f, MMA role %.0f cycles in %.0f ns (%.0f MHz) = %.0f per tile "
                        "(waiting for slabs %.0f, for a free accumulator %.0f; inside the MMA issue loops %.0f, inside commits %.0f), producer waiting for a free stage %.0f, epilogue role %.0f "
                        "(waiting for the accumulator %.0f, for the staging buffer %.0f, inside tcgen05.wait::ld %.0f, drain loop + accumulator hand-back %.0f, proxy fence + tile hand-over %.0f)\n",
                NA, STACK, p.S, p.H, p.W, p.Cin, p.Cout, k.resw, a[7], a[2], a[6], a[6] > 0 ? 1e3 * a[2] / a[6] : 0.0, a[2] / (a[7] > 0 ? a[7] : 1), a[0], a[1], a[8], a[9], a[5], a[4], a[3], a[11], a[10], a[12], a[13]);
    }
    if (trace_on) {  // watchdog: a hang dumps where every role of every CTA stopped, then the process exits
        for (int ms = 0; ms < 5000; ++ms) {
            if (cudaStreamQuery(ctx->stream) != cudaErrorNotReady) return QMRI_OK;
            struct timespec ts = {0, 1000000};
            nanosleep(&ts, nullptr);
        }
        fprintf(stderr, "[qmri trace] pair kernel NA=%d STACK=%d hung: S=%d H=%d W=%d Cin=%d Cout=%d tile %dx%d nsplit=%d clusters=%d work=%d\n", NA, STACK,
                p.S, p.H, p.W, p.Cin, p.Cout, p.BW, p.BH, nsplit, nclusters, work);
        for (int c = 0; c < nclusters * 2; ++c) {
            const int* t = trace_host + c * 8;
            fprintf(stderr, "[qmri trace] cta %3d prod %08x mma %08x epi %08x %08x %08x %08x end %08x sync %d\n", c, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
        }
        _exit(3);
    }
    return QMRI_OK;
}

// ------------------------------------------------------------------------------------------------
// Operand-swapped 3x3 conv for the 64 -> 64 layers (Cin = Cout = 64).
//
// (Experiment, off by default, NOT brought forward to the warp-uniform issue loop / staging-tile epilogue of the pair kernel.)  Built when
// MMA issue looked like a ~103-cycle hardware floor - it was ptxas' register-to-uniform loop around UTCHMMAs issued from divergent code,
// see tc2_mma_elect - under which a layer with only N = 64 output channels could not feed the tensor pipe from the N side.  Here the
// roles are swapped: the stacked weights [w_hi ; w_lo] are the M = 128 A operand and live in TENSOR MEMORY for the whole kernel
// (576 bf16 per row = 288 columns, written once per CTA with tcgen05.st), the PIXELS are the N dimension - 224 of them per tile (14
// rows x 16), the widest N the remaining 224 TMEM columns allow for the accumulator - and only the activation slabs stream through
// shared memory.  Per K step two N = 224 MMAs (B = the hi and the lo activation plane; row 2 c of D collects w_hi[c] (a_hi + a_lo),
// row 2 c + 1 w_lo[c] (a_hi + a_lo) - the fourth product a_lo w_lo comes for free).  The accumulator is single-buffered; the epilogue
// therefore drains it quickly (tcgen05.ld, the hi / lo rows of a channel meet through one lane shuffle, the tile goes to shared
// memory as [channel][pixel] fp32) and hands it back, and the slow part - residual, ReLU, hi-lo split, 16-byte stores - runs from
// shared memory while the next tile's MMAs are under way.
// ------------------------------------------------------------------------------------------------
constexpr int SW_BW = 16, SW_BH = 14, SW_N = SW_BW * SW_BH;              // 224 pixels per tile
constexpr int SW_STAGES = 2;
constexpr uint32_t SW_PLANE = (SW_BH + 2) * SW_BW * 128;                  // 32 KB: (14 + 2) x 16 rows of 128 B
constexpr uint32_t SW_STAGE = 2 * SW_PLANE;                               // hi + lo
constexpr size_t SW_SMEM = (size_t)SW_STAGES * SW_STAGE + 64 * SW_N * 4 /*the finished tile, [channel][pixel] fp32*/ + 1024 + 256;
constexpr uint32_t SW_ACOL = 224;                                         // TMEM: D in columns [0, 224), weights in [224, 512)

struct SwapK {
    uint16_t *out_hi, *out_lo;
    const uint16_t *res1_hi, *res1_lo, *res2_hi, *res2_lo;
    const uint16_t *w_hi, *w_lo;   // [64][576] K-major (tap-major, then input channel)
    int S, H, W, tiles_x, tiles_y, relu;
};

__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
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
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the four epilogue warps

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv64_swap_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, const SwapK p) {
    extern __shared__ unsigned char tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    unsigned char* smem = tc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    float* tile = reinterpret_cast<float*>(smem + (size_t)SW_STAGES * SW_STAGE);   // [64 channels][224 pixels]; pixel j of a 32-pixel chunk of
                                                                                    // channel c sits at (j + c) & 31 (bank-conflict-free both ways)
    uint64_t* bars = reinterpret_cast<uint64_t*>(tile + 64 * SW_N);
    uint64_t* full = bars;                  // [SW_STAGES]
    uint64_t* empty = bars + SW_STAGES;     // [SW_STAGES]
    uint64_t* dfull = bars + 2 * SW_STAGES; // accumulator complete
    uint64_t* dempty = dfull + 1;           // accumulator drained (4 epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int ntiles = tiles_xy * p.S;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SW_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(dfull, 1);
        mbar_init(dempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmA_hi);
        tma_prefetch_desc(&tmA_lo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int lg = warp & 3;  // TMEM lane quarter of an epilogue warp (warps 2-5 -> quarters 2, 3, 0, 1)
    if (warp >= 2) {
        // the layer's weights into TMEM: lane 2 c = w_hi[c], lane 2 c + 1 = w_lo[c]; 288 columns of bf16 pairs
        const int r = lg * 32 + lane;
        const uint4* src = reinterpret_cast<const uint4*>(((r & 1) ? p.w_lo : p.w_hi) + (size_t)(r >> 1) * 576);
#pragma unroll 1
        for (int c = 0; c < 9; ++c) {
            uint32_t v[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 t = __ldg(src + c * 8 + q);
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
            tc_st32(tmem_base + ((uint32_t)(lg * 32) << 16) + SW_ACOL + 32 * c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();  // everything below reads activations produced by earlier kernels

    if (warp == 0) {
        // ================= TMA producer: one (16 + 2... rows) x 16 slab pair per dx =================
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int w = blockIdx.x; w < ntiles; w += gridDim.x) {
                const int s = w / tiles_xy, txy = w - s * tiles_xy;
                const int x0 = (txy % p.tiles_x) * SW_BW, y0 = (txy / p.tiles_x) * SW_BH;
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(&empty[st], ph ^ 1);
                    unsigned char* sp = smem + (size_t)st * SW_STAGE;
                    mbar_expect_tx(&full[st], SW_STAGE);
                    tma_load_4d(sp, &tmA_hi, &full[st], 0, x0 + dxi - 1, y0 - 1, s);
                    tma_load_4d(sp + SW_PLANE, &tmA_lo, &full[st], 0, x0 + dxi - 1, y0 - 1, s);
                    if (++st == SW_STAGES) {
                        st = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // D = f32, A = B = bf16, K-major, M = 128, N = 224
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SW_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t tmem_d = tmem_base, tmem_a = tmem_base + SW_ACOL;
            int st = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int w = blockIdx.x; w < ntiles; w += gridDim.x, ++it) {
                mbar_wait(dempty, (uint32_t)(it & 1) ^ 1);  // the epilogue has read the previous tile out of the accumulator
                tc_fence_after();
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + (size_t)st * SW_STAGE);
#pragma unroll
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        const uint64_t b_hi = umma_desc(sb + dyi * (SW_BW * 128)), b_lo = umma_desc(sb + SW_PLANE + dyi * (SW_BW * 128));
                        const uint32_t a_tap = tmem_a + (uint32_t)((dyi * 3 + dxi) * 32);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {
                            tc_mma_ts(tmem_d, a_tap + 8 * k, b_hi + (uint64_t)(2 * k), idesc, (dxi | dyi | k) ? 1u : 0u);
                            tc_mma_ts(tmem_d, a_tap + 8 * k, b_lo + (uint64_t)(2 * k), idesc, 1u);
                        }
                    }
                    tc_commit(&empty[st]);
                    if (++st == SW_STAGES) {
                        st = 0;
                        ph ^= 1;
                    }
                }
                tc_commit(dfull);
            }
        }
    } else {
        // ================= epilogue warps =================
        const int c = 16 * lg + (lane >> 1);   // output channel of this lane pair (even lane: w_hi part, odd lane: w_lo part)
        const int pr = lane & 1;               // after the shuffle both lanes hold the sum; lane parity pr stores columns 16 pr .. 16 pr + 15
        const int seg = warp - 2;              // store phase: warp = 16-channel segment, lane = pixel (mod 32)
        int it = 0;
        for (int w = blockIdx.x; w < ntiles; w += gridDim.x, ++it) {
            const int s = w / tiles_xy, txy = w - s * tiles_xy;
            const int x0 = (txy % p.tiles_x) * SW_BW, y0 = (txy / p.tiles_x) * SW_BH;
            epi_bar();  // the previous tile's store phase has finished reading `tile`
            mbar_wait(dfull, (uint32_t)(it & 1));
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < SW_N; c0 += 32) {
                uint32_t a[32];
                tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + c0, a);
                tc_wait_ld();
                float* row = tile + c * SW_N + c0;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const float lo16 = __uint_as_float(a[t]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(a[t]), 1);
                    const float hi16 = __uint_as_float(a[16 + t]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(a[16 + t]), 1);
                    row[(16 * pr + t + c) & 31] = pr ? hi16 : lo16;
                }
            }
            // the accumulator has been read out: the next tile's MMAs may start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty);
            epi_bar();  // `tile` is complete
#pragma unroll 1
            for (int q = 0; q < SW_N / 32; ++q) {
                const int n = 32 * q + lane;
                const int x = x0 + (n & 15), y = y0 + (n >> 4);
                if (x < p.W && y < p.H) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int co = 16 * seg + i;
                        v[i] = tile[co * SW_N + 32 * q + ((lane + co) & 31)];
                    }
                    const size_t o = (((size_t)s * p.H + y) * p.W + x) * 64 + 16 * seg;
                    if (p.res1_hi) {
                        const uint4* gh = reinterpret_cast<const uint4*>(p.res1_hi + o);
                        const uint4* gl = reinterpret_cast<const uint4*>(p.res1_lo + o);
                        add_split8(v, gh[0], gl[0]);
                        add_split8(v + 8, gh[1], gl[1]);
                    }
                    if (p.res2_hi) {
                        const uint4* gh = reinterpret_cast<const uint4*>(p.res2_hi + o);
                        const uint4* gl = reinterpret_cast<const uint4*>(p.res2_lo + o);
                        add_split8(v, gh[0], gl[0]);
                        add_split8(v + 8, gh[1], gl[1]);
                    }
                    store_split16(p.out_hi + o, p.out_lo + o, v, p.relu);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// Host side: mapA_*[0] = activation maps with box (16 px, 16 rows) of 64 channels; w_hi / w_lo = the layer's raw K-major weight planes.
bool conv64_swap_supported(int W, int H, int Cin, int Cout) { return Cin == 64 && Cout == 64 && W % SW_BW == 0 && H % SW_BH == 0; }
int conv64_swap(qmri_ctx* ctx, const TcConvParams& p, const uint16_t* w_hi, const uint16_t* w_lo) {
    if (!conv64_swap_supported(p.W, p.H, p.Cin, p.Cout)) return qmri_fail(QMRI_EINVAL, "conv64_swap: %d x %d, %d -> %d channels", p.W, p.H, p.Cin, p.Cout);
    if (!p.mapA_hi[0] || !p.mapA_lo[0] || !w_hi || !w_lo) return qmri_fail(QMRI_EINVAL, "conv64_swap: missing tensor map / weights");
    static bool configured_dev[QMRI_MAX_DEV] = {};
    bool& configured = configured_dev[qmri_dev_slot(ctx)];
    if (!configured) {
        QCUDA(cudaFuncSetAttribute(tc_conv64_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM));
        configured = true;
    }
    SwapK k;
    k.out_hi = p.out_hi; k.out_lo = p.out_lo;
    k.res1_hi = p.res1_hi; k.res1_lo = p.res1_lo;
    k.res2_hi = p.res2_hi; k.res2_lo = p.res2_lo;
    k.w_hi = w_hi; k.w_lo = w_lo;
    k.S = p.S; k.H = p.H; k.W = p.W;
    k.tiles_x = p.W / SW_BW; k.tiles_y = p.H / SW_BH;
    k.relu = p.relu;
    const int ntiles = k.tiles_x * k.tiles_y * p.S;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ntiles < ctx->sm_count ? ntiles : ctx->sm_count);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = SW_SMEM;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tc_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QCUDA(cudaLaunchKernelEx(&cfg, tc_conv64_swap_kernel, *(const CUtensorMap*)p.mapA_hi[0], *(const CUtensorMap*)p.mapA_lo[0], (const SwapK)k));
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

// 3x3 conv through the CTA-pair kernel.  mapA_*[0]: activation maps with box (BW, BH + 2); mapB_hi / mapB_lo / mapB_h2:
// weight maps with the box rows of tc_pair_weight_boxes().
int conv3x3_tc_pair(qmri_ctx* ctx, const TcConvParams& p) {
    if (p.Cin % TC_BK) return qmri_fail(QMRI_EINVAL, "conv3x3_tc_pair: Cin %% 64");
    if (p.Cout != 64 && p.Cout != 128 && p.Cout % 256) return qmri_fail(QMRI_EINVAL, "conv3x3_tc_pair: Cout %d", p.Cout);
    if (p.BW % 8 || p.BW * p.BH != TC_BM || (p.BH + 2) * p.BW * 128 > (int)PairCfg<256, 0>::A_PLANE)
        return qmri_fail(QMRI_EINVAL, "conv3x3_tc_pair: tile %d x %d not supported", p.BW, p.BH);
    if (!p.mapA_hi[0] || !p.mapA_lo[0] || !p.mapB_hi || !p.mapB_lo || !p.mapB_h2) return qmri_fail(QMRI_EINVAL, "conv3x3_tc_pair: missing tensor map");
    TcK k;
    k.out_hi = p.out_hi; k.out_lo = p.out_lo;
    k.res1_hi = p.res1_hi; k.res1_lo = p.res1_lo;
    k.res2_hi = p.res2_hi; k.res2_lo = p.res2_lo;
    k.partial = p.partial; k.tickets = p.tickets;
    k.S = p.S; k.H = p.H; k.W = p.W; k.Cin = p.Cin; k.Cout = p.Cout;
    k.BW = p.BW; k.BH = p.BH; k.tiles_x = p.tiles_x; k.tiles_y = p.tiles_y;
    k.relu = p.relu; k.nsplit = 1; k.trace = nullptr; k.cluster_splitk = 0; k.dual = 0; k.resw = 0; k.tma_out = 0; k.res_tma = 0; k.l2pf = 0; k.prof = nullptr;
    if (p.Cout == 64) return launch_pair<128, 1>(ctx, p, k);
    if (p.Cout == 128) return launch_pair<256, 1>(ctx, p, k);
    return launch_pair<256, 0>(ctx, p, k);
}
