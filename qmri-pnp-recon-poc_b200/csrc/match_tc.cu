// K2, tensor-pipe variant - dictionary matching as a tcgen05 kind::tf32 contraction with a fused group-maximum epilogue and an
// FP32 rescore; the K x B score matrix of the reference is never materialised.
//
// Reference being replaced: main_files/dictionary_matching/mrf_dtm_cpu.m:84-98
//   ip = dict.D * ctranspose(x(cind,:));  [mt, dm] = max(abs(ip),[],1)
//
// Numerics (profiles/tools/k2_tf32_emulation.py): single-pass TF32 operands put ~800 atoms of a dense dictionary within the
// rounding error of the maximum, so every fp32 value a is split a = hi + lo, hi = tf32(a), lo = tf32(a - hi), and the C <= 10
// channel contraction is laid out along a K = 32 axis
//      pixel row  [ x_hi | x_lo | x_hi | 0 0 ]      atom row  [ d_hi | d_hi | d_lo | 0 0 ]
// = x_hi d_hi + x_lo d_hi + x_hi d_lo (the dropped x_lo d_lo term is 2^-22), fp32 accumulation in TMEM: scores as accurate as
// the fp32 FMA kernel's (1e-6 |x|^2).  4 MMAs of K = 8 per operand part; complex pixels = two accumulators (real, imaginary).
//
// Kernel: persistent, warp-specialised.  Work item = (tile of 128 pixels, atom range).  Warp 0 (TMA): the pixel tile (real and
// imaginary K-major rows, 2 x 16 KB, loaded once per work item) and a ring of 128-atom tiles (16 KB each) into 128B-swizzled
// shared memory.  Warp 1 and the last warp (MMA): per atom tile 4 tcgen05.mma (M128 x N128 x K8) each into one of two TMEM accumulator
// pairs - warp 1 the real part, the last warp the imaginary part (one thread issues an MMA every ~103 cycles at best).
// Warps 2-5 (epilogue; warps 2-9 for atom ranges >= 65 536: two per TMEM lane quarter, each taking 64 of the tile's 128 atom
// columns): thread = pixel; tcgen05.ld of
// the real and imaginary accumulators (issued one 32-column chunk ahead of the arithmetic), s = re^2 + im^2, the maximum of each
// 16-atom group as a depth-4 tree (3 FP32 instructions per score; with one or two warps per scheduler a 16-long dependent
// FMNMX chain was the limiter: ncu showed the epilogue warps issuing 41 % of the cycles).  The thread keeps the TWO BEST GROUPS
// by group maximum (strict comparisons: earlier groups stay ahead on exact ties).  A per-atom top-k list degenerates on real
// dictionaries, whose scores rise smoothly along the atom index so that nearly every group enters the list: measured 5.5e11
// px-atoms/s on the benchmark's dictionary against 1.8e12 on random atoms, and 3.1e12 for group tracking on either.
// At the end of the work item the thread's 32 candidate atoms are rescored with k2_score() - the FP32 FMA kernel's exact
// expression on the ORIGINAL fp32 atoms and pixels - and the best becomes the packed key
//     float_bits(score) << 32 | (0xFFFFFFFF - atom)
// merged with atomicMax like the FMA kernel's (same keys, so the two column halves of a pixel, atom ranges, kernels and ranks
// mix freely).  The approximate scores only have to bring the true maximum's group into the top two: they are within 5e-7
// (relative) of the FP32 scores, so an atom can be missed only inside a reference near-tie (top-2 gap < 1e-6).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "match_kernel.h"
#include "match_score.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int MT_THREADS_MAX = 96 + 32 * 8;    // TMA warp + MMA warp + 4 or 8 epilogue warps (EW) + second MMA warp
constexpr int MT_BM = 128;                  // pixels per tile
constexpr int MT_BN = 128;                  // atoms per tile
constexpr int MT_KF = 32;                   // floats per operand row (128 B = one swizzle row)
constexpr int MT_STAGES = 6;
constexpr uint32_t MT_A_PART = MT_BM * MT_KF * 4;   // 16 KB
constexpr uint32_t MT_B_TILE = MT_BN * MT_KF * 4;   // 16 KB
constexpr uint32_t MT_TMEM_COLS = 512;              // 2 buffers x (real, imaginary) x 128 columns
constexpr size_t MT_SMEM = 2 * MT_A_PART + (size_t)MT_STAGES * MT_B_TILE + 1024 /*alignment*/ + 256 /*barriers*/;

struct MtParams {
    const float* x_re;   // [C][npix] planar, the original fp32 pixels (rescoring)
    const float* x_im;   // may be null
    int64_t npix;
    int64_t npix_pad;    // rows per part of the staged A matrix
    const float* Dp;     // [K][CP] original fp32 atoms, indexed by GLOBAL atom number
    int64_t a0, a1;      // atoms scored: tile t of the packed matrix holds atoms a0 + 128 t ...
    int ntiles;          // atom tiles of the packed matrix
    int nsplit;          // atom-range splits per pixel tile
    int issuers;         // MMA-issuing warps for complex pixels: 2 (default) or 1
    int debug;           // profiling only (QMRI_K2_DEBUG): 1 = epilogue skips reads and arithmetic, 2 = MMA warp skips the MMAs
    int ptiles;
    int CP;
    unsigned long long* keys;
};

// x planar [C][npix] -> K-major operand rows [part][npix_pad][32] = [hi | lo | hi | 0 0] (part 0 = real, 1 = imaginary)
__global__ void match_prep_kernel(const float* __restrict__ x_re, const float* __restrict__ x_im, int64_t npix, int64_t npix_pad, int C,
                                  float* __restrict__ A) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int part = blockIdx.y;
    if (pix >= npix_pad) return;
    const float* x = part ? x_im : x_re;
    float4* row = reinterpret_cast<float4*>(A + ((int64_t)part * npix_pad + pix) * MT_KF);
    float v[MT_KF];
#pragma unroll
    for (int j = 0; j < MT_KF; ++j) v[j] = 0.f;
    if (pix < npix && x) {
        float m = 0.f;  // per-pixel power-of-two scale over both parts (match_score.cuh)
        for (int c = 0; c < C; ++c) {
            m = fmaxf(m, fabsf(__ldg(x_re + (int64_t)c * npix + pix)));
            if (x_im) m = fmaxf(m, fabsf(__ldg(x_im + (int64_t)c * npix + pix)));
        }
        const float sc = k2_pixel_scale(m);
        for (int c = 0; c < C; ++c) {
            const float a = sc * __ldg(x + (int64_t)c * npix + pix);
            const float hi = to_tf32(a);
            v[c] = hi;
            v[C + c] = to_tf32(a - hi);
            v[2 * C + c] = hi;
        }
    }
#pragma unroll
    for (int j = 0; j < MT_KF / 4; ++j) row[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

template <int C, bool CPLX, int EW>
__global__ void __launch_bounds__(96 + 32 * EW, 1) match_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                const MtParams p) {
    extern __shared__ unsigned char mt_smem_raw[];
    const uint32_t raw = smem_u32(mt_smem_raw);
    unsigned char* smem = mt_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* smA = smem;                       // [2 parts][128 rows][128 B]
    unsigned char* smB = smem + 2 * MT_A_PART;       // [STAGES][128 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)MT_STAGES * MT_B_TILE);
    uint64_t* b_full = bars;                         // [STAGES]
    uint64_t* b_empty = bars + MT_STAGES;            // [STAGES]
    uint64_t* a_full = bars + 2 * MT_STAGES;         // [1]
    uint64_t* a_empty = a_full + 1;                  // [1]
    uint64_t* tfull = a_full + 2;                    // [2]
    uint64_t* tempty = a_full + 4;                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 6);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // shuffle: warp-uniform for ptxas
    const int total_work = p.ptiles * p.nsplit;

    // Two MMA-issuing warps for complex pixels (built against a ~103-cycle "issue floor" that turned out to be ptxas' register-to-uniform
    // loop around every UTCHMMA issued from divergent code - tc_ptx.cuh; with the warp-uniform loop one issuer does as well, two are kept):
    // and the real and the imaginary accumulator are independent - warp 1 issues the real part's MMAs, the last warp the imaginary
    // part's.  Both commit on the stage / accumulator / pixel-tile barriers (count 2).  QMRI_K2_ISSUERS=1 keeps a single issuer.
    const bool two = CPLX && p.issuers == 2;
    if (threadIdx.x == 0) {
        const uint32_t ni = two ? 2u : 1u;
        for (int s = 0; s < MT_STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], ni);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, ni);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], ni);
            mbar_init(&tempty[a], EW);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(tmem_slot, MT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int pt = w / p.nsplit, split = w - pt * p.nsplit;
                const int t0 = (int)(((int64_t)split * p.ntiles) / p.nsplit), t1 = (int)(((int64_t)(split + 1) * p.ntiles) / p.nsplit);
                mbar_wait(a_empty, aphase ^ 1);  // the MMAs of the previous work item have read the pixel tile
                mbar_expect_tx(a_full, CPLX ? 2 * MT_A_PART : MT_A_PART);
                tma_load_2d(smA, &tmA, a_full, 0, pt * MT_BM);
                if (CPLX) tma_load_2d(smA + MT_A_PART, &tmA, a_full, 0, (int)(p.npix_pad + (int64_t)pt * MT_BM));
                aphase ^= 1;
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&b_empty[stage], phase ^ 1);
                    mbar_expect_tx(&b_full[stage], MT_B_TILE);
                    tma_load_2d(smB + (size_t)stage * MT_B_TILE, &tmB, &b_full[stage], 0, t * MT_BN);
                    if (++stage == MT_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 2 + EW) {
        // ================= MMA issuer(s): warp 1 (real part; both parts with a single issuer), warp 2 + EW (imaginary part) =================
        const bool second = warp != 1;
        if (!second || two) {  // the whole warp runs the issue loop, one elected lane issues (tc_ptx.cuh)
            const bool do_re = !second, do_im = CPLX && (second || !two);
            // instruction descriptor: D = f32, A = B = tf32, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(MT_BN >> 3) << 17) | ((uint32_t)(MT_BM >> 4) << 24);
            const uint64_t a_re = umma_desc_sw128(smem_u32(smA)), a_im = umma_desc_sw128(smem_u32(smA + MT_A_PART));
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            int it = 0;  // accumulator-buffer sequence number over all atom tiles of this CTA
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int pt = w / p.nsplit, split = w - pt * p.nsplit;
                const int t0 = (int)(((int64_t)split * p.ntiles) / p.nsplit), t1 = (int)(((int64_t)(split + 1) * p.ntiles) / p.nsplit);
                mbar_wait(a_full, aphase);
                aphase ^= 1;
                tc_fence_after();
                for (int t = t0; t < t1; ++t, ++it) {
                    // (the shuffles mark the loop-carried values as warp-uniform for ptxas: operands stay in uniform registers)
                    const int ab = __shfl_sync(0xffffffffu, it & 1, 0);
                    mbar_wait(&tempty[ab], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator pair
                    mbar_wait(&b_full[stage], phase);
                    tc_fence_after();
                    const uint64_t b = umma_desc_sw128(__shfl_sync(0xffffffffu, smem_u32(smB + (size_t)stage * MT_B_TILE), 0));
                    const uint32_t d_re = tmem_base + (uint32_t)(ab * 2 * MT_BN), d_im = d_re + MT_BN;
                    if (p.debug != 2) {
                        if (do_re) {
#pragma unroll
                            for (int k = 0; k < MT_KF / 8; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);  // 32 bytes along K inside the swizzle row (16-byte units)
                                mma_tf32_ss_elect(d_re, a_re + ko, b + ko, idesc, k ? 1u : 0u);
                            }
                        }
                        if (do_im) {
#pragma unroll
                            for (int k = 0; k < MT_KF / 8; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);
                                mma_tf32_ss_elect(d_im, a_im + ko, b + ko, idesc, k ? 1u : 0u);
                            }
                        }
                    }
                    tc_commit_elect(&b_empty[stage]);
                    tc_commit_elect(&tfull[ab]);
                    if (++stage == MT_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit_elect(a_empty);  // the pixel tile may be overwritten once every MMA above has completed
            }
        }
    } else {
        // ================= epilogue warps: thread = (pixel = TMEM lane, half of the tile's atom columns) =================
        const int lg = warp & 3;                 // TMEM lane quarter this warp may read
        constexpr int NCH = 4 / (EW / 4);        // 32-column chunks of a tile per warp: 4 (EW = 4) or 2 (EW = 8, two warps per lane quarter)
        const int half = EW == 8 ? (warp - 2) >> 2 : 0;  // EW = 8: columns [64 half, 64 half + 64) of every accumulator
        const int row = lg * 32 + lane;
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const int pt = w / p.nsplit, split = w - pt * p.nsplit;
            const int t0 = (int)(((int64_t)split * p.ntiles) / p.nsplit), t1 = (int)(((int64_t)(split + 1) * p.ntiles) / p.nsplit);
            // the two best 16-atom groups by their largest score (strict comparisons: the earlier group stays ahead on exact ties)
            float g1s = -1.f, g2s = -1.f;
            int g1k = -1, g2k = -1;
            auto score_chunk = [&](const uint32_t (&re)[32], const uint32_t (&im)[32], int gid) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    float sc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float a = __uint_as_float(re[16 * g + j]);
                        if (CPLX) {
                            const float b = __uint_as_float(im[16 * g + j]);
                            sc[j] = fmaf(b, b, a * a);
                        } else {
                            sc[j] = a * a;
                        }
                    }
                    // maximum as a tree (depth 4): a single warp per scheduler cannot hide a 16-long dependent chain
#pragma unroll
                    for (int st = 8; st > 0; st >>= 1)
#pragma unroll
                        for (int j = 0; j < st; ++j) sc[j] = fmaxf(sc[j], sc[j + st]);  // fmaxf drops NaN
                    const float gmax = fmaxf(sc[0], -1.f);
                    if (gmax > g2s) {
                        if (gmax > g1s) {
                            g2s = g1s;
                            g2k = g1k;
                            g1s = gmax;
                            g1k = gid + g;
                        } else {
                            g2s = gmax;
                            g2k = gid + g;
                        }
                    }
                }
            };
            auto load_chunk = [&](uint32_t taddr, uint32_t (&re)[32], uint32_t (&im)[32]) {
                tc_ld32(taddr, re);
                if (CPLX) tc_ld32(taddr + MT_BN, im);
            };
            // Accumulator reads run one 32-column chunk ahead of the arithmetic: a chunk's tcgen05.ld is issued before the previous
            // chunk is scored and waited for afterwards (tcgen05.wait::ld covers every outstanding load of the thread).
            uint32_t reA[32], imA[32], reB[32], imB[32];
            auto tile_addr = [&](int i) { return tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((i & 1) * 2 * MT_BN + 32 * NCH * half); };
            if (t0 < t1) {
                mbar_wait(&tfull[it & 1], (it >> 1) & 1);
                tc_fence_after();
                load_chunk(tile_addr(it), reA, imA);
                tc_wait_ld();
            }
            for (int t = t0; t < t1; ++t, ++it) {
                const int ab = it & 1;
                const int gid = (t - t0) * (MT_BN / 16) + 2 * NCH * half;
                if (p.debug == 1) {  // profiling: hand the buffer straight back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[ab]);
                    if (t + 1 < t1) {
                        mbar_wait(&tfull[(it + 1) & 1], ((it + 1) >> 1) & 1);
                        tc_fence_after();
                    }
                    continue;
                }
                if (NCH == 4) {
                    load_chunk(tile_addr(it) + 32, reB, imB);
                    score_chunk(reA, imA, gid);
                    tc_wait_ld();
                    load_chunk(tile_addr(it) + 64, reA, imA);
                    score_chunk(reB, imB, gid + 2);
                    tc_wait_ld();
                }
                load_chunk(tile_addr(it) + 32 * (NCH - 1), reB, imB);
                score_chunk(reA, imA, gid + 2 * (NCH - 2));
                tc_wait_ld();
                // this warp's share of the tile is in registers: hand the buffer back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[ab]);
                if (t + 1 < t1) {
                    mbar_wait(&tfull[(it + 1) & 1], ((it + 1) >> 1) & 1);
                    tc_fence_after();
                    load_chunk(tile_addr(it + 1), reA, imA);
                }
                score_chunk(reB, imB, gid + 2 * (NCH - 1));
                if (t + 1 < t1) tc_wait_ld();
            }
            // rescore the 2 x 16 candidates with the FMA kernel's exact expression on the original fp32 data
            const int64_t pix = (int64_t)pt * MT_BM + row;
            if (pix < p.npix && g1k >= 0) {
                float xr[C], xi[CPLX ? C : 1];
                float m = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    xr[c] = __ldg(p.x_re + (int64_t)c * p.npix + pix);
                    m = fmaxf(m, fabsf(xr[c]));
                    if (CPLX) {
                        xi[c] = __ldg(p.x_im + (int64_t)c * p.npix + pix);
                        m = fmaxf(m, fabsf(xi[c]));
                    }
                }
                const float psc = k2_pixel_scale(m);  // the same per-pixel scale as the operand pre-pass and the FMA kernel
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    xr[c] *= psc;
                    if (CPLX) xi[c] *= psc;
                }
                float best = -1.f;
                int64_t win = -1;
#pragma unroll 1
                for (int i = 0; i < 2; ++i) {
                    const int gk = i ? g2k : g1k;
                    if (gk < 0) continue;
                    const int64_t kbase = p.a0 + ((int64_t)t0 * MT_BN) + (int64_t)gk * 16;
#pragma unroll 4
                    for (int j = 0; j < 16; ++j) {
                        const int64_t k = kbase + j;
                        if (k >= p.a1) break;  // zero rows that pad the last atom tile
                        float d[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) d[c] = __ldg(p.Dp + k * p.CP + c);
                        const float s = k2_score<C, CPLX>(d, xr, xi);
                        if (s > best || (s == best && k < win)) {
                            best = s;
                            win = k;
                        }
                    }
                }
                if (win >= 0 && best >= 0.f) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)win);
                    atomicMax(p.keys + pix, key);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, MT_TMEM_COLS);
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
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}
// [rows][32] fp32, K-major, box = 32 x 128 rows, 128-byte swizzle
int make_rows_map(CUtensorMap* out, const float* base, int64_t rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)MT_KF, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)MT_KF * 4};
    cuuint32_t box[2] = {(cuuint32_t)MT_KF, (cuuint32_t)MT_BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return qmri_fail(QMRI_ECUDA, "cuTensorMapEncodeTiled(%lld x 32 fp32) failed: %d", (long long)rows, (int)r);
    return QMRI_OK;
}

float host_tf32(float a) {  // round to nearest (ties away, like cvt.rna) to 10 explicit mantissa bits
    uint32_t b;
    memcpy(&b, &a, 4);
    if ((b & 0x7f800000u) == 0x7f800000u) return a;  // inf / nan unchanged
    b = (b + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &b, 4);
    return r;
}

template <int C>
int launch_c(qmri_ctx* ctx, const K2TcDict& d, const K2Params& p, float* A, int64_t npix_pad) {
    CUtensorMap tmA;
    QCHECK(make_rows_map(&tmA, A, 2 * npix_pad));
    MtParams m = {};
    m.x_re = p.x_re; m.x_im = p.x_im; m.npix = p.npix; m.npix_pad = npix_pad;
    m.Dp = p.Dp; m.a0 = p.a0; m.a1 = p.a1; m.ntiles = d.ntiles; m.CP = p.CP; m.keys = p.keys;
    m.ptiles = (int)(npix_pad / MT_BM);
    // split the atom range so that the work items fill whole waves of SMs (at least 8 atom tiles per item)
    int nsplit = 1;
    double best_eff = 0.0;
    const int max_split = std::max(1, std::min(d.ntiles / 8, 64));
    for (int cand = 1; cand <= max_split; ++cand) {
        const int64_t items = (int64_t)m.ptiles * cand;
        const int64_t waves = (items + ctx->sm_count - 1) / ctx->sm_count;
        const double eff = (double)items / (double)(waves * ctx->sm_count);
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            nsplit = cand;
        }
        if (waves >= 6) break;
    }
    m.nsplit = nsplit;
    static const int dbg = getenv("QMRI_K2_DEBUG") ? atoi(getenv("QMRI_K2_DEBUG")) : 0;  // attribution of the tile time: wrong results by design
    m.debug = dbg;
    static const int iss = getenv("QMRI_K2_ISSUERS") ? atoi(getenv("QMRI_K2_ISSUERS")) : 2;
    m.issuers = iss == 1 ? 1 : 2;
    const int grid = (int)std::min<int64_t>((int64_t)m.ptiles * nsplit, ctx->sm_count);
    // Epilogue warps: 8 (two per TMEM lane quarter) once the MMA side has two issuers and the atom range is long, 4 otherwise; measured on
    // the benchmark's dictionary, two issuers: 10 000 / 100 000 / 1 048 576 atoms: 4 warps 1.98 / 3.02 / 3.26, 8 warps 1.65 / 3.15 / 3.46
    // x 10^12 px-atoms/s (profiles/r02o_k2_issuers.txt).  QMRI_K2_EPI_WARPS=4|8 forces one.
    static const int ew_env = getenv("QMRI_K2_EPI_WARPS") ? atoi(getenv("QMRI_K2_EPI_WARPS")) : 0;
    const int ew = ew_env == 8 ? 8 : ew_env == 4 ? 4 : (p.x_im && m.issuers == 2 && (p.a1 - p.a0) >= 65536 ? 8 : 4);
    static bool configured_dev[QMRI_MAX_DEV][4] = {};
    bool& conf = configured_dev[qmri_dev_slot(ctx)][(p.x_im ? 1 : 0) + (ew == 8 ? 2 : 0)];
    const CUtensorMap& tmB = *reinterpret_cast<const CUtensorMap*>(d.mapB);
#define MT_LAUNCH(CPLX, EW)                                                                                                      \
    do {                                                                                                                          \
        if (!conf) QCUDA(cudaFuncSetAttribute(match_tc_kernel<C, CPLX, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM)); \
        conf = true;                                                                                                              \
        match_tc_kernel<C, CPLX, EW><<<grid, 96 + 32 * EW, MT_SMEM, ctx->stream>>>(tmA, tmB, m);                                  \
    } while (0)
    if (p.x_im) {
        if (ew == 8) MT_LAUNCH(true, 8);
        else MT_LAUNCH(true, 4);
    } else {
        if (ew == 8) MT_LAUNCH(false, 8);
        else MT_LAUNCH(false, 4);
    }
#undef MT_LAUNCH
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

}  // namespace

bool k2tc_supported(int C) { return C >= 1 && 3 * C <= MT_KF && get_encode() != nullptr; }

// Packs atoms [a0, a1) of D (host, K x C column-major with `ldD` rows resident starting at atom `row0`) as tf32-split K-major
// rows and uploads them; rows past a1 - a0 up to a multiple of 128 are zero.
int k2tc_dict_build(qmri_ctx* ctx, const float* D, int64_t ldD, int64_t row0, int C, int64_t a0, int64_t a1, K2TcDict* out) {
    static_assert(sizeof(CUtensorMap) <= sizeof(out->mapB), "tensor map storage");
    const int64_t n = a1 - a0;
    const int ntiles = (int)((n + MT_BN - 1) / MT_BN);
    std::vector<float> rows((size_t)ntiles * MT_BN * MT_KF, 0.f);
    for (int64_t k = 0; k < n; ++k) {
        float* r = rows.data() + (size_t)k * MT_KF;
        for (int c = 0; c < C; ++c) {
            const float v = D[(size_t)c * ldD + (a0 - row0) + k];
            const float hi = host_tf32(v);
            r[c] = hi;
            r[C + c] = hi;
            r[2 * C + c] = host_tf32(v - hi);
        }
    }
    QCHECK(dev_alloc(&out->Dt, rows.size()));
    QCUDA(cudaMemcpy(out->Dt, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
    out->ntiles = ntiles;
    return make_rows_map(reinterpret_cast<CUtensorMap*>(out->mapB), out->Dt, (int64_t)ntiles * MT_BN);
}
void k2tc_dict_free(K2TcDict* d) {
    if (d->Dt) cudaFree(d->Dt);
    d->Dt = nullptr;
    d->ntiles = 0;
}

size_t k2tc_stage_elems(int64_t npix) { return (size_t)2 * ((npix + MT_BM - 1) / MT_BM * MT_BM) * MT_KF; }

// keys (zero-initialised by the caller) <- max over atoms [p.a0, p.a1) - the range the packed dictionary `d` was built for
int k2tc_launch_keys(qmri_ctx* ctx, const K2TcDict& d, const K2Params& p, float* stage) {
    if (p.npix <= 0 || p.a1 <= p.a0) return QMRI_OK;
    const int64_t npix_pad = (p.npix + MT_BM - 1) / MT_BM * MT_BM;
    match_prep_kernel<<<dim3((unsigned)((npix_pad + 127) / 128), p.x_im ? 2 : 1), 128, 0, ctx->stream>>>(p.x_re, p.x_im, p.npix, npix_pad, p.C, stage);
    QLAUNCH_CHECK(ctx);
    switch (p.C) {
        case 1: return launch_c<1>(ctx, d, p, stage, npix_pad);
        case 2: return launch_c<2>(ctx, d, p, stage, npix_pad);
        case 3: return launch_c<3>(ctx, d, p, stage, npix_pad);
        case 4: return launch_c<4>(ctx, d, p, stage, npix_pad);
        case 5: return launch_c<5>(ctx, d, p, stage, npix_pad);
        case 6: return launch_c<6>(ctx, d, p, stage, npix_pad);
        case 7: return launch_c<7>(ctx, d, p, stage, npix_pad);
        case 8: return launch_c<8>(ctx, d, p, stage, npix_pad);
        case 9: return launch_c<9>(ctx, d, p, stage, npix_pad);
        case 10: return launch_c<10>(ctx, d, p, stage, npix_pad);
    }
    return qmri_fail(QMRI_EUNSUPPORTED, "tensor-pipe matching packs 3 C <= 32 operand slots: C = %d is served by the FP32 FMA kernel", p.C);
}
