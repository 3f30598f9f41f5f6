// K1 "streaming" variant - the fused x-update for large slice batches (sm_100a).
//
// Reference being replaced: the same lines as xupdate_kernel.cu -
//   main_files/algorithms/PnP_ADMM/PnP_ADMM.m:102-103 (lsqr x-update), :115-118 (v = real(x+u)),
//   :121 (global min/max for the 0-1 normalisation), :144 (u += x - v),
// with F.forward / F.adjoint of main_recon_tsmis_FFT.m:228-229 - in the closed form
//   x = z + A^H (y - A z) / (1 + rho)          (A A^H = I, see xupdate_phases.cuh).
//
// Work decomposition: ONE CTA per (slice, channel).  The CTA walks the eight 28-column slabs of the
// 224 x 224 image twice:
//   pass A  slab -> registers -> FFT along n -> sampled DFT along m, accumulated into the <= ~700 sample
//           values of the channel (shared memory);       then the data-consistency solve on those samples;
//   pass B  sparse inverse DFT along m -> inverse FFT along n -> registers -> w' = v + corr -> global.
// Compared with the cluster kernel (xupdate_kernel.cu) there is no cluster barrier and no DSMEM exchange,
// the operator tables are read once per image instead of once per slab, and the shared-memory traffic is
// about a third: the first / last FFT stage works on registers filled from / drained to global memory,
// and the sparse sums keep a k-space row in registers and rotate the twiddle instead of looking it up.
// HBM traffic per pixel-channel: read w (8 B) + v (4 B) in pass A, write w' (8 B) in pass B; pass B reads
// v a second time (4 B), normally from L2 - the CTA touched it a few tens of microseconds earlier.
// The kernel needs every k-space row of a frame to hold at most RMAX_STREAM samples (true for the spiral
// masks); line-sampled masks (EPI) and small slice batches stay on the cluster kernel.
#include <math.h>

#include "common.cuh"
#include "xupdate_kernel.h"
#include "xupdate_phases.cuh"

using namespace k1;

namespace {

constexpr int MC = 28;                // slab width (columns)
constexpr int THREADS = 16 * MC;      // 448: 28 column FFTs x 16 lanes; in the sparse passes thread = (k-space row, half of the slab)
constexpr int SLABS = NF / MC;        // 8
constexpr int HC = MC / 2;            // 14 columns per half slab

__device__ __forceinline__ void prefetch_l1(const float* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) xupdate_stream_kernel(K1Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* ws = reinterpret_cast<float2*>(smem_raw);                  // [MC][CS] slab workspace
    float2* tw = ws + MC * CS;                                         // [224]
    float2* tw2 = tw + NF;                                             // [16][16]
    float2* pc = tw2 + 256;                                            // [2][ns_max] sample accumulators (one per half slab), then c in [0]
    uint32_t* s_items = reinterpret_cast<uint32_t*>(pc + 2 * p.ns_max);  // [ns_max]
    uint16_t* s_rptr = reinterpret_cast<uint16_t*>(s_items + p.ns_max);  // [225]
    __shared__ float red_min[THREADS / 32], red_max[THREADS / 32];

    const int tid = threadIdx.x;
    const int c = blockIdx.x;
    const int s = blockIdx.y;
    const int f0 = p.frame_ptr[c];
    const int ns = p.frame_ptr[c + 1] - f0;
    const size_t img = ((size_t)s * p.C + c) * (size_t)(NF * NF);
    const float inv_n = 1.0f / (float)NF;  // unitary scaling 1/sqrt(N*M), once per transform direction

    for (int i = tid; i < NF; i += THREADS) tw[i] = p.tw[i];
    for (int i = tid; i < 256; i += THREADS) tw2[i] = p.tw2[i];
    for (int i = tid; i < ns; i += THREADS) {
        s_items[i] = p.items[f0 + i];
        if (MODE == K1_ADJOINT) {
            const float2 y = p.y[(size_t)s * p.nmeas + f0 + i];
            pc[i] = make_float2(y.x * inv_n, y.y * inv_n);
        } else {
            pc[i] = make_float2(0.f, 0.f);
            pc[p.ns_max + i] = make_float2(0.f, 0.f);
        }
    }
    for (int i = tid; i <= NF; i += THREADS) s_rptr[i] = p.row_ptr[(size_t)c * (NF + 1) + i];
    const int half = tid >= NF ? 1 : 0;   // warps 0-6: columns 0-13 of the slab, warps 7-13: columns 14-27
    const int rtid = tid - half * NF;
    const int k1row = p.rowmap[(size_t)c * NF + rtid];
    __syncthreads();
    const uint32_t* my_items = s_items + s_rptr[rtid];
    const int my_cnt = s_rptr[rtid + 1] - s_rptr[rtid];
    const int l16 = tid & 15;
    const int col = tid >> 4;

    // ---------------- pass A: forward transform sampled on the mask -------------------------------------
    if (MODE != K1_ADJOINT) {
#pragma unroll 1
        for (int slab = 0; slab < SLABS; ++slab) {
            const int m0 = slab * MC;
            {
                const size_t g = img + (size_t)(m0 + col) * NF + l16;
                float2 a[16];
                if (MODE == K1_ADMM) {  // z = 2 v - w
                    float wr[14], wi[14], vv[14];
#pragma unroll
                    for (int n1 = 0; n1 < 14; ++n1) {
                        wr[n1] = __ldcs(p.in_re + g + 16 * n1);
                        wi[n1] = __ldcs(p.in_im + g + 16 * n1);
                        vv[n1] = __ldg(p.v + g + 16 * n1);
                    }
#pragma unroll
                    for (int n1 = 0; n1 < 14; ++n1) a[n1] = make_float2(2.f * vv[n1] - wr[n1], -wi[n1]);
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < 14; ++n1)
                        a[n1] = make_float2(__ldcs(p.in_re + g + 16 * n1), p.in_im ? __ldcs(p.in_im + g + 16 * n1) : 0.f);
                }
                float2* colp = ws + col * CS;
                fwd_s1_regs(a, l16, tw2);
                fft_s1_store(colp, l16, a);
                __syncwarp();
                if (l16 < 14) fft_s2_load(colp, l16, a);
                __syncwarp();
                if (l16 < 14) fft_s2_store<false>(colp, l16, a);
            }
            __syncthreads();
            p3_row<HC>(ws + half * HC * CS, k1row, my_items, my_cnt, tw, m0 + half * HC, pc + half * p.ns_max);
            __syncthreads();
        }
        // data-consistency solve on the samples (each sample is owned by the thread of its row: no race above)
        for (int j = tid; j < ns; j += THREADS) {
            const size_t yi = (size_t)s * p.nmeas + f0 + j;
            const float2 az = make_float2((pc[j].x + pc[p.ns_max + j].x) * inv_n, (pc[j].y + pc[p.ns_max + j].y) * inv_n);
            if (MODE == K1_FORWARD) {
                p.y_out[yi] = az;
            } else {
                const float2 y = p.y[yi];
                const float gsc = p.inv_1p_rho * inv_n;  // (y - A z)/(1 + rho), pre-scaled for the inverse transform
                pc[j] = make_float2((y.x - az.x) * gsc, (y.y - az.y) * gsc);
            }
        }
        if (MODE == K1_FORWARD) return;
        __syncthreads();
    }

    // ---------------- pass B: corr = A^H c, epilogue ------------------------------------------------------
    float lmin = INFINITY, lmax = -INFINITY;
#pragma unroll 1
    for (int slab = 0; slab < SLABS; ++slab) {
        const int m0 = slab * MC;
        p4_row<HC>(ws + half * HC * CS, k1row, my_items, my_cnt, tw, m0 + half * HC, pc);
        __syncthreads();
        {
            const size_t g = img + (size_t)(m0 + col) * NF + l16;
            // what the correction is added to (v in ADMM mode, z in SOLVE mode) is wanted only after the transform: ask for
            // the lines now (no register cost), load them in the epilogue
            if (MODE == K1_ADMM) {
#pragma unroll
                for (int d = 0; d < 14; ++d) prefetch_l1(p.v + g + 16 * d);
            } else if (MODE == K1_SOLVE) {
#pragma unroll
                for (int d = 0; d < 14; ++d) {
                    prefetch_l1(p.in_re + g + 16 * d);
                    if (p.in_im) prefetch_l1(p.in_im + g + 16 * d);
                }
            }
            float2* colp = ws + col * CS;
            float2 a[16];
            if (l16 < 14) inv_s1_load(colp, l16, tw2, a);
            __syncwarp();
            if (l16 < 14) inv_s1_store(colp, l16, a);
            __syncwarp();
            inv_s2_regs(colp, l16, a);
            float br[14], bi[14];
            if (MODE == K1_ADMM) {
#pragma unroll
                for (int d = 0; d < 14; ++d) br[d] = __ldg(p.v + g + 16 * d);
            } else if (MODE == K1_SOLVE) {
#pragma unroll
                for (int d = 0; d < 14; ++d) {
                    br[d] = __ldg(p.in_re + g + 16 * d);
                    bi[d] = p.in_im ? __ldg(p.in_im + g + 16 * d) : 0.f;
                }
            }
            if (MODE == K1_ADMM && p.x_re) {  // last iteration: x = z + corr = 2 v - w + corr; w' may overwrite w in place, so x goes first
#pragma unroll
                for (int d = 0; d < 14; ++d) {
                    const float wr = __ldcs(p.in_re + g + 16 * d), wi = __ldcs(p.in_im + g + 16 * d);
                    p.x_re[g + 16 * d] = 2.f * br[d] - wr + a[d].x;
                    p.x_im[g + 16 * d] = a[d].y - wi;
                }
            }
#pragma unroll
            for (int d = 0; d < 14; ++d) {
                float ore, oim;
                if (MODE == K1_ADMM) {  // w' = v + corr
                    ore = br[d] + a[d].x;
                    oim = a[d].y;
                } else if (MODE == K1_SOLVE) {  // x = z + corr
                    ore = br[d] + a[d].x;
                    oim = bi[d] + a[d].y;
                } else {  // adjoint
                    ore = a[d].x;
                    oim = a[d].y;
                }
                p.out_re[g + 16 * d] = ore;
                p.out_im[g + 16 * d] = oim;
                lmin = fminf(lmin, ore);
                lmax = fmaxf(lmax, ore);
            }
        }
        __syncthreads();
    }
    if (p.minmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        if ((tid & 31) == 0) {
            red_min[tid >> 5] = lmin;
            red_max[tid >> 5] = lmax;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < THREADS / 32; ++w) {
                lmin = fminf(lmin, red_min[w]);
                lmax = fmaxf(lmax, red_max[w]);
            }
            atomicMin(p.minmax + 2 * s, float_to_ordered(lmin));
            atomicMax(p.minmax + 2 * s + 1, float_to_ordered(lmax));
        }
    }
}

template <int MODE>
int launch_mode(qmri_ctx* ctx, const K1Params& p, int S, size_t smem) {
    static size_t configured = 0;
    if (smem > configured) {
        QCUDA(cudaFuncSetAttribute(xupdate_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    xupdate_stream_kernel<MODE><<<dim3(p.C, S), THREADS, smem, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

}  // namespace

bool k1_stream_supported(int max_row, int ns_max) { return max_row <= RMAX_STREAM && ns_max <= 4096; }

int k1_stream_launch(qmri_ctx* ctx, const K1Params& p_in, int S, int ns_max) {
    if (S <= 0) return QMRI_OK;
    if (S > 65535) return qmri_fail(QMRI_EINVAL, "x-update: at most 65535 slices per launch (got %d)", S);
    K1Params p = p_in;
    p.ns_max = ns_max;
    const size_t smem = (size_t)(MC * CS + NF + 256 + 2 * ns_max) * sizeof(float2) + (size_t)ns_max * 4 + (size_t)(NF + 1) * 2 + 16;
    switch (p.mode) {
        case K1_ADMM: return launch_mode<K1_ADMM>(ctx, p, S, smem);
        case K1_SOLVE: return launch_mode<K1_SOLVE>(ctx, p, S, smem);
        case K1_FORWARD: return launch_mode<K1_FORWARD>(ctx, p, S, smem);
        default: return launch_mode<K1_ADJOINT>(ctx, p, S, smem);
    }
}
