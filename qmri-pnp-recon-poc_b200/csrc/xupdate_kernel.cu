// K1 - the fused x-update kernel (sm_100a, thread-block clusters + distributed shared memory).
//
// Reference being replaced: the whole of Step 1 and Step 3 of the ADMM loop,
//   main_files/algorithms/PnP_ADMM/PnP_ADMM.m:102-103 (lsqr x-update), :115-118 (v = real(x+u)),
//   :121 (global min/max for the 0-1 normalisation), :144 (u += x - v),
// with F.forward / F.adjoint of main_recon_tsmis_FFT.m:228-229.  See xupdate_phases.cuh for
// the maths.  One cluster of CL = 224/MC CTAs per (slice, channel); each slice-channel crosses
// HBM once per iteration: read w (8 B) + v (4 B), write w' (8 B) per pixel-channel.
#include <cooperative_groups.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "xupdate_kernel.h"
#include "xupdate_phases.cuh"

namespace cg = cooperative_groups;
using namespace k1;

template <int MC>
struct K1Cfg {
    static constexpr int THREADS = 8 * MC;          // 16 threads per column FFT, 2 rounds
    static constexpr int CL = NF / MC;              // cluster size
    static constexpr int GROUPS = THREADS / 16;     // concurrent column FFTs
    static constexpr int ROUNDS = MC / GROUPS;      // == 2
    static constexpr int F4_PER_THREAD = (MC * NF / 4) / THREADS;  // == 7
};

__device__ __forceinline__ float4 ldg_stream(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}

template <int MC>
__global__ void __launch_bounds__(K1Cfg<MC>::THREADS) xupdate_kernel(K1Params p) {
    using Cfg = K1Cfg<MC>;
    constexpr int THREADS = Cfg::THREADS;
    constexpr int CL = Cfg::CL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* cols = reinterpret_cast<float2*>(smem_raw);            // [MC][CS]
    float2* tw = cols + MC * CS;                                   // [224] twiddles (FFT stages)
    float2* twp = tw + NF;                                         // [238] padded copy (sparse passes)
    float2* pc = twp + TWP;                                        // [ns_max + 1] partial sums, then c; pc[ns_max] == 0
    uint32_t* s_p4 = reinterpret_cast<uint32_t*>(pc + p.ns_max + 1);  // [8][p4_len] work lists of the sparse inverse pass
    uint16_t* s_samp = reinterpret_cast<uint16_t*>(s_p4 + K1_PHASES * p.p4_len);  // [ns_max] k1 | k2 << 8
    __shared__ float red_min[32], red_max[32];

    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int rank = blockIdx.x;  // == cluster.block_rank(): cluster spans gridDim.x
    const int c = blockIdx.y;
    const int s = blockIdx.z;
    const int m0 = rank * MC;
    const int mode = p.mode;
    const int f0 = p.frame_ptr[c];
    const int ns = p.frame_ptr[c + 1] - f0;
    const size_t plane = (size_t)NF * NF;
    const size_t slab = ((size_t)(s * p.C + c)) * plane + (size_t)m0 * NF;  // contiguous MC*224 floats

    // per-frame operator tables -> shared memory (the sparse passes walk them once per column / sample)
    for (int i = tid; i < NF; i += THREADS) {
        const float2 t = p.tw[i];
        tw[i] = t;
        twp[i + (i >> 4)] = t;
    }
    for (int i = tid; i < ns; i += THREADS) s_samp[i] = p.samp[f0 + i];
    for (int i = tid; i < K1_PHASES * p.p4_len; i += THREADS) s_p4[i] = p.p4tab[(size_t)c * K1_PHASES * p.p4_len + i];
    if (tid == 0) pc[p.ns_max] = make_float2(0.f, 0.f);

    // ---- P1: load z into shared memory ---------------------------------------------
    if (mode != K1_ADJOINT) {
#pragma unroll
        for (int i = 0; i < Cfg::F4_PER_THREAD; ++i) {
            int q = tid + i * THREADS;
            int e = 4 * q;
            int mm = e / NF, n = e - mm * NF;
            float4 re = ldg_stream(p.in_re + slab + e);
            float4 im = p.in_im ? ldg_stream(p.in_im + slab + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (mode == K1_ADMM) {  // z = 2 v - w
                float4 v = __ldg(reinterpret_cast<const float4*>(p.v + slab + e));
                re = make_float4(2.f * v.x - re.x, 2.f * v.y - re.y, 2.f * v.z - re.z, 2.f * v.w - re.w);
                im = make_float4(-im.x, -im.y, -im.z, -im.w);
            }
            float2* d = cols + mm * CS + n;
            d[0] = make_float2(re.x, im.x);
            d[1] = make_float2(re.y, im.y);
            d[2] = make_float2(re.z, im.z);
            d[3] = make_float2(re.w, im.w);
        }
        __syncthreads();

        // ---- P2: forward FFT along n for every column of the slab ---------------------
        {
            const int l16 = tid & 15;
#pragma unroll 1
            for (int rd = 0; rd < Cfg::ROUNDS; ++rd) {
                float2* col = cols + (rd * Cfg::GROUPS + (tid >> 4)) * CS;
                float2 a[16];
                fft_s1_load<false>(col, l16, tw, a);
                __syncwarp();
                fft_s1_store(col, l16, a);
                __syncwarp();
                if (l16 < 14) fft_s2_load(col, l16, a);
                __syncwarp();
                if (l16 < 14) fft_s2_store<false>(col, l16, a);
            }
        }
        __syncthreads();

        // ---- P3: sampled DFT along m (partial over this slab) ---------------------------
        for (int j = tid; j < ns; j += THREADS) {
            uint16_t kk = s_samp[j];
            pc[j] = sampled_dft_partial<MC>(cols, twp, m0, kk & 0xff, kk >> 8);
        }
        cluster.sync();

        // ---- reduce-scatter over the cluster, data-consistency solve ---------------------
        const int chunk = (ns + CL - 1) / CL;
        const int j0 = rank * chunk, j1 = min(ns, j0 + chunk);
        const float inv_n = 1.0f / (float)NF;  // unitary scaling 1/sqrt(N*M)
        for (int j = j0 + tid; j < j1; j += THREADS) {
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int r = 0; r < CL; ++r) {
                const float2* rp = cluster.map_shared_rank(pc, r);
                float2 t = rp[j];
                sx += t.x;
                sy += t.y;
            }
            sx *= inv_n;
            sy *= inv_n;
            const size_t yi = (size_t)s * p.nmeas + f0 + j;
            if (mode == K1_FORWARD) {
                p.y_out[yi] = make_float2(sx, sy);
            } else {
                float2 y = p.y[yi];
                float g = p.inv_1p_rho * inv_n;  // (y - Az)/(1+rho), pre-scaled for the inverse transform
                pc[j] = make_float2((y.x - sx) * g, (y.y - sy) * g);
            }
        }
        cluster.sync();
        if (mode == K1_FORWARD) return;  // uniform across the cluster; no remote access after the sync above
        // ---- all-gather of c ----------------------------------------------------------------
#pragma unroll 1
        for (int r = 0; r < CL; ++r) {
            if (r == rank) continue;
            const float2* rp = cluster.map_shared_rank(pc, r);
            const int a0 = r * chunk, a1 = min(ns, a0 + chunk);
            for (int j = a0 + tid; j < a1; j += THREADS) pc[j] = rp[j];
        }
        cluster.sync();  // also the last remote access: CTAs may finish independently from here
    } else {
        const float inv_n = 1.0f / (float)NF;
        for (int j = tid; j < ns; j += THREADS) {
            float2 y = p.y[(size_t)s * p.nmeas + f0 + j];
            pc[j] = make_float2(y.x * inv_n, y.y * inv_n);
        }
        __syncthreads();
    }

    // ---- P4: sparse inverse DFT along m: every (column, row) of the slab is rewritten ------
    {
        const int mm = tid % MC, ph = tid / MC;  // THREADS = 8 MC: column mm, work list ph
        sparse_idft_flat(cols + mm * CS, pc, twp, s_p4 + ph * p.p4_len, p.p4_len, m0 + mm);
    }
    __syncthreads();

    // ---- P5: inverse FFT along n -----------------------------------------------------------
    {
        const int l16 = tid & 15;
#pragma unroll 1
        for (int rd = 0; rd < Cfg::ROUNDS; ++rd) {
            float2* col = cols + (rd * Cfg::GROUPS + (tid >> 4)) * CS;
            float2 a[16];
            fft_s1_load<true>(col, l16, tw, a);
            __syncwarp();
            fft_s1_store(col, l16, a);
            __syncwarp();
            if (l16 < 14) fft_s2_load(col, l16, a);
            __syncwarp();
            if (l16 < 14) fft_s2_store<true>(col, l16, a);
        }
    }
    __syncthreads();

    // ---- P6: epilogue: w' = v + corr (dual update folded in), real-part min/max -------------
    float lmin = INFINITY, lmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < Cfg::F4_PER_THREAD; ++i) {
        int q = tid + i * THREADS;
        int e = 4 * q;
        int mm = e / NF, n = e - mm * NF;
        const float2* d = cols + mm * CS + n;
        float2 c0 = d[0], c1 = d[1], c2 = d[2], c3 = d[3];
        float4 ore, oim;
        if (mode == K1_ADMM) {
            float4 v = __ldg(reinterpret_cast<const float4*>(p.v + slab + e));  // L2 hit: read in P1
            ore = make_float4(v.x + c0.x, v.y + c1.x, v.z + c2.x, v.w + c3.x);
            oim = make_float4(c0.y, c1.y, c2.y, c3.y);
            if (p.x_re) {  // last iteration: x = z + corr = 2 v - w + corr
                float4 wr = ldg_stream(p.in_re + slab + e);
                float4 wi = ldg_stream(p.in_im + slab + e);
                float4 xr = make_float4(2.f * v.x - wr.x + c0.x, 2.f * v.y - wr.y + c1.x, 2.f * v.z - wr.z + c2.x,
                                        2.f * v.w - wr.w + c3.x);
                float4 xi = make_float4(c0.y - wi.x, c1.y - wi.y, c2.y - wi.z, c3.y - wi.w);
                *reinterpret_cast<float4*>(p.x_re + slab + e) = xr;
                *reinterpret_cast<float4*>(p.x_im + slab + e) = xi;
            }
        } else if (mode == K1_SOLVE) {  // x = z + corr
            float4 zr = ldg_stream(p.in_re + slab + e);
            float4 zi = p.in_im ? ldg_stream(p.in_im + slab + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            ore = make_float4(zr.x + c0.x, zr.y + c1.x, zr.z + c2.x, zr.w + c3.x);
            oim = make_float4(zi.x + c0.y, zi.y + c1.y, zi.z + c2.y, zi.w + c3.y);
        } else {  // adjoint
            ore = make_float4(c0.x, c1.x, c2.x, c3.x);
            oim = make_float4(c0.y, c1.y, c2.y, c3.y);
        }
        *reinterpret_cast<float4*>(p.out_re + slab + e) = ore;
        *reinterpret_cast<float4*>(p.out_im + slab + e) = oim;
        lmin = fminf(lmin, fminf(fminf(ore.x, ore.y), fminf(ore.z, ore.w)));
        lmax = fmaxf(lmax, fmaxf(fmaxf(ore.x, ore.y), fmaxf(ore.z, ore.w)));
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

template <int MC>
static int launch_mc(qmri_ctx* ctx, const K1Params& p, int S, int ns_max) {
    using Cfg = K1Cfg<MC>;
    static_assert(Cfg::THREADS / MC == K1_PHASES, "one work list per row phase");
    size_t smem = (size_t)(MC * CS + NF + TWP + ns_max + 1) * sizeof(float2) + (size_t)K1_PHASES * p.p4_len * 4 + (size_t)ns_max * 2 + 16;
    static size_t configured_dev[QMRI_MAX_DEV] = {};  // the table sizes depend on the operator: raise the limit when a larger one comes along
    size_t& configured = configured_dev[qmri_dev_slot(ctx)];
    if (smem > configured) {
        QCUDA(cudaFuncSetAttribute(xupdate_kernel<MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(Cfg::CL, p.C, S);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QCUDA(cudaLaunchKernelEx(&cfg, xupdate_kernel<MC>, p));
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

int k1_launch(qmri_ctx* ctx, const K1Params& p_in, int S, int ns_max, int mc) {
    if (S <= 0) return QMRI_OK;
    K1Params p = p_in;
    p.ns_max = ns_max;
    if (mc == 56) return launch_mc<56>(ctx, p, S, ns_max);
    return launch_mc<28>(ctx, p, S, ns_max);
}

__global__ void minmax_init_kernel(int* mm, int S) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) {
        mm[2 * i] = 0x7fffffff;
        mm[2 * i + 1] = (int)0x80000000;
    }
}

int k1_minmax_init(qmri_ctx* ctx, int* minmax, int S) {
    minmax_init_kernel<<<(S + 127) / 128, 128, 0, ctx->stream>>>(minmax, S);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
