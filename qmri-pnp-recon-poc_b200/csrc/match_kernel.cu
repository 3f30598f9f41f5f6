// K2 - dictionary matching: fused contraction + |.|^2 + running (max, idx); the K x B score
// matrix of the reference is never materialised.
//
// Reference being replaced: main_files/dictionary_matching/mrf_dtm_cpu.m:84-98
//   ip = dict.D * ctranspose(x(cind,:));  [mt, dm] = max(abs(ip),[],1);  pd = ip(dm) ./ normD(dm)
// and :136-148 (LUT gather, NaN -> 0).
//
// FP32-FMA pipe: every thread keeps PX pixels' C-channel complex signatures in registers
// (2*C*PX floats); atoms are staged through shared memory in chunks and read as warp-wide
// broadcasts (LDS.128), so per (pixel, atom) the SM issues 2*C FFMA + 2 for |.|^2 + a compare/
// select.  The running best is carried as a packed 64-bit key
//     key = float_bits(score^2) << 32 | (0xFFFFFFFF - atom_index)
// whose integer max is "largest score, lowest index on ties" = MATLAB's first-index rule; the
// same key is what atom-sharded ranks reduce with an integer max (qmri.h, BASELINE config 5).
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "match_kernel.h"

namespace {

constexpr int K2_THREADS = 256;
constexpr int K2_CHUNK = 512;  // atoms per shared-memory stage

template <int C, int CP, int PX, bool CPLX>
__global__ void __launch_bounds__(K2_THREADS) match_kernel(K2Params p) {
    __shared__ __align__(16) float atoms[K2_CHUNK * CP];
    const int tid = threadIdx.x;
    const int64_t pix0 = (int64_t)blockIdx.x * (K2_THREADS * PX);
    // this block's atom range
    const int64_t per = (p.a1 - p.a0 + gridDim.y - 1) / gridDim.y;
    const int64_t ka = p.a0 + (int64_t)blockIdx.y * per;
    const int64_t kb = min(p.a1, ka + per);

    float xr[PX][C], xi[PX][C];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        int64_t pix = pix0 + tid + (int64_t)i * K2_THREADS;
        bool ok = pix < p.npix;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            xr[i][c] = (ok && c < p.C) ? __ldg(p.x_re + (int64_t)c * p.npix + pix) : 0.f;
            xi[i][c] = (CPLX && ok && c < p.C) ? __ldg(p.x_im + (int64_t)c * p.npix + pix) : 0.f;
        }
    }
    float best[PX];
    int bidx[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        best[i] = -1.f;
        bidx[i] = 0;
    }

    for (int64_t k0 = ka; k0 < kb; k0 += K2_CHUNK) {
        const int nk = (int)min((int64_t)K2_CHUNK, kb - k0);
        __syncthreads();
        {
            const float4* src = reinterpret_cast<const float4*>(p.Dp + k0 * CP);
            float4* dst = reinterpret_cast<float4*>(atoms);
            const int n4 = nk * CP / 4;
            for (int i = tid; i < n4; i += K2_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
#pragma unroll 2
        for (int a = 0; a < nk; ++a) {
            float d[CP];
#pragma unroll
            for (int q = 0; q < CP / 4; ++q) {
                float4 t = *reinterpret_cast<const float4*>(atoms + a * CP + 4 * q);
                d[4 * q] = t.x;
                d[4 * q + 1] = t.y;
                d[4 * q + 2] = t.z;
                d[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                float sr = 0.f, si = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    sr = fmaf(d[c], xr[i][c], sr);
                    if (CPLX) si = fmaf(d[c], xi[i][c], si);
                }
                float sc = CPLX ? fmaf(si, si, sr * sr) : sr * sr;
                if (sc > best[i]) {  // strict: the first (lowest) index wins ties
                    best[i] = sc;
                    bidx[i] = (int)(k0 - p.a0) + a;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        int64_t pix = pix0 + tid + (int64_t)i * K2_THREADS;
        if (pix < p.npix && best[i] >= 0.f) {
            unsigned long long key = ((unsigned long long)__float_as_uint(best[i]) << 32) |
                                     (unsigned long long)(0xFFFFFFFFu - (unsigned)(p.a0 + bidx[i]));
            atomicMax(p.keys + pix, key);
        }
    }
}

// From the (reduced) keys: recompute <d, x> for the winner in FP32 and gather the outputs.
__global__ void match_finish_kernel(K2Finish p) {
    int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.npix) return;
    unsigned long long key = p.keys[pix];
    int64_t idx = (int64_t)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    if (key == 0ull || idx >= p.K) idx = 0;  // all-NaN pixel: MATLAB's max returns index 1
    float sr = 0.f, si = 0.f;
    for (int c = 0; c < p.C; ++c) {
        float d = __ldg(p.Dp + idx * p.CP + c);
        sr = fmaf(d, __ldg(p.x_re + (int64_t)c * p.npix + pix), sr);
        if (p.x_im) si = fmaf(d, __ldg(p.x_im + (int64_t)c * p.npix + pix), si);
    }
    // ip = D * ctranspose(x): the pixel signature enters conjugated
    if (p.pd) {
        float nd = __ldg(p.normD + idx);
        p.pd[2 * pix] = sr / nd;
        p.pd[2 * pix + 1] = -si / nd;
    }
    if (p.mt) p.mt[pix] = sqrtf(fmaf(si, si, sr * sr));
    if (p.dm) p.dm[pix] = (int32_t)idx + 1;
    if (p.qmap)
        for (int q = 0; q < p.Q; ++q) {
            float v = __ldg(p.lut + (int64_t)q * p.K + idx);
            p.qmap[(int64_t)q * p.npix + pix] = isnan(v) ? 0.f : v;
        }
}

template <int C, int CP>
int launch_c(qmri_ctx* ctx, const K2Params& p) {
    constexpr int PX = 2;
    const int64_t ppb = (int64_t)K2_THREADS * PX;
    int64_t gx = (p.npix + ppb - 1) / ppb;
    // split the atom range so that small pixel counts still fill the machine (>= 2 waves)
    int64_t natoms = p.a1 - p.a0;
    int64_t want = (2LL * ctx->sm_count * 2 + gx - 1) / gx;  // 2 CTAs/SM resident
    int64_t maxsplit = std::max<int64_t>(1, natoms / (4 * K2_CHUNK));
    int gy = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, maxsplit), 65535));
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (p.x_im) match_kernel<C, CP, PX, true><<<grid, K2_THREADS, 0, ctx->stream>>>(p);
    else match_kernel<C, CP, PX, false><<<grid, K2_THREADS, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

}  // namespace

int k2_padded_channels(int C) { return (C + 3) / 4 * 4; }

int k2_launch_keys(qmri_ctx* ctx, const K2Params& p) {
    if (p.npix <= 0 || p.a1 <= p.a0) return QMRI_OK;
    switch (p.CP) {
        case 4: return launch_c<4, 4>(ctx, p);
        case 8: return launch_c<8, 8>(ctx, p);
        case 12: return p.C == 10 ? launch_c<10, 12>(ctx, p) : launch_c<12, 12>(ctx, p);
        case 16: return launch_c<16, 16>(ctx, p);
    }
    return qmri_fail(QMRI_EUNSUPPORTED, "dictionary matching supports at most 16 channels (got %d)", p.C);
}

int k2_launch_finish(qmri_ctx* ctx, const K2Finish& p) {
    if (p.npix <= 0) return QMRI_OK;
    match_finish_kernel<<<(unsigned)((p.npix + 255) / 256), 256, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
