// K2 - dictionary matching: fused contraction + |.|^2 + running (max, idx); the K x B score
// matrix of the reference is never materialised.
//
// Reference being replaced: main_files/dictionary_matching/mrf_dtm_cpu.m:84-98
//   ip = dict.D * ctranspose(x(cind,:));  [mt, dm] = max(abs(ip),[],1);  pd = ip(dm) ./ normD(dm)
// and :136-148 (LUT gather, NaN -> 0).
//
// FP32-FMA pipe: every thread keeps PX pixels' C-channel complex signatures in registers
// (2*C*PX floats); atoms are staged through shared memory in chunks and read as warp-wide
// broadcasts (LDS.128).  Per (pixel, atom) the SM issues 2*C FFMA + FMUL + FFMA for |.|^2 and ONE
// FMNMX: only the running maximum of a group of 16 atoms is tracked in the inner loop, the
// (max, group) compare/select happens once per group, and the winning group is rescanned at the end
// with bit-identical arithmetic to recover the first atom attaining the maximum (MATLAB's
// first-index tie rule: strict > across groups keeps the first group, the rescan the first atom).
// The result is carried as a packed 64-bit key
//     key = float_bits(score^2) << 32 | (0xFFFFFFFF - atom_index)
// whose integer max is "largest score, lowest index on ties"; the same key is what atom-sharded
// ranks reduce with an integer max (qmri.h, BASELINE config 5).
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "match_kernel.h"
#include "match_score.cuh"

namespace {

constexpr int K2_THREADS = 256;
constexpr int K2_CHUNK = 512;  // atoms per shared-memory stage
constexpr int K2_GROUP = 16;   // atoms per (max, index) bookkeeping step


template <int C, int CP, int PX, bool CPLX>
__global__ void __launch_bounds__(K2_THREADS) match_kernel(K2Params p) {
    __shared__ __align__(16) float atoms[K2_CHUNK * CP];
    const int tid = threadIdx.x;
    const int64_t pix0 = (int64_t)blockIdx.x * (K2_THREADS * PX);
    // this block's atom range
    const int64_t per = (p.a1 - p.a0 + gridDim.y - 1) / gridDim.y;
    const int64_t ka = p.a0 + (int64_t)blockIdx.y * per;
    const int64_t kb = min(p.a1, ka + per);

    float xr[PX][C], xi[PX][CPLX ? C : 1];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        int64_t pix = pix0 + tid + (int64_t)i * K2_THREADS;
        bool ok = pix < p.npix;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            xr[i][c] = (ok && c < p.C) ? __ldg(p.x_re + (int64_t)c * p.npix + pix) : 0.f;
            if (CPLX) xi[i][c] = (ok && c < p.C) ? __ldg(p.x_im + (int64_t)c * p.npix + pix) : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < PX; ++i) {  // per-pixel power-of-two scale (match_score.cuh): keeps |<d, x>|^2 inside the fp32 range
        float m = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            m = fmaxf(m, fabsf(xr[i][c]));
            if (CPLX) m = fmaxf(m, fabsf(xi[i][c]));
        }
        const float sc = k2_pixel_scale(m);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            xr[i][c] *= sc;
            if (CPLX) xi[i][c] *= sc;
        }
    }
    float best[PX];
    int bgrp[PX];  // winning group, counted from ka in units of K2_GROUP atoms
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        best[i] = -1.f;
        bgrp[i] = 0;
    }

    for (int64_t k0 = ka; k0 < kb; k0 += K2_CHUNK) {
        const int nk = (int)min((int64_t)K2_CHUNK, kb - k0);
        const int ngrp = (nk + K2_GROUP - 1) / K2_GROUP;
        __syncthreads();
        {
            const float4* src = reinterpret_cast<const float4*>(p.Dp + k0 * CP);
            float4* dst = reinterpret_cast<float4*>(atoms);
            const int n4 = nk * CP / 4, n4pad = ngrp * K2_GROUP * CP / 4;
            // atoms past the range are zero: their score 0 can never be strictly greater than a real score
            for (int i = tid; i < n4pad; i += K2_THREADS) dst[i] = (i < n4) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const int g0 = (int)((k0 - ka) / K2_GROUP);
#pragma unroll 1
        for (int g = 0; g < ngrp; ++g) {
            float gmax[PX];
#pragma unroll
            for (int i = 0; i < PX; ++i) gmax[i] = -1.f;
#pragma unroll 4
            for (int a = 0; a < K2_GROUP; ++a) {
                float d[CP];
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    float4 t = *reinterpret_cast<const float4*>(atoms + (g * K2_GROUP + a) * CP + 4 * q);
                    d[4 * q] = t.x;
                    d[4 * q + 1] = t.y;
                    d[4 * q + 2] = t.z;
                    d[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < PX; ++i) gmax[i] = fmaxf(gmax[i], k2_score<C, CPLX>(d, xr[i], xi[i]));  // fmaxf drops NaN
            }
#pragma unroll
            for (int i = 0; i < PX; ++i)
                if (gmax[i] > best[i]) {  // strict: the first group wins ties
                    best[i] = gmax[i];
                    bgrp[i] = g0 + g;
                }
        }
    }
    // rescan the winning group: first atom whose score equals the maximum
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        int64_t pix = pix0 + tid + (int64_t)i * K2_THREADS;
        if (pix >= p.npix || !(best[i] >= 0.f)) continue;
        const int64_t kg = ka + (int64_t)bgrp[i] * K2_GROUP;
        int64_t win = kg;
        bool found = false;
        for (int a = 0; a < K2_GROUP; ++a) {
            const int64_t k = kg + a;
            if (k >= kb || found) break;
            float d[CP];
#pragma unroll
            for (int q = 0; q < CP / 4; ++q) {
                float4 t = __ldg(reinterpret_cast<const float4*>(p.Dp + k * CP + 4 * q));
                d[4 * q] = t.x;
                d[4 * q + 1] = t.y;
                d[4 * q + 2] = t.z;
                d[4 * q + 3] = t.w;
            }
            if (k2_score<C, CPLX>(d, xr[i], xi[i]) == best[i]) {
                win = k;
                found = true;
            }
        }
        unsigned long long key = ((unsigned long long)__float_as_uint(best[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)win);
        atomicMax(p.keys + pix, key);
    }
}

// From the (reduced) keys: recompute <d, x> for the winner in FP32 and gather the outputs.
__global__ void match_finish_kernel(K2Finish p) {
    int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.npix) return;
    unsigned long long key = p.keys[pix];
    int64_t idx = (int64_t)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    if (key == 0ull || idx >= p.K) idx = 0;  // all-NaN pixel: MATLAB's max returns index 1
    if (idx < p.own0 || idx >= p.own1) {       // atom-sharded dictionary: the owner of the winning atom writes this pixel
        if (p.pd) p.pd[2 * pix] = p.pd[2 * pix + 1] = 0.f;
        if (p.mt) p.mt[pix] = 0.f;
        if (p.dm) p.dm[pix] = 0;
        if (p.qmap)
            for (int q = 0; q < p.Q; ++q) p.qmap[(int64_t)q * p.npix + pix] = 0.f;
        return;
    }
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
    if (p.mt) p.mt[pix] = hypotf(sr, si);  // no intermediate square: |ip| keeps the whole fp32 range, like abs(ip) in the reference
    if (p.dm) p.dm[pix] = (int32_t)idx + 1;
    if (p.qmap)
        for (int q = 0; q < p.Q; ++q) {
            float v = __ldg(p.lut + (int64_t)q * p.K + idx);
            p.qmap[(int64_t)q * p.npix + pix] = isnan(v) ? 0.f : v;
        }
}

template <int C, int CP>
int launch_c(qmri_ctx* ctx, const K2Params& p) {
    constexpr int PX = 4;
    const int64_t ppb = (int64_t)K2_THREADS * PX;
    int64_t gx = (p.npix + ppb - 1) / ppb;
    // split the atom range so that the grid fills whole waves of resident CTAs (small pixel counts would
    // otherwise leave SMs idle or end with a mostly empty last wave)
    static int per_sm_dev[QMRI_MAX_DEV] = {};
    int& per_sm = per_sm_dev[qmri_dev_slot(ctx)];
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel<C, CP, PX, true>, K2_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const int64_t slots = (int64_t)ctx->sm_count * per_sm;
    int64_t natoms = p.a1 - p.a0;
    int64_t maxsplit = std::max<int64_t>(1, std::min<int64_t>(natoms / (2 * K2_CHUNK), 4096));
    int gy = 1;
    double best_eff = 0.0;
    for (int64_t cand = 1; cand <= maxsplit; ++cand) {
        const int64_t ctas = gx * cand;
        const int64_t waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots);
        if (eff > best_eff + 0.02) {  // prefer fewer splits unless clearly better
            best_eff = eff;
            gy = (int)cand;
        }
        if (waves >= 8) break;
    }
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (p.x_im) match_kernel<C, CP, PX, true><<<grid, K2_THREADS, 0, ctx->stream>>>(p);
    else match_kernel<C, CP, PX, false><<<grid, K2_THREADS, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// K4 - TSMI synthesis, the mirror image of K2 (argmin of a distance instead of argmax of a correlation).
//
// Reference being replaced: main_synthesize_tsmis.m:84-98
//   I = knnsearch(KDTreeSearcher(dict.lut), qm(:,1:2));  X = real(dict.D(I,1:10)) .* dict.normD(I) .* abs(qm(:,3));
//   X = X .* sign(X(:,:,1))
// The KD-tree is replaced by an exhaustive fused scan: every thread keeps PX pixels' (T1, T2) in registers, the
// (T1, T2) pairs of the atoms stream through shared memory, per (pixel, atom) 2 FADD + FMUL + FFMA + one FMNMX
// (running minimum of a 16-atom group; the group is rescanned for the first atom attaining it).  The packed key
// float_bits(d^2) << 32 | atom merges atom ranges with atomicMin (smallest distance, lowest index on exact ties).
// ------------------------------------------------------------------------------------------------
constexpr int K4_CHUNK = 2048;

template <int PX>
__global__ void __launch_bounds__(K2_THREADS) synth_nn_kernel(K4Params p) {
    __shared__ float s1[K4_CHUNK], s2[K4_CHUNK];
    const int tid = threadIdx.x;
    const int64_t pix0 = (int64_t)blockIdx.x * (K2_THREADS * PX);
    const int64_t per = (p.K + gridDim.y - 1) / gridDim.y;
    const int64_t ka = (int64_t)blockIdx.y * per, kb = min(p.K, ka + per);
    float a1[PX], a2[PX], best[PX];
    int bgrp[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        const int64_t pix = pix0 + tid + (int64_t)i * K2_THREADS;
        const bool ok = pix < p.npix;
        a1[i] = ok ? __ldg(p.t1 + pix) : 0.f;
        a2[i] = ok ? __ldg(p.t2 + pix) : 0.f;
        best[i] = INFINITY;
        bgrp[i] = 0;
    }
    for (int64_t k0 = ka; k0 < kb; k0 += K4_CHUNK) {
        const int nk = (int)min((int64_t)K4_CHUNK, kb - k0);
        const int ngrp = (nk + K2_GROUP - 1) / K2_GROUP;
        __syncthreads();
        for (int i = tid; i < ngrp * K2_GROUP; i += K2_THREADS) {  // atoms past the range sit at infinity: never nearest
            s1[i] = (i < nk) ? __ldg(p.lut + k0 + i) : INFINITY;
            s2[i] = (i < nk) ? __ldg(p.lut + p.K + k0 + i) : INFINITY;
        }
        __syncthreads();
        const int g0 = (int)((k0 - ka) / K2_GROUP);
#pragma unroll 1
        for (int g = 0; g < ngrp; ++g) {
            float gmin[PX];
#pragma unroll
            for (int i = 0; i < PX; ++i) gmin[i] = INFINITY;
#pragma unroll
            for (int a = 0; a < K2_GROUP; ++a) {
                const float l1 = s1[g * K2_GROUP + a], l2 = s2[g * K2_GROUP + a];
#pragma unroll
                for (int i = 0; i < PX; ++i) {
                    const float d1 = a1[i] - l1, d2 = a2[i] - l2;
                    gmin[i] = fminf(gmin[i], fmaf(d2, d2, d1 * d1));  // fminf drops NaN (NaN LUT rows never match)
                }
            }
#pragma unroll
            for (int i = 0; i < PX; ++i)
                if (gmin[i] < best[i]) {  // strict: the first group wins exact ties
                    best[i] = gmin[i];
                    bgrp[i] = g0 + g;
                }
        }
    }
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        const int64_t pix = pix0 + tid + (int64_t)i * K2_THREADS;
        if (pix >= p.npix || !(best[i] < INFINITY)) continue;
        const int64_t kg = ka + (int64_t)bgrp[i] * K2_GROUP;
        int64_t win = kg;
        for (int a = 0; a < K2_GROUP; ++a) {
            const int64_t k = kg + a;
            if (k >= kb) break;
            const float d1 = a1[i] - __ldg(p.lut + k), d2 = a2[i] - __ldg(p.lut + p.K + k);
            if (fmaf(d2, d2, d1 * d1) == best[i]) {
                win = k;
                break;
            }
        }
        const unsigned long long key = ((unsigned long long)__float_as_uint(best[i]) << 32) | (unsigned long long)(unsigned)win;
        atomicMin(p.keys + pix, key);
    }
}

// X = D(I,:) .* normD(I) .* |PD|, then X .* sign(X(:,1))   (main_synthesize_tsmis.m:89-98)
__global__ void synth_render_kernel(K4Params p) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.npix) return;
    const unsigned long long key = p.keys[pix];
    int64_t idx = (int64_t)(key & 0xFFFFFFFFull);
    if (key == ~0ull || idx >= p.K) idx = 0;  // NaN query: knnsearch still returns an index; take the first atom
    const float sc = __ldg(p.normD + idx) * fabsf(__ldg(p.pd + pix));
    const float x0 = __ldg(p.Dp + idx * p.CP) * sc;
    const float sg = (x0 > 0.f) ? 1.f : (x0 < 0.f ? -1.f : 0.f);  // MATLAB sign(): sign(0) = 0
    for (int c = 0; c < p.C; ++c) p.X[(int64_t)c * p.npix + pix] = __ldg(p.Dp + idx * p.CP + c) * sc * sg;
    if (p.index) p.index[pix] = (int32_t)idx + 1;
}

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

int k4_launch(qmri_ctx* ctx, const K4Params& p) {
    if (p.npix <= 0) return QMRI_OK;
    constexpr int PX = 4;
    const int64_t ppb = (int64_t)K2_THREADS * PX;
    const int64_t gx = (p.npix + ppb - 1) / ppb;
    int64_t gy = std::max<int64_t>(1, std::min<int64_t>((2LL * ctx->sm_count + gx - 1) / gx, std::max<int64_t>(1, p.K / K4_CHUNK)));
    QCUDA(cudaMemsetAsync(p.keys, 0xFF, (size_t)p.npix * 8, ctx->stream));
    synth_nn_kernel<PX><<<dim3((unsigned)gx, (unsigned)gy), K2_THREADS, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    synth_render_kernel<<<(unsigned)((p.npix + 255) / 256), 256, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
