// Kernels either side of the reconstruction loop (SURVEY.md 8f-1, 8f-4 and the P.for / P.adj handles of 8b):
//
//  * measurement noise      Y = awgn(Y, snr, 'measured')                         main_recon_tsmis_FFT.m:243
//  * the bare sampling matrix  P.for(vec) = P * vec, P.adj(y) = P' * y           setup_subsampling_spiralgrided.m:36-42,
//                                                                              setup_subsampling_epi.m:31-35
//  * foreground mask        getmask_fromPD(PD, thresh)                           main_files/utils/getmask_fromPD.m:1-15
//  * quality metrics        masked MAE, psnr(), ssim() of T1 / T2 / PD and TSMIs  main_recon_tsmis_FFT.m:328-384
//
// None of this is on the per-iteration hot path: a slice is touched once.  Everything runs in double precision (the
// reference computes these in double) with fixed-order reductions, so results are reproducible run to run.
#include <math.h>

#include "aux_kernels.h"
#include "common.cuh"

namespace {

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// ------------------------------------------------------------------------------------------------
// layout conversion: interleaved host dtype <-> planar fp64
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void unpack_f64_kernel(const T* __restrict__ src, double* __restrict__ re, double* __restrict__ im, size_t n, int cplx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (cplx) {
        re[i] = (double)src[2 * i];
        if (im) im[i] = (double)src[2 * i + 1];
    } else {
        re[i] = (double)src[i];
        if (im) im[i] = 0.0;
    }
}
template <typename T>
__global__ void pack_f64_kernel(T* __restrict__ dst, const double* __restrict__ re, const double* __restrict__ im, size_t n, int cplx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (cplx) {
        dst[2 * i] = (T)re[i];
        dst[2 * i + 1] = im ? (T)im[i] : (T)0;
    } else {
        dst[i] = (T)re[i];
    }
}

// fixed-order block reduction (sum or max) of one double per thread; result valid in thread 0
template <bool MAX>
__device__ __forceinline__ double block_reduce(double v, double* sh) {
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) sh[tid] = MAX ? fmax(sh[tid], sh[tid + s]) : sh[tid] + sh[tid + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------------
// awgn(Y, snr, 'measured'): complex white noise of power mean(|Y|^2) / 10^(snr/10), per slice
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void power_kernel(const T* __restrict__ y, int64_t nmeas, double* __restrict__ power) {
    __shared__ double sh[1024];
    const T* ys = y + 2 * (size_t)blockIdx.x * nmeas;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < nmeas; i += blockDim.x) {
        const double a = (double)ys[2 * i], b = (double)ys[2 * i + 1];
        acc += a * a + b * b;
    }
    const double tot = block_reduce<false>(acc, sh);
    if (threadIdx.x == 0) power[blockIdx.x] = nmeas > 0 ? tot / (double)nmeas : 0.0;
}

// Philox-4x32-10 (Salmon et al., SC'11): counter-based, so sample i of slice s depends only on (seed, s, i)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <typename T>
__global__ void awgn_kernel(T* __restrict__ y, int64_t nmeas, int S, const double* __restrict__ power, double inv_snr_lin, uint32_t k0,
                            uint32_t k1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= nmeas) return;
    uint32_t r[4];
    philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)s, 0u, k0, k1, r);
    const double u1 = ((double)r[0] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)r[1] + 0.5) * (1.0 / 4294967296.0);
    const double sigma = sqrt(0.5 * power[s] * inv_snr_lin);  // real and imaginary parts carry half of the noise power each
    const double rad = sqrt(-2.0 * log(u1)) * sigma;
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    T* p = y + 2 * ((size_t)s * nmeas + i);
    p[0] = (T)((double)p[0] + rad * cs);
    p[1] = (T)((double)p[1] + rad * sn);
}

// ------------------------------------------------------------------------------------------------
// P.for / P.adj
// ------------------------------------------------------------------------------------------------
// y_j = sum_c conj(V(i,c)) X_c(k_j) for sample j of frame i  (V real; identity: y_j = X_i(k_j))
__global__ void p_for_kernel(PMatrix P, const double* __restrict__ k_re, const double* __restrict__ k_im, double* __restrict__ y_re,
                             double* __restrict__ y_im) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.nmeas) return;
    int lo = 0, hi = P.L;  // frame of sample j: frame_ptr[f] <= j < frame_ptr[f + 1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)P.frame_ptr[mid] <= j) lo = mid; else hi = mid;
    }
    const int f = lo;
    const int64_t k = P.idx[j];
    double ar = 0.0, ai = 0.0;
    if (!P.V) {
        ar = k_re[(int64_t)f * P.NM + k];
        ai = k_im[(int64_t)f * P.NM + k];
    } else {
        for (int c = 0; c < P.C; ++c) {
            const double v = P.V[f + (size_t)P.L * c];
            ar = fma(v, k_re[(int64_t)c * P.NM + k], ar);
            ai = fma(v, k_im[(int64_t)c * P.NM + k], ai);
        }
    }
    y_re[j] = ar;
    y_im[j] = ai;
}
// X_c(k_j) += V(f,c) y_j over the samples [j0, j1) of ONE frame (locations are unique inside a frame: no conflicts)
__global__ void p_adj_frame_kernel(PMatrix P, int f, int j0, int j1, const double* __restrict__ y_re, const double* __restrict__ y_im,
                                   double* __restrict__ k_re, double* __restrict__ k_im) {
    const int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= j1) return;
    const int64_t k = P.idx[j];
    for (int c = 0; c < P.C; ++c) {
        const double v = P.V[f + (size_t)P.L * c];
        k_re[(int64_t)c * P.NM + k] += v * y_re[j];
        k_im[(int64_t)c * P.NM + k] += v * y_im[j];
    }
}
// identity V: frame f writes channel f only -> every (k, c) receives at most one sample
__global__ void p_adj_identity_kernel(PMatrix P, const double* __restrict__ y_re, const double* __restrict__ y_im, double* __restrict__ k_re,
                                      double* __restrict__ k_im) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.nmeas) return;
    int lo = 0, hi = P.L;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)P.frame_ptr[mid] <= j) lo = mid; else hi = mid;
    }
    const int64_t o = (int64_t)lo * P.NM + P.idx[j];
    k_re[o] = y_re[j];
    k_im[o] = y_im[j];
}

// ------------------------------------------------------------------------------------------------
// getmask_fromPD: pd = |PD| / max, pd(pd < thresh) = 0, imfill(pd, 8, 'holes'), mask(mask > 0) = 1.
// Grey-scale hole filling followed by "> 0" keeps exactly the pixels that the zero background cannot reach from the
// image border through 8-connected zero pixels: one CTA floods the background until nothing changes.
// ------------------------------------------------------------------------------------------------
__global__ void mask_kernel(const double* __restrict__ pd_abs, int N, int M, double thresh, double* __restrict__ mask,
                            unsigned char* zero, unsigned char* reach_) {
    __shared__ double sh[1024];
    volatile unsigned char* reach = reach_;
    const int n = N * M;
    double mx = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, pd_abs[i]);
    mx = block_reduce<true>(mx, sh);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = pd_abs[i] / mx;            // 0 / 0 = NaN compares false below, like MATLAB's pd(pd < thresh) = 0
        const bool z = (v < thresh) || (v == 0.0);
        const int r = i % N, c = i / N;
        zero[i] = z;
        reach[i] = z && (r == 0 || r == N - 1 || c == 0 || c == M - 1);
    }
    __syncthreads();
    for (int it = 0; it < n; ++it) {
        int changed = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (!zero[i] || reach[i]) continue;
            const int r = i % N, c = i / N;
            bool hit = false;
            for (int dc = -1; dc <= 1 && !hit; ++dc)
                for (int dr = -1; dr <= 1; ++dr) {
                    const int rr = r + dr, cc = c + dc;
                    if (rr >= 0 && rr < N && cc >= 0 && cc < M && reach[cc * N + rr]) { hit = true; break; }
                }
            if (hit) {
                reach[i] = 1;
                changed = 1;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) mask[i] = reach[i] ? 0.0 : 1.0;
}

// ------------------------------------------------------------------------------------------------
// metrics
// ------------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const double* __restrict__ v, size_t n, double* __restrict__ out) {
    __shared__ double sh[1024];
    double mx = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, fabs(v[i]));
    mx = block_reduce<true>(mx, sh);
    if (threadIdx.x == 0) *out = mx;
}
__global__ void abs_scale_kernel(const double* re, const double* im, const double* mask, const double* scale, double* out, size_t n) {  // out may alias re
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = im ? hypot(re[i], im[i]) : fabs(re[i]);
    if (mask) a *= mask[i];
    if (scale) a /= *scale;
    out[i] = a;
}

__global__ void mul_mask_kernel(double* __restrict__ v, const double* __restrict__ mask, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[(size_t)blockIdx.y * n + i] *= mask[i];
}

constexpr int SSIM_R = 5;  // radius = ceil(3 * 1.5): 11 x 11 window, sigma 1.5 (MATLAB ssim defaults)
struct GaussW { double w[2 * SSIM_R + 1]; };

// One thread per pixel: the five Gaussian-weighted moments over the replicate-padded window, the SSIM map value
// (Wang et al. eq. 13: C1 = 0.01^2, C2 = 0.03^2, dynamic range 1 for double images) and the error terms.
// partial[((pair * nb) + block) * 3 + {0,1,2}] = sum |a-b| over the mask, sum (a-b)^2, sum ssim
__global__ void pair_metrics_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ mask, int N, int M,
                                    GaussW g, double* __restrict__ partial) {
    __shared__ double sh[256];
    const int pair = blockIdx.y, n = N * M;
    const double* A = a + (size_t)pair * n;
    const double* B = b + (size_t)pair * n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e_abs = 0.0, e_sq = 0.0, ss = 0.0;
    if (i < n) {
        const int r = i % N, c = i / N;
        double mx = 0, my = 0, mxx = 0, myy = 0, mxy = 0;
        for (int dc = -SSIM_R; dc <= SSIM_R; ++dc) {
            const int cc = min(max(c + dc, 0), M - 1);
            double rx = 0, ry = 0, rxx = 0, ryy = 0, rxy = 0;
            for (int dr = -SSIM_R; dr <= SSIM_R; ++dr) {
                const int rr = min(max(r + dr, 0), N - 1);
                const double x = A[cc * N + rr], y = B[cc * N + rr], w = g.w[dr + SSIM_R];
                rx = fma(w, x, rx);
                ry = fma(w, y, ry);
                rxx = fma(w, x * x, rxx);
                ryy = fma(w, y * y, ryy);
                rxy = fma(w, x * y, rxy);
            }
            const double w = g.w[dc + SSIM_R];
            mx = fma(w, rx, mx);
            my = fma(w, ry, my);
            mxx = fma(w, rxx, mxx);
            myy = fma(w, ryy, myy);
            mxy = fma(w, rxy, mxy);
        }
        const double C1 = 1e-4, C2 = 9e-4;
        const double mux2 = mx * mx, muy2 = my * my, muxy = mx * my;
        const double sx2 = mxx - mux2, sy2 = myy - muy2, sxy = mxy - muxy;
        ss = ((2.0 * muxy + C1) * (2.0 * sxy + C2)) / ((mux2 + muy2 + C1) * (sx2 + sy2 + C2));
        const double d = A[i] - B[i];
        e_sq = d * d;
        e_abs = (!mask || mask[i] > 0.0) ? fabs(d) : 0.0;
    }
    double* out = partial + ((size_t)pair * gridDim.x + blockIdx.x) * 3;
    double t = block_reduce<false>(e_abs, sh);
    if (threadIdx.x == 0) out[0] = t;
    t = block_reduce<false>(e_sq, sh);
    if (threadIdx.x == 0) out[1] = t;
    t = block_reduce<false>(ss, sh);
    if (threadIdx.x == 0) out[2] = t;
}
// one thread per pair: block partials in block order, then MAE / PSNR / SSIM
__global__ void pair_finalize_kernel(const double* __restrict__ partial, const double* __restrict__ mask, int npairs, int nb, int n,
                                     double* __restrict__ out) {
    __shared__ double sh[1024];
    double cnt = 0.0;  // number of mask pixels (find(foreground_mask > 0))
    if (mask) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += mask[i] > 0.0 ? 1.0 : 0.0;
        cnt = block_reduce<false>(cnt, sh);
    } else {
        cnt = (double)n;
    }
    if ((int)threadIdx.x < npairs) {
        const double* p = partial + (size_t)threadIdx.x * nb * 3;
        double s_abs = 0, s_sq = 0, s_ss = 0;
        for (int k = 0; k < nb; ++k) {
            s_abs += p[3 * k];
            s_sq += p[3 * k + 1];
            s_ss += p[3 * k + 2];
        }
        out[3 * threadIdx.x] = s_abs / cnt;                          // mean(abs(a(ind) - b(ind)))
        out[3 * threadIdx.x + 1] = 10.0 * log10(1.0 / (s_sq / n));   // psnr(): peak 1 for double images; Inf when identical
        out[3 * threadIdx.x + 2] = s_ss / n;                         // mean(ssimmap(:))
    }
}

}  // namespace

int aux_unpack_f64(qmri_ctx* ctx, const void* raw, int dtype, double* re, double* im, size_t n) {
    if (!n) return QMRI_OK;
    const int cplx = dtype_is_complex(dtype);
    if (dtype == QMRI_F32 || dtype == QMRI_C64) unpack_f64_kernel<float><<<nblk(n), 256, 0, ctx->stream>>>((const float*)raw, re, im, n, cplx);
    else unpack_f64_kernel<double><<<nblk(n), 256, 0, ctx->stream>>>((const double*)raw, re, im, n, cplx);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int aux_pack_f64(qmri_ctx* ctx, void* raw, int dtype, const double* re, const double* im, size_t n) {
    if (!n) return QMRI_OK;
    const int cplx = dtype_is_complex(dtype);
    if (dtype == QMRI_F32 || dtype == QMRI_C64) pack_f64_kernel<float><<<nblk(n), 256, 0, ctx->stream>>>((float*)raw, re, im, n, cplx);
    else pack_f64_kernel<double><<<nblk(n), 256, 0, ctx->stream>>>((double*)raw, re, im, n, cplx);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

int aux_awgn(qmri_ctx* ctx, void* y_raw, int y_dtype, int64_t nmeas, int S, double snr_db, uint64_t seed, double* power) {
    if (nmeas <= 0 || S <= 0) return QMRI_OK;
    const double inv_snr = pow(10.0, -snr_db / 10.0);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    dim3 grid(nblk((size_t)nmeas), S);
    if (y_dtype == QMRI_C64) {
        power_kernel<float><<<S, 1024, 0, ctx->stream>>>((const float*)y_raw, nmeas, power);
        QLAUNCH_CHECK(ctx);
        awgn_kernel<float><<<grid, 256, 0, ctx->stream>>>((float*)y_raw, nmeas, S, power, inv_snr, k0, k1);
    } else {
        power_kernel<double><<<S, 1024, 0, ctx->stream>>>((const double*)y_raw, nmeas, power);
        QLAUNCH_CHECK(ctx);
        awgn_kernel<double><<<grid, 256, 0, ctx->stream>>>((double*)y_raw, nmeas, S, power, inv_snr, k0, k1);
    }
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

int aux_p_for(qmri_ctx* ctx, const PMatrix& P, const double* k_re, const double* k_im, double* y_re, double* y_im) {
    if (P.nmeas <= 0) return QMRI_OK;
    p_for_kernel<<<nblk((size_t)P.nmeas), 256, 0, ctx->stream>>>(P, k_re, k_im, y_re, y_im);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int aux_p_adj(qmri_ctx* ctx, const PMatrix& P, const int* frame_ptr_host, const double* y_re, const double* y_im, double* k_re, double* k_im) {
    const size_t n = (size_t)P.NM * P.C;
    QCUDA(cudaMemsetAsync(k_re, 0, n * sizeof(double), ctx->stream));
    QCUDA(cudaMemsetAsync(k_im, 0, n * sizeof(double), ctx->stream));
    if (P.nmeas <= 0) return QMRI_OK;
    if (!P.V) {
        p_adj_identity_kernel<<<nblk((size_t)P.nmeas), 256, 0, ctx->stream>>>(P, y_re, y_im, k_re, k_im);
        QLAUNCH_CHECK(ctx);
        return QMRI_OK;
    }
    for (int f = 0; f < P.L; ++f) {  // frames in order: overlapping k-space locations accumulate deterministically
        const int j0 = frame_ptr_host[f], j1 = frame_ptr_host[f + 1];
        if (j1 <= j0) continue;
        p_adj_frame_kernel<<<nblk((size_t)(j1 - j0)), 256, 0, ctx->stream>>>(P, f, j0, j1, y_re, y_im, k_re, k_im);
        QLAUNCH_CHECK(ctx);
    }
    return QMRI_OK;
}

int aux_foreground_mask(qmri_ctx* ctx, const double* pd_abs, int N, int M, double thresh, double* mask, unsigned char* scratch) {
    mask_kernel<<<1, 1024, 0, ctx->stream>>>(pd_abs, N, M, thresh, mask, scratch, scratch + (size_t)N * M);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

size_t aux_pair_metrics_partial_elems(int npairs, int N, int M) { return (size_t)npairs * nblk((size_t)N * M) * 3; }

int aux_pair_metrics(qmri_ctx* ctx, const double* a, const double* b, const double* mask, int npairs, int N, int M, double* partial,
                     double* out) {
    if (npairs < 1 || npairs > 1024) return qmri_fail(QMRI_EINVAL, "metrics: 1..1024 image pairs per call (got %d)", npairs);
    GaussW g;
    double sum = 0.0;
    for (int k = -SSIM_R; k <= SSIM_R; ++k) sum += (g.w[k + SSIM_R] = exp(-(double)(k * k) / (2.0 * 1.5 * 1.5)));
    for (int k = 0; k <= 2 * SSIM_R; ++k) g.w[k] /= sum;
    const int nb = (int)nblk((size_t)N * M);
    pair_metrics_kernel<<<dim3(nb, npairs), 256, 0, ctx->stream>>>(a, b, mask, N, M, g, partial);
    QLAUNCH_CHECK(ctx);
    pair_finalize_kernel<<<1, 1024, 0, ctx->stream>>>(partial, mask, npairs, nb, N * M, out);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

int aux_absmax(qmri_ctx* ctx, const double* v, size_t n, double* out_dev) {
    absmax_kernel<<<1, 1024, 0, ctx->stream>>>(v, n, out_dev);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int aux_abs_scale(qmri_ctx* ctx, const double* re, const double* im, const double* mask, const double* scale_dev, double* out, size_t n) {
    if (!n) return QMRI_OK;
    abs_scale_kernel<<<nblk(n), 256, 0, ctx->stream>>>(re, im, mask, scale_dev, out, n);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int aux_mul_mask(qmri_ctx* ctx, double* v, const double* mask, size_t n, int planes) {
    if (!n || planes < 1) return QMRI_OK;
    mul_mask_kernel<<<dim3(nblk(n), planes), 256, 0, ctx->stream>>>(v, mask, n);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
