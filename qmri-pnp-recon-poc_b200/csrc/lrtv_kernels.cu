// LRTV comparison baseline - device pieces.
//
// Reference being replaced: main_files/algorithms/LRTV/FISTA_deep.m:53-102 (outer FISTA with backtracking) and the TV prox it
// calls, unlocbox/prox/prox_tv.m:156-193 (fast gradient projection on the dual, Beck & Teboulle 2009) with the forward
// differences / divergence / isotropic TV norm of unlocbox/utils.  The TV image is the reference's stacked real array
// [reshape(real(x),N,[]); reshape(imag(x),N,[])] of 2N x (M L) pixels, column-major.  Double precision throughout (the
// reference runs in double and both loops stop on relative objective changes); reductions go through per-block partial sums
// added in block order, so runs are reproducible.  Not a hot path: LRTV is the quality baseline the driver script compares with.
#include <math.h>

#include "common.cuh"
#include "lrtv_kernels.h"

namespace {

constexpr int LT = 256;
inline unsigned nblk(size_t n) { return (unsigned)((n + LT - 1) / LT); }

__device__ __forceinline__ void block_sum_store(double v, double* partial) {
    __shared__ double sh[LT];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = LT >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void reduce_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sh[LT];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += LT) acc += partial[i];  // fixed assignment of blocks to threads
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = LT >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

__global__ void stack_kernel(const double* __restrict__ re, const double* __restrict__ im, int N, size_t n2, double* __restrict__ st) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;  // index into the stacked array
    if (e >= n2) return;
    const size_t j = e / (2 * N);
    const int i = (int)(e - j * 2 * N);
    st[e] = i < N ? re[i + (size_t)N * j] : im[i - N + (size_t)N * j];
}
__global__ void unstack_kernel(const double* __restrict__ st, int N, size_t n2, double* __restrict__ re, double* __restrict__ im) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    if (e >= n2) return;
    const size_t j = e / (2 * N);
    const int i = (int)(e - j * 2 * N);
    if (i < N) re[i + (size_t)N * j] = st[e];
    else im[i - N + (size_t)N * j] = st[e];
}

// div_op: adjoint of the forward differences with the toolbox's boundary rows / columns
__global__ void tv_sol_kernel(const double* __restrict__ b, const double* __restrict__ r, const double* __restrict__ s, double gamma, int R, int Cc,
                              double* __restrict__ sol, double* __restrict__ partial) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    double d2 = 0.0;
    if (e < (size_t)R * Cc) {
        const int j = (int)(e / R), i = (int)(e - (size_t)j * R);
        double dv;
        if (i == 0) dv = r[e];
        else if (i == R - 1) dv = -r[e - 1];
        else dv = r[e] - r[e - 1];
        if (j == 0) dv += s[e];
        else if (j == Cc - 1) dv += -s[e - R];
        else dv += s[e] - s[e - R];
        const double v = b[e] - gamma * dv;
        sol[e] = v;
        d2 = (b[e] - v) * (b[e] - v);
    }
    block_sum_store(d2, partial);
}
__global__ void tv_update_kernel(const double* __restrict__ sol, double* __restrict__ r, double* __restrict__ s, double* __restrict__ pold,
                                 double* __restrict__ qold, double inv8g, double mom, int R, int Cc, double* __restrict__ partial) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    double tv = 0.0;
    if (e < (size_t)R * Cc) {
        const int j = (int)(e / R), i = (int)(e - (size_t)j * R);
        const double c = sol[e];
        const double dx = (i < R - 1) ? sol[e + 1] - c : 0.0;
        const double dy = (j < Cc - 1) ? sol[e + R] - c : 0.0;
        tv = sqrt(dx * dx + dy * dy);
        const double rr = r[e] - inv8g * dx, ss = s[e] - inv8g * dy;
        const double w = fmax(1.0, sqrt(rr * rr + ss * ss));
        const double p = rr / w, q = ss / w;
        r[e] = p + mom * (p - pold[e]);
        s[e] = q + mom * (q - qold[e]);
        pold[e] = p;
        qold[e] = q;
    }
    block_sum_store(tv, partial);
}
__global__ void norm_tv_kernel(const double* __restrict__ I, int R, int Cc, double* __restrict__ partial) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    double tv = 0.0;
    if (e < (size_t)R * Cc) {
        const int j = (int)(e / R), i = (int)(e - (size_t)j * R);
        const double c = I[e];
        const double dx = (i < R - 1) ? I[e + 1] - c : 0.0;
        const double dy = (j < Cc - 1) ? I[e + R] - c : 0.0;
        tv = sqrt(dx * dx + dy * dy);
    }
    block_sum_store(tv, partial);
}
__global__ void grad_step_kernel(const double* __restrict__ xr, const double* __restrict__ xi, const float* __restrict__ gr, const float* __restrict__ gi,
                                 double step, size_t n, double* __restrict__ x2r, double* __restrict__ x2i) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    if (e >= n) return;
    x2r[e] = xr[e] - step * (double)gr[e];
    x2i[e] = xi[e] - step * (double)gi[e];
}
__global__ void backtrack_kernel(const double* __restrict__ xr, const double* __restrict__ xi, const double* __restrict__ x2r, const double* __restrict__ x2i,
                                 const float* __restrict__ gr, const float* __restrict__ gi, size_t n, double* __restrict__ p_dot, double* __restrict__ p_nrm) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    double dot = 0.0, nrm = 0.0;
    if (e < n) {
        const double dr = x2r[e] - xr[e], di = x2i[e] - xi[e];
        dot = (double)gr[e] * dr + (double)gi[e] * di;   // real(grad1(:)' * (x2(:) - x(:)))
        nrm = dr * dr + di * di;
    }
    block_sum_store(dot, p_dot);
    __syncthreads();
    block_sum_store(nrm, p_nrm);
}
__global__ void momentum_kernel(double* __restrict__ xr, double* __restrict__ xi, const double* __restrict__ x2r, const double* __restrict__ x2i,
                                double* __restrict__ pr, double* __restrict__ pi, double beta, size_t n, float* __restrict__ fr, float* __restrict__ fi) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    if (e >= n) return;
    const double a = x2r[e], b = x2i[e];
    const double nr = a + beta * (a - pr[e]), ni = b + beta * (b - pi[e]);
    xr[e] = nr; xi[e] = ni;
    pr[e] = a; pi[e] = b;
    fr[e] = (float)nr; fi[e] = (float)ni;
}
__global__ void to_float_kernel(const double* __restrict__ a, const double* __restrict__ b, size_t n, float* __restrict__ fa, float* __restrict__ fb) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    if (e >= n) return;
    fa[e] = (float)a[e];
    fb[e] = (float)b[e];
}
__global__ void residual_kernel(float2* __restrict__ fx, const float2* __restrict__ y, size_t n, double* __restrict__ partial) {
    const size_t e = (size_t)blockIdx.x * LT + threadIdx.x;
    double v = 0.0;
    if (e < n) {
        const float2 d = make_float2(fx[e].x - y[e].x, fx[e].y - y[e].y);
        fx[e] = d;
        v = (double)d.x * d.x + (double)d.y * d.y;
    }
    block_sum_store(v, partial);
}

int reduce_to(qmri_ctx* ctx, const double* partial, size_t n, double* out) {
    reduce_kernel<<<1, LT, 0, ctx->stream>>>(partial, (int)nblk(n), out);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

}  // namespace

size_t lrtv_partial_elems(size_t n) { return 2 * (size_t)nblk(n); }

int lrtv_stack(qmri_ctx* ctx, const double* re, const double* im, int N, int cols, double* st) {
    const size_t n2 = (size_t)2 * N * cols;
    stack_kernel<<<nblk(n2), LT, 0, ctx->stream>>>(re, im, N, n2, st);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int lrtv_unstack(qmri_ctx* ctx, const double* st, int N, int cols, double* re, double* im) {
    const size_t n2 = (size_t)2 * N * cols;
    unstack_kernel<<<nblk(n2), LT, 0, ctx->stream>>>(st, N, n2, re, im);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int lrtv_tv_sol(qmri_ctx* ctx, const double* b, const double* r, const double* s, double gamma, int R, int Cc, double* sol, double* partial, double* sums) {
    const size_t n = (size_t)R * Cc;
    tv_sol_kernel<<<nblk(n), LT, 0, ctx->stream>>>(b, r, s, gamma, R, Cc, sol, partial);
    QLAUNCH_CHECK(ctx);
    return reduce_to(ctx, partial, n, sums);
}
int lrtv_tv_update(qmri_ctx* ctx, const double* sol, double* r, double* s, double* pold, double* qold, double gamma, double mom, int R, int Cc,
                   double* partial, double* sums) {
    const size_t n = (size_t)R * Cc;
    tv_update_kernel<<<nblk(n), LT, 0, ctx->stream>>>(sol, r, s, pold, qold, 1.0 / (8.0 * gamma), mom, R, Cc, partial);
    QLAUNCH_CHECK(ctx);
    return reduce_to(ctx, partial, n, sums + 1);
}
int lrtv_norm_tv(qmri_ctx* ctx, const double* I, int R, int Cc, double* partial, double* sums, int slot) {
    const size_t n = (size_t)R * Cc;
    norm_tv_kernel<<<nblk(n), LT, 0, ctx->stream>>>(I, R, Cc, partial);
    QLAUNCH_CHECK(ctx);
    return reduce_to(ctx, partial, n, sums + slot);
}
int lrtv_grad_step(qmri_ctx* ctx, const double* xr, const double* xi, const float* gr, const float* gi, double step, size_t n, double* x2r, double* x2i) {
    grad_step_kernel<<<nblk(n), LT, 0, ctx->stream>>>(xr, xi, gr, gi, step, n, x2r, x2i);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int lrtv_backtrack_terms(qmri_ctx* ctx, const double* xr, const double* xi, const double* x2r, const double* x2i, const float* gr, const float* gi, size_t n,
                         double* partial, double* sums) {
    double* p2 = partial + nblk(n);
    backtrack_kernel<<<nblk(n), LT, 0, ctx->stream>>>(xr, xi, x2r, x2i, gr, gi, n, partial, p2);
    QLAUNCH_CHECK(ctx);
    QCHECK(reduce_to(ctx, partial, n, sums));
    return reduce_to(ctx, p2, n, sums + 1);
}
int lrtv_momentum(qmri_ctx* ctx, double* xr, double* xi, const double* x2r, const double* x2i, double* pr, double* pi, double beta, size_t n, float* fr, float* fi) {
    momentum_kernel<<<nblk(n), LT, 0, ctx->stream>>>(xr, xi, x2r, x2i, pr, pi, beta, n, fr, fi);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int lrtv_to_float(qmri_ctx* ctx, const double* a, const double* b, size_t n, float* fa, float* fb) {
    to_float_kernel<<<nblk(n), LT, 0, ctx->stream>>>(a, b, n, fa, fb);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int lrtv_residual(qmri_ctx* ctx, float2* fx, const float2* y, size_t n, double* partial, double* sums, int slot) {
    residual_kernel<<<nblk(n), LT, 0, ctx->stream>>>(fx, y, n, partial);
    QLAUNCH_CHECK(ctx);
    return reduce_to(ctx, partial, n, sums + slot);
}
