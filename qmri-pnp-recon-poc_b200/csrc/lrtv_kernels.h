// Host-visible interface of the LRTV kernels (lrtv_kernels.cu): total-variation prox pieces on the stacked 2N x (M L) real
// image and the vector updates of the outer FISTA loop.  All double precision, fixed-order reductions.
#pragma once
#include <stddef.h>

struct qmri_ctx;

// planar (re, im) [N][cols] column-major  <->  stacked [2N][cols]
int lrtv_stack(qmri_ctx* ctx, const double* re, const double* im, int N, int cols, double* st);
int lrtv_unstack(qmri_ctx* ctx, const double* st, int N, int cols, double* re, double* im);
// sol = b - gamma div(r, s); sums[0] = sum (b - sol)^2
int lrtv_tv_sol(qmri_ctx* ctx, const double* b, const double* r, const double* s, double gamma, int R, int Cc, double* sol, double* partial, double* sums);
// sums[1] = sum sqrt(dx^2 + dy^2) of sol; dual update (in place): r -= dx / (8 gamma), s -= dy / (8 gamma), projection onto the unit
// ball, momentum `mom` against (pold, qold), which receive the projected values
int lrtv_tv_update(qmri_ctx* ctx, const double* sol, double* r, double* s, double* pold, double* qold, double gamma, double mom, int R, int Cc,
                   double* partial, double* sums);
// sums[slot] = sum sqrt(dx^2 + dy^2) of I
int lrtv_norm_tv(qmri_ctx* ctx, const double* I, int R, int Cc, double* partial, double* sums, int slot);
size_t lrtv_partial_elems(size_t n);
// x2 = x - step * g (g single precision planes)
int lrtv_grad_step(qmri_ctx* ctx, const double* xr, const double* xi, const float* gr, const float* gi, double step, size_t n, double* x2r, double* x2i);
// sums[0] = Re <g, x2 - x>, sums[1] = |x2 - x|^2
int lrtv_backtrack_terms(qmri_ctx* ctx, const double* xr, const double* xi, const double* x2r, const double* x2i, const float* gr, const float* gi, size_t n,
                         double* partial, double* sums);
// x = x2 + beta (x2 - x2prev); x2prev = x2; also the single-precision copy of x handed to the operator
int lrtv_momentum(qmri_ctx* ctx, double* xr, double* xi, const double* x2r, const double* x2i, double* pr, double* pi, double beta, size_t n, float* fr, float* fi);
int lrtv_to_float(qmri_ctx* ctx, const double* a, const double* b, size_t n, float* fa, float* fb);
// err = Fx - y (in place on Fx); sums[slot] = |err|^2
int lrtv_residual(qmri_ctx* ctx, float2* fx, const float2* y, size_t n, double* partial, double* sums, int slot);
