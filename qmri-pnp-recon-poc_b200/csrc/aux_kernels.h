// Host-visible interface of the kernels either side of the reconstruction loop (see aux_kernels.cu):
// measurement noise, the bare sampling matrix P, the foreground mask and the quality metrics of the driver script.
#pragma once
#include <stdint.h>

struct qmri_ctx;

// raw host-layout array (already on the device) <-> planar fp64 (re, im); im may be null
int aux_unpack_f64(qmri_ctx* ctx, const void* raw, int dtype, double* re, double* im, size_t n);
int aux_pack_f64(qmri_ctx* ctx, void* raw, int dtype, const double* re, const double* im, size_t n);

// awgn(Y, snr, 'measured') per slice, in place on an interleaved complex array of element type float / double
// power: S doubles of scratch (receives mean |y|^2 of every slice)
int aux_awgn(qmri_ctx* ctx, void* y_raw, int y_dtype, int64_t nmeas, int S, double snr_db, uint64_t seed, double* power);

// P.for / P.adj on planar fp64 k-space [C][NM] and measurements [nmeas]
struct PMatrix {
    const int32_t* idx;      // [nmeas] 0-based column-major k index, frame-major
    const int* frame_ptr;    // [L + 1]
    const double* V;         // [L][C] column-major L x C, or null for the identity
    int L, C;
    int64_t NM, nmeas;
};
int aux_p_for(qmri_ctx* ctx, const PMatrix& P, const double* k_re, const double* k_im, double* y_re, double* y_im);
// frame_ptr_host: the same offsets on the host (general V runs one launch per frame: frames may overlap in k-space)
int aux_p_adj(qmri_ctx* ctx, const PMatrix& P, const int* frame_ptr_host, const double* y_re, const double* y_im, double* k_re, double* k_im);

// getmask_fromPD: pd_abs [N*M] column-major (already |PD|) -> mask 0/1; scratch: N*M bytes x 2
int aux_foreground_mask(qmri_ctx* ctx, const double* pd_abs, int N, int M, double thresh, double* mask, unsigned char* scratch);

// Metrics of main_recon_tsmis_FFT.m:328-384 on `npairs` image pairs [npairs][N*M] (a = estimate, b = reference):
// out[3 * p + {0,1,2}] = masked MAE (mask may be null: all pixels), PSNR (peak 1), SSIM (MATLAB defaults)
int aux_pair_metrics(qmri_ctx* ctx, const double* a, const double* b, const double* mask, int npairs, int N, int M, double* partial,
                     double* out);
size_t aux_pair_metrics_partial_elems(int npairs, int N, int M);
// max over an image of |v| (one launch per image), result in *out_dev
int aux_absmax(qmri_ctx* ctx, const double* v, size_t n, double* out_dev);
// a[i] = |re + i im| * (mask ? mask[i] : 1) / (*scale_dev, if given)
int aux_abs_scale(qmri_ctx* ctx, const double* re, const double* im, const double* mask, const double* scale_dev, double* out, size_t n);
// v[k][i] *= mask[i] for k < planes
int aux_mul_mask(qmri_ctx* ctx, double* v, const double* mask, size_t n, int planes);
