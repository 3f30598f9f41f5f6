// Host-visible interface of the K2 kernels (see match_kernel.cu).
#pragma once
#include <stdint.h>

struct qmri_ctx;

struct K2Params {
    const float* x_re;   // [C][npix] planar pixel signatures
    const float* x_im;   // may be null (real data)
    int64_t npix;
    const float* Dp;     // [K][CP] atoms, rows padded with zeros to CP = 4*ceil(C/4)
    int64_t a0, a1;      // atom range scored by this launch
    unsigned long long* keys;  // [npix], zero-initialised by the caller
    int C, CP;
};

struct K2Finish {
    const float* x_re;
    const float* x_im;
    int64_t npix;
    const float* Dp;
    const float* normD;  // [K]
    const float* lut;    // [Q][K] (column-major K x Q)
    int64_t K;
    int C, CP, Q;
    const unsigned long long* keys;
    float* qmap;         // [Q][npix]
    float* pd;           // [npix][2]
    float* mt;           // [npix]
    int32_t* dm;         // [npix], 1-based
    int64_t own0, own1;  // atoms whose rows are resident here: pixels won by another rank's atom get zeros (summed over ranks later)
};

int k2_padded_channels(int C);
int k2_launch_keys(qmri_ctx* ctx, const K2Params& p);
int k2_launch_finish(qmri_ctx* ctx, const K2Finish& p);

// K4 - TSMI synthesis (main_synthesize_tsmis.m:84-98): nearest (T1, T2) atom + render
struct K4Params {
    const float* t1;     // [npix]
    const float* t2;     // [npix]
    const float* pd;     // [npix]
    int64_t npix;
    const float* lut;    // [Q][K] column-major K x Q, Q >= 2: lut[0][k] = T1, lut[1][k] = T2
    const float* Dp;     // [K][CP]
    const float* normD;  // [K]
    int64_t K;
    int C, CP;
    unsigned long long* keys;  // [npix] scratch
    float* X;            // [C][npix] planar output
    int32_t* index;      // optional [npix], 1-based nearest atom
};
int k4_launch(qmri_ctx* ctx, const K4Params& p);

// K2, tensor-pipe variant (match_tc.cu): atoms [a0, a1) packed as tf32-split K-major rows + their TMA descriptor
struct K2TcDict {
    float* Dt = nullptr;   // [ntiles * 128][32]
    int ntiles = 0;
    alignas(64) unsigned char mapB[128];
};
bool k2tc_supported(int C);
int k2tc_dict_build(qmri_ctx* ctx, const float* D, int64_t ldD, int64_t row0, int C, int64_t a0, int64_t a1, K2TcDict* out);
void k2tc_dict_free(K2TcDict* d);
size_t k2tc_stage_elems(int64_t npix);
int k2tc_launch_keys(qmri_ctx* ctx, const K2TcDict& d, const K2Params& p, float* stage);
