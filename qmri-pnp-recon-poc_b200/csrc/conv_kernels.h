// Host-visible interface of the K3 convolution kernels (conv_fp32.cu, conv_tc.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

struct qmri_ctx;

struct ConvParams {
    const float* in;    // [S][H][W][Cin]
    const float* w;     // packed weights (layout depends on the kernel, see unetres.cu)
    float* out;         // [S][Ho][Wo][Cout]
    const float* res1;  // optional residual, same shape as out (ResBlock input)
    const float* res2;  // optional second residual (U-skip)
    int S, H, W;        // input spatial size
    int Cin, Cout;
    int relu;
    int mode;           // resample kernel: 0 = 2x2 s2 conv, 1 = 2x2 s2 transposed conv
    // split-bf16 activation planes (tensor mode); when set they replace in / out
    const uint16_t* in_hi = nullptr;
    const uint16_t* in_lo = nullptr;
    uint16_t* out_hi = nullptr;
    uint16_t* out_lo = nullptr;
};

struct HeadTailParams {
    const float* planar_in;   // head: [S][Cpl][H][W]
    const float* noise_map;   // head: optional [H][W] plane appended as the last input channel
    float* planar_out;        // tail: [S][10][H][W]
    float* nhwc;              // head: output [S][H][W][64]; tail: input
    const float* w;           // head: [9][Cin][64]; tail: [9][64][10]
    const float* minmax;      // optional [S][2] per-slice min / max (affine folded in)
    int S, H, W, Cin;
    // split-bf16 planes replacing nhwc (tensor mode)
    uint16_t* nhwc_hi = nullptr;
    uint16_t* nhwc_lo = nullptr;
};

// tcgen05 implicit-GEMM 3x3 conv on split-bf16 planes (conv_tc.cu)
struct TcConvParams {
    uint16_t* out_hi;
    uint16_t* out_lo;
    const uint16_t* res1_hi;   // optional residual (may alias out)
    const uint16_t* res1_lo;
    const uint16_t* res2_hi;   // optional U-skip
    const uint16_t* res2_lo;
    int S, H, W, Cin, Cout;
    int BW, BH, tiles_x, tiles_y;
    int relu;
};

int conv3x3_fp32(qmri_ctx* ctx, const ConvParams& p);
int resample_fp32(qmri_ctx* ctx, const ConvParams& p);
int head_fp32(qmri_ctx* ctx, const HeadTailParams& p);
int tail_fp32(qmri_ctx* ctx, const HeadTailParams& p);
int tc_make_act_map(void* out_map, const void* base, int S, int H, int W, int C, int BW, int BH);
int tc_make_weight_map(void* out_map, const void* base, int K, int Cout, int BN);
int tc_tile_shape(int W, int H, int* BW, int* BH);
int tc_block_n(int Cout);
int conv3x3_tc(qmri_ctx* ctx, const void* mapA_hi, const void* mapA_lo, const void* mapB_hi, const void* mapB_lo,
               const TcConvParams& p);
int normalize_planar(qmri_ctx* ctx, const float* in, float* out, const float* minmax, size_t per_slice, int S, int undo);
