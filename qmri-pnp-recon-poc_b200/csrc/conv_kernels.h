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

// tcgen05 implicit-GEMM convolutions on split-bf16 planes (conv_tc.cu)
enum { TC_CONV3X3 = 0, TC_DOWN2X2 = 1, TC_UP2X2 = 2 };
struct TcConvParams {
    uint16_t* out_hi;
    uint16_t* out_lo;
    const uint16_t* res1_hi;   // optional residual (may alias out)
    const uint16_t* res1_lo;
    const uint16_t* res2_hi;   // optional U-skip
    const uint16_t* res2_lo;
    int mode;                  // TC_CONV3X3 / TC_DOWN2X2 / TC_UP2X2
    int S, H, W;               // pixel grid of the GEMM's M dimension: conv3x3 image, down OUTPUT, up INPUT
    int Cin, Cout;             // channels per tap / per output pixel
    int BW, BH, tiles_x, tiles_y;
    int relu;
    float* partial;            // split-K workspace (tc_partial_elems floats) and tickets (tc_ticket_count ints, zeroed once)
    int* tickets;
    const void* mapA_hi[4];    // CUtensorMap*: [0] for conv3x3 / up, one per tap (dy*2+dx) for down
    const void* mapA_lo[4];
    const void* mapB_hi;       // weights [N][K] K-major, box = tc_block_n(Cout) rows
    const void* mapB_lo;
    const void* mapB_h2;       // pair kernel, stacked mode: w_hi with a box of Cout / 2 rows
    const void* mapB_l2;       // pair kernel, resident-weight mode: w_lo with a box of Cout / 2 rows
    const void* mapO_hi;       // pair kernel, 64 -> 64 layers: OUTPUT maps with box (BW, BH) for the bulk tensor stores of the epilogue
    const void* mapO_lo;
    const void* mapR_hi;       // ... and the same kind of maps over the res1 buffer (residual tile by bulk tensor load)
    const void* mapR_lo;
};

int conv3x3_fp32(qmri_ctx* ctx, const ConvParams& p);
int resample_fp32(qmri_ctx* ctx, const ConvParams& p);
int head_fp32(qmri_ctx* ctx, const HeadTailParams& p);
int tail_fp32(qmri_ctx* ctx, const HeadTailParams& p);
int head_pack_tc(qmri_ctx* ctx, const HeadTailParams& p);     // tensor-mode head: planar input -> 64-channel hi / lo planes (conv follows)
int tail_unpack_tc(qmri_ctx* ctx, const HeadTailParams& p);   // tensor-mode tail: channels 0 .. 9 of the conv output -> planar fp32
int tc_make_act_map(void* out_map, const void* base, int S, int H, int W, int C, int BW, int BH);
int tc_make_down_map(void* out_map, const void* base, int S, int H, int W, int C, int dy, int dx, int BW, int BH);
int tc_make_weight_map(void* out_map, const void* base, int K, int N, int BN);
int tc_tile_shape(int W, int H, int* BW, int* BH);
int tc_block_n(int Cout);
size_t tc_partial_elems(int sm_count);
size_t tc_ticket_count(int sm_count);
int conv_tc(qmri_ctx* ctx, const TcConvParams& p);
int tc_slab_tile_shape(int W, int H, int* BW, int* BH);
void tc_pair_weight_boxes(int Cout, int* rows_main, int* rows_h2);
int conv3x3_tc_pair(qmri_ctx* ctx, const TcConvParams& p);
// operand-swapped 64 -> 64 conv (weights in TMEM, 224 pixels as the MMA's N): mapA_*[0] = activation maps with box (16, 16)
bool conv64_swap_supported(int W, int H, int Cin, int Cout);
int conv64_swap(qmri_ctx* ctx, const TcConvParams& p, const uint16_t* w_hi, const uint16_t* w_lo);
int normalize_planar(qmri_ctx* ctx, const float* in, float* out, const float* minmax, size_t per_slice, int S, int undo);
