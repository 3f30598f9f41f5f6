// Internal interface of the denoiser (see unetres.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

struct qmri_ctx;

struct alignas(64) TMap {
    unsigned char bytes[128];  // CUtensorMap
};

struct qmri_net {
    qmri_ctx* ctx = nullptr;
    int in_nc = 10;
    int precision = 1;          // 0 = fp32 CUDA cores (exact mode), 1 = tcgen05 split-bf16 (default; 5e-6 rel-L2)
    bool tc_available = true;
    std::vector<float*> w[2];   // packed device weights per layer, [0] PyTorch planes, [1] MATLAB planes
    float* ws = nullptr;        // activation workspace
    size_t ws_elems = 0;
    int max_chunk = 128;        // slices evaluated per pass through the network (72 MB of workspace per slice).  Measured on B200,
                                // 120 slices: passes of 15 / 30 / 60 / 120 slices -> 80.2 / 78.6 / 77.8 / 76.9 ms per forward, bit-identical
                                // outputs (deep levels fill whole waves of CTA pairs, 64 instead of 512 launches)
    int chunk = 0;
    float* io = nullptr;        // staging for the host entry points
    size_t io_elems = 0;
    // tensor mode (tcgen05, split bf16): K-major weights [N][K] as hi / lo planes and their TMA maps
    // (3x3: [Cout][9*Cin]; down 2x2: [Cout][4*Cin]; up 2x2 transposed: [4*Cout][Cin])
    std::vector<uint16_t*> wtc_hi[2], wtc_lo[2];
    std::vector<TMap> wmap_hi[2], wmap_lo[2];
    // activation TMA maps per workspace buffer (X, A, T) x level x plane (hi, lo); rebuilt when the workspace moves
    TMap amap[3][4][2];
    TMap amap_slab[3][4][2];     // (BH + 2) x BW slabs for the CTA-pair 3x3 kernel
    TMap amap_tile[3][2];        // level 0: BH x BW tiles of the CTA-pair kernel (output side: bulk tensor stores of the 64 -> 64 layers)
    TMap amap_swap[3][2];        // level 0: (14 + 2) x 16 slabs for the operand-swapped 64 -> 64 kernel (conv64_swap)
    int tc_swap = 0;             // 1 = the 64 -> 64 convs of level 0 run through conv64_swap when the batch fills the machine (QMRI_TC_SWAP)
    std::vector<TMap> wmapp_hi[2], wmapp_lo[2], wmapp_h2[2], wmapp_l2[2];  // 3x3 weights with the box rows of the CTA-pair kernel
    int tc_pair = 15;            // 3x3 convs, bit l = level l: 1 = CTA-pair kernel (cta_group::2, slab A reuse), 0 = single-CTA per-tap kernel
    TMap amap_down[3][3][4][2];  // buffer x input level x tap (dy*2+dx) x plane: stride-2 views for the 2x2 s2 convs
    float* tc_partial = nullptr; // split-K workspace and tickets (conv_tc.cu)
    int* tc_tickets = nullptr;
    const float* amap_ws = nullptr;
    int amap_chunk = 0, amap_H = 0, amap_W = 0;
};

int unetres_create(qmri_ctx* ctx, int in_nc, const float* const* weights, int n_weights, qmri_net** out);
void unetres_free(qmri_net* net);
int unetres_reserve(qmri_net* net, int S, int H, int W);
// planes [S][C][H][W], W fastest; orient 0: PyTorch (rows = h), 1: MATLAB (rows = w, h fastest)
int unetres_forward_dev(qmri_net* net, const float* in, float* out, const float* minmax, const float* noise_map,
                        int S, int H, int W, int orient);
double unetres_flops(int in_nc, int S, int H, int W);
size_t unetres_weight_count(int in_nc, int layer);
