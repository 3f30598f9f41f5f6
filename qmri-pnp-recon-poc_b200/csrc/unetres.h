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
    int precision = 0;          // 0 = fp32 CUDA cores, 1 = tcgen05 split-bf16
    bool tc_available = true;
    std::vector<float*> w[2];   // packed device weights per layer, [0] PyTorch planes, [1] MATLAB planes
    float* ws = nullptr;        // activation workspace
    size_t ws_elems = 0;
    int max_chunk = 8;          // slices evaluated per pass through the network
    int chunk = 0;
    float* io = nullptr;        // staging for the host entry points
    size_t io_elems = 0;
    // tensor mode (tcgen05, split bf16): K-major weights [Cout][9*Cin] as hi / lo planes and their TMA maps
    std::vector<uint16_t*> wtc_hi[2], wtc_lo[2];
    std::vector<TMap> wmap_hi[2], wmap_lo[2];
    // activation TMA maps per workspace buffer (X, A, T) x level x plane (hi, lo); rebuilt when the workspace moves
    TMap amap[3][4][2];
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
