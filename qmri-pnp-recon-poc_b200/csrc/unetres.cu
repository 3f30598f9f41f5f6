// K3 orchestration - the DRUNet denoiser (UNetRes) on device.
//
// Reference being replaced: UNetRes.forward,
//   PyTorch_Denoiser/zhang_dpir_testing_code/network_unet.py:106-117, configured as in
//   PyTorch_Denoiser/main_train.py:247 (nc=[64,128,256,512], nb=4, 'R', strideconv, convtranspose),
// called by main_files/utils/denoiseImage_PnP_ADMM.m:88 through param.net (PnP_ADMM.m:128,133).
//
// Layer order == state_dict() key order (64 bias-free weight tensors):
//   0 head | 1-8 down1 res, 9 down1 strideconv | 10-17, 18 | 19-26, 27 | 28-35 body |
//   36 up3 convT, 37-44 res | 45, 46-53 | 54, 55-62 | 63 tail
#include <string.h>

#include <vector>

#include "common.cuh"
#include "conv_kernels.h"
#include "unetres.h"

static const int NC[4] = {64, 128, 256, 512};

namespace {

struct LayerDesc {
    int kind;  // 0 conv3x3, 1 down 2x2, 2 up 2x2 (transposed), 3 head, 4 tail
    int cin, cout;
};

std::vector<LayerDesc> layer_table(int in_nc) {
    std::vector<LayerDesc> L;
    L.push_back({3, in_nc, NC[0]});
    for (int lvl = 0; lvl < 3; ++lvl) {
        for (int i = 0; i < 8; ++i) L.push_back({0, NC[lvl], NC[lvl]});
        L.push_back({1, NC[lvl], NC[lvl + 1]});
    }
    for (int i = 0; i < 8; ++i) L.push_back({0, NC[3], NC[3]});
    for (int lvl = 2; lvl >= 0; --lvl) {
        L.push_back({2, NC[lvl + 1], NC[lvl]});
        for (int i = 0; i < 8; ++i) L.push_back({0, NC[lvl], NC[lvl]});
    }
    L.push_back({4, NC[0], 10});
    return L;
}

// Pack one PyTorch weight tensor for the FP32 kernels.  `orient` = 0: planes are [h][w]
// (PyTorch), internal Y = h, X = w.  orient = 1: planes are MATLAB-ordered ([w][h], h fastest):
// internal Y = w, X = h, so every tap (r along h, s along w) lands at (ry, sx) = (s, r).
void pack_layer(const LayerDesc& d, const float* w, int orient, std::vector<float>& out) {
    const int ci_n = d.cin, co_n = d.cout;
    if (d.kind == 0 || d.kind == 3 || d.kind == 4) {  // [co][ci][3][3] -> [tap][ci][co]
        out.assign((size_t)9 * ci_n * co_n, 0.f);
        for (int co = 0; co < co_n; ++co)
            for (int ci = 0; ci < ci_n; ++ci)
                for (int r = 0; r < 3; ++r)
                    for (int s = 0; s < 3; ++s) {
                        int ry = orient ? s : r, sx = orient ? r : s;
                        out[((size_t)(ry * 3 + sx) * ci_n + ci) * co_n + co] = w[(((size_t)co * ci_n + ci) * 3 + r) * 3 + s];
                    }
    } else if (d.kind == 1) {  // Conv2d [co][ci][2][2] -> B[(tap, ci)][co]
        out.assign((size_t)4 * ci_n * co_n, 0.f);
        for (int co = 0; co < co_n; ++co)
            for (int ci = 0; ci < ci_n; ++ci)
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        int ay = orient ? b : a, bx = orient ? a : b;
                        out[((size_t)(ay * 2 + bx) * ci_n + ci) * co_n + co] = w[(((size_t)co * ci_n + ci) * 2 + a) * 2 + b];
                    }
    } else {  // ConvTranspose2d [ci][co][2][2] -> B[ci][(tap, co)]
        out.assign((size_t)4 * ci_n * co_n, 0.f);
        for (int ci = 0; ci < ci_n; ++ci)
            for (int co = 0; co < co_n; ++co)
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        int ay = orient ? b : a, bx = orient ? a : b;
                        out[((size_t)ci * 4 + (ay * 2 + bx)) * co_n + co] = w[(((size_t)ci * co_n + co) * 2 + a) * 2 + b];
                    }
    }
}

}  // namespace

size_t unetres_weight_count(int in_nc, int layer) {
    auto L = layer_table(in_nc);
    const LayerDesc& d = L[layer];
    int k = (d.kind == 1 || d.kind == 2) ? 4 : 9;
    return (size_t)k * d.cin * d.cout;
}

int unetres_create(qmri_ctx* ctx, int in_nc, const float* const* weights, int n_weights, qmri_net** out) {
    if (!ctx || !weights || !out) return qmri_fail(QMRI_EINVAL, "qmri_unetres_load: null argument");
    if (in_nc != 10 && in_nc != 11)
        return qmri_fail(QMRI_EUNSUPPORTED, "UNetRes in_nc must be 10 (single_level) or 11 (multi_level), got %d", in_nc);
    if (n_weights != 64) return qmri_fail(QMRI_EINVAL, "UNetRes state_dict has 64 weight tensors, got %d", n_weights);
    DevSetter ds(ctx->device);
    qmri_net* net = new qmri_net();
    net->ctx = ctx;
    net->in_nc = in_nc;
    auto L = layer_table(in_nc);
    std::vector<float> packed;
    for (int o = 0; o < 2; ++o) {
        net->w[o].assign(64, nullptr);
        for (int l = 0; l < 64; ++l) {
            if (!weights[l]) {
                unetres_free(net);
                return qmri_fail(QMRI_EINVAL, "qmri_unetres_load: weight tensor %d is null", l);
            }
            pack_layer(L[l], weights[l], o, packed);
            int r = dev_alloc(&net->w[o][l], packed.size());
            if (r) {
                unetres_free(net);
                return r;
            }
            cudaError_t e = cudaMemcpy(net->w[o][l], packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) {
                unetres_free(net);
                return qmri_fail(QMRI_ECUDA, "weight upload failed: %s", cudaGetErrorString(e));
            }
        }
    }
    *out = net;
    return QMRI_OK;
}

void unetres_free(qmri_net* net) {
    if (!net) return;
    DevSetter ds(net->ctx->device);
    for (int o = 0; o < 2; ++o)
        for (float* p : net->w[o])
            if (p) cudaFree(p);
    if (net->ws) cudaFree(net->ws);
    if (net->io) cudaFree(net->io);
    delete net;
}

static size_t level_elems(int lvl, int H, int W) { return (size_t)(H >> lvl) * (W >> lvl) * NC[lvl]; }

int unetres_reserve(qmri_net* net, int S, int H, int W) {
    size_t per_slice = 0;
    for (int l = 0; l < 4; ++l) per_slice += 3 * level_elems(l, H, W);
    // chunk the slice batch so the workspace stays modest and activations stay L2-friendly
    int chunk = S < net->max_chunk ? S : net->max_chunk;
    size_t need = per_slice * chunk;
    if (need > net->ws_elems) {
        DevSetter ds(net->ctx->device);
        if (net->ws) cudaFree(net->ws);
        net->ws = nullptr;
        net->ws_elems = 0;
        QCHECK(dev_alloc(&net->ws, need));
        net->ws_elems = need;
    }
    net->chunk = chunk;
    return QMRI_OK;
}

// one forward over `S` slices already resident on the device; planes are [S][C][H][W] with W fastest
static int forward_chunk(qmri_net* net, const float* in, float* out, const float* minmax, const float* noise_map,
                         int S, int H, int W, int orient) {
    qmri_ctx* ctx = net->ctx;
    float* base = net->ws;
    float *X[4], *A[4], *T[4];
    for (int l = 0; l < 4; ++l) {
        size_t n = level_elems(l, H, W) * S;
        X[l] = base;
        A[l] = base + n;
        T[l] = base + 2 * n;
        base += 3 * n;
    }
    const std::vector<float*>& w = net->w[orient];
    int li = 0;
    HeadTailParams hp = {};
    hp.planar_in = in;
    hp.noise_map = (net->in_nc == 11) ? noise_map : nullptr;
    if (net->in_nc == 11 && !noise_map) return qmri_fail(QMRI_EINVAL, "11-channel denoiser needs a noise map");
    hp.nhwc = X[0];
    hp.w = w[li++];
    hp.minmax = minmax;
    hp.S = S; hp.H = H; hp.W = W; hp.Cin = net->in_nc;
    QCHECK(head_fp32(ctx, hp));

    auto conv = [&](int lvl, const float* src, float* dst, const float* wt, int relu, const float* r1, const float* r2) {
        ConvParams p = {};
        p.in = src; p.w = wt; p.out = dst; p.res1 = r1; p.res2 = r2;
        p.S = S; p.H = H >> lvl; p.W = W >> lvl; p.Cin = NC[lvl]; p.Cout = NC[lvl]; p.relu = relu;
        return conv3x3_fp32(ctx, p);
    };
    // four ResBlocks on `src` (left intact), result in A[lvl]; `skip` is added by the last conv
    auto resblocks = [&](int lvl, const float* src, const float* skip) {
        for (int b = 0; b < 4; ++b) {
            const float* xin = (b == 0) ? src : A[lvl];
            QCHECK(conv(lvl, xin, T[lvl], w[li++], 1, nullptr, nullptr));
            QCHECK(conv(lvl, T[lvl], A[lvl], w[li++], 0, xin, (b == 3) ? skip : nullptr));
        }
        return (int)QMRI_OK;
    };
    for (int lvl = 0; lvl < 3; ++lvl) {
        QCHECK(resblocks(lvl, X[lvl], nullptr));
        ConvParams p = {};
        p.in = A[lvl]; p.w = w[li++]; p.out = X[lvl + 1];
        p.S = S; p.H = H >> lvl; p.W = W >> lvl; p.Cin = NC[lvl]; p.Cout = NC[lvl + 1]; p.mode = 0;
        QCHECK(resample_fp32(ctx, p));
    }
    QCHECK(resblocks(3, X[3], X[3]));  // body(x4) + x4 feeds up3 (network_unet.py:112)
    for (int lvl = 2; lvl >= 0; --lvl) {
        ConvParams p = {};
        p.in = A[lvl + 1]; p.w = w[li++]; p.out = T[lvl];  // convT output parked in T, first ResBlock reads it
        p.S = S; p.H = H >> (lvl + 1); p.W = W >> (lvl + 1); p.Cin = NC[lvl + 1]; p.Cout = NC[lvl]; p.mode = 1;
        QCHECK(resample_fp32(ctx, p));
        // the first ResBlock needs its input intact while T is reused: move roles - input in A, temp in T
        // (swap pointers instead of copying)
        float* tmp = A[lvl]; A[lvl] = T[lvl]; T[lvl] = tmp;
        // ResBlocks in place on A[lvl] (b == 0 reads src == A[lvl]: element-wise residual, safe in place)
        for (int b = 0; b < 4; ++b) {
            QCHECK(conv(lvl, A[lvl], T[lvl], w[li++], 1, nullptr, nullptr));
            QCHECK(conv(lvl, T[lvl], A[lvl], w[li++], 0, A[lvl], (b == 3) ? X[lvl] : nullptr));
        }
    }
    HeadTailParams tp = {};
    tp.nhwc = A[0];
    tp.planar_out = out;
    tp.w = w[li++];
    tp.minmax = minmax;
    tp.S = S; tp.H = H; tp.W = W; tp.Cin = 64;
    QCHECK(tail_fp32(ctx, tp));
    return QMRI_OK;
}

int unetres_forward_dev(qmri_net* net, const float* in, float* out, const float* minmax, const float* noise_map,
                        int S, int H, int W, int orient) {
    if (!net || !in || !out) return qmri_fail(QMRI_EINVAL, "unetres forward: null argument");
    if (S <= 0) return QMRI_OK;
    if (H % 8 || W % 8 || H < 8 || W < 8) return qmri_fail(QMRI_EINVAL, "UNetRes needs H, W multiples of 8 (got %d x %d)", H, W);
    if (net->precision != 0)
        return qmri_fail(QMRI_EUNSUPPORTED, "denoiser precision mode %d is not available in this build", net->precision);
    DevSetter ds(net->ctx->device);
    QCHECK(unetres_reserve(net, S, H, W));
    const int Cpl = (net->in_nc == 11) ? 10 : net->in_nc;
    for (int s0 = 0; s0 < S; s0 += net->chunk) {
        int n = (S - s0 < net->chunk) ? S - s0 : net->chunk;
        QCHECK(forward_chunk(net, in + (size_t)s0 * Cpl * H * W, out + (size_t)s0 * 10 * H * W,
                             minmax ? minmax + 2 * s0 : nullptr, noise_map, n, H, W, orient));
    }
    return QMRI_OK;
}

double unetres_flops(int in_nc, int S, int H, int W) {
    auto L = layer_table(in_nc);
    double f = 0;
    int lvl = 0;
    for (const auto& d : L) {
        double px;
        if (d.kind == 3 || d.kind == 4) px = (double)H * W;
        else if (d.kind == 0) {
            for (lvl = 0; lvl < 4; ++lvl)
                if (NC[lvl] == d.cin) break;
            px = (double)(H >> lvl) * (W >> lvl);
        } else if (d.kind == 1) {
            for (lvl = 0; lvl < 4; ++lvl)
                if (NC[lvl] == d.cout) break;
            px = (double)(H >> lvl) * (W >> lvl);  // output pixels, K = 4 cin
            f += 2.0 * px * 4 * d.cin * d.cout;
            continue;
        } else {
            for (lvl = 0; lvl < 4; ++lvl)
                if (NC[lvl] == d.cin) break;
            px = (double)(H >> lvl) * (W >> lvl);  // input pixels, N = 4 cout
            f += 2.0 * px * 4 * d.cin * d.cout;
            continue;
        }
        f += 2.0 * px * 9 * d.cin * d.cout;
    }
    return f * S;
}
