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
#include <stdlib.h>
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

uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);  // round to nearest even
}
float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// weights for the tensor path as K-major GEMM rows, split hi / lo:
//   3x3 conv   [co][tap][ci]            (tap-major K)
//   down 2x2   [co][tap][ci]
//   up 2x2 T   [phase * Cout + co][ci]  (phase = output sub-pixel (dy, dx))
void pack_layer_tc(const LayerDesc& d, const float* w, int orient, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo) {
    const int ci_n = d.cin, co_n = d.cout;
    const int kk = (d.kind == 0) ? 3 : 2, taps = kk * kk;
    hi.assign((size_t)taps * ci_n * co_n, 0);
    lo.assign((size_t)taps * ci_n * co_n, 0);
    for (int co = 0; co < co_n; ++co)
        for (int ci = 0; ci < ci_n; ++ci)
            for (int r = 0; r < kk; ++r)
                for (int s = 0; s < kk; ++s) {
                    const int ry = orient ? s : r, sx = orient ? r : s;
                    const int tap = ry * kk + sx;
                    float v;
                    size_t o;
                    if (d.kind == 2) {  // ConvTranspose2d weight is [ci][co][kh][kw]
                        v = w[(((size_t)ci * co_n + co) * kk + r) * kk + s];
                        o = ((size_t)tap * co_n + co) * ci_n + ci;
                    } else {
                        v = w[(((size_t)co * ci_n + ci) * kk + r) * kk + s];
                        o = ((size_t)co * taps + tap) * ci_n + ci;
                    }
                    hi[o] = f2bf(v);
                    lo[o] = f2bf(v - bf2f(hi[o]));
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
    if (const char* e = getenv("QMRI_TC_PAIR")) net->tc_pair = atoi(e);  // A/B timing experiments: bit l = U-Net level l uses the pair kernel
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
        // tensor-mode copies of the 58 3x3 convs and the six 2x2 stride-2 resampling convs
        net->wtc_hi[o].assign(64, nullptr);
        net->wtc_lo[o].assign(64, nullptr);
        net->wmap_hi[o].resize(64);
        net->wmap_lo[o].resize(64);
        net->wmapp_hi[o].resize(64);
        net->wmapp_lo[o].resize(64);
        net->wmapp_h2[o].resize(64);
        net->wmapp_l2[o].resize(64);
        std::vector<uint16_t> hi, lo;
        std::vector<float> wpad;
        for (int l = 0; l < 64 && net->tc_available; ++l) {
            LayerDesc dl = L[l];
            const float* wsrc = weights[l];
            if (dl.kind > 2) {
                // head / tail in tensor mode: the same conv as a zero-padded 64 -> 64 layer ([co][ci][3][3], PyTorch order) - the
                // 64 -> 64 tensor-core kernel is faster on 6 x the flops than the fp32 CUDA-core kernels were on the real ones
                wpad.assign((size_t)64 * 64 * 9, 0.f);
                for (int co = 0; co < dl.cout; ++co)
                    for (int ci = 0; ci < dl.cin; ++ci)
                        for (int t = 0; t < 9; ++t) wpad[((size_t)co * 64 + ci) * 9 + t] = weights[l][((size_t)co * dl.cin + ci) * 9 + t];
                dl = LayerDesc{0, 64, 64};
                wsrc = wpad.data();
            }
            pack_layer_tc(dl, wsrc, o, hi, lo);
            int r = dev_alloc(&net->wtc_hi[o][l], hi.size()) | dev_alloc(&net->wtc_lo[o][l], lo.size());
            if (r) {
                unetres_free(net);
                return r;
            }
            cudaMemcpy(net->wtc_hi[o][l], hi.data(), hi.size() * 2, cudaMemcpyHostToDevice);
            cudaMemcpy(net->wtc_lo[o][l], lo.data(), lo.size() * 2, cudaMemcpyHostToDevice);
            const int BN = tc_block_n(dl.cout);
            const int K = (dl.kind == 0 ? 9 : dl.kind == 1 ? 4 : 1) * dl.cin;
            const int N = (dl.kind == 2 ? 4 : 1) * dl.cout;
            r = tc_make_weight_map(&net->wmap_hi[o][l], net->wtc_hi[o][l], K, N, BN) |
                tc_make_weight_map(&net->wmap_lo[o][l], net->wtc_lo[o][l], K, N, BN);
            if (!r && dl.kind == 0) {
                int rows_main, rows_h2;
                tc_pair_weight_boxes(dl.cout, &rows_main, &rows_h2);
                r = tc_make_weight_map(&net->wmapp_hi[o][l], net->wtc_hi[o][l], K, N, rows_main) |
                    tc_make_weight_map(&net->wmapp_lo[o][l], net->wtc_lo[o][l], K, N, rows_main) |
                    tc_make_weight_map(&net->wmapp_h2[o][l], net->wtc_hi[o][l], K, N, rows_h2) |
                    tc_make_weight_map(&net->wmapp_l2[o][l], net->wtc_lo[o][l], K, N, rows_h2);
            }
            if (r) net->tc_available = false;  // driver entry point missing: tensor mode off, exact fp32 mode still works
        }
    }
    if (net->tc_available) {
        const size_t np = tc_partial_elems(ctx->sm_count), nt = tc_ticket_count(ctx->sm_count);
        int r = dev_alloc(&net->tc_partial, np) | dev_alloc(&net->tc_tickets, nt);
        if (r) {
            unetres_free(net);
            return r;
        }
        cudaMemset(net->tc_tickets, 0, nt * sizeof(int));
    } else {
        net->precision = 0;
    }
    *out = net;
    return QMRI_OK;
}

void unetres_free(qmri_net* net) {
    if (!net) return;
    DevSetter ds(net->ctx->device);
    for (int o = 0; o < 2; ++o) {
        for (float* p : net->w[o])
            if (p) cudaFree(p);
        for (uint16_t* p : net->wtc_hi[o])
            if (p) cudaFree(p);
        for (uint16_t* p : net->wtc_lo[o])
            if (p) cudaFree(p);
    }
    if (net->ws) cudaFree(net->ws);
    if (net->io) cudaFree(net->io);
    if (net->tc_partial) cudaFree(net->tc_partial);
    if (net->tc_tickets) cudaFree(net->tc_tickets);
    delete net;
}

static size_t level_elems(int lvl, int H, int W) { return (size_t)(H >> lvl) * (W >> lvl) * NC[lvl]; }

int unetres_reserve(qmri_net* net, int S, int H, int W) {
    size_t per_slice = 0;
    for (int l = 0; l < 4; ++l) per_slice += 3 * level_elems(l, H, W);
    // chunk very large slice batches so that the workspace stays bounded (9 GB at 128 slices); balanced chunks: 200 slices run as
    // 2 x 100, not 128 + 72
    const char* env_chunk = getenv("QMRI_NET_CHUNK");  // tuning knob: most slices per pass (default unetres.h max_chunk)
    if (env_chunk && atoi(env_chunk) > 0) net->max_chunk = atoi(env_chunk);
    const int nchunks = (S + net->max_chunk - 1) / net->max_chunk;
    int chunk = nchunks > 0 ? (S + nchunks - 1) / nchunks : S;
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

// ---- tensor mode ---------------------------------------------------------------------------
// Workspace buffers are laid out for the FULL chunk (net->chunk slices): buffer b of level l holds
// level_elems * chunk elements; as split bf16 that is a hi plane followed by a lo plane.
static void tc_buffers(qmri_net* net, int H, int W, uint16_t* hi[3][4], uint16_t* lo[3][4]) {
    float* base = net->ws;
    for (int l = 0; l < 4; ++l) {
        size_t n = level_elems(l, H, W) * net->chunk;
        for (int b = 0; b < 3; ++b) {
            hi[b][l] = reinterpret_cast<uint16_t*>(base + (size_t)b * n);
            lo[b][l] = hi[b][l] + n;
        }
        base += 3 * n;
    }
}

static int build_act_maps(qmri_net* net, int H, int W) {
    if (net->amap_ws == net->ws && net->amap_chunk == net->chunk && net->amap_H == H && net->amap_W == W) return QMRI_OK;
    uint16_t *hi[3][4], *lo[3][4];
    tc_buffers(net, H, W, hi, lo);
    for (int l = 0; l < 4; ++l) {
        int BW, BH;
        tc_tile_shape(W >> l, H >> l, &BW, &BH);
        for (int b = 0; b < 3; ++b) {
            QCHECK(tc_make_act_map(&net->amap[b][l][0], hi[b][l], net->chunk, H >> l, W >> l, NC[l], BW, BH));
            QCHECK(tc_make_act_map(&net->amap[b][l][1], lo[b][l], net->chunk, H >> l, W >> l, NC[l], BW, BH));
            int SW, SH;
            tc_slab_tile_shape(W >> l, H >> l, &SW, &SH);
            QCHECK(tc_make_act_map(&net->amap_slab[b][l][0], hi[b][l], net->chunk, H >> l, W >> l, NC[l], SW, SH + 2));
            QCHECK(tc_make_act_map(&net->amap_slab[b][l][1], lo[b][l], net->chunk, H >> l, W >> l, NC[l], SW, SH + 2));
            if (l == 0) {
                QCHECK(tc_make_act_map(&net->amap_tile[b][0], hi[b][0], net->chunk, H, W, NC[0], SW, SH));
                QCHECK(tc_make_act_map(&net->amap_tile[b][1], lo[b][0], net->chunk, H, W, NC[0], SW, SH));
            }
            if (l == 0 && conv64_swap_supported(W, H, NC[0], NC[0])) {
                QCHECK(tc_make_act_map(&net->amap_swap[b][0], hi[b][0], net->chunk, H, W, NC[0], 16, 16));
                QCHECK(tc_make_act_map(&net->amap_swap[b][1], lo[b][0], net->chunk, H, W, NC[0], 16, 16));
            }
            if (l < 3) {  // stride-2 tap views feeding the down conv to level l + 1
                int DW, DH;
                tc_tile_shape(W >> (l + 1), H >> (l + 1), &DW, &DH);
                for (int tap = 0; tap < 4; ++tap) {
                    QCHECK(tc_make_down_map(&net->amap_down[b][l][tap][0], hi[b][l], net->chunk, H >> l, W >> l, NC[l], tap >> 1, tap & 1, DW, DH));
                    QCHECK(tc_make_down_map(&net->amap_down[b][l][tap][1], lo[b][l], net->chunk, H >> l, W >> l, NC[l], tap >> 1, tap & 1, DW, DH));
                }
            }
        }
    }
    net->amap_ws = net->ws;
    net->amap_chunk = net->chunk;
    net->amap_H = H;
    net->amap_W = W;
    return QMRI_OK;
}

struct LayerProf {
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    bool on = false;
    void mark(qmri_ctx* ctx, const char* name) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, ctx->stream);
        ev.push_back(e);
        names.push_back(name);
    }
    void report(int S) {
        if (!on || ev.empty()) return;
        cudaEventSynchronize(ev.back());
        double tot = 0;
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            fprintf(stderr, "[qmri profile] S=%d %-10s %8.1f us\n", S, names[i], ms * 1e3);
            tot += ms;
        }
        fprintf(stderr, "[qmri profile] S=%d total %8.1f us\n", S, tot * 1e3);
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        names.clear();
    }
};

static int forward_chunk_tc(qmri_net* net, const float* in, float* out, const float* minmax, const float* noise_map,
                            int S, int H, int W, int orient) {
    qmri_ctx* ctx = net->ctx;
    LayerProf prof;
    prof.on = getenv("QMRI_PROFILE") != nullptr;
    prof.mark(ctx, "start");
    uint16_t *hi[3][4], *lo[3][4];
    tc_buffers(net, H, W, hi, lo);
    enum { BX = 0, BA = 1, BT = 2 };
    int roleA[4] = {BA, BA, BA, BA}, roleT[4] = {BT, BT, BT, BT};  // which physical buffer plays A / T per level
    const std::vector<float*>& w = net->w[orient];
    int li = 0;
    if (net->in_nc == 11 && !noise_map) return qmri_fail(QMRI_EINVAL, "11-channel denoiser needs a noise map");
    HeadTailParams hp = {};
    hp.planar_in = in;
    hp.noise_map = (net->in_nc == 11) ? noise_map : nullptr;
    hp.nhwc_hi = hi[BX][0];
    hp.nhwc_lo = lo[BX][0];
    hp.w = w[li++];
    hp.minmax = minmax;
    hp.S = S; hp.H = H; hp.W = W; hp.Cin = net->in_nc;
    // QMRI_TC_HEADTAIL=0: the fp32 CUDA-core head / tail kernels of round 1 (A/B runs)
    const char* ht_env = getenv("QMRI_TC_HEADTAIL");
    const bool ht_tc = !(ht_env && atoi(ht_env) == 0);
    if (!ht_tc) {
        QCHECK(head_fp32(ctx, hp));
        prof.mark(ctx, "head");
    }

    auto tiles = [&](TcConvParams& p) {
        tc_tile_shape(p.W, p.H, &p.BW, &p.BH);
        p.tiles_x = (p.W + p.BW - 1) / p.BW;
        p.tiles_y = (p.H + p.BH - 1) / p.BH;
        p.partial = net->tc_partial;
        p.tickets = net->tc_tickets;
    };
    // 3x3 conv: src buffer -> dst buffer at `lvl`, optional residual buffers (physical ids, -1 = none)
    auto conv = [&](int lvl, int src, int dst, int layer, int relu, int r1, int r2) {
        TcConvParams p = {};
        p.mode = TC_CONV3X3;
        p.out_hi = hi[dst][lvl]; p.out_lo = lo[dst][lvl];
        p.res1_hi = r1 >= 0 ? hi[r1][lvl] : nullptr; p.res1_lo = r1 >= 0 ? lo[r1][lvl] : nullptr;
        p.res2_hi = r2 >= 0 ? hi[r2][lvl] : nullptr; p.res2_lo = r2 >= 0 ? lo[r2][lvl] : nullptr;
        p.S = S; p.H = H >> lvl; p.W = W >> lvl; p.Cin = NC[lvl]; p.Cout = NC[lvl];
        p.relu = relu;
        tiles(p);
        // CTA-pair kernel when the layer fills the machine with pair tiles; the single-CTA kernel (finer tiles, split-K)
        // for the small layers of small slice batches, which are latency- rather than throughput-bound
        {
            // operand-swapped kernel for the 64 -> 64 layers of a batch that fills the machine (one 224-pixel tile per CTA at least)
            const char* sw_env = getenv("QMRI_TC_SWAP");  // read per call: tests toggle it.  Off by default: measured on par with the pair
            const bool sw_on = sw_env ? atoi(sw_env) != 0 : net->tc_swap != 0;  // kernel (profiles/r02_conv64_experiments.md)
            if (lvl == 0 && sw_on && conv64_swap_supported(p.W, p.H, p.Cin, p.Cout) && (p.W / 16) * (p.H / 14) * S >= ctx->sm_count) {
                p.mapA_hi[0] = &net->amap_swap[src][0]; p.mapA_lo[0] = &net->amap_swap[src][1];
                return conv64_swap(ctx, p, net->wtc_hi[orient][layer], net->wtc_lo[orient][layer]);
            }
        }
        int SW, SH;
        tc_slab_tile_shape(p.W, p.H, &SW, &SH);
        const int pair_groups = ((((p.W + SW - 1) / SW) * ((p.H + SH - 1) / SH) * S + 1) / 2) * (NC[lvl] >= 256 ? NC[lvl] / 256 : 1);
        if (((net->tc_pair >> lvl) & 1) && (pair_groups >= ctx->sm_count / 2 || (net->tc_pair & 16))) {
            p.BW = SW;
            p.BH = SH;
            p.tiles_x = (p.W + p.BW - 1) / p.BW;
            p.tiles_y = (p.H + p.BH - 1) / p.BH;
            p.mapA_hi[0] = &net->amap_slab[src][lvl][0]; p.mapA_lo[0] = &net->amap_slab[src][lvl][1];
            p.mapB_hi = &net->wmapp_hi[orient][layer]; p.mapB_lo = &net->wmapp_lo[orient][layer];
            p.mapB_h2 = &net->wmapp_h2[orient][layer];
            p.mapB_l2 = &net->wmapp_l2[orient][layer];
            if (lvl == 0) {
                p.mapO_hi = &net->amap_tile[dst][0]; p.mapO_lo = &net->amap_tile[dst][1];
                if (r1 >= 0) {
                    p.mapR_hi = &net->amap_tile[r1][0]; p.mapR_lo = &net->amap_tile[r1][1];
                }
            }
            return conv3x3_tc_pair(ctx, p);
        }
        p.mapA_hi[0] = &net->amap[src][lvl][0]; p.mapA_lo[0] = &net->amap[src][lvl][1];
        p.mapB_hi = &net->wmap_hi[orient][layer]; p.mapB_lo = &net->wmap_lo[orient][layer];
        return conv_tc(ctx, p);
    };
    // 2x2 stride-2 conv (up = 0: lvl_in -> lvl_in + 1) or transposed conv (up = 1: lvl_in -> lvl_in - 1)
    auto resample = [&](int lvl_in, int src, int lvl_out, int dst, int layer, int up) {
        TcConvParams p = {};
        p.mode = up ? TC_UP2X2 : TC_DOWN2X2;
        p.out_hi = hi[dst][lvl_out]; p.out_lo = lo[dst][lvl_out];
        const int lm = up ? lvl_in : lvl_out;  // level of the GEMM's pixel grid
        p.S = S; p.H = H >> lm; p.W = W >> lm; p.Cin = NC[lvl_in]; p.Cout = NC[lvl_out];
        tiles(p);
        if (up) {
            p.mapA_hi[0] = &net->amap[src][lvl_in][0]; p.mapA_lo[0] = &net->amap[src][lvl_in][1];
        } else {
            for (int tap = 0; tap < 4; ++tap) {
                p.mapA_hi[tap] = &net->amap_down[src][lvl_in][tap][0];
                p.mapA_lo[tap] = &net->amap_down[src][lvl_in][tap][1];
            }
        }
        p.mapB_hi = &net->wmap_hi[orient][layer]; p.mapB_lo = &net->wmap_lo[orient][layer];
        return conv_tc(ctx, p);
    };
    if (ht_tc) {  // head = pack to 64 channels (zeros above Cin) + the 64 -> 64 tensor conv with zero-padded weights
        hp.nhwc_hi = hi[BT][0];
        hp.nhwc_lo = lo[BT][0];
        QCHECK(head_pack_tc(ctx, hp));
        QCHECK(conv(0, BT, BX, 0, 0, -1, -1));
        prof.mark(ctx, "head");
    }
    for (int lvl = 0; lvl < 3; ++lvl) {
        for (int b = 0; b < 4; ++b) {
            int xin = (b == 0) ? BX : roleA[lvl];
            QCHECK(conv(lvl, xin, roleT[lvl], li, 1, -1, -1)); ++li;
            prof.mark(ctx, lvl == 0 ? "conv L0" : lvl == 1 ? "conv L1" : "conv L2");
            QCHECK(conv(lvl, roleT[lvl], roleA[lvl], li, 0, xin, -1)); ++li;
            prof.mark(ctx, lvl == 0 ? "conv L0 r" : lvl == 1 ? "conv L1 r" : "conv L2 r");
        }
        QCHECK(resample(lvl, roleA[lvl], lvl + 1, BX, li, 0)); ++li;
        prof.mark(ctx, "down");
    }
    for (int b = 0; b < 4; ++b) {
        int xin = (b == 0) ? BX : roleA[3];
        QCHECK(conv(3, xin, roleT[3], li, 1, -1, -1)); ++li;
        prof.mark(ctx, "conv L3");
        QCHECK(conv(3, roleT[3], roleA[3], li, 0, xin, (b == 3) ? BX : -1)); ++li;
        prof.mark(ctx, "conv L3 r");
    }
    for (int lvl = 2; lvl >= 0; --lvl) {
        QCHECK(resample(lvl + 1, roleA[lvl + 1], lvl, roleT[lvl], li, 1)); ++li;
        prof.mark(ctx, "up");
        int tmp = roleA[lvl]; roleA[lvl] = roleT[lvl]; roleT[lvl] = tmp;
        for (int b = 0; b < 4; ++b) {
            QCHECK(conv(lvl, roleA[lvl], roleT[lvl], li, 1, -1, -1)); ++li;
            prof.mark(ctx, lvl == 0 ? "conv L0" : lvl == 1 ? "conv L1" : "conv L2");
            QCHECK(conv(lvl, roleT[lvl], roleA[lvl], li, 0, roleA[lvl], (b == 3) ? BX : -1)); ++li;
            prof.mark(ctx, lvl == 0 ? "conv L0 r" : lvl == 1 ? "conv L1 r" : "conv L2 r");
        }
    }
    HeadTailParams tp = {};
    tp.nhwc_hi = hi[roleA[0]][0];
    tp.nhwc_lo = lo[roleA[0]][0];
    tp.planar_out = out;
    tp.w = w[li++];
    tp.minmax = minmax;
    tp.S = S; tp.H = H; tp.W = W; tp.Cin = 64;
    if (ht_tc) {  // tail = the 64 -> 64 tensor conv (output channels >= 10 are zero weights) + unpack of channels 0 .. 9
        QCHECK(conv(0, roleA[0], roleT[0], li - 1, 0, -1, -1));
        tp.nhwc_hi = hi[roleT[0]][0];
        tp.nhwc_lo = lo[roleT[0]][0];
        QCHECK(tail_unpack_tc(ctx, tp));
    } else {
        QCHECK(tail_fp32(ctx, tp));
    }
    prof.mark(ctx, "tail");
    prof.report(S);
    return QMRI_OK;
}

int unetres_forward_dev(qmri_net* net, const float* in, float* out, const float* minmax, const float* noise_map,
                        int S, int H, int W, int orient) {
    if (!net || !in || !out) return qmri_fail(QMRI_EINVAL, "unetres forward: null argument");
    if (S <= 0) return QMRI_OK;
    if (H % 8 || W % 8 || H < 8 || W < 8) return qmri_fail(QMRI_EINVAL, "UNetRes needs H, W multiples of 8 (got %d x %d)", H, W);
    if (net->precision == 1 && !net->tc_available)
        return qmri_fail(QMRI_EUNSUPPORTED, "tensor mode unavailable: cuTensorMapEncodeTiled could not be resolved");
    DevSetter ds(net->ctx->device);
    QCHECK(unetres_reserve(net, S, H, W));
    if (net->precision == 1) QCHECK(build_act_maps(net, H, W));
    const int Cpl = (net->in_nc == 11) ? 10 : net->in_nc;
    for (int s0 = 0; s0 < S; s0 += net->chunk) {
        int n = (S - s0 < net->chunk) ? S - s0 : net->chunk;
        if (net->precision == 1)
            QCHECK(forward_chunk_tc(net, in + (size_t)s0 * Cpl * H * W, out + (size_t)s0 * 10 * H * W,
                                    minmax ? minmax + 2 * s0 : nullptr, noise_map, n, H, W, orient));
        else
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
