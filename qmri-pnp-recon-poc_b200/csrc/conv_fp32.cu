// K3 (exact mode) - DRUNet convolutions as FP32 CUDA-core implicit GEMMs.
//
// Reference being replaced: the layers of UNetRes
//   PyTorch_Denoiser/zhang_dpir_testing_code/network_unet.py:68-117, basicblock.py:61-98 (conv),
//   :211-223 (ResBlock), :413-419 (ConvTranspose2d 2x2 s2), :437-443 (Conv2d 2x2 s2)
// evaluated by denoiseImage_PnP_ADMM.m:88.  All layers are bias-free.
//
// Activations live in HBM as [S][Y][X][Ch] fp32 (channels innermost), X being the fastest
// spatial index of whatever plane layout the caller uses; the weight packer (unetres.cu)
// transposes the taps when the planes are MATLAB-ordered.  ReLU, the ResBlock residual and the
// U-skip adds are fused into the producing conv's epilogue; the per-slice 0-1 normalisation
// of PnP_ADMM.m:121 / :138 is folded into the head load and the tail store.
#include <math.h>
#include <string.h>

#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_kernels.h"

namespace {

// activation access that understands both storage formats: fp32, or split bf16 (hi + lo planes)
__device__ __forceinline__ float bf16_bits_to_float(uint16_t u) { return __uint_as_float((uint32_t)u << 16); }
__device__ __forceinline__ uint16_t float_to_bf16_bits(float f) {
    __nv_bfloat16 b = __float2bfloat16_rn(f);
    return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ float4 ld_act4(const float* f32, const uint16_t* hi, const uint16_t* lo, size_t idx) {
    if (!hi) return __ldg(reinterpret_cast<const float4*>(f32 + idx));
    uint2 a = __ldg(reinterpret_cast<const uint2*>(hi + idx)), b = __ldg(reinterpret_cast<const uint2*>(lo + idx));
    float4 r;
    r.x = __uint_as_float(a.x << 16) + __uint_as_float(b.x << 16);
    r.y = __uint_as_float(a.x & 0xffff0000u) + __uint_as_float(b.x & 0xffff0000u);
    r.z = __uint_as_float(a.y << 16) + __uint_as_float(b.y << 16);
    r.w = __uint_as_float(a.y & 0xffff0000u) + __uint_as_float(b.y & 0xffff0000u);
    return r;
}
__device__ __forceinline__ void st_act4(float* f32, uint16_t* hi, uint16_t* lo, size_t idx, float4 v) {
    if (!hi) {
        *reinterpret_cast<float4*>(f32 + idx) = v;
        return;
    }
    float f[4] = {v.x, v.y, v.z, v.w};
    uint16_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = float_to_bf16_bits(f[i]);
        l[i] = float_to_bf16_bits(f[i] - bf16_bits_to_float(h[i]));
    }
    *reinterpret_cast<uint2*>(hi + idx) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    *reinterpret_cast<uint2*>(lo + idx) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

// ---------------------------------------------------------------------------------------
// 3x3, stride 1, pad 1, Cin % 8 == 0, Cout % 64 == 0.  CTA: 8 x 16 output pixels x 64 couts.
// ---------------------------------------------------------------------------------------
constexpr int C3_KC = 8;
constexpr int C3_PW = 20;  // patch row stride (18 used), keeps rows 16-byte aligned

__global__ void __launch_bounds__(256) conv3x3_fp32_kernel(ConvParams p) {
    __shared__ __align__(16) float patch[C3_KC][10][C3_PW];
    __shared__ __align__(16) float wts[9][C3_KC][64];
    const int tid = threadIdx.x;
    const int tiles_x = (p.W + 15) / 16;
    const int ty0 = (blockIdx.x / tiles_x) * 8, tx0 = (blockIdx.x % tiles_x) * 16;
    const int co0 = blockIdx.y * 64;
    const int s = blockIdx.z;
    const int pg = tid >> 4, cgp = tid & 15;
    const int row = pg >> 1, xh = (pg & 1) * 8;
    const float* in = p.in + (size_t)s * p.H * p.W * p.Cin;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int ci0 = 0; ci0 < p.Cin; ci0 += C3_KC) {
        __syncthreads();
        // input patch (10 x 18 pixels x 8 channels), zero padded
        for (int t = tid; t < 360; t += 256) {
            int pix = t % 180, half = t / 180;
            int py = pix / 18, px = pix % 18;
            int y = ty0 + py - 1, x = tx0 + px - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < p.H && x >= 0 && x < p.W)
                v = __ldg(reinterpret_cast<const float4*>(in + ((size_t)y * p.W + x) * p.Cin + ci0 + 4 * half));
            patch[4 * half + 0][py][px] = v.x;
            patch[4 * half + 1][py][px] = v.y;
            patch[4 * half + 2][py][px] = v.z;
            patch[4 * half + 3][py][px] = v.w;
        }
        // weights [tap][ci0..ci0+8][co0..co0+64]
        for (int t = tid; t < 9 * C3_KC * 16; t += 256) {
            int tap = t / (C3_KC * 16), r = t % (C3_KC * 16);
            int k = r / 16, q = r % 16;
            float4 v = __ldg(reinterpret_cast<const float4*>(p.w + ((size_t)tap * p.Cin + ci0 + k) * p.Cout + co0 + 4 * q));
            *reinterpret_cast<float4*>(&wts[tap][k][4 * q]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int ry = 0; ry < 3; ++ry) {
#pragma unroll
            for (int k = 0; k < C3_KC; ++k) {
                float a[12];
                const float* pr = &patch[k][row + ry][xh];
                *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(pr);
                *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(pr + 4);
                *reinterpret_cast<float2*>(a + 8) = *reinterpret_cast<const float2*>(pr + 8);
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    float4 b = *reinterpret_cast<const float4*>(&wts[ry * 3 + sx][k][4 * cgp]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        acc[i][0] = fmaf(a[i + sx], b.x, acc[i][0]);
                        acc[i][1] = fmaf(a[i + sx], b.y, acc[i][1]);
                        acc[i][2] = fmaf(a[i + sx], b.z, acc[i][2]);
                        acc[i][3] = fmaf(a[i + sx], b.w, acc[i][3]);
                    }
                }
            }
        }
    }
    const int y = ty0 + row;
    if (y < p.H) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int x = tx0 + xh + i;
            if (x >= p.W) continue;
            size_t o = (((size_t)s * p.H + y) * p.W + x) * p.Cout + co0 + 4 * cgp;
            float4 r = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (p.res1) {
                float4 t = *reinterpret_cast<const float4*>(p.res1 + o);  // may alias out (in-place ResBlock)
                r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
            }
            if (p.res2) {
                float4 t = *reinterpret_cast<const float4*>(p.res2 + o);
                r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w;
            }
            if (p.relu) {
                r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
            }
            *reinterpret_cast<float4*>(p.out + o) = r;
        }
    }
}

// ---------------------------------------------------------------------------------------
// 2x2 stride-2 conv (mode 0) and 2x2 stride-2 transposed conv (mode 1) as per-pixel GEMMs.
//   down: M = S*Ho*Wo, K = 4*Cin (tap-major), N = Cout
//   up:   M = S*Hi*Wi, K = Cin,               N = 4*Cout (tap-major), scattered to (2y+a, 2x+b)
// 64 x 64 tile, 256 threads, 4 x 4 per thread, K chunks of 16.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resample_gemm_fp32_kernel(ConvParams p) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64];
    const int tid = threadIdx.x;
    const int up = p.mode;
    const int Hi = p.H, Wi = p.W;  // input spatial size
    const int Ho = up ? 2 * Hi : Hi / 2, Wo = up ? 2 * Wi : Wi / 2;
    const int Hm = up ? Hi : Ho, Wm = up ? Wi : Wo;  // spatial size the GEMM rows enumerate
    const int64_t Mtot = (int64_t)p.S * Hm * Wm;
    const int K = up ? p.Cin : 4 * p.Cin;
    const int N = up ? 4 * p.Cout : p.Cout;
    const int64_t m0 = (int64_t)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    const int tm = tid >> 4, tn = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // A loader: thread -> (row = tid / 4, 4 consecutive k = (tid % 4) * 4)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int64_t lm = m0 + lr;
    int ls = 0, ly = 0, lx = 0;
    if (lm < Mtot) {
        ls = (int)(lm / ((int64_t)Hm * Wm));
        int rem = (int)(lm % ((int64_t)Hm * Wm));
        ly = rem / Wm;
        lx = rem % Wm;
    }
    for (int k0 = 0; k0 < K; k0 += 16) {
        __syncthreads();
        {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lm < Mtot) {
                int k = k0 + lk;
                size_t src;
                if (up) {
                    src = (((size_t)ls * Hi + ly) * Wi + lx) * p.Cin + k;
                } else {
                    int tap = k / p.Cin, ci = k % p.Cin;
                    src = (((size_t)ls * Hi + 2 * ly + (tap >> 1)) * Wi + 2 * lx + (tap & 1)) * p.Cin + ci;
                }
                v = ld_act4(p.in, p.in_hi, p.in_lo, src);
            }
            As[lk + 0][lr] = v.x;
            As[lk + 1][lr] = v.y;
            As[lk + 2][lr] = v.z;
            As[lk + 3][lr] = v.w;
            // B: 16 x 64 floats = 256 float4
            int bk = tid >> 4, bq = tid & 15;
            float4 w = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)(k0 + bk) * N + n0 + 4 * bq));
            *reinterpret_cast<float4*>(&Bs[bk][4 * bq]) = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][4 * tm]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][4 * tn]);
            float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + 4 * tm + i;
        if (m >= Mtot) continue;
        int s = (int)(m / ((int64_t)Hm * Wm));
        int rem = (int)(m % ((int64_t)Hm * Wm));
        int y = rem / Wm, x = rem % Wm;
        int n = n0 + 4 * tn;
        size_t o;
        if (up) {
            int tap = n / p.Cout, co = n % p.Cout;
            o = (((size_t)s * Ho + 2 * y + (tap >> 1)) * Wo + 2 * x + (tap & 1)) * p.Cout + co;
        } else {
            o = (((size_t)s * Ho + y) * Wo + x) * p.Cout + n;
        }
        st_act4(p.out, p.out_hi, p.out_lo, o, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
}

// ---------------------------------------------------------------------------------------
// head: planar [S][Cin][H][W] (+ optional shared noise-map plane as last channel) -> [S][H][W][64]
// with v_in = (x - min) / range folded into the load (PnP_ADMM.m:121,177-182; no zero-range guard).
// ---------------------------------------------------------------------------------------
// Register-blocked: a thread owns 4 consecutive pixels of a row x 16 of the 64 output channels (cb = tid & 3 picks the 16), so the
// six patch values of a (row, input channel) serve 3 taps x 4 pixels and every weight load feeds 64 FMAs (a thread per pixel
// issued 5 shared-memory loads per 16 FMAs and staged its patch with one scalar load in flight: 51 us for one slice).
// CTA = 256 threads = a 32 x 8 pixel tile x 64 channels, 196 tiles per slice; the four cb lanes of a pixel group write the four
// 32-byte quarters of the same 128-byte NHWC line.
constexpr int HT_TW = 32, HT_PITCH = 36;   // tile width; patch row pitch in floats (34 used; 36 keeps rows 16-byte aligned)
constexpr int HEAD_WPITCH = 80;            // weights [9][Cin][4 cb][20]: 20-float blocks put the four cb lanes on different banks
// TH = tile height (8, or 4 for one or two slices: twice the CTAs for 148 SMs); CTA = 32 TH threads.
// (d0, d1) += a * (b0, b1) as ONE packed FP32 instruction (FFMA2, sm_100): same rounding as two fmaf, half the issue slots
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rd, {%0, %1};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rd;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}

template <int TH>
__global__ void __launch_bounds__(32 * TH, (TH == 8 ? 3 : 1)) head_fp32_kernel(HeadTailParams p) {
    constexpr int HEAD_TH = TH, HEAD_ROWS = TH + 2, NT = 32 * TH;
    extern __shared__ __align__(16) float hsm[];
    const int Cin = p.Cin;
    float* patch = hsm;                                  // [Cin][10][36]
    float* wts = hsm + Cin * HEAD_ROWS * HT_PITCH;       // [9][Cin][80]
    const int tid = threadIdx.x;
    const int tiles_x = (p.W + HT_TW - 1) / HT_TW;
    const int ty0 = (blockIdx.x / tiles_x) * HEAD_TH, tx0 = (blockIdx.x % tiles_x) * HT_TW;
    const int s = blockIdx.z;
    float mn = 0.f, inv = 1.f;
    if (p.minmax) {
        mn = p.minmax[2 * s];
        inv = 1.0f / (p.minmax[2 * s + 1] - mn);
    }
    const int Cpl = p.noise_map ? Cin - 1 : Cin;  // channels stored in the planar input
    // patch staging: a warp takes whole rows (34 pixels: lane -> x, lanes 0-1 also the last two), four rows in flight
    {
        const int lane = tid & 31, wrp = tid >> 5;
        const int nrows = Cin * HEAD_ROWS;
        auto fetch = [&](int r, int px) -> float {
            const int c = r / HEAD_ROWS, y = ty0 + r % HEAD_ROWS - 1, x = tx0 + px - 1;
            if (r >= nrows || y < 0 || y >= p.H || x < 0 || x >= p.W) return 0.f;
            if (c < Cpl) return (__ldg(p.planar_in + (((size_t)s * Cpl + c) * p.H + y) * p.W + x) - mn) * inv;
            return __ldg(p.noise_map + (size_t)y * p.W + x);
        };
        for (int r0 = wrp; r0 < nrows; r0 += 4 * TH) {
            float v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v0[u] = fetch(r0 + TH * u, lane);
                v1[u] = lane < 2 ? fetch(r0 + TH * u, 32 + lane) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + TH * u;
                if (r < nrows) {
                    patch[r * HT_PITCH + lane] = v0[u];
                    if (lane < 4) patch[r * HT_PITCH + 32 + lane] = v1[u];  // columns 34, 35 are padding (zero)
                }
            }
        }
    }
    for (int t = tid; t < 9 * Cin * 16; t += NT) {  // float4 t = row (tap, c) x 16 quads
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w) + t);
        const int r = t >> 4, q16 = t & 15;
        *reinterpret_cast<float4*>(wts + r * HEAD_WPITCH + (q16 >> 2) * 20 + (q16 & 3) * 4) = w;
    }
    __syncthreads();
    const int cbq = tid & 3, tx8 = (tid >> 2) & 7, ty = tid >> 5;
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < Cin; ++c) {
#pragma unroll
        for (int ry = 0; ry < 3; ++ry) {
            const float* ar = patch + (c * HEAD_ROWS + ty + ry) * HT_PITCH + 4 * tx8;
            const float4 a0 = *reinterpret_cast<const float4*>(ar);
            const float2 a1 = *reinterpret_cast<const float2*>(ar + 4);
            const float a[6] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
#pragma unroll
            for (int sx = 0; sx < 3; ++sx) {
                const float4* wp = reinterpret_cast<const float4*>(wts + ((ry * 3 + sx) * Cin + c) * HEAD_WPITCH + cbq * 20);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b = wp[q];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ffma2(acc[i][4 * q + 0], acc[i][4 * q + 1], a[i + sx], b.x, b.y);
                        ffma2(acc[i][4 * q + 2], acc[i][4 * q + 3], a[i + sx], b.z, b.w);
                    }
                }
            }
        }
    }
    const int y = ty0 + ty;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = tx0 + 4 * tx8 + i;
        if (y < p.H && x < p.W) {
            const size_t o = (((size_t)s * p.H + y) * p.W + x) * 64 + cbq * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                st_act4(p.nhwc, p.nhwc_hi, p.nhwc_lo, o + 4 * q, make_float4(acc[i][4 * q], acc[i][4 * q + 1], acc[i][4 * q + 2], acc[i][4 * q + 3]));
        }
    }
}

// ---------------------------------------------------------------------------------------
// tail: [S][H][W][64] -> planar [S][10][H][W], v = out * range + min folded into the store
// (PnP_ADMM.m:138,187-192).
// ---------------------------------------------------------------------------------------
// Register-blocked like the head: a thread owns 4 consecutive pixels of a row x all 10 outputs, for a quarter of the input
// channels (kq = tid & 3 takes channels 4 k + kq); the four partial sums of a pixel group meet through two shuffles (fixed order:
// deterministic), then lane kq stores pixel kq, so a warp writes 32 consecutive pixels of a plane.  CTA = 256 threads = a 32 x 8
// pixel tile (196 per slice); K = 64 is staged in four chunks of 16 channels.  Patch planes are 10 x 36 floats: 360 = 8 mod 32
// puts the four kq lanes on different banks.
template <int TH>
__global__ void __launch_bounds__(32 * TH, (TH == 8 ? 4 : 1)) tail_fp32_kernel(HeadTailParams p) {
    constexpr int TAIL_TH = TH, TAIL_ROWS = TH + 2, TAIL_PLANE = TAIL_ROWS * HT_PITCH, NT = 32 * TH;
    static_assert(TAIL_PLANE % 32 == 8 || TAIL_PLANE % 32 == 24, "the four kq lanes must land on different banks");
    __shared__ __align__(16) float patch[16 * TAIL_PLANE];   // [16][10][36]
    __shared__ __align__(16) float wts[9 * 16 * 12];         // [9][16][12]
    const int tid = threadIdx.x;
    const int tiles_x = (p.W + HT_TW - 1) / HT_TW;
    const int ty0 = (blockIdx.x / tiles_x) * TAIL_TH, tx0 = (blockIdx.x % tiles_x) * HT_TW;
    const int s = blockIdx.z;
    const size_t in_off = (size_t)s * p.H * p.W * 64;
    const int kq = tid & 3, tx8 = (tid >> 2) & 7, ty = tid >> 5;
    float acc[4][10];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 10; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
        __syncthreads();
        for (int t = tid; t < TAIL_ROWS * (HT_TW + 2) * 4; t += NT) {
            int pix = t >> 2, q = t & 3;
            int py = pix / (HT_TW + 2), px = pix % (HT_TW + 2);
            int y = ty0 + py - 1, x = tx0 + px - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < p.H && x >= 0 && x < p.W)
                v = ld_act4(p.nhwc, p.nhwc_hi, p.nhwc_lo, in_off + ((size_t)y * p.W + x) * 64 + c0 + 4 * q);
            float* d = patch + (4 * q) * TAIL_PLANE + py * HT_PITCH + px;
            d[0] = v.x;
            d[TAIL_PLANE] = v.y;
            d[2 * TAIL_PLANE] = v.z;
            d[3 * TAIL_PLANE] = v.w;
        }
        for (int t = tid; t < 9 * 16 * 12; t += NT) {
            int tap = t / 192, r = t % 192;
            int k = r / 12, co = r % 12;
            wts[t] = co < 10 ? __ldg(p.w + ((size_t)tap * 64 + c0 + k) * 10 + co) : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int k4 = 0; k4 < 4; ++k4) {
            const int ch = 4 * k4 + kq;
#pragma unroll
            for (int ry = 0; ry < 3; ++ry) {
                const float* ar = patch + ch * TAIL_PLANE + (ty + ry) * HT_PITCH + 4 * tx8;
                const float4 a0 = *reinterpret_cast<const float4*>(ar);
                const float2 a1 = *reinterpret_cast<const float2*>(ar + 4);
                const float a[6] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    const float* wp = wts + ((ry * 3 + sx) * 16 + ch) * 12;
                    const float4 b0 = *reinterpret_cast<const float4*>(wp);
                    const float4 b1 = *reinterpret_cast<const float4*>(wp + 4);
                    const float2 b2 = *reinterpret_cast<const float2*>(wp + 8);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float av = a[i + sx];
                        ffma2(acc[i][0], acc[i][1], av, b0.x, b0.y);
                        ffma2(acc[i][2], acc[i][3], av, b0.z, b0.w);
                        ffma2(acc[i][4], acc[i][5], av, b1.x, b1.y);
                        ffma2(acc[i][6], acc[i][7], av, b1.z, b1.w);
                        ffma2(acc[i][8], acc[i][9], av, b2.x, b2.y);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 1);
            acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 2);
        }
    const int y = ty0 + ty, x = tx0 + 4 * tx8 + kq;  // all four lanes hold all sums; lane kq stores pixel kq
    if (y < p.H && x < p.W) {
        float mn = 0.f, rg = 1.f;
        if (p.minmax) {
            mn = p.minmax[2 * s];
            rg = p.minmax[2 * s + 1] - mn;
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float v = kq == 0 ? acc[0][j] : kq == 1 ? acc[1][j] : kq == 2 ? acc[2][j] : acc[3][j];
            p.planar_out[(((size_t)s * 10 + j) * p.H + y) * p.W + x] = v * rg + mn;
        }
    }
}

// planar affine kernels for the callback-denoiser path
__global__ void normalize_kernel(const float* in, float* out, const float* minmax, size_t per_slice, int S, int undo) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int s = blockIdx.y;
    if (i >= per_slice || s >= S) return;
    float mn = minmax[2 * s];
    float rg = minmax[2 * s + 1] - mn;
    float v = in[(size_t)s * per_slice + i];
    out[(size_t)s * per_slice + i] = undo ? v * rg + mn : (v - mn) / rg;
}

}  // namespace

int conv3x3_fp32(qmri_ctx* ctx, const ConvParams& p) {
    if (p.Cin % C3_KC || p.Cout % 64) return qmri_fail(QMRI_EINVAL, "conv3x3_fp32: Cin %% 8 / Cout %% 64");
    dim3 grid(((p.H + 7) / 8) * ((p.W + 15) / 16), p.Cout / 64, p.S);
    conv3x3_fp32_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

int resample_fp32(qmri_ctx* ctx, const ConvParams& p) {
    const int up = p.mode;
    const int64_t M = up ? (int64_t)p.S * p.H * p.W : (int64_t)p.S * (p.H / 2) * (p.W / 2);
    const int N = up ? 4 * p.Cout : p.Cout;
    if (p.Cin % 16 || N % 64) return qmri_fail(QMRI_EINVAL, "resample_fp32: Cin %% 16 / N %% 64");
    dim3 grid((unsigned)((M + 63) / 64), N / 64);
    resample_gemm_fp32_kernel<<<grid, 256, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

// tile height: 4 rows for one or two slices (392 CTAs per slice), 8 otherwise
static inline int ht_tile_h(const HeadTailParams& p) { return p.S <= 2 ? 4 : 8; }

template <int TH>
static int head_launch(qmri_ctx* ctx, const HeadTailParams& p) {
    size_t smem = (size_t)(p.Cin * (TH + 2) * HT_PITCH + 9 * p.Cin * HEAD_WPITCH) * sizeof(float);
    static size_t configured_dev[QMRI_MAX_DEV] = {};
    size_t& configured = configured_dev[qmri_dev_slot(ctx)];
    if (smem > 48 * 1024 && smem > configured) {
        QCUDA(cudaFuncSetAttribute(head_fp32_kernel<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(((p.H + TH - 1) / TH) * ((p.W + HT_TW - 1) / HT_TW), 1, p.S);
    head_fp32_kernel<TH><<<grid, 32 * TH, smem, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int head_fp32(qmri_ctx* ctx, const HeadTailParams& p) { return ht_tile_h(p) == 4 ? head_launch<4>(ctx, p) : head_launch<8>(ctx, p); }

template <int TH>
static int tail_launch(qmri_ctx* ctx, const HeadTailParams& p) {
    dim3 grid(((p.H + TH - 1) / TH) * ((p.W + HT_TW - 1) / HT_TW), 1, p.S);
    tail_fp32_kernel<TH><<<grid, 32 * TH, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int tail_fp32(qmri_ctx* ctx, const HeadTailParams& p) { return ht_tile_h(p) == 4 ? tail_launch<4>(ctx, p) : tail_launch<8>(ctx, p); }


// ---------------------------------------------------------------------------------------
// Tensor-mode head / tail: the 10 (11) -> 64 and 64 -> 10 convs run through the 64 -> 64 tensor-core conv with zero-padded
// weights; these two streaming kernels are what is left of them.
//   head_pack:   planar fp32 [S][Cpl][H][W] (+ noise map), (v - min) / range folded in -> [S][H][W][64] hi / lo bf16, channels
//                >= Cin zero.  A warp takes 32 consecutive pixels: coalesced plane reads (one pixel per lane), the values meet
//                through shuffles, and every store instruction writes 512 contiguous bytes (lane = (pixel, 16-byte chunk)).
//   tail_unpack: [S][H][W][64] hi / lo (channels 0 .. 9 hold the conv output) -> planar fp32 [S][10][H][W], v * range + min.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ht_pack2(float a, float b) {  // two floats -> bf16 pair (round to nearest even), a in the low half
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__global__ void __launch_bounds__(256) head_pack_kernel(HeadTailParams p) {
    const int lane = threadIdx.x & 31;
    const size_t HW = (size_t)p.H * p.W;
    const size_t npx = (size_t)p.S * HW;
    const size_t px0 = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;   // first pixel of this warp
    if (px0 >= npx) return;
    const int Cpl = p.noise_map ? p.Cin - 1 : p.Cin;
    float v[11];
    {
        const size_t px = px0 + lane;
        const bool ok = px < npx;
        const size_t s = ok ? px / HW : 0, r = ok ? px - s * HW : 0;
        float mn = 0.f, inv = 1.f;
        if (p.minmax) {
            mn = p.minmax[2 * s];
            inv = 1.0f / (p.minmax[2 * s + 1] - mn);
        }
#pragma unroll
        for (int c = 0; c < 11; ++c) {
            float x = 0.f;
            if (ok && c < Cpl) x = (__ldg(p.planar_in + (s * Cpl + c) * HW + r) - mn) * inv;
            else if (ok && c == Cpl && p.noise_map) x = __ldg(p.noise_map + r);
            v[c] = x;
        }
    }
    const int j = lane & 7;   // 16-byte chunk = channels 8 j .. 8 j + 7
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int q = (lane >> 3) + 4 * k;   // pixel of the warp whose chunk j this lane stores
        float w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float a = __shfl_sync(0xffffffffu, v[c], q);
            const float b = c < 3 ? __shfl_sync(0xffffffffu, v[8 + c], q) : 0.f;
            w[c] = j == 0 ? a : (j == 1 ? b : 0.f);
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            hi[c] = ht_pack2(w[2 * c], w[2 * c + 1]);
            const float h0 = __uint_as_float(hi[c] << 16), h1 = __uint_as_float(hi[c] & 0xffff0000u);
            lo[c] = ht_pack2(w[2 * c] - h0, w[2 * c + 1] - h1);
        }
        const size_t px = px0 + q;
        if (px < npx) {
            *reinterpret_cast<uint4*>(p.nhwc_hi + px * 64 + 8 * j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(p.nhwc_lo + px * 64 + 8 * j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}
__global__ void __launch_bounds__(256) tail_unpack_kernel(HeadTailParams p) {
    const size_t HW = (size_t)p.H * p.W;
    const size_t npx = (size_t)p.S * HW;
    const size_t px = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= npx) return;
    const size_t s = px / HW, r = px - s * HW;
    float mn = 0.f, rg = 1.f;
    if (p.minmax) {
        mn = p.minmax[2 * s];
        rg = p.minmax[2 * s + 1] - mn;
    }
    const uint4 h0 = *reinterpret_cast<const uint4*>(p.nhwc_hi + px * 64), h1 = *reinterpret_cast<const uint4*>(p.nhwc_hi + px * 64 + 8);
    const uint4 l0 = *reinterpret_cast<const uint4*>(p.nhwc_lo + px * 64), l1 = *reinterpret_cast<const uint4*>(p.nhwc_lo + px * 64 + 8);
    const uint32_t hh[5] = {h0.x, h0.y, h0.z, h0.w, h1.x}, ll[5] = {l0.x, l0.y, l0.z, l0.w, l1.x};
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const float a = __uint_as_float(hh[c] << 16) + __uint_as_float(ll[c] << 16);
        const float b = __uint_as_float(hh[c] & 0xffff0000u) + __uint_as_float(ll[c] & 0xffff0000u);
        p.planar_out[(s * 10 + 2 * c) * HW + r] = a * rg + mn;
        p.planar_out[(s * 10 + 2 * c + 1) * HW + r] = b * rg + mn;
    }
}
int head_pack_tc(qmri_ctx* ctx, const HeadTailParams& p) {
    if (p.Cin > 11 || !p.nhwc_hi || !p.nhwc_lo) return qmri_fail(QMRI_EINVAL, "head_pack_tc: Cin %d / missing planes", p.Cin);
    const size_t npx = (size_t)p.S * p.H * p.W;
    head_pack_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
int tail_unpack_tc(qmri_ctx* ctx, const HeadTailParams& p) {
    if (!p.nhwc_hi || !p.nhwc_lo) return qmri_fail(QMRI_EINVAL, "tail_unpack_tc: missing planes");
    const size_t npx = (size_t)p.S * p.H * p.W;
    tail_unpack_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(p);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}

int normalize_planar(qmri_ctx* ctx, const float* in, float* out, const float* minmax, size_t per_slice, int S, int undo) {
    dim3 grid((unsigned)((per_slice + 255) / 256), S);
    normalize_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, minmax, per_slice, S, undo);
    QLAUNCH_CHECK(ctx);
    return QMRI_OK;
}
