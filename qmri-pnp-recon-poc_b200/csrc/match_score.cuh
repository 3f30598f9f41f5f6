// |<d, x>|^2 with ONE fixed operation order, shared by the FP32-FMA kernel (main loop and rescan) and by the rescoring step of
// the tensor-core kernel: every path that decides an atom index evaluates exactly this expression, so keys produced by
// different kernels / atom ranges / ranks compare bit for bit.
#pragma once

template <int C, bool CPLX>
__device__ __forceinline__ float k2_score(const float* d, const float* xr, const float* xi) {
    float sr = 0.f, si = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        sr = fmaf(d[c], xr[c], sr);
        if (CPLX) si = fmaf(d[c], xi[c], si);
    }
    return CPLX ? fmaf(si, si, sr * sr) : sr * sr;
}

// Power of two that brings the largest component of a pixel's signature into [1, 2) (|x| >= 2^127: [2, 4)).  The argmax over atoms is
// invariant to a per-pixel scale, and with unit-norm atoms |<d, x s>|^2 <= C * 4 then stays inside the fp32 range for every finite
// pixel - unscaled, signatures below ~1e-19 squared to zero / denormals and above ~1e19 to infinity, and every atom tied (the
// reference ranks by abs(ip) and has twice the exponent range).  A power of two is exact, so the ranking of the fp32 scores is the one
// the unscaled expression would give wherever that one does not leave the range.  Every kernel that builds a key scales with this
// function of the pixel data alone, so keys of different kernels / atom ranges / ranks stay comparable; mt and pd are recomputed
// from the original data (match_finish_kernel).  Zero, denormal-only, infinite and NaN pixels are left alone.
__device__ __forceinline__ float k2_pixel_scale(float maxabs) {
    const unsigned e = (__float_as_uint(maxabs) >> 23) & 0xffu;
    if (e == 0u || e == 255u) return 1.0f;
    return __uint_as_float((e >= 253u ? 1u : 254u - e) << 23);
}
