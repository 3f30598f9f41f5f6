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
