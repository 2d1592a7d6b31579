// m2s_span.cuh — exact coverage of one pixel row of a triangle's candidate box as an interval.
//
// The rasteriser's rule (GL 4.6 14.6.1, restated in oracle/m2s_oracle.c): pixel (x, y) of the box is covered iff
// E'_k(x, y) = E_k(x, y) - (edge k owns its zero set ? 0 : 1) >= 0 for the three sign-normalised integer edge
// functions E_k = A_k x + B_k y + C_k.  For a fixed row the covered set is the intersection of three half-lines,
// i.e. ONE interval [xl, xl + n): it is computed here with three estimated divisions, each corrected by exact
// int64 evaluations, so the result is bit-identical to testing every pixel centre — at O(1) per row instead
// of O(width).  Used by raster_kernel (to count) and fragment_kernel (to enumerate); also compiled for the host
// by tests/test_span_host.py (g++), which checks it against the brute-force per-pixel test.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define M2S_HD __host__ __device__ __forceinline__
#else
#define M2S_HD inline
#endif

namespace m2s {

struct RowState {   // one triangle, relative to the origin (x0, y0) of its candidate box
    long long E[3]; // E'_k at the box origin (ownership bias folded in: covered <=> all E' >= 0)
    int A[3];       // dE_k/dx per pixel
    int B[3];       // dE_k/dy per pixel
    int w;          // box width in pixels (1..4096)
};

M2S_HD float span_quot(long long num, int den) {  // estimate of num / den, num >= 0, den > 0
#if defined(__CUDA_ARCH__)
    return __fdividef(__ll2float_rn(num), __int2float_rn(den));
#else
    return (float)num / (float)den;
#endif
}

// Row `yrel` (0-based inside the box): returns the number of covered pixels, xl = first covered column.
M2S_HD uint32_t span_row(const RowState& s, int yrel, int& xl) {
    int lo = 0, hi = s.w - 1;
    const float wlim = (float)s.w + 2.0f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; ++k) {
        const long long D = s.E[k] + (long long)s.B[k] * yrel;  // E'_k at column 0 of this row
        const int A = s.A[k];
        if (A == 0) {
            if (D < 0) hi = -1;
        } else if (A > 0) {             // covered for x >= ceil(-D / A)
            if (D < 0) {
                const float qf = span_quot(-D, A);
                if (!(qf < wlim)) hi = -1;      // the boundary lies right of the box (also catches inf/NaN)
                else {
                    int x = (int)qf;
                    long long e = D + (long long)A * x;
                    while (e < 0) { ++x; e += A; }                     // exact fix-up of the estimate
                    while (x > 0 && e - A >= 0) { --x; e -= A; }
                    lo = lo > x ? lo : x;
                }
            }
        } else {                        // covered for x <= floor(D / -A)
            if (D < 0) hi = -1;
            else {
                const float qf = span_quot(D, -A);
                if (qf < wlim) {                // else: no constraint inside the box
                    int x = (int)qf + 1;
                    long long e = D + (long long)A * x;
                    while (e < 0) { --x; e -= A; }                     // A < 0: stepping left raises e
                    while (e + A >= 0) { ++x; e += A; }
                    hi = hi < x ? hi : x;
                }
            }
        }
    }
    xl = lo;
    return hi >= lo ? (uint32_t)(hi - lo + 1) : 0u;
}

}  // namespace m2s
