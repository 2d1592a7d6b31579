// Host check of mesh2splat_b200/csrc/m2s_span.cuh: span_row() against the brute-force per-pixel coverage test,
// on random triangles set up exactly as raster_kernel does (24.8 fixed point, sign-normalised int64 edge
// functions, top-left ownership, candidate box).  Built and run by tests/test_span_host.py (g++).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include "../mesh2splat_b200/csrc/m2s_span.cuh"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static int rnd_range(int lo, int hi) { return lo + (int)(rnd() % (uint64_t)(hi - lo + 1)); }

int main(int argc, char** argv) {
    const long cases = argc > 1 ? atol(argv[1]) : 200000;
    long rows_checked = 0, tris = 0;
    for (long c = 0; c < cases; ++c) {
        const int R = (c % 5 == 0) ? 4096 : (c % 5 == 1 ? 64 : (c % 5 == 2 ? 512 : (c % 5 == 3 ? 2048 : 97)));
        int X[3], Y[3];
        const int mode = (int)(rnd() % 6);
        const int cx = rnd_range(-200, R * 256 + 200), cy = rnd_range(-200, R * 256 + 200);
        for (int k = 0; k < 3; ++k) {
            if (mode == 0) { X[k] = rnd_range(-8192 * 256, 8192 * 256); Y[k] = rnd_range(-8192 * 256, 8192 * 256); }       // guard band extremes
            else if (mode == 1) { X[k] = rnd_range(0, R * 256); Y[k] = rnd_range(0, R * 256); }                            // big
            else if (mode == 2) { X[k] = cx + rnd_range(-600, 600); Y[k] = cy + rnd_range(-600, 600); }                    // small
            else if (mode == 3) { X[k] = cx + rnd_range(-30000, 30000); Y[k] = cy + rnd_range(-300, 300); }                // flat sliver
            else if (mode == 4) { X[k] = (cx + rnd_range(-5000, 5000)) & ~255 | 128; Y[k] = (cy + rnd_range(-5000, 5000)) & ~255 | 128; }  // vertices ON pixel centres: ties
            else { X[k] = cx + rnd_range(-300, 300); Y[k] = cy + rnd_range(-40000, 40000); }                               // tall sliver
        }
        const long long area2 = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(X[2] - X[0]) * (Y[1] - Y[0]);
        if (area2 == 0) continue;
        const int sg = area2 < 0 ? -1 : 1;
        long long C[3]; int A[3], B[3]; unsigned incl = 0;
        for (int k = 0; k < 3; ++k) {
            const int va = (k + 1) % 3, vb = (k + 2) % 3;
            const int dx = X[vb] - X[va], dy = Y[vb] - Y[va];
            A[k] = sg * (-dy * 256); B[k] = sg * (dx * 256);
            C[k] = (long long)sg * ((long long)dx * (128 - Y[va]) - (long long)dy * (128 - X[va]));
            if (A[k] > 0 || (A[k] == 0 && B[k] > 0)) incl |= 1u << k;
        }
        const int xmin = std::min(X[0], std::min(X[1], X[2])), xmax = std::max(X[0], std::max(X[1], X[2]));
        const int ymin = std::min(Y[0], std::min(Y[1], Y[2])), ymax = std::max(Y[0], std::max(Y[1], Y[2]));
        const int x0 = std::max(0, (xmin + 127) >> 8), x1 = std::min(R - 1, (xmax - 128) >> 8);
        const int y0 = std::max(0, (ymin + 127) >> 8), y1 = std::min(R - 1, (ymax - 128) >> 8);
        if (x1 < x0 || y1 < y0) continue;
        ++tris;
        m2s::RowState s;
        for (int k = 0; k < 3; ++k) {
            s.E[k] = C[k] + (long long)A[k] * x0 + (long long)B[k] * y0 - ((incl >> k) & 1 ? 0 : 1);
            s.A[k] = A[k]; s.B[k] = B[k];
        }
        s.w = x1 - x0 + 1;
        const int h = y1 - y0 + 1;
        const int step = h > 64 ? h / 48 : 1;  // sample the rows of tall boxes
        for (int yr = 0; yr < h; yr += step) {
            int xl = -1;
            const uint32_t n = m2s::span_row(s, yr, xl);
            // brute force
            int bl = -1, bn = 0; bool contiguous = true;
            for (int xr = 0; xr < s.w; ++xr) {
                bool in = true;
                for (int k = 0; k < 3; ++k) in = in && (s.E[k] + (long long)s.A[k] * xr + (long long)s.B[k] * yr >= 0);
                if (in) { if (bl < 0) bl = xr; else if (bl + bn != xr) contiguous = false; ++bn; }
            }
            if (!contiguous) { printf("row coverage not an interval?! case %ld\n", c); return 2; }
            if ((uint32_t)bn != n || (bn && bl != xl)) {
                printf("MISMATCH case %ld R %d row %d: span (%d,%u) brute (%d,%d) w %d\n", c, R, yr, xl, n, bl, bn, s.w);
                return 1;
            }
            ++rows_checked;
        }
    }
    printf("ok %ld triangles %ld rows\n", tris, rows_checked);
    return 0;
}
