// Probe (measurement aid, not product code): semantics and throughput of tex2Dgather on sm_100a.
//  1. which texel lands in which component of the float4, with WRAP addressing at footprint-centre coordinates
//  2. is the unorm8 -> float conversion exactly b/255 (rounded) ?
//  3. throughput: 8 gathers per thread (4 channels x 2 levels) vs 8 scalar 4-byte loads, coherent access pattern
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Tex { cudaArray_t arr; cudaTextureObject_t obj; int W, H; std::vector<unsigned char> host; uint32_t* lin; };

static int make_tex(Tex& t, int W, int H, unsigned seed) {
    t.W = W; t.H = H; t.host.resize((size_t)W * H * 4);
    for (auto& b : t.host) { seed = seed * 1664525u + 1013904223u; b = (unsigned char)(seed >> 24); }
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
    CK(cudaMallocArray(&t.arr, &cd, W, H, cudaArrayTextureGather));
    CK(cudaMemcpy2DToArray(t.arr, 0, 0, t.host.data(), (size_t)W * 4, (size_t)W * 4, H, cudaMemcpyHostToDevice));
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = t.arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    CK(cudaCreateTextureObject(&t.obj, &rd, &td, nullptr));
    CK(cudaMalloc(&t.lin, (size_t)W * H * 4));
    CK(cudaMemcpy(t.lin, t.host.data(), (size_t)W * H * 4, cudaMemcpyHostToDevice));
    return 0;
}

__global__ void sem_kernel(cudaTextureObject_t tex, int W, int H, const int2* pts, int n, float4* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float cu = (float)(pts[i].x + 1) * (1.0f / (float)W), cv = (float)(pts[i].y + 1) * (1.0f / (float)H);
    for (int c = 0; c < 4; ++c) out[i * 4 + c] = tex2Dgather<float4>(tex, cu, cv, c);
}

// coherent pattern: thread = pixel of a 2-D grid, uv advances slowly (like fragments of neighbouring pixels)
__global__ void thr_gather(cudaTextureObject_t t0, cudaTextureObject_t t1, int n, float du, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = i & 1023, y = i >> 10;
    const float u = x * du + 0.013f, v = y * du + 0.021f;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 a = tex2Dgather<float4>(t0, u, v, c);
        const float4 b = tex2Dgather<float4>(t1, u, v, c);
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    }
    out[i] = acc;
}
__global__ void thr_ldg(const uint32_t* __restrict__ l0, const uint32_t* __restrict__ l1, int W0, int W1, int n, float du, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = i & 1023, y = i >> 10;
    const float u = x * du + 0.013f, v = y * du + 0.021f;
    const int x0 = (int)(u * W0) & (W0 - 2), y0 = (int)(v * W0) & (W0 - 2), x1 = (int)(u * W1) & (W1 - 2), y1 = (int)(v * W1) & (W1 - 2);
    uint32_t s = __ldg(l0 + y0 * W0 + x0) + __ldg(l0 + y0 * W0 + x0 + 1) + __ldg(l0 + (y0 + 1) * W0 + x0) + __ldg(l0 + (y0 + 1) * W0 + x0 + 1);
    s += __ldg(l1 + y1 * W1 + x1) + __ldg(l1 + y1 * W1 + x1 + 1) + __ldg(l1 + (y1 + 1) * W1 + x1) + __ldg(l1 + (y1 + 1) * W1 + x1 + 1);
    out[i] = (float)s;
}

int main() {
    Tex big, half, npot;
    if (make_tex(big, 2048, 2048, 1) || make_tex(half, 1024, 1024, 2) || make_tex(npot, 100, 60, 3)) return 1;
    // ---- semantics ----
    for (Tex* t : {&npot, &big}) {
        std::vector<int2> pts;
        for (int y = -1; y < t->H; y += (t->H > 100 ? 97 : 1))
            for (int x = -1; x < t->W; x += (t->W > 100 ? 89 : 1)) pts.push_back(make_int2(x, y));
        pts.push_back(make_int2(t->W - 1, t->H - 1)); pts.push_back(make_int2(-1, t->H - 1)); pts.push_back(make_int2(t->W - 1, -1));
        int2* dp; float4* dout; const int n = (int)pts.size();
        CK(cudaMalloc(&dp, n * sizeof(int2))); CK(cudaMalloc(&dout, (size_t)n * 4 * sizeof(float4)));
        CK(cudaMemcpy(dp, pts.data(), n * sizeof(int2), cudaMemcpyHostToDevice));
        sem_kernel<<<(n + 127) / 128, 128>>>(t->obj, t->W, t->H, dp, n, dout);
        CK(cudaDeviceSynchronize());
        std::vector<float4> out((size_t)n * 4);
        CK(cudaMemcpy(out.data(), dout, out.size() * sizeof(float4), cudaMemcpyDeviceToHost));
        // candidate orders: component k of the float4 = texel (dx[k], dy[k]) of the footprint
        long bad_order = 0, bad_div = 0, bad_mul = 0;
        for (int i = 0; i < n; ++i) {
            const int x0 = (pts[i].x + t->W) % t->W, y0 = (pts[i].y + t->H) % t->H, x1 = (x0 + 1) % t->W, y1 = (y0 + 1) % t->H;
            for (int c = 0; c < 4; ++c) {
                const float4 g = out[(size_t)i * 4 + c];
                auto tx = [&](int x, int y) { return t->host[((size_t)y * t->W + x) * 4 + c]; };
                const unsigned char e[4] = {tx(x0, y1), tx(x1, y1), tx(x1, y0), tx(x0, y0)};  // CUDA doc order: x=(0,1) y=(1,1) z=(1,0) w=(0,0)
                const float gv[4] = {g.x, g.y, g.z, g.w};
                for (int k = 0; k < 4; ++k) {
                    if (fabsf(gv[k] - e[k] / 255.0f) > 1e-6f) ++bad_order;
                    if (gv[k] != (float)e[k] / 255.0f) ++bad_div;
                    if (gv[k] != (float)e[k] * (1.0f / 255.0f)) ++bad_mul;
                }
            }
        }
        printf("semantics %dx%d: %d footprints, wrong texel (doc order x=(0,1) y=(1,1) z=(1,0) w=(0,0)): %ld, != b/255.f: %ld, != b*(1/255.f): %ld\n",
               t->W, t->H, n, bad_order, bad_div, bad_mul);
        cudaFree(dp); cudaFree(dout);
    }
    // ---- throughput ----
    const int n = 1024 * 4096;
    float* dout; CK(cudaMalloc(&dout, (size_t)n * 4));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    for (float du : {1.0f / 2048.f, 1.0f / 1024.f, 1.0f / 512.f}) {
        float best_g = 1e9f, best_l = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(a); thr_gather<<<n / 128, 128>>>(big.obj, half.obj, n, du, dout); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); best_g = fminf(best_g, ms);
            cudaEventRecord(a); thr_ldg<<<n / 128, 128>>>(big.lin, half.lin, 2048, 1024, n, du, dout); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            cudaEventElapsedTime(&ms, a, b); best_l = fminf(best_l, ms);
        }
        const double clk = 1.965e9;
        printf("du=1/%g: gather x8: %.1f us  (%.2f thread-gathers/clk/SM)   ldg x8: %.1f us (%.2f thread-loads/clk/SM)\n", 1.0 / du, best_g * 1e3,
               (double)n * 8 / (best_g * 1e-3 * clk * prop.multiProcessorCount), best_l * 1e3, (double)n * 8 / (best_l * 1e-3 * clk * prop.multiProcessorCount));
    }
    return 0;
}
