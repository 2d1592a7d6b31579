// Probe (measurement aid): what does a kernel launch cost on the event clock, as a function of dynamic shared memory and of
// the kernel that ran before it (shared-memory carve-out changes)?
#include <cstdio>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k_empty(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_touch(int* p) { extern __shared__ int sm[]; sm[threadIdx.x] = threadIdx.x; __syncthreads(); if (p && sm[(threadIdx.x + 1) % blockDim.x] == -1) *p = 1; }
__global__ void k_fill(int4* p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_int4(0, 0, 0, 0); }
int main() {
    int* d; CK(cudaMalloc(&d, 4));
    int4* big; const size_t nbig = (256u << 20) / 16; CK(cudaMalloc(&big, nbig * 16));
    CK(cudaFuncSetAttribute(k_touch, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char* name, int smem, int threads, int grid, bool flush, bool pre_small) -> int {
        std::vector<float> ts;
        for (int i = 0; i < 30; ++i) {
            if (flush) k_fill<<<1184, 256, 0, st>>>(big, nbig);
            if (pre_small) k_touch<<<1184, 128, 24 * 1024, st>>>(nullptr);
            CK(cudaStreamSynchronize(st));
            cudaEventRecord(a, st);
            if (smem < 0) k_empty<<<grid, threads, 0, st>>>(nullptr); else k_touch<<<grid, threads, smem, st>>>(nullptr);
            cudaEventRecord(b, st);
            CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms * 1e3f);
        }
        std::sort(ts.begin(), ts.end());
        printf("%-60s median %6.2f us  min %6.2f us\n", name, ts[ts.size() / 2], ts[0]);
        return 0;
    };
    run("empty 148 x 512, no smem", -1, 512, 148, false, false);
    run("touch 148 x 512, 227 KB smem", 227 * 1024, 512, 148, false, false);
    run("touch 148 x 512, 227 KB smem, after a 256 MB fill", 227 * 1024, 512, 148, true, false);
    run("touch 148 x 512, 227 KB smem, after fill + 1184 x 128 / 24 KB", 227 * 1024, 512, 148, true, true);
    run("touch 148 x 512, 100 KB smem, after a 256 MB fill", 100 * 1024, 512, 148, true, false);
    run("touch 1184 x 128, 24 KB smem, after a 256 MB fill", 24 * 1024, 128, 1184, true, false);
    run("empty 1184 x 128, after a 256 MB fill", -1, 128, 1184, true, false);
    return 0;
}
