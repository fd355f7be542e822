// pcie2d_probe.cu -- how fast are strided host->device copies (the halo shell of one level) and cudaHostRegister?
// nvcc -O2 -o build/pcie2d_probe tools/pcie2d_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const int n = 512, np = n + 8;
    const size_t plane = (size_t)np * np, lvl = plane * np;
    float *h = nullptr, *d = nullptr;
    CK(cudaMallocHost(&h, lvl * sizeof(float)));
    CK(cudaMalloc(&d, lvl * sizeof(float)));
    memset(h, 0, lvl * sizeof(float));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        // whole level
        cudaEventRecord(a, s);
        CK(cudaMemcpyAsync(d, h, lvl * sizeof(float), cudaMemcpyHostToDevice, s));
        cudaEventRecord(b, s);
        CK(cudaStreamSynchronize(s));
        cudaEventElapsedTime(&ms, a, b);
        printf("full level %.1f MB: %.2f ms (%.1f GB/s)\n", lvl * 4 / 1e6, ms, lvl * 4 / ms / 1e6);
        // seams: 32 B per row, all rows of the interior planes, in chunks of 24 planes
        double t0 = now();
        cudaEventRecord(a, s);
        for (int x = 4; x < 4 + n; x += 24) {
            const int nx = (x + 24 <= 4 + n) ? 24 : 4 + n - x;
            CK(cudaMemcpy2DAsync(d + x * plane + (np - 4), np * 4, h + x * plane + (np - 4), np * 4, 32, (size_t)nx * np, cudaMemcpyHostToDevice, s));
        }
        double t1 = now();
        cudaEventRecord(b, s);
        CK(cudaStreamSynchronize(s));
        cudaEventElapsedTime(&ms, a, b);
        printf("seams 32 B x %zu rows (%.1f MB): %.2f ms device, %.2f ms host enqueue\n", (size_t)n * np, n * np * 32 / 1e6, ms, (t1 - t0) * 1e3);
        // bands: 8 rows per plane boundary
        t0 = now();
        cudaEventRecord(a, s);
        for (int x = 4; x < 4 + n; x += 24) {
            const int nx = (x + 24 <= 4 + n) ? 24 : 4 + n - x;
            CK(cudaMemcpy2DAsync(d + x * plane + (size_t)(np - 4) * np, plane * 4, h + x * plane + (size_t)(np - 4) * np, plane * 4, 8 * np * 4, nx, cudaMemcpyHostToDevice, s));
        }
        t1 = now();
        cudaEventRecord(b, s);
        CK(cudaStreamSynchronize(s));
        cudaEventElapsedTime(&ms, a, b);
        printf("bands %d B x %d planes (%.1f MB): %.2f ms device, %.2f ms host enqueue\n", 8 * np * 4, n, n * 8.0 * np * 4 / 1e6, ms, (t1 - t0) * 1e3);
    }
    // page-locking a pageable array of the size of u + m at 512^3
    const size_t bytes = 4 * lvl * sizeof(float);
    float *p = (float *)malloc(bytes);
    memset(p, 1, bytes);
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now();
        CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
        double t1 = now();
        CK(cudaHostUnregister(p));
        double t2 = now();
        printf("cudaHostRegister %.2f GB: %.1f ms, unregister %.1f ms\n", bytes / 1e9, (t1 - t0) * 1e3, (t2 - t1) * 1e3);
    }
    return 0;
}
