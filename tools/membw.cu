// membw.cu -- practical HBM ceilings for this path's access pattern (tuning aid).
//   copy     : 1 read + 1 write stream (what MEASURED_PEAKS.json's hbm_gbs measures)
//   r3w1     : 3 read streams + 1 write stream, float4, the stencil's compulsory pattern (16 B/pt)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/membw.cu -o gpurun_out/membw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_copy(const float4 *__restrict__ a, float4 *__restrict__ o, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = a[i];
}
__global__ void k_r3w1(const float4 *__restrict__ a, const float4 *__restrict__ b, const float4 *__restrict__ c,
                       float4 *__restrict__ o, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 x = a[i], y = b[i], z = c[i];
        o[i] = make_float4(x.x + y.x * z.x, x.y + y.y * z.y, x.z + y.z * z.z, x.w + y.w * z.w);
    }
}
__global__ void k_r3w1_flat(const float4 *__restrict__ a, const float4 *__restrict__ b, const float4 *__restrict__ c,
                            float4 *__restrict__ o, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) {
        float4 x = a[i], y = b[i], z = c[i];
        o[i] = make_float4(x.x + y.x * z.x, x.y + y.y * z.y, x.z + y.z * z.z, x.w + y.w * z.w);
    }
}

int main()
{
    const size_t n = (size_t)520 * 520 * 520 / 4;  // one 512^3 level (padded), in float4
    float4 *a, *b, *c, *o;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&c, n * 16); cudaMalloc(&o, n * 16);
    cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16); cudaMemset(c, 0, n * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e9;
            for (int r = 0; r < 8; ++r) {
                cudaEventRecord(e0);
                if (mode == 0) k_copy<<<grid, 256>>>(a, o, n);
                else if (mode == 1) k_r3w1<<<grid, 256>>>(a, b, c, o, n);
                else k_r3w1_flat<<<(unsigned)((n + 255) / 256), 256>>>(a, b, c, o, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (r > 1 && ms < best) best = ms;
            }
            const double bytes = (mode == 0 ? 2.0 : 4.0) * n * 16;
            printf("%-10s grid %5d: %8.3f ms  %8.1f GB/s%s\n", mode == 0 ? "copy" : mode == 1 ? "r3w1" : "r3w1_flat", grid,
                   best, bytes / best / 1e6, mode ? "   (= Gpts/s x16)" : "");
        }
    }
    printf("cudaError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
