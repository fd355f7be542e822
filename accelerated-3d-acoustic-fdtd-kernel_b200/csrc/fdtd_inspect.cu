// fdtd_inspect.cu -- looking at a resident field without pulling it to the host: windows and checksums.
//
// The 1024^3 / 2048^3 configurations hold 13-104 GB of u on the devices (SURVEY 8b "Resident API":
// _download_level / _checksum / _window).  What the reference's driver compares after a run is all three
// levels of u (main.cpp:573-604); these accessors let a harness do the same on cropped windows (around the
// sources, across slab seams) and prove with an order-independent checksum that nothing outside the windows
// differs: integer sums of bit patterns add up over slabs, so N slabs can be compared with one GPU.
#include "fdtd_plan.h"

#include <string.h>

namespace {

struct Window {
    int x0, x1, y0, y1, z0, z1;  // padded local coordinates, half open
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// One warp per (x, y) row of the window, lanes stride over z.  Integer sums are exact and order independent;
// sum_sq is accumulated in double (its last bits depend on the order of the atomics -- it is a norm, not a hash).
__global__ void checksum_kernel(const float *__restrict__ lvl, fdtd::Grid g, Window w, long long x_offset,
                                unsigned long long *out_u64, double *out_sq, unsigned *out_max)
{
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)(w.x1 - w.x0) * (w.y1 - w.y0);
    unsigned long long bit_sum = 0, pos_sum = 0, nonzero = 0, nonfinite = 0;
    double sq = 0.0;
    float mx = 0.f;
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < nrows;
         row += (long long)gridDim.x * (blockDim.x >> 5)) {
        const int X = w.x0 + (int)(row / (w.y1 - w.y0)), Y = w.y0 + (int)(row % (w.y1 - w.y0));
        const long long base = ((long long)X * g.nyp + Y) * g.nzp;
        const unsigned long long gbase = (unsigned long long)(((long long)X + x_offset) * g.nyp + Y) * (unsigned long long)g.nzp;
        for (int Z = w.z0 + lane; Z < w.z1; Z += 32) {
            const float v = lvl[base + Z];
            const unsigned b = __float_as_uint(v);
            bit_sum += b;
            pos_sum += (unsigned long long)b * (gbase + (unsigned long long)Z + 1ull);
            nonzero += (b & 0x7fffffffu) != 0u;
            const bool fin = (b & 0x7f800000u) != 0x7f800000u;
            nonfinite += !fin;
            if (fin) {
                sq += (double)v * (double)v;
                mx = fmaxf(mx, fabsf(v));
            }
        }
    }
    bit_sum = warp_sum(bit_sum);
    pos_sum = warp_sum(pos_sum);
    nonzero = warp_sum(nonzero);
    nonfinite = warp_sum(nonfinite);
    sq = warp_sum(sq);
    mx = warp_max(mx);
    if (lane == 0) {
        atomicAdd(out_u64 + 0, bit_sum);
        atomicAdd(out_u64 + 1, pos_sum);
        atomicAdd(out_u64 + 2, nonzero);
        atomicAdd(out_u64 + 3, nonfinite);
        atomicAdd(out_sq, sq);
        atomicMax(out_max, __float_as_uint(mx));  // non-negative floats order like their bit patterns
    }
}

bool window_ok(const fdtd_b200_plan *p, const Window &w)
{
    return w.x0 >= 0 && w.y0 >= 0 && w.z0 >= 0 && w.x1 <= p->g.nxp && w.y1 <= p->g.nyp && w.z1 <= p->g.nzp && w.x0 < w.x1 &&
           w.y0 < w.y1 && w.z0 < w.z1;
}

}  // namespace

extern "C" int fdtd_b200_plan_download_window(fdtd_b200_plan *p, int ring_level, int x0, int x1, int y0, int y1, int z0,
                                              int z1, float *host)
{
    if (!p || !host || ring_level < 0 || ring_level > 2) return (int)cudaErrorInvalidValue;
    const Window w{x0, x1, y0, y1, z0, z1};
    if (!window_ok(p, w)) return (int)cudaErrorInvalidValue;
    FDTD_CHECK(cudaSetDevice(p->dev));
    // one strided 3-D copy: a level is a pitched array (row = nzp floats, nyp rows per plane, nxp planes)
    cudaMemcpy3DParms c{};
    c.srcPtr = make_cudaPitchedPtr(fdtd_b200_plan_level(p, ring_level), (size_t)p->g.nzp * sizeof(float), (size_t)p->g.nzp,
                                   (size_t)p->g.nyp);
    c.srcPos = make_cudaPos((size_t)z0 * sizeof(float), (size_t)y0, (size_t)x0);
    c.dstPtr = make_cudaPitchedPtr(host, (size_t)(z1 - z0) * sizeof(float), (size_t)(z1 - z0), (size_t)(y1 - y0));
    c.dstPos = make_cudaPos(0, 0, 0);
    c.extent = make_cudaExtent((size_t)(z1 - z0) * sizeof(float), (size_t)(y1 - y0), (size_t)(x1 - x0));
    c.kind = cudaMemcpyDeviceToHost;
    FDTD_CHECK(cudaMemcpy3DAsync(&c, p->stream));
    FDTD_CHECK(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int fdtd_b200_plan_checksum(fdtd_b200_plan *p, int ring_level, int x0, int x1, int y0, int y1, int z0, int z1,
                                       fdtd_b200_checksum *out)
{
    if (!p || !out || ring_level < 0 || ring_level > 2) return (int)cudaErrorInvalidValue;
    const Window w{x0, x1, y0, y1, z0, z1};
    if (!window_ok(p, w)) return (int)cudaErrorInvalidValue;
    FDTD_CHECK(cudaSetDevice(p->dev));
    // 48 bytes of scratch behind the slab-protocol words in the flag block (words 16..27 of 64)
    unsigned long long *d_u64 = reinterpret_cast<unsigned long long *>(p->d_flags + 16);
    double *d_sq = reinterpret_cast<double *>(d_u64 + 4);
    unsigned *d_max = reinterpret_cast<unsigned *>(d_u64 + 5);
    FDTD_CHECK(cudaMemsetAsync(d_u64, 0, 48, p->stream));
    const long long nrows = (long long)(x1 - x0) * (y1 - y0);
    const int warps_per_block = 8;
    long long blocks = (nrows + warps_per_block - 1) / warps_per_block;
    if (blocks > 148 * 16) blocks = 148 * 16;
    checksum_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, p->stream>>>(fdtd_b200_plan_level(p, ring_level), p->g, w,
                                                                             (long long)p->shape.x_offset, d_u64, d_sq, d_max);
    FDTD_CHECK(cudaGetLastError());
    unsigned long long h[6];
    FDTD_CHECK(cudaMemcpyAsync(h, d_u64, 48, cudaMemcpyDeviceToHost, p->stream));
    FDTD_CHECK(cudaStreamSynchronize(p->stream));
    out->bit_sum = h[0];
    out->pos_sum = h[1];
    out->nonzero = h[2];
    out->nonfinite = h[3];
    memcpy(&out->sum_sq, &h[4], sizeof(double));
    const unsigned mb = (unsigned)(h[5] & 0xffffffffull);
    memcpy(&out->max_abs, &mb, sizeof(float));
    out->reserved = 0;
    return 0;
}
