// stencil_generic.cu -- Section0 for ANY extents/alignment (one point per thread, neighbour reuse
// through L1/L2), the Section1 stand-alone scatter, and small utility kernels.
//
// Replaces reference cuda.cu:61-108 (stencil_update_kernel_1step) and cuda.cu:112-170 /
// cuda_optimized.cu:241-260 (source_inject_kernel).  Unlike the reference's plain kernel the
// thread x index walks the CONTIGUOUS z axis, and the scatter is atomics-free: one thread owns
// one cell and adds its contributions in p_src order (the serial order of openacc.cpp:116-136).
#include "fdtd_arith.cuh"
#include "fdtd_kernels.cuh"
#include "tma_ptx.cuh"

namespace fdtd {

template <bool EXACT>
__global__ void __launch_bounds__(256) stencil_generic_kernel(StepArgs a)
{
    const int Z = a.g.Z0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = a.g.Y0 + blockIdx.y * blockDim.y + threadIdx.y;
    const int X = a.g.X0 + blockIdx.z;
    if (Z >= a.g.Z1 || Y >= a.g.Y1) return;

    const long long sy = a.g.nzp, sx = (long long)a.g.nyp * a.g.nzp;
    const long long c = (long long)X * sx + (long long)Y * sy + Z;
    const float *__restrict__ u0 = a.u + a.t0 * a.g.lvl;
    const float *__restrict__ u1 = a.u + a.t1 * a.g.lvl;
    float *__restrict__ u2 = a.u + a.t2 * a.g.lvl;

    // the same per-point function as the streaming kernel: both kernels agree bit for bit in both modes
    float v = point<EXACT>(__ldg(u0 + c), __ldg(u0 + c - 2 * sx), __ldg(u0 + c - sx), __ldg(u0 + c + sx), __ldg(u0 + c + 2 * sx),
                           __ldg(u0 + c - 2 * sy), __ldg(u0 + c - sy), __ldg(u0 + c + sy), __ldg(u0 + c + 2 * sy),
                           __ldg(u0 + c - 2), __ldg(u0 + c - 1), __ldg(u0 + c + 1), __ldg(u0 + c + 2), __ldg(u1 + c),
                           __ldg(a.m + c), a.k);

    if (a.sv.ncells > 0) {  // fused Section1: the owner of a source cell adds its contributions
        const int c0 = a.sv.plane_off[X], c1 = a.sv.plane_off[X + 1];
        for (int i = c0; i < c1; ++i) {
            const SourceCell cell = a.sv.cells[i];
            if (cell.Y == Y && cell.Z == Z) v = apply_cell(v, cell, a.sv);
        }
    }
    u2[c] = v;
}

int launch_stencil_generic(const StepArgs &a, bool exact, cudaStream_t stream)
{
    const int nz = a.g.Z1 - a.g.Z0, ny = a.g.Y1 - a.g.Y0, nx = a.g.X1 - a.g.X0;
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    dim3 block(64, 4, 1);
    if (nz <= 32) block = dim3(32, 8, 1);
    dim3 grid((nz + block.x - 1) / block.x, (ny + block.y - 1) / block.y, nx);
    if (grid.y > 65535 || grid.z > 65535) return (int)cudaErrorInvalidValue;
    if (exact)
        stencil_generic_kernel<true><<<grid, block, 0, stream>>>(a);
    else
        stencil_generic_kernel<false><<<grid, block, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------- space orders 6..12
// One point per thread like the kernel above; R pairs of neighbours per axis.  EXACT replays the oracle's operation
// order: the outermost neighbour pair first, which at R = 2 is the reference's (openacc.cpp:102-107).
template <int R, bool EXACT>
__global__ void __launch_bounds__(256) stencil_order_kernel(StepArgs a, OrderCoef oc)
{
    const int Z = a.g.Z0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = a.g.Y0 + blockIdx.y * blockDim.y + threadIdx.y;
    const int X = a.g.X0 + blockIdx.z;
    if (Z >= a.g.Z1 || Y >= a.g.Y1) return;
    const long long sy = a.g.nzp, sx = (long long)a.g.nyp * a.g.nzp;
    const long long c = (long long)X * sx + (long long)Y * sy + Z;
    const float *__restrict__ u0 = a.u + a.t0 * a.g.lvl;
    const float *__restrict__ u1 = a.u + a.t1 * a.g.lvl;
    float *__restrict__ u2 = a.u + a.t2 * a.g.lvl;
    const float uc = __ldg(u0 + c), up = __ldg(u1 + c), mm = __ldg(a.m + c);
    float v;
    if (EXACT) {
        const float r5 = __fmul_rn(oc.c[0], uc);
        float dx = r5, dy = r5, dz = r5;
#pragma unroll
        for (int k = R; k >= 1; --k) {
            dx = __fadd_rn(dx, __fmul_rn(oc.c[k], __fadd_rn(__ldg(u0 + c - k * sx), __ldg(u0 + c + k * sx))));
            dy = __fadd_rn(dy, __fmul_rn(oc.c[k], __fadd_rn(__ldg(u0 + c - k * sy), __ldg(u0 + c + k * sy))));
            dz = __fadd_rn(dz, __fmul_rn(oc.c[k], __fadd_rn(__ldg(u0 + c - k), __ldg(u0 + c + k))));
        }
        v = leapfrog_exact(uc, dx, dy, dz, up, mm, a.k);
    } else {
        float acc = oc.f0 * uc;
#pragma unroll
        for (int k = R; k >= 1; --k) acc = fmaf(oc.fx[k], __ldg(u0 + c - k * sx) + __ldg(u0 + c + k * sx), acc);
#pragma unroll
        for (int k = R; k >= 1; --k) acc = fmaf(oc.fy[k], __ldg(u0 + c - k * sy) + __ldg(u0 + c + k * sy), acc);
#pragma unroll
        for (int k = R; k >= 1; --k) acc = fmaf(oc.fz[k], __ldg(u0 + c - k) + __ldg(u0 + c + k), acc);
        float rm;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rm) : "f"(mm));
        v = fmaf(acc, rm, fmaf(2.0f, uc, -up));
    }
    if (a.sv.ncells > 0) {
        const int c0 = a.sv.plane_off[X], c1 = a.sv.plane_off[X + 1];
        for (int i = c0; i < c1; ++i) {
            const SourceCell cell = a.sv.cells[i];
            if (cell.Y == Y && cell.Z == Z) v = apply_cell(v, cell, a.sv);
        }
    }
    u2[c] = v;
}

int launch_stencil_order(const StepArgs &a, const OrderCoef &oc, bool exact, cudaStream_t stream)
{
    const int nz = a.g.Z1 - a.g.Z0, ny = a.g.Y1 - a.g.Y0, nx = a.g.X1 - a.g.X0;
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    dim3 block(64, 4, 1);
    if (nz <= 32) block = dim3(32, 8, 1);
    dim3 grid((nz + block.x - 1) / block.x, (ny + block.y - 1) / block.y, nx);
    if (grid.y > 65535 || grid.z > 65535) return (int)cudaErrorInvalidValue;
#define FDTD_ORDER_CASE(R_)                                                          \
    case R_:                                                                         \
        if (exact)                                                                   \
            stencil_order_kernel<R_, true><<<grid, block, 0, stream>>>(a, oc);       \
        else                                                                         \
            stencil_order_kernel<R_, false><<<grid, block, 0, stream>>>(a, oc);      \
        break;
    switch (oc.R) {
        FDTD_ORDER_CASE(2)
        FDTD_ORDER_CASE(3)
        FDTD_ORDER_CASE(4)
        FDTD_ORDER_CASE(5)
        FDTD_ORDER_CASE(6)
    default:
        return (int)cudaErrorInvalidValue;
    }
#undef FDTD_ORDER_CASE
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------- receiver sampling
// One thread per receiver; the eight corners are summed in the oracle's order (x outermost, z innermost), every
// weight product as ((wx*wy)*wz)*u with one rounding per operation, so the trace is bit-identical to the oracle's.
__global__ void sample_receivers_kernel(const float *__restrict__ u, Grid g, const ReceiverPoint *__restrict__ pts, int npts,
                                        float *__restrict__ rec_row)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const ReceiverPoint r = pts[i];
    float sum = 0.0f;
#pragma unroll
    for (int rx = 0; rx <= 1; ++rx)
#pragma unroll
        for (int ry = 0; ry <= 1; ++ry)
#pragma unroll
            for (int rz = 0; rz <= 1; ++rz) {
                if (!((r.mask >> (rx * 4 + ry * 2 + rz)) & 1u)) continue;
                // r*p + (1 - r)*(1 - p) with r in {0, 1} is exactly p or 1 - p (x + 0 == x for the x >= 0 that occur)
                const float wx = rx ? r.fx : __fsub_rn(1.0f, r.fx);
                const float wy = ry ? r.fy : __fsub_rn(1.0f, r.fy);
                const float wz = rz ? r.fz : __fsub_rn(1.0f, r.fz);
                const float v = u[((long long)(r.X + rx) * g.nyp + (r.Y + ry)) * g.nzp + (r.Z + rz)];
                sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(__fmul_rn(wx, wy), wz), v));
            }
    rec_row[r.index] = sum;
}

int launch_sample_receivers(const float *u_level, const Grid &g, const ReceiverPoint *pts, int npts, float *rec_row,
                            cudaStream_t stream)
{
    if (npts <= 0) return 0;
    sample_receivers_kernel<<<(npts + 127) / 128, 128, 0, stream>>>(u_level, g, pts, npts, rec_row);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------- Section1 scatter
__global__ void scatter_kernel(float *__restrict__ u2, Grid g, const SourceCell *__restrict__ cells, int ncells,
                               SourceView sv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncells) return;
    const SourceCell cell = cells[i];
    const long long c = ((long long)cell.X * g.nyp + cell.Y) * g.nzp + cell.Z;
    u2[c] = apply_cell(u2[c], cell, sv);
}

int launch_scatter(float *u_level, const Grid &g, const SourceCell *cells, int ncells,
                   const SourceContrib *contribs, const float *src_row, const float *mbase,
                   cudaStream_t stream)
{
    if (ncells <= 0) return 0;
    SourceView sv{};
    sv.contribs = contribs;
    sv.src_row = src_row;
    sv.mbase = mbase;
    sv.ncells = ncells;
    scatter_kernel<<<(ncells + 127) / 128, 128, 0, stream>>>(u_level, g, cells, ncells, sv);
    return (int)cudaGetLastError();
}

__global__ void gather_mbase_kernel(const float *__restrict__ m, const long long *__restrict__ idx,
                                    float *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = idx[i] >= 0 ? m[idx[i]] : 1.0f;
}

int launch_gather_mbase(const float *m, const long long *base_idx, float *mbase, int n, cudaStream_t stream)
{
    if (n <= 0) return 0;
    gather_mbase_kernel<<<(n + 127) / 128, 128, 0, stream>>>(m, base_idx, mbase, n);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------- fills
__global__ void fill_kernel(float4 *p4, float *p, size_t n4, size_t n, float v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4 v4 = make_float4(v, v, v, v);
    for (size_t i = t; i < n4; i += stride) p4[i] = v4;
    for (size_t i = n4 * 4 + t; i < n; i += stride) p[i] = v;
}

int launch_fill(float *p, size_t n, float v, cudaStream_t stream)
{
    if (n == 0) return 0;
    const size_t n4 = (((uintptr_t)p & 15) == 0) ? n / 4 : 0;
    fill_kernel<<<148 * 8, 256, 0, stream>>>((float4 *)p, p, n4, n, v);
    return (int)cudaGetLastError();
}

// Dense parity field of reference main.cpp:525-532: m = 1.5, u[0] = u[1] = sin(i*0.001f)*10+100 over
// the whole padded volume (halos included), i = GLOBAL linear index; u[2] = 0.
__global__ void fill_dense_kernel(float *__restrict__ u, float *__restrict__ m, size_t volp, size_t goff)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < volp; i += stride) {
        const float val = __fadd_rn(__fmul_rn(sinf(__fmul_rn((float)(i + goff), 0.001f)), 10.0f), 100.0f);
        m[i] = 1.5f;
        u[i] = val;
        u[volp + i] = val;
        u[2 * volp + i] = 0.0f;
    }
}

int launch_fill_dense(float *u, float *m, int nxp, int nyp, int nzp, long long x_plane_offset, cudaStream_t stream)
{
    const size_t volp = (size_t)nxp * nyp * nzp;
    fill_dense_kernel<<<148 * 8, 256, 0, stream>>>(u, m, volp, (size_t)x_plane_offset * nyp * nzp);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------- ghost refresh (pull protocol)
__global__ void ghost_refresh_kernel(float *u, Grid g, SlabLink lk, int epoch)
{
    const int side = blockIdx.y;
    if (!lk.peer_u[side]) return;
    if (threadIdx.x == 0) wait_flag(lk.my_flag[side], epoch, lk.err);  // the neighbour's last launch has written its boundary planes
    __syncthreads();
    const long long plane = (long long)g.nyp * g.nzp;
    const int my_x = side == 0 ? g.X0 - FDTD_HALO : g.X1;                      // first of this slab's ghost planes on that side
    const int peer_x = side == 0 ? lk.peer_edge[0] - FDTD_HALO : lk.peer_edge[1];  // the same planes where the neighbour computes them
    const long long n4 = FDTD_HALO * plane / 4;
    for (int l = 0; l < FDTD_LEVELS; ++l) {
        const float4 *src = reinterpret_cast<const float4 *>(lk.peer_u[side] + l * lk.peer_lvl[side] + peer_x * plane);
        float4 *dst = reinterpret_cast<float4 *>(u + l * g.lvl + my_x * plane);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
    }
}

int launch_ghost_refresh(float *u, const Grid &g, const SlabLink &lk, int epoch, cudaStream_t stream)
{
    if (!lk.peer_u[0] && !lk.peer_u[1]) return 0;
    ghost_refresh_kernel<<<dim3(128, 2), 256, 0, stream>>>(u, g, lk, epoch);
    return (int)cudaGetLastError();
}

}  // namespace fdtd
