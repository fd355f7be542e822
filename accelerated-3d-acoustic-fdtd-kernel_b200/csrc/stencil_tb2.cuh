// stencil_tb2.cuh -- what the two-step kernels (stencil_tb2.cu, stencil_tb2l.cu) share: launch arguments, the
// per-plane source injection and the instantiation-table entry.
#pragma once
#include "fdtd_arith.cuh"
#include "fdtd_kernels.cuh"
#include "tma_ptx.cuh"

namespace fdtd {

struct Tb2Args {
    alignas(64) CUtensorMap map_cur;
    alignas(64) CUtensorMap map_prev;
    alignas(64) CUtensorMap map_m;
    alignas(64) CUtensorMap map_cur_peer[2], map_prev_peer[2];  // pull mode: the neighbours' u
    Tb2Step s;
    int tiles_z, tiles_y, xchunk;
    int edge;  // > 0: the first and last chunk are `edge` planes long (slabs with neighbours)
    int prefetch;  // lean kernel: stages the producer prefetches into L2 ahead of its loads (0 = none)
};

// Source cells of one plane that fall into this thread's float4: add their contributions in p_src order.
__device__ __forceinline__ void inject_plane(float4 &r, int X, int Y, int Z, const SourceView &sv)
{
    const int c0 = sv.plane_off[X], c1 = sv.plane_off[X + 1];
    for (int q = c0; q < c1; ++q) {
        const SourceCell cell = sv.cells[q];
        const int dzc = cell.Z - Z;
        if (cell.Y == Y && dzc >= 0 && dzc < 4) {
            float v = dzc == 0 ? r.x : dzc == 1 ? r.y : dzc == 2 ? r.z : r.w;
            v = apply_cell(v, cell, sv);
            r.x = dzc == 0 ? v : r.x;
            r.y = dzc == 1 ? v : r.y;
            r.z = dzc == 2 ? v : r.z;
            r.w = dzc == 3 ? v : r.w;
        }
    }
}

typedef void (*Tb2KernelFn)(const Tb2Args);
struct Tb2Variant {
    int er, ec, rows;
    bool exact;
    Tb2KernelFn fn;
    int nt;
    size_t smem;
};
// instantiations of the lean kernel (stencil_tb2l.cu); same meaning of (er, ec) as stencil_tb2.cu's table
const Tb2Variant *tb2l_variants(int *n);

}  // namespace fdtd
