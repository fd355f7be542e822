// fdtd_kernels.cuh -- launch interface of the device kernels (implemented in stencil_*.cu, inject.cu).
#pragma once
#include "fdtd_common.cuh"

namespace fdtd {

// Everything one Section0 launch needs besides the kernel-variant specifics.
struct StepArgs {
    float *u;          // base of u[3][nxp][nyp][nzp]
    const float *m;    // m[nxp][nyp][nzp]
    Grid g;
    Coef k;
    int t0, t1, t2;    // ring: current, previous, next
    SourceView sv;     // sv.ncells == 0 -> no fused injection
    SlabLink link;     // x-slab neighbours (all null for a single slab)
};

// --- generic kernel: any extents / alignment, one point per thread, loads through L1/L2.
int launch_stencil_generic(const StepArgs &a, bool exact, cudaStream_t stream);

// --- the same for space orders 6..12 (radius 3..6, halo = order cells): exact mode follows the oracle's generalisation
// of openacc.cpp:102-107 (outermost neighbour pair first), contracted mode the FMA form.
int launch_stencil_order(const StepArgs &a, const OrderCoef &oc, bool exact, cudaStream_t stream);
// Receiver sampling: rec_row[p] = sum over the in-range trilinear corners of ((wx*wy)*wz)*u[corner],
// x outermost / z innermost; pos/frac are precomputed on the host in IEEE fp32 (same arithmetic as the injection).
struct ReceiverPoint {
    int X, Y, Z;          // padded local base corner
    float fx, fy, fz;     // fractions
    unsigned mask;        // bit rx*4+ry*2+rz set = corner in range (openacc.cpp:132 bounds) and owned data
    int index;            // column of rec[time][index]
};
int launch_sample_receivers(const float *u_level, const Grid &g, const ReceiverPoint *pts, int npts, float *rec_row,
                            cudaStream_t stream);

// --- 2.5D x-streaming kernel: TMA -> mbarrier ring in shared memory -> register queue along x.
struct TmaConfig {
    int ty = 0, tz = 0;  // (y,z) tile; 0 = auto
    int rows = 0;        // y rows per consumer thread (1, 2 or 4); 0 = auto
    int stages = 0;      // halo-plane ring depth (5 or 10); 0 = 5
    int xchunk = 0;      // x planes per CTA; 0 = auto
    int lean = 1;        // two-step passes: 1 = stencil_tb2l.cu (lean inner loop), 0 = stencil_tb2.cu
};
struct TmaPlan {         // built once per (arrays, config): tensor maps + launch shape
    alignas(64) CUtensorMap map_halo;  // u as (z,y,x,t), box (tz+8, ty+4, 1, 1)
    alignas(64) CUtensorMap map_ctr;   // u as (z,y,x,t), box (tz, ty, 1, 1)
    alignas(64) CUtensorMap map_m;     // m as (z,y,x),   box (tz, ty, 1)
    alignas(64) CUtensorMap map_halo_peer[2];  // the neighbours' u, same box as map_halo (pull mode)
    int ty, tz, rows, stages, xchunk;
    int variant;         // index into the instantiation table
    size_t smem_bytes;
    bool valid = false;
};
// Can the TMA kernel run on this geometry?  (row pitch and z origin 16-byte aligned, nz % 4 == 0)
bool tma_supported(const Grid &g);
int tma_plan_build(TmaPlan &p, float *u, const float *m, const Grid &g, const TmaConfig &cfg, bool exact,
                   int sm_count, const SlabLink *link = nullptr);
int launch_stencil_tma(const TmaPlan &p, const StepArgs &a, bool exact, cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda): fp32, no swizzle, OOB = 0.
int encode_tensor_map(CUtensorMap *map, const float *base, int rank, const cuuint64_t *dims, const cuuint32_t *box);

// --- two time steps per pass (temporal blocking): u^{n+1} and u^{n+2} from u^{n-1}, u^n and m read once.
// Levels are indices into the plan's FDTD_LEVELS-deep array.
struct Tb2Plan {
    alignas(64) CUtensorMap map_cur;   // u as (z,y,x,level), box (tz+8, ty+8, 1, 1): u^n with the radius-4 halo
    alignas(64) CUtensorMap map_prev;  // u as (z,y,x,level), box (tz+8, ty+4, 1, 1): u^{n-1} on the extended tile
    alignas(64) CUtensorMap map_m;     // m as (z,y,x), box (tz+8, ty+4, 1)
    alignas(64) CUtensorMap map_cur_peer[2], map_prev_peer[2];  // the neighbours' u, same boxes (pull mode)
    int ty, tz, rows, xchunk, variant;  // output tile, rows per thread, x planes per CTA
    bool lean = false;                  // variant indexes stencil_tb2l.cu's table
    size_t smem_bytes;
    bool valid = false;
};
struct Tb2Step {
    float *u;
    Grid g;
    Coef k;
    int l_prev, l_cur, l_n1, l_n2;     // level indices of u^{n-1}, u^n (inputs) and u^{n+1}, u^{n+2} (outputs)
    SourceView sv;                     // cells incl. the neighbours' two nearest planes; sv.src_row = src row of step n
    const float *src_row2;             // src row of step n+1
    SlabLink link;                     // x-slab neighbours (all null for a single slab)
};
// --- the same two-step pass as a time-step pipeline across a 2-CTA cluster (stencil_tc2.cu): CTA 0 computes u^{n+1}
// and streams it into CTA 1's shared memory (DSMEM), CTA 1 computes u^{n+2}.  Unlinked slabs only.
struct Tc2Plan {
    alignas(64) CUtensorMap map_cur, map_prev, map_m;  // step 1: as Tb2Plan
    alignas(64) CUtensorMap map_ctr, map_mc;           // step 2: u^n and m on the output tile, boxes (tz, ty)
    int ty, tz, rows, xchunk, variant;
    int npairs;                                        // resident CTA pairs (SMs / 2): each loops over (tile, chunk) items
    size_t smem_bytes;
    bool valid = false;
};
int tc2_plan_build(Tc2Plan &p, float *u, const float *m, const Grid &g, const TmaConfig &cfg, bool exact, int sm_count);
int launch_stencil_tc2(const Tc2Plan &p, const Tb2Step &a, bool exact, cudaStream_t stream);

constexpr int kSlabEdgePlanes = 8;     // shortest slab side that can run linked two-step passes is 4 x this
// Length of the two boundary chunks of a linked slab of nx planes (they are dispatched first and raise the neighbours'
// flags): FDTD_B200_SLAB_EDGE overrides; 2*edge == nx means "no chunks in between" (the slabs then run in lock step).
int slab_edge_planes(int nx, int xchunk, int tiles, int slots);
int tb2_plan_build(Tb2Plan &p, float *u, const float *m, const Grid &g, const TmaConfig &cfg, bool exact, int sm_count,
                   const SlabLink *link = nullptr);
int launch_stencil_tb2(const Tb2Plan &p, const Tb2Step &a, bool exact, cudaStream_t stream);
// 1 in *flag (device) unless the shells (every padded cell outside g's box) of levels 0..2 are bit-identical
int launch_shell_check(float *u, const Grid &g, int *flag, cudaStream_t stream);
// shell of level `from` -> level `to`
int launch_shell_copy(float *u, const Grid &g, int from, int to, cudaStream_t stream);

// Pull protocol: at the end of a run, copy the neighbours' four boundary planes of every device level into this slab's ghost planes
// (waits in-kernel for the neighbours' last launch, `epoch`, to have raised this slab's whole-boundary flags).
int launch_ghost_refresh(float *u, const Grid &g, const SlabLink &lk, int epoch, cudaStream_t stream);

// --- Section1 stand-alone scatter: one thread per cell, contributions added in p_src order.
int launch_scatter(float *u_level, const Grid &g, const SourceCell *cells, int ncells,
                   const SourceContrib *contribs, const float *src_row, const float *mbase,
                   cudaStream_t stream);
// mbase[i] = m[base_idx[i]] (base corner of every source), i < n.
int launch_gather_mbase(const float *m, const long long *base_idx, float *mbase, int n, cudaStream_t stream);
// Fills.
int launch_fill(float *p, size_t n, float v, cudaStream_t stream);
int launch_fill_dense(float *u, float *m, int nxp, int nyp, int nzp, long long x_plane_offset, cudaStream_t stream);

}  // namespace fdtd
