// fdtd_common.cuh -- shared types of the B200 FDTD hot path (device + host).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define FDTD_HALO 4          // reference main.cpp:27-32: HALO == STENCIL_ORDER == 4 cells
#define FDTD_WARMUP_STEPS 5  // reference openacc.cpp:5
#define FDTD_LEVELS 4        // device levels of u: the ABI's 3-level ring + 1 work level for two-step passes

#define FDTD_CHECK(expr)                         \
    do {                                         \
        cudaError_t _e = (expr);                 \
        if (_e != cudaSuccess) return (int)_e;   \
    } while (0)

namespace fdtd {

// Scalar coefficients of one step, all fp32, computed on the host exactly as openacc.cpp:84-87.
struct Coef {
    float dt2;   // dt*dt
    float r1;    // 1/(dt*dt)
    float n2r1;  // -2.0f*r1
    float r2, r3, r4;  // 1/h_x^2, 1/h_y^2, 1/h_z^2
    // contracted form only: dt2*r{2,3,4}*{4/3, -1/12} per axis and dt2*(r2+r3+r4)*(-5/2)
    float fx1, fx2, fy1, fy2, fz1, fz2, f0;
};

// Space orders beyond the reference's 4 (main.cpp is parameterised by STENCIL_ORDER, HALO == order):
// second-derivative weights c[0..R] of order 2R (R = order/2 <= 6) and their contracted-form products.
#define FDTD_MAX_RADIUS 6
struct OrderCoef {
    int R;
    float c[FDTD_MAX_RADIUS + 1];                                                  // exact form: c[0] centre, c[k] = +-k
    float fx[FDTD_MAX_RADIUS + 1], fy[FDTD_MAX_RADIUS + 1], fz[FDTD_MAX_RADIUS + 1];  // dt2*r{2,3,4}*c[k]
    float f0;                                                                      // dt2*(r2+r3+r4)*c[0]
};

// One grid cell that receives source contributions (padded local coordinates).
struct SourceCell {
    int X, Y, Z;
    int first, count;  // range in the contribution list, ascending p_src (the serial oracle's order)
};
struct SourceContrib {
    int p;     // source index (column of src[time][p], row of mbase)
    float w;   // ((1e-2f*wx)*wy)*wz  -- openacc.cpp:134, time independent
};

// Device view of the per-step scatter work.
struct SourceView {
    const int *plane_off;          // [nxp+1] offsets into cells[] by padded X, interior cells only
    const SourceCell *cells;       // interior cells sorted by (X,Y,Z)
    const SourceContrib *contribs;
    const float *src_row;          // src + time*pstride
    const float *mbase;            // m at each source's base corner, indexed by p
    int ncells;
};

// Neighbours of this slab along x (side 0 = lower planes, side 1 = upper planes).  The stencil kernel
// stores its two boundary planes straight into the neighbour's ghost planes through a peer-mapped
// pointer (NVLink) and, when the last CTA touching that boundary is done, raises the neighbour's
// ready flag with the step's epoch; the next step's producer threads wait on their own flags before
// loading ghost planes.  No separate exchange kernel, copy or collective exists.
struct SlabLink {
    float *peer_u[2];        // neighbour's u base (peer-mapped); nullptr = physical boundary
    long long peer_lvl[2];   // neighbour's elements per level
    int peer_edge[2];        // side 0: the neighbour's X1 (my plane X0+i lands on its ghost plane X1+i);
                             // side 1: the neighbour's X0 (my plane X1-i lands on its ghost plane X0-i)
    int *peer_flag[2];       // flag in the neighbour's memory that THIS slab raises
    int *my_flag[2];         // flags in this slab's memory raised by the neighbours
    int *counter;            // [2] CTAs of this launch that finished each boundary
    int *err;                // set to 1 if a flag wait timed out
    int expect[2];           // CTAs touching each boundary in this launch
    int epoch;               // sequence number of this step (same on every slab)
    int wait;                // 1: ghost planes of u[t0] were produced by the neighbours' step epoch-1
    int depth;               // boundary planes pushed per side: 2 (one step per pass), 4 when two-step passes are in use
    // Per-tile completion flags (one int per (y,z) tile and side, in the tail of every slab's u allocation): a boundary CTA
    // stores the launch's epoch into the NEIGHBOUR's array when it is done (boundary planes stored, ghost planes read).  When
    // the previous launch had the same kernel and tile grid (tile_mode), a boundary CTA waits only for the 3 x 3 tiles around
    // its own instead of for the whole boundary, so slabs need neither short boundary chunks nor lock step.
    int *my_tile[2];         // arrays in this slab's memory raised by the neighbours
    int *peer_tile[2];       // arrays in the neighbours' memory that THIS slab raises
    int tile_mode;           // 1: wait on my_tile[side][3x3 around the tile] >= epoch-1 instead of my_flag[side]
    // PULL mode (option "halo_pull", off by default): instead of storing its boundary planes into the neighbours' ghost
    // planes, a slab leaves them where they are and the neighbours' TMA producers read them straight from this slab's
    // memory through peer tensor maps.  Same flags, no remote stores to wait for -- but the remote loads sit on the
    // boundary CTAs' critical path, and on 8 GPUs pushing is 2-5 % faster (profiles/r02_slab_protocol_ab_8gpu.txt).
    int pull;
    int peer_nxp[2];         // padded planes of the neighbours' arrays (the peer tensor maps' x extent)
};
constexpr int kMaxFlagTiles = 16384;  // tiles per side the per-tile flag arrays can hold

// Field geometry of one slab as the kernels see it.
struct Grid {
    int nxp, nyp, nzp;       // padded extents
    int X0, X1;              // padded x range [X0, X1) updated by this launch
    int Y0, Y1, Z0, Z1;      // padded y / z interior ranges [.., ..)
    long long lvl;           // elements per level
};

}  // namespace fdtd
