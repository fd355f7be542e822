// fdtd_plan.h -- the resident plan object behind include/fdtd_b200.h (internal).
#pragma once
#include "../../include/fdtd_b200.h"
#include "fdtd_kernels.cuh"

#include <vector>

namespace fdtd {

// Array shape and update extents exactly as the reference ABI passes them (unpadded, inclusive).
struct PlanShape {
    int nxp, nyp, nzp;
    int x_m, x_M, y_m, y_M, z_m, z_M;  // local extents of this slab
    float dt, h_x, h_y, h_z, o_x, o_y, o_z;
    int x_offset;      // global index of local x = 0
    int gx_m, gx_M;    // GLOBAL x extents (== x_m, x_M for one slab)
    int deviceid;
    int space_order = 4;  // 4 (the reference's kernels) .. 12; halo cells per side == space_order (main.cpp:27-32)
};

// cache_buffers: take the field buffers from / return them to the per-process device-buffer cache (used by
// the Kernel_* entry points, whose callers run the same size repeatedly: cudaMalloc/cudaFree of GBs costs
// more than the 50 time steps).
int plan_create_internal(const PlanShape &s, fdtd_b200_plan **out, bool cache_buffers = false);

// shared between fdtd_plan.cu and fdtd_staged.cu
int env_int(const char *key, int fallback);
// kernel choice, tensor maps, two-step feasibility, mbase gather: what every run does before its first launch
int plan_prepare(fdtd_b200_plan *p);
// ring level r lives in device level r again (after an upload / fill); shell_state: 0 unknown, 1 identical shells
void reset_placement(fdtd_b200_plan *p, int shell_state);

}  // namespace fdtd

struct fdtd_b200_plan {
    fdtd::PlanShape shape{};
    int dev = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    fdtd::Grid g{};
    fdtd::Coef k{};
    float *d_u = nullptr, *d_m = nullptr;
    bool cache_buffers = false;
    size_t u_bytes = 0, m_bytes = 0;

    // sources: cells [0, ncells_int) are inside the Section0 write range (fusable), the rest are halo cells
    float *d_src = nullptr;
    int src_size0 = 0, pstride = 1;
    fdtd::SourceCell *d_cells = nullptr;
    fdtd::SourceContrib *d_contribs = nullptr;
    int *d_plane_off = nullptr;
    float *d_mbase = nullptr;
    long long *d_base_idx = nullptr;
    int ncells_int = 0, ncells_halo = 0, ncells_all = 0, n_mbase = 0;
    // for two-step passes: the interior cells plus the cells on the neighbour slabs' two nearest planes (a pass
    // recomputes those planes), sorted by (X,Y,Z); src_halo_global: some source of the GLOBAL grid touches a halo
    // cell (every slab computes the same answer, so slabs agree on whether two-step passes are possible)
    fdtd::SourceCell *d_cells2 = nullptr;
    int *d_plane_off2 = nullptr;
    int ncells2 = 0;
    bool src_halo_global = false;
    std::vector<long long> h_base_idx;  // host copy of d_base_idx (the staged run gathers mbase from the host's m)

    // space orders 6..12 (generic kernel only) and receivers
    fdtd::OrderCoef oc{};
    fdtd::ReceiverPoint *d_rec_pts = nullptr;
    int nrec_total = 0, nrec_owned = 0;
    std::vector<int> rec_owned;        // [nrec_total] 1 = this slab samples the receiver
    float *d_rec = nullptr;            // [rec_rows_cap][nrec_total]
    int rec_rows_cap = 0, rec_rows = 0, rec_time_m = 0;

    // options
    int opt_kernel = 0, opt_exact = 1, opt_fuse = 1, opt_t_fuse = 1;
    int opt_stage_planes = -1;       // x planes per block of the staged (pipelined H2D / compute / D2H) run; -1 = auto, 0 = off
    fdtd::TmaConfig cfg{};
    fdtd::TmaPlan tma{};
    fdtd::Tb2Plan tb2{};
    fdtd::Tc2Plan tc2{};             // two-step passes on 2-CTA clusters (option "cluster", unlinked slabs)
    int opt_cluster = 0;
    bool t_fuse_explicit = false;    // "t_fuse" came from set_option (honoured as is) rather than from the driver's hook / env
    bool use_tc2 = false;            // the current run's passes go through stencil_tc2
    int kernel_used = 0;
    int t_fuse_used = 1;             // time steps per pass of the current run (1 or 2)
    int t_fuse_agreed = -1;          // linked slabs: depth all slabs agreed on (-1 = not negotiated -> 1)

    // Placement of the ABI's 3-level ring in the FDTD_LEVELS device levels: ring level r lives in device level
    // phys[r], `work` is the spare one.  A two-step pass writes u^{n+2} into the spare level (it cannot overwrite
    // u^{n-1} in place: neighbouring tiles still read it) and the two swap roles.  upload/fill reset to identity.
    int phys[3] = {0, 1, 2};
    int work = 3;
    int shell_state = 0;             // 0 = unknown, 1 = halo shells of all device levels identical, 2 = they differ

    // x-slab neighbours: flag words live in the tail of the u allocation (one IPC handle covers both)
    int *d_flags = nullptr;          // [0] ready-from-lower, [1] ready-from-upper, [2..3] CTA counters, [4] error
    size_t flags_offset = 0;         // byte offset of d_flags inside the u allocation
    size_t arena_offset = 0, arena_used = 0;  // small-array arena behind the flag words (fdtd_plan.cu)
    fdtd::SlabLink link{};           // peers (null = physical boundary)
    void *ipc_base[2] = {nullptr, nullptr};  // mappings opened with cudaIpcOpenMemHandle
    int epoch = 0;                   // step sequence number, identical on every slab
    size_t tile_flags_offset = 0;    // byte offset of the per-tile flag arrays [2][kMaxFlagTiles] inside the u allocation
    int last_kind = 0;               // kernel of the previous launch of this run: 0 none, 1 one-step streaming, 2 two-step
    int opt_halo_pull = -1;          // 1: linked slabs read the neighbours' boundary planes in place instead of receiving them; 0: never;
                                     // -1 (default): for runs of the lean two-step kernel (FDTD_B200_HALO_PULL; plan_prepare)
    int opt_tile_flags = 1;          // 0: always the whole-boundary flags (FDTD_B200_TILE_FLAGS)

    // statistics of the last run
    long last_launches = 0;
    double last_kernel_seconds = 0.0;
};
