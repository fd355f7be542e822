/*
 * fdtd_b200.h -- C ABI of libfdtd_b200.so: the B200 (sm_100a) implementation of the reference's
 * 3D acoustic FDTD hot path (Section0 stencil + leapfrog, Section1 source injection).
 *
 * Part 1 is the reference's own operator boundary, byte for byte: a maintainer links this
 * library in place of cuda_optimized.o and main.cpp runs unchanged.
 * Part 2 is an additive device-resident API (plans) needed for grids that cannot be host-staged
 * and for the x-slab multi-GPU decomposition.  Plain pointers and sizes only; no C++/torch types.
 *
 * All functions return 0 on success, otherwise a cudaError_t cast to int (the reference's
 * convention, cuda.cu:280-284; cudaErrorInvalidValue == 1 for bad arguments).
 */
#ifndef FDTD_B200_H
#define FDTD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ Part 1: reference ABI */

/* reference main.cpp:35-45 (dup openacc.cpp:11-22, cuda.cu:21-31).  Only data/size/nbytes are read. */
struct dataobj {
    void *__restrict data;
    int *size;
    unsigned long nbytes;
    unsigned long *npsize;
    unsigned long *dsize;
    int *hsize;
    int *hofs;
    int *oofs;
    void *dmap;
};

/* reference main.cpp:47-50: seconds of the steps time >= time_m + 5. */
struct profiler {
    double section0; /* stencil + leapfrog (includes the fused injection) */
    double section1; /* stand-alone source scatter (0 when every source cell is fused) */
};

/*
 * Replaces Kernel_CUDA_Optimized (reference cuda_optimized.cu:282-287; declared main.cpp:67-72,
 * called through KernelFunc main.cpp:75-80,394-399).  Same 24 arguments, max before min.
 * u: fp32 [3][nxp][nyp][nzp] in/out (all three levels copied back), m: fp32 [nxp][nyp][nzp],
 * src: fp32 [T][src.size[1]], src_coords: fp32 [nsrc][src_coords.size[1]], 4-cell halos.
 * "no sources": p_src_M < p_src_m, or src.size[0]*src.size[1] == 0, or src.data == NULL.
 * timers: section0/section1 are OVERWRITTEN (cuda_optimized.cu:290,469-470) with measured
 * device seconds -- no fake 85/15 split.
 */
int Kernel_CUDA_Optimized(struct dataobj *__restrict m_vec, struct dataobj *__restrict src_vec,
                          struct dataobj *__restrict src_coords_vec, struct dataobj *__restrict u_vec,
                          const int x_M, const int x_m, const int y_M, const int y_m, const int z_M,
                          const int z_m, const float dt, const float h_x, const float h_y,
                          const float h_z, const float o_x, const float o_y, const float o_z,
                          const int p_src_M, const int p_src_m, const int time_M, const int time_m,
                          const int deviceid, const int devicerm, struct profiler *timers);

/* Same operator under its own name, for linking side by side with the reference's cuda_optimized.o. */
int Kernel_B200(struct dataobj *__restrict m_vec, struct dataobj *__restrict src_vec,
                struct dataobj *__restrict src_coords_vec, struct dataobj *__restrict u_vec,
                const int x_M, const int x_m, const int y_M, const int y_m, const int z_M,
                const int z_m, const float dt, const float h_x, const float h_y, const float h_z,
                const float o_x, const float o_y, const float o_z, const int p_src_M,
                const int p_src_m, const int time_M, const int time_m, const int deviceid,
                const int devicerm, struct profiler *timers);

/*
 * The optional hook main.cpp declares weak (main.cpp:84) and calls once per method
 * (main.cpp:271-276).  use_tc is ignored (no tensor cores: the stencil is not a contraction),
 * t_fuse selects the temporal-blocking depth (1 = one step per pass), nfields must be 1.
 */
void FDTD_SetRuntimeConfig(int use_tc, int t_fuse, int nfields);

/* ------------------------------------------------------------------ Part 2: resident plans */

typedef struct fdtd_b200_plan fdtd_b200_plan;

/*
 * One x-slab of the grid resident on one GPU (x = the reference's slowest axis).  A single-GPU
 * run is the slab [0, nx_global).  Local arrays are u[3][nx+8][ny+8][nz+8] and m[nx+8][ny+8][nz+8]:
 * the same 4-cell halo layout as the reference, so planes 2..3 and nx+4..nx+5 are the ghost planes
 * a neighbour slab fills (or the fixed physical halo at the two global ends).
 */
typedef struct {
    int nx, ny, nz;      /* interior extents of this slab */
    int x_offset;        /* global index of this slab's first interior x plane */
    int nx_global;       /* global interior x extent (== nx for one GPU) */
    float dt, h_x, h_y, h_z;
    float o_x, o_y, o_z; /* GLOBAL origin */
    int deviceid;        /* -1: keep the current device */
} fdtd_b200_geometry;

int fdtd_b200_plan_create(const fdtd_b200_geometry *geom, fdtd_b200_plan **out);
int fdtd_b200_plan_destroy(fdtd_b200_plan *plan);
/*
 * The same at space order 4, 6, 8, 10 or 12.  The reference's driver is parameterised by
 * STENCIL_ORDER with HALO == STENCIL_ORDER cells per side (main.cpp:2,27-32) but ships order-4 kernels only; here the
 * arrays are u[3][nx+2*order][..][..], the stencil has radius order/2 with the standard central second-derivative
 * weights (correctly rounded floats of the exact rationals; order 4 = the reference's literals) summed outermost pair
 * first like openacc.cpp:104-106.  Orders above 4 run the one-point-per-thread kernel on a single slab
 * (cudaErrorNotSupported for x-slabs).  Kernel_* reads the order off the padding, or from FDTD_B200_STENCIL_ORDER.
 */
int fdtd_b200_plan_create_order(const fdtd_b200_geometry *geom, int space_order, fdtd_b200_plan **out);

/* Device pointers of the slab's arrays (for wrapping as torch tensors / IPC export). */
float *fdtd_b200_plan_u(fdtd_b200_plan *plan);
float *fdtd_b200_plan_m(fdtd_b200_plan *plan);
/* Device pointer of ring level 0..2 (u[level] of the reference ABI).  The allocation holds one spare level
 * beyond the ABI's three; two-step passes rotate it through the ring, so after a run with t_fuse = 2 ring
 * level r is NOT necessarily at plan_u + r*level_elems -- use this accessor (download does). */
float *fdtd_b200_plan_level(fdtd_b200_plan *plan, int ring_level);
size_t fdtd_b200_plan_level_elems(fdtd_b200_plan *plan); /* (nx+8)(ny+8)(nz+8) */

/* Host <-> device staging of whole arrays (either pointer may be NULL to skip it). */
int fdtd_b200_plan_upload(fdtd_b200_plan *plan, const float *h_u, const float *h_m);
int fdtd_b200_plan_download(fdtd_b200_plan *plan, float *h_u);
/*
 * A window of one ring level, [x0,x1) x [y0,y1) x [z0,z1) in PADDED LOCAL coordinates (interior starts at 4),
 * copied to host as a dense float[x1-x0][y1-y0][z1-z0] (SURVEY 8b "_window"): what a harness needs to compare a
 * 1024^3 / 2048^3 run with the reference on cropped grids around the sources and across slab seams
 * (main.cpp:573-604 compares all three levels) without pulling 13-104 GB.
 */
int fdtd_b200_plan_download_window(fdtd_b200_plan *plan, int ring_level, int x0, int x1, int y0, int y1, int z0,
                                   int z1, float *host);
/*
 * Checksum of a window of one ring level (SURVEY 8b "_checksum"), computed on the device.  bit_sum / pos_sum /
 * nonzero / nonfinite are integer sums (mod 2^64) and therefore independent of the summation order: the sums of
 * the slabs' interior windows equal the single-GPU sums.  pos_sum weights every bit pattern with 1 + its GLOBAL
 * padded linear index ((X + x_offset)*nyp + Y)*nzp + Z, so a misplaced plane changes it.  sum_sq (double) and
 * max_abs skip non-finite cells.
 */
typedef struct {
    unsigned long long bit_sum;
    unsigned long long pos_sum;
    unsigned long long nonzero;   /* cells whose value is not +-0 */
    unsigned long long nonfinite; /* Inf / NaN cells */
    double sum_sq;
    float max_abs;
    int reserved;
} fdtd_b200_checksum;
int fdtd_b200_plan_checksum(fdtd_b200_plan *plan, int ring_level, int x0, int x1, int y0, int y1, int z0, int z1,
                            fdtd_b200_checksum *out);
/* Device-side constant fill of all three levels / of m (the driver's synthetic init, main.cpp:351-352). */
int fdtd_b200_plan_fill(fdtd_b200_plan *plan, float u_value, float m_value);
/* Dense parity field of main.cpp:525-532 generated on the device from the GLOBAL linear index. */
int fdtd_b200_plan_fill_dense(fdtd_b200_plan *plan);

/*
 * Sources (host arrays, copied): src [src_size0][pstride], coords [ncoords][cstride] in GLOBAL
 * physical coordinates; the active range is [p_src_m, p_src_M].  Builds the per-cell scatter
 * table (positions, fractions and weights in IEEE fp32, bit-exact with openacc.cpp:125-134);
 * a slab keeps only the cells it owns.
 */
int fdtd_b200_plan_set_sources(fdtd_b200_plan *plan, const float *src, int src_size0, int pstride,
                               const float *coords, int ncoords, int cstride, int p_src_m,
                               int p_src_M);

/*
 * Receivers (the reference has none -- the operator it was generated from samples them after the
 * injection): rec[time][r] = sum over the in-range trilinear corners, x outermost / z innermost, of
 * ((wx*wy)*wz)*u[time % 3][corner], positions / fractions / bounds exactly as the injection's (openacc.cpp:125-132).
 * coords [nrec][cstride] in GLOBAL physical coordinates (copied).  Every run samples all its steps; the time is
 * reported under section1.  download: host [rows][nrec] of the last run (rows = its steps, also returned), owned[r] = 1
 * where THIS slab sampled receiver r (x-slabs: exactly one slab owns a receiver; the others leave 0).
 */
int fdtd_b200_plan_set_receivers(fdtd_b200_plan *plan, const float *coords, int nrec, int cstride);
int fdtd_b200_plan_download_receivers(fdtd_b200_plan *plan, float *host, int *owned, int *rows);

/*
 * Run time steps time_m..time_M inclusive on this slab (ring phase = time % 3).  The first
 * min(5, T) steps are untimed (openacc.cpp:90-92,148); timers receive device seconds of the rest.
 * Blocking.  For a multi-slab run every rank calls this with peers attached (below).
 */
int fdtd_b200_plan_run(fdtd_b200_plan *plan, int time_m, int time_M, struct profiler *timers);

/*
 * upload + run + download as one pipeline, for host arrays h_u [3][nxp][nyp][nzp] (in/out) and h_m (what
 * Kernel_* does with its caller's buffers): the arrays travel in chunks of x planes, the time loop is skewed
 * along x so a block of planes runs all its steps as soon as its chunk has landed, and finished planes go
 * back while later chunks are still arriving (PCIe is full duplex).  Bit-identical to upload, run, download.
 * Option "stage_planes" = planes per block (-1 = auto: grids of >= 8M points, blocks sized for ~1600 launches;
 * 0 = off).  Returns cudaErrorNotSupported (801) when the three-phase path must be used instead: linked slabs,
 * sources that touch halo cells, fewer than 2 blocks, or (auto) a grid too small to gain.
 */
int fdtd_b200_plan_run_staged(fdtd_b200_plan *plan, float *h_u, const float *h_m, int time_m, int time_M,
                              struct profiler *timers);

/* Number of kernel launches issued by the last run (stencil + scatter + halo kernels). */
long fdtd_b200_plan_last_launches(fdtd_b200_plan *plan);
/* Average stencil-kernel seconds per launch in the timed region of the last run. */
double fdtd_b200_plan_last_kernel_seconds(fdtd_b200_plan *plan);

/*
 * Options: "kernel" 0 = auto, 1 = generic (any extents), 2 = tma (2.5D x-streaming, TMA ring);
 * "exact" 1 = replay the reference's fp32 operation order (0 ulp vs the host build), 0 = contracted;
 * "fuse_inject" 1 = scatter inside the stencil epilogue; "tile_y","tile_z","rows","xchunk";
 * "t_fuse" temporal-blocking depth: 1 = one time step per pass, >= 2 = two time steps per pass (u^{n+1} and
 * u^{n+2} from one read of u^{n-1}, u^n, m; bit-identical to two one-step passes).  Two-step passes need the
 * halo shells of the three levels to be identical and no source corner in a halo cell; otherwise, and for the
 * steps that do not pair up (the untimed/timed boundary, an odd remainder), one-step passes run.
 * get_option also answers "kernel_used", "t_fuse_used", "tile_y_used", "tile_z_used", "rows_used", "xchunk_used",
 * "ncells_fused", "ncells_halo" for the last run.
 */
int fdtd_b200_plan_set_option(fdtd_b200_plan *plan, const char *key, int value);
int fdtd_b200_plan_get_option(fdtd_b200_plan *plan, const char *key, int *value);

/*
 * Depth (1 or 2) this slab could run with its current options, sources and field contents.  Linked slabs must
 * all use the same depth: one process per GPU calls this on every rank, takes the minimum (an all-reduce on
 * the host side) and sets option "t_fuse_agreed" before the run; fdtd_b200_run_slabs does it internally.
 */
int fdtd_b200_plan_probe_fuse(fdtd_b200_plan *plan, int *t_fuse);

/* ---- x-slab neighbours (one process per GPU; handles travel over torch.distributed) */
#define FDTD_B200_IPC_BYTES 160
/* Fill `blob` (FDTD_B200_IPC_BYTES) with this slab's exportable handles. */
int fdtd_b200_plan_ipc_export(fdtd_b200_plan *plan, void *blob);
/* Attach the neighbour on side 0 (x-, lower planes) or 1 (x+) from its exported blob. */
int fdtd_b200_plan_ipc_attach(fdtd_b200_plan *plan, int side, const void *blob);
/* Same-process attachment (several slabs driven by one process, peer access enabled). */
int fdtd_b200_plan_attach_local(fdtd_b200_plan *plan, int side, fdtd_b200_plan *neighbour);

/*
 * Run the same time steps on several slabs driven by ONE process (one plan per device, or several
 * plans on one device for tests), neighbours attached with fdtd_b200_plan_attach_local.  timers
 * receive the slowest slab's device seconds.
 */
int fdtd_b200_run_slabs(fdtd_b200_plan **plans, int nplans, int time_m, int time_M, struct profiler *timers);

/* ---- host-only helpers (no GPU needed) */

/* Source table of openacc.cpp:125-134 for one source: pos[3], frac[3], 8 corner weights
 * w[rx*4+ry*2+rz] = 1e-2f*wx*wy*wz, in_range[8] per openacc.cpp:132. */
int fdtd_b200_source_table(const float coord[3], const float o[3], const float h[3], const int lo[3],
                           const int hi[3], int pos[3], float frac[3], float w[8], int in_range[8]);
/* The scatter table a slab would build from these sources (host only): cells[i] = {X, Y, Z, first, count}
 * in padded local coordinates, the first *ncells_int cells lie inside the Section0 write range; every
 * cell sums contrib_w[first..first+count) * src[t][contrib_p[..]] / m[base_idx[p]] in that order. */
int fdtd_b200_slab_source_cells(const fdtd_b200_geometry *geom, const float *coords, int ncoords, int cstride,
                                int p_src_m, int p_src_M, int max_cells, int *cells, int *ncells_int,
                                int *ncells_all, int max_contribs, int *contrib_p, float *contrib_w,
                                int *ncontribs, long long *base_idx);
/* The table two-step launches use (host only): interior cells plus the cells on the neighbour slabs' two nearest
 * planes, sorted by (X,Y,Z); *halo_global = 1 when some source corner of the GLOBAL grid is a halo cell. */
int fdtd_b200_slab_source_cells2(const fdtd_b200_geometry *geom, const float *coords, int ncoords, int cstride,
                                 int p_src_m, int p_src_M, int max_cells, int *cells, int *ncells,
                                 int max_contribs, int *contrib_p, float *contrib_w, int *ncontribs,
                                 int *halo_global);
/* The driver's input generators (main.cpp:290-298 and 301-325), fp32. */
void fdtd_b200_fill_ricker(float *src, int T, int S, float dt);
void fdtd_b200_fill_source_coords(float *coords, int S, int nx, int ny, int nz, float h_x, float h_y,
                                  float h_z);
/* Append one row in the reference's benchmark.csv schema (main.cpp:201-249, 24 columns). */
int fdtd_b200_write_benchmark_csv(const char *filename, const char *method, double total_s,
                                  double total_std, double s0_s, double s0_std, double s1_s,
                                  double s1_std, double device_s, double device_std,
                                  double overhead_s, double overhead_std, double gflops,
                                  double gflops_std, double gbps, double gbps_std, double peak_fp32_gf,
                                  double peak_bw_gbs, double ai, int nx, int ny, int nz, int timesteps,
                                  int nsrc, int stencil_order);
const char *fdtd_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FDTD_B200_H */
