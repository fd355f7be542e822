// stencil_tma.cu -- Section0 (+ fused Section1) as a 2.5D x-streaming kernel for sm_100a.
//
// Replaces reference cuda_optimized.cu:63-238 (stencil_update_h100_scalar_pipelined_kernel).
// Design (B200-first, not a port):
//   * a CTA owns one (y,z) tile of TY x TZ points and walks a chunk of x planes (x = the
//     reference's slowest axis; one x plane is nyp*nzp contiguous floats);
//   * one producer thread feeds an mbarrier-pipelined shared-memory ring with TMA
//     (cp.async.bulk.tensor): per stage the u[t0] plane tile WITH its radius-2 halo
//     ((TY+4) x (TZ+8) box, z start 16-byte aligned) and, two planes behind it, the centre
//     tiles of u[t1] and m (TY x TZ boxes).  No consumer thread ever issues a global load and
//     there is no __syncthreads in the steady state: consumers wait on full[slot] and release
//     with one mbarrier arrive per warp;
//   * every consumer thread owns a float4 of z at one y and keeps its own column's
//     x-2..x+2 values in a register queue, so only the centre plane is read from shared
//     memory for the y/z neighbours (4 LDS.128 + 2 LDS.64 per 4 points);
//   * results leave through 128-bit coalesced stores; the thread that owns a source cell
//     adds the source terms in p_src order before the store (atomics-free fused Section1).
// Stage s of a chunk starting at padded plane Xa carries u[t0] plane Xa-2+s and, for s >= 4,
// u[t1]/m plane Xa+s-4.  Iteration j (output plane Xa+j) reads its own column from stage j+4,
// the neighbours from stage j+2, and then releases stage j+2 (halo slot (j+2)%S0 and centre
// slot j%(S0-2) become free together, so one empty barrier per halo slot suffices).
#include "fdtd_arith.cuh"
#include "fdtd_kernels.cuh"
#include "tma_ptx.cuh"

#include <cudaTypedefs.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace fdtd {

struct TmaArgs {
    alignas(64) CUtensorMap map_halo;
    alignas(64) CUtensorMap map_ctr;
    alignas(64) CUtensorMap map_m;
    alignas(64) CUtensorMap map_halo_peer[2];  // pull mode: the neighbours' u
    StepArgs s;
    int tiles_z, tiles_y, xchunk;
    int edge;  // > 0: the first and last chunk are only `edge` planes long (slabs with neighbours)
};

// ---------------------------------------------------------------------------- geometry of one variant
// S0_ = halo-plane ring depth; the plane loop is unrolled by S0_ and the register queue has period 5,
// so S0_ must be a multiple of 5 (5: two stages of prefetch, 10: seven).
template <int TY, int TZ, int RY, int S0_>
struct TileShape {
    static constexpr int S0 = S0_;
    static_assert(S0_ % 5 == 0, "ring depth must be a multiple of the queue period");
    static constexpr int S1 = S0 - 2;               // centre-ring slots
    static constexpr int ZQ = TZ / 4;               // float4 columns per row
    static constexpr int NC = (TY / RY) * ZQ;       // consumer threads: RY rows x 4 z each
    static constexpr int NCW = NC / 32;             // consumer warps
    static constexpr int NT = NC + 32;              // + one producer warp
    static constexpr int HP = TZ + 8;               // halo-tile pitch (floats)
    static constexpr int HROWS = TY + 4;
    static constexpr int HBYTES = HROWS * HP * 4;   // bytes one halo box delivers
    static constexpr int HSLOT = (HBYTES + 127) / 128 * 128;
    static constexpr int CBYTES = TY * TZ * 4;      // bytes one centre box delivers
    static constexpr int SMEM = S0 * HSLOT + 2 * S1 * CBYTES + 2 * S0 * 8;
    static_assert(TZ % 4 == 0 && TY % RY == 0 && NC % 32 == 0, "tile must give whole warps of float4 columns");
    static_assert(NT <= 1024, "too many threads");
    static_assert(CBYTES % 128 == 0, "centre slots must stay 128-byte aligned");
};

template <int I>
__device__ __forceinline__ float f4get(const float4 &v)
{
    return I == 0 ? v.x : I == 1 ? v.y : I == 2 ? v.z : v.w;
}

template <int TY, int TZ, int RY, int NS, bool EXACT, int MINB>
__global__ void __launch_bounds__(TileShape<TY, TZ, RY, NS>::NT, MINB)
    stencil_tma_kernel(const __grid_constant__ TmaArgs a)
{
    using T = TileShape<TY, TZ, RY, NS>;
    constexpr int S0 = T::S0, S1 = T::S1;
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sH = reinterpret_cast<float *>(smem);
    float *sU1 = reinterpret_cast<float *>(smem + S0 * T::HSLOT);
    float *sM = sU1 + S1 * TY * TZ;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S0 * T::HSLOT + 2 * S1 * T::CBYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S0);

    const Grid &g = a.s.g;
    const int tz = blockIdx.x % a.tiles_z, ty = blockIdx.x / a.tiles_z;
    // chunk order: the two chunks that hold the slab's boundary planes are dispatched first, so the
    // neighbours' ghost planes are written (and their flags raised) early in the step
    const int nch = gridDim.y, by = blockIdx.y;
    const int chunk = by == 0 ? 0 : (by == 1 ? nch - 1 : by - 1);
    int Xa, Xb;
    if (a.edge == 0) {
        Xa = g.X0 + chunk * a.xchunk;
        Xb = min(g.X1, Xa + a.xchunk);
    } else if (chunk == 0) {       // short boundary chunks: the neighbours get their ghost planes (and flags)
        Xa = g.X0;                  // within the first wave of CTAs instead of after a full-length chunk
        Xb = g.X0 + a.edge;
    } else if (chunk == nch - 1) {
        Xa = g.X1 - a.edge;
        Xb = g.X1;
    } else {
        Xa = g.X0 + a.edge + (chunk - 1) * a.xchunk;
        Xb = min(g.X1 - a.edge, Xa + a.xchunk);
    }
    const int np = Xb - Xa;         // output planes of this CTA (>= 1 by construction)
    const SlabLink &lk = a.s.link;
    // the chunk with the slab's UPPER boundary streams downwards: its boundary planes leave first (see stencil_tb2.cu)
    const bool rev = lk.peer_u[1] != nullptr && nch > 1 && chunk == nch - 1;
    const int Yt = g.Y0 + ty * TY;  // padded origin of the tile
    const int Zt = g.Z0 + tz * TZ;

    // Does any source cell of this chunk lie inside this TILE?  (Asking per chunk only sent every tile of the source's chunks --
    // a quarter of the CTAs at 512^3 -- through two dependent global loads per plane.)
    __shared__ int s_tile_has_src;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S0; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, T::NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_tile_has_src = 0;
    }
    __syncthreads();
    if (a.s.sv.ncells > 0) {
        const int c0 = a.s.sv.plane_off[Xa], c1 = a.s.sv.plane_off[Xb];
        for (int q = c0 + (int)threadIdx.x; q < c1; q += T::NT) {
            const SourceCell cell = a.s.sv.cells[q];
            if (cell.Y >= Yt && cell.Y < Yt + TY && cell.Z >= Zt && cell.Z < Zt + TZ) s_tile_has_src = 1;
        }
        __syncthreads();
    }
    const bool chunk_has_src = s_tile_has_src != 0;

    if (threadIdx.x >= T::NC) {
        // ------------------------------------------------------------------ producer (one thread)
        // per-tile flags: the whole producer warp polls the 3 x 3 tiles around its own on the sides whose ghost planes
        // this chunk reads (and whose neighbour it will write to: same tiles, same condition)
        const bool tiled = lk.tile_mode && lk.wait;
        if (tiled) {
            if (lk.peer_u[0] && Xa - 2 < g.X0) wait_tiles(lk.my_tile[0], ty, tz, a.tiles_y, a.tiles_z, lk.epoch - 1, lk.err);
            if (lk.peer_u[1] && Xb + 2 > g.X1) wait_tiles(lk.my_tile[1], ty, tz, a.tiles_y, a.tiles_z, lk.epoch - 1, lk.err);
        }
        if (threadIdx.x == T::NC) {
            const int nst = np + 4;
            int slot = 0, use = 0, cslot = 0;
            bool waited[2] = {tiled || !(lk.wait && lk.peer_u[0]), tiled || !(lk.wait && lk.peer_u[1])};
            for (int s = 0; s < nst; ++s) {
                const int Xp = rev ? Xb + 1 - s : Xa - 2 + s;  // u[t0] plane of this stage: a ghost plane outside [X0, X1)
                const int Xc = rev ? Xb + 3 - s : Xa + s - 4;  // u[t1] / m plane of this stage (output plane of iteration s-4)
                const int side = Xp < g.X0 ? 0 : (Xp >= g.X1 ? 1 : -1);
                if (side >= 0 && !waited[side]) {
                    wait_flag(lk.my_flag[side], lk.epoch - 1, lk.err);
                    waited[side] = true;
                }
                if (use > 0) mbar_wait(empty0 + 8 * slot, (use - 1) & 1);
                const uint32_t bar = full0 + 8 * slot;
                const bool ctr = s >= 4;
                mbar_expect_tx(bar, T::HBYTES + (ctr ? 2 * T::CBYTES : 0));
                if (lk.pull && side >= 0 && lk.peer_u[side])  // a neighbour's plane, read where it lies (same level placement on every slab)
                    tma_load_4d(smem_u32(sH) + slot * T::HSLOT, &a.map_halo_peer[side], bar, Zt - 4, Yt - 2,
                                lk.peer_edge[side] + Xp - (side == 0 ? g.X0 : g.X1), a.s.t0);
                else
                    tma_load_4d(smem_u32(sH) + slot * T::HSLOT, &a.map_halo, bar, Zt - 4, Yt - 2, Xp, a.s.t0);
                if (ctr) {
                    tma_load_4d(smem_u32(sU1) + cslot * T::CBYTES, &a.map_ctr, bar, Zt, Yt, Xc, a.s.t1);
                    tma_load_3d(smem_u32(sM) + cslot * T::CBYTES, &a.map_m, bar, Zt, Yt, Xc);
                    if (++cslot == S1) cslot = 0;
                }
                if (++slot == S0) {
                    slot = 0;
                    ++use;
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int zq = threadIdx.x % T::ZQ, yr = threadIdx.x / T::ZQ;
    const int lane = threadIdx.x & 31;
    const int Y = Yt + yr * RY, Z = Zt + 4 * zq;       // first of this thread's RY rows
    const bool z_ok = Z < g.Z1;
    const int own = (yr * RY + 2) * T::HP + 4 + 4 * zq;  // own column inside a halo slot (floats)
    const int ctr = yr * RY * TZ + 4 * zq;               // own column inside a centre slot
    constexpr int HSLOT_F = T::HSLOT / 4;

    const SourceView &sv = a.s.sv;

    // Register queue: q[s % 5][r] = this thread's float4 of u[t0] plane (stage s), row r.  All indices are
    // compile-time constants because the plane loop is unrolled by the ring period.
    float4 q[5][RY];
#pragma unroll
    for (int s = 0; s < 4; ++s) {  // prologue: planes Xa-2 .. Xa+1
        mbar_wait(full0 + 8 * s, 0);
#pragma unroll
        for (int r = 0; r < RY; ++r) q[s][r] = lds128(sH + s * HSLOT_F + own + r * T::HP);
        if (s == 1) {  // stages 0 and 1 are never a centre plane: release them now
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty0 + 0);
                mbar_arrive(empty0 + 8);
            }
        }
    }

    // boundary planes also go to the neighbours' ghost planes (peer stores over NVLink)
    const bool cta_lo = lk.peer_u[0] != nullptr && Xa < g.X0 + lk.depth;
    const bool cta_hi = lk.peer_u[1] != nullptr && Xb > g.X1 - lk.depth;
    const long long row0 = (long long)Y * g.nzp + Z;

    int ms = 0;  // centre-ring slot j % S1
    float *__restrict__ out = a.s.u + (long long)a.s.t2 * g.lvl + ((long long)(rev ? Xb - 1 : Xa) * g.nyp + Y) * g.nzp + Z;
    const long long plane = (long long)g.nyp * g.nzp;
    const long long out_step = rev ? -plane : plane;

    for (int j0 = 0; j0 < np; j0 += S0) {
        const uint32_t par = (uint32_t)(j0 / S0) & 1u;
#pragma unroll
        for (int k = 0; k < S0; ++k) {
            const int j = j0 + k;
            if (j >= np) break;
            const int fsl = (k + 4) % S0, csl = (k + 2) % S0;  // front (stage j+4) and centre (stage j+2) slots
            mbar_wait(full0 + 8 * fsl, par ^ (uint32_t)((k + 4) / S0));
            float4 u1v[RY], mv[RY];
            float2 zl[RY], zr[RY];
            const float *F = sH + fsl * HSLOT_F + own;
            const float *P = sH + csl * HSLOT_F + own;
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                q[(k + 4) % 5][r] = lds128(F + r * T::HP);
                u1v[r] = lds128(sU1 + ms * (TY * TZ) + ctr + r * TZ);
                mv[r] = lds128(sM + ms * (TY * TZ) + ctr + r * TZ);
                zl[r] = lds64(P + r * T::HP - 2);
                zr[r] = lds64(P + r * T::HP + 4);
            }
            // the y column of this thread on the centre plane: 2 rows below, own rows (from the queue), 2 above
            float4 col[RY + 4];
            col[0] = lds128(P - 2 * T::HP);
            col[1] = lds128(P - T::HP);
            col[RY + 2] = lds128(P + RY * T::HP);
            col[RY + 3] = lds128(P + (RY + 1) * T::HP);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * csl);  // stage j+2 (and centre slot j % S1) free
#pragma unroll
            for (int r = 0; r < RY; ++r) col[r + 2] = q[(k + 2) % 5][r];

            float4 o[RY];
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                const float4 c = col[r + 2];
                const float4 xm2 = q[k % 5][r], xm1 = q[(k + 1) % 5][r], xp1 = q[(k + 3) % 5][r], xp2 = q[(k + 4) % 5][r];
                const float4 ym2 = col[r], ym1 = col[r + 1], yp1 = col[r + 3], yp2 = col[r + 4];
                o[r] = column4<EXACT>(c, xm2, xm1, xp1, xp2, ym2, ym1, yp1, yp2, zl[r], zr[r], u1v[r], mv[r], a.s.k);
            }

            if (chunk_has_src) {  // fused Section1: rare path, only chunks that contain a source cell
                const int X = rev ? Xb - 1 - j : Xa + j;
                const int c0 = sv.plane_off[X], c1 = sv.plane_off[X + 1];
                for (int i = c0; i < c1; ++i) {
                    const SourceCell cell = sv.cells[i];
                    const int dyc = cell.Y - Y, dzc = cell.Z - Z;
                    if (dyc >= 0 && dyc < RY && dzc >= 0 && dzc < 4) {
                        // extract / re-insert with selects so o[] stays in registers
                        float v = 0.f;
#pragma unroll
                        for (int r = 0; r < RY; ++r) {
                            const float e = dzc == 0 ? o[r].x : dzc == 1 ? o[r].y : dzc == 2 ? o[r].z : o[r].w;
                            v = dyc == r ? e : v;
                        }
                        v = apply_cell(v, cell, sv);
#pragma unroll
                        for (int r = 0; r < RY; ++r) {
                            const bool hit = dyc == r;
                            o[r].x = (hit && dzc == 0) ? v : o[r].x;
                            o[r].y = (hit && dzc == 1) ? v : o[r].y;
                            o[r].z = (hit && dzc == 2) ? v : o[r].z;
                            o[r].w = (hit && dzc == 3) ? v : o[r].w;
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RY; ++r)
                if (z_ok && Y + r < g.Y1) *reinterpret_cast<float4 *>(out + (long long)r * g.nzp) = o[r];
            out += out_step;
            if ((cta_lo || cta_hi) && !lk.pull) {
                const int X = rev ? Xb - 1 - j : Xa + j;
                if (cta_lo && X < g.X0 + lk.depth) {
                    float *dst = lk.peer_u[0] + a.s.t2 * lk.peer_lvl[0] + (long long)(lk.peer_edge[0] + X - g.X0) * plane + row0;
#pragma unroll
                    for (int r = 0; r < RY; ++r)
                        if (z_ok && Y + r < g.Y1) *reinterpret_cast<float4 *>(dst + (long long)r * g.nzp) = o[r];
                }
                if (cta_hi && X >= g.X1 - lk.depth) {
                    float *dst = lk.peer_u[1] + a.s.t2 * lk.peer_lvl[1] + (long long)(lk.peer_edge[1] + X - g.X1) * plane + row0;
#pragma unroll
                    for (int r = 0; r < RY; ++r)
                        if (z_ok && Y + r < g.Y1) *reinterpret_cast<float4 *>(dst + (long long)r * g.nzp) = o[r];
                }
            }
            if (++ms == S1) ms = 0;
        }
    }

    if (cta_lo || cta_hi) {
        // every consumer thread of this CTA has issued its peer stores: count the CTA, and let the last
        // CTA of a boundary publish the step's epoch in the neighbour's flag
        asm volatile("bar.sync 1, %0;" ::"r"(T::NC) : "memory");
        if (threadIdx.x == 0) {
            __threadfence_system();
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!(side == 0 ? cta_lo : cta_hi)) continue;
                raise_flag_fenced(lk.peer_tile[side] + blockIdx.x, lk.epoch);  // this tile's boundary is done (blockIdx.x = ty*tiles_z + tz)
                const int done = atomicAdd(lk.counter + side, 1);
                if (done == lk.expect[side] - 1) {
                    atomicExch(lk.counter + side, 0);
                    __threadfence_system();
                    raise_flag_fenced(lk.peer_flag[side], lk.epoch);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------- host side
typedef void (*TmaKernelFn)(const TmaArgs);
struct Variant {
    int ty, tz, rows, stages;
    bool exact;
    TmaKernelFn fn;
    int nt;
    size_t smem;
    int minb;
};

#define FDTD_V1(TY_, TZ_, RY_, NS_, EX_, MINB_)                                                             \
    {TY_, TZ_, RY_, NS_, EX_, stencil_tma_kernel<TY_, TZ_, RY_, NS_, EX_, MINB_>,                           \
     TileShape<TY_, TZ_, RY_, NS_>::NT, (size_t)TileShape<TY_, TZ_, RY_, NS_>::SMEM, MINB_}
#define FDTD_VARIANT(TY_, TZ_, RY_, MINB_) FDTD_V1(TY_, TZ_, RY_, 5, true, MINB_), FDTD_V1(TY_, TZ_, RY_, 5, false, MINB_)

static const Variant g_variants[] = {
    // (TY, TZ, rows per thread, min CTAs/SM), ring depth 5, both arithmetic modes.  Auto selection takes
    // the first match, so the preferred row count of a tile comes first.
    FDTD_VARIANT(8, 64, 2, 4),   FDTD_VARIANT(8, 64, 1, 4),   FDTD_VARIANT(8, 128, 1, 2),
    FDTD_VARIANT(8, 128, 2, 3),  FDTD_VARIANT(16, 64, 2, 3),  FDTD_VARIANT(16, 64, 1, 2),
    FDTD_VARIANT(16, 128, 2, 2), FDTD_VARIANT(16, 128, 1, 1), FDTD_VARIANT(32, 64, 2, 2),
    FDTD_VARIANT(32, 64, 1, 1),  FDTD_VARIANT(16, 32, 2, 4),  FDTD_VARIANT(16, 32, 1, 4),
    FDTD_VARIANT(8, 32, 1, 6),   FDTD_VARIANT(32, 32, 2, 3),  FDTD_VARIANT(32, 128, 2, 1),
    FDTD_VARIANT(32, 64, 4, 1),  FDTD_VARIANT(32, 128, 4, 1), FDTD_VARIANT(14, 64, 1, 2),
    FDTD_VARIANT(14, 128, 1, 1), FDTD_VARIANT(28, 64, 2, 2),
    // deeper prefetch (ring depth 10), contracted arithmetic only (the exact loop body is too large to unroll x10)
    FDTD_V1(8, 64, 2, 10, false, 3),  FDTD_V1(8, 128, 2, 10, false, 2), FDTD_V1(16, 64, 2, 10, false, 2),
    FDTD_V1(8, 64, 1, 10, false, 3),  FDTD_V1(16, 128, 2, 10, false, 1),
};
static const int g_nvariants = (int)(sizeof(g_variants) / sizeof(g_variants[0]));

int slab_edge_planes(int nx, int xchunk, int tiles, int slots)
{
    (void)xchunk; (void)tiles; (void)slots;
    static const int forced = [] {
        const char *v = getenv("FDTD_B200_SLAB_EDGE");
        return (v && *v) ? atoi(v) : 0;
    }();
    int e = forced > 0 ? forced : kSlabEdgePlanes;
    if (forced < 0) e = nx / 2;            // -1: two chunks of half a slab each
    e = e < 4 ? 4 : e;
    return e > nx / 2 ? nx / 2 : e;
}

bool tma_supported(const Grid &g)
{
    return (g.nzp % 4 == 0) && (g.Z0 % 4 == 0) && ((g.Z1 - g.Z0) % 4 == 0) && (g.Z1 > g.Z0) && (g.Y1 > g.Y0);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

int encode_tensor_map(CUtensorMap *map, const float *base, int rank, const cuuint64_t *dims, const cuuint32_t *box)
{
    PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t strides[3];
    cuuint64_t acc = sizeof(float);
    for (int i = 0; i < rank - 1; ++i) {
        acc *= dims[i];
        strides[i] = acc;  // bytes between consecutive indices of dimension i+1
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

int tma_plan_build(TmaPlan &p, float *u, const float *m, const Grid &g, const TmaConfig &cfg, bool exact,
                   int sm_count, const SlabLink *link)
{
    p.valid = false;
    if (!tma_supported(g)) return (int)cudaErrorInvalidValue;
    const int ny = g.Y1 - g.Y0, nz = g.Z1 - g.Z0, nx = g.X1 - g.X0;

    // ---- tile choice: explicit, or the measured optimum of the 512^3 sweep (profiles/r01_sweep512.txt):
    // small y tiles keep many CTAs per SM in flight; the exact path is issue-bound and wants one row per
    // thread (92 registers), the contracted path is bandwidth-bound and wants two rows of a 128-wide tile.
    int ty = cfg.ty, tz = cfg.tz, rows = cfg.rows;
    if (tz <= 0) tz = exact ? (nz >= 64 ? 64 : 32) : (nz >= 128 ? 128 : (nz >= 64 ? 64 : 32));
    if (ty <= 0) ty = 8;
    if (rows <= 0) rows = (exact || tz == 32) ? 1 : 2;
    int vi = -1;
    for (int pass = 0; pass < 2 && vi < 0; ++pass)  // pass 1: any row count, if the caller did not ask for one
        for (int i = 0; i < g_nvariants && vi < 0; ++i)
            if (g_variants[i].ty == ty && g_variants[i].tz == tz && g_variants[i].exact == exact &&
                (g_variants[i].rows == rows || (pass == 1 && cfg.rows <= 0)) &&
                (cfg.stages <= 0 ? g_variants[i].stages == 5 : g_variants[i].stages == cfg.stages))
                vi = i;
    if (vi < 0) return (int)cudaErrorInvalidValue;
    const Variant &v = g_variants[vi];

    cuuint64_t dims_u[4] = {(cuuint64_t)g.nzp, (cuuint64_t)g.nyp, (cuuint64_t)g.nxp, (cuuint64_t)FDTD_LEVELS};
    cuuint32_t box_h[4] = {(cuuint32_t)(tz + 8), (cuuint32_t)(ty + 4), 1, 1};
    cuuint32_t box_c[4] = {(cuuint32_t)tz, (cuuint32_t)ty, 1, 1};
    int rc;
    if ((rc = encode_tensor_map(&p.map_halo, u, 4, dims_u, box_h))) return rc;
    if ((rc = encode_tensor_map(&p.map_ctr, u, 4, dims_u, box_c))) return rc;
    if ((rc = encode_tensor_map(&p.map_m, m, 3, dims_u, box_c))) return rc;
    for (int side = 0; side < 2; ++side) {
        p.map_halo_peer[side] = p.map_halo;
        if (!link || !link->peer_u[side]) continue;
        cuuint64_t dims_p[4] = {(cuuint64_t)g.nzp, (cuuint64_t)g.nyp, (cuuint64_t)link->peer_nxp[side], (cuuint64_t)FDTD_LEVELS};
        if ((rc = encode_tensor_map(&p.map_halo_peer[side], link->peer_u[side], 4, dims_p, box_h))) return rc;
    }

    cudaError_t e = cudaFuncSetAttribute((const void *)v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)v.fn, v.nt, v.smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return (int)cudaErrorInvalidConfiguration;

    // ---- x chunking.  CTAs = tiles x chunks are dispatched in waves of `slots` = SMs x resident CTAs.
    // Several waves let the hardware scheduler even out the SMs (one wave leaves the issue-bound exact path
    // at 85% of roofline, ~7 reach 97%), but a partially filled last wave idles the machine and every chunk
    // re-reads 4 pipeline-prologue planes of u[t0] (a quarter of the traffic).  Score each chunk count by
    // wave efficiency / prologue overhead and keep the best.
    const int tiles = ((ny + ty - 1) / ty) * ((nz + tz - 1) / tz);
    int xchunk = cfg.xchunk;
    if (xchunk <= 0) {
        const double slots = (double)sm_count * occ;
        double best = -1.0;
        for (int nch = 1; nch <= nx; ++nch) {
            const int xc = (nx + nch - 1) / nch;
            if (xc < 8 && nch > 1) break;
            if ((nx + xc - 1) / xc != nch) continue;  // not a distinct chunking
            const double waves = tiles * (double)nch / slots;
            const double full = ceil(waves);
            double eff = waves / full;          // last-wave quantisation
            eff *= full / (full + 0.25);        // ramp-up / drain of the launch, amortised over the waves
            eff *= xc / (xc + 3.0);             // pipeline prologue of every chunk (4 extra u[t0] planes + fill latency)
            if (eff > best) {
                best = eff;
                xchunk = xc;
            }
        }
    }
    p.ty = ty;
    p.tz = tz;
    p.rows = v.rows;
    p.stages = v.stages;
    p.xchunk = xchunk;
    p.variant = vi;
    p.smem_bytes = v.smem;
    p.valid = true;
    return 0;
}

int launch_stencil_tma(const TmaPlan &p, const StepArgs &a, bool exact, cudaStream_t stream)
{
    if (!p.valid) return (int)cudaErrorInvalidValue;
    const Variant &v = g_variants[p.variant];
    if (v.exact != exact) return (int)cudaErrorInvalidValue;
    const int ny = a.g.Y1 - a.g.Y0, nz = a.g.Z1 - a.g.Z0, nx = a.g.X1 - a.g.X0;
    if (nx <= 0) return 0;
    TmaArgs args;
    args.map_halo = p.map_halo;
    args.map_ctr = p.map_ctr;
    args.map_m = p.map_m;
    args.map_halo_peer[0] = p.map_halo_peer[0];
    args.map_halo_peer[1] = p.map_halo_peer[1];
    args.s = a;
    args.tiles_z = (nz + p.tz - 1) / p.tz;
    args.tiles_y = (ny + p.ty - 1) / p.ty;
    args.xchunk = p.xchunk;
    args.edge = 0;
    int nchunks = (nx + p.xchunk - 1) / p.xchunk;
    const bool linked = a.link.peer_u[0] != nullptr || a.link.peer_u[1] != nullptr;
    if (linked) {
        // exactly ONE chunk per side may hold boundary planes (its CTA raises the tile's flag when it is done): every
        // chunk at least 4 planes long, else fewer and longer chunks
        while (nchunks > 1 && (args.xchunk < 4 || nx - (nchunks - 1) * args.xchunk < 4)) {
            --nchunks;
            args.xchunk = (nx + nchunks - 1) / nchunks;
        }
        nchunks = (nx + args.xchunk - 1) / args.xchunk;
    }
    if (linked && nx >= 4 * kSlabEdgePlanes && !a.link.tile_mode) {  // short boundary chunks + the usual chunks in between
        args.edge = slab_edge_planes(nx, p.xchunk, args.tiles_z * args.tiles_y, 0);
        args.xchunk = p.xchunk;
        nchunks = 2 + (nx - 2 * args.edge + p.xchunk - 1) / p.xchunk;
    }
    dim3 grid(args.tiles_z * args.tiles_y, nchunks, 1);
    if (grid.y > 65535) return (int)cudaErrorInvalidValue;
    // CTAs whose chunk holds one of the two lowest / two highest planes of the slab
    const int tiles = (int)grid.x;
    if (args.edge) {
        args.s.link.expect[0] = args.s.link.expect[1] = tiles;
    } else {
        args.s.link.expect[0] = args.s.link.expect[1] = tiles;  // one chunk per side (see above)
    }
    v.fn<<<grid, v.nt, v.smem, stream>>>(args);
    return (int)cudaGetLastError();
}

}  // namespace fdtd
