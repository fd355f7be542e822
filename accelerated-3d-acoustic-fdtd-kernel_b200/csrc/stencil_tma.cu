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

#include <cudaTypedefs.h>
#include <stdio.h>

namespace fdtd {

struct TmaArgs {
    alignas(64) CUtensorMap map_halo;
    alignas(64) CUtensorMap map_ctr;
    alignas(64) CUtensorMap map_m;
    StepArgs s;
    int tiles_z, tiles_y, xchunk;
};

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float2 lds64(const float *p) { return *reinterpret_cast<const float2 *>(p); }

// ---------------------------------------------------------------------------- geometry of one variant
template <int TY, int TZ, int S0>
struct TileShape {
    static constexpr int ZQ = TZ / 4;               // float4 columns per row
    static constexpr int NC = TY * ZQ;              // consumer threads
    static constexpr int NCW = NC / 32;             // consumer warps
    static constexpr int NT = NC + 32;              // + one producer warp
    static constexpr int S1 = S0 - 2;               // centre-ring slots
    static constexpr int HP = TZ + 8;               // halo-tile pitch (floats)
    static constexpr int HROWS = TY + 4;
    static constexpr int HBYTES = HROWS * HP * 4;   // bytes one halo box delivers
    static constexpr int HSLOT = (HBYTES + 127) / 128 * 128;
    static constexpr int CBYTES = TY * TZ * 4;      // bytes one centre box delivers
    static constexpr int SMEM = S0 * HSLOT + 2 * S1 * CBYTES + 2 * S0 * 8;
    static_assert(TZ % 4 == 0 && NC % 32 == 0, "tile must give whole warps of float4 columns");
    static_assert(S0 >= 5, "ring must hold stages j+2..j+4 plus prefetch");
    static_assert(CBYTES % 128 == 0, "centre slots must stay 128-byte aligned");
};

template <bool EXACT>
__device__ __forceinline__ float point(float c, float r_xm2, float r_xm1, float r_xp1, float r_xp2, float ym2,
                                       float ym1, float yp1, float yp2, float zm2, float zm1, float zp1, float zp2,
                                       float u1, float m, const Coef &k)
{
    const float r5 = EXACT ? __fmul_rn(FDTD_C0, c) : FDTD_C0 * c;
    const float dx = axis_term<EXACT>(r5, r_xm2, r_xm1, r_xp1, r_xp2);
    const float dy = axis_term<EXACT>(r5, ym2, ym1, yp1, yp2);
    const float dz = axis_term<EXACT>(r5, zm2, zm1, zp1, zp2);
    return leapfrog<EXACT>(c, dx, dy, dz, u1, m, k);
}

template <int TY, int TZ, int S0, bool EXACT, int MINB>
__global__ void __launch_bounds__(TileShape<TY, TZ, S0>::NT, MINB)
    stencil_tma_kernel(const __grid_constant__ TmaArgs a)
{
    using T = TileShape<TY, TZ, S0>;
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sH = reinterpret_cast<float *>(smem);
    float *sU1 = reinterpret_cast<float *>(smem + S0 * T::HSLOT);
    float *sM = sU1 + T::S1 * TY * TZ;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S0 * T::HSLOT + 2 * T::S1 * T::CBYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S0);

    const Grid &g = a.s.g;
    const int tz = blockIdx.x % a.tiles_z, ty = blockIdx.x / a.tiles_z;
    const int Xa = g.X0 + blockIdx.y * a.xchunk;
    const int Xb = min(g.X1, Xa + a.xchunk);
    const int np = Xb - Xa;        // output planes of this CTA (>= 1 by construction)
    const int Yt = g.Y0 + ty * TY;  // padded origin of the tile
    const int Zt = g.Z0 + tz * TZ;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S0; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, T::NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= T::NC) {
        // ------------------------------------------------------------------ producer (one thread)
        if (threadIdx.x == T::NC) {
            const int nst = np + 4;
            int slot = 0, use = 0, cslot = 0;
            for (int s = 0; s < nst; ++s) {
                if (use > 0) mbar_wait(empty0 + 8 * slot, (use - 1) & 1);
                const uint32_t bar = full0 + 8 * slot;
                const bool ctr = s >= 4;
                mbar_expect_tx(bar, T::HBYTES + (ctr ? 2 * T::CBYTES : 0));
                tma_load_4d(smem_u32(sH) + slot * T::HSLOT, &a.map_halo, bar, Zt - 4, Yt - 2, Xa - 2 + s, a.s.t0);
                if (ctr) {
                    tma_load_4d(smem_u32(sU1) + cslot * T::CBYTES, &a.map_ctr, bar, Zt, Yt, Xa + s - 4, a.s.t1);
                    tma_load_3d(smem_u32(sM) + cslot * T::CBYTES, &a.map_m, bar, Zt, Yt, Xa + s - 4);
                    if (++cslot == T::S1) cslot = 0;
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
    const int zq = threadIdx.x % T::ZQ, yy = threadIdx.x / T::ZQ;
    const int lane = threadIdx.x & 31;
    const int Y = Yt + yy, Z = Zt + 4 * zq;
    const bool store_ok = (Y < g.Y1) && (Z < g.Z1);
    const int own = (yy + 2) * T::HP + 4 + 4 * zq;  // own column inside a halo slot (floats)
    const int ctr = yy * TZ + 4 * zq;               // own column inside a centre slot
    constexpr int HSLOT_F = T::HSLOT / 4;

    const SourceView &sv = a.s.sv;
    bool chunk_has_src = false;
    if (sv.ncells > 0) chunk_has_src = (sv.plane_off[Xb] - sv.plane_off[Xa]) > 0;

    // prologue: own-column values of planes Xa-2 .. Xa+1 (stages 0..3) into the register queue
    float4 qm2, qm1, qc, qp1;
    mbar_wait(full0 + 0, 0);
    qm2 = lds128(sH + 0 * HSLOT_F + own);
    mbar_wait(full0 + 8, 0);
    qm1 = lds128(sH + 1 * HSLOT_F + own);
    __syncwarp();
    if (lane == 0) {  // stages 0 and 1 are never a centre plane: release them now
        mbar_arrive(empty0 + 0);
        mbar_arrive(empty0 + 8);
    }
    mbar_wait(full0 + 16, 0);
    qc = lds128(sH + 2 * HSLOT_F + own);
    mbar_wait(full0 + 24, 0);
    qp1 = lds128(sH + 3 * HSLOT_F + own);

    int fs = 4 % S0, fpar = (4 / S0) & 1;  // front stage j+4: slot and phase parity
    int cs = 2;                            // centre stage j+2: slot
    int ms = 0;                            // centre-ring slot j % S1
    float *__restrict__ out = a.s.u + (long long)a.s.t2 * g.lvl + ((long long)Xa * g.nyp + Y) * g.nzp + Z;
    const long long plane = (long long)g.nyp * g.nzp;

    for (int j = 0; j < np; ++j) {
        mbar_wait(full0 + 8 * fs, fpar);
        const float4 qp2 = lds128(sH + fs * HSLOT_F + own);
        const float4 u1v = lds128(sU1 + ms * (TY * TZ) + ctr);
        const float4 mv = lds128(sM + ms * (TY * TZ) + ctr);
        const float *P = sH + cs * HSLOT_F + own;
        const float4 ym2 = lds128(P - 2 * T::HP), ym1 = lds128(P - T::HP);
        const float4 yp1 = lds128(P + T::HP), yp2 = lds128(P + 2 * T::HP);
        const float2 zl = lds64(P - 2), zr = lds64(P + 4);

        float4 o;
        o.x = point<EXACT>(qc.x, qm2.x, qm1.x, qp1.x, qp2.x, ym2.x, ym1.x, yp1.x, yp2.x, zl.x, zl.y, qc.y, qc.z, u1v.x, mv.x, a.s.k);
        o.y = point<EXACT>(qc.y, qm2.y, qm1.y, qp1.y, qp2.y, ym2.y, ym1.y, yp1.y, yp2.y, zl.y, qc.x, qc.z, qc.w, u1v.y, mv.y, a.s.k);
        o.z = point<EXACT>(qc.z, qm2.z, qm1.z, qp1.z, qp2.z, ym2.z, ym1.z, yp1.z, yp2.z, qc.x, qc.y, qc.w, zr.x, u1v.z, mv.z, a.s.k);
        o.w = point<EXACT>(qc.w, qm2.w, qm1.w, qp1.w, qp2.w, ym2.w, ym1.w, yp1.w, yp2.w, qc.y, qc.z, zr.x, zr.y, u1v.w, mv.w, a.s.k);

        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * cs);  // stage j+2 (and centre slot j%S1) free

        if (chunk_has_src) {  // fused Section1: rare path, only chunks that contain a source cell
            const int X = Xa + j;
            const int c0 = sv.plane_off[X], c1 = sv.plane_off[X + 1];
            for (int i = c0; i < c1; ++i) {
                const SourceCell cell = sv.cells[i];
                const int dzc = cell.Z - Z;
                if (cell.Y == Y && dzc >= 0 && dzc < 4) {
                    if (dzc == 0) o.x = apply_cell(o.x, cell, sv);
                    else if (dzc == 1) o.y = apply_cell(o.y, cell, sv);
                    else if (dzc == 2) o.z = apply_cell(o.z, cell, sv);
                    else o.w = apply_cell(o.w, cell, sv);
                }
            }
        }
        if (store_ok) *reinterpret_cast<float4 *>(out) = o;
        out += plane;

        qm2 = qm1; qm1 = qc; qc = qp1; qp1 = qp2;
        if (++fs == S0) { fs = 0; fpar ^= 1; }
        if (++cs == S0) cs = 0;
        if (++ms == T::S1) ms = 0;
    }
}

// ---------------------------------------------------------------------------- host side
typedef void (*TmaKernelFn)(const TmaArgs);
struct Variant {
    int ty, tz, stages;
    bool exact;
    TmaKernelFn fn;
    int nt;
    size_t smem;
    int minb;
};

#define FDTD_VARIANT(TY_, TZ_, S_, MINB_)                                                                   \
    {TY_, TZ_, S_, true, stencil_tma_kernel<TY_, TZ_, S_, true, MINB_>, TileShape<TY_, TZ_, S_>::NT,         \
     (size_t)TileShape<TY_, TZ_, S_>::SMEM, MINB_},                                                          \
    {TY_, TZ_, S_, false, stencil_tma_kernel<TY_, TZ_, S_, false, MINB_>, TileShape<TY_, TZ_, S_>::NT,       \
     (size_t)TileShape<TY_, TZ_, S_>::SMEM, MINB_}

static const Variant g_variants[] = {
    // (TY, TZ, stages, min CTAs/SM).  MINB is chosen so the register cap stays >= 72 (no spills):
    // the steady-state loop keeps a 4-plane float4 queue + 6 neighbour vectors live.
    // For each tile the preferred stage count comes first (auto selection takes the first match).
    FDTD_VARIANT(16, 64, 6, 3),  FDTD_VARIANT(16, 64, 5, 2),  FDTD_VARIANT(32, 64, 6, 1),
    FDTD_VARIANT(32, 64, 8, 1),  FDTD_VARIANT(16, 128, 6, 1), FDTD_VARIANT(16, 128, 8, 1),
    FDTD_VARIANT(8, 128, 6, 3),  FDTD_VARIANT(8, 128, 5, 2),  FDTD_VARIANT(8, 64, 6, 5),
    FDTD_VARIANT(8, 64, 5, 4),   FDTD_VARIANT(16, 32, 6, 5),  FDTD_VARIANT(16, 32, 5, 4),
    FDTD_VARIANT(8, 32, 6, 8),   FDTD_VARIANT(32, 32, 6, 3),
};
static const int g_nvariants = (int)(sizeof(g_variants) / sizeof(g_variants[0]));

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

static int encode_map(CUtensorMap *map, const float *base, int rank, const cuuint64_t *dims, const cuuint32_t *box)
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
                   int sm_count)
{
    p.valid = false;
    if (!tma_supported(g)) return (int)cudaErrorInvalidValue;
    const int ny = g.Y1 - g.Y0, nz = g.Z1 - g.Z0, nx = g.X1 - g.X0;

    // ---- tile choice: explicit, or the auto heuristic (largest z tile that the row fills, then the
    // y tile / stage count with the most bytes in flight that still gives >= 2 CTAs per SM).
    int ty = cfg.ty, tz = cfg.tz;
    if (tz <= 0) tz = nz >= 64 ? 64 : 32;
    if (ty <= 0) ty = ny >= 16 ? 16 : 8;
    int vi = -1;
    for (int i = 0; i < g_nvariants && vi < 0; ++i)
        if (g_variants[i].ty == ty && g_variants[i].tz == tz && g_variants[i].exact == exact &&
            (cfg.stages <= 0 || g_variants[i].stages == cfg.stages))
            vi = i;
    if (vi < 0) return (int)cudaErrorInvalidValue;
    const int stages = g_variants[vi].stages;
    const Variant &v = g_variants[vi];

    cuuint64_t dims_u[4] = {(cuuint64_t)g.nzp, (cuuint64_t)g.nyp, (cuuint64_t)g.nxp, 3};
    cuuint32_t box_h[4] = {(cuuint32_t)(tz + 8), (cuuint32_t)(ty + 4), 1, 1};
    cuuint32_t box_c[4] = {(cuuint32_t)tz, (cuuint32_t)ty, 1, 1};
    int rc;
    if ((rc = encode_map(&p.map_halo, u, 4, dims_u, box_h))) return rc;
    if ((rc = encode_map(&p.map_ctr, u, 4, dims_u, box_c))) return rc;
    if ((rc = encode_map(&p.map_m, m, 3, dims_u, box_c))) return rc;

    cudaError_t e = cudaFuncSetAttribute((const void *)v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)v.fn, v.nt, v.smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return (int)cudaErrorInvalidConfiguration;

    // ---- x chunking: enough CTAs to fill every SM `occ` deep in one wave, no more.
    const int tiles = ((ny + ty - 1) / ty) * ((nz + tz - 1) / tz);
    int xchunk = cfg.xchunk;
    if (xchunk <= 0) {
        const int slots = sm_count * occ;
        int nchunks = slots / tiles;
        if (nchunks < 1) nchunks = 1;
        if (nchunks > nx) nchunks = nx;
        xchunk = (nx + nchunks - 1) / nchunks;
        if (xchunk < 8 && nx >= 8) xchunk = 8;
    }
    p.ty = ty;
    p.tz = tz;
    p.stages = stages;
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
    args.s = a;
    args.tiles_z = (nz + p.tz - 1) / p.tz;
    args.tiles_y = (ny + p.ty - 1) / p.ty;
    args.xchunk = p.xchunk;
    dim3 grid(args.tiles_z * args.tiles_y, (nx + p.xchunk - 1) / p.xchunk, 1);
    if (grid.y > 65535) return (int)cudaErrorInvalidValue;
    v.fn<<<grid, v.nt, v.smem, stream>>>(args);
    return (int)cudaGetLastError();
}

}  // namespace fdtd
