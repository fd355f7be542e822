// stencil_tc2.cu -- two leapfrog steps per pass as a TIME-STEP PIPELINE ACROSS A THREAD-BLOCK CLUSTER (sm_100a).
//
// Same contract as stencil_tb2.cu (one pass reads u^{n-1}, u^n, m once and writes u^{n+1} and u^{n+2}; bit-identical
// to two one-step launches), different machine mapping.  stencil_tb2 runs both steps in ONE CTA: every thread carries
// two register queues, the 80-register cap forces the compiler to rematerialise addresses, and the CTA's 22 warps wait
// on each other through a shared-memory ring (profiles/r02_ncu_tb2_source_counters.txt: 65 lane-instructions per
// point-update, a quarter of them barrier spinning).  Here the two steps live on the TWO CTAs OF A CLUSTER, i.e. on two
// SMs, and the intermediate level travels through distributed shared memory:
//   * CTA rank 0 ("A", step 1) is the first half of stencil_tb2: TMA ring of u^n tiles with a radius-4 halo and of
//     u^{n-1} / m on the tile extended by 2 rows and one float4 column per side, one register queue per thread.  Its
//     u^{n+1} values go to global memory (the output tile) AND, for the whole extended tile, straight into a ring in
//     the PARTNER's shared memory (st.async: DSMEM stores counted on the partner's mbarrier like TMA bytes), plus one
//     remote arrive.expect_tx per warp;
//   * CTA rank 1 ("B", step 2) is the one-step streaming kernel (stencil_tma.cu) whose halo-plane ring is filled by A
//     instead of by TMA: it waits on its own mbarriers, keeps its column of u^{n+1} in a register queue, reads the y/z
//     neighbours from the ring, loads the centre tiles of u^n ("previous" level of step 2) and m with its own TMA
//     producer, and hands ring slots back with a remote arrive on A's `bfree` barriers.
// Each SM runs a lean one-step-like loop at ~92 registers; the DRAM traffic is that of stencil_tb2 (B's u^n and m
// tiles hit in L2, A fetched them 4-6 planes earlier).  No __syncthreads in the steady state; the only cluster-wide
// barriers are after the mbarrier initialisation and before exit (a CTA's shared memory must outlive its partner's
// accesses).  Unlinked slabs only (linked slabs use stencil_tb2).
#include "fdtd_arith.cuh"
#include "fdtd_kernels.cuh"
#include "tma_ptx.cuh"

#include <math.h>
#include <stdlib.h>

namespace fdtd {

struct Tc2Args {
    alignas(64) CUtensorMap map_cur;   // A: u as (z,y,x,level), box (HP, ER+4): u^n with the radius-4 halo
    alignas(64) CUtensorMap map_prev;  // A: u, box (HP, ER): u^{n-1} on the extended tile
    alignas(64) CUtensorMap map_m;     // A: m, box (HP, ER)
    alignas(64) CUtensorMap map_ctr;   // B: u, box (TZ, TY): u^n on the output tile
    alignas(64) CUtensorMap map_mc;    // B: m, box (TZ, TY)
    Tb2Step s;
    int tiles_z, tiles_y, xchunk, nchunks;
};

template <int TY_, int TZ_, int RY_>
struct Tc2Shape {
    static constexpr int TY = TY_, TZ = TZ_, RY = RY_;    // RY = rows per consumer thread (1 or 2)
    static constexpr int ER = TY + 4, EC = TZ / 4 + 2;   // extended tile: rows, float4 columns
    static constexpr int HP = 4 * EC;                     // pitch of every extended-tile slot (floats)
    static_assert(TY % RY == 0 && ER % RY == 0 && (RY == 1 || RY == 2), "rows per thread must divide both tiles");
    static constexpr int NCA = (ER / RY) * EC;            // A: one consumer thread per RY rows x one float4 column of the extended tile
    static constexpr int NCB = (TY / RY) * (TZ / 4);      // B: the same on the output tile
    static constexpr int NWA = (NCA + 31) / 32, NWB = (NCB + 31) / 32;
    static constexpr int NC = NWA * 32, NT = NC + 32;     // + one producer warp (both roles)
    // rings: u^n (A; 4 of its slots are the x look-ahead of the stencil, the rest is prefetch), u^{n-1} and m (A; they become
    // free two iterations before the u^n slot of the same stage, so SU - 2 slots share the u^n ring's `empty` barriers, as
    // in stencil_tma.cu), step-1 planes and centre tiles (B)
    static constexpr int SU = 7, SP = SU - 2, SB = 8, SC = 6;
    static constexpr int UBYTES = (ER + 4) * HP * 4, USLOT = (UBYTES + 127) / 128 * 128;
    static constexpr int CBYTES = ER * HP * 4, CSLOT = (CBYTES + 127) / 128 * 128;
    static constexpr int TBYTES = TY * TZ * 4;            // one centre tile
    // barrier block (bytes from the start of shared memory; the same layout in both CTAs, each uses its part)
    static constexpr int B_FULL = 0, B_EMPTY = B_FULL + 8 * SU, B_BFREE = B_EMPTY + 8 * SU;  // A
    static constexpr int B_BFULL = B_BFREE + 8 * SB, B_CFULL = B_BFULL + 8 * SB, B_CEMPTY = B_CFULL + 8 * SC;  // B
    static constexpr int DATA0 = 512;                     // also the guard in front of the rings (column 0 reads 2 floats to its left)
    static_assert(B_CEMPTY + 8 * SC <= DATA0 - 16, "barrier block overflows");
    static constexpr int SMEM_A = DATA0 + SU * USLOT + 2 * SP * CSLOT;
    static constexpr int SMEM_B = DATA0 + SB * CSLOT + 2 * SC * TBYTES;
    static constexpr int SMEM = (SMEM_A > SMEM_B ? SMEM_A : SMEM_B) + 128;
    static_assert(TZ % 4 == 0 && TBYTES % 128 == 0, "centre tiles must stay 128-byte aligned");
    static_assert(NT <= 1024 && SMEM <= 232448, "CTA too large");
};

// ---------------------------------------------------------------------------- cluster PTX
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// 16 bytes into the partner's shared memory through the async proxy; the bytes are counted on the partner's mbarrier
// (complete_tx), so a plain wait on that barrier sees them -- like TMA data, no fence on either side.  (A generic
// st.shared::cluster + arrive.release.cluster costs MEMBAR.ALL.GPU per warp and iteration, and acquire.cluster waits
// add a CCTL.IVALL each: 2.3x slower than stencil_tb2 when tried, profiles/r02_sweep512_tc2_fences.txt.)
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, const float4 &v, uint32_t cluster_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(cluster_bar)
                 : "memory");
}
// one arrival on the partner's barrier that also announces `bytes` of st.async data (relaxed: the data is ordered by
// the barrier's transaction count, not by this arrive)
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t cluster_bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
// plain arrival on the partner's barrier ("I have read the slot"), the form CUTLASS's cluster pipelines use
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::
                     : "memory");
}

// Source cells of one plane that fall into this thread's float4 (same as stencil_tb2.cu).
__device__ __forceinline__ void tc2_inject_plane(float4 &r, int X, int Y, int Z, const SourceView &sv)
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

template <int TY, int TZ, int RY, bool EXACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Shape<TY, TZ, RY>::NT, 1) stencil_tc2_kernel(const __grid_constant__ Tc2Args a)
{
    using T = Tc2Shape<TY, TZ, RY>;
    constexpr int SU = T::SU, SP = T::SP, SB = T::SB, SC = T::SC, HP = T::HP, ER = T::ER, EC = T::EC;
    constexpr int USLOT_F = T::USLOT / 4, CSLOT_F = T::CSLOT / 4, TILE_F = TY * TZ;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t rank = cluster_ctarank();

    const Grid &g = a.s.g;
    const int lane = threadIdx.x & 31;
    // PERSISTENT pairs: pair p works on items p, p + npairs, ... (item = chunk * tiles + tile, so the pairs running at the
    // same time sit on neighbouring tiles of one x chunk).  All rings and barrier phases simply continue from one item to
    // the next -- the partner CTA trails by a few planes, and with one item per CTA pair that lag (plus the two cluster
    // barriers) idled each SM for ~10 % of its time (profiles/r02_ncu_tc2_nonpersistent.txt: 11 % of the samples on UCGABAR_WAIT).
    const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int tiles = a.tiles_z * a.tiles_y;
    const int nitems = tiles * a.nchunks;

    if (threadIdx.x == 0) {
        if (rank == 0) {
            for (int i = 0; i < SU; ++i) {
                mbar_init(smem0 + T::B_FULL + 8 * i, 1);
                mbar_init(smem0 + T::B_EMPTY + 8 * i, T::NWA);
            }
            for (int i = 0; i < SB; ++i) mbar_init(smem0 + T::B_BFREE + 8 * i, T::NWB);  // arrived by B's warps
        } else {
            for (int i = 0; i < SB; ++i) mbar_init(smem0 + T::B_BFULL + 8 * i, T::NWA);  // arrived by A's warps
            for (int i = 0; i < SC; ++i) {
                mbar_init(smem0 + T::B_CFULL + 8 * i, 1);
                mbar_init(smem0 + T::B_CEMPTY + 8 * i, T::NWB);
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();  // both CTAs' barriers exist before anyone arrives on them remotely

    const SourceView &sv = a.s.sv;
    const long long plane = (long long)g.nyp * g.nzp;

    if (rank == 0) {
        // ================================================================================== A: step 1
        float *sU = reinterpret_cast<float *>(smem + T::DATA0);
        float *sP = sU + SU * USLOT_F;
        float *sM = sP + SP * CSLOT_F;
        const uint32_t full0 = smem0 + T::B_FULL, empty0 = smem0 + T::B_EMPTY, bfree0 = smem0 + T::B_BFREE;
        const uint32_t r_bfull0 = mapa_u32(smem0 + T::B_BFULL, 1);   // partner's barriers and ring, cluster addresses
        const uint32_t r_ring0 = mapa_u32(smem0 + T::DATA0, 1);
        if (threadIdx.x >= T::NC) {
            if (threadIdx.x == T::NC) {
                // producer: stage s of an item carries u^n plane Xa-4+s and, for s >= 4, u^{n-1} / m plane Xa-6+s (iteration
                // s-4).  A u^n slot is free once the stage in it has been read as a centre plane (or the item ended); the
                // u^{n-1} / m slot of the same stage was last read two iterations earlier.
                int us = 0, ps = 0, use = 0;
                for (int item = pair; item < nitems; item += npairs) {
                    const int tile = item % tiles, chunk = item / tiles;
                    const int Xa = g.X0 + chunk * a.xchunk, np = min(g.X1, Xa + a.xchunk) - Xa;
                    const int Yt = g.Y0 + (tile / a.tiles_z) * TY, Zt = g.Z0 + (tile % a.tiles_z) * TZ;
                    const int nst = np + 8;
                    for (int s = 0; s < nst; ++s) {
                        const int Xp = Xa - 4 + s;
                        if (use > 0) mbar_wait(empty0 + 8 * us, (use - 1) & 1);
                        const uint32_t bar = full0 + 8 * us;
                        const bool ctr = s >= 4;
                        mbar_expect_tx(bar, T::UBYTES + (ctr ? 2 * T::CBYTES : 0));
                        tma_load_4d(smem_u32(sU) + us * T::USLOT, &a.map_cur, bar, Zt - 4, Yt - 4, Xp, a.s.l_cur);
                        if (ctr) {
                            tma_load_4d(smem_u32(sP) + ps * T::CSLOT, &a.map_prev, bar, Zt - 4, Yt - 2, Xp - 2, a.s.l_prev);
                            tma_load_3d(smem_u32(sM) + ps * T::CSLOT, &a.map_m, bar, Zt - 4, Yt - 2, Xp - 2);
                            if (++ps == SP) ps = 0;
                        }
                        if (++us == SU) {
                            us = 0;
                            ++use;
                        }
                    }
                }
            }
        } else {
            const bool active = threadIdx.x < T::NCA;
            const int er = active ? (threadIdx.x / EC) * RY : 0, ec = active ? threadIdx.x % EC : 0;  // first of this thread's rows
            const int ownU = (er + 2) * HP + 4 * ec;  // own column in a u^n slot (rows start at Yt-4)
            const int ownC = er * HP + 4 * ec;        // own column in an extended-tile slot (rows start at Yt-2)
            const uint32_t r_own = r_ring0 + 4u * (uint32_t)ownC;
            // bytes this warp sends per plane: 16 per row of an active lane (the last warp of the extended tile is partly idle)
            const int warp_first = (int)(threadIdx.x & ~31u);
            const uint32_t warp_bytes = 16u * RY * (uint32_t)max(0, min(32, T::NCA - warp_first));
            int us = 0, pp = 0, b8 = 0;   // running slots: next u^n stage to read, u^{n-1} / m, the partner's ring
            uint32_t upar = 0;            // parity of the ring use `us` is in
            uint32_t fpar = 0;            // parity to wait for on bfree[b8] from the second ring use on
            bool lapped = false;          // the partner's ring has been written once around

            for (int item = pair; item < nitems; item += npairs) {
                const int tile = item % tiles, chunk = item / tiles;
                const int Xa = g.X0 + chunk * a.xchunk, Xb = min(g.X1, Xa + a.xchunk), np = Xb - Xa;
                const int Yt = g.Y0 + (tile / a.tiles_z) * TY, Zt = g.Z0 + (tile % a.tiles_z) * TZ;
                const int nit = np + 4;  // step-1 planes Xa-2 .. Xb+1
                const int Y = Yt - 2 + er, Z = Zt - 4 + 4 * ec;
                const bool z_in = active && Z >= g.Z0 && Z < g.Z1;
                const bool rows_core = er >= 2 && er < ER - 2 && ec >= 1 && ec < EC - 1;  // rows of the output tile (2 | RY-aligned)
                bool inb[RY], core[RY], any_inb = false;
#pragma unroll
                for (int r = 0; r < RY; ++r) {
                    inb[r] = z_in && Y + r >= g.Y0 && Y + r < g.Y1;  // interior in (y,z)
                    core[r] = inb[r] && rows_core;                    // inside the output tile
                    any_inb |= inb[r];
                }
                bool chunk_has_src = false;
                if (sv.ncells > 0) chunk_has_src = (sv.plane_off[min(Xb + 2, g.nxp)] - sv.plane_off[max(Xa - 2, 0)]) > 0;
                float *__restrict__ out1 = a.s.u + (long long)a.s.l_n1 * g.lvl + (long long)(Xa - 2) * plane + (long long)Y * g.nzp + Z;

                // prologue: stages 0..3 of the item; stages 0 and 1 are never a centre plane and go back at once
                float4 qU[5][RY];
                int uc = us;  // slot of the item's stage 0; after the prologue it trails `us` by two stages (the centre plane)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    mbar_wait(full0 + 8 * us, upar);
#pragma unroll
                    for (int r = 0; r < RY; ++r) qU[s][r] = lds128(sU + us * USLOT_F + ownU + r * HP);
                    if (s < 2) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty0 + 8 * us);
                        if (++uc == SU) uc = 0;
                    }
                    if (++us == SU) {
                        us = 0;
                        upar ^= 1;
                    }
                }
                for (int i0 = 0; i0 < nit; i0 += 5) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int i = i0 + k;
                        if (i >= nit) break;
                        mbar_wait(full0 + 8 * us, upar);   // front stage i+4
#pragma unroll
                        for (int r = 0; r < RY; ++r) qU[(k + 4) % 5][r] = lds128(sU + us * USLOT_F + ownU + r * HP);
                        const int P1 = Xa - 2 + i;
                        float4 res[RY];  // halo cells keep their value (identical in every level by construction)
#pragma unroll
                        for (int r = 0; r < RY; ++r) res[r] = qU[(k + 2) % 5][r];
                        if (any_inb && P1 >= g.X0 && P1 < g.X1) {
                            const float *P = sU + uc * USLOT_F + ownU;
                            // this thread's y column on the centre plane: 2 rows above, own rows (registers), 2 rows below
                            float4 col[RY + 4];
                            col[0] = lds128(P - 2 * HP);
                            col[1] = lds128(P - HP);
                            col[RY + 2] = lds128(P + RY * HP);
                            col[RY + 3] = lds128(P + (RY + 1) * HP);
#pragma unroll
                            for (int r = 0; r < RY; ++r) col[r + 2] = qU[(k + 2) % 5][r];
#pragma unroll
                            for (int r = 0; r < RY; ++r) {
                                const float2 zl = lds64(P + r * HP - 2), zr = lds64(P + r * HP + 4);
                                const float4 pv = lds128(sP + pp * CSLOT_F + ownC + r * HP);
                                const float4 mv = lds128(sM + pp * CSLOT_F + ownC + r * HP);
                                float4 v = column4<EXACT>(col[r + 2], qU[k % 5][r], qU[(k + 1) % 5][r], qU[(k + 3) % 5][r], qU[(k + 4) % 5][r],
                                                          col[r], col[r + 1], col[r + 3], col[r + 4], zl, zr, pv, mv, a.s.k);
                                if (chunk_has_src) tc2_inject_plane(v, P1, Y + r, Z, sv);  // rare: source cells of step n (ghost zone too)
                                if (inb[r]) res[r] = v;
                                if (core[r] && P1 >= Xa && P1 < Xb) *reinterpret_cast<float4 *>(out1 + r * g.nzp) = v;
                            }
                        }
                        out1 += plane;
                        // the step-1 plane goes into the next slot of the partner's ring
                        if (lapped) mbar_wait(bfree0 + 8 * b8, fpar);
                        if (active) {
#pragma unroll
                            for (int r = 0; r < RY; ++r) st_async_v4(r_own + (uint32_t)(b8 * T::CSLOT + r * HP * 4), res[r], r_bfull0 + 8 * b8);
                        }
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive_expect_tx_remote(r_bfull0 + 8 * b8, warp_bytes);  // this warp's part of the plane is on its way
                            mbar_arrive(empty0 + 8 * uc);  // stage i+2 (and the u^{n-1} / m slot of iteration i) free
                        }
                        if (++us == SU) {
                            us = 0;
                            upar ^= 1;
                        }
                        if (++uc == SU) uc = 0;
                        if (++pp == SP) pp = 0;
                        if (++b8 == SB) {
                            b8 = 0;
                            if (lapped) fpar ^= 1;
                            lapped = true;
                        }
                    }
                }
                // stages nit+2 and nit+3 were read as front planes only: hand their slots back
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(empty0 + 8 * uc);
                    mbar_arrive(empty0 + 8 * (uc + 1 == SU ? 0 : uc + 1));
                }
            }
        }
    } else {
        // ================================================================================== B: step 2
        float *sB = reinterpret_cast<float *>(smem + T::DATA0);
        float *sV = sB + SB * CSLOT_F;       // u^n on the output tile ("previous" level of step 2)
        float *sMB = sV + SC * TILE_F;       // m on the output tile
        const uint32_t bfull0 = smem0 + T::B_BFULL, cfull0 = smem0 + T::B_CFULL, cempty0 = smem0 + T::B_CEMPTY;
        const uint32_t r_bfree0 = mapa_u32(smem0 + T::B_BFREE, 0);
        if (threadIdx.x >= T::NC) {
            if (threadIdx.x == T::NC) {
                int cs = 0, use = 0;
                for (int item = pair; item < nitems; item += npairs) {
                    const int tile = item % tiles, chunk = item / tiles;
                    const int Xa = g.X0 + chunk * a.xchunk, np = min(g.X1, Xa + a.xchunk) - Xa;
                    const int Yt = g.Y0 + (tile / a.tiles_z) * TY, Zt = g.Z0 + (tile % a.tiles_z) * TZ;
                    for (int j = 0; j < np; ++j) {
                        if (use > 0) mbar_wait(cempty0 + 8 * cs, (use - 1) & 1);
                        const uint32_t bar = cfull0 + 8 * cs;
                        mbar_expect_tx(bar, 2 * T::TBYTES);
                        tma_load_4d(smem_u32(sV) + cs * T::TBYTES, &a.map_ctr, bar, Zt, Yt, Xa + j, a.s.l_cur);
                        tma_load_3d(smem_u32(sMB) + cs * T::TBYTES, &a.map_mc, bar, Zt, Yt, Xa + j);
                        if (++cs == SC) {
                            cs = 0;
                            ++use;
                        }
                    }
                }
            }
        } else if (threadIdx.x < T::NWB * 32) {
            const bool active = threadIdx.x < T::NCB;
            constexpr int ZQ = TZ / 4;
            const int yr = active ? (threadIdx.x / ZQ) * RY : 0, zq = active ? threadIdx.x % ZQ : 0;  // first of this thread's rows
            const int ownC = (yr + 2) * HP + 4 * (zq + 1);  // own column in an extended-tile slot
            const int ctr = yr * TZ + 4 * zq;               // own column in a centre tile
            SourceView sv2 = sv;
            sv2.src_row = a.s.src_row2;
            int f8 = 0, cs = 0;              // running slots: next step-1 plane to read, centre tiles
            uint32_t fpar = 0, cpar = 0;     // parities of those ring uses

            for (int item = pair; item < nitems; item += npairs) {
                const int tile = item % tiles, chunk = item / tiles;
                const int Xa = g.X0 + chunk * a.xchunk, Xb = min(g.X1, Xa + a.xchunk), np = Xb - Xa;
                const int Yt = g.Y0 + (tile / a.tiles_z) * TY, Zt = g.Z0 + (tile % a.tiles_z) * TZ;
                const int Y = Yt + yr, Z = Zt + 4 * zq;
                const bool z_ok = active && Z < g.Z1;
                bool chunk_has_src = false;
                if (sv.ncells > 0) chunk_has_src = (sv.plane_off[Xb] - sv.plane_off[Xa]) > 0;
                float *__restrict__ out2 = a.s.u + (long long)a.s.l_n2 * g.lvl + (long long)Xa * plane + (long long)Y * g.nzp + Z;

                // qR[j % 5] = own rows of the item's step-1 plane j (A's iteration j, plane Xa-2+j)
                float4 qR[5][RY];
                int c8 = f8;  // slot of the item's plane 0; after the prologue it trails `f8` by two planes (the centre plane)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mbar_wait(bfull0 + 8 * f8, fpar);
#pragma unroll
                    for (int r = 0; r < RY; ++r) qR[j][r] = lds128(sB + f8 * CSLOT_F + ownC + r * HP);
                    if (j < 2) {  // planes 0 and 1 are never a centre plane: hand their slots back now
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(r_bfree0 + 8 * f8);
                        if (++c8 == SB) c8 = 0;
                    }
                    if (++f8 == SB) {
                        f8 = 0;
                        fpar ^= 1;
                    }
                }
                for (int i0 = 0; i0 < np; i0 += 5) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int i = i0 + k;
                        if (i >= np) break;
                        mbar_wait(bfull0 + 8 * f8, fpar);  // front plane i+4
#pragma unroll
                        for (int r = 0; r < RY; ++r) qR[(k + 4) % 5][r] = lds128(sB + f8 * CSLOT_F + ownC + r * HP);
                        const float *P = sB + c8 * CSLOT_F + ownC;
                        float4 col[RY + 4];
                        col[0] = lds128(P - 2 * HP);
                        col[1] = lds128(P - HP);
                        col[RY + 2] = lds128(P + RY * HP);
                        col[RY + 3] = lds128(P + (RY + 1) * HP);
                        float2 zl[RY], zr[RY];
#pragma unroll
                        for (int r = 0; r < RY; ++r) zl[r] = lds64(P + r * HP - 2), zr[r] = lds64(P + r * HP + 4);
                        mbar_wait(cfull0 + 8 * cs, cpar);
                        float4 pv[RY], mv[RY];
#pragma unroll
                        for (int r = 0; r < RY; ++r) {
                            pv[r] = lds128(sV + cs * TILE_F + ctr + r * TZ);
                            mv[r] = lds128(sMB + cs * TILE_F + ctr + r * TZ);
                        }
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive_remote(r_bfree0 + 8 * c8);  // A may overwrite the slot of step-1 plane i+2
                            mbar_arrive(cempty0 + 8 * cs);
                        }
#pragma unroll
                        for (int r = 0; r < RY; ++r) col[r + 2] = qR[(k + 2) % 5][r];
#pragma unroll
                        for (int r = 0; r < RY; ++r) {
                            float4 o = column4<EXACT>(col[r + 2], qR[k % 5][r], qR[(k + 1) % 5][r], qR[(k + 3) % 5][r], qR[(k + 4) % 5][r], col[r],
                                                      col[r + 1], col[r + 3], col[r + 4], zl[r], zr[r], pv[r], mv[r], a.s.k);
                            if (chunk_has_src) tc2_inject_plane(o, Xa + i, Y + r, Z, sv2);  // source cells of step n+1
                            if (z_ok && Y + r < g.Y1) *reinterpret_cast<float4 *>(out2 + r * g.nzp) = o;
                        }
                        out2 += plane;
                        if (++f8 == SB) {
                            f8 = 0;
                            fpar ^= 1;
                        }
                        if (++c8 == SB) c8 = 0;
                        if (++cs == SC) {
                            cs = 0;
                            cpar ^= 1;
                        }
                    }
                }
                // planes np+2 and np+3 were read as front planes only: hand their slots back
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_remote(r_bfree0 + 8 * c8);
                    mbar_arrive_remote(r_bfree0 + 8 * (c8 + 1 == SB ? 0 : c8 + 1));
                }
            }
        }
    }
    __syncwarp();
    cluster_sync_all();  // neither CTA may exit while its partner can still write its shared memory / barriers
}

// ---------------------------------------------------------------------------- host side
typedef void (*Tc2KernelFn)(const Tc2Args);
struct Tc2Variant {
    int ty, tz, rows;
    bool exact;
    Tc2KernelFn fn;
    int nt;
    size_t smem;
};
#define FDTD_TC2_1(TY_, TZ_, RY_, EX_) \
    {TY_, TZ_, RY_, EX_, stencil_tc2_kernel<TY_, TZ_, RY_, EX_>, Tc2Shape<TY_, TZ_, RY_>::NT, (size_t)Tc2Shape<TY_, TZ_, RY_>::SMEM}
#define FDTD_TC2(TY_, TZ_, RY_) FDTD_TC2_1(TY_, TZ_, RY_, false), FDTD_TC2_1(TY_, TZ_, RY_, true)
static const Tc2Variant g_tc2[] = {
    // output tile, rows per thread; the first match that divides the grid wins (else the first match)
    FDTD_TC2(16, 128, 2), FDTD_TC2(32, 64, 2), FDTD_TC2(16, 128, 1), FDTD_TC2(32, 64, 1), FDTD_TC2(28, 64, 1), FDTD_TC2(24, 64, 1),
    FDTD_TC2(28, 64, 2), FDTD_TC2(24, 64, 2), FDTD_TC2(16, 64, 1), FDTD_TC2(12, 128, 1), FDTD_TC2(40, 64, 2),
};
static const int g_ntc2 = (int)(sizeof(g_tc2) / sizeof(g_tc2[0]));

// FDTD_B200_TC2_PAIRS: resident CTA pairs (default SMs / 2); -1 = one pair per item (no persistence)
static int env_pairs()
{
    static const int v = [] {
        const char *e = getenv("FDTD_B200_TC2_PAIRS");
        return (e && *e) ? atoi(e) : 0;
    }();
    return v;
}

int tc2_plan_build(Tc2Plan &p, float *u, const float *m, const Grid &g, const TmaConfig &cfg, bool exact, int sm_count)
{
    p.valid = false;
    if (!tma_supported(g)) return (int)cudaErrorInvalidValue;
    const int ny = g.Y1 - g.Y0, nz = g.Z1 - g.Z0, nx = g.X1 - g.X0;
    int vi = -1;
    for (int pass = 0; pass < 2 && vi < 0; ++pass)
        for (int i = 0; i < g_ntc2 && vi < 0; ++i)
            if (g_tc2[i].exact == exact && (cfg.ty <= 0 || g_tc2[i].ty == cfg.ty) && (cfg.tz <= 0 || g_tc2[i].tz == cfg.tz) &&
                (cfg.rows <= 0 || g_tc2[i].rows == cfg.rows) &&
                (pass == 1 || cfg.ty > 0 || cfg.tz > 0 || (ny % g_tc2[i].ty == 0 && nz % g_tc2[i].tz == 0)))
                vi = i;
    if (vi < 0) return (int)cudaErrorInvalidValue;
    const Tc2Variant &v = g_tc2[vi];
    const int er = v.ty + 4, hp = v.tz + 8;
    cuuint64_t dims_u[4] = {(cuuint64_t)g.nzp, (cuuint64_t)g.nyp, (cuuint64_t)g.nxp, (cuuint64_t)FDTD_LEVELS};
    cuuint32_t box_u[4] = {(cuuint32_t)hp, (cuuint32_t)(er + 4), 1, 1};
    cuuint32_t box_e[4] = {(cuuint32_t)hp, (cuuint32_t)er, 1, 1};
    cuuint32_t box_c[4] = {(cuuint32_t)v.tz, (cuuint32_t)v.ty, 1, 1};
    int rc;
    if ((rc = encode_tensor_map(&p.map_cur, u, 4, dims_u, box_u))) return rc;
    if ((rc = encode_tensor_map(&p.map_prev, u, 4, dims_u, box_e))) return rc;
    if ((rc = encode_tensor_map(&p.map_m, m, 3, dims_u, box_e))) return rc;
    if ((rc = encode_tensor_map(&p.map_ctr, u, 4, dims_u, box_c))) return rc;
    if ((rc = encode_tensor_map(&p.map_mc, m, 3, dims_u, box_c))) return rc;
    cudaError_t e = cudaFuncSetAttribute((const void *)v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem);
    if (e != cudaSuccess) return (int)e;
    // x chunking: one CTA PAIR per two SMs; several full waves, chunks long enough to amortise the 8-plane prologue
    const int tiles = ((ny + v.ty - 1) / v.ty) * ((nz + v.tz - 1) / v.tz);
    int xchunk = cfg.xchunk;
    if (xchunk <= 0) {
        const double slots = sm_count / 2;
        double best = -1.0;
        for (int nch = 1; nch <= nx; ++nch) {
            const int xc = (nx + nch - 1) / nch;
            if (xc < 16 && nch > 1) break;
            if ((nx + xc - 1) / xc != nch) continue;
            const double waves = tiles * (double)nch / slots, full = ceil(waves);
            const double eff = waves / full * full / (full + 0.25) * xc / (xc + 8.0);
            if (eff > best) {
                best = eff;
                xchunk = xc;
            }
        }
    }
    p.ty = v.ty;
    p.tz = v.tz;
    p.rows = v.rows;
    p.npairs = env_pairs() > 0 ? env_pairs() : (env_pairs() < 0 ? 1 << 30 : sm_count / 2);
    p.xchunk = xchunk;
    p.variant = vi;
    p.smem_bytes = v.smem;
    p.valid = true;
    return 0;
}

int launch_stencil_tc2(const Tc2Plan &p, const Tb2Step &a, bool exact, cudaStream_t stream)
{
    if (!p.valid) return (int)cudaErrorInvalidValue;
    const Tc2Variant &v = g_tc2[p.variant];
    if (v.exact != exact) return (int)cudaErrorInvalidValue;
    if (a.link.peer_u[0] || a.link.peer_u[1]) return (int)cudaErrorNotSupported;
    const int ny = a.g.Y1 - a.g.Y0, nz = a.g.Z1 - a.g.Z0, nx = a.g.X1 - a.g.X0;
    if (nx <= 0) return 0;
    Tc2Args args;
    args.map_cur = p.map_cur;
    args.map_prev = p.map_prev;
    args.map_m = p.map_m;
    args.map_ctr = p.map_ctr;
    args.map_mc = p.map_mc;
    args.s = a;
    args.tiles_z = (nz + p.tz - 1) / p.tz;
    args.tiles_y = (ny + p.ty - 1) / p.ty;
    args.xchunk = p.xchunk;
    args.nchunks = (nx + p.xchunk - 1) / p.xchunk;
    const int nitems = args.tiles_z * args.tiles_y * args.nchunks;
    const int npairs = nitems < p.npairs ? nitems : p.npairs;
    dim3 grid(2 * npairs, 1, 1);  // blockIdx.x = 2*pair + role; __cluster_dims__(2,1,1) pairs them; pairs loop over the items
    v.fn<<<grid, v.nt, v.smem, stream>>>(args);
    return (int)cudaGetLastError();
}

}  // namespace fdtd
