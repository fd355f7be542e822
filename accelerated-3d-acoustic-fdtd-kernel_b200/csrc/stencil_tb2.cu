// stencil_tb2.cu -- temporal blocking: TWO leapfrog steps per pass (t_fuse = 2), sm_100a.
//
// One pass reads u^{n-1}, u^n and m once and writes u^{n+1} and u^{n+2}: ~20-24 B per point per two
// steps instead of 2 x 16 B -- below the single-step compulsory floor.  No reference kernel does this
// (the "temporal blocking" comment in cuda_optimized.cu:52-55 is x look-ahead).
//
// Structure: overlapped (ghost-zone) tiles in (y,z), streaming in x, both steps in one CTA.
//   * a CTA owns a TY x TZ output tile and computes step 1 (u^{n+1}) on the tile extended by 2 rows and
//     one float4 column on every side (ER x EC float4 columns = one consumer thread each), step 2 (u^{n+2})
//     on the tile itself, two planes behind step 1;
//   * TMA + mbarrier ring exactly as in stencil_tma.cu: u^n tiles with a radius-4 halo, u^{n-1} and m on the
//     extended tile, one producer thread, `done[]` barriers (one arrive per warp and iteration) free the slots;
//   * a thread keeps its own column of u^n (5 planes) AND of its step-1 results (5 planes) in registers, so
//     step 2 takes its x neighbours, its centre and its "previous" value (u^n) from registers; only the
//     y/z neighbours of the step-1 plane come from a 6-slot shared-memory ring that the CTA's threads fill
//     (STS + one mbarrier arrive per warp); the consumer of a plane runs two iterations behind its producer,
//     so nobody waits in the steady state and no __syncthreads exists;
//   * points outside the interior box keep their halo value: the host only selects this kernel when the
//     halo shells of all levels are bit-identical (launch_shell_check) and no source cell lies in a halo;
//   * sources are injected into BOTH steps in p_src order (every CTA that recomputes a step-1 cell in its
//     ghost zone injects it too), with the src rows of step n and n+1.
// Arithmetic: the same point<EXACT>() as the single-step kernels => a two-step pass is BIT-IDENTICAL to two
// single-step contracted launches (tests/test_tb2_gpu.py).
#include "fdtd_arith.cuh"
#include "fdtd_kernels.cuh"
#include "tma_ptx.cuh"
#include "stencil_tb2.cuh"

#include <math.h>
#include <stdlib.h>

namespace fdtd {

template <int ER, int EC, int RY>
struct Tb2Shape {
    static_assert(ER % RY == 0 && (RY == 1 || RY == 2), "rows per thread must divide the extended tile (and its 2-row ghost zone)");
    static constexpr int TY = ER - 4, TZ = 4 * EC - 8;  // output tile
    static constexpr int TR = ER / RY;                  // thread rows
    static constexpr int NCA = TR * EC;                 // active consumer threads: RY rows x one float4 column each
    static constexpr int NC = (NCA + 31) / 32 * 32;
    static constexpr int NCW = NC / 32;
    static constexpr int NT = NC + 32;                  // + producer warp
    static constexpr int HP = 4 * EC;                   // pitch of every slot (floats)
    static constexpr int SU = 5, SP = 3, SM = 5, SB = 6, ND = 8;  // ring depths: u^n, u^{n-1}, m, step-1 planes, done[]
    static constexpr int UBYTES = (ER + 4) * HP * 4;
    static constexpr int USLOT = (UBYTES + 127) / 128 * 128;
    static constexpr int CBYTES = ER * HP * 4;
    static constexpr int CSLOT = (CBYTES + 127) / 128 * 128;
    static constexpr int PAD = 128;  // guard before the first ring: column 0 reads two floats to its left
    static constexpr int SMEM = PAD + SU * USLOT + (SP + SM + SB) * CSLOT + (SU + ND + SB + 1) * 8 + 128;
    static_assert(NT <= 1024, "too many threads");
    static_assert(SMEM <= 232448, "shared memory of one CTA exceeds 227 KB");
};

template <int ER, int EC, int RY, bool EXACT>
__global__ void __launch_bounds__(Tb2Shape<ER, EC, RY>::NT, 1) stencil_tb2_kernel(const __grid_constant__ Tb2Args a)
{
    using T = Tb2Shape<ER, EC, RY>;
    constexpr int SU = T::SU, SP = T::SP, SM = T::SM, SB = T::SB, ND = T::ND, HP = T::HP;
    constexpr int USLOT_F = T::USLOT / 4, CSLOT_F = T::CSLOT / 4;
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sU = reinterpret_cast<float *>(smem + T::PAD);
    float *sP = sU + SU * USLOT_F;
    float *sM = sP + SP * CSLOT_F;
    float *sB = sM + SM * CSLOT_F;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + SB * CSLOT_F);
    const uint32_t full0 = smem_u32(bars), done0 = smem_u32(bars + SU), bfull0 = smem_u32(bars + SU + ND);
    const uint32_t pro0 = smem_u32(bars + SU + ND + SB);  // prologue done: stages 0 and 1 have been read

    const Grid &g = a.s.g;
    const SlabLink &lk = a.s.link;
    const int tz = blockIdx.x % a.tiles_z, ty = blockIdx.x / a.tiles_z;
    // chunk order and the short boundary chunks: as in stencil_tma.cu
    const int nch = gridDim.y, by = blockIdx.y;
    const int chunk = by == 0 ? 0 : (by == 1 ? nch - 1 : by - 1);
    int Xa, Xb;
    if (a.edge == 0) {
        Xa = g.X0 + chunk * a.xchunk;
        Xb = min(g.X1, Xa + a.xchunk);
    } else if (chunk == 0) {
        Xa = g.X0;
        Xb = g.X0 + a.edge;
    } else if (chunk == nch - 1) {
        Xa = g.X1 - a.edge;
        Xb = g.X1;
    } else {
        Xa = g.X0 + a.edge + (chunk - 1) * a.xchunk;
        Xb = min(g.X1 - a.edge, Xa + a.xchunk);
    }
    const int np = Xb - Xa;
    const int Yt = g.Y0 + ty * T::TY, Zt = g.Z0 + tz * T::TZ;  // padded origin of the OUTPUT tile
    // step 1 is computed on [XC0, XC1): the slab's planes plus, towards a neighbour slab, its two nearest planes
    // (their u^n, u^{n-1} and m sit in this slab's ghost planes); towards a physical boundary the halo is kept
    const int XC0 = g.X0 - (lk.peer_u[0] ? 2 : 0), XC1 = g.X1 + (lk.peer_u[1] ? 2 : 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < SU; ++i) mbar_init(full0 + 8 * i, 1);
        for (int i = 0; i < ND; ++i) mbar_init(done0 + 8 * i, T::NCW);
        for (int i = 0; i < SB; ++i) mbar_init(bfull0 + 8 * i, T::NCW);
        mbar_init(pro0, T::NCW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nit = np + 4;  // step-1 iterations: planes Xa-2 .. Xb+1
    if (threadIdx.x >= T::NC) {
        // ------------------------------------------------------------------ producer (one thread)
        // stage s: u^n plane Xa-4+s; for s >= 4 also u^{n-1} and m plane Xa-6+s (= step-1 plane of iteration s-4)
        const bool tiled = lk.tile_mode && lk.wait;  // per-tile flags: see stencil_tma.cu
        if (tiled) {
            if (lk.peer_u[0] && Xa - 4 < g.X0) wait_tiles(lk.my_tile[0], ty, tz, a.tiles_y, a.tiles_z, lk.epoch - 1, lk.err);
            if (lk.peer_u[1] && Xb + 4 > g.X1) wait_tiles(lk.my_tile[1], ty, tz, a.tiles_y, a.tiles_z, lk.epoch - 1, lk.err);
        }
        if (threadIdx.x == T::NC) {
            const int nst = nit + 4;
            int us = 0, ps = 0, ms = 0;
            bool waited[2] = {tiled || !(lk.wait && lk.peer_u[0]), tiled || !(lk.wait && lk.peer_u[1])};
            for (int s = 0; s < nst; ++s) {
                const int Xp = Xa - 4 + s;  // ghost planes (outside [X0, X1)) are written by the neighbours' previous pass
                const int side = Xp < g.X0 ? 0 : (Xp >= g.X1 ? 1 : -1);
                if (side >= 0 && !waited[side]) {
                    wait_flag(lk.my_flag[side], lk.epoch - 1, lk.err);
                    waited[side] = true;
                }
                // slot reuse: stage s overwrites u^n stage s-5 (last read as the centre plane of iteration s-7), and the
                // u^{n-1} / m slots of iterations s-7 / s-9.  Stages 5 and 6 overwrite stages 0 and 1, which only the
                // prologue reads.  (Re-arming a full barrier whose previous phase is still pending is undefined.)
                if (s == 5) mbar_wait(pro0, 0);
                if (s >= 7) mbar_wait(done0 + 8 * ((s - 7) % ND), ((s - 7) / ND) & 1);
                const uint32_t bar = full0 + 8 * us;
                const bool ctr = s >= 4;
                mbar_expect_tx(bar, T::UBYTES + (ctr ? 2 * T::CBYTES : 0));
                if (lk.pull && side >= 0 && lk.peer_u[side])  // a neighbour's plane, read where it lies (same level placement on every slab)
                    tma_load_4d(smem_u32(sU) + us * T::USLOT, &a.map_cur_peer[side], bar, Zt - 4, Yt - 4,
                                lk.peer_edge[side] + Xp - (side == 0 ? g.X0 : g.X1), a.s.l_cur);
                else
                    tma_load_4d(smem_u32(sU) + us * T::USLOT, &a.map_cur, bar, Zt - 4, Yt - 4, Xp, a.s.l_cur);
                if (ctr) {
                    const int Xq = Xp - 2, sq = Xq < g.X0 ? 0 : (Xq >= g.X1 ? 1 : -1);
                    if (lk.pull && sq >= 0 && lk.peer_u[sq])
                        tma_load_4d(smem_u32(sP) + ps * T::CSLOT, &a.map_prev_peer[sq], bar, Zt - 4, Yt - 2,
                                    lk.peer_edge[sq] + Xq - (sq == 0 ? g.X0 : g.X1), a.s.l_prev);
                    else
                        tma_load_4d(smem_u32(sP) + ps * T::CSLOT, &a.map_prev, bar, Zt - 4, Yt - 2, Xp - 2, a.s.l_prev);
                    tma_load_3d(smem_u32(sM) + ms * T::CSLOT, &a.map_m, bar, Zt - 4, Yt - 2, Xp - 2);
                    if (++ps == SP) ps = 0;
                    if (++ms == SM) ms = 0;
                }
                if (++us == SU) us = 0;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    // a thread owns RY consecutive rows x one float4 column of the extended tile
    const bool active = threadIdx.x < T::NCA;
    const int tr = active ? threadIdx.x / EC : 0, ec = active ? threadIdx.x % EC : 0;
    const int er = tr * RY;  // first row
    const int lane = threadIdx.x & 31;
    const int Y = Yt - 2 + er, Z = Zt - 4 + 4 * ec;
    const bool z_in = active && Z >= g.Z0 && Z < g.Z1;                  // interior in z: whole float4
    const bool rows_core = er >= 2 && er < ER - 2 && ec >= 1 && ec < EC - 1;  // rows of the output tile (2 | RY-aligned)
    bool inb[RY], core[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) {
        inb[r] = z_in && Y + r >= g.Y0 && Y + r < g.Y1;  // interior in (y,z)
        core[r] = inb[r] && rows_core;
    }
    bool any_inb = false, any_core = false;
#pragma unroll
    for (int r = 0; r < RY; ++r) any_inb |= inb[r], any_core |= core[r];
    const int ownU = (er + 2) * HP + 4 * ec;  // own first row in a u^n slot (rows start at Yt-4)
    const int ownC = er * HP + 4 * ec;        // own first row in a u^{n-1} / m / step-1 slot (rows start at Yt-2)

    const SourceView &sv = a.s.sv;
    bool chunk_has_src = false;
    if (sv.ncells > 0) chunk_has_src = (sv.plane_off[min(Xb + 2, g.nxp)] - sv.plane_off[max(Xa - 2, 0)]) > 0;
    SourceView sv2 = sv;
    sv2.src_row = a.s.src_row2;

    const long long plane = (long long)g.nyp * g.nzp;
    const long long row0 = (long long)Y * g.nzp + Z;
    float *__restrict__ out1 = a.s.u + (long long)a.s.l_n1 * g.lvl + row0;  // + X*plane + r*nzp
    float *__restrict__ out2 = a.s.u + (long long)a.s.l_n2 * g.lvl + row0;
    // boundary planes also go to the neighbours' ghost planes (peer stores over NVLink): the two outermost
    // planes of u^{n+1} and the four outermost planes of u^{n+2}
    const bool cta_lo = lk.peer_u[0] != nullptr && Xa < g.X0 + 4;
    const bool cta_hi = lk.peer_u[1] != nullptr && Xb > g.X1 - 4;
    const bool push_lo = cta_lo && !lk.pull, push_hi = cta_hi && !lk.pull;  // pull mode: the neighbours read our planes in place

    // register queues: qU[s % 5][r] = own rows of u^n, stage s (plane Xa-4+s); qR[i % 5][r] = own step-1 results of
    // iteration i (plane Xa-2+i).  The loop is unrolled by 5 so every index is a compile-time constant.
    float4 qU[5][RY], qR[5][RY];
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int r = 0; r < RY; ++r) qR[s][r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        mbar_wait(full0 + 8 * s, 0);
#pragma unroll
        for (int r = 0; r < RY; ++r) qU[s][r] = lds128(sU + s * USLOT_F + ownU + r * HP);
        if (s == 1) {  // stages 0 and 1 are never a centre plane: release them now
            __syncwarp();
            if (lane == 0) mbar_arrive(pro0);
        }
    }

    int p3 = 0, b6 = 0, d8 = 0;   // i % 3, i % 6, i % 8
    int b6c = SB - 2;              // (i - 2) % 6
    uint32_t bpar_c = 1;           // parity of (i-2)/6, valid from i >= 2 (becomes 0 when b6c wraps to 0)
    for (int i0 = 0; i0 < nit; i0 += 5) {
        const uint32_t par = (uint32_t)(i0 / 5) & 1u;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int i = i0 + k;
            if (i >= nit) break;
            // ---------------- step 1: u^{n+1} on plane P1 = Xa-2+i, extended tile
            const int fsl = (k + 4) % 5, csl = (k + 2) % 5;
            // both waits first: the two steps of an iteration are independent instruction streams (step 2 works on
            // the step-1 plane of iteration i-2), so with no barrier operation between them their shared-memory
            // loads and arithmetic overlap
            mbar_wait(full0 + 8 * fsl, par ^ (uint32_t)((k + 4) / 5));
            if (i >= 4) mbar_wait(bfull0 + 8 * b6c, bpar_c);
#pragma unroll
            for (int r = 0; r < RY; ++r) qU[fsl][r] = lds128(sU + fsl * USLOT_F + ownU + r * HP);
            const int P1 = Xa - 2 + i;
            float4 res[RY];  // halo cells keep their value (identical in every level by construction)
#pragma unroll
            for (int r = 0; r < RY; ++r) res[r] = qU[csl][r];
            if (any_inb && P1 >= XC0 && P1 < XC1) {
                const float *P = sU + csl * USLOT_F + ownU;
                // this thread's y column on the centre plane: 2 rows above, own rows (registers), 2 rows below
                float4 col[RY + 4];
                col[0] = lds128(P - 2 * HP);
                col[1] = lds128(P - HP);
                col[RY + 2] = lds128(P + RY * HP);
                col[RY + 3] = lds128(P + (RY + 1) * HP);
#pragma unroll
                for (int r = 0; r < RY; ++r) col[r + 2] = qU[csl][r];
#pragma unroll
                for (int r = 0; r < RY; ++r) {
                    const float2 zl = lds64(P + r * HP - 2), zr = lds64(P + r * HP + 4);
                    const float4 pv = lds128(sP + p3 * CSLOT_F + ownC + r * HP);
                    const float4 mv = lds128(sM + k * CSLOT_F + ownC + r * HP);  // m slot (s-4) % 5 = i % 5 = k
                    float4 v = column4<EXACT>(col[r + 2], qU[k % 5][r], qU[(k + 1) % 5][r], qU[(k + 3) % 5][r], qU[fsl][r], col[r],
                                              col[r + 1], col[r + 3], col[r + 4], zl, zr, pv, mv, a.s.k);
                    if (chunk_has_src) inject_plane(v, P1, Y + r, Z, sv);  // rare: source cells of step n (also in the ghost zone)
                    if (inb[r]) res[r] = v;
                    if (core[r] && P1 >= Xa && P1 < Xb) {
                        *reinterpret_cast<float4 *>(out1 + (long long)P1 * plane + r * g.nzp) = v;
                        if (push_lo && P1 < g.X0 + 2)
                            *reinterpret_cast<float4 *>(lk.peer_u[0] + a.s.l_n1 * lk.peer_lvl[0] + (long long)(lk.peer_edge[0] + P1 - g.X0) * plane + row0 + r * g.nzp) = v;
                        if (push_hi && P1 >= g.X1 - 2)
                            *reinterpret_cast<float4 *>(lk.peer_u[1] + a.s.l_n1 * lk.peer_lvl[1] + (long long)(lk.peer_edge[1] + P1 - g.X1) * plane + row0 + r * g.nzp) = v;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                qR[k % 5][r] = res[r];
                if (active) *reinterpret_cast<float4 *>(sB + b6 * CSLOT_F + ownC + r * HP) = res[r];
            }

            // ---------------- step 2: u^{n+2} on plane X = Xa+i-4 (centre = step-1 plane of iteration i-2)
            if (i >= 4) {
                if (any_core) {
                    const int X = Xa + i - 4;
                    const float *P = sB + b6c * CSLOT_F + ownC;
                    float4 col[RY + 4];
                    col[0] = lds128(P - 2 * HP);
                    col[1] = lds128(P - HP);
                    col[RY + 2] = lds128(P + RY * HP);
                    col[RY + 3] = lds128(P + (RY + 1) * HP);
#pragma unroll
                    for (int r = 0; r < RY; ++r) col[r + 2] = qR[(k + 3) % 5][r];
#pragma unroll
                    for (int r = 0; r < RY; ++r) {
                        const float2 zl = lds64(P + r * HP - 2), zr = lds64(P + r * HP + 4);
                        const float4 mv = lds128(sM + ((k + 3) % 5) * CSLOT_F + ownC + r * HP);  // m of plane X: slot (i-2) % 5
                        // x neighbours and centre from the step-1 queue; "previous" level = u^n on this plane (stage i)
                        float4 o = column4<EXACT>(col[r + 2], qR[(k + 1) % 5][r], qR[(k + 2) % 5][r], qR[(k + 4) % 5][r], qR[k % 5][r], col[r],
                                                  col[r + 1], col[r + 3], col[r + 4], zl, zr, qU[k % 5][r], mv, a.s.k);
                        if (chunk_has_src) inject_plane(o, X, Y + r, Z, sv2);  // source cells of step n+1
                        if (core[r]) {
                            *reinterpret_cast<float4 *>(out2 + (long long)X * plane + r * g.nzp) = o;
                            if (push_lo && X < g.X0 + 4)
                                *reinterpret_cast<float4 *>(lk.peer_u[0] + a.s.l_n2 * lk.peer_lvl[0] + (long long)(lk.peer_edge[0] + X - g.X0) * plane + row0 + r * g.nzp) = o;
                            if (push_hi && X >= g.X1 - 4)
                                *reinterpret_cast<float4 *>(lk.peer_u[1] + a.s.l_n2 * lk.peer_lvl[1] + (long long)(lk.peer_edge[1] + X - g.X1) * plane + row0 + r * g.nzp) = o;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bfull0 + 8 * b6);  // this warp's part of step-1 plane i is in shared memory
                mbar_arrive(done0 + 8 * d8);   // iteration i done: u^n stage i+2, u^{n-1} and m slots reusable
            }

            if (++p3 == SP) p3 = 0;
            if (++b6 == SB) b6 = 0;
            if (++d8 == ND) d8 = 0;
            if (++b6c == SB) {
                b6c = 0;
                bpar_c ^= 1;
            }
        }
    }

    if (cta_lo || cta_hi) {
        // every consumer thread of this CTA has issued its peer stores: count the CTA, and let the last CTA of
        // a boundary publish the pass's epoch in the neighbour's flag (same protocol as stencil_tma.cu)
        asm volatile("bar.sync 1, %0;" ::"r"(T::NC) : "memory");
        if (threadIdx.x == 0) {
            __threadfence_system();
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!(side == 0 ? cta_lo : cta_hi)) continue;
                raise_flag_fenced(lk.peer_tile[side] + blockIdx.x, lk.epoch);  // this tile's boundary is done
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

// ---------------------------------------------------------------------------- halo-shell check / copy
// The shell = every padded cell outside the box [X0,X1) x [Y0,Y1) x [Z0,Z1).  Section0 never writes it, so it
// is constant in time; a two-step pass moves ring levels between physical levels and therefore needs the
// shells of all levels to be the same.  One (x,y) row per threadIdx.y, z strided over threadIdx.x.
// check: *flag |= 1 if the shells of levels 0, 1 and 2 differ anywhere (bit compare).
// copy : shell of level `from` -> level `to`.
template <bool COPY>
__global__ void shell_kernel(float *__restrict__ u, Grid g, int from, int to, int *flag)
{
    const long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    if (row >= (long long)g.nxp * g.nyp) return;
    const int X = (int)(row / g.nyp), Y = (int)(row % g.nyp);
    const bool whole = X < g.X0 || X >= g.X1 || Y < g.Y0 || Y >= g.Y1;
    unsigned *a = reinterpret_cast<unsigned *>(u) + row * g.nzp;
    bool bad = false;
    for (int z = threadIdx.x; z < g.nzp; z += blockDim.x) {
        if (!whole && z >= g.Z0 && z < g.Z1) continue;
        if (COPY) {
            a[(long long)to * g.lvl + z] = a[(long long)from * g.lvl + z];
        } else {
            const unsigned v = a[z];
            bad |= (v != a[g.lvl + z]) || (v != a[2 * g.lvl + z]);
        }
    }
    if (!COPY && bad) atomicOr(flag, 1);
}

int launch_shell_check(float *u, const Grid &g, int *flag, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), stream);
    if (e != cudaSuccess) return (int)e;
    dim3 block(32, 8);
    const long long rows = (long long)g.nxp * g.nyp;
    shell_kernel<false><<<(unsigned)((rows + block.y - 1) / block.y), block, 0, stream>>>(u, g, 0, 0, flag);
    return (int)cudaGetLastError();
}

int launch_shell_copy(float *u, const Grid &g, int from, int to, cudaStream_t stream)
{
    dim3 block(32, 8);
    const long long rows = (long long)g.nxp * g.nyp;
    shell_kernel<true><<<(unsigned)((rows + block.y - 1) / block.y), block, 0, stream>>>(u, g, from, to, nullptr);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------- host side
#define FDTD_TB2_1(ER_, EC_, RY_, EX_) \
    {ER_, EC_, RY_, EX_, stencil_tb2_kernel<ER_, EC_, RY_, EX_>, Tb2Shape<ER_, EC_, RY_>::NT, (size_t)Tb2Shape<ER_, EC_, RY_>::SMEM}
#define FDTD_TB2(ER_, EC_, RY_) FDTD_TB2_1(ER_, EC_, RY_, false), FDTD_TB2_1(ER_, EC_, RY_, true)
static const Tb2Variant g_tb2[] = {
    // extended tile (rows, float4 columns), rows per thread -> output tile (ER-4) x (4*EC-8); first match wins.
    // Two rows per thread halve the barrier / address work per point but leave 12 warps per SM instead of 22: the
    // loop is latency-bound and runs 25% slower (profiles/r01_sweep512_twostep_rows.txt) -- kept for the record.
    FDTD_TB2(36, 18, 1),  // 32 x 64
    FDTD_TB2(32, 18, 1),  // 28 x 64
    FDTD_TB2(20, 34, 1),  // 16 x 128
    FDTD_TB2(20, 18, 1),  // 16 x 64
    FDTD_TB2(36, 18, 2),
    FDTD_TB2(20, 34, 2),
};
static const int g_ntb2 = (int)(sizeof(g_tb2) / sizeof(g_tb2[0]));

int tb2_plan_build(Tb2Plan &p, float *u, const float *m, const Grid &g, const TmaConfig &cfg, bool exact, int sm_count,
                   const SlabLink *link)
{
    p.valid = false;
    if (!tma_supported(g)) return (int)cudaErrorInvalidValue;
    const int ny = g.Y1 - g.Y0, nz = g.Z1 - g.Z0, nx = g.X1 - g.X0;
    // tile: explicit, or 16 x 128 where it divides the grid (measured 1% ahead of 32 x 64 at 512^3: longer rows per
    // TMA box), else 32 x 64
    int want_ty = cfg.ty, want_tz = cfg.tz;
    if (want_ty <= 0 && want_tz <= 0 && ny % 16 == 0 && nz % 128 == 0) want_ty = 16, want_tz = 128;
    // the lean kernel (stencil_tb2l.cu) has one row per thread; two rows per thread exist in the first kernel only
    const bool lean = cfg.lean != 0 && cfg.rows != 2;
    int ntab = g_ntb2;
    const Tb2Variant *tab = g_tb2;
    if (lean) tab = tb2l_variants(&ntab);
    int vi = -1;
    for (int i = 0; i < ntab && vi < 0; ++i)
        if (tab[i].exact == exact && (want_ty <= 0 || tab[i].er - 4 == want_ty) && (want_tz <= 0 || 4 * tab[i].ec - 8 == want_tz) &&
            (cfg.rows <= 0 ? tab[i].rows <= 2 : tab[i].rows == cfg.rows))
            vi = i;
    if (vi < 0) return (int)cudaErrorInvalidValue;
    const Tb2Variant &v = tab[vi];
    const int ty = v.er - 4, tz = 4 * v.ec - 8;

    cuuint64_t dims_u[4] = {(cuuint64_t)g.nzp, (cuuint64_t)g.nyp, (cuuint64_t)g.nxp, (cuuint64_t)FDTD_LEVELS};
    cuuint32_t box_u[4] = {(cuuint32_t)(4 * v.ec), (cuuint32_t)(v.er + 4), 1, 1};
    cuuint32_t box_c[4] = {(cuuint32_t)(4 * v.ec), (cuuint32_t)v.er, 1, 1};
    int rc;
    if ((rc = encode_tensor_map(&p.map_cur, u, 4, dims_u, box_u))) return rc;
    if ((rc = encode_tensor_map(&p.map_prev, u, 4, dims_u, box_c))) return rc;
    if ((rc = encode_tensor_map(&p.map_m, m, 3, dims_u, box_c))) return rc;
    for (int side = 0; side < 2; ++side) {
        p.map_cur_peer[side] = p.map_cur;
        p.map_prev_peer[side] = p.map_prev;
        if (!link || !link->peer_u[side]) continue;
        cuuint64_t dims_p[4] = {(cuuint64_t)g.nzp, (cuuint64_t)g.nyp, (cuuint64_t)link->peer_nxp[side], (cuuint64_t)FDTD_LEVELS};
        if ((rc = encode_tensor_map(&p.map_cur_peer[side], link->peer_u[side], 4, dims_p, box_u))) return rc;
        if ((rc = encode_tensor_map(&p.map_prev_peer[side], link->peer_u[side], 4, dims_p, box_c))) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute((const void *)v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem);
    if (e != cudaSuccess) return (int)e;

    // x chunking: one CTA per SM; several full waves, chunks long enough to amortise the 8-plane prologue
    const int tiles = ((ny + ty - 1) / ty) * ((nz + tz - 1) / tz);
    int xchunk = cfg.xchunk;
    if (xchunk <= 0) {
        double best = -1.0;
        // chunks longer than 128 planes run slower (1024^3, lean kernel: 549 Gpts/s at 64 and 128 planes, 481 at 256, 440 at 512 --
        // profiles/r02_sweep1024_lean.txt: neighbouring tiles drift apart along x and stop sharing their halo rows in L2)
        for (int nch = (nx + 127) / 128; nch <= nx; ++nch) {
            const int xc = (nx + nch - 1) / nch;
            if (xc < 16 && nch > 1) break;
            if ((nx + xc - 1) / xc != nch) continue;
            const double waves = tiles * (double)nch / sm_count, full = ceil(waves);
            const double eff = waves / full * full / (full + 0.25) * xc / (xc + 8.0);
            if (eff > best) {
                best = eff;
                xchunk = xc;
            }
        }
    }
    p.ty = ty;
    p.tz = tz;
    p.rows = v.rows;
    p.xchunk = xchunk;
    p.variant = vi;
    p.lean = lean;
    p.smem_bytes = v.smem;
    p.valid = true;
    return 0;
}

int launch_stencil_tb2(const Tb2Plan &p, const Tb2Step &a, bool exact, cudaStream_t stream)
{
    if (!p.valid) return (int)cudaErrorInvalidValue;
    int ntab = g_ntb2;
    const Tb2Variant *tab = g_tb2;
    if (p.lean) tab = tb2l_variants(&ntab);
    const Tb2Variant &v = tab[p.variant];
    if (v.exact != exact) return (int)cudaErrorInvalidValue;
    const int ny = a.g.Y1 - a.g.Y0, nz = a.g.Z1 - a.g.Z0, nx = a.g.X1 - a.g.X0;
    if (nx <= 0) return 0;
    Tb2Args args;
    args.map_cur = p.map_cur;
    args.map_prev = p.map_prev;
    args.map_m = p.map_m;
    for (int side = 0; side < 2; ++side) {
        args.map_cur_peer[side] = p.map_cur_peer[side];
        args.map_prev_peer[side] = p.map_prev_peer[side];
    }
    args.s = a;
    args.tiles_z = (nz + p.tz - 1) / p.tz;
    args.tiles_y = (ny + p.ty - 1) / p.ty;
    args.xchunk = p.xchunk;
    args.edge = 0;
    // L2 prefetch distance of the lean kernel's producer: 2 stages in contracted arithmetic (509 -> 547 Gpts/s at 512^3), none in
    // exact arithmetic (issue-bound: 413 without, 395 with) -- profiles/r02_sweep512_lean_prefetch.txt
    static const int pf = [] { const char *e = getenv("FDTD_B200_TB2_PREFETCH"); return e ? atoi(e) : -1; }();
    args.prefetch = pf >= 0 ? pf : (exact ? 0 : 2);
    int nchunks = (nx + p.xchunk - 1) / p.xchunk;
    const bool linked = a.link.peer_u[0] != nullptr || a.link.peer_u[1] != nullptr;
    if (linked) {  // short boundary chunks hold the 4 planes a neighbour needs; the usual chunks lie in between
        if (nx < 4 * kSlabEdgePlanes) return (int)cudaErrorInvalidValue;
        if (a.link.tile_mode) {
            // per-tile flags: the usual chunks, but exactly ONE chunk per side may hold boundary planes (>= 4 planes each)
            while (nchunks > 1 && (args.xchunk < 4 || nx - (nchunks - 1) * args.xchunk < 4)) {
                --nchunks;
                args.xchunk = (nx + nchunks - 1) / nchunks;
            }
            nchunks = (nx + args.xchunk - 1) / args.xchunk;
            // lean kernel: a CTA that holds BOTH boundaries runs the loop copy with calls (stencil_tb2l.cu, MODE 2); two chunks keep
            // every CTA of a slab with two neighbours on the one-boundary copy (2048^3 on 8 GPUs: 1320 -> see profiles)
            if (p.lean && nchunks == 1 && a.link.peer_u[0] != nullptr && a.link.peer_u[1] != nullptr && nx >= 16) {
                nchunks = 2;
                args.xchunk = (nx + 1) / 2;
            }
        } else {
            args.edge = slab_edge_planes(nx, p.xchunk, args.tiles_z * args.tiles_y, 0);
            nchunks = 2 + (nx - 2 * args.edge + p.xchunk - 1) / p.xchunk;
        }
        args.s.link.expect[0] = args.s.link.expect[1] = args.tiles_z * args.tiles_y;
    }
    dim3 grid(args.tiles_z * args.tiles_y, nchunks, 1);
    if (grid.y > 65535) return (int)cudaErrorInvalidValue;
    v.fn<<<grid, v.nt, v.smem, stream>>>(args);
    return (int)cudaGetLastError();
}

}  // namespace fdtd
