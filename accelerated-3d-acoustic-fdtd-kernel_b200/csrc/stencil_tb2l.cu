// stencil_tb2l.cu -- the two-step pass of stencil_tb2.cu with a LEAN inner loop (same inputs, outputs, level rotation, slab
// protocol and bits; option "tb2_lean", default on).
//
// Why: stencil_tb2_kernel is bound by instruction issue, not by memory (profiles/r02_ncu_tb2_source_counters.txt: 382 warp
// instructions per warp and x plane, 112 of them arithmetic; DRAM 59 % busy).  Its SASS recomputes the shared-memory base
// (S2R SR_CgaCtaId + LEA), both 64-bit global store addresses (a 25-instruction IMAD chain each), the slab-link predicates and
// four ring counters in EVERY iteration, because at 80 registers the compiler rematerialises rather than keeps.  This kernel
// removes the causes instead of the symptoms:
//   * every ring (u^n, u^{n-1}, m, step-1 planes, barriers) is 5 deep, the loop is unrolled by 5, so every slot index is a
//     compile-time constant and every shared-memory access is `[sb + immediate]` off ONE per-thread base register (inline PTX,
//     so the base cannot be rematerialised);
//   * one `done` barrier per iteration instead of two: "all warps finished iteration i" is both what frees the TMA slots
//     (producer) and what makes step-1 plane i readable (consumers);
//   * one running 64-bit pointer for both global stores (u^{n+2} = the same pointer + a launch-constant offset);
//   * the first group of 5 iterations (no step 2 before i = 4, no stores before i = 2) is peeled, so the steady-state body has
//     two integer compares; step 1 is computed unconditionally and de-selected per thread where it must not apply (halo
//     cells, planes beyond a physical boundary) instead of being branched around;
//   * the rare work is compiled into separate copies of the loop, chosen per CTA: source cells inside the TILE (found by a
//     pre-pass over the chunk's cells; a non-inlined call per step, on exactly the planes that hold such a cell), peer stores of
//     a boundary CTA in the push protocol (the neighbour's address is the own one + a launch constant); the common copy has
//     neither, so they cost no registers and no instruction-cache footprint in the steady state;
//   * the TMA producer prefetches a few stages ahead into L2 (cp.async.bulk.prefetch.tensor): DRAM latency and its jitter are
//     hidden without a deeper shared-memory ring.
// Linked slabs pull by default with this kernel (the neighbours' TMA producers read the boundary planes in place): its warps
// advance in lock step, so a stalling peer store holds up the whole tile (fdtd_plan.cu, plan_prepare).
// Arithmetic is the shared column4<EXACT>() => bit-identical to stencil_tb2_kernel and to two one-step launches
// (tests/test_tb2_gpu.py runs every case through both kernels).  512^3 on one B200: 542-549 Gpts/s contracted (first kernel
// 452-465), 410-416 bit-exact (245); DESIGN.md 4.5.
#include "stencil_tb2.cuh"

#include <type_traits>

namespace fdtd {

template <int ER, int EC>
struct Tb2LShape {
    static constexpr int TY = ER - 4, TZ = 4 * EC - 8;  // output tile
    static constexpr int NCA = ER * EC;                 // consumer threads in use: one float4 column each
    static constexpr int NC = (NCA + 31) / 32 * 32;
    static constexpr int NCW = NC / 32;
    static constexpr int NT = NC + 32;                  // + producer warp
    static constexpr int HP = 4 * EC;                   // pitch of every slot (floats)
    static constexpr int D = 5;                         // depth of every ring
    static constexpr int UBYTES = (ER + 4) * HP * 4;
    static constexpr int USLOT = (UBYTES + 127) / 128 * 128;
    static constexpr int CBYTES = ER * HP * 4;
    static constexpr int CSLOT = (CBYTES + 127) / 128 * 128;
    static constexpr int OFF_U = 128;  // guard: column 0 reads two floats to its left
    static constexpr int OFF_P = OFF_U + D * USLOT;
    static constexpr int OFF_M = OFF_P + D * CSLOT;
    static constexpr int OFF_B = OFF_M + D * CSLOT;
    // full[5], done[5], pro -- behind two rows of slack: in step 2 the ghost-row threads of a warp that also holds rows of the
    // output tile read two rows past their slot (values they never use)
    static constexpr int OFF_BAR = OFF_B + D * CSLOT + 2 * HP * 4;
    static constexpr int SMEM = OFF_BAR + 128;
    static_assert(NT <= 1024, "too many threads");
    static_assert(SMEM <= 232448, "shared memory of one CTA exceeds 227 KB");
};

constexpr int kSrcMaskWords = 64;  // planes per chunk the per-tile source-plane mask covers (32 per word)

// ---- shared-memory accesses as [register + immediate]
template <int OFF>
__device__ __forceinline__ float4 lds4(uint32_t base)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ float2 lds2(uint32_t base)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(base), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts4(uint32_t base, const float4 &v)
{
    asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(base), "n"(OFF), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int OFF>
__device__ __forceinline__ void mbar_wait_at(uint32_t base, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0+%1], %2, %3;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(base),
        "n"(OFF), "r"(parity), "r"(0x989680)
        : "memory");
}
template <int OFF>
__device__ __forceinline__ void mbar_arrive_at(uint32_t base)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+%1];" ::"r"(base), "n"(OFF) : "memory");
}

// one elected lane of the (converged) warp arrives
template <int OFF>
__device__ __forceinline__ void elect_arrive_at(uint32_t base)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0+%1];\n"
        "}\n" ::"r"(base),
        "n"(OFF)
        : "memory");
}
__device__ __forceinline__ void stg4(unsigned long long addr, const float4 &v)
{
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The rare work of one step of one thread, kept out of line: source cells of plane X that fall into this float4 (p_src
// order, as inject_plane), then -- for a CTA on a slab boundary -- the copy of a boundary plane into the neighbour's ghost
// plane (2 planes of u^{n+1}, 4 planes of u^{n+2} per side).  Returns the (possibly injected) value.
__device__ __noinline__ float4 tb2l_rare(float4 v, const Tb2Args *a, int step, int X, int Y, int Z, int store)
{
    SourceView sv = a->s.sv;
    if (step) sv.src_row = a->s.src_row2;
    if (sv.ncells > 0) inject_plane(v, X, Y, Z, sv);
    const SlabLink &lk = a->s.link;
    if (store && !lk.pull) {
        const Grid &g = a->s.g;
        const int depth = step ? 4 : 2, lvl = step ? a->s.l_n2 : a->s.l_n1;
        const long long plane = (long long)g.nyp * g.nzp, row0 = (long long)Y * g.nzp + Z;
        if (lk.peer_u[0] && X < g.X0 + depth)
            *reinterpret_cast<float4 *>(lk.peer_u[0] + lvl * lk.peer_lvl[0] + (long long)(lk.peer_edge[0] + X - g.X0) * plane + row0) = v;
        if (lk.peer_u[1] && X >= g.X1 - depth)
            *reinterpret_cast<float4 *>(lk.peer_u[1] + lvl * lk.peer_lvl[1] + (long long)(lk.peer_edge[1] + X - g.X1) * plane + row0) = v;
    }
    return v;
}

// MODE 0: neither source cells in the chunk nor a slab boundary; 1: a CTA on ONE boundary of a linked slab (copies its boundary
// planes into that neighbour's ghost planes, whose address is the own one + a launch constant); 2: source cells in the chunk, or
// both boundaries in one chunk (tb2l_rare() per step).
template <int ER, int EC, bool EXACT, int MODE>
__device__ __forceinline__ void tb2l_consume(const Tb2Args &a, const uint32_t s0, const int Xa, const int Xb, const int Yt, const int Zt,
                                             const int XC0, const int XC1, const int xs_lo, const int xs_hi, const unsigned *srcmask)
{
    using T = Tb2LShape<ER, EC>;
    constexpr int HP = T::HP;
    constexpr int FULL = T::OFF_BAR, DONE = T::OFF_BAR + 8 * T::D, PRO = T::OFF_BAR + 16 * T::D;
    const Grid &g = a.s.g;
    const bool live = threadIdx.x < T::NCA;
    const int t = live ? threadIdx.x : 0;
    const int er = t / EC, ec = t % EC;
    const int Y = Yt - 2 + er, Z = Zt - 4 + 4 * ec;
    const bool inb = Z >= g.Z0 && Z < g.Z1 && Y >= g.Y0 && Y < g.Y1;                     // interior in (y,z): whole float4
    const bool core = live && inb && er >= 2 && er < ER - 2 && ec >= 1 && ec < EC - 1;  // a point of the output tile
    const bool warp_core = __any_sync(0xffffffffu, core);                               // warps of ghost rows skip step 2

    // Per-thread base: own float4 in U slot 0.  Everything else in shared memory is a compile-time offset from it:
    //   U(slot, drow, dfl)  u^n ring, rows start at Yt-4;   C(ring, slot, drow, dfl)  u^{n-1} / m / step-1 rings, rows start at Yt-2
    // The launch constants the loop needs are made opaque (asm volatile), so that the compiler keeps them instead of sinking
    // their computation into every iteration.
    uint32_t sb;
    asm volatile("mov.u32 %0, %1;" : "=r"(sb) : "r"(s0 + T::OFF_U + ((er + 2) * HP + 4 * ec) * 4));
#define TB2L_U(slot, drow, dfl) ((slot) * T::USLOT + ((drow) * HP + (dfl)) * 4)
#define TB2L_C(ring, slot, drow, dfl) ((ring) - T::OFF_U - 2 * HP * 4 + (slot) * T::CSLOT + ((drow) * HP + (dfl)) * 4)
    const long long plane = (long long)g.nyp * g.nzp;
    // p1 -> this thread's float4 of u^{n+1} on the step-1 plane of the current iteration (bytes); u^{n+2} of the same iteration
    // (two planes behind, another level) lies d2 bytes further
    unsigned long long p1, stride, d2;
    asm volatile("mov.b64 %0, %1;" : "=l"(p1) : "l"(a.s.u + (long long)a.s.l_n1 * g.lvl + (long long)(Xa - 2) * plane + (long long)Y * g.nzp + Z));
    asm volatile("mov.b64 %0, %1;" : "=l"(stride) : "l"(plane * 4));
    asm volatile("mov.b64 %0, %1;" : "=l"(d2) : "l"(((long long)(a.s.l_n2 - a.s.l_n1) * g.lvl - 2 * plane) * 4));
    // step 1 is stored while more than 2 iterations remain (planes [Xa, Xb)), applies from iteration ilo on (planes >= XC0) and
    // while more than `cut` iterations remain (planes < XC1)
    const int np = Xb - Xa;
    int rem, cut, ilo;
    asm volatile("mov.u32 %0, %1;" : "=r"(rem) : "r"(np + 4));
    asm volatile("mov.u32 %0, %1;" : "=r"(cut) : "r"(Xb + 2 - XC1));
    asm volatile("mov.u32 %0, %1;" : "=r"(ilo) : "r"(XC0 - Xa + 2));
    int X1 = Xa - 2;  // step-1 plane of the current iteration (MODE >= 1 only)
    // MODE 1, peer stores: the neighbour's copy of a boundary plane lies a launch-constant number of bytes away from this slab's
    // own (dq1: u^{n+1} relative to p1, dq2: u^{n+2} relative to p1).  With W = X1 towards the lower neighbour and -X1 towards the
    // upper one, step 1 is copied while W < thr1 and step 2 while W < thr2.
    const SlabLink &lk = a.s.link;
    // MODE 2: a boundary CTA (one side with source cells, or both sides in one chunk) copies its boundary planes inside tb2l_rare()
    const bool pushes = MODE == 2 && !lk.pull && ((lk.peer_u[0] != nullptr && Xa < g.X0 + 4) || (lk.peer_u[1] != nullptr && Xb > g.X1 - 4));
    // MODE 2: does plane X hold a source cell inside this tile?  (a 64-source lattice puts cells into many tiles, but into few
    // planes of each: asking per plane keeps the call out of almost every iteration)
    auto plane_has_src = [&](int X) -> bool {
        if (X < xs_lo || X > xs_hi) return false;
        const int j = X - (Xa - 2);
        return j >= 32 * kSrcMaskWords || ((srcmask[j >> 5] >> (j & 31)) & 1u);
    };
    unsigned long long dq1 = 0, dq2 = 0;
    int W = 0, wstep = 0, thr1 = 0, thr2 = 0;
    if (MODE == 1) {
        const int side = (lk.peer_u[0] != nullptr && Xa < g.X0 + 4) ? 0 : 1;
        const unsigned long long e = (unsigned long long)lk.peer_u[side] - (unsigned long long)a.s.u +
                                     (unsigned long long)((long long)(lk.peer_edge[side] - (side ? g.X1 : g.X0)) * plane * 4);
        asm volatile("mov.b64 %0, %1;" : "=l"(dq1) : "l"(e + (unsigned long long)(a.s.l_n1 * (lk.peer_lvl[side] - g.lvl) * 4)));
        asm volatile("mov.b64 %0, %1;" : "=l"(dq2) : "l"(e + (unsigned long long)(a.s.l_n2 * (lk.peer_lvl[side] - g.lvl) * 4) + d2));
        asm volatile("mov.u32 %0, %1;" : "=r"(wstep) : "r"(side ? -1 : 1));
        asm volatile("mov.u32 %0, %1;" : "=r"(thr1) : "r"(side ? 3 - g.X1 : g.X0 + 2));
        asm volatile("mov.u32 %0, %1;" : "=r"(thr2) : "r"(side ? 3 - g.X1 : g.X0 + 6));
        W = side ? -X1 : X1;
    }

    // register queues: qU[s % 5] = own float4 of u^n, stage s (plane Xa-4+s); qR[i % 5] = own step-1 result of iteration i
    float4 qU[5], qR[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) qR[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    mbar_wait_at<FULL + 0>(s0, 0);
    qU[0] = lds4<TB2L_U(0, 0, 0)>(sb);
    mbar_wait_at<FULL + 8>(s0, 0);
    qU[1] = lds4<TB2L_U(1, 0, 0)>(sb);
    __syncwarp();
    elect_arrive_at<PRO>(s0);  // stages 0 and 1 are never a centre plane: release them now
    mbar_wait_at<FULL + 16>(s0, 0);
    qU[2] = lds4<TB2L_U(2, 0, 0)>(sb);
    mbar_wait_at<FULL + 24>(s0, 0);
    qU[3] = lds4<TB2L_U(3, 0, 0)>(sb);

    // The two floats left (zl) and right (zr) of the own float4: 8-byte loads at a 16-byte lane stride touch half of the banks and
    // take 4 wavefronts instead of 2 -- all of the kernel's bank conflicts (10.7 M of 94 M wavefronts per pass,
    // profiles/r02_ncu_tb2l_16x128_exact0.txt).  Letting odd groups of 8 lanes read zr first and zl second makes both loads
    // conflict-free, but the selects and the two extra base registers cost what the wavefronts gain (539 vs 542 Gpts/s,
    // profiles/r02_sweep512_zswap.txt): not kept.
    auto load_z = [&](float2 &zl, float2 &zr, auto offc) __attribute__((always_inline)) {
        constexpr int OFF = decltype(offc)::value;
        zl = lds2<OFF - 8>(sb);
        zr = lds2<OFF + 16>(sb);
    };

    // Iteration i (k = i % 5): new u^n stage i+4 -> slot (k+4)%5; centre plane of step 1 = stage i+2 -> slot (k+2)%5; u^{n-1}, m and
    // the step-1 result of iteration i -> slot k; step 2 works on the step-1 plane of iteration i-2 -> slot (k+3)%5.  The first
    // five iterations are peeled (I0 = 0), then the loop is unrolled by ten (i = 5 + 10*G + J) so that the barrier parities
    // are compile-time constants too.
    auto body = [&](auto jc, auto firstc, auto pushc) __attribute__((always_inline)) {
        constexpr bool FIRST = decltype(firstc)::value;
        constexpr bool PUSH = MODE == 1 && decltype(pushc)::value;  // this group of iterations may hold planes the neighbour needs
        constexpr int J = decltype(jc)::value, I = (FIRST ? 0 : 5) + J, k = I % 5;  // I = i modulo 10
        constexpr bool STEP2 = !FIRST || J == 4;
        constexpr int fsl = (k + 4) % 5, csl = (k + 2) % 5, bsl = (k + 3) % 5;
        // both waits first: the two steps of an iteration are independent instruction streams
        mbar_wait_at<FULL + 8 * fsl>(s0, ((I + 4) / 5) & 1);
        if (STEP2) mbar_wait_at<DONE + 8 * bsl>(s0, ((I - 2) / 5) & 1);  // all warps have finished iteration i-2
        qU[fsl] = lds4<TB2L_U(fsl, 0, 0)>(sb);
        // ---------------- step 1: u^{n+1} on plane Xa-2+i, extended tile
        float4 res;
        {
            const float4 ym2 = lds4<TB2L_U(csl, -2, 0)>(sb), ym1 = lds4<TB2L_U(csl, -1, 0)>(sb);
            const float4 yp1 = lds4<TB2L_U(csl, 1, 0)>(sb), yp2 = lds4<TB2L_U(csl, 2, 0)>(sb);
            float2 zl, zr;
            load_z(zl, zr, std::integral_constant<int, TB2L_U(csl, 0, 0)>{});
            const float4 pv = lds4<TB2L_C(T::OFF_P, k, 0, 0)>(sb), mv = lds4<TB2L_C(T::OFF_M, k, 0, 0)>(sb);
            float4 v = column4<EXACT>(qU[csl], qU[k], qU[(k + 1) % 5], qU[(k + 3) % 5], qU[fsl], ym2, ym1, yp1, yp2, zl, zr, pv, mv, a.s.k);
            const bool st = core && (FIRST ? J >= 2 : true) && rem > 2;
            if (MODE == 2) {
                if (plane_has_src(X1) || (pushes && (X1 < g.X0 + 2 || X1 >= g.X1 - 2))) v = tb2l_rare(v, &a, 0, X1, Y, Z, st ? 1 : 0);
            }
            if (st) stg4(p1, v);
            if (PUSH) {
                if (W < thr1 && st) stg4(p1 + dq1, v);
            }
            // halo cells and planes beyond a physical boundary keep their value (identical in every level by construction)
            const bool apply = inb && (FIRST && J < 2 ? J >= ilo : true) && rem > cut;
            res.x = apply ? v.x : qU[csl].x;
            res.y = apply ? v.y : qU[csl].y;
            res.z = apply ? v.z : qU[csl].z;
            res.w = apply ? v.w : qU[csl].w;
        }
        qR[k] = res;
        sts4<TB2L_C(T::OFF_B, k, 0, 0)>(sb, res);
        // ---------------- step 2: u^{n+2} on plane Xa+i-4 (centre = step-1 plane of iteration i-2)
        if (STEP2) {
            if (warp_core) {
                const float4 ym2 = lds4<TB2L_C(T::OFF_B, bsl, -2, 0)>(sb), ym1 = lds4<TB2L_C(T::OFF_B, bsl, -1, 0)>(sb);
                const float4 yp1 = lds4<TB2L_C(T::OFF_B, bsl, 1, 0)>(sb), yp2 = lds4<TB2L_C(T::OFF_B, bsl, 2, 0)>(sb);
                float2 zl, zr;
                load_z(zl, zr, std::integral_constant<int, TB2L_C(T::OFF_B, bsl, 0, 0)>{});
                const float4 mv = lds4<TB2L_C(T::OFF_M, bsl, 0, 0)>(sb);  // m of plane Xa+i-4: loaded with iteration i-2
                // x neighbours and centre from the step-1 queue; "previous" level = u^n on this plane (stage i, slot k)
                float4 o = column4<EXACT>(qR[bsl], qR[(k + 1) % 5], qR[(k + 2) % 5], qR[(k + 4) % 5], qR[k], ym2, ym1, yp1, yp2, zl, zr, qU[k], mv, a.s.k);
                if (MODE == 2) {
                    if (plane_has_src(X1 - 2) || (pushes && (X1 < g.X0 + 6 || X1 >= g.X1 - 2))) o = tb2l_rare(o, &a, 1, X1 - 2, Y, Z, core ? 1 : 0);
                }
                if (core) stg4(p1 + d2, o);
                if (PUSH) {
                    if (W < thr2 && core) stg4(p1 + dq2, o);
                }
            }
        }
        __syncwarp();
        elect_arrive_at<DONE + 8 * k>(s0);  // this warp's part of step-1 plane i is in shared memory; its reads of iteration i are done
        asm volatile("add.s64 %0, %0, %1;" : "+l"(p1) : "l"(stride));
        --rem;
        ++X1;
        W += wstep;
    };
    using std::integral_constant;
    typedef integral_constant<bool, true> Yes;
    typedef integral_constant<bool, false> No;
    body(integral_constant<int, 0>{}, Yes{}, Yes{});  // rem = np + 4 >= 5
    body(integral_constant<int, 1>{}, Yes{}, Yes{});
    body(integral_constant<int, 2>{}, Yes{}, Yes{});
    body(integral_constant<int, 3>{}, Yes{}, Yes{});
    body(integral_constant<int, 4>{}, Yes{}, Yes{});
    // ten iterations; true when the chunk is finished
    auto group = [&](auto pushc) __attribute__((always_inline)) -> bool {
        if (rem <= 0) return true;
        body(integral_constant<int, 0>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 1>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 2>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 3>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 4>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 5>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 6>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 7>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 8>{}, No{}, pushc);
        if (rem <= 0) return true;
        body(integral_constant<int, 9>{}, No{}, pushc);
        return false;
    };
    for (;;) {
        // MODE 1: only the groups at the slab boundary (the first after the peeled one towards the lower neighbour, the last one or
        // two towards the upper one) carry the peer stores; the others are the common loop
        if (MODE == 1 && min(W, W + 9 * wstep) < thr2) {
            if (group(Yes{})) break;
        } else {
            if (group(No{})) break;
        }
    }
#undef TB2L_U
#undef TB2L_C
}

template <int ER, int EC, bool EXACT>
__global__ void __launch_bounds__(Tb2LShape<ER, EC>::NT, 1) stencil_tb2l_kernel(const __grid_constant__ Tb2Args a)
{
    using T = Tb2LShape<ER, EC>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t s0;  // shared-window address of the CTA's dynamic shared memory, opaque to the compiler (one register, never recomputed)
    asm volatile("mov.u32 %0, %1;" : "=r"(s0) : "r"(smem_u32(smem)));
    constexpr int FULL = T::OFF_BAR, DONE = T::OFF_BAR + 8 * T::D, PRO = T::OFF_BAR + 16 * T::D;

    const Grid &g = a.s.g;
    const SlabLink &lk = a.s.link;
    const int tz = blockIdx.x % a.tiles_z, ty = blockIdx.x / a.tiles_z;
    // chunk order and the short boundary chunks: as in stencil_tb2.cu / stencil_tma.cu
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
    const int XC0 = g.X0 - (lk.peer_u[0] ? 2 : 0), XC1 = g.X1 + (lk.peer_u[1] ? 2 : 0);

    // Which planes of this chunk (incl. its ghost-zone planes) hold source cells inside this tile's extended (y,z) range?  Almost
    // always none: then the CTA runs the copy of the loop without injection.
    __shared__ int s_xsrc[2];
    __shared__ unsigned s_srcmask[kSrcMaskWords];  // bit j: plane Xa-2+j holds such a cell (chunks longer than the mask: every plane is asked)
    if (threadIdx.x < kSrcMaskWords) s_srcmask[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < T::D; ++i) mbar_init(s0 + FULL + 8 * i, 1);
        for (int i = 0; i < T::D; ++i) mbar_init(s0 + DONE + 8 * i, T::NCW);
        mbar_init(s0 + PRO, T::NCW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_xsrc[0] = 0x7fffffff;
        s_xsrc[1] = -1;
    }
    __syncthreads();
    if (a.s.sv.ncells > 0) {
        const int c0 = a.s.sv.plane_off[max(Xa - 2, 0)], c1 = a.s.sv.plane_off[min(Xb + 2, g.nxp)];
        for (int q = c0 + (int)threadIdx.x; q < c1; q += T::NT) {
            const SourceCell cell = a.s.sv.cells[q];
            if (cell.Y >= Yt - 2 && cell.Y < Yt + T::TY + 2 && cell.Z >= Zt - 4 && cell.Z < Zt + T::TZ + 4) {
                atomicMin(&s_xsrc[0], cell.X);
                atomicMax(&s_xsrc[1], cell.X);
                const int j = cell.X - (Xa - 2);
                if (j < 32 * kSrcMaskWords) atomicOr(&s_srcmask[j >> 5], 1u << (j & 31));
            }
        }
        __syncthreads();
    }
    const int xs_lo = s_xsrc[0], xs_hi = s_xsrc[1];

    const int nit = np + 4;  // iterations: step-1 planes Xa-2 .. Xb+1
    if (threadIdx.x >= T::NC) {
        // ------------------------------------------------------------------ producer (one thread)
        // stage s: u^n plane Xa-4+s -> U slot s % 5; for s >= 4 also u^{n-1} and m of plane Xa-6+s (= step-1 plane of iteration
        // s-4) -> P / M slot (s-4) % 5; all on full[s % 5]
        const bool tiled = lk.tile_mode && lk.wait;  // per-tile flags: see stencil_tma.cu
        if (tiled) {
            if (lk.peer_u[0] && Xa - 4 < g.X0) wait_tiles(lk.my_tile[0], ty, tz, a.tiles_y, a.tiles_z, lk.epoch - 1, lk.err);
            if (lk.peer_u[1] && Xb + 4 > g.X1) wait_tiles(lk.my_tile[1], ty, tz, a.tiles_y, a.tiles_z, lk.epoch - 1, lk.err);
        }
        if (threadIdx.x == T::NC) {
            const int nst = nit + 4;
            int us = 0, cs = 0;
            bool waited[2] = {tiled || !(lk.wait && lk.peer_u[0]), tiled || !(lk.wait && lk.peer_u[1])};
            // L2 prefetch of stage t (own planes only: a neighbour's ghost planes may not have been written yet)
            auto prefetch = [&](int t) {
                const int Xt = Xa - 4 + t;
                if (Xt >= g.X0 && Xt < g.X1) tma_prefetch_4d(&a.map_cur, Zt - 4, Yt - 4, Xt, a.s.l_cur);
                if (Xt - 2 >= g.X0 && Xt - 2 < g.X1) {
                    tma_prefetch_4d(&a.map_prev, Zt - 4, Yt - 2, Xt - 2, a.s.l_prev);
                    tma_prefetch_3d(&a.map_m, Zt - 4, Yt - 2, Xt - 2);
                }
            };
            const int pf = a.prefetch;
            for (int t = 5; t < 5 + pf && t < nst; ++t) prefetch(t);
            for (int s = 0; s < nst; ++s) {
                if (pf > 0 && s >= 5 && s + pf < nst) prefetch(s + pf);
                const int Xp = Xa - 4 + s;  // ghost planes (outside [X0, X1)) are written by the neighbours' previous pass
                const int side = Xp < g.X0 ? 0 : (Xp >= g.X1 ? 1 : -1);
                if (side >= 0 && !waited[side]) {
                    wait_flag(lk.my_flag[side], lk.epoch - 1, lk.err);
                    waited[side] = true;
                }
                // slot reuse: stage s overwrites u^n stage s-5 (last read as the centre plane of iteration s-7) and the u^{n-1} / m
                // slots of iteration s-9 (m is last read by step 2 of iteration s-7).  Stages 5 and 6 overwrite stages 0 and 1,
                // which only the prologue reads.
                if (s == 5) mbar_wait(s0 + PRO, 0);
                if (s >= 7) mbar_wait(s0 + DONE + 8 * ((s - 7) % T::D), ((s - 7) / T::D) & 1);
                const uint32_t bar = s0 + FULL + 8 * us;
                const bool ctr = s >= 4;
                mbar_expect_tx(bar, T::UBYTES + (ctr ? 2 * T::CBYTES : 0));
                if (lk.pull && side >= 0 && lk.peer_u[side])  // a neighbour's plane, read where it lies
                    tma_load_4d(s0 + T::OFF_U + us * T::USLOT, &a.map_cur_peer[side], bar, Zt - 4, Yt - 4,
                                lk.peer_edge[side] + Xp - (side == 0 ? g.X0 : g.X1), a.s.l_cur);
                else
                    tma_load_4d(s0 + T::OFF_U + us * T::USLOT, &a.map_cur, bar, Zt - 4, Yt - 4, Xp, a.s.l_cur);
                if (ctr) {
                    const int Xq = Xp - 2, sq = Xq < g.X0 ? 0 : (Xq >= g.X1 ? 1 : -1);
                    if (lk.pull && sq >= 0 && lk.peer_u[sq])
                        tma_load_4d(s0 + T::OFF_P + cs * T::CSLOT, &a.map_prev_peer[sq], bar, Zt - 4, Yt - 2,
                                    lk.peer_edge[sq] + Xq - (sq == 0 ? g.X0 : g.X1), a.s.l_prev);
                    else
                        tma_load_4d(s0 + T::OFF_P + cs * T::CSLOT, &a.map_prev, bar, Zt - 4, Yt - 2, Xq, a.s.l_prev);
                    tma_load_3d(s0 + T::OFF_M + cs * T::CSLOT, &a.map_m, bar, Zt - 4, Yt - 2, Xq);
                    if (++cs == T::D) cs = 0;
                }
                if (++us == T::D) us = 0;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    // the rare work of a CTA: source cells in its chunk (incl. the ghost-zone planes), peer stores of a boundary CTA.  Each kind of
    // CTA runs its own copy of the loop; the common one has neither calls nor peer stores.
    const bool has_src = xs_lo <= xs_hi;
    const bool cta_lo = lk.peer_u[0] != nullptr && Xa < g.X0 + 4;
    const bool cta_hi = lk.peer_u[1] != nullptr && Xb > g.X1 - 4;
    if (has_src || (cta_lo && cta_hi && !lk.pull))
        tb2l_consume<ER, EC, EXACT, 2>(a, s0, Xa, Xb, Yt, Zt, XC0, XC1, xs_lo, xs_hi, s_srcmask);
    else if ((cta_lo || cta_hi) && !lk.pull)
        tb2l_consume<ER, EC, EXACT, 1>(a, s0, Xa, Xb, Yt, Zt, XC0, XC1, 0, -1, s_srcmask);
    else
        tb2l_consume<ER, EC, EXACT, 0>(a, s0, Xa, Xb, Yt, Zt, XC0, XC1, 0, -1, s_srcmask);

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

#define FDTD_TB2L_1(ER_, EC_, EX_) \
    {ER_, EC_, 1, EX_, stencil_tb2l_kernel<ER_, EC_, EX_>, Tb2LShape<ER_, EC_>::NT, (size_t)Tb2LShape<ER_, EC_>::SMEM}
#define FDTD_TB2L(ER_, EC_) FDTD_TB2L_1(ER_, EC_, false), FDTD_TB2L_1(ER_, EC_, true)
static const Tb2Variant g_tb2l[] = {
    // extended tile (rows, float4 columns) -> output tile (ER-4) x (4*EC-8); first match wins
    FDTD_TB2L(36, 18),  // 32 x 64
    FDTD_TB2L(32, 18),  // 28 x 64
    FDTD_TB2L(20, 34),  // 16 x 128
    FDTD_TB2L(20, 18),  // 16 x 64
};
const Tb2Variant *tb2l_variants(int *n)
{
    *n = (int)(sizeof(g_tb2l) / sizeof(g_tb2l[0]));
    return g_tb2l;
}

}  // namespace fdtd
