// fdtd_arith.cuh -- the per-point update of Section0 and the per-cell source sum of Section1.
//
// EXACT = true replays the reference's fp32 operation order (openacc.cpp:102-107) with
// round-to-nearest intrinsics, which the compiler may not contract into FMAs: the result is
// bit-identical to the reference built for the host with -ffp-contract=off.
// EXACT = false is the algebraically equal leapfrog form of the reference's CUDA paths
// (cuda.cu:105: 2*uc - um1 + dt^2*lap/m) with pre-multiplied coefficients, FMA and one MUFU.RCP; it
// differs from the oracle by relative L2 ~1e-6 (tolerance 1e-4, README.md:33).  Neither form flushes
// denormal field values.  Every kernel calls the same point<EXACT>(), so kernels agree bit for bit.
#pragma once
#include "fdtd_common.cuh"

namespace fdtd {

#define FDTD_C2 (-8.33333333e-2F)  // -1/12  (openacc.cpp:104)
#define FDTD_C1 (1.333333330F)     //  4/3
#define FDTD_C0 (-2.50F)           // -5/2

// (r5 + c2*(u[-2] + u[+2])) + c1*(u[-1] + u[+1])  -- one axis of openacc.cpp:104-106, reference order
__device__ __forceinline__ float axis_term_exact(float r5, float m2, float m1, float p1, float p2)
{
    return __fadd_rn(__fadd_rn(r5, __fmul_rn(FDTD_C2, __fadd_rn(m2, p2))), __fmul_rn(FDTD_C1, __fadd_rn(m1, p1)));
}

// dt*dt*( r2*dx + r3*dy + r4*dz - ((-2*r1)*u0 + r1*u1)*m ) / m  -- openacc.cpp:103-107, reference order
__device__ __forceinline__ float leapfrog_exact(float c, float dx, float dy, float dz, float u1, float m, const Coef &k)
{
    const float lap = __fadd_rn(__fadd_rn(__fmul_rn(k.r2, dx), __fmul_rn(k.r3, dy)), __fmul_rn(k.r4, dz));
    const float d = __fmul_rn(__fadd_rn(__fmul_rn(k.n2r1, c), __fmul_rn(k.r1, u1)), m);
    const float num = __fmul_rn(k.dt2, __fsub_rn(lap, d));
    // (+-0)/m == (+-0)*m for finite m != 0: skip the IEEE division where the field is still zero
    // (its FCHK guard sends zero dividends to the slow path); warp-uniform in quiescent regions.
    if (num == 0.0f) return __fmul_rn(num, m);
    // Tiny (incl. denormal) numerators -- the fringe of the wavefield -- would take the fp32 division's
    // software slow path.  Dividing in fp64 and rounding once more to fp32 is still the correctly rounded
    // fp32 quotient (double rounding is innocuous for division when p' >= 2p+2: 53 >= 50), and fp32
    // denormals are normal fp64 numbers, so this branch has no slow path.
    if (fabsf(num) < 0x1p-80f) return __double2float_rn(__ddiv_rn((double)num, (double)m));
    return __fdiv_rn(num, m);
}

// One output point.  x*/y*/z* are the radius-2 neighbours along each axis.
template <bool EXACT>
__device__ __forceinline__ float point(float c, float xm2, float xm1, float xp1, float xp2, float ym2, float ym1,
                                       float yp1, float yp2, float zm2, float zm1, float zp1, float zp2, float u1,
                                       float m, const Coef &k)
{
    if (EXACT) {
        const float r5 = __fmul_rn(FDTD_C0, c);
        const float dx = axis_term_exact(r5, xm2, xm1, xp1, xp2);
        const float dy = axis_term_exact(r5, ym2, ym1, yp1, yp2);
        const float dz = axis_term_exact(r5, zm2, zm1, zp1, zp2);
        return leapfrog_exact(c, dx, dy, dz, u1, m, k);
    } else {
        // minimal-operation form: dt^2*lap accumulated with pre-multiplied coefficients, one MUFU.RCP
        float acc = k.f0 * c;
        acc = fmaf(k.fx2, xm2 + xp2, acc);
        acc = fmaf(k.fx1, xm1 + xp1, acc);
        acc = fmaf(k.fy2, ym2 + yp2, acc);
        acc = fmaf(k.fy1, ym1 + yp1, acc);
        acc = fmaf(k.fz2, zm2 + zp2, acc);
        acc = fmaf(k.fz1, zm1 + zp1, acc);
        float rm;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rm) : "f"(m));
        return fmaf(acc, rm, fmaf(2.0f, c, -u1));
    }
}

// Value one source adds to one of its corner cells at this step (openacc.cpp:134):
// (w * src[time][p]) / m[base corner], w = ((1e-2f*wx)*wy)*wz.
__device__ __forceinline__ float source_term(const SourceContrib &sc, const float *__restrict__ src_row,
                                             const float *__restrict__ mbase)
{
    return __fdiv_rn(__fmul_rn(sc.w, src_row[sc.p]), mbase[sc.p]);
}

// Sequentially add every contribution of `cell` to v, in p_src order (the serial order of
// openacc.cpp:116-136), one rounding per addition.
__device__ __forceinline__ float apply_cell(float v, const SourceCell &cell, const SourceView &sv)
{
    for (int i = 0; i < cell.count; ++i)
        v = __fadd_rn(v, source_term(sv.contribs[cell.first + i], sv.src_row, sv.mbase));
    return v;
}

}  // namespace fdtd
