// fdtd_arith.cuh -- the per-point update of Section0 and the per-cell source sum of Section1.
//
// EXACT = true replays the reference's fp32 operation order (openacc.cpp:102-107) with
// round-to-nearest intrinsics, which the compiler may not contract into FMAs: the result is
// bit-identical to the reference built for the host with -ffp-contract=off.
// EXACT = false is the algebraically equal leapfrog form of the reference's CUDA paths
// (cuda.cu:105: 2*uc - um1 + dt^2*lap/m) with FMA contraction; it differs from the oracle by
// relative L2 ~1e-6 (tolerance 1e-4, README.md:33).  Neither form flushes denormals.
#pragma once
#include "fdtd_common.cuh"

namespace fdtd {

#define FDTD_C2 (-8.33333333e-2F)  // -1/12  (openacc.cpp:104)
#define FDTD_C1 (1.333333330F)     //  4/3
#define FDTD_C0 (-2.50F)           // -5/2

template <bool EXACT>
__device__ __forceinline__ float axis_term(float r5, float m2, float m1, float p1, float p2)
{
    if (EXACT) {
        // (r5 + c2*(u[-2] + u[+2])) + c1*(u[-1] + u[+1])
        return __fadd_rn(__fadd_rn(r5, __fmul_rn(FDTD_C2, __fadd_rn(m2, p2))),
                         __fmul_rn(FDTD_C1, __fadd_rn(m1, p1)));
    } else {
        return fmaf(FDTD_C1, m1 + p1, fmaf(FDTD_C2, m2 + p2, r5));
    }
}

// c: u[t0] centre; dx,dy,dz: the three axis terms; u1: u[t1] centre; m: squared slowness.
template <bool EXACT>
__device__ __forceinline__ float leapfrog(float c, float dx, float dy, float dz, float u1, float m,
                                          const Coef &k)
{
    if (EXACT) {
        // dt*dt*( r2*dx + r3*dy + r4*dz - ((-2*r1)*u0 + r1*u1)*m ) / m
        const float lap = __fadd_rn(__fadd_rn(__fmul_rn(k.r2, dx), __fmul_rn(k.r3, dy)), __fmul_rn(k.r4, dz));
        const float d = __fmul_rn(__fadd_rn(__fmul_rn(k.n2r1, c), __fmul_rn(k.r1, u1)), m);
        const float num = __fmul_rn(k.dt2, __fsub_rn(lap, d));
        // (+-0)/m == (+-0)*m for finite m != 0: skip the IEEE division where the field is still zero
        // (its FCHK guard sends zero dividends to the slow path); warp-uniform in quiescent regions.
        if (num == 0.0f) return __fmul_rn(num, m);
        return __fdiv_rn(num, m);
    } else {
        const float lap = fmaf(k.r4, dz, fmaf(k.r3, dy, k.r2 * dx));
        return fmaf(2.0f, c, -u1) + __fdividef(k.dt2 * lap, m);
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
