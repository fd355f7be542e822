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

// dt*dt*( r2*dx + r3*dy + r4*dz - ((-2*r1)*u0 + r1*u1)*m )  -- the dividend of openacc.cpp:103-107, reference order
__device__ __forceinline__ float numerator_exact(float c, float dx, float dy, float dz, float u1, float m, const Coef &k)
{
    const float lap = __fadd_rn(__fadd_rn(__fmul_rn(k.r2, dx), __fmul_rn(k.r3, dy)), __fmul_rn(k.r4, dz));
    const float d = __fmul_rn(__fadd_rn(__fmul_rn(k.n2r1, c), __fmul_rn(k.r1, u1)), m);
    return __fmul_rn(k.dt2, __fsub_rn(lap, d));
}

// num / m, correctly rounded (IEEE), without the fp32 division's slow paths.
__device__ __forceinline__ float divide_exact(float num, float m)
{
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

// The same for the four points of a float4 with ONE classification instead of eight tests: all four dividends zero
// (the quiescent part of the benchmark field), all four well inside the normal range (a dense field), or mixed
// (the fringe of the wavefield: point by point).  Same operations per point, so the bits do not change.
__device__ __forceinline__ float4 divide4_exact(const float4 &n, const float4 &m)
{
    const unsigned any = (__float_as_uint(n.x) | __float_as_uint(n.y) | __float_as_uint(n.z) | __float_as_uint(n.w)) & 0x7fffffffu;
    if (any == 0u) return make_float4(__fmul_rn(n.x, m.x), __fmul_rn(n.y, m.y), __fmul_rn(n.z, m.z), __fmul_rn(n.w, m.w));
    const float amin = fminf(fminf(fabsf(n.x), fabsf(n.y)), fminf(fabsf(n.z), fabsf(n.w)));
    if (amin >= 0x1p-80f) return make_float4(__fdiv_rn(n.x, m.x), __fdiv_rn(n.y, m.y), __fdiv_rn(n.z, m.z), __fdiv_rn(n.w, m.w));
    return make_float4(divide_exact(n.x, m.x), divide_exact(n.y, m.y), divide_exact(n.z, m.z), divide_exact(n.w, m.w));
}

// dt*dt*( r2*dx + r3*dy + r4*dz - ((-2*r1)*u0 + r1*u1)*m ) / m  -- openacc.cpp:103-107, reference order
__device__ __forceinline__ float leapfrog_exact(float c, float dx, float dy, float dz, float u1, float m, const Coef &k)
{
    return divide_exact(numerator_exact(c, dx, dy, dz, u1, m, k), m);
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

// The dividend of one point in exact arithmetic (point<true>() without its final division).
__device__ __forceinline__ float point_numerator_exact(float c, float xm2, float xm1, float xp1, float xp2, float ym2, float ym1,
                                                       float yp1, float yp2, float zm2, float zm1, float zp1, float zp2, float u1,
                                                       float m, const Coef &k)
{
    const float r5 = __fmul_rn(FDTD_C0, c);
    const float dx = axis_term_exact(r5, xm2, xm1, xp1, xp2);
    const float dy = axis_term_exact(r5, ym2, ym1, yp1, yp2);
    const float dz = axis_term_exact(r5, zm2, zm1, zp1, zp2);
    return numerator_exact(c, dx, dy, dz, u1, m, k);
}

// The four points of one float4 column: c = centre, xm2..xp2 = the same column on the neighbouring planes, ym2..yp2 =
// the rows above / below, zl / zr = the two floats left / right of the column, u1 = previous time level.
template <bool EXACT>
__device__ __forceinline__ float4 column4(const float4 &c, const float4 &xm2, const float4 &xm1, const float4 &xp1,
                                          const float4 &xp2, const float4 &ym2, const float4 &ym1, const float4 &yp1,
                                          const float4 &yp2, const float2 &zl, const float2 &zr, const float4 &u1,
                                          const float4 &m, const Coef &k)
{
    float4 o;
    if (EXACT) {
        o.x = point_numerator_exact(c.x, xm2.x, xm1.x, xp1.x, xp2.x, ym2.x, ym1.x, yp1.x, yp2.x, zl.x, zl.y, c.y, c.z, u1.x, m.x, k);
        o.y = point_numerator_exact(c.y, xm2.y, xm1.y, xp1.y, xp2.y, ym2.y, ym1.y, yp1.y, yp2.y, zl.y, c.x, c.z, c.w, u1.y, m.y, k);
        o.z = point_numerator_exact(c.z, xm2.z, xm1.z, xp1.z, xp2.z, ym2.z, ym1.z, yp1.z, yp2.z, c.x, c.y, c.w, zr.x, u1.z, m.z, k);
        o.w = point_numerator_exact(c.w, xm2.w, xm1.w, xp1.w, xp2.w, ym2.w, ym1.w, yp1.w, yp2.w, c.y, c.z, zr.x, zr.y, u1.w, m.w, k);
        return divide4_exact(o, m);
    }
    o.x = point<false>(c.x, xm2.x, xm1.x, xp1.x, xp2.x, ym2.x, ym1.x, yp1.x, yp2.x, zl.x, zl.y, c.y, c.z, u1.x, m.x, k);
    o.y = point<false>(c.y, xm2.y, xm1.y, xp1.y, xp2.y, ym2.y, ym1.y, yp1.y, yp2.y, zl.y, c.x, c.z, c.w, u1.y, m.y, k);
    o.z = point<false>(c.z, xm2.z, xm1.z, xp1.z, xp2.z, ym2.z, ym1.z, yp1.z, yp2.z, c.x, c.y, c.w, zr.x, u1.z, m.z, k);
    o.w = point<false>(c.w, xm2.w, xm1.w, xp1.w, xp2.w, ym2.w, ym1.w, yp1.w, yp2.w, c.y, c.z, zr.x, zr.y, u1.w, m.w, k);
    return o;
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
