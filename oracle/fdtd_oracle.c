/*
 * fdtd_oracle.c -- CPU restatement of the reference's 3D acoustic FDTD hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA library
 * libfdtd_b200.so or its Python host mirror) may import, link or call this file.
 * It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and as the reported CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 *   (a) bit-for-bit against the unmodified /root/reference/openacc.cpp compiled for the
 *       host into oracle/_ref/libref_openacc.so (oracle/Makefile), when that file exists, and
 *   (b) against the golden fixtures in tests/golden/ that were generated from that same
 *       reference build by tests/golden/make_golden.py.
 *
 * Build (oracle/Makefile): gcc -O3 -ffp-contract=off  -- contraction must stay off so that
 * every fp32 operation rounds exactly like the reference built with g++ -ffp-contract=off.
 * Gradual underflow must stay on (no -ffast-math, no FTZ/DAZ): a third of the benchmark
 * wavefield is denormal.
 *
 * Layout (reference main.cpp:331-367): u is float[3][nxp][nyp][nzp], z contiguous,
 * nxp = nx + 8 (HALO = 4 cells either side); m is float[nxp][nyp][nzp];
 * src is float[T][pstride]; src_coords is float[nsrc][cstride] (x, y, z physical).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <sys/time.h>

#define ORACLE_HALO 4
#define ORACLE_WARMUP_STEPS 5 /* openacc.cpp:5 */

typedef struct {
    int nxp, nyp, nzp;          /* padded extents (u.size[1..3]) */
    int x_m, x_M, y_m, y_M, z_m, z_M;
    float dt, h_x, h_y, h_z, o_x, o_y, o_z;
} oracle_geom;

static inline size_t idx3(const oracle_geom *g, int X, int Y, int Z)
{
    return ((size_t)X * g->nyp + (size_t)Y) * g->nzp + (size_t)Z;
}

/*
 * Section0: 4th-order Laplacian + leapfrog update, OpenACC-form arithmetic.
 * Follows openacc.cpp:84-87 (r1..r4) and openacc.cpp:102-107 (== :159-164) operation by
 * operation.  C's left-to-right evaluation gives:
 *   r5  = -2.5f*u0[c]
 *   dX  = (r5 + c2*(u0[-2] + u0[+2])) + c1*(u0[-1] + u0[+1])            per axis
 *   num = ((r2*dx + r3*dy) + r4*dz) - (((-2.0f*r1)*u0[c]) + r1*u1[c])*m[c]
 *   u2  = ((dt*dt)*num)/m[c]
 * x_lo/x_hi restrict the x range (inclusive, unpadded) so the OpenMP baseline and the
 * slab tests can call it on sub-ranges; the reference always runs [x_m, x_M].
 */
void oracle_section0_range(const oracle_geom *g, const float *m, const float *u0,
                           const float *u1, float *u2, int x_lo, int x_hi)
{
    const float dt = g->dt;
    const float r1 = 1.0F / (dt * dt);
    const float r2 = 1.0F / (g->h_x * g->h_x);
    const float r3 = 1.0F / (g->h_y * g->h_y);
    const float r4 = 1.0F / (g->h_z * g->h_z);
    const size_t sx = (size_t)g->nyp * g->nzp, sy = (size_t)g->nzp;

    for (int x = x_lo; x <= x_hi; x += 1) {
        for (int y = g->y_m; y <= g->y_M; y += 1) {
            for (int z = g->z_m; z <= g->z_M; z += 1) {
                const size_t c = idx3(g, x + 4, y + 4, z + 4);
                float r5 = -2.50F * u0[c];
                u2[c] = dt * dt * (r2 * (r5 + (-8.33333333e-2F) * (u0[c - 2 * sx] + u0[c + 2 * sx]) + 1.333333330F * (u0[c - sx] + u0[c + sx]))
                                   + r3 * (r5 + (-8.33333333e-2F) * (u0[c - 2 * sy] + u0[c + 2 * sy]) + 1.333333330F * (u0[c - sy] + u0[c + sy]))
                                   + r4 * (r5 + (-8.33333333e-2F) * (u0[c - 2] + u0[c + 2]) + 1.333333330F * (u0[c - 1] + u0[c + 1]))
                                   - (-2.0F * r1 * u0[c] + r1 * u1[c]) * m[c]) / m[c];
            }
        }
    }
}

void oracle_section0(const oracle_geom *g, const float *m, const float *u0,
                     const float *u1, float *u2)
{
    oracle_section0_range(g, m, u0, u1, u2, g->x_m, g->x_M);
}

/*
 * Source position: grid cell and trilinear fraction for one axis.
 * Follows openacc.cpp:125-131: g = (-o + c)/h ; pos = (int)floor(g) ; frac = -floor(g) + g.
 */
void oracle_source_pos(float coord, float o, float h, int *pos, float *frac)
{
    const float gq = (-o + coord) / h;
    const float fl = floorf(gq);
    *pos = (int)fl;
    *frac = -fl + gq;
}

/*
 * Section1: trilinear 8-corner scatter of every source into u2, serial p_src order.
 * Follows openacc.cpp:113-143 (== :173-203).  The weight of one axis is
 * r*p + (1-r)*(1-p) with integer r in {0,1} promoted to float; the product is evaluated
 * left to right: ((((1e-2f*wx)*wy)*wz)*src[time][p]) / m[base corner].
 */
void oracle_section1(const oracle_geom *g, const float *m, const float *src, int src_size0,
                     int pstride, const float *coords, int cstride, int p_src_m, int p_src_M,
                     int time, float *u2)
{
    if (!(src_size0 * pstride > 0 && p_src_M - p_src_m + 1 > 0))
        return;
    for (int p_src = p_src_m; p_src <= p_src_M; p_src += 1) {
        for (int rsrcx = 0; rsrcx <= 1; rsrcx += 1) {
            for (int rsrcy = 0; rsrcy <= 1; rsrcy += 1) {
                for (int rsrcz = 0; rsrcz <= 1; rsrcz += 1) {
                    int posx, posy, posz;
                    float px, py, pz;
                    oracle_source_pos(coords[(size_t)p_src * cstride + 0], g->o_x, g->h_x, &posx, &px);
                    oracle_source_pos(coords[(size_t)p_src * cstride + 1], g->o_y, g->h_y, &posy, &py);
                    oracle_source_pos(coords[(size_t)p_src * cstride + 2], g->o_z, g->h_z, &posz, &pz);
                    if (rsrcx + posx >= g->x_m - 1 && rsrcy + posy >= g->y_m - 1 && rsrcz + posz >= g->z_m - 1 &&
                        rsrcx + posx <= g->x_M + 1 && rsrcy + posy <= g->y_M + 1 && rsrcz + posz <= g->z_M + 1) {
                        float r0 = 1.0e-2F * (rsrcx * px + (1 - rsrcx) * (1 - px)) * (rsrcy * py + (1 - rsrcy) * (1 - py)) *
                                   (rsrcz * pz + (1 - rsrcz) * (1 - pz)) * src[(size_t)time * pstride + p_src] /
                                   m[idx3(g, posx + 4, posy + 4, posz + 4)];
                        u2[idx3(g, rsrcx + posx + 4, rsrcy + posy + 4, rsrcz + posz + 4)] += r0;
                    }
                }
            }
        }
    }
}

static double now_s(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + (double)tv.tv_usec / 1e6;
}

/*
 * The operator: time loop with the 3-level ring (openacc.cpp:90-92,148):
 * t0 = time%3 (current), t1 = (time+2)%3 (previous), t2 = (time+1)%3 (next).
 * The first min(5, T) steps are real steps run outside the timers; timers[0]/[1]
 * accumulate (+=) section0/section1 seconds of the remaining steps (openacc.cpp:2-3).
 * nthreads > 1 splits Section0 over x with OpenMP when built with -fopenmp
 * (liboracle_omp.so); the result is bit-identical because points are independent.
 */
int oracle_run(const oracle_geom *g, const float *m, float *u, const float *src, int src_size0,
               int pstride, const float *coords, int cstride, int p_src_m, int p_src_M,
               int time_m, int time_M, double *timers)
{
    const size_t lvl = (size_t)g->nxp * g->nyp * g->nzp;
    for (int time = time_m; time <= time_M; time += 1) {
        const int t0 = time % 3, t1 = (time + 2) % 3, t2 = (time + 1) % 3;
        const int timed = time >= time_m + ORACLE_WARMUP_STEPS;
        double a = now_s();
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
        for (int x = g->x_m; x <= g->x_M; x += 1)
            oracle_section0_range(g, m, u + t0 * lvl, u + t1 * lvl, u + t2 * lvl, x, x);
#else
        oracle_section0(g, m, u + t0 * lvl, u + t1 * lvl, u + t2 * lvl);
#endif
        double b = now_s();
        oracle_section1(g, m, src, src_size0, pstride, coords, cstride, p_src_m, p_src_M, time,
                        u + t2 * lvl);
        double c = now_s();
        if (timed && timers) {
            timers[0] += b - a;
            timers[1] += c - b;
        }
    }
    return 0;
}

/*
 * Input synthesis of the reference's benchmark driver (main.cpp:285-325): Ricker wavelet
 * and source lattice, all in fp32.  `expf` is what std::exp(float) resolves to.
 */
void oracle_fill_ricker(float *src, int T, int S, float dt)
{
    const float f0 = 10.0f;
    for (int t = 0; t < T; ++t) {
        const float tshift = t * dt - 1.0f / f0;
        const float a = (float)M_PI * (float)M_PI * f0 * f0 * tshift * tshift;
        const float val = (1.0f - 2.0f * a) * expf(-a);
        for (int s = 0; s < S; ++s)
            src[(size_t)t * S + s] = val;
    }
}

void oracle_fill_source_coords(float *coords, int S, int nx, int ny, int nz, float h_x, float h_y,
                               float h_z)
{
    const float frac[3] = {0.25f, 0.50f, 0.75f};
    const float h = 0.1f; /* main.cpp:304 -- the lattice uses a literal 0.1f, not h_x */
    const float Lx = (nx - 1) * h, Ly = (ny - 1) * h, Lz = (nz - 1) * h;
    int placed = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) {
                if (placed >= S)
                    break; /* main.cpp:314 breaks the innermost loop only; harmless here */
                coords[3 * placed + 0] = frac[i] * Lx;
                coords[3 * placed + 1] = frac[j] * Ly;
                coords[3 * placed + 2] = frac[k] * Lz;
                ++placed;
            }
    for (; placed < S; ++placed) {
        coords[3 * placed + 0] = 0.5f * (nx - 1) * h_x;
        coords[3 * placed + 1] = 0.5f * (ny - 1) * h_y;
        coords[3 * placed + 2] = 0.5f * (nz - 1) * h_z;
    }
}

/* Dense parity field of the reference's correctness test (main.cpp:525-532), level 2 zeroed. */
void oracle_fill_dense(float *u, float *m, size_t volp)
{
    for (size_t i = 0; i < volp; ++i) {
        m[i] = 1.5f;
        float val = sinf(i * 0.001f) * 10.0f + 100.0f;
        u[i] = u[volp + i] = val;
        u[2 * volp + i] = 0.0f;
    }
}

/* ==========================================================================================
 * Generalisations for the SURVEY 8(f) "next" rows.  The reference ships order-4 kernels only; its
 * driver is parameterised by STENCIL_ORDER with HALO == STENCIL_ORDER cells (main.cpp:2,27-32,129-136)
 * and the operator it was generated from (Devito's acoustic example) also samples receivers.  What
 * follows is THIS repo's definition of those rows, written the way the generator writes order 4:
 * parity of the CUDA path for them is against this file; at order 4 the functions below are
 * bit-identical to the pinned ones above (tests/test_oracle.py).
 * ========================================================================================== */

/* Central second-derivative weights of space order 2R (R = 2..6), c[0] = centre, c[k] = +-k, each the
 * correctly rounded float of the exact rational.  Order 4 reproduces the reference's literals
 * (openacc.cpp:104: -2.50F, 1.333333330F, -8.33333333e-2F bit for bit). */
int oracle_fd_coeffs(int space_order, float *c)
{
    static const float tab[5][7] = {
        {-2.5F, 1.33333337F, -0.0833333358F},
        {-2.72222233F, 1.5F, -0.150000006F, 0.0111111114F},
        {-2.84722233F, 1.60000002F, -0.200000003F, 0.0253968257F, -0.0017857143F},
        {-2.92722225F, 1.66666663F, -0.238095239F, 0.039682541F, -0.00496031763F, 0.000317460304F},
        {-2.98277783F, 1.71428573F, -0.267857134F, 0.0529100522F, -0.00892857183F, 0.001038961F, -6.01250613e-05F}};
    if (space_order < 4 || space_order > 12 || (space_order & 1))
        return -1;
    const int R = space_order / 2;
    for (int k = 0; k <= R; ++k)
        c[k] = tab[R - 2][k];
    return R;
}

/* Section0 at space order 2R, halo = space_order cells.  Per axis, outermost pair first (the order of
 * openacc.cpp:104-106 at R = 2):  d = ((r5 + c_R*(u[-R] + u[+R])) + ...) + c_1*(u[-1] + u[+1]),  r5 = c_0*u[c]. */
void oracle_section0_order(const oracle_geom *g, int space_order, const float *m, const float *u0, const float *u1,
                           float *u2, int x_lo, int x_hi)
{
    float cf[7];
    const int R = oracle_fd_coeffs(space_order, cf), H = space_order;
    const float dt = g->dt;
    const float r1 = 1.0F / (dt * dt);
    const float r2 = 1.0F / (g->h_x * g->h_x);
    const float r3 = 1.0F / (g->h_y * g->h_y);
    const float r4 = 1.0F / (g->h_z * g->h_z);
    const size_t sx = (size_t)g->nyp * g->nzp, sy = (size_t)g->nzp;
    if (R < 0)
        return;
    for (int x = x_lo; x <= x_hi; x += 1)
        for (int y = g->y_m; y <= g->y_M; y += 1)
            for (int z = g->z_m; z <= g->z_M; z += 1) {
                const size_t c = idx3(g, x + H, y + H, z + H);
                const float r5 = cf[0] * u0[c];
                float dx = r5, dy = r5, dz = r5;
                for (int k = R; k >= 1; --k) {
                    dx = dx + cf[k] * (u0[c - k * sx] + u0[c + k * sx]);
                    dy = dy + cf[k] * (u0[c - k * sy] + u0[c + k * sy]);
                    dz = dz + cf[k] * (u0[c - k] + u0[c + k]);
                }
                u2[c] = dt * dt * (r2 * dx + r3 * dy + r4 * dz - (-2.0F * r1 * u0[c] + r1 * u1[c]) * m[c]) / m[c];
            }
}

/* Section1 with a halo of H cells (openacc.cpp:113-143 with "+ 4" replaced by "+ H"). */
void oracle_section1_halo(const oracle_geom *g, int H, const float *m, const float *src, int src_size0, int pstride,
                          const float *coords, int cstride, int p_src_m, int p_src_M, int time, float *u2)
{
    if (!(src_size0 * pstride > 0 && p_src_M - p_src_m + 1 > 0))
        return;
    for (int p_src = p_src_m; p_src <= p_src_M; p_src += 1)
        for (int rsrcx = 0; rsrcx <= 1; rsrcx += 1)
            for (int rsrcy = 0; rsrcy <= 1; rsrcy += 1)
                for (int rsrcz = 0; rsrcz <= 1; rsrcz += 1) {
                    int posx, posy, posz;
                    float px, py, pz;
                    oracle_source_pos(coords[(size_t)p_src * cstride + 0], g->o_x, g->h_x, &posx, &px);
                    oracle_source_pos(coords[(size_t)p_src * cstride + 1], g->o_y, g->h_y, &posy, &py);
                    oracle_source_pos(coords[(size_t)p_src * cstride + 2], g->o_z, g->h_z, &posz, &pz);
                    if (rsrcx + posx >= g->x_m - 1 && rsrcy + posy >= g->y_m - 1 && rsrcz + posz >= g->z_m - 1 &&
                        rsrcx + posx <= g->x_M + 1 && rsrcy + posy <= g->y_M + 1 && rsrcz + posz <= g->z_M + 1) {
                        float r0 = 1.0e-2F * (rsrcx * px + (1 - rsrcx) * (1 - px)) * (rsrcy * py + (1 - rsrcy) * (1 - py)) *
                                   (rsrcz * pz + (1 - rsrcz) * (1 - pz)) * src[(size_t)time * pstride + p_src] /
                                   m[idx3(g, posx + H, posy + H, posz + H)];
                        u2[idx3(g, rsrcx + posx + H, rsrcy + posy + H, rsrcz + posz + H)] += r0;
                    }
                }
}

/* Receiver sampling ("Section2" of the generated operator this path comes from: rec.interpolate(expr=u)):
 * rec[time][p] = sum over the 8 trilinear corners, x outermost / z innermost, of ((wx*wy)*wz)*u[t0][corner],
 * same positions, fractions and bounds test as the injection; a corner outside [m-1, M+1] contributes nothing. */
void oracle_section2_halo(const oracle_geom *g, int H, const float *u0, const float *coords, int nrec, int cstride,
                          float *rec_row)
{
    for (int p = 0; p < nrec; p += 1) {
        int posx, posy, posz;
        float px, py, pz;
        oracle_source_pos(coords[(size_t)p * cstride + 0], g->o_x, g->h_x, &posx, &px);
        oracle_source_pos(coords[(size_t)p * cstride + 1], g->o_y, g->h_y, &posy, &py);
        oracle_source_pos(coords[(size_t)p * cstride + 2], g->o_z, g->h_z, &posz, &pz);
        float sum = 0.0F;
        for (int rx = 0; rx <= 1; rx += 1)
            for (int ry = 0; ry <= 1; ry += 1)
                for (int rz = 0; rz <= 1; rz += 1)
                    if (rx + posx >= g->x_m - 1 && ry + posy >= g->y_m - 1 && rz + posz >= g->z_m - 1 &&
                        rx + posx <= g->x_M + 1 && ry + posy <= g->y_M + 1 && rz + posz <= g->z_M + 1)
                        sum += (rx * px + (1 - rx) * (1 - px)) * (ry * py + (1 - ry) * (1 - py)) * (rz * pz + (1 - rz) * (1 - pz)) *
                               u0[idx3(g, rx + posx + H, ry + posy + H, rz + posz + H)];
        rec_row[p] = sum;
    }
}

/* The operator at space order 2R with optional receivers: per step Section0, Section1, then Section2 samples the
 * CURRENT level u[t0] into rec[time - time_m][0..nrec).  Same ring and timers as oracle_run. */
int oracle_run_order(const oracle_geom *g, int space_order, const float *m, float *u, const float *src, int src_size0,
                     int pstride, const float *coords, int cstride, int p_src_m, int p_src_M, int time_m, int time_M,
                     const float *rec_coords, int nrec, int rec_cstride, float *rec, double *timers)
{
    float cf[7];
    if (oracle_fd_coeffs(space_order, cf) < 0)
        return -1;
    const size_t lvl = (size_t)g->nxp * g->nyp * g->nzp;
    for (int time = time_m; time <= time_M; time += 1) {
        const int t0 = time % 3, t1 = (time + 2) % 3, t2 = (time + 1) % 3;
        const int timed = time >= time_m + ORACLE_WARMUP_STEPS;
        double a = now_s();
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int x = g->x_m; x <= g->x_M; x += 1)
            oracle_section0_order(g, space_order, m, u + t0 * lvl, u + t1 * lvl, u + t2 * lvl, x, x);
        double b = now_s();
        oracle_section1_halo(g, space_order, m, src, src_size0, pstride, coords, cstride, p_src_m, p_src_M, time,
                             u + t2 * lvl);
        if (rec && nrec > 0)
            oracle_section2_halo(g, space_order, u + t0 * lvl, rec_coords, nrec, rec_cstride,
                                 rec + (size_t)(time - time_m) * nrec);
        double c = now_s();
        if (timed && timers) {
            timers[0] += b - a;
            timers[1] += c - b;
        }
    }
    return 0;
}
