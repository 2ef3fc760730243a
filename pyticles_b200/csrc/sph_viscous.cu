// pyticles_b200 -- velocity gradient and Newtonian viscous pair force for sm_100a.
//
// BUILDER-DEFINED ARITHMETIC.  The reference routes the irreversible stress through the Fortran routine
// sphforce3d.calc_sphforce3d (spam_complete_force.py:158-165), whose source is not in the reference tree,
// so there is nothing to be bit-compatible with.  What is kept from the reference:
//   velocity gradient  gradv_i[a][b] = sum_j w_j (v_j - v_i)_a dW_ij/dx_b          properties.py:95-98,
//                      with the volume weight w_j = m_j / rho_j of c_properties.pyx:166-188 and the FINAL
//                      summation density (the reference uses the running one, which makes its result depend
//                      on the pair order)
//                      NOTE the sign: dW_ij/dx_b is the kernel gradient with respect to r_j - r_i, as everywhere
//                      in the reference, so this sum estimates MINUS the velocity gradient (v = A r gives
//                      gradv ~ -A); p.gradv keeps the reference's convention.
//   stress             pi = -2 eta symmetric_traceless(grad v) - zeta (div v) I with grad v = -gradv, i.e.
//                      pi = 2 eta symmetric_traceless(gradv) + zeta tr(gradv) I      tensor.py:5-16, the eta / zeta
//                      arguments of SpamComplete (spam_complete_force.py:36-60); dissipative: shear heats
//   pair force         a = (P_i / rho_i^2 + P_j / rho_j^2) . dW_ij, +a to i, -a to j, no mass factor;
//                      du = a . dv / 2, udot_i += du m_j, udot_j += du m_i          forces.py:353-368 with the
//                      scalar pressure replaced by the stress tensor
// Checked against oracle/oracle.py (gradv_two_pass, viscous_force) and by invariants (tests/test_gpu_viscous.py).
//
// Same structure as the pressure force pass: one thread per Morton-sorted particle over its warp-transposed
// ELL row, neighbours gathered as 32-byte rows one stage ahead, fp64 accumulation in registers, fixed order.
#include "sph_device.cuh"

namespace {

constexpr int kVU = 2;               // neighbours gathered per pipeline stage

// aux4[a] = (vx, vy, vz, m / rho) of sorted particle a
__global__ void __launch_bounds__(kBlock)
volume_term_kernel(int n, const int32_t *__restrict__ perm, const double *__restrict__ pos4,
                   const double *__restrict__ vel4, const double *__restrict__ rho, double *__restrict__ aux4)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    double x, y, z, m, vx, vy, vz, w;
    load4(pos4 + 4 * (size_t)a, x, y, z, m);
    load4(vel4 + 4 * (size_t)a, vx, vy, vz, w);
    store4(aux4 + 4 * (size_t)a, vx, vy, vz, m / rho[perm[a]]);
}

struct Geom {
    double dx, dy, dz, rr, fac;      // r_other - r_self, its length, and dW/dx_a = fac * d_a
    bool in;
};

template <bool UNIFORM_H, bool WRAP>
__device__ __forceinline__ Geom pair_geom(const sph_grid &g, double px, double py, double pz, double bx, double by,
                                          double bz, const int32_t *__restrict__ perm,
                                          const double *__restrict__ h_orig, int orig, int j, double hinv, double c2)
{
    Geom q;
    q.dx = bx - px; q.dy = by - py; q.dz = bz - pz;
    if (WRAP) {
        q.dx = min_image(q.dx, g.box[0], g.box[0] / 2.);
        q.dy = min_image(q.dy, g.box[1], g.box[1] / 2.);
        q.dz = min_image(q.dz, g.box[2], g.box[2] / 2.);
    }
    q.rr = sqrt(rsq_exact(q.dx, q.dy, q.dz));
    double hi = hinv, cc = c2;
    if (!UNIFORM_H) {
        const int oj = perm[j];
        const double h = h_orig[oj < orig ? oj : orig];             // properties.py:88: h of the first member
        hi = 1.0 / h;
        cc = -12.0 * lucy_norm3(h) * hi * hi;
    }
    const double s = q.rr * hi;
    const double t = 1.0 - s;
    q.in = s < 1.0;                                                  // spkernel.py:106
    q.fac = cc * (t * t);                                            // spkernel.py:113-114 divided by r
    return q;
}

template <bool UNIFORM_H, bool WRAP>
__device__ __forceinline__ void gradv_row(const sph_grid &g, const double *__restrict__ pos4,
                                          const double *__restrict__ aux4, const int32_t *__restrict__ perm,
                                          const double *__restrict__ h_orig, const int32_t *__restrict__ row,
                                          int count, int orig, int self, double px, double py, double pz, double vx,
                                          double vy, double vz, double hinv, double c2, double G[9])
{
    int jn[kVU];
#pragma unroll
    for (int u = 0; u < kVU; ++u) jn[u] = u < count ? row[(size_t)u * 32] : self;
    for (int k0 = 0; k0 < count; k0 += kVU) {
        double bx[kVU], by[kVU], bz[kVU], bm[kVU], wx[kVU], wy[kVU], wz[kVU], ww[kVU];
        int j[kVU];
#pragma unroll
        for (int u = 0; u < kVU; ++u) {
            j[u] = jn[u];
            load4(pos4 + 4 * (size_t)j[u], bx[u], by[u], bz[u], bm[u]);
            load4(aux4 + 4 * (size_t)j[u], wx[u], wy[u], wz[u], ww[u]);
        }
#pragma unroll
        for (int u = 0; u < kVU; ++u) {
            const int kn = k0 + kVU + u;
            jn[u] = kn < count ? row[(size_t)kn * 32] : self;
        }
#pragma unroll
        for (int u = 0; u < kVU; ++u) {
            const Geom q = pair_geom<UNIFORM_H, WRAP>(g, px, py, pz, bx[u], by[u], bz[u], perm, h_orig, orig, j[u],
                                                      hinv, c2);
            if (q.in && k0 + u < count) {
                const double dw[3] = {q.fac * q.dx, q.fac * q.dy, q.fac * q.dz};
                const double dv[3] = {ww[u] * (wx[u] - vx), ww[u] * (wy[u] - vy), ww[u] * (wz[u] - vz)};
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) G[3 * a + b] += dv[a] * dw[b];
            }
        }
    }
}

template <bool UNIFORM_H>
__global__ void __launch_bounds__(kBlock)
gradv_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
             const double *__restrict__ aux4, const float *__restrict__ rel4, const int32_t *__restrict__ perm,
             const int32_t *__restrict__ nbr, const int32_t *__restrict__ cnt,
             const sph_status *__restrict__ status, const double *__restrict__ h_orig, int list_fresh,
             double *__restrict__ gradv)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = a < n;
    double px = 0, py = 0, pz = 0, pm = 0, vx = 0, vy = 0, vz = 0, vw = 0;
    int count = 0, orig = 0;
    bool interior = true;
    if (active) {
        load4(pos4 + 4 * (size_t)a, px, py, pz, pm);
        load4(aux4 + 4 * (size_t)a, vx, vy, vz, vw);
        count = min(cnt[a], K);
        orig = perm[a];
        interior = (__float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w) & 1u) != 0u;
    }
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
    const double h0 = h_orig[0];
    const double hinv = 1.0 / h0, c2 = -12.0 * lucy_norm3(h0) * hinv * hinv;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (skip) gradv_row<UNIFORM_H, false>(g, pos4, aux4, perm, h_orig, row, count, orig, a, px, py, pz, vx, vy, vz, hinv, c2, G);
    else gradv_row<UNIFORM_H, true>(g, pos4, aux4, perm, h_orig, row, count, orig, a, px, py, pz, vx, vy, vz, hinv, c2, G);
    (void)pm; (void)vw;
    if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) gradv[9 * (size_t)orig + k] = G[k];
    }
}

// aux8[a] = pi / rho^2 of sorted particle a as (xx, yy, zz, xy, xz, yz, 0, 0);
// pi = 2 eta symmetric_traceless(gradv) + zeta tr(gradv) I   (tensor.py:5-16; gradv = -grad v, see the top)
__global__ void __launch_bounds__(kBlock)
stress_term_kernel(int n, const int32_t *__restrict__ perm, const double *__restrict__ gradv,
                   const double *__restrict__ rho, double eta, double zeta, double *__restrict__ aux8)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const size_t o = (size_t)perm[a];
    const double *G = gradv + 9 * o;
    const double tr = (G[0] + G[4]) + G[8];
    const double d = rho[o], inv = 1.0 / (d * d);
    const double iso = tr / 3.0;
    const double sxx = 2.0 * eta * (G[0] - iso) + zeta * tr;
    const double syy = 2.0 * eta * (G[4] - iso) + zeta * tr;
    const double szz = 2.0 * eta * (G[8] - iso) + zeta * tr;
    const double sxy = 2.0 * eta * (0.5 * (G[1] + G[3]));
    const double sxz = 2.0 * eta * (0.5 * (G[2] + G[6]));
    const double syz = 2.0 * eta * (0.5 * (G[5] + G[7]));
    store4(aux8 + 8 * (size_t)a, sxx * inv, syy * inv, szz * inv, sxy * inv);
    store4(aux8 + 8 * (size_t)a + 4, sxz * inv, syz * inv, 0.0, 0.0);
}

struct VAcc { double ax, ay, az, du; };

template <bool UNIFORM_H, bool WRAP>
__device__ __forceinline__ VAcc viscous_row(const sph_grid &g, const double *__restrict__ pos4,
                                            const double *__restrict__ vel4, const double *__restrict__ aux8,
                                            const int32_t *__restrict__ perm, const double *__restrict__ h_orig,
                                            const int32_t *__restrict__ row, int count, int orig, int self,
                                            double px, double py, double pz, double vx, double vy, double vz,
                                            const double S[6], double hinv, double c2, double fcutsq)
{
    VAcc f = {0.0, 0.0, 0.0, 0.0};
    int jn = count > 0 ? row[0] : self;
    for (int k = 0; k < count; ++k) {
        const int j = jn;
        double bx, by, bz, bm, wx, wy, wz, ww, t0, t1, t2, t3, t4, t5, t6, t7;
        load4(pos4 + 4 * (size_t)j, bx, by, bz, bm);
        load4(vel4 + 4 * (size_t)j, wx, wy, wz, ww);
        load4(aux8 + 8 * (size_t)j, t0, t1, t2, t3);
        load4(aux8 + 8 * (size_t)j + 4, t4, t5, t6, t7);
        jn = k + 1 < count ? row[(size_t)(k + 1) * 32] : self;
        const Geom q = pair_geom<UNIFORM_H, WRAP>(g, px, py, pz, bx, by, bz, perm, h_orig, orig, j, hinv, c2);
        if (q.in && q.rr * q.rr <= fcutsq) {                         // forces.py:40
            const double dwx = q.fac * q.dx, dwy = q.fac * q.dy, dwz = q.fac * q.dz;
            const double xx = S[0] + t0, yy = S[1] + t1, zz = S[2] + t2, xy = S[3] + t3, xz = S[4] + t4, yz = S[5] + t5;
            const double gx = (xx * dwx + xy * dwy) + xz * dwz;
            const double gy = (xy * dwx + yy * dwy) + yz * dwz;
            const double gz = (xz * dwx + yz * dwy) + zz * dwz;
            f.ax += gx;
            f.ay += gy;
            f.az += gz;
            const double dot = (gx * (wx - vx) + gy * (wy - vy)) + gz * (wz - vz);
            f.du += (0.5 * dot) * bm;
        }
        (void)ww; (void)t6; (void)t7;
    }
    return f;
}

template <bool UNIFORM_H>
__global__ void __launch_bounds__(kBlock)
viscous_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
               const double *__restrict__ vel4, const double *__restrict__ aux8, const float *__restrict__ rel4,
               const int32_t *__restrict__ perm, const int32_t *__restrict__ nbr, const int32_t *__restrict__ cnt,
               const sph_status *__restrict__ status, const double *__restrict__ h_orig, int list_fresh,
               double fcutsq, double *__restrict__ vdot, double *__restrict__ udot)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = a < n;
    double px = 0, py = 0, pz = 0, pm = 0, vx = 0, vy = 0, vz = 0, vw = 0;
    double S[6] = {0, 0, 0, 0, 0, 0}, s6 = 0, s7 = 0;
    int count = 0, orig = 0;
    bool interior = true;
    if (active) {
        load4(pos4 + 4 * (size_t)a, px, py, pz, pm);
        load4(vel4 + 4 * (size_t)a, vx, vy, vz, vw);
        load4(aux8 + 8 * (size_t)a, S[0], S[1], S[2], S[3]);
        load4(aux8 + 8 * (size_t)a + 4, S[4], S[5], s6, s7);
        count = min(cnt[a], K);
        orig = perm[a];
        interior = (__float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w) & 1u) != 0u;
    }
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
    const double h0 = h_orig[0];
    const double hinv = 1.0 / h0, c2 = -12.0 * lucy_norm3(h0) * hinv * hinv;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    VAcc f;
    if (skip) f = viscous_row<UNIFORM_H, false>(g, pos4, vel4, aux8, perm, h_orig, row, count, orig, a, px, py, pz, vx, vy, vz, S, hinv, c2, fcutsq);
    else f = viscous_row<UNIFORM_H, true>(g, pos4, vel4, aux8, perm, h_orig, row, count, orig, a, px, py, pz, vx, vy, vz, S, hinv, c2, fcutsq);
    (void)pm; (void)vw; (void)s6; (void)s7;
    if (active) {
        vdot[3 * (size_t)orig] += f.ax;
        vdot[3 * (size_t)orig + 1] += f.ay;
        vdot[3 * (size_t)orig + 2] += f.az;
        udot[orig] += f.du;
    }
}

// ------------------------------------------------------------------ SPH gradient of a per-particle scalar
// BUILDER-DEFINED (the reference hands grad_rho_lr and the heat flux jq to the absent Fortran routine,
// spam_complete_force.py:113-115,158-165):
//     out_i = sum_j wgt_j (f_j - c f_i) grad_i W_ij,   grad_i W_ij = -dW_ij/d(r_j - r_i)
// With f = 1, c = 0, wgt = m this is the density gradient sum_j m_j grad_i W_ij; with f = T, c = 1,
// wgt = m / rho it is the usual difference form of grad T.  aux4[a] = (f, wgt, 0, 0) of sorted particle a.
__global__ void __launch_bounds__(kBlock)
scalar_term_kernel(int n, const int32_t *__restrict__ perm, const double *__restrict__ f,
                   const double *__restrict__ wgt, double *__restrict__ aux4)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const size_t o = (size_t)perm[a];
    store4(aux4 + 4 * (size_t)a, f ? f[o] : 1.0, wgt[o], 0.0, 0.0);
}

template <bool UNIFORM_H, bool WRAP>
__device__ __forceinline__ void gradient_row(const sph_grid &g, const double *__restrict__ pos4,
                                             const double *__restrict__ aux4, const int32_t *__restrict__ perm,
                                             const double *__restrict__ h_orig, const int32_t *__restrict__ row,
                                             int count, int orig, int self, double px, double py, double pz,
                                             double fself, double hinv, double c2, double G[3])
{
    int jn = count > 0 ? row[0] : self;
    for (int k = 0; k < count; ++k) {
        const int j = jn;
        double bx, by, bz, bm, fj, wj, e2, e3;
        load4(pos4 + 4 * (size_t)j, bx, by, bz, bm);
        load4(aux4 + 4 * (size_t)j, fj, wj, e2, e3);
        jn = k + 1 < count ? row[(size_t)(k + 1) * 32] : self;
        const Geom q = pair_geom<UNIFORM_H, WRAP>(g, px, py, pz, bx, by, bz, perm, h_orig, orig, j, hinv, c2);
        if (q.in) {
            const double w = wj * (fj - fself);
            G[0] -= w * (q.fac * q.dx);
            G[1] -= w * (q.fac * q.dy);
            G[2] -= w * (q.fac * q.dz);
        }
        (void)bm; (void)e2; (void)e3;
    }
}

template <bool UNIFORM_H>
__global__ void __launch_bounds__(kBlock)
gradient_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
                const double *__restrict__ aux4, const float *__restrict__ rel4, const int32_t *__restrict__ perm,
                const int32_t *__restrict__ nbr, const int32_t *__restrict__ cnt,
                const sph_status *__restrict__ status, const double *__restrict__ h_orig, int list_fresh,
                int subtract_self, double *__restrict__ out)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = a < n;
    double px = 0, py = 0, pz = 0, pm = 0, fs = 0, ws = 0, e2 = 0, e3 = 0;
    int count = 0, orig = 0;
    bool interior = true;
    if (active) {
        load4(pos4 + 4 * (size_t)a, px, py, pz, pm);
        load4(aux4 + 4 * (size_t)a, fs, ws, e2, e3);
        count = min(cnt[a], K);
        orig = perm[a];
        interior = (__float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w) & 1u) != 0u;
    }
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
    const double h0 = h_orig[0];
    const double hinv = 1.0 / h0, c2 = -12.0 * lucy_norm3(h0) * hinv * hinv;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    double G[3] = {0, 0, 0};
    const double fself = subtract_self ? fs : 0.0;
    if (skip) gradient_row<UNIFORM_H, false>(g, pos4, aux4, perm, h_orig, row, count, orig, a, px, py, pz, fself, hinv, c2, G);
    else gradient_row<UNIFORM_H, true>(g, pos4, aux4, perm, h_orig, row, count, orig, a, px, py, pz, fself, hinv, c2, G);
    (void)pm; (void)ws; (void)e2; (void)e3;
    if (active) {
        out[3 * (size_t)orig] = G[0];
        out[3 * (size_t)orig + 1] = G[1];
        out[3 * (size_t)orig + 2] = G[2];
    }
}

// aux8[a] = S / rho^2 of sorted particle a as (xx, yy, zz, xy, xz, yz, 0, 0) for a symmetric stress S[n,3,3]
__global__ void __launch_bounds__(kBlock)
stress_pack_kernel(int n, const int32_t *__restrict__ perm, const double *__restrict__ stress,
                   const double *__restrict__ rho, double *__restrict__ aux8)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const size_t o = (size_t)perm[a];
    const double *S = stress + 9 * o;
    const double d = rho[o], inv = 1.0 / (d * d);
    store4(aux8 + 8 * (size_t)a, S[0] * inv, S[4] * inv, S[8] * inv, (0.5 * (S[1] + S[3])) * inv);
    store4(aux8 + 8 * (size_t)a + 4, (0.5 * (S[2] + S[6])) * inv, (0.5 * (S[5] + S[7])) * inv, 0.0, 0.0);
}

// ------------------------------------------------------------------ repulsive core
// BUILDER-DEFINED (SpamComplete's `sigma` = core size, `rcoef` = core strength, spam_complete_force.py:28-29,52-53,
// arithmetic in the absent Fortran).  Pair potential per unit mass phi(r) = rcoef (1 - r^2/sigma^2)^4 for r < sigma:
//     a = -(8 rcoef / sigma^2) (1 - r^2/sigma^2)^3 (r_j - r_i),  +a to i, -a to j (no mass factor, like the pressure
//     force), du = a . dv / 2, udot_i += du m_j, udot_j += du m_i   (the conventions of forces.py:353-368)
template <bool WRAP>
__device__ __forceinline__ VAcc core_row(const sph_grid &g, const double *__restrict__ pos4,
                                         const double *__restrict__ vel4, const int32_t *__restrict__ row, int count,
                                         int self, double px, double py, double pz, double vx, double vy, double vz,
                                         double inv_s2, double coef)
{
    VAcc f = {0.0, 0.0, 0.0, 0.0};
    int jn = count > 0 ? row[0] : self;
    for (int k = 0; k < count; ++k) {
        const int j = jn;
        double bx, by, bz, bm, wx, wy, wz, ww;
        load4(pos4 + 4 * (size_t)j, bx, by, bz, bm);
        load4(vel4 + 4 * (size_t)j, wx, wy, wz, ww);
        jn = k + 1 < count ? row[(size_t)(k + 1) * 32] : self;
        double dx = bx - px, dy = by - py, dz = bz - pz;
        if (WRAP) {
            dx = min_image(dx, g.box[0], g.box[0] / 2.);
            dy = min_image(dy, g.box[1], g.box[1] / 2.);
            dz = min_image(dz, g.box[2], g.box[2] / 2.);
        }
        const double s = rsq_exact(dx, dy, dz) * inv_s2;
        if (s < 1.0) {
            const double t = 1.0 - s;
            const double fac = coef * (t * t * t);
            const double gx = fac * dx, gy = fac * dy, gz = fac * dz;
            f.ax += gx;
            f.ay += gy;
            f.az += gz;
            const double dot = (gx * (wx - vx) + gy * (wy - vy)) + gz * (wz - vz);
            f.du += (0.5 * dot) * bm;
        }
        (void)ww;
    }
    return f;
}

__global__ void __launch_bounds__(kBlock)
core_kernel(const __grid_constant__ sph_grid g, int n, int K, const double *__restrict__ pos4,
            const double *__restrict__ vel4, const float *__restrict__ rel4, const int32_t *__restrict__ perm,
            const int32_t *__restrict__ nbr, const int32_t *__restrict__ cnt, const sph_status *__restrict__ status,
            int list_fresh, double inv_s2, double coef, double *__restrict__ vdot, double *__restrict__ udot)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = a < n;
    double px = 0, py = 0, pz = 0, pm = 0, vx = 0, vy = 0, vz = 0, vw = 0;
    int count = 0, orig = 0;
    bool interior = true;
    if (active) {
        load4(pos4 + 4 * (size_t)a, px, py, pz, pm);
        load4(vel4 + 4 * (size_t)a, vx, vy, vz, vw);
        count = min(cnt[a], K);
        orig = perm[a];
        interior = (__float_as_uint(reinterpret_cast<const float4 *>(rel4)[a].w) & 1u) != 0u;
    }
    const bool can_skip = list_fresh && !(status->flags & (SPH_F_OUT_OF_BOX | SPH_F_OUT_OF_RANGE | SPH_F_NONFINITE));
    const bool skip = __all_sync(0xffffffffu, interior) && can_skip;
    const int32_t *row = nbr + ((size_t)(a >> 5) * (size_t)K) * 32 + (a & 31);
    VAcc f;
    if (skip) f = core_row<false>(g, pos4, vel4, row, count, a, px, py, pz, vx, vy, vz, inv_s2, coef);
    else f = core_row<true>(g, pos4, vel4, row, count, a, px, py, pz, vx, vy, vz, inv_s2, coef);
    (void)pm; (void)vw;
    if (active) {
        vdot[3 * (size_t)orig] += f.ax;
        vdot[3 * (size_t)orig + 1] += f.ay;
        vdot[3 * (size_t)orig + 2] += f.az;
        udot[orig] += f.du;
    }
}

inline int blocks_of(int64_t n) { return (int)((n + kBlock - 1) / kBlock); }

inline int status_of_launch()
{
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPH_OK : (int)e;
}

}  // namespace

extern "C" {

int sph_gradv(const sph_grid *g, const sph_buffers *b, const double *d_rho, const double *d_h_orig, int h_uniform,
              int list_fresh, double *d_aux4, double *d_gradv, void *stream)
{
    if (!g || !b || !d_rho || !d_h_orig || !d_aux4 || !d_gradv) return SPH_E_BADARG;
    if (!b->pos4 || !b->vel4 || !b->rel4 || !b->perm || !b->nbr || !b->cnt || !b->status) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_of(b->n);
    volume_term_kernel<<<nb, kBlock, 0, s>>>(b->n, b->perm, b->pos4, b->vel4, d_rho, d_aux4);
    if (h_uniform)
        gradv_kernel<true><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, d_aux4, b->rel4, b->perm, b->nbr,
                                                 b->cnt, b->status, d_h_orig, list_fresh, d_gradv);
    else
        gradv_kernel<false><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, d_aux4, b->rel4, b->perm, b->nbr,
                                                  b->cnt, b->status, d_h_orig, list_fresh, d_gradv);
    return status_of_launch();
}

int sph_viscous_force(const sph_grid *g, const sph_buffers *b, const double *d_gradv, const double *d_rho, double eta,
                      double zeta, const double *d_h_orig, int h_uniform, int list_fresh, double fcutoff,
                      double *d_aux8, double *d_vdot, double *d_udot, void *stream)
{
    if (!g || !b || !d_gradv || !d_rho || !d_h_orig || !d_aux8 || !d_vdot || !d_udot) return SPH_E_BADARG;
    if (!b->pos4 || !b->vel4 || !b->rel4 || !b->perm || !b->nbr || !b->cnt || !b->status) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_of(b->n);
    const double fcutsq = fcutoff * fcutoff;                         // forces.py:36
    stress_term_kernel<<<nb, kBlock, 0, s>>>(b->n, b->perm, d_gradv, d_rho, eta, zeta, d_aux8);
    if (h_uniform)
        viscous_kernel<true><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, b->vel4, d_aux8, b->rel4, b->perm,
                                                   b->nbr, b->cnt, b->status, d_h_orig, list_fresh, fcutsq, d_vdot,
                                                   d_udot);
    else
        viscous_kernel<false><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, b->vel4, d_aux8, b->rel4, b->perm,
                                                    b->nbr, b->cnt, b->status, d_h_orig, list_fresh, fcutsq, d_vdot,
                                                    d_udot);
    return status_of_launch();
}

int sph_gradient(const sph_grid *g, const sph_buffers *b, const double *d_f, const double *d_wgt, int subtract_self,
                 const double *d_h_orig, int h_uniform, int list_fresh, double *d_aux4, double *d_out, void *stream)
{
    if (!g || !b || !d_wgt || !d_h_orig || !d_aux4 || !d_out) return SPH_E_BADARG;
    if (!b->pos4 || !b->rel4 || !b->perm || !b->nbr || !b->cnt || !b->status) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_of(b->n);
    scalar_term_kernel<<<nb, kBlock, 0, s>>>(b->n, b->perm, d_f, d_wgt, d_aux4);
    if (h_uniform)
        gradient_kernel<true><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, d_aux4, b->rel4, b->perm, b->nbr,
                                                    b->cnt, b->status, d_h_orig, list_fresh, subtract_self, d_out);
    else
        gradient_kernel<false><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, d_aux4, b->rel4, b->perm, b->nbr,
                                                     b->cnt, b->status, d_h_orig, list_fresh, subtract_self, d_out);
    return status_of_launch();
}

int sph_stress_force(const sph_grid *g, const sph_buffers *b, const double *d_stress, const double *d_rho,
                     const double *d_h_orig, int h_uniform, int list_fresh, double fcutoff, double *d_aux8,
                     double *d_vdot, double *d_udot, void *stream)
{
    if (!g || !b || !d_stress || !d_rho || !d_h_orig || !d_aux8 || !d_vdot || !d_udot) return SPH_E_BADARG;
    if (!b->pos4 || !b->vel4 || !b->rel4 || !b->perm || !b->nbr || !b->cnt || !b->status) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_of(b->n);
    const double fcutsq = fcutoff * fcutoff;                         // forces.py:36
    stress_pack_kernel<<<nb, kBlock, 0, s>>>(b->n, b->perm, d_stress, d_rho, d_aux8);
    if (h_uniform)
        viscous_kernel<true><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, b->vel4, d_aux8, b->rel4, b->perm,
                                                   b->nbr, b->cnt, b->status, d_h_orig, list_fresh, fcutsq, d_vdot,
                                                   d_udot);
    else
        viscous_kernel<false><<<nb, kBlock, 0, s>>>(*g, b->n, b->max_nbrs, b->pos4, b->vel4, d_aux8, b->rel4, b->perm,
                                                    b->nbr, b->cnt, b->status, d_h_orig, list_fresh, fcutsq, d_vdot,
                                                    d_udot);
    return status_of_launch();
}

int sph_core_force(const sph_grid *g, const sph_buffers *b, double sigma, double rcoef, int list_fresh, double *d_vdot,
                   double *d_udot, void *stream)
{
    if (!g || !b || !d_vdot || !d_udot || !(sigma > 0.0)) return SPH_E_BADARG;
    if (!b->pos4 || !b->vel4 || !b->rel4 || !b->perm || !b->nbr || !b->cnt || !b->status) return SPH_E_BADARG;
    if (b->n == 0) return SPH_OK;
    const double inv_s2 = 1.0 / (sigma * sigma);
    core_kernel<<<blocks_of(b->n), kBlock, 0, (cudaStream_t)stream>>>(
        *g, b->n, b->max_nbrs, b->pos4, b->vel4, b->rel4, b->perm, b->nbr, b->cnt, b->status, list_fresh, inv_s2,
        -8.0 * rcoef * inv_s2, d_vdot, d_udot);
    return status_of_launch();
}

}  // extern "C"
