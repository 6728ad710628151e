// Temporal blocking: TWO time steps per launch, i.e. per HBM round trip of the state.
//
// FK_DTI at the 2x up-sampled grid (BASELINE.json configs[4], 28 M voxels) streams 5 words per voxel-update
// and runs at the HBM roofline with the one-step kernels; the only way past it is to read the state once
// for several steps.  dti_step2_tb marches along x like the one-step kernels and keeps two pipelines in
// flight on overlapped tiles:
//     U1[x]   = step(U0[x-1], U0[x], U0[x+1])      on the tile extended by one voxel in y and z (E1)
//     U2[x-1] = step(U1[x-2], U1[x-1], U1[x])      on the tile itself, written to the other state buffer
// Every thread owns one (y, z) column of E1: 32 x TYT threads, of which the inner 30 x (TYT-2) also produce
// output.  U0 of the current plane goes through a shared-memory copy of E1 plus a ring loaded by designated
// threads (E2), U1 through a second shared-memory plane; both double buffered, two barriers per plane.
// The x-neighbours and the coefficients stay in registers.  Columns outside the box are periodic images
// (wrapped coordinates), so ragged tiles and np.roll semantics need no special cases.
//
// Every arithmetic operation is the one-step kernel's (dti_step_v4's `one`, individually rounded), hence the
// result after two steps is bitwise the result of two one-step launches; only the association of the
// per-step volume sums differs (per-block partials of a different grid).
#pragma once
#include "tgtk_kernels_v3.cuh"

namespace tgtk {

// per-block partial sums of the launch's two steps into ring slots a.slot and a.slot + 1
template <typename T> __device__ __forceinline__ void finish_light_pair(const StepArgs<T> &a, double s1, double s2)
{
    __shared__ double red3[2][32];
    const int tid = threadIdx.x + blockDim.x * threadIdx.y;
    const int nthreads = blockDim.x * blockDim.y;
    const int nwarps = (nthreads + 31) >> 5;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { red3[0][warp] = s1; red3[1][warp] = s2; }
    __syncthreads();
    if (warp == 0) {
        double v1 = (lane < nwarps) ? red3[0][lane] : 0.0;
        double v2 = (lane < nwarps) ? red3[1][lane] : 0.0;
        v1 = warp_sum(v1);
        v2 = warp_sum(v2);
        if (lane == 0) {
            a.partials[(size_t)a.slot * a.ring_stride + bid] = v1;
            a.partials[(size_t)(a.slot + 1) * a.ring_stride + bid] = v2;
        }
    }
}

template <typename T, int TYT, int MINB>
__global__ void __launch_bounds__(32 * TYT, MINB) dti_step2_tb(const StepArgs<T> a, const int lx)
{
    using R = Rn<T>;
    constexpr int TZT = 32, OZ = TZT - 2, OY = TYT - 2;
    constexpr int RS0 = TZT + 2, PL0 = (TYT + 2) * RS0;      // U0 plane: E1 + ring
    constexpr int RS1 = TZT, PL1 = TYT * RS1;                // U1 plane: E1
    __shared__ T s0[2][PL0];
    __shared__ T s1[2][PL1];
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ dxa = a.coef[0];
    const T *__restrict__ dya = a.coef[1];
    const T *__restrict__ dza = a.coef[2];
    const int tz = threadIdx.x, ty = threadIdx.y, tid = ty * TZT + tz;
    const int z0 = blockIdx.x * OZ, y0 = blockIdx.y * OY;
    const int zc = z0 - 1 + tz, yc = y0 - 1 + ty;            // this thread's column of E1
    const bool is_out = tz >= 1 && tz <= OZ && ty >= 1 && ty <= OY && zc < a.Z && yc < a.Y;
    const int x0 = blockIdx.z * lx, x1 = min(x0 + lx, a.X);
    const uint32_t sx = (uint32_t)a.sx, sy = (uint32_t)a.sy;
    const uint32_t col = (uint32_t)wrap_coord(yc, a.Y) * sy + (uint32_t)wrap_coord(zc, a.Z);
    // ring of E1 (the part of E2 no thread owns): TZT above, TZT below, TYT left, TYT right
    const bool ring = tid < 2 * TZT + 2 * TYT;
    int hr = 1, hc = 1;
    if (tid < TZT) { hr = 0; hc = tid + 1; }
    else if (tid < 2 * TZT) { hr = TYT + 1; hc = tid - TZT + 1; }
    else if (tid < 2 * TZT + TYT) { hr = tid - 2 * TZT + 1; hc = 0; }
    else if (ring) { hr = tid - 2 * TZT - TYT + 1; hc = TZT + 1; }
    const int hs = hr * RS0 + hc;
    const uint32_t hcol = (uint32_t)wrap_coord(y0 - 2 + hr, a.Y) * sy + (uint32_t)wrap_coord(z0 - 2 + hc, a.Z);
    const int o0 = (ty + 1) * RS0 + tz + 1, o1 = ty * RS1 + tz;
    auto plane = [&](int x) { return (uint32_t)wrap_coord(x, a.X) * sx; };
    // L2 prefetch of the rows this y-tile reads (clipped to the box), by one thread of the blockIdx.x == 0 tile
    L2Strip pf;
    {
        const int r0 = max(y0 - 2, 0), r1 = min(y0 + TYT, a.Y);
        pf.on = a.l2pf > 0 && blockIdx.x == 0 && tid == 0 && r1 > r0;
        pf.row0 = (uint32_t)r0 * sy;
        pf.count = (uint32_t)max(r1 - r0, 0) * sy;
        pf.plane = sx;
        pf.total = (uint32_t)a.X * sx;
    }
    // FK_DTI.py:39-45 + FK.py:36-41, D_plus aliasing D_minus -- the arithmetic of dti_step_v4
    auto one = [&](T c, T d0, T d1, T d2, T um, T up, T uym, T uyp, T uzm, T uzp) {
        const T fx = R::mul(a.cx, R::sub(R::mul(d0, R::sub(um, c)), R::mul(d0, R::sub(c, up))));
        const T fy = R::mul(a.cy, R::sub(R::mul(d1, R::sub(uym, c)), R::mul(d1, R::sub(c, uyp))));
        const T fz = R::mul(a.cz, R::sub(R::mul(d2, R::sub(uzm, c)), R::mul(d2, R::sub(c, uzp))));
        const T sp = R::add(R::add(fx, fy), fz);
        const T react = R::mul(a.rho, R::mul(c, R::sub(T(1), c)));
        return R::add(c, R::mul(R::add(sp, react), a.dt));
    };
    double sum1 = 0.0, sum2 = 0.0;
    griddep_launch();
    griddep_wait();

    // pipeline state when U1[xi] is computed: U0 of planes xi-1, xi, xi+1 (own column), the coefficients of
    // planes xi (step 1) and xi-1 (step 2), U1 of planes xi-2, xi-1 (own column)
    int xi = x0 - 1;
    T u0m = u[plane(xi - 1) + col], u0c = u[plane(xi) + col], u0p = u[plane(xi + 1) + col];
    T c0 = dxa[plane(xi) + col], c1 = dya[plane(xi) + col], c2 = dza[plane(xi) + col];
    // plane offsets advance incrementally (one compare per plane instead of a modulo)
    const uint32_t last = (uint32_t)(a.X - 1) * sx;
    auto adv = [&](uint32_t o) { return (o == last) ? 0u : o + sx; };
    uint32_t pw = plane(xi - 1), pn = plane(xi + 1), pn2 = plane(xi + 2);   // planes xi-1 (written), xi+1, xi+2
    T m0 = T(0), m1 = T(0), m2 = T(0);
    T u1mm = T(0), u1m = T(0);
    T hu = T(0);
    if (ring) hu = u[plane(xi) + hcol];
    int buf = 0;
    for (; xi <= x1; ++xi) {
        T *t0 = s0[buf], *t1 = s1[buf];
        t0[o0] = u0c;
        if (ring) t0[hs] = hu;
        // requests for the next iteration: U0[xi+2], the coefficients and the ring of plane xi+1
        const T u0n = u[pn2 + col];
        const T n0 = dxa[pn + col], n1 = dya[pn + col], n2 = dza[pn + col];
        if (ring) hu = u[pn + hcol];
        if (pf.on) {
            const uint32_t q = plane(xi + 2 + a.l2pf) + pf.row0;
            l2_prefetch_strip(u, q, pf.count, pf.total);
            const uint32_t qc = plane(xi + 1 + a.l2pf) + pf.row0;
            l2_prefetch_strip(dxa, qc, pf.count, pf.total); l2_prefetch_strip(dya, qc, pf.count, pf.total);
            l2_prefetch_strip(dza, qc, pf.count, pf.total);
        }
        __syncthreads();
        const T u1c = one(u0c, c0, c1, c2, u0m, u0p, t0[o0 - RS0], t0[o0 + RS0], t0[o0 - 1], t0[o0 + 1]);
        t1[o1] = u1c;
        if (is_out && xi >= x0 && xi < x1 && counts(a, xi)) sum1 += (double)u1c;
        __syncthreads();
        if (is_out && xi > x0) {             // U2[xi-1] from U1[xi-2], U1[xi-1], U1[xi]
            const T *tp = s1[buf ^ 1];
            const T u2 = one(u1m, m0, m1, m2, u1mm, u1c, tp[o1 - RS1], tp[o1 + RS1], tp[o1 - 1], tp[o1 + 1]);
            out[pw + col] = u2;
            if (counts(a, xi - 1)) sum2 += (double)u2;
        }
        u0m = u0c; u0c = u0p; u0p = u0n;
        m0 = c0; m1 = c1; m2 = c2;
        c0 = n0; c1 = n1; c2 = n2;
        u1mm = u1m; u1m = u1c;
        pw = adv(pw); pn = adv(pn); pn2 = adv(pn2);
        buf ^= 1;
    }
    finish_light_pair(a, sum1, sum2);
}

// FK with two time steps per launch: the same two pipelines as dti_step2_tb.  The face coefficients are
// relu(k(c) + k(c')) (tgtk_kernels_v3.cuh), so the U0 plane in shared memory carries (u, k) pairs on the
// extended tile plus ring, and the U1 plane carries (u1, k) pairs too (k copied along: U2's neighbours then
// come with one 128-bit read each, and the U0 buffer that the next iteration overwrites is never read late).
// Own-column k of planes x-2 .. x+1 stays in registers.
template <typename T, int TYT, int MINB>
__global__ void __launch_bounds__(32 * TYT, MINB) fk_step2_tb(const StepArgs<T> a, const int lx)
{
    constexpr int TZT = 32, OZ = TZT - 2, OY = TYT - 2;
    constexpr int RS0 = TZT + 2, PL0 = (TYT + 2) * RS0;
    constexpr int RS1 = TZT, PL1 = TYT * RS1;
    __shared__ Pair<T> s0[2][PL0];         // (u0, k) of plane xi on E1 + ring
    __shared__ Pair<T> s1[2][PL1];         // (u1, k) of plane xi on E1
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ kc = a.coef[0];
    const int tz = threadIdx.x, ty = threadIdx.y, tid = ty * TZT + tz;
    const int z0 = blockIdx.x * OZ, y0 = blockIdx.y * OY;
    const int zc = z0 - 1 + tz, yc = y0 - 1 + ty;
    const bool is_out = tz >= 1 && tz <= OZ && ty >= 1 && ty <= OY && zc < a.Z && yc < a.Y;
    const int x0 = blockIdx.z * lx, x1 = min(x0 + lx, a.X);
    const uint32_t sx = (uint32_t)a.sx, sy = (uint32_t)a.sy;
    const uint32_t col = (uint32_t)wrap_coord(yc, a.Y) * sy + (uint32_t)wrap_coord(zc, a.Z);
    const bool ring = tid < 2 * TZT + 2 * TYT;
    int hr = 1, hc = 1;
    if (tid < TZT) { hr = 0; hc = tid + 1; }
    else if (tid < 2 * TZT) { hr = TYT + 1; hc = tid - TZT + 1; }
    else if (tid < 2 * TZT + TYT) { hr = tid - 2 * TZT + 1; hc = 0; }
    else if (ring) { hr = tid - 2 * TZT - TYT + 1; hc = TZT + 1; }
    const int hs = hr * RS0 + hc;
    const uint32_t hcol = (uint32_t)wrap_coord(y0 - 2 + hr, a.Y) * sy + (uint32_t)wrap_coord(z0 - 2 + hc, a.Z);
    const int o0 = (ty + 1) * RS0 + tz + 1, o1 = ty * RS1 + tz;
    auto plane = [&](int x) { return (uint32_t)wrap_coord(x, a.X) * sx; };
    L2Strip pf;
    {
        const int r0 = max(y0 - 2, 0), r1 = min(y0 + TYT, a.Y);
        pf.on = a.l2pf > 0 && blockIdx.x == 0 && tid == 0 && r1 > r0;
        pf.row0 = (uint32_t)r0 * sy;
        pf.count = (uint32_t)max(r1 - r0, 0) * sy;
        pf.plane = sx;
        pf.total = (uint32_t)a.X * sx;
    }
    // FK.py:36-41 with the face coefficients of FK.py:12-26 in the compact form of fk_step_v4
    auto one = [&](T c, T k, T um, T km, T up, T kp, T uym, T kym, T uyp, T kyp, T uzm, T kzm, T uzp, T kzp) {
        T sp = a.cx * (relu(km + k) * (um - c) - relu(k + kp) * (c - up));
        sp += a.cy * (relu(kym + k) * (uym - c) - relu(k + kyp) * (c - uyp));
        sp += a.cz * (relu(kzm + k) * (uzm - c) - relu(k + kzp) * (c - uzp));
        return c + a.dt * (sp + a.rho * c * (T(1) - c));
    };
    double sum1 = 0.0, sum2 = 0.0;
    griddep_launch();
    griddep_wait();

    int xi = x0 - 1;
    T u0m = u[plane(xi - 1) + col], u0c = u[plane(xi) + col], u0p = u[plane(xi + 1) + col];
    T kmm = kc[plane(xi - 2) + col], km = kc[plane(xi - 1) + col], kk = kc[plane(xi) + col], kp = kc[plane(xi + 1) + col];
    const uint32_t last = (uint32_t)(a.X - 1) * sx;
    auto adv = [&](uint32_t o) { return (o == last) ? 0u : o + sx; };
    uint32_t pw = plane(xi - 1), pn = plane(xi + 1), pn2 = plane(xi + 2);
    T u1mm = T(0), u1m = T(0);
    T hu = T(0), hk = T(0);
    if (ring) { hu = u[plane(xi) + hcol]; hk = kc[plane(xi) + hcol]; }
    int buf = 0;
    for (; xi <= x1; ++xi) {
        Pair<T> *t0 = s0[buf], *t1 = s1[buf];
        t0[o0] = Pair<T>{u0c, kk};
        if (ring) t0[hs] = Pair<T>{hu, hk};
        const T u0n = u[pn2 + col], kn = kc[pn2 + col];
        if (ring) { hu = u[pn + hcol]; hk = kc[pn + hcol]; }
        if (pf.on) {
            const uint32_t q = plane(xi + 2 + a.l2pf) + pf.row0;
            l2_prefetch_strip(u, q, pf.count, pf.total); l2_prefetch_strip(kc, q, pf.count, pf.total);
        }
        __syncthreads();
        const Pair<T> ym = t0[o0 - RS0], yp = t0[o0 + RS0], zm = t0[o0 - 1], zp = t0[o0 + 1];
        const T u1c = one(u0c, kk, u0m, km, u0p, kp, ym.a, ym.b, yp.a, yp.b, zm.a, zm.b, zp.a, zp.b);
        t1[o1] = Pair<T>{u1c, kk};
        if (is_out && xi >= x0 && xi < x1 && counts(a, xi)) sum1 += (double)u1c;
        __syncthreads();
        if (is_out && xi > x0) {             // U2[xi-1]: u1 of planes xi-2, xi-1, xi; k of the same planes
            const Pair<T> *tp = s1[buf ^ 1];
            const Pair<T> vm = tp[o1 - RS1], vp = tp[o1 + RS1], wm = tp[o1 - 1], wp = tp[o1 + 1];
            const T u2 = one(u1m, km, u1mm, kmm, u1c, kk, vm.a, vm.b, vp.a, vp.b, wm.a, wm.b, wp.a, wp.b);
            out[pw + col] = u2;
            if (counts(a, xi - 1)) sum2 += (double)u2;
        }
        u0m = u0c; u0c = u0p; u0p = u0n;
        kmm = km; km = kk; kk = kp; kp = kn;
        u1mm = u1m; u1m = u1c;
        pw = adv(pw); pn = adv(pn); pn2 = adv(pn2);
        buf ^= 1;
    }
    finish_light_pair(a, sum1, sum2);
}

}  // namespace tgtk

// =====================================================================================================
// Full-tensor FK_DTI (19-point stencil, the opt-in mode of dti_full_step_v1) as a staged marching kernel.
// The mixed-derivative terms read the four in-plane diagonals of plane x and the y/z neighbours of planes
// x-1 and x+1, so the block keeps a rolling window of FOUR shared-memory planes (tile + a complete ring,
// corners included): three are read while the fourth is being published for the next plane -- one barrier
// per plane, no register copy of the neighbours.  The own voxel's u and six tensor components of plane x+1
// are requested one plane ahead.  Expressions and their order are dti_full_step_v1's: bitwise equal to it.
// =====================================================================================================
namespace tgtk {

template <typename T, int TZ, int TY>
__global__ void __launch_bounds__(TZ *TY) dti_full_step_v3(const StepArgs<T> a, const int lx)
{
    using R = Rn<T>;
    constexpr int RS = TZ + 2, PL = (TY + 2) * RS, NRING = 2 * RS + 2 * TY;
    __shared__ T tile[4][PL];
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const int tz = threadIdx.x, ty = threadIdx.y, tid = ty * TZ + tz;
    const int z0 = blockIdx.x * TZ, y0 = blockIdx.y * TY;
    const bool valid = (z0 + tz < a.Z) && (y0 + ty < a.Y);
    const int x0 = blockIdx.z * lx, x1 = min(x0 + lx, a.X);
    const uint32_t sx = (uint32_t)a.sx, sy = (uint32_t)a.sy;
    const uint32_t col = (uint32_t)wrap_coord(y0 + ty, a.Y) * sy + (uint32_t)wrap_coord(z0 + tz, a.Z);
    // ring with corners: RS above, RS below, TY left, TY right
    const bool ring = tid < NRING;
    int hr = 1, hc = 1;
    if (tid < RS) { hr = 0; hc = tid; }
    else if (tid < 2 * RS) { hr = TY + 1; hc = tid - RS; }
    else if (tid < 2 * RS + TY) { hr = tid - 2 * RS + 1; hc = 0; }
    else if (ring) { hr = tid - 2 * RS - TY + 1; hc = TZ + 1; }
    const int hs = hr * RS + hc;
    const uint32_t hcol = (uint32_t)wrap_coord(y0 - 1 + hr, a.Y) * sy + (uint32_t)wrap_coord(z0 - 1 + hc, a.Z);
    const int o = (ty + 1) * RS + tz + 1;
    const uint32_t last = (uint32_t)(a.X - 1) * sx;
    auto adv = [&](uint32_t p) { return (p == last) ? 0u : p + sx; };
    auto plane = [&](int x) { return (uint32_t)wrap_coord(x, a.X) * sx; };
    double s = 0.0;
    griddep_launch();
    griddep_wait();

    // planes x0-1 and x0 go straight into the window; plane x0+1 waits in registers
    uint32_t pc = plane(x0), pn = plane(x0 + 1);
    {
        const uint32_t pm = plane(x0 - 1);
        tile[(x0 + 3) & 3][o] = u[pm + col];        // slot of plane x is x & 3; x0-1 = x0+3 (mod 4)
        tile[x0 & 3][o] = u[pc + col];
        if (ring) {
            tile[(x0 + 3) & 3][hs] = u[pm + hcol];
            tile[x0 & 3][hs] = u[pc + hcol];
        }
    }
    T un = u[pn + col], hn = T(0);
    if (ring) hn = u[pn + hcol];
    T d[6], dnext[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) d[q] = a.coef[q][pc + col];
    for (int x = x0; x < x1; ++x) {
        T *tp = tile[(x + 1) & 3];
        tp[o] = un;
        if (ring) tp[hs] = hn;
        const uint32_t pn2 = adv(pn);
        un = u[pn2 + col];
        if (ring) hn = u[pn2 + hcol];
#pragma unroll
        for (int q = 0; q < 6; ++q) dnext[q] = a.coef[q][pn + col];
        __syncthreads();
        const T *tm = tile[(x + 3) & 3], *tc = tile[x & 3];
        const T uc = tc[o];
        const T fx = R::mul(a.cx, R::sub(R::mul(d[0], R::sub(tm[o], uc)), R::mul(d[0], R::sub(uc, tp[o]))));
        const T fy = R::mul(a.cy, R::sub(R::mul(d[1], R::sub(tc[o - RS], uc)), R::mul(d[1], R::sub(uc, tc[o + RS]))));
        const T fz = R::mul(a.cz, R::sub(R::mul(d[2], R::sub(tc[o - 1], uc)), R::mul(d[2], R::sub(uc, tc[o + 1]))));
        const T sp = R::add(R::add(fx, fy), fz);
        // mixed terms (u[+a,+b] - u[+a,-b]) - (u[-a,+b] - u[-a,-b]); ex, ey, ez carry 2/(4 h_a h_b)
        const T mxy = R::sub(R::sub(tp[o + RS], tp[o - RS]), R::sub(tm[o + RS], tm[o - RS]));
        const T mxz = R::sub(R::sub(tp[o + 1], tp[o - 1]), R::sub(tm[o + 1], tm[o - 1]));
        const T myz = R::sub(R::sub(tc[o + RS + 1], tc[o + RS - 1]), R::sub(tc[o - RS + 1], tc[o - RS - 1]));
        const T cross = R::add(R::add(R::mul(R::mul(a.ex, d[3]), mxy), R::mul(R::mul(a.ey, d[4]), mxz)),
                               R::mul(R::mul(a.ez, d[5]), myz));
        const T unew = R::add(uc, R::mul(R::add(R::add(sp, cross), R::mul(a.rho, R::mul(uc, R::sub(T(1), uc)))), a.dt));
        if (valid) {
            out[pc + col] = unew;
            if (counts(a, x)) s += (double)unew;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) d[q] = dnext[q];
        pc = pn; pn = pn2;
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

}  // namespace tgtk
