// Baseline step kernels: one voxel per thread, neighbours straight from global memory
// through L1/L2.  They are the simplest correct statement of each operator on the GPU and
// stay in the library as the in-device cross-check of the tuned kernels (tgtk_kernels_v2.cuh).
#pragma once
#include "tgtk_common.cuh"

namespace tgtk {

struct Nbr {
    int64_t c, xm, xp, ym, yp, zm, zp;
};

template <typename T>
__device__ __forceinline__ bool voxel_index(const StepArgs<T> &a, int &i, int &j, int &k, Nbr &n)
{
    k = blockIdx.x * blockDim.x + threadIdx.x;
    j = blockIdx.y * blockDim.y + threadIdx.y;
    i = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= a.X || j >= a.Y || k >= a.Z) return false;
    // periodic neighbours (np.roll): FK/FK.py:36-38
    const int im = (i == 0) ? a.X - 1 : i - 1, ip = (i + 1 == a.X) ? 0 : i + 1;
    const int jm = (j == 0) ? a.Y - 1 : j - 1, jp = (j + 1 == a.Y) ? 0 : j + 1;
    const int km = (k == 0) ? a.Z - 1 : k - 1, kp = (k + 1 == a.Z) ? 0 : k + 1;
    const int64_t row = (int64_t)i * a.sx + (int64_t)j * a.sy;
    n.c = row + k;
    n.xm = (int64_t)im * a.sx + (int64_t)j * a.sy + k;
    n.xp = (int64_t)ip * a.sx + (int64_t)j * a.sy + k;
    n.ym = (int64_t)i * a.sx + (int64_t)jm * a.sy + k;
    n.yp = (int64_t)i * a.sx + (int64_t)jp * a.sy + k;
    n.zm = row + km;
    n.zp = row + kp;
    return true;
}

// ---- FK, compact coefficients ------------------------------------------------------------
// coef[0] = k(c) = (WM + GM/ratio)/2 where WM+GM >= th_matter, else -Big.  The face
// coefficient of FK/FK.py:10-26 is Dw*max(k(c)+k(c'),0); Dw/h^2 is folded into cx,cy,cz.
template <typename T> __global__ void __launch_bounds__(256) fk_step_v1(const StepArgs<T> a)
{
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ o = a.buf[sc.par ^ 1][0];
    const T *__restrict__ kc = a.coef[0];
    int i, j, k;
    Nbr n;
    double s = 0.0;
    if (voxel_index(a, i, j, k, n)) {
        const T uc = u[n.c], kk = kc[n.c];
        const T fxm = tg_max(kc[n.xm] + kk, T(0)), fxp = tg_max(kk + kc[n.xp], T(0));
        const T fym = tg_max(kc[n.ym] + kk, T(0)), fyp = tg_max(kk + kc[n.yp], T(0));
        const T fzm = tg_max(kc[n.zm] + kk, T(0)), fzp = tg_max(kk + kc[n.zp], T(0));
        T sp = a.cx * (fxm * (u[n.xm] - uc) - fxp * (uc - u[n.xp]));
        sp += a.cy * (fym * (u[n.ym] - uc) - fyp * (uc - u[n.yp]));
        sp += a.cz * (fzm * (u[n.zm] - uc) - fzp * (uc - u[n.zp]));
        const T un = uc + a.dt * (sp + a.rho * uc * (T(1) - uc));
        o[n.c] = un;
        if (counts(a, i)) s = (double)un;
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- explicit coefficient arrays, reference rounding order --------------------------------
// kMode 0: COEF6   coef[0..2] = D_plus_{x,y,z}, coef[3..5] = D_minus_{x,y,z}    FK/FK.py:34-42
// kMode 1: FACES   coef[0..2] = D_minus_{x,y,z}; D_plus is D_minus one voxel down FK/FK.py:28-30
// kMode 2: DTI     coef[0..2] = D_{x,y,z} = rgb*Dw; D_plus aliases D_minus  FK_DTI/FK_DTI.py:39-45
template <typename T, int kMode> __global__ void __launch_bounds__(256) coef_step_v1(const StepArgs<T> a)
{
    using R = Rn<T>;
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ o = a.buf[sc.par ^ 1][0];
    int i, j, k;
    Nbr n;
    double s = 0.0;
    if (voxel_index(a, i, j, k, n)) {
        const T uc = u[n.c];
        T dpx, dpy, dpz, dmx, dmy, dmz;
        if (kMode == 0) {
            dpx = a.coef[0][n.c]; dpy = a.coef[1][n.c]; dpz = a.coef[2][n.c];
            dmx = a.coef[3][n.c]; dmy = a.coef[4][n.c]; dmz = a.coef[5][n.c];
        } else if (kMode == 1) {
            dmx = a.coef[0][n.c]; dmy = a.coef[1][n.c]; dmz = a.coef[2][n.c];
            dpx = a.coef[0][n.xm]; dpy = a.coef[1][n.ym]; dpz = a.coef[2][n.zm];
        } else {
            dmx = dpx = a.coef[0][n.c]; dmy = dpy = a.coef[1][n.c]; dmz = dpz = a.coef[2][n.c];
        }
        // 1/(h*h) * (D_plus*(u[c-1]-u) - D_minus*(u-u[c+1]))
        const T sx = R::mul(a.cx, R::sub(R::mul(dpx, R::sub(u[n.xm], uc)), R::mul(dmx, R::sub(uc, u[n.xp]))));
        const T sy = R::mul(a.cy, R::sub(R::mul(dpy, R::sub(u[n.ym], uc)), R::mul(dmy, R::sub(uc, u[n.yp]))));
        const T sz = R::mul(a.cz, R::sub(R::mul(dpz, R::sub(u[n.zm], uc)), R::mul(dmz, R::sub(uc, u[n.zp]))));
        const T sp = R::add(R::add(sx, sy), sz);
        // A += (SP + f*A*(1-A)) * dt
        const T un = R::add(uc, R::mul(R::add(sp, R::mul(a.rho, R::mul(uc, R::sub(T(1), uc)))), a.dt));
        o[n.c] = un;
        if (counts(a, i)) s = (double)un;
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK_DTI with the FULL diffusion tensor (opt-in; NOT in the packaged reference) ------------
// du/dt = sum_a D_aa d2_a u + 2 sum_{a<b} D_ab d_a d_b u + rho u (1-u): the reference's diagonal
// operator (FK_DTI/FK_DTI.py:39-45 + FK/FK.py:36-41, evaluated exactly as coef_step_v1 mode 2)
// plus the mixed-derivative terms on the 12 edge neighbours (19-point stencil),
//   d_a d_b u = (u[+a,+b] - u[+a,-b] - u[-a,+b] + u[-a,-b]) / (4 h_a h_b).
// coef[0..2] = D_xx, D_yy, D_zz ; coef[3..5] = D_xy, D_xz, D_yz (all already times Dw).
// With zero off-diagonals the result equals the reference operator bit for bit.  The only code in
// the reference that touches off-diagonals is an unpackaged scratch script
// (FK_DTI/testNewDiffusionStepWithTensor.py:101-119): parity of this mode is unpinned.
template <typename T> __global__ void __launch_bounds__(256) dti_full_step_v1(const StepArgs<T> a)
{
    using R = Rn<T>;
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ o = a.buf[sc.par ^ 1][0];
    int i, j, k;
    Nbr n;
    double s = 0.0;
    if (voxel_index(a, i, j, k, n)) {
        const int im = (i == 0) ? a.X - 1 : i - 1, ip = (i + 1 == a.X) ? 0 : i + 1;
        const int jm = (j == 0) ? a.Y - 1 : j - 1, jp = (j + 1 == a.Y) ? 0 : j + 1;
        const int km = (k == 0) ? a.Z - 1 : k - 1, kp = (k + 1 == a.Z) ? 0 : k + 1;
        auto at = [&](int x, int y, int z) { return u[(int64_t)x * a.sx + (int64_t)y * a.sy + z]; };
        const T uc = u[n.c];
        const T dxx = a.coef[0][n.c], dyy = a.coef[1][n.c], dzz = a.coef[2][n.c];
        const T dxy = a.coef[3][n.c], dxz = a.coef[4][n.c], dyz = a.coef[5][n.c];
        const T sx = R::mul(a.cx, R::sub(R::mul(dxx, R::sub(u[n.xm], uc)), R::mul(dxx, R::sub(uc, u[n.xp]))));
        const T sy = R::mul(a.cy, R::sub(R::mul(dyy, R::sub(u[n.ym], uc)), R::mul(dyy, R::sub(uc, u[n.yp]))));
        const T sz = R::mul(a.cz, R::sub(R::mul(dzz, R::sub(u[n.zm], uc)), R::mul(dzz, R::sub(uc, u[n.zp]))));
        const T sp = R::add(R::add(sx, sy), sz);
        // mixed terms; ex, ey, ez carry 2/(4 h_x h_y), 2/(4 h_x h_z), 2/(4 h_y h_z)
        const T mxy = R::sub(R::sub(at(ip, jp, k), at(ip, jm, k)), R::sub(at(im, jp, k), at(im, jm, k)));
        const T mxz = R::sub(R::sub(at(ip, j, kp), at(ip, j, km)), R::sub(at(im, j, kp), at(im, j, km)));
        const T myz = R::sub(R::sub(at(i, jp, kp), at(i, jp, km)), R::sub(at(i, jm, kp), at(i, jm, km)));
        const T cross = R::add(R::add(R::mul(R::mul(a.ex, dxy), mxy), R::mul(R::mul(a.ey, dxz), mxz)),
                               R::mul(R::mul(a.ez, dyz), myz));
        const T un = R::add(uc, R::mul(R::add(R::add(sp, cross), R::mul(a.rho, R::mul(uc, R::sub(T(1), uc)))), a.dt));
        o[n.c] = un;
        if (counts(a, i)) s = (double)un;
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK_2c, compact coefficients ----------------------------------------------------------
// coef[0] = k(c)  (P diffusivity half-sum, -Big outside tissue), coef[1] = b(c) = (WM+GM)/2
// (-Big outside tissue).  The necrotic closure (P+N > th_necro, FK_2c/FK_2c.py:13-15) is
// evaluated from the CURRENT P,N of both voxels of a face, i.e. the get_D_with_necro call of
// FK_2c.py:230 fused into the next step.  Ordering of FK_2c.py:58-71: stencils on old P,S;
// N and the S sink use the new P.
template <typename T> __global__ void __launch_bounds__(256) fk2c_step_v1(const StepArgs<T> a)
{
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ P = a.buf[sc.par][0];
    const T *__restrict__ N = a.buf[sc.par][1];
    const T *__restrict__ S = a.buf[sc.par][2];
    T *__restrict__ Po = a.buf[sc.par ^ 1][0];
    T *__restrict__ No = a.buf[sc.par ^ 1][1];
    T *__restrict__ So = a.buf[sc.par ^ 1][2];
    const T *__restrict__ kc = a.coef[0];
    const T *__restrict__ bc = a.coef[1];
    int i, j, k;
    Nbr n;
    double sP = 0.0, sN = 0.0;
    if (voxel_index(a, i, j, k, n)) {
        const T nb = -Big<T>::v;
        const T p = P[n.c], nn = N[n.c], s = S[n.c];
        auto keff = [&](int64_t q, T pq) { return (pq + N[q] <= a.th_necro) ? kc[q] : nb; };
        const T pxm = P[n.xm], pxp = P[n.xp], pym = P[n.ym], pyp = P[n.yp], pzm = P[n.zm], pzp = P[n.zp];
        const T kk = (p + nn <= a.th_necro) ? kc[n.c] : nb;
        T sp = a.cx * (tg_max(keff(n.xm, pxm) + kk, T(0)) * (pxm - p) - tg_max(kk + keff(n.xp, pxp), T(0)) * (p - pxp));
        sp += a.cy * (tg_max(keff(n.ym, pym) + kk, T(0)) * (pym - p) - tg_max(kk + keff(n.yp, pyp), T(0)) * (p - pyp));
        sp += a.cz * (tg_max(keff(n.zm, pzm) + kk, T(0)) * (pzm - p) - tg_max(kk + keff(n.zp, pzp), T(0)) * (p - pzp));
        const T bb = bc[n.c];
        T ss = a.ex * (tg_max(bc[n.xm] + bb, T(0)) * (S[n.xm] - s) - tg_max(bb + bc[n.xp], T(0)) * (s - S[n.xp]));
        ss += a.ey * (tg_max(bc[n.ym] + bb, T(0)) * (S[n.ym] - s) - tg_max(bb + bc[n.yp], T(0)) * (s - S[n.yp]));
        ss += a.ez * (tg_max(bc[n.zm] + bb, T(0)) * (S[n.zm] - s) - tg_max(bb + bc[n.zp], T(0)) * (s - S[n.zp]));
        const T H = heaviside50(s, a.sigma_np);
        const T pn = p + a.dt * (sp + a.rho * (s * p) * ((T(1) - p) - nn) - (a.lambda_np * p) * H);
        const T nnew = nn + a.dt * ((a.lambda_np * pn) * H);
        const T snew = s + a.dt * (ss - (a.lambda_s * s) * pn);
        Po[n.c] = pn;
        No[n.c] = nnew;
        So[n.c] = snew;
        if (counts(a, i)) {
            sP = (double)pn;
            sN = (double)nnew;
        }
    }
    step_end<T, 2>(a, sc, sP, sN);
}

// ---- FK_2c with twelve explicit coefficient arrays, reference rounding order ---------------
// coef[0..5] = D_plus_{x,y,z}, D_minus_{x,y,z} for P; coef[6..11] the same for S
// (the public FK_2c_update method, FK_2c/FK_2c.py:48-73).
template <typename T> __global__ void __launch_bounds__(256) fk2c_coef_step_v1(const StepArgs<T> a)
{
    using R = Rn<T>;
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ P = a.buf[sc.par][0];
    const T *__restrict__ N = a.buf[sc.par][1];
    const T *__restrict__ S = a.buf[sc.par][2];
    int i, j, k;
    Nbr n;
    double sP = 0.0, sN = 0.0;
    if (voxel_index(a, i, j, k, n)) {
        const T p = P[n.c], nn = N[n.c], s = S[n.c];
        auto term = [&](T inv, T dp, T dm, T um, T u, T up) {
            return R::mul(inv, R::sub(R::mul(dp, R::sub(um, u)), R::mul(dm, R::sub(u, up))));
        };
        const T H = R::sub(T(1), T(1) / R::add(T(1), tg_exp(R::mul(T(-50), R::sub(s, a.sigma_np)))));
        const T sp = R::add(R::add(term(a.cx, a.coef[0][n.c], a.coef[3][n.c], P[n.xm], p, P[n.xp]),
                                   term(a.cy, a.coef[1][n.c], a.coef[4][n.c], P[n.ym], p, P[n.yp])),
                            term(a.cz, a.coef[2][n.c], a.coef[5][n.c], P[n.zm], p, P[n.zp]));
        const T growth = R::mul(R::mul(a.rho, R::mul(s, p)), R::sub(R::sub(T(1), p), nn));
        const T death = R::mul(R::mul(a.lambda_np, p), H);
        const T pn = R::add(p, R::mul(R::sub(R::add(sp, growth), death), a.dt));
        const T nnew = R::add(nn, R::mul(R::mul(R::mul(a.lambda_np, pn), H), a.dt));
        const T ss = R::add(R::add(term(a.cx, a.coef[6][n.c], a.coef[9][n.c], S[n.xm], s, S[n.xp]),
                                   term(a.cy, a.coef[7][n.c], a.coef[10][n.c], S[n.ym], s, S[n.yp])),
                            term(a.cz, a.coef[8][n.c], a.coef[11][n.c], S[n.zm], s, S[n.zp]));
        const T snew = R::add(s, R::mul(R::sub(ss, R::mul(R::mul(a.lambda_s, s), pn)), a.dt));
        a.buf[sc.par ^ 1][0][n.c] = pn;
        a.buf[sc.par ^ 1][1][n.c] = nnew;
        a.buf[sc.par ^ 1][2][n.c] = snew;
        if (counts(a, i)) {
            sP = (double)pn;
            sN = (double)nnew;
        }
    }
    step_end<T, 2>(a, sc, sP, sN);
}

}  // namespace tgtk
