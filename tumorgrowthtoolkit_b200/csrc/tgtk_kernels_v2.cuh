// Tuned step kernels: 2.5-D marching.  A block owns a (TY x TZ) tile of the y-z plane and a
// segment of LX consecutive x planes; each thread walks its column along x (the slowest
// axis) keeping the x-1 / x / x+1 values of every field in registers, so each field is read
// from L2/HBM once per voxel and only the four in-plane neighbours come through L1.
// Periodic wrap (np.roll semantics) is handled by index arithmetic, also across segments.
//
// Element offsets are 32-bit and advanced incrementally (a field has < 2^31 elements; checked
// on the host): one IMAD.WIDE per load instead of 64-bit index arithmetic, which otherwise
// dominates the instruction stream of a 25-load stencil.
#pragma once
#include "tgtk_common.cuh"

namespace tgtk {

struct March {
    int x0, x1;                     // planes [x0, x1) of this block
    bool valid;                     // thread maps to a voxel column inside the box
    uint32_t c, ym, yp, zm, zp;     // element offsets of the current plane's five stencil points
    uint32_t prev;                  // own column in plane x0-1 (wrapped)
    uint32_t sx, wrap_back;         // plane pitch; (X-1)*sx
};

template <typename T> __device__ __forceinline__ March march_setup(const StepArgs<T> &a, int lx)
{
    March m;
    int z = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    m.valid = (z < a.Z) && (y < a.Y);
    if (z >= a.Z) z = a.Z - 1;      // out-of-box threads shadow the last column: loads stay legal,
    if (y >= a.Y) y = a.Y - 1;      // nothing is stored, nothing is summed
    m.x0 = blockIdx.z * lx;
    m.x1 = min(m.x0 + lx, a.X);
    const int yl = (y == 0) ? a.Y - 1 : y - 1, yh = (y + 1 == a.Y) ? 0 : y + 1;
    const int zl = (z == 0) ? a.Z - 1 : z - 1, zh = (z + 1 == a.Z) ? 0 : z + 1;
    const uint32_t sy = (uint32_t)a.sy;
    m.sx = (uint32_t)a.sx;
    m.wrap_back = (uint32_t)(a.X - 1) * m.sx;
    const uint32_t base = (uint32_t)m.x0 * m.sx;
    m.c = base + (uint32_t)y * sy + z;
    m.ym = base + (uint32_t)yl * sy + z;
    m.yp = base + (uint32_t)yh * sy + z;
    m.zm = base + (uint32_t)y * sy + zl;
    m.zp = base + (uint32_t)y * sy + zh;
    m.prev = (m.x0 == 0) ? m.c + m.wrap_back : m.c - m.sx;
    return m;
}

// offset of the own column one plane up (wrapping to plane 0 after plane X-1)
template <typename T> __device__ __forceinline__ uint32_t march_next(const StepArgs<T> &a, const March &m, int x)
{
    return (x + 1 == a.X) ? m.c - m.wrap_back : m.c + m.sx;
}

__device__ __forceinline__ void march_advance(March &m, uint32_t next)
{
    const uint32_t d = next - m.c;   // +sx, or -(X-1)*sx modulo 2^32 at the wrap
    m.c = next; m.ym += d; m.yp += d; m.zm += d; m.zp += d;
}

// 1/x for x in [1, 1e300]: hardware seed + two Newton steps (no slow path, <= 1 ulp).
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) { return __frcp_rn(x); }

// smooth_heaviside (FK_2c.py:75-76) = 1 - 1/(1+e), e = exp(-50 (s - sigma)).  The argument is
// clamped to +-600 so 1+e stays finite and the reciprocal needs no special cases; for
// e > 1e260 the reference gives 1 - 1/inf = 1 and so does this.
template <typename T> __device__ __forceinline__ T heaviside50_fast(T s, T sigma)
{
    T arg = T(-50) * (s - sigma);
    arg = (arg > T(600)) ? T(600) : arg;
    if (sizeof(T) == 4) arg = (arg > T(80)) ? T(80) : arg;
    return T(1) - fast_rcp(T(1) + tg_exp(arg));
}

// ---- FK (compact coefficient k, see fk_step_v1) -------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) fk_step_v2(const StepArgs<T> a, const int lx)
{
    const int t = a.ctrl->t;
    if (a.ctrl->stop_reason != 0) return;
    const T *__restrict__ u = a.buf[t & 1][0];
    T *__restrict__ o = a.buf[(t + 1) & 1][0];
    const T *__restrict__ kc = a.coef[0];
    March m = march_setup(a, lx);
    double s = 0.0;
    if (m.x0 < m.x1) {
        T uc = u[m.c], kk = kc[m.c];
        // flux through the face below (x-1 | x); the face above is handed on to the next plane
        T flux_lo = relu(kc[m.prev] + kk) * (u[m.prev] - uc);
#pragma unroll 2
        for (int x = m.x0; x < m.x1; ++x) {
            const uint32_t nx = march_next(a, m, x);
            const T up = u[nx], kp = kc[nx];
            const T uym = u[m.ym], uyp = u[m.yp], uzm = u[m.zm], uzp = u[m.zp];
            const T kym = kc[m.ym], kyp = kc[m.yp], kzm = kc[m.zm], kzp = kc[m.zp];
            const T flux_hi = relu(kk + kp) * (uc - up);
            T sp = a.cx * (flux_lo - flux_hi);
            sp += a.cy * (relu(kym + kk) * (uym - uc) - relu(kk + kyp) * (uc - uyp));
            sp += a.cz * (relu(kzm + kk) * (uzm - uc) - relu(kk + kzp) * (uc - uzp));
            const T un = uc + a.dt * (sp + a.rho * uc * (T(1) - uc));
            if (m.valid) {
                o[m.c] = un;
                s += (double)un;
            }
            flux_lo = flux_hi; uc = up; kk = kp;
            march_advance(m, nx);
        }
    }
    finish_step<T, 1>(a, t, s, 0.0);
}

// ---- FK_DTI: coef[0..2] = D_{x,y,z}(c), reference rounding order (see coef_step_v1 mode 2) -----
template <typename T> __global__ void __launch_bounds__(256) dti_step_v2(const StepArgs<T> a, const int lx)
{
    using R = Rn<T>;
    const int t = a.ctrl->t;
    if (a.ctrl->stop_reason != 0) return;
    const T *__restrict__ u = a.buf[t & 1][0];
    T *__restrict__ o = a.buf[(t + 1) & 1][0];
    const T *__restrict__ dxa = a.coef[0];
    const T *__restrict__ dya = a.coef[1];
    const T *__restrict__ dza = a.coef[2];
    March m = march_setup(a, lx);
    double s = 0.0;
    if (m.x0 < m.x1) {
        T um = u[m.prev], uc = u[m.c];
        for (int x = m.x0; x < m.x1; ++x) {
            const uint32_t nx = march_next(a, m, x);
            const T up = u[nx];
            const T d0 = dxa[m.c], d1 = dya[m.c], d2 = dza[m.c];
            const T uym = u[m.ym], uyp = u[m.yp], uzm = u[m.zm], uzp = u[m.zp];
            const T sx = R::mul(a.cx, R::sub(R::mul(d0, R::sub(um, uc)), R::mul(d0, R::sub(uc, up))));
            const T sy = R::mul(a.cy, R::sub(R::mul(d1, R::sub(uym, uc)), R::mul(d1, R::sub(uc, uyp))));
            const T sz = R::mul(a.cz, R::sub(R::mul(d2, R::sub(uzm, uc)), R::mul(d2, R::sub(uc, uzp))));
            const T sp = R::add(R::add(sx, sy), sz);
            const T un = R::add(uc, R::mul(R::add(sp, R::mul(a.rho, R::mul(uc, R::sub(T(1), uc)))), a.dt));
            if (m.valid) {
                o[m.c] = un;
                s += (double)un;
            }
            um = uc; uc = up;
            march_advance(m, nx);
        }
    }
    finish_step<T, 1>(a, t, s, 0.0);
}

// ---- FK_2c (compact coefficients k, b; see fk2c_step_v1) ----------------------------------------
template <typename T> __global__ void __launch_bounds__(256) fk2c_step_v2(const StepArgs<T> a, const int lx)
{
    const int t = a.ctrl->t;
    if (a.ctrl->stop_reason != 0) return;
    const T *__restrict__ P = a.buf[t & 1][0];
    const T *__restrict__ N = a.buf[t & 1][1];
    const T *__restrict__ S = a.buf[t & 1][2];
    T *__restrict__ Po = a.buf[(t + 1) & 1][0];
    T *__restrict__ No = a.buf[(t + 1) & 1][1];
    T *__restrict__ So = a.buf[(t + 1) & 1][2];
    const T *__restrict__ kc = a.coef[0];
    const T *__restrict__ bc = a.coef[1];
    March m = march_setup(a, lx);
    const T closed = -Big<T>::v;
    double sP = 0.0, sN = 0.0;
    // k with the necrotic closure of FK_2c.py:13-15 applied: -Big where P+N > th_necro
    auto keff = [&](uint32_t q, T pq) { return (pq + N[q] <= a.th_necro) ? kc[q] : closed; };
    if (m.x0 < m.x1) {
        T p_c = P[m.c], n_c = N[m.c], s_c = S[m.c], b_c = bc[m.c];
        T k_c = (p_c + n_c <= a.th_necro) ? kc[m.c] : closed;
        // fluxes through the face below (x-1 | x); the face above is handed on to the next plane
        T fp_lo, fs_lo;
        {
            const T p_m = P[m.prev];
            fp_lo = relu(keff(m.prev, p_m) + k_c) * (p_m - p_c);
            fs_lo = relu(bc[m.prev] + b_c) * (S[m.prev] - s_c);
        }
#pragma unroll 2
        for (int x = m.x0; x < m.x1; ++x) {
            const uint32_t nx = march_next(a, m, x);
            const T p_p = P[nx], n_p = N[nx], s_p = S[nx], b_p = bc[nx];
            const T k_p = (p_p + n_p <= a.th_necro) ? kc[nx] : closed;
            const T pym = P[m.ym], pyp = P[m.yp], pzm = P[m.zm], pzp = P[m.zp];
            const T kym = keff(m.ym, pym), kyp = keff(m.yp, pyp);
            const T kzm = keff(m.zm, pzm), kzp = keff(m.zp, pzp);
            const T fp_hi = relu(k_c + k_p) * (p_c - p_p);
            const T fs_hi = relu(b_c + b_p) * (s_c - s_p);
            T sp = a.cx * (fp_lo - fp_hi);
            sp += a.cy * (relu(kym + k_c) * (pym - p_c) - relu(k_c + kyp) * (p_c - pyp));
            sp += a.cz * (relu(kzm + k_c) * (pzm - p_c) - relu(k_c + kzp) * (p_c - pzp));
            T ss = a.ex * (fs_lo - fs_hi);
            ss += a.ey * (relu(bc[m.ym] + b_c) * (S[m.ym] - s_c) - relu(b_c + bc[m.yp]) * (s_c - S[m.yp]));
            ss += a.ez * (relu(bc[m.zm] + b_c) * (S[m.zm] - s_c) - relu(b_c + bc[m.zp]) * (s_c - S[m.zp]));
            const T H = heaviside50_fast(s_c, a.sigma_np);
            const T pn = p_c + a.dt * (sp + a.rho * (s_c * p_c) * ((T(1) - p_c) - n_c) - (a.lambda_np * p_c) * H);
            const T nn = n_c + a.dt * ((a.lambda_np * pn) * H);
            const T sn = s_c + a.dt * (ss - (a.lambda_s * s_c) * pn);
            if (m.valid) {
                Po[m.c] = pn;
                No[m.c] = nn;
                So[m.c] = sn;
                sP += (double)pn;
                sN += (double)nn;
            }
            fp_lo = fp_hi; fs_lo = fs_hi;
            p_c = p_p; n_c = n_p; s_c = s_p; b_c = b_p; k_c = k_p;
            march_advance(m, nx);
        }
    }
    finish_step<T, 2>(a, t, sP, sN);
}

}  // namespace tgtk
