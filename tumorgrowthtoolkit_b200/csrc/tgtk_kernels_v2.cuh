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

// offset of a column one plane above plane x (wrapping to plane 0 after plane X-1)
template <typename T>
__device__ __forceinline__ uint32_t plane_up(const StepArgs<T> &a, const March &m, uint32_t off, int x)
{
    return (x + 1 == a.X) ? off - m.wrap_back : off + m.sx;
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
// exp(x) for x in [-700, 600] (callers clamp): Cody-Waite reduction x = n ln2 + r, |r| <= ln2/2,
// degree-12 Taylor polynomial (truncation 1.7e-16 relative), 2^n applied to the exponent field.
// No range checks and the coefficients come from the constant bank as DFMA operands (libdevice's
// exp() re-materialises 11 coefficients in uniform registers every call: 22 UMOV per voxel).
__constant__ double kExpTaylor[11] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                      1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,     1.0 / 120.0,
                                      1.0 / 24.0,        1.0 / 6.0,        0.5};

__device__ __forceinline__ double fast_exp(double x)
{
    const double magic = 6755399441055744.0;                      // 1.5 * 2^52: rounds to nearest integer
    const double t = fma(x, 1.4426950408889634, magic);
    const int n = __double2loint(t);
    const double nf = t - magic;
    double r = fma(nf, -6.93147180369123816490e-01, x);           // ln2 high part
    r = fma(nf, -1.90821492927058770002e-10, r);                  // ln2 low part
#ifndef TGTK_EXP_ESTRIN
    double p = kExpTaylor[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = fma(p, r, kExpTaylor[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
#else
    // Estrin's scheme: the same degree-12 polynomial, dependency depth 4 instead of 13 (the kernel that calls this
    // is bound by fp64 latency chains, not by the two extra multiplications); c_i = 1/i! = kExpTaylor[12 - i]
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = fma(r, 1.0, 1.0), a1 = fma(kExpTaylor[9], r, kExpTaylor[10]), a2 = fma(kExpTaylor[7], r, kExpTaylor[8]);
    const double a3 = fma(kExpTaylor[5], r, kExpTaylor[6]), a4 = fma(kExpTaylor[3], r, kExpTaylor[4]);
    const double a5 = fma(kExpTaylor[1], r, kExpTaylor[2]);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double d0 = fma(b1, r4, b0), d1 = fma(kExpTaylor[0], r4, b2);
    const double p = fma(d1, r8, d0);
#endif
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
__device__ __forceinline__ float fast_exp(float x) { return expf(x); }

template <typename T> __device__ __forceinline__ T heaviside50_fast(T s, T sigma)
{
    T arg = T(-50) * (s - sigma);
    arg = (arg > T(600)) ? T(600) : arg;
    arg = (arg < T(-700)) ? T(-700) : arg;                        // exp < 1e-304: 1 + e == 1 either way
    if (sizeof(T) == 4) arg = (arg > T(80)) ? T(80) : arg;
    return T(1) - fast_rcp(T(1) + fast_exp(arg));
}

// fp32: 1 - 1/(1 + 2^t), t = -50 log2(e) (s - sigma) capped so that 2^t stays finite, with the SFU's ex2 and
// rcp (relative error 2^-22, far inside the fp32 mode's 1e-4 contract): 7 instructions instead of expf's range
// handling plus the correctly rounded reciprocal's slow-path call.
__device__ __forceinline__ float heaviside50_fast(float s, float sigma)
{
    const float c = 72.13475204444817f;   // 50 log2(e)
    const float t = fminf(fmaf(s, -c, c * sigma), 115.0f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return 1.0f - r;
}

// All three kernels are software-pipelined along x: the own-column values of plane x+2 are
// requested while plane x is computed (they are the loads that miss L1 and come from L2/HBM);
// the in-plane neighbours of plane x are L1 hits issued at the top of the iteration.  The loop
// is unrolled three times with the plane roles (current / next / prefetched) rotating through
// three register sets, so no register moves (which would wait on the prefetch) are needed.

// ---- FK (compact coefficient k, see fk_step_v1) -------------------------------------------
template <typename T> struct ColFk { T u, k; };

template <typename T> __global__ void __launch_bounds__(256) fk_step_v2(const StepArgs<T> a, const int lx)
{
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ o = a.buf[sc.par ^ 1][0];
    const T *__restrict__ kc = a.coef[0];
    March m = march_setup(a, lx);
    double s = 0.0;
    if (m.x0 < m.x1) {
        ColFk<T> A, B, C;
        A.u = u[m.c]; A.k = kc[m.c];
        // flux through the face below (x-1 | x); the face above is handed on to the next plane
        T flux_lo = relu(kc[m.prev] + A.k) * (u[m.prev] - A.u);
        uint32_t n1 = plane_up(a, m, m.c, m.x0);
        B.u = u[n1]; B.k = kc[n1];
        auto step = [&](int x, const ColFk<T> &c, const ColFk<T> &p, ColFk<T> &q) {
            const int x1 = (x + 1 == a.X) ? 0 : x + 1;
            const uint32_t n2 = plane_up(a, m, n1, x1);
            q.u = u[n2]; q.k = kc[n2];                              // prefetch plane x+2
            const T uym = u[m.ym], uyp = u[m.yp], uzm = u[m.zm], uzp = u[m.zp];
            const T kym = kc[m.ym], kyp = kc[m.yp], kzm = kc[m.zm], kzp = kc[m.zp];
            const T flux_hi = relu(c.k + p.k) * (c.u - p.u);
            T sp = a.cx * (flux_lo - flux_hi);
            sp += a.cy * (relu(kym + c.k) * (uym - c.u) - relu(c.k + kyp) * (c.u - uyp));
            sp += a.cz * (relu(kzm + c.k) * (uzm - c.u) - relu(c.k + kzp) * (c.u - uzp));
            const T un = c.u + a.dt * (sp + a.rho * c.u * (T(1) - c.u));
            if (m.valid) {
                o[m.c] = un;
                if (counts(a, x)) s += (double)un;
            }
            flux_lo = flux_hi;
            march_advance(m, n1);
            n1 = n2;
        };
        for (int x = m.x0; x < m.x1; x += 3) {
            step(x, A, B, C);
            if (x + 1 >= m.x1) break;
            step(x + 1, B, C, A);
            if (x + 2 >= m.x1) break;
            step(x + 2, C, A, B);
        }
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK_DTI: coef[0..2] = D_{x,y,z}(c), reference rounding order (see coef_step_v1 mode 2) -----
template <typename T> struct ColDti { T u, d0, d1, d2; };

template <typename T> __global__ void __launch_bounds__(256) dti_step_v2(const StepArgs<T> a, const int lx)
{
    using R = Rn<T>;
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ o = a.buf[sc.par ^ 1][0];
    const T *__restrict__ dxa = a.coef[0];
    const T *__restrict__ dya = a.coef[1];
    const T *__restrict__ dza = a.coef[2];
    March m = march_setup(a, lx);
    double s = 0.0;
    if (m.x0 < m.x1) {
        ColDti<T> A, B, C;
        T um = u[m.prev];
        A.u = u[m.c]; A.d0 = dxa[m.c]; A.d1 = dya[m.c]; A.d2 = dza[m.c];
        uint32_t n1 = plane_up(a, m, m.c, m.x0);
        B.u = u[n1]; B.d0 = dxa[n1]; B.d1 = dya[n1]; B.d2 = dza[n1];
        auto step = [&](int x, const ColDti<T> &c, const ColDti<T> &p, ColDti<T> &q) {
            const int x1 = (x + 1 == a.X) ? 0 : x + 1;
            const uint32_t n2 = plane_up(a, m, n1, x1);
            q.u = u[n2]; q.d0 = dxa[n2]; q.d1 = dya[n2]; q.d2 = dza[n2];      // prefetch plane x+2
            const T uym = u[m.ym], uyp = u[m.yp], uzm = u[m.zm], uzp = u[m.zp];
            const T sx = R::mul(a.cx, R::sub(R::mul(c.d0, R::sub(um, c.u)), R::mul(c.d0, R::sub(c.u, p.u))));
            const T sy = R::mul(a.cy, R::sub(R::mul(c.d1, R::sub(uym, c.u)), R::mul(c.d1, R::sub(c.u, uyp))));
            const T sz = R::mul(a.cz, R::sub(R::mul(c.d2, R::sub(uzm, c.u)), R::mul(c.d2, R::sub(c.u, uzp))));
            const T sp = R::add(R::add(sx, sy), sz);
            const T un = R::add(c.u, R::mul(R::add(sp, R::mul(a.rho, R::mul(c.u, R::sub(T(1), c.u)))), a.dt));
            if (m.valid) {
                o[m.c] = un;
                if (counts(a, x)) s += (double)un;
            }
            um = c.u;
            march_advance(m, n1);
            n1 = n2;
        };
        for (int x = m.x0; x < m.x1; x += 3) {
            step(x, A, B, C);
            if (x + 1 >= m.x1) break;
            step(x + 1, B, C, A);
            if (x + 2 >= m.x1) break;
            step(x + 2, C, A, B);
        }
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK_2c (compact coefficients k, b; see fk2c_step_v1) ----------------------------------------
template <typename T> struct Col2c { T p, n, s, b, k; };   // k: with the necrotic closure applied

template <typename T> __global__ void __launch_bounds__(256) fk2c_step_v2(const StepArgs<T> a, const int lx)
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
    March m = march_setup(a, lx);
    const T closed = -Big<T>::v;
    double sP = 0.0, sN = 0.0;
    // k with the necrotic closure of FK_2c.py:13-15 applied: -Big where P+N > th_necro.  All
    // three loads are issued unconditionally so none of them waits for another.
    auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
    if (m.x0 < m.x1) {
        Col2c<T> A, B, C;
        A.p = P[m.c]; A.n = N[m.c]; A.s = S[m.c]; A.b = bc[m.c]; A.k = keff(A.p, A.n, kc[m.c]);
        // fluxes through the face below (x-1 | x); the face above is handed on to the next plane
        T fp_lo, fs_lo;
        {
            const T p_m = P[m.prev];
            fp_lo = relu(keff(p_m, N[m.prev], kc[m.prev]) + A.k) * (p_m - A.p);
            fs_lo = relu(bc[m.prev] + A.b) * (S[m.prev] - A.s);
        }
        uint32_t n1 = plane_up(a, m, m.c, m.x0);
        B.p = P[n1]; B.n = N[n1]; B.s = S[n1]; B.b = bc[n1]; B.k = kc[n1];   // B.k raw until its turn
        auto step = [&](int x, const Col2c<T> &c, Col2c<T> &p, Col2c<T> &q) {
            const int x1 = (x + 1 == a.X) ? 0 : x + 1;
            const uint32_t n2 = plane_up(a, m, n1, x1);
            q.p = P[n2]; q.n = N[n2]; q.s = S[n2]; q.b = bc[n2]; q.k = kc[n2];          // prefetch x+2
            const T pym = P[m.ym], pyp = P[m.yp], pzm = P[m.zm], pzp = P[m.zp];
            const T nym = N[m.ym], nyp = N[m.yp], nzm = N[m.zm], nzp = N[m.zp];
            const T rym = kc[m.ym], ryp = kc[m.yp], rzm = kc[m.zm], rzp = kc[m.zp];
            const T sym = S[m.ym], syp = S[m.yp], szm = S[m.zm], szp = S[m.zp];
            const T bym = bc[m.ym], byp = bc[m.yp], bzm = bc[m.zm], bzp = bc[m.zp];
            // work that needs no in-plane neighbour first: Heaviside, reaction, x faces
            const T H = heaviside50_fast(c.s, a.sigma_np);
            p.k = keff(p.p, p.n, p.k);
            const T fp_hi = relu(c.k + p.k) * (c.p - p.p);
            const T fs_hi = relu(c.b + p.b) * (c.s - p.s);
            const T react = a.rho * (c.s * c.p) * ((T(1) - c.p) - c.n) - (a.lambda_np * c.p) * H;
            T sp = a.cx * (fp_lo - fp_hi);
            T ss = a.ex * (fs_lo - fs_hi);
            ss += a.ey * (relu(bym + c.b) * (sym - c.s) - relu(c.b + byp) * (c.s - syp));
            ss += a.ez * (relu(bzm + c.b) * (szm - c.s) - relu(c.b + bzp) * (c.s - szp));
            sp += a.cy * (relu(keff(pym, nym, rym) + c.k) * (pym - c.p) - relu(c.k + keff(pyp, nyp, ryp)) * (c.p - pyp));
            sp += a.cz * (relu(keff(pzm, nzm, rzm) + c.k) * (pzm - c.p) - relu(c.k + keff(pzp, nzp, rzp)) * (c.p - pzp));
            const T pn = c.p + a.dt * (sp + react);
            const T nn = c.n + a.dt * ((a.lambda_np * pn) * H);
            const T sn = c.s + a.dt * (ss - (a.lambda_s * c.s) * pn);
            if (m.valid) {
                Po[m.c] = pn;
                No[m.c] = nn;
                So[m.c] = sn;
                if (counts(a, x)) {
                    sP += (double)pn;
                    sN += (double)nn;
                }
            }
            fp_lo = fp_hi; fs_lo = fs_hi;
            march_advance(m, n1);
            n1 = n2;
        };
        for (int x = m.x0; x < m.x1; x += 3) {
            step(x, A, B, C);
            if (x + 1 >= m.x1) break;
            step(x + 1, B, C, A);
            if (x + 2 >= m.x1) break;
            step(x + 2, C, A, B);
        }
    }
    step_end<T, 2>(a, sc, sP, sN);
}

}  // namespace tgtk
