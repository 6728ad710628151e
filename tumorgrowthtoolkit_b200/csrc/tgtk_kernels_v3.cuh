// Shared-memory staged marching kernels (the default).
//
// Same decomposition as tgtk_kernels_v2.cuh -- a block owns a (TY x TZ) column tile and walks
// LX planes along x, own-column values of planes x, x+1, x+2 live in three rotating register
// sets -- but the four in-plane neighbours of plane x are read from a shared-memory copy of the
// plane (tile + a one-voxel ring) instead of L1:
//   * every global load of the kernel is issued at least one full plane ahead of its first
//     use (own column: two planes; ring: one plane), so no L2/HBM latency is exposed;
//   * per-voxel derived quantities (FK_2c: the diffusivity with the necrotic closure applied)
//     are computed once by the owning thread and shared, not recomputed by four neighbours.
// The plane copy is double buffered: one __syncthreads() per plane.
//
// Out-of-box threads of ragged edge tiles and the ring loaders use periodically wrapped
// coordinates, so the slot next to the last in-box voxel always holds its np.roll neighbour.
#pragma once
#include "tgtk_common.cuh"
#include "tgtk_kernels_v2.cuh"   // fast_rcp, heaviside50_fast

namespace tgtk {

template <typename T, int TZ> struct TileGeom {
    // row stride in elements; for 16-wide fp32 rows the two half-warps are put 16 banks apart
    static constexpr int RS = (sizeof(T) == 4 && TZ == 16) ? 48 : TZ + 2;
};

// two fields of one voxel side by side: one 128-bit (fp64) / 64-bit (fp32) shared-memory access
// per neighbour instead of two
template <typename T> struct alignas(2 * sizeof(T)) Pair {
    T a, b;
};

struct Stage {
    int x0, x1;
    bool valid;          // own column is inside the box (else: a wrapped shadow, nothing stored)
    bool ring;           // this thread also loads one element of the ring
    uint32_t c;          // element offset of the own column in plane x
    uint32_t prev;       // own column in plane x0-1
    uint32_t h;          // ring element offset in the plane being prefetched
    uint32_t sx, wrap_back;
    int so, hs;          // shared-memory slot of the own column / of the ring element
};

__device__ __forceinline__ int wrap_coord(int v, int n)
{
    v %= n;
    return (v < 0) ? v + n : v;
}

template <typename T, int TZ, int TY, int RS = TileGeom<T, TZ>::RS>
__device__ __forceinline__ Stage stage_setup(const StepArgs<T> &a, int lx)
{
    Stage g;
    const int tz = threadIdx.x, ty = threadIdx.y, tid = ty * TZ + tz;
    const int z0 = blockIdx.x * TZ, y0 = blockIdx.y * TY;
    g.valid = (z0 + tz < a.Z) && (y0 + ty < a.Y);
    g.x0 = blockIdx.z * lx;
    g.x1 = min(g.x0 + lx, a.X);
    g.sx = (uint32_t)a.sx;
    g.wrap_back = (uint32_t)(a.X - 1) * g.sx;
    const uint32_t sy = (uint32_t)a.sy, base = (uint32_t)g.x0 * g.sx;
    g.c = base + (uint32_t)wrap_coord(y0 + ty, a.Y) * sy + (uint32_t)wrap_coord(z0 + tz, a.Z);
    g.prev = (g.x0 == 0) ? g.c + g.wrap_back : g.c - g.sx;
    g.so = (ty + 1) * RS + tz + 1;
    // ring: TZ elements above, TZ below, TY left, TY right
    int hr = 0, hc = 0;
    g.ring = tid < 2 * TZ + 2 * TY;
    if (tid < TZ) { hr = 0; hc = tid + 1; }
    else if (tid < 2 * TZ) { hr = TY + 1; hc = tid - TZ + 1; }
    else if (tid < 2 * TZ + TY) { hr = tid - 2 * TZ + 1; hc = 0; }
    else { hr = tid - 2 * TZ - TY + 1; hc = TZ + 1; }
    if (!g.ring) { hr = 1; hc = 1; }
    g.hs = hr * RS + hc;
    g.h = base + (uint32_t)wrap_coord(y0 - 1 + hr, a.Y) * sy + (uint32_t)wrap_coord(z0 - 1 + hc, a.Z);
    return g;
}

template <typename T>
__device__ __forceinline__ uint32_t stage_up(const StepArgs<T> &a, const Stage &g, uint32_t off, int x)
{
    return (x + 1 == a.X) ? off - g.wrap_back : off + g.sx;
}

// ---- FK_2c ---------------------------------------------------------------------------------------
template <typename T> struct Col2cS { T p, n, s, b, k; };

template <typename T, int TZ, int TY, int MINB = 2>
__global__ void __launch_bounds__(TZ *TY, MINB) fk2c_step_v3(const StepArgs<T> a, const int lx)
{
    constexpr int RS = TZ + 2, PL = (TY + 2) * RS;
    __shared__ Pair<T> tile[2][2][PL];     // [buffer][(P, k_eff), (S, b)][slot]
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
    Stage g = stage_setup<T, TZ, TY, RS>(a, lx);
    const T closed = -Big<T>::v;
    // diffusivity half-coefficient with the necrotic closure of FK_2c.py:13-15 applied
    auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
    double sP = 0.0, sN = 0.0;

    Col2cS<T> A, B, C;
    A.p = P[g.c]; A.n = N[g.c]; A.s = S[g.c]; A.b = bc[g.c]; A.k = keff(A.p, A.n, kc[g.c]);
    T fp_lo, fs_lo;     // fluxes through the face below (x-1 | x); the upper face is handed on
    {
        const T p_m = P[g.prev];
        fp_lo = relu(keff(p_m, N[g.prev], kc[g.prev]) + A.k) * (p_m - A.p);
        fs_lo = relu(bc[g.prev] + A.b) * (S[g.prev] - A.s);
    }
    uint32_t n1 = stage_up(a, g, g.c, g.x0);
    B.p = P[n1]; B.n = N[n1]; B.s = S[n1]; B.b = bc[n1]; B.k = kc[n1];     // B.k raw until its turn
    // ring values of the plane about to be published (one register set, refilled right after
    // it has been stored)
    T hp = 0, hn = 0, hk = 0, hsv = 0, hb = 0;
    if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
    int buf = 0;

    auto step = [&](int x, const Col2cS<T> &c, Col2cS<T> &p, Col2cS<T> &q) {
        Pair<T>(*tl)[PL] = tile[buf];
        // publish plane x
        tl[0][g.so] = Pair<T>{c.p, c.k};
        tl[1][g.so] = Pair<T>{c.s, c.b};
        if (g.ring) {
            tl[0][g.hs] = Pair<T>{hp, keff(hp, hn, hk)};
            tl[1][g.hs] = Pair<T>{hsv, hb};
        }
        // request plane x+2 (own column) and the ring of plane x+1
        const int x1 = (x + 1 == a.X) ? 0 : x + 1;
        const uint32_t n2 = stage_up(a, g, n1, x1);
        q.p = P[n2]; q.n = N[n2]; q.s = S[n2]; q.b = bc[n2]; q.k = kc[n2];
        g.h = stage_up(a, g, g.h, x);
        if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
        // work that needs no neighbour: Heaviside, reaction, x faces
        const T H = heaviside50_fast(c.s, a.sigma_np);
        p.k = keff(p.p, p.n, p.k);
        const T fp_hi = relu(c.k + p.k) * (c.p - p.p);
        const T fs_hi = relu(c.b + p.b) * (c.s - p.s);
        const T react = a.rho * (c.s * c.p) * ((T(1) - c.p) - c.n) - (a.lambda_np * c.p) * H;
        T sp = a.cx * (fp_lo - fp_hi);
        T ss = a.ex * (fs_lo - fs_hi);
        __syncthreads();
        const int o = g.so;
        const Pair<T> pym = tl[0][o - RS], pyp = tl[0][o + RS], pzm = tl[0][o - 1], pzp = tl[0][o + 1];   // (P, k)
        const Pair<T> sym = tl[1][o - RS], syp = tl[1][o + RS], szm = tl[1][o - 1], szp = tl[1][o + 1];   // (S, b)
        sp += a.cy * (relu(pym.b + c.k) * (pym.a - c.p) - relu(c.k + pyp.b) * (c.p - pyp.a));
        sp += a.cz * (relu(pzm.b + c.k) * (pzm.a - c.p) - relu(c.k + pzp.b) * (c.p - pzp.a));
        ss += a.ey * (relu(sym.b + c.b) * (sym.a - c.s) - relu(c.b + syp.b) * (c.s - syp.a));
        ss += a.ez * (relu(szm.b + c.b) * (szm.a - c.s) - relu(c.b + szp.b) * (c.s - szp.a));
        const T pn = c.p + a.dt * (sp + react);
        const T nn = c.n + a.dt * ((a.lambda_np * pn) * H);
        const T sn = c.s + a.dt * (ss - (a.lambda_s * c.s) * pn);
        if (g.valid) {
            Po[g.c] = pn;
            No[g.c] = nn;
            So[g.c] = sn;
            if (counts(a, x)) {
                sP += (double)pn;
                sN += (double)nn;
            }
        }
        fp_lo = fp_hi; fs_lo = fs_hi;
        g.c = n1; n1 = n2;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 3) {
        step(x, A, B, C);
        if (x + 1 >= g.x1) break;
        step(x + 1, B, C, A);
        if (x + 2 >= g.x1) break;
        step(x + 2, C, A, B);
    }
    step_end<T, 2>(a, sc, sP, sN);
}

// Same kernel with a prefetch distance of ONE plane (two register sets instead of three): the
// own-column loads of plane x+1 are issued at the top of step x and first used by the x-face flux at
// its end.  Ten registers fewer per thread -> four resident blocks per SM instead of three.
template <typename T, int TZ, int TY, int MINB = 4>
__global__ void __launch_bounds__(TZ *TY, MINB) fk2c_step_v3s(const StepArgs<T> a, const int lx)
{
    constexpr int RS = TZ + 2, PL = (TY + 2) * RS;
    __shared__ Pair<T> tile[2][2][PL];     // [buffer][(P, k_eff), (S, b)][slot]
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
    Stage g = stage_setup<T, TZ, TY, RS>(a, lx);
    const T closed = -Big<T>::v;
    // diffusivity half-coefficient with the necrotic closure of FK_2c.py:13-15 applied
    auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
    double sP = 0.0, sN = 0.0;

    Col2cS<T> A, B;
    A.p = P[g.c]; A.n = N[g.c]; A.s = S[g.c]; A.b = bc[g.c]; A.k = keff(A.p, A.n, kc[g.c]);
    T fp_lo, fs_lo;     // fluxes through the face below (x-1 | x); the upper face is handed on
    {
        const T p_m = P[g.prev];
        fp_lo = relu(keff(p_m, N[g.prev], kc[g.prev]) + A.k) * (p_m - A.p);
        fs_lo = relu(bc[g.prev] + A.b) * (S[g.prev] - A.s);
    }
    // ring values of the plane about to be published (one register set, refilled right after
    // it has been stored)
    T hp = 0, hn = 0, hk = 0, hsv = 0, hb = 0;
    if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
    int buf = 0;

    auto step = [&](int x, const Col2cS<T> &c, Col2cS<T> &p) {
        Pair<T>(*tl)[PL] = tile[buf];
        // publish plane x
        tl[0][g.so] = Pair<T>{c.p, c.k};
        tl[1][g.so] = Pair<T>{c.s, c.b};
        if (g.ring) {
            tl[0][g.hs] = Pair<T>{hp, keff(hp, hn, hk)};
            tl[1][g.hs] = Pair<T>{hsv, hb};
        }
        // request plane x+1: own column and ring
        const uint32_t n1 = stage_up(a, g, g.c, x);
        p.p = P[n1]; p.n = N[n1]; p.s = S[n1]; p.b = bc[n1]; p.k = kc[n1];
        g.h = stage_up(a, g, g.h, x);
        if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
        // work that needs neither neighbours nor plane x+1: Heaviside, reaction
        const T H = heaviside50_fast(c.s, a.sigma_np);
        const T react = a.rho * (c.s * c.p) * ((T(1) - c.p) - c.n) - (a.lambda_np * c.p) * H;
        T sp = T(0), ss = T(0);
        __syncthreads();
        const int o = g.so;
        const Pair<T> pym = tl[0][o - RS], pyp = tl[0][o + RS], pzm = tl[0][o - 1], pzp = tl[0][o + 1];   // (P, k)
        const Pair<T> sym = tl[1][o - RS], syp = tl[1][o + RS], szm = tl[1][o - 1], szp = tl[1][o + 1];   // (S, b)
        sp += a.cy * (relu(pym.b + c.k) * (pym.a - c.p) - relu(c.k + pyp.b) * (c.p - pyp.a));
        sp += a.cz * (relu(pzm.b + c.k) * (pzm.a - c.p) - relu(c.k + pzp.b) * (c.p - pzp.a));
        ss += a.ey * (relu(sym.b + c.b) * (sym.a - c.s) - relu(c.b + syp.b) * (c.s - syp.a));
        ss += a.ez * (relu(szm.b + c.b) * (szm.a - c.s) - relu(c.b + szp.b) * (c.s - szp.a));
        // x faces last: plane x+1 has had the whole step to arrive
        p.k = keff(p.p, p.n, p.k);
        const T fp_hi = relu(c.k + p.k) * (c.p - p.p);
        const T fs_hi = relu(c.b + p.b) * (c.s - p.s);
        sp += a.cx * (fp_lo - fp_hi);
        ss += a.ex * (fs_lo - fs_hi);
        const T pn = c.p + a.dt * (sp + react);
        const T nn = c.n + a.dt * ((a.lambda_np * pn) * H);
        const T sn = c.s + a.dt * (ss - (a.lambda_s * c.s) * pn);
        if (g.valid) {
            Po[g.c] = pn;
            No[g.c] = nn;
            So[g.c] = sn;
            if (counts(a, x)) {
                sP += (double)pn;
                sN += (double)nn;
            }
        }
        fp_lo = fp_hi; fs_lo = fs_hi;
        g.c = n1;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 2) {
        step(x, A, B);
        if (x + 1 >= g.x1) break;
        step(x + 1, B, A);
    }
    step_end<T, 2>(a, sc, sP, sN);
}

// ---- FK_2c, two voxels per thread -------------------------------------------------------------------
// Same staging as fk2c_step_v3s, but a thread owns two rows (y, y+1) of its column tile: the face
// between them is evaluated once and needs no shared-memory read, the tile is 32 x 16 voxels for the
// same 256 threads (ring 96 elements per 512 voxels instead of 80 per 256), and loop / address /
// barrier overhead is paid once per two voxels.  The two voxels give each warp two independent
// instruction streams, which replaces the occupancy the doubled register state costs.
struct Stage2 {
    int x0, x1;
    bool valid0, valid1, ring;
    uint32_t c0, c1, prev0, prev1, h;
    uint32_t sx, wrap_back;
    int so, hs;          // shared-memory slot of the upper-row voxel (the lower one is so + RS) / ring slot
};

// SLAB = 1: the instantiation the fused slab driver launches (restricted plane range, boundary segments last, halo push,
// optional serpentine sweep).  The ordinary instantiation carries none of it: its mere presence in the plane loop cost
// 5-6 % on FK / FK_2c (round-2 bisect, profiles/r2_sweep_misc.txt).
template <typename T, int TZ, int TYT, int RS, int SLAB = 0>
__device__ __forceinline__ Stage2 stage2_setup(const StepArgs<T> &a, int lx)
{
    constexpr int TYV = 2 * TYT;          // voxel rows per tile
    Stage2 g;
    const int tz = threadIdx.x, ty = threadIdx.y, tid = ty * TZ + tz;
    const int z0 = blockIdx.x * TZ, y0 = blockIdx.y * TYV;
    const int r0 = y0 + 2 * ty;
    g.valid0 = (z0 + tz < a.Z) && (r0 < a.Y);
    g.valid1 = (z0 + tz < a.Z) && (r0 + 1 < a.Y);
    // fused slab mode: the two segments that read a ghost plane get the highest block indices -- they are scheduled
    // last, when the neighbours' words have long arrived
    int seg = blockIdx.z;
    if (SLAB) {
        if (a.slab && gridDim.z > 2) seg = (seg + 2 < (int)gridDim.z) ? seg + 1 : (seg + 2 == (int)gridDim.z ? 0 : seg);
        g.x0 = a.x_lo + seg * lx;
        g.x1 = min(g.x0 + lx, a.x_hi);
    } else {
        g.x0 = seg * lx;
        g.x1 = min(g.x0 + lx, a.X);
    }
    g.sx = (uint32_t)a.sx;
    g.wrap_back = (uint32_t)(a.X - 1) * g.sx;
    const uint32_t sy = (uint32_t)a.sy, base = (uint32_t)g.x0 * g.sx;
    const uint32_t zc = (uint32_t)wrap_coord(z0 + tz, a.Z);
    g.c0 = base + (uint32_t)wrap_coord(r0, a.Y) * sy + zc;
    g.c1 = base + (uint32_t)wrap_coord(r0 + 1, a.Y) * sy + zc;
    g.prev0 = (g.x0 == 0) ? g.c0 + g.wrap_back : g.c0 - g.sx;
    g.prev1 = (g.x0 == 0) ? g.c1 + g.wrap_back : g.c1 - g.sx;
    g.so = (2 * ty + 1) * RS + tz + 1;
    int hr = 1, hc = 1;
    g.ring = tid < 2 * TZ + 2 * TYV;
    if (tid < TZ) { hr = 0; hc = tid + 1; }
    else if (tid < 2 * TZ) { hr = TYV + 1; hc = tid - TZ + 1; }
    else if (tid < 2 * TZ + TYV) { hr = tid - 2 * TZ + 1; hc = 0; }
    else if (g.ring) { hr = tid - 2 * TZ - TYV + 1; hc = TZ + 1; }
    g.hs = hr * RS + hc;
    g.h = base + (uint32_t)wrap_coord(y0 - 1 + hr, a.Y) * sy + (uint32_t)wrap_coord(z0 - 1 + hc, a.Z);
    return g;
}

// Marching direction of a block: up (x0 -> x1-1) or down.  The face fluxes are antisymmetric and the sums
// commutative, so a voxel's result does not depend on the direction (bitwise).
struct March2 {
    bool down;
    int xs, dx, n;            // first plane, +-1, number of planes
    uint32_t dsx;             // +-sx modulo 2^32
    int wrapx;                // plane index after which the offset wraps back (upward marching over the whole box)
    uint32_t wrapd;
};

template <typename T, int SLAB> __device__ __forceinline__ March2 march2_setup(const StepArgs<T> &a, Stage2 &g, int par)
{
    March2 m;
    // serpentine sweep (TGTK_ZIGZAG=1, measured +-1 %): odd steps march downwards, so a step starts on the planes the
    // previous step wrote last -- the part of the state that is still in L2 (the ping-pong parity is the step's parity)
    m.down = SLAB && a.zigzag && (par & 1);
    m.n = g.x1 - g.x0;
    m.xs = g.x0;
    m.dx = 1;
    m.dsx = g.sx;
    m.wrapx = a.X;            // upwards: the plane after X-1 is plane 0
    m.wrapd = 0u - g.wrap_back;
    if (m.down) {             // start at the segment's top plane; "previous" is the plane above it, the plane after 0 is X-1
        const uint32_t up = (uint32_t)(m.n - 1) * g.sx;
        g.c0 += up; g.c1 += up; g.h += up;
        const uint32_t behind = (g.x1 == a.X) ? 0u - g.wrap_back : g.sx;
        g.prev0 = g.c0 + behind; g.prev1 = g.c1 + behind;
        m.xs = g.x1 - 1;
        m.dx = -1;
        m.dsx = 0u - g.sx;
        m.wrapx = 1;          // (x + 1 == wrapx) <=> x == 0
        m.wrapd = g.wrap_back;
    }
    return m;
}

template <typename T> struct Vox2c { T p, n, s, b, k; };

// LL = 1 ("late loads"): the own-column values of the next plane are requested AFTER the barrier instead of at the top
// of the step -- their registers are then not live across the in-plane section, which is what lets the kernel fit
// 5 blocks per SM (96 registers) without spilling; the bulk L2 prefetch keeps the shortened distance cheap.
template <typename T, int TZ, int TYT, int MINB = 2, int LL = 0, int SLAB = 0>
__global__ void __launch_bounds__(TZ *TYT, MINB) fk2c_step_v4(const StepArgs<T> a, const int lx)
{
    constexpr int TYV = 2 * TYT, RS = TZ + 2, PL = (TYV + 2) * RS;
    __shared__ Pair<T> tile[2][2][PL];     // [buffer][(P, k_eff), (S, b)][slot]
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
    Stage2 g = stage2_setup<T, TZ, TYT, RS, SLAB>(a, lx);
    const March2 mm = march2_setup<T, SLAB>(a, g, sc.par);
    const T closed = -Big<T>::v;
    auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
    auto load = [&](Vox2c<T> &v, uint32_t off) { v.p = P[off]; v.n = N[off]; v.s = S[off]; v.b = bc[off]; v.k = kc[off]; };
    double sP = 0.0, sN = 0.0;
    griddep_launch();
    griddep_wait();
    const SlabCtl sl = SLAB ? slab_begin(a, g.x0, g.x1) : SlabCtl{false, false, 0};
    auto push = [&](T *const *peer, uint32_t off, T pn0, T nn0, T sn0, T pn1, T nn1, T sn1) {
        if (g.valid0) { peer[0][g.c0 + off] = pn0; peer[1][g.c0 + off] = nn0; peer[2][g.c0 + off] = sn0; }
        if (g.valid1) { peer[0][g.c1 + off] = pn1; peer[1][g.c1 + off] = nn1; peer[2][g.c1 + off] = sn1; }
    };

    Vox2c<T> A0, A1, B0, B1;               // rows (y, y+1) of plane x (A) and plane x+1 (B); roles swap
    load(A0, g.c0); A0.k = keff(A0.p, A0.n, A0.k);
    load(A1, g.c1); A1.k = keff(A1.p, A1.n, A1.k);
    T fp_lo0, fs_lo0, fp_lo1, fs_lo1;      // fluxes through the faces below (x-1 | x)
    {
        const T pm0 = P[g.prev0], pm1 = P[g.prev1];
        fp_lo0 = relu2(keff(pm0, N[g.prev0], kc[g.prev0]) + A0.k) * (pm0 - A0.p);
        fs_lo0 = relu2(bc[g.prev0] + A0.b) * (S[g.prev0] - A0.s);
        fp_lo1 = relu2(keff(pm1, N[g.prev1], kc[g.prev1]) + A1.k) * (pm1 - A1.p);
        fs_lo1 = relu2(bc[g.prev1] + A1.b) * (S[g.prev1] - A1.s);
    }
    T hp = 0, hn = 0, hk = 0, hsv = 0, hb = 0;
    if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto step = [&](int x, const Vox2c<T> &c0, const Vox2c<T> &c1, Vox2c<T> &p0, Vox2c<T> &p1) {
        Pair<T>(*tl)[PL] = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        tl[0][o0] = Pair<T>{c0.p, c0.k};
        tl[1][o0] = Pair<T>{c0.s, c0.b};
        tl[0][o1] = Pair<T>{c1.p, c1.k};
        tl[1][o1] = Pair<T>{c1.s, c1.b};
        if (g.ring) {
            tl[0][g.hs] = Pair<T>{hp, keff(hp, hn, hk)};
            tl[1][g.hs] = Pair<T>{hsv, hb};
        }
        // request the next plane of the march: both own voxels and the ring
        const uint32_t adv = (x + 1 == mm.wrapx) ? mm.wrapd : mm.dsx;
        const uint32_t n0 = g.c0 + adv, n1 = g.c1 + adv;
        if (!LL) {
            load(p0, n0);
            load(p1, n1);
        }
        g.h += adv;
        if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
        if (pf.on) {                           // this tile strip of plane x + l2pf into L2
            const uint32_t q = l2_strip_at(a, pf, x, mm.dx);
            l2_prefetch_strip(P, q, pf.count, pf.total); l2_prefetch_strip(N, q, pf.count, pf.total);
            l2_prefetch_strip(S, q, pf.count, pf.total); l2_prefetch_strip(kc, q, pf.count, pf.total);
            l2_prefetch_strip(bc, q, pf.count, pf.total);
        }
        // Two exact per-warp shortcuts (decided by warp votes, so nothing diverges; TGTK_CLOSED_SKIP=0 switches them off
        // at run time for A/B runs).  (1) The smooth Heaviside (28 fp64-pipe instructions per voxel: exp, reciprocal)
        // only ever multiplies P and the new P: a warp whose 64 voxels all hold P == 0 skips it -- 0 * H = 0 for the
        // finite H of a finite S.  (2) A voxel outside the tissue has b = k = "closed", so every one of its faces
        // carries the coefficient max(closed + k', 0) = 0 whatever the neighbour holds; with P == 0 on top, the
        // reference's update returns P' = 0, N' = N, S' = S bit for bit (0 * finite = +-0; x + +-0 = x).  A warp whose
        // 64 voxels are all of that kind -- 60 % of the cropped atlas box is background -- still loads, publishes its
        // tile rows for the neighbouring warps, takes the barrier and STORES its voxels (the HBM traffic is unchanged),
        // but skips the shared-memory reads and the arithmetic.
        const bool warp_has_p = !a.skip || __any_sync(0xffffffffu, (c0.p != T(0)) || (c1.p != T(0)));
        const bool warp_closed = !warp_has_p && __all_sync(0xffffffffu, (c0.b == closed) && (c1.b == closed));
        T pn0, pn1, nn0, nn1, sn0, sn1;
        if (warp_closed) {
            __syncthreads();
            if (LL) {
                load(p0, n0);
                load(p1, n1);
            }
            p0.k = keff(p0.p, p0.n, p0.k);
            p1.k = keff(p1.p, p1.n, p1.k);
            pn0 = c0.p; pn1 = c1.p; nn0 = c0.n; nn1 = c1.n; sn0 = c0.s; sn1 = c1.s;
            fp_lo0 = fs_lo0 = fp_lo1 = fs_lo1 = T(0);      // the faces towards the next plane are closed from this side
        } else {
        T H0 = T(0), H1 = T(0);
        if (warp_has_p) { H0 = heaviside50_fast(c0.s, a.sigma_np); H1 = heaviside50_fast(c1.s, a.sigma_np); }
        const T react0 = a.rho * (c0.s * c0.p) * ((T(1) - c0.p) - c0.n) - (a.lambda_np * c0.p) * H0;
        const T react1 = a.rho * (c1.s * c1.p) * ((T(1) - c1.p) - c1.n) - (a.lambda_np * c1.p) * H1;
        // the face between the two own rows: no shared memory involved
        const T fp_mid = relu2(c0.k + c1.k) * (c0.p - c1.p);
        const T fs_mid = relu2(c0.b + c1.b) * (c0.s - c1.s);
        __syncthreads();
        const Pair<T> pu = tl[0][o0 - RS], pd = tl[0][o1 + RS];                       // rows y-1 and y+2: (P, k)
        const Pair<T> su = tl[1][o0 - RS], sd = tl[1][o1 + RS];                       // (S, b)
        const Pair<T> pl0 = tl[0][o0 - 1], pr0 = tl[0][o0 + 1], pl1 = tl[0][o1 - 1], pr1 = tl[0][o1 + 1];
        const Pair<T> sl0 = tl[1][o0 - 1], sr0 = tl[1][o0 + 1], sl1 = tl[1][o1 - 1], sr1 = tl[1][o1 + 1];
        T sp0 = a.hy * (relu2(pu.b + c0.k) * (pu.a - c0.p) - fp_mid);
        T sp1 = a.hy * (fp_mid - relu2(c1.k + pd.b) * (c1.p - pd.a));
        T ss0 = a.gy * (relu2(su.b + c0.b) * (su.a - c0.s) - fs_mid);
        T ss1 = a.gy * (fs_mid - relu2(c1.b + sd.b) * (c1.s - sd.a));
        if (LL) {
            load(p0, n0);
            load(p1, n1);
        }
        sp0 += a.hz * (relu2(pl0.b + c0.k) * (pl0.a - c0.p) - relu2(c0.k + pr0.b) * (c0.p - pr0.a));
        sp1 += a.hz * (relu2(pl1.b + c1.k) * (pl1.a - c1.p) - relu2(c1.k + pr1.b) * (c1.p - pr1.a));
        ss0 += a.gz * (relu2(sl0.b + c0.b) * (sl0.a - c0.s) - relu2(c0.b + sr0.b) * (c0.s - sr0.a));
        ss1 += a.gz * (relu2(sl1.b + c1.b) * (sl1.a - c1.s) - relu2(c1.b + sr1.b) * (c1.s - sr1.a));
        // x faces last: plane x+1 has had the whole step to arrive
        p0.k = keff(p0.p, p0.n, p0.k);
        p1.k = keff(p1.p, p1.n, p1.k);
        const T fp_hi0 = relu2(c0.k + p0.k) * (c0.p - p0.p), fs_hi0 = relu2(c0.b + p0.b) * (c0.s - p0.s);
        const T fp_hi1 = relu2(c1.k + p1.k) * (c1.p - p1.p), fs_hi1 = relu2(c1.b + p1.b) * (c1.s - p1.s);
        sp0 += a.hx * (fp_lo0 - fp_hi0);
        sp1 += a.hx * (fp_lo1 - fp_hi1);
        ss0 += a.gx * (fs_lo0 - fs_hi0);
        ss1 += a.gx * (fs_lo1 - fs_hi1);
        pn0 = c0.p + a.dt * (sp0 + react0); pn1 = c1.p + a.dt * (sp1 + react1);
        // P has just entered a warp that held none (the front moves one voxel per step): now the Heaviside is needed
        if (!warp_has_p && __any_sync(0xffffffffu, (pn0 != T(0)) || (pn1 != T(0)))) {
            H0 = heaviside50_fast(c0.s, a.sigma_np); H1 = heaviside50_fast(c1.s, a.sigma_np);
        }
        nn0 = c0.n + a.dt * ((a.lambda_np * pn0) * H0); nn1 = c1.n + a.dt * ((a.lambda_np * pn1) * H1);
        sn0 = c0.s + a.dt * (ss0 - (a.lambda_s * c0.s) * pn0); sn1 = c1.s + a.dt * (ss1 - (a.lambda_s * c1.s) * pn1);
        fp_lo0 = fp_hi0; fs_lo0 = fs_hi0; fp_lo1 = fp_hi1; fs_lo1 = fs_hi1;
        }
        const bool cnt = counts(a, x);
        if (g.valid0) {
            Po[g.c0] = pn0; No[g.c0] = nn0; So[g.c0] = sn0;
            if (cnt) { sP += (double)pn0; sN += (double)nn0; }
        }
        if (g.valid1) {
            Po[g.c1] = pn1; No[g.c1] = nn1; So[g.c1] = sn1;
            if (cnt) { sP += (double)pn1; sN += (double)nn1; }
        }
        if (SLAB && sl.lo && x == a.x_lo) { push(a.peer_lo[sc.par ^ 1], a.off_lo, pn0, nn0, sn0, pn1, nn1, sn1); }
        if (SLAB && sl.hi && x == a.x_hi - 1) { push(a.peer_hi[sc.par ^ 1], a.off_hi, pn0, nn0, sn0, pn1, nn1, sn1); }
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
    };
    for (int i = 0; i < mm.n; i += 2) {
        step(mm.xs + i * mm.dx, A0, A1, B0, B1);
        if (i + 1 >= mm.n) break;
        step(mm.xs + (i + 1) * mm.dx, B0, B1, A0, A1);
    }
    step_end<T, 2>(a, sc, sP, sN);
}

// ---- FK, two voxels per thread (see fk2c_step_v4) -------------------------------------------------
template <typename T> struct VoxFk { T u, k; };

template <typename T, int TZ, int TYT, int MINB = 4, int SLAB = 0>
__global__ void __launch_bounds__(TZ *TYT, MINB) fk_step_v4(const StepArgs<T> a, const int lx)
{
    constexpr int TYV = 2 * TYT, RS = TZ + 2, PL = (TYV + 2) * RS;
    __shared__ Pair<T> tile[2][PL];        // [buffer][slot] of (u, k)
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ kc = a.coef[0];
    Stage2 g = stage2_setup<T, TZ, TYT, RS, SLAB>(a, lx);
    const March2 mm = march2_setup<T, SLAB>(a, g, sc.par);
    double s = 0.0;
    griddep_launch();
    griddep_wait();
    const SlabCtl sl = SLAB ? slab_begin(a, g.x0, g.x1) : SlabCtl{false, false, 0};
    auto push = [&](T *const *peer, uint32_t off, T v0, T v1) {
        if (g.valid0) peer[0][g.c0 + off] = v0;
        if (g.valid1) peer[0][g.c1 + off] = v1;
    };

    VoxFk<T> A0, A1, B0, B1;
    A0.u = u[g.c0]; A0.k = kc[g.c0];
    A1.u = u[g.c1]; A1.k = kc[g.c1];
    T f_lo0 = relu2(kc[g.prev0] + A0.k) * (u[g.prev0] - A0.u);
    T f_lo1 = relu2(kc[g.prev1] + A1.k) * (u[g.prev1] - A1.u);
    T hu = 0, hk = 0;
    if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto step = [&](int x, const VoxFk<T> &c0, const VoxFk<T> &c1, VoxFk<T> &p0, VoxFk<T> &p1) {
        Pair<T> *tl = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        tl[o0] = Pair<T>{c0.u, c0.k};
        tl[o1] = Pair<T>{c1.u, c1.k};
        if (g.ring) tl[g.hs] = Pair<T>{hu, hk};
        const uint32_t adv = (x + 1 == mm.wrapx) ? mm.wrapd : mm.dsx;
        const uint32_t n0 = g.c0 + adv, n1 = g.c1 + adv;
        p0.u = u[n0]; p0.k = kc[n0];
        p1.u = u[n1]; p1.k = kc[n1];
        g.h += adv;
        if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x, mm.dx);
            l2_prefetch_strip(u, q, pf.count, pf.total); l2_prefetch_strip(kc, q, pf.count, pf.total);
        }
        // a warp whose 64 voxels all lie outside the tissue (k = "closed": every face coefficient is max(closed + k', 0) = 0)
        // and hold u == 0 returns u' = 0 in the reference's arithmetic: it loads, publishes, takes the barrier and stores,
        // but skips the shared-memory reads and the arithmetic (see fk2c_step_v4)
        const T closed_k = -Big<T>::v;
        const bool warp_closed = a.skip && __all_sync(0xffffffffu, (c0.k == closed_k) && (c1.k == closed_k) && (c0.u == T(0)) && (c1.u == T(0)));
        T un0, un1;
        if (warp_closed) {
            __syncthreads();
            un0 = c0.u; un1 = c1.u;
            f_lo0 = f_lo1 = T(0);
        } else {
        const T react0 = a.rho * c0.u * (T(1) - c0.u), react1 = a.rho * c1.u * (T(1) - c1.u);
        const T f_mid = relu2(c0.k + c1.k) * (c0.u - c1.u);
        __syncthreads();
        const Pair<T> up = tl[o0 - RS], dn = tl[o1 + RS];
        const Pair<T> l0 = tl[o0 - 1], r0 = tl[o0 + 1], l1 = tl[o1 - 1], r1 = tl[o1 + 1];
        T sp0 = a.hy * (relu2(up.b + c0.k) * (up.a - c0.u) - f_mid);
        T sp1 = a.hy * (f_mid - relu2(c1.k + dn.b) * (c1.u - dn.a));
        sp0 += a.hz * (relu2(l0.b + c0.k) * (l0.a - c0.u) - relu2(c0.k + r0.b) * (c0.u - r0.a));
        sp1 += a.hz * (relu2(l1.b + c1.k) * (l1.a - c1.u) - relu2(c1.k + r1.b) * (c1.u - r1.a));
        const T f_hi0 = relu2(c0.k + p0.k) * (c0.u - p0.u), f_hi1 = relu2(c1.k + p1.k) * (c1.u - p1.u);
        sp0 += a.hx * (f_lo0 - f_hi0);
        sp1 += a.hx * (f_lo1 - f_hi1);
        un0 = c0.u + a.dt * (sp0 + react0); un1 = c1.u + a.dt * (sp1 + react1);
        f_lo0 = f_hi0; f_lo1 = f_hi1;
        }
        const bool cnt = counts(a, x);
        if (g.valid0) { out[g.c0] = un0; if (cnt) s += (double)un0; }
        if (g.valid1) { out[g.c1] = un1; if (cnt) s += (double)un1; }
        if (SLAB && sl.lo && x == a.x_lo) { push(a.peer_lo[sc.par ^ 1], a.off_lo, un0, un1); }
        if (SLAB && sl.hi && x == a.x_hi - 1) { push(a.peer_hi[sc.par ^ 1], a.off_hi, un0, un1); }
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
    };
    for (int i = 0; i < mm.n; i += 2) {
        step(mm.xs + i * mm.dx, A0, A1, B0, B1);
        if (i + 1 >= mm.n) break;
        step(mm.xs + (i + 1) * mm.dx, B0, B1, A0, A1);
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK_DTI, two voxels per thread (reference rounding order, bitwise equal to dti_step_v3) ----------
template <typename T> struct VoxDti { T u, d0, d1, d2; };

template <typename T, int TZ, int TYT, int MINB = 4, int SLAB = 0>
__global__ void __launch_bounds__(TZ *TYT, MINB) dti_step_v4(const StepArgs<T> a, const int lx)
{
    using R = Rn<T>;
    constexpr int TYV = 2 * TYT, RS = TZ + 2, PL = (TYV + 2) * RS;
    __shared__ T tile[2][PL];
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ dxa = a.coef[0];
    const T *__restrict__ dya = a.coef[1];
    const T *__restrict__ dza = a.coef[2];
    Stage2 g = stage2_setup<T, TZ, TYT, RS, SLAB>(a, lx);
    const March2 mm = march2_setup<T, SLAB>(a, g, sc.par);
    double s = 0.0;
    auto load = [&](VoxDti<T> &v, uint32_t off) { v.u = u[off]; v.d0 = dxa[off]; v.d1 = dya[off]; v.d2 = dza[off]; };

    VoxDti<T> A0, A1, B0, B1;
    griddep_launch();
    griddep_wait();
    const SlabCtl sl = SLAB ? slab_begin(a, g.x0, g.x1) : SlabCtl{false, false, 0};
    auto push = [&](T *const *peer, uint32_t off, T v0, T v1) {
        if (g.valid0) peer[0][g.c0 + off] = v0;
        if (g.valid1) peer[0][g.c1 + off] = v1;
    };
    load(A0, g.c0);
    load(A1, g.c1);
    T um0 = u[g.prev0], um1 = u[g.prev1];
    T hu = 0;
    if (g.ring) hu = u[g.h];
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto one = [&](const VoxDti<T> &c, T um, T up, T uym, T uyp, T uzm, T uzp) {
        const T sx = R::mul(a.cx, R::sub(R::mul(c.d0, R::sub(um, c.u)), R::mul(c.d0, R::sub(c.u, up))));
        const T sy = R::mul(a.cy, R::sub(R::mul(c.d1, R::sub(uym, c.u)), R::mul(c.d1, R::sub(c.u, uyp))));
        const T sz = R::mul(a.cz, R::sub(R::mul(c.d2, R::sub(uzm, c.u)), R::mul(c.d2, R::sub(c.u, uzp))));
        const T sp = R::add(R::add(sx, sy), sz);
        const T react = R::mul(a.rho, R::mul(c.u, R::sub(T(1), c.u)));
        return R::add(c.u, R::mul(R::add(sp, react), a.dt));
    };
    auto step = [&](int x, const VoxDti<T> &c0, const VoxDti<T> &c1, VoxDti<T> &p0, VoxDti<T> &p1) {
        T *tl = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        tl[o0] = c0.u;
        tl[o1] = c1.u;
        if (g.ring) tl[g.hs] = hu;
        const uint32_t adv = (x + 1 == mm.wrapx) ? mm.wrapd : mm.dsx;
        const uint32_t n0 = g.c0 + adv, n1 = g.c1 + adv;
        load(p0, n0);
        load(p1, n1);
        g.h += adv;
        if (g.ring) hu = u[g.h];
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x, mm.dx);
            l2_prefetch_strip(u, q, pf.count, pf.total); l2_prefetch_strip(dxa, q, pf.count, pf.total);
            l2_prefetch_strip(dya, q, pf.count, pf.total); l2_prefetch_strip(dza, q, pf.count, pf.total);
        }
        __syncthreads();
        // (no closed-warp shortcut here: FK_DTI streams at 0.85 of the HBM roofline, and the vote over eight compares cost
        // 3 % on the atlas box -- profiles/r2_sweep_misc.txt)
        // um* = the plane behind the march, p*.u the plane ahead: x-1 / x+1 when marching up, swapped when marching down
        const T un0 = one(c0, mm.down ? p0.u : um0, mm.down ? um0 : p0.u, tl[o0 - RS], c1.u, tl[o0 - 1], tl[o0 + 1]);
        const T un1 = one(c1, mm.down ? p1.u : um1, mm.down ? um1 : p1.u, c0.u, tl[o1 + RS], tl[o1 - 1], tl[o1 + 1]);
        const bool cnt = counts(a, x);
        if (g.valid0) { out[g.c0] = un0; if (cnt) s += (double)un0; }
        if (g.valid1) { out[g.c1] = un1; if (cnt) s += (double)un1; }
        if (SLAB && sl.lo && x == a.x_lo) { push(a.peer_lo[sc.par ^ 1], a.off_lo, un0, un1); }
        if (SLAB && sl.hi && x == a.x_hi - 1) { push(a.peer_hi[sc.par ^ 1], a.off_hi, un0, un1); }
        um0 = c0.u; um1 = c1.u;
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
    };
    for (int i = 0; i < mm.n; i += 2) {
        step(mm.xs + i * mm.dx, A0, A1, B0, B1);
        if (i + 1 >= mm.n) break;
        step(mm.xs + (i + 1) * mm.dx, B0, B1, A0, A1);
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK --------------------------------------------------------------------------------------------
template <typename T> struct ColFkS { T u, k; };

template <typename T, int TZ, int TY>
__global__ void __launch_bounds__(TZ *TY) fk_step_v3(const StepArgs<T> a, const int lx)
{
    constexpr int RS = TZ + 2, PL = (TY + 2) * RS;
    __shared__ Pair<T> tile[2][PL];        // [buffer][slot] of (u, k)
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ kc = a.coef[0];
    Stage g = stage_setup<T, TZ, TY, RS>(a, lx);
    double s = 0.0;

    ColFkS<T> A, B, C;
    A.u = u[g.c]; A.k = kc[g.c];
    T flux_lo = relu(kc[g.prev] + A.k) * (u[g.prev] - A.u);
    uint32_t n1 = stage_up(a, g, g.c, g.x0);
    B.u = u[n1]; B.k = kc[n1];
    T hu = 0, hk = 0;
    if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
    int buf = 0;

    auto step = [&](int x, const ColFkS<T> &c, const ColFkS<T> &p, ColFkS<T> &q) {
        Pair<T> *tl = tile[buf];
        tl[g.so] = Pair<T>{c.u, c.k};
        if (g.ring) tl[g.hs] = Pair<T>{hu, hk};
        const int x1 = (x + 1 == a.X) ? 0 : x + 1;
        const uint32_t n2 = stage_up(a, g, n1, x1);
        q.u = u[n2]; q.k = kc[n2];
        g.h = stage_up(a, g, g.h, x);
        if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
        const T flux_hi = relu(c.k + p.k) * (c.u - p.u);
        T sp = a.cx * (flux_lo - flux_hi);
        const T react = a.rho * c.u * (T(1) - c.u);
        __syncthreads();
        const int o = g.so;
        const Pair<T> ym = tl[o - RS], yp = tl[o + RS], zm = tl[o - 1], zp = tl[o + 1];
        sp += a.cy * (relu(ym.b + c.k) * (ym.a - c.u) - relu(c.k + yp.b) * (c.u - yp.a));
        sp += a.cz * (relu(zm.b + c.k) * (zm.a - c.u) - relu(c.k + zp.b) * (c.u - zp.a));
        const T un = c.u + a.dt * (sp + react);
        if (g.valid) {
            out[g.c] = un;
            if (counts(a, x)) s += (double)un;
        }
        flux_lo = flux_hi;
        g.c = n1; n1 = n2;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 3) {
        step(x, A, B, C);
        if (x + 1 >= g.x1) break;
        step(x + 1, B, C, A);
        if (x + 2 >= g.x1) break;
        step(x + 2, C, A, B);
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK_DTI (reference rounding order, see coef_step_v1 mode 2) -------------------------------------
template <typename T> struct ColDtiS { T u, d0, d1, d2; };

template <typename T, int TZ, int TY>
__global__ void __launch_bounds__(TZ *TY) dti_step_v3(const StepArgs<T> a, const int lx)
{
    using R = Rn<T>;
    constexpr int RS = TileGeom<T, TZ>::RS, PL = (TY + 2) * RS;
    __shared__ T tile[2][PL];
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ dxa = a.coef[0];
    const T *__restrict__ dya = a.coef[1];
    const T *__restrict__ dza = a.coef[2];
    Stage g = stage_setup<T, TZ, TY>(a, lx);
    double s = 0.0;

    ColDtiS<T> A, B, C;
    T um = u[g.prev];
    A.u = u[g.c]; A.d0 = dxa[g.c]; A.d1 = dya[g.c]; A.d2 = dza[g.c];
    uint32_t n1 = stage_up(a, g, g.c, g.x0);
    B.u = u[n1]; B.d0 = dxa[n1]; B.d1 = dya[n1]; B.d2 = dza[n1];
    T hu = 0;
    if (g.ring) hu = u[g.h];
    int buf = 0;

    auto step = [&](int x, const ColDtiS<T> &c, const ColDtiS<T> &p, ColDtiS<T> &q) {
        T *tl = tile[buf];
        tl[g.so] = c.u;
        if (g.ring) tl[g.hs] = hu;
        const int x1 = (x + 1 == a.X) ? 0 : x + 1;
        const uint32_t n2 = stage_up(a, g, n1, x1);
        q.u = u[n2]; q.d0 = dxa[n2]; q.d1 = dya[n2]; q.d2 = dza[n2];
        g.h = stage_up(a, g, g.h, x);
        if (g.ring) hu = u[g.h];
        const T sx = R::mul(a.cx, R::sub(R::mul(c.d0, R::sub(um, c.u)), R::mul(c.d0, R::sub(c.u, p.u))));
        const T react = R::mul(a.rho, R::mul(c.u, R::sub(T(1), c.u)));
        __syncthreads();
        const int o = g.so;
        const T sy = R::mul(a.cy, R::sub(R::mul(c.d1, R::sub(tl[o - RS], c.u)), R::mul(c.d1, R::sub(c.u, tl[o + RS]))));
        const T sz = R::mul(a.cz, R::sub(R::mul(c.d2, R::sub(tl[o - 1], c.u)), R::mul(c.d2, R::sub(c.u, tl[o + 1]))));
        const T sp = R::add(R::add(sx, sy), sz);
        const T un = R::add(c.u, R::mul(R::add(sp, react), a.dt));
        if (g.valid) {
            out[g.c] = un;
            if (counts(a, x)) s += (double)un;
        }
        um = c.u;
        g.c = n1; n1 = n2;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 3) {
        step(x, A, B, C);
        if (x + 1 >= g.x1) break;
        step(x + 1, B, C, A);
        if (x + 2 >= g.x1) break;
        step(x + 2, C, A, B);
    }
    step_end<T, 1>(a, sc, s, 0.0);
}

}  // namespace tgtk
