// Sparse launches of the two-voxel marching kernels (variant 4 on a grid with dead space).
//
// The solvers run on the bounding box of the brain (FK.py:155-170, FK_2c.py:170-185, FK_DTI.py:169-171): on the SRI atlas
// 64 % of that box is background.  A voxel outside the tissue carries the coefficient "closed" -- each of its six faces
// has the coefficient max(closed + k', 0) = 0 -- so with P = N = 0 the reference's update returns it unchanged, bit for
// bit, in every step of the loop: it is a fixed point that no neighbour can disturb.  A WARP TILE (2 rows x 32 voxels of
// one plane, the voxels one warp of the two-voxel kernels owns) that consists of such voxels only is "dead".  Both
// ping-pong buffers are given the same initial content, after which a dead tile is
//   * never loaded   (its owner fills the shared-memory tile with P = S = 0, k = b = closed: that is all a neighbour's
//                     face towards it can see, any finite value times the coefficient 0),
//   * never computed,
//   * never stored   (both buffers already hold its value),
// and the launch grid is a WORK LIST: per 8 x 32 column of the box the plane ranges that contain live tiles, cut into items
// of equal weight (tgtk_api.cu: build_sparse), so no block is scheduled for the dead space and the SMs stay evenly loaded
// although the tissue is not evenly spread.  Results are bitwise those of the dense kernel (tests/test_gpu_sparse.py).
//
// live[(x * nzt + zt) * nyt + yt]: one 32-bit word per block tile of a plane, byte w = warp w's two rows are live.
#pragma once
#include "tgtk_kernels_v3.cuh"

namespace tgtk {

struct SparseArgs {
    const uint32_t *live;     // [X][nzt][nyt]
    const int4 *work;         // {y tile, z tile, x0, x1}
    int nzt, nyt;
};

// one thread per warp tile; a word is assembled by the four threads of a block tile
template <typename T, int NF>
__global__ void __launch_bounds__(128) sparse_mask_kernel(const T *__restrict__ f0, const T *__restrict__ f1, const T *__restrict__ c0,
                                                           const T *__restrict__ c1, uint32_t *__restrict__ live, int X, int Y, int Z,
                                                           int64_t sx, int64_t sy, int nzt, int nyt, int mode)
{
    // block = (32 lanes, 4 warps): lanes scan z, warp w scans rows 2w, 2w+1 of the tile
    const int yt = blockIdx.x % nyt, zt = (blockIdx.x / nyt) % nzt, x = blockIdx.x / (nyt * nzt);
    const int z = zt * 32 + threadIdx.x, w = threadIdx.y;
    bool alive = false;
    const T closed = -Big<T>::v;
    if (z < Z) {
        for (int r = 0; r < 2; ++r) {
            const int y = yt * 8 + 2 * w + r;
            if (y >= Y) continue;
            const int64_t q = x * sx + y * sy + z;
            bool dead;
            if (mode == 0) {          // FK_2c: P, N; k, b
                dead = (c0[q] == closed) && (c1[q] == closed) && (f0[q] == T(0)) && (f1[q] == T(0));
            } else {                  // FK: u; k
                dead = (c0[q] == closed) && (f0[q] == T(0));
            }
            alive |= !dead;
        }
    }
    const unsigned any = __any_sync(0xffffffffu, alive) ? 1u : 0u;
    __shared__ unsigned bytes[4];
    if (threadIdx.x == 0) bytes[w] = any;
    __syncthreads();
    if (threadIdx.x == 0 && w == 0)
        live[blockIdx.x] = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
}

// Stage2 of a work item (see stage2_setup): tile (yt, zt), planes [x0, x1)
template <typename T, int RS>
__device__ __forceinline__ Stage2 stage2_item(const StepArgs<T> &a, const int4 it)
{
    constexpr int TZ = 32, TYV = 8;
    Stage2 g;
    const int tz = threadIdx.x, ty = threadIdx.y, tid = ty * TZ + tz;
    const int z0 = it.y * TZ, y0 = it.x * TYV;
    const int r0 = y0 + 2 * ty;
    g.valid0 = (z0 + tz < a.Z) && (r0 < a.Y);
    g.valid1 = (z0 + tz < a.Z) && (r0 + 1 < a.Y);
    g.x0 = it.z;
    g.x1 = it.w;
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

// the words of planes x+1 .. x+3 travel in registers, so the decision "load plane x+1 or not" never waits
struct LiveWords {
    const uint32_t *col;      // &live[(0 * nzt + zt) * nyt + yt]
    int stride, X;            // words per plane
    uint32_t w0, w1, w2, w3;  // planes x, x+1, x+2, x+3 (periodic)
    __device__ __forceinline__ uint32_t at(int x) const
    {
        if (x >= X) x -= X;
        if (x < 0) x += X;
        return __ldg(col + (size_t)x * stride);
    }
    __device__ __forceinline__ void start(int x)
    {
        w0 = at(x); w1 = at(x + 1); w2 = at(x + 2); w3 = at(x + 3);
    }
    __device__ __forceinline__ void advance(int x)      // x = the plane just finished
    {
        w0 = w1; w1 = w2; w2 = w3; w3 = at(x + 4);
    }
};

// L2 prefetch of the live rows of the own tile in plane xp: lanes 0 .. 2*NA-1 of a warp issue one bulk request each
// (row lane / NA of the warp's two, array lane % NA -- `base` is that array, picked once per kernel)
template <typename T, int NA>
__device__ __forceinline__ void l2_prefetch_tile(const T *base, const StepArgs<T> &a, int xp, int y0, int z0, uint32_t word)
{
    const int lane = threadIdx.x, ty = threadIdx.y;
    if (lane >= 2 * NA || !((word >> (8 * ty)) & 0xffu)) return;
    const int row = y0 + 2 * ty + lane / NA;
    if (row >= a.Y) return;
    if (xp >= a.X) xp -= a.X;
    const uint32_t first = (uint32_t)xp * (uint32_t)a.sx + (uint32_t)row * (uint32_t)a.sy + (uint32_t)z0;
    const uint32_t count = (uint32_t)min(32, a.Z - z0);
    l2_prefetch_strip(base, first, count, (uint32_t)a.X * (uint32_t)a.sx);
}

// ---- FK_2c ---------------------------------------------------------------------------------------------------------
template <typename T, int MINB>
__global__ void __launch_bounds__(128, MINB) fk2c_step_sp(const StepArgs<T> a, const SparseArgs sp)
{
    constexpr int TZ = 32, TYV = 8, RS = TZ + 2, PL = (TYV + 2) * RS;
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
    const int4 item = sp.work[blockIdx.x];
    Stage2 g = stage2_item<T, RS>(a, item);
    LiveWords lw;
    lw.col = sp.live + (size_t)item.y * sp.nyt + item.x;
    lw.stride = sp.nzt * sp.nyt;
    lw.X = a.X;
    lw.start(g.x0);
    const int sh = 8 * threadIdx.y;
    const T closed = -Big<T>::v;
    auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
    auto load = [&](Vox2c<T> &v, uint32_t off) { v.p = P[off]; v.n = N[off]; v.s = S[off]; v.b = bc[off]; v.k = kc[off]; };
    auto none = [&](Vox2c<T> &v) { v.p = T(0); v.n = T(0); v.s = T(0); v.b = closed; v.k = closed; };
    double sP = 0.0, sN = 0.0;
    griddep_launch();
    griddep_wait();

    Vox2c<T> A0, A1, B0, B1;               // rows (y, y+1) of plane x (A) and plane x+1 (B); roles swap
    bool liveA = (lw.w0 >> sh) & 0xffu, liveB = false;
    if (liveA) { load(A0, g.c0); load(A1, g.c1); } else { none(A0); none(A1); }
    A0.k = keff(A0.p, A0.n, A0.k);
    A1.k = keff(A1.p, A1.n, A1.k);
    T fp_lo0 = T(0), fs_lo0 = T(0), fp_lo1 = T(0), fs_lo1 = T(0);      // fluxes through the faces below (x-1 | x)
    if (liveA) {
        // (a dead tile below: P = 0 and closed coefficients there, which the loads return as well)
        const T pm0 = P[g.prev0], pm1 = P[g.prev1];
        fp_lo0 = relu2(keff(pm0, N[g.prev0], kc[g.prev0]) + A0.k) * (pm0 - A0.p);
        fs_lo0 = relu2(bc[g.prev0] + A0.b) * (S[g.prev0] - A0.s);
        fp_lo1 = relu2(keff(pm1, N[g.prev1], kc[g.prev1]) + A1.k) * (pm1 - A1.p);
        fs_lo1 = relu2(bc[g.prev1] + A1.b) * (S[g.prev1] - A1.s);
    }
    T hp = 0, hn = 0, hk = 0, hsv = 0, hb = 0;
    if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
    int buf = 0;
    const int pfi = threadIdx.x % 5;
    const T *pf_base = (pfi == 0) ? P : (pfi == 1) ? N : (pfi == 2) ? S : (pfi == 3) ? kc : bc;
    const int y0 = item.x * TYV, z0 = item.y * TZ;

    auto step = [&](int x, const bool live, bool &live_n, const Vox2c<T> &c0, const Vox2c<T> &c1, Vox2c<T> &p0, Vox2c<T> &p1) {
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
        // request the next plane of the march: both own voxels (if that tile is live) and the ring
        const uint32_t adv = (x + 1 == a.X) ? 0u - g.wrap_back : g.sx;
        const uint32_t n0 = g.c0 + adv, n1 = g.c1 + adv;
        live_n = (lw.w1 >> sh) & 0xffu;
        if (live_n) { load(p0, n0); load(p1, n1); } else { none(p0); none(p1); }
        g.h += adv;
        if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
        if (a.l2pf) l2_prefetch_tile<T, 5>(pf_base, a, x + 2, y0, z0, lw.w2);
        T pn0, pn1, nn0, nn1, sn0, sn1;
        if (!live) {
            __syncthreads();
            p0.k = keff(p0.p, p0.n, p0.k);
            p1.k = keff(p1.p, p1.n, p1.k);
            fp_lo0 = fs_lo0 = fp_lo1 = fs_lo1 = T(0);      // the faces towards the next plane are closed from this side
        } else {
            // a warp whose 64 voxels hold no tumour skips the Heaviside (see fk2c_step_v4)
            const bool warp_has_p = !a.skip || __any_sync(0xffffffffu, (c0.p != T(0)) || (c1.p != T(0)));
            T H0 = T(0), H1 = T(0);
            if (warp_has_p) { H0 = heaviside50_fast(c0.s, a.sigma_np); H1 = heaviside50_fast(c1.s, a.sigma_np); }
            const T react0 = a.rho * (c0.s * c0.p) * ((T(1) - c0.p) - c0.n) - (a.lambda_np * c0.p) * H0;
            const T react1 = a.rho * (c1.s * c1.p) * ((T(1) - c1.p) - c1.n) - (a.lambda_np * c1.p) * H1;
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
            if (!warp_has_p && __any_sync(0xffffffffu, (pn0 != T(0)) || (pn1 != T(0)))) {
                H0 = heaviside50_fast(c0.s, a.sigma_np); H1 = heaviside50_fast(c1.s, a.sigma_np);
            }
            nn0 = c0.n + a.dt * ((a.lambda_np * pn0) * H0); nn1 = c1.n + a.dt * ((a.lambda_np * pn1) * H1);
            sn0 = c0.s + a.dt * (ss0 - (a.lambda_s * c0.s) * pn0); sn1 = c1.s + a.dt * (ss1 - (a.lambda_s * c1.s) * pn1);
            fp_lo0 = fp_hi0; fs_lo0 = fs_hi0; fp_lo1 = fp_hi1; fs_lo1 = fs_hi1;
            const bool cnt = counts(a, x);
            if (g.valid0) {
                Po[g.c0] = pn0; No[g.c0] = nn0; So[g.c0] = sn0;
                if (cnt) { sP += (double)pn0; sN += (double)nn0; }
            }
            if (g.valid1) {
                Po[g.c1] = pn1; No[g.c1] = nn1; So[g.c1] = sn1;
                if (cnt) { sP += (double)pn1; sN += (double)nn1; }
            }
        }
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
        lw.advance(x);
    };
    const int n = g.x1 - g.x0;
    for (int i = 0; i < n; i += 2) {
        step(g.x0 + i, liveA, liveB, A0, A1, B0, B1);
        if (i + 1 >= n) break;
        step(g.x0 + i + 1, liveB, liveA, B0, B1, A0, A1);
    }
    step_end<T, 2>(a, sc, sP, sN);
}

}  // namespace tgtk
