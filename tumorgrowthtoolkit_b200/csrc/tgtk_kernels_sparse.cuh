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
    const int4 *work;         // two per item: {y tile, z tile, x0, x1}, {live words of planes x0 .. x0+3}
    int ahead;                // items between an item and the one whose first planes it asks into L2 (0: none)
    int nzt, nyt;
    unsigned long long *trace;     // nullptr, or [items][4]: globaltimer at block start / after the prologue / at the end, SM id
};

// one thread per warp tile; a word is assembled by the four threads of a block tile
template <typename T, int NF>
__global__ void __launch_bounds__(128) sparse_mask_kernel(const T *__restrict__ f0, const T *__restrict__ f1, const T *__restrict__ c0,
                                                           const T *__restrict__ c1, const T *__restrict__ c2, uint32_t *__restrict__ live, int X, int Y, int Z,
                                                           int64_t sx, int64_t sy, int nzt, int nyt, int mode)
{
    // block = (32 lanes, 4 warps): lanes scan z, warp w scans rows 2w, 2w+1 of the tile
    const int yt = blockIdx.x % nyt, zt = (blockIdx.x / nyt) % nzt, x = blockIdx.x / (nyt * nzt);
    const int z = zt * 32 + threadIdx.x, w = threadIdx.y;
    bool alive = false;
    const T closed = -Big<T>::v;
    if (z < Z) {
        for (int r = 0; r < 2; ++r) {
            // Rows beyond the box (the last y tile of a box whose Y is not a multiple of 8) hold the periodically WRAPPED rows
            // in the step kernels, and the FIRST of them (y == Y, mirroring row 0) is the lower neighbour of the box's last
            // row: where it starts a warp of its own (Y even) that warp is live when row 0 is, or a box whose tissue touches
            // its y faces would lose the coupling (tests/test_gpu_sparse.py::test_sparse_on_boxes_narrower_than_a_tile).
            // Such a warp stores nothing.  (Y odd: row Y is the second row of the warp that owns row Y-1 and is loaded with
            // it.)  Later mirror rows are nobody's neighbour.
            int y = yt * 8 + 2 * w + r;
            if (y >= Y) {
                if (y != Y) continue;
                y = 0;
            }
            const int64_t q = x * sx + y * sy + z;
            bool dead;
            if (mode == 0) {          // FK_2c: P, N; k, b
                dead = (c0[q] == closed) && (c1[q] == closed) && (f0[q] == T(0)) && (f1[q] == T(0));
            } else if (mode == 1) {   // FK: u; k
                dead = (c0[q] == closed) && (f0[q] == T(0));
            } else {                  // FK_DTI: u; the voxel's own three coefficients (both faces of an axis carry the voxel's
                                      // own coefficient, FK_DTI.py:174: D = 0 and u = 0 is a fixed point whatever the neighbours hold)
                dead = (c0[q] == T(0)) && (c1[q] == T(0)) && (c2[q] == T(0)) && (f0[q] == T(0));
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

// wrap_coord for v in [-1, n + 31]: no integer division unless the box is narrower than a tile
__device__ __forceinline__ int wrap_near(int v, int n)
{
    if (v < 0) return v + n;
    if (v >= n) {
        v -= n;
        if (v >= n) v %= n;
    }
    return v;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void sparse_trace(const SparseArgs &sp, int slot)
{
    if (sp.trace && threadIdx.x == 0 && threadIdx.y == 0) {
        sp.trace[4 * (size_t)blockIdx.x + slot] = global_ns();
        if (slot == 0) {
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            sp.trace[4 * (size_t)blockIdx.x + 3] = sm;
        }
    }
}

// -DTGTK_SP_PHASES (diagnostic build, tools/sparse_phases.py): cycles every warp of fk2c_step_sp spends in the phases of a
// plane step, summed over the item's live planes -> trace[(item * 4 + warp) * 8 + phase]
#ifdef TGTK_SP_PHASES
#define SP_CLOCK(var, dep) long long var; asm volatile("mov.u64 %0, %%clock64;" : "=l"(var) : "d"((double)(dep)) : "memory")
#else
#define SP_CLOCK(var, dep)
#endif

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
    const uint32_t zc = (uint32_t)wrap_near(z0 + tz, a.Z);
    g.c0 = base + (uint32_t)wrap_near(r0, a.Y) * sy + zc;
    g.c1 = base + (uint32_t)wrap_near(r0 + 1, a.Y) * sy + zc;
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
    g.h = base + (uint32_t)wrap_near(y0 - 1 + hr, a.Y) * sy + (uint32_t)wrap_near(z0 - 1 + hc, a.Z);
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

// L2 prefetch of the live rows of the own tile in plane xp: ONE prefetch.global.L2 per warp and plane.  Lane l < 6 * NA
// asks for 128-byte line (l % 3) of row (l % 6) / 3 of the warp's two, array l / 6 -- a 32-voxel row segment spans at most
// three lines whatever its alignment.  (A bulk request per row through the TMA unit -- cp.async.bulk.prefetch.L2 -- takes
// uniform operands: ten lanes with ten addresses became a ten-trip loop, 17 % of the kernel's instructions.)
struct TilePrefetch {
    const char *base;       // the lane's array, nullptr: lane takes no part
    uint32_t row_off;       // element offset of the lane's row segment inside a plane
    int line;               // which 128-byte line of the segment
    uint32_t seg_bytes;
};
template <typename T, int NA>
__device__ __forceinline__ TilePrefetch tile_prefetch_setup(const StepArgs<T> &a, int y0, int z0, const T *p0, const T *p1,
                                                            const T *p2 = nullptr, const T *p3 = nullptr, const T *p4 = nullptr)
{
    TilePrefetch t;
    const int lane = threadIdx.x, row = y0 + 2 * (int)threadIdx.y + (lane % 6) / 3, arr = lane / 6;
    const T *p = (arr == 0) ? p0 : (arr == 1) ? p1 : (arr == 2) ? p2 : (arr == 3) ? p3 : p4;
    t.base = (lane < 6 * NA && row < a.Y) ? (const char *)p : nullptr;
    t.row_off = (uint32_t)row * (uint32_t)a.sy + (uint32_t)z0;
    t.line = lane % 3;
    t.seg_bytes = (uint32_t)min(32, a.Z - z0) * (uint32_t)sizeof(T);
    return t;
}
template <typename T>
__device__ __forceinline__ void tile_prefetch(const TilePrefetch &t, const StepArgs<T> &a, int xp, bool live)
{
    if (!live || !t.base) return;
    if (xp >= a.X) xp -= a.X;
    const char *first = t.base + ((size_t)((uint32_t)xp * (uint32_t)a.sx + t.row_off)) * sizeof(T);
    const char *line = (const char *)((uintptr_t)first & ~(uintptr_t)127) + 128 * t.line;
    if (line < first + t.seg_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
}

// Prologue help for a LATER item (the one `ahead` places further down the list -- about the time this block ends, that one
// starts on some SM): its first planes go to L2 now, so that its prologue -- item record, first plane, plane below, ring; three
// dependent trips to memory -- finds them there.  Thread t: array t % 5, tile row t / 5 - 1 (-1 .. 8, the ring rows included).
template <typename T>
__device__ __forceinline__ void sparse_prefetch_item(const StepArgs<T> &a, const SparseArgs &sp, int par)
{
    const int j = (int)blockIdx.x + sp.ahead;
    if (sp.ahead <= 0 || j >= (int)gridDim.x) return;
    const int t = threadIdx.y * 32 + threadIdx.x;
    if (t == 127) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp.work + 2 * j));
    if (t >= 50) return;
    const int4 it = sp.work[2 * j];
    const int arr = t % 5, row = wrap_near(it.x * 8 + t / 5 - 1, a.Y);
    const T *base = (arr == 0) ? a.buf[par][0] : (arr == 1) ? a.buf[par][1] : (arr == 2) ? a.buf[par][2] : (arr == 3) ? a.coef[0] : a.coef[1];
    const int z0 = max(it.y * 32 - 1, 0);
    const uint32_t seg = (uint32_t)min(34, a.Z - z0) * (uint32_t)sizeof(T);
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
        const int x = wrap_near(it.z + dx, a.X);
        const char *first = (const char *)(base + ((size_t)x * a.sx + (size_t)row * a.sy + z0));
        const char *line = (const char *)((uintptr_t)first & ~(uintptr_t)127);
        for (; line < first + seg; line += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
    }
}

// ---- FK_2c ---------------------------------------------------------------------------------------------------------
template <typename T, int MINB, int SLAB = 0>
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
    sparse_trace(sp, 0);
    const int4 item = sp.work[2 * blockIdx.x], item_w = sp.work[2 * blockIdx.x + 1];
    sparse_prefetch_item<T>(a, sp, sc.par);
    Stage2 g = stage2_item<T, RS>(a, item);
    LiveWords lw;
    lw.col = sp.live + (size_t)item.y * sp.nyt + item.x;
    lw.stride = sp.nzt * sp.nyt;
    lw.X = a.X;
    lw.w0 = item_w.x; lw.w1 = item_w.y; lw.w2 = item_w.z; lw.w3 = item_w.w;
    const int sh = 8 * threadIdx.y;
    const T closed = -Big<T>::v;
    auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
    auto load = [&](Vox2c<T> &v, uint32_t off) { v.p = P[off]; v.n = N[off]; v.s = S[off]; v.b = bc[off]; v.k = kc[off]; };
    auto none = [&](Vox2c<T> &v) { v.p = T(0); v.n = T(0); v.s = T(0); v.b = closed; v.k = closed; };
    double sP = 0.0, sN = 0.0;
    griddep_launch();
    griddep_wait();
    // fused slab mode (tgtk_common.cuh: slab_begin): item 0 publishes the epoch, an item that starts on the rank's lowest /
    // ends on its highest owned plane waits for that neighbour's word, reads its ghost plane and pushes its boundary plane
    const SlabCtl sl = SLAB ? slab_begin(a, g.x0, g.x1) : SlabCtl{false, false, 0};
    auto push = [&](T *const *peer, uint32_t off, T pn0, T nn0, T sn0, T pn1, T nn1, T sn1) {
        if (g.valid0) { peer[0][g.c0 + off] = pn0; peer[1][g.c0 + off] = nn0; peer[2][g.c0 + off] = sn0; }
        if (g.valid1) { peer[0][g.c1 + off] = pn1; peer[1][g.c1 + off] = nn1; peer[2][g.c1 + off] = sn1; }
    };

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
    const TilePrefetch pf = tile_prefetch_setup<T, 5>(a, item.x * TYV, item.y * TZ, P, N, S, kc, bc);

#ifdef TGTK_SP_PHASES
    long long ph[6] = {0, 0, 0, 0, 0, 0};
    int ph_n = 0;
#endif
    auto step = [&](int x, const bool live, bool &live_n, const Vox2c<T> &c0, const Vox2c<T> &c1, Vox2c<T> &p0, Vox2c<T> &p1) {
        SP_CLOCK(k0, c0.p);
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
        if (a.l2pf) tile_prefetch(pf, a, x + 2, (lw.w2 >> sh) & 0xffu);
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
            SP_CLOCK(k1, fs_mid + react0 + react1);
            __syncthreads();
            SP_CLOCK(k2, fs_mid);
            const Pair<T> pu = tl[0][o0 - RS], pd = tl[0][o1 + RS];                       // rows y-1 and y+2: (P, k)
            const Pair<T> su = tl[1][o0 - RS], sd = tl[1][o1 + RS];                       // (S, b)
#ifdef TGTK_SP_SHFL
            // experiment: the z neighbours of lanes 1..30 by warp shuffle, shared memory only at the two ends of the warp
            Pair<T> pl0 = {__shfl_up_sync(0xffffffffu, c0.p, 1), __shfl_up_sync(0xffffffffu, c0.k, 1)};
            Pair<T> pr0 = {__shfl_down_sync(0xffffffffu, c0.p, 1), __shfl_down_sync(0xffffffffu, c0.k, 1)};
            Pair<T> pl1 = {__shfl_up_sync(0xffffffffu, c1.p, 1), __shfl_up_sync(0xffffffffu, c1.k, 1)};
            Pair<T> pr1 = {__shfl_down_sync(0xffffffffu, c1.p, 1), __shfl_down_sync(0xffffffffu, c1.k, 1)};
            Pair<T> sl0 = {__shfl_up_sync(0xffffffffu, c0.s, 1), __shfl_up_sync(0xffffffffu, c0.b, 1)};
            Pair<T> sr0 = {__shfl_down_sync(0xffffffffu, c0.s, 1), __shfl_down_sync(0xffffffffu, c0.b, 1)};
            Pair<T> sl1 = {__shfl_up_sync(0xffffffffu, c1.s, 1), __shfl_up_sync(0xffffffffu, c1.b, 1)};
            Pair<T> sr1 = {__shfl_down_sync(0xffffffffu, c1.s, 1), __shfl_down_sync(0xffffffffu, c1.b, 1)};
            if (threadIdx.x == 0) { pl0 = tl[0][o0 - 1]; pl1 = tl[0][o1 - 1]; sl0 = tl[1][o0 - 1]; sl1 = tl[1][o1 - 1]; }
            if (threadIdx.x == 31) { pr0 = tl[0][o0 + 1]; pr1 = tl[0][o1 + 1]; sr0 = tl[1][o0 + 1]; sr1 = tl[1][o1 + 1]; }
#else
            const Pair<T> pl0 = tl[0][o0 - 1], pr0 = tl[0][o0 + 1], pl1 = tl[0][o1 - 1], pr1 = tl[0][o1 + 1];
            const Pair<T> sl0 = tl[1][o0 - 1], sr0 = tl[1][o0 + 1], sl1 = tl[1][o1 - 1], sr1 = tl[1][o1 + 1];
#endif
            T sp0 = a.hy * (relu2(pu.b + c0.k) * (pu.a - c0.p) - fp_mid);
            T sp1 = a.hy * (fp_mid - relu2(c1.k + pd.b) * (c1.p - pd.a));
            T ss0 = a.gy * (relu2(su.b + c0.b) * (su.a - c0.s) - fs_mid);
            T ss1 = a.gy * (fs_mid - relu2(c1.b + sd.b) * (c1.s - sd.a));
            sp0 += a.hz * (relu2(pl0.b + c0.k) * (pl0.a - c0.p) - relu2(c0.k + pr0.b) * (c0.p - pr0.a));
            sp1 += a.hz * (relu2(pl1.b + c1.k) * (pl1.a - c1.p) - relu2(c1.k + pr1.b) * (c1.p - pr1.a));
            ss0 += a.gz * (relu2(sl0.b + c0.b) * (sl0.a - c0.s) - relu2(c0.b + sr0.b) * (c0.s - sr0.a));
            ss1 += a.gz * (relu2(sl1.b + c1.b) * (sl1.a - c1.s) - relu2(c1.b + sr1.b) * (c1.s - sr1.a));
            SP_CLOCK(k3, sp0 + sp1 + ss0 + ss1);
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
            SP_CLOCK(k4, fp_hi0 + fp_hi1 + fs_hi0 + fs_hi1);
            SP_CLOCK(k5, sn0 + sn1 + nn0 + nn1);
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
#ifdef TGTK_SP_PHASES
            SP_CLOCK(k6, sP);
            ph[0] += k1 - k0; ph[1] += k2 - k1; ph[2] += k3 - k2; ph[3] += k4 - k3; ph[4] += k5 - k4; ph[5] += k6 - k5;
            ++ph_n;
#endif
        }
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
        lw.advance(x);
    };
    sparse_trace(sp, 1);
    const int n = g.x1 - g.x0;
    for (int i = 0; i < n; i += 2) {
        step(g.x0 + i, liveA, liveB, A0, A1, B0, B1);
        if (i + 1 >= n) break;
        step(g.x0 + i + 1, liveB, liveA, B0, B1, A0, A1);
    }
    sparse_trace(sp, 2);
#ifdef TGTK_SP_PHASES
    if (sp.trace && threadIdx.x == 0) {
        unsigned long long *q = sp.trace + 4 * (size_t)gridDim.x + ((size_t)blockIdx.x * 4 + threadIdx.y) * 8;
        for (int i = 0; i < 6; ++i) q[i] = (unsigned long long)ph[i];
        q[6] = (unsigned long long)ph_n;
    }
#endif
    step_end<T, 2>(a, sc, sP, sN);
}

// ---- FK_2c, own voxels staged through shared memory by cp.async ------------------------------------------------------
// fk2c_step_sp asks for plane x+1 at the top of step x, into registers, and needs it at the end of the same step: a step
// can never be shorter than one trip to memory, and the 20 registers of that plane are live through the whole in-plane
// section.  Here every thread copies ITS OWN ten values of plane x+D (D = 2) into its private shared-memory slots with
// cp.async -- no registers, no waiting -- and picks plane x+1 up from there just before the x faces need it
// (cp.async.wait_group: a thread only ever reads what it copied itself, so no barrier is involved).
__device__ __forceinline__ void cp_async_elem(double *dst_smem, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(float *dst_smem, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int MINB, int D>
__global__ void __launch_bounds__(128, MINB) fk2c_step_spa(const StepArgs<T> a, const SparseArgs sp)
{
    constexpr int TZ = 32, TYV = 8, RS = TZ + 2, PL = (TYV + 2) * RS;
    __shared__ Pair<T> tile[2][2][PL];     // [buffer][(P, k_eff), (S, b)][slot]
    __shared__ T stg[D][10][128];          // [plane % D][row * 5 + (P, N, S, b, k)][thread]
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
    sparse_trace(sp, 0);
    const int4 item = sp.work[2 * blockIdx.x], item_w = sp.work[2 * blockIdx.x + 1];
    sparse_prefetch_item<T>(a, sp, sc.par);
    Stage2 g = stage2_item<T, RS>(a, item);
    LiveWords lw;
    lw.col = sp.live + (size_t)item.y * sp.nyt + item.x;
    lw.stride = sp.nzt * sp.nyt;
    lw.X = a.X;
    lw.w0 = item_w.x; lw.w1 = item_w.y; lw.w2 = item_w.z; lw.w3 = item_w.w;
    const int sh = 8 * threadIdx.y, tid = threadIdx.y * TZ + threadIdx.x;
    const T closed = -Big<T>::v;
    auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
    auto load = [&](Vox2c<T> &v, uint32_t off) { v.p = P[off]; v.n = N[off]; v.s = S[off]; v.b = bc[off]; v.k = kc[off]; };
    auto none = [&](Vox2c<T> &v) { v.p = T(0); v.n = T(0); v.s = T(0); v.b = closed; v.k = closed; };
    // the copy of one plane's own voxels (element offsets o0, o1) into staging slot `slot`
    auto request = [&](int slot, uint32_t o0, uint32_t o1) {
        T(*q)[128] = stg[slot];
        cp_async_elem(&q[0][tid], P + o0); cp_async_elem(&q[1][tid], N + o0); cp_async_elem(&q[2][tid], S + o0);
        cp_async_elem(&q[3][tid], bc + o0); cp_async_elem(&q[4][tid], kc + o0);
        cp_async_elem(&q[5][tid], P + o1); cp_async_elem(&q[6][tid], N + o1); cp_async_elem(&q[7][tid], S + o1);
        cp_async_elem(&q[8][tid], bc + o1); cp_async_elem(&q[9][tid], kc + o1);
    };
    auto pick = [&](int slot, Vox2c<T> &v0, Vox2c<T> &v1) {
        T(*q)[128] = stg[slot];
        v0.p = q[0][tid]; v0.n = q[1][tid]; v0.s = q[2][tid]; v0.b = q[3][tid]; v0.k = q[4][tid];
        v1.p = q[5][tid]; v1.n = q[6][tid]; v1.s = q[7][tid]; v1.b = q[8][tid]; v1.k = q[9][tid];
    };
    double sP = 0.0, sN = 0.0;
    griddep_launch();
    griddep_wait();

    Vox2c<T> A0, A1, B0, B1;               // rows (y, y+1) of plane x (A) and plane x+1 (B); roles swap
    const bool live0 = (lw.w0 >> sh) & 0xffu;
    if (live0) { load(A0, g.c0); load(A1, g.c1); } else { none(A0); none(A1); }
    // planes x0+1 .. x0+D-1 into the staging ring; off1/off2: element offsets of the plane requested next
    uint32_t ahead0 = g.c0, ahead1 = g.c1;     // own voxels in the plane most recently requested
    auto advance = [&](int x_from) {          // offsets of plane x_from + 1 from those of plane x_from (periodic)
        const uint32_t adv = (x_from + 1 == a.X) ? 0u - g.wrap_back : g.sx;
        ahead0 += adv; ahead1 += adv;
    };
    {
        const uint32_t words[3] = {lw.w1, lw.w2, lw.w3};
#pragma unroll
        for (int d = 1; d < D; ++d) {
            int xq = g.x0 + d - 1;
            if (xq >= a.X) xq -= a.X;
            advance(xq);
            if ((words[d - 1] >> sh) & 0xffu) request(d % D, ahead0, ahead1);
            cp_async_commit();
        }
    }
    A0.k = keff(A0.p, A0.n, A0.k);
    A1.k = keff(A1.p, A1.n, A1.k);
    T fp_lo0 = T(0), fs_lo0 = T(0), fp_lo1 = T(0), fs_lo1 = T(0);      // fluxes through the faces below (x-1 | x)
    if (live0) {
        const T pm0 = P[g.prev0], pm1 = P[g.prev1];
        fp_lo0 = relu2(keff(pm0, N[g.prev0], kc[g.prev0]) + A0.k) * (pm0 - A0.p);
        fs_lo0 = relu2(bc[g.prev0] + A0.b) * (S[g.prev0] - A0.s);
        fp_lo1 = relu2(keff(pm1, N[g.prev1], kc[g.prev1]) + A1.k) * (pm1 - A1.p);
        fs_lo1 = relu2(bc[g.prev1] + A1.b) * (S[g.prev1] - A1.s);
    }
    // ring values two planes ahead as well: two register sets, the one published at the top of a step is refilled at once
    struct Ring { T p, n, k, s, b; };
    Ring RE = {0, 0, 0, 0, 0}, RO = {0, 0, 0, 0, 0};
    auto ring_load = [&](Ring &r) { r.p = P[g.h]; r.n = N[g.h]; r.k = kc[g.h]; r.s = S[g.h]; r.b = bc[g.h]; };
    if (g.ring) {
        ring_load(RE);
        g.h += (g.x0 + 1 == a.X) ? 0u - g.wrap_back : g.sx;
        ring_load(RO);
    }
    int buf = 0;
    const TilePrefetch pf = tile_prefetch_setup<T, 5>(a, item.x * TYV, item.y * TZ, P, N, S, kc, bc);

    auto step = [&](int i, int x, Ring &r, const Vox2c<T> &c0, const Vox2c<T> &c1, Vox2c<T> &p0, Vox2c<T> &p1) {
        Pair<T>(*tl)[PL] = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        const bool live = (lw.w0 >> sh) & 0xffu, live_n = (lw.w1 >> sh) & 0xffu;
        tl[0][o0] = Pair<T>{c0.p, c0.k};
        tl[1][o0] = Pair<T>{c0.s, c0.b};
        tl[0][o1] = Pair<T>{c1.p, c1.k};
        tl[1][o1] = Pair<T>{c1.s, c1.b};
        if (g.ring) {
            tl[0][g.hs] = Pair<T>{r.p, keff(r.p, r.n, r.k)};
            tl[1][g.hs] = Pair<T>{r.s, r.b};
            int xq = x + 1;                      // g.h: ring element in plane x+1 -> plane x+2
            if (xq >= a.X) xq -= a.X;
            g.h += (xq + 1 == a.X) ? 0u - g.wrap_back : g.sx;
            ring_load(r);
        }
        // own voxels of plane x+D -> staging slot (the one plane x was picked up from); ring of plane x+1 -> registers
        {
            int xq = x + D - 1;
            if (xq >= a.X) xq -= a.X;
            advance(xq);
            const uint32_t wd = (D == 1) ? lw.w1 : (D == 2) ? lw.w2 : lw.w3;
            if ((wd >> sh) & 0xffu) request(i % D, ahead0, ahead1);
            cp_async_commit();
        }
        const uint32_t adv = (x + 1 == a.X) ? 0u - g.wrap_back : g.sx;
        if (a.l2pf) {
            const uint32_t wf = (D == 1) ? lw.w2 : lw.w3;
            tile_prefetch(pf, a, x + D + 1, (wf >> sh) & 0xffu);
        }
        if (!live) {
            __syncthreads();
            cp_async_wait<D - 1>();
            if (live_n) pick((i + 1) % D, p0, p1); else { none(p0); none(p1); }
            p0.k = keff(p0.p, p0.n, p0.k);
            p1.k = keff(p1.p, p1.n, p1.k);
            fp_lo0 = fs_lo0 = fp_lo1 = fs_lo1 = T(0);      // the faces towards the next plane are closed from this side
        } else {
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
            // x faces last: plane x+1 from the staging ring
            cp_async_wait<D - 1>();
            if (live_n) pick((i + 1) % D, p0, p1); else { none(p0); none(p1); }
            p0.k = keff(p0.p, p0.n, p0.k);
            p1.k = keff(p1.p, p1.n, p1.k);
            const T fp_hi0 = relu2(c0.k + p0.k) * (c0.p - p0.p), fs_hi0 = relu2(c0.b + p0.b) * (c0.s - p0.s);
            const T fp_hi1 = relu2(c1.k + p1.k) * (c1.p - p1.p), fs_hi1 = relu2(c1.b + p1.b) * (c1.s - p1.s);
            sp0 += a.hx * (fp_lo0 - fp_hi0);
            sp1 += a.hx * (fp_lo1 - fp_hi1);
            ss0 += a.gx * (fs_lo0 - fs_hi0);
            ss1 += a.gx * (fs_lo1 - fs_hi1);
            const T pn0 = c0.p + a.dt * (sp0 + react0), pn1 = c1.p + a.dt * (sp1 + react1);
            if (!warp_has_p && __any_sync(0xffffffffu, (pn0 != T(0)) || (pn1 != T(0)))) {
                H0 = heaviside50_fast(c0.s, a.sigma_np); H1 = heaviside50_fast(c1.s, a.sigma_np);
            }
            const T nn0 = c0.n + a.dt * ((a.lambda_np * pn0) * H0), nn1 = c1.n + a.dt * ((a.lambda_np * pn1) * H1);
            const T sn0 = c0.s + a.dt * (ss0 - (a.lambda_s * c0.s) * pn0), sn1 = c1.s + a.dt * (ss1 - (a.lambda_s * c1.s) * pn1);
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
        g.c0 += adv; g.c1 += adv;
        buf ^= 1;
        lw.advance(x);
    };
    sparse_trace(sp, 1);
    const int n = g.x1 - g.x0;
    for (int i = 0; i < n; i += 2) {
        step(i, g.x0 + i, RE, A0, A1, B0, B1);
        if (i + 1 >= n) break;
        step(i + 1, g.x0 + i + 1, RO, B0, B1, A0, A1);
    }
    cp_async_wait<0>();
    sparse_trace(sp, 2);
    step_end<T, 2>(a, sc, sP, sN);
}

// ---- FK ----------------------------------------------------------------------------------------------------------------
template <typename T, int MINB, int SLAB = 0>
__global__ void __launch_bounds__(128, MINB) fk_step_sp(const StepArgs<T> a, const SparseArgs sp)
{
    constexpr int TZ = 32, TYV = 8, RS = TZ + 2, PL = (TYV + 2) * RS;
    __shared__ Pair<T> tile[2][PL];        // [buffer][slot] of (u, k)
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ kc = a.coef[0];
    sparse_trace(sp, 0);
    const int4 item = sp.work[2 * blockIdx.x], item_w = sp.work[2 * blockIdx.x + 1];
    Stage2 g = stage2_item<T, RS>(a, item);
    LiveWords lw;
    lw.col = sp.live + (size_t)item.y * sp.nyt + item.x;
    lw.stride = sp.nzt * sp.nyt;
    lw.X = a.X;
    lw.w0 = item_w.x; lw.w1 = item_w.y; lw.w2 = item_w.z; lw.w3 = item_w.w;
    const int sh = 8 * threadIdx.y;
    const T closed = -Big<T>::v;
    auto load = [&](VoxFk<T> &v, uint32_t off) { v.u = u[off]; v.k = kc[off]; };
    auto none = [&](VoxFk<T> &v) { v.u = T(0); v.k = closed; };
    double s = 0.0;
    griddep_launch();
    griddep_wait();
    const SlabCtl sl = SLAB ? slab_begin(a, g.x0, g.x1) : SlabCtl{false, false, 0};
    auto push = [&](T *const *peer, uint32_t off, T v0, T v1) {
        if (g.valid0) peer[0][g.c0 + off] = v0;
        if (g.valid1) peer[0][g.c1 + off] = v1;
    };

    VoxFk<T> A0, A1, B0, B1;
    bool liveA = (lw.w0 >> sh) & 0xffu, liveB = false;
    if (liveA) { load(A0, g.c0); load(A1, g.c1); } else { none(A0); none(A1); }
    T f_lo0 = T(0), f_lo1 = T(0);
    if (liveA) {
        f_lo0 = relu2(kc[g.prev0] + A0.k) * (u[g.prev0] - A0.u);
        f_lo1 = relu2(kc[g.prev1] + A1.k) * (u[g.prev1] - A1.u);
    }
    T hu = 0, hk = 0;
    if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
    int buf = 0;
    const TilePrefetch pf = tile_prefetch_setup<T, 2>(a, item.x * TYV, item.y * TZ, u, kc);

    auto step = [&](int x, const bool live, bool &live_n, const VoxFk<T> &c0, const VoxFk<T> &c1, VoxFk<T> &p0, VoxFk<T> &p1) {
        Pair<T> *tl = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        tl[o0] = Pair<T>{c0.u, c0.k};
        tl[o1] = Pair<T>{c1.u, c1.k};
        if (g.ring) tl[g.hs] = Pair<T>{hu, hk};
        const uint32_t adv = (x + 1 == a.X) ? 0u - g.wrap_back : g.sx;
        const uint32_t n0 = g.c0 + adv, n1 = g.c1 + adv;
        live_n = (lw.w1 >> sh) & 0xffu;
        if (live_n) { load(p0, n0); load(p1, n1); } else { none(p0); none(p1); }
        g.h += adv;
        if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
        if (a.l2pf) tile_prefetch(pf, a, x + 2, (lw.w2 >> sh) & 0xffu);
        if (!live) {
            __syncthreads();
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
            const T un0 = c0.u + a.dt * (sp0 + react0), un1 = c1.u + a.dt * (sp1 + react1);
            f_lo0 = f_hi0; f_lo1 = f_hi1;
            const bool cnt = counts(a, x);
            if (g.valid0) { out[g.c0] = un0; if (cnt) s += (double)un0; }
            if (g.valid1) { out[g.c1] = un1; if (cnt) s += (double)un1; }
            if (SLAB && sl.lo && x == a.x_lo) { push(a.peer_lo[sc.par ^ 1], a.off_lo, un0, un1); }
            if (SLAB && sl.hi && x == a.x_hi - 1) { push(a.peer_hi[sc.par ^ 1], a.off_hi, un0, un1); }
        }
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
        lw.advance(x);
    };
    sparse_trace(sp, 1);
    const int n = g.x1 - g.x0;
    for (int i = 0; i < n; i += 2) {
        step(g.x0 + i, liveA, liveB, A0, A1, B0, B1);
        if (i + 1 >= n) break;
        step(g.x0 + i + 1, liveB, liveA, B0, B1, A0, A1);
    }
    sparse_trace(sp, 2);
    step_end<T, 1>(a, sc, s, 0.0);
}

// ---- FK_DTI (reference rounding order, bitwise equal to dti_step_v4) -----------------------------------------------------
template <typename T, int MINB, int SLAB = 0>
__global__ void __launch_bounds__(128, MINB) dti_step_sp(const StepArgs<T> a, const SparseArgs sp)
{
    using R = Rn<T>;
    constexpr int TZ = 32, TYV = 8, RS = TZ + 2, PL = (TYV + 2) * RS;
    __shared__ T tile[2][PL];
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const T *__restrict__ u = a.buf[sc.par][0];
    T *__restrict__ out = a.buf[sc.par ^ 1][0];
    const T *__restrict__ dxa = a.coef[0];
    const T *__restrict__ dya = a.coef[1];
    const T *__restrict__ dza = a.coef[2];
    sparse_trace(sp, 0);
    const int4 item = sp.work[2 * blockIdx.x], item_w = sp.work[2 * blockIdx.x + 1];
    Stage2 g = stage2_item<T, RS>(a, item);
    LiveWords lw;
    lw.col = sp.live + (size_t)item.y * sp.nyt + item.x;
    lw.stride = sp.nzt * sp.nyt;
    lw.X = a.X;
    lw.w0 = item_w.x; lw.w1 = item_w.y; lw.w2 = item_w.z; lw.w3 = item_w.w;
    const int sh = 8 * threadIdx.y;
    double s = 0.0;
    auto load = [&](VoxDti<T> &v, uint32_t off) { v.u = u[off]; v.d0 = dxa[off]; v.d1 = dya[off]; v.d2 = dza[off]; };
    auto none = [&](VoxDti<T> &v) { v.u = T(0); v.d0 = T(0); v.d1 = T(0); v.d2 = T(0); };
    griddep_launch();
    griddep_wait();
    const SlabCtl sl = SLAB ? slab_begin(a, g.x0, g.x1) : SlabCtl{false, false, 0};
    auto push = [&](T *const *peer, uint32_t off, T v0, T v1) {
        if (g.valid0) peer[0][g.c0 + off] = v0;
        if (g.valid1) peer[0][g.c1 + off] = v1;
    };

    VoxDti<T> A0, A1, B0, B1;
    bool liveA = (lw.w0 >> sh) & 0xffu, liveB = false;
    if (liveA) { load(A0, g.c0); load(A1, g.c1); } else { none(A0); none(A1); }
    T um0 = T(0), um1 = T(0);
    if (liveA) { um0 = u[g.prev0]; um1 = u[g.prev1]; }
    T hu = 0;
    if (g.ring) hu = u[g.h];
    int buf = 0;
    const TilePrefetch pf = tile_prefetch_setup<T, 4>(a, item.x * TYV, item.y * TZ, u, dxa, dya, dza);

    auto one = [&](const VoxDti<T> &c, T um, T up, T uym, T uyp, T uzm, T uzp) {
        const T sx = R::mul(a.cx, R::sub(R::mul(c.d0, R::sub(um, c.u)), R::mul(c.d0, R::sub(c.u, up))));
        const T sy = R::mul(a.cy, R::sub(R::mul(c.d1, R::sub(uym, c.u)), R::mul(c.d1, R::sub(c.u, uyp))));
        const T sz = R::mul(a.cz, R::sub(R::mul(c.d2, R::sub(uzm, c.u)), R::mul(c.d2, R::sub(c.u, uzp))));
        const T spv = R::add(R::add(sx, sy), sz);
        const T react = R::mul(a.rho, R::mul(c.u, R::sub(T(1), c.u)));
        return R::add(c.u, R::mul(R::add(spv, react), a.dt));
    };
    auto step = [&](int x, const bool live, bool &live_n, const VoxDti<T> &c0, const VoxDti<T> &c1, VoxDti<T> &p0, VoxDti<T> &p1) {
        T *tl = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        tl[o0] = c0.u;
        tl[o1] = c1.u;
        if (g.ring) tl[g.hs] = hu;
        const uint32_t adv = (x + 1 == a.X) ? 0u - g.wrap_back : g.sx;
        const uint32_t n0 = g.c0 + adv, n1 = g.c1 + adv;
        live_n = (lw.w1 >> sh) & 0xffu;
        // the plane ahead: a live tile needs its u even where the tile ahead is dead -- but there u = 0 by definition
        if (live_n) { load(p0, n0); load(p1, n1); } else { none(p0); none(p1); }
        g.h += adv;
        if (g.ring) hu = u[g.h];
        if (a.l2pf) tile_prefetch(pf, a, x + 2, (lw.w2 >> sh) & 0xffu);
        __syncthreads();
        if (live) {
            const T un0 = one(c0, um0, p0.u, tl[o0 - RS], c1.u, tl[o0 - 1], tl[o0 + 1]);
            const T un1 = one(c1, um1, p1.u, c0.u, tl[o1 + RS], tl[o1 - 1], tl[o1 + 1]);
            const bool cnt = counts(a, x);
            if (g.valid0) { out[g.c0] = un0; if (cnt) s += (double)un0; }
            if (g.valid1) { out[g.c1] = un1; if (cnt) s += (double)un1; }
            if (SLAB && sl.lo && x == a.x_lo) { push(a.peer_lo[sc.par ^ 1], a.off_lo, un0, un1); }
            if (SLAB && sl.hi && x == a.x_hi - 1) { push(a.peer_hi[sc.par ^ 1], a.off_hi, un0, un1); }
        }
        um0 = c0.u; um1 = c1.u;
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
        lw.advance(x);
    };
    sparse_trace(sp, 1);
    const int n = g.x1 - g.x0;
    for (int i = 0; i < n; i += 2) {
        step(g.x0 + i, liveA, liveB, A0, A1, B0, B1);
        if (i + 1 >= n) break;
        step(g.x0 + i + 1, liveB, liveA, B0, B1, A0, A1);
    }
    sparse_trace(sp, 2);
    step_end<T, 1>(a, sc, s, 0.0);
}

}  // namespace tgtk
