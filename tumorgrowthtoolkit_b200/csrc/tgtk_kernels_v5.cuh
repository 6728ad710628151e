// fp32 marching kernels with four voxels per thread: a z pair (one 64-bit global access, one
// FADD2 / FMUL2 / FFMA2 per arithmetic step -- Blackwell's packed fp32 pipe) times the two rows
// (y, y+1) of tgtk_kernels_v3.cuh's two-voxel kernels.  Same staging as *_step_v4: own columns of
// plane x+1 requested one plane ahead into a second register set, in-plane neighbours through a
// double-buffered shared-memory copy of the plane, one barrier per plane.
//
// Tile: 2*TZL (z) x 2*TYT (y) voxels on TZL x TYT threads (32 x 16 on 16 x 8).  Per z pair the
// shared-memory copy holds one float4 (f_a, f_b, g_a, g_b) -- field f and its coefficient g of the
// voxels z and z+1 -- so the rows above / below arrive as register pairs ready for packed
// arithmetic (LDS.128) and the left / right neighbours are single floats of the adjacent float4.
//
// Layout contract (tgtk_api.cu): fp32 arrays have an even row pitch (sy = Z rounded up to even), so
// every (row, even z) pair is 8-byte aligned.  When Z is odd the pad element [.., y, Z] MIRRORS
// [.., y, 0]: the pair that starts at z = Z-1 then carries its periodic right neighbour (np.roll)
// in its second lane.  The thread that owns z = 0 keeps the mirror of the state arrays up to date;
// the static arrays are mirrored once by tgtk_sim_prepare (mirror_pad_kernel).
#pragma once
#include "tgtk_kernels_v3.cuh"

namespace tgtk {

// ---- packed pair arithmetic ---------------------------------------------------------------------
__device__ __forceinline__ uint64_t pk2(float2 v)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
    return r;
}
__device__ __forceinline__ float2 upk2(uint64_t r)
{
    float2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
    return upk2(r);
}
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 relu2(float2 a) { return make_float2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)); }
__device__ __forceinline__ float2 ld2(const float *p, uint32_t off) { return *reinterpret_cast<const float2 *>(p + off); }

// smooth_heaviside (FK_2c.py:75-76) of a pair: 1 - 1/(1 + 2^t), t = -50 log2(e) (s - sigma), with the
// SFU's ex2 / rcp (relative error 2^-22: far inside the fp32 mode's 1e-4 contract).  t is capped so
// that 2^t stays finite; for large t the result is 1 - 0 like the reference's 1 - 1/inf.
__device__ __forceinline__ float2 heaviside50_pair(float2 s, float sigma)
{
    const float c = 72.13475204444817f;   // 50 log2(e)
    float2 t = fma2(s, bc2(-c), bc2(c * sigma));
    t.x = fminf(t.x, 115.0f);
    t.y = fminf(t.y, 115.0f);
    float2 e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
    const float2 d = add2(e, bc2(1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
    return fma2(r, bc2(-1.0f), bc2(1.0f));
}

// Left / right neighbours of a z pair: the voxel left of my a-element is the b-element of the lane below, the voxel
// right of my b-element the a-element of the lane above -- one SHFL each instead of a scalar shared-memory read at a
// 16-byte lane stride (4-way bank conflict, what made round 1's z-pair FK / FK_2c kernels lose to the two-voxel ones).
// Only the first / last lane of a tile row (TZL lanes; a warp holds 32 / TZL rows) reads the ring slot beside the tile.
template <int TZL> __device__ __forceinline__ float from_left(float mine_b) { return __shfl_up_sync(0xffffffffu, mine_b, 1, TZL); }
template <int TZL> __device__ __forceinline__ float from_right(float mine_a) { return __shfl_down_sync(0xffffffffu, mine_a, 1, TZL); }

// ---- geometry -------------------------------------------------------------------------------------
struct Stage4 {
    int x0, x1;
    bool va0, vb0, va1, vb1;   // voxels (z, y), (z+1, y), (z, y+1), (z+1, y+1) lie inside the box
    bool ring, mirror;         // loads one ring element / owns z = 0 of a box with odd Z
    uint32_t c0, c1, prev0, prev1, h;
    uint32_t sx, wrap_back, zpad;
    int so;                    // slot (in pair units) of the upper own row; the lower one is so + RS
    int hs;                    // ring element: float index of its field value (coefficient: + 2 in float4 tiles)
};

// UNIT = floats per pair slot of the shared tile (4: field + coefficient, 2: field only);
// ROWS = voxel rows per thread (2: *_step_v5, 1: *_step_v6, where c1 / va1 / vb1 are unused)
template <int TZL, int TYT, int UNIT, int ROWS = 2>
__device__ __forceinline__ Stage4 stage4_setup(const StepArgs<float> &a, int lx)
{
    constexpr int TZV = 2 * TZL, TYV = ROWS * TYT, RS = TZL + 2;
    Stage4 g;
    const int tz = threadIdx.x, ty = threadIdx.y, tid = ty * TZL + tz;
    const int z0 = blockIdx.x * TZV, y0 = blockIdx.y * TYV;
    const int z = z0 + 2 * tz, r0 = y0 + ROWS * ty;
    const bool ra = r0 < a.Y, rb = r0 + 1 < a.Y;
    g.va0 = ra && z < a.Z;
    g.vb0 = ra && z + 1 < a.Z;
    g.va1 = rb && z < a.Z;
    g.vb1 = rb && z + 1 < a.Z;
    g.x0 = blockIdx.z * lx;
    g.x1 = min(g.x0 + lx, a.X);
    g.sx = (uint32_t)a.sx;
    g.wrap_back = (uint32_t)(a.X - 1) * g.sx;
    g.zpad = (uint32_t)a.Z;
    g.mirror = (a.Z & 1) && z == 0;
    const uint32_t sy = (uint32_t)a.sy, base = (uint32_t)g.x0 * g.sx;
    // inside the box: the pair at z (odd Z: the pair at Z-1 ends in the mirror of z = 0).  Outside: a
    // shadow pair; the one at z == Z (even Z) must supply z = 0, the others are never looked at.
    const uint32_t zc = (z < a.Z) ? (uint32_t)z : ((uint32_t)wrap_coord(z, a.Z) & ~1u);
    g.c0 = base + (uint32_t)wrap_coord(r0, a.Y) * sy + zc;
    g.c1 = base + (uint32_t)wrap_coord(r0 + 1, a.Y) * sy + zc;
    g.prev0 = (g.x0 == 0) ? g.c0 + g.wrap_back : g.c0 - g.sx;
    g.prev1 = (g.x0 == 0) ? g.c1 + g.wrap_back : g.c1 - g.sx;
    g.so = (ROWS * ty + 1) * RS + tz + 1;
    // ring: 2*TZL elements above, as many below, 2*TYT on the left, as many on the right (scalars)
    g.ring = tid < 2 * TZV + 2 * TYV;
    int row = 1, unit = 1, lane = 0, yr = y0, zr = z0;
    if (tid < TZV) { row = 0; unit = (tid >> 1) + 1; lane = tid & 1; yr = y0 - 1; zr = z0 + tid; }
    else if (tid < 2 * TZV) { const int i = tid - TZV; row = TYV + 1; unit = (i >> 1) + 1; lane = i & 1; yr = y0 + TYV; zr = z0 + i; }
    else if (tid < 2 * TZV + TYV) { const int i = tid - 2 * TZV; row = i + 1; unit = 0; lane = 1; yr = y0 + i; zr = z0 - 1; }
    else if (g.ring) { const int i = tid - 2 * TZV - TYV; row = i + 1; unit = TZL + 1; lane = 0; yr = y0 + i; zr = z0 + TZV; }
    g.hs = (row * RS + unit) * UNIT + lane;
    g.h = base + (uint32_t)wrap_coord(yr, a.Y) * sy + (uint32_t)wrap_coord(zr, a.Z);
    return g;
}

// 64-bit store of a pair, or of its first lane only where the second lies outside the box
__device__ __forceinline__ void store_pair(float *q, uint32_t off, float2 v, bool va, bool vb)
{
    if (vb) *reinterpret_cast<float2 *>(q + off) = v;
    else if (va) q[off] = v.x;
}

// ---- FK_2c ------------------------------------------------------------------------------------------
struct Vox2c4 { float2 p, n, s, b, k; };

template <int TZL, int TYT, int MINB>
__global__ void __launch_bounds__(TZL *TYT, MINB) fk2c_step_v5(const StepArgs<float> a, const int lx)
{
    constexpr int TYV = 2 * TYT, RS = TZL + 2, PL = (TYV + 2) * RS;
    __shared__ float4 tile[2][2][PL];      // [buffer][(P_a, P_b, k_a, k_b) | (S_a, S_b, b_a, b_b)][slot]
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const float *__restrict__ P = a.buf[sc.par][0];
    const float *__restrict__ N = a.buf[sc.par][1];
    const float *__restrict__ S = a.buf[sc.par][2];
    float *__restrict__ Po = a.buf[sc.par ^ 1][0];
    float *__restrict__ No = a.buf[sc.par ^ 1][1];
    float *__restrict__ So = a.buf[sc.par ^ 1][2];
    const float *__restrict__ kc = a.coef[0];
    const float *__restrict__ bc = a.coef[1];
    Stage4 g = stage4_setup<TZL, TYT, 4>(a, lx);
    const float closed = -Big<float>::v;
    const float th = a.th_necro;
    auto keff1 = [&](float pq, float nq, float kq) { return (pq + nq <= th) ? kq : closed; };
    auto keff2 = [&](float2 pq, float2 nq, float2 kq) {
        const float2 t = add2(pq, nq);
        return make_float2((t.x <= th) ? kq.x : closed, (t.y <= th) ? kq.y : closed);
    };
    auto load = [&](Vox2c4 &v, uint32_t off) {
        v.p = ld2(P, off); v.n = ld2(N, off); v.s = ld2(S, off); v.b = ld2(bc, off); v.k = ld2(kc, off);
    };
    const float2 cx2 = bc2(a.cx), cy2 = bc2(a.cy), cz2 = bc2(a.cz), ex2 = bc2(a.ex), ey2 = bc2(a.ey), ez2 = bc2(a.ez);
    const float2 dt2 = bc2(a.dt), rho2 = bc2(a.rho), lam2 = bc2(a.lambda_np), lams2 = bc2(a.lambda_s), one2 = bc2(1.0f);
    double sP = 0.0, sN = 0.0;
    griddep_launch();
    griddep_wait();

    Vox2c4 A0, A1, B0, B1;                 // rows (y, y+1) of plane x (A) and plane x+1 (B); roles swap
    load(A0, g.c0); A0.k = keff2(A0.p, A0.n, A0.k);
    load(A1, g.c1); A1.k = keff2(A1.p, A1.n, A1.k);
    float2 fp_lo0, fs_lo0, fp_lo1, fs_lo1; // fluxes through the faces below (x-1 | x)
    {
        const float2 pm0 = ld2(P, g.prev0), pm1 = ld2(P, g.prev1);
        fp_lo0 = mul2(relu2(add2(keff2(pm0, ld2(N, g.prev0), ld2(kc, g.prev0)), A0.k)), sub2(pm0, A0.p));
        fs_lo0 = mul2(relu2(add2(ld2(bc, g.prev0), A0.b)), sub2(ld2(S, g.prev0), A0.s));
        fp_lo1 = mul2(relu2(add2(keff2(pm1, ld2(N, g.prev1), ld2(kc, g.prev1)), A1.k)), sub2(pm1, A1.p));
        fs_lo1 = mul2(relu2(add2(ld2(bc, g.prev1), A1.b)), sub2(ld2(S, g.prev1), A1.s));
    }
    float hp = 0, hn = 0, hk = 0, hsv = 0, hb = 0;
    if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto step = [&](int x, const Vox2c4 &c0, const Vox2c4 &c1, Vox2c4 &p0, Vox2c4 &p1) {
        float4 *t0 = tile[buf][0], *t1 = tile[buf][1];
        const int o0 = g.so, o1 = g.so + RS;
        t0[o0] = make_float4(c0.p.x, c0.p.y, c0.k.x, c0.k.y);
        t1[o0] = make_float4(c0.s.x, c0.s.y, c0.b.x, c0.b.y);
        t0[o1] = make_float4(c1.p.x, c1.p.y, c1.k.x, c1.k.y);
        t1[o1] = make_float4(c1.s.x, c1.s.y, c1.b.x, c1.b.y);
        if (g.ring) {
            float *f0 = reinterpret_cast<float *>(t0), *f1 = reinterpret_cast<float *>(t1);
            f0[g.hs] = hp; f0[g.hs + 2] = keff1(hp, hn, hk);
            f1[g.hs] = hsv; f1[g.hs + 2] = hb;
        }
        // request plane x+1: both own pairs and the ring
        const uint32_t n0 = (x + 1 == a.X) ? g.c0 - g.wrap_back : g.c0 + g.sx;
        const uint32_t n1 = (x + 1 == a.X) ? g.c1 - g.wrap_back : g.c1 + g.sx;
        load(p0, n0);
        load(p1, n1);
        g.h = (x + 1 == a.X) ? g.h - g.wrap_back : g.h + g.sx;
        if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x);
            l2_prefetch_strip(P, q, pf.count, pf.total); l2_prefetch_strip(N, q, pf.count, pf.total);
            l2_prefetch_strip(S, q, pf.count, pf.total); l2_prefetch_strip(kc, q, pf.count, pf.total);
            l2_prefetch_strip(bc, q, pf.count, pf.total);
        }
        // work that needs neither neighbours nor plane x+1
        const float2 H0 = heaviside50_pair(c0.s, a.sigma_np), H1 = heaviside50_pair(c1.s, a.sigma_np);
        const float2 react0 = sub2(mul2(mul2(rho2, mul2(c0.s, c0.p)), sub2(sub2(one2, c0.p), c0.n)), mul2(mul2(lam2, c0.p), H0));
        const float2 react1 = sub2(mul2(mul2(rho2, mul2(c1.s, c1.p)), sub2(sub2(one2, c1.p), c1.n)), mul2(mul2(lam2, c1.p), H1));
        // faces between the four own voxels: y between the rows (pairs), z inside each pair (scalars)
        const float2 fp_mid = mul2(relu2(add2(c0.k, c1.k)), sub2(c0.p, c1.p));
        const float2 fs_mid = mul2(relu2(add2(c0.b, c1.b)), sub2(c0.s, c1.s));
        const float gp0 = relu(c0.k.x + c0.k.y) * (c0.p.x - c0.p.y), gs0 = relu(c0.b.x + c0.b.y) * (c0.s.x - c0.s.y);
        const float gp1 = relu(c1.k.x + c1.k.y) * (c1.p.x - c1.p.y), gs1 = relu(c1.b.x + c1.b.y) * (c1.s.x - c1.s.y);
        __syncthreads();
        const float4 pu = t0[o0 - RS], pd = t0[o1 + RS];       // rows y-1 and y+2: (P_a, P_b, k_a, k_b)
        const float4 su = t1[o0 - RS], sd = t1[o1 + RS];       // (S_a, S_b, b_a, b_b)
        float pl0 = from_left<TZL>(c0.p.y), kl0 = from_left<TZL>(c0.k.y), sl0 = from_left<TZL>(c0.s.y), bl0 = from_left<TZL>(c0.b.y);
        float pr0 = from_right<TZL>(c0.p.x), kr0 = from_right<TZL>(c0.k.x), sr0 = from_right<TZL>(c0.s.x), br0 = from_right<TZL>(c0.b.x);
        float pl1 = from_left<TZL>(c1.p.y), kl1 = from_left<TZL>(c1.k.y), sl1 = from_left<TZL>(c1.s.y), bl1 = from_left<TZL>(c1.b.y);
        float pr1 = from_right<TZL>(c1.p.x), kr1 = from_right<TZL>(c1.k.x), sr1 = from_right<TZL>(c1.s.x), br1 = from_right<TZL>(c1.b.x);
        if (threadIdx.x == 0) {
            const float4 e0 = t0[o0 - 1], e1 = t1[o0 - 1], e2 = t0[o1 - 1], e3 = t1[o1 - 1];
            pl0 = e0.y; kl0 = e0.w; sl0 = e1.y; bl0 = e1.w; pl1 = e2.y; kl1 = e2.w; sl1 = e3.y; bl1 = e3.w;
        }
        if (threadIdx.x == TZL - 1) {
            const float4 e0 = t0[o0 + 1], e1 = t1[o0 + 1], e2 = t0[o1 + 1], e3 = t1[o1 + 1];
            pr0 = e0.x; kr0 = e0.z; sr0 = e1.x; br0 = e1.z; pr1 = e2.x; kr1 = e2.z; sr1 = e3.x; br1 = e3.z;
        }
        // y
        float2 sp0 = mul2(cy2, sub2(mul2(relu2(add2(make_float2(pu.z, pu.w), c0.k)), sub2(make_float2(pu.x, pu.y), c0.p)), fp_mid));
        float2 sp1 = mul2(cy2, sub2(fp_mid, mul2(relu2(add2(c1.k, make_float2(pd.z, pd.w))), sub2(c1.p, make_float2(pd.x, pd.y)))));
        float2 ss0 = mul2(ey2, sub2(mul2(relu2(add2(make_float2(su.z, su.w), c0.b)), sub2(make_float2(su.x, su.y), c0.s)), fs_mid));
        float2 ss1 = mul2(ey2, sub2(fs_mid, mul2(relu2(add2(c1.b, make_float2(sd.z, sd.w))), sub2(c1.s, make_float2(sd.x, sd.y)))));
        // z: left | a | b | right
        {
            const float fl0 = relu(kl0 + c0.k.x) * (pl0 - c0.p.x), fr0 = relu(c0.k.y + kr0) * (c0.p.y - pr0);
            const float fl1 = relu(kl1 + c1.k.x) * (pl1 - c1.p.x), fr1 = relu(c1.k.y + kr1) * (c1.p.y - pr1);
            sp0 = fma2(cz2, make_float2(fl0 - gp0, gp0 - fr0), sp0);
            sp1 = fma2(cz2, make_float2(fl1 - gp1, gp1 - fr1), sp1);
            const float hl0 = relu(bl0 + c0.b.x) * (sl0 - c0.s.x), hr0 = relu(c0.b.y + br0) * (c0.s.y - sr0);
            const float hl1 = relu(bl1 + c1.b.x) * (sl1 - c1.s.x), hr1 = relu(c1.b.y + br1) * (c1.s.y - sr1);
            ss0 = fma2(ez2, make_float2(hl0 - gs0, gs0 - hr0), ss0);
            ss1 = fma2(ez2, make_float2(hl1 - gs1, gs1 - hr1), ss1);
        }
        // x faces last: plane x+1 has had the whole step to arrive
        p0.k = keff2(p0.p, p0.n, p0.k);
        p1.k = keff2(p1.p, p1.n, p1.k);
        const float2 fp_hi0 = mul2(relu2(add2(c0.k, p0.k)), sub2(c0.p, p0.p)), fs_hi0 = mul2(relu2(add2(c0.b, p0.b)), sub2(c0.s, p0.s));
        const float2 fp_hi1 = mul2(relu2(add2(c1.k, p1.k)), sub2(c1.p, p1.p)), fs_hi1 = mul2(relu2(add2(c1.b, p1.b)), sub2(c1.s, p1.s));
        sp0 = fma2(cx2, sub2(fp_lo0, fp_hi0), sp0);
        sp1 = fma2(cx2, sub2(fp_lo1, fp_hi1), sp1);
        ss0 = fma2(ex2, sub2(fs_lo0, fs_hi0), ss0);
        ss1 = fma2(ex2, sub2(fs_lo1, fs_hi1), ss1);
        // FK_2c.py:54-71: P, then N with the new P, then S with the new P
        const float2 pn0 = fma2(dt2, add2(sp0, react0), c0.p), pn1 = fma2(dt2, add2(sp1, react1), c1.p);
        const float2 nn0 = fma2(dt2, mul2(mul2(lam2, pn0), H0), c0.n), nn1 = fma2(dt2, mul2(mul2(lam2, pn1), H1), c1.n);
        const float2 sn0 = fma2(dt2, sub2(ss0, mul2(mul2(lams2, c0.s), pn0)), c0.s);
        const float2 sn1 = fma2(dt2, sub2(ss1, mul2(mul2(lams2, c1.s), pn1)), c1.s);
        store_pair(Po, g.c0, pn0, g.va0, g.vb0); store_pair(No, g.c0, nn0, g.va0, g.vb0); store_pair(So, g.c0, sn0, g.va0, g.vb0);
        store_pair(Po, g.c1, pn1, g.va1, g.vb1); store_pair(No, g.c1, nn1, g.va1, g.vb1); store_pair(So, g.c1, sn1, g.va1, g.vb1);
        if (g.mirror) {                        // odd Z: the pad element of the row mirrors z = 0
            if (g.va0) { Po[g.c0 + g.zpad] = pn0.x; No[g.c0 + g.zpad] = nn0.x; So[g.c0 + g.zpad] = sn0.x; }
            if (g.va1) { Po[g.c1 + g.zpad] = pn1.x; No[g.c1 + g.zpad] = nn1.x; So[g.c1 + g.zpad] = sn1.x; }
        }
        if (counts(a, x)) {
            if (g.va0) { sP += (double)pn0.x; sN += (double)nn0.x; }
            if (g.vb0) { sP += (double)pn0.y; sN += (double)nn0.y; }
            if (g.va1) { sP += (double)pn1.x; sN += (double)nn1.x; }
            if (g.vb1) { sP += (double)pn1.y; sN += (double)nn1.y; }
        }
        fp_lo0 = fp_hi0; fs_lo0 = fs_hi0; fp_lo1 = fp_hi1; fs_lo1 = fs_hi1;
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 2) {
        step(x, A0, A1, B0, B1);
        if (x + 1 >= g.x1) break;
        step(x + 1, B0, B1, A0, A1);
    }
    step_end<float, 2>(a, sc, sP, sN);
}

// ---- FK ---------------------------------------------------------------------------------------------
struct VoxFk4 { float2 u, k; };

template <int TZL, int TYT, int MINB>
__global__ void __launch_bounds__(TZL *TYT, MINB) fk_step_v5(const StepArgs<float> a, const int lx)
{
    constexpr int TYV = 2 * TYT, RS = TZL + 2, PL = (TYV + 2) * RS;
    __shared__ float4 tile[2][PL];         // [buffer][slot] of (u_a, u_b, k_a, k_b)
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const float *__restrict__ u = a.buf[sc.par][0];
    float *__restrict__ out = a.buf[sc.par ^ 1][0];
    const float *__restrict__ kc = a.coef[0];
    Stage4 g = stage4_setup<TZL, TYT, 4>(a, lx);
    const float2 cx2 = bc2(a.cx), cy2 = bc2(a.cy), cz2 = bc2(a.cz), dt2 = bc2(a.dt), rho2 = bc2(a.rho), one2 = bc2(1.0f);
    double s = 0.0;
    griddep_launch();
    griddep_wait();

    VoxFk4 A0, A1, B0, B1;
    A0.u = ld2(u, g.c0); A0.k = ld2(kc, g.c0);
    A1.u = ld2(u, g.c1); A1.k = ld2(kc, g.c1);
    float2 f_lo0 = mul2(relu2(add2(ld2(kc, g.prev0), A0.k)), sub2(ld2(u, g.prev0), A0.u));
    float2 f_lo1 = mul2(relu2(add2(ld2(kc, g.prev1), A1.k)), sub2(ld2(u, g.prev1), A1.u));
    float hu = 0, hk = 0;
    if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto step = [&](int x, const VoxFk4 &c0, const VoxFk4 &c1, VoxFk4 &p0, VoxFk4 &p1) {
        float4 *tl = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        tl[o0] = make_float4(c0.u.x, c0.u.y, c0.k.x, c0.k.y);
        tl[o1] = make_float4(c1.u.x, c1.u.y, c1.k.x, c1.k.y);
        if (g.ring) {
            float *f = reinterpret_cast<float *>(tl);
            f[g.hs] = hu; f[g.hs + 2] = hk;
        }
        const uint32_t n0 = (x + 1 == a.X) ? g.c0 - g.wrap_back : g.c0 + g.sx;
        const uint32_t n1 = (x + 1 == a.X) ? g.c1 - g.wrap_back : g.c1 + g.sx;
        p0.u = ld2(u, n0); p0.k = ld2(kc, n0);
        p1.u = ld2(u, n1); p1.k = ld2(kc, n1);
        g.h = (x + 1 == a.X) ? g.h - g.wrap_back : g.h + g.sx;
        if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x);
            l2_prefetch_strip(u, q, pf.count, pf.total); l2_prefetch_strip(kc, q, pf.count, pf.total);
        }
        const float2 react0 = mul2(mul2(rho2, c0.u), sub2(one2, c0.u)), react1 = mul2(mul2(rho2, c1.u), sub2(one2, c1.u));
        const float2 f_mid = mul2(relu2(add2(c0.k, c1.k)), sub2(c0.u, c1.u));
        const float g0 = relu(c0.k.x + c0.k.y) * (c0.u.x - c0.u.y), g1 = relu(c1.k.x + c1.k.y) * (c1.u.x - c1.u.y);
        __syncthreads();
        const float4 up = tl[o0 - RS], dn = tl[o1 + RS];
        float ul0 = from_left<TZL>(c0.u.y), kl0 = from_left<TZL>(c0.k.y), ur0 = from_right<TZL>(c0.u.x), kr0 = from_right<TZL>(c0.k.x);
        float ul1 = from_left<TZL>(c1.u.y), kl1 = from_left<TZL>(c1.k.y), ur1 = from_right<TZL>(c1.u.x), kr1 = from_right<TZL>(c1.k.x);
        if (threadIdx.x == 0) { const float4 e0 = tl[o0 - 1], e1 = tl[o1 - 1]; ul0 = e0.y; kl0 = e0.w; ul1 = e1.y; kl1 = e1.w; }
        if (threadIdx.x == TZL - 1) { const float4 e0 = tl[o0 + 1], e1 = tl[o1 + 1]; ur0 = e0.x; kr0 = e0.z; ur1 = e1.x; kr1 = e1.z; }
        float2 sp0 = mul2(cy2, sub2(mul2(relu2(add2(make_float2(up.z, up.w), c0.k)), sub2(make_float2(up.x, up.y), c0.u)), f_mid));
        float2 sp1 = mul2(cy2, sub2(f_mid, mul2(relu2(add2(c1.k, make_float2(dn.z, dn.w))), sub2(c1.u, make_float2(dn.x, dn.y)))));
        const float fl0 = relu(kl0 + c0.k.x) * (ul0 - c0.u.x), fr0 = relu(c0.k.y + kr0) * (c0.u.y - ur0);
        const float fl1 = relu(kl1 + c1.k.x) * (ul1 - c1.u.x), fr1 = relu(c1.k.y + kr1) * (c1.u.y - ur1);
        sp0 = fma2(cz2, make_float2(fl0 - g0, g0 - fr0), sp0);
        sp1 = fma2(cz2, make_float2(fl1 - g1, g1 - fr1), sp1);
        const float2 f_hi0 = mul2(relu2(add2(c0.k, p0.k)), sub2(c0.u, p0.u)), f_hi1 = mul2(relu2(add2(c1.k, p1.k)), sub2(c1.u, p1.u));
        sp0 = fma2(cx2, sub2(f_lo0, f_hi0), sp0);
        sp1 = fma2(cx2, sub2(f_lo1, f_hi1), sp1);
        const float2 un0 = fma2(dt2, add2(sp0, react0), c0.u), un1 = fma2(dt2, add2(sp1, react1), c1.u);
        store_pair(out, g.c0, un0, g.va0, g.vb0);
        store_pair(out, g.c1, un1, g.va1, g.vb1);
        if (g.mirror) {
            if (g.va0) out[g.c0 + g.zpad] = un0.x;
            if (g.va1) out[g.c1 + g.zpad] = un1.x;
        }
        if (counts(a, x)) {
            if (g.va0) s += (double)un0.x;
            if (g.vb0) s += (double)un0.y;
            if (g.va1) s += (double)un1.x;
            if (g.vb1) s += (double)un1.y;
        }
        f_lo0 = f_hi0; f_lo1 = f_hi1;
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 2) {
        step(x, A0, A1, B0, B1);
        if (x + 1 >= g.x1) break;
        step(x + 1, B0, B1, A0, A1);
    }
    step_end<float, 1>(a, sc, s, 0.0);
}

// ---- FK_DTI (every operation individually rounded as in dti_step_v4: bitwise equal to it) -----------
struct VoxDti4 { float2 u, d0, d1, d2; };

template <int TZL, int TYT, int MINB>
__global__ void __launch_bounds__(TZL *TYT, MINB) dti_step_v5(const StepArgs<float> a, const int lx)
{
    using R = Rn<float>;
    constexpr int TYV = 2 * TYT, RS = TZL + 2, PL = (TYV + 2) * RS;
    __shared__ float2 tile[2][PL];         // [buffer][slot] of (u_a, u_b)
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const float *__restrict__ u = a.buf[sc.par][0];
    float *__restrict__ out = a.buf[sc.par ^ 1][0];
    const float *__restrict__ dxa = a.coef[0];
    const float *__restrict__ dya = a.coef[1];
    const float *__restrict__ dza = a.coef[2];
    Stage4 g = stage4_setup<TZL, TYT, 2>(a, lx);
    const float2 cx2 = bc2(a.cx), cy2 = bc2(a.cy), dt2 = bc2(a.dt), rho2 = bc2(a.rho), one2 = bc2(1.0f);
    double s = 0.0;
    auto load = [&](VoxDti4 &v, uint32_t off) { v.u = ld2(u, off); v.d0 = ld2(dxa, off); v.d1 = ld2(dya, off); v.d2 = ld2(dza, off); };
    griddep_launch();
    griddep_wait();

    VoxDti4 A0, A1, B0, B1;
    load(A0, g.c0);
    load(A1, g.c1);
    float2 um0 = ld2(u, g.prev0), um1 = ld2(u, g.prev1);
    float hu = 0;
    if (g.ring) hu = u[g.h];
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    // FK_DTI.py:39-45 + FK.py:36-41 with D_plus aliasing D_minus: D (u(c-a) - u(c)) - D (u(c) - u(c+a))
    auto one = [&](const VoxDti4 &c, float2 um, float2 up, float2 uym, float2 uyp, float uzl, float uzr) {
        const float2 sx = mul2(cx2, sub2(mul2(c.d0, sub2(um, c.u)), mul2(c.d0, sub2(c.u, up))));
        const float2 sy = mul2(cy2, sub2(mul2(c.d1, sub2(uym, c.u)), mul2(c.d1, sub2(c.u, uyp))));
        float2 sz;
        sz.x = R::mul(a.cz, R::sub(R::mul(c.d2.x, R::sub(uzl, c.u.x)), R::mul(c.d2.x, R::sub(c.u.x, c.u.y))));
        sz.y = R::mul(a.cz, R::sub(R::mul(c.d2.y, R::sub(c.u.x, c.u.y)), R::mul(c.d2.y, R::sub(c.u.y, uzr))));
        const float2 sp = add2(add2(sx, sy), sz);
        const float2 react = mul2(rho2, mul2(c.u, sub2(one2, c.u)));
        return add2(c.u, mul2(add2(sp, react), dt2));
    };
    auto step = [&](int x, const VoxDti4 &c0, const VoxDti4 &c1, VoxDti4 &p0, VoxDti4 &p1) {
        float2 *tl = tile[buf];
        const int o0 = g.so, o1 = g.so + RS;
        tl[o0] = c0.u;
        tl[o1] = c1.u;
        if (g.ring) reinterpret_cast<float *>(tl)[g.hs] = hu;
        const uint32_t n0 = (x + 1 == a.X) ? g.c0 - g.wrap_back : g.c0 + g.sx;
        const uint32_t n1 = (x + 1 == a.X) ? g.c1 - g.wrap_back : g.c1 + g.sx;
        load(p0, n0);
        load(p1, n1);
        g.h = (x + 1 == a.X) ? g.h - g.wrap_back : g.h + g.sx;
        if (g.ring) hu = u[g.h];
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x);
            l2_prefetch_strip(u, q, pf.count, pf.total); l2_prefetch_strip(dxa, q, pf.count, pf.total);
            l2_prefetch_strip(dya, q, pf.count, pf.total); l2_prefetch_strip(dza, q, pf.count, pf.total);
        }
        __syncthreads();
        const float *f = reinterpret_cast<const float *>(tl);
        const float2 un0 = one(c0, um0, p0.u, tl[o0 - RS], c1.u, f[(o0 - 1) * 2 + 1], f[(o0 + 1) * 2]);
        const float2 un1 = one(c1, um1, p1.u, c0.u, tl[o1 + RS], f[(o1 - 1) * 2 + 1], f[(o1 + 1) * 2]);
        store_pair(out, g.c0, un0, g.va0, g.vb0);
        store_pair(out, g.c1, un1, g.va1, g.vb1);
        if (g.mirror) {
            if (g.va0) out[g.c0 + g.zpad] = un0.x;
            if (g.va1) out[g.c1 + g.zpad] = un1.x;
        }
        if (counts(a, x)) {
            if (g.va0) s += (double)un0.x;
            if (g.vb0) s += (double)un0.y;
            if (g.va1) s += (double)un1.x;
            if (g.vb1) s += (double)un1.y;
        }
        um0 = c0.u; um1 = c1.u;
        g.c0 = n0; g.c1 = n1;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 2) {
        step(x, A0, A1, B0, B1);
        if (x + 1 >= g.x1) break;
        step(x + 1, B0, B1, A0, A1);
    }
    step_end<float, 1>(a, sc, s, 0.0);
}

// =====================================================================================================
// *_step_v6: the same z-pair kernels with ONE row per thread (two voxels per thread, 32 x TYT voxel tile
// on 16 x TYT threads).  Half the register state of *_step_v5, so the occupancy of the two-voxel
// kernels of tgtk_kernels_v3.cuh with the 64-bit accesses and the packed arithmetic of v5.
// =====================================================================================================
template <int TZL, int TYT, int MINB>
__global__ void __launch_bounds__(TZL *TYT, MINB) fk2c_step_v6(const StepArgs<float> a, const int lx)
{
    constexpr int TYV = TYT, RS = TZL + 2, PL = (TYV + 2) * RS;
    __shared__ float4 tile[2][2][PL];      // [buffer][(P_a, P_b, k_a, k_b) | (S_a, S_b, b_a, b_b)][slot]
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const float *__restrict__ P = a.buf[sc.par][0];
    const float *__restrict__ N = a.buf[sc.par][1];
    const float *__restrict__ S = a.buf[sc.par][2];
    float *__restrict__ Po = a.buf[sc.par ^ 1][0];
    float *__restrict__ No = a.buf[sc.par ^ 1][1];
    float *__restrict__ So = a.buf[sc.par ^ 1][2];
    const float *__restrict__ kc = a.coef[0];
    const float *__restrict__ bc = a.coef[1];
    Stage4 g = stage4_setup<TZL, TYT, 4, 1>(a, lx);
    const float closed = -Big<float>::v;
    const float th = a.th_necro;
    auto keff1 = [&](float pq, float nq, float kq) { return (pq + nq <= th) ? kq : closed; };
    auto keff2 = [&](float2 pq, float2 nq, float2 kq) {
        const float2 t = add2(pq, nq);
        return make_float2((t.x <= th) ? kq.x : closed, (t.y <= th) ? kq.y : closed);
    };
    auto load = [&](Vox2c4 &v, uint32_t off) {
        v.p = ld2(P, off); v.n = ld2(N, off); v.s = ld2(S, off); v.b = ld2(bc, off); v.k = ld2(kc, off);
    };
    const float2 cx2 = bc2(a.cx), cy2 = bc2(a.cy), cz2 = bc2(a.cz), ex2 = bc2(a.ex), ey2 = bc2(a.ey), ez2 = bc2(a.ez);
    const float2 dt2 = bc2(a.dt), rho2 = bc2(a.rho), lam2 = bc2(a.lambda_np), lams2 = bc2(a.lambda_s), one2 = bc2(1.0f);
    double sP = 0.0, sN = 0.0;
    griddep_launch();
    griddep_wait();

    Vox2c4 A, B;                           // the own pair in plane x (A) and plane x+1 (B); roles swap
    load(A, g.c0); A.k = keff2(A.p, A.n, A.k);
    float2 fp_lo, fs_lo;                   // fluxes through the faces below (x-1 | x)
    {
        const float2 pm = ld2(P, g.prev0);
        fp_lo = mul2(relu2(add2(keff2(pm, ld2(N, g.prev0), ld2(kc, g.prev0)), A.k)), sub2(pm, A.p));
        fs_lo = mul2(relu2(add2(ld2(bc, g.prev0), A.b)), sub2(ld2(S, g.prev0), A.s));
    }
    float hp = 0, hn = 0, hk = 0, hsv = 0, hb = 0;
    if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto step = [&](int x, const Vox2c4 &c, Vox2c4 &p) {
        float4 *t0 = tile[buf][0], *t1 = tile[buf][1];
        const int o = g.so;
        t0[o] = make_float4(c.p.x, c.p.y, c.k.x, c.k.y);
        t1[o] = make_float4(c.s.x, c.s.y, c.b.x, c.b.y);
        if (g.ring) {
            float *f0 = reinterpret_cast<float *>(t0), *f1 = reinterpret_cast<float *>(t1);
            f0[g.hs] = hp; f0[g.hs + 2] = keff1(hp, hn, hk);
            f1[g.hs] = hsv; f1[g.hs + 2] = hb;
        }
        const uint32_t n0 = (x + 1 == a.X) ? g.c0 - g.wrap_back : g.c0 + g.sx;
        load(p, n0);
        g.h = (x + 1 == a.X) ? g.h - g.wrap_back : g.h + g.sx;
        if (g.ring) { hp = P[g.h]; hn = N[g.h]; hk = kc[g.h]; hsv = S[g.h]; hb = bc[g.h]; }
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x);
            l2_prefetch_strip(P, q, pf.count, pf.total); l2_prefetch_strip(N, q, pf.count, pf.total);
            l2_prefetch_strip(S, q, pf.count, pf.total); l2_prefetch_strip(kc, q, pf.count, pf.total);
            l2_prefetch_strip(bc, q, pf.count, pf.total);
        }
        const float2 H = heaviside50_pair(c.s, a.sigma_np);
        const float2 react = sub2(mul2(mul2(rho2, mul2(c.s, c.p)), sub2(sub2(one2, c.p), c.n)), mul2(mul2(lam2, c.p), H));
        const float gp = relu(c.k.x + c.k.y) * (c.p.x - c.p.y), gs = relu(c.b.x + c.b.y) * (c.s.x - c.s.y);
        __syncthreads();
        const float4 pu = t0[o - RS], pd = t0[o + RS], su = t1[o - RS], sd = t1[o + RS];
        float pl = from_left<TZL>(c.p.y), kl = from_left<TZL>(c.k.y), sl = from_left<TZL>(c.s.y), bl = from_left<TZL>(c.b.y);
        float pr = from_right<TZL>(c.p.x), kr = from_right<TZL>(c.k.x), sr = from_right<TZL>(c.s.x), br = from_right<TZL>(c.b.x);
        if (threadIdx.x == 0) { const float4 e0 = t0[o - 1], e1 = t1[o - 1]; pl = e0.y; kl = e0.w; sl = e1.y; bl = e1.w; }
        if (threadIdx.x == TZL - 1) { const float4 e0 = t0[o + 1], e1 = t1[o + 1]; pr = e0.x; kr = e0.z; sr = e1.x; br = e1.z; }
        float2 sp = mul2(cy2, sub2(mul2(relu2(add2(make_float2(pu.z, pu.w), c.k)), sub2(make_float2(pu.x, pu.y), c.p)),
                                   mul2(relu2(add2(c.k, make_float2(pd.z, pd.w))), sub2(c.p, make_float2(pd.x, pd.y)))));
        float2 ss = mul2(ey2, sub2(mul2(relu2(add2(make_float2(su.z, su.w), c.b)), sub2(make_float2(su.x, su.y), c.s)),
                                   mul2(relu2(add2(c.b, make_float2(sd.z, sd.w))), sub2(c.s, make_float2(sd.x, sd.y)))));
        {
            const float fl = relu(kl + c.k.x) * (pl - c.p.x), fr = relu(c.k.y + kr) * (c.p.y - pr);
            sp = fma2(cz2, make_float2(fl - gp, gp - fr), sp);
            const float hl = relu(bl + c.b.x) * (sl - c.s.x), hr = relu(c.b.y + br) * (c.s.y - sr);
            ss = fma2(ez2, make_float2(hl - gs, gs - hr), ss);
        }
        p.k = keff2(p.p, p.n, p.k);
        const float2 fp_hi = mul2(relu2(add2(c.k, p.k)), sub2(c.p, p.p)), fs_hi = mul2(relu2(add2(c.b, p.b)), sub2(c.s, p.s));
        sp = fma2(cx2, sub2(fp_lo, fp_hi), sp);
        ss = fma2(ex2, sub2(fs_lo, fs_hi), ss);
        const float2 pn = fma2(dt2, add2(sp, react), c.p);
        const float2 nn = fma2(dt2, mul2(mul2(lam2, pn), H), c.n);
        const float2 sn = fma2(dt2, sub2(ss, mul2(mul2(lams2, c.s), pn)), c.s);
        store_pair(Po, g.c0, pn, g.va0, g.vb0); store_pair(No, g.c0, nn, g.va0, g.vb0); store_pair(So, g.c0, sn, g.va0, g.vb0);
        if (g.mirror && g.va0) { Po[g.c0 + g.zpad] = pn.x; No[g.c0 + g.zpad] = nn.x; So[g.c0 + g.zpad] = sn.x; }
        if (counts(a, x)) {
            if (g.va0) { sP += (double)pn.x; sN += (double)nn.x; }
            if (g.vb0) { sP += (double)pn.y; sN += (double)nn.y; }
        }
        fp_lo = fp_hi; fs_lo = fs_hi;
        g.c0 = n0;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 2) {
        step(x, A, B);
        if (x + 1 >= g.x1) break;
        step(x + 1, B, A);
    }
    step_end<float, 2>(a, sc, sP, sN);
}

template <int TZL, int TYT, int MINB>
__global__ void __launch_bounds__(TZL *TYT, MINB) fk_step_v6(const StepArgs<float> a, const int lx)
{
    constexpr int TYV = TYT, RS = TZL + 2, PL = (TYV + 2) * RS;
    __shared__ float4 tile[2][PL];         // [buffer][slot] of (u_a, u_b, k_a, k_b)
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const float *__restrict__ u = a.buf[sc.par][0];
    float *__restrict__ out = a.buf[sc.par ^ 1][0];
    const float *__restrict__ kc = a.coef[0];
    Stage4 g = stage4_setup<TZL, TYT, 4, 1>(a, lx);
    const float2 cx2 = bc2(a.cx), cy2 = bc2(a.cy), cz2 = bc2(a.cz), dt2 = bc2(a.dt), rho2 = bc2(a.rho), one2 = bc2(1.0f);
    double s = 0.0;
    griddep_launch();
    griddep_wait();

    VoxFk4 A, B;
    A.u = ld2(u, g.c0); A.k = ld2(kc, g.c0);
    float2 f_lo = mul2(relu2(add2(ld2(kc, g.prev0), A.k)), sub2(ld2(u, g.prev0), A.u));
    float hu = 0, hk = 0;
    if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto step = [&](int x, const VoxFk4 &c, VoxFk4 &p) {
        float4 *tl = tile[buf];
        const int o = g.so;
        tl[o] = make_float4(c.u.x, c.u.y, c.k.x, c.k.y);
        if (g.ring) {
            float *f = reinterpret_cast<float *>(tl);
            f[g.hs] = hu; f[g.hs + 2] = hk;
        }
        const uint32_t n0 = (x + 1 == a.X) ? g.c0 - g.wrap_back : g.c0 + g.sx;
        p.u = ld2(u, n0); p.k = ld2(kc, n0);
        g.h = (x + 1 == a.X) ? g.h - g.wrap_back : g.h + g.sx;
        if (g.ring) { hu = u[g.h]; hk = kc[g.h]; }
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x);
            l2_prefetch_strip(u, q, pf.count, pf.total); l2_prefetch_strip(kc, q, pf.count, pf.total);
        }
        const float2 react = mul2(mul2(rho2, c.u), sub2(one2, c.u));
        const float g0 = relu(c.k.x + c.k.y) * (c.u.x - c.u.y);
        __syncthreads();
        const float4 up = tl[o - RS], dn = tl[o + RS];
        float ul = from_left<TZL>(c.u.y), kl = from_left<TZL>(c.k.y), ur = from_right<TZL>(c.u.x), kr = from_right<TZL>(c.k.x);
        if (threadIdx.x == 0) { const float4 e = tl[o - 1]; ul = e.y; kl = e.w; }
        if (threadIdx.x == TZL - 1) { const float4 e = tl[o + 1]; ur = e.x; kr = e.z; }
        float2 sp = mul2(cy2, sub2(mul2(relu2(add2(make_float2(up.z, up.w), c.k)), sub2(make_float2(up.x, up.y), c.u)),
                                   mul2(relu2(add2(c.k, make_float2(dn.z, dn.w))), sub2(c.u, make_float2(dn.x, dn.y)))));
        const float fl = relu(kl + c.k.x) * (ul - c.u.x), fr = relu(c.k.y + kr) * (c.u.y - ur);
        sp = fma2(cz2, make_float2(fl - g0, g0 - fr), sp);
        const float2 f_hi = mul2(relu2(add2(c.k, p.k)), sub2(c.u, p.u));
        sp = fma2(cx2, sub2(f_lo, f_hi), sp);
        const float2 un = fma2(dt2, add2(sp, react), c.u);
        store_pair(out, g.c0, un, g.va0, g.vb0);
        if (g.mirror && g.va0) out[g.c0 + g.zpad] = un.x;
        if (counts(a, x)) {
            if (g.va0) s += (double)un.x;
            if (g.vb0) s += (double)un.y;
        }
        f_lo = f_hi;
        g.c0 = n0;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 2) {
        step(x, A, B);
        if (x + 1 >= g.x1) break;
        step(x + 1, B, A);
    }
    step_end<float, 1>(a, sc, s, 0.0);
}

template <int TZL, int TYT, int MINB>
__global__ void __launch_bounds__(TZL *TYT, MINB) dti_step_v6(const StepArgs<float> a, const int lx)
{
    using R = Rn<float>;
    constexpr int TYV = TYT, RS = TZL + 2, PL = (TYV + 2) * RS;
    __shared__ float2 tile[2][PL];         // [buffer][slot] of (u_a, u_b)
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    const float *__restrict__ u = a.buf[sc.par][0];
    float *__restrict__ out = a.buf[sc.par ^ 1][0];
    const float *__restrict__ dxa = a.coef[0];
    const float *__restrict__ dya = a.coef[1];
    const float *__restrict__ dza = a.coef[2];
    Stage4 g = stage4_setup<TZL, TYT, 2, 1>(a, lx);
    const float2 cx2 = bc2(a.cx), cy2 = bc2(a.cy), dt2 = bc2(a.dt), rho2 = bc2(a.rho), one2 = bc2(1.0f);
    double s = 0.0;
    auto load = [&](VoxDti4 &v, uint32_t off) { v.u = ld2(u, off); v.d0 = ld2(dxa, off); v.d1 = ld2(dya, off); v.d2 = ld2(dza, off); };
    griddep_launch();
    griddep_wait();

    VoxDti4 A, B;
    load(A, g.c0);
    float2 um = ld2(u, g.prev0);
    float hu = 0;
    if (g.ring) hu = u[g.h];
    int buf = 0;
    const L2Strip pf = l2_strip_setup(a, TYV);

    auto step = [&](int x, const VoxDti4 &c, VoxDti4 &p) {
        float2 *tl = tile[buf];
        const int o = g.so;
        tl[o] = c.u;
        if (g.ring) reinterpret_cast<float *>(tl)[g.hs] = hu;
        const uint32_t n0 = (x + 1 == a.X) ? g.c0 - g.wrap_back : g.c0 + g.sx;
        load(p, n0);
        g.h = (x + 1 == a.X) ? g.h - g.wrap_back : g.h + g.sx;
        if (g.ring) hu = u[g.h];
        if (pf.on) {
            const uint32_t q = l2_strip_at(a, pf, x);
            l2_prefetch_strip(u, q, pf.count, pf.total); l2_prefetch_strip(dxa, q, pf.count, pf.total);
            l2_prefetch_strip(dya, q, pf.count, pf.total); l2_prefetch_strip(dza, q, pf.count, pf.total);
        }
        __syncthreads();
        const float *f = reinterpret_cast<const float *>(tl);
        const float2 uym = tl[o - RS], uyp = tl[o + RS];
        const float uzl = f[(o - 1) * 2 + 1], uzr = f[(o + 1) * 2];
        const float2 sx = mul2(cx2, sub2(mul2(c.d0, sub2(um, c.u)), mul2(c.d0, sub2(c.u, p.u))));
        const float2 sy = mul2(cy2, sub2(mul2(c.d1, sub2(uym, c.u)), mul2(c.d1, sub2(c.u, uyp))));
        float2 sz;
        sz.x = R::mul(a.cz, R::sub(R::mul(c.d2.x, R::sub(uzl, c.u.x)), R::mul(c.d2.x, R::sub(c.u.x, c.u.y))));
        sz.y = R::mul(a.cz, R::sub(R::mul(c.d2.y, R::sub(c.u.x, c.u.y)), R::mul(c.d2.y, R::sub(c.u.y, uzr))));
        const float2 sp = add2(add2(sx, sy), sz);
        const float2 react = mul2(rho2, mul2(c.u, sub2(one2, c.u)));
        const float2 un = add2(c.u, mul2(add2(sp, react), dt2));
        store_pair(out, g.c0, un, g.va0, g.vb0);
        if (g.mirror && g.va0) out[g.c0 + g.zpad] = un.x;
        if (counts(a, x)) {
            if (g.va0) s += (double)un.x;
            if (g.vb0) s += (double)un.y;
        }
        um = c.u;
        g.c0 = n0;
        buf ^= 1;
    };
    for (int x = g.x0; x < g.x1; x += 2) {
        step(x, A, B);
        if (x + 1 >= g.x1) break;
        step(x + 1, B, A);
    }
    step_end<float, 1>(a, sc, s, 0.0);
}

}  // namespace tgtk
