// FK_2c step kernel with TMA tile staging and a warp-specialised producer (variant 7).
//
// Same decomposition as fk2c_step_v4 -- a block owns a 32(z) x 2*TYT(y) column tile and a segment of x planes, a
// consumer thread owns the rows (y, y+1) of one z -- but no thread of the block issues a global LOAD:
//   * a producer warp asks the TMA unit for the (TZ+2) x (TYV+2) tile of plane x of each of the five arrays
//     (cp.async.bulk.tensor.3d, SASS UTMALDG), NST planes deep, straight into shared memory; completion is counted in
//     bytes on an mbarrier per stage ("full"), a second mbarrier per stage ("empty", one arrival per consumer warp)
//     hands the slot back.  There is no __syncthreads() in the plane loop: consumer warps only ever wait for data.
//   * own-column values, the four in-plane neighbours and the plane ahead are all read from the staged tiles, so the
//     second register set and the 80 divergent ring loaders of fk2c_step_v4 are gone.
//   * periodic wrap (np.roll, FK/FK.py:12-18,36-38): TMA zero-fills what lies outside the tensor, so the blocks on a
//     box face additionally fetch the opposite face as a thin strip (a 16-byte column strip for z, one row for y) and
//     the affected threads read their neighbour there: every thread keeps eight byte offsets (own rows, up, down,
//     left / right of both rows) that are fixed for the whole march, so the plane loop has no edge branches.
// The arithmetic is fk2c_step_v4's, expression by expression (bitwise equal results).
//
// Requirements (checked by the host, which otherwise keeps variant 4): even row pitch (16-byte global strides),
// 32-bit element offsets.  Measured on B200 (tools/probes/tma_probe.cu): the innermost START COORDINATE of a box must
// be a multiple of 16 bytes too -- an odd z start for fp64 raises "illegal instruction" -- negative and out-of-range
// aligned starts are fine (zero fill).  Hence the tile starts CW = 16 / sizeof(T) elements left of z0, not one.
#pragma once
#include <cuda.h>
#include "tgtk_common.cuh"
#include "tgtk_kernels_v3.cuh"

namespace tgtk {

struct TmaSet {
    CUtensorMap main[5], col[5], row[5];      // P, N, S, k, b of the buffer a launch reads
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <typename T, int TYT> struct TmaGeom {
    static constexpr int TZ = 32, TYV = 2 * TYT;
    static constexpr int CW = 16 / (int)sizeof(T);                  // elements of a 16-byte column strip
    static constexpr int RS = TZ + 2 * CW;                          // tile row in elements: z0-CW .. z0+TZ+CW-1 (16-byte aligned start)
    static constexpr int MAINB = ((RS * (TYV + 2) * (int)sizeof(T) + 127) / 128) * 128;
    static constexpr int COLB = ((16 * (TYV + 2) + 127) / 128) * 128;
    static constexpr int ROWB = ((RS * (int)sizeof(T) + 127) / 128) * 128;
    static constexpr int MAIN_TX = RS * (TYV + 2) * (int)sizeof(T), COL_TX = 16 * (TYV + 2), ROW_TX = RS * (int)sizeof(T);
    // one array's region of a stage: the tile, two column strips (z = -1, z = Z), two row strips (y = -1, y = Y); a
    // compile-time constant so that every staged value is one LDS with an immediate field offset
    static constexpr int FS = MAINB + 2 * COLB + 2 * ROWB, SS = 5 * FS;
    static __host__ __device__ constexpr int smem_bytes(int nst) { return 128 + 128 + nst * SS; }
};

template <typename T, int TYT, int NST, int MINB>
__global__ void __launch_bounds__(32 * (TYT + 1), MINB)
    fk2c_step_tma(const StepArgs<T> a, const int lx, const __grid_constant__ TmaSet maps)
{
    using G = TmaGeom<T, TYT>;
    constexpr int TZ = G::TZ, TYV = G::TYV, RS = G::RS, ES = (int)sizeof(T);
    extern __shared__ unsigned char smem_raw[];
    const StepCtl sc = step_begin(a);
    if (sc.stop) return;
    // dynamic shared memory: [full[NST] | empty[NST]] (128 B), then the stages, 128-byte aligned (TMA destination)
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    unsigned char *const sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_full = base, bar_empty = base + 8 * NST;
    constexpr int FS = G::FS, SS = G::SS;
    unsigned char *const stages = sm + 128;
    const uint32_t stages_u32 = base + 128;

    const int tz = threadIdx.x, ty = threadIdx.y;
    const bool producer = (ty == TYT);
    const int z0 = blockIdx.x * TZ, y0 = blockIdx.y * TYV;
    const int x0 = a.x_lo + blockIdx.z * lx, x1 = min(x0 + lx, a.x_hi);
    const int nload = x1 - x0 + 2;                                   // planes x0-1 .. x1, in march order
    // block-uniform: which box faces this tile touches
    const bool zlo = (z0 == 0), zhi = (z0 + TZ >= a.Z), ylo = (y0 == 0), yhi = (y0 + TYV >= a.Y);
    const int zlast = ((a.Z - 1) / G::CW) * G::CW;                   // aligned start of the column strip that holds z = Z-1

    if (tz == 0 && ty == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, TYT);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    griddep_launch();
    griddep_wait();

    double sP = 0.0, sN = 0.0;
    if (producer) {
        // ---- producer warp: one lane drives the TMA unit (UTMALDG takes uniform operands) --------------------------
        const uint32_t tx = 5u * (uint32_t)(G::MAIN_TX + ((zlo ? 1 : 0) + (zhi ? 1 : 0)) * G::COL_TX + ((ylo ? 1 : 0) + (yhi ? 1 : 0)) * G::ROW_TX);
        const L2Strip pf = l2_strip_setup(a, TYV);
        const bool pf_on = a.l2pf > 0 && blockIdx.x == 0 && tz == 5;  // a spare lane of the z = 0 tiles' producer
        const T *arrs[5] = {a.buf[sc.par][0], a.buf[sc.par][1], a.buf[sc.par][2], a.coef[0], a.coef[1]};
        int xp = (x0 == 0) ? a.X - 1 : x0 - 1;
        int s = 0;
        uint32_t ph = 0;                                             // parity of the phase load j completes on full[s]
        for (int j = 0; j < nload; ++j) {
            if (tz == 0) {
                if (j >= NST) mbar_wait(bar_empty + 8 * s, ph ^ 1);  // the consumers are done with the previous tenant of the slot
                const uint32_t bar = bar_full + 8 * s;
                mbar_expect_tx(bar, tx);
#pragma unroll
                for (int f = 0; f < 5; ++f) {
                    const uint32_t dst = stages_u32 + s * SS + f * FS;
                    tma_load_3d(dst, &maps.main[f], bar, z0 - G::CW, y0 - 1, xp);
                    if (zlo) tma_load_3d(dst + G::MAINB, &maps.col[f], bar, zlast, y0 - 1, xp);
                    if (zhi) tma_load_3d(dst + G::MAINB + G::COLB, &maps.col[f], bar, 0, y0 - 1, xp);
                    if (ylo) tma_load_3d(dst + G::MAINB + 2 * G::COLB, &maps.row[f], bar, z0 - G::CW, a.Y - 1, xp);
                    if (yhi) tma_load_3d(dst + G::MAINB + 2 * G::COLB + G::ROWB, &maps.row[f], bar, z0 - G::CW, 0, xp);
                }
            }
            if (++s == NST) { s = 0; ph ^= 1; }
            if (pf_on) {                                             // the strip of the plane l2pf loads ahead into L2
                const int xq = (xp + a.l2pf) % a.X;
                const uint32_t q = (uint32_t)xq * pf.plane + pf.row0;
#pragma unroll
                for (int k = 0; k < 5; ++k) l2_prefetch_strip(arrs[k], q, pf.count, pf.total);
            }
            xp = (xp + 1 == a.X) ? 0 : xp + 1;
        }
    } else {
        // ---- consumer warps ----------------------------------------------------------------------------------
        T *__restrict__ Po = a.buf[sc.par ^ 1][0];
        T *__restrict__ No = a.buf[sc.par ^ 1][1];
        T *__restrict__ So = a.buf[sc.par ^ 1][2];
        const int r0 = y0 + 2 * ty, z = z0 + tz;
        const bool valid0 = (z < a.Z) && (r0 < a.Y), valid1 = (z < a.Z) && (r0 + 1 < a.Y);
        // byte offsets inside one array's region of a stage; fixed for the whole march
        constexpr int col_lo = G::MAINB, col_hi = G::MAINB + G::COLB;
        constexpr int row_lo = G::MAINB + 2 * G::COLB, row_hi = row_lo + G::ROWB;
        const int t0 = ((2 * ty + 1) * RS + tz + G::CW) * ES;         // tile slot of (r0, z)
        int o0 = t0, o1 = t0 + RS * ES, oup = t0 - RS * ES, odn = t0 + 2 * RS * ES;
        int ol0 = o0 - ES, or0 = o0 + ES, ol1 = o1 - ES, or1 = o1 + ES;
        if (r0 == 0) oup = row_lo + (tz + G::CW) * ES;                // y = -1 -> Y-1
        if (r0 + 1 == a.Y) o1 = ol1 = or1 = odn = row_hi + (tz + G::CW) * ES;   // the lower own row is y = Y -> 0 (shadow, nothing stored)
        else if (r0 + 2 == a.Y) odn = row_hi + (tz + G::CW) * ES;    // y = Y -> 0
        if (z == 0) {                                                 // z = -1 -> Z-1, inside the strip that starts at zlast
            const int e = a.Z - 1 - zlast;
            ol0 = col_lo + ((2 * ty + 1) * G::CW + e) * ES;
            if (r0 + 1 != a.Y) ol1 = col_lo + ((2 * ty + 2) * G::CW + e) * ES;
        }
        if (z + 1 == a.Z) {                                           // z = Z -> 0: first element of the strip that starts at 0
            or0 = col_hi + ((2 * ty + 1) * G::CW) * ES;
            if (r0 + 1 != a.Y) or1 = col_hi + ((2 * ty + 2) * G::CW) * ES;
        }
        uint32_t g0 = (uint32_t)x0 * (uint32_t)a.sx + (uint32_t)min(r0, a.Y - 1) * (uint32_t)a.sy + (uint32_t)min(z, a.Z - 1);
        uint32_t g1 = g0 + (uint32_t)a.sy;
        const uint32_t gsx = (uint32_t)a.sx;
        const T closed = -Big<T>::v;
        auto keff = [&](T pq, T nq, T kq) { return (pq + nq <= a.th_necro) ? kq : closed; };
        auto at = [&](const unsigned char *fb, int off) { return *(const T *)(fb + off); };
        // own voxel of one row from a stage: (p, n, s, b, keff)
        auto own = [&](Vox2c<T> &v, const unsigned char *st, int off) {
            v.p = at(st, off); v.n = at(st + FS, off); v.s = at(st + 2 * FS, off);
            v.k = keff(v.p, v.n, at(st + 3 * FS, off)); v.b = at(st + 4 * FS, off);
        };
        const int lane0 = (tz == 0);

        Vox2c<T> A0, A1, B0, B1;
        T fp_lo0, fs_lo0, fp_lo1, fs_lo1;
        // stage / phase parity of the load that holds the current plane (load 0 = plane x0-1, load 1 = plane x0, ...)
        int sj = 0;
        uint32_t phj = 0;
        auto next_stage = [&](int &sn, uint32_t &phn) { sn = (sj + 1 == NST) ? 0 : sj + 1; phn = phj ^ (sn == 0 ? 1u : 0u); };
        {   // prologue: plane x0-1 only contributes the fluxes through the faces below plane x0
            mbar_wait(bar_full + 8 * sj, phj);
            const unsigned char *sb = stages + sj * SS;
            Vox2c<T> M0, M1;
            own(M0, sb, o0); own(M1, sb, o1);
            int sn; uint32_t phn;
            next_stage(sn, phn);
            mbar_wait(bar_full + 8 * sn, phn);
            const unsigned char *sa = stages + sn * SS;
            own(A0, sa, o0); own(A1, sa, o1);
            fp_lo0 = relu(M0.k + A0.k) * (M0.p - A0.p);
            fs_lo0 = relu(M0.b + A0.b) * (M0.s - A0.s);
            fp_lo1 = relu(M1.k + A1.k) * (M1.p - A1.p);
            fs_lo1 = relu(M1.b + A1.b) * (M1.s - A1.s);
            __syncwarp();
            if (lane0) mbar_arrive(bar_empty + 8 * sj);
            sj = sn; phj = phn;
        }
        auto step = [&](int x, const Vox2c<T> &c0, const Vox2c<T> &c1, Vox2c<T> &p0, Vox2c<T> &p1) {
            // plane x sits in stage sj (already waited for), plane x+1 in the next one
            const unsigned char *sb = stages + sj * SS;
            int sn_; uint32_t phn;
            next_stage(sn_, phn);
            mbar_wait(bar_full + 8 * sn_, phn);
            const unsigned char *sn = stages + sn_ * SS;
            own(p0, sn, o0); own(p1, sn, o1);
            // work that needs no neighbour
            const T H0 = heaviside50_fast(c0.s, a.sigma_np), H1 = heaviside50_fast(c1.s, a.sigma_np);
            const T react0 = a.rho * (c0.s * c0.p) * ((T(1) - c0.p) - c0.n) - (a.lambda_np * c0.p) * H0;
            const T react1 = a.rho * (c1.s * c1.p) * ((T(1) - c1.p) - c1.n) - (a.lambda_np * c1.p) * H1;
            const T fp_mid = relu(c0.k + c1.k) * (c0.p - c1.p);
            const T fs_mid = relu(c0.b + c1.b) * (c0.s - c1.s);
            // in-plane neighbours of plane x: (P, k_eff) and (S, b)
            const unsigned char *fP = sb, *fN = sb + FS, *fS = sb + 2 * FS, *fK = sb + 3 * FS, *fB = sb + 4 * FS;
            auto nbP = [&](int off, T &pv, T &kv) { pv = at(fP, off); kv = keff(pv, at(fN, off), at(fK, off)); };
            T pu, ku, pd, kd, pl0, kl0, pr0, kr0, pl1, kl1, pr1, kr1;
            nbP(oup, pu, ku); nbP(odn, pd, kd); nbP(ol0, pl0, kl0); nbP(or0, pr0, kr0); nbP(ol1, pl1, kl1); nbP(or1, pr1, kr1);
            T sp0 = a.cy * (relu(ku + c0.k) * (pu - c0.p) - fp_mid);
            T sp1 = a.cy * (fp_mid - relu(c1.k + kd) * (c1.p - pd));
            T ss0 = a.ey * (relu(at(fB, oup) + c0.b) * (at(fS, oup) - c0.s) - fs_mid);
            T ss1 = a.ey * (fs_mid - relu(c1.b + at(fB, odn)) * (c1.s - at(fS, odn)));
            sp0 += a.cz * (relu(kl0 + c0.k) * (pl0 - c0.p) - relu(c0.k + kr0) * (c0.p - pr0));
            sp1 += a.cz * (relu(kl1 + c1.k) * (pl1 - c1.p) - relu(c1.k + kr1) * (c1.p - pr1));
            ss0 += a.ez * (relu(at(fB, ol0) + c0.b) * (at(fS, ol0) - c0.s) - relu(c0.b + at(fB, or0)) * (c0.s - at(fS, or0)));
            ss1 += a.ez * (relu(at(fB, ol1) + c1.b) * (at(fS, ol1) - c1.s) - relu(c1.b + at(fB, or1)) * (c1.s - at(fS, or1)));
            // every read of plane x's stage is done: hand the slot back to the producer
            __syncwarp();
            if (lane0) mbar_arrive(bar_empty + 8 * sj);
            // x faces
            const T fp_hi0 = relu(c0.k + p0.k) * (c0.p - p0.p), fs_hi0 = relu(c0.b + p0.b) * (c0.s - p0.s);
            const T fp_hi1 = relu(c1.k + p1.k) * (c1.p - p1.p), fs_hi1 = relu(c1.b + p1.b) * (c1.s - p1.s);
            sp0 += a.cx * (fp_lo0 - fp_hi0);
            sp1 += a.cx * (fp_lo1 - fp_hi1);
            ss0 += a.ex * (fs_lo0 - fs_hi0);
            ss1 += a.ex * (fs_lo1 - fs_hi1);
            const T pn0 = c0.p + a.dt * (sp0 + react0), pn1 = c1.p + a.dt * (sp1 + react1);
            const T nn0 = c0.n + a.dt * ((a.lambda_np * pn0) * H0), nn1 = c1.n + a.dt * ((a.lambda_np * pn1) * H1);
            const T sn0 = c0.s + a.dt * (ss0 - (a.lambda_s * c0.s) * pn0), sn1 = c1.s + a.dt * (ss1 - (a.lambda_s * c1.s) * pn1);
            const bool cnt = counts(a, x);
            if (valid0) {
                Po[g0] = pn0; No[g0] = nn0; So[g0] = sn0;
                if (cnt) { sP += (double)pn0; sN += (double)nn0; }
            }
            if (valid1) {
                Po[g1] = pn1; No[g1] = nn1; So[g1] = sn1;
                if (cnt) { sP += (double)pn1; sN += (double)nn1; }
            }
            fp_lo0 = fp_hi0; fs_lo0 = fs_hi0; fp_lo1 = fp_hi1; fs_lo1 = fs_hi1;
            g0 += gsx; g1 += gsx;
            sj = sn_; phj = phn;
        };
        for (int x = x0; x < x1; x += 2) {
            step(x, A0, A1, B0, B1);
            if (x + 1 >= x1) break;
            step(x + 1, B0, B1, A0, A1);
        }
        // the last "ahead" plane (load nload-1) was only read for the x faces: release it too (nobody waits; keeps the protocol symmetric)
    }
    step_end<T, 2>(a, sc, sP, sN);
}

}  // namespace tgtk
