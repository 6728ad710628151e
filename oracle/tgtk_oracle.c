/*
 * tgtk_oracle.c -- CPU restatement of TumorGrowthToolkit's explicit time-stepping path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call it, and only as the checker / CPU baseline.  The product path
 * (tumorgrowthtoolkit_b200/) never imports it and fails loudly without its CUDA library.
 *
 * Parity status: PINNED.  tests/golden/ holds vectors produced by running the reference
 * itself (tests/make_golden.py imports /root/reference); tests/test_oracle_golden.py
 * checks every function below against them.
 *
 * Form: the reference is whole-array numpy (np.roll / np.where temporaries).  This file
 * restates the same arithmetic per voxel, in the reference's floating-point evaluation
 * order (compile with -ffp-contract=off so no FMA is formed), so FK and FK_DTI steps match
 * numpy bit for bit; FK_2c differs only through libm exp() vs numpy's SIMD exp (<= 1 ulp).
 *
 * Layout everywhere: C-order (X,Y,Z), axis 2 fastest, contiguous float64.  Neighbour
 * access wraps modulo the box (np.roll semantics, reference FK/FK.py:12-18,36-38).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j, k) (((size_t)(i) * Y + (size_t)(j)) * Z + (size_t)(k))

int tgo_abi_version(void) { return 1; }

int tgo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void tgo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/*
 * Face coefficients ("D_minus_a": the face between voxel c and c+1 along axis a).
 *   FK/FK.py:10-20  (m_Tildas)        tissue condition at both ends of the face
 *   FK/FK.py:22-26  (get_D)           D_minus = Dw*(WM_t + GM_t/ratio)
 *   FK_2c/FK_2c.py:10-40              extra condition (P+N) <= th_necro at both ends
 * D_plus_a is np.roll(D_minus_a, +1, a) (FK.py:28-30): the same numbers one voxel up, so
 * it is not materialised here; tgo_roll_plus() makes it when a caller wants the dict.
 * PC/PN may be NULL (plain FK faces).
 */
void tgo_faces(const double *WM, const double *GM, const double *PC, const double *PN,
               double th, double th_necro, double Dcoef, double ratio,
               int X, int Y, int Z, double *Dmx, double *Dmy, double *Dmz)
{
    const int necro = (PC != NULL && PN != NULL);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            for (int k = 0; k < Z; ++k) {
                const size_t c = IDX(i, j, k);
                const size_t nb[3] = { IDX((i + 1 == X) ? 0 : i + 1, j, k),
                                       IDX(i, (j + 1 == Y) ? 0 : j + 1, k),
                                       IDX(i, j, (k + 1 == Z) ? 0 : k + 1) };
                double *out[3] = { Dmx, Dmy, Dmz };
                const int here_ok = (WM[c] + GM[c] >= th) &&
                                    (!necro || (PC[c] + PN[c] <= th_necro));
                for (int a = 0; a < 3; ++a) {
                    const size_t n = nb[a];
                    const int there_ok = (WM[n] + GM[n] >= th) &&
                                         (!necro || (PC[n] + PN[n] <= th_necro));
                    double wt = 0.0, gt = 0.0;
                    if (here_ok && there_ok) {
                        wt = (WM[n] + WM[c]) / 2;
                        gt = (GM[n] + GM[c]) / 2;
                    }
                    out[a][c] = Dcoef * (wt + gt / ratio);
                }
            }
}

/* np.roll(src, +1, axis): dst[c] = src[c-1 along axis] with wrap (FK.py:28-30). */
void tgo_roll_plus(const double *src, int axis, int X, int Y, int Z, double *dst)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            for (int k = 0; k < Z; ++k) {
                int ii = i, jj = j, kk = k;
                if (axis == 0) ii = (i == 0) ? X - 1 : i - 1;
                if (axis == 1) jj = (j == 0) ? Y - 1 : j - 1;
                if (axis == 2) kk = (k == 0) ? Z - 1 : k - 1;
                dst[IDX(i, j, k)] = src[IDX(ii, jj, kk)];
            }
}

static inline double stencil_term(double inv_h2, double dp, double dm, double um, double u,
                                  double up)
{
    /* 1/(h*h) * (D_plus*(u[c-1]-u) - D_minus*(u - u[c+1]))      FK.py:36-38 */
    return inv_h2 * (dp * (um - u) - dm * (u - up));
}

/*
 * One FK_update (FK/FK.py:34-42) with the six coefficient arrays given explicitly, exactly
 * as the reference method takes them (so it also serves FK_DTI, where D_plus aliases
 * D_minus, FK_DTI/FK_DTI.py:39-45).  src and dst must not alias.
 */
void tgo_fk_step(const double *u, const double *Dpx, const double *Dpy, const double *Dpz,
                 const double *Dmx, const double *Dmy, const double *Dmz,
                 double rho, double dt, double dx, double dy, double dz,
                 int X, int Y, int Z, double *out)
{
    const double ix = 1 / (dx * dx), iy = 1 / (dy * dy), iz = 1 / (dz * dz);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            const int im = (i == 0) ? X - 1 : i - 1, ip = (i + 1 == X) ? 0 : i + 1;
            const int jm = (j == 0) ? Y - 1 : j - 1, jp = (j + 1 == Y) ? 0 : j + 1;
            for (int k = 0; k < Z; ++k) {
                const int km = (k == 0) ? Z - 1 : k - 1, kp = (k + 1 == Z) ? 0 : k + 1;
                const size_t c = IDX(i, j, k);
                const double a = u[c];
                const double sx = stencil_term(ix, Dpx[c], Dmx[c], u[IDX(im, j, k)], a, u[IDX(ip, j, k)]);
                const double sy = stencil_term(iy, Dpy[c], Dmy[c], u[IDX(i, jm, k)], a, u[IDX(i, jp, k)]);
                const double sz = stencil_term(iz, Dpz[c], Dmz[c], u[IDX(i, j, km)], a, u[IDX(i, j, kp)]);
                const double sp = (sx + sy) + sz;
                /* diff_A = (SP + f*np.multiply(A,1-A)) * dt ; A += diff_A   FK.py:40-41 */
                out[c] = a + (sp + rho * (a * (1 - a))) * dt;
            }
        }
}

/*
 * FK step with D_plus taken as the rolled D_minus (what FK.Solver.solve feeds FK_update,
 * FK.py:200,213): saves the three rolled copies.
 */
void tgo_fk_step_faces(const double *u, const double *Dmx, const double *Dmy, const double *Dmz,
                       double rho, double dt, double dx, double dy, double dz,
                       int X, int Y, int Z, double *out)
{
    const double ix = 1 / (dx * dx), iy = 1 / (dy * dy), iz = 1 / (dz * dz);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            const int im = (i == 0) ? X - 1 : i - 1, ip = (i + 1 == X) ? 0 : i + 1;
            const int jm = (j == 0) ? Y - 1 : j - 1, jp = (j + 1 == Y) ? 0 : j + 1;
            for (int k = 0; k < Z; ++k) {
                const int km = (k == 0) ? Z - 1 : k - 1, kp = (k + 1 == Z) ? 0 : k + 1;
                const size_t c = IDX(i, j, k);
                const double a = u[c];
                const double sx = stencil_term(ix, Dmx[IDX(im, j, k)], Dmx[c], u[IDX(im, j, k)], a, u[IDX(ip, j, k)]);
                const double sy = stencil_term(iy, Dmy[IDX(i, jm, k)], Dmy[c], u[IDX(i, jm, k)], a, u[IDX(i, jp, k)]);
                const double sz = stencil_term(iz, Dmz[IDX(i, j, km)], Dmz[c], u[IDX(i, j, km)], a, u[IDX(i, j, kp)]);
                const double sp = (sx + sy) + sz;
                out[c] = a + (sp + rho * (a * (1 - a))) * dt;
            }
        }
}

/* smooth_heaviside, FK_2c/FK_2c.py:75-76:  1 - 1/(1 + exp(-k*(x - sigma))),  k = 50 */
static inline double heaviside50(double s, double sigma)
{
    return 1 - 1 / (1 + exp(-50.0 * (s - sigma)));
}

/*
 * One FK_2c_update (FK_2c/FK_2c.py:48-73) with explicit coefficient arrays (D for P, Ds for
 * S; plus = face below, minus = face above).  In-place ordering of the reference: every
 * stencil reads the OLD P and S; N and the S sink use the NEW P (FK_2c.py:58-71).
 * Outputs must not alias inputs.
 */
void tgo_fk2c_step(const double *P, const double *N, const double *S,
                   const double *Dpx, const double *Dpy, const double *Dpz,
                   const double *Dmx, const double *Dmy, const double *Dmz,
                   const double *Epx, const double *Epy, const double *Epz,
                   const double *Emx, const double *Emy, const double *Emz,
                   double rho, double lambda_np, double sigma_np, double lambda_s,
                   double dt, double dx, double dy, double dz,
                   int X, int Y, int Z, double *Po, double *No, double *So)
{
    const double ix = 1 / (dx * dx), iy = 1 / (dy * dy), iz = 1 / (dz * dz);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            const int im = (i == 0) ? X - 1 : i - 1, ip = (i + 1 == X) ? 0 : i + 1;
            const int jm = (j == 0) ? Y - 1 : j - 1, jp = (j + 1 == Y) ? 0 : j + 1;
            for (int k = 0; k < Z; ++k) {
                const int km = (k == 0) ? Z - 1 : k - 1, kp = (k + 1 == Z) ? 0 : k + 1;
                const size_t c = IDX(i, j, k);
                const size_t xm = IDX(im, j, k), xp = IDX(ip, j, k);
                const size_t ym = IDX(i, jm, k), yp = IDX(i, jp, k);
                const size_t zm = IDX(i, j, km), zp = IDX(i, j, kp);
                const double p = P[c], n = N[c], s = S[c];
                const double H = heaviside50(s, sigma_np);
                const double spx = stencil_term(ix, Dpx[c], Dmx[c], P[xm], p, P[xp]);
                const double spy = stencil_term(iy, Dpy[c], Dmy[c], P[ym], p, P[yp]);
                const double spz = stencil_term(iz, Dpz[c], Dmz[c], P[zm], p, P[zp]);
                const double sp = (spx + spy) + spz;
                /* (SP + f*np.multiply(S,P)*(1-P-N) - lambda_np*P*H) * dt        FK_2c.py:58 */
                const double dP = ((sp + (rho * (s * p)) * ((1 - p) - n)) - (lambda_np * p) * H) * dt;
                const double pn = p + dP;
                /* lambda_np * P(new) * H * dt                                    FK_2c.py:62 */
                const double dN = ((lambda_np * pn) * H) * dt;
                const double ssx = stencil_term(ix, Epx[c], Emx[c], S[xm], s, S[xp]);
                const double ssy = stencil_term(iy, Epy[c], Emy[c], S[ym], s, S[yp]);
                const double ssz = stencil_term(iz, Epz[c], Emz[c], S[zm], s, S[zp]);
                const double ss = (ssx + ssy) + ssz;
                /* (SS - lambda_s*S*P(new)) * dt                                  FK_2c.py:70 */
                const double dS = (ss - (lambda_s * s) * pn) * dt;
                Po[c] = pn;
                No[c] = n + dN;
                So[c] = s + dS;
            }
        }
}

/*
 * The FK_2c step as FK_2c.Solver.solve runs it (FK_2c.py:215-216,229-230): P faces rebuilt
 * from the current P,N (necro mask) and the tissue maps, S faces static with ratio 1.
 * Dm* / Em* are scratch (X*Y*Z each) so the faces are built exactly like tgo_faces().
 */
void tgo_fk2c_step_tissue(const double *P, const double *N, const double *S,
                          const double *WM, const double *GM,
                          double th, double th_necro, double Dw, double ratio, double Ds,
                          double rho, double lambda_np, double sigma_np, double lambda_s,
                          double dt, double dx, double dy, double dz,
                          int X, int Y, int Z,
                          double *Dmx, double *Dmy, double *Dmz,
                          double *Emx, double *Emy, double *Emz, int rebuild_static,
                          double *Po, double *No, double *So)
{
    tgo_faces(WM, GM, P, N, th, th_necro, Dw, ratio, X, Y, Z, Dmx, Dmy, Dmz);
    if (rebuild_static)
        tgo_faces(WM, GM, NULL, NULL, th, 0.0, Ds, 1.0, X, Y, Z, Emx, Emy, Emz);
    const double ix = 1 / (dx * dx), iy = 1 / (dy * dy), iz = 1 / (dz * dz);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            const int im = (i == 0) ? X - 1 : i - 1, ip = (i + 1 == X) ? 0 : i + 1;
            const int jm = (j == 0) ? Y - 1 : j - 1, jp = (j + 1 == Y) ? 0 : j + 1;
            for (int k = 0; k < Z; ++k) {
                const int km = (k == 0) ? Z - 1 : k - 1, kp = (k + 1 == Z) ? 0 : k + 1;
                const size_t c = IDX(i, j, k);
                const size_t xm = IDX(im, j, k), xp = IDX(ip, j, k);
                const size_t ym = IDX(i, jm, k), yp = IDX(i, jp, k);
                const size_t zm = IDX(i, j, km), zp = IDX(i, j, kp);
                const double p = P[c], n = N[c], s = S[c];
                const double H = heaviside50(s, sigma_np);
                const double spx = stencil_term(ix, Dmx[xm], Dmx[c], P[xm], p, P[xp]);
                const double spy = stencil_term(iy, Dmy[ym], Dmy[c], P[ym], p, P[yp]);
                const double spz = stencil_term(iz, Dmz[zm], Dmz[c], P[zm], p, P[zp]);
                const double sp = (spx + spy) + spz;
                const double dP = ((sp + (rho * (s * p)) * ((1 - p) - n)) - (lambda_np * p) * H) * dt;
                const double pn = p + dP;
                const double dN = ((lambda_np * pn) * H) * dt;
                const double ssx = stencil_term(ix, Emx[xm], Emx[c], S[xm], s, S[xp]);
                const double ssy = stencil_term(iy, Emy[ym], Emy[c], S[ym], s, S[yp]);
                const double ssz = stencil_term(iz, Emz[zm], Emz[c], S[zm], s, S[zp]);
                const double ss = (ssx + ssy) + ssz;
                const double dS = (ss - (lambda_s * s) * pn) * dt;
                Po[c] = pn;
                No[c] = n + dN;
                So[c] = s + dS;
            }
        }
}

/* Plain sum of a field (the volume tests use a tolerance; np.sum is pairwise, this is a
 * per-row Kahan-free double sum reduced over threads -- agreement ~1e-15 relative). */
double tgo_sum(const double *a, size_t n)
{
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (size_t i = 0; i < n; ++i) s += a[i];
    return s;
}

/*
 * Separable linear resample = scipy.ndimage.zoom(x, f, order=1) as the reference calls it
 * (FK.py:150-151,233-238): output size round(n*f) is chosen by the caller, coordinate map
 * in = out*(n_in-1)/(n_out-1) ("align corners", grid_mode=False), linear weights.
 */
void tgo_zoom_linear(const double *src, int X, int Y, int Z, double *dst, int OX, int OY, int OZ)
{
    const double fx = (OX > 1) ? (double)(X - 1) / (double)(OX - 1) : 0.0;
    const double fy = (OY > 1) ? (double)(Y - 1) / (double)(OY - 1) : 0.0;
    const double fz = (OZ > 1) ? (double)(Z - 1) / (double)(OZ - 1) : 0.0;
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < OX; ++i)
        for (int j = 0; j < OY; ++j) {
            const double xi = i * fx, yj = j * fy;
            int i0 = (int)floor(xi), j0 = (int)floor(yj);
            if (i0 > X - 2) i0 = (X >= 2) ? X - 2 : 0;
            if (j0 > Y - 2) j0 = (Y >= 2) ? Y - 2 : 0;
            const double tx = xi - i0, ty = yj - j0;
            const int i1 = (X >= 2) ? i0 + 1 : i0, j1 = (Y >= 2) ? j0 + 1 : j0;
            for (int k = 0; k < OZ; ++k) {
                const double zk = k * fz;
                int k0 = (int)floor(zk);
                if (k0 > Z - 2) k0 = (Z >= 2) ? Z - 2 : 0;
                const double tz = zk - k0;
                const int k1 = (Z >= 2) ? k0 + 1 : k0;
                double acc = 0.0;
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b)
                        for (int c2 = 0; c2 < 2; ++c2) {
                            const double w = (a ? tx : 1 - tx) * (b ? ty : 1 - ty) * (c2 ? tz : 1 - tz);
                            acc += w * src[IDX(a ? i1 : i0, b ? j1 : j0, c2 ? k1 : k0)];
                        }
                dst[((size_t)i * OY + j) * OZ + k] = acc;
            }
        }
}
