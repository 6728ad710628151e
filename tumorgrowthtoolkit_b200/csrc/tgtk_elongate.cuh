// Batched 3x3 symmetric eigen-elongation: the preprocessing FK_DTI_Solver runs when
// diffusionEllipsoidScaling != 1 -- reference TumorGrowthToolkit/FK_DTI/tools.py:109-164, a float32
// torch.linalg.eigh over ~9 M voxels on the CPU.  Per tensor A (rounded to float32 like the reference's input):
//     lambda, v  = largest eigenvalue and its unit eigenvector
//     adj        = (lambda*s - lambda) / 2                         tools.py:123-129
//     e_other'   = e_other - adj + (total - sum(e_adjusted)) / 3   tools.py:134-139; the bracket is 2*adj because
//                                                                  e_adjusted still holds the unscaled maximum
//     e_max'     = lambda * s                                      tools.py:142
//     A'         = V diag(e') V^T = A - (adj/3) (I - v v^T) + (lambda*s - lambda) v v^T     tools.py:145
// so only the top eigenpair is needed.  It is computed in float64 (trigonometric eigenvalues, eigenvector from
// the largest cross product of the rows of A - lambda I); the reference's own float32 LAPACK noise (1e-7
// relative) is the agreement one can ask for.  Where the largest eigenvalue is degenerate the reference's
// eigenvector is whatever LAPACK returns: for exactly diagonal tensors (isotropic grey matter, background) that
// is deterministic -- ssteqr's selection sort of the diagonal, then argmax's first index on ties -- and is
// reproduced here; for non-diagonal tensors with a degenerate top pair any vector of the eigenplane is used.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tgtk {

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// unit vector spanning the null space of the symmetric rank-2 matrix with rows r0, r1, r2; false if the
// matrix has (numerically) rank < 2
__device__ __forceinline__ bool null_vector(const double r0[3], const double r1[3], const double r2[3], double scale2, double v[3])
{
    double c[3][3];
    cross3(r0, r1, c[0]);
    cross3(r0, r2, c[1]);
    cross3(r1, r2, c[2]);
    int best = 0;
    double nb = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double n = c[i][0] * c[i][0] + c[i][1] * c[i][1] + c[i][2] * c[i][2];
        if (n > nb) { nb = n; best = i; }
    }
    if (!(nb > 1e-24 * scale2 * scale2)) return false;      // |cross|^2 ~ scale^4
    const double inv = rsqrt(nb);
    v[0] = c[best][0] * inv; v[1] = c[best][1] * inv; v[2] = c[best][2] * inv;
    return true;
}

__device__ __forceinline__ void top_eigenpair(double a00, double a01, double a02, double a11, double a12, double a22,
                                              double &lam, double v[3])
{
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    if (p1 == 0.0) {
        // diagonal: LAPACK sorts the diagonal ascending by selection sort (strict <), swapping the eigenvector
        // columns along; torch.argmax then takes the first of equal maxima
        double d[3] = {a00, a11, a22};
        int ax[3] = {0, 1, 2};
        for (int i = 0; i < 2; ++i) {
            int k = i;
            for (int j = i + 1; j < 3; ++j)
                if (d[j] < d[k]) k = j;
            if (k != i) {
                const double t = d[k]; d[k] = d[i]; d[i] = t;
                const int u = ax[k]; ax[k] = ax[i]; ax[i] = u;
            }
        }
        int im = 0;
        if (d[1] > d[im]) im = 1;
        if (d[2] > d[im]) im = 2;
        lam = d[im];
        v[0] = (ax[im] == 0); v[1] = (ax[im] == 1); v[2] = (ax[im] == 2);
        return;
    }
    const double q = (a00 + a11 + a22) / 3.0;
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    const double p = sqrt(p2 / 6.0);
    const double det = b00 * (b11 * b22 - a12 * a12) - a01 * (a01 * b22 - a12 * a02) + a02 * (a01 * a12 - b11 * a02);
    double r = det / (2.0 * p * p * p);
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) / 3.0;
    lam = q + 2.0 * p * cos(phi);
    const double scale2 = p2 + q * q;
    {
        const double r0[3] = {a00 - lam, a01, a02}, r1[3] = {a01, a11 - lam, a12}, r2[3] = {a02, a12, a22 - lam};
        if (null_vector(r0, r1, r2, scale2, v)) return;
    }
    // degenerate top pair: any unit vector of the plane orthogonal to the smallest eigenvalue's eigenvector,
    // i.e. any non-zero row of A - lambda_min I
    const double lmin = q + 2.0 * p * cos(phi + 2.0943951023931953);
    const double r0[3] = {a00 - lmin, a01, a02}, r1[3] = {a01, a11 - lmin, a12}, r2[3] = {a02, a12, a22 - lmin};
    const double n0 = r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2];
    const double n1 = r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2];
    const double n2 = r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2];
    const double *rb = (n0 >= n1 && n0 >= n2) ? r0 : (n1 >= n2 ? r1 : r2);
    const double nb = fmax(n0, fmax(n1, n2));
    const double inv = (nb > 0.0) ? rsqrt(nb) : 0.0;
    v[0] = rb[0] * inv; v[1] = rb[1] * inv; v[2] = rb[2] * inv;
}

// in: n tensors of 9 float64 (C order 3x3), out: n tensors of 9 float32.  A block stages its 256 tensors
// through shared memory so that both the loads and the stores of the 72-byte records are coalesced.
__global__ void __launch_bounds__(256) elongate_kernel(const double *__restrict__ in, float *__restrict__ out, int64_t n,
                                                       double s)
{
    __shared__ double stage[256 * 9];
    const int64_t first = (int64_t)blockIdx.x * 256;
    const int count = (int)min((int64_t)256, n - first);
    if (count <= 0) return;
    for (int i = threadIdx.x; i < count * 9; i += 256) stage[i] = in[first * 9 + i];
    __syncthreads();
    float o[9];
    if ((int)threadIdx.x < count) {
        const double *a = stage + threadIdx.x * 9;
        // the reference casts to float32 before eigh (tools.py:114) and eigh reads the lower triangle
        const double a00 = (double)(float)a[0], a11 = (double)(float)a[4], a22 = (double)(float)a[8];
        const double a01 = (double)(float)a[3], a02 = (double)(float)a[6], a12 = (double)(float)a[7];
        double lam, v[3];
        top_eigenpair(a00, a01, a02, a11, a12, a22, lam, v);
        const double grow = lam * s - lam;          // e_max' - e_max
        const double shrink = -(grow / 2.0) / 3.0;  // e_other' - e_other
        const double w = grow - shrink;             // coefficient of v v^T
        const double A[9] = {a00, a01, a02, a01, a11, a12, a02, a12, a22};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) o[i * 3 + j] = (float)(A[i * 3 + j] + (i == j ? shrink : 0.0) + w * v[i] * v[j]);
    }
    __syncthreads();
    float *fs = reinterpret_cast<float *>(stage);
    if ((int)threadIdx.x < count)
        for (int i = 0; i < 9; ++i) fs[threadIdx.x * 9 + i] = o[i];
    __syncthreads();
    for (int i = threadIdx.x; i < count * 9; i += 256) out[first * 9 + i] = fs[i];
}

}  // namespace tgtk
