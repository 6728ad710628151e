// Stand-alone probe of the TMA pieces tgtk_kernels_tma.cuh uses: a tensor map over a pitched (X,Y,Z) array, one
// cp.async.bulk.tensor.3d box with negative / out-of-range start coordinates, mbarrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tma_probe tma_probe.cu
//   ./tma_probe <elem: 8|4> <box z> <box y> <cz> <cy> <cx> <desc: param|global> <form: cluster|cta> <l2promo 0..3> <rank 3|2>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../tumorgrowthtoolkit_b200/csrc/tgtk_kernels_tma.cuh"
using namespace tgtk;

struct OneMap { CUtensorMap m; };

__global__ void probe(const __grid_constant__ OneMap maps, const CUtensorMap *gmap, unsigned char *out, int bytes, int cz, int cy, int cx,
                      int use_global, int cta_form, int rank)
{
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 127u) & ~127u;
    unsigned char *sm = raw + (base - smem_u32(raw));
    const uint32_t bar = base, dst = base + 128;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *m = use_global ? gmap : &maps.m;
        mbar_expect_tx(bar, bytes);
        if (rank == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                         "l"(m), "r"(bar), "r"(cz), "r"(cy) : "memory");
        else if (cta_form)
            asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                         "l"(m), "r"(bar), "r"(cz), "r"(cy), "r"(cx) : "memory");
        else tma_load_3d(dst, m, bar, cz, cy, cx);
    }
    mbar_wait(bar, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[128 + i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char **argv)
{
    const int es = argc > 1 ? atoi(argv[1]) : 8, bz = argc > 2 ? atoi(argv[2]) : 34, by = argc > 3 ? atoi(argv[3]) : 10;
    const int cz = argc > 4 ? atoi(argv[4]) : -1, cy = argc > 5 ? atoi(argv[5]) : -1, cx = argc > 6 ? atoi(argv[6]) : 2;
    const int use_global = argc > 7 && !strcmp(argv[7], "global"), cta_form = argc > 8 && !strcmp(argv[8], "cta");
    const int promo = argc > 9 ? atoi(argv[9]) : 2, rank = argc > 10 ? atoi(argv[10]) : 3;
    const int X = 5, Y = 21, Z = 139, sy = 140, sx = Y * sy;
    std::vector<double> h8((size_t)X * sx);
    std::vector<float> h4((size_t)X * sx);
    for (int i = 0; i < X; ++i) for (int j = 0; j < Y; ++j) for (int k = 0; k < sy; ++k) {
        const double v = (k < Z) ? i * 1e4 + j * 1e2 + k * 0.5 : -7.0;
        h8[(size_t)i * sx + j * sy + k] = v; h4[(size_t)i * sx + j * sy + k] = (float)v;
    }
    void *d; unsigned char *o;
    CK(cudaMalloc(&d, h8.size() * es));
    CK(cudaMemcpy(d, es == 8 ? (void *)h8.data() : (void *)h4.data(), h8.size() * es, cudaMemcpyHostToDevice));
    const int bytes = bz * by * es;
    CK(cudaMalloc(&o, bytes));
    void *fn = nullptr; cudaDriverEntryPointQueryResult st;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st));
    typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    OneMap mp;
    const cuuint64_t dims[3] = {(cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X}, strides[2] = {(cuuint64_t)sy * es, (cuuint64_t)sx * es};
    const cuuint32_t box[3] = {(cuuint32_t)bz, (cuuint32_t)by, 1}, est[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(&mp.m, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, est,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("es %d box %dx%d at (%d,%d,%d) desc %s form %s promo %d rank %d: encode -> %d, ", es, bz, by, cz, cy, cx, use_global ? "global" : "param",
           cta_form ? "cta" : "cluster", promo, rank, (int)r);
    CUtensorMap *gm;
    CK(cudaMalloc(&gm, sizeof(CUtensorMap)));
    CK(cudaMemcpy(gm, &mp.m, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    probe<<<1, 128, 128 + 128 + bytes>>>(mp, gm, o, bytes, cz, cy, cx, use_global, cta_form, rank);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<unsigned char> res(bytes);
    CK(cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int j = 0; j < by; ++j) for (int k = 0; k < bz; ++k) {
        const int yy = cy + j, zz = cz + k, xx = rank == 2 ? 0 : cx;
        const double want = (yy < 0 || zz < 0 || yy >= Y || zz >= Z) ? 0.0 : xx * 1e4 + yy * 1e2 + zz * 0.5;
        const double got = es == 8 ? ((double *)res.data())[j * bz + k] : (double)((float *)res.data())[j * bz + k];
        if (got != want) ++bad;
    }
    printf(", %d mismatches\n", bad);
    return 0;
}
