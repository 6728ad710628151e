// C ABI of libtgtk_b200 (see include/tgtk_b200.h).  Host-side driver of the time loop:
// device memory, uploads/downloads, launch sequencing, snapshots, status.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tgtk_b200.h"
#include "tgtk_common.cuh"
#include "tgtk_kernels_v1.cuh"
#include "tgtk_kernels_v2.cuh"
#include "tgtk_kernels_v3.cuh"
#include "tgtk_kernels_v5.cuh"
#include "tgtk_kernels_tb.cuh"
#include "tgtk_kernels_tma.cuh"
#include "tgtk_kernels_sparse.cuh"
#include "tgtk_elongate.cuh"

using namespace tgtk;

// steps per CUDA-graph replay = slots of the partial-sum ring of the host-scheduled mode
static const int kGraphSteps = 32;

static thread_local std::string g_err;

static int fail(const char *what, const char *detail)
{
    g_err = std::string(what) + ": " + detail;
    return 1;
}

#define CU(call)                                                         \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return fail(#call, cudaGetErrorString(e__)); \
    } while (0)

struct tgtk_sim {
    tgtk_params p;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t n = 0;          // voxels
    size_t elem = 8;       // sizeof state type
    int64_t sx = 0, sy = 0;
    int nfields = 1, ncoef = 0;
    void *state[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    void *coef[12] = {nullptr};
    double *stage = nullptr;  // dense float64 staging for uploads / downloads
    double *resample_out = nullptr;   // device output of tgtk_sim_download_resampled
    size_t resample_cap = 0;
    double *resample_async[3] = {nullptr, nullptr, nullptr};   // per-field outputs of the asynchronous downloads
    uint8_t *seg_out = nullptr;
    size_t resample_async_cap[3] = {0, 0, 0}, seg_cap = 0;
    void *initial[3] = {nullptr, nullptr, nullptr};  // state at prepare time (tgtk_sim_reset)
    double *wm = nullptr, *gm = nullptr, *pc = nullptr, *pn = nullptr;
    int faces_what = 0;  // tgtk_faces_host: 0 D_minus, 1 WM_tilda, 2 GM_tilda
    Ctrl *ctrl = nullptr;
    Ctrl *ctrl_host = nullptr;  // pinned mirror
    double *partials = nullptr;
    double *volumes = nullptr;
    int volumes_cap = 0;
    dim3 grid, block;
    unsigned nblocks = 0;
    int t_enqueued = 0;
    bool prepared = false;
    int variant = 0;       // 0: auto (default), 1: baseline, 2: marching via L1, 3: marching via shared memory
    int lx = 1;            // planes per block of the marching kernels
    double narrow_penalty = 0.25;   // 16-wide v4 tiles are chosen when they waste this much less than 32-wide
    int v4_threads = 128;  // fk2c_step_v4 block size: 128 (32x4 threads, 32x8 voxels) or 256 (32x8 threads, 32x16 voxels)
    int pf = 1;            // FK_2c staged kernel: prefetch distance in planes (1: two register sets, 4 blocks/SM; 2: three sets)
    int force_tz = 0;      // 0: pick the tile width by lane waste
    int force_lx = 0;      // 0: planes per segment from the wave model (configure_launch)
    int l2pf = -1;         // v4 kernels: L2 prefetch distance in planes (TGTK_L2PF); 0 = off, -1 = auto:
                           // 2 planes when the arrays a step touches do not fit in L2, else off (measured:
                           // the requests only cost time when the working set is L2 resident)
    int l2_bytes = 0;      // cudaDevAttrL2CacheSize
    int pdl = -1;          // host-scheduled launches overlap their prologue with the previous step's tail (TGTK_PDL):
                           // 1 on, 0 off, -1 auto = on when the step's arrays fit in L2 (short launches, where the
                           // launch gap matters; measured slightly negative for the long HBM-streaming launches)
    bool use_graph = true;
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};   // by parity of the first step
    cudaGraphExec_t slab_graph[2] = {nullptr, nullptr};   // tgtk_sim_run_slab: one block, by parity
    void *slab_key_comm = nullptr; int slab_key_halo = 0, slab_key_owned = 0; double *slab_key_vol = nullptr;
    // peer-to-peer slab mode: state buffers from cudaMalloc (exportable with CUDA IPC), the neighbours'
    // buffers and flag words mapped into this process
    bool ipc = false;
    int *slab_flags = nullptr;             // [0..1] done epoch from prev/next, [2..3] pushed epoch, [4] error, [5..6] own counters
    void *peer_state[2][2][3] = {};        // [prev|next][ping-pong][field]
    int *peer_flags[2] = {nullptr, nullptr};
    void *peer_opened[2][7] = {};          // pointers to close with cudaIpcCloseMemHandle
    // fused slab mode (tgtk_sim_slab_attach): planes [1, 1 + slab_owned) are computed, the boundary planes are stored
    // into the neighbours' ghost planes from inside the step kernels, volumes are all-reduced per chunk
    bool slab_on = false;
    int slab_owned = 0, slab_prev_owned = 0;
    void *slab_comm = nullptr;             // NCCL communicator for the per-chunk all-reduce (nullptr: one rank)
    double *chunkvol = nullptr;            // the chunk's local volumes, summed over the ranks in place
    bool host_sched = false;   // no stop criterion: parity baked into launches (see StepArgs)
    // A stop criterion is active: run host-scheduled all the same, chunk by chunk, behind a checkpoint of the
    // state; chunk_finish_kernel runs the stop tests afterwards and a trip is replayed from the checkpoint up to
    // the stopping step (fetch_ctrl).  TGTK_SPECULATE=0 selects the device-scheduled launches instead.
    bool spec = false, replaying = false;
    bool ctrl_valid = false;   // the pinned mirror of Ctrl is current (nothing was enqueued since it was fetched)
    void *ckpt[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;   // around a replay
    float replay_ms = 0.f;
    int sum_x0 = 0, sum_x1 = -1;     // planes counted in the volume (-1: all)
    int cur_par = 0, cur_slot = 0;   // launch-time values of StepArgs::par / slot
    // fp32 rows are padded to an even pitch; for odd Z the pad element mirrors z = 0 (tgtk_kernels_v5.cuh).
    // Only the v5 kernels keep the mirror current: anything else that writes state marks it stale.
    bool pads_dirty = true;
    int minb = 0;          // FK_2c staged kernel: resident blocks per SM the register budget targets
    unsigned partials_cap = 0;
    // temporal blocking (tgtk_kernels_tb.cuh): two steps per launch where a kernel exists (FK, FK_DTI) and the
    // run is host-scheduled or speculative.  TGTK_TB: 1 on, 0 off, -1 auto (fp64 boxes of more than twice the L2)
    int closed_skip = 1;   // exact per-warp shortcuts of the two-voxel FK / FK_2c kernels (TGTK_CLOSED_SKIP=0: off)
    // Sparse launches (tgtk_kernels_sparse.cuh): warp tiles outside the tissue that hold no tumour are never loaded, computed
    // or stored, and the grid is a work list over the rest.  TGTK_SPARSE: 1 on, 0 off, -1 auto (FK_2c boxes beyond L2 with at
    // least 10 % dead tiles).  Rebuilt by tgtk_sim_run whenever a state or tissue array was uploaded since.
    int sparse = -1;
    bool sp_dirty = true, sp_active = false, sp_forbid = false;
    uint32_t *sp_live = nullptr;
    int4 *sp_work = nullptr;
    size_t sp_live_cap = 0, sp_work_cap = 0;
    int sp_items = 0, sp_nzt = 0, sp_nyt = 0;
    int64_t sp_live_voxels = 0;      // voxels in live warp tiles
    bool sp_waves_set = false;
    double sp_waves = 1.0;           // work items per resident-block slot (TGTK_SP_WAVES); measured best: one (every item pays a 3 us prologue)
    int sp_carveout = -1;
    int64_t sp_min_voxels = 1 << 18;  // TGTK_SPARSE_MIN_VOXELS
    int sp_ahead = 0;                // TGTK_SP_AHEAD: prologue prefetch for a later item (measured: no gain), -1 = resident blocks of the device
    double sp_split = 0.5;           // TGTK_SP_SPLIT: share of a slot's work in its first item
    unsigned long long *sp_trace = nullptr;      // tgtk_sim_sparse_trace: per-item timestamps of the next launches
    int sp_async = 0, sp_minb = 4, sp_per_sm = 0;   // TGTK_SP_ASYNC=2: own voxels staged by cp.async two planes ahead; TGTK_SP_MINB
    int zigzag = 0;        // serpentine sweep of the two-voxel kernels (TGTK_ZIGZAG=1): measured +-1 %, off by default
    int tb = -1;
    int tb_rows = 18;      // thread rows of the two-step kernel's block: 18 (16 output rows, 576 threads) or 10 (8, 320)
    dim3 tb_grid, tb_block;
    int tb_lx = 0;
    unsigned tb_nblocks = 0, ring_stride = 0, chunk_tb_mask = 0;
    // variant 7 (tgtk_kernels_tma.cuh): tensor maps of the five arrays an FK_2c launch reads, per ping-pong parity
    TmaSet tma[2];
    bool tma_ready = false;
    int tma_rows = 4, tma_stages = 3, tma_minb = 3;   // thread rows of the consumer tile (4: 32x8 voxels, 8: 32x16), pipeline depth
    std::vector<void *> snaps;     // one allocation per snapshot: nfields * n * elem
    std::vector<int> snap_steps;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int64_t launches = 0;       // step-kernel launches (graph nodes included)
    int graph_launches = kGraphSteps;   // step-kernel nodes of one captured chunk
    int slab_block_launches = 0;        // ... of one captured slab block
    int64_t aux_launches = 0;   // chunk reductions and other helper kernels, as enqueued or captured
};

// Device memory comes from the stream-ordered allocator: unlike cudaMalloc / cudaFree it never
// synchronises the device, so creating or destroying one simulation does not stall the step
// kernels of the others that an ensemble keeps in flight on neighbouring streams.
static cudaError_t sim_malloc(tgtk_sim *s, void **ptr, size_t bytes)
{
    return cudaMallocAsync(ptr, bytes ? bytes : 8, s->stream);
}

static cudaError_t sim_free(tgtk_sim *s, void *ptr)
{
    return ptr ? cudaFreeAsync(ptr, s->stream) : cudaSuccess;
}

// pinned host mirrors of Ctrl are recycled: cudaMallocHost costs far more than a whole small solve
static std::mutex g_pin_mu;
static std::vector<Ctrl *> g_pin_free;

static Ctrl *pinned_ctrl_get()
{
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) {
            Ctrl *c = g_pin_free.back();
            g_pin_free.pop_back();
            return c;
        }
    }
    Ctrl *c = nullptr;
    if (cudaMallocHost(&c, sizeof(Ctrl)) != cudaSuccess) return nullptr;
    return c;
}

static void pinned_ctrl_put(Ctrl *c)
{
    if (!c) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_free.push_back(c);
}

static void keep_pool_memory(int device)
{
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;     // cache freed blocks instead of returning them to the OS
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[device] = true;
}

// ------------------------------------------------------------------------------------------
// layout conversion and coefficient set-up kernels
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_kernel(const double *__restrict__ src, T *__restrict__ dst, int X, int Y, int Z,
                            int64_t sx, int64_t sy)
{
    const int64_t n = (int64_t)X * Y * Z;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(q % Z);
        const int64_t r = q / Z;
        const int j = (int)(r % Y);
        const int i = (int)(r / Y);
        dst[i * sx + j * sy + k] = (T)src[q];
    }
}

template <typename T>
__global__ void unpack_kernel(const T *__restrict__ src, double *__restrict__ dst, int X, int Y, int Z,
                              int64_t sx, int64_t sy)
{
    const int64_t n = (int64_t)X * Y * Z;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(q % Z);
        const int64_t r = q / Z;
        const int j = (int)(r % Y);
        const int i = (int)(r / Y);
        dst[q] = (double)src[i * sx + j * sy + k];
    }
}

// odd-Z fp32 boxes: copy element z = 0 of every row into the pad element z = Z (tgtk_kernels_v5.cuh)
template <typename T> __global__ void mirror_pad_kernel(T *__restrict__ arr, int X, int Y, int Z, int64_t sx, int64_t sy)
{
    const int64_t n = (int64_t)X * Y;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = (q / Y) * sx + (q % Y) * sy;
        arr[d + Z] = arr[d];
    }
}

// k(c) and b(c) of the compact FK / FK_2c kernels, evaluated in float64 from the float64
// tissue maps and rounded once to the state type.
template <typename T>
__global__ void prep_compact_kernel(const double *__restrict__ wm, const double *__restrict__ gm, T *__restrict__ kc,
                                    T *__restrict__ bc, double th, double ratio, int X, int Y, int Z, int64_t sx,
                                    int64_t sy)
{
    const int64_t n = (int64_t)X * Y * Z;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(q % Z);
        const int64_t r = q / Z;
        const int j = (int)(r % Y);
        const int i = (int)(r / Y);
        const double w = wm[q], g = gm[q];
        const bool tissue = (w + g >= th);
        const int64_t d = i * sx + j * sy + k;
        kc[d] = tissue ? (T)(0.5 * (w + g / ratio)) : -Big<T>::v;
        if (bc) bc[d] = tissue ? (T)(0.5 * (w + g)) : -Big<T>::v;
    }
}

// D_minus_{x,y,z} in the reference's own rounding order (FK/FK.py:12-26): used by the
// bit-exact FACES model.  logical_or: numpy bool semantics of '+' (SURVEY.md 7).
template <typename T>
__global__ void prep_faces_kernel(const double *__restrict__ wm, const double *__restrict__ gm,
                                  const double *__restrict__ pc, const double *__restrict__ pn, T *__restrict__ fx,
                                  T *__restrict__ fy, T *__restrict__ fz, double th, double th_necro, double Dw,
                                  double ratio, int logical_or, int what, int X, int Y, int Z, int64_t sx, int64_t sy)
{
    const int64_t n = (int64_t)X * Y * Z;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(q % Z);
        const int64_t r = q / Z;
        const int j = (int)(r % Y);
        const int i = (int)(r / Y);
        const int ip = (i + 1 == X) ? 0 : i + 1, jp = (j + 1 == Y) ? 0 : j + 1, kp = (k + 1 == Z) ? 0 : k + 1;
        const int64_t nb[3] = {((int64_t)ip * Y + j) * Z + k, ((int64_t)i * Y + jp) * Z + k, ((int64_t)i * Y + j) * Z + kp};
        T *out[3] = {fx, fy, fz};
        const double w = wm[q], g = gm[q];
        auto plus = [&](double a, double b) {
            return logical_or ? ((a != 0.0 || b != 0.0) ? 1.0 : 0.0) : __dadd_rn(a, b);
        };
        // FK_2c/FK_2c.py:13-15: both ends of the face need (P+N) <= th_necro
        const bool here = plus(w, g) >= th && (!pc || __dadd_rn(pc[q], pn[q]) <= th_necro);
        for (int a = 0; a < 3; ++a) {
            const double w2 = wm[nb[a]], g2 = gm[nb[a]];
            double wt = 0.0, gt = 0.0;
            if (here && plus(w2, g2) >= th && (!pc || __dadd_rn(pc[nb[a]], pn[nb[a]]) <= th_necro)) {
                wt = plus(w2, w) / 2;
                gt = plus(g2, g) / 2;
            }
            const double v = (what == 1) ? wt : (what == 2) ? gt : __dmul_rn(Dw, __dadd_rn(wt, __ddiv_rn(gt, ratio)));
            out[a][i * sx + j * sy + k] = (T)v;
        }
    }
}

// ------------------------------------------------------------------------------------------
static inline int grid_1d(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 148 * 16); }

static bool exceeds_l2(const tgtk_sim *s)
{
    const tgtk_params &p = s->p;
    const double touched = (double)(2 * s->nfields + s->ncoef) * (double)p.X * (double)s->sx * (double)s->elem;
    return touched > 0.8 * (double)s->l2_bytes;
}
static int effective_l2pf(const tgtk_sim *s)
{
    // sparse launches only touch the live tiles: what has to fit in L2 is that share of the arrays (FK_DTI on the atlas box:
    // 143 MB dense, 79 MB live -> no prefetch, 18.0 us per step against 18.4 with it)
    bool beyond = exceeds_l2(s);
    if (beyond && s->sp_active && s->n > 0) {
        const tgtk_params &p = s->p;
        const double touched = (double)(2 * s->nfields + s->ncoef) * (double)p.X * (double)s->sx * (double)s->elem;
        const double owned = s->slab_on ? (double)s->slab_owned * p.Y * p.Z : (double)s->n;
        beyond = touched * ((double)s->sp_live_voxels / owned) > 0.8 * (double)s->l2_bytes;
    }
    const int l2pf = (s->l2pf < 0) ? (beyond ? 2 : 0) : s->l2pf;
    return (l2pf > 0 && l2pf < s->p.X) ? l2pf : 0;
}
static bool effective_pdl(const tgtk_sim *s)
{
    if (!s->host_sched && !s->spec) return false;
    return (s->pdl < 0) ? !exceeds_l2(s) : (s->pdl != 0);
}

template <typename T> static void fill_args(const tgtk_sim *s, StepArgs<T> &a)
{
    const tgtk_params &p = s->p;
    a.X = p.X; a.Y = p.Y; a.Z = p.Z;
    a.sx = s->sx; a.sy = s->sy;
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < 3; ++f) a.buf[b][f] = (T *)s->state[b][f];
    for (int c = 0; c < 12; ++c) a.coef[c] = (const T *)s->coef[c];
    const double ix = 1.0 / (p.h[0] * p.h[0]), iy = 1.0 / (p.h[1] * p.h[1]), iz = 1.0 / (p.h[2] * p.h[2]);
    const bool fold = (p.model == TGTK_MODEL_FK || p.model == TGTK_MODEL_FK_2C);
    a.cx = (T)(fold ? p.Dw * ix : ix);
    a.cy = (T)(fold ? p.Dw * iy : iy);
    a.cz = (T)(fold ? p.Dw * iz : iz);
    a.ex = (T)(p.Ds * ix); a.ey = (T)(p.Ds * iy); a.ez = (T)(p.Ds * iz);
    if (p.model == TGTK_MODEL_DTI_FULL) {   // mixed-derivative weights 2/(4 h_a h_b)
        a.ex = (T)(2.0 / (4.0 * p.h[0] * p.h[1]));
        a.ey = (T)(2.0 / (4.0 * p.h[0] * p.h[2]));
        a.ez = (T)(2.0 / (4.0 * p.h[1] * p.h[2]));
    }
    const T half = (T)Relu2<T>::scale;
    a.hx = half * a.cx; a.hy = half * a.cy; a.hz = half * a.cz;
    a.gx = half * a.ex; a.gy = half * a.ey; a.gz = half * a.ez;
    a.dt = (T)p.dt; a.rho = (T)p.rho;
    a.th_necro = (T)p.th_necro; a.lambda_np = (T)p.lambda_np; a.sigma_np = (T)p.sigma_np; a.lambda_s = (T)p.lambda_s;
    a.vol_scale_p = p.h[0] * p.h[1] * p.h[2];
    a.vol_scale_n = 1.0;
    a.stop_volume = p.stopping_volume;
    a.small_volume = p.small_volume;
    a.ctrl = s->ctrl;
    a.partials = s->partials;
    a.ring_stride = s->ring_stride;
    a.volumes = s->volumes;
    a.volumes_cap = s->volumes_cap;
    a.host_sched = (s->host_sched || s->spec) ? 1 : 0;
    a.spec = (s->spec && !s->replaying) ? 1 : 0;
    a.par = s->cur_par;
    a.slot = s->cur_slot;
    a.sum_x0 = s->sum_x0;
    a.sum_x1 = (s->sum_x1 < 0) ? p.X : s->sum_x1;
    a.l2pf = effective_l2pf(s);
    a.x_lo = 0;
    a.x_hi = p.X;
    a.zigzag = s->zigzag;
    a.skip = s->closed_skip;
    a.slab = 0;
    if (s->slab_on) {
        a.slab = 1;
        a.x_lo = 1;
        a.x_hi = 1 + s->slab_owned;
        for (int b = 0; b < 2; ++b)
            for (int f = 0; f < 3; ++f) {
                a.peer_lo[b][f] = (T *)s->peer_state[0][b][f];
                a.peer_hi[b][f] = (T *)s->peer_state[1][b][f];
            }
        // my plane 1 is the previous rank's upper ghost (its plane 1 + prev_owned); my plane `owned` is the next rank's plane 0
        a.off_lo = (uint32_t)((int64_t)s->slab_prev_owned * s->sx);
        a.off_hi = (uint32_t)(-(int64_t)s->slab_owned * s->sx);
        a.flags = s->slab_flags;
        a.flag_lo = s->peer_flags[0] + kSlabFromNext;     // I am the previous rank's "next"
        a.flag_hi = s->peer_flags[1] + kSlabFromPrev;
    }
}

// variant 0 = auto = the staged kernels with two voxels per thread (v4), fastest for every model and
// dtype measured on B200; 1, 2, 3 select the earlier families (cross-checks)
static bool has_v5(const tgtk_sim *s)
{
    const int m = s->p.model;
    return s->p.dtype == TGTK_F32 && (m == TGTK_MODEL_FK || m == TGTK_MODEL_FK_2C || m == TGTK_MODEL_DTI);
}

static bool slab_kernels(const tgtk_sim *s);

// 7 = FK_2c with TMA tile staging and a producer warp (tgtk_kernels_tma.cuh); needs 16-byte global strides
static bool tma_ok(const tgtk_sim *s)
{
    return s->p.model == TGTK_MODEL_FK_2C && ((s->sy * s->elem) % 16) == 0 && (uint64_t)s->p.X * (uint64_t)s->sx < (1ull << 31);
}

// 5 / 6 = the fp32 z-pair kernels with four / two voxels per thread (tgtk_kernels_v5.cuh); fp64 has no
// such family and runs the two-voxel kernels (4) instead.  Auto (0): measured on B200 the
// z-pair kernels win for FK_DTI (four arrays streamed per voxel, one through shared memory: 64-bit
// accesses halve the load count) and lose for FK / FK_2c, whose tiles carry the coefficient as well and
// pay for the scalar left / right reads with shared-memory bank conflicts.
static int effective_variant(const tgtk_sim *s)
{
    if (s->slab_on) return 4;              // the in-kernel halo push lives in the two-voxel kernels
    if (s->variant == 0) {
        if (!has_v5(s)) return 4;
        if (s->p.model == TGTK_MODEL_DTI) return s->sp_active ? 4 : 6;      // sparse launches exist for the two-voxel kernels only, and win
        // FK / FK_2c fp32: the two-voxel kernels.  Round 2 moved the z-pair kernels' left / right neighbours to warp shuffles,
        // which makes them 13 % faster than the two-voxel FK kernel on all-tissue random boxes -- but on the atlas, where 60 %
        // of the box is background, the two-voxel kernel's closed-warp shortcut wins (281 vs 267 GLUPS): they stay opt-in
        return 4;
    }
    if (s->variant == 7) return tma_ok(s) ? 7 : 4;
    if (s->variant >= 5) return has_v5(s) ? s->variant : 4;
    return s->variant;
}
static bool zpair_variant(const tgtk_sim *s) { const int v = effective_variant(s); return v == 5 || v == 6; }

// refresh the z = 0 mirrors in the pad column of the current state and of the static arrays
static int refresh_pads(tgtk_sim *s)
{
    const tgtk_params &p = s->p;
    s->pads_dirty = false;
    if (s->sy == p.Z || p.dtype != TGTK_F32) return 0;
    const int cur = s->t_enqueued & 1;
    const int g = grid_1d((size_t)p.X * p.Y);
    for (int f = 0; f < s->nfields; ++f)
        mirror_pad_kernel<float><<<g, 256, 0, s->stream>>>((float *)s->state[cur][f], p.X, p.Y, p.Z, s->sx, s->sy);
    for (int c = 0; c < s->ncoef; ++c)
        mirror_pad_kernel<float><<<g, 256, 0, s->stream>>>((float *)s->coef[c], p.X, p.Y, p.Z, s->sx, s->sy);
    CU(cudaGetLastError());
    s->aux_launches += s->nfields + s->ncoef;
    return 0;
}

// v4 launches go through cudaLaunchKernelEx so that host-scheduled steps can carry the programmatic
// stream serialization attribute (PDL): step t+1 is scheduled while step t drains, runs its index set-up
// and static-coefficient loads, and blocks in griddepcontrol.wait until step t has completed and flushed.
template <typename... KA, typename... A> static cudaError_t launch_v4(tgtk_sim *s, void (*kern)(KA...), A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = s->grid;
    cfg.blockDim = s->block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (effective_pdl(s) && s->cur_slot > 0) ? 1 : 0;   // slot 0 follows a non-step node
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---- variant 7: tensor maps and launch -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {   // the driver entry point through the runtime: the library does not link libcuda
        void *q = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)q;
    }
    return fn;
}

// (X, Y, Z) array with pitches (sx, sy) as a rank-3 tensor {Z, Y, X}; box = {bz, by, 1} elements; outside -> zeros
static int make_map(tgtk_sim *s, CUtensorMap *m, void *base, int bz, int by)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail("cuTensorMapEncodeTiled", "driver entry point not available");
    const cuuint64_t dims[3] = {(cuuint64_t)s->p.Z, (cuuint64_t)s->p.Y, (cuuint64_t)s->p.X};
    const cuuint64_t strides[2] = {(cuuint64_t)s->sy * s->elem, (cuuint64_t)s->sx * s->elem};
    const cuuint32_t box[3] = {(cuuint32_t)bz, (cuuint32_t)by, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(m, s->elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           by > 1 && bz > 4 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled", "encoding failed");
    return 0;
}

static int build_tma(tgtk_sim *s)
{
    const int tyv = 2 * s->tma_rows, cw = 16 / (int)s->elem, rs = 32 + 2 * cw;      // TmaGeom::RS
    for (int par = 0; par < 2; ++par) {
        void *arr[5] = {s->state[par][0], s->state[par][1], s->state[par][2], s->coef[0], s->coef[1]};
        for (int f = 0; f < 5; ++f) {
            if (make_map(s, &s->tma[par].main[f], arr[f], rs, tyv + 2)) return 1;
            if (make_map(s, &s->tma[par].col[f], arr[f], cw, tyv + 2)) return 1;
            if (make_map(s, &s->tma[par].row[f], arr[f], rs, 1)) return 1;
        }
    }
    s->tma_ready = true;
    return 0;
}

template <typename T, int TYT, int NST, int MINB> static cudaError_t launch_tma_k(tgtk_sim *s, const StepArgs<T> &a)
{
    auto kern = fk2c_step_tma<T, TYT, NST, MINB>;
    const int smem = TmaGeom<T, TYT>::smem_bytes(NST);
    static bool attr_set = false;      // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = s->grid;
    cfg.blockDim = s->block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (effective_pdl(s) && s->cur_slot > 0) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a, s->lx, s->tma[s->cur_par]);
}

template <typename T> static cudaError_t launch_tma(tgtk_sim *s, const StepArgs<T> &a)
{
    if (!s->tma_ready && build_tma(s)) return cudaErrorInvalidValue;
    if (s->tma_rows == 8) return (s->tma_stages == 2) ? launch_tma_k<T, 8, 2, 2>(s, a) : launch_tma_k<T, 8, 3, 2>(s, a);
    if (s->tma_stages == 2) return launch_tma_k<T, 4, 2, 4>(s, a);
    return (s->tma_minb >= 4) ? launch_tma_k<T, 4, 3, 4>(s, a) : launch_tma_k<T, 4, 3, 3>(s, a);
}

template <typename T> static const void *tma_kernel(const tgtk_sim *s, int *smem)
{
    if (s->tma_rows == 8) {
        *smem = TmaGeom<T, 8>::smem_bytes(s->tma_stages == 2 ? 2 : 3);
        return (s->tma_stages == 2) ? (const void *)fk2c_step_tma<T, 8, 2, 2> : (const void *)fk2c_step_tma<T, 8, 3, 2>;
    }
    *smem = TmaGeom<T, 4>::smem_bytes(s->tma_stages == 2 ? 2 : 3);
    if (s->tma_stages == 2) return (const void *)fk2c_step_tma<T, 4, 2, 4>;
    return (s->tma_minb >= 4) ? (const void *)fk2c_step_tma<T, 4, 3, 4> : (const void *)fk2c_step_tma<T, 4, 3, 3>;
}

template <typename T> static cudaError_t launch_step(tgtk_sim *s)
{
    StepArgs<T> a;
    fill_args(s, a);
    if (s->lx > 0 && effective_variant(s) == 7) {
        if (s->sy != s->p.Z) s->pads_dirty = true;
        return launch_tma<T>(s, a);
    }
    if constexpr (sizeof(T) == 4) {
        if (s->lx > 0 && effective_variant(s) == 5) {   // fp32: four voxels per thread
            const int mb = s->minb ? s->minb : 4;
            switch (s->p.model) {
            case TGTK_MODEL_FK: return launch_v4(s, fk_step_v5<16, 8, 4>, a, s->lx);
            case TGTK_MODEL_DTI: return launch_v4(s, dti_step_v5<16, 8, 4>, a, s->lx);
            default:
                if (mb >= 4) return launch_v4(s, fk2c_step_v5<16, 8, 4>, a, s->lx);
                return launch_v4(s, fk2c_step_v5<16, 8, 3>, a, s->lx);
            }
        }
        if (s->lx > 0 && effective_variant(s) == 6) {   // fp32: one z pair per thread
            const int mb = s->minb ? s->minb : 6;
            switch (s->p.model) {
            case TGTK_MODEL_FK: return launch_v4(s, fk_step_v6<16, 8, 8>, a, s->lx);
            case TGTK_MODEL_DTI: return launch_v4(s, dti_step_v6<16, 8, 8>, a, s->lx);
            default:
                if (mb >= 6) return launch_v4(s, fk2c_step_v6<16, 8, 6>, a, s->lx);
                return launch_v4(s, fk2c_step_v6<16, 8, 4>, a, s->lx);
            }
        }
    }
    if (s->sy != s->p.Z) s->pads_dirty = true;   // no other kernel maintains the z = 0 mirror
    if (s->sp_active && s->lx > 0 && effective_variant(s) == 4) {   // work list over the live tiles
        const dim3 g = s->grid, b = s->block;
        s->grid = dim3((unsigned)s->sp_items, 1, 1);
        s->block = dim3(32, 4, 1);
        const SparseArgs sp = {s->sp_live, s->sp_work, s->sp_ahead, s->sp_nzt, s->sp_nyt, s->sp_trace};
        cudaError_t e;
        if (s->slab_on) {           // fused slab driver: the instantiations with the halo push and the flag protocol
            if (s->p.model == TGTK_MODEL_FK) e = launch_v4(s, fk_step_sp<T, 8, 1>, a, sp);
            else if (s->p.model == TGTK_MODEL_DTI) e = launch_v4(s, dti_step_sp<T, 8, 1>, a, sp);
            else if constexpr (sizeof(T) == 8) e = launch_v4(s, fk2c_step_sp<T, 4, 1>, a, sp);
            else e = launch_v4(s, fk2c_step_sp<T, 6, 1>, a, sp);
        } else if (s->p.model == TGTK_MODEL_FK) e = launch_v4(s, fk_step_sp<T, 8>, a, sp);
        else if (s->p.model == TGTK_MODEL_DTI) e = launch_v4(s, dti_step_sp<T, 8>, a, sp);
        else if constexpr (sizeof(T) == 8) {
            if (s->sp_async == 2 && s->sp_minb == 5) e = launch_v4(s, fk2c_step_spa<T, 5, 2>, a, sp);
            else if (s->sp_async == 2) e = launch_v4(s, fk2c_step_spa<T, 4, 2>, a, sp);
            else e = launch_v4(s, fk2c_step_sp<T, 4>, a, sp);
        } else {
            if (s->sp_async == 2) e = launch_v4(s, fk2c_step_spa<T, 6, 2>, a, sp);
            else e = launch_v4(s, fk2c_step_sp<T, 6>, a, sp);
        }
        s->grid = g;
        s->block = b;
        return e;
    }
    // the fused slab driver (and the serpentine sweep) run dedicated instantiations of the two-voxel kernels, 32 x 4 threads
    if (s->lx > 0 && effective_variant(s) == 4 && slab_kernels(s)) {
        switch (s->p.model) {
        case TGTK_MODEL_FK: return launch_v4(s, fk_step_v4<T, 32, 4, 8, 1>, a, s->lx);
        case TGTK_MODEL_DTI: return launch_v4(s, dti_step_v4<T, 32, 4, 8, 1>, a, s->lx);
        case TGTK_MODEL_FK_2C:
            if (sizeof(T) == 8) return launch_v4(s, fk2c_step_v4<T, 32, 4, 4, 0, 1>, a, s->lx);
            return launch_v4(s, fk2c_step_v4<T, 32, 4, 6, 0, 1>, a, s->lx);
        default: break;
        }
    }
    if (s->lx > 0 && effective_variant(s) == 4 && s->p.model == TGTK_MODEL_FK) {
        if (s->block.x == 16) return launch_v4(s, fk_step_v4<T, 16, 8, 8>, a, s->lx);
        else if (s->block.y == 4) return launch_v4(s, fk_step_v4<T, 32, 4, 8>, a, s->lx);
        else return launch_v4(s, fk_step_v4<T, 32, 8, 4>, a, s->lx);
        return cudaGetLastError();
    }
    if (s->lx > 0 && effective_variant(s) == 4 && s->p.model == TGTK_MODEL_DTI) {
        if (s->block.x == 16) return launch_v4(s, dti_step_v4<T, 16, 8, 8>, a, s->lx);
        else if (s->block.y == 4) return launch_v4(s, dti_step_v4<T, 32, 4, 8>, a, s->lx);
        else return launch_v4(s, dti_step_v4<T, 32, 8, 4>, a, s->lx);
        return cudaGetLastError();
    }
    if (s->lx > 0 && effective_variant(s) == 4 && s->p.model == TGTK_MODEL_FK_2C) {   // two voxels per thread
        const int mb = s->minb ? s->minb : ((sizeof(T) == 8) ? 2 : 3);
        if (s->block.x == 16) {
            if (mb >= 3) return launch_v4(s, fk2c_step_v4<T, 16, 8, 6>, a, s->lx);
            else return launch_v4(s, fk2c_step_v4<T, 16, 8, 4>, a, s->lx);
        } else if (s->block.y == 4) {
            if (mb == 5) return launch_v4(s, fk2c_step_v4<T, 32, 4, 5, 1>, a, s->lx);
            if (mb == 14) return launch_v4(s, fk2c_step_v4<T, 32, 4, 4, 1>, a, s->lx);
            if (mb >= 3) return launch_v4(s, fk2c_step_v4<T, 32, 4, 6>, a, s->lx);
            else return launch_v4(s, fk2c_step_v4<T, 32, 4, 4>, a, s->lx);
        } else {
            if (mb >= 3) return launch_v4(s, fk2c_step_v4<T, 32, 8, 3>, a, s->lx);
            else return launch_v4(s, fk2c_step_v4<T, 32, 8, 2>, a, s->lx);
        }
        return cudaGetLastError();
    }
    if (s->lx > 0 && s->p.model == TGTK_MODEL_DTI_FULL)   // 19-point stencil, four-plane shared-memory window
        return launch_v4(s, dti_full_step_v3<T, 32, 8>, a, s->lx);
    if (s->lx > 0 && effective_variant(s) >= 3) {   // shared-memory staged marching kernels
        const bool wide = (s->block.x == 32);
        switch (s->p.model) {
        case TGTK_MODEL_FK:
            if (wide) fk_step_v3<T, 32, 8><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            else fk_step_v3<T, 16, 16><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            return cudaGetLastError();
        case TGTK_MODEL_FK_2C:
            if (s->pf == 1) {
                if (wide) fk2c_step_v3s<T, 32, 8, 4><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
                else fk2c_step_v3s<T, 16, 16, 4><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            } else if (s->minb == 4) {
                if (wide) fk2c_step_v3<T, 32, 8, 4><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
                else fk2c_step_v3<T, 16, 16, 4><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            } else if (s->minb == 3 || s->minb == 0) {
                if (wide) fk2c_step_v3<T, 32, 8, 3><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
                else fk2c_step_v3<T, 16, 16, 3><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            } else {
                if (wide) fk2c_step_v3<T, 32, 8><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
                else fk2c_step_v3<T, 16, 16><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            }
            return cudaGetLastError();
        case TGTK_MODEL_DTI:
            if (wide) dti_step_v3<T, 32, 8><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            else dti_step_v3<T, 16, 16><<<s->grid, s->block, 0, s->stream>>>(a, s->lx);
            return cudaGetLastError();
        default: break;
        }
    }
    if (s->lx > 0) {   // marching geometry chosen by configure_launch
        switch (s->p.model) {
        case TGTK_MODEL_FK: fk_step_v2<T><<<s->grid, s->block, 0, s->stream>>>(a, s->lx); return cudaGetLastError();
        case TGTK_MODEL_FK_2C: fk2c_step_v2<T><<<s->grid, s->block, 0, s->stream>>>(a, s->lx); return cudaGetLastError();
        case TGTK_MODEL_DTI: dti_step_v2<T><<<s->grid, s->block, 0, s->stream>>>(a, s->lx); return cudaGetLastError();
        default: break;
        }
    }
    switch (s->p.model) {
    case TGTK_MODEL_FK: fk_step_v1<T><<<s->grid, s->block, 0, s->stream>>>(a); break;
    case TGTK_MODEL_FK_2C: fk2c_step_v1<T><<<s->grid, s->block, 0, s->stream>>>(a); break;
    case TGTK_MODEL_DTI: coef_step_v1<T, 2><<<s->grid, s->block, 0, s->stream>>>(a); break;
    case TGTK_MODEL_COEF6: coef_step_v1<T, 0><<<s->grid, s->block, 0, s->stream>>>(a); break;
    case TGTK_MODEL_FK_FACES: coef_step_v1<T, 1><<<s->grid, s->block, 0, s->stream>>>(a); break;
    case TGTK_MODEL_FK2C_COEF12: fk2c_coef_step_v1<T><<<s->grid, s->block, 0, s->stream>>>(a); break;
    case TGTK_MODEL_DTI_FULL: dti_full_step_v1<T><<<s->grid, s->block, 0, s->stream>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <typename T> static const void *v2_kernel(int model)
{
    switch (model) {
    case TGTK_MODEL_FK: return (const void *)fk_step_v2<T>;
    case TGTK_MODEL_FK_2C: return (const void *)fk2c_step_v2<T>;
    default: return (const void *)dti_step_v2<T>;
    }
}

template <typename T> static const void *v3_kernel(int model, bool wide, int minb)
{
    switch (model) {
    case TGTK_MODEL_FK: return wide ? (const void *)fk_step_v3<T, 32, 8> : (const void *)fk_step_v3<T, 16, 16>;
    case TGTK_MODEL_FK_2C:
        if (minb == 14) return wide ? (const void *)fk2c_step_v3s<T, 32, 8, 4> : (const void *)fk2c_step_v3s<T, 16, 16, 4>;
        if (minb == 4) return wide ? (const void *)fk2c_step_v3<T, 32, 8, 4> : (const void *)fk2c_step_v3<T, 16, 16, 4>;
        if (minb == 3 || minb == 0) return wide ? (const void *)fk2c_step_v3<T, 32, 8, 3> : (const void *)fk2c_step_v3<T, 16, 16, 3>;
        return wide ? (const void *)fk2c_step_v3<T, 32, 8> : (const void *)fk2c_step_v3<T, 16, 16>;
    default: return wide ? (const void *)dti_step_v3<T, 32, 8> : (const void *)dti_step_v3<T, 16, 16>;
    }
}

static void slab_p2p_close(tgtk_sim *s);

static void drop_graph(tgtk_sim *s)
{
    for (int i = 0; i < 2; ++i) {
        if (s->graph_exec[i]) cudaGraphExecDestroy(s->graph_exec[i]);
        s->graph_exec[i] = nullptr;
        if (s->slab_graph[i]) cudaGraphExecDestroy(s->slab_graph[i]);
        s->slab_graph[i] = nullptr;
    }
}

static const void *v5_kernel(int model, int minb)
{
    if (model == TGTK_MODEL_FK) return (const void *)fk_step_v5<16, 8, 4>;
    if (model == TGTK_MODEL_DTI) return (const void *)dti_step_v5<16, 8, 4>;
    return (minb == 0 || minb >= 4) ? (const void *)fk2c_step_v5<16, 8, 4> : (const void *)fk2c_step_v5<16, 8, 3>;
}

static const void *v6_kernel(int model, int minb)
{
    if (model == TGTK_MODEL_FK) return (const void *)fk_step_v6<16, 8, 8>;
    if (model == TGTK_MODEL_DTI) return (const void *)dti_step_v6<16, 8, 8>;
    return (minb == 0 || minb >= 6) ? (const void *)fk2c_step_v6<16, 8, 6> : (const void *)fk2c_step_v6<16, 8, 4>;
}

static bool slab_kernels(const tgtk_sim *s)
{
    const int m = s->p.model;
    return (s->slab_on || s->zigzag) && (m == TGTK_MODEL_FK || m == TGTK_MODEL_FK_2C || m == TGTK_MODEL_DTI);
}

static const void *v4_slab_kernel(int model, int dtype)
{
    const bool d = (dtype == TGTK_F64);
    if (model == TGTK_MODEL_FK) return d ? (const void *)fk_step_v4<double, 32, 4, 8, 1> : (const void *)fk_step_v4<float, 32, 4, 8, 1>;
    if (model == TGTK_MODEL_DTI) return d ? (const void *)dti_step_v4<double, 32, 4, 8, 1> : (const void *)dti_step_v4<float, 32, 4, 8, 1>;
    return d ? (const void *)fk2c_step_v4<double, 32, 4, 4, 0, 1> : (const void *)fk2c_step_v4<float, 32, 4, 6, 0, 1>;
}

static const void *v4_kernel(int model, int dtype, int rows, int minb)
{
    if (rows == 16) {   // 16-wide tiles
        if (model == TGTK_MODEL_FK) return dtype == TGTK_F64 ? (const void *)fk_step_v4<double, 16, 8, 8> : (const void *)fk_step_v4<float, 16, 8, 8>;
        if (model == TGTK_MODEL_DTI) return dtype == TGTK_F64 ? (const void *)dti_step_v4<double, 16, 8, 8> : (const void *)dti_step_v4<float, 16, 8, 8>;
        const int mb = minb ? minb : ((dtype == TGTK_F64) ? 2 : 3);
        if (dtype == TGTK_F64) return mb >= 3 ? (const void *)fk2c_step_v4<double, 16, 8, 6> : (const void *)fk2c_step_v4<double, 16, 8, 4>;
        return mb >= 3 ? (const void *)fk2c_step_v4<float, 16, 8, 6> : (const void *)fk2c_step_v4<float, 16, 8, 4>;
    }
    if (model == TGTK_MODEL_FK) {
        if (dtype == TGTK_F64) return rows == 4 ? (const void *)fk_step_v4<double, 32, 4, 8> : (const void *)fk_step_v4<double, 32, 8, 4>;
        return rows == 4 ? (const void *)fk_step_v4<float, 32, 4, 8> : (const void *)fk_step_v4<float, 32, 8, 4>;
    }
    if (model == TGTK_MODEL_DTI) {
        if (dtype == TGTK_F64) return rows == 4 ? (const void *)dti_step_v4<double, 32, 4, 8> : (const void *)dti_step_v4<double, 32, 8, 4>;
        return rows == 4 ? (const void *)dti_step_v4<float, 32, 4, 8> : (const void *)dti_step_v4<float, 32, 8, 4>;
    }
    if (minb == 0) minb = (dtype == TGTK_F64) ? 2 : 3;   // fp64: 120 registers, fp32: 76 (measured faster)
    if (rows == 4 && minb == 5) return dtype == TGTK_F64 ? (const void *)fk2c_step_v4<double, 32, 4, 5, 1> : (const void *)fk2c_step_v4<float, 32, 4, 5, 1>;
    if (rows == 4 && minb == 14) return dtype == TGTK_F64 ? (const void *)fk2c_step_v4<double, 32, 4, 4, 1> : (const void *)fk2c_step_v4<float, 32, 4, 4, 1>;
    if (dtype == TGTK_F64) {
        if (rows == 4) return minb >= 3 ? (const void *)fk2c_step_v4<double, 32, 4, 6> : (const void *)fk2c_step_v4<double, 32, 4, 4>;
        return minb >= 3 ? (const void *)fk2c_step_v4<double, 32, 8, 3> : (const void *)fk2c_step_v4<double, 32, 8, 2>;
    }
    if (rows == 4) return minb >= 3 ? (const void *)fk2c_step_v4<float, 32, 4, 6> : (const void *)fk2c_step_v4<float, 32, 4, 4>;
    return minb >= 3 ? (const void *)fk2c_step_v4<float, 32, 8, 3> : (const void *)fk2c_step_v4<float, 32, 8, 2>;
}

static bool has_v2(int model) { return model == TGTK_MODEL_FK || model == TGTK_MODEL_FK_2C || model == TGTK_MODEL_DTI; }

// Launch geometry.  Baseline: 32x4x2 voxels per block.  Marching: a (TZ x TY) column tile per
// block, TZ in {64,32,16} picked to waste the fewest lanes on the ragged z edge, and the x
// axis cut into segments so that the grid is a few waves of the 148 SMs.
static int configure_launch(tgtk_sim *s)
{
    const tgtk_params &p = s->p;
    const bool fits32 = (uint64_t)p.X * (uint64_t)s->sx < (1ull << 31);   // 32-bit element offsets
    const int variant = effective_variant(s);
    const int NX = s->slab_on ? s->slab_owned : p.X;       // planes a launch computes
    s->tma_ready = false;
    if (variant == 7) {
        // 32 x 2*rows voxel tile on 32 x rows consumer threads + one producer warp; x cut into segments for whole waves
        const int rows = s->tma_rows, tyv = 2 * rows;
        const int tiles = ((p.Z + 31) / 32) * ((p.Y + tyv - 1) / tyv);
        int per_sm = 2, sms = 148, smem = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        const void *fn = (p.dtype == TGTK_F64) ? tma_kernel<double>(s, &smem) : tma_kernel<float>(s, &smem);
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 32 * (rows + 1), smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        const double slots = (double)sms * per_sm;
        int best_lx = NX;
        double best_eff = -1.0;
        for (int lx = 1; lx <= NX; ++lx) {
            const int ns = (NX + lx - 1) / lx;
            const double waves = tiles * (double)ns / slots;
            const double eff = waves / std::ceil(waves) * (lx / (lx + 1.0)) * ((double)NX / ((double)ns * lx));   // two extra tile loads per segment
            if (eff > best_eff + 1e-9) { best_eff = eff; best_lx = lx; }
        }
        s->lx = best_lx;
        s->block = dim3(32, rows + 1, 1);
        s->grid = dim3((p.Z + 31) / 32, (p.Y + tyv - 1) / tyv, (NX + best_lx - 1) / best_lx);
    } else
    if (variant >= 2 && has_v2(p.model) && fits32) {
        int best_tz = 32;
        double best = 1e30;
        for (int tz : {64, 32, 16}) {
            if (variant >= 3 && tz == 64) continue;   // staged kernels: 32x8 or 16x16 tiles
            if (s->force_tz && tz != s->force_tz) continue;
            const int ty = 256 / tz;
            const double padded = (double)((p.Z + tz - 1) / tz * tz) * ((p.Y + ty - 1) / ty * ty);
            const double cost = padded / ((double)p.Z * p.Y) + 0.08 * (tz == 16) + 0.03 * (tz == 64);   // full-warp rows measure faster
            if (cost < best) { best = cost; best_tz = tz; }
        }
        const bool four = (variant == 5);       // *_step_v5 (fp32): 32 x 16 voxel tile on 16 x 8 threads
        const bool zp1 = (variant == 6);        // *_step_v6 (fp32): 32 x 8 voxel tile on 16 x 8 threads
        const bool two_rows = (variant == 4) || four || zp1;   // *_step_v4: 32 x 8 or 32 x 16 voxel tile, two rows per thread
        // v4 tiles: 32(z) x 8(y) voxels on 32x4 threads, 16 x 16 voxels on 16x8 threads (narrow boxes waste
        // fewer lanes), or 32 x 16 voxels on 32x8 threads (TGTK_V4_THREADS=256)
        bool narrow = false;
        if (two_rows && s->v4_threads == 128) {
            const double w32 = (double)((p.Z + 31) / 32 * 32) * ((p.Y + 7) / 8 * 8);
            const double w16 = (double)((p.Z + 15) / 16 * 16) * ((p.Y + 15) / 16 * 16);
            narrow = s->force_tz ? (s->force_tz == 16) : (w16 * (1.0 + s->narrow_penalty) < w32);
        }
        if (four || zp1) narrow = false;
        const bool slabk = (variant == 4) && slab_kernels(s);
        if (slabk) narrow = false;
        const int v4rows = slabk ? 4 : (narrow || four || zp1) ? 8 : ((s->v4_threads == 128) ? 4 : 8);   // thread rows; voxel rows = 2x
        const int tz = two_rows ? (narrow ? 16 : 32) : best_tz, ty = zp1 ? 8 : two_rows ? 2 * v4rows : 256 / tz;   // voxels per tile
        const int tiles = ((p.Z + tz - 1) / tz) * ((p.Y + ty - 1) / ty);
        // Segment length: trade the tail of the last wave (blocks / resident slots not an
        // integer) against the two extra planes every segment loads to prime its pipeline.
        int per_sm = 2, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        const void *fn = four ? v5_kernel(p.model, s->minb) : zp1 ? v6_kernel(p.model, s->minb)
                         : slabk ? v4_slab_kernel(p.model, p.dtype)
                         : two_rows ? v4_kernel(p.model, p.dtype, narrow ? 16 : v4rows, s->minb)
                         : (variant >= 3)
                             ? ((p.dtype == TGTK_F64) ? v3_kernel<double>(p.model, tz == 32, s->pf == 1 ? 14 : s->minb) : v3_kernel<float>(p.model, tz == 32, s->pf == 1 ? 14 : s->minb))
                             : ((p.dtype == TGTK_F64) ? v2_kernel<double>(p.model) : v2_kernel<float>(p.model));
        const int nthreads = (four || zp1) ? 128 : two_rows ? tz * v4rows : 256;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, nthreads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        const double slots = (double)sms * per_sm;
        int best_lx = NX;
        double best_eff = -1.0;
        const int lx_max = NX;
        for (int lx = 1; lx <= lx_max; ++lx) {
            const int ns = (NX + lx - 1) / lx;
            const double waves = tiles * (double)ns / slots;
            const double eff = waves / std::ceil(waves) * (lx / (lx + 0.6)) * ((double)NX / ((double)ns * lx));
            if (eff > best_eff + 1e-9) { best_eff = eff; best_lx = lx; }
        }
        if (s->force_lx > 0) best_lx = std::min(s->force_lx, NX);      // TGTK_LX: experiments
        s->lx = best_lx;
        const int nseg = (NX + s->lx - 1) / s->lx;
        s->block = dim3((four || zp1) ? 16 : tz, two_rows ? v4rows : ty, 1);
        s->grid = dim3((p.Z + tz - 1) / tz, (p.Y + ty - 1) / ty, nseg);
    } else if (p.model == TGTK_MODEL_DTI_FULL && variant >= 3 && fits32) {
        // dti_full_step_v3: 32 x 8 tile, x cut into segments so that the grid is a whole number of waves
        const int tiles = ((p.Z + 31) / 32) * ((p.Y + 7) / 8);
        int per_sm = 2, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        const void *fn = (p.dtype == TGTK_F64) ? (const void *)dti_full_step_v3<double, 32, 8> : (const void *)dti_full_step_v3<float, 32, 8>;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        const double slots = (double)sms * per_sm;
        int best_lx = p.X;
        double best_eff = -1.0;
        for (int lx = 1; lx <= p.X; ++lx) {
            const int ns = (p.X + lx - 1) / lx;
            const double waves = tiles * (double)ns / slots;
            const double eff = waves / std::ceil(waves) * (lx / (lx + 2.0)) * ((double)p.X / ((double)ns * lx));
            if (eff > best_eff + 1e-9) { best_eff = eff; best_lx = lx; }
        }
        s->lx = best_lx;
        s->block = dim3(32, 8, 1);
        s->grid = dim3((p.Z + 31) / 32, (p.Y + 7) / 8, (p.X + best_lx - 1) / best_lx);
    } else {
        s->lx = 0;
        s->block = dim3(32, 4, 2);
        s->grid = dim3((p.Z + 31) / 32, (p.Y + 3) / 4, (p.X + 1) / 2);
    }
    s->nblocks = s->grid.x * s->grid.y * s->grid.z;
    drop_graph(s);
    // two-step kernel geometry: 30 x (rows - 2) output voxels per 32 x rows thread block, x cut into segments like above
    s->tb_nblocks = 0;
    if ((p.model == TGTK_MODEL_DTI || p.model == TGTK_MODEL_FK) && fits32) {
        const int rows = (s->tb_rows == 10 && p.model == TGTK_MODEL_DTI) ? 10 : 18, oy = rows - 2;
        const int tiles = ((p.Z + 29) / 30) * ((p.Y + oy - 1) / oy);
        int per_sm = 2, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        const void *fn = (p.model == TGTK_MODEL_FK)
                             ? ((p.dtype == TGTK_F64) ? (const void *)fk_step2_tb<double, 18, 2> : (const void *)fk_step2_tb<float, 18, 2>)
                         : (p.dtype == TGTK_F64)
                             ? (rows == 10 ? (const void *)dti_step2_tb<double, 10, 3> : (const void *)dti_step2_tb<double, 18, 2>)
                             : (rows == 10 ? (const void *)dti_step2_tb<float, 10, 3> : (const void *)dti_step2_tb<float, 18, 2>);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 32 * rows, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        const double slots = (double)sms * per_sm;
        int best_lx = p.X;
        double best_eff = -1.0;
        for (int lx = 1; lx <= p.X; ++lx) {
            const int ns = (p.X + lx - 1) / lx;
            const double waves = tiles * (double)ns / slots;
            const double eff = waves / std::ceil(waves) * (lx / (lx + 2.0)) * ((double)p.X / ((double)ns * lx));   // two extra planes per segment
            if (eff > best_eff + 1e-9) { best_eff = eff; best_lx = lx; }
        }
        s->tb_lx = best_lx;
        s->tb_block = dim3(32, rows, 1);
        s->tb_grid = dim3((p.Z + 29) / 30, (p.Y + oy - 1) / oy, (p.X + best_lx - 1) / best_lx);
        s->tb_nblocks = s->tb_grid.x * s->tb_grid.y * s->tb_grid.z;
    }
    const unsigned need = std::max(s->nblocks, s->tb_nblocks);
    if (need > s->partials_cap) {
        if (s->partials) CU(sim_free(s, s->partials));
        s->partials = nullptr;
        CU(sim_malloc(s, (void **)(&s->partials), sizeof(double) * 2 * need * kGraphSteps));   // ring of kGraphSteps slots
        s->partials_cap = need;
    }
    s->ring_stride = 2 * s->partials_cap;
    return 0;
}

// ---- sparse launches: the work list (pure host code; exported for the CPU tests as tgtk_sparse_plan_host) ------------
struct PlanOpts {
    bool slab = false;          // fused slab driver: items that touch plane xa / end on plane xb go last
    bool waves_set = false;
    double waves = 1.0, split = 0.5;
};
static inline int live_warps(uint32_t w) { return ((w & 0xffu) != 0) + ((w & 0xff00u) != 0) + ((w & 0xff0000u) != 0) + ((w >> 24) != 0); }

// live: [X][nzt][nyt] words; planes [xa, xb) are computed; `slots` = resident blocks of the device for the kernel.
// Out: two int4 per item -- {y tile, z tile, x0, x1}, {live words of planes x0 .. x0+3 (periodic in X)}.
static void plan_work_list(const uint32_t *live, int X, int nzt, int nyt, int xa, int xb, int slots, const PlanOpts &o,
                           std::vector<int4> &work)
{
    auto at = [&](int x, int zt, int yt) { return live[((size_t)x * nzt + zt) * nyt + yt]; };
    // runs of planes with live tiles per (y tile, z tile) column; a gap of kGap dead planes ends a run
    struct Run { int yt, zt, x0, x1; double w; };
    const int kGap = 4;
    // Relative cost of a block plane.  Measured (tools/sparse_trace.py): a plane takes a block 1.5-1.9 us whether one or four
    // of its warps are live -- the barrier makes the block as slow as its slowest warp -- and a plane without live warps a
    // fraction of that.  TGTK_SP_COST="base:per_warp:dead".
    double kBase = 3.4, kWarp = 0.15, kDeadPlane = 0.4;
    if (const char *v = getenv("TGTK_SP_COST")) sscanf(v, "%lf:%lf:%lf", &kBase, &kWarp, &kDeadPlane);
    auto cost = [&](uint32_t w) { return w ? kBase + kWarp * live_warps(w) : kDeadPlane; };
    std::vector<Run> runs;
    double total = 0.0;
    for (int zt = 0; zt < nzt; ++zt)
        for (int yt = 0; yt < nyt; ++yt) {
            int x = xa;
            while (x < xb) {
                while (x < xb && !at(x, zt, yt)) ++x;
                if (x >= xb) break;
                int x0 = x, last = x;
                double w = 0.0;
                for (; x < xb && x - last <= kGap; ++x)
                    if (at(x, zt, yt)) last = x;
                for (int q = x0; q <= last; ++q) w += cost(at(q, zt, yt));
                runs.push_back({yt, zt, x0, last + 1, w});
                total += w;
                x = last + 1;
            }
        }
    // The list: every column's runs cut into items of a target size.  For long shares (big boxes) the size is decided by
    // trying a few -- a slot's share of the work in 1, 1.5, 2, 3, 4, 6 or 8 items -- and playing each list through the block scheduler
    // (blocks start in list order on the slot that frees first; an item costs its planes plus a prologue worth ~1.75 planes,
    // tools/sparse_trace.py): one item per slot avoids prologues but ends unevenly and falls off a cliff when the list is
    // a handful of items longer than the resident slots; many short items balance by themselves but pay for it.
    const double plane = kBase + 2.5 * kWarp, prologue = 1.75 * plane;
    const double floor_w = 4.0 * plane;                        // at least ~4 full planes per item
    
    const double share = total / slots;
    auto order = [&](std::vector<Run> &v) {
        std::stable_sort(v.begin(), v.end(), [](const Run &a, const Run &b) { return a.w > b.w; });   // longest first
        if (o.slab)     // ... and the items that wait for a neighbour's word last: by then it has long arrived
            std::stable_partition(v.begin(), v.end(), [&](const Run &r) { return r.x0 != xa && r.x1 != xb; });
    };
    auto cut = [&](double waves, std::vector<Run> &out, double grow = 1.0) {
        const bool two = std::fabs(waves - 2.0) < 1e-9;
        const double tgt[2] = {std::max((two ? share * o.split : share / waves) * grow, floor_w),
                               std::max((two ? share * (1.0 - o.split) : share / waves) * grow, floor_w)};
        out.clear();
        int turn = 0;
        for (const Run &r : runs) {
            int x0 = r.x0;
            double acc = 0.0, left = r.w;
            for (int q = r.x0; q < r.x1; ++q) {
                const double c = cost(at(q, r.zt, r.yt));
                acc += c;
                left -= c;
                const bool last = (q + 1 == r.x1);
                // cut when the item has reached its size and what is left of the run is worth an item of its own
                if (last || (acc >= tgt[turn] - 1e-9 && left >= 0.5 * std::min(tgt[0], tgt[1]))) {
                    out.push_back({r.yt, r.zt, x0, q + 1, acc});
                    x0 = q + 1;
                    acc = 0.0;
                    turn ^= 1;
                }
            }
        }
        order(out);
    };
    auto makespan = [&](const std::vector<Run> &v) {
        std::vector<double> heap((size_t)slots, 0.0);          // min-heap of the slots' free times
        auto cmp = [](double a, double b) { return a > b; };
        double end = 0.0;
        for (const Run &r : v) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            heap.back() += r.w + prologue;
            end = std::max(end, heap.back());
            std::push_heap(heap.begin(), heap.end(), cmp);
        }
        return end;
    };
    std::vector<Run> items, trial;
    if (o.waves_set) {
        cut(o.waves, items);
    } else if (share < 20.0 * plane) {
        // short shares: ONE item per slot, sized up until the list fits the resident slots.  (Measured on the atlas box: the
        // model below would pick two rounds for FK_2c and lose 2 % -- it does not know that the blocks left on an SM speed
        // up when their neighbours have finished, which evens out a single round by itself.)
        for (double grow = 1.0; grow < 3.0; grow *= 1.04) {
            cut(1.0, items, grow);
            if ((int)items.size() <= slots) break;
        }
    } else {
        double best = 1e300;
        for (double w : {1.0, 1.5, 2.0, 3.0, 4.0, 6.0, 8.0}) {
            cut(w, trial);
            const double t = makespan(trial);
            if (t < best * 0.95) { best = t; items.swap(trial); }      // (a finer list has to win by 5 %)
        }
    }
    work.assign(2 * items.size(), make_int4(0, 0, 0, 0));
    auto word = [&](const Run &r, int d) { return (int)at((r.x0 + d) % X, r.zt, r.yt); };
    for (size_t i = 0; i < items.size(); ++i) {
        work[2 * i] = make_int4(items[i].yt, items[i].zt, items[i].x0, items[i].x1);
        work[2 * i + 1] = make_int4(word(items[i], 0), word(items[i], 1), word(items[i], 2), word(items[i], 3));
    }
}

// ---- sparse launches: live-tile mask and work list (tgtk_kernels_sparse.cuh) ---------------------------------------
static bool tb_enabled(const tgtk_sim *s);

static bool sparse_eligible(const tgtk_sim *s)
{
    const tgtk_params &p = s->p;
    if (s->sparse == 0 || (s->sp_forbid && !s->slab_on)) return false;      // (the fused slab driver only READS through its buffer views)
    if (p.model != TGTK_MODEL_FK_2C && p.model != TGTK_MODEL_FK && p.model != TGTK_MODEL_DTI) return false;
    const bool dti32_auto = s->variant == 0 && p.dtype == TGTK_F32 && p.model == TGTK_MODEL_DTI;      // runs the z-pair kernels unless sparse
    if (s->lx <= 0 || (!dti32_auto && effective_variant(s) != 4) || p.X < 8) return false;
    if ((uint64_t)p.X * (uint64_t)s->sx >= (1ull << 31)) return false;
    if (s->sparse < 0 && s->n < (size_t)s->sp_min_voxels) return false;      // auto: not for boxes whose loop is a few launch latencies long
    return !tb_enabled(s);
}


// The mask of the CURRENT state, the copy of that state into the other ping-pong buffer (a dead tile is never written
// again: both buffers must hold it), the work list.  Synchronises the stream; once per upload.
static int build_sparse_inner(tgtk_sim *s)
{
    const tgtk_params &p = s->p;
    const bool was = s->sp_active;
    s->sp_dirty = false;
    s->sp_active = false;
    if (!sparse_eligible(s)) { if (was) drop_graph(s); return 0; }
    const int nyt = (p.Y + 7) / 8, nzt = (p.Z + 31) / 32;
    const size_t words = (size_t)p.X * nyt * nzt;
    if (words > s->sp_live_cap) {
        if (s->sp_live) CU(sim_free(s, s->sp_live));
        s->sp_live = nullptr;
        CU(sim_malloc(s, (void **)(&s->sp_live), words * sizeof(uint32_t)));
        s->sp_live_cap = words;
    }
    const int cur = s->t_enqueued & 1;
    const size_t bytes = (size_t)p.X * s->sx * s->elem;
    for (int f = 0; f < s->nfields; ++f)
        CU(cudaMemcpyAsync(s->state[cur ^ 1][f], s->state[cur][f], bytes, cudaMemcpyDeviceToDevice, s->stream));
    const int mode = (p.model == TGTK_MODEL_FK_2C) ? 0 : (p.model == TGTK_MODEL_FK) ? 1 : 2;
    if (p.dtype == TGTK_F64)
        sparse_mask_kernel<double, 2><<<(unsigned)words, dim3(32, 4, 1), 0, s->stream>>>(
            (const double *)s->state[cur][0], (const double *)s->state[cur][1], (const double *)s->coef[0], (const double *)s->coef[1],
            (const double *)s->coef[2], s->sp_live, p.X, p.Y, p.Z, s->sx, s->sy, nzt, nyt, mode);
    else
        sparse_mask_kernel<float, 2><<<(unsigned)words, dim3(32, 4, 1), 0, s->stream>>>(
            (const float *)s->state[cur][0], (const float *)s->state[cur][1], (const float *)s->coef[0], (const float *)s->coef[1],
            (const float *)s->coef[2], s->sp_live, p.X, p.Y, p.Z, s->sx, s->sy, nzt, nyt, mode);
    CU(cudaGetLastError());
    s->aux_launches += 1;
    std::vector<uint32_t> live(words);
    CU(cudaMemcpyAsync(live.data(), s->sp_live, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));

    // voxels in live warp tiles (the bytes a sparse launch has to move are proportional to it); fused slab mode: of the
    // planes this rank owns -- the ghost planes are the neighbours' to compute
    const int xa = s->slab_on ? 1 : 0, xb = s->slab_on ? 1 + s->slab_owned : p.X;
    int64_t live_vox = 0;
    auto at = [&](int x, int zt, int yt) { return live[((size_t)x * nzt + zt) * nyt + yt]; };
    for (int x = xa; x < xb; ++x)
        for (int zt = 0; zt < nzt; ++zt)
            for (int yt = 0; yt < nyt; ++yt) {
                const uint32_t w = at(x, zt, yt);
                const int zs = std::min(32, p.Z - zt * 32);
                for (int k = 0; k < 4; ++k)
                    if ((w >> (8 * k)) & 0xffu) live_vox += (int64_t)zs * std::max(0, std::min(2, p.Y - (yt * 8 + 2 * k)));
            }
    s->sp_live_voxels = live_vox;
    const double owned_vox = (double)(xb - xa) * p.Y * p.Z;
    if (s->sparse < 0 && (double)live_vox > 0.9 * owned_vox) { if (was) drop_graph(s); return 0; }   // nothing to gain
    if (live_vox == 0) { if (was) drop_graph(s); return 0; }                                             // (and nothing to launch)

    int per_sm = 4, sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    const bool f64 = (p.dtype == TGTK_F64);
    const void *fn;
    if (p.model == TGTK_MODEL_FK) fn = f64 ? (const void *)fk_step_sp<double, 8> : (const void *)fk_step_sp<float, 8>;
    else if (p.model == TGTK_MODEL_DTI) fn = f64 ? (const void *)dti_step_sp<double, 8> : (const void *)dti_step_sp<float, 8>;
    else
        fn = f64 ? (s->sp_async == 2 ? (s->sp_minb == 5 ? (const void *)fk2c_step_spa<double, 5, 2> : (const void *)fk2c_step_spa<double, 4, 2>)
                                     : (const void *)fk2c_step_sp<double, 4>)
                 : (s->sp_async == 2 ? (const void *)fk2c_step_spa<float, 6, 2> : (const void *)fk2c_step_sp<float, 6>);
    if (s->sp_carveout >= 0) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, s->sp_carveout);   // TGTK_SP_CARVEOUT: experiments
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    s->sp_per_sm = per_sm;
    if (s->sp_ahead < 0) s->sp_ahead = sms * per_sm;
    PlanOpts o;
    o.slab = s->slab_on; o.waves_set = s->sp_waves_set; o.waves = s->sp_waves; o.split = s->sp_split;
    std::vector<int4> work;
    plan_work_list(live.data(), p.X, nzt, nyt, xa, xb, sms * per_sm, o, work);
    if (work.size() > s->sp_work_cap) {
        if (s->sp_work) CU(sim_free(s, s->sp_work));
        s->sp_work = nullptr;
        CU(sim_malloc(s, (void **)(&s->sp_work), work.size() * sizeof(int4)));
        s->sp_work_cap = work.size();
    }
    CU(cudaMemcpyAsync(s->sp_work, work.data(), work.size() * sizeof(int4), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->sp_items = (int)(work.size() / 2);
    s->sp_nzt = nzt;
    s->sp_nyt = nyt;
    if ((unsigned)s->sp_items > s->partials_cap) {
        if (s->partials) CU(sim_free(s, s->partials));
        s->partials = nullptr;
        CU(sim_malloc(s, (void **)(&s->partials), sizeof(double) * 2 * s->sp_items * kGraphSteps));
        s->partials_cap = (unsigned)s->sp_items;
        s->ring_stride = 2 * s->partials_cap;
    }
    s->sp_active = true;
    drop_graph(s);
    return 0;
}

static int build_sparse(tgtk_sim *s)
{
    const bool was = s->sp_active;
    const int v0 = effective_variant(s);
    if (build_sparse_inner(s)) return 1;
    if (s->sp_active != was && effective_variant(s) != v0) {      // FK_DTI fp32: the kernel family follows the mode
        if (configure_launch(s)) return 1;
        if (s->sp_active && (unsigned)s->sp_items > s->partials_cap) return fail("sparse", "partial-sum ring too small");
    }
    return 0;
}

static int ensure_volumes(tgtk_sim *s, int need)
{
    if (need <= s->volumes_cap) return 0;
    int cap = std::max(need, std::max(1024, s->volumes_cap * 2));
    double *nv = nullptr;
    CU(sim_malloc(s, (void **)(&nv), sizeof(double) * cap));
    CU(cudaMemsetAsync(nv, 0, sizeof(double) * cap, s->stream));
    if (s->volumes) {
        CU(cudaMemcpyAsync(nv, s->volumes, sizeof(double) * s->volumes_cap, cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        CU(sim_free(s, s->volumes));
    }
    s->volumes = nv;
    s->volumes_cap = cap;
    drop_graph(s);   // the log pointer is a kernel argument
    return 0;
}

// kGraphSteps consecutive step launches captured once and replayed: every launch has the same
// arguments because the step index and the ping-pong parity are read from Ctrl on the device.

static cudaError_t launch_one(tgtk_sim *s, int par, int slot)
{
    s->cur_par = par;
    s->cur_slot = slot;
    return (s->p.dtype == TGTK_F64) ? launch_step<double>(s) : launch_step<float>(s);
}

static bool tb_enabled(const tgtk_sim *s)
{
    if (s->tb_nblocks == 0 || !(s->host_sched || s->spec) || s->replaying || s->slab_on) return false;
    // the two-step kernels do not write the z = Z mirror element the fp32 z-pair kernels read as their periodic
    // neighbour: a chunk that mixes both families on an odd-Z box would step its leftovers on a stale mirror
    if (zpair_variant(s) && s->sy != s->p.Z) return false;
    if (s->tb >= 0) return s->tb != 0;
    // auto: measured on B200 the two-step kernel wins where the one-step kernel sits on the HBM
    // roofline -- fp64 boxes well beyond L2 (+20 % at 28 M voxels) -- and is level or behind elsewhere
    const tgtk_params &p = s->p;
    const double touched = (double)(2 * s->nfields + s->ncoef) * (double)p.X * (double)s->sx * (double)s->elem;
    // (FK: the two-step kernel is correct but 14 % behind the two-voxel one-step kernel even at 28 M voxels -- opt-in only)
    return p.model == TGTK_MODEL_DTI && p.dtype == TGTK_F64 && touched > 2.0 * (double)s->l2_bytes;
}

// two steps in one launch: reads state[par], writes state[par ^ 1], partial sums into ring slots slot, slot + 1
static cudaError_t launch_pair(tgtk_sim *s, int par, int slot)
{
    s->cur_par = par;
    s->cur_slot = slot;
    const dim3 g = s->grid, b = s->block;
    s->grid = s->tb_grid;
    s->block = s->tb_block;
    cudaError_t e;
    const bool fk = (s->p.model == TGTK_MODEL_FK);
    if (s->p.dtype == TGTK_F64) {
        StepArgs<double> a;
        fill_args(s, a);
        if (fk) e = launch_v4(s, fk_step2_tb<double, 18, 2>, a, s->tb_lx);
        else e = (s->tb_rows == 10) ? launch_v4(s, dti_step2_tb<double, 10, 3>, a, s->tb_lx) : launch_v4(s, dti_step2_tb<double, 18, 2>, a, s->tb_lx);
    } else {
        StepArgs<float> a;
        fill_args(s, a);
        if (fk) e = launch_v4(s, fk_step2_tb<float, 18, 2>, a, s->tb_lx);
        else e = (s->tb_rows == 10) ? launch_v4(s, dti_step2_tb<float, 10, 3>, a, s->tb_lx) : launch_v4(s, dti_step2_tb<float, 18, 2>, a, s->tb_lx);
    }
    s->grid = g;
    s->block = b;
    if (s->sy != s->p.Z) s->pads_dirty = true;
    return e;
}

// n consecutive steps of one chunk starting with ping-pong parity par0, ring slots 0..n-1.  With temporal
// blocking an EVEN number of two-step launches comes first: each one writes the buffer the state did not
// live in, so the two buffer pointers are swapped after it (state after t steps stays in state[t & 1]); an
// even count leaves the pointers as they were, which is what captured graphs and the peers of a slab rely on.
static cudaError_t launch_steps(tgtk_sim *s, int par0, int n)
{
    int i = 0;
    cudaError_t e = cudaSuccess;
    if (tb_enabled(s)) {
        const int pairs = (n / 4) * 2;
        for (int j = 0; j < pairs && e == cudaSuccess; ++j, i += 2) {
            e = launch_pair(s, par0, i);
            for (int f = 0; f < s->nfields; ++f) std::swap(s->state[0][f], s->state[1][f]);
            s->chunk_tb_mask |= 3u << i;
            ++s->launches;
        }
    }
    for (; i < n && e == cudaSuccess; ++i) {
        e = launch_one(s, (par0 + i) & 1, i);
        ++s->launches;
    }
    return e;
}

static cudaError_t slab_allreduce(tgtk_sim *s, int count);   // tgtk_nccl.inc

static cudaError_t launch_chunk_finish(tgtk_sim *s, int count)
{
    const tgtk_params &p = s->p;
    const int spec = (s->spec && !s->replaying) ? 1 : 0;
    ++s->aux_launches;
    chunk_finish_kernel<<<count, 256, 0, s->stream>>>(s->ctrl, s->partials, s->volumes, s->volumes_cap,
                                                   s->sp_active ? (unsigned)s->sp_items : s->nblocks, s->nfields,
                                                   count, p.h[0] * p.h[1] * p.h[2], 1.0, spec,
                                                   p.stopping_volume, p.small_volume, s->ring_stride, s->tb_nblocks,
                                                   s->chunk_tb_mask, s->slab_on ? s->chunkvol : nullptr);
    s->chunk_tb_mask = 0;
    cudaError_t e = cudaGetLastError();
    if (!s->slab_on || e != cudaSuccess) return e;
    // fused slab mode: sum the chunk's volumes over the ranks (only a stop criterion needs them during the loop; otherwise
    // the log keeps the rank's own sums and the caller reduces it once at the end), then log them, run the stop tests,
    // advance the counters
    if (s->spec) e = slab_allreduce(s, count);
    if (e != cudaSuccess) return e;
    ++s->aux_launches;
    chunk_commit_kernel<<<1, 32, 0, s->stream>>>(s->ctrl, s->chunkvol, s->volumes, s->volumes_cap, count, spec, p.stopping_volume,
                                               p.small_volume, s->slab_flags, s->peer_flags[0] + kSlabFromNext,
                                               s->peer_flags[1] + kSlabFromPrev);
    return cudaGetLastError();
}

// speculative mode: copy the state held by ping-pong buffer `par` aside before a chunk starts
static cudaError_t launch_checkpoint(tgtk_sim *s, int par)
{
    const size_t n8 = (size_t)s->p.X * s->sx * s->elem / 8;
    ++s->aux_launches;
    checkpoint_kernel<<<grid_1d(n8), 256, 0, s->stream>>>(s->ctrl, (const uint2 *)s->state[par][0], (const uint2 *)s->state[par][1],
                                                       (const uint2 *)s->state[par][2], (uint2 *)s->ckpt[0], (uint2 *)s->ckpt[1],
                                                       (uint2 *)s->ckpt[2], n8, s->slab_on ? s->slab_flags : nullptr);
    return cudaGetLastError();
}

// One graph per parity of its first step: in host-scheduled mode the parity is a launch argument.
static int build_graph(tgtk_sim *s, int par0)
{
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = s->spec ? launch_checkpoint(s, par0) : cudaSuccess;
    const int64_t before = s->launches;
    if (e == cudaSuccess) e = launch_steps(s, par0, kGraphSteps);
    s->graph_launches = (int)(s->launches - before);
    s->launches = before;                      // capture is not execution: tgtk_sim_run counts the replays
    if (e == cudaSuccess && (s->host_sched || s->spec)) e = launch_chunk_finish(s, kGraphSteps);
    cudaError_t e2 = cudaStreamEndCapture(s->stream, &graph);
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail("graph capture", cudaGetErrorString(e)); }
    if (e2 != cudaSuccess) return fail("cudaStreamEndCapture", cudaGetErrorString(e2));
    e = cudaGraphInstantiate(&s->graph_exec[par0], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { s->graph_exec[par0] = nullptr; return fail("cudaGraphInstantiate", cudaGetErrorString(e)); }
    return 0;
}

static void *field_ptr(tgtk_sim *s, int field, int which_buf)
{
    if (field >= 0 && field < s->nfields) return s->state[which_buf][field];
    if (field >= TGTK_FIELD_COEF0 && field < TGTK_FIELD_COEF0 + 12) return s->coef[field - TGTK_FIELD_COEF0];
    return nullptr;
}

// host (float64, strided) <-> dense float64 device buffer of shape (X,Y,Z)
static int copy_host_dev(cudaStream_t stream, int X, int Y, int Z, double *dev, double *host, const int64_t st[3],
                         bool to_device)
{
    const size_t n = (size_t)X * Y * Z;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (st[2] == 1 && st[1] == Z && st[0] == (int64_t)Y * Z) {
        if (to_device) CU(cudaMemcpyAsync(dev, host, n * 8, kind, stream));
        else CU(cudaMemcpyAsync(host, dev, n * 8, kind, stream));
        return 0;
    }
    if (st[2] == 1 && st[1] >= Z && st[0] % st[1] == 0 && st[0] / st[1] >= Y) {
        cudaMemcpy3DParms c;
        memset(&c, 0, sizeof(c));
        cudaPitchedPtr hp = make_cudaPitchedPtr(host, (size_t)st[1] * 8, (size_t)Z * 8, (size_t)(st[0] / st[1]));
        cudaPitchedPtr dp = make_cudaPitchedPtr(dev, (size_t)Z * 8, (size_t)Z * 8, (size_t)Y);
        c.srcPtr = to_device ? hp : dp;
        c.dstPtr = to_device ? dp : hp;
        c.extent = make_cudaExtent((size_t)Z * 8, (size_t)Y, (size_t)X);
        c.kind = kind;
        CU(cudaMemcpy3DAsync(&c, stream));
        return 0;
    }
    // arbitrary strides: gather / scatter on the host
    std::vector<double> tmp(n);
    if (to_device) {
        size_t q = 0;
        for (int i = 0; i < X; ++i)
            for (int j = 0; j < Y; ++j)
                for (int k = 0; k < Z; ++k) tmp[q++] = host[i * st[0] + j * st[1] + k * st[2]];
        CU(cudaMemcpyAsync(dev, tmp.data(), n * 8, kind, stream));
        CU(cudaStreamSynchronize(stream));
    } else {
        CU(cudaMemcpyAsync(tmp.data(), dev, n * 8, kind, stream));
        CU(cudaStreamSynchronize(stream));
        size_t q = 0;
        for (int i = 0; i < X; ++i)
            for (int j = 0; j < Y; ++j)
                for (int k = 0; k < Z; ++k) host[i * st[0] + j * st[1] + k * st[2]] = tmp[q++];
    }
    return 0;
}

static int copy_host_stage(tgtk_sim *s, double *host, const int64_t st[3], bool to_device)
{
    return copy_host_dev(s->stream, s->p.X, s->p.Y, s->p.Z, s->stage, host, st, to_device);
}

// scipy.ndimage.zoom(v, f, order=1) of a virtual volume `v` of shape (VX,VY,VZ) that is zero
// except for the box [lo, lo+in_shape) holding `in` -- i.e. restore_tumor (FK/FK.py:75-90) fused
// with the resample of FK/FK.py:233-238.  Coordinate map in = out*(n_in-1)/(n_out-1).
__global__ void resample_linear_kernel(const double *__restrict__ in, int IX, int IY, int IZ, int lx, int ly, int lz,
                                       int VX, int VY, int VZ, double *__restrict__ out, int OX, int OY, int OZ)
{
    const double fx = (OX > 1) ? (double)(VX - 1) / (double)(OX - 1) : 0.0;
    const double fy = (OY > 1) ? (double)(VY - 1) / (double)(OY - 1) : 0.0;
    const double fz = (OZ > 1) ? (double)(VZ - 1) / (double)(OZ - 1) : 0.0;
    const int64_t n = (int64_t)OX * OY * OZ;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(q % OZ);
        const int64_t r = q / OZ;
        const int j = (int)(r % OY);
        const int i = (int)(r / OY);
        const double xi = i * fx, yj = j * fy, zk = k * fz;
        int i0 = (int)floor(xi), j0 = (int)floor(yj), k0 = (int)floor(zk);
        i0 = min(i0, max(VX - 2, 0)); j0 = min(j0, max(VY - 2, 0)); k0 = min(k0, max(VZ - 2, 0));
        const double tx = xi - i0, ty = yj - j0, tz = zk - k0;
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int ii = min(i0 + a, VX - 1) - lx, jj = min(j0 + b, VY - 1) - ly, kk = min(k0 + c, VZ - 1) - lz;
                    const double w = (a ? tx : 1.0 - tx) * (b ? ty : 1.0 - ty) * (c ? tz : 1.0 - tz);
                    double v = 0.0;
                    if (ii >= 0 && ii < IX && jj >= 0 && jj < IY && kk >= 0 && kk < IZ) v = in[((int64_t)ii * IY + jj) * IZ + kk];
                    acc += w * v;
                }
        out[q] = acc;
    }
}

// The same interpolation read straight from a pitched state array of the simulation's dtype (no staging):
// the value at output voxel (i, j, k) of zoom(restore(field), order=1).  Same expressions as
// resample_linear_kernel, so both paths return identical bits.
struct ResampleGeom {
    int IX, IY, IZ, lx, ly, lz, VX, VY, VZ, OX, OY, OZ;
    int64_t sx, sy;
    double fx, fy, fz;
};

template <typename T> __device__ __forceinline__ double resample_at(const T *__restrict__ in, const ResampleGeom &g, int i, int j, int k)
{
    const double xi = i * g.fx, yj = j * g.fy, zk = k * g.fz;
    int i0 = (int)floor(xi), j0 = (int)floor(yj), k0 = (int)floor(zk);
    i0 = min(i0, max(g.VX - 2, 0)); j0 = min(j0, max(g.VY - 2, 0)); k0 = min(k0, max(g.VZ - 2, 0));
    const double tx = xi - i0, ty = yj - j0, tz = zk - k0;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int ii = min(i0 + a, g.VX - 1) - g.lx, jj = min(j0 + b, g.VY - 1) - g.ly, kk = min(k0 + c, g.VZ - 1) - g.lz;
                const double w = (a ? tx : 1.0 - tx) * (b ? ty : 1.0 - ty) * (c ? tz : 1.0 - tz);
                double v = 0.0;
                if (ii >= 0 && ii < g.IX && jj >= 0 && jj < g.IY && kk >= 0 && kk < g.IZ) v = (double)in[(int64_t)ii * g.sx + (int64_t)jj * g.sy + kk];
                acc += w * v;
            }
    return acc;
}

template <typename T> __global__ void resample_state_kernel(const T *__restrict__ in, ResampleGeom g, double *__restrict__ out)
{
    const int64_t n = (int64_t)g.OX * g.OY * g.OZ;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(q % g.OZ);
        const int64_t r = q / g.OZ;
        out[q] = resample_at(in, g, (int)(r / g.OY), (int)(r % g.OY), k);
    }
}

// create_segmentation_map (synthetic_gens/generatePatient_FK_2c.py:67-83) of the full-resolution P and N,
// evaluated on the fly: 3 edema (th_edema_p <= P < th_enhancing_p), 1 enhancing (P >= th_enhancing_p),
// 4 necrotic (N > th_necro_n, wins), 0 elsewhere.
template <typename T>
__global__ void segmentation_kernel(const T *__restrict__ P, const T *__restrict__ N, ResampleGeom g, double th_necro_n,
                                    double th_enhancing_p, double th_edema_p, uint8_t *__restrict__ out)
{
    const int64_t n = (int64_t)g.OX * g.OY * g.OZ;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(q % g.OZ);
        const int64_t r = q / g.OZ;
        const int j = (int)(r % g.OY), i = (int)(r / g.OY);
        const double p = resample_at(P, g, i, j, k), nn = resample_at(N, g, i, j, k);
        uint8_t lab = 0;
        if (p >= th_edema_p && p < th_enhancing_p) lab = 3;
        if (p >= th_enhancing_p) lab = 1;
        if (nn > th_necro_n) lab = 4;
        out[q] = lab;
    }
}

// ------------------------------------------------------------------------------------------
extern "C" {

int tgtk_abi_version(void) { return TGTK_ABI_VERSION; }

const char *tgtk_last_error(void) { return g_err.c_str(); }

int tgtk_device_count(int *count)
{
    CU(cudaGetDeviceCount(count));
    return 0;
}

int tgtk_sim_create(tgtk_sim **out, const tgtk_params *params, int device, void *stream)
{
    return tgtk_sim_create_ex(out, params, device, stream, 0);
}

int tgtk_sim_create_ex(tgtk_sim **out, const tgtk_params *params, int device, void *stream, int flags)
{
    if (!out || !params) return fail("tgtk_sim_create", "null argument");
    const tgtk_params &p = *params;
    if (p.X < 1 || p.Y < 1 || p.Z < 1) return fail("tgtk_sim_create", "empty box");
    if (p.model < 0 || p.model > TGTK_MODEL_DTI_FULL) return fail("tgtk_sim_create", "unknown model");
    if (p.dtype != TGTK_F64 && p.dtype != TGTK_F32) return fail("tgtk_sim_create", "unknown dtype");
    if (!(p.h[0] > 0 && p.h[1] > 0 && p.h[2] > 0)) return fail("tgtk_sim_create", "non-positive grid spacing");
    CU(cudaSetDevice(device));
    keep_pool_memory(device);
    tgtk_sim *s = new (std::nothrow) tgtk_sim();
    if (!s) return fail("tgtk_sim_create", "out of host memory");
    s->p = p;
    s->ipc = (flags & TGTK_CREATE_IPC) != 0;
    s->device = device;
    if (stream) {
        s->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete s; return fail("cudaStreamCreate", cudaGetErrorString(e)); }
        s->own_stream = true;
    }
    s->n = (size_t)p.X * p.Y * p.Z;
    s->elem = (p.dtype == TGTK_F64) ? 8 : 4;
    // fp32 rows have an even pitch: every (row, even z) pair is 8-byte aligned for the z-pair kernels.  fp64 rows are
    // dense -- except when the TMA-staged kernel is asked for at creation (TGTK_VARIANT=7), whose tensor maps need
    // 16-byte global strides; the padded fp64 pitch costs the default kernels 1-2 % (profiles/r2_sweep_misc.txt)
    {
        const char *v = getenv("TGTK_VARIANT"), *e = getenv("TGTK_EVEN_PITCH64");
        const bool even64 = e ? atoi(e) != 0 : (v && atoi(v) == 7);
        s->sy = p.Z + ((p.dtype == TGTK_F32 || even64) ? (p.Z & 1) : 0);
        if (even64 && p.dtype == TGTK_F32) s->sy = (p.Z + 3) / 4 * 4;      // 16-byte rows in fp32 too
        // TGTK_PITCH_ALIGN=bytes: rows start on a multiple of that many bytes (128: a warp's 32-voxel row segment then touches
        // 2 lines in fp64 / 1 in fp32 instead of 3 / 2 -- the marching kernels are bound by L1 wavefronts)
        if (const char *pa = getenv("TGTK_PITCH_ALIGN")) {
            const int64_t el = std::max<int64_t>(1, atoi(pa) / (int64_t)s->elem);
            s->sy = (s->sy + el - 1) / el * el;
        }
    }
    s->sx = (int64_t)p.Y * s->sy;
    s->nfields = (p.model == TGTK_MODEL_FK_2C || p.model == TGTK_MODEL_FK2C_COEF12) ? 3 : 1;
    switch (p.model) {
    case TGTK_MODEL_FK: s->ncoef = 1; break;
    case TGTK_MODEL_FK_2C: s->ncoef = 2; break;
    case TGTK_MODEL_DTI: s->ncoef = 3; break;
    case TGTK_MODEL_COEF6: s->ncoef = 6; break;
    case TGTK_MODEL_FK_FACES: s->ncoef = 3; break;
    case TGTK_MODEL_FK2C_COEF12: s->ncoef = 12; break;
    case TGTK_MODEL_DTI_FULL: s->ncoef = 6; break;
    }
    if (const char *v = getenv("TGTK_VARIANT")) s->variant = atoi(v);
    if (const char *v = getenv("TGTK_MINB")) s->minb = atoi(v);
    if (const char *v = getenv("TGTK_PF")) s->pf = atoi(v);
    if (const char *v = getenv("TGTK_V4_THREADS")) s->v4_threads = atoi(v);
    if (const char *v = getenv("TGTK_NARROW_PENALTY")) s->narrow_penalty = atof(v);
    if (const char *v = getenv("TGTK_GRAPH")) s->use_graph = atoi(v) != 0;
    if (const char *v = getenv("TGTK_TZ")) s->force_tz = atoi(v);
    if (const char *v = getenv("TGTK_LX")) s->force_lx = atoi(v);
    if (const char *v = getenv("TGTK_L2PF")) s->l2pf = atoi(v);
    if (const char *v = getenv("TGTK_PDL")) s->pdl = atoi(v);
    if (const char *v = getenv("TGTK_TB")) s->tb = atoi(v);
    if (const char *v = getenv("TGTK_ZIGZAG")) s->zigzag = atoi(v) != 0;
    if (const char *v = getenv("TGTK_CLOSED_SKIP")) s->closed_skip = atoi(v) != 0;
    if (const char *v = getenv("TGTK_SPARSE")) s->sparse = atoi(v);
    if (const char *v = getenv("TGTK_SP_WAVES")) { s->sp_waves = std::max(0.25, atof(v)); s->sp_waves_set = true; }
    if (const char *v = getenv("TGTK_SP_ASYNC")) s->sp_async = atoi(v);
    if (const char *v = getenv("TGTK_SPARSE_MIN_VOXELS")) s->sp_min_voxels = atoll(v);
    if (const char *v = getenv("TGTK_SP_AHEAD")) s->sp_ahead = atoi(v);
    if (const char *v = getenv("TGTK_SP_SPLIT")) s->sp_split = std::min(0.9, std::max(0.1, atof(v)));
    if (const char *v = getenv("TGTK_SP_CARVEOUT")) s->sp_carveout = atoi(v);
    if (const char *v = getenv("TGTK_SP_MINB")) s->sp_minb = atoi(v);
    if (const char *v = getenv("TGTK_TB_ROWS")) s->tb_rows = atoi(v);
    if (const char *v = getenv("TGTK_TMA_ROWS")) s->tma_rows = (atoi(v) == 8) ? 8 : 4;
    if (const char *v = getenv("TGTK_TMA_STAGES")) s->tma_stages = (atoi(v) == 2) ? 2 : (atoi(v) == 4) ? 4 : 3;
    if (const char *v = getenv("TGTK_TMA_MINB")) s->tma_minb = atoi(v);
    if (cudaDeviceGetAttribute(&s->l2_bytes, cudaDevAttrL2CacheSize, s->device) != cudaSuccess) s->l2_bytes = 0;
    *out = s;  // from here on tgtk_sim_destroy cleans up on failure
    if (configure_launch(s)) return 1;
    const size_t bytes = (size_t)p.X * s->sx * s->elem;
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < s->nfields; ++f) {
            if (s->ipc) CU(cudaMalloc(&s->state[b][f], bytes));   // pool memory cannot be exported over IPC
            else CU(sim_malloc(s, (void **)(&s->state[b][f]), bytes));
            CU(cudaMemsetAsync(s->state[b][f], 0, bytes, s->stream));
        }
    for (int c = 0; c < s->ncoef; ++c) {
        CU(sim_malloc(s, (void **)(&s->coef[c]), bytes));
        CU(cudaMemsetAsync(s->coef[c], 0, bytes, s->stream));
    }
    CU(sim_malloc(s, (void **)(&s->stage), s->n * 8));
    CU(sim_malloc(s, (void **)(&s->ctrl), sizeof(Ctrl)));
    s->ctrl_host = pinned_ctrl_get();
    if (!s->ctrl_host) return fail("cudaMallocHost", "cannot allocate the pinned control mirror");
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    Ctrl init = {0, 0, -1, 0u, 0.0};
    *s->ctrl_host = init;
    CU(cudaMemcpyAsync(s->ctrl, s->ctrl_host, sizeof(Ctrl), cudaMemcpyHostToDevice, s->stream));
    if (ensure_volumes(s, 1024)) return 1;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

void tgtk_sim_destroy(tgtk_sim *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < 3; ++f) {
            if (s->ipc) cudaFree(s->state[b][f]);
            else sim_free(s, s->state[b][f]);
        }
    for (int c = 0; c < 12; ++c) sim_free(s, s->coef[c]);
    for (void *q : s->snaps) sim_free(s, q);
    sim_free(s, s->stage);
    sim_free(s, s->sp_live);
    sim_free(s, s->sp_work);
    sim_free(s, s->resample_out);
    for (int f = 0; f < 3; ++f) sim_free(s, s->resample_async[f]);
    sim_free(s, s->seg_out);
    for (int f = 0; f < 3; ++f) sim_free(s, s->initial[f]);
    for (int f = 0; f < 3; ++f) sim_free(s, s->ckpt[f]);
    if (s->ev2) cudaEventDestroy(s->ev2);
    if (s->ev3) cudaEventDestroy(s->ev3);
    sim_free(s, s->wm);
    sim_free(s, s->gm);
    sim_free(s, s->pc);
    sim_free(s, s->pn);
    sim_free(s, s->ctrl);
    sim_free(s, s->partials);
    sim_free(s, s->volumes);
    sim_free(s, s->chunkvol);
    pinned_ctrl_put(s->ctrl_host);
    slab_p2p_close(s);
    drop_graph(s);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int tgtk_sim_set_variant(tgtk_sim *s, int variant)
{
    if (!s) return fail("tgtk_sim_set_variant", "null argument");
    if (variant < 0 || variant > 7) return fail("tgtk_sim_set_variant", "variant must be 0 (auto) or 1..7");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    s->variant = variant;
    s->pads_dirty = true;
    return configure_launch(s);
}

int tgtk_sim_get_tuning(tgtk_sim *s, int32_t out[10])
{
    if (!s || !out) return fail("tgtk_sim_get_tuning", "null argument");
    const bool marching = s->lx > 0;
    out[0] = marching ? effective_variant(s) : 1;
    out[1] = s->lx;
    out[2] = (marching && effective_variant(s) >= 4) ? effective_l2pf(s) : 0;
    out[3] = (marching && effective_variant(s) >= 4 && effective_pdl(s)) ? 1 : 0;
    out[4] = (int32_t)s->grid.x; out[5] = (int32_t)s->grid.y; out[6] = (int32_t)s->grid.z;
    out[7] = (int32_t)(s->block.x * s->block.y * s->block.z);
    out[8] = tb_enabled(s) ? 2 : 1;
    out[9] = s->sp_active ? s->sp_items : 0;       // work items of a sparse launch (0: dense grid)
    return 0;
}

int tgtk_sparse_plan_host(const uint32_t *mask, int X, int nzt, int nyt, int x_lo, int x_hi, int slots, int slab, int32_t *items,
                          int cap, int *count)
{
    if (!mask || !count || X < 1 || nzt < 1 || nyt < 1 || x_lo < 0 || x_hi > X || x_lo >= x_hi || slots < 1)
        return fail("tgtk_sparse_plan_host", "bad argument");
    PlanOpts o;
    o.slab = slab != 0;
    if (const char *v = getenv("TGTK_SP_WAVES")) { o.waves = std::max(0.25, atof(v)); o.waves_set = true; }
    if (const char *v = getenv("TGTK_SP_SPLIT")) o.split = std::min(0.9, std::max(0.1, atof(v)));
    std::vector<int4> work;
    plan_work_list(mask, X, nzt, nyt, x_lo, x_hi, slots, o, work);
    *count = (int)(work.size() / 2);
    if (items) {
        if (cap < *count) return fail("tgtk_sparse_plan_host", "output too small");
        for (int i = 0; i < *count; ++i) {
            items[4 * i] = work[2 * i].x; items[4 * i + 1] = work[2 * i].y; items[4 * i + 2] = work[2 * i].z; items[4 * i + 3] = work[2 * i].w;
        }
    }
    return 0;
}

int tgtk_sim_sparse_info(tgtk_sim *s, int64_t out[4])
{
    if (!s || !out) return fail("tgtk_sim_sparse_info", "null argument");
    if (!s->prepared) return fail("tgtk_sim_sparse_info", "call tgtk_sim_prepare first");
    CU(cudaSetDevice(s->device));
    if (!s->slab_on && (s->sp_dirty || s->sp_active != sparse_eligible(s)) && (s->sp_active || sparse_eligible(s)) && build_sparse(s)) return 1;
    out[0] = s->sp_active ? s->sp_items : 0;
    out[1] = s->sp_active ? s->sp_live_voxels : (int64_t)s->n;
    out[2] = (int64_t)s->n;
    out[3] = s->sp_active ? (int64_t)s->sp_per_sm : 0;      // resident blocks per SM of the sparse kernel
    return 0;
}

/* diagnostic: run ONE more step with per-item timestamps (nanoseconds of the global timer, relative to the earliest block
 * start): out[4*i + 0..3] = block start, end of the prologue, end of the plane loop, SM id; plus the item's planes in
 * planes[2*i + 0..1] */
int tgtk_sim_sparse_trace(tgtk_sim *s, int64_t *out, int32_t *planes, int cap)
{
    if (!s || !out) return fail("tgtk_sim_sparse_trace", "null argument");
    if (!s->sp_active) return fail("tgtk_sim_sparse_trace", "the simulation does not run sparse launches");
    if (cap < s->sp_items) return fail("tgtk_sim_sparse_trace", "output too small");
    CU(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->sp_items * 4 * sizeof(unsigned long long);
    const size_t bytes_all = bytes + (size_t)s->sp_items * 32 * sizeof(unsigned long long);     // + the phase counters of a -DTGTK_SP_PHASES build
    CU(sim_malloc(s, (void **)(&s->sp_trace), bytes_all));
    CU(cudaMemsetAsync(s->sp_trace, 0, bytes_all, s->stream));
    const bool g = s->use_graph;
    s->use_graph = false;
    int rc = tgtk_sim_run(s, 1, nullptr, 0);
    s->use_graph = g;
    std::vector<unsigned long long> t((size_t)s->sp_items * 4);
    std::vector<int4> w2((size_t)s->sp_items * 2);
    if (!rc) {
        CU(cudaMemcpyAsync(t.data(), s->sp_trace, bytes, cudaMemcpyDeviceToHost, s->stream));
        if (const char *v = getenv("TGTK_SP_PHASES_OUT")) {      // diagnostic build: raw phase counters to a file
            std::vector<unsigned long long> ph((size_t)s->sp_items * 32);
            CU(cudaMemcpyAsync(ph.data(), (char *)s->sp_trace + bytes, ph.size() * 8, cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            if (FILE *f = fopen(v, "wb")) { fwrite(ph.data(), 8, ph.size(), f); fclose(f); }
        }
        CU(cudaMemcpyAsync(w2.data(), s->sp_work, w2.size() * sizeof(int4), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    sim_free(s, s->sp_trace);
    s->sp_trace = nullptr;
    if (rc) return rc;
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < s->sp_items; ++i) t0 = std::min(t0, t[4 * (size_t)i]);
    for (int i = 0; i < s->sp_items; ++i) {
        for (int k = 0; k < 3; ++k) out[4 * (size_t)i + k] = (int64_t)(t[4 * (size_t)i + k] - t0);
        out[4 * (size_t)i + 3] = (int64_t)t[4 * (size_t)i + 3];
        if (planes) { planes[2 * i] = w2[2 * i].z; planes[2 * i + 1] = w2[2 * i].w; }
    }
    return 0;
}

int tgtk_sim_set_sum_range(tgtk_sim *s, int x0, int x1)
{
    if (!s) return fail("tgtk_sim_set_sum_range", "null argument");
    if (x0 < 0 || x1 > s->p.X || x0 > x1) return fail("tgtk_sim_set_sum_range", "range outside the box");
    CU(cudaSetDevice(s->device));
    s->sum_x0 = x0;
    s->sum_x1 = x1;
    drop_graph(s);
    return 0;
}

int tgtk_sim_buffer_ptr(tgtk_sim *s, int field, int which, unsigned long long *ptr)
{
    if (!s || !ptr) return fail("tgtk_sim_buffer_ptr", "null argument");
    if (field < 0 || field >= s->nfields || which < 0 || which > 1) return fail("tgtk_sim_buffer_ptr", "no such buffer");
    *ptr = (unsigned long long)(uintptr_t)s->state[which][field];
    s->sp_forbid = true;       // the caller writes the buffers behind the library's back (halo exchanges): every tile is stored
    return 0;
}

int tgtk_sim_upload(tgtk_sim *s, int field, const double *host, const int64_t strides[3], int pinned)
{
    (void)pinned;
    if (!s || !host || !strides) return fail("tgtk_sim_upload", "null argument");
    CU(cudaSetDevice(s->device));
    const tgtk_params &p = s->p;
    if (field >= TGTK_FIELD_WM && field <= TGTK_FIELD_NC) {
        if (p.model != TGTK_MODEL_FK && p.model != TGTK_MODEL_FK_2C && p.model != TGTK_MODEL_FK_FACES)
            return fail("tgtk_sim_upload", "model takes no tissue maps");
        double **dst = (field == TGTK_FIELD_WM) ? &s->wm : (field == TGTK_FIELD_GM) ? &s->gm
                       : (field == TGTK_FIELD_PC) ? &s->pc : &s->pn;
        if (!*dst) CU(sim_malloc(s, (void **)(dst), s->n * 8));
        if (copy_host_stage(s, const_cast<double *>(host), strides, true)) return 1;
        CU(cudaMemcpyAsync(*dst, s->stage, s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
        s->prepared = false;
        s->sp_dirty = true;
        return 0;
    }
    const int cur = s->t_enqueued & 1;
    s->sp_dirty = true;
    void *dst = field_ptr(s, field, cur);
    if (!dst) return fail("tgtk_sim_upload", "field not present in this model");
    if (field >= TGTK_FIELD_COEF0 && p.model != TGTK_MODEL_DTI && p.model != TGTK_MODEL_COEF6 &&
        p.model != TGTK_MODEL_FK2C_COEF12 && p.model != TGTK_MODEL_DTI_FULL)
        return fail("tgtk_sim_upload", "coefficient arrays of this model are built by tgtk_sim_prepare");
    if (copy_host_stage(s, const_cast<double *>(host), strides, true)) return 1;
    s->pads_dirty = true;
    if (p.dtype == TGTK_F64)
        pack_kernel<double><<<grid_1d(s->n), 256, 0, s->stream>>>(s->stage, (double *)dst, p.X, p.Y, p.Z, s->sx, s->sy);
    else
        pack_kernel<float><<<grid_1d(s->n), 256, 0, s->stream>>>(s->stage, (float *)dst, p.X, p.Y, p.Z, s->sx, s->sy);
    CU(cudaGetLastError());
    return 0;
}

static int download_ptr(tgtk_sim *s, const void *src, double *host, const int64_t strides[3])
{
    const tgtk_params &p = s->p;
    if (p.dtype == TGTK_F64)
        unpack_kernel<double><<<grid_1d(s->n), 256, 0, s->stream>>>((const double *)src, s->stage, p.X, p.Y, p.Z, s->sx, s->sy);
    else
        unpack_kernel<float><<<grid_1d(s->n), 256, 0, s->stream>>>((const float *)src, s->stage, p.X, p.Y, p.Z, s->sx, s->sy);
    CU(cudaGetLastError());
    if (copy_host_stage(s, host, strides, false)) return 1;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// index of the buffer that holds the current state: step t writes buf[(t+1)&1]
static int current_buf(tgtk_sim *s, int *steps_run)
{
    const Ctrl &c = *s->ctrl_host;
    if (steps_run) *steps_run = c.t;
    return c.t & 1;
}

// Speculative mode: a stop test tripped inside a chunk whose later steps had already run.  Go back to the
// checkpoint taken before the chunk and redo exactly the steps up to and including the stopping one
// (FK.py:216-218: the state after the step whose volume tripped the test is the final state).
static int replay_to_stop(tgtk_sim *s)
{
    const int t0 = s->ctrl_host->t, n = s->ctrl_host->replay;
    if (n < 1 || n > kGraphSteps) return fail("replay", "inconsistent loop control");
    const size_t bytes = (size_t)s->p.X * s->sx * s->elem;
    CU(cudaEventRecord(s->ev2, s->stream));
    for (int f = 0; f < s->nfields; ++f)
        CU(cudaMemcpyAsync(s->state[t0 & 1][f], s->ckpt[f], bytes, cudaMemcpyDeviceToDevice, s->stream));
    s->replaying = true;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n && e == cudaSuccess; ++i) e = launch_one(s, (t0 + i) & 1, i);
    if (e == cudaSuccess) e = launch_chunk_finish(s, n);   // logs the volumes again, t = t0 + n, replay = 0
    s->replaying = false;
    if (e != cudaSuccess) return fail("replay launch", cudaGetErrorString(e));
    CU(cudaEventRecord(s->ev3, s->stream));
    s->launches += n;
    CU(cudaMemcpyAsync(s->ctrl_host, s->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev2, s->ev3));
    s->replay_ms += ms;
    s->t_enqueued = s->ctrl_host->t;
    s->pads_dirty = true;
    return 0;
}

static int fetch_ctrl(tgtk_sim *s)
{
    CU(cudaMemcpyAsync(s->ctrl_host, s->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (s->spec && s->ctrl_host->replay > 0 && replay_to_stop(s)) return 1;
    s->ctrl_valid = true;
    return 0;
}

int tgtk_sim_download(tgtk_sim *s, int field, double *host, const int64_t strides[3])
{
    if (!s || !host || !strides) return fail("tgtk_sim_download", "null argument");
    CU(cudaSetDevice(s->device));
    if (fetch_ctrl(s)) return 1;
    void *src = field_ptr(s, field, current_buf(s, nullptr));
    if (!src) return fail("tgtk_sim_download", "field not present in this model");
    return download_ptr(s, src, host, strides);
}

int tgtk_sim_prepare(tgtk_sim *s)
{
    if (!s) return fail("tgtk_sim_prepare", "null argument");
    CU(cudaSetDevice(s->device));
    const tgtk_params &p = s->p;
    const bool tissue_model = (p.model == TGTK_MODEL_FK || p.model == TGTK_MODEL_FK_2C || p.model == TGTK_MODEL_FK_FACES);
    if (tissue_model) {
        if (!s->wm || !s->gm) return fail("tgtk_sim_prepare", "upload TGTK_FIELD_WM and TGTK_FIELD_GM first");
        const int g = grid_1d(s->n);
        if (p.model == TGTK_MODEL_FK_FACES) {
            if (p.dtype == TGTK_F64)
                prep_faces_kernel<double><<<g, 256, 0, s->stream>>>(s->wm, s->gm, s->pc, s->pn, (double *)s->coef[0],
                                                                   (double *)s->coef[1], (double *)s->coef[2], p.th_matter,
                                                                   p.th_necro, p.Dw, p.ratio, p.logical_or_faces,
                                                                   s->faces_what, p.X, p.Y, p.Z, s->sx, s->sy);
            else
                prep_faces_kernel<float><<<g, 256, 0, s->stream>>>(s->wm, s->gm, s->pc, s->pn, (float *)s->coef[0],
                                                                  (float *)s->coef[1], (float *)s->coef[2], p.th_matter,
                                                                  p.th_necro, p.Dw, p.ratio, p.logical_or_faces,
                                                                  s->faces_what, p.X, p.Y, p.Z, s->sx, s->sy);
        } else {
            if (p.dtype == TGTK_F64)
                prep_compact_kernel<double><<<g, 256, 0, s->stream>>>(s->wm, s->gm, (double *)s->coef[0], (double *)s->coef[1],
                                                                     p.th_matter, p.ratio, p.X, p.Y, p.Z, s->sx, s->sy);
            else
                prep_compact_kernel<float><<<g, 256, 0, s->stream>>>(s->wm, s->gm, (float *)s->coef[0], (float *)s->coef[1],
                                                                    p.th_matter, p.ratio, p.X, p.Y, p.Z, s->sx, s->sy);
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s->stream));
        CU(sim_free(s, s->wm));
        CU(sim_free(s, s->gm));
        CU(sim_free(s, s->pc));
        CU(sim_free(s, s->pn));
        s->wm = s->gm = s->pc = s->pn = nullptr;
    }
    // no stop criterion active -> host-scheduled launches (see StepArgs::host_sched)
    {
        const bool want = std::isinf(p.stopping_volume) && p.stopping_volume > 0 && !(p.small_volume > 0);
        const char *v = getenv("TGTK_HOST_SCHED");
        const bool host_sched = want && !(v && atoi(v) == 0);
        const char *w = getenv("TGTK_SPECULATE");
        const bool spec = !want && (s->slab_on || !(w && atoi(w) == 0));   // a decomposed run has no device-scheduled mode
        if (host_sched != s->host_sched || spec != s->spec) drop_graph(s);
        s->host_sched = host_sched;
        s->spec = spec;
        if (spec) {
            const size_t bytes = (size_t)p.X * s->sx * s->elem;
            for (int f = 0; f < s->nfields; ++f)
                if (!s->ckpt[f]) CU(sim_malloc(s, &s->ckpt[f], bytes));
            if (!s->ev2) { CU(cudaEventCreate(&s->ev2)); CU(cudaEventCreate(&s->ev3)); }
        }
    }
    if (refresh_pads(s)) return 1;
    // keep the initial state so a finished loop can be rerun without another upload
    const size_t bytes = (size_t)p.X * s->sx * s->elem;
    const int cur = s->t_enqueued & 1;
    for (int f = 0; f < s->nfields; ++f) {
        if (!s->initial[f]) CU(sim_malloc(s, (void **)(&s->initial[f]), bytes));
        CU(cudaMemcpyAsync(s->initial[f], s->state[cur][f], bytes, cudaMemcpyDeviceToDevice, s->stream));
    }
    s->prepared = true;
    s->sp_dirty = true;
    return 0;
}

int tgtk_sim_reset(tgtk_sim *s)
{
    if (!s) return fail("tgtk_sim_reset", "null argument");
    if (!s->prepared) return fail("tgtk_sim_reset", "call tgtk_sim_prepare first");
    CU(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->p.X * s->sx * s->elem;
    for (int f = 0; f < s->nfields; ++f)
        CU(cudaMemcpyAsync(s->state[0][f], s->initial[f], bytes, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));   // ctrl_host is about to be rewritten
    s->ctrl_valid = false;
    Ctrl init = {0, 0, -1, 0u, 0.0};
    *s->ctrl_host = init;
    CU(cudaMemcpyAsync(s->ctrl, s->ctrl_host, sizeof(Ctrl), cudaMemcpyHostToDevice, s->stream));
    for (void *q : s->snaps) sim_free(s, q);
    s->snaps.clear();
    s->snap_steps.clear();
    s->t_enqueued = 0;
    return 0;
}

int tgtk_sim_run(tgtk_sim *s, int nsteps, const int32_t *record_steps, int nrecord)
{
    if (!s) return fail("tgtk_sim_run", "null argument");
    if (!s->prepared) return fail("tgtk_sim_run", "call tgtk_sim_prepare first");
    if (nsteps < 0) return fail("tgtk_sim_run", "negative step count");
    CU(cudaSetDevice(s->device));
    if (s->pads_dirty && zpair_variant(s) && refresh_pads(s)) return 1;
    if (s->slab_on) {
        if (s->sp_dirty && s->sp_active) { s->sp_active = false; drop_graph(s); }      // see tgtk_sim_slab_attach
    } else if ((s->sp_dirty || s->sp_active != sparse_eligible(s)) && (s->sp_active || sparse_eligible(s)) && build_sparse(s)) return 1;
    if (ensure_volumes(s, s->t_enqueued + nsteps)) return 1;
    s->ctrl_valid = false;
    const size_t bytes = (size_t)s->p.X * s->sx * s->elem;
    int r = 0;
    while (r < nrecord && record_steps[r] < s->t_enqueued) ++r;
    // snapshot arenas are allocated up front: cudaMalloc inside the loop would serialise it
    int nsnap = 0;
    for (int i = r; i < nrecord; ++i)
        if (record_steps[i] < s->t_enqueued + nsteps && (i == r || record_steps[i] != record_steps[i - 1])) ++nsnap;
    std::vector<void *> arenas(nsnap, nullptr);
    int used = 0;
    // arenas not handed to s->snaps go back to the pool on EVERY exit, the error returns included
    struct ArenaGuard {
        tgtk_sim *s; std::vector<void *> &a; int &used;
        ~ArenaGuard() { for (size_t i = (size_t)used; i < a.size(); ++i) sim_free(s, a[i]); }
    } arena_guard{s, arenas, used};
    for (int i = 0; i < nsnap; ++i) CU(sim_malloc(s, (void **)(&arenas[i]), bytes * s->nfields));
    s->replay_ms = 0.f;
    CU(cudaEventRecord(s->ev0, s->stream));
    int q = 0;
    while (q < nsteps) {
        // steps up to and including the next record step (or the end of this run)
        const int t0 = s->t_enqueued;
        const int next_rec = (r < nrecord) ? record_steps[r] : 0x7fffffff;
        const int run_len = (int)std::min<int64_t>(nsteps - q, (int64_t)next_rec - t0 + 1);
        int done = 0;
        // identical launches (the step index lives in device memory) replayed from a CUDA graph
        while (s->use_graph && run_len - done >= kGraphSteps) {
            const int par0 = (t0 + done) & 1;
            if (!s->graph_exec[par0] && build_graph(s, par0)) return 1;
            CU(cudaGraphLaunch(s->graph_exec[par0], s->stream));
            s->launches += s->graph_launches;
            done += kGraphSteps;
        }
        while (done < run_len) {
            // leftovers: plain launches, in host-scheduled mode closed by one chunk_finish
            const int n = std::min(run_len - done, kGraphSteps);
            cudaError_t e = s->spec ? launch_checkpoint(s, (t0 + done) & 1) : cudaSuccess;
            if (e == cudaSuccess) e = launch_steps(s, (t0 + done) & 1, n);
            if (e == cudaSuccess && (s->host_sched || s->spec)) e = launch_chunk_finish(s, n);
            if (e != cudaSuccess) return fail("step launch", cudaGetErrorString(e));
            done += n;
        }
        s->t_enqueued += run_len;
        q += run_len;
        const int t = s->t_enqueued - 1;
        if (t == next_rec) {
            // state after step t lives in buf[(t+1)&1]; FK/FK.py:221-223
            void *arena = arenas[used++];
            for (int f = 0; f < s->nfields; ++f)
                CU(cudaMemcpyAsync((char *)arena + bytes * f, s->state[(t + 1) & 1][f], bytes, cudaMemcpyDeviceToDevice,
                                   s->stream));
            s->snaps.push_back(arena);
            s->snap_steps.push_back(t);
            while (r < nrecord && record_steps[r] == t) ++r;
        }
    }
    CU(cudaEventRecord(s->ev1, s->stream));
    s->timed = true;
    return 0;
}

int tgtk_sim_sync(tgtk_sim *s, tgtk_status *st)
{
    if (!s) return fail("tgtk_sim_sync", "null argument");
    CU(cudaSetDevice(s->device));
    if (fetch_ctrl(s)) return 1;
    if (st) {
        const Ctrl &c = *s->ctrl_host;
        st->steps_run = c.t;
        st->stop_reason = c.stop_reason;
        st->stop_step = c.stop_step;
        st->last_volume = c.last_volume;
        int valid = 0;
        // a snapshot of step t is valid only if the loop got past step t's stop test
        for (int q : s->snap_steps)
            if (q < c.t && !(c.stop_reason != 0 && q >= c.stop_step)) ++valid;
        st->snapshots = valid;
        float ms = 0.f;
        if (s->timed) CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        st->loop_ms = ms + s->replay_ms;   // a speculative run's replay belongs to the loop
        st->kernel_launches = s->launches;
    }
    return 0;
}

int tgtk_sim_volumes(tgtk_sim *s, int first, int count, double *out)
{
    if (!s || !out) return fail("tgtk_sim_volumes", "null argument");
    if (first < 0 || count < 0 || first + count > s->volumes_cap) return fail("tgtk_sim_volumes", "range outside the log");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(out, s->volumes + first, sizeof(double) * count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int tgtk_sim_snapshot(tgtk_sim *s, int index, int field, double *host, const int64_t strides[3])
{
    if (!s || !host || !strides) return fail("tgtk_sim_snapshot", "null argument");
    if (index < 0 || index >= (int)s->snaps.size()) return fail("tgtk_sim_snapshot", "no such snapshot");
    if (field < 0 || field >= s->nfields) return fail("tgtk_sim_snapshot", "no such field");
    CU(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->p.X * s->sx * s->elem;
    return download_ptr(s, (char *)s->snaps[index] + bytes * field, host, strides);
}

int tgtk_sim_device_ptr(tgtk_sim *s, int field, unsigned long long *ptr, int64_t pitches[3], int *elem_size)
{
    if (!s || !ptr) return fail("tgtk_sim_device_ptr", "null argument");
    CU(cudaSetDevice(s->device));
    if (fetch_ctrl(s)) return 1;
    void *q = field_ptr(s, field, current_buf(s, nullptr));
    if (!q) return fail("tgtk_sim_device_ptr", "field not present in this model");
    *ptr = (unsigned long long)(uintptr_t)q;
    if (pitches) { pitches[0] = s->sx; pitches[1] = s->sy; pitches[2] = 1; }
    if (elem_size) *elem_size = (int)s->elem;
    return 0;
}

int tgtk_fk_update_host(const tgtk_params *params, int device, double *A, const int64_t a_strides[3],
                        const double *const coef[6], const int64_t coef_strides[6][3])
{
    if (!params || !A || !coef) return fail("tgtk_fk_update_host", "null argument");
    tgtk_params p = *params;
    p.model = TGTK_MODEL_COEF6;
    p.stopping_volume = INFINITY;
    p.small_volume = 0.0;
    tgtk_sim *s = nullptr;
    int rc = tgtk_sim_create(&s, &p, device, nullptr);
    if (!rc) rc = tgtk_sim_upload(s, TGTK_FIELD_U, A, a_strides, 0);
    for (int c = 0; c < 6 && !rc; ++c) rc = tgtk_sim_upload(s, TGTK_FIELD_COEF0 + c, coef[c], coef_strides[c], 0);
    if (!rc) rc = tgtk_sim_prepare(s);
    if (!rc) rc = tgtk_sim_run(s, 1, nullptr, 0);
    if (!rc) rc = tgtk_sim_download(s, TGTK_FIELD_U, A, a_strides);
    tgtk_sim_destroy(s);
    return rc;
}

int tgtk_fk2c_update_host(const tgtk_params *params, int device, double *P, const int64_t p_strides[3],
                          double *N, const int64_t n_strides[3], double *S, const int64_t s_strides[3],
                          const double *const D[6], const int64_t d_strides[6][3],
                          const double *const Ds[6], const int64_t ds_strides[6][3])
{
    if (!params || !P || !N || !S || !D || !Ds) return fail("tgtk_fk2c_update_host", "null argument");
    tgtk_params p = *params;
    p.model = TGTK_MODEL_FK2C_COEF12;
    p.stopping_volume = INFINITY;
    p.small_volume = 0.0;
    tgtk_sim *s = nullptr;
    int rc = tgtk_sim_create(&s, &p, device, nullptr);
    if (!rc) rc = tgtk_sim_upload(s, TGTK_FIELD_P, P, p_strides, 0);
    if (!rc) rc = tgtk_sim_upload(s, TGTK_FIELD_N, N, n_strides, 0);
    if (!rc) rc = tgtk_sim_upload(s, TGTK_FIELD_S, S, s_strides, 0);
    for (int c = 0; c < 6 && !rc; ++c) rc = tgtk_sim_upload(s, TGTK_FIELD_COEF0 + c, D[c], d_strides[c], 0);
    for (int c = 0; c < 6 && !rc; ++c) rc = tgtk_sim_upload(s, TGTK_FIELD_COEF0 + 6 + c, Ds[c], ds_strides[c], 0);
    if (!rc) rc = tgtk_sim_prepare(s);
    if (!rc) rc = tgtk_sim_run(s, 1, nullptr, 0);
    if (!rc) rc = tgtk_sim_download(s, TGTK_FIELD_P, P, p_strides);
    if (!rc) rc = tgtk_sim_download(s, TGTK_FIELD_N, N, n_strides);
    if (!rc) rc = tgtk_sim_download(s, TGTK_FIELD_S, S, s_strides);
    tgtk_sim_destroy(s);
    return rc;
}

int tgtk_faces_host(int device, int X, int Y, int Z, const double *wm, const int64_t wm_strides[3],
                    const double *gm, const int64_t gm_strides[3], const double *P, const int64_t p_strides[3],
                    const double *N, const int64_t n_strides[3], double th, double th_necro, double Dw,
                    double ratio, int logical_or, int what, double *out_x, double *out_y, double *out_z)
{
    if (!wm || !gm || !out_x || !out_y || !out_z) return fail("tgtk_faces_host", "null argument");
    if ((P == nullptr) != (N == nullptr)) return fail("tgtk_faces_host", "P and N must be given together");
    tgtk_params p;
    memset(&p, 0, sizeof(p));
    p.model = TGTK_MODEL_FK_FACES;
    p.dtype = TGTK_F64;
    p.X = X; p.Y = Y; p.Z = Z;
    p.logical_or_faces = logical_or;
    p.h[0] = p.h[1] = p.h[2] = 1.0;
    p.Dw = Dw; p.ratio = ratio; p.th_matter = th; p.th_necro = th_necro;
    p.stopping_volume = INFINITY;
    tgtk_sim *s = nullptr;
    int rc = tgtk_sim_create(&s, &p, device, nullptr);
    if (!rc) s->faces_what = what;
    if (!rc) rc = tgtk_sim_upload(s, TGTK_FIELD_WM, wm, wm_strides, 0);
    if (!rc) rc = tgtk_sim_upload(s, TGTK_FIELD_GM, gm, gm_strides, 0);
    if (!rc && P) rc = tgtk_sim_upload(s, TGTK_FIELD_PC, P, p_strides, 0);
    if (!rc && N) rc = tgtk_sim_upload(s, TGTK_FIELD_NC, N, n_strides, 0);
    if (!rc) rc = tgtk_sim_prepare(s);
    const int64_t dense_st[3] = {(int64_t)Y * Z, Z, 1};
    double *outs[3] = {out_x, out_y, out_z};
    for (int a = 0; a < 3 && !rc; ++a) rc = tgtk_sim_download(s, TGTK_FIELD_COEF0 + a, outs[a], dense_st);
    tgtk_sim_destroy(s);
    return rc;
}

int tgtk_sim_download_resampled(tgtk_sim *s, int field, const int32_t lo[3], const int32_t virt_shape[3], double *out,
                                const int32_t out_shape[3])
{
    if (!s || !lo || !virt_shape || !out || !out_shape) return fail("tgtk_sim_download_resampled", "null argument");
    const tgtk_params &p = s->p;
    const int in_shape[3] = {p.X, p.Y, p.Z};
    for (int d = 0; d < 3; ++d)
        if (virt_shape[d] < 1 || out_shape[d] < 1 || lo[d] < 0 || lo[d] + in_shape[d] > virt_shape[d])
            return fail("tgtk_sim_download_resampled", "bad shape or box");
    CU(cudaSetDevice(s->device));
    if (fetch_ctrl(s)) return 1;
    const void *src = field_ptr(s, field, current_buf(s, nullptr));
    if (!src) return fail("tgtk_sim_download_resampled", "field not present in this model");
    if (p.dtype == TGTK_F64)
        unpack_kernel<double><<<grid_1d(s->n), 256, 0, s->stream>>>((const double *)src, s->stage, p.X, p.Y, p.Z, s->sx, s->sy);
    else
        unpack_kernel<float><<<grid_1d(s->n), 256, 0, s->stream>>>((const float *)src, s->stage, p.X, p.Y, p.Z, s->sx, s->sy);
    CU(cudaGetLastError());
    const size_t nout = (size_t)out_shape[0] * out_shape[1] * out_shape[2];
    if (nout > s->resample_cap) {
        if (s->resample_out) CU(sim_free(s, s->resample_out));
        s->resample_out = nullptr;
        s->resample_cap = 0;
        CU(sim_malloc(s, (void **)(&s->resample_out), nout * 8));
        s->resample_cap = nout;
    }
    resample_linear_kernel<<<grid_1d(nout), 256, 0, s->stream>>>(s->stage, p.X, p.Y, p.Z, lo[0], lo[1], lo[2], virt_shape[0],
                                                               virt_shape[1], virt_shape[2], s->resample_out, out_shape[0],
                                                               out_shape[1], out_shape[2]);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, s->resample_out, nout * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

static int resample_geom(tgtk_sim *s, const char *who, const int32_t lo[3], const int32_t virt_shape[3], const int32_t out_shape[3],
                         ResampleGeom &g)
{
    const tgtk_params &p = s->p;
    const int in_shape[3] = {p.X, p.Y, p.Z};
    for (int d = 0; d < 3; ++d)
        if (virt_shape[d] < 1 || out_shape[d] < 1 || lo[d] < 0 || lo[d] + in_shape[d] > virt_shape[d]) return fail(who, "bad shape or box");
    g.IX = p.X; g.IY = p.Y; g.IZ = p.Z;
    g.lx = lo[0]; g.ly = lo[1]; g.lz = lo[2];
    g.VX = virt_shape[0]; g.VY = virt_shape[1]; g.VZ = virt_shape[2];
    g.OX = out_shape[0]; g.OY = out_shape[1]; g.OZ = out_shape[2];
    g.sx = s->sx; g.sy = s->sy;
    g.fx = (g.OX > 1) ? (double)(g.VX - 1) / (double)(g.OX - 1) : 0.0;
    g.fy = (g.OY > 1) ? (double)(g.VY - 1) / (double)(g.OY - 1) : 0.0;
    g.fz = (g.OZ > 1) ? (double)(g.VZ - 1) / (double)(g.OZ - 1) : 0.0;
    return 0;
}

int tgtk_sim_download_resampled_async(tgtk_sim *s, int field, const int32_t lo[3], const int32_t virt_shape[3], double *out,
                                      const int32_t out_shape[3])
{
    if (!s || !lo || !virt_shape || !out || !out_shape) return fail("tgtk_sim_download_resampled_async", "null argument");
    if (field < 0 || field >= s->nfields) return fail("tgtk_sim_download_resampled_async", "not a state field of this model");
    ResampleGeom g;
    if (resample_geom(s, "tgtk_sim_download_resampled_async", lo, virt_shape, out_shape, g)) return 1;
    CU(cudaSetDevice(s->device));
    if (!s->ctrl_valid && fetch_ctrl(s)) return 1;   // the loop is over: which buffer holds the state, replay of a tripped stop
    const void *src = s->state[current_buf(s, nullptr)][field];
    const size_t nout = (size_t)g.OX * g.OY * g.OZ;
    if (nout > s->resample_async_cap[field]) {
        if (s->resample_async[field]) CU(sim_free(s, s->resample_async[field]));
        s->resample_async[field] = nullptr;
        s->resample_async_cap[field] = 0;
        CU(sim_malloc(s, (void **)(&s->resample_async[field]), nout * 8));
        s->resample_async_cap[field] = nout;
    }
    if (s->p.dtype == TGTK_F64) resample_state_kernel<double><<<grid_1d(nout), 256, 0, s->stream>>>((const double *)src, g, s->resample_async[field]);
    else resample_state_kernel<float><<<grid_1d(nout), 256, 0, s->stream>>>((const float *)src, g, s->resample_async[field]);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, s->resample_async[field], nout * 8, cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

int tgtk_sim_segmentation_async(tgtk_sim *s, const int32_t lo[3], const int32_t virt_shape[3], const int32_t out_shape[3],
                                double th_necro_n, double th_enhancing_p, double th_edema_p, uint8_t *out)
{
    if (!s || !lo || !virt_shape || !out || !out_shape) return fail("tgtk_sim_segmentation_async", "null argument");
    if (s->p.model != TGTK_MODEL_FK_2C && s->p.model != TGTK_MODEL_FK2C_COEF12)
        return fail("tgtk_sim_segmentation_async", "the segmentation map needs the P and N fields of FK_2c");
    ResampleGeom g;
    if (resample_geom(s, "tgtk_sim_segmentation_async", lo, virt_shape, out_shape, g)) return 1;
    CU(cudaSetDevice(s->device));
    if (!s->ctrl_valid && fetch_ctrl(s)) return 1;
    const int cur = current_buf(s, nullptr);
    const size_t nout = (size_t)g.OX * g.OY * g.OZ;
    if (nout > s->seg_cap) {
        if (s->seg_out) CU(sim_free(s, s->seg_out));
        s->seg_out = nullptr;
        s->seg_cap = 0;
        CU(sim_malloc(s, (void **)(&s->seg_out), nout));
        s->seg_cap = nout;
    }
    if (s->p.dtype == TGTK_F64)
        segmentation_kernel<double><<<grid_1d(nout), 256, 0, s->stream>>>((const double *)s->state[cur][0], (const double *)s->state[cur][1], g,
                                                                        th_necro_n, th_enhancing_p, th_edema_p, s->seg_out);
    else
        segmentation_kernel<float><<<grid_1d(nout), 256, 0, s->stream>>>((const float *)s->state[cur][0], (const float *)s->state[cur][1], g,
                                                                       th_necro_n, th_enhancing_p, th_edema_p, s->seg_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, s->seg_out, nout, cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

// ---- tools.makeXYZ_rgb_from_tensor (FK_DTI/tools.py:166-187) on the device ------------------------------------------
namespace {
// pass 1: diagonal of every 3x3 tensor -> rgb, brain mask (max of the diagonal > 0, tools.py:172), per-block sum and
// count of the entries of brain voxels
__global__ void __launch_bounds__(256) rgb_extract_kernel(const double *__restrict__ t, double *__restrict__ rgb,
                                                          unsigned char *__restrict__ brain, int64_t n, double *__restrict__ psum,
                                                          unsigned long long *__restrict__ pcnt)
{
    __shared__ double rs[8];
    __shared__ unsigned long long rc[8];
    double s = 0.0;
    unsigned long long c = 0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const double a = t[q * 9], b = t[q * 9 + 4], d = t[q * 9 + 8];
        rgb[q * 3] = a; rgb[q * 3 + 1] = b; rgb[q * 3 + 2] = d;
        const bool in = fmax(fmax(a, b), d) > 0.0;
        brain[q] = in;
        if (in) { s += (a + b) + d; c += 3; }
    }
    s = warp_sum(s);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0; unsigned long long b = 0;
        for (int w = 0; w < 8; ++w) { a += rs[w]; b += rc[w]; }
        psum[blockIdx.x] = a; pcnt[blockIdx.x] = b;
    }
}

// one block, fixed order: stat[which] = sum(psum) / count; the count is summed from pcnt (pass 1) or reused (pass 2)
__global__ void __launch_bounds__(256) rgb_mean_kernel(const double *psum, const unsigned long long *pcnt, int nparts, double *stat,
                                                       unsigned long long *count, int which)
{
    __shared__ double rs[256];
    __shared__ unsigned long long rc[256];
    double a = 0.0; unsigned long long b = 0;
    for (int i = threadIdx.x; i < nparts; i += 256) { a += psum[i]; if (pcnt) b += pcnt[i]; }
    rs[threadIdx.x] = a; rc[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0; unsigned long long tb = 0;
        for (int i = 0; i < 256; ++i) { ta += rs[i]; tb += rc[i]; }
        if (pcnt) *count = tb;
        stat[which] = ta / (double)(*count);
    }
}

// x /= mean; x[x > 10] = 10; x[x < 0] = 0 (tools.py:174-178).  stage 0 then applies x**exponent + x*linear (:183) and
// sums the result over the brain voxels for the second normalisation (:185)
__global__ void __launch_bounds__(256) rgb_normalise_kernel(double *__restrict__ rgb, const unsigned char *__restrict__ brain, int64_t n,
                                                            const double *__restrict__ stat, int stage, double exponent, double linear,
                                                            double *__restrict__ psum)
{
    __shared__ double rs[8];
    const double m = stat[stage];
    double s = 0.0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        double v[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double x = rgb[q * 3 + a] / m;
            x = (x > 10.0) ? 10.0 : x;
            x = (x < 0.0) ? 0.0 : x;
            if (stage == 0) x = ((exponent == 1.0) ? x : pow(x, exponent)) + x * linear;
            v[a] = x;
            rgb[q * 3 + a] = x;
        }
        if (stage == 0 && brain[q]) s += (v[0] + v[1]) + v[2];
    }
    if (stage == 0) {
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) { double a = 0.0; for (int w = 0; w < 8; ++w) a += rs[w]; psum[blockIdx.x] = a; }
    }
}
}  // namespace

int tgtk_dti_rgb_host(int device, const double *tensors, int64_t n, double exponent, double linear, double *rgb)
{
    if (!tensors || !rgb) return fail("tgtk_dti_rgb_host", "null argument");
    if (n < 1) return fail("tgtk_dti_rgb_host", "empty tensor volume");
    CU(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    // the tensor volume (72 bytes per voxel, 643 MB for 240x240x155) goes through in slabs
    const int64_t chunk = std::min<int64_t>(n, (int64_t)1 << 21);
    const int nslab = (int)((n + chunk - 1) / chunk), per = grid_1d((size_t)chunk), nb = grid_1d((size_t)n);
    const int nparts = std::max(per * nslab, nb);
    double *dt = nullptr, *drgb = nullptr, *dsum = nullptr, *dstat = nullptr;
    unsigned long long *dcnt = nullptr;
    unsigned char *dbrain = nullptr;
    int rc = 0;
    cudaError_t e = cudaMallocAsync((void **)&drgb, (size_t)n * 24, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dbrain, (size_t)n, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dt, (size_t)chunk * 72, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dsum, sizeof(double) * nparts, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dcnt, sizeof(unsigned long long) * (nparts + 1), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dstat, sizeof(double) * 2, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dsum, 0, sizeof(double) * nparts, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dcnt, 0, sizeof(unsigned long long) * (nparts + 1), st);
    for (int k = 0; k < nslab && e == cudaSuccess; ++k) {
        const int64_t q = (int64_t)k * chunk, m = std::min(chunk, n - q);
        e = cudaMemcpyAsync(dt, tensors + q * 9, (size_t)m * 72, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            rgb_extract_kernel<<<grid_1d((size_t)m), 256, 0, st>>>(dt, drgb + q * 3, dbrain + q, m, dsum + (size_t)k * per, dcnt + (size_t)k * per);
            e = cudaGetLastError();
        }
    }
    unsigned long long *dcount = dcnt + nparts;
    if (e == cudaSuccess) { rgb_mean_kernel<<<1, 256, 0, st>>>(dsum, dcnt, per * nslab, dstat, dcount, 0); e = cudaGetLastError(); }
    if (e == cudaSuccess) { rgb_normalise_kernel<<<nb, 256, 0, st>>>(drgb, dbrain, n, dstat, 0, exponent, linear, dsum); e = cudaGetLastError(); }
    if (e == cudaSuccess) { rgb_mean_kernel<<<1, 256, 0, st>>>(dsum, nullptr, nb, dstat, dcount, 1); e = cudaGetLastError(); }
    if (e == cudaSuccess) { rgb_normalise_kernel<<<nb, 256, 0, st>>>(drgb, dbrain, n, dstat, 1, exponent, linear, dsum); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgb, drgb, (size_t)n * 24, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail("tgtk_dti_rgb_host", cudaGetErrorString(e));
    if (dt) cudaFreeAsync(dt, st);
    if (drgb) cudaFreeAsync(drgb, st);
    if (dbrain) cudaFreeAsync(dbrain, st);
    if (dsum) cudaFreeAsync(dsum, st);
    if (dcnt) cudaFreeAsync(dcnt, st);
    if (dstat) cudaFreeAsync(dstat, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

// ---- batched restore + zoom of snapshots / the final state into pageable host memory ------------------------------
namespace {
struct PinnedRing {            // page-locked staging buffers, kept for the life of the process (page-locking 71 MB costs ~25 ms)
    std::mutex mu;
    void *buf[3] = {nullptr, nullptr, nullptr};
    size_t bytes = 0;
    bool busy = false;
};
PinnedRing g_ring;

// dst[0..n) = src[0..n) with `nthreads` threads: a fresh numpy array is untouched memory, one thread spends more time
// in page faults than in the copy
void parallel_copy(char *dst, const char *src, size_t n, int nthreads, std::vector<std::thread> &pool)
{
    const size_t chunk = ((n / nthreads) + 4095) & ~(size_t)4095;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t len = std::min(chunk, n - off);
        pool.emplace_back([=] { memcpy(dst + off, src + off, len); });
    }
}
}  // namespace

int tgtk_sim_fetch_series(tgtk_sim *s, int field, int first, int count, const int32_t lo[3], const int32_t virt_shape[3],
                          const int32_t out_shape[3], double *out)
{
    if (!s || !lo || !virt_shape || !out_shape || !out) return fail("tgtk_sim_fetch_series", "null argument");
    if (field < 0 || field >= s->nfields) return fail("tgtk_sim_fetch_series", "not a state field of this model");
    if (count < 0 || first < -1 || (first == -1 && count > 1) || (first >= 0 && first + count > (int)s->snaps.size()))
        return fail("tgtk_sim_fetch_series", "no such snapshot");
    if (count == 0) return 0;
    ResampleGeom g;
    if (resample_geom(s, "tgtk_sim_fetch_series", lo, virt_shape, out_shape, g)) return 1;
    CU(cudaSetDevice(s->device));
    if (!s->ctrl_valid && fetch_ctrl(s)) return 1;
    const size_t nout = (size_t)g.OX * g.OY * g.OZ, bytes = nout * 8, state_bytes = (size_t)s->p.X * s->sx * s->elem;
    // two device output buffers, so that the resample of item i+1 does not wait for the copy of item i
    for (int k = 0; k < 2; ++k)
        if (nout > s->resample_async_cap[k] || !s->resample_async[k]) {
            if (s->resample_async[k]) CU(sim_free(s, s->resample_async[k]));
            s->resample_async[k] = nullptr;
            s->resample_async_cap[k] = 0;
            CU(sim_malloc(s, (void **)(&s->resample_async[k]), bytes));
            s->resample_async_cap[k] = nout;
        }
    // the staging ring is process-wide; a concurrent caller (another simulation's thread) falls back to its own buffers
    bool own_ring = false;
    void *ring[3] = {nullptr, nullptr, nullptr};
    {
        std::lock_guard<std::mutex> lk(g_ring.mu);
        if (!g_ring.busy) {
            if (g_ring.bytes < bytes) {
                for (int k = 0; k < 3; ++k) { if (g_ring.buf[k]) cudaFreeHost(g_ring.buf[k]); g_ring.buf[k] = nullptr; }
                g_ring.bytes = 0;
                bool ok = true;
                for (int k = 0; k < 3 && ok; ++k) ok = (cudaMallocHost(&g_ring.buf[k], bytes) == cudaSuccess);
                if (ok) g_ring.bytes = bytes;
            }
            if (g_ring.bytes >= bytes) { g_ring.busy = true; for (int k = 0; k < 3; ++k) ring[k] = g_ring.buf[k]; }
        }
    }
    if (!ring[0]) {
        own_ring = true;
        for (int k = 0; k < 3; ++k)
            if (cudaMallocHost(&ring[k], bytes) != cudaSuccess) {
                for (int j = 0; j < k; ++j) cudaFreeHost(ring[j]);
                return fail("tgtk_sim_fetch_series", "cannot allocate page-locked staging buffers");
            }
    }
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    int rc = 0;
    for (int k = 0; k < 3 && !rc; ++k)
        if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) rc = fail("tgtk_sim_fetch_series", "cudaEventCreate failed");
    const int nthreads = std::max(1, std::min(8, (int)std::thread::hardware_concurrency()));
    std::vector<std::thread> jobs[3];
    auto join_slot = [&](int k) { for (auto &t : jobs[k]) t.join(); jobs[k].clear(); };
    auto enqueue = [&](int i) -> int {         // resample item i on the device, copy it into ring slot i % 3
        const int k = i % 3;
        join_slot(k);                            // the copy-out that last read this slot
        const void *src = (first < 0) ? s->state[current_buf(s, nullptr)][field] : (const void *)((char *)s->snaps[first + i] + state_bytes * field);
        double *dev = s->resample_async[i & 1];
        if (s->p.dtype == TGTK_F64) resample_state_kernel<double><<<grid_1d(nout), 256, 0, s->stream>>>((const double *)src, g, dev);
        else resample_state_kernel<float><<<grid_1d(nout), 256, 0, s->stream>>>((const float *)src, g, dev);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(ring[k], dev, bytes, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaEventRecord(ev[k], s->stream));
        return 0;
    };
    if (!rc) rc = enqueue(0);
    for (int i = 0; i < count && !rc; ++i) {
        if (i + 1 < count) rc = enqueue(i + 1);
        if (rc) break;
        const int k = i % 3;
        if (cudaEventSynchronize(ev[k]) != cudaSuccess) { rc = fail("tgtk_sim_fetch_series", "device-to-host copy failed"); break; }
        parallel_copy((char *)(out + (size_t)i * nout), (const char *)ring[k], bytes, nthreads, jobs[k]);
    }
    for (int k = 0; k < 3; ++k) join_slot(k);
    cudaStreamSynchronize(s->stream);
    for (int k = 0; k < 3; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
    if (own_ring) { for (int k = 0; k < 3; ++k) cudaFreeHost(ring[k]); }
    else { std::lock_guard<std::mutex> lk(g_ring.mu); g_ring.busy = false; }
    return rc;
}

void *tgtk_host_alloc(size_t bytes)
{
    void *q = nullptr;
    if (cudaMallocHost(&q, bytes ? bytes : 8) != cudaSuccess) {
        fail("tgtk_host_alloc", "cudaMallocHost failed");
        return nullptr;
    }
    return q;
}

void tgtk_host_free(void *q)
{
    if (q) cudaFreeHost(q);
}

int tgtk_resample_linear_host(int device, const double *in, const int64_t in_strides[3], const int32_t in_shape[3],
                              const int32_t lo[3], const int32_t virt_shape[3], double *out, const int32_t out_shape[3])
{
    if (!in || !in_strides || !in_shape || !lo || !virt_shape || !out || !out_shape)
        return fail("tgtk_resample_linear_host", "null argument");
    for (int d = 0; d < 3; ++d)
        if (in_shape[d] < 1 || virt_shape[d] < 1 || out_shape[d] < 1 || lo[d] < 0 || lo[d] + in_shape[d] > virt_shape[d])
            return fail("tgtk_resample_linear_host", "bad shape or box");
    CU(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t nin = (size_t)in_shape[0] * in_shape[1] * in_shape[2];
    const size_t nout = (size_t)out_shape[0] * out_shape[1] * out_shape[2];
    double *din = nullptr, *dout = nullptr;
    int rc = 0;
    cudaError_t e = cudaMallocAsync((void **)&din, nin * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dout, nout * 8, st);
    if (e != cudaSuccess) rc = fail("cudaMalloc", cudaGetErrorString(e));
    if (!rc) rc = copy_host_dev(st, in_shape[0], in_shape[1], in_shape[2], din, const_cast<double *>(in), in_strides, true);
    if (!rc) {
        resample_linear_kernel<<<grid_1d(nout), 256, 0, st>>>(din, in_shape[0], in_shape[1], in_shape[2], lo[0], lo[1], lo[2],
                                                            virt_shape[0], virt_shape[1], virt_shape[2], dout, out_shape[0],
                                                            out_shape[1], out_shape[2]);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, nout * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail("resample", cudaGetErrorString(e));
    }
    if (din) cudaFreeAsync(din, st);
    if (dout) cudaFreeAsync(dout, st);
    cudaStreamDestroy(st);
    return rc;
}

int tgtk_elongate_tensors_host(int device, const double *tensors, int64_t n, double scale_factor, float *out)
{
    if (!tensors || !out) return fail("tgtk_elongate_tensors_host", "null argument");
    if (n < 0) return fail("tgtk_elongate_tensors_host", "negative tensor count");
    if (n == 0) return 0;
    CU(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    // chunks bound the device footprint (108 bytes per tensor) whatever the volume: 240x240x155 is 8.9 M tensors
    const int64_t chunk = std::min<int64_t>(n, (int64_t)1 << 21);
    double *din = nullptr;
    float *dout = nullptr;
    int rc = 0;
    cudaError_t e = cudaMallocAsync((void **)&din, (size_t)chunk * 72, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dout, (size_t)chunk * 36, st);
    if (e != cudaSuccess) rc = fail("cudaMalloc", cudaGetErrorString(e));
    for (int64_t q = 0; q < n && !rc; q += chunk) {
        const int64_t m = std::min(chunk, n - q);
        e = cudaMemcpyAsync(din, tensors + q * 9, (size_t)m * 72, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            tgtk::elongate_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(din, dout, m, scale_factor);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + q * 9, dout, (size_t)m * 36, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail("elongate", cudaGetErrorString(e));
    }
    if (din) cudaFreeAsync(din, st);
    if (dout) cudaFreeAsync(dout, st);
    cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
#include "tgtk_nccl.inc"
