// Shared device-side definitions of libtgtk_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tgtk {

// Sentinel for "no flux through any face of this voxel": face = max(k(c) + k(c'), 0), so one
// closed end drives the sum far below zero and the face coefficient to exactly 0.
template <typename T> struct Big;
template <> struct Big<double> { static constexpr double v = 1.0e300; };
template <> struct Big<float> { static constexpr float v = 1.0e30f; };

// Device-resident loop control: lets identical launches be enqueued back to back (and
// captured in CUDA graphs) while the stop test of the reference loops
// (FK/FK.py:216-218, FK_DTI/FK_DTI.py:199-206) is evaluated on the device after every step.
struct Ctrl {
    int t;            // index of the next step to execute
    int stop_reason;  // tgtk_stop; non-zero freezes the state (later launches return at once)
    int stop_step;    // step index that tripped the test, -1 if none
    unsigned ticket;  // last-block-done counter of the running launch
    double last_volume;
    int replay;       // speculative mode: steps to redo from the checkpoint (state before step t) to reach the stop step
};

// Everything a step kernel needs; passed by value.
template <typename T> struct StepArgs {
    int X, Y, Z;
    int64_t sx, sy;        // element pitches of axis 0 and 1 (axis 2 pitch is 1)
    T *buf[2][3];          // ping-pong state: step t reads buf[t&1], writes buf[(t+1)&1]
    const T *coef[12];     // model dependent, see tgtk_api.cu
    T cx, cy, cz;          // FK/FK_2c: Dw/h^2 ; DTI/COEF6/FACES: 1/h^2
    T ex, ey, ez;          // FK_2c: Ds/h^2
    T hx, hy, hz, gx, gy, gz;   // the same six constants for the two-voxel kernels; with -DTGTK_RELU2 (experiment, off) they carry a
                                // factor 1/2 and max(x, 0) is taken as x + |x| = 2 max(x, 0): exact, every product bit for bit the same
    T dt, rho;
    T th_necro, lambda_np, sigma_np, lambda_s;
    double vol_scale_p;    // dx*dy*dz
    double vol_scale_n;    // FK_2c: 1.0 (sic, FK_2c.py:235); others unused
    double stop_volume;
    double small_volume;
    Ctrl *ctrl;
    double *partials;      // [2][nblocks] per-block sums (field 0, field 1); host-scheduled: a ring of such slots
    unsigned ring_stride;  // doubles per ring slot (fits the largest grid of the simulation's kernels)
    double *volumes;       // per-step volume log
    int volumes_cap;
    // Scheduling.  host_sched = 0: the step index and the stop flag live on the device (Ctrl)
    // and are read at kernel start; the last block to finish reduces the volume and runs the
    // reference's stop tests.  host_sched = 1 (no stop criterion active): the ping-pong parity
    // is baked into the launch, blocks only deposit their partial sums in ring slot `slot`, and
    // chunk_finish_kernel reduces a whole chunk of steps later -- nothing on the step-to-step
    // critical path depends on a device-side read-modify-write.
    int host_sched, par, slot;
    // host_sched launches of a run WITH a stop criterion (speculative mode, tgtk_api.cu): chunk_finish_kernel
    // runs the stop tests for a whole chunk afterwards; once one has tripped every later launch is void
    int spec;
    // planes [sum_x0, sum_x1) count towards the volume (slab mode: the owned planes, not the ghosts)
    int sum_x0, sum_x1;
    // marching kernels: planes ahead of the current one whose lines are requested into L2
    // (prefetch.global.L2, fire and forget); 0 = off
    int l2pf;
    // planes [x_lo, x_hi) of axis 0 are computed by this launch (two-voxel kernels; default: the whole box)
    int x_lo, x_hi;
    int zigzag;               // two-voxel kernels: odd steps march axis 0 downwards (L2 reuse of what the previous step wrote last)
    int skip;                 // two-voxel FK / FK_2c kernels: exact per-warp shortcuts for warps that hold no tumour / lie outside
                              // the tissue (TGTK_CLOSED_SKIP=0 switches them off, e.g. to measure the dense path)
    // Fused slab mode (tgtk_nccl.inc, one rank per GPU, axis 0 decomposed, one ghost plane per side): the blocks
    // that compute the rank's lowest / highest owned plane ALSO store it into the neighbouring rank's ghost plane
    // (peer-mapped memory, NVLink).  A launch's stores are visible system-wide when the launch has completed, so
    // the NEXT launch of the stream (step, checkpoint or chunk commit) publishes "epoch e-1 is complete" in both
    // neighbours' flag words as its first action, and only the blocks that read a ghost plane -- scheduled last --
    // wait for the neighbours' word.  No exchange launches, no copies, no fences inside the step; interior blocks
    // never wait.
    int slab;                 // 0: off
    T *peer_lo[2][3];         // previous rank's state buffers [ping-pong][field]
    T *peer_hi[2][3];         // next rank's
    uint32_t off_lo, off_hi;  // added (mod 2^32) to an element offset in my plane x_lo / x_hi-1: the neighbour's ghost element
    int *flags;               // my words (SlabWord): written by the neighbours
    int *flag_lo, *flag_hi;   // the word this rank publishes in the previous / next rank's memory
};

// words of the slab flag array (one cudaMalloc per rank, mapped by both neighbours)
enum SlabWord { kSlabFromPrev = 8, kSlabFromNext = 9, kSlabEpoch = 10, kSlabError = 4 };

__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one thread: wait until the words `w0`, `w1` (-1: skip) of `flags` have reached `epoch`; gives up after 20 s
__device__ __forceinline__ void slab_spin(int *flags, int w0, int w1, int epoch)
{
    unsigned long long t0 = 0, t1;
    for (unsigned it = 0;; ++it) {
        const bool ok0 = w0 < 0 || ld_acquire_sys(flags + w0) >= epoch;
        const bool ok1 = w1 < 0 || ld_acquire_sys(flags + w1) >= epoch;
        if (ok0 && ok1) break;
        if ((it & 63u) == 0u) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 20000000000ull) { ((volatile int *)flags)[kSlabError] = 3; break; }   // report instead of hanging
        }
    }
}

// first action of every launch of a decomposed run: whatever the stream ran before this launch has completed, its
// stores into the neighbours' ghost planes included -- say so
__device__ __forceinline__ void slab_publish(int *flag_lo, int *flag_hi, int done_epoch)
{
    st_release_sys(flag_lo, done_epoch);
    st_release_sys(flag_hi, done_epoch);
}

struct SlabCtl {
    bool lo, hi;   // this block computes the rank's lowest / highest owned plane (reads that ghost, pushes that plane)
    int epoch;     // of this launch
};

// after griddep_wait(), before the first state access of the block
template <typename T> __device__ __forceinline__ SlabCtl slab_begin(const StepArgs<T> &a, int x0, int x1)
{
    SlabCtl c;
    c.lo = c.hi = false;
    c.epoch = 0;
    if (!a.slab) return c;
    c.lo = (x0 == a.x_lo);
    c.hi = (x1 == a.x_hi);
    c.epoch = a.flags[kSlabEpoch] + a.slot + 1;      // only chunk_commit_kernel writes the base, never beside a step launch
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0)
        slab_publish(a.flag_lo, a.flag_hi, c.epoch - 1);
    if (c.lo || c.hi) {
        if (threadIdx.x == 0 && threadIdx.y == 0) slab_spin(a.flags, c.lo ? kSlabFromPrev : -1, c.hi ? kSlabFromNext : -1, c.epoch - 1);
        __syncthreads();
    }
    return c;
}

// L2 prefetch of the rows [y0, y0+rows) of plane xp of one array: rows of a plane are contiguous
// (axis 2 fastest), so the whole strip -- every z tile of it -- is ONE bulk request to the TMA unit
// (cp.async.bulk.prefetch.L2), issued by a single thread of the blockIdx.x == 0 tile.  Start and
// end are rounded to the instruction's 16-byte granularity inside the array.
template <typename T>
__device__ __forceinline__ void l2_prefetch_strip(const T *base, uint32_t first, uint32_t count, uint32_t total)
{
    const uint64_t b0 = ((uint64_t)first * sizeof(T)) & ~(uint64_t)15;
    uint64_t b1 = ((uint64_t)(first + count) * sizeof(T) + 15) & ~(uint64_t)15;
    const uint64_t cap = ((uint64_t)total * sizeof(T)) & ~(uint64_t)15;
    if (b1 > cap) b1 = cap;
    if (b1 <= b0) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char *)base + b0), "r"((uint32_t)(b1 - b0))
                 : "memory");
}

// per-kernel prefetch state of the thread that issues the strips (others: on = false)
struct L2Strip {
    bool on;
    uint32_t row0, count, plane, total;   // elements: offset of the strip in a plane, its length, plane pitch, array size
};
template <typename T> __device__ __forceinline__ L2Strip l2_strip_setup(const StepArgs<T> &a, int rows_per_tile)
{
    L2Strip s;
    const int tid = threadIdx.x + blockDim.x * threadIdx.y;
    s.on = a.l2pf > 0 && blockIdx.x == 0 && tid == 0;
    const int y0 = blockIdx.y * rows_per_tile;
    const int rows = min(rows_per_tile, a.Y - y0);
    s.row0 = (uint32_t)y0 * (uint32_t)a.sy;
    s.count = (uint32_t)rows * (uint32_t)a.sy;
    s.plane = (uint32_t)a.sx;
    s.total = (uint32_t)a.X * (uint32_t)a.sx;
    return s;
}
// element offset of the strip in plane x + l2pf (periodic); dir = -1: the block marches downwards, plane x - l2pf
template <typename T> __device__ __forceinline__ uint32_t l2_strip_at(const StepArgs<T> &a, const L2Strip &s, int x, int dir = 1)
{
    int xp = x + dir * a.l2pf;
    if (xp >= a.X) xp -= a.X;
    if (xp < 0) xp += a.X;
    return (uint32_t)xp * s.plane + s.row0;
}

// Programmatic dependent launch (tgtk_api.cu: launch_v4).  Both are no-ops in a launch without the
// attribute.  griddep_wait() returns once the preceding grid has completed and its writes are visible:
// nothing that grid writes (the state it produces) or reads (the state this grid overwrites) may be
// touched before it.
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename T> __device__ __forceinline__ bool counts(const StepArgs<T> &a, int x)
{
    return x >= a.sum_x0 && x < a.sum_x1;
}

struct StepCtl {
    int t;       // step index (device-scheduled mode), -1 otherwise
    int par;     // index of the buffer holding the current state
    bool stop;
};

template <typename T> __device__ __forceinline__ StepCtl step_begin(const StepArgs<T> &a)
{
    StepCtl c;
    if (a.host_sched) {
        // the flag is only ever written by a chunk_finish_kernel, which no step launch overlaps
        c.t = -1; c.par = a.par; c.stop = a.spec && (*(volatile const int *)&a.ctrl->stop_reason != 0);
    } else {
        c.t = a.ctrl->t;
        c.par = c.t & 1;
        c.stop = (a.ctrl->stop_reason != 0);
    }
    return c;
}

// Arithmetic in the reference's rounding order: no FMA contraction (numpy evaluates every
// ufunc separately), used where bit-exact agreement with the float64 reference is offered.
template <typename T> struct Rn;
template <> struct Rn<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
};
template <> struct Rn<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};

__device__ __forceinline__ double tg_exp(double x) { return exp(x); }
__device__ __forceinline__ float tg_exp(float x) { return expf(x); }
__device__ __forceinline__ double tg_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float tg_max(float a, float b) { return fmaxf(a, b); }

// max(x, 0) for finite x.  fp64 fmax() compiles to 7 SASS instructions (DSETP.MAX plus NaN
// fix-up) and sits on the fp64 pipe; masking with the sign bit is 3 integer ops.
__device__ __forceinline__ double relu(double x)
{
#ifdef TGTK_RELU_ABS
    return 0.5 * (x + fabs(x));     // experiment: |x| is an operand modifier, two fp64-pipe instructions instead of three ALU ones
#endif
    const int hi = __double2hiint(x);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(x) & keep);
}
__device__ __forceinline__ float relu(float x) { return fmaxf(x, 0.0f); }

// 2 max(x, 0) for the two-voxel kernels (see StepArgs::hx): x + |x|.  Compile-time experiment (-DTGTK_RELU2), OFF by
// default: it removes 18 ALU instructions per FK_2c voxel but adds 9 to the fp64 pipe, and measured 2.4 % SLOWER on the
// B200 (profiles/r2_sweep_misc.txt) -- the kernel is not short of issue slots, it is short of fp64-pipe latency cover.
#ifdef TGTK_RELU2
template <typename T> struct Relu2 { static constexpr double scale = 0.5; };
__device__ __forceinline__ double relu2(double x) { return x + fabs(x); }
__device__ __forceinline__ float relu2(float x) { return x + fabsf(x); }
#else
template <typename T> struct Relu2 { static constexpr double scale = 1.0; };
__device__ __forceinline__ double relu2(double x) { return relu(x); }
__device__ __forceinline__ float relu2(float x) { return relu(x); }
#endif

// smooth_heaviside, FK_2c/FK_2c.py:75-76 (k = 50)
template <typename T> __device__ __forceinline__ T heaviside50(T s, T sigma)
{
    const T e = tg_exp(T(-50) * (s - sigma));
    return T(1) - T(1) / (T(1) + e);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of up to two doubles per thread, then the deterministic last-block-done
// final reduction, the volume / stop test of the reference loop and the step counter bump.
// Every thread of every block of the launch must call this exactly once.
template <typename T, int kFields>
__device__ __forceinline__ void finish_step(const StepArgs<T> &a, int t, double s0, double s1)
{
    __shared__ double red[2][32];
    __shared__ bool is_last;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int nwarps = (nthreads + 31) >> 5;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);

    s0 = warp_sum(s0);
    if (kFields > 1) s1 = warp_sum(s1);
    if (lane == 0) {
        red[0][warp] = s0;
        red[1][warp] = s1;
    }
    __syncthreads();
    if (warp == 0) {
        double v0 = (lane < nwarps) ? red[0][lane] : 0.0;
        double v1 = (kFields > 1 && lane < nwarps) ? red[1][lane] : 0.0;
        v0 = warp_sum(v0);
        if (kFields > 1) v1 = warp_sum(v1);
        if (lane == 0) {
            a.partials[bid] = v0;
            if (kFields > 1) a.partials[nblocks + bid] = v1;
            __threadfence();
            const unsigned done = atomicAdd(&a.ctrl->ticket, 1u);
            is_last = (done == nblocks - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;

    // last block: fixed-order reduction of the per-block partials
    __threadfence();
    double v0 = 0.0, v1 = 0.0;
    for (unsigned i = tid; i < nblocks; i += nthreads) {
        v0 += __ldcg(&a.partials[i]);
        if (kFields > 1) v1 += __ldcg(&a.partials[nblocks + i]);
    }
    v0 = warp_sum(v0);
    if (kFields > 1) v1 = warp_sum(v1);
    __syncthreads();
    if (lane == 0) {
        red[0][warp] = v0;
        red[1][warp] = v1;
    }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < nwarps; ++w) {
            t0 += red[0][w];
            t1 += red[1][w];
        }
        // FK/FK.py:215  volume = dx*dy*dz*sum(A);  FK_2c.py:235  dx*dy*dz*sum(P) + sum(N)
        double vol = a.vol_scale_p * t0;
        if (kFields > 1) vol += a.vol_scale_n * t1;
        if (t < a.volumes_cap) a.volumes[t] = vol;
        a.ctrl->last_volume = vol;
        if (vol >= a.stop_volume) {
            a.ctrl->stop_reason = 1;
            a.ctrl->stop_step = t;
        } else if (vol < a.small_volume) {
            a.ctrl->stop_reason = 2;
            a.ctrl->stop_step = t;
        }
        a.ctrl->t = t + 1;
        a.ctrl->ticket = 0u;
        __threadfence();
    }
}

// host-scheduled mode: per-block partial sums into ring slot a.slot, nothing else
template <typename T, int kFields>
__device__ __forceinline__ void finish_light(const StepArgs<T> &a, double s0, double s1)
{
    __shared__ double red2[2][32];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int nwarps = (nthreads + 31) >> 5;
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    s0 = warp_sum(s0);
    if (kFields > 1) s1 = warp_sum(s1);
    if (lane == 0) {
        red2[0][warp] = s0;
        red2[1][warp] = s1;
    }
    __syncthreads();
    if (warp == 0) {
        double v0 = (lane < nwarps) ? red2[0][lane] : 0.0;
        double v1 = (kFields > 1 && lane < nwarps) ? red2[1][lane] : 0.0;
        v0 = warp_sum(v0);
        if (kFields > 1) v1 = warp_sum(v1);
        if (lane == 0) {
            double *ring = a.partials + (size_t)a.slot * a.ring_stride;
            ring[bid] = v0;
            if (kFields > 1) ring[nblocks + bid] = v1;
        }
    }
}

template <typename T, int kFields>
__device__ __forceinline__ void step_end(const StepArgs<T> &a, const StepCtl &c, double s0, double s1)
{
    if (a.host_sched) finish_light<T, kFields>(a, s0, s1);
    else finish_step<T, kFields>(a, c.t, s0, s1);
}

// Reduces the partial sums of `count` consecutive host-scheduled steps in a fixed order, logs the
// volumes and advances the device step counter.  One block per step (ring slot = blockIdx.x); the
// last block to finish (ticket) publishes the new step index, after every block has read the old.
// spec != 0 (a run with a stop criterion, executed speculatively): the last block also runs the
// reference's stop tests (FK.py:216-218, FK_DTI.py:199-206) over the chunk in step order.  On a trip
// the step counter stays at the chunk's first step -- the state the checkpoint holds -- and `replay`
// tells the host how many steps lead from there to the stopping step; later launches see the flag.
__global__ void __launch_bounds__(256) chunk_finish_kernel(Ctrl *ctrl, const double *partials, double *volumes,
                                                           int volumes_cap, unsigned nblocks, int nfields, int count,
                                                           double scale_p, double scale_n, int spec, double stop_volume,
                                                           double small_volume, unsigned ring_stride, unsigned nblocks_tb,
                                                           unsigned tb_mask, double *chunkvol)
{
    // chunkvol != nullptr (fused slab mode): only deposit the chunk's LOCAL volumes there; chunk_commit_kernel logs
    // them, runs the stop tests and advances the counters once they have been summed over the ranks
    __shared__ double red[2][8];
    if (spec && *(volatile const int *)&ctrl->stop_reason != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x;
    const int t0 = ctrl->t;
    const double *ring = partials + (size_t)j * ring_stride;
    if ((tb_mask >> j) & 1u) nblocks = nblocks_tb;       // slot written by a two-step launch (its own grid)
    double v0 = 0.0, v1 = 0.0;
    for (unsigned i = tid; i < nblocks; i += 256) {
        v0 += ring[i];
        if (nfields > 1) v1 += ring[nblocks + i];
    }
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    if (lane == 0) { red[0][warp] = v0; red[1][warp] = v1; }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int w = 0; w < 8; ++w) { a0 += red[0][w]; a1 += red[1][w]; }
        double vol = scale_p * a0;
        if (nfields > 1) vol += scale_n * a1;
        if (chunkvol) { chunkvol[j] = vol; return; }
        if (t0 + j < volumes_cap) volumes[t0 + j] = vol;
        if (!spec && j == count - 1) ctrl->last_volume = vol;
        __threadfence();
        if (atomicAdd(&ctrl->ticket, 1u) == (unsigned)count - 1) {
            ctrl->ticket = 0u;
            if (!spec) {
                ctrl->t = t0 + count;
                ctrl->replay = 0;
            } else {
                __threadfence();
                int trip = -1, reason = 0;
                double v = 0.0;
                for (int i = 0; i < count && t0 + i < volumes_cap; ++i) {
                    v = __ldcg(&volumes[t0 + i]);
                    if (v >= stop_volume) { trip = i; reason = 1; break; }
                    if (v < small_volume) { trip = i; reason = 2; break; }
                }
                ctrl->last_volume = v;
                if (trip >= 0) {
                    ctrl->stop_step = t0 + trip;
                    ctrl->replay = trip + 1;
                    __threadfence();
                    ctrl->stop_reason = reason;
                } else {
                    ctrl->t = t0 + count;
                }
            }
        }
    }
}

// fused slab mode: the second half of chunk_finish_kernel, after the chunk's volumes have been all-reduced
__global__ void chunk_commit_kernel(Ctrl *ctrl, const double *chunkvol, double *volumes, int volumes_cap, int count, int spec,
                                    double stop_volume, double small_volume, int *flags, int *flag_lo, int *flag_hi)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (spec && ctrl->stop_reason != 0) return;       // a voided chunk: its launches did nothing, the epoch stays
    slab_publish(flag_lo, flag_hi, flags[kSlabEpoch] + count);      // the chunk's last step launch has completed
    const int t0 = ctrl->t;
    int trip = -1, reason = 0;
    double v = 0.0;
    for (int i = 0; i < count; ++i) {
        v = chunkvol[i];
        if (t0 + i < volumes_cap) volumes[t0 + i] = v;
        if (spec && trip < 0) {
            if (v >= stop_volume) { trip = i; reason = 1; }
            else if (v < small_volume) { trip = i; reason = 2; }
            if (trip >= 0) ctrl->last_volume = v;
        }
    }
    flags[kSlabEpoch] += count;
    if (trip >= 0) {
        ctrl->stop_step = t0 + trip;
        ctrl->replay = trip + 1;
        __threadfence();
        ctrl->stop_reason = reason;
    } else {
        ctrl->last_volume = v;
        ctrl->t = t0 + count;
        ctrl->replay = 0;
    }
}

// speculative mode: the state before a chunk, kept until the chunk's stop tests have passed
// (flags != nullptr, fused slab mode: the ghost planes are part of that state -- wait until both neighbours have
// said that every launch up to the current epoch, and with it every store into this rank's ghost planes, is complete;
// chunk_commit_kernel, which precedes this launch in the stream, said the same for this rank)
__global__ void __launch_bounds__(256) checkpoint_kernel(const Ctrl *ctrl, const uint2 *s0, const uint2 *s1, const uint2 *s2,
                                                         uint2 *d0, uint2 *d1, uint2 *d2, size_t n8, int *flags)
{
    if (*(volatile const int *)&ctrl->stop_reason != 0) return;
    if (flags) {
        if (threadIdx.x == 0) slab_spin(flags, kSlabFromPrev, kSlabFromNext, flags[kSlabEpoch]);
        __syncthreads();
    }
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n8; q += (size_t)gridDim.x * blockDim.x) {
        d0[q] = s0[q];
        if (s1) d1[q] = s1[q];
        if (s2) d2[q] = s2[q];
    }
}

}  // namespace tgtk
