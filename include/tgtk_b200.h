/*
 * tgtk_b200.h -- C ABI of the B200-native time-stepping library (libtgtk_b200.so).
 *
 * The reference (m1balcerak/TumorGrowthToolkit) is pure Python and has NO FFI / plugin
 * layer: its boundary for this path is the Python class API.  This header
 * is the C ABI introduced BENEATH that API; every entry point names the reference code it
 * replaces.  The Python mirror (tumorgrowthtoolkit_b200/) binds it with ctypes; the stub a
 * reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, doubles; no torch / CUDA types in signatures
 *     (streams and device pointers travel as void* / unsigned long long).
 *   - every call returns 0 on success, non-zero on failure; tgtk_last_error() returns the
 *     message of the calling thread's last failure.  No exceptions cross the boundary.
 *   - host arrays are float64 (the reference's dtype) described by ELEMENT strides, so the
 *     reference's non-contiguous crop views (FK/FK.py:69-71) can be passed as they are.
 *     The library owns all device memory of a simulation; the caller owns host memory.
 *   - a simulation is bound to one device and one stream; calls on one simulation must
 *     come from one thread at a time.  No global mutable state besides the error string.
 *   - boundary conditions are periodic in all three axes (np.roll, FK/FK.py:36-38).
 */
#ifndef TGTK_B200_H
#define TGTK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGTK_ABI_VERSION 1

/* which reference operator a simulation steps */
enum tgtk_model {
    TGTK_MODEL_FK = 0,     /* FK.Solver: get_D + FK_update            FK/FK.py:10-42,212-223        */
    TGTK_MODEL_FK_2C = 1,  /* FK_2c.Solver: get_D_with_necro + FK_2c_update  FK_2c/FK_2c.py:10-76,228-243 */
    TGTK_MODEL_DTI = 2,    /* FK_DTI_Solver: get_D_from_DTI + FK_update  FK_DTI/FK_DTI.py:25-47,195-211 */
    TGTK_MODEL_COEF6 = 3,  /* FK_update with six caller-given coefficient arrays   FK/FK.py:34-42 */
    TGTK_MODEL_FK_FACES = 4, /* FK with the three face arrays built in the reference's exact
                               rounding order (bit-exact fp64 mode)    FK/FK.py:10-32 */
    TGTK_MODEL_FK2C_COEF12 = 5, /* FK_2c_update with twelve caller-given coefficient arrays
                                  FK_2c/FK_2c.py:48-73 */
    TGTK_MODEL_DTI_FULL = 6 /* opt-in: FK_DTI with all six tensor components (19-point stencil incl.
                               the mixed derivatives).  NOT in the packaged reference, which keeps
                               the diagonal only (FK_DTI/tools.py:168-171); parity unpinned.
                               Fields 20..22 = D_xx,D_yy,D_zz, 23..25 = D_xy,D_xz,D_yz (times Dw). */
};

enum tgtk_dtype { TGTK_F64 = 0, TGTK_F32 = 1 };

/* arrays a caller can upload / download (tgtk_sim_upload / tgtk_sim_download) */
enum tgtk_field {
    TGTK_FIELD_U = 0, /* FK / DTI / COEF6 tumour density; FK_2c: P */
    TGTK_FIELD_P = 0,
    TGTK_FIELD_N = 1,
    TGTK_FIELD_S = 2,
    TGTK_FIELD_WM = 10,
    TGTK_FIELD_GM = 11,
    TGTK_FIELD_PC = 12, /* tgtk_faces_host only: P and N of the necrotic closure */
    TGTK_FIELD_NC = 13,
    TGTK_FIELD_COEF0 = 20 /* DTI: 20..22 = D_x,D_y,D_z (= rgb[...,a]*Dw);
                             COEF6: 20..25 = D_plus_x,y,z, D_minus_x,y,z;
                             FK2C_COEF12: 20..25 as COEF6 for P, 26..31 the same for S */
};

enum tgtk_stop {
    TGTK_STOP_NONE = 0,
    TGTK_STOP_VOLUME = 1, /* volume >= stopping_volume                  FK/FK.py:216-218 */
    TGTK_STOP_SMALL = 2   /* FK_DTI only: volume < small_volume (1e-6)  FK_DTI/FK_DTI.py:203-206 */
};

typedef struct tgtk_params {
    int32_t model;           /* enum tgtk_model */
    int32_t dtype;           /* enum tgtk_dtype: storage + arithmetic type of the state */
    int32_t X, Y, Z;         /* cropped box, C order, axis 2 fastest */
    int32_t logical_or_faces; /* 1: tissue maps are booleans and '+' is logical OR (numpy bool
                                 quirk); FK_FACES model only */
    double h[3];             /* dx, dy, dz (already divided by resolution_factor) */
    double dt;               /* stopping_time / Nt                      FK/FK.py:186 */
    double rho;              /* proliferation rate 'f'                  FK/FK.py:122 */
    double Dw;               /* FK, FK_2c (DTI: folded into COEF arrays by the caller) */
    double ratio;            /* RatioDw_Dg                              FK/FK.py:130 */
    double th_matter;        /* FK/FK.py:131 */
    double th_necro;         /* FK_2c/FK_2c.py:154 */
    double Ds;               /* D_s        FK_2c/FK_2c.py:143 */
    double lambda_np;        /* FK_2c/FK_2c.py:141 */
    double sigma_np;         /* FK_2c/FK_2c.py:142 */
    double lambda_s;         /* FK_2c/FK_2c.py:144 */
    double stopping_volume;  /* +inf: never                             FK/FK.py:119 */
    double small_volume;     /* 0: off; FK_DTI passes 1e-6              FK_DTI/FK_DTI.py:203 */
} tgtk_params;

typedef struct tgtk_status {
    int32_t steps_run;   /* number of update steps executed (= index of the stop step + 1) */
    int32_t stop_reason; /* enum tgtk_stop */
    int32_t stop_step;   /* t of the step that tripped the stop test, -1 if none */
    int32_t snapshots;   /* number of valid snapshots recorded so far */
    double last_volume;  /* FK/DTI: dx*dy*dz*sum(u); FK_2c: dx*dy*dz*sum(P)+sum(N) (sic, FK_2c.py:235) */
    double loop_ms;      /* device time of the launches issued by the last tgtk_sim_run (CUDA events) */
    int64_t kernel_launches; /* step-kernel launches issued so far */
} tgtk_status;

typedef struct tgtk_sim tgtk_sim;

int tgtk_abi_version(void);
const char *tgtk_last_error(void);
int tgtk_device_count(int *count);

/* Create a simulation on `device`.  `stream` is a cudaStream_t passed as void* (NULL: the
 * library creates its own non-blocking stream).  Replaces the array set-up part of
 * Solver.solve() after cropping (FK/FK.py:197-208). */
int tgtk_sim_create(tgtk_sim **out, const tgtk_params *params, int device, void *stream);
/* flags: TGTK_CREATE_IPC allocates the state buffers so that they can be shared with the
 * neighbouring ranks' processes (tgtk_sim_ipc_export / tgtk_sim_ipc_connect). */
#define TGTK_CREATE_IPC 1
int tgtk_sim_create_ex(tgtk_sim **out, const tgtk_params *params, int device, void *stream, int flags);
void tgtk_sim_destroy(tgtk_sim *sim);

/* Kernel family: 0 (default) = auto (the fastest measured for the model / dtype), 4 = marching kernels
 * staged through shared memory with two voxels (rows y, y+1) per thread, 5 / 6 = fp32 only: z pairs with
 * 64-bit accesses and packed f32x2 arithmetic, four / two voxels per thread, 3 = staged, one voxel per
 * thread, 2 = marching kernels reading neighbours through L1, 1 = baseline one-voxel-per-thread kernels
 * kept as the in-device cross-check.  The environment variable TGTK_VARIANT sets the default.
 * All families compute the same operator. */
int tgtk_sim_set_variant(tgtk_sim *sim, int variant);
/* What the launch heuristics chose, for logs and benchmarks: out[0] kernel family in effect, out[1] planes
 * per block segment, out[2] L2 prefetch distance in planes (0 = off), out[3] programmatic dependent launch
 * (0/1), out[4..6] grid x, y, z, out[7] threads per block, out[8] temporal blocking: time steps per
 * launch (1 or 2), out[9] work items of a sparse launch (0: dense grid; see tgtk_sim_sparse_info). */
int tgtk_sim_get_tuning(tgtk_sim *sim, int32_t out[10]);
/* Sparse launches (FK_2c; TGTK_SPARSE = 1 on / 0 off / unset: boxes beyond L2 with at least 10 % dead space).  The solvers
 * run on the bounding box of the brain, most of which is background (FK_2c.py:170-185).  A voxel outside the tissue with
 * P = N = 0 is a fixed point of the reference's update that no neighbour can disturb (all its face coefficients are 0), so
 * the 2 x 32 voxel tiles that consist of such voxels only are neither loaded, computed nor stored, and the launch grid is a
 * work list over the rest: same results bit for bit, proportionally less HBM traffic.  Decided from the current state the
 * next time tgtk_sim_run is called after an upload (or by this call).  out[0] work items per launch (0: dense launches),
 * out[1] voxels in live tiles, out[2] voxels of the box, out[3] resident blocks per SM of the sparse kernel. */
int tgtk_sim_sparse_info(tgtk_sim *sim, int64_t out[4]);
/* Diagnostic (tools/sparse_trace.py): runs one more step and reports, per work item i, out[4*i + 0..2] = nanoseconds from the
 * first block's start to this block's start / to the end of its prologue / to the end of its plane loop, out[4*i + 3] = the SM
 * it ran on; planes[2*i + 0..1] = its plane range (may be NULL).  cap: items the arrays hold (>= tgtk_sim_sparse_info out[0]). */
int tgtk_sim_sparse_trace(tgtk_sim *sim, int64_t *out, int32_t *planes, int cap);
/* The work-list planner of the sparse launches as a pure host function (no device needed; tests/test_sparse_plan.py).
 * mask[(x * nzt + zt) * nyt + yt]: one word per 8 x 32 block tile of plane x, byte w = the tile's rows 2w, 2w+1 hold a live
 * voxel.  Planes [x_lo, x_hi) are to be computed (a slab rank: its owned planes), `slots` = resident blocks of the device
 * for the kernel, slab != 0: items that start on x_lo or end on x_hi go last.  items[4*i + 0..3] = y tile, z tile, first
 * plane, end plane of item i (launch order); *count = number of items (items may be NULL to query it). */
int tgtk_sparse_plan_host(const uint32_t *mask, int X, int nzt, int nyt, int x_lo, int x_hi, int slots, int slab, int32_t *items,
                          int cap, int *count);

/* Slab decomposition support.  set_sum_range: only planes [x0, x1) of axis 0 count towards the
 * volume (a rank's owned planes; its ghost planes are stepped but not summed).  buffer_ptr: device
 * address of ping-pong buffer `which` (0/1) of a state field, so halo planes can be exchanged in
 * place (NCCL send/recv on views of these buffers); the layout is C order (X,Y,Z) of the
 * simulation's dtype with the pitches tgtk_sim_device_ptr reports (fp32 rows are padded to an even length).  After n enqueued steps the current state is buffer n & 1. */
int tgtk_sim_set_sum_range(tgtk_sim *sim, int x0, int x1);
int tgtk_sim_buffer_ptr(tgtk_sim *sim, int field, int which, unsigned long long *ptr);

/* Native slab driver (BASELINE.json configs[4]).  The simulation holds one rank's extended slab:
 * `halo` ghost planes, `owned` planes, `halo` ghost planes along axis 0.  tgtk_sim_run_slab
 * enqueues the whole decomposed loop on the simulation's stream: blocks of `halo` steps of the
 * ordinary periodic kernels, each followed by one grouped ncclSend/ncclRecv ring exchange of the
 * owned boundary planes into the neighbours' ghost planes (in place, no staging).  comm = NULL
 * closes the ring on the rank itself (one GPU).  NCCL is dlopen'ed; the unique id made on rank 0
 * travels to the other ranks by any means the caller has (the Python mirror broadcasts it with
 * torch.distributed).  Only for runs without a stop criterion. */
int tgtk_nccl_unique_id(char id[128]);
int tgtk_nccl_comm_create(void **comm, const char id[128], int rank, int world, int device);
void tgtk_nccl_comm_destroy(void *comm);
int tgtk_sim_run_slab(tgtk_sim *sim, void *comm, int nsteps, int halo, int owned, int prev, int next);

/* Peer-to-peer slab driver: no NCCL.  Each rank maps its neighbours' state buffers and a few flag
 * words (CUDA IPC, NVLink / NVSwitch peer access) and, after every block of `halo` steps, PUSHES
 * its owned boundary planes straight into the neighbours' ghost planes with peer copies, fenced
 * by two tiny signal-and-wait kernels (epoch counters in peer memory: "my block is done, you may
 * overwrite my ghosts" / "your ghosts are filled").  The whole block is replayed from a CUDA graph.
 *   tgtk_sim_ipc_export : TGTK_IPC_BYTES bytes describing this rank's buffers (send to neighbours)
 *   tgtk_sim_ipc_connect: map the previous / next rank's buffers (same_peer: world == 2, both
 *                          neighbours are one process; NULL handles: world == 1)
 * prev_owned: planes the previous rank owns (slabs differ by one plane when axis 0 does not divide).
 * The simulation must have been created with TGTK_CREATE_IPC.  Spin-waits are bounded (20 s); a time-out sets an error that tgtk_sim_sync reports. */
#define TGTK_IPC_BYTES (7 * 64)
int tgtk_sim_ipc_export(tgtk_sim *sim, unsigned char *out);
int tgtk_sim_ipc_connect(tgtk_sim *sim, const unsigned char *prev, const unsigned char *next, int same_peer);
int tgtk_sim_run_slab_p2p(tgtk_sim *sim, int nsteps, int halo, int owned, int prev_owned);
/* Fused slab driver (the default of slab.py): X = owned + 2, one ghost plane per side.  After this call tgtk_sim_run
 * steps planes [1, 1 + owned) only; the blocks of the STEP KERNEL that compute a boundary plane also store it into the
 * neighbouring rank's ghost plane (peer-mapped memory over NVLink) and publish the launch's epoch in the neighbour's
 * flag word, the blocks that read a ghost plane wait for the neighbour's previous epoch -- no exchange launches, no
 * copies between steps, interior blocks never wait.  The lower half of the plane segments marches upwards and the upper
 * half downwards, so both boundary planes leave at the START of a step.  When a stop criterion is active
 * (stopping_volume / small_volume) the volumes of each chunk of steps are summed over the ranks
 * (ncclAllReduce on `comm`; NULL: one rank) before the reference's stop tests (FK/FK.py:216-218,
 * FK_DTI/FK_DTI.py:199-206) run on them, so stopping_volume, the `volume < 1e-6` exit, snapshots and tgtk_sim_reset
 * work as on one GPU; without one tgtk_sim_volumes returns the rank's own sums.  All ranks must issue the same
 * run / sync / reset sequence. */
int tgtk_sim_slab_attach(tgtk_sim *sim, void *comm, int owned, int prev_owned);
/* 0, or which handshake timed out (p2p driver: 1 block done, 2 ghosts pushed; fused driver: 3). */
int tgtk_sim_slab_error(tgtk_sim *sim, int *err);

/* Host -> device.  `host` is float64 with element strides (sx,sy,sz); any strides are
 * accepted (rows with sz==1 are copied directly, otherwise gathered on the host).  If
 * `pinned` is non-zero the caller promises page-locked memory and the copy is fully
 * asynchronous on the simulation's stream. */
int tgtk_sim_upload(tgtk_sim *sim, int field, const double *host, const int64_t strides[3], int pinned);
/* Device -> host (float64, element strides). Synchronises the stream. */
int tgtk_sim_download(tgtk_sim *sim, int field, double *host, const int64_t strides[3]);

/* Finish set-up once all inputs are uploaded: builds the per-voxel coefficient arrays
 * (get_D, FK/FK.py:22-32; get_D(…,D_s,1), FK_2c/FK_2c.py:216), resets the step counter. */
int tgtk_sim_prepare(tgtk_sim *sim);

/* Restore the state saved by tgtk_sim_prepare and restart the step counter at 0 (device to
 * device; lets one uploaded problem be stepped repeatedly, e.g. by bench.py). */
int tgtk_sim_reset(tgtk_sim *sim);

/* Enqueue `nsteps` update steps (the body of the reference's `for t in range(...)` loops:
 * update, volume, stop test, snapshot).  `record_steps` (sorted, may be NULL) are absolute
 * step indices after which the state is snapshotted (FK/FK.py:221-223).  Asynchronous. */
int tgtk_sim_run(tgtk_sim *sim, int nsteps, const int32_t *record_steps, int nrecord);

/* Wait for the stream and report where the loop stands. */
int tgtk_sim_sync(tgtk_sim *sim, tgtk_status *status);

/* Per-step volumes of steps [first, first+count) as the device computed them. */
int tgtk_sim_volumes(tgtk_sim *sim, int first, int count, double *out);

/* Copy snapshot `index` (0-based, in record order) of `field` to the host. */
int tgtk_sim_snapshot(tgtk_sim *sim, int index, int field, double *host, const int64_t strides[3]);

/* Device pointer of the CURRENT state buffer of `field` (for zero-copy interop, e.g. halo
 * exchange with torch.distributed), its element pitches and dtype size. */
int tgtk_sim_device_ptr(tgtk_sim *sim, int field, unsigned long long *ptr, int64_t pitches[3], int *elem_size);

/* One-call convenience used by the Python mirror's public per-step methods
 * (FK_update(A, D_domain, ...), FK/FK.py:34-42): upload, one step, download. */
int tgtk_fk_update_host(const tgtk_params *params, int device, double *A, const int64_t a_strides[3],
                        const double *const coef[6], const int64_t coef_strides[6][3]);

/* FK_2c_update(P,N,S,D_domain,...,D_s_domain,...) on host arrays, in place
 * (FK_2c/FK_2c.py:48-73): D = D_plus_x,y,z,D_minus_x,y,z for P, Ds the same for S. */
int tgtk_fk2c_update_host(const tgtk_params *params, int device, double *P, const int64_t p_strides[3],
                          double *N, const int64_t n_strides[3], double *S, const int64_t s_strides[3],
                          const double *const D[6], const int64_t d_strides[6][3],
                          const double *const Ds[6], const int64_t ds_strides[6][3]);

/* m_Tildas / get_D / m_Tildas_with_necro / get_D_with_necro on the device, in the
 * reference's rounding order (FK/FK.py:10-32, FK_2c/FK_2c.py:10-46).  `what`: 0 = D_minus
 * faces Dw*(WM_t + GM_t/ratio), 1 = WM_tilda, 2 = GM_tilda.  P and N may both be NULL (no
 * necrotic closure).  Outputs are dense float64 (X,Y,Z); D_plus_a is np.roll(D_minus_a,1,a). */
int tgtk_faces_host(int device, int X, int Y, int Z, const double *wm, const int64_t wm_strides[3],
                    const double *gm, const int64_t gm_strides[3], const double *P, const int64_t p_strides[3],
                    const double *N, const int64_t n_strides[3], double th, double th_necro, double Dw,
                    double ratio, int logical_or, int what, double *out_x, double *out_y, double *out_z);

/* Device -> host with restore_tumor + zoom(order=1) applied on the way (FK/FK.py:229-238): the
 * current state of `field` is placed at `lo` in a zero volume of shape virt_shape and resampled
 * to out_shape on the device; only the full-resolution result crosses PCIe. */
int tgtk_sim_download_resampled(tgtk_sim *sim, int field, const int32_t lo[3], const int32_t virt_shape[3], double *out,
                                const int32_t out_shape[3]);

/* Asynchronous patient outputs (synthetic_gens/generatePatient_FK_2c.py:141-155 does the
 * same on the CPU after solve()).  Both calls only ENQUEUE on the simulation's stream -- restore_tumor + zoom(order=1)
 * read straight from the state (bit-identical to tgtk_sim_download_resampled), device-to-host copy -- and return;
 * `out` is valid after the next tgtk_sim_sync.  Pass buffers from tgtk_host_alloc (page-locked) so that the copies
 * run at link speed and overlap other simulations' kernels.  tgtk_sim_segmentation_async evaluates
 * create_segmentation_map (generatePatient_FK_2c.py:67-83) on the resampled P and N on the fly and returns labels
 * 0 / 1 / 3 / 4 as uint8 (the reference's map is int64 with the same values). */
int tgtk_sim_download_resampled_async(tgtk_sim *sim, int field, const int32_t lo[3], const int32_t virt_shape[3], double *out,
                                      const int32_t out_shape[3]);
int tgtk_sim_segmentation_async(tgtk_sim *sim, const int32_t lo[3], const int32_t virt_shape[3], const int32_t out_shape[3],
                                double th_necro_n, double th_enhancing_p, double th_edema_p, uint8_t *out);
/* What Solver.solve() does with the final state and with every recorded snapshot -- restore_tumor + zoom(order=1),
 * FK/FK.py:229-238 -- for `count` volumes in one call: snapshots [first, first + count) of `field` (first = -1, count = 1:
 * the current state) are resampled on the device straight from their arenas, copied through a ring of page-locked
 * staging buffers and written to `out` (count consecutive dense float64 volumes of out_shape, ordinary pageable memory,
 * e.g. a fresh numpy array) by several host threads while the next volume is in flight.  Synchronous. */
int tgtk_sim_fetch_series(tgtk_sim *sim, int field, int first, int count, const int32_t lo[3], const int32_t virt_shape[3],
                          const int32_t out_shape[3], double *out);
/* Page-locked host memory for the buffers of the host-pointer entry points (NULL on failure). */
void *tgtk_host_alloc(size_t bytes);
void tgtk_host_free(void *ptr);

/* scipy.ndimage.zoom(v, f, order=1) on the device (the resample of FK/FK.py:150-151,233-238).
 * `v` is a virtual float64 volume of shape virt_shape that is zero except for the box starting
 * at `lo` that holds `in` (restore_tumor, FK/FK.py:75-90, fused in; pass lo = 0 and
 * virt_shape = in_shape for a plain zoom).  `out` is dense float64 of out_shape, which the
 * caller chooses as round(n*f) like scipy does. */
int tgtk_resample_linear_host(int device, const double *in, const int64_t in_strides[3], const int32_t in_shape[3],
                              const int32_t lo[3], const int32_t virt_shape[3], double *out, const int32_t out_shape[3]);

/* tools.makeXYZ_rgb_from_tensor (TumorGrowthToolkit/FK_DTI/tools.py:166-187; FK_DTI_Solver.solve calls it on the whole
 * (240,240,155,3,3) volume before the loop, FK_DTI/FK_DTI.py:115) on the device: tensors = n C-order 3x3 float64 tensors
 * (host), rgb = n x 3 float64 (host): the tensor diagonal divided by its mean over the brain voxels (max of the diagonal
 * > 0), clipped to [0, 10], x**exponent + linear*x, normalised and clipped again.  Same arithmetic as the reference except
 * for the summation order of the two means (deterministic here, pairwise in numpy): agreement ~1e-15 relative. */
int tgtk_dti_rgb_host(int device, const double *tensors, int64_t n, double exponent, double linear, double *rgb);

/* Batched 3x3 symmetric eigen-elongation on the device: tensors = n C-order 3x3 float64 tensors (host),
 * out = n 3x3 float32 tensors (host), the dtype the reference returns.  Replaces
 * tools.elongate_tensor_along_main_axis_torch (TumorGrowthToolkit/FK_DTI/tools.py:109-164; called from
 * FK_DTI_Solver.solve, FK_DTI/FK_DTI.py:107-113, when diffusionEllipsoidScaling != 1): the largest
 * eigenvalue is multiplied by scale_factor, the other two are shifted as the reference shifts them, the
 * tensor is rebuilt.  Agreement with the reference's float32 LAPACK path is at float32 rounding level. */
int tgtk_elongate_tensors_host(int device, const double *tensors, int64_t n, double scale_factor, float *out);

#ifdef __cplusplus
}
#endif
#endif /* TGTK_B200_H */
