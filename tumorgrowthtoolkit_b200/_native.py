"""ctypes binding of libtgtk_b200.so (C ABI: include/tgtk_b200.h).

There is no CPU implementation behind this module: if the shared library is missing, or no
CUDA device is usable, every entry point raises.  Build the library with
`python -c "import __graft_entry__ as g; g.build()"` or `make -C tumorgrowthtoolkit_b200/csrc`.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TGTK_LIB_PATH") or os.path.join(_HERE, "csrc", "libtgtk_b200.so")   # override: A/B builds

MODEL_FK, MODEL_FK_2C, MODEL_DTI, MODEL_COEF6, MODEL_FK_FACES, MODEL_FK2C_COEF12, MODEL_DTI_FULL = 0, 1, 2, 3, 4, 5, 6
F64, F32 = 0, 1
FIELD_U = FIELD_P = 0
FIELD_N = 1
FIELD_S = 2
FIELD_WM, FIELD_GM = 10, 11
FIELD_COEF0 = 20
STOP_NONE, STOP_VOLUME, STOP_SMALL = 0, 1, 2

_I64x3 = ctypes.c_int64 * 3


class NativeError(RuntimeError):
    """A call into libtgtk_b200 failed (CUDA error, bad argument, missing library)."""


class Params(ctypes.Structure):
    _fields_ = [("model", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("X", ctypes.c_int32), ("Y", ctypes.c_int32), ("Z", ctypes.c_int32),
                ("logical_or_faces", ctypes.c_int32),
                ("h", ctypes.c_double * 3), ("dt", ctypes.c_double), ("rho", ctypes.c_double),
                ("Dw", ctypes.c_double), ("ratio", ctypes.c_double), ("th_matter", ctypes.c_double),
                ("th_necro", ctypes.c_double), ("Ds", ctypes.c_double), ("lambda_np", ctypes.c_double),
                ("sigma_np", ctypes.c_double), ("lambda_s", ctypes.c_double),
                ("stopping_volume", ctypes.c_double), ("small_volume", ctypes.c_double)]


class Status(ctypes.Structure):
    _fields_ = [("steps_run", ctypes.c_int32), ("stop_reason", ctypes.c_int32),
                ("stop_step", ctypes.c_int32), ("snapshots", ctypes.c_int32),
                ("last_volume", ctypes.c_double), ("loop_ms", ctypes.c_double),
                ("kernel_launches", ctypes.c_int64)]


_lib = None


def lib():
    """The loaded shared library; raises NativeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run __graft_entry__.build() or `make -C tumorgrowthtoolkit_b200/csrc`). "
            "There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    if os.environ.get("TGTK_LIB_PATH"):
        # A/B runs against an older build of the library (tools/sweep_env.py): entry points it lacks are never called there
        class _Tolerant:
            def __init__(self, lib):
                object.__setattr__(self, "_lib", lib)

            def __getattr__(self, name):
                try:
                    return getattr(self._lib, name)
                except AttributeError:
                    class _Missing:
                        argtypes = restype = None
                    return _Missing()
        L = _Tolerant(L)
    L.tgtk_last_error.restype = ctypes.c_char_p
    L.tgtk_abi_version.restype = ctypes.c_int
    vp, dp, ip = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    L.tgtk_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.tgtk_sim_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(Params), ctypes.c_int, vp]
    L.tgtk_sim_destroy.argtypes = [vp]
    L.tgtk_sim_destroy.restype = None
    L.tgtk_sim_upload.argtypes = [vp, ctypes.c_int, vp, _I64x3, ctypes.c_int]
    L.tgtk_sim_download.argtypes = [vp, ctypes.c_int, vp, _I64x3]
    L.tgtk_sim_prepare.argtypes = [vp]
    L.tgtk_sim_create_ex.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(Params), ctypes.c_int, vp, ctypes.c_int]
    L.tgtk_sim_ipc_export.argtypes = [vp, ctypes.c_char_p]
    L.tgtk_sim_ipc_connect.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    L.tgtk_sim_run_slab_p2p.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.tgtk_sim_slab_error.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    L.tgtk_sim_slab_attach.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
    L.tgtk_sim_run_slab.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.tgtk_nccl_unique_id.argtypes = [ctypes.c_char_p]
    L.tgtk_nccl_comm_create.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.tgtk_nccl_comm_destroy.argtypes = [vp]
    L.tgtk_nccl_comm_destroy.restype = None
    L.tgtk_sim_set_sum_range.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    L.tgtk_sim_buffer_ptr.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
    i3 = ctypes.c_int32 * 3
    L.tgtk_sim_download_resampled.argtypes = [vp, ctypes.c_int, i3, i3, vp, i3]
    L.tgtk_sim_set_variant.argtypes = [vp, ctypes.c_int]
    i3p = ctypes.POINTER(ctypes.c_int32 * 3)
    L.tgtk_sim_download_resampled_async.argtypes = [vp, ctypes.c_int, i3p, i3p, ctypes.c_void_p, i3p]
    L.tgtk_sim_segmentation_async.argtypes = [vp, i3p, i3p, i3p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    L.tgtk_sim_fetch_series.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, i3p, i3p, i3p, ctypes.c_void_p]
    L.tgtk_host_alloc.argtypes = [ctypes.c_size_t]
    L.tgtk_host_alloc.restype = ctypes.c_void_p
    L.tgtk_host_free.argtypes = [ctypes.c_void_p]
    L.tgtk_host_free.restype = None
    L.tgtk_sim_get_tuning.argtypes = [vp, ctypes.POINTER(ctypes.c_int32 * 10)]
    L.tgtk_sim_sparse_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int64 * 4)]
    L.tgtk_sim_sparse_trace.argtypes = [vp, ctypes.POINTER(ctypes.c_int64), ip, ctypes.c_int]
    L.tgtk_sparse_plan_host.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.tgtk_sim_reset.argtypes = [vp]
    L.tgtk_sim_run.argtypes = [vp, ctypes.c_int, ip, ctypes.c_int]
    L.tgtk_sim_sync.argtypes = [vp, ctypes.POINTER(Status)]
    L.tgtk_sim_volumes.argtypes = [vp, ctypes.c_int, ctypes.c_int, dp]
    L.tgtk_sim_snapshot.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, _I64x3]
    L.tgtk_sim_device_ptr.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong), _I64x3,
                                      ctypes.POINTER(ctypes.c_int)]
    # host-pointer entry points behind the reference's public per-step methods (_ops.py)
    c_int, c_dbl = ctypes.c_int, ctypes.c_double
    i64x6x3, ptr6 = (ctypes.c_int64 * 3) * 6, ctypes.c_void_p * 6
    L.tgtk_fk_update_host.argtypes = [ctypes.POINTER(Params), c_int, vp, _I64x3, ptr6, i64x6x3]
    L.tgtk_fk2c_update_host.argtypes = [ctypes.POINTER(Params), c_int, vp, _I64x3, vp, _I64x3, vp, _I64x3,
                                        ptr6, i64x6x3, ptr6, i64x6x3]
    L.tgtk_faces_host.argtypes = [c_int, c_int, c_int, c_int, vp, _I64x3, vp, _I64x3, vp, _I64x3, vp, _I64x3,
                                  c_dbl, c_dbl, c_dbl, c_dbl, c_int, c_int, vp, vp, vp]
    L.tgtk_resample_linear_host.argtypes = [c_int, vp, _I64x3, i3, i3, i3, vp, i3]
    L.tgtk_elongate_tensors_host.argtypes = [c_int, vp, ctypes.c_int64, c_dbl, vp]
    L.tgtk_dti_rgb_host.argtypes = [c_int, vp, ctypes.c_int64, c_dbl, c_dbl, vp]
    if L.tgtk_abi_version() != 1:
        raise NativeError("libtgtk_b200.so ABI version mismatch")
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise NativeError(lib().tgtk_last_error().decode("utf-8", "replace"))


def device_count():
    n = ctypes.c_int(0)
    _check(lib().tgtk_device_count(ctypes.byref(n)))
    return n.value


def _as_f64(a):
    """float64 view of `a` without copying when it already is float64 (any strides)."""
    a = np.asarray(a)
    if a.dtype != np.float64:
        a = a.astype(np.float64)
    return a


def _strides(a):
    return _I64x3(*[s // a.itemsize for s in a.strides])


def sparse_plan(mask, x_lo=0, x_hi=None, slots=592, slab=False):
    """The work list of a sparse launch for a tile mask of shape (X, nzt, nyt), uint32 (include/tgtk_b200.h:
    tgtk_sparse_plan_host; host code only).  Returns an (items, 4) int32 array: y tile, z tile, first plane, end plane."""
    mask = np.ascontiguousarray(mask, dtype=np.uint32)
    X, nzt, nyt = mask.shape
    x_hi = X if x_hi is None else x_hi
    n = ctypes.c_int(0)
    mp = mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    _check(lib().tgtk_sparse_plan_host(mp, X, nzt, nyt, int(x_lo), int(x_hi), int(slots), int(bool(slab)), None, 0, ctypes.byref(n)))
    out = np.zeros((n.value, 4), dtype=np.int32)
    _check(lib().tgtk_sparse_plan_host(mp, X, nzt, nyt, int(x_lo), int(x_hi), int(slots), int(bool(slab)),
                                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n.value, ctypes.byref(n)))
    return out


def make_params(model, dtype, shape, h, dt, rho=0.0, Dw=0.0, ratio=1.0, th_matter=0.0, th_necro=0.0,
                Ds=0.0, lambda_np=0.0, sigma_np=0.0, lambda_s=0.0, stopping_volume=np.inf,
                small_volume=0.0, logical_or_faces=0):
    p = Params()
    p.model, p.dtype = int(model), int(dtype)
    p.X, p.Y, p.Z = (int(s) for s in shape)
    p.logical_or_faces = int(logical_or_faces)
    p.h[0], p.h[1], p.h[2] = (float(v) for v in h)
    p.dt, p.rho, p.Dw, p.ratio, p.th_matter = float(dt), float(rho), float(Dw), float(ratio), float(th_matter)
    p.th_necro, p.Ds, p.lambda_np = float(th_necro), float(Ds), float(lambda_np)
    p.sigma_np, p.lambda_s = float(sigma_np), float(lambda_s)
    p.stopping_volume, p.small_volume = float(stopping_volume), float(small_volume)
    return p


class _PinnedBlock:
    """Owns one tgtk_host_alloc allocation; numpy arrays made by `pinned_empty` keep it alive through .base."""

    def __init__(self, nbytes):
        self.ptr = lib().tgtk_host_alloc(int(nbytes))
        if not self.ptr:
            raise NativeError("tgtk_host_alloc: " + lib().tgtk_last_error().decode())
        self.nbytes = int(nbytes)
        self.__array_interface__ = dict(shape=(max(self.nbytes, 1),), typestr="|u1", data=(int(self.ptr), False), version=3)

    def __del__(self):
        try:
            if self.ptr:
                lib().tgtk_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (tgtk_host_alloc): device-to-host copies into it run at link speed
    and asynchronously.  The memory is released when the last view of the array is garbage collected."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    raw = np.asarray(_PinnedBlock(n))
    return raw[:n].view(dtype).reshape(shape)


class Sim:
    """One device-resident simulation (tgtk_sim*)."""

    def __init__(self, params, device=0, stream=None, ipc=False):
        self._h = ctypes.c_void_p()
        self.params = params
        self.shape = (params.X, params.Y, params.Z)
        _check(lib().tgtk_sim_create_ex(ctypes.byref(self._h), ctypes.byref(params), int(device),
                                        ctypes.c_void_p(stream) if stream else None, 1 if ipc else 0))

    def close(self):
        if self._h:
            lib().tgtk_sim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, field, array, pinned=False):
        a = _as_f64(array)
        if a.shape != self.shape:
            raise ValueError(f"array shape {a.shape} does not match the simulation box {self.shape}")
        _check(lib().tgtk_sim_upload(self._h, int(field), a.ctypes.data_as(ctypes.c_void_p), _strides(a),
                                     1 if pinned else 0))
        # uploads from pageable memory have completed on return; pinned ones are asynchronous and
        # the caller keeps `array` alive until the next sync()

    def prepare(self):
        _check(lib().tgtk_sim_prepare(self._h))

    def set_variant(self, variant):
        _check(lib().tgtk_sim_set_variant(self._h, int(variant)))

    def tuning(self):
        """What the launch heuristics chose (after prepare): kernel family, planes per segment, L2 prefetch
        distance, programmatic dependent launch, grid, threads per block."""
        out = (ctypes.c_int32 * 10)()
        _check(lib().tgtk_sim_get_tuning(self._h, ctypes.byref(out)))
        return dict(variant=out[0], planes_per_segment=out[1], l2_prefetch_planes=out[2], pdl=bool(out[3]),
                    grid=(out[4], out[5], out[6]), threads=out[7], steps_per_launch=out[8])

    def sparse_info(self):
        """Sparse launches (include/tgtk_b200.h): work items per launch (0 = dense), voxels in live tiles, voxels."""
        out = (ctypes.c_int64 * 4)()
        fn = lib().tgtk_sim_sparse_info
        if not callable(fn):            # an older build loaded through TGTK_LIB_PATH
            n = int(np.prod(self.shape))
            return dict(items=0, live_voxels=n, voxels=n, blocks_per_sm=0)
        _check(fn(self._h, ctypes.byref(out)))
        return dict(items=int(out[0]), live_voxels=int(out[1]), voxels=int(out[2]), blocks_per_sm=int(out[3]))

    def sparse_trace(self):
        """One more step with per-item timestamps: (items, 4) int64 [start ns, prologue end ns, loop end ns, SM], (items, 2) planes."""
        n = self.sparse_info()["items"]
        out = np.zeros((n, 4), dtype=np.int64)
        planes = np.zeros((n, 2), dtype=np.int32)
        _check(lib().tgtk_sim_sparse_trace(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                           planes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n))
        return out, planes

    def run_slab(self, comm, nsteps, halo, owned, prev, nxt):
        _check(lib().tgtk_sim_run_slab(self._h, comm, int(nsteps), int(halo), int(owned), int(prev), int(nxt)))

    IPC_BYTES = 7 * 64

    def ipc_export(self):
        buf = ctypes.create_string_buffer(self.IPC_BYTES)
        _check(lib().tgtk_sim_ipc_export(self._h, buf))
        return buf.raw

    def ipc_connect(self, prev, nxt, same_peer=False):
        _check(lib().tgtk_sim_ipc_connect(self._h, prev, nxt, 1 if same_peer else 0))

    def run_slab_p2p(self, nsteps, halo, owned, prev_owned):
        _check(lib().tgtk_sim_run_slab_p2p(self._h, int(nsteps), int(halo), int(owned), int(prev_owned)))

    def slab_attach(self, comm, owned, prev_owned):
        _check(lib().tgtk_sim_slab_attach(self._h, comm, int(owned), int(prev_owned)))

    def slab_error(self):
        e = ctypes.c_int(0)
        _check(lib().tgtk_sim_slab_error(self._h, ctypes.byref(e)))
        return e.value

    def set_sum_range(self, x0, x1):
        _check(lib().tgtk_sim_set_sum_range(self._h, int(x0), int(x1)))

    def buffer_ptr(self, field, which):
        ptr = ctypes.c_ulonglong(0)
        _check(lib().tgtk_sim_buffer_ptr(self._h, int(field), int(which), ctypes.byref(ptr)))
        return ptr.value

    def reset(self):
        _check(lib().tgtk_sim_reset(self._h))

    def run(self, nsteps, record_steps=None):
        if record_steps is None or len(record_steps) == 0:
            _check(lib().tgtk_sim_run(self._h, int(nsteps), None, 0))
        else:
            rec = np.ascontiguousarray(np.unique(np.asarray(record_steps)), dtype=np.int32)
            _check(lib().tgtk_sim_run(self._h, int(nsteps),
                                      rec.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(rec.size)))

    def sync(self):
        st = Status()
        _check(lib().tgtk_sim_sync(self._h, ctypes.byref(st)))
        return st

    def download(self, field, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        _check(lib().tgtk_sim_download(self._h, int(field), out.ctypes.data_as(ctypes.c_void_p), _strides(out)))
        return out

    def download_resampled(self, field, lo, virt_shape, out_shape, out=None):
        """restore_tumor + zoom(order=1) on the device, then D2H of the full-resolution field."""
        i3 = ctypes.c_int32 * 3
        if out is None:
            out = np.empty(tuple(int(v) for v in out_shape), dtype=np.float64)
        _check(lib().tgtk_sim_download_resampled(self._h, int(field), i3(*[int(v) for v in lo]),
                                                 i3(*[int(v) for v in virt_shape]), out.ctypes.data_as(ctypes.c_void_p),
                                                 i3(*out.shape)))
        return out

    def download_resampled_async(self, field, lo, virt_shape, out):
        """Enqueue restore_tumor + zoom(order=1) + the copy into `out` (float64, C-contiguous, ideally from
        pinned_empty) on the simulation's stream; `out` is valid after the next sync()."""
        i3 = ctypes.c_int32 * 3
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.ndim == 3
        _check(lib().tgtk_sim_download_resampled_async(self._h, int(field), ctypes.byref(i3(*[int(v) for v in lo])),
                                                       ctypes.byref(i3(*[int(v) for v in virt_shape])), out.ctypes.data,
                                                       ctypes.byref(i3(*out.shape))))
        return out

    def fetch_series(self, field, first, count, lo, virt_shape, out):
        """restore_tumor + zoom(order=1) of snapshots [first, first + count) of `field` (first = -1, count = 1: the current
        state) into `out` (float64, C-contiguous, (count, X, Y, Z) or (X, Y, Z)): resampled on the device, staged through
        page-locked buffers, written by several host threads.  Returns `out`."""
        i3 = ctypes.c_int32 * 3
        assert out.dtype == np.float64 and out.flags.c_contiguous
        shape = out.shape[-3:]
        assert out.size == int(count) * int(np.prod(shape, dtype=np.int64))
        _check(lib().tgtk_sim_fetch_series(self._h, int(field), int(first), int(count), ctypes.byref(i3(*[int(v) for v in lo])),
                                           ctypes.byref(i3(*[int(v) for v in virt_shape])), ctypes.byref(i3(*[int(v) for v in shape])),
                                           out.ctypes.data))
        return out

    def segmentation_async(self, lo, virt_shape, th_necro_n, th_enhancing_p, th_edema_p, out):
        """Enqueue create_segmentation_map (generatePatient_FK_2c.py:67-83) of the full-resolution P and N, computed on
        the device, into `out` (uint8 labels 0/1/3/4); valid after the next sync()."""
        i3 = ctypes.c_int32 * 3
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.ndim == 3
        _check(lib().tgtk_sim_segmentation_async(self._h, ctypes.byref(i3(*[int(v) for v in lo])),
                                                 ctypes.byref(i3(*[int(v) for v in virt_shape])), ctypes.byref(i3(*out.shape)),
                                                 float(th_necro_n), float(th_enhancing_p), float(th_edema_p), out.ctypes.data))
        return out

    def snapshot(self, index, field, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        _check(lib().tgtk_sim_snapshot(self._h, int(index), int(field), out.ctypes.data_as(ctypes.c_void_p),
                                       _strides(out)))
        return out

    def volumes(self, first, count):
        out = np.empty(int(count), dtype=np.float64)
        if count:
            _check(lib().tgtk_sim_volumes(self._h, int(first), int(count),
                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out

    def device_ptr(self, field):
        ptr = ctypes.c_ulonglong(0)
        pitches = _I64x3()
        es = ctypes.c_int(0)
        _check(lib().tgtk_sim_device_ptr(self._h, int(field), ctypes.byref(ptr), pitches, ctypes.byref(es)))
        return ptr.value, tuple(pitches), es.value


def nccl_unique_id():
    buf = ctypes.create_string_buffer(128)
    _check(lib().tgtk_nccl_unique_id(buf))
    return buf.raw


def nccl_comm_create(uid, rank, world, device):
    comm = ctypes.c_void_p()
    _check(lib().tgtk_nccl_comm_create(ctypes.byref(comm), uid, int(rank), int(world), int(device)))
    return comm


def nccl_comm_destroy(comm):
    if comm:
        lib().tgtk_nccl_comm_destroy(comm)
