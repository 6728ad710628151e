"""Slab decomposition of one solve over the GPUs of a box (BASELINE.json configs[4]).

The reference has no domain decomposition; this is the spatial analogue of its
single time loop (FK/FK.py:212-226, FK_2c/FK_2c.py:227-245, FK_DTI/FK_DTI.py:195-214) for grids
that exceed one GPU's useful working set, e.g. the 2x up-sampled 480x480x310 atlas.

  * axis 0 (slowest, so a halo plane is one contiguous chunk) is split into `world` slabs, one
    process per GPU; the ring is periodic (rank 0 <-> rank n-1) because the reference's
    neighbour access is np.roll (SURVEY.md finding 2);
  * every rank carries `halo` ghost planes on each side and steps its whole extended slab with
    the ordinary periodic kernels for `halo` time steps between exchanges: after j steps the j
    outermost ghost planes are stale, the owned planes never are (overlapped halo recompute).
    One exchange per `halo` steps amortises its latency, which for a 1/8 slab is longer than a
    step (SURVEY.md 7 "halo latency");
  * the exchange is in place on the simulations' own device buffers, no staging copies.  Four drivers:
    "fused" (default, halo 1): the STEP KERNEL itself stores the rank's two boundary planes into the neighbours'
    ghost planes (CUDA-IPC mapped peer memory, NVLink) and publishes an epoch word there; only the thread blocks
    that read a ghost plane wait for the neighbour's previous epoch, interior blocks never do -- no exchange
    launches, no ghost recompute, and the transfer of step t overlaps the interior of step t+1.  Stop criteria
    (`stopping_volume`, FK_DTI's `volume < 1e-6`), snapshots and reset work as on one GPU: each chunk's volumes
    are summed over the ranks with one small ncclAllReduce before the reference's stop tests run on them.
    "p2p" maps the neighbours' buffers the same way and PUSHES the boundary planes with peer copies between
    graph-replayed blocks of `halo` steps, fenced by epoch flags; "native" enqueues grouped ncclSend/ncclRecv
    from C; "torch" uses torch.distributed send/recv per block (NCCL on GPU tensors, gloo in the CPU tests);
  * volumes are summed over owned planes only and reduced over the ranks.
"""
import numpy as np


class SlabPlan:
    """Which planes of axis 0 a rank owns and where its halos go (pure host logic)."""

    def __init__(self, X, world, rank, halo):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank / world")
        base, rem = divmod(int(X), world)
        sizes = [base + (1 if r < rem else 0) for r in range(world)]
        if min(sizes) < 1:
            raise ValueError(f"{X} planes cannot be split over {world} ranks")
        halo = int(halo)
        if halo < 1 or halo > min(sizes):
            raise ValueError(f"halo depth {halo} must be between 1 and the smallest slab ({min(sizes)} planes)")
        self.X, self.world, self.rank, self.halo, self.sizes = int(X), world, rank, halo, sizes
        self.x0 = sum(sizes[:rank])
        self.n = sizes[rank]
        self.x1 = self.x0 + self.n
        self.prev, self.next = (rank - 1) % world, (rank + 1) % world
        self.local_planes = self.n + 2 * halo

    def indices(self):
        """Global plane index of every local plane (ghosts wrap periodically)."""
        return np.arange(self.x0 - self.halo, self.x1 + self.halo) % self.X

    def extend(self, a):
        """Local extended slab (owned planes + ghosts) of a global array split along axis 0."""
        return np.ascontiguousarray(np.asarray(a)[self.indices()])

    @property
    def owned(self):
        return slice(self.halo, self.halo + self.n)

    # local plane ranges of the four halo messages
    @property
    def send_up(self):      # my top owned planes -> next rank's lower ghosts
        return slice(self.n, self.n + self.halo)

    @property
    def send_down(self):    # my bottom owned planes -> previous rank's upper ghosts
        return slice(self.halo, 2 * self.halo)

    @property
    def ghost_lo(self):
        return slice(0, self.halo)

    @property
    def ghost_hi(self):
        return slice(self.halo + self.n, 2 * self.halo + self.n)


def exchange_halos(plan, tensors, dist):
    """In-place ring exchange of `plan.halo` planes per side for every tensor (axis 0 = planes).
    `dist` is torch.distributed (NCCL on GPU tensors, gloo on CPU tensors) or None when
    world == 1 (the ring closes on the rank itself: a periodic copy)."""
    if plan.world == 1:
        for t in tensors:
            t[plan.ghost_lo] = t[plan.send_up]
            t[plan.ghost_hi] = t[plan.send_down]
        return
    ops = []
    for t in tensors:
        ops.append(dist.P2POp(dist.isend, t[plan.send_up], plan.next))
        ops.append(dist.P2POp(dist.isend, t[plan.send_down], plan.prev))
    for t in tensors:
        ops.append(dist.P2POp(dist.irecv, t[plan.ghost_lo], plan.prev))
        ops.append(dist.P2POp(dist.irecv, t[plan.ghost_hi], plan.next))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def run_blocks(plan, stepper, nsteps, dist):
    """The decomposed time loop: `halo` steps, one exchange, repeat.  `stepper.run(k)` advances
    the local extended slab k steps; `stepper.tensors()` returns the current state tensors."""
    done = 0
    while done < nsteps:
        k = min(plan.halo, nsteps - done)
        stepper.run(k)
        done += k
        if done < nsteps:
            with stepper.context():      # the stream the step kernels run on must be current
                exchange_halos(plan, stepper.tensors(), dist)


class _RawCuda:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


class CudaSlabStepper:
    """One rank's extended slab as a libtgtk_b200 simulation on torch's current CUDA stream."""

    def __init__(self, plan, params, uploads, state_fields, device, ipc=False):
        import torch
        from . import _native as nat
        self.torch, self.nat, self.plan, self.device = torch, nat, plan, device
        torch.cuda.set_device(device)
        # a dedicated torch stream: the library launches on it and torch.distributed orders its
        # NCCL transfers after / before the work of the stream that is current when they are issued
        self.tstream = torch.cuda.Stream(device=device)
        self.sim = nat.Sim(params, device, stream=self.tstream.cuda_stream, ipc=ipc)
        for field, arr in uploads:
            self.sim.upload(field, arr)
        self.sim.prepare()
        self.sim.set_sum_range(plan.halo, plan.halo + plan.n)
        self.fields = state_fields
        self.steps = 0
        # planes as flat rows of the library's plane pitch (fp32 rows are padded to an even length): the
        # exchange only slices axis 0, and whole planes -- pad elements included -- are what must travel
        _, pitches, _ = self.sim.device_ptr(state_fields[0])
        shape = (params.X, int(pitches[0]))
        typestr = "<f8" if params.dtype == nat.F64 else "<f4"
        self.views = {(f, w): torch.as_tensor(_RawCuda(self.sim.buffer_ptr(f, w), shape, typestr), device=f"cuda:{device}")
                      for f in state_fields for w in (0, 1)}

    def context(self):
        return self.torch.cuda.stream(self.tstream)

    def run(self, k):
        self.sim.run(k)
        self.steps += k

    def tensors(self):
        return [self.views[(f, self.steps & 1)] for f in self.fields]

    def finish(self):
        st = self.sim.sync()
        vols = self.sim.volumes(0, self.steps)
        owned = {f: self.sim.download(f)[self.plan.owned] for f in self.fields}
        self.sim.close()
        return st, vols, owned


def gather_planes(plan, own, dist):
    """All ranks' owned planes (a torch tensor, planes on axis 0) concatenated in rank order on
    every rank.  Slabs may differ by one plane, all_gather wants equal sizes: pad, then trim."""
    import torch
    if plan.world == 1:
        return own
    nmax = max(plan.sizes)
    padded = torch.zeros((nmax,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    padded[:own.shape[0]] = own
    parts = [torch.empty_like(padded) for _ in range(plan.world)]
    dist.all_gather(parts, padded)
    return torch.cat([part[:n] for part, n in zip(parts, plan.sizes)], dim=0)


def _gather_planes(plan, owned, dist, torch, device):
    t = torch.from_numpy(np.ascontiguousarray(owned))
    if plan.world > 1 and device is not None:
        t = t.to(f"cuda:{device}")
    return gather_planes(plan, t, dist).cpu().numpy()


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


_COMMS = {}


def _native_comm(dist, rank, world, device):
    """One NCCL communicator per (world, device) owned by the library; the unique id made on
    rank 0 is broadcast through the torch.distributed group the caller initialised."""
    import torch
    from . import _native as nat
    key = (world, rank, device)
    if key not in _COMMS:
        t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{device}")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(nat.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        _COMMS[key] = nat.nccl_comm_create(bytes(t.cpu().numpy().tobytes()), rank, world, device)
    return _COMMS[key]


class SlabSession:
    """One rank's share of a slab-decomposed solve, kept alive across runs (bench.py times repeated loops on it;
    `_run_cuda` is create + one run + gather).  All ranks of the torch.distributed group construct it together."""

    def __init__(self, model, fields_in, statics, scalars, shape, h, dt, halo, precision, device, names, driver="p2p"):
        import torch
        from . import _native as nat
        from ._loop import dtype_code
        self.torch, self.nat, self.driver, self.device, self.names = torch, nat, driver, device, names
        self.dist, self.rank, self.world = _dist()
        if driver == "fused" and halo != 1:
            raise ValueError("the fused driver exchanges one plane per side every step: halo must be 1")
        self.plan = plan = SlabPlan(shape[0], self.world, self.rank, halo)
        local_shape = (plan.local_planes,) + tuple(shape[1:])
        p = nat.make_params(model, dtype_code(precision), local_shape, h, dt, **scalars)
        # a stop criterion makes the library sum every chunk's volumes over the ranks on the device; without one the
        # log holds the rank's own sums and `volumes()` reduces it once
        self.has_stop = bool(np.isfinite(scalars.get("stopping_volume", np.inf)) or scalars.get("small_volume", 0.0) > 0)
        uploads = [(f, plan.extend(a)) for f, a in statics] + [(f, plan.extend(a)) for f, a in fields_in]
        self.fields = [f for f, _ in fields_in]
        self.stepper = CudaSlabStepper(plan, p, uploads, self.fields, device, ipc=(driver in ("p2p", "fused")))
        self.sim = self.stepper.sim
        self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.comm = None
        if driver in ("p2p", "fused"):
            # peer-to-peer pushes over NVLink: map the neighbours' buffers (CUDA IPC handles travel over the
            # caller's torch.distributed group)
            mine = self.sim.ipc_export()
            if self.world > 1:
                t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(f"cuda:{device}")
                allh = [torch.empty_like(t) for _ in range(self.world)]
                self.dist.all_gather(allh, t)
                hb = [bytes(h.cpu().numpy().tobytes()) for h in allh]
                self.sim.ipc_connect(hb[plan.prev], hb[plan.next], same_peer=(plan.prev == plan.next))
            else:
                self.sim.ipc_connect(None, None)
            if driver == "fused":
                # stop criteria need the global volume: one small NCCL all-reduce per chunk of steps
                self.comm = _native_comm(self.dist, self.rank, self.world, device) if self.world > 1 else None
                self.sim.slab_attach(self.comm, plan.n, plan.sizes[plan.prev])
        elif driver == "native":
            self.comm = _native_comm(self.dist, self.rank, self.world, device) if self.world > 1 else None
        self._barrier()
        self._torch_ms = 0.0

    def _barrier(self):
        self.stepper.tstream.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def reset(self):
        """Back to the uploaded initial state (ghost planes included); collective."""
        self.sim.reset()
        self.stepper.steps = 0
        self._barrier()

    def run(self, nsteps, record_steps=None):
        plan, sim = self.plan, self.sim
        if record_steps is not None and self.driver != "fused":
            raise ValueError("snapshots inside a decomposed run need driver='fused'")
        if self.driver == "fused":
            sim.run(nsteps, record_steps)                # tgtk_sim_run in slab mode: pushes and flag waits inside the step kernels
            self.stepper.steps += nsteps
        elif self.driver == "p2p":
            sim.run_slab_p2p(nsteps, plan.halo, plan.n, plan.sizes[plan.prev])
            self.stepper.steps += nsteps
        elif self.driver == "native":
            sim.run_slab(self.comm, nsteps, plan.halo, plan.n, plan.prev, plan.next)
            self.stepper.steps += nsteps
        else:
            self.ev0.record(self.stepper.tstream)
            run_blocks(plan, self.stepper, nsteps, self.dist)       # torch.distributed send/recv per block
            self.ev1.record(self.stepper.tstream)

    def sync(self):
        """Wait for the enqueued loop; returns the library's status (collective for the fused driver: a tripped
        stop criterion is replayed from the checkpoint by all ranks in lockstep)."""
        st = self.sim.sync()
        if self.driver in ("p2p", "fused"):
            err = self.sim.slab_error()
            if err:
                raise self.nat.NativeError(f"peer-to-peer halo handshake {err} timed out (a neighbouring rank stalled)")
        self.status = st
        return st

    def sync_ms(self):
        st = self.sync()
        if self.driver == "torch":
            self.stepper.tstream.synchronize()
            return float(self.ev0.elapsed_time(self.ev1))
        return float(st.loop_ms)

    def launches(self):
        return int(self.sim.sync().kernel_launches)

    def volumes(self):
        """Per-step volumes summed over ranks."""
        torch = self.torch
        st = self.sync()
        vols = self.sim.volumes(0, st.steps_run)
        if self.world > 1 and not (self.driver == "fused" and self.has_stop):
            v = torch.from_numpy(vols).to(f"cuda:{self.device}")
            self.dist.all_reduce(v)
            vols = v.cpu().numpy()
        return vols

    def volume(self):
        v = self.volumes()
        return float(v[-1]) if len(v) else None

    def gather(self, snapshot=None):
        """The global state (or snapshot `snapshot`) on every rank: array, or {'P','N','S'}."""
        self.sync()
        out = {}
        for n, f in zip(self.names, self.fields):
            loc = self.sim.download(f) if snapshot is None else self.sim.snapshot(snapshot, f)
            out[n] = _gather_planes(self.plan, loc[self.plan.owned], self.dist, self.torch, self.device)
        return out[self.names[0]] if len(self.names) == 1 else out

    def close(self):
        if self.sim is not None:
            self.stepper.tstream.synchronize()
            if self.world > 1:
                self.dist.barrier()      # nobody unmaps buffers a neighbour may still be pushing into
            self.sim.close()
            self.sim = None


def _run_cuda(model, fields_in, statics, scalars, shape, h, dt, nsteps, halo, precision, device, names, driver="p2p",
              record_steps=None):
    sess = SlabSession(model, fields_in, statics, scalars, shape, h, dt, halo, precision, device, names, driver)
    sess.run(nsteps, record_steps)
    loop_ms = sess.sync_ms()
    st = sess.status
    vols = sess.volumes()
    state = sess.gather()
    snaps = None
    if record_steps is not None:
        per = [sess.gather(snapshot=i) for i in range(st.snapshots)]
        snaps = per if len(names) == 1 else {n: [q[n] for q in per] for n in names}
    sess.close()
    nat = sess.nat
    stop = {nat.STOP_NONE: "time", nat.STOP_VOLUME: "volume", nat.STOP_SMALL: "small"}[st.stop_reason]
    steps_run = int(st.steps_run) if driver == "fused" else int(nsteps)
    return dict(state=state, volume=float(vols[-1]) if len(vols) else None, steps_run=steps_run, stop=stop,
                t_stop=(int(st.stop_step) if st.stop_reason == nat.STOP_VOLUME else None), snapshots=snaps, volumes=vols,
                loop_ms=float(loop_ms), launches=int(st.kernel_launches), plan=sess.plan)


def _scalars(stopping_volume=np.inf, small_volume=0.0, **kw):
    if not np.isinf(stopping_volume):
        kw["stopping_volume"] = stopping_volume
    if small_volume:
        kw["small_volume"] = small_volume
    return kw


def _stop_args(driver, stopping_volume, small_volume, record_steps):
    if driver != "fused" and (not np.isinf(stopping_volume) or small_volume or record_steps is not None):
        raise ValueError("stop criteria and snapshots inside a decomposed run need driver='fused'")


def fk_loop(u0, WM, GM, th, Dw, ratio, rho, dt, dx, dy, dz, nsteps, halo=1, precision="fp64", device=0, driver="fused",
            stopping_volume=np.inf, record_steps=None):
    """FK.py:200,212-226 on a slab-decomposed grid (all ranks pass the same global arrays)."""
    from . import _native as nat
    _stop_args(driver, stopping_volume, 0.0, record_steps)
    return _run_cuda(nat.MODEL_FK, [(nat.FIELD_U, u0)], [(nat.FIELD_WM, WM), (nat.FIELD_GM, GM)],
                     _scalars(stopping_volume, rho=rho, Dw=Dw, ratio=ratio, th_matter=th), np.shape(u0), (dx, dy, dz), dt, nsteps,
                     halo, precision, device, ("u",), driver, record_steps)


def _dti_args(u0, rgb, Dw, rho, stopping_volume, small_volume):
    from . import _native as nat
    rgb = np.asarray(rgb)
    statics = [(nat.FIELD_COEF0 + a, rgb[..., a] * Dw) for a in range(3)]
    return nat.MODEL_DTI, [(nat.FIELD_U, u0)], statics, _scalars(stopping_volume, small_volume, rho=rho, Dw=Dw)


def dti_loop(u0, rgb, Dw, rho, dt, dx, dy, dz, nsteps, halo=1, precision="fp64", device=0, driver="fused",
             stopping_volume=np.inf, small_volume=0.0, record_steps=None):
    """FK_DTI.py:174,195-214 on a slab-decomposed grid.  small_volume=1e-6 enables the reference's
    `volume < 1e-6` exit (FK_DTI.py:203-206; fused driver only)."""
    _stop_args(driver, stopping_volume, small_volume, record_steps)
    model, fields, statics, scalars = _dti_args(u0, rgb, Dw, rho, stopping_volume, small_volume)
    return _run_cuda(model, fields, statics, scalars, np.shape(u0), (dx, dy, dz), dt, nsteps, halo, precision, device, ("u",),
                     driver, record_steps)


def session_dti(u0, rgb, Dw, rho, dt, dx, dy, dz, halo=1, precision="fp64", device=0, driver="fused"):
    """A reusable decomposed FK_DTI problem (bench.py: reset / run / sync_ms repeatedly)."""
    model, fields, statics, scalars = _dti_args(u0, rgb, Dw, rho, np.inf, 0.0)
    return SlabSession(model, fields, statics, scalars, np.shape(u0), (dx, dy, dz), dt, halo, precision, device, ("u",), driver)


def session_fk2c(P0, N0, S0, WM, GM, th, th_necro, Dw, ratio, Ds, rho, lambda_np, sigma_np, lambda_s, dt, dx, dy, dz,
                 halo=1, precision="fp64", device=0, driver="fused"):
    """A reusable decomposed FK_2c problem (bench.py)."""
    from . import _native as nat
    return SlabSession(nat.MODEL_FK_2C, [(nat.FIELD_P, P0), (nat.FIELD_N, N0), (nat.FIELD_S, S0)],
                       [(nat.FIELD_WM, WM), (nat.FIELD_GM, GM)],
                       _scalars(rho=rho, Dw=Dw, ratio=ratio, th_matter=th, th_necro=th_necro, Ds=Ds, lambda_np=lambda_np,
                                sigma_np=sigma_np, lambda_s=lambda_s), np.shape(P0), (dx, dy, dz), dt, halo, precision, device,
                       ("P", "N", "S"), driver)


def fk2c_loop(P0, N0, S0, WM, GM, th, th_necro, Dw, ratio, Ds, rho, lambda_np, sigma_np, lambda_s, dt, dx, dy, dz,
              nsteps, halo=1, precision="fp64", device=0, driver="fused", stopping_volume=np.inf, record_steps=None):
    """FK_2c.py:215-216,227-245 on a slab-decomposed grid: P, N and S boundary planes are exchanged."""
    from . import _native as nat
    _stop_args(driver, stopping_volume, 0.0, record_steps)
    return _run_cuda(nat.MODEL_FK_2C, [(nat.FIELD_P, P0), (nat.FIELD_N, N0), (nat.FIELD_S, S0)],
                     [(nat.FIELD_WM, WM), (nat.FIELD_GM, GM)],
                     _scalars(stopping_volume, rho=rho, Dw=Dw, ratio=ratio, th_matter=th, th_necro=th_necro, Ds=Ds,
                              lambda_np=lambda_np, sigma_np=sigma_np, lambda_s=lambda_s), np.shape(P0), (dx, dy, dz), dt,
                     nsteps, halo, precision, device, ("P", "N", "S"), driver, record_steps)
