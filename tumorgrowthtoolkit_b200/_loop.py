"""Device time loops: the `for t in range(N_simulation_steps)` bodies of the three reference
solvers (FK/FK.py:212-226, FK_2c/FK_2c.py:227-245, FK_DTI/FK_DTI.py:195-214) executed by
libtgtk_b200 on one GPU.

Each function uploads the cropped float64 arrays the reference would loop over (strided crop
views are passed as they are), enqueues all steps, and returns

    {'state': array | {'P','N','S'}, 'volume': float, 'steps_run': int,
     'stop': 'time' | 'volume' | 'small', 't_stop': int | None,
     'snapshots': list | {'P': [...], 'N': [...], 'S': [...]} | None,
     'loop_ms': float, 'launches': int}

The update, the per-step volume, the stop tests and the snapshot copies all run on the
device; the host only enqueues launches and reads the final status.
"""
import numpy as np

from . import _native as nat

_DTYPES = {"fp64": nat.F64, "float64": nat.F64, "f64": nat.F64, "double": nat.F64,
           "fp32": nat.F32, "float32": nat.F32, "f32": nat.F32, "single": nat.F32}


def dtype_code(precision):
    try:
        return _DTYPES[str(precision).lower()]
    except KeyError:
        raise ValueError(f"precision must be 'fp64' or 'fp32', got {precision!r}") from None


def _finish(sim, nsteps, record_steps, fields, names, output=None):
    """Runs the loop and collects its results.  `output` = (lo, virt_shape, out_shape): what solve() does with the final
    state and with every snapshot -- restore_tumor into the low-resolution grid and zoom back to the input grid,
    FK/FK.py:229-238 -- happens on the device while the simulation is still open (tgtk_sim_fetch_series); the result then
    carries 'state_full' / 'snapshots_full' instead of the cropped low-resolution arrays."""
    sim.run(nsteps, record_steps)
    st = sim.sync()
    stop = {nat.STOP_NONE: "time", nat.STOP_VOLUME: "volume", nat.STOP_SMALL: "small"}[st.stop_reason]
    res = dict(volume=(float(st.last_volume) if st.steps_run > 0 else None),
               steps_run=int(st.steps_run), stop=stop,
               t_stop=(int(st.stop_step) if st.stop_reason == nat.STOP_VOLUME else None),
               loop_ms=float(st.loop_ms), launches=int(st.kernel_launches), voxels=int(np.prod(sim.shape)))
    sp = sim.sparse_info()
    res.update(sparse_items=sp["items"], sparse_live_fraction=sp["live_voxels"] / max(sp["voxels"], 1))
    if output is not None:
        lo, virt_shape, out_shape = output
        out_shape = tuple(int(v) for v in out_shape)
        full = {n: sim.fetch_series(f, -1, 1, lo, virt_shape, np.empty(out_shape)) for n, f in zip(names, fields)}
        snaps_full = None
        if record_steps is not None:
            snaps_full = {n: sim.fetch_series(f, 0, st.snapshots, lo, virt_shape, np.empty((st.snapshots,) + out_shape))
                          for n, f in zip(names, fields)}
        res.update(state=None, snapshots=None, state_full=full, snapshots_full=snaps_full)
        if len(names) == 1:
            res["state_full"] = full[names[0]]
            if snaps_full is not None:
                res["snapshots_full"] = snaps_full[names[0]]
        return res
    state = {n: sim.download(f) for n, f in zip(names, fields)}
    snaps = None
    if record_steps is not None:
        snaps = {n: [sim.snapshot(i, f) for i in range(st.snapshots)] for n, f in zip(names, fields)}
    res.update(state=state, snapshots=snaps)
    if len(names) == 1:
        res["state"] = state[names[0]]
        if snaps is not None:
            res["snapshots"] = snaps[names[0]]
    return res


def _slab_requested(devices):
    """params['devices'] == 'slab' (or TGTK_DEVICES=slab): decompose this solve over the ranks of the caller's
    torch.distributed group (one process per GPU, every rank calls solve() with the same parameters and gets
    the same result dict).  Anything else, or no initialised group of more than one rank: one GPU."""
    import os
    if (devices or os.environ.get("TGTK_DEVICES", "")) != "slab":
        return False
    try:
        import torch.distributed as dist
    except ImportError:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def fk_loop(u0, WM, GM, th, Dw, ratio, rho, dt, dx, dy, dz, nsteps, stopping_volume=np.inf,
            record_steps=None, precision="fp64", device=0, bit_exact=False, devices=None, output=None):
    """FK: get_D (FK.py:200) fused into the step kernel + the loop of FK.py:212-226.
    bit_exact=True builds the three face arrays in the reference's rounding order and steps
    them with unfused arithmetic (fp64 results then equal numpy's bit for bit)."""
    logical_or = bool(np.asarray(WM).dtype == np.bool_ and np.asarray(GM).dtype == np.bool_)
    if _slab_requested(devices):
        if bit_exact or logical_or:
            raise ValueError("devices='slab' steps the compact FK kernel: bit_exact / boolean tissue maps are single-GPU modes")
        from . import slab
        return slab.fk_loop(u0, WM, GM, th, Dw, ratio, rho, dt, dx, dy, dz, nsteps, precision=precision, device=device,
                            stopping_volume=stopping_volume, record_steps=record_steps)
    model = nat.MODEL_FK_FACES if (bit_exact or logical_or) else nat.MODEL_FK
    p = nat.make_params(model, dtype_code(precision), np.shape(u0), (dx, dy, dz), dt, rho=rho, Dw=Dw,
                        ratio=ratio, th_matter=th, stopping_volume=stopping_volume,
                        logical_or_faces=int(logical_or))
    with nat.Sim(p, device) as sim:
        sim.upload(nat.FIELD_WM, WM)
        sim.upload(nat.FIELD_GM, GM)
        sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        return _finish(sim, nsteps, record_steps, (nat.FIELD_U,), ("u",), output)


def dti_loop(u0, rgb, Dw, rho, dt, dx, dy, dz, nsteps, stopping_volume=np.inf, record_steps=None,
             precision="fp64", device=0, devices=None, output=None):
    """FK_DTI: get_D_from_DTI (FK_DTI.py:174: D_a = rgb[...,a]*Dw, D_plus aliasing D_minus)
    + the loop of FK_DTI.py:195-214 incl. the `volume < 1e-6` exit."""
    if _slab_requested(devices):
        from . import slab
        return slab.dti_loop(u0, rgb, Dw, rho, dt, dx, dy, dz, nsteps, precision=precision, device=device,
                             stopping_volume=stopping_volume, small_volume=0.000001, record_steps=record_steps)
    p = nat.make_params(nat.MODEL_DTI, dtype_code(precision), np.shape(u0), (dx, dy, dz), dt, rho=rho, Dw=Dw,
                        stopping_volume=stopping_volume, small_volume=0.000001)
    rgb = np.asarray(rgb)
    with nat.Sim(p, device) as sim:
        for a in range(3):
            sim.upload(nat.FIELD_COEF0 + a, rgb[..., a] * Dw)
        sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        return _finish(sim, nsteps, record_steps, (nat.FIELD_U,), ("u",), output)


def dti_full_loop(u0, rgb, offdiag, Dw, rho, dt, dx, dy, dz, nsteps, stopping_volume=np.inf, record_steps=None,
                  precision="fp64", device=0, output=None):
    """Opt-in full-tensor FK_DTI (NOT in the packaged reference, SURVEY.md finding 1): the
    reference's diagonal operator plus 2*sum_{a<b} D_ab d_a d_b u on a 19-point stencil.
    rgb[...,a] = normalised D_aa as in dti_loop; offdiag[...,(xy,xz,yz)] in the same units."""
    p = nat.make_params(nat.MODEL_DTI_FULL, dtype_code(precision), np.shape(u0), (dx, dy, dz), dt, rho=rho, Dw=Dw,
                        stopping_volume=stopping_volume, small_volume=0.000001)
    rgb, offdiag = np.asarray(rgb), np.asarray(offdiag)
    with nat.Sim(p, device) as sim:
        for a in range(3):
            sim.upload(nat.FIELD_COEF0 + a, rgb[..., a] * Dw)
            sim.upload(nat.FIELD_COEF0 + 3 + a, offdiag[..., a] * Dw)
        sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        return _finish(sim, nsteps, record_steps, (nat.FIELD_U,), ("u",), output)


def fk2c_loop(P0, N0, S0, WM, GM, th, th_necro, Dw, ratio, Ds, rho, lambda_np, sigma_np, lambda_s,
              dt, dx, dy, dz, nsteps, stopping_volume=np.inf, record_steps=None, precision="fp64", device=0, devices=None,
              output=None):
    """FK_2c: get_D_with_necro + get_D(...,D_s,1) + FK_2c_update fused per step
    (FK_2c.py:215-216,229-230) + the loop of FK_2c.py:227-245."""
    if _slab_requested(devices):
        from . import slab
        return slab.fk2c_loop(P0, N0, S0, WM, GM, th, th_necro, Dw, ratio, Ds, rho, lambda_np, sigma_np, lambda_s, dt, dx, dy, dz,
                              nsteps, precision=precision, device=device, stopping_volume=stopping_volume,
                              record_steps=record_steps)
    p = nat.make_params(nat.MODEL_FK_2C, dtype_code(precision), np.shape(P0), (dx, dy, dz), dt, rho=rho, Dw=Dw,
                        ratio=ratio, th_matter=th, th_necro=th_necro, Ds=Ds, lambda_np=lambda_np,
                        sigma_np=sigma_np, lambda_s=lambda_s, stopping_volume=stopping_volume)
    with nat.Sim(p, device) as sim:
        sim.upload(nat.FIELD_WM, WM)
        sim.upload(nat.FIELD_GM, GM)
        sim.upload(nat.FIELD_P, P0)
        sim.upload(nat.FIELD_N, N0)
        sim.upload(nat.FIELD_S, S0)
        sim.prepare()
        return _finish(sim, nsteps, record_steps, (nat.FIELD_P, nat.FIELD_N, nat.FIELD_S), ("P", "N", "S"), output)
