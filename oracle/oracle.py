"""ctypes front end of the CPU oracle (oracle/tgtk_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of tgtk_oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the product
package never imports this module.

Parity status: pinned against reference-generated vectors in tests/golden/
(tests/test_oracle_golden.py).

The loop drivers restate the three reference time loops:
  FK      TumorGrowthToolkit/FK/FK.py:212-226
  FK_2c   TumorGrowthToolkit/FK_2c/FK_2c.py:227-245
  FK_DTI  TumorGrowthToolkit/FK_DTI/FK_DTI.py:195-214
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtgtk_oracle.so")
_lib = None

_D = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile the C restatement (gcc only; seconds)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "tgtk_oracle.c"))):
        subprocess.run(["make", "-C", _HERE, "-B", "libtgtk_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.tgo_sum.restype = ctypes.c_double
        _lib.tgo_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_D)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return int(lib().tgo_max_threads())


def set_threads(n):
    lib().tgo_set_threads(ctypes.c_int(int(n)))


def faces(WM, GM, th, Dcoef, ratio, P=None, N=None, th_necro=0.0):
    """D_minus_{x,y,z} of get_D / get_D_with_necro (FK.py:22-26, FK_2c.py:36-40)."""
    WM, GM = _c(WM), _c(GM)
    X, Y, Z = WM.shape
    out = [np.empty_like(WM) for _ in range(3)]
    Pc = _c(P) if P is not None else None
    Nc = _c(N) if N is not None else None
    lib().tgo_faces(_p(WM), _p(GM), _p(Pc), _p(Nc), ctypes.c_double(th), ctypes.c_double(th_necro),
                    ctypes.c_double(Dcoef), ctypes.c_double(ratio), X, Y, Z, *[_p(o) for o in out])
    return out


def roll_plus(a, axis):
    a = _c(a)
    X, Y, Z = a.shape
    out = np.empty_like(a)
    lib().tgo_roll_plus(_p(a), int(axis), X, Y, Z, _p(out))
    return out


def get_D(WM, GM, th, Dw, ratio, P=None, N=None, th_necro=0.0):
    """The reference's six-entry coefficient dict (FK.py:22-32 / FK_2c.py:36-46)."""
    m = faces(WM, GM, th, Dw, ratio, P, N, th_necro)
    d = {}
    for a, name in enumerate("xyz"):
        d["D_minus_" + name] = m[a]
        d["D_plus_" + name] = roll_plus(m[a], a)
    return d


def fk_step(u, D, rho, dt, dx, dy, dz):
    """FK_update (FK.py:34-42) with an explicit coefficient dict; returns a new array."""
    u = _c(u)
    X, Y, Z = u.shape
    arrs = [_c(D[k]) for k in ("D_plus_x", "D_plus_y", "D_plus_z", "D_minus_x", "D_minus_y", "D_minus_z")]
    out = np.empty_like(u)
    lib().tgo_fk_step(_p(u), *[_p(a) for a in arrs], ctypes.c_double(rho), ctypes.c_double(dt),
                      ctypes.c_double(dx), ctypes.c_double(dy), ctypes.c_double(dz), X, Y, Z, _p(out))
    return out


def fk_step_faces(u, Dm, rho, dt, dx, dy, dz):
    u = _c(u)
    X, Y, Z = u.shape
    out = np.empty_like(u)
    lib().tgo_fk_step_faces(_p(u), _p(Dm[0]), _p(Dm[1]), _p(Dm[2]), ctypes.c_double(rho),
                            ctypes.c_double(dt), ctypes.c_double(dx), ctypes.c_double(dy),
                            ctypes.c_double(dz), X, Y, Z, _p(out))
    return out


def fk2c_step(P, N, S, D, Ds, rho, lambda_np, sigma_np, lambda_s, dt, dx, dy, dz):
    """FK_2c_update (FK_2c.py:48-73) with explicit coefficient dicts; returns new arrays."""
    P, N, S = _c(P), _c(N), _c(S)
    X, Y, Z = P.shape
    keys = ("D_plus_x", "D_plus_y", "D_plus_z", "D_minus_x", "D_minus_y", "D_minus_z")
    da = [_c(D[k]) for k in keys]
    ea = [_c(Ds[k]) for k in keys]
    Po, No, So = np.empty_like(P), np.empty_like(P), np.empty_like(P)
    lib().tgo_fk2c_step(_p(P), _p(N), _p(S), *[_p(a) for a in da], *[_p(a) for a in ea],
                        ctypes.c_double(rho), ctypes.c_double(lambda_np), ctypes.c_double(sigma_np),
                        ctypes.c_double(lambda_s), ctypes.c_double(dt), ctypes.c_double(dx),
                        ctypes.c_double(dy), ctypes.c_double(dz), X, Y, Z, _p(Po), _p(No), _p(So))
    return Po, No, So


class _Fk2cTissueStepper:
    """Holds the scratch face arrays of tgo_fk2c_step_tissue across steps."""

    def __init__(self, WM, GM, th, th_necro, Dw, ratio, Ds):
        self.WM, self.GM = _c(WM), _c(GM)
        self.sc = (th, th_necro, Dw, ratio, Ds)
        self.scratch = [np.empty_like(self.WM) for _ in range(6)]
        self.first = True

    def step(self, P, N, S, rho, lambda_np, sigma_np, lambda_s, dt, dx, dy, dz):
        X, Y, Z = P.shape
        th, th_necro, Dw, ratio, Ds = self.sc
        Po, No, So = np.empty_like(P), np.empty_like(P), np.empty_like(P)
        lib().tgo_fk2c_step_tissue(
            _p(P), _p(N), _p(S), _p(self.WM), _p(self.GM), ctypes.c_double(th), ctypes.c_double(th_necro),
            ctypes.c_double(Dw), ctypes.c_double(ratio), ctypes.c_double(Ds), ctypes.c_double(rho),
            ctypes.c_double(lambda_np), ctypes.c_double(sigma_np), ctypes.c_double(lambda_s),
            ctypes.c_double(dt), ctypes.c_double(dx), ctypes.c_double(dy), ctypes.c_double(dz), X, Y, Z,
            *[_p(s) for s in self.scratch], ctypes.c_int(1 if self.first else 0), _p(Po), _p(No), _p(So))
        self.first = False
        return Po, No, So


def zoom_linear(a, out_shape):
    a = _c(a)
    out = np.empty(tuple(int(s) for s in out_shape), dtype=np.float64)
    lib().tgo_zoom_linear(_p(a), *a.shape, _p(out), *out.shape)
    return out


# ---------------------------------------------------------------------------------------
# Time loops.  Each returns a dict:
#   state (final cropped field(s)), volume (last computed), steps_run (number of updates
#   executed), stop ('volume' | 'time' | 'small'), snapshots (list, or None)
# ---------------------------------------------------------------------------------------

def _record_set(record_steps):
    return None if record_steps is None else set(int(t) for t in record_steps)


def fk_loop(u0, WM, GM, th, Dw, ratio, rho, dt, dx, dy, dz, nsteps, stopping_volume=np.inf,
            record_steps=None):
    """FK.py:200,212-226."""
    Dm = faces(WM, GM, th, Dw, ratio)
    u = _c(u0).copy()
    rec, snaps = _record_set(record_steps), ([] if record_steps is not None else None)
    volume, stop, t_stop, steps_run = None, "time", None, 0
    for t in range(int(nsteps)):
        u = fk_step_faces(u, Dm, rho, dt, dx, dy, dz)
        steps_run = t + 1
        volume = dx * dy * dz * float(np.sum(u))
        if volume >= stopping_volume:
            stop, t_stop = "volume", t
            break
        if rec is not None and t in rec:
            snaps.append(u.copy())
    return dict(state=u, volume=volume, steps_run=steps_run, stop=stop, t_stop=t_stop, snapshots=snaps)


def dti_loop(u0, rgb, Dw, rho, dt, dx, dy, dz, nsteps, stopping_volume=np.inf, record_steps=None):
    """FK_DTI.py:174,195-214 (D_plus aliases D_minus = rgb[...,a]*Dw, FK_DTI.py:39-45)."""
    D = {}
    for a, name in enumerate("xyz"):
        D["D_minus_" + name] = _c(rgb[..., a]) * Dw
        D["D_plus_" + name] = D["D_minus_" + name]
    u = _c(u0).copy()
    rec, snaps = _record_set(record_steps), ([] if record_steps is not None else None)
    volume, stop, t_stop, steps_run = None, "time", None, 0
    for t in range(int(nsteps)):
        u = fk_step(u, D, rho, dt, dx, dy, dz)
        steps_run = t + 1
        volume = dx * dy * dz * float(np.sum(u))
        if volume >= stopping_volume:
            stop, t_stop = "volume", t
            break
        if volume < 0.000001:
            stop = "small"
            break
        if rec is not None and t in rec:
            snaps.append(u.copy())
    return dict(state=u, volume=volume, steps_run=steps_run, stop=stop, t_stop=t_stop, snapshots=snaps)


def fk2c_loop(P0, N0, S0, WM, GM, th, th_necro, Dw, ratio, Ds, rho, lambda_np, sigma_np, lambda_s,
              dt, dx, dy, dz, nsteps, stopping_volume=np.inf, record_steps=None):
    """FK_2c.py:215-216,227-245 (volume = dx*dy*dz*sum(P) + sum(N), sic, FK_2c.py:235)."""
    st = _Fk2cTissueStepper(WM, GM, th, th_necro, Dw, ratio, Ds)
    P, N, S = _c(P0).copy(), _c(N0).copy(), _c(S0).copy()
    rec = _record_set(record_steps)
    snaps = ({"P": [], "N": [], "S": []} if record_steps is not None else None)
    volume, stop, t_stop, steps_run = None, "time", None, 0
    for t in range(int(nsteps)):
        P, N, S = st.step(P, N, S, rho, lambda_np, sigma_np, lambda_s, dt, dx, dy, dz)
        steps_run = t + 1
        volume = dx * dy * dz * float(np.sum(P)) + float(np.sum(N))
        if volume >= stopping_volume:
            stop, t_stop = "volume", t
            break
        if rec is not None and t in rec:
            snaps["P"].append(P.copy())
            snaps["N"].append(N.copy())
            snaps["S"].append(S.copy())
    return dict(state={"P": P, "N": N, "S": S}, volume=volume, steps_run=steps_run, stop=stop,
                t_stop=t_stop, snapshots=snaps)
