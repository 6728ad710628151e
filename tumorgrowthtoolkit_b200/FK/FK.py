"""Single-species Fisher-Kolmogorov solver on the B200.

Same constructor / solve() / result-dict contract as the reference's
TumorGrowthToolkit/FK/FK.py (class Solver, :6-250); the time loop (:212-226) and the per-step
methods run in libtgtk_b200's CUDA kernels.  Extra optional parameters (all default to the
reference's behaviour):

    precision   'fp64' (default) | 'fp32'   storage + arithmetic type of the loop
    device      CUDA device ordinal (default 0)
    devices     'slab': decompose the grid along axis 0 over the ranks of the caller's torch.distributed group
                (one process per GPU; every rank calls solve() and receives the same result dict)
    bit_exact   True: fp64 results equal numpy's bit for bit (slower face-array kernel)
"""
import numpy as np

from .. import _host, _loop, _ops
from ..base_solver import BaseSolver


def _full_state(run, grid, box, device, key=None):
    """The loop's final state on the input grid (FK.py:229-238): resampled on the device by the loop itself when it could
    (`state_full`), otherwise -- slab-decomposed runs, a test double standing in for the loop in the host-logic tests --
    restore + zoom of the cropped low-resolution array."""
    if run.get('state_full') is not None:
        return run['state_full'] if key is None else run['state_full'][key]
    return grid.upsample_crop(run['state'] if key is None else run['state'][key], box, device=device)


def _full_series(run, grid, box, device, key=None):
    """The recorded snapshots on the input grid, as the (n, X, Y, Z) array FK.py:238 builds."""
    if run.get('snapshots_full') is not None:
        arr = run['snapshots_full'] if key is None else run['snapshots_full'][key]
        return arr if len(arr) else np.array([])
    snaps = run['snapshots'] if key is None else run['snapshots'][key]
    return np.array([grid.upsample_crop(s, box, device=device) for s in snaps])


def _loop_stats(run):
    """Not part of the reference's result dict: kept on the solver instance for benchmarks."""
    st = run['state'] if run.get('state') is not None else None
    if isinstance(st, dict):
        st = st['P']
    voxels = run.get('voxels') or (int(np.prod(st.shape)) if st is not None else None)
    return {'steps_run': run['steps_run'], 'loop_ms': run.get('loop_ms'), 'voxels': voxels}


class Solver(BaseSolver):
    def __init__(self, params):
        super().__init__(params)

    # ---- per-step operators (public in the reference; kept callable) ------------------------
    def m_Tildas(self, WM, GM, th):
        """Face-averaged tissue maps, FK.py:10-20 (device kernel, reference rounding)."""
        return _ops.tissue_tildas(WM, GM, th, device=self._device())

    def get_D(self, WM, GM, th, Dw, Dw_ratio):
        """Six face-coefficient arrays, FK.py:22-32."""
        return _ops.face_coefficients(WM, GM, th, Dw, Dw_ratio, device=self._device())

    def FK_update(self, A, D_domain, f, dt, dx, dy, dz):
        """One explicit Euler step, in place, FK.py:34-42.  Returns the same array object."""
        return _ops.fk_update(A, D_domain, f, dt, dx, dy, dz, device=self._device())

    # ---- set-up helpers ---------------------------------------------------------------------
    def crop_tissues_and_tumor(self, GM, WM, tumor_initial, margin=2, threshold=0.05):
        """Crop to the bounding box of GM+WM >= threshold plus margin (FK.py:44-73)."""
        box = _host.bounding_box(GM + WM >= threshold, margin, GM.shape)
        return _host.crop(GM, box), _host.crop(WM, box), _host.crop(tumor_initial, box), box

    def restore_tumor(self, original_shape, tumor, crop_coords):
        """Paste a cropped field back into a zero volume (FK.py:75-90)."""
        return _host.paste(original_shape, tumor, crop_coords)

    def gauss_sol3d(self, x, y, z, dx, dy, dz, init_scale):
        """Clipped Gaussian seed on caller-given coordinate arrays (FK.py:92-105)."""
        r2 = np.power(x * dx / init_scale, 2) + np.power(y * dy / init_scale, 2) + np.power(z * dz / init_scale, 2)
        g = 250 / np.power(4 * np.pi * 5.0, 3 / 2) * np.exp(-r2 / (4 * 5.0))
        g = np.where(g > 0.1, g, 0)
        return np.where(g > 1, np.float64(1), g)

    def get_initial_configuration(self, Nx, Ny, Nz, NxT, NyT, NzT, r):
        """Box-shaped seed (FK.py:107-114; unused by solve(), kept for API parity)."""
        A = np.zeros([Nx, Ny, Nz])
        if r == 0:
            A[NxT, NyT, NzT] = 1
        else:
            A[NxT - r:NxT + r, NyT - r:NyT + r, NzT - r:NzT + r] = 1
        return A

    def _device(self):
        p = self.params if isinstance(self.params, dict) else {}
        if 'device' in p:
            return int(p['device'])
        if p.get('devices') == 'slab':          # one process per GPU (torchrun): each rank steps its slab on its own device
            import os
            return int(os.environ.get('LOCAL_RANK', 0))
        return 0

    # ---- solve ------------------------------------------------------------------------------
    def solve(self):
        p = self.params
        stopping_time = p.get('stopping_time', 100)
        stopping_volume = p.get('stopping_volume', np.inf)
        Dw, rho = p['Dw'], p['rho']
        gm, wm = p['gm'], p['wm']
        pct = (p['NxT1_pct'], p['NyT1_pct'], p['NzT1_pct'])
        res_factor = p['resolution_factor']
        ratio = p.get('RatioDw_Dg', 10.)
        th_matter = p.get('th_matter', 0.1)
        spacing = (p.get('dx_mm', 1.), p.get('dy_mm', 1.), p.get('dz_mm', 1.))
        init_scale = p.get('init_scale', 1.)
        n_series = p.get('time_series_solution_Nt', None)
        verbose = p.get('verbose', False)

        _host.check_tissue_inputs(gm, wm, pct)

        gm_lo = _host.resample_linear(gm, res_factor, device=self._device())
        wm_lo = _host.resample_linear(wm, res_factor, device=self._device())
        grid = _host.LowResGrid(gm_lo.shape, gm.shape, res_factor, *spacing, pct)
        dx, dy, dz = grid.h

        result = {}
        if gm_lo[grid.seed] == 0 and wm_lo[grid.seed] == 0:          # FK.py:177-182
            result['error'] = 'Initial tumor position is outside the brain matter'
            result['success'] = False
            if verbose:
                print('Initial tumor position is outside the brain matter')
            return result

        # FK.py:185: Nt = max(T*Dw/min(h)^2*8 + 100, T*rho*1.1)
        nt = np.max([stopping_time * Dw / np.power(np.min([dx, dy, dz]), 2) * 8 + 100, stopping_time * rho * 1.1])
        nsteps, dt = _host.steps_and_dt(nt, stopping_time)
        if verbose:
            print(f'Number of simulation timesteps: {nsteps}')

        initial = _host.seed_gaussian(grid, init_scale)
        box = _host.bounding_box(gm_lo + wm_lo >= 0.5, 2, gm_lo.shape)   # FK.py:197
        record = _host.record_schedule(nsteps, n_series)

        try:
            run = _loop.fk_loop(_host.crop(initial, box), _host.crop(wm_lo, box), _host.crop(gm_lo, box),
                                th_matter, Dw, ratio, rho, dt, dx, dy, dz, nsteps,
                                stopping_volume=stopping_volume, record_steps=record,
                                precision=p.get('precision', 'fp64'), device=self._device(),
                                bit_exact=bool(p.get('bit_exact', False)), devices=p.get('devices'),
                                output=grid.output_spec(box))
            volume = np.float64(run['volume'])
            final_time = run['t_stop'] * dt if run['stop'] == 'volume' else stopping_time

            result['initial_state'] = grid.upsample(initial, device=self._device())
            result['final_state'] = _full_state(run, grid, box, self._device())
            result['final_time'] = final_time
            result['final_volume'] = volume
            result['stopping_criteria'] = 'volume' if volume >= stopping_volume else 'time'
            result['time_series'] = _full_series(run, grid, box, self._device()) if record is not None else None
            result['Dw'] = Dw
            result['rho'] = rho
            result['success'] = True
            result['NxT1_pct'], result['NyT1_pct'], result['NzT1_pct'] = pct
            # not part of the reference's result dict: kept on the instance for benchmarks
            self.last_loop_stats = _loop_stats(run)
        except Exception as e:                                        # FK.py:246-248
            result['error'] = str(e)
            result['success'] = False
        return result
