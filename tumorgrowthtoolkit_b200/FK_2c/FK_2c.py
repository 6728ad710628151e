"""Proliferative / necrotic / nutrient (P, N, S) solver on the B200.

Same contract as the reference's TumorGrowthToolkit/FK_2c/FK_2c.py (class Solver, :6-279).
The per-step pair FK_2c_update + get_D_with_necro (:229-230) is ONE fused CUDA kernel pass;
exceptions propagate (the reference's try/except is commented out, :226,275-277).
Optional extra parameters: precision ('fp64' | 'fp32'), device.
"""
import numpy as np

from .. import _host, _loop, _ops
from ..FK.FK import Solver as FK_Solver, _full_state, _full_series, _loop_stats


class Solver(FK_Solver):
    def __init__(self, params):
        super().__init__(params)

    # ---- per-step operators -----------------------------------------------------------------
    def m_Tildas_with_necro(self, WM, GM, PC, PN, th, th_necro):
        """FK_2c.py:10-34."""
        return _ops.tissue_tildas(WM, GM, th, PC, PN, th_necro, device=self._device())

    def get_D_with_necro(self, WM, GM, th, Dw, Dw_ratio, PC, NC, th_necro):
        """FK_2c.py:36-46."""
        return _ops.face_coefficients(WM, GM, th, Dw, Dw_ratio, PC, NC, th_necro, device=self._device())

    def FK_2c_update(self, P, N, S, D_domain, f, lambda_np, sigma_np, D_s_domain, lambda_s, dt, dx, dy, dz):
        """One step of the coupled system, in place, FK_2c.py:48-73."""
        return _ops.fk2c_update(P, N, S, D_domain, f, lambda_np, sigma_np, D_s_domain, lambda_s, dt, dx, dy, dz,
                                device=self._device())

    def smooth_heaviside(self, x, sigma_np, k=50):
        """FK_2c.py:75-76."""
        return 1 - 1 / (1 + np.exp(-k * (x - sigma_np)))

    def crop_tissues_and_tumor(self, GM, WM, tumor_initial, necrotic_initial, nutrient_field, margin=2,
                               threshold=0.1):
        """FK_2c.py:79-118."""
        box = _host.bounding_box(GM + WM >= threshold, margin, GM.shape)
        states = {'P': _host.crop(tumor_initial, box), 'N': _host.crop(necrotic_initial, box),
                  'S': _host.crop(nutrient_field, box)}
        return _host.crop(GM, box), _host.crop(WM, box), states, box

    def get_initial_configuration(self, NxT, NyT, NzT, Nx, Ny, Nz, dx, dy, dz, init_scale, GM, WM, th_matter):
        """P = clipped Gaussian, N = 0, S = 1 inside tissue (FK_2c.py:121-131)."""
        xv, yv, zv = np.meshgrid(np.arange(0, Nx), np.arange(0, Ny), np.arange(0, Nz), indexing='ij')
        P = np.array(self.gauss_sol3d(xv - NxT, yv - NyT, zv - NzT, dx, dy, dz, init_scale))
        return {'P': P, 'N': np.zeros(P.shape), 'S': np.where(WM + GM >= th_matter, np.ones(P.shape), 0)}

    # ---- solve ------------------------------------------------------------------------------
    def solve(self):
        p = self.params
        stopping_time = p.get('stopping_time', 100)
        stopping_volume = p.get('stopping_volume', np.inf)
        Dw, rho = p['Dw'], p['rho']
        lambda_np, sigma_np, D_s, lambda_s = p['lambda_np'], p['sigma_np'], p['D_s'], p['lambda_s']
        gm, wm = p['gm'], p['wm']
        pct = (p['NxT1_pct'], p['NyT1_pct'], p['NzT1_pct'])
        res_factor = p['resolution_factor']
        ratio = p.get('RatioDw_Dg', 10.)
        th_matter = p.get('th_matter', 0.1)
        th_necro = p.get('th_necro', 0.9)
        spacing = (p.get('dx_mm', 1.), p.get('dy_mm', 1.), p.get('dz_mm', 1.))
        init_scale = p.get('init_scale', 1.)
        n_series = p.get('time_series_solution_Nt', None)
        nt_multiplier = p.get('Nt_multiplier', 8)
        verbose = p.get('verbose', False)

        _host.check_tissue_inputs(gm, wm, pct)

        gm_lo = _host.resample_linear(gm, res_factor, device=self._device())
        wm_lo = _host.resample_linear(wm, res_factor, device=self._device())
        grid = _host.LowResGrid(gm_lo.shape, gm.shape, res_factor, *spacing, pct)
        dx, dy, dz = grid.h

        # FK_2c.py:200: Nt = T*max(Dw,D_s)/min(h)^2*Nt_multiplier + 300
        nt = stopping_time * np.max([Dw, D_s]) / np.power(np.min([dx, dy, dz]), 2) * nt_multiplier + 300
        nsteps, dt = _host.steps_and_dt(nt, stopping_time)
        if verbose:
            print(f'Number of simulation timesteps: {nsteps}')

        initial = {'P': _host.seed_gaussian(grid, init_scale), 'N': np.zeros(grid.shape),
                   'S': np.where(wm_lo + gm_lo >= th_matter, np.ones(grid.shape), 0)}   # FK_2c.py:121-131
        box = _host.bounding_box(gm_lo + wm_lo >= 0.5, 2, gm_lo.shape)                  # FK_2c.py:213
        record = _host.record_schedule(nsteps, n_series)

        run = _loop.fk2c_loop(_host.crop(initial['P'], box), _host.crop(initial['N'], box),
                              _host.crop(initial['S'], box), _host.crop(wm_lo, box), _host.crop(gm_lo, box),
                              th_matter, th_necro, Dw, ratio, D_s, rho, lambda_np, sigma_np, lambda_s,
                              dt, dx, dy, dz, nsteps, stopping_volume=stopping_volume, record_steps=record,
                              precision=p.get('precision', 'fp64'), device=self._device(), devices=p.get('devices'),
                              output=grid.output_spec(box))
        volume = np.float64(run['volume'])
        final_time = run['t_stop'] * dt if run['stop'] == 'volume' else stopping_time

        result = {}
        result['initial_state'] = {k: grid.upsample(v, device=self._device()) for k, v in initial.items()}
        result['final_state'] = {k: _full_state(run, grid, box, self._device(), k) for k in 'PNS'}
        if record is not None:
            result['time_series'] = {k: list(_full_series(run, grid, box, self._device(), k)) for k in 'PNS'}
        else:
            result['time_series'] = None
        result['Dw'] = Dw
        result['rho'] = rho
        result['success'] = True
        result['NxT1_pct'], result['NyT1_pct'], result['NzT1_pct'] = pct
        result['final_time'] = final_time
        result['final_volume'] = volume
        result['stopping_criteria'] = 'volume' if volume >= stopping_volume else 'time'
        self.last_loop_stats = _loop_stats(run)
        return result
