"""DTI-driven Fisher-Kolmogorov solver on the B200.

Same contract as the reference's TumorGrowthToolkit/FK_DTI/FK_DTI.py (class FK_DTI_Solver,
:21-237).  As in the reference, only the tensor DIAGONAL drives the diffusion and D_plus
aliases D_minus (:39-45), i.e. the operator is sum_a D_aa(c)*(u[c-1]-2u[c]+u[c+1])/h_a^2.
Optional extra parameters: precision ('fp64' | 'fp32'), device, and
dti_full_tensor (default False): opt-in 19-point operator that also uses the off-diagonal tensor
components (mixed derivatives).  The packaged reference has no such mode, so it has no
reference parity; with zero off-diagonals it equals the reference operator bit for bit.
"""
import numpy as np

from .. import _host, _loop
from ..FK.FK import Solver as FK_Solver, _full_state, _full_series, _loop_stats
from . import tools


class FK_DTI_Solver(FK_Solver):
    def __init__(self, params):
        super().__init__(params)

    def get_D_from_DTI(self, dtiRGB, DwMax, maskOut=None):
        """Per-axis coefficients from the normalised tensor diagonal (FK_DTI.py:25-47).
        Like the reference, D_plus_a IS D_minus_a (same object) and maskOut zeroes dtiRGB in place."""
        if maskOut is not None:
            dtiRGB[maskOut] = 0
        out = {}
        for a, name in enumerate("xyz"):
            out["D_minus_" + name] = dtiRGB[:, :, :, a] * DwMax
        for name in "xyz":
            out["D_plus_" + name] = out["D_minus_" + name]
        return out

    def crop_tissues_and_tumor(self, tissue, tumor_initial, brainmask, margin=2, threshold=0.0):
        """Crop to the bounding box of brainmask > threshold plus margin (FK_DTI.py:49-74)."""
        box = _host.bounding_box(brainmask > threshold, margin, brainmask.shape)
        return _host.crop(tissue, box), _host.crop(tumor_initial, box), box

    def solve(self, doPlot=False):
        p = self.params
        stopping_time = p.get('stopping_time', 100)
        stopping_volume = p.get('stopping_volume', np.inf)
        Dw, rho = p['Dw'], p['rho']
        scaling = p['diffusionEllipsoidScaling']
        print(f'diffusionEllipsoidScaling: {scaling}')                 # FK_DTI.py:86 (unconditional)
        exponent = p.get('diffusionTensorExponent', 1)
        linear = p.get('diffusionTensorLinear', 0)
        pct = (p['NxT1_pct'], p['NyT1_pct'], p['NzT1_pct'])
        res_factor = p['resolution_factor']
        spacing = (p.get('dx_mm', 1.), p.get('dy_mm', 1.), p.get('dz_mm', 1.))
        init_scale = p.get('init_scale', 1.)
        n_series = p.get('time_series_solution_Nt', None)
        verbose = p.get('verbose', False)

        tensors = p["diffusionTensors"]
        if scaling != 1:                                                # FK_DTI.py:107-110
            # 'cuda' (default): csrc/tgtk_elongate.cuh, equal to the reference's float32 LAPACK path at float32
            # rounding (1e-6 of the largest entry) and ~180x faster on a full volume; 'torch': opt-out, the
            # reference's own float32 CPU eigh restated literally (bit-identical preprocessing, needs torch)
            backend = p.get('elongation_backend', 'cuda')
            if backend == 'cuda':
                tensors = tools.elongate_tensor_along_main_axis_cuda(tensors, scaling, device=self._device())
            elif backend == 'torch':
                tensors = tools.elongate_tensor_along_main_axis_torch(tensors, scaling)
            else:
                raise ValueError(f"elongation_backend must be 'torch' or 'cuda', got {backend!r}")
        rgb = tools.makeXYZ_rgb_from_tensor_cuda(tensors, exponent=exponent, linear=linear, device=self._device())
        full_tensor = bool(p.get('dti_full_tensor', False))
        if full_tensor and (exponent != 1 or linear != 0):
            raise ValueError("dti_full_tensor needs diffusionTensorExponent=1 and diffusionTensorLinear=0")

        assert isinstance(rgb, np.ndarray), "sRGB must be a numpy array"
        assert rgb.ndim == 4, "sRGB must be a 4D numpy array, with the last dimension being 3 (RGB)"
        _host.check_seed(pct)

        rgb_lo = _host.resample_linear(rgb, [res_factor, res_factor, res_factor, 1], device=self._device())
        if doPlot:          # FK_DTI.py:128-150 shows three matplotlib figures and carries on; plotting is out of
            import warnings  # scope here, so the request is ignored rather than failing the solve
            warnings.warn("doPlot=True is ignored: plotting is not part of the time-stepping path", stacklevel=2)
        grid = _host.LowResGrid(rgb_lo.shape, rgb.shape, res_factor, *spacing, pct)
        dx, dy, dz = grid.h

        # FK_DTI.py:156: the max runs over the FULL-resolution rgb
        nt = np.max([stopping_time * Dw * np.max(rgb) / np.power(np.min([dx, dy, dz]), 2) * 8 + 100,
                     stopping_time * rho * 1.1])
        nsteps, dt = _host.steps_and_dt(nt, stopping_time)
        if verbose:
            print(f'Number of simulation timesteps: {nsteps}')

        initial = _host.seed_gaussian(grid, init_scale)
        print("init: ", initial.shape, "Volume of init Tumor", np.sum(initial))   # FK_DTI.py:164
        brainmask = _host.any_positive_last(rgb_lo)                                # FK_DTI.py:169: np.max(rgb, axis=-1) > 0
        box = _host.bounding_box(brainmask > 0.5, 2, brainmask.shape)             # FK_DTI.py:171
        record = _host.record_schedule(nsteps, n_series)

        result = {}
        try:
            result['success'] = False
            if not brainmask[grid.seed]:
                raise ValueError("Origin not within brainmask")
            if full_tensor:
                off_lo = _host.resample_linear(tools.offdiagonal_from_tensor(tensors),
                                               [res_factor, res_factor, res_factor, 1], device=self._device())
                run = _loop.dti_full_loop(_host.crop(initial, box), _host.crop(rgb_lo, box), _host.crop(off_lo, box),
                                          Dw, rho, dt, dx, dy, dz, nsteps, stopping_volume=stopping_volume,
                                          record_steps=record, precision=p.get('precision', 'fp64'),
                                          device=self._device(), output=grid.output_spec(box))
            else:
                run = _loop.dti_loop(_host.crop(initial, box), _host.crop(rgb_lo, box), Dw, rho, dt, dx, dy, dz, nsteps,
                                     stopping_volume=stopping_volume, record_steps=record,
                                     precision=p.get('precision', 'fp64'), device=self._device(), devices=p.get('devices'),
                                     output=grid.output_spec(box))
            if run['stop'] == 'small':
                print("Volume is to small")                             # FK_DTI.py:203-206 (sic)
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
            result['success'] = True        # also after the "too small" exit, as in FK_DTI.py:229
            self.last_loop_stats = _loop_stats(run)
        except Exception as e:              # FK_DTI.py:232-235
            print(e)
            result['error'] = str(e)
            result['success'] = False
        return result
