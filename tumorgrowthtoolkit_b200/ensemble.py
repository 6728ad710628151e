"""Synthetic-patient ensembles of FK_2c solves, sharded over the GPUs of one box.

Reference: synthetic_gens/run_gen_FK_2c.py (16 parameter sets drawn with np.random.seed(42000),
one `python generatePatient_FK_2c.py` subprocess each, :40-61) and
synthetic_gens/generatePatient_FK_2c.py (:97-157: Solver(parameters).solve(), segmentation map).
The parameter law and the per-patient solver contract are kept; the process-per-patient
launcher is replaced by device-resident simulations:

  * the tissue maps are resampled and cropped ONCE per ensemble (they are shared by all
    patients of the reference script: same files, same resolution_factor);
  * each GPU keeps `concurrency` simulations in flight on separate CUDA streams, so the
    step kernels of several patients (about one wave of thread blocks each at this grid size)
    overlap and the per-launch latency of one patient is hidden behind the others;
  * ranks shard the parameter list with no data-path collective: patients are sorted by step
    count and dealt round-robin (longest processing time first).

FET-PET surrogate, Gibbs noise and NIfTI output of the reference script are out of scope
(SURVEY.md 2): a patient here is the final P, N, S plus the thresholded segmentation.
"""
import time

import numpy as np

from . import _host, _native as nat
from ._loop import dtype_code


def generate_random_parameters(rs=np.random):
    """The sampling law of run_gen_FK_2c.py:5-26: twelve uniform draws in dict order."""
    return {
        'Dw': rs.uniform(low=0.05, high=4), 'rho': rs.uniform(low=0.01, high=0.8),
        'lambda_np': rs.uniform(0.05, 0.7), 'sigma_np': rs.uniform(0.05, 0.7),
        'D_s': rs.uniform(0.05, 4), 'lambda_s': rs.uniform(0.01, 0.5),
        'NxT1_pct': rs.uniform(low=0.3, high=0.7), 'NyT1_pct': rs.uniform(low=0.3, high=0.7),
        'NzT1_pct': rs.uniform(low=0.3, high=0.7),
        'resolution_factor': 0.65, 'RatioDw_Dg': 10,
        'th_necro_n': rs.uniform(low=0.05, high=0.2), 'th_enhancing_p': rs.uniform(low=0.25, high=0.6),
        'th_edema_p': rs.uniform(low=0.12, high=0.20),
    }


def draw_ensemble(n, seed=42000):
    """n parameter sets exactly as run_gen_FK_2c.py:41-48 draws them (legacy global RandomState)."""
    rs = np.random.RandomState(seed)
    return [generate_random_parameters(rs) for _ in range(n)]


def create_segmentation_map(P, N, th_necro_n, th_enhancing_p, th_edema_p):
    """Labels 3 (edema) / 1 (enhancing) / 4 (necrotic) from thresholds,
    generatePatient_FK_2c.py:67-83."""
    seg = np.zeros(P.shape, dtype=int)
    seg[(P >= th_edema_p) & (P < th_enhancing_p)] = 3
    seg[P >= th_enhancing_p] = 1
    seg[N > th_necro_n] = 4
    return seg


def step_count(p, spacing=(1.0, 1.0, 1.0), stopping_time=100, nt_multiplier=10):
    """FK_2c.py:200 with generatePatient_FK_2c.py's Nt_multiplier default of 10 (:184)."""
    h = min(s / p['resolution_factor'] for s in spacing)
    nt = stopping_time * np.max([p['Dw'], p['D_s']]) / np.power(h, 2) * nt_multiplier + 300
    return int(np.ceil(nt)), nt


def shard(params, rank, world, spacing=(1.0, 1.0, 1.0)):
    """Indices of the patients rank `rank` of `world` runs: sorted by step count, dealt
    round-robin, so every rank gets the same mix of long and short solves."""
    order = sorted(range(len(params)), key=lambda i: -step_count(params[i], spacing)[0])
    return order[rank::world]


class EnsembleRunner:
    """Runs FK_2c patients that share tissue maps on one GPU.

    Parameters mirror generatePatient_FK_2c.py's argparse defaults (:160-187): th_matter 0.1,
    th_necro 0.9, Nt_multiplier 10, init_scale 1, 1 mm voxels, stopping_time 100.
    """

    def __init__(self, wm, gm, resolution_factor=0.65, device=0, precision="fp64", concurrency=4,
                 th_matter=0.1, th_necro=0.9, nt_multiplier=10, stopping_time=100, init_scale=1.0,
                 spacing=(1.0, 1.0, 1.0), full_resolution=True):
        self.device, self.precision, self.concurrency = int(device), precision, max(1, int(concurrency))
        self.th_matter, self.th_necro, self.nt_multiplier = th_matter, th_necro, nt_multiplier
        self.stopping_time, self.init_scale, self.spacing = stopping_time, init_scale, spacing
        self.res = resolution_factor
        self.full_resolution = full_resolution
        self.full_shape = wm.shape
        # shared once per ensemble: FK_2c.py:174-175,213
        self.gm_lo = _host.resample_linear(np.asarray(gm, dtype=np.float64), resolution_factor, device=self.device)
        self.wm_lo = _host.resample_linear(np.asarray(wm, dtype=np.float64), resolution_factor, device=self.device)
        self.box = _host.bounding_box(self.gm_lo + self.wm_lo >= 0.5, 2, self.gm_lo.shape)
        self.cw = np.ascontiguousarray(_host.crop(self.wm_lo, self.box))
        self.cg = np.ascontiguousarray(_host.crop(self.gm_lo, self.box))
        self.S0 = np.ascontiguousarray(_host.crop(np.where(self.wm_lo + self.gm_lo >= th_matter, 1.0, 0.0), self.box))
        self.N0 = np.zeros_like(self.S0)
        self.voxels = int(self.cw.size)
        # page-locked copies of what every patient uploads, and a ring of page-locked output sets: the device-to-host
        # copies of one patient's full-resolution P, N, S and label map then overlap the other patients' kernels
        for name in ("cw", "cg", "S0", "N0"):
            pin = nat.pinned_empty(getattr(self, name).shape)
            pin[...] = getattr(self, name)
            setattr(self, name, pin)
        self._outs, self._next_out = [], 0

    def _start(self, p):
        grid = _host.LowResGrid(self.gm_lo.shape, self.full_shape, self.res, *self.spacing,
                                (p['NxT1_pct'], p['NyT1_pct'], p['NzT1_pct']))
        dx, dy, dz = grid.h
        nt = self.stopping_time * np.max([p['Dw'], p['D_s']]) / np.power(np.min([dx, dy, dz]), 2) * self.nt_multiplier + 300
        nsteps, dt = _host.steps_and_dt(nt, self.stopping_time)
        P0 = _host.crop(_host.seed_gaussian(grid, self.init_scale), self.box)
        prm = nat.make_params(nat.MODEL_FK_2C, dtype_code(self.precision), self.cw.shape, grid.h, dt, rho=p['rho'],
                              Dw=p['Dw'], ratio=p.get('RatioDw_Dg', 10.), th_matter=self.th_matter,
                              th_necro=self.th_necro, Ds=p['D_s'], lambda_np=p['lambda_np'], sigma_np=p['sigma_np'],
                              lambda_s=p['lambda_s'])
        sim = nat.Sim(prm, self.device)
        sim.upload(nat.FIELD_WM, self.cw)
        sim.upload(nat.FIELD_GM, self.cg)
        sim.upload(nat.FIELD_P, P0)
        sim.upload(nat.FIELD_N, self.N0)
        sim.upload(nat.FIELD_S, self.S0)
        sim.prepare()
        sim.run(nsteps)          # asynchronous: the steps of several patients overlap on the GPU
        return dict(sim=sim, grid=grid, nsteps=nsteps, params=p)

    def _output_set(self, shape):
        """The next page-locked output set (P, N, S float64 + uint8 labels) of a ring of `concurrency + 1`."""
        if not self._outs:      # page-locking is slow (~0.1 s per set): all sets at the first patient, none later
            self._outs = [{k: nat.pinned_empty(shape) for k in "PNS"} | {"seg": nat.pinned_empty(shape, np.uint8)}
                          for _ in range(self.concurrency + 1)]
        self._next_out = (self._next_out + 1) % len(self._outs)
        return self._outs[self._next_out]

    def _finish(self, job, copy=True):
        sim, grid, p = job['sim'], job['grid'], job['params']
        th = (p.get('th_necro_n', 0.12), p.get('th_enhancing_p', 0.4), p.get('th_edema_p', 0.15))
        st = sim.sync()
        assert st.steps_run == job['nsteps']
        if self.full_resolution:
            # restore + zoom(order=1) of P, N, S and the label map (generatePatient_FK_2c.py:67-83,141-155) on the
            # device, four asynchronous copies into page-locked memory, one wait
            pin = self._output_set(grid.full_shape)
            for name, f in (('P', nat.FIELD_P), ('N', nat.FIELD_N), ('S', nat.FIELD_S)):
                sim.download_resampled_async(f, self.box[0], grid.shape, pin[name])
            sim.segmentation_async(self.box[0], grid.shape, *th, pin['seg'])
            sim.sync()
            # results handed out are the caller's own: copy out of the ring unless they are dropped right away
            out = {k: (pin[k].copy() if copy else pin[k]) for k in "PNS"}
            seg = pin['seg'].copy() if copy else pin['seg']
        else:
            out = {name: sim.download(f) for name, f in (('P', nat.FIELD_P), ('N', nat.FIELD_N), ('S', nat.FIELD_S))}
            seg = create_segmentation_map(out['P'], out['N'], *th)
        sim.close()
        return dict(final_state=out, segmentation=seg, final_volume=float(st.last_volume), steps=job['nsteps'],
                    loop_ms=float(st.loop_ms), params=p)

    def run(self, params_list, keep_results=True):
        """Runs every parameter set; returns (results, stats).  Results are in input order."""
        results = [None] * len(params_list)
        inflight = []
        t0 = time.perf_counter()
        updates = 0
        for i, p in enumerate(params_list):
            inflight.append((i, self._start(p)))
            if len(inflight) >= self.concurrency:
                j, job = inflight.pop(0)
                r = self._finish(job, copy=keep_results)
                updates += r['steps'] * self.voxels
                results[j] = r if keep_results else dict(final_volume=r['final_volume'], steps=r['steps'])
        for j, job in inflight:
            r = self._finish(job, copy=keep_results)
            updates += r['steps'] * self.voxels
            results[j] = r if keep_results else dict(final_volume=r['final_volume'], steps=r['steps'])
        wall = time.perf_counter() - t0
        stats = dict(patients=len(params_list), wall_s=wall, voxel_updates=updates,
                     patients_per_hour=3600.0 * len(params_list) / wall if wall > 0 else float('nan'),
                     glups=updates / wall / 1e9 if wall > 0 else float('nan'), voxels=self.voxels)
        return results, stats
