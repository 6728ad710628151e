"""Slab decomposition host logic on the CPU: world_size-2 and -3 gloo groups run the decomposed
time loop with the CPU oracle as the local stepper (the oracle as the checker) and must
reproduce the single-domain result BITWISE -- same operator, same order, exact halos."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tumorgrowthtoolkit_b200 import slab


def test_plan_partitions_ring_and_halos():
    for X, world, halo in ((10, 2, 3), (17, 3, 2), (8, 8, 1), (9, 1, 4)):
        plans = [slab.SlabPlan(X, world, r, halo) for r in range(world)]
        assert [p.x0 for p in plans][0] == 0 and plans[-1].x1 == X
        assert sum(p.n for p in plans) == X
        for p in plans:
            idx = p.indices()
            assert len(idx) == p.n + 2 * halo
            assert list(idx[p.owned]) == list(range(p.x0, p.x1))
            assert idx[0] == (p.x0 - halo) % X and idx[-1] == (p.x1 + halo - 1) % X
            # what I send up is what the next rank holds as lower ghosts
            nxt = plans[p.next]
            assert list(idx[p.send_up]) == list(nxt.indices()[nxt.ghost_lo])
            prv = plans[p.prev]
            assert list(idx[p.send_down]) == list(prv.indices()[prv.ghost_hi])
    with pytest.raises(ValueError):
        slab.SlabPlan(4, 8, 0, 1)
    with pytest.raises(ValueError):
        slab.SlabPlan(10, 2, 0, 6)


class OracleStepper:
    """Local extended slab stepped by the CPU oracle with periodic wrap on the LOCAL array, which
    is exactly what the CUDA kernels do with a rank's slab."""

    def __init__(self, orc, kind, local, scal):
        self.orc, self.kind, self.scal = orc, kind, scal
        self.state = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in local["state"].items()}
        self.static = local["static"]

    def context(self):
        import contextlib
        return contextlib.nullcontext()

    def run(self, k):
        o, s = self.orc, self.scal
        if self.kind == "fk":
            r = o.fk_loop(self.state["u"].numpy(), self.static["wm"], self.static["gm"], s["th"], s["Dw"], s["ratio"],
                          s["rho"], s["dt"], *s["h"], k)
            self.state["u"].copy_(torch.from_numpy(r["state"]))
        else:
            r = o.fk2c_loop(self.state["P"].numpy(), self.state["N"].numpy(), self.state["S"].numpy(), self.static["wm"],
                            self.static["gm"], s["th"], 0.9, s["Dw"], s["ratio"], 1.3, s["rho"], 0.35, 0.5, 0.05, s["dt"],
                            *s["h"], k)
            for n in "PNS":
                self.state[n].copy_(torch.from_numpy(r["state"][n]))

    def tensors(self):
        return list(self.state.values())


def _problem():
    rng = np.random.default_rng(42)
    shape = (13, 6, 7)
    wm = rng.uniform(0.2, 0.7, shape)
    gm = 1.0 - wm                       # tissue everywhere: the periodic ring is fully active
    u0 = rng.uniform(0, 0.6, shape)
    scal = dict(th=0.1, Dw=1.0, ratio=10.0, rho=0.3, dt=0.02, h=(1.0, 1.2, 0.9))
    return shape, wm, gm, u0, scal


def _worker(rank, world, port, kind, halo, nsteps, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    orc.set_threads(1)
    shape, wm, gm, u0, scal = _problem()
    plan = slab.SlabPlan(shape[0], world, rank, halo)
    static = dict(wm=plan.extend(wm), gm=plan.extend(gm))
    if kind == "fk":
        state = dict(u=plan.extend(u0))
    else:
        state = dict(P=plan.extend(u0), N=plan.extend(0.3 * u0[::-1]), S=plan.extend(np.ones(shape)))
    st = OracleStepper(orc, kind, dict(state=state, static=static), scal)
    slab.run_blocks(plan, st, nsteps, dist)
    parts = {}
    for name, t in st.state.items():
        parts[name] = slab.gather_planes(plan, t[plan.owned].contiguous(), dist).numpy()
    if rank == 0:
        np.savez(out, **parts)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,kind,halo,nsteps", [(2, "fk", 3, 10), (3, "fk", 2, 7), (2, "fk2c", 2, 6)])
def test_decomposed_loop_is_bitwise_equal_to_single_domain(tmp_path, oracle, world, kind, halo, nsteps):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), kind, halo, nsteps, out), nprocs=world, join=True)
    got = np.load(out)
    shape, wm, gm, u0, s = _problem()
    if kind == "fk":
        ref = oracle.fk_loop(u0, wm, gm, s["th"], s["Dw"], s["ratio"], s["rho"], s["dt"], *s["h"], nsteps)
        assert np.array_equal(got["u"], ref["state"])
    else:
        ref = oracle.fk2c_loop(u0, 0.3 * u0[::-1], np.ones(shape), wm, gm, s["th"], 0.9, s["Dw"], s["ratio"], 1.3,
                               s["rho"], 0.35, 0.5, 0.05, s["dt"], *s["h"], nsteps)
        for n in "PNS":
            assert np.array_equal(got[n], ref["state"][n]), n


def test_single_rank_ring_closes_on_itself(oracle):
    shape, wm, gm, u0, s = _problem()
    plan = slab.SlabPlan(shape[0], 1, 0, 4)
    st = OracleStepper(oracle, "fk", dict(state=dict(u=plan.extend(u0)), static=dict(wm=plan.extend(wm), gm=plan.extend(gm))), s)
    slab.run_blocks(plan, st, 9, None)
    ref = oracle.fk_loop(u0, wm, gm, s["th"], s["Dw"], s["ratio"], s["rho"], s["dt"], *s["h"], 9)
    assert np.array_equal(st.state["u"][plan.owned].numpy(), ref["state"])
