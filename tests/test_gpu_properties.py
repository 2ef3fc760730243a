"""GPU suite: size-independent properties of the hot path at sizes the oracle does not reach
(2 Mi particles): determinism, conservation, invariance under relabelling and periodic shifts."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_system(r, v, box, h=2.0):
    from pyticles_b200 import forces, neighbour_list, particles
    n = r.shape[0]
    p = particles.SmoothParticleSystem(n, d=3, maxn=n, xmax=box[0], ymax=box[1], zmax=box[2], hshort=h)
    p.r[0:n, :] = r
    p.v[0:n, :] = v
    nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=0.0)
    f = forces.SpamForce(p, nl)
    return p, nl, f


def evaluate(p, nl, f):
    from pyticles_b200 import properties
    nl.build()
    nl.separations()
    properties.spam_properties(p, nl)
    p.vdot[:, :] = 0.0
    p.udot[:] = 0.0
    f.apply()
    torch.cuda.synchronize()


@pytest.fixture(scope="module")
def big():
    sys.path.insert(0, ROOT)
    import bench
    dev = torch.device("cuda", 0)
    r, v = bench.lattice_on_device((128, 128, 128), 0, dev, 77)
    return r, v, (128.0, 128.0, 128.0)


def test_bitwise_reproducible(big):
    r, v, box = big
    p, nl, f = build_system(r, v, box)
    evaluate(p, nl, f)
    rho1, vdot1, iap1 = p.rho.clone(), p.vdot.clone(), nl.iap.clone()
    evaluate(p, nl, f)
    assert torch.equal(nl.iap, iap1)
    assert torch.equal(p.rho, rho1) and torch.equal(p.vdot, vdot1)      # fixed summation order
    # pair list properties: i < j, lexicographic, no duplicates
    iap = iap1.to(torch.int64)
    assert bool((iap[:, 0] < iap[:, 1]).all())
    key = iap[:, 0] * r.shape[0] + iap[:, 1]
    assert bool((key[1:] > key[:-1]).all())
    assert nl.nip == iap.shape[0]
    # about 4/3 pi 2^3 / 2 pairs per particle on a unit-density lattice with jitter
    assert 14.0 < iap.shape[0] / r.shape[0] < 15.0


def test_momentum_and_energy_symmetry(big):
    r, v, box = big
    p, nl, f = build_system(r, v, box)
    evaluate(p, nl, f)
    # a_ij = -a_ji and no mass factor (forces.py:359-364): the accelerations sum to zero
    tot = p.vdot.sum(dim=0).abs().max().item()
    scale = p.vdot.abs().sum().item()
    assert tot < 1e-11 * scale
    # rho, p finite and positive density everywhere
    assert bool(torch.isfinite(p.rho).all()) and float(p.rho.min()) > 0.5


def test_relabelling_and_periodic_shift_invariance(big):
    r, v, box = big
    n = r.shape[0]
    p, nl, f = build_system(r, v, box)
    evaluate(p, nl, f)
    rho0, vdot0, nip0 = p.rho.clone(), p.vdot.clone(), nl.nip
    # relabel the particles: same physics, same pair count, same fields up to summation order
    g = torch.Generator(device=r.device)
    g.manual_seed(5)
    perm = torch.randperm(n, device=r.device, generator=g)
    p2, nl2, f2 = build_system(r[perm], v[perm], box)
    evaluate(p2, nl2, f2)
    assert nl2.nip == nip0
    assert float((p2.rho - rho0[perm]).abs().max()) < 1e-12
    assert float((p2.vdot - vdot0[perm]).abs().max()) < 1e-11
    # shift every particle by a lattice vector through the periodic faces
    shift = torch.tensor([37.0, 64.0, 5.0], dtype=torch.float64, device=r.device)
    L = torch.tensor(box, dtype=torch.float64, device=r.device)
    r3 = torch.remainder(r + shift, L)
    p3, nl3, f3 = build_system(r3, v, box)
    evaluate(p3, nl3, f3)
    assert abs(nl3.nip - nip0) <= 2                      # a shift re-rounds coordinates: knife-edge pairs may flip
    assert float((p3.rho - rho0).abs().max()) < 1e-9
    assert float((p3.vdot - vdot0).abs().max()) < 1e-8


def test_bench_emits_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "small", "--steps", "3",
                          "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["roofline"]["frac"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0


def test_bspana_style_front_end_runs(tmp_path):
    """The text of run_scripts/bspana.py:38-62 with only the import lines changed (examples/bspana.py):
    SpamComplete with its DEFAULT arguments (cgrad = 1, eta = 1, zeta = 0.1) next to the collision force, build,
    properties, update loop with thermostat, trajectory output."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "bspana.py"), "12", "8"], cwd=str(tmp_path),
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Completed 11 steps" in out.stdout and "nan" not in out.stdout          # the script prints the last index
    from scipy.io import netcdf_file
    f = netcdf_file(str(tmp_path / "output.nc"), "r", mmap=False)
    assert f.variables["position"].shape == (2, 512, 3)
    f.close()


def test_create_particle_thermostat_and_euler():
    from pyticles_b200 import forces, neighbour_list, particles, properties
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    r, v, box = O.lattice_workload(6, 6, 6, seed=3, jitter=0.2)
    n = r.shape[0]
    particles.SPROPS = True
    try:
        p = particles.SmoothParticleSystem(n, d=3, maxn=n + 4, xmax=box[0], ymax=box[1], zmax=box[2], hshort=2.0,
                                           integrator='euler', thermostat=True, thermostat_temp=2.0)
        p.r[0:n, :] = r
        p.v[0:n, :] = v
        nl = neighbour_list.VerletList(p, cutoff=2.0, tolerance=1.0)
        p.nlists.append(nl)
        p.nl_default = nl
        p.forces.append(forces.SpamForce(p, nl))
        nl.build()
        nip0 = nl.nip
        p.create_particle((3.0, 3.0, 3.0))                       # particles.py:171-178: grows n, rebuilds
        assert p.n == n + 1 and nl.nip > nip0
        nl.separations()
        properties.spam_properties(p, nl)
        r0, v0 = p.r[:p.n].clone(), p.v[:p.n].clone()
        p.update(0.01)
        vd = p.vdot[:p.n].clone()
        # euler (integrator.py:14-41): x += xdot * dt with the derivatives of the START state; update()
        # evaluates them once before the step (particles.py:479) and the step evaluates them again
        assert float((p.r[:p.n] - (r0 + 0.01 * v0)).abs().max()) < 1e-13
        assert float(p.t.mean()) == pytest.approx(2.0, rel=1e-12)  # scaling thermostat (particles.py:450-457)
        assert bool(torch.isfinite(vd).all())
    finally:
        particles.SPROPS = False
