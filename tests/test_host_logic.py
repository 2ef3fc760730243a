"""CPU suite: host-side pieces of the pyticles-shaped API that need no device -- the explicit steppers against
the reference's own integrator module (tests/golden/integrators.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch


def _rhs(x):
    return torch.stack([x[1], -torch.sin(x[0]) - 0.1 * x[1] * x[2], 0.5 * x[0] * x[1] - 0.2 * x[2]])


@pytest.mark.parametrize("name", ["euler", "imp_euler", "rk4"])
def test_steppers_match_reference(golden_dir, name):
    """integrator.euler / imp_euler / rk4 (integrator.py:14-103) through the callback protocol, bit for bit."""
    from pyticles_b200 import integrator
    g = np.load(os.path.join(golden_dir, "integrators.npz"))
    box = {"x": torch.from_numpy(g["x0"].copy()), "xdot": None}

    def calc():
        box["xdot"] = _rhs(box["x"])

    def setx(x):
        box["x"] = x.clone()

    for _ in range(int(g["steps"])):
        getattr(integrator, name)(lambda: box["x"], calc, lambda: box["xdot"], setx, float(g["dt"]))
    assert np.array_equal(box["x"].numpy(), g[name])


def test_scalar_kernel_and_eos_known_answers(golden_dir):
    """spkernel.lucy_kernel (spkernel.py:86-118) and properties.vdw / vdw_energy / vdw_temp (properties.py:38-49)
    of the product's host API against the reference's answers (tests/golden/kat.npz) and SURVEY.md section 8c."""
    from pyticles_b200 import properties, spkernel
    k = np.load(os.path.join(golden_dir, "kat.npz"))
    seen = 0
    for name in k.files:
        if not name.startswith("lucy_"):
            continue
        r, h, dim, dx0, dx1, dx2, w, g0, g1, g2 = k[name]
        dx = (dx0, dx1, dx2)[:int(dim)]
        pw, pg = spkernel.lucy_kernel(r, dx, h)
        assert pw == pytest.approx(w, rel=1e-14, abs=1e-300)
        pg = [pg] * int(dim) if not isinstance(pg, (list, tuple)) else pg       # r == 0: the reference returns a scalar 0
        assert np.allclose(pg, (g0, g1, g2)[:int(dim)], rtol=1e-14, atol=1e-300)
        seen += 1
    assert seen >= 5
    assert spkernel.lucy_kernel(0.0, (0., 0., 0.), 2.0)[0] == pytest.approx(0.2611135785101408, rel=1e-15)
    assert spkernel.lucy_kernel(2.5, (2.5, 0., 0.), 2.0) == (0, [0, 0, 0])
    assert spkernel.kernel(1.0, (1., 0., 0.), 2.0, 'lucy')[0] == pytest.approx(0.081597993284419, rel=1e-15)
    with pytest.raises(NotImplementedError):
        spkernel.kernel(1.0, (1., 0., 0.), 2.0, 'gaussian')
    t = lambda x: torch.tensor(x, dtype=torch.float64)
    p, pco = properties.vdw(t(1.0), t(1.0))
    assert (float(p), float(pco)) == tuple(k["vdw_1_1"]) == (2.0, -2.0)
    p, pco = properties.vdw(t(0.5), t(1.5))
    assert (float(p), float(pco)) == tuple(k["vdw_05_15"])
    assert float(properties.vdw_energy(t(1.0), t(5.0))) == float(k["vdw_energy_1_5"]) == 3.0
    assert float(properties.vdw_temp(t(1.0), t(3.0))) == float(k["vdw_temp_1_3"]) == 5.0


def test_host_minimum_image_rows():
    """neighbour_list.py:105-123 on rows: one shift, strict comparisons against L/2 -- including values exactly on
    +-L/2 (not shifted) and beyond 3L/2 (shifted once only)."""
    from oracle import oracle as O
    from pyticles_b200.neighbour_list import _minimum_image_rows
    L = (10.0, 6.0, 7.5)
    rng = np.random.default_rng(3)
    d = rng.uniform(-2.0, 2.0, size=(500, 3)) * np.array(L)
    d[:6] = [[5.0, 3.0, 3.75], [-5.0, -3.0, -3.75], [5.000000000000001, -3.0000000000000004, 3.7500000000000004],
             [16.0, -10.0, 12.0], [0.0, 0.0, 0.0], [-15.1, 9.1, -11.3]]
    got = _minimum_image_rows(torch.from_numpy(d), *L).numpy()
    assert np.array_equal(got, O.minimum_image(d, *L))
    assert np.array_equal(got[0], d[0]) and np.array_equal(got[1], d[1])          # exactly on the boundary: kept
    assert np.array_equal(got[3], [6.0, -4.0, 4.5])                                # one shift only


def test_particle_system_storage_matches_reference(golden_dir):
    """SmoothParticleSystem on CPU tensors: constructor defaults, integrator / box choice, timing keys and the
    [11, maxn] state and derivative mappings (particles.py:69-169,256-359,496-542) against the reference's
    (tests/golden/particle_system.npz)."""
    from pyticles_b200 import particles
    g = np.load(os.path.join(golden_dir, "particle_system.npz"))
    p = particles.SmoothParticleSystem(27, d=3, maxn=40, xmax=6.0, ymax=5.0, zmax=4.0, vmax=0.0, mass=0.1386,
                                       temperature=0.8, thermostat_temp=0.9, thermostat=True, hshort=1.5, hlong=3.0,
                                       integrator='rk4', device="cpu")
    for k in ("m", "v", "t", "u", "h", "hlr", "rho", "p", "pco", "udot", "vdot"):
        assert np.array_equal(getattr(p, k).numpy(), g["defaults_" + k]), k
    assert p.step.__name__ == str(g["step_name"]) and type(p.box).__name__ == str(g["box_type"])
    assert set(g["timing_keys"].tolist()) <= set(p.timing)
    assert p.thermostat is True and p.thermostat_temp == 0.9 and (p.n, p.maxn, p.dim) == (27, 40, 3)
    for k in ("r", "v", "rdot", "vdot", "m", "mdot", "rho", "rhodot", "p", "pco", "u", "udot"):
        getattr(p, k)[...] = torch.from_numpy(g["in_" + k])
    assert np.array_equal(p.gather_state().numpy(), g["x"])
    assert np.array_equal(p.gather_derivatives().numpy(), g["xdot"])
    p.scatter_state(torch.from_numpy(g["x2"]))
    for k in ("m", "r", "v", "rho", "p", "pco", "u"):
        assert np.array_equal(getattr(p, k).numpy(), g["scattered_" + k]), k
