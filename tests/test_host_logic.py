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
