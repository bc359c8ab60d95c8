"""CPU: the plain-C restatement against the golden vectors recorded from the compiled reference.

This is the pin that travels: it holds on any box, with or without /root/reference or oracle/_ref.
"""
import numpy as np
import pytest

import golden_check as gc
from oracle import port_backend


@pytest.mark.parametrize("name", gc.golden_cases())
def test_port_phases(name):
    gc.check_phases(port_backend.PortSim, name, tol_j=1e-13)


@pytest.mark.parametrize("name", gc.golden_cases())
def test_port_multistep(name):
    sim, g = gc.check_multistep(port_backend.PortSim, name)
    sim.deposit_moment()
    assert np.allclose(sim.get_energy(), g["end_energy"], rtol=1e-10, atol=1e-12)


def test_golden_present():
    assert len(gc.golden_cases()) >= 6
