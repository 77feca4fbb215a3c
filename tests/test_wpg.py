"""Gait scheduler parity against the fixture recorded from the reference's own wpg.py
(tests/golden/make_wpg_golden.py; SURVEY.md fixture W)."""
import os

import numpy as np
import pytest

from srbd_horizon_b200 import prb as P
from srbd_horizon_b200 import wpg

GOLD = os.path.join(os.path.dirname(__file__), "golden", "wpg_golden.npz")


@pytest.mark.parametrize("c_init_z", [0.0, 0.013])
def test_tables_and_schedule_match_reference(c_init_z):
    g = dict(np.load(GOLD))
    tag = f"z{int(c_init_z * 1000):03d}"
    robot = P.RobotConstants()
    foot = list(robot.foot)
    for i in range(4):
        foot[3 * i + 2] = c_init_z
    robot.foot = tuple(foot)
    srbd = P.SRBDProblem(robot)
    srbd.createSRBDProblem(20, 1.0)
    gen = wpg.steps_phase(srbd.f, srbd.c, srbd.cdot, float(srbd.initial_foot_position[0][2]), srbd.c_ref, srbd.w_ref,
                          srbd.orientation_tracking_gain, srbd.cdot_switch, 20, number_of_legs=2,
                          contact_model=srbd.contact_model)
    assert gen.step_nodes == 10 and len(gen.l_cycle) == 21
    for name, mine in (("l_cycle", gen.l_cycle), ("l_switch", gen.l_cdot_switch),
                       ("r_cycle", gen.r_cycle), ("r_switch", gen.r_cdot_switch)):
        np.testing.assert_array_equal(np.array(mine), g[f"{tag}_{name}"])
    for t, action in enumerate(g[tag + "_actions"]):
        gen.set(str(action))
        c_ref = np.concatenate([srbd.c_ref[i].getValues() for i in range(4)], axis=0)
        sw = np.concatenate([srbd.cdot_switch[i].getValues() for i in range(4)], axis=0)
        np.testing.assert_array_equal(c_ref, g[tag + "_c_ref"][t])
        np.testing.assert_array_equal(sw, g[tag + "_switch"][t])
        np.testing.assert_array_equal(srbd.orientation_tracking_gain.getValues(), g[tag + "_otg"][t])
        np.testing.assert_array_equal(srbd.w_ref.getValues(), g[tag + "_w_ref"][t])


def test_fixture_W_values():
    """SURVEY.md section 8c fixture W."""
    lc, ls, rc, rs = wpg.gait_tables(0.0)
    assert list(ls) == [1, 1] + [0] * 8 + [1] * 11
    assert list(rs) == [1] * 12 + [0] * 8 + [1]
    np.testing.assert_allclose(lc[2:10], [0.00641, 0.01279, 0.01912, 0.02537, 0.03151, 0.03753, 0.04339, 0.04907], atol=5e-6)
    np.testing.assert_array_equal(lc[2:10], rc[12:20])


def test_flat_parameter_order():
    """ddp.py:165-177 flattening order == the p[19] / p[11] layouts of config.py."""
    srbd = P.SRBDProblem(); srbd.createSRBDProblem(20, 1.0)
    srbd.rdot_ref.assign([1, 2, 3], nodes=5)
    srbd.c_ref[2].assign(0.04, nodes=5); srbd.cdot_switch[3].assign(0.0, nodes=5)
    p = srbd.prb.flat_parameters()
    assert p.shape == (21, 19)
    np.testing.assert_array_equal(p[5], [1, 2, 3, 0, 0, 0, 10, 0, 1, 0, 1, 0.04, 1, 0, 0, 0, 0, 0, 1])
    lip = P.LIPProblem(); lip.createLIPProblem(20, 1.0)
    assert lip.prb.flat_parameters().shape == (21, 11)
    assert lip.prb.getDt() == pytest.approx(0.05)
