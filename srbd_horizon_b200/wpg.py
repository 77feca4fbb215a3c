"""Gait scheduler: the host-side mirror of /root/reference/python/wpg.py (class steps_phase).

Same constructor signature and `set(action)` semantics, so the reference's example
loops can drive it unchanged (dsrbd_example.py:79-80,126-131).  Behaviour kept,
including the reference's quirks:

* cycle = 2 steps of `step_duration`=0.5 s at dt=0.05 s -> 20 nodes, tables hold 21
  entries (wpg.py:19-64); left foot swings in the first step, right foot in the second;
* swing height = 0.1*sin over the first 8 interior samples of a **50-point**
  half-sine (`np.linspace(0, pi)` default length, wpg.py:28,37);
* `set` shifts c_ref / cdot_switch one node back, then writes node N:
  "step" from the tables (wpg.py:80-88), "jump" zeroes the switches and leaves c_ref
  untouched (wpg.py:89-93), anything else is stance with switches 1 and **c_ref = 0**
  (wpg.py:94-99).  w_ref = 0 and orientation_tracking_gain in {1e2, 0} go to node N too.

A batched, device-side version of the same tables lives in problems.py
(`gait_tables`, `schedule_params`).
"""
from __future__ import annotations

import numpy as np

STEP_DURATION = 0.5
GAIT_DT = 0.05
SS_SHARE = 0.8
DS_SHARE = 0.2
SWING_AMPLITUDE = 0.1


def gait_tables(c_init_z: float = 0.0):
    """(l_cycle, l_switch, r_cycle, r_switch), each of length 2*step_nodes + 1 (wpg.py:25-64)."""
    step_nodes = int(STEP_DURATION / GAIT_DT)
    ss = int(SS_SHARE * step_nodes)
    ds = int(DS_SHARE * step_nodes)
    bump = SWING_AMPLITUDE * np.sin(np.linspace(0, np.pi))[1:ss + 1]   # 50-sample half sine
    flat_ds, flat_ss = np.zeros(ds), np.zeros(ss)
    swing_first = np.concatenate([flat_ds, bump, flat_ds, flat_ss, [0.0]])
    swing_second = np.concatenate([flat_ds, flat_ss, flat_ds, bump, [0.0]])
    sw_first = np.concatenate([np.ones(ds), np.zeros(ss), np.ones(ds), np.ones(ss), [1.0]])
    sw_second = np.concatenate([np.ones(ds), np.ones(ss), np.ones(ds), np.zeros(ss), [1.0]])
    return c_init_z + swing_first, sw_first, c_init_z + swing_second, sw_second


class steps_phase:
    def __init__(self, f, c, cdot, c_init_z, c_ref, w_ref, orientation_tracking_gain, cdot_switch, nodes,
                 number_of_legs, contact_model):
        self.f, self.c, self.cdot = f, c, cdot
        self.c_ref, self.cdot_switch = c_ref, cdot_switch
        self.w_ref = w_ref
        self.orientation_tracking_gain = orientation_tracking_gain
        self.number_of_legs, self.contact_model = number_of_legs, contact_model
        self.nodes = nodes
        self.step_counter = 0
        self.step_duration, self.dt = STEP_DURATION, GAIT_DT
        self.ss_share, self.ds_share = SS_SHARE, DS_SHARE
        self.step_nodes = int(self.step_duration / self.dt)
        l_cycle, l_sw, r_cycle, r_sw = gait_tables(c_init_z)
        self.l_cycle, self.l_cdot_switch = list(l_cycle), list(l_sw)
        self.r_cycle, self.r_cdot_switch = list(r_cycle), list(r_sw)
        self.action = ""

    def set(self, action):
        self.action = action
        ref_id = self.step_counter % (2 * self.step_nodes)
        last = self.nodes
        n_contacts = self.contact_model * self.number_of_legs
        for i in range(n_contacts):            # shift the contact plan one node back
            for par in (self.cdot_switch[i], self.c_ref[i]):
                vals = par.getValues()
                for j in range(1, last + 1):
                    par.assign(vals[:, j], nodes=j - 1)
        self.w_ref.assign([0.0, 0.0, 0.0], nodes=last)
        if action == "step":
            self.orientation_tracking_gain.assign(1e2, nodes=last)
            for i in range(n_contacts):
                left = i < self.contact_model
                self.cdot_switch[i].assign((self.l_cdot_switch if left else self.r_cdot_switch)[ref_id], nodes=last)
                self.c_ref[i].assign((self.l_cycle if left else self.r_cycle)[ref_id], nodes=last)
        elif action == "jump":
            self.orientation_tracking_gain.assign(0.0, nodes=last)
            for i in range(len(self.c)):
                self.cdot_switch[i].assign(0.0, nodes=last)
        else:  # stance
            self.orientation_tracking_gain.assign(1e2, nodes=last)
            for i in range(len(self.c)):
                self.cdot_switch[i].assign(1.0, nodes=last)
                self.c_ref[i].assign(0.0, nodes=last)
        self.step_counter += 1
