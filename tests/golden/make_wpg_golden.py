#!/usr/bin/env python
"""Generate tests/golden/wpg_golden.npz by RUNNING the reference's gait scheduler
(/root/reference/python/wpg.py, pure numpy) in this container with a minimal stand-in
for horizon's Parameter (assign / getValues only).  The reference tree does not exist
on the GPU box, hence the committed fixture.

Recorded: the four 21-entry tables and, for a scripted action sequence, the
c_ref / cdot_switch / w_ref / orientation_tracking_gain arrays after every `set`.

Run from the repo root:  python tests/golden/make_wpg_golden.py
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/python/wpg.py"


class StubParam:
    def __init__(self, dim, nodes, init=0.0):
        self.v = np.full((dim, nodes), float(init))

    def assign(self, val, nodes=None):
        val = np.asarray(val, dtype=float).reshape(-1)
        if nodes is None:
            self.v[:, :] = val[:, None]
        else:
            self.v[:, nodes] = val

    def getValues(self, nodes=None):
        return self.v.copy() if nodes is None else self.v[:, nodes].copy()


def main():
    spec = importlib.util.spec_from_file_location("ref_wpg", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ns, nc = 20, 4
    out = {}
    for c_init_z in (0.0, 0.013):
        tag = f"z{int(c_init_z * 1000):03d}"
        c_ref = {i: StubParam(1, ns + 1, c_init_z) for i in range(nc)}
        sw = {i: StubParam(1, ns + 1, 1.0) for i in range(nc)}
        w_ref = StubParam(3, ns + 1)
        otg = StubParam(1, ns + 1, 10.0)
        dummy = {i: None for i in range(nc)}
        g = ref.steps_phase(dummy, dummy, dummy, c_init_z, c_ref, w_ref, otg, sw, ns, number_of_legs=2, contact_model=2)
        out[tag + "_l_cycle"] = np.array(g.l_cycle)
        out[tag + "_l_switch"] = np.array(g.l_cdot_switch)
        out[tag + "_r_cycle"] = np.array(g.r_cycle)
        out[tag + "_r_switch"] = np.array(g.r_cdot_switch)
        actions = ["standing"] * 3 + ["step"] * 27 + ["jump"] * 4 + ["step"] * 5 + ["standing"] * 2
        rec_c, rec_s, rec_o, rec_w = [], [], [], []
        for a in actions:
            g.set(a)
            rec_c.append(np.concatenate([c_ref[i].v for i in range(nc)], axis=0))
            rec_s.append(np.concatenate([sw[i].v for i in range(nc)], axis=0))
            rec_o.append(otg.v.copy())
            rec_w.append(w_ref.v.copy())
        out[tag + "_actions"] = np.array(actions)
        out[tag + "_c_ref"] = np.array(rec_c)
        out[tag + "_switch"] = np.array(rec_s)
        out[tag + "_otg"] = np.array(rec_o)
        out[tag + "_w_ref"] = np.array(rec_w)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wpg_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))
    print(out["z000_l_switch"], out["z000_r_switch"], out["z000_l_cycle"][2:10])


if __name__ == "__main__":
    main()
