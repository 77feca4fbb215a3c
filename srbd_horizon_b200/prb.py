"""Problem definitions: the host-side mirror of /root/reference/python/prb.py.

The reference builds a symbolic `horizon.problem.Problem` (CasADi SX) and lets
`ddp.DDPSolver` turn it into per-node CasADi functions.  Here the dynamics and
cost of the two problems are hand-written CUDA (csrc/), so a problem object only
carries what the adapter and the callers read:

* variable names / dimensions in creation order (prb.py:32-68, 264-295) -- used to slice the
  solution into the `{name: dim x nodes}` dict of ddp.py:125-151;
* parameters with `assign(val, nodes=)` / `getValues(nodes=)` -- all that wpg.py:74-99 and
  dsrbd_example.py:102-122 touch;
* dt, number of nodes, the model tag and the numeric constants.

`SRBDProblem` / `LIPProblem` keep the reference's attribute names
(`prb, f, c, cdot, c_ref, w_ref, rdot_ref, oref, cdot_switch, orientation_tracking_gain,
initial_foot_position, com, force_scaling, m, I, nc, contact_model`).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Union

import numpy as np

from .config import (Gains, MODEL_LIP, MODEL_SRBD, RobotConstants)

NodeSpec = Union[None, int, Iterable[int]]


def _node_index(nodes: NodeSpec, n_nodes: int):
    if nodes is None:
        return slice(None)
    if isinstance(nodes, (int, np.integer)):
        if not 0 <= int(nodes) < n_nodes:
            raise IndexError(f"node {nodes} outside 0..{n_nodes - 1}")
        return int(nodes)
    idx = np.fromiter((int(n) for n in nodes), dtype=np.int64)
    if idx.size and (idx.min() < 0 or idx.max() >= n_nodes):
        raise IndexError(f"nodes outside 0..{n_nodes - 1}")
    return idx


class Variable:
    """A named state or input variable (only name and dimension are needed)."""

    def __init__(self, name: str, dim: int):
        self._name, self._dim = name, int(dim)

    def getName(self) -> str:
        return self._name

    def getDim(self) -> int:
        return self._dim

    def size(self):
        return (self._dim, 1)

    @property
    def shape(self):
        return (self._dim, 1)

    def __repr__(self):
        return f"Variable({self._name!r}, {self._dim})"


class Parameter(Variable):
    """Per-node parameter, values stored as dim x nodes (what ddp.py:173-177 iterates over)."""

    def __init__(self, name: str, dim: int, n_nodes: int):
        super().__init__(name, dim)
        self._values = np.zeros((self._dim, int(n_nodes)))

    def assign(self, val, nodes: NodeSpec = None) -> None:
        v = np.asarray(val, dtype=np.float64).reshape(-1)
        if v.size == 1 and self._dim > 1:
            v = np.full(self._dim, float(v[0]))
        if v.size != self._dim:
            raise ValueError(f"parameter {self._name!r}: expected {self._dim} values, got {v.size}")
        idx = _node_index(nodes, self._values.shape[1])
        if isinstance(idx, int):
            self._values[:, idx] = v
        else:
            self._values[:, idx] = v[:, None]

    def getValues(self, nodes: NodeSpec = None) -> np.ndarray:
        idx = _node_index(nodes, self._values.shape[1])
        return self._values[:, idx].copy()


class _Aggregate:
    def __init__(self, variables: List[Variable]):
        self._vars = variables

    def getVars(self) -> List[Variable]:
        return list(self._vars)

    def size(self):
        return (sum(v.getDim() for v in self._vars), 1)


class _VarContainer:
    def __init__(self, prb: "Problem"):
        self._prb = prb

    def getVarList(self, offset: bool = False) -> List[Variable]:
        return self._prb._state + self._prb._input


class Problem:
    """Light stand-in for horizon.problem.Problem (prb.py:21): N intervals, N+1 nodes."""

    def __init__(self, N: int, model: int, robot: RobotConstants, gains: Gains):
        self.N = int(N)
        self.nodes = self.N + 1          # ddp.py:83,90 use prb.nodes as the node count
        self.model = int(model)
        self.robot, self.gains = robot, gains
        self._state: List[Variable] = []
        self._input: List[Variable] = []
        self._params: Dict[str, Parameter] = {}
        self._dt: Optional[float] = None
        self.var_container = _VarContainer(self)

    def createStateVariable(self, name: str, dim: int) -> Variable:
        v = Variable(name, dim)
        self._state.append(v)
        return v

    def createInputVariable(self, name: str, dim: int) -> Variable:
        v = Variable(name, dim)
        self._input.append(v)
        return v

    def createParameter(self, name: str, dim: int) -> Parameter:
        p = Parameter(name, dim, self.nodes)
        self._params[name] = p
        return p

    def getState(self) -> _Aggregate:
        return _Aggregate(self._state)

    def getInput(self) -> _Aggregate:
        return _Aggregate(self._input)

    def getParameters(self) -> Dict[str, Parameter]:
        return self._params

    def getNNodes(self) -> int:
        return self.nodes

    def setDt(self, dt: float) -> None:
        self._dt = float(dt)

    def getDt(self) -> float:
        return self._dt

    def flat_parameters(self) -> np.ndarray:
        """[nodes, np] in the order ddp.py:165-177 flattens them (creation order, row-wise)."""
        return np.concatenate([p._values for p in self._params.values()], axis=0).T.copy()


class _ProblemBase:
    def __init__(self, robot: Optional[RobotConstants] = None, gains: Optional[Gains] = None):
        self.robot = robot or RobotConstants()
        self.gains = gains or Gains()

    def _common(self, prb: Problem, ns: int):
        foot = np.asarray(self.robot.foot, dtype=np.float64).reshape(4, 3)
        self.initial_foot_position = {i: foot[i].copy() for i in range(4)}
        self.com = np.asarray(self.robot.com, dtype=np.float64)
        self.force_scaling = float(self.robot.force_scaling)
        self.m = float(self.robot.mass)
        self.I = np.asarray(self.robot.inertia, dtype=np.float64).reshape(3, 3)
        self.contact_model = 2    # launch/SRBD_kangaroo_line_feet.launch:16
        self.nc = 4               # number_of_legs * contact_model (prb.py:41)
        self.prb = prb


class SRBDProblem(_ProblemBase):
    """prb.py:16-246 (nx=37, nu=24, np=19 at contact_model=2, number_of_legs=2)."""

    def createSRBDProblem(self, ns: int, T: float) -> None:
        prb = Problem(ns, MODEL_SRBD, self.robot, self.gains)
        self._common(prb, ns)
        prb.createStateVariable("r", 3)
        prb.createStateVariable("o", 4)
        self.c = {i: prb.createStateVariable(f"c{i}", 3) for i in range(self.nc)}
        prb.createStateVariable("rdot", 3)
        prb.createStateVariable("w", 3)
        self.cdot = {i: prb.createStateVariable(f"cdot{i}", 3) for i in range(self.nc)}
        self.f = {}
        for i in range(self.nc):
            prb.createInputVariable(f"cddot{i}", 3)
            self.f[i] = prb.createInputVariable(f"f{i}", 3)
        self.rdot_ref = prb.createParameter("rdot_ref", 3)
        self.w_ref = prb.createParameter("w_ref", 3)
        prb.setDt(T / ns)
        self.orientation_tracking_gain = prb.createParameter("orientation_tracking_gain", 1)
        self.orientation_tracking_gain.assign(1e1)                      # prb.py:143-144
        self.c_ref, self.cdot_switch = {}, {}
        for i in range(self.nc):                                         # prb.py:159-163
            self.c_ref[i] = prb.createParameter(f"c_ref{i}", 1)
            self.c_ref[i].assign(self.initial_foot_position[i][2])
            self.cdot_switch[i] = prb.createParameter(f"cdot_switch{i}", 1)
            self.cdot_switch[i].assign(1.0)
        self.oref = prb.createParameter("oref", 4)
        self.oref.assign([0.0, 0.0, 0.0, 1.0])                           # prb.py:185-186

    def getInitialState(self) -> np.ndarray:                             # prb.py:224-240
        x = np.zeros(37)
        x[0:3] = self.com
        x[6] = 1.0
        for i in range(4):
            x[7 + 3 * i:10 + 3 * i] = self.initial_foot_position[i]
        return x

    def getStaticInput(self) -> np.ndarray:                              # prb.py:242-246
        u = np.zeros(24)
        for i in range(4):
            u[6 * i + 5] = self.m * 9.81 / self.force_scaling / 4
        return u


class LIPProblem(_ProblemBase):
    """prb.py:248-441 (nx=30, nu=15, np=11)."""

    def createLIPProblem(self, ns: int, T: float) -> None:
        prb = Problem(ns, MODEL_LIP, self.robot, self.gains)
        self._common(prb, ns)
        prb.createStateVariable("r", 3)
        self.c = {i: prb.createStateVariable(f"c{i}", 3) for i in range(self.nc)}
        prb.createStateVariable("rdot", 3)
        self.cdot = {i: prb.createStateVariable(f"cdot{i}", 3) for i in range(self.nc)}
        prb.createInputVariable("z", 3)
        for i in range(self.nc):
            prb.createInputVariable(f"cddot{i}", 3)
        self.rdot_ref = prb.createParameter("rdot_ref", 3)
        prb.setDt(T / ns)
        self.c_ref, self.cdot_switch = {}, {}
        for i in range(self.nc):                                         # prb.py:372-376
            self.c_ref[i] = prb.createParameter(f"c_ref{i}", 1)
            self.c_ref[i].assign(self.initial_foot_position[i][2])
            self.cdot_switch[i] = prb.createParameter(f"cdot_switch{i}", 1)
            self.cdot_switch[i].assign(1.0)

    def getInitialState(self) -> np.ndarray:                             # prb.py:420-434
        x = np.zeros(30)
        x[0:3] = self.com
        for i in range(4):
            x[3 + 3 * i:6 + 3 * i] = self.initial_foot_position[i]
        return x

    def getStaticInput(self) -> np.ndarray:                              # prb.py:436-441
        u = np.zeros(15)
        u[0:2] = self.com[0:2]
        return u
