import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch
N, B = 50, 4736
mode = sys.argv[1]
cfg = make_config(MODEL_SRBD, N, 0.05, {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3})
s = BatchedDDP(cfg)
t = lambda v: torch.as_tensor(v, dtype=torch.float64, device="cuda")
b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True)
sel = np.arange(B) if mode == "mixed" else np.zeros(B, dtype=int)
x0, p, X0, U0 = (t(b[k][sel]) for k in ("x0", "params", "X0", "U0"))
for rep in range(2):
    r = s.solve(x0, p, X0, U0, gains=True, history=False, order="schedule" if mode == "mixed" else None)
torch.cuda.synchronize()
print(mode, float(r.iters.double().sum()) * N)
