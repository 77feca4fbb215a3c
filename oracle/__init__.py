"""CPU oracle for the srbd_horizon DDP hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See oracle/sddp_oracle.h.  PARITY UNPINNED for the
DDP iteration (pyddp is absent); model functions pinned to sympy goldens.
"""
