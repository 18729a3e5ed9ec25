"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/gsm_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  PARITY UNPINNED: it restates SPEC.md, not the withheld GS-MARL
sources; only the linear-assignment routine is pinned (to scipy in this image).
"""
