"""TEST INFRASTRUCTURE ONLY: CPU oracle (plain C) + door onto the reference's own kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product never does.
"""
from .d2t_oracle import *  # noqa: F401,F403
