"""CPU parity oracle for the DFTB+ hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under dftbplus_b200/ does.  See oracle/README.md.
"""
from .oracle import *  # noqa: F401,F403
