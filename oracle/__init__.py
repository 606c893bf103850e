"""Test-infrastructure package: CPU oracle of the GDSS sampling hot path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import from here.  See oracle/gdss_oracle.py.
"""
