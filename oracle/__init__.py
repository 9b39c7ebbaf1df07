"""oracle/ -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import anything from here.  See oracle/README.md.
"""
