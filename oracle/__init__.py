"""CPU oracle of the hot path -- TEST INFRASTRUCTURE ONLY (parity unpinned; see oracle/oracle.c).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
