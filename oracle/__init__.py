"""TEST INFRASTRUCTURE ONLY: CPU oracle for the tiddit_b200 hot path.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this package; the product package `tiddit_b200` never does.
"""
