"""oracle/ — CPU restatements of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from this package.  The product (gymrl_b200/) never does and has no CPU fallback.
"""
