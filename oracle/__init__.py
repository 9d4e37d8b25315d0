"""CPU oracle for the per-SNP REML scan.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (pygemma_b200/) never does.
"""
