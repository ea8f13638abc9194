"""CPU oracle for treat_b200 -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this package.  The product (treat_b200/) never does.
"""
