"""oracle/ -- TEST INFRASTRUCTURE (CPU checker), never imported by the product package.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package (see the header of hnh_oracle.c).
"""
