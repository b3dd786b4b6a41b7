"""Test oracle for the set-abstraction path -- TEST INFRASTRUCTURE ONLY (see ssd3d_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
