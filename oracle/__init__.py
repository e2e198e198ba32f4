"""TEST INFRASTRUCTURE: the CPU oracle of libsbn_b200.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing under libsbn_b200/ does.
"""
