"""CPU oracle for the NanoRepeat repeat-size hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  Nothing under nanorepeat_b200/ does.  PARITY UNPINNED: see the header of nr_oracle.c.
"""
