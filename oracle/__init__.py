"""TEST INFRASTRUCTURE ONLY -- CPU checkers for the stable-fluids step.

`oracle.sfo`   ctypes binding of oracle/stable_fluids_oracle.c (the plain-C restatement).
`oracle.refs`  ctypes bindings of oracle/_ref/libref_{cpu,gpu}.so (the UNMODIFIED reference
               solvers compiled out of /root/reference by oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (fluid-2d_b200) never does.
"""
