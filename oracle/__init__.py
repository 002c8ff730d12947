"""TEST INFRASTRUCTURE: CPU restatement of the reference (oracle.py / adcensus_oracle.c), the
shim that builds the reference itself (refshim/, Makefile, _ref/) and its driver (refdriver.py).
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this package."""
