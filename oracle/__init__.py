"""TEST INFRASTRUCTURE ONLY.  CPU checkers for the hot path; see oracle/pt_oracle.cpp and
oracle/ref_bridge.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this package; nothing under gdpathtracing_b200/ does."""
