"""TEST INFRASTRUCTURE -- the CPU oracle of the GRASP hot path.

Nothing in the product (grasp_b200/, modeling_grasp.py, grasp.py, tools/utils_func.py) may import
this package.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs use it, and only as the checker or the timed CPU baseline.

  restate.py     from-spec CPU restatement of the reference path (torch CPU fp32, LAPACK SVD),
                 every function citing the reference file:line it follows
  ref_import.py  imports the real reference from /root/reference (this container only)
  make_golden.py runs the REAL reference on small seeded inputs and writes tests/golden/*.pt;
                 also asserts that restate.py reproduces the reference on the same inputs

Pinning: the reference ships no tests, golden vectors or fixtures for this path ("parity
unpinned" by the reference's own tests).  The oracle is therefore pinned
against outputs of the reference itself, generated here by make_golden.py and committed under
tests/golden/ (tests/test_oracle_golden.py re-checks restate.py against them on every run).
"""
