"""ORACLE -- test infrastructure only.

CPU restatement of the reference's algorithms for the statevector hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (fhsim, models, operators, linalg) never does.
"""
