"""ORACLE -- test infrastructure only.

CPU restatement of the reference's algorithms for the statevector hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (fhsim, models, operators, linalg) never does.

Pinning status
  * host tables (pool order/content, k-space occupation, sector index order, HVA layer colouring,
    Fourier matrix, H split): PINNED against fixtures produced by executing the reference's own
    Python (tests/golden/make_golden.py -> tests/golden/reference_host_tables.json,
    checked in tests/test_golden_reference.py for the oracle and for the product).
  * statevector arithmetic (PennyLane default.qubit.torch, torch autograd, OpenFermion jordan_wigner /
    get_sparse_operator, scipy eigsh): the reference holds no tests or golden vectors and its
    dependencies are not installable here -> PARITY UNPINNED by the reference; pinned instead to analytic
    known answers (tests/test_oracle_known_answers.py) and by two independent formulations that must
    agree (oracle/literal.py gate-by-gate + autograd  ==  oracle/statevector.py closed form).
"""
