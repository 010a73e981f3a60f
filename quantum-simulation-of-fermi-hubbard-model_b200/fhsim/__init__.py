"""fhsim: B200-native complex128 statevector backend for Fermi-Hubbard VQE studies.

Host side (this package): symbolic operators, Pauli-table compiler, circuit programs and the
ctypes binding to ``libfhsim.so`` (hand-written sm_100a CUDA behind the C-ABI in
``include/fhsim.h``).  There is no CPU execution path: every statevector operation goes
through the shared library and raises if it is missing.
"""
__version__ = "0.1.0"
