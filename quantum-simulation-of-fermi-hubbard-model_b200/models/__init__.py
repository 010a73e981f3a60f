"""Drop-in for the reference ``models/`` package: ADAPT / HVA / IQCC / VQE drivers on the fhsim backend."""
