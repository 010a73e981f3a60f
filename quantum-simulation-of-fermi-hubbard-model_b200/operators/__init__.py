"""Drop-in for the reference ``operators/`` package (pool, fourier, tools)."""
