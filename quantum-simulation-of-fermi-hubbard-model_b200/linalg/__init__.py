"""Drop-in for the reference ``linalg/`` package (sector-restricted exact diagonalisation on the GPU)."""
