"""Real <-> momentum space maps (reference ``operators/fourier.py``).

``fourier_transform_matrix`` (reference :13-37) is the single-particle DFT that feeds the
Givens network; ``fourier_transform`` / ``inverse_fourier_transform`` (:39-115) rewrite a
FermionOperator mode by mode.  Conventions: site s = x + y*Nx, spin orbital 2s+spin;
creation operators pick up exp(-2 pi i m.n/N), annihilation operators the conjugate.
"""
import numpy as np

from fhsim.symbolic import FermionOperator, normal_ordered


def round_operator(op, decimals=6):
    """Round every coefficient to ``decimals`` places (reference :4-11)."""
    rounded = FermionOperator()
    for term, coeff in op.terms.items():
        rounded += FermionOperator(term, np.round(coeff, decimals=decimals))
    return rounded


def _site_xy(site, Nx):
    return site % Nx, site // Nx


def fourier_transform_matrix(x_dimension, y_dimension):
    """(2N x 2N) unitary, block diagonal in spin for the interleaved orbital order."""
    Nx, Ny = x_dimension, y_dimension
    n_sites = Nx * Ny
    matrix = np.zeros((2 * n_sites, 2 * n_sites), dtype=complex)
    for n_site in range(n_sites):
        nx, ny = _site_xy(n_site, Nx)
        for m_site in range(n_sites):
            mx, my = _site_xy(m_site, Nx)
            # same operation order as the reference expression so the entries agree to the last bit
            value = np.exp(-1j * 2 * np.pi * mx * nx / Nx) * np.exp(-1j * 2 * np.pi * my * ny / Ny)
            matrix[2 * n_site, 2 * m_site] = value
            matrix[2 * n_site + 1, 2 * m_site + 1] = value
    return matrix / np.sqrt(n_sites)


def _mode_expansion(Nx, Ny, sign):
    """Return f(index, ladder) -> FermionOperator: the expansion of one ladder operator."""
    n_sites = Nx * Ny
    norm = np.sqrt(n_sites)
    cache = {}

    def expand(index, ladder):
        key = (index, ladder)
        if key not in cache:
            nx, ny = _site_xy(index // 2, Nx)
            spin = index % 2
            direction = -sign if ladder else sign
            basis = FermionOperator()
            for m in range(n_sites):
                mx, my = _site_xy(m, Nx)
                angle = 2 * np.pi * (mx * nx / Nx + my * ny / Ny)
                basis += FermionOperator((2 * m + spin, ladder), np.exp(1j * direction * angle) / norm)
            cache[key] = basis
        return cache[key]

    return expand


def _transform(hamiltonian, Nx, Ny, sign):
    expand = _mode_expansion(Nx, Ny, sign)
    result = FermionOperator()
    for term, coeff in hamiltonian.terms.items():
        product = FermionOperator.identity()
        for index, ladder in term:
            product *= expand(index, ladder)
        result += product * coeff
        result = normal_ordered(result)
    result.compress()
    return round_operator(result)


def fourier_transform(hamiltonian, Nx, Ny):
    """a†_n -> sum_m exp(-2 pi i m.n/N) a†_m / sqrt(N)   (a_n with the conjugate phase)."""
    return _transform(hamiltonian, Nx, Ny, sign=+1)


def inverse_fourier_transform(hamiltonian, Nx, Ny):
    """Inverse map: a†_m -> sum_n exp(+2 pi i m.n/N) a†_n / sqrt(N)."""
    return _transform(hamiltonian, Nx, Ny, sign=-1)
