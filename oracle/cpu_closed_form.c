/* ORACLE / CPU BASELINE (test infrastructure only -- never linked into or called by the product path).
 *
 * Plain C + OpenMP restatement, on the host cores, of the closed form the GPU path evaluates for one ADAPT screening
 * (reference ADAPT.select_operator, models/adapt_vqe.py:297-323, via SURVEY 3.2's identity
 *      g_k = d<H>/d e_k |_{e=0} = 2 Im <lambda| G_k |psi>,  lambda = W^dagger H W psi ):
 *   cf_pair   2x2 block on index pairs (i, i^x) selected by (i & fixmask) == fixval with parity sign   (a rotation
 *             exp(-i theta G) of reference Trotterize_generator, adapt_vqe.py:87-98, or a Givens of W, :347-354)
 *   cf_diag   exp(-i sum_m a_m (-1)^popcount(i & z_m))                                              (RZ layers, :344-345)
 *   cf_table  out = sum_t c_t P_t in, P_t = i^k X^x Z^z        (qml.expval(Hamiltonian) as a matvec, :357-361)
 *   cf_pool   g_k for every pool entry
 * It is what bench.py reports as "cpu_closed_form" (1 thread / all cores): the SAME algorithm as the CUDA path on a CPU,
 * so the algorithmic gain (closed form vs append-and-backprop) can be told apart from the kernel/hardware gain.
 * Checked against oracle/statevector.py (numpy) in tests/test_cpu_closed_form.py.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC -o oracle/_build/libcpu_closed_form.so oracle/cpu_closed_form.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

static inline int parity64(uint64_t v) { return __builtin_popcountll(v) & 1; }

void cf_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int cf_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* dense counter -> index with the fixed bits cleared (ascending positions) */
static inline uint64_t deposit(uint64_t v, const int *pos, int npos) {
    for (int k = 0; k < npos; ++k) {
        const uint64_t low = v & ((1ull << pos[k]) - 1ull);
        v = ((v >> pos[k]) << (pos[k] + 1)) | low;
    }
    return v;
}

/* psi_i <- m00 psi_i + s m01 psi_j ; psi_j <- s m10 psi_i + m11 psi_j, j = i^x, s = (-1)^popcount(i & zeta) */
void cf_pair(cplx *psi, int n, uint64_t x, uint64_t fixmask, uint64_t fixval, uint64_t zeta, const double *m) {
    int pos[64], npos = 0;
    for (int b = 0; b < n; ++b)
        if (fixmask >> b & 1ull) pos[npos++] = b;
    const int64_t npairs = (int64_t)1 << (n - npos);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < npairs; ++k) {
        const uint64_t i = deposit((uint64_t)k, pos, npos) | fixval, j = i ^ x;
        const double s = parity64(i & zeta) ? -1.0 : 1.0;
        const cplx a = psi[i], b = psi[j];
        const double m01r = s * m[2], m01i = s * m[3], m10r = s * m[4], m10i = s * m[5];
        psi[i].re = m[0] * a.re - m[1] * a.im + m01r * b.re - m01i * b.im;
        psi[i].im = m[0] * a.im + m[1] * a.re + m01r * b.im + m01i * b.re;
        psi[j].re = m10r * a.re - m10i * a.im + m[6] * b.re - m[7] * b.im;
        psi[j].im = m10r * a.im + m10i * a.re + m[6] * b.im + m[7] * b.re;
    }
}

void cf_diag(cplx *psi, int n, int nterms, const uint64_t *z, const double *angle) {
    const int64_t dim = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < dim; ++i) {
        double tot = 0.0;
        for (int t = 0; t < nterms; ++t) tot += parity64((uint64_t)i & z[t]) ? -angle[t] : angle[t];
        const double c = cos(tot), s = -sin(tot);
        const cplx a = psi[i];
        psi[i].re = c * a.re - s * a.im;
        psi[i].im = c * a.im + s * a.re;
    }
}

/* out[i] = sum_t c_t i^{k_t} (-1)^{popcount((i^x_t) & z_t)} in[i^x_t] */
void cf_table(const cplx *in, cplx *out, int n, int nterms, const uint64_t *x, const uint64_t *z, const double *cre,
              const double *cim) {
    const int64_t dim = (int64_t)1 << n;
    static const double ipr[4] = {1, 0, -1, 0}, ipi[4] = {0, 1, 0, -1};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < dim; ++i) {
        double ar = 0.0, ai = 0.0;
        for (int t = 0; t < nterms; ++t) {
            const uint64_t j = (uint64_t)i ^ x[t];
            const int k = __builtin_popcountll(x[t] & z[t]) & 3;
            const double s = parity64(j & z[t]) ? -1.0 : 1.0;
            const double wr = s * (cre[t] * ipr[k] - cim[t] * ipi[k]), wi = s * (cre[t] * ipi[k] + cim[t] * ipr[k]);
            ar += wr * in[j].re - wi * in[j].im;
            ai += wr * in[j].im + wi * in[j].re;
        }
        out[i].re = ar;
        out[i].im = ai;
    }
}

double cf_inner_re(const cplx *a, const cplx *b, int n) {
    const int64_t dim = (int64_t)1 << n;
    double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t i = 0; i < dim; ++i) acc += a[i].re * b[i].re + a[i].im * b[i].im;
    return acc;
}

/* g_e = 2 Im sum_{pairs} s ( conj(lam_i) B psi_j + conj(lam_j) conj(B) psi_i ) for every entry e (one pattern-pinned
 * pair piece per pool operator: the 8 JW strings of i(a+a+aa - h.c.) act on index patterns 1100 <-> 0011 only) */
void cf_pool(const cplx *psi, const cplx *lam, int n, int nentries, const uint64_t *x, const uint64_t *fixmask,
             const uint64_t *fixval, const uint64_t *zeta, const double *br, const double *bi, double *out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int e = 0; e < nentries; ++e) {
        int pos[64], npos = 0;
        for (int b = 0; b < n; ++b)
            if (fixmask[e] >> b & 1ull) pos[npos++] = b;
        const int64_t npairs = (int64_t)1 << (n - npos);
        double acc = 0.0;
        for (int64_t k = 0; k < npairs; ++k) {
            const uint64_t i = deposit((uint64_t)k, pos, npos) | fixval[e], j = i ^ x[e];
            const double s = parity64(i & zeta[e]) ? -1.0 : 1.0;
            const cplx a = psi[i], b = psi[j], la = lam[i], lb = lam[j];
            /* B b and conj(B) a */
            const double gr_i = br[e] * b.re - bi[e] * b.im, gi_i = br[e] * b.im + bi[e] * b.re;
            const double gr_j = br[e] * a.re + bi[e] * a.im, gi_j = br[e] * a.im - bi[e] * a.re;
            acc += s * ((la.re * gi_i - la.im * gr_i) + (lb.re * gi_j - lb.im * gr_j));
        }
        out[e] = 2.0 * acc;
    }
}
