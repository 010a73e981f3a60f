// Device-side helpers shared by the kernel translation units (complex arithmetic, bit deposit, deterministic sums).
#pragma once
#include "common.cuh"

#define FULL 0xffffffffu

// ----------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// Im(conj(a) * b)
__device__ __forceinline__ double im_conj_mul(double2 a, double2 b) { return a.x * b.y - a.y * b.x; }

// insert a zero bit at each (ascending) position: maps a dense counter onto indices whose fixed bits are 0
__device__ __forceinline__ u64 deposit_zeros(u64 v, const unsigned char *pos, int npos) {
    for (int k = 0; k < npos; ++k) {
        const unsigned p = pos[k];
        const u64 low = v & ((1ull << p) - 1ull);
        v = ((v >> p) << (p + 1)) | low;
    }
    return v;
}

__device__ __forceinline__ double sign_of(u64 masked) { return (__popcll(masked) & 1) ? -1.0 : 1.0; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// deterministic block sum (fixed tree); result valid in thread 0
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sh /* NT/32 doubles */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < NT / 32) ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

__device__ __forceinline__ void load_pair_op(PairOp *dst, const PairOp *src) {
    const int words = sizeof(PairOp) / 8;
    for (int t = threadIdx.x; t < words; t += blockDim.x)
        reinterpret_cast<u64 *>(dst)[t] = __ldg(reinterpret_cast<const u64 *>(src) + t);
    __syncthreads();
}

struct Mat2 {
    double2 m00, m01, m10, m11;
};

__device__ __forceinline__ Mat2 op_matrix(const PairOp &op, int dagger) {
    Mat2 M;
    M.m00 = make_double2(op.m[0], op.m[1]);
    M.m01 = make_double2(op.m[2], op.m[3]);
    M.m10 = make_double2(op.m[4], op.m[5]);
    M.m11 = make_double2(op.m[6], op.m[7]);
    if (dagger) {
        const double2 t01 = cconj(M.m10), t10 = cconj(M.m01);
        M.m00 = cconj(M.m00);
        M.m11 = cconj(M.m11);
        M.m01 = t01;
        M.m10 = t10;
    }
    return M;
}

__device__ __forceinline__ void rot2(const Mat2 &M, double sgn, double2 &a, double2 &b) {
    const double2 sb = cscale(b, sgn), sa = cscale(a, sgn);
    const double2 ra = cadd(cmul(M.m00, a), cmul(M.m01, sb));
    const double2 rb = cadd(cmul(M.m10, sa), cmul(M.m11, b));
    a = ra;
    b = rb;
}


// Tile geometry helpers.  TileLaunch arrives as a kernel argument (constant bank); indexing its bits[] with a run-time
// subscript makes the compiler copy the struct to local memory and every access becomes a dependent LDL (measured
// with clock64 at 18 qubits: 850 cycles for the two masks, 1 450 for one deposit_zeros, 840 for the scatter tables of
// a ~10 000-cycle launch).  All loops below are fully unrolled with compile-time subscripts, so bits[] stays in the
// constant bank.
#define TILE_BITS_CAP 16
__device__ __forceinline__ unsigned tile_scatter(const TileLaunch &tl, int T, unsigned v, int lo, int hi) {
    unsigned g = 0;             // sum over local bits b in [lo, hi) of ((v >> (b - lo)) & 1) << bits[b]
#pragma unroll
    for (int b = 0; b < TILE_BITS_CAP; ++b)
        if (b >= lo && b < hi && b < T) g |= ((v >> (b - lo)) & 1u) << tl.bits[b];
    return g;
}

__device__ __forceinline__ unsigned tile_mask(const TileLaunch &tl, int T, int lo, int hi) {
    unsigned m = 0;
#pragma unroll
    for (int b = 0; b < TILE_BITS_CAP; ++b)
        if (b >= lo && b < hi && b < T) m |= 1u << tl.bits[b];
    return m;
}

// deposit_zeros(v, tl.bits, T) with compile-time subscripts
__device__ __forceinline__ u64 tile_base(const TileLaunch &tl, int T, u64 v) {
#pragma unroll
    for (int k = 0; k < TILE_BITS_CAP; ++k)
        if (k < T) {
            const unsigned p = tl.bits[k];
            const u64 low = v & ((1ull << p) - 1ull);
            v = ((v >> p) << (p + 1)) | low;
        }
    return v;
}

