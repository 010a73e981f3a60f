// Probe (measurement tool, not product): TMA tile gather of a 2^n complex128 state by arbitrary index bits.
//   1. does cuTensorMapEncodeTiled accept OVERLAPPING strides (dim 0 = the whole flat state, higher dims = single tile
//      bits with stride 16 B << bit)?  That turns "gather all amplitudes that differ only in the tile bits" into one
//      cp.async.bulk.tensor per <= 5 runs of tile bits.
//   2. what is the shared-memory layout with CU_TENSOR_MAP_SWIZZLE_128B and 16-byte elements (expected: 16-byte chunk
//      index c of box-linear offset l lands at l ^ ((l >> 3) & 7)), also when the inner box row is shorter than 128 B?
//   3. load + store round trip, and a first timing of TMA tile traffic on an L2-resident 18-qubit state.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/probe_tma tools/probes/probe_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k_probe(const __grid_constant__ CUtensorMap map, double2 *out_smem_image, int box_elems, int c0, int do_store,
                        const __grid_constant__ CUtensorMap map_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    double2 *buf = reinterpret_cast<double2 *>(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_elems * 16));
        if (RANK == 5)
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4, %4, %4}], [%2];"
                         ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(0) : "memory");
        else if (RANK == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4, %4}], [%2];"
                         ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4}], [%2];"
                         ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(0) : "memory");
    }
    // wait phase 0
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W;\n}" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < box_elems; i += blockDim.x) out_smem_image[i] = buf[i];
    if (do_store) {
        for (int i = threadIdx.x; i < box_elems; i += blockDim.x) buf[i].y = buf[i].x + 0.5;   // mark
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            if (RANK == 5)
                asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %3, %3, %3}], [%1];"
                             ::"l"(&map_out), "r"(smem_u32(buf)), "r"(c0), "r"(0) : "memory");
            else if (RANK == 4)
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %3, %3}], [%1];"
                             ::"l"(&map_out), "r"(smem_u32(buf)), "r"(c0), "r"(0) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %3}], [%1];"
                             ::"l"(&map_out), "r"(smem_u32(buf)), "r"(c0), "r"(0) : "memory");
            asm volatile("cp.async.bulk.commit_group;");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

// throughput: every CTA gathers tile blockIdx.x (11 tile bits) of an 18-qubit state and stores it back, `reps` times
__global__ void k_tma_copy_tiles(const __grid_constant__ CUtensorMap map, int tile_elems, int nissue, int issue_elems,
                                 const unsigned *__restrict__ c0_of_tile_issue, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    double2 *buf = reinterpret_cast<double2 *>(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    unsigned phase = 0;
    for (int r = 0; r < reps; ++r) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(tile_elems * 16));
            for (int q = 0; q < nissue; ++q) {
                const int c0 = (int)c0_of_tile_issue[blockIdx.x * nissue + q];
                asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4, %4, %4}], [%2];"
                             ::"r"(smem_u32(buf + q * issue_elems)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(0) : "memory");
            }
        }
        asm volatile("{\n.reg .pred p;\nW2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W2;\n}" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
        // touch: one op on the tile so the copy is not optimised away conceptually (TMA is opaque anyway)
        for (int i = threadIdx.x; i < tile_elems; i += blockDim.x) buf[i].x += 1.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 0; q < nissue; ++q) {
                const int c0 = (int)c0_of_tile_issue[blockIdx.x * nissue + q];
                asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %3, %3, %3}], [%1];"
                             ::"l"(&map), "r"(smem_u32(buf + q * issue_elems)), "r"(c0), "r"(0) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncthreads();
    }
}

// the same traffic with plain 128-bit loads/stores through registers (what k_tile does today)
__global__ void k_ldg_copy_tiles(double2 *psi, int T, const unsigned char *bits, int n, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double2 *buf = reinterpret_cast<double2 *>(smem);
    const unsigned L = 1u << T;
    unsigned base = blockIdx.x;
    for (int k = 0; k < T; ++k) {
        const unsigned p = bits[k];
        base = ((base >> p) << (p + 1)) | (base & ((1u << p) - 1u));
    }
    for (int r = 0; r < reps; ++r) {
        for (unsigned l = threadIdx.x; l < L; l += blockDim.x) {
            unsigned g = base;
            for (int b = 0; b < T; ++b) g |= ((l >> b) & 1u) << bits[b];
            buf[l] = psi[g];
        }
        __syncthreads();
        for (unsigned l = threadIdx.x; l < L; l += blockDim.x) buf[l].x += 1.0;
        __syncthreads();
        for (unsigned l = threadIdx.x; l < L; l += blockDim.x) {
            unsigned g = base;
            for (int b = 0; b < T; ++b) g |= ((l >> b) & 1u) << bits[b];
            psi[g] = buf[l];
        }
        __syncthreads();
    }
}

static int encode(CUtensorMap *m, double2 *base, int n, const std::vector<std::pair<int, int>> &runs /* (start bit, len) */,
                  CUtensorMapSwizzle sw, int inner_split) {
    // dim 0: the flat state in doubles (2^(n+1)); its box covers the low bits of run 0 (at most 7 bits = 256 doubles, or
    // `inner_split` bits).  Every other dim covers <= 8 consecutive tile bits with stride 16 B << start.  Always padded
    // to rank 5 with size-1 dims.  Returns -1 when more than 5 dims would be needed.
    std::vector<std::pair<int, int>> dims;     // (start bit, bits)
    int len0 = runs[0].second, first = inner_split > 0 ? inner_split : 7;
    if (first > len0) first = len0;
    dims.push_back({0, first});
    for (int done = first; done < len0;) { int take = len0 - done > 8 ? 8 : len0 - done; dims.push_back({done, take}); done += take; }
    for (size_t k = 1; k < runs.size(); ++k)
        for (int done = 0; done < runs[k].second;) {
            int take = runs[k].second - done > 8 ? 8 : runs[k].second - done;
            dims.push_back({runs[k].first + done, take});
            done += take;
        }
    if (dims.size() > 5) { printf("  needs %zu dims\n", dims.size()); return -1; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    gdim[0] = 2ull << n; box[0] = 2u << dims[0].second;
    for (int k = 1; k < 5; ++k) {
        if (k < (int)dims.size()) { gdim[k] = 1ull << dims[k].second; box[k] = 1u << dims[k].second; gstride[k - 1] = 16ull << dims[k].first; }
        else { gdim[k] = 1; box[k] = 1; gstride[k - 1] = 16ull << n; }
    }
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("  cuTensorMapEncodeTiled FAILED rc=%d\n", (int)r); return -1; }
    return 5;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDriverEntryPointQueryResult qres;
    void *fn = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    g_encode = (EncodeFn)fn;
    if (!g_encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int n = 18;
    const size_t dim = 1ull << n;
    std::vector<double2> h(dim);
    for (size_t i = 0; i < dim; ++i) h[i] = make_double2((double)i, 0.0);
    double2 *d, *d2, *dimg;
    CK(cudaMalloc(&d, dim * 16)); CK(cudaMalloc(&d2, dim * 16)); CK(cudaMalloc(&dimg, 8192 * 16));
    CK(cudaMemcpy(d, h.data(), dim * 16, cudaMemcpyHostToDevice));
    CK(cudaMemset(d2, 0, dim * 16));
    CK(cudaFuncSetAttribute(k_probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(k_probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(k_probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));

    struct Case { const char *name; std::vector<std::pair<int, int>> runs; CUtensorMapSwizzle sw; int split; unsigned base; };
    std::vector<Case> cases = {
        {"A: runs [0-2],[5],[9-10],[14] no swizzle", {{0, 3}, {5, 1}, {9, 2}, {14, 1}}, CU_TENSOR_MAP_SWIZZLE_NONE, 0, (1u << 3) | (1u << 12)},
        {"B: same, SWIZZLE_128B (inner = 128 B)", {{0, 3}, {5, 1}, {9, 2}, {14, 1}}, CU_TENSOR_MAP_SWIZZLE_128B, 0, (1u << 3) | (1u << 12)},
        {"C: runs [0-4],[7-8],[12] SWIZZLE_128B, run0 split 3+2", {{0, 5}, {7, 2}, {12, 1}}, CU_TENSOR_MAP_SWIZZLE_128B, 3, (1u << 5) | (1u << 17)},
        {"D: runs [0-1],[4],[6],[9-10] no swizzle (inner = 32 B rows)", {{0, 2}, {4, 1}, {6, 1}, {9, 2}}, CU_TENSOR_MAP_SWIZZLE_NONE, 0, (1u << 2) | (1u << 16)},
        {"E: runs [0],[3],[5],[7],[9] no swizzle (inner = 32 B)", {{0, 1}, {3, 1}, {5, 1}, {7, 1}, {9, 1}}, CU_TENSOR_MAP_SWIZZLE_NONE, 0, (1u << 1)},
        {"F: runs [0-3],[5],[7],[9],[11] no swizzle (W-like)", {{0, 4}, {5, 1}, {7, 1}, {9, 1}, {11, 1}}, CU_TENSOR_MAP_SWIZZLE_NONE, 0, (1u << 13) | (1u << 4)},
    };
    for (auto &c : cases) {
        printf("%s\n", c.name);
        CUtensorMap m, mo;
        int rank = encode(&m, d, n, c.runs, c.sw, c.split);
        int rank2 = encode(&mo, d2, n, c.runs, c.sw, c.split);
        if (rank < 0 || rank2 < 0) continue;
        int T = 0;
        std::vector<int> tbits;
        for (auto &r : c.runs) for (int b = 0; b < r.second; ++b) { tbits.push_back(r.first + b); ++T; }
        const int L = 1 << T;
        CK(cudaMemset(dimg, 0xff, 8192 * 16));
        const int c0 = (int)(c.base * 2);      // dim-0 coordinate in doubles
        if (rank == 5) k_probe<5><<<1, 128, L * 16 + 1024>>>(m, dimg, L, c0, 1, mo);
        else if (rank == 4) k_probe<4><<<1, 128, L * 16 + 1024>>>(m, dimg, L, c0, 1, mo);
        else k_probe<3><<<1, 128, L * 16 + 1024>>>(m, dimg, L, c0, 1, mo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<double2> img(L);
        CK(cudaMemcpy(img.data(), dimg, L * 16, cudaMemcpyDeviceToHost));
        // expected global index of box-linear offset l
        int ok_lin = 1, ok_sw = 1, ok_swrow = 1;
        for (int l = 0; l < L; ++l) {
            unsigned g = c.base;
            for (int b = 0; b < T; ++b) g |= ((l >> b) & 1u) << tbits[b];
            if (img[l].x != (double)g) ok_lin = 0;
            if (img[l ^ ((l >> 3) & 7)].x != (double)g) ok_sw = 0;
        }
        printf("  rank %d, tile 2^%d: layout linear=%d  swizzled(l ^ ((l>>3)&7))=%d\n", rank, T, ok_lin, ok_sw);
        if (!ok_lin && !ok_sw) {
            printf("  first 32 slots hold global indices:");
            for (int l = 0; l < 32 && l < L; ++l) printf(" %d", (int)img[l].x - (int)c.base);
            printf("\n");
        }
        (void)ok_swrow;
        // store round trip: d2 must hold (g, g + 0.5) on the tile and zeros elsewhere
        std::vector<double2> back(dim);
        CK(cudaMemcpy(back.data(), d2, dim * 16, cudaMemcpyDeviceToHost));
        size_t bad = 0, set = 0;
        for (size_t i = 0; i < dim; ++i) {
            unsigned tm = 0; for (int b : tbits) tm |= 1u << b;
            const bool in_tile = ((unsigned)i & ~tm) == c.base;
            if (in_tile) { set++; if (back[i].x != (double)i || back[i].y != (double)i + 0.5) bad++; }
            else if (back[i].x != 0.0 || back[i].y != 0.0) bad++;
        }
        printf("  store round trip: %zu tile amplitudes, %zu mismatches\n", set, bad);
        CK(cudaMemset(d2, 0, dim * 16));
    }

    // ---- timing: 128 CTAs x (gather 32 KB tile + store), bench-like tile bit sets, L2-resident ----
    {
        struct TS { const char *name; std::vector<int> bits; };
        std::vector<TS> sets = {
            {"ansatz tile [0,1,2,3,4,6,7,9,12,14,17]", {0, 1, 2, 3, 4, 6, 7, 9, 12, 14, 17}},
            {"W tile [0,1,2,3,5,7,9,11,13,15,17]", {0, 1, 2, 3, 5, 7, 9, 11, 13, 15, 17}},
            {"low tile [0..10]", {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10}},
        };
        CK(cudaFuncSetAttribute(k_tma_copy_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CK(cudaFuncSetAttribute(k_ldg_copy_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        for (auto &ts : sets) {
            const int T = (int)ts.bits.size();
            // runs of tile bits
            std::vector<std::pair<int, int>> runs;
            for (int b : ts.bits) {
                if (!runs.empty() && runs.back().first + runs.back().second == b) runs.back().second++;
                else runs.push_back({b, 1});
            }
            // first 5 runs (with run 0 possibly long) go into the map; the remaining tile bits are iterated
            // dims needed by run 0 (7 bits in dim 0, then 8 per dim); remaining runs one dim each (all <= 8 bits here)
            size_t dims0 = runs[0].second <= 7 ? 1 : 1 + (runs[0].second - 7 + 7) / 8;
            size_t take = runs.size();
            while (dims0 + (take - 1) > 5) --take;
            std::vector<std::pair<int, int>> in_map(runs.begin(), runs.begin() + take);
            std::vector<int> extra;
            for (size_t k = take; k < runs.size(); ++k) for (int b = 0; b < runs[k].second; ++b) extra.push_back(runs[k].first + b);
            int map_bits = 0; for (auto &r : in_map) map_bits += r.second;
            const int nissue = 1 << extra.size(), issue_elems = 1 << map_bits;
            CUtensorMap m;
            if (encode(&m, d, n, in_map, CU_TENSOR_MAP_SWIZZLE_NONE, 0) < 0) continue;
            const int ntiles = 1 << (n - T);
            std::vector<unsigned> c0(ntiles * nissue);
            std::vector<unsigned char> hb(ts.bits.begin(), ts.bits.end());
            for (int t = 0; t < ntiles; ++t) {
                unsigned base = t;
                for (int k = 0; k < T; ++k) { const unsigned p = ts.bits[k]; base = ((base >> p) << (p + 1)) | (base & ((1u << p) - 1u)); }
                for (int q = 0; q < nissue; ++q) {
                    unsigned g = base;
                    for (size_t e = 0; e < extra.size(); ++e) g |= ((q >> e) & 1u) << extra[e];
                    c0[t * nissue + q] = g * 2;
                }
            }
            unsigned *dc0; unsigned char *dbits;
            CK(cudaMalloc(&dc0, c0.size() * 4)); CK(cudaMemcpy(dc0, c0.data(), c0.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMalloc(&dbits, 16)); CK(cudaMemcpy(dbits, hb.data(), T, cudaMemcpyHostToDevice));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int variant = 0; variant < 2; ++variant) {
                for (int reps : {1, 9}) {
                    float best = 1e9f;
                    for (int it = 0; it < 6; ++it) {
                        cudaEventRecord(e0);
                        if (variant == 0) k_tma_copy_tiles<<<ntiles, 256, (1 << T) * 16 + 1024>>>(m, 1 << T, nissue, issue_elems, dc0, reps);
                        else k_ldg_copy_tiles<<<ntiles, 512, (1 << T) * 16 + 1024>>>(d, T, dbits, n, reps);
                        cudaEventRecord(e1);
                        CK(cudaEventSynchronize(e1));
                        float ms; cudaEventElapsedTime(&ms, e0, e1);
                        if (it > 0 && ms < best) best = ms;
                    }
                    printf("  %-44s %s reps=%d: %.2f us (%d issues of %d B per tile)\n", ts.name, variant == 0 ? "TMA" : "LDG", reps,
                           best * 1e3, nissue, issue_elems * 16);
                }
            }
            cudaFree(dc0); cudaFree(dbits);
        }
    }
    printf("done\n");
    return 0;
}
