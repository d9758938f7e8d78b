// tma_probe.cu — stand-alone probe of FP64 3-D TMA box loads with halo (negative / out-of-range
// coordinates, boxes larger than the tensor).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -o tools/tma_probe tools/tma_probe.cu ; run on the GPU box.  Development aid, not product.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, double* out) {
    __shared__ __align__(128) double buf[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int m = threadIdx.x; m < BW * BH; m += blockDim.x) buf[m] = -777.;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(BW * BH * 8)) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                smem_u32(buf)),
            "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
    }
    // wait phase 0
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(&bar)),
        "r"(0)
        : "memory");
    for (int m = threadIdx.x; m < BW * BH; m += blockDim.x) out[m] = buf[m];
}

static PFN_encodeTiled enc;

template <int BW, int BH>
static int run(const char* name, double* d, int d0, int d1, int d2, int sJ, int sK, int c0, int c1, int c2, const std::vector<double>& h) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)sJ * 8, (cuuint64_t)sK * 8};
    cuuint32_t box[3] = {BW, BH, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-28s encode FAILED (%d)\n", name, (int)r); return 1; }
    double* out;
    cudaMalloc(&out, BW * BH * 8);
    probe<BW, BH><<<1, 128>>>(m, c0, c1, c2, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s kernel FAILED: %s\n", name, cudaGetErrorString(e)); return 2; }
    std::vector<double> o(BW * BH);
    cudaMemcpy(o.data(), out, BW * BH * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int jj = 0; jj < BH; ++jj)
        for (int ii = 0; ii < BW; ++ii) {
            int i = c0 + ii, j = c1 + jj, k = c2;
            double want = (i >= 0 && i < d0 && j >= 0 && j < d1 && k >= 0 && k < d2) ? h[i + (size_t)sJ * j + (size_t)sK * k] : 0.;
            if (o[jj * BW + ii] != want) ++bad;
        }
    printf("%-28s ok, mismatches %d of %d\n", name, bad, BW * BH);
    cudaFree(out);
    return bad != 0;
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : -1;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode\n"); return 1; }
    enc = (PFN_encodeTiled)p;
    const int nI = 70, nJ = 40, nK = 9, sJ = 80, sK = 80 * 40;
    std::vector<double> h((size_t)sK * nK);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 1. + (double)i;
    double* d;
    cudaMalloc(&d, h.size() * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    int rc = 0;
    // one case per process: a TMA fault poisons the context
    switch (which) {
        case 0: rc = run<32, 16>("inside 32x16", d, nI, nJ, nK, sJ, sK, 2, 3, 4, h); break;
        case 1: rc = run<34, 18>("c0=-1 (odd) 34x18", d, nI, nJ, nK, sJ, sK, -1, -1, 4, h); break;
        case 2: rc = run<36, 18>("c0=-2 c1=-1 36x18", d, nI, nJ, nK, sJ, sK, -2, -1, 4, h); break;
        case 3: rc = run<36, 18>("c0=-2 plane -1", d, nI, nJ, nK, sJ, sK, -2, -1, -1, h); break;
        case 4: rc = run<36, 18>("beyond: c0=62 c1=31 plane nK", d, nI, nJ, nK, sJ, sK, 62, 31, nK, h); break;
        case 5: rc = run<36, 18>("far corner c0=62 c1=31", d, nI, nJ, nK, sJ, sK, 62, 31, nK - 1, h); break;
        case 6: rc = run<36, 18>("box > tensor (11,9,7)", d, 11, 9, 7, sJ, sK, -2, -1, 2, h); break;
        case 7: rc = run<68, 10>("68x10 c0=-2", d, nI, nJ, nK, sJ, sK, -2, -1, 2, h); break;
        case 8: rc = run<34, 18>("c0=1 (odd, inside)", d, nI, nJ, nK, sJ, sK, 1, 1, 2, h); break;
        case 9: rc = run<36, 18>("tiny tensor (2,2,2)", d, 2, 2, 2, sJ, sK, -2, -1, 0, h); break;
        default: printf("usage: tma_probe <case 0..9>\n"); return 0;
    }
    printf("probe rc %d\n", rc);
    return rc;
}
