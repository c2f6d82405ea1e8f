// probe of TMA im2col-mode semantics on sm_100a (scratch tool; not part of the product)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_im2col tools/probe_im2col.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*PFN_im2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                               cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int PIX>
__global__ void probe(const __grid_constant__ CUtensorMap tm, __half* out, int c, int w, int h, int n, int offw, int offh)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bar;
    unsigned char* tile = (unsigned char*)(((unsigned long long)smem + 1023) & ~1023ull);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(PIX * 128));
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
            ::"r"(smem_u32(tile)), "l"(&tm), "r"(smem_u32(&bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"((unsigned short)offw), "h"((unsigned short)offh)
            : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PIX * 64; i += blockDim.x) out[i] = ((__half*)tile)[i];
}

int main()
{
    const int N = 3, H = 5, W = 7, C = 64, PIX = 32;
    std::vector<__half> host((size_t)N * H * W * C);
    // value encodes the pixel: n*100 + h*10 + w (exact in fp16 up to 2048), same for all channels except ch0 (= value), ch1 = -1 marker
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int c = 0; c < C; ++c) host[(((size_t)n * H + h) * W + w) * C + c] = __float2half((float)(n * 100 + h * 10 + w + 1));
    __half* d; cudaMalloc(&d, host.size() * 2);
    cudaMemcpy(d, host.data(), host.size() * 2, cudaMemcpyHostToDevice);
    __half* dout; cudaMalloc(&dout, PIX * 64 * 2);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fp, cudaEnableDefault, &q);
    if (!fp) { printf("no cuTensorMapEncodeIm2col\n"); return 1; }
    PFN_im2col enc = (PFN_im2col)fp;
    const int pad = 1; // 3x3
    for (int variant = 0; variant < 2; ++variant) {
        CUtensorMap tm;
        cuuint64_t dims[4] = { C, W, H, N };
        cuuint64_t strides[3] = { (cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2 };
        int lower[2] = { -pad, -pad }, upper[2] = { pad - 2, pad - 2 }; // upper = pad - (R-1)
        cuuint32_t estr[4] = { 1, 1, 1, 1 };
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, lower, upper, 64, PIX, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         variant ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d encode rc=%d\n", variant, (int)r);
        if (r) continue;
        struct { int w, h, n, ow, oh; } cases[] = { { -1, -1, 0, 0, 0 }, { -1, -1, 0, 1, 1 }, { 0, 0, 0, 0, 0 }, { 2, 1, 0, 2, 2 }, { 3, 3, 2, 1, 1 }, { -1, 2, 1, 2, 0 } };
        for (auto& cs : cases) {
            cudaMemset(dout, 0xff, PIX * 64 * 2);
            probe<PIX><<<1, 128, PIX * 128 + 2048>>>(tm, dout, 0, cs.w, cs.h, cs.n, cs.ow, cs.oh);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<__half> o(PIX * 64);
            cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
            printf("start(w=%d,h=%d,n=%d) off(%d,%d):", cs.w, cs.h, cs.n, cs.ow, cs.oh);
            for (int p = 0; p < PIX; ++p) {
                // un-swizzle: chunk 0 of row p sits at chunk (0 ^ (p&7)) when swizzled
                const int chunk = variant ? (0 ^ (p & 7)) : 0;
                printf(" %g", __half2float(o[p * 64 + chunk * 8]));
            }
            printf("\n");
        }
    }
    return 0;
}
