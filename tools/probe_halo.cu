// probe: can tcgen05.mma read its A operand from a HALO BOX that one tiled-mode TMA load wrote (18 x 10 pixels x 64 channels,
// 128B-swizzled), addressing filter tap (r, s) by nothing but the descriptor start address (+ (r*10 + s) * 128 B) and a
// stride-byte-offset of 10 pixels (1280 B) between 8-row groups?  Two descriptor variants: base_offset field 0, or
// (start >> 7) & 7.  (scratch tool; not part of the product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -I hyperpose_b200/csrc -o tools/probe_halo tools/probe_halo.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv_tcgen05.cuh"

using namespace hpb;

typedef CUresult (*PFN_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int TH = 16, TW = 8, R = 3, BH = TH + R - 1, BW = TW + R - 1; // box 18 x 10
constexpr int BOX_BYTES = BH * BW * 128;                               // 23040
constexpr int B_OFF = 23 * 1024;
constexpr int B_TILE = 64 * 128;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t base_off)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, float* out, int h0, int w0, int variant)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(ptx::smem_u32(&bar_ld), 1);
        ptx::mbar_init(ptx::smem_u32(&bar_mma), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 64);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t bar = ptx::smem_u32(&bar_ld);
        ptx::mbar_expect_tx(bar, BOX_BYTES + 9 * B_TILE);
        ptx::tma_load_4d(ptx::smem_u32(smem), &tm_x, bar, 0, w0 - 1, h0 - 1, 0);
        for (int t = 0; t < 9; ++t) ptx::tma_load_2d(ptx::smem_u32(smem + B_OFF + t * B_TILE), &tm_w, bar, 0, t * 64);
        ptx::mbar_wait(bar, 0);
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::make_idesc_f16(128, 64);
        for (int t = 0; t < 9; ++t) {
            const int r = t / 3, s = t % 3;
            const uint32_t a_addr = ptx::smem_u32(smem) + (uint32_t)((r * BW + s) * 128);
            const uint32_t bo = variant ? ((a_addr >> 7) & 7) : 0;
            const uint64_t da = make_desc(a_addr, BW * 128, bo);
            const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + B_OFF + t * B_TILE));
            for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (t | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(ptx::smem_u32(&bar_mma));
    }
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar_mma), 0);
    ptx::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int q = 0; q < 4; ++q) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * 16), v);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[row * 64 + q * 16 + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

int main()
{
    const int H = 20, W = 12, C = 64;
    std::vector<__half> x((size_t)H * W * C), w((size_t)9 * 64 * 64);
    srand(1);
    for (auto& v : x) v = __float2half((float)(rand() % 5 - 2));
    for (auto& v : w) v = __float2half((float)(rand() % 5 - 2));
    __half *dx, *dw; float* dout;
    cudaMalloc(&dx, x.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    if (!fp) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    PFN_tiled enc = (PFN_tiled)fp;
    CUtensorMap tmx, tmw;
    {
        cuuint64_t dims[4] = { C, W, H, 1 };
        cuuint64_t strides[3] = { (cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2 };
        cuuint32_t box[4] = { 64, BW, BH, 1 }, estr[4] = { 1, 1, 1, 1 };
        CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode x rc=%d\n", (int)r);
        if (r) return 1;
    }
    {
        cuuint64_t dims[2] = { 64, 9 * 64 };
        cuuint64_t strides[1] = { 128 };
        cuuint32_t box[2] = { 64, 64 }, estr[2] = { 1, 1 };
        CUresult r = enc(&tmw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode w rc=%d\n", (int)r);
        if (r) return 1;
    }
    const size_t smem = 1024 + B_OFF + 9 * B_TILE;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles[3][2] = { { 0, 0 }, { 4, 4 }, { 2, 3 } };
    for (int variant = 0; variant < 2; ++variant)
        for (auto& t : tiles) {
            const int h0 = t[0], w0 = t[1];
            cudaMemset(dout, 0, 128 * 64 * 4);
            probe<<<1, 128, smem>>>(tmx, tmw, dout, h0, w0, variant);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<float> o(128 * 64);
            cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0; double maxerr = 0; int first = -1;
            for (int y = 0; y < TH; ++y)
                for (int xx = 0; xx < TW; ++xx)
                    for (int oc = 0; oc < 64; ++oc) {
                        float ref = 0;
                        for (int r = 0; r < 3; ++r)
                            for (int s = 0; s < 3; ++s) {
                                const int hh = h0 + y + r - 1, ww = w0 + xx + s - 1;
                                if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                                for (int c = 0; c < C; ++c)
                                    ref += __half2float(x[((size_t)hh * W + ww) * C + c]) * __half2float(w[((size_t)(r * 3 + s) * 64 + oc) * 64 + c]);
                            }
                        const float got = o[(y * TW + xx) * 64 + oc];
                        const double err = fabs((double)got - ref);
                        if (err > 1e-3) { if (first < 0) first = (y * TW + xx) * 64 + oc; ++bad; }
                        if (err > maxerr) maxerr = err;
                    }
            printf("variant %d (base_offset %s) tile (h0=%d,w0=%d): mismatches %d / 8192, max err %g, first bad idx %d\n", variant,
                   variant ? "(addr>>7)&7" : "0", h0, w0, bad, maxerr, first);
        }
    return 0;
}
