// conv_tcgen05.cuh -- implicit-GEMM convolution on the sm_100a 5th-gen tensor cores.
//
// Replaces the TensorRT-executed CNN of the reference (IExecutionContext::executeV2,
// src/tensorrt.cpp:387-396; layer shapes from hyperpose/Model/backbones.py:447-509 and
// hyperpose/Model/openpose/model/openpose.py:36-199).
//
//   D[128 pixels, BN out-channels] += A[128 pixels, 64 in-channels] * B[BN, 64]^T
//   summed over (filter tap r,s) x (64-channel chunk): one "k-step" per (r, s, chunk).
//
// * activations are fp16 NHWC; the A tile of a k-step is ONE im2col-mode TMA load
//   (cp.async.bulk.tensor.4d...im2col): 128 consecutive output pixels (in n,h,w order, wrapping over rows
//   and images inside the TMA unit) x 64 channels at filter offset (s, r); out-of-image elements are
//   zero-filled by the TMA unit = "SAME" padding, no im2col buffer in HBM and no tile padding waste;
// * weights are fp16 [G][Cout_pad][R][S][Cin_g] (K-major); the B tile is a 2-D TMA box {64, BN};
// * both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes;
// * accumulators live in TMEM (2 x BN fp32 columns, double-buffered so the epilogue of tile i
//   overlaps the MMAs of tile i+1);
// * persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer
//   (single elected thread), warp 2 = TMEM allocator, warps 4-7 = epilogue
//   (tcgen05.ld -> bias + PReLU/ReLU -> fp16 NHWC (or fp32 NCHW for the parser) stores).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace hpb {

constexpr int CONV_BLOCK_M = 128;   // pixels per tile == TMEM lanes == UMMA M
constexpr int CONV_BLOCK_K = 64;    // fp16 channels per k-step == one 128-byte swizzle row
constexpr int CONV_UMMA_K = 16;     // K of one tcgen05.mma.kind::f16
constexpr int CONV_MAX_STAGES = 8;
constexpr int CONV_THREADS = 256;
constexpr int CONV_IM2COL_THREADS = 384;   // conv_tcgen05_kernel: warps 0-3 roles, 4-7 epilogue, 8-11 second epilogue set
constexpr int CONV_A_BYTES = CONV_BLOCK_M * CONV_BLOCK_K * 2; // 16 KiB
constexpr size_t CONV_SMEM_LIMIT = 227 * 1024;

enum ConvOutMode : int {
    OUT_F16_NHWC = 0,       // fp16, out[pixel * ld + ch_off + g * cout_g + n]
    OUT_F32_NCHW_SPLIT = 1, // fp32 planar; channels [0,split) -> out, [split,cout_g) -> out2 (conf / paf for the parser)
};

struct ConvParams {
    int Nb, H, W;              // batch, spatial size (stride 1, "SAME" padding: output size == input size)
    int R, S;                  // filter taps
    int groups, cin_g;         // cin_g: multiple of 64
    int cout_g, cout_g_pad;    // real / padded (multiple of BN) output channels per group
    int BN;                    // tile N == UMMA N (multiple of 16, <= 256)
    int m_tiles;               // ceil(Nb * H * W / 128): a tile is 128 CONSECUTIVE output pixels in (n, h, w) order
    int in_ch_off;             // first input channel inside the input buffer
    int num_stages;
    int tmem_cols;             // power of two >= 2 * BN
    const float* bias;         // [groups * cout_g_pad]
    const float* alpha;        // [groups * cout_g_pad]   y = v > 0 ? v : alpha * v   (0 => ReLU, 1 => linear)
    int out_mode;
    void* out; void* out2;
    int out_ld, out_ch_off;    // NHWC: channels per pixel of the output buffer / first channel written
    int split;                 // NCHW_SPLIT: channels [0,split) go to out, the rest to out2
    int tma_store;             // 1: fp16 NHWC output goes through swizzled smem staging + TMA tensor stores (BN % 64 == 0)
    int swap_ab;               // 1: conv_tcgen05_swap_kernel (cout_g_pad % 128 == 0, TMA store)
    int npx;                   // swap kernel: pixels per unit (UMMA N), multiple of 16, 128 < npx <= 256
    const __half* res;         // residual input [pixels, res_ld] (+ res_ch_off), or nullptr
    int res_ld, res_ch_off;
    int res_mode;              // 1: y = act(v + res)   2: y = act(v) + res
    int relu_only;             // 1: every slope of the layer is 0 (plain ReLU): the epilogue skips the slope loads
    int epi_warps;             // conv_tcgen05_kernel: 8 = two epilogue warps per TMEM lane quarter (each takes half of the channels), 4 = one
    // a 1x1 "depthwise" op that follows the conv (per-channel scale + bias + PReLU: the filter_size (1,1) separable blocks of
    // MobilenetThin-OpenPose) applied in the epilogue, with the fp16 rounding of the tensor in between kept: bit-identical with the two launches
    const float* post_w; const float* post_b; const float* post_a;   // [groups * cout_g] each, or nullptr
    int epi_one_bar;               // conv_tcgen05_kernel TMA-store epilogue: one named barrier per 64-channel sub-tile instead of two (see there)
    int res_stages;                // depth of the residual-tile ring of conv_tcgen05_kernel's TMA-store epilogue (2 or 4)
    int split_from, total_items;   // conv_tcgen05_kernel work list: items >= split_from are N-halves (see decode_tile); total_items = tiles + (tiles - split_from)
};

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// im2col-mode load: {c, w, h, n} = channel + base pixel (output pixel minus padding); {offw, offh} = filter tap
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n, int offw, int offh)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((unsigned short)offw), "h"((unsigned short)offh)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// multicast variants (thread-block clusters): the tile lands at the SAME shared-memory offset in every CTA of cta_mask and
// completes bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// arrives on the mbarrier at this offset in every CTA of cta_mask once the previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// programmatic dependent launch (no-ops when the kernel was launched without the attribute)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc];  accumulate = 0 overwrites D
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (count 1) on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (= 1024 B between 8-row groups) | [46,48) version = 1 (sm_100)
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, A/B = fp16 K-major, D = fp32:
//   [4,6) c_format = 1 (F32) | [7,10) a_format = 0 (F16) | [10,13) b_format = 0 (F16)
//   [15] a_major = 0 (K) | [16] b_major = 0 (K) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

} // namespace ptx

struct ConvTile {
    int p0, g, n0; // first output pixel (flattened n,h,w), group, first output channel of the tile
    int bn;        // channels of the tile: BN, or BN / 2 for the N-halves of the ragged last round (split_from)
};

// Work item -> tile.  Items [0, split_from) are the full 128 x BN tiles of the rounds that fill the persistent grid; the tiles of
// a ragged last round (fewer than half a wave) are cut into two N-halves each, items split_from + 2 j + {0, 1}: the tail of the
// kernel runs twice as many CTAs for ~0.6 of a tile time (472 tiles on 148 SMs: 3 rounds + 28 tiles -> 3 rounds + 56 half tiles).
__device__ __forceinline__ ConvTile decode_tile(const ConvParams& p, int item, int n_tiles_g)
{
    ConvTile t;
    int tile = item, half = 0;
    t.bn = p.BN;
    if (item >= p.split_from) { const int j = item - p.split_from; tile = p.split_from + (j >> 1); half = j & 1; t.bn = p.BN >> 1; }
    const int n_tiles_total = p.groups * n_tiles_g;
    const int nt = tile % n_tiles_total;
    const int mt = tile / n_tiles_total;
    t.g = nt / n_tiles_g;
    t.n0 = (nt - t.g * n_tiles_g) * p.BN + half * (p.BN >> 1);
    t.p0 = mt * CONV_BLOCK_M;
    return t;
}

struct PixelPos {
    int n, h, w;
};
__device__ __forceinline__ PixelPos unflatten(const ConvParams& p, int px)
{
    PixelPos q;
    const int hw = p.H * p.W;
    q.n = px / hw;
    const int rem = px - q.n * hw;
    q.h = rem / p.W;
    q.w = rem - q.h * p.W;
    return q;
}


// Epilogue arithmetic of one pixel for 16 consecutive output channels: bias (+ residual) + PReLU -> fp16 pairs.
// relu_only layers compute max(a, 0) with the sign of a negative input kept on the zero, which is what a * 0.f gives
// (bit-identical with the general form, one FMNMX + LOP3 instead of FSETP + FMUL + FSEL, and no slope loads).
template <bool kRes>
__device__ __forceinline__ void conv_epilogue16(const uint32_t (&v)[16], const float (&bv)[16], const float (&av)[16], const bool relu_only,
                                                const int res_mode, const __half2 (&rs)[8], uint32_t (&pk)[8])
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a0 = __uint_as_float(v[2 * j]) + bv[2 * j];
        float a1 = __uint_as_float(v[2 * j + 1]) + bv[2 * j + 1];
        float r0 = 0.f, r1 = 0.f;
        if (kRes && res_mode) { const float2 rf = __half22float2(rs[j]); r0 = rf.x; r1 = rf.y; }
        if (kRes && res_mode == 1) { a0 += r0; a1 += r1; }
        if (relu_only) {
            a0 = __uint_as_float(__float_as_uint(fmaxf(a0, 0.f)) | (__float_as_uint(a0) & 0x80000000u));
            a1 = __uint_as_float(__float_as_uint(fmaxf(a1, 0.f)) | (__float_as_uint(a1) & 0x80000000u));
        } else {
            a0 = a0 > 0.f ? a0 : a0 * av[2 * j];
            a1 = a1 > 0.f ? a1 : a1 * av[2 * j + 1];
        }
        if (kRes && res_mode == 2) { a0 += r0; a1 += r1; }
        const __half2 h2 = __floats2half2_rn(a0, a1);
        pk[j] = *(const uint32_t*)&h2;
    }
}
// the fused 1x1 depthwise stage on 16 packed channels: x = the fp16 value the conv would have stored; y = fp16(act(x * w + b))
// with dwconv_kernel<1, 1>'s arithmetic (fmaf(x, w, 0) + b, slope, round)
__device__ __forceinline__ void conv_post16(uint32_t (&pk)[8], const float* __restrict__ pw, const float* __restrict__ pb, const float* __restrict__ pa)
{
    float w[16], b[16], a[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        *(float4*)&w[4 * j] = __ldg((const float4*)pw + j);
        *(float4*)&b[4 * j] = __ldg((const float4*)pb + j);
        *(float4*)&a[4 * j] = __ldg((const float4*)pa + j);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 x = __half22float2(*(const __half2*)&pk[j]);
        float y0 = fmaf(x.x, w[2 * j], 0.f) + b[2 * j], y1 = fmaf(x.y, w[2 * j + 1], 0.f) + b[2 * j + 1];
        y0 = y0 > 0.f ? y0 : y0 * a[2 * j];
        y1 = y1 > 0.f ? y1 : y1 * a[2 * j + 1];
        const __half2 h2 = __floats2half2_rn(y0, y1);
        pk[j] = *(const uint32_t*)&h2;
    }
}

template <bool kRes> // kRes: residual epilogue compiled in (ResNet / LW-OpenPose blocks); false keeps the plain epilogue lean
__global__ void __launch_bounds__(CONV_IM2COL_THREADS, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_r,
                    const __grid_constant__ CUtensorMap tmap_bh /* weight box of BN / 2 rows (N-halves) */, const ConvParams p)
{
    extern __shared__ uint8_t smem_raw[];
    ptx::pdl_launch_dependents();   // the next kernel may start its prologue on every SM this grid has left
    // 1024-byte alignment: required by the 128B swizzle atoms shared by TMA and UMMA
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.BN * CONV_BLOCK_K * 2;
    const int stage_bytes = CONV_A_BYTES + b_bytes;
    uint8_t* bar_base = smem + (size_t)p.num_stages * stage_bytes;
    uint64_t* full_bar = (uint64_t*)bar_base;                  // [stages]  TMA -> MMA
    uint64_t* empty_bar = full_bar + CONV_MAX_STAGES;          // [stages]  MMA -> TMA
    uint64_t* tfull_bar = empty_bar + CONV_MAX_STAGES;         // [2]       MMA -> epilogue
    uint64_t* tempty_bar = tfull_bar + 2;                      // [2]       epilogue -> MMA
    uint64_t* res_bar = tempty_bar + 2;                        // [4]       residual TMA loads -> epilogue
    uint32_t* tmem_slot = (uint32_t*)(res_bar + 4);
    // two 16 KiB staging tiles (128 pixels x 64 channels fp16, 128B-swizzled) for the TMA-store epilogue
    uint8_t* out_stage = (uint8_t*)(((uintptr_t)(tmem_slot + 4) + 1023) & ~(uintptr_t)1023);
    // residual epilogue (res_tma): res_stages (2 or 4) more 16 KiB tiles, filled by TMA loads the epilogue leader issues ahead of use
    uint8_t* res_stage = out_stage + 2 * CONV_A_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_g = p.cout_g_pad / p.BN;
    const int total_tiles = p.total_items;   // work items: full tiles, then the N-halves of a ragged last round
    const int chunks = p.cin_g / CONV_BLOCK_K;
    const int ksteps = p.R * p.S * chunks;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        if (p.tma_store) ptx::prefetch_tmap(&tmap_o);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.num_stages; ++i) {
            ptx::mbar_init(ptx::smem_u32(full_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(tfull_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(tempty_bar + i), (uint32_t)p.epi_warps); // one arrive per epilogue warp
        }
        for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(res_bar + i), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const int pad_h = p.R / 2, pad_w = p.S / 2;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const ConvTile t = decode_tile(p, tile, n_tiles_g);
                const PixelPos q0 = unflatten(p, t.p0);
                const int a_ch0 = p.in_ch_off + t.g * p.cin_g;
                const int b_row = t.g * p.cout_g_pad + t.n0;
                int kcol = 0;
                for (int r = 0; r < p.R; ++r)
                    for (int s = 0; s < p.S; ++s)
                        for (int c = 0; c < chunks; ++c, kcol += CONV_BLOCK_K) {
                            ptx::mbar_wait(ptx::smem_u32(empty_bar + stage), phase ^ 1);
                            const uint32_t fb = ptx::smem_u32(full_bar + stage);
                            uint8_t* sa = smem + (size_t)stage * stage_bytes;
                            ptx::mbar_expect_tx(fb, (uint32_t)(CONV_A_BYTES + t.bn * CONV_BLOCK_K * 2));
                            ptx::tma_load_im2col_4d(ptx::smem_u32(sa), &tmap_a, fb, a_ch0 + c * CONV_BLOCK_K,
                                                    q0.w - pad_w, q0.h - pad_h, q0.n, s, r);
                            ptx::tma_load_2d(ptx::smem_u32(sa + CONV_A_BYTES), t.bn == p.BN ? &tmap_b : &tmap_bh, fb, kcol, b_row);
                            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (ptx::elect_one()) {
            const uint32_t idesc_full = ptx::make_idesc_f16(CONV_BLOCK_M, p.BN), idesc_half = ptx::make_idesc_f16(CONV_BLOCK_M, p.BN >> 1);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint64_t da_first = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem)), stage_step = (uint64_t)(stage_bytes >> 4);
            const uint32_t full_first = ptx::smem_u32(full_bar), empty_first = ptx::smem_u32(empty_bar);
            uint64_t da_cur = da_first;
            uint32_t full_cur = full_first, empty_cur = empty_first;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(ptx::smem_u32(tempty_bar + acc), acc_phase ^ 1); // epilogue drained this accumulator
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                const uint32_t idesc = tile < p.split_from ? idesc_full : idesc_half;
                for (int ks = 0; ks < ksteps; ++ks) {
                    // descriptors and barrier addresses of the stage are ready before the wait (advanced incrementally: the stage buffers are
                    // 1024-byte multiples); only the MMAs sit between the arrival of the data and their issue
                    const uint64_t da = da_cur, db = da_cur + (uint64_t)(CONV_A_BYTES >> 4);
                    const uint32_t fb = full_cur, eb = empty_cur;
                    ptx::mbar_wait(fb, phase);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int k = 0; k < CONV_BLOCK_K / CONV_UMMA_K; ++k) {
                        // advancing K by 16 fp16 = 32 bytes inside the 128-byte swizzled row: +2 in the (>>4) address field
                        ptx::umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(eb); // frees the smem slot when these MMAs retire
                    da_cur += stage_step; full_cur += 8u; empty_cur += 8u;
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1; da_cur = da_first; full_cur = full_first; empty_cur = empty_first; }
                }
                ptx::umma_commit(ptx::smem_u32(tfull_bar + acc)); // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 4 + p.epi_warps) {
        // ===================== epilogue: TMEM -> registers -> global =====================
        // A lone epilogue warp per scheduler cannot hide its own latencies (tcgen05.ld -> wait -> loads -> stores: ~0.16 IPC measured on
        // the residual 1x1 layers), so with epi_warps == 8 every TMEM lane quarter has TWO warps, each taking half of the channels.
        const int ew = (warp - 4) & 3;            // == warp % 4: the TMEM lane quarter this warp may access
        const int eh = (warp - 4) >> 2;           // which half of every 64-channel sub-tile (always 0 with 4 epilogue warps)
        const int q_step = p.epi_warps == 8 ? 2 : 4, epi_threads = 32 * p.epi_warps;
        const int row = ew * 32 + lane;           // accumulator row == pixel index inside the tile
        const int total_px = p.Nb * p.H * p.W;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t stage_ctr = 0;
        // residual tiles: sub-tile k of this CTA's (tile, sub) sequence lands in res_stage[k & 1]; the leader keeps two in flight
        const bool res_tma = kRes && p.res_mode != 0 && p.tma_store != 0;
        uint32_t res_issued = 0, res_used = 0;
        const uint32_t res_mask = p.res_stages == 4 ? 3u : 1u, res_shift = p.res_stages == 4 ? 2u : 1u;
        int ri_tile = blockIdx.x, ri_sub = 0;
        ConvTile rt = decode_tile(p, min(ri_tile, total_tiles - 1), n_tiles_g);   // tile of the next residual load (decoded once per tile: the
        auto issue_residual = [&]() { // leader only                                  // leader issues between two barriers the other warps wait at)
            if (ri_tile >= total_tiles) return;
            const uint32_t rb = ptx::smem_u32(res_bar + (res_issued & res_mask));
            ptx::mbar_expect_tx(rb, (uint32_t)CONV_A_BYTES);
            ptx::tma_load_2d(ptx::smem_u32(res_stage + (res_issued & res_mask) * CONV_A_BYTES), &tmap_r, rb,
                             p.res_ch_off + rt.g * p.cout_g + rt.n0 + ri_sub * 64, rt.p0);
            ++res_issued;
            if (++ri_sub == rt.bn / 64) {
                ri_sub = 0; ri_tile += gridDim.x;
                if (ri_tile < total_tiles) rt = decode_tile(p, ri_tile, n_tiles_g);
            }
        };
        if (kRes && res_tma && warp == 4 && lane == 0)
            for (int i = 0; i <= (int)res_mask; ++i) issue_residual();
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const ConvTile t = decode_tile(p, tile, n_tiles_g);
            const bool in_img = (t.p0 + row) < total_px;
            const PixelPos q = unflatten(p, in_img ? t.p0 + row : 0);
            const int h = q.h, w = q.w;
            ptx::mbar_wait(ptx::smem_u32(tfull_bar + acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.BN);
            const float* bias = p.bias + t.g * p.cout_g_pad + t.n0;
            const float* alpha = p.alpha + t.g * p.cout_g_pad + t.n0;
            const int n_valid = min(t.bn, p.cout_g - t.n0); // real (unpadded) channels of this tile
            const size_t pix = (size_t)(t.p0 + row);
            if (p.tma_store) {
                // 64 channels at a time: registers -> swizzled smem tile -> one TMA tensor store (coalesced, clipped
                // at the image border by the TMA unit); double-buffered so the store of sub-tile k overlaps the
                // TMEM reads of sub-tile k+1.
                const bool leader = (warp == 4 && lane == 0);
                for (int sub = 0; sub < t.bn / 64; ++sub, ++stage_ctr) {
                    uint8_t* sbuf = out_stage + (stage_ctr & 1) * CONV_A_BYTES;
                    // Two-barrier form: the leader waits until the store that last used this buffer has drained it, everybody syncs, fills
                    // the buffer, syncs again, the leader stores.  One-barrier form (epi_one_bar): the leader drains ALL earlier stores just
                    // before the second barrier of the PREVIOUS sub-tile (they were issued a whole sub-tile ago, so this rarely waits); past that
                    // barrier every warp knows the other buffer is free, and the leader's serial work after it (store, commit, next residual
                    // load) overlaps the other warps' next sub-tile instead of holding them at a barrier.
                    if (!p.epi_one_bar) {
                        if (leader) ptx::bulk_wait_group_read<1>();
                        ptx::named_bar_sync(1, epi_threads);
                    }
                    const uint32_t srow = ptx::smem_u32(sbuf) + (uint32_t)row * 128u;
                    const uint32_t rrow = ptx::smem_u32(res_stage) + (res_used & res_mask) * (uint32_t)CONV_A_BYTES + (uint32_t)row * 128u;
                    if (kRes && res_tma) ptx::mbar_wait(ptx::smem_u32(res_bar + (res_used & res_mask)), (res_used >> res_shift) & 1);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {   // 16 channels per TMEM load; the other loads of the step are issued before the wait
                        if (qq >= q_step) break;
                        const int q = eh * q_step + qq;
                        const int c0 = sub * 64 + q * 16;
                        uint32_t v[16];
                        ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
                        __half2 rs[8];
                        if (kRes && res_tma) { // the residual tile sits in smem in the same 128B-swizzled layout as the output tile
                            *(uint4*)&rs[0] = ptx::ld_shared_v4(rrow + (uint32_t)(((q * 2) ^ (row & 7)) * 16));
                            *(uint4*)&rs[4] = ptx::ld_shared_v4(rrow + (uint32_t)(((q * 2 + 1) ^ (row & 7)) * 16));
                        } else if (kRes && p.res_mode) { // 16 residual channels of this pixel: two 16-byte loads
                            if (in_img) {
                                const uint4* rp = (const uint4*)(p.res + pix * p.res_ld + p.res_ch_off + t.g * p.cout_g + t.n0 + c0);
                                *(uint4*)&rs[0] = __ldg(rp);
                                *(uint4*)&rs[4] = __ldg(rp + 1);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) rs[j] = __floats2half2_rn(0.f, 0.f);
                            }
                        }
                        // bias / PReLU slope of these 16 channels (the tile origin is a multiple of 64 channels)
                        float bv[16], av[16];
#pragma unroll
                        for (int j = 0; j < 4; ++j) *(float4*)&bv[4 * j] = __ldg((const float4*)(bias + c0) + j);
                        if (!p.relu_only) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) *(float4*)&av[4 * j] = __ldg((const float4*)(alpha + c0) + j);
                        }
                        ptx::tmem_ld_wait();
                        uint32_t pk[8];
                        conv_epilogue16<kRes>(v, bv, av, p.relu_only != 0, p.res_mode, rs, pk);
                        if (p.post_w) {
                            const int pc = t.g * p.cout_g + t.n0 + c0;
                            conv_post16(pk, p.post_w + pc, p.post_b + pc, p.post_a + pc);
                        }
                        const int ch0 = q * 2; // 16-byte chunk index inside the 128-byte row
                        ptx::st_shared_v4(srow + (uint32_t)(((ch0) ^ (row & 7)) * 16), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                        ptx::st_shared_v4(srow + (uint32_t)(((ch0 + 1) ^ (row & 7)) * 16), make_uint4(pk[4], pk[5], pk[6], pk[7]));
                    }
                    ptx::fence_proxy_async(); // generic-proxy smem writes -> visible to the TMA (async proxy)
                    if (p.epi_one_bar && leader) ptx::bulk_wait_group_read<0>(); // the other staging buffer is free for the next sub-tile
                    ptx::named_bar_sync(1, epi_threads);
                    if (leader) {
                        ptx::tma_store_2d(&tmap_o, ptx::smem_u32(sbuf), p.out_ch_off + t.g * p.cout_g + t.n0 + sub * 64, t.p0);
                        ptx::bulk_commit_group();
                        if (kRes && res_tma) issue_residual(); // everybody is past the barrier: this sub-tile's residual buffer is free again
                    }
                    ++res_used;
                }
            } else
            for (int c0 = eh * 16; c0 < t.bn; c0 += (p.epi_warps == 8 ? 32 : 16)) {
                if (c0 >= n_valid) break; // warp-uniform: the remaining columns are padding
                uint32_t v[16];
                ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v); // warp-collective: executed by all lanes
                ptx::tmem_ld_wait();
                if (!in_img) continue;
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float a = __uint_as_float(v[j]) + __ldg(bias + c0 + j);
                    float r = 0.f;
                    if (kRes && p.res_mode && c0 + j < n_valid) r = __half2float(p.res[pix * p.res_ld + p.res_ch_off + t.g * p.cout_g + t.n0 + c0 + j]);
                    if (kRes && p.res_mode == 1) a += r;
                    a = a > 0.f ? a : a * __ldg(alpha + c0 + j);
                    if (kRes && p.res_mode == 2) a += r;
                    y[j] = a;
                }
                if (p.out_mode == OUT_F16_NHWC) {
                    __half* o = (__half*)p.out + pix * p.out_ld + p.out_ch_off + t.g * p.cout_g + t.n0 + c0;
                    const int nv = min(16, n_valid - c0);
                    if (nv == 16 && ((uintptr_t)o & 15) == 0) {
                        uint4 q0, q1;
                        __half2 h2;
                        h2 = __floats2half2_rn(y[0], y[1]);   q0.x = *(uint32_t*)&h2;
                        h2 = __floats2half2_rn(y[2], y[3]);   q0.y = *(uint32_t*)&h2;
                        h2 = __floats2half2_rn(y[4], y[5]);   q0.z = *(uint32_t*)&h2;
                        h2 = __floats2half2_rn(y[6], y[7]);   q0.w = *(uint32_t*)&h2;
                        h2 = __floats2half2_rn(y[8], y[9]);   q1.x = *(uint32_t*)&h2;
                        h2 = __floats2half2_rn(y[10], y[11]); q1.y = *(uint32_t*)&h2;
                        h2 = __floats2half2_rn(y[12], y[13]); q1.z = *(uint32_t*)&h2;
                        h2 = __floats2half2_rn(y[14], y[15]); q1.w = *(uint32_t*)&h2;
                        ((uint4*)o)[0] = q0;
                        ((uint4*)o)[1] = q1;
                    } else {
                        for (int j = 0; j < nv; ++j) o[j] = __float2half_rn(y[j]);
                    }
                } else {
                    const int nv = min(16, n_valid - c0);
                    for (int j = 0; j < nv; ++j) {
                        const int ch = t.n0 + c0 + j;
                        if (ch < p.split) {
                            ((float*)p.out)[(((size_t)q.n * p.split + ch) * p.H + h) * p.W + w] = y[j];
                        } else {
                            const int c2 = ch - p.split, n2 = p.cout_g - p.split;
                            ((float*)p.out2)[(((size_t)q.n * n2 + c2) * p.H + h) * p.W + w] = y[j];
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(tempty_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.tma_store && warp == 4 && lane == 0) ptx::bulk_wait_group_read<0>(); // smem must outlive the last stores
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// Swapped-operand variant for layers whose output channels come in blocks of 128 and whose k-loop is long
// (the 7x7 refinement convs: 68 % of the model's FLOPs, and the 3x3 init convs).
//   D^T[128 channels, NPX pixels] += W[128 ch, 64 k] * X[NPX px, 64 k]^T        (UMMA M = 128, N = NPX <= 256)
// Why: (1) with M = 128 pixels x N = 128 channels one MMA reads 4 + 4 KiB of shared memory per 64 tensor cycles,
// the full 128 B/clk (ncu: sm__mem_tensor_cycles_active 78 %); with N = NPX >= 192 pixels it is <= 107 B/clk.
// (2) NPX is free (any multiple of 16), so the unit size is chosen per layer to fill the last wave of the
// persistent grid: 46x82x16 pixels x 2 groups at NPX = 256 is 472 units = 3.19 waves of 148 CTAs (4 rounds),
// at NPX = 208 it is 582 units = 3.93 waves (4 rounds of a 19 % smaller unit).
// One im2col TMA load brings the NPX consecutive pixels of a unit.  TMEM lanes are channels, columns are pixels;
// the epilogue transposes through the swizzled staging tiles (one 2-byte shared store per value) and issues
// TMA stores of rows [0,128) and [128,NPX) (second tensor map with an NPX-128 row box).
// ---------------------------------------------------------------------------------------------
// kMC (weight multicast): launched as clusters of TWO CTAs that walk the same (group, channel block) over two neighbouring pixel
// units in lockstep.  The 16 KiB weight tile of a k-step is the same for both, so each CTA loads HALF of it (64 rows) and
// multicasts that half into both CTAs' stage; per CTA and k-step the L2 -> SM traffic drops from 16 + NPX/8 KiB to 8 + NPX/8 KiB
// (42.6 -> 34.6 KiB at NPX = 208) -- and that port, not the tensor pipe, is what bounds these layers (profiles/r01_s2_layer_bounds.md).
// A stage may be refilled only when BOTH CTAs' MMAs have released it: the empty barriers count two arrivals, and every
// tcgen05.commit arrives on the barrier of both CTAs.  tmap_bh = the weight map with a 64-row box.
template <bool kMC>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tcgen05_swap_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o2,
                         const __grid_constant__ CUtensorMap tmap_bh, const ConvParams p)
{
    extern __shared__ uint8_t smem_raw[];
    ptx::pdl_launch_dependents();   // the next kernel may start its prologue on every SM this grid has left
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int W_BYTES = 128 * CONV_BLOCK_K * 2;          // 16 KiB weight tile (128 channels x 64 k)
    const int npx = p.npx;
    const int stage_bytes = W_BYTES + npx * 128;             // + NPX pixel rows of 128 B
    uint8_t* bar_base = smem + (size_t)p.num_stages * stage_bytes;
    uint64_t* full_bar = (uint64_t*)bar_base;
    uint64_t* empty_bar = full_bar + CONV_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + CONV_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
    uint8_t* out_stage = (uint8_t*)(((uintptr_t)(tmem_slot + 4) + 1023) & ~(uintptr_t)1023); // 2 x 16 KiB

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_px = p.Nb * p.H * p.W;
    const int units = (total_px + npx - 1) / npx;
    const int c_tiles = p.cout_g_pad / 128;                  // 128-channel blocks per group
    const int gc = p.groups * c_tiles;
    const int chunks = p.cin_g / CONV_BLOCK_K;
    const int ksteps = p.R * p.S * chunks;
    // work list.  plain: tile = (unit, sub) with sub fastest, CTA b takes tiles b, b + grid, ...
    //            kMC  : item = (unit PAIR, sub) with sub fastest, cluster c takes items c, c + clusters, ...; CTA rank r of the
    //                   cluster works on unit 2 * pair + r (a unit past the end computes on TMA zero-fill and stores nothing)
    const uint32_t crank = kMC ? ptx::cluster_ctarank() : 0u;
    const int my_first = kMC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int my_step = kMC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int total_items = kMC ? ((units + 1) / 2) * gc : units * gc;
    auto decode = [&](int item, int& unit, int& sub) {
        sub = item % gc;
        unit = kMC ? 2 * (item / gc) + (int)crank : item / gc;
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_o);
        ptx::prefetch_tmap(&tmap_o2);
        if (kMC) ptx::prefetch_tmap(&tmap_bh);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.num_stages; ++i) {
            ptx::mbar_init(ptx::smem_u32(full_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(empty_bar + i), kMC ? 2 : 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(tfull_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(tempty_bar + i), 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512u);
    ptx::tc_fence_before();
    __syncthreads();
    if (kMC) ptx::cluster_sync();   // the peer's barriers exist before anything is multicast into them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

    if (warp == 0) {
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const int pad_h = p.R / 2, pad_w = p.S / 2;
            for (int item = my_first; item < total_items; item += my_step) {
                int unit, sub;
                decode(item, unit, sub);
                const int g = sub / c_tiles, ct = sub - g * c_tiles;
                // pixels past the last image are zero-filled by the TMA unit and never stored
                const PixelPos q0 = unflatten(p, unit * npx);
                const int a_ch0 = p.in_ch_off + g * p.cin_g;
                const int b_row = g * p.cout_g_pad + ct * 128;
                int kcol = 0;
                for (int r = 0; r < p.R; ++r)
                    for (int s = 0; s < p.S; ++s)
                        for (int c = 0; c < chunks; ++c, kcol += CONV_BLOCK_K) {
                            ptx::mbar_wait(ptx::smem_u32(empty_bar + stage), phase ^ 1);
                            const uint32_t fb = ptx::smem_u32(full_bar + stage);
                            const uint32_t sw = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
                            ptx::mbar_expect_tx(fb, (uint32_t)stage_bytes);
                            if (kMC) ptx::tma_load_2d_mc(sw + crank * (W_BYTES / 2), &tmap_bh, fb, kcol, b_row + (int)crank * 64, (uint16_t)3);
                            else ptx::tma_load_2d(sw, &tmap_b, fb, kcol, b_row);
                            ptx::tma_load_im2col_4d(sw + W_BYTES, &tmap_a, fb, a_ch0 + c * CONV_BLOCK_K, q0.w - pad_w, q0.h - pad_h, q0.n, s, r);
                            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_f16(128, npx);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            // stage 0's weight-tile descriptor / barrier addresses and the per-stage increments (the stage buffers are 1024-byte multiples,
            // so the descriptor's address field just advances by stage_bytes >> 4; the pixel tile follows the weight tile by W_BYTES)
            const uint64_t dw_first = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem)), stage_step = (uint64_t)(stage_bytes >> 4);
            const uint32_t full_first = ptx::smem_u32(full_bar), empty_first = ptx::smem_u32(empty_bar);
            uint64_t dw_cur = dw_first;
            uint32_t full_cur = full_first, empty_cur = empty_first;
            for (int item = my_first; item < total_items; item += my_step) {
                ptx::mbar_wait(ptx::smem_u32(tempty_bar + acc), acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int ks = 0; ks < ksteps; ++ks) {
                    // the descriptors of this stage are ready BEFORE the wait: nothing but the four MMAs sits between the arrival of the
                    // data and their issue (they were computed after the wait: ~30 dependent uniform-datapath instructions per k-step)
                    const uint64_t dw = dw_cur, dx = dw_cur + (uint64_t)(W_BYTES >> 4);
                    const uint32_t fb = full_cur, eb = empty_cur;
                    ptx::mbar_wait(fb, phase);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int k = 0; k < CONV_BLOCK_K / CONV_UMMA_K; ++k)
                        ptx::umma_f16(d_tmem, dw + (uint64_t)(k * 2), dx + (uint64_t)(k * 2), idesc, (ks | k) != 0 ? 1u : 0u);
                    if (kMC) ptx::umma_commit_mc(eb, (uint16_t)3);   // frees the slot in BOTH CTAs' view
                    else ptx::umma_commit(eb);
                    dw_cur += stage_step; full_cur += 8u; empty_cur += 8u;
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1; dw_cur = dw_first; full_cur = full_first; empty_cur = empty_first; }
                }
                ptx::umma_commit(ptx::smem_u32(tfull_bar + acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int ch = ew * 32 + lane;                 // TMEM lane == output channel inside the 128-channel block
        const bool leader = (warp == 4 && lane == 0);
        uint8_t* sbuf = out_stage + (ch >> 6) * CONV_A_BYTES; // staging tile of this channel's 64-channel half
        const uint32_t sbase = ptx::smem_u32(sbuf) + (uint32_t)((ch & 7) * 2);
        const int chunk = (ch & 63) >> 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = my_first; item < total_items; item += my_step) {
            int unit, sub;
            decode(item, unit, sub);
            const int g = sub / c_tiles, ct = sub - g * c_tiles;
            const float bias = __ldg(p.bias + g * p.cout_g_pad + ct * 128 + ch);
            const float alpha = __ldg(p.alpha + g * p.cout_g_pad + ct * 128 + ch);
            ptx::mbar_wait(ptx::smem_u32(tfull_bar + acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * 256);
            for (int half = 0; half * 128 < npx; ++half) {
                const int rows = min(128, npx - half * 128);
                const int p0 = unit * npx + half * 128;
                if (leader) ptx::bulk_wait_group_read<0>(); // previous stores have drained both staging tiles
                ptx::named_bar_sync(1, 128);
                for (int q = 0; q * 16 < rows; ++q) {
                    uint32_t v[16];
                    ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)(half * 128 + q * 16), v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int px = q * 16 + j;
                        float a = __uint_as_float(v[j]) + bias;
                        a = a > 0.f ? a : a * alpha;
                        const __half hv = __float2half_rn(a);
                        const uint32_t addr = sbase + (uint32_t)(px * 128 + ((chunk ^ (px & 7)) * 16));
                        asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*(const unsigned short*)&hv) : "memory");
                    }
                }
                ptx::fence_proxy_async();
                ptx::named_bar_sync(1, 128);
                if (leader && p0 < total_px) {
                    const int c0 = p.out_ch_off + g * p.cout_g + ct * 128;
                    const CUtensorMap* tm = (rows == 128) ? &tmap_o : &tmap_o2;
                    ptx::tma_store_2d(tm, ptx::smem_u32(out_stage), c0, p0);
                    ptx::tma_store_2d(tm, ptx::smem_u32(out_stage + CONV_A_BYTES), c0 + 64, p0);
                    ptx::bulk_commit_group();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(tempty_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (leader) ptx::bulk_wait_group_read<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (kMC) ptx::cluster_sync();   // nobody leaves while the peer may still multicast into this CTA or arrive on its barriers
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512u);
    }
}

// ---------------------------------------------------------------------------------------------
// Halo-box variant for RxS convolutions on large images (the early VGG layers).
// The im2col-mode kernels above fetch the A operand once PER FILTER TAP: a 3x3 layer pulls every input pixel nine times
// from L2 into shared memory, and ncu shows those layers pinned at the L2 output limit (lts2xbar ~11 TB/s, xbar->SM
// 62 B/clk/SM) with the MMA issuer waiting on the full barrier half of the time (profiles/r01_s2_ncu_conv1_2_im2col.txt).
// Here a work item is a SPATIAL tile of 16 rows x 8 columns of output pixels (= the 128 UMMA rows).  Per 64-channel
// chunk ONE tiled-mode TMA load brings the (16+R-1) x (8+S-1) pixel halo box (out-of-image pixels zero-filled by the
// TMA unit = "SAME" padding) into a 128B-swizzled buffer, and every filter tap (r, s) multiplies straight out of that
// buffer: its A descriptor is the box address + (r * box_width + s) * 128 B, with a stride-byte-offset of one box row
// (box_width * 128 B) between the 8-pixel row groups -- the 128B swizzle is a function of the absolute shared-memory
// address, so a start address that is not 1024-byte aligned reads back exactly what the TMA wrote (verified on
// hardware: tools/probe_halo.cu, bit-exact for interior, corner and edge tiles).  A traffic drops from R*S x 16 KiB
// to one 22.5 KiB box per chunk (3x3: 6.4x less); the weights stream through their own ring, one BN x 64 tile per
// (tap, chunk).  Everything else -- TMEM double buffering, warp roles, bias + PReLU epilogue through swizzled staging
// tiles -- is as in conv_tcgen05_kernel; the output goes out as a 4-D TMA box {64 ch, 8, 16, 1} that the TMA unit
// clips at the image border.  Used when the tile grid wastes little of the image (engine.cu: halo_eligible).
// ---------------------------------------------------------------------------------------------
constexpr int HALO_TH = 16, HALO_TW = 8;
constexpr int HALO_MAX_BOXES = 3;

struct HaloParams {
    int Nb, H, W;
    int R, S, groups, cin_g;
    int cout_g, cout_g_pad, BN;
    int in_ch_off, out_ch_off;
    int tiles_x, tiles_y;
    int num_boxes, box_bytes;   // box_bytes: (16+R-1) * (8+S-1) * 128 rounded up to a multiple of 1024
    int num_b_stages;           // weight ring depth; in resident mode: R * S * chunks tiles, loaded once
    int b_resident;             // 1: the layer's whole weight matrix (one group, one n-tile) stays in shared memory
    int tmem_cols;
    const float* bias; const float* alpha;
    int relu_only;              // 1: every slope of the layer is 0 (plain ReLU)
    int epi_warps;              // 8: two epilogue warps per TMEM lane quarter (each takes half of the channels), 4: one
};

// kPool: the 2x2 / stride-2 max-pool that follows the layer is taken in the epilogue (VGG conv1_2 -> maxpool_1: the un-pooled
// 494 MB activation is never written and the pool kernel disappears).  A tile is 16 x 8 pixels with TMEM lane j = pixel (j >> 3, j & 7),
// so a 2x2 window is the lanes j, j^1, j^8, j^9 of ONE warp.  The maximum is taken on the raw fp32 accumulators BEFORE bias /
// activation / rounding -- all three are monotone (slopes >= 0 are required), so fp16(act(max(v) + b)) == max(fp16(act(v + b))) bit for
// bit -- by a transposing butterfly: in the x step a lane keeps 8 of its 16 channels and trades the other 8 with its partner, in the y
// step 4 of 8, so afterwards EVERY lane holds 4 channels of one pooled pixel and none is idle.  The pooled tile (8 x 4 pixels x 64
// channels) goes out as one TMA box {64, 4, 8, 1} of the pooled tensor.
template <int KR, bool kPool = false> // KR > 0: R == S == KR, tap loops unrolled (descriptor offsets become immediates); 0: run-time R, S
__global__ void __launch_bounds__(CONV_IM2COL_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_o, const HaloParams p)
{
    extern __shared__ uint8_t smem_raw[];
    ptx::pdl_launch_dependents();   // the next kernel may start its prologue on every SM this grid has left
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.BN * CONV_BLOCK_K * 2;
    uint8_t* s_box = smem;                                              // [num_boxes][box_bytes]
    uint8_t* s_b = s_box + (size_t)p.num_boxes * p.box_bytes;           // [num_b_stages][BN x 128 B]
    uint8_t* out_stage = s_b + (size_t)p.num_b_stages * b_bytes;        // 2 x 16 KiB (1024-aligned: all sizes are multiples of 1024)
    uint64_t* a_full = (uint64_t*)(out_stage + 2 * CONV_A_BYTES);       // [HALO_MAX_BOXES]
    uint64_t* a_empty = a_full + HALO_MAX_BOXES;
    uint64_t* b_full = a_empty + HALO_MAX_BOXES;                        // [CONV_MAX_STAGES]
    uint64_t* b_empty = b_full + CONV_MAX_STAGES;
    uint64_t* tfull_bar = b_empty + CONV_MAX_STAGES;                    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                               // [2]
    uint64_t* w_bar = tempty_bar + 2;                                   // resident weights landed
    uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_g = p.cout_g_pad / p.BN;
    const int nt_total = p.groups * n_tiles_g;
    const int sp_per_img = p.tiles_x * p.tiles_y;
    const int total_items = p.Nb * sp_per_img * nt_total;
    const int chunks = p.cin_g / CONV_BLOCK_K;
    const int taps = p.R * p.S;
    const int BW = KR > 0 ? HALO_TW + KR - 1 : HALO_TW + p.S - 1;
    const int pad_h = p.R / 2, pad_w = p.S / 2;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_x);
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_o);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.num_boxes; ++i) {
            ptx::mbar_init(ptx::smem_u32(a_full + i), 1);
            ptx::mbar_init(ptx::smem_u32(a_empty + i), 1);
        }
        if (!p.b_resident)
            for (int i = 0; i < p.num_b_stages; ++i) {
                ptx::mbar_init(ptx::smem_u32(b_full + i), 1);
                ptx::mbar_init(ptx::smem_u32(b_empty + i), 1);
            }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(tfull_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(tempty_bar + i), (uint32_t)p.epi_warps);
        }
        ptx::mbar_init(ptx::smem_u32(w_bar), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

    // item -> (spatial tile, n-tile); n-tiles vary fastest so that neighbouring CTAs share a halo box in L2
    auto decode = [&](int item, int& n, int& y0, int& x0, int& g, int& n0) {
        const int nt = item % nt_total, sp = item / nt_total;
        g = nt / n_tiles_g;
        n0 = (nt - g * n_tiles_g) * p.BN;
        n = sp / sp_per_img;
        const int t = sp - n * sp_per_img;
        const int ty = t / p.tiles_x;
        y0 = ty * HALO_TH;
        x0 = (t - ty * p.tiles_x) * HALO_TW;
    };

    if (warp == 0) {
        // ===================== TMA producer: halo boxes (one per item x chunk) and weight tiles (one per tap) =====================
        if (ptx::elect_one()) {
            int box = 0; uint32_t box_phase = 0;       // next box slot to fill
            int st = 0; uint32_t st_phase = 0;
            // the box of A-step i+1 is requested before the weight tiles of A-step i, so it is in flight a whole chunk early
            int nx_item = blockIdx.x, nx_chunk = 0;
            auto issue_box = [&]() {
                if (nx_item >= total_items) return;
                int n, y0, x0, g, n0;
                decode(nx_item, n, y0, x0, g, n0);
                ptx::mbar_wait(ptx::smem_u32(a_empty + box), box_phase ^ 1);
                const uint32_t fb = ptx::smem_u32(a_full + box);
                ptx::mbar_expect_tx(fb, (uint32_t)((HALO_TH + p.R - 1) * BW * 128));
                ptx::tma_load_4d(ptx::smem_u32(s_box + (size_t)box * p.box_bytes), &tmap_x, fb,
                                 p.in_ch_off + g * p.cin_g + nx_chunk * CONV_BLOCK_K, x0 - pad_w, y0 - pad_h, n);
                if (++box == p.num_boxes) { box = 0; box_phase ^= 1; }
                if (++nx_chunk == chunks) { nx_chunk = 0; nx_item += gridDim.x; }
            };
            if (p.b_resident) { // tile (t, c) of the weight matrix at s_b + (t * chunks + c) * b_bytes, all on one barrier
                ptx::mbar_expect_tx(ptx::smem_u32(w_bar), (uint32_t)(taps * chunks * b_bytes));
                for (int t = 0; t < taps; ++t)
                    for (int c = 0; c < chunks; ++c)
                        ptx::tma_load_2d(ptx::smem_u32(s_b + (size_t)(t * chunks + c) * b_bytes), &tmap_b, ptx::smem_u32(w_bar), t * p.cin_g + c * CONV_BLOCK_K, 0);
            }
            issue_box();
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int n, y0, x0, g, n0;
                decode(item, n, y0, x0, g, n0);
                const int b_row = g * p.cout_g_pad + n0;
                for (int c = 0; c < chunks; ++c) {
                    issue_box(); // look-ahead: the next A-step's box
                    if (p.b_resident) continue;
                    for (int t = 0; t < taps; ++t) {
                        ptx::mbar_wait(ptx::smem_u32(b_empty + st), st_phase ^ 1);
                        const uint32_t fb = ptx::smem_u32(b_full + st);
                        ptx::mbar_expect_tx(fb, (uint32_t)b_bytes);
                        ptx::tma_load_2d(ptx::smem_u32(s_b + (size_t)st * b_bytes), &tmap_b, fb, t * p.cin_g + c * CONV_BLOCK_K, b_row);
                        if (++st == p.num_b_stages) { st = 0; st_phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // This one thread's instruction stream bounds the small-N layers (ncu on conv1_2: ~44 scalar instructions per k-step at
        // ~10 cycles each against 128 cycles of tensor work), so with KR > 0 the tap loops are unrolled: every descriptor is a
        // hoisted 64-bit base plus an immediate, two UIADD3.64 + one UTCHMMA per MMA.
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_f16(CONV_BLOCK_M, p.BN);
            const uint64_t a_hi = ((uint64_t)((uint32_t)(BW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            const uint64_t db0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(s_b));
            const uint64_t b_step = (uint64_t)(b_bytes >> 4);               // one weight tile, in descriptor address units
            const uint64_t tap_step = b_step * (uint64_t)chunks;            // resident layout: tile (t, c) at (t * chunks + c)
            // resident / streamed weights are two instantiations of the loop (the per-tap test of a run-time flag kept the streamed path's
            // barrier wait and commit in the resident stream: ~45 instructions per tap on a 9-tap tile that is paced by exactly this thread)
            auto issue = [&](auto resident_tag) {
                constexpr bool kResident = decltype(resident_tag)::value;
                int box = 0; uint32_t box_phase = 0;
                int st = 0; uint32_t st_phase = 0;
                int acc = 0; uint32_t acc_phase = 0;
                if (kResident) ptx::mbar_wait(ptx::smem_u32(w_bar), 0);
                for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                    ptx::mbar_wait(ptx::smem_u32(tempty_bar + acc), acc_phase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                    for (int c = 0; c < chunks; ++c) {
                        const uint32_t box_addr = ptx::smem_u32(s_box + (size_t)box * p.box_bytes);
                        // A: rows = the 16 x 8 pixels at box offset (r, s2); K-major, 128B swizzle, row groups one box row apart
                        const uint64_t da0 = (uint64_t)((box_addr & 0x3ffffu) >> 4) | a_hi;
                        uint64_t db_res = db0 + b_step * (uint64_t)c;
                        ptx::mbar_wait(ptx::smem_u32(a_full + box), box_phase);
                        ptx::tc_fence_after();
                        auto tap = [&](int r, int s2, bool first) {
                            uint64_t db;
                            if (kResident) { db = db_res; db_res += tap_step; }
                            else {
                                ptx::mbar_wait(ptx::smem_u32(b_full + st), st_phase);
                                ptx::tc_fence_after();
                                db = db0 + b_step * (uint64_t)st;
                            }
                            const uint64_t da = da0 + (uint64_t)((r * BW + s2) * 8);
#pragma unroll
                            for (int k = 0; k < CONV_BLOCK_K / CONV_UMMA_K; ++k)
                                ptx::umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (first && k == 0) ? 0u : 1u);
                            if (!kResident) {
                                ptx::umma_commit(ptx::smem_u32(b_empty + st));
                                if (++st == p.num_b_stages) { st = 0; st_phase ^= 1; }
                            }
                        };
                        if (KR > 0) {
#pragma unroll
                            for (int r = 0; r < (KR > 0 ? KR : 1); ++r)
#pragma unroll
                                for (int s2 = 0; s2 < (KR > 0 ? KR : 1); ++s2) tap(r, s2, c == 0 && r == 0 && s2 == 0);
                        } else {
                            for (int r = 0; r < p.R; ++r)
                                for (int s2 = 0; s2 < p.S; ++s2) tap(r, s2, c == 0 && r == 0 && s2 == 0);
                        }
                        ptx::umma_commit(ptx::smem_u32(a_empty + box)); // every tap of this chunk has read the box
                        if (++box == p.num_boxes) { box = 0; box_phase ^= 1; }
                    }
                    ptx::umma_commit(ptx::smem_u32(tfull_bar + acc));
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            };
            if (p.b_resident) issue(std::true_type{}); else issue(std::false_type{});
        }
    } else if (warp >= 4 && warp < 4 + p.epi_warps) {
        // ===================== epilogue: TMEM lane j = pixel (j >> 3, j & 7) of the tile =====================
        // (epi_warps == 8: two warps per lane quarter, each converting half of the channels -- a 9-k-step tile is paced by its epilogue)
        const int ew = (warp - 4) & 3, eh = (warp - 4) >> 2, row = ew * 32 + lane;
        const int q_step = p.epi_warps == 8 ? 2 : 4, epi_threads = 32 * p.epi_warps;
        const bool leader = (warp == 4 && lane == 0);
        int acc = 0; uint32_t acc_phase = 0, stage_ctr = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            int n, y0, x0, g, n0;
            decode(item, n, y0, x0, g, n0);
            ptx::mbar_wait(ptx::smem_u32(tfull_bar + acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.BN);
            const float* bias = p.bias + g * p.cout_g_pad + n0;
            const float* alpha = p.alpha + g * p.cout_g_pad + n0;
            if (kPool) {
                const int lx = lane & 7, ly = lane >> 3;                       // pixel of this lane inside the warp's 4 x 8 strip
                const bool xo = (lx & 1) != 0, yo = (ly & 1) != 0;
                const int cbase = (xo ? 8 : 0) + (yo ? 4 : 0);                  // the 4 channels (of every 16) this lane ends up with
                const int row_p = ((ew * 4 + ly) >> 1) * 4 + (lx >> 1);         // pooled pixel inside the 8 x 4 pooled tile
                for (int sub = 0; sub < p.BN / 64; ++sub, ++stage_ctr) {
                    uint8_t* sbuf = out_stage + (stage_ctr & 1) * CONV_A_BYTES;
                    if (leader) ptx::bulk_wait_group_read<1>();
                    ptx::named_bar_sync(1, epi_threads);
                    const uint32_t srow = ptx::smem_u32(sbuf) + (uint32_t)row_p * 128u;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        if (qq >= q_step) break;
                        const int q = eh * q_step + qq;
                        const int c0 = sub * 64 + q * 16;
                        uint32_t v[16];
                        ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
                        const float4 bv = __ldg((const float4*)(bias + c0 + cbase));
                        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (!p.relu_only) av = __ldg((const float4*)(alpha + c0 + cbase));
                        ptx::tmem_ld_wait();
                        float m1[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {   // x step: even columns keep channels 0..7, odd columns 8..15
                            const float send = xo ? __uint_as_float(v[i]) : __uint_as_float(v[8 + i]);
                            const float keep = xo ? __uint_as_float(v[8 + i]) : __uint_as_float(v[i]);
                            m1[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                        }
                        float m2[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {   // y step: even rows keep the first 4 of those, odd rows the last 4
                            const float send = yo ? m1[i] : m1[4 + i];
                            const float keep = yo ? m1[4 + i] : m1[i];
                            m2[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
                        }
                        float a0 = m2[0] + bv.x, a1 = m2[1] + bv.y, a2 = m2[2] + bv.z, a3 = m2[3] + bv.w;
                        if (p.relu_only) {
                            a0 = __uint_as_float(__float_as_uint(fmaxf(a0, 0.f)) | (__float_as_uint(a0) & 0x80000000u));
                            a1 = __uint_as_float(__float_as_uint(fmaxf(a1, 0.f)) | (__float_as_uint(a1) & 0x80000000u));
                            a2 = __uint_as_float(__float_as_uint(fmaxf(a2, 0.f)) | (__float_as_uint(a2) & 0x80000000u));
                            a3 = __uint_as_float(__float_as_uint(fmaxf(a3, 0.f)) | (__float_as_uint(a3) & 0x80000000u));
                        } else {
                            a0 = a0 > 0.f ? a0 : a0 * av.x; a1 = a1 > 0.f ? a1 : a1 * av.y;
                            a2 = a2 > 0.f ? a2 : a2 * av.z; a3 = a3 > 0.f ? a3 : a3 * av.w;
                        }
                        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
                        const uint32_t chunk = (uint32_t)(q * 2 + (cbase >> 3));
                        const uint32_t addr = srow + ((chunk ^ (uint32_t)(row_p & 7)) * 16u) + (uint32_t)((cbase & 4) * 2);
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*(const uint32_t*)&h01), "r"(*(const uint32_t*)&h23) : "memory");
                    }
                    ptx::fence_proxy_async();
                    ptx::named_bar_sync(1, epi_threads);
                    if (leader) {
                        // box {64 ch, 4, 8, 1} of the pooled tensor; pooled pixels outside it are clipped by the TMA unit
                        ptx::tma_store_4d(&tmap_o, ptx::smem_u32(sbuf), p.out_ch_off + g * p.cout_g + n0 + sub * 64, x0 >> 1, y0 >> 1, n);
                        ptx::bulk_commit_group();
                    }
                }
            } else
            for (int sub = 0; sub < p.BN / 64; ++sub, ++stage_ctr) {
                uint8_t* sbuf = out_stage + (stage_ctr & 1) * CONV_A_BYTES;
                if (leader) ptx::bulk_wait_group_read<1>();
                ptx::named_bar_sync(1, epi_threads);
                const uint32_t srow = ptx::smem_u32(sbuf) + (uint32_t)row * 128u;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    if (qq >= q_step) break;
                    const int q = eh * q_step + qq;
                    const int c0 = sub * 64 + q * 16;
                    uint32_t v[16];
                    ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
                    float bv[16], av[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) *(float4*)&bv[4 * j] = __ldg((const float4*)(bias + c0) + j);
                    if (!p.relu_only) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) *(float4*)&av[4 * j] = __ldg((const float4*)(alpha + c0) + j);
                    }
                    ptx::tmem_ld_wait();
                    uint32_t pk[8];
                    __half2 rs[8];
                    conv_epilogue16<false>(v, bv, av, p.relu_only != 0, 0, rs, pk);
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2) ^ (row & 7)) * 16), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2 + 1) ^ (row & 7)) * 16), make_uint4(pk[4], pk[5], pk[6], pk[7]));
                }
                ptx::fence_proxy_async();
                ptx::named_bar_sync(1, epi_threads);
                if (leader) {
                    // box {64 ch, 8, 16, 1}: staging row y * 8 + x == TMEM lane; pixels outside the image are clipped by the TMA unit
                    ptx::tma_store_4d(&tmap_o, ptx::smem_u32(sbuf), p.out_ch_off + g * p.cout_g + n0 + sub * 64, x0, y0, n);
                    ptx::bulk_commit_group();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(tempty_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (leader) ptx::bulk_wait_group_read<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

inline int halo_box_bytes(int R, int S) { return ((HALO_TH + R - 1) * (HALO_TW + S - 1) * 128 + 1023) & ~1023; }
inline size_t conv_halo_smem_bytes(int R, int S, int BN, int boxes, int b_stages)
{
    return 1024 + (size_t)boxes * halo_box_bytes(R, S) + (size_t)b_stages * BN * 128 + 2 * CONV_A_BYTES + (2 * HALO_MAX_BOXES + 2 * CONV_MAX_STAGES + 5) * 8 + 16;
}
inline int conv_halo_pick_b_stages(int R, int S, int BN, int boxes)
{
    const size_t fixed = conv_halo_smem_bytes(R, S, BN, boxes, 0);
    if (fixed >= CONV_SMEM_LIMIT) return 0;
    const int st = (int)((CONV_SMEM_LIMIT - fixed) / ((size_t)BN * 128));
    return st > CONV_MAX_STAGES ? CONV_MAX_STAGES : st;
}

// ---------------------------------------------------------------------------------------------
// Stem variant: the first convolution of a network (RxR x 3 channels, R in {3,7}, stride 1/2) reading the u8 frames
// DIRECTLY.  The generic path materialises the im2col patches in HBM (128 B/pixel for 3x3: 494 MB written and read again
// per cfg3 step); here four producer warps (one thread per pixel of the tile) gather the RxRx3 bytes of their pixel,
// normalise (u8 * factor in double, R/B swap, - mean: src/data.cpp:21-51, backbones.py:455) and write the fp16 patch row
// straight into the 128B-swizzled A tile that tcgen05.mma reads -- the patches never exist outside shared memory.
// The weight tile (BN x KCH*64) is loaded once by TMA and stays resident.  Warp roles: 0-3 producers, 4-7 epilogue,
// 8 MMA issuer + TMEM allocator.  Epilogue = bias + PReLU -> fp16 -> swizzled staging -> TMA tensor store.
// ---------------------------------------------------------------------------------------------
struct StemParams {
    const uint8_t* frames; // [N, H, W, 3] u8
    int Nb, H, W;          // input frame geometry
    int OH, OW;            // output geometry (= ceil(H / stride))
    int stride, pad_h, pad_w;
    double factor; int flip;
    float m0, m1, m2;
    int BN;                // padded output channels (multiple of 64, <= 128), single N tile
    int cout;
    const float* bias; const float* alpha;
    int out_ch_off;
};

constexpr int STEM_THREADS = 288;
constexpr int STEM_STAGES = 3;

template <int R>
__global__ void __launch_bounds__(STEM_THREADS, 1)
conv_stem_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_o, const StemParams p)
{
    constexpr int KTOT = R * R * 3;
    constexpr int KCH = (KTOT + 63) / 64;           // 64-wide k chunks (1 for 3x3, 3 for 7x7)
    constexpr int A_BYTES = KCH * CONV_A_BYTES;     // one stage of patches: KCH tiles of 128 rows x 128 B
    extern __shared__ uint8_t smem_raw[];
    ptx::pdl_launch_dependents();   // the next kernel may start its prologue on every SM this grid has left
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                         // [STEM_STAGES][KCH][128 x 128 B]
    uint8_t* sB = sA + (size_t)STEM_STAGES * A_BYTES;           // [KCH][BN x 128 B] weights, resident
    uint8_t* sOut = sB + (size_t)KCH * p.BN * 128;              // 2 x 16 KiB staging
    uint64_t* full_bar = (uint64_t*)(sOut + 2 * CONV_A_BYTES);  // [STAGES] producers -> MMA (128 arrivals)
    uint64_t* empty_bar = full_bar + STEM_STAGES;               // [STAGES] MMA -> producers
    uint64_t* tfull_bar = empty_bar + STEM_STAGES;              // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                       // [2]
    uint64_t* w_bar = tempty_bar + 2;                           // weights landed
    uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_px = p.Nb * p.OH * p.OW;
    const int total_tiles = (total_px + CONV_BLOCK_M - 1) / CONV_BLOCK_M;
    const int tmem_cols = p.BN <= 64 ? 128 : 256;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_o);
        for (int i = 0; i < STEM_STAGES; ++i) {
            ptx::mbar_init(ptx::smem_u32(full_bar + i), 128);
            ptx::mbar_init(ptx::smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(tfull_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(tempty_bar + i), 4);
        }
        ptx::mbar_init(ptx::smem_u32(w_bar), 1);
        ptx::fence_barrier_init();
    }
    // zero the patch tiles once: the k slots beyond R*R*3 are never written again and must read as 0
    for (int i = threadIdx.x; i < STEM_STAGES * A_BYTES / 16; i += STEM_THREADS) ((uint4*)sA)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 8) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), (uint32_t)tmem_cols);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

    if (warp < 4) {
        // ===================== producers: one thread per pixel of the tile =====================
        const int row = threadIdx.x; // 0..127
        const float mean[3] = { p.m0, p.m1, p.m2 };
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int px = tile * CONV_BLOCK_M + row;
            const bool valid = px < total_px;
            int n = 0, oh = 0, ow = 0;
            if (valid) { n = px / (p.OH * p.OW); const int rem = px - n * p.OH * p.OW; oh = rem / p.OW; ow = rem - oh * p.OW; }
            const int h0 = oh * p.stride - p.pad_h, w0 = ow * p.stride - p.pad_w;
            // gather + normalise into registers first (global latency), then wait for the stage and store
            __align__(16) __half vals[((KTOT + 7) / 8) * 8];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int hh = h0 + r;
                const bool rok = valid && hh >= 0 && hh < p.H;
                const uint8_t* rp = p.frames + ((size_t)n * p.H + (rok ? hh : 0)) * p.W * 3;
#pragma unroll
                for (int s2 = 0; s2 < R; ++s2) {
                    const int ww = w0 + s2;
                    const bool ok = rok && ww >= 0 && ww < p.W;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float v = 0.f;
                        if (ok) v = (float)((double)__ldg(rp + (size_t)ww * 3 + (p.flip ? 2 - c : c)) * p.factor) - mean[c];
                        vals[(r * R + s2) * 3 + c] = __float2half_rn(v);
                    }
                }
            }
#pragma unroll
            for (int k = KTOT; k < ((KTOT + 7) / 8) * 8; ++k) vals[k] = __float2half_rn(0.f);
            ptx::mbar_wait(ptx::smem_u32(empty_bar + stage), phase ^ 1);
            uint8_t* a = sA + (size_t)stage * A_BYTES;
#pragma unroll
            for (int q = 0; q < (KTOT + 7) / 8; ++q) { // 16-byte chunk q of the patch row
                const int kc = q >> 3, j = q & 7;
                const uint32_t addr = ptx::smem_u32(a + (size_t)kc * CONV_A_BYTES) + (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) * 16);
                ptx::st_shared_v4(addr, ((const uint4*)vals)[q]);
            }
            ptx::fence_proxy_async(); // generic-proxy writes -> visible to the tensor core (async proxy)
            ptx::mbar_arrive(ptx::smem_u32(full_bar + stage));
            if (++stage == STEM_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 8) {
        // ===================== weights (once) + MMA issuer =====================
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(ptx::smem_u32(w_bar), (uint32_t)(KCH * p.BN * 128));
            for (int kc = 0; kc < KCH; ++kc) ptx::tma_load_2d(ptx::smem_u32(sB + (size_t)kc * p.BN * 128), &tmap_b, ptx::smem_u32(w_bar), kc * 64, 0);
            ptx::mbar_wait(ptx::smem_u32(w_bar), 0);
            const uint32_t idesc = ptx::make_idesc_f16(CONV_BLOCK_M, p.BN);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(ptx::smem_u32(tempty_bar + acc), acc_phase ^ 1);
                ptx::mbar_wait(ptx::smem_u32(full_bar + stage), phase);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
#pragma unroll
                for (int kc = 0; kc < KCH; ++kc) {
                    const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + (size_t)stage * A_BYTES + (size_t)kc * CONV_A_BYTES));
                    const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB + (size_t)kc * p.BN * 128));
                    constexpr int KK = (KCH == 1) ? (KTOT + 15) / 16 : 4; // 3x3: only the first two 16-wide k slices are non-zero
#pragma unroll
                    for (int k = 0; k < KK; ++k) ptx::umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kc | k) != 0 ? 1u : 0u);
                }
                ptx::umma_commit(ptx::smem_u32(empty_bar + stage));
                ptx::umma_commit(ptx::smem_u32(tfull_bar + acc));
                if (++stage == STEM_STAGES) { stage = 0; phase ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (warps 4..7) =====================
        const int ew = warp - 4, row = ew * 32 + lane;
        const bool leader = (warp == 4 && lane == 0);
        int acc = 0; uint32_t acc_phase = 0, stage_ctr = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int p0 = tile * CONV_BLOCK_M;
            ptx::mbar_wait(ptx::smem_u32(tfull_bar + acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int sub = 0; sub < p.BN / 64; ++sub, ++stage_ctr) {
                uint8_t* sbuf = sOut + (stage_ctr & 1) * CONV_A_BYTES;
                if (leader) ptx::bulk_wait_group_read<1>();
                ptx::named_bar_sync(1, 128);
                const uint32_t srow = ptx::smem_u32(sbuf) + (uint32_t)row * 128u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c0 = sub * 64 + q * 16;
                    uint32_t v[16];
                    ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
                    ptx::tmem_ld_wait();
                    uint32_t pk[8];
                    float bv[16], av[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        *(float4*)&bv[4 * j] = __ldg((const float4*)(p.bias + c0) + j);
                        *(float4*)&av[4 * j] = __ldg((const float4*)(p.alpha + c0) + j);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float a0 = __uint_as_float(v[2 * j]) + bv[2 * j];
                        float a1 = __uint_as_float(v[2 * j + 1]) + bv[2 * j + 1];
                        a0 = a0 > 0.f ? a0 : a0 * av[2 * j];
                        a1 = a1 > 0.f ? a1 : a1 * av[2 * j + 1];
                        const __half2 h2 = __floats2half2_rn(a0, a1);
                        pk[j] = *(const uint32_t*)&h2;
                    }
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2) ^ (row & 7)) * 16), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2 + 1) ^ (row & 7)) * 16), make_uint4(pk[4], pk[5], pk[6], pk[7]));
                }
                ptx::fence_proxy_async();
                ptx::named_bar_sync(1, 128);
                if (leader) {
                    ptx::tma_store_2d(&tmap_o, ptx::smem_u32(sbuf), p.out_ch_off + sub * 64, p0);
                    ptx::bulk_commit_group();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(tempty_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (leader) ptx::bulk_wait_group_read<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// 3x3 stem, second version (VGG / MobileNet first layer, stride 1 or 2): same idea as conv_stem_kernel -- the patches
// exist only in shared memory -- but built for instruction count and occupancy, which is what bounded the first one
// (ncu: 168 registers -> 1 CTA per SM, 4 gather warps per SM, ~500 instructions per pixel, issue slot busy 20 %):
//   * the u8 -> fp16 normalisation (u8 * factor in double, - mean: src/data.cpp:48, backbones.py:455) is a 3 x 256 entry
//     table in shared memory, exact by construction, instead of I2F.F64 / DMUL / F2F per value;
//   * the 9 bytes of a filter row are contiguous in the BGR frame: three aligned 32-bit loads + funnel shifts replace
//     nine byte loads; out-of-image taps select a zero;
//   * only k = 0..31 of the 128-byte patch row is written (27 taps + 5 zeros) and only two K=16 MMA slices are issued;
//   * 113 registers, ~107 KB of shared memory: two CTAs per SM, so two gather groups overlap each other's latency.
// Warp roles: 0-3 gather, 4-7 epilogue, 8 weights + MMA issue + TMEM.  FLIP = R/B swap of tensorrt.cpp's flip_rgb.
// ---------------------------------------------------------------------------------------------
constexpr int STEM3_STAGES = 4;

template <bool FLIP>
__global__ void __launch_bounds__(STEM_THREADS, 2)
conv_stem3_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_o, const StemParams p)
{
    extern __shared__ uint8_t smem_raw[];
    ptx::pdl_launch_dependents();   // the next kernel may start its prologue on every SM this grid has left
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                          // [STAGES][128 rows x 128 B]
    uint8_t* sB = sA + (size_t)STEM3_STAGES * CONV_A_BYTES;      // [BN x 128 B] weights, resident
    uint8_t* sOut = sB + (size_t)p.BN * 128;                     // 2 x 16 KiB staging
    __half* lut = (__half*)(sOut + 2 * CONV_A_BYTES);            // [3][256] + one zero entry (padded to 1552 B)
    uint64_t* full_bar = (uint64_t*)((uint8_t*)lut + 1552);      // [STAGES] gather -> MMA (128 arrivals)
    uint64_t* empty_bar = full_bar + STEM3_STAGES;               // [STAGES] MMA -> gather
    uint64_t* tfull_bar = empty_bar + STEM3_STAGES;              // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                        // [2]
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_px = p.Nb * p.OH * p.OW;
    const int total_tiles = (total_px + CONV_BLOCK_M - 1) / CONV_BLOCK_M;
    const int tmem_cols = p.BN <= 64 ? 128 : 256;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_o);
        for (int i = 0; i < STEM3_STAGES; ++i) {
            ptx::mbar_init(ptx::smem_u32(full_bar + i), 128);
            ptx::mbar_init(ptx::smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(tfull_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(tempty_bar + i), 4);
        }
        ptx::mbar_init(ptx::smem_u32(w_bar), 1);
        ptx::fence_barrier_init();
    }
    // normalisation table: lut[c][u] = half((float)((double)u * factor) - mean[c]), c = model channel
    for (int i = threadIdx.x; i < 3 * 256 + 8; i += STEM_THREADS) {
        float v = 0.f;
        if (i < 768) {
            const int c = i >> 8, u = i & 255;
            v = (float)((double)u * p.factor) - (c == 0 ? p.m0 : c == 1 ? p.m1 : p.m2);
        }
        lut[i] = __float2half_rn(v);
    }
    // the patch rows' k slots >= 32 are never written or read; slots 27..31 are written as zeros by every gather
    if (warp == 8) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), (uint32_t)tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

    if (warp < 4) {
        // ===================== gather: one thread per pixel of the tile =====================
        const int row = threadIdx.x; // 0..127
        const uint32_t lut_s = ptx::smem_u32(lut);
        const uint32_t swz = (uint32_t)(row & 7);
        const long long total_bytes = (long long)p.Nb * p.H * p.W * 3;
        const uint8_t* frames = p.frames; // 4-byte aligned (cudaMalloc), readable up to the next multiple of 4
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int px = tile * CONV_BLOCK_M + row;
            const bool valid = px < total_px;
            int n = 0, oh = 0, ow = 0;
            if (valid) { n = px / (p.OH * p.OW); const int rem = px - n * p.OH * p.OW; oh = rem / p.OW; ow = rem - oh * p.OW; }
            const int h0 = oh * p.stride - p.pad_h, w0 = ow * p.stride - p.pad_w;
            bool cok[3];
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) cok[s2] = (w0 + s2) >= 0 && (w0 + s2) < p.W;
            uint32_t hv[27]; // fp16 bit patterns in the low halves
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int hh = h0 + r;
                const bool rok = valid && hh >= 0 && hh < p.H;
                // the 9 bytes of (row hh, columns w0 .. w0+2): three aligned words, re-aligned by a funnel shift
                const long long off = (((long long)n * p.H + (rok ? hh : 0)) * p.W + w0) * 3;
                const long long a0 = off & ~3ll;
                const uint32_t sh = (uint32_t)(off & 3) * 8u;
                uint32_t wd[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const long long a = a0 + 4 * i;
                    wd[i] = (rok && a >= 0 && a < total_bytes) ? __ldg((const uint32_t*)(frames + a)) : 0u;
                }
                uint32_t seg[3];
                seg[0] = __funnelshift_r(wd[0], wd[1], sh);
                seg[1] = __funnelshift_r(wd[1], wd[2], sh);
                seg[2] = wd[2] >> sh;
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2) {
                    const bool ok = rok && cok[s2];
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const int j = s2 * 3 + b;                 // byte of the segment (memory order B, G, R)
                        const int c = FLIP ? 2 - b : b;           // model channel fed by this byte
                        const uint32_t byte = __byte_perm(seg[j >> 2], 0u, 0x4440u | (uint32_t)(j & 3));
                        const uint32_t idx = ok ? (byte * 2u + (uint32_t)(c * 512)) : 1536u; // entry 768 is zero
                        uint32_t v;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(lut_s + idx));
                        hv[(r * 3 + s2) * 3 + c] = v;
                    }
                }
            }
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 13; ++i) pk[i] = hv[2 * i] | (hv[2 * i + 1] << 16);
            pk[13] = hv[26];
            pk[14] = 0u; pk[15] = 0u;
            ptx::mbar_wait(ptx::smem_u32(empty_bar + stage), phase ^ 1);
            const uint32_t arow = ptx::smem_u32(sA + (size_t)stage * CONV_A_BYTES) + (uint32_t)row * 128u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                ptx::st_shared_v4(arow + ((uint32_t)q ^ swz) * 16u, make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]));
            ptx::fence_proxy_async(); // generic-proxy writes -> visible to the tensor core (async proxy)
            ptx::mbar_arrive(ptx::smem_u32(full_bar + stage));
            if (++stage == STEM3_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 8) {
        // ===================== weights (once) + MMA issuer =====================
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(ptx::smem_u32(w_bar), (uint32_t)(p.BN * 128));
            ptx::tma_load_2d(ptx::smem_u32(sB), &tmap_b, ptx::smem_u32(w_bar), 0, 0);
            ptx::mbar_wait(ptx::smem_u32(w_bar), 0);
            const uint32_t idesc = ptx::make_idesc_f16(CONV_BLOCK_M, p.BN);
            const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB));
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(ptx::smem_u32(tempty_bar + acc), acc_phase ^ 1);
                ptx::mbar_wait(ptx::smem_u32(full_bar + stage), phase);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + (size_t)stage * CONV_A_BYTES));
                ptx::umma_f16(d_tmem, da, db, idesc, 0u);                 // k = 0..15
                ptx::umma_f16(d_tmem, da + 2ull, db + 2ull, idesc, 1u);   // k = 16..31 (taps 16..26 + zeros)
                ptx::umma_commit(ptx::smem_u32(empty_bar + stage));
                ptx::umma_commit(ptx::smem_u32(tfull_bar + acc));
                if (++stage == STEM3_STAGES) { stage = 0; phase ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (warps 4..7): bias + PReLU -> fp16 -> swizzled staging -> TMA store =====================
        const int ew = warp - 4, row = ew * 32 + lane;
        const bool leader = (warp == 4 && lane == 0);
        int acc = 0; uint32_t acc_phase = 0, stage_ctr = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int p0 = tile * CONV_BLOCK_M;
            ptx::mbar_wait(ptx::smem_u32(tfull_bar + acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int sub = 0; sub < p.BN / 64; ++sub, ++stage_ctr) {
                uint8_t* sbuf = sOut + (stage_ctr & 1) * CONV_A_BYTES;
                if (leader) ptx::bulk_wait_group_read<1>();
                ptx::named_bar_sync(1, 128);
                const uint32_t srow = ptx::smem_u32(sbuf) + (uint32_t)row * 128u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c0 = sub * 64 + q * 16;
                    uint32_t v[16];
                    ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
                    ptx::tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) { // four channels at a time keeps the register count under the 2-CTA limit
                        const float4 bv = __ldg((const float4*)(p.bias + c0) + j4);
                        const float4 av = __ldg((const float4*)(p.alpha + c0) + j4);
                        float a0 = __uint_as_float(v[4 * j4]) + bv.x, a1 = __uint_as_float(v[4 * j4 + 1]) + bv.y;
                        float a2 = __uint_as_float(v[4 * j4 + 2]) + bv.z, a3 = __uint_as_float(v[4 * j4 + 3]) + bv.w;
                        a0 = a0 > 0.f ? a0 : a0 * av.x; a1 = a1 > 0.f ? a1 : a1 * av.y;
                        a2 = a2 > 0.f ? a2 : a2 * av.z; a3 = a3 > 0.f ? a3 : a3 * av.w;
                        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
                        pk[2 * j4] = *(const uint32_t*)&h01;
                        pk[2 * j4 + 1] = *(const uint32_t*)&h23;
                    }
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2) ^ (row & 7)) * 16), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2 + 1) ^ (row & 7)) * 16), make_uint4(pk[4], pk[5], pk[6], pk[7]));
                }
                ptx::fence_proxy_async();
                ptx::named_bar_sync(1, 128);
                if (leader) {
                    ptx::tma_store_2d(&tmap_o, ptx::smem_u32(sbuf), p.out_ch_off + sub * 64, p0);
                    ptx::bulk_commit_group();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(tempty_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (leader) ptx::bulk_wait_group_read<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// 7x7 / stride-2 stem (ResNet-50 first layer), built like conv_stem3_kernel: table-driven u8 -> fp16 normalisation, the 21 bytes
// of a filter row (7 pixels x BGR, contiguous in the frame) as six aligned 32-bit loads + funnel shifts, patches only ever in
// shared memory.  conv_stem_kernel<7> (the first version: ~2600 instructions per pixel, 168 registers, one CTA per SM) ran at
// 36 TFLOP/s = 0.68 ms per 32-frame batch at cfg4, 10 % of the step, for 0.2 % of the FLOPs.
// K layout (chosen here, the weights are packed to match at load time): filter row r owns 24 slots, slot r * 24 + s * 3 + c
// (3 pad slots per row), i.e. 48 bytes = three 16-byte pieces that never straddle a 64-slot chunk; 7 rows = 168 slots in three
// chunks; the pieces never written (slots 168..191) are zeroed once and stay zero.  One A stage (48 KiB) per CTA and two CTAs per
// SM: the second CTA's gather overlaps this CTA's MMA / epilogue.  11 MMA slices per tile (4 + 4 + 3).
// Warp roles as in conv_stem3_kernel: 0-3 gather (one thread per pixel), 4-7 epilogue, 8 weights + MMA issue + TMEM.
// ---------------------------------------------------------------------------------------------
constexpr int STEM7_ROW_SLOTS = 24, STEM7_CHUNKS = 3;
constexpr int STEM7_A_BYTES = STEM7_CHUNKS * CONV_A_BYTES;   // 48 KiB

template <bool FLIP>
__global__ void __launch_bounds__(STEM_THREADS, 2)
conv_stem7_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_o, const StemParams p)
{
    extern __shared__ uint8_t smem_raw[];
    ptx::pdl_launch_dependents();   // the next kernel may start its prologue on every SM this grid has left
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                          // [3 chunks][128 rows x 128 B]
    uint8_t* sB = sA + STEM7_A_BYTES;                            // [3 chunks][BN x 128 B] weights, resident
    uint8_t* sOut = sB + (size_t)STEM7_CHUNKS * p.BN * 128;      // 2 x 16 KiB staging
    __half* lut = (__half*)(sOut + 2 * CONV_A_BYTES);            // [3][256] + one zero entry (padded to 1552 B)
    uint64_t* full_bar = (uint64_t*)((uint8_t*)lut + 1552);      // gather -> MMA (128 arrivals)
    uint64_t* empty_bar = full_bar + 1;                          // MMA -> gather
    uint64_t* tfull_bar = empty_bar + 1;                         // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                        // [2]
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_px = p.Nb * p.OH * p.OW;
    const int total_tiles = (total_px + CONV_BLOCK_M - 1) / CONV_BLOCK_M;
    const int tmem_cols = p.BN <= 64 ? 128 : 256;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_b);
        ptx::prefetch_tmap(&tmap_o);
        ptx::mbar_init(ptx::smem_u32(full_bar), 128);
        ptx::mbar_init(ptx::smem_u32(empty_bar), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(tfull_bar + i), 1);
            ptx::mbar_init(ptx::smem_u32(tempty_bar + i), 4);
        }
        ptx::mbar_init(ptx::smem_u32(w_bar), 1);
        ptx::fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 3 * 256 + 8; i += STEM_THREADS) {
        float v = 0.f;
        if (i < 768) {
            const int c = i >> 8, u = i & 255;
            v = (float)((double)u * p.factor) - (c == 0 ? p.m0 : c == 1 ? p.m1 : p.m2);
        }
        lut[i] = __float2half_rn(v);
    }
    // slots 168..191 (pieces 5..7 of chunk 2) are never written by the gather: zero the tile once
    for (int i = threadIdx.x; i < STEM7_A_BYTES / 16; i += STEM_THREADS) ((uint4*)sA)[i] = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();
    if (warp == 8) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), (uint32_t)tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

    if (warp < 4) {
        // ===================== gather: one thread per pixel of the tile =====================
        const int row = threadIdx.x; // 0..127
        const uint32_t lut_s = ptx::smem_u32(lut);
        const uint32_t swz = (uint32_t)(row & 7);
        const long long total_bytes = (long long)p.Nb * p.H * p.W * 3;
        const uint8_t* frames = p.frames;
        const uint32_t arow = ptx::smem_u32(sA) + (uint32_t)row * 128u;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int px = tile * CONV_BLOCK_M + row;
            const bool valid = px < total_px;
            int n = 0, oh = 0, ow = 0;
            if (valid) { n = px / (p.OH * p.OW); const int rem = px - n * p.OH * p.OW; oh = rem / p.OW; ow = rem - oh * p.OW; }
            const int h0 = oh * p.stride - p.pad_h, w0 = ow * p.stride - p.pad_w;
            unsigned cmask = 0u;   // bit s: column w0 + s lies inside the frame
#pragma unroll
            for (int s2 = 0; s2 < 7; ++s2) cmask |= ((w0 + s2) >= 0 && (w0 + s2) < p.W) ? (1u << s2) : 0u;
            ptx::mbar_wait(ptx::smem_u32(empty_bar), phase ^ 1);   // the previous tile's MMAs have read the stage
#pragma unroll 1
            for (int r = 0; r < 7; ++r) {
                const int hh = h0 + r;
                const bool rok = valid && hh >= 0 && hh < p.H;
                const long long off = (((long long)n * p.H + (rok ? hh : 0)) * p.W + w0) * 3;
                const long long a0 = off & ~3ll;
                const uint32_t sh = (uint32_t)(off & 3) * 8u;
                uint32_t wd[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const long long a = a0 + 4 * i;
                    wd[i] = (rok && a >= 0 && a < total_bytes) ? __ldg((const uint32_t*)(frames + a)) : 0u;
                }
                uint32_t seg[6];
#pragma unroll
                for (int i = 0; i < 5; ++i) seg[i] = __funnelshift_r(wd[i], wd[i + 1], sh);
                seg[5] = wd[5] >> sh;
                uint32_t hv[STEM7_ROW_SLOTS];
#pragma unroll
                for (int s2 = 0; s2 < 7; ++s2) {
                    const bool ok = rok && ((cmask >> s2) & 1u);
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const int j = s2 * 3 + b;                 // byte of the row segment (memory order B, G, R)
                        const int c = FLIP ? 2 - b : b;           // model channel fed by this byte
                        const uint32_t byte = __byte_perm(seg[j >> 2], 0u, 0x4440u | (uint32_t)(j & 3));
                        const uint32_t idx = ok ? (byte * 2u + (uint32_t)(c * 512)) : 1536u; // entry 768 is zero
                        uint32_t v;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(lut_s + idx));
                        hv[s2 * 3 + c] = v;
                    }
                }
                hv[21] = 0u; hv[22] = 0u; hv[23] = 0u;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int piece = r * 3 + i;                  // 16-byte piece of the 384-byte patch row
                    const uint32_t a = arow + (uint32_t)(piece >> 3) * (uint32_t)CONV_A_BYTES + (((uint32_t)(piece & 7)) ^ swz) * 16u;
                    ptx::st_shared_v4(a, make_uint4(hv[8 * i] | (hv[8 * i + 1] << 16), hv[8 * i + 2] | (hv[8 * i + 3] << 16),
                                                    hv[8 * i + 4] | (hv[8 * i + 5] << 16), hv[8 * i + 6] | (hv[8 * i + 7] << 16)));
                }
            }
            ptx::fence_proxy_async(); // generic-proxy writes -> visible to the tensor core (async proxy)
            ptx::mbar_arrive(ptx::smem_u32(full_bar));
            phase ^= 1;
        }
    } else if (warp == 8) {
        // ===================== weights (once) + MMA issuer =====================
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(ptx::smem_u32(w_bar), (uint32_t)(STEM7_CHUNKS * p.BN * 128));
            for (int c = 0; c < STEM7_CHUNKS; ++c)
                ptx::tma_load_2d(ptx::smem_u32(sB + (size_t)c * p.BN * 128), &tmap_b, ptx::smem_u32(w_bar), c * CONV_BLOCK_K, 0);
            ptx::mbar_wait(ptx::smem_u32(w_bar), 0);
            const uint32_t idesc = ptx::make_idesc_f16(CONV_BLOCK_M, p.BN);
            uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(ptx::smem_u32(tempty_bar + acc), acc_phase ^ 1);
                ptx::mbar_wait(ptx::smem_u32(full_bar), phase);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
#pragma unroll
                for (int c = 0; c < STEM7_CHUNKS; ++c) {
                    const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + (size_t)c * CONV_A_BYTES));
                    const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB + (size_t)c * p.BN * 128));
#pragma unroll
                    for (int k = 0; k < (c == 2 ? 3 : 4); ++k)     // chunk 2: slots 128..175 (168..175 are zeros)
                        ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (c | k) ? 1u : 0u);
                }
                ptx::umma_commit(ptx::smem_u32(empty_bar));
                ptx::umma_commit(ptx::smem_u32(tfull_bar + acc));
                phase ^= 1;
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (warps 4..7): bias + PReLU -> fp16 -> swizzled staging -> TMA store =====================
        const int ew = warp - 4, row = ew * 32 + lane;
        const bool leader = (warp == 4 && lane == 0);
        int acc = 0; uint32_t acc_phase = 0, stage_ctr = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int p0 = tile * CONV_BLOCK_M;
            ptx::mbar_wait(ptx::smem_u32(tfull_bar + acc), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int sub = 0; sub < p.BN / 64; ++sub, ++stage_ctr) {
                uint8_t* sbuf = sOut + (stage_ctr & 1) * CONV_A_BYTES;
                if (leader) ptx::bulk_wait_group_read<1>();
                ptx::named_bar_sync(1, 128);
                const uint32_t srow = ptx::smem_u32(sbuf) + (uint32_t)row * 128u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c0 = sub * 64 + q * 16;
                    uint32_t v[16];
                    ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
                    ptx::tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 bv = __ldg((const float4*)(p.bias + c0) + j4);
                        const float4 av = __ldg((const float4*)(p.alpha + c0) + j4);
                        float a0 = __uint_as_float(v[4 * j4]) + bv.x, a1 = __uint_as_float(v[4 * j4 + 1]) + bv.y;
                        float a2 = __uint_as_float(v[4 * j4 + 2]) + bv.z, a3 = __uint_as_float(v[4 * j4 + 3]) + bv.w;
                        a0 = a0 > 0.f ? a0 : a0 * av.x; a1 = a1 > 0.f ? a1 : a1 * av.y;
                        a2 = a2 > 0.f ? a2 : a2 * av.z; a3 = a3 > 0.f ? a3 : a3 * av.w;
                        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
                        pk[2 * j4] = *(const uint32_t*)&h01;
                        pk[2 * j4 + 1] = *(const uint32_t*)&h23;
                    }
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2) ^ (row & 7)) * 16), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    ptx::st_shared_v4(srow + (uint32_t)(((q * 2 + 1) ^ (row & 7)) * 16), make_uint4(pk[4], pk[5], pk[6], pk[7]));
                }
                ptx::fence_proxy_async();
                ptx::named_bar_sync(1, 128);
                if (leader) {
                    ptx::tma_store_2d(&tmap_o, ptx::smem_u32(sbuf), p.out_ch_off + sub * 64, p0);
                    ptx::bulk_commit_group();
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(tempty_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (leader) ptx::bulk_wait_group_read<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
    }
}

inline size_t conv_stem7_smem_bytes(int BN)
{
    return 1024 + (size_t)STEM7_A_BYTES + (size_t)STEM7_CHUNKS * BN * 128 + 2 * CONV_A_BYTES + 1552 + 7 * 8 + 16;
}

inline size_t conv_stem3_smem_bytes(int BN)
{
    return 1024 + (size_t)STEM3_STAGES * CONV_A_BYTES + (size_t)BN * 128 + 2 * CONV_A_BYTES + 1552 + (2 * STEM3_STAGES + 5) * 8 + 16;
}

inline size_t conv_stem_smem_bytes(int R, int BN)
{
    const int kch = (R * R * 3 + 63) / 64;
    return 1024 + (size_t)STEM_STAGES * kch * CONV_A_BYTES + (size_t)kch * BN * 128 + 2 * CONV_A_BYTES + (2 * STEM_STAGES + 5) * 8 + 16;
}

constexpr size_t CONV_SMEM_FIXED = 1024 /*base alignment*/ + (2 * CONV_MAX_STAGES + 8) * 8 + 16 /*tmem slot*/ + 1024 /*staging alignment*/;
inline size_t conv_smem_bytes(int BN, int stages, bool tma_store, bool res_tma = false, int res_stages = 2)
{
    return CONV_SMEM_FIXED + (size_t)stages * (CONV_A_BYTES + BN * CONV_BLOCK_K * 2) + (tma_store ? 2 * CONV_A_BYTES : 0) + (res_tma ? res_stages * CONV_A_BYTES : 0);
}
inline size_t conv_swap_smem_bytes(int npx, int stages) { return CONV_SMEM_FIXED + (size_t)stages * (CONV_A_BYTES + npx * 128) + 2 * CONV_A_BYTES; }
inline int conv_swap_pick_stages(int npx)
{
    const size_t avail = CONV_SMEM_LIMIT - CONV_SMEM_FIXED - 2 * CONV_A_BYTES;
    const int st = (int)(avail / (size_t)(CONV_A_BYTES + npx * 128));
    return st > CONV_MAX_STAGES ? CONV_MAX_STAGES : st;
}
// unit size that fills the last wave of the persistent grid best: minimise rounds x (npx + fixed per-unit cost)
inline int conv_swap_pick_npx(long total_px, int gc, int num_sms)
{
    int best = 256;
    long best_cost = -1;
    for (int npx = 256; npx >= 144; npx -= 16) {
        const long units = (total_px + npx - 1) / npx * gc;
        const long rounds = (units + num_sms - 1) / num_sms;
        const long cost = rounds * (npx + 6) + (npx < 192 ? rounds * (192 - npx) / 4 : 0); // small units: smem operand feed
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = npx; }
    }
    return best;
}
inline int conv_pick_stages(int BN, bool tma_store, bool res_tma = false, int res_stages = 2)
{
    const size_t avail = CONV_SMEM_LIMIT - CONV_SMEM_FIXED - (tma_store ? 2 * CONV_A_BYTES : 0) - (res_tma ? res_stages * CONV_A_BYTES : 0);
    const int st = (int)(avail / (size_t)(CONV_A_BYTES + BN * CONV_BLOCK_K * 2));
    return st > CONV_MAX_STAGES ? CONV_MAX_STAGES : st;
}

} // namespace hpb
