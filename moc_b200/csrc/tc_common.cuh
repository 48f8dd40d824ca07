// tcgen05 / TMEM helpers and tile constants shared by the dense 3xTF32 kernels (linear_tc.cu: Y = act(X W^T + b);
// wgrad_tc.cu: dW = G^T X).  sm_100a only.
#pragma once
#include "common.cuh"

namespace moc {

constexpr int LT_M = 128;                 // rows per work item (UMMA M)
constexpr int LT_N = 128;                 // output columns per work item (UMMA N)
constexpr int LT_KB = 32;                 // K elements per stage (one 128-byte swizzle row)
constexpr int LT_STAGES = 3;
constexpr int LT_PF = 4;                  // K-block register sets per producer thread
constexpr int LT_A_BYTES = LT_M * 128;    // 16 KB per component
constexpr int LT_B_BYTES = LT_N * 128;    // 16 KB per component
constexpr int LT_STAGE_BYTES = 2 * LT_A_BYTES + 2 * LT_B_BYTES;  // 64 KB
constexpr int LT_EPI_WARPS = 4, LT_PROD_WARPS = 8;
constexpr int LT_WARP_MMA = LT_EPI_WARPS + LT_PROD_WARPS;  // 12
constexpr int LT_WARP_B = LT_WARP_MMA + 1;                 // 13
constexpr int LT_THREADS = (LT_WARP_B + 1) * 32;           // 448
constexpr int LT_TMEM_COLS = 256;                          // two 128-column accumulators
constexpr size_t LT_SMEM = (size_t)LT_STAGES * LT_STAGE_BYTES + 1024;

constexpr uint32_t LT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(LT_N >> 3) << 17) | ((uint32_t)(LT_M >> 4) << 24);

__device__ __forceinline__ uint64_t lt_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void lt_umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(LT_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void lt_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void lt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void lt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void lt_tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
        "[%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo with hi = x rounded to TF32 (nearest, ties away: the tensor core itself just drops the low 13 bits,
// which would bias every product the same way) and lo = x - hi exact in fp32.
__device__ __forceinline__ float lt_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void lt_split(const float4& v, float4& hi, float4& lo) {
    hi = make_float4(lt_hi(v.x), lt_hi(v.y), lt_hi(v.z), lt_hi(v.w));
    lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

__device__ __forceinline__ float lt_act(float z, int act) {
    switch (act) {
        case MOC_ACT_RELU: return fmaxf(z, 0.f);
        case MOC_ACT_TANH: return tanhf(z);
        case MOC_ACT_SIGMOID: return sigmoidf_exact(z);
        default: return z;
    }
}

}  // namespace moc
