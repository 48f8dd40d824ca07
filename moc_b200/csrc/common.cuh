// Shared device/host helpers for the moc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "moc_b200.h"

namespace moc {

constexpr int D = MOC_FEAT_DIM;      // 512 floats = 2048 B per patch
constexpr int ROW_BYTES = D * 4;
constexpr unsigned FULL = 0xffffffffu;

// ---- error plumbing (host) -------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MOC_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            moc::set_error(__VA_ARGS__);    \
            return MOC_E_ARG;               \
        }                                   \
    } while (0)
#define MOC_CHECK_SHAPE(cond, ...)          \
    do {                                    \
        if (!(cond)) {                      \
            moc::set_error(__VA_ARGS__);    \
            return MOC_E_SHAPE;             \
        }                                   \
    } while (0)
#define MOC_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return moc::cuda_fail(e__, #call); \
    } while (0)
#define MOC_LAUNCH_CHECK(name)                                \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return moc::cuda_fail(e__, name); \
    } while (0)

int sm_count();

// ---- mbarrier / bulk-copy (TMA 1-D) primitives -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint
// expires) instead of spinning through the loop and taking issue slots from the warps it is waiting for.
constexpr uint32_t MBAR_SUSPEND_NS = 1000000u;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(MBAR_SUSPEND_NS)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared bulk copy (SASS: UBLKCP), completion counted in bytes on `bar`.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// explicit shared-space accesses by 32-bit shared address: pointers derived through integer alignment arithmetic
// lose their address space and would otherwise compile to generic LD.E / ST.E
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)
                 : "memory");
    return v;
}
// bulk copy with 32-bit shared destination / barrier addresses
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(MBAR_SUSPEND_NS)
        : "memory");
}

// ---- key-plane layout (include/moc_b200.h, top) ------------------------------------
// full: [0,C) L | [C,2C) softmax | 2C diff | 2C+1 bg sum | 2C+2 bg max;  compact (C >= MOC_KEYS_COMPACT_MIN_CLASSES):
// [0,C) L | C lse = log sum exp(L) | C+1 diff | C+2 bg sum | C+3 bg max.
struct KeyLayout {
    int n_planes, diff, bg_sum, bg_max;
    int softmax0;       // full layout only (-1 otherwise)
    int lse;            // compact layout only (-1 otherwise)
    bool compact;
};
__host__ __device__ inline KeyLayout key_layout(int C) {
    KeyLayout k;
    k.compact = C >= MOC_KEYS_COMPACT_MIN_CLASSES;
    if (k.compact) {
        k.lse = C; k.diff = C + 1; k.bg_sum = C + 2; k.bg_max = C + 3;
        k.softmax0 = -1; k.n_planes = C + 4;
    } else {
        k.softmax0 = C; k.diff = 2 * C; k.bg_sum = 2 * C + 1; k.bg_max = 2 * C + 2;
        k.lse = -1; k.n_planes = 2 * C + 3;
    }
    return k;
}
// What a consumer needs of one row to form its softmax values: full layout - nothing (they are loaded); compact - lse.
struct RowSoftmax {
    float lse;
    __device__ __forceinline__ float of(float l) const { return expf(l - lse); }
};
__device__ __forceinline__ RowSoftmax row_softmax_of(const float* __restrict__ kp, int64_t key_stride, const KeyLayout& kl) {
    RowSoftmax r;
    r.lse = kp[(int64_t)kl.lse * key_stride];
    return r;
}
// the lse plane's value from the row maximum and sum exp(L - max) the scoring epilogues have at hand
__device__ __forceinline__ float lse_of(float row_max, float exp_sum) { return row_max + logf(exp_sum); }

// Gated combination of one selected row's four score planes for every class (main_moc.py:391-405): the association of
// the reference, ((g0*L + g1*P) + g2*delta) + g3*bg, with no fma contraction.  KU classes per batch of key loads: every
// load is a 32-byte sector of its own in a [planes, rows] array far larger than the L2, so with many classes the
// dependent batches of DRAM latency were what a row tile waited for (C = 30 on the full layout, 62 loads per row:
// four classes per batch 3.6 ms per 400 EBRAINS slides, eight 3.1 ms, sixteen 2.7 ms).
template <int KU, bool COMPACT>
__device__ __forceinline__ void combine_classes(const float* __restrict__ kp, int64_t key_stride, int C, RowSoftmax rs,
                                                const float (&g)[4], float a0, float a1, float a2, float a3, float dlt,
                                                float bgm, float* __restrict__ out) {
    for (int c0 = 0; c0 < C; c0 += KU) {
        float lt[KU], ls[COMPACT ? 1 : KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            const int cc = c0 + u < C ? c0 + u : C - 1;
            lt[u] = kp[(int64_t)cc * key_stride];
            if (!COMPACT) ls[u] = kp[(int64_t)(C + cc) * key_stride];
        }
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            if (c0 + u < C) {
                float f = a0 * __fmul_rn(g[0], lt[u]);
                f = __fadd_rn(f, a1 * __fmul_rn(g[1], COMPACT ? rs.of(lt[u]) : ls[u]));
                f = __fadd_rn(f, a2 * __fmul_rn(g[2], dlt));
                f = __fadd_rn(f, a3 * __fmul_rn(g[3], bgm));
                out[c0 + u] = f;
            }
        }
    }
}
#ifndef MOC_COMBINE_KU_COMPACT
#define MOC_COMBINE_KU_COMPACT 16
#endif
// kp = keys + row, out = final_scores + slot * C
__device__ __forceinline__ void combine_row(const float* __restrict__ kp, int64_t key_stride, int C, const float (&g)[4],
                                            float a0, float a1, float a2, float a3, float* __restrict__ out) {
    const KeyLayout kl = key_layout(C);
    const float dlt = kp[(int64_t)kl.diff * key_stride];
    const float bgm = kp[(int64_t)kl.bg_max * key_stride];
    RowSoftmax rs = {0.f};
    if (kl.compact) {
        rs = row_softmax_of(kp, key_stride, kl);
        combine_classes<MOC_COMBINE_KU_COMPACT, true>(kp, key_stride, C, rs, g, a0, a1, a2, a3, dlt, bgm, out);
    } else if (C <= 4) {
        combine_classes<4, false>(kp, key_stride, C, rs, g, a0, a1, a2, a3, dlt, bgm, out);
    } else {
        combine_classes<8, false>(kp, key_stride, C, rs, g, a0, a1, a2, a3, dlt, bgm, out);
    }
}

// ---- order-preserving float <-> uint32 map (for radix selection) -------------
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// ReLU that keeps NaN, like torch.relu (fmaxf(NaN, 0) is 0): non-finite features must surface as non-finite gates
__device__ __forceinline__ float relu_nan(float v) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float sigmoidf_exact(float z) { return 1.0f / (1.0f + expf(-z)); }

}  // namespace moc
