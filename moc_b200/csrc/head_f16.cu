// Gate MLP of the head on the 5th-generation tensor cores with FP16x3 operands (fp32-accurate), W1 resident in
// shared memory.  Same contract as head_tc.cu (senet's Linear(512,64) for every selected patch, main_moc.py:303, fused
// with ReLU, the 64->4 layer, the sigmoid and the classifier-bank combination, main_moc.py:390-403).  Two kernels:
// head_rows_f16t_kernel (further down; the default at every class count) keeps the split patch tile in TENSOR
// memory, head_rows_f16_kernel (round 1's default, MOC_HEAD_A=smem) stages it in shared memory.  head_tc.cu (3xTF32)
// serves features outside the FP16 split's range and stays selectable with MOC_HEAD_IMPL=tf32.
//
// Why: an M128 N64 K8 tf32 MMA reads 6 KB of operands for 32 cycles of tensor work, more than the 128 B/clk the
// SM's shared memory delivers, and the producers' stores and the W1 bulk copies share that port (measured on the
// TF32 kernel: no loads -> same time, no MMAs -> -26 %, neither -> -45 %).  FP16 operands halve every one of those
// byte streams per unit of K, halve the number of producer -> MMA hand-offs per row, and the whole split W1
// (64 x 512 x 2 halves = 128 KB) then fits in shared memory once per CTA instead of being streamed from L2 for every
// row tile.  Measured: 0.274 -> 0.225 ms per 200 NSCLC slides for the shared-memory-staged kernel, 0.164 ms with the A
// operand in tensor memory.
//
// Precision: x * 2^4 = a0 + a1 and w * 2^SW = b0 + b1 with a0, b0 the nearest FP16 and a1, b1 the FP16 of the
// remainder (exact to 2^-22 relative, or 2^-29 absolute on x where a1 becomes subnormal); D = a0 b0 + a0 b1 + a1 b0
// with fp32 accumulation in TMEM - FP16 products are exact in fp32 - then scaled back by 2^-(4+SW).  SW is chosen on
// the device from max|W1|.  Same error class as 3xTF32 (measured ~1e-6 relative on the gate pre-activations).
// Domain: |x| < 4094; larger (or non-finite) features give non-finite gates, hence non-finite bag logits.
//
// Structure (persistent, one CTA per SM, 21 warps; work item = 128 selected-row slots):
//   warps 0-3   epilogue: tcgen05.ld the 128x64 accumulator (thread = row), descale, +b1, ReLU, 4x64 second layer,
//               sigmoid, gated combination with the row's key planes, stores
//   warps 4-19  A producers: per 64-wide K-block gather 256 B of 8 rows each (32 B per thread and row, HF_PF
//               K-blocks of loads in flight per thread, across tile boundaries), split into FP16 halves in registers,
//               one 16-byte store each into the 128B-swizzled K-major a0 / a1 tiles.  Sixteen warps, not eight: a
//               producer's ~100 dependent ALU / store / fence instructions per K-block are latency-bound, and two
//               warps per scheduler could not hide that (0.2 IPC per scheduler in the TF32 kernel)
//   warp 20     MMA issuer (one elected lane): 4 k-steps x 3 products of tcgen05.mma.kind::f16 M128 N64 K16
// Three A stages of {a0 16K, a1 16K}; W1 image 8 x {b0 8K, b1 8K}; two TMEM accumulators.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace moc {

constexpr int HF_H = MOC_HIDDEN;   // 64 (UMMA N)
constexpr int HF_G = MOC_GATES;
constexpr int HF_M = 128;          // rows per tile (UMMA M)
constexpr int HF_KB = 64;          // K elements per stage (one 128-byte swizzle row of halves)
constexpr int HF_NKB = D / HF_KB;  // 8
constexpr int HF_STAGES = 3;
constexpr int HF_PF = 2;           // K-block register sets (4 x 16 B) a producer thread keeps in flight: 64 KB per CTA
constexpr int HF_RPT = 2;          // (row, 32-byte chunk) items per producer thread and K-block
constexpr int HF_SX = 4;           // features are split as x * 2^4
constexpr int HF_A_BYTES = HF_M * 128;            // 16 KB per component
constexpr int HF_STAGE_BYTES = 2 * HF_A_BYTES;    // 32 KB
constexpr int HF_B_BYTES = HF_H * 128;            // 8 KB per component and K-block
constexpr int HF_BIMG_BYTES = HF_NKB * 2 * HF_B_BYTES;  // 128 KB
constexpr int HF_EPI_WARPS = 4, HF_PROD_WARPS = 16;  // 16 producer warps: 4 per scheduler, enough to hide their own ALU / store latency
constexpr int HF_WARP_MMA = HF_EPI_WARPS + HF_PROD_WARPS;  // 20
constexpr int HF_THREADS = (HF_WARP_MMA + 1) * 32;         // 672
constexpr int HF_TMEM_COLS = 128;
constexpr int HF_DEFAULT_A_IN_TMEM = 1;   // head_rows_f16t_kernel (A in tensor memory) is the faster one: 0.75 vs 0.93 ms per 1 000 NSCLC slides
constexpr size_t HF_SMEM = (size_t)HF_BIMG_BYTES + (size_t)HF_STAGES * HF_STAGE_BYTES + 1024;

struct HeadF16Tail {   // after the W1 image in the workspace
    float scale, descale;
    int pad0, pad1;
};

// D=f32, A=B=f16, both K-major, N=64, M=128
constexpr uint32_t HF_IDESC = (1u << 4) | ((uint32_t)(HF_H >> 3) << 17) | ((uint32_t)(HF_M >> 4) << 24);

__device__ __forceinline__ uint64_t hf_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void hf_umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(HF_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void hf_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void hf_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hf_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hf_tmem_ld32(uint32_t taddr, float (&v)[32]) {
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

// 8 floats (already scaled) -> 8 halves hi + 8 halves lo, packed as two 16-byte chunks
__device__ __forceinline__ void hf_split8(const float4& u, const float4& v, float s, uint4& hi, uint4& lo) {
    const float x[8] = {u.x * s, u.y * s, u.z * s, u.w * s, v.x * s, v.y * s, v.z * s, v.w * s};
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __half2 a = __floats2half2_rn(x[2 * p], x[2 * p + 1]);
        const float2 f = __half22float2(a);
        const __half2 b = __floats2half2_rn(x[2 * p] - f.x, x[2 * p + 1] - f.y);
        h[p] = *reinterpret_cast<const uint32_t*>(&a);
        l[p] = *reinterpret_cast<const uint32_t*>(&b);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One block.  SW = 13 - floor(log2(max|W1|)) so that max|w| * 2^SW lies in [2^13, 2^14); W1 [64][512] -> per K-block
// (64) one tile {b0: 64 rows x 128 B | b1: 64 rows x 128 B} in the swizzled K-major layout, then the tail.
__global__ void __launch_bounds__(1024)
head_f16_prep_kernel(const float* __restrict__ w1, unsigned char* __restrict__ img) {
    __shared__ float red[32];
    __shared__ float scale_s;
    float m = 0.f;
    for (int i = threadIdx.x; i < HF_H * D; i += blockDim.x) {
        const float a = fabsf(w1[i]);
        if (a <= 3.0e38f) m = fmaxf(m, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
        int sw = 0;
        if (m > 0.f) sw = 13 - ilogbf(m);
        sw = sw > 100 ? 100 : (sw < -100 ? -100 : sw);
        scale_s = ldexpf(1.0f, sw);
        HeadF16Tail* tail = reinterpret_cast<HeadF16Tail*>(img + HF_BIMG_BYTES);
        tail->scale = scale_s;
        tail->descale = ldexpf(1.0f, -(HF_SX + sw));
        tail->pad0 = tail->pad1 = 0;
    }
    __syncthreads();
    const float sc = scale_s;
    for (int i = threadIdx.x; i < HF_H * (D / 8); i += blockDim.x) {   // one 8-element chunk of W1
        const int n = i / (D / 8), c8 = i % (D / 8);
        const int kb = c8 / 8, chunk = c8 % 8;
        const float4 u = reinterpret_cast<const float4*>(w1 + (size_t)n * D)[2 * c8];
        const float4 v = reinterpret_cast<const float4*>(w1 + (size_t)n * D)[2 * c8 + 1];
        uint4 hi, lo;
        hf_split8(u, v, sc, hi, lo);
        unsigned char* tile = img + (size_t)kb * 2 * HF_B_BYTES;
        const int off = n * 128 + ((chunk ^ (n & 7)) << 4);
        *reinterpret_cast<uint4*>(tile + off) = hi;
        *reinterpret_cast<uint4*>(tile + HF_B_BYTES + off) = lo;
    }
}

__device__ __forceinline__ bool hf_tile_has_rows(const int32_t* __restrict__ sel_rows, int64_t slot0, int64_t n_slots,
                                                 int lane) {
    if (sel_rows == nullptr) return true;
    bool any = false;
#pragma unroll
    for (int i = 0; i < HF_M / 32; ++i) {
        const int64_t s = slot0 + lane + 32 * i;
        any |= (s < n_slots) && (sel_rows[s] >= 0);
    }
    return __any_sync(FULL, any);
}

__global__ void __launch_bounds__(HF_THREADS, 1)
head_rows_f16_kernel(const float* __restrict__ feat, const float* __restrict__ keys, int64_t key_stride, int C,
                     const int32_t* __restrict__ sel_rows, int64_t n_slots, const unsigned char* __restrict__ img,
                     const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     unsigned active_mask, float* __restrict__ gate, float* __restrict__ final_scores,
                     int* __restrict__ domain_flag) {
    extern __shared__ unsigned char hf_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[HF_STAGES], empty_bar[HF_STAGES], tfull_bar[2], tempty_bar[2], b_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float w2s[HF_G * HF_H], b1s[HF_H], b2s[HF_G];

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(hf_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* bsm = smem;                       // resident W1 image
    unsigned char* asm_ = smem + HF_BIMG_BYTES;      // A stages
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < HF_G * HF_H) w2s[tid] = w2[tid];
    if (tid < HF_H) b1s[tid] = b1[tid];
    if (tid < HF_G) b2s[tid] = b2[tid];
    if (tid == 0) {
        for (int s = 0; s < HF_STAGES; ++s) {
            mbar_init(&full_bar[s], HF_PROD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], HF_EPI_WARPS);
        }
        mbar_init(&b_bar, 1);
        fence_mbar_init();
    }
    if (warp == HF_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)HF_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    hf_fence_before();
    __syncthreads();
    hf_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int64_t n_tiles = (n_slots + HF_M - 1) / HF_M;
    const float descale = reinterpret_cast<const HeadF16Tail*>(img + HF_BIMG_BYTES)->descale;

    if (warp >= HF_EPI_WARPS && warp < HF_WARP_MMA) {
        // =============================== A producers ===============================================
        const int pw = warp - HF_EPI_WARPS;      // 0..15 : rows 8*pw .. 8*pw+7 of the tile
        const int rsub = lane >> 3, chunk = lane & 7;
        uint32_t roff[HF_RPT];
#pragma unroll
        for (int i = 0; i < HF_RPT; ++i) {
            const int r = pw * (4 * HF_RPT) + i * 4 + rsub;
            roff[i] = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
        }
        const uint32_t a_base = smem_u32(asm_);
        const float sx = (float)(1 << HF_SX);
        int64_t ltile = blockIdx.x;
        int lkb = 0;
        const float4* lsrc[HF_RPT];
        auto seek = [&]() {  // move ltile to the next tile with rows and fetch its row pointers
            while (ltile < n_tiles && !hf_tile_has_rows(sel_rows, ltile * HF_M, n_slots, lane)) ltile += gridDim.x;
            if (ltile >= n_tiles) return;
#pragma unroll
            for (int i = 0; i < HF_RPT; ++i) {
                const int64_t sl = ltile * HF_M + pw * (4 * HF_RPT) + i * 4 + rsub;
                int64_t row = -1;
                if (sl < n_slots) row = sel_rows ? (int64_t)sel_rows[sl] : sl;
                lsrc[i] = row >= 0 ? reinterpret_cast<const float4*>(feat + row * D) + chunk * 2 : nullptr;
            }
        };
        auto issue = [&](float4 (&b)[2 * HF_RPT]) -> bool {  // load the cursor's step into b and advance; false when done
            if (ltile >= n_tiles) return false;
#pragma unroll
            for (int i = 0; i < HF_RPT; ++i) {
                if (lsrc[i]) {
                    b[2 * i] = __ldg(lsrc[i] + lkb * (HF_KB / 4));
                    b[2 * i + 1] = __ldg(lsrc[i] + lkb * (HF_KB / 4) + 1);
                } else {
                    b[2 * i] = b[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (++lkb == HF_NKB) {
                lkb = 0;
                ltile += gridDim.x;
                seek();
            }
            return true;
        };
        float4 buf[HF_PF][2 * HF_RPT];
        bool pending[HF_PF];
        seek();
#pragma unroll
        for (int s = 0; s < HF_PF; ++s) pending[s] = issue(buf[s]);
        int stage = 0;
        uint32_t parity = 0;
        while (pending[0]) {
#pragma unroll
            for (int s = 0; s < HF_PF; ++s) {
                if (pending[s]) {
                    uint4 hi[HF_RPT], lo[HF_RPT];
#pragma unroll
                    for (int i = 0; i < HF_RPT; ++i) hf_split8(buf[s][2 * i], buf[s][2 * i + 1], sx, hi[i], lo[i]);
                    mbar_wait(&empty_bar[stage], parity ^ 1u);
                    const uint32_t a0 = a_base + stage * HF_STAGE_BYTES, a1 = a0 + HF_A_BYTES;
#pragma unroll
                    for (int i = 0; i < HF_RPT; ++i) {
                        sts128u(a0 + roff[i], hi[i]);
                        sts128u(a1 + roff[i], lo[i]);
                    }
                    fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                    if (++stage == HF_STAGES) { stage = 0; parity ^= 1u; }
                    pending[s] = issue(buf[s]);
                }
            }
        }
    } else if (warp == HF_WARP_MMA) {
        // =============================== MMA issuer ================================================
        if (lane == 0) {   // W1 image -> shared memory, once per CTA
            const uint64_t policy = l2_policy_evict_last();
            mbar_arrive_expect_tx(&b_bar, HF_BIMG_BYTES);
            for (int kb = 0; kb < HF_NKB; ++kb)
                bulk_g2s(bsm + (size_t)kb * 2 * HF_B_BYTES, img + (size_t)kb * 2 * HF_B_BYTES, 2 * HF_B_BYTES, &b_bar, policy);
            mbar_wait(&b_bar, 0);
        }
        __syncwarp();
        int stage = 0, acc = 0;
        uint32_t parity = 0, acc_parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!hf_tile_has_rows(sel_rows, tile * HF_M, n_slots, lane)) continue;
            if (lane == 0) {
                mbar_wait(&tempty_bar[acc], acc_parity ^ 1u);  // epilogue has drained this accumulator
                hf_fence_after();
            }
            __syncwarp();
            const uint32_t tmem_d = tmem_base + acc * HF_H;
            for (int kb = 0; kb < HF_NKB; ++kb) {
                if (lane == 0) {
                    mbar_wait(&full_bar[stage], parity);
                    hf_fence_after();
                    const uint32_t a0 = smem_u32(asm_ + (size_t)stage * HF_STAGE_BYTES);
                    const uint32_t a1 = a0 + HF_A_BYTES;
                    const uint32_t b0 = smem_u32(bsm + (size_t)kb * 2 * HF_B_BYTES);
                    const uint32_t bl = b0 + HF_B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < HF_KB / 16; ++ks) {
                        const uint32_t o = ks * 32;  // 16 halves = 32 bytes along K inside the swizzled row
                        const uint64_t da0 = hf_desc_sw128(a0 + o), da1 = hf_desc_sw128(a1 + o);
                        const uint64_t db0 = hf_desc_sw128(b0 + o), db1 = hf_desc_sw128(bl + o);
                        hf_umma(tmem_d, da0, db1, (kb | ks) != 0 ? 1u : 0u);
                        hf_umma(tmem_d, da1, db0, 1u);
                        hf_umma(tmem_d, da0, db0, 1u);
                    }
                    hf_commit(&empty_bar[stage]);               // stage free once these MMAs have read it
                    if (kb == HF_NKB - 1) hf_commit(&tfull_bar[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == HF_STAGES) { stage = 0; parity ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
    } else {
        // =============================== epilogue (warps 0-3) =======================================
        const float a0 = (active_mask & MOC_CLS_TOPK) ? 1.f : 0.f;
        const float a1 = (active_mask & MOC_CLS_DELTA_SOFTMAX) ? 1.f : 0.f;
        const float a2 = (active_mask & MOC_CLS_DELTA_DIFF) ? 1.f : 0.f;
        const float a3 = (active_mask & MOC_CLS_BOTTOMK) ? 1.f : 0.f;
        int acc = 0;
        uint32_t acc_parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t slot0 = tile * HF_M;
            if (!hf_tile_has_rows(sel_rows, slot0, n_slots, lane)) continue;
            const int64_t slot = slot0 + warp * 32 + lane;
            int64_t row = -1;
            if (slot < n_slots) row = sel_rows ? (int64_t)sel_rows[slot] : slot;
            mbar_wait(&tfull_bar[acc], acc_parity);
            hf_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * HF_H;
            float z[HF_G] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float d[32];
                hf_tmem_ld32(taddr + half * 32, d);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int jj = half * 32 + j;
                    const float h = relu_nan(fmaf(d[j], descale, b1s[jj]));
#pragma unroll
                    for (int m = 0; m < HF_G; ++m) z[m] = fmaf(h, w2s[m * HF_H + jj], z[m]);
                }
            }
            hf_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);  // accumulator is in registers: MMA may reuse it
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
            if (row < 0) continue;
            // a feature outside the FP16 split's range (|x| >= 4094) or a non-finite one turns the pre-activations
            // non-finite: raise the caller's flag so the pass can be redone on the range-free 3xTF32 kernel
            if (domain_flag != nullptr && !(fabsf(z[0] + z[1] + z[2] + z[3]) <= 3.0e38f)) atomicOr(domain_flag, 1);
            float g[HF_G];
#pragma unroll
            for (int m = 0; m < HF_G; ++m) g[m] = sigmoidf_exact(z[m] + b2s[m]);
            if (gate != nullptr) *reinterpret_cast<float4*>(gate + slot * HF_G) = make_float4(g[0], g[1], g[2], g[3]);
            if (final_scores == nullptr) continue;
            combine_row(keys + row, key_stride, C, g, a0, a1, a2, a3, final_scores + slot * C);
        }
    }

    hf_fence_before();
    __syncthreads();
    if (warp == HF_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)HF_TMEM_COLS)
                     : "memory");
    }
}

// =====================================================================================================
// The A operand in TENSOR MEMORY: the default gate kernel (MOC_HEAD_A=smem selects the kernel above).
//
// What bounds head_rows_f16_kernel is memory-level parallelism, not request size: tools/probe_gather_kblock.cu gathers
// the same sparse rows K-block by K-block at 4.5 TB/s with 64 KB of loads in flight per SM (what 16 producer warps x
// two 64-byte register sets give: the kernel reaches 3.9) and at 6.5 TB/s with 128 KB.  The registers for deeper
// prefetch went into the (row, 32-byte chunk) work split: 8 lanes shared one row piece so that their 16-byte stores
// filled a swizzled shared-memory row.  With A read from tensor memory there is no shared-memory layout to serve: a
// producer thread owns ONE ROW of the tile (its TMEM lane), loads 128 contiguous bytes of it per step (32 K-elements,
// two steps in flight = 256 B per thread, 128 KB per SM), splits them into 16 + 16 packed half2 registers and writes
// them with two tcgen05.st into the stage's a0 / a1 columns - no STS, no fence.proxy.async, and the MMA's A reads leave
// the shared-memory port to the W1 image.  Twelve producer warps (17 warps: registers are allocated for 20, which leaves 96 per thread - two
// sets of 32 loaded floats plus 16 packed results at a time) form 3 groups (warp / 4) of 4 lane quadrants
// (warp % 4, as tcgen05.st requires); step s of the CTA's flat (tile, 32-wide K-slice) stream belongs to group s % 3
// and to TMEM stage s % 3, so every group owns one stage and one full / empty barrier pair; 96 KB of loads in flight.
constexpr int HT_KS = 32;                    // K elements per step (128 bytes of a row)
constexpr int HT_NKS = D / HT_KS;            // 16 steps per tile
constexpr int HT_GROUPS = 3;                 // producer groups of four lane-quadrant warps
constexpr int HT_STAGES = 2 * HT_GROUPS;     // TMEM A stages: step s -> group s % 3, stage s % 6 (two stages per group)
#ifndef MOC_HT_SETS
#define MOC_HT_SETS 2
#endif
#ifndef MOC_HT_PREFETCH
#define MOC_HT_PREFETCH 0
#endif
constexpr int HT_SETS = MOC_HT_SETS;         // 128-byte register sets of loads in flight per producer thread (2: 96 KB per SM)
constexpr int HT_PROD_WARPS = 4 * HT_GROUPS; // 12
constexpr int HT_WARP_MMA = HF_EPI_WARPS + HT_PROD_WARPS;   // 16
// Five warpgroups: epilogue (warps 0-3), three producer groups (4-15), and one whose first warp issues the MMAs (its
// other three warps only exist so that setmaxnreg, a warpgroup-wide instruction, can be executed).  The CTA is
// launched with 96 registers per thread (640 threads = 61 440 registers: that is the pool setmaxnreg moves registers
// in - the SM's remaining 4 096 do not belong to it, and an .inc that the pool cannot satisfy spins forever); the
// epilogue shrinks to 80, the MMA group to 40, and the producers grow to 120: 80*128 + 40*128 + 120*384 = 61 440.
constexpr int HT_THREADS = (HT_WARP_MMA + 4) * 32;          // 640
constexpr int HT_REGS_EPI = 80, HT_REGS_MMA = 40, HT_REGS_PROD = 120;
constexpr int HT_STAGE_COLS = 32;            // a0: 16 columns of packed half2, a1: 16 columns
constexpr int HT_ACC_COL0 = 0;               // two accumulators of 64 columns
constexpr int HT_A_COL0 = 128;               // four A stages of 32 columns
constexpr int HT_TMEM_COLS = 512;
constexpr size_t HT_SMEM = (size_t)HF_BIMG_BYTES + 1024;

__device__ __forceinline__ void ht_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(HF_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void ht_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void ht_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// 16 lanes x 16 packed columns: registers {4 rep + 2 sel + j} <-> (lane t/4 + 8 sel, column 8 rep + 2 (t%4) + j)
__device__ __forceinline__ void ht_tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// 4 floats (scaled here) -> 2 packed half2 hi + 2 packed half2 lo
__device__ __forceinline__ void hf_split4(const float4& u, float s, uint2& hi, uint2& lo) {
    const float x[4] = {u.x * s, u.y * s, u.z * s, u.w * s};
    uint32_t h[2], l[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const __half2 a = __floats2half2_rn(x[2 * p], x[2 * p + 1]);
        const float2 f = __half22float2(a);
        const __half2 b = __floats2half2_rn(x[2 * p] - f.x, x[2 * p + 1] - f.y);
        h[p] = *reinterpret_cast<const uint32_t*>(&a);
        l[p] = *reinterpret_cast<const uint32_t*>(&b);
    }
    hi = make_uint2(h[0], h[1]);
    lo = make_uint2(l[0], l[1]);
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void ht_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(HT_THREADS, 1)
head_rows_f16t_kernel(const float* __restrict__ feat, const float* __restrict__ keys, int64_t key_stride, int C,
                      const int32_t* __restrict__ sel_rows, int64_t n_slots, const unsigned char* __restrict__ img,
                      const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                      unsigned active_mask, float* __restrict__ gate, float* __restrict__ final_scores,
                      int* __restrict__ domain_flag) {
    extern __shared__ unsigned char ht_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[HT_STAGES], empty_bar[HT_STAGES], tfull_bar[2], tempty_bar[2], b_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float w2s[HF_G * HF_H], b1s[HF_H], b2s[HF_G];

    unsigned char* bsm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ht_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < HF_G * HF_H) w2s[tid] = w2[tid];
    if (tid < HF_H) b1s[tid] = b1[tid];
    if (tid < HF_G) b2s[tid] = b2[tid];
    if (tid == 0) {
        for (int s = 0; s < HT_STAGES; ++s) {
            mbar_init(&full_bar[s], 4);        // the group's four lane-quadrant warps
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], HF_EPI_WARPS);
        }
        mbar_init(&b_bar, 1);
        fence_mbar_init();
    }
    if (warp == HT_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)HT_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    hf_fence_before();
    __syncthreads();
    hf_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int64_t n_tiles = (n_slots + HF_M - 1) / HF_M;
    const float descale = reinterpret_cast<const HeadF16Tail*>(img + HF_BIMG_BYTES)->descale;

    if (warp >= HF_EPI_WARPS && warp < HT_WARP_MMA) {
        // =============================== A producers ===============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HT_REGS_PROD));
        const int pw = warp - HF_EPI_WARPS;          // 0..11
        const int quad = warp & 3;                   // TMEM lane quadrant this warp may access (warp id % 4)
        const int grp = pw >> 2;                     // owns steps s with s % HT_GROUPS == grp, TMEM stages grp and grp + HT_GROUPS
        // tcgen05.st.16x256b fragment: thread t holds, for the 16-lane half hh of the quadrant, rows 16 hh + t/4 and
        // + 8, and of each row the packed columns 8 rep + 2 (t % 4) + {0, 1} = K elements 16 rep + 4 (t % 4) .. + 3 =
        // float4 number 4 rep + t % 4 of the row's 32-element slice.  So the four threads of a row read 64 contiguous
        // bytes per load instruction and a warp-wide load touches 8 rows (not 32: the one-row-per-thread version
        // spent 77 % of the L1's tag throughput fetching every sector twice).
        const int rq = lane >> 2, cq = lane & 3;
        static_assert(HF_SX == 4, "the producers scale by the literal 16");
        const uint32_t a_taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + HT_A_COL0 + grp * HT_STAGE_COLS;
        // cursor over this group's steps of the flat stream (non-empty tiles of this CTA x 16 K-slices): every
        // HT_GROUPS-th one.  32-bit arithmetic throughout (slots and rows are int32 quantities): registers are what
        // this role is short of.
        const int nt = (int)n_tiles, ns = (int)n_slots, gstride = (int)gridDim.x;
        int ltile = blockIdx.x;
        int lks = grp;
        int32_t lrow[4];                             // feature rows (hh, sel): tile rows 16 hh + rq + 8 sel of the quadrant; -1 = padding
        const float4* fbase = reinterpret_cast<const float4*>(feat) + cq;
        auto seek = [&]() {   // make ltile the next tile with rows, fetch this thread's four row numbers
            while (ltile < nt && !hf_tile_has_rows(sel_rows, (int64_t)ltile * HF_M, n_slots, lane)) ltile += gstride;
            if (ltile >= nt) return;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int sl = ltile * HF_M + quad * 32 + (i >> 1) * 16 + rq + (i & 1) * 8;
                lrow[i] = sl < ns ? (sel_rows ? sel_rows[sl] : sl) : -1;
#if MOC_HT_PREFETCH
                // whole row towards the L2 now (one 2 KB request per row, issued by one of its four threads): the
                // sixteen 128-byte slices that follow over the tile's lifetime then miss only in the L1
                if (cq == 0 && lrow[i] >= 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(feat + (int64_t)lrow[i] * D), "r"(ROW_BYTES) : "memory");
#endif
            }
        };
        auto issue = [&](float4 (&b)[8]) -> int {   // b[2 i + rep] = float4 4 rep + cq of row i's slice; 1 if a step was issued
            if (ltile >= nt) return 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (lrow[i] >= 0) {   // streamed once: keep the rows out of the L1
                    const float4* p = fbase + (int64_t)lrow[i] * (D / 4) + lks * (HT_KS / 4);
                    b[2 * i] = ldg_stream(p);
                    b[2 * i + 1] = ldg_stream(p + 4);
                } else {
                    b[2 * i] = b[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            lks += HT_GROUPS;
            if (lks >= HT_NKS) {
                lks -= HT_NKS;
                ltile += gstride;
                seek();
            }
            return 1;
        };
        float4 buf[HT_SETS][8];
        seek();
        int ahead = 0;                               // steps issued and not yet written to tensor memory
#pragma unroll
        for (int s = 0; s < HT_SETS; ++s) ahead += issue(buf[s]);
        uint32_t parity = 0;
        int half = 0;                                // which of the group's two stages the next step fills
        while (ahead > 0) {
#pragma unroll
            for (int s = 0; s < HT_SETS; ++s) {
                if (ahead > 0) {
                    const int stage = grp + HT_GROUPS * half;
                    mbar_wait(&empty_bar[stage], parity ^ 1u);   // the MMAs that read this stage have completed
                    hf_fence_after();
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {           // one 16-lane half of the quadrant at a time
                        // registers {4 rep + 2 sel + j}: repetition rep (8 columns), row rq + 8 sel, column 2 cq + j
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int sel = 0; sel < 2; ++sel)
#pragma unroll
                            for (int rep = 0; rep < 2; ++rep) {
                                uint2 h, l;
                                hf_split4(buf[s][2 * (2 * hh + sel) + rep], 16.0f, h, l);
                                hi[4 * rep + 2 * sel] = h.x; hi[4 * rep + 2 * sel + 1] = h.y;
                                lo[4 * rep + 2 * sel] = l.x; lo[4 * rep + 2 * sel + 1] = l.y;
                            }
                        const uint32_t t = a_taddr + ((uint32_t)(hh * 16) << 16) + half * HT_GROUPS * HT_STAGE_COLS;
                        ht_tmem_st_16x256b_x2(t, hi);
                        ht_tmem_st_16x256b_x2(t + 16, lo);
                    }
                    ahead += issue(buf[s]) - 1;            // the registers are free again: next loads leave now
                    ht_wait_st();
                    hf_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                    half ^= 1;
                    if (half == 0) parity ^= 1u;       // both of the group's stages have been used once more
                }
            }
        }
    } else if (warp >= HT_WARP_MMA) {
        // =============================== MMA issuer (first warp of its warpgroup) ==================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HT_REGS_MMA));
        if (warp == HT_WARP_MMA) {
        if (lane == 0) {   // W1 image -> shared memory, once per CTA
            const uint64_t policy = l2_policy_evict_last();
            mbar_arrive_expect_tx(&b_bar, HF_BIMG_BYTES);
            for (int kb = 0; kb < HF_NKB; ++kb)
                bulk_g2s(bsm + (size_t)kb * 2 * HF_B_BYTES, img + (size_t)kb * 2 * HF_B_BYTES, 2 * HF_B_BYTES, &b_bar, policy);
            mbar_wait(&b_bar, 0);
        }
        __syncwarp();
        int acc = 0, st = 0;
        uint32_t acc_parity = 0, parity = 0;      // all stages flip together every HT_STAGES steps
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!hf_tile_has_rows(sel_rows, tile * HF_M, n_slots, lane)) continue;
            if (lane == 0) {
                mbar_wait(&tempty_bar[acc], acc_parity ^ 1u);
                hf_fence_after();
            }
            __syncwarp();
            const uint32_t tmem_d = tmem_base + HT_ACC_COL0 + acc * HF_H;
            for (int ks = 0; ks < HT_NKS; ++ks) {
                if (lane == 0) {
                    mbar_wait(&full_bar[st], parity);
                    hf_fence_after();
                    const uint32_t a0 = tmem_base + HT_A_COL0 + st * HT_STAGE_COLS, a1 = a0 + 16;
                    const uint32_t b0 = smem_u32(bsm + (size_t)(ks >> 1) * 2 * HF_B_BYTES) + (uint32_t)(ks & 1) * 64u;
                    const uint32_t bl = b0 + HF_B_BYTES;
#pragma unroll
                    for (int k16 = 0; k16 < HT_KS / 16; ++k16) {
                        const uint64_t db0 = hf_desc_sw128(b0 + k16 * 32), db1 = hf_desc_sw128(bl + k16 * 32);
                        ht_umma_ts(tmem_d, a0 + k16 * 8, db1, (ks | k16) != 0 ? 1u : 0u);
                        ht_umma_ts(tmem_d, a1 + k16 * 8, db0, 1u);
                        ht_umma_ts(tmem_d, a0 + k16 * 8, db0, 1u);
                    }
                    hf_commit(&empty_bar[st]);
                    if (ks == HT_NKS - 1) hf_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++st == HT_STAGES) { st = 0; parity ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
        }
    } else {
        // =============================== epilogue (warps 0-3): as in head_rows_f16_kernel ==========
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HT_REGS_EPI));
        const float a0 = (active_mask & MOC_CLS_TOPK) ? 1.f : 0.f;
        const float a1 = (active_mask & MOC_CLS_DELTA_SOFTMAX) ? 1.f : 0.f;
        const float a2 = (active_mask & MOC_CLS_DELTA_DIFF) ? 1.f : 0.f;
        const float a3 = (active_mask & MOC_CLS_BOTTOMK) ? 1.f : 0.f;
        int acc = 0;
        uint32_t acc_parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t slot0 = tile * HF_M;
            if (!hf_tile_has_rows(sel_rows, slot0, n_slots, lane)) continue;
            const int64_t slot = slot0 + warp * 32 + lane;
            int64_t row = -1;
            if (slot < n_slots) row = sel_rows ? (int64_t)sel_rows[slot] : slot;
            mbar_wait(&tfull_bar[acc], acc_parity);
            hf_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + HT_ACC_COL0 + acc * HF_H;
            float z[HF_G] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float d[32];
                hf_tmem_ld32(taddr + half * 32, d);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int jj = half * 32 + j;
                    const float h = relu_nan(fmaf(d[j], descale, b1s[jj]));
#pragma unroll
                    for (int m = 0; m < HF_G; ++m) z[m] = fmaf(h, w2s[m * HF_H + jj], z[m]);
                }
            }
            hf_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
            if (row < 0) continue;
            if (domain_flag != nullptr && !(fabsf(z[0] + z[1] + z[2] + z[3]) <= 3.0e38f)) atomicOr(domain_flag, 1);
            float g[HF_G];
#pragma unroll
            for (int m = 0; m < HF_G; ++m) g[m] = sigmoidf_exact(z[m] + b2s[m]);
            if (gate != nullptr) *reinterpret_cast<float4*>(gate + slot * HF_G) = make_float4(g[0], g[1], g[2], g[3]);
            if (final_scores == nullptr) continue;
            combine_row(keys + row, key_stride, C, g, a0, a1, a2, a3, final_scores + slot * C);
        }
    }

    hf_fence_before();
    __syncthreads();
    if (warp == HT_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)HT_TMEM_COLS)
                     : "memory");
    }
}

size_t head_f16_workspace_bytes() { return (size_t)HF_BIMG_BYTES + sizeof(HeadF16Tail); }

int launch_head_rows_f16(const float* feat, const float* keys, int64_t key_stride, int C, const int32_t* sel_rows,
                         int64_t n_slots, const float* w1, const float* b1, const float* w2, const float* b2,
                         unsigned active_mask, float* gate, float* final_scores, void* workspace, int* domain_flag,
                         cudaStream_t st) {
    unsigned char* img = reinterpret_cast<unsigned char*>(workspace);
    head_f16_prep_kernel<<<1, 1024, 0, st>>>(w1, img);
    MOC_LAUNCH_CHECK("head_f16_prep_kernel");
    MOC_CUDA(cudaFuncSetAttribute(head_rows_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HF_SMEM));
    const int64_t n_tiles = (n_slots + HF_M - 1) / HF_M;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    static int a_in_tmem = -1;          // MOC_HEAD_A=tmem|smem picks where the split patch tile lives
    if (a_in_tmem < 0) {
        const char* e = getenv("MOC_HEAD_A");
        a_in_tmem = (e && e[0] == 's') ? 0 : (e && e[0] == 't') ? 1 : HF_DEFAULT_A_IN_TMEM;
    }
    if (a_in_tmem) {
        MOC_CUDA(cudaFuncSetAttribute(head_rows_f16t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_SMEM));
        head_rows_f16t_kernel<<<grid, HT_THREADS, HT_SMEM, st>>>(feat, keys, key_stride, C, sel_rows, n_slots, img, b1, w2,
                                                                b2, active_mask, gate, final_scores, domain_flag);
        MOC_LAUNCH_CHECK("head_rows_f16t_kernel");
        return MOC_OK;
    }
    head_rows_f16_kernel<<<grid, HF_THREADS, HF_SMEM, st>>>(feat, keys, key_stride, C, sel_rows, n_slots, img, b1, w2, b2,
                                                           active_mask, gate, final_scores, domain_flag);
    MOC_LAUNCH_CHECK("head_rows_f16_kernel");
    return MOC_OK;
}

}  // namespace moc
