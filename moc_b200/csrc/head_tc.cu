// Gate MLP of the head on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate via 3xTF32.
//
// Replaces the first layer of senet (main_moc.py:303: Linear(512,64)) for every selected patch, fused with
// ReLU, the second layer, the sigmoid and the classifier-bank combination (main_moc.py:390-403).  This is the
// one dense contraction on the MOC path: [S,512] x [512,64] per slide, S ~ 1.7k (C=2) .. 13k (C=30).
//
// Precision: fp32 operands are split x = hi + lo with hi = x & 0xffffe000 (exactly TF32) and lo = x - hi
// (exact in fp32); D = A_hi B_hi + A_lo B_hi + A_hi B_lo accumulated in fp32 in TMEM.  The dropped lo*lo term
// and the TF32 truncation of the lo operands are ~2^-22 relative per product, i.e. fp32-level.
//
// Structure (persistent, one CTA per SM, 14 warps; a work item is a super-tile of 256 selected-row slots = two
// M128 row tiles that share every K-block of W1, which halves both the W1 bytes streamed from L2 per row and the
// number of producer -> MMA -> producer barrier round trips per row - with 128-row items and four 48 KB stages
// those round trips, not the gather or the tensor pipe, set the pace):
//   warps 0-3   epilogue: tcgen05.ld the two 128x64 accumulators (thread = one row of each), +b1, ReLU, 4x64
//               second layer, sigmoid, gated combination with the row's key planes, stores
//   warps 4-11  A producers: gather 128 B of 2 x 16 rows each per K-block straight from HBM/L2 (TC_PF K-blocks
//               of loads in flight per thread, across item boundaries), split hi/lo in registers, store into
//               the 128B-swizzled K-major smem tiles the UMMA descriptor describes
//   warp 12     MMA issuer (one elected lane): per K-block 4 k-steps x 2 row tiles x 3 products of
//               tcgen05.mma.kind::tf32 M128 N64 K8
//   warp 13     B copier: one 16 KB bulk copy per K-block of the pre-split, pre-swizzled W1 (hi|lo)
// Two smem stages of {A0_hi, A0_lo, A1_hi, A1_lo 16K each, B_hi 8K, B_lo 8K}; full/empty mbarriers; two pairs of
// TMEM accumulators so the epilogue of item i overlaps the MMAs of item i+1.
#include "common.cuh"

namespace moc {

constexpr int TC_H = MOC_HIDDEN;   // 64  (UMMA N)
constexpr int TC_G = MOC_GATES;
constexpr int TC_M = 128;          // rows per tile (UMMA M)
constexpr int TC_KB = 32;          // K elements per stage (one 128-byte swizzle row)
constexpr int TC_NKB = D / TC_KB;  // 16 K-blocks
constexpr int TC_SUB = 2;           // row tiles per work item
constexpr int TC_STAGES = 2;
constexpr int TC_PF = 2;            // K-block register sets (8 x 16 B each) a producer thread keeps in flight: 64 KB per CTA
constexpr int TC_A_BYTES = TC_M * 128;            // 16 KB per component
constexpr int TC_B_BYTES = TC_H * 128;            // 8 KB per component
constexpr int TC_STAGE_BYTES = TC_SUB * 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // 80 KB
constexpr int TC_EPI_WARPS = 4, TC_PROD_WARPS = 8;
constexpr int TC_WARP_MMA = TC_EPI_WARPS + TC_PROD_WARPS;  // 12
constexpr int TC_WARP_B = TC_WARP_MMA + 1;                 // 13
constexpr int TC_THREADS = (TC_WARP_B + 1) * 32;           // 448
constexpr int TC_TMEM_COLS = 256;                          // two pairs of 64-column accumulators
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024;  // + alignment slack

// tcgen05 instruction descriptor: D=f32, A=B=tf32, both K-major, N=64, M=128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_H >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// K-major, 128-byte swizzle: LBO = 1 (unused), SBO = 1024 B between 8-row groups, version 1 (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
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

// W1 [64][512] -> per K-block tile [hi 8 KB | lo 8 KB] in the swizzled K-major layout, ready for one bulk copy.
__global__ void head_tc_prep_kernel(const float* __restrict__ w1, float* __restrict__ w1_split) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk of W1
    if (i >= TC_H * (D / 4)) return;
    const int n = i / (D / 4), c4 = i % (D / 4);
    const int kb = c4 / 8, chunk = c4 % 8;
    const float4 v = reinterpret_cast<const float4*>(w1)[i];
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
    char* tile = reinterpret_cast<char*>(w1_split) + (size_t)kb * 2 * TC_B_BYTES;
    const int off = n * 128 + ((chunk ^ (n & 7)) << 4);
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + TC_B_BYTES + off) = lo;
}

__device__ __forceinline__ bool tile_has_rows(const int32_t* __restrict__ sel_rows, int64_t slot0, int64_t n_slots,
                                              int lane) {
    if (sel_rows == nullptr) return true;
    bool any = false;
#pragma unroll
    for (int i = 0; i < TC_SUB * TC_M / 32; ++i) {
        const int64_t s = slot0 + lane + 32 * i;
        any |= (s < n_slots) && (sel_rows[s] >= 0);
    }
    return __any_sync(FULL, any);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
head_rows_tc_kernel(const float* __restrict__ feat, const float* __restrict__ keys, int64_t key_stride, int C,
                    const int32_t* __restrict__ sel_rows, int64_t n_slots, const float* __restrict__ w1_split,
                    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                    unsigned active_mask, float* __restrict__ gate, float* __restrict__ final_scores) {
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float w2s[TC_G * TC_H], b1s[TC_H], b2s[TC_G];

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < TC_G * TC_H) w2s[tid] = w2[tid];
    if (tid < TC_H) b1s[tid] = b1[tid];
    if (tid < TC_G) b2s[tid] = b2[tid];
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], TC_PROD_WARPS + 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], TC_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == TC_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int64_t n_tiles = (n_slots + TC_SUB * TC_M - 1) / (TC_SUB * TC_M);   // work items

    if (warp >= TC_EPI_WARPS && warp < TC_WARP_MMA) {
        // =============================== A producers ===============================================
        // The CTA's non-empty tiles form one flat stream of (tile, K-block) steps.  A thread owns 16 bytes of 4
        // rows in every step and keeps TC_PF steps of gather loads in flight (a ring of register sets that
        // runs across tile boundaries): the gather is latency-bound, so bytes in flight are what buys bandwidth.
        const int pw = warp - TC_EPI_WARPS;      // 0..7 : rows 16*pw .. 16*pw+15 of the tile
        const int rsub = lane >> 3, chunk = lane & 7;
        uint32_t roff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = pw * 16 + i * 4 + rsub;
            roff[i] = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
        }
        const uint32_t smem_base = smem_u32(smem);
        // load cursor
        int64_t ltile = blockIdx.x;
        int lkb = 0;
        const float4* lsrc[TC_SUB * 4];
        auto seek = [&]() {  // move ltile to the next item with rows and fetch its row pointers
            while (ltile < n_tiles && !tile_has_rows(sel_rows, ltile * (TC_SUB * TC_M), n_slots, lane)) ltile += gridDim.x;
            if (ltile >= n_tiles) return;
#pragma unroll
            for (int q = 0; q < TC_SUB * 4; ++q) {
                const int64_t sl = ltile * (TC_SUB * TC_M) + (q >> 2) * TC_M + pw * 16 + (q & 3) * 4 + rsub;
                int64_t row = -1;
                if (sl < n_slots) row = sel_rows ? (int64_t)sel_rows[sl] : sl;
                lsrc[q] = row >= 0 ? reinterpret_cast<const float4*>(feat + row * D) + chunk : nullptr;
            }
        };
        auto issue = [&](float4 (&b)[TC_SUB * 4]) -> bool {  // load the cursor's step into b and advance; false when done
            if (ltile >= n_tiles) return false;
#pragma unroll
            for (int q = 0; q < TC_SUB * 4; ++q)
                b[q] = lsrc[q] ? __ldg(lsrc[q] + lkb * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (++lkb == TC_NKB) {
                lkb = 0;
                ltile += gridDim.x;
                seek();
            }
            return true;
        };
        float4 buf[TC_PF][TC_SUB * 4];
        bool pending[TC_PF];
        seek();
#pragma unroll
        for (int s = 0; s < TC_PF; ++s) pending[s] = issue(buf[s]);
        int stage = 0;
        uint32_t parity = 0;
        while (pending[0]) {
#pragma unroll
            for (int s = 0; s < TC_PF; ++s) {
                if (pending[s]) {
                    mbar_wait(&empty_bar[stage], parity ^ 1u);
                    const uint32_t st0 = smem_base + stage * TC_STAGE_BYTES;
#pragma unroll
                    for (int q = 0; q < TC_SUB * 4; ++q) {
                        const uint32_t a_hi = st0 + (q >> 2) * 2 * TC_A_BYTES, a_lo = a_hi + TC_A_BYTES;
                        const float4 cur = buf[s][q];
                        float4 hi, lo;
                        hi.x = __uint_as_float(__float_as_uint(cur.x) & 0xffffe000u);
                        hi.y = __uint_as_float(__float_as_uint(cur.y) & 0xffffe000u);
                        hi.z = __uint_as_float(__float_as_uint(cur.z) & 0xffffe000u);
                        hi.w = __uint_as_float(__float_as_uint(cur.w) & 0xffffe000u);
                        lo = make_float4(cur.x - hi.x, cur.y - hi.y, cur.z - hi.z, cur.w - hi.w);
                        sts128(a_hi + roff[q & 3], hi);   // explicit st.shared: the aligned-by-arithmetic base pointer would
                        sts128(a_lo + roff[q & 3], lo);   // otherwise compile to generic ST.E
                    }
                    fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                    if (++stage == TC_STAGES) { stage = 0; parity ^= 1u; }
                    pending[s] = issue(buf[s]);
                }
            }
        }
    } else if (warp == TC_WARP_B) {
        // =============================== B copier ==================================================
        const uint64_t policy = l2_policy_evict_last();
        int stage = 0;
        uint32_t parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!tile_has_rows(sel_rows, tile * (TC_SUB * TC_M), n_slots, lane)) continue;
            for (int kb = 0; kb < TC_NKB; ++kb) {
                if (lane == 0) {
                    mbar_wait(&empty_bar[stage], parity ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[stage], 2 * TC_B_BYTES);
                    bulk_g2s(smem + (size_t)stage * TC_STAGE_BYTES + TC_SUB * 2 * TC_A_BYTES,
                             reinterpret_cast<const char*>(w1_split) + (size_t)kb * 2 * TC_B_BYTES, 2 * TC_B_BYTES,
                             &full_bar[stage], policy);
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; parity ^= 1u; }
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // =============================== MMA issuer ================================================
        int stage = 0, acc = 0;
        uint32_t parity = 0, acc_parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!tile_has_rows(sel_rows, tile * (TC_SUB * TC_M), n_slots, lane)) continue;
            if (lane == 0) {
                mbar_wait(&tempty_bar[acc], acc_parity ^ 1u);  // epilogue has drained this accumulator pair
                tc_fence_after();
            }
            __syncwarp();
            const uint32_t tmem_d = tmem_base + acc * (TC_SUB * TC_H);
            for (int kb = 0; kb < TC_NKB; ++kb) {
                if (lane == 0) {
                    mbar_wait(&full_bar[stage], parity);
                    tc_fence_after();
                    const uint32_t st0 = smem_u32(smem + (size_t)stage * TC_STAGE_BYTES);
                    const uint32_t b_hi = st0 + TC_SUB * 2 * TC_A_BYTES;
                    const uint32_t b_lo = b_hi + TC_B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_KB / 8; ++ks) {
                        const uint32_t o = ks * 32;  // 8 tf32 = 32 bytes along K inside the swizzled row
                        const uint64_t dbh = umma_desc_sw128(b_hi + o), dbl = umma_desc_sw128(b_lo + o);
#pragma unroll
                        for (int sub = 0; sub < TC_SUB; ++sub) {
                            const uint32_t a_hi = st0 + sub * 2 * TC_A_BYTES, a_lo = a_hi + TC_A_BYTES;
                            const uint64_t dah = umma_desc_sw128(a_hi + o), dal = umma_desc_sw128(a_lo + o);
                            umma_tf32(tmem_d + sub * TC_H, dal, dbh, (kb | ks) != 0 ? 1u : 0u);
                            umma_tf32(tmem_d + sub * TC_H, dah, dbl, 1u);
                            umma_tf32(tmem_d + sub * TC_H, dah, dbh, 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);               // stage free once these MMAs have read it
                    if (kb == TC_NKB - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; parity ^= 1u; }
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
            const int64_t slot0 = tile * (TC_SUB * TC_M);
            if (!tile_has_rows(sel_rows, slot0, n_slots, lane)) continue;
            mbar_wait(&tfull_bar[acc], acc_parity);
            tc_fence_after();
            float z[TC_SUB][TC_G];
#pragma unroll
            for (int sub = 0; sub < TC_SUB; ++sub) {
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * (TC_SUB * TC_H) + sub * TC_H;
#pragma unroll
                for (int m = 0; m < TC_G; ++m) z[sub][m] = 0.f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float d[32];
                    tmem_ld32(taddr + half * 32, d);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int jj = half * 32 + j;
                        const float h = relu_nan(d[j] + b1s[jj]);
#pragma unroll
                        for (int m = 0; m < TC_G; ++m) z[sub][m] = fmaf(h, w2s[m * TC_H + jj], z[sub][m]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);  // accumulators are in registers: MMA may reuse them
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
#pragma unroll
            for (int sub = 0; sub < TC_SUB; ++sub) {
                const int64_t slot = slot0 + sub * TC_M + warp * 32 + lane;
                int64_t row = -1;
                if (slot < n_slots) row = sel_rows ? (int64_t)sel_rows[slot] : slot;
                if (row < 0) continue;
                float g[TC_G];
#pragma unroll
                for (int m = 0; m < TC_G; ++m) g[m] = sigmoidf_exact(z[sub][m] + b2s[m]);
                if (gate != nullptr) *reinterpret_cast<float4*>(gate + slot * TC_G) = make_float4(g[0], g[1], g[2], g[3]);
                if (final_scores == nullptr) continue;
                combine_row(keys + row, key_stride, C, g, a0, a1, a2, a3, final_scores + slot * C);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

size_t head_tc_workspace_bytes() { return (size_t)TC_NKB * 2 * TC_B_BYTES; }  // 256 KB: W1 split + swizzled

int launch_head_rows_tc(const float* feat, const float* keys, int64_t key_stride, int C, const int32_t* sel_rows,
                        int64_t n_slots, const float* w1, const float* b1, const float* w2, const float* b2,
                        unsigned active_mask, float* gate, float* final_scores, void* workspace, cudaStream_t st) {
    float* w1_split = reinterpret_cast<float*>(workspace);
    head_tc_prep_kernel<<<(TC_H * (D / 4) + 255) / 256, 256, 0, st>>>(w1, w1_split);
    MOC_LAUNCH_CHECK("head_tc_prep_kernel");
    MOC_CUDA(cudaFuncSetAttribute(head_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    const int64_t n_tiles = (n_slots + TC_SUB * TC_M - 1) / (TC_SUB * TC_M);
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    head_rows_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(feat, keys, key_stride, C, sel_rows, n_slots, w1_split, b1, w2,
                                                           b2, active_mask, gate, final_scores);
    MOC_LAUNCH_CHECK("head_rows_tc_kernel");
    return MOC_OK;
}

}  // namespace moc
