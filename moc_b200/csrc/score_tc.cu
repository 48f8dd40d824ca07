// Streaming patch-prompt scoring for WIDE prompt sets (9..64 columns, e.g. EBRAINS-30: 30 classes + 4
// background prompts) on the tensor cores.  Same contract as score_keys.cu (main_moc.py:336-337 + the per-row
// key arithmetic); with this many columns the path is a dense contraction (2*512*cols FLOP per 2 KB patch) and
// CUDA-core FMAs cannot keep up with HBM.
//
// Precision: fp32-level through a three-product FP16 decomposition,
//     x = a0 + a1,  w*2^SW = b0 + b1,   a0 = fp16(x), a1 = fp16(x - a0), likewise b   (a1 may be an FP16 subnormal:
//     the split is exact to max(2^-22 |x|, 2^-25)),
//     score = 2^-SW * ( a0.b0 + a1.b0 + a0.b1 ),  fp32 accumulation in TMEM.
// FP16 rather than TF32 because both split operands of the prompt matrix must stay resident in shared memory
// next to the patch stages: 64 columns x 512 x (2+2) B = 128 KB (TF32 would need 256 KB), and the MMA count and
// the shared-memory read traffic per patch halve (K=16 per tcgen05.mma).  Patches are split unscaled, so |x| must
// stay below 65504 (CONCH embeddings are O(1)); beyond that - or for non-finite inputs, where the reference
// produces non-finite scores too - the scores come out non-finite and the kernel raises a flag in the prompt
// image that the host can poll (moc_prompts_tc_flag_offset).  The prompt scale 2^SW is chosen on the device from
// max|w| when the image is prepared, so no call synchronises.
//
// Structure: persistent CTA per SM, 13 warps.
//   warps 4-11  producers.  Each owns a private ring of 1 KB shared-memory slots (4 patches x one 256-byte
//               K-block piece) that it fills itself with bulk copies (cp.async.bulk -> UBLKCP, mbarrier
//               complete_tx, L2 evict-first): the copy engine keeps the whole ring (all the shared memory the
//               prompt image leaves free, 50-80 KB per SM) in flight, which plain LDG cannot (the LSU caps
//               outstanding misses: a register-prefetch version of this kernel stalled at 2.8 TB/s).  The warp
//               then reads its slots (conflict-free LDS.128), scales, splits into (a0, a1) and stores them into
//               the 128B-swizzled K-major A stage the UMMA descriptors describe.
//   warp 12     MMA issuer (one elected lane), two TMEM accumulators.
//   warps 0-3   epilogue, thread = patch: softmax / top-2 / background sum+max, key planes written as full
//               128-byte lines.
// The prompt tile is [b0 ; b1] stacked along N (b1 starts at row NPa = round16(cols)); a0 x [b0 ; b1] is ONE MMA of
// width 2 NPa (accumulator columns [0, NPa) = a0 b0, [NPa, 2 NPa) = a0 b1) and a1 x b0 a second one of width NPa
// into the first half; the epilogue adds the two halves.  (Three MMAs of width NPa into the same columns - the
// first version - read the a0 tile twice; measured, the two forms run at the same speed: the producers bound it.)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stddef.h>
#include <stdlib.h>

#include "common.cuh"

namespace moc {

constexpr int ST_M = 128;                 // patches per tile
constexpr int ST_KB = 64;                 // K elements per stage (128 B of fp16)
constexpr int ST_NKB = D / ST_KB;         // 8
constexpr int ST_A_BYTES = ST_M * 128;    // 16 KB per component
constexpr int ST_STAGE_BYTES = 2 * ST_A_BYTES;
constexpr int ST_EPI_WARPS = 4, ST_PROD_WARPS = 8;
constexpr int ST_WARP_MMA = ST_EPI_WARPS + ST_PROD_WARPS;
constexpr int ST_THREADS = (ST_WARP_MMA + 1) * 32;  // 416
constexpr int ST_SX = 0;                  // patches are split unscaled (|x| < 65504); only the prompts are scaled
constexpr int ST_SLOT_ROWS = 8;           // patches per raw slot
constexpr int ST_SLOT_BYTES = ST_SLOT_ROWS * ST_KB * 4;  // 2 KB
constexpr int ST_MAX_SLOTS = 8;           // slots per producer warp (ring), upper bound
constexpr int ST_A_STAGES = 2;

__device__ __forceinline__ uint64_t st_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t st_idesc_f16(int n) {  // D=f32, A=B=f16, K-major, M=128
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(ST_M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void st_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void st_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_tmem_ld32(uint32_t taddr, float (&v)[32]) {
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

__device__ __forceinline__ void st_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 2-D tiled tensor copy global -> shared (SASS: UTMALDG), completion counted in bytes on the mbarrier at `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
        : "memory");
}

// fp32 x4 -> (a0, a1): two 8-byte half-chunks of 4 halves with x = a0 + a1 up to 2^-25 absolute (a1 may be an
// FP16 subnormal) or 2^-22 relative, whichever is larger.
__device__ __forceinline__ void split4(const float4& v, uint2& c0, uint2& c1) {
    const __half2 a01 = __floats2half2_rn(v.x, v.y), a23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(a01), f23 = __half22float2(a23);
    const __half2 b01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), b23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    c0 = make_uint2(*reinterpret_cast<const uint32_t*>(&a01), *reinterpret_cast<const uint32_t*>(&a23));
    c1 = make_uint2(*reinterpret_cast<const uint32_t*>(&b01), *reinterpret_cast<const uint32_t*>(&b23));
}

// Tail of the prompt image (after the tiles): {float scale = 2^SW, float descale = 2^-SW, int flag, pad}.
struct ScoreTcTail {
    float scale, descale;
    int flag, pad;
};

struct ScoreTcGeom {
    int npa;       // row of b1 inside a tile = columns rounded up to 8
    int n_wide;    // rows of one K-block tile: [b0 ; b1] rounded up to 16
    int n_narrow;  // MMA width (N) of each of the three products
    __host__ __device__ size_t tile_bytes() const { return (size_t)n_wide * 128; }
    __host__ __device__ size_t b_bytes() const { return (size_t)ST_NKB * tile_bytes(); }
};
__host__ __device__ inline ScoreTcGeom score_tc_geom(int n_cols) {
    ScoreTcGeom g;
    g.npa = (n_cols + 15) & ~15;     // a multiple of 16: the narrow MMA (N = npa) must stop exactly where b1 starts
    g.n_wide = 2 * g.npa;
    g.n_narrow = g.npa;
    return g;
}

// One block: SW = 14 - floor(log2(max|w|)), so that max|w| * 2^SW lies in [2^14, 2^15).
__global__ void score_tc_scale_kernel(const float* __restrict__ packed, int n, ScoreTcTail* __restrict__ tail) {
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = fabsf(packed[i]);
        if (a <= 3.0e38f) m = fmaxf(m, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
        int sw = 0;
        if (m > 0.f) sw = 14 - ilogbf(m);
        sw = sw > 100 ? 100 : (sw < -100 ? -100 : sw);
        tail->scale = ldexpf(1.0f, sw);
        tail->descale = ldexpf(1.0f, -(ST_SX + sw));
        tail->flag = 0;
        tail->pad = 0;
    }
}

// packed fp32 K-major prompts [cols_pad][512] -> per K-block (64) tile of n_wide rows x 128 B: rows [0,npa) b0,
// rows [npa, 2 npa) b1, 128B-swizzled, scaled by 2^SW (the image is zero-filled beforehand).
__global__ void score_tc_prep_kernel(const float* __restrict__ packed, int n_cols, ScoreTcGeom g,
                                     const ScoreTcTail* __restrict__ tail, unsigned char* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (col n, 8-byte half-chunk of 4 k-elements)
    if (i >= n_cols * (D / 4)) return;
    const int n = i / (D / 4), qq = i % (D / 4);
    const int kb = qq / 16, q = qq % 16;
    const float4 v = reinterpret_cast<const float4*>(packed + (size_t)n * D)[qq];
    const float sc = tail->scale;
    uint2 c0, c1;
    split4(make_float4(v.x * sc, v.y * sc, v.z * sc, v.w * sc), c0, c1);
    unsigned char* tile = out + (size_t)kb * g.tile_bytes();
    const int n1 = g.npa + n;
    *reinterpret_cast<uint2*>(tile + n * 128 + (((q >> 1) ^ (n & 7)) << 4) + ((q & 1) << 3)) = c0;
    *reinterpret_cast<uint2*>(tile + n1 * 128 + (((q >> 1) ^ (n1 & 7)) << 4) + ((q & 1) << 3)) = c1;
}

// PW = producer warps.  8: one CTA of 13 warps, 128 registers each (the first version).  16: twice the producers, each
// converting 8 instead of 16 patches per K-block step - the producers' LDS -> split -> STS chain is latency-bound with
// two warps per scheduler - in a CTA of six warpgroups launched at 80 registers per thread whose pool setmaxnreg
// redistributes: producers 64, the MMA issuer's group 40, the epilogue (64 live scores per thread) 160.
template <int PW> struct StRoles {
    static constexpr int WARP_MMA = ST_EPI_WARPS + PW;
    static constexpr int THREADS = (PW == 16 ? WARP_MMA + 4 : WARP_MMA + 1) * 32;   // 768 / 416
    static constexpr int ROWS_PER_WARP = ST_M / PW;                                  // 8 / 16
    static constexpr int SLOTS_PER_STEP = ROWS_PER_WARP / ST_SLOT_ROWS;              // 1 / 2
};
constexpr int ST_REGS_PROD16 = 64, ST_REGS_MMA16 = 40, ST_REGS_EPI16 = 160;

template <int NCHUNK, bool NORM, int PW>  // NCHUNK = ceil(n_cols/32): 1 or 2
__global__ void __launch_bounds__(StRoles<PW>::THREADS, 1)
score_keys_tc_kernel(const __grid_constant__ CUtensorMap feat_map, const float* __restrict__ feat, int64_t n_rows,
                     const unsigned char* __restrict__ btiles,
                     int n_classes, int n_cols, ScoreTcGeom g, int ring_slots, float* __restrict__ keys,
                     int64_t key_stride, ScoreTcTail* __restrict__ tail) {
    extern __shared__ unsigned char st_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[ST_A_STAGES], empty_bar[ST_A_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ __align__(8) uint64_t raw_bar[PW][ST_MAX_SLOTS];
    __shared__ uint32_t tmem_base_s;
    constexpr int WARP_MMA = StRoles<PW>::WARP_MMA, THREADS = StRoles<PW>::THREADS;
    constexpr int RPW = StRoles<PW>::ROWS_PER_WARP, SPS = StRoles<PW>::SLOTS_PER_STEP;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b_bytes = (int)g.b_bytes();
    unsigned char* bsm = smem;                                    // resident prompt tiles
    unsigned char* asm_ = smem + b_bytes;                         // A stages (a0 | a1)
    unsigned char* rawsm = asm_ + ST_A_STAGES * ST_STAGE_BYTES;   // per-warp raw fp32 rings
    const int tmem_cols = 256;                                    // two accumulators of n_wide (<= 128) columns

    // resident prompt tiles (already swizzled): plain copy, then make them visible to the async proxy
    for (int i = tid; i < b_bytes / 16; i += THREADS)
        reinterpret_cast<uint4*>(bsm)[i] = __ldg(reinterpret_cast<const uint4*>(btiles) + i);
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int s = 0; s < ST_A_STAGES; ++s) {
            mbar_init(&full_bar[s], PW);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], ST_EPI_WARPS);
        }
        for (int w = 0; w < PW; ++w)
            for (int s = 0; s < ring_slots; ++s) mbar_init(&raw_bar[w][s], 1);
        fence_mbar_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    st_fence_before();
    __syncthreads();
    st_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int acc_cols = tmem_cols / 2;
    const int64_t n_tiles = (n_rows + ST_M - 1) / ST_M;

    if (warp >= ST_EPI_WARPS && warp < WARP_MMA) {
        // =============================== producers ================================================
        if (PW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ST_REGS_PROD16));
        // The CTA's tiles form one flat stream of (tile, K-block) steps; in a step a warp converts the 256-byte
        // K-block piece of its 16 patches = 2 slots of 8 patches.  Slot u of the warp's stream lands in ring
        // position u % ring_slots; as soon as the warp has read a slot it re-arms it for slot u + ring_slots.
        const int pw = warp - ST_EPI_WARPS, rhalf = lane >> 4, q = lane & 15;
        const uint64_t policy = l2_policy_evict_first();
        const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t n_slots_total = my_tiles * ST_NKB * SPS;
        const uint32_t ring = smem_u32(rawsm) + (uint32_t)(pw * ring_slots * ST_SLOT_BYTES);
        const uint32_t bars = smem_u32(&raw_bar[pw][0]);
        // issue cursor: one elected lane arms the slot's barrier and launches ONE 2-D tensor copy (8 patches x
        // 64 floats; rows past the end of feat are zero-filled by the copy engine)
        int64_t i_tile = blockIdx.x, i_left = n_slots_total;
        int i_sub = 0;  // (kb, slot of the step) = (i_sub / SPS, i_sub % SPS) inside the tile
        auto issue = [&](int pos) {
            if (lane == 0) {
                const int64_t row0 = i_tile * ST_M + pw * RPW + (i_sub % SPS) * ST_SLOT_ROWS;
                const uint32_t bar = bars + pos * 8;
                mbar_arrive_expect_tx_a(bar, ST_SLOT_BYTES);
                tma_load_2d(ring + pos * ST_SLOT_BYTES, &feat_map, (i_sub / SPS) * ST_KB, (int)row0, bar, policy);
            }
            if (++i_sub == SPS * ST_NKB) { i_sub = 0; i_tile += gridDim.x; }
            --i_left;
        };
        for (int s = 0; s < ring_slots; ++s)
            if (i_left > 0) issue(s);
        uint32_t roff[4 * SPS];
#pragma unroll
        for (int i = 0; i < 4 * SPS; ++i) {
            const int r = pw * RPW + i * 2 + rhalf;
            roff[i] = (uint32_t)(r * 128 + (((q >> 1) ^ (r & 7)) << 4) + ((q & 1) << 3));
        }
        const uint32_t a_base = smem_u32(asm_);
        const uint32_t lds_off = (uint32_t)(rhalf * (ST_KB * 4) + q * 16);
        int stage = 0, pos = 0;
        uint32_t parity = 0, rparity = 0;
        for (int64_t step = 0; step < my_tiles * ST_NKB; ++step) {
            uint2 c0[4 * SPS], c1[4 * SPS];
#pragma unroll
            for (int j = 0; j < SPS; ++j) {
                mbar_wait_a(bars + pos * 8, rparity);
                const uint32_t sp = ring + pos * ST_SLOT_BYTES + lds_off;
#pragma unroll
                for (int i = 0; i < 4; ++i) split4(lds128(sp + i * 2 * (ST_KB * 4)), c0[4 * j + i], c1[4 * j + i]);
                __syncwarp();  // every lane has read the slot: hand it back to the copy engine
                if (i_left > 0) issue(pos);
                if (++pos == ring_slots) { pos = 0; rparity ^= 1u; }
            }
            mbar_wait(&empty_bar[stage], parity ^ 1u);
            const uint32_t a0 = a_base + stage * ST_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 4 * SPS; ++i) {
                sts64(a0 + roff[i], c0[i]);
                sts64(a0 + ST_A_BYTES + roff[i], c1[i]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == ST_A_STAGES) { stage = 0; parity ^= 1u; }
        }
    } else if (warp >= WARP_MMA) {
        // =============================== MMA issuer (first warp of its group) ======================
        if (PW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ST_REGS_MMA16));
        if (warp == WARP_MMA) {
        // a0 x [b0 ; b1] as ONE MMA of width 2 npa (columns [0, npa) = a0 b0, [npa, 2 npa) = a0 b1), then a1 x b0 of
        // width npa accumulating into the first half: the a0 tile is read from shared memory once instead of twice
        // (the kernel is bound by shared-memory bandwidth: TMA writes, the producers' LDS / STS and these operand
        // reads share 128 B/clk), and there are two MMAs per k-step instead of three.  The epilogue adds the halves.
        const uint32_t idesc_wide = st_idesc_f16(g.n_wide), idesc_narrow = st_idesc_f16(g.n_narrow);
        int stage = 0, acc = 0;
        uint32_t parity = 0, acc_parity = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (lane == 0) {
                mbar_wait(&tempty_bar[acc], acc_parity ^ 1u);
                st_fence_after();
            }
            __syncwarp();
            const uint32_t tmem_d = tmem_base + acc * acc_cols;
            for (int kb = 0; kb < ST_NKB; ++kb) {
                if (lane == 0) {
                    mbar_wait(&full_bar[stage], parity);
                    st_fence_after();
                    const uint32_t a0 = smem_u32(asm_ + (size_t)stage * ST_STAGE_BYTES);
                    const uint32_t a1 = a0 + ST_A_BYTES;
                    const uint32_t bt = smem_u32(bsm + (size_t)kb * g.tile_bytes());
#pragma unroll
                    for (int ks = 0; ks < ST_KB / 16; ++ks) {
                        const uint32_t o = ks * 32;  // 16 halves = 32 bytes along K inside the swizzled row
                        const uint64_t da0 = st_desc_sw128(a0 + o), da1 = st_desc_sw128(a1 + o);
                        const uint64_t db = st_desc_sw128(bt + o);
                        umma_f16(tmem_d, da0, db, idesc_wide, (kb | ks) != 0 ? 1u : 0u);
                        umma_f16(tmem_d, da1, db, idesc_narrow, 1u);
                    }
                    st_commit(&empty_bar[stage]);
                    if (kb == ST_NKB - 1) st_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == ST_A_STAGES) { stage = 0; parity ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
        }
    } else {
        // =============================== epilogue (warps 0-3): thread = patch ========================
        if (PW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ST_REGS_EPI16));
        const int C = n_classes;
        const KeyLayout kl = key_layout(C);
        const float descale = tail->descale;
        int acc = 0;
        uint32_t acc_parity = 0;
        bool bad = false;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t row = tile * ST_M + warp * 32 + lane;
            mbar_wait(&tfull_bar[acc], acc_parity);
            st_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * acc_cols;
            // Columns are handled in groups of 8 behind warp-uniform guards: whole groups of classes run without
            // per-element predicates (C = 30 -> three of four), only the group straddling C / n_cols is predicated.
            // column c of the scores = (a0 b0 + a1 b0)[c] + (a0 b1)[c]: accumulator columns c and npa + c.  npa is a
            // multiple of 16, so the second half starts at 16-column granularity: it is read with x16 loads.
            float v[NCHUNK * 32];
#pragma unroll
            for (int q = 0; q < NCHUNK * 2; ++q) {
                if (q * 16 < n_cols) {
                    float d0[16], d1[16];
                    st_tmem_ld16(taddr + q * 16, d0);
                    st_tmem_ld16(taddr + g.npa + q * 16, d1);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[q * 16 + e] = (d0[e] + d1[e]) * descale;
                }
            }
            st_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
            if (row >= n_rows) continue;
            if (NORM) {
                // |x|^2 is not a prompt column here: read the patch once more (mostly L2: it was just streamed)
                const float4* xp = reinterpret_cast<const float4*>(feat + row * D);
                float ss = 0.f;
                for (int i = 0; i < D / 4; ++i) {
                    const float4 t = __ldg(xp + i);
                    ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
                }
                const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
                for (int i0 = 0; i0 < NCHUNK * 32; i0 += 8) {
                    if (i0 < n_cols) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[i0 + e] *= inv;
                    }
                }
            }
            // pass 1: top-2 over the classes, sum / max over the background, raw similarities out, finiteness probe
            // (0 * v is NaN exactly when v is NaN or infinite)
            float m1 = -INFINITY, m2 = -INFINITY, bsum = 0.f, bmax = -INFINITY, probe = 0.f;
            float* kp = keys + row;
#pragma unroll
            for (int c0 = 0; c0 < NCHUNK * 32; c0 += 8) {
                if (c0 < n_cols) {
                    if (c0 + 8 <= C) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float x = v[c0 + e];
                            probe = fmaf(x, 0.f, probe);
                            m2 = fmaxf(m2, fminf(m1, x));
                            m1 = fmaxf(m1, x);
                            kp[(int64_t)(c0 + e) * key_stride] = x;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = c0 + e;
                            const float x = v[c];
                            if (c < C) {
                                probe = fmaf(x, 0.f, probe);
                                m2 = fmaxf(m2, fminf(m1, x));
                                m1 = fmaxf(m1, x);
                                kp[(int64_t)c * key_stride] = x;
                            } else if (c < n_cols) {
                                probe = fmaf(x, 0.f, probe);
                                bsum += x;
                                bmax = fmaxf(bmax, x);
                            }
                        }
                    }
                }
            }
            bad |= (probe != probe);
            // pass 2: exp(v - max) and its sum; pass 3: the normalised softmax planes
            float esum = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < NCHUNK * 32; c0 += 8) {
                if (c0 < C) {
                    if (c0 + 8 <= C) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            v[c0 + e] = expf(v[c0 + e] - m1);
                            esum += v[c0 + e];
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (c0 + e < C) {
                                v[c0 + e] = expf(v[c0 + e] - m1);
                                esum += v[c0 + e];
                            }
                        }
                    }
                }
            }
            const float inv_sum = 1.0f / esum;
            if (kl.compact) {
                // wide class sets: the softmax planes are not stored, only what rebuilds them, expf(L - lse)
                // (C + 4 planes: 136 instead of 252 bytes per patch at C = 30)
                kp[(int64_t)kl.lse * key_stride] = lse_of(m1, esum);
            } else {
                float* ks = kp + (int64_t)C * key_stride;
#pragma unroll
                for (int c0 = 0; c0 < NCHUNK * 32; c0 += 8) {
                    if (c0 < C) {
                        if (c0 + 8 <= C) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) ks[(int64_t)(c0 + e) * key_stride] = v[c0 + e] * inv_sum;
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (c0 + e < C) ks[(int64_t)(c0 + e) * key_stride] = v[c0 + e] * inv_sum;
                        }
                    }
                }
            }
            kp[(int64_t)kl.diff * key_stride] = fabsf(m1 - m2);
            kp[(int64_t)kl.bg_sum * key_stride] = bsum;
            kp[(int64_t)kl.bg_max * key_stride] = bmax;
        }
        if (bad) atomicExch(&tail->flag, 1);
    }

    st_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols)
                     : "memory");
    }
}

// =====================================================================================================
// Un-collapsed prompt BANK (BASELINE.json configs[2]: >= 64 prompts per class kept as W_bank [512, C*P]).
//
// The reference collapses a bank offline to one unit column per class (utils/zeroshot_utils.py:38-44: normalise every
// prompt, mean over the class, renormalise) and scores against that.  Scoring is linear, so the class logit equals
//     L_c = ( sum_{p in class c} x . w^_p ) / ( P_c * || mean_p w^_p || )            w^_p = prompt p, L2-normalised
// This kernel keeps the bank un-collapsed: every prompt is a column of the tensor-core contraction (the dense
// stress case of SURVEY.md section 8d: 2*512*(C*P + n_bg) FLOP per 2 KB patch, ~100 FLOP/B - tensor-bound, not
// HBM-bound) and the group sum + the 1/(P_c ||mean||) rescale happen in the epilogue; keys come out in the usual
// layout.  Same FP16x3 operand split, same producers as score_keys_tc_kernel.  Differences:
//   * up to 256 columns: one tcgen05.mma of N = round16(columns); two TMEM accumulators of 256 columns each
//   * the split prompt image ([b0 ; b1] per 64-wide K-block: 2 * N * 128 B, 52 KB at 196 columns) no longer fits
//     shared memory for all eight K-blocks: a B-loader warp streams one K-block tile per bulk copy (UBLKCP, L2
//     evict-last: all CTAs read the same 416 KB image) into a two-stage ring, in step with the A stages
//   * epilogue: thread = patch; the accumulator is read 32 columns at a time; a chunk that lies inside one class
//     (almost all do) is tree-summed and added to that class's sum, boundary chunks go element by element
static_assert(MOC_BANK_MAX_CLASSES < MOC_KEYS_COMPACT_MIN_CLASSES, "the bank kernel writes the full 2C+3-plane key layout");
struct BankTail {                      // after the eight K-block tiles of the image
    float scale, descale;              // same first 16 bytes as ScoreTcTail (score_tc_scale_kernel writes them)
    int flag, pad;
    int col_off[MOC_BANK_MAX_CLASSES + 1];   // first column of every class; col_off[C] = number of bank prompts
    float cls_scale[MOC_BANK_MAX_CLASSES];   // 1 / (P_c * ||mean_p w^_p||)
};

struct BankGeom {
    int n_cols;    // bank prompts + background prompts
    int npa;       // columns rounded up to 16 = MMA width N = row of b1 inside a K-block tile
    __host__ __device__ size_t tile_bytes() const { return (size_t)(2 * npa) * 128; }
    __host__ __device__ size_t b_bytes() const { return (size_t)ST_NKB * tile_bytes(); }
    __host__ __device__ size_t packed_offset() const { return (b_bytes() + sizeof(BankTail) + 255) & ~(size_t)255; }
    __host__ __device__ size_t image_bytes() const { return packed_offset() + (size_t)n_cols * D * sizeof(float); }
};
__host__ __device__ inline BankGeom bank_geom(int n_prompts, int n_bg) {
    BankGeom g;
    g.n_cols = n_prompts + n_bg;
    g.npa = (g.n_cols + 15) & ~15;
    return g;
}

constexpr int SB_B_STAGES = 2;
constexpr int SB_ACC_COLS = 256;
constexpr int SB_WARP_B = ST_WARP_MMA + 1;          // 13: B loader
constexpr int SB_THREADS = (SB_WARP_B + 1) * 32;    // 448

// rows of the packed K-major matrix the image is built from: bank prompts L2-normalised (F.normalize, eps 1e-12),
// background prompts as given.  One block of 512 threads per column.
__global__ void __launch_bounds__(D) bank_pack_kernel(const float* __restrict__ bank, int n_prompts,
                                                     const float* __restrict__ bg, int n_bg,
                                                     float* __restrict__ packed) {
    __shared__ float red[D / 32];
    __shared__ float total;
    const int col = blockIdx.x, k = threadIdx.x;
    if (col >= n_prompts) {
        packed[(size_t)col * D + k] = bg[(size_t)(col - n_prompts) * D + k];
        return;
    }
    const float v = bank[(size_t)col * D + k];
    const float s = warp_sum(v * v);
    if ((k & 31) == 0) red[k >> 5] = s;
    __syncthreads();
    if (k == 0) {
        float t = 0.f;
        for (int w = 0; w < D / 32; ++w) t += red[w];
        total = fmaxf(sqrtf(t), 1e-12f);
    }
    __syncthreads();
    packed[(size_t)col * D + k] = v / total;
}

// One block of 512 threads per class: || mean of the class's normalised prompts ||, column range, rescale factor.
__global__ void __launch_bounds__(D) bank_class_scale_kernel(const float* __restrict__ packed,
                                                            const int32_t* __restrict__ class_offsets, int n_classes,
                                                            BankTail* __restrict__ tail) {
    __shared__ float red[D / 32];
    const int c = blockIdx.x, k = threadIdx.x;
    const int p0 = class_offsets[c], p1 = class_offsets[c + 1];
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) acc += packed[(size_t)p * D + k];
    const float m = acc / (float)(p1 - p0);
    const float s = warp_sum(m * m);
    if ((k & 31) == 0) red[k >> 5] = s;
    __syncthreads();
    if (k == 0) {
        float t = 0.f;
        for (int w = 0; w < D / 32; ++w) t += red[w];
        tail->col_off[c] = p0;
        if (c == n_classes - 1) tail->col_off[n_classes] = p1;
        tail->cls_scale[c] = 1.0f / ((float)(p1 - p0) * sqrtf(t));
    }
}

template <bool NORM>
__global__ void __launch_bounds__(SB_THREADS, 1)
score_bank_tc_kernel(const __grid_constant__ CUtensorMap feat_map, const float* __restrict__ feat, int64_t n_rows,
                     const unsigned char* __restrict__ image, int n_classes, BankGeom g, int ring_slots,
                     float* __restrict__ keys, int64_t key_stride, BankTail* __restrict__ tail) {
    extern __shared__ unsigned char sb_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[ST_A_STAGES], empty_bar[ST_A_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ __align__(8) uint64_t bfull_bar[SB_B_STAGES], bempty_bar[SB_B_STAGES];
    __shared__ __align__(8) uint64_t raw_bar[ST_PROD_WARPS][ST_MAX_SLOTS];
    __shared__ uint32_t tmem_base_s;
    __shared__ int col_off_s[MOC_BANK_MAX_CLASSES + 1];
    __shared__ float cls_scale_s[MOC_BANK_MAX_CLASSES];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sb_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tile_bytes = (uint32_t)g.tile_bytes();
    unsigned char* bsm = smem;                                         // B ring: SB_B_STAGES K-block tiles
    unsigned char* asm_ = smem + (size_t)SB_B_STAGES * tile_bytes;     // A stages (a0 | a1)
    unsigned char* rawsm = asm_ + ST_A_STAGES * ST_STAGE_BYTES;        // per-warp raw fp32 rings
    constexpr int tmem_cols = 2 * SB_ACC_COLS;                         // the whole tensor memory

    if (tid <= n_classes) col_off_s[tid] = tail->col_off[tid];
    if (tid < n_classes) cls_scale_s[tid] = tail->cls_scale[tid];
    if (tid == 0) {
        for (int s = 0; s < ST_A_STAGES; ++s) {
            mbar_init(&full_bar[s], ST_PROD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < SB_B_STAGES; ++s) {
            mbar_init(&bfull_bar[s], 1);
            mbar_init(&bempty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], ST_EPI_WARPS);
        }
        for (int w = 0; w < ST_PROD_WARPS; ++w)
            for (int s = 0; s < ring_slots; ++s) mbar_init(&raw_bar[w][s], 1);
        fence_mbar_init();
    }
    if (warp == ST_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    st_fence_before();
    __syncthreads();
    st_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int64_t n_tiles = (n_rows + ST_M - 1) / ST_M;
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp >= ST_EPI_WARPS && warp < ST_WARP_MMA) {
        // =============================== producers (as in score_keys_tc_kernel) ====================
        const int pw = warp - ST_EPI_WARPS, rhalf = lane >> 4, q = lane & 15;
        const uint64_t policy = l2_policy_evict_first();
        const int64_t n_slots_total = my_tiles * ST_NKB * 2;
        const uint32_t ring = smem_u32(rawsm) + (uint32_t)(pw * ring_slots * ST_SLOT_BYTES);
        const uint32_t bars = smem_u32(&raw_bar[pw][0]);
        int64_t i_tile = blockIdx.x, i_left = n_slots_total;
        int i_sub = 0;
        auto issue = [&](int pos) {
            if (lane == 0) {
                const int64_t row0 = i_tile * ST_M + pw * 16 + (i_sub & 1) * ST_SLOT_ROWS;
                const uint32_t bar = bars + pos * 8;
                mbar_arrive_expect_tx_a(bar, ST_SLOT_BYTES);
                tma_load_2d(ring + pos * ST_SLOT_BYTES, &feat_map, (i_sub >> 1) * ST_KB, (int)row0, bar, policy);
            }
            if (++i_sub == 2 * ST_NKB) { i_sub = 0; i_tile += gridDim.x; }
            --i_left;
        };
        for (int s = 0; s < ring_slots; ++s)
            if (i_left > 0) issue(s);
        uint32_t roff[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = pw * 16 + i * 2 + rhalf;
            roff[i] = (uint32_t)(r * 128 + (((q >> 1) ^ (r & 7)) << 4) + ((q & 1) << 3));
        }
        const uint32_t a_base = smem_u32(asm_);
        const uint32_t lds_off = (uint32_t)(rhalf * (ST_KB * 4) + q * 16);
        int stage = 0, pos = 0;
        uint32_t parity = 0, rparity = 0;
        for (int64_t step = 0; step < my_tiles * ST_NKB; ++step) {
            uint2 c0[8], c1[8];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                mbar_wait_a(bars + pos * 8, rparity);
                const uint32_t sp = ring + pos * ST_SLOT_BYTES + lds_off;
#pragma unroll
                for (int i = 0; i < 4; ++i) split4(lds128(sp + i * 2 * (ST_KB * 4)), c0[4 * j + i], c1[4 * j + i]);
                __syncwarp();
                if (i_left > 0) issue(pos);
                if (++pos == ring_slots) { pos = 0; rparity ^= 1u; }
            }
            mbar_wait(&empty_bar[stage], parity ^ 1u);
            const uint32_t a0 = a_base + stage * ST_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sts64(a0 + roff[i], c0[i]);
                sts64(a0 + ST_A_BYTES + roff[i], c1[i]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == ST_A_STAGES) { stage = 0; parity ^= 1u; }
        }
    } else if (warp == SB_WARP_B) {
        // =============================== B loader: one K-block tile of the image per step ===========
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_last();
            int bs = 0;
            uint32_t bparity = 0;
            for (int64_t step = 0; step < my_tiles * ST_NKB; ++step) {
                const int kb = (int)(step % ST_NKB);
                mbar_wait(&bempty_bar[bs], bparity ^ 1u);
                mbar_arrive_expect_tx(&bfull_bar[bs], tile_bytes);
                for (uint32_t o = 0; o < tile_bytes; o += 16384u)      // 16 KB pieces, all counted on the one barrier
                    bulk_g2s(bsm + (size_t)bs * tile_bytes + o, image + (size_t)kb * tile_bytes + o,
                             tile_bytes - o < 16384u ? tile_bytes - o : 16384u, &bfull_bar[bs], policy);
                if (++bs == SB_B_STAGES) { bs = 0; bparity ^= 1u; }
            }
        }
    } else if (warp == ST_WARP_MMA) {
        // =============================== MMA issuer ================================================
        const uint32_t idesc = st_idesc_f16(g.npa);
        const uint32_t b1_off = (uint32_t)g.npa * 128u;
        int stage = 0, acc = 0, bs = 0;
        uint32_t parity = 0, acc_parity = 0, bparity = 0;
        for (int64_t t = 0; t < my_tiles; ++t) {
            if (lane == 0) {
                mbar_wait(&tempty_bar[acc], acc_parity ^ 1u);
                st_fence_after();
            }
            __syncwarp();
            const uint32_t tmem_d = tmem_base + acc * SB_ACC_COLS;
            for (int kb = 0; kb < ST_NKB; ++kb) {
                if (lane == 0) {
                    mbar_wait(&full_bar[stage], parity);
                    mbar_wait(&bfull_bar[bs], bparity);
                    st_fence_after();
                    const uint32_t a0 = smem_u32(asm_ + (size_t)stage * ST_STAGE_BYTES);
                    const uint32_t a1 = a0 + ST_A_BYTES;
                    const uint32_t bt = smem_u32(bsm + (size_t)bs * tile_bytes);
#pragma unroll
                    for (int ks = 0; ks < ST_KB / 16; ++ks) {
                        const uint32_t o = ks * 32;
                        const uint64_t da0 = st_desc_sw128(a0 + o), da1 = st_desc_sw128(a1 + o);
                        const uint64_t db0 = st_desc_sw128(bt + o), db1 = st_desc_sw128(bt + b1_off + o);
                        umma_f16(tmem_d, da0, db1, idesc, (kb | ks) != 0 ? 1u : 0u);
                        umma_f16(tmem_d, da1, db0, idesc, 1u);
                        umma_f16(tmem_d, da0, db0, idesc, 1u);
                    }
                    st_commit(&empty_bar[stage]);
                    st_commit(&bempty_bar[bs]);
                    if (kb == ST_NKB - 1) st_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == ST_A_STAGES) { stage = 0; parity ^= 1u; }
                if (++bs == SB_B_STAGES) { bs = 0; bparity ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
    } else {
        // =============================== epilogue (warps 0-3): thread = patch ========================
        const int C = n_classes;
        const int n_bank = col_off_s[C], n_cols = g.n_cols;
        const int n_chunks = (n_cols + 31) >> 5;
        const float descale = tail->descale;
        int acc = 0;
        uint32_t acc_parity = 0;
        bool bad = false;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t row = tile * ST_M + warp * 32 + lane;
            mbar_wait(&tfull_bar[acc], acc_parity);
            st_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * SB_ACC_COLS;
            float cs[MOC_BANK_MAX_CLASSES];
#pragma unroll
            for (int c = 0; c < MOC_BANK_MAX_CLASSES; ++c) cs[c] = 0.f;
            float bsum = 0.f, bmax = -INFINITY, probe = 0.f;
            int cls = 0;                                    // class of the current chunk's first column (warp-uniform)
            for (int ch = 0; ch < n_chunks; ++ch) {
                float d[32];
                st_tmem_ld32(taddr + ch * 32, d);
                const int c0 = ch * 32;
                while (cls < C && c0 >= col_off_s[cls + 1]) ++cls;
                if (cls < C && c0 + 32 <= col_off_s[cls + 1]) {
                    // the whole chunk belongs to class `cls`: fixed-shape tree sum, one add into that class
#pragma unroll
                    for (int w = 16; w > 0; w >>= 1)
#pragma unroll
                        for (int e = 0; e < w; ++e) d[e] += d[e + w];
                    probe = fmaf(d[0], 0.f, probe);
#pragma unroll
                    for (int c = 0; c < MOC_BANK_MAX_CLASSES; ++c)
                        if (c == cls) cs[c] += d[0];
                } else {
                    // a chunk straddling class boundaries, the background columns or the padding
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int col = c0 + e;
                        const float x = d[e];
                        if (col < n_bank) {
                            probe = fmaf(x, 0.f, probe);
#pragma unroll
                            for (int c = 0; c < MOC_BANK_MAX_CLASSES; ++c)
                                if (c < C && col >= col_off_s[c] && col < col_off_s[c + 1]) cs[c] += x;
                        } else if (col < n_cols) {
                            const float b = x * descale;
                            probe = fmaf(b, 0.f, probe);
                            bsum += b;
                            bmax = fmaxf(bmax, b);
                        }
                    }
                }
            }
            st_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
            if (row >= n_rows) continue;
            float inv = 1.0f;
            if (NORM) {
                const float4* xp = reinterpret_cast<const float4*>(feat + row * D);
                float ss = 0.f;
                for (int i = 0; i < D / 4; ++i) {
                    const float4 t = __ldg(xp + i);
                    ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
                }
                inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                bsum *= inv;
                bmax *= inv;
            }
            bad |= (probe != probe);
            float m1 = -INFINITY, m2 = -INFINITY;
            float* kp = keys + row;
#pragma unroll
            for (int c = 0; c < MOC_BANK_MAX_CLASSES; ++c) {
                if (c < C) {
                    const float l = cs[c] * descale * cls_scale_s[c] * inv;
                    cs[c] = l;
                    m2 = fmaxf(m2, fminf(m1, l));
                    m1 = fmaxf(m1, l);
                    kp[(int64_t)c * key_stride] = l;
                }
            }
            float esum = 0.f;
#pragma unroll
            for (int c = 0; c < MOC_BANK_MAX_CLASSES; ++c) {
                if (c < C) {
                    cs[c] = expf(cs[c] - m1);
                    esum += cs[c];
                }
            }
            const float inv_sum = 1.0f / esum;
#pragma unroll
            for (int c = 0; c < MOC_BANK_MAX_CLASSES; ++c)
                if (c < C) kp[(int64_t)(C + c) * key_stride] = cs[c] * inv_sum;
            kp[(int64_t)(2 * C) * key_stride] = fabsf(m1 - m2);
            kp[(int64_t)(2 * C + 1) * key_stride] = bsum;
            kp[(int64_t)(2 * C + 2) * key_stride] = bmax;
        }
        if (bad) atomicExch(&tail->flag, 1);
    }

    st_fence_before();
    __syncthreads();
    if (warp == ST_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols)
                     : "memory");
    }
}

// feat viewed as a 2-D fp32 tensor [n_rows][512]; box = one raw slot (ST_SLOT_ROWS patches x ST_KB floats).
static int make_feat_map(CUtensorMap* map, const float* feat, int64_t n_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        MOC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (fn == nullptr || q != cudaDriverEntryPointSuccess) {
            set_error("moc_score_keys_tc: the driver does not export cuTensorMapEncodeTiled");
            return MOC_E_CUDA;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ROW_BYTES};
    const cuuint32_t box[2] = {(cuuint32_t)ST_KB, (cuuint32_t)ST_SLOT_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("moc_score_keys_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return MOC_E_CUDA;
    }
    return MOC_OK;
}

template <int NCHUNK, bool NORM, int PW>
static int launch_tc_pw(const CUtensorMap& map, const float* feat, int64_t n_rows, const unsigned char* prep, int C, int n_cols,
                        const ScoreTcGeom& g, int ring_slots, size_t smem, float* keys, int64_t key_stride, cudaStream_t st) {
    MOC_CUDA(cudaFuncSetAttribute(score_keys_tc_kernel<NCHUNK, NORM, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int64_t n_tiles = (n_rows + ST_M - 1) / ST_M;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    ScoreTcTail* tail = reinterpret_cast<ScoreTcTail*>(const_cast<unsigned char*>(prep) + g.b_bytes());
    score_keys_tc_kernel<NCHUNK, NORM, PW><<<grid, StRoles<PW>::THREADS, smem, st>>>(map, feat, n_rows, prep, C, n_cols, g,
                                                                                    ring_slots, keys, key_stride, tail);
    MOC_LAUNCH_CHECK("score_keys_tc_kernel");
    return MOC_OK;
}

template <int NCHUNK, bool NORM>
static int launch_tc(const float* feat, int64_t n_rows, const unsigned char* prep, int C, int n_cols, float* keys,
                     int64_t key_stride, cudaStream_t st) {
    const ScoreTcGeom g = score_tc_geom(n_cols);
    const size_t b_bytes = g.b_bytes();
    const size_t budget = 227 * 1024 - 1024 - 2048;  // alignment slack, static shared memory
    const size_t fixed = b_bytes + (size_t)ST_A_STAGES * ST_STAGE_BYTES;
    auto slots_for = [&](int pw) {
        int n = fixed < budget ? (int)((budget - fixed) / ((size_t)pw * ST_SLOT_BYTES)) : 0;
        return n > ST_MAX_SLOTS ? ST_MAX_SLOTS : n;
    };
    static int force_pw = -1;       // MOC_SCORE_TC_PRODUCERS=8|16 (developer A/B switch)
    if (force_pw < 0) {
        const char* e = getenv("MOC_SCORE_TC_PRODUCERS");
        force_pw = e ? atoi(e) : 0;
    }
    // sixteen producer warps whenever their rings still hold two slots each
    const int pw = force_pw == 8 ? 8 : (slots_for(16) >= 2 ? 16 : 8);
    const int ring_slots = slots_for(pw);
    MOC_CHECK_SHAPE(ring_slots >= 2, "moc_score_keys_tc: %d prompt columns do not fit the tensor-core kernel", n_cols);
    const size_t smem = fixed + (size_t)ring_slots * pw * ST_SLOT_BYTES + 1024;
    CUtensorMap map;
    const int rc = make_feat_map(&map, feat, n_rows);
    if (rc != MOC_OK) return rc;
    return pw == 16 ? launch_tc_pw<NCHUNK, NORM, 16>(map, feat, n_rows, prep, C, n_cols, g, ring_slots, smem, keys, key_stride, st)
                    : launch_tc_pw<NCHUNK, NORM, 8>(map, feat, n_rows, prep, C, n_cols, g, ring_slots, smem, keys, key_stride, st);
}

template <bool NORM>
static int launch_bank(const float* feat, int64_t n_rows, const unsigned char* image, int C, const BankGeom& g, float* keys,
                       int64_t key_stride, cudaStream_t st) {
    const size_t budget = 227 * 1024 - 1024 - 2048;  // alignment slack, static shared memory
    const size_t fixed = (size_t)SB_B_STAGES * g.tile_bytes() + (size_t)ST_A_STAGES * ST_STAGE_BYTES;
    int ring_slots = fixed < budget ? (int)((budget - fixed) / ((size_t)ST_PROD_WARPS * ST_SLOT_BYTES)) : 0;
    if (ring_slots > ST_MAX_SLOTS) ring_slots = ST_MAX_SLOTS;
    MOC_CHECK_SHAPE(ring_slots >= 2, "moc_score_keys_bank_tc: %d columns do not fit the kernel's shared memory", g.n_cols);
    const size_t smem = fixed + (size_t)ring_slots * ST_PROD_WARPS * ST_SLOT_BYTES + 1024;
    MOC_CUDA(cudaFuncSetAttribute(score_bank_tc_kernel<NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tiles = (n_rows + ST_M - 1) / ST_M;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    BankTail* tail = reinterpret_cast<BankTail*>(const_cast<unsigned char*>(image) + g.b_bytes());
    CUtensorMap map;
    const int rc = make_feat_map(&map, feat, n_rows);
    if (rc != MOC_OK) return rc;
    score_bank_tc_kernel<NORM><<<grid, SB_THREADS, smem, st>>>(map, feat, n_rows, image, C, g, ring_slots, keys, key_stride,
                                                              tail);
    MOC_LAUNCH_CHECK("score_bank_tc_kernel");
    return MOC_OK;
}

static int bank_args_ok(int n_classes, int n_prompts, int n_bg) {
    MOC_CHECK_SHAPE(n_classes >= 2 && n_classes <= MOC_BANK_MAX_CLASSES,
                    "prompt bank: 2..%d classes supported, got %d", MOC_BANK_MAX_CLASSES, n_classes);
    MOC_CHECK_SHAPE(n_prompts >= n_classes && n_bg >= 1 && n_prompts + n_bg <= MOC_BANK_MAX_COLS,
                    "prompt bank: need >= 1 prompt per class, >= 1 background prompt and at most %d columns in all, got "
                    "%d + %d", MOC_BANK_MAX_COLS, n_prompts, n_bg);
    return MOC_OK;
}

}  // namespace moc

using namespace moc;

extern "C" size_t moc_prompt_bank_tc_bytes(int n_prompts, int n_bg) {
    if (n_prompts < 1 || n_bg < 1 || n_prompts + n_bg > MOC_BANK_MAX_COLS) return 0;
    return bank_geom(n_prompts, n_bg).image_bytes();
}

extern "C" size_t moc_prompt_bank_tc_flag_offset(int n_prompts, int n_bg) {
    return bank_geom(n_prompts, n_bg).b_bytes() + offsetof(BankTail, flag);
}

extern "C" int moc_prepare_prompt_bank_tc(const float* bank, const int32_t* class_offsets, int n_classes, int n_prompts,
                                          const float* bg, int n_bg, void* image, size_t image_bytes, void* stream) {
    MOC_CHECK_ARG(bank && class_offsets && bg && image, "moc_prepare_prompt_bank_tc: null pointer");
    {
        const int rc = bank_args_ok(n_classes, n_prompts, n_bg);
        if (rc != MOC_OK) return rc;
    }
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 15) == 0, "moc_prepare_prompt_bank_tc: image must be 16-byte aligned");
    const BankGeom g = bank_geom(n_prompts, n_bg);
    if (image_bytes < g.image_bytes()) {
        set_error("moc_prepare_prompt_bank_tc: image needs %zu bytes, got %zu", g.image_bytes(), image_bytes);
        return MOC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* out = reinterpret_cast<unsigned char*>(image);
    BankTail* tail = reinterpret_cast<BankTail*>(out + g.b_bytes());
    float* packed = reinterpret_cast<float*>(out + g.packed_offset());
    MOC_CUDA(cudaMemsetAsync(out, 0, g.packed_offset(), st));
    bank_pack_kernel<<<g.n_cols, D, 0, st>>>(bank, n_prompts, bg, n_bg, packed);
    MOC_LAUNCH_CHECK("bank_pack_kernel");
    score_tc_scale_kernel<<<1, 1024, 0, st>>>(packed, g.n_cols * D, reinterpret_cast<ScoreTcTail*>(tail));
    MOC_LAUNCH_CHECK("score_tc_scale_kernel");
    ScoreTcGeom tg;                       // the tile layout of score_tc_prep_kernel with b1 at row npa
    tg.npa = g.npa;
    tg.n_wide = 2 * g.npa;
    tg.n_narrow = g.npa;
    score_tc_prep_kernel<<<(g.n_cols * (D / 4) + 255) / 256, 256, 0, st>>>(packed, g.n_cols, tg,
                                                                         reinterpret_cast<const ScoreTcTail*>(tail), out);
    MOC_LAUNCH_CHECK("score_tc_prep_kernel");
    bank_class_scale_kernel<<<n_classes, D, 0, st>>>(packed, class_offsets, n_classes, tail);
    MOC_LAUNCH_CHECK("bank_class_scale_kernel");
    return MOC_OK;
}

extern "C" int moc_score_keys_bank_tc(const float* feat, int64_t n_rows, const void* image, int n_classes, int n_prompts,
                                      int n_bg, int normalize, float* keys, int64_t key_stride, void* stream) {
    MOC_CHECK_ARG(feat && image && keys, "moc_score_keys_bank_tc: null pointer");
    MOC_CHECK_ARG(n_rows >= 0 && key_stride >= n_rows, "moc_score_keys_bank_tc: bad n_rows / key_stride");
    {
        const int rc = bank_args_ok(n_classes, n_prompts, n_bg);
        if (rc != MOC_OK) return rc;
    }
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(feat) & 15) == 0, "moc_score_keys_bank_tc: feat must be 16-byte aligned");
    if (n_rows == 0) return MOC_OK;
    const BankGeom g = bank_geom(n_prompts, n_bg);
    const unsigned char* p = reinterpret_cast<const unsigned char*>(image);
    return normalize ? launch_bank<true>(feat, n_rows, p, n_classes, g, keys, key_stride, (cudaStream_t)stream)
                     : launch_bank<false>(feat, n_rows, p, n_classes, g, keys, key_stride, (cudaStream_t)stream);
}

extern "C" size_t moc_prompts_tc_bytes(int n_classes, int n_ext) {
    (void)n_classes;
    if (n_ext < 1 || n_ext > MOC_MAX_COLS) return 0;
    return score_tc_geom(n_ext).b_bytes() + sizeof(ScoreTcTail);
}

extern "C" size_t moc_prompts_tc_flag_offset(int n_classes, int n_ext) {
    (void)n_classes;
    return score_tc_geom(n_ext).b_bytes() + offsetof(ScoreTcTail, flag);
}

extern "C" int moc_prepare_prompts_tc(const float* packed, int n_classes, int n_ext, void* prompts_tc,
                                      size_t prompts_tc_bytes, void* stream) {
    MOC_CHECK_ARG(packed && prompts_tc, "moc_prepare_prompts_tc: null pointer");
    MOC_CHECK_SHAPE(n_classes >= 2 && n_ext > n_classes && n_ext <= MOC_MAX_COLS,
                    "moc_prepare_prompts_tc: need 2 <= C < C_ext <= %d, got C=%d C_ext=%d", MOC_MAX_COLS, n_classes, n_ext);
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(prompts_tc) & 15) == 0, "moc_prepare_prompts_tc: image must be 16-byte aligned");
    if (prompts_tc_bytes < moc_prompts_tc_bytes(n_classes, n_ext)) {
        set_error("moc_prepare_prompts_tc: image needs %zu bytes, got %zu", moc_prompts_tc_bytes(n_classes, n_ext),
                  prompts_tc_bytes);
        return MOC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const ScoreTcGeom g = score_tc_geom(n_ext);
    unsigned char* out = reinterpret_cast<unsigned char*>(prompts_tc);
    ScoreTcTail* tail = reinterpret_cast<ScoreTcTail*>(out + g.b_bytes());
    MOC_CUDA(cudaMemsetAsync(out, 0, g.b_bytes(), st));
    score_tc_scale_kernel<<<1, 1024, 0, st>>>(packed, n_ext * D, tail);
    MOC_LAUNCH_CHECK("score_tc_scale_kernel");
    score_tc_prep_kernel<<<(n_ext * (D / 4) + 255) / 256, 256, 0, st>>>(packed, n_ext, g, tail, out);
    MOC_LAUNCH_CHECK("score_tc_prep_kernel");
    return MOC_OK;
}

extern "C" int moc_score_keys_tc(const float* feat, int64_t n_rows, const void* prompts_tc, int n_classes, int n_ext,
                                 int normalize, float* keys, int64_t key_stride, void* stream) {
    MOC_CHECK_ARG(feat && prompts_tc && keys, "moc_score_keys_tc: null pointer");
    MOC_CHECK_ARG(n_rows >= 0 && key_stride >= n_rows, "moc_score_keys_tc: bad n_rows / key_stride");
    MOC_CHECK_SHAPE(n_classes >= 2 && n_ext > n_classes && n_ext <= MOC_MAX_COLS,
                    "moc_score_keys_tc: need 2 <= C < C_ext <= %d, got C=%d C_ext=%d", MOC_MAX_COLS, n_classes, n_ext);
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(feat) & 15) == 0, "moc_score_keys_tc: feat must be 16-byte aligned");
    if (n_rows == 0) return MOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(prompts_tc);
    if (n_ext <= 32)
        return normalize ? launch_tc<1, true>(feat, n_rows, p, n_classes, n_ext, keys, key_stride, st)
                         : launch_tc<1, false>(feat, n_rows, p, n_classes, n_ext, keys, key_stride, st);
    return normalize ? launch_tc<2, true>(feat, n_rows, p, n_classes, n_ext, keys, key_stride, st)
                     : launch_tc<2, false>(feat, n_rows, p, n_classes, n_ext, keys, key_stride, st);
}
