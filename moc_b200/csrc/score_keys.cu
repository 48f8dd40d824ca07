// Streaming patch-prompt scoring + selection keys.
//
// Replaces main_moc.py:336-337 (`feat @ zeroshot_weights`, `feat @ zeroshot_weights_ext`) and the
// per-row arithmetic of utils/patch_selection_classifier_index.py:34,46-48,71-75 and main_moc.py:359-366.
//
// Roofline: HBM.  One fp32 512-vector (2048 B) is read per patch, exactly once; the outputs are 2C+3 floats.
// Design (B200): persistent grid of one CTA per SM; every warp owns a private ring of STAGES x 8 KB shared
// memory stages that it fills itself with 1-D bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier),
// so there is no CTA-wide synchronisation anywhere in the steady state and ~128 KB are in flight per SM.
// A stage holds RP=4 consecutive patches.  The 32 lanes split the 512-long dot products (16 elements each,
// conflict-free LDS.128); with few prompt columns the lane's slice of every column lives in registers, so
// the inner loop is pure (two-wide) FMA.  Partial sums are combined with a two-level transposing butterfly
// (rows are scattered over lane groups while they are reduced) followed by a three-level all-reduce inside
// each group of 8 lanes.
// Measured on B200 (16.4 GB, C=2): 6.2-6.3 TB/s; the same ring with no compute and no stores reads 7.45 TB/s,
// dropping only the 28 B/patch key stores gives 6.65 TB/s (any DRAM write mixed into the read stream costs
// ~5 %, independent of store cache policy, locality or wave size), dropping 3/4 of the FMAs gives 6.55 TB/s.
#include <stdlib.h>
#include "common.cuh"

namespace moc {

static_assert(8 <= MOC_KEYS_COMPACT_MIN_CLASSES, "score_keys_regw_kernel (C + n_bg <= 8) writes the full 2C+3-plane key layout");
constexpr int RP = 4;                       // patches per stage
constexpr int STAGE_BYTES = RP * ROW_BYTES; // 8 KB
#ifndef MOC_SK_WARPS
#define MOC_SK_WARPS 8
#endif
#ifndef MOC_SK_STAGES
#define MOC_SK_STAGES 2
#endif
constexpr int SK_WARPS = MOC_SK_WARPS;
constexpr int SK_STAGES = MOC_SK_STAGES;

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}

// acc[r][c] partial sums of this lane for RP=4 rows.  On return u[c] holds the full 32-lane sum for row
// `row_of_lane()` in every lane.
template <int NV>
__device__ __forceinline__ void butterfly_rows(float (&acc)[RP][NV], float (&u)[NV], int lane) {
    const bool up16 = (lane & 16) != 0;
    const bool up8 = (lane & 8) != 0;
    float t[2][NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = up16 ? acc[i][c] : acc[i + 2][c];
            const float keep = up16 ? acc[i + 2][c] : acc[i][c];
            t[i][c] = keep + __shfl_xor_sync(FULL, send, 16);
        }
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        const float send = up8 ? t[0][c] : t[1][c];
        const float keep = up8 ? t[1][c] : t[0][c];
        float v = keep + __shfl_xor_sync(FULL, send, 8);
        v += __shfl_xor_sync(FULL, v, 4);
        v += __shfl_xor_sync(FULL, v, 2);
        v += __shfl_xor_sync(FULL, v, 1);
        u[c] = v;
    }
}
__device__ __forceinline__ int row_of_lane(int lane) { return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1); }

struct WarpRing {
    float* stage0;
    uint64_t* bars;
    uint64_t policy;
    __device__ __forceinline__ void issue(const float* feat, int64_t n_rows, int64_t group, int stage) const {
        const int64_t row0 = group * RP;
        const int64_t left = n_rows - row0;
        const uint32_t bytes = static_cast<uint32_t>((left < RP ? left : RP) * ROW_BYTES);
        mbar_arrive_expect_tx(&bars[stage], bytes);
        bulk_g2s(stage0 + stage * (RP * D), feat + row0 * D, bytes, &bars[stage], policy);
    }
};

// ------------------------------------------------------------------------------------------------
// Few prompt columns (NC = C + n_bg <= 8): prompt slices in registers.
//
// Work unit of a warp = a super-group of 32 consecutive patches (8 stages of 4).  Per stage the lane holds
// its 16-element slice of 4 patches; slot s of lane L holds patch (s ^ p(L)), p = 2*bit4(L) + bit3(L), so the
// two transposing butterfly levels need no selects: a lane always keeps slots {0,1} (then {0}) and sends
// {2,3} (then {1}).  FMAs are issued two-wide (FFMA2).  The reduced sums of a stage are parked in a per-warp
// smem scratch; after 8 stages every lane finishes one patch (softmax, top-2, background sum/max) and the key
// planes are written with full 128-byte stores.
// ------------------------------------------------------------------------------------------------
constexpr int SG_STAGES = 8;               // stages (of RP patches) per super-group
constexpr int SG_ROWS = SG_STAGES * RP;    // 32

__device__ __forceinline__ float2 ffma2(float a0, float a1, float b0, float b1, float2 c) {
    return __ffma2_rn(make_float2(a0, a1), make_float2(b0, b1), c);
}

template <int NC, bool NORM>
__global__ void __launch_bounds__(SK_WARPS * 32, 1)
score_keys_regw_kernel(const float* __restrict__ feat, int64_t n_rows, const float* __restrict__ packed,
                       int n_classes, float* __restrict__ keys, int64_t key_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NV = NC + (NORM ? 1 : 0);
    constexpr int SLD = (NV % 2 == 0) ? NV + 1 : NV;  // odd scratch stride: conflict-free row-per-lane reads

    WarpRing ring;
    ring.stage0 = reinterpret_cast<float*>(smem + (size_t)warp * SK_STAGES * STAGE_BYTES);
    unsigned char* tail = smem + (size_t)SK_WARPS * SK_STAGES * STAGE_BYTES;
    ring.bars = reinterpret_cast<uint64_t*>(tail) + warp * SK_STAGES;
    float* scratch = reinterpret_cast<float*>(tail + SK_WARPS * SK_STAGES * 8) + warp * (SG_ROWS * SLD);
    ring.policy = l2_policy_evict_first();

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SK_STAGES; ++s) mbar_init(&ring.bars[s], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncwarp();

    // groups (stages) are enumerated super-group by super-group: local index t -> group of 4 patches
    const int64_t n_groups = (n_rows + RP - 1) / RP;
    const int64_t n_sg = (n_rows + SG_ROWS - 1) / SG_ROWS;
    const int64_t sg_stride = (int64_t)gridDim.x * SK_WARPS;
    const int64_t sg0 = (int64_t)blockIdx.x * SK_WARPS + warp;
    auto group_of = [&](int64_t t) -> int64_t {  // t-th stage this warp processes
        const int64_t sg = sg0 + (t / SG_STAGES) * sg_stride;
        return sg < n_sg ? sg * SG_STAGES + (t % SG_STAGES) : n_groups;
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SK_STAGES; ++s) {
            const int64_t gg = group_of(s);
            if (gg < n_groups) ring.issue(feat, n_rows, gg, s);
        }
    }

    // this lane's slice of every prompt column: elements q*128 + lane*4 + {0..3}, q = 0..3
    float4 w[NC][4];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            w[c][q] = __ldg(reinterpret_cast<const float4*>(packed + c * D + q * 128 + lane * 4));

    const int C = n_classes;
    const int perm = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);  // slot s holds patch s ^ perm; reduced row = perm
    int stage = 0;
    uint32_t parity = 0;
    for (int64_t t = 0;; ++t) {
        const int64_t g = group_of(t);
        if (g >= n_groups) break;
        const int j = (int)(t % SG_STAGES);
        mbar_wait(&ring.bars[stage], parity);
        const float4* xs = reinterpret_cast<const float4*>(ring.stage0 + stage * (RP * D));
        float2 acc[RP][NV];
#pragma unroll
        for (int r = 0; r < RP; ++r)
#pragma unroll
            for (int c = 0; c < NV; ++c) acc[r][c] = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 xv[RP];
#pragma unroll
            for (int r = 0; r < RP; ++r) xv[r] = xs[(r ^ perm) * (D / 4) + q * 32 + lane];
#pragma unroll
            for (int r = 0; r < RP; ++r) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    acc[r][c] = ffma2(xv[r].x, xv[r].y, w[c][q].x, w[c][q].y, acc[r][c]);
                    acc[r][c] = ffma2(xv[r].z, xv[r].w, w[c][q].z, w[c][q].w, acc[r][c]);
                }
                if (NORM) {
                    acc[r][NV - 1] = ffma2(xv[r].x, xv[r].y, xv[r].x, xv[r].y, acc[r][NV - 1]);
                    acc[r][NV - 1] = ffma2(xv[r].z, xv[r].w, xv[r].z, xv[r].w, acc[r][NV - 1]);
                }
            }
        }
        __syncwarp();  // every lane has consumed the stage: hand it back to the copy engine
        if (lane == 0) {
            const int64_t gn = group_of(t + SK_STAGES);
            if (gn < n_groups) ring.issue(feat, n_rows, gn, stage);
        }
        // select-free transposing butterfly: keep slots {0,1}, send {2,3}; then keep {0}, send {1}
        float u[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const float a0 = acc[0][c].x + acc[0][c].y, a1 = acc[1][c].x + acc[1][c].y;
            const float a2 = acc[2][c].x + acc[2][c].y, a3 = acc[3][c].x + acc[3][c].y;
            const float t0 = a0 + __shfl_xor_sync(FULL, a2, 16);
            const float t1 = a1 + __shfl_xor_sync(FULL, a3, 16);
            float v = t0 + __shfl_xor_sync(FULL, t1, 8);
            v += __shfl_xor_sync(FULL, v, 4);
            v += __shfl_xor_sync(FULL, v, 2);
            v += __shfl_xor_sync(FULL, v, 1);
            u[c] = v;
        }
        if ((lane & 7) == 0) {
            float* dst = scratch + (j * RP + perm) * SLD;
#pragma unroll
            for (int c = 0; c < NV; ++c) dst[c] = u[c];
        }
        if (++stage == SK_STAGES) { stage = 0; parity ^= 1u; }

        const bool last_of_sg = (j == SG_STAGES - 1) || (g == n_groups - 1);
        if (!last_of_sg) continue;
        __syncwarp();
        // one patch per lane
        const int64_t row = (g / SG_STAGES) * SG_ROWS + lane;
        if (row < n_rows) {
            const float* sr = scratch + lane * SLD;
            float v[NV];
#pragma unroll
            for (int c = 0; c < NV; ++c) v[c] = sr[c];
            if (NORM) {
                const float inv = 1.0f / fmaxf(sqrtf(v[NV - 1]), 1e-12f);
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] *= inv;
            }
            float m1 = -INFINITY, m2 = -INFINITY, bsum = 0.f, bmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c < C) {
                    if (v[c] > m1) { m2 = m1; m1 = v[c]; } else if (v[c] > m2) { m2 = v[c]; }
                } else {
                    bsum += v[c];
                    bmax = fmaxf(bmax, v[c]);
                }
            }
            float e[NC];
            float esum = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                e[c] = (c < C) ? expf(v[c] - m1) : 0.f;
                esum += e[c];
            }
            const float inv_sum = 1.0f / esum;
            float* kp = keys + row;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c < C) {
                    kp[(int64_t)c * key_stride] = v[c];
                    kp[(int64_t)(C + c) * key_stride] = e[c] * inv_sum;
                }
            }
            kp[(int64_t)(2 * C) * key_stride] = fabsf(m1 - m2);
            kp[(int64_t)(2 * C + 1) * key_stride] = bsum;
            kp[(int64_t)(2 * C + 2) * key_stride] = bmax;
        }
        __syncwarp();  // scratch is reused by the next super-group
    }
}

// ------------------------------------------------------------------------------------------------
// Any column count up to MOC_MAX_COLS: prompts resident in shared memory, patches in registers.
// CUDA-core fallback for wide prompt sets (the dense-contraction case).
// ------------------------------------------------------------------------------------------------
template <bool NORM>
__global__ void __launch_bounds__(SK_WARPS * 32, 1)
score_keys_smemw_kernel(const float* __restrict__ feat, int64_t n_rows, const float* __restrict__ packed,
                        int n_classes, int n_cols, int n_cols_pad, int stages, float* __restrict__ keys,
                        int64_t key_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wsm = reinterpret_cast<float*>(smem);
    unsigned char* p = smem + (size_t)n_cols_pad * ROW_BYTES;
    WarpRing ring;
    ring.stage0 = reinterpret_cast<float*>(p + (size_t)warp * stages * STAGE_BYTES);
    p += (size_t)SK_WARPS * stages * STAGE_BYTES;
    float* scratch = reinterpret_cast<float*>(p) + warp * RP * n_cols_pad;
    p += (size_t)SK_WARPS * RP * n_cols_pad * 4;
    ring.bars = reinterpret_cast<uint64_t*>(p) + warp * stages;
    ring.policy = l2_policy_evict_first();

    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&ring.bars[s], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncwarp();
    const int64_t n_groups = (n_rows + RP - 1) / RP;
    const int64_t gstride = (int64_t)gridDim.x * SK_WARPS;
    int64_t g = (int64_t)blockIdx.x * SK_WARPS + warp;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) {
            const int64_t gg = g + s * gstride;
            if (gg < n_groups) ring.issue(feat, n_rows, gg, s);
        }
    }
    for (int i = threadIdx.x; i < n_cols_pad * (D / 4); i += blockDim.x)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(packed) + i);
    __syncthreads();

    const int C = n_classes;
    const int j8 = lane & 7;
    const int myrow = row_of_lane(lane);
    int stage = 0;
    uint32_t parity = 0;
    for (; g < n_groups; g += gstride) {
        mbar_wait(&ring.bars[stage], parity);
        const float4* xs = reinterpret_cast<const float4*>(ring.stage0 + stage * (RP * D));
        float4 xr[RP][4];
#pragma unroll
        for (int r = 0; r < RP; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) xr[r][q] = xs[r * (D / 4) + q * 32 + lane];
        __syncwarp();
        if (lane == 0) {
            const int64_t gn = g + (int64_t)stages * gstride;
            if (gn < n_groups) ring.issue(feat, n_rows, gn, stage);
        }
        float inv_norm = 1.f;
        if (NORM) {
            float ss[RP][1];
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                ss[r][0] = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) ss[r][0] = dot4(xr[r][q], xr[r][q], ss[r][0]);
            }
            float un[1];
            butterfly_rows<1>(ss, un, lane);
            inv_norm = 1.0f / fmaxf(sqrtf(un[0]), 1e-12f);
        }
        for (int c0 = 0; c0 < n_cols_pad; c0 += 4) {
            float acc[RP][4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float4* wp = reinterpret_cast<const float4*>(wsm + (c0 + cc) * D);
                const float4 w0 = wp[lane], w1 = wp[32 + lane], w2 = wp[64 + lane], w3 = wp[96 + lane];
#pragma unroll
                for (int r = 0; r < RP; ++r) {
                    float a = dot4(xr[r][0], w0, 0.f);
                    a = dot4(xr[r][1], w1, a);
                    a = dot4(xr[r][2], w2, a);
                    acc[r][cc] = dot4(xr[r][3], w3, a);
                }
            }
            float u[4];
            butterfly_rows<4>(acc, u, lane);
            if (j8 == 0)
                *reinterpret_cast<float4*>(scratch + myrow * n_cols_pad + c0) =
                    make_float4(u[0] * inv_norm, u[1] * inv_norm, u[2] * inv_norm, u[3] * inv_norm);
        }
        __syncwarp();
        // epilogue: the 8 lanes of a group share one row
        const int64_t row = g * RP + myrow;
        const float* sr = scratch + myrow * n_cols_pad;
        float m1 = -INFINITY, m2 = -INFINITY;
        for (int c = j8; c < C; c += 8) {
            const float v = sr[c];
            if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) { m2 = v; }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float om1 = __shfl_xor_sync(FULL, m1, o), om2 = __shfl_xor_sync(FULL, m2, o);
            const float hi = fmaxf(m1, om1), lo = fminf(m1, om1);
            m2 = fmaxf(lo, fmaxf(m2, om2));
            m1 = hi;
        }
        float esum = 0.f;
        for (int c = j8; c < C; c += 8) esum += expf(sr[c] - m1);
        float bsum = 0.f, bmax = -INFINITY;
        for (int c = C + j8; c < n_cols; c += 8) {
            bsum += sr[c];
            bmax = fmaxf(bmax, sr[c]);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            esum += __shfl_xor_sync(FULL, esum, o);
            bsum += __shfl_xor_sync(FULL, bsum, o);
            bmax = fmaxf(bmax, __shfl_xor_sync(FULL, bmax, o));
        }
        if (row < n_rows) {
            const float inv_sum = 1.0f / esum;
            const KeyLayout kl = key_layout(C);
            float* kp = keys + row;
            for (int c = j8; c < C; c += 8) {
                const float v = sr[c];
                kp[(int64_t)c * key_stride] = v;
                if (!kl.compact) kp[(int64_t)(C + c) * key_stride] = expf(v - m1) * inv_sum;
            }
            if (j8 == 0) {
                if (kl.compact) kp[(int64_t)kl.lse * key_stride] = lse_of(m1, esum);   // readers rebuild softmax: expf(L - lse)
                kp[(int64_t)kl.diff * key_stride] = fabsf(m1 - m2);
                kp[(int64_t)kl.bg_sum * key_stride] = bsum;
                kp[(int64_t)kl.bg_max * key_stride] = bmax;
            }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; parity ^= 1u; }
    }
}

// packed[col][512] from W [512,C] and the background columns of W_ext [512,n_ext]
__global__ void pack_prompts_kernel(const float* __restrict__ w, int C, const float* __restrict__ w_ext, int n_ext,
                                    int n_cols_pad, float* __restrict__ packed) {
    const int total = n_cols_pad * D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int col = i / D, k = i % D;
        float v = 0.f;
        if (col < C) v = w[k * C + col];
        else if (col < n_ext) v = w_ext[k * n_ext + col];
        packed[i] = v;
    }
}

// One block per class: column c of W = normalise( mean_p normalise(bank[p]) ) over the class's prompt rows
// (utils/zeroshot_utils.py:38-44).  Thread k owns embedding component k; norms are block reductions.
__global__ void __launch_bounds__(D) collapse_bank_kernel(const float* __restrict__ bank, const int32_t* __restrict__ class_offsets,
                                                         int n_classes, float* __restrict__ w_out) {
    __shared__ float red[D / 32];
    __shared__ float total;
    const int c = blockIdx.x, k = threadIdx.x, lane = k & 31, warp = k >> 5;
    auto block_norm = [&](float v) -> float {
        float s = warp_sum(v * v);
        __syncthreads();
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (k == 0) {
            float t = 0.f;
            for (int w = 0; w < D / 32; ++w) t += red[w];
            total = sqrtf(t);
        }
        __syncthreads();
        return total;
    };
    const int p0 = class_offsets[c], p1 = class_offsets[c + 1];
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) {
        const float v = bank[(size_t)p * D + k];
        acc += v / fmaxf(block_norm(v), 1e-12f);   // F.normalize(x, dim=-1): x / max(||x||, eps)
    }
    const float m = acc / (float)(p1 - p0);
    w_out[(size_t)k * n_classes + c] = m / block_norm(m);
}

// The pure streaming kernel reads HBM fastest with somewhat fewer CTAs than SMs: on the 148-SM B200, 108..132
// persistent CTAs read 6.74-6.81 TB/s, 134..148 only 6.33-6.42 TB/s (the step is sharp between 132 and 134).
static int64_t stream_cta_cap() {
    static int64_t cached = 0;
    if (cached == 0) {
        const int n = sm_count();
        cached = n == 148 ? 132 : n;
        if (const char* e = getenv("MOC_SCORE_CTAS")) {   // developer override for A/B runs
            const int v = atoi(e);
            if (v > 0 && v <= n) cached = v;
        }
    }
    return cached;
}

template <int NC, bool NORM>
static int launch_regw(const float* feat, int64_t n_rows, const float* packed, int C, float* keys,
                       int64_t key_stride, int max_ctas, cudaStream_t st) {
    constexpr int NV = NC + (NORM ? 1 : 0);
    constexpr int SLD = (NV % 2 == 0) ? NV + 1 : NV;
    constexpr size_t smem = (size_t)SK_WARPS * SK_STAGES * STAGE_BYTES + SK_WARPS * SK_STAGES * 8 +
                            (size_t)SK_WARPS * SG_ROWS * SLD * 4;
    MOC_CUDA(cudaFuncSetAttribute(score_keys_regw_kernel<NC, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int64_t n_sg = (n_rows + SG_ROWS - 1) / SG_ROWS;
    int64_t ctas = (n_sg + SK_WARPS - 1) / SK_WARPS;
    const int64_t cap = max_ctas > 0 ? (max_ctas < sm_count() ? max_ctas : sm_count()) : stream_cta_cap();
    if (ctas > cap) ctas = cap;
    score_keys_regw_kernel<NC, NORM><<<(unsigned)ctas, SK_WARPS * 32, smem, st>>>(feat, n_rows, packed, C, keys,
                                                                                key_stride);
    MOC_LAUNCH_CHECK("score_keys_regw_kernel");
    return MOC_OK;
}

template <bool NORM>
static int launch_smemw(const float* feat, int64_t n_rows, const float* packed, int C, int n_cols, float* keys,
                        int64_t key_stride, cudaStream_t st) {
    const int n_cols_pad = (n_cols + 3) & ~3;
    const size_t fixed = (size_t)n_cols_pad * ROW_BYTES + (size_t)SK_WARPS * RP * n_cols_pad * 4;
    const size_t budget = 227 * 1024;
    int stages = 4;
    while (stages > 1 && fixed + (size_t)SK_WARPS * stages * (STAGE_BYTES + 8) > budget) --stages;
    MOC_CHECK_SHAPE(fixed + (size_t)SK_WARPS * stages * (STAGE_BYTES + 8) <= budget,
                    "moc_score_keys: %d prompt columns do not fit in shared memory", n_cols);
    const size_t smem = fixed + (size_t)SK_WARPS * stages * (STAGE_BYTES + 8);
    static size_t configured = 0;
    if (configured < smem) {
        MOC_CUDA(cudaFuncSetAttribute(score_keys_smemw_kernel<NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)budget));
        configured = budget;
    }
    const int64_t n_groups = (n_rows + RP - 1) / RP;
    int64_t ctas = (n_groups + SK_WARPS - 1) / SK_WARPS;
    if (ctas > sm_count()) ctas = sm_count();
    score_keys_smemw_kernel<NORM><<<(unsigned)ctas, SK_WARPS * 32, smem, st>>>(feat, n_rows, packed, C, n_cols,
                                                                              n_cols_pad, stages, keys, key_stride);
    MOC_LAUNCH_CHECK("score_keys_smemw_kernel");
    return MOC_OK;
}

}  // namespace moc

using namespace moc;

extern "C" int moc_packed_cols(int n_classes, int n_ext) {
    (void)n_classes;
    return (n_ext + 3) & ~3;
}

extern "C" int moc_pack_prompts(const float* w, int n_classes, const float* w_ext, int n_ext, float* packed,
                                void* stream) {
    MOC_CHECK_ARG(w && w_ext && packed, "moc_pack_prompts: null pointer");
    MOC_CHECK_SHAPE(n_classes >= 2 && n_ext > n_classes && n_ext <= MOC_MAX_COLS,
                    "moc_pack_prompts: need 2 <= C < C_ext <= %d, got C=%d C_ext=%d", MOC_MAX_COLS, n_classes, n_ext);
    const int pad = moc_packed_cols(n_classes, n_ext);
    pack_prompts_kernel<<<(pad * D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_classes, w_ext, n_ext, pad,
                                                                                packed);
    MOC_LAUNCH_CHECK("pack_prompts_kernel");
    return MOC_OK;
}

extern "C" int moc_collapse_prompt_bank(const float* bank, const int32_t* class_offsets, int n_classes, float* w_out,
                                        void* stream) {
    MOC_CHECK_ARG(bank && class_offsets && w_out, "moc_collapse_prompt_bank: null pointer");
    MOC_CHECK_SHAPE(n_classes >= 1 && n_classes <= MOC_MAX_COLS, "moc_collapse_prompt_bank: bad class count %d", n_classes);
    collapse_bank_kernel<<<n_classes, D, 0, (cudaStream_t)stream>>>(bank, class_offsets, n_classes, w_out);
    MOC_LAUNCH_CHECK("collapse_bank_kernel");
    return MOC_OK;
}

extern "C" int moc_score_keys(const float* feat, int64_t n_rows, const float* packed, int n_classes, int n_ext,
                              int normalize, float* keys, int64_t key_stride, void* stream) {
    return moc_score_keys_ex(feat, n_rows, packed, n_classes, n_ext, normalize, keys, key_stride, 0, stream);
}

extern "C" int moc_score_keys_ex(const float* feat, int64_t n_rows, const float* packed, int n_classes, int n_ext,
                                 int normalize, float* keys, int64_t key_stride, int max_ctas, void* stream) {
    MOC_CHECK_ARG(feat && packed && keys, "moc_score_keys: null pointer");
    MOC_CHECK_ARG(n_rows >= 0 && key_stride >= n_rows, "moc_score_keys: bad n_rows / key_stride");
    MOC_CHECK_SHAPE(n_classes >= 2 && n_ext > n_classes && n_ext <= MOC_MAX_COLS,
                    "moc_score_keys: need 2 <= C < C_ext <= %d, got C=%d C_ext=%d", MOC_MAX_COLS, n_classes, n_ext);
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(feat) & 15) == 0, "moc_score_keys: feat must be 16-byte aligned");
    if (n_rows == 0) return MOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define MOC_REGW(NCV)                                                                                  \
    case NCV:                                                                                          \
        return normalize ? launch_regw<NCV, true>(feat, n_rows, packed, n_classes, keys, key_stride, max_ctas, st) \
                         : launch_regw<NCV, false>(feat, n_rows, packed, n_classes, keys, key_stride, max_ctas, st);
    switch (n_ext) {
        MOC_REGW(3)
        MOC_REGW(4)
        MOC_REGW(5)
        MOC_REGW(6)
        MOC_REGW(7)
        MOC_REGW(8)
        default:
            break;
    }
#undef MOC_REGW
    return normalize ? launch_smemw<true>(feat, n_rows, packed, n_classes, n_ext, keys, key_stride, st)
                     : launch_smemw<false>(feat, n_rows, packed, n_classes, n_ext, keys, key_stride, st);
}
