// Weight gradient of a dense per-patch layer on the 5th-generation tensor cores:  dW = G^T X,  fp32 in, fp32 out.
//
// For y = x W^T (+ b) applied to every patch of a bag, autograd's weight gradient is dW[m][k] = sum_n G[n][m] X[n][k]
// with G = d(loss)/dy  -  the contraction runs over the N patches of the bag (20 000 .. 100 000), the output is only
// [n_out][k] (e.g. 768 x 512).  This is the backward of the ABMIL layers models/model_clam.py:83-91 (fc 512->512) and
// :44-49 (gated-attention branches 512->384 x2), which the reference trains through torch autograd
// (utils/core_utils.py:391-416: loss.backward(); optimizer.step()).
//
// Both operands are stored with the contraction index n as the slow dimension, the opposite of what the K-major
// tensor-core tiles want, so the producers transpose on the way through registers: a thread loads a 4(n) x 4(m) block
// with four 16-byte loads, splits every value into hi/lo TF32 parts and stores four 16-byte pieces (4 consecutive n
// of one m) into the 128B-swizzled K-major tile; the 8 lanes of a quarter-warp hold the 8 different n-groups of one
// row, so the stores are bank-conflict free and the loads still cover whole 64-byte segments.
//
// Work item = (128 x 128 output tile, slice of the bag).  The bag is cut into as many slices as it takes to give
// every SM a work item (split-K); each item writes its partial tile to the workspace and wgrad_reduce_kernel adds the
// slices in a fixed order (deterministic - no atomics).
//   warps 0-3   epilogue: tcgen05.ld the 128x128 accumulator, 16-byte stores of the partial tile
//   warps 4-11  producers of both operands (2 K-blocks of loads in flight per thread)
//   warp 12     MMA issuer (one elected lane): per K-block 4 k-steps x 3 products of tcgen05.mma.kind::tf32
// Precision as linear_tc.cu: 3xTF32, fp32 accumulation in TMEM.
#include "tc_common.cuh"

namespace moc {

constexpr int WG_THREADS = (LT_EPI_WARPS + LT_PROD_WARPS + 1) * 32;  // 416
constexpr int WG_WARP_MMA = LT_EPI_WARPS + LT_PROD_WARPS;           // 12
constexpr int WG_MAX_KB = 64;   // K-blocks (of 32 patches) accumulated in TMEM per work item
constexpr size_t WG_SMEM = (size_t)LT_STAGES * LT_STAGE_BYTES + 1024;

struct WgBlock {
    float4 g[4], x[4];  // rows n0+4*ng+j (j = 0..3): 4 consecutive columns of G and of X
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const float* __restrict__ g, int64_t ldg, int M, const float* __restrict__ x, int64_t ldx, int K,
                int64_t n_rows, int n_mt, int n_kt, int n_slices, int64_t kb_per_slice, float* __restrict__ part) {
    extern __shared__ unsigned char wg_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[LT_STAGES], empty_bar[LT_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(wg_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < LT_STAGES; ++s) {
            mbar_init(&full_bar[s], LT_PROD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], LT_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == WG_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)LT_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    lt_fence_before();
    __syncthreads();
    lt_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int64_t total_kb = (n_rows + LT_KB - 1) / LT_KB;
    const int64_t n_work = (int64_t)n_mt * n_kt * n_slices;   // w = (slice, m tile, k tile), k tile fastest
    const uint32_t smem_base = smem_u32(smem);
    const int64_t Mp = (int64_t)n_mt * LT_M, Kp = (int64_t)n_kt * LT_N;

    if (warp >= LT_EPI_WARPS && warp < WG_WARP_MMA) {
        // =============================== producers =================================================
        const int pw = warp - LT_EPI_WARPS, ng = lane & 7, cg = pw * 4 + (lane >> 3);  // n-group 0..7, column group 0..31
        uint32_t soff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = cg * 4 + i;   // tile row = output row (m) / output column (k) index inside the tile
            soff[i] = (uint32_t)(r * 128 + ((ng ^ (r & 7)) << 4));
        }
        int stage = 0;
        uint32_t parity = 0;
        for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int kt = (int)(wk % n_kt), mt = (int)((wk / n_kt) % n_mt);
            const int64_t sl = wk / ((int64_t)n_kt * n_mt);
            const int64_t kb0 = sl * kb_per_slice, kb1 = kb0 + kb_per_slice < total_kb ? kb0 + kb_per_slice : total_kb;
            const int mcol = mt * LT_M + cg * 4, kcol = kt * LT_N + cg * 4;
            const bool m_ok = mcol < M, k_ok = kcol < K;   // M, K are multiples of 4
            auto load = [&](WgBlock& b, int64_t kb) {
                const int64_t r0 = kb * LT_KB + ng * 4;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool ok = r0 + j < n_rows;
                    b.g[j] = ok && m_ok ? __ldg(reinterpret_cast<const float4*>(g + (r0 + j) * ldg + mcol)) : z;
                    b.x[j] = ok && k_ok ? __ldg(reinterpret_cast<const float4*>(x + (r0 + j) * ldx + kcol)) : z;
                }
            };
            auto store = [&](const WgBlock& b) {
                mbar_wait(&empty_bar[stage], parity ^ 1u);
                const uint32_t a_hi = smem_base + stage * LT_STAGE_BYTES, a_lo = a_hi + LT_A_BYTES;
                const uint32_t b_hi = a_hi + 2 * LT_A_BYTES, b_lo = b_hi + LT_B_BYTES;
                float4 hi, lo;
                // transposed 4x4 blocks: piece i = column i of the four loaded rows
                lt_split(make_float4(b.g[0].x, b.g[1].x, b.g[2].x, b.g[3].x), hi, lo); sts128(a_hi + soff[0], hi); sts128(a_lo + soff[0], lo);
                lt_split(make_float4(b.g[0].y, b.g[1].y, b.g[2].y, b.g[3].y), hi, lo); sts128(a_hi + soff[1], hi); sts128(a_lo + soff[1], lo);
                lt_split(make_float4(b.g[0].z, b.g[1].z, b.g[2].z, b.g[3].z), hi, lo); sts128(a_hi + soff[2], hi); sts128(a_lo + soff[2], lo);
                lt_split(make_float4(b.g[0].w, b.g[1].w, b.g[2].w, b.g[3].w), hi, lo); sts128(a_hi + soff[3], hi); sts128(a_lo + soff[3], lo);
                lt_split(make_float4(b.x[0].x, b.x[1].x, b.x[2].x, b.x[3].x), hi, lo); sts128(b_hi + soff[0], hi); sts128(b_lo + soff[0], lo);
                lt_split(make_float4(b.x[0].y, b.x[1].y, b.x[2].y, b.x[3].y), hi, lo); sts128(b_hi + soff[1], hi); sts128(b_lo + soff[1], lo);
                lt_split(make_float4(b.x[0].z, b.x[1].z, b.x[2].z, b.x[3].z), hi, lo); sts128(b_hi + soff[2], hi); sts128(b_lo + soff[2], lo);
                lt_split(make_float4(b.x[0].w, b.x[1].w, b.x[2].w, b.x[3].w), hi, lo); sts128(b_hi + soff[3], hi); sts128(b_lo + soff[3], lo);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
                if (++stage == LT_STAGES) { stage = 0; parity ^= 1u; }
            };
            WgBlock b0, b1;
            load(b0, kb0);
            if (kb0 + 1 < kb1) load(b1, kb0 + 1);
            for (int64_t kb = kb0; kb < kb1; kb += 2) {
                store(b0);
                if (kb + 2 < kb1) load(b0, kb + 2);
                if (kb + 1 < kb1) {
                    store(b1);
                    if (kb + 3 < kb1) load(b1, kb + 3);
                }
            }
        }
    } else if (warp == WG_WARP_MMA) {
        // =============================== MMA issuer ================================================
        int stage = 0, acc = 0;
        uint32_t parity = 0, acc_parity = 0;
        for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int64_t sl = wk / ((int64_t)n_kt * n_mt);
            const int64_t kb0 = sl * kb_per_slice, kb1 = kb0 + kb_per_slice < total_kb ? kb0 + kb_per_slice : total_kb;
            if (lane == 0) {
                mbar_wait(&tempty_bar[acc], acc_parity ^ 1u);
                lt_fence_after();
            }
            __syncwarp();
            const uint32_t tmem_d = tmem_base + acc * LT_N;
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                if (lane == 0) {
                    mbar_wait(&full_bar[stage], parity);
                    lt_fence_after();
                    const uint32_t a_hi = smem_base + stage * LT_STAGE_BYTES;
                    const uint32_t a_lo = a_hi + LT_A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * LT_A_BYTES;
                    const uint32_t b_lo = b_hi + LT_B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < LT_KB / 8; ++ks) {
                        const uint32_t o = ks * 32;
                        const uint64_t dah = lt_desc_sw128(a_hi + o), dal = lt_desc_sw128(a_lo + o);
                        const uint64_t dbh = lt_desc_sw128(b_hi + o), dbl = lt_desc_sw128(b_lo + o);
                        lt_umma_tf32(tmem_d, dal, dbh, (kb != kb0 || ks != 0) ? 1u : 0u);
                        lt_umma_tf32(tmem_d, dah, dbl, 1u);
                        lt_umma_tf32(tmem_d, dah, dbh, 1u);
                    }
                    lt_commit(&empty_bar[stage]);
                    if (kb == kb1 - 1) lt_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == LT_STAGES) { stage = 0; parity ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
    } else {
        // =============================== epilogue (warps 0-3): thread = output row m =================
        int acc = 0;
        uint32_t acc_parity = 0;
        for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int kt = (int)(wk % n_kt), mt = (int)((wk / n_kt) % n_mt);
            const int64_t sl = wk / ((int64_t)n_kt * n_mt);
            mbar_wait(&tfull_bar[acc], acc_parity);
            lt_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * LT_N;
            float* out = part + (sl * Mp + (int64_t)mt * LT_M + warp * 32 + lane) * Kp + (int64_t)kt * LT_N;
#pragma unroll 1
            for (int q = 0; q < LT_N / 32; ++q) {
                float d[32];
                lt_tmem_ld32(taddr + q * 32, d);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(out + q * 32 + j) = make_float4(d[j], d[j + 1], d[j + 2], d[j + 3]);
            }
            lt_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
    }

    lt_fence_before();
    __syncthreads();
    if (warp == WG_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)LT_TMEM_COLS)
                     : "memory");
    }
}

// dW[m][k] (+)= sum over slices, in slice order
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int n_slices, int64_t Mp, int64_t Kp, int M, int K,
                                    float* __restrict__ dw, int64_t lddw, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * (K / 4)) return;
    const int m = (int)(i / (K / 4)), k4 = (int)(i % (K / 4));
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sl = 0; sl < n_slices; ++sl) {
        const float4 v = *reinterpret_cast<const float4*>(part + ((int64_t)sl * Mp + m) * Kp + 4 * k4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = dw + (int64_t)m * lddw + 4 * k4;
    if (accumulate) { s.x += o[0]; s.y += o[1]; s.z += o[2]; s.w += o[3]; }
    o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w;
}

struct WgPlan {
    int n_mt, n_kt, n_slices;
    int64_t kb_per_slice;
};
static WgPlan wg_plan(int64_t n_rows, int M, int K) {
    WgPlan p;
    p.n_mt = (M + LT_M - 1) / LT_M;
    p.n_kt = (K + LT_N - 1) / LT_N;
    const int64_t total_kb = (n_rows + LT_KB - 1) / LT_KB;
    const int tiles = p.n_mt * p.n_kt;
    // Slices: whole waves of work items over the SMs (k * SMs / tiles of them), and short enough (<= WG_MAX_KB
    // K-blocks) that the tensor core's truncating fp32 accumulation stays ~1e-5 relative; the slices are then added
    // with round-to-nearest in wgrad_reduce_kernel.
    int64_t slices = 1;
    for (int k = 1;; ++k) {
        slices = (int64_t)k * sm_count() / tiles;
        if (slices < 1) slices = 1;
        if (slices >= total_kb) { slices = total_kb > 0 ? total_kb : 1; break; }
        if ((total_kb + slices - 1) / slices <= WG_MAX_KB) break;
    }
    p.kb_per_slice = (total_kb + slices - 1) / slices;
    if (p.kb_per_slice < 1) p.kb_per_slice = 1;
    p.n_slices = (int)((total_kb + p.kb_per_slice - 1) / p.kb_per_slice);   // no empty slice
    if (p.n_slices < 1) p.n_slices = 1;
    return p;
}

}  // namespace moc

using namespace moc;

extern "C" size_t moc_linear_wgrad_workspace_bytes(int64_t n_rows, int n_out, int k) {
    if (n_rows < 0 || n_out < 1 || k < 1) return 0;
    const WgPlan p = wg_plan(n_rows, n_out, k);
    return (size_t)p.n_slices * p.n_mt * LT_M * p.n_kt * LT_N * sizeof(float);
}

extern "C" int moc_linear_wgrad(const float* g, int64_t ldg, int n_out, const float* x, int64_t ldx, int k, int64_t n_rows,
                                float* dw, int64_t lddw, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    MOC_CHECK_ARG(g && x && dw && workspace, "moc_linear_wgrad: null pointer");
    MOC_CHECK_ARG(n_rows >= 0 && ldg >= n_out && ldx >= k && lddw >= k, "moc_linear_wgrad: bad n_rows / leading dimensions");
    MOC_CHECK_SHAPE(n_out >= 4 && n_out % 4 == 0 && n_out <= 8192 && k >= 4 && k % 4 == 0 && k <= 8192,
                    "moc_linear_wgrad: out_features (%d) and in_features (%d) must be multiples of 4", n_out, k);
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (ldg & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "moc_linear_wgrad: g, x and the workspace must be 16-byte aligned, ldg / ldx multiples of 4");
    const size_t need = moc_linear_wgrad_workspace_bytes(n_rows, n_out, k);
    if (workspace_bytes < need) {
        set_error("moc_linear_wgrad: workspace %zu B < required %zu B", workspace_bytes, need);
        return MOC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) {
        if (!accumulate) MOC_CUDA(cudaMemset2DAsync(dw, (size_t)lddw * sizeof(float), 0, (size_t)k * sizeof(float), n_out, st));
        return MOC_OK;
    }
    const WgPlan p = wg_plan(n_rows, n_out, k);
    MOC_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    const int64_t n_work = (int64_t)p.n_mt * p.n_kt * p.n_slices;
    const int grid = (int)(n_work < sm_count() ? n_work : sm_count());
    float* part = reinterpret_cast<float*>(workspace);
    wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM, st>>>(g, ldg, n_out, x, ldx, k, n_rows, p.n_mt, p.n_kt, p.n_slices,
                                                      p.kb_per_slice, part);
    MOC_LAUNCH_CHECK("wgrad_tc_kernel");
    const int64_t items = (int64_t)n_out * (k / 4);
    wgrad_reduce_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(part, p.n_slices, (int64_t)p.n_mt * LT_M,
                                                                         (int64_t)p.n_kt * LT_N, n_out, k, dw, lddw, accumulate);
    MOC_LAUNCH_CHECK("wgrad_reduce_kernel");
    return MOC_OK;
}
