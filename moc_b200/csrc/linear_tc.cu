// Dense per-patch layer on the 5th-generation tensor cores:  Y = act(X W^T + b),  fp32 in, fp32 out.
//
// Building block of the secondary MIL heads that north_star names next to MOC proper - every nn.Linear applied
// to all N patches of a bag:
//   Conch_CLIP_Ada.adapter   512 -> 128 -> 512          (models/model_adapters.py:152-157, :186)
//   CLAM_SB / ABMIL          fc 512 -> 512, gated attention 512 -> 384 (x2)   (models/model_clam.py:83-91, :41-64)
//   MIL_fc                   384 -> 512 -> 2            (models/model_mil.py:17-23)
// These are dense contractions (0.26 - 1.3 MFLOP per 2 KB patch), i.e. tensor-core work, unlike the streaming
// scoring kernel.
//
// Precision: 3xTF32 (x = hi + lo, hi = x rounded to TF32; D = A_lo B_hi + A_hi B_lo + A_hi B_hi with fp32
// accumulation in TMEM).  The split is exact to 2^-21; measured end-to-end error against float64 is a few 1e-6
// relative (the tensor core's own accumulation), two orders inside the 1e-3 parity bar even through the three or
// four stacked layers of these heads.
//
// Structure (persistent, one CTA per SM, 14 warps; a work item is a 128-row x 128-column block of Y):
//   warps 0-3   epilogue: tcgen05.ld the 128x128 accumulator (thread = row), + bias, activation, 16-byte stores
//   warps 4-11  A producers: 128 B of 16 rows each per K-block (LT_PF K-blocks of loads in flight per thread,
//               across work items), hi/lo split in registers, stores into the 128B-swizzled K-major tiles
//   warp 12     MMA issuer (one elected lane): 4 k-steps x 3 products of tcgen05.mma.kind::tf32 M128 N128 K8
//   warp 13     B copier: one 32 KB bulk copy per K-block of the pre-split, pre-swizzled weight block (hi|lo)
// Three smem stages of {A_hi 16K, A_lo 16K, B_hi 16K, B_lo 16K}; two TMEM accumulators.
#include "tc_common.cuh"

namespace moc {

// W [n_out][K] row-major (nn.Linear.weight) -> for every (column block cb, K-block kb) one tile
// [hi 16 KB | lo 16 KB] in the swizzled K-major layout; rows past n_out are zero.
__global__ void linear_tc_prep_kernel(const float* __restrict__ w, int n_out, int K, int n_cb, float* __restrict__ w_split) {
    const int n_kb = K / LT_KB;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk of the padded W
    const int64_t total = (int64_t)n_cb * LT_N * (K / 4);
    if (i >= total) return;
    const int n = (int)(i / (K / 4)), c4 = (int)(i % (K / 4));
    const int kb = c4 / 8, chunk = c4 % 8;
    const int cb = n / LT_N, nl = n % LT_N;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < n_out) v = reinterpret_cast<const float4*>(w + (size_t)n * K)[c4];
    float4 hi, lo;
    lt_split(v, hi, lo);
    char* tile = reinterpret_cast<char*>(w_split) + ((size_t)cb * n_kb + kb) * 2 * LT_B_BYTES;
    const int off = nl * 128 + ((chunk ^ (nl & 7)) << 4);
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + LT_B_BYTES + off) = lo;
}

__global__ void __launch_bounds__(LT_THREADS, 1)
linear_tc_kernel(const float* __restrict__ x, int64_t ldx, int64_t n_rows, int K, const float* __restrict__ w_split,
                 const float* __restrict__ bias, int n_out, int n_cb, int act0, int split, int act1,
                 float* __restrict__ y, int64_t ldy) {
    extern __shared__ unsigned char lt_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[LT_STAGES], empty_bar[LT_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(lt_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_kb = K / LT_KB;

    if (tid == 0) {
        for (int s = 0; s < LT_STAGES; ++s) {
            mbar_init(&full_bar[s], LT_PROD_WARPS + 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], LT_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == LT_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)LT_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    lt_fence_before();
    __syncthreads();
    lt_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int64_t n_tiles = (n_rows + LT_M - 1) / LT_M;
    const int64_t n_work = n_tiles * n_cb;   // work item w = (row tile w / n_cb, column block w % n_cb)
    const uint32_t smem_base = smem_u32(smem);

    if (warp >= LT_EPI_WARPS && warp < LT_WARP_MMA) {
        // =============================== A producers ===============================================
        const int pw = warp - LT_EPI_WARPS, rsub = lane >> 3, chunk = lane & 7;
        uint32_t roff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = pw * 16 + i * 4 + rsub;
            roff[i] = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
        }
        const int64_t my_work = (int64_t)blockIdx.x < n_work ? (n_work - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t n_steps = my_work * n_kb;
        // load cursor over the flat (work item, K-block) stream
        int64_t l_work = blockIdx.x, l_left = n_steps;
        int l_kb = 0;
        auto issue = [&](float4 (&b)[4]) {
            const int64_t row0 = (l_work / n_cb) * LT_M + pw * 16 + rsub;
            const float4* p = reinterpret_cast<const float4*>(x + row0 * ldx + l_kb * LT_KB) + chunk;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                b[i] = row0 + i * 4 < n_rows ? __ldg(p + i * ldx) : make_float4(0.f, 0.f, 0.f, 0.f);  // 4 rows = 4*ldx floats = ldx float4
            if (++l_kb == n_kb) { l_kb = 0; l_work += gridDim.x; }
            --l_left;
        };
        float4 buf[LT_PF][4];
#pragma unroll
        for (int s = 0; s < LT_PF; ++s)
            if (l_left > 0) issue(buf[s]);
        int stage = 0;
        uint32_t parity = 0;
        for (int64_t step0 = 0; step0 < n_steps; step0 += LT_PF) {
#pragma unroll
            for (int s = 0; s < LT_PF; ++s) {
                if (step0 + s < n_steps) {
                    mbar_wait(&empty_bar[stage], parity ^ 1u);
                    const uint32_t a_hi = smem_base + stage * LT_STAGE_BYTES, a_lo = a_hi + LT_A_BYTES;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 cur = buf[s][i];
                        float4 hi, lo;
                        lt_split(cur, hi, lo);
                        sts128(a_hi + roff[i], hi);
                        sts128(a_lo + roff[i], lo);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                    if (++stage == LT_STAGES) { stage = 0; parity ^= 1u; }
                    if (l_left > 0) issue(buf[s]);
                }
            }
        }
    } else if (warp == LT_WARP_B) {
        // =============================== B copier ==================================================
        const uint64_t policy = l2_policy_evict_last();
        int stage = 0;
        uint32_t parity = 0;
        for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int cb = (int)(wk % n_cb);
            for (int kb = 0; kb < n_kb; ++kb) {
                if (lane == 0) {
                    mbar_wait(&empty_bar[stage], parity ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[stage], 2 * LT_B_BYTES);
                    bulk_g2s(smem + (size_t)stage * LT_STAGE_BYTES + 2 * LT_A_BYTES,
                             reinterpret_cast<const char*>(w_split) + ((size_t)cb * n_kb + kb) * 2 * LT_B_BYTES,
                             2 * LT_B_BYTES, &full_bar[stage], policy);
                }
                __syncwarp();
                if (++stage == LT_STAGES) { stage = 0; parity ^= 1u; }
            }
        }
    } else if (warp == LT_WARP_MMA) {
        // =============================== MMA issuer ================================================
        int stage = 0, acc = 0;
        uint32_t parity = 0, acc_parity = 0;
        for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            if (lane == 0) {
                mbar_wait(&tempty_bar[acc], acc_parity ^ 1u);
                lt_fence_after();
            }
            __syncwarp();
            const uint32_t tmem_d = tmem_base + acc * LT_N;
            for (int kb = 0; kb < n_kb; ++kb) {
                if (lane == 0) {
                    mbar_wait(&full_bar[stage], parity);
                    lt_fence_after();
                    const uint32_t a_hi = smem_base + stage * LT_STAGE_BYTES;
                    const uint32_t a_lo = a_hi + LT_A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * LT_A_BYTES;
                    const uint32_t b_lo = b_hi + LT_B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < LT_KB / 8; ++ks) {
                        const uint32_t o = ks * 32;  // 8 tf32 = 32 bytes along K inside the swizzled row
                        const uint64_t dah = lt_desc_sw128(a_hi + o), dal = lt_desc_sw128(a_lo + o);
                        const uint64_t dbh = lt_desc_sw128(b_hi + o), dbl = lt_desc_sw128(b_lo + o);
                        lt_umma_tf32(tmem_d, dal, dbh, (kb | ks) != 0 ? 1u : 0u);
                        lt_umma_tf32(tmem_d, dah, dbl, 1u);
                        lt_umma_tf32(tmem_d, dah, dbh, 1u);
                    }
                    lt_commit(&empty_bar[stage]);
                    if (kb == n_kb - 1) lt_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == LT_STAGES) { stage = 0; parity ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
    } else {
        // =============================== epilogue (warps 0-3): thread = row ==========================
        int acc = 0;
        uint32_t acc_parity = 0;
        for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int64_t row = (wk / n_cb) * LT_M + warp * 32 + lane;
            const int col0 = (int)(wk % n_cb) * LT_N;
            mbar_wait(&tfull_bar[acc], acc_parity);
            lt_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * LT_N;
#pragma unroll 1
            for (int q = 0; q < LT_N / 32; ++q) {
                float d[32];
                lt_tmem_ld32(taddr + q * 32, d);   // (all lanes: the load is warp-collective)
                const int c0 = col0 + q * 32;
                if (row < n_rows && c0 < n_out) {
                    float* yp = y + row * ldy + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = c0 + j + e;
                            const float z = d[j + e] + ((bias != nullptr && c < n_out) ? __ldg(bias + c) : 0.f);
                            o[e] = lt_act(z, c < split ? act0 : act1);
                        }
                        if (c0 + j + 3 < n_out && (ldy & 3) == 0) {
                            *reinterpret_cast<float4*>(yp + j) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c0 + j + e < n_out) yp[j + e] = o[e];
                        }
                    }
                }
            }
            lt_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_parity ^= 1u; }
        }
    }

    lt_fence_before();
    __syncthreads();
    if (warp == LT_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)LT_TMEM_COLS)
                     : "memory");
    }
}

}  // namespace moc

using namespace moc;

extern "C" size_t moc_linear_workspace_bytes(int n_out, int k) {
    if (n_out < 1 || k < LT_KB || k % LT_KB) return 0;
    const size_t n_cb = (size_t)(n_out + LT_N - 1) / LT_N;
    return n_cb * (size_t)(k / LT_KB) * 2 * LT_B_BYTES;
}

extern "C" int moc_linear_forward(const float* x, int64_t ldx, int64_t n_rows, int k, const float* w, const float* bias,
                                  int n_out, int act0, int split, int act1, float* y, int64_t ldy, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    MOC_CHECK_ARG(x && w && y && workspace, "moc_linear_forward: null pointer");
    MOC_CHECK_ARG(n_rows >= 0 && ldx >= k && ldy >= n_out, "moc_linear_forward: bad n_rows / leading dimensions");
    MOC_CHECK_SHAPE(k >= LT_KB && k % LT_KB == 0 && k <= 4096, "moc_linear_forward: in_features must be a multiple of %d, got %d",
                    LT_KB, k);
    MOC_CHECK_SHAPE(n_out >= 1 && n_out <= 8192, "moc_linear_forward: bad out_features %d", n_out);
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "moc_linear_forward: x, w, y and the workspace must be 16-byte aligned, ldx a multiple of 4");
    MOC_CHECK_ARG(act0 >= 0 && act0 <= MOC_ACT_SIGMOID && act1 >= 0 && act1 <= MOC_ACT_SIGMOID, "moc_linear_forward: bad activation");
    const size_t need = moc_linear_workspace_bytes(n_out, k);
    if (workspace_bytes < need) {
        set_error("moc_linear_forward: workspace %zu B < required %zu B", workspace_bytes, need);
        return MOC_E_WORKSPACE;
    }
    if (n_rows == 0) return MOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_cb = (n_out + LT_N - 1) / LT_N;
    float* w_split = reinterpret_cast<float*>(workspace);
    const int64_t chunks = (int64_t)n_cb * LT_N * (k / 4);
    linear_tc_prep_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(w, n_out, k, n_cb, w_split);
    MOC_LAUNCH_CHECK("linear_tc_prep_kernel");
    MOC_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM));
    const int64_t n_work = ((n_rows + LT_M - 1) / LT_M) * n_cb;
    const int grid = (int)(n_work < sm_count() ? n_work : sm_count());
    linear_tc_kernel<<<grid, LT_THREADS, LT_SMEM, st>>>(x, ldx, n_rows, k, w_split, bias, n_out, n_cb, act0, split, act1, y,
                                                       ldy);
    MOC_LAUNCH_CHECK("linear_tc_kernel");
    return MOC_OK;
}
