// Row-wise pieces of the secondary MIL heads around the dense layers of linear_tc.cu.
//
//   moc_adapter_scores          Conch_CLIP_Ada.forward / forward_disable_ada after the adapter MLP: residual blend,
//                               L2 normalisation, similarity to the classifier matrix (models/model_adapters.py:186-190,
//                               :211-214); the per-class top-j mean that follows is moc_pool_topk.
//   moc_gated_attention_scores  Attn_Net_Gated: A = (tanh branch * sigmoid branch) . w_c + b_c  (models/model_clam.py:59-62)
//   moc_attention_pool          CLAM_SB.forward_single: softmax of A over the bag, M = A h, classifier, softmax, argmax
//                               (models/model_clam.py:181, :209-212)
//   moc_row_softmax             MIL_fc: per-patch softmax of the instance logits (models/model_mil.py:38)
// All of these are HBM-bound single passes over [N, 512..768] activations with fp32 accumulation.
#include "common.cuh"

namespace moc {

// ---- a15: logits[c][n] = < normalise(ratio * a[n] + (1 - ratio) * x[n]) , classifier[:, c] > -------------------
// warp per patch, lanes split the 512 components (4 x float4 each); classifier [512][C] staged K-major in smem.
constexpr int AS_WARPS = 8;
template <int MAXC>
__global__ void __launch_bounds__(AS_WARPS * 32)
adapter_scores_kernel(const float* __restrict__ x, const float* __restrict__ a, float ratio,
                      const float* __restrict__ classifier, int C, int64_t n_rows, float* __restrict__ logits, int64_t ld) {
    extern __shared__ float as_w[];  // [C][512]
    for (int i = threadIdx.x; i < C * D; i += blockDim.x) {
        const int c = i / D, k = i % D;
        as_w[i] = classifier[(size_t)k * C + c];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float keep = 1.0f - ratio;
    for (int64_t row = (int64_t)blockIdx.x * AS_WARPS + warp; row < n_rows; row += (int64_t)gridDim.x * AS_WARPS) {
        float4 f[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + row * D) + q * 32 + lane);
            if (a != nullptr) {
                const float4 av = __ldcs(reinterpret_cast<const float4*>(a + row * D) + q * 32 + lane);
                // adapted * clip_ratio + feat * (1 - clip_ratio), in that order (models/model_adapters.py:187)
                f[q] = make_float4(__fadd_rn(__fmul_rn(av.x, ratio), __fmul_rn(xv.x, keep)),
                                   __fadd_rn(__fmul_rn(av.y, ratio), __fmul_rn(xv.y, keep)),
                                   __fadd_rn(__fmul_rn(av.z, ratio), __fmul_rn(xv.z, keep)),
                                   __fadd_rn(__fmul_rn(av.w, ratio), __fmul_rn(xv.w, keep)));
            } else {
                f[q] = xv;
            }
        }
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) ss += f[q].x * f[q].x + f[q].y * f[q].y + f[q].z * f[q].z + f[q].w * f[q].w;
        const float inv = 1.0f / sqrtf(warp_sum(ss));   // x / x.norm(): no epsilon in the reference
#pragma unroll
        for (int q = 0; q < 4; ++q) f[q] = make_float4(f[q].x * inv, f[q].y * inv, f[q].z * inv, f[q].w * inv);
        for (int c = 0; c < C; ++c) {
            const float4* wc = reinterpret_cast<const float4*>(as_w + c * D);
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = wc[q * 32 + lane];
                s = fmaf(f[q].x, wv.x, s); s = fmaf(f[q].y, wv.y, s); s = fmaf(f[q].z, wv.z, s); s = fmaf(f[q].w, wv.w, s);
            }
            s = warp_sum(s);
            if (lane == 0) logits[(int64_t)c * ld + row] = s;
        }
    }
}

// ---- a16: A[n] = sum_d ab[n][d] * ab[n][Dh + d] * wc[d] + bc ---------------------------------------------------
__global__ void __launch_bounds__(256)
gated_attention_scores_kernel(const float* __restrict__ ab, int64_t ld, int Dh, const float* __restrict__ wc, float bc,
                              int64_t n_rows, float* __restrict__ a_raw) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n_rows; row += (int64_t)gridDim.x * 8) {
        const float* p = ab + row * ld;
        float s = 0.f;
        for (int d = lane; d < Dh; d += 32) s = fmaf(p[d] * p[Dh + d], __ldg(wc + d), s);
        s = warp_sum(s);
        if (lane == 0) a_raw[row] = s + bc;
    }
}

// ---- a16: softmax over the bag + attention-weighted mean of h -------------------------------------------------
// Pass 1: block b reduces rows [b*chunk, (b+1)*chunk) to (running max m_b, sum of exp s_b, M_b[L] = sum exp(A-m_b) h).
// Pass 2: one block merges the partials in block order (deterministic), then the bag classifier, softmax, argmax.
constexpr int AP_THREADS = 256;
__global__ void __launch_bounds__(AP_THREADS)
attention_pool_partial_kernel(const float* __restrict__ a_raw, const float* __restrict__ h, int64_t ldh, int L,
                              int64_t n_rows, int64_t chunk, float* __restrict__ part /* [blocks][L + 2] */) {
    __shared__ float red[AP_THREADS / 32];
    __shared__ float m_s;
    const int64_t r0 = (int64_t)blockIdx.x * chunk;
    const int64_t r1 = r0 + chunk < n_rows ? r0 + chunk : n_rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int64_t r = r0 + tid; r < r1; r += AP_THREADS) m = fmaxf(m, a_raw[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        float t = red[0];
        for (int w = 1; w < AP_THREADS / 32; ++w) t = fmaxf(t, red[w]);
        m_s = t;
    }
    __syncthreads();
    m = m_s;
    float* out = part + (size_t)blockIdx.x * (L + 2);
    // thread t owns columns t, t + 256, ...: a fixed row order per column keeps the sum reproducible
    float s = 0.f;
    for (int c = tid; c < L; c += AP_THREADS) {
        float acc = 0.f;
        for (int64_t r = r0; r < r1; ++r) acc = fmaf(expf(a_raw[r] - m), h[r * ldh + c], acc);
        out[c] = acc;
    }
    if (tid == 0) {
        for (int64_t r = r0; r < r1; ++r) s += expf(a_raw[r] - m);
        out[L] = m;
        out[L + 1] = s;
    }
}

__global__ void __launch_bounds__(AP_THREADS)
attention_pool_final_kernel(const float* __restrict__ part, int n_blocks, int L, const float* __restrict__ wcls /* [C][L] */,
                            const float* __restrict__ bcls, int C, float* __restrict__ m_out /* [L] */,
                            float* __restrict__ logits, float* __restrict__ probs, int32_t* __restrict__ y_hat) {
    extern __shared__ float ap_m[];  // [L] pooled feature, then [C] logits
    __shared__ float gmax_s, gsum_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        float gm = -INFINITY;
        for (int b = 0; b < n_blocks; ++b) gm = fmaxf(gm, part[(size_t)b * (L + 2) + L]);
        float gs = 0.f;
        for (int b = 0; b < n_blocks; ++b) gs += part[(size_t)b * (L + 2) + L + 1] * expf(part[(size_t)b * (L + 2) + L] - gm);
        gmax_s = gm;
        gsum_s = gs;
    }
    __syncthreads();
    const float gm = gmax_s, inv = 1.0f / gsum_s;
    for (int c = tid; c < L; c += AP_THREADS) {
        float acc = 0.f;
        for (int b = 0; b < n_blocks; ++b)
            acc = fmaf(part[(size_t)b * (L + 2) + c], expf(part[(size_t)b * (L + 2) + L] - gm), acc);
        const float v = acc * inv;
        ap_m[c] = v;
        if (m_out) m_out[c] = v;
    }
    __syncthreads();
    float* lg = ap_m + L;
    for (int c = warp; c < C; c += AP_THREADS / 32) {
        float s = 0.f;
        for (int k = lane; k < L; k += 32) s = fmaf(ap_m[k], wcls[(size_t)c * L + k], s);
        s = warp_sum(s);
        if (lane == 0) {
            lg[c] = s + (bcls ? bcls[c] : 0.f);
            logits[c] = lg[c];
        }
    }
    __syncthreads();
    if (tid == 0) {
        float mx = -INFINITY;
        int arg = 0;
        for (int c = 0; c < C; ++c)
            if (lg[c] > mx) { mx = lg[c]; arg = c; }
        float es = 0.f;
        for (int c = 0; c < C; ++c) es += expf(lg[c] - mx);
        if (probs)
            for (int c = 0; c < C; ++c) probs[c] = expf(lg[c] - mx) / es;
        if (y_hat) y_hat[0] = arg;
    }
}

// ---- a17: row softmax of small logit rows [N][C] --------------------------------------------------------------
__global__ void row_softmax_kernel(const float* __restrict__ logits, int64_t ld, int C, int64_t n_rows,
                                   float* __restrict__ probs, int64_t ldp) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const float* p = logits + r * ld;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, p[c]);
    float es = 0.f;
    for (int c = 0; c < C; ++c) es += expf(p[c] - mx);
    for (int c = 0; c < C; ++c) probs[r * ldp + c] = expf(p[c] - mx) / es;
}

}  // namespace moc

using namespace moc;

extern "C" int moc_adapter_scores(const float* x, const float* adapted, float clip_ratio, const float* classifier,
                                  int n_classes, int64_t n_rows, float* logits, int64_t ld, void* stream) {
    MOC_CHECK_ARG(x && classifier && logits, "moc_adapter_scores: null pointer");
    MOC_CHECK_ARG(n_rows >= 0 && ld >= n_rows, "moc_adapter_scores: bad n_rows / ld");
    MOC_CHECK_SHAPE(n_classes >= 1 && n_classes <= MOC_MAX_COLS, "moc_adapter_scores: bad class count %d", n_classes);
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(adapted) & 15) == 0,
                  "moc_adapter_scores: inputs must be 16-byte aligned");
    if (n_rows == 0) return MOC_OK;
    const size_t smem = (size_t)n_classes * D * sizeof(float);
    MOC_CUDA(cudaFuncSetAttribute(adapter_scores_kernel<MOC_MAX_COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  MOC_MAX_COLS * D * (int)sizeof(float)));
    int64_t blocks = (n_rows + AS_WARPS - 1) / AS_WARPS;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    adapter_scores_kernel<MOC_MAX_COLS><<<(unsigned)blocks, AS_WARPS * 32, smem, (cudaStream_t)stream>>>(
        x, adapted, clip_ratio, classifier, n_classes, n_rows, logits, ld);
    MOC_LAUNCH_CHECK("adapter_scores_kernel");
    return MOC_OK;
}

extern "C" int moc_gated_attention_scores(const float* ab, int64_t ld, int hidden, const float* wc, float bc, int64_t n_rows,
                                          float* a_raw, void* stream) {
    MOC_CHECK_ARG(ab && wc && a_raw && hidden >= 1 && ld >= 2 * hidden && n_rows >= 0, "moc_gated_attention_scores: bad arguments");
    if (n_rows == 0) return MOC_OK;
    int64_t blocks = (n_rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    gated_attention_scores_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ab, ld, hidden, wc, bc, n_rows, a_raw);
    MOC_LAUNCH_CHECK("gated_attention_scores_kernel");
    return MOC_OK;
}

static int attention_pool_blocks(int64_t n_rows) {
    int64_t b = (n_rows + 63) / 64;   // at least 64 patches per partial
    const int64_t cap = (int64_t)sm_count() * 4;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

extern "C" size_t moc_attention_pool_workspace_bytes(int64_t n_rows, int width) {
    return (size_t)attention_pool_blocks(n_rows) * (size_t)(width + 2) * sizeof(float);
}

extern "C" int moc_attention_pool(const float* a_raw, const float* h, int64_t ldh, int width, int64_t n_rows,
                                  const float* w_cls, const float* b_cls, int n_classes, float* pooled, float* logits,
                                  float* probs, int32_t* y_hat, void* workspace, size_t workspace_bytes, void* stream) {
    MOC_CHECK_ARG(a_raw && h && w_cls && logits && workspace, "moc_attention_pool: null pointer");
    MOC_CHECK_ARG(n_rows >= 1 && width >= 1 && ldh >= width && n_classes >= 1, "moc_attention_pool: bad sizes");
    MOC_CHECK_SHAPE(width <= 4096 && n_classes <= 1024, "moc_attention_pool: width %d / classes %d too large", width, n_classes);
    const size_t need = moc_attention_pool_workspace_bytes(n_rows, width);
    if (workspace_bytes < need) {
        set_error("moc_attention_pool: workspace %zu B < required %zu B", workspace_bytes, need);
        return MOC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = attention_pool_blocks(n_rows);
    const int64_t chunk = (n_rows + blocks - 1) / blocks;
    float* part = reinterpret_cast<float*>(workspace);
    attention_pool_partial_kernel<<<blocks, AP_THREADS, 0, st>>>(a_raw, h, ldh, width, n_rows, chunk, part);
    MOC_LAUNCH_CHECK("attention_pool_partial_kernel");
    const int used = (int)((n_rows + chunk - 1) / chunk);
    attention_pool_final_kernel<<<1, AP_THREADS, (size_t)(width + n_classes) * sizeof(float), st>>>(
        part, used, width, w_cls, b_cls, n_classes, pooled, logits, probs, y_hat);
    MOC_LAUNCH_CHECK("attention_pool_final_kernel");
    return MOC_OK;
}

extern "C" int moc_row_softmax(const float* logits, int64_t ld, int n_cols, int64_t n_rows, float* probs, int64_t ldp,
                               void* stream) {
    MOC_CHECK_ARG(logits && probs && n_cols >= 1 && ld >= n_cols && ldp >= n_cols && n_rows >= 0, "moc_row_softmax: bad arguments");
    if (n_rows == 0) return MOC_OK;
    row_softmax_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, ld, n_cols, n_rows, probs, ldp);
    MOC_LAUNCH_CHECK("row_softmax_kernel");
    return MOC_OK;
}
