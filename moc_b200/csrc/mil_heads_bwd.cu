// Backward of ABMIL = CLAM_SB(instance_loss_fn=None) for one bag (models/model_clam.py:175-219 forward; the reference
// trains it through torch autograd in utils/core_utils.py:391-416: loss = CE(logits, label); loss.backward()).
//
// Forward (saved by the caller):  h1 = relu(x Wfc^T + bfc) [N,L];  ab = [tanh(h1 Wa^T + ba) | sigmoid(h1 Wb^T + bb)] [N,2D];
//   A_n = sum_d a_nd b_nd wc_d + bc;  p = softmax_n(A);  M = sum_n p_n h1_n [L];  logits = Wcls M + bcls.
// Backward from dlogits [C]:
//   dWcls = dlogits (x) M, dbcls = dlogits, dM = Wcls^T dlogits
//   dA_n = p_n (dM.h1_n - dM.M)                                   (softmax over the bag)
//   dwc_d = sum_n dA_n a_nd b_nd, dbc = sum_n dA_n
//   dZa_nd = dA_n wc_d b_nd (1 - a_nd^2),  dZb_nd = dA_n wc_d a_nd b_nd (1 - b_nd)      dZ = [dZa | dZb] [N,2D]
//   d[Wa;Wb] = dZ^T h1 (tensor cores, wgrad_tc.cu), d[ba;bb] = column sums of dZ
//   dh1 = dZ [Wa;Wb] (tensor cores, linear_tc.cu on the transposed weights) + p_n dM;  dz1 = dh1 * (h1 > 0)
//   dWfc = dz1^T x (tensor cores), dbfc = column sums of dz1
// Every reduction over the bag is two-stage with a fixed order (per-block partials, then one block in block order):
// deterministic, fp32 accumulation.
#include "common.cuh"

namespace moc {

constexpr int AB_THREADS = 256;
constexpr int AB_WARPS = AB_THREADS / 32;
constexpr int AB_MAXI = 16;   // attention width D <= 512

// ---- step 1 (one block): softmax statistics of a_raw, dM, dM.M, classifier gradients ---------------------------
__global__ void __launch_bounds__(1024)
abmil_bwd_prep_kernel(const float* __restrict__ a_raw, int64_t n_rows, const float* __restrict__ pooled, int L,
                      const float* __restrict__ wcls, int C, const float* __restrict__ dlogits,
                      float* __restrict__ dM /* [L] */, float* __restrict__ scal /* gmax, 1/gsum, dM.M */,
                      float* __restrict__ d_wcls, float* __restrict__ d_bcls) {
    __shared__ float red[32];
    __shared__ float bcast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int64_t r = tid; r < n_rows; r += blockDim.x) m = fmaxf(m, a_raw[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        float t = red[0];
        for (int w = 1; w < (int)blockDim.x / 32; ++w) t = fmaxf(t, red[w]);
        bcast = t;
    }
    __syncthreads();
    const float gmax = bcast;
    float s = 0.f;
    for (int64_t r = tid; r < n_rows; r += blockDim.x) s += expf(a_raw[r] - gmax);
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w];
        scal[0] = gmax;
        scal[1] = 1.0f / t;
    }
    float dot = 0.f;
    for (int l = tid; l < L; l += blockDim.x) {
        float v = 0.f;
        for (int c = 0; c < C; ++c) v = fmaf(wcls[(size_t)c * L + l], dlogits[c], v);
        dM[l] = v;
        dot = fmaf(v, pooled[l], dot);
        for (int c = 0; c < C; ++c) d_wcls[(size_t)c * L + l] = dlogits[c] * pooled[l];
    }
    dot = warp_sum(dot);
    __syncthreads();
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w];
        scal[2] = t;
    }
    if (tid < C) d_bcls[tid] = dlogits[tid];
}

// ---- step 2: per patch dA, dZ; per-block partial sums of dwc / d[ba;bb] / dbc -----------------------------------
__global__ void __launch_bounds__(AB_THREADS)
abmil_bwd_rows_kernel(const float* __restrict__ h1, int64_t ldh, int L, const float* __restrict__ ab, int64_t ldab, int Dh,
                      const float* __restrict__ a_raw, const float* __restrict__ wc, const float* __restrict__ dM,
                      const float* __restrict__ scal, int64_t n_rows, int64_t rows_per_block,
                      float* __restrict__ dZ /* [N][2Dh] */, float* __restrict__ p_out,
                      float* __restrict__ part /* [blocks][3Dh+1] */) {
    extern __shared__ float ab_s[];   // [AB_WARPS][3*Dh+1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float gmax = scal[0], inv = scal[1], dmm = scal[2];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < n_rows ? r0 + rows_per_block : n_rows;
    float acc_wc[AB_MAXI], acc_a[AB_MAXI], acc_b[AB_MAXI], acc_c = 0.f;
#pragma unroll
    for (int i = 0; i < AB_MAXI; ++i) acc_wc[i] = acc_a[i] = acc_b[i] = 0.f;
    for (int64_t row = r0 + warp; row < r1; row += AB_WARPS) {
        const float4* hp = reinterpret_cast<const float4*>(h1 + row * ldh);
        const float4* dp = reinterpret_cast<const float4*>(dM);
        float s = 0.f;
        for (int q = lane; q < L / 4; q += 32) {
            const float4 hv = __ldg(hp + q), dv = __ldg(dp + q);
            s = fmaf(hv.x, dv.x, s); s = fmaf(hv.y, dv.y, s); s = fmaf(hv.z, dv.z, s); s = fmaf(hv.w, dv.w, s);
        }
        s = warp_sum(s);
        const float p = expf(a_raw[row] - gmax) * inv;
        const float dA = p * (s - dmm);
        if (lane == 0) p_out[row] = p;
        acc_c += dA;
        const float* abp = ab + row * ldab;
        float* zp = dZ + row * (2 * (int64_t)Dh);
#pragma unroll
        for (int i = 0; i < AB_MAXI; ++i) {
            const int d = lane + 32 * i;
            if (d < Dh) {
                const float a = abp[d], b = abp[Dh + d];
                const float t = dA * __ldg(wc + d);
                const float dza = (t * b) * (1.0f - a * a);
                const float dzb = ((t * a) * (1.0f - b)) * b;
                zp[d] = dza;
                zp[Dh + d] = dzb;
                acc_wc[i] = fmaf(dA, a * b, acc_wc[i]);
                acc_a[i] += dza;
                acc_b[i] += dzb;
            }
        }
    }
    float* mine = ab_s + (size_t)warp * (3 * Dh + 1);
#pragma unroll
    for (int i = 0; i < AB_MAXI; ++i) {
        const int d = lane + 32 * i;
        if (d < Dh) {
            mine[d] = acc_wc[i];
            mine[Dh + d] = acc_a[i];
            mine[2 * Dh + d] = acc_b[i];
        }
    }
    if (lane == 0) mine[3 * Dh] = acc_c;
    __syncthreads();
    float* out = part + (size_t)blockIdx.x * (3 * Dh + 1);
    for (int j = tid; j < 3 * Dh + 1; j += AB_THREADS) {
        float t = 0.f;
        for (int w = 0; w < AB_WARPS; ++w) t += ab_s[(size_t)w * (3 * Dh + 1) + j];
        out[j] = t;
    }
}

// out[j] = sum_b part[b][j] in block order
__global__ void colsum_final_kernel(const float* __restrict__ part, int n_blocks, int width, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= width) return;
    float t = 0.f;
    for (int b = 0; b < n_blocks; ++b) t += part[(size_t)b * width + j];
    out[j] = t;
}

// ---- [rows][cols] -> [cols][rows] -------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

// ---- step 4: dz1 = (G1 + p_n dM) * (h1 > 0) in place; per-block column sums --------------------------------------
// thread = one float4 of columns; a block walks its rows in order.
__global__ void abmil_bwd_dz1_kernel(float* __restrict__ g1, const float* __restrict__ h1, int64_t ldh, int L,
                                     const float* __restrict__ p, const float* __restrict__ dM, int64_t n_rows,
                                     int64_t rows_per_block, float* __restrict__ part /* [blocks][L] */) {
    const int q = threadIdx.x;   // blockDim.x == L / 4
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < n_rows ? r0 + rows_per_block : n_rows;
    const float4 dm = reinterpret_cast<const float4*>(dM)[q];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0; r < r1; ++r) {
        float4* gp = reinterpret_cast<float4*>(g1 + r * L) + q;
        const float4 hv = __ldg(reinterpret_cast<const float4*>(h1 + r * ldh) + q);
        const float pr = p[r];
        float4 v = *gp;
        v.x = hv.x > 0.f ? v.x + pr * dm.x : 0.f;
        v.y = hv.y > 0.f ? v.y + pr * dm.y : 0.f;
        v.z = hv.z > 0.f ? v.z + pr * dm.z : 0.f;
        v.w = hv.w > 0.f ? v.w + pr * dm.w : 0.f;
        *gp = v;
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(part + (size_t)blockIdx.x * L)[q] = s;
}

struct AbmilPlan {
    int blocks_rows, blocks_dz1;
    int64_t rpb_rows, rpb_dz1;
    size_t off_dM, off_scal, off_p, off_dZ, off_g1, off_part, off_wT, off_lin, off_wg, total;
    size_t lin_bytes, wg_bytes;
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static AbmilPlan abmil_plan(int64_t n_rows, int k_in, int L, int Dh) {
    AbmilPlan p;
    const int64_t cap = (int64_t)sm_count() * 4;
    int64_t b = (n_rows + 63) / 64;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    p.rpb_rows = (n_rows + b - 1) / b;
    if (p.rpb_rows < 1) p.rpb_rows = 1;
    p.blocks_rows = (int)((n_rows + p.rpb_rows - 1) / p.rpb_rows);
    if (p.blocks_rows < 1) p.blocks_rows = 1;
    p.rpb_dz1 = p.rpb_rows;
    p.blocks_dz1 = p.blocks_rows;
    size_t o = 0;
    p.off_dM = o;   o += align256((size_t)L * 4);
    p.off_scal = o; o += 256;
    p.off_p = o;    o += align256((size_t)(n_rows > 0 ? n_rows : 1) * 4);
    p.off_dZ = o;   o += align256((size_t)(n_rows > 0 ? n_rows : 1) * 2 * Dh * 4);
    p.off_g1 = o;   o += align256((size_t)(n_rows > 0 ? n_rows : 1) * L * 4);
    const size_t w1 = (size_t)p.blocks_rows * (3 * Dh + 1), w2 = (size_t)p.blocks_dz1 * L;
    p.off_part = o; o += align256((w1 > w2 ? w1 : w2) * 4);
    p.off_wT = o;   o += align256((size_t)L * 2 * Dh * 4);
    p.lin_bytes = moc_linear_workspace_bytes(L, 2 * Dh);
    p.off_lin = o;  o += align256(p.lin_bytes);
    const size_t g1 = moc_linear_wgrad_workspace_bytes(n_rows, 2 * Dh, L), g2 = moc_linear_wgrad_workspace_bytes(n_rows, L, k_in);
    p.wg_bytes = g1 > g2 ? g1 : g2;
    p.off_wg = o;   o += align256(p.wg_bytes);
    p.total = o;
    return p;
}

}  // namespace moc

using namespace moc;

extern "C" size_t moc_abmil_backward_workspace_bytes(int64_t n_rows, int k_in, int width, int hidden) {
    if (n_rows < 0 || k_in < 1 || width < 1 || hidden < 1) return 0;
    return abmil_plan(n_rows, k_in, width, hidden).total;
}

extern "C" int moc_abmil_backward(const float* x, int64_t ldx, int k_in, int64_t n_rows, const float* h1, int64_t ldh,
                                  int width, const float* ab, int64_t ldab, int hidden, const float* a_raw,
                                  const float* pooled, const float* w_ab, const float* wc, const float* w_cls,
                                  int n_classes, const float* dlogits, float* d_wfc, float* d_bfc, float* d_wab,
                                  float* d_bab, float* d_wc, float* d_bc, float* d_wcls, float* d_bcls, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    MOC_CHECK_ARG(x && h1 && ab && a_raw && pooled && w_ab && wc && w_cls && dlogits && d_wfc && d_bfc && d_wab && d_bab &&
                      d_wc && d_bc && d_wcls && d_bcls && workspace,
                  "moc_abmil_backward: null pointer");
    MOC_CHECK_ARG(n_rows >= 1 && ldx >= k_in && ldh >= width && ldab >= 2 * hidden, "moc_abmil_backward: bad sizes");
    MOC_CHECK_SHAPE(width % 128 == 0 && width <= 4096, "moc_abmil_backward: hidden width %d must be a multiple of 128", width);
    MOC_CHECK_SHAPE(hidden >= 16 && hidden % 16 == 0 && hidden <= 32 * AB_MAXI,
                    "moc_abmil_backward: attention width %d must be a multiple of 16, at most %d", hidden, 32 * AB_MAXI);
    MOC_CHECK_SHAPE(k_in % 4 == 0 && n_classes >= 1 && n_classes <= 1024, "moc_abmil_backward: bad in_features %d / classes %d",
                    k_in, n_classes);
    MOC_CHECK_ARG((reinterpret_cast<uintptr_t>(h1) & 15) == 0 && (ldh & 3) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                  "moc_abmil_backward: h1 must be 16-byte aligned with ldh a multiple of 4, the workspace 256-byte aligned");
    const AbmilPlan p = abmil_plan(n_rows, k_in, width, hidden);
    if (workspace_bytes < p.total) {
        set_error("moc_abmil_backward: workspace %zu B < required %zu B", workspace_bytes, p.total);
        return MOC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = reinterpret_cast<char*>(workspace);
    float* dM = reinterpret_cast<float*>(ws + p.off_dM);
    float* scal = reinterpret_cast<float*>(ws + p.off_scal);
    float* pr = reinterpret_cast<float*>(ws + p.off_p);
    float* dZ = reinterpret_cast<float*>(ws + p.off_dZ);
    float* g1 = reinterpret_cast<float*>(ws + p.off_g1);
    float* part = reinterpret_cast<float*>(ws + p.off_part);
    float* wT = reinterpret_cast<float*>(ws + p.off_wT);
    const int Dh = hidden, L = width;

    abmil_bwd_prep_kernel<<<1, 1024, 0, st>>>(a_raw, n_rows, pooled, L, w_cls, n_classes, dlogits, dM, scal, d_wcls, d_bcls);
    MOC_LAUNCH_CHECK("abmil_bwd_prep_kernel");
    const size_t smem = (size_t)AB_WARPS * (3 * Dh + 1) * sizeof(float);   // 36 KB at Dh = 384, 49 KB at the 512 maximum
    MOC_CUDA(cudaFuncSetAttribute(abmil_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((size_t)AB_WARPS * (3 * 32 * AB_MAXI + 1) * sizeof(float))));
    abmil_bwd_rows_kernel<<<p.blocks_rows, AB_THREADS, smem, st>>>(h1, ldh, L, ab, ldab, Dh, a_raw, wc, dM, scal, n_rows,
                                                                   p.rpb_rows, dZ, pr, part);
    MOC_LAUNCH_CHECK("abmil_bwd_rows_kernel");
    // part columns: [dwc Dh | d ba Dh | d bb Dh | dbc]
    colsum_final_kernel<<<(3 * Dh + 1 + 127) / 128, 128, 0, st>>>(part, p.blocks_rows, 3 * Dh + 1, g1 /* scratch */);
    MOC_LAUNCH_CHECK("colsum_final_kernel");
    MOC_CUDA(cudaMemcpyAsync(d_wc, g1, (size_t)Dh * 4, cudaMemcpyDeviceToDevice, st));
    MOC_CUDA(cudaMemcpyAsync(d_bab, g1 + Dh, (size_t)2 * Dh * 4, cudaMemcpyDeviceToDevice, st));
    MOC_CUDA(cudaMemcpyAsync(d_bc, g1 + 3 * Dh, 4, cudaMemcpyDeviceToDevice, st));

    // dh1 (through the attention branches) = dZ [Wa;Wb]:  a linear layer with weight [Wa;Wb]^T  [L][2Dh]
    transpose_kernel<<<dim3((L + 31) / 32, (2 * Dh + 31) / 32), dim3(32, 8), 0, st>>>(w_ab, 2 * Dh, L, wT);
    MOC_LAUNCH_CHECK("transpose_kernel");
    int rc = moc_linear_forward(dZ, 2 * Dh, n_rows, 2 * Dh, wT, nullptr, L, MOC_ACT_NONE, L, MOC_ACT_NONE, g1, L, ws + p.off_lin,
                                p.lin_bytes, stream);
    if (rc != MOC_OK) return rc;
    abmil_bwd_dz1_kernel<<<p.blocks_dz1, L / 4, 0, st>>>(g1, h1, ldh, L, pr, dM, n_rows, p.rpb_dz1, part);
    MOC_LAUNCH_CHECK("abmil_bwd_dz1_kernel");
    colsum_final_kernel<<<(L + 127) / 128, 128, 0, st>>>(part, p.blocks_dz1, L, d_bfc);
    MOC_LAUNCH_CHECK("colsum_final_kernel");

    rc = moc_linear_wgrad(dZ, 2 * Dh, 2 * Dh, h1, ldh, L, n_rows, d_wab, L, 0, ws + p.off_wg, p.wg_bytes, stream);
    if (rc != MOC_OK) return rc;
    return moc_linear_wgrad(g1, L, L, x, ldx, k_in, n_rows, d_wfc, k_in, 0, ws + p.off_wg, p.wg_bytes, stream);
}

// =================================================================================================================
// Conch_CLIP_Ada and MIL_fc: the rows that carry gradient are few (the top-j rows of each class; the single
// max-probability instance), so their backward is a handful of row kernels around the dense tensor-core layers.
// =================================================================================================================
namespace moc {

// Conch_CLIP_Ada.forward (models/model_adapters.py:185-193) backwards, for "virtual rows" v = (row pooled for class
// c_v): f = ratio * a2 + (1 - ratio) * x, fn = f / |f|, logit = fn . classifier[:, c];  given g_v = d(loss)/d(logit)
// the gradient at the adapter output is  da2 = ratio * (cls_c - fn (fn . cls_c)) g_v / |f|  where a2 > 0 (its ReLU).
// One warp per virtual row; x / a2 are the gathered rows [R][512].
__global__ void __launch_bounds__(256)
adapter_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ a2, float ratio,
                        const float* __restrict__ classifier /* [512][C] */, int C, const int32_t* __restrict__ cls_of_row,
                        const float* __restrict__ g_of_row, int64_t n_rows, float* __restrict__ da2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float keep = 1.0f - ratio;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n_rows; row += (int64_t)gridDim.x * 8) {
        const int c = cls_of_row[row];
        const float g = g_of_row[row];
        float f[16], av[16], w[16];
        float ss = 0.f, dot = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k = q * 32 + lane;
            av[q] = a2[row * D + k];
            f[q] = __fadd_rn(__fmul_rn(av[q], ratio), __fmul_rn(x[row * D + k], keep));
            w[q] = __ldg(classifier + (size_t)k * C + c);
            ss = fmaf(f[q], f[q], ss);
        }
        const float nrm = sqrtf(warp_sum(ss));
        const float inv = 1.0f / nrm;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            f[q] *= inv;                      // fn
            dot = fmaf(f[q], w[q], dot);
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float df = (w[q] - f[q] * dot) * (g * inv);
            da2[row * D + q * 32 + lane] = av[q] > 0.f ? ratio * df : 0.f;
        }
    }
}

// g[i] = ref[i] > 0 ? g[i] : 0   (ReLU backward on a small dense block)
__global__ void mask_positive_kernel(float* __restrict__ g, const float* __restrict__ ref, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(ref[i] > 0.f)) g[i] = 0.f;
}

// MIL_fc (models/model_mil.py:30-51): only the selected instance carries gradient.  x [K0] its features, hid [H1] its
// hidden activations (after ReLU), dtop [C] = d(loss)/d(top_instance logits).  Block j handles hidden unit j.
__global__ void __launch_bounds__(128)
mil_fc_backward_kernel(const float* __restrict__ x, int K0, const float* __restrict__ hid, int H1,
                       const float* __restrict__ w_last /* [C][H1] */, int C, const float* __restrict__ dtop,
                       float* __restrict__ d_w0 /* [H1][K0] */, float* __restrict__ d_b0, float* __restrict__ d_wl,
                       float* __restrict__ d_bl) {
    const int j = blockIdx.x;
    const float h = hid[j];
    float dh = 0.f;
    for (int c = 0; c < C; ++c) dh = fmaf(dtop[c], w_last[(size_t)c * H1 + j], dh);
    if (!(h > 0.f)) dh = 0.f;
    for (int k = threadIdx.x; k < K0; k += blockDim.x) d_w0[(size_t)j * K0 + k] = dh * x[k];
    if (threadIdx.x == 0) {
        d_b0[j] = dh;
        for (int c = 0; c < C; ++c) d_wl[(size_t)c * H1 + j] = dtop[c] * h;
        if (j == 0)
            for (int c = 0; c < C; ++c) d_bl[c] = dtop[c];
    }
}

}  // namespace moc

extern "C" int moc_transpose(const float* in, int rows, int cols, float* out, void* stream) {
    MOC_CHECK_ARG(in && out && rows >= 1 && cols >= 1, "moc_transpose: bad arguments");
    transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(in, rows, cols, out);
    MOC_LAUNCH_CHECK("transpose_kernel");
    return MOC_OK;
}

extern "C" int moc_adapter_backward_rows(const float* x_rows, const float* a2_rows, float clip_ratio, const float* classifier,
                                         int n_classes, const int32_t* cls_of_row, const float* g_of_row, int64_t n_rows,
                                         float* da2, void* stream) {
    MOC_CHECK_ARG(x_rows && a2_rows && classifier && cls_of_row && g_of_row && da2 && n_rows >= 0 && n_classes >= 1,
                  "moc_adapter_backward_rows: bad arguments");
    if (n_rows == 0) return MOC_OK;
    int64_t blocks = (n_rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    adapter_bwd_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x_rows, a2_rows, clip_ratio, classifier, n_classes,
                                                                                 cls_of_row, g_of_row, n_rows, da2);
    MOC_LAUNCH_CHECK("adapter_bwd_rows_kernel");
    return MOC_OK;
}

extern "C" int moc_mask_positive(float* g, const float* ref, int64_t n, void* stream) {
    MOC_CHECK_ARG(g && ref && n >= 0, "moc_mask_positive: bad arguments");
    if (n == 0) return MOC_OK;
    mask_positive_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, ref, n);
    MOC_LAUNCH_CHECK("mask_positive_kernel");
    return MOC_OK;
}

extern "C" int moc_mil_fc_backward(const float* x_row, int k_in, const float* hid_row, int width, const float* w_last,
                                   int n_classes, const float* dtop, float* d_w0, float* d_b0, float* d_wl, float* d_bl,
                                   void* stream) {
    MOC_CHECK_ARG(x_row && hid_row && w_last && dtop && d_w0 && d_b0 && d_wl && d_bl, "moc_mil_fc_backward: null pointer");
    MOC_CHECK_ARG(k_in >= 1 && width >= 1 && n_classes >= 1, "moc_mil_fc_backward: bad sizes");
    mil_fc_backward_kernel<<<width, 128, 0, (cudaStream_t)stream>>>(x_row, k_in, hid_row, width, w_last, n_classes, dtop, d_w0,
                                                                    d_b0, d_wl, d_bl);
    MOC_LAUNCH_CHECK("mil_fc_backward_kernel");
    return MOC_OK;
}
