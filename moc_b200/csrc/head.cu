// Head of the MOC path: meta-learner gate on the selected patches, classifier-bank combination, top-K
// pooling into bag logits, cross-entropy, backward and Adam.
//
// Replaces main_moc.py:299-312 (senet), :390-405 / :481-493 (gate, gated sum, topj_pooling), :406-410
// (cross_entropy, backward, optimizer.step) and utils/patch_selection_classifier.py:18-32.
//
// The four score planes of a selected row are not recomputed from the features (main_moc.py:356-366 does
// `selected_feat @ W` again): they are the key planes the streaming kernel already wrote for that row.
// All arithmetic is fp32 with fp32 accumulation.  Reductions that decide results (pooling order, gradient
// sums) run in a fixed order, so repeated runs are bit-identical.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace moc {

// head_tc.cu: the tcgen05 (3xTF32) implementation of the gate MLP
size_t head_tc_workspace_bytes();
int launch_head_rows_tc(const float* feat, const float* keys, int64_t key_stride, int C, const int32_t* sel_rows,
                        int64_t n_slots, const float* w1, const float* b1, const float* w2, const float* b2,
                        unsigned active_mask, float* gate, float* final_scores, void* workspace, cudaStream_t st);

// head_f16.cu: the tcgen05 FP16x3 implementation (default): half the shared-memory bytes per unit of K, W1 resident
size_t head_f16_workspace_bytes();
int launch_head_rows_f16(const float* feat, const float* keys, int64_t key_stride, int C, const int32_t* sel_rows,
                         int64_t n_slots, const float* w1, const float* b1, const float* w2, const float* b2,
                         unsigned active_mask, float* gate, float* final_scores, void* workspace, int* domain_flag,
                         cudaStream_t st);

// MOC_HEAD_IMPL = auto (default: FP16x3, see head_f16.cu) | f16 | tf32 | simt.  The 3xTF32 kernel serves features
// outside the FP16x3 kernel's |x| < 4094 range (MOC_HEAD_WIDE_DOMAIN); it and the CUDA-core kernel are also the
// in-tree cross-checks.
static int head_impl() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("MOC_HEAD_IMPL");
        cached = (e && e[0] == 's') ? 2 : (e && e[0] == 't') ? 1 : (e && e[0] == 'f') ? 0 : 3;
    }
    return cached;
}
static bool use_simt_head() { return head_impl() == 2; }
static size_t head_img_bytes() {
    const size_t a = head_tc_workspace_bytes(), b = head_f16_workspace_bytes();
    return ((a > b ? a : b) + 15) & ~(size_t)15;
}
// workspace = [ W1 image of whichever kernel runs | int32 domain flag (+ padding) ]
static size_t head_ws_bytes() { return head_img_bytes() + 16; }
static int launch_head_rows_mma(const float* feat, const float* keys, int64_t key_stride, int C, const int32_t* sel_rows,
                                int64_t n_slots, const float* w1, const float* b1, const float* w2, const float* b2,
                                unsigned active_mask, float* gate, float* final_scores, void* workspace, cudaStream_t st) {
    const int impl = head_impl();
    const bool wide = (active_mask & MOC_HEAD_WIDE_DOMAIN) != 0;   // the caller asks for the range-free kernel
    int* flag = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(workspace) + head_img_bytes());
    (void)C;   // the FP16x3 kernel with its A operand in tensor memory is ahead at every class count (C = 30: 5.98 vs 6.55 ms)
    return (impl == 1 || wide)
               ? launch_head_rows_tc(feat, keys, key_stride, C, sel_rows, n_slots, w1, b1, w2, b2, active_mask, gate,
                                     final_scores, workspace, st)
               : launch_head_rows_f16(feat, keys, key_stride, C, sel_rows, n_slots, w1, b1, w2, b2, active_mask, gate,
                                      final_scores, workspace, flag, st);
}

constexpr int H = MOC_HIDDEN;  // 64
constexpr int G = MOC_GATES;   // 4

// ------------------------------------------------------------------------------------------------
// Gate + combination for every selected row (slot).  Persistent CTAs keep W1 resident in shared memory.
// Tile = 32 slots; 256 threads; thread (ty,tx) owns rows {ty, ty+16} x hidden units {tx, tx+16, tx+32, tx+48}.
// ------------------------------------------------------------------------------------------------
constexpr int HR_TM = 32;
constexpr int HR_THREADS = 256;
constexpr int HR_LD = D + 4;  // padded row stride (floats): 516/4 = 129 is odd -> conflict-free LDS.128

__global__ void __launch_bounds__(HR_THREADS, 1)
head_rows_kernel(const float* __restrict__ feat, const float* __restrict__ keys, int64_t key_stride, int C,
                 const int32_t* __restrict__ sel_rows, int64_t n_slots, const float* __restrict__ w1,
                 const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                 unsigned active_mask, float* __restrict__ gate, float* __restrict__ final_scores) {
    extern __shared__ __align__(16) float hsm[];
    float* w1s = hsm;                  // [64][516]
    float* xs = hsm + H * HR_LD;       // [32][516]
    __shared__ int32_t rows_s[HR_TM];
    __shared__ float w2s[G * H], b1s[H], b2s[G];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ty = tid >> 4, tx = tid & 15;
    for (int i = tid; i < H * (D / 4); i += HR_THREADS) {
        const int j = i / (D / 4), k4 = i % (D / 4);
        *reinterpret_cast<float4*>(w1s + j * HR_LD + k4 * 4) = __ldg(reinterpret_cast<const float4*>(w1) + i);
    }
    if (tid < G * H) w2s[tid] = w2[tid];
    if (tid < H) b1s[tid] = b1[tid];
    if (tid < G) b2s[tid] = b2[tid];
    __syncthreads();

    const float a0 = (active_mask & MOC_CLS_TOPK) ? 1.f : 0.f;
    const float a1 = (active_mask & MOC_CLS_DELTA_SOFTMAX) ? 1.f : 0.f;
    const float a2 = (active_mask & MOC_CLS_DELTA_DIFF) ? 1.f : 0.f;
    const float a3 = (active_mask & MOC_CLS_BOTTOMK) ? 1.f : 0.f;

    const int64_t n_tiles = (n_slots + HR_TM - 1) / HR_TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t slot0 = tile * HR_TM;
        if (tid < HR_TM) {
            const int64_t s = slot0 + tid;
            rows_s[tid] = s < n_slots ? (sel_rows ? sel_rows[s] : (int32_t)s) : -1;
        }
        __syncthreads();
        // gather: warp w loads rows w, w+8, w+16, w+24 (2 KB each, coalesced)
        bool any = false;
#pragma unroll
        for (int rr = 0; rr < HR_TM / 8; ++rr) {
            const int r = warp + rr * 8;
            const int32_t row = rows_s[r];
            float4 v[4];
            if (row >= 0) {
                any = true;
                const float4* src = reinterpret_cast<const float4*>(feat + (int64_t)row * D);
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __ldg(src + q * 32 + lane);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(xs + r * HR_LD + q * 128 + lane * 4) = v[q];
        }
        const int tile_any = __syncthreads_or(any);
        if (tile_any) {
            float acc[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[r][u] = 0.f;
            const float* xa = xs + ty * HR_LD;
            const float* xb = xs + (ty + 16) * HR_LD;
            const float* wp = w1s + tx * HR_LD;
#pragma unroll 4
            for (int k4 = 0; k4 < D / 4; ++k4) {
                const float4 va = *reinterpret_cast<const float4*>(xa + k4 * 4);
                const float4 vb = *reinterpret_cast<const float4*>(xb + k4 * 4);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 wv = *reinterpret_cast<const float4*>(wp + u * 16 * HR_LD + k4 * 4);
                    acc[0][u] = fmaf(va.x, wv.x, acc[0][u]);
                    acc[0][u] = fmaf(va.y, wv.y, acc[0][u]);
                    acc[0][u] = fmaf(va.z, wv.z, acc[0][u]);
                    acc[0][u] = fmaf(va.w, wv.w, acc[0][u]);
                    acc[1][u] = fmaf(vb.x, wv.x, acc[1][u]);
                    acc[1][u] = fmaf(vb.y, wv.y, acc[1][u]);
                    acc[1][u] = fmaf(vb.z, wv.z, acc[1][u]);
                    acc[1][u] = fmaf(vb.w, wv.w, acc[1][u]);
                }
            }
            // second layer: partial over this thread's 4 hidden units, then all-reduce over the 16 tx lanes
            float z[2][G];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int m = 0; m < G; ++m) z[r][m] = 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = tx + 16 * u;
                    const float h = relu_nan(acc[r][u] + b1s[j]);
#pragma unroll
                    for (int m = 0; m < G; ++m) z[r][m] = fmaf(h, w2s[m * H + j], z[r][m]);
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int m = 0; m < G; ++m) z[r][m] += __shfl_xor_sync(FULL, z[r][m], o);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int rt = ty + 16 * r;
                const int32_t row = rows_s[rt];
                if (row < 0) continue;
                const int64_t slot = slot0 + rt;
                float g[G];
#pragma unroll
                for (int m = 0; m < G; ++m) g[m] = sigmoidf_exact(z[r][m] + b2s[m]);
                if (gate != nullptr && tx < G) gate[slot * G + tx] = tx == 0 ? g[0] : tx == 1 ? g[1] : tx == 2 ? g[2] : g[3];
                if (final_scores == nullptr) continue;
                const float* kp = keys + row;
                const KeyLayout kl = key_layout(C);
                const float dlt = kp[(int64_t)kl.diff * key_stride];
                const float bgm = kp[(int64_t)kl.bg_max * key_stride];
                RowSoftmax rs = {0.f};
                if (kl.compact) rs = row_softmax_of(kp, key_stride, kl);
                for (int c = tx; c < C; c += 16) {
                    // same association as the reference: ((g0*L + g1*P) + g2*delta) + g3*bg, no fma contraction
                    const float lt = kp[(int64_t)c * key_stride];
                    const float ls = kl.compact ? rs.of(lt) : kp[(int64_t)(C + c) * key_stride];
                    float f = a0 * __fmul_rn(g[0], lt);
                    f = __fadd_rn(f, a1 * __fmul_rn(g[1], ls));
                    f = __fadd_rn(f, a2 * __fmul_rn(g[2], dlt));
                    f = __fadd_rn(f, a3 * __fmul_rn(g[3], bgm));
                    final_scores[slot * C + c] = f;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Per (slide, class): mean of the min(topk, S) largest combined scores; remembers which rows won.
// One warp per (slide, class); K rounds of "largest key strictly below the previous winner".
// ------------------------------------------------------------------------------------------------
constexpr int POOL_WARPS = 4;
__global__ void __launch_bounds__(POOL_WARPS * 32)
pool_final_kernel(const float* __restrict__ final_scores, const int64_t* __restrict__ sel_base,
                  const int32_t* __restrict__ sel_count, int n_slides, int C, int topk,
                  float* __restrict__ bag_logits, int32_t* __restrict__ pool_pos) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * POOL_WARPS + (threadIdx.x >> 5);
    if (item >= (int64_t)n_slides * C) return;
    const int slide = (int)(item / C), c = (int)(item % C);
    const int S = sel_count[slide];
    const float* f = final_scores + sel_base[slide] * C + c;
    const int k_eff = topk < S ? topk : S;
    unsigned long long prev = ~0ull;
    float sum = 0.f;
    int32_t* pp = pool_pos ? pool_pos + ((int64_t)slide * C + c) * topk : nullptr;
    for (int r = 0; r < k_eff; ++r) {
        unsigned long long best = 0ull;
        for (int i = lane; i < S; i += 32) {
            const unsigned long long key =
                ((unsigned long long)f2ord(f[(int64_t)i * C]) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
            if (key < prev && key > best) best = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, best, o);
            best = other > best ? other : best;
        }
        prev = best;
        const uint32_t pos = 0xffffffffu - (uint32_t)(best & 0xffffffffull);
        sum += ord2f((uint32_t)(best >> 32));
        if (pp && lane == 0) pp[r] = (int32_t)pos;
    }
    if (lane == 0) {
        bag_logits[(int64_t)slide * C + c] = k_eff > 0 ? sum / (float)k_eff : 0.f;
        if (pp)
            for (int r = k_eff; r < topk; ++r) pp[r] = -1;
    }
}

// Block per slide, one coalesced pass over the slide's [S][C] block of combined scores (the version above reads it
// K times with a stride of C floats: 30 x the sectors at C = 30).  The block uses T = C * floor(512 / C) threads, so a
// thread striding the flat array by T always meets the same class; it keeps its POOL_MAXK best 64-bit keys
// (value, lowest position first) sorted in registers, loading 8 elements at a time so that 8 x T loads are in flight.
// Then K rounds: every thread offers the head of its list to a shared-memory atomicMax per class, the owner of the
// winning key pops it.  Same keys and the same order as pool_final_kernel, so results and pool_pos are identical.
constexpr int POOL_MAXK = 16;
constexpr int POOL_T = 512;
constexpr int POOL_BATCH = 8;
__global__ void __launch_bounds__(POOL_T)
pool_final_block_kernel(const float* __restrict__ final_scores, const int64_t* __restrict__ sel_base,
                        const int32_t* __restrict__ sel_count, int C, int topk, float* __restrict__ bag_logits,
                        int32_t* __restrict__ pool_pos) {
    extern __shared__ unsigned long long pool_best[];   // [C] round winners, then float sums [C]
    const int slide = blockIdx.x, tid = threadIdx.x;
    const int per = POOL_T / C, T = per * C;            // threads in use; thread t owns class t % C
    const int S = sel_count[slide];
    const float* f = final_scores + sel_base[slide] * C;
    const int64_t total = (int64_t)S * C;
    float* sums = reinterpret_cast<float*>(pool_best + C);
    unsigned long long* thr = pool_best + 2 * C;          // [C] admission thresholds
    unsigned long long* tmax = thr + C;                   // [T] per-thread maxima
    const int k_eff = topk < S ? topk : S;
    unsigned long long lst[POOL_MAXK];
#pragma unroll
    for (int j = 0; j < POOL_MAXK; ++j) lst[j] = 0ull;
    if (tid < C) { pool_best[tid] = 0ull; sums[tid] = 0.f; }
    // Pass 1: every thread's largest key.  The K-th largest of the per threads' maxima of a class is a lower bound of
    // the class's K-th largest key (K keys at least that large exist), so pass 2 only has to look at keys above it:
    // a dozen per class instead of a sorted-list insertion at almost every element (with 32 lanes sharing a branch,
    // "some lane inserts" was true for ~95 % of the elements: 0.32 ms per 100 C=30 slides).
    unsigned long long mx = 0ull;
    if (tid < T) {
        uint32_t i = (uint32_t)(tid / C);               // row of element e = tid + m * T is tid / C + m * per
        for (int64_t e = tid; e < total; e += (int64_t)POOL_BATCH * T, i += (uint32_t)(POOL_BATCH * per)) {
            float v[POOL_BATCH];
#pragma unroll
            for (int u = 0; u < POOL_BATCH; ++u) v[u] = e + (int64_t)u * T < total ? f[e + (int64_t)u * T] : 0.f;
#pragma unroll
            for (int u = 0; u < POOL_BATCH; ++u) {
                if (e + (int64_t)u * T < total) {
                    const unsigned long long key = ((unsigned long long)f2ord(v[u]) << 32) |
                                                   (unsigned long long)(0xffffffffu - (i + (uint32_t)(u * per)));
                    mx = key > mx ? key : mx;
                }
            }
        }
        tmax[tid] = mx;
    }
    __syncthreads();
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int cc = warp; cc < C; cc += POOL_T / 32) {  // warp per class: K rounds of "largest below the previous"
            unsigned long long prev = ~0ull;
            for (int r = 0; r < k_eff; ++r) {
                unsigned long long best = 0ull;
                for (int j = lane; j < per; j += 32) {
                    const unsigned long long key = tmax[cc + j * C];
                    if (key < prev && key > best) best = key;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(FULL, best, o);
                    best = other > best ? other : best;
                }
                prev = best;                              // 0 once the maxima run out: no filtering
            }
            if (lane == 0) thr[cc] = k_eff > 0 ? prev : 0ull;
        }
    }
    __syncthreads();
    // Pass 2: sorted top lists of the keys at or above the class threshold
    if (tid < T) {
        const unsigned long long lim = thr[tid % C];
        uint32_t i = (uint32_t)(tid / C);
        for (int64_t e = tid; e < total; e += (int64_t)POOL_BATCH * T, i += (uint32_t)(POOL_BATCH * per)) {
            float v[POOL_BATCH];
#pragma unroll
            for (int u = 0; u < POOL_BATCH; ++u) v[u] = e + (int64_t)u * T < total ? f[e + (int64_t)u * T] : 0.f;
#pragma unroll
            for (int u = 0; u < POOL_BATCH; ++u) {
                if (e + (int64_t)u * T < total) {
                    const unsigned long long key = ((unsigned long long)f2ord(v[u]) << 32) |
                                                   (unsigned long long)(0xffffffffu - (i + (uint32_t)(u * per)));
                    if (key >= lim && key > lst[POOL_MAXK - 1]) {
                        lst[POOL_MAXK - 1] = key;
#pragma unroll
                        for (int j = POOL_MAXK - 1; j > 0; --j) {
                            const unsigned long long a = lst[j - 1], b = lst[j];
                            lst[j - 1] = a > b ? a : b;
                            lst[j] = a > b ? b : a;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    const int c = tid % C;
    for (int r = 0; r < k_eff; ++r) {
        if (tid < T && lst[0] != 0ull) atomicMax(&pool_best[c], lst[0]);
        __syncthreads();
        const unsigned long long win = pool_best[c];
        const bool mine = tid < T && lst[0] == win && win != 0ull;   // keys are unique: exactly one owner per class
        __syncthreads();
        if (mine) {
#pragma unroll
            for (int j = 0; j + 1 < POOL_MAXK; ++j) lst[j] = lst[j + 1];
            lst[POOL_MAXK - 1] = 0ull;
            sums[c] += ord2f((uint32_t)(win >> 32));      // one writer per class and round: the summation order is fixed
            if (pool_pos) pool_pos[((int64_t)slide * C + c) * topk + r] = (int32_t)(0xffffffffu - (uint32_t)(win & 0xffffffffull));
            pool_best[c] = 0ull;
        }
        __syncthreads();
    }
    if (tid < C) {
        bag_logits[(int64_t)slide * C + tid] = k_eff > 0 ? sums[tid] / (float)k_eff : 0.f;
        if (pool_pos)
            for (int r = k_eff; r < topk; ++r) pool_pos[((int64_t)slide * C + tid) * topk + r] = -1;
    }
}

static int launch_pool_final(const float* final_scores, const int64_t* sel_base, const int32_t* sel_count, int n_slides,
                             int C, int topk, float* bag_logits, int32_t* pool_pos, cudaStream_t st) {
    if (topk <= POOL_MAXK && C <= POOL_T) {
        // round winners [C] (8 B) | sums [C] (4 B, padded to 8) | thresholds [C] | per-thread maxima [POOL_T]
        const size_t smem = (size_t)(3 * C + POOL_T) * sizeof(unsigned long long);
        pool_final_block_kernel<<<n_slides, POOL_T, smem, st>>>(final_scores, sel_base, sel_count, C, topk, bag_logits, pool_pos);
        MOC_LAUNCH_CHECK("pool_final_block_kernel");
        return MOC_OK;
    }
    const int64_t items = (int64_t)n_slides * C;
    pool_final_kernel<<<(unsigned)((items + POOL_WARPS - 1) / POOL_WARPS), POOL_WARPS * 32, 0, st>>>(
        final_scores, sel_base, sel_count, n_slides, C, topk, bag_logits, pool_pos);
    MOC_LAUNCH_CHECK("pool_final_kernel");
    return MOC_OK;
}

// ------------------------------------------------------------------------------------------------
// cross_entropy(logits[1,C], label) per slide; optional gradient and argmax.
// ------------------------------------------------------------------------------------------------
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int n, int C,
                                     float grad_scale, float* __restrict__ loss, float* __restrict__ dlogits,
                                     int32_t* __restrict__ pred) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = logits + (int64_t)i * C;
    float m = x[0];
    int am = 0;
    for (int c = 1; c < C; ++c)
        if (x[c] > m) { m = x[c]; am = c; }
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(x[c] - m);
    const float lse = logf(s);
    const int y = (int)labels[i];
    if (loss) loss[i] = -((x[y] - m) - lse);
    if (pred) pred[i] = am;
    if (dlogits) {
        for (int c = 0; c < C; ++c) {
            const float p = expf((x[c] - m) - lse);
            dlogits[(int64_t)i * C + c] = grad_scale * (p - (c == y ? 1.f : 0.f));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  Every pooled (slide, class, rank) entry is one "virtual row" carrying gradient
// dlogits[slide,c]/k_eff in column c only; gradients are linear in it, so a row pooled for several classes
// is simply handled once per class.  Stage 1 recomputes the gate of each virtual row and writes
// dz1[64], dz2[4], h[64]; stage 2 reduces the outer products over virtual rows in a fixed order.
// ------------------------------------------------------------------------------------------------
constexpr int BW_THREADS = 256;
struct BwdScratch {  // per virtual row
    float dz1[H];
    float h[H];
    float dz2[G];
    int32_t row;
    int32_t pad[3];
};

__global__ void __launch_bounds__(BW_THREADS)
head_bwd_rows_kernel(const float* __restrict__ feat, const float* __restrict__ keys, int64_t key_stride, int C,
                     const int64_t* __restrict__ sel_base, const int32_t* __restrict__ sel_rows,
                     const int32_t* __restrict__ sel_count, const float* __restrict__ w1,
                     const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     unsigned active_mask, int topk, const int32_t* __restrict__ pool_pos,
                     const float* __restrict__ dlogits, const float* __restrict__ dgate_dense,
                     BwdScratch* __restrict__ scratch) {
    __shared__ float z1s[H], dz2s[G];
    const int v = blockIdx.x;  // pooled mode: virtual row = (slide*C + c)*topk + r; dense mode: row of feat
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    BwdScratch& out = scratch[v];
    int32_t row;
    int c = 0, slide = 0;
    if (dgate_dense != nullptr) {
        const float4 dg = *reinterpret_cast<const float4*>(dgate_dense + (int64_t)v * G);
        if (dg.x == 0.f && dg.y == 0.f && dg.z == 0.f && dg.w == 0.f) {
            if (tid == 0) out.row = -1;
            return;
        }
        row = v;
    } else {
        c = (v / topk) % C;
        slide = v / (topk * C);
        const int32_t pos = pool_pos[v];
        if (pos < 0) {
            if (tid == 0) out.row = -1;
            return;
        }
        row = sel_rows[sel_base[slide] + pos];
    }
    const float4* xp = reinterpret_cast<const float4*>(feat + (int64_t)row * D);
    float4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = __ldg(xp + q * 32 + lane);
    // z1: warp w computes hidden units w*8 .. w*8+7
#pragma unroll
    for (int jj = 0; jj < H / (BW_THREADS / 32); ++jj) {
        const int j = warp * (H / (BW_THREADS / 32)) + jj;
        const float4* wp = reinterpret_cast<const float4*>(w1 + (int64_t)j * D);
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 wv = __ldg(wp + q * 32 + lane);
            a = fmaf(x[q].x, wv.x, a);
            a = fmaf(x[q].y, wv.y, a);
            a = fmaf(x[q].z, wv.z, a);
            a = fmaf(x[q].w, wv.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) z1s[j] = a + b1[j];
    }
    __syncthreads();
    if (warp < G) {
        const int m = warp;
        float a = fmaxf(z1s[lane], 0.f) * w2[m * H + lane] + fmaxf(z1s[lane + 32], 0.f) * w2[m * H + lane + 32];
        a = warp_sum(a);
        if (lane == 0) {
            const float g = sigmoidf_exact(a + b2[m]);
            float dg;
            if (dgate_dense != nullptr) {
                dg = dgate_dense[(int64_t)v * G + m];
            } else {
                const int S = sel_count[slide];
                const int k_eff = topk < S ? topk : S;
                const float dF = dlogits[(int64_t)slide * C + c] / (float)k_eff;
                const float* kp = keys + row;
                const KeyLayout kl = key_layout(C);
                float psi;
                if (m == 0) psi = kp[(int64_t)c * key_stride];
                else if (m == 1) psi = kl.compact ? row_softmax_of(kp, key_stride, kl).of(kp[(int64_t)c * key_stride])
                                                  : kp[(int64_t)(C + c) * key_stride];
                else if (m == 2) psi = kp[(int64_t)kl.diff * key_stride];
                else psi = kp[(int64_t)kl.bg_max * key_stride];
                dg = ((active_mask >> m) & 1u) ? dF * psi : 0.f;
            }
            dz2s[m] = dg * g * (1.f - g);
        }
    }
    __syncthreads();
    if (tid < H) {
        const float z = z1s[tid];
        float dh = 0.f;
#pragma unroll
        for (int m = 0; m < G; ++m) dh = fmaf(dz2s[m], w2[m * H + tid], dh);
        out.dz1[tid] = z > 0.f ? dh : 0.f;
        out.h[tid] = fmaxf(z, 0.f);
    }
    if (tid < G) out.dz2[tid] = dz2s[tid];
    if (tid == 0) out.row = row;
}

// grid H+1: block j < 64 -> dW1[j][:] and db1[j]; block 64 -> dW2, db2.
__global__ void __launch_bounds__(128)
head_bwd_reduce_kernel(const float* __restrict__ feat, const BwdScratch* __restrict__ scratch, int n_virtual,
                       float* __restrict__ grads) {
    const int tid = threadIdx.x;
    float* dw1 = grads;
    float* db1 = grads + H * D;
    float* dw2 = db1 + H;
    float* db2 = dw2 + G * H;
    if (blockIdx.x < H) {
        const int j = blockIdx.x;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float bsum = 0.f;
        for (int v = 0; v < n_virtual; ++v) {
            const int32_t row = scratch[v].row;
            if (row < 0) continue;
            const float d = scratch[v].dz1[j];
            const float4 xv = __ldg(reinterpret_cast<const float4*>(feat + (int64_t)row * D) + tid);
            acc.x = fmaf(d, xv.x, acc.x);
            acc.y = fmaf(d, xv.y, acc.y);
            acc.z = fmaf(d, xv.z, acc.z);
            acc.w = fmaf(d, xv.w, acc.w);
            bsum += d;
        }
        reinterpret_cast<float4*>(dw1 + (int64_t)j * D)[tid] = acc;
        if (tid == 0) db1[j] = bsum;
    } else {
        for (int o = tid; o < G * H; o += blockDim.x) {
            const int m = o / H, j = o % H;
            float a = 0.f;
            for (int v = 0; v < n_virtual; ++v)
                if (scratch[v].row >= 0) a = fmaf(scratch[v].dz2[m], scratch[v].h[j], a);
            dw2[o] = a;
        }
        if (tid < G) {
            float a = 0.f;
            for (int v = 0; v < n_virtual; ++v)
                if (scratch[v].row >= 0) a += scratch[v].dz2[tid];
            db2[tid] = a;
        }
    }
}


// ablation_evaluation (main_moc.py:538-553): un-gated avg / sum / max of the four planes per selected row.
// mode 0 = avg (0.25 each), 1 = sum, 2 = max.  One thread per (slot, class).
__global__ void ablation_rows_kernel(const float* __restrict__ keys, int64_t key_stride, int C,
                                     const int32_t* __restrict__ sel_rows, int64_t n_slots, int mode,
                                     float* __restrict__ final_scores) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_slots * C) return;
    const int64_t slot = t / C;
    const int c = (int)(t % C);
    const int32_t row = sel_rows[slot];
    if (row < 0) return;
    const float* kp = keys + row;
    const KeyLayout kl = key_layout(C);
    const float l = kp[(int64_t)c * key_stride];
    const float p = kl.compact ? row_softmax_of(kp, key_stride, kl).of(l) : kp[(int64_t)(C + c) * key_stride];
    const float d = kp[(int64_t)kl.diff * key_stride], b = kp[(int64_t)kl.bg_max * key_stride];
    float f;
    if (mode == 2) {
        f = fmaxf(fmaxf(l, p), fmaxf(d, b));
    } else {
        const float g = mode == 0 ? 0.25f : 1.0f;
        f = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g, l), __fmul_rn(g, p)), __fmul_rn(g, d)), __fmul_rn(g, b));
    }
    final_scores[t] = f;
}

// selected_feat = feat[selected_index] and the four [S,C] score planes (main_moc.py:355-366) for callers that
// want the reference's dense tensors.  One warp per selected row.
__global__ void __launch_bounds__(256)
gather_selected_kernel(const float* __restrict__ feat, const float* __restrict__ keys, int64_t key_stride, int C,
                       const int32_t* __restrict__ sel_rows, int64_t n_sel, float* __restrict__ out_feat,
                       float* __restrict__ p_top, float* __restrict__ p_softmax, float* __restrict__ p_diff,
                       float* __restrict__ p_bg) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= n_sel) return;
    const int32_t row = sel_rows[s];
    if (row < 0) return;
    if (out_feat) {
        const float4* src = reinterpret_cast<const float4*>(feat + (int64_t)row * D);
        float4* dst = reinterpret_cast<float4*>(out_feat + s * D);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q * 32 + lane] = __ldg(src + q * 32 + lane);
    }
    if (p_top) {
        const float* kp = keys + row;
        const KeyLayout kl = key_layout(C);
        const float dlt = kp[(int64_t)kl.diff * key_stride], bgm = kp[(int64_t)kl.bg_max * key_stride];
        RowSoftmax rs = {0.f};
        if (kl.compact) rs = row_softmax_of(kp, key_stride, kl);
        for (int c = lane; c < C; c += 32) {
            const float lt = kp[(int64_t)c * key_stride];
            p_top[s * C + c] = lt;
            p_softmax[s * C + c] = kl.compact ? rs.of(lt) : kp[(int64_t)(C + c) * key_stride];
            p_diff[s * C + c] = dlt;
            p_bg[s * C + c] = bgm;
        }
    }
}

// torch.optim.Adam, single tensor, weight decay folded into the gradient (torch/optim/adam.py _single_tensor_adam)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                            float wd, float step_size, float bc2_sqrt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pi = p[i];
    const float gi = __fmaf_rn(wd, pi, g[i]);
    // exp_avg.lerp_(grad, 1-beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float mi = __fmaf_rn(1.f - beta1, gi - m[i], m[i]);
    const float vi = __fmaf_rn(__fmul_rn(gi, gi), 1.f - beta2, __fmul_rn(v[i], beta2));
    m[i] = mi;
    v[i] = vi;
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
    p[i] = pi - step_size * __fdiv_rn(mi, denom);
}

// Adam with the step counter in device memory, so that a captured CUDA graph of a training step can be replayed:
// the prepare kernel advances the counter and derives the two bias-correction scalars (same double-precision
// arithmetic as moc_adam_step does on the host), the apply kernel is adam_kernel reading them from memory.
__global__ void adam_prepare_dev_kernel(int64_t* __restrict__ step, float* __restrict__ scalars, float lr, float beta1,
                                        float beta2) {
    const int64_t t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    scalars[0] = (float)((double)lr / bc1);
    scalars[1] = (float)sqrt(bc2);
}
__global__ void adam_apply_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                      float* __restrict__ v, int64_t n, const float* __restrict__ scalars, float beta1,
                                      float beta2, float eps, float wd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float step_size = scalars[0], bc2_sqrt = scalars[1];
    const float pi = p[i];
    const float gi = __fmaf_rn(wd, pi, g[i]);
    const float mi = __fmaf_rn(1.f - beta1, gi - m[i], m[i]);
    const float vi = __fmaf_rn(__fmul_rn(gi, gi), 1.f - beta2, __fmul_rn(v[i], beta2));
    m[i] = mi;
    v[i] = vi;
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
    p[i] = pi - step_size * __fdiv_rn(mi, denom);
}

// dst += src (gradient accumulation over the slides of a data-parallel micro-batch)
__global__ void accumulate_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __fadd_rn(dst[i], src[i]);
}

}  // namespace moc

using namespace moc;

extern "C" int moc_adam_prepare_dev(int64_t* step, float* scalars, float lr, float beta1, float beta2, void* stream) {
    MOC_CHECK_ARG(step && scalars, "moc_adam_prepare_dev: null pointer");
    adam_prepare_dev_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, scalars, lr, beta1, beta2);
    MOC_LAUNCH_CHECK("adam_prepare_dev_kernel");
    return MOC_OK;
}

extern "C" int moc_adam_apply_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                  const float* scalars, float beta1, float beta2, float eps, float weight_decay,
                                  void* stream) {
    MOC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && scalars && n >= 0, "moc_adam_apply_dev: bad arguments");
    if (n == 0) return MOC_OK;
    adam_apply_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        params, grads, exp_avg, exp_avg_sq, n, scalars, beta1, beta2, eps, weight_decay);
    MOC_LAUNCH_CHECK("adam_apply_dev_kernel");
    return MOC_OK;
}

extern "C" int moc_accumulate(float* dst, const float* src, int64_t n, void* stream) {
    MOC_CHECK_ARG(dst && src && n >= 0, "moc_accumulate: bad arguments");
    if (n == 0) return MOC_OK;
    accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    MOC_LAUNCH_CHECK("accumulate_kernel");
    return MOC_OK;
}

extern "C" int moc_head_forward(const float* feat, const float* keys, int64_t key_stride, int n_classes,
                                const int64_t* sel_base, const int32_t* sel_rows, const int32_t* sel_count,
                                int n_slides, int64_t sel_capacity_total, const float* w1, const float* b1,
                                const float* w2, const float* b2, unsigned active_mask, int topk, float* gate,
                                float* final_scores, float* bag_logits, int32_t* pool_pos, void* workspace,
                                size_t workspace_bytes, void* stream) {
    MOC_CHECK_ARG(feat && keys && sel_base && sel_rows && sel_count && w1 && b1 && w2 && b2 && final_scores &&
                      bag_logits,
                  "moc_head_forward: null pointer");
    MOC_CHECK_ARG(n_slides >= 0 && sel_capacity_total >= 0 && topk >= 1, "moc_head_forward: bad sizes");
    MOC_CHECK_SHAPE(n_classes >= 2 && n_classes < MOC_MAX_COLS, "moc_head_forward: bad class count %d", n_classes);
    if (n_slides == 0) return MOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (sel_capacity_total > 0) {
        if (use_simt_head()) {
            const size_t smem = (size_t)(H + HR_TM) * HR_LD * sizeof(float);
            MOC_CUDA(cudaFuncSetAttribute(head_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int64_t n_tiles = (sel_capacity_total + HR_TM - 1) / HR_TM;
            const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
            head_rows_kernel<<<grid, HR_THREADS, smem, st>>>(feat, keys, key_stride, n_classes, sel_rows,
                                                             sel_capacity_total, w1, b1, w2, b2, active_mask, gate,
                                                             final_scores);
            MOC_LAUNCH_CHECK("head_rows_kernel");
        } else {
            if (workspace == nullptr || workspace_bytes < head_ws_bytes()) {
                set_error("moc_head_forward: workspace %zu B < required %zu B", workspace_bytes, head_ws_bytes());
                return MOC_E_WORKSPACE;
            }
            const int rc = launch_head_rows_mma(feat, keys, key_stride, n_classes, sel_rows, sel_capacity_total, w1, b1,
                                               w2, b2, active_mask, gate, final_scores, workspace, st);
            if (rc != MOC_OK) return rc;
        }
    }
    {
        const int rc = launch_pool_final(final_scores, sel_base, sel_count, n_slides, n_classes, topk, bag_logits, pool_pos, st);
        if (rc != MOC_OK) return rc;
    }
    return MOC_OK;
}

extern "C" int moc_cross_entropy(const float* bag_logits, const int64_t* labels, int n_slides, int n_classes,
                                 float grad_scale, float* loss, float* dlogits, int32_t* pred, void* stream) {
    MOC_CHECK_ARG(bag_logits && labels, "moc_cross_entropy: null pointer");
    MOC_CHECK_ARG(n_slides >= 0 && n_classes >= 1, "moc_cross_entropy: bad sizes");
    if (n_slides == 0) return MOC_OK;
    cross_entropy_kernel<<<(n_slides + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bag_logits, labels, n_slides,
                                                                                  n_classes, grad_scale, loss, dlogits,
                                                                                  pred);
    MOC_LAUNCH_CHECK("cross_entropy_kernel");
    return MOC_OK;
}

extern "C" size_t moc_head_backward_workspace_bytes(int n_slides, int n_classes, int topk) {
    return (size_t)n_slides * n_classes * topk * sizeof(BwdScratch);
}

extern "C" int moc_head_backward(const float* feat, const float* keys, int64_t key_stride, int n_classes,
                                 const int64_t* sel_base, const int32_t* sel_rows, const int32_t* sel_count,
                                 int n_slides, const float* w1, const float* b1, const float* w2, const float* b2,
                                 unsigned active_mask, int topk, const int32_t* pool_pos, const float* dlogits,
                                 float* grads, void* workspace, size_t workspace_bytes, void* stream) {
    MOC_CHECK_ARG(feat && keys && sel_base && sel_rows && sel_count && w1 && b1 && w2 && b2 && pool_pos && dlogits &&
                      grads && workspace,
                  "moc_head_backward: null pointer");
    MOC_CHECK_ARG(n_slides >= 1 && topk >= 1, "moc_head_backward: bad sizes");
    MOC_CHECK_SHAPE(n_classes >= 2 && n_classes < MOC_MAX_COLS, "moc_head_backward: bad class count %d", n_classes);
    const size_t need = moc_head_backward_workspace_bytes(n_slides, n_classes, topk);
    if (workspace_bytes < need) {
        set_error("moc_head_backward: workspace %zu B < required %zu B", workspace_bytes, need);
        return MOC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_virtual = n_slides * n_classes * topk;
    BwdScratch* scratch = reinterpret_cast<BwdScratch*>(workspace);
    head_bwd_rows_kernel<<<n_virtual, BW_THREADS, 0, st>>>(feat, keys, key_stride, n_classes, sel_base, sel_rows,
                                                          sel_count, w1, b1, w2, b2, active_mask, topk, pool_pos,
                                                          dlogits, nullptr, scratch);
    MOC_LAUNCH_CHECK("head_bwd_rows_kernel");
    head_bwd_reduce_kernel<<<H + 1, 128, 0, st>>>(feat, scratch, n_virtual, grads);
    MOC_LAUNCH_CHECK("head_bwd_reduce_kernel");
    return MOC_OK;
}

extern "C" int moc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                             int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                             void* stream) {
    MOC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "moc_adam_step: null pointer");
    MOC_CHECK_ARG(n >= 0 && step >= 1, "moc_adam_step: bad n / step");
    if (n == 0) return MOC_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)((double)lr / bc1),
        (float)sqrt(bc2));
    MOC_LAUNCH_CHECK("adam_kernel");
    return MOC_OK;
}

extern "C" int moc_gather_selected(const float* feat, const float* keys, int64_t key_stride, int n_classes,
                                   const int32_t* sel_rows, int64_t n_sel, float* out_feat, float* plane_top,
                                   float* plane_softmax, float* plane_diff, float* plane_bg, void* stream) {
    MOC_CHECK_ARG(feat && sel_rows && n_sel >= 0, "moc_gather_selected: bad arguments");
    MOC_CHECK_ARG(!plane_top || (keys && plane_softmax && plane_diff && plane_bg),
                  "moc_gather_selected: the four planes come together");
    if (n_sel == 0) return MOC_OK;
    gather_selected_kernel<<<(unsigned)((n_sel + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        feat, keys, key_stride, n_classes, sel_rows, n_sel, out_feat, plane_top, plane_softmax, plane_diff, plane_bg);
    MOC_LAUNCH_CHECK("gather_selected_kernel");
    return MOC_OK;
}

extern "C" size_t moc_head_forward_workspace_bytes(void) { return head_ws_bytes(); }
extern "C" size_t moc_head_domain_flag_offset(void) { return head_img_bytes(); }

extern "C" int moc_senet_forward(const float* x, int64_t n_rows, const float* w1, const float* b1, const float* w2,
                                 const float* b2, float* gate, unsigned flags, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    MOC_CHECK_ARG(x && w1 && b1 && w2 && b2 && gate && n_rows >= 0, "moc_senet_forward: bad arguments");
    MOC_CHECK_SHAPE(n_rows < (1ll << 31), "moc_senet_forward: too many rows");
    if (n_rows == 0) return MOC_OK;
    if (!use_simt_head()) {
        if (workspace == nullptr || workspace_bytes < head_ws_bytes()) {
            set_error("moc_senet_forward: workspace %zu B < required %zu B", workspace_bytes, head_ws_bytes());
            return MOC_E_WORKSPACE;
        }
        return launch_head_rows_mma(x, nullptr, 0, 0, nullptr, n_rows, w1, b1, w2, b2, flags & MOC_HEAD_WIDE_DOMAIN, gate, nullptr, workspace,
                                   (cudaStream_t)stream);
    }
    const size_t smem = (size_t)(H + HR_TM) * HR_LD * sizeof(float);
    MOC_CUDA(cudaFuncSetAttribute(head_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tiles = (n_rows + HR_TM - 1) / HR_TM;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    head_rows_kernel<<<grid, HR_THREADS, smem, (cudaStream_t)stream>>>(x, nullptr, 0, 0, nullptr, n_rows, w1, b1, w2, b2,
                                                                      0u, gate, nullptr);
    MOC_LAUNCH_CHECK("head_rows_kernel");
    return MOC_OK;
}

extern "C" size_t moc_senet_backward_workspace_bytes(int64_t n_rows) { return (size_t)n_rows * sizeof(BwdScratch); }

extern "C" int moc_senet_backward(const float* x, int64_t n_rows, const float* dgate, const float* w1, const float* b1,
                                  const float* w2, const float* b2, float* grads, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    MOC_CHECK_ARG(x && dgate && w1 && b1 && w2 && b2 && grads && workspace && n_rows >= 1,
                  "moc_senet_backward: bad arguments");
    MOC_CHECK_SHAPE(n_rows < (1ll << 30), "moc_senet_backward: too many rows");
    if (workspace_bytes < moc_senet_backward_workspace_bytes(n_rows)) {
        set_error("moc_senet_backward: workspace too small");
        return MOC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    BwdScratch* scratch = reinterpret_cast<BwdScratch*>(workspace);
    head_bwd_rows_kernel<<<(unsigned)n_rows, BW_THREADS, 0, st>>>(x, nullptr, 0, 0, nullptr, nullptr, nullptr, w1, b1, w2,
                                                                 b2, 0u, 1, nullptr, nullptr, dgate, scratch);
    MOC_LAUNCH_CHECK("head_bwd_rows_kernel");
    head_bwd_reduce_kernel<<<H + 1, 128, 0, st>>>(x, scratch, (int)n_rows, grads);
    MOC_LAUNCH_CHECK("head_bwd_reduce_kernel");
    return MOC_OK;
}

extern "C" int moc_ablation_forward(const float* keys, int64_t key_stride, int n_classes, const int64_t* sel_base,
                                    const int32_t* sel_rows, const int32_t* sel_count, int n_slides,
                                    int64_t sel_capacity_total, int mode, int topk, float* final_scores,
                                    float* bag_logits, void* stream) {
    MOC_CHECK_ARG(keys && sel_base && sel_rows && sel_count && final_scores && bag_logits,
                  "moc_ablation_forward: null pointer");
    MOC_CHECK_ARG(mode >= 0 && mode <= 2 && topk >= 1 && n_slides >= 0, "moc_ablation_forward: bad arguments");
    if (n_slides == 0) return MOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t items = sel_capacity_total * n_classes;
    if (items > 0) {
        ablation_rows_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(keys, key_stride, n_classes, sel_rows,
                                                                            sel_capacity_total, mode, final_scores);
        MOC_LAUNCH_CHECK("ablation_rows_kernel");
    }
    return launch_pool_final(final_scores, sel_base, sel_count, n_slides, n_classes, topk, bag_logits, nullptr, st);
}
