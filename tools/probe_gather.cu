// Dev probe: bandwidth of gathering sparse 2 KB rows (the head's selected patches) by different mechanisms.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// each warp: ring of SLOTS slots, a slot = ROWS_PER_SLOT rows, each row fetched as PIECES copies of 2048/PIECES bytes
template <int WARPS, int SLOTS, int ROWS_PER_SLOT, int PIECES>
__global__ void __launch_bounds__(WARPS * 32, 1) gather_bulk(const char* __restrict__ feat, const int* __restrict__ rows, int64_t n_sel, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int SLOT_BYTES = ROWS_PER_SLOT * 2048;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* st0 = smem + (size_t)warp * SLOTS * SLOT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * SLOTS * SLOT_BYTES) + warp * SLOTS;
    if (lane == 0) {
        for (int s = 0; s < SLOTS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t n_groups = n_sel / ROWS_PER_SLOT;
    const int64_t stride = (int64_t)gridDim.x * WARPS;
    int64_t g = (int64_t)blockIdx.x * WARPS + warp;
    auto issue = [&](int64_t gg, int s) {
        if (lane == 0) mbar_expect(&bars[s], SLOT_BYTES);
        __syncwarp();
        for (int c = lane; c < ROWS_PER_SLOT * PIECES; c += 32) {
            const int r = c / PIECES, p = c % PIECES;
            const int64_t row = rows[gg * ROWS_PER_SLOT + r];
            bulk(st0 + s * SLOT_BYTES + r * 2048 + p * (2048 / PIECES), feat + row * 2048 + p * (2048 / PIECES), 2048 / PIECES, &bars[s]);
        }
    };
    for (int s = 0; s < SLOTS; ++s) if (g + s * stride < n_groups) issue(g + s * stride, s);
    int stage = 0; uint32_t par = 0; float acc = 0.f;
    for (; g < n_groups; g += stride) {
        mbar_wait(&bars[stage], par);
        const float4* p = reinterpret_cast<const float4*>(st0 + stage * SLOT_BYTES);
        for (int i = lane; i < SLOT_BYTES / 16; i += 32) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        __syncwarp();
        const int64_t gn = g + (int64_t)SLOTS * stride;
        if (gn < n_groups) issue(gn, stage);
        if (++stage == SLOTS) { stage = 0; par ^= 1; }
    }
    if (acc == 123.456f) out[0] = acc;
}

// LDG: a warp reads a row with 4 x LDG.128 per lane; UNROLL rows in flight per warp
template <int UNROLL>
__global__ void __launch_bounds__(256) gather_ldg(const char* __restrict__ feat, const int* __restrict__ rows, int64_t n_sel, float* out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int64_t g = warp * UNROLL; g + UNROLL <= n_sel; g += nw * UNROLL) {
        float4 v[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const float4* p = reinterpret_cast<const float4*>(feat + (int64_t)rows[g + u] * 2048) + lane;
#pragma unroll
            for (int q = 0; q < 4; ++q) v[u][q] = __ldg(p + q * 32);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += v[u][q].x + v[u][q].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
void timeit(const char* name, int64_t bytes, F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    printf("%-44s %.3f ms  %.0f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int W, int S, int R, int P>
void run_bulk(const char* name, const char* feat, const int* rows, int64_t n_sel, float* out, int sms) {
    size_t smem = (size_t)W * S * R * 2048 + W * S * 8;
    cudaFuncSetAttribute(gather_bulk<W, S, R, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    timeit(name, n_sel * 2048, [&] { gather_bulk<W, S, R, P><<<sms, W * 32, smem>>>(feat, rows, n_sel, out); });
}

int main() {
    const int64_t n_rows = 8ll << 20;  // 16 GiB of rows
    const double density = 0.083;
    char* feat; float* out; cudaMalloc(&feat, n_rows * 2048); cudaMalloc(&out, 4); cudaMemset(feat, 1, n_rows * 2048);
    std::vector<int> sel; srand(1);
    for (int64_t r = 0; r < n_rows; ++r) if (rand() < density * RAND_MAX) sel.push_back((int)r);
    int64_t n_sel = sel.size() / 64 * 64;
    int* rows; cudaMalloc(&rows, n_sel * 4); cudaMemcpy(rows, sel.data(), n_sel * 4, cudaMemcpyHostToDevice);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("%lld selected rows of %lld (%.1f MB gathered)\n", (long long)n_sel, (long long)n_rows, n_sel * 2048 / 1e6);
    run_bulk<8, 4, 2, 1>("bulk 2KB rows: 8w x 4 slots x 2 rows", feat, rows, n_sel, out, sms);
    run_bulk<8, 8, 1, 1>("bulk 2KB rows: 8w x 8 slots x 1 row", feat, rows, n_sel, out, sms);
    run_bulk<8, 3, 4, 1>("bulk 2KB rows: 8w x 3 slots x 4 rows", feat, rows, n_sel, out, sms);
    run_bulk<16, 3, 2, 1>("bulk 2KB rows: 16w x 3 slots x 2 rows", feat, rows, n_sel, out, sms);
    run_bulk<8, 4, 2, 8>("bulk 256B pieces: 8w x 4 slots x 2 rows", feat, rows, n_sel, out, sms);
    run_bulk<8, 4, 2, 16>("bulk 128B pieces: 8w x 4 slots x 2 rows", feat, rows, n_sel, out, sms);
    run_bulk<8, 2, 4, 16>("bulk 128B pieces: 8w x 2 slots x 4 rows", feat, rows, n_sel, out, sms);
    timeit("ldg: 1 row / warp in flight, 148x8 CTAs", n_sel * 2048, [&] { gather_ldg<1><<<sms * 8, 256>>>(feat, rows, n_sel, out); });
    timeit("ldg: 2 rows / warp in flight, 148x8 CTAs", n_sel * 2048, [&] { gather_ldg<2><<<sms * 8, 256>>>(feat, rows, n_sel, out); });
    timeit("ldg: 4 rows / warp in flight, 148x8 CTAs", n_sel * 2048, [&] { gather_ldg<4><<<sms * 8, 256>>>(feat, rows, n_sel, out); });
    timeit("ldg: 4 rows / warp in flight, 148x2 CTAs", n_sel * 2048, [&] { gather_ldg<4><<<sms * 2, 256>>>(feat, rows, n_sel, out); });
    return 0;
}
