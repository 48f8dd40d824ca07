// Dev probe: how fast can a CTA gather sparse 2 KB rows when - like the gate kernel's producers - it visits a tile of
// 128 rows one K-block at a time, i.e. a row's 2 KB arrive as 2048/PIECE separate requests spread over the tile's
// lifetime?  PIECE = 256 B is what head_rows_f16_kernel does (64 floats per K-block); 512 / 1024 / 2048 B show what a
// wider K-block would buy.  512 threads per CTA, 1 CTA per SM, two steps of loads in flight per thread.
//   nvcc -O3 -arch=sm_100a -o tools/probe_gather_kblock tools/probe_gather_kblock.cu && tools/probe_gather_kblock
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

template <int PIECE, int THREADS, int DEPTH>
__global__ void __launch_bounds__(THREADS, 1) gather_kblock(const char* __restrict__ feat, const int* __restrict__ rows,
                                                            int64_t n_sel, float* out) {
    constexpr int TILE = 128;
    constexpr int LANES_PER_ROW = PIECE / 16 < THREADS / TILE ? PIECE / 16 : THREADS / TILE;   // threads sharing a row piece
    constexpr int ROWS_PER_PASS = THREADS / LANES_PER_ROW;
    constexpr int PASSES = TILE / ROWS_PER_PASS;                 // row groups per step
    constexpr int V = PIECE / 16 / LANES_PER_ROW;                // float4 per thread, row and step
    constexpr int STEPS = 2048 / PIECE;
    const int t = threadIdx.x, rsub = t / LANES_PER_ROW, l = t % LANES_PER_ROW;
    const int64_t n_tiles = n_sel / TILE;
    float acc = 0.f;
    float4 buf[DEPTH][PASSES * V];
    // flat stream of (tile, step)
    int64_t tile = blockIdx.x;
    int step = 0;
    const float4* src[PASSES];
    auto seek = [&]() {
        if (tile >= n_tiles) return;
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
            src[p] = reinterpret_cast<const float4*>(feat + (int64_t)rows[tile * TILE + p * ROWS_PER_PASS + rsub] * 2048) + l * V;
    };
    auto issue = [&](float4* b) -> bool {
        if (tile >= n_tiles) return false;
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
#pragma unroll
            for (int v = 0; v < V; ++v) b[p * V + v] = __ldg(src[p] + step * (PIECE / 16) + v);
        if (++step == STEPS) { step = 0; tile += gridDim.x; seek(); }
        return true;
    };
    bool pend[DEPTH];
    seek();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) pend[d] = issue(buf[d]);
    while (pend[0]) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (pend[d]) {
#pragma unroll
                for (int i = 0; i < PASSES * V; ++i) acc += buf[d][i].x + buf[d][i].w;
                pend[d] = issue(buf[d]);
            }
        }
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
    printf("%-52s %.3f ms  %.0f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int64_t n_rows = 8ll << 20;  // 16 GiB of rows
    const double density = 0.083;
    char* feat; float* out; cudaMalloc(&feat, n_rows * 2048); cudaMalloc(&out, 4); cudaMemset(feat, 1, n_rows * 2048);
    std::vector<int> sel; srand(1);
    for (int64_t r = 0; r < n_rows; ++r) if (rand() < density * RAND_MAX) sel.push_back((int)r);
    int64_t n_sel = sel.size() / 128 * 128;
    int* rows; cudaMalloc(&rows, n_sel * 4); cudaMemcpy(rows, sel.data(), n_sel * 4, cudaMemcpyHostToDevice);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("%lld selected rows of %lld (%.1f MB gathered), %d SMs\n", (long long)n_sel, (long long)n_rows, n_sel * 2048 / 1e6, sms);
#define RUN(P, T, D) timeit("kblock gather: piece " #P " B, " #T " threads, depth " #D, n_sel * 2048, [&] { gather_kblock<P, T, D><<<sms, T>>>(feat, rows, n_sel, out); })
    RUN(256, 512, 2);
    RUN(256, 512, 4);
    RUN(256, 1024, 2);
    RUN(512, 512, 2);
    RUN(512, 512, 4);
    RUN(512, 1024, 2);
    RUN(1024, 512, 2);
    RUN(1024, 1024, 2);
    RUN(2048, 512, 1);
    RUN(2048, 512, 2);
    RUN(2048, 1024, 1);
    return 0;
}
