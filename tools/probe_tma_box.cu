// Dev probe: read-only HBM bandwidth of a ring of 2-D / 3-D TMA tensor copies over feat [n_rows][512] fp32 as a function
// of the BOX SHAPE - how many contiguous bytes of a patch row one copy fetches - with no compute at all.  Every CTA walks
// 128-row tiles (persistent, tile = blockIdx.x + i * gridDim.x); a tile is cut into boxes of R rows x S 32-float slices
// (S * 128 contiguous bytes per row); the CTA's 16 warps take the tile's boxes round-robin, each through its own ring.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_tma_box tools/probe_tma_box.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(dst), "l"((uint64_t)m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

constexpr int WARPS = 16;

// R rows x S slices per box; SLOTS ring slots per warp; SLICE_FAST: consecutive boxes of a tile advance along the row first
template <int R, int S, int SLOTS, bool SLICE_FAST>
__global__ void __launch_bounds__(WARPS * 32, 1) box_read(const __grid_constant__ CUtensorMap map, int64_t n_tiles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int BOX_BYTES = R * S * 128;
    constexpr int NB_R = 128 / R, NB_S = 16 / S, NB = NB_R * NB_S;     // boxes per tile
    static_assert(NB % WARPS == 0 || WARPS % NB == 0, "boxes per tile vs warps");
    __shared__ __align__(8) uint64_t bars[WARPS][SLOTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        for (int s = 0; s < SLOTS; ++s) mbar_init(&bars[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane != 0) return;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const uint32_t ring = smem_u32(smem) + warp * SLOTS * BOX_BYTES;
    // flat stream of this CTA's boxes: index u -> tile blockIdx.x + (u / NB) * gridDim.x, box u % NB; the warp takes u = warp, warp + 16, ...
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * NB;
    auto issue = [&](int64_t u, int slot) {
        const int64_t tile = blockIdx.x + (u / NB) * gridDim.x;
        const int b = (int)(u % NB);
        const int rb = SLICE_FAST ? b / NB_S : b % NB_R, sb = SLICE_FAST ? b % NB_S : b / NB_R;
        mbar_expect(&bars[warp][slot], BOX_BYTES);
        tma3(ring + slot * BOX_BYTES, &map, 0, sb * S, (int)(tile * 128 + rb * R), &bars[warp][slot], pol);
    };
    int64_t u = warp;
    for (int s = 0; s < SLOTS; ++s)
        if (u + (int64_t)s * WARPS < total) issue(u + (int64_t)s * WARPS, s);
    int slot = 0;
    uint32_t par = 0;
    for (; u < total; u += WARPS) {
        mbar_wait(&bars[warp][slot], par);
        const int64_t un = u + (int64_t)SLOTS * WARPS;
        if (un < total) issue(un, slot);
        if (++slot == SLOTS) { slot = 0; par ^= 1; }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode;

template <int R, int S, int SLOTS, bool SF>
void run(const char* name, float* buf, int64_t n_rows, int sms) {
    CUtensorMap map;
    const cuuint64_t dims[3] = {32, 16, (cuuint64_t)n_rows};
    const cuuint64_t strides[2] = {128, 2048};
    const cuuint32_t box[3] = {32, S, R};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-44s encode failed %d\n", name, (int)r); return; }
    const size_t smem = (size_t)WARPS * SLOTS * R * S * 128 + 1024;
    if (smem > 227 * 1024) { printf("%-44s needs %zu KB\n", name, smem / 1024); return; }
    cudaFuncSetAttribute(box_read<R, S, SLOTS, SF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(a);
        box_read<R, S, SLOTS, SF><<<sms, WARPS * 32, smem>>>(map, n_rows / 128);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    printf("%-44s ring=%4zu KB  %.3f ms  %5.0f GB/s  (%s)\n", name, (smem - 1024) / 1024, best, n_rows * 2048.0 / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int64_t n_rows = 8ll << 20;   // 16 GiB
    float* buf; cudaMalloc(&buf, n_rows * 2048); cudaMemset(buf, 1, n_rows * 2048);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (argc > 1) sms = atoi(argv[1]);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    encode = (EncodeFn)fn;
    printf("CTAs = %d, 16 warps each; box = R rows x S*128 contiguous bytes\n", sms);
    //   R   S  slots  slice-fast
    run<32, 1, 2, true >("32 rows x 128 B  (4 KB) x2  slice-fast", buf, n_rows, sms);    // what score_keys_tct_kernel issues
    run<32, 1, 2, false>("32 rows x 128 B  (4 KB) x2  row-fast", buf, n_rows, sms);
    run<16, 2, 2, true >("16 rows x 256 B  (4 KB) x2", buf, n_rows, sms);
    run< 8, 2, 4, true >(" 8 rows x 256 B  (2 KB) x4", buf, n_rows, sms);               // ~ score_keys_tc_kernel's boxes
    run< 8, 4, 2, true >(" 8 rows x 512 B  (4 KB) x2", buf, n_rows, sms);
    run< 4, 8, 2, true >(" 4 rows x 1 KB   (4 KB) x2", buf, n_rows, sms);
    run< 2, 16, 2, true>(" 2 rows x 2 KB   (4 KB) x2", buf, n_rows, sms);
    run< 4, 16, 1, true>(" 4 rows x 2 KB   (8 KB) x1", buf, n_rows, sms);
    run< 2, 16, 1, true>(" 2 rows x 2 KB   (4 KB) x1  (64 KB ring)", buf, n_rows, sms);
    run< 1, 16, 2, true>(" 1 row  x 2 KB   (2 KB) x2  (64 KB ring)", buf, n_rows, sms);
    run< 1, 16, 4, true>(" 1 row  x 2 KB   (2 KB) x4", buf, n_rows, sms);
    run<32, 1, 1, true >("32 rows x 128 B  (4 KB) x1  (64 KB ring)", buf, n_rows, sms);
    run< 8, 2, 2, true >(" 8 rows x 256 B  (2 KB) x2  (64 KB ring)", buf, n_rows, sms);
    run<32, 2, 1, true >("32 rows x 256 B  (8 KB) x1", buf, n_rows, sms);
    run<32, 4, 1, false>("32 rows x 512 B (16 KB) x1  (256 KB: skip)", buf, n_rows, sms);
    return 0;
}
