// Dev probe: what read-only HBM bandwidth can the bulk-copy ring reach with no compute at all?
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
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

template <int WARPS, int STAGES, int STAGE_BYTES, bool TOUCH>
__global__ void __launch_bounds__(WARPS * 32, 1) ring_read(const char* __restrict__ src, int64_t n_chunks, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* st0 = smem + (size_t)warp * STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * STAGE_BYTES) + warp * STAGES;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const int64_t stride = (int64_t)gridDim.x * WARPS;
    int64_t g = (int64_t)blockIdx.x * WARPS + warp;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) {
            int64_t gg = g + s * stride;
            if (gg < n_chunks) { mbar_expect(&bars[s], STAGE_BYTES); bulk(st0 + s * STAGE_BYTES, src + gg * STAGE_BYTES, STAGE_BYTES, &bars[s], pol); }
        }
    int stage = 0; uint32_t par = 0; float acc = 0.f;
    for (; g < n_chunks; g += stride) {
        mbar_wait(&bars[stage], par);
        if (TOUCH) {
            const float4* p = reinterpret_cast<const float4*>(st0 + stage * STAGE_BYTES);
            for (int i = lane; i < STAGE_BYTES / 16; i += 32) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        }
        __syncwarp();
        if (lane == 0) {
            int64_t gn = g + (int64_t)STAGES * stride;
            if (gn < n_chunks) { mbar_expect(&bars[stage], STAGE_BYTES); bulk(st0 + stage * STAGE_BYTES, src + gn * STAGE_BYTES, STAGE_BYTES, &bars[stage], pol); }
        }
        if (++stage == STAGES) { stage = 0; par ^= 1; }
    }
    if (TOUCH && acc == 123.456f) out[0] = acc;
}

__global__ void ldg_read(const float4* __restrict__ src, int64_t n, float* out) {
    float acc = 0.f;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int W, int S, int B, bool T>
void run(const char* name, const char* buf, int64_t bytes, float* out, int sms) {
    size_t smem = (size_t)W * S * B + W * S * 8;
    cudaFuncSetAttribute(ring_read<W, S, B, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        ring_read<W, S, B, T><<<sms, W * 32, smem>>>(buf, bytes / B, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    printf("%-34s smem=%6zu KB  %.3f ms  %.0f GB/s  (%s)\n", name, smem / 1024, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int64_t bytes = 16ll << 30;
    char* buf; float* out; cudaMalloc(&buf, bytes); cudaMalloc(&out, 4); cudaMemset(buf, 1, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<8, 3, 8192, false>("ring 8w x 3st x 8K  no-touch", buf, bytes, out, sms);
    run<8, 3, 8192, true>("ring 8w x 3st x 8K  touch", buf, bytes, out, sms);
    run<8, 2, 8192, false>("ring 8w x 2st x 8K  no-touch", buf, bytes, out, sms);
    run<4, 3, 16384, false>("ring 4w x 3st x 16K no-touch", buf, bytes, out, sms);
    run<8, 6, 4096, false>("ring 8w x 6st x 4K  no-touch", buf, bytes, out, sms);
    run<16, 3, 4096, false>("ring 16w x 3st x 4K no-touch", buf, bytes, out, sms);
    run<8, 1, 8192, false>("ring 8w x 1st x 8K  no-touch", buf, bytes, out, sms);
    run<4, 2, 8192, false>("ring 4w x 2st x 8K  no-touch", buf, bytes, out, sms);
    run<2, 3, 32768, false>("ring 2w x 3st x 32K no-touch", buf, bytes, out, sms);
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int blocks : {sms * 4, sms * 8, sms * 16}) {
            float best = 1e9;
            for (int it = 0; it < 5; ++it) {
                cudaEventRecord(a); ldg_read<<<blocks, 512>>>((const float4*)buf, bytes / 16, out); cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
            }
            printf("ldg.cs x4 unroll  blocks=%d            %.3f ms  %.0f GB/s\n", blocks, best, bytes / best / 1e6);
        }
        float best = 1e9; char* dst; cudaMalloc(&dst, bytes / 2);
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a); cudaMemcpyAsync(dst, buf, bytes / 2, cudaMemcpyDeviceToDevice); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
        }
        printf("cudaMemcpy D2D 8 GiB (read+write)        %.3f ms  %.0f GB/s\n", best, (double)bytes / best / 1e6);
    }
    return 0;
}
