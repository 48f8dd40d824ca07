// Error reporting and small queries of the C ABI (include/moc_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace moc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return MOC_E_CUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace moc

extern "C" const char* moc_last_error(void) { return moc::g_err; }
extern "C" int moc_version(void) { return 100; }
extern "C" int moc_num_key_planes(int n_classes) { return moc::key_layout(n_classes).n_planes; }

extern "C" int moc_key_plane(int n_classes, int which) {
    const moc::KeyLayout k = moc::key_layout(n_classes);
    switch (which) {
        case MOC_PLANE_TOP0: return 0;
        case MOC_PLANE_SOFTMAX0: return k.softmax0;
        case MOC_PLANE_DIFF: return k.diff;
        case MOC_PLANE_BG_SUM: return k.bg_sum;
        case MOC_PLANE_BG_MAX: return k.bg_max;
        case MOC_PLANE_LSE: return k.lse;
        default: return -1;
    }
}

namespace moc {
// thread = row: the full 2C+3 planes of the row from whichever layout the class count uses
__global__ void expand_keys_kernel(const float* __restrict__ keys, int64_t key_stride, int C, int64_t n_rows,
                                   float* __restrict__ full, int64_t full_stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const KeyLayout kl = key_layout(C);
    const float* kp = keys + i;
    float* fp = full + i;
    RowSoftmax rs = {0.f};
    if (kl.compact) rs = row_softmax_of(kp, key_stride, kl);
    for (int c = 0; c < C; ++c) {
        const float l = kp[(int64_t)c * key_stride];
        fp[(int64_t)c * full_stride] = l;
        fp[(int64_t)(C + c) * full_stride] = kl.compact ? rs.of(l) : kp[(int64_t)(kl.softmax0 + c) * key_stride];
    }
    fp[(int64_t)(2 * C) * full_stride] = kp[(int64_t)kl.diff * key_stride];
    fp[(int64_t)(2 * C + 1) * full_stride] = kp[(int64_t)kl.bg_sum * key_stride];
    fp[(int64_t)(2 * C + 2) * full_stride] = kp[(int64_t)kl.bg_max * key_stride];
}
}  // namespace moc

extern "C" int moc_expand_keys(const float* keys, int64_t key_stride, int n_classes, int64_t n_rows, float* full,
                               int64_t full_stride, void* stream) {
    MOC_CHECK_ARG(keys && full && n_rows >= 0 && key_stride >= n_rows && full_stride >= n_rows, "moc_expand_keys: bad arguments");
    MOC_CHECK_SHAPE(n_classes >= 1 && n_classes < MOC_MAX_COLS, "moc_expand_keys: bad class count %d", n_classes);
    if (n_rows == 0) return MOC_OK;
    moc::expand_keys_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keys, key_stride, n_classes,
                                                                                              n_rows, full, full_stride);
    MOC_LAUNCH_CHECK("expand_keys_kernel");
    return MOC_OK;
}
