// Library-wide plumbing: error string, launch counter, device info, cached scratch buffers.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace icrl {

static thread_local char g_err[512] = "";
static int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// Everything cached here is keyed by the CURRENT device: the API allows one object per device (cpg passes --cn_device to
// ConstraintNet.load and --device to PPOLagrangian; every call site runs under th.cuda.device(obj_dev)).
constexpr int kMaxDevices = 16;
static int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    return dev;
}

int sm_count() {
    static int cached[kMaxDevices] = {0};
    const int dev = current_device();
    if (cached[dev] <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
        cached[dev] = n;
    }
    return cached[dev];
}

struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
};
static Scratch g_dev[kMaxDevices][SLOT_COUNT], g_pin[kMaxDevices][SLOT_COUNT];

int device_scratch(Slot s, size_t bytes, void** ptr) {
    Scratch& b = g_dev[current_device()][s];
    if (bytes > b.cap) {
        if (b.ptr) {
            ICRL_CUDA(cudaDeviceSynchronize());
            ICRL_CUDA(cudaFree(b.ptr));
            b.ptr = nullptr;
            b.cap = 0;
        }
        size_t want = bytes + bytes / 4 + 4096;
        ICRL_CUDA(cudaMalloc(&b.ptr, want));
        b.cap = want;
    }
    *ptr = b.ptr;
    return 0;
}

int pinned_scratch(Slot s, size_t bytes, void** ptr) {
    Scratch& b = g_pin[current_device()][s];
    if (bytes > b.cap) {
        if (b.ptr) {
            ICRL_CUDA(cudaDeviceSynchronize());
            ICRL_CUDA(cudaFreeHost(b.ptr));
            b.ptr = nullptr;
            b.cap = 0;
        }
        size_t want = bytes + bytes / 4 + 4096;
        ICRL_CUDA(cudaHostAlloc(&b.ptr, want, cudaHostAllocDefault));
        b.cap = want;
    }
    *ptr = b.ptr;
    return 0;
}

}  // namespace icrl

extern "C" {

int icrl_abi_version(void) { return ICRL_ABI_VERSION; }
const char* icrl_last_error(void) { return icrl::g_err; }
int64_t icrl_launch_count(void) { return icrl::g_launches; }

int icrl_device_info(int32_t* sm_count, int32_t* cc) {
    int dev = 0, major = 0, minor = 0, sms = 0;
    ICRL_CUDA(cudaGetDevice(&dev));
    ICRL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    ICRL_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    ICRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sm_count) *sm_count = sms;
    if (cc) *cc = major * 10 + minor;
    return 0;
}

}  // extern "C"
