// common.cu -- library-level state: last-error string, device info, Part-1 scratch cache.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace d2t {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
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

namespace {
constexpr int kMaxDev = 64;
constexpr int kSlots = 4;
struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
    std::mutex mu;
};
Scratch g_scratch[kMaxDev][kSlots];
}  // namespace

bool lease_scratch(int slot, size_t bytes, ScratchLease& out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kSlots) {
        set_error("lease_scratch: bad device/slot");
        return false;
    }
    Scratch& s = g_scratch[dev][slot];
    out.lock = std::unique_lock<std::mutex>(s.mu);
    if (s.bytes < bytes) {
        if (s.ptr) {
            cudaDeviceSynchronize();  // earlier enqueued work may still use the old block
            cudaFree(s.ptr);
            s.ptr = nullptr;
            s.bytes = 0;
        }
        size_t want = align_up(bytes + bytes / 2, 1 << 20);
        cudaError_t e = cudaMalloc(&s.ptr, want);
        if (e != cudaSuccess) {
            set_error("lease_scratch: cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
            s.ptr = nullptr;
            return false;
        }
        s.bytes = want;
    }
    out.ptr = s.ptr;
    out.bytes = s.bytes;
    return true;
}

}  // namespace d2t

extern "C" {
const char* d2t_version(void) { return "d2t_b200 0.1 (sm_100a)"; }
const char* d2t_last_error(void) { return d2t::g_err; }
int d2t_device_sm_count(void) { return d2t::sm_count(); }
}
