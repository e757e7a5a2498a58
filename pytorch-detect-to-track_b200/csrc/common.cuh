// common.cuh -- shared helpers for libd2t_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>

#include "../../include/d2t_b200.h"

namespace d2t {

void set_error(const char* fmt, ...);
int sm_count();

// Per-device scratch for the reference-named (Part 1) entry points.  Grown on demand;
// the returned lock keeps it exclusive for the duration of the enqueue.
struct ScratchLease {
    void* ptr = nullptr;
    size_t bytes = 0;
    std::unique_lock<std::mutex> lock;
};
bool lease_scratch(int slot, size_t bytes, ScratchLease& out);

#define D2T_CHECK_LAUNCH(what)                                                      \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            d2t::set_error("%s: %s", what, cudaGetErrorString(e__));                \
            return 0;                                                               \
        }                                                                           \
    } while (0)

#define D2T_CUDA_OK(call, what)                                                     \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            d2t::set_error("%s: %s", what, cudaGetErrorString(e__));                \
            return 0;                                                               \
        }                                                                           \
    } while (0)

#define D2T_REQUIRE(cond, ...)                                                      \
    do {                                                                            \
        if (!(cond)) {                                                              \
            d2t::set_error(__VA_ARGS__);                                            \
            return 0;                                                               \
        }                                                                           \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Opt a kernel into > 48 KB of dynamic shared memory, once per device.
struct SmemAttrOnce {
    bool done[64] = {};
    template <typename F>
    bool ensure(F func, size_t bytes, const char* what) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        if (done[dev]) return true;
        cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();   // do not leave it behind for the next launch check
            set_error("%s: %s", what, cudaGetErrorString(e));
            return false;
        }
        done[dev] = true;
        return true;
    }
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP).  dst, src 16-byte aligned,
// bytes a multiple of 16.  Completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 3xFP16 operand scale (conv.cu): sa = 2^ea puts a tensor's max |x| (`amax`, one float in device memory) into
// [2^14, 2^15) -- the fp16 (hi, lo) split of x * sa then neither overflows nor loses its low bits
__device__ __forceinline__ int act_exp(const float* amax) {
    int eb = (int)((__float_as_uint(__ldcg(amax)) >> 23) & 0xffu);
    eb = eb < 15 ? 15 : (eb > 254 ? 254 : eb);
    return 141 - eb;
}
__device__ __forceinline__ float pow2f(int e) {
    e = e < -126 ? -126 : (e > 127 ? 127 : e);
    return __uint_as_float((uint32_t)(e + 127) << 23);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace d2t
