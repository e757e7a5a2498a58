#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
// microbench 2: who limits unique L2 -> SM streaming?  (a) LDG.256 rings with 4/8/16 warps on 16/74/148 SMs; (b) 1-D bulk TMA with a 3-stage ring
__device__ __forceinline__ void ldg256(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(v[0]),"=f"(v[1]),"=f"(v[2]),"=f"(v[3]),"=f"(v[4]),"=f"(v[5]),"=f"(v[6]),"=f"(v[7]) : "l"(p));
}
__global__ void __launch_bounds__(512, 1) kl(const float* __restrict__ x, size_t nrows, int nwarps, int iters, float* out, long long* cyc) {
    const int warp = threadIdx.x >> 5;
    if (warp >= nwarps) return;
    float buf[64];
    float acc = 0.f;
    const size_t rows_per_it = (size_t)gridDim.x * nwarps * 32;
    size_t r = (size_t)blockIdx.x * nwarps * 32 + threadIdx.x;
    long long t0 = clock64();
    const float* row = x + r * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) ldg256(row + j * 8, buf + j * 8);
    for (int it = 0; it < iters; ++it) {
        r += rows_per_it; if (r >= nrows) r -= nrows;
        const float* nxt = x + r * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc += buf[j * 8 + e];
            ldg256(nxt + j * 8, buf + j * 8);
        }
    }
#pragma unroll
    for (int e = 0; e < 64; ++e) acc += buf[e];
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) out[0] = acc;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) kt(const float* __restrict__ x, size_t nbytes, int chunk, int stages, int iters, long long* cyc) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    size_t off = (size_t)blockIdx.x * chunk;
    const size_t step = (size_t)gridDim.x * chunk;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(x);
    for (int it = 0; it < iters + stages; ++it) {
        const int s = it % stages;
        if (it >= stages) {
            const uint32_t par = ((it / stages) - 1) & 1;
            uint32_t ok = 0;
            while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(s32(&bar[s])), "r"(par) : "memory");
        }
        if (it < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)s * chunk)), "l"(base + off), "r"(chunk), "r"(s32(&bar[s])) : "memory");
            off += step; if (off + chunk > nbytes) off = (size_t)blockIdx.x * chunk;
        }
    }
    cyc[blockIdx.x] = clock64() - t0;
}
int main() {
    const size_t nrows = 38304 * 4;     // x 256 B = 39.2 MB (L2 resident)
    float* x; cudaMalloc(&x, nrows * 256); cudaMemset(x, 0, nrows * 256);
    float* out; cudaMalloc(&out, 4);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    long long h[148];
    for (int grid : {16, 74, 148}) for (int nw : {4, 8, 16}) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) kl<<<grid, 512>>>(x, nrows, nw, iters, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
        printf("LDG.256 grid %3d warps %2d: %.1f B/clk/SM  (%.2f KB/clk chip)  %s\n", grid, nw, nw * 32 * 256.0 * iters / mean, grid * nw * 32 * 256.0 * iters / mean / 1024, cudaGetErrorString(cudaGetLastError()));
    }
    for (int grid : {16, 74, 148}) for (int chunk : {8192, 32768}) for (int stages : {3, 6}) {
        if (chunk * stages > 200 * 1024) continue;
        const int iters = 4000 * 8192 / chunk * 4;
        cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int rep = 0; rep < 2; ++rep) kt<<<grid, 128, chunk * stages>>>(x, nrows * 256, chunk, stages, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
        printf("bulk TMA grid %3d chunk %5d stages %d: %.1f B/clk/SM  (%.2f KB/clk chip)  %s\n", grid, chunk, stages, (double)chunk * iters / mean, grid * (double)chunk * iters / mean / 1024, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
