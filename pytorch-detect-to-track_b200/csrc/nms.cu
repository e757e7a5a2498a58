// nms.cu -- greedy IoU NMS entirely on the device, batched over images, for sm_100a.
//
// Replaces nms_kernel + the host-side part of nms_cuda_compute
// (/root/reference/lib/model/nms/src/nms_cuda_kernel.cu:31-39 devIoU, :41-85 mask kernel,
// :87-161 cudaMalloc / D2H of the whole mask / serial CPU sweep / H2D / cudaFree).
//
// Bit-exactness: devIoU is evaluated with the exact instruction sequence nvcc emits for the
// reference source on sm_100a (checked with cuobjdump -sass on oracle/_ref):
//     Sa = FMUL(a2-a0+1, a3-a1+1)      a = the EARLIER box (the reference's row / cur_box)
//     t  = FFMA(b2-b0+1, b3-b1+1, Sa)  b = the LATER box   (the reference's column box)
//     I  = FMUL(w, h);  den = FADD(t, -I);  iou = I / den (IEEE);  suppress iff iou > thresh
// The roles matter: FFMA(wb,hb,wa*ha) and FFMA(wa,ha,wb*hb) can differ in the last ulp.
//
// Design (B200):
//   kernel 1 (nms_mask): 64x64 IoU tiles of the UPPER triangle only (the reference computes the
//     full square), all images in one launch; boxes of the column block staged in shared memory
//     as SoA; pairs with empty intersection skip the division.  Diagonal tiles store the full
//     symmetric 64-bit word per box in `diag`, off-diagonal tiles store mask[row][colblock].
//   kernel 2 (nms_sweep): one CTA per image walks the 64-box chunks in order.  Warp 0 resolves
//     the intra-chunk dependency chain with a ballot fix-point on the symmetric diag words
//     (a box is kept once every earlier box that overlaps it is known to be removed; it is
//     removed once one of them is known to be kept) instead of a 64-step serial loop; the kept
//     rows of the chunk are then OR-ed into the shared-memory `removed` bitmap by the whole CTA
//     with every load in flight at once.  Early exit once max_keep boxes are kept.
//   No cudaMalloc, no host round trip, stream-ordered.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {
namespace {

typedef unsigned long long u64;

// fl(inter / den) > thresh -- the reference's test (nms_cuda_kernel.cu:31-39 divides in fp32) -- decided WITHOUT the division
// unless the quotient is within 2^-21 of thresh: the rounding of thresh * den and the rounding of the quotient are 2^-24
// each, so outside that band the comparison of inter with thresh * den (1 +- 2^-21) already fixes the outcome; inside it
// (and for a non-positive, tiny or non-finite denominator, or thresh <= 0) the IEEE division decides.  Same bits always.
__device__ __forceinline__ bool quotient_gt(float inter, float den, float thresh) {
    const float p = __fmul_rn(thresh, den);
    const bool up = inter > __fmul_rn(p, 1.00000048f);               // 1 + 2^-21: certainly above
    const bool dn = inter < __fmul_rn(p, 0.99999952f);               // 1 - 2^-21: certainly not above
    bool r = up;
    if (!((up || dn) && den > 1e-10f && thresh > 0.f)) r = __fdiv_rn(inter, den) > thresh;     // (rare)
    return r;
}

// a = earlier box, b = later box; both as (x1,y1,x2,y2).
__device__ __forceinline__ bool iou_gt(float4 a, float Sa, float4 b, float thresh, bool fast_reject) {
    float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
    float h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
    float inter = __fmul_rn(w, h);
    // inter == 0 -> iou is 0, -0 or NaN, none of which is > thresh when thresh >= 0
    if (fast_reject && inter == 0.f) return false;
    float wb = __fadd_rn(__fsub_rn(b.z, b.x), 1.f), hb = __fadd_rn(__fsub_rn(b.w, b.y), 1.f);
    float t = __fmaf_rn(wb, hb, Sa);
    float den = __fsub_rn(t, inter);
    return quotient_gt(inter, den, thresh);
}
// the same decision with the later box's width and height (wb, hb: the expressions above) supplied by the caller
__device__ __forceinline__ bool iou_gt_wh(float4 a, float Sa, float4 b, float wb, float hb, float thresh, bool fast_reject) {
    float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
    float h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
    float inter = __fmul_rn(w, h);
    // (inter == 0: `dn` of quotient_gt holds for every positive denominator, and 0 / den is 0, -0 or NaN otherwise -- never
    // above a thresh >= 0: no separate early exit on this path)
    (void)fast_reject;
    float den = __fsub_rn(__fmaf_rn(wb, hb, Sa), inter);
    return quotient_gt(inter, den, thresh);
}
__device__ __forceinline__ float box_area(float4 a) {
    return __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
}
__device__ __forceinline__ float4 load_box(const float* p) {
    return make_float4(p[0], p[1], p[2], p[3]);
}

// grid (G, B): a block walks the 64x64 tiles t = r * cb + c of its image with stride G and skips the ones below the
// diagonal.  `limit` (a multiple of 64, or >= N) restricts the problem to the first `limit` boxes of every list -- greedy
// NMS on a sorted list decides a prefix without looking past it -- `skip_cb` skips the tiles an earlier prefix pass
// already wrote, and images whose `done` flag is set leave at once (see d2t_nms_batched).
__global__ void __launch_bounds__(64, 16)
nms_mask(const float* __restrict__ boxes, const int* __restrict__ n_valid, int N, int box_dim, int cb,
         float thresh, u64* __restrict__ mask, u64* __restrict__ diag, int limit, int skip_cb,
         const int* __restrict__ done) {
    const int img = blockIdx.y;
    if (done && done[img]) return;
    const int n = min(n_valid ? min(n_valid[img], N) : N, limit);
    const int nb = (n + 63) >> 6;                     // tiles per side that hold boxes
    const float* bx = boxes + (size_t)img * N * box_dim;
    const int tid = threadIdx.x;
    const bool fast = thresh >= 0.f;
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    __shared__ float2 cwh[64];       // width / height of the column boxes (the `wb`, `hb` of iou_gt: once per box, not per pair)
    // upper-triangle tiles only, row-major: row r starts at t0(r) = r * nb - r (r - 1) / 2
    const int ntri = nb * (nb + 1) / 2;
    for (int t = blockIdx.x; t < ntri; t += gridDim.x) {
        int r = (int)(((float)(2 * nb + 1) - sqrtf((float)(2 * nb + 1) * (float)(2 * nb + 1) - 8.f * (float)t)) * 0.5f);
        r = max(0, min(r, nb - 1));
        while (r > 0 && r * nb - r * (r - 1) / 2 > t) --r;                     // (float rounding)
        while ((r + 1) * nb - (r + 1) * r / 2 <= t) ++r;
        const int c = r + (t - (r * nb - r * (r - 1) / 2));
        if (r < skip_cb && c < skip_cb) continue;
        const int col_size = min(n - c * 64, 64), row_size = min(n - r * 64, 64);
        __syncthreads();                               // the previous tile's column boxes are no longer read
        if (tid < col_size) {
            float4 b = load_box(bx + (size_t)(c * 64 + tid) * box_dim);
            cbox[tid] = b;
            carea[tid] = box_area(b);
            cwh[tid] = make_float2(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
        }
        __syncthreads();
        if (tid < row_size) {
            const int row = r * 64 + tid;
            const float4 a = (r == c) ? cbox[tid] : load_box(bx + (size_t)row * box_dim);
            const float Sa = (r == c) ? carea[tid] : box_area(a);
            u64 bits = 0;
            if (r == c) {
                for (int j = 0; j < col_size; ++j) {
                    if (j == tid) continue;
                    bool sp = (j > tid) ? iou_gt(a, Sa, cbox[j], thresh, fast) : iou_gt(cbox[j], carea[j], a, thresh, fast);
                    if (sp) bits |= 1ull << j;
                }
                diag[(size_t)img * N + row] = bits;
            } else {
                if (col_size == 64 && fast) {              // full tile: eight column boxes at a time, constant bit positions
                    unsigned half[2] = {0u, 0u};
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
#pragma unroll 1
                        for (int g = 0; g < 4; ++g) {
                            unsigned m8 = 0u;
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int j = hf * 32 + g * 8 + u;
                                const float2 wh = cwh[j];
                                if (iou_gt_wh(a, Sa, cbox[j], wh.x, wh.y, thresh, true)) m8 |= 1u << u;
                            }
                            half[hf] |= m8 << (g * 8);
                        }
                    }
                    bits = (u64)half[0] | ((u64)half[1] << 32);
                } else {
#pragma unroll 4
                    for (int j = 0; j < col_size; ++j)
                        if (iou_gt(a, Sa, cbox[j], thresh, fast)) bits |= 1ull << j;
                }
                mask[((size_t)img * N + row) * cb + c] = bits;
            }
        }
    }
}

constexpr int kSweepThreads = 512;

__global__ void __launch_bounds__(kSweepThreads)
nms_sweep(const u64* __restrict__ mask, const u64* __restrict__ diag, const int* __restrict__ n_valid, int N,
          int cb, int max_keep, int* __restrict__ keep, int keep_stride, int* __restrict__ num_keep, int limit,
          int* __restrict__ done, int set_done) {
    extern __shared__ u64 removed[];  // [cb]
    __shared__ int s_list[64];
    __shared__ int s_nk, s_count;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (done && !set_done && done[img]) return;       // the prefix pass already produced this image's answer
    const int n_full = n_valid ? min(n_valid[img], N) : N;
    const int n = min(n_full, limit);
    const int nchunks = (n + 63) >> 6;
    const u64* dg = diag + (size_t)img * N;
    const u64* mk = mask + (size_t)img * N * cb;
    int* kp = keep + (size_t)img * keep_stride;
    const int cap = max_keep > 0 ? min(max_keep, keep_stride) : keep_stride;

    for (int t = tid; t < cb; t += blockDim.x) removed[t] = 0;
    if (tid == 0) s_count = 0;
    __syncthreads();

    int count = 0;  // meaningful in warp 0
    for (int k = 0; k < nchunks; ++k) {
        if (warp == 0) {
            const int base = k << 6;
            const int i0 = base + lane, i1 = i0 + 32;
            const u64 S0 = i0 < n ? dg[i0] : 0ull, S1 = i1 < n ? dg[i1] : 0ull;
            const int rem = n - base;
            const u64 valid = rem >= 64 ? ~0ull : ((1ull << rem) - 1ull);
            u64 U = ~removed[k] & valid, K = 0;
            const u64 P0 = S0 & ((1ull << lane) - 1ull);
            const u64 P1 = S1 & ((1ull << (lane + 32)) - 1ull);
            while (U != 0ull) {
                const bool in0 = (U >> lane) & 1ull, in1 = (U >> (lane + 32)) & 1ull;
                const bool rm0 = in0 && (P0 & K), rm1 = in1 && (P1 & K);
                const bool kp0 = in0 && !rm0 && !(P0 & U), kp1 = in1 && !rm1 && !(P1 & U);
                const u64 newK = (u64)__ballot_sync(0xffffffffu, kp0) | ((u64)__ballot_sync(0xffffffffu, kp1) << 32);
                const u64 dec = (u64)__ballot_sync(0xffffffffu, kp0 || rm0) |
                                ((u64)__ballot_sync(0xffffffffu, kp1 || rm1) << 32);
                K |= newK;
                U &= ~dec;
            }
            // keep at most (cap - count) of them, lowest indices first
            const int room = cap - count;
            const int rank0 = __popcll(K & ((1ull << lane) - 1ull));
            const int rank1 = __popcll(K & ((1ull << (lane + 32)) - 1ull));
            const bool w0 = ((K >> lane) & 1ull) && rank0 < room;
            const bool w1 = ((K >> (lane + 32)) & 1ull) && rank1 < room;
            if (w0) { kp[count + rank0] = i0; s_list[rank0] = i0; }
            if (w1) { kp[count + rank1] = i1; s_list[rank1] = i1; }
            const int nk = min(__popcll(K), room);
            count += nk;
            if (lane == 0) { s_nk = nk; s_count = count; }
        }
        __syncthreads();
        const int nk = s_nk;
        const bool done = s_count >= cap;
        if (done || k + 1 >= nchunks) break;
        // OR the kept rows of this chunk into removed[k+1 .. nchunks)
        const int ncols = nchunks - (k + 1);
        for (int ri = warp; ri < nk; ri += kSweepThreads / 32) {
            const u64* rowp = mk + (size_t)s_list[ri] * cb + (k + 1);
#pragma unroll 2
            for (int t = lane; t < ncols; t += 32) {
                const u64 v = rowp[t];
                if (v) atomicOr(&removed[k + 1 + t], v);
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        num_keep[img] = s_count;
        // prefix pass: final iff it already holds max_keep boxes or the prefix was the whole list
        if (set_done) done[img] = (s_count >= cap || limit >= n_full) ? 1 : 0;
    }
}

// ---- capped NMS without the N x N mask (the proposal step: keep the first 300 of 6000, or 2000 of 12000) ----
// Greedy NMS only ever compares a candidate with the boxes KEPT before it, and the proposal step stops at max_keep of them:
// 6000 x 300 IoUs per image instead of the 18 M of the upper triangle, no mask in HBM, one launch.  One 1024-thread CTA
// per image keeps the survivors (box + area) in shared memory and walks the candidates in chunks of 64: 16 threads per
// candidate test it against the kept list (early out at the first hit) and build the chunk's symmetric 64 x 64 overlap
// words; warp 0 then resolves the intra-chunk chain with the same ballot fix-point as nms_sweep and appends the new
// survivors.  Same iou_gt, same roles (a = the earlier box) => the same keep-set, bit for bit.
constexpr int kGreedyMaxKeep = 2048;
constexpr int kGreedyThreads = 1024;

__global__ void __launch_bounds__(kGreedyThreads)
nms_greedy(const float* __restrict__ boxes, const int* __restrict__ n_valid, int N, int box_dim, float thresh, int cap,
           int* __restrict__ keep, int keep_stride, int* __restrict__ num_keep) {
    extern __shared__ float4 kbox[];                 // [cap] kept boxes, then [cap] areas
    float* karea = reinterpret_cast<float*>(kbox + cap);
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    __shared__ u64 S[64];
    __shared__ int s_hit[64];
    __shared__ int s_count;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = n_valid ? min(n_valid[img], N) : N;
    const float* bx = boxes + (size_t)img * N * box_dim;
    int* kp = keep + (size_t)img * keep_stride;
    const bool fast = thresh >= 0.f;
    const int j = tid >> 4, sub = tid & 15;          // candidate of the chunk / position in its 16-thread group
    if (tid == 0) s_count = 0;
    int count = 0;
    const int nchunks = (n + 63) >> 6;
    for (int k = 0; k < nchunks; ++k) {
        const int base = k << 6;
        __syncthreads();                              // the previous chunk's candidates / words are no longer read
        if (tid < 64 && base + tid < n) {
            const float4 b = load_box(bx + (size_t)(base + tid) * box_dim);
            cbox[tid] = b;
            carea[tid] = box_area(b);
        }
        __syncthreads();
        count = s_count;
        const bool valid = base + j < n;
        bool hit = false;
        u64 bits = 0;
        if (valid) {
            const float4 b = cbox[j];
            for (int q = sub; q < count; q += 16)
                if (iou_gt(kbox[q], karea[q], b, thresh, fast)) {
                    hit = true;
                    break;
                }
            const float Sb = carea[j];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = sub + 16 * q;
                if (i == j || base + i >= n) continue;
                const bool sp = i > j ? iou_gt(b, Sb, cbox[i], thresh, fast) : iou_gt(cbox[i], carea[i], b, thresh, fast);
                if (sp) bits |= 1ull << i;
            }
        }
        // combine the 16 threads of a candidate (two candidates per warp)
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
        if (sub == 0) {
            S[j] = bits;
            s_hit[j] = (!valid || ((hm >> (lane & 16)) & 0xffffu)) ? 1 : 0;
        }
        __syncthreads();
        if (warp == 0) {
            const u64 S0 = S[lane], S1 = S[lane + 32];
            u64 U = (u64)__ballot_sync(0xffffffffu, !s_hit[lane]) | ((u64)__ballot_sync(0xffffffffu, !s_hit[lane + 32]) << 32);
            u64 K = 0;
            const u64 P0 = S0 & ((1ull << lane) - 1ull);
            const u64 P1 = S1 & ((1ull << (lane + 32)) - 1ull);
            while (U != 0ull) {
                const bool in0 = (U >> lane) & 1ull, in1 = (U >> (lane + 32)) & 1ull;
                const bool rm0 = in0 && (P0 & K), rm1 = in1 && (P1 & K);
                const bool kp0 = in0 && !rm0 && !(P0 & U), kp1 = in1 && !rm1 && !(P1 & U);
                const u64 newK = (u64)__ballot_sync(0xffffffffu, kp0) | ((u64)__ballot_sync(0xffffffffu, kp1) << 32);
                const u64 dec = (u64)__ballot_sync(0xffffffffu, kp0 || rm0) |
                                ((u64)__ballot_sync(0xffffffffu, kp1 || rm1) << 32);
                K |= newK;
                U &= ~dec;
            }
            const int room = cap - count;
            const int rank0 = __popcll(K & ((1ull << lane) - 1ull));
            const int rank1 = __popcll(K & ((1ull << (lane + 32)) - 1ull));
            if (((K >> lane) & 1ull) && rank0 < room) {
                kp[count + rank0] = base + lane;
                kbox[count + rank0] = cbox[lane];
                karea[count + rank0] = carea[lane];
            }
            if (((K >> (lane + 32)) & 1ull) && rank1 < room) {
                kp[count + rank1] = base + lane + 32;
                kbox[count + rank1] = cbox[lane + 32];
                karea[count + rank1] = carea[lane + 32];
            }
            if (lane == 0) s_count = count + min(__popcll(K), room);
        }
        __syncthreads();
        if (s_count >= cap) break;
    }
    if (tid == 0) num_keep[img] = s_count;
}

}  // namespace
}  // namespace d2t

using namespace d2t;

extern "C" int d2t_nms_prefix(int N, int max_keep);
// MEASURED on B200 (profiles/r02_nms_capped_kernel.json): 6000 spread boxes, keep 300: 6.7 us/image against 37 for mask +
// sweep; but on heavily clustered lists (the random-weight model of bench.py: the cap is reached late or never, every
// chunk meets a long kept list on ONE SM) the step is slower -- 5.75 against 5.67 ms eval, 28.3 against 26.9 ms training.
// So the capped kernel is opt-in: D2T_NMS_GREEDY=1 or d2t_nms_set_mode(1).
static int g_nms_greedy = -1;      // -1: environment default (off unless D2T_NMS_GREEDY=1), 0 / 1: set by d2t_nms_set_mode
static bool nms_greedy_enabled() {
    static const bool env_on = [] { const char* e = getenv("D2T_NMS_GREEDY"); return e && e[0] == '1'; }();
    const int m = g_nms_greedy;
    return m < 0 ? env_on : m != 0;
}
extern "C" int d2t_nms_set_mode(int greedy) {
    g_nms_greedy = greedy < 0 ? -1 : (greedy ? 1 : 0);
    return 1;
}
// kernels d2t_nms_batched launches for this problem (1: the mask-free capped kernel; 2: mask + sweep; 4: with a prefix pass)
extern "C" int d2t_nms_launch_count(int N, int max_keep) {
    if (N <= 0) return 0;
    if (nms_greedy_enabled() && max_keep > 0 && max_keep <= kGreedyMaxKeep) return 1;
    return d2t_nms_prefix(N, max_keep) > 0 ? 4 : 2;
}

extern "C" size_t d2t_nms_workspace_bytes(int B, int N) {
    size_t cb = ((size_t)N + 63) / 64;
    return align_up((size_t)B * N * cb * 8 + (size_t)B * N * 8 + (size_t)B * 4, 256);
}

// boxes decided by the prefix pass of d2t_nms_batched (0: no prefix pass); a multiple of 64
extern "C" int d2t_nms_prefix(int N, int max_keep) {
    if (max_keep <= 0) return 0;
    // Opt-in (D2T_NMS_PREFIX=1).  Measured on B200: 12.5 instead of 36.8 us/image at N = 6000 / keep 300 when the survivors sit
    // in the prefix (spread boxes), +17 % when they do not and both passes run (heavily clustered boxes) -- and the
    // synthetic random-weight model of bench.py is of the second kind (5.95 vs 5.92 ms/step), so it is off by default.
    const char* env = getenv("D2T_NMS_PREFIX");
    if (!env || atoi(env) == 0) return 0;
    int p = ((4 * max_keep + 63) / 64) * 64;
    if (p < 1024) p = 1024;
    return 2 * p <= N ? p : 0;
}

extern "C" int d2t_nms_batched(const float* boxes, const int* n_valid, int B, int N, int box_dim, float thresh,
                               int max_keep, int* keep, int keep_stride, int* num_keep, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream) {
    D2T_REQUIRE(B >= 0 && N >= 0 && box_dim >= 4, "d2t_nms_batched: bad sizes");
    if (B == 0) return 1;
    D2T_REQUIRE(num_keep, "d2t_nms_batched: null num_keep");
    if (N == 0) {
        D2T_CUDA_OK(cudaMemsetAsync(num_keep, 0, sizeof(int) * B, stream), "nms memset");
        return 1;
    }
    D2T_REQUIRE(boxes && keep && workspace, "d2t_nms_batched: null pointer");
    D2T_REQUIRE(workspace_bytes >= d2t_nms_workspace_bytes(B, N), "d2t_nms_batched: workspace too small");
    D2T_REQUIRE(max_keep > 0 ? keep_stride >= 1 : keep_stride >= N, "d2t_nms_batched: keep_stride too small");
    // capped lists (the proposal step) can take the mask-free kernel (opt-in, see nms_greedy_enabled)
    if (nms_greedy_enabled() && max_keep > 0 && max_keep <= kGreedyMaxKeep) {
        const int cap = max_keep < keep_stride ? max_keep : keep_stride;
        const size_t smem = (size_t)cap * 20;
        if (smem > 40 * 1024) {
            static SmemAttrOnce once;
            if (!once.ensure(nms_greedy, kGreedyMaxKeep * 20, "nms_greedy smem attr")) return 0;
        }
        nms_greedy<<<B, kGreedyThreads, smem, stream>>>(boxes, n_valid, N, box_dim, thresh, cap, keep, keep_stride, num_keep);
        D2T_CHECK_LAUNCH("nms_greedy");
        return 1;
    }
    const int cb = (N + 63) / 64;
    D2T_REQUIRE(cb <= 65535 && B <= 65535, "d2t_nms_batched: too many boxes/images");
    D2T_REQUIRE((size_t)cb * 8 <= 200 * 1024, "d2t_nms_batched: N too large for the sweep bitmap");
    u64* mask = reinterpret_cast<u64*>(workspace);
    u64* diag = mask + (size_t)B * N * cb;
    int* done = reinterpret_cast<int*>(diag + (size_t)B * N);
    size_t smem = (size_t)cb * 8;
    if (smem > 40 * 1024) {
        static SmemAttrOnce once;
        if (!once.ensure(nms_sweep, 200 * 1024, "nms_sweep smem attr")) return 0;
    }
    const int gx_cap = sm_count() * 32;   // 64-thread blocks: 32 resident per SM
    // Keeping only the first max_keep survivors (the proposal step: 300 of 6000) may not need the whole list: an optional
    // first pass (d2t_nms_prefix) decides the prefix of 4 * max_keep boxes (exact: greedy NMS never looks ahead) and marks
    // the images it finished; the full pass then runs only for the others and reuses the prefix tiles.
    const int prefix = d2t_nms_prefix(N, max_keep);
    int skip_cb = 0;
    if (prefix > 0) {
        const int pcb = prefix / 64;
        nms_mask<<<dim3(pcb * (pcb + 1) / 2 < gx_cap ? pcb * (pcb + 1) / 2 : gx_cap, B), 64, 0, stream>>>(boxes, n_valid, N, box_dim, cb, thresh, mask,
                                                                                       diag, prefix, 0, nullptr);
        D2T_CHECK_LAUNCH("nms_mask (prefix)");
        nms_sweep<<<B, kSweepThreads, smem, stream>>>(mask, diag, n_valid, N, cb, max_keep, keep, keep_stride, num_keep, prefix,
                                                      done, 1);
        D2T_CHECK_LAUNCH("nms_sweep (prefix)");
        skip_cb = pcb;
    }
    const int* done_c = prefix > 0 ? done : nullptr;
    nms_mask<<<dim3(cb * (cb + 1) / 2 < gx_cap ? cb * (cb + 1) / 2 : gx_cap, B), 64, 0, stream>>>(boxes, n_valid, N, box_dim, cb, thresh, mask, diag,
                                                                             0x7fffffff, skip_cb, done_c);
    D2T_CHECK_LAUNCH("nms_mask");
    nms_sweep<<<B, kSweepThreads, smem, stream>>>(mask, diag, n_valid, N, cb, max_keep, keep, keep_stride, num_keep, 0x7fffffff,
                                                  prefix > 0 ? done : nullptr, 0);
    D2T_CHECK_LAUNCH("nms_sweep");
    return 1;
}

// nms_cuda_kernel.h:5-6 -- device pointers, synchronous, legacy default stream.
extern "C" void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim,
                                 float nms_overlap_thresh) {
    if (boxes_num <= 0) {
        if (num_out) cudaMemsetAsync(num_out, 0, sizeof(int), 0);
        cudaStreamSynchronize(0);
        return;
    }
    ScratchLease lease;
    if (!lease_scratch(1, d2t_nms_workspace_bytes(1, boxes_num), lease)) return;
    if (d2t_nms_batched(boxes_host, nullptr, 1, boxes_num, boxes_dim, nms_overlap_thresh, 0, keep_out, boxes_num,
                        num_out, lease.ptr, lease.bytes, 0))
        cudaStreamSynchronize(0);
}
