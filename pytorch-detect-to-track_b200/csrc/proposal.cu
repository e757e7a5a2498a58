// proposal.cu -- the RPN proposal step around NMS: anchor enumeration + box decode + clip fused
// into one kernel, the top-N gather, and the padded roi write-out.
//
// Replaces the tensor-op chain of /root/reference/lib/model/rpn/proposal_layer.py:67-159
// (numpy shift construction + H2D every call :80-93, two permute+contiguous copies :98-103,
// bbox_transform_inv = 16 elementwise kernels, bbox_transform.py:108-134, a python loop of
// clamp_ calls :156-173, and a python loop over images around index_select / cat / nms / slicing
// :128-159).  Arithmetic is evaluated operation by operation in fp32 with explicit
// round-to-nearest intrinsics (no FMA contraction), i.e. exactly what the reference's chain of
// separate torch kernels computes, so boxes are bit-identical to it on the GPU.
#include "common.cuh"

namespace d2t {
namespace {

// index i = (y*W + x)*A + a  (proposal_layer.py:91-93: anchors.view(1,A,4) + shifts.view(K,1,4))
__global__ void proposal_decode(const float* __restrict__ anchors, int A, const float* __restrict__ deltas,
                                const float* __restrict__ cls_prob, const float* __restrict__ im_info, int B, int H,
                                int W, int stride, float* __restrict__ boxes, float* __restrict__ scores) {
    const int HW = H * W;
    const size_t per = (size_t)HW * A, total = per * B;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / per);
        const int i = (int)(idx % per);
        const int a = i % A, pos = i / A, x = pos % W, y = pos / W;
        const float sx = (float)(x * stride), sy = (float)(y * stride);
        const float ax1 = __fadd_rn(anchors[a * 4 + 0], sx), ay1 = __fadd_rn(anchors[a * 4 + 1], sy);
        const float ax2 = __fadd_rn(anchors[a * 4 + 2], sx), ay2 = __fadd_rn(anchors[a * 4 + 3], sy);
        // bbox_transform.py:109-112
        const float w = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), h = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
        const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, w)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, h));
        const float* d = deltas + ((size_t)b * 4 * A + 4 * a) * HW + pos;
        const float dx = __ldg(d), dy = __ldg(d + HW), dw = __ldg(d + 2 * (size_t)HW), dh = __ldg(d + 3 * (size_t)HW);
        // :119-122
        const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
        const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
        const float hpw = __fmul_rn(0.5f, pw), hph = __fmul_rn(0.5f, ph);
        // :126-132 then clip_boxes :167-171 (im_info = (height, width, scale))
        const float xmax = __fsub_rn(im_info[b * 3 + 1], 1.f), ymax = __fsub_rn(im_info[b * 3 + 0], 1.f);
        float4 o;
        o.x = fminf(fmaxf(__fsub_rn(pcx, hpw), 0.f), xmax);
        o.y = fminf(fmaxf(__fsub_rn(pcy, hph), 0.f), ymax);
        o.z = fminf(fmaxf(__fadd_rn(pcx, hpw), 0.f), xmax);
        o.w = fminf(fmaxf(__fadd_rn(pcy, hph), 0.f), ymax);
        reinterpret_cast<float4*>(boxes)[idx] = o;
        // fg scores are the second A channels (proposal_layer.py:67)
        scores[idx] = __ldg(cls_prob + ((size_t)b * 2 * A + A + a) * HW + pos);
    }
}

__global__ void proposal_gather(const float* __restrict__ boxes, const float* __restrict__ scores,
                                const int64_t* __restrict__ order, int B, int n_total, int order_stride, int n_take,
                                float* __restrict__ dets) {
    const size_t total = (size_t)B * n_take;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / n_take), i = (int)(idx % n_take);
        const int64_t src = order[(size_t)b * order_stride + i];
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        float sc = 0.f;
        if (src >= 0 && src < n_total) {
            bx = reinterpret_cast<const float4*>(boxes)[(size_t)b * n_total + src];
            sc = scores[(size_t)b * n_total + src];
        }
        float* o = dets + idx * 5;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = sc;
    }
}

__global__ void proposal_write_rois(const float* __restrict__ dets, const int* __restrict__ keep, int keep_stride,
                                    const int* __restrict__ num_keep, int B, int n_take, int post,
                                    float* __restrict__ rois) {
    const size_t total = (size_t)B * post;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / post), j = (int)(idx % post);
        float* o = rois + idx * 5;
        o[0] = (float)b;                                   // proposal_layer.py:158
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);        // zero padding, :127
        if (j < min(num_keep[b], keep_stride)) {
            const int k = keep[(size_t)b * keep_stride + j];
            if (k >= 0 && k < n_take) {
                const float* s = dets + ((size_t)b * n_take + k) * 5;
                bx = make_float4(s[0], s[1], s[2], s[3]);
            }
        }
        o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
    }
}

// ---- top-N selection + stable descending sort + gather in ONE kernel (proposal_layer.py:115-137: the reference sorts all
// ~29 k scores of an image and keeps the first pre_nms_topN; here torch.sort -- a cub segmented radix sort over everything --
// and proposal_gather become one launch).  One 1024-thread CTA per image:
//   1. radix select (4 passes of 8 bits, per-warp histograms, equal keys of a warp aggregated with match.any) finds the
//      n_take-th key in descending order and how many of its equals are needed;
//   2. compaction: every score above the threshold, and the first `needed` equal ones in index order, as 64-bit composites
//      (descending-order key << 32 | index) in shared memory;
//   3. bitonic sort of the composites (all distinct: equal scores stay in index order = torch.sort(stable=True));
//   4. gather: dets[b][j] = (box, score) of the j-th composite.
// Keys: float bits mapped so that unsigned ascending order = descending float order; +-0 compare equal and NaN sorts first,
// as in torch.
constexpr int kTopkThreads = 1024;
constexpr int kTopkMaxList = 16384;
constexpr int kTopkKeysPerThread = 32;                            // the image's keys stay in registers: n_total <= 32768

__device__ __forceinline__ uint32_t desc_key(float s) {
    uint32_t u = __float_as_uint(s);
    if ((u << 1) == 0u) u = 0u;                                   // -0 -> +0
    if ((u & 0x7fffffffu) > 0x7f800000u) u = 0x7fc00000u;         // every NaN is the largest value
    const uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending float order
    return ~k;                                                    // descending
}

// SPLIT: steps 1 and 2 only, the n_take composites go to glist[b][n_take] in global memory (no order) and
// proposal_rank_gather finishes the job on the whole device.
template <bool SPLIT>
__global__ void __launch_bounds__(kTopkThreads)
proposal_topk_gather(const float* __restrict__ boxes, const float* __restrict__ scores, int n_total, int n_take, int P,
                     float* __restrict__ dets, unsigned long long* __restrict__ glist) {
    extern __shared__ unsigned long long list_smem[];             // [P] composites (!SPLIT)
    unsigned long long* list = SPLIT ? glist + (size_t)blockIdx.x * n_take : list_smem;
    __shared__ uint32_t whist[32][256];                           // per-warp histograms
    __shared__ uint32_t s_prefix, s_remaining, s_cnt, s_eqbase;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* sc = scores + (size_t)b * n_total;
    // element i = r * 1024 + tid lives in key[r] (coalesced loads; "index order" = r ascending, then tid ascending)
    uint32_t key[kTopkKeysPerThread];
#pragma unroll
    for (int r = 0; r < kTopkKeysPerThread; ++r) {
        const int i = r * kTopkThreads + tid;
        key[r] = i < n_total ? desc_key(__ldg(sc + i)) : 0xffffffffu;      // (no real key is all ones)
    }
    const int nr = (n_total + kTopkThreads - 1) / kTopkThreads;   // rounds that hold elements
    uint32_t prefix = 0, mask = 0, remaining = (uint32_t)n_take;
    const bool all = n_take >= n_total;
    if (!all) {
        for (int pass = 3; pass >= 0; --pass) {
            for (int i = tid; i < 32 * 256; i += kTopkThreads) (&whist[0][0])[i] = 0u;
            __syncthreads();
            const int shift = 8 * pass;
#pragma unroll
            for (int r = 0; r < kTopkKeysPerThread; ++r) {
                if (r < nr) {                                     // (uniform)
                    const uint32_t kd = key[r];
                    const bool in = kd != 0xffffffffu && (kd & mask) == prefix;
                    const uint32_t bin = (kd >> shift) & 255u;
                    // the leading digits of neighbouring scores mostly agree: one add for a warp whose lanes share the bin
                    const uint32_t act = __ballot_sync(0xffffffffu, in);
                    if (act) {
                        const uint32_t ref = __shfl_sync(0xffffffffu, bin, __ffs(act) - 1);
                        const uint32_t agree = __ballot_sync(0xffffffffu, in && bin == ref);
                        if (agree == act) {
                            if (lane == 0) whist[warp][ref] += __popc(act);
                        } else if (in) {
                            atomicAdd(&whist[warp][bin], 1u);
                        }
                    }
                }
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t t = 0;
#pragma unroll 8
                for (int w = 0; w < 32; ++w) t += whist[w][tid];
                whist[0][tid] = t;
            }
            __syncthreads();
            if (warp == 0) {
                // 256 bins, 8 per lane: find the bin in which the running count reaches `remaining`
                uint32_t h[8], local = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    h[q] = whist[0][lane * 8 + q];
                    local += h[q];
                }
                uint32_t incl = local;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += n;
                }
                const uint32_t excl = incl - local;
                const bool mine = excl < remaining && remaining <= incl;       // exactly one lane (total >= remaining)
                if (mine) {
                    uint32_t cum = excl, sel = 7;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (cum + h[q] >= remaining) {
                            sel = q;
                            break;
                        }
                        cum += h[q];
                    }
                    s_prefix = prefix | ((uint32_t)(lane * 8 + sel) << shift);
                    s_remaining = remaining - cum;
                }
            }
            __syncthreads();
            prefix = s_prefix;
            remaining = s_remaining;
            mask |= 255u << shift;
        }
    }
    // prefix = the threshold key T; `remaining` = how many scores equal to T are taken (the first ones in index order)
    if (tid == 0) {
        s_cnt = 0u;
        s_eqbase = 0u;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kTopkKeysPerThread; ++r) {
        if (r < nr) {
            const int i = r * kTopkThreads + tid;
            const uint32_t kd = key[r];
            const bool real = kd != 0xffffffffu;
            const bool lt = real && (all || kd < prefix);
            const bool eq = real && !all && kd == prefix;
            // elements above the threshold: any slot (the sort follows); one counter add per warp
            const uint32_t blt = __ballot_sync(0xffffffffu, lt);
            uint32_t base = 0;
            if (lane == 0 && blt) base = atomicAdd(&s_cnt, (uint32_t)__popc(blt));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lt) list[base + __popc(blt & ((1u << lane) - 1u))] = ((unsigned long long)kd << 32) | (uint32_t)i;
            // equal elements: strictly in index order -> block-wide exclusive count of the round (rare: skipped when the
            // whole block has none)
            const uint32_t beq = __ballot_sync(0xffffffffu, eq);
            if (__syncthreads_or(beq != 0u)) {
                if (lane == 0) whist[1][warp] = (uint32_t)__popc(beq);
                __syncthreads();
                if (warp == 0) {
                    uint32_t v = whist[1][lane], incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += n;
                    }
                    whist[2][lane] = incl - v;                     // exclusive prefix of the warps
                    if (lane == 31) whist[3][0] = incl;            // round total
                }
                __syncthreads();
                if (eq) {
                    const uint32_t rank = s_eqbase + whist[2][warp] + __popc(beq & ((1u << lane) - 1u));
                    if (rank < remaining) {
                        const uint32_t pos = atomicAdd(&s_cnt, 1u);
                        list[pos] = ((unsigned long long)kd << 32) | (uint32_t)i;
                    }
                }
                __syncthreads();
                if (tid == 0) s_eqbase += whist[3][0];
            }
        }
    }
    __syncthreads();
    if constexpr (SPLIT) return;
    const int cnt = (int)s_cnt;                            // == n_take
    for (int i = cnt + tid; i < P; i += kTopkThreads) list[i] = ~0ull;
    __syncthreads();
    // bitonic sort, ascending.  Stages whose partner distance is < 32 stay inside a warp's 64 consecutive elements... kept
    // simple: every stage goes through shared memory, one barrier per stage
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            // thread t handles comparator t of each group of 1024: i = element with bit j clear
            for (int c = tid; c < (P >> 1); c += kTopkThreads) {
                const int i = ((c & ~(j - 1)) << 1) | (c & (j - 1));
                const int l = i | j;
                const unsigned long long a = list[i], d = list[l];
                const bool up = (i & k) == 0;
                if ((a > d) == up) {
                    list[i] = d;
                    list[l] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < n_take; j += kTopkThreads) {
        const uint32_t src = (uint32_t)(list[j] & 0xffffffffull);
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        float v = 0.f;
        if (j < cnt && src < (uint32_t)n_total) {
            bx = reinterpret_cast<const float4*>(boxes)[(size_t)b * n_total + src];
            v = sc[src];
        }
        float* o = dets + ((size_t)b * n_take + j) * 5;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = v;
    }
}

inline int grid_for(size_t total) {
    size_t blocks = (total + 255) / 256, cap = (size_t)sm_count() * 8;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace
}  // namespace d2t

using namespace d2t;

extern "C" int d2t_proposal_decode(const float* anchors, int A, const float* deltas, const float* cls_prob,
                                   const float* im_info, int B, int H, int W, int feat_stride, float* boxes,
                                   float* scores_out, cudaStream_t stream) {
    D2T_REQUIRE(A > 0 && B > 0 && H > 0 && W > 0, "d2t_proposal_decode: bad sizes");
    D2T_REQUIRE(anchors && deltas && cls_prob && im_info && boxes && scores_out, "d2t_proposal_decode: null pointer");
    D2T_REQUIRE(((uintptr_t)boxes & 15) == 0, "d2t_proposal_decode: boxes must be 16-byte aligned");
    const size_t total = (size_t)B * H * W * A;
    proposal_decode<<<grid_for(total), 256, 0, stream>>>(anchors, A, deltas, cls_prob, im_info, B, H, W, feat_stride,
                                                        boxes, scores_out);
    D2T_CHECK_LAUNCH("proposal_decode");
    return 1;
}

extern "C" int d2t_proposal_gather(const float* boxes, const float* scores, const int64_t* order, int B, int n_total,
                                   int order_stride, int n_take, float* dets, cudaStream_t stream) {
    D2T_REQUIRE(B > 0 && n_total > 0 && n_take > 0 && n_take <= order_stride, "d2t_proposal_gather: bad sizes");
    D2T_REQUIRE(boxes && scores && order && dets, "d2t_proposal_gather: null pointer");
    proposal_gather<<<grid_for((size_t)B * n_take), 256, 0, stream>>>(boxes, scores, order, B, n_total, order_stride,
                                                                     n_take, dets);
    D2T_CHECK_LAUNCH("proposal_gather");
    return 1;
}

extern "C" int d2t_proposal_write_rois(const float* dets, const int* keep, int keep_stride, const int* num_keep, int B,
                                       int n_take, int post, float* rois, cudaStream_t stream) {
    D2T_REQUIRE(B > 0 && post > 0 && n_take > 0, "d2t_proposal_write_rois: bad sizes");
    D2T_REQUIRE(dets && keep && num_keep && rois, "d2t_proposal_write_rois: null pointer");
    proposal_write_rois<<<grid_for((size_t)B * post), 256, 0, stream>>>(dets, keep, keep_stride, num_keep, B, n_take,
                                                                       post, rois);
    D2T_CHECK_LAUNCH("proposal_write_rois");
    return 1;
}

// 1 if d2t_proposal_topk_gather takes this problem (the sorted list must fit one CTA's shared memory)
extern "C" int d2t_proposal_topk_supported(int n_total, int n_take) {
    int P = 1;
    while (P < n_take) P <<= 1;
    return n_total > 0 && n_total <= kTopkKeysPerThread * kTopkThreads && n_take > 0 && n_take <= n_total && P <= kTopkMaxList ? 1 : 0;
}

// dets[b][j] = (x1, y1, x2, y2, score) of the j-th highest score of image b, j < n_take: what torch.sort(scores, 1,
// descending=True, stable=True)[1][:, :n_take] followed by d2t_proposal_gather produces (proposal_layer.py:115-137).
extern "C" int d2t_proposal_topk_gather(const float* boxes, const float* scores, int B, int n_total, int n_take, float* dets,
                                        cudaStream_t stream) {
    D2T_REQUIRE(B > 0 && boxes && scores && dets && ((uintptr_t)boxes & 15) == 0 && d2t_proposal_topk_supported(n_total, n_take),
                "d2t_proposal_topk_gather: bad arguments (n_take <= n_total, n_take <= 16384, boxes 16-byte aligned)");
    int P = 1;
    while (P < n_take) P <<= 1;
    const size_t smem = (size_t)P * 8;
    static SmemAttrOnce once;
    if (!once.ensure(proposal_topk_gather<false>, kTopkMaxList * 8, "proposal_topk_gather smem attr")) return 0;
    proposal_topk_gather<false><<<B, kTopkThreads, smem, stream>>>(boxes, scores, n_total, n_take, P, dets, nullptr);
    D2T_CHECK_LAUNCH("proposal_topk_gather");
    return 1;
}

// ---- the same result in two launches that use the whole device: select + compaction (one CTA per image, above), then a
// RANK sort -- the composites are all distinct, so the place of one of them in the sorted list is the number of composites
// below it: n_take^2 independent 64-bit comparisons per image spread over all SMs instead of 91 barrier-separated bitonic
// stages on one -- fused with the gather.
namespace d2t {
namespace {
constexpr int kRankThreads = 128, kRankTile = 2048;
__global__ void __launch_bounds__(kRankThreads)
proposal_rank_gather(const unsigned long long* __restrict__ glist, const float* __restrict__ boxes,
                     const float* __restrict__ scores, int n_total, int n_take, float* __restrict__ dets) {
    __shared__ ulonglong2 tile[kRankTile / 2];
    const int b = blockIdx.y, i = blockIdx.x * kRankThreads + threadIdx.x;
    const unsigned long long* lst = glist + (size_t)b * n_take;
    const unsigned long long mine = i < n_take ? lst[i] : ~0ull;
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;                           // (four independent counters: no serial add chain)
    for (int t0 = 0; t0 < n_take; t0 += kRankTile) {
        __syncthreads();
        for (int j = threadIdx.x; j < kRankTile; j += kRankThreads)
            reinterpret_cast<unsigned long long*>(tile)[j] = t0 + j < n_take ? lst[t0 + j] : ~0ull;   // (padding: never below)
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kRankTile / 2; j += 2) {
            const ulonglong2 v = tile[j], w = tile[j + 1];        // (every thread reads the same 16 bytes: a broadcast)
            r0 += v.x < mine;
            r1 += v.y < mine;
            r2 += w.x < mine;
            r3 += w.y < mine;
        }
    }
    const int rank = (r0 + r1) + (r2 + r3);
    if (i < n_take) {
        const uint32_t src = (uint32_t)(mine & 0xffffffffull);
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        float v = 0.f;
        if (src < (uint32_t)n_total) {
            bx = reinterpret_cast<const float4*>(boxes)[(size_t)b * n_total + src];
            v = scores[(size_t)b * n_total + src];
        }
        float* o = dets + ((size_t)b * n_take + rank) * 5;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = v;
    }
}
}  // namespace
}  // namespace d2t

// ---- the selection itself spread over the device: two 16-bit histogram levels in global memory instead of four 8-bit
// radix passes inside one 1024-thread CTA per image (which holds four SMs for tens of microseconds -- long enough to delay the
// four CTAs of a persistent conv kernel that starts beside it).  Six short launches: hist (level 1) -> scan -> hist (level 2,
// keys inside the threshold bucket) -> scan -> compaction -> rank sort + gather.
namespace d2t {
namespace {
constexpr int kSelBins = 65536;
struct TopkSel {          // per image, in the scratch
    unsigned prefix16;    // the 16 leading bits of the threshold key T
    unsigned remaining;   // how many keys inside the level-1 bucket are taken
    unsigned T;           // the n_take-th key in descending score order
    unsigned need_eq;     // how many keys equal to T are taken (the first ones in index order)
};

// level 1: all keys by their leading 16 bits; level 2: the keys of the threshold bucket by their trailing 16 bits
template <int LEVEL>
__global__ void __launch_bounds__(256)
topk_hist(const float* __restrict__ scores, int n_total, unsigned* __restrict__ hist, const TopkSel* __restrict__ sel) {
    const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_total) return;
    const uint32_t k = desc_key(__ldg(scores + (size_t)b * n_total + i));
    unsigned* h = hist + (size_t)b * kSelBins;
    if (LEVEL == 1) atomicAdd(h + (k >> 16), 1u);
    else if ((k >> 16) == sel[b].prefix16) atomicAdd(h + (k & 0xffffu), 1u);
}

// the bin in which the running count (ascending bins) reaches `want`: one CTA per image, a warp per 2048 consecutive bins
template <int LEVEL>
__global__ void __launch_bounds__(1024)
topk_scan(const unsigned* __restrict__ hist, TopkSel* __restrict__ sel, int n_take) {
    __shared__ unsigned wtot[32];
    __shared__ unsigned s_warp, s_before;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned* h = hist + (size_t)b * kSelBins + warp * 2048;
    const unsigned want = LEVEL == 1 ? (unsigned)n_take : sel[b].remaining;
    unsigned part = 0;
#pragma unroll 8
    for (int r = 0; r < 64; ++r) part += h[r * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) wtot[warp] = part;
    __syncthreads();
    if (warp == 0) {
        const unsigned v = wtot[lane];
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (incl - v < want && want <= incl) {       // exactly one lane (the total is >= want)
            s_warp = (unsigned)lane;
            s_before = incl - v;
        }
    }
    __syncthreads();
    if (warp != (int)s_warp) return;
    unsigned before = s_before;
    for (int r = 0; r < 64; ++r) {                   // 32 consecutive bins per round
        const unsigned v = h[r * 32 + lane];
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const bool hit = before + incl - v < want && want <= before + incl;
        if (__any_sync(0xffffffffu, hit)) {
            if (hit) {
                const unsigned bin = (unsigned)(warp * 2048 + r * 32 + lane), cum = before + incl - v;
                if (LEVEL == 1) {
                    sel[b].prefix16 = bin;
                    sel[b].remaining = want - cum;
                } else {
                    sel[b].T = (sel[b].prefix16 << 16) | bin;
                    sel[b].need_eq = want - cum;
                }
            }
            return;
        }
        before += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// compaction: blocks x < gridDim.x - 1 append the keys above the threshold of their 1024 elements (any order: the rank sort
// follows) behind a per-image counter; the last block of every image walks ALL elements in index order and places the first
// need_eq keys equal to the threshold behind the n_take - need_eq keys above it.
__global__ void __launch_bounds__(256)
topk_compact(const float* __restrict__ scores, int n_total, int n_take, const TopkSel* __restrict__ sel,
             unsigned* __restrict__ counter, unsigned long long* __restrict__ glist) {
    __shared__ unsigned wcnt[8], wbase[8], s_total;
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* sc = scores + (size_t)b * n_total;
    const unsigned T = sel[b].T, need = sel[b].need_eq;
    unsigned long long* lst = glist + (size_t)b * n_take;
    if ((int)blockIdx.x < (int)gridDim.x - 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = blockIdx.x * 1024 + q * 256 + threadIdx.x;
            const uint32_t k = i < n_total ? desc_key(__ldg(sc + i)) : 0xffffffffu;
            const bool lt = i < n_total && k < T;
            const unsigned m = __ballot_sync(0xffffffffu, lt);
            unsigned base = 0;
            if (lane == 0 && m) base = atomicAdd(counter + b, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lt) lst[base + __popc(m & ((1u << lane) - 1u))] = ((unsigned long long)k << 32) | (uint32_t)i;
        }
        return;
    }
    const unsigned n_lt = (unsigned)n_take - need;
    unsigned seen = 0;                                // equal keys before this round (block-uniform)
    for (int i1 = 0; i1 < n_total && seen < need; i1 += 2048) {
      uint32_t kk[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {                   // eight rounds' loads in flight together
          const int i = i1 + q * 256 + threadIdx.x;
          kk[q] = i < n_total ? desc_key(__ldg(sc + i)) : 0xffffffffu;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = i1 + q * 256 + threadIdx.x;
        const bool eq = i < n_total && kk[q] == T;
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        if (!__syncthreads_or(m != 0u)) continue;
        if (lane == 0) wcnt[warp] = (unsigned)__popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < 8; ++w) {
                wbase[w] = t;
                t += wcnt[w];
            }
            s_total = t;
        }
        __syncthreads();
        if (eq) {
            const unsigned rank = seen + wbase[warp] + __popc(m & ((1u << lane) - 1u));
            if (rank < need) lst[n_lt + rank] = ((unsigned long long)T << 32) | (uint32_t)i;
        }
        seen += s_total;
        __syncthreads();
      }
    }
}
}  // namespace
}  // namespace d2t

// [B][n_take] composites | [B] TopkSel | [B] counters (+ pad) | 2 x [B][65536] histogram levels
static size_t topk_list_bytes(int B, int n_take) {
    return ((size_t)(B > 0 ? B : 0) * (size_t)(n_take > 0 ? n_take : 0) * sizeof(unsigned long long) + 255) / 256 * 256;
}
extern "C" size_t d2t_proposal_topk_scratch_bytes(int B, int n_take) {
    const size_t b = (size_t)(B > 0 ? B : 0);
    return topk_list_bytes(B, n_take) + (b * (sizeof(TopkSel) + sizeof(unsigned)) + 255) / 256 * 256 + 2 * b * kSelBins * sizeof(unsigned);
}

// d2t_proposal_topk_gather in six short launches (any n_take <= n_total <= 32768); scratch: d2t_proposal_topk_scratch_bytes(B,
// n_take) bytes, 8-byte aligned.
extern "C" int d2t_proposal_topk_gather_split(const float* boxes, const float* scores, int B, int n_total, int n_take,
                                              float* dets, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    D2T_REQUIRE(B > 0 && boxes && scores && dets && ((uintptr_t)boxes & 15) == 0 && n_total > 0 &&
                    n_total <= kTopkKeysPerThread * kTopkThreads && n_take > 0 && n_take <= n_total,
                "d2t_proposal_topk_gather_split: bad arguments (n_take <= n_total <= 32768, boxes 16-byte aligned)");
    D2T_REQUIRE(scratch && ((uintptr_t)scratch & 7) == 0 && scratch_bytes >= d2t_proposal_topk_scratch_bytes(B, n_take),
                "d2t_proposal_topk_gather_split: scratch too small or misaligned");
    unsigned long long* glist = reinterpret_cast<unsigned long long*>(scratch);
    char* p = reinterpret_cast<char*>(scratch) + topk_list_bytes(B, n_take);
    TopkSel* sel = reinterpret_cast<TopkSel*>(p);
    unsigned* counter = reinterpret_cast<unsigned*>(sel + B);
    const size_t meta = ((size_t)B * (sizeof(TopkSel) + sizeof(unsigned)) + 255) / 256 * 256;
    unsigned* hist1 = reinterpret_cast<unsigned*>(p + meta);
    unsigned* hist2 = hist1 + (size_t)B * kSelBins;
    if (n_take >= n_total) {           // everything is taken: no selection (T = the all-ones key no real score maps to)
        D2T_CUDA_OK(cudaMemsetAsync(p, 0xff, meta, stream), "topk scratch memset");
        D2T_CUDA_OK(cudaMemsetAsync(counter, 0, (size_t)B * sizeof(unsigned), stream), "topk counter memset");
    } else {
        D2T_CUDA_OK(cudaMemsetAsync(p, 0, meta + 2 * (size_t)B * kSelBins * sizeof(unsigned), stream), "topk scratch memset");
        const dim3 ge((n_total + 255) / 256, B);
        topk_hist<1><<<ge, 256, 0, stream>>>(scores, n_total, hist1, sel);
        D2T_CHECK_LAUNCH("topk_hist<1>");
        topk_scan<1><<<B, 1024, 0, stream>>>(hist1, sel, n_take);
        D2T_CHECK_LAUNCH("topk_scan<1>");
        topk_hist<2><<<ge, 256, 0, stream>>>(scores, n_total, hist2, sel);
        D2T_CHECK_LAUNCH("topk_hist<2>");
        topk_scan<2><<<B, 1024, 0, stream>>>(hist2, sel, n_take);
        D2T_CHECK_LAUNCH("topk_scan<2>");
    }
    topk_compact<<<dim3((n_total + 1023) / 1024 + 1, B), 256, 0, stream>>>(scores, n_total, n_take, sel, counter, glist);
    D2T_CHECK_LAUNCH("topk_compact");
    proposal_rank_gather<<<dim3((n_take + kRankThreads - 1) / kRankThreads, B), kRankThreads, 0, stream>>>(
        glist, boxes, scores, n_total, n_take, dets);
    D2T_CHECK_LAUNCH("proposal_rank_gather");
    return 1;
}
