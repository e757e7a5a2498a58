// roi_ops.cu -- RoIAlign, RoIPool and RoICrop (bilinear grid sampler), forward + backward.
//
// Replace, with the same semantics:
//   ROIAlignForward/Backward      /root/reference/lib/model/roi_align/src/roi_align_kernel.cu:15-70, 94-143
//   ROIPoolForward/Backward       /root/reference/lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93, 128-203
//   bilinearSamplingFromGrid / backwardBilinearSampling
//                                 /root/reference/lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:47-109, 111-194
// These three are on the north-star list but not on the D&T graph (_RFCN wires PSRoI only,
// rfcn.py:40-43); they are gather kernels with one thread per output element, grid-stride over
// a grid sized from the SM count, coalesced along the innermost output axis.  Arithmetic
// follows what nvcc emits for the reference on sm_100a (explicit FMA where it contracts;
// RoIAlign keeps the reference's double-precision bin size and interpolation).
// Differences from the reference, all in its favour:
//   * RoIAlign's image offset is computed in integers (the reference multiplies it out in
//     fp32, roi_align_kernel.cu:49, which is inexact beyond 2^24 elements);
//   * RoIPool backward scatters top_diff to its argmax instead of scanning every RoI for every
//     input element (O(R*C*P*P) instead of O(B*C*H*W*R)); same sums, different add order;
//   * RoICrop writes 0 when all four neighbours are outside (the reference leaves the
//     pre-zeroed output untouched).
#include "common.cuh"

namespace d2t {
namespace {

inline int grid_for(size_t total) {
    size_t blocks = (total + 255) / 256, cap = (size_t)sm_count() * 16;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

#define D2T_GRID_STRIDE(i, n)                                                          \
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n);           \
         i += (size_t)gridDim.x * blockDim.x)

// ------------------------------------------------------------------ deterministic backward (fixed point)
// The reference's three backward kernels add fp32 terms with atomicAdd in whatever order the threads arrive: the result
// changes in its last bits from run to run.  The *_det entry points accumulate every term as a 64-bit fixed-point integer
// instead -- q = rint(term * 2^k) added with the L2's native 64-bit integer atomic (RED.ADD.U64; integer addition is
// associative, so the order does not matter) -- and convert the finished sums to float once.  k = 62 - NB - (eb - 126) with
// eb the biased exponent of max |top_diff| (one reduction pass) and 2^NB > the number of terms (every interpolation weight
// is <= 1): no sum can leave 63 bits, and a term is resolved to 2^-(62 - NB) of max |top_diff| -- below 2^-37 for 2^24 terms.
// A NaN / Inf gradient cannot be scaled: such a call takes the float atomics (and propagates it like the reference).
// scratch: [0] = max |top_diff| bits, [1] = unused, [2 ...] = one int64 accumulator per element of the gradient tensor.
struct DetScale {
    double sc, isc;    // 2^k, 2^-k
    bool ok;           // false: non-finite gradients, use the float atomics
};
__device__ __forceinline__ DetScale det_scale(const unsigned long long* scratch, int nb) {
    const unsigned bits = (unsigned)scratch[0];
    const int eb = (int)(bits >> 23);
    DetScale d;
    d.ok = eb < 255;
    const int k = 62 - nb - (eb - 126);                        // eb in [0, 254], nb in [1, 48]: k in [-114, 187]
    d.sc = __longlong_as_double((long long)(1023 + k) << 52);
    d.isc = __longlong_as_double((long long)(1023 - k) << 52);
    return d;
}
__device__ __forceinline__ void det_add(unsigned long long* acc, size_t i, double term, const DetScale& d) {
    atomicAdd(acc + i, (unsigned long long)__double2ll_rn(term * d.sc));
}

__global__ void det_amax(const float* __restrict__ x, size_t n, unsigned long long* __restrict__ scratch) {
    unsigned m = 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(x[i]) & 0x7fffffffu);        // (unsigned order = float order; Inf / NaN sort on top)
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m > (unsigned)*reinterpret_cast<volatile unsigned long long*>(scratch))
        atomicMax(scratch, (unsigned long long)m);
}

// out (+)= float(acc * 2^-k); nothing to do when the scatter kernel took the float atomics
__global__ void det_finish(const unsigned long long* __restrict__ scratch, int nb, size_t n, float* __restrict__ out, int assign) {
    const DetScale d = det_scale(scratch, nb);
    if (!d.ok) return;
    const long long* acc = reinterpret_cast<const long long*>(scratch + 2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = (float)((double)acc[i] * d.isc);
        out[i] = assign ? v : out[i] + v;
    }
}

// ------------------------------------------------------------------ RoIAlign
struct AlignSample {
    bool inside;
    int hs, ws;
    float hr, wr;
};

__device__ __forceinline__ AlignSample align_sample(const float* roi, float scale, int AH, int AW, int H, int W,
                                                    int ph, int pw) {
    AlignSample s;
    const float sw = __fmul_rn(roi[1], scale), sh = __fmul_rn(roi[2], scale);
    const float rw = fmaxf(__fadd_rn(__fmaf_rn(roi[3], scale, -sw), 1.f), 0.f);
    const float rh = fmaxf(__fadd_rn(__fmaf_rn(roi[4], scale, -sh), 1.f), 0.f);
    const float bh = (float)((double)rh / ((double)AH - 1.));
    const float bw = (float)((double)rw / ((double)AW - 1.));
    const float h = __fmaf_rn((float)ph, bh, sh), w = __fmaf_rn((float)pw, bw, sw);
    s.inside = !(h < 0 || h >= H || w < 0 || w >= W);
    s.hs = (int)fminf(floorf(h), (float)(H - 2));
    s.ws = (int)fminf(floorf(w), (float)(W - 2));
    s.hr = __fsub_rn(h, (float)s.hs);
    s.wr = __fsub_rn(w, (float)s.ws);
    return s;
}

__global__ void roi_align_fwd(const float* __restrict__ feat, float scale, int R, int H, int W, int C, int AH,
                              int AW, const float* __restrict__ rois, float* __restrict__ top) {
    const size_t total = (size_t)R * C * AH * AW;
    D2T_GRID_STRIDE(index, total) {
        const int pw = (int)(index % AW), ph = (int)((index / AW) % AH);
        const int c = (int)((index / AW / AH) % C), n = (int)(index / AW / AH / C);
        const float* roi = rois + (size_t)n * 5;
        const AlignSample s = align_sample(roi, scale, AH, AW, H, W, ph, pw);
        if (!s.inside) {
            top[index] = 0.f;
            continue;
        }
        const int b = (int)roi[0];
        const float* p = feat + (((size_t)b * C + c) * H + s.hs) * W + s.ws;
        const double hr = s.hr, wr = s.wr;
        const double v = (double)__ldg(p) * (1. - hr) * (1. - wr) + (double)__ldg(p + 1) * (1. - hr) * wr +
                         (double)__ldg(p + W) * hr * (1. - wr) + (double)__ldg(p + W + 1) * hr * wr;
        top[index] = (float)v;
    }
}

template <bool DET>
__global__ void roi_align_bwd(const float* __restrict__ top_diff, float scale, int R, int H, int W, int C, int AH,
                              int AW, const float* __restrict__ rois, float* __restrict__ bottom_diff,
                              unsigned long long* __restrict__ scratch, int nb) {
    const size_t total = (size_t)R * C * AH * AW;
    DetScale d = {};
    if constexpr (DET) d = det_scale(scratch, nb);
    D2T_GRID_STRIDE(index, total) {
        const int pw = (int)(index % AW), ph = (int)((index / AW) % AH);
        const int c = (int)((index / AW / AH) % C), n = (int)(index / AW / AH / C);
        const float* roi = rois + (size_t)n * 5;
        const AlignSample s = align_sample(roi, scale, AH, AW, H, W, ph, pw);
        if (!s.inside) continue;
        const int b = (int)roi[0];
        float* p = bottom_diff + (((size_t)b * C + c) * H + s.hs) * W + s.ws;
        const double g = top_diff[index], hr = s.hr, wr = s.wr;
        if (DET && d.ok) {
            const size_t i = (size_t)(p - bottom_diff);
            det_add(scratch + 2, i, g * (1. - hr) * (1. - wr), d);
            det_add(scratch + 2, i + 1, g * (1. - hr) * wr, d);
            det_add(scratch + 2, i + W, g * hr * (1. - wr), d);
            det_add(scratch + 2, i + W + 1, g * hr * wr, d);
            continue;
        }
        atomicAdd(p, (float)(g * (1. - hr) * (1. - wr)));
        atomicAdd(p + 1, (float)(g * (1. - hr) * wr));
        atomicAdd(p + W, (float)(g * hr * (1. - wr)));
        atomicAdd(p + W + 1, (float)(g * hr * wr));
    }
}

// ------------------------------------------------------------------ RoIPool
__global__ void roi_pool_fwd(const float* __restrict__ feat, float scale, int R, int H, int W, int C, int PH,
                             int PW, const float* __restrict__ rois, float* __restrict__ top,
                             int* __restrict__ argmax) {
    const size_t total = (size_t)R * C * PH * PW;
    D2T_GRID_STRIDE(index, total) {
        const int pw = (int)(index % PW), ph = (int)((index / PW) % PH);
        const int c = (int)((index / PW / PH) % C), n = (int)(index / PW / PH / C);
        const float* roi = rois + (size_t)n * 5;
        const int b = (int)roi[0];
        const int rsw = (int)roundf(__fmul_rn(roi[1], scale)), rsh = (int)roundf(__fmul_rn(roi[2], scale));
        const int rew = (int)roundf(__fmul_rn(roi[3], scale)), reh = (int)roundf(__fmul_rn(roi[4], scale));
        const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);
        const float bh = __fdiv_rn((float)rh, (float)PH), bw = __fdiv_rn((float)rw, (float)PW);
        int hs = __float2int_rd(__fmul_rn((float)ph, bh)), ws = __float2int_rd(__fmul_rn((float)pw, bw));
        int he = __float2int_ru(__fmul_rn((float)(ph + 1), bh)), we = __float2int_ru(__fmul_rn((float)(pw + 1), bw));
        hs = min(max(hs + rsh, 0), H);
        he = min(max(he + rsh, 0), H);
        ws = min(max(ws + rsw, 0), W);
        we = min(max(we + rsw, 0), W);
        const bool empty = (he <= hs) || (we <= ws);
        float mv = empty ? 0.f : -3.402823466e+38f;
        int mi = -1;
        const int base = (b * C + c) * H * W;   // int, like the reference's argmax contract
        for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
                const float v = __ldg(feat + (size_t)base + h * W + w);
                if (v > mv) {
                    mv = v;
                    mi = base + h * W + w;
                }
            }
        top[index] = mv;
        if (argmax) argmax[index] = mi;
    }
}

template <bool DET>
__global__ void roi_pool_bwd(const float* __restrict__ top_diff, const int* __restrict__ argmax, size_t total,
                             size_t limit, float* __restrict__ bottom_diff, unsigned long long* __restrict__ scratch, int nb) {
    DetScale d = {};
    if constexpr (DET) d = det_scale(scratch, nb);
    D2T_GRID_STRIDE(index, total) {
        const int a = argmax[index];
        if (a < 0 || (size_t)a >= limit) continue;
        if (DET && d.ok) det_add(scratch + 2, (size_t)a, (double)top_diff[index], d);
        else atomicAdd(bottom_diff + a, top_diff[index]);
    }
}

// ------------------------------------------------------------------ RoICrop
__device__ __forceinline__ void top_left(float x, int width, int& point, float& weight) {
    // roi_crop_cuda_kernel.cu:11-22: xcoord = (x + 1) * (width - 1) / 2
    const float xc = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), (float)(width - 1)), 0.5f);
    point = __float2int_rd(xc);
    weight = __fsub_rn(1.f, __fsub_rn(xc, (float)point));
}

struct CropSample {
    int x0, y0;
    float wx, wy;
    bool tl, tr, bl, br;
};

__device__ __forceinline__ CropSample crop_sample(const float* g, int H, int W) {
    CropSample s;
    top_left(g[1], W, s.x0, s.wx);
    top_left(g[0], H, s.y0, s.wy);
    const bool x0in = s.x0 >= 0 && s.x0 <= W - 1, x1in = s.x0 + 1 >= 0 && s.x0 + 1 <= W - 1;
    const bool y0in = s.y0 >= 0 && s.y0 <= H - 1, y1in = s.y0 + 1 >= 0 && s.y0 + 1 <= H - 1;
    s.tl = x0in && y0in;
    s.tr = x1in && y0in;
    s.bl = x0in && y1in;
    s.br = x1in && y1in;
    return s;
}

__global__ void roi_crop_fwd(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                             int C, int H, int W, int R, int gh, int gw, int per_image) {
    const size_t total = (size_t)R * C * gh * gw;
    D2T_GRID_STRIDE(index, total) {
        const int xo = (int)(index % gw), yo = (int)((index / gw) % gh);
        const int c = (int)((index / gw / gh) % C), b = (int)(index / gw / gh / C);
        const CropSample s = crop_sample(grid + (((size_t)b * gh + yo) * gw + xo) * 2, H, W);
        const float* p = img + ((size_t)(b / per_image) * C + c) * H * W + (ptrdiff_t)s.y0 * W + s.x0;
        const float tl = s.tl ? __ldg(p) : 0.f, tr = s.tr ? __ldg(p + 1) : 0.f;
        const float bl = s.bl ? __ldg(p + W) : 0.f, br = s.br ? __ldg(p + W + 1) : 0.f;
        const float wx = s.wx, wy = s.wy, ux = 1.f - wx, uy = 1.f - wy;
        out[index] = wx * wy * tl + ux * wy * tr + wx * uy * bl + ux * uy * br;
    }
}

template <bool DET>
__global__ void roi_crop_bwd(const float* __restrict__ grid, const float* __restrict__ gout, float* __restrict__ gimg,
                             int C, int H, int W, int R, int gh, int gw, int per_image,
                             unsigned long long* __restrict__ scratch, int nb) {
    const size_t total = (size_t)R * C * gh * gw;
    DetScale d = {};
    if constexpr (DET) d = det_scale(scratch, nb);
    D2T_GRID_STRIDE(index, total) {
        const int xo = (int)(index % gw), yo = (int)((index / gw) % gh);
        const int c = (int)((index / gw / gh) % C), b = (int)(index / gw / gh / C);
        const CropSample s = crop_sample(grid + (((size_t)b * gh + yo) * gw + xo) * 2, H, W);
        float* p = gimg + ((size_t)(b / per_image) * C + c) * H * W + (ptrdiff_t)s.y0 * W + s.x0;
        const float go = gout[index];
        const float wx = s.wx, wy = s.wy, ux = 1.f - wx, uy = 1.f - wy;
        if (DET && d.ok) {                               // (the same fp32 products as below, summed exactly)
            const ptrdiff_t i = p - gimg;
            if (s.tl) det_add(scratch + 2, (size_t)i, (double)(wx * wy * go), d);
            if (s.tr) det_add(scratch + 2, (size_t)(i + 1), (double)(ux * wy * go), d);
            if (s.bl) det_add(scratch + 2, (size_t)(i + W), (double)(wx * uy * go), d);
            if (s.br) det_add(scratch + 2, (size_t)(i + W + 1), (double)(ux * uy * go), d);
            continue;
        }
        if (s.tl) atomicAdd(p, wx * wy * go);
        if (s.tr) atomicAdd(p + 1, ux * wy * go);
        if (s.bl) atomicAdd(p + W, wx * uy * go);
        if (s.br) atomicAdd(p + W + 1, ux * uy * go);
    }
}

}  // namespace
}  // namespace d2t

using namespace d2t;

extern "C" int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                      const int height, const int width, const int channels,
                                      const int aligned_height, const int aligned_width, const float* bottom_rois,
                                      float* top_data, cudaStream_t stream) {
    if (num_rois <= 0) return 1;
    D2T_REQUIRE(bottom_data && bottom_rois && top_data, "ROIAlignForwardLaucher: null pointer");
    D2T_REQUIRE(aligned_height > 1 && aligned_width > 1 && height >= 2 && width >= 2,
                "ROIAlignForwardLaucher: aligned size and feature size must be >= 2");
    const size_t total = (size_t)num_rois * channels * aligned_height * aligned_width;
    roi_align_fwd<<<grid_for(total), 256, 0, stream>>>(bottom_data, spatial_scale, num_rois, height, width, channels,
                                                      aligned_height, aligned_width, bottom_rois, top_data);
    D2T_CHECK_LAUNCH("roi_align_fwd");
    return 1;
}

extern "C" int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                       const int num_rois, const int height, const int width, const int channels,
                                       const int aligned_height, const int aligned_width, const float* bottom_rois,
                                       float* bottom_diff, cudaStream_t stream) {
    (void)batch_size;
    if (num_rois <= 0) return 1;
    D2T_REQUIRE(top_diff && bottom_rois && bottom_diff, "ROIAlignBackwardLaucher: null pointer");
    D2T_REQUIRE(aligned_height > 1 && aligned_width > 1 && height >= 2 && width >= 2,
                "ROIAlignBackwardLaucher: aligned size and feature size must be >= 2");
    const size_t total = (size_t)num_rois * channels * aligned_height * aligned_width;
    roi_align_bwd<false><<<grid_for(total), 256, 0, stream>>>(top_diff, spatial_scale, num_rois, height, width, channels,
                                                             aligned_height, aligned_width, bottom_rois, bottom_diff,
                                                             nullptr, 0);
    D2T_CHECK_LAUNCH("roi_align_bwd");
    return 1;
}

// ---- deterministic variants (fixed-point accumulation, see det_scale above) ----
extern "C" size_t d2t_roi_backward_scratch_bytes(size_t grad_elems) { return (grad_elems + 2) * sizeof(unsigned long long); }

namespace {
// zero the scratch, find max |top_diff|; returns NB (2^NB > number of terms)
int det_begin(const float* top_diff, size_t total, size_t grad_elems, void* scratch, size_t scratch_bytes, cudaStream_t stream,
              const char* what, int* nb) {
    if (!scratch || ((uintptr_t)scratch & 7) || scratch_bytes < d2t_roi_backward_scratch_bytes(grad_elems)) {
        set_error("%s: scratch must be 8-byte aligned and hold d2t_roi_backward_scratch_bytes(elements of the gradient)", what);
        return 0;
    }
    D2T_CUDA_OK(cudaMemsetAsync(scratch, 0, d2t_roi_backward_scratch_bytes(grad_elems), stream), "roi backward scratch memset");
    det_amax<<<grid_for(total), 256, 0, stream>>>(top_diff, total, reinterpret_cast<unsigned long long*>(scratch));
    D2T_CHECK_LAUNCH("det_amax");
    int b = 1;
    while (((size_t)1 << b) <= total && b < 48) ++b;
    *nb = b;
    return 1;
}
}  // namespace

// bottom_diff += the RoIAlign gradient (like the reference launcher, which adds into a caller-zeroed tensor)
extern "C" int d2t_roi_align_backward_det(const float* top_diff, float spatial_scale, int batch_size, int num_rois, int height,
                                          int width, int channels, int aligned_height, int aligned_width,
                                          const float* bottom_rois, float* bottom_diff, void* scratch, size_t scratch_bytes,
                                          cudaStream_t stream) {
    if (num_rois <= 0) return 1;
    D2T_REQUIRE(top_diff && bottom_rois && bottom_diff, "d2t_roi_align_backward_det: null pointer");
    D2T_REQUIRE(aligned_height > 1 && aligned_width > 1 && height >= 2 && width >= 2 && batch_size > 0,
                "d2t_roi_align_backward_det: aligned size and feature size must be >= 2");
    const size_t total = (size_t)num_rois * channels * aligned_height * aligned_width;
    const size_t elems = (size_t)batch_size * channels * height * width;
    int nb = 0;
    if (!det_begin(top_diff, total, elems, scratch, scratch_bytes, stream, "d2t_roi_align_backward_det", &nb)) return 0;
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(scratch);
    roi_align_bwd<true><<<grid_for(total), 256, 0, stream>>>(top_diff, spatial_scale, num_rois, height, width, channels,
                                                            aligned_height, aligned_width, bottom_rois, bottom_diff, sc, nb);
    D2T_CHECK_LAUNCH("roi_align_bwd<det>");
    det_finish<<<grid_for(elems), 256, 0, stream>>>(sc, nb, elems, bottom_diff, 0);
    D2T_CHECK_LAUNCH("det_finish");
    return 1;
}

// bottom_diff = the RoIPool gradient (assigned, like ROIPoolBackwardLaucher)
extern "C" int d2t_roi_pool_backward_det(const float* top_diff, int batch_size, int num_rois, int height, int width,
                                         int channels, int pooled_height, int pooled_width, float* bottom_diff,
                                         const int* argmax_data, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    D2T_REQUIRE(bottom_diff && batch_size > 0, "d2t_roi_pool_backward_det: null bottom_diff");
    const size_t limit = (size_t)batch_size * channels * height * width;
    D2T_CUDA_OK(cudaMemsetAsync(bottom_diff, 0, limit * sizeof(float), stream), "roi_pool_bwd memset");
    if (num_rois <= 0) return 1;
    D2T_REQUIRE(top_diff && argmax_data, "d2t_roi_pool_backward_det: null pointer");
    const size_t total = (size_t)num_rois * channels * pooled_height * pooled_width;
    int nb = 0;
    if (!det_begin(top_diff, total, limit, scratch, scratch_bytes, stream, "d2t_roi_pool_backward_det", &nb)) return 0;
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(scratch);
    roi_pool_bwd<true><<<grid_for(total), 256, 0, stream>>>(top_diff, argmax_data, total, limit, bottom_diff, sc, nb);
    D2T_CHECK_LAUNCH("roi_pool_bwd<det>");
    det_finish<<<grid_for(limit), 256, 0, stream>>>(sc, nb, limit, bottom_diff, 0);
    D2T_CHECK_LAUNCH("det_finish");
    return 1;
}

// grad_images += the RoICrop (bilinear sampler) gradient; contiguous NCHW images, [R, gh, gw, 2] grid, R = rois of ALL images
// (R / batch_size per image, like BilinearSamplerBHWD_updateGradInput_cuda_kernel)
extern "C" int d2t_roi_crop_backward_det(const float* grids, const float* grad_output, float* grad_images, int batch_size,
                                         int channels, int height, int width, int num_rois, int grid_h, int grid_w,
                                         void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (num_rois <= 0) return 1;
    D2T_REQUIRE(grids && grad_output && grad_images, "d2t_roi_crop_backward_det: null pointer");
    D2T_REQUIRE(batch_size > 0 && num_rois % batch_size == 0, "d2t_roi_crop_backward_det: rois must divide evenly over images");
    const size_t total = (size_t)num_rois * channels * grid_h * grid_w;
    const size_t elems = (size_t)batch_size * channels * height * width;
    int nb = 0;
    if (!det_begin(grad_output, total, elems, scratch, scratch_bytes, stream, "d2t_roi_crop_backward_det", &nb)) return 0;
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(scratch);
    roi_crop_bwd<true><<<grid_for(total), 256, 0, stream>>>(grids, grad_output, grad_images, channels, height, width, num_rois,
                                                           grid_h, grid_w, num_rois / batch_size, sc, nb);
    D2T_CHECK_LAUNCH("roi_crop_bwd<det>");
    det_finish<<<grid_for(elems), 256, 0, stream>>>(sc, nb, elems, grad_images, 0);
    D2T_CHECK_LAUNCH("det_finish");
    return 1;
}

extern "C" int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                     const int height, const int width, const int channels, const int pooled_height,
                                     const int pooled_width, const float* bottom_rois, float* top_data,
                                     int* argmax_data, cudaStream_t stream) {
    if (num_rois <= 0) return 1;
    D2T_REQUIRE(bottom_data && bottom_rois && top_data, "ROIPoolForwardLaucher: null pointer");
    const size_t total = (size_t)num_rois * channels * pooled_height * pooled_width;
    roi_pool_fwd<<<grid_for(total), 256, 0, stream>>>(bottom_data, spatial_scale, num_rois, height, width, channels,
                                                     pooled_height, pooled_width, bottom_rois, top_data, argmax_data);
    D2T_CHECK_LAUNCH("roi_pool_fwd");
    return 1;
}

extern "C" int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                      const int num_rois, const int height, const int width, const int channels,
                                      const int pooled_height, const int pooled_width, const float* bottom_rois,
                                      float* bottom_diff, const int* argmax_data, cudaStream_t stream) {
    (void)spatial_scale; (void)bottom_rois;
    D2T_REQUIRE(bottom_diff, "ROIPoolBackwardLaucher: null bottom_diff");
    const size_t limit = (size_t)batch_size * channels * height * width;
    // the reference kernel assigns every element of bottom_diff (roi_pooling_kernel.cu:201)
    D2T_CUDA_OK(cudaMemsetAsync(bottom_diff, 0, limit * sizeof(float), stream), "roi_pool_bwd memset");
    if (num_rois <= 0) return 1;
    D2T_REQUIRE(top_diff && argmax_data, "ROIPoolBackwardLaucher: null pointer");
    const size_t total = (size_t)num_rois * channels * pooled_height * pooled_width;
    roi_pool_bwd<false><<<grid_for(total), 256, 0, stream>>>(top_diff, argmax_data, total, limit, bottom_diff, nullptr, 0);
    D2T_CHECK_LAUNCH("roi_pool_bwd");
    return 1;
}

static int crop_layout_ok(int C, int H, int W, int sb, int sc, int sh, int sw) {
    return sw == 1 && sh == W && sc == H * W && sb == C * H * W;
}

extern "C" int BilinearSamplerBHWD_updateOutput_cuda_kernel(int oc, int ow, int oh, int ob, int ic, int ih, int iw,
                                                            int ib, float* inputImages, int isb, int isc, int ish,
                                                            int isw, float* grids, int gsb, int gsc, int gsh, int gsw,
                                                            float* output, int osb, int osc, int osh, int osw,
                                                            cudaStream_t stream) {
    if (ob <= 0) return 1;
    D2T_REQUIRE(inputImages && grids && output, "BilinearSamplerBHWD_updateOutput: null pointer");
    D2T_REQUIRE(ib > 0 && ob % ib == 0 && oc == ic, "BilinearSamplerBHWD_updateOutput: rois must divide evenly over images");
    D2T_REQUIRE(crop_layout_ok(ic, ih, iw, isb, isc, ish, isw) && crop_layout_ok(oc, oh, ow, osb, osc, osh, osw) &&
                    gsc == 1 && gsw == 2 && gsh == 2 * ow && gsb == 2 * ow * oh,
                "BilinearSamplerBHWD_updateOutput: tensors must be contiguous (NCHW images, [R,h,w,2] grid)");
    const size_t total = (size_t)ob * oc * oh * ow;
    roi_crop_fwd<<<grid_for(total), 256, 0, stream>>>(inputImages, grids, output, ic, ih, iw, ob, oh, ow, ob / ib);
    D2T_CHECK_LAUNCH("roi_crop_fwd");
    return 1;
}

extern "C" int BilinearSamplerBHWD_updateGradInput_cuda_kernel(
    int goc, int gow, int goh, int gob, int ic, int ih, int iw, int ib, float* inputImages, int isb, int isc, int ish,
    int isw, float* grids, int gsb, int gsc, int gsh, int gsw, float* gradInputImages, int gisb, int gisc, int gish,
    int gisw, float* gradGrids, int ggsb, int ggsc, int ggsh, int ggsw, float* gradOutput, int gosb, int gosc,
    int gosh, int gosw, cudaStream_t stream) {
    (void)inputImages; (void)isb; (void)isc; (void)ish; (void)isw;
    (void)gradGrids; (void)ggsb; (void)ggsc; (void)ggsh; (void)ggsw;   // never written (kernel.cu:111-194)
    if (gob <= 0) return 1;
    D2T_REQUIRE(grids && gradInputImages && gradOutput, "BilinearSamplerBHWD_updateGradInput: null pointer");
    D2T_REQUIRE(ib > 0 && gob % ib == 0 && goc == ic, "BilinearSamplerBHWD_updateGradInput: rois must divide evenly over images");
    D2T_REQUIRE(crop_layout_ok(ic, ih, iw, gisb, gisc, gish, gisw) && crop_layout_ok(goc, goh, gow, gosb, gosc, gosh, gosw) &&
                    gsc == 1 && gsw == 2 && gsh == 2 * gow && gsb == 2 * gow * goh,
                "BilinearSamplerBHWD_updateGradInput: tensors must be contiguous");
    const size_t total = (size_t)gob * goc * goh * gow;
    roi_crop_bwd<false><<<grid_for(total), 256, 0, stream>>>(grids, gradOutput, gradInputImages, ic, ih, iw, gob, goh, gow,
                                                            gob / ib, nullptr, 0);
    D2T_CHECK_LAUNCH("roi_crop_bwd");
    return 1;
}
