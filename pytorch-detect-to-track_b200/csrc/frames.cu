// frames.cu -- frame preparation on the device (the data format ahead of the hot path):
// uint8 BGR frames as cv2.imread returns them -> mean-subtracted, bilinearly resized float32 network input.
//
// Replaces, per frame batch, the reference's host chain
//   im.astype(float32) -= PIXEL_MEANS ; cv2.resize(fx, fy, INTER_LINEAR)      lib/model/utils/blob.py:35-52
//   im[:, ::-1, :] (flipped roidb entries)                                    lib/roi_data_layer/minibatch.py:77-78
//   im_list_to_blob (zero padding to the blob's size)                         lib/model/utils/blob.py:20-33
//   .permute(0, 3, 1, 2)                                                      lib/roi_data_layer/roibatchLoader.py:183
// with one launch that reads 3 bytes per source pixel and writes the blob once; the host->device copy carries the
// uint8 frame (2.8 MB for 720x1280) instead of the float blob (7.2 MB for 600x1000).
//
// Arithmetic = OpenCV's own float32 INTER_LINEAR code (resize.cpp), restated in oracle/frames.py: coordinates in double,
// rounded to float, two fp32 passes (horizontal, then vertical), no fused multiply-adds -- bit-identical to the oracle.
// HBM-bound byte work: algorithmic bytes per frame = 3*src_h*src_w (read) + 12*blob_h*blob_w (written).
#include "common.cuh"

#include <math.h>

namespace {

struct Means { double m[3]; };

__device__ __forceinline__ void axis_coord(int d, double scale, int n_src, int& s, float& a) {
    const double pos = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);   // (d + 0.5) * scale - 0.5, as two roundings
    const float f = __double2float_rn(pos);
    const float fl = floorf(f);
    s = (int)fl;
    a = __fsub_rn(f, fl);
    if (s < 0) { s = 0; a = 0.f; }
    if (s >= n_src - 1) { s = n_src - 1; a = 0.f; }
}

// float32(double(u8) - mean) without a table and without fp64: mean = hi + lo with hi a multiple of 2^-8 (so that u8 - hi is
// exact in fp32) and lo = float(mean - hi); (u8 - hi) - lo then rounds once.  lo itself carries a 2^-33 error, so the host
// checks all 3 x 256 values against the double formula before choosing this path (means_split); otherwise the table is used.
struct MeanSplit { float hi32k[3], lo[3]; };     // 32768 + hi, lo

// One thread: kX destination columns (blockDim apart: a warp reads neighbouring source bytes and writes 128 contiguous
// bytes per plane) of one destination row, all three channels.
template <int kX, bool kLut, bool kNhwc>
__global__ void __launch_bounds__(256)
frames_prep_kernel(const uint8_t* __restrict__ frames, int src_h, int src_w, Means means, MeanSplit split, double scale,
                   int flipped, int dst_h, int dst_w, float* __restrict__ blob, int blob_h, int blob_w) {
    __shared__ float lut[kLut ? 3 : 1][256];                                // float32(double(u8) - mean[c])
    if (kLut) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x)
            lut[i >> 8][i & 255] = __double2float_rn(__dsub_rn((double)(i & 255), means.m[i >> 8]));
        __syncthreads();
    }
    const int y = blockIdx.y, n = blockIdx.z;
    const bool row_live = y < dst_h;
    int sy = 0; float b = 0.f;
    if (row_live) axis_coord(y, scale, src_h, sy, b);
    const int sy1 = min(sy + 1, src_h - 1);
    const float b0 = __fsub_rn(1.f, b);
    const uint8_t* img = frames + (size_t)n * src_h * src_w * 3;
    const uint8_t* row0 = img + (size_t)sy * src_w * 3;
    const uint8_t* row1 = img + (size_t)sy1 * src_w * 3;
    // (the int -> float conversion instruction runs on the quarter-rate XU pipe, which ncu showed 60 % busy with 12
    // conversions per pixel: 32768 + u is built with one byte permute instead -- the byte lands in mantissa bits 8..15 of
    // 0x47000000 = 32768.0f, whose ulp is 2^-8 -- and hi32k = 32768 + hi, exact because hi is a multiple of 2^-8, is
    // what gets subtracted)
    auto sample = [&](const uint8_t* __restrict__ row, unsigned off, int c) -> float {
        const unsigned u = __ldg(row + off + c);
        if (kLut) return lut[c][u];
        return __fsub_rn(__fsub_rn(__uint_as_float(__byte_perm(u, 0x47000000u, 0x7604)), split.hi32k[c]), split.lo[c]);
    };
#pragma unroll
    for (int j = 0; j < kX; ++j) {
        const int x = (blockIdx.x * kX + j) * blockDim.x + threadIdx.x;
        if (x >= blob_w) break;
        float v[3] = {0.f, 0.f, 0.f};
        if (row_live && x < dst_w) {
            int sx; float a;
            axis_coord(x, scale, src_w, sx, a);
            int sx1 = min(sx + 1, src_w - 1);
            if (flipped) { sx = src_w - 1 - sx; sx1 = src_w - 1 - sx1; }
            const float a0 = __fsub_rn(1.f, a);
            const unsigned o0 = (unsigned)sx * 3u, o1 = (unsigned)sx1 * 3u;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s00 = sample(row0, o0, c), s01 = sample(row0, o1, c);
                const float s10 = sample(row1, o0, c), s11 = sample(row1, o1, c);
                const float r0 = __fadd_rn(__fmul_rn(s00, a0), __fmul_rn(s01, a));
                const float r1 = __fadd_rn(__fmul_rn(s10, a0), __fmul_rn(s11, a));
                v[c] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b));
            }
        }
        if (kNhwc) {
            float* o = blob + (((size_t)n * blob_h + y) * blob_w + x) * 3;
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) blob[(((size_t)n * 3 + c) * blob_h + y) * blob_w + x] = v[c];
        }
    }
}

// hi / lo split of the three means; false when (u8 - hi) - lo is not float32(double(u8) - mean) for every u8
bool means_split(const double* m, MeanSplit& out) {
    for (int c = 0; c < 3; ++c) {
        const double hi = nearbyint(m[c] * 256.0) / 256.0;
        if (!(hi > -16384.0 && hi < 16384.0)) return false;                  // 32768 + hi and u - hi stay exact in fp32
        out.hi32k[c] = (float)(32768.0 + hi);
        out.lo[c] = (float)(m[c] - hi);
        for (int i = 0; i < 256; ++i) {
            volatile float t = (float)(32768 + i) - out.hi32k[c];
            volatile float got = t - out.lo[c];
            const float want = (float)((double)i - m[c]);
            if (got != want) return false;
        }
    }
    return true;
}

}  // namespace

extern "C" {

int d2t_frames_resized_shape(int src_h, int src_w, int target_size, int max_size, int cap, int* out_hw,
                                     double* im_scale) {
    D2T_REQUIRE(src_h > 0 && src_w > 0 && target_size > 0 && out_hw && im_scale, "d2t_frames_resized_shape: bad arguments");
    const int size_min = src_h < src_w ? src_h : src_w, size_max = src_h < src_w ? src_w : src_h;
    double s = (double)target_size / (double)size_min;                      // blob.py:43
    if (cap && nearbyint(s * size_max) > (double)max_size)                  // demo.py:273-274 (np.round: half to even)
        s = (double)max_size / (double)size_max;
    out_hw[0] = (int)nearbyint(src_h * s);                                  // cv::resize: saturate_cast<int>(rows * fy)
    out_hw[1] = (int)nearbyint(src_w * s);
    *im_scale = s;
    D2T_REQUIRE(out_hw[0] > 0 && out_hw[1] > 0, "d2t_frames_resized_shape: empty destination");
    return 1;
}

int d2t_frames_prep(const uint8_t* frames, int n, int src_h, int src_w, const double* pixel_means, double im_scale,
                            int flipped, int dst_h, int dst_w, float* blob, int blob_h, int blob_w, int nhwc,
                            cudaStream_t stream) {
    D2T_REQUIRE(frames && blob && pixel_means, "d2t_frames_prep: null pointer");
    D2T_REQUIRE(n > 0 && src_h > 0 && src_w > 0 && im_scale > 0.0, "d2t_frames_prep: bad source geometry");
    D2T_REQUIRE(dst_h == (int)nearbyint(src_h * im_scale) && dst_w == (int)nearbyint(src_w * im_scale),
                "d2t_frames_prep: dst %dx%d is not round(src * im_scale) = %dx%d", dst_h, dst_w,
                (int)nearbyint(src_h * im_scale), (int)nearbyint(src_w * im_scale));
    D2T_REQUIRE(dst_h > 0 && dst_w > 0 && blob_h >= dst_h && blob_w >= dst_w, "d2t_frames_prep: blob %dx%d smaller than the resized frame %dx%d",
                blob_h, blob_w, dst_h, dst_w);
    D2T_REQUIRE(blob_h <= 65535 && n <= 65535, "d2t_frames_prep: grid limits (blob_h, n <= 65535)");
    Means means{{pixel_means[0], pixel_means[1], pixel_means[2]}};
    MeanSplit split{};
    const bool lut = !means_split(pixel_means, split);
    const double scale = 1.0 / im_scale;                                    // cv::resize: scale_x = 1 / inv_scale_x
    constexpr int kX = 4;
    const int threads = blob_w >= 1024 ? 256 : (blob_w >= 512 ? 128 : 64);
    dim3 grid((blob_w + threads * kX - 1) / (threads * kX), blob_h, n);
#define D2T_FRAMES_LAUNCH(LUT, NHWC)                                                                                    \
    frames_prep_kernel<kX, LUT, NHWC><<<grid, threads, 0, stream>>>(frames, src_h, src_w, means, split, scale, flipped, \
                                                                    dst_h, dst_w, blob, blob_h, blob_w)
    if (lut) { if (nhwc) D2T_FRAMES_LAUNCH(true, true); else D2T_FRAMES_LAUNCH(true, false); }
    else     { if (nhwc) D2T_FRAMES_LAUNCH(false, true); else D2T_FRAMES_LAUNCH(false, false); }
#undef D2T_FRAMES_LAUNCH
    D2T_CHECK_LAUNCH("d2t_frames_prep");
    return 1;
}

}  // extern "C"
