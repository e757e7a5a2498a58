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

// One thread: one destination column of kRows consecutive destination rows, all three channels -- the two passes kept
// separate as in OpenCV.  The horizontal pass of a SOURCE row (6 byte loads, mean subtraction, 3 two-tap sums) is computed
// once and held in registers while consecutive destination rows need it: a reduction by 1.28 touches 1.28 new source rows
// per destination row instead of 2 (ncu of the one-pass form: issue-bound at 165 instructions per pixel, profiles/
// r02_frames_prep_ncu.txt).  Which rows are new depends on the row only, so the branches are uniform over the block; the
// column coordinate (fp64 pipe) is computed once per thread, the block's row coordinates by its first kRows threads.
// (Measured and dropped: the block's output rows staged in shared memory and written by 1-D TMA bulk stores -- same time.)
template <int kRows, bool kLut, bool kNhwc>
__global__ void __launch_bounds__(128)
frames_prep_kernel(const uint8_t* __restrict__ frames, int src_h, int src_w, Means means, MeanSplit split, double scale,
                   int flipped, int dst_h, int dst_w, float* __restrict__ blob, int blob_h, int blob_w) {
    __shared__ float lut[kLut ? 3 : 1][256];                                // float32(double(u8) - mean[c])
    __shared__ int row_s[kRows];
    __shared__ float row_b[kRows];
    if (kLut)
        for (int i = threadIdx.x; i < 768; i += blockDim.x)
            lut[i >> 8][i & 255] = __double2float_rn(__dsub_rn((double)(i & 255), means.m[i >> 8]));
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * kRows, n = blockIdx.z;
    if (threadIdx.x < kRows) {
        int sy = 0; float b = 0.f;
        if (y0 + threadIdx.x < dst_h) axis_coord(y0 + threadIdx.x, scale, src_h, sy, b);
        row_s[threadIdx.x] = sy;
        row_b[threadIdx.x] = b;
    }
    unsigned o0 = 0, o1 = 0; float a = 0.f;
    const bool col_live = x < dst_w;
    if (col_live) {
        int sx;
        axis_coord(x, scale, src_w, sx, a);
        int sx1 = min(sx + 1, src_w - 1);
        if (flipped) { sx = src_w - 1 - sx; sx1 = src_w - 1 - sx1; }
        o0 = (unsigned)sx * 3u; o1 = (unsigned)sx1 * 3u;
    }
    const float a0 = __fsub_rn(1.f, a);
    __syncthreads();
    const bool active = x < blob_w;
    const uint8_t* __restrict__ img = frames + (size_t)n * src_h * src_w * 3;
    // (the int -> float conversion instruction runs on the quarter-rate XU pipe: 32768 + u is built with one byte permute
    // instead -- the byte lands in mantissa bits 8..15 of 0x47000000 = 32768.0f, whose ulp is 2^-8 -- and hi32k =
    // 32768 + hi, exact because hi is a multiple of 2^-8, is what gets subtracted)
    auto sample = [&](const uint8_t* __restrict__ row, unsigned off, int c) -> float {
        const unsigned u = __ldg(row + off + c);
        if (kLut) return lut[c][u];
        return __fsub_rn(__fsub_rn(__uint_as_float(__byte_perm(u, 0x47000000u, 0x7604)), split.hi32k[c]), split.lo[c]);
    };
    // address arithmetic in 32 bits (d2t_frames_prep checks that a frame and a blob image stay below 2^31 bytes /
    // elements): 64-bit multiplies per row and channel were two thirds of the instructions ncu counted
    const unsigned row_bytes = (unsigned)src_w * 3u;
    auto hpass = [&](int sy, float (&h)[3]) {
        const unsigned q0 = (unsigned)sy * row_bytes + o0, q1 = (unsigned)sy * row_bytes + o1;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            h[c] = __fadd_rn(__fmul_rn(sample(img, q0, c), a0), __fmul_rn(sample(img, q1, c), a));
    };
    const int rows = min(kRows, blob_h - y0);
    // The rows of the loop below are a dependent chain (load -> blend -> store, row after row): without help a thread
    // waits out one cold DRAM access per new source row, ~11 in sequence.  All source rows of the tile are requested up
    // front instead (both taps of a column share a sector, or neighbouring ones).
    if (col_live && y0 < dst_h) {
        const int first = row_s[0], last = min(row_s[min(rows, dst_h - y0) - 1] + 1, src_h - 1);
        for (int sy = first; sy <= last; ++sy)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(img + (unsigned)sy * row_bytes + min(o0, o1)));
    }
    float h0[3] = {0.f, 0.f, 0.f}, h1[3] = {0.f, 0.f, 0.f};
    int have0 = -1, have1 = -1;                                             // source rows held in h0 / h1
    const unsigned plane = (unsigned)blob_h * (unsigned)blob_w;
    float* __restrict__ out = kNhwc ? blob + ((size_t)n * plane + (size_t)y0 * blob_w + x) * 3
                                    : blob + (size_t)n * 3 * plane + (size_t)y0 * blob_w + x;
    for (int r = 0; r < rows; ++r) {
        float v[3] = {0.f, 0.f, 0.f};
        if (col_live && y0 + r < dst_h) {
            const int sy = row_s[r], sy1 = min(sy + 1, src_h - 1);
            const float b = row_b[r], b0 = __fsub_rn(1.f, b);
            if (sy != have0) {
                if (sy == have1) { h0[0] = h1[0]; h0[1] = h1[1]; h0[2] = h1[2]; }
                else hpass(sy, h0);
                have0 = sy;
            }
            if (sy1 != have1) {
                if (sy1 == have0) { h1[0] = h0[0]; h1[1] = h0[1]; h1[2] = h0[2]; }
                else hpass(sy1, h1);
                have1 = sy1;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = __fadd_rn(__fmul_rn(h0[c], b0), __fmul_rn(h1[c], b));
        }
        if (active) {
            if (kNhwc) {
                out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
                out += (size_t)blob_w * 3;
            } else {
                out[0] = v[0]; out[plane] = v[1]; out[2u * plane] = v[2];
                out += blob_w;
            }
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
    D2T_REQUIRE((double)src_h * src_w * 3 < 2147483648.0 && (double)blob_h * blob_w * 3 < 2147483648.0,
                "d2t_frames_prep: a frame / a blob image must stay below 2^31 bytes / elements");
    D2T_REQUIRE(blob_h <= 65535 * 8 && n <= 65535, "d2t_frames_prep: grid limits (blob_h <= 524280, n <= 65535)");
    Means means{{pixel_means[0], pixel_means[1], pixel_means[2]}};
    MeanSplit split{};
    const bool lut = !means_split(pixel_means, split);
    const double scale = 1.0 / im_scale;                                    // cv::resize: scale_x = 1 / inv_scale_x
    constexpr int kRows = 8;
    const int threads = blob_w >= 128 ? 128 : 64;
    dim3 grid((blob_w + threads - 1) / threads, (blob_h + kRows - 1) / kRows, n);
#define D2T_FRAMES_LAUNCH(LUT, NHWC)                                                                                  \
    frames_prep_kernel<kRows, LUT, NHWC><<<grid, threads, 0, stream>>>(frames, src_h, src_w, means, split, scale,     \
                                                                       flipped, dst_h, dst_w, blob, blob_h, blob_w)
    if (lut) { if (nhwc) D2T_FRAMES_LAUNCH(true, true); else D2T_FRAMES_LAUNCH(true, false); }
    else     { if (nhwc) D2T_FRAMES_LAUNCH(false, true); else D2T_FRAMES_LAUNCH(false, false); }
#undef D2T_FRAMES_LAUNCH
    D2T_CHECK_LAUNCH("d2t_frames_prep");
    return 1;
}

}  // extern "C"
