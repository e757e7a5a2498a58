// conv_util.cu -- layout helpers around the tcgen05 conv kernels (conv.cu).
//
// The conv engine keeps activations as plain fp32 NHWC tensors; these kernels move data between that layout and
// the reference's NCHW tensors, pack OIHW weights into the [Cout, R*S*Cin_pad] K-major (w, w_lo) matrices the
// weight TMA reads (w_lo = w - trunc13(w): the exact remainder of the tensor core's own TF32 truncation, see
// conv.cu), and implement the stem's MaxPool2d(3, stride 2, padding 0, ceil_mode=True)
// (/root/reference/lib/model/faster_rcnn/resnet.py:120).
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"

namespace d2t {
namespace {

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

inline int grid_for(size_t total, int per_block = 256) {
    size_t blocks = (total + per_block - 1) / per_block, cap = (size_t)sm_count() * 16;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

// image [N,C<=4,H,W] -> zero-bordered NHWC4 [N, Hp, Wp, 4] for the stem (border 3, unused channels 0);
// one thread per padded pixel, every element of the buffer is written.
__global__ void stem_pack_input(const float* __restrict__ x, int N, int C, int H, int W, int Hp, int Wp,
                                float4* __restrict__ out) {
    const size_t total = (size_t)N * Hp * Wp;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int wp = (int)(idx % Wp), hp = (int)((idx / Wp) % Hp), n = (int)(idx / Wp / Hp);
        const int h = hp - 3, w = wp - 3;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (h >= 0 && h < H && w >= 0 && w < W)
            for (int c = 0; c < C && c < 4; ++c) v[c] = __ldg(x + (((size_t)n * C + c) * H + h) * W + w);
        out[idx] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// the same with max |x| folded into *amax (the 3xFP16 stem's operand scale)
__global__ void __launch_bounds__(256)
stem_pack_input_amax(const float* __restrict__ x, int N, int C, int H, int W, int Hp, int Wp, float4* __restrict__ out,
                     float* __restrict__ amax, int pairs) {
    const size_t total = (size_t)N * Hp * Wp;
    float m = 0.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int wp = (int)(idx % Wp), hp = (int)((idx / Wp) % Hp), n = (int)(idx / Wp / Hp);
        const int h = hp - 3, w = wp - 3;
        // pairs > 0: x is [pairs][2 legs] frames (the reference's im_data), the packed batch is leg-major -- the permute
        // the engine used to run as a copy of its own
        const int ns = pairs > 0 ? (n % pairs) * 2 + n / pairs : n;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (h >= 0 && h < H && w >= 0 && w < W)
            for (int c = 0; c < C && c < 4; ++c) v[c] = __ldg(x + (((size_t)ns * C + c) * H + h) * W + w);
        out[idx] = make_float4(v[0], v[1], v[2], v[3]);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));
    }
    __shared__ unsigned int smax;
    if (threadIdx.x == 0) smax = 0u;
    __syncthreads();
    const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && wm) atomicMax(&smax, wm);
    __syncthreads();
    if (threadIdx.x == 0 && smax) atomicMax(reinterpret_cast<unsigned int*>(amax), smax);
}

// stem weights [O, C<=4, 7, 7] -> [O][7 rows][32] with column s*4 + c (zeros elsewhere)
__global__ void stem_pack_weights(const float* __restrict__ w, int O, int C, float* __restrict__ hi, float* __restrict__ lo) {
    const int total = O * 7 * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int col = idx % 32, r = (idx / 32) % 7, o = idx / 32 / 7;
        const int s = col >> 2, c = col & 3;
        const float v = (s < 7 && c < C) ? __ldg(w + (((size_t)o * C + c) * 7 + r) * 7 + s) : 0.f;
        hi[idx] = v;
        if (lo) lo[idx] = v - tf32_hi(v);
    }
}

// [N,C,H,W] -> channels [coff, coff + cw) of [N,H,W,cs]: x's C channels, then zeros.  32x32 smem transpose per (n, h).
__global__ void __launch_bounds__(256)
nchw_to_nhwc(const float* __restrict__ x, int N, int C, int H, int W, int cs, int coff, int cw, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int wt = blockIdx.x * 32, ct = blockIdx.y * 32, nh = blockIdx.z, n = nh / H, h = nh % H;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = ct + j, w = wt + tx;
        tile[j][tx] = (c < C && w < W) ? __ldg(x + (((size_t)n * C + c) * H + h) * W + w) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int w = wt + j, c = ct + tx;
        if (w < W && c < cw) out[(((size_t)n * H + h) * W + w) * cs + coff + c] = tile[tx][j];
    }
}

// the same with the tensor's max |x| folded into *amax (one atomicMax per block; values >= 0: unsigned order is float order)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_amax(const float* __restrict__ x, int N, int C, int H, int W, int cs, int coff, int cw, float* __restrict__ out,
                  float* __restrict__ amax) {
    __shared__ float tile[32][33];
    __shared__ unsigned int smax;
    const int wt = blockIdx.x * 32, ct = blockIdx.y * 32, nh = blockIdx.z, n = nh / H, h = nh % H;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    if (threadIdx.x == 0) smax = 0u;
    float m = 0.f;
    for (int j = ty; j < 32; j += 8) {
        const int c = ct + j, w = wt + tx;
        const float v = (c < C && w < W) ? __ldg(x + (((size_t)n * C + c) * H + h) * W + w) : 0.f;
        tile[j][tx] = v;
        m = fmaxf(m, fabsf(v));
    }
    const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    __syncthreads();
    if (tx == 0 && wm) atomicMax(&smax, wm);
    for (int j = ty; j < 32; j += 8) {
        const int w = wt + j, c = ct + tx;
        if (w < W && c < cw) out[(((size_t)n * H + h) * W + w) * cs + coff + c] = tile[tx][j];
    }
    __syncthreads();
    if (threadIdx.x == 0 && smax) atomicMax(reinterpret_cast<unsigned int*>(amax), smax);
}

// channels [coff, coff + C) of [N,H,W,cs] -> [N,C,H,W]
__global__ void __launch_bounds__(256)
nhwc_to_nchw(const float* __restrict__ x, int N, int C, int H, int W, int cs, int coff, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int wt = blockIdx.x * 32, ct = blockIdx.y * 32, nh = blockIdx.z, n = nh / H, h = nh % H;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int w = wt + j, c = ct + tx;
        tile[j][tx] = (w < W && c < C) ? __ldg(x + (((size_t)n * H + h) * W + w) * cs + coff + c) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = ct + j, w = wt + tx;
        if (c < C && w < W) out[(((size_t)n * C + c) * H + h) * W + w] = tile[tx][j];
    }
}

// OIHW -> [O][R*S][cin_pad] (w, w_lo), zero pad
__global__ void pack_weights(const float* __restrict__ w, int O, int I, int R, int S, int cin_pad,
                             float* __restrict__ hi, float* __restrict__ lo) {
    const size_t total = (size_t)O * R * S * cin_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cin_pad);
        const int rs = (int)((idx / cin_pad) % (R * S));
        const int o = (int)(idx / cin_pad / (R * S));
        const float v = c < I ? __ldg(w + ((size_t)o * I + c) * R * S + rs) : 0.f;
        hi[idx] = v;
        if (lo) lo[idx] = v - tf32_hi(v);
    }
}

// OIHW -> [O][R*S][cin_pad] fp16 (hi, lo) of w * 2^w_exp (the 3xFP16 operand format, conv.cu), zero pad
__global__ void pack_weights_f16(const float* __restrict__ w, const float* __restrict__ scale, int O, int I, int R, int S,
                                 int cin_pad, float sw, __half* __restrict__ hi, __half* __restrict__ lo) {
    const size_t total = (size_t)O * R * S * cin_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cin_pad);
        const int rs = (int)((idx / cin_pad) % (R * S));
        const int o = (int)(idx / cin_pad / (R * S));
        float v = c < I ? __ldg(w + ((size_t)o * I + c) * R * S + rs) * sw : 0.f;
        if (scale) v *= __ldg(scale + o);
        const __half h = __float2half_rn(v);
        hi[idx] = h;
        lo[idx] = __float2half_rn(v - __half2float(h));
    }
}

// MaxPool 3x3 stride 2, padding 0, ceil_mode (windows clipped at the border), NHWC, 4 channels per thread
__global__ void maxpool3x3s2_nhwc(const float* __restrict__ in, int N, int H, int W, int C, int OH, int OW,
                                  float* __restrict__ out) {
    const int C4 = C >> 2;
    const size_t total = (size_t)N * OH * OW * C4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % C4);
        const int ow = (int)((idx / C4) % OW);
        const int oh = (int)((idx / C4 / OW) % OH);
        const int n = (int)(idx / C4 / OW / OH);
        float4 m = make_float4(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
        for (int r = 0; r < 3; ++r) {
            const int h = oh * 2 + r;
            if (h >= H) break;
            for (int s = 0; s < 3; ++s) {
                const int w = ow * 2 + s;
                if (w >= W) break;
                const float4 v = __ldg(reinterpret_cast<const float4*>(in) + ((((size_t)n * H + h) * W + w) * C >> 2) + c4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        reinterpret_cast<float4*>(out)[idx] = m;
    }
}


// ---------------------------------------------------------------- training-path helpers (backward-data / weight-gradient)
// OIHW -> forward layout [O][R*S][cin_pad] fp16 (hi, lo) of w * 2^k, k derived on the device from *amax (= max |w|, kept
// by the caller): the per-step re-pack of a weight that the optimizer has just changed needs no host round trip
__global__ void pack_weights_f16_dev(const float* __restrict__ w, const float* __restrict__ scale, int O, int I, int R, int S,
                                     int cin_pad, const float* __restrict__ amax, __half* __restrict__ hi,
                                     __half* __restrict__ lo) {
    const float sw = pow2f(act_exp(amax));
    const size_t total = (size_t)O * R * S * cin_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cin_pad);
        const int rs = (int)((idx / cin_pad) % (R * S));
        const int o = (int)(idx / cin_pad / (R * S));
        float v = c < I ? __ldg(w + ((size_t)o * I + c) * R * S + rs) * sw : 0.f;
        if (scale) v *= __ldg(scale + o);
        const __half h = __float2half_rn(v);
        hi[idx] = h;
        lo[idx] = __float2half_rn(v - __half2float(h));
    }
}

// OIHW -> the backward-data operand [I][R*S][cout_pad]: the transposed, spatially flipped filter with the folded
// BatchNorm scale of the OUTPUT channel multiplied in,  wt[ci][r'][s'][co] = w[co][ci][R-1-r'][S-1-s'] * scale[co],
// fp16 (hi, lo) of wt * 2^k with k from *amax (an upper bound of max |wt|)
__global__ void pack_weights_f16_dgrad(const float* __restrict__ w, const float* __restrict__ scale, int O, int I, int rows,
                                       int R, int S, int cout_pad, const float* __restrict__ amax,
                                       __half* __restrict__ hi, __half* __restrict__ lo) {
    const float sw = pow2f(act_exp(amax));
    const size_t total = (size_t)rows * R * S * cout_pad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx % cout_pad);
        const int rs = (int)((idx / cout_pad) % (R * S));
        const int i = (int)(idx / cout_pad / (R * S));
        float v = 0.f;
        if (o < O && i < I) {
            v = __ldg(w + ((size_t)o * I + i) * R * S + (R * S - 1 - rs)) * sw;
            if (scale) v *= __ldg(scale + o);
        }
        const __half h = __float2half_rn(v);
        hi[idx] = h;
        lo[idx] = __float2half_rn(v - __half2float(h));
    }
}

// ONE launch for every packed operand of a training engine (d2t_conv_repack_many): after an optimizer step a Res-101 D&T
// engine re-packs ~330 forward operands and ~300 backward-data operands; as one small kernel each they are launch-bound
// (1.1 ms per step for 0.5 GB of traffic).  Block b belongs to the item whose [first_block, first_block + blocks) contains
// it (binary search over the items' first blocks) and packs 1024 consecutive output elements.
struct RepackItem {
    const float* w;        // OIHW weights
    const float* scale;    // [O] or null
    const float* amax;     // device scalar: max |w * scale| bound the 2^k is derived from
    __half* hi;
    __half* lo;
    int O, I, R, S;
    int pad;               // forward: cin_pad; backward-data: cout_pad
    int rows;              // backward-data: rows of the transposed operand (>= I); forward: unused
    int dgrad;             // 0: [O][R*S][cin_pad] of w;  1: [rows][R*S][cout_pad] of the flipped transpose
    int first_block;
};

// A block re-packs one tile of 32 output channels x 32 input channels x all R*S taps through shared memory: the OIHW source
// is read in runs of 32 * R*S consecutive floats (one output channel's slice), the packed operand -- [O][R*S][cin_pad] for
// the forward convolution, [rows][R*S][cout_pad] of the flipped transpose for backward-data -- is written in runs of 32
// consecutive fp16 values.  (Element-by-element, the backward-data operand read its source with a stride of I*R*S floats
// and every index took three 64-bit divisions: 0.65 ms per step for 0.4 GB of traffic.)  Taps: R*S <= kRepackMaxRS.
constexpr int kRepackMaxRS = 9;
__host__ __device__ inline int repack_item_blocks(int O, int I, int RS, int pad, int rows, int dgrad) {
    (void)RS;
    return dgrad ? ((rows + 31) / 32) * ((pad + 31) / 32) : ((O + 31) / 32) * ((pad + 31) / 32);
}

// (RS_ is a compile-time constant for the two filter sizes of the network -- the divisions by R*S below were most of the
// kernel's instructions when it was a run-time value -- 0 = any R*S <= kRepackMaxRS)
template <int RS_>
__device__ __forceinline__ void repack_tile(const RepackItem& it, int tb, float sw, float (*tile)[32 * kRepackMaxRS + 1]) {
    const int RS = RS_ ? RS_ : it.R * it.S, tid = threadIdx.x;
    const int tiles_c = (it.pad + 31) / 32;                // tiles along the packed operand's channel (fastest) axis
    const int tr = tb / tiles_c, tc = tb - tr * tiles_c;
    // forward: rows of the operand = output channels o, its channel axis = input channels i; backward-data: rows = i, channels = o
    const int o0 = (it.dgrad ? tc : tr) * 32, i0 = (it.dgrad ? tr : tc) * 32;
    const int run = 32 * RS;                               // floats of one output channel's slice in the tile
    for (int e = tid; e < 32 * run; e += 256) {
        const int ol = e / run, j = e - ol * run;
        const int o = o0 + ol, i = i0 + j / RS;
        float v = 0.f;
        if (o < it.O && i < it.I) {
            v = __ldg(it.w + ((size_t)o * it.I + i0) * RS + j) * sw;
            if (it.scale) v *= __ldg(it.scale + o);
        }
        tile[ol][j] = v;
    }
    __syncthreads();
    for (int e = tid; e < 32 * run; e += 256) {
        // 32 consecutive values of the packed operand's channel axis per run
        const int cl = e & 31, rest = e >> 5, rl = rest / RS, rs = rest - rl * RS;
        float v;
        size_t idx;
        if (!it.dgrad) {
            const int o = o0 + rl, c = i0 + cl;
            if (o >= it.O || c >= it.pad) continue;
            v = tile[rl][cl * RS + rs];
            idx = ((size_t)o * RS + rs) * it.pad + c;
        } else {
            const int r0 = i0 + rl, c = o0 + cl;
            if (r0 >= it.rows || c >= it.pad) continue;
            v = tile[cl][rl * RS + (RS - 1 - rs)];
            idx = ((size_t)r0 * RS + rs) * it.pad + c;
        }
        const __half h = __float2half_rn(v);
        it.hi[idx] = h;
        it.lo[idx] = __float2half_rn(v - __half2float(h));
    }
}

__global__ void __launch_bounds__(256) repack_many(const RepackItem* __restrict__ items, int n_items,
                                                   const int* __restrict__ block_item) {
    __shared__ float tile[32][32 * kRepackMaxRS + 1];
    int lo_i = 0, hi_i = n_items - 1;
    const int b = (int)blockIdx.x;
    if (block_item) {                                      // the block's item, tabulated by the host: ONE load instead of the
        lo_i = hi_i = __ldg(block_item + b);               // ~10 dependent ones of the search (most of a small block's life)
    }
    while (lo_i < hi_i) {                                  // last item with first_block <= b
        const int mid = (lo_i + hi_i + 1) >> 1;
        if (__ldg(&items[mid].first_block) <= b) lo_i = mid;
        else hi_i = mid - 1;
    }
    const RepackItem it = items[lo_i];
    const float sw = pow2f(act_exp(it.amax));
    const int RS = it.R * it.S, tb = b - it.first_block;
    if (RS == 1) repack_tile<1>(it, tb, sw, tile);
    else if (RS == 9) repack_tile<9>(it, tb, sw, tile);
    else repack_tile<0>(it, tb, sw, tile);
}

// channels [0, C) of NHWC [N, H, W, cs], sampled at (oy * stride, ox * stride), -> fp32 planes [N][C][OH][pitch]
// (columns [OW, pitch) zero): the K-contiguous A operand of the weight-gradient GEMM.  A CTA transposes 64 positions x
// 32 channels of one output row through shared memory: 128-byte loads (32 channels of a pixel), 256-byte stores (a lane
// writes two neighbouring positions of one channel).  pitch is even.
__global__ void __launch_bounds__(256)
nhwc_to_planes(const float* __restrict__ x, int N, int H, int W, int cs, int C, int stride, int OH, int OW, int pitch,
               float* __restrict__ out) {
    __shared__ float tile[64][33];
    const int wt = blockIdx.x * 64, ct = blockIdx.y * 32, nh = blockIdx.z, n = nh / OH, oy = nh % OH;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // programmatic dependent launch on both sides: this grid's blocks are scheduled while the kernel before it drains and
    // wait here for its results; the kernel after it (the weight-gradient GEMM, which waits for THIS grid's completion before
    // it reads the planes) may set up its CTAs as soon as SMs free up
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
#pragma unroll
    for (int j = ty; j < 64; j += 8) {
        const int xo = wt + j, c = ct + tx;
        tile[j][tx] = (xo < OW && c < C) ? __ldg(x + (((size_t)n * H + (size_t)oy * stride) * W + (size_t)xo * stride) * cs + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int c = ct + j, xo = wt + 2 * tx;
        if (c < C && xo < pitch)
            *reinterpret_cast<float2*>(out + (((size_t)n * C + c) * OH + oy) * pitch + xo) = make_float2(tile[2 * tx][j], tile[2 * tx + 1][j]);
    }
}

// gradient NHWC [N, OH, OW, cs] -> the B operand of the weight-gradient GEMM: S copies of fp16 (hi, lo) planes
// [S][N][C][OH][pitch] of g * 2^k (k from *amax), copy s shifted by dx_s = s * dil - pad columns:
// plane_s[u] = g[u - dx_s] (zero outside [0, OW)) -- the filter column's offset cannot be a TMA start coordinate (see
// conv.cu, WGRAD), so it is materialised here.  The 64-position tile is loaded with an 8-column halo on both sides
// (|dx| <= 8); a lane writes two neighbouring positions (half2: 128-byte stores per warp).  Also used, with a source
// stride, for the correlation backward's planes (S = 1).
__global__ void __launch_bounds__(256)
nhwc_to_planes_split(const float* __restrict__ g, int N, int Hs, int Ws, int stride, int OH, int OW, int cs, int C, int pitch,
                     int S, int dil, int pad, const float* __restrict__ amax, __half* __restrict__ out_hi,
                     __half* __restrict__ out_lo) {
    __shared__ float tile[80][33];
    const int wt = blockIdx.x * 64, ct = blockIdx.y * 32, nh = blockIdx.z, n = nh / OH, oy = nh % OH;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    asm volatile("griddepcontrol.wait;" ::: "memory");           // (see nhwc_to_planes)
    asm volatile("griddepcontrol.launch_dependents;");
    const float sa = pow2f(act_exp(amax));
#pragma unroll
    for (int j = ty; j < 80; j += 8) {
        const int xo = wt - 8 + j, c = ct + tx;
        tile[j][tx] = (xo >= 0 && xo < OW && c < C)
                          ? __ldg(g + (((size_t)n * Hs + (size_t)oy * stride) * Ws + (size_t)xo * stride) * cs + c) * sa : 0.f;
    }
    __syncthreads();
    const size_t copy = (size_t)N * C * OH * pitch;
    for (int s = 0; s < S; ++s) {
        const int dx = s * dil - pad;
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int c = ct + j, u = wt + 2 * tx;
            if (c < C && u < pitch) {
                const float v0 = tile[2 * tx + 8 - dx][j], v1 = tile[2 * tx + 9 - dx][j];
                const size_t o = s * copy + (((size_t)n * C + c) * OH + oy) * pitch + u;
                const __half2 h = __floats2half2_rn(v0, v1);
                const float2 hf = __half22float2(h);
                *reinterpret_cast<__half2*>(out_hi + o) = h;
                *reinterpret_cast<__half2*>(out_lo + o) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            }
        }
    }
}

// correlation backward, A operand: the gradient of the correlation output gO [N, H, W, D*D] (channels [coff, coff + D*D)
// of an NHWC buffer), expanded to [N][H][W][D][64] fp16 (hi, lo) of gO * 2^k:  row (pixel, tj + r) holds the D values
// gO[pixel, (tj + r) D + ti + r] at halo columns hc = (x mod 32) + 16 + ti, zeros elsewhere -- the position the other
// frame's pixel (x + ti) has inside the 64-column halo row [x0 - 16, x0 + 48) of the pixel's 32-wide tile, so that a
// tile row's slice of the band matrix is one TMA box (conv.cu, CORRB).  flipped: the band of the gradient w.r.t. the
// SECOND frame, E[p, t] = gO[p + t, -t] (zero where p + t leaves the map).  One warp per row, one half2 per lane.
__global__ void __launch_bounds__(256)
corr_band_pack(const float* __restrict__ g, int N, int H, int W, int cs, int coff, int r, int flipped,
               const float* __restrict__ amax, __half2* __restrict__ e_hi, __half2* __restrict__ e_lo) {
    const int D = 2 * r + 1;
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (size_t)N * H * W * D) return;
    const int lane = threadIdx.x & 31;
    const int tjr = (int)(row % D);
    const size_t pix = row / D;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / W / H);
    const float sa = pow2f(act_exp(amax));
    const int hc0 = (x & 31) + 16 - r;                 // halo column of ti = -r
    float v[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int tir = 2 * lane + q - hc0;
        float val = 0.f;
        if (tir >= 0 && tir < D) {
            if (!flipped) {
                val = __ldg(g + (size_t)pix * cs + coff + tjr * D + tir);
            } else {
                const int ys = y + tjr - r, xs = x + tir - r;          // source pixel p + t
                if (ys >= 0 && ys < H && xs >= 0 && xs < W)
                    val = __ldg(g + (((size_t)n * H + ys) * W + xs) * cs + coff + (2 * r - tjr) * D + (2 * r - tir));
            }
        }
        v[q] = val * sa;
    }
    const __half2 h = __floats2half2_rn(v[0], v[1]);
    const float2 hf = __half22float2(h);
    e_hi[row * 32 + lane] = h;
    e_lo[row * 32 + lane] = __floats2half2_rn(v[0] - hf.x, v[1] - hf.y);
}

// backward of a stride-2 1x1 convolution's input + ReLU:  out[n, y, x, :] = mask > 0 ? (even(y, x) ? low[n, y/2, x/2, :] : 0)
// + extra[n, y, x, :] : 0, all NHWC with the same channel stride; max |out| folded into *amax
__global__ void upsample2_add_mask(const float4* __restrict__ low, int LH, int LW, const float4* __restrict__ extra,
                                   const float4* __restrict__ mask, int N, int H, int W, int C4, float4* __restrict__ out,
                                   float* __restrict__ amax) {
    const size_t total = (size_t)N * H * W * C4;
    float m = 0.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % C4);
        const int x = (int)((idx / C4) % W), y = (int)((idx / C4 / W) % H), n = (int)(idx / C4 / W / H);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!(x & 1) && !(y & 1) && (y >> 1) < LH && (x >> 1) < LW)
            v = __ldg(low + (((size_t)n * LH + (y >> 1)) * LW + (x >> 1)) * C4 + c4);
        if (extra) {
            const float4 e = __ldg(extra + idx);
            v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        }
        if (mask) {
            const float4 k = __ldg(mask + idx);
            v.x = k.x > 0.f ? v.x : 0.f; v.y = k.y > 0.f ? v.y : 0.f; v.z = k.z > 0.f ? v.z : 0.f; v.w = k.w > 0.f ? v.w : 0.f;
        }
        out[idx] = v;
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (amax) {
        const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        if ((threadIdx.x & 31) == 0 && wm != 0u) atomicMax(reinterpret_cast<unsigned int*>(amax), wm);
    }
}

}  // namespace
}  // namespace d2t

using namespace d2t;

extern "C" int d2t_nchw_to_nhwc(const float* x, int N, int C, int H, int W, int c_stride, int c_offset, int c_width,
                                float* out, cudaStream_t stream) {
    D2T_REQUIRE(x && out && N > 0 && C > 0 && H > 0 && W > 0 && c_offset >= 0 && c_width >= C &&
                    c_stride >= c_offset + c_width,
                "d2t_nchw_to_nhwc: bad arguments");
    dim3 grid((W + 31) / 32, (c_width + 31) / 32, N * H);
    D2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "d2t_nchw_to_nhwc: tensor too large for the launch grid");
    nchw_to_nhwc<<<grid, 256, 0, stream>>>(x, N, C, H, W, c_stride, c_offset, c_width, out);
    D2T_CHECK_LAUNCH("nchw_to_nhwc");
    return 1;
}

extern "C" int d2t_nchw_to_nhwc_amax(const float* x, int N, int C, int H, int W, int c_stride, int c_offset, int c_width,
                                     float* out, float* amax, cudaStream_t stream) {
    D2T_REQUIRE(x && out && amax && N > 0 && C > 0 && H > 0 && W > 0 && c_offset >= 0 && c_width >= C &&
                    c_stride >= c_offset + c_width,
                "d2t_nchw_to_nhwc_amax: bad arguments");
    dim3 grid((W + 31) / 32, (c_width + 31) / 32, N * H);
    D2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "d2t_nchw_to_nhwc_amax: tensor too large for the launch grid");
    nchw_to_nhwc_amax<<<grid, 256, 0, stream>>>(x, N, C, H, W, c_stride, c_offset, c_width, out, amax);
    D2T_CHECK_LAUNCH("nchw_to_nhwc_amax");
    return 1;
}

extern "C" int d2t_nhwc_to_nchw(const float* x, int N, int C, int H, int W, int c_stride, int c_offset, float* out,
                                cudaStream_t stream) {
    D2T_REQUIRE(x && out && N > 0 && C > 0 && H > 0 && W > 0 && c_stride >= c_offset + C, "d2t_nhwc_to_nchw: bad arguments");
    dim3 grid((W + 31) / 32, (C + 31) / 32, N * H);
    D2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "d2t_nhwc_to_nchw: tensor too large for the launch grid");
    nhwc_to_nchw<<<grid, 256, 0, stream>>>(x, N, C, H, W, c_stride, c_offset, out);
    D2T_CHECK_LAUNCH("nhwc_to_nchw");
    return 1;
}

extern "C" int d2t_conv_pack_weights(const float* w_oihw, int Cout, int Cin, int R, int S, int cin_pad, float* w_hi,
                                     float* w_lo, cudaStream_t stream) {
    D2T_REQUIRE(w_oihw && w_hi && Cout > 0 && Cin > 0 && R > 0 && S > 0 && cin_pad >= Cin,
                "d2t_conv_pack_weights: bad arguments");
    const size_t total = (size_t)Cout * R * S * cin_pad;
    pack_weights<<<grid_for(total), 256, 0, stream>>>(w_oihw, Cout, Cin, R, S, cin_pad, w_hi, w_lo);
    D2T_CHECK_LAUNCH("pack_weights");
    return 1;
}

extern "C" int d2t_conv_pack_weights_f16(const float* w_oihw, const float* scale, int Cout, int Cin, int R, int S, int cin_pad,
                                         int w_exp, void* w_hi, void* w_lo, cudaStream_t stream) {
    D2T_REQUIRE(w_oihw && w_hi && w_lo && Cout > 0 && Cin > 0 && R > 0 && S > 0 && cin_pad >= Cin && cin_pad % 64 == 0 &&
                    w_exp >= -126 && w_exp <= 127,
                "d2t_conv_pack_weights_f16: bad arguments");
    const size_t total = (size_t)Cout * R * S * cin_pad;
    pack_weights_f16<<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, Cout, Cin, R, S, cin_pad, ldexpf(1.f, w_exp),
                                                          reinterpret_cast<__half*>(w_hi), reinterpret_cast<__half*>(w_lo));
    D2T_CHECK_LAUNCH("pack_weights_f16");
    return 1;
}


extern "C" int d2t_conv_pack_weights_f16_dev(const float* w_oihw, const float* scale, int Cout, int Cin, int R, int S,
                                             int cin_pad, const float* amax_w, void* w_hi, void* w_lo, cudaStream_t stream) {
    D2T_REQUIRE(w_oihw && w_hi && w_lo && amax_w && Cout > 0 && Cin > 0 && R > 0 && S > 0 && cin_pad >= Cin && cin_pad % 64 == 0,
                "d2t_conv_pack_weights_f16_dev: bad arguments");
    const size_t total = (size_t)Cout * R * S * cin_pad;
    pack_weights_f16_dev<<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, Cout, Cin, R, S, cin_pad, amax_w,
                                                              reinterpret_cast<__half*>(w_hi), reinterpret_cast<__half*>(w_lo));
    D2T_CHECK_LAUNCH("pack_weights_f16_dev");
    return 1;
}

extern "C" int d2t_conv_pack_weights_f16_dgrad(const float* w_oihw, const float* scale, int Cout, int Cin, int rows, int R,
                                               int S, int cout_pad, const float* amax_wt, void* wt_hi, void* wt_lo,
                                               cudaStream_t stream) {
    D2T_REQUIRE(w_oihw && wt_hi && wt_lo && amax_wt && Cout > 0 && Cin > 0 && rows >= Cin && R > 0 && S > 0 &&
                    cout_pad >= Cout && cout_pad % 64 == 0,
                "d2t_conv_pack_weights_f16_dgrad: bad arguments");
    const size_t total = (size_t)rows * R * S * cout_pad;
    pack_weights_f16_dgrad<<<grid_for(total), 256, 0, stream>>>(w_oihw, scale, Cout, Cin, rows, R, S, cout_pad, amax_wt,
                                                                reinterpret_cast<__half*>(wt_hi), reinterpret_cast<__half*>(wt_lo));
    D2T_CHECK_LAUNCH("pack_weights_f16_dgrad");
    return 1;
}

// items: device array of n_items descriptors, 8 x 8-byte words each (see d2t_b200.h); total_blocks = sum of the items' blocks
extern "C" size_t d2t_conv_repack_item_bytes(void) { return sizeof(RepackItem); }

// blocks of d2t_conv_repack_many that one item takes (its first_block is the running sum of these); 0: the item is not supported
extern "C" int d2t_conv_repack_item_blocks(int O, int I, int R, int S, int pad, int rows, int dgrad) {
    if (O <= 0 || I <= 0 || R <= 0 || S <= 0 || R * S > kRepackMaxRS || pad <= 0) return 0;
    return repack_item_blocks(O, I, R * S, pad, rows, dgrad);
}

extern "C" int d2t_conv_repack_many(const void* items, int n_items, int total_blocks, const int* block_item,
                                    cudaStream_t stream) {
    D2T_REQUIRE(items && n_items > 0 && total_blocks > 0, "d2t_conv_repack_many: bad arguments");
    repack_many<<<total_blocks, 256, 0, stream>>>(reinterpret_cast<const RepackItem*>(items), n_items, block_item);
    D2T_CHECK_LAUNCH("repack_many");
    return 1;
}

extern "C" int d2t_wgrad_pack_input(const float* x, int N, int H, int W, int c_stride, int C, int stride, int OH, int OW,
                                    int pitch, float* xt, cudaStream_t stream) {
    D2T_REQUIRE(x && xt && N > 0 && H > 0 && W > 0 && C > 0 && c_stride >= C && stride > 0 && OH > 0 && OW > 0 &&
                    pitch >= OW && pitch % 4 == 0 && (OH - 1) * stride < H && (OW - 1) * stride < W,
                "d2t_wgrad_pack_input: bad arguments (row pitch a multiple of 4)");
    dim3 grid((pitch + 63) / 64, (C + 31) / 32, N * OH);
    D2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "d2t_wgrad_pack_input: tensor too large for the launch grid");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    D2T_CUDA_OK(cudaLaunchKernelEx(&cfg, nhwc_to_planes, x, N, H, W, c_stride, C, stride, OH, OW, pitch, xt), "nhwc_to_planes launch");
    return 1;
}

extern "C" int d2t_wgrad_pack_grad(const float* g, int N, int OH, int OW, int c_stride, int C, int pitch, int S, int dil,
                                   int pad, const float* amax_g, void* g_hi, void* g_lo, cudaStream_t stream) {
    D2T_REQUIRE(g && g_hi && g_lo && amax_g && N > 0 && OH > 0 && OW > 0 && C > 0 && c_stride >= C && pitch >= OW &&
                    pitch % 8 == 0 && S >= 1 && dil >= 1 && pad >= 0 && pad <= 8 && (S - 1) * dil - pad <= 8,
                "d2t_wgrad_pack_grad: bad arguments (row pitch a multiple of 8, column shifts within +-8)");
    dim3 grid((pitch + 63) / 64, (C + 31) / 32, N * OH);
    D2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "d2t_wgrad_pack_grad: tensor too large for the launch grid");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    D2T_CUDA_OK(cudaLaunchKernelEx(&cfg, nhwc_to_planes_split, g, N, OH, OW, 1, OH, OW, c_stride, C, pitch, S, dil, pad, amax_g,
                                   reinterpret_cast<__half*>(g_hi), reinterpret_cast<__half*>(g_lo)),
                "nhwc_to_planes_split launch");
    return 1;
}

// ---- correlation backward (conv.cu, CORRB): operand packers
// the other frame's features, sampled on the correlation lattice (stride 2 for conv3), as fp16 (hi, lo) planes
extern "C" int d2t_corrb_pack_other(const float* x, int N, int H, int W, int c_stride, int C, int stride, int OH, int OW,
                                    int pitch, const float* amax_x, void* o_hi, void* o_lo, cudaStream_t stream) {
    D2T_REQUIRE(x && o_hi && o_lo && amax_x && N > 0 && H > 0 && W > 0 && C > 0 && c_stride >= C && stride > 0 && OH > 0 &&
                    OW > 0 && pitch >= OW && pitch % 8 == 0 && (OH - 1) * stride < H && (OW - 1) * stride < W,
                "d2t_corrb_pack_other: bad arguments (row pitch a multiple of 8)");
    dim3 grid((pitch + 63) / 64, (C + 31) / 32, N * OH);
    D2T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "d2t_corrb_pack_other: tensor too large for the launch grid");
    nhwc_to_planes_split<<<grid, 256, 0, stream>>>(x, N, H, W, stride, OH, OW, c_stride, C, pitch, 1, 1, 0, amax_x,
                                                   reinterpret_cast<__half*>(o_hi), reinterpret_cast<__half*>(o_lo));
    D2T_CHECK_LAUNCH("nhwc_to_planes_split (corrb)");
    return 1;
}

extern "C" int d2t_corrb_pack_band(const float* g, int N, int H, int W, int c_stride, int c_offset, int r, int flipped,
                                   const float* amax_g, void* e_hi, void* e_lo, cudaStream_t stream) {
    D2T_REQUIRE(g && e_hi && e_lo && amax_g && N > 0 && H > 0 && W > 0 && r >= 1 && r <= 8 && c_offset >= 0 &&
                    c_stride >= c_offset + (2 * r + 1) * (2 * r + 1),
                "d2t_corrb_pack_band: bad arguments");
    const size_t rows = (size_t)N * H * W * (2 * r + 1);
    D2T_REQUIRE(rows < ((size_t)1 << 31), "d2t_corrb_pack_band: tensor too large");
    corr_band_pack<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(g, N, H, W, c_stride, c_offset, r, flipped, amax_g,
                                                                 reinterpret_cast<__half2*>(e_hi), reinterpret_cast<__half2*>(e_lo));
    D2T_CHECK_LAUNCH("corr_band_pack");
    return 1;
}

extern "C" int d2t_upsample2_add_mask(const float* low, int LH, int LW, const float* extra, const float* mask, int N, int H,
                                      int W, int C, float* out, float* amax_out, cudaStream_t stream) {
    D2T_REQUIRE(low && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && LH == (H + 1) / 2 && LW == (W + 1) / 2,
                "d2t_upsample2_add_mask: bad arguments (C % 4 == 0, low = ceil(full / 2))");
    const size_t total = (size_t)N * H * W * (C / 4);
    upsample2_add_mask<<<grid_for(total), 256, 0, stream>>>(reinterpret_cast<const float4*>(low), LH, LW,
                                                            reinterpret_cast<const float4*>(extra),
                                                            reinterpret_cast<const float4*>(mask), N, H, W, C / 4,
                                                            reinterpret_cast<float4*>(out), amax_out);
    D2T_CHECK_LAUNCH("upsample2_add_mask");
    return 1;
}

extern "C" int d2t_stem_pack_input(const float* x, int N, int C, int H, int W, float* packed, cudaStream_t stream) {
    D2T_REQUIRE(x && packed && N > 0 && C > 0 && C <= 4 && H > 0 && W > 0, "d2t_stem_pack_input: bad arguments");
    const int Hp = (H + 7) & ~1, Wp = W + 8;
    const size_t total = (size_t)N * Hp * Wp;
    stem_pack_input<<<grid_for(total), 256, 0, stream>>>(x, N, C, H, W, Hp, Wp, reinterpret_cast<float4*>(packed));
    D2T_CHECK_LAUNCH("stem_pack_input");
    return 1;
}

extern "C" int d2t_stem_pack_input_amax(const float* x, int N, int C, int H, int W, float* packed, float* amax, int pairs,
                                        cudaStream_t stream) {
    D2T_REQUIRE(x && packed && amax && N > 0 && C > 0 && C <= 4 && H > 0 && W > 0 && (pairs == 0 || (pairs > 0 && N == 2 * pairs)),
                "d2t_stem_pack_input_amax: bad arguments (pairs: 0, or N / 2 for a [pairs][2] frame batch)");
    const int Hp = (H + 7) & ~1, Wp = W + 8;
    const size_t total = (size_t)N * Hp * Wp;
    stem_pack_input_amax<<<grid_for(total), 256, 0, stream>>>(x, N, C, H, W, Hp, Wp, reinterpret_cast<float4*>(packed), amax, pairs);
    D2T_CHECK_LAUNCH("stem_pack_input_amax");
    return 1;
}

extern "C" int d2t_stem_pack_weights(const float* w, int Cout, int Cin, float* w_hi, float* w_lo, cudaStream_t stream) {
    D2T_REQUIRE(w && w_hi && Cout > 0 && Cin > 0 && Cin <= 4, "d2t_stem_pack_weights: bad arguments");
    stem_pack_weights<<<grid_for((size_t)Cout * 7 * 32), 256, 0, stream>>>(w, Cout, Cin, w_hi, w_lo);
    D2T_CHECK_LAUNCH("stem_pack_weights");
    return 1;
}

extern "C" int d2t_maxpool3x3s2_nhwc(const float* in, int N, int H, int W, int C, float* out, cudaStream_t stream) {
    D2T_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "d2t_maxpool3x3s2_nhwc: bad arguments");
    // ceil((H - 3) / 2) + 1, and the last window must start inside the input (torch pooling rule)
    int OH = (H - 3 + 1) / 2 + 1, OW = (W - 3 + 1) / 2 + 1;
    if ((OH - 1) * 2 >= H) --OH;
    if ((OW - 1) * 2 >= W) --OW;
    D2T_REQUIRE(OH > 0 && OW > 0, "d2t_maxpool3x3s2_nhwc: input too small");
    const size_t total = (size_t)N * OH * OW * (C / 4);
    maxpool3x3s2_nhwc<<<grid_for(total), 256, 0, stream>>>(in, N, H, W, C, OH, OW, out);
    D2T_CHECK_LAUNCH("maxpool3x3s2_nhwc");
    return 1;
}
