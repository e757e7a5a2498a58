// correlation.cu -- FlowNet-style cross-frame correlation, forward + backward, for sm_100a.
//
// Replaces channels_first / Correlation_forward / Correlation_backward_input1/2 and their
// launchers (/root/reference/lib/model/correlation/src/correlation_cuda_kernel.cu:10-32, 34-106,
// 108-198, 200-290, 296-369, 371-477) and the shape logic of correlation_cuda.c:20-38.
// Semantics: SURVEY.md App. A.1.
//
//   out[n, (tj+r)*D + (ti+r), y, x] = (1/(k*k*C)) * sum_{j,i in [-kr,kr]} sum_c
//         in1p[n, c, y1+j, x1+i] * in2p[n, c, y1+tj*s2+j, x1+ti*s2+i]
//   y1 = y*s1 + md + kr (padded coordinates), inXp = zero-padded input, r = md/s2, D = 2r+1.
//
// Design (B200).  The reference first transposes both inputs into zero-padded NHWC scratch
// (two extra full passes + three memsets), then runs one 32-thread block per output pixel that
// re-reads both channel vectors from global memory for each of the D*D displacements.  Here:
//   * no scratch, no transposes: NCHW rows are read directly (coalesced along x) and padding is
//     synthesised by zero-filling the shared-memory halo;
//   * the D&T configurations (k = 1, stride1 == stride2; rfcn.py:58-60) run a register-tiled
//     kernel: a CTA owns one output row, 64 x-positions and all D*D displacements; thread
//     (strip, tj) keeps an 8-position x D accumulator tile in registers and per channel reads
//     8 in1 values + an (8+2r)-wide in2 window from shared memory with 128-bit loads, i.e. 8*D
//     FMAs per (2 + (8+2r)/4) LDS.128 -- FMA-pipe bound rather than load bound;
//   * small batches are spread over all 148 SMs by splitting the channel range across a
//     thread-block CLUSTER (1/2/4/8 CTAs); partial tiles are summed through distributed shared
//     memory in a fixed rank order (deterministic, no atomics, no workspace);
//   * every other (k, pad, md, s1, s2) goes through a generic coalesced kernel.
// Backward is the exact adjoint of the forward for any parameters (see DESIGN.md for the
// places where the reference backward kernels are not), written as a gather so that every
// gradient element is produced by exactly one thread: deterministic and no pre-zeroing.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace d2t {
namespace {

struct CorrShape {
    int kr, r, D, oc, oh, ow;
};

__host__ __device__ inline CorrShape corr_shape(int H, int W, int pad, int k, int md, int s1, int s2) {
    CorrShape s;
    s.kr = (k - 1) / 2;
    int br = s.kr + md;
    s.r = md / s2;
    s.D = 2 * s.r + 1;
    s.oc = s.D * s.D;
    int nh = H + 2 * pad - 2 * br, nw = W + 2 * pad - 2 * br;
    s.oh = nh > 0 ? (nh + s1 - 1) / s1 : 0;   // ceil((pH - 2*br) / s1), correlation_cuda.c:33-34
    s.ow = nw > 0 ? (nw + s1 - 1) / s1 : 0;
    return s;
}

// ------------------------------------------------------------------ tuned forward, k == 1, s1 == s2
constexpr int kTX = 64;   // x positions per CTA (8 strips of 8)
constexpr int kCC = 16;   // channels staged per step

template <int R>
struct FwdCfg {
    static constexpr int D = 2 * R + 1;
    static constexpr int PITCH = kTX + 2 * R;            // 80 (R=8) / 72 (R=4): multiple of 4
    static constexpr int THREADS = 8 * D;
    static constexpr int STAGE_FLOATS = kCC * (kTX + D * PITCH);
    static constexpr int RED_FLOATS = D * D * kTX;
    static constexpr int SMEM_FLOATS = STAGE_FLOATS > RED_FLOATS ? STAGE_FLOATS : RED_FLOATS;
};

template <int R>
__global__ void __launch_bounds__(FwdCfg<R>::THREADS, 2)
corr_fwd_k1(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out, int C, int H,
            int W, int oh, int ow, int s, int o, int S, int c_per) {
    using Cfg = FwdCfg<R>;
    constexpr int D = Cfg::D, PITCH = Cfg::PITCH, NT = Cfg::THREADS;
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    float* in1s = sm;                       // [kCC][kTX]
    float* in2s = sm + kCC * kTX;           // [kCC][D][PITCH]

    const int rank = blockIdx.x % S, xtile = blockIdx.x / S;
    const int y = blockIdx.y, n = blockIdx.z;
    const int tid = threadIdx.x, strip = tid & 7, tjx = tid >> 3;
    const int xbase = xtile * kTX;
    const size_t HW = (size_t)H * W;
    const float* a_img = in1 + (size_t)n * C * HW;
    const float* b_img = in2 + (size_t)n * C * HW;
    const int c_begin = rank * c_per, c_end = min(C, c_begin + c_per);

    float acc[8][D];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int t = 0; t < D; ++t) acc[p][t] = 0.f;

    const int gy1 = y * s + o;
    for (int c0 = c_begin; c0 < c_end; c0 += kCC) {
        __syncthreads();
        // ---- stage in1 row segment: in1[c, y*s+o, (xbase+xx)*s+o]
        for (int idx = tid; idx < kCC * kTX; idx += NT) {
            const int cc = idx / kTX, xx = idx % kTX;
            const int c = c0 + cc, gx = (xbase + xx) * s + o;
            float v = 0.f;
            if (c < c_end && xbase + xx < ow && gy1 >= 0 && gy1 < H && gx >= 0 && gx < W)
                v = __ldg(a_img + (size_t)c * HW + (size_t)gy1 * W + gx);
            in1s[idx] = v;
        }
        // ---- stage in2 window rows: in2[c, (y+row-R)*s+o, (xbase+col-R)*s+o]
        for (int idx = tid; idx < kCC * D * PITCH; idx += NT) {
            const int cc = idx / (D * PITCH), rem = idx % (D * PITCH);
            const int row = rem / PITCH, col = rem % PITCH;
            const int c = c0 + cc;
            const int gy = (y + row - R) * s + o, gx = (xbase + col - R) * s + o;
            float v = 0.f;
            if (c < c_end && gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = __ldg(b_img + (size_t)c * HW + (size_t)gy * W + gx);
            in2s[idx] = v;
        }
        __syncthreads();
        // ---- 8 x D register tile
#pragma unroll 2
        for (int cc = 0; cc < kCC; ++cc) {
            float a[8], b[8 + 2 * R];
            const float4* ap = reinterpret_cast<const float4*>(in1s + cc * kTX + strip * 8);
            const float4* bp = reinterpret_cast<const float4*>(in2s + (cc * D + tjx) * PITCH + strip * 8);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float4 v = ap[q];
                a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < (8 + 2 * R) / 4; ++q) {
                float4 v = bp[q];
                b[4 * q] = v.x; b[4 * q + 1] = v.y; b[4 * q + 2] = v.z; b[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int t = 0; t < D; ++t) acc[p][t] = fmaf(a[p], b[p + t], acc[p][t]);
        }
    }

    const float nelems = (float)C;   // k == 1 (correlation_cuda_kernel.cu:63)
    float* o_img = out + (size_t)n * D * D * oh * ow;
    if (S == 1) {
#pragma unroll
        for (int t = 0; t < D; ++t) {
            float* dst = o_img + ((size_t)(tjx * D + t) * oh + y) * ow + xbase + strip * 8;
#pragma unroll
            for (int p = 0; p < 8; ++p)
                if (xbase + strip * 8 + p < ow) dst[p] = __fdiv_rn(acc[p][t], nelems);
        }
        return;
    }
    // ---- cluster reduction through distributed shared memory
    cg::cluster_group cluster = cg::this_cluster();
    __syncthreads();   // staging buffers are dead; reuse them as the partial tile
    float* red = sm;   // [D*D][kTX]
#pragma unroll
    for (int t = 0; t < D; ++t) {
        float4* dst = reinterpret_cast<float4*>(red + (tjx * D + t) * kTX + strip * 8);
        dst[0] = make_float4(acc[0][t], acc[1][t], acc[2][t], acc[3][t]);
        dst[1] = make_float4(acc[4][t], acc[5][t], acc[6][t], acc[7][t]);
    }
    cluster.sync();
    const int xs = min(kTX, ow - xbase);
    for (int tc = rank; tc < D * D; tc += S) {
        for (int xx = tid; xx < xs; xx += NT) {
            float sum = 0.f;
            for (int q = 0; q < S; ++q) sum += cluster.map_shared_rank(red, q)[tc * kTX + xx];
            o_img[((size_t)tc * oh + y) * ow + xbase + xx] = __fdiv_rn(sum, nelems);
        }
    }
    cluster.sync();   // peers may still be reading our tile
}

// ------------------------------------------------------------------ generic forward (any parameters)
__global__ void corr_fwd_generic(const float* __restrict__ in1, const float* __restrict__ in2,
                                 float* __restrict__ out, int B, int C, int H, int W, int pad, int k, int md,
                                 int s1, int s2, CorrShape sh) {
    const size_t total = (size_t)B * sh.oc * sh.oh * sh.ow;
    const size_t HW = (size_t)H * W;
    const float nelems = (float)(k * k * C);
    for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < total;
         index += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(index % sh.ow);
        const int y = (int)((index / sh.ow) % sh.oh);
        const int tc = (int)((index / sh.ow / sh.oh) % sh.oc);
        const int n = (int)(index / sh.ow / sh.oh / sh.oc);
        const int ti = tc % sh.D - sh.r, tj = tc / sh.D - sh.r;
        const int y1 = y * s1 + md + sh.kr - pad, x1 = x * s1 + md + sh.kr - pad;   // unpadded
        const int y2 = y1 + tj * s2, x2 = x1 + ti * s2;
        const float* a = in1 + (size_t)n * C * HW;
        const float* b = in2 + (size_t)n * C * HW;
        float acc = 0.f;
        for (int j = -sh.kr; j <= sh.kr; ++j)
            for (int i = -sh.kr; i <= sh.kr; ++i) {
                const int ya = y1 + j, xa = x1 + i, yb = y2 + j, xb = x2 + i;
                if (ya < 0 || ya >= H || xa < 0 || xa >= W || yb < 0 || yb >= H || xb < 0 || xb >= W) continue;
                const float* pa = a + (size_t)ya * W + xa;
                const float* pb = b + (size_t)yb * W + xb;
                for (int c = 0; c < C; ++c) acc = fmaf(__ldg(pa + c * HW), __ldg(pb + c * HW), acc);
            }
        out[index] = __fdiv_rn(acc, nelems);
    }
}

// ------------------------------------------------------------------ backward (exact adjoint, gather form)
// grad1[n,c,Y,X] = (1/nelems) * sum_{tj,ti} in2p[c, Yp + tj*s2, Xp + ti*s2] * G(tc; Y, X)
//   G = sum of gradOutput[n,tc,y,x] over the output positions whose k*k window covers (Y,X):
//       y*s1 + md + kr + j == Yp for some j in [-kr, kr]   (Yp = Y + pad)
// grad2[n,c,Y,X] = (1/nelems) * sum_{tj,ti} in1p[c, Yp - tj*s2, Xp - ti*s2] * G2(tc; Y, X)
//   G2 = sum over y with y*s1 + md + kr + tj*s2 + j == Yp.
// Thread = one (n, c, Y, X); x fastest so gradOutput / input reads coalesce; blockDim.y spans
// channels so the gradOutput lines are shared through L1 by the channel-threads of a CTA.
template <int WHICH>
__global__ void __launch_bounds__(256)
corr_bwd_gather(const float* __restrict__ other, const float* __restrict__ gout, float* __restrict__ grad, int C,
                int H, int W, int pad, int k, int md, int s1, int s2, CorrShape sh) {
    const int X = blockIdx.x * 32 + threadIdx.x;
    const int Y = blockIdx.y;
    const int nc = blockIdx.z * blockDim.y + threadIdx.y;   // n*C + c
    if (X >= W) return;
    const int n = nc / C;
    const size_t HW = (size_t)H * W;
    const float* oth = other + (size_t)nc * HW;
    const float* go = gout + (size_t)n * sh.oc * sh.oh * sh.ow;
    const int Yp = Y + pad, Xp = X + pad;
    const float nelems = (float)(k * k * C);
    float acc = 0.f;
    for (int tjx = 0; tjx < sh.D; ++tjx) {
        const int dj = (tjx - sh.r) * s2;
        const int yo = (WHICH == 1 ? Yp + dj : Yp - dj) - pad;   // partner row (unpadded)
        if (yo < 0 || yo >= H) continue;
        // output rows y with y*s1 + md + kr + (WHICH==2 ? dj : 0) + j == Yp
        const int ybase = Yp - md - sh.kr - (WHICH == 2 ? dj : 0);
        for (int tix = 0; tix < sh.D; ++tix) {
            const int di = (tix - sh.r) * s2;
            const int xo = (WHICH == 1 ? Xp + di : Xp - di) - pad;
            if (xo < 0 || xo >= W) continue;
            const int xbase = Xp - md - sh.kr - (WHICH == 2 ? di : 0);
            const float* g = go + (size_t)(tjx * sh.D + tix) * sh.oh * sh.ow;
            float gs = 0.f;
            for (int j = -sh.kr; j <= sh.kr; ++j) {
                const int yn = ybase - j;
                if (yn < 0 || yn % s1 != 0) continue;
                const int y = yn / s1;
                if (y >= sh.oh) continue;
                for (int i = -sh.kr; i <= sh.kr; ++i) {
                    const int xn = xbase - i;
                    if (xn < 0 || xn % s1 != 0) continue;
                    const int x = xn / s1;
                    if (x >= sh.ow) continue;
                    gs += __ldg(g + (size_t)y * sh.ow + x);
                }
            }
            acc = fmaf(gs, __ldg(oth + (size_t)yo * W + xo), acc);
        }
    }
    grad[(size_t)nc * HW + (size_t)Y * W + X] = __fdiv_rn(acc, nelems);
}

template <int R>
int launch_fwd_k1(const float* in1, const float* in2, float* out, int B, int C, int H, int W, int s, int o,
                  const CorrShape& sh, cudaStream_t stream) {
    using Cfg = FwdCfg<R>;
    const int xtiles = (sh.ow + kTX - 1) / kTX;
    const int base_blocks = xtiles * sh.oh * B;
    int S = 1;
    const int sms = sm_count();
    while (S < 8 && base_blocks * S * 2 <= sms * 2 + sms / 2 && C / (S * 2) >= 2 * kCC) S *= 2;
    int c_per = (C + S - 1) / S;
    c_per = (c_per + kCC - 1) / kCC * kCC;
    const size_t smem = (size_t)Cfg::SMEM_FLOATS * sizeof(float);
    static SmemAttrOnce once;
    if (!once.ensure(corr_fwd_k1<R>, smem, "corr_fwd smem attr")) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(xtiles * S, sh.oh, B);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    D2T_CUDA_OK(cudaLaunchKernelEx(&cfg, corr_fwd_k1<R>, in1, in2, out, C, H, W, sh.oh, sh.ow, s, o, S, c_per),
                "corr_fwd_k1 launch");
    return 1;
}

}  // namespace
}  // namespace d2t

using namespace d2t;

extern "C" int d2t_correlation_shape(int H, int W, int pad, int k, int md, int s1, int s2, int* out3) {
    D2T_REQUIRE(k >= 1 && (k & 1) && s1 >= 1 && s2 >= 1 && md >= 0 && pad >= 0, "d2t_correlation_shape: bad params");
    CorrShape sh = corr_shape(H, W, pad, k, md, s1, s2);
    out3[0] = sh.oc;
    out3[1] = sh.oh;
    out3[2] = sh.ow;
    return 1;
}

extern "C" int d2t_correlation_forward(const float* in1, const float* in2, int B, int C, int H, int W, int pad,
                                       int k, int md, int s1, int s2, float* out, cudaStream_t stream) {
    D2T_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "d2t_correlation_forward: bad sizes");
    D2T_REQUIRE(k >= 1 && (k & 1) && s1 >= 1 && s2 >= 1 && md >= 0 && pad >= 0,
                "d2t_correlation_forward: bad params (kernel_size must be odd)");
    D2T_REQUIRE(in1 && in2 && out, "d2t_correlation_forward: null pointer");
    CorrShape sh = corr_shape(H, W, pad, k, md, s1, s2);
    D2T_REQUIRE(sh.oh > 0 && sh.ow > 0, "d2t_correlation_forward: empty output");
    if (k == 1 && s1 == s2 && B <= 65535 && sh.oh <= 65535) {
        const int o = md - pad;
        if (sh.r == 8) return launch_fwd_k1<8>(in1, in2, out, B, C, H, W, s1, o, sh, stream);
        if (sh.r == 4) return launch_fwd_k1<4>(in1, in2, out, B, C, H, W, s1, o, sh, stream);
    }
    const size_t total = (size_t)B * sh.oc * sh.oh * sh.ow;
    size_t blocks = (total + 255) / 256, cap = (size_t)sm_count() * 32;
    corr_fwd_generic<<<(int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(in1, in2, out, B, C, H, W, pad, k, md,
                                                                            s1, s2, sh);
    D2T_CHECK_LAUNCH("corr_fwd_generic");
    return 1;
}

extern "C" int d2t_correlation_backward(const float* in1, const float* in2, const float* grad_out, int B, int C,
                                        int H, int W, int pad, int k, int md, int s1, int s2, float* grad1,
                                        float* grad2, cudaStream_t stream) {
    D2T_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "d2t_correlation_backward: bad sizes");
    D2T_REQUIRE(k >= 1 && (k & 1) && s1 >= 1 && s2 >= 1 && md >= 0 && pad >= 0,
                "d2t_correlation_backward: bad params");
    D2T_REQUIRE(in1 && in2 && grad_out, "d2t_correlation_backward: null pointer");
    CorrShape sh = corr_shape(H, W, pad, k, md, s1, s2);
    D2T_REQUIRE(sh.oh > 0 && sh.ow > 0, "d2t_correlation_backward: empty output");
    const long long nc = (long long)B * C;
    int cy = 8;
    while (cy > 1 && nc % cy != 0) cy >>= 1;
    D2T_REQUIRE(nc / cy <= 65535 && H <= 65535, "d2t_correlation_backward: tensor too large for the launch grid");
    dim3 block(32, cy), grid((W + 31) / 32, H, (unsigned)(nc / cy));
    if (grad1) {
        corr_bwd_gather<1><<<grid, block, 0, stream>>>(in2, grad_out, grad1, C, H, W, pad, k, md, s1, s2, sh);
        D2T_CHECK_LAUNCH("corr_bwd_gather<1>");
    }
    if (grad2) {
        corr_bwd_gather<2><<<grid, block, 0, stream>>>(in1, grad_out, grad2, C, H, W, pad, k, md, s1, s2, sh);
        D2T_CHECK_LAUNCH("corr_bwd_gather<2>");
    }
    return 1;
}

// ---- reference-named launchers (correlation_cuda_kernel.h:5-88) ----
extern "C" int Correlation_forward_cuda_kernel(float* output, int ob, int oc, int oh, int ow, int osb, int osc,
                                               int osh, int osw, float* input1, int ic, int ih, int iw, int isb,
                                               int isc, int ish, int isw, float* input2, int gc, int gsb, int gsc,
                                               int gsh, int gsw, float* rInput1, float* rInput2, int pad_size,
                                               int kernel_size, int max_displacement, int stride1, int stride2,
                                               int corr_type_multiply, cudaStream_t stream) {
    (void)rInput1; (void)rInput2; (void)corr_type_multiply; (void)gc;
    D2T_REQUIRE(isw == 1 && ish == iw && isc == ih * iw && isb == ic * ih * iw && gsw == 1 && gsh == iw &&
                    gsc == ih * iw && gsb == ic * ih * iw,
                "Correlation_forward_cuda_kernel: inputs must be contiguous NCHW");
    CorrShape sh = corr_shape(ih, iw, pad_size, kernel_size, max_displacement, stride1, stride2);
    D2T_REQUIRE(sh.oc == oc && sh.oh == oh && sh.ow == ow && osw == 1 && osh == ow && osc == oh * ow &&
                    osb == oc * oh * ow,
                "Correlation_forward_cuda_kernel: output must be contiguous [B,%d,%d,%d]", sh.oc, sh.oh, sh.ow);
    return d2t_correlation_forward(input1, input2, ob, ic, ih, iw, pad_size, kernel_size, max_displacement, stride1,
                                   stride2, output, stream);
}

extern "C" int Correlation_backward_cuda_kernel(
    float* gradOutput, int gob, int goc, int goh, int gow, int gosb, int gosc, int gosh, int gosw, float* input1,
    int ic, int ih, int iw, int isb, int isc, int ish, int isw, float* input2, int gsb, int gsc, int gsh, int gsw,
    float* gradInput1, int gisb, int gisc, int gish, int gisw, float* gradInput2, int ggc, int ggsb, int ggsc,
    int ggsh, int ggsw, float* rInput1, float* rInput2, int pad_size, int kernel_size, int max_displacement,
    int stride1, int stride2, int corr_type_multiply, cudaStream_t stream) {
    (void)rInput1; (void)rInput2; (void)corr_type_multiply; (void)ggc;
    const int chw = ic * ih * iw, hw = ih * iw;
    D2T_REQUIRE(isw == 1 && ish == iw && isc == hw && isb == chw && gsw == 1 && gsh == iw && gsc == hw && gsb == chw &&
                    gisw == 1 && gish == iw && gisc == hw && gisb == chw && ggsw == 1 && ggsh == iw && ggsc == hw &&
                    ggsb == chw,
                "Correlation_backward_cuda_kernel: inputs/gradInputs must be contiguous NCHW");
    CorrShape sh = corr_shape(ih, iw, pad_size, kernel_size, max_displacement, stride1, stride2);
    D2T_REQUIRE(sh.oc == goc && sh.oh == goh && sh.ow == gow && gosw == 1 && gosh == gow && gosc == goh * gow &&
                    gosb == goc * goh * gow,
                "Correlation_backward_cuda_kernel: gradOutput must be contiguous [B,%d,%d,%d]", sh.oc, sh.oh, sh.ow);
    return d2t_correlation_backward(input1, input2, gradOutput, gob, ic, ih, iw, pad_size, kernel_size,
                                    max_displacement, stride1, stride2, gradInput1, gradInput2, stream);
}
