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
