/*
 * oracle_cpu.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's native operators on the Detect-to-Track
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libd2t_b200.so) never does.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/lib/model/).  Where the reference's *compiled* arithmetic differs from
 * the source text (FMA contraction chosen by nvcc for sm_100a -- checked with
 * `cuobjdump -sass oracle/_ref/obj/*.o`), the restatement uses explicit fmaf() and this
 * file is built with -ffp-contract=off so nothing else is fused.  `contract` arguments
 * select 1 = as compiled (default everywhere) / 0 = as written, so tests can count how
 * often the two differ.
 *
 * Parity status: PINNED.  tests/golden/ holds outputs of the reference's own kernels
 * (oracle/_ref/libref_oracle.so, built from the unmodified reference .cu files) run on a
 * B200 by tests/golden/make_golden_gpu.py, and of the reference's own Python RPN code
 * run on CPU by tests/golden/make_golden_rpn.py; tests/test_oracle_golden.py checks this
 * restatement against them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include <pthread.h>

#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------
 * Host threading for the CPU-baseline timings (no OpenMP runtime in this image): the outer
 * loop of the forward restatements is split into contiguous index ranges, one per thread.
 * Every output element is still produced by exactly one thread with the same arithmetic, so
 * results do not depend on the thread count.  oracle_set_threads(1) (the default) = serial.
 * ------------------------------------------------------------------------------------ */
static int g_threads = 1;
void oracle_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int oracle_get_threads(void) { return g_threads; }

typedef void (*range_fn)(int begin, int end, void* ctx);
typedef struct { range_fn fn; void* ctx; int begin, end; } range_job;
static void* range_tramp(void* p) { range_job* j = (range_job*)p; j->fn(j->begin, j->end, j->ctx); return NULL; }
static void parallel_range(int n, range_fn fn, void* ctx)
{
    int T = IMIN(g_threads, n);
    if (T <= 1) { fn(0, n, ctx); return; }
    pthread_t th[256]; range_job jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].begin = (int)((long long)n * t / T); jobs[t].end = (int)((long long)n * (t + 1) / T);
        pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------
 * PSRoI pooling   psroi_pooling/src/psroi_pooling_kernel.cu:15-79 (fwd), :109-170 (bwd)
 * ------------------------------------------------------------------------------------ */

/* [hstart,hend) x [wstart,wend) of bin (ph,pw) of one roi -- kernel.cu:31-62.
 * Compiled form (SASS of PSROIPoolForward, sm_100a):
 *   start = FMUL(roundf(x1), scale)
 *   width = FFMA(roundf(x2)+1, scale, -start)       <- contraction of end - start
 *   width = max(width, 0.1)  (done in double; identical to fmaxf(width, 0.1f))
 *   bin   = width / P  (IEEE)
 *   hstart = F2I.FLOOR(FFMA(ph, bin, start)) ; hend = F2I.CEIL(FFMA(ph+1, bin, start)) */
void oracle_psroi_bin(const float* roi, float scale, int PH, int PW, int H, int W,
                      int ph, int pw, int contract, int* out4)
{
    float rsw = roundf(roi[1]) * scale;
    float rsh = roundf(roi[2]) * scale;
    float tw = roundf(roi[3]) + 1.f;
    float th = roundf(roi[4]) + 1.f;
    float rw, rh;
    if (contract) {
        rw = fmaf(tw, scale, -rsw);
        rh = fmaf(th, scale, -rsh);
    } else {
        float rew = tw * scale, reh = th * scale;
        rw = rew - rsw;
        rh = reh - rsh;
    }
    rw = fmaxf(rw, 0.1f);
    rh = fmaxf(rh, 0.1f);
    float bh = rh / (float)PH;
    float bw = rw / (float)PW;
    float fhs, fws, fhe, fwe;
    if (contract) {
        fhs = fmaf((float)ph, bh, rsh);
        fws = fmaf((float)pw, bw, rsw);
        fhe = fmaf((float)(ph + 1), bh, rsh);
        fwe = fmaf((float)(pw + 1), bw, rsw);
    } else {
        float a = (float)ph * bh;        fhs = a + rsh;
        float b = (float)pw * bw;        fws = b + rsw;
        float c = (float)(ph + 1) * bh;  fhe = c + rsh;
        float d = (float)(pw + 1) * bw;  fwe = d + rsw;
    }
    int hs = (int)floorf(fhs), ws = (int)floorf(fws);
    int he = (int)ceilf(fhe), we = (int)ceilf(fwe);
    out4[0] = IMIN(IMAX(hs, 0), H);
    out4[1] = IMIN(IMAX(he, 0), H);
    out4[2] = IMIN(IMAX(ws, 0), W);
    out4[3] = IMIN(IMAX(we, 0), W);
}

/* bins: optional int32 [R, PH, PW, 4] dump of the integer windows (may be NULL).
 * mapping: optional int32 [R, D, PH, PW]. */
typedef struct { const float* feat; int B, C, H, W; const float* rois; float scale; int PH, PW, G, D, contract;
                 float* top; int32_t* mapping; int32_t* bins; } psroi_ctx;
static void psroi_fwd_range(int n0, int n1, void* vp)
{
    psroi_ctx* q = (psroi_ctx*)vp;
    const float* feat = q->feat; const float* rois = q->rois; float scale = q->scale;
    int C = q->C, H = q->H, W = q->W, PH = q->PH, PW = q->PW, G = q->G, D = q->D, contract = q->contract;
    float* top = q->top; int32_t* mapping = q->mapping; int32_t* bins = q->bins;
    for (int n = n0; n < n1; ++n) {
        const float* roi = rois + 5 * n;
        int b = (int)roi[0];
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw) {
                int w4[4];
                oracle_psroi_bin(roi, scale, PH, PW, H, W, ph, pw, contract, w4);
                if (bins) memcpy(bins + ((size_t)(n * PH + ph) * PW + pw) * 4, w4, sizeof w4);
                int hs = w4[0], he = w4[1], ws = w4[2], we = w4[3];
                int empty = (he <= hs) || (we <= ws);
                float area = (float)((he - hs) * (we - ws));
                for (int ct = 0; ct < D; ++ct) {
                    int c = (ct * G + ph) * G + pw;              /* kernel.cu:64-66 */
                    const float* p = feat + ((size_t)b * C + c) * H * W;
                    float s = 0.f;
                    for (int h = hs; h < he; ++h)                  /* kernel.cu:69-74 */
                        for (int w = ws; w < we; ++w) s += p[h * W + w];
                    size_t idx = (((size_t)n * D + ct) * PH + ph) * PW + pw;
                    top[idx] = empty ? 0.f : s / area;             /* kernel.cu:76 */
                    if (mapping) mapping[idx] = c;
                }
            }
    }
}
void oracle_psroi_forward(const float* feat, int B, int C, int H, int W,
                          const float* rois, int R, float scale, int PH, int PW,
                          int G, int D, int contract,
                          float* top, int32_t* mapping, int32_t* bins)
{
    psroi_ctx q = { feat, B, C, H, W, rois, scale, PH, PW, G, D, contract, top, mapping, bins };
    parallel_range(R, psroi_fwd_range, &q);
}

/* bottom_diff must be zero-filled by the caller (functions/psroi_pool.py:40).
 * The reference scatters with float atomicAdd in an undefined order (kernel.cu:166);
 * this restatement adds in output-index order, so parity is tolerance-level. */
void oracle_psroi_backward(const float* top_diff, int B, int C, int H, int W,
                           const float* rois, int R, float scale, int PH, int PW,
                           int G, int D, int contract, float* bottom_diff)
{
    (void)B;
    for (int n = 0; n < R; ++n) {
        const float* roi = rois + 5 * n;
        int b = (int)roi[0];
        for (int ct = 0; ct < D; ++ct)
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    int w4[4];
                    oracle_psroi_bin(roi, scale, PH, PW, H, W, ph, pw, contract, w4);
                    int hs = w4[0], he = w4[1], ws = w4[2], we = w4[3];
                    int empty = (he <= hs) || (we <= ws);
                    if (empty) continue;
                    int c = (ct * G + ph) * G + pw;
                    float area = (float)((he - hs) * (we - ws));
                    size_t idx = (((size_t)n * D + ct) * PH + ph) * PW + pw;
                    float dv = top_diff[idx] / area;               /* kernel.cu:161 */
                    float* p = bottom_diff + ((size_t)b * C + c) * H * W;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w) p[h * W + w] += dv;
                }
    }
}

/* ------------------------------------------------------------------------------------
 * NMS   nms/src/nms_cuda_kernel.cu:31-39 (devIoU), :68-84 (mask), :123-144 (host sweep)
 * ------------------------------------------------------------------------------------ */

/* a = the row box (cur_box, kernel.cu:70), b = the column box (block_boxes + i*5).
 * Compiled form (SASS of nms_kernel, sm_100a):
 *   Sa = FMUL(a2-a0+1, a3-a1+1); t = FFMA(b2-b0+1, b3-b1+1, Sa); I = FMUL(w, h);
 *   den = FADD(t, -I); iou = I / den (IEEE); suppressed iff iou > thresh. */
float oracle_iou(const float* a, const float* b, int contract)
{
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float w = fmaxf(right - left + 1.f, 0.f), h = fmaxf(bottom - top + 1.f, 0.f);
    float inter = w * h;
    float Sa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
    float wb = b[2] - b[0] + 1.f, hb = b[3] - b[1] + 1.f;
    float t;
    if (contract) t = fmaf(wb, hb, Sa);
    else { float Sb = wb * hb; t = Sa + Sb; }
    return inter / (t - inter);
}

/* Greedy sweep in index order over caller-sorted boxes; equivalent to building the
 * 64-wide suppression masks (:68-84) and sweeping them (:132-144).  keep_out gets the
 * kept indices ascending; returns their count.  max_keep <= 0 means no cap. */
int oracle_nms(const float* boxes, int N, int box_dim, float thresh, int contract,
               int max_keep, int32_t* keep_out)
{
    uint8_t* dead = (uint8_t*)calloc((size_t)(N > 0 ? N : 1), 1);
    int k = 0;
    for (int i = 0; i < N; ++i) {
        if (dead[i]) continue;
        keep_out[k++] = i;
        if (max_keep > 0 && k >= max_keep) break;
        const float* a = boxes + (size_t)i * box_dim;
        for (int j = i + 1; j < N; ++j)
            if (!dead[j] && oracle_iou(a, boxes + (size_t)j * box_dim, contract) > thresh)
                dead[j] = 1;
    }
    free(dead);
    return k;
}

/* ------------------------------------------------------------------------------------
 * Correlation   correlation/src/correlation_cuda.c:20-38 (shapes),
 *               correlation_cuda_kernel.cu:34-106 (fwd), :108-198, :200-290 (bwd)
 * ------------------------------------------------------------------------------------ */
void oracle_correlation_shape(int H, int W, int pad, int k, int md, int s1, int s2,
                              int* out3 /* oc, oh, ow */)
{
    int kr = (k - 1) / 2, br = kr + md;
    int pH = H + 2 * pad, pW = W + 2 * pad;
    int r = md / s2, Dd = 2 * r + 1;
    out3[0] = Dd * Dd;
    out3[1] = (int)ceilf((float)(pH - 2 * br) / (float)s1);
    out3[2] = (int)ceilf((float)(pW - 2 * br) / (float)s1);
}

static inline float padded_at(const float* in, int C, int H, int W, int n, int c,
                              int yp, int xp, int pad)
{   /* value of the zero-padded NHWC scratch rInput[n, yp, xp, c] (kernel.cu:10-32) */
    int y = yp - pad, x = xp - pad;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;
    return in[(((size_t)n * C + c) * H + y) * W + x];
}

/* Reduction order follows the reference: lane l of the 32-thread block accumulates
 * channels l, l+32, ... over the k*k window with FFMA (kernel.cu:78-90), lane 0 sums the
 * 32 partials in order and divides by nelems (:93-100). */
typedef struct { const float* in1; const float* in2; int B, C, H, W, pad, k, md, s1, s2; float* out; } corr_ctx;
static void corr_fwd_range(int row0, int row1, void* vp)
{
    corr_ctx* q = (corr_ctx*)vp;
    const float* in1 = q->in1; const float* in2 = q->in2; float* out = q->out;
    int C = q->C, H = q->H, W = q->W, pad = q->pad, k = q->k, md = q->md, s1 = q->s1, s2 = q->s2;
    int sh[3];
    oracle_correlation_shape(H, W, pad, k, md, s1, s2, sh);
    int oc = sh[0], oh = sh[1], ow = sh[2];
    int kr = (k - 1) / 2, r = md / s2, Dd = 2 * r + 1;
    float nelems = (float)(k * k * C);
    for (int row = row0; row < row1; ++row) {
        int n = row / oh, y = row % oh;
            for (int x = 0; x < ow; ++x) {
                int y1 = y * s1 + md + kr, x1 = x * s1 + md + kr;
                for (int tj = -r; tj <= r; ++tj)
                    for (int ti = -r; ti <= r; ++ti) {
                        int x2 = x1 + ti * s2, y2 = y1 + tj * s2;
                        float part[32];
                        for (int l = 0; l < 32; ++l) part[l] = 0.f;
                        for (int j = -kr; j <= kr; ++j)
                            for (int i = -kr; i <= kr; ++i)
                                for (int ch = 0; ch < C; ++ch) {
                                    float a = padded_at(in1, C, H, W, n, ch, y1 + j, x1 + i, pad);
                                    float b = padded_at(in2, C, H, W, n, ch, y2 + j, x2 + i, pad);
                                    part[ch & 31] = fmaf(a, b, part[ch & 31]);
                                }
                        float s = 0.f;
                        for (int l = 0; l < 32; ++l) s += part[l];
                        int tc = (tj + r) * Dd + (ti + r);
                        out[(((size_t)n * oc + tc) * oh + y) * ow + x] = s / nelems;
                    }
            }
    }
}
void oracle_correlation_forward(const float* in1, const float* in2, int B, int C, int H,
                                int W, int pad, int k, int md, int s1, int s2, float* out)
{
    int sh[3];
    oracle_correlation_shape(H, W, pad, k, md, s1, s2, sh);
    corr_ctx q = { in1, in2, B, C, H, W, pad, k, md, s1, s2, out };
    parallel_range(B * sh[1], corr_fwd_range, &q);
}

/* Faithful restatement of Correlation_backward_input1/_input2 including their launch
 * geometry (grid (H, W, C), position = blockIdx*stride1 + pad, kernel.cu:120-121, 212-213,
 * 437-463).  For stride1 > 1 the reference computes flat write offsets beyond the
 * (H, W) plane; the caller passes the number of floats available in each gradient buffer
 * (`cap`), writes landing at >= cap are dropped and counted in oob[0] (grad1), oob[1]
 * (grad2) -- in the reference they corrupt whatever follows the tensor.  Buffers must be
 * zero-filled by the caller (correlation_cuda.c:111-114). */
void oracle_correlation_backward_ref(const float* in1, const float* in2, const float* gout,
                                     int B, int C, int H, int W, int pad, int k, int md,
                                     int s1, int s2, float* g1, float* g2, size_t cap,
                                     int64_t* oob)
{
    int sh[3];
    oracle_correlation_shape(H, W, pad, k, md, s1, s2, sh);
    int oc = sh[0], oh = sh[1], ow = sh[2];
    int kr = (k - 1) / 2, r = md / s2, Dd = 2 * r + 1;
    float nelems = (float)(k * k * C);
    int pH = H + 2 * pad, pW = W + 2 * pad;
    oob[0] = oob[1] = 0;
    for (int n = 0; n < B; ++n)
        for (int by = 0; by < H; ++by)
            for (int bx = 0; bx < W; ++bx) {
                int y = by * s1 + pad, x = bx * s1 + pad;
                /* ---- input1 (kernel.cu:128-197) ---- */
                int xmin = (x - kr - md) / s1, ymin = (y - kr - md) / s1;
                int xmax = (x + kr - md) / s1, ymax = (y + kr - md) / s1;
                int skip1 = (xmax < 0 || ymax < 0 || xmin >= ow || ymin >= oh) ||
                            (xmin > xmax || ymin > ymax);
                int xmin1 = IMAX(0, xmin), xmax1 = IMIN(ow - 1, xmax);
                int ymin1 = IMAX(0, ymin), ymax1 = IMIN(oh - 1, ymax);
                for (int c = 0; c < C; ++c) {
                    if (!skip1) {
                        float part[32];
                        for (int l = 0; l < 32; ++l) part[l] = 0.f;
                        for (int tc = 0; tc < oc; ++tc) {
                            int i2 = (tc % Dd - r) * s2, j2 = (tc / Dd - r) * s2;
                            int yy = y + j2, xx = x + i2;
                            float v2 = (yy >= 0 && yy < pH && xx >= 0 && xx < pW)
                                           ? padded_at(in2, C, H, W, n, c, yy, xx, pad) : 0.f;
                            for (int j = ymin1; j <= ymax1; ++j)
                                for (int i = xmin1; i <= xmax1; ++i)
                                    part[tc & 31] = fmaf(gout[(((size_t)n * oc + tc) * oh + j) * ow + i], v2, part[tc & 31]);
                        }
                        float s = 0.f;
                        for (int l = 0; l < 32; ++l) s += part[l];
                        size_t idx = ((size_t)n * C + c) * H * W + (size_t)(y - pad) * W + (x - pad);
                        if (idx < cap) g1[idx] = s / nelems; else oob[0]++;
                    }
                    /* ---- input2 (kernel.cu:240-288) ---- */
                    {
                        float part[32];
                        for (int l = 0; l < 32; ++l) part[l] = 0.f;
                        for (int tc = 0; tc < oc; ++tc) {
                            int i2 = (tc % Dd - r) * s2, j2 = (tc / Dd - r) * s2;
                            int xmn = (x - kr - md - i2) / s1, ymn = (y - kr - md - j2) / s1;
                            int xmx = (x + kr - md - i2) / s1, ymx = (y + kr - md - j2) / s1;
                            if (xmx < 0 || ymx < 0 || xmn >= ow || ymn >= oh) continue;
                            if (xmn > xmx || ymn > ymx) continue;
                            xmn = IMAX(0, xmn); xmx = IMIN(ow - 1, xmx);
                            ymn = IMAX(0, ymn); ymx = IMIN(oh - 1, ymx);
                            int yy = y - j2, xx = x - i2;
                            float v1 = (yy >= 0 && yy < pH && xx >= 0 && xx < pW)
                                           ? padded_at(in1, C, H, W, n, c, yy, xx, pad) : 0.f;
                            for (int j = ymn; j <= ymx; ++j)
                                for (int i = xmn; i <= xmx; ++i)
                                    part[tc & 31] = fmaf(gout[(((size_t)n * oc + tc) * oh + j) * ow + i], v1, part[tc & 31]);
                        }
                        float s = 0.f;
                        for (int l = 0; l < 32; ++l) s += part[l];
                        size_t idx = ((size_t)n * C + c) * H * W + (size_t)(y - pad) * W + (x - pad);
                        if (idx < cap) g2[idx] = s / nelems; else oob[1]++;
                    }
                }
            }
}

/* Mathematically exact adjoint of oracle_correlation_forward (double accumulation): the
 * gradient the product kernels must reproduce for ANY (pad, k, md, s1, s2).  It agrees
 * with oracle_correlation_backward_ref wherever the reference is well defined. */
void oracle_correlation_backward_true(const float* in1, const float* in2, const float* gout,
                                      int B, int C, int H, int W, int pad, int k, int md,
                                      int s1, int s2, float* g1, float* g2)
{
    int sh[3];
    oracle_correlation_shape(H, W, pad, k, md, s1, s2, sh);
    int oc = sh[0], oh = sh[1], ow = sh[2];
    int kr = (k - 1) / 2, r = md / s2, Dd = 2 * r + 1;
    double nelems = (double)(k * k * C);
    size_t tot = (size_t)B * C * H * W;
    double* a1 = (double*)calloc(tot, sizeof(double));
    double* a2 = (double*)calloc(tot, sizeof(double));
    for (int n = 0; n < B; ++n)
        for (int tc = 0; tc < oc; ++tc) {
            int ti = tc % Dd - r, tj = tc / Dd - r;
            for (int y = 0; y < oh; ++y)
                for (int x = 0; x < ow; ++x) {
                    double go = gout[(((size_t)n * oc + tc) * oh + y) * ow + x] / nelems;
                    int y1 = y * s1 + md + kr - pad, x1 = x * s1 + md + kr - pad;
                    int y2 = y1 + tj * s2, x2 = x1 + ti * s2;
                    for (int j = -kr; j <= kr; ++j)
                        for (int i = -kr; i <= kr; ++i) {
                            int ya = y1 + j, xa = x1 + i, yb = y2 + j, xb = x2 + i;
                            if (ya < 0 || ya >= H || xa < 0 || xa >= W) continue;
                            if (yb < 0 || yb >= H || xb < 0 || xb >= W) continue;
                            for (int c = 0; c < C; ++c) {
                                size_t ia = (((size_t)n * C + c) * H + ya) * W + xa;
                                size_t ib = (((size_t)n * C + c) * H + yb) * W + xb;
                                a1[ia] += go * in2[ib];
                                a2[ib] += go * in1[ia];
                            }
                        }
                }
        }
    for (size_t i = 0; i < tot; ++i) { g1[i] = (float)a1[i]; g2[i] = (float)a2[i]; }
    free(a1); free(a2);
}

/* ------------------------------------------------------------------------------------
 * RoIAlign   roi_align/src/roi_align_kernel.cu:15-70 (fwd), :94-143 (bwd)
 * Compiled form: start = FMUL(x, scale); extent = fmaxf(FFMA(x2, scale, -start) + 1, 0);
 * bin = (float)((double)extent / (double)(A - 1)); h = FFMA(ph, bin, start);
 * interpolation evaluated in double and rounded once to float.
 * ------------------------------------------------------------------------------------ */
static inline int roi_align_sample(const float* roi, float scale, int AH, int AW, int H,
                                   int W, int ph, int pw, int* hs, int* ws, float* hr,
                                   float* wr)
{
    float sw = roi[1] * scale, sh = roi[2] * scale;
    float rw = fmaxf(fmaf(roi[3], scale, -sw) + 1.f, 0.f);
    float rh = fmaxf(fmaf(roi[4], scale, -sh) + 1.f, 0.f);
    float bh = (float)((double)rh / ((double)AH - 1.));
    float bw = (float)((double)rw / ((double)AW - 1.));
    float h = fmaf((float)ph, bh, sh), w = fmaf((float)pw, bw, sw);
    if (h < 0 || h >= H || w < 0 || w >= W) return 0;
    *hs = (int)fminf(floorf(h), (float)(H - 2));
    *ws = (int)fminf(floorf(w), (float)(W - 2));
    *hr = h - (float)(*hs);
    *wr = w - (float)(*ws);
    return 1;
}

void oracle_roi_align_forward(const float* feat, int B, int C, int H, int W,
                              const float* rois, int R, float scale, int AH, int AW,
                              float* top)
{
    (void)B;
    for (int n = 0; n < R; ++n) {
        const float* roi = rois + 5 * n;
        int b = (int)roi[0];
        for (int ph = 0; ph < AH; ++ph)
            for (int pw = 0; pw < AW; ++pw) {
                int hs = 0, ws = 0; float hr = 0, wr = 0;
                int in = roi_align_sample(roi, scale, AH, AW, H, W, ph, pw, &hs, &ws, &hr, &wr);
                for (int c = 0; c < C; ++c) {
                    size_t idx = (((size_t)n * C + c) * AH + ph) * AW + pw;
                    if (!in) { top[idx] = 0.f; continue; }
                    const float* p = feat + (((size_t)b * C + c) * H + hs) * W + ws;
                    double v = (double)p[0] * (1. - hr) * (1. - wr) + (double)p[1] * (1. - hr) * wr +
                               (double)p[W] * hr * (1. - wr) + (double)p[W + 1] * hr * wr;
                    top[idx] = (float)v;
                }
            }
    }
}

void oracle_roi_align_backward(const float* top_diff, int B, int C, int H, int W,
                               const float* rois, int R, float scale, int AH, int AW,
                               float* bottom_diff)
{
    (void)B;
    for (int n = 0; n < R; ++n) {
        const float* roi = rois + 5 * n;
        int b = (int)roi[0];
        for (int c = 0; c < C; ++c)
            for (int ph = 0; ph < AH; ++ph)
                for (int pw = 0; pw < AW; ++pw) {
                    int hs = 0, ws = 0; float hr = 0, wr = 0;
                    if (!roi_align_sample(roi, scale, AH, AW, H, W, ph, pw, &hs, &ws, &hr, &wr)) continue;
                    double g = top_diff[(((size_t)n * C + c) * AH + ph) * AW + pw];
                    float* p = bottom_diff + (((size_t)b * C + c) * H + hs) * W + ws;
                    p[0] += (float)(g * (1. - hr) * (1 - wr));
                    p[1] += (float)(g * (1. - hr) * wr);
                    p[W] += (float)(g * hr * (1 - wr));
                    p[W + 1] += (float)(g * hr * wr);
                }
    }
}

/* ------------------------------------------------------------------------------------
 * RoIPool   roi_pooling/src/roi_pooling_kernel.cu:24-93 (fwd), :128-203 (bwd)
 * ------------------------------------------------------------------------------------ */
static inline void roi_pool_bin(const float* roi, float scale, int PH, int PW, int H, int W,
                                int ph, int pw, int* w4)
{
    int rsw = (int)roundf(roi[1] * scale), rsh = (int)roundf(roi[2] * scale);
    int rew = (int)roundf(roi[3] * scale), reh = (int)roundf(roi[4] * scale);
    int rw = IMAX(rew - rsw + 1, 1), rh = IMAX(reh - rsh + 1, 1);
    float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
    int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
    w4[0] = IMIN(IMAX(hs + rsh, 0), H);
    w4[1] = IMIN(IMAX(he + rsh, 0), H);
    w4[2] = IMIN(IMAX(ws + rsw, 0), W);
    w4[3] = IMIN(IMAX(we + rsw, 0), W);
}

void oracle_roi_pool_forward(const float* feat, int B, int C, int H, int W,
                             const float* rois, int R, float scale, int PH, int PW,
                             float* top, int32_t* argmax)
{
    (void)B;
    for (int n = 0; n < R; ++n) {
        const float* roi = rois + 5 * n;
        int b = (int)roi[0];
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw) {
                int w4[4];
                roi_pool_bin(roi, scale, PH, PW, H, W, ph, pw, w4);
                int empty = (w4[1] <= w4[0]) || (w4[3] <= w4[2]);
                for (int c = 0; c < C; ++c) {
                    int base = (b * C + c) * H * W;
                    float mv = empty ? 0.f : -FLT_MAX;
                    int mi = -1;
                    for (int h = w4[0]; h < w4[1]; ++h)
                        for (int w = w4[2]; w < w4[3]; ++w) {
                            float v = feat[(size_t)base + h * W + w];
                            if (v > mv) { mv = v; mi = base + h * W + w; }
                        }
                    size_t idx = (((size_t)n * C + c) * PH + ph) * PW + pw;
                    top[idx] = mv;
                    argmax[idx] = mi;
                }
            }
    }
}

/* The reference gathers per input element over all rois (:137-201); summing top_diff into
 * argmax positions is the same function (order of the float adds aside). */
void oracle_roi_pool_backward(const float* top_diff, const int32_t* argmax, int B, int C,
                              int H, int W, int R, int PH, int PW, float* bottom_diff)
{
    size_t tot = (size_t)B * C * H * W;
    memset(bottom_diff, 0, tot * sizeof(float));
    size_t nout = (size_t)R * C * PH * PW;
    for (size_t i = 0; i < nout; ++i)
        if (argmax[i] >= 0) bottom_diff[argmax[i]] += top_diff[i];
}

/* ------------------------------------------------------------------------------------
 * RoICrop (bilinear grid sampler)   roi_crop/src/roi_crop_cuda_kernel.cu:11-22, 47-109,
 * 111-194.  grid is [R, gh, gw, 2] with (y, x) order in [-1, 1]; image index = b / (R/B).
 * ------------------------------------------------------------------------------------ */
static inline void top_left(float x, int width, int* point, float* weight)
{
    float xc = (x + 1.f) * (float)(width - 1) / 2.f;
    *point = (int)floorf(xc);
    *weight = 1.f - (xc - (float)(*point));
}

void oracle_roi_crop_forward(const float* img, int B, int C, int H, int W,
                             const float* grid, int R, int gh, int gw, float* out)
{
    int per = R / B;
    for (int b = 0; b < R; ++b) {
        int bi = b / per;
        for (int yo = 0; yo < gh; ++yo)
            for (int xo = 0; xo < gw; ++xo) {
                const float* g = grid + (((size_t)b * gh + yo) * gw + xo) * 2;
                int y0, x0; float wy, wx;
                top_left(g[1], W, &x0, &wx);
                top_left(g[0], H, &y0, &wy);
                int xin0 = x0 >= 0 && x0 <= W - 1, xin1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
                int yin0 = y0 >= 0 && y0 <= H - 1, yin1 = y0 + 1 >= 0 && y0 + 1 <= H - 1;
                int any = (xin0 || xin1) && (yin0 || yin1);
                for (int c = 0; c < C; ++c) {
                    size_t o = (((size_t)b * C + c) * gh + yo) * gw + xo;
                    if (!any) { out[o] = 0.f; continue; }   /* output is pre-zeroed, :84-85 */
                    const float* p = img + ((size_t)bi * C + c) * H * W;
                    float tl = (xin0 && yin0) ? p[y0 * W + x0] : 0.f;
                    float tr = (xin1 && yin0) ? p[y0 * W + x0 + 1] : 0.f;
                    float bl = (xin0 && yin1) ? p[(y0 + 1) * W + x0] : 0.f;
                    float br = (xin1 && yin1) ? p[(y0 + 1) * W + x0 + 1] : 0.f;
                    out[o] = wx * wy * tl + (1 - wx) * wy * tr + wx * (1 - wy) * bl +
                             (1 - wx) * (1 - wy) * br;
                }
            }
    }
}

void oracle_roi_crop_backward(const float* gout, int B, int C, int H, int W,
                              const float* grid, int R, int gh, int gw, float* gimg)
{
    int per = R / B;
    for (int b = 0; b < R; ++b) {
        int bi = b / per;
        for (int c = 0; c < C; ++c)
            for (int yo = 0; yo < gh; ++yo)
                for (int xo = 0; xo < gw; ++xo) {
                    const float* g = grid + (((size_t)b * gh + yo) * gw + xo) * 2;
                    int y0, x0; float wy, wx;
                    top_left(g[1], W, &x0, &wx);
                    top_left(g[0], H, &y0, &wy);
                    int xin0 = x0 >= 0 && x0 <= W - 1, xin1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
                    int yin0 = y0 >= 0 && y0 <= H - 1, yin1 = y0 + 1 >= 0 && y0 + 1 <= H - 1;
                    float go = gout[(((size_t)b * C + c) * gh + yo) * gw + xo];
                    float* p = gimg + ((size_t)bi * C + c) * H * W;
                    if (xin0 && yin0) p[y0 * W + x0] += wx * wy * go;
                    if (xin1 && yin0) p[y0 * W + x0 + 1] += (1 - wx) * wy * go;
                    if (xin0 && yin1) p[(y0 + 1) * W + x0] += wx * (1 - wy) * go;
                    if (xin1 && yin1) p[(y0 + 1) * W + x0 + 1] += (1 - wx) * (1 - wy) * go;
                }
    }
}

/* ------------------------------------------------------------------------------------
 * RPN proposal decode + clip   rpn/bbox_transform.py:108-134, :156-173 and the anchor
 * enumeration of rpn/proposal_layer.py:80-103 (shift (x*stride, y*stride), y outer,
 * x inner, anchor index fastest).  deltas [B, 4A, H, W], scores taken by the caller.
 * Arithmetic as torch evaluates it op by op in fp32 (no fusion).
 * ------------------------------------------------------------------------------------ */
void oracle_proposal_decode(const float* anchors /* [A,4] */, int A, const float* deltas,
                            const float* im_info /* [B,3] */, int B, int H, int W,
                            int stride, float* boxes /* [B, H*W*A, 4] */)
{
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int a = 0; a < A; ++a) {
                    float ax1 = anchors[a * 4 + 0] + (float)(x * stride);
                    float ay1 = anchors[a * 4 + 1] + (float)(y * stride);
                    float ax2 = anchors[a * 4 + 2] + (float)(x * stride);
                    float ay2 = anchors[a * 4 + 3] + (float)(y * stride);
                    float w = ax2 - ax1 + 1.0f, h = ay2 - ay1 + 1.0f;
                    float hw = 0.5f * w, hh = 0.5f * h;
                    float cx = ax1 + hw, cy = ay1 + hh;
                    const float* d = deltas + ((size_t)b * 4 * A + 4 * a) * H * W + (size_t)y * W + x;
                    float dx = d[0], dy = d[(size_t)H * W], dw = d[2 * (size_t)H * W], dh = d[3 * (size_t)H * W];
                    float t0 = dx * w, t1 = dy * h;
                    float pcx = t0 + cx, pcy = t1 + cy;
                    float pw = expf(dw) * w, ph = expf(dh) * h;
                    float hpw = 0.5f * pw, hph = 0.5f * ph;
                    float xmax = im_info[b * 3 + 1] - 1.f, ymax = im_info[b * 3 + 0] - 1.f;
                    float* o = boxes + (((size_t)b * H * W + (size_t)y * W + x) * A + a) * 4;
                    o[0] = fminf(fmaxf(pcx - hpw, 0.f), xmax);
                    o[1] = fminf(fmaxf(pcy - hph, 0.f), ymax);
                    o[2] = fminf(fmaxf(pcx + hpw, 0.f), xmax);
                    o[3] = fminf(fmaxf(pcy + hph, 0.f), ymax);
                }
}
