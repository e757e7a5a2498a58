// conv.cu -- fused convolution (+ folded BatchNorm / bias, + residual, + ReLU) as a TMA-fed
// tcgen05 implicit GEMM for sm_100a.
//
// Replaces the cuDNN calls the reference makes through torch.nn for the dilated ResNet-101 trunk
// and the R-FCN / RPN / tracking heads (/root/reference/lib/model/faster_rcnn/resnet.py:66-129,
// 258-312, 333-344; rfcn.py:49-53; rpn/rpn.py:28-36, 62-71): 1x1 convs (stride 1 and 2), 3x3 convs
// with dilation 1 / 2 / 6, each followed by an eval-mode BatchNorm (a per-channel affine,
// resnet.py:290-295) or a bias, an optional residual add and an optional ReLU.
//
// GEMM view.  D[pixel, cout] = sum over (tap r,s) and cin of  X[n, oh*st - pad + r*dil,
// ow*st - pad + s*dil, cin] * W[cout, r, s, cin].
//   M = 128 output pixels = a TH x TW box of one image (TW a power of two, TH*TW = 128);
//   N = BN output channels; K runs over taps x 32-channel blocks.
// Activations are NHWC, so for one (tap, channel block) the A operand of a tile is a
// [1, TH, TW, 32] box of the input tensor: ONE 4-D TMA copy (cp.async.bulk.tensor, 128-byte
// swizzle) whose start coordinate carries the tap offset, whose element strides carry the conv
// stride, and whose out-of-bounds zero fill IS the conv padding -- no im2col buffer, no
// per-element address math.  B is a [BN, 32] box of the packed weight matrix [Cout, R*S*Cin].
// Both land in shared memory in exactly the K-major SWIZZLE_128B layout tcgen05.mma consumes.
//
// Precision.  The reference computes these convolutions in fp32.  tcgen05 has no fp32 kind;
// kind::tf32 reads fp32 containers and IGNORES the low 13 mantissa bits (verified on B200:
// scripts/tf32_trunc_probe.py gives bit-identical results with those bits cleared or not).  An
// operand x is therefore used twice: as itself -- which the tensor core reads as hi = trunc13(x) --
// and as lo = x - trunc13(x), exactly representable.  A K-block issues THREE MMAs,
// hi*hi + hi*lo + lo*hi ("3xTF32"); the dropped lo*lo term is <= 2^-22 relative, i.e. fp32-level
// accuracy at one third of the TF32 rate -- still ~5x the fp32 SIMT pipe.  Weights are packed once
// as (w, w_lo) matrices; ACTIVATIONS ARE PLAIN fp32 NHWC TENSORS in HBM: only x travels through
// the TMA, and two "converter" warps derive the lo tile in shared memory (same swizzled address,
// elementwise) while the stage waits for its turn -- activation traffic is not doubled.
// PASSES = 1 runs the plain single-pass TF32 conv (~1e-3 relative) and is reported separately,
// never as the parity number.
//
// PASSES = 16 ("3xFP16") is the same three-product scheme on kind::f16, which runs at TWICE the TF32 rate:
// x*sa = hi + lo with hi = the top 11 significant bits and lo the rounded remainder, both fp16.  fp16 has
// fp32's precision budget for such a pair (11 + 11 bits) but not its range, so every operand tensor carries a
// power-of-two scale: weights are packed once with sw = 2^w_exp (max |w| * sw in [2^14, 2^15)); activations
// keep a per-tensor running max |x| (one float in HBM, written by the producing kernel's epilogue with one
// atomicMax per warp and tile) from which the consumer derives sa = 2^ea the same way.  Scaling by a power of
// two is exact, the epilogue multiplies the fp32 accumulator by 2^-(ea + w_exp) before the folded BatchNorm.
// The fp32 activation tile travels through the TMA unchanged (two 32-channel sub-tiles per 64-channel K block);
// four converter warps rewrite it IN PLACE as the (hi | lo) fp16 operand tiles -- 4 bytes per element either
// way, and a thread pair owns one pixel row, so no second buffer is needed.
//
// Kernel shape (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer (one lane per operand copy), NS-stage ring of {A x, A lo, B x, B lo}
//   warp 1   : MMA issuer  (one elected lane), tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8
//   warp 2   : TMEM allocation (two ping-pong chunk accumulators + two cross-term accumulators)
//   warps 2-3: converters (3-pass mode): lo = x - trunc13(x) for the activation tile(s) of each landed stage
//   warps 4-7: accumulate + epilogue, one TMEM lane (= one output pixel) per thread: every finished
//              K chunk is pulled out of TMEM (tcgen05.ld) and added into fp32 registers, then
//              scale/shift -> + residual -> ReLU -> re-split into hi/lo -> NHWC stores (and / or a
//              plain fp32 NCHW copy for the consumers that keep the reference's layout: correlation,
//              PSRoI, the proposal step) -- all while the tensor core is already on the next tile.
#include <cuda.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstddef>
#include <cstring>
#include <new>

#include "common.cuh"

namespace d2t {
namespace {

constexpr int kBlockM = 128;       // output pixels per tile (TMEM lanes)
constexpr int kBoxC = 32;          // fp32 channels per activation TMA box = one 128-byte swizzle row
constexpr int kEpiWarp0 = 4;       // warps 4..11 are the epilogue: two groups of 4 (warp % 4 = TMEM lane quarter),
constexpr int kEpiThreads = 256;   // group g owns columns [g*BN/2, (g+1)*BN/2) of the tile
constexpr int kChunkC = 256;       // channels (x taps) accumulated in TMEM before the partial sum moves to registers

// channels per K block / K blocks per TMEM chunk for a given operand mode (PASSES: 1 = TF32, 3 = 3xTF32, 16 = 3xFP16)
__host__ __device__ constexpr int kblk_of(int passes) { return passes == 16 ? 64 : 32; }
__host__ __device__ constexpr int chunk_of(int passes) { return kChunkC / kblk_of(passes); }
// Stream-K work unit in K blocks: one accumulation chunk.  The scheduler and every role take any unit (a segment may start
// or end inside a chunk; the accumulation chunks stay aligned to absolute K positions), and units of ONE K block were
// measured: 608 chunks on 148 CTAs leave 16 CTAs with 5 chunks against 4, which K-block units even out -- but the layer
// times did not move (5.57 -> 5.55 ms per step, within noise: the stragglers are the CTAs with two fix-up epilogues, not
// the ones with an extra chunk, profiles/r02_chain_timeline_trace.txt), so the unit stays the chunk.
#ifndef D2T_SK_UNIT_KBLOCK
__host__ __device__ constexpr int unit_of(int /*k_iters*/, int chunk) { return chunk; }
#else
__host__ __device__ constexpr int unit_of(int k_iters, int chunk) { return k_iters <= chunk ? chunk : 1; }
#endif

struct ConvArgs {
    int N, OH, OW, Cout;
    int R, S, stride, pad, dil;
    int kc_blocks;                 // Cin / 32
    int TW_log2, TH;               // tile = TH x (1 << TW_log2) pixels
    int tiles_h, tiles_w, m_tiles, n_tiles;
    int stem;                      // 1: A comes through the rank-5 "row window" map of the 7x7 stride-2 stem
    const float* scale;            // [Cout] or null (= 1)
    const float* shift;            // [Cout] or null (= 0)
    const float* res;              // NHWC [N, OH, OW, res_cstride] or null
    int res_cstride;
    int relu;
    float* out;                    // NHWC [N, OH, OW, out_cstride], channels [out_coffset, +Cout); or null
    int out_cstride, out_coffset;
    float* out_nchw;               // plain fp32 [N, Cout, OH, OW] or null
    // correlation mode (CORR): B operand = halo rows of the second frame, see corr section below
    int corr_r, corr_D;            // displacement radius (lattice units) and 2r+1
    float corr_nelems;             // kernel_size^2 * C
    // stream-K fix-up (see Sched): per-CTA partial tiles [grid][BN][128] and their "published" flags
    float* sk_scratch;
    int* sk_flags;
    int sk_epoch;
    // per-tensor running max |x| (one float each): amax_in scales the fp16-split activation operand, amax_out
    // collects this layer's output for its consumers; w_exp = log2 of the packed weights' scale (3xFP16 only)
    const float* amax_in;
    float* amax_out;
    int w_exp;
    // 3xFP16 with a B operand whose scale lives on the device (weights re-packed every training step, or -- WGRAD -- the
    // pre-split gradient planes): its max |b| (one float); the packer and this kernel derive the same 2^k from it
    const float* amax_b;
    // backward-data: ReLU mask -- out = mask[pixel, channel] > 0 ? value : 0, applied after the residual add (the mask
    // tensor is the forward activation whose gradient this launch produces; same geometry as the output)
    const float* mask;
    int mask_cstride;
    // WGRAD (see the kernel comment): K blocks per output row, K blocks in total, input channels, OIHW gradient
    int wg_xblocks, wg_kiters, wg_cin;
    float* wg_out;
    // WGRAD with a separate reduction launch (d2t_wgrad_plan_set_partials): a weight-gradient GEMM has few tiles and a very
    // long K (all pixels), so every tile is split over ~9 CTAs and the in-kernel finisher would read 8 partial tiles one
    // after the other through one SM's load path (30 B/clk: ~20 us per launch, profiles/r02_mb_*).  Instead every CTA
    // stores its partial tiles ([cta][slot: 0 = its first tile, 1 = its last][column][row]) and wgrad_reduce, a device-wide
    // launch, adds them in CTA order.
    float* wg_partials;
    // optional completion hand-shake between consecutive launches of one chain (d2t_conv_plan_set_done): every CTA adds
    // 1 to *done_self when all its outputs are globally visible; a launch whose done_prev is set polls that counter up
    // to done_target (= the previous launch's grid) INSTEAD of griddepcontrol.wait, i.e. it does not sit through the
    // hardware's grid-completion latency (~3.4 us per dependent launch, DESIGN section 6 finding 4)
    const int* done_prev;
    int done_target;
    int* done_self;
    // stand-alone launches: the weight (B operand) copies of the first pipeline stages are issued BEFORE griddepcontrol.wait --
    // packed weights do not depend on the previous launch, and their first touch per forward is a DRAM miss that otherwise
    // sits in front of the first MMA (d2t_conv_plan_set_early_weights; never for plans whose B operand is produced upstream)
    int early_b;
    long long* trace;              // debug builds only (D2T_CONV_TRACE): per-CTA wait-cycle counters
    int exp;                       // debug builds only: experiment bit mask (env D2T_CONV_EXP)
};
#ifdef D2T_CONV_TRACE
#define EXP(bit) (p.exp & (bit))
#else
#define EXP(bit) false
#endif

// Work distribution ("stream-K").  A layer is tiles x cpt units, a unit = one K chunk (256 channels x taps) of
// one output tile.  CTA c owns the contiguous unit range [c*U/G, (c+1)*U/G): every CTA gets the same amount of
// tensor-core work (+-1 chunk) whatever the tile count -- with whole tiles, 152 or 304 tiles on 148 SMs cost
// 2 or 3 rounds for 1.03 or 2.05 rounds of work.  A tile split across CTAs is finished by the CTA holding its
// LAST chunk: the others publish their fp32 partial tile (registers -> L2-resident scratch) and the finisher
// adds them in k order (deterministic) before the fused epilogue.  A CTA runs its trailing partial tile
// FIRST and its leading partial tile LAST, so a finisher never waits on a partial that is not long published,
// and waits only ever point to lower CTA indices (all CTAs are co-resident: grid <= #SMs, 1 CTA/SM).
struct Seg {
    int tile, c0, c1, role;        // chunks [c0, c1) of `tile`; role 0 = whole tile, 1 = publish partial, 2 = finish
};
struct Sched {
    int cpt, nseg, first_tile, tail_pub, head_fin;
    long long u0, u1;
    __device__ Sched(int tiles, int k_iters, int chunk, int cta, int G) {
        cpt = (k_iters + chunk - 1) / chunk;
        const long long U = (long long)tiles * cpt;
        u0 = U * cta / G;
        u1 = U * (cta + 1) / G;
        nseg = 0; first_tile = 0; tail_pub = 0; head_fin = 0;
        if (u1 > u0) {
            first_tile = (int)(u0 / cpt);
            const int last_tile = (int)((u1 - 1) / cpt);
            nseg = last_tile - first_tile + 1;
            tail_pub = (u1 - (long long)last_tile * cpt) < cpt;
            head_fin = (u0 - (long long)first_tile * cpt) > 0 && (nseg > 1 || !tail_pub);
        }
    }
    __device__ Seg get(int e) const {          // e-th segment in EXECUTION order
        int nat;
        int e2 = e;
        if (tail_pub && e == 0) {
            nat = nseg - 1;
        } else {
            if (tail_pub) e2 = e - 1;
            const int mid = (tail_pub ? nseg - 1 : nseg) - head_fin;
            nat = e2 < mid ? e2 + head_fin : 0;
        }
        Seg g;
        g.tile = first_tile + nat;
        const long long base = (long long)g.tile * cpt;
        g.c0 = (int)((u0 > base ? u0 : base) - base);
        g.c1 = (int)((u1 < base + cpt ? u1 : base + cpt) - base);
        g.role = g.c1 < cpt ? 1 : (g.c0 > 0 ? 2 : 0);
        return g;
    }
};

// ARES ("A resident", an EPI2 sub-variant for 1x1 layers of at most 4 K blocks): the activation tile of an m tile is converted
// into tensor memory ONCE and reused by all the n tiles the CTA runs on it (see the kernel comment).  Shared memory then
// holds one fp32 staging buffer for A, a ring of weight-only stages and the full-tile output staging.
template <int BN, int PASSES, bool PAIR = false, bool EPI2 = false, bool ARES = false>
struct Cfg {
    static constexpr bool F16 = PASSES == 16;
    static constexpr bool SPLIT = PASSES != 1;                        // three products per K step
    static constexpr int KBLK = kblk_of(PASSES);                      // channels per K block
    static constexpr int CHUNK = chunk_of(PASSES);                    // K blocks per TMEM chunk
    // 3xFP16: 16 warps -- warps 12-13 are two more converters (14-15 idle) -- so that the register file can be
    // re-split by warpgroup (setmaxnreg): 80 registers for the TMA / MMA / converter warps, 176 for the epilogue
    static constexpr int THREADS = F16 ? 512 : 384;
    static constexpr int CVT_THREADS = F16 ? 128 : 64;
    // 3xFP16: A = two fp32 [128 x 32] sub-tiles as loaded, rewritten in place as fp16 [128 x 64] hi | lo
    static constexpr int A_BYTES = kBlockM * KBLK * 4;                // 16 KB (32 KB)
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * KBLK * (F16 ? 2 : 4);   // (a CTA pair holds half of B each)
    static constexpr int NOPER = PASSES == 3 ? 2 : 1;                 // TF32: hi (+ lo) copies of each operand
    // smem per stage -- TF32: A x | A lo | B x | B lo;  FP16: A (hi | lo) | B hi | B lo
    static constexpr int STAGE_BYTES = ARES ? 2 * B_BYTES : (F16 ? A_BYTES + 2 * B_BYTES : NOPER * (A_BYTES + B_BYTES));
    static constexpr int A_STAGING = ARES ? A_BYTES : 0;              // ARES: one fp32 activation K block in front of the ring
    static constexpr int OFF_ALO = F16 ? A_BYTES / 2 : A_BYTES;
    static constexpr int OFF_BHI = F16 ? A_BYTES : NOPER * A_BYTES;
    static constexpr int OFF_BLO = OFF_BHI + B_BYTES;
    // output staging: [128 x 32] fp32 slabs (TMA store; EPI2: also the TMA-loaded residual).  Default: one slab per
    // epilogue group, reused by the BN/64 slabs of a tile.  EPI2 (layers whose time is the epilogue, not the K loop: short
    // K, residual): one buffer for EVERY slab of the tile, paid for with one pipeline stage.
    static constexpr int OUT_SLABS = EPI2 ? BN / 64 : 1;             // buffers per group
    static constexpr int OUT_BUF_BYTES = 2 * OUT_SLABS * kBlockM * 128;
    // ARES: TWO full-tile buffers -- the residual of tile i + 1 lands while tile i is being added / stored, and no TMA wait
    // sits on the epilogue's critical path
    static constexpr int OUT_STAGE_BYTES = (ARES ? 2 : 1) * OUT_BUF_BYTES;
    static constexpr int STAGES_RAW = (228 * 1024 - OUT_STAGE_BYTES - A_STAGING) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int OUT_OFF = A_STAGING + STAGES * STAGE_BYTES;  // output staging follows the operand stages
    static constexpr int SMEM_BYTES = OUT_OFF + OUT_STAGE_BYTES + 1024 /*align slack*/ + 384 /*barriers*/ +
                                      1024 /*scale | shift of the current n tile*/ + (ARES ? 512 : 0) /*third shift buffer*/;
    // barrier block (8-byte slots from `bars`)
    static constexpr int NCVT = ARES ? 4 : STAGES;
    static constexpr int BAR_CVT = 2 * STAGES + 6, BAR_RFULL = BAR_CVT + NCVT, BAR_TSLOT = BAR_RFULL + 2;
    static constexpr int BAR_AFULL = BAR_TSLOT + 1, BAR_AEMPTY = BAR_AFULL + 1, BAR_AFREE = BAR_AEMPTY + 1, BAR_RFULL2 = BAR_AFREE + 4;
    static_assert((BAR_RFULL2 + 2) * 8 <= 384, "barrier block");
    // TMEM columns: main[2] chunk buffers (+ cross[2] whole-tile buffers in 3-pass mode), BN each
    // TMEM columns.  TF32: main[2] chunk accumulators (+ cross[2] whole-tile accumulators in 3-pass mode), BN each.
    // 3xFP16: main[2] only (the cross terms join the chunk accumulator) + the A OPERAND RING: per pipeline stage 32 columns
    // of hi and 32 of lo (128 lanes = pixel rows x 64 channels of packed fp16), written by the converter warps with
    // tcgen05.st and read by tcgen05.mma directly -- the tensor core's A reads leave shared memory altogether.
    static constexpr bool ATMEM = F16;
    static constexpr int A_TMEM_COLS = 64;
    static constexpr int A_TMEM_BASE = 2 * BN;
    static constexpr int A_SLOTS = ARES ? 4 : STAGES;                 // ARES: slot = K block of the tile, resident across n tiles
    static constexpr int TMEM_USED = ATMEM ? 2 * BN + A_SLOTS * A_TMEM_COLS : (SPLIT ? 4 : 2) * BN;
    static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512);
    static_assert(TMEM_USED <= 512, "tensor memory budget");
    static_assert(!ARES || (EPI2 && F16 && !PAIR && BN == 128 && STAGES >= 2), "ARES is an EPI2 sub-variant");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d_nocommit(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// ---- CTA-pair (cta_group::2) forms.  `mbar_leader` is the shared::cluster address of the LEADER CTA's barrier.
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t mbar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes
// apart (SBO), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;              // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
    d |= (uint64_t)1 << 46;              // version
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = fp32, A = B = tf32, both K-major, M = 128, N = BN
template <int BN, int M = kBlockM>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D = fp32, A = B = fp16, both K-major
template <int BN, int M = kBlockM>
__device__ __forceinline__ constexpr uint32_t make_idesc_f16() {
    return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 MMA with the A operand in tensor memory (lane = row, 2 fp16 per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same across a CTA pair: M = 256 (each CTA's tensor memory holds its own 128 rows of A and of D), each CTA's shared
// memory holds N/2 rows of B
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// M = 256 across the CTA pair: rows 0..127 come from / go to the leader, 128..255 the peer; each CTA supplies N/2 rows of B
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {     // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the same load without the wait: several may be in flight before one tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// mbarrier.try_wait already suspends the thread for a hardware-chosen interval, so the poll loop is a plain spin
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Debug builds (make trace: -DD2T_CONV_TRACE, libd2t_b200_trace.so) account the cycles every role spends in each of its
// waits; slot layout in scripts/conv_trace.py.  Compiled out of the product library.
#ifdef D2T_CONV_TRACE
#define TRACE_DECL long long trace_t0__ = 0, trace_acc__[4] = {0, 0, 0, 0}
#define TRACED_WAIT(slot, bar, parity)            \
    do {                                          \
        const long long t__ = clock64();          \
        mbar_wait_sleep(bar, parity);             \
        trace_acc__[slot] += clock64() - t__;     \
    } while (0)
#define TRACE_BEGIN() trace_t0__ = clock64()
#define TRACE_FLUSH(role)                                                                        \
    do {                                                                                         \
        if (p.trace && lane == 0) {                                                              \
            long long* d__ = p.trace + ((size_t)blockIdx.x * 8 + (role)) * 8;                    \
            d__[0] = clock64() - trace_t0__;                                                     \
            d__[1] = trace_acc__[0]; d__[2] = trace_acc__[1]; d__[3] = trace_acc__[2]; d__[4] = trace_acc__[3]; \
        }                                                                                        \
    } while (0)
#else
#define TRACE_DECL
#define TRACED_WAIT(slot, bar, parity) mbar_wait_sleep(bar, parity)
#define TRACE_BEGIN()
#define TRACE_FLUSH(role)
#endif

// ------------------------------------------------------------------ the kernel
// CORR = true turns the same pipeline into the cross-frame correlation (correlation/src/
// correlation_cuda_kernel.cu:34-106, kernel_size 1, stride1 == stride2): for a tile of 8x16 positions p of
// frame t and a chunk of 4x32 "halo" positions q of frame t+tau, D[p, q] = sum_c in1[p, c] * in2[q, c] is
// a Gram block -- a GEMM whose B operand is a second 4-D activation box instead of a weight tile.  The
// (2r+1)^2 displacements of p are the q with |q - p| <= r in both axes; the epilogue keeps those (38 % of the
// block for r = 8) and writes out[n, (tj+r)*D + (ti+r), y, x] = D / C.  Zero padding outside the frame is
// again the TMA out-of-bounds fill; strided correlation (conv3: stride 2) is the TMA element stride.
//
// PAIR = true runs the kernel as CTA pairs (2-CTA clusters, tcgen05 cta_group::2): one MMA covers 256 output
// pixels (two adjacent pixel tiles, one per CTA) x BN channels; each CTA stages only its own A tile and HALF of
// the weight tile, the tensor cores of the two SMs exchange the B halves: 25 % less TMA traffic per useful FLOP
// and one more pipeline stage.  MEASURED on B200 (round 1): no faster than single-CTA mode on these layers
// (head conv 0.89 vs 0.87 ms) -- the MMA pipe itself reaches ~95 % of the TF32 peak when fed (MMA-repeat
// experiment), and the ~40 % main-loop overhead is additive per K block and insensitive to operand bytes (pair
// mode) and to L2 reads (2x2 TMA-multicast clusters were tried too).  So pairs are OFF by default
// (D2T_CONV_PAIR=1 enables them) and kept as the validated cta_group::2 path.  The leader CTA issues the MMAs and owns the
// `full` / `tempty` / `xempty` barriers (both CTAs' TMA and epilogue warps signal them remotely); its commits
// are multicast to both CTAs' `empty` / `tfull` barriers.
//
// WGRAD = true (3xFP16 only) is the weight gradient of a stride-1 convolution on the same pipeline,
//   dW[co, ci, r, s] = scale[co] * sum over (image, oy, ox) of  X[image, ci, oy + r*dil - pad, ox + s*dil - pad] * G[image, co, oy, ox]
// as the GEMM  D[m = ci, n = co] per filter tap with K running over the output PIXELS.  Both operands are read from
// PLANAR copies (channel-major planes, row pitch a multiple of 16 bytes) so that K is the contiguous axis, exactly the
// K-major SWIZZLE_128B tiles the forward pass uses: A = a [128 channels x 64 pixels] fp32 box of X (two 32-pixel TMA
// boxes), split into fp16 (hi, lo) by the converter warps into tensor memory like a forward activation tile; B =
// [BN channels x 64 pixels] boxes of the gradient's pre-split fp16 (hi, lo) planes (d2t_wgrad_pack_grad).  The filter
// tap's ROW offset is the start coordinate of the A box in y (zero padding = the TMA's out-of-bounds fill); its COLUMN
// offset cannot be a start coordinate -- a box must start on a 16-byte boundary of the contiguous axis (an odd start
// raises an illegal-instruction fault, measured) -- so the packer writes one copy of the gradient planes per filter
// column, shifted by that column's offset, and K runs over the columns u = ox + dx of X.  A K block is 64 consecutive
// columns of one row;
// tiles are (128 input channels, tap) x (BN output channels); stream-K and the drain are unchanged; the epilogue
// scales column co by the folded BatchNorm scale and writes the OIHW gradient.
//
// CORRB = true (3xFP16 only) is the BACKWARD of the cross-frame correlation (correlation_cuda_kernel.cu:108-290) on the
// same pipeline.  With t = (tj, ti) a displacement,  g1[p, c] = (1/C) sum_t gO[p, t] * in2[p + t, c]  is, for a tile of
// 4 x 32 positions p, the GEMM  D[p, c] = sum_q Band[p, q] * in2[q, c]  over the halo positions q (K = one 64-wide halo
// row per K block, 4 + 2r of them), where Band[p, q] = gO[p, q - p] inside the (2r+1)^2 window and 0 outside.  Both
// operands arrive pre-split as fp16 (hi, lo): B = boxes of the other frame's channel-major planes (like WGRAD); A = the
// band, expanded by d2t_corrb_pack_band to [pixel][tj + r][64 halo columns] so that a tile row's K block is ONE TMA box
// whose displacement-row coordinate is (halo row - tile row) -- rows outside the window are the TMA's zero fill.  The
// converter warps only move the landed A tile into tensor memory.  The gradient w.r.t. the second frame is the same GEMM
// on the flipped band  gO[p + t, -t]  and the first frame's planes.  Epilogue: the ordinary NHWC output path.
//
// MASK = true compiles the backward-data epilogue (ReLU mask of the forward activation, d2t_conv_plan_set_mask) in.  It is
// a template parameter, not a run-time test, because the forward pass pays for the extra code in its hottest loop even
// when the pointer is null: measured on one B200 box, 5.27 ms against 4.98 ms for the 110 forward launches of the step.
// CHAIN = true compiles the body for the multi-layer persistent kernel (conv_chain below): the pipeline barriers are
// (re)initialised here, per layer; tensors written by earlier layers of the same launch are only read through the TMA or
// with L2-coherent loads; a publisher waits for its stream-K flag to be consumed before it reuses its scratch slot (layers
// that do not depend on each other run without a grid-wide barrier between them); no setmaxnreg.
template <int BN, int PASSES, bool CORR, bool PAIR, bool EPI2, bool WGRAD, bool CORRB, bool MASK, bool CHAIN, bool ARES = false>
__device__ __forceinline__ void conv_body(const CUtensorMap& tmA,      // activation (A operand)
                                          const CUtensorMap& tmB_hi,   // weights w (CORR: the second frame's activation)
                                          const CUtensorMap& tmB_lo,   // weights w_lo (unused in CORR / 1-pass mode)
                                          const CUtensorMap& tmO,      // NHWC output (TMA store)
                                          const CUtensorMap& tmR,      // NHWC residual (EPI2: TMA load into the output slabs)
                                          const ConvArgs& p, uint8_t* smem /*1024-B aligned*/, uint64_t* bars,
                                          const uint32_t tmem_base, const uint32_t crank, const int prev_nbars,
                                          const Sched* pre_sched = nullptr) {
    using C = Cfg<BN, PASSES, PAIR, EPI2, ARES>;
    static_assert(!EPI2 || (PASSES == 16 && !CORR && !PAIR), "EPI2 is a 3xFP16 convolution variant");
    static_assert(!ARES || (!CHAIN && !WGRAD && !CORRB && !CORR), "ARES is a stand-alone convolution variant");
    constexpr bool F16 = C::F16, SPLIT = C::SPLIT, ATMEM = C::ATMEM;
    constexpr int kChunkK = C::CHUNK, kBlockK = C::KBLK, kCvtThreads = C::CVT_THREADS;
    static_assert(!(F16 && PAIR) || (BN == 128 && !CORR && !WGRAD && !CORRB && !MASK), "3xFP16 pairs: plain forward convolutions");
    static_assert(!WGRAD || (F16 && !EPI2), "the weight-gradient mode is a plain 3xFP16 variant");
    static_assert(!CORRB || (F16 && !EPI2 && !WGRAD && BN == 128), "the correlation-backward mode is a plain 3xFP16 variant");
    static_assert(!CHAIN || (F16 && !CORR && !PAIR && !WGRAD && !CORRB), "the chain kernel runs 3xFP16 convolutions");
    uint8_t* out_stage = smem + C::OUT_OFF;                                   // [2 groups][128 rows][128 B], swizzled
    uint64_t* full = bars;                        // [STAGES]   TMA -> MMA
    uint64_t* empty = bars + C::STAGES;           // [STAGES]   MMA -> TMA
    uint64_t* tfull = bars + 2 * C::STAGES;       // [2]        MMA -> epilogue: chunk buffer complete
    uint64_t* tempty = tfull + 2;                 // [2]        epilogue -> MMA: chunk buffer drained
    uint64_t* xempty = tempty + 2;                // [2]        epilogue -> MMA: cross-term buffer read
    uint64_t* cvt = bars + C::BAR_CVT;            // [STAGES]   converters -> MMA: lo tile(s) of the stage written (ARES: [4], per A slot)
    uint64_t* rfull = bars + C::BAR_RFULL;        // [2]        TMA -> epilogue group: residual slabs landed (EPI2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::BAR_TSLOT);
    // ARES only: the A staging buffer's hand-shake, "slot k may be overwritten" (the MMAs that read it have retired), and the
    // residual barriers of the odd tiles (two output buffers in flight)
    uint64_t* afull = bars + C::BAR_AFULL;        // [1]        TMA -> converters: fp32 K block landed
    uint64_t* aempty = bars + C::BAR_AEMPTY;      // [1]        converters -> TMA: staging buffer read
    uint64_t* afree = bars + C::BAR_AFREE;        // [4]        MMA -> converters
    uint64_t* rfull2 = bars + C::BAR_RFULL2;      // [2]
    (void)afull; (void)aempty; (void)afree; (void)rfull2;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CS = PAIR ? 2 : 1;               // CTAs per work unit
    const int tiles = ((p.m_tiles + CS - 1) / CS) * p.n_tiles;     // (pair-)tiles
    const int k_iters = WGRAD ? p.wg_kiters : (CORRB ? p.TH + 2 * p.corr_r : p.R * p.S * p.kc_blocks);
    // the stream-K split needs every participating CTA to own at least one unit (a finisher waits for ALL CTAs between the
    // tile's first chunk and itself): the stand-alone launch sizes its grid accordingly (sk_grid), a chain layer with fewer
    // units than CTAs leaves the surplus CTAs idle
    const int kUnit = unit_of(k_iters, kChunkK);
    const long long units_total = (long long)tiles * ((k_iters + kUnit - 1) / kUnit);
    const int unit_id = blockIdx.x / CS;
    const int n_units = CHAIN ? (int)(units_total < (long long)gridDim.x ? units_total : (long long)gridDim.x) : gridDim.x / CS;
    if constexpr (CHAIN) {
        // per-layer pipeline reset: every role has left the previous layer (the caller's __syncthreads), no arrival is
        // pending on any barrier (the MMA thread waited for its last commits), so the barriers restart at phase 0
        if (warp == 1 && lane == 0) {
            for (int i = 0; i < prev_nbars; ++i)          // (the previous layer's variant may have laid out more or fewer)
                asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bars + i)) : "memory");
            for (int i = 0; i < C::STAGES; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < C::NCVT; ++i) mbar_init(&cvt[i], kCvtThreads);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&tfull[i], 1);
                mbar_init(&tempty[i], kEpiThreads);
                mbar_init(&xempty[i], kEpiThreads);
                mbar_init(&rfull[i], 1);
            }
            fence_mbar_init();
        }
        if (warp == 0 && lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmB_hi);
            prefetch_tmap(&tmB_lo);
            if (p.out) prefetch_tmap(&tmO);
            if (EPI2 && p.res) prefetch_tmap(&tmR);
        }
        __syncthreads();
        asm volatile("fence.proxy.async;" ::: "memory");      // TMA reads of this layer come after the caller's acquire
        tc_fence_after();
#ifdef D2T_CONV_TRACE
        if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 64 + 5] = clock64();          // role 0 value 5: body start
#endif
    }
    // (a stand-alone launch computes its schedule -- 64-bit divisions -- before griddepcontrol.wait and hands it in)
    Sched sched = pre_sched ? *pre_sched : Sched(tiles, k_iters, kUnit, CHAIN && unit_id >= n_units ? 0 : unit_id, n_units);
    if (CHAIN && unit_id >= n_units) sched.nseg = 0;

    if (warp < kEpiWarp0 || warp >= kEpiWarp0 + 8) {
    // ---- warpgroups 0 (and 3 in 3xFP16 mode): TMA producer, MMA issuer, converters
    if constexpr (F16) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if constexpr (ARES) {
        // ===================== A-resident variant: 1x1 layers of 2..4 K blocks, several n tiles per m tile =====================
        // Tiles run n-inner (t = m_tile * n_tiles + n_tile), every tile is whole (one chunk), so consecutive tiles of a CTA
        // mostly share their m tile: its activation K blocks are loaded + converted into tensor memory slots 0..k_iters-1 ONCE
        // ("fresh" tile) and the following tiles only stream weights -- per tile 128 KB of TMA traffic and 32 KB of converter
        // reads less, on layers whose time is shared-memory bandwidth.  Rings: A staging (1 buffer,
        // afull / aempty), weight stages (4, full / empty), A slots (cvt[k]: converted, afree[k]: last reader retired).
        const int nt = p.n_tiles;
        if (warp == 0) {
            // ---- weight producer: lanes 0 / 1 = hi / lo halves of the stage
            if (lane < 2) {
                const CUtensorMap* map = lane == 0 ? &tmB_hi : &tmB_lo;
                int stage = 0;
                uint32_t phase = 0;
                for (int e = 0; e < sched.nseg; ++e) {
                    const int n0 = (sched.get(e).tile % nt) * BN;
                    for (int k = 0; k < k_iters; ++k) {
                        mbar_wait_sleep(&empty[stage], phase ^ 1);
                        uint8_t* dst = smem + C::A_STAGING + stage * C::STAGE_BYTES + lane * C::B_BYTES;
                        if (lane == 0) mbar_expect_tx(&full[stage], 2 * C::B_BYTES);
                        tma_load_2d(dst, map, &full[stage], k * kBlockK, n0);
                        if (++stage == C::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        } else if (warp == kEpiWarp0 + 10) {
            // ---- activation producer (warp 14): the two 32-channel sub-tiles of a K block, fresh tiles only
            if (lane < 2) {
                uint32_t aphase = 0;
                int prev_m = -1;
                for (int e = 0; e < sched.nseg; ++e) {
                    const int m_tile = sched.get(e).tile / nt;
                    if (m_tile == prev_m) continue;
                    prev_m = m_tile;
                    const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, img = m_tile / (p.tiles_w * p.tiles_h);
                    const int iw0 = (tw << p.TW_log2) * p.stride - p.pad, ih0 = th * p.TH * p.stride - p.pad;
                    for (int k = 0; k < k_iters; ++k) {
                        mbar_wait_sleep(aempty, aphase ^ 1);           // the converters have read the previous K block
                        aphase ^= 1;
                        if (lane == 0) mbar_expect_tx(afull, C::A_BYTES);
                        tma_load_4d(smem + lane * (kBlockM * kBoxC * 4), &tmA, afull, k * kBlockK + lane * kBoxC, iw0, ih0, img);
                    }
                }
            }
        } else if (warp == 1) {
            // ---- MMA issuer
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc_f16<BN>();
                const uint64_t desc0 = make_smem_desc(smem_u32(smem + C::A_STAGING));
                int stage = 0, cbuf = 0, prev_m = -1, fresh_idx = -1;
                uint32_t phase = 0, cphase = 0;
                for (int e = 0; e < sched.nseg; ++e) {
                    const int m_tile = sched.get(e).tile / nt;
                    const bool fresh = m_tile != prev_m;
                    prev_m = m_tile;
                    if (fresh) ++fresh_idx;
                    const bool last_of_group = e + 1 == sched.nseg || sched.get(e + 1).tile / nt != m_tile;
                    mbar_wait_sleep(&tempty[cbuf], cphase ^ 1);        // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d_main = tmem_base + cbuf * BN;
                    for (int k = 0; k < k_iters; ++k) {
                        mbar_wait_sleep(&full[stage], phase);
                        if (fresh) mbar_wait_sleep(&cvt[k], (uint32_t)(fresh_idx & 1));
                        tc_fence_after();
                        const uint64_t b_hi = desc0 + (uint64_t)(stage * (C::STAGE_BYTES >> 4));
                        const uint64_t b_lo = b_hi + (C::B_BYTES >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t o = (uint64_t)(kk * 32 >> 4);
                            const uint32_t at_hi = tmem_base + C::A_TMEM_BASE + k * C::A_TMEM_COLS + kk * 8;
                            umma_f16_ts(d_main, at_hi, b_hi + o, idesc, (k | kk) != 0);
                            umma_f16_ts(d_main, at_hi, b_lo + o, idesc, 1);
                            umma_f16_ts(d_main, at_hi + 32, b_hi + o, idesc, 1);
                        }
                        umma_commit(&empty[stage]);
                        if (last_of_group) umma_commit(&afree[k]);     // slot k may be overwritten once these have retired
                        if (++stage == C::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(&tfull[cbuf]);
                    cbuf ^= 1;
                    if (cbuf == 0) cphase ^= 1;
                }
            }
        } else if (warp < kEpiWarp0 || warp == kEpiWarp0 + 8 || warp == kEpiWarp0 + 9) {
            // ---- converters (warps 2, 3, 12, 13 = tensor-memory lane quarters 2, 3, 0, 1): fresh tiles only
            const int m = (warp & 3) * 32 + lane;
            const uint32_t sw = (uint32_t)m & 7u;
            const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::A_TMEM_BASE;
            const float sa = pow2f(act_exp(p.amax_in));
            uint32_t aphase = 0;
            int prev_m = -1, fresh_idx = -1;
            for (int e = 0; e < sched.nseg; ++e) {
                const int m_tile = sched.get(e).tile / nt;
                if (m_tile == prev_m) continue;
                prev_m = m_tile;
                ++fresh_idx;
                for (int k = 0; k < k_iters; ++k) {
                    if (fresh_idx > 0) mbar_wait_sleep(&afree[k], (uint32_t)((fresh_idx - 1) & 1));
                    mbar_wait_sleep(afull, aphase);
                    aphase ^= 1;
                    tc_fence_after();
                    const uint8_t* row = smem + m * 128;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint8_t* src = row + half * (kBlockM * kBoxC * 4);
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 v = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sw) << 4));
                            const float a0 = v.x * sa, a1 = v.y * sa, a2 = v.z * sa, a3 = v.w * sa;
                            const float h0 = __uint_as_float(__float_as_uint(a0) & 0xffffe000u);
                            const float h1 = __uint_as_float(__float_as_uint(a1) & 0xffffe000u);
                            const float h2 = __uint_as_float(__float_as_uint(a2) & 0xffffe000u);
                            const float h3 = __uint_as_float(__float_as_uint(a3) & 0xffffe000u);
                            __half2 t;
                            t = __floats2half2_rn(h0, h1); hi[2 * c] = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2half2_rn(h2, h3); hi[2 * c + 1] = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2half2_rn(a0 - h0, a1 - h1); lo[2 * c] = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2half2_rn(a2 - h2, a3 - h3); lo[2 * c + 1] = *reinterpret_cast<uint32_t*>(&t);
                        }
                        const uint32_t slot = a_lane + k * C::A_TMEM_COLS + half * 16;
                        tmem_st16(slot, hi);
                        tmem_st16(slot + 32, lo);
                    }
                    mbar_arrive(aempty);                               // (the staging rows are in registers / tensor memory)
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    mbar_arrive(&cvt[k]);
                }
            }
        }
    } else
    if (warp == 0) {
        // ===================== TMA producer =====================
        // One lane per operand copy (A x, A lo, B x, B lo), all four walking the same K loop; (r, s, kc) advance
        // incrementally -- no divisions in the loop.  Measured (round 1): the feed alone (no MMAs issued) runs at
        // ~950 cycles per 64 KB K block with the 3-stage ring = TMA latency x bytes in flight, against ~870 cycles
        // of 3xTF32 MMA work per K block; the two overlap only partly (1450 cycles per K block end to end).
        // TMA copies per stage: A x (3xFP16: its two 32-channel sub-tiles), B x (, B lo: weights only)
        constexpr int NA = F16 ? 2 : 1;
        // B copies: weights (hi, lo); correlation: the second frame's tile -- ONE fp32 box in the TF32 kinds, two 32-channel
        // sub-tiles in 3xFP16 (they land in the B hi / B lo slots and are rewritten there as the fp16 pair, see the converters)
        constexpr int NL = NA + ((SPLIT && !CORR) || (CORR && F16) ? 2 : 1);
        constexpr int TMA_BYTES = C::A_BYTES + (NL - NA) * C::B_BYTES;
        if constexpr (CORRB) {
            // ten copies per stage: the band's (hi, lo) boxes of the four tile rows (32 pixels x 64 halo columns each),
            // the other frame's (hi, lo) planes (BN channels x 64 halo columns of one halo row)
            if (lane < 10) {
                const bool is_a = lane < 8;
                const int yl = lane & 3;
                const CUtensorMap* map = lane < 4 ? &tmA : (lane < 8 ? &tmR : (lane == 8 ? &tmB_hi : &tmB_lo));
                const int dst_off = lane < 4 ? yl * 4096 : (lane < 8 ? C::OFF_ALO + yl * 4096 : (lane == 8 ? C::OFF_BHI : C::OFF_BLO));
                int stage = 0;
                uint32_t phase = 0;
                for (int e = 0; e < sched.nseg; ++e) {
                    const Seg sg = sched.get(e);
                    const int t = sg.tile;
                    const int n_tile = t % p.n_tiles, m_tile = t / p.n_tiles;
                    const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, img = m_tile / (p.tiles_w * p.tiles_h);
                    const int x0 = tw << p.TW_log2, y0 = th * p.TH;
                    const int k_beg = sg.c0 * kUnit, k_end = min(sg.c1 * kUnit, k_iters);
                    for (int hr = k_beg; hr < k_end; ++hr) {        // K block = halo row y0 - r + hr
                        mbar_wait_sleep(&empty[stage], phase ^ 1);
                        uint8_t* dst = smem + stage * C::STAGE_BYTES + dst_off;
                        uint64_t* fbar = &full[stage];
                        if (lane == 0) mbar_expect_tx(fbar, TMA_BYTES);
                        if (is_a) tma_load_5d(dst, map, fbar, 0, hr - yl, x0, y0 + yl, img);
                        else tma_load_4d(dst, map, fbar, x0 - 16, y0 - p.corr_r + hr, n_tile * BN, img);
                        if (++stage == C::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        } else
        if (lane < NL) {
            const bool is_a = lane < NA;
            const CUtensorMap* map = is_a ? &tmA : (lane == NA ? &tmB_hi : &tmB_lo);
            const int dst_off = is_a ? lane * (kBlockM * kBoxC * 4) : (lane == NA ? C::OFF_BHI : C::OFF_BLO);
            const int a_c0 = is_a ? lane * kBoxC : 0;            // first channel of this lane's sub-tile in the K block
            int stage = 0;
            uint32_t phase = 0;
            TRACE_DECL;
            TRACE_BEGIN();
            for (int e = 0; e < sched.nseg; ++e) {
                const Seg sg = sched.get(e);
                const int t = sg.tile;
                const int n_tile = t % p.n_tiles, m_tile = (t / p.n_tiles) * CS + (int)crank;
                if constexpr (WGRAD) {
                    // tile = (tap, 128 input channels) x (BN output channels); K block k = 64 pixels of output row
                    // (image, oy) starting at column 64 * xb
                    const int RS = p.R * p.S, tap = m_tile % RS, ci0 = (m_tile / RS) * kBlockM;
                    const int dy = (tap / p.S) * p.dil - p.pad, sx = tap % p.S;     // (column shift sx: pre-shifted B planes)
                    const int k_beg = sg.c0 * kUnit, k_end = min(sg.c1 * kUnit, k_iters);
                    int xb = k_beg % p.wg_xblocks, oy = (k_beg / p.wg_xblocks) % p.OH, im = k_beg / (p.wg_xblocks * p.OH);
                    for (int k = k_beg; k < k_end; ++k) {
                        TRACED_WAIT(0, &empty[stage], phase ^ 1);
                        uint8_t* dst = smem + stage * C::STAGE_BYTES + dst_off;
                        uint64_t* fbar = &full[stage];
                        if (lane == 0) mbar_expect_tx(fbar, TMA_BYTES);
                        if (is_a) tma_load_4d(dst, map, fbar, xb * kBlockK + a_c0, oy + dy, ci0, im);
                        else tma_load_5d(dst, map, fbar, xb * kBlockK, oy, n_tile * BN, im, sx);
                        if (++stage == C::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                        if (++xb == p.wg_xblocks) {
                            xb = 0;
                            if (++oy == p.OH) {
                                oy = 0;
                                ++im;
                            }
                        }
                    }
                    continue;
                }
                const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, img = m_tile / (p.tiles_w * p.tiles_h);
                const int ow0 = tw << p.TW_log2, oh0 = th * p.TH;
                const int iw0 = ow0 * p.stride - p.pad, ih0 = oh0 * p.stride - p.pad;
                const int n0 = n_tile * BN;
                // correlation: chunk n_tile = halo rows [4*n_tile, 4*n_tile + 4) x 32 columns of frame t+tau
                const int bw0 = (ow0 - p.corr_r) * p.stride - p.pad;
                const int bh0 = (oh0 - p.corr_r + 4 * n_tile) * p.stride - p.pad;
                const int k_beg = sg.c0 * kUnit, k_end = min(sg.c1 * kUnit, k_iters);
                const int kcb = p.kc_blocks, fS = p.S, dil = p.dil, stem = p.stem;
                int kc = k_beg % kcb, rs = k_beg / kcb, s = rs % fS, r = rs / fS;
                for (int k = k_beg; k < k_end; ++k) {
                    TRACED_WAIT(0, &empty[stage], phase ^ 1);
                    uint8_t* dst = smem + stage * C::STAGE_BYTES + dst_off;
                    uint64_t* fbar = &full[stage];
                    if (PAIR) {
                        // each CTA stages its own A tile and its half of the weight tile (completion: own barrier)
                        if (lane == 0) mbar_expect_tx(fbar, TMA_BYTES);
                        if (is_a) tma_load_4d(dst, map, fbar, kc * kBlockK + a_c0, iw0 + s * dil, ih0 + r * dil, img);
                        else tma_load_2d(dst, map, fbar, k * kBlockK, n0 + (int)crank * (BN / 2));
                    } else {
                        if (lane == 0) mbar_expect_tx(fbar, TMA_BYTES);
                        if (is_a) {
                            // stem: filter row r = 32 consecutive floats (8 pixels x 4 channels) of padded input row
                            // 2*oh + r starting at padded pixel 2*ow; rows are indexed (pair, parity)
                            // (3xFP16: a K block = TWO filter rows, one per 32-float sub-tile: rows 2r and 2r + 1)
                            if (stem) tma_load_5d(dst, map, fbar, 0, ow0, (F16 ? 2 * r + lane : r) & 1, oh0 + ((F16 ? 2 * r + lane : r) >> 1), img);
                            else tma_load_4d(dst, map, fbar, kc * kBlockK + a_c0, iw0 + s * dil, ih0 + r * dil, img);
                        } else if (CORR) {
                            tma_load_4d(dst, map, fbar, kc * kBlockK + (F16 ? (lane - NA) * kBoxC : 0), bw0, bh0, img);
                        } else {
                            tma_load_2d(dst, map, fbar, k * kBlockK, n0);
                        }
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                    if (++kc == kcb) {
                        kc = 0;
                        if (++s == fS) {
                            s = 0;
                            ++r;
                        }
                    }
                }
            }
            TRACE_FLUSH(0);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // fp32 accumulation inside the tensor core truncates on every add, a bias that grows with the
        // number of accumulation steps (measured: ~2e-8 of the result per step).  So (1) the main
        // hi*hi products are accumulated in TMEM only over a CHUNK of kChunkK k-blocks (32 steps);
        // each finished chunk is drained by the epilogue warps into fp32 REGISTERS with
        // round-to-nearest adds while the next chunk runs into the other TMEM buffer; (2) the two
        // cross terms (2^-11 of the result, so their truncation is negligible) accumulate over the
        // whole tile in their own TMEM buffer.  TMEM: main[2] | cross[2], BN columns each.
        if (lane == 0 && crank == 0) {
            constexpr uint32_t idesc = F16 ? (PAIR ? make_idesc_f16<BN, 256>() : make_idesc_f16<BN>())
                                           : (PAIR ? make_idesc<BN, 256>() : make_idesc<BN>());
            const uint64_t desc0 = make_smem_desc(smem_u32(smem));     // stage s / operand o: + (byte offset >> 4)
            int stage = 0, cbuf = 0, local = 0, issued = 0;
            uint32_t phase = 0, cphase = 0;
            TRACE_DECL;
            TRACE_BEGIN();
            for (; local < sched.nseg; ++local) {
                const Seg sg = sched.get(local);
                const int xacc = local & 1;
                issued += min(sg.c1 * kUnit, k_iters) - sg.c0 * kUnit;
                const uint32_t d_cross = tmem_base + (2 + xacc) * BN;
                if (SPLIT && !ATMEM) {
                    TRACED_WAIT(0, &xempty[xacc], ((local >> 1) & 1) ^ 1);   // epilogue has read this cross buffer
                    tc_fence_after();
                }
                const int k_beg = sg.c0 * kUnit, k_end = min(sg.c1 * kUnit, k_iters);
                for (int k = k_beg; k < k_end; ++k) {
                    const int kin = k % kChunkK;
                    const bool chunk_start = kin == 0 || k == k_beg;         // (a segment may begin inside a chunk)
                    if (chunk_start) {
                        TRACED_WAIT(1, &tempty[cbuf], cphase ^ 1);          // epilogue drained this chunk buffer
                        tc_fence_after();
                    }
                    const uint32_t d_main = tmem_base + cbuf * BN;
                    TRACED_WAIT(2, &full[stage], phase);
                    if (SPLIT || PAIR) TRACED_WAIT(3, &cvt[stage], phase);   // lo tiles written (pair: peer landed too)
                    tc_fence_after();
#ifdef D2T_CONV_TRACE
                    if (CHAIN && p.trace && local == 0 && k == k_beg) p.trace[(size_t)blockIdx.x * 64 + 8 + 5] = clock64();   // role 1 value 5: first MMA
#endif
                    const uint64_t a_hi = desc0 + (uint64_t)(stage * (C::STAGE_BYTES >> 4));
                    if (SPLIT) {
                        const uint64_t a_lo = a_hi + (C::OFF_ALO >> 4);
                        const uint64_t b_hi = a_hi + (C::OFF_BHI >> 4);
                        const uint64_t b_lo = a_hi + (C::OFF_BLO >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {                  // 4 K steps of 32 bytes per 128-byte row
                            const uint64_t o = (uint64_t)(kk * 32 >> 4);
                            if (F16) {
                                // A from tensor memory (8 columns = 16 fp16 per K step); all three products of the K step
                                // go to the chunk accumulator (48 accumulation steps per 256-channel chunk)
                                const uint32_t at_hi = tmem_base + C::A_TMEM_BASE + stage * C::A_TMEM_COLS + kk * 8;
                                if (PAIR) {
                                    umma_f16_ts_pair(d_main, at_hi, b_hi + o, idesc, !(chunk_start && kk == 0));
                                    umma_f16_ts_pair(d_main, at_hi, b_lo + o, idesc, 1);
                                    umma_f16_ts_pair(d_main, at_hi + 32, b_hi + o, idesc, 1);
                                } else {
                                    umma_f16_ts(d_main, at_hi, b_hi + o, idesc, !(chunk_start && kk == 0));
                                    umma_f16_ts(d_main, at_hi, b_lo + o, idesc, 1);
                                    umma_f16_ts(d_main, at_hi + 32, b_hi + o, idesc, 1);
                                }
                            } else if (PAIR) {
                                umma_tf32_pair(d_cross, a_lo + o, b_hi + o, idesc, ((k - k_beg) | kk) != 0);
                                umma_tf32_pair(d_cross, a_hi + o, b_lo + o, idesc, 1);
                                umma_tf32_pair(d_main, a_hi + o, b_hi + o, idesc, !(chunk_start && kk == 0));
                            } else {
                                umma_tf32(d_cross, a_lo + o, b_hi + o, idesc, ((k - k_beg) | kk) != 0);
                                umma_tf32(d_cross, a_hi + o, b_lo + o, idesc, 1);
                                umma_tf32(d_main, a_hi + o, b_hi + o, idesc, !(chunk_start && kk == 0));
                            }
                        }
                    } else {
                        const uint64_t b_hi = a_hi + (C::OFF_BHI >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t o = (uint64_t)(kk * 32 >> 4);
                            if (PAIR) umma_tf32_pair(d_main, a_hi + o, b_hi + o, idesc, !(chunk_start && kk == 0));
                            else umma_tf32(d_main, a_hi + o, b_hi + o, idesc, !(chunk_start && kk == 0));
                        }
                    }
                    // smem slot free once these MMAs retire (in a pair: in BOTH CTAs)
                    if (PAIR) umma_commit_pair(&empty[stage]);
                    else umma_commit(&empty[stage]);
                    if (kin == kChunkK - 1 || k == k_end - 1) {          // chunk complete (covers the cross MMAs too)
                        if (PAIR) umma_commit_pair(&tfull[cbuf]);
                        else umma_commit(&tfull[cbuf]);
                        cbuf ^= 1;
                        if (cbuf == 0) cphase ^= 1;
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if constexpr (CHAIN) {
                // the barriers are re-initialised for the next layer: no commit may still be on its way to `empty`
                for (int s = 0; s < C::STAGES; ++s)
                    if (issued > s) mbar_wait_sleep(&empty[s], (uint32_t)(((issued - s - 1) / C::STAGES) & 1));
            }
            TRACE_FLUSH(1);
        }
    } else if (warp < kEpiWarp0 || warp == kEpiWarp0 + 8 || warp == kEpiWarp0 + 9) {
        // ===================== converters (warps 2-3; 3xFP16: + warps 12-13) =====================
        // lo = x - trunc13(x) for the activation tile of every landed stage (and for the second frame's tile in
        // correlation mode): same swizzled address in the stage's "lo" slot, so the layout needs no thought.
        // In a CTA pair the converters also carry the "my copies have landed" signal to the leader: each CTA's TMA
        // completes on its OWN `full` barrier, and the leader's MMA thread waits for both CTAs' cvt arrivals.
        if constexpr (F16) {
            // 3xFP16: the landed activation tile is two fp32 [128 rows x 32 channels] sub-tiles (128-byte rows,
            // SWIZZLE_128B).  A converter thread owns ONE pixel row -- the row whose TMEM lane its warp may address (warps
            // 2, 3, 12, 13 = lane quarters 2, 3, 0, 1) -- reads its 64 floats (quarter warps touch 8 consecutive rows: every
            // 16-byte access is conflict-free), scales by sa (exact), splits x*sa = hi + lo (hi = top 11 significant bits,
            // lo = the remainder rounded to fp16) and stores the packed pairs into the stage's A slot in tensor memory.
            const int cw = warp < kEpiWarp0 ? warp - 2 : warp - (kEpiWarp0 + 8) + 2;      // converter warp 0..3 (trace slot)
            (void)cw;
            const int m = (warp & 3) * 32 + lane;
            const uint32_t sw = (uint32_t)m & 7u;
            const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::A_TMEM_BASE;
            const float sa = pow2f(act_exp(p.amax_in));
            const float sb = CORR ? pow2f(act_exp(p.amax_b)) : 1.f;
            (void)sb;
            int stage = 0;
            uint32_t phase = 0;
            TRACE_DECL;
            TRACE_BEGIN();
            for (int e = 0; e < sched.nseg; ++e) {
                const Seg sg = sched.get(e);
                const int k_beg = sg.c0 * kUnit, k_end = min(sg.c1 * kUnit, k_iters);
                for (int k = k_beg; k < k_end; ++k) {
                    // (the slot's previous MMAs have retired: the TMA refilled this stage only after their commit)
                    TRACED_WAIT(0, &full[stage], phase);
                    tc_fence_after();
                    const uint8_t* row = smem + stage * C::STAGE_BYTES + m * 128;
                    if constexpr (CORRB) {
                        // the A tile landed as fp16 already: [128 rows x 64 halves] hi, then lo -- move it to tensor memory
#pragma unroll
                        for (int part = 0; part < 2; ++part) {         // 0 = hi, 1 = lo
                            const uint8_t* src = row + part * C::OFF_ALO;
#pragma unroll
                            for (int half = 0; half < 2; ++half) {     // K elements [32 half, 32 half + 32)
                                uint32_t w[16];
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const uint4 v = *reinterpret_cast<const uint4*>(src + (((uint32_t)(half * 4 + c) ^ sw) << 4));
                                    w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
                                }
                                tmem_st16(a_lane + stage * C::A_TMEM_COLS + part * 32 + half * 16, w);
                            }
                        }
                    } else
#pragma unroll
                    for (int half = 0; half < 2; ++half) {             // sub-tile = channels [32 half, 32 half + 32)
                        const uint8_t* src = row + half * (kBlockM * kBoxC * 4);
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 v = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sw) << 4));
                            const float a0 = v.x * sa, a1 = v.y * sa, a2 = v.z * sa, a3 = v.w * sa;
                            const float h0 = __uint_as_float(__float_as_uint(a0) & 0xffffe000u);
                            const float h1 = __uint_as_float(__float_as_uint(a1) & 0xffffe000u);
                            const float h2 = __uint_as_float(__float_as_uint(a2) & 0xffffe000u);
                            const float h3 = __uint_as_float(__float_as_uint(a3) & 0xffffe000u);
                            __half2 t;
                            t = __floats2half2_rn(h0, h1); hi[2 * c] = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2half2_rn(h2, h3); hi[2 * c + 1] = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2half2_rn(a0 - h0, a1 - h1); lo[2 * c] = *reinterpret_cast<uint32_t*>(&t);
                            t = __floats2half2_rn(a2 - h2, a3 - h3); lo[2 * c + 1] = *reinterpret_cast<uint32_t*>(&t);
                        }
                        const uint32_t slot = a_lane + stage * C::A_TMEM_COLS + half * 16;
                        tmem_st16(slot, hi);
                        tmem_st16(slot + 32, lo);
                    }
                    if constexpr (CORR) {
                        // correlation: the B operand is an activation tile too.  Its two fp32 sub-tiles sit in the B hi / B lo
                        // slots; row m's 64 values are split like A's and written back IN PLACE as row m of the fp16 hi tile
                        // (first slot) and of the lo tile (second slot): the same 2 x 128 bytes the thread has just read.
                        uint8_t* b0 = smem + stage * C::STAGE_BYTES + C::OFF_BHI + m * 128;
                        uint8_t* b1 = smem + stage * C::STAGE_BYTES + C::OFF_BLO + m * 128;
                        uint4 bh[8], bl[8];
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const uint8_t* src = half ? b1 : b0;
#pragma unroll
                            for (int c = 0; c < 8; c += 2) {
                                const float4 v0 = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sw) << 4));
                                const float4 v1 = *reinterpret_cast<const float4*>(src + (((uint32_t)(c + 1) ^ sw) << 4));
                                const float a[8] = {v0.x * sb, v0.y * sb, v0.z * sb, v0.w * sb, v1.x * sb, v1.y * sb, v1.z * sb, v1.w * sb};
                                uint32_t wh[4], wl[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const __half2 h = __floats2half2_rn(a[2 * q], a[2 * q + 1]);
                                    const float2 hf = __half22float2(h);
                                    const __half2 l = __floats2half2_rn(a[2 * q] - hf.x, a[2 * q + 1] - hf.y);
                                    wh[q] = *reinterpret_cast<const uint32_t*>(&h);
                                    wl[q] = *reinterpret_cast<const uint32_t*>(&l);
                                }
                                bh[half * 4 + (c >> 1)] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                                bl[half * 4 + (c >> 1)] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            *reinterpret_cast<uint4*>(b0 + (((uint32_t)j ^ sw) << 4)) = bh[j];
                            *reinterpret_cast<uint4*>(b1 + (((uint32_t)j ^ sw) << 4)) = bl[j];
                        }
                        fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    if (PAIR) {
                        // the leader issues the pair's MMAs.  ONE arrival per CTA (128 remote arrivals per K block measured
                        // ~40 % slower than single-CTA mode): the four converter warps meet at a named barrier first
                        asm volatile("bar.sync 4, 128;" ::: "memory");
                        if ((warp & 3) == 2 && warp < kEpiWarp0 && lane == 0) {
                            tc_fence_after();
                            tc_fence_before();
                            mbar_arrive_remote(map_to_cta(smem_u32(&cvt[stage]), 0));
                        }
                    } else {
                        mbar_arrive(&cvt[stage]);
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            TRACE_FLUSH(2 + cw);
        } else if (PASSES == 3 || PAIR) {
            const int ct = threadIdx.x - 64;                       // 0..63
            int stage = 0;
            uint32_t phase = 0;
            for (int e = 0; e < sched.nseg; ++e) {
                const Seg sg = sched.get(e);
                const int k_beg = sg.c0 * kUnit, k_end = min(sg.c1 * kUnit, k_iters);
                for (int k = k_beg; k < k_end; ++k) {
                    mbar_wait_sleep(&full[stage], phase);           // this CTA's copies of the stage have landed
                    uint8_t* st = smem + stage * C::STAGE_BYTES;
                    const float4* ax = reinterpret_cast<const float4*>(st);
                    float4* al = reinterpret_cast<float4*>(st + C::A_BYTES);
#pragma unroll 4
                    for (int i = ct; i < (PASSES == 3 ? C::A_BYTES / 16 : 0); i += kCvtThreads) {
                        const float4 v = ax[i];
                        al[i] = make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u),
                                            v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                                            v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u),
                                            v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
                    }
                    if (CORR && PASSES == 3) {
                        const float4* bx = reinterpret_cast<const float4*>(st + C::OFF_BHI);
                        float4* bl = reinterpret_cast<float4*>(st + C::OFF_BLO);
#pragma unroll 4
                        for (int i = ct; i < C::B_BYTES / 16; i += kCvtThreads) {
                            const float4 v = bx[i];
                            bl[i] = make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u),
                                                v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                                                v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u),
                                                v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
                        }
                    }
                    fence_proxy_async();                           // generic-proxy writes -> visible to the tensor core
                    if (PAIR) mbar_arrive_remote(map_to_cta(smem_u32(&cvt[stage]), 0));
                    else mbar_arrive(&cvt[stage]);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }
    } else {
        // ===================== epilogue (warpgroups 1-2) =====================
        // 3xFP16: these warps hold a 64-column accumulator row AND the prefetched residual row in registers
        if constexpr (F16) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        const int q = (warp - kEpiWarp0) & 3;              // TMEM lane quarter of this warp
        const int grp = (warp - kEpiWarp0) >> 2;           // column half of the tile this warp owns
        constexpr int HN = BN / 2;                         // columns per thread
        const int cofs = grp * HN;
        const int m = q * 32 + lane;                       // tile row = TMEM lane = output pixel in the tile
        const int TW = 1 << p.TW_log2;
        const int hl = m >> p.TW_log2, wl = m & (TW - 1);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        // 3xFP16: the accumulator holds (x * 2^ea) . (w * 2^w_exp); the drain undoes both (exact: fma by a power of two)
        const float descale = F16 ? pow2f(-(act_exp(p.amax_in) + (p.amax_b ? act_exp(p.amax_b) : p.w_exp))) : 1.f;
        int local = 0, cbuf = 0;
        uint32_t cphase = 0, rphase = 0;
        TRACE_DECL;
        TRACE_BEGIN();
        float tmax = 0.f;                                   // max |value| this thread wrote (all its tiles)
        // Folded BatchNorm / bias of the current n tile live in shared memory ([BN] scale | [BN] shift): every thread needs
        // all of its 64 channels' values every tile, and as global loads they paid an L2 round trip per 16 channels
        // (measured: the affine loop was 5-7k cycles per tile).  Thread et fetches ONE value a tile ahead (register),
        // so the latency hides behind the previous tile; channels past Cout read a clamped index and are never stored.
        // 3xFP16 convolutions (forward and backward-data): the per-channel scale lives in the PACKED WEIGHTS (folded by the
        // packers) and the shift INITIALISES the accumulator, so a finished tile needs no affine pass at all -- the traced
        // epilogue spent 3.4 k of its 8.5 k cycles per tile there on the residual 1x1 layers (profiles/r02_conv_trace.txt).
        // Every other variant (TF32 kinds, WGRAD's column scale, CORRB's 1/C) keeps the scale / shift loop below.
        constexpr bool FOLDED = F16 && !WGRAD && !CORRB && !CORR;
        float* scsh = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 384);
        const int et = (warp - kEpiWarp0) * 32 + lane;     // 0..255
        auto fetch_scsh = [&](int tile) -> float {
            if (CORR || et >= 2 * BN) return 0.f;
            const bool is_shift = et >= BN;
            if (FOLDED && !is_shift) return 1.f;
            const float* src = is_shift ? p.shift : p.scale;
            const int chn = min((tile % p.n_tiles) * BN + (is_shift ? et - BN : et), p.Cout - 1);
            return src ? __ldg(src + chn) : (is_shift ? 0.f : 1.f);
        };
        float scsh_next = sched.nseg > 0 ? fetch_scsh(sched.get(0).tile) : 0.f;
        // ARES: the epilogue is software-pipelined over the CTA's tiles.  Output / residual staging is double-buffered (tile i
        // uses buffer i & 1) and the residual of tile i + 1 is requested during tile i, once the buffer's previous store has
        // been read; the shifts sit in THREE small buffers written one tile ahead, so one named barrier per tile orders them.
        auto res_request = [&](int e2) {                     // (thread m == 0 of each group) residual slabs of segment e2
            const int t2 = sched.get(e2).tile;
            const int n02 = (t2 % p.n_tiles) * BN, m2 = t2 / p.n_tiles;
            const int tw2 = m2 % p.tiles_w, th2 = (m2 / p.tiles_w) % p.tiles_h, img2 = m2 / (p.tiles_w * p.tiles_h);
            uint64_t* rb = (e2 & 1) ? &rfull2[grp] : &rfull[grp];
            int nsl = 0;
#pragma unroll
            for (int sl = 0; sl < HN / 32; ++sl) nsl += (n02 + cofs + sl * 32 < p.Cout) ? 1 : 0;
            mbar_expect_tx(rb, nsl * (kBlockM * 128));
#pragma unroll
            for (int sl = 0; sl < HN / 32; ++sl) {
                const int chs = n02 + cofs + sl * 32;
                if (chs < p.Cout)
                    tma_load_4d(out_stage + (e2 & 1) * C::OUT_BUF_BYTES + (grp * C::OUT_SLABS + sl) * (kBlockM * 128), &tmR, rb, chs,
                                tw2 << p.TW_log2, th2 * p.TH, img2);
            }
        };
        if constexpr (ARES) {
            if (sched.nseg > 0) {
                if (p.res != nullptr && m == 0) res_request(0);
                if (et >= BN && et < 2 * BN) scsh[et - BN] = scsh_next;                 // shift buffer 0 <- tile 0
                scsh_next = sched.nseg > 1 ? fetch_scsh(sched.get(1).tile) : 0.f;        // (fetched one more tile ahead)
            }
        }
        for (; local < sched.nseg; ++local) {
            const Seg sg = sched.get(local);
            const float scsh_cur = scsh_next;
            if (local + (ARES ? 2 : 1) < sched.nseg) scsh_next = fetch_scsh(sched.get(local + (ARES ? 2 : 1)).tile);
            uint8_t* const obuf = out_stage + (ARES ? (local & 1) * C::OUT_BUF_BYTES : 0);
            const int t = sg.tile;
            const int seg_k0 = sg.c0 * kUnit, seg_k1 = min(sg.c1 * kUnit, k_iters);
            const int nchunks = (seg_k1 - 1) / kChunkK - seg_k0 / kChunkK + 1;      // accumulation chunks the segment touches
            const int xacc = local & 1;
            const int n_tile = t % p.n_tiles, m_tile = (t / p.n_tiles) * CS + (int)crank;
            const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, img = m_tile / (p.tiles_w * p.tiles_h);
            const int oh = th * p.TH + hl, ow = (tw << p.TW_log2) + wl;
            // WGRAD: tile row m = input channel ci0 + m of filter tap `tap` (tiles_w = tiles_h = 1 there)
            const int wg_rs = WGRAD ? p.R * p.S : 1, wg_tap = m_tile % wg_rs, wg_ci = (m_tile / wg_rs) * kBlockM + m;
            const bool pix_ok = WGRAD ? wg_ci < p.wg_cin
                                      : (oh < p.OH && ow < p.OW && img < p.N);   // (a pair's second tile may not exist)
            const size_t pix = ((size_t)img * p.OH + oh) * p.OW + ow;
            const int n0 = n_tile * BN;
            // EPI2: the residual tile is requested NOW -- before the accumulator is even complete -- as bulk tensor
            // loads straight into this group's output slabs (the epilogue then adds and overwrites in place, each thread
            // its own row): 64 KB per SM in flight on the TMA engine, which plain loads cannot sustain (the LSU keeps
            // ~16 KB in flight per SM: measured ~10 B/clk/SM for register-prefetched residual rows)
            const bool res_tma = EPI2 && !CORR && p.res != nullptr && sg.role != 1;
            if constexpr (EPI2 && !ARES) {
                if (res_tma && m == 0) {
                    tma_store_wait_read();                  // the previous tile's stores have read the slabs
                    int nsl = 0;
#pragma unroll
                    for (int sl = 0; sl < HN / 32; ++sl) nsl += (n0 + cofs + sl * 32 < p.Cout) ? 1 : 0;
                    mbar_expect_tx(&rfull[grp], nsl * (kBlockM * 128));
#pragma unroll
                    for (int sl = 0; sl < HN / 32; ++sl) {
                        const int chs = n0 + cofs + sl * 32;
                        if (chs < p.Cout)
                            tma_load_4d(out_stage + (grp * C::OUT_SLABS + sl) * (kBlockM * 128), &tmR, &rfull[grp], chs,
                                        tw << p.TW_log2, th * p.TH, img);
                    }
                }
            }
            float acc[HN];
            if constexpr (FOLDED) {
                if constexpr (ARES) {
                    // shift buffer (i + 1) % 3 <- tile i + 1 (its last readers, tile i - 2, passed the previous barrier)
                    if (local + 1 < sched.nseg && et >= BN && et < 2 * BN) scsh[((local + 1) % 3) * BN + et - BN] = scsh_cur;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                } else {
                asm volatile("bar.sync 1, 256;" ::: "memory");   // every epilogue thread has read the previous tile's shifts
                if (et >= BN && et < 2 * BN) scsh[et] = scsh_cur;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                if (sg.role != 1) {                              // (a published partial tile must not carry the shift)
                    const float4* sh4 = reinterpret_cast<const float4*>(scsh + (ARES ? (local % 3) * BN : BN) + cofs);
#pragma unroll
                    for (int j = 0; j < HN; j += 4) {
                        const float4 sh = sh4[j >> 2];
                        acc[j] = sh.x; acc[j + 1] = sh.y; acc[j + 2] = sh.z; acc[j + 3] = sh.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < HN; ++j) acc[j] = 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < HN; ++j) acc[j] = 0.f;
            }
            if constexpr (ARES) {
                // one chunk per tile: all four tensor-memory loads of the row in flight before ONE wait (the generic drain
                // below pays a round trip per 16 columns), then value = accumulator * 2^-(ea + w_exp) + shift
                TRACED_WAIT(0, &tfull[cbuf], cphase);
                tc_fence_after();
                uint32_t raw[HN];
#pragma unroll
                for (int c = 0; c < HN / 16; ++c) tmem_ld16_nowait(lane_base + cbuf * BN + cofs + c * 16, raw + c * 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < HN; ++j) acc[j] = fmaf(__uint_as_float(raw[j]), descale, acc[j]);
                tc_fence_before();
                mbar_arrive(&tempty[cbuf]);
                cbuf ^= 1;
                if (cbuf == 0) cphase ^= 1;
            } else
            for (int ch = 0; ch < nchunks; ++ch) {          // drain finished chunks: fp32 round-to-nearest adds
                TRACED_WAIT(0, &tfull[cbuf], cphase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < HN / 16; ++c) {
                    float v[16];
                    tmem_ld16(lane_base + cbuf * BN + cofs + c * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c * 16 + j] = F16 ? fmaf(v[j], descale, acc[c * 16 + j]) : acc[c * 16 + j] + v[j];
                }
                tc_fence_before();
                if (PAIR) mbar_arrive_remote(map_to_cta(smem_u32(&tempty[cbuf]), 0));
                else mbar_arrive(&tempty[cbuf]);
                cbuf ^= 1;
                if (cbuf == 0) cphase ^= 1;
            }
            if (SPLIT && !ATMEM) {                           // + the cross terms of the whole tile
#pragma unroll
                for (int c = 0; c < HN / 16; ++c) {
                    float v[16];
                    tmem_ld16(lane_base + (2 + xacc) * BN + cofs + c * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c * 16 + j] = F16 ? fmaf(v[j], descale, acc[c * 16 + j]) : acc[c * 16 + j] + v[j];
                }
                tc_fence_before();
                if (PAIR) mbar_arrive_remote(map_to_cta(smem_u32(&xempty[xacc]), 0));
                else mbar_arrive(&xempty[xacc]);
            }
            // ---- from here on the tile lives in registers; the tensor core is already on the next segment
            if constexpr (ARES) {
                if (p.res != nullptr && m == 0 && local + 1 < sched.nseg) {
                    tma_store_wait_read();                   // buffer (i + 1) & 1: the store of tile i - 1 (issued a drain ago) has read it
                    res_request(local + 1);
                }
            }
#ifdef D2T_CONV_TRACE
            const long long t_post__ = clock64();
#endif
            if constexpr (WGRAD) {
                if (p.wg_partials != nullptr && sg.role != 0) {
                    // split tile: park the partial sum for wgrad_reduce (no flags: the kernel boundary orders it)
                    const int slot = sg.tile == sched.first_tile ? 0 : 1;
                    float* dst = p.wg_partials + (((size_t)blockIdx.x * 2 + slot) * BN + cofs) * kBlockM + m;
#pragma unroll
                    for (int j = 0; j < HN; ++j) dst[j * kBlockM] = acc[j];
                    continue;
                }
            }
            if (sg.role == 1) {
                // partial tile: publish registers -> scratch[cta][column][row] (coalesced across the warp)
                float* dst = p.sk_scratch + ((size_t)blockIdx.x * BN + cofs) * kBlockM + m;
                if constexpr (CHAIN) {
                    // no grid-wide barrier separates independent layers: the finisher of my previous partial clears the flag
                    // once it has read the slot
                    if (m == 0 && grp == 0) {
                        int v;
                        do {
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.sk_flags + blockIdx.x) : "memory");
                            if (v != 0) __nanosleep(64);
                        } while (v != 0);
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
#pragma unroll
                for (int j = 0; j < HN; ++j) dst[j * kBlockM] = acc[j];
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 epilogue warps
                if (m == 0 && grp == 0) {
                    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + blockIdx.x), "r"(p.sk_epoch) : "memory");
                }
                continue;
            }
            if (sg.role == 2) {
                // finisher: add the partials of the CTAs that ran the earlier chunks of this tile, in k order
                const long long ufirst = (long long)t * sched.cpt;
                const long long U = (long long)tiles * sched.cpt;
                const int c_first = (int)(((ufirst + 1) * n_units + U - 1) / U) - 1;
                for (int cc = c_first; cc < unit_id; ++cc) {
                    const int c = cc * CS + (int)crank;    // the CTA of unit cc that ran my tile position
                    if (m == 0 && grp == 0) {
                        int v;
                        do {
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.sk_flags + c) : "memory");
                            if (v != p.sk_epoch) __nanosleep(64);
                        } while (v != p.sk_epoch);
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    __threadfence();
                    const float* src = p.sk_scratch + ((size_t)c * BN + cofs) * kBlockM + m;
#pragma unroll
                    for (int j = 0; j < HN; ++j) acc[j] += __ldcg(src + j * kBlockM);
                }
                // the consumer clears the flags it used (each publisher has exactly one finisher), so a launch never
                // finds its own epoch left over from an earlier launch with the same argument: a captured CUDA graph
                // replays the same epochs.  The next writer of these flags starts after this grid has completed.
                if (c_first < unit_id) {
                    asm volatile("bar.sync 1, 256;" ::: "memory");        // every partial of this tile has been read
                    if (m == 0 && grp == 0)
                        for (int cc = c_first; cc < unit_id; ++cc) {
                            if constexpr (CHAIN) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + cc), "r"(0) : "memory");
                            else p.sk_flags[cc * CS + (int)crank] = 0;
                        }
                }
            }
            if constexpr (CORR) {
                // tile row m = position (yl, xl) of the 8x16 tile; column n = halo (rl, cl) of this 4x32 chunk: the in-range
                // displacements of position m are the columns cl = ti + r + xl of halo rows rl = tj + r + yl - 4 n_tile.
                // Written straight from the accumulator rows, every store instruction would scatter 32 single floats over 32
                // sectors (lanes = positions, and the channel depends on the lane): measured, the band loop took as long as
                // the whole K loop (profiles/r02_corr_trace.txt).  So each halo row goes through a swizzled [128 x 32] staging
                // tile (the output slabs, idle in this mode) and is written out transposed: NHWC -- lanes = the 2r+1
                // consecutive channels of one pixel (one 68-byte run per instruction); NCHW -- lanes = consecutive x of one
                // channel plane.
                const int D = p.corr_D;
                float* stage = reinterpret_cast<float*>(out_stage) + grp * (kBlockM * 32);       // 16 KB per group
                const int wq = (warp - kEpiWarp0) & 3;                                            // rows wq*32 .. +31 of the tile
                const int tile_y0 = th * p.TH, tile_x0 = tw << p.TW_log2;
                const float nel = p.corr_nelems;
#pragma unroll
                for (int rh = 0; rh < HN / 32; ++rh) {
                    const int rl = grp * (HN / 32) + rh;             // halo row of the chunk (this group's half)
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");      // the previous pass has been read
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4)
                        *reinterpret_cast<float4*>(stage + m * 32 + ((c4 ^ (m & 7)) << 2)) =
                            make_float4(acc[rh * 32 + 4 * c4], acc[rh * 32 + 4 * c4 + 1], acc[rh * 32 + 4 * c4 + 2], acc[rh * 32 + 4 * c4 + 3]);
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                    if (p.out) {
                        // one pixel per instruction: lane = ti + r
                        for (int i = 0; i < 32; ++i) {
                            const int m2 = wq * 32 + i, hl2 = m2 >> p.TW_log2, wl2 = m2 & (TW - 1);
                            const int oh2 = tile_y0 + hl2, ow2 = tile_x0 + wl2, tjr = 4 * n_tile + rl - hl2;
                            if (oh2 >= p.OH || ow2 >= p.OW || img >= p.N || tjr < 0 || tjr >= D) continue;   // (warp-uniform)
                            if (lane < D) {
                                const int col = lane + wl2;
                                const float val = __fdiv_rn(stage[m2 * 32 + (((col >> 2) ^ (m2 & 7)) << 2) + (col & 3)], nel);   // kernel.cu:100
                                tmax = fmaxf(tmax, fabsf(val));
                                p.out[(((size_t)img * p.OH + oh2) * p.OW + ow2) * p.out_cstride + p.out_coffset + tjr * D + lane] = val;
                            }
                        }
                    }
                    if (p.out_nchw) {
                        // one displacement per instruction: lane = position (consecutive x of one channel plane)
                        const int m2 = wq * 32 + lane, hl2 = m2 >> p.TW_log2, wl2 = m2 & (TW - 1);
                        const int oh2 = tile_y0 + hl2, ow2 = tile_x0 + wl2, tjr = 4 * n_tile + rl - hl2;
                        const bool ok = oh2 < p.OH && ow2 < p.OW && img < p.N && tjr >= 0 && tjr < D;
                        float* o = p.out_nchw + (((size_t)img * D * D + (size_t)(ok ? tjr : 0) * D) * p.OH + oh2) * p.OW + ow2;
                        const size_t cs = (size_t)p.OH * p.OW;
                        for (int tir = 0; tir < D; ++tir) {
                            const int col = tir + wl2;
                            const float val = __fdiv_rn(stage[m2 * 32 + (((col >> 2) ^ (m2 & 7)) << 2) + (col & 3)], nel);
                            if (ok) {
                                if (!p.out) tmax = fmaxf(tmax, fabsf(val));
                                o[tir * cs] = val;
                            }
                        }
                    }
                }
            } else {
            if constexpr (!FOLDED) {
            asm volatile("bar.sync 1, 256;" ::: "memory");   // every epilogue thread is done with the previous tile's values
            if (et < 2 * BN) scsh[et] = scsh_cur;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if constexpr (EPI2) {
                if (res_tma) {                               // the residual slabs of this group have landed
                    if constexpr (ARES) {
                        mbar_wait_sleep((local & 1) ? &rfull2[grp] : &rfull[grp], (uint32_t)((local >> 1) & 1));
                    } else {
                        TRACED_WAIT(2, &rfull[grp], rphase);
                        rphase ^= 1;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < HN / 16; ++c) {
                float* v = acc + c * 16;
                const int ch0 = n0 + cofs + c * 16;
                if (ch0 >= p.Cout) continue;                 // (warp-uniform)
                const bool full16 = ch0 + 16 <= p.Cout;
                if constexpr (!FOLDED) {                     // per-channel affine (folded BN / bias) from shared memory
                    const float4* sc4 = reinterpret_cast<const float4*>(scsh + cofs + c * 16);
                    const float4* sh4 = reinterpret_cast<const float4*>(scsh + BN + cofs + c * 16);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 sc = sc4[j >> 2], sh = sh4[j >> 2];
                        v[j] = fmaf(v[j], sc.x, sh.x); v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
                        v[j + 2] = fmaf(v[j + 2], sc.z, sh.z); v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
                    }
                }
                if (res_tma) {
                    // (channels / pixels outside the tensor were zero-filled by the TMA)
                    const uint8_t* rrow = obuf + (grp * C::OUT_SLABS + (c >> 1)) * (kBlockM * 128) + (uint32_t)m * 128u;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(rrow + (((uint32_t)((c & 1) * 4 + (j >> 2)) ^ ((uint32_t)m & 7u)) << 4));
                        v[j] += a.x; v[j + 1] += a.y; v[j + 2] += a.z; v[j + 3] += a.w;
                    }
                } else if (p.res && pix_ok) {
                    const float* rh = p.res + pix * p.res_cstride + ch0;
                    if (full16) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 a = CHAIN ? __ldcg(reinterpret_cast<const float4*>(rh + j)) : __ldg(reinterpret_cast<const float4*>(rh + j));
                            v[j] += a.x; v[j + 1] += a.y; v[j + 2] += a.z; v[j + 3] += a.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ch0 + j < p.Cout) v[j] += CHAIN ? __ldcg(rh + j) : __ldg(rh + j);
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if constexpr (WGRAD) {
                    if (pix_ok) {       // dW[co][ci][tap], co = ch0 + j
                        const size_t cs = (size_t)p.wg_cin * wg_rs;
                        float* o = p.wg_out + ((size_t)ch0 * p.wg_cin + wg_ci) * wg_rs + wg_tap;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (full16 || ch0 + j < p.Cout) o[j * cs] = v[j];
                    }
                    continue;
                }
                if constexpr (MASK) {
                if (p.mask && pix_ok) {                      // backward-data: ReLU mask of the forward activation
                    const float* mr = p.mask + pix * p.mask_cstride + ch0;
                    if (full16) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(mr + j));
                            v[j] = a.x > 0.f ? v[j] : 0.f; v[j + 1] = a.y > 0.f ? v[j + 1] : 0.f;
                            v[j + 2] = a.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = a.w > 0.f ? v[j + 3] : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ch0 + j < p.Cout) v[j] = __ldg(mr + j) > 0.f ? v[j] : 0.f;
                    }
                }
                }
                if (p.amax_out && pix_ok) {
                    if (full16 && p.relu) {                  // (after a ReLU the values are their own magnitudes)
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            tmax = fmaxf(fmaxf(tmax, fmaxf(v[j], v[j + 1])), fmaxf(v[j + 2], v[j + 3]));
                    } else if (full16) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            tmax = fmaxf(fmaxf(tmax, fmaxf(fabsf(v[j]), fabsf(v[j + 1]))), fmaxf(fabsf(v[j + 2]), fabsf(v[j + 3])));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ch0 + j < p.Cout) tmax = fmaxf(tmax, fabsf(v[j]));
                    }
                }
                if (p.out_nchw && pix_ok) {
                    float* o = p.out_nchw + (((size_t)img * p.Cout + ch0) * p.OH + oh) * p.OW + ow;
                    const size_t cs = (size_t)p.OH * p.OW;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (full16 || ch0 + j < p.Cout) o[j * cs] = v[j];     // lanes = consecutive ow: coalesced
                }
            }
#ifdef D2T_CONV_TRACE
            trace_acc__[3] += clock64() - t_post__;          // affine / residual / relu / amax part
#endif
            if (p.out) {
                // NHWC output through shared memory + TMA store: the 128 threads of this group lay their rows
                // (32 channels = 128 B each) into a SWIZZLE_128B slab, then ONE bulk tensor store writes the
                // [TH x TW x 32] box as full lines; pixels / channels outside the tensor are clipped by the TMA.
                // EPI2: every slab of the tile has its own buffer, so the stores of a whole tile are issued back to back
                // and their smem reads are only waited for one tile later (with a residual: before its loads above).
                const uint32_t row_off = (uint32_t)m * 128u, sw = (uint32_t)m & 7u;
                if constexpr (EPI2) {
                    if (!res_tma) {
                        // the slabs' previous store has read them (ARES: that was two tiles ago -- one group may stay pending)
                        if (m == 0) {
                            if constexpr (ARES) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            else tma_store_wait_read();
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                    }
#pragma unroll
                    for (int sl = 0; sl < HN / 32; ++sl) {
                        if (n0 + cofs + sl * 32 >= p.Cout) continue;   // (uniform over the group)
                        uint8_t* slab = obuf + (grp * C::OUT_SLABS + sl) * (kBlockM * 128);
                        const float* v = acc + sl * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(slab + row_off + ((((uint32_t)j >> 2) ^ sw) << 4)) =
                                make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                    if (m == 0) {
#pragma unroll
                        for (int sl = 0; sl < HN / 32; ++sl) {
                            const int chs = n0 + cofs + sl * 32;
                            if (chs < p.Cout)
                                tma_store_4d_nocommit(&tmO, obuf + (grp * C::OUT_SLABS + sl) * (kBlockM * 128), chs,
                                                      tw << p.TW_log2, th * p.TH, img);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
                uint8_t* slab = out_stage + grp * (kBlockM * 128);
#pragma unroll
                for (int sl = 0; sl < HN / 32; ++sl) {
                    const int chs = n0 + cofs + sl * 32;
                    if (chs >= p.Cout) continue;              // (uniform over the group)
                    const float* v = acc + sl * 32;
                    if (m == 0) tma_store_wait_read();        // the previous store has read the slab
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(slab + row_off + ((((uint32_t)j >> 2) ^ sw) << 4)) =
                            make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    fence_proxy_async();
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                    if (m == 0) tma_store_4d(&tmO, slab, chs, tw << p.TW_log2, th * p.TH, img);
                }
                }
            }
            }   // !CORR
#ifdef D2T_CONV_TRACE
            trace_acc__[1] += clock64() - t_post__;
#endif
        }
#ifdef D2T_CONV_TRACE
        if ((warp - kEpiWarp0) % 4 == 0) TRACE_FLUSH(6 + (warp - kEpiWarp0) / 4);
        if (CHAIN && p.trace && warp == kEpiWarp0 && lane == 0) p.trace[(size_t)blockIdx.x * 64 + 48 + 5] = clock64();   // role 6 value 5: epilogue done
#endif
        if (p.amax_out) {
            // running max |x| of the output tensor: ONE atomic per CTA and layer (values are >= 0, so unsigned integer
            // order is float order) -- same-address atomics serialise in the L2 slice and would back up the store path
            const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(tmax));
            uint32_t* red = tmem_slot + 1;                    // spare word next to the TMEM address
            if (warp == kEpiWarp0 && lane == 0) *red = 0u;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (lane == 0) atomicMax(red, wmax);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == kEpiWarp0 && lane == 0 && *red != 0u && !EXP(8)) atomicMax(reinterpret_cast<unsigned int*>(p.amax_out), *red);
        }
    }
    if (warp >= kEpiWarp0 && lane == 0 && ((warp - kEpiWarp0) & 3) == 0) {
        tma_store_wait_all();
        if (CHAIN || p.done_self) {                           // my bulk stores are complete: order them before the count / barrier
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
        }
    }
    tc_fence_before();
}

template <int BN, int PASSES, bool CORR, bool PAIR, bool EPI2, bool WGRAD = false, bool CORRB = false, bool MASK = false,
          bool ARES = false>
__global__ void __launch_bounds__((Cfg<BN, PASSES, PAIR, EPI2, ARES>::THREADS), 1)
conv_igemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB_hi,
           const __grid_constant__ CUtensorMap tmB_lo, const __grid_constant__ CUtensorMap tmO,
           const __grid_constant__ CUtensorMap tmR, const __grid_constant__ ConvArgs p) {
#ifdef D2T_CONV_TRACE
    const long long t_entry__ = clock64();
#endif
    using C = Cfg<BN, PASSES, PAIR, EPI2, ARES>;
    constexpr bool SPLIT = C::SPLIT;
    constexpr int kCvtThreads = C::CVT_THREADS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);      // 1024-B aligned (swizzle atom)
    uint8_t* out_stage = smem + C::OUT_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + C::OUT_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* xempty = tempty + 2;
    uint64_t* cvt = bars + C::BAR_CVT;
    uint64_t* rfull = bars + C::BAR_RFULL;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::BAR_TSLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CS = PAIR ? 2 : 1;
    uint32_t crank = 0;                            // 0 = leader
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB_hi);
        if (SPLIT && !CORR) prefetch_tmap(&tmB_lo);
        if (!CORR && p.out) prefetch_tmap(&tmO);
        if ((EPI2 && p.res) || CORRB) prefetch_tmap(&tmR);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < C::NCVT; ++i)
            mbar_init(&cvt[i], (PAIR && C::F16) ? 2 : CS * kCvtThreads);   // (leader's copy collects both CTAs' converters; 3xFP16 pairs: one elected arrival per CTA)
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], CS * kEpiThreads);       // (leader's copy collects both CTAs' epilogue threads)
            mbar_init(&xempty[i], CS * kEpiThreads);
            mbar_init(&rfull[i], 1);
        }
        if (ARES) {                                        // (see conv_body)
            mbar_init(bars + C::BAR_AFULL, 1);
            mbar_init(bars + C::BAR_AEMPTY, kCvtThreads);
            for (int i = 0; i < 4; ++i) mbar_init(bars + C::BAR_AFREE + i, 1);
            for (int i = 0; i < 2; ++i) mbar_init(bars + C::BAR_RFULL2 + i, 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "n"(C::TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "n"(C::TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                  // the peer's barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int k_iters_ = WGRAD ? p.wg_kiters : (CORRB ? p.TH + 2 * p.corr_r : p.R * p.S * p.kc_blocks);
    const Sched sched(((p.m_tiles + CS - 1) / CS) * p.n_tiles, k_iters_, unit_of(k_iters_, C::CHUNK), (int)blockIdx.x / CS,
                      (int)gridDim.x / CS);
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch)
    // overlapped the tail of the previous layer; from here on we touch its output.
#ifdef D2T_CONV_TRACE
    const long long t_prol__ = clock64();
#endif
    if (p.done_prev) {
        // the previous launch of this chain counts its finished CTAs (see the end of this kernel): poll instead of waiting
        // for the hardware's notion of grid completion; everything older in the stream was complete before that launch
        // could finish (it waited the same way, or with griddepcontrol.wait)
        if (threadIdx.x == 0) {
            int v;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.done_prev) : "memory");
                if (v < p.done_target) __nanosleep(40);
            } while (v < p.done_target);
        }
        __syncthreads();
        asm volatile("fence.proxy.async;" ::: "memory");      // my TMA loads come after the acquire
    } else {
        // (the weight lanes of the producer warp run ahead when the plan allows it: see ConvArgs::early_b)
        constexpr int NA_ = C::F16 ? 2 : 1, NL_ = NA_ + ((SPLIT && !CORR) || (CORR && C::F16) ? 2 : 1);
        const bool runs_ahead = !CORR && !PAIR && !WGRAD && !CORRB && p.early_b && warp == 0 && lane >= NA_ && lane < NL_;
        if (!runs_ahead) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    asm volatile("griddepcontrol.launch_dependents;");
#ifdef D2T_CONV_TRACE
    const long long t_dep__ = clock64();
#endif

    conv_body<BN, PASSES, CORR, PAIR, EPI2, WGRAD, CORRB, MASK, false, ARES>(tmA, tmB_hi, tmB_lo, tmO, tmR, p, smem, bars,
                                                                              tmem_base, crank, 0, &sched);
    __syncthreads();
    if (p.done_self && threadIdx.x == 0) {         // every thread's stores (and the flag resets above) precede the barrier
        __threadfence();
        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p.done_self) : "memory");
    }
    if (PAIR) cluster_sync_all();                  // the peer may still signal my barriers / feed my tensor core
#ifdef D2T_CONV_TRACE
    const long long t_sync__ = clock64();
#endif
    if (warp == 2) {
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
#ifdef D2T_CONV_TRACE
        const long long t_de__ = clock64();
        if (p.trace && lane == 0) p.trace[(size_t)blockIdx.x * 64 + 16 + 7] = t_de__ - t_sync__;   // role 2 value 7: the dealloc itself
        if (p.trace && lane == 0) {     // kernel-level timeline of this CTA: role slot 0 (producer), values 5..7 and role 1 value 5
            long long* d__ = p.trace + (size_t)blockIdx.x * 64;
            d__[5] = t_prol__ - t_entry__;      // entry -> prologue done (barriers, TMEM, descriptors)
            d__[6] = t_dep__ - t_entry__;       // entry -> the previous grid is complete
            d__[7] = t_sync__ - t_entry__;      // entry -> every role done (stores waited for)
            d__[8 + 5] = clock64() - t_entry__; // entry -> TMEM released
        }
#endif
    }
}

// ------------------------------------------------------------------ split-K reduction of the weight-gradient GEMM
// One block per (tile, 8 output channels): thread = tile row (an input channel); the partial tiles of the CTAs that shared
// the tile are added in CTA order (deterministic), scaled by the folded BatchNorm factor and written as OIHW.  The unit ->
// CTA arithmetic is the scheduler's (struct Sched).
__global__ void __launch_bounds__(128)
wgrad_reduce(const ConvArgs p, int BN, int G, int cpt) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");     // (the next layer's operand packers wait for this grid's completion)
    const int cols_per_block = 8;
    const int blocks_per_tile = BN / cols_per_block;
    const int t = (int)blockIdx.x / blocks_per_tile, cb = (int)blockIdx.x % blocks_per_tile;
    const int tiles = p.m_tiles * p.n_tiles;
    const long long U = (long long)tiles * cpt;
    const long long ufirst = (long long)t * cpt, ulast = ufirst + cpt - 1;
    const int c_first = (int)(((ufirst + 1) * G + U - 1) / U) - 1;
    const int c_last = (int)(((ulast + 1) * G + U - 1) / U) - 1;
    if (c_first == c_last) return;                           // the tile was not split: its CTA wrote the gradient itself
    const int m = threadIdx.x;
    const int n_tile = t % p.n_tiles, m_tile = t / p.n_tiles;
    const int rs = p.R * p.S, tap = m_tile % rs, ci = (m_tile / rs) * kBlockM + m;
    float sum[cols_per_block];
#pragma unroll
    for (int j = 0; j < cols_per_block; ++j) sum[j] = 0.f;
    for (int c = c_first; c <= c_last; ++c) {
        const int first_tile_c = (int)((U * c / G) / cpt);
        const int slot = t == first_tile_c ? 0 : 1;
        const float* src = p.wg_partials + (((size_t)c * 2 + slot) * BN + cb * cols_per_block) * kBlockM + m;
#pragma unroll
        for (int j = 0; j < cols_per_block; ++j) sum[j] += __ldcg(src + j * kBlockM);
    }
    if (ci >= p.wg_cin) return;
#pragma unroll
    for (int j = 0; j < cols_per_block; ++j) {
        const int co = n_tile * BN + cb * cols_per_block + j;
        if (co < p.Cout) p.wg_out[((size_t)co * p.wg_cin + ci) * rs + tap] = fmaf(sum[j], p.scale ? __ldg(p.scale + co) : 1.f, 0.f);
    }
}

// ------------------------------------------------------------------ the multi-layer persistent kernel
// Every dependent launch of the layer chain costs ~6 us in which the SMs idle (the next grid's CTAs become
// resident as this grid's exit, then sit in griddepcontrol.wait until the LAST CTA is done and its writes are flushed) on top
// of the prologue (barrier init, TMEM allocation).  conv_chain runs a LIST of 3xFP16 convolution layers in one launch: one
// CTA per SM stays resident, keeps its tensor memory, and walks the list; each layer is the same conv_body, with its tensor
// maps and arguments read from a descriptor array in global memory.  Where layer l reads what an earlier layer of the launch
// wrote, the CTAs meet at a grid-wide barrier (self-resetting counter + generation word: safe under CUDA-graph replay);
// independent neighbours (the four heads on base_feat, a block's downsample branch next to its first 1x1) run back to back.
// Launched cooperatively, so all CTAs are co-resident by construction.
struct alignas(128) ChainLayer {
    CUtensorMap tmA, tmB_hi, tmB_lo, tmO, tmR;
    ConvArgs args;
    int variant;                   // 0: BN = 128, 1: BN = 128 with full-tile output staging (EPI2), 2: BN = 64
    int sync_before;               // grid-wide barrier before this layer
};
constexpr int kChainBarsOff = 224 * 1024;     // the largest variant layout: 3 x 64 KB stages + 32 KB output staging
constexpr int kChainSmem = kChainBarsOff + 1024 /*align slack*/ + 384 /*barriers*/ + 1024 /*shift of the current n tile*/ +
                           2 * 304 /*ConvArgs + variant + sync_before of the current and the next layer*/;
static_assert(sizeof(ConvArgs) + 8 <= 304 && sizeof(ConvArgs) % 4 == 0 && offsetof(ChainLayer, variant) == offsetof(ChainLayer, args) + sizeof(ConvArgs) &&
              offsetof(ChainLayer, sync_before) == offsetof(ChainLayer, variant) + 4 && kChainSmem <= 227 * 1024, "chain shared-memory budget");
static_assert(Cfg<128, 16, false, false>::STAGES * Cfg<128, 16, false, false>::STAGE_BYTES + Cfg<128, 16, false, false>::OUT_STAGE_BYTES <= kChainBarsOff &&
              Cfg<128, 16, false, true>::STAGES * Cfg<128, 16, false, true>::STAGE_BYTES + Cfg<128, 16, false, true>::OUT_STAGE_BYTES <= kChainBarsOff &&
              Cfg<64, 16, false, false>::STAGES * Cfg<64, 16, false, false>::STAGE_BYTES + Cfg<64, 16, false, false>::OUT_STAGE_BYTES <= kChainBarsOff,
              "chain shared-memory layout");
static_assert(3 * Cfg<64, 16, false, false>::STAGES + 8 <= 40, "chain barrier block");

__device__ __forceinline__ void grid_barrier(int* gsync, int nblocks) {      // one thread per CTA
    __threadfence();
    int gen, old;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(gen) : "l"(gsync + 1) : "memory");     // BEFORE arriving
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(gsync) : "memory");
    if (old == nblocks - 1) {
        asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(gsync), "r"(0) : "memory");          // ready for the next barrier
        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(gsync + 1) : "memory");
    } else {
        int g;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(g) : "l"(gsync + 1) : "memory");
            if (g == gen) __nanosleep(32);
        } while (g == gen);
    }
}

__global__ void __launch_bounds__(512, 1) conv_chain(const ChainLayer* __restrict__ layers, const int n_layers, int* gsync) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kChainBarsOff);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 44);          // past every variant's barriers (and its spare words)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot;
    // The arguments of a layer (ConvArgs | variant | sync_before) are read from shared memory -- as kernel parameters they
    // would sit in the constant bank, a register copy spills -- and they are fetched ONE LAYER AHEAD into the other of two
    // slots, so that no global-memory round trip sits between two layers.
    constexpr int kArgWords = (int)(sizeof(ConvArgs) + 8) / 4;
    uint8_t* argbuf = reinterpret_cast<uint8_t*>(bars) + 384 + 1024;
    if (threadIdx.x < kArgWords)
        reinterpret_cast<uint32_t*>(argbuf)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(&layers[0].args) + threadIdx.x);
    __syncthreads();
    int prev_nbars = 0;
    for (int l = 0; l < n_layers; ++l) {
        const ChainLayer* L = layers + l;
        const ConvArgs* sp = reinterpret_cast<const ConvArgs*>(argbuf + (l & 1) * 304);
        const int variant = reinterpret_cast<const int*>(sp + 1)[0], sync_before = reinterpret_cast<const int*>(sp + 1)[1];
        if (l > 0) {
            __syncthreads();                   // every role of this CTA has left layer l - 1: stores complete, fences done
#ifdef D2T_CONV_TRACE
            if (sp->trace && threadIdx.x == 0) sp->trace[(size_t)blockIdx.x * 64 + 6] = clock64();   // role 0 value 6: previous layer left
#endif
            if (sync_before && warp == 0) {
                if (lane == 0) grid_barrier(gsync, (int)gridDim.x);
                __syncwarp();
            }
#ifdef D2T_CONV_TRACE
            if (sp->trace && threadIdx.x == 0) sp->trace[(size_t)blockIdx.x * 64 + 7] = clock64();   // role 0 value 7: barrier passed
#endif
            // (the body's barrier after its pipeline reset releases the other warps)
        }
        if (l + 1 < n_layers && threadIdx.x >= 128 && threadIdx.x < 128 + kArgWords)      // (epilogue warps: idle at a layer's start)
            reinterpret_cast<uint32_t*>(argbuf + ((l + 1) & 1) * 304)[threadIdx.x - 128] =
                __ldg(reinterpret_cast<const uint32_t*>(&L[1].args) + (threadIdx.x - 128));
        if (warp == 3 && lane == 0 && l + 1 < n_layers) {       // the next layer's descriptors: fetched while this one runs
            prefetch_tmap(&L[1].tmA);
            prefetch_tmap(&L[1].tmB_hi);
            prefetch_tmap(&L[1].tmB_lo);
            prefetch_tmap(&L[1].tmO);
            prefetch_tmap(&L[1].tmR);
        }
        const ConvArgs& p = *sp;
        if (variant == 0)
            conv_body<128, 16, false, false, false, false, false, false, true>(L->tmA, L->tmB_hi, L->tmB_lo, L->tmO, L->tmR, p, smem, bars, tmem_base, 0u, prev_nbars);
        else if (variant == 1)
            conv_body<128, 16, false, false, true, false, false, false, true>(L->tmA, L->tmB_hi, L->tmB_lo, L->tmO, L->tmR, p, smem, bars, tmem_base, 0u, prev_nbars);
        else
            conv_body<64, 16, false, false, false, false, false, false, true>(L->tmA, L->tmB_hi, L->tmB_lo, L->tmO, L->tmR, p, smem, bars, tmem_base, 0u, prev_nbars);
        prev_nbars = 3 * (variant == 0 ? Cfg<128, 16, false, false>::STAGES : (variant == 1 ? Cfg<128, 16, false, true>::STAGES : Cfg<64, 16, false, false>::STAGES)) + 8;
    }
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

bool encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box, const cuuint32_t* estr, const char* what, bool f16 = false) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return false;
    }
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
        return false;
    }
    return true;
}

}  // namespace
}  // namespace d2t

using namespace d2t;

namespace d2t {
namespace {
// Per-device DEFAULT stream-K scratch: one fp32 partial tile [128 x 128] and one flag per SM.  Allocated once, at the
// first plan creation on a device (never on the launch path).  Plans that were not given a private scratch
// (d2t_conv_plan_set_scratch; D2TEngine allocates one per engine) share it, and two launches that share it must not
// overlap: d2t_conv_plan_run therefore keeps the launches on the default scratch STREAM-ORDERED -- when such a plan
// arrives on another stream than the previous one, the new stream first waits (event) for the previous stream's work,
// and the whole guard + launch runs under a mutex so that host threads cannot interleave.
struct SkScratch {
    float* partial = nullptr;
    int* flags = nullptr;
    cudaStream_t last_stream = nullptr;
    bool used = false;
    cudaEvent_t ev = nullptr;
};
SkScratch g_sk[64];
std::mutex g_sk_mu;
std::atomic<int> g_sk_epoch{0};

bool sk_scratch(SkScratch* out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        set_error("stream-K scratch: bad device");
        return false;
    }
    std::lock_guard<std::mutex> lock(g_sk_mu);
    if (!g_sk[dev].partial) {
        const size_t n = (size_t)sm_count();
        cudaError_t e = cudaMalloc(&g_sk[dev].partial, n * kBlockM * 128 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&g_sk[dev].flags, n * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(g_sk[dev].flags, 0, n * sizeof(int));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_sk[dev].ev, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            set_error("stream-K scratch: %s", cudaGetErrorString(e));
            g_sk[dev] = SkScratch();
            return false;
        }
    }
    *out = g_sk[dev];
    return true;
}

// (g_sk_mu held) order a launch that uses the device's default scratch after the previous such launch
bool sk_serialize(cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    SkScratch& g = g_sk[dev];
    if (g.used && g.last_stream != stream && g.ev) {
        cudaError_t e = cudaEventRecord(g.ev, g.last_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, g.ev, 0);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            set_error("conv: cannot order the shared stream-K scratch across streams (%s); give the plan a private scratch "
                      "(d2t_conv_plan_set_scratch)", cudaGetErrorString(e));
            return false;
        }
    }
    g.last_stream = stream;
    g.used = true;
    return true;
}

// grid = one CTA per SM, fewer only when the layer has fewer K chunks than SMs
int sk_grid(int tiles, int k_iters, int passes) {
    const int unit = unit_of(k_iters, chunk_of(passes));
    const long long units = (long long)tiles * ((k_iters + unit - 1) / unit);
    return (int)(units < sm_count() ? units : sm_count());
}
}  // namespace
}  // namespace d2t

namespace d2t {
namespace {
// NHWC output [N, OH, OW, cstride], channels [coff, coff + Cout): box = one 32-channel slab of a tile
bool encode_out_map(CUtensorMap* map, float* out, int N, int OH, int OW, int Cout, int cstride, int coff, int TH,
                    int TW) {
    const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    const cuuint64_t str[3] = {(cuuint64_t)cstride * 4, (cuuint64_t)OW * cstride * 4, (cuuint64_t)OH * OW * cstride * 4};
    const cuuint32_t box[4] = {32u, (cuuint32_t)TW, (cuuint32_t)TH, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return encode(map, out + coff, 4, dims, str, box, estr, "out");
}
}  // namespace
}  // namespace d2t

struct d2t_conv_plan {
    alignas(64) CUtensorMap tmA;
    alignas(64) CUtensorMap tmB_hi;
    alignas(64) CUtensorMap tmB_lo;
    alignas(64) CUtensorMap tmO;
    alignas(64) CUtensorMap tmR;
    alignas(64) CUtensorMap tmBh_hi;   // pair mode: boxes of BN / 2 weight rows (each CTA of a pair stages half of the tile)
    alignas(64) CUtensorMap tmBh_lo;
    ConvArgs args;
    int grid_single;                   // grid of the single-CTA kernel (a pair plan falls back to it: d2t_conv_plan_set_mask)
    int BN, passes, grid, corr;
    int epi2;                  // 3xFP16, BN = 128: the full-tile output staging / TMA residual variant
    int pair;                  // run as CTA pairs (tcgen05 cta_group::2)
    int private_scratch;       // 1: the caller supplied the stream-K scratch (no cross-stream guard needed)
    int ares;                  // EPI2 sub-variant: activation tile resident in tensor memory across the n tiles of an m tile
    int wgrad;                 // weight-gradient plan (d2t_wgrad_plan_create)
    int corrb;                 // correlation-backward plan (d2t_corrb_plan_create)
};

template <int BN, int PASSES, bool CORR, bool PAIR, bool EPI2 = false, bool WGRAD = false, bool CORRB = false, bool MASK = false,
          bool ARES = false>
static int launch_conv(const d2t_conv_plan* pl, cudaStream_t stream) {
    using C = Cfg<BN, PASSES, PAIR, EPI2, ARES>;
    static SmemAttrOnce once;
    if (!once.ensure(conv_igemm<BN, PASSES, CORR, PAIR, EPI2, WGRAD, CORRB, MASK, ARES>, C::SMEM_BYTES, "conv smem attr")) return 0;
    ConvArgs args = pl->args;
    args.sk_epoch = ++g_sk_epoch;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PAIR || pl->grid_single <= 0 ? pl->grid : pl->grid_single);
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    // D2T_CONV_PDL=0 (experiments): plain stream order -- the next launch of a chain is not resident (holding an SM at
    // griddepcontrol.wait) while another stream's kernel could use that SM
    static const bool pdl = [] { const char* e = getenv("D2T_CONV_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous kernel's tail
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (PAIR) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    D2T_REQUIRE(PASSES != 16 || args.amax_in, "conv plan: the fp16-split mode needs the input's amax (d2t_conv_plan_set_amax)");
    D2T_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_igemm<BN, PASSES, CORR, PAIR, EPI2, WGRAD, CORRB, MASK, ARES>, pl->tmA,
                                   PAIR ? pl->tmBh_hi : pl->tmB_hi, PAIR ? pl->tmBh_lo : pl->tmB_lo, pl->tmO, pl->tmR, args),
                "conv_igemm launch");
    return 1;
}

// how many CTA pairs of the kernel can be resident at once (the persistent grid must not exceed it)
template <int BN, int PASSES>
static int max_pairs() {
    using C = Cfg<BN, PASSES, true>;
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cached[dev]) return cached[dev];
    auto kern = conv_igemm<BN, PASSES, false, true, false>;
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 64);
        cfg.blockDim = dim3(C::THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    (void)cudaGetLastError();
    if (n <= 0) n = sm_count() / 2 * 3 / 4;       // conservative fallback
    if (n > sm_count() / 2) n = sm_count() / 2;
    cached[dev] = n;
    return n;
}

extern "C" d2t_conv_plan* d2t_conv_plan_create(const d2t_conv_desc* d, const float* in,
                                               const void* w_hi, const void* w_lo, const float* scale,
                                               const float* shift, const float* res, float* out, float* out_nchw) {
    if (!d || !in || !w_hi || d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->R <= 0 ||
        d->S <= 0 || d->stride <= 0 || d->dil <= 0 || d->pad < 0) {
        set_error("d2t_conv_plan_create: bad descriptor");
        return nullptr;
    }
    if (d->Cin % kBoxC != 0 || d->in_cstride % 4 != 0 || d->in_cstride < d->Cin) {
        set_error("d2t_conv_plan_create: Cin must be a multiple of 32 (zero-pad) and in_cstride a multiple of 4");
        return nullptr;
    }
    if (d->passes != 1 && d->passes != 3 && d->passes != 16) {
        set_error("d2t_conv_plan_create: passes must be 1 (TF32), 3 (3xTF32, fp32-accurate) or 16 (3xFP16, fp32-accurate)");
        return nullptr;
    }
    const bool f16 = d->passes == 16;
    const int kblk = kblk_of(d->passes);
    if (f16 && scale) {
        set_error("d2t_conv_plan_create: a 3xFP16 plan takes its per-channel scale folded into the packed weights "
                  "(d2t_conv_pack_weights_f16 / _dev with `scale`), not as an epilogue vector");
        return nullptr;
    }
    if (d->passes != 1 && !w_lo) {
        set_error("d2t_conv_plan_create: 3-pass mode needs the lo half of the packed weights");
        return nullptr;
    }
    if ((out && (d->out_cstride % 4 != 0 || d->out_coffset % 4 != 0 || d->Cout % 4 != 0)) || (!out && !out_nchw)) {
        set_error("d2t_conv_plan_create: need an output (NHWC with Cout and the channel stride/offset multiples of 4, and/or NCHW)");
        return nullptr;
    }
    if (res && ((d->res_cstride > 0 ? d->res_cstride : d->Cout) % 4 != 0 || d->Cout % 4 != 0)) {
        set_error("d2t_conv_plan_create: the residual needs Cout and its channel stride to be multiples of 4");
        return nullptr;
    }
    const int OH = (d->H + 2 * d->pad - d->dil * (d->R - 1) - 1) / d->stride + 1;
    const int OW = (d->W + 2 * d->pad - d->dil * (d->S - 1) - 1) / d->stride + 1;
    if (OH <= 0 || OW <= 0) {
        set_error("d2t_conv_plan_create: empty output");
        return nullptr;
    }
    int twl = 0;
    while ((1 << twl) < OW && twl < 7) ++twl;            // TW = min(128, next pow2 >= OW)
    const int TW = 1 << twl, TH = kBlockM / TW;
    if ((TW - 1) * d->stride + 1 > 256 || (TH - 1) * d->stride + 1 > 256) {
        set_error("d2t_conv_plan_create: stride too large for the TMA box");
        return nullptr;
    }
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(d2t_conv_plan)) != 0) {
        set_error("d2t_conv_plan_create: out of memory");
        return nullptr;
    }
    d2t_conv_plan* pl = new (mem) d2t_conv_plan();
    ConvArgs& a = pl->args;
    a.N = d->N; a.OH = OH; a.OW = OW; a.Cout = d->Cout;
    a.R = d->R; a.S = d->S; a.stride = d->stride; a.pad = d->pad; a.dil = d->dil;
    a.kc_blocks = (d->Cin + kblk - 1) / kblk;         // (3xFP16: a trailing half block reads zeros past Cin)
    a.TW_log2 = twl; a.TH = TH; a.stem = 0;
    a.tiles_w = (OW + TW - 1) / TW; a.tiles_h = (OH + TH - 1) / TH;
    a.m_tiles = d->N * a.tiles_h * a.tiles_w;
    // N tile: 128 unless the layer is narrow
    pl->BN = d->Cout <= 64 ? 64 : 128;
    a.n_tiles = (d->Cout + pl->BN - 1) / pl->BN;
    a.scale = scale; a.shift = shift;
    a.res = res; a.res_cstride = d->res_cstride > 0 ? d->res_cstride : d->Cout;
    a.relu = d->relu;
    a.out = out; a.out_cstride = d->out_cstride; a.out_coffset = d->out_coffset;
    a.out_nchw = out_nchw;
    a.amax_in = nullptr; a.amax_out = nullptr; a.w_exp = d->w_exp; a.trace = nullptr; a.exp = 0; a.done_prev = nullptr; a.done_target = 0; a.done_self = nullptr;
    pl->passes = d->passes; pl->corr = 0;
    // CTA pairs (cta_group::2) are opt-in: measured no faster than single-CTA mode (see the kernel comment)
    // 3xFP16 pairs (plain BN = 128 forward layers): D2T_CONV_PAIR=1 every eligible layer, =2 only the K-loop-bound ones
    // (more than two K chunks per tile and no residual: the layers that do not take the EPI2 variant)
    const int pair_env = getenv("D2T_CONV_PAIR") ? atoi(getenv("D2T_CONV_PAIR")) : 0;
    const bool pair_f16_ok = f16 && pl->BN == 128 && out && d->Cout >= 128 &&
                             (pair_env == 1 || (pair_env == 2 && !res && a.R * a.S * a.kc_blocks > 2 * chunk_of(16)));
    pl->pair = (a.m_tiles >= 2 && ((!f16 && pair_env == 1) || pair_f16_ok)) ? 1 : 0;
    if (pl->pair) {
        const int pair_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
        const int ch = unit_of(a.R * a.S * a.kc_blocks, chunk_of(d->passes));
        const long long units = (long long)pair_tiles * ((a.R * a.S * a.kc_blocks + ch - 1) / ch);
        const int maxp = f16 ? max_pairs<128, 16>()
                             : (d->passes == 3 ? (pl->BN == 64 ? max_pairs<64, 3>() : max_pairs<128, 3>())
                                               : (pl->BN == 64 ? max_pairs<64, 1>() : max_pairs<128, 1>()));
        pl->grid = 2 * (int)(units < maxp ? units : maxp);
    } else {
        pl->grid = sk_grid(a.m_tiles * a.n_tiles, a.R * a.S * a.kc_blocks, d->passes);
    }
    pl->grid_single = sk_grid(a.m_tiles * a.n_tiles, a.R * a.S * a.kc_blocks, d->passes);
    SkScratch sk;
    if (!sk_scratch(&sk)) {
        free(pl);
        return nullptr;
    }
    a.sk_scratch = sk.partial; a.sk_flags = sk.flags; a.sk_epoch = 0;

    // A: NHWC activation, dims innermost first {C, W, H, N}
    const cuuint64_t adims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    const cuuint64_t astr[3] = {(cuuint64_t)d->in_cstride * 4, (cuuint64_t)d->W * d->in_cstride * 4,
                                (cuuint64_t)d->H * d->W * d->in_cstride * 4};
    const cuuint32_t abox[4] = {(cuuint32_t)kBoxC, (cuuint32_t)((TW - 1) * d->stride + 1),
                                (cuuint32_t)((TH - 1) * d->stride + 1), 1u};
    const cuuint32_t aestr[4] = {1u, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1u};
    // B: packed weights [Cout, R*S*cin_pad], cin_pad = kc_blocks * (channels per K block); fp32 or fp16 elements
    const cuuint64_t ktot = (cuuint64_t)d->R * d->S * a.kc_blocks * kblk;
    const cuuint64_t bdims[2] = {ktot, (cuuint64_t)d->Cout};
    const cuuint64_t bstr[1] = {ktot * (f16 ? 2 : 4)};
    const cuuint32_t bbox[2] = {(cuuint32_t)kblk, (cuuint32_t)pl->BN};
    const cuuint32_t bbox_half[2] = {(cuuint32_t)kblk, (cuuint32_t)(pl->BN / 2)};
    const cuuint32_t bestr[2] = {1u, 1u};
    bool ok = encode(&pl->tmA, in, 4, adims, astr, abox, aestr, "A") &&
              encode(&pl->tmB_hi, w_hi, 2, bdims, bstr, bbox, bestr, "B hi", f16);
    if (ok && d->passes != 1) ok = encode(&pl->tmB_lo, w_lo, 2, bdims, bstr, bbox, bestr, "B lo", f16);
    if (ok && d->passes == 1) pl->tmB_lo = pl->tmB_hi;
    pl->tmBh_hi = pl->tmB_hi;
    pl->tmBh_lo = pl->tmB_lo;
    if (ok && pl->pair) {
        ok = encode(&pl->tmBh_hi, w_hi, 2, bdims, bstr, bbox_half, bestr, "B hi (pair)", f16);
        if (ok && d->passes != 1) ok = encode(&pl->tmBh_lo, w_lo, 2, bdims, bstr, bbox_half, bestr, "B lo (pair)", f16);
        if (ok && d->passes == 1) pl->tmBh_lo = pl->tmBh_hi;
    }
    if (ok && out) ok = encode_out_map(&pl->tmO, out, d->N, OH, OW, d->Cout, d->out_cstride, d->out_coffset, TH, TW);
    else if (ok) pl->tmO = pl->tmA;
    // EPI2 variant (3xFP16, BN = 128, NHWC output): layers whose time is the epilogue rather than the K loop -- a residual
    // to add, or at most two K chunks per tile -- trade one pipeline stage for full-tile output staging and take the
    // residual through the TMA.  D2T_CONV_EPI2=0/1 forces the choice (experiments).
    pl->epi2 = (f16 && pl->BN == 128 && out && (res || a.R * a.S * a.kc_blocks <= 2 * chunk_of(16))) ? 1 : 0;
    if (f16 && pl->BN == 128 && out && getenv("D2T_CONV_EPI2")) pl->epi2 = atoi(getenv("D2T_CONV_EPI2")) ? 1 : 0;
    // (a pair plan that later receives a ReLU mask falls back to the single-CTA kernel, d2t_conv_plan_set_mask: EPI2 is
    // decided as if it had never been a pair)
    // A-resident sub-variant: 1x1 layers of at most 4 K blocks with several n tiles per m tile (the residual 1x1 convs that
    // close a bottleneck in layer1 / layer2 / layer3, the shortcut convs of those stages).  D2T_CONV_ARES=0 turns it off (experiments).
    pl->ares = (pl->epi2 && !pl->pair && a.R == 1 && a.S == 1 && a.kc_blocks >= 1 && a.kc_blocks <= 4 && a.n_tiles >= 2 &&
                !(getenv("D2T_CONV_ARES") && atoi(getenv("D2T_CONV_ARES")) == 0)) ? 1 : 0;
    pl->tmR = pl->tmA;
    if (ok && pl->epi2 && res)
        ok = encode_out_map(&pl->tmR, const_cast<float*>(res), d->N, OH, OW, d->Cout, a.res_cstride, 0, TH, TW);
    if (!ok) {
        free(pl);
        return nullptr;
    }
    return pl;
}

// The 7x7 stride-2 pad-3 stem (faster_rcnn/resnet.py:116) has 3 input channels: far too thin for
// a 32-channel K block.  Instead the image is stored NHWC with 4 channels and an explicit 3-pixel
// zero border, so that the 7 taps of one filter ROW are 28 consecutive floats; a K block is then
// "filter row r" = 32 consecutive floats (the 8th pixel meets zero weights).  Consecutive output
// pixels start 2 pixels = 32 bytes apart, which a rank-5 tensor map {32, OW, 2, Hp/2, N} with
// strides {32 B, row, 2 rows, image} expresses directly (overlapping windows): still one TMA copy
// per (tile, filter row), same kernel, 7 K blocks.
extern "C" d2t_conv_plan* d2t_conv_stem_plan_create(int N, int H, int W, int Cout, int passes, const float* in,
                                                    const float* w_hi, const float* w_lo, const float* scale,
                                                    const float* shift, int relu, float* out, int out_cstride) {
    if (N <= 0 || H < 7 || W < 7 || Cout <= 0 || !in || !w_hi || !out || out_cstride % 4 != 0 ||
        (passes != 1 && passes != 3 && passes != 16) || (passes != 1 && !w_lo) || (passes == 16 && (scale || Cout > 64))) {
        set_error("d2t_conv_stem_plan_create: bad arguments (passes = 16: fp16 (hi, lo) weights [Cout][4][64] with the scale "
                  "folded in, Cout <= 64, then d2t_conv_plan_set_amax + d2t_conv_plan_set_weight_amax)");
        return nullptr;
    }
    const bool f16 = passes == 16;
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const int Hp = (H + 7) & ~1, Wp = W + 8;                  // padded buffer geometry (d2t_stem_pack_input)
    int twl = 0;
    while ((1 << twl) < OW && twl < 7) ++twl;
    const int TW = 1 << twl, TH = kBlockM / TW;
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(d2t_conv_plan)) != 0) {
        set_error("d2t_conv_stem_plan_create: out of memory");
        return nullptr;
    }
    d2t_conv_plan* pl = new (mem) d2t_conv_plan();
    ConvArgs& a = pl->args;
    a.N = N; a.OH = OH; a.OW = OW; a.Cout = Cout;
    a.R = f16 ? 4 : 7; a.S = 1; a.stride = 2; a.pad = 3; a.dil = 1; a.kc_blocks = 1;   // 3xFP16: K blocks of two filter rows
    a.TW_log2 = twl; a.TH = TH; a.stem = 1;
    a.tiles_w = (OW + TW - 1) / TW; a.tiles_h = (OH + TH - 1) / TH;
    a.m_tiles = N * a.tiles_h * a.tiles_w;
    pl->BN = Cout <= 64 ? 64 : 128;
    a.n_tiles = (Cout + pl->BN - 1) / pl->BN;
    a.scale = scale; a.shift = shift; a.res = nullptr; a.res_cstride = Cout; a.relu = relu;
    a.out = out; a.out_cstride = out_cstride; a.out_coffset = 0; a.out_nchw = nullptr;
    pl->passes = passes; pl->corr = 0; pl->pair = 0;
    a.amax_in = nullptr; a.amax_out = nullptr; a.w_exp = 0; a.trace = nullptr; a.exp = 0; a.done_prev = nullptr; a.done_target = 0; a.done_self = nullptr;
    pl->grid = sk_grid(a.m_tiles * a.n_tiles, a.R, passes);
    {
        SkScratch sk;
        if (!sk_scratch(&sk)) {
            free(pl);
            return nullptr;
        }
        a.sk_scratch = sk.partial; a.sk_flags = sk.flags; a.sk_epoch = 0;
    }
    const cuuint64_t row = (cuuint64_t)Wp * 16;
    const cuuint64_t adims[5] = {32, (cuuint64_t)OW, 2, (cuuint64_t)(Hp / 2), (cuuint64_t)N};
    const cuuint64_t astr[4] = {32, row, 2 * row, (cuuint64_t)Hp * row};
    const cuuint32_t abox[5] = {32u, (cuuint32_t)TW, 1u, (cuuint32_t)TH, 1u};
    const cuuint32_t aestr[5] = {1u, 1u, 1u, 1u, 1u};
    const cuuint64_t bdims[2] = {(cuuint64_t)(f16 ? 4 * 64 : 7 * 32), (cuuint64_t)Cout};
    const cuuint64_t bstr[1] = {(cuuint64_t)(f16 ? 4 * 64 * 2 : 7 * 32 * 4)};
    const cuuint32_t bbox[2] = {f16 ? 64u : 32u, (cuuint32_t)pl->BN};
    const cuuint32_t bestr[2] = {1u, 1u};
    bool ok = encode(&pl->tmA, in, 5, adims, astr, abox, aestr, "stem A") &&
              encode(&pl->tmB_hi, w_hi, 2, bdims, bstr, bbox, bestr, "stem B hi", f16);
    if (ok && passes != 1) ok = encode(&pl->tmB_lo, w_lo, 2, bdims, bstr, bbox, bestr, "stem B lo", f16);
    if (ok && passes == 1) pl->tmB_lo = pl->tmB_hi;
    if (ok) ok = encode_out_map(&pl->tmO, out, N, OH, OW, Cout, out_cstride, 0, TH, TW);
    if (!ok) {
        free(pl);
        return nullptr;
    }
    return pl;
}

extern "C" d2t_conv_plan* d2t_corr_plan_create(int N, int C, int c_real, int H, int W, int in_cstride, int pad, int md, int stride,
                                               int passes, const float* in1, const float* in2, float* out,
                                               int out_cstride, int out_coffset, float* out_nchw) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || pad < 0 || md < 0 || stride <= 0 || !in1 || !in2 ||
        (passes != 1 && passes != 3 && passes != 16) || (!out && !out_nchw)) {
        set_error("d2t_corr_plan_create: bad arguments");
        return nullptr;
    }
    const int r = md / stride;
    if (C % kBoxC != 0 || in_cstride % 4 != 0 || in_cstride < C || r > 8 || r < 1) {
        set_error("d2t_corr_plan_create: needs C %% 32 == 0, 16-byte pixel stride and 1 <= max_displacement/stride2 <= 8");
        return nullptr;
    }
    const int nh = H + 2 * pad - 2 * md, nw = W + 2 * pad - 2 * md;   // correlation_cuda.c:33-34, kernel_size 1
    if (nh <= 0 || nw <= 0) {
        set_error("d2t_corr_plan_create: empty output");
        return nullptr;
    }
    const int OH = (nh + stride - 1) / stride, OW = (nw + stride - 1) / stride;
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(d2t_conv_plan)) != 0) {
        set_error("d2t_corr_plan_create: out of memory");
        return nullptr;
    }
    d2t_conv_plan* pl = new (mem) d2t_conv_plan();
    ConvArgs& a = pl->args;
    a.N = N; a.OH = OH; a.OW = OW; a.Cout = (2 * r + 1) * (2 * r + 1);
    a.R = 1; a.S = 1; a.stride = stride; a.pad = pad - md; a.dil = 1;       // element = lattice*stride + (md - pad)
    a.kc_blocks = (C + kblk_of(passes) - 1) / kblk_of(passes);   // (3xFP16: a trailing half block reads zeros past C)
    a.TW_log2 = 4; a.TH = 8; a.stem = 0;
    a.tiles_w = (OW + 15) / 16; a.tiles_h = (OH + 7) / 8;
    a.m_tiles = N * a.tiles_h * a.tiles_w;
    a.n_tiles = (8 + 2 * r + 3) / 4;                                         // halo chunks of 4 rows
    a.scale = nullptr; a.shift = nullptr; a.res = nullptr; a.res_cstride = 0; a.relu = 0;
    a.out = out; a.out_cstride = out_cstride; a.out_coffset = out_coffset; a.out_nchw = out_nchw;
    a.corr_r = r; a.corr_D = 2 * r + 1; a.corr_nelems = (float)(c_real > 0 ? c_real : C);
    pl->BN = 128; pl->passes = passes; pl->corr = 1; pl->pair = 0;
    a.amax_in = nullptr; a.amax_out = nullptr; a.w_exp = 0; a.trace = nullptr; a.exp = 0; a.done_prev = nullptr; a.done_target = 0; a.done_self = nullptr;
    pl->grid = sk_grid(a.m_tiles * a.n_tiles, a.kc_blocks, passes);
    {
        SkScratch sk;
        if (!sk_scratch(&sk)) {
            free(pl);
            return nullptr;
        }
        a.sk_scratch = sk.partial; a.sk_flags = sk.flags; a.sk_epoch = 0;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t str[3] = {(cuuint64_t)in_cstride * 4, (cuuint64_t)W * in_cstride * 4, (cuuint64_t)H * W * in_cstride * 4};
    const cuuint32_t abox[4] = {(cuuint32_t)kBoxC, (cuuint32_t)(15 * stride + 1), (cuuint32_t)(7 * stride + 1), 1u};
    const cuuint32_t bbox[4] = {(cuuint32_t)kBoxC, (cuuint32_t)(31 * stride + 1), (cuuint32_t)(3 * stride + 1), 1u};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
    bool ok = encode(&pl->tmA, in1, 4, dims, str, abox, estr, "corr A") &&
              encode(&pl->tmB_hi, in2, 4, dims, str, bbox, estr, "corr B");
    pl->tmB_lo = pl->tmB_hi;   // unused: both lo tiles are derived in shared memory
    pl->tmO = pl->tmA;         // unused in correlation mode (direct stores)
    if (!ok) {
        free(pl);
        return nullptr;
    }
    return pl;
}


// Weight gradient of a stride-1 convolution (see the WGRAD kernel comment).  xt: planar fp32 [N][Cin][xh][xt_pitch] (the
// forward input; for a strided 1x1 conv the caller packs the sub-sampled positions, d2t_wgrad_pack_input); g_hi / g_lo:
// planar fp16 [S][N][Cout][OH][g_pitch] = the (hi, lo) split of G * 2^k, k derived from *amax_g, copy s shifted right by
// s*dil - pad columns (d2t_wgrad_pack_grad);
// dw: OIHW fp32 [Cout][Cin][R][S], every element written (no accumulation).  Pitches in elements: xt_pitch % 4 == 0,
// g_pitch % 8 == 0.  amax_x / amax_g: the device scalars holding max |X| and max |G|.
extern "C" d2t_conv_plan* d2t_wgrad_plan_create(int N, int Cin, int Cout, int xh, int xw, int xt_pitch, int OH, int OW,
                                                int g_pitch, int R, int S, int pad, int dil, const float* xt,
                                                const void* g_hi, const void* g_lo, const float* amax_x,
                                                const float* amax_g, const float* scale, float* dw) {
    if (N <= 0 || Cin <= 0 || Cout <= 0 || xh <= 0 || xw <= 0 || OH <= 0 || OW <= 0 || R <= 0 || S <= 0 || pad < 0 ||
        dil <= 0 || !xt || !g_hi || !g_lo || !amax_x || !amax_g || !dw || xt_pitch % 4 != 0 || xt_pitch < xw ||
        g_pitch % 8 != 0 || g_pitch < OW) {
        set_error("d2t_wgrad_plan_create: bad arguments");
        return nullptr;
    }
    if (OH != xh + 2 * pad - dil * (R - 1) || OW != xw + 2 * pad - dil * (S - 1) || g_pitch < xw) {
        set_error("d2t_wgrad_plan_create: output geometry does not match a stride-1 convolution of the input");
        return nullptr;
    }
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(d2t_conv_plan)) != 0) {
        set_error("d2t_wgrad_plan_create: out of memory");
        return nullptr;
    }
    d2t_conv_plan* pl = new (mem) d2t_conv_plan();
    ConvArgs& a = pl->args;
    constexpr int KB = kblk_of(16);                          // 64 pixels per K block
    a.N = N; a.OH = OH; a.OW = OW; a.Cout = Cout;
    a.R = R; a.S = S; a.stride = 1; a.pad = pad; a.dil = dil;
    a.kc_blocks = 1; a.TW_log2 = 7; a.TH = 1; a.stem = 0;
    a.tiles_w = 1; a.tiles_h = 1;
    a.m_tiles = ((Cin + kBlockM - 1) / kBlockM) * R * S;
    pl->BN = Cout <= 64 ? 64 : 128;
    a.n_tiles = (Cout + pl->BN - 1) / pl->BN;
    a.scale = scale; a.shift = nullptr; a.res = nullptr; a.res_cstride = 0; a.relu = 0;
    a.out = nullptr; a.out_nchw = nullptr; a.out_cstride = 0; a.out_coffset = 0;
    a.amax_in = amax_x; a.amax_b = amax_g; a.amax_out = nullptr; a.w_exp = 0;
    a.wg_xblocks = (xw + KB - 1) / KB;                       // K runs over the columns of X
    a.wg_kiters = N * OH * a.wg_xblocks;
    a.wg_cin = Cin; a.wg_out = dw;
    pl->passes = 16; pl->corr = 0; pl->pair = 0; pl->epi2 = 0; pl->wgrad = 1;
    pl->grid = sk_grid(a.m_tiles * a.n_tiles, a.wg_kiters, 16);
    SkScratch sk;
    if (!sk_scratch(&sk)) {
        free(pl);
        return nullptr;
    }
    a.sk_scratch = sk.partial; a.sk_flags = sk.flags; a.sk_epoch = 0;
    // A: planes of X, dims innermost first {x, y, channel, image}; box = 32 pixels of one row x 128 channels
    const cuuint64_t adims[4] = {(cuuint64_t)xw, (cuuint64_t)xh, (cuuint64_t)Cin, (cuuint64_t)N};
    const cuuint64_t astr[3] = {(cuuint64_t)xt_pitch * 4, (cuuint64_t)xh * xt_pitch * 4, (cuuint64_t)Cin * xh * xt_pitch * 4};
    const cuuint32_t abox[4] = {(cuuint32_t)kBoxC, 1u, (cuuint32_t)kBlockM, 1u};
    // B: the S column-shifted copies of the planes of G (fp16), indexed by X's columns; box = 64 columns of one row x BN channels
    const cuuint64_t bdims[5] = {(cuuint64_t)xw, (cuuint64_t)OH, (cuuint64_t)Cout, (cuuint64_t)N, (cuuint64_t)S};
    const cuuint64_t bstr[4] = {(cuuint64_t)g_pitch * 2, (cuuint64_t)OH * g_pitch * 2, (cuuint64_t)Cout * OH * g_pitch * 2,
                                (cuuint64_t)N * Cout * OH * g_pitch * 2};
    const cuuint32_t bbox[5] = {(cuuint32_t)KB, 1u, (cuuint32_t)pl->BN, 1u, 1u};
    const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    bool ok = encode(&pl->tmA, xt, 4, adims, astr, abox, estr, "wgrad A") &&
              encode(&pl->tmB_hi, g_hi, 5, bdims, bstr, bbox, estr, "wgrad B hi", true) &&
              encode(&pl->tmB_lo, g_lo, 5, bdims, bstr, bbox, estr, "wgrad B lo", true);
    pl->tmO = pl->tmA;
    pl->tmR = pl->tmA;
    if (!ok) {
        free(pl);
        return nullptr;
    }
    return pl;
}


// Correlation backward (see the CORRB kernel comment): out[n, y, x, c] (+= nothing: every element written) =
// scale[c] * sum over the (2r+1)^2 window of band * other.  e_hi / e_lo: the expanded band [N][H][W][D][64] fp16
// (d2t_corrb_pack_band); o_hi / o_lo: the other frame's planes [N][C][H][o_pitch] fp16 (d2t_corrb_pack_other);
// amax_e / amax_o: the device scalars the two packers scaled with.
extern "C" d2t_conv_plan* d2t_corrb_plan_create(int N, int C, int H, int W, int r, const void* e_hi, const void* e_lo,
                                                const void* o_hi, const void* o_lo, int o_pitch, const float* amax_e,
                                                const float* amax_o, const float* scale, float* out, int out_cstride,
                                                int out_coffset) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || r < 1 || r > 8 || !e_hi || !e_lo || !o_hi || !o_lo || !amax_e || !amax_o ||
        !out || o_pitch % 8 != 0 || o_pitch < W || out_cstride % 4 != 0 || out_coffset % 4 != 0 || C % 4 != 0) {
        set_error("d2t_corrb_plan_create: bad arguments");
        return nullptr;
    }
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(d2t_conv_plan)) != 0) {
        set_error("d2t_corrb_plan_create: out of memory");
        return nullptr;
    }
    d2t_conv_plan* pl = new (mem) d2t_conv_plan();
    ConvArgs& a = pl->args;
    const int D = 2 * r + 1, TH = 4, TW = 32;
    a.N = N; a.OH = H; a.OW = W; a.Cout = C;
    a.R = 1; a.S = 1; a.stride = 1; a.pad = 0; a.dil = 1; a.kc_blocks = 1;
    a.TW_log2 = 5; a.TH = TH; a.stem = 0;
    a.tiles_w = (W + TW - 1) / TW; a.tiles_h = (H + TH - 1) / TH;
    a.m_tiles = N * a.tiles_h * a.tiles_w;
    pl->BN = 128;
    a.n_tiles = (C + 127) / 128;
    a.scale = scale; a.shift = nullptr; a.res = nullptr; a.res_cstride = C; a.relu = 0;
    a.out = out; a.out_cstride = out_cstride; a.out_coffset = out_coffset; a.out_nchw = nullptr;
    a.corr_r = r; a.corr_D = D;
    a.amax_in = amax_e; a.amax_b = amax_o; a.amax_out = nullptr; a.w_exp = 0;
    pl->passes = 16; pl->corrb = 1;
    pl->grid = sk_grid(a.m_tiles * a.n_tiles, TH + 2 * r, 16);
    SkScratch sk;
    if (!sk_scratch(&sk)) {
        free(pl);
        return nullptr;
    }
    a.sk_scratch = sk.partial; a.sk_flags = sk.flags; a.sk_epoch = 0;
    // A: expanded band, dims innermost first {halo column, displacement row, x, y, image}; box = 32 pixels of one tile row
    const cuuint64_t edims[5] = {64, (cuuint64_t)D, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t estrd[4] = {128, (cuuint64_t)D * 128, (cuuint64_t)W * D * 128, (cuuint64_t)H * W * D * 128};
    const cuuint32_t ebox[5] = {64u, 1u, (cuuint32_t)TW, 1u, 1u};
    const cuuint32_t one5[5] = {1u, 1u, 1u, 1u, 1u};
    // B: planes of the other frame, box = 64 halo columns of one halo row x 128 channels
    const cuuint64_t odims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    const cuuint64_t ostrd[3] = {(cuuint64_t)o_pitch * 2, (cuuint64_t)H * o_pitch * 2, (cuuint64_t)C * H * o_pitch * 2};
    const cuuint32_t obox[4] = {64u, 1u, 128u, 1u};
    bool ok = encode(&pl->tmA, e_hi, 5, edims, estrd, ebox, one5, "corrb band hi", true) &&
              encode(&pl->tmR, e_lo, 5, edims, estrd, ebox, one5, "corrb band lo", true) &&
              encode(&pl->tmB_hi, o_hi, 4, odims, ostrd, obox, one5, "corrb other hi", true) &&
              encode(&pl->tmB_lo, o_lo, 4, odims, ostrd, obox, one5, "corrb other lo", true) &&
              encode_out_map(&pl->tmO, out, N, H, W, C, out_cstride, out_coffset, TH, TW);
    if (!ok) {
        free(pl);
        return nullptr;
    }
    return pl;
}

extern "C" size_t d2t_wgrad_partials_bytes(void) {
    return (size_t)sm_count() * 2 * 128 * kBlockM * sizeof(float);
}

extern "C" int d2t_wgrad_plan_set_partials(d2t_conv_plan* pl, void* partials, size_t bytes) {
    D2T_REQUIRE(pl && pl->wgrad && (!partials || (((uintptr_t)partials & 15) == 0 && bytes >= d2t_wgrad_partials_bytes())),
                "d2t_wgrad_plan_set_partials: needs a weight-gradient plan and a 16-byte aligned buffer of d2t_wgrad_partials_bytes()");
    pl->args.wg_partials = reinterpret_cast<float*>(partials);
    return 1;
}

extern "C" int d2t_conv_plan_set_mask(d2t_conv_plan* pl, const float* mask, int mask_cstride) {
    D2T_REQUIRE(pl && pl->passes == 16 && !pl->corr && !pl->wgrad && !pl->corrb && (!mask || (mask_cstride % 4 == 0 && mask_cstride >= pl->args.Cout)),
                "d2t_conv_plan_set_mask: needs a convolution plan and a mask with a channel stride that is a multiple of 4");
    pl->args.mask = mask;
    pl->args.mask_cstride = mask_cstride;
    if (mask) pl->pair = 0;             // the backward-data epilogue is a single-CTA variant
    return 1;
}

extern "C" int d2t_conv_plan_set_weight_amax(d2t_conv_plan* pl, const float* amax_w) {
    D2T_REQUIRE(pl && pl->passes == 16, "d2t_conv_plan_set_weight_amax: needs a 3xFP16 plan (convolution: max |packed weights|; "
                                        "correlation: max |second input|)");
    pl->args.amax_b = amax_w;
    return 1;
}

extern "C" int d2t_conv_plan_set_early_weights(d2t_conv_plan* pl, int on) {
    D2T_REQUIRE(pl && !pl->corr && !pl->wgrad && !pl->corrb,
                "d2t_conv_plan_set_early_weights: needs a convolution plan (its B operand must be the packed weights)");
    pl->args.early_b = on ? 1 : 0;
    return 1;
}

extern "C" size_t d2t_conv_scratch_bytes(void) {
    return (size_t)sm_count() * kBlockM * 128 * sizeof(float) + (size_t)sm_count() * sizeof(int);
}

extern "C" int d2t_conv_plan_set_scratch(d2t_conv_plan* pl, void* scratch, size_t bytes) {
    D2T_REQUIRE(pl && scratch && ((uintptr_t)scratch & 15) == 0 && bytes >= d2t_conv_scratch_bytes(),
                "d2t_conv_plan_set_scratch: need a 16-byte aligned, zero-initialised buffer of d2t_conv_scratch_bytes()");
    pl->args.sk_scratch = reinterpret_cast<float*>(scratch);
    pl->args.sk_flags = reinterpret_cast<int*>(reinterpret_cast<char*>(scratch) + (size_t)sm_count() * kBlockM * 128 * sizeof(float));
    pl->private_scratch = 1;
    return 1;
}

extern "C" int d2t_conv_plan_set_amax(d2t_conv_plan* pl, const float* amax_in, float* amax_out) {
    D2T_REQUIRE(pl, "d2t_conv_plan_set_amax: null plan");
    pl->args.amax_in = amax_in;
    pl->args.amax_out = amax_out;
    return 1;
}

extern "C" int d2t_conv_plan_set_done(d2t_conv_plan* pl, const d2t_conv_plan* prev, const int* prev_counter,
                                      int* self_counter) {
    D2T_REQUIRE(pl && (!prev_counter || prev), "d2t_conv_plan_set_done: a counter to wait on needs the plan that fills it");
    pl->args.done_prev = prev_counter;
    pl->args.done_target = prev_counter ? prev->grid : 0;
    pl->args.done_self = self_counter;
    return 1;
}

#ifdef D2T_CONV_TRACE
// debug builds only: [grid][8 roles][8] cycle counters (scripts/conv_trace.py)
extern "C" __attribute__((visibility("default"))) int d2t_conv_plan_set_trace(d2t_conv_plan* pl, long long* trace) {
    pl->args.trace = trace;
    pl->args.exp = getenv("D2T_CONV_EXP") ? atoi(getenv("D2T_CONV_EXP")) : 0;
    return pl->grid;
}
#endif

extern "C" void d2t_conv_plan_destroy(d2t_conv_plan* pl) {
    if (pl) free(pl);
}

extern "C" int d2t_conv_plan_info(const d2t_conv_plan* pl, int* out8) {
    D2T_REQUIRE(pl && out8, "d2t_conv_plan_info: null");
    out8[0] = pl->args.OH; out8[1] = pl->args.OW; out8[2] = pl->args.TH; out8[3] = 1 << pl->args.TW_log2;
    out8[4] = pl->BN; out8[5] = pl->args.m_tiles; out8[6] = pl->args.n_tiles; out8[7] = pl->grid * 10 + pl->pair;
    return 1;
}

static int conv_plan_dispatch(const d2t_conv_plan* pl, cudaStream_t stream);

extern "C" int d2t_conv_plan_run(const d2t_conv_plan* pl, cudaStream_t stream) {
    D2T_REQUIRE(pl, "d2t_conv_plan_run: null plan");
    if (pl->private_scratch) return conv_plan_dispatch(pl, stream);
    std::lock_guard<std::mutex> lock(g_sk_mu);       // default scratch: launches stay stream-ordered (see SkScratch)
    if (!sk_serialize(stream)) return 0;
    return conv_plan_dispatch(pl, stream);
}

static int conv_plan_dispatch(const d2t_conv_plan* pl, cudaStream_t stream) {
    if (pl->wgrad) {
        const int ok = pl->BN == 64 ? launch_conv<64, 16, false, false, false, true>(pl, stream)
                                    : launch_conv<128, 16, false, false, false, true>(pl, stream);
        if (!ok || !pl->args.wg_partials) return ok;
        const int unit = unit_of(pl->args.wg_kiters, chunk_of(16));
        const int cpt = (pl->args.wg_kiters + unit - 1) / unit;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pl->args.m_tiles * pl->args.n_tiles * (pl->BN / 8));
        cfg.blockDim = dim3(128);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        D2T_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_reduce, pl->args, pl->BN, pl->grid, cpt), "wgrad_reduce launch");
        return 1;
    }
    if (pl->corrb) return launch_conv<128, 16, false, false, false, false, true>(pl, stream);
    if (pl->corr) {
        if (pl->passes == 16) {
            D2T_REQUIRE(pl->args.amax_b, "correlation plan: the fp16-split mode needs both inputs' amax (d2t_conv_plan_set_weight_amax for the second)");
            return launch_conv<128, 16, true, false>(pl, stream);
        }
        return pl->passes == 3 ? launch_conv<128, 3, true, false>(pl, stream) : launch_conv<128, 1, true, false>(pl, stream);
    }
#define D2T_RUN(bn, ps) (pl->pair ? launch_conv<bn, ps, false, true>(pl, stream) : launch_conv<bn, ps, false, false>(pl, stream))
    if (pl->passes == 16) {
        if (pl->args.mask) {        // backward-data plans: the epilogue with the ReLU mask compiled in (never pairs)
            if (pl->BN == 64) return launch_conv<64, 16, false, false, false, false, false, true>(pl, stream);
            if (pl->epi2 && pl->ares) return launch_conv<128, 16, false, false, true, false, false, true, true>(pl, stream);
            return pl->epi2 ? launch_conv<128, 16, false, false, true, false, false, true>(pl, stream)
                            : launch_conv<128, 16, false, false, false, false, false, true>(pl, stream);
        }
        if (pl->BN == 64) return launch_conv<64, 16, false, false>(pl, stream);
        if (pl->pair) return launch_conv<128, 16, false, true>(pl, stream);
        if (pl->epi2 && pl->ares) return launch_conv<128, 16, false, false, true, false, false, false, true>(pl, stream);
        return pl->epi2 ? launch_conv<128, 16, false, false, true>(pl, stream) : launch_conv<128, 16, false, false>(pl, stream);
    }
    if (pl->passes == 3) return pl->BN == 64 ? D2T_RUN(64, 3) : D2T_RUN(128, 3);
    return pl->BN == 64 ? D2T_RUN(64, 1) : D2T_RUN(128, 1);
#undef D2T_RUN
}

// ------------------------------------------------------------------ layer chains (conv_chain)
struct d2t_conv_chain {
    ChainLayer* dev;
    int* gsync;
    int n;
    int private_scratch;
};

extern "C" int d2t_conv_plan_chainable(const d2t_conv_plan* pl) {
    return pl && pl->passes == 16 && !pl->corr && !pl->pair && !pl->wgrad && !pl->corrb && !pl->args.mask && !pl->args.stem &&
           !pl->args.done_prev && !pl->args.done_self && pl->args.amax_in && (pl->BN == 128 || (pl->BN == 64 && !pl->epi2)) ? 1 : 0;
}

extern "C" size_t d2t_conv_chain_bytes(int n_layers) {
    return 256 + (size_t)(n_layers > 0 ? n_layers : 0) * sizeof(ChainLayer);
}

// plans[i] runs as layer i.  sync_before[i] != 0: layer i reads (input, residual) what one of the layers since the last such
// mark writes -- the CTAs meet at a grid-wide barrier first.  dev_buf: d2t_conv_chain_bytes(n) bytes of device memory,
// 256-byte aligned, owned by the caller for the life of the chain.  The plans' tensor maps and arguments are COPIED: build
// the chain after every d2t_conv_plan_set_* call.  All plans must use the same stream-K scratch.
extern "C" d2t_conv_chain* d2t_conv_chain_create(const d2t_conv_plan* const* plans, const int* sync_before, int n, void* dev_buf,
                                                 size_t bytes) {
    if (!plans || n <= 0 || !dev_buf || ((uintptr_t)dev_buf & 255) != 0 || bytes < d2t_conv_chain_bytes(n)) {
        set_error("d2t_conv_chain_create: bad arguments");
        return nullptr;
    }
    ChainLayer* host = nullptr;
    if (posix_memalign(reinterpret_cast<void**>(&host), 128, sizeof(ChainLayer) * (size_t)n) != 0) {
        set_error("d2t_conv_chain_create: out of memory");
        return nullptr;
    }
    for (int i = 0; i < n; ++i) {
        const d2t_conv_plan* pl = plans[i];
        if (!d2t_conv_plan_chainable(pl) || pl->args.sk_scratch != plans[0]->args.sk_scratch) {
            set_error("d2t_conv_chain_create: layer %d is not a plain 3xFP16 convolution plan on the chain's scratch", i);
            free(host);
            return nullptr;
        }
        ChainLayer& L = host[i];
        memset(&L, 0, sizeof(L));
        L.tmA = pl->tmA; L.tmB_hi = pl->tmB_hi; L.tmB_lo = pl->tmB_lo; L.tmO = pl->tmO; L.tmR = pl->tmR;
        L.args = pl->args;
        L.args.sk_epoch = i + 1;            // unique per layer: un-barriered neighbours must not mistake each other's partials
        L.variant = pl->BN == 64 ? 2 : (pl->epi2 ? 1 : 0);      // (the A-resident sub-variant is a stand-alone kernel: plain EPI2 here)
        L.sync_before = sync_before ? (sync_before[i] != 0) : 1;
        if (getenv("D2T_CHAIN_NOSYNC") && atoi(getenv("D2T_CHAIN_NOSYNC")) == 1) L.sync_before = 0;    // timing experiments only (wrong results)
    }
    d2t_conv_chain* ch = new (std::nothrow) d2t_conv_chain();
    if (!ch) {
        free(host);
        set_error("d2t_conv_chain_create: out of memory");
        return nullptr;
    }
    ch->gsync = reinterpret_cast<int*>(dev_buf);
    ch->dev = reinterpret_cast<ChainLayer*>(reinterpret_cast<char*>(dev_buf) + 256);
    ch->n = n;
    ch->private_scratch = plans[0]->private_scratch;
    cudaError_t e = cudaMemset(dev_buf, 0, 256);
    if (e == cudaSuccess) e = cudaMemcpy(ch->dev, host, sizeof(ChainLayer) * (size_t)n, cudaMemcpyHostToDevice);
    free(host);
    if (e != cudaSuccess) {
        set_error("d2t_conv_chain_create: %s", cudaGetErrorString(e));
        delete ch;
        return nullptr;
    }
    return ch;
}

extern "C" int d2t_conv_chain_run(const d2t_conv_chain* ch, cudaStream_t stream) {
    D2T_REQUIRE(ch, "d2t_conv_chain_run: null chain");
    static SmemAttrOnce once;
    if (!once.ensure(conv_chain, kChainSmem, "conv chain smem attr")) return 0;
    std::unique_lock<std::mutex> lock(g_sk_mu, std::defer_lock);
    if (!ch->private_scratch) {                       // default scratch: launches stay stream-ordered (see SkScratch)
        lock.lock();
        if (!sk_serialize(stream)) return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count());                   // one CTA per SM: the grid-wide barrier needs every CTA resident
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = kChainSmem;
    cfg.stream = stream;
    // D2T_CHAIN_COOP=0 (experiments): plain launch -- co-residency then rests on nothing else occupying the SMs
    static const bool coop = [] { const char* e = getenv("D2T_CHAIN_COOP"); return !(e && e[0] == '0'); }();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = coop ? 1 : 0;
    const ChainLayer* layers = ch->dev;
    D2T_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_chain, layers, ch->n, ch->gsync), "conv_chain launch");
    return 1;
}

extern "C" void d2t_conv_chain_destroy(d2t_conv_chain* ch) {
    delete ch;
}
