"""GPU parity tests: the sm_100a kernels of libd2t_b200.so (through the reference-shaped operator
API / the C ABI) against (i) the CPU oracle, (ii) the reference's own CUDA kernels recompiled for
sm_100a (oracle/_ref/libref_oracle.so) on identical inputs, (iii) the committed golden fixtures.
Bars: integer outputs (bins, mapping channels, argmax, NMS keep sets) bit-exact; floating point
within the tolerance written at each assert (north star: 1e-4 relative)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import cases
import common
from d2t_b200 import ops
from d2t_b200._lib import lib
from oracle import ref_cuda

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # BASELINE.json north_star: "fp outputs within 1e-4 rel of the reference"


# PSRoI values: the tuned kernel sums each bin exactly (fp64 summed-area table) and rounds once;
# the reference adds sequentially in fp32, so they differ by the reference's own rounding error.
PS_RTOL, PS_ATOL = 1e-5, 2e-6


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


def close(a, b, rtol=RTOL, atol=1e-6, msg=""):
    np.testing.assert_allclose(npy(a) if torch.is_tensor(a) else a, npy(b) if torch.is_tensor(b) else b,
                               rtol=rtol, atol=atol, err_msg=msg)


def have_ref():
    return ref_cuda.available()


# =========================================================================== PSRoI
def _psroi_module(D):
    from model.psroi_pooling.modules.psroi_pool import _PSRoIPooling
    return _PSRoIPooling(7, 7, 1.0 / 16.0, 7, D)


def test_psroi_bins_bit_exact_vs_oracle(oracle):
    rois = np.concatenate([common.make_rois(2000, 2, seed=21), cases.psroi_cases()["r7"]["rois"]], 0)
    bins = npy(ops.psroi_bins(cu(rois), 7, 7, 1.0 / 16.0, 38, 63))
    feat = np.zeros((2, 49, 38, 63), np.float32)
    _, _, want = oracle.psroi_forward(feat, rois, 1.0 / 16.0, 7, 7, 7, 1, want_bins=True)
    np.testing.assert_array_equal(bins, want)
    # how often would the as-written (unfused) arithmetic have differed?  
    _, _, unfused = oracle.psroi_forward(feat, rois, 1.0 / 16.0, 7, 7, 7, 1, contract=0, want_bins=True)
    print("fused != unfused bins:", int((unfused != want).any(-1).sum()), "of", want.shape[0] * 49)


def test_psroi_golden_case_forward_backward(oracle, golden_cuda):
    c = cases.psroi_cases()["r7"]
    feat, rois = cu(c["feat"]).requires_grad_(True), cu(c["rois"])
    from model.psroi_pooling.functions.psroi_pool import PSRoIPoolFunction
    fn = PSRoIPoolFunction(7, 7, c["scale"], 7, c["D"])
    top = fn(feat, rois)
    close(top, golden_cuda["psroi_r7_top"], rtol=PS_RTOL, atol=PS_ATOL)
    np.testing.assert_array_equal(npy(fn.mappingchannel), golden_cuda["psroi_r7_map"])
    _, kmap = ops.psroi_forward(feat.detach(), rois, 7, 7, c["scale"], 7, c["D"], want_mapping=True)
    np.testing.assert_array_equal(npy(kmap), golden_cuda["psroi_r7_map"])           # kernel-written map, bit-exact
    assert fn.rois is rois and tuple(fn.feature_size) == tuple(feat.shape)
    top.backward(cu(c["gtop"]))
    close(feat.grad, golden_cuda["psroi_r7_grad"], rtol=1e-5)
    want = oracle.psroi_backward(c["gtop"], c["rois"], c["feat"].shape, c["scale"], 7, 7, 7, c["D"])
    close(feat.grad, want, rtol=1e-5)


@pytest.mark.parametrize("D,B,R", [(4, 2, 300), (31, 2, 300), (30, 1, 2000)])
def test_psroi_full_size_vs_reference_kernel(D, B, R):
    """BASELINE configs 2 and 5 shapes against the reference kernel itself on the same inputs."""
    if not have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built")
    torch.manual_seed(20)
    feat = torch.randn(B, D * 49, 38, 63, device="cuda")
    rois = cu(common.make_rois(R, B, seed=21, shuffle=(D == 4)))
    top, mapping = ops.psroi_forward(feat, rois, 7, 7, 1.0 / 16.0, 7, D, want_mapping=True)
    rtop, rmap = ref_cuda.psroi_forward(feat, rois, 1.0 / 16.0, 7, 7, 7, D)
    assert torch.equal(mapping, rmap)                                                # bit-exact channel map
    close(top, rtop, rtol=PS_RTOL, atol=PS_ATOL)
    # and against the exact window means (fp64 on the host) the tuned kernel must be the closer one
    assert float((top - rtop).abs().max()) < 4e-6
    gt = torch.randn_like(top)
    g = ops.psroi_backward(gt, rois, feat.shape, 7, 7, 1.0 / 16.0, 7, D)
    rg = ref_cuda.psroi_backward(gt, rmap, rois, feat.shape, 1.0 / 16.0, 7, 7, D)
    close(g, rg, rtol=1e-4, atol=1e-5)


def test_psroi_integer_tables_against_exact_tables(monkeypatch):
    """The forward has two table kernels (csrc/psroi.cu): fp64 summed-area tables (exactly rounded bin means,
    D2T_PSROI_INT=0) and, chosen by default when there is more than one item per SM, per-plane fixed-point int32 tables
    (window sums exact in integers; the only error is the quantisation, <= 2^-30 of the plane's L1 norm per cell).  Same
    bins, same mapping; values within that bound -- also for large-magnitude features, unsorted rois, ragged counts."""
    for B, D, R, shuffle, amp in ((2, 30, 2000, False, 1.0), (2, 30, 500, True, 300.0), (3, 8, 77, True, 1e-3),
                                  (4, 31, 300, False, 1.0)):
        torch.manual_seed(20 + R)
        feat = torch.randn(B, D * 49, 38, 63, device="cuda") * amp
        feat[0, 5] += 3.0 * amp                                        # a plane with a large mean (no cancellation)
        rois = cu(common.make_rois(R, B, seed=21, shuffle=shuffle))
        lib().d2t_psroi_set_mode(0, 0)
        exact, map_x = ops.psroi_forward(feat, rois, 7, 7, 1.0 / 16.0, 7, D, want_mapping=True)
        lib().d2t_psroi_set_mode(4, 0)
        fixed, map_i = ops.psroi_forward(feat, rois, 7, 7, 1.0 / 16.0, 7, D, want_mapping=True)
        lib().d2t_psroi_set_mode(-1, 0)
        auto, _ = ops.psroi_forward(feat, rois, 7, 7, 1.0 / 16.0, 7, D)
        assert torch.equal(map_x, map_i)
        l1 = feat.abs().sum((2, 3))[:, :D * 49].max()                   # largest plane L1 norm
        bound = float(l1) * 2.0 ** -30 * 1.05 + 3e-7 * float(exact.abs().max())   # quantisation + fp32 rounding of the mean
        err = float((fixed - exact).abs().max())
        assert err <= bound, (err, bound, B, D, R)
        # the default is one of the two
        assert torch.equal(auto, fixed) or torch.equal(auto, exact)


@pytest.mark.parametrize("B,D,H,W", [(2, 2, 10, 12), (1, 3, 5, 7), (2, 2, 30, 64), (3, 1, 45, 33), (2, 4, 19, 63),
                                     (1, 2, 1, 1), (2, 31, 24, 40)])
def test_psroi_integer_tables_other_geometries(oracle, monkeypatch, B, D, H, W):
    """The integer-table kernel is selected for any 7x7 geometry with planes up to 64 wide once there is more than one
    item per SM; its scans have separate code for odd widths (thread per row), even widths (warp shuffles), rows / columns
    shorter than one batch.  Forced here (D2T_PSROI_INT=4) on small planes and compared with the CPU oracle: bins and
    mapping bit-exact, values within the quantisation bound; the last case is selected by the library on its own."""
    torch.manual_seed(H * 100 + W)
    feat = torch.randn(B, D * 49, H, W, device="cuda")
    rois = common.make_rois(40, B, height=H * 16, width=W * 16, seed=H + W, lo=8, hi=max(64.0, 12.0 * min(H, W)),
                            shuffle=True)
    top, mapping = ops.psroi_forward(feat, cu(rois), 7, 7, 1 / 16., 7, D, want_mapping=True)
    lib().d2t_psroi_set_mode(-1, 0)
    want, wmap = oracle.psroi_forward(npy(feat), rois, 1 / 16., 7, 7, 7, D)
    np.testing.assert_array_equal(npy(mapping), wmap)
    l1 = float(feat.abs().sum((2, 3)).max())
    np.testing.assert_allclose(npy(top), want, rtol=PS_RTOL, atol=PS_ATOL + l1 * 2.0 ** -30)
    lib().d2t_psroi_set_mode(0, 0)
    exact, _ = ops.psroi_forward(feat, cu(rois), 7, 7, 1 / 16., 7, D)
    lib().d2t_psroi_set_mode(-1, 0)
    assert float((top - exact).abs().max()) <= l1 * 2.0 ** -30 * 1.05 + 3e-7 * float(exact.abs().max())


def test_psroi_backward_integer_tables_experiment(monkeypatch):
    """D2T_PSROI_BWD_INT=1 (csrc/psroi.cu, psroi_bwd_isat_mc) against the default fp64 difference tables: every dv is
    rounded to a multiple of 2^-k with sum |dv| 2^k < 2^30, so a cell covered by n bins is within n 2^-30 sum|dv|."""
    for B, D, R, shuffle in ((2, 30, 2000, False), (3, 4, 77, True), (1, 2, 5, False)):
        torch.manual_seed(R)
        rois = cu(common.make_rois(R, B, seed=21, shuffle=shuffle))
        gt = torch.randn(rois.size(0), D, 7, 7, device="cuda")
        shape = (B, D * 49 + 3, 38, 63)
        lib().d2t_psroi_set_mode(-1, 2)
        want = ops.psroi_backward(gt, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
        lib().d2t_psroi_set_mode(-1, 1)
        got = ops.psroi_backward(gt, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
        lib().d2t_psroi_set_mode(-1, 0)
        assert float(got[:, D * 49:].abs().max()) == 0.0
        err = float((got - want).abs().max())
        assert err <= 2e-5 * max(1.0, float(want.abs().max())), err


def test_psroi_backward_two_limb_tables_exact():
    """The default backward (csrc/psroi.cu, psroi_bwd_limb): every dv = top_diff / area is split into two 32-bit limbs of a
    fixed-point integer and accumulated with native shared-memory integer atomics -- exact, order-independent sums, rounded
    to fp32 once.  Against the fp64 difference tables (mode 2, themselves exact to 1e-13): equal to one fp32 rounding, for large / tiny gradient magnitudes, unsorted rois, accumulate=1; bit-identical from run to run; a NaN / Inf
    gradient reaches exactly the cells the fp64 kernel gives it to."""
    for B, D, R, shuffle, amp in ((2, 30, 2000, False, 1.0), (3, 4, 77, True, 3e4), (1, 2, 5, False, 1e-6), (2, 8, 300, True, 1.0)):
        torch.manual_seed(R)
        rois = cu(common.make_rois(R, B, seed=21, shuffle=shuffle))
        gt = torch.randn(rois.size(0), D, 7, 7, device="cuda") * amp
        shape = (B, D * 49 + 3, 38, 63)
        lib().d2t_psroi_set_mode(-1, 2)
        want = ops.psroi_backward(gt, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
        for mode in (0,):
            lib().d2t_psroi_set_mode(-1, mode)
            got = ops.psroi_backward(gt, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
            again = ops.psroi_backward(gt, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
            assert torch.equal(got, again)                                         # deterministic
            assert float(got[:, D * 49:].abs().max()) == 0.0
            # one fp32 rounding of the same exact sum on both sides: <= 1 ulp of the cell, + the 2^-34 max|g| term bound
            err = (got - want).abs()
            bound = 1.2e-7 * want.abs() + 1e-9 * float(gt.abs().max())
            assert bool((err <= bound).all()), float((err - bound).max())
        lib().d2t_psroi_set_mode(-1, 0)
    # non-finite gradients: the item falls back to the fp64 loops, everything else stays on the integer path
    B, D, R = 2, 4, 120
    rois = cu(common.make_rois(R, B, seed=3))
    gt = torch.randn(R * B, D, 7, 7, device="cuda")
    gt[7, 1, 3, 2] = float("nan")
    gt[130, 2, 0, 6] = float("inf")
    shape = (B, D * 49, 38, 63)
    lib().d2t_psroi_set_mode(-1, 2)
    want = ops.psroi_backward(gt, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
    lib().d2t_psroi_set_mode(-1, 0)
    got = ops.psroi_backward(gt, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and bool(torch.isnan(got).any())
    assert torch.equal(torch.isinf(got), torch.isinf(want))
    fin = torch.isfinite(want)
    assert float((got[fin] - want[fin]).abs().max()) <= 1e-6
    # accumulate = 1 adds to what is there
    base = torch.randn(shape, device="cuda")
    gt2 = torch.randn(R * B, D, 7, 7, device="cuda")
    g0 = ops.psroi_backward(gt2, rois, shape, 7, 7, 1.0 / 16.0, 7, D)
    acc = base.clone()
    ws = torch.empty(lib().d2t_psroi_workspace_bytes(R * B, B, 7, 7), dtype=torch.uint8, device="cuda")
    from d2t_b200._lib import check
    check(lib().d2t_psroi_backward(gt2.data_ptr(), B, D * 49, 38, 63, rois.data_ptr(), R * B, 1.0 / 16.0, 7, 7, 7, D,
                                   acc.data_ptr(), 1, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
          "d2t_psroi_backward")
    assert float((acc - (base + g0)).abs().max()) <= 1e-6


def test_psroi_edge_cases(oracle):
    feat = torch.randn(2, 196, 9, 11, device="cuda")
    # empty roi list
    top, _ = ops.psroi_forward(feat, torch.zeros(0, 5, device="cuda"), 7, 7, 1 / 16., 7, 4)
    assert top.shape == (0, 4, 7, 7)
    g = ops.psroi_backward(torch.zeros(0, 4, 7, 7, device="cuda"), torch.zeros(0, 5, device="cuda"), feat.shape, 7, 7,
                           1 / 16., 7, 4)
    assert float(g.abs().max()) == 0.0
    # all rois on image 1, none on image 0; ragged & unsorted image indices
    rois = common.make_rois(17, 2, height=144, width=176, seed=5, lo=4, hi=200, shuffle=True)
    rois1 = rois.copy(); rois1[:, 0] = 1
    for r in (rois, rois1):
        top, mapping = ops.psroi_forward(feat, cu(r), 7, 7, 1 / 16., 7, 4, want_mapping=True)
        want, wmap = oracle.psroi_forward(npy(feat), r, 1 / 16., 7, 7, 7, 4)
        close(top, want, rtol=PS_RTOL, atol=PS_ATOL)
        np.testing.assert_array_equal(npy(mapping), wmap)
    # out-of-range image indices give zero rows (the reference would read out of bounds)
    rbad = rois.copy(); rbad[::3, 0] = 7; rbad[1::3, 0] = -2
    top, _ = ops.psroi_forward(feat, cu(rbad), 7, 7, 1 / 16., 7, 4)
    assert float(top[::3].abs().max()) == 0.0 and float(top[1::3].abs().max()) == 0.0
    want, _ = oracle.psroi_forward(npy(feat), rois[2::3], 1 / 16., 7, 7, 7, 4)
    close(top[2::3], want, rtol=PS_RTOL, atol=PS_ATOL)
    # generic path: pooled 3x3 / group 3, and a plane too large for shared memory
    feat3 = torch.randn(1, 18, 20, 30, device="cuda")
    r3 = common.make_rois(33, 1, height=320, width=480, seed=6)
    top, mapping = ops.psroi_forward(feat3, cu(r3), 3, 3, 1 / 16., 3, 2, want_mapping=True)
    want, wmap = oracle.psroi_forward(npy(feat3), r3, 1 / 16., 3, 3, 3, 2)
    np.testing.assert_array_equal(npy(top), want)     # generic kernel: the reference's own summation order
    np.testing.assert_array_equal(npy(mapping), wmap)
    g = ops.psroi_backward(cu(common.randn(want.shape, 7)), cu(r3), feat3.shape, 3, 3, 1 / 16., 3, 2)
    close(g, oracle.psroi_backward(common.randn(want.shape, 7), r3, feat3.shape, 1 / 16., 3, 3, 3, 2), rtol=1e-5)
    big = torch.randn(1, 49, 100, 160, device="cuda")
    rb = common.make_rois(50, 1, height=1600, width=2560, seed=8, hi=1500)
    top, _ = ops.psroi_forward(big, cu(rb), 7, 7, 1 / 16., 7, 1)
    np.testing.assert_array_equal(npy(top), oracle.psroi_forward(npy(big), rb, 1 / 16., 7, 7, 7, 1)[0])


def test_psroi_legacy_launcher_symbols(oracle):
    """PSROIPoolForwardLauncher / PSROIPoolBackwardLauncher with the reference's argument order
    (psroi_pooling_kernel.h:8-14: pooled_width BEFORE pooled_height in backward)."""
    c = cases.psroi_cases()["r7"]
    feat, rois = cu(c["feat"]), cu(c["rois"])
    R, D = rois.shape[0], c["D"]
    top = torch.full((R, D, 7, 7), 7.0, device="cuda")
    mapping = torch.zeros(R, D, 7, 7, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert lib().PSROIPoolForwardLauncher(feat.data_ptr(), c["scale"], R, 20, 30, D * 49, 7, 7, rois.data_ptr(), 7, D,
                                          top.data_ptr(), mapping.data_ptr(), st) == 1
    want, wmap = oracle.psroi_forward(c["feat"], c["rois"], c["scale"], 7, 7, 7, D)
    np.testing.assert_array_equal(npy(top), want)
    np.testing.assert_array_equal(npy(mapping), wmap)
    grad = torch.zeros_like(feat)
    gt = cu(c["gtop"])
    assert lib().PSROIPoolBackwardLauncher(gt.data_ptr(), mapping.data_ptr(), 2, R, c["scale"], D * 49, 20, 30, 7, 7, D,
                                           grad.data_ptr(), rois.data_ptr(), st) == 1
    close(grad, oracle.psroi_backward(c["gtop"], c["rois"], c["feat"].shape, c["scale"], 7, 7, 7, D), rtol=1e-5)


# =========================================================================== NMS
def test_nms_golden_keep_sets(golden_cuda):
    from model.nms.nms_wrapper import nms
    for name, (dets, thresh) in cases.nms_cases().items():
        keep = nms(cu(dets), thresh)
        assert keep.dtype == torch.int32 and keep.dim() == 2 and keep.shape[1] == 1
        np.testing.assert_array_equal(npy(keep).reshape(-1), golden_cuda["nms_%s" % name], err_msg=name)


@pytest.mark.parametrize("n,thresh,clustered", [(2, 0.7, False), (63, 0.5, True), (64, 0.5, True), (65, 0.3, True),
                                                (129, 0.7, True), (300, 0.3, False), (2000, 0.7, False),
                                                (6000, 0.7, True), (12000, 0.7, False), (12000, 0.7, True)])
def test_nms_vs_oracle_and_reference_kernel(oracle, n, thresh, clustered):
    dets = common.make_clustered_dets(n, seed=100 + n) if clustered else common.make_dets(n, seed=100 + n)
    keep = npy(ops.nms(cu(dets), thresh)).reshape(-1)
    np.testing.assert_array_equal(keep, oracle.nms(dets, thresh))
    if have_ref():
        np.testing.assert_array_equal(keep, npy(ref_cuda.nms(cu(dets), thresh)).reshape(-1))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_nms_quotients_on_the_threshold(oracle, seed):
    """The mask kernel decides `inter / den > thresh` without the division unless the quotient is within 2^-21 of the
    threshold (csrc/nms.cu: quotient_gt).  Boxes on a coarse integer lattice share a handful of IoU values; with the threshold
    set to the fp32 quotient of one such pair (and to its fp32 neighbours) thousands of comparisons land exactly on, one ulp
    above and one ulp below it -- the keep-sets must still be the oracle's and the reference kernel's, bit for bit."""
    rng = np.random.RandomState(900 + seed)
    n = 1500
    x1 = rng.randint(0, 30, n).astype(np.float32)
    y1 = rng.randint(0, 30, n).astype(np.float32)
    w = rng.randint(4, 20, n).astype(np.float32)
    h = rng.randint(4, 20, n).astype(np.float32)
    score = np.sort(rng.rand(n).astype(np.float32))[::-1]
    dets = np.stack([x1, y1, x1 + w, y1 + h, score], 1).astype(np.float32)
    # the fp32 IoU of a few overlapping pairs, computed as the kernels do (nms_cuda_kernel.cu:31-39)
    cands = []
    for i, j in rng.randint(0, n, (400, 2)):
        a, b = dets[i], dets[j]
        iw = np.float32(max(min(a[2], b[2]) - max(a[0], b[0]) + 1, 0))
        ih = np.float32(max(min(a[3], b[3]) - max(a[1], b[1]) + 1, 0))
        inter = np.float32(iw * ih)
        sa = np.float32((a[2] - a[0] + 1) * (a[3] - a[1] + 1))
        sb = np.float32((b[2] - b[0] + 1) * (b[3] - b[1] + 1))
        if inter > 0 and i != j:
            cands.append(np.float32(inter / np.float32(sa + sb - inter)))
    cands = [c for c in cands if 0.2 < c < 0.8][:4]
    assert cands
    for c in cands:
        for thresh in (c, np.nextafter(c, np.float32(1)), np.nextafter(c, np.float32(0))):
            keep = npy(ops.nms(cu(dets), float(thresh))).reshape(-1)
            np.testing.assert_array_equal(keep, oracle.nms(dets, float(thresh)))
            if have_ref():
                np.testing.assert_array_equal(keep, npy(ref_cuda.nms(cu(dets), float(thresh))).reshape(-1))


def test_nms_batched_ragged_and_capped(oracle):
    B, N = 5, 700
    dets = np.stack([common.make_clustered_dets(N, seed=200 + b) for b in range(B)])
    n_valid = np.array([700, 0, 1, 65, 333], np.int32)
    keep, num = ops.nms_batched(cu(dets), 0.7, max_keep=0, n_valid=cu(n_valid))
    keep, num = npy(keep), npy(num)
    for b in range(B):
        want = oracle.nms(dets[b, : n_valid[b]], 0.7)
        assert num[b] == len(want)
        np.testing.assert_array_equal(keep[b, : num[b]], want)
    keep, num = ops.nms_batched(cu(dets), 0.7, max_keep=40)
    for b in range(B):
        want = oracle.nms(dets[b], 0.7)[:40]
        assert int(num[b]) == len(want)
        np.testing.assert_array_equal(npy(keep[b, : len(want)]), want)


@pytest.mark.parametrize("N,K,thresh", [(6000, 300, 0.7), (12000, 2000, 0.7), (6000, 300, 0.05), (700, 40, 0.3), (130, 2048, 0.7),
                                        (6000, 64, 1.0), (6000, 300, 0.0)])
def test_nms_capped_mask_free_kernel(oracle, N, K, thresh):
    """capped lists (the proposal step) run on nms_greedy -- no N x N mask, one CTA per list, candidates tested against the
    kept boxes only -- and must give the keep-sets of the oracle and of the mask + sweep pair, bit for bit: spread and
    clustered boxes, ragged / empty lists, degenerate boxes, caps that are and are not reached."""
    from d2t_b200._lib import lib
    dets = np.stack([common.make_dets(N, seed=400), common.make_clustered_dets(N, seed=401), common.make_dets(N, seed=402),
                     common.make_clustered_dets(N, seed=403), common.make_clustered_dets(N, seed=404)])
    dets[2, 10:20, 2] = dets[2, 10:20, 0] - 5          # inverted boxes
    dets[2, 30:40, :4] = dets[2, 30, :4]               # exact duplicates
    dets[2, 50:60, :4] = 0.0                           # zero boxes
    dets[2, 60:64, :4] = np.array([5, 5, 4, 4])        # 0/0 IoU = NaN, never suppresses
    n_valid = np.array([N, N, min(N, 1000), N - 37, 0], np.int32)
    outs = []
    try:
        for mode in (1, 0):
            lib().d2t_nms_set_mode(mode)
            assert lib().d2t_nms_launch_count(N, K) == (1 if mode else 2)
            keep, num = ops.nms_batched(cu(dets), thresh, max_keep=K, n_valid=cu(n_valid))
            outs.append((npy(keep), npy(num)))
    finally:
        lib().d2t_nms_set_mode(-1)
    for b in range(5):
        want = oracle.nms(dets[b, : n_valid[b]], thresh)[:K]
        for keep, num in outs:
            assert num[b] == len(want), (b, num[b], len(want))
            np.testing.assert_array_equal(keep[b, : num[b]], want)


def test_nms_prefix_pass_and_fallback(oracle, monkeypatch):
    """max_keep > 0 on long lists with D2T_NMS_PREFIX=1: d2t_nms_batched first decides a prefix (4 * max_keep boxes); lists
    that reach max_keep inside it are final, the others take the full pass.  One batch mixes short / empty / long lists."""
    from d2t_b200._lib import lib
    assert lib().d2t_nms_prefix(6000, 300) == 0           # opt-in
    monkeypatch.setenv("D2T_NMS_PREFIX", "1")
    lib().d2t_nms_set_mode(0)                             # (the mask + sweep pair: capped lists default to the mask-free kernel)
    monkeypatch.setattr(ops, "_nms_mode_restore", lib().d2t_nms_set_mode, raising=False)
    B, N, K = 5, 6000, 300
    assert lib().d2t_nms_prefix(N, K) == 1216 and lib().d2t_nms_prefix(N, 0) == 0 and lib().d2t_nms_prefix(2000, K) == 0
    dets = np.stack([common.make_dets(N, seed=300), common.make_clustered_dets(N, seed=301), common.make_dets(N, seed=302),
                     common.make_clustered_dets(N, seed=303), common.make_clustered_dets(N, seed=304)])
    n_valid = np.array([6000, 6000, 1000, 5000, 0], np.int32)
    keep, num = ops.nms_batched(cu(dets), 0.7, max_keep=K, n_valid=cu(n_valid))
    keep, num = npy(keep), npy(num)
    in_prefix = []
    for b in range(B):
        full = oracle.nms(dets[b, : n_valid[b]], 0.7)
        want = full[:K]
        assert num[b] == len(want), (b, num[b], len(want))
        np.testing.assert_array_equal(keep[b, : num[b]], want)
        in_prefix.append(len(want) == K and want[-1] < 1216)
    assert in_prefix[0]                                # (spread boxes: final after the prefix pass)
    # a threshold that suppresses almost everything: fewer than max_keep survivors inside the prefix, so the full pass
    # runs and must find the late ones
    keep, num = ops.nms_batched(cu(dets[1:2]), 0.05, max_keep=K)
    want = oracle.nms(dets[1], 0.05)[:K]
    assert int(num[0]) == len(want) and (len(want) < K or want[-1] >= 1216)
    np.testing.assert_array_equal(npy(keep)[0, : len(want)], want)
    lib().d2t_nms_set_mode(-1)


def test_nms_degenerate_boxes(oracle):
    rng = np.random.RandomState(9)
    dets = common.make_dets(200, seed=9)
    dets[10:20, 2] = dets[10:20, 0] - 5          # inverted boxes (negative width)
    dets[30:40, :4] = dets[30, :4]               # exact duplicates
    dets[50:60, :4] = 0.0                        # zero boxes
    dets[60:64, :4] = np.array([5, 5, 4, 4])     # width 0 -> 0/0 IoU = NaN, never suppresses
    for thresh in (0.0, 0.3, 0.7, 1.0):
        np.testing.assert_array_equal(npy(ops.nms(cu(dets), thresh)).reshape(-1), oracle.nms(dets, thresh))


def test_nms_legacy_symbol(oracle):
    dets = common.make_dets(500, seed=77)
    d = cu(dets)
    keep = torch.zeros(500, dtype=torch.int32, device="cuda")
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    lib().nms_cuda_compute(keep.data_ptr(), num.data_ptr(), d.data_ptr(), 500, 5, 0.7)
    want = oracle.nms(dets, 0.7)
    assert int(num.item()) == len(want)
    np.testing.assert_array_equal(npy(keep)[: len(want)], want)


# =========================================================================== correlation
def test_correlation_golden_cases(oracle, golden_cuda):
    from model.correlation.modules.correlation import Correlation
    for name, c in cases.corr_cases().items():
        p = c["params"]
        a, b = cu(c["in1"]).requires_grad_(True), cu(c["in2"]).requires_grad_(True)
        out = Correlation(pad_size=p[0], kernel_size=p[1], max_displacement=p[2], stride1=p[3], stride2=p[4])(a, b)
        close(out, golden_cuda["corr_%s_out" % name], rtol=RTOL, atol=1e-6, msg=name)
        go = common.randn(tuple(out.shape), c["gseed"])
        out.backward(cu(go))
        t1, t2 = oracle.correlation_backward_true(c["in1"], c["in2"], go, *p)
        close(a.grad, t1, rtol=RTOL, atol=1e-6, msg=name)
        close(b.grad, t2, rtol=RTOL, atol=1e-6, msg=name)
        if p[0] == p[2] and p[1] == 1 and p[3] == 1:   # the reference backward is the adjoint here: must agree with it too
            close(a.grad, golden_cuda["corr_%s_g1" % name], rtol=RTOL, atol=1e-6, msg=name)
            close(b.grad, golden_cuda["corr_%s_g2" % name], rtol=RTOL, atol=1e-6, msg=name)


@pytest.mark.parametrize("C,H,W,p,B", [(1024, 38, 63, (8, 1, 8, 1, 1), 1), (2048, 38, 63, (8, 1, 8, 1, 1), 2),
                                       (512, 75, 125, (8, 1, 8, 2, 2), 2), (1024, 38, 63, (8, 1, 8, 1, 1), 8)])
def test_correlation_full_size_vs_reference_kernel(C, H, W, p, B):
    """The three D&T correlation shapes (rfcn.py:58-60) against the reference kernel itself."""
    if not have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built")
    torch.manual_seed(10)
    a, b = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    out = ops.correlation_forward(a, b, *p)
    ref = ref_cuda.correlation_forward(a, b, *p)
    assert out.shape == ref.shape
    # |out| ~ 1/sqrt(C); relative to the output scale, as the north star states it
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < RTOL, err
    close(out, ref, rtol=1e-3, atol=2e-6)


def test_correlation_backward_full_size_vs_reference_kernel():
    if not have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built")
    torch.manual_seed(11)
    B, C, H, W, p = 1, 256, 38, 63, (8, 1, 8, 1, 1)
    a, b = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    go = torch.randn(B, 289, 38, 63, device="cuda")
    g1, g2 = ops.correlation_backward(a, b, go, *p)
    r1, r2 = ref_cuda.correlation_backward(a, b, go, *p)
    for g, r in ((g1, r1), (g2, r2)):
        assert float((g - r).abs().max() / r.abs().max()) < RTOL
    # stride-2 (conv3) case: grad1 is well defined in the reference
    B, C, H, W, p = 1, 64, 75, 125, (8, 1, 8, 2, 2)
    a, b = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    go = torch.randn(B, 81, 38, 63, device="cuda")
    g1, g2 = ops.correlation_backward(a, b, go, *p)
    r1, _ = ref_cuda.correlation_backward(a, b, go, *p, slack=4 * a.numel())
    assert float((g1 - r1).abs().max() / r1.abs().max()) < RTOL
    assert float(g2[:, :, 1::2, :].abs().max()) == 0.0 and float(g2[:, :, :, 1::2].abs().max()) == 0.0


@pytest.mark.parametrize("C,H,W,p,B", [(1024, 38, 63, (8, 1, 8, 1, 1), 2), (512, 75, 125, (8, 1, 8, 2, 2), 2), (96, 21, 30, (4, 1, 4, 1, 1), 3)])
def test_correlation_backward_tensor_core_path_vs_simt(C, H, W, p, B):
    """ops.correlation_backward takes the tensor-core banded-GEMM path (CORRB) for every D&T configuration; the exact-adjoint
    fp32 SIMT gather kernels are the comparator (both deterministic)."""
    torch.manual_seed(13)
    a, b = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    out = ops.correlation_forward(a, b, *p)
    g = torch.randn_like(out)
    t1, t2 = ops.correlation_backward(a, b, g, *p)
    ops.TENSOR_CORE_CORRELATION = False
    try:
        s1, s2 = ops.correlation_backward(a, b, g, *p)
    finally:
        ops.TENSOR_CORE_CORRELATION = True
    for t_, s_ in ((t1, s1), (t2, s2)):
        assert t_.shape == s_.shape
        assert float((t_ - s_).abs().max() / s_.abs().max()) < 2e-5
    only1, none2 = ops.correlation_backward(a, b, g, *p, need2=False)
    assert none2 is None and torch.equal(only1, t1)


def test_correlation_adjoint_property():
    """<corr(a,b), g> is bilinear: d/da <out, g> = grad1 exactly => <out, g> == <a, grad1> == <b, grad2>."""
    torch.manual_seed(12)
    for (C, H, W, p) in [(1024, 38, 63, (8, 1, 8, 1, 1)), (512, 75, 125, (8, 1, 8, 2, 2)), (24, 17, 19, (3, 3, 4, 2, 1))]:
        a, b = torch.randn(2, C, H, W, device="cuda"), torch.randn(2, C, H, W, device="cuda")
        out = ops.correlation_forward(a, b, *p)
        g = torch.randn_like(out)
        g1, g2 = ops.correlation_backward(a, b, g, *p)
        lhs = float((out.double() * g.double()).sum())
        assert abs(lhs - float((a.double() * g1.double()).sum())) < 1e-4 * max(1.0, abs(lhs))
        assert abs(lhs - float((b.double() * g2.double()).sum())) < 1e-4 * max(1.0, abs(lhs))


def test_correlation_legacy_launcher_symbol(oracle):
    c = cases.corr_cases()["d2t_s1"]
    a, b = cu(c["in1"]), cu(c["in2"])
    B, Cc, H, W = a.shape
    out = torch.empty(B, 289, H, W, device="cuda")
    ok = lib().Correlation_forward_cuda_kernel(out.data_ptr(), B, 289, H, W, *out.stride(), a.data_ptr(), Cc, H, W,
                                               *a.stride(), b.data_ptr(), Cc, *b.stride(), None, None, 8, 1, 8, 1, 1, 1,
                                               torch.cuda.current_stream().cuda_stream)
    assert ok == 1
    close(out, oracle.correlation_forward(c["in1"], c["in2"], 8, 1, 8, 1, 1), rtol=RTOL, atol=1e-6)
    # non-contiguous strides are refused with a message instead of silently misreading
    ok = lib().Correlation_forward_cuda_kernel(out.data_ptr(), B, 289, H, W, *out.stride(), a.data_ptr(), Cc, H, W,
                                               1, 2, 3, 4, b.data_ptr(), Cc, *b.stride(), None, None, 8, 1, 8, 1, 1, 1,
                                               torch.cuda.current_stream().cuda_stream)
    assert ok == 0 and b"contiguous" in lib().d2t_last_error()


# =========================================================================== RoIAlign / RoIPool / RoICrop
def test_roi_align_pool_crop_golden(oracle, golden_cuda):
    from model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg
    from model.roi_pooling.modules.roi_pool import _RoIPooling
    from model.roi_crop.modules.roi_crop import _RoICrop
    c = cases.roi_cases()
    rois, grid, scale = cu(c["rois"]), cu(c["grid"]), c["scale"]
    for ah in (7, 8):
        feat = cu(c["feat"]).requires_grad_(True)
        top = RoIAlign(ah, ah, scale)(feat, rois)
        close(top, golden_cuda["align%d_top" % ah], rtol=1e-6, atol=1e-7)
        top.backward(cu(common.randn(tuple(top.shape), 80 + ah)))
        close(feat.grad, golden_cuda["align%d_grad" % ah], rtol=RTOL, atol=1e-5)   # float atomics: order differs
    avg = RoIAlignAvg(7, 7, scale)(cu(c["feat"]), rois)     # 8x8 samples + 2x2 average (roi_align.py:26-29)
    close(avg, torch.nn.functional.avg_pool2d(cu(golden_cuda["align8_top"]), 2, 1), rtol=1e-6, atol=1e-7)
    feat = cu(c["feat"]).requires_grad_(True)
    from model.roi_pooling.functions.roi_pool import RoIPoolFunction
    fn = RoIPoolFunction(7, 7, scale)
    top = fn(feat, rois)
    np.testing.assert_array_equal(npy(top), golden_cuda["pool_top"])
    np.testing.assert_array_equal(npy(fn.argmax), golden_cuda["pool_arg"])
    top.backward(cu(common.randn(tuple(top.shape), 90)))
    close(feat.grad, golden_cuda["pool_grad"], rtol=RTOL, atol=1e-5)
    close(_RoIPooling(7, 7, scale)(cu(c["feat"]), rois), golden_cuda["pool_top"], rtol=0, atol=0)
    feat = cu(c["feat"]).requires_grad_(True)
    g = grid.clone().requires_grad_(True)
    out = _RoICrop()(feat, g)
    close(out, golden_cuda["crop_out"], rtol=1e-5, atol=1e-6)
    out.backward(cu(common.randn(tuple(out.shape), 91)))
    close(feat.grad, golden_cuda["crop_gimg"], rtol=RTOL, atol=1e-5)
    assert float(g.grad.abs().max()) == 0.0


def test_roi_backward_deterministic(oracle, golden_cuda):
    """d2t_roi_{align,pool,crop}_backward_det: every term accumulated as a 64-bit fixed-point integer (order-independent)
    and converted once.  Bit-identical from run to run; equal to the reference goldens and to the float-atomic launchers within
    the float atomics' own noise; tighter than they are against an fp64 host sum; NaN gradients still propagate."""
    c = cases.roi_cases()
    rois, grid, scale = cu(c["rois"]), cu(c["grid"]), c["scale"]
    feat = cu(c["feat"])
    shape = tuple(feat.shape)
    for ah in (7, 8):
        gt = cu(common.randn((rois.size(0), shape[1], ah, ah), 80 + ah))
        a = ops.roi_align_backward(gt, rois, shape, ah, ah, scale, deterministic=True)
        b = ops.roi_align_backward(gt, rois, shape, ah, ah, scale, deterministic=True)
        assert torch.equal(a, b)
        close(a, golden_cuda["align%d_grad" % ah], rtol=RTOL, atol=1e-5)
        close(a, ops.roi_align_backward(gt, rois, shape, ah, ah, scale, deterministic=False), rtol=1e-5, atol=1e-5)
    top, argmax = ops.roi_pool_forward(feat, rois, 7, 7, scale)
    gt = cu(common.randn(tuple(top.shape), 90))
    a = ops.roi_pool_backward(gt, argmax, rois, shape, 7, 7, scale, deterministic=True)
    assert torch.equal(a, ops.roi_pool_backward(gt, argmax, rois, shape, 7, 7, scale, deterministic=True))
    close(a, golden_cuda["pool_grad"], rtol=RTOL, atol=1e-5)
    # exactness: the fixed-point sum against an fp64 scatter-add on the host is within one fp32 rounding
    want = torch.zeros(feat.numel(), dtype=torch.float64)
    am = argmax.flatten().cpu().long()
    ok = am >= 0
    want.index_add_(0, am[ok], gt.flatten().cpu().double()[ok])
    err = (a.flatten().cpu().double() - want).abs()
    assert bool((err <= 6.0e-8 * want.abs() + 1e-10).all()), float(err.max())
    out = ops.roi_crop_forward(feat, grid)
    go = cu(common.randn(tuple(out.shape), 91))
    gi, gg = ops.roi_crop_backward(feat, grid, go, deterministic=True)
    gi2, _ = ops.roi_crop_backward(feat, grid, go, deterministic=True)
    assert torch.equal(gi, gi2) and float(gg.abs().max()) == 0.0
    close(gi, golden_cuda["crop_gimg"], rtol=RTOL, atol=1e-5)
    # full size, twice, plus a NaN gradient: falls back to the float atomics and propagates it
    torch.manual_seed(5)
    featL = torch.randn(2, 256, 38, 63, device="cuda")
    roisL = cu(common.make_rois(256, 2, seed=24))
    gtL = torch.randn(512, 256, 7, 7, device="cuda") * 37.0
    a = ops.roi_align_backward(gtL, roisL, featL.shape, 7, 7, 1 / 16., deterministic=True)
    assert torch.equal(a, ops.roi_align_backward(gtL, roisL, featL.shape, 7, 7, 1 / 16., deterministic=True))
    b = ops.roi_align_backward(gtL, roisL, featL.shape, 7, 7, 1 / 16., deterministic=False)
    assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
    gtL[3, 5, 2, 2] = float("nan")
    a = ops.roi_align_backward(gtL, roisL, featL.shape, 7, 7, 1 / 16., deterministic=True)
    b = ops.roi_align_backward(gtL, roisL, featL.shape, 7, 7, 1 / 16., deterministic=False)
    assert bool(torch.isnan(a).any()) and torch.equal(torch.isnan(a), torch.isnan(b))


def test_roi_crop_is_grid_sample_align_corners():
    """net_utils.py:198-224 names F.grid_sample (torch 0.3 semantics = align_corners=True, zeros
    padding) with the grid's last axis swapped as the comparator for RoICrop."""
    c = cases.roi_cases()
    feat, grid = cu(c["feat"]), cu(c["grid"])
    out = ops.roi_crop_forward(feat, grid)
    per = grid.shape[0] // feat.shape[0]
    src = feat.repeat_interleave(per, 0)
    want = torch.nn.functional.grid_sample(src, grid.flip(-1), mode="bilinear", padding_mode="zeros", align_corners=True)
    close(out, want, rtol=1e-4, atol=1e-5)


def test_roi_ops_full_size_vs_reference_kernels():
    if not have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built")
    torch.manual_seed(13)
    feat = torch.randn(2, 512, 38, 63, device="cuda")
    rois = cu(common.make_rois(300, 2, seed=24))
    a, r = ops.roi_align_forward(feat, rois, 8, 8, 1 / 16.), ref_cuda.roi_align_forward(feat, rois, 1 / 16., 8, 8)
    close(a, r, rtol=1e-5, atol=1e-6)   # double-precision interpolation, contraction order may differ
    gt = torch.randn_like(a)
    close(ops.roi_align_backward(gt, rois, feat.shape, 8, 8, 1 / 16.),
          ref_cuda.roi_align_backward(gt, rois, feat.shape, 1 / 16., 8, 8), rtol=1e-4, atol=1e-5)
    (t, am), (rt, ram) = ops.roi_pool_forward(feat, rois, 7, 7, 1 / 16.), ref_cuda.roi_pool_forward(feat, rois, 1 / 16., 7, 7)
    assert torch.equal(t, rt) and torch.equal(am, ram)
    gt = torch.randn_like(t)
    close(ops.roi_pool_backward(gt, am, rois, feat.shape, 7, 7, 1 / 16.),
          ref_cuda.roi_pool_backward(gt, ram, rois, feat.shape, 1 / 16., 7, 7), rtol=1e-4, atol=1e-5)


# =========================================================================== proposal step
def test_proposal_layer_vs_reference_python_golden(golden_rpn):
    from model.rpn.proposal_layer import _ProposalLayer
    from model.utils.config import cfg
    layer = _ProposalLayer(16, cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS).cuda()
    np.testing.assert_array_equal(npy(layer._anchors), golden_rpn["anchors_d2t"].astype(np.float32))
    cases_ = {"small": dict(B=2, H=10, W=14, seed=30), "config1": dict(B=1, H=19, W=32, seed=31, im_h=300, im_w=500),
              "full": dict(B=1, H=38, W=63, seed=32, im_h=600, im_w=1000)}
    for name, kw in cases_.items():
        prob, deltas, im_info = common.make_rpn_inputs(**kw)
        for key in ("TEST", "TRAIN"):
            rois = layer((cu(prob), cu(deltas), cu(im_info), key))
            ref = golden_rpn["rois_%s_%s" % (name, key)]
            assert rois.shape == ref.shape
            # same proposals in the same order; coordinates differ only by exp() rounding
            close(rois, ref, rtol=1e-5, atol=1e-3, msg="%s %s" % (name, key))


def test_proposal_decode_bit_exact_vs_torch_chain():
    """The fused decode+clip kernel against the reference's op-by-op torch chain run on the GPU."""
    from model.rpn.bbox_transform import bbox_transform_inv, clip_boxes
    from model.rpn.generate_anchors import generate_anchors
    prob, deltas, im_info = common.make_rpn_inputs(B=2, H=38, W=63, seed=35, im_h=600, im_w=1000)
    anchors = torch.from_numpy(generate_anchors(scales=np.array([4, 8, 16, 32]))).float().cuda()
    boxes, scores = ops.proposal_decode(anchors, cu(deltas), cu(prob), cu(im_info), 16)
    A, H, W = 12, 38, 63
    sx, sy = torch.meshgrid(torch.arange(W) * 16, torch.arange(H) * 16, indexing="xy")
    shifts = torch.stack([sx.reshape(-1), sy.reshape(-1), sx.reshape(-1), sy.reshape(-1)], 1).float().cuda()
    all_anchors = (anchors.view(1, A, 4) + shifts.view(-1, 1, 4)).view(1, -1, 4).expand(2, -1, 4)
    d = cu(deltas).permute(0, 2, 3, 1).contiguous().view(2, -1, 4)
    want = clip_boxes(bbox_transform_inv(all_anchors, d, 2), cu(im_info), 2)
    assert torch.equal(boxes, want)
    assert torch.equal(scores, cu(prob)[:, A:].permute(0, 2, 3, 1).contiguous().view(2, -1))


@pytest.mark.parametrize("B,D,R", [(2, 31, 300), (4, 4, 300), (1, 30, 2000), (3, 4, 77)])
def test_psroi_vote_fused(B, D, R):
    """fused PSRoI + 7x7 vote (+ softmax) against AvgPool2d(7) of the unfused operator (rfcn.py:133-140)"""
    torch.manual_seed(R + D)
    feat = torch.randn(B, D * 49, 38, 63, device="cuda")
    rois = cu(common.make_rois(R, B, seed=21, shuffle=(R == 77)))
    rois[5, 0] = 99.0                                   # a roi of no image: zeros like the unfused operator
    top, _ = ops.psroi_forward(feat, rois, 7, 7, 1.0 / 16.0, 7, D)
    want = top.mean((2, 3))
    got = ops.psroi_vote(feat, rois, 7, 7, 1.0 / 16.0, 7, D)
    assert got.shape == (rois.size(0), D)
    assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    sm = ops.psroi_vote(feat, rois, 7, 7, 1.0 / 16.0, 7, D, softmax=True)
    assert float((sm - torch.softmax(want, 1)).abs().max()) < 1e-5


def test_psroi_integer_tables_outlier_guard():
    """a plane with one huge value (or a NaN) must not cost the other bins their accuracy: the multi-CTA kernel detects
    the dynamic range and pools such an item by direct summation in the reference's order (bit-identical to the oracle's
    arithmetic); all other items keep the integer tables"""
    from oracle import cpu as oracle_mod
    B, D, R = 2, 30, 400
    torch.manual_seed(5)
    feat = torch.randn(B, D * 49, 38, 63, device="cuda")
    feat[0, 7 * 49 + 3, 11, 20] = 1.0e9                  # one cell 1e9 x the rest of its plane
    feat[1, 2 * 49 + 10, 5, 5] = -3.0e7
    rois_np = common.make_rois(R, B, seed=21)
    rois = cu(rois_np)
    top, _ = ops.psroi_forward(feat, rois, 7, 7, 1.0 / 16.0, 7, D)
    want, _ = oracle_mod.psroi_forward(feat.cpu().numpy(), rois_np, 1.0 / 16.0, 7, 7, 7, D)
    want = torch.from_numpy(want).cuda()
    # elementwise: bins that do not contain the outlier keep their own relative accuracy
    err = (top - want).abs()
    tol = 1e-5 * want.abs() + 2e-6
    assert bool((err <= tol).all()), (float(err.max()), int((err > tol).sum()))
    # the guarded items are bit-identical to the reference arithmetic
    assert torch.equal(top[:R, 7, 0, 3], want[:R, 7, 0, 3])
    feat[0, 3 * 49 + 1, 0, 0] = float("nan")
    top, _ = ops.psroi_forward(feat, rois, 7, 7, 1.0 / 16.0, 7, D)
    first = top[:R, 3, 0, 1]                              # bins of that plane covering cell (0, 0) are NaN, like the reference
    bins = ops.psroi_bins(rois, 7, 7, 1.0 / 16.0, 38, 63)[:R, 0, 1]
    covers = (bins[:, 0] == 0) & (bins[:, 2] == 0) & (bins[:, 1] > 0) & (bins[:, 3] > 0)
    assert bool(torch.isnan(first[covers]).all()) and not bool(torch.isnan(first[~covers]).any())
    assert not bool(torch.isnan(top[:, 4]).any())


def test_correlation_backward_legacy_launcher_symbol():
    """Correlation_backward_cuda_kernel with the reference's own 40-argument order (correlation_cuda_kernel.h:47-88)
    against the reference kernel itself compiled unmodified (oracle/_ref), for the D&T stride-1 geometry where the
    reference backward is the adjoint of its forward."""
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs /root/reference at build time)")
    g = torch.Generator(device="cuda").manual_seed(9)
    B, Cc, H, W, pad, k, md, s1, s2 = 2, 64, 20, 30, 8, 1, 8, 1, 1
    a = torch.randn(B, Cc, H, W, device="cuda", generator=g)
    b = torch.randn(B, Cc, H, W, device="cuda", generator=g)
    gout = torch.randn(B, 289, H, W, device="cuda", generator=g)
    want1, want2 = ref_cuda.correlation_backward(a, b, gout, pad, k, md, s1, s2)
    g1, g2 = torch.full_like(a, float("nan")), torch.full_like(a, float("nan"))      # (the replacement needs no pre-zeroing)
    ok = lib().Correlation_backward_cuda_kernel(
        gout.data_ptr(), *gout.shape, *gout.stride(), a.data_ptr(), Cc, H, W, *a.stride(), b.data_ptr(), *b.stride(),
        g1.data_ptr(), *g1.stride(), g2.data_ptr(), Cc, *g2.stride(), None, None, pad, k, md, s1, s2, 1,
        torch.cuda.current_stream().cuda_stream)
    assert ok == 1
    torch.cuda.synchronize()
    for got, want in ((g1, want1), (g2, want2)):
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max())
    # non-contiguous strides are refused with a message
    ok = lib().Correlation_backward_cuda_kernel(
        gout.data_ptr(), *gout.shape, *gout.stride(), a.data_ptr(), Cc, H, W, 1, 2, 3, 4, b.data_ptr(), *b.stride(),
        g1.data_ptr(), *g1.stride(), g2.data_ptr(), Cc, *g2.stride(), None, None, pad, k, md, s1, s2, 1,
        torch.cuda.current_stream().cuda_stream)
    assert ok == 0 and b"contiguous" in lib().d2t_last_error()


@pytest.mark.parametrize("n_total,n_take,B", [(28728, 6000, 4), (28728, 12000, 2), (28728, 300, 3), (7296, 6000, 2), (500, 500, 2),
                                              (1000, 1, 2), (32768, 16384, 1)])
def test_proposal_topk_gather_matches_stable_sort(n_total, n_take, B):
    """d2t_proposal_topk_gather (radix select + compaction + bitonic sort + gather, one CTA per image) against what it
    replaces: torch.sort(descending, stable) over all scores, the first n_take, d2t_proposal_gather -- bit for bit, with
    heavy ties (scores on a coarse lattice: the tie order is the index order), zeros of both signs, and a few NaN / inf."""
    from d2t_b200._lib import lib
    assert lib().d2t_proposal_topk_supported(n_total, n_take) == 1 and lib().d2t_proposal_topk_supported(100, 200) == 0
    assert lib().d2t_proposal_topk_supported(40000, 6000) == 0          # keys live in registers: n_total <= 32768
    g = torch.Generator(device="cuda").manual_seed(70 + n_take)
    for kind in ("smooth", "ties", "special"):
        scores = torch.rand(B, n_total, device="cuda", generator=g)
        if kind == "ties":
            scores = (scores * 50).round() / 50
        if kind == "special":
            scores = (scores * 20).round() / 20 - 0.5
            scores[:, 3::97] = 0.0
            scores[:, 5::89] = -0.0
            scores[0, 11] = float("nan")
            scores[0, n_total - 1] = float("nan")
            scores[B - 1, 7] = float("inf")
            scores[B - 1, 8] = float("-inf")
        boxes = torch.rand(B, n_total, 4, device="cuda", generator=g) * 100
        order = torch.sort(scores, dim=1, descending=True, stable=True)[1]
        want = ops.proposal_gather(boxes, scores, order, n_take)
        for split in (False, True):          # one launch (bitonic sort in one CTA) / two launches (device-wide rank sort)
            got = ops.proposal_topk_gather(boxes, scores, n_take, split=split)
            torch.cuda.synchronize()
            same = (got == want) | (torch.isnan(got) & torch.isnan(want))
            assert bool(same.all()), (kind, split, n_total, n_take, int((~same).sum()))


def test_proposals_same_with_hand_written_topk():
    """ops.proposals with the one-launch select + sort + gather kernel == with torch.sort + d2t_proposal_gather"""
    prob, deltas, im_info = common.make_rpn_inputs(B=2, H=38, W=63, seed=33, im_h=600, im_w=1000)
    from model.rpn.generate_anchors import generate_anchors
    from model.utils.config import cfg
    anchors = torch.from_numpy(generate_anchors(scales=np.array(cfg.ANCHOR_SCALES), ratios=np.array(cfg.ANCHOR_RATIOS))).float().cuda()
    outs = []
    saved = ops.HAND_WRITTEN_TOPK, ops.SPLIT_TOPK
    try:
        for flag, split in ((False, False), (True, False), (False, True)):
            ops.HAND_WRITTEN_TOPK, ops.SPLIT_TOPK = flag, split
            outs.append(ops.proposals(anchors, cu(deltas), cu(prob), cu(im_info), 16, 6000, 300, 0.7))
    finally:
        ops.HAND_WRITTEN_TOPK, ops.SPLIT_TOPK = saved
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
