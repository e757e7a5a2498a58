"""CPU model of the arithmetic of the default PSRoI forward kernel (csrc/psroi.cu, psroi_fwd_isat_mc): per-plane
power-of-two fixed point, in-place int32 2-D inclusive prefix sums with wrap-around, four-corner window sums with the
"row / column -1 contributes 0" masks.  It pins the two facts the kernel's precision claim rests on, without a GPU:
  * window sums are exact in integers even when the running prefix wraps around int32;
  * the pooled value differs from the exact bin mean (the oracle's bins, float64 arithmetic) by at most
    2^-30 * L1(plane) -- the bound quoted in DESIGN.md / include/d2t_b200.h and asserted on the GPU in
    tests/test_ops_gpu.py::test_psroi_integer_tables_against_exact_tables.
The bins come from the C oracle (bit-exact with the reference kernel, tests/test_oracle_golden.py)."""
import numpy as np

import common


def int_tables(plane):
    """plane [H, W] float32 -> (int32 inclusive 2-D prefix sums with wrap-around, k) as the kernel builds them"""
    l1 = np.float32(np.abs(plane).astype(np.float32).sum(dtype=np.float32))
    eb = (l1.view(np.uint32) >> 23) & 0xff
    k = int(156 - int(eb)) if 0 < eb < 255 else 0
    k = max(-96, min(120, k))
    q = np.rint(plane.astype(np.float32) * np.float32(2.0 ** k)).astype(np.int64)
    assert np.abs(q).sum() < 2 ** 31
    t = np.cumsum(np.cumsum(q, axis=1), axis=0)
    return t.astype(np.int64), k


def lookup(t32, hs, he, ws, we):
    """four-corner sum in wrap-around int32 arithmetic with the kernel's masks"""
    def at(r, c):
        return np.int64(t32[r, c])
    a11 = at(he - 1, we - 1)
    a01 = at(max(hs - 1, 0), we - 1) if hs > 0 else 0
    a10 = at(he - 1, max(ws - 1, 0)) if ws > 0 else 0
    a00 = at(max(hs - 1, 0), max(ws - 1, 0)) if (hs > 0 and ws > 0) else 0
    s = (a11 - a10) - (a01 - a00)
    return int(((int(s) + 2 ** 31) % 2 ** 32) - 2 ** 31)          # the int32 result of the kernel's subtractions


def test_integer_table_window_sums_exact_and_bounded(oracle):
    rng = np.random.RandomState(3)
    G, D, H, W = 7, 2, 38, 63
    for amp, offset in ((1.0, 0.0), (250.0, 0.0), (1.0, 4.0), (1e-3, 0.0)):
        feat = (rng.randn(1, D * G * G, H, W) * amp + offset).astype(np.float32)
        rois = common.make_rois(60, 1, seed=11)
        _, _, bins = oracle.psroi_forward(feat, rois, 1 / 16., G, G, G, D, want_bins=True)
        worst, worst_bound = 0.0, 0.0
        for c in range(0, D * G * G, 5):
            plane = feat[0, c]
            t, k = int_tables(plane)
            wrapped = ((t + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int64)       # what an int32 table holds
            ctop, rem = divmod(c, G * G)
            ph, pw = divmod(rem, G)
            l1 = float(np.abs(plane.astype(np.float64)).sum())
            for n in range(rois.shape[0]):
                hs, he, ws, we = [int(v) for v in bins[n, ph, pw]]
                if he <= hs or we <= ws:
                    continue
                s = lookup(wrapped, hs, he, ws, we)
                q_exact = int(np.rint(plane[hs:he, ws:we].astype(np.float32) * np.float32(2.0 ** k)).astype(np.int64).sum())
                assert s == q_exact                                              # exact integer window sum
                area = (he - hs) * (we - ws)
                got = s * 2.0 ** -k / area
                want = float(plane[hs:he, ws:we].astype(np.float64).mean())
                worst = max(worst, abs(got - want))
                worst_bound = max(worst_bound, l1 * 2.0 ** -30)
                assert abs(got - want) <= l1 * 2.0 ** -30, (amp, offset, c, n)
        assert worst > 0.0 and worst <= worst_bound


def test_integer_table_wraparound_is_harmless():
    """A running prefix that leaves int32 does not matter as long as the window sum itself fits: differences of
    wrapped values are the differences of the true values modulo 2^32."""
    rng = np.random.RandomState(5)
    q = rng.randint(-2 ** 20, 2 ** 20, size=(38, 63)).astype(np.int64)
    t = np.cumsum(np.cumsum(q, axis=1), axis=0) + (2 ** 33 + 2 ** 31 - 7)          # far outside int32
    wrapped = (t + 2 ** 31) % 2 ** 32 - 2 ** 31
    for _ in range(200):
        hs, ws = rng.randint(1, 30), rng.randint(1, 50)
        he, we = hs + rng.randint(1, 8), ws + rng.randint(1, 12)
        assert lookup(wrapped, hs, he, ws, we) == int(q[hs:he, ws:we].sum())


def test_integer_difference_table_backward_model(oracle):
    """CPU model of psroi_bwd_isat_mc (the opt-in integer-table backward): per item one scale 2^k from sum |dv|, the four
    corner updates of every bin on a [H+1][Wp] int32 table, row scan over x < W, column scan over h < H, planes * 2^-k.
    Against the C oracle's backward (the reference's per-cell accumulation): within n 2^-30 sum|dv| for n covering bins."""
    rng = np.random.RandomState(7)
    G, D, H, W, B = 7, 2, 38, 63, 2
    feat_shape = (B, D * G * G, H, W)
    rois = common.make_rois(40, B, seed=13, shuffle=True)
    gtop = rng.randn(rois.shape[0], D, G, G).astype(np.float32)
    want = oracle.psroi_backward(gtop, rois, feat_shape, 1 / 16., G, G, G, D)
    _, _, bins = oracle.psroi_forward(np.zeros(feat_shape, np.float32), rois, 1 / 16., G, G, G, D, want_bins=True)
    Wp = (W + 1) | 1
    got = np.zeros(feat_shape, np.float32)
    for b in range(B):
        mine = np.nonzero(rois[:, 0].astype(int) == b)[0]
        for ctop in range(D):
            for ph in range(G):
                dvs = []
                for n in mine:
                    for pw in range(G):
                        hs, he, ws, we = [int(v) for v in bins[n, ph, pw]]
                        if he > hs and we > ws:
                            dv = np.float32(gtop[n, ctop, ph, pw]) / np.float32((he - hs) * (we - ws))
                            dvs.append((pw, hs, he, ws, we, np.float32(dv)))
                tot = np.float32(sum(abs(float(d[5])) for d in dvs))
                eb = (np.float32(tot).view(np.uint32) >> 23) & 0xff
                k = int(156 - int(eb)) if 0 < eb < 255 else 0
                k = max(-96, min(120, k))
                T = np.zeros((G, H + 1, Wp), np.int64)
                for pw, hs, he, ws, we, dv in dvs:
                    q = int(np.rint(np.float32(dv * np.float32(2.0 ** k))))
                    T[pw, hs, ws] += q
                    T[pw, hs, we] -= q
                    T[pw, he, ws] -= q
                    T[pw, he, we] += q
                assert sum(abs(int(np.rint(np.float32(d[5] * np.float32(2.0 ** k))))) for d in dvs) < 2 ** 31
                T = ((T + 2 ** 31) % 2 ** 32) - 2 ** 31                       # int32 storage
                P = np.cumsum(np.cumsum(T[:, :H, :W], axis=2), axis=1)       # row scan over x < W, column scan over h < H
                P = ((P + 2 ** 31) % 2 ** 32) - 2 ** 31
                c0 = (ctop * G + ph) * G
                got[b, c0:c0 + G] = (P.astype(np.float32) * np.float32(2.0 ** -k))
                bound = max(len(dvs), 1) * 2.0 ** -30 * max(float(tot), 1e-30) + 1e-6 * float(np.abs(want[b, c0:c0 + G]).max())
                assert float(np.abs(got[b, c0:c0 + G] - want[b, c0:c0 + G]).max()) <= bound
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
