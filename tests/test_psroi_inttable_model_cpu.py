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
