"""PSRoI forward: the fp64 summed-area-table kernel (D2T_PSROI_INT=0) against the integer-table variants, BASELINE config 5.
  D2T_PSROI_INT unset  the library's own choice (integer tables, mode 4, when there is more than one item per SM)
  D2T_PSROI_INT=1  in-place int32 tables, one 1024-thread CTA per SM, triple-buffered
  D2T_PSROI_INT=2  same tables, one item per CTA, three CTAs per SM, lane -> (roi, pw) lookups
  D2T_PSROI_INT=3  same, lane -> roi lookups with the outputs transposed through shared memory
  D2T_PSROI_INT=4  same as 3 with thread-per-row norm / quantise / row scan
  D2T_PSROI_THREADS=256|384  CTA width of modes 2 / 3
Prints one JSON line per (mode, shape): time, GB/s against the algorithmic bytes, max |difference| to the fp64 kernel.
(The library reads the variables at every launch, so one process can switch between them.)  Run under `timeout`."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
ARGS = sys.argv[1:]
sys.argv = ["bench.py"]
import bench
import common
from d2t_b200._lib import lib

torch.cuda.set_device(0)
flush = torch.zeros(128 * 1024 * 1024, device="cuda")
HBM = bench.peaks()[0]
MODES = [("fp64", "0", None), ("auto", None, None), ("int1", "1", None), ("int2_256", "2", "256"), ("int2_384", "2", "384"),
         ("int3_256", "3", "256"), ("int3_384", "3", "384"), ("int4_256", "4", "256"), ("int4_288", "4", "288"), ("int4_320", "4", "320"),
         ("int4_384", "4", "384")]


def set_mode(mode, threads):
    for k, v in (("D2T_PSROI_INT", mode), ("D2T_PSROI_THREADS", threads)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def run_shape(B, D, R, shuffle=False, modes=MODES, iters=20, want_mapping=False):
    torch.manual_seed(20)
    C = D * 49
    feat = torch.randn(B, C, 38, 63, device="cuda")
    rois = torch.from_numpy(common.make_rois(R, B, seed=21, shuffle=shuffle)).cuda()
    n = rois.size(0)
    ws = torch.empty(lib().d2t_psroi_workspace_bytes(n, B, 7, 7), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    alg = 4.0 * (C * 2394 + 5 * R + R * D * 49) * B
    ref = ref_map = None
    for name, mode, threads in modes:
        set_mode(mode, threads)
        top = torch.full((n, D, 7, 7), float("nan"), device="cuda")
        mp = torch.full((n, D, 7, 7), -7, dtype=torch.int32, device="cuda") if want_mapping else None

        def call():
            ok = lib().d2t_psroi_forward(feat.data_ptr(), B, C, 38, 63, rois.data_ptr(), n, 1 / 16., 7, 7, 7, D,
                                         top.data_ptr(), mp.data_ptr() if mp is not None else None, ws.data_ptr(),
                                         ws.numel(), st)
            assert ok == 1, lib().d2t_last_error()
        ms = bench.time_kernel(call, iters, flush)
        torch.cuda.synchronize()
        if ref is None:
            ref, ref_map = top.clone(), (mp.clone() if mp is not None else None)
        rec = {"mode": name, "B": B, "D": D, "R_per_img": R, "shuffled": shuffle, "us": ms * 1e3,
               "gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / HBM,
               "max_abs_diff_vs_fp64": float((top - ref).abs().max()), "nan": bool(torch.isnan(top).any()),
               "viol_rtol1e-5_atol2e-6": int(((top - ref).abs() > 2e-6 + 1e-5 * ref.abs()).sum())}
        if want_mapping:
            rec["mapping_equal"] = bool(torch.equal(mp, ref_map))
        print(json.dumps(rec), flush=True)
    set_mode(None, None)


if __name__ == "__main__":
    quick = bool(ARGS) and ARGS[0] == "quick"
    fast = [m for m in MODES if m[0] in (("fp64", "auto", "int4_288") if quick else ("fp64", "auto", "int4_256", "int4_320", "int4_384"))]
    run_shape(2, 30, 2000, modes=fast if quick else MODES)     # BASELINE config 5 at B = 2 (bench.py's shape)
    run_shape(2, 30, 2000, shuffle=True, modes=fast, iters=5, want_mapping=True)   # unsorted rois + mapping output
    run_shape(3, 4, 77, modes=fast, iters=5, want_mapping=True)     # ragged: R not a multiple of 32, D = 4
    run_shape(4, 31, 300, modes=fast, iters=10)                # the model's cls head: 4 frames x 300 rois, D = 31
    for B in (1, 8, 32):                                       # batch sweep of config 5
        run_shape(B, 30, 2000, modes=fast, iters=10)
