"""Loader for libd2t_b200.so, the C-ABI library declared in include/d2t_b200.h.

There is no CPU fallback and no alternative backend: if the shared library has not been
built (``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C pytorch-detect-to-track_b200/csrc``) every operator raises ``D2TLibraryMissing``.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("D2T_B200_LIB") or os.path.join(_HERE, "libd2t_b200.so")   # (override: debug builds only)
CSRC_DIR = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
HEADER = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "d2t_b200.h"))


class D2TLibraryMissing(RuntimeError):
    pass


class D2TError(RuntimeError):
    pass


_i, _f, _p, _sz = C.c_int, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); order and meaning exactly as include/d2t_b200.h
_SIGS = {
    "d2t_version": (C.c_char_p, []),
    "d2t_last_error": (C.c_char_p, []),
    "d2t_device_sm_count": (_i, []),
    # ---- Part 1: reference launcher symbols
    "Correlation_forward_cuda_kernel": (_i, [_p] + [_i] * 8 + [_p] + [_i] * 7 + [_p] + [_i] * 5 + [_p, _p] + [_i] * 6 + [_p]),
    "Correlation_backward_cuda_kernel": (_i, [_p] + [_i] * 8 + [_p] + [_i] * 7 + [_p] + [_i] * 4 + [_p] + [_i] * 4 +
                                         [_p] + [_i] * 5 + [_p, _p] + [_i] * 6 + [_p]),
    "PSROIPoolForwardLauncher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p]),
    "PSROIPoolBackwardLauncher": (_i, [_p, _p, _i, _i, _f, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ROIAlignForwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "d2t_roi_backward_scratch_bytes": (_sz, [_sz]),
    "d2t_roi_align_backward_det": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "d2t_roi_pool_backward_det": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "d2t_roi_crop_backward_det": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "ROIAlignBackwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ROIPoolForwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ROIPoolBackwardLaucher": (_i, [_p, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "BilinearSamplerBHWD_updateOutput_cuda_kernel": (_i, [_i] * 8 + [_p] + [_i] * 4 + [_p] + [_i] * 4 + [_p] + [_i] * 4 + [_p]),
    "BilinearSamplerBHWD_updateGradInput_cuda_kernel": (_i, [_i] * 8 + [_p] + [_i] * 4 + [_p] + [_i] * 4 + [_p] + [_i] * 4 +
                                                        [_p] + [_i] * 4 + [_p] + [_i] * 4 + [_p]),
    "nms_cuda_compute": (None, [_p, _p, _p, _i, _i, _f]),
    # ---- Part 2: stream-ordered surface
    "d2t_nms_workspace_bytes": (_sz, [_i, _i]),
    "d2t_nms_prefix": (_i, [_i, _i]),
    "d2t_nms_set_mode": (_i, [_i]),
    "d2t_nms_launch_count": (_i, [_i, _i]),
    "d2t_nms_batched": (_i, [_p, _p, _i, _i, _i, _f, _i, _p, _i, _p, _p, _sz, _p]),
    "d2t_psroi_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "d2t_psroi_forward": (_i, [_p, _i, _i, _i, _i, _p, _i, _f, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "d2t_psroi_backward": (_i, [_p, _i, _i, _i, _i, _p, _i, _f, _i, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "d2t_psroi_bins": (_i, [_p, _i, _f, _i, _i, _i, _i, _p, _p]),
    "d2t_correlation_shape": (_i, [_i] * 7 + [C.POINTER(_i)]),
    "d2t_correlation_forward": (_i, [_p, _p] + [_i] * 9 + [_p, _p]),
    "d2t_correlation_backward": (_i, [_p, _p, _p] + [_i] * 9 + [_p, _p, _p]),
    "d2t_proposal_decode": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "d2t_proposal_gather": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "d2t_proposal_write_rois": (_i, [_p, _p, _i, _p, _i, _i, _i, _p, _p]),
    "d2t_proposal_topk_supported": (_i, [_i, _i]),
    "d2t_proposal_topk_gather": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "d2t_proposal_topk_scratch_bytes": (_sz, [_i, _i]),
    "d2t_proposal_topk_gather_split": (_i, [_p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    # ---- convolution engine
    "d2t_conv_plan_create": (_p, [_p] * 9),
    "d2t_conv_plan_destroy": (None, [_p]),
    "d2t_conv_plan_set_amax": (_i, [_p, _p, _p]),
    "d2t_conv_scratch_bytes": (_sz, []),
    "d2t_conv_plan_set_scratch": (_i, [_p, _p, _sz]),
    "d2t_conv_plan_set_done": (_i, [_p, _p, _p, _p]),
    "d2t_conv_plan_set_early_weights": (_i, [_p, _i]),
    "d2t_conv_plan_chainable": (_i, [_p]),
    "d2t_conv_chain_bytes": (_sz, [_i]),
    "d2t_conv_chain_create": (_p, [_p, _p, _i, _p, _sz]),
    "d2t_conv_chain_run": (_i, [_p, _p]),
    "d2t_conv_chain_destroy": (None, [_p]),
    "d2t_conv_pack_weights_f16": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "d2t_corr_plan_create": (_p, [_i] * 10 + [_p] * 3 + [_i, _i, _p]),
    "d2t_conv_plan_info": (_i, [_p, C.POINTER(_i)]),
    "d2t_conv_plan_run": (_i, [_p, _p]),
    "d2t_conv_pack_weights": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "d2t_nchw_to_nhwc": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "d2t_nchw_to_nhwc_amax": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "d2t_conv_stem_plan_create": (_p, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _i]),
    "d2t_stem_pack_input": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "d2t_stem_pack_input_amax": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p]),
    "d2t_stem_pack_weights": (_i, [_p, _i, _i, _p, _p, _p]),
    "d2t_nhwc_to_nchw": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "d2t_maxpool3x3s2_nhwc": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    # ---- frame preparation (blob.py / minibatch.py on the device)
    "d2t_frames_resized_shape": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(C.c_double)]),
    "d2t_frames_prep": (_i, [_p, _i, _i, _i, C.POINTER(C.c_double), C.c_double, _i, _i, _i, _p, _i, _i, _i, _p]),
    # ---- training path (backward-data / weight-gradient)
    "d2t_conv_plan_set_mask": (_i, [_p, _p, _i]),
    "d2t_conv_plan_set_weight_amax": (_i, [_p, _p]),
    "d2t_conv_pack_weights_f16_dev": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "d2t_conv_pack_weights_f16_dgrad": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "d2t_conv_repack_item_bytes": (_sz, []),
    "d2t_conv_repack_item_blocks": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "d2t_conv_repack_many": (_i, [_p, _i, _i, _p, _p]),
    "d2t_upsample2_add_mask": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "d2t_wgrad_pack_input": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "d2t_wgrad_pack_grad": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "d2t_wgrad_plan_create": (_p, [_i] * 13 + [_p] * 7),
    "d2t_wgrad_partials_bytes": (_sz, []),
    "d2t_wgrad_plan_set_partials": (_i, [_p, _p, _sz]),
    "d2t_corrb_pack_band": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "d2t_corrb_pack_other": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "d2t_corrb_plan_create": (_p, [_i] * 5 + [_p] * 4 + [_i] + [_p] * 4 + [_i, _i]),
    "d2t_psroi_set_mode": (_i, [_i, _i]),
    "d2t_psroi_vote_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "d2t_psroi_vote_forward": (_i, [_p, _i, _i, _i, _i, _p, _i, _f, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
}

_OPTIONAL = set()


class ConvDesc(C.Structure):
    """struct d2t_conv_desc (include/d2t_b200.h)"""
    _fields_ = [(n, C.c_int) for n in ("N", "H", "W", "Cin", "in_cstride", "Cout", "R", "S", "stride", "pad", "dil",
                                       "passes", "relu", "out_cstride", "out_coffset", "res_cstride", "w_exp")]

_lib = None


def build(verbose=False):
    """Compile libd2t_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise D2TLibraryMissing(
                "%s is missing: build it with `make -C %s` (there is no CPU or PyTorch fallback)" % (SO_PATH, CSRC_DIR))
        handle = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                if name in _OPTIONAL:
                    continue
                raise D2TLibraryMissing("%s does not export %s: rebuild it" % (SO_PATH, name))
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def declared_symbols():
    """Function names declared in include/d2t_b200.h (used by the export test)."""
    import re
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)) - {"defined", "push", "visibility"})


def check(ok, what):
    if ok != 1:
        raise D2TError("%s failed: %s" % (what, lib().d2t_last_error().decode() or "unknown error"))
