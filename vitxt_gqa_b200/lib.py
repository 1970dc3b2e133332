"""ctypes binding of libt2s_sm100.so (include/t2s_b200.h).

The shared library is the product; this module only loads it, declares the
argument types and turns non-zero return codes into exceptions.  There is no
fallback: if the library is missing or fails to load, importing callers get a
RuntimeError that says how to build it.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libt2s_sm100.so")
CSRC = os.path.join(_HERE, "csrc")

GEMM_GELU, GEMM_OUT_F32, GEMM_RES_F32, GEMM_OUT_SPLIT = 1, 2, 4, 8
GEMM_DGELU = 16
GEMM_SM_CAP_SHIFT = 8     # bits 8..15 of the GEMM flags: cap on the persistent grid

_p, _i, _ll, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
_d = ctypes.c_double
_ull, _u = ctypes.c_ulonglong, ctypes.c_uint

# name -> argtypes, in the order of include/t2s_b200.h
SIGNATURES = {
    "t2s_gemm_bf16": [_p, _ll, _p, _ll, _p, _p, _ll, _p, _ll, _i, _i, _i, _i, _i, _p],
    "t2s_gemm_bf16x3": [_p, _ll, _p, _ll, _p, _p, _ll, _p, _ll, _i, _i, _i, _i, _i, _p],
    "t2s_gemm_wgrad_bf16": [_p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p],
    "t2s_split_bf16": [_p, _ll, _i, _i, _i, _p, _ll, _i, _i, _i, _p],
    "t2s_gemm_f32": [_p, _ll, _p, _ll, _p, _p, _ll, _p, _ll, _i, _i, _i, _i, _i, _i, _i, _p],
    "t2s_attn_f32": [_p, _ll, _i, _i, _i, _i, _p, _p, _i, _p, _ll, _p, _ll, _p],
    "t2s_attn_x3": [_p, _ll, _i, _i, _i, _i, _i, _p, _p, _i, _p, _ll, _p],
    "t2s_attn_tc": [_p, _ll, _i, _i, _i, _i, _i, _p, _p, _i, _p, _ll, _p],
    "t2s_attn_bf16": [_p, _ll, _i, _i, _i, _i, _p, _p, _i, _p, _ll, _p],
    "t2s_attn_dec": [_p, _ll, _i, _p, _ll, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _ll, _p],
    "t2s_bert_embed_ln": [_p, _i, _i, _i, _p, _p, _p, _p, _p, _f, _p, _ll, _p],
    "t2s_feat_concat": [_p, _i, _p, _i, _p, _p, _p, _p, _i, _i, _p, _ll, _i, _p, _ll, _p],
    "t2s_add_ln": [_p, _i, _ll, _p, _i, _ll, _p, _p, _f, _i, _i, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _p],
    "t2s_add_ln_split": [_p, _i, _ll, _p, _i, _ll, _p, _p, _f, _i, _i, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _p],
    "t2s_ocr_finish": [_p, _ll, _p, _p, _p, _p, _p, _p, _p, _f, _i, _i, _p, _ll, _i, _i, _i, _p],
    "t2s_prev_embed": [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _ll, _ll, _p, _p, _p, _p, _p, _p, _p, _p, _f,
                       _p, _p, _ll, _i, _p],
    "t2s_cast_rows_bf16": [_p, _ll, _i, _i, _p, _ll, _i, _i, _i, _p],
    "t2s_mask_prep": [_p, _p, _p, _i, _i, _i, _i, _p, _p],
    "t2s_build_keys": [_p, _i, _i, _p, _p, _i, _p],
    "t2s_question_pool": [_p, _i, _i, _i, _p, _p, _p, _i, _p, _p],
    "t2s_sim_scores": [_p, _p, _ll, _ll, _i, _i, _i, _i, _p, _p],
    "t2s_temporal_select": [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "t2s_spatial_select": [_p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p],
    "t2s_middle_frame_slots": [_p, _p, _i, _i, _p, _p],
    "t2s_frame_slots": [_p, _i, _p, _i, _i, _p, _p],
    "t2s_frames_from_ocr": [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p],
    "t2s_ptr_score": [_p, _ll, _i, _i, _i, _i, _p, _ll, _ll, _i, _i, _p, _ll, _p, _ll, _i, _p],
    "t2s_argmax_feedback": [_p, _ll, _i, _i, _i, _i, _i, _p, _i, _p, _p],
    "t2s_pos_bce_loss": [_p, _p, _p, _i, _i, _i, _p, _p, _p],
    "t2s_info_nce_loss": [_p, _p, _p, _i, _i, _i, _f, _p, _p, _p],
    # K8 training step
    "t2s_ln_bwd": [_p, _i, _ll, _p, _i, _ll, _i, _i, _i, _p, _p, _f, _i, _i, _i, _p, _i, _ll, _p, _p, _p, _p],
    "t2s_colsum": [_p, _i, _ll, _i, _i, _p, _p],
    "t2s_rows_add": [_p, _p, _p, _ll, _i, _i, _p, _ll, _i, _i, _i, _i, _p],
    "t2s_gelu_rows": [_p, _i, _ll, _i, _i, _p, _ll, _i, _p],
    "t2s_embed_scatter_add": [_p, _i, _ll, _i, _i, _p, _i, _ll, _p, _ll, _p],
    "t2s_ptr_score_bwd": [_p, _ll, _i, _i, _i, _p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _p, _ll, _ll, _p],
    "t2s_prev_embed_bwd": [_p, _ll, _p, _i, _i, _i, _i, _i, _p, _p, _ll, _ll, _p, _p, _p, _p, _p, _f,
                           _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p],
    "t2s_ocr_finish_bwd": [_p, _ll, _p, _p, _p, _p, _p, _f, _i, _i, _p, _ll, _i, _i, _i, _p, _ll, _p, _ll,
                           _p, _p, _p, _p, _p, _p, _p, _p],
    "t2s_bert_embed_bwd": [_p, _ll, _p, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p],
    "t2s_attn_bwd": [_p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _i,
                     _p, _p, _i, _i, _p, _p],
    "t2s_nce_rowstats": [_p, _p, _p, _i, _i, _p, _p],
    "t2s_pos_bce_loss_bwd": [_p, _p, _p, _i, _i, _i, _p, _p, _i, _p],
    "t2s_info_nce_loss_bwd": [_p, _p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p, _i, _p],
    "t2s_sumsq": [_p, _ll, _p, _p, _p],
    "t2s_repack_weights": [_p, _p, _i, _i, _p],
    # input featurisation
    "t2s_phoc_build": [_p, _p, _i, _i, _p, _ll, _p],
    "t2s_phoc_build_fixed": [_p, _i, _i, _p, _ll, _p],
    "t2s_pack_ocr_frames": [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    # evaluation step
    "t2s_answer_decode": [_p, _ll, _i, _i, _i, _i, _i, _p, _p, _p],
    "t2s_ground_metrics": [_p, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _d, _d, _p, _p, _p, _p, _p, _p, _p],
    "t2s_adam_step": [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _i, _p, _f, _f, _p],
    # dropout of the training step (masks recomputed from (p, seed, site))
    "t2s_add_ln_dropout": [_p, _i, _ll, _p, _i, _ll, _p, _p, _f, _i, _i, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p,
                           _f, _ull, _u, _p],
    "t2s_dropout_rows": [_p, _i, _ll, _i, _i, _i, _i, _i, _f, _ull, _u, _p],
    "t2s_dropout_mask": [_p, _i, _i, _i, _i, _i, _f, _ull, _u, _p],
    "t2s_ln_bwd_dropout": [_p, _i, _ll, _p, _i, _ll, _i, _i, _i, _p, _p, _f, _i, _i, _i, _p, _i, _ll, _p, _p, _p, _p,
                           _f, _ull, _u, _p],
    "t2s_attn_tc_dropout": [_p, _ll, _i, _i, _i, _i, _i, _p, _p, _i, _p, _ll, _f, _ull, _u, _p, _i, _p],
    "t2s_attn_dec_dropout": [_p, _ll, _i, _p, _ll, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _ll, _f, _ull, _u, _p, _p],
    "t2s_attn_bwd_dropout": [_p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _i,
                             _p, _p, _i, _i, _p, _f, _ull, _u, _p, _p],
}
PLAIN = {"t2s_abi_version": (_i, []), "t2s_last_error": (ctypes.c_char_p, []),
         "t2s_loss_workspace_bytes": (_ll, [_i, _i]), "t2s_loss_bwd_workspace_bytes": (_ll, [_i, _i]),
         "t2s_attn_bwd_workspace_bytes": (_ll, [_i, _i, _i, _i]),
         "t2s_tmap_cache_stats": (_ll, [_i])}

EXPORTED_SYMBOLS = sorted(list(SIGNATURES) + list(PLAIN))


class T2SLibraryError(RuntimeError):
    pass


def build_library(verbose=False):
    """Compile csrc/*.cu for sm_100a into libt2s_sm100.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise T2SLibraryError("building libt2s_sm100.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise T2SLibraryError(
                "libt2s_sm100.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C vitxt_gqa_b200/csrc`. There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.launches = 0
        self.timing = None
        # host-side default for the SM cap bits of the tcgen05 GEMM flags (0 = none): the pipelined eval forward
        # (model.submit) sets it so that every throughput GEMM leaves SMs to the decode stream
        self.gemm_cap = 0
        for name, (res, args) in PLAIN.items():
            fn = getattr(self.cdll, name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name[4:], fn)
        if self.cdll.t2s_abi_version() != 1:
            raise T2SLibraryError("libt2s_sm100.so ABI version mismatch")
        for name, args in SIGNATURES.items():
            fn = getattr(self.cdll, name)
            fn.restype, fn.argtypes = _i, args
            setattr(self, name[4:], self._checked(name, fn))

    def _checked(self, name, fn):
        capped = name in ("t2s_gemm_bf16", "t2s_gemm_bf16x3")     # flags = argument 12 (include/t2s_b200.h)

        def call(*args):
            if capped and self.gemm_cap and not (args[12] >> GEMM_SM_CAP_SHIFT) & 0xff:
                args = args[:12] + (args[12] | ((self.gemm_cap & 0xff) << GEMM_SM_CAP_SHIFT),) + args[13:]
            rec = self.timing
            if rec is not None:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()          # torch's current stream == the stream every caller passes down
            rc = fn(*args)
            if rc != 0:
                raise T2SLibraryError("%s failed (rc=%d): %s" % (name, rc, self.cdll.t2s_last_error().decode()))
            self.launches += 1
            if rec is not None:
                e1.record()
                rec.append((name, args, e0, e1))
        call.__name__ = name
        return call

    def start_timing(self):
        """Measurement aid for bench.py: bracket every launch with CUDA events on the launching stream."""
        self.timing = []

    def stop_timing(self):
        """-> [(entry point, args, milliseconds)] after synchronising."""
        import torch
        torch.cuda.synchronize()
        rec, self.timing = self.timing or [], None
        return [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in rec]


_LIB = None


def get_lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB
