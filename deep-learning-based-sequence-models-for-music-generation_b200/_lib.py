"""ctypes binding of the C-ABI in include/mamba_b200.h.  Fails loudly: no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libmamba_b200.so"

MAMBA_F32, MAMBA_BF16 = 0, 1
FLAG_HAS_Z, FLAG_DELTA_SOFTPLUS, FLAG_HAS_DELTA_BIAS, FLAG_HAS_D, FLAG_A_IS_LOG = 1, 2, 4, 8, 16

EXPORTS = [
    "mamba_abi_version", "mamba_last_error", "mamba_launch_count",
    "mamba_scan_fwd", "mamba_scan_ckpt_elems", "mamba_scan_bwd", "mamba_scan_bwd_workspace_bytes",
    "mamba_conv1d_silu_fwd", "mamba_conv1d_silu_bwd", "mamba_conv1d_bwd_workspace_bytes",
    "mamba_conv_step", "mamba_ssm_step", "mamba_linear_step", "mamba_fused_linear_step", "mamba_sample_step",
    "mamba_decode_token", "mamba_decode_token_scratch_bytes", "mamba_decode_token_barrier_bytes",
    "mamba_rmsnorm_fwd", "mamba_rmsnorm_bwd", "mamba_rmsnorm_bwd_workspace_bytes",
    "mamba_filtered_ce_fwd", "mamba_filtered_ce_bwd", "mamba_filtered_ce_workspace_bytes",
]

i32, i64, vp, fp, sz = C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t


def _t(name):  # tensor triple: pointer, batch stride, seq stride
    return [(name, vp), (name + "_bs", i64), (name + "_ls", i64)]


class ScanFwdArgs(C.Structure):
    _fields_ = ([("struct_size", i32), ("dtype", i32), ("batch", i32), ("seqlen", i32), ("dim", i32), ("dstate", i32),
                 ("chunk", i32), ("flags", i32), ("variant", i32), ("reserved", i32)]
                + _t("u") + _t("delta") + [("A", fp)] + _t("B") + _t("C") + [("D", fp)] + _t("z")
                + [("delta_bias", fp)] + _t("out") + [("ckpt", fp), ("h_last", fp), ("h_init", fp)] + _t("y_pre"))


class ScanBwdArgs(C.Structure):
    _fields_ = ([("struct_size", i32), ("dtype", i32), ("batch", i32), ("seqlen", i32), ("dim", i32), ("dstate", i32),
                 ("chunk", i32), ("flags", i32), ("variant", i32), ("reserved", i32)]
                + _t("u") + _t("delta") + [("A", fp)] + _t("B") + _t("C") + [("D", fp)] + _t("z")
                + [("delta_bias", fp)] + _t("dout") + [("ckpt", fp)]
                + _t("du") + _t("ddelta") + _t("dz") + _t("dB") + _t("dC")
                + [("dA", fp), ("dD", fp), ("ddelta_bias", fp), ("workspace", vp), ("workspace_bytes", sz)]
                + _t("y_pre"))


class ConvArgs(C.Structure):
    _fields_ = ([("struct_size", i32), ("dtype", i32), ("batch", i32), ("seqlen", i32), ("dim", i32), ("width", i32)]
                + _t("x") + [("weight", fp), ("bias", fp)] + _t("out") + [("final_state", vp)]
                + _t("dout") + _t("dx") + [("dweight", fp), ("dbias", fp), ("workspace", vp), ("workspace_bytes", sz)])


class StepArgs(C.Structure):
    _fields_ = [("struct_size", i32), ("dtype", i32), ("batch", i32), ("dim", i32), ("dstate", i32), ("width", i32),
                ("dt_rank", i32), ("flags", i32),
                ("x", vp), ("x_bs", i64), ("conv_state", vp), ("conv_weight", fp), ("conv_bias", fp),
                ("xc", vp), ("xc_bs", i64),
                ("dt_in", vp), ("dt_in_bs", i64), ("Bv", vp), ("Bv_bs", i64), ("Cv", vp), ("Cv_bs", i64),
                ("dt_weight", fp), ("dt_bias", fp), ("A", fp), ("D", fp), ("z", vp), ("z_bs", i64),
                ("ssm_state", fp), ("y", vp), ("y_bs", i64)]


class NormArgs(C.Structure):
    _fields_ = [("struct_size", i32), ("dtype", i32), ("resid_dtype", i32), ("dim", i32), ("rows", i64),
                ("eps", C.c_float), ("reserved", i32),
                ("x", vp), ("residual", vp), ("weight", fp), ("y", vp), ("resid_out", vp), ("rstd", fp),
                ("dy", vp), ("dresid_in", vp), ("dx", vp), ("dresid_out", vp), ("dweight", fp), ("workspace", vp),
                ("workspace_bytes", sz)]


class LinearStepArgs(C.Structure):
    _fields_ = [("struct_size", i32), ("dtype", i32), ("w_dtype", i32), ("batch", i32), ("in_features", i32),
                ("out_features", i32), ("x", vp), ("x_bs", i64), ("weight", vp), ("bias", vp), ("y", vp), ("y_bs", i64)]


class FusedLinearStepArgs(C.Structure):
    _fields_ = [("struct_size", i32), ("dtype", i32), ("w_dtype", i32), ("batch", i32), ("in_features", i32),
                ("out_features", i32), ("x", vp), ("x_bs", i64), ("weight", vp), ("bias", vp), ("y", vp), ("y_bs", i64),
                ("norm_weight", fp), ("eps", C.c_float), ("conv_dim", i32),
                ("residual_in", fp), ("residual_in_bs", i64), ("residual_out", fp), ("residual_out_bs", i64),
                ("conv_width", i32), ("reserved", i32), ("conv_state", vp), ("conv_weight", fp), ("conv_bias", fp),
                ("conv_out", vp), ("conv_out_bs", i64)]


class SampleStepArgs(C.Structure):
    _fields_ = [("struct_size", i32), ("mode", i32), ("batch", i32), ("vocab", i32), ("bucket_bounds", i32 * 4),
                ("class_bounds", i32 * 4), ("pen_rule", i32 * 5), ("prompt_len", i32), ("time_budget", i32),
                ("reserved", i32), ("pen_table", fp),
                ("logits", fp), ("logits_bs", i64), ("lse", fp), ("dist", fp), ("counts", vp),
                ("generated", vp), ("generated_bs", i64), ("gen_len", vp), ("next_token", vp), ("uniforms", fp),
                ("win_q", vp), ("win_sum", vp)]


class DecodeLayer(C.Structure):
    _fields_ = [("norm_weight", fp), ("in_proj_weight", vp), ("in_proj_bias", vp), ("conv_weight", fp), ("conv_bias", fp),
                ("conv_state", fp), ("x_proj_weight", vp), ("dt_weight", fp), ("dt_bias", fp), ("A", fp), ("D", fp),
                ("ssm_state", fp), ("out_proj_weight", vp), ("out_proj_bias", vp)]


class DecodeTokenArgs(C.Structure):
    _fields_ = [("struct_size", i32), ("w_dtype", i32), ("batch", i32), ("n_layers", i32), ("vocab", i32),
                ("d_model", i32), ("d_inner", i32), ("d_state", i32), ("dt_rank", i32), ("d_conv", i32),
                ("eps", C.c_float), ("flags", i32), ("token", vp), ("embedding", vp), ("layers", vp),
                ("norm_f_weight", fp), ("head_weight", vp), ("head_bias", vp), ("logits", fp), ("logits_bs", i64),
                ("scratch", fp), ("scratch_bytes", sz), ("barrier", vp)]


class LossArgs(C.Structure):
    _fields_ = [("struct_size", i32), ("dtype", i32), ("batch", i32), ("seqlen", i32), ("vocab", i32),
                ("boundaries", i32 * 4), ("reserved", i32),
                ("logits", vp), ("logits_bs", i64), ("logits_ts", i64), ("src", vp), ("trg", vp), ("table", fp),
                ("col_lse", fp), ("row_lse", fp), ("loss", fp), ("grad_out", fp),
                ("dlogits", vp), ("dlogits_bs", i64), ("dlogits_ts", i64), ("workspace", vp), ("workspace_bytes", sz)]


class MambaLibError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libmamba_b200.so (built by build.py).  Raises if it is missing — by design."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MambaLibError(
            f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built "
            "(run `python __graft_entry__.py build`); there is no CPU fallback for the Mamba hot path")
    L = C.CDLL(str(LIB_PATH))
    L.mamba_abi_version.restype = C.c_int
    L.mamba_last_error.restype = C.c_char_p
    L.mamba_launch_count.restype = C.c_uint64
    for name, argt in (("mamba_scan_fwd", ScanFwdArgs), ("mamba_scan_bwd", ScanBwdArgs),
                       ("mamba_conv1d_silu_fwd", ConvArgs), ("mamba_conv1d_silu_bwd", ConvArgs),
                       ("mamba_conv_step", StepArgs), ("mamba_ssm_step", StepArgs),
                       ("mamba_rmsnorm_fwd", NormArgs), ("mamba_rmsnorm_bwd", NormArgs),
                       ("mamba_filtered_ce_fwd", LossArgs), ("mamba_filtered_ce_bwd", LossArgs),
                       ("mamba_linear_step", LinearStepArgs), ("mamba_fused_linear_step", FusedLinearStepArgs),
                       ("mamba_sample_step", SampleStepArgs), ("mamba_decode_token", DecodeTokenArgs)):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [C.POINTER(argt), C.c_void_p]
    L.mamba_decode_token_scratch_bytes.restype = sz
    L.mamba_decode_token_scratch_bytes.argtypes = [C.c_int] * 4
    L.mamba_decode_token_barrier_bytes.restype = sz
    L.mamba_decode_token_barrier_bytes.argtypes = [C.c_int]
    L.mamba_scan_ckpt_elems.restype = sz
    L.mamba_scan_ckpt_elems.argtypes = [C.c_int] * 5
    L.mamba_scan_bwd_workspace_bytes.restype = sz
    L.mamba_scan_bwd_workspace_bytes.argtypes = [C.c_int] * 4
    L.mamba_conv1d_bwd_workspace_bytes.restype = sz
    L.mamba_conv1d_bwd_workspace_bytes.argtypes = [C.c_int] * 4
    L.mamba_filtered_ce_workspace_bytes.restype = sz
    L.mamba_filtered_ce_workspace_bytes.argtypes = [C.c_int] * 3
    L.mamba_rmsnorm_bwd_workspace_bytes.restype = sz
    L.mamba_rmsnorm_bwd_workspace_bytes.argtypes = [C.c_int64, C.c_int]
    if L.mamba_abi_version() != 3:
        raise MambaLibError(f"ABI version mismatch: library {L.mamba_abi_version()} != binding 3")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mamba_last_error().decode(errors="replace")
        raise MambaLibError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().mamba_launch_count())
