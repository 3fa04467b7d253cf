"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/mamba_b200.h declares,
and the ctypes mirrors of the argument structs have the C compiler's layout.  No compute calls."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "mamba_b200.h"


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|size_t|uint64_t|char\s*\*)\s*\*?\s*(mamba_\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for must in ("mamba_scan_fwd", "mamba_scan_bwd", "mamba_conv1d_silu_fwd", "mamba_conv1d_silu_bwd",
                 "mamba_conv_step", "mamba_ssm_step", "mamba_rmsnorm_fwd", "mamba_rmsnorm_bwd", "mamba_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from mamba_b200 import _lib
    for name in _declared_functions():
        assert hasattr(lib, name), f"libmamba_b200.so does not export {name}"
    assert set(_lib.EXPORTS) == set(_declared_functions())
    assert lib.mamba_abi_version() == 3


def test_ctypes_struct_layout_matches_c(tmp_path):
    """Compile a C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    from mamba_b200 import _lib
    structs = {"MambaScanFwdArgs": _lib.ScanFwdArgs, "MambaScanBwdArgs": _lib.ScanBwdArgs,
               "MambaConvArgs": _lib.ConvArgs, "MambaStepArgs": _lib.StepArgs, "MambaNormArgs": _lib.NormArgs,
               "MambaLossArgs": _lib.LossArgs, "MambaLinearStepArgs": _lib.LinearStepArgs,
               "MambaFusedLinearStepArgs": _lib.FusedLinearStepArgs, "MambaSampleStepArgs": _lib.SampleStepArgs,
               "MambaDecodeLayer": _lib.DecodeLayer, "MambaDecodeTokenArgs": _lib.DecodeTokenArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_flag_and_dtype_constants_match_the_header(tmp_path):
    """The python mirrors of the header's enums (flags, dtypes) are compiled against the header itself."""
    from mamba_b200 import _lib
    names = {"MAMBA_FLAG_HAS_Z": _lib.FLAG_HAS_Z, "MAMBA_FLAG_DELTA_SOFTPLUS": _lib.FLAG_DELTA_SOFTPLUS,
             "MAMBA_FLAG_HAS_DELTA_BIAS": _lib.FLAG_HAS_DELTA_BIAS, "MAMBA_FLAG_HAS_D": _lib.FLAG_HAS_D,
             "MAMBA_FLAG_A_IS_LOG": _lib.FLAG_A_IS_LOG, "MAMBA_F32": _lib.MAMBA_F32, "MAMBA_BF16": _lib.MAMBA_BF16}
    lines = ['#include <stdio.h>', f'#include "{HEADER}"', "int main(void){"]
    lines += [f'printf("{n} %d\\n", (int){n});' for n in names]
    lines.append("return 0;}")
    src = tmp_path / "consts.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "consts"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for n, v in names.items():
        assert int(got[n]) == v, n


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call, with a message (error behaviour of the boundary)."""
    from mamba_b200 import _lib
    a = _lib.ScanFwdArgs()
    a.struct_size = 3
    assert lib.mamba_scan_fwd(C.byref(a), None) == -1
    assert b"struct_size" in lib.mamba_last_error()
    a.struct_size = C.sizeof(_lib.ScanFwdArgs)
    a.batch, a.seqlen, a.dim, a.dstate = 1, 0, 4, 4
    assert lib.mamba_scan_fwd(C.byref(a), None) == -1
    assert b"positive" in lib.mamba_last_error()
    n = _lib.NormArgs()
    n.struct_size = C.sizeof(_lib.NormArgs)
    n.rows, n.dim, n.dtype, n.resid_dtype = 4, 8, 0, 1
    assert lib.mamba_rmsnorm_fwd(C.byref(n), None) == -2
    assert lib.mamba_scan_ckpt_elems(2, 100, 64, 16, 16) == 2 * 7 * 16 * 64
    assert lib.mamba_scan_ckpt_elems(1, 16, 32, 5, 16) == 1 * 1 * 8 * 32   # d_state padded to a multiple of 4
    assert lib.mamba_scan_bwd_workspace_bytes(0, 1, 1, 1) == 0


def test_ops_fail_loudly_on_cpu_tensors():
    import torch
    from mamba_b200 import ops
    x = torch.randn(1, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.causal_conv1d_silu_fn(x, torch.randn(8, 1, 4), torch.randn(8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.selective_scan_fn(x, x, -torch.ones(8, 2), torch.randn(1, 4, 2), torch.randn(1, 4, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.rmsnorm_fn(x, torch.ones(8))


def test_product_package_never_imports_oracle():
    pkg = ROOT / "deep-learning-based-sequence-models-for-music-generation_b200"
    for py in pkg.rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), py
