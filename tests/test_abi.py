"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/mcnerf.h declares."""
import ctypes
import os

from mc_nerf_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m mc_nerf_b200.build` (the driver's build() does)"
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    protos = _lib.parse_header()
    assert len(protos) >= 26
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/mcnerf.h but not exported"
    assert _lib.lib().cdll.mcnerf_abi_version() == 1


def test_bad_arguments_return_error_codes_not_crashes():
    L = _lib.lib()
    rc = L.cdll.mcnerf_se3_fwd(None, 0, None, None)
    assert rc == 10001
    assert b"bad argument" in L.cdll.mcnerf_last_error()
    try:
        L.call("mcnerf_se3_fwd", None, 0, None, None)
    except _lib.McnerfError as e:
        assert "mcnerf_se3_fwd" in str(e)
    else:
        raise AssertionError("expected McnerfError")


def test_cpu_tensors_are_rejected_loudly():
    import pytest
    import torch
    from mc_nerf_b200 import ops
    with pytest.raises(_lib.McnerfError):
        ops.SE3Fn.apply(torch.ones(2, 6))
