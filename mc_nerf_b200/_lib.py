"""ctypes binding of libmcnerf.so (the C ABI declared in include/mcnerf.h).

The prototypes are parsed from the header itself so the Python side cannot drift from the ABI.
There is NO fallback: if the shared library is missing or a call fails, we raise.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "mcnerf.h")
LIB_PATH = os.path.join(HERE, "csrc", "libmcnerf.so")

MAX_DEPTH = 16
MAX_FREQS = 16
_fp = ctypes.c_void_p


class Sampling(ctypes.Structure):
    _fields_ = [("near_", ctypes.c_float), ("far_", ctypes.c_float), ("S", ctypes.c_int),
                ("n_freqs", ctypes.c_int), ("band_w", ctypes.c_float * MAX_FREQS), ("band_w_dev", ctypes.c_void_p)]


class MlpParams(ctypes.Structure):
    _fields_ = [("depth", ctypes.c_int), ("width", ctypes.c_int), ("in_ch", ctypes.c_int),
                ("skip_mask", ctypes.c_uint32), ("sh_dim", ctypes.c_int),
                ("W", _fp * MAX_DEPTH), ("b", _fp * MAX_DEPTH),
                ("W_sigma0", _fp), ("b_sigma0", _fp), ("W_sigma2", _fp), ("b_sigma2", _fp),
                ("W_sh0", _fp), ("b_sh0", _fp), ("W_sh2", _fp), ("b_sh2", _fp)]


class MlpGrads(ctypes.Structure):
    _fields_ = [("W", _fp * MAX_DEPTH), ("b", _fp * MAX_DEPTH),
                ("W_sigma0", _fp), ("b_sigma0", _fp), ("W_sigma2", _fp), ("b_sigma2", _fp),
                ("W_sh0", _fp), ("b_sh0", _fp), ("W_sh2", _fp), ("b_sh2", _fp)]


class Dirs(ctypes.Structure):
    _fields_ = [("dirs", _fp), ("dir_idx", _fp), ("dir_S", ctypes.c_int)]


class CompositeCfg(ctypes.Structure):
    _fields_ = [("near_", ctypes.c_float), ("far_", ctypes.c_float), ("S", ctypes.c_int),
                ("white_back", ctypes.c_int)]


class TcInput(ctypes.Structure):
    _fields_ = [("rays_o", _fp), ("rays_d", _fp), ("jitter", _fp), ("n_rays", ctypes.c_int), ("smp", Sampling),
                ("sample_idx", _fp), ("n_rows", ctypes.c_int), ("n_rows_dev", _fp),
                ("x_enc", _fp), ("ld_enc", ctypes.c_int), ("dirs_rows", _fp),
                ("ray_offsets", _fp), ("ordered_ray_grads", ctypes.c_int)]


class P2P(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int), ("n_ranks", ctypes.c_int), ("n_ctas", ctypes.c_int),
                ("buf", _fp * 8), ("flags", _fp * 8), ("epoch", _fp)]


_STRUCTS = {"mcnerf_p2p": P2P, "mcnerf_tc_input": TcInput, "mcnerf_sampling": Sampling, "mcnerf_mlp_params": MlpParams, "mcnerf_mlp_grads": MlpGrads,
            "mcnerf_dirs": Dirs, "mcnerf_composite_cfg": CompositeCfg}
_SCALARS = {"int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t,
            "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32, "void": None}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#[^\n]*", " ", src, flags=re.M)
    src = re.sub(r'extern\s+"C"\s*\{', " ", src)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(mcnerf_\w+)\s*\(([^;{}]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        else:
            restype = _SCALARS[ret.replace("const", "").strip()]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.replace("const", " ").split())
                if "*" in a:
                    base = a.split("*")[0].strip()
                    argtypes.append(ctypes.POINTER(_STRUCTS[base]) if base in _STRUCTS else ctypes.c_void_p)
                else:
                    argtypes.append(_SCALARS[a.split()[0]])
        protos[name] = (restype, argtypes)
    return protos


class McnerfError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise McnerfError(
                f"{LIB_PATH} is missing: build it with `python -m mc_nerf_b200.build` "
                "(there is no CPU or PyTorch fallback for the MC-NeRF hot path)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)      # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if self.cdll.mcnerf_abi_version() != 1:
            raise McnerfError("libmcnerf.so ABI version mismatch")

    def profile_begin(self):
        """Record a CUDA-event pair around every subsequent call (on the launching stream); bench.py uses this
        for the per-kernel device times behind the roofline numbers."""
        self._prof = []

    def profile_end(self):
        """-> {entry name: total milliseconds} and stop recording."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self._prof:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        self._prof = None
        return out

    def profiling(self):
        return getattr(self, "_prof", None) is not None

    def call(self, name, *args, label=None):
        """Call an int-returning entry; raise McnerfError(mcnerf_last_error()) on failure."""
        if getattr(self, "_prof", None) is not None:
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = getattr(self.cdll, name)(*args)
            e1.record()
            self._prof.append((label or name, e0, e1))
        else:
            rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            msg = self.cdll.mcnerf_last_error()
            raise McnerfError(f"{name} failed (rc={rc}): {msg.decode() if msg else ''}")

    def launch_count(self):
        return int(self.cdll.mcnerf_launch_count())


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB
