"""Import the UNMODIFIED reference (SkylerGao/MC_NeRF) from baseline/_ref/ (or /root/reference in the build
container) on a machine that lacks six of its imports.

Used by: bench.py --impl reference / cpu_baseline (the reference's own CPU path as the timed baseline),
tests/golden/make_golden.py (fixture generation) and tests/test_main_integration*.py (the reference's main.py driving
the drop-in package).  The product (mc_nerf_b200/) never imports this.

The stand-ins below are inert: plotting handles, the LPIPS metric, the AprilTag detector and the pretty-table
printer are never exercised on the train/render path.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


def reference_root():
    """Directory holding the reference's main.py, or None."""
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "main.py")) and os.path.isfile(os.path.join(c, "model", "mc_nerf.py")):
            return c
    return None


class _Inert:
    """Callable / indexable / addable no-op used for plotting handles."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __getitem__(self, k):
        return _Inert()

    def __setitem__(self, k, v):
        pass

    def __add__(self, o):
        return _Inert()

    __radd__ = __add__

    def __iter__(self):
        return iter(())


def _stub(name, **attrs):
    m = types.ModuleType(name)

    def _getattr(attr, _name=name):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)      # inspect / importlib probe dunder names
        return _Inert()

    m.__getattr__ = _getattr
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _NoLPIPS:
    """lpips.LPIPS stand-in: the perceptual metric needs downloaded AlexNet weights; it reports 0."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, a, b):
        import torch
        return torch.zeros(())


def _importable(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install_import_shims():
    """Register stand-ins for the packages the reference imports that this image does not have."""
    if "lpips" not in sys.modules and not _importable("lpips"):
        _stub("lpips", LPIPS=_NoLPIPS)
    if "apriltag" not in sys.modules and not _importable("apriltag"):
        _stub("apriltag")
    if "prettytable" not in sys.modules and not _importable("prettytable"):
        _stub("prettytable", PrettyTable=_Inert)
    if "matplotlib" not in sys.modules and not _importable("matplotlib"):
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.cm = _stub("matplotlib.cm")
    if "mpl_toolkits" not in sys.modules and not _importable("mpl_toolkits.mplot3d"):
        tk = _stub("mpl_toolkits")
        tk.mplot3d = _stub("mpl_toolkits.mplot3d", Axes3D=_Inert)


def _purge(prefixes):
    for k in [k for k in sys.modules if any(k == p or k.startswith(p + ".") for p in prefixes)]:
        del sys.modules[k]


def import_reference(root=None):
    """-> (MC_Model, NeRF_Model, MC_NeRF_Loss, net_block, net_utils) of the unmodified reference."""
    root = root or reference_root()
    if root is None:
        raise ImportError("reference not installed: run `python baseline/install_ref.py` in the build container")
    install_import_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    _purge(("model",))          # make sure "model" resolves to the reference's package
    from model.mc_nerf import MC_Model, NeRF_Model  # noqa
    from model.loss import MC_NeRF_Loss  # noqa
    import model.net_block as net_block  # noqa
    import model.net_utils as net_utils  # noqa
    return MC_Model, NeRF_Model, MC_NeRF_Loss, net_block, net_utils


def import_reference_main(model_package=None, data_module=None, root=None):
    """Import the reference's main.py as module `main` (its Model_Engine drives training / rendering).

    model_package: module object to serve as `model` (e.g. mc_nerf_b200.model - the drop-in), or None for the
                   reference's own model package;
    data_module:   module object to serve as `data` (must export Data_set and Data_loader), or None for the
                   reference's own (needs cv2 + apriltag + a dataset on disk).
    """
    root = root or reference_root()
    if root is None:
        raise ImportError("reference not installed: run `python baseline/install_ref.py` in the build container")
    install_import_shims()
    _purge(("main", "model", "data", "config", "utils"))
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    if model_package is not None:
        sys.modules["model"] = model_package
        for sub in ("mc_nerf", "net_block", "net_utils", "loss", "external", "external.pohsun_ssim",
                    "external.pohsun_ssim.pytorch_ssim"):
            full = model_package.__name__ + "." + sub
            try:
                sys.modules["model." + sub] = importlib.import_module(full)
            except ImportError:
                pass
    if data_module is not None:
        sys.modules["data"] = data_module
    return importlib.import_module("main")
