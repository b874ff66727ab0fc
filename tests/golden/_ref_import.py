"""Import the UNMODIFIED reference modules from /root/reference on CPU.

Only used by tests/golden/make_golden.py (fixture generation, in the build
container).  Nothing in tests/, bench.py or smoke() imports this at run time:
/root/reference does not exist on the GPU box.

The reference imports six packages that are not installed here (lpips,
prettytable, matplotlib(.pyplot/.cm), mpl_toolkits.mplot3d, apriltag); we
register inert stand-ins so `from model.mc_nerf import ...` succeeds.  The
stand-ins are never exercised by the hot path.
"""
import sys
import types

REF_ROOT = "/root/reference"


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
            raise AttributeError(attr)
        return _Inert()

    m.__getattr__ = _getattr
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (MC_Model, NeRF_Model, MC_NeRF_Loss, net_block, net_utils)."""
    for name in ("lpips", "apriltag"):
        if name not in sys.modules:
            _stub(name)
    if "prettytable" not in sys.modules:
        _stub("prettytable", PrettyTable=_Inert)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.cm = _stub("matplotlib.cm")
    if "mpl_toolkits" not in sys.modules:
        tk = _stub("mpl_toolkits")
        tk.mplot3d = _stub("mpl_toolkits.mplot3d", Axes3D=_Inert)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # make sure "model" resolves to the reference's package, not ours
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    from model.mc_nerf import MC_Model, NeRF_Model  # noqa
    from model.loss import MC_NeRF_Loss  # noqa
    import model.net_block as net_block  # noqa
    import model.net_utils as net_utils  # noqa
    return MC_Model, NeRF_Model, MC_NeRF_Loss, net_block, net_utils
