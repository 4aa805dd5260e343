"""TEST INFRASTRUCTURE ONLY: lets the reference's own `config.py` (imported from /root/reference where it lies, never
copied) be imported in a container that lacks the reference's third-party stack.

The reference imports pytorch3d, frnn, trimesh, plyfile, easydict, matplotlib, skimage ... at module scope all over its
package.  `install()` registers a meta-path finder that serves permissive stub modules for those
top-level names -- any attribute is a stub class that can be subclassed, called and decorated with -- except for the
handful of names the plugin boundary really uses, which are mapped to the real implementations of this repo:

    pytorch3d.renderer.FoVPerspectiveCameras (+ .cameras)  -> dss_b200.core.camera.FoVPerspectiveCameras
    pytorch3d.renderer.NormWeightedCompositor              -> dss_b200.core.renderer.NormWeightedCompositor
    easydict.EasyDict                                       -> a 15-line attribute dict
    DSS._C (the reference's pybind module, not built in its tree) -> dss_b200._C, the drop-in of INTEGRATION.md

With that in place `config.create_renderer(cfg.renderer)` (config.py:241-262) runs UNMODIFIED and builds whatever classes
the YAML names -- tests/test_reference_factory.py points it at dss_b200's, as INTEGRATION.md tells a user to.
"""
import importlib.abc
import importlib.machinery
import sys
import types

STUB_ROOTS = ("pytorch3d", "frnn", "trimesh", "plyfile", "easydict", "matplotlib", "skimage", "imageio", "open3d",
              "pymeshlab", "point_cloud_utils", "torch_batch_svd", "prefix_sum", "git", "cv2", "plotly", "seaborn",
              "tensorboard", "tensorboardX", "OpenEXR", "Imath", "sklearn", "scipy_stub_never")


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(name)

    def __call__(cls, *a, **k):
        # used as a decorator -> hand the function back; otherwise build an inert instance
        if len(a) == 1 and not k and callable(a[0]) and not isinstance(a[0], type):
            return a[0]
        return super().__call__(*a, **k)


class _StubBase(metaclass=_StubMeta):
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _StubBase()

    def __call__(self, *a, **k):
        return _StubBase()

    def __iter__(self):
        return iter(())


def _make_stub(name):
    return _StubMeta(name, (_StubBase,), {})


class EasyDict(dict):
    """attribute access on a (nested) dict -- what config.py needs from easydict"""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class _StubModule(types.ModuleType):
    def __init__(self, name, real):
        super().__init__(name)
        self.__path__ = []
        self.__dict__["_real"] = real
        self.__dict__["_cache"] = {}

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in self._real:
            return self._real[full]
        if name not in self._cache:
            self._cache[name] = _make_stub(name)
        return self._cache[name]


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, real):
        self.real = real

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        if fullname == "DSS._C":     # the reference's native module is not built in its tree: our drop-in takes its place
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        if spec.name == "DSS._C":
            import dss_b200._C as native
            proxy = types.ModuleType(spec.name, native.__doc__)        # same callables under the reference's module name
            proxy.__dict__.update({k: v for k, v in vars(native).items() if not k.startswith("__")})
            return proxy
        return _StubModule(spec.name, self.real)

    def exec_module(self, module):
        pass


_installed = None


def install():
    """idempotent; returns the finder"""
    global _installed
    if _installed is not None:
        return _installed
    from dss_b200.core.camera import FoVPerspectiveCameras
    from dss_b200.core.renderer import NormWeightedCompositor
    real = {
        "pytorch3d.renderer.FoVPerspectiveCameras": FoVPerspectiveCameras,
        "pytorch3d.renderer.cameras.FoVPerspectiveCameras": FoVPerspectiveCameras,
        "pytorch3d.renderer.NormWeightedCompositor": NormWeightedCompositor,
        "easydict.EasyDict": EasyDict,
    }
    missing = []
    for root in STUB_ROOTS:
        try:
            if importlib.util.find_spec(root) is None:
                missing.append(root)
        except (ImportError, ValueError):
            missing.append(root)
    f = _Finder(real)
    # only names that are really absent are stubbed; an installed package is left alone
    globals()["STUB_ROOTS"] = tuple(missing)
    sys.meta_path.append(f)
    _installed = f
    return f


def uninstall():
    global _installed
    if _installed is not None:
        sys.meta_path.remove(_installed)
        for m in [m for m in sys.modules if m.split(".")[0] in STUB_ROOTS]:
            del sys.modules[m]
        _installed = None
