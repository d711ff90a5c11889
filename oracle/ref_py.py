"""Run the REFERENCE's own Python (pcdet/**) unmodified — TEST INFRASTRUCTURE, never imported by com_b200.

The reference package cannot simply be imported: `pcdet/__init__.py` wants a generated `version.py` run through
setup.py, the package `__init__`s pull in every detector / pointnet2 / roipoint pybind module, and four third-party
modules are absent from this image (easydict, skimage, SharedArray, tensorboardX — SURVEY.md §8c).  This loader
therefore

  * registers four tiny shims for those names (only what the hot-path modules touch at import time),
  * serves `pcdet.<a>.<b>` from the reference tree BY FILE PATH, treating every directory as an empty namespace
    package (the reference's package `__init__.py` files are not executed), and
  * on the GPU box, where /root/reference does not exist, serves the same modules from byte-code compiled here by
    `build_pyc()` into oracle/_ref/pcdet_bc/ (git-ignored like the compiled C++ reference next to it; produced from
    the sources where they lie, no reference source is copied into the repository).

Registries that the skipped `__init__`s would have defined (`backbones_3d.__all__` ...) are synthesised by
`registry()` from the hot-path classes only.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COM_REFERENCE", "/root/reference")
PYC_ROOT = os.path.join(HERE, "_ref", "pcdet_bc")      # byte-code files are named *.bc: *.pyc is filtered out of gpurun snapshots


def source_available():
    return os.path.isdir(os.path.join(REF, "pcdet", "models"))


def pyc_available():
    return os.path.exists(os.path.join(PYC_ROOT, "pcdet", "models", "backbones_3d", "spconv_backbone.bc"))


def available():
    return source_available() or pyc_available()


# the reference modules the hot-path tests execute (everything else in pcdet/ stays out of oracle/_ref)
MODULES = [
    "pcdet/config.py",
    "pcdet/utils/common_utils.py", "pcdet/utils/spconv_utils.py", "pcdet/utils/box_utils.py",
    "pcdet/utils/loss_utils.py",
    "pcdet/ops/iou3d_nms/iou3d_nms_utils.py", "pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py",
    "pcdet/datasets/processor/data_processor.py",
    "pcdet/models/backbones_3d/spconv_backbone.py",
    "pcdet/models/backbones_3d/vfe/vfe_template.py", "pcdet/models/backbones_3d/vfe/mean_vfe.py",
    "pcdet/models/backbones_2d/map_to_bev/height_compression.py", "pcdet/models/backbones_2d/base_bev_backbone.py",
    "pcdet/models/model_utils/model_nms_utils.py", "pcdet/models/model_utils/centernet_utils.py",
    "pcdet/models/dense_heads/center_head.py", "pcdet/models/dense_heads/curriculum_center_head.py",
    # CurriculumCenterHead_x5 (the COM head of BASELINE configs[2]) and what head_zoo.py imports next to it
    "pcdet/models/dense_heads/head_zoo.py", "pcdet/models/dense_heads/curri_anchor_head_single.py",
    "pcdet/models/dense_heads/anchor_head_curriculum.py",
    "pcdet/models/dense_heads/target_assigner/anchor_generator.py",
    "pcdet/models/dense_heads/target_assigner/atss_target_assigner.py",
    "pcdet/models/dense_heads/target_assigner/axis_aligned_target_assigner.py",
    "pcdet/models/dense_heads/target_assigner/curri_axis_aligned_target_assigner.py",
    "pcdet/utils/box_coder_utils.py",
    "pcdet/models/detectors/detector3d_template.py", "pcdet/models/detectors/centerpoint.py",
]
# directories that must exist as (empty) packages although no module of theirs is compiled
EXTRA_PACKAGES = ["pcdet/models/backbones_3d/pfe", "pcdet/models/roi_heads"]


def build_pyc():
    """Byte-compile MODULES from the reference tree into oracle/_ref/pcdet_bc (no-op without /root/reference)."""
    if not source_available():
        return False
    import warnings
    n = 0
    for rel in MODULES:
        dst = os.path.join(PYC_ROOT, rel[:-3] + ".bc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            py_compile.compile(os.path.join(REF, rel), cfile=dst, dfile=os.path.join("<reference>", rel), doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        n += 1
    for rel in EXTRA_PACKAGES:
        os.makedirs(os.path.join(PYC_ROOT, rel), exist_ok=True)
    return n


# ---------------------------------------------------------------------------------------------------------- shims
class EasyDict(dict):
    """easydict.EasyDict as pcdet/config.py and the model configs use it: attribute access, nested dicts wrapped."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {}, **kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = type(value)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in value)
        elif isinstance(value, dict) and not isinstance(value, EasyDict):
            value = EasyDict(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)


def _install_shims():
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__comb_shim__ = True
        sys.modules[name] = m
        return m

    try:
        import easydict  # noqa: F401
    except ImportError:
        mod("easydict", EasyDict=EasyDict)
    try:
        import skimage  # noqa: F401
    except ImportError:
        def _absent(*a, **k):
            raise RuntimeError("skimage is not installed (shim): image transforms are outside the hot path")
        tr = mod("skimage.transform", downscale_local_mean=_absent, resize=_absent)
        io = mod("skimage.io", imread=_absent, imsave=_absent)
        mod("skimage", transform=tr, io=io)
    try:
        import SharedArray  # noqa: F401
    except ImportError:
        def _absent_sa(*a, **k):
            raise RuntimeError("SharedArray is not installed (shim): shared-memory GT databases are outside the hot path")
        mod("SharedArray", attach=_absent_sa, create=_absent_sa, delete=_absent_sa)
    try:
        import tensorboardX  # noqa: F401
    except ImportError:
        class SummaryWriter:                          # tools/train.py only
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass
        mod("tensorboardX", SummaryWriter=SummaryWriter)


# ---------------------------------------------------------------------------------------------------------- finder
class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        if name != "pcdet" and not name.startswith("pcdet."):
            return None
        rel = name.replace(".", os.sep)
        roots = []
        if source_available():
            roots.append((REF, ".py", importlib.machinery.SourceFileLoader))
        if pyc_available():
            roots.append((PYC_ROOT, ".bc", importlib.machinery.SourcelessFileLoader))
        for root, ext, loader_cls in roots:
            f = os.path.join(root, rel + ext)
            if os.path.isfile(f):
                return importlib.util.spec_from_file_location(name, f, loader=loader_cls(name, f))
            d = os.path.join(root, rel)
            if os.path.isdir(d):                      # a package: empty namespace, the reference __init__ is skipped
                spec = importlib.machinery.ModuleSpec(name, None, is_package=True)
                spec.submodule_search_locations = [d]
                return spec
        return None


_installed = False


def install(dropins=True, accelerate=True):
    """Shims + (optionally) com_b200.install_dropins() + the by-path finder for `pcdet.*`."""
    global _installed
    if not available():
        raise FileNotFoundError("neither %s nor %s exists: run oracle/build_ref.py where the reference is mounted"
                                % (REF, PYC_ROOT))
    _install_shims()
    if not _installed:
        # ahead of the standard PathFinder (which would execute the reference's package __init__s); com_b200's
        # post-import finder, inserted at position 0 by install_dropins below, delegates to this one
        sys.meta_path.insert(0, _RefFinder())
        _installed = True
    if dropins:
        import com_b200
        com_b200.install_dropins(accelerate=accelerate)


def load(name):
    """import_module of a reference module by dotted name, e.g. 'pcdet.models.backbones_3d.spconv_backbone'."""
    install()
    return importlib.import_module(name)


def registry():
    """The `__all__` registries of the package `__init__`s this loader skips, restricted to the hot-path classes
    (pcdet/models/backbones_3d/__init__.py:6-13, vfe/__init__.py, map_to_bev/__init__.py, backbones_2d/__init__.py,
    dense_heads/__init__.py) — set on the synthetic packages so that Detector3DTemplate.build_networks finds them."""
    install()
    b3 = importlib.import_module("pcdet.models.backbones_3d")
    b3.__all__ = {"VoxelResBackBone8x": load("pcdet.models.backbones_3d.spconv_backbone").VoxelResBackBone8x,
                  "VoxelBackBone8x": load("pcdet.models.backbones_3d.spconv_backbone").VoxelBackBone8x}
    vfe = importlib.import_module("pcdet.models.backbones_3d.vfe")
    vfe.__all__ = {"MeanVFE": load("pcdet.models.backbones_3d.vfe.mean_vfe").MeanVFE,
                   "VFETemplate": load("pcdet.models.backbones_3d.vfe.vfe_template").VFETemplate}
    pfe = importlib.import_module("pcdet.models.backbones_3d.pfe")
    pfe.__all__ = {}
    b2 = importlib.import_module("pcdet.models.backbones_2d")
    b2.__all__ = {"BaseBEVBackbone": load("pcdet.models.backbones_2d.base_bev_backbone").BaseBEVBackbone}
    m2b = importlib.import_module("pcdet.models.backbones_2d.map_to_bev")
    m2b.__all__ = {"HeightCompression": load("pcdet.models.backbones_2d.map_to_bev.height_compression").HeightCompression}
    dh = importlib.import_module("pcdet.models.dense_heads")
    dh.__all__ = {"CenterHead": load("pcdet.models.dense_heads.center_head").CenterHead}
    try:
        for modname in ("pcdet.models.dense_heads.curriculum_center_head", "pcdet.models.dense_heads.head_zoo"):
            cur = load(modname)
            for k in dir(cur):
                if k.startswith("CurriculumCenterHead"):
                    dh.__all__[k] = getattr(cur, k)
    except Exception as e:  # pragma: no cover  (reported by the test that needs it)
        dh.__comb_curriculum_error__ = e
    rh = importlib.import_module("pcdet.models.roi_heads")
    rh.__all__ = {}
    return {"backbones_3d": b3, "vfe": vfe, "pfe": pfe, "backbones_2d": b2, "map_to_bev": m2b, "dense_heads": dh,
            "roi_heads": rh}


if __name__ == "__main__":
    print("byte-compiled %s reference modules into %s" % (build_pyc(), PYC_ROOT))
