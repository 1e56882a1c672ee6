"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference Python modules from /root/reference on
torch-CPU so that golden vectors can be minted from the reference's own code
(SURVEY.md section 8c, Appendix E).  /root/reference only exists in the build
container; nothing that runs on the GPU box may call into this file -- the
minted vectors travel as tests/golden/*.npz instead.

The reference pulls in ~16 third-party modules that are absent here.  None of
them participate in the arithmetic of the hot path, so they are replaced by
inert stand-ins *before* `train_shot` / `train_dino` are imported:

  pytorch_lightning.LightningModule -> torch.nn.Module
  torch_scatter.scatter_add         -> zeros(dim_size).scatter_add_ (1-D use only:
                                       train_dino.py:204, eval.py:265)
  hydra.main                        -> identity decorator (train_shot.py:133)
  everything else                   -> empty module objects

`get_topk_dir` lives in eval.py (eval.py:37-51), whose import would need
fire/lietorch/DINOv2; only that FunctionDef is lifted with `ast`, with the
`.cuda()` / device='cuda' placements stripped so it runs on torch-CPU.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CPPF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "train_dino.py"))


class _Anything:
    """Absorbs any attribute access / call made at import time."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behaves as a package so that `import a.b` works
    def _missing(attr, _name=name):
        if attr.startswith("__"):
            raise AttributeError(f"stub module {_name!r} has no attribute {attr!r}")
        return _Anything

    m.__getattr__ = _missing  # type: ignore[attr-defined]
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    import torch

    def scatter_add(src, index, dim=-1, out=None, dim_size=None):
        if dim_size is None:
            dim_size = int(index.max()) + 1
        return torch.zeros(int(dim_size), dtype=src.dtype).scatter_add_(0, index, src)

    def hydra_main(*a, **k):
        return lambda fn: fn

    _mod("pytorch_lightning", LightningModule=torch.nn.Module, Trainer=_Anything)
    _mod("pytorch_lightning.loggers", TensorBoardLogger=_Anything)
    _mod("pytorch_lightning.callbacks", ModelCheckpoint=_Anything)
    _mod("hydra", main=hydra_main)
    _mod("hydra.utils", to_absolute_path=lambda p: p)
    _mod("omegaconf", OmegaConf=_Anything)
    _mod("torch_scatter", scatter_add=scatter_add)
    for name in (
        "src_shot", "src_shot.build", "src_shot.build.shot", "open3d", "trimesh", "zmq",
        "wandb", "icecream", "skimage", "skimage.color", "pycocotools", "pycocotools.cocoeval",
        "pycocotools.mask", "matplotlib", "matplotlib.pyplot", "scipy.misc", "pyrender", "lietorch",
        "fire",
    ):
        _mod(name)
    _mod("visdom", Visdom=_Anything)

    class ImageOnlyTransform:  # utils/util.py:122,140 subclass it at import time
        def __init__(self, *a, **k):
            pass

    _mod("albumentations", ImageOnlyTransform=ImageOnlyTransform)
    _mod("albumentations.core", transforms_interface=_Anything)
    _mod("albumentations.core.transforms_interface", ImageOnlyTransform=ImageOnlyTransform)
    _mod("albumentations.pytorch", ToTensorV2=_Anything)


_loaded = None


def load_reference():
    """Returns a namespace with the reference's own hot-path callables (torch-CPU)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    train_shot = importlib.import_module("train_shot")
    train_dino = importlib.import_module("train_dino")
    dataset = importlib.import_module("dataset")
    util = importlib.import_module("utils.util")

    import numpy as np
    import torch

    if not hasattr(np, "product"):  # eval.py:229 uses the numpy-1 spelling
        np.product = np.prod  # type: ignore[attr-defined]

    # lift get_topk_dir out of eval.py without importing eval.py
    src = open(os.path.join(REFERENCE_ROOT, "eval.py")).read()
    tree = ast.parse(src)
    fn_src = None
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "get_topk_dir":
            fn_src = ast.get_source_segment(src, node)
    assert fn_src is not None
    fn_src = fn_src.replace(".cuda()", "").replace(", device='cuda'", "")
    scope = {"torch": torch, "np": np}
    exec(compile(fn_src, "eval.py::get_topk_dir", "exec"), scope)

    ns = types.SimpleNamespace(
        BeyondCPPFSHOT=train_shot.BeyondCPPF,
        BeyondCPPFDINO=train_dino.BeyondCPPF,
        vote_center=train_dino.vote_center,
        vote_rotation=train_dino.vote_rotation,
        generate_target_pairs=dataset.generate_target_pairs,
        fibonacci_sphere=util.fibonacci_sphere,
        backproject=util.backproject,
        get_topk_dir=scope["get_topk_dir"],
        shapenet_obj_scales=dataset.shapenet_obj_scales,
    )
    _loaded = ns
    return ns


if __name__ == "__main__":
    ref = load_reference()
    print("reference import ok:", [k for k in vars(ref)])
