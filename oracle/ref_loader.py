"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference on CPU.

Only usable in the build container (the GPU box has no /root/reference).  Used to (i) validate the
restatement in ``oracle/oracle.py`` and (ii) generate the golden fixtures under ``tests/golden/``
(``oracle/make_golden.py``).  Nothing in the product path imports this file.

How the reference is made importable:
  * ``oracle/shims`` provides torch_scatter / torchsparse.utils.quantize / open3d stand-ins,
  * ``chamfer_distance/chamfer_distance.py`` JIT-builds its extension at import time with paths
    relative to the reference root, so the import happens with cwd=/root/reference,
    ``TORCH_EXTENSIONS_DIR=oracle/_ref`` (git-ignored build output) and
    ``TORCH_CUDA_ARCH_LIST=10.0a`` (no GPU here to auto-detect).
"""
import contextlib
import os
import sys

REF_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(_HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_loaded = {}


def load():
    """Import the reference modules; returns a namespace dict."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    os.makedirs(REF_BUILD, exist_ok=True)
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", REF_BUILD)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    shims = os.path.join(_HERE, "shims")
    for p in (REF_ROOT, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    with _cwd(REF_ROOT):
        from models.motionnet import MotionNet  # noqa
        from libs.voxel_generator import Voxelization  # noqa
        from libs.dataloader import collate_fn  # noqa
        from chamfer_distance.chamfer_distance import ChamferDistance  # noqa
        import toolbox.register_utils as register_utils  # noqa
        import toolbox.utils as utils  # noqa
    _loaded.update(MotionNet=MotionNet, Voxelization=Voxelization, collate_fn=collate_fn,
                   ChamferDistance=ChamferDistance, register_utils=register_utils, utils=utils)
    return _loaded


def reference_config(dataset="waymo", mode="test", overrides=None):
    """default.yaml (+) dataset yaml (+) overrides, then main.py:10-14's update_config."""
    import yaml

    with open(os.path.join(REF_ROOT, "configs/default.yaml")) as f:
        cfg = yaml.safe_load(f)
    with open(os.path.join(REF_ROOT, f"configs/{dataset}/{dataset}.yaml")) as f:
        special = yaml.safe_load(f)
    sys.path.insert(0, REF_ROOT)
    from toolbox.config import update_recursive

    update_recursive(cfg, special)
    if overrides:
        update_recursive(cfg, overrides)
    cfg["misc"]["mode"] = mode
    cfg["pillar_encoder"]["voxel_size"] = cfg["voxel_generator"]["voxel_size"]
    cfg["pillar_encoder"]["pc_range"] = cfg["voxel_generator"]["range"]
    cfg["pillar_encoder"]["n_sweeps"] = cfg["voxel_generator"]["n_sweeps"]
    return cfg
