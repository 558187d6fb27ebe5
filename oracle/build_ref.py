"""Stage the reference's own implementation of the path for the CPU timing legs.   *** TEST INFRASTRUCTURE ***

The reference is pure Python (no build step): "building" it means placing the UNMODIFIED files the path needs --
``models.py`` and ``configure/`` (cfgs.py reads its defaults through yacs) -- under the git-ignored ``oracle/_ref/`` so that
they travel to the GPU box with the snapshot (``/root/reference`` does not exist there).  Nothing is copied into the tracked
tree.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg import ``oracle/_ref/models.py`` when it is present
(``kind: "reference"``) and fall back to the oracle port otherwise (``kind: "port"``).

    python oracle/build_ref.py        (also called by __graft_entry__.build())
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["models.py", os.path.join("configure", "__init__.py"), os.path.join("configure", "cfgs.py")]
# the reference's training loop and what it imports: staged for the drop-in test that drives the CUDA model through
# train_funcs.train_autoencoder_dataloader itself (tests/test_gpu_dropin_loop.py)
LOOP_FILES = ["train_funcs.py", "utils_SH.py", "utils_distance.py", "mesh_sampling.py", "shape_data.py", "utils_spiral.py"]


def build(verbose=False):
    """Returns True if oracle/_ref/ holds the reference files (freshly staged or already there)."""
    if not os.path.isdir(REF):
        return available()
    for rel in FILES + LOOP_FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", rel)
    return available()


def available(loop=False):
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES + (LOOP_FILES if loop else []))


def import_reference_module(name):
    """Import a staged reference module by name (e.g. "train_funcs") with the inert third-party stand-ins of
    tests/golden/_ref_stubs.py (yacs, psbody, opendr, trimesh, torch_scatter, tensorboardX)."""
    if not available(loop=True):
        raise ImportError("oracle/_ref does not hold the reference's training loop (run oracle/build_ref.py)")
    import importlib

    sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
    import _ref_stubs

    _ref_stubs.install()
    while "/root/reference" in sys.path:  # the staged copy is the one under test, here and on the GPU box
        sys.path.remove("/root/reference")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    return importlib.import_module(name)


def import_reference_models():
    """Import oracle/_ref/models.py (the reference's file, unmodified) with an inert stand-in for yacs."""
    if not available():
        raise ImportError("oracle/_ref is not staged (run oracle/build_ref.py in the build container)")
    import types

    if "yacs" not in sys.modules:
        class CfgNode(dict):
            def __init__(self, init_dict=None, new_allowed=False, **_):
                super().__init__()

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v

        yacs = types.ModuleType("yacs")
        yacs.config = types.ModuleType("yacs.config")
        yacs.config.CfgNode = CfgNode
        sys.modules["yacs"], sys.modules["yacs.config"] = yacs, yacs.config
    import importlib.util

    if DST not in sys.path:
        sys.path.insert(0, DST)  # models.py does `from configure.cfgs import cfg`
    spec = importlib.util.spec_from_file_location("_reference_models", os.path.join(DST, "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("oracle/_ref staged:", build(verbose=True))
