"""Harness for BASELINE configs[3] / configs[4]: run the reference's OWN model code on this repo's kernels.

`stage()` (called by `__graft_entry__.build()` where /root/reference exists) copies the reference's Python
sources (models/, config/, myutils/, loss/, dataloader/, datalist/; *.py and *.yml only) into the git-ignored
`baseline/_ref/ebfi_be/`, which travels to the GPU box with the snapshot. Nothing there is product code and
nothing is committed.

`load()` puts that tree on sys.path, installs import stubs for the plotting / IO packages the reference imports at
module scope but does not need for a forward pass (matplotlib, mpl_toolkits, open3d, h5py,
`torchvision.models.utils`; SURVEY.md §8c "import harness"), registers this repo's `_ext` / `kernelconv2d_cuda`
shims, and imports `models.Ours.model_singleframe` UNCHANGED. `build_model()` constructs
`EVFIAutoEx(**config/train_ours.yml:model.args)` (reference: train_ours.py:221-232).

`use_reference_cuda_fac(True)` swaps `kernelconv2d_cuda` for the reference's own CUDA kernels compiled for sm_100a
(`oracle/_ref/fac_cuda`, test infrastructure) so that the same model can be timed A/B.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
STAGE = os.path.join(ROOT, "baseline", "_ref", "ebfi_be")
SUBDIRS = ("models", "config", "myutils", "loss", "dataloader", "datalist")


def stage(verbose=False):
    """Copy the reference's *.py / *.yml of the model tree into baseline/_ref/ebfi_be (no-op without /root/reference)."""
    if not os.path.isdir(REF):
        return False
    n = 0
    for sub in SUBDIRS:
        for dirpath, _, files in os.walk(os.path.join(REF, sub)):
            rel = os.path.relpath(dirpath, REF)
            for f in files:
                if f.endswith((".py", ".yml", ".yaml")):
                    dst = os.path.join(STAGE, rel, f)
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(os.path.join(dirpath, f), dst)
                    n += 1
    if verbose:
        print(f"staged {n} reference files into {STAGE}")
    return True


def available():
    return os.path.isfile(os.path.join(STAGE, "models", "Ours", "model_singleframe.py"))


class _Stub(types.ModuleType):
    """Module / object stand-in: any attribute is another stub, calling it returns a stub."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        s = _Stub(self.__name__ + "." + name)
        setattr(self, name, s)
        return s

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


STUB_ROOTS = ("matplotlib", "mpl_toolkits", "open3d", "h5py", "esim_py", "lpips", "thop", "skimage", "IPython",
              "tensorboardX", "tensorboard", "seaborn", "mmcv")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolves `import X.y.z` to a stub for every X in STUB_ROOTS that is not installed."""

    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def install_stubs():
    missing = [r for r in STUB_ROOTS if importlib.util.find_spec(r) is None]
    if missing and not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder(missing))
    # torch.utils.tensorboard imports the `tensorboard` package at module scope
    if importlib.util.find_spec("tensorboard") is None and "torch.utils.tensorboard" not in sys.modules:
        sys.modules["torch.utils.tensorboard"] = _Stub("torch.utils.tensorboard")
    # torchvision.models.utils was removed from torchvision; resnet_3D.py:3 imports load_state_dict_from_url from it
    if "torchvision.models.utils" not in sys.modules:
        try:
            importlib.import_module("torchvision.models.utils")
        except Exception:
            m = types.ModuleType("torchvision.models.utils")
            from torch.hub import load_state_dict_from_url
            m.load_state_dict_from_url = load_state_dict_from_url
            sys.modules["torchvision.models.utils"] = m


_loaded = None


def load():
    """Import the staged, unmodified `models.Ours.model_singleframe` on this repo's shims."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"{STAGE} is missing: run __graft_entry__.build() where /root/reference exists")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import ebfi_be_b200
    ebfi_be_b200.install_shims()
    install_stubs()
    if STAGE not in sys.path:
        sys.path.insert(0, STAGE)
    _loaded = importlib.import_module("models.Ours.model_singleframe")
    return _loaded


def model_args():
    import yaml
    with open(os.path.join(STAGE, "config", "train_ours.yml")) as f:
        cfg = yaml.safe_load(f)
    return cfg["model"]["args"], cfg


def build_model(seed=0, **override):
    """EVFIAutoEx(**train_ours.yml model.args), random init (train_ours.py:221-232)."""
    import torch
    mod = load()
    args, _ = model_args()
    args = dict(args)
    args.update(override)
    torch.manual_seed(seed)
    return mod.EVFIAutoEx(**args)


def reference_wrappers():
    """The reference's unmodified operator wrappers (models/DCNv2/dcn_v2.py, models/FAC/kernelconv2d/KernelConv2D.py),
    imported from the staged tree on this repo's shims."""
    load()
    return importlib.import_module("models.DCNv2.dcn_v2"), importlib.import_module("models.FAC.kernelconv2d.KernelConv2D")


def use_reference_cuda_fac(on):
    """A/B switch for the FAC extension the staged KernelConv2D.py calls: this repo's shim, or the reference's own
    kernels compiled for sm_100a (oracle/_ref/fac_cuda). Returns False when the latter did not travel / load."""
    load()
    kc_mod = importlib.import_module("models.FAC.kernelconv2d.KernelConv2D")
    if not on:
        from ebfi_be_b200.shims import kernelconv2d_cuda as ours
        kc_mod.kernelconv2d_cuda = ours
        return True
    d = os.path.join(ROOT, "oracle", "_ref", "fac_cuda")
    for f in (os.listdir(d) if os.path.isdir(d) else []):
        if f.startswith("kernelconv2d_cuda.") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("kernelconv2d_cuda_ref", os.path.join(d, f))
            # the extension's init symbol is PyInit_kernelconv2d_cuda: load under its own name, keep ours in sys.modules
            spec = importlib.util.spec_from_file_location("kernelconv2d_cuda", os.path.join(d, f))
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
            except ImportError:
                return False
            kc_mod.kernelconv2d_cuda = mod
            return True
    return False


if __name__ == "__main__":
    stage(verbose=True)
    m = build_model()
    print(type(m).__name__, sum(p.numel() for p in m.parameters()), "parameters")
