"""Import the UNMODIFIED reference from /root/reference on the CPU (this container only).

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py to produce tests/golden/*.pt and by
nothing else: /root/reference does not exist on the GPU box, so no test, smoke() or bench
imports this module at run time.

The reference's inference arithmetic needs only torch + torchvision, but its modules import
training / visualisation packages that are not installed (pytorch_lightning, matplotlib,
dex_ycb_toolkit, manopth, pycocotools, ...).  They are replaced by MagicMock packages; the
two ImageNet downloads (fcos_utils/fcos.py:476, a2j/resnet.py:196) are neutralised.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

REFERENCE_ROOT = "/root/reference"
_STUBBED = ("pytorch_lightning", "matplotlib", "dex_ycb_toolkit", "manopth", "pycocotools", "sklearn",
            "roi_data_layer", "zmq", "tomlkit", "easydict", "plotly", "cv2", "wandb", "model")


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def __init__(self, names):
        self.names = set(names)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


_installed = False


def install():
    """Idempotently make ``import fcos_utils.fcos`` etc. resolve to the reference."""
    global _installed
    if _installed:
        return
    import torch
    import torch.nn as nn
    import torch.utils.model_zoo
    import torchvision

    import importlib.util
    missing = [n for n in _STUBBED if n == "pytorch_lightning" or importlib.util.find_spec(n) is None]
    sys.meta_path.insert(0, _StubFinder(missing))
    # a fake pytorch_lightning with real base classes so class statements in the reference work
    pl = types.ModuleType("pytorch_lightning")
    pl.__path__ = []

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    class LightningDataModule:
        pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    sys.modules["pytorch_lightning"] = pl

    torch.utils.model_zoo.load_url = lambda *a, **k: None
    torchvision.models._api.WeightsEnum.get_state_dict = lambda *a, **k: None
    _orig = nn.Module.load_state_dict

    def _load_state_dict(self, state_dict, *a, **k):
        if state_dict is None:
            return None
        return _orig(self, state_dict, *a, **k)

    nn.Module.load_state_dict = _load_state_dict
    # the product package mirrors the reference's module paths; make sure the reference wins here
    sys.path[:] = [p for p in sys.path if "handnet-pipeline_b200" not in p]
    for name in list(sys.modules):
        if name.split(".")[0] in ("fcos_utils", "a2j", "handnet_pipeline", "utils", "datasets3d"):
            del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
