"""Drop-in wiring for the reference's own scripts (detect.py / test.py / hubconf.py run unchanged).

``install()`` rebinds, inside the reference's modules when they are importable (reference repo on ``sys.path``):
    models.experimental.attempt_load      -> attempt_load        (models/experimental.py:237-260)
    utils.torch_utils.TracedModel         -> TracedModel         (utils/torch_utils.py:343-377; pass-through, nothing to trace)
    utils.general.non_max_suppression     -> non_max_suppression (utils/general.py:953)
and the copies that ``from x import y`` left in already-imported modules (detect, test, models.common).
Nothing here is needed when the package is used directly (``repyolo_b200.Model`` / ``non_max_suppression``).
"""
from __future__ import annotations

import sys

import torch
import torch.nn as nn

from .model import Model
from . import preproc as _preproc
from .nms import non_max_suppression


def from_reference(ref_model) -> Model:
    """Native model from an (unfused) reference ``models.yolo.Model`` instance: same yaml, same state_dict keys."""
    m = Model(dict(ref_model.yaml))
    m.load_state_dict({k: v.float() for k, v in ref_model.state_dict().items()}, strict=True)
    if hasattr(ref_model, 'names'):
        m.names = ref_model.names
    return m


class Ensemble(nn.ModuleList):
    """models/experimental.py:69-81: concatenates the members' predictions before NMS."""

    def forward(self, x, augment=False):
        y = [module(x, augment=augment)[0] for module in self]
        return torch.cat(y, 1), None


def attempt_load(weights, map_location=None):
    """models/experimental.py:237-260 with the native model substituted for ``ckpt['ema' or 'model'].float().fuse().eval()``.
    The checkpoint is the reference's pickled nn.Module, so the reference package must be importable to unpickle it."""
    model = Ensemble()
    for w in weights if isinstance(weights, list) else [weights]:
        ckpt = torch.load(w, map_location='cpu', weights_only=False)
        ref = ckpt['ema' if ckpt.get('ema') else 'model']
        m = from_reference(ref).fuse().eval()
        if map_location is not None and str(map_location) != 'cpu':
            m = m.to(map_location)
        model.append(m)
    if len(model) == 1:
        return model[-1]
    for k in ('names', 'stride'):
        setattr(model, k, getattr(model[-1], k))
    return model


class TracedModel(nn.Module):
    """utils/torch_utils.py:343-377 stand-in: the native engine is already a static plan, there is nothing to trace."""

    def __init__(self, model=None, device=None, img_size=(640, 640)):
        super().__init__()
        self.stride, self.names, self.model = model.stride, model.names, model
        self.detect_layer = model.model[-1]

    def forward(self, x, augment=False, profile=False):
        return self.model(x, augment=augment)


def install() -> list:
    """Rebinds the reference's entry points to the native ones; returns the list of (module, attribute) patched."""
    import importlib
    patched = []
    targets = {'attempt_load': attempt_load, 'TracedModel': TracedModel, 'non_max_suppression': non_max_suppression,
               'Ensemble': Ensemble}
    for modname in ('models.experimental', 'utils.torch_utils', 'utils.general'):
        try:
            importlib.import_module(modname)
        except Exception:
            continue
    ref_general = sys.modules.get('utils.general')
    if ref_general is not None and hasattr(ref_general, 'scale_coords') and not getattr(ref_general.scale_coords, '_ry_native', False):
        ref_scale_coords = ref_general.scale_coords

        def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
            # detections (fp32 CUDA tensors, detect.py:114 / test.py:141) go to the native kernel; anything else (host label
            # arrays of the dataset code) stays with the reference's own function
            if torch.is_tensor(coords) and coords.is_cuda and coords.dtype == torch.float32 and coords.dim() == 2:
                return _preproc.scale_coords(img1_shape, coords, img0_shape, ratio_pad)
            return ref_scale_coords(img1_shape, coords, img0_shape, ratio_pad)

        scale_coords._ry_native = True
        scale_coords.__module__ = 'repyolo_b200.compat'
        targets['scale_coords'] = scale_coords
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname.split('.')[0] in ('models', 'utils', 'detect', 'test', 'hubconf', '__main__')):
            continue
        for attr, fn in targets.items():
            if hasattr(mod, attr) and getattr(mod, attr) is not fn and getattr(getattr(mod, attr), '__module__', '').split('.')[0] in (
                    'models', 'utils'):
                setattr(mod, attr, fn)
                patched.append((modname, attr))
    return patched
