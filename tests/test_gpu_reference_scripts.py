"""-m gpu: the reference's own UNMODIFIED detect.py on the native backend.

``tools/make_baseline_ref.py`` (build container) stages models/ utils/ cfg/ detect.py + dog.jpg of the reference into the
git-ignored ``baseline/_ref/``; it travels to the GPU box.  Here: a pickled checkpoint of the REFERENCE's ``models.yolo.Model``
(the format experimental.py:242-243 loads) is written on the box, ``detect.py --device 0`` runs in a subprocess after
``repyolo_b200.compat.install()`` -- default options, i.e. WITH its TracedModel step -- and the label files it saves must equal,
character for character, what the direct API (``preprocess`` -> ``Model`` -> ``non_max_suppression`` -> ``scale_coords``)
produces for the same picture and weights.  Then the same with two checkpoints (``Ensemble``, experimental.py:69-81).
"""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT
from oracle import repyolo_oracle as O

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, 'baseline', '_ref')
IMG_DIR = os.path.join(REF, 'inference', 'images')


def _run_detect(tmp, name, weights, make=None, extra=()):
    cmd = [sys.executable, os.path.join(ROOT, 'tools', 'run_reference_script.py'), '--ref', REF, '--native']
    if make:
        cmd += ['--make-ckpt', ','.join(make)]
    cmd += ['--', '--weights', *weights, '--source', IMG_DIR, '--device', '0', '--save-txt', '--save-conf', '--nosave',
            '--project', os.path.join(tmp, 'out'), '--name', name, '--exist-ok', *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')][-1]
    return json.loads(line[len('RESULT '):]), r.stdout


def _direct_labels(models, im0):
    """detect.py:70-131 through the package's own API; formatting = detect.py:121-126."""
    import repyolo_b200 as R
    img, _, _ = R.preprocess(im0, 640, 32)                       # LoadImages: letterbox(auto) + BGR->RGB + HWC->CHW
    x = img.float()
    x /= 255.0
    x = x.unsqueeze(0)
    net = models[0] if len(models) == 1 else R.Ensemble(models)
    pred = net(x, augment=False)[0]
    det = R.non_max_suppression(pred, 0.25, 0.45, classes=None, agnostic=False)[0]
    lines = []
    if len(det):
        det[:, :4] = R.scale_coords(x.shape[2:], det[:, :4], im0.shape).round()
        gn = torch.tensor(im0.shape)[[1, 0, 1, 0]]
        for *xyxy, conf, cls in reversed(det):
            b = torch.tensor(xyxy).view(1, 4)
            xywh = torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1)   # general.py:256-262
            xywh = (xywh / gn).view(-1).tolist()
            line = (cls, *xywh, conf)
            lines.append(('%g ' * len(line)).rstrip() % line + '\n')
    return ''.join(lines), int(len(det))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'detect.py')), reason='baseline/_ref not staged (tools/make_baseline_ref.py)')
def test_unmodified_detect_py_runs_on_native_backend(tmp_path):
    import cv2
    import repyolo_b200 as R
    tmp = str(tmp_path)
    wa, wb = os.path.join(tmp, 'repyolo_seed0.pt'), os.path.join(tmp, 'repyolo_seed1.pt')
    res, log = _run_detect(tmp, 'single', [wa], make=[wa, wb])
    assert res['native'] and ['models.experimental', 'attempt_load'] in res['patched'], res['patched']
    assert ['utils.general', 'non_max_suppression'] in res['patched'] and ['utils.torch_utils', 'TracedModel'] in res['patched']
    assert 'dog.txt' in res['labels'], (res, log[-2000:])

    models = []
    for seed in (0, 1):
        _, _, sd, _ = O.make_model(seed=seed, mode='calibrated')
        m = R.Model()
        m.load_state_dict(sd, strict=True)
        m.names = ['person']
        models.append(m.fuse().eval().to('cuda:0'))
    im0 = cv2.imread(os.path.join(IMG_DIR, 'dog.jpg'))
    want, n = _direct_labels(models[:1], im0)
    assert n > 0
    assert res['labels']['dog.txt'] == want, (res['labels']['dog.txt'][:400], want[:400])

    # Ensemble of two checkpoints: predictions concatenated before NMS (experimental.py:69-81, 254-260)
    # (--no-trace: the reference's own TracedModel reads `model.model[-1]`, which an Ensemble does not have -- torch_utils.py:357)
    res2, log2 = _run_detect(tmp, 'ensemble', [wa, wb], extra=('--no-trace',))
    want2, n2 = _direct_labels(models, im0)
    assert n2 > 0 and res2['labels']['dog.txt'] == want2
