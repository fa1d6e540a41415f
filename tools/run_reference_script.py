"""Runs the reference's UNMODIFIED ``detect.py`` (staged by tools/make_baseline_ref.py into baseline/_ref/) in this process,
either on its own stock code path or on the native backend after ``repyolo_b200.compat.install()``.

    python tools/run_reference_script.py --ref baseline/_ref [--native] -- --weights W.pt --source DIR --device 0 --save-txt ...

The only environment shims are import stubs for the plotting packages this image does not have (matplotlib, seaborn:
utils/plots.py:11-15 imports them at module scope; detect.py only uses plot_one_box, which is cv2) -- detect.py and every
module it imports are executed as they are.  ``--native``: the reference modules are imported, ``compat.install()`` rebinds
attempt_load / TracedModel / non_max_suppression / scale_coords inside them, and detect.py's own ``from ... import`` lines then
pick the native callables up (the binding a maintainer would add is those two lines).
Prints one JSON line: {"labels": {file: text}, "native": bool, "patched": [...]}.
"""
import json
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_plot_modules():
    for name in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.rc = mod.use = lambda *a, **k: None
        sys.modules[name] = mod
    if isinstance(sys.modules.get('matplotlib'), types.ModuleType) and not hasattr(sys.modules['matplotlib'], 'pyplot'):
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


def write_checkpoint(ref_dir, path, seed=0):
    """The pickled checkpoint detect.py expects (train.py's format: experimental.py:242-243 takes ckpt['ema'] when present):
    the REFERENCE's own models.yolo.Model built from its yaml, holding the oracle's synthetic calibrated weights of ``seed``,
    parameters frozen like an EMA copy.  ~110 MB, so it is written where it is used (tmp dir) instead of travelling."""
    import copy
    import logging
    import torch
    ref_dir = os.path.abspath(ref_dir)
    stub_plot_modules()
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    if ROOT not in sys.path:
        sys.path.append(ROOT)
    logging.disable(logging.CRITICAL)
    from models.yolo import Model
    from oracle import repyolo_oracle as O
    _, _, sd, _ = O.make_model(seed=seed, mode='calibrated')
    m = Model(os.path.join(ref_dir, 'cfg', 'training', 'Rep-YOLO.yaml'), ch=3, nc=1)
    m.load_state_dict(sd, strict=True)
    m.names = ['person']
    ema = copy.deepcopy(m).eval()
    for p in ema.parameters():
        p.requires_grad_(False)
    torch.save({'model': None, 'ema': ema, 'epoch': -1}, path)
    logging.disable(logging.NOTSET)
    return path


def run_detect(ref_dir, argv, native):
    """Executes <ref_dir>/detect.py as __main__ with ``argv``; returns {'labels': {name: text}, 'patched': [...]}."""
    ref_dir = os.path.abspath(ref_dir)
    stub_plot_modules()
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    if ROOT not in sys.path:
        sys.path.append(ROOT)
    patched = []
    if not native:
        # the reference predates torch 2.6: its torch.load(w) of a pickled nn.Module needs the old weights_only=False default
        os.environ['TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD'] = '1'
    if native:
        import repyolo_b200 as R
        import models.experimental, utils.general, utils.torch_utils      # noqa: E401,F401  (the reference's, from ref_dir)
        patched = R.compat.install()
    project = name = None
    for i, a in enumerate(argv):
        if a == '--project':
            project = argv[i + 1]
        if a == '--name':
            name = argv[i + 1]
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [os.path.join(ref_dir, 'detect.py')] + list(argv)
    os.makedirs(os.path.join(ref_dir, '_run'), exist_ok=True)
    os.chdir(os.path.join(ref_dir, '_run'))                # TracedModel (stock path) drops traced_model.pt into the cwd
    try:
        runpy.run_path(os.path.join(ref_dir, 'detect.py'), run_name='__main__')
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    labels = {}
    ldir = os.path.join(project or 'runs/detect', name or 'exp', 'labels')
    if os.path.isdir(ldir):
        for f in sorted(os.listdir(ldir)):
            with open(os.path.join(ldir, f)) as fh:
                labels[f] = fh.read()
    return {'labels': labels, 'native': bool(native), 'patched': [list(p) for p in patched]}


def main():
    args = sys.argv[1:]
    split = args.index('--')
    own, rest = args[:split], args[split + 1:]
    ref = own[own.index('--ref') + 1]
    if '--make-ckpt' in own:                              # --make-ckpt PATH[,PATH2]: seeds 0, 1, ...
        for seed, path in enumerate(own[own.index('--make-ckpt') + 1].split(',')):
            write_checkpoint(ref, path, seed)
    out = run_detect(ref, rest, '--native' in own)
    print('RESULT ' + json.dumps(out))


if __name__ == '__main__':
    main()
