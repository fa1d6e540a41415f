"""Stages the files the reference's own ``detect.py`` needs into git-ignored ``baseline/_ref/``, so the
GPU box (where /root/reference does not exist) can run the UNMODIFIED script on the native backend.

Run in the build container:  python tools/make_baseline_ref.py

What it does
  1. copies models/, utils/, cfg/, detect.py and the one sample picture of the reference (deploy/.../dog.jpg) from
     /root/reference into baseline/_ref/ -- byte copies, never committed (baseline/_ref/ is in .gitignore)
  2. builds the reference's own ``models.yolo.Model(cfg/training/Rep-YOLO.yaml)`` on the CPU with the oracle's synthetic
     weights (seed 0, calibrated init) and writes the pickled checkpoint detect.py expects into a temp dir
     (tools/run_reference_script.py::write_checkpoint; ~110 MB each, so the GPU test writes its own on the box)
  3. runs the reference detect.py itself on the CPU (its own attempt_load / Model / NMS) on dog.jpg and stores the labels it
     writes under baseline/_ref/expected_cpu/ (fp32 reference detections, for information: random-weight Rep-YOLO is chaotic
     under bf16, so the GPU test compares detect.py-on-native with the native direct API, not with these)
"""
import copy
import os
import shutil
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')
sys.path.insert(0, ROOT)


def main():
    os.makedirs(DST, exist_ok=True)
    for d in ('models', 'utils', 'cfg'):
        shutil.copytree(os.path.join(REF, d), os.path.join(DST, d), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    shutil.copy2(os.path.join(REF, 'detect.py'), os.path.join(DST, 'detect.py'))
    os.makedirs(os.path.join(DST, 'inference', 'images'), exist_ok=True)
    shutil.copy2(os.path.join(REF, 'deploy', 'triton-inference-server', 'data', 'dog.jpg'),
                 os.path.join(DST, 'inference', 'images', 'dog.jpg'))

    import tempfile
    from tools.run_reference_script import run_detect, write_checkpoint
    with tempfile.TemporaryDirectory() as tmp:
        w = write_checkpoint(DST, os.path.join(tmp, 'repyolo_seed0.pt'), seed=0)
        # the reference itself, CPU, its own code path end to end
        out = run_detect(DST, ['--weights', w, '--source', os.path.join(DST, 'inference', 'images'), '--device', 'cpu', '--save-txt',
                               '--save-conf', '--nosave', '--no-trace', '--project', os.path.join(DST, 'expected_cpu'), '--name', 'exp',
                               '--exist-ok'], native=False)
    print('reference CPU run:', out)


if __name__ == '__main__':
    main()
