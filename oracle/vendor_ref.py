"""Vendor the reference's two MokA layer packages into oracle/_ref/ (build container only).

TEST / BASELINE INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so the CPU baseline
(``bench.py --impl reference``, ``cpu_baseline``), the eager-CUDA comparator and the INTEGRATION.md patch test
would otherwise have to fall back to the oracle's port.  This script copies, byte for byte and *.py only,

  * ``AudioVisualText/peft_hyper``       -> ``oracle/_ref/avt/peft_hyper``   (Linear: tuners/lora.py:277-532)
  * ``VisualText/modified_peft``         -> ``oracle/_ref/vt/peft``          (Linear: tuners/lora/layer.py:389-685;
                                                                              the package imports itself as ``peft``)

and writes ``oracle/_ref/MANIFEST.json`` (source path + sha256 per file).  ``oracle/_ref/`` is git-ignored (no
reference source ever enters the history) but NOT gpurun-ignored, so it travels to the GPU box next to the built
``libmoka.so``.  ``__graft_entry__.build()`` runs this whenever ``/root/reference`` is present.
Nothing in ``moka_b200/`` imports from ``oracle/_ref``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('MOKA_REFERENCE_ROOT', '/root/reference')
DST = os.path.join(HERE, '_ref')
TREES = (('AudioVisualText/peft_hyper', 'avt/peft_hyper'), ('VisualText/modified_peft', 'vt/peft'))


def vendor(ref_root: str = REF, dst: str = DST, quiet: bool = False) -> dict:
    if not os.path.isdir(ref_root):
        raise FileNotFoundError(f'{ref_root} not found: the reference tree only exists in the build container')
    manifest = {}
    for src_rel, dst_rel in TREES:
        src_dir = os.path.join(ref_root, src_rel)
        out_dir = os.path.join(dst, dst_rel)
        if os.path.isdir(out_dir):
            shutil.rmtree(out_dir)
        for base, _dirs, files in os.walk(src_dir):
            for f in sorted(files):
                if not f.endswith('.py'):
                    continue
                s = os.path.join(base, f)
                rel = os.path.relpath(s, src_dir)
                t = os.path.join(out_dir, rel)
                os.makedirs(os.path.dirname(t), exist_ok=True)
                shutil.copyfile(s, t)
                with open(s, 'rb') as fh:
                    manifest[os.path.join(dst_rel, rel)] = dict(source=os.path.join(src_rel, rel),
                                                                sha256=hashlib.sha256(fh.read()).hexdigest())
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as fh:
        json.dump(dict(reference_root=ref_root, files=manifest), fh, indent=1, sort_keys=True)
    if not quiet:
        print(f'vendored {len(manifest)} reference files into {dst}')
    return manifest


if __name__ == '__main__':
    vendor()
    sys.exit(0)
