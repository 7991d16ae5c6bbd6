"""In-tree build of libmoka.so (nvcc, sm_100a only).  The built library travels to the GPU
box with the repo snapshot; nothing is JIT-compiled at run time."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmoka.so')
STAMP = os.path.join(HERE, '.libmoka.stamp')

SOURCES = ['moka_abi.cu', 'moka_simt.cu', 'moka_meta.cu', 'moka_optim.cu', 'moka_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']
# --use_fast_math only for the translation units that hold the attention softmax (exp2f / division: errors
# ~1e-7 rel, inside the fp32 bar).  NOT for moka_optim.cu: AdamW keeps IEEE division / sqrt and denormals on the
# fp32 master weights and moments, RoPE keeps the full-range sincosf.
FAST_MATH = {'moka_simt.cu', 'moka_tc.cu'}


def _nvcc() -> str:
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, 'include', 'moka.h')]
    for f in files:
        with open(f, 'rb') as fh:
            h.update(f.encode()); h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS + sorted(FAST_MATH)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into moka_b200/libmoka.so.  Rebuilds only when a source changed."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace('.cu', '.o'))
        cmd = [_nvcc(), *NVCC_FLAGS, *(['--use_fast_math'] if src in FAST_MATH else []),
               '-I', os.path.join(ROOT, 'include'), '-I', CSRC, '-c',
               os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'== {src}\n{out}')
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f'nvcc failed on {src}')
    with open(os.path.join(bdir, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    cmd = [_nvcc(), '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.run(cmd, check=True)
    with open(STAMP, 'w') as f:
        f.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
