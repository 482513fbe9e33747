"""Build libkgc_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build().

    python kgc-gcn_b200/build.py [--force] [--verbose]

Every .cu under csrc/ is compiled to an object (in parallel) with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
and linked into kgc-gcn_b200/libkgc_b200.so next to this file, so the library travels with the
repo snapshot to the GPU box.  nvcc cross-compiles without a GPU.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libkgc_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
         '--expt-relaxed-constexpr', '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stamp(src):
    h = hashlib.sha1()
    deps = [src] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h'))]
    deps.append(os.path.join(ROOT, 'include', 'kgc_b200.h'))
    for d in deps:
        with open(d, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(ARCH + FLAGS).encode())
    return h.hexdigest()


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
    stamp_file = obj + '.stamp'
    stamp = _stamp(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False, ''
    cmd = [NVCC] + ARCH + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}'.format(src, p.stdout))
    with open(stamp_file, 'w') as f:
        f.write(stamp)
    return obj, True, p.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [r[0] for r in res]
    rebuilt = any(r[1] for r in res)
    if verbose:
        for r in res:
            if r[2]:
                print(r[2])
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', LIB] + objs
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError('link failed:\n' + p.stdout)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(a.force, a.verbose))
