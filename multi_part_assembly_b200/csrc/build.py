"""Build libmpa_b200.so in-tree with plain nvcc for sm_100a (no torch headers).

    python multi_part_assembly_b200/csrc/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['mpa_runtime.cu', 'chamfer.cu', 'se3.cu', 'pointnet.cu', 'pointnet_bwd.cu', 'linear.cu', 'knn.cu', 'loss.cu']
HEADERS = ['mpa_common.cuh', 'tc05.cuh', os.path.join('..', '..', 'include', 'mpa_b200.h')]
TARGET = os.path.join(HERE, 'libmpa_b200.so')
NVCC = os.environ.get('MPA_NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-std=c++17', '-Xcompiler', '-fPIC', '--use_fast_math=false'
]
FLAGS = [f for f in FLAGS if not f.startswith('--use_fast_math')]


def _stale():
    if not os.path.exists(TARGET):
        return True
    t = os.path.getmtime(TARGET)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return TARGET
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, 'build', s.replace('.cu', '.o'))
        os.makedirs(os.path.dirname(o), exist_ok=True)
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
            ['-c', os.path.join(HERE, s), '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- {s}\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed building libmpa_b200.so')
    subprocess.check_call([NVCC, '-shared', '-o', TARGET] + objs +
                          ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return TARGET


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
