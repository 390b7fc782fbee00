"""Build libmpa_b200.so in-tree with plain nvcc for sm_100a (no torch headers).

    python multi_part_assembly_b200/csrc/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['mpa_runtime.cu', 'chamfer.cu', 'se3.cu', 'pointnet.cu', 'pointnet_bwd.cu', 'linear.cu', 'knn.cu', 'loss.cu', 'pointnet2.cu']
HEADERS = ['mpa_common.cuh', 'tc05.cuh', os.path.join('..', '..', 'include', 'mpa_b200.h')]
TARGET = os.path.join(HERE, 'libmpa_b200.so')
NVCC = os.environ.get('MPA_NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-std=c++17', '-Xcompiler', '-fPIC', '--use_fast_math=false'
]
FLAGS = [f for f in FLAGS if not f.startswith('--use_fast_math')]


def _source_hash():
    import hashlib
    h = hashlib.sha256(' '.join(FLAGS).encode())
    for d in SOURCES + HEADERS:
        with open(os.path.join(HERE, d), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


STAMP = os.path.join(HERE, 'build', '.srchash')


def _stale():
    """True when the library is missing or was built from other sources.  Content hash,
    not mtimes: the snapshot that carries the tree to a GPU box does not keep them."""
    if not os.path.exists(TARGET) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile when the library is missing or older than a source.  Concurrent callers
    (ranks of one torchrun) serialise on a file lock; the library is written to a
    temporary name and renamed, so a reader never maps a partial file."""
    if not force and not _stale():
        return TARGET
    import fcntl
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    with open(os.path.join(HERE, 'build', '.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():  # another process built it while we waited
            return TARGET
        return _build_locked(verbose)


def _build_locked(verbose):
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
    tmp = TARGET + f'.tmp{os.getpid()}'
    subprocess.check_call([NVCC, '-shared', '-o', tmp] + objs +
                          ['-gencode', 'arch=compute_100a,code=sm_100a'])
    os.replace(tmp, TARGET)
    with open(STAMP, 'w') as f:
        f.write(_source_hash())
    return TARGET


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
