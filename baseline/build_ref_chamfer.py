"""Build the REFERENCE's own Chamfer CUDA extension for sm_100a (measurement
baseline only).  The sources are read where they lie under
/root/reference, patched IN MEMORY for the torch 2.x C++ API (kernel bodies and
launch configuration untouched) and compiled into baseline/_ref/ (git-ignored):

  THC/THC.h include           -> c10/cuda/CUDAException.h + ATen/cuda/CUDAContext.h
  x.type().is_cuda()          -> x.is_cuda()
  THCudaCheck(...)            -> C10_CUDA_CHECK(...)
  THArgCheck(c, 1, msg)       -> TORCH_CHECK(c, msg)
  .data<T>()                  -> .data_ptr<T>()
  CHECK_EQ                    -> TORCH_CHECK_EQ
  at::zeros(sz, t.type()...)  -> at::zeros(sz, t.options()...)

    python baseline/build_ref_chamfer.py      # build container only
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/multi_part_assembly/utils/chamfer/cuda'
OUT = os.path.join(HERE, '_ref')
BUILD = os.path.join(OUT, '_build')


def patched_kernel():
    s = open(os.path.join(SRC, 'chamfer_kernel.cu')).read()
    s = s.replace('#include <THC/THC.h>',
                  '#include <c10/cuda/CUDAException.h>\n#include <ATen/cuda/CUDAContext.h>')
    s = s.replace('x.type().is_cuda()', 'x.is_cuda()')
    s = s.replace('THCudaCheck(', 'C10_CUDA_CHECK(')
    s = re.sub(r'THArgCheck\((.*?), 1, (".*?")\);', r'TORCH_CHECK(\1, \2);', s)
    s = re.sub(r'\.data<', '.data_ptr<', s)
    s = s.replace('CHECK_EQ(', 'TORCH_CHECK_EQ(')
    s = s.replace('xyz1.type().toScalarType(at::kLong)', 'xyz1.options().dtype(at::kLong)')
    s = s.replace('xyz2.type().toScalarType(at::kLong)', 'xyz2.options().dtype(at::kLong)')
    s = re.sub(r'(at::zeros\(\{[^}]*\}, \w+)\.type\(\)\)', r'\1.options())', s)
    return s


def main():
    if not os.path.isdir(SRC):
        print('reference tree not present; nothing to build')
        return 0
    os.makedirs(BUILD, exist_ok=True)
    open(os.path.join(BUILD, 'chamfer_kernel.cu'), 'w').write(patched_kernel())
    shutil.copy(os.path.join(SRC, 'chamfer.cpp'), os.path.join(BUILD, 'chamfer.cpp'))
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ['CC'] = '/usr/bin/gcc'
    os.environ['CXX'] = '/usr/bin/g++'
    from torch.utils.cpp_extension import load
    load(name='chamfer_cuda', sources=[os.path.join(BUILD, 'chamfer.cpp'),
                                       os.path.join(BUILD, 'chamfer_kernel.cu')],
         build_directory=BUILD, extra_cuda_cflags=['-O2'], verbose=False, is_python_module=True)
    so = [f for f in os.listdir(BUILD) if f.startswith('chamfer_cuda') and f.endswith('.so')][0]
    shutil.copy(os.path.join(BUILD, so), os.path.join(OUT, 'chamfer_cuda.so'))
    print('built', os.path.join(OUT, 'chamfer_cuda.so'))
    return 0


if __name__ == '__main__':
    sys.exit(main())
