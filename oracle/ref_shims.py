"""Import the UNMODIFIED reference package from /root/reference on CPU.

TEST INFRASTRUCTURE, build container only (the reference tree does not travel
to the GPU box).  Used by oracle/make_golden.py to produce tests/golden/*.npz
and by tests that pin the oracle against the reference when it is present.

The reference needs third-party modules that are not installed here
(SURVEY.md 8c / appendix C); this installs stand-ins into sys.modules:
  pytorch3d.transforms  <- restated semantics (oracle/torch_ref.py + the
                           conversion helpers of the product package)
  pytorch_lightning     <- nn.Module based LightningModule (product compat)
  yacs.config           <- dict based CfgNode (product compat)
  chamfer_cuda          <- the C oracle (the reference has no CPU Chamfer)
  pointnet2_ops, pyntcloud, trimesh, wandb <- empty stubs
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = '/root/reference'


def available(root=None):
    return os.path.isdir(os.path.join(root or REFERENCE_ROOT, 'multi_part_assembly'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install(root=None, cuda_chamfer=False):
    """Returns the imported reference package `multi_part_assembly`.

    root: directory holding the package (default /root/reference; the GPU
    baseline tool passes baseline/_ref, the pip-installed unmodified copy).
    cuda_chamfer: use the reference's own compiled `chamfer_cuda` extension
    (must be importable) instead of the CPU oracle stand-in."""
    global REFERENCE_ROOT
    if root is not None:
        REFERENCE_ROOT = root
    if not available():
        raise RuntimeError('reference tree not present')
    if 'multi_part_assembly' in sys.modules and \
            getattr(sys.modules['multi_part_assembly'], '__file__', '').startswith(REFERENCE_ROOT):
        return sys.modules['multi_part_assembly']
    assert 'multi_part_assembly' not in sys.modules, \
        'the product alias is installed in this process; use a fresh interpreter'
    from . import torch_ref, cpu
    from multi_part_assembly_b200.utils import rotation_conversions as rc
    from multi_part_assembly_b200.compat import yacs_config, lightning

    p3d = _stub('pytorch3d')
    tr = _stub(
        'pytorch3d.transforms',
        quaternion_invert=torch_ref.quaternion_invert,
        quaternion_apply=torch_ref.quaternion_apply,
        quaternion_raw_multiply=torch_ref.quaternion_raw_multiply,
        random_quaternions=rc.random_quaternions,
        matrix_to_quaternion=rc.matrix_to_quaternion,
        matrix_to_axis_angle=rc.matrix_to_axis_angle,
        quaternion_to_matrix=rc.quaternion_to_matrix,
        quaternion_to_axis_angle=rc.quaternion_to_axis_angle,
        axis_angle_to_quaternion=rc.axis_angle_to_quaternion,
        axis_angle_to_matrix=rc.axis_angle_to_matrix,
        quaternion_multiply=rc.quaternion_multiply,
        rotation_6d_to_matrix=rc.rotation_6d_to_matrix)
    p3d.transforms = tr

    yacs = _stub('yacs')
    yacs.config = yacs_config
    sys.modules['yacs.config'] = yacs_config

    pl = _stub('pytorch_lightning', LightningModule=lightning.LightningModule,
               Trainer=lightning.Trainer, Callback=lightning.Callback)
    pl.callbacks = lightning._Callbacks

    def chamfer_forward(xyz1, xyz2):
        d1, i1, d2, i2 = cpu.chamfer_forward(xyz1.detach().numpy(), xyz2.detach().numpy())
        return [torch.from_numpy(a) for a in (d1, i1, d2, i2)]

    def chamfer_backward(g1, g2, xyz1, xyz2, idx1, idx2):
        a, b = cpu.chamfer_backward(g1.numpy(), g2.numpy(), xyz1.detach().numpy(),
                                    xyz2.detach().numpy(), idx1.numpy(), idx2.numpy())
        return [torch.from_numpy(a), torch.from_numpy(b)]

    if not cuda_chamfer:
        _stub('chamfer_cuda', chamfer_forward=chamfer_forward, chamfer_backward=chamfer_backward)

    class _Missing:

        def __init__(self, *a, **k):
            raise RuntimeError('stubbed third-party class')

    _stub('pointnet2_ops')
    _stub('pointnet2_ops.pointnet2_modules', PointnetSAModule=_Missing,
          PointnetSAModuleMSG=_Missing)
    _stub('pyntcloud', PyntCloud=_Missing)
    _stub('trimesh')
    if importlib.util.find_spec('wandb') is None:
        _stub('wandb')

    sys.path.insert(0, REFERENCE_ROOT)
    ref = importlib.import_module('multi_part_assembly')
    assert ref.__file__.startswith(REFERENCE_ROOT)
    # the reference asserts .is_cuda in its autograd Function (chamfer.py:18);
    # swap in the CPU oracle Function (same forward/backward arithmetic)
    if not cuda_chamfer:
        ch = importlib.import_module('multi_part_assembly.utils.chamfer.chamfer')
        ch.ChamferDistanceFunction = torch_ref._Chamfer
    return ref


def load_reference_test_functions():
    """bpdist2 / nn_distance_torch from the reference's own test file
    (utils/chamfer/test_chamfer.py:8-31), executed from the file where it lies
    (the module itself cannot be imported: it needs chamfer_cuda and runs a
    test at import time, :136)."""
    import ast
    path = os.path.join(REFERENCE_ROOT, 'multi_part_assembly/utils/chamfer/test_chamfer.py')
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and
            n.name in ('bpdist2', 'nn_distance_torch')]
    ns = {'torch': torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, 'exec'), ns)
    return ns['bpdist2'], ns['nn_distance_torch']
