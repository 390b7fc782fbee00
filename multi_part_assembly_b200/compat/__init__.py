"""Drop-in glue: lets the reference's unmodified config files (which do
`from yacs.config import CfgNode` and `from multi_part_assembly.utils import
merge_cfg`) and user code written against `multi_part_assembly` run on this
package.  Call `install()` once (idempotent)."""
import importlib
import sys
import types


def install():
    import multi_part_assembly_b200 as pkg
    # 1. yacs / pytorch_lightning stand-ins, only when the real ones are absent
    try:
        importlib.import_module('yacs.config')
    except ImportError:
        from . import yacs_config
        yacs = types.ModuleType('yacs')
        yacs.config = yacs_config
        sys.modules['yacs'] = yacs
        sys.modules['yacs.config'] = yacs_config
    try:
        importlib.import_module('pytorch_lightning')
    except ImportError:
        from . import lightning
        pl = types.ModuleType('pytorch_lightning')
        pl.LightningModule = lightning.LightningModule
        pl.Trainer = lightning.Trainer
        pl.Callback = lightning.Callback
        pl.callbacks = lightning._Callbacks
        sys.modules['pytorch_lightning'] = pl
    # 2. the package under the reference's name
    if 'multi_part_assembly' not in sys.modules:
        sys.modules['multi_part_assembly'] = pkg
        for sub in ('utils', 'models', 'datasets'):
            mod = importlib.import_module(f'multi_part_assembly_b200.{sub}')
            sys.modules[f'multi_part_assembly.{sub}'] = mod
    return pkg
