"""Config inheritance, reference API of utils/config_utils.py:6-19."""
import importlib.util
import os


def _load_module(path):
    spec = importlib.util.spec_from_file_location(
        'mpa_cfg_' + os.path.basename(path)[:-3] + f'_{abs(hash(path))}', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def merge_cfg(base_cfg, base_dir, cfg_lst):
    """Merge the `_base_` sub-configs {key: relative .py path} into `base_cfg`
    without overwriting keys the child config already set."""
    for k, v in cfg_lst.items():
        sub_cfg = _load_module(os.path.join(base_dir, v)).get_cfg_defaults()
        if k not in base_cfg:
            base_cfg[k] = sub_cfg
        else:
            for key, value in sub_cfg.items():
                if key not in base_cfg[k]:
                    base_cfg[k][key] = value
    return base_cfg
