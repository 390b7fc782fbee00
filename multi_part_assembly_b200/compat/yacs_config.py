"""Minimal stand-in for `yacs.config.CfgNode` (yacs is not installed in this
image).  Supports what the reference's config files and models use: attribute
and item access, nested nodes, `clone`, `freeze`/`defrost`, `get`, `in`,
`items`, `merge_from_list`, None values ("Config")."""
import copy


class CfgNode(dict):
    IMMUTABLE = '__immutable__'

    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(
                f'Attempted to set {name} to {value}, but CfgNode is immutable')
        self[name] = value

    def __setitem__(self, key, value):
        if self.__dict__.get(CfgNode.IMMUTABLE, False):
            raise AttributeError(
                f'Attempted to set {key} to {value}, but CfgNode is immutable')
        super().__setitem__(key, value)

    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def _set_immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

    def freeze(self):
        self._set_immutable(True)

    def defrost(self):
        self._set_immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        new = CfgNode()
        for k, v in self.items():
            dict.__setitem__(new, k, copy.deepcopy(v, memo))
        new.__dict__[CfgNode.IMMUTABLE] = self.is_frozen()
        return new

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            keys = full_key.split('.')
            for k in keys[:-1]:
                node = node[k]
            node[keys[-1]] = v

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            if isinstance(v, CfgNode) and isinstance(self.get(k), CfgNode):
                self[k].merge_from_other_cfg(v)
            else:
                self[k] = copy.deepcopy(v)

    def dump(self, **kwargs):
        import yaml

        def plain(n):
            return {k: plain(v) if isinstance(v, CfgNode) else
                    (list(v) if isinstance(v, tuple) else v) for k, v in n.items()}

        return yaml.safe_dump(plain(self), **kwargs)

    def __str__(self):
        return self.dump()

    __repr__ = dict.__repr__
