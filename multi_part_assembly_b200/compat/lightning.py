"""Minimal `pytorch_lightning` surface for the reference's BaseModel hooks
(PL 1.6 API: training_step / validation_step / *_epoch_end / log_dict /
local_rank / trainer.profiler).  pytorch_lightning is not
installed in this image; the Trainer runtime itself is out of scope (SURVEY 8,
L3).  One process per GPU: rank and world size come from torch.distributed."""
import os

import torch
import torch.nn as nn


class _Profiler:
    """Stands in for `trainer.profiler`: base_model.py:139-144 reads the last
    recorded duration of the '...prepare_data...' action."""

    def __init__(self):
        self.recorded_durations = {'[TrainingEpochLoop].prepare_data': [0.0]}


class Trainer:
    """Just enough of a Trainer for `LightningModule.trainer` lookups."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.profiler = _Profiler()
        self.logged = {}
        self.current_epoch = 0
        self.global_step = 0


class LightningModule(nn.Module):

    def __init__(self):
        super().__init__()
        self.trainer = None
        self._logged = {}

    @property
    def local_rank(self):
        return int(os.environ.get('LOCAL_RANK', 0))

    @property
    def global_rank(self):
        return int(os.environ.get('RANK', 0))

    def log_dict(self, dictionary, sync_dist=False, **kwargs):
        """Record scalars; with sync_dist=True average them over the process
        group (PL semantics used at base_model.py:84)."""
        out = {}
        for k, v in dictionary.items():
            if sync_dist and torch.distributed.is_available() and \
                    torch.distributed.is_initialized():
                t = v.detach().clone() if isinstance(v, torch.Tensor) else \
                    torch.tensor(float(v))
                if torch.distributed.get_backend() == 'nccl':
                    t = t.cuda()
                torch.distributed.all_reduce(t)
                v = t / torch.distributed.get_world_size()
            out[k] = v
        self._logged.update(out)
        if self.trainer is not None:
            self.trainer.logged.update(out)

    def log(self, name, value, **kwargs):
        self.log_dict({name: value}, **kwargs)


class Callback:
    pass


class _Callbacks:
    Callback = Callback
