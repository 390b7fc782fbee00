"""Cosine learning-rate schedule with linear warm-up, stepped per epoch
(reference utils/lr.py:26-125; optimiser side, outside the fwd+loss hot path)."""
import math

from torch.optim.lr_scheduler import _LRScheduler


class CosineAnnealingWarmupRestarts(_LRScheduler):

    def __init__(self, optimizer, first_cycle_steps, cycle_mult=1., max_lr=0.1, min_lr=0.001,
                 warmup_steps=0, gamma=1., last_epoch=-1):
        assert warmup_steps < first_cycle_steps
        self.first_cycle_steps = first_cycle_steps
        self.cycle_mult = cycle_mult
        self.base_max_lr = max_lr
        self.max_lr = max_lr
        self.min_lr = min_lr
        self.warmup_steps = warmup_steps
        self.gamma = gamma
        self.cur_cycle_steps = first_cycle_steps
        self.cycle = 0
        self.step_in_cycle = last_epoch
        super().__init__(optimizer, last_epoch)
        self.base_lrs = []
        for group in self.optimizer.param_groups:
            group['lr'] = self.min_lr
            self.base_lrs.append(self.min_lr)

    def get_lr(self):
        if self.step_in_cycle == -1:
            return self.base_lrs
        if self.step_in_cycle < self.warmup_steps:
            return [(self.max_lr - b) * self.step_in_cycle / self.warmup_steps + b
                    for b in self.base_lrs]
        frac = (self.step_in_cycle - self.warmup_steps) / (self.cur_cycle_steps - self.warmup_steps)
        return [b + (self.max_lr - b) * (1 + math.cos(math.pi * frac)) / 2 for b in self.base_lrs]

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
            self.step_in_cycle += 1
            if self.step_in_cycle >= self.cur_cycle_steps:
                self.cycle += 1
                self.step_in_cycle -= self.cur_cycle_steps
                self.cur_cycle_steps = int((self.cur_cycle_steps - self.warmup_steps) *
                                           self.cycle_mult) + self.warmup_steps
        elif epoch >= self.first_cycle_steps:
            if self.cycle_mult == 1.:
                self.step_in_cycle = epoch % self.first_cycle_steps
                self.cycle = epoch // self.first_cycle_steps
            else:
                n = int(math.log(epoch / self.first_cycle_steps * (self.cycle_mult - 1) + 1,
                                 self.cycle_mult))
                self.cycle = n
                self.step_in_cycle = epoch - int(
                    self.first_cycle_steps * (self.cycle_mult**n - 1) / (self.cycle_mult - 1))
                self.cur_cycle_steps = self.first_cycle_steps * self.cycle_mult**n
        else:
            self.cur_cycle_steps = self.first_cycle_steps
            self.step_in_cycle = epoch
        self.max_lr = self.base_max_lr * (self.gamma**self.cycle)
        self.last_epoch = math.floor(epoch)
        for group, lr in zip(self.optimizer.param_groups, self.get_lr()):
            group['lr'] = lr
