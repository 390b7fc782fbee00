"""CUDA-graph execution of a model step.

The reference drives every step from Python through ~150 small kernel
launches; on a B200 the forward+loss of a 32-shape batch is ~1-3 ms of GPU
work, less than the host needs to issue it.  `GraphedStep` captures
`BaseModel.forward_pass` (forward + all loss terms) once for a fixed batch
shape into a CUDA graph and replays it per batch: inputs are copied into
static device buffers (from pinned host memory when the batch is on the
host), one graph launch, and the loss dict is read from static outputs.

The capture runs without autograd recording (forward + loss only, the scope
of BASELINE.json's metric); training steps use the eager path.
"""
import torch


def _check_capturable(model):
    """A CUDA graph replays fixed device work: anything the step draws from the CPU
    generator or reads back on the host would be frozen into the graph.  The semantic
    (PartNet) models do both -- `torch.randperm` per matching group and a host read of
    `match_ids` (models/modules/base_model.py `_match_parts`), `torch.randn` pose noise
    (models/modules/regressor.py) -- so Min-of-N sampling would silently repeat one
    sample.  They take the eager path."""
    if getattr(model, 'semantic', False):
        raise ValueError('Graphed steps do not support semantic (Hungarian-matching) models: '
                         'the matching draws CPU random numbers and reads match_ids on the host')
    for m in model.modules():
        if getattr(m, 'noise_dim', 0):
            raise ValueError('Graphed steps do not support pose heads with noise_dim > 0: the noise '
                             'is drawn from the CPU generator and would be baked into the graph')


def _flat_views(example, device):
    """One contiguous device buffer holding a tensor per entry of `example` (256-byte aligned
    views): a whole batch then moves between staging and static inputs with ONE copy."""
    offs, total = {}, 0
    for k, v in example.items():
        offs[k] = total
        total += (v.numel() * v.element_size() + 255) // 256 * 256
    flat = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
    views = {k: flat[offs[k]:offs[k] + v.numel() * v.element_size()].view(v.dtype).view(v.shape)
             for k, v in example.items()}
    return flat, views


class _HostRead:
    """A device->host copy in flight (pinned destination + event); `value()` waits for it.  The
    destination is one of a small ring of pinned scalars: a handle that is read only after the
    ring has wrapped around raises instead of returning a later step's value."""

    def __init__(self, slot, event):
        self._slot, self._gen, self._event = slot, slot['gen'], event

    def value(self):
        self._event.synchronize()
        if self._slot['gen'] != self._gen:
            raise RuntimeError('read_async: this result was overwritten (more reads in flight than ring slots)')
        return float(self._slot['host'])


class GraphedStep:

    def __init__(self, model, example_batch, mode='train', autocast_dtype=torch.bfloat16,
                 warmup=3):
        _check_capturable(model)
        self.model = model
        self.mode = mode
        self.autocast_dtype = autocast_dtype
        dev = next(model.parameters()).device
        self.device = dev
        self._static_flat, self.static_in = _flat_views(example_batch, dev)
        for k, v in example_batch.items():
            self.static_in[k].copy_(v)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._step()
            # the poses the losses were computed from ([B, P, 3], [B, P, 4]); static graph
            # buffers like the losses: valid until the next replay
            self.static_pred = getattr(model, '_last_pred', None)
        torch.cuda.synchronize(dev)

    def _step(self):
        with torch.no_grad():
            with torch.autocast('cuda', dtype=self.autocast_dtype or torch.bfloat16,
                                enabled=self.autocast_dtype is not None):
                out = self.model.forward_pass(dict(self.static_in), mode=self.mode,
                                              optimizer_idx=-1)
        return {k: v for k, v in out.items() if isinstance(v, torch.Tensor)}

    def __call__(self, batch=None):
        """Run one step; `batch` tensors (host or device) are copied into the
        static inputs first.  Returns the dict of loss tensors (static buffers:
        read them before the next call)."""
        if batch is not None:
            for k, dst in self.static_in.items():
                dst.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.static_out

    # ---- input prefetch: hide the host->device copy of the NEXT batch behind the
    # current step (what a data loader with a prefetch queue does) -------------
    def prefetch(self, batch):
        """Start copying `batch` (pinned host tensors) into a staging buffer on a
        side stream.  Call `run_prefetched()` afterwards to step on it."""
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            flats = [_flat_views(self.static_in, self.device) for _ in range(2)]
            self._staging_flat = [f for f, _ in flats]
            self._staging = [v for _, v in flats]
            self._staged = [None, None]
            self._slot = 0
        slot = self._slot
        self._slot ^= 1
        ev_free = self._staged[slot]
        with torch.cuda.stream(self._copy_stream):
            if ev_free is not None:  # the step that consumed this slot must be done with it
                self._copy_stream.wait_event(ev_free)
            for k, dst in self._staging[slot].items():
                dst.copy_(batch[k], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        self._pending = (slot, ready)

    def run_prefetched(self):
        slot, ready = self._pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        self._static_flat.copy_(self._staging_flat[slot], non_blocking=True)  # ONE device-to-device copy
        consumed = torch.cuda.Event()
        consumed.record(cur)
        self._staged[slot] = consumed
        self.graph.replay()
        return self.static_out

    def read_async(self, key='loss'):
        """Start the device->host copy of one scalar output of the step just replayed and
        return a handle; `handle.value()` waits for that copy only.  Lets the host enqueue the
        next step before it looks at this one's loss (a training loop that logs one step late)."""
        if not hasattr(self, '_host_ring'):
            self._host_ring = [{'host': torch.empty((), dtype=torch.float32, pin_memory=True), 'gen': 0}
                               for _ in range(8)]
            self._host_slot = 0
        slot = self._host_ring[self._host_slot]
        self._host_slot = (self._host_slot + 1) % len(self._host_ring)
        slot['gen'] += 1
        slot['host'].copy_(self.static_out[key].detach().reshape(()).float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return _HostRead(slot, ev)


def allreduce_gradients(params, group=None):
    """Average the gradients of `params` over the ranks of `group`: the one data-path
    collective of data-parallel training (the reference gets it implicitly from DDP,
    scripts/train.py:85,141).  One flat fp32 buffer, ONE all-reduce (3.3 M elements =
    13 MB for pn_transformer: latency-bound on NVLink, so no bucketing), scattered back
    with a multi-tensor copy.  Safe to record into a CUDA graph (NCCL)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    # every rank must reduce the same buffer: a parameter without a gradient on this rank
    # (unused in this step) contributes zeros instead of shortening the buffer
    params = [p for p in params if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    torch._foreach_copy_(grads, views)


class GraphedTrainStep:
    """Whole training step -- forward, loss, backward, optimizer -- as one CUDA graph.

    Eager PyTorch needs ~15-20 ms of host time to issue the ~1500 launches of one
    pn_transformer training step, several times what the GPU needs to run them.  The
    step is captured once for a fixed batch shape (the optimizer must keep its state on
    the device: Adam/AdamW are switched to `capturable`) and replayed per batch; inputs
    go through static buffers like `GraphedStep`.  With torch.distributed initialised the
    gradient all-reduce of data-parallel training is part of the graph.  Reference loop being replaced: the
    PL training loop around `BaseModel.training_step` (models/modules/base_model.py:60-63)
    with automatic optimisation and `--fp16` autocast (scripts/train.py:88).
    """

    def __init__(self, model, optimizer, example_batch, autocast_dtype=torch.bfloat16, warmup=3,
                 group=None):
        _check_capturable(model)
        self.model = model
        self.optimizer = optimizer
        self.autocast_dtype = autocast_dtype
        self.group = group
        self._params = [p for p in model.parameters() if p.requires_grad]
        dev = next(model.parameters()).device
        self.device = dev
        for group in optimizer.param_groups:
            if 'capturable' in group:
                group['capturable'] = True
        self.static_in = {k: v.to(dev).clone() for k, v in example_batch.items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._fwd_bwd()
            optimizer.step()
        torch.cuda.synchronize(dev)

    def _fwd_bwd(self):
        with torch.autocast('cuda', dtype=self.autocast_dtype or torch.bfloat16,
                            enabled=self.autocast_dtype is not None):
            loss = self.model.training_step(dict(self.static_in), 0)
        loss.backward()
        allreduce_gradients(self._params, self.group)  # no-op on a single rank
        return loss.detach()

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self._fwd_bwd()
        self.optimizer.step()
        return loss

    def __call__(self, batch=None):
        """One optimisation step on `batch` (host or device tensors; None = the static
        buffers as they are).  Returns the (static) loss tensor."""
        if batch is not None:
            for k, dst in self.static_in.items():
                dst.copy_(batch[k], non_blocking=True)
        self.graph.replay()  # gradients live in the graph's private pool and are rewritten in place
        return self.static_loss
