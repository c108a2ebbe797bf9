"""Data-parallel gradient exchange for the UNet train step (reference: accelerate's DDP wrapper around the module,
training/trainer.py:1022-1044 -- one gradient all-reduce(mean) per optimizer step).

The backward schedule (modules/unets/unet_train.py) writes parameter gradients into one flat fp32 buffer laid out in
backward-completion order; `GradAllReducer` all-reduces each bucket over NCCL on its own stream as soon as the
bucket's weight-norm backward was enqueued, so the exchange of the decoder's gradients overlaps the encoder's
backward.  Parameters' `.grad` are views of that flat buffer (no per-parameter copies, no autograd accumulation).

    net.grad_sync = GradAllReducer()          # instead of wrapping the module in torch DDP
    ...
    with net.grad_sync.no_sync(): loss.backward()   # micro-steps of gradient accumulation
    loss.backward()                                  # last micro-step: overlapped all-reduce(mean)
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, process_group: Optional["dist.ProcessGroup"] = None) -> None:
        self.pg = process_group
        self.sync_now = True
        self.stream: Optional[torch.cuda.Stream] = None
        self.bytes_reduced = 0
        # DD_DDP_TRACE=1 (tuning): CUDA events around every bucket's all-reduce and at the end of the backward schedule;
        # `trace_report()` turns the last step's events into (bucket MB, start, end) relative to the backward's end
        self.trace = os.environ.get("DD_DDP_TRACE", "0") == "1"
        self._events: list = []

    def trace_report(self) -> list:
        if not self._events:
            return []
        torch.cuda.synchronize()
        t_end = self._events[-1][2]
        rows = [(mb, -e0.elapsed_time(t_end), -e1.elapsed_time(t_end)) for mb, e0, e1 in self._events[:-1]]
        return [(round(mb, 1), round(a, 3), round(b, 3)) for mb, a, b in rows]

    @contextlib.contextmanager
    def no_sync(self):
        old, self.sync_now = self.sync_now, False
        try:
            yield
        finally:
            self.sync_now = old

    def world_size(self) -> int:
        return dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1

    def _all_reduce_mean(self, t: torch.Tensor) -> None:
        if dist.get_backend(self.pg) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg)
        else:                                   # gloo (CPU tests of the host logic) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            t.mul_(1.0 / self.world_size())

    @staticmethod
    def extra_params(net, plan, ts) -> list:
        """Trainable parameters of `net` that are neither a slot of the flat gradient buffer nor one of its scalar
        gains.  Together the three sets must cover every parameter that requires grad (checked here), otherwise a
        replica would keep a rank-local gradient and the weights diverge silently: everything not covered by the flat
        buffer is exchanged by reduce_extras at the end of the backward pass."""
        if net is None:
            return []
        covered = {id(s.param) for s in ts.slots.values()} | {id(p) for p in plan.gain_params}
        return [p for p in net.parameters() if p.requires_grad and id(p) not in covered]

    def reduce_extras(self, extras) -> None:
        """all-reduce(mean) of the .grad of `extras` as one small flat message on the current stream."""
        grads = [p.grad for p in extras if p.grad is not None]
        if not grads or self.world_size() == 1:
            return
        flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
        self._all_reduce_mean(flat)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        self.bytes_reduced += flat.numel() * 4

    def run_backward(self, net, plan, ts, saved, dD, backward_fn=None) -> torch.Tensor:
        """Runs the backward schedule with bucket-wise gradient exchange.  `backward_fn` defaults to
        unet_train.train_backward (tests substitute a host-only stand-in)."""
        if backward_fn is None:
            from .modules.unets.unet_train import train_backward as backward_fn
        params = [s for s in ts.slots.values()]
        installed = all(s.param.grad is not None and s.param.grad.data_ptr() == s.grad.data_ptr() for s in params)
        on_gpu = ts.grad_flat.is_cuda
        exchange = self.sync_now and self.world_size() > 1
        if on_gpu:
            main = torch.cuda.current_stream(ts.grad_flat.device)
            if self.stream is None:
                # high priority: the all-reduce kernels of a finished bucket should not queue behind the backward's grids
                self.stream = torch.cuda.Stream(device=ts.grad_flat.device, priority=-1)

        def bucket_done(i: int) -> None:
            if not exchange:
                return
            lo, hi = ts.bucket_ranges[i]
            if on_gpu:
                self.stream.wait_stream(main)
                with torch.cuda.stream(self.stream):
                    if self.trace:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                    self._all_reduce_mean(ts.grad_flat[lo:hi])
                    if self.trace:
                        e1.record()
                        self._events.append(((hi - lo) * 4 / 2 ** 20, e0, e1))
            else:
                self._all_reduce_mean(ts.grad_flat[lo:hi])
            self.bytes_reduced += (hi - lo) * 4

        if self.trace:
            self._events = []
        dlabel = backward_fn(net, plan, saved, dD, accumulate=installed, bucket_done=bucket_done)
        if self.trace and exchange and on_gpu:
            e = torch.cuda.Event(enable_timing=True)
            e.record()                                   # end of the backward schedule on the main stream
            self._events.append((0.0, None, e))
        if exchange and on_gpu:
            main.wait_stream(self.stream)
        if exchange:
            # Parameters whose gradients do not come out of this node (emb_label / emb_label_unconditional /
            # logvar_linear: LabelEmbeddingFunction and SigmaLogvarFunction hand theirs to autograd) are exchanged once
            # the whole backward pass has accumulated them into .grad -- the same point at which torch DDP finalises.
            extras = self.extra_params(net, plan, ts)
            if extras:
                try:
                    torch.autograd.Variable._execution_engine.queue_callback(lambda: self.reduce_extras(extras))
                except RuntimeError:            # not inside an autograd backward pass (host-logic tests): caller's job
                    pass
        if not installed:
            for s in params:
                if s.param.grad is None:
                    s.param.grad = s.grad
                elif s.param.grad.data_ptr() != s.grad.data_ptr():
                    s.param.grad.add_(s.grad)
            for i, p in enumerate(plan.gain_params):
                gview = ts.dgains[i]
                if p.grad is None:
                    p.grad = gview
                elif p.grad.data_ptr() != gview.data_ptr():
                    p.grad.add_(gview)
        return dlabel
