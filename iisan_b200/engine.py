"""Train-step runner for the hot path: zero_grad -> ModelMM.forward -> backward -> (gradient all-reduce) -> Adam, with the
whole step captured once in a CUDA graph and replayed (the step is ~100 kernel launches of a few microseconds each, so
launch latency and Python overhead would otherwise dominate; every kernel of the library is stream-ordered, allocation-free
and sync-free, which is what makes the capture legal).

This replaces the body of the reference's batch loop (Code_Cached/run.py:368-385):

    step = TrainStep(model, optimizer)                      # model = iisan_b200.model.ModelMM with the SAN installed
    for ids, image, text, log_mask in train_dl:             # host (pinned) or device tensors, reference shapes
        loss = step(ids, image, text, log_mask)             # 0-dim device tensor; loss.item() only when it is logged

Inputs are copied into static device buffers (asynchronous H2D from pinned memory), so the DataLoader side is unchanged.
Data parallelism: pass ``group``; gradients are all-reduced (mean) over the ranks inside the step (NCCL, one flat bucket)
instead of through a DDP wrapper, and ``model.negatives = "global"`` adds the item-embedding all-gather.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class TrainStep:
    def __init__(self, model, optimizer, use_graph=True, group=None, warmup=3, autocast_dtype=None):
        self.model = model
        self.opt = optimizer
        self.use_graph = bool(use_graph)
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None or (dist.is_available() and dist.is_initialized())) else 1
        if self.world > 1 and group is None:
            self.group = dist.group.WORLD
        self.warmup = int(warmup)
        self.graph = None
        self.static_in = None
        self.static_loss = None
        self._params = [p for p in model.parameters() if p.requires_grad]
        self._flat = None

    # ------------------------------------------------------------------------------------------------------------
    def _allreduce(self):
        if self.world == 1:
            return
        grads = [p.grad for p in self._params if p.grad is not None]
        if self._flat is None:
            self._flat = torch.empty(sum(g.numel() for g in grads), dtype=torch.float32, device=grads[0].device)
        flat = self._flat
        torch._foreach_copy_(list(torch.split(flat, [g.numel() for g in grads])), [g.reshape(-1) for g in grads])
        dist.all_reduce(flat, group=self.group)
        flat.div_(self.world)
        torch._foreach_copy_([g.reshape(-1) for g in grads], list(torch.split(flat, [g.numel() for g in grads])))

    def _snapshot(self):
        params = [p.detach().clone() for p in self._params]
        state = []
        for p in self._params:
            st = self.opt.state.get(p, None)
            state.append(None if not st else {k: v.detach().clone() for k, v in st.items() if torch.is_tensor(v)})
        return params, state

    @torch.no_grad()
    def _restore(self, saved):
        params, state = saved
        for p, q in zip(self._params, params):
            p.copy_(q)
        for p, old in zip(self._params, state):
            st = self.opt.state.get(p, None)
            if not st:
                continue
            for k, v in st.items():
                if torch.is_tensor(v):
                    if old is not None and k in old:
                        v.copy_(old[k])
                    else:
                        v.zero_()            # the state did not exist before the warm-up: back to its initial value

    def _eager(self, ids, image, text, log_mask):
        self.opt.zero_grad(set_to_none=True)
        loss = self.model(ids, image, text, log_mask, ids.device)
        loss.backward()
        self._allreduce()
        self.opt.step()
        return loss.detach()

    def _stage(self, ids, image, text, log_mask, device):
        src = (ids.reshape(-1), image, text, log_mask)
        if self.static_in is None:
            self.static_in = tuple(torch.empty(t.shape, dtype=t.dtype, device=device) for t in src)
        for dst, s in zip(self.static_in, src):
            if dst is s or (dst.data_ptr() == s.data_ptr() and dst.shape == s.shape):
                continue
            if dst.shape != s.shape or dst.dtype != s.dtype:
                raise ValueError("TrainStep was captured for batch tensors of shape "
                                 f"{tuple(dst.shape)}/{dst.dtype}; got {tuple(s.shape)}/{s.dtype} (drop_last the loader or use_graph=False)")
            dst.copy_(s, non_blocking=True)
        return self.static_in

    def capture(self, ids, image, text, log_mask):
        """Adopt the given DEVICE tensors as the static inputs (no staging copy) and capture the step on them; afterwards
        ``replay()`` re-runs the step on whatever those tensors then hold.  Applies one step."""
        self.static_in = (ids.reshape(-1), image, text, log_mask)
        return self.__call__(*self.static_in)

    def replay(self):
        self.graph.replay()
        return self.static_loss

    def __call__(self, ids, image, text, log_mask):
        device = next(self.model.parameters()).device
        if not self.use_graph:
            args = tuple(t.to(device, non_blocking=True) for t in (ids.reshape(-1), image, text, log_mask))
            return self._eager(*args)
        inputs = self._stage(ids, image, text, log_mask, device)
        if self.graph is None:
            if not all(g.get("capturable", False) for g in self.opt.param_groups):
                raise ValueError("TrainStep(use_graph=True) needs a capturable optimizer, e.g. "
                                 "torch.optim.Adam(groups, fused=True, capturable=True)")
            # warm-up on a side stream (lazy initialisation of kernels / NCCL / optimizer state), then capture.  Parameters
            # and optimizer state are restored afterwards so that this call applies exactly ONE step like every later call.
            saved = self._snapshot()
            s = torch.cuda.Stream(device=device)
            s.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(s):
                for _ in range(self.warmup):
                    self._eager(*inputs)
            torch.cuda.current_stream(device).wait_stream(s)
            torch.cuda.synchronize(device)
            self.opt.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread may query its events while this thread captures
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.static_loss = self._eager(*inputs)
            self._restore(saved)
        self.graph.replay()
        return self.static_loss
