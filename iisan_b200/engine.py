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
    def __init__(self, model, optimizer, use_graph=True, group=None, warmup=3, autocast_dtype=None, store=None, packed=False):
        self.model = model
        self.store = store            # iisan_b200.store.CachedStateStore: batches are (ids, log_mask) only, states gathered on the device
        self.packed = bool(packed)    # the image / text tensors hold only the selected layers [N, A, d] (already gathered from a store)
        self.opt = optimizer
        self.use_graph = bool(use_graph)
        self.group = group
        if group is False:            # explicit "no communication" (measurements: the step without its collectives)
            self.group, self.world = None, 1
        else:
            self.world = dist.get_world_size(group) if (group is not None or (dist.is_available() and dist.is_initialized())) else 1
            if self.world > 1 and group is None:
                self.group = dist.group.WORLD
        if self.world > 1:
            # Generation 3 of the fused side-adapter path (san_chain3.cu / san_lr.cu) faulted intermittently on 8 x B200
            # ("unspecified launch failure" on one rank, 3 of 3 runs; 1 and 2 GPUs and every compute-sanitizer run are clean) and the
            # cause is not found yet: data-parallel steps run generation 2 (measured cost at 8 GPUs: 1.30 -> 1.37 ms per step).
            # IISAN_B200_DP_CHAIN_GEN=3 overrides (for the investigation).
            import os
            from . import _lib as L
            L.load().iisan_debug_chain_generation(int(os.environ.get("IISAN_B200_DP_CHAIN_GEN", "2")))
        self.warmup = int(warmup)
        self.graph = None
        self.static_in = None
        self.static_loss = None
        self._params = [p for p in model.parameters() if p.requires_grad]
        self._grad_arena = None

    # ------------------------------------------------------------------------------------------------------------
    def _arena(self, device):
        """Gradient arena of the step (ops.GradArena): every parameter gradient is a slice of one buffer -- ONE memset per step
        instead of a fill kernel per backward function, and (N > 1) one collective over the used part."""
        if self._grad_arena is None:
            from .ops import GradArena
            n = sum((p.numel() + 63) // 64 * 64 for p in self._params) + 64 * 64
            self._grad_arena = GradArena(n, device)
        return self._grad_arena

    def _allreduce(self):
        """Mean of the gradients over the ranks, in place and without staging copies.  With the gradient arena active (the
        normal case) this is ONE collective over the used part of the arena; gradients that live elsewhere (a foreign autograd
        function) are reduced through their own storage."""
        if self.world == 1:
            return
        bases = {}
        arena = self._grad_arena
        if arena is not None and arena.off > 0:
            bases[arena.buf.untyped_storage().data_ptr()] = arena.used()
        for p in self._params:
            g = p.grad
            if g is None:
                continue
            st = g.untyped_storage()
            key = st.data_ptr()
            if key not in bases:
                # the whole storage as one flat tensor: the arena of SAN / SASRec gradients, the tensor itself otherwise
                bases[key] = torch.empty(0, dtype=g.dtype, device=g.device).set_(st)
        nccl = dist.get_backend(self.group) == "nccl"
        for b in bases.values():
            if nccl:
                dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group)       # mean inside the collective
            else:
                dist.all_reduce(b, group=self.group)
                b.div_(self.world)

    def _snapshot(self):
        params = [p.detach().clone() for p in self._params]
        state = []
        for p in self._params:
            st = self.opt.state.get(p, None)
            state.append(None if not st else {k: v.detach().clone() for k, v in st.items() if torch.is_tensor(v)})
        step_dev = getattr(self.opt, "_step_dev", None)
        return params, state, (None if step_dev is None else step_dev.detach().clone())

    @torch.no_grad()
    def _restore(self, saved):
        params, state, step_dev = saved
        cur = getattr(self.opt, "_step_dev", None)        # iisan_b200.optim.FusedAdam keeps its step count on the device
        if cur is not None:
            cur.copy_(step_dev) if step_dev is not None else cur.zero_()
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
        from .ops import GradArena
        self.opt.zero_grad(set_to_none=True)
        arena = self._arena(ids.device)
        if arena is not None:
            arena.reset()
        prev, GradArena.active = GradArena.active, arena
        try:
            if self.store is not None:
                image, text = self.store.gather(ids)
                loss = self.model(ids, image, text, log_mask, ids.device, packed=True)
            else:
                loss = self.model(ids, image, text, log_mask, ids.device, packed=self.packed)
            loss.backward()
        finally:
            GradArena.active = prev
        self._allreduce()
        self.opt.step()
        return loss.detach()

    def _stage(self, ids, image, text, log_mask, device):
        src = (ids.reshape(-1), image, text, log_mask)
        if self.static_in is None:
            self.static_in = tuple(None if t is None else torch.empty(t.shape, dtype=t.dtype, device=device) for t in src)
        for dst, s in zip(self.static_in, src):
            if dst is None:
                continue
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
        return self.__call__(ids.reshape(-1), image, text, log_mask)

    def replay(self):
        self.graph.replay()
        return self.static_loss

    def __call__(self, ids, image, text, log_mask):
        device = next(self.model.parameters()).device
        if not self.use_graph:
            args = tuple(None if t is None else t.to(device, non_blocking=True) for t in (ids.reshape(-1), image, text, log_mask))
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


def stage_states_h2d(dst, src, sel, stream=None):
    """Copy the layers ``sel`` of a pinned host batch ``src`` [..., layers, d] into the device buffer ``dst`` of the same shape
    (iisan_stage_states_h2d: one 2-D DMA per run of adjacent layers).  The other layers of ``dst`` keep their old content --
    the kernels never read them (6 of the 13 cached layers are never used)."""
    import ctypes as C
    from . import _lib as L
    lib = L.load()
    if not src.is_pinned() or not src.is_contiguous() or not dst.is_contiguous() or src.shape != dst.shape or src.dtype != dst.dtype:
        raise ValueError("stage_states_h2d needs a contiguous pinned host tensor and a device tensor of the same shape/dtype")
    layers, d = src.shape[-2], src.shape[-1]
    n_rows = src.numel() // (layers * d)
    sel = sorted(set(int(v) for v in sel))
    arr = (C.c_int32 * len(sel))(*sel)
    st = stream if stream is not None else torch.cuda.current_stream(dst.device)
    L.check(lib.iisan_stage_states_h2d(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), n_rows, layers, d,
                                       L.torch_dtype_code(src.dtype), arr, len(sel), C.c_void_p(st.cuda_stream)),
            "iisan_stage_states_h2d")


class PipelinedTrainStep:
    """Double-buffered TrainStep for host-resident batches: while the graph of batch i runs, the selected layers of batch
    i+1 stream over the host link on a copy stream.

        pipe = PipelinedTrainStep(model, optimizer)
        pipe.submit(first_batch)
        for nxt in loader:                      # (ids, image, text, log_mask): pinned host tensors of the reference shapes
            pipe.submit(nxt)                    # asynchronous H2D of the NEXT batch
            loss = pipe.run()                   # step on the batch submitted before it
        loss = pipe.run()

    ``run()`` returns the loss as a 0-dim DEVICE tensor (reading it with ``.item()`` waits for the step).  ``run_logged()`` is the
    logging-friendly variant: it also enqueues the device-to-host copy of that loss into pinned memory and returns the VALUE of
    the step before it (a float; None for the first call), so the host thread never waits for the step it has just launched --
    the reference accumulates the loss on the device and logs it every ``steps_for_log`` batches (Code_Cached/run.py:382, 390-392; its
    per-step ``torch.isnan`` test at :387 is what synchronises its loop); here every step's loss reaches the host, one step late.
    ``last_loss()`` waits for and returns the newest one.
    """

    def __init__(self, model, optimizer, group=None, use_graph=True, depth=2, store=None, prefetch_gather=True):
        self.model, self.opt, self.group, self.use_graph = model, optimizer, group, use_graph
        self.depth = depth
        self.store = store
        # With a store: the per-item gather of batch i + 1 (HBM-bound: every selected layer read and written once) is issued by
        # submit() on the copy stream, behind the H2D copy of its ids, and so runs WHILE step i computes (whose kernels are
        # latency-, not bandwidth-bound); the captured step then consumes the gathered [N, A, d] buffers (packed=True).
        # prefetch_gather=False keeps the gather inside the captured step (one stream, no overlap).
        self.prefetch = store is not None and bool(prefetch_gather)
        if self.prefetch:
            self.steps = [TrainStep(model, optimizer, use_graph=use_graph, group=group, packed=True) for _ in range(depth)]
        else:
            self.steps = [TrainStep(model, optimizer, use_graph=use_graph, group=group, store=store) for _ in range(depth)]
        self.bufs = [None] * depth
        self.copied = [None] * depth
        self.done = [None] * depth
        self.copy_stream = None
        self.n_sub = 0
        self.n_run = 0
        self._loss_host = None        # pinned [depth] fp32: asynchronous loss read-back of run_logged()
        self._loss_ev = [None] * depth
        self._n_logged = 0
        enc = getattr(model, "mm_encoder", None)
        plan = getattr(enc, "plan", None)
        self.sel_img = list(plan.layers_img_read) if plan is not None else None      # incl. layer 0 under remove_first
        self.sel_text = list(plan.layers_text_read) if plan is not None else None

    def submit(self, ids, image=None, text=None, log_mask=None):
        """With a store attached the batch is ``submit(ids, log_mask=log_mask)``: only ids and mask cross the host link."""
        if self.n_sub - self.n_run >= self.depth:
            raise RuntimeError("PipelinedTrainStep: run() the submitted batches before submitting more")
        device = next(self.model.parameters()).device
        k = self.n_sub % self.depth
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=device)
        if self.bufs[k] is None:
            self.bufs[k] = tuple(None if t is None else torch.zeros(t.shape, dtype=t.dtype, device=device)
                                 for t in (ids.reshape(-1), image, text, log_mask))
            if self.prefetch:
                gi, gt = self.store.batch_buffers(ids.numel())
                self.bufs[k] = (self.bufs[k][0], gi, gt, self.bufs[k][3])
            self.copied[k] = torch.cuda.Event()
            self.done[k] = torch.cuda.Event()
            self.done[k].record(torch.cuda.current_stream(device))
        cs = self.copy_stream
        cs.wait_event(self.done[k])                           # the previous step on this buffer set has consumed it
        d_ids, d_img, d_txt, d_lm = self.bufs[k]
        with torch.cuda.stream(cs):
            d_ids.copy_(ids.reshape(-1), non_blocking=True)
            d_lm.copy_(log_mask, non_blocking=True)
            if self.prefetch:
                self.store.gather(d_ids, out=(d_img, d_txt))
                image = text = None
            for dst, src, sel in ((d_img, image, self.sel_img), (d_txt, text, self.sel_text)):
                if dst is None or src is None:
                    continue
                if (not src.is_cuda) and src.is_pinned() and sel is not None and src.dim() >= 3 and len(sel) < src.shape[-2]:
                    stage_states_h2d(dst, src, sel, cs)
                else:
                    dst.copy_(src, non_blocking=True)
            self.copied[k].record(cs)
        self.n_sub += 1

    def run(self):
        if self.n_run >= self.n_sub:
            raise RuntimeError("PipelinedTrainStep.run(): nothing submitted")
        k = self.n_run % self.depth
        device = self.bufs[k][0].device
        cur = torch.cuda.current_stream(device)
        cur.wait_event(self.copied[k])
        st = self.steps[k]
        if st.graph is None and self.use_graph:
            loss = st.capture(*self.bufs[k])
        elif self.use_graph:
            loss = st.replay()
        else:
            loss = st(*self.bufs[k])
        self.done[k].record(cur)
        self.n_run += 1
        return loss

    def run_logged(self):
        """run() + asynchronous loss read-back; returns the loss value (float) of the PREVIOUS run_logged() step, None at first."""
        k = self.n_run % self.depth
        loss = self.run()
        if self._loss_host is None:
            self._loss_host = torch.zeros(self.depth, dtype=torch.float32).pin_memory()
        if self._loss_ev[k] is None:
            self._loss_ev[k] = torch.cuda.Event()
        prev = self.last_loss() if self._n_logged > 0 else None        # waits for step i - 1 only: step i is already queued
        self._loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self._loss_ev[k].record(torch.cuda.current_stream(loss.device))
        self._n_logged += 1
        self._last_k = k
        return prev

    def last_loss(self):
        """Value of the newest loss whose read-back was enqueued by run_logged() (waits for that copy)."""
        if self._n_logged == 0:
            raise RuntimeError("PipelinedTrainStep.last_loss(): no run_logged() step yet")
        self._loss_ev[self._last_k].synchronize()
        return float(self._loss_host[self._last_k])
