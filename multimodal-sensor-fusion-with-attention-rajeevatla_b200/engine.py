"""Fused HybridFusion train step / inference pass with flat arenas, CUDA graphs
and batch-sharded data parallelism.

One process per GPU.  Everything a training step of the reference does around
the fusion model (``train.MultimodalFusionModule.training_step`` +
``configure_gradient_clipping`` + ``AdamW``; src/train.py:302-324,378-382,
416-430) is enqueued as one stream of msf_b200 kernels:

    forward -> CE(label smoothing) -> backward -> [NCCL all-reduce of the
    gradient arena] -> global grad norm -> clip + AdamW -> (bf16 re-pack)

and captured once into a CUDA graph; ``step()`` copies the batch into static
buffers and replays it.  Dropout masks and the Adam step count live in a small
device-side state so every replay draws fresh masks.

Windows are independent, so data parallelism shards the batch across ranks with
replicated parameters; the only collective is the gradient all-reduce
(SURVEY.md §8e).  Evaluation / ECE are shard-local followed by one integer
all-reduce of the bin counts.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import _native as N
from . import ops


class FusionEngine:
    def __init__(self, model, batch: int, *, precision: str = "bf16", lr: float = 1e-3,
                 weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 label_smoothing: float = 0.05, max_grad_norm: float = 1.0, seed: int = 1234,
                 process_group=None, use_graph: bool = True, device: Optional[torch.device] = None,
                 comm: str = "auto", feature_dtype: torch.dtype = torch.float32):
        """``feature_dtype=torch.bfloat16`` (tensor-core path): the input slots hold the features as bf16 rows
        (msf_fusion_call.x_bf16) — the projection kernel rounds them to bf16 anyway, and a host batch that is
        already bf16 (``pinned_batch()``) crosses PCIe with half the bytes."""
        self.dev = device or ops.require_cuda("FusionEngine")
        if feature_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("feature_dtype must be torch.float32 or torch.bfloat16")
        if feature_dtype == torch.bfloat16 and precision != "bf16":
            raise N.MsfError("bf16 features need the tensor-core path (precision='bf16')")
        self.feature_dtype = feature_dtype
        self.arena_bf16 = None
        self.model = model.to(self.dev)
        self.plan = model._plan()
        self.batch = int(batch)
        self.prec = ops.PRECISIONS[precision]
        self.precision = precision
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.smoothing, self.max_norm = label_smoothing, max_grad_norm
        self.p = float(model.dropout.p)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.use_graph = use_graph
        plan, dev, B = self.plan, self.dev, self.batch
        f32 = dict(dtype=torch.float32, device=dev)

        # flat master arena; the module's parameters become views into it
        own = dict(self.model.named_parameters())
        self.arena = torch.zeros(plan.total, **f32)
        with torch.no_grad():
            for key, off, shape in plan.slots:
                prm = own[key]
                n = prm.numel()
                self.arena[off:off + n].copy_(prm.detach().reshape(-1))
                prm.data = self.arena[off:off + n].view(shape)
        # gradient exchange under data parallelism: "p2p" = fused reduce + AdamW kernels over NVLink peer
        # memory (msf_dp_optimizer_step), "nccl" = all-reduce then optimizer, "auto" = p2p when available
        self.comm, self.dp_comm = "none", None
        self.grad = None
        # "zshard": sharded optimizer over NVLink peer memory (msf_dpz_optimizer_step_packed, tensor-core path only):
        # every rank reduces and updates only the optimizer units it owns and pushes their bf16 copies to all ranks;
        # "auto" prefers it, then "p2p" (replicated AdamW from pushed reduced gradients), then NCCL.
        want = [comm] if comm != "auto" else (["zshard", "p2p"] if self.prec == N.MSF_PREC_BF16 else ["p2p"])
        if self.world > 1:
            for mode in want:
                if mode not in ("zshard", "p2p"):
                    continue
                try:
                    if mode == "zshard":
                        if self.prec != N.MSF_PREC_BF16:
                            raise N.MsfError("the sharded step needs the tensor-core path")
                        self._setup_sharded()
                    else:
                        self._setup_peer_memory()
                    self.comm = mode
                    break
                except Exception as exc:  # noqa: BLE001 - symmetric memory is optional plumbing
                    if comm == mode:
                        raise
                    import sys
                    print(f"msf_b200: {mode} gradient exchange unavailable ({exc})", file=sys.stderr)
        if self.world > 1 and self.comm == "none":
            self.comm = "nccl"
        if self.grad is None:
            self.grad = torch.zeros(plan.total, **f32)
        self.exp_avg = torch.zeros(plan.total, **f32)
        self.exp_avg_sq = torch.zeros(plan.total, **f32)
        if self.prec == N.MSF_PREC_BF16 and getattr(self, "arena_bf16", None) is None:
            self.arena_bf16 = plan.pack_bf16(self.arena)
        elif self.prec != N.MSF_PREC_BF16:
            self.arena_bf16 = None
        self.ws = torch.empty(plan.workspace_bytes(B, self.prec), dtype=torch.uint8, device=dev)

        # static I/O buffers (graph inputs / outputs).  Two input slots: the host-facing stream API
        # (train_stream) copies batch i+1 into one slot on a copy stream while the graph captured over the
        # other slot computes batch i; the resident API uses slot 0 only.
        self._slots = [([torch.zeros(B, d, dtype=feature_dtype, device=dev) for d in plan.dims], torch.ones(B, plan.M, **f32),
                        torch.zeros(B, dtype=torch.int64, device=dev)) for _ in range(2)]
        self.x, self.mask, self.labels = self._slots[0]
        self.logits = torch.zeros(B, plan.C, **f32)
        self.dlogits = torch.zeros(B, plan.C, **f32)
        self.row_loss = torch.zeros(B, **f32)
        self.loss = torch.zeros(1, **f32)
        self.sq_norm = torch.zeros(2, dtype=torch.float64, device=dev)   # [used by the optimizer, from the pass]
        self.conf = torch.zeros(B, **f32)
        self.pred = torch.zeros(B, dtype=torch.int64, device=dev)
        # {seed, offset, step, lr bits}; step is 1-based when the optimizer kernel reads it.  The learning rate lives
        # on the device (low word of state[3], fp32 bits) and the optimizer launches are given lr = -1, so a captured
        # graph follows set_lr() without being re-captured (the reference steps a cosine schedule per epoch,
        # src/train.py:395-404).  Data-parallel replicas draw different dropout masks: the rank is folded in.
        rank = torch.distributed.get_rank(process_group) if self.world > 1 else 0
        self.state = torch.tensor([seed + 7919 * rank, 0, 1, 0], dtype=torch.int64, device=dev)
        self.set_lr(lr)
        self._train_graph = None
        self._train_graphs = [None, None]
        self._loss_ptr = {}          # slot -> where that slot's graph writes the mean loss (default: self.loss)
        self._infer_graph = None
        self._subset_graphs = {}
        self._packed_slots = False   # set by pinned_batch(): train_stream's input slots as one buffer each
        self._epoch_graphs = {}      # tuple of slots -> (graph over one step per slot, per-step losses)
        self._subset_masks = {}
        self._fold_wov = None        # fold_for_inference(): Wo Wv per attention module, for infer_sweep
        self._copy_stream = None
        self.launches_per_step = 0

    def set_lr(self, lr: float) -> None:
        """Learning rate of every later optimizer step, captured graphs included (stream-ordered device write)."""
        import struct
        self.lr = float(lr)
        bits = struct.unpack("<I", struct.pack("<f", self.lr))[0]
        self.state[3:4].copy_(torch.tensor([bits], dtype=torch.int64), non_blocking=False)

    # -- enqueue helpers (all on the current stream) ---------------------------------
    def set_input_norms(self, norms) -> None:
        """Per-modality ``nn.LayerNorm`` applied to the inputs inside the projection kernel of every later INFERENCE
        pass (src/train.py:170-171,267-268: the LayerNorm between encoder and fusion): ``load_batch`` / ``infer`` are
        then fed the raw encoder outputs.  ``norms``: mapping or sequence in ``plan.names`` order (None entries = no
        LayerNorm for that modality); ``None`` switches it off.  Captured inference graphs are dropped."""
        self._infer_graph, self._subset_graphs = None, {}
        if norms is None:
            self._ln, self._ln_eps = None, 1e-5
            return
        keyed = hasattr(norms, "keys")
        seq = [(norms[m] if m in norms else None) if keyed else norms[i] for i, m in enumerate(self.plan.names)]
        if not ops.layer_norm_fused(self.plan, self.prec):
            raise N.MsfError("input LayerNorm is fused on the tensor-core path only (msf_fusion_layer_norm_fused)")
        eps = {float(n.eps) for n in seq if n is not None}
        if len(eps) > 1:
            raise ValueError("the fused input LayerNorm takes one eps for all modalities")
        put = lambda t: None if t is None else t.detach().to(device=self.dev, dtype=torch.float32).contiguous()  # noqa: E731
        self._ln = [None if n is None else (put(n.weight), put(n.bias)) for n in seq]
        self._ln_eps = eps.pop() if eps else 1e-5

    def _call(self, training: bool, slot: int = 0) -> N.FusionCall:
        x, mask, _ = self._slots[slot]
        ln = getattr(self, "_ln", None)
        if training and ln is not None:
            raise N.MsfError("FusionEngine trains the fusion model only: input LayerNorm parameters belong to the "
                             "caller's optimizer (use HybridFusion.forward(input_norms=...) / pipeline.EncodeFuse)")
        c = ops._make_call(self.plan, self.batch, self.prec, training and self.p > 0, self.p, 0, 0,
                           self.arena, self.arena_bf16, x, mask, self.ws, ln=ln, ln_eps=getattr(self, "_ln_eps", 1e-5))
        c.rng_state = self.state.data_ptr()
        return c

    def _enqueue_forward(self, training: bool, slot: int = 0) -> None:
        c = self._call(training, slot)
        c.logits = self.logits.data_ptr()
        N.check(N.lib().msf_fusion_forward(ctypes_ref(self.plan.shape), ctypes_ref(c), ops._stream()))

    def _enqueue_train_step(self, slot: int = 0) -> None:
        norm_given = self._enqueue_pass(slot)
        self._enqueue_optimizer(norm_given)

    def _enqueue_pass(self, slot: int = 0, grad: Optional[torch.Tensor] = None, micro_batches: int = 1) -> bool:
        """Forward + CE(label smoothing) + backward of one batch as one enqueue (msf_fusion_train_pass); the gradient
        arena (``grad`` or self.grad) is overwritten.  Returns whether the pass also produced the gradient square norm."""
        lib = N.lib()
        st = ops._stream()
        # Mean over the GLOBAL batch (times the micro-batches of an accumulated step): each rank scales by
        # 1/(B*world*micro_batches), the gradient exchange / accumulation sums.
        c = self._call(True, slot)
        c.logits, c.grad_params = self.logits.data_ptr(), (self.grad if grad is None else grad).data_ptr()
        # single GPU, fused pass: the kernels that write the gradients also sum their squares, so the optimizer
        # launch needs no pass over the gradient arena before it can clip (MSF_OPT_NORM_GIVEN)
        norm_given = (self.world == 1 and self.arena_bf16 is not None and grad is None
                      and lib.msf_fusion_train_pass_is_fused(ctypes_ref(self.plan.shape), self.prec) == 1)
        if norm_given:
            c.grad_sq = self.sq_norm.data_ptr() + 8
        N.check(lib.msf_fusion_train_pass(ctypes_ref(self.plan.shape), ctypes_ref(c),
                                          self._slots[slot][2].data_ptr(), self.smoothing,
                                          1.0 / (self.batch * self.world * micro_batches), self.row_loss.data_ptr(),
                                          self._loss_ptr.get(slot, self.loss.data_ptr()), self.dlogits.data_ptr(),
                                          N.MSF_TRAIN_DEAD_SLOTS_ZERO if grad is None else 0,   # self.grad is zero-
                                          st))                  # initialised and only ever written by this entry point
        return norm_given

    def _enqueue_optimizer(self, norm_given: bool) -> None:
        lib = N.lib()
        st = ops._stream()
        if self.comm == "zshard":
            N.check(lib.msf_dpz_optimizer_step_packed(
                ctypes_ref(self.plan.shape), ctypes_ref(self.dp_comm), self.arena.data_ptr(), self.grad.data_ptr(),
                self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.state.data_ptr(), -1.0, self.betas[0],
                self.betas[1], self.eps, self.wd, 1.0, self.max_norm, 1, st))
            return
        if self.comm == "p2p":
            # reduce-scatter + norm over NVLink peer memory, then clip + AdamW (+ bf16 re-pack + state advance
            # in the same launch on the tensor-core path) from the local reduced arena
            args = (ctypes_ref(self.plan.shape), ctypes_ref(self.dp_comm), self.arena.data_ptr(),
                    self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.state.data_ptr(), -1.0,
                    self.betas[0], self.betas[1], self.eps, self.wd, 1.0, self.max_norm)
            if self.arena_bf16 is not None:
                N.check(lib.msf_dp_optimizer_step_packed(*args, self.arena_bf16.data_ptr(), 1, st))
                return
            N.check(lib.msf_dp_optimizer_step(*args, st))
        else:
            if self._nccl_step(lib, st, norm_given):
                return   # optimizer, bf16 re-pack and state advance were one launch
        if self.arena_bf16 is not None:
            N.check(lib.msf_fusion_pack_bf16(ctypes_ref(self.plan.shape), self.arena.data_ptr(),
                                             self.arena_bf16.data_ptr(), st))
        N.check(lib.msf_train_state_advance(self.state.data_ptr(), st))

    def _nccl_step(self, lib, st, norm_given: bool = False) -> bool:
        if self.world > 1:
            self._all_reduce_gradients()
        if self.arena_bf16 is not None:
            # clip + AdamW + bf16 re-pack + train-state advance in one kernel
            N.check(lib.msf_fusion_optimizer_step_packed(
                ctypes_ref(self.plan.shape), self.arena.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                self.exp_avg_sq.data_ptr(), self.state.data_ptr(), -1.0, self.betas[0], self.betas[1], self.eps,
                self.wd, 1.0, self.max_norm, self.sq_norm.data_ptr(), self.arena_bf16.data_ptr(),
                1 | (N.MSF_OPT_NORM_GIVEN if norm_given else 0), st))
            return True
        # global-norm clip + AdamW, skipping the dead q/k slots' moments (their gradients are exact zeros)
        N.check(lib.msf_fusion_optimizer_step(ctypes_ref(self.plan.shape), self.arena.data_ptr(),
                                              self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), self.state.data_ptr(), -1.0,
                                              self.betas[0], self.betas[1], self.eps, self.wd, 1.0, self.max_norm,
                                              self.sq_norm.data_ptr(), st))
        return False

    def _setup_peer_memory(self) -> None:
        """Gradient / reduced-gradient arenas and the signal block in symmetric memory, peer pointers
        gathered through torch's rendezvous (plumbing only: the kernels are ours)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.pg if self.pg is not None else dist.group.WORLD
        n = self.plan.total
        self.grad = symm_mem.empty(n, dtype=torch.float32, device=self.dev)
        self.stage = symm_mem.empty(n, dtype=torch.float32, device=self.dev)
        self.red = symm_mem.empty(n, dtype=torch.float32, device=self.dev)
        self.sig = symm_mem.empty(64, dtype=torch.int64, device=self.dev)
        handles = [symm_mem.rendezvous(t, group) for t in (self.grad, self.red, self.sig, self.stage)]
        self.grad.zero_()
        self.stage.zero_()
        self.red.zero_()
        self.sig.zero_()
        torch.cuda.synchronize(self.dev)
        dist.barrier(group)  # nobody signals into a block that is not zeroed yet
        comm = N.DpComm()
        comm.rank, comm.world = dist.get_rank(group), self.world
        for r in range(self.world):
            comm.grads[r] = int(handles[0].buffer_ptrs[r])
            comm.reds[r] = int(handles[1].buffer_ptrs[r])
            comm.sigs[r] = int(handles[2].buffer_ptrs[r])
            comm.stages[r] = int(handles[3].buffer_ptrs[r])
        self._symm_handles = handles
        self.dp_comm = comm

    def _setup_sharded(self) -> None:
        """Master arena, compute arena, staging (world x arena) and signal block in symmetric memory; the module's
        parameters are re-pointed into the symmetric master arena."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.pg if self.pg is not None else dist.group.WORLD
        n, plan = self.plan.total, self.plan
        arena = symm_mem.empty(n, dtype=torch.float32, device=self.dev)
        arena.copy_(self.arena)
        packed = plan.pack_bf16(arena)
        arena16 = symm_mem.empty(packed.numel(), dtype=packed.dtype, device=self.dev)
        arena16.copy_(packed)
        self.zstage = symm_mem.empty((n + 3) // 4 * 4 * self.world, dtype=torch.float32, device=self.dev)   # slices 16-byte aligned
        self.sig = symm_mem.empty(64, dtype=torch.int64, device=self.dev)
        grad = symm_mem.empty(n, dtype=torch.float32, device=self.dev)
        handles = [symm_mem.rendezvous(t, group) for t in (arena, arena16, self.zstage, self.sig, grad)]
        self.zstage.zero_()
        self.sig.zero_()
        grad.zero_()
        own = dict(self.model.named_parameters())
        with torch.no_grad():
            for key, off, shape in plan.slots:
                own[key].data = arena[off:off + own[key].numel()].view(shape)
        self.arena, self.arena_bf16, self.grad = arena, arena16, grad
        torch.cuda.synchronize(self.dev)
        dist.barrier(group)  # nobody signals into a block that is not zeroed yet
        comm = N.DpzComm()
        comm.rank, comm.world = dist.get_rank(group), self.world
        for r in range(self.world):
            comm.params[r] = int(handles[0].buffer_ptrs[r])
            comm.arenas_bf16[r] = int(handles[1].buffer_ptrs[r])
            comm.stages[r] = int(handles[2].buffer_ptrs[r])
            comm.sigs[r] = int(handles[3].buffer_ptrs[r])
        # NVLink multicast (NVLS) windows of the gradient / compute / master arenas, opt-in (MSF_DP_MULTICAST=1): the
        # owner of a unit reads its gradient summed by the switch and stores every updated weight once for all ranks.
        # Measured: 194 vs 186 us/step at 2 GPUs, 200 vs 203 at 8 — the push path stays the default (fixed-order sums)
        import os
        mc = [int(getattr(handles[i], "multicast_ptr", 0) or 0) for i in (4, 1, 0)]
        local_ok = all(int(handles[i].buffer_ptrs[comm.rank]) == t.data_ptr()
                       for i, t in ((4, grad), (1, arena16), (0, arena)))
        self.multicast = all(mc) and local_ok and os.environ.get("MSF_DP_MULTICAST", "0") == "1"
        flag = torch.tensor([int(self.multicast)], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # every rank takes the same path
        self.multicast = bool(int(flag.item()))
        if self.multicast:
            comm.mc_grad, comm.mc_arena_bf16, comm.mc_params = mc
        self._symm_handles = handles
        self.dp_comm = comm
        owner = torch.empty(n, dtype=torch.int8)
        N.check(N.lib().msf_dpz_owner_map(ctypes_ref(plan.shape), self.world, owner.data_ptr()))
        self._owner = owner.to(self.dev)
        self._rank = comm.rank

    def gather_parameters(self) -> None:
        """Sharded step only: make the fp32 master arena (the module's parameters) complete on every rank — each
        weight tile is copied from the rank that owns it (one all-reduce of the owned values; sums with zeros are
        exact).  Call before reading parameters for a checkpoint / state_dict; a no-op in every other mode.  Adam
        moments stay sharded (`exp_avg` / `exp_avg_sq` hold the owner's values for owned tiles)."""
        if self.comm != "zshard":
            return
        mine = self._owner == self._rank
        owned_anywhere = self._owner >= 0
        buf = torch.where(mine, self.arena, torch.zeros_like(self.arena))
        torch.distributed.all_reduce(buf, group=self.pg)
        self.arena.copy_(torch.where(owned_anywhere, buf, self.arena))

    def _live_gradient_views(self):
        """Views of the gradient arena that can be non-zero: everything except the query/key projection
        slots of the pair modules (exact zeros on every rank) — 53 % of the arena."""
        if getattr(self, "_live_views", None) is None:
            H, plan = self.plan.H, self.plan
            spans = []
            for key, off, shape in plan.slots:
                n = 1
                for d in shape:
                    n *= int(d)
                if ".query_proj." in key or ".key_proj." in key:
                    continue
                if spans and spans[-1][1] == off:
                    spans[-1][1] = off + n      # merge adjacent live tensors into one span
                else:
                    spans.append([off, off + n])
            self._live_views = [self.grad[a:b] for a, b in spans]
        return self._live_views

    def _all_reduce_gradients(self) -> None:
        """Sum the flat gradient arena over the ranks (each rank pre-scaled by 1/(B*world)).  One NCCL call:
        a grouped call over the live spans only was measured slower (335 vs 318 us/step at 2 GPUs)."""
        torch.distributed.all_reduce(self.grad, group=self.pg)

    def _enqueue_inference(self, present_hint: int = 0) -> None:
        c = self._call(False)
        c.logits = self.logits.data_ptr()
        N.check(N.lib().msf_fusion_infer_pass(ctypes_ref(self.plan.shape), ctypes_ref(c), self.conf.data_ptr(),
                                              self.pred.data_ptr(), present_hint, ops._stream()))

    def _capture(self, fn):
        if not self.use_graph:
            return None
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            fn()  # warm-up outside capture (lazy module loading, NCCL channel setup)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fn()
        return graph

    def _one_buffer_batch(self, host: bool):
        """(features, mask, labels) as views of ONE allocation, laid out in the order load_batch copies them.  A host
        batch from pinned_batch() and a train_stream input slot are then adjacent on both sides and load_batch
        sends the whole batch as a single transfer."""
        B, plan = self.batch, self.plan
        fsz = 2 if self.feature_dtype == torch.bfloat16 else 4
        sizes = [B * d * fsz for d in plan.dims] + [B * plan.M * 4, B * 8]
        if any(sz % 16 for sz in sizes[:-1]):
            return None   # a tensor would start off a 16-byte boundary (bulk copies, int64 labels): separate buffers
        total = sum(sizes)
        buf = (torch.zeros(total, dtype=torch.uint8).pin_memory() if host
               else torch.zeros(total, dtype=torch.uint8, device=self.dev))
        parts, off = [], 0
        for sz in sizes:
            parts.append(buf[off:off + sz])
            off += sz
        feats = [p.view(self.feature_dtype).view(B, d) for p, d in zip(parts, plan.dims)]
        mask = parts[-2].view(torch.float32).view(B, plan.M)
        labels = parts[-1].view(torch.int64)
        mask.fill_(1.0)
        return feats, mask, labels

    def pinned_batch(self):
        """A host batch in page-locked memory shaped for this engine: ``(features list, mask, labels)``, views of one
        pinned allocation in the engine's copy order.  Fill it (a DataLoader collate function can write straight
        into it) and hand it to train_stream: it crosses PCIe as one transfer instead of one per tensor.  Call it
        before the first train_stream (the stream's device slots are laid out to match when they are created)."""
        if self._copy_stream is None:
            self._packed_slots = True   # train_stream lays its two input slots out the same way (first call)
        got = self._one_buffer_batch(host=True)
        if got is None:
            B, plan = self.batch, self.plan
            got = ([torch.zeros(B, d, dtype=self.feature_dtype).pin_memory() for d in plan.dims], torch.ones(B, plan.M).pin_memory(),
                   torch.zeros(B, dtype=torch.int64).pin_memory())
        return got

    # -- public API --------------------------------------------------------------------
    def load_batch(self, features: Dict[str, torch.Tensor] | Sequence[torch.Tensor],
                   mask: Optional[torch.Tensor], labels: Optional[torch.Tensor] = None, slot: int = 0) -> int:
        """Copy one batch (host-pinned or device tensors) into the static buffers of `slot` on the current
        stream.  Returns the bytes that crossed host->device."""
        feats = [features[m] for m in self.plan.names] if isinstance(features, dict) else list(features)
        x, msk, lab = self._slots[slot]
        pairs = list(zip(x, feats))
        if mask is None:
            msk.fill_(1.0)
        else:
            pairs.append((msk, mask))
        if labels is not None:
            if labels.device.type == "cpu" and labels.numel():   # a host batch: cheap to check before it is copied
                lo, hi = int(labels.min()), int(labels.max())
                if lo < 0 or hi >= self.plan.C:
                    raise IndexError(f"Target {lo if lo < 0 else hi} is out of bounds.")   # F.cross_entropy's message
            pairs.append((lab, labels))
        moved = sum(src.numel() * src.element_size() for _, src in pairs if src.device.type == "cpu")
        if all(src.dtype == dst.dtype and src.is_contiguous() and src.numel() == dst.numel()
               and (src.device.type != "cpu" or src.is_pinned()) for dst, src in pairs):
            # one library call issues all the asynchronous copies (no per-tensor dispatch on the host)
            import ctypes
            triples = [(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size()) for dst, src in pairs]
            one_alloc = lambda ts: len({t.untyped_storage().data_ptr() for t in ts}) == 1   # noqa: E731
            if self._packed_slots and one_alloc([d for d, _ in pairs]) and one_alloc([s_ for _, s_ in pairs]):
                # pinned_batch() source and a one-buffer slot (views of ONE allocation on each side — tensors of
                # separate allocations that merely sit next to each other must not be merged): runs adjacent on both
                # sides go out as one transfer
                merged = [triples[0]]
                for d, s_, sz in triples[1:]:
                    d0, s0, sz0 = merged[-1]
                    if d == d0 + sz0 and s_ == s0 + sz0:
                        merged[-1] = (d0, s0, sz0 + sz)
                    else:
                        merged.append((d, s_, sz))
                triples = merged
            n = len(triples)
            dsts = (ctypes.c_void_p * n)(*[t[0] for t in triples])
            srcs = (ctypes.c_void_p * n)(*[t[1] for t in triples])
            sizes = (ctypes.c_size_t * n)(*[t[2] for t in triples])
            N.check(N.lib().msf_memcpy_batch(dsts, srcs, sizes, n, ops._stream()))
        else:
            for dst, src in pairs:
                dst.copy_(src, non_blocking=True)
        return moved

    def _replay_train(self, slot: int = 0) -> None:
        if not self.use_graph:
            self._enqueue_train_step(slot)
            return
        if self._train_graphs[slot] is None:
            self._replay_train_capture_only(slot)
        self._train_graphs[slot].replay()

    def add_resident_batch(self, features, mask: Optional[torch.Tensor], labels: torch.Tensor) -> int:
        """Register a batch that already lives in device memory (fp32 features (B, D_m), fp32 mask (B, M),
        int64 labels (B)) as an input slot of its own: `train_step_slot(slot)` then trains on it in place —
        no copy into the static buffers.  For datasets that fit in HBM (PAMAP2 does many times over)."""
        feats = [features[m] for m in self.plan.names] if isinstance(features, dict) else list(features)
        B = self.batch
        xs = []
        for f, d in zip(feats, self.plan.dims):
            ok_dtype = f.dtype == feats[0].dtype and (f.dtype == torch.float32 or (
                f.dtype == torch.bfloat16 and self.prec == N.MSF_PREC_BF16))
            if f.device != self.dev or not ok_dtype or tuple(f.shape) != (B, d) or not f.is_contiguous():
                raise ValueError("resident features must be contiguous fp32 (or, on the tensor-core path, bf16) "
                                 "(batch, in_dim) tensors on the engine's device, all of one dtype")
            xs.append(f)
        msk = torch.ones(B, self.plan.M, dtype=torch.float32, device=self.dev) if mask is None else mask
        if msk.device != self.dev or msk.dtype != torch.float32 or tuple(msk.shape) != (B, self.plan.M) \
                or not msk.is_contiguous():
            raise ValueError("resident mask must be a contiguous fp32 (batch, modalities) tensor on the engine's device")
        if labels.device != self.dev or labels.dtype != torch.int64 or tuple(labels.shape) != (B,):
            raise ValueError("resident labels must be an int64 (batch,) tensor on the engine's device")
        self._slots.append((xs, msk, labels))
        self._train_graphs.append(None)
        return len(self._slots) - 1

    def train_step_slot(self, slot: int) -> torch.Tensor:
        """One optimizer step on the batch of input slot `slot` (see add_resident_batch)."""
        self._replay_train(slot)
        return self.loss

    def train_slots(self, slots: Sequence[int]) -> torch.Tensor:
        """One optimizer step per listed input slot, in order, as ONE graph launch (an epoch over batches that live
        in HBM, `add_resident_batch`).  Inside the graph the first kernel of step i+1 is a programmatic dependent of
        the optimizer launch of step i, exactly as in the eager stream order, so the gap between two graph launches
        is paid once per call instead of once per step.  Returns the per-step mean losses (device tensor,
        `len(slots)`); every step does the full work of `train_step_slot`."""
        key = tuple(int(s) for s in slots)
        if not key:
            return torch.empty(0, dtype=torch.float32, device=self.dev)
        if not self.use_graph:
            out = torch.empty(len(key), dtype=torch.float32, device=self.dev)
            for i, s in enumerate(key):
                self._enqueue_train_step(s)
                out[i:i + 1].copy_(self.loss)
            return out
        entry = self._epoch_graphs.get(key)
        if entry is None:
            losses = torch.zeros(len(key), dtype=torch.float32, device=self.dev)

            def enqueue():
                saved = dict(self._loss_ptr)
                try:
                    for i, s in enumerate(key):
                        self._loss_ptr[s] = losses[i:i + 1].data_ptr()
                        self._enqueue_train_step(s)
                finally:
                    self._loss_ptr = saved

            snap = self._snapshot()
            graph = self._capture(enqueue)
            self._restore(snap)
            entry = self._epoch_graphs[key] = (graph, losses)
        entry[0].replay()
        return entry[1]

    def train_step_resident(self) -> torch.Tensor:
        """One optimizer step on the batch already in the static buffers.
        Returns the (device) mean loss of this rank's shard scaled to the global batch."""
        self._replay_train(0)
        return self.loss

    def train_step(self, features, mask, labels) -> torch.Tensor:
        self.load_batch(features, mask, labels)
        return self.train_step_resident()

    def train_step_accumulated(self, micro_batches) -> torch.Tensor:
        """One optimizer step over several micro-batches (config/base.yaml:75 ``accumulate_grad_batches``,
        src/train.py:519-521): every micro-batch (features, mask, labels), each of the engine's batch size, runs
        forward + backward with its loss scaled by 1 / len(micro_batches); the gradients are summed on the device
        (msf_grad_accumulate) and clipped / applied once.  Returns the mean of the micro-batch losses (device)."""
        micro_batches = list(micro_batches)
        k = len(micro_batches)
        if k == 0:
            raise ValueError("train_step_accumulated needs at least one micro-batch")
        if getattr(self, "_grad_micro", None) is None:
            self._grad_micro = torch.zeros_like(self.grad)
            self._loss_acc = torch.zeros(1, dtype=torch.float32, device=self.dev)
        lib, n = N.lib(), self.plan.total
        for i, (f, m, y) in enumerate(micro_batches):
            self.load_batch(f, m, y)
            self._enqueue_pass(0, grad=self._grad_micro, micro_batches=k)
            N.check(lib.msf_grad_accumulate(self.grad.data_ptr(), self._grad_micro.data_ptr(), n,
                                            self._loss_acc.data_ptr(), self.loss.data_ptr(), 1.0 / k, int(i == 0),
                                            ops._stream()))
            if i + 1 < k:
                self.state[1] += 1   # the next micro-batch draws fresh dropout masks (Philox offset; stream-ordered)
        self._enqueue_optimizer(False)
        return self._loss_acc

    def train_stream(self, batches):
        """Host-facing training loop: `batches` yields (features, mask, labels) on the host (pinned memory
        makes the copies asynchronous); yields one Python-float loss per batch, in order.

        Two-deep pipeline over two private input slots: while the graph captured over slot s computes batch i,
        the host->device copies of batch i+1 run on a copy stream into the other slot, and the loss of batch i
        is picked up after batch i+1 has been enqueued, so neither PCIe direction stalls the compute stream.
        The step writes its loss straight into pinned host memory (the loss pointer of these two graphs is a
        mapped host address), so no copy call sits between steps.  Every batch still crosses host->device and
        every loss device->host."""
        dev = self.dev
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
            self._ev_ready = [torch.cuda.Event() for _ in range(2)]   # slot inputs landed
            self._ev_done = [torch.cuda.Event() for _ in range(2)]    # graph over the slot finished (loss on the host)
            f32 = dict(dtype=torch.float32, device=dev)
            self._stream_slots = []
            for s in range(2):
                self._slots.append((self._one_buffer_batch(host=False) if self._packed_slots else None)
                                   or ([torch.zeros(self.batch, d, dtype=self.feature_dtype, device=dev)
                                        for d in self.plan.dims],
                                       torch.ones(self.batch, self.plan.M, **f32),
                                       torch.zeros(self.batch, dtype=torch.int64, device=dev)))
                self._train_graphs.append(None)
                slot = len(self._slots) - 1
                self._stream_slots.append(slot)
                self._loss_ptr[slot] = self._loss_host[s:s + 1].data_ptr()   # UVA: host-pinned == device-visible
        if self.use_graph:
            for slot in self._stream_slots:                           # capture before the pipeline starts
                if self._train_graphs[slot] is None:
                    self._replay_train_capture_only(slot)
        cs = self._copy_stream
        main = torch.cuda.current_stream(dev)
        cs.wait_stream(main)

        import ctypes
        lib = N.lib()
        cs_handle = ctypes.c_void_p(cs.cuda_stream)
        loss_np = self._loss_host.numpy()        # same pinned memory, cheap scalar reads
        dst_tables = []
        for slot in self._stream_slots:          # destination pointers / sizes never change
            x, msk, lab = self._slots[slot]
            tens = list(x) + [msk, lab]
            dst_tables.append(((ctypes.c_void_p * len(tens))(*[t.data_ptr() for t in tens]),
                               (ctypes.c_size_t * len(tens))(*[t.numel() * t.element_size() for t in tens]),
                               [(t.dtype, t.numel()) for t in tens]))
        one_dst, one_size = (ctypes.c_void_p * 1)(), (ctypes.c_size_t * 1)()

        def adjacent(tens):   # views of one allocation laid out back to back in this order (pinned_batch / one-buffer slots)
            return (len({t.untyped_storage().data_ptr() for t in tens}) == 1 and
                    all(b.data_ptr() == a.data_ptr() + a.numel() * a.element_size() for a, b in zip(tens, tens[1:])))

        slot_adjacent = [self._packed_slots and adjacent(list(self._slots[s][0]) + list(self._slots[s][1:]))
                         for s in self._stream_slots]

        def stage(batch, k):
            cs.wait_event(self._ev_done[k])      # no-op until the slot's first step has been recorded
            feats = [batch[0][m] for m in self.plan.names] if isinstance(batch[0], dict) else list(batch[0])
            srcs = feats + [batch[1], batch[2]]
            dsts, sizes, meta = dst_tables[k]
            if batch[1] is not None and all(
                    t.dtype == d and t.numel() == n and t.is_contiguous() and (t.device.type != "cpu" or t.is_pinned())
                    for t, (d, n) in zip(srcs, meta)):
                # one library call issues every host->device copy of the batch on the copy stream
                if slot_adjacent[k] and adjacent(srcs):   # the whole batch as one transfer
                    one_dst[0], one_size[0] = dsts[0], sum(sizes)
                    ptrs = (ctypes.c_void_p * 1)(srcs[0].data_ptr())
                    N.check(lib.msf_memcpy_batch(one_dst, ptrs, one_size, 1, cs_handle))
                else:
                    ptrs = (ctypes.c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
                    N.check(lib.msf_memcpy_batch(dsts, ptrs, sizes, len(srcs), cs_handle))
            else:
                with torch.cuda.stream(cs):
                    self.load_batch(batch[0], batch[1], batch[2], slot=self._stream_slots[k])
            self._ev_ready[k].record(cs)

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        stage(nxt, 0)
        k, pending = 0, None
        while nxt is not None:
            main.wait_event(self._ev_ready[k])
            self._replay_train(self._stream_slots[k])
            self._ev_done[k].record(main)
            nxt = next(it, None)
            if nxt is not None:
                stage(nxt, k ^ 1)
            if pending is not None:
                self._ev_done[pending].synchronize()
                yield float(loss_np[pending])
            pending, k = k, k ^ 1
        self._ev_done[pending].synchronize()
        yield float(loss_np[pending])

    def _snapshot(self):
        """Everything a train step changes.  The compute arena is saved as it is, not re-packed from the master:
        under the sharded step a rank's master is only current for the tiles it owns."""
        return tuple(t.clone() for t in (self.arena, self.exp_avg, self.exp_avg_sq, self.state)
                     + ((self.arena_bf16,) if self.arena_bf16 is not None else ()))

    def _restore(self, snap) -> None:
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            torch.distributed.barrier(self.pg)   # no peer is still pushing into the arenas being restored
        for dst, src in zip((self.arena, self.exp_avg, self.exp_avg_sq, self.state)
                            + ((self.arena_bf16,) if self.arena_bf16 is not None else ()), snap):
            dst.copy_(src)
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            torch.distributed.barrier(self.pg)

    def _replay_train_capture_only(self, slot: int) -> None:
        """Capture the train graph over `slot` without leaving a net model update behind."""
        snap = self._snapshot()
        self._train_graphs[slot] = self._capture(lambda: self._enqueue_train_step(slot))
        self._restore(snap)
        self._train_graph = self._train_graphs[0]

    def infer_resident(self):
        if self.use_graph:
            if self._infer_graph is None:
                self._infer_graph = self._capture(self._enqueue_inference)
            self._infer_graph.replay()
        else:
            self._enqueue_inference()
        return self.logits, self.conf, self.pred

    def infer(self, features, mask):
        self.load_batch(features, mask)
        return self.infer_resident()

    def infer_subset(self, features, present: Sequence[int]):
        """Inference with only the modalities `present` (indices into plan.names) available for the WHOLE batch:
        the missing-modality sweep of src/eval.py:342-404.  Same result as `infer` with the corresponding
        uniform mask; the work of the absent modalities is skipped (msf_fusion_infer_pass present_hint).
        `features=None` keeps the features already in the static buffers (a sweep over the same batch)."""
        bits = 0
        for m in present:
            bits |= 1 << int(m)
        if bits == 0 or bits >= (1 << self.plan.M):
            raise ValueError("present must name at least one valid modality")
        mask = self._subset_masks.get(bits)
        if mask is None:
            mask = torch.zeros(self.batch, self.plan.M, dtype=torch.float32, device=self.dev)
            mask[:, [int(m) for m in present]] = 1.0
            self._subset_masks[bits] = mask
        if features is None:
            self.mask.copy_(mask, non_blocking=True)
        else:
            self.load_batch(features, mask)
        if not self.use_graph or self.prec != N.MSF_PREC_BF16:
            self._enqueue_inference(bits if self.prec == N.MSF_PREC_BF16 else 0)
        else:
            if bits not in self._subset_graphs:
                self._subset_graphs[bits] = self._capture(lambda: self._enqueue_inference(bits))
            self._subset_graphs[bits].replay()
        return self.logits, self.conf, self.pred

    def fold_for_inference(self) -> None:
        """Fold every attention module's value_proj -> out_proj pair into one matrix for the subset sweep
        (msf_fusion_infer_folded): Wov = Wo Wv (bf16 [pairs][H][H]) and cv = Wo bv + bo, from the CURRENT fp32 master
        weights.  Call again after the weights change (training steps, load_state_dict)."""
        plan, H, M = self.plan, self.plan.H, self.plan.M
        names = plan.names
        mods = self.model.attention_modules
        wov = torch.zeros(M * (M - 1), H, H, dtype=torch.float32, device=self.dev)
        self._fold_cv = torch.zeros(M, M, H, dtype=torch.float32, device=self.dev)   # [q][k]: Wo bv + bo
        self._fold_bo = torch.zeros(M, M, H, dtype=torch.float32, device=self.dev)   # [q][k]: bo alone
        with torch.no_grad():
            for q in range(M):
                for k in range(M):
                    key = f"{names[q]}_to_{names[k]}"
                    if q == k or key not in mods:
                        continue
                    att = mods[key]
                    pi = q * (M - 1) + (k if k < q else k - 1)
                    wo, wv = att.out_proj.weight.detach().float(), att.value_proj.weight.detach().float()
                    wov[pi] = wo @ wv
                    self._fold_cv[q, k] = wo @ att.value_proj.bias.detach().float() + att.out_proj.bias.detach().float()
                    self._fold_bo[q, k] = att.out_proj.bias.detach().float()
        self._fold_wov = wov.to(torch.bfloat16).contiguous()
        self._fold_bias = {}
        self._fold_graphs = {}

    def _fold_bias_for(self, bits: int) -> torch.Tensor:
        got = self._fold_bias.get(bits)
        if got is None:
            M = self.plan.M
            pres = torch.tensor([(bits >> m) & 1 for m in range(M)], dtype=torch.float32, device=self.dev)
            # row q: present keys contribute Wo bv + bo, absent keys their out_proj bias (attention gate 0)
            got = (self._fold_cv * pres.view(1, M, 1) + self._fold_bo * (1.0 - pres).view(1, M, 1)).sum(1).contiguous()
            self._fold_bias[bits] = got
        return got

    def _enqueue_inference_folded(self, bits: int, flags: int) -> None:
        c = self._call(False)
        c.logits = self.logits.data_ptr()
        bias = self._fold_bias_for(bits)
        N.check(N.lib().msf_fusion_infer_folded(ctypes_ref(self.plan.shape), ctypes_ref(c), self._fold_wov.data_ptr(),
                                                bias.data_ptr(), bits, flags, self.conf.data_ptr(), self.pred.data_ptr(),
                                                ops._stream()))

    def infer_sweep(self, features, subsets: Sequence[Sequence[int]], on_subset=None):
        """The missing-modality sweep of src/eval.py:342-404 over ONE batch: for every subset (modality indices present
        in every window) logits / confidences / predictions as `infer_subset` gives them, with the work the subsets
        share done once — the projections of all modalities are computed a single time, and every subset then costs
        one folded pair GEMM (msf_fusion_infer_folded: value_proj and out_proj of each module as one matrix) plus
        the head kernel.  `on_subset(index, subset)` runs after each subset on the same stream (e.g. eval_update /
        ece_bins on self.logits).  `features=None` keeps the batch already in the static buffers."""
        if self.prec != N.MSF_PREC_BF16:
            raise N.MsfError("infer_sweep needs the tensor-core path (precision='bf16')")
        if getattr(self, "_fold_wov", None) is None:
            self.fold_for_inference()
        M = self.plan.M
        full = (1 << M) - 1
        ones = self._subset_mask(full)
        if features is not None:
            self.load_batch(features, ones)
        else:
            self.mask.copy_(ones, non_blocking=True)
        self._run_folded(full, N.MSF_FOLD_PROJECTIONS_ONLY)        # every modality projected once for the whole sweep
        for i, sub in enumerate(subsets):
            bits = 0
            for m in sub:
                bits |= 1 << int(m)
            if bits == 0 or bits > full:
                raise ValueError("every subset must name at least one valid modality")
            self.mask.copy_(self._subset_mask(bits), non_blocking=True)
            self._run_folded(bits, N.MSF_FOLD_REUSE_PROJECTIONS)
            if on_subset is not None:
                on_subset(i, sub)
        return self.logits, self.conf, self.pred

    def _subset_mask(self, bits: int) -> torch.Tensor:
        mask = self._subset_masks.get(bits)
        if mask is None:
            mask = torch.zeros(self.batch, self.plan.M, dtype=torch.float32, device=self.dev)
            mask[:, [m for m in range(self.plan.M) if (bits >> m) & 1]] = 1.0
            self._subset_masks[bits] = mask
        return mask

    def _run_folded(self, bits: int, flags: int) -> None:
        if not self.use_graph:
            self._enqueue_inference_folded(bits, flags)
            return
        key = (bits, flags)
        if key not in self._fold_graphs:
            self._fold_bias_for(bits)   # allocate outside the capture
            self._fold_graphs[key] = self._capture(lambda: self._enqueue_inference_folded(bits, flags))
        self._fold_graphs[key].replay()

    def eval_update(self, labels: torch.Tensor, stats: Optional["ops.EvalStats"] = None) -> "ops.EvalStats":
        """Fold the last inference pass (self.logits) and its labels into device-side evaluation statistics
        (confusion counts -> accuracy / macro-F1, NLL sum, ECE bins: ops.EvalStats, msf_eval_accumulate) — the
        per-batch body of evaluate_model (src/eval.py:39-130) without host round trips.  Under data parallelism call
        ``stats.all_reduce()`` once at the end (integer sums)."""
        if stats is None:
            stats = ops.EvalStats(self.plan.C, device=self.dev)
        stats.update(self.logits, labels)
        return stats

    def ece_bins(self, labels: torch.Tensor, edges: Sequence[float], out=None) -> torch.Tensor:
        """Shard-local binning of the last inference, then one integer all-reduce
        of the (3, num_bins) statistics."""
        stats = ops.ece_bin(self.conf, self.pred, labels.to(self.dev), edges, out=out)
        if self.world > 1:
            torch.distributed.all_reduce(stats, group=self.pg)
        return stats


def ctypes_ref(obj):
    import ctypes
    return ctypes.byref(obj)


def shard_batch(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous, equal shards (equal sizes keep mean-of-means == global mean)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)
