"""torch-facing wrappers over the C ABI: device-pointer plumbing, workspaces and
``torch.autograd.Function`` glue.  PyTorch only owns memory and streams here;
every FLOP of the path runs in ``libmsf_b200.so``.
"""
from __future__ import annotations

import ctypes
import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N

PRECISIONS = {"fp32": N.MSF_PREC_F32, "bf16": N.MSF_PREC_BF16}


def require_cuda(what: str = "this operation") -> torch.device:
    """The product has no CPU path: fail loudly instead of falling back."""
    if not torch.cuda.is_available():
        raise N.MsfError(
            f"{what} needs a CUDA device (sm_100a); msf_b200 has no CPU or PyTorch-eager fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


# ---------------------------------------------------------------------------
# HybridFusion plan: shape struct + master-arena slots
# ---------------------------------------------------------------------------
class FusionPlan:
    """Static description of one HybridFusion instance (src/fusion.py:267-329)."""

    def __init__(self, names: Sequence[str], dims: Sequence[int], hidden: int, heads: int,
                 classes: int, present_pairs: Sequence[Tuple[int, int]]):
        self.names = list(names)
        self.dims = [int(d) for d in dims]
        self.M, self.H, self.heads, self.C = len(self.names), int(hidden), int(heads), int(classes)
        if not 1 <= self.M <= N.MSF_MAX_MODALITIES:
            raise ValueError(f"msf_b200 supports 1..{N.MSF_MAX_MODALITIES} modalities, got {self.M}")
        self.present = sorted(set(present_pairs))
        bits = 0
        for q, k in self.present:
            bits |= 1 << (q * self.M + k)
        self.shape = N.FusionShape()
        self.shape.num_modalities, self.shape.hidden = self.M, self.H
        self.shape.num_heads, self.shape.num_classes = self.heads, self.C
        for i, d in enumerate(self.dims):
            self.shape.in_dims[i] = d
        self.shape.pair_present = bits
        cnt = ctypes.c_int64()
        N.check(N.lib().msf_fusion_param_count(ctypes.byref(self.shape), ctypes.byref(cnt)))
        self.total = cnt.value
        self.slots = self._slots()  # [(state_dict key, offset, shape)] in master order, present tensors only
        self._tables: Dict[tuple, torch.Tensor] = {}

    def key(self):
        return (tuple(self.names), tuple(self.dims), self.H, self.heads, self.C, tuple(self.present))

    def _offset(self, kind: int, idx: int) -> int:
        off = ctypes.c_int64()
        N.check(N.lib().msf_fusion_param_offset(ctypes.byref(self.shape), kind, idx, ctypes.byref(off)))
        return off.value

    def _slots(self):
        H, M, C = self.H, self.M, self.C
        out = []
        for m, name in enumerate(self.names):
            out.append((f"projections.{name}.0.weight", self._offset(0, m), (H, self.dims[m])))
            out.append((f"projections.{name}.0.bias", self._offset(1, m), (H,)))
        for q, k in itertools.permutations(range(M), 2):
            if (q, k) not in self.present:
                continue
            pre = f"attention_modules.{self.names[q]}_to_{self.names[k]}"
            for w, proj in enumerate(("query_proj", "key_proj", "value_proj", "out_proj")):
                out.append((f"{pre}.{proj}.weight", self._offset(2 + 2 * w, q * M + k), (H, H)))
                out.append((f"{pre}.{proj}.bias", self._offset(3 + 2 * w, q * M + k), (H,)))
        for m, name in enumerate(self.names):
            out.append((f"gating_layers.{name}.weight", self._offset(10, m), (1, H)))
            out.append((f"gating_layers.{name}.bias", self._offset(11, m), (1,)))
        out.append(("classifier.0.weight", self._offset(12, 0), (H, H)))
        out.append(("classifier.0.bias", self._offset(13, 0), (H,)))
        out.append(("classifier.3.weight", self._offset(14, 0), (C, H)))
        out.append(("classifier.3.bias", self._offset(15, 0), (C,)))
        return out

    @property
    def dense(self) -> bool:
        return len(self.present) == self.M * (self.M - 1)

    def workspace_bytes(self, batch: int, precision: int) -> int:
        b = ctypes.c_size_t()
        N.check(N.lib().msf_fusion_workspace_bytes(ctypes.byref(self.shape), batch, precision, ctypes.byref(b)))
        return b.value

    def bf16_arena_bytes(self) -> int:
        b = ctypes.c_size_t()
        N.check(N.lib().msf_fusion_bf16_arena_bytes(ctypes.byref(self.shape), ctypes.byref(b)))
        return b.value

    # -- per-tensor storage <-> flat master arena --------------------------------
    def _table(self, tensors: Sequence[torch.Tensor]) -> torch.Tensor:
        ptrs = tuple(t.data_ptr() for t in tensors)
        tab = self._tables.get(ptrs)
        if tab is None:
            rows = [[t.data_ptr(), off, t.numel()] for t, (_, off, _) in zip(tensors, self.slots)]
            tab = torch.tensor(rows, dtype=torch.int64).to(tensors[0].device)
            if len(self._tables) > 16:
                self._tables.clear()
            self._tables[ptrs] = tab
        return tab

    def gather(self, tensors: Sequence[torch.Tensor]) -> torch.Tensor:
        """Pack per-parameter tensors (slot order) into a fresh master arena."""
        dev = tensors[0].device
        alloc = torch.empty if self.dense else torch.zeros
        arena = alloc(self.total, dtype=torch.float32, device=dev)
        N.check(N.lib().msf_arena_gather(_p(self._table(tensors)), len(tensors), self.total, _p(arena), _stream()))
        return arena

    def scatter(self, arena: torch.Tensor, tensors: Sequence[torch.Tensor]) -> None:
        N.check(N.lib().msf_arena_scatter(_p(self._table(tensors)), len(tensors), self.total, _p(arena), _stream()))

    def pack_bf16(self, arena: torch.Tensor) -> torch.Tensor:
        out = torch.zeros(self.bf16_arena_bytes(), dtype=torch.uint8, device=arena.device)
        N.check(N.lib().msf_fusion_pack_bf16(ctypes.byref(self.shape), _p(arena), _p(out), _stream()))
        return out


_PLANS: Dict[tuple, FusionPlan] = {}


def get_plan(names, dims, hidden, heads, classes, present) -> FusionPlan:
    """Process-wide plan cache: modules stay free of ctypes state (deepcopy / pickle safe)."""
    key = (tuple(names), tuple(int(d) for d in dims), int(hidden), int(heads), int(classes),
           tuple(sorted(set(present))))
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) > 64:
            _PLANS.clear()
        plan = _PLANS[key] = FusionPlan(names, dims, hidden, heads, classes, present)
    return plan


def layer_norm_fused(plan: "FusionPlan", precision: int) -> bool:
    """True if the projection kernel applies msf_fusion_call.ln_* itself for this shape / precision."""
    return N.lib().msf_fusion_layer_norm_fused(ctypes.byref(plan.shape), precision) == 1


def _make_call(plan: FusionPlan, batch: int, precision: int, training: bool, p: float, seed: int,
               offset: int, arena: torch.Tensor, arena_bf16: Optional[torch.Tensor],
               xs: Sequence[torch.Tensor], mask: Optional[torch.Tensor], ws: torch.Tensor,
               ln: Optional[Sequence[Optional[tuple]]] = None, ln_eps: float = 1e-5) -> N.FusionCall:
    """``ln``: per modality ``(weight, bias)`` fp32 device tensors (either may be None) or None — the LayerNorm
    train.py puts between encoder and fusion, applied inside the projection kernel (layer_norm_fused())."""
    c = N.FusionCall()
    if ln is not None:
        for i, wb in enumerate(ln):
            if wb is not None:
                c.ln_weight[i], c.ln_bias[i] = _p(wb[0]), _p(wb[1])
        c.ln_eps = float(ln_eps)
    c.batch, c.precision, c.training, c.dropout_p = batch, precision, int(bool(training)), float(p)
    c.seed, c.offset = seed & (2**64 - 1), offset & (2**64 - 1)
    c.params, c.params_bf16 = _p(arena), _p(arena_bf16)
    for i, x in enumerate(xs):
        c.x[i] = _p(x)
    kinds = {x.dtype for x in xs}
    if kinds == {torch.bfloat16}:
        c.x_bf16 = 1      # features stored as bf16 rows (msf_fusion_call.x_bf16)
    elif kinds != {torch.float32}:
        raise N.MsfError(f"features must be all float32 or all bfloat16, got {sorted(str(k) for k in kinds)}")
    c.mask = _p(mask)
    c.workspace, c.workspace_bytes = _p(ws), ws.numel()
    return c


def fusion_forward_raw(plan: FusionPlan, arena: torch.Tensor, xs: Sequence[torch.Tensor],
                       mask: Optional[torch.Tensor], *, precision: int = N.MSF_PREC_F32,
                       training: bool = False, p: float = 0.0, seed: int = 0, offset: int = 0,
                       arena_bf16: Optional[torch.Tensor] = None, want_aux: bool = True,
                       workspace: Optional[torch.Tensor] = None, ln=None, ln_eps: float = 1e-5):
    """One msf_fusion_forward call on already-prepared device buffers."""
    dev = arena.device
    B = xs[0].shape[0]
    if workspace is None:
        workspace = torch.empty(plan.workspace_bytes(B, precision), dtype=torch.uint8, device=dev)
    logits = torch.empty(B, plan.C, dtype=torch.float32, device=dev)
    fw = gates = None
    call = _make_call(plan, B, precision, training, p, seed, offset, arena, arena_bf16, xs, mask, workspace,
                      ln=ln, ln_eps=ln_eps)
    call.logits = _p(logits)
    if want_aux:
        fw = torch.empty(B, plan.M, dtype=torch.float32, device=dev)
        gates = torch.empty(plan.M * (plan.M - 1), B, plan.heads, dtype=torch.float32, device=dev)
        call.fusion_weights, call.attn_gates = _p(fw), _p(gates)
    N.check(N.lib().msf_fusion_forward(ctypes.byref(plan.shape), ctypes.byref(call), _stream()))
    return logits, fw, gates, workspace


def fusion_backward_raw(plan: FusionPlan, arena: torch.Tensor, xs: Sequence[torch.Tensor],
                        mask: Optional[torch.Tensor], workspace: torch.Tensor, grad_logits: torch.Tensor,
                        *, precision: int = N.MSF_PREC_F32, training: bool = False, p: float = 0.0,
                        seed: int = 0, offset: int = 0, arena_bf16: Optional[torch.Tensor] = None,
                        need_dx: Sequence[bool] = (), grad_arena: Optional[torch.Tensor] = None):
    dev = arena.device
    B = xs[0].shape[0]
    if grad_arena is None:
        grad_arena = torch.empty(plan.total, dtype=torch.float32, device=dev)
    call = _make_call(plan, B, precision, training, p, seed, offset, arena, arena_bf16, xs, mask, workspace)
    call.grad_logits, call.grad_params = _p(grad_logits), _p(grad_arena)
    dxs: List[Optional[torch.Tensor]] = []
    for i, x in enumerate(xs):
        need = bool(need_dx[i]) if i < len(need_dx) else False
        dx = torch.empty_like(x) if need else None
        call.grad_x[i] = _p(dx)
        dxs.append(dx)
    N.check(N.lib().msf_fusion_backward(ctypes.byref(plan.shape), ctypes.byref(call), _stream()))
    return grad_arena, dxs


def fusion_train_pass_raw(plan: FusionPlan, arena: torch.Tensor, xs: Sequence[torch.Tensor],
                          mask: Optional[torch.Tensor], labels: torch.Tensor, *, smoothing: float = 0.0,
                          grad_scale: Optional[float] = None, precision: int = N.MSF_PREC_F32,
                          training: bool = True, p: float = 0.0, seed: int = 0, offset: int = 0,
                          arena_bf16: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                          grad_sq: Optional[torch.Tensor] = None):
    """One msf_fusion_train_pass call: forward + CE(label smoothing) + backward.
    Returns ``(logits, loss[1], grad_arena, fusion_weights, gates)``.  ``grad_sq`` (float64[1], fused path only)
    receives the sum of squares of the gradient arena."""
    dev = arena.device
    B = xs[0].shape[0]
    if workspace is None:
        workspace = torch.empty(plan.workspace_bytes(B, precision), dtype=torch.uint8, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    logits = torch.empty(B, plan.C, **f32)
    scratch = torch.empty(B, plan.C, **f32)
    row = torch.empty(B, **f32)
    loss = torch.empty(1, **f32)
    grad = torch.empty(plan.total, **f32)
    fw = torch.empty(B, plan.M, **f32)
    gates = torch.empty(plan.M * (plan.M - 1), B, plan.heads, **f32)
    call = _make_call(plan, B, precision, training, p, seed, offset, arena, arena_bf16, xs, mask, workspace)
    call.logits, call.grad_params = _p(logits), _p(grad)
    call.fusion_weights, call.attn_gates = _p(fw), _p(gates)
    call.grad_sq = _p(grad_sq)
    labels = labels.to(torch.int64).contiguous()
    scale = (1.0 / B) if grad_scale is None else grad_scale
    N.check(N.lib().msf_fusion_train_pass(ctypes.byref(plan.shape), ctypes.byref(call), _p(labels),
                                          float(smoothing), float(scale), _p(row), _p(loss), _p(scratch),
                                          0, _stream()))
    return logits, loss, grad, fw, gates


def fusion_infer_pass_raw(plan: FusionPlan, arena: torch.Tensor, xs: Sequence[torch.Tensor],
                          mask: Optional[torch.Tensor], *, precision: int = N.MSF_PREC_F32,
                          arena_bf16: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                          present_hint: int = 0, ln=None, ln_eps: float = 1e-5):
    """One msf_fusion_infer_pass call: ``(logits, conf, pred)`` (src/eval.py:84-90).  ``present_hint``: bit set
    of the modalities a batch-uniform mask marks present (0 = no hint)."""
    dev = arena.device
    B = xs[0].shape[0]
    if workspace is None:
        workspace = torch.empty(plan.workspace_bytes(B, precision), dtype=torch.uint8, device=dev)
    logits = torch.empty(B, plan.C, dtype=torch.float32, device=dev)
    conf = torch.empty(B, dtype=torch.float32, device=dev)
    pred = torch.empty(B, dtype=torch.int64, device=dev)
    call = _make_call(plan, B, precision, False, 0.0, 0, 0, arena, arena_bf16, xs, mask, workspace, ln=ln, ln_eps=ln_eps)
    call.logits = _p(logits)
    N.check(N.lib().msf_fusion_infer_pass(ctypes.byref(plan.shape), ctypes.byref(call), _p(conf), _p(pred),
                                          int(present_hint), _stream()))
    return logits, conf, pred


class HybridFusionFunction(torch.autograd.Function):
    """Autograd node for HybridFusion.forward (src/fusion.py:331-427).

    ``tensors`` = the M feature tensors followed by the parameters in
    ``plan.slots`` order, all fp32 CUDA.  Dead query/key projection parameters
    receive exact-zero gradient tensors (never ``None``), matching the
    reference's autograd (SURVEY.md §7 hard part 2).
    """

    @staticmethod
    def forward(ctx, plan: FusionPlan, cfg: dict, mask: Optional[torch.Tensor], *tensors):
        M = plan.M
        n_par = len(plan.slots)
        xs = [t.contiguous() for t in tensors[:M]]
        params = [t.contiguous() for t in tensors[M:M + n_par]]
        # optional fused input LayerNorm (cfg["ln"]): 2 * M trailing tensors, weight_m then bias_m (None = absent)
        ln = None
        if cfg.get("ln"):
            extra = tensors[M + n_par:]
            ln = [None if extra[m] is None and extra[M + m] is None else
                  (None if extra[m] is None else extra[m].contiguous(),
                   None if extra[M + m] is None else extra[M + m].contiguous()) for m in range(M)]
        precision = cfg["precision"]
        arena = plan.gather(params)
        arena_bf16 = plan.pack_bf16(arena) if precision == N.MSF_PREC_BF16 else None
        logits, fw, gates, ws = fusion_forward_raw(
            plan, arena, xs, mask, precision=precision, training=cfg["training"], p=cfg["p"],
            seed=cfg["seed"], offset=cfg["offset"], arena_bf16=arena_bf16, want_aux=True, ln=ln,
            ln_eps=cfg.get("ln_eps", 1e-5))
        ctx.plan, ctx.cfg, ctx.mask = plan, cfg, mask
        ctx.saved = (arena, arena_bf16, xs, ws, ln)
        ctx.param_shapes = [t.shape for t in tensors[M:M + n_par]]
        ctx.mark_non_differentiable(fw, gates)
        return logits, fw, gates

    @staticmethod
    def backward(ctx, grad_logits, _gfw, _ggates):
        plan, cfg = ctx.plan, ctx.cfg
        arena, arena_bf16, xs, ws, ln = ctx.saved
        M = plan.M
        n_par = len(plan.slots)
        need_dx = list(ctx.needs_input_grad[3:3 + M])
        need_ln = [False] * (2 * M)
        if ln is not None:
            need_ln = list(ctx.needs_input_grad[3 + M + n_par:3 + M + n_par + 2 * M])
            for m in range(M):   # the LayerNorm gradients need the gradient with respect to the normalised rows
                need_dx[m] = need_dx[m] or (ln[m] is not None and (need_ln[m] or need_ln[M + m]))
        g = grad_logits.to(torch.float32).contiguous()
        grad_arena, dxs = fusion_backward_raw(
            plan, arena, xs, ctx.mask, ws, g, precision=cfg["precision"], training=cfg["training"],
            p=cfg["p"], seed=cfg["seed"], offset=cfg["offset"], arena_bf16=arena_bf16, need_dx=need_dx)
        grads = [grad_arena[off:off + int(torch.Size(shape).numel())].view(shape)
                 for (_, off, _), shape in zip(plan.slots, ctx.param_shapes)]
        ln_grads = []
        if ln is not None:
            dws, dbs = [None] * M, [None] * M
            for m in range(M):
                if ln[m] is None or dxs[m] is None:
                    continue
                # dxs[m] is the gradient with respect to the normalised rows (include/msf_b200.h: ln_*)
                dxs[m], dws[m], dbs[m] = layer_norm_backward_raw(xs[m], ln[m][0], dxs[m], cfg.get("ln_eps", 1e-5))
                if ln[m][0] is None:
                    dws[m] = None
                if ln[m][1] is None:
                    dbs[m] = None
            ln_grads = dws + dbs
            dxs = [dx if ctx.needs_input_grad[3 + m] else None for m, dx in enumerate(dxs)]
        return (None, None, None, *dxs, *grads, *ln_grads)


def layer_norm_forward_raw(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                           eps: float = 1e-5) -> torch.Tensor:
    rows = x.numel() // x.shape[-1]
    y = torch.empty_like(x)
    N.check(N.lib().msf_layer_norm_forward(_p(x), _p(weight), _p(bias), _p(y), rows, x.shape[-1], float(eps), _stream()))
    return y


def layer_norm_backward_raw(x: torch.Tensor, weight: Optional[torch.Tensor], dy: torch.Tensor, eps: float = 1e-5):
    """``(dx, dweight, dbias)`` of y = LayerNorm(x) given dy (msf_layer_norm_backward)."""
    dim = x.shape[-1]
    rows = x.numel() // dim
    dx = torch.empty_like(x)
    dw = torch.empty(dim, dtype=torch.float32, device=x.device)
    db = torch.empty(dim, dtype=torch.float32, device=x.device)
    N.check(N.lib().msf_layer_norm_backward(_p(x), _p(weight), _p(dy.contiguous()), _p(dx), _p(dw), _p(db), rows, dim,
                                            float(eps), _stream()))
    return dx, dw, db


class LayerNormFunction(torch.autograd.Function):
    """nn.LayerNorm over the last dimension on msf_layer_norm_* (src/train.py:170-171,267-268)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps: float):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.eps, ctx.has = eps, (weight is not None, bias is not None)
        return layer_norm_forward_raw(x, weight, bias, eps)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        dx, dw, db = layer_norm_backward_raw(x, weight, gy.to(torch.float32), ctx.eps)
        return dx, (dw if ctx.has[0] else None), (db if ctx.has[1] else None), None


def layer_norm(x: torch.Tensor, weight: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
               eps: float = 1e-5) -> torch.Tensor:
    return LayerNormFunction.apply(x, weight, bias, eps)


class FramePoolFunction(torch.autograd.Function):
    """Temporal pooling of FrameEncoder (src/encoders.py:258-336) on msf_frame_pool_*: ``mode`` 0 attention (scores
    ``x . w + b``, masked softmax over the frames, weighted sum), 1 (masked) average, 2 (masked) maximum."""

    @staticmethod
    def forward(ctx, x, mask, w, b, mode: int):
        x = x.to(torch.float32).contiguous()
        B, T, H = x.shape
        dev = x.device
        m = None if mask is None else mask.detach().to(device=dev, dtype=torch.float32).expand(B, T).contiguous()
        wv = None if w is None else w.detach().to(torch.float32).reshape(-1).contiguous()
        bv = None if b is None else b.detach().to(torch.float32).reshape(-1).contiguous()
        pooled = torch.empty(B, H, dtype=torch.float32, device=dev)
        weights = torch.empty(B, T, dtype=torch.float32, device=dev) if mode != 2 else None
        argmax = torch.empty(B, H, dtype=torch.int32, device=dev) if mode == 2 else None
        N.check(N.lib().msf_frame_pool_forward(_p(x), _p(m), _p(wv), _p(bv), B, T, H, mode, _p(pooled), _p(weights),
                                               _p(argmax), _stream()))
        ctx.save_for_backward(x, wv, weights, argmax)
        ctx.mode, ctx.has = mode, (w is not None, b is not None)
        ctx.w_shape = None if w is None else w.shape
        ctx.b_shape = None if b is None else b.shape
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        x, wv, weights, argmax = ctx.saved_tensors
        B, T, H = x.shape
        dev = x.device
        dp = d_pooled.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        dw = db = scratch = None
        if ctx.mode == 0:
            dw = torch.empty(H, dtype=torch.float32, device=dev)
            db = torch.empty(1, dtype=torch.float32, device=dev)
            scratch = torch.empty(B * H + B, dtype=torch.float32, device=dev)
        N.check(N.lib().msf_frame_pool_backward(_p(x), _p(wv), _p(weights), _p(argmax), _p(dp), B, T, H, ctx.mode, _p(dx),
                                                _p(dw), _p(db), _p(scratch), _stream()))
        gw = dw.reshape(ctx.w_shape) if (ctx.mode == 0 and ctx.has[0]) else None
        gb = db.reshape(ctx.b_shape) if (ctx.mode == 0 and ctx.has[1]) else None
        return dx, None, gw, gb, None


def frame_pool(x: torch.Tensor, mask: Optional[torch.Tensor], mode: str, weight: Optional[torch.Tensor] = None,
               bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``(B, T, H) -> (B, H)``: ``mode`` "attention" (``weight`` (1, H) / ``bias`` (1,) of the scoring layer),
    "average" or "max", with an optional frame mask (B, T)."""
    require_cuda("frame_pool")
    return FramePoolFunction.apply(x, mask, weight, bias, {"attention": 0, "average": 1, "max": 2}[mode])


class BatchNormActFunction(torch.autograd.Function):
    """``dropout(relu(batch_norm(y)))`` behind a Linear layer on msf_bn_act_* (src/encoders.py:374-377)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, momentum, eps, training, relu, p, seed):
        y = y.contiguous()
        rows, cols = y.shape
        out = torch.empty_like(y)
        mean = torch.empty(cols, dtype=torch.float32, device=y.device)
        invstd = torch.empty(cols, dtype=torch.float32, device=y.device)
        scratch = torch.zeros(2 * cols, dtype=torch.float64, device=y.device)
        N.check(N.lib().msf_bn_act_forward(_p(y), _p(out), rows, cols, _p(gamma), _p(beta), _p(running_mean),
                                           _p(running_var), float(momentum), float(eps), int(training), int(relu),
                                           float(p), seed & (2**64 - 1), _p(mean), _p(invstd), _p(scratch), _stream()))
        ctx.save_for_backward(y, out, gamma, mean, invstd)
        ctx.cfg = (int(training), int(relu), float(p), gamma is not None, beta is not None)
        return out

    @staticmethod
    def backward(ctx, gout):
        y, out, gamma, mean, invstd = ctx.saved_tensors
        training, relu, p, has_g, has_b = ctx.cfg
        rows, cols = y.shape
        dy = torch.empty_like(y)
        dg = torch.empty(cols, dtype=torch.float32, device=y.device) if has_g else None
        db = torch.empty(cols, dtype=torch.float32, device=y.device) if has_b else None
        scratch = torch.zeros(2 * cols, dtype=torch.float64, device=y.device)
        N.check(N.lib().msf_bn_act_backward(_p(gout.to(torch.float32).contiguous()), _p(y), _p(out), rows, cols, _p(gamma),
                                            _p(mean), _p(invstd), training, relu, p, _p(dy), _p(dg), _p(db), _p(scratch),
                                            _stream()))
        return dy, dg, db, None, None, None, None, None, None, None, None


def batch_norm_act(y: torch.Tensor, weight, bias, running_mean, running_var, momentum: float, eps: float,
                   use_batch_stats: bool, relu: bool = True, p: float = 0.0) -> torch.Tensor:
    """``dropout_p(relu(batch_norm(y)))`` for a 2-D fp32 CUDA tensor through msf_bn_act_* (differentiable in ``y``,
    ``weight``, ``bias``).  ``use_batch_stats``: normalise with the batch statistics and, if running tensors are
    given, move them on (nn.BatchNorm1d training mode); otherwise normalise with the running statistics.  ``p`` > 0
    draws a fresh Philox dropout mask (only with batch statistics, i.e. in training mode)."""
    seed = int(torch.randint(0, 2**62, (1,)).item()) if (use_batch_stats and p > 0.0) else 0
    return BatchNormActFunction.apply(y, weight, bias, running_mean, running_var, momentum, eps, use_batch_stats, relu,
                                      p if use_batch_stats else 0.0, seed)


def adaptive_weights(feats: Sequence[torch.Tensor], gate_w: Sequence[torch.Tensor],
                     gate_b: Sequence[torch.Tensor], mask: torch.Tensor) -> torch.Tensor:
    """HybridFusion.compute_adaptive_weights (src/fusion.py:429-479) on the tail kernel."""
    M, (B, H) = len(feats), feats[0].shape
    agg = torch.stack([f.contiguous() for f in feats], dim=0).contiguous()        # [M][B][H]
    gw = torch.stack([w.reshape(-1) for w in gate_w], dim=0).contiguous()          # [M][H]
    gb = torch.cat([b.reshape(-1) for b in gate_b]).contiguous()                   # [M]
    out = torch.empty(B, M, dtype=torch.float32, device=agg.device)
    N.check(N.lib().msf_adaptive_weights(_p(agg), _p(gw), _p(gb), _p(mask.contiguous()), B, M, H, _p(out),
                                         _stream()))
    return out


def dropout_mask(seed: int, offset: int, site: int, sub: int, rows: int, cols: int, p: float,
                 device=None) -> torch.Tensor:
    """The multipliers the kernels draw for one dropout site (for oracle injection)."""
    dev = device or require_cuda("dropout_mask")
    out = torch.empty(rows, cols, dtype=torch.float32, device=dev)
    N.check(N.lib().msf_dropout_mask(seed & (2**64 - 1), offset & (2**64 - 1), site, sub, rows, cols,
                                     float(p), _p(out), _stream()))
    return out


# ---------------------------------------------------------------------------
# nn.Linear on the fp32 FFMA path (stand-alone attention / encoder projections)
# ---------------------------------------------------------------------------
class LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).to(torch.float32).contiguous()
        w = weight.to(torch.float32).contiguous()
        b = None if bias is None else bias.to(torch.float32).contiguous()
        y = torch.empty(x2.shape[0], w.shape[0], dtype=torch.float32, device=x2.device)
        N.check(N.lib().msf_linear_forward(_p(x2), _p(w), _p(b), _p(y), x2.shape[0], w.shape[1],
                                           w.shape[0], int(relu), _stream()))
        ctx.save_for_backward(x2, w, y if relu else None)
        ctx.relu, ctx.has_bias, ctx.in_shape = relu, bias is not None, x.shape
        return y.reshape(*lead, w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w, y = ctx.saved_tensors
        g = gy.reshape(-1, w.shape[0]).to(torch.float32).contiguous()
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty(w.shape[0], dtype=torch.float32, device=w.device) if ctx.has_bias else None
        scratch = torch.empty_like(g) if ctx.relu else None
        N.check(N.lib().msf_linear_backward(_p(x2), _p(w), _p(y), _p(g), _p(scratch), _p(dx), _p(dw),
                                            _p(db), x2.shape[0], w.shape[1], w.shape[0], int(ctx.relu),
                                            _stream()))
        return (None if dx is None else dx.reshape(ctx.in_shape)), dw, db, None


def linear(x, weight, bias=None, relu: bool = False):
    return LinearFunction.apply(x, weight, bias, relu)


class AttentionCoreFunction(torch.autograd.Function):
    """Scores -> key mask -> softmax -> NaN->0 -> dropout -> weights . V
    (src/attention.py:108-139) for projected q (B,Lq,H), k/v (B,Lk,H)."""

    @staticmethod
    def forward(ctx, q, k, v, mask, heads: int, p: float, training: bool, seed: int):
        q, k, v = (t.to(torch.float32).contiguous() for t in (q, k, v))
        B, Lq, H = q.shape
        Lk = k.shape[1]
        if mask is not None:
            mask = mask.detach().to(device=q.device, dtype=torch.float32).expand(B, Lk).contiguous()
        weights = torch.empty(B, heads, Lq, Lk, dtype=torch.float32, device=q.device)
        out = torch.empty(B, Lq, H, dtype=torch.float32, device=q.device)
        N.check(N.lib().msf_attention_core_forward(_p(q), _p(k), _p(v), _p(mask), B, Lq, Lk, H, heads,
                                                   float(p), int(training), seed, 0, _p(weights), _p(out),
                                                   _stream()))
        ctx.save_for_backward(q, k, v, mask, weights)
        ctx.cfg = (heads, float(p), int(training), seed)
        ctx.mark_non_differentiable(weights)
        return out, weights

    @staticmethod
    def backward(ctx, gout, _gw):
        q, k, v, mask, weights = ctx.saved_tensors
        heads, p, training, seed = ctx.cfg
        B, Lq, H = q.shape
        Lk = k.shape[1]
        g = gout.to(torch.float32).contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        scratch = torch.empty_like(weights)
        N.check(N.lib().msf_attention_core_backward(_p(q), _p(k), _p(v), _p(mask), B, Lq, Lk, H, heads, p,
                                                    training, seed, 0, _p(weights), _p(g), _p(scratch),
                                                    _p(dq), _p(dk), _p(dv), _stream()))
        return dq, dk, dv, None, None, None, None, None


def attention_core(q, k, v, mask, heads: int, p: float = 0.0, training: bool = False, seed: int = 0):
    return AttentionCoreFunction.apply(q, k, v, mask, heads, p, training, seed)


# ---------------------------------------------------------------------------
# loss / confidence / calibration binning / optimizer
# ---------------------------------------------------------------------------
def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, smoothing: float = 0.0,
                  grad_scale: Optional[float] = None):
    """Mean CE with label smoothing and its gradient w.r.t. logits
    (src/train.py:185-186,310).  Returns ``(loss[1], grad_logits)``."""
    B, C = logits.shape
    logits = logits.to(torch.float32).contiguous()
    labels = labels.to(torch.int64).contiguous()
    row = torch.empty(B, dtype=torch.float32, device=logits.device)
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits)
    scale = (1.0 / B) if grad_scale is None else grad_scale
    N.check(N.lib().msf_cross_entropy(_p(logits), _p(labels), B, C, float(smoothing), float(scale),
                                      _p(row), _p(loss), _p(grad), _stream()))
    return loss, grad


def softmax_conf_pred(logits: torch.Tensor):
    """``conf, pred = max(softmax(logits, 1), 1)`` (src/eval.py:89-90)."""
    B, C = logits.shape
    logits = logits.to(torch.float32).contiguous()
    conf = torch.empty(B, dtype=torch.float32, device=logits.device)
    pred = torch.empty(B, dtype=torch.int64, device=logits.device)
    N.check(N.lib().msf_softmax_conf_pred(_p(logits), B, C, _p(conf), _p(pred), _stream()))
    return conf, pred


def ece_bin(conf: torch.Tensor, pred: torch.Tensor, label: torch.Tensor, edges: Sequence[float],
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Single-pass binning (src/uncertainty.py:113-126,231-241).  Returns an
    int64 tensor ``(3, num_bins)`` on the device: counts, correct, and the
    Q32 fixed-point confidence sums; accumulates into ``out`` when given."""
    nb = len(edges) - 1
    dev = conf.device
    conf = conf.detach().to(torch.float32).contiguous()
    pred = pred.detach().to(torch.int64).contiguous()
    label = label.detach().to(torch.int64).contiguous()
    if out is None:
        out = torch.zeros(3, nb, dtype=torch.int64, device=dev)
    e = (ctypes.c_double * (nb + 1))(*[float(v) for v in edges])
    N.check(N.lib().msf_ece_bin(_p(conf), _p(pred), _p(label), conf.numel(), e, nb, _p(out[0]),
                                _p(out[1]), _p(out[2]), _stream()))
    return out


class EvalStats:
    """Device-side accumulators of msf_eval_accumulate: confusion matrix, NLL sum, ECE bins.  ``update(logits,
    labels)`` adds one batch in a single pass over the logits; ``metrics()`` reads the few integers back and returns
    what evaluate_model / compute_calibration_metrics report (src/eval.py:39-130, src/uncertainty.py:495-553):
    accuracy, f1_macro (sklearn's macro average with zero_division=0 over the classes seen in labels or
    predictions), loss (mean NLL per window), ece, mce, num_samples."""

    def __init__(self, classes: int, num_bins: int = 15, device=None):
        self.dev = device or require_cuda("EvalStats")
        self.C, self.nb = int(classes), int(num_bins)
        self.edges = torch.linspace(0.0, 1.0, self.nb + 1).double().tolist()   # fp32 linspace edges, uncertainty.py:113
        self.confusion = torch.zeros(self.C, self.C, dtype=torch.int64, device=self.dev)
        self.scalars = torch.zeros(3, dtype=torch.int64, device=self.dev)
        self.bins = torch.zeros(3, self.nb, dtype=torch.int64, device=self.dev)

    def reset(self) -> None:
        self.confusion.zero_()
        self.scalars.zero_()
        self.bins.zero_()

    def update(self, logits: torch.Tensor, labels: torch.Tensor, conf: Optional[torch.Tensor] = None,
               pred: Optional[torch.Tensor] = None) -> None:
        logits = logits.detach().to(device=self.dev, dtype=torch.float32).contiguous()
        labels = labels.detach().to(device=self.dev, dtype=torch.int64).contiguous()
        assert logits.dim() == 2 and logits.shape[1] == self.C and labels.numel() == logits.shape[0]
        e = (ctypes.c_double * (self.nb + 1))(*self.edges)
        N.check(N.lib().msf_eval_accumulate(_p(logits), _p(labels), logits.shape[0], self.C, e, self.nb, _p(conf), _p(pred),
                                            _p(self.confusion), _p(self.scalars), _p(self.bins), _stream()))

    def all_reduce(self, group=None) -> None:
        """Shard-local statistics -> global ones: integer sums, exact in any order."""
        for t in (self.confusion, self.scalars, self.bins):
            torch.distributed.all_reduce(t, group=group)

    def metrics(self) -> Dict[str, float]:
        cm = self.confusion.cpu().double()
        seen, nll_q24, bad = (int(v) for v in self.scalars.cpu().tolist())
        n = int(cm.sum().item())
        tp = cm.diag()
        support, predicted = cm.sum(1), cm.sum(0)
        present = (support + predicted) > 0
        denom = support + predicted
        f1 = torch.where(denom > 0, 2.0 * tp / denom.clamp_min(1.0), torch.zeros_like(tp))
        count, correct, csum = (self.bins[i].cpu().double() for i in range(3))
        csum = csum / 4294967296.0
        gap = torch.where(count > 0, (csum / count.clamp_min(1.0) - correct / count.clamp_min(1.0)).abs(), torch.zeros_like(count))
        total = float(count.sum().item())
        return {"accuracy": float(tp.sum().item()) / max(n, 1),
                "f1_macro": float(f1[present].mean().item()) if bool(present.any()) else 0.0,
                "loss": (nll_q24 / 16777216.0) / max(n, 1),
                "ece": float((gap * count).sum().item()) / max(total, 1.0),
                "mce": float(gap.max().item()) if total > 0 else 0.0,
                "num_samples": seen, "out_of_range_labels": bad}


def grad_sq_norm(grad: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.zeros(1, dtype=torch.float64, device=grad.device)
    N.check(N.lib().msf_grad_sq_norm(_p(grad), grad.numel(), _p(out), _stream()))
    return out


def adamw_step(params, grad, exp_avg, exp_avg_sq, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8,
               weight_decay=1e-4, grad_scale=1.0, max_norm=0.0, sq_norm: Optional[torch.Tensor] = None):
    """AdamW over flat fp32 arenas with optional global-norm clip
    (src/train.py:378-382,416-430)."""
    N.check(N.lib().msf_adamw_step(_p(params), _p(grad), _p(exp_avg), _p(exp_avg_sq), params.numel(),
                                   int(step), lr, beta1, beta2, eps, weight_decay, grad_scale, max_norm,
                                   _p(sq_norm), _stream()))


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, *, mn_major: bool = False, bias: Optional[torch.Tensor] = None,
              relu: bool = False, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """tcgen05 GEMM building block.  ``mn_major=False``: ``a (m,k) @ b (n,k).T`` (nn.Linear forward /
    dgrad); ``mn_major=True``: ``a (k,m).T @ b (k,n)`` (weight gradient).  bf16 operands, fp32 accumulate."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    if mn_major:
        (k, m), (k2, n) = a.shape, b.shape
    else:
        (m, k), (n, k2) = a.shape, b.shape
    assert k == k2, "contraction lengths differ"
    d = torch.empty(m, n, dtype=out_dtype, device=a.device)
    if bias is not None:
        bias = bias.to(torch.float32).contiguous()
    N.check(N.lib().msf_gemm_bf16(_p(a), _p(b), _p(d), int(out_dtype == torch.bfloat16), m, n, k, a.stride(0),
                                  b.stride(0), d.stride(0), int(mn_major), _p(bias), int(relu), _stream()))
    return d


def fusion_optimizer_step(plan: FusionPlan, params, grad, exp_avg, exp_avg_sq, train_state, lr=1e-3, beta1=0.9,
                          beta2=0.999, eps=1e-8, weight_decay=1e-4, grad_scale=1.0, max_norm=0.0,
                          sq_norm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Global-norm clip + AdamW over a HybridFusion master arena (src/train.py:378-382,416-430) with the
    dead query/key slots handled as pure weight decay.  ``train_state`` = int64 {seed, offset, step}
    on the device (step 1-based).  Returns the gradient square-norm scratch (float64[1])."""
    if sq_norm is None:
        sq_norm = torch.zeros(1, dtype=torch.float64, device=params.device)
    N.check(N.lib().msf_fusion_optimizer_step(ctypes.byref(plan.shape), _p(params), _p(grad), _p(exp_avg),
                                              _p(exp_avg_sq), _p(train_state), lr, beta1, beta2, eps,
                                              weight_decay, grad_scale, max_norm, _p(sq_norm), _stream()))
    return sq_norm


def fusion_optimizer_step_packed(plan: FusionPlan, params, grad, exp_avg, exp_avg_sq, train_state, arena_bf16,
                                 lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-4, grad_scale=1.0,
                                 max_norm=0.0, sq_norm: Optional[torch.Tensor] = None,
                                 advance: bool = True, norm_given: bool = False) -> torch.Tensor:
    """``fusion_optimizer_step`` fused with the bf16 re-pack of the compute arena and the advance of
    ``train_state`` (one launch instead of three).  ``norm_given``: ``sq_norm`` is float64[2] and ``sq_norm[1]``
    already holds the gradient square norm (``fusion_train_pass_raw(grad_sq=
    sq_norm[1:])``), see MSF_OPT_NORM_GIVEN."""
    if sq_norm is None:
        sq_norm = torch.zeros(2 if norm_given else 1, dtype=torch.float64, device=params.device)
    if norm_given and sq_norm.numel() < 2:
        raise ValueError("norm_given needs sq_norm with two elements")
    N.check(N.lib().msf_fusion_optimizer_step_packed(
        ctypes.byref(plan.shape), _p(params), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(train_state), lr, beta1,
        beta2, eps, weight_decay, grad_scale, max_norm, _p(sq_norm), _p(arena_bf16),
        int(advance) | (N.MSF_OPT_NORM_GIVEN if norm_given else 0), _stream()))
    return sq_norm


# ---------------------------------------------------------------------------
# LSTM recurrence of SequenceEncoder on the tensor cores (inference path)
# ---------------------------------------------------------------------------
def lstm_pack_weights(weight_ih: torch.Tensor, weight_hh: torch.Tensor, bias_ih: Optional[torch.Tensor],
                      bias_hh: Optional[torch.Tensor]):
    """nn.LSTM layer-0 parameters -> the operand layouts of msf_lstm_forward (include/msf_b200.h):
    gate-interleaved rows (4u+g), bf16, W_hh split into 64-wide k-blocks, W_ih zero-padded to 64 columns."""
    H4, F = weight_ih.shape
    H = H4 // 4
    if H % 64 != 0 or F > 64:
        raise N.MsfError(f"msf_lstm_forward needs hidden % 64 == 0 and input_dim <= 64 (got {H}, {F})")
    dev = weight_hh.device
    inter = lambda w: w.detach().to(torch.float32).view(4, H, -1).permute(1, 0, 2).reshape(4 * H, -1)
    w_hh = inter(weight_hh).view(4 * H, H // 64, 64).permute(1, 0, 2).contiguous().to(torch.bfloat16)
    w_ih = torch.zeros(4 * H, 64, dtype=torch.float32, device=dev)
    w_ih[:, :F] = inter(weight_ih)
    bias = torch.zeros(4 * H, dtype=torch.float32, device=dev)
    if bias_ih is not None:
        bias = bias + bias_ih.detach().to(torch.float32)
    if bias_hh is not None:
        bias = bias + bias_hh.detach().to(torch.float32)
    bias = bias.view(4, H).t().reshape(4 * H).contiguous()
    return w_hh, w_ih.to(torch.bfloat16).contiguous(), bias


def lstm_pack_input(x: torch.Tensor, ones_column: bool = False) -> torch.Tensor:
    """(B, T, F) fp32 windows -> time-major bf16 [T][B][64] with the features zero-padded to 64 columns.
    ``ones_column``: column F carries 1.0 (its weight column is zero), so that the weight-gradient GEMM of
    msf_lstm_backward yields the bias gradient as column F of d W_ih."""
    B, T, F = x.shape
    if x.is_cuda and F <= 64:   # one kernel: pad, transpose to time-major, round to bf16 (msf_lstm_pack_input)
        src = x.detach().to(torch.float32).contiguous()
        out = torch.empty(T, B, 64, dtype=torch.bfloat16, device=x.device)
        with torch.cuda.device(x.device):
            N.check(N.lib().msf_lstm_pack_input(_p(src), B, T, F, int(ones_column and F < 64), _p(out), _stream()))
        return out
    out = torch.zeros(T, B, 64, dtype=torch.bfloat16, device=x.device)   # host tensors (CPU tests of the layout)
    out[:, :, :F] = x.detach().transpose(0, 1)
    if ones_column and F < 64:   # F == 64: no spare column, msf_lstm_backward sums the columns of d a instead
        out[:, :, F] = 1.0
    return out


def gru_as_four_gates(weight_ih, weight_hh, bias_ih, bias_hh):
    """nn.GRU layer parameters (gate order r, z, n; (3H, .)) -> the four-gate form the recurrence kernel takes with
    ``cell_type = 1`` (msf_lstm_seq): gates (r, z, n_x, n_h) with W_ih's n_h rows and W_hh's n_x rows zero, so that the
    input's and the recurrent share of the candidate gate arrive in separate accumulator columns.  The result goes
    through the LSTM packers (lstm_pack_weights / lstm_pack_upper)."""
    H = weight_hh.shape[1]
    w_ih, w_hh = weight_ih.detach().to(torch.float32), weight_hh.detach().to(torch.float32)
    zi, zh = torch.zeros_like(w_ih[:H]), torch.zeros_like(w_hh[:H])
    w_ih4 = torch.cat([w_ih[:2 * H], w_ih[2 * H:], zi], 0)
    w_hh4 = torch.cat([w_hh[:2 * H], zh, w_hh[2 * H:]], 0)
    zb = torch.zeros(H, dtype=torch.float32, device=w_hh.device)
    b_ih = torch.zeros(3 * H, dtype=torch.float32, device=w_hh.device) if bias_ih is None else bias_ih.detach().float()
    b_hh = torch.zeros(3 * H, dtype=torch.float32, device=w_hh.device) if bias_hh is None else bias_hh.detach().float()
    return w_ih4, w_hh4, torch.cat([b_ih[:2 * H], b_ih[2 * H:], zb], 0), torch.cat([b_hh[:2 * H], zb, b_hh[2 * H:]], 0)


def lstm_forward(xs: Sequence[torch.Tensor], packed: Sequence[tuple], hidden: int,
                 lengths: Optional[Sequence[Optional[torch.Tensor]]] = None, cell: str = "lstm") -> List[torch.Tensor]:
    """``h_T`` (fp32, (B, hidden)) of up to 4 single-layer LSTM encoders over packed inputs ``xs``
    (``lstm_pack_input``) with packed weights (``lstm_pack_weights``): one tensor-core launch per time step
    for all of them (msf_lstm_forward); hidden <= 256: one persistent launch over all steps.  ``lengths``: per
    encoder an integer tensor (B,) of valid steps (1..T) or None — the state stops after a window's last valid step,
    as ``pack_padded_sequence`` makes nn.LSTM do (src/encoders.py:140-152)."""
    require_cuda("lstm_forward")
    n = len(xs)
    T, B, _ = xs[0].shape
    dev = xs[0].device
    KBH = hidden // 64
    seqs = (N.LstmSeq * n)()
    keep, outs = [], []
    for i, (x, (w_hh, w_ih, bias)) in enumerate(zip(xs, packed)):
        h_a = torch.zeros(KBH, B, 64, dtype=torch.bfloat16, device=dev)
        h_b = torch.zeros(KBH, B, 64, dtype=torch.bfloat16, device=dev)
        cell = torch.zeros(B, hidden, dtype=torch.float32, device=dev)
        h_out = torch.empty(B, hidden, dtype=torch.float32, device=dev)
        keep += [h_a, h_b, cell]
        outs.append(h_out)
        seqs[i].x_bf16, seqs[i].w_hh, seqs[i].w_ih, seqs[i].bias = _p(x), _p(w_hh), _p(w_ih), _p(bias)
        seqs[i].h_a, seqs[i].h_b, seqs[i].cell, seqs[i].h_out = _p(h_a), _p(h_b), _p(cell), _p(h_out)
        seqs[i].cell_type = 1 if cell == "gru" else 0
        if lengths is not None and lengths[i] is not None:
            ln = lengths[i].to(device=dev, dtype=torch.int32).contiguous()
            if ln.numel() != B or int(ln.min()) < 1 or int(ln.max()) > T:
                raise N.MsfError(f"lengths must be {B} integers in [1, {T}]")
            keep.append(ln)
            seqs[i].lengths = _p(ln)
    N.check(N.lib().msf_lstm_forward(seqs, n, B, T, hidden, _stream()))
    for t in keep:   # the launches are asynchronous: keep the scratch alive on this stream
        t.record_stream(torch.cuda.current_stream(dev))
    return outs


class LstmTape:
    """What msf_lstm_forward keeps in training mode for msf_lstm_backward (one layer of one encoder): the layer's
    input as the weight-gradient GEMM reads it (``inp``: [T][B][64] padded windows with the ones column for the first
    layer, the [T][B][H] (dropped-out) hidden states of the layer below otherwise), h_{t-1} / gate activations / cell
    states of every step, and the transposed weights of the backward GEMMs."""

    def __init__(self, inp, in_cols, features, w_hh_t, w_ih_t, h_all, gates, c_all, h_out, lengths, hidden, layer=0,
                 drop=None, cell="lstm"):
        self.cell = cell
        self.inp, self.in_cols, self.features = inp, in_cols, features
        self.w_hh_t, self.w_ih_t = w_hh_t, w_ih_t
        self.h_all, self.gates, self.c_all, self.h_out = h_all, gates, c_all, h_out
        self.lengths, self.hidden, self.layer, self.drop = lengths, hidden, layer, drop


def _lstm_interleave(w: torch.Tensor, hidden: int) -> torch.Tensor:
    """(4H, K) rows in nn.LSTM's gate-major order -> rows 4u+g (fp32)."""
    return w.detach().to(torch.float32).view(4, hidden, -1).permute(1, 0, 2).reshape(4 * hidden, -1)


def lstm_pack_weights_t(weight: torch.Tensor) -> torch.Tensor:
    """weight_hh (4H, H) or an upper layer's weight_ih (4H, H) -> the operand of the backward GEMMs: [K][4H] bf16
    with ``out[n][4u+g] = weight[g*H+u][n]`` (gate-interleaved columns, like the gate buffers)."""
    return _lstm_interleave(weight, weight.shape[0] // 4).t().contiguous().to(torch.bfloat16)


def lstm_pack_upper(weight_ih, weight_hh, bias_ih, bias_hh):
    """Parameters of a layer above the first -> (w_hh k-blocked [H/64][4H][64], w_ih [4H][H] for the input GEMM, bias
    [4H]), rows gate-interleaved, bf16 / fp32."""
    H = weight_hh.shape[1]
    if H % 64 != 0:
        raise N.MsfError(f"msf_lstm_forward needs hidden % 64 == 0 (got {H})")
    w_hh = _lstm_interleave(weight_hh, H).view(4 * H, H // 64, 64).permute(1, 0, 2).contiguous().to(torch.bfloat16)
    w_ih = _lstm_interleave(weight_ih, H).contiguous().to(torch.bfloat16)
    bias = torch.zeros(4 * H, dtype=torch.float32, device=weight_hh.device)
    for b in (bias_ih, bias_hh):
        if b is not None:
            bias = bias + b.detach().to(torch.float32)
    return w_hh, w_ih, bias.view(4, H).t().reshape(4 * H).contiguous()


def _lstm_lengths(lengths, B, T, dev):
    if lengths is None:
        return None
    ln = lengths.to(device=dev, dtype=torch.int32).contiguous()
    if ln.numel() != B or int(ln.min()) < 1 or int(ln.max()) > T:
        raise N.MsfError(f"lengths must be {B} integers in [1, {T}]")
    return ln


def lstm_forward_stack_group(xs: Sequence[torch.Tensor], layers_list: Sequence[Sequence[tuple]], hidden: int,
                             lengths_list: Optional[Sequence[Optional[torch.Tensor]]] = None,
                             cell: str = "lstm") -> List[torch.Tensor]:
    """``h_n[-1]`` of up to 4 stacked nn.LSTM / nn.GRU encoders of the same depth, hidden size, batch and length in
    inference mode (src/encoders.py:54-72,135-166): ``xs[i]`` (B, T, F_i) fp32, ``layers_list[i]`` = per layer
    ``(weight_ih, weight_hh, bias_ih, bias_hh)``.  Every LAYER is one persistent launch for all encoders
    (msf_lstm_forward, n sequences); a layer above the first gets the input's share of its gate pre-activations from
    ONE tensor-core GEMM per encoder over all steps of the layer below (msf_gemm_bf16 -> z_in).
    ``cell="gru"``: nn.GRU layers (src/encoders.py:66-72), passed through ``gru_as_four_gates``."""
    require_cuda("lstm_forward_stack_group")
    n = len(xs)
    B, T, _ = xs[0].shape
    dev = xs[0].device
    depth = len(layers_list[0])
    if any(len(ls) != depth for ls in layers_list) or any(tuple(x.shape[:2]) != (B, T) for x in xs):
        raise N.MsfError("lstm_forward_stack_group: the encoders of one call share depth, batch and length")
    lns = [_lstm_lengths(None if lengths_list is None else lengths_list[i], B, T, dev) for i in range(n)]
    prev: List[Optional[torch.Tensor]] = [None] * n
    h_outs: List[torch.Tensor] = []
    for l in range(depth):
        last = l == depth - 1
        seqs = (N.LstmSeq * n)()
        keep, h_outs, h_alls = [], [], []
        for i in range(n):
            q, w = seqs[i], layers_list[i][l]
            if cell == "gru":
                w = gru_as_four_gates(*w)
                q.cell_type = 1
            if l == 0:
                w_hh, w_ih, bias = lstm_pack_weights(*w)
                xp = lstm_pack_input(xs[i].to(torch.float32))
                q.x_bf16, q.w_ih = _p(xp), _p(w_ih)
                keep += [xp, w_ih]
            else:
                w_hh, w_ih, bias = lstm_pack_upper(*w)
                z = gemm_bf16(prev[i][1:].view(T * B, hidden), w_ih, out_dtype=torch.bfloat16)
                q.z_in = _p(z)
                keep += [z]
            h_out = torch.empty(B, hidden, dtype=torch.float32, device=dev)
            state = torch.zeros(B, hidden, dtype=torch.float32, device=dev)   # c (LSTM) / h (GRU) in fp32
            q.w_hh, q.bias, q.cell, q.h_out = _p(w_hh), _p(bias), _p(state), _p(h_out)
            keep += [w_hh, bias, state]
            h_outs.append(h_out)
            if lns[i] is not None:
                q.lengths = _p(lns[i])
            if last:
                h_a = torch.zeros(hidden // 64, B, 64, dtype=torch.bfloat16, device=dev)
                h_b = torch.zeros(hidden // 64, B, 64, dtype=torch.bfloat16, device=dev)
                q.h_a, q.h_b = _p(h_a), _p(h_b)
                keep += [h_a, h_b]
            else:
                h_all = torch.empty(T + 1, B, hidden, dtype=torch.bfloat16, device=dev)
                h_all[0].zero_()
                q.h_all = _p(h_all)
                h_alls.append(h_all)
        N.check(N.lib().msf_lstm_forward(seqs, n, B, T, hidden, _stream()))
        prev = h_alls if not last else [None] * n
        del keep
    return h_outs


def lstm_forward_stack(x: torch.Tensor, layers: Sequence[tuple], hidden: int,
                       lengths: Optional[torch.Tensor] = None, cell: str = "lstm") -> torch.Tensor:
    """``lstm_forward_stack_group`` for one encoder."""
    return lstm_forward_stack_group([x], [layers], hidden, None if lengths is None else [lengths], cell)[0]


def lstm_train_forward(xs: Sequence[torch.Tensor], weights: Sequence[tuple], hidden: int,
                       lengths: Optional[Sequence[Optional[torch.Tensor]]] = None, cell: str = "lstm") -> List[LstmTape]:
    """Training-mode forward of up to 4 single-layer LSTM encoders (msf_lstm_forward with the training buffers set):
    ``xs`` are (B, T, F) fp32 windows, ``weights`` per encoder ``(weight_ih, weight_hh, bias_ih, bias_hh)`` as
    nn.LSTM holds them.  One persistent launch; the returned tapes hold ``h_out`` (B, hidden) fp32 and what
    ``lstm_backward`` needs: h_{t-1} (bf16), the gate activations (bf16) and the cell states (fp32) of every step."""
    require_cuda("lstm_train_forward")
    n = len(xs)
    B, T, _ = xs[0].shape
    dev = xs[0].device
    seqs = (N.LstmSeq * n)()
    tapes, keep = [], []
    for i, (x, (w_ih, w_hh, b_ih, b_hh)) in enumerate(zip(xs, weights)):
        F = x.shape[2]
        xp = lstm_pack_input(x.to(torch.float32), ones_column=True)
        if cell == "gru":   # nn.GRU: four-gate form (r, z, n_x, n_h), fp32 hidden state in `cell`, no cell-state tape
            w_ih, w_hh, b_ih, b_hh = gru_as_four_gates(w_ih, w_hh, b_ih, b_hh)
            seqs[i].cell_type = 1
            state = torch.zeros(B * hidden, dtype=torch.float32, device=dev)
            seqs[i].cell = _p(state)
            keep.append(state)
        packed = lstm_pack_weights(w_ih, w_hh, b_ih, b_hh)
        h_all = torch.empty(T + 1, B, hidden, dtype=torch.bfloat16, device=dev)
        h_all[0].zero_()
        gates = torch.empty(T, B, 4 * hidden, dtype=torch.bfloat16, device=dev)
        c_all = None if cell == "gru" else torch.empty(T, B * hidden, dtype=torch.float32, device=dev)
        h_out = torch.empty(B, hidden, dtype=torch.float32, device=dev)
        ln = _lstm_lengths(None if lengths is None else lengths[i], B, T, dev)
        if ln is not None:
            seqs[i].lengths = _p(ln)
        seqs[i].x_bf16, seqs[i].w_hh, seqs[i].w_ih, seqs[i].bias = _p(xp), _p(packed[0]), _p(packed[1]), _p(packed[2])
        seqs[i].h_all, seqs[i].gates, seqs[i].c_all, seqs[i].h_out = _p(h_all), _p(gates), _p(c_all), _p(h_out)
        keep.append(packed)
        tapes.append(LstmTape(xp, 0, F, lstm_pack_weights_t(w_hh), None, h_all, gates, c_all, h_out, ln, hidden, cell=cell))
    N.check(N.lib().msf_lstm_forward(seqs, n, B, T, hidden, _stream()))
    return tapes


def lstm_train_forward_stack_group(xs: Sequence[torch.Tensor], layers_list: Sequence[Sequence[tuple]], hidden: int,
                                   lengths_list: Optional[Sequence[Optional[torch.Tensor]]] = None,
                                   dropout_p: float = 0.0, seeds: Optional[Sequence[int]] = None,
                                   cell: str = "lstm") -> List[List[LstmTape]]:
    """Training-mode forward of up to 4 stacked nn.LSTM encoders of the same depth / hidden size / batch / length:
    per encoder one tape per layer, bottom first (``tapes[i][-1].h_out`` is ``h_n[-1]`` of encoder i); every layer is
    ONE persistent launch for all encoders.  Between the layers nn.LSTM's dropout (p = ``dropout_p``) is applied with
    the library's Philox multipliers (msf_lstm_dropout: site 4, sub = layer, keyed by ``seeds[i]``).  A layer above the
    first reads the input's share of its pre-activations from one GEMM over all steps, written straight into its gate
    buffer (z_in == gates: read, then overwritten in place by the activations)."""
    require_cuda("lstm_train_forward_stack_group")
    n = len(xs)
    B, T, _ = xs[0].shape
    dev = xs[0].device
    depth = len(layers_list[0])
    if any(len(ls) != depth for ls in layers_list) or any(tuple(x.shape[:2]) != (B, T) for x in xs):
        raise N.MsfError("lstm_train_forward_stack_group: the encoders of one call share depth, batch and length")
    seeds = list(seeds) if seeds is not None else [0] * n
    lns = [_lstm_lengths(None if lengths_list is None else lengths_list[i], B, T, dev) for i in range(n)]
    tapes = [[tp] for tp in lstm_train_forward(xs, [ls[0] for ls in layers_list], hidden, lns, cell)]
    for l in range(1, depth):
        seqs = (N.LstmSeq * n)()
        keep, new = [], []
        for i in range(n):
            w = layers_list[i][l]
            below = tapes[i][-1].h_all[1:].view(T * B, hidden)
            drop = None
            if dropout_p > 0.0:
                inp = torch.empty_like(below)
                N.check(N.lib().msf_lstm_dropout(_p(below), _p(inp), T * B, hidden, float(dropout_p),
                                                 seeds[i] & (2**64 - 1), 0, l, _stream()))
                drop = (float(dropout_p), seeds[i], 0)
            else:
                inp = below
            q = seqs[i]
            if cell == "gru":
                w = gru_as_four_gates(*w)
                q.cell_type = 1
                state = torch.zeros(B * hidden, dtype=torch.float32, device=dev)
                q.cell = _p(state)
                keep.append(state)
            w_hh, w_ih, bias = lstm_pack_upper(*w)
            gates = gemm_bf16(inp, w_ih, out_dtype=torch.bfloat16).view(T, B, 4 * hidden)
            h_all = torch.empty(T + 1, B, hidden, dtype=torch.bfloat16, device=dev)
            h_all[0].zero_()
            c_all = None if cell == "gru" else torch.empty(T, B * hidden, dtype=torch.float32, device=dev)
            h_out = torch.empty(B, hidden, dtype=torch.float32, device=dev)
            q.w_hh, q.bias, q.z_in, q.gates = _p(w_hh), _p(bias), _p(gates), _p(gates)
            q.h_all, q.c_all, q.h_out = _p(h_all), _p(c_all), _p(h_out)
            if lns[i] is not None:
                q.lengths = _p(lns[i])
            keep += [w_hh, bias]
            new.append(LstmTape(inp, hidden, hidden, lstm_pack_weights_t(w[1]), lstm_pack_weights_t(w[0]), h_all, gates,
                                c_all, h_out, lns[i], hidden, layer=l, drop=drop, cell=cell))
        N.check(N.lib().msf_lstm_forward(seqs, n, B, T, hidden, _stream()))
        for i in range(n):
            tapes[i].append(new[i])
        del keep
    return tapes


def lstm_train_forward_stack(x: torch.Tensor, layers: Sequence[tuple], hidden: int,
                             lengths: Optional[torch.Tensor] = None, dropout_p: float = 0.0, seed: int = 0,
                             offset: int = 0) -> List[LstmTape]:
    """``lstm_train_forward_stack_group`` for one encoder (``offset`` is kept for callers of the first version: 0)."""
    return lstm_train_forward_stack_group([x], [layers], hidden, None if lengths is None else [lengths], dropout_p,
                                          [seed], "lstm")[0]


def lstm_backward(tapes: Sequence[LstmTape], d_h_out: Sequence[Optional[torch.Tensor]],
                  d_h_all: Optional[Sequence[Optional[torch.Tensor]]] = None, release: bool = True):
    """Gradients of the recurrences recorded by ``lstm_train_forward`` (one layer each, up to 4 per call): per tape
    ``(d weight_ih, d weight_hh, d bias)`` in nn.LSTM's layout (``d bias`` is the gradient of bias_ih and of bias_hh
    alike).  ``d_h_out[i]``: gradient of the tape's ``h_out`` or None; ``d_h_all[i]``: [T][B][H] bf16 gradient of
    every step's hidden state (from the layer above) or None.  Overwrites the tapes' gate buffers with the gradients
    of the gate pre-activations (msf_lstm_backward): a tape can be walked backwards once."""
    require_cuda("lstm_backward")
    n = len(tapes)
    T1, B, hidden = tapes[0].h_all.shape
    T = T1 - 1
    dev = tapes[0].h_all.device
    nbytes = ctypes.c_size_t(0)
    N.check(N.lib().msf_lstm_backward_scratch_bytes(B, T, hidden, ctypes.byref(nbytes)))
    seqs = (N.LstmSeq * n)()
    keep, outs = [], []
    for i, tp in enumerate(tapes):
        if tp.gates is None:
            raise N.MsfError("lstm_backward: this tape was already walked backwards")
        dc = torch.zeros(B * hidden, dtype=torch.float32, device=dev)
        partial = torch.empty(nbytes.value // 4, dtype=torch.float32, device=dev)
        d_w_ih = torch.empty(4 * hidden, tp.features, dtype=torch.float32, device=dev)
        d_w_hh = torch.empty(4 * hidden, hidden, dtype=torch.float32, device=dev)
        d_b = torch.empty(4 * hidden, dtype=torch.float32, device=dev)
        keep += [dc, partial]
        outs.append((d_w_ih, d_w_hh, d_b))
        q = seqs[i]
        q.x_bf16, q.h_all, q.gates, q.c_all, q.w_hh_t = _p(tp.inp), _p(tp.h_all), _p(tp.gates), _p(tp.c_all), _p(tp.w_hh_t)
        q.dc, q.partial = _p(dc), _p(partial)
        q.d_w_ih, q.d_w_hh, q.d_bias, q.features, q.in_cols = _p(d_w_ih), _p(d_w_hh), _p(d_b), tp.features, tp.in_cols
        q.cell_type = 1 if tp.cell == "gru" else 0
        if d_h_out[i] is not None:
            dh = d_h_out[i].to(device=dev, dtype=torch.float32).contiguous()
            keep.append(dh)
            q.d_h_out = _p(dh)
        if d_h_all is not None and d_h_all[i] is not None:
            da = d_h_all[i]
            assert da.dtype == torch.bfloat16 and da.is_contiguous() and da.numel() == T * B * hidden
            keep.append(da)
            q.d_h_all = _p(da)
        if tp.lengths is not None:
            q.lengths = _p(tp.lengths)
    N.check(N.lib().msf_lstm_backward(seqs, n, B, T, hidden, _stream()))
    if release:
        for tp in tapes:
            tp.gates = None
    return outs


def lstm_backward_stack_group(tapes_list: Sequence[Sequence[LstmTape]], d_h_outs: Sequence[torch.Tensor]):
    """Backward pass through the tapes of ``lstm_train_forward_stack_group`` (top layer first, every layer ONE
    persistent launch + the weight-gradient GEMMs for all encoders): per encoder and layer
    ``(d weight_ih, d weight_hh, d bias)``, bottom first.  Between two layers the gradient of the lower layer's hidden
    states is one GEMM over all steps (d a of the upper layer times its W_ih) through the same dropout mask."""
    n, depth = len(tapes_list), len(tapes_list[0])
    grads: List[List[Optional[tuple]]] = [[None] * depth for _ in range(n)]
    d_all: List[Optional[torch.Tensor]] = [None] * n
    for l in reversed(range(depth)):
        tps = [tapes_list[i][l] for i in range(n)]
        T1, B, hidden = tps[0].h_all.shape
        out = lstm_backward(tps, [d_h_outs[i] if l == depth - 1 else None for i in range(n)], d_all, release=False)
        for i in range(n):
            grads[i][l] = out[i]
            tp = tps[i]
            if l > 0:
                d_all[i] = gemm_bf16(tp.gates.view((T1 - 1) * B, 4 * hidden), tp.w_ih_t, out_dtype=torch.bfloat16)
                if tp.drop is not None:
                    p, seed, offset = tp.drop
                    N.check(N.lib().msf_lstm_dropout(_p(d_all[i]), _p(d_all[i]), (T1 - 1) * B, hidden, p,
                                                     seed & (2**64 - 1), offset & (2**64 - 1), l, _stream()))
            tp.gates = tp.c_all = tp.h_all = tp.inp = None
    return grads


def lstm_backward_stack(tapes: Sequence[LstmTape], d_h_out: torch.Tensor):
    """``lstm_backward_stack_group`` for one encoder."""
    return lstm_backward_stack_group([tapes], [d_h_out])[0]


class LstmLastHiddenGroup(torch.autograd.Function):
    """``h_n[-1]`` of up to 4 (stacked) nn.LSTM encoders of the same depth / hidden size / batch / length in one go,
    with gradients for their parameters: every layer of the forward and of the backward pass is ONE persistent launch
    for all encoders.  ``tensors`` = the n input sequences, then the parameters (weight_ih, weight_hh, bias_ih,
    bias_hh per layer, bottom first) of encoder 0, encoder 1, ...  Returns n tensors (B, hidden)."""

    @staticmethod
    def forward(ctx, n, depth, lengths_list, dropout_p, seeds, cell, *tensors):
        xs, params = tensors[:n], tensors[n:]
        ctx.cell = cell
        layers_list = [[tuple(params[(i * depth + l) * 4:(i * depth + l) * 4 + 4]) for l in range(depth)] for i in range(n)]
        hidden = layers_list[0][0][1].shape[1]
        tapes = lstm_train_forward_stack_group(xs, layers_list, hidden, lengths_list, dropout_p, seeds, cell)
        ctx.tapes, ctx.n = tapes, n
        ctx.has = [p is not None for p in params]
        return tuple(tp[-1].h_out for tp in tapes)

    @staticmethod
    def backward(ctx, *d_hs):
        outs = [tp[-1].h_out for tp in ctx.tapes]
        d_hs = [torch.zeros_like(o) if d is None else d for d, o in zip(d_hs, outs)]
        grads = lstm_backward_stack_group(ctx.tapes, d_hs)
        ctx.tapes = None
        flat = []
        for per_encoder in grads:
            for (d_w_ih, d_w_hh, d_b) in per_encoder:
                if ctx.cell == "gru":   # four-gate rows (r | z | n_x | n_h) back to nn.GRU's (r | z | n)
                    H = d_w_hh.shape[1]
                    flat += [torch.cat([d_w_ih[:2 * H], d_w_ih[2 * H:3 * H]], 0), torch.cat([d_w_hh[:2 * H], d_w_hh[3 * H:]], 0),
                             torch.cat([d_b[:2 * H], d_b[2 * H:3 * H]], 0), torch.cat([d_b[:2 * H], d_b[3 * H:]], 0)]
                else:
                    flat += [d_w_ih, d_w_hh, d_b, d_b.clone()]
        return (None, None, None, None, None, None, *([None] * ctx.n), *[g if has else None for g, has in zip(flat, ctx.has)])


class LstmLastHiddenF32(torch.autograd.Function):
    """``h_n[-1]`` of a (stacked) nn.LSTM or nn.GRU (``cell`` = "lstm" / "gru") on the fp32 parity kernels
    (msf_lstm_f32_* / msf_gru_f32_*, one call per layer), with gradients for every parameter AND the input sequence.
    ``params``: (weight_ih, weight_hh, bias_ih, bias_hh) per layer, flattened, bottom first.  Training-mode
    inter-layer dropout (``dropout_p`` > 0): the library's Philox multipliers of site 4 (msf_dropout_mask), keyed by
    ``seed``."""

    @staticmethod
    def forward(ctx, x, lengths, dropout_p, seed, cell, *params):
        require_cuda("LstmLastHiddenF32")
        B, T, _ = x.shape
        dev = x.device
        n_layers = len(params) // 4
        gru = cell == "gru"
        ln = _lstm_lengths(lengths, B, T, dev)
        inp = x.detach().to(torch.float32).transpose(0, 1).contiguous()   # time-major [T][B][F]
        layers = []
        for l in range(n_layers):
            w_ih, w_hh, b_ih, b_hh = (None if p is None else p.detach().to(torch.float32).contiguous()
                                      for p in params[4 * l:4 * l + 4])
            H = w_hh.shape[1]
            mask = None
            if l > 0 and dropout_p > 0.0:
                mask = dropout_mask(seed, 0, 4, l, T * B, H, dropout_p, dev).view(T, B, H)
                inp = inp * mask
            h_seq = torch.empty(T + 1, B, H, dtype=torch.float32, device=dev)
            h_seq[0].zero_()
            gates = torch.empty(T, B, 4 * H, dtype=torch.float32, device=dev)
            scratch = torch.empty(B * 4 * H, dtype=torch.float32, device=dev)
            if gru:
                c_seq = None
                N.check(N.lib().msf_gru_f32_forward(_p(inp), inp.shape[2], _p(w_ih), _p(w_hh), _p(b_ih), _p(b_hh), _p(ln), B, T,
                                                    H, _p(h_seq), _p(gates), _p(scratch), _stream()))
            else:
                c_seq = torch.empty(T + 1, B, H, dtype=torch.float32, device=dev)
                c_seq[0].zero_()
                N.check(N.lib().msf_lstm_f32_forward(_p(inp), inp.shape[2], _p(w_ih), _p(w_hh), _p(b_ih), _p(b_hh), _p(ln), B,
                                                     T, H, _p(h_seq), _p(c_seq), _p(gates), _p(scratch), _stream()))
            layers.append((inp, w_ih, w_hh, h_seq, c_seq, gates, mask))
            inp = h_seq[1:]
        ctx.layers, ctx.ln, ctx.gru = layers, ln, gru
        ctx.has = [p is not None for p in params]
        ctx.want_dx = x.requires_grad
        return layers[-1][3][T].clone()   # finished windows carry their state: slice T holds every window's last state

    @staticmethod
    def backward(ctx, d_h):
        layers, ln, gru = ctx.layers, ctx.ln, ctx.gru
        T1, B, H = layers[0][3].shape
        T = T1 - 1
        dev = d_h.device
        d_last = d_h.to(torch.float32).contiguous()
        d_seq = None
        flat = [None] * (4 * len(layers))
        d_x = None
        for l in reversed(range(len(layers))):
            inp, w_ih, w_hh, h_seq, c_seq, gates, mask = layers[l]
            F = inp.shape[2]
            need_dx = l > 0 or ctx.want_dx
            dx = torch.empty(T, B, F, dtype=torch.float32, device=dev) if need_dx else None
            d_w_ih, d_w_hh = torch.empty_like(w_ih), torch.empty_like(w_hh)
            scratch = torch.empty(3 * B * H, dtype=torch.float32, device=dev)
            top = d_last if l == len(layers) - 1 else None
            if gru:
                d_b_ih = torch.empty(3 * H, dtype=torch.float32, device=dev)
                d_b_hh = torch.empty(3 * H, dtype=torch.float32, device=dev)
                dzh = torch.empty(T, B, 3 * H, dtype=torch.float32, device=dev)
                N.check(N.lib().msf_gru_f32_backward(_p(inp), F, _p(w_ih), _p(w_hh), _p(ln), B, T, H, _p(h_seq), _p(gates),
                                                     _p(dzh), _p(top), _p(d_seq), _p(scratch), _p(dx), _p(d_w_ih), _p(d_w_hh),
                                                     _p(d_b_ih), _p(d_b_hh), _stream()))
                flat[4 * l:4 * l + 4] = [d_w_ih, d_w_hh, d_b_ih, d_b_hh]
            else:
                d_b = torch.empty(4 * H, dtype=torch.float32, device=dev)
                N.check(N.lib().msf_lstm_f32_backward(_p(inp), F, _p(w_ih), _p(w_hh), _p(ln), B, T, H, _p(h_seq), _p(c_seq),
                                                      _p(gates), _p(top), _p(d_seq), _p(scratch), _p(dx), _p(d_w_ih),
                                                      _p(d_w_hh), _p(d_b), _stream()))
                flat[4 * l:4 * l + 4] = [d_w_ih, d_w_hh, d_b, d_b.clone()]
            if l > 0:
                d_seq = dx if mask is None else dx * mask
            else:
                d_x = dx
        ctx.layers = None
        gx = None if d_x is None else d_x.transpose(0, 1).contiguous()
        return (gx, None, None, None, None, *[g if has else None for g, has in zip(flat, ctx.has)])


class LstmLastHidden:
    """``h_n[-1]`` of ONE (stacked) nn.LSTM / nn.GRU with gradients for its parameters (none for ``x``: the encoders'
    inputs are data): ``LstmLastHiddenGroup`` with one encoder.  ``params`` are (weight_ih, weight_hh, bias_ih, bias_hh)
    of every layer, flattened, bottom layer first."""

    @staticmethod
    def apply(x, lengths, dropout_p, seed, *params, cell: str = "lstm"):
        return LstmLastHiddenGroup.apply(1, len(params) // 4, None if lengths is None else [lengths], dropout_p, [seed],
                                         cell, x, *params)[0]
