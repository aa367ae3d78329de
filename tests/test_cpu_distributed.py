"""World-size-2 gloo run (CPU) of the data-parallel host logic: contiguous equal
batch shards, each rank's gradient pre-scaled by 1/(B_global) and summed by all-reduce equals the
single-process global-batch gradient; shard-local ECE bin statistics merged by one integer
all-reduce equal the unsharded statistics bit for bit.  The arithmetic here is the CPU oracle —
the test covers the sharding / reduction contract the GPU engine implements with NCCL."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_pkg
from oracle import ece_oracle, fusion_oracle

DIMS = {"a": 8, "b": 8, "c": 8}
HIDDEN, HEADS, CLASSES, BATCH = 16, 2, 5, 24


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    import sys
    from conftest import dropin_src
    if dropin_src() not in sys.path:
        sys.path.insert(0, dropin_src())
    fusion = importlib.import_module("fusion")
    torch.manual_seed(3)
    model = fusion.HybridFusion(DIMS, hidden_dim=HIDDEN, num_classes=CLASSES, num_heads=HEADS, dropout=0.0)
    gen = torch.Generator().manual_seed(4)
    feats = {m: torch.randn(BATCH, d, generator=gen) for m, d in DIMS.items()}
    mask = (torch.rand(BATCH, len(DIMS), generator=gen) < 0.8).float()
    labels = torch.randint(0, CLASSES, (BATCH,), generator=gen)
    return model, feats, mask, labels


def _grads(sd, names, feats, mask, labels, scale_batch):
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits, _ = fusion_oracle.hybrid_fusion_forward(sd, names, HEADS, feats, mask)
    # mean over the GLOBAL batch: shard loss sum / B_global (engine: grad_scale = 1 / (B * world))
    loss = fusion_oracle.cross_entropy_label_smoothing(logits, labels, 0.05) * (labels.numel() / scale_batch)
    loss.backward()
    return torch.cat([v.grad.reshape(-1) for v in sd.values()]), logits.detach()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        engine = importlib.import_module(load_pkg().__name__ + ".engine")
        model, feats, mask, labels = _problem()
        names = list(DIMS)
        sl = engine.shard_batch(BATCH, rank, world)
        g, logits = _grads(model.state_dict(), names, {k: v[sl] for k, v in feats.items()}, mask[sl], labels[sl], BATCH)
        dist.all_reduce(g)  # SUM of pre-scaled shard gradients
        probs = torch.softmax(logits, 1)
        conf, pred = probs.max(1)
        edges = ece_oracle.linspace_f32(15).astype(np.float64)
        cnt, cor, csum = ece_oracle.bin_masks(conf.numpy(), pred.numpy(), labels[sl].numpy(), edges)
        stats = torch.from_numpy(np.stack([cnt, cor]).astype(np.int64))
        dist.all_reduce(stats)
        if rank == 0:
            g_full, logits_full = _grads(model.state_dict(), names, feats, mask, labels, BATCH)
            conf_f, pred_f = torch.softmax(logits_full, 1).max(1)
            cnt_f, cor_f, _ = ece_oracle.bin_masks(conf_f.numpy(), pred_f.numpy(), labels.numpy(), edges)
            out["grad_err"] = float((g - g_full).abs().max())
            out["grad_scale"] = float(g_full.abs().max())
            out["bins_equal"] = bool(np.array_equal(stats[0].numpy(), cnt_f) and np.array_equal(stats[1].numpy(), cor_f))
            out["total"] = int(stats[0].sum())
    finally:
        dist.destroy_process_group()


def test_two_rank_gradients_and_bins_match_single_process():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert out["grad_scale"] > 0 and out["grad_err"] <= 1e-6, dict(out)
        assert out["bins_equal"] and out["total"] == BATCH


def test_shard_batch_contract():
    engine = importlib.import_module(load_pkg().__name__ + ".engine")
    assert [engine.shard_batch(32, r, 4) for r in range(4)] == [slice(0, 8), slice(8, 16), slice(16, 24), slice(24, 32)]
    with pytest.raises(ValueError, match="not divisible"):
        engine.shard_batch(10, 0, 4)
