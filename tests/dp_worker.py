"""Worker of tests/test_gpu_dp.py (launched with torch.distributed.run, one process per GPU): steps the fused
train step on batch shards with the peer-memory exchange and with NCCL, and rank 0 also steps one process on
the global batch.  Prints one JSON line (rank 0)."""
import faulthandler
import importlib
import json
import os
import sys

faulthandler.dump_traceback_later(100, exit=True)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from conftest import load_pkg  # noqa: E402
from helpers import PAMAP2, seeded_case  # noqa: E402

GLOBAL_B, STEPS = 512, 3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    engine = importlib.import_module(load_pkg().__name__ + ".engine")
    sl = engine.shard_batch(GLOBAL_B, rank, world)
    res = {}
    for comm in ("p2p", "nccl"):
        model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, GLOBAL_B, seed=31, device=dev)
        eng = engine.FusionEngine(model, GLOBAL_B // world, precision="fp32", seed=9, use_graph=True, comm=comm)
        assert eng.comm == comm, eng.comm
        eng.p = 0.0
        shard = ({k: v[sl] for k, v in feats.items()}, mask[sl], labels[sl])
        for _ in range(STEPS):
            eng.train_step(*shard)
        torch.cuda.synchronize()
        res[comm] = eng.arena.clone()
    mine = res["p2p"]
    other = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    out = {"replicas_identical": all(bool(torch.equal(o, mine)) for o in other),
           "p2p_vs_nccl": float((res["p2p"] - res["nccl"]).abs().max())}
    # tensor-core path: the exchange is fused with AdamW + bf16 re-pack + state advance (one launch)
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, GLOBAL_B, seed=31, device=dev)
    eng = engine.FusionEngine(model, GLOBAL_B // world, precision="bf16", seed=9, use_graph=True, comm="p2p")
    eng.p = 0.0
    shard = ({k: v[sl] for k, v in feats.items()}, mask[sl], labels[sl])
    for _ in range(STEPS):
        eng.train_step(*shard)
    torch.cuda.synchronize()
    mine16 = eng.arena.clone()
    other = [torch.empty_like(mine16) for _ in range(world)]
    dist.all_gather(other, mine16)
    out["bf16_replicas_identical"] = all(bool(torch.equal(o, mine16)) for o in other)
    out["bf16_pack_consistent"] = bool(torch.equal(eng.plan.pack_bf16(eng.arena).view(torch.int16),
                                                   eng.arena_bf16.view(torch.int16)))
    out["bf16_state"] = eng.state.tolist()
    out["bf16_vs_fp32"] = float((mine16 - res["p2p"]).abs().max())
    # sharded optimizer (msf_dpz_optimizer_step_packed): every rank reduces / updates only the tiles it owns and
    # pushes their bf16 copies everywhere.  Same arithmetic as the replicated p2p step: after gather_parameters()
    # the masters equal the p2p engine's bit for bit, and every rank's compute arena is the pack of those masters.
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, GLOBAL_B, seed=31, device=dev)
    engz = engine.FusionEngine(model, GLOBAL_B // world, precision="bf16", seed=9, use_graph=True, comm="zshard")
    assert engz.comm == "zshard", engz.comm
    engz.p = 0.0
    for _ in range(STEPS):
        engz.train_step(*shard)
    torch.cuda.synchronize()
    dist.barrier()
    bf16_local = engz.arena_bf16.clone()
    engz.gather_parameters()
    torch.cuda.synchronize()
    mz = engz.arena.clone()
    other = [torch.empty_like(mz) for _ in range(world)]
    dist.all_gather(other, mz)
    out["zshard_replicas_identical"] = all(bool(torch.equal(o, mz)) for o in other)
    other16 = [torch.empty_like(bf16_local) for _ in range(world)]
    dist.all_gather(other16, bf16_local)
    out["zshard_bf16_identical"] = all(bool(torch.equal(o.view(torch.int16), bf16_local.view(torch.int16))) for o in other16)
    out["zshard_pack_consistent"] = bool(torch.equal(engz.plan.pack_bf16(engz.arena).view(torch.int16),
                                                     bf16_local.view(torch.int16)))
    out["zshard_vs_p2p"] = float((mz - mine16).abs().max())
    out["zshard_state"] = engz.state.tolist()
    out["zshard_module_param_is_arena_view"] = bool(
        next(iter(engz.model.parameters())).data_ptr() == engz.arena.data_ptr())
    if rank == 0:
        model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, GLOBAL_B, seed=31, device=dev)
        start = torch.cat([p.detach().flatten() for p in model.parameters()]).clone()
        single = engine.FusionEngine(model, GLOBAL_B, precision="fp32", seed=9, use_graph=False, comm="nccl")
        single.world, single.comm, single.p = 1, "none", 0.0   # one process, the whole batch, no exchange
        for _ in range(STEPS):
            single.train_step(feats, mask, labels)
        torch.cuda.synchronize()
        out["p2p_vs_single"] = float((res["p2p"] - single.arena).abs().max())
        out["moved"] = float((res["p2p"] - start).abs().max())
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)  # communicator teardown with live CUDA graphs can block (see bench.py)


if __name__ == "__main__":
    main()
