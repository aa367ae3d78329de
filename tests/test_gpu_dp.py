"""Two-GPU data parallelism of the fused train step (needs >= 2 CUDA devices, skipped otherwise): the
peer-memory path (msf_dp_optimizer_step: reduce-scatter + norm and all-gather + clip + AdamW over NVLink)
against the NCCL all-reduce path and against one process stepping on the global batch."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_step_matches_nccl_and_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "dp_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["replicas_identical"], res       # every rank applied the same reduced gradient and norm
    assert res["moved"] > 1e-4, res             # the optimizer did step
    assert res["p2p_vs_nccl"] <= 1e-6, res      # same sums (two ranks: a + b is order-free), same AdamW
    assert res["p2p_vs_single"] <= 2e-5, res    # shard sums vs one global sum differ only in rounding
    # bf16 engine: exchange + AdamW + bf16 re-pack + state advance in one launch (msf_dp_optimizer_step_packed)
    assert res["bf16_replicas_identical"], res
    assert res["bf16_pack_consistent"], res     # the compute arena is exactly the pack of the updated parameters
    assert res["bf16_state"][2] >= 3 and res["bf16_state"][1] == res["bf16_state"][2] - 1, res
    assert res["bf16_vs_fp32"] <= 1.3e-2, res   # Adam turns sign flips of ~0 gradients into 2*lr per step
    # sharded optimizer (msf_dpz_optimizer_step_packed): owner-computes, bf16 weights pushed to every rank
    assert res["zshard_replicas_identical"], res          # masters agree once gathered from their owners
    assert res["zshard_bf16_identical"], res              # every rank computes with the same bf16 weights
    assert res["zshard_pack_consistent"], res             # ... which are exactly the pack of the masters
    assert res["zshard_vs_p2p"] <= 2e-5, res              # same sums in the same rank order, same AdamW: the two
                                                          # exchanges differ only where fp32 atomics order the bias sums
    assert res["zshard_state"][2] == res["bf16_state"][2], res
    assert res["zshard_module_param_is_arena_view"], res
