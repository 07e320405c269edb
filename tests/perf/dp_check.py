"""Multi-GPU check of the fused data-parallel optimiser (b200_dp_adam_step: reduce-scatter + Adam + all-gather over
NVLink peer memory) against the baseline formulation (NCCL all-reduce of [gradients || losses] + local Adam):
same shards, same index batches, same initial state -> parameters, Adam moments and loss vectors must agree to fp32
summation order.  Run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/perf/dp_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
from b200 import _native as N, atlas as A, synth          # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    H, W, T, B = 96, 160, 16, 4000
    data = synth.throughput_set(H, W, T, seed=0)
    t0, t1 = A.frame_range(rank, world, T)
    vid = A.DeviceVideo.from_reference_layout(data, dev, t0, t1)
    prec = N.PREC_TC if N.lib().b200_device_supports_tc() else N.PREC_FP32
    out = {"world": world}
    trainers = {}
    for name, fused in (("nccl", False), ("fused", True)):
        tr = A.AtlasTrainer(vid, {"samples_batch": B}, precision=prec, device=dev, process_group=dist.group.WORLD,
                            fused_dp=fused)
        torch.manual_seed(0)
        tr.init_like_reference()
        dist.broadcast(tr.params, 0)
        trainers[name] = tr
    assert trainers["fused"]._dp is not None and trainers["nccl"]._dp is None
    # ---- 1. the exchange itself, on identical synthetic partial buffers: tight
    gg = torch.Generator().manual_seed(100 + rank)
    synth_grads = (torch.randn(trainers["nccl"].n_params + N.LOSS_FLOATS, generator=gg) * 1e-2).to(dev)
    for step in range(3):
        for name, tr in trainers.items():
            tr.grad_loss.copy_(synth_grads * (step + 1))
            if tr._dp is not None:
                tr.dp_adam()
            else:
                tr.all_reduce(); tr.adam()
    torch.cuda.synchronize()
    out["exchange_param_diff"] = float((trainers["fused"].params - trainers["nccl"].params).abs().max())
    out["exchange_loss_diff"] = float((trainers["fused"].losses - trainers["nccl"].losses).abs().max())
    trainers["fused"].gather_moments()
    out["exchange_moment_rel"] = float((trainers["fused"].exp_avg - trainers["nccl"].exp_avg).abs().max() /
                                       trainers["nccl"].exp_avg.abs().max())
    # ---- 2. real iterations (gradients come from fp32 atomics, so two runs differ by summation order): loose
    g = torch.Generator().manual_seed(5)
    worst = {"params": 0.0, "losses": 0.0, "moments": 0.0}
    for it, use_graph in ((0, False), (1, False), (2, True), (3, True), (6000, True), (6001, True), (4, True)):
        inds = torch.randint(H * W * T, (B, 1), generator=g)
        res = {}
        for name, tr in trainers.items():
            res[name] = tr.step_host(inds, it, use_graph=use_graph)
        worst["losses"] = max(worst["losses"], float(abs(res["fused"][:6] - res["nccl"][:6]).max() / abs(res["nccl"][0])))
        assert res["fused"][6] == res["nccl"][6] and res["fused"][7] == res["nccl"][7]
        worst["params"] = max(worst["params"], float((trainers["fused"].params - trainers["nccl"].params).abs().max()))
    # identical parameters on every rank
    ref = trainers["fused"].params.clone()
    dist.broadcast(ref, 0)
    out["rank_param_diff"] = float((ref - trainers["fused"].params).abs().max())
    trainers["fused"].gather_moments()
    worst["moments"] = float((trainers["fused"].exp_avg - trainers["nccl"].exp_avg).abs().max() /
                             trainers["nccl"].exp_avg.abs().max())
    out.update(worst)
    out["steps"] = [int(trainers[n].step_count) for n in ("nccl", "fused")]
    ok = (out["rank_param_diff"] == 0.0 and out["exchange_param_diff"] <= 1e-6 and out["exchange_loss_diff"] <= 1e-5 and
          out["exchange_moment_rel"] <= 1e-5 and worst["params"] <= 1.5e-3 and worst["losses"] <= 1e-4 and
          out["steps"] == [10, 10])
    out["ok"] = bool(ok)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out["all_ranks_ok"] = bool(t.item() == 1.0)
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize(); dist.barrier()
    sys.stdout.flush()
    os._exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
