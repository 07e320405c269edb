#!/usr/bin/env python
"""Throughput of the stage-1 neural-atlas loop (BASELINE.json metric: atlas iters/sec, 80 frames
768x432, 10 000 points per iteration) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision tc|fp32]

One "step" = one loop trip of src/stage1_neural_atlas.py:151-231 (sampling, 7 mapping + 3 atlas
evaluations, 4 losses, backward, Adam).  Prints ONE JSON line on rank 0.  Keys beyond the driver's
contract: `roofline` (dominant kernel, timed live with CUDA events around its launch site inside the
replayed step), `cpu_baseline` (the oracle on this box's host cores, bounded sample),
`e2e` (same metric through AtlasTrainer.step_host: pinned H2D of the index batch + D2H of the loss
vector + sync every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))

H, W, T, BATCH = 432, 768, 80, 10000          # BASELINE.json configs[1]
MAC_MAP, MAC_ATLAS = 263424, 414584           # MACs per row


def algorithmic_flop(with_global: bool, n_f: float, n_b: float) -> float:
    """6 x (R_map x 263 424 + R_atlas x 414 584), SURVEY.md §8(d)."""
    r_map = (7 if with_global else 5) * BATCH + n_f + n_b
    return 6.0 * (r_map * MAC_MAP + 3 * BATCH * MAC_ATLAS)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.15)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm)}


CPU_THREADS = 32      # the oracle gets SLOWER beyond this on the B200 hosts (128 threads: 0.03-0.08 it/s, measured)


def run_oracle_steps(data, steps, warmup, threads=None, budget_s=None, fraction=1.0):
    """The CPU restatement of the reference loop on the host cores (oracle = test/baseline code).
    Returns (seconds, steps actually timed).  `fraction` < 1 times a sub-batch of the 10 000 samples per
    step (bounded sample); `budget_s` stops early once that much time has been spent (>= 3 steps)."""
    from oracle import atlas_oracle as O
    if threads:
        torch.set_num_threads(threads)
    video = O.Video(**{k: v for k, v in data.items() if k != "clean"})
    torch.manual_seed(0)
    mp = [p.requires_grad_(True) for p in O.init_mlp(O.MAPPING_SPEC)]
    ap = [p.requires_grad_(True) for p in O.init_mlp(O.ATLAS_SPEC)]
    opt = O.make_optimizer(mp, ap)
    npix = video.H * video.W * video.T
    g = torch.Generator().manual_seed(1)
    batch = max(64, int(BATCH * fraction))
    total, done = 0.0, 0
    for i in range(warmup + steps):
        it = 0 if i < warmup + (steps + 1) // 2 else 6000           # half with / half without global rigidity
        inds = torch.randint(npix, (batch, 1), generator=g)
        t0 = time.perf_counter()
        O.train_iteration(video, mp, ap, opt, inds, it)
        if i >= warmup:
            total += time.perf_counter() - t0
            done += 1
            if budget_s is not None and done >= 3 and total >= budget_s:
                break
    return float(total), done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("B200_PRECISION", "auto"), choices=["auto", "tc", "fp32"])
    ap.add_argument("--cpu-sample-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    K, Wm = args.steps, max(args.warmup, 0)
    from b200 import synth

    config = {"workload": "stage-1 atlas loop, 80 frames 768x432, 10000 samples/iter, config_flow_100.json "
                          "coefficients, no segmentation (BASELINE.json configs[1])",
              "frames": T, "height": H, "width": W, "samples_batch": BATCH,
              "regime": "first half of the timed steps with the global rigidity term (i<=5000), second half without",
              "l2": "per-step working set (~1 GB of activations + random gathers from 1.7 GB of pixel records) "
                    "exceeds the 126 MB L2; no explicit flush"}

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        data = synth.throughput_set(H, W, T, seed=0)
        cores = os.cpu_count() or 1
        threads = min(cores, CPU_THREADS)
        # calibrate one full iteration, then size the per-step sample so that K steps take ~150 s
        t_full, _ = run_oracle_steps(data, 1, 1, threads=threads)
        fraction = min(1.0, 150.0 / max(K * t_full, 1e-9))
        secs, done = run_oracle_steps(data, K, 1, threads=threads, fraction=fraction)
        batch = max(64, int(BATCH * fraction))
        val = done * (batch / BATCH) / secs          # full-iteration equivalents per second
        line = {"impl": "reference", "metric": "atlas_iters_per_sec", "value": val, "unit": "it/s", "n_gpus": args.gpus,
                "steps": K, "warmup": 1, "ms_per_step": 1000.0 * secs / done, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "it/s", "cores": threads, "kind": "port",
                                 "sample": f"{done} steps of {batch} samples each ({batch / BATCH:.3f} of an iteration; "
                                           f"cost is linear in the samples) of the oracle restatement of "
                                           f"src/stage1_neural_atlas.py:151-231, torch CPU fp32, {threads} threads of "
                                           f"{cores} cores"},
                "e2e": {"value": val, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    from b200 import _native as N
    from b200 import atlas as A
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("B200_NCCL_DEBUG", "WARN")   # keep stdout to the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    lib = N.lib()
    prec = args.precision
    if prec == "auto":
        prec = "tc" if lib.b200_device_supports_tc() else "fp32"
    precision = N.PREC_TC if prec == "tc" else N.PREC_FP32

    data = synth.throughput_set(H, W, T, seed=0)
    t0, t1 = A.frame_range(rank, world, T)
    video = A.DeviceVideo.from_reference_layout(data, dev, t0, t1)
    trainer = A.AtlasTrainer(video, {"samples_batch": BATCH}, precision=precision, device=dev, process_group=pg)
    torch.manual_seed(0)
    trainer.init_like_reference()
    if world > 1:
        dist.broadcast(trainer.params, 0)
    npix = H * W * T
    gen = torch.Generator().manual_seed(1)            # same stream on every rank -> identical index batches
    total = Wm + K
    inds_cpu = torch.randint(npix, (total, BATCH), generator=gen)
    inds_dev = inds_cpu.to(dev)
    if rank != 0 or args.no_cpu_baseline:
        del data
    half = Wm + (K + 1) // 2
    it_of = lambda i: 0 if i < half else 6000

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernel timer on the dominant launch site
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); ev1.record()
    tag = 1   # B200_TAG_MAP_FWD
    lib.b200_set_kernel_timer(C_void(ev0), C_void(ev1), tag)

    # both graphs are captured during warm-up
    launches0 = lib.b200_launch_count()
    trainer.indices.copy_(inds_dev[0]); trainer.step(0)
    n_g = lib.b200_launch_count() - launches0
    launches0 = lib.b200_launch_count()
    trainer.indices.copy_(inds_dev[0]); trainer.step(6000)
    n_ng = lib.b200_launch_count() - launches0
    # each first call = 1 eager warm-up + 1 capture
    per_step = {True: n_g // 2, False: n_ng // 2}
    for i in range(Wm):
        trainer.indices.copy_(inds_dev[i]); trainer.step(it_of(i))
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler: sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(Wm, total):
        trainer.indices.copy_(inds_dev[i]); trainer.step(it_of(i))
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    losses_last = trainer.losses.cpu().numpy().copy()
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    value = K / (ms / 1000.0)

    # ---- e2e: host index batches in, loss vector out, every step
    k_e2e = max(10, K // 2)
    kern_ms = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(k_e2e):
        i = Wm + (j * 2) % K
        trainer.step_host(inds_cpu[i], it_of(i))
        kern_ms.append(ev0.elapsed_time(ev1))      # step_host synchronises: the tagged kernel's time for this step
    e1.record()
    barrier()
    if sampler:
        sampler.stop_flag = True
    ems = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ems], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ems = float(tms)
    e2e_val = k_e2e / (ems / 1000.0)
    lib.b200_set_kernel_timer(None, None, 0)

    if rank == 0:
        peaks, how = measured_peaks()
        n_f = float(losses_last[6]) / world
        n_b = float(losses_last[7]) / world
        flop_step = 0.5 * (algorithmic_flop(True, n_f, n_b) + algorithmic_flop(False, n_f, n_b))
        k_ms = float(np.median(kern_ms)) if kern_ms else None
        # tagged kernel: one 256x256 hidden layer of the mapping forward over all row groups
        # (fp32 path) / the whole fused mapping forward (tensor-core path)
        cap = (BATCH + 127) // 128 * 128
        if precision == N.PREC_FP32:
            kern_flop = 2.0 * 256 * 256 * ((9 + 7) / 2.0 * BATCH / world)
            kern_name = "sgemm_kernel<true,true> (mapping hidden layer, fwd)"
        else:
            kern_flop = 2.0 * MAC_MAP * ((9 + 7) / 2.0 * BATCH / world)
            kern_name = "tc_mlp_forward<mapping>"
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        achieved = kern_flop / (k_ms * 1e-3) / 1e12 if k_ms else None
        line = {"metric": "atlas_iters_per_sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": K,
                "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None,
                "dtype": "fp32" if precision == N.PREC_FP32 else "fp32 (2-term fp16 split on tcgen05, fp32 accumulate)",
                "data": "synthetic", "config": dict(config, parallelism=f"frame-sharded dp{world}",
                                                    precision=prec, cuda_graph=True),
                "e2e": {"value": e2e_val, "unit": "it/s", "h2d_bytes_per_step": BATCH * 8,
                        "d2h_bytes_per_step": N.LOSS_FLOATS * 4, "steps": k_e2e},
                "gpu_launches": int(sum(per_step[it_of(i) == 0] for i in range(Wm, total))),
                "launches_per_step": {"with_global": per_step[True], "without": per_step[False]},
                "roofline": {"bound": "tensor", "kernel": kern_name, "achieved": achieved, "peak": peak_tf,
                             "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                             "kernel_ms": k_ms, "peak_source": f"{how} bf16_tflops_sustained",
                             # dram__bytes_read.sum + dram__bytes_write.sum of one tc_fwd_kernel<mapping> launch, from the
                             # ncu --set full capture in profiles/r1_tc_v1_ncu_full_summary.csv (4.9 MB + 314.1 MB: the fp16
                             # activation images written for the backward pass; algorithmic image bytes = rows*256*4*5 B)
                             "traffic": (319.0e6 if precision != N.PREC_FP32 and world == 1 else None),
                             "step_algorithmic_gflop": flop_step / 1e9,
                             "step_tflops": flop_step * value / 1e12},
                "clocks": sampler.summary() if sampler else None,
                "losses_last": [float(x) for x in losses_last[:6]]}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            threads = min(cores, CPU_THREADS)
            secs, done = run_oracle_steps(data, args.cpu_sample_steps, 1, threads=threads, budget_s=20.0)
            line["cpu_baseline"] = {"value": done / secs, "unit": "it/s", "cores": threads, "kind": "port",
                                    "sample": f"{done} full iterations of the oracle (torch CPU fp32 restatement of "
                                              f"src/stage1_neural_atlas.py:151-231) on the same synthetic video, "
                                              f"{threads} threads of {cores} cores"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # orderly teardown: every rank is past its last collective; drop the captured graphs (they hold NCCL
        # work) before the communicator, and never block the launcher on a straggling destructor
        torch.cuda.synchronize()
        dist.barrier()
        trainer._graphs.clear()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def C_void(ev):
    import ctypes
    return ctypes.c_void_p(ev.cuda_event)


if __name__ == "__main__":
    main()
