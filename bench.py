#!/usr/bin/env python
"""Benchmark of the FloWaveNet flow pass on B200 (contract: see the task's bench.py section).

Default workload = BASELINE config C3: hparams8000.py model (5 blocks x 6 flows x 2 layers, 80 mels, hop 96),
INVERSE SYNTHESIS of 32 utterances x 10.008 s (T = 80 064 samples) per GPU, mixed precision
(bf16 operands / fp32 accumulate on tcgen05; fp32 flow variable).  Weak scaling: every GPU synthesises its own
32 utterances, no data-path collective.  One "step" = one full reverse pass over the batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c4] [--impl ours|reference]

Prints ONE JSON line (rank 0).  --impl reference times the CPU restatement of the reference graph
(oracle/flowavenet_oracle.py; TensorFlow 1.12 itself cannot be installed here) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, direction, B, n_frames, dtype, description)
    "c3": ("hparams8000", "reverse", 32, 834, "bfloat16", "C3: hparams8000 inverse synthesis, 32 utterances x 10.008 s (T=80064) per GPU, mixed precision"),
    "c1": ("hparams", "reverse", 1, 87, "float32", "C1: hparams inverse synthesis, 1 utterance x 1.01 s (T=22272), fp32"),
    "c2": ("hparams", "forward", 8, 63, "float32", "C2: hparams forward log-likelihood, 8 x 16128 samples, fp32"),
    "c4": ("hparams", "reverse", 1, 5168, "bfloat16", "C4: hparams inverse synthesis, 1 utterance x 60 s (T=1323008), mixed precision, single GPU"),
    "c1m": ("hparams", "reverse", 1, 87, "bfloat16", "C1 shape in mixed precision"),
    "c4s": ("hparams", "reverse", 1, 5168, "bfloat16", "C4: hparams inverse synthesis, ONE 60 s utterance (T=1323008) sharded by time chunk across the GPUs, "
            "15616-sample receptive-field halos exchanged with NCCL send/recv, mixed precision"),
}
TRAIN_WORKLOADS = {
    # name: (preset, B per GPU, n_frames, gin_channels, n_speakers, description)
    "c5": ("hparams", 8, 25, 16, 7, "C5: hparams training step (forward + backward + tower-average + clip + Adam + re-pack), 8 utterances x 6400 samples "
           "per GPU (hparams.py:28,36), fp32 variables/activations/accumulation, all GEMMs (forward, dgrad, wgrad) on tcgen05 via 3-way bf16 operand split"),
}
MFLOP_PER_SAMPLE = {"hparams": 17.31174, "hparams8000": 15.448592}  # SURVEY 8d algorithmic 2*MAC per audio sample


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_step(hp_kw, direction, B, n_frames, seed=1234):
    """One pass of the CPU restatement (the `port` CPU baseline).  Returns (samples, seconds, threads)."""
    import torch
    from oracle import flowavenet_oracle as O
    hp = O.HP(**hp_kw)
    params = O.synthetic_params(hp, seed)
    a, c = O.synthetic_inputs(hp, B, n_frames, seed + 1, "z" if direction == "reverse" else "x")
    torch.set_num_threads(os.cpu_count() or 1)
    fn = (lambda: O.reverse(params, hp, a, c, torch.float32)) if direction == "reverse" else (lambda: O.forward(params, hp, a, c, torch.float32))
    with torch.no_grad():
        fn()  # warm-up
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
    return B * n_frames * hp.hop, best, torch.get_num_threads()


def cpu_sample_shape(preset, B, n_frames):
    # bounded sample of the same workload: ~10-30 s of CPU work in total (1 warm-up + 3 timed passes)
    if preset == "hparams8000":
        return min(B, 2), min(n_frames, 250)   # 2 x 3 s of 8 kHz audio
    return 1, min(n_frames, 87)                # 1 x 1.01 s of 22.05 kHz audio


def oracle_train_step(hp_kw, B, n_frames, seed=1234):
    """loss + gradients of one tower on the CPU restatement (the `port` baseline of the training step)."""
    import torch
    from oracle import flowavenet_oracle as O
    from oracle import flowavenet_train_oracle as TO
    hp = O.HP(**hp_kw)
    params = O.synthetic_params(hp, seed)
    x, c = O.synthetic_inputs(hp, B, n_frames, seed + 1, "x")
    torch.set_num_threads(os.cpu_count() or 1)
    TO.loss_and_grads(params, hp, x, c, torch.float32)  # warm-up
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        TO.loss_and_grads(params, hp, x, c, torch.float32)
        best = min(best, time.perf_counter() - t0)
    return B * n_frames * hp.hop, best, torch.get_num_threads()


def main_train(args):
    """C5: one training step per `step` (train.py:236), data-parallel towers = processes, gradients averaged with one NCCL all-reduce."""
    preset, B, n_frames, gin, nspk, desc = TRAIN_WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric = "training audio samples/sec"
    import tf_flowavenet_b200 as P
    hp_ref = getattr(P, preset)
    hop = int(np.prod(hp_ref.upsample_scales))
    T = n_frames * hop
    hp_kw = dict(n_block=hp_ref.n_block, upsample_scales=tuple(hp_ref.upsample_scales))
    config = {"workload": desc, "preset": preset, "direction": "train", "utterances_per_gpu": B, "samples_per_utterance": T,
              "global_batch": B * max(world, args.gpus), "split_terms": args.split_terms, "gin_channels": gin, "n_speakers": nspk, "sample_rate": hp_ref.sample_rate,
              "parallelism": "dp%d: one tower per GPU, all-reduce(avg) of the flat fp32 gradient (181 M floats)" % max(world, args.gpus),
              "l2_policy": "per-step working set (tape ~3 GB + 2.2 GB of weights and operands) >> 126 MB L2; no explicit flush needed"}
    if args.impl == "reference":
        if rank != 0:
            return
        samples, secs, thr = oracle_train_step(hp_kw, 1, n_frames)
        val = samples / secs
        sample = "1 utterance x %d frames (%d samples): loss + autograd gradients of the same model, fp32, best of 2 after 1 warm-up" % (n_frames, samples)
        emit({"impl": "reference", "metric": metric, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "samples/s", "cores": thr, "kind": "port", "sample": sample,
                                           "note": "CPU restatement (PyTorch-CPU autograd) of the reference training graph; optimizer update not included"},
                          "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    import torch
    import torch.distributed as dist
    from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params
    from tf_flowavenet_b200.train import Trainer
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    net = P.FloWaveNet(P.HParams(**{**hp_ref.values(), "dtype": "float32", "gin_channels": gin, "n_speakers": nspk}), variables=P.VariableStore())
    net.load_variables(synthetic_params(net.variable_shapes(), seed=1234))   # same variables on every tower
    x_np, c_np = synthetic_inputs(hop, 80, B, n_frames, 1234 + 5 + rank, "x")  # each tower draws its own batch (dataset.py:34-38)
    g_np = np.random.default_rng(77 + rank).integers(0, nspk, size=(B,)).astype(np.int32)
    x_pin, c_pin, g_pin = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(c_np).pin_memory(), torch.from_numpy(g_np).pin_memory()
    x_dev, c_dev, g_dev = x_pin.cuda(), c_pin.cuda(), g_pin.cuda()
    tr = Trainer(net, split_terms=args.split_terms)
    tr.train_step(x_dev, c_dev, g_dev, init=True)  # ActNorm DDI step (train.py:221,229)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    phase = [0.0, 0.0, 0.0]

    def step_dev(timed=False):
        if timed:
            ev[0].record()
        tr.loss_and_grads(x_dev, c_dev, g_dev)
        if timed:
            ev[1].record()
        tr.average_gradients()
        if timed:
            ev[2].record()
        tr.apply_gradients()
        if timed:
            ev[3].record()
            ev[3].synchronize()
            for i in range(3):
                phase[i] += ev[i].elapsed_time(ev[i + 1])

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = net.last_launches() * args.steps
    for _ in range(args.steps):   # same steps again with CUDA events between the phases
        step_dev(timed=True)
    barrier()
    # end to end: pinned host batch -> device, step, loss back to the host, every step
    def step_e2e():
        xd, cd, gd = x_pin.cuda(non_blocking=True), c_pin.cuda(non_blocking=True), g_pin.cuda(non_blocking=True)
        info = tr.train_step(xd, cd, gd)
        return float(info["loss"])
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    samples_per_step = B * T * world
    value = samples_per_step * args.steps / (ms * 1e-3)
    flop_step = 3.0 * MFLOP_PER_SAMPLE[preset] * 1e6 * B * T   # forward + dgrad + wgrad, per GPU
    ach = flop_step / (phase[0] / args.steps * 1e-3) / 1e12
    roof = {"kernel": "forward+backward of one tower (tc3_gemm_kernel: fwd + dgrad GEMMs; wgrad_kernel: CUDA-core wgrad)", "bound": "tensor",
            "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
            "peak_source": "%s bf16_tflops_sustained; the fp32-storage GEMMs spend %d bf16 MMA terms per product, so the attainable fraction "
                           "of this peak is 1/%d for the tensor-core GEMMs" % (pk["src"], args.split_terms, args.split_terms),
            "traffic": None, "algorithmic_flop_per_step_per_gpu": flop_step,
            "phases_ms_per_step": {"loss_and_grads": phase[0] / args.steps, "allreduce_avg": phase[1] / args.steps,
                                   "clip_adam_repack": phase[2] / args.steps}}
    tp = os.path.join(ROOT, "profiles", "r1_c5_traffic.json")
    if os.path.exists(tp):   # DRAM bytes of ONE ncu --set full capture of the kernel with the largest share of the step
        tj = json.load(open(tp))
        roof["traffic"] = tj["traffic_bytes_per_launch"]
        roof["traffic_detail"] = tj
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "roofline": roof, "clocks": clk,
            "e2e": {"value": samples_per_step * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": x_pin.numel() * 4 + c_pin.numel() * 4 + g_pin.numel() * 4, "d2h_bytes_per_step": 4,
                    "api": "Trainer.train_step -> fwn_loss_and_grads + all_reduce + fwn_apply_gradients", "last_loss": loss},
            "gpu_launches": int(launches)}
    if not args.no_cpu_baseline and world == 1:
        samples, secs, thr = oracle_train_step(hp_kw, 1, n_frames)
        line["cpu_baseline"] = {"value": samples / secs, "unit": "samples/s", "cores": thr, "kind": "port",
                                "sample": "1 utterance x %d frames (%d samples): loss + autograd gradients, fp32, best of 2 after 1 warm-up" % (n_frames, samples),
                                "note": "CPU restatement (PyTorch-CPU autograd) of the reference training graph; baseline only"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Route fd 1 to stderr while the benchmark runs (NCCL prints its version banner to stdout on the first multi-rank collective);
    emit() writes the one JSON line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + sorted(TRAIN_WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split-terms", type=int, default=3, choices=[3, 6], help="c5: bf16 products per fp32 product in the training GEMMs")
    args = ap.parse_args()
    if args.workload in TRAIN_WORKLOADS:
        return main_train(args)
    preset, direction, B, n_frames, dtype, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric = "synthesis audio samples/sec" if direction == "reverse" else "forward log-likelihood audio samples/sec"

    import tf_flowavenet_b200 as P
    hp_ref = getattr(P, preset)
    hp_kw = dict(n_block=hp_ref.n_block, upsample_scales=tuple(hp_ref.upsample_scales))
    hop = int(np.prod(hp_ref.upsample_scales))
    T = n_frames * hop
    config = {"workload": desc, "preset": preset, "direction": direction, "utterances_per_gpu": B, "samples_per_utterance": T,
              "global_utterances": B * max(world, args.gpus), "sample_rate": hp_ref.sample_rate, "parallelism": "utterance-sharded x%d, no collective" % max(world, args.gpus),
              "l2_policy": "working set (%.1f GB of activations per pass) >> 126 MB L2; no explicit flush needed" % (B * T / 2 * 256 * 2 * 5 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        sb, sf = cpu_sample_shape(preset, B, n_frames)
        samples, secs, thr = oracle_step(hp_kw, direction, sb, sf)
        val = samples / secs
        sample = "%d utterance(s) x %d frames (%d samples) of the same model, fp32, best of 3 after 1 warm-up" % (sb, sf, samples)
        emit({"impl": "reference", "metric": metric, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "samples/s", "cores": thr, "kind": "port", "sample": sample,
                                           "note": "CPU restatement of the reference TF-1.12 graph (PyTorch-CPU); TF 1.12 is not installable on Python 3.12"},
                          "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "xrt_at_22050": val / 22050.0})
        return

    import torch
    import torch.distributed as dist
    from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    net = P.FloWaveNet(P.HParams(**{**hp_ref.values(), "dtype": dtype}), variables=P.VariableStore())
    net.load_variables(synthetic_params(net.variable_shapes(), seed=1234))
    sharded = args.workload == "c4s"
    if sharded:
        assert n_frames % world == 0, "c4s needs the frame count to divide by the number of GPUs"
        full_a, full_c = synthetic_inputs(hop, 80, B, n_frames, 1234 + 3, "z")     # same utterance on every rank ...
        n_frames //= world                                                          # ... of which this rank owns one time chunk
        a_np = np.ascontiguousarray(full_a[:, rank * n_frames * hop:(rank + 1) * n_frames * hop])
        c_np = np.ascontiguousarray(full_c[:, rank * n_frames:(rank + 1) * n_frames])
        T = n_frames * hop
        config.update(scaling_note="strong scaling: total work fixed (one utterance), chunk per GPU = %d samples + halos" % T,
                      parallelism="time-chunk sharded x%d, halo exchange = NCCL P2P with rank+-1" % world)
    else:
        a_np, c_np = synthetic_inputs(hop, 80, B, n_frames, 1234 + 3 + rank, "z" if direction == "reverse" else "x")
    # ActNorm data-dependent init on a small batch of the same distribution (train.py:221,229)
    xi, ci = synthetic_inputs(hop, 80, min(B, 2), min(n_frames, 64), 99, "x")
    net.initialize_actnorm(torch.from_numpy(xi).cuda(), torch.from_numpy(ci).cuda())
    a_pin, c_pin = torch.from_numpy(a_np).pin_memory(), torch.from_numpy(c_np).pin_memory()
    a_dev, c_dev = a_pin.cuda(), c_pin.cuda()
    out_pin = torch.empty(B, T, 1, dtype=torch.float32).pin_memory()

    side = torch.cuda.Stream()   # a real (capturable) stream: the library replays each pass as a CUDA graph on it

    def step_dev():
        with torch.cuda.stream(side):
            return _step_dev()

    def _step_dev():
        if sharded:
            return net.reverse_sharded(a_dev, c_dev, rank, world)
        return net.reverse(a_dev, c_dev) if direction == "reverse" else net.forward(a_dev, c_dev)

    def step_e2e():
        if sharded:  # host chunk in, halo exchange on device, host chunk out
            x = net.reverse_sharded(a_pin.cuda(non_blocking=True), c_pin.cuda(non_blocking=True), rank, world)
            out_pin.copy_(x, non_blocking=True)
            torch.cuda.synchronize()
        elif direction == "reverse":
            net.reverse_host(a_pin, c_pin, out_pin)
        else:
            net.forward_host(a_pin, c_pin)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(side)
    for _ in range(args.steps):
        step_dev()
    ev1.record(side)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    launches = net.last_launches() * args.steps
    # same K steps again with a CUDA-event pair around every kernel launch of the pass (per-family durations for the roofline);
    # the instrumented loop launches eagerly (the un-instrumented one replays the pass as a CUDA graph)
    net.profile(True)
    net.profile_read()
    barrier()
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(side)
    for _ in range(args.steps):
        step_dev()
    evp1.record(side)
    barrier()
    ms_prof = evp0.elapsed_time(evp1)
    prof = net.profile_read()
    net.profile(False)

    # end to end through the public host-buffer API: H2D of z and mel, pass, D2H of the waveform, every step
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    samples_per_step = B * T * world
    value = samples_per_step * args.steps / (ms * 1e-3)
    g_ms, g_n, g_flop = prof["gate_gemm"]
    ach = g_flop / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    roof = {"kernel": "tc_gemm_kernel<EPI_GATE,256> (dilated conv k=3 + cond 1x1 + tanh*sigmoid)" if dtype == "bfloat16" else
            ("simt_gemm_kernel<EPI_GATE>" if os.environ.get("FWN_FP32_ENGINE") == "simt" else
             "tc3_gemm_kernel<EPI_GATE,128> (fp32 parity mode: 6 bf16 MMA terms per product, so <= 1/6 of the bf16 peak is attainable)"),
            "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
            "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % pk["src"], "traffic": None,
            "launches": g_n, "avg_launch_ms": g_ms / max(g_n, 1), "share_of_step": g_ms / ms_prof,
            "instrumented_ms_per_step": ms_prof / args.steps,
            "families": {k: {"ms": v[0], "launches": v[1], "achieved": (v[2] / (v[0] * 1e-3) / (1e9 if k == "upsample" else 1e12)) if v[0] > 0 else 0.0,
                             "unit": "GB/s" if k == "upsample" else "TFLOP/s"} for k, v in prof.items()},
            "whole_pass_tflops": value * MFLOP_PER_SAMPLE[preset] * 1e6 / 1e12 / world}
    tp = os.path.join(ROOT, "profiles", "r1_gate_traffic.json")
    if dtype == "bfloat16" and args.workload == "c3" and os.path.exists(tp):
        tj = json.load(open(tp))
        # DRAM bytes of ONE ncu --set full capture of this kernel (its block-0 launch), next to that launch's algorithmic bytes
        roof["traffic"] = tj["traffic_bytes_per_launch"]
        roof["traffic_detail"] = {k: tj[k] for k in ("launch", "source", "algorithmic_bytes_per_launch", "algorithmic_flop_per_launch")}
    ups = prof["upsample"]
    if ups[0] > 0:
        roof["upsample_hbm_frac"] = ups[2] / (ups[0] * 1e-3) / 1e9 / pk["hbm_gbs"]
    bytes_in = a_pin.numel() * 4 + c_pin.numel() * 4
    bytes_out = out_pin.numel() * 4 if direction == "reverse" else 8
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "bf16" if dtype == "bfloat16" else "f32", "data": "synthetic", "config": config, "roofline": roof, "clocks": clk,
            "e2e": {"value": samples_per_step * args.steps / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": bytes_in,
                    "d2h_bytes_per_step": bytes_out, "api": "FloWaveNet.reverse_host -> fwn_reverse_host (pinned host buffers)"},
            "gpu_launches": int(launches), "xrt_at_22050": value / 22050.0, "xrt_at_native_rate": value / hp_ref.sample_rate,
            "xrt_per_gpu_at_22050": value / 22050.0 / world}
    if not args.no_cpu_baseline and world == 1:
        sb, sf = cpu_sample_shape(preset, B, n_frames)
        samples, secs, thr = oracle_step(hp_kw, direction, sb, sf)
        line["cpu_baseline"] = {"value": samples / secs, "unit": "samples/s", "cores": thr, "kind": "port",
                                "sample": "%d utterance(s) x %d frames (%d samples) of the same model, fp32, best of 3 after 1 warm-up" % (sb, sf, samples),
                                "note": "CPU restatement of the reference TF-1.12 graph (PyTorch-CPU); baseline only"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
