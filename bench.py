#!/usr/bin/env python
"""Benchmark of the FloWaveNet flow pass on B200 (contract: see the task's bench.py section).

Headline workload = BASELINE config C3: hparams8000.py model (5 blocks x 6 flows x 2 layers, 80 mels, hop 96),
INVERSE SYNTHESIS of 32 utterances x 10.008 s (T = 80 064 samples) per GPU, mixed precision
(16-bit operands / fp32 accumulate on tcgen05; fp32 flow variable).  Weak scaling: every GPU synthesises its own
32 utterances, no data-path collective.  One "step" = one full reverse pass over the batch.

The default invocation then runs, at the same N, the two workloads of BASELINE.json that COMMUNICATE and attaches them under
`secondary` of the one JSON line:
  c4s  one 60 s utterance sharded by time chunk over the N GPUs, receptive-field halos exchanged with NCCL send/recv  (strong scaling)
  c5   data-parallel training step, one tower per GPU, gradient all-reduce overlapped with the backward pass          (weak scaling)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c1m|c2|c4|c4s|c5] [--impl ours|reference] [--no-secondary]

Prints ONE JSON line (rank 0).  --impl reference times the CPU restatement of the reference graph
(oracle/flowavenet_oracle.py; TensorFlow 1.12 itself cannot be installed here) on the box's host cores: every step is one pass
over a bounded excerpt of the workload, W warm-up + K timed steps exactly as asked.
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
           "per GPU (hparams.py:28,36), global-condition multi-speaker inputs"),
}
MFLOP_PER_SAMPLE = {"hparams": 17.31174, "hparams8000": 15.448592}  # SURVEY 8d algorithmic 2*MAC per audio sample
DT_NAME = {"bfloat16": "bf16", "float16": "f16", "float32": "f32"}


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


class Dist:
    """torch.distributed plumbing shared by every workload of one invocation (one process per GPU, NCCL)."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.up = False

    def init(self):
        import torch
        import torch.distributed as dist
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        if self.world > 1 and not self.up:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.up = True

    def barrier(self):
        import torch
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        import torch
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return [float(v) for v in t]

    def close(self):
        if self.up:
            import torch.distributed as dist
            dist.destroy_process_group()
            self.up = False


# ------------------------------------------------------------------------------------------------ CPU restatement (baseline / reference arm)
def oracle_runner(hp_kw, direction, B, n_frames, seed=1234):
    """-> (fn, samples per call, threads): one pass of the CPU restatement over a bounded excerpt."""
    import torch
    from oracle import flowavenet_oracle as O
    hp = O.HP(**hp_kw)
    params = O.synthetic_params(hp, seed)
    a, c = O.synthetic_inputs(hp, B, n_frames, seed + 1, "z" if direction == "reverse" else "x")
    torch.set_num_threads(os.cpu_count() or 1)
    if direction == "reverse":
        def fn():
            with torch.no_grad():
                O.reverse(params, hp, a, c, torch.float32)
    elif direction == "forward":
        def fn():
            with torch.no_grad():
                O.forward(params, hp, a, c, torch.float32)
    else:
        from oracle import flowavenet_train_oracle as TO

        def fn():
            TO.loss_and_grads(params, hp, a, c, torch.float32)
    return fn, B * n_frames * hp.hop, torch.get_num_threads()


def time_cpu(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return time.perf_counter() - t0


def cpu_sample_shape(preset, direction, B, n_frames):
    # bounded excerpt of the same workload: about a second of CPU work per pass, so W + K passes end within a few minutes
    if direction == "train":
        return 1, min(n_frames, 25)             # 1 utterance x 6400 samples (one of the tower's 8)
    if preset == "hparams8000":
        return min(B, 2), min(n_frames, 250)    # 2 x 3 s of 8 kHz audio
    return 1, min(n_frames, 87)                 # 1 x 1.01 s of 22.05 kHz audio


def cpu_baseline(preset, hp_kw, direction, B, n_frames, warmup=1, steps=3):
    sb, sf = cpu_sample_shape(preset, direction, B, n_frames)
    fn, samples, thr = oracle_runner(hp_kw, direction, sb, sf)
    secs = time_cpu(fn, warmup, steps)
    what = "loss + autograd gradients" if direction == "train" else "one %s pass" % direction
    return {"value": samples * steps / secs, "unit": "samples/s", "cores": thr, "kind": "port",
            "sample": "%d utterance(s) x %d frames (%d samples) of the same model, fp32: %s per step, %d timed steps after %d warm-up" %
                      (sb, sf, samples, what, steps, warmup),
            "note": "CPU restatement of the reference TF-1.12 graph (PyTorch-CPU); TF 1.12 is not installable on Python 3.12; baseline only",
            "ms_per_step": secs / steps * 1e3}


def reference_arm(args, dist, metric, config, preset, hp_kw, direction, B, n_frames):
    """bench.py --impl reference: exactly W warm-up + K timed steps, each one pass of the CPU restatement over the bounded excerpt."""
    if dist.rank != 0:
        return
    warm, steps = max(args.warmup, 1), max(args.steps, 1)
    cb = cpu_baseline(preset, hp_kw, direction, B, n_frames, warm, steps)
    val = cb["value"]
    cfg = dict(config)
    cfg["measured_on"] = "bounded excerpt per step (see cpu_baseline.sample); throughput is per-sample, the full-size batch is this excerpt repeated"
    emit({"impl": "reference", "metric": metric, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
          "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": cfg, "cpu_baseline": cb, "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "xrt_at_22050": val / 22050.0})


# ------------------------------------------------------------------------------------------------ training step (C5)
def run_train(args, dist, workload="c5", with_cpu=True):
    """C5: one training step per `step` (train.py:236), data-parallel towers = processes, bucketed gradient all-reduce overlapped
    with the backward pass.  Returns the JSON line (rank 0) or None."""
    preset, B, n_frames, gin, nspk, desc = TRAIN_WORKLOADS[workload]
    rank, world = dist.rank, dist.world
    metric = "training audio samples/sec"
    import tf_flowavenet_b200 as P
    hp_ref = getattr(P, preset)
    hop = int(np.prod(hp_ref.upsample_scales))
    T = n_frames * hop
    hp_kw = dict(n_block=hp_ref.n_block, upsample_scales=tuple(hp_ref.upsample_scales))
    tdt = args.train_dtype
    config = {"workload": desc + "; compute dtype " + tdt, "preset": preset, "direction": "train", "utterances_per_gpu": B, "samples_per_utterance": T,
              "global_batch": B * max(world, args.gpus), "train_dtype": tdt, "split_terms": args.split_terms if tdt == "float32" else None,
              "gin_channels": gin, "n_speakers": nspk, "sample_rate": hp_ref.sample_rate,
              "parallelism": "dp%d: one tower per GPU, all-reduce(avg) of the flat fp32 gradient (181 M floats) in per-block buckets overlapped "
                             "with the backward pass" % max(world, args.gpus),
              "l2_policy": "per-step working set (tape + weights, several GB) >> 126 MB L2; no explicit flush needed"}
    if args.impl == "reference":
        return reference_arm(args, dist, metric, config, preset, hp_kw, "train", B, n_frames)

    import torch
    from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params
    from tf_flowavenet_b200.train import Trainer
    dist.init()
    net = P.FloWaveNet(P.HParams(**{**hp_ref.values(), "dtype": "float32", "gin_channels": gin, "n_speakers": nspk}), variables=P.VariableStore())
    net.load_variables(synthetic_params(net.variable_shapes(), seed=1234))   # same variables on every tower
    x_np, c_np = synthetic_inputs(hop, 80, B, n_frames, 1234 + 5 + rank, "x")  # each tower draws its own batch (dataset.py:34-38)
    g_np = np.random.default_rng(77 + rank).integers(0, nspk, size=(B,)).astype(np.int32)
    x_pin, c_pin, g_pin = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(c_np).pin_memory(), torch.from_numpy(g_np).pin_memory()
    x_dev, c_dev, g_dev = x_pin.cuda(), c_pin.cuda(), g_pin.cuda()
    tr = Trainer(net, split_terms=args.split_terms, compute_dtype=tdt)
    tr.train_step(x_dev, c_dev, g_dev, init=True)  # ActNorm DDI step (train.py:221,229)

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

    steps, warm = args.steps, max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    dist.barrier()
    clocks = ClockSampler(dist.local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(steps):
        step_dev()
    e1.record()
    dist.barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = net.last_launches() * steps
    for _ in range(steps):   # same steps again with CUDA events between the phases
        step_dev(timed=True)
    dist.barrier()

    # end to end: pinned host batch -> device, step, loss back to the host, every step
    def step_e2e():
        xd, cd, gd = x_pin.cuda(non_blocking=True), c_pin.cuda(non_blocking=True), g_pin.cuda(non_blocking=True)
        info = tr.train_step(xd, cd, gd)
        return float(info["loss"])
    step_e2e()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step_e2e()
    dist.barrier()
    e2e_s = time.perf_counter() - t0
    ms, e2e_ms = dist.max_over_ranks([ms, e2e_s * 1e3])
    if rank != 0:
        return None
    pk = peaks()
    samples_per_step = B * T * world
    value = samples_per_step * steps / (ms * 1e-3)
    flop_step = 3.0 * MFLOP_PER_SAMPLE[preset] * 1e6 * B * T   # forward + dgrad + wgrad, per GPU
    ach = flop_step / (phase[0] / steps * 1e-3) / 1e12
    terms = args.split_terms if tdt == "float32" else 1
    roof = {"kernel": ("forward+backward of one tower: tc3_gemm_kernel (fwd + dgrad, 3-way bf16 split of fp32 operands) + wgrad_tc3_kernel" if tdt == "float32"
                       else "forward+backward of one tower: tc_gemm_kernel family (bf16 operands) + wgrad_tc_kernel"),
            "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
            "peak_source": "%s bf16_tflops_sustained; %d bf16 MMA term(s) per product" % (pk["src"], terms),
            "traffic": None, "algorithmic_flop_per_step_per_gpu": flop_step,
            "phases_ms_per_step": {"loss_and_grads": phase[0] / steps, "allreduce_avg_exposed": phase[1] / steps,
                                   "clip_adam_repack": phase[2] / steps}}
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DT_NAME[tdt],
            "data": "synthetic", "config": config, "roofline": roof, "clocks": clk,
            "allreduce_bytes_per_step": int(tr.param_floats()) * 4,
            "e2e": {"value": samples_per_step * steps / (e2e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": x_pin.numel() * 4 + c_pin.numel() * 4 + g_pin.numel() * 4, "d2h_bytes_per_step": 4,
                    "api": "Trainer.train_step -> fwn_loss_and_grads + bucketed all_reduce + fwn_apply_gradients", "last_loss": loss},
            "gpu_launches": int(launches)}
    if with_cpu and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(preset, hp_kw, "train", B, n_frames, 1, 2)
    del tr, net
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------ forward / inverse passes
def run_pass(args, dist, workload, with_cpu=True, steps=None):
    preset, direction, B, n_frames, dtype, desc = WORKLOADS[workload]
    if args.dtype and dtype != "float32":
        dtype = args.dtype
    rank, world = dist.rank, dist.world
    metric = "synthesis audio samples/sec" if direction == "reverse" else "forward log-likelihood audio samples/sec"
    steps = steps or args.steps

    import tf_flowavenet_b200 as P
    hp_ref = getattr(P, preset)
    hp_kw = dict(n_block=hp_ref.n_block, upsample_scales=tuple(hp_ref.upsample_scales))
    hop = int(np.prod(hp_ref.upsample_scales))
    T = n_frames * hop
    config = {"workload": desc + (" (%s operands)" % dtype if dtype != "float32" else ""), "preset": preset, "direction": direction,
              "utterances_per_gpu": B, "samples_per_utterance": T,
              "global_utterances": B * max(world, args.gpus), "sample_rate": hp_ref.sample_rate, "parallelism": "utterance-sharded x%d, no collective" % max(world, args.gpus),
              "l2_policy": "working set (%.1f GB of activations per pass) >> 126 MB L2; no explicit flush needed" % (B * T / 2 * 256 * 2 * 5 / 1e9)}
    if args.impl == "reference":
        return reference_arm(args, dist, metric, config, preset, hp_kw, direction, B, n_frames)

    import torch
    from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params
    dist.init()
    net = P.FloWaveNet(P.HParams(**{**hp_ref.values(), "dtype": dtype}), variables=P.VariableStore())
    net.load_variables(synthetic_params(net.variable_shapes(), seed=1234))
    sharded = workload == "c4s"
    halo_bytes = 0
    if sharded:
        assert n_frames % world == 0, "c4s needs the frame count to divide by the number of GPUs"
        full_a, full_c = synthetic_inputs(hop, 80, B, n_frames, 1234 + 3, "z")     # same utterance on every rank ...
        n_frames //= world                                                          # ... of which this rank owns one time chunk
        a_np = np.ascontiguousarray(full_a[:, rank * n_frames * hop:(rank + 1) * n_frames * hop])
        c_np = np.ascontiguousarray(full_c[:, rank * n_frames:(rank + 1) * n_frames])
        T = n_frames * hop
        halo = net.receptive_halo()
        sides = (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)
        halo_bytes = sides * B * (halo * 4 + (halo // hop) * 80 * 4)   # received per pass: z halo + mel halo, fp32
        config.update(scaling_note="strong scaling: total work fixed (one utterance), chunk per GPU = %d samples + halos" % T,
                      parallelism="time-chunk sharded x%d, halo exchange = NCCL P2P with rank+-1" % world, halo_samples_per_side=halo)
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
            if sharded:
                return net.reverse_sharded(a_dev, c_dev, rank, world)
            return net.reverse(a_dev, c_dev) if direction == "reverse" else net.forward(a_dev, c_dev)

    def step_e2e():
        if sharded:  # host chunk in, halo exchange on device, host chunk out
            x = net.reverse_sharded(a_pin.cuda(non_blocking=True), c_pin.cuda(non_blocking=True), rank, world)
            out_pin.copy_(x, non_blocking=True)
            torch.cuda.synchronize()
        elif direction == "reverse":
            net.reverse_host(a_pin, c_pin, out=out_pin)
        else:
            net.forward_host(a_pin, c_pin)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    dist.barrier()
    clocks = ClockSampler(dist.local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    ev0.record(side)
    for _ in range(steps):
        step_dev()
    ev1.record(side)
    dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    launches = net.last_launches() * steps
    # same K steps again with a CUDA-event pair around every kernel launch of the pass (per-family durations for the roofline);
    # the instrumented loop launches eagerly (the un-instrumented one replays the pass as a CUDA graph)
    net.profile(True)
    net.profile_read()
    dist.barrier()
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(side)
    for _ in range(steps):
        step_dev()
    evp1.record(side)
    dist.barrier()
    ms_prof = evp0.elapsed_time(evp1)
    prof = net.profile_read()
    net.profile(False)

    # end to end through the public host-buffer API: H2D of z and mel, pass, D2H of the waveform, every step
    for _ in range(2):
        step_e2e()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    dist.barrier()
    e2e_s = time.perf_counter() - t0
    ms, e2e_ms = dist.max_over_ranks([ms, e2e_s * 1e3])
    if rank != 0:
        return None

    pk = peaks()
    samples_per_step = B * T * world
    value = samples_per_step * steps / (ms * 1e-3)
    g_ms, g_n, g_flop = prof["gate_gemm"]
    ach = g_flop / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    mixed = dtype != "float32"
    # mixed modes, large launches: each ResBlock layer is ONE kernel (gate GEMM -> tanh*sigmoid -> res|skip 1x1, csrc/layer_tc.cu); its
    # family time and FLOP count then cover both GEMMs and the res_skip family is empty.  Small launches keep the two-launch path.
    fused = mixed and prof["res_skip_gemm"][1] < prof["gate_gemm"][1]
    roof = {"kernel": ("layer_kernel<2> (fused ResBlock layer: dilated conv k=3 + cond 1x1 + tanh*sigmoid + res|skip 1x1; deep-block launches: "
                       "tc_gemm_kernel<EPI_GATE,256> + <EPI_RES_SKIP>)" if fused else
                       "tc_gemm_kernel<EPI_GATE,256> (dilated conv k=3 + cond 1x1 + tanh*sigmoid)") if mixed else
            ("simt_gemm_kernel<EPI_GATE>" if os.environ.get("FWN_FP32_ENGINE") == "simt" else
             "tc3_gemm_kernel<EPI_GATE,128> (fp32 parity mode: 6 bf16 MMA terms per product, so <= 1/6 of the bf16 peak is attainable)"),
            "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
            "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % pk["src"], "traffic": None,
            "launches": g_n, "avg_launch_ms": g_ms / max(g_n, 1), "share_of_step": g_ms / ms_prof,
            "instrumented_ms_per_step": ms_prof / steps, "fused_layers": bool(fused),
            "families": {k: {"ms": v[0], "launches": v[1], "achieved": (v[2] / (v[0] * 1e-3) / (1e9 if k == "upsample" else 1e12)) if v[0] > 0 else 0.0,
                             "unit": "GB/s" if k == "upsample" else "TFLOP/s"} for k, v in prof.items()},
            "whole_pass_tflops": value * MFLOP_PER_SAMPLE[preset] * 1e6 / 1e12 / world}
    if fused:
        roof["note"] = ("gate_gemm family = the fused layer kernel (gate + res|skip FLOPs, HBM-side epilogue I/O included; the last layer's launch "
                        "also carries the WaveNet tail = final 1x1 + zero conv + affine, whose FLOPs are counted here); final_conv family = the "
                        "separate fused tail kernel where the last layer is not fused.  With the 1x1 epilogue I/O switched off the layer kernel "
                        "runs at the power-capped tensor rate (profiles/r2_ncu_fused.md)")
    tp = os.path.join(ROOT, "profiles", "r2_layer_traffic.json" if fused else "r1_gate_traffic.json")
    if mixed and workload == "c3" and os.path.exists(tp):
        tj = json.load(open(tp))
        # DRAM bytes of ONE ncu --set full capture of this kernel (its block-0 launch), next to that launch's algorithmic bytes:
        # a committed capture, NOT re-measured by this run
        roof["traffic"] = tj["traffic_bytes_per_launch"]
        roof["traffic_source"] = "static ncu capture (%s): one block-0 launch on one GPU" % tj["source"]
        roof["traffic_detail"] = {k: tj[k] for k in ("launch", "source", "algorithmic_bytes_per_launch", "algorithmic_flop_per_launch")}
    ups = prof["upsample"]
    if ups[0] > 0:
        roof["upsample_hbm_frac"] = ups[2] / (ups[0] * 1e-3) / 1e9 / pk["hbm_gbs"]
    bytes_in = a_pin.numel() * 4 + c_pin.numel() * 4
    bytes_out = out_pin.numel() * 4 if direction == "reverse" else 8
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": DT_NAME[dtype], "data": "synthetic", "config": config, "roofline": roof, "clocks": clk,
            "e2e": {"value": samples_per_step * steps / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": bytes_in,
                    "d2h_bytes_per_step": bytes_out, "api": "FloWaveNet.reverse_host -> fwn_reverse_host (pinned host buffers)"},
            "gpu_launches": int(launches), "xrt_at_22050": value / 22050.0, "xrt_at_native_rate": value / hp_ref.sample_rate,
            "xrt_per_gpu_at_22050": value / 22050.0 / world}
    if sharded:
        line["halo_bytes_received_per_pass"] = halo_bytes
    if with_cpu and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(preset, hp_kw, direction, B, n_frames)
    del net
    torch.cuda.empty_cache()
    return line


def brief(line):
    """What a secondary workload contributes to the headline line."""
    if line is None:
        return None
    keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "clocks", "gpu_launches", "e2e",
            "xrt_at_22050", "halo_bytes_received_per_pass", "allreduce_bytes_per_step")
    out = {k: line[k] for k in keep if k in line}
    out["workload"] = line["config"]["workload"]
    out["parallelism"] = line["config"]["parallelism"]
    r = line["roofline"]
    out["roofline"] = {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "share_of_step", "phases_ms_per_step",
                                         "whole_pass_tflops") if k in r}
    return out


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
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + sorted(TRAIN_WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="default invocation: skip the c4s / c5 legs")
    ap.add_argument("--dtype", default=None, choices=["bfloat16", "float16"], help="operand type of the mixed-precision workloads")
    ap.add_argument("--train-dtype", default="bfloat16", choices=["bfloat16", "float32"],
                    help="c5: compute dtype of the training step (BASELINE config 5 is bf16; float32 = the fp32-accurate parity mode)")
    ap.add_argument("--split-terms", type=int, default=3, choices=[3, 6], help="c5 in float32: bf16 products per fp32 product in the training GEMMs")
    args = ap.parse_args()
    dist = Dist()
    default_run = args.workload is None
    workload = args.workload or "c3"
    try:
        if workload in TRAIN_WORKLOADS:
            line = run_train(args, dist, workload)
        else:
            line = run_pass(args, dist, workload)
        if default_run and args.impl == "ours" and not args.no_secondary:
            # the workloads that communicate, at the same N, inside the same driver-run record
            sec = {}
            ssteps = min(args.steps, 10)
            saved = args.steps
            args.steps = ssteps
            for name in ("c4s", "c5"):
                try:
                    l2 = run_train(args, dist, name, with_cpu=False) if name in TRAIN_WORKLOADS else run_pass(args, dist, name, with_cpu=False)
                    if dist.rank == 0:
                        sec[name] = brief(l2)
                except Exception as e:  # a failed secondary leg must not void the headline
                    if dist.rank == 0:
                        sec[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            args.steps = saved
            if line is not None:
                line["secondary"] = sec
        if line is not None:
            emit(line)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
