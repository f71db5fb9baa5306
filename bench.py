#!/usr/bin/env python
"""Benchmark of the Pangu-Weather 24 h forecast step (PanguModel.forward, 721x1440, batch 1, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode replicas]

A "step" is one full forward of the model over one synthetic ERA5-shaped sample (BASELINE.json
configs[1]).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is defined.

  value     steps/s with inputs resident in HBM, CUDA-event timed, max over ranks, whole job
  e2e       the same through the public API (models.pangu_model.PanguModel.forward) starting from pinned
            HOST buffers, H2D of the 287 MB inputs and D2H of the 287 MB outputs inside the timed region
  roofline  the dominant kernel class (tcgen05 GEMM), FLOPs / CUDA-event time vs the measured bf16 peak
  cpu_baseline  ONE whole forward of the UNMODIFIED reference (baseline/_ref, see baseline/run_reference.py) on all
            host cores, run in a subprocess

--impl reference runs the unmodified reference PanguModel from the git-ignored baseline/_ref (installed by
__graft_entry__.build() where /root/reference exists; it travels with the gpurun snapshot) on all host cores:
1 warm-up + up to 3 timed WHOLE forwards (the cap is stated in config.reference_arm; `steps`/`warmup` in the line
are what was executed, `steps_requested`/`warmup_requested` what was asked).  Fallback if baseline/_ref is
absent: the oracle port's full forward, labelled kind "port".  Nothing is extrapolated.
"""
import argparse
import json
import os
import statistics as pystats
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pangu-pytorch-demo_b200")
sys.path.insert(0, PKG)

METRIC = "24h forecast steps/s at 721x1440 bf16"
UNIT = "steps/s"
FLOPS_TOTAL = 8.421e12            # algorithmic FLOPs of one forward
FLOPS_ATTN_MLP = 8.132e12         # attention + MLP blocks


TRAFFIC_FILE = next((f for f in ("profiles/r2d_ncu_traffic.json", "profiles/r2c_ncu_traffic.json", "profiles/r2b_ncu_traffic.json", "profiles/r2_ncu_traffic.json")
                     if os.path.exists(os.path.join(ROOT, f))), "profiles/r1_ncu_traffic.json")   # newest committed capture


def ncu_traffic(kernel_tag):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (dram__bytes_read.sum + dram__bytes_write.sum of the un-sharded, full-grid launch), or None."""
    p = os.path.join(ROOT, TRAFFIC_FILE)
    want = {"mlp_fused_bf16[C=384]": ("tc::mlp_fused_kernel<384, 0>", 148), "mlp_fused_bf16[C=192]": ("tc::mlp_fused_kernel<192, 0>", 148),
            "attn_proj_mlp_bf16[C=384]": ("tc::mlp_fused_kernel<384, 1>", 148)}
    if kernel_tag not in want or not os.path.exists(p):
        return None
    name, grid = want[kernel_tag]
    for k, lst in json.load(open(p))["kernels"].items():
        if name in k:
            for e in lst:
                if e["grid"] == grid:
                    return e["dram_read_bytes"] + e["dram_write_bytes"]
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# --------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background, killed by exact PID)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": pystats.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (baseline/_ref, see baseline/run_reference.py), whole forwards, all host cores
# --------------------------------------------------------------------------------------------
REF_ARM_MAX_STEPS, REF_ARM_MAX_WARMUP = 3, 1
# `config` is byte-identical in both arms (the driver compares them); what actually happened is in `launch_used`
LAUNCH_NOTE = {"on": "b200 arm: cuda-graph replay, 1 graph launch per step (eager ctypes launches if capture fails, see launch_used)",
               "off": "b200 arm: eager ctypes launches"}
REF_ARM_NOTE = ("CPU arm = the UNMODIFIED reference PanguModel.eval() (baseline/_ref/models/{layers,pangu_model}.py) under "
                "torch.no_grad(), fp32, all host cores, WHOLE forwards timed with perf_counter; --impl reference caps the "
                f"run at {REF_ARM_MAX_WARMUP} warm-up + {REF_ARM_MAX_STEPS} timed forwards whatever --steps/--warmup say "
                "(a forward takes ~20-60 s); the cpu_baseline leg of the default run times ONE forward without warm-up")


def run_cpu_reference(args, as_impl):
    """Times the reference's own CPU implementation of the path.  Preferred: the unmodified reference files in
    baseline/_ref (kind "reference"), run in a subprocess (its package is called `models` like ours).  Fallback
    when baseline/_ref did not travel: the oracle port's full forward (kind "port").  Nothing is extrapolated:
    every reported time is a whole PanguModel forward that was actually executed."""
    steps, warm = ((max(1, min(args.steps, REF_ARM_MAX_STEPS)), min(args.warmup, REF_ARM_MAX_WARMUP)) if as_impl else (1, 0))
    script = os.path.join(ROOT, "baseline", "run_reference.py")
    r = None
    if os.path.exists(script):
        p = subprocess.run([sys.executable, script, "--steps", str(steps), "--warmup", str(warm)],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            r = json.loads(p.stdout.strip().splitlines()[-1])
        except (IndexError, ValueError):
            sys.stderr.write(f"[bench] reference arm failed (rc {p.returncode}): {p.stderr[-2000:]}\n")
            r = None
        if r is not None and "unavailable" in r:
            sys.stderr.write(f"[bench] {r['unavailable']}; falling back to the oracle port\n")
            r = None
    if r is None:
        r = run_cpu_port(steps, warm)
    sec = r["seconds_per_forward"]
    what = ("the UNMODIFIED reference PanguModel.eval() forward (baseline/_ref), " if r["kind"] == "reference" else
            "the oracle port's pangu_forward (baseline/_ref absent on this box), ")
    sample = (what + f"{r['warmup']} warm-up + {r['steps']} timed WHOLE forwards at 721x1440 batch 1 fp32 under no_grad, "
              f"{r['cores']} threads on {r.get('cpu_model', 'unknown CPU')}, torch {r['torch']}; per-forward seconds: "
              + ", ".join(f"{t:.1f}" for t in r["times_s"]))
    return {"value": 1.0 / sec, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample,
            "seconds_per_forward": sec, "steps": r["steps"], "warmup": r["warmup"], "times": r["times_s"]}


def run_cpu_port(steps, warm):
    """Fallback CPU arm: the oracle's restatement of the whole forward (oracle/pangu_oracle.py pangu_forward)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import pangu_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = orc.synth_params(0)
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(1)
    times = []
    with torch.no_grad():
        for i in range(warm + steps):
            t0 = time.perf_counter()
            orc.pangu_forward(params, inp, inp_s, stats, maps, const_h)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return {"kind": "port", "cores": cores, "torch": torch.__version__, "steps": steps, "warmup": warm,
            "times_s": times, "seconds_per_forward": sum(times) / len(times)}


# --------------------------------------------------------------------------------------------
# fine-tune step (BASELINE configs[4]): fwd + loss + bwd + gradient all-reduce + Adam, one sample per GPU
# --------------------------------------------------------------------------------------------
def run_finetune(args, rank, world, dev, torch, dist, orc, PanguModel, ops):
    from pangu_b200.loss import weighted_l1_loss
    torch.manual_seed(1234 + rank)                          # DropPath draws differ per rank, like under DDP
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.to(dev).train().set_compute_dtype("bf16")
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1 + rank)
    g = torch.Generator().manual_seed(100 + rank)
    tgt, tgt_s = torch.randn(inp.shape, generator=g), torch.randn(inp_s.shape, generator=g)
    h_bufs = [t.pin_memory() for t in (inp, inp_s, tgt, tgt_s)]
    d_bufs = [t.to(dev) for t in (inp, inp_s, tgt, tgt_s)]
    stats = tuple(s.to(dev) for s in stats)
    maps, const_h = maps.to(dev), const_h.to(dev)
    # era5_data/config.py:45-46 (Adam, lr 2e-5, weight decay 3e-6), finetune/finetune_fully.py:202
    opt = torch.optim.Adam(model.parameters(), lr=2e-5, weight_decay=3e-6, fused=True)
    reducer, fwd = None, model
    if world > 1:
        if args.dp == "ddp":
            from torch.nn.parallel import DistributedDataParallel as DDP
            # --bucket-mb: DDP's bucket size.  64 MB overlaps the all-reduce with the backward, but the persistent kernels
            # (one CTA per SM, ~220 KB of shared memory each) then queue behind resident NCCL CTAs; a bucket larger than the
            # 1.1 GB of gradients defers the (2-3 ms over NVLink) all-reduce to the end of the backward: no contention.
            fwd = DDP(model, device_ids=[dev.index], bucket_cap_mb=args.bucket_mb, gradient_as_bucket_view=True)
        else:
            from pangu_b200.dist import GradientAllReducer
            reducer = GradientAllReducer(model)
    h_loss = torch.zeros((), dtype=torch.float32).pin_memory()

    def train_step(bufs):
        a, b, t, ts = bufs
        opt.zero_grad(set_to_none=True)
        o, os_ = fwd(a, b, stats, maps, const_h)
        loss = weighted_l1_loss(o, os_, t, ts)              # targets are already normalised (synthetic)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    def step_resident():
        return train_step(d_bufs)

    def step_e2e():
        loss = train_step([h.to(dev, non_blocking=True) for h in h_bufs])
        h_loss.copy_(loss.detach(), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    ops.LAUNCHES = 0
    total_ms = timed(step_resident, args.steps)
    launches = ops.LAUNCHES
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.steps / (total_ms / 1000.0)
    step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    kernels = None
    if not args.no_kernel_times:
        ops.start_kernel_timing()
        step_resident()
        table = ops.stop_kernel_timing()
        kernels = {k: {"calls_per_step": v[0], "ms_per_step": v[1],
                       "tflops": (v[2] / (v[1] / 1000.0) / 1e12) if v[1] > 0 and v[2] > 0 else None,
                       "gbs": (v[3] / (v[1] / 1000.0) / 1e9) if v[1] > 0 and v[3] > 0 else None}
                   for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])}
    peaks = load_peaks()
    if rank == 0:
        ms = total_ms / args.steps
        line = {"metric": "fine-tune steps/s at 721x1440 bf16 (fwd + L1 loss + bwd + gradient all-reduce + Adam), batch 1 per GPU",
                "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic (seeded random-init weights, ERA5-shaped inputs and targets)",
                "config": {"workload": "finetune_fully.py fwd+bwd bf16, batch 1 per GPU, data-parallel NCCL all-reduce (BASELINE.json configs[4])",
                           "parallelism": f"dp{world} ({f'torch DDP, bucket_cap_mb={args.bucket_mb}' if args.dp == 'ddp' else 'GradientAllReducer: flat fp32 buckets, NCCL all-reduce overlapped with the backward'})",
                           "params": 276659936, "grad_bytes": 276659936 * 4, "optimizer": "Adam lr 2e-5 wd 3e-6 (torch fused)",
                           "activations": "saved (no re-computation)" if os.environ.get("PANGU_B200_TRAIN_RECOMPUTE", "0") == "0" else "re-computed per block",
                           "peak_mem_gb": peak_gb, "cache": "activations per step (>= 35 GB) exceed the 126 MB L2; no flush needed"},
                "e2e": {"value": world * args.steps / (e2e_ms / 1000.0), "unit": "steps/s", "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": sum(h.numel() for h in h_bufs) * 4, "d2h_bytes_per_step": 4,
                        "api": "PanguModel(...).train() forward, pangu_b200.loss.weighted_l1_loss, loss.backward(), optimizer.step() from pinned host samples"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": 3 * FLOPS_TOTAL / (ms / 1000.0) / 1e12, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                             "frac": 3 * FLOPS_TOTAL / (ms / 1000.0) / 1e12 / peaks["tf_sust"], "traffic": None,
                             "note": "whole step: 3 x 8.421 TFLOP (fwd + dgrad + wgrad, SURVEY 8d) per sample / step time"},
                "cpu_baseline": None, "kernels": kernels}
        print(json.dumps(line))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        threading.Timer(15.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


# --------------------------------------------------------------------------------------------
# iterative rollout (BASELINE configs[2]): the 24 h model chained 7 times, state resident on the device
# --------------------------------------------------------------------------------------------
def run_rollout(args, rank, world, dev, torch, dist, orc, PanguModel, ops, chain=7):
    from pangu_b200.rollout import Rollout
    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.to(dev).eval().set_compute_dtype(args.dtype)
    inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1 + rank)
    sm, ss, um, us = stats                                   # weatherStatistics_output layout, era5_data/utils_data.py:395-421
    f = lambda t: t.flip(0).permute(1, 3, 0, 2).unsqueeze(-1).contiguous()
    last = (sm.view(1, 4, 1, 1), ss.view(1, 4, 1, 1), f(um), f(us))
    ro = Rollout(model, stats, last, maps, const_h, graph=args.graph == "on")
    h_in = (inp.pin_memory(), inp_s.pin_memory())
    d_in = (inp.to(dev), inp_s.to(dev))
    h_out = [(torch.empty((5, 13, 721, 1440)).pin_memory(), torch.empty((4, 721, 1440)).pin_memory()) for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def chain_resident():
        for _o, _os in ro.run(d_in[0], d_in[1], steps=chain):
            pass

    def chain_e2e():
        a, b = h_in[0].to(dev, non_blocking=True), h_in[1].to(dev, non_blocking=True)
        for k, (o, os_) in enumerate(ro.run(a, b, steps=chain)):        # every lead time goes back to the host
            h_out[k % 2][0].copy_(o.reshape(5, 13, 721, 1440), non_blocking=True)
            h_out[k % 2][1].copy_(os_.reshape(4, 721, 1440), non_blocking=True)

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    chains = max(1, (args.steps + chain - 1) // chain)       # --steps counts forecast steps; whole chains are timed
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    for _ in range(max(1, (args.warmup + chain - 1) // chain)):
        chain_resident()
    ops.LAUNCHES = 0
    total_ms = timed(chain_resident, chains)
    launches = ops.LAUNCHES
    clocks = sampler.stop() if rank == 0 else None
    steps = chains * chain
    chain_e2e()
    e2e_ms = timed(chain_e2e, chains)
    peaks = load_peaks()
    if rank == 0:
        value = world * steps / (total_ms / 1000.0)
        line = {"metric": "24h forecast steps/s at 721x1440 bf16, 7-step iterative rollout", "value": value, "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": args.warmup, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic (seeded random-init weights, ERA5-shaped inputs)",
                "config": {"workload": "Iterative rollout: 24h model chained 7 steps, batch 1, state resident on the device (BASELINE.json configs[2])",
                           "chain": chain, "chains_timed": chains, "parallelism": f"replicas x{world}",
                           "launch": LAUNCH_NOTE[args.graph], "de-normalisation": "fused into the patch-recover scatter (normBackData)",
                           "cache": "activations per step (>= 6 GB) exceed the 126 MB L2; no flush needed"},
                "e2e": {"value": world * steps / (e2e_ms / 1000.0), "unit": UNIT, "ms_per_step": e2e_ms / steps,
                        "h2d_bytes_per_step": (inp.numel() + inp_s.numel()) * 4 // chain, "d2h_bytes_per_step": (inp.numel() + inp_s.numel()) * 4,
                        "api": "pangu_b200.rollout.Rollout.run: initial state from pinned host memory once per chain, every lead time copied back to pinned host memory"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": FLOPS_TOTAL * value / world / 1e12, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                             "frac": FLOPS_TOTAL * value / world / 1e12 / peaks["tf_sust"], "traffic": None,
                             "note": "whole model: 8.421 TFLOP per forecast step / step time (per-kernel figures: default mode)"},
                "cpu_baseline": None}
        print(json.dumps(line))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        threading.Timer(15.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "replicas", "bands", "finetune", "rollout"],
                    help="multi-GPU partition: replicas = one forecast per GPU (weak scaling); bands = ONE forecast "
                         "sharded over latitude bands with NCCL halo exchange (strong scaling, BASELINE configs[3]); "
                         "finetune = BASELINE configs[4]: fwd + loss + bwd + NCCL gradient all-reduce + Adam, one sample per "
                         "GPU (data parallel, weak scaling); rollout = BASELINE configs[2]: the 24 h model chained 7 times on "
                         "the device; auto = bands when N > 1")
    ap.add_argument("--dp", default="ddp", choices=["buckets", "ddp"],
                    help="--mode finetune gradient all-reduce: pangu_b200.dist.GradientAllReducer or torch DDP (what the "
                         "reference uses, finetune/finetune_fully.py:220)")
    ap.add_argument("--bucket-mb", type=int, default=2048,
                    help="--mode finetune --dp ddp: DDP bucket_cap_mb (2048 = one bucket after the backward; 64 = overlapped)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="replay the step from a CUDA graph (pangu_b200.graph.GraphedForward) instead of ~100 "
                         "ctypes launches per step")
    ap.add_argument("--scheme", default="redundant", choices=["redundant", "sendback"],
                    help="halo exchange scheme of --mode bands (pangu_b200/dist.py)")
    ap.add_argument("--e2e-depth", type=int, default=0,
                    help="slots (captured graphs + pinned output buffers) of the streamed e2e leg; 0 = 2")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not bind each rank's host thread to its GPU's NUMA node")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-times", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    mode = args.mode if args.mode != "auto" else ("bands" if world > 1 else "replicas")
    if mode in ("finetune", "rollout") and args.impl == "reference":
        raise SystemExit("--impl reference times the forward (the headline metric); --mode finetune / rollout have no CPU arm")
    if mode == "bands" and world not in (1, 2, 4, 8):
        mode = "replicas"
    par = (f"replicas x{world} (one forecast per GPU, no data-path collective)" if mode == "replicas" else
           f"latitude bands x{world} (ONE forecast sharded over {world} GPUs; NCCL neighbour halo exchange of 3 rows "
           "around each of the 8 shifted-window blocks)")
    config = {"workload": "PanguModel 24h forward full-res bf16, batch 1 (BASELINE.json configs[1]"
                          + ("" if mode == "replicas" else " sharded as configs[3]") + ")",
              "grid": "13x721x1440 upper-air x5 + 721x1440 surface x4", "tokens": "521280@C192 + 131040@C384",
              "blocks": 16, "params": 276659936, "parallelism": par, "mode": mode,
              "cache": "inputs+activations per step (>= 6 GB) exceed the 126 MB L2; no flush needed",
              "reference_arm": REF_ARM_NOTE}

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_cpu_reference(args, as_impl=True)
        config["launch"] = LAUNCH_NOTE[args.graph]
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": r["steps"], "warmup": r["warmup"], "steps_requested": args.steps, "warmup_requested": args.warmup,
                "ms_per_step": 1000.0 * r["seconds_per_forward"],
                "higher_is_better": True, "scaling": "weak" if mode == "replicas" else "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic (seeded random-init weights, ERA5-shaped inputs)",
                "config": config, "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_cpus = None
    if world > 1 and not args.no_numa_bind:
        # one process per GPU: bind each rank to its GPU's NUMA node BEFORE any pinned buffer exists (see prefetch.py)
        from pangu_b200.prefetch import bind_host_thread_to_gpu
        numa_cpus = bind_host_thread_to_gpu(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sys.path.insert(0, os.path.join(ROOT, "oracle"))       # synthetic weights/inputs only (seeded generators)
    import pangu_oracle as orc
    from models.pangu_model import PanguModel
    from pangu_b200 import ops

    if mode == "finetune":
        return run_finetune(args, rank, world, dev, torch, dist, orc, PanguModel, ops)
    if mode == "rollout":
        return run_rollout(args, rank, world, dev, torch, dist, orc, PanguModel, ops)

    model = PanguModel(device="cpu")
    model.load_state_dict(orc.synth_params(seed=0), strict=True)
    model = model.to(dev).eval().set_compute_dtype(args.dtype)
    if mode == "bands":
        # every rank holds its latitude band of ONE sample (same seed on all ranks)
        from pangu_b200.dist import BandedPangu, BandPlan
        inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1)
        plan = BandPlan(world, rank)
        full_inputs = (inp, inp_s, maps, const_h) if world > 1 else None       # for the post-timing band check
        inp, inp_s, maps, const_h = plan.slice_inputs(inp[0], inp_s[0], maps, const_h)
        forward = BandedPangu(model, scheme=args.scheme) if world > 1 else model
        config["halo_scheme"] = args.scheme
        if world == 1:
            inp, inp_s = inp[None], inp_s[None]
    else:
        inp, inp_s, stats, maps, const_h = orc.synth_inputs(seed=1 + rank)
        forward = model
    h_inp, h_inp_s = inp.pin_memory(), inp_s.pin_memory()
    d_inp, d_inp_s = inp.to(dev), inp_s.to(dev)
    stats = tuple(s.to(dev) for s in stats)
    maps, const_h = maps.to(dev), const_h.to(dev)
    lat = inp.shape[-2]
    h_out = torch.empty((1, 5, 13, lat, 1440), dtype=torch.float32).pin_memory()
    h_out_s = torch.empty((1, 4, lat, 1440), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = None
    if args.graph == "on":
        from pangu_b200.graph import GraphedForward
        try:
            graphed = GraphedForward(forward, (d_inp, d_inp_s, stats, maps, const_h))
        except Exception as exc:                                  # noqa: BLE001 -- report and measure the eager path
            sys.stderr.write(f"[bench] CUDA-graph capture failed on rank {rank} ({type(exc).__name__}: {exc}); eager launches\n")
            graphed = None
        if world > 1:                                             # all ranks must take the same path
            flag = torch.tensor([1 if graphed is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                graphed = None
    config["launch"] = LAUNCH_NOTE[args.graph]
    launch_used = "cuda-graph replay (1 graph launch per step)" if graphed is not None else "eager ctypes launches"

    def step_resident():
        if graphed is not None:
            return graphed.replay()
        with torch.no_grad():
            return forward(d_inp, d_inp_s, stats, maps, const_h)

    def step_e2e():
        with torch.no_grad():
            if graphed is not None:
                o, os_ = graphed(h_inp, h_inp_s)                   # H2D into the static buffers + replay
            else:
                a = h_inp.to(dev, non_blocking=True)
                b = h_inp_s.to(dev, non_blocking=True)
                o, os_ = forward(a, b, stats, maps, const_h)
            h_out.copy_(o, non_blocking=True)
            h_out_s.copy_(os_, non_blocking=True)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                    # samples cover warm-up + the timed region (both under load)
    for _ in range(args.warmup):
        step_resident()
    ops.LAUNCHES = 0
    total_ms = timed(step_resident, args.steps)
    launches = ops.LAUNCHES
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    jobs = world if mode == "replicas" else 1              # forecasts completed per step across the job
    value = jobs * args.steps / (total_ms / 1000.0)

    streamer = None
    if graphed is not None:
        # public streaming API: H2D of sample i+1, forward of sample i and D2H of sample i-1 overlap
        # (every sample still pays both of its transfers inside the timed region)
        from pangu_b200.pipeline import StreamedForecaster
        try:
            # (a third slot -- the host a full step ahead of the device on every rank -- measured neutral in band mode at 4 GPUs:
            # profiles/r2_e2e_bands_transfers.md)
            depth = args.e2e_depth if args.e2e_depth > 0 else 2
            streamer = StreamedForecaster(forward, (d_inp, d_inp_s, stats, maps, const_h), depth=depth)
        except Exception as exc:                                  # noqa: BLE001
            sys.stderr.write(f"[bench] StreamedForecaster unavailable on rank {rank} ({type(exc).__name__}: {exc})\n")
        if world > 1:
            flag = torch.tensor([1 if streamer is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                streamer = None

    def e2e_run(steps):
        if streamer is None:
            for _ in range(steps):
                step_e2e()
        else:
            for _ in range(steps):
                streamer.submit(h_inp, h_inp_s)
            streamer.flush()                                      # last results are on the host when timing stops

    e2e_run(2)
    e2e_ms = timed(lambda: e2e_run(args.steps), 1)
    e2e_value = jobs * args.steps / (e2e_ms / 1000.0)
    h2d = (h_inp.numel() + h_inp_s.numel()) * 4            # per rank
    d2h = (h_out.numel() + h_out_s.numel()) * 4

    # the two transfers of one step ALONE (all ranks at the same time, nothing else running): what the host side of this
    # box sustains; explains how much of e2e - value is transfer time that a 2-deep pipeline cannot hide
    def copy_in():
        d_inp.copy_(h_inp.reshape(d_inp.shape), non_blocking=True)
        d_inp_s.copy_(h_inp_s.reshape(d_inp_s.shape), non_blocking=True)

    def copy_out():
        h_out.copy_(step_out[0].reshape(h_out.shape), non_blocking=True)
        h_out_s.copy_(step_out[1].reshape(h_out_s.shape), non_blocking=True)

    step_out = step_resident()
    copy_in(); copy_out()
    h2d_ms = timed(copy_in, 10) / 10
    d2h_ms = timed(copy_out, 10) / 10

    kernels, roofline = None, None
    peaks = load_peaks()
    if not args.no_kernel_times:
        ops.start_kernel_timing()                          # per-kernel CUDA events need the eager launch path
        _g, graphed = graphed, None
        inst_ms = timed(step_resident, args.steps)
        graphed = _g
        table = ops.stop_kernel_timing()               # {op-class: (calls, total_ms, flops, bytes)}
        kernels = {k: {"calls_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps,
                       "tflops": (v[2] / (v[1] / 1000.0) / 1e12) if v[1] > 0 and v[2] > 0 else None,
                       "gbs": (v[3] / (v[1] / 1000.0) / 1e9) if v[1] > 0 and v[3] > 0 else None}
                   for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])}
        gemm = {k: v for k, v in table.items() if k.startswith("gemm") or k.startswith("mlp_fused") or k.startswith("attn_proj_mlp")}
        if gemm:
            dom = max(gemm, key=lambda k: gemm[k][1])
            calls, tms, fl, _ = gemm[dom]
            ach = fl / (tms / 1000.0) / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tf_sust"], "traffic": ncu_traffic(dom) if world == 1 else None,
                        "traffic_source": (TRAFFIC_FILE + " (ncu --set full capture of this kernel at this shape, committed; "
                                           "not re-measured by this run)" if world == 1 and ncu_traffic(dom) else
                                           "null: the committed ncu capture is of the un-sharded launch, not of this band-sized one"
                                           if world > 1 else None),
                        "algorithmic_bytes_per_launch": gemm[dom][3] / calls, "peak_source": peaks["src"] + " (sustained cuBLAS bf16)",
                        "avg_launch_ms": tms / calls, "flops_per_launch": fl / calls,
                        "all_gemm_tflops": sum(v[2] for v in gemm.values()) / (sum(v[1] for v in gemm.values()) / 1000.0) / 1e12,
                        "model_attn_mlp_frac": FLOPS_ATTN_MLP * value / world / 1e12 / peaks["tf_sust"],
                        "instrumented_ms_per_step": inst_ms / args.steps}

    band_check = None
    if mode == "bands" and world > 1:
        # outside the timed regions: every rank also runs the UN-sharded forward and compares its own rows of the
        # banded result (over real NCCL) with it, bit for bit; the verdict is the minimum over the ranks
        with torch.no_grad():
            bo, bos = forward(d_inp, d_inp_s, stats, maps, const_h)
            fo, fos = model(full_inputs[0].to(dev), full_inputs[1].to(dev), stats, full_inputs[2].to(dev), full_inputs[3].to(dev))
        p0, p1 = plan.pix
        same = bool(torch.equal(bo, fo[..., p0:p1, :])) and bool(torch.equal(bos, fos[..., p0:p1, :]))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        band_check = {"bit_identical_to_unsharded": bool(int(flag.item())), "ranks": world,
                      "how": "each rank: torch.equal(own rows of the banded forward over NCCL, same rows of its own un-sharded forward)"}
        del bo, bos, fo, fos

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu_reference(args, as_impl=False)
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak" if mode == "replicas" else "strong", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic (seeded random-init weights, ERA5-shaped inputs)", "config": config,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps,
                        "host_thread_numa_bound_cpus": len(numa_cpus) if numa_cpus else None,
                        "transfers_alone": {"h2d_ms": h2d_ms, "d2h_ms": d2h_ms, "h2d_gbs": h2d / h2d_ms / 1e6, "d2h_gbs": d2h / d2h_ms / 1e6,
                                            "note": "one step's pinned-host copies timed alone, all ranks at once, max over ranks"},
                        "api": ("pangu_b200.pipeline.StreamedForecaster(PanguModel / BandedPangu): pinned host in -> pinned host out, "
                                "transfers of neighbouring samples overlap the forward, %d slots" % streamer.depth if streamer is not None else
                                "models.pangu_model.PanguModel.forward on pinned host inputs, serial H2D / forward / D2H")},
                "gpu_launches": launches, "launch_used": launch_used, "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "band_check": band_check, "kernels": kernels, "tflops_model_per_gpu": FLOPS_TOTAL * value / world / 1e12}
        print(json.dumps(line))
    if world > 1:
        # NCCL teardown while a captured graph still references the communicator can block: release the graph,
        # drain the device, and bound the teardown (the JSON line is already out)
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        threading.Timer(15.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    main()
