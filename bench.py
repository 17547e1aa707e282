#!/usr/bin/env python
"""bench.py - generated clips/sec (50-step VSampler) for SyncFusion's diffusion sampling path on B200.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU implementation (oracle port)

A "step" is ONE pass of the hot path over one batch of synthetic input: a full ``model.sample()`` call - the
50-step v-diffusion loop of exp/train_diffusion_gh.yaml shapes (L = 262144 @ 48 kHz) - for B clips per GPU.
Workload = BASELINE.json configs[1]: random-init UNet1d of exp/model/diffusion.yaml, batch 16 per GPU, 50-step
VSampler, bf16 operands (fp32 residual stream), random unit-norm 512-d "CLAP" embeddings, synthetic onset pyramid.
Multi-GPU: clips shard by rank with no data-path collective (weak scaling); the one NCCL all-gather of the
finished waveforms is inside the timed region.

Rank 0 prints ONE JSON line.  `value` = whole-job clips/s with inputs resident in HBM; `e2e` = the same through the
public API with pinned HOST inputs (H2D + D2H inside the timed region); `roofline` = the dominant kernel
(sk_kernel: streaming-K fused tcgen05 implicit GEMM) - algorithmic FLOPs of its launches / their CUDA-event durations,
measured live on a profiled replica of the step; `cpu_baseline` = the oracle on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_EVAL_GFLOP = {"nearest": 192.8, "transpose": 185.6}   # algorithmic GFLOP / clip / U-Net evaluation at L = 2^18


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                        hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (NVML; nvidia-smi is not in this image)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v:
                    names[v] = k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and bin(bit).count("1") == 1:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        reasons = sorted(r for r in self.reasons if r not in ("None", "GpuIdle", "ApplicationsClocksSetting"))
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons}


def cpu_reference_clips_per_s(sample_steps, scale, steps_timed, warm, L, threads):
    """Oracle (CPU restatement of audio_diffusion_pytorch / a_unet) on the host cores: 1 clip, `steps_timed` sampler
    steps of the `sample_steps`-step schedule, extrapolated to the full loop."""
    from oracle import DiffusionModel, UNetConfig
    from tests.util import make_encoder, make_inputs
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    om = DiffusionModel(UNetConfig()).eval()
    x, ch, e = make_inputs(om.net.cfg, 1, L, encoder=make_encoder(om.net.cfg))
    sig = torch.linspace(1.0, 0.0, sample_steps + 1)
    times = []
    with torch.no_grad():
        for i in range(warm + steps_timed):
            t0 = time.perf_counter()
            om.net(x, sig[i % sample_steps].reshape(1), embedding=e, embedding_scale=scale, channels=ch)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    per_step = sum(times) / len(times)
    return 1.0 / (per_step * sample_steps), per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    K, W = max(args.steps, 1), max(args.warmup, 0)
    t0 = time.perf_counter()
    cps, per_step = cpu_reference_clips_per_s(args.sample_steps, args.scale, K, min(W, 1), args.length, threads)
    line = {
        "impl": "reference", "metric": "generated clips/sec (50-step VSampler)", "value": cps, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": per_step * args.sample_steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": f"oracle port (reference packages not installable offline), 1 clip, L={args.length}, "
                                   f"{K} of {args.sample_steps} sampler steps timed, extrapolated to the full loop"},
        "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {"workload": "BASELINE.json configs[1]: SyncFusion diffusion UNet1d (exp/model/diffusion.yaml), random-init, "
                        f"batch {batch} per GPU, {args.sample_steps}-step VSampler",
            "batch_per_gpu": batch, "length": args.length, "sample_steps": args.sample_steps,
            "embedding_scale": args.scale, "evals_per_step": 2 if args.scale != 1.0 else 1,
            "precision": args.precision, "upsample_mode": args.upsample_mode,
            "l2": "inputs_exceed_l2 (activations per evaluation >> 126 MB)", "parallelism": f"dp{args.gpus} (clips sharded)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--length", type=int, default=262144)
    ap.add_argument("--sample-steps", type=int, default=50)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--upsample-mode", default="nearest", choices=["nearest", "transpose"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import syncfusion_b200 as sf
    from syncfusion_b200.synth import random_state_dict, synthetic_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the sampling path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # connect every NCCL transport inside init_process_group instead of lazily inside the first collectives: no
        # peer / VMM mapping work is left for the time the sampling kernels run (open issue)
        os.environ.setdefault("NCCL_RUNTIME_CONNECT", "0")
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    B, L, NS = args.batch, args.length, args.sample_steps
    peaks = measured_peaks()

    # random-init weights of the named architecture (seed 0) and synthetic inputs (seed 12345 + rank)
    cfg = sf.UNetConfig(precision=args.precision, upsample_mode=args.upsample_mode)
    model = sf.DiffusionModel(cfg, dev)
    model.load_state_dict(random_state_dict(cfg, seed=0))
    x_h, ch_h, e_h = synthetic_inputs(cfg, B, L, seed=12345 + rank)
    x_p, e_p = x_h.pin_memory(), e_h.pin_memory()
    ch_p = [c.pin_memory() for c in ch_h]
    x_d, e_d = x_p.to(dev), e_p.to(dev)
    ch_d = [c.to(dev) for c in ch_p]
    out_host = torch.empty(B, 1, L).pin_memory()
    if world > 1:
        # NCCL sets its transports up lazily inside the first collective of each kind / size (peer mappings, channel
        # buffers).  Do that here, on an idle GPU and with the timed loop's exact collectives, so no sampling kernel ever
        # runs while a rank is still mapping memory (open issue).
        warm = torch.zeros(B, 1, L, device=dev)
        for _ in range(2):
            sf.gather_waveforms(warm, B * world)
            dist.barrier()
            t_ = torch.zeros(1, device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            torch.cuda.synchronize()
        del warm

    def step_resident():
        out = model.sample(x_noisy=x_d, num_steps=NS, channels=ch_d, embedding=e_d, embedding_scale=args.scale)
        return sf.gather_waveforms(out, B * world) if world > 1 else out

    def step_e2e():
        xd = x_p.to(dev, non_blocking=True)
        ed = e_p.to(dev, non_blocking=True)
        cd = [c.to(dev, non_blocking=True) for c in ch_p]
        out = model.sample(x_noisy=xd, num_steps=NS, channels=cd, embedding=ed, embedding_scale=args.scale)
        out_host.copy_(out, non_blocking=True)
        return out

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(k):
            fn()
        ev1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with ClockSampler(local) as clk:
        ms_total = timed(step_resident, K, W)
    launches = model.net.last_launch_count * K
    clips = B * world * K
    value = clips / (ms_total / 1e3)
    e2e = None
    if not args.no_e2e:
        ms_e2e = timed(step_e2e, K, 1)
        h2d = x_p.numel() * 4 + e_p.numel() * 4 + sum(c.numel() * 4 for c in ch_p)
        e2e = {"value": clips / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_host.numel() * 4}

    # roofline of the dominant kernel (gemm_tc_kernel) from a profiled replica of the step
    model.net.profile(True)
    model.sample(x_noisy=x_d, num_steps=2, channels=ch_d, embedding=e_d, embedding_scale=args.scale)
    torch.cuda.synchronize()
    rep = model.net.profile_report()
    model.net.profile(False)
    by = {}
    for r in rep:
        k = r["kind"]
        a = by.setdefault(k, dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
        a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]; a["n"] += 1
    eval_ms = sum(a["ms"] for a in by.values())
    # dominant kernel: the streaming-K fused tcgen05 GEMM (sk_kernel; fp32 mode: gemm_tc_kernel) - every conv / inject /
    # projection / down / up of depths 3-7, i.e. ~95 % of the algorithmic FLOPs
    dom = "sk" if "sk" in by else "gemm"
    g = by.get(dom, dict(ms=1e-9, flops=0, n=1))
    ach = g["flops"] / (g["ms"] / 1e3) / 1e12
    peak = peaks["bf16_sustained"] if args.precision == "bf16" else peaks["bf16_sustained"] / 2
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_sk_traffic.json")       # dram bytes of the same launches from `ncu --set full`
    if os.path.exists(tpath) and args.precision == "bf16" and B == 16 and L == 262144:
        try:
            traffic = json.load(open(tpath))["dram_bytes_per_launch_avg"]
        except Exception:
            traffic = None
    roofline = {"kernel": "sk_kernel (streaming-K fused tcgen05 implicit GEMM: GN/LN prologue, conv3 / inject / qkv / out / down / up)"
                          if dom == "sk" else "gemm_tc_kernel (tcgen05 implicit GEMM)", "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "algorithmic_flops_per_launch_avg": g["flops"] / max(g["n"], 1), "launches_per_eval": g["n"],
                "avg_launch_ms": g["ms"] / max(g["n"], 1), "share_of_eval_time": g["ms"] / eval_ms,
                "per_kernel": {k: {"ms_per_eval": round(a["ms"], 4), "launches": a["n"],
                                   "tflops": round(a["flops"] / (a["ms"] / 1e3) / 1e12, 2) if a["ms"] > 0 else None,
                                   "gbs": round(a["bytes"] / (a["ms"] / 1e3) / 1e9, 1) if a["ms"] > 0 else None,
                                   "share": round(a["ms"] / eval_ms, 4)} for k, a in sorted(by.items())},
                "hbm_peak_gbs": peaks["hbm"],
                "hbm_bound_kernel": {"kernel": "rk_kernel (resident-weight fused items, depths 1-2)",
                                     "achieved_gbs": (by["rk"]["bytes"] / (by["rk"]["ms"] / 1e3) / 1e9) if "rk" in by else None,
                                     "frac_of_measured_hbm": (by["rk"]["bytes"] / (by["rk"]["ms"] / 1e3) / 1e9 / peaks["hbm"]) if "rk" in by else None}}
    evals = 2 if args.scale != 1.0 else 1
    f_eval = F_EVAL_GFLOP[args.upsample_mode] * (L / 262144.0) * 1e9
    roofline["end_to_end_frac_of_bf16_peak"] = value / world * NS * evals * f_eval / (peaks["bf16_sustained"] * 1e12)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cps, per = cpu_reference_clips_per_s(NS, args.scale, 2, 1, L, threads)
        cpu = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
               "sample": f"oracle port, 1 clip, L={L}, 2 of {NS} sampler steps timed ({per:.2f} s/step), extrapolated to the full loop"}

    if rank == 0:
        line = {"metric": "generated clips/sec (50-step VSampler)", "value": value, "unit": "clips/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "tf32", "data": "synthetic",
                "config": workload_config(args, B), "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
