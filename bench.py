#!/usr/bin/env python
"""bench.py - generated clips/sec (50-step VSampler) for SyncFusion's diffusion sampling path on B200.

    python bench.py --gpus N --steps K --warmup W                   # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU implementation (oracle port)

A "step" is ONE pass of the hot path over one batch of synthetic input: a full ``model.sample()`` call - the
50-step v-diffusion loop of exp/train_diffusion_gh.yaml shapes (L = 262144 @ 48 kHz) - for B clips per GPU.
Workload = BASELINE.json configs[1]: random-init UNet1d of exp/model/diffusion.yaml, batch 16 per GPU, 50-step
VSampler, bf16 operands (fp32 residual stream), random unit-norm 512-d "CLAP" embeddings, synthetic onset pyramid.
Multi-GPU: clips shard by rank with no data-path collective (weak scaling, 16 clips per GPU); the one NCCL all-gather
of the finished waveforms is inside the timed region.  At N > 1 the line also carries ``config4``: BASELINE.json
configs[3] (256 clips in total sharded over the N GPUs, classifier-free guidance 2.0), one timed call.

Rank 0 prints ONE JSON line, always: every leg after the primary timing is wrapped, so a failing leg adds an entry
to ``errors`` instead of losing the record, a device fault adds the decoded barrier wait log, and a watchdog prints
an error line if the run outlives ``--watchdog`` seconds.  `value` = whole-job clips/s with inputs resident in HBM;
`e2e` = the same through the public API with pinned HOST inputs (H2D + D2H inside the timed region); `roofline` =
the dominant kernel (sk_kernel: streaming-K fused tcgen05 implicit GEMM) - algorithmic FLOPs of its launches / their
CUDA-event durations, measured live on a profiled replica of the step; `cpu_baseline` = the oracle on this box's host
cores on a bounded sample, which also checks the first U-Net evaluation of clip 0 against the GPU (``parity``).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_EVAL_GFLOP = {"nearest": 192.8, "transpose": 185.6}   # algorithmic GFLOP / clip / U-Net evaluation at L = 2^18
METRIC = "generated clips/sec (50-step VSampler)"


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


def csrc_sha() -> str:
    """Hash of the kernel sources: a committed ncu traffic figure is only quoted for the code it was measured on."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "syncfusion_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


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


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def oracle_setup(L, threads, upsample_mode):
    """The oracle (CPU restatement of audio_diffusion_pytorch / a_unet, default init seed 0) and one synthetic clip."""
    from oracle import DiffusionModel, UNetConfig
    from tests.util import make_encoder, make_inputs
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    om = DiffusionModel(UNetConfig(upsample_mode=upsample_mode)).eval()
    x, ch, e = make_inputs(om.net.cfg, 1, L, encoder=make_encoder(om.net.cfg))
    return om, x, ch, e


def oracle_time_evals(om, x, ch, e, sample_steps, scale, n_timed, n_warm):
    """Wall time of single-clip U-Net evaluations (one per sampler step; two network passes each under CFG)."""
    sig = torch.linspace(1.0, 0.0, sample_steps + 1)
    times, v0 = [], None
    with torch.no_grad():
        for i in range(n_warm + n_timed):
            t0 = time.perf_counter()
            v = om.net(x, sig[i % sample_steps].reshape(1), embedding=e, embedding_scale=scale, channels=ch)
            dt = time.perf_counter() - t0
            if i == 0:
                v0 = v
            if i >= n_warm:
                times.append(dt)
    return times, v0


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path = the oracle port (its pip packages are not
    installable offline).  One bench "step" here is a BOUNDED SAMPLE of the workload: one U-Net
    evaluation of one clip (1 of the `sample_steps` sampler steps of 1 of the clips); `ms_per_step` is the time really
    measured per such step, `value` extrapolates it to clips/s of the full 50-step loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    K, W = max(args.steps, 1), max(args.warmup, 0)
    t0 = time.perf_counter()
    om, x, ch, e = oracle_setup(args.length, threads, args.upsample_mode)
    times, _ = oracle_time_evals(om, x, ch, e, args.sample_steps, args.scale, K, W)
    per_eval = sum(times) / len(times)
    cps = 1.0 / (per_eval * args.sample_steps)
    cfgd = workload_config(args, args.batch)
    cfgd["precision"] = "fp32 (CPU)"
    cfgd["step_definition"] = "1 clip x 1 sampler step (one U-Net evaluation) = 1 / (batch x sample_steps) of the GPU arm's step"
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": per_eval * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfgd,
        "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": f"oracle port (reference packages not installable offline), 1 clip, L={args.length}, "
                                   f"{K} of {args.sample_steps} sampler steps timed ({per_eval:.3f} s each), "
                                   f"value = 1 / (per-step time x {args.sample_steps})"},
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


# ------------------------------------------------------------------------------------------------ GPU arm
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
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--watchdog", type=float, default=1500.0, help="seconds after which an error line is printed and the run aborts")
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
    K, W = max(args.steps, 1), max(args.warmup, 0)
    B, L, NS = args.batch, args.length, args.sample_steps
    peaks = measured_peaks()
    line = {"metric": METRIC, "value": None, "unit": "clips/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "tf32", "data": "synthetic", "config": workload_config(args, B),
            "clocks": None, "e2e": None, "gpu_launches": 0, "roofline": None, "cpu_baseline": None}
    errors = {}
    state = {"model": None, "printed": False}
    lock = threading.Lock()

    def emit(rc=None):
        """Print the one JSON line (rank 0, once); with rc: exit right away (no interpreter teardown on a dead context)."""
        with lock:
            if rank == 0 and not state["printed"]:
                if errors:
                    line["errors"] = errors
                    m = state["model"]
                    try:
                        wl = m.net.wait_log() if m is not None else ""
                    except Exception:   # noqa: BLE001
                        wl = ""
                    if wl:
                        line["wait_log"] = wl
                print(json.dumps(line), flush=True)
                state["printed"] = True
        if rc is not None:
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(rc)

    def on_watchdog():
        errors["watchdog"] = f"run exceeded {args.watchdog:.0f} s"
        emit(4)

    dog = threading.Timer(args.watchdog, on_watchdog)
    dog.daemon = True
    dog.start()

    try:
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        # random-init weights of the named architecture (seed 0) and synthetic inputs (seed 12345 + rank)
        cfg = sf.UNetConfig(precision=args.precision, upsample_mode=args.upsample_mode)
        model = sf.DiffusionModel(cfg, dev)
        state["model"] = model
        model.load_state_dict(random_state_dict(cfg, seed=0))
        x_h, ch_h, e_h = synthetic_inputs(cfg, B, L, seed=12345 + rank)
        x_p, e_p = x_h.pin_memory(), e_h.pin_memory()
        ch_p = [c.pin_memory() for c in ch_h]
        x_d, e_d = x_p.to(dev), e_p.to(dev)
        ch_d = [c.to(dev) for c in ch_p]
        out_host = torch.empty(B, 1, L).pin_memory()

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

        # ---- primary: resident inputs, K back-to-back sample() calls, no host sync in between
        with ClockSampler(local) as clk:
            ms_total = timed(step_resident, K, W)
        clips = B * world * K
        value = clips / (ms_total / 1e3)
        evals = 2 if args.scale != 1.0 else 1
        f_eval = F_EVAL_GFLOP[args.upsample_mode] * (L / 262144.0) * 1e9
        line.update(value=value, ms_per_step=ms_total / K, clocks=clk.summary(),
                    gpu_launches=int(model.net.last_launch_count * K))
    except BaseException as ex:   # noqa: BLE001  - the primary timing failed: still one parsable line, non-zero exit
        errors["primary"] = f"{type(ex).__name__}: {str(ex)[:1500]}"
        traceback.print_exc()
        emit(1)
        return

    # ---- enrichment legs: each may fail without losing the record
    def leg(name, fn):
        try:
            fn()
        except BaseException as ex:   # noqa: BLE001
            errors[name] = f"{type(ex).__name__}: {str(ex)[:1500]}"
            traceback.print_exc()

    def leg_e2e():
        ms_e2e = timed(step_e2e, K, 1)
        h2d = x_p.numel() * 4 + e_p.numel() * 4 + sum(c.numel() * 4 for c in ch_p)
        line["e2e"] = {"value": clips / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": out_host.numel() * 4}

    def leg_roofline():
        # roofline of the dominant kernel from a profiled replica of the step (CUDA events around every launch)
        model.net.profile(True)
        model.sample(x_noisy=x_d, num_steps=2, channels=ch_d, embedding=e_d, embedding_scale=args.scale)
        torch.cuda.synchronize()
        rep = model.net.profile_report()
        model.net.profile(False)
        by = {}
        for r in rep:
            a = by.setdefault(r["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
            a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]; a["n"] += 1
        eval_ms = sum(a["ms"] for a in by.values())
        # dominant kernel: the streaming-K fused tcgen05 GEMM (sk_kernel; fp32 mode: gemm_tc_kernel) - every conv / inject /
        # projection / down / up of depths 3-7, i.e. ~95 % of the algorithmic FLOPs
        dom = "sk" if "sk" in by else "gemm"
        g = by.get(dom, dict(ms=1e-9, flops=0, n=1))
        ach = g["flops"] / (g["ms"] / 1e3) / 1e12
        peak = peaks["bf16_sustained"] if args.precision == "bf16" else peaks["bf16_sustained"] / 2
        traffic, traffic_note = None, "not measured in this run (needs ncu)"
        tpath = os.path.join(ROOT, "profiles", "sk_traffic.json")     # dram bytes per launch of the same launches, `ncu --set full`
        if os.path.exists(tpath):
            try:
                t = json.load(open(tpath))
                if t.get("csrc_sha") == csrc_sha() and t.get("batch") == B and t.get("length") == L and t.get("precision") == args.precision:
                    traffic, traffic_note = t["dram_bytes_per_launch_avg"], f"ncu capture of this exact kernel source ({tpath})"
                else:
                    traffic_note = "committed ncu capture is of a different kernel source / shape: not quoted"
            except Exception:   # noqa: BLE001
                pass
        line["roofline"] = {
            "kernel": "sk_kernel (streaming-K fused tcgen05 implicit GEMM: GN/LN prologue, conv3 / inject / qkv / out / down / up)"
                      if dom == "sk" else "gemm_tc_kernel (tcgen05 implicit GEMM)", "bound": "tensor",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "traffic_note": traffic_note,
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
                                 "frac_of_measured_hbm": (by["rk"]["bytes"] / (by["rk"]["ms"] / 1e3) / 1e9 / peaks["hbm"]) if "rk" in by else None},
            "end_to_end_frac_of_bf16_peak": value / world * NS * evals * f_eval / (peaks["bf16_sustained"] * 1e12)}

    def leg_config4():
        # BASELINE.json configs[3]: 256 clips in total over the N GPUs, classifier-free guidance 2.0, + waveform gather
        total = 256
        b4 = total // world
        x4, ch4, e4 = synthetic_inputs(cfg, b4, L, seed=777 + rank)
        x4, e4, ch4 = x4.to(dev), e4.to(dev), [c.to(dev) for c in ch4]

        def step4():
            out = model.sample(x_noisy=x4, num_steps=NS, channels=ch4, embedding=e4, embedding_scale=2.0)
            return sf.gather_waveforms(out, total)

        ms4 = timed(step4, 1, 1)
        line["config4"] = {"workload": "BASELINE.json configs[3]: 256 clips sharded over the GPUs, 50 steps, CFG 2.0, NCCL waveform gather",
                           "batch_total": total, "batch_per_gpu": b4, "embedding_scale": 2.0, "value": total / (ms4 / 1e3),
                           "unit": "clips/s", "ms_per_step": ms4, "steps": 1, "warmup": 1, "scaling": "strong",
                           "frac_of_bf16_peak_end_to_end": total / (ms4 / 1e3) / world * NS * 2 * f_eval / (peaks["bf16_sustained"] * 1e12)}

    def leg_cpu():
        threads = os.cpu_count() or 1
        om, x1, ch1, e1 = oracle_setup(L, threads, args.upsample_mode)
        n_cpu = min(16, NS)             # bounded sample: ~10 s of host work at the bench shape (0.5-0.8 s per evaluation)
        times, v_ref = oracle_time_evals(om, x1, ch1, e1, NS, args.scale, n_cpu, 1)
        per = sum(times) / len(times)
        line["cpu_baseline"] = {"value": 1.0 / (per * NS), "unit": "clips/s", "cores": threads, "kind": "port",
                                "sample": f"oracle port, 1 clip, L={L}, {n_cpu} of {NS} sampler steps timed ({per:.2f} s each), "
                                          f"value = 1 / (per-step time x {NS})"}
        # parity of the shipped CUDA path on the same clip: first U-Net evaluation (sigma = 1) vs the oracle's, same weights
        pm = sf.DiffusionModel(cfg, dev)
        pm.load_state_dict(om.net.state_dict())
        v_gpu = pm.net(x1.to(dev), torch.ones(1, device=dev), embedding=e1.to(dev), embedding_scale=args.scale,
                       channels=[c.to(dev) for c in ch1])
        torch.cuda.synchronize()
        d = (v_gpu.cpu().double() - v_ref.double()).norm() / v_ref.double().norm()
        tol = 2e-2 if args.precision == "bf16" else 1e-3
        line["parity"] = {"check": "first U-Net evaluation of one clip at the full shape, CUDA path vs CPU oracle (same weights)",
                          "rel_l2_v": float(d), "tolerance": tol, "ok": bool(d < tol)}

    if not args.no_e2e:
        leg("e2e", leg_e2e)
    leg("roofline", leg_roofline)
    if world > 1 and not args.no_config4 and 256 % world == 0:
        leg("config4", leg_config4)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        leg("cpu_baseline", leg_cpu)
    dog.cancel()
    emit()
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:   # noqa: BLE001
            pass
    if errors:
        sys.stdout.flush()
        os._exit(0 if line["value"] is not None else 1)


if __name__ == "__main__":
    main()
