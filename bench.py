#!/usr/bin/env python
"""bench.py -- Robust U-Net hot-path benchmark (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference arithmetic on the host cores

Own arm.  A step is one pass of the training hot path over one synthetic batch: forward, BCE loss (+ confusion
counts), backward, (N > 1: bucketed gradient all-reduce overlapped with the backward) and the Adam update of
Main_Final.py:552,573-582.  Workload at every N: BASELINE.json configs[1] per GPU -- bf16 storage / fp32 accumulate,
batch 64 at 256x256, 3 channels (weak scaling).  `value` is images/s with the batch resident in HBM; `e2e` is the same
loop fed from pinned host memory (H2D of images + masks and D2H of the loss inside the timed region) through the public
nn.Module API.  `--workload infer` times config 4 (eval forward + thresholded counts) instead.

Reference arm.  The oracle port of the reference (oracle/robust_unet_ref.py, fp32 torch CPU ops = the reference's own
CPU path) on all host threads, same metric/unit, on a bounded sample of the workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRAIN_GFLOP_PER_IMG_256 = 323.6      # SURVEY.md §2.2 / BASELINE.md: fwd + dgrad + wgrad conv FLOPs per image at 256x256
FWD_GFLOP_PER_IMG_256 = 107.9


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs.  NVML is initialised once, up
    front (nvmlInit inside the timed region stalls kernel launches for ~100 ms)."""
    _nv = None
    _handles = {}

    @classmethod
    def prepare(cls, index):
        try:
            import pynvml as nv
            if cls._nv is None:
                nv.nvmlInit()
                cls._nv = nv
            if index not in cls._handles:
                cls._handles[index] = nv.nvmlDeviceGetHandleByIndex(index)
        except Exception as e:
            cls._nv = None
            cls._err = f"nvml_unavailable:{type(e).__name__}"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.recording = False     # the thread polls from before the warm-up (first NVML queries are slow and hold a
                                   # driver lock); samples and throttle reasons count only inside the timed region

    def run(self):
        nv = ClockSampler._nv
        if nv is None:
            self.reasons.add(getattr(ClockSampler, "_err", "nvml_unavailable"))
            return
        try:
            h = ClockSampler._handles[self.index]
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.recording:
                    self.samples.append(mhz)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:   # NVML trouble: report that instead of failing the bench
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ own arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import rbunet
    from rbunet import _lib
    from oracle import robust_unet_ref as R          # synthetic input generator + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().rbu_device_check(), "rbu_device_check")
    ClockSampler.prepare(local)

    infer = args.workload == "infer"
    unet = args.workload == "unet"          # the plain 2-class U-Net of train_water_segmentation.py
    B = args.batch or (32 if infer else 64)
    S = args.size or (1024 if infer else 256)
    nc = args.channels
    torch.manual_seed(0)
    model = (rbunet.UNet(nc, 2) if unet else rbunet.RobustUNet(nc, 1, 64)).to(dev)
    crit = rbunet.CrossEntropyArgmaxLoss() if unet else rbunet.RobustBCEDiceLoss()
    net = model
    if infer:
        model.eval()
    else:
        model.train()
        if world > 1 and not unet:
            net = rbunet.DataParallel(model)
        opt = rbunet.FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)      # = torch.optim.Adam of Main_Final.py:552
    x_cpu, y_cpu = R.synthetic_inputs(B, nc, S, S, seed=123 + rank, blobby=True)
    if unet:
        y_cpu = y_cpu[:, 0].long()           # class-index masks, nn.CrossEntropyLoss style
    x_pin, y_pin = x_cpu.pin_memory(), y_cpu.pin_memory()
    x_dev, y_dev = x_cpu.to(dev), y_cpu.to(dev)
    # e2e: double-buffered input staging -- the H2D copy of step i+1 runs on a copy stream while step i computes (what a
    # pin_memory DataLoader + prefetcher does); every step's inputs are copied from pinned host memory inside the
    # timed region and every step ends with a D2H read of its result
    stage = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {"i": 0, "ev": None}

    def stage_next():
        xs, ys = stage[staged["i"] & 1]
        copy_stream.wait_stream(torch.cuda.current_stream(dev))      # the buffer's previous consumer has been enqueued
        with torch.cuda.stream(copy_stream):
            xs.copy_(x_pin, non_blocking=True)
            ys.copy_(y_pin, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["ev"] = ev
        return xs, ys

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    flush.zero_()      # first use loads torch's fill-kernel module (lazy CUDA module loading: 200-400 ms); the inference
                       # workload has no other fill before the timed region and its first timed step used to absorb that

    def step(x, y):
        if infer:
            with torch.no_grad():
                p = net(x)
                counts = rbunet.confusion_counts(p, y)
            return counts
        opt.zero_grad(set_to_none=True)
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        return loss

    def step_e2e():
        if staged["ev"] is None:
            staged["cur"] = stage_next()                              # first step: nothing to overlap with
        xs, ys = staged["cur"]
        torch.cuda.current_stream(dev).wait_event(staged["ev"])
        staged["i"] += 1
        staged["cur"] = stage_next()                                  # next step's inputs, behind this step's compute
        out = step(xs, ys)
        return out.cpu() if infer else out.item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        sampler = ClockSampler(local)
        sampler.start()
        for _ in range(warmup):
            fn()
        barrier()
        sampler.recording = True
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        marks = []
        for _ in range(steps):
            if not args.no_flush:
                flush.zero_()               # L2 flush between timed iterations
            fn()
            m = torch.cuda.Event(enable_timing=True)
            m.record()
            marks.append(m)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prev, per = e0, []
        for m in marks:                     # per-step durations (diagnostic: outliers show up here, not in the mean)
            per.append(prev.elapsed_time(m))
            prev = m
        timed.per_step = [round(v, 2) for v in per]
        launches = _lib.launch_count() - l0
        clocks = sampler.result()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, clocks

    ms, launches, clocks = timed(lambda: step(x_dev, y_dev), args.steps, args.warmup)
    per_step_value = list(timed.per_step)
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    # host time to ENQUEUE one step (no synchronisation inside): how close the launching thread is to being the bottleneck
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(x_dev, y_dev)
    host_enqueue_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()

    # per-kernel-class device times of one extra step (CUDA events around every C-ABI call on the launching stream)
    # The weight-gradient side stream is switched off for this pass: kernels running concurrently slow each other down and
    # the per-kernel times (the roofline numerators) would be those of the mix, not of the kernel.
    prof = _lib.Profiler()
    eng = getattr(model, "engine", None)
    was_overlap = getattr(eng, "overlap_wgrad", None)
    if was_overlap:
        eng.overlap_wgrad = False
    _lib.PROFILER = prof
    step(x_dev, y_dev)
    _lib.PROFILER = None
    if was_overlap:
        eng.overlap_wgrad = True
    summ = prof.summary()
    if args.detail and rank == 0:
        with open(args.detail, "w") as f:
            for lab, ms_, fl, nb_ in prof.detail():
                f.write(f"{ms_:8.3f} ms  {fl / ms_ / 1e9 if fl and ms_ else 0:8.1f} TF/s  {lab}\n")
    peaks = load_peaks()
    gemm = {k: v for k, v in summ.items() if v["flops"] > 0}
    gemm_ms = sum(v["ms"] for v in gemm.values())
    gemm_flops = sum(v["flops"] for v in gemm.values())
    lib_ms = sum(v["ms"] for v in summ.values())
    top = max(gemm.items(), key=lambda kv: kv[1]["ms"])
    peak_tf = peaks["bf16_tflops_sustained"]
    roof = {"bound": "tensor", "kernel": top[0], "achieved": top[1]["flops"] / top[1]["ms"] / 1e9 if top[1]["ms"] else None,
            "peak": peak_tf, "unit": "TFLOP/s", "frac": None, "traffic": None, "peak_source": peaks["source"] + " (sustained)",
            "launches_per_step": top[1]["calls"], "avg_launch_ms": top[1]["ms"] / max(1, top[1]["calls"]),
            "share_of_step": top[1]["ms"] / lib_ms if lib_ms else None,
            "all_gemm_tflops": gemm_flops / gemm_ms / 1e9 if gemm_ms else None}
    if roof["achieved"]:
        roof["frac"] = roof["achieved"] / peak_tf
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture (dram__bytes_read/write.sum) of the same step
    # (tools/ncu_traffic.py -> profiles/traffic.json: {class: {"dram_bytes_per_launch": ..., "launches": ..., "source": ...}})
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f).get(top[0])
        if tr and not infer and not unet and B == 64 and S == 256:
            roof["traffic"] = tr["dram_bytes_per_launch"]
            roof["traffic_source"] = tr["source"]
            roof["algorithmic_bytes_per_launch"] = top[1]["bytes"] / max(1, top[1]["calls"]) if top[1].get("bytes") else None
    except (OSError, ValueError):
        pass
    classes = {k: {"calls": v["calls"], "ms": round(v["ms"], 3), "share": round(v["ms"] / lib_ms, 4) if lib_ms else None,
                   **({"tflops": round(v["flops"] / v["ms"] / 1e9, 1),
                       "tensor_frac": round(v["flops"] / v["ms"] / 1e9 / peak_tf, 3)} if v["flops"] and v["ms"] else {}),
                   **({"gbps": round(v["bytes"] / v["ms"] / 1e6, 0),
                       "hbm_frac": round(v["bytes"] / v["ms"] / 1e6 / peaks["hbm_gbs"], 3)} if v["bytes"] and v["ms"] else {})}
               for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}
    ew_ms = sum(v["ms"] for v in summ.values() if v["bytes"] and not v["flops"])
    ew_bytes = sum(v["bytes"] for v in summ.values() if v["bytes"] and not v["flops"])

    out = None
    if rank == 0:
        imgs = B * world * args.steps
        gflop_img = (FWD_GFLOP_PER_IMG_256 if infer else (289.2 if unet else TRAIN_GFLOP_PER_IMG_256)) * (S * S) / (256 * 256)
        value = imgs / (ms / 1e3)
        h2d = x_pin.numel() * 4 + y_pin.numel() * y_pin.element_size()
        d2h = (B * 4 * 8) if infer else 4
        out = {"metric": "infer_images_per_sec" if infer else "train_images_per_sec", "value": round(value, 2), "unit": "img/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": (f"Robust U-Net inference {S}x{S} batch {B}/GPU, thresholded counts" if infer else
                                       f"plain 2-class U-Net (train_water_segmentation.py) training step (fwd+CE+bwd+Adam), "
                                       f"batch {B}/GPU at {S}x{S}" if unet else
                                       f"Robust U-Net bf16 training step (fwd+BCE+bwd+Adam), batch {B}/GPU at {S}x{S}, "
                                       f"{nc} channels, base 64"),
                          "global_batch": B * world, "image": [nc, S, S], "parallelism": f"dp{world}",
                          "l2": "256 MiB flush buffer written between timed steps; per-step activations exceed L2",
                          "weights": "reference init (seed 0), random", "optimizer": "rbunet.FusedAdam (= torch.optim.Adam, coupled L2) lr 1e-4 wd 1e-4"},
               "clocks": clocks, "ms_each_step": per_step_value, "host_enqueue_ms_per_step": round(host_enqueue_ms, 2),
               "e2e": {"value": round(imgs / (ms_e2e / 1e3), 2), "unit": "img/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / args.steps, 3)},
               "gpu_launches": int(launches),
               "model_tflops": round(value * gflop_img / 1e3, 1),
               "model_frac_of_sustained_peak": round(value * gflop_img / 1e3 / peak_tf, 4),
               "roofline": roof,
               "hbm_kernels": {"ms": round(ew_ms, 3), "gbps": round(ew_bytes / ew_ms / 1e6, 0) if ew_ms else None,
                               "hbm_frac": round(ew_bytes / ew_ms / 1e6 / peaks["hbm_gbs"], 3) if ew_ms else None,
                               "peak": peaks["hbm_gbs"], "peak_source": peaks["source"],
                               "note": "all bandwidth-bound kernels of one step: algorithmic bytes / summed CUDA-event time"},
               "kernel_classes": classes}
        if world == 1 and not args.no_cpu_baseline and not unet:
            out["cpu_baseline"] = cpu_port(R, infer, nc, S, budget_s=20.0)
        if world == 1 and args.eager_baseline and not infer:
            del model, net
            torch.cuda.empty_cache()
            out["torch_eager_same_gpu"] = eager_cuda_leg(R, nc, B, S, dev)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------ CPU port
def cpu_port(R, infer, nc, S, budget_s, steps=None, warmup=1, batch=None):
    """Times the oracle port (fp32 torch CPU ops, all host threads) on a bounded sample of the workload."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = R.synthetic_state_dict(R.robust_unet_shapes(nc, 1, 64), seed=0)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    params = [sd[n].requires_grad_(not infer) for n in names]
    opt = None if infer else torch.optim.Adam(params, lr=1e-4, weight_decay=1e-4)
    b = batch or 2
    masks = R.synthetic_drop_masks(64, 64, seed=7)

    def one(bs):
        x, y = R.synthetic_inputs(bs, nc, S, S, seed=5, blobby=True)
        t0 = time.perf_counter()
        if infer:
            with torch.no_grad():
                p = R.robust_unet_forward(sd, x, training=False)
            R.confusion_counts(p.numpy(), y.numpy())
        else:
            opt.zero_grad()
            dm = {k: v[:bs] for k, v in masks.items()}
            p = R.robust_unet_forward(sd, x, training=True, drop_masks=dm, new_buffers={})
            loss = R.bce_loss(p, y)
            loss.backward()
            opt.step()
        return time.perf_counter() - t0

    t_probe = one(b)                                   # also the warm-up
    if batch is None:                                  # size the sample to the time budget
        n_steps = steps or 2
        per_img = t_probe / b
        b = int(max(1, min(8, budget_s / max(per_img * (n_steps + warmup), 1e-9))))
    for _ in range(max(0, warmup - 1)):
        one(b)
    n_steps = steps or 2
    times = [one(b) for _ in range(n_steps)]
    total = sum(times)
    return {"value": round(b * n_steps / total, 4), "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"{n_steps} steps of batch {b} at {S}x{S} ({'eval forward + counts' if infer else 'fwd+BCE+bwd+Adam'}), "
                      f"fp32 torch CPU ops ({torch.__version__}), {cores} threads",
            "ms_per_step": round(1e3 * total / n_steps, 1), "batch": b}


def eager_cuda_leg(R, nc, B, S, dev, steps=3):
    """Informational: the reference arithmetic (the oracle's torch ops = what the unmodified reference dispatches)
    executed by eager PyTorch on the SAME GPU -- cuDNN/ATen kernels, autocast(bfloat16) + channels_last and plain fp32
    (SURVEY.md §2.1 calls this the existing-Blackwell-kernel bar).  Not part of the timed arm."""
    import torch
    out = {}
    for mode in ("bf16_autocast_channels_last", "fp32"):
        try:
            sd = {k: v.to(dev) for k, v in R.synthetic_state_dict(R.robust_unet_shapes(nc, 1, 64), seed=0).items()}
            names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
            params = [sd[n].requires_grad_(True) for n in names]
            opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-4, fused=True)
            x, y = R.synthetic_inputs(B, nc, S, S, seed=5, blobby=True)
            x, y = x.to(dev), y.to(dev)
            if mode != "fp32":
                x = x.contiguous(memory_format=torch.channels_last)

            def one():
                opt.zero_grad(set_to_none=True)
                if mode == "fp32":
                    p = R.robust_unet_forward(sd, x, training=True, new_buffers={})
                else:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        z = R.robust_unet_forward(sd, x, training=True, new_buffers={}, return_logits=True)
                    p = torch.sigmoid(z.float())
                loss = R.bce_loss(p, y)
                loss.backward()
                opt.step()

            one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"img_per_s": round(B / ms * 1e3, 1), "ms_per_step": round(ms, 2), "batch": B}
            del sd, params, opt
            torch.cuda.empty_cache()
        except Exception as e:      # e.g. out of memory at this batch: report, do not fail the bench
            out[mode] = {"error": f"{type(e).__name__}: {str(e)[:80]}"}
            torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import robust_unet_ref as R
    infer = args.workload == "infer"
    S = args.size or (1024 if infer else 256)
    nc = args.channels
    steps, warmup = args.steps, max(1, args.warmup)
    budget = 150.0
    r = cpu_port(R, infer, nc, S, budget_s=budget, steps=steps, warmup=warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": "infer_images_per_sec" if infer else "train_images_per_sec", "value": r["value"],
           "unit": "img/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": (f"Robust U-Net inference {S}x{S}" if infer else
                                   f"Robust U-Net training step (fwd+BCE+bwd+Adam) at {S}x{S}, {nc} channels, base 64")
                      + f"; bounded sample: batch {r['batch']} per step on the host CPU", "parallelism": "cpu"},
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": r["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer", "unet"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the 256 MiB L2-flush write between timed steps")
    ap.add_argument("--eager-baseline", action="store_true",
                    help="also time the reference arithmetic through eager PyTorch (cuDNN) on the same GPU (informational)")
    ap.add_argument("--detail", default="", help="write the per-call device times of one profiled step to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if args.workload == "infer" and args.warmup < 6:
            args.warmup = 6        # 1024x1024 activations: the caching allocator needs a few passes to settle
        run_ours(args)


if __name__ == "__main__":
    main()
