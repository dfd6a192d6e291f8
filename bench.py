#!/usr/bin/env python
"""bench.py -- end-to-end 512x512 slices/s of the per-slice segmentation path (RAW u16 -> UNet -> polygons).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32] [--head binary|argmax]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic CT-like slices (BASELINE.json
configs[1]: batch 32 of 512x512, binary UNet, bf16).  Slices are independent, so ranks shard the slices
with no data-path collective ("scaling": "weak": every rank processes its own batch per step); the only
torch.distributed traffic is the barrier and the max-over-ranks of the timing.

Keys of the JSON line (rank 0):
  value / ms_per_step : whole-job slices/s with inputs already resident in HBM, CUDA-event timed
  e2e                 : same metric through the C-ABI call with pinned HOST buffers (H2D + polygons D2H inside)
  roofline            : the dominant kernel instantiation (by time): algorithmic FLOP of its launches / their CUDA-event
                        durations INSIDE the step vs the measured sustained cuBLAS bf16 peak; `alone` = the same launches
                        timed alone vs the burst peak; `forward_pass` = all UNet launches; `by_kernel` = every instantiation
  cpu_baseline        : the oracle port of the reference's CPU pipeline on a bounded sample (rank 0, N=1 only)
  clocks, gpu_launches, p50_ms_per_slice (batch-1 latency)
`--impl reference` times the oracle port (the reference cannot be built here: OpenCV C++ SDK + TensorRT
are absent) on the host cores, on the same config and metric.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="slices per step per GPU (cfg2: 32)")
    ap.add_argument("--head", default="binary", choices=["binary", "argmax"], help="cfg2 is the binary head; argmax = reference's 3-class")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-sample", type=int, default=0, help="slices in the CPU baseline sample (0 = auto, ~10-30 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layer-table", default="", help="write the per-layer roofline table to this JSON file")
    ap.add_argument("--volume-slices", type=int, default=256, help="cfg3: slices of the one-call volume run (0 = skip)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def unet_layer_bytes(names, B, S, n_classes):
    """Algorithmic HBM bytes per UNet layer of the canonical architecture at batch B, S x S input:
    every input map read once, every output map written once (bf16 NHWC; the u8 slice in, the u8 mask out), weights once.
    The skip halves of the concat buffers are counted where they are read (the decoder's first conv)."""
    px = lambda lvl: B * (S >> lvl) * (S >> lvl)
    out = {"enc1a": px(0) * (1 + 64 * 2) + 64 * 9 * 4}
    enc = {1: (64, 64), 2: (64, 128), 3: (128, 256), 4: (256, 512)}
    out["enc1b"] = px(0) * (64 * 2 + 64 * 2) + px(1) * 64 * 2 + 64 * 64 * 9 * 2
    for lvl, name in ((1, "enc2"), (2, "enc3"), (3, "enc4")):
        cin, cout = enc[lvl + 1]
        out[name + "a"] = px(lvl) * (cin + cout) * 2 + cin * cout * 9 * 2
        out[name + "b"] = px(lvl) * (cout + cout) * 2 + px(lvl + 1) * cout * 2 + cout * cout * 9 * 2
    out["bott_a"] = px(4) * (512 + 1024) * 2 + 512 * 1024 * 9 * 2
    out["bott_b"] = px(4) * (1024 + 1024) * 2 + 1024 * 1024 * 9 * 2
    for k, lvl, c in ((4, 3, 512), (3, 2, 256), (2, 1, 128), (1, 0, 64)):      # up_k: level lvl + 1 -> lvl; dec_k at level lvl
        out["up%d" % k] = px(lvl + 1) * 2 * c * 2 + px(lvl) * c * 2 + 2 * c * c * 4 * 2
        out["dec%da" % k] = px(lvl) * (2 * c + c) * 2 + 2 * c * c * 9 * 2
        out["dec%db" % k] = px(lvl) * (c + c) * 2 + c * c * 9 * 2
    out["dec1b_head"] = px(0) * (64 * 2 + 1) + 64 * 64 * 9 * 2
    out.pop("dec1b", None)
    return {n: out.get(n, 0) for n in names}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs"), "bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # the first queries can take ~100 ms: pay for them before the timed region starts
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            self.nv = None
        self.proc = None
        if self.nv is None:   # no NVML binding: let nvidia-smi sample (the recipe's clocks line)
            try:
                import subprocess
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active",
                                              "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            except Exception:
                self.proc = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                out = self.proc.communicate(timeout=5)[0]
            except Exception:
                out = ""
            masks = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            for line in out.splitlines():
                f = [x.strip() for x in line.split(",")]
                try:
                    self.samples.append(float(f[0]))
                    self.max_mhz = int(float(f[1]))
                    bits = int(f[2], 16)
                    self.reasons |= {n for b, n in masks.items() if bits & b}
                except Exception:
                    pass
        elif self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ reference arm
def oracle_runner(n_classes, head, size, with_files=False):
    """The reference's CPU pipeline as restated by the oracle (oracle/pipeline.py + torch-CPU UNet)."""
    import torch
    import medseg_b200 as ms
    from medseg_b200 import weights as W
    from oracle import pipeline as op
    from oracle.unet_torch import load_unet
    import cv2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    net = load_unet(W.make_weights(1234, n_classes), n_classes)

    def run(slices):
        out = []
        for s in slices:
            r = op.process_slice(s, net, head="argmax" if head == "argmax" else "binary", n_classes=n_classes)
            out.append(op.generate_json(r["mapped"], "s", s.shape[1], s.shape[0]) if r["mapped"] else "")
        return out

    def run_files(paths, out_dir):   # the reference as shipped: through PNG / JSON files
        for p in paths:
            op.process_single_image_files(p, size, size, out_dir, net, head="argmax" if head == "argmax" else "binary")
    return (run, cores, run_files) if with_files else (run, cores)


def run_reference(args, rank, world):
    from medseg_b200 import synth
    if rank != 0:
        return
    n_classes = 1 if args.head == "binary" else 3
    run, cores = oracle_runner(n_classes, args.head, args.size)
    vol = synth.ct_volume(args.batch, args.size, args.size)       # the same slices our arm's rank 0 processes per step
    t0 = time.perf_counter()
    for _ in range(max(1, args.warmup)):
        run(vol[:1])
    per_slice = (time.perf_counter() - t0) / max(1, args.warmup)
    # the step is the whole batch when the run then still ends within a few minutes (~4), else a bounded prefix of it
    sample = args.cpu_sample or max(1, min(args.batch, int(240.0 / (max(1, args.steps) * per_slice))))
    vol = vol[:sample]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(vol)
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    line = {"impl": "reference", "metric": "slices_per_sec", "value": value, "unit": "slices/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg2: batch {args.batch} of {args.size}x{args.size} slices, {args.head} UNet -> polygons",
                       "sample_per_step": sample, "same_sample_as_ours": sample == args.batch, "note": "reference CPU pipeline = oracle port (cv2 4.13 + torch-CPU fp32 UNet); "
                       "the reference binary needs OpenCV C++ SDK + TensorRT, neither is installable here"},
            "cpu_baseline": {"value": value, "unit": "slices/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} slice(s) of {args.size}x{args.size} per step x {args.steps} steps, in memory, JSON text included"},
            "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import medseg_b200 as ms
    from medseg_b200 import synth
    from medseg_b200.sharding import max_over_ranks as _max_over_ranks, shard_range

    use_dist = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    cpu_group = None
    if use_dist:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # host-side barrier for the one-process volume run: an NCCL barrier parks a spinning kernel on every waiting GPU,
        # and rank 0's work on those GPUs would then be time-sliced against it (different processes do not share a GPU
        # concurrently)
        cpu_group = dist.new_group(backend="gloo")
    n_classes = 1 if args.head == "binary" else 3
    B, S = args.batch, args.size

    td = tempfile.mkdtemp(prefix=f"medseg_bench_r{rank}_")
    blob = ms.make_weight_blob(os.path.join(td, "unet.msegw"), n_classes=n_classes, seed=1234)
    cfg = {"weights": blob, "max_batch": B, "device": local, "net_h": S, "net_w": S}
    if args.head == "binary":
        cfg["head"] = "binary"
    eng = ms.Engine(cfg)
    # a dedicated (non-default) stream: the library launches on the stream handle it is given and
    # torch.cuda.Event only sees torch's current stream, so both must be this one
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # synthetic input: R distinct batches per rank, slices seeded by global slice index (cfg3 sharding:
    # contiguous block of slices per rank)
    R = 4
    lo, hi = shard_range(world * R * B, world, rank)   # this rank's contiguous block of the synthetic volume
    host = [torch.from_numpy(synth.ct_volume(B, S, S, first_seed=lo + r * B)).pin_memory() for r in range(R)]
    assert hi - lo == R * B
    dev = [h.cuda(non_blocking=True) for h in host]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return _max_over_ranks(x, device="cuda")

    # ---------------- device-resident throughput ("value")
    for i in range(args.warmup):
        eng.process_batch_dev(dev[i % R].data_ptr(), S, S, B, stream)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.process_batch_dev(dev[i % R].data_ptr(), S, S, B, stream, wait=False)    # no host round trip inside a step
    e1.record()
    barrier()
    clocks = sampler.result()
    n_pts, n_cnt = eng.last_counts()                  # validates the last step's header (overflow / trace errors raise)
    launches = eng.launch_count() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    # the same loop for >= 3 s: the 20-step region lasts ~0.2 s and runs nearer burst clocks
    sus_steps = max(args.steps, int(3.2e3 / ms_per_step) + 1)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(sus_steps):
        eng.process_batch_dev(dev[i % R].data_ptr(), S, S, B, stream, wait=False)
    s1.record()
    barrier()
    eng.last_counts()
    sus_ms = max_over_ranks(s0.elapsed_time(s1))
    value_sustained = world * B * sus_steps / (sus_ms / 1e3)

    # ---------------- end to end through the host-buffer C-ABI calls ("e2e"): the double-buffered streaming form
    # (ms_submit_batch_host / ms_wait_batch); every step's H2D (pinned u16 slices) and D2H (polygons) is inside
    polys = None
    hnp = [h.numpy() for h in host]
    # warm both slots: first use allocates a slot's buffers, second use captures its CUDA graph, later uses replay it
    for i in range(max(args.warmup, 6)):
        eng.submit_batch(i % 2, hnp[i % R])
        polys = eng.wait_batch(i % 2)
    polys, _, _ = eng.process_batch(hnp[0])
    barrier()
    x0 = eng.transfer_bytes()
    t0 = time.perf_counter()
    eng.submit_batch(0, hnp[0])
    for i in range(args.steps):
        if i + 1 < args.steps:
            eng.submit_batch((i + 1) % 2, hnp[(i + 1) % R])
        polys = eng.wait_batch(i % 2)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    x1 = eng.transfer_bytes()
    e2e_value = world * B * args.steps / dt
    # bytes the library actually copied per step (counted at its cudaMemcpyAsync calls), whole job like `value`
    h2d = world * (x1[0] - x0[0]) // args.steps
    d2h = world * (x1[1] - x0[1]) // args.steps
    d2h_useful = world * int(polys.n_points * 8 + (polys.n_contours + 1) * 4 + (B + 1) * 4 + 32)
    # the same through the synchronous single call, for reference
    t0 = time.perf_counter()
    for i in range(args.steps):
        polys, _, _ = eng.process_batch(hnp[i % R])
    e2e_sync = world * B * args.steps / max_over_ranks(time.perf_counter() - t0)

    # ---------------- JSON serialisation rate, reported separately: the LabelMe documents of the
    # last batch's polygons through the threaded formatter, all host cores
    names = [f"slice_{i:04d}" for i in range(B)]
    jbuf = np.empty(int(polys.n_points) * 100 + B * 1024, np.uint8)
    ms.polygons_to_json_batch(polys, names, S, S, buf=jbuf)
    t0 = time.perf_counter()
    jreps = 0
    while time.perf_counter() - t0 < 0.5:
        jtxt, joffs = ms.polygons_to_json_batch(polys, names, S, S, buf=jbuf)
        jreps += 1
    jdt = (time.perf_counter() - t0) / jreps
    json_rate = {"json_text_slices_per_s": B / jdt, "threads": os.cpu_count(), "bytes_per_slice": int(joffs[-1]) // B,
                 "api": "ms_polygons_to_json_batch (byte-exact with the reference's nlohmann dump(4))"}

    # ---------------- batch-1 latency (p50 ms/slice), host buffers
    lat = []
    one = host[0].numpy()[:1]
    for i in range(3 + 20):
        t = time.perf_counter()
        eng.process_batch(one)
        if i >= 3:
            lat.append((time.perf_counter() - t) * 1e3)
    p50_sync = float(np.median(lat))
    # the same slice through submit + wait: the kernel chain of a slot is one CUDA graph launch
    lat = []
    for i in range(5 + 30):
        t = time.perf_counter()
        eng.submit_batch(0, one)
        eng.wait_batch(0)
        if i >= 5:
            lat.append((time.perf_counter() - t) * 1e3)
    p50 = float(np.median(lat))

    # ---------------- roofline of the dominant kernel, live: CUDA events around every layer launch INSIDE the step
    # (ms_profile_layers_*: the same ms_process_batch_dev loop as the timed region, run again right after it so the
    # events do not perturb `value`; the GPU is in the same sustained state), plus each layer timed alone.
    peaks = measured_peaks()
    burst, sustained = peaks["bf16_burst"], peaks["bf16_sustained"]
    names, kernels = eng.layer_names(), eng.layer_kernels()
    eng.profile_layers_begin(args.steps)
    for i in range(args.steps):
        eng.process_batch_dev(dev[i % R].data_ptr(), S, S, B, stream, wait=False)
    in_step_ms, passes = eng.profile_layers_read()
    stage_ms, stage_passes = eng.profile_stages_read()
    # HBM-bound stages, in-step: ALGORITHMIC bytes / CUDA-event time inside the same loop
    hbm = peaks["hbm_gbs"]
    alg = {"K1": B * (2 * S * S + 2 * S * S), "K5": B * 2 * S * S, "K6": B * S * S + 8 * n_pts + 4 * n_cnt}
    stages = []
    fused_slices = stage_ms["K5"] < 0.01      # one kernel per slice does K5 and K6 (slice_fused.cuh): only the sum is observable
    stage_ms["K5+K6"] = stage_ms["K5"] + stage_ms["K6"]
    alg["K5+K6"] = alg["K5"] + alg["K6"]
    rows = [("K1", "preprocess: u16 in, u8 out (+ bf16 conversion fused into the first conv)")]
    if fused_slices:
        rows.append(("K5+K6", "postprocess + mask2polygon, one CTA per slice in shared memory + finalize (2 launches); algorithmic bytes = "
                              "the two stages' sum (3 B/px + 8 B/vertex), although the fused kernel never re-reads the clean mask"))
    else:
        rows += [("K5", "postprocess: hole fill + 3x3 open + area filter"), ("K6", "mask2polygon: labels, contour order, mapped vertices")]
    for k, what in rows:
        t = stage_ms[k]
        stages.append({"stage": k, "what": what, "ms": t, "alg_bytes": int(alg[k]), "achieved": alg[k] / t / 1e6 if t > 0 else None,
                       "peak": hbm, "unit": "GB/s", "frac": alg[k] / t / 1e6 / hbm if t > 0 else None, "share_of_step": t / ms_per_step})
    table, by_kernel = [], {}
    layer_bytes = unet_layer_bytes(names, B, S, eng.info.n_classes)
    for li, name in enumerate(names):
        ms_alone, fl = eng.time_layer(li, B, iters=max(3, min(10, args.steps)))
        ms_l = in_step_ms[li]
        table.append({"layer": name, "kernel": kernels[li], "ms_in_step": ms_l, "ms_alone": ms_alone, "gflop": fl / 1e9,
                      "tflops_in_step": fl / ms_l / 1e9 if ms_l > 0 else None})
        k = by_kernel.setdefault(kernels[li], {"kernel": kernels[li], "launches": 0, "ms": 0.0, "ms_alone": 0.0, "flop": 0.0, "alg_bytes": 0})
        k["launches"] += 1
        k["ms"] += ms_l
        k["ms_alone"] += ms_alone
        k["flop"] += fl
        k["alg_bytes"] += layer_bytes.get(name, 0)
    kernel_rows = sorted(by_kernel.values(), key=lambda r: -r["ms"])
    for r in kernel_rows:   # per kernel instantiation: algorithmic FLOP of its launches / their event-timed durations
        r["achieved"] = r["flop"] / (r["ms"] / 1e3) / 1e12                 # inside the step -> sustained peak
        r["frac"] = r["achieved"] / sustained
        r["achieved_alone"] = r["flop"] / (r["ms_alone"] / 1e3) / 1e12     # timed alone -> burst peak
        r["frac_alone"] = r["achieved_alone"] / burst
        r["share_of_step"] = r["ms"] / ms_per_step
        # the other roofline: ALGORITHMIC activation + weight bytes of its launches (every map read once, written once) / the same
        # in-step time, against the measured HBM peak; `bound` names the roofline this instantiation sits closer to
        r["hbm_gbs"] = r["alg_bytes"] / (r["ms"] / 1e3) / 1e9 if r["alg_bytes"] else None
        r["hbm_frac"] = r["hbm_gbs"] / hbm if r["alg_bytes"] else None
        r["bound"] = "hbm" if (r["hbm_frac"] or 0.0) > r["frac"] else "tensor"
        del r["flop"]
    dominant = kernel_rows[0]
    fwd_ms = float(sum(in_step_ms))
    fwd_tflops = eng.info.flops_per_slice * B / (fwd_ms / 1e3) / 1e12
    # DRAM bytes from the committed `ncu --set full` capture of one forward pass at this batch (tools/forward_once.py)
    traffic, traffic_all, traffic_src = None, None, None
    prof = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_ncu_full_forward_b32.json")
    if os.path.exists(prof):
        with open(prof) as f:
            pj = json.load(f)
        if pj.get("batch") == B and S == 512 and len(pj["layers"]) == len(names):
            per_layer = [l["dram_read_bytes"] + l["dram_write_bytes"] for l in pj["layers"]]
            traffic = sum(b for b, k in zip(per_layer, kernels) if k == dominant["kernel"])
            traffic_all = pj["tcgen05_dram_bytes"]
            traffic_src = "profiles/r2_ncu_full_forward_b32.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over this kernel's launches of one step"
    roofline = {"bound": "tensor", "kernel": dominant["kernel"], "launches_per_step": dominant["launches"],
                "achieved": dominant["achieved"], "peak": sustained, "unit": "TFLOP/s", "frac": dominant["achieved"] / sustained,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks["source"] + " cuBLAS bf16, sustained (the kernel is timed inside the step); burst %.1f, nominal dense bf16 2250" % burst,
                "frac_of_nominal": dominant["achieved"] / 2250.0, "share_of_step": dominant["share_of_step"],
                "timing": "CUDA events around every layer launch inside %d steps of the timed loop (ms_profile_layers_*), summed over the kernel's layers" % passes,
                "alone": {"achieved": dominant["achieved_alone"], "peak": burst, "frac": dominant["frac_alone"],
                          "what": "the same launches timed alone (back to back, ms_time_layer) vs burst cuBLAS bf16"},
                "forward_pass": {"achieved": fwd_tflops, "peak": sustained, "frac": fwd_tflops / sustained, "ms": fwd_ms,
                                 "what": "all 22 UNet launches inside the step, vs sustained cuBLAS bf16",
                                 "traffic": traffic_all, "share_of_step": fwd_ms / ms_per_step},
                "by_kernel": kernel_rows,
                "stages": stages, "stages_timing": "CUDA events around K1 | UNet | K5 | K6 inside %d steps of the same loop" % stage_passes,
                "unet_ms_in_step": stage_ms["unet"]}
    if args.layer_table and rank == 0:
        with open(args.layer_table, "w") as f:
            json.dump({"batch": B, "layers": table}, f, indent=1)

    # ---------------- parity of what was just timed: slices of the LAST step's batch against the oracle
    parity = None
    if rank == 0:
        from medseg_b200 import weights as W
        from oracle import pipeline as op
        from oracle.unet_torch import load_unet
        last = hnp[(args.steps - 1) % R]
        pp, pn, pm = eng.process_batch(last, want_norm=True, want_mask=True)
        same_as_timed = (pp.n_points, pp.n_contours) == (n_pts, n_cnt) or R > 1
        net = load_unet(W.make_weights(1234, n_classes), n_classes)
        idx = [0, B - 1] if B > 1 else [0]
        ok, worst = True, 1.0
        for i in idx:
            ref = op.process_slice(last[i], net, head="argmax" if args.head == "argmax" else "binary", n_classes=n_classes)
            agree = float((pm[i] == ref["mask"]).mean())
            worst = min(worst, agree)
            want = op.map_contour_points(op.extract_contours(op.mask_to_image(pm[i])), 1.0, 1.0)
            got = pp.slice(i)
            ok = ok and bool((pn[i] == ref["norm"]).all()) and agree >= 0.999 and len(got) == len(want) and \
                all(a.shape == b.shape and bool((a == b).all()) for a, b in zip(got, want))
        # every slice of the batch: integer stages bit-exact on the kernel's own masks
        for i in range(B):
            want = op.extract_contours(op.mask_to_image(pm[i]))
            got = pp.slice(i)
            ok = ok and len(got) == len(want) and all(a.shape == b.shape and bool((a == b).all()) for a, b in zip(got, want))
        parity = {"slices": len(idx), "ok": bool(ok and same_as_timed), "min_mask_agreement": worst, "polygon_slices_checked": B,
                  "what": "last step's batch: preprocess bit-exact, final mask >= 99.9 % vs the fp32 oracle, polygons bit-exact vs cv2 on the same mask"}

    # ---------------- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        run, cores, run_files = oracle_runner(n_classes, args.head, S, with_files=True)
        vol = host[0].numpy()
        run(vol[:1])
        sample = args.cpu_sample or min(32, len(vol))
        t0 = time.perf_counter()
        run(vol[:sample])
        dtc = time.perf_counter() - t0
        cpu = {"value": sample / dtc, "unit": "slices/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} slices of the step's batch, in memory (no PNG round trips), JSON text included, {dtc:.1f} s"}
        # the reference as shipped: the same call sequence with every PNG / JSON round trip (src/process.cpp:188-262)
        n_files = min(8, sample)
        fdir = os.path.join(td, "as_shipped")
        os.makedirs(fdir, exist_ok=True)
        for i in range(n_files):
            vol[i].tofile(os.path.join(fdir, f"s{i:03d}.raw"))
        t0 = time.perf_counter()
        run_files([os.path.join(fdir, f"s{i:03d}.raw") for i in range(n_files)], os.path.join(fdir, "out"))
        cpu["as_shipped_value"] = n_files / (time.perf_counter() - t0)
        cpu["as_shipped_sample"] = f"{n_files} slices through files: RAW -> PNG/JSON artefacts -> re-read, as process_single_image does"

    info = eng.info
    eng.cleanup()

    # ---------------- cfg3 as written: ONE 256-slice volume through ONE call of ONE process, sharded by contiguous slice
    # blocks over the N GPUs (ms_process_volume_host, one host thread + two streams per GPU).  Rank 0 runs it while the
    # other ranks wait at the barrier with their engines released; strong scaling: the volume is fixed as N grows.
    volume = None
    if args.volume_slices > 0:
        barrier()
        if rank == 0:
            nv = args.volume_slices
            vcfg = dict(cfg)
            vcfg.pop("device", None)
            vcfg["devices"] = list(range(world))
            veng = ms.Engine(vcfg)
            vhost = torch.from_numpy(synth.ct_volume(nv, S, S, first_seed=0)).pin_memory()
            vnp = vhost.numpy()
            for _ in range(3):
                vp, _, _ = veng.process_volume(vnp)
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                vp, _, _ = veng.process_volume(vnp)
            dtv = (time.perf_counter() - t0) / reps
            volume = {"slices": nv, "gpus": veng.device_count(), "value": nv / dtv, "unit": "slices/s", "ms_per_volume": dtv * 1e3,
                      "scaling": "strong", "contours": int(vp.n_contours), "points": int(vp.n_points),
                      "api": "ms_process_volume_host: one process, one call, contiguous slice blocks per GPU, host buffers (H2D + polygons D2H inside)"}
            veng.cleanup()
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier(group=cpu_group)
    if rank == 0:
        line = {"metric": "slices_per_sec", "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"cfg2: batch {B} of {S}x{S} CT-like slices per GPU, {args.head} UNet (31.0M params, random-init blob) -> polygons",
                           "global_batch": world * B, "parallelism": f"slice-sharded x{world}, no collective",
                           "l2": "per-step working set (~0.29 GB of activations per slice) >> 126 MB L2; input batches rotate"},
                "value_sustained": value_sustained, "sustained_steps": sus_steps, "sustained_seconds": sus_ms / 1e3,
                "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "d2h_useful_bytes_per_step": d2h_useful,
                        "api": "ms_submit_batch_host + ms_wait_batch (double buffered)", "sync_call_value": e2e_sync},
                "gpu_launches": int(launches), "clocks": clocks, "p50_ms_per_slice": p50, "p50_api": "ms_submit_batch_host + ms_wait_batch, batch 1 (CUDA graph replay)",
                "p50_sync_call_ms": p50_sync, "roofline": roofline,
                "polygons_last_step": {"contours": int(n_cnt), "points": int(n_pts)},
                "flops_per_slice": int(info.flops_per_slice)}
        line["json_text"] = json_rate
        if volume:
            line["volume"] = volume
        if parity:
            line["parity_check"] = parity
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def main():
    args = parse()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
