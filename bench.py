#!/usr/bin/env python3
"""bench.py — DenseNet-121 img/s through the B200-native engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp8|bf16|fp32] [--batch 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)
    python bench.py --impl reference ...      (the reference's CPU path: oracle graph interpreter on host cores)

A "step" is one forward pass of one batch (default 256 images, 3x224x224, synthetic) per GPU.
  value        whole-job img/s with the batch already resident in HBM (device events on the engine's stream)
  e2e          same metric through the reference-facing C-ABI call `ModelInfer` with pinned HOST buffers
               (H2D of the fp32 NCHW input and D2H of the logits inside the timed region); `e2e_uint8`: raw uint8 pixels
  roofline     the whole forward (one CUDA graph of the engine's own kernels): algorithmic FLOPs / ms_per_step against the
               tensor peak measured on this box for the same operand type; the HBM view and the per-family split ride along
  legs         the other BASELINE configurations in the same run: fp32 bs256 (like-for-like with the CPU arm), bf16 bs256, bf16 bs64
  strong_scaling  (--gpus N > 1) ONE process, all GPUs: a 256-image request split 256/N per GPU by the engine's own scheduler
  cpu_baseline the CPU oracle ("torch-CPU stand-in for ORT-CPU 1.21.0") on a bounded sample, same box (N=1 only)
Scaling is weak: every rank/GPU processes its own `--batch` images; no collective is on the data path.
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "densenet121_images_per_second"
UNIT = "img/s"
FLOPS_PER_IMAGE = 5.668e9
# algorithmic HBM bytes per image, unfused layer-by-layer dataflow
BYTES_PER_IMAGE = {"fp32": 89.6e6, "bf16": 44.8e6, "fp8": 22.4e6}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measure_box_peaks(device: int):
    """Dense matmul throughput of THIS box for the operand types MEASURED_PEAKS.json does not cover (it has bf16 only): an 8192^3
    cuBLASLt GEMM in fp8-e4m3 (torch._scaled_mm), tf32 and bf16, best of 8 after warm-up (burst figures).  Library GEMMs are
    the roofline denominators, never on the engine's path."""
    out = {}
    try:
        import torch
        dev = torch.device("cuda", device)
        n = 8192
        flops = 2.0 * n ** 3

        def best(fn, reps=8):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            return flops / min(ts) / 1e12

        with torch.cuda.device(dev):
            a = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
            b = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
            out["bf16_tflops"] = best(lambda: torch.matmul(a, b))
            try:
                a8, b8 = a.to(torch.float8_e4m3fn), b.t().contiguous().to(torch.float8_e4m3fn).t()
                one = torch.ones((), device=dev, dtype=torch.float32)
                out["fp8_tflops"] = best(lambda: torch._scaled_mm(a8, b8, scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
            except Exception as e:  # noqa: BLE001
                out["fp8_error"] = repr(e)[:200]
            try:
                torch.backends.cuda.matmul.allow_tf32 = True
                af, bf = a.float(), b.float()
                out["tf32_tflops"] = best(lambda: torch.matmul(af, bf), reps=4)
                torch.backends.cuda.matmul.allow_tf32 = False
            except Exception as e:  # noqa: BLE001
                out["tf32_error"] = repr(e)[:200]
            del a, b
            torch.cuda.empty_cache()
        out["how"] = "torch 8192^3 GEMMs on this box, best of 8 (burst): bf16 matmul, fp8-e4m3 _scaled_mm, tf32 matmul"
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[:200]
    return out


def _ncu_traffic(precision: str, batch: int):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE forward, summed over its launches, from the newest
    committed ncu capture of this configuration (profiles/*traffic_<precision>_bs<batch>.csv, written by tools/ncu_traffic.sh).
    None when no capture of this configuration is committed."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*traffic_{precision}_bs{batch}.csv")))
    if not files:
        return None, None
    path = files[-1]
    total = 0.0
    try:
        with open(path, newline="") as fh:
            rows = [r for r in csv.reader(fh) if r]
        hdr = next(i for i, r in enumerate(rows) if "Metric Name" in r)
        cols = {name: k for k, name in enumerate(rows[hdr])}
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        launches = set()
        for r in rows[hdr + 1:]:
            if len(r) <= cols["Metric Value"] or not r[cols["Metric Name"]].startswith("dram__bytes_"):
                continue
            total += float(r[cols["Metric Value"]].replace(",", "")) * mult.get(r[cols["Metric Unit"]], 1.0)
            launches.add(r[cols["ID"]])
        meta = os.path.splitext(path)[0] + ".json"
        forwards = 1
        if os.path.exists(meta):
            with open(meta) as fh:
                forwards = int(json.load(fh).get("forwards", 1))
        return total / max(1, forwards), os.path.relpath(path, ROOT)
    except Exception:  # noqa: BLE001
        return None, os.path.relpath(path, ROOT)


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled from a
    thread every 10 ms (nvidia-smi -lms is too coarse for a 50-100 ms timed region); falls back to one
    nvidia-smi query when NVML is unavailable."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._nv = None

    def _uuid_index(self):
        # CUDA_VISIBLE_DEVICES may remap indices; NVML enumerates all GPUs of the box
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._uuid_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._nv = None
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def stop(self):
        if self._nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                        "note": "NVML unavailable: single nvidia-smi sample after the timed region"}
            except Exception:  # noqa: BLE001
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        self._stop.set()
        self._t.join(timeout=1)
        sm = self.samples
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm)}


def _dist():
    """(rank, local_rank, world), a MAX-reduce, a device barrier and a HOST barrier; torch.distributed only under torchrun.
    The host barrier runs on a gloo group: ranks that wait for rank 0's single-process legs must not spin inside an NCCL
    kernel on their GPU (that showed up as 91 % "GPU busy" on idle ranks and slowed rank 0's CPU leg)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 0, 1, (lambda v: v), (lambda: None), (lambda: None)
    import torch
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    backend = os.environ.get("B200_BENCH_BACKEND", "nccl")  # "gloo" for the CPU-only tests of this plumbing
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dev = f"cuda:{local}"
        host_group = dist.new_group(backend="gloo")
    else:
        dist.init_process_group(backend)
        dev = "cpu"
        host_group = None

    def reduce_max(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        dist.barrier()
        if dev != "cpu":
            torch.cuda.synchronize()

    def host_barrier():
        dist.barrier(group=host_group) if host_group is not None else dist.barrier()

    return rank, local, world, reduce_max, barrier, host_barrier


def _workload(args, n_gpus: int):
    """The `config` object: identical for the engine arm and the reference arm (same model file, same synthetic images, same
    batch); what differs between the arms is WHERE it runs, and that is in `cpu_baseline` / `dtype`, not here."""
    return {"workload": f"DenseNet-121 3x224x224 forward, batch {args.batch} per GPU (BASELINE.json configs[3])",
            "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus,
            "parallelism": f"dp{n_gpus} (replicated weights, no collective)",
            "l2": "L2 flushed (256 MiB write) before every timed step; activation arena > L2"}


def cpu_oracle_throughput(n_images: int, batch: int, threads: int | None = None, min_seconds: float = 0.0):
    """img/s of the CPU oracle on a bounded sample of the bench workload (same model file, same inputs)."""
    import torch
    from oracle.onnx_oracle import OnnxOracle
    from tools import synth
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    o = OnnxOracle(os.path.join(ROOT, "models", "densenet_onnx", "1", "model.onnx"))
    x = synth.to_model_input(synth.synthetic_images_u8(min(batch, n_images), start=0))
    o.run({"data_0": x[:2]})  # warm-up (thread pools, primitive caches)
    done, t0 = 0, time.perf_counter()
    while done < n_images or time.perf_counter() - t0 < min_seconds:
        o.run({"data_0": x})
        done += len(x)
    dt = time.perf_counter() - t0
    return done / dt, cores, done, dt


def cpu_oracle_latency_1thread(calls: int = 12):
    """bs1 latency of the CPU oracle on ONE thread: the reference pins ONNX Runtime to `SetIntraOpNumThreads(1)` +
    `ORT_SEQUENTIAL` (inference_engine/src/model.cpp:899,902), so this is the leg that mirrors its session options."""
    import numpy as np
    import torch
    from oracle.onnx_oracle import OnnxOracle
    from tools import synth
    prev = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        o = OnnxOracle(os.path.join(ROOT, "models", "densenet_onnx", "1", "model.onnx"))
        x = synth.to_model_input(synth.synthetic_images_u8(1, start=0))
        o.run({"data_0": x})
        ts = []
        for _ in range(calls):
            t0 = time.perf_counter()
            o.run({"data_0": x})
            ts.append(time.perf_counter() - t0)
        p50 = float(np.median(ts))
        return {"bs1_p50_ms": p50 * 1e3, "value": 1.0 / p50, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{calls} single-image forwards, torch.set_num_threads(1) (mirrors SetIntraOpNumThreads(1) + ORT_SEQUENTIAL, "
                          f"reference model.cpp:899,902)"}
    finally:
        torch.set_num_threads(prev)


def run_reference(args):
    """The reference's own CPU implementation of the path: ONNX Runtime cannot be installed here (no wheel,
    no network), so this is the oracle port — a graph interpreter of the same .onnx on torch-CPU kernels,
    all host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as ge
    ge.ensure_fixtures()
    sample = 32  # images per step: a bounded sample of the 256-image batch, run as ONE bs32 forward
    for _ in range(args.warmup):
        cpu_oracle_throughput(sample, sample, threads=os.cpu_count())
    per = []
    cores = 0
    for _ in range(args.steps):
        ips, cores, done, dt = cpu_oracle_throughput(sample, sample, threads=os.cpu_count())
        per.append(dt)
    value = sample * len(per) / sum(per)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(per) / len(per), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _workload(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} images (one bs{sample} forward) per step x {args.steps} steps of the {args.batch}-image "
                                       f"workload; fp32; torch-CPU stand-in for ORT-CPU 1.21.0 (not installable here)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def _device_leg(pkg, synth, precision: str, batch: int, steps: int, warmup: int, device: str, peaks, box):
    """One extra configuration, device-resident, on a fresh model instance (rank 0 only, outside the headline's timed region)."""
    import numpy as np
    os.environ["B200_ENGINE_PRECISION"] = precision
    os.environ["B200_ENGINE_MAX_BATCH"] = str(batch)
    os.environ["B200_ENGINE_DEVICES"] = device
    mgr = pkg.InferenceManager(os.path.join(ROOT, "models"))
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        base = synth.to_model_input(synth.synthetic_images_u8(min(batch, 32), start=0))
        x = np.concatenate([base] * ((batch + len(base) - 1) // len(base)))[:batch]
        m.stage_input(pkg.TensorData("data_0", x))
        m.forward_device(batch, max(3, warmup), True)
        ms = m.forward_device(batch, steps, True)
        value = batch * steps / (float(ms.sum()) * 1e-3)
        tpeak, tnote = _tensor_peak(precision, peaks, box)
        return {"value": value, "unit": UNIT, "ms_per_step": float(ms.mean()), "steps": steps, "batch": batch, "precision": precision,
                "tensor_frac": value * FLOPS_PER_IMAGE / 1e12 / tpeak, "tensor_peak_tflops": tpeak, "tensor_peak_note": tnote,
                "hbm_frac": value * BYTES_PER_IMAGE[precision] / 1e9 / peaks["hbm_gbs"]}
    finally:
        mgr.shutdown()


def _tensor_peak(precision: str, peaks, box):
    """TFLOP/s denominator for a precision mode: measured on this box when available, else derived from the driver-measured bf16."""
    if precision == "fp8":
        if box.get("fp8_tflops"):
            return box["fp8_tflops"], "fp8-e4m3 8192^3 _scaled_mm measured in this run (burst)"
        return 2.0 * peaks["bf16_tflops"], "2 x MEASURED_PEAKS bf16 burst (no fp8 GEMM could be measured)"
    if precision == "bf16":
        return peaks["bf16_tflops"], "MEASURED_PEAKS bf16 burst"
    # fp32 mode = three bf16 MMAs per product: its ceiling on the tensor pipe is bf16 / 3; reported against the tf32 GEMM as well
    return peaks["bf16_tflops"] / 3.0, "MEASURED_PEAKS bf16 burst / 3 (FP32 mode issues three bf16 MMAs per product)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--precision", default=os.environ.get("B200_BENCH_PRECISION", "fp8"), choices=["fp32", "bf16", "fp8"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the extra configurations (fp32/bf16 legs, box peaks, latency)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    rank, local, world, reduce_max, barrier, host_barrier = _dist()
    in_process_multi = world == 1 and args.gpus > 1
    os.environ["B200_ENGINE_PRECISION"] = args.precision
    os.environ["B200_ENGINE_MAX_BATCH"] = str(args.batch)
    os.environ["B200_ENGINE_DEVICES"] = ",".join(str(i) for i in range(args.gpus)) if in_process_multi else str(local)

    import numpy as np
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if rank == 0:
        ge.ensure_fixtures()
    barrier()
    ge.ensure_fixtures()
    from tools import synth

    if not pkg.is_cuda_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    mgr = pkg.InferenceManager(os.path.join(ROOT, "models"))
    mgr.load_model("densenet_onnx")
    model = mgr.get_model("densenet_onnx")
    B = args.batch
    base = synth.to_model_input(synth.synthetic_images_u8(min(B, 32), start=rank * 32))
    x = np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B]

    # pinned host buffers for the end-to-end leg
    import torch
    x_pinned_t = torch.from_numpy(x).pin_memory()
    x_pinned = x_pinned_t.numpy()
    inp = pkg.TensorData("data_0", x_pinned)
    outc = [pkg.OutputConfig("fc6_1", [B, 1000])]

    # ---------------- device-resident leg ----------------
    model.stage_input(inp)
    model.forward_device(B, args.warmup, True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = pkg.kernel_launch_count()
    t_wall0 = time.perf_counter()
    ms = model.forward_device(B, args.steps, True)  # CUDA events on the engine's stream, L2 flushed between steps
    t_wall = time.perf_counter() - t_wall0
    launches = pkg.kernel_launch_count() - launches0
    barrier()
    clocks = sampler.stop()
    total_ms = reduce_max(float(ms.sum()))
    n_gpus = args.gpus
    images = B * args.steps * n_gpus
    value = images / (total_ms * 1e-3)
    logits = model.read_output(B * 1000).reshape(B, 1000)
    assert np.isfinite(logits).all()

    # ---------------- end-to-end leg (host buffers through ModelInfer) ----------------
    # Every step is one blocking ModelInfer call: H2D of that step's pinned input, forward, D2H of its logits.  The headline
    # runs the K calls from `E2E_CLIENTS` client threads (a server has several requests in flight; the engine keeps several
    # execution instances per GPU so one call's PCIe transfer overlaps another's forward); the strictly serial figure (one
    # call at a time) is reported next to it.
    E2E_CLIENTS = 2
    e2e_steps = max(4, min(args.steps, 20))

    def e2e_leg(m, make_input, clients, out_cfg, batch, scale):
        ins = [make_input() for _ in range(clients)]
        outs = [None] * clients
        todo = [e2e_steps // clients + (1 if c < e2e_steps % clients else 0) for c in range(clients)]

        def client(c, calls):
            for _ in range(calls):
                outs[c] = m.infer([ins[c]], out_cfg)[0].data

        def run(calls):
            ths = [threading.Thread(target=client, args=(c, calls[c])) for c in range(1, clients)]
            t0 = time.perf_counter()
            for t in ths:
                t.start()
            client(0, calls[0])
            for t in ths:
                t.join()
            return time.perf_counter() - t0

        run([4] * clients)  # untimed: every execution instance this pattern reaches has its graphs and staging buffers
        barrier()
        dt = reduce_max(run(todo))
        barrier()
        return batch * e2e_steps * scale / dt, outs[0]

    scale = 1 if in_process_multi else n_gpus
    mk_f32 = lambda: pkg.TensorData("data_0", torch.from_numpy(x).pin_memory().numpy())  # noqa: E731
    e2e_serial, out = e2e_leg(model, mk_f32, 1, outc, B, scale)
    e2e_value, out2 = e2e_leg(model, mk_f32, E2E_CLIENTS, outc, B, scale)
    e2e_rel = float(max(np.abs(out - logits).max(), np.abs(out2 - logits).max()) / np.abs(logits).max())

    # ---------------- same call with raw uint8 HWC pixels (the section-8f ingestion extension: 4x fewer PCIe bytes) ----------------
    u8 = np.ascontiguousarray(np.concatenate([synth.synthetic_images_u8(min(B, 32), start=rank * 32)] * ((B + 31) // 32))[:B])
    mk_u8 = lambda: pkg.TensorData("data_0", torch.from_numpy(u8).pin_memory().numpy(), pkg.DataType.UINT8)  # noqa: E731
    u8_serial, out_u8 = e2e_leg(model, mk_u8, 1, outc, B, scale)
    u8_value, out_u8b = e2e_leg(model, mk_u8, E2E_CLIENTS, outc, B, scale)
    u8_equal = bool(np.array_equal(out_u8, out) and np.array_equal(out_u8b, out))

    line = None
    prof = None
    lat = {}
    if rank == 0:
        prof = model.profile_steps(B, 2)
        if not args.no_legs:
            # bs1 latency (p50 over >= 1000 calls after 100 warm-ups) for the headline precision, device + e2e
            try:
                one = pkg.TensorData("data_0", x_pinned[:1])
                model.stage_input(one)
                l_ms = model.forward_device(1, 1100, False)[100:]
                t = []
                one_cfg = [pkg.OutputConfig("fc6_1", [1, 1000])]
                for i in range(1100):
                    t0 = time.perf_counter()
                    model.infer([one], one_cfg)
                    t.append(time.perf_counter() - t0)
                lat = {"precision": args.precision, "calls": 1000, "bs1_p50_ms_device": float(np.median(l_ms)),
                       "bs1_p99_ms_device": float(np.percentile(l_ms, 99)),
                       "bs1_p50_ms_e2e": float(np.median(t[100:]) * 1e3), "bs1_p99_ms_e2e": float(np.percentile(t[100:], 99) * 1e3)}
            except Exception as e:  # noqa: BLE001
                lat = {"error": repr(e)}
    mgr.shutdown()
    barrier()

    # ---------------- strong scaling: ONE process, every GPU, one 256-image request split by the engine's scheduler ----------------
    strong = None
    if args.gpus > 1 and rank == 0:
        try:
            os.environ["B200_ENGINE_PRECISION"] = args.precision
            os.environ["B200_ENGINE_MAX_BATCH"] = str(B)
            os.environ["B200_ENGINE_DEVICES"] = ",".join(str(i) for i in range(args.gpus))
            os.environ["B200_ENGINE_MIN_SHARD"] = str(max(1, min(32, B // args.gpus)))
            mgr2 = pkg.InferenceManager(os.path.join(ROOT, "models"))
            try:
                mgr2.load_model("densenet_onnx")
                m2 = mgr2.get_model("densenet_onnx")
                per = B // args.gpus
                m2.stage_input(pkg.TensorData("data_0", x_pinned))
                m2.forward_device(per, 5, True)
                sms = m2.forward_device(per, args.steps, True)     # every replica runs its 256/N shard, max over replicas per step

                def leg(make_input, clients):
                    ins = [make_input() for _ in range(clients)]
                    todo = [e2e_steps // clients + (1 if c < e2e_steps % clients else 0) for c in range(clients)]

                    def client(c, calls):
                        for _ in range(calls):
                            m2.infer([ins[c]], outc)

                    def run(calls):
                        ths = [threading.Thread(target=client, args=(c, calls[c])) for c in range(1, clients)]
                        t0 = time.perf_counter()
                        [t.start() for t in ths]
                        client(0, calls[0])
                        [t.join() for t in ths]
                        return time.perf_counter() - t0
                    run([3] * clients)
                    return B * e2e_steps / run(todo)
                h2d = {}
                try:
                    for wc in (False, True):
                        one, allg = pkg.measure_h2d(args.gpus, 256, 8, wc)
                        h2d["write_combined" if wc else "pinned"] = {"one_gpu_gbs": one, "all_gpus_gbs": allg}
                    h2d["note"] = ("host->device copy ceiling of this box (B200MeasureH2D: 256 MB x 8 per GPU, GPU 0 alone, then all GPUs at once): "
                                   "e2e with fp32 input needs 0.602 MB per image")
                except Exception as e:  # noqa: BLE001
                    h2d = {"error": repr(e)[:200]}
                strong = {"n_gpus": args.gpus, "global_batch": B, "host_to_device": h2d, "shard_per_gpu": per, "scheduler": "ModelImpl::Execute (persistent per-GPU workers)",
                          "device_value": B * args.steps / (float(sms.sum()) * 1e-3), "device_ms_per_step": float(sms.mean()),
                          "e2e_fp32": leg(mk_f32, E2E_CLIENTS), "e2e_fp32_serial": leg(mk_f32, 1),
                          "e2e_uint8": leg(mk_u8, E2E_CLIENTS), "e2e_uint8_serial": leg(mk_u8, 1), "unit": UNIT,
                          "faulted_replicas": m2.faulted_replicas(),
                          "note": "one process, B200_ENGINE_DEVICES=all: a 256-image ModelInfer call is cut into 256/N contiguous shards, "
                                  "one per GPU, H2D from / D2H into the caller's buffers at the shard offsets; no collective"}
            finally:
                mgr2.shutdown()
        except Exception as e:  # noqa: BLE001
            strong = {"error": repr(e)[:300]}
    # ---------------- BASELINE configs[4]: mixed batch sizes 1..128 from concurrent request threads through ModelInfer ----------------
    # tools/rest_replay.cpp is the stand-in for the Go REST handler (one OS thread per in-flight request, C-ABI only); it runs
    # as its own process over every GPU of the job, pinned fp32 request buffers and raw uint8 pixels.
    replay = None
    if rank == 0 and not args.no_legs:
        exe = os.path.join(ROOT, "build", "rest_replay")
        if os.path.exists(exe):
            replay = {}
            env = dict(os.environ, B200_ENGINE_PRECISION=args.precision, B200_ENGINE_MAX_BATCH=str(max(128, B)),
                       B200_ENGINE_DEVICES=",".join(str(i) for i in range(args.gpus)) if args.gpus > 1 else str(local))
            env.pop("B200_ENGINE_MIN_SHARD", None)
            nthreads = str(min(64, 24 + 8 * args.gpus))
            for key, extra in (("pinned_fp32", ["--pinned"]), ("pinned_uint8", ["--pinned", "--uint8"])):
                try:
                    r = subprocess.run([exe, "--repo", os.path.join(ROOT, "models"), "--threads", nthreads, "--requests", str(1500 * args.gpus)] + extra,
                                       capture_output=True, text=True, timeout=300, env=env)
                    ln = [x for x in r.stdout.splitlines() if x.startswith("{")]
                    j = json.loads(ln[-1])
                    replay[key] = {k: j[k] for k in ("images_per_s", "requests_per_s", "latency_ms_p50", "latency_ms_p99", "threads", "requests", "failed", "gpus_visible")}
                except Exception as e:  # noqa: BLE001
                    replay[key] = {"error": repr(e)[:200]}
            replay["note"] = "closed loop, batch sizes drawn uniformly from {1,2,4,...,128}; tools/rest_replay.cpp (BASELINE.json configs[4])"

    if rank == 0:
        peaks = _peaks()
        box = {} if args.no_legs else measure_box_peaks(local)
        tensor_peak, tensor_note = _tensor_peak(args.precision, peaks, box)
        ms_per_step = total_ms / args.steps
        per_gpu_ips = B / (ms_per_step * 1e-3)
        tfl = per_gpu_ips * FLOPS_PER_IMAGE / 1e12
        gbs = per_gpu_ips * BYTES_PER_IMAGE[args.precision] / 1e9
        traffic, traffic_src = _ncu_traffic(args.precision, B)
        fam = {}
        for p in prof:
            if p["kind"] == "conv":
                key = ("stem" if p["R"] == 7 else "transition" if "trans" in p["name"] else
                       "classifier" if p["H"] == 1 else f"conv{p['R']}x{p['R']}_H{p['H']}")
            else:
                key = p["kind"]
            a = fam.setdefault(key, {"ms": 0.0, "gflop": 0.0})
            a["ms"] += p["ms"]
            a["gflop"] += p["flops"] / 1e9
        fam_ms = sum(a["ms"] for a in fam.values())
        roofline = {"bound": "tensor",
                    "kernel": f"whole forward: one CUDA graph of {int(launches) // max(1, args.steps)} launches of the engine's own kernels "
                              "(stem_conv7x7 / conv1x1_tma / conv3x3_tma / dense_stream / dense_block / pool_bn_relu / maxpool / gap / fc)",
                    "achieved": tfl, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tfl / tensor_peak,
                    "peak_source": tensor_note, "traffic": traffic, "traffic_source": traffic_src,
                    "how": "achieved = 5.668 GFLOP/image (SURVEY.md 8d) x batch / ms_per_step (CUDA events on the engine's stream, per GPU)",
                    "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                            "algorithmic_bytes_per_image": BYTES_PER_IMAGE[args.precision], "peak_source": peaks["source"] + " (MEASURED_PEAKS.json)"},
                    "limiter": "no single saturated resource (ncu, conv3x3 block 1: DRAM 37 %, tensor pipe 46 %, L2 28 %, issue 31 %). Measured "
                               "facts (tools/ubench, profiles/r02q_ubench_*): tcgen05.mma with both operands in shared memory is operand-fetch bound at "
                               "64 B/clk - N=96 91 cycles, N=128 107 cycles per dispatch against nominal 48 / 64, and DenseNet's GEMMs have N = Cout = "
                               "128 / 32; CTA pairs (cta_group::2) or an A operand in tensor memory restore the nominal rate, but neither made the 3x3 or "
                               "the fused dense-layer kernel faster, so the convs are not tensor-bound either; the 1x1 kernels of blocks 1-2 stream at "
                               "~4.8 TB/s of HBM traffic plus a fixed ~10 us per launch; the 14x14 / 7x7 blocks are a per-layer dependency chain "
                               "(epilogue 1 -> 3x3 MMAs -> epilogue 2 + fences). See DESIGN.md sections 5 and 10",
                    "families_ms": {k: round(v["ms"], 4) for k, v in fam.items()},
                    "families_sum_ms": fam_ms,
                    "families_note": "per-step events WITHOUT programmatic-launch overlap; they sum to more than ms_per_step"}
        legs = {}
        if not args.no_legs and args.gpus == 1:   # per-GPU figures: the N=1 run carries them
            for name, prec, bs in (("fp32_bs256", "fp32", 256), ("bf16_bs256", "bf16", 256), ("bf16_bs64", "bf16", 64)):
                if prec == args.precision and bs == B:
                    continue
                try:
                    legs[name] = _device_leg(pkg, synth, prec, bs, max(5, args.steps // 2), 3, str(local), peaks, box)
                except Exception as e:  # noqa: BLE001
                    legs[name] = {"error": repr(e)[:300]}
            if box.get("tf32_tflops") and "fp32_bs256" in legs and "value" in legs["fp32_bs256"]:
                legs["fp32_bs256"]["vs_tf32_gemm_peak"] = legs["fp32_bs256"]["value"] * FLOPS_PER_IMAGE / 1e12 / box["tf32_tflops"]
        cpu = None
        cpu1 = None
        if not args.no_cpu_baseline and world == 1:
            # >= 64 images as bs32 forwards, at least ~10 s of CPU work; rank 0 of a single-process run only (under torchrun the
            # reference arm is the CPU number: OMP is pinned to one thread there and N-1 ranks would idle behind it)
            ips, cores, done, dt = cpu_oracle_throughput(64, 32, threads=os.cpu_count(), min_seconds=10.0)
            cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{done} images as bs32 forwards of the same workload in {dt:.1f}s; fp32; oracle graph interpreter on torch-CPU "
                             f"(stand-in for ORT-CPU 1.21.0, which cannot be installed here)"}
            try:
                cpu1 = cpu_oracle_latency_1thread()
            except Exception as e:  # noqa: BLE001
                cpu1 = {"error": repr(e)[:200]}
        cfg = _workload(args, n_gpus)   # identical to the reference arm's config: same workload on both arms
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp8": "fp8-e4m3 (fp32 accumulate)", "bf16": "bf16 (fp32 accumulate)",
                          "fp32": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate)"}[args.precision],
                "data": "synthetic", "config": cfg, "precision": args.precision,
                "launch": "torchrun one rank per GPU" if world > 1 else ("in-process replicas" if in_process_multi else "single process"),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(B * 1000 * 4),
                        "steps": e2e_steps, "api": "ModelInfer (C-ABI) with pinned host buffers", "requests_in_flight": E2E_CLIENTS,
                        "serial_value": e2e_serial, "max_rel_vs_device_leg": e2e_rel},
                "e2e_uint8": {"value": u8_value, "unit": UNIT, "h2d_bytes_per_step": int(u8.nbytes), "d2h_bytes_per_step": int(B * 1000 * 4),
                              "steps": e2e_steps, "requests_in_flight": E2E_CLIENTS, "serial_value": u8_serial,
                              "api": "ModelInfer with DATATYPE_UINT8 [N,H,W,3] pixels (extension; value/255 + layout on the GPU)",
                              "logits_identical_to_float_path": u8_equal},
                "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_1thread": cpu1, "latency": lat, "legs": legs,
                "strong_scaling": strong, "mixed_replay": replay, "box_peaks": box,
                "wall_clock_check_ms_per_step": 1e3 * t_wall / args.steps}
    if line is not None:
        print(json.dumps(line))
    if world > 1:
        host_barrier()  # idle ranks wait here on the host (gloo) while rank 0 runs its single-process legs, not inside an NCCL kernel
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
