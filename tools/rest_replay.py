"""Replay of BASELINE.json configs[4]: mixed batch sizes 1..128 arriving concurrently, as the Go REST server's
/models/densenet_onnx/infer handler would produce them (one goroutine / OS thread per request, each calling the cgo
`ModelInfer`), driven here through the same C-ABI with Python threads (ctypes releases the GIL during the call).

usage: python tools/rest_replay.py [--threads 32] [--requests 600] [--precision fp8] [--devices all] [--coalesce-us 200]
                                   [--sizes 1,2,4,8,16,32,64,128 | --sizes 1]
Prints one JSON line: images/s, requests/s, p50/p99 request latency, coalescer batches/requests."""
import argparse, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=32)
    ap.add_argument("--requests", type=int, default=600)
    ap.add_argument("--precision", default="fp8")
    ap.add_argument("--devices", default="all")
    ap.add_argument("--coalesce-us", type=int, default=0)
    ap.add_argument("--sizes", default="1,2,4,8,16,32,64,128")
    ap.add_argument("--uint8", action="store_true")
    a = ap.parse_args()
    os.environ["B200_ENGINE_PRECISION"] = a.precision
    os.environ["B200_ENGINE_DEVICES"] = a.devices
    os.environ["B200_ENGINE_MAX_BATCH"] = "256"
    os.environ["B200_ENGINE_COALESCE_US"] = str(a.coalesce_us)
    import __graft_entry__ as ge
    pkg = ge.load_package(); ge.ensure_fixtures()
    from tools import synth
    sizes = [int(s) for s in a.sizes.split(",")]
    u8 = synth.synthetic_images_u8(max(sizes), start=8000)
    x = synth.to_model_input(u8)
    mgr = pkg.InferenceManager(os.path.join(ROOT, "models"))
    mgr.load_model("densenet_onnx")
    m = mgr.get_model("densenet_onnx")
    rng = np.random.default_rng(0)
    plan = [int(rng.choice(sizes)) for _ in range(a.requests)]
    for s in sorted(set(plan)):   # warm every batch size once (CUDA graphs, tensor maps)
        m.infer([pkg.TensorData("data_0", x[:s])], [pkg.OutputConfig("fc6_1", [s, 1000])])
    lat, lock, nxt = [], threading.Lock(), [0]

    def worker():
        while True:
            with lock:
                i = nxt[0]; nxt[0] += 1
            if i >= len(plan):
                return
            s = plan[i]
            t = pkg.TensorData("data_0", np.ascontiguousarray(u8[:s]), pkg.DataType.UINT8) if a.uint8 else pkg.TensorData("data_0", x[:s])
            t0 = time.perf_counter()
            m.infer([t], [pkg.OutputConfig("fc6_1", [s, 1000])])
            lat.append(time.perf_counter() - t0)

    b0, r0 = m.coalesce_stats()
    ts = [threading.Thread(target=worker) for _ in range(a.threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]; [t.join() for t in ts]
    dt = time.perf_counter() - t0
    b1, r1 = m.coalesce_stats()
    lat.sort()
    print(json.dumps({"workload": "mixed-batch replay through ModelInfer (C-ABI)", "precision": a.precision, "gpus": pkg.get_device_count() if a.devices == "all" else len(a.devices.split(",")),
                      "threads": a.threads, "requests": len(plan), "sizes": sizes, "uint8": a.uint8, "coalesce_us": a.coalesce_us,
                      "images_per_s": sum(plan) / dt, "requests_per_s": len(plan) / dt,
                      "latency_ms_p50": 1e3 * lat[len(lat) // 2], "latency_ms_p99": 1e3 * lat[int(len(lat) * 0.99) - 1],
                      "coalesced_batches": b1 - b0, "coalesced_requests": r1 - r0}))
    mgr.shutdown()


main()
