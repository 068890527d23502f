"""Ad-hoc GPU measurement: device-resident forward timing + per-step profile.
usage: python tools/gpu_probe.py <precision> <batch> [iters] [--profile] [--check]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge

def main():
    precision, batch = sys.argv[1], int(sys.argv[2])
    iters = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 10
    os.environ["B200_ENGINE_PRECISION"] = precision
    os.environ.setdefault("B200_ENGINE_DEVICES", "0")
    os.environ["B200_ENGINE_MAX_BATCH"] = str(batch)
    pkg = ge.load_package(); ge.ensure_fixtures()
    from tools import synth
    mgr = pkg.InferenceManager(os.path.join(ROOT, "models"))
    t0 = time.time(); mgr.load_model("densenet_onnx"); print(f"load {time.time()-t0:.2f}s")
    m = mgr.get_model("densenet_onnx")
    nimg = min(batch, 16)
    imgs = synth.to_model_input(synth.synthetic_images_u8(nimg, start=500))
    x = np.concatenate([imgs] * ((batch + nimg - 1) // nimg))[:batch]
    m.stage_input(pkg.TensorData("data_0", x))
    m.forward_device(batch, 3, False)
    ms = m.forward_device(batch, iters, True)
    print(f"{precision} bs{batch}: median {np.median(ms):.3f} ms  min {ms.min():.3f}  -> {batch/np.median(ms)*1e3:.0f} img/s (device-resident)")
    if "--check" in sys.argv:
        from oracle.onnx_oracle import OnnxOracle
        got = m.read_output(batch * 1000).reshape(batch, 1000)[:nimg]
        ref = OnnxOracle(os.path.join(ROOT, "models/densenet_onnx/1/model.onnx")).run({"data_0": imgs})[0]
        print("max rel err", np.abs(got - ref).max() / np.abs(ref).max(), "top1 agree", (got.argmax(1) == ref.argmax(1)).mean(),
              "ref top1 in top5", np.mean([r in np.argsort(-g)[:5] for r, g in zip(ref.argmax(1), got)]))
    if "--e2e" in sys.argv:
        t = []
        for _ in range(5):
            t0 = time.perf_counter()
            mgr.run_inference("densenet_onnx", "", [pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [batch, 1000])])
            t.append(time.perf_counter() - t0)
        print(f"e2e host->host: median {np.median(t)*1e3:.2f} ms -> {batch/np.median(t):.0f} img/s")
    if "--profile" in sys.argv:
        prof = m.profile_steps(batch, 3)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"steps_{precision}_bs{batch}.json"), "w") as fh:
            json.dump(prof, fh)
        tot = sum(p["ms"] for p in prof)
        by = {}
        for p in prof:
            k = p["kind"] + (f"{p['R']}x{p['R']}" if p["kind"] == "conv" else "")
            by.setdefault(k, [0.0, 0.0, 0.0]); by[k][0] += p["ms"]; by[k][1] += p["flops"]; by[k][2] += p["bytes"]
        print(f"sum of steps {tot:.3f} ms")
        for k, (t, f, b) in sorted(by.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:16s} {t:8.3f} ms {100*t/tot:5.1f}%  {f/t/1e9 if t else 0:9.1f} TFLOP/s  {b/t/1e6 if t else 0:9.1f} GB/s")
    mgr.shutdown()

main()
