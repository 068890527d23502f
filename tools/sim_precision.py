"""Design tool: simulate the engine's reduced-precision pipelines on CPU (torch) to choose
storage/compute formats before writing kernels.  Not part of the product or the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from tools import onnx_lite, synth

E4M3_MAX = 448.0

def q_bf16(x): return x.to(torch.bfloat16).to(torch.float32)
def q_fp8(x, scale):  # scale: multiply before cast; returns dequantised value
    return (x * scale).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).to(torch.float32) / scale
def q_tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF   # round-to-nearest (ties away) to 10 mantissa bits
    return i.view(torch.float32)

class Sim:
    def __init__(self, path, mode, store="bf16", act_scale_mode="dyn"):
        self.m = onnx_lite.load(path); self.mode = mode; self.store = store
        self.calib = {}  # tensor name -> amax
    def qstore(self, x, name):
        if self.mode in ("fp32", "tf32", "bf16x3"): return x
        if self.store == "bf16" or self.mode == "bf16": return q_bf16(x)
        amax = float(x.abs().max()); s = E4M3_MAX / max(amax, 1e-12) / 2
        return q_fp8(x, s)
    def qact(self, x, name):   # MMA A operand
        if self.mode == "fp32": return x
        if self.mode == "tf32": return q_tf32(x)
        if self.mode == "bf16": return q_bf16(x)
        amax = float(x.abs().max()); s = E4M3_MAX / max(amax, 1e-12) / 2
        return q_fp8(x, s)
    def qw(self, w):
        if self.mode == "fp32": return w
        if self.mode == "tf32": return q_tf32(w)
        if self.mode == "bf16": return q_bf16(w)
        amax = w.abs().flatten(1).max(1).values.clamp_min(1e-12).view(-1,1,1,1)
        return q_fp8(w, E4M3_MAX / amax)
    def run(self, x):
        g = self.m.graph
        env = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in g.initializers.items()}
        env[g.inputs[0].name] = self.qstore(torch.from_numpy(x), "in")
        prod = {}
        for n in g.nodes:
            for o in n.outputs: prod[o] = n
        with torch.no_grad():
            for n in g.nodes:
                a = n.attrs; xs = [env[i] for i in n.inputs]; op = n.op_type
                if op == "Conv" and self.mode == "bf16x3":
                    # fp32 emulated on bf16 tensor cores: a = a0 + a1, w = w0 + w1 (bf16 each); a0*w0 + a0*w1 + a1*w0, fp32 accumulate
                    a0 = q_bf16(xs[0]); a1 = q_bf16(xs[0] - a0)
                    w0 = q_bf16(xs[1]); w1 = q_bf16(xs[1] - w0)
                    kw = dict(stride=a["strides"], padding=a["pads"][:2])
                    y = F.conv2d(a1, w0, None, **kw) + F.conv2d(a0, w1, None, **kw) + F.conv2d(a0, w0, xs[2] if len(xs) > 2 else None, **kw)
                    env[n.outputs[0]] = y
                    continue
                if op == "Conv":
                    inp = self.qact(xs[0], n.inputs[0])
                    w = self.qw(xs[1])
                    first = n.inputs[0] == g.inputs[0].name
                    if first and self.mode == "fp8":   # stem kept in bf16
                        inp = q_bf16(xs[0]); w = q_bf16(xs[1])
                    y = F.conv2d(inp, w, xs[2] if len(xs) > 2 else None, stride=a["strides"], padding=a["pads"][:2])
                    # stored after epilogue: if the consumer is Relu, store after relu (handled at Relu)
                    env[n.outputs[0]] = y
                    continue
                elif op == "BatchNormalization":
                    y = F.batch_norm(xs[0], xs[3], xs[4], xs[1], xs[2], False, eps=a.get("epsilon", 1e-5))
                elif op == "Relu":
                    y = torch.relu(xs[0])
                    p = prod[n.inputs[0]]
                    if p.op_type == "Conv": y = self.qstore(y, n.outputs[0])  # conv+relu epilogue output stored
                elif op == "Concat":
                    xs2 = []
                    for nm, t in zip(n.inputs, xs):
                        p = prod.get(nm)
                        if p is not None and p.op_type == "Conv": t = self.qstore(t, nm)  # raw conv output stored
                        xs2.append(t)
                    y = torch.cat(xs2, 1)
                elif op == "MaxPool":
                    y = self.qstore(F.max_pool2d(xs[0], a["kernel_shape"], a["strides"], a["pads"][:2]), n.outputs[0])
                elif op == "AveragePool":
                    y = self.qstore(F.avg_pool2d(xs[0], a["kernel_shape"], a["strides"]), n.outputs[0])
                elif op == "GlobalAveragePool": y = xs[0].mean((2, 3), keepdim=True)
                elif op == "Flatten": y = xs[0].flatten(1)
                elif op == "Gemm": y = xs[0] @ xs[1].t() + xs[2]
                else: raise NotImplementedError(op)
                env[n.outputs[0]] = y
        return env[g.outputs[0].name].numpy()

if __name__ == "__main__":
    path = "models/densenet_onnx/1/model.onnx"
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    x = synth.to_model_input(synth.synthetic_images_u8(N, start=200))
    ref = Sim(path, "fp32").run(x)
    top5 = np.argsort(-ref, 1)[:, :5]
    modes = [("bf16x3", "fp32")] if len(sys.argv) > 2 and sys.argv[2] == "x3" else None
    for mode, store in modes or [("tf32", "fp32"), ("bf16", "bf16"), ("fp8", "bf16"), ("fp8", "fp8")]:
        t = time.time(); y = Sim(path, mode, store).run(x)
        t5 = np.argsort(-y, 1)[:, :5]
        agree_set = np.mean([len(set(a) & set(b)) / 5 for a, b in zip(top5, t5)])
        top1 = np.mean(ref.argmax(1) == y.argmax(1))
        print(f"{mode}/{store}: max|d|={np.abs(y-ref).max():.4f} rel={np.abs(y-ref).max()/np.abs(ref).max():.2e} "
              f"rms={np.sqrt(np.mean((y-ref)**2)):.4f} top1={top1:.3f} top5set={agree_set:.3f} ({time.time()-t:.0f}s)")
