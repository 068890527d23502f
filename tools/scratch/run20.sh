timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "operator_graphs or low_precision or golden or uint8" 2>&1 | tail -3
for q in 0 1; do
B200_ENGINE_POOL_QUAD=$q timeout 300 python - <<'P'
import os, json, numpy as np
os.environ["B200_ENGINE_PRECISION"]="fp8"; os.environ["B200_ENGINE_DEVICES"]="0"; os.environ["B200_ENGINE_INSTANCES"]="1"
import __graft_entry__ as ge
pkg=ge.load_package(); ge.ensure_fixtures()
from tools import synth
mgr=pkg.InferenceManager("models"); mgr.load_model("densenet_onnx"); m=mgr.get_model("densenet_onnx")
x=synth.to_model_input(synth.synthetic_images_u8(32,start=0)); x=np.concatenate([x]*8)
m.stage_input(pkg.TensorData("data_0",x))
ms=m.forward_device(256,10,True)
prof=m.profile_steps(256,3)
out=m.read_output(256*1000)
print("QUAD",os.environ["B200_ENGINE_POOL_QUAD"], float(ms[3:].mean()), [(p["step"],round(p["ms"]*1e3,1)) for p in prof if p["step"] in (0,1)], float(out.sum()), float(np.abs(out).max()))
mgr.shutdown()
P
done
