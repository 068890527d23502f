timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for tf in 1 0; do
B200_ENGINE_TILEFUSE=$tf timeout 300 python - <<'P'
import os, json, numpy as np
os.environ["B200_ENGINE_PRECISION"]="fp8"; os.environ["B200_ENGINE_DEVICES"]="0"; os.environ["B200_ENGINE_INSTANCES"]="1"
import __graft_entry__ as ge
pkg=ge.load_package(); ge.ensure_fixtures()
from tools import synth
mgr=pkg.InferenceManager("models"); mgr.load_model("densenet_onnx"); m=mgr.get_model("densenet_onnx")
x=synth.to_model_input(synth.synthetic_images_u8(32,start=0)); x=np.concatenate([x]*8)
m.stage_input(pkg.TensorData("data_0",x))
m.forward_device(256,5,True)
ms=m.forward_device(256,20,True)
print("TILEFUSE",os.environ["B200_ENGINE_TILEFUSE"],"ms/step",float(ms.mean()), "img/s", 256/float(ms.mean())*1e3)
prof=m.profile_steps(256,3)
json.dump(prof,open("gpurun_out/steps_fp8_tile%s.json"%os.environ["B200_ENGINE_TILEFUSE"],"w"))
tot=0
for p in prof:
    if p["ms"]>0.01: print(p["step"],p["name"][-40:],round(p["ms"]*1e3,1))
mgr.shutdown()
P
done
