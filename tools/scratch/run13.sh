B200_DENSE_TRACE=1 B200_ENGINE_GRAPHS=0 timeout 300 python - <<'P'
import os, json, numpy as np
os.environ["B200_ENGINE_PRECISION"]="fp8"; os.environ["B200_ENGINE_DEVICES"]="0"; os.environ["B200_ENGINE_INSTANCES"]="1"
import __graft_entry__ as ge
pkg=ge.load_package(); ge.ensure_fixtures()
from tools import synth
mgr=pkg.InferenceManager("models"); mgr.load_model("densenet_onnx"); m=mgr.get_model("densenet_onnx")
x=synth.to_model_input(synth.synthetic_images_u8(32,start=0)); x=np.concatenate([x]*8)
m.stage_input(pkg.TensorData("data_0",x))
m.forward_device(256,1,True)
mgr.shutdown()
P
