// api_test.cpp — run-time test of the C++ API the reference exposes next to the C bridge:
//   inference::Tensor            reference inference_engine/include/model.h:93-126, src/model.cpp:30-436
//   inference::InferenceManager  reference inference_engine/src/inference_manager.cpp:283-384 (load state machine),
//                                :674-707 (RunInference)
// Prints one line per check ("ok <name>" / "FAIL <name>: why") and the RunInference results as "KAT <v0> <v1>" /
// "DENSE <argmax> <max logit>" lines that tests/test_gpu_parity.py compares with the oracle.  Exit code = failed checks.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <future>
#include <string>
#include <vector>

#include "inference_manager.h"
#include "model.h"

using namespace inference;

static int g_fail = 0;
#define CHECK(name, cond)                                              \
    do {                                                               \
        if (cond) printf("ok %s\n", name);                             \
        else { printf("FAIL %s: %s\n", name, #cond); ++g_fail; }       \
    } while (0)

static void TensorChecks() {
    Shape s;
    s.dims = {2, 3};
    Tensor t("x", DataType::FLOAT32, s);
    CHECK("tensor.meta", t.GetName() == "x" && t.GetDataType() == DataType::FLOAT32 && t.GetShape().dims == s.dims && t.ByteSize() == 24);
    std::vector<float> v = {1, 2, 3, 4, 5, 6}, back;
    CHECK("tensor.set", t.SetData(v));
    CHECK("tensor.get", t.GetData(back) && back == v);
    std::vector<float> wrong(5, 0.f);
    CHECK("tensor.set_wrong_size_rejected", !t.SetData(wrong));
    std::vector<int> as_int;
    CHECK("tensor.get_wrong_type_rejected", !t.GetData(as_int));
    // deep copy: the copy must not alias the source (the reference shallow-copies its device mirror, model.cpp:205)
    Tensor c(t);
    std::vector<float> v2 = {9, 9, 9, 9, 9, 9};
    CHECK("tensor.copy_set", c.SetData(v2));
    CHECK("tensor.copy_is_deep", t.GetData(back) && back == v);
    Tensor a;
    a = t;
    CHECK("tensor.assign", a.GetData(back) && back == v && a.GetName() == "x");
    // Reshape with the same element count updates the shape and keeps the data (the reference's version is a silent
    // no-op there); a different count reallocates
    Shape r;
    r.dims = {3, 2};
    CHECK("tensor.reshape_same_count", t.Reshape(r) && t.GetShape().dims == r.dims && t.GetData(back) && back == v);
    Shape big;
    big.dims = {4, 4};
    CHECK("tensor.reshape_grow", t.Reshape(big) && t.ByteSize() == 64);
    // device mirror round trip
    Tensor g("g", DataType::FLOAT32, s);
    g.SetData(v);
    const bool up = g.toGPU(0);
    CHECK("tensor.toGPU", up);
    CHECK("tensor.toCPU", g.toCPU() && g.GetData(back) && back == v);
    Tensor i64("i", DataType::INT64, s);
    std::vector<long> lv = {1, 2, 3, 4, 5, 6}, lb;
    CHECK("tensor.int64", i64.ByteSize() == 48 && i64.SetData(lv) && i64.GetData(lb) && lb == lv);
    Tensor moved(std::move(g));
    CHECK("tensor.move", moved.GetData(back) && back == v);
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <model_repository> [densenet_input.f32 n]\n", argv[0]);
        return 2;
    }
    TensorChecks();

    InferenceManager mgr(argv[1], 2);
    CHECK("mgr.initialize", mgr.Initialize());
    std::vector<std::string> models = mgr.ListModels();
    bool has_test = false;
    for (auto& m : models) has_test |= (m == "test_model");
    CHECK("mgr.list", has_test);
    CHECK("mgr.state_before", mgr.GetModelState("test_model", "1") != ModelState::LOADED && !mgr.IsModelLoaded("test_model", "1"));
    CHECK("mgr.load", mgr.LoadModel("test_model", "1"));
    CHECK("mgr.loaded", mgr.IsModelLoaded("test_model", "1") && mgr.GetModelState("test_model", "1") == ModelState::LOADED);
    CHECK("mgr.status_json", mgr.GetModelStatus("test_model", "1").find("LOADED") != std::string::npos);
    CHECK("mgr.load_missing_fails", !mgr.LoadModel("no_such_model", "1") && !mgr.GetLastError().empty());

    // RunInference: the input of the reference's test/onnx_test.cpp:92 ([[1, 1, 1]])
    Shape s;
    s.dims = {1, 3};
    Tensor in("input", DataType::FLOAT32, s);
    in.SetData(std::vector<float>(3, 1.0f));
    std::vector<Tensor> outs;
    const bool ok = mgr.RunInference("test_model", "1", {in}, outs);
    CHECK("mgr.run_inference", ok && outs.size() == 1);
    if (ok && !outs.empty()) {
        std::vector<float> y;
        outs[0].GetData(y);
        CHECK("mgr.output_shape", outs[0].GetShape().dims == std::vector<int64_t>({1, 2}) && y.size() == 2);
        if (y.size() == 2) printf("KAT %.9g %.9g\n", y[0], y[1]);
    } else {
        printf("error: %s\n", mgr.GetLastError().c_str());
    }
    std::vector<Tensor> none;
    CHECK("mgr.run_unloaded_model_fails", !mgr.RunInference("densenet_onnx", "1", {in}, none));
    Tensor bad("data", DataType::FLOAT32, s);
    bad.SetData(std::vector<float>(3, 1.0f));
    CHECK("mgr.run_bad_input_name_fails", !mgr.RunInference("test_model", "1", {bad}, none));

    // async unload + callback
    std::promise<bool> done;
    CHECK("mgr.unload_async", mgr.UnloadModelAsync("test_model", "1", [&](bool success, const std::string&, const std::string&) { done.set_value(success); }));
    CHECK("mgr.unload_callback", done.get_future().get());
    CHECK("mgr.unloaded", !mgr.IsModelLoaded("test_model", "1"));

    // DenseNet through RunInference when an input file is given: raw fp32 [n,3,224,224]
    if (argc >= 4) {
        const int n = atoi(argv[3]);
        std::vector<float> x((size_t)n * 3 * 224 * 224);
        FILE* f = fopen(argv[2], "rb");
        const bool read_ok = f && fread(x.data(), 4, x.size(), f) == x.size();
        if (f) fclose(f);
        CHECK("dense.input_file", read_ok);
        std::promise<bool> loaded;
        mgr.LoadModelAsync("densenet_onnx", "", [&](bool success, const std::string&, const std::string&) { loaded.set_value(success); });
        CHECK("dense.load_async", loaded.get_future().get());
        Shape ds;
        ds.dims = {n, 3, 224, 224};
        Tensor din("data_0", DataType::FLOAT32, ds);
        din.SetData(x);
        std::vector<Tensor> douts;
        const bool dok = mgr.RunInference("densenet_onnx", "", {din}, douts);
        CHECK("dense.run_inference", dok && douts.size() == 1);
        if (dok && !douts.empty()) {
            std::vector<float> y;
            douts[0].GetData(y);
            CHECK("dense.output_size", y.size() == (size_t)n * 1000);
            for (int i = 0; i < n && y.size() == (size_t)n * 1000; ++i) {
                int arg = 0;
                for (int c = 1; c < 1000; ++c)
                    if (y[(size_t)i * 1000 + c] > y[(size_t)i * 1000 + arg]) arg = c;
                printf("DENSE %d %d %.9g\n", i, arg, y[(size_t)i * 1000 + arg]);
            }
        } else {
            printf("error: %s\n", mgr.GetLastError().c_str());
        }
        CHECK("dense.unload", mgr.UnloadModel("densenet_onnx", ""));
    }
    mgr.Shutdown();
    printf("failed checks: %d\n", g_fail);
    return g_fail;
}
