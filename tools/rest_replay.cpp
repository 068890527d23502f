// rest_replay.cpp — native replay of BASELINE.json configs[4]: mixed batch sizes 1..128 arriving concurrently, the way the
// Go REST server's /models/densenet_onnx/infer handler produces them: one OS thread per in-flight request (gin goroutines
// blocked in cgo), each filling a C buffer and calling `ModelInfer` (reference server/main.go -> inference_binding.go:650-730
// -> inference_bridge.cpp:692).  Only the C-ABI of include/inference_bridge.h (+ the pinned-allocation extension) is used.
//
//   build:  g++ -O2 -std=c++17 -pthread -Iinclude tools/rest_replay.cpp -o build/rest_replay \
//           -Lgpu-ai-inference-server_b200/lib -linference_engine -Wl,-rpath,'$ORIGIN/../gpu-ai-inference-server_b200/lib'
//   usage:  build/rest_replay [--repo models] [--model densenet_onnx] [--threads 32] [--requests 2000]
//                             [--sizes 1,2,4,8,16,32,64,128] [--uint8] [--pinned] [--seed 0]
//   env:    B200_ENGINE_PRECISION / _DEVICES / _COALESCE_US / _INSTANCES as for the library.
// Prints one JSON line (images/s, requests/s, p50/p99 request latency, coalescer counters).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "b200_engine.h"

static double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    std::string repo = "models", model = "densenet_onnx", sizes_s = "1,2,4,8,16,32,64,128";
    int threads = 32, requests = 2000, seed = 0;
    bool u8 = false, pinned = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--repo") repo = next();
        else if (a == "--model") model = next();
        else if (a == "--threads") threads = atoi(next());
        else if (a == "--requests") requests = atoi(next());
        else if (a == "--sizes") sizes_s = next();
        else if (a == "--seed") seed = atoi(next());
        else if (a == "--uint8") u8 = true;
        else if (a == "--pinned") pinned = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    std::vector<int> sizes;
    for (size_t p = 0; p < sizes_s.size();) {
        size_t q = sizes_s.find(',', p);
        if (q == std::string::npos) q = sizes_s.size();
        sizes.push_back(atoi(sizes_s.substr(p, q - p).c_str()));
        p = q + 1;
    }
    const int max_n = *std::max_element(sizes.begin(), sizes.end());
    const size_t px = 3 * 224 * 224, in_elem = u8 ? 1 : 4;

    InferenceManagerHandle mgr = InferenceInitialize(repo.c_str());
    ErrorMessage err = nullptr;
    if (!mgr || !InferenceLoadModel(mgr, model.c_str(), nullptr, &err)) { fprintf(stderr, "load failed: %s\n", err ? err : "?"); return 1; }
    ModelHandle h = GetModelHandle(mgr, model.c_str(), nullptr, &err);
    if (!h) { fprintf(stderr, "no handle: %s\n", err ? err : "?"); return 1; }

    std::mt19937 rng(seed);
    std::vector<int> plan(requests);
    for (auto& s : plan) s = sizes[rng() % sizes.size()];
    long long images = 0;
    for (int s : plan) images += s;

    // one request buffer pair per worker, as a server keeps per-connection buffers: malloc (pageable, what cgo's C.malloc
    // gives the reference) or the engine's pinned allocator
    struct Worker { void* in; float* out; };
    std::vector<Worker> W(threads);
    for (auto& w : W) {
        const size_t ib = (size_t)max_n * px * in_elem, ob = (size_t)max_n * 1000 * 4;
        w.in = pinned ? B200HostAlloc(ib) : malloc(ib);
        w.out = (float*)(pinned ? B200HostAlloc(ob) : malloc(ob));
        if (!w.in || !w.out) { fprintf(stderr, "allocation failed\n"); return 1; }
        if (u8) { uint8_t* p = (uint8_t*)w.in; for (size_t i = 0; i < ib; ++i) p[i] = (uint8_t)((i * 2654435761u) >> 24); }
        else { float* p = (float*)w.in; for (size_t i = 0; i < ib / 4; ++i) p[i] = (float)((i * 2654435761u) >> 24) / 255.f; }
    }
    auto call = [&](Worker& w, int n) -> bool {
        int64_t idims[4] = {n, 3, 224, 224}, udims[4] = {n, 224, 224, 3}, odims[4] = {n, 1000, 1, 1};
        TensorData in{}, out{};
        in.name = "data_0"; in.data_type = u8 ? DATATYPE_UINT8 : DATATYPE_FLOAT32;
        in.shape.dims = u8 ? udims : idims; in.shape.num_dims = 4;
        in.data = w.in; in.data_size = (size_t)n * px * in_elem;
        out.name = "fc6_1"; out.data_type = DATATYPE_FLOAT32; out.shape.dims = odims; out.shape.num_dims = 4;
        out.data = w.out; out.data_size = (size_t)n * 1000 * 4;
        ErrorMessage e = nullptr;
        bool ok = ModelInfer(h, &in, 1, &out, 1, &e);
        if (!ok) { fprintf(stderr, "ModelInfer(%d) failed: %s\n", n, e ? e : "?"); if (e) FreeErrorMessage(e); }
        return ok;
    };
    for (int s : sizes) if (!call(W[0], s)) return 1;  // warm every batch size once (CUDA graphs, tensor maps)
    for (int t = 1; t < threads && t < 8; ++t) call(W[t], sizes[0]);

    int64_t b0 = 0, r0 = 0, b1 = 0, r1 = 0;
    B200ModelCoalesceStats(h, &b0, &r0);
    std::atomic<int> next{0}, failed{0};
    std::vector<std::vector<double>> lat(threads);
    std::vector<std::thread> pool;
    const double t0 = Now();
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= requests) return;
                double a = Now();
                if (!call(W[t], plan[i])) failed.fetch_add(1);
                lat[t].push_back(Now() - a);
            }
        });
    for (auto& th : pool) th.join();
    const double dt = Now() - t0;
    B200ModelCoalesceStats(h, &b1, &r1);
    std::vector<double> all;
    for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    const char* prec = getenv("B200_ENGINE_PRECISION");
    const char* co = getenv("B200_ENGINE_COALESCE_US");
    const char* inst = getenv("B200_ENGINE_INSTANCES");
    printf("{\"workload\": \"mixed-batch replay through ModelInfer (C-ABI, native threads)\", \"precision\": \"%s\", \"gpus_visible\": %d, "
           "\"threads\": %d, \"requests\": %d, \"failed\": %d, \"sizes\": \"%s\", \"uint8\": %s, \"host_buffers\": \"%s\", \"coalesce_us\": %s, \"instances\": \"%s\", "
           "\"images_per_s\": %.1f, \"requests_per_s\": %.1f, \"latency_ms_p50\": %.3f, \"latency_ms_p99\": %.3f, "
           "\"coalesced_batches\": %lld, \"coalesced_requests\": %lld}\n",
           prec ? prec : "default", GetDeviceCount(), threads, requests, failed.load(), sizes_s.c_str(), u8 ? "true" : "false",
           pinned ? "pinned (B200HostAlloc)" : "pageable (malloc)", co ? co : "0", inst ? inst : "default",
           images / dt, requests / dt, 1e3 * all[all.size() / 2], 1e3 * all[std::min(all.size() - 1, (size_t)(all.size() * 0.99))],
           (long long)(b1 - b0), (long long)(r1 - r0));
    InferenceShutdown(mgr);
    return failed.load() ? 1 : 0;
}
