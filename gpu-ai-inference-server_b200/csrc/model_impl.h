// model_impl.h — private implementation behind inference::Model (PIMPL), shared with the C bridge
// so that ModelInfer can hand caller buffers to the GPU without intermediate copies.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "engine.h"
#include "model.h"
#include "plan.h"

namespace inference {

// Borrowed description of one caller tensor (no ownership, no copy).
struct IoDesc {
    std::string name;
    DataType dtype = DataType::FLOAT32;
    std::vector<int64_t> dims;
    const void* data = nullptr;  // host memory
    size_t bytes = 0;
};
struct OutDesc {
    std::string name;           // filled by the engine (graph output name)
    std::vector<int64_t> dims;  // filled by the engine
    void* data = nullptr;       // caller buffer (may be null: nothing is copied)
    size_t capacity = 0;        // bytes available at `data`
    size_t produced = 0;        // bytes the engine produced for this output
    // extension (B200ModelInferTopK, output 0 only): softmax/top-k on the GPU, k (class, score) pairs per sample
    int topk = 0;
    bool topk_softmax = false;
    int32_t* topk_idx = nullptr;
    float* topk_val = nullptr;
};

// Persistent host workers of the multi-GPU batch scheduler ("one host worker thread ... per GPU"): a
// request that is split over GPUs hands each shard to the queue of its GPU instead of creating and joining a thread per request
// and GPU.  `per_gpu` threads serve one GPU's queue, so shards of concurrent requests can still occupy every execution instance.
class GpuWorkers {
public:
    GpuWorkers(int gpus, int per_gpu);
    ~GpuWorkers();
    GpuWorkers(const GpuWorkers&) = delete;
    GpuWorkers& operator=(const GpuWorkers&) = delete;
    std::future<void> Submit(int gpu, std::function<void()> fn);

private:
    struct Queue {
        std::mutex mu;
        std::condition_variable cv;
        std::deque<std::packaged_task<void()>> tasks;
        bool stop = false;
    };
    std::vector<std::unique_ptr<Queue>> queues_;
    std::vector<std::thread> threads_;
};

// One contiguous piece of a batch assigned to one GPU replica.
struct ShardPlan {
    int replica, off, cnt;
};
std::vector<ShardPlan> PlanShards(int n, int G, int max_batch, int min_shard, int round_robin);

class ModelImpl {
public:
    ModelImpl(const std::string& model_path, ModelType type, const ModelConfig& config, DeviceType device, int device_id);
    ~ModelImpl();

    bool Load();
    void Unload();
    bool IsLoaded() const { return loaded_.load(std::memory_order_acquire); }
    bool Infer(const std::vector<Tensor>& inputs, std::vector<Tensor>& outputs);
    // Zero-copy variant used by the C bridge.  `outs` has one entry per caller-provided output slot
    // (matched by POSITION, reference inference_bridge.cpp:794-812).
    bool InferBorrowed(const std::vector<IoDesc>& ins, std::vector<OutDesc>& outs);

    ModelMetadata GetMetadata() const;
    Model::Stats GetStats() const;
    std::string GetLastError() const;
    void SetLastError(const std::string& e) const;

    // extension API (b200_engine.h)
    struct Loaded {
        std::shared_ptr<const b200::Plan> plan;
        std::vector<std::unique_ptr<b200::Replica>> replicas;  // one per GPU (instance 0)
        std::vector<float> last_ms;
        // Further execution instances per GPU (the reference's dead `instance_count` field, model.h:70): each has its own
        // stream, arena and graph cache, so the H2D copy of one caller's batch overlaps the forward of another's.
        int instances = 1;
        std::vector<std::unique_ptr<b200::Replica>> extra;  // [g * (instances - 1) + j]
        std::mutex pick_mu;
        std::condition_variable pick_cv;
        std::vector<int> busy;                               // [g * instances + j]: 1 while a Run call owns the instance
        std::vector<uint64_t> next_ticket, serving;          // per GPU: callers are served strictly first come, first served
        // a replica whose GPU raised a CUDA error is dropped from the shard set ("a failed GPU replica should be
        // dropped from the shard set, not crash"); requests keep being served by the others
        std::vector<std::atomic<int>> faulted;               // [g]
        std::vector<int> Healthy() const;                    // GPU (replica) indices still in the shard set
        std::unique_ptr<GpuWorkers> workers;                 // created when there is more than one replica
        int inject_fault = -1;                               // test hook (B200_ENGINE_FAULT_REPLICA): this replica fails every Run
        int Slots() const { return (int)replicas.size() * instances; }
        b200::Replica* Acquire(int g, int* slot, bool* alone = nullptr);  // blocks until an instance of GPU g is free (FIFO)
        void Release(int slot);
    };
    std::shared_ptr<Loaded> Pin() const;
    int staged_batch = 0;
    void CoalesceStats(int64_t* batches, int64_t* requests) const { *batches = co_batches_.load(); *requests = co_requests_.load(); }
    int FaultedReplicas() const;

private:
    bool ValidateInputs(const std::vector<IoDesc>& ins) const;
    bool Execute(Loaded& st, int n, const std::vector<const void*>& in_ptrs, std::vector<OutDesc>& outs, unsigned u8_mask = 0);

    std::string model_path_;
    ModelType type_;
    ModelConfig config_;
    DeviceType device_type_;
    int device_id_;
    std::atomic<bool> loaded_{false};
    ModelMetadata metadata_;
    std::atomic<int64_t> inference_count_{0}, total_ns_{0}, last_ns_{0};
    std::atomic<size_t> memory_bytes_{0};
    mutable std::mutex err_mu_;
    mutable std::string last_error_;
    mutable std::mutex state_mu_;
    std::shared_ptr<Loaded> state_;
    std::atomic<unsigned> round_robin_{0};

    // ---- request coalescer: concurrent small requests that arrive within a short window are
    // executed as ONE batch.  Off unless config.json says "dynamic_batching": true or B200_ENGINE_COALESCE_US > 0.
    struct Pending {
        int n = 0;
        unsigned u8_mask = 0;
        std::vector<const void*> ptrs;
        std::vector<OutDesc>* outs = nullptr;
        bool done = false, ok = false, promoted = false;
        std::string err;
        std::condition_variable cv;
    };
    bool Coalesce(const std::shared_ptr<Loaded>& st, int n, const std::vector<const void*>& ptrs, std::vector<OutDesc>& outs, unsigned u8_mask);
    void RunCoalesced(Loaded& st, const std::vector<Pending*>& batch);
    int min_shard_ = 32;          // smallest per-GPU shard of a split request (B200_ENGINE_MIN_SHARD)
    bool stage_pageable_ = true;  // copy pageable request buffers into pinned staging on the caller's thread
    int coalesce_us_ = 0;      // collection window
    int coalesce_small_ = 8;   // only requests of at most this many samples are coalesced
    std::mutex co_mu_;
    std::condition_variable co_cv_;
    std::vector<Pending*> co_queue_;
    bool co_leader_ = false;
    int co_inflight_ = 0;      // coalesced batches currently executing (guarded by co_mu_)
    std::atomic<int64_t> co_batches_{0}, co_requests_{0};
};

}  // namespace inference
