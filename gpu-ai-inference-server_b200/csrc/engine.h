// engine.h — per-GPU executor of a Plan ("replica"): weight replica, activation arena, stream,
// CUDA-graph cache.  One Replica per visible GPU; batches are sharded across replicas with no
// collective (images are independent, weights replicated) — SURVEY.md §8e.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.h"
#include "plan.h"

namespace b200 {

struct CudaError : public std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};
void CudaCheck(cudaError_t e, const char* what);

// Execution instances of one GPU run their FORWARDS one after the other (their host<->device copies overlap freely): the
// tcgen05 kernels are persistent one-CTA-per-SM kernels chained with programmatic dependent launch, and two such chains
// interleaved on the same SMs park each other's pre-launched CTAs (measured: two concurrent forwards took 1.6x as long each).
// The chain is a device-side event hand-over between the instances' streams; the host never blocks on it.
struct ComputeChain {
    explicit ComputeChain(int device);
    ~ComputeChain();
    std::mutex mu;
    cudaEvent_t ev = nullptr;
    int device;
};

// Process-wide pool of page-locked staging buffers: a request whose input lies in pageable memory (cgo's C.malloc) is
// copied into one by the CALLER's thread before it queues for a GPU, so the slow pageable leg runs in parallel across request
// threads instead of inside the driver's serial staging path while an execution instance is held.
class PinnedPool {
public:
    static PinnedPool& Get();
    static bool IsPageable(const void* p);
    void* Take(size_t bytes);  // nullptr when the budget (B200_ENGINE_STAGING_MB, default 2048) is exhausted
    void Give(void* p);
    void Trim();               // frees every idle buffer
    ~PinnedPool();
private:
    struct Buf { void* p; size_t cap; bool used; };
    std::mutex mu_;
    std::vector<Buf> bufs_;
    size_t total_ = 0, budget_ = 0;
};

class Replica {
public:
    static constexpr int kMaxTopK = 64;
    // `weights_of`: another Replica of the SAME plan on the SAME device whose device-resident weights, constants, tensor maps and
    // dense-layer tables this one borrows (execution instances of one GPU share one read-only weight copy: one upload, one
    // quantisation pass, one L2 footprint).  The lender must outlive the borrower.
    Replica(int device, std::shared_ptr<const Plan> plan, bool use_graphs, std::shared_ptr<ComputeChain> chain = nullptr,
            const Replica* weights_of = nullptr);
    ~Replica();
    Replica(const Replica&) = delete;
    Replica& operator=(const Replica&) = delete;

    int device() const { return device_; }
    size_t DeviceBytes() const { return device_bytes_; }
    const Plan& plan() const { return *plan_; }

    // Host-to-host forward of `n` samples (n <= plan.max_batch): H2D of every graph input, all steps,
    // D2H of every graph output (min(capacity, produced) bytes each).  Blocking.  Thread-safe (serialised).
    // `u8_mask` bit i: graph input i is raw uint8 [n][H][W][C] pixels (value / 255 is applied on the GPU) instead of the
    // graph's fp32 NCHW tensor - the uint8-ingestion extension of SURVEY.md section 8f.
    // `alone` = no other execution instance of this GPU is busy: only then is the batch cut into H2D/forward sub-batches
    // (with other requests in flight their forwards already hide this one's copy, and whole batches run more efficiently).
    // `topk` (optional): softmax/top-k of graph output 0 is computed on the GPU after the forward and only k (class, score)
    // pairs per sample travel back (host arrays of n * k entries).
    struct TopK {
        int k = 0;
        bool softmax = false;
        int32_t* idx = nullptr;
        float* val = nullptr;
    };
    void Run(int n, const std::vector<const void*>& host_inputs, const std::vector<void*>& host_outputs,
             const std::vector<size_t>& out_capacity_bytes, unsigned u8_mask = 0, bool alone = true, const TopK* topk = nullptr);

    // Several callers' requests as ONE batch (request coalescing): every segment's inputs are
    // copied from its own host buffers to consecutive sample offsets, one forward of the total runs, and every segment's
    // outputs go straight back to its own host buffers.  sum(n) <= plan.max_batch.
    struct Segment {
        int n = 0;
        std::vector<const void*> in;   // per graph input
        std::vector<void*> out;        // per graph output (may hold nulls)
        std::vector<size_t> cap;       // bytes available at each out
    };
    void RunSegments(const std::vector<Segment>& segs, unsigned u8_mask = 0);

    // Measurement helpers (extension API).
    void StageInput(int input_index, const void* host, int n);
    float ForwardTimed(int n, bool flush_l2);  // ms between events on this replica's stream
    void ReadOutput(int output_index, void* host, size_t bytes);
    std::string ProfileSteps(int n, int repeats);
    int64_t ReadValue(const std::string& value_name, float* out, size_t capacity, int n);

    std::mutex& mutex() { return mu_; }

private:
    struct Prepared {
        kernels::ConvArgs conv;        // Conv
        const float* w_kn = nullptr;   // SIMT weights
        kernels::UmmaWeights umma;     // tcgen05 weights
        bool use_umma = false;
        bool use_f32x3 = false;        // FP32 mode on tcgen05: bf16-split operands (kernels_f32x3.cu); weights live in `umma`
        kernels::View in, in2, out;
        const float* scale = nullptr;
        const float* shift = nullptr;
        size_t in_io_stride = 0, in2_io_stride = 0, out_io_stride = 0;  // bytes per sample when the view is graph I/O
        int fused_run = -1;  // index into dense_runs_ when this step starts a run executed by the dense-block kernel
        bool split_pool = false;  // wide transition: pooled BN+ReLU A operand materialised once, then a plain 1x1 conv
        bool stream_pair = false; // this 1x1 conv and the 3x3 conv after it run as ONE streaming dense-layer kernel (kernels_dense_stream.cu)
        int stream_min_batch = 0; // ... for batches of at least this many samples (auto mode: below it the kernel pair is as fast)
        std::vector<float> h_out_scale;  // host copy of umma.out_scale (kernel-parameter constants of the streaming kernel)
    };
    // Consecutive (1x1 conv, 3x3 conv) step pairs of one dense block executed by ONE persistent kernel.
    struct DenseRun {
        size_t first_step = 0;
        int num_layers = 0;
        kernels::DenseBlockArgs args;
    };

    // Every step on stream_ for samples [off, off+n) of the staged batch (captured into a graph when enabled).
    // Only graph-input / graph-output buffers are indexed by `off`; all intermediate buffers are reused.
    void Enqueue(int n, int off = 0, unsigned u8_mask = 0);
    void EnqueueStep(size_t i, int n, int off = 0, unsigned u8_mask = 0);
    size_t EnqueueAt(size_t i, int n, int off, unsigned u8_mask = 0);
    const uint8_t* U8Source(int tensor, int off, unsigned u8_mask);  // device uint8 staging of a graph input, or null  // runs step i (or the fused run starting there); returns steps consumed
    void BuildDenseRuns();
    void BorrowDenseRuns(const Replica& lender);
    void MarkStreamPairs();
    kernels::View MakeView(int tensor) const;
    void* BufferPtr(int buffer) const;
    void* Upload(const void* host, size_t bytes);

    int device_;
    std::shared_ptr<const Plan> plan_;
    bool use_graphs_;
    std::shared_ptr<ComputeChain> chain_;
    const Replica* weights_of_ = nullptr;  // lender of the weight set, or null when this replica owns its own
    cudaStream_t stream_ = nullptr;
    cudaStream_t copy_stream_ = nullptr;  // H2D of later sub-batches overlaps the forward of earlier ones
    std::vector<cudaEvent_t> copy_events_;
    int chain_min_batch_ = 64;  // forwards of fewer samples leave SMs free and may overlap other instances' (B200_ENGINE_CHAIN_MIN_BATCH)
    int pipeline_chunk_ = 128;  // sub-batch size of the H2D/compute pipeline (B200_ENGINE_PIPELINE_CHUNK, 0 = off)
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    char* arena_ = nullptr;
    char* arena_alloc_ = nullptr;  // cudaMalloc'ed block; arena_ starts kArenaGuard bytes into it (TMA boxes that begin one pixel early)
    void* flush_buf_ = nullptr;
    void* splitk_scratch_ = nullptr;  // fp32 SIMT split-K: tile counters + partial sums (small batches)
    size_t splitk_bytes_ = 0;
    char* topk_dev_ = nullptr;      // [max_batch][kMaxTopK] indices then values (allocated on first use)
    void* pool_scratch_ = nullptr;  // [max_batch][Ho][Wo][Cin] of the widest split transition
    size_t flush_bytes_ = 0;
    std::vector<void*> allocations_;
    std::vector<const float*> dconst_;  // fp32 device copy of every Plan::consts entry that is used as a vector
    std::vector<Prepared> prepared_;
    std::vector<DenseRun> dense_runs_;
    std::vector<uint8_t*> u8_stage_;  // per graph input: device staging for uint8 ingestion (allocated on first use)
    std::map<int64_t, cudaGraphExec_t> graphs_;  // key: (off << 20) | n
    std::map<int64_t, int> graph_launches_;  // kernels per captured forward, for the launch counter
    int launches_per_forward_ = 0;
    size_t device_bytes_ = 0;
    std::mutex mu_;
};

}  // namespace b200
