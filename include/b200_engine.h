/*
 * b200_engine.h — extension entry points of libinference_engine.so that the reference ABI has
 * no slot for.  Nothing here is needed by the Go server; these exist for measurement
 * (bench.py), parity tests and plan inspection.  Plain C types only.
 *
 * Precision selection (no slot in reference ModelConfig, see SURVEY.md §5 "Config / flags"):
 *   env B200_ENGINE_PRECISION = fp32 | bf16 | fp8      (default fp32), read at Model::Load;
 *   env B200_ENGINE_DEVICES   = "all" | "0,2,3"        (default: all visible GPUs);
 *   env B200_ENGINE_MAX_BATCH = per-GPU arena batch     (default 256).
 *   `config.json` next to model.onnx may carry "precision" / "max_batch_size" with the same
 *   meaning (environment wins).
 * Serving knobs (all read at Model::Load):
 *   env B200_ENGINE_INSTANCES       = execution instances per GPU, 1..8 (default 4; `instance_count` of ModelConfig/config.json);
 *   env B200_ENGINE_COALESCE_US     = request-coalescing window in microseconds (default 0, or 200 when "dynamic_batching": true);
 *   env B200_ENGINE_MIN_SHARD       = smallest per-GPU shard of a split request (default 32);
 *   env B200_ENGINE_STAGE_PAGEABLE  = 0 disables the pinned staging of pageable request buffers; B200_ENGINE_STAGING_MB caps the pool (2048);
 *   env B200_ENGINE_CHAIN / B200_ENGINE_CHAIN_MIN_BATCH = device-side serialisation of big forwards across instances (default on, 64);
 *   env B200_ENGINE_PIPELINE_CHUNK  = sub-batch of the H2D/forward pipeline of a lone request (default 128, 0 = off).
 * Kernel selection (debug / A-B measurements): B200_ENGINE_GRAPHS=0, B200_ENGINE_DENSEFUSE=0, B200_ENGINE_SPLIT_TRANSITION=0,
 *   B200_ENGINE_RESB=0, B200_ENGINE_L1TMA=0 / B200_ENGINE_C3TMA=0 / B200_ENGINE_HALO=0 (force the generic gather kernels),
 *   B200_ENGINE_FP32_EXACT=1 (FP32 mode on the exact FFMA kernels instead of tcgen05 with bf16-split operands),
 *   B200_ENGINE_LAYERFUSE=0|1|auto (e4m3: dense layers of the 56x56 / 28x28 blocks as ONE streaming kernel, kernels_dense_stream.cu;
 *   default auto = the layers it is faster for: 56x56 with one K chunk; 1 = all of both blocks, 0 = none; B200_ENGINE_LAYERFUSE_TSA=0|1
 *   forces its conv1 operand path: 1 = through tensor memory, 0 = in-place transform in shared memory),
 *   B200_ENGINE_L1CSTP=0 (1x1 kernels: epilogue constants from shared memory instead of the kernel-parameter bank),
 *   B200_ENGINE_C3PAIR=1 (e4m3: 3x3 convs as CTA pairs, tcgen05.mma.cta_group::2), B200_DENSE_INTERLEAVE=0 (dense-block megakernel
 *   without the layer-by-layer interleave of two image groups); every variant is bit-identical to the others.
 */
#ifndef B200_ENGINE_H
#define B200_ENGINE_H

#include "inference_bridge.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Version string of the engine build ("b200-engine <n> sm_100a"). Static storage. */
const char* B200EngineVersion(void);

/* Parse + plan `<model_dir>/model.onnx` for `precision` and `max_batch` WITHOUT touching CUDA and
 * return a malloc'd JSON description (steps, buffers, arena bytes, algorithmic FLOPs and HBM bytes
 * per image).  NULL + *error on failure.  Caller frees with free(). */
char* B200PlanDescribe(const char* model_dir, const char* precision, int max_batch, ErrorMessage* error);

/* The multi-GPU batch scheduler's shard plan as a pure function (no CUDA): how `n` samples are split over
 * `gpus` replicas with an arena of `max_batch` samples each.  Writes up to `capacity` (replica, offset, count)
 * triples and returns the number of shards. */
int B200PlanShards(int n, int gpus, int max_batch, int min_shard, int round_robin, int* triples, int capacity);

/* Number of kernels this library has launched in this process since load (all streams/devices). */
uint64_t B200KernelLaunchCount(void);

/* Device-resident forward: input `input_name` must already have been staged with
 * B200ModelStageInput (host -> device copy happens there, outside the timed region).  Runs
 * `iters` forwards of batch `batch` on every replica of the model concurrently (each replica gets
 * the same staged batch) and writes per-iteration device time in milliseconds (CUDA events on the
 * replica's own stream, max over replicas) to ms_out[0..iters).  If l2_flush != 0 a buffer larger
 * than L2 is overwritten between iterations (outside the event pair). */
bool B200ModelStageInput(ModelHandle handle, const TensorData* input, ErrorMessage* error);
/* Number of GPU replicas of this model that raised a CUDA error and were dropped from the shard set (0 on a healthy model). */
int B200ModelFaultedReplicas(ModelHandle handle);
/* Request coalescer counters: batches executed and requests they carried (requests / batches = mean coalesced size). */
bool B200ModelCoalesceStats(ModelHandle handle, int64_t* batches, int64_t* requests);
bool B200ModelForwardDevice(ModelHandle handle, int batch, int iters, int l2_flush, float* ms_out,
                            ErrorMessage* error);
/* Copies the logits of the last B200ModelForwardDevice from replica 0 into `out` (fp32). */
bool B200ModelReadOutput(ModelHandle handle, float* out, size_t out_elems, ErrorMessage* error);

/* Per-step device timing of one forward at `batch` on replica 0 (events between steps).
 * Returns malloc'd JSON [{"step":i,"kind":"conv","name":"…","ms":…,"flops":…,"bytes":…},…]. */
char* B200ModelProfileSteps(ModelHandle handle, int batch, int repeats, ErrorMessage* error);

/* Debug/parity: copy the device tensor that holds ONNX value `value_name` after the last forward on
 * replica 0 back to host as fp32 NCHW (or row-major 2D).  `out_elems` is the capacity of `out`;
 * returns the number of elements written, or -1 on error. */
int64_t B200ModelReadValue(ModelHandle handle, const char* value_name, float* out, size_t out_elems,
                           ErrorMessage* error);

/* ModelInfer with the classification tail on the GPU: runs the forward exactly like ModelInfer
 * (same input rules, FLOAT32 NCHW or UINT8 NHWC pixels), then softmax (if apply_softmax != 0) and top-k of graph output 0 on the
 * device; only k (class index, score) pairs per sample come back.  `classes` / `scores` are caller arrays of N * k entries, row
 * major, best first (equal scores: lowest class index first).  Replaces the full sort of 1000 floats per request in the Go
 * handler (reference server/main.go:744-786).  1 <= k <= 64. */
bool B200ModelInferTopK(ModelHandle handle, const TensorData* inputs, int num_inputs, int k, int apply_softmax, int32_t* classes,
                        float* scores, ErrorMessage* error);

/* Page-locked host memory for request buffers (copy elimination at the boundary): a caller
 * that fills buffers from B200HostAlloc (cgo: C.B200HostAlloc instead of C.malloc, inference_binding.go:670-700) lets
 * ModelInfer DMA straight from / into them; pageable buffers work too but go through the driver's staging copy.
 * NULL on failure.  B200HostFree(NULL) is a no-op. */
void* B200HostAlloc(size_t bytes);
void B200HostFree(void* ptr);

/* Host->device copy bandwidth of this box, the ceiling of every end-to-end figure: `bytes` per GPU are copied `iters` times from
 * page-locked host memory (write-combined when `write_combined` != 0) to each of the first `gpus` devices, first on GPU 0 alone
 * (*gbs_single), then on all of them at once from one host thread per GPU (*gbs_all, aggregate).  GB/s = 1e9 bytes/s. */
bool B200MeasureH2D(int gpus, size_t bytes, int iters, int write_combined, double* gbs_single, double* gbs_all, ErrorMessage* error);

#ifdef __cplusplus
}
#endif

#endif /* B200_ENGINE_H */
