// cuda_utils.cu — device queries and the vector-add smoke kernel.  Contract of reference
// inference_engine/src/cuda_utils.cu (:17-57 availability/count/info string, :63-127 VectorAdd,
// :129-177 memory info).
#include "cuda_utils.h"

#include <cuda_runtime.h>

#include "kernels.h"

namespace inference {
namespace cuda {

bool IsCudaAvailable() { return GetDeviceCount() > 0; }

int GetDeviceCount() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();  // clear the sticky "no device" status
        return 0;
    }
    return n;
}

std::string GetDeviceInfo(int device_id) {
    cudaDeviceProp prop;
    if (device_id < 0 || device_id >= GetDeviceCount() || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) {
        cudaGetLastError();
        return "Unknown device";
    }
    return "Device " + std::to_string(device_id) + ": " + prop.name + " (Compute Capability " + std::to_string(prop.major) + "." +
           std::to_string(prop.minor) + ")";
}

bool VectorAdd(const std::vector<float>& a, const std::vector<float>& b, std::vector<float>& result) {
    if (a.size() != b.size() || !IsCudaAvailable()) return false;
    size_t n = a.size(), bytes = n * sizeof(float);
    result.resize(n);
    if (n == 0) return true;
    float *da = nullptr, *db = nullptr, *dr = nullptr;
    bool ok = cudaMalloc((void**)&da, bytes) == cudaSuccess && cudaMalloc((void**)&db, bytes) == cudaSuccess &&
              cudaMalloc((void**)&dr, bytes) == cudaSuccess;
    ok = ok && cudaMemcpy(da, a.data(), bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(db, b.data(), bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && b200::kernels::VectorAddF32(da, db, dr, n, 0) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    ok = ok && cudaMemcpy(result.data(), dr, bytes, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(da);
    cudaFree(db);
    cudaFree(dr);
    if (!ok) cudaGetLastError();
    return ok;
}

MemoryInfo GetMemoryInfo(int device_id) {
    MemoryInfo info{0, 0, 0};
    if (device_id < 0 || device_id >= GetDeviceCount()) return info;
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device_id) == cudaSuccess) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            info.total = total_b;
            info.free = free_b;
            info.used = total_b - free_b;
        }
    }
    cudaSetDevice(prev);
    return info;
}

}  // namespace cuda
}  // namespace inference
