// cuda_utils.h — device queries; API of reference `inference_engine/include/cuda_utils.h:8-42`.
#ifndef CUDA_UTILS_H
#define CUDA_UTILS_H

#include <string>
#include <vector>

namespace inference {
namespace cuda {

bool IsCudaAvailable();
int GetDeviceCount();
std::string GetDeviceInfo(int device_id = 0);

// result = a + b on the GPU; the reference's smoke kernel (cuda_utils.cu:10-15,63-127).
bool VectorAdd(const std::vector<float>& a, const std::vector<float>& b, std::vector<float>& result);

struct MemoryInfo {
    size_t total;
    size_t free;
    size_t used;
};
MemoryInfo GetMemoryInfo(int device_id = 0);

}  // namespace cuda
}  // namespace inference

#endif  // CUDA_UTILS_H
