// placeholder, replaced below
#include "kernels.h"
namespace b200 { namespace kernels {
int UmmaKChunkElems(DType d) { return d == DType::FP8 ? 128 : 64; }
int UmmaPaddedCin(int Cin, int, int, DType d) { int kc = UmmaKChunkElems(d); return (Cin + kc - 1) / kc * kc; }
bool UmmaSupported(const ConvArgs&) { return false; }
cudaError_t ConvUmma(const ConvArgs&, const UmmaWeights&, cudaStream_t) { return cudaErrorNotSupported; }
}}
