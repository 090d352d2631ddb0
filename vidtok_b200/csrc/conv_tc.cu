#include "kernels.h"
namespace vt {
bool conv_tc_supported(const ConvP&) { return false; }
cudaError_t launch_conv_tc(const ConvP&, const bf16*, const bf16*, int, bf16*, cudaStream_t) { return cudaErrorNotSupported; }
const char* conv_tc_last_error() { return ""; }
}
