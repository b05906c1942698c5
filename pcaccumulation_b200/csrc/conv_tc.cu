// tcgen05 tensor-core convolution path (placeholder until the kernel lands: reports "unsupported").
#include "common.cuh"
#include "pcab200.h"

extern "C" int pcab_conv3x3_tc_supported(int, int, int, int, int, int, int) { return 0; }
extern "C" size_t pcab_conv3x3_tc_pack_floats(int, int) { return 0; }
extern "C" int pcab_conv3x3_tc(const float*, int, const float*, int, const float*, int, int, const float*, const float*,
                               const float*, const float*, int, float*, int, int, int, int, int, int, cudaStream_t) {
  pcab_set_error("pcab_conv3x3_tc: not built");
  return PCAB_ERR_ARG;
}
