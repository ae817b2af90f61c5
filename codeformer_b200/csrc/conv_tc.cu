// tcgen05 implicit-GEMM convolution engine (sm_100a) -- see conv_tc.cuh.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_tc.cuh"

namespace cfb {

__global__ void tc_split_weights_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo,
                                        int Cout, int Cin, int k) {
  const int64_t total = (int64_t)Cout * Cin * k * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes out [tap][co][ci]
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int tap = (int)(i / ((int64_t)Cout * Cin));
    const float v = w[((int64_t)co * Cin + ci) * k * k + tap];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

int tc_split_weights(const float* oihw, __half* hi, __half* lo, int Cout, int Cin, int k, cudaStream_t st) {
  const int64_t total = (int64_t)Cout * Cin * k * k;
  const int64_t blocks = (total + 255) / 256;
  tc_split_weights_kernel<<<(unsigned)(blocks > 4096 ? 4096 : blocks), 256, 0, st>>>(oihw, hi, lo, Cout, Cin, k);
  CFB_LAUNCH_CHECK();
  return 0;
}

bool tc_supported(const ConvArgs& a) { (void)a; return false; }
size_t tc_scratch_bytes(const ConvArgs& a) { (void)a; return 0; }
int conv_tc(const ConvArgs& a, void* scratch, int sm_count, cudaStream_t st) {
  (void)a; (void)scratch; (void)sm_count; (void)st;
  set_error("conv_tc: tcgen05 engine not built yet");
  return 1;
}

}  // namespace cfb
