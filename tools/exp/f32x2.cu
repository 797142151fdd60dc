// throughput probe: scalar FMUL/FADD vs packed FMUL2/FADD2 (sm_100a), independent chains
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float2 x[4], y[4];
  for (int i = 0; i < 4; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i); y[i] = make_float2(1.0f + i, 2.0f + i); }
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (MODE == 0) {
        x[i].x = x[i].x * a; x[i].y = x[i].y * a;
        y[i].x = y[i].x + b; y[i].y = y[i].y + b;
      } else {
        x[i] = __fmul2_rn(x[i], A);
        y[i] = __fadd2_rn(y[i], B);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 4; ++i) s += x[i].x + x[i].y + y[i].x + y[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, 1.0000001f, 1e-9f, iters); else k<1><<<148 * 8, 256>>>(d, 1.0000001f, 1e-9f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double flops = double(148) * 8 * 256 * iters * 16.0;
      if (rep) printf("mode %d (%s): %.3f ms  %.2f Tflop-ops/s (fp32 mul/add ops)\n", mode, mode ? "packed f32x2" : "scalar", ms, flops / ms / 1e9);
    }
  }
  return 0;
}
