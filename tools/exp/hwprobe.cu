// hwprobe.cu — B200 micro-measurements that decide the structure of the push/deposit kernel
// (DESIGN.md §3.1).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hwprobe hwprobe.cu
//   1. deposit of 48-byte cell records: 3 x RED.128 per lane  vs  smem staging + cp.reduce.async.bulk (TMA)
//   2. scalar FMUL/FADD/FFMA vs packed FMUL2/FFMA2 issue throughput
//   3. nodal gathers: LDG.128/LDG.64 from an L2-resident array vs LDS.128/LDS.64 from a staged box
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// ---- 1a. RED.128 x3 per record; `active` lanes of every warp deposit, cells spread over `ncells`
__global__ void __launch_bounds__(256) k_red128(float4* __restrict__ Jc, unsigned ncells, int active, int iters) {
  const unsigned lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    if (int(lane) < active) {
      // neighbouring lanes hit neighbouring cells (what a nearly sorted container produces)
      const unsigned c = (hash(gw * 131u + it) + lane * 3u) % ncells;
      const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
      atomicAdd(&Jc[3 * size_t(c) + 0], v);
      atomicAdd(&Jc[3 * size_t(c) + 1], v);
      atomicAdd(&Jc[3 * size_t(c) + 2], v);
    }
  }
}
// ---- 1a'. 12 scalar REDs per record
__global__ void __launch_bounds__(256) k_red32(float* __restrict__ Jc, unsigned ncells, int active, int iters) {
  const unsigned lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    if (int(lane) < active) {
      const unsigned c = (hash(gw * 131u + it) + lane * 3u) % ncells;
#pragma unroll
      for (int q = 0; q < 12; ++q) atomicAdd(&Jc[12 * size_t(c) + q], 1.0f + q);
    }
  }
}

// ---- 1b. the same records through shared memory + cp.reduce.async.bulk (48 B each)
__global__ void __launch_bounds__(256) k_bulkred(float4* __restrict__ Jc, unsigned ncells, int active, int iters) {
  __shared__ __align__(128) float4 stage[8][2][32 * 3];   // per warp: two banks of 32 records
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    const int bank = it & 1;
    // the bank was handed to the TMA two iterations ago: wait until it has been read
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    if (int(lane) < active) {
      const unsigned c = (hash(gw * 131u + it) + lane * 3u) % ncells;
      const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
      float4* rec = &stage[w][bank][lane * 3];
      rec[0] = v; rec[1] = v; rec[2] = v;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const unsigned saddr = unsigned(__cvta_generic_to_shared(rec));
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 48;"
                   :: "l"(Jc + 3 * size_t(c)), "r"(saddr) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- 2. FP issue throughput: MODE 0 scalar mul+add, 1 scalar fma, 2 packed FMUL2 + FFMA2(p,1,q), 3 packed FFMA2 only
template <int MODE>
__global__ void __launch_bounds__(256) k_fp(float* out, float a, float b, float one, int iters) {
  float2 x[6], y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i); y[i] = make_float2(1.0f + i, 2.0f + i); }
  const float2 A = make_float2(a, a), B = make_float2(b, b), O = make_float2(one, one);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (MODE == 0) { x[i].x = __fmul_rn(x[i].x, a); x[i].y = __fmul_rn(x[i].y, a); y[i].x = __fadd_rn(y[i].x, b); y[i].y = __fadd_rn(y[i].y, b); }
      else if (MODE == 1) { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a, b); y[i].x = __fmaf_rn(y[i].x, a, b); y[i].y = __fmaf_rn(y[i].y, a, b); }
      else if (MODE == 2) { x[i] = __fmul2_rn(x[i], A); y[i] = __ffma2_rn(y[i], O, B); }
      else { x[i] = __ffma2_rn(x[i], A, B); y[i] = __ffma2_rn(y[i], A, B); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += x[i].x + x[i].y + y[i].x + y[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 3. nodal gathers: 8 corners x (16 B + 8 B) per "particle"; lanes of a warp spread over a few neighbouring cells
constexpr int BX = 4, BY = 8, HZ = 70;                  // staged box: 4 x 8 rows of 70 nodes
__global__ void __launch_bounds__(256) k_gather_global(const float4* __restrict__ nodA, const float2* __restrict__ nodB,
                                                       unsigned nnodes, float* out, int iters) {
  const unsigned lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  const unsigned sj = HZ, si = 70 * 70;
  for (int it = 0; it < iters; ++it) {
    const unsigned h = hash(gw * 977u + it);
    const unsigned base = h % (nnodes - 2 * si - 64);
    const unsigned hl = hash(h + lane);
    const unsigned n = base + (lane >> 4) + (hl & 1u) + ((hl >> 1) & 1u) * sj + ((hl >> 2) & 1u) * si;   // ~2 sort cells, drifted by <= 1
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const unsigned o = n + (c & 1) + ((c >> 1) & 1) * sj + ((c >> 2) & 1) * si;
      const float4 a = __ldg(nodA + o);
      const float2 b = __ldg(nodB + o);
      acc += a.x + a.y + a.z + a.w + b.x + b.y;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void __launch_bounds__(256) k_gather_shared(const float4* __restrict__ nodA, const float2* __restrict__ nodB,
                                                       unsigned nnodes, float* out, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* sA = reinterpret_cast<float4*>(smem);
  float2* sB = reinterpret_cast<float2*>(sA + BX * BY * HZ);
  __shared__ __align__(8) unsigned long long bar;
  const unsigned lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned sbar = unsigned(__cvta_generic_to_shared(&bar));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  unsigned phase = 0;
  const unsigned sj = HZ, si = BY * HZ;
  for (int outer = 0; outer < iters; outer += 16) {
    // stage the box: one bulk copy per row and array (what a block does once per row group)
    if (threadIdx.x < 32) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sbar), "r"(unsigned(BX * BY * HZ * 24)) : "memory");
      }
      __syncwarp();
      const unsigned rbase = (hash(blockIdx.x * 31u + outer) % (nnodes / HZ - 70 * BX - BY)) ;
      for (int r = lane; r < BX * BY; r += 32) {
        const unsigned grow = rbase + (r / BY) * 70 + (r % BY);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(unsigned(__cvta_generic_to_shared(sA + r * HZ))), "l"(nodA + size_t(grow) * HZ), "r"(unsigned(HZ * 16)), "r"(sbar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(unsigned(__cvta_generic_to_shared(sB + r * HZ))), "l"(nodB + size_t(grow) * HZ), "r"(unsigned(HZ * 8)), "r"(sbar) : "memory");
      }
    }
    {
      unsigned done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(sbar), "r"(phase) : "memory");
      phase ^= 1;
    }
    for (int it = outer; it < outer + 16 && it < iters; ++it) {
      const unsigned h = hash(gw * 977u + it);
      const unsigned base = (1 * BY + 2) * HZ + (h % (HZ - 4));
      const unsigned hl = hash(h + lane);
      const unsigned n = base + (lane >> 4) + (hl & 1u) + ((hl >> 1) & 1u) * sj + ((hl >> 2) & 1u) * si;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const unsigned o = n + (c & 1) + ((c >> 1) & 1) * sj + ((c >> 2) & 1) * si;
        const float4 a = sA[o];
        const float2 b = sB[o];
        acc += a.x + a.y + a.z + a.w + b.x + b.y;
      }
    }
    __syncthreads();   // everybody is done with the box before it is overwritten
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F> static float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const unsigned ncells = 343000 * 8;                       // 8 tiles of 70^3 cells: 132 MB of cell-edge records (~L2 size)
  float4* Jc; CK(cudaMalloc(&Jc, size_t(ncells) * 48));
  CK(cudaMemset(Jc, 0, size_t(ncells) * 48));
  const int nblk = 148 * 8, iters = 200;
  printf("== 1. deposit of 48-byte records (per SM: 8 blocks x 8 warps), %d iterations\n", iters);
  for (int active : { 32, 16, 8, 4 }) {
    const double recs = double(nblk) * 8 * active * iters;
    float a = time_ms([&] { k_red128<<<nblk, 256>>>(Jc, ncells, active, iters); });
    float b = time_ms([&] { k_red32<<<nblk, 256>>>(reinterpret_cast<float*>(Jc), ncells, active, iters); });
    float c = time_ms([&] { k_bulkred<<<nblk, 256>>>(Jc, ncells, active, iters); });
    CK(cudaGetLastError());
    printf("  active lanes %2d: RED.128x3 %.3f ms (%.2f Grec/s, %.1f SM-cyc/rec @1.9GHz)  RED.32x12 %.3f ms (%.2f Grec/s)  bulk-reduce %.3f ms (%.2f Grec/s, %.1f SM-cyc/rec)\n",
           active, a, recs / a / 1e6, a * 1e-3 * 1.9e9 * 148 / recs, b, recs / b / 1e6, c, recs / c / 1e6, c * 1e-3 * 1.9e9 * 148 / recs);
  }
  {
    // small footprint: everything L2 resident (one tile)
    const unsigned nc1 = 343000;
    for (int active : { 32, 8 }) {
      const double recs = double(nblk) * 8 * active * iters;
      float a = time_ms([&] { k_red128<<<nblk, 256>>>(Jc, nc1, active, iters); });
      float c = time_ms([&] { k_bulkred<<<nblk, 256>>>(Jc, nc1, active, iters); });
      printf("  one tile (16 MB), active %2d: RED.128x3 %.3f ms (%.2f Grec/s)  bulk-reduce %.3f ms (%.2f Grec/s)\n", active, a, recs / a / 1e6, c, recs / c / 1e6);
    }
  }
  float* out; CK(cudaMalloc(&out, size_t(nblk) * 256 * 4));
  printf("== 2. fp32 issue throughput (24 fp32 lanes-ops per thread-iteration, 6 independent chains)\n");
  {
    const int it2 = 20000;
    const double ops = double(nblk) * 256 * it2 * 24.0;
    float t0 = time_ms([&] { k_fp<0><<<nblk, 256>>>(out, 1.0000001f, 1e-9f, 1.0f, it2); });
    float t1 = time_ms([&] { k_fp<1><<<nblk, 256>>>(out, 1.0000001f, 1e-9f, 1.0f, it2); });
    float t2 = time_ms([&] { k_fp<2><<<nblk, 256>>>(out, 1.0000001f, 1e-9f, 1.0f, it2); });
    float t3 = time_ms([&] { k_fp<3><<<nblk, 256>>>(out, 1.0000001f, 1e-9f, 1.0f, it2); });
    printf("  scalar FMUL+FADD %.3f ms (%.2f Tops/s)   scalar FFMA %.3f ms (%.2f)   FMUL2 + FFMA2(p,1,q) %.3f ms (%.2f)   FFMA2 %.3f ms (%.2f)\n",
           t0, ops / t0 / 1e9, t1, ops / t1 / 1e9, t2, ops / t2 / 1e9, t3, ops / t3 / 1e9);
  }
  printf("== 3. nodal gathers, 8 corners x 24 B per lane\n");
  {
    const unsigned nnodes = 343000 * 8;
    float4* nodA; float2* nodB;
    CK(cudaMalloc(&nodA, size_t(nnodes) * 16)); CK(cudaMalloc(&nodB, size_t(nnodes) * 8));
    CK(cudaMemset(nodA, 0, size_t(nnodes) * 16)); CK(cudaMemset(nodB, 0, size_t(nnodes) * 8));
    const int it3 = 256;
    const double g = double(nblk) * 256 * it3;
    float tg = time_ms([&] { k_gather_global<<<nblk, 256>>>(nodA, nodB, nnodes, out, it3); });
    const size_t sm = size_t(BX) * BY * HZ * 24;
    CK(cudaFuncSetAttribute(k_gather_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    float ts = time_ms([&] { k_gather_shared<<<nblk, 256, sm>>>(nodA, nodB, nnodes, out, it3); });
    CK(cudaGetLastError());
    printf("  global (L1/L2) %.3f ms (%.2f Gparticle-gathers/s, %.1f SM-cyc per warp-gather)   shared box (%zu B, restaged every 16 gathers) %.3f ms (%.2f G/s, %.1f SM-cyc)\n",
           tg, g / tg / 1e6, tg * 1e-3 * 1.9e9 * 148 / (g / 32), sm, ts, g / ts / 1e6, ts * 1e-3 * 1.9e9 * 148 / (g / 32));
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
