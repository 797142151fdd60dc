// common.cuh — shared host/device declarations of libb200pic (sm_100a).
// Memory layout:
//   fields   : per tile, E/B/J each 3*Ch fp32, component-major, k fastest:
//              buf[c*Ch + (i*Hy + j)*Hz + k], H* = N* + 6 (halo 3 each side)
//   particles: per (tile, species) SoA streams x,y,z,ux,uy,uz (fp32) + id (u64)
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200pic.h"

namespace b2p {

constexpr int H = B2P_HALO;
constexpr unsigned long long DEAD = ~0ull;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define B2P_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      throw ::b2p::Error(B2P_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// process context: one device, one stream (all work is stream-ordered)
struct Context {
  int device = -1;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  unsigned long long launches = 0;
  unsigned long long h2d_bytes = 0, d2h_bytes = 0;
  double host_wait_ms = 0.0;   // wall time the host has spent blocked in stream synchronisations (b2p_host_wait_ms)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
Context& ctx();          // initialises lazily on device 0; throws if no GPU
void count_launch(int n = 1);
#define B2P_LAUNCH_CHECK()                  \
  do {                                      \
    ::b2p::count_launch();                  \
    B2P_CUDA(cudaGetLastError());           \
  } while (0)

// Optional per-kernel-class timing with CUDA events on the launch stream
// (bench.py's roofline leg).  Off by default: zero overhead on the hot path.
enum KernelClass {
  KC_NODAL, KC_PUSH, KC_DEPOSIT, KC_SORT_KEYS, KC_RADIX_SORT, KC_GATHER, KC_DETECT, KC_GATHER_OUT, KC_APPEND,
  KC_ZERO, KC_PUSH_B, KC_PUSH_E, KC_ADD_CURRENT, KC_FILTER, KC_HALO, KC_J_EXCHANGE, KC_ENERGY, KC_EDGE_GATHER, KC_OTHER, KC_NCCL, KC_COUNT
};
const char* kernel_class_name(int k);
// cudaStreamSynchronize on the library's current stream, with the blocked wall time booked to Context::host_wait_ms
void timed_stream_sync();

struct ProfScope {
  int idx = -1;
  explicit ProfScope(KernelClass k, double units = 0.0);
  ~ProfScope();
};

// Kernel-variant knobs (b2p_set_option / env B2P_OPTS="name=value,..."): results are identical for
// every setting, only the launch shape / aggregation strategy changes.
struct Tuning {
  int push_minb = 6;      // __launch_bounds__(256, minb) variant of k_push: 5, 6 or 8 resident blocks per SM
  int deposit_minb = 4;   // same for k_deposit_zigzag: 4, 6 or 8
  int deposit_agg = 1;    // warp-level run aggregation before the REDs
  int agg_min = 2;        // ... a step is taken when at least this many lanes of the warp fold
  int filter_chunk = 35;  // i-planes per thread column of k_filter_binomial2
  int filter_ahead = 4;   // planes ahead the pair filter requests its own row into L2 (0 = off; measured 10.3 -> 7.9 ms per 1024^3 pass)
  int stencil_minb = 5;   // resident blocks per SM the extended-stencil push_b is compiled for (2..5; measured 21.1 / 19.1 / 16.1 / 15.6 ms per shock step)
  int energy_cache = 1;   // kinetic-energy account kept by push / pack / append (particles.cuh: KE_SLOTS); 0 = always sum the containers
  int push_streams = 1;   // worker streams the groups of the particle phase are round-robined over (1 = library stream only: with 32 tiles per launch measured 3 % faster than 2)
  int sort_streams = 2;   // worker streams the sort's batches alternate over (0: library stream; at most 2)
  int comm_overlap = 1;   // multi-rank step_pic: 0 = blocking exchanges; 1 = the B exchange before the push flies under the interior tiles' pushes;
                          // 2 = also every exchange of the field phase (boundary tiles first) - bit-identical, measured equal to 1 on 2 and 4 GPUs
                          // (194.9 / 195.1 vs 194.5 / 195.2 ms per step at full size): what the exchanges cost is the wait for the slowest rank
  int filter_pairs = 1;   // binomial2 on lattices with even Hz: two k-adjacent outputs per thread, packed fp32x2 (0: one output per thread)
  int sort_batch = 16;    // containers per launch of the counting-sort kernels (scratch: ~200 MB per 4 M-slot container)
  int sort_overlap = 1;   // b2p_grid_step_pic leaves the sort running on the worker streams under the field phase of the lap
  int push_block = 128;       // threads per block of k_push (128 or 256; 128 measured 1 % faster: finer-grained tail)
  int push_group = 32;    // tiles per group of the particle phase: one launch each of nodal means, scratch clear, push (all containers), edge gather
  int sort_counting = 1;  // counting sort by cell (0: always the general radix sort)
  int defer_tile_calls = 1;   // batch consecutive per-tile calls of one kind (host.cu: deferred per-tile calls)
  int fuse_deposit = 1;   // deposit the stayers' current inside the push kernel (arrivals deposit on append)
};
Tuning& tuning();

void* dmalloc(size_t bytes);   // stream-ordered (cudaMallocAsync)
void dfree(void* p);
template <class T> T* dalloc(size_t n) { return static_cast<T*>(dmalloc(n * sizeof(T))); }

// growable device array bound to the context stream
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~DBuf() { release(); }
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
    return *this;
  }
  void release() { if (p) { dfree(p); p = nullptr; cap = 0; } }
  // ensure capacity >= n; contents are NOT preserved unless keep > 0 (first `keep` elements)
  void reserve(size_t n, size_t keep = 0) {
    if (n <= cap) return;
    reserve_exact(n + n / 8 + 64, keep);
  }
  void reserve_exact(size_t ncap, size_t keep = 0) {
    if (ncap <= cap) return;
    T* q = dalloc<T>(ncap);
    if (keep && p) B2P_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, ctx().stream));
    if (p) dfree(p);
    p = q; cap = ncap;
  }
};

// Pointers read from a job table in device memory are generic to the compiler; telling it that they address global
// memory turns LD / ST / ATOM (+ an address-space query per atomic) into LDG / STG / RED.
#define B2P_GLOBAL(p) __builtin_assume(__isGlobal(p))
#define B2P_GLOBAL_SPECIES(s)                                                                               \
  do {                                                                                                      \
    B2P_GLOBAL((s).x); B2P_GLOBAL((s).y); B2P_GLOBAL((s).z); B2P_GLOBAL((s).ux); B2P_GLOBAL((s).uy);        \
    B2P_GLOBAL((s).uz); B2P_GLOBAL((s).id);                                                                 \
  } while (0)

// ---- device-visible descriptors -------------------------------------------
struct Geom {                 // identical for every tile of a grid
  int N[3];                   // interior cells
  int Hx[3];                  // with halo
  unsigned Ch;                // Hx*Hy*Hz
};

struct FieldPtrs { float* E; float* B; float* J; };

// A packed comp-major slab of a lattice region (multi-GPU staging, comm.cu):
// value(c, ii, jj, kk) = base[c*vol + (ii*dims[1] + jj)*dims[2] + kk]
struct SlabDesc {
  float* base;
  float* field;               // pack only: source lattice (3*Ch)
  int begin[3];               // pack only: first lattice index of the region
  int dims[3];
};

struct Species {              // one particle container on the device
  float* x; float* y; float* z; float* ux; float* uy; float* uz;
  unsigned long long* id;
  unsigned n;                 // container size (dead or alive)
};

}  // namespace b2p
