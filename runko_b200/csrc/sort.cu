// sort.cu — ParticleContainer::sort (pic/particle.h:575-703, key from pic/tile.c++:430-435): stable sort by
// cell key, dead slots last.
//
// Hot path: a stable counting sort by cell, BATCHED over containers — every kernel below is launched once for a
// whole batch (blockIdx.y indexes a device table of SortJob), so sorting the 1024 containers of a 512^3 block is a
// few dozen launches instead of several thousand:
//   1. k_sort_count   key[n]; rank[n] = arrival order among the particles of that key (one returning atomic on
//                     cnt[key] per run of equal keys inside a warp — NOT slot order yet); dead slots are not ranked
//   2. k_sort_chunk_sums + k_sort_scan_chunks   offs = exclusive scan of cnt (nkeys + 2 counters) and the largest
//                     cell population (a routing hint for the container's NEXT sort)
//   3. k_sort_scatter members[offs[key] + rank] = n: the slots of every cell, in arrival order
//   4. k_sort_cells   one thread per cell puts the cell's list into slot order in shared memory (insertion sort of a
//                     few tens of entries); lists longer than 64 are ranked by the whole block
//   5. k_sort_place   one thread per destination slot: dst[e] = src[members[e]] for the seven streams; destination
//                     slots behind the alive particles get the dead id.
// Result: exactly the stable order for any input.  Step 4 is quadratic in the population of a cell, so the host
// routes containers whose last known largest cell exceeded SORT_RADIX_POP to a general radix sort of (key, slot)
// pairs (CUB; crowded containers only — never the steady state of a plasma).
#include "particles.cuh"
#include "pmath.cuh"
#include "sortnet.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

namespace b2p {

// ------------------------------------------------------------------- sort --
// pic/tile.c++:430-435 + pic/particle.h:607-614: key = layout_right cell index in
// the haloed lattice (uint32), dead -> UINT32_MAX.
__global__ void __launch_bounds__(256)
k_sort_keys(const Species s, const Geom g, const float3 origo, unsigned* __restrict__ keys, unsigned* __restrict__ idx,
            const unsigned dead_key) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= s.n) return;
  unsigned key = dead_key;
  if (s.id[n] != DEAD) {
    const unsigned i = __float2uint_rz(s.x[n] - origo.x);
    const unsigned j = __float2uint_rz(s.y[n] - origo.y);
    const unsigned k = __float2uint_rz(s.z[n] - origo.z);
    key = (i * unsigned(g.Hx[1]) + j) * unsigned(g.Hx[2]) + k;
    if (dead_key != 0xFFFFFFFFu && key > dead_key) key = dead_key;   // sort path: clamp to the dead key (= Ch)
  }
  keys[n] = key;
  if (idx) idx[n] = n;
}

// gather all seven streams through the sort permutation (pic/particle.h:640-701)
__global__ void __launch_bounds__(256)
k_gather(const Species src, const Species dst, const unsigned* __restrict__ perm) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= src.n) return;
  const unsigned p = perm[n];
  dst.x[n] = src.x[p]; dst.y[n] = src.y[p]; dst.z[n] = src.z[p];
  dst.ux[n] = src.ux[p]; dst.uy[n] = src.uy[p]; dst.uz[n] = src.uz[p];
  dst.id[n] = src.id[p];
}


// ------------------------------------------------------------ batched counting sort --
constexpr int SORT_SLOTS_PER_THREAD = 4;

__device__ __forceinline__ unsigned cell_key(const float px, const float py, const float pz, const float3 origo, const Geom& g,
                                             const unsigned dead_key) {
  const unsigned i = __float2uint_rz(px - origo.x);
  const unsigned j = __float2uint_rz(py - origo.y);
  const unsigned k = __float2uint_rz(pz - origo.z);
  const unsigned key = (i * unsigned(g.Hx[1]) + j) * unsigned(g.Hx[2]) + k;
  return key > dead_key ? dead_key : key;   // positions outside the lattice (undefined in the reference) share the dead key
}

__global__ void __launch_bounds__(256)
k_sort_count(const SortJob* __restrict__ jobs, const Geom g, const unsigned dead_key) {
  const SortJob jb = jobs[blockIdx.y];
  const unsigned n_total = jb.src.n;
  const unsigned first = blockIdx.x * (256u * SORT_SLOTS_PER_THREAD);
  if (first >= n_total) return;
  const Species s = jb.src;
  B2P_GLOBAL_SPECIES(s); B2P_GLOBAL(jb.cnt); B2P_GLOBAL(jb.keys); B2P_GLOBAL(jb.rank);
  const float3 origo = jb.origo;
  unsigned* __restrict__ cnt = jb.cnt;
  // SORT_SLOTS_PER_THREAD slots per thread, 256 apart: all loads, then all atomics, are in flight together
  const unsigned lane = threadIdx.x & 31;
  unsigned n[SORT_SLOTS_PER_THREAD], key[SORT_SLOTS_PER_THREAD], start[SORT_SLOTS_PER_THREAD], base[SORT_SLOTS_PER_THREAD];
  bool in[SORT_SLOTS_PER_THREAD];
  unsigned long long id[SORT_SLOTS_PER_THREAD];
  float px[SORT_SLOTS_PER_THREAD], py[SORT_SLOTS_PER_THREAD], pz[SORT_SLOTS_PER_THREAD];
#pragma unroll
  for (int r = 0; r < SORT_SLOTS_PER_THREAD; ++r) {
    n[r] = first + 256u * r + threadIdx.x;
    in[r] = n[r] < n_total;
    id[r] = DEAD; px[r] = py[r] = pz[r] = 0.f;
    if (in[r]) { id[r] = ld_pinned(s.id + n[r]); px[r] = ld_pinned(s.x + n[r]); py[r] = ld_pinned(s.y + n[r]); pz[r] = ld_pinned(s.z + n[r]); }
  }
#pragma unroll
  for (int r = 0; r < SORT_SLOTS_PER_THREAD; ++r) {
    key[r] = id[r] != DEAD ? cell_key(px[r], py[r], pz[r], origo, g, dead_key) : dead_key;
    // one atomic per run of equal keys in the warp (a container sorted a few laps ago is made of such runs)
    const unsigned prev = __shfl_up_sync(0xffffffffu, key[r], 1);
    const bool head = lane == 0 || key[r] != prev || !in[r];
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    start[r] = 31u - __clz(hm & (0xFFFFFFFFu >> (31u - lane)));                 // my run's first lane
    const unsigned above = lane == 31 ? 0u : (hm >> (lane + 1));
    const unsigned len = above ? unsigned(__ffs(above)) : 32u - lane;            // for a head: length of its run
    base[r] = 0;
    if (head && key[r] != dead_key) base[r] = atomicAdd(&cnt[key[r]], len);
  }
#pragma unroll
  for (int r = 0; r < SORT_SLOTS_PER_THREAD; ++r) {
    const unsigned b0 = __shfl_sync(0xffffffffu, base[r], start[r]);
    if (in[r]) {
      jb.keys[n[r]] = key[r];
      jb.rank[n[r]] = b0 + (lane - start[r]);
    }
  }
}

// Exclusive scan of the ncount = nkeys + 2 counters of every container of the batch, in two fully parallel
// kernels over chunks of SCAN_CHUNK counters: (a) chunk totals (+ the largest population among the alive keys),
// (b) every chunk adds up the totals of the chunks before it (at most a few hundred values) and scans itself.
constexpr unsigned SCAN_CHUNK = 2048;        // counters per block of 256 threads (8 per thread)
__device__ __forceinline__ unsigned block_exclusive_scan_256(const unsigned v, unsigned* total) {
  __shared__ unsigned wsum[8];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= unsigned(o)) incl += t; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  unsigned before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { const unsigned s = wsum[w]; if (unsigned(w) < wid) before += s; all += s; }
  *total = all;
  return before + incl - v;
}
__global__ void __launch_bounds__(256)
k_sort_chunk_sums(const SortJob* __restrict__ jobs, const unsigned nkeys, const unsigned nchunks) {
  const SortJob jb = jobs[blockIdx.y];
  B2P_GLOBAL(jb.cnt); B2P_GLOBAL(jb.chunk_sums);
  const unsigned ncount = nkeys + 2u;
  const unsigned i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * 8u;
  unsigned sum = 0, mx = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const unsigned i = i0 + q;
    const unsigned v = i < ncount ? jb.cnt[i] : 0u;
    sum += v;
    if (i < nkeys) mx = max(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  __shared__ unsigned ss[8], sm[8];
  if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = sum; sm[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned a = 0, m = 0;
    for (int w = 0; w < 8; ++w) { a += ss[w]; m = max(m, sm[w]); }
    jb.chunk_sums[blockIdx.x] = a;
    if (m) atomicMax(jb.chunk_sums + nchunks, m);                  // [nchunks]: largest population (zeroed with cnt)
  }
}
__global__ void __launch_bounds__(256)
k_sort_scan_chunks(const SortJob* __restrict__ jobs, const unsigned nkeys, const unsigned nchunks) {
  const SortJob jb = jobs[blockIdx.y];
  B2P_GLOBAL(jb.cnt); B2P_GLOBAL(jb.chunk_sums); B2P_GLOBAL(jb.offs);
  const unsigned ncount = nkeys + 2u;
  // totals of the chunks before mine
  unsigned part = 0;
  for (unsigned c = threadIdx.x; c < blockIdx.x; c += 256u) part += jb.chunk_sums[c];
  __shared__ unsigned sbase;
  if (threadIdx.x == 0) sbase = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0 && part) atomicAdd(&sbase, part);
  __syncthreads();
  const unsigned base = sbase;
  const unsigned i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * 8u;
  unsigned v[8], sum = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) { v[q] = i0 + q < ncount ? jb.cnt[i0 + q] : 0u; sum += v[q]; }
  unsigned total;
  unsigned run = base + block_exclusive_scan_256(sum, &total);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (i0 + q < ncount) jb.offs[i0 + q] = run;
    run += v[q];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *jb.max_pop = jb.chunk_sums[nchunks];   // page-locked hint slot (zero-copy store)
}

__global__ void __launch_bounds__(256)
k_sort_scatter(const SortJob* __restrict__ jobs, const unsigned dead_key) {
  const SortJob& jb = jobs[blockIdx.y];
  const unsigned n_total = jb.src.n;
  const unsigned first = blockIdx.x * (256u * SORT_SLOTS_PER_THREAD);
  if (first >= n_total) return;
  const unsigned* __restrict__ keys = jb.keys;
  const unsigned* __restrict__ rank = jb.rank;
  const unsigned* __restrict__ offs = jb.offs;
  unsigned n[SORT_SLOTS_PER_THREAD], k[SORT_SLOTS_PER_THREAD], r_[SORT_SLOTS_PER_THREAD], o[SORT_SLOTS_PER_THREAD];
#pragma unroll
  for (int r = 0; r < SORT_SLOTS_PER_THREAD; ++r) {
    n[r] = first + 256u * r + threadIdx.x;
    const bool in = n[r] < n_total;
    k[r] = in ? keys[n[r]] : dead_key;
    r_[r] = in ? rank[n[r]] : 0u;
  }
#pragma unroll
  for (int r = 0; r < SORT_SLOTS_PER_THREAD; ++r) o[r] = k[r] != dead_key ? offs[k[r]] : 0u;
#pragma unroll
  for (int r = 0; r < SORT_SLOTS_PER_THREAD; ++r)
    if (k[r] != dead_key) jb.members[o[r] + r_[r]] = n[r];
}

// Step 4: every cell's member list, which the scatter left in ARRIVAL order, is put into slot order — the stable
// order of the reference's sort (pic/particle.h:607-640).  One thread per cell: a block stages the lists of its 256
// consecutive cells (one contiguous range of `members`) in shared memory with coalesced loads, each thread sorts its
// own list (registers, a fixed min/max network for up to 32 entries; an insertion sort up to 64), and the range is
// written back.  Lists longer than CELLS_SMALL_POP are ranked by the whole block afterwards (every entry counts the smaller
// entries of its list; `rank`, dead since the scatter, is the staging buffer); ranges that do not fit the staging
// buffer are ordered in global memory.
constexpr unsigned CELLS_SMALL_POP = 64;
constexpr unsigned CELLS_STAGE = 6144;       // entries staged per block
// Lists of a uniform plasma start 16 entries apart, i.e. on two of the 32 banks: one padding word per 16 entries
// (start 17 apart) spreads the threads' lists over all banks.
__device__ __forceinline__ unsigned pad16(const unsigned i) { return i + (i >> 4); }

template <class At>
__device__ __forceinline__ void insertion_sort(At at, const unsigned n) {
  for (unsigned i = 1; i < n; ++i) {
    const unsigned v = at(i);
    unsigned j = i;
    while (j > 0) {
      const unsigned w = at(j - 1);
      if (w <= v) break;
      at(j) = w;
      --j;
    }
    at(j) = v;
  }
}

// Batcher's odd-even merge sort of N = 2^m values held in registers: a fixed network of min/max pairs (63 for 16,
// 191 for 32), no branches and no memory traffic — the lanes of a warp sort their lists in lockstep whatever the
// arrival order was (an insertion sort pays the longest list and the worst order of the 32 lanes at every step).
template <int N, class At>
__device__ __forceinline__ void network_sort(At at, const unsigned n, const bool active) {
  static_assert(N == 16 || N == 32, "networks are generated for 16 and 32 inputs (sortnet.cuh)");
  unsigned v[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    v[k] = 0xFFFFFFFFu;
    if (active && unsigned(k) < n) v[k] = at(unsigned(k));
  }
#define B2P_CE(a, b) { const unsigned lo_ = min(v[a], v[b]), hi_ = max(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
  if (N == 16) { B2P_SORTNET16(B2P_CE) } else { B2P_SORTNET32(B2P_CE) }
#undef B2P_CE
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (active && unsigned(k) < n) at(unsigned(k)) = v[k];
}

// one thread orders one list of 2..CELLS_SMALL_POP entries; the network width is a warp-uniform choice
template <class At>
__device__ __forceinline__ void order_list(At at, const unsigned pop) {
  const bool net = pop >= 2u && pop <= 32u;
  const unsigned wmax = __reduce_max_sync(0xffffffffu, net ? pop : 0u);
  if (wmax > 16u) network_sort<32>(at, pop, net);
  else if (wmax >= 2u) network_sort<16>(at, pop, net);
  if (pop > 32u && pop <= CELLS_SMALL_POP) insertion_sort(at, pop);
}

// 4 blocks per SM (64 registers, the 32-value network spills 144 B): 2 and 3 measured 15 % / 3 % slower
__global__ void __launch_bounds__(256, 4)
k_sort_cells(const SortJob* __restrict__ jobs, const unsigned nkeys) {
  const SortJob jb = jobs[blockIdx.y];
  B2P_GLOBAL(jb.offs); B2P_GLOBAL(jb.members); B2P_GLOBAL(jb.rank);
  const unsigned c0 = blockIdx.x * 256u;
  if (c0 >= nkeys) return;
  __shared__ unsigned s_off[257];
  __shared__ unsigned buf[CELLS_STAGE + CELLS_STAGE / 16 + 1];
  __shared__ unsigned n_large;
  __shared__ unsigned short large[256];
  const unsigned c = c0 + threadIdx.x;
  s_off[threadIdx.x] = jb.offs[min(c, nkeys)];
  if (threadIdx.x == 255) s_off[256] = jb.offs[min(c + 1u, nkeys)];
  if (threadIdx.x == 0) n_large = 0;
  __syncthreads();
  const unsigned blo = s_off[0], B = s_off[256] - blo;
  if (B == 0) return;                                               // halo cells: nothing lives there
  const unsigned lo = s_off[threadIdx.x], pop = s_off[threadIdx.x + 1] - lo;
  unsigned* __restrict__ members = jb.members;
  if (pop > CELLS_SMALL_POP) large[atomicAdd(&n_large, 1u)] = (unsigned short)threadIdx.x;
  if (B <= CELLS_STAGE) {
    for (unsigned i = threadIdx.x; i < B; i += 256u) buf[pad16(i)] = members[blo + i];
    __syncthreads();
    const unsigned base = lo - blo;
    order_list([&](const unsigned k) -> unsigned& { return buf[pad16(base + k)]; }, pop);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < B; i += 256u) members[blo + i] = buf[pad16(i)];
  } else {
    order_list([&](const unsigned k) -> unsigned& { return members[lo + k]; }, pop);
  }
  __syncthreads();
  // crowded cells: stable rank by counting, the whole block per cell
  unsigned* __restrict__ tmp = jb.rank;
  for (unsigned q = 0; q < n_large; ++q) {
    const unsigned t = large[q], l0 = s_off[t], l1 = s_off[t + 1];
    for (unsigned i = l0 + threadIdx.x; i < l1; i += 256u) {
      const unsigned v = members[i];
      unsigned before = 0;
      for (unsigned k = l0; k < l1; ++k) before += unsigned(members[k] < v);
      tmp[l0 + before] = v;
    }
    __syncthreads();
    for (unsigned i = l0 + threadIdx.x; i < l1; i += 256u) members[i] = tmp[i];
    __syncthreads();
  }
}

// Step 5: one thread per destination slot moves the seven streams of its member; destination slots behind the alive
// particles get the dead id.
constexpr int PLACE_SLOTS_PER_THREAD = 4;   // 2 measured 5 % slower (48 registers, fewer loads in flight per thread)
__global__ void __launch_bounds__(256)
k_sort_place(const SortJob* __restrict__ jobs, const unsigned nkeys) {
  const SortJob jb = jobs[blockIdx.y];
  const Species src = jb.src, dst = jb.dst;
  B2P_GLOBAL_SPECIES(src); B2P_GLOBAL_SPECIES(dst); B2P_GLOBAL(jb.members); B2P_GLOBAL(jb.offs);
  const unsigned first = blockIdx.x * (256u * PLACE_SLOTS_PER_THREAD);
  if (first >= src.n) return;
  const unsigned* __restrict__ members = jb.members;
  const unsigned na = jb.offs[nkeys];                              // alive particles
  unsigned e[PLACE_SLOTS_PER_THREAD], p[PLACE_SLOTS_PER_THREAD];
  bool a[PLACE_SLOTS_PER_THREAD];
  float f[PLACE_SLOTS_PER_THREAD][6];
  unsigned long long id[PLACE_SLOTS_PER_THREAD];
#pragma unroll
  for (int r = 0; r < PLACE_SLOTS_PER_THREAD; ++r) {
    e[r] = first + 256u * r + threadIdx.x;
    a[r] = e[r] < na;
    p[r] = a[r] ? members[e[r]] : 0u;
  }
#pragma unroll
  for (int r = 0; r < PLACE_SLOTS_PER_THREAD; ++r) {
    id[r] = DEAD;
    if (a[r]) {
      f[r][0] = src.x[p[r]]; f[r][1] = src.y[p[r]]; f[r][2] = src.z[p[r]];
      f[r][3] = src.ux[p[r]]; f[r][4] = src.uy[p[r]]; f[r][5] = src.uz[p[r]];
      id[r] = src.id[p[r]];
    }
  }
#pragma unroll
  for (int r = 0; r < PLACE_SLOTS_PER_THREAD; ++r) {
    if (a[r]) {
      dst.x[e[r]] = f[r][0]; dst.y[e[r]] = f[r][1]; dst.z[e[r]] = f[r][2];
      dst.ux[e[r]] = f[r][3]; dst.uy[e[r]] = f[r][4]; dst.uz[e[r]] = f[r][5];
      dst.id[e[r]] = id[r];
    } else if (e[r] < src.n) {
      dst.id[e[r]] = DEAD;
    }
  }
}

// ---------------------------------------------------------------- launchers --
static unsigned blocks_for(size_t n) { return unsigned((n + 255) / 256); }

void launch_sort_keys(const Species& s, const Geom& g, const float origo[3], unsigned* keys, unsigned* idx, unsigned dead_key) {
  ProfScope prof_(KC_SORT_KEYS, double(s.n));
  if (!s.n) return;
  k_sort_keys<<<blocks_for(s.n), 256, 0, ctx().stream>>>(s, g, make_float3(origo[0], origo[1], origo[2]), keys, idx, dead_key);
  B2P_LAUNCH_CHECK();
}

size_t sort_pairs_temp_bytes(unsigned n, int end_bit) {
  size_t bytes = 0;
  cub::DoubleBuffer<unsigned> k(nullptr, nullptr), v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, int(n), 0, end_bit, ctx().stream);
  return bytes;
}
// stable LSD radix sort of (key, slot) pairs; returns which half of the double
// buffers holds the result
int sort_pairs(void* temp, size_t temp_bytes, unsigned* keys[2], unsigned* vals[2], unsigned n, int end_bit) {
  ProfScope prof_(KC_RADIX_SORT, double(n));
  cub::DoubleBuffer<unsigned> k(keys[0], keys[1]), v(vals[0], vals[1]);
  B2P_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, int(n), 0, end_bit, ctx().stream));
  count_launch((end_bit + 7) / 8 + 2);
  return v.selector;
}
void launch_gather(const Species& src, const Species& dst, const unsigned* perm) {
  ProfScope prof_(KC_GATHER, double(src.n));
  if (!src.n) return;
  k_gather<<<blocks_for(src.n), 256, 0, ctx().stream>>>(src, dst, perm);
  B2P_LAUNCH_CHECK();
}


// steps 1-2 for the njobs containers of the DEVICE table `jobs` (cnt zeroed by the caller); max_n = the largest container
void launch_sort_count_scan(const SortJob* jobs, int njobs, unsigned max_n, double total_slots, const Geom& g, unsigned nkeys) {
  if (!njobs || !max_n) return;
  {
    ProfScope prof_(KC_SORT_KEYS, total_slots);
    const dim3 grid((max_n + 256 * SORT_SLOTS_PER_THREAD - 1) / (256 * SORT_SLOTS_PER_THREAD), unsigned(njobs));
    k_sort_count<<<grid, 256, 0, ctx().stream>>>(jobs, g, nkeys);
    B2P_LAUNCH_CHECK();
  }
  ProfScope prof_(KC_RADIX_SORT, total_slots);
  const unsigned nchunks = sort_scan_chunks(nkeys);
  k_sort_chunk_sums<<<dim3(nchunks, unsigned(njobs)), 256, 0, ctx().stream>>>(jobs, nkeys, nchunks);
  B2P_LAUNCH_CHECK();
  k_sort_scan_chunks<<<dim3(nchunks, unsigned(njobs)), 256, 0, ctx().stream>>>(jobs, nkeys, nchunks);
  B2P_LAUNCH_CHECK();
}
unsigned sort_scan_chunks(unsigned nkeys) { return (nkeys + 2u + SCAN_CHUNK - 1u) / SCAN_CHUNK; }
// steps 3-4
void launch_sort_scatter_place(const SortJob* jobs, int njobs, unsigned max_n, double total_slots, unsigned nkeys) {
  if (!njobs || !max_n) return;
  const dim3 grid((max_n + 256 * SORT_SLOTS_PER_THREAD - 1) / (256 * SORT_SLOTS_PER_THREAD), unsigned(njobs));
  {
    ProfScope prof_(KC_RADIX_SORT, total_slots);
    k_sort_scatter<<<grid, 256, 0, ctx().stream>>>(jobs, nkeys);
    B2P_LAUNCH_CHECK();
  }
  {
    ProfScope prof_(KC_RADIX_SORT, total_slots);
    k_sort_cells<<<dim3((nkeys + 255u) / 256u, unsigned(njobs)), 256, 0, ctx().stream>>>(jobs, nkeys);
    B2P_LAUNCH_CHECK();
  }
  ProfScope prof_(KC_GATHER, total_slots);
  const dim3 pgrid((max_n + 256 * PLACE_SLOTS_PER_THREAD - 1) / (256 * PLACE_SLOTS_PER_THREAD), unsigned(njobs));
  k_sort_place<<<pgrid, 256, 0, ctx().stream>>>(jobs, nkeys);
  B2P_LAUNCH_CHECK();
}

}  // namespace b2p
