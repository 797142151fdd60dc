// particles.cu — particle kernels: nodal-mean staging + interpolation + the three
// relativistic pushers, zigzag current deposition, cell-key sort, leaver
// detection / migration packing, append with periodic wrap, kinetic energy,
// synthetic thermal injection.
//
// Arithmetic contract: fp32, no FMA contraction (-fmad=false), IEEE div/sqrt
// (nvcc defaults -prec-div=true -prec-sqrt=true), operation order of the
// reference (citations per function) => push, keys and migration lists are
// bit-identical to the reference's unfused CPU build.
#include "particles.cuh"
#include "pmath.cuh"


#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

namespace b2p {


// ------------------------------------------------------- nodal field means --
// The interpolator's per-corner staggered averages
// (emf/yee_lattice_interpolate_linear_1st.h:91-113) depend only on the lattice
// node (ii,jj,kk), not on the particle, so they are computed once per tile into
// two node-major arrays, nodA[n] = {Ex,Ey,Ez,Bx} (float4) and nodB[n] = {By,Bz} (float2, stored
// right behind nodA: 24 B/node): a particle then needs 8 corners x (LDG.128 + LDG.64) instead of
// 144 scalar gathers, and the L1 data pipe — the unit that bounds the push — returns 192 B per
// lane instead of 256.  Same operands and same association (2-term sum /2; 4-term right fold /4)
// => same bits.
// staggered means of one lattice node; zero where a neighbour would fall outside the lattice
__device__ __forceinline__ void node_means(const float* __restrict__ E, const float* __restrict__ B, const Geom& g, const int i,
                                           const int j, const int k, float4& a, float2& b) {
  a = make_float4(0.f, 0.f, 0.f, 0.f);
  b = make_float2(0.f, 0.f);
  if (i < 1 || j < 1 || k < 1 || k >= g.Hx[2]) return;
  const size_t sj = g.Hx[2], si = size_t(g.Hx[1]) * g.Hx[2], Ch = g.Ch;
  const size_t n = (size_t(i) * g.Hx[1] + j) * g.Hx[2] + k;
  const float* Ex = E; const float* Ey = E + Ch; const float* Ez = E + 2 * Ch;
  const float* Bx = B; const float* By = B + Ch; const float* Bz = B + 2 * Ch;
  a.x = (Ex[n - si] + Ex[n]) / 2.0f;
  a.y = (Ey[n - sj] + Ey[n]) / 2.0f;
  a.z = (Ez[n - 1] + Ez[n]) / 2.0f;
  a.w = (Bx[n] + (Bx[n - sj] + (Bx[n - 1] + Bx[n - sj - 1]))) / 4.0f;
  b.x = (By[n] + (By[n - si] + (By[n - 1] + By[n - si - 1]))) / 4.0f;
  b.y = (Bz[n] + (Bz[n - si] + (Bz[n - sj] + Bz[n - si - sj]))) / 4.0f;
}

// nodal staging layout (float4 units per tile, n = node index): A[n] = {Ex,Ey,Ez,Bx} at nod[n],
// {By,Bz}[n] as float2 right behind the Ch float4s (24 B per node)
size_t nodal_float4_per_node() { return 2; }

__global__ void __launch_bounds__(256)
k_nodal_means(const NodalBatch bt, const Geom g) {
  const int tile = blockIdx.y;
  const float* __restrict__ E = bt.E[tile];
  const float* __restrict__ B = bt.B[tile];
  float4* __restrict__ nod = bt.nod[tile];
  const int kblocks = (g.Hx[2] + 31) / 32;
  const int k = (blockIdx.x % kblocks) * blockDim.x + threadIdx.x;
  const int j = (blockIdx.x / kblocks) * blockDim.y + threadIdx.y;
  if (k >= g.Hx[2] || j >= g.Hx[1]) return;
  for (int i = blockIdx.z; i < g.Hx[0]; i += gridDim.z) {
    const size_t n = (size_t(i) * g.Hx[1] + j) * g.Hx[2] + k;
    float4 a;
    float2 b;
    node_means(E, B, g, i, j, k, a, b);
    nod[n] = a;
    reinterpret_cast<float2*>(nod + g.Ch)[n] = b;
  }
}

// ---------------------------------------------------------------- deposit --
struct DepositArgs {
  int agg_min;     // fewest folding lanes for which a warp aggregation step pays (reduce_runs_and_red)
  Species s;
  float4* Jc;      // cell-edge accumulators: 3 float4 per lattice cell (see below)
  Geom g;
  float3 origo;
  float cfl;
  float charge;
};


// Standalone deposit of a whole container: one thread per particle.
template <int AGG, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_deposit_zigzag(const DepositArgs a) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool alive = n < a.s.n && a.s.id[n] != DEAD;
  Zigzag z;
  if (alive)
    z = zigzag_split(V3{ a.s.x[n], a.s.y[n], a.s.z[n] }, V3{ a.s.ux[n], a.s.uy[n], a.s.uz[n] }, a.origo, a.cfl, a.charge, a.g);
  deposit_split<AGG>(alive, z, a.Jc, a.agg_min);
}


// ------------------------------------------------------------ edge gather --
// Fold the cell-edge records into the nodal current and write ALL of J (this is also
// the reference's clear_current + `J += generated_J`, pic/tile.c++:371,405):
//   Jx[i,j,k] = c(i,j,k).x0 + c(i,j-1,k).x1 + c(i,j,k-1).x2 + c(i,j-1,k-1).x3   etc.
__global__ void __launch_bounds__(256)
k_edge_gather(const EdgeBatch bt, const Geom g) {
  const int tile = blockIdx.y;
  const float4* __restrict__ Jc = bt.Jc[tile];
  float* __restrict__ J = bt.J[tile];
  const int kblocks = (g.Hx[2] + 31) / 32;
  const int k = (blockIdx.x % kblocks) * blockDim.x + threadIdx.x;
  const int j = (blockIdx.x / kblocks) * blockDim.y + threadIdx.y;
  if (k >= g.Hx[2] || j >= g.Hx[1]) return;
  for (int i = blockIdx.z; i < g.Hx[0]; i += gridDim.z) {
  const long sj = g.Hx[2], si = long(g.Hx[1]) * g.Hx[2];
  const long n = (long(i) * g.Hx[1] + j) * g.Hx[2] + k;
  const bool pi = i > 0, pj = j > 0, pk = k > 0;
  float jx = Jc[3 * n + 0].x, jy = Jc[3 * n + 1].x, jz = Jc[3 * n + 2].x;
  if (pj) jx += Jc[3 * (n - sj) + 0].y;
  if (pk) jx += Jc[3 * (n - 1) + 0].z;
  if (pj && pk) jx += Jc[3 * (n - sj - 1) + 0].w;
  if (pi) jy += Jc[3 * (n - si) + 1].y;
  if (pk) jy += Jc[3 * (n - 1) + 1].z;
  if (pi && pk) jy += Jc[3 * (n - si - 1) + 1].w;
  if (pi) jz += Jc[3 * (n - si) + 2].y;
  if (pj) jz += Jc[3 * (n - sj) + 2].z;
  if (pi && pj) jz += Jc[3 * (n - si - sj) + 2].w;
  J[n] = jx; J[size_t(g.Ch) + n] = jy; J[2 * size_t(g.Ch) + n] = jz;
  }
}

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

// ------------------------------------------------------- counting sort (fast path) --
// The contract of ParticleContainer::sort is "stable sort by cell key, dead slots last"
// (pic/particle.h:575-703).  Keys are lattice cell indices < Ch with a few tens of particles per
// key, so instead of a general radix sort of (key, slot) pairs the container is sorted by counting:
//   1. k_sort_count   key[n], rank[n] = arrival order among the particles of that key (one atomic
//                     on cnt[key] per run of equal keys inside a warp; NOT in slot order yet)
//   2. exclusive scan of cnt -> offs (CUB DeviceScan over Ch + 2 counters) and the largest
//                     population of a cell (a hint for the NEXT sort of this container)
//   3. k_sort_scatter members[offs[key] + rank] = n: the slots of every cell, in arrival order
//   4. k_sort_fix     one thread per cell puts its (short, almost ordered) member list into
//                     ascending slot order by insertion; cells with more than SORT_THREAD_POP
//                     members are queued for k_sort_fix_big (one block per cell, rank by counting)
//   5. k_gather       dst[p] = src[members[p]]: coalesced writes, reads within a few hundred slots
//                     of p for a container that was sorted a few laps ago.
// Result: exactly the stable order for any input.  Dead slots (key Ch, clamped like the radix
// path) keep their arrival order — the contents of dead slots are unspecified in the reference.
// Step 4 is quadratic in the population of a cell, so the host routes containers whose last known
// largest cell population exceeds SORT_RADIX_POP to the general radix sort instead.
__global__ void __launch_bounds__(256)
k_sort_count(const Species s, const Geom g, const float3 origo, unsigned* __restrict__ keys, unsigned* __restrict__ rank,
             unsigned* __restrict__ cnt, const unsigned dead_key) {
  // two slots per thread, 256 apart: both slots' loads, then both atomics, are in flight together
  const unsigned lane = threadIdx.x & 31;
  unsigned n[2], key[2], start[2], base[2];
  bool in[2];
  unsigned long long id[2];
  float px[2], py[2], pz[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    n[r] = blockIdx.x * 512u + 256u * r + threadIdx.x;
    in[r] = n[r] < s.n;
    id[r] = DEAD; px[r] = py[r] = pz[r] = 0.f;
    if (in[r]) { id[r] = ld_pinned(s.id + n[r]); px[r] = ld_pinned(s.x + n[r]); py[r] = ld_pinned(s.y + n[r]); pz[r] = ld_pinned(s.z + n[r]); }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    key[r] = dead_key;
    if (id[r] != DEAD) {
      const unsigned i = __float2uint_rz(px[r] - origo.x);
      const unsigned j = __float2uint_rz(py[r] - origo.y);
      const unsigned k = __float2uint_rz(pz[r] - origo.z);
      key[r] = (i * unsigned(g.Hx[1]) + j) * unsigned(g.Hx[2]) + k;
      if (key[r] > dead_key) key[r] = dead_key;
    }
    // one atomic per run of equal keys in the warp (a container sorted a few laps ago is made of such runs)
    const unsigned prev = __shfl_up_sync(0xffffffffu, key[r], 1);
    const bool head = lane == 0 || key[r] != prev || !in[r];
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    start[r] = 31u - __clz(hm & (0xFFFFFFFFu >> (31u - lane)));                 // my run's first lane
    const unsigned above = lane == 31 ? 0u : (hm >> (lane + 1));
    const unsigned len = above ? unsigned(__ffs(above)) : 32u - lane;            // for a head: length of its run
    // Dead slots are not ranked: they end up behind the alive particles in any order (k_sort_gather).
    base[r] = 0;
    if (head && key[r] != dead_key) base[r] = atomicAdd(&cnt[key[r]], len);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const unsigned b0 = __shfl_sync(0xffffffffu, base[r], start[r]);
    if (in[r]) {
      keys[n[r]] = key[r];
      rank[n[r]] = b0 + (lane - start[r]);
    }
  }
}

// largest population among the alive keys [0, nkeys)
__global__ void __launch_bounds__(256)
k_max_count(const unsigned* __restrict__ cnt, const unsigned nkeys, unsigned* __restrict__ out) {
  unsigned m = 0;
  for (unsigned c = blockIdx.x * blockDim.x + threadIdx.x; c < nkeys; c += gridDim.x * blockDim.x) m = max(m, cnt[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__global__ void __launch_bounds__(256)
k_sort_scatter(const unsigned* __restrict__ keys, const unsigned* __restrict__ rank, const unsigned* __restrict__ offs,
               unsigned* __restrict__ members, const unsigned n_total, const unsigned dead_key) {
  const unsigned n0 = blockIdx.x * 512u + threadIdx.x, n1 = n0 + 256u;   // two slots per thread: both lookups in flight together
  const bool i0 = n0 < n_total, i1 = n1 < n_total;
  const unsigned k0 = i0 ? keys[n0] : dead_key, k1 = i1 ? keys[n1] : dead_key;
  const unsigned r0 = i0 ? rank[n0] : 0u, r1 = i1 ? rank[n1] : 0u;
  const unsigned o0 = k0 != dead_key ? offs[k0] : 0u, o1 = k1 != dead_key ? offs[k1] : 0u;
  if (k0 != dead_key) members[o0 + r0] = n0;
  if (k1 != dead_key) members[o1 + r1] = n1;
}

// Step 5: dst[p] = src[members[p]] for the offs[dead_key] alive particles; the slots behind them are
// dead.  SORT_GATHER_SLOTS slots per thread (256 apart) keep 7 x SORT_GATHER_SLOTS independent gathers in flight.
constexpr int SORT_GATHER_SLOTS = 4;
__global__ void __launch_bounds__(256)
k_sort_gather(const Species src, const Species dst, const unsigned* __restrict__ members, const unsigned* __restrict__ n_alive) {
  const unsigned na = *n_alive;
  unsigned n[SORT_GATHER_SLOTS], p[SORT_GATHER_SLOTS];
  bool a[SORT_GATHER_SLOTS];
  float f[SORT_GATHER_SLOTS][6];
  unsigned long long id[SORT_GATHER_SLOTS];
#pragma unroll
  for (int r = 0; r < SORT_GATHER_SLOTS; ++r) {
    n[r] = blockIdx.x * (256u * SORT_GATHER_SLOTS) + 256u * r + threadIdx.x;
    a[r] = n[r] < na;
    p[r] = a[r] ? members[n[r]] : 0u;
  }
#pragma unroll
  for (int r = 0; r < SORT_GATHER_SLOTS; ++r) {
    id[r] = DEAD;
    if (a[r]) {
      f[r][0] = src.x[p[r]]; f[r][1] = src.y[p[r]]; f[r][2] = src.z[p[r]];
      f[r][3] = src.ux[p[r]]; f[r][4] = src.uy[p[r]]; f[r][5] = src.uz[p[r]];
      id[r] = src.id[p[r]];
    }
  }
#pragma unroll
  for (int r = 0; r < SORT_GATHER_SLOTS; ++r) {
    if (a[r]) {
      dst.x[n[r]] = f[r][0]; dst.y[n[r]] = f[r][1]; dst.z[n[r]] = f[r][2];
      dst.ux[n[r]] = f[r][3]; dst.uy[n[r]] = f[r][4]; dst.uz[n[r]] = f[r][5];
    }
    if (n[r] < src.n) dst.id[n[r]] = id[r];
  }
}

// Step 4: ascending slot order inside every alive cell.  A block takes blockDim.x (<= 256)
// consecutive cells — one contiguous piece of `members` — and stages the piece and the cells'
// offsets in shared memory with coalesced loads.  One thread per cell labels its members with the
// cell's local index; then one thread per MEMBER flags its cell if it sits behind a larger slot
// index, and the members of flagged cells count the members of their cell with a smaller slot
// index (the lanes of a warp read at most a few distinct shared-memory words per step:
// broadcasts) and write themselves to their stable position.  Cells with more than
// SORT_THREAD_POP members, and all cells of a piece that does not fit the staging buffer, are
// queued in `big` = {count, cells...} for k_sort_fix_big.
constexpr unsigned SORT_FIX_STAGE = 8192;   // entries (32 KB + 8 KB of labels)
__global__ void __launch_bounds__(256)
k_sort_fix(const unsigned* __restrict__ offs, unsigned* __restrict__ members, const unsigned nkeys, unsigned* __restrict__ big,
           unsigned* __restrict__ max_pop) {
  __shared__ unsigned sm[SORT_FIX_STAGE];
  __shared__ unsigned char cellof[SORT_FIX_STAGE];
  __shared__ unsigned so[257];
  __shared__ unsigned char unsorted[256];
  const unsigned c0 = blockIdx.x * blockDim.x, nc = min(blockDim.x, nkeys - c0);
  for (unsigned t = threadIdx.x; t <= nc; t += blockDim.x) so[t] = offs[c0 + t];
  unsorted[threadIdx.x] = 0;
  __syncthreads();
  const unsigned lo0 = so[0], total = so[nc] - lo0;
  const bool staged = total <= SORT_FIX_STAGE;
  {   // largest cell population of the container (the hint for its next sort)
    unsigned m = threadIdx.x < nc ? so[threadIdx.x + 1] - so[threadIdx.x] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 1) atomicMax(max_pop, m);
  }
  if (threadIdx.x < nc) {
    const unsigned lo = so[threadIdx.x] - lo0, hi = so[threadIdx.x + 1] - lo0;
    if (hi - lo >= 2 && (!staged || hi - lo > SORT_THREAD_POP)) big[1 + atomicAdd(big, 1u)] = c0 + threadIdx.x;
    if (staged)
      for (unsigned t = lo; t < hi; ++t) cellof[t] = static_cast<unsigned char>(threadIdx.x);
  }
  if (!staged) return;
  for (unsigned t = threadIdx.x; t < total; t += blockDim.x) sm[t] = members[lo0 + t];
  __syncthreads();
  for (unsigned e = threadIdx.x; e < total; e += blockDim.x) {
    const unsigned a = cellof[e];
    if (e > so[a] - lo0 && sm[e - 1] > sm[e]) unsorted[a] = 1;
  }
  __syncthreads();
  for (unsigned e = threadIdx.x; e < total; e += blockDim.x) {
    const unsigned a = cellof[e];
    if (!unsorted[a]) continue;
    const unsigned lo = so[a] - lo0, hi = so[a + 1] - lo0;
    if (hi - lo > SORT_THREAD_POP) continue;
    const unsigned v = sm[e];
    unsigned before = 0;
    for (unsigned q = lo; q < hi; ++q) before += unsigned(sm[q] < v);
    if (lo + before != e) members[lo0 + lo + before] = v;
  }
}

// Queued cells: one warp per cell, rank by counting.  Up to 32 members sit one per lane and are
// ranked over shuffles; up to 128 sit four per lane and are ranked against the list re-read
// through L1 (warp-uniform addresses); larger cells go through `tmp` (>= container size; the
// arrival ranks are no longer needed).  All ranks are known before the first member is rewritten.
__global__ void __launch_bounds__(256)
k_sort_fix_big(const unsigned* __restrict__ offs, unsigned* __restrict__ members, unsigned* __restrict__ tmp,
               const unsigned* __restrict__ big) {
  const unsigned nbig = big[0];
  const unsigned lane = threadIdx.x & 31;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < nbig; e += nwarps) {
    const unsigned c = big[1 + e];
    const unsigned lo = offs[c], pop = offs[c + 1] - lo;
    if (pop <= 32) {
      const unsigned v = lane < pop ? members[lo + lane] : 0xFFFFFFFFu;
      unsigned before = 0;
      for (unsigned q = 0; q < pop; ++q) before += unsigned(__shfl_sync(0xffffffffu, v, int(q)) < v);
      if (lane < pop && before != lane) members[lo + before] = v;
    } else if (pop <= 128) {
      unsigned v[4], before[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r] = lane + 32u * r < pop ? members[lo + lane + 32u * r] : 0xFFFFFFFFu;
      for (unsigned q = 0; q < pop; ++q) {
        const unsigned w = members[lo + q];
#pragma unroll
        for (int r = 0; r < 4; ++r) before[r] += unsigned(w < v[r]);
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (lane + 32u * r < pop) members[lo + before[r]] = v[r];
    } else {
      for (unsigned t = lane; t < pop; t += 32) {
        const unsigned v = members[lo + t];
        unsigned before = 0;
        for (unsigned q = 0; q < pop; ++q) before += unsigned(members[lo + q] < v);
        tmp[lo + before] = v;
      }
      __syncwarp();
      for (unsigned t = lane; t < pop; t += 32) members[lo + t] = tmp[lo + t];
    }
    __syncwarp();
  }
}

// -------------------------------------------------------------- migration --
// standalone mask pass (containers whose masks are stale: injected / uploaded / appended)
__global__ void __launch_bounds__(256)
k_make_masks(const Species s, uint2* __restrict__ masks, const float3 mn, const float3 mx) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool alive = n < s.n && s.id[n] != DEAD;
  float x = 0.f, y = 0.f, z = 0.f;
  if (alive) { x = s.x[n]; y = s.y[n]; z = s.z[n]; }
  publish_masks(alive, inside_box(x, y, z, mn, mx), n, masks);
}

struct CollectJob {           // one container
  const uint2* masks;
  unsigned nwords;
  Species s;
  float3 mn, mx;
};

// Mask words of all containers -> key list (container << 37) | (subregion << 32) | slot,
// per-container leaver counts and P = 1 + the largest slot that stays alive
// (ParticleContainer::append, pic/particle.h:469-488).  One thread per mask word, one
// list-position atomic per block.  blockIdx.y = container.
__global__ void __launch_bounds__(256)
k_collect_leavers(const CollectJob* __restrict__ jobs, const unsigned first_container, unsigned long long* __restrict__ list,
                  unsigned* __restrict__ list_count, const unsigned list_cap, unsigned* __restrict__ last_alive,
                  unsigned* __restrict__ cont_count) {
  const unsigned c = first_container + blockIdx.y;
  const CollectJob jb = jobs[c];
  __shared__ unsigned sh_cnt[8], sh_base, sh_last;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (unsigned w0 = blockIdx.x * blockDim.x; w0 < jb.nwords; w0 += gridDim.x * blockDim.x) {
    const unsigned w = w0 + threadIdx.x;
    uint2 m = make_uint2(0u, 0u);
    if (w < jb.nwords) m = jb.masks[w];
    const unsigned cnt = __popc(m.x);
    // exclusive scan of cnt over the block
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= unsigned(o)) incl += t; }
    unsigned last = m.y ? w * 32u + (32u - __clz(m.y)) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    if (threadIdx.x == 0) sh_last = 0;
    if (lane == 31) sh_cnt[wid] = incl;
    __syncthreads();
    if (lane == 0 && last) atomicMax(&sh_last, last);
    if (threadIdx.x == 0) {
      unsigned total = 0;
      for (int q = 0; q < 8; ++q) { const unsigned v = sh_cnt[q]; sh_cnt[q] = total; total += v; }
      unsigned base = 0;
      if (total) { base = atomicAdd(list_count, total); atomicAdd(&cont_count[c], total); }
      sh_base = base;
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh_last) atomicMax(&last_alive[c], sh_last);
    unsigned pos = sh_base + sh_cnt[wid] + incl - cnt;
    unsigned bits = m.x;
    while (bits) {
      const unsigned b = __ffs(bits) - 1;
      bits &= bits - 1;
      const unsigned n = w * 32u + b;
      const int sub = subregion_of(jb.s.x[n], jb.s.y[n], jb.s.z[n], jb.mn, jb.mx);
      if (pos < list_cap)
        list[pos] = (static_cast<unsigned long long>(c) << 37) | (static_cast<unsigned long long>(sub) << 32) | n;
      ++pos;
    }
    __syncthreads();
  }
}

struct OutTile {              // per container: where its leavers go
  b2p_particle_state* buf;    // tile's outgoing AoS buffer
  unsigned long long base;    // index of this tile's first entry in the sorted list
  Species s;
};

// Sorted leaver list -> AoS ParticleState (pic/particle.c++:268-291) + mark the
// source slots dead (:304-312) + per-(container, subregion) counts (:336-346).
__global__ void __launch_bounds__(256)
k_gather_outgoing(const unsigned long long* __restrict__ sorted, const unsigned total, const OutTile* __restrict__ cont,
                  unsigned* __restrict__ counts /*[ncont][27]*/) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const unsigned long long key = sorted[q];
  const unsigned c = unsigned(key >> 37), sub = unsigned(key >> 32) & 31u, n = unsigned(key);
  const OutTile t = cont[c];
  b2p_particle_state st;
  st.pos[0] = t.s.x[n]; st.pos[1] = t.s.y[n]; st.pos[2] = t.s.z[n];
  st.vel[0] = t.s.ux[n]; st.vel[1] = t.s.uy[n]; st.vel[2] = t.s.uz[n];
  st.id = t.s.id[n];
  t.buf[q - t.base] = st;
  t.s.id[n] = DEAD;
  // the list is sorted by (container, subregion): one count atomic per run inside the warp
  const unsigned grp = unsigned(key >> 32);
  const unsigned peers = __match_any_sync(__activemask(), grp);
  if ((threadIdx.x & 31) == unsigned(__ffs(peers) - 1)) atomicAdd(&counts[c * 27 + sub], unsigned(__popc(peers)));
}

// Full-pass fallback for P = 1 + last alive slot (pic/particle.h:469-488), used
// when no detection pass has produced it (inject outside the lap).
__global__ void __launch_bounds__(256)
k_last_alive(const unsigned long long* __restrict__ id, const unsigned n_total, unsigned* __restrict__ last_alive) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool alive = n < n_total && id[n] != DEAD;
  const unsigned m = __ballot_sync(0xffffffffu, alive);
  if (m && (threadIdx.x & 31) == 0) atomicMax(last_alive, (n & ~31u) + (32 - __clz(m)));
}

struct AppendJob {            // one incoming span (pic/tile_communication.c++:133-181)
  const b2p_particle_state* src;
  unsigned count;
  unsigned dst_offset;        // P + exclusive scan of span sizes (pic/particle.h:490-509)
  Species dst;
  float* Jpend;               // != nullptr: the destination tile's pending nodal J (fused push+deposit)
  float3 origo;
  float charge;
};

// ParticleContainer::append (pic/particle.h:511-571): AoS -> SoA after the last
// alive particle; with `wrap` the global periodic wrap
// (x<0 ? max : min) + fmodf(x, L) of :534-549.
// Arrivals of a tile whose stayers were deposited by the fused push add their zigzag current
// to that tile's pending J here (scalar atomics: arrivals are ~1% of the particles).
__global__ void __launch_bounds__(256)
k_append(const AppendJob* __restrict__ jobs, const int wrap, const float3 wmin, const float3 wmax, const Geom g, const float cfl) {
  const AppendJob jb = jobs[blockIdx.y];
  const float Lx = wmax.x - wmin.x, Ly = wmax.y - wmin.y, Lz = wmax.z - wmin.z;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < jb.count; i += gridDim.x * blockDim.x) {
    const b2p_particle_state st = jb.src[i];
    const unsigned j = jb.dst_offset + i;
    V3 p = { st.pos[0], st.pos[1], st.pos[2] };
    if (wrap) {
      p.x = (st.pos[0] < 0 ? wmax.x : wmin.x) + fmodf(st.pos[0], Lx);
      p.y = (st.pos[1] < 0 ? wmax.y : wmin.y) + fmodf(st.pos[1], Ly);
      p.z = (st.pos[2] < 0 ? wmax.z : wmin.z) + fmodf(st.pos[2], Lz);
    }
    jb.dst.x[j] = p.x; jb.dst.y[j] = p.y; jb.dst.z[j] = p.z;
    jb.dst.ux[j] = st.vel[0]; jb.dst.uy[j] = st.vel[1]; jb.dst.uz[j] = st.vel[2];
    jb.dst.id[j] = st.id;
    if (jb.Jpend && st.id != DEAD)
      deposit_split_nodal(zigzag_split(p, V3{ st.vel[0], st.vel[1], st.vel[2] }, jb.origo, cfl, jb.charge, g), jb.Jpend, g);
  }
}

__global__ void __launch_bounds__(256)
k_fill_dead(unsigned long long* __restrict__ id, const unsigned begin, const unsigned end) {
  const unsigned n = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (n < end) id[n] = DEAD;
}

// pic/particle.c++:352-377: Σ alive (sqrt(1+u·u) − 1), fp32 per particle, fp64 sum.
__global__ void __launch_bounds__(256)
k_kinetic_energy(const Species s, double* __restrict__ out) {
  double acc = 0.0;
  for (unsigned n = blockIdx.x * blockDim.x + threadIdx.x; n < s.n; n += gridDim.x * blockDim.x) {
    const V3 v = { s.ux[n], s.uy[n], s.uz[n] };
    const float e = sqrtf(1.0f + dot(v, v)) - 1.0f;
    if (s.id[n] != DEAD) acc += double(e);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int q = 0; q < int(blockDim.x >> 5); ++q) t += sh[q];
    atomicAdd(out, t);
  }
}

// ----------------------------------------------------------- reflector wall --
// pic/reflector_wall.c++:35-118: zigzag deposit of one sub-trajectory x1 -> x2 (lattice-local
// coordinates) into the nodal correction lattice.  Same operands and association as the
// reference; the scalar atomics land in arbitrary order (stated deposit tolerance).
__device__ __forceinline__ void zigzag_deposit_single(float* __restrict__ J, const Geom& g, const V3 x1, const V3 x2,
                                                      const float charge) {
  const V3 fi1 = { floorf(x1.x), floorf(x1.y), floorf(x1.z) };
  const V3 fi2 = { floorf(x2.x), floorf(x2.y), floorf(x2.z) };
  auto relay = [](const float f1, const float f2, const float p1, const float p2) {
    const float a = (f1 < f2 ? f1 : f2) + 1.0f;
    const float m = f1 > f2 ? f1 : f2;
    const float h = 0.5f * (p1 + p2);
    const float b = m > h ? m : h;
    return a < b ? a : b;
  };
  const V3 xr = { relay(fi1.x, fi2.x, x1.x, x2.x), relay(fi1.y, fi2.y, x1.y, x2.y), relay(fi1.z, fi2.z, x1.z, x2.z) };
  const V3 F1 = charge * (xr - x1);
  const V3 F2 = charge * (x2 - xr);
  const V3 W1 = 0.5f * (x1 + xr) - fi1;
  const V3 W2 = 0.5f * (x2 + xr) - fi2;
  const unsigned Hy = unsigned(g.Hx[1]), Hz = unsigned(g.Hx[2]);
  Zigzag z;
  z.n1 = (__float2uint_rz(fi1.x) * Hy + __float2uint_rz(fi1.y)) * Hz + __float2uint_rz(fi1.z);
  z.n2 = (__float2uint_rz(fi2.x) * Hy + __float2uint_rz(fi2.y)) * Hz + __float2uint_rz(fi2.z);
  const float one = 1.0f;
#define EDGES(F, W, ex, ey, ez)                                                                                \
  ex = make_float4(F.x * (one - W.y) * (one - W.z), F.x * W.y * (one - W.z), F.x * (one - W.y) * W.z, F.x * W.y * W.z); \
  ey = make_float4(F.y * (one - W.x) * (one - W.z), F.y * W.x * (one - W.z), F.y * (one - W.x) * W.z, F.y * W.x * W.z); \
  ez = make_float4(F.z * (one - W.x) * (one - W.y), F.z * W.x * (one - W.y), F.z * (one - W.x) * W.y, F.z * W.x * W.y);
  EDGES(F1, W1, z.ax, z.ay, z.az)
  EDGES(F2, W2, z.bx, z.by, z.bz)
#undef EDGES
  deposit_split_nodal(z, J, g);
}

// ParticleContainer::reflect_at_wall (pic/reflector_wall.c++:126-222), one thread per slot.  The
// reference evaluates every alive particle branch-free with 0/1 float masks; the same expressions
// are evaluated here (positions and velocities are bit-identical), and only the two correction
// deposits — which the reference multiplies by mask_refl = 0 for everything but the reflected
// particles — are skipped when the mask is zero (adding +-0 does not change the lattice).
__global__ void __launch_bounds__(256)
k_reflect_at_wall(const Species s, float* __restrict__ corrJ, const Geom g, const float3 origo_, const float c,
                  const float walloc, const float betawall, const float gammawall, const float charge) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= s.n || s.id[n] == DEAD) return;
  const float EPS = 1e-10f;
  const float walloc0 = walloc - betawall * c;
  const V3 origo = { origo_.x, origo_.y, origo_.z };
  const V3 pos_1 = { s.x[n], s.y[n], s.z[n] };
  const V3 u = { s.ux[n], s.uy[n], s.uz[n] };
  const float gam = sqrtf(1.0f + dot(u, u));
  const float invgam = 1.0f / gam;
  const V3 pos_0 = pos_1 - c * invgam * u;
  const float mask_skip = (pos_1.x >= walloc) ? 1.0f : 0.0f;
  const float mask_close = (walloc0 - pos_0.x <= c) ? 1.0f : 0.0f;
  const float denom = betawall * c - c * u.x * invgam;
  const float dt = fabsf((pos_0.x - walloc0) / (denom + EPS));
  const float mask_crossed = (dt <= 1.0f) ? 1.0f : 0.0f;
  const float mask_refl = (1.0f - mask_skip) * mask_close * mask_crossed;
  const float mask_park = (1.0f - mask_skip) - mask_refl;
  const V3 pos_col = pos_0 + c * dt * invgam * u;
  const float ux_new = gammawall * gammawall * gam * (2.0f * betawall - u.x * invgam * (1.0f + betawall * betawall));
  const V3 u_new = { ux_new, u.y, u.z };
  const float gam_new = sqrtf(1.0f + dot(u_new, u_new));
  const float invgam_new = 1.0f / gam_new;
  const float ratio = fabsf((pos_1.x - pos_col.x) / (pos_1.x - pos_0.x + EPS));
  const float dt_refl = 1.0f < ratio ? 1.0f : ratio;
  const V3 pos_refl = pos_col + c * dt_refl * invgam_new * u_new;
  if (mask_refl != 0.0f) {
    const V3 p1l = pos_1 - origo;
    const V3 dep_fwd_from = p1l + mask_refl * (pos_0 - pos_1);
    const V3 dep_fwd_to = p1l + mask_refl * (pos_col - pos_1);
    zigzag_deposit_single(corrJ, g, dep_fwd_from, dep_fwd_to, mask_refl * charge);
    const V3 x1_deposit = pos_refl - c * invgam_new * u_new;
    const V3 dep_rev_from = p1l + mask_refl * (x1_deposit - pos_1);
    const V3 dep_rev_to = p1l + mask_refl * (pos_col - pos_1);
    zigzag_deposit_single(corrJ, g, dep_rev_from, dep_rev_to, mask_refl * (-charge));
  }
  s.x[n] = (1.0f - mask_refl) * pos_1.x + mask_refl * pos_refl.x;
  s.y[n] = (1.0f - mask_refl) * pos_1.y + mask_refl * pos_refl.y;
  s.z[n] = (1.0f - mask_refl) * pos_1.z + mask_refl * pos_refl.z;
  s.ux[n] = (1.0f - mask_refl) * u.x + mask_refl * ux_new;
  if (mask_park > 0.5f) s.id[n] = DEAD;
}

// FieldsWriter<3>::pack_tile, density part (io/snapshots/mpiio_fields.c++:277-316): alive particles per
// coarse cell, float atomics (exact for counts < 2^24)
__global__ void __launch_bounds__(256)
k_snapshot_density(const Species s, const float3 mins, const float inv_stride, const int nxt, const int nyt, const int nzt,
                   float* __restrict__ n_out) {
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= s.n || s.id[n] == DEAD) return;
  const float fx = floorf((s.x[n] - mins.x) * inv_stride), fy = floorf((s.y[n] - mins.y) * inv_stride),
              fz = floorf((s.z[n] - mins.z) * inv_stride);
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < float(nxt) && fy < float(nyt) && fz < float(nzt))) return;
  atomicAdd(&n_out[(size_t(fz) * nyt + size_t(fy)) * nxt + size_t(fx)], 1.0f);
}

// number of alive slots (id != dead) of a container, added to *out
__global__ void __launch_bounds__(256)
k_count_alive(const Species s, unsigned long long* __restrict__ out) {
  unsigned c = 0;
  for (unsigned n = blockIdx.x * blockDim.x + threadIdx.x; n < s.n; n += gridDim.x * blockDim.x) c += unsigned(s.id[n] != DEAD);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, static_cast<unsigned long long>(c));
}

// ------------------------------------------------- synthetic thermal plasma --
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct Rng {
  unsigned long long state;
  __device__ float uniform() {   // (0,1)
    state = mix64(state);
    return (float(unsigned(state >> 40)) + 0.5f) * (1.0f / 16777216.0f);
  }
};

// Bench-only generator (no reference equivalent): slot p = round*Ncells + cell,
// cells ordered i->j->k like pic::Tile::batch_inject_in_x_stripe (pic/tile.c++:264-275);
// position = cell corner + U[0,1)^3 (the same position for every species, as in
// projects/pic-turbulence/pic.py:141-156); momentum = Sobol sampling of a
// Juttner-Synge distribution for theta > 0.2, Maxwellian below
// (runko/sample_thermal_distributions.py:58-127, statistically — not bitwise).
__global__ void __launch_bounds__(256)
k_inject_thermal(const Species s, const Geom g, const float3 mins, const unsigned ppc, const float theta,
                 const unsigned long long seed_pos, const unsigned long long seed_vel, const unsigned long long id_base) {
  const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned ncell = unsigned(g.N[0]) * g.N[1] * g.N[2];
  if (p >= ncell * ppc) return;
  const unsigned cell = p % ncell;
  const unsigned k = cell % g.N[2], j = (cell / g.N[2]) % g.N[1], i = cell / (g.N[2] * g.N[1]);
  Rng rp{ mix64(seed_pos ^ (static_cast<unsigned long long>(p) * 0xD6E8FEB86659FD93ull)) };
  s.x[p] = mins.x + float(i) + rp.uniform() * 0.999999f;
  s.y[p] = mins.y + float(j) + rp.uniform() * 0.999999f;
  s.z[p] = mins.z + float(k) + rp.uniform() * 0.999999f;
  Rng rv{ mix64(seed_vel ^ (static_cast<unsigned long long>(p) * 0xD6E8FEB86659FD93ull)) };
  float umag;
  if (theta > 0.2f) {
    for (;;) {
      const float x4 = rv.uniform(), x5 = rv.uniform(), x6 = rv.uniform(), x7 = rv.uniform();
      const float uu = -theta * logf(x4 * x5 * x6);
      const float eta = -theta * logf(x4 * x5 * x6 * x7);
      if (eta * eta - uu * uu > 1.0f) { umag = uu; break; }
    }
    const float mu = 2.0f * rv.uniform() - 1.0f, phi = 6.2831853f * rv.uniform();
    const float st = sqrtf(fmaxf(0.0f, 1.0f - mu * mu));
    s.ux[p] = umag * st * cosf(phi); s.uy[p] = umag * st * sinf(phi); s.uz[p] = umag * mu;
  } else {
    const float sig = sqrtf(theta);
    const float r1 = sqrtf(-2.0f * logf(rv.uniform())), a1 = 6.2831853f * rv.uniform();
    const float r2 = sqrtf(-2.0f * logf(rv.uniform())), a2 = 6.2831853f * rv.uniform();
    s.ux[p] = sig * r1 * cosf(a1); s.uy[p] = sig * r1 * sinf(a1); s.uz[p] = sig * r2 * cosf(a2);
  }
  s.id[p] = id_base + p;
}

__global__ void __launch_bounds__(256)
k_selfcheck_divc(const float* __restrict__ x, const unsigned long long n, const float c, float* __restrict__ out, float* __restrict__ ref) {
  const DivC d(c);
  for (unsigned long long i = (blockIdx.x * blockDim.x + threadIdx.x) * 3ull; i < n; i += gridDim.x * blockDim.x * 3ull) {
    const V3 v = { x[i], i + 1 < n ? x[i + 1] : 1.0f, i + 2 < n ? x[i + 2] : 1.0f };
    const V3 q = d(v);
    const V3 r = v / c;
    out[i] = q.x; ref[i] = r.x;
    if (i + 1 < n) { out[i + 1] = q.y; ref[i + 1] = r.y; }
    if (i + 2 < n) { out[i + 2] = q.z; ref[i + 2] = r.z; }
  }
}
void launch_selfcheck_divc(const float* x, unsigned long long n, float c, float* out, float* ref) {
  k_selfcheck_divc<<<1184, 256, 0, ctx().stream>>>(x, n, c, out, ref);
  B2P_LAUNCH_CHECK();
}

// ---------------------------------------------------------------- launchers --
static unsigned blocks_for(size_t n) { return unsigned((n + 255) / 256); }

void launch_nodal_means(const NodalBatch& bt, const Geom& g) {
  ProfScope prof_(KC_NODAL, double(g.Ch) * bt.n);
  if (!bt.n) return;
  const dim3 grid(((g.Hx[2] + 31) / 32) * ((g.Hx[1] + 7) / 8), bt.n, std::min(g.Hx[0], 32768));
  k_nodal_means<<<grid, dim3(32, 8, 1), 0, ctx().stream>>>(bt, g);
  B2P_LAUNCH_CHECK();
}
void launch_nodal_means(const float* E, const float* B, const Geom& g, float4* nod) {
  NodalBatch bt{};
  bt.E[0] = E; bt.B[0] = B; bt.nod[0] = nod; bt.n = 1;
  launch_nodal_means(bt, g);
}

void launch_deposit(const Species& s, float4* Jc, const Geom& g, const float origo[3], float cfl, float charge) {
  ProfScope prof_(KC_DEPOSIT, double(s.n));
  if (!s.n) return;
  DepositArgs a{ tuning().agg_min, s, Jc, g, make_float3(origo[0], origo[1], origo[2]), cfl, charge };
  const unsigned nb = blocks_for(s.n);
  const int minb = tuning().deposit_minb, agg = tuning().deposit_agg;
  if (agg) {
    if (minb >= 8) k_deposit_zigzag<1, 8><<<nb, 256, 0, ctx().stream>>>(a);
    else if (minb >= 6) k_deposit_zigzag<1, 6><<<nb, 256, 0, ctx().stream>>>(a);
    else k_deposit_zigzag<1, 4><<<nb, 256, 0, ctx().stream>>>(a);
  } else {
    if (minb >= 8) k_deposit_zigzag<0, 8><<<nb, 256, 0, ctx().stream>>>(a);
    else if (minb >= 6) k_deposit_zigzag<0, 6><<<nb, 256, 0, ctx().stream>>>(a);
    else k_deposit_zigzag<0, 4><<<nb, 256, 0, ctx().stream>>>(a);
  }
  B2P_LAUNCH_CHECK();
}

void launch_edge_gather(const EdgeBatch& bt, const Geom& g) {
  ProfScope prof_(KC_EDGE_GATHER, double(g.Ch) * bt.n);
  if (!bt.n) return;
  const dim3 grid(((g.Hx[2] + 31) / 32) * ((g.Hx[1] + 7) / 8), bt.n, std::min(g.Hx[0], 32768));
  k_edge_gather<<<grid, dim3(32, 8, 1), 0, ctx().stream>>>(bt, g);
  B2P_LAUNCH_CHECK();
}
void launch_edge_gather(const float4* Jc, float* J, const Geom& g) {
  EdgeBatch bt{};
  bt.Jc[0] = Jc; bt.J[0] = J; bt.n = 1;
  launch_edge_gather(bt, g);
}

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
size_t sort_keys64_temp_bytes(unsigned n, int end_bit) {
  size_t bytes = 0;
  cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, k, int(n), 0, end_bit, ctx().stream);
  return bytes;
}
int sort_keys64(void* temp, size_t temp_bytes, unsigned long long* keys[2], unsigned n, int end_bit) {
  ProfScope prof_(KC_RADIX_SORT, double(n));
  cub::DoubleBuffer<unsigned long long> k(keys[0], keys[1]);
  B2P_CUDA(cub::DeviceRadixSort::SortKeys(temp, temp_bytes, k, int(n), 0, end_bit, ctx().stream));
  count_launch((end_bit + 7) / 8 + 2);
  return k.selector;
}

size_t scan_temp_bytes(unsigned n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, static_cast<unsigned*>(nullptr), static_cast<unsigned*>(nullptr), int(n), ctx().stream);
  return bytes;
}
// steps 1-2 of the counting sort: keys, arrival ranks, offs = exclusive scan of the per-key
// populations (nkeys alive keys + the dead key + one pad entry), *max_pop = largest alive population
void launch_sort_count_scan(const Species& s, const Geom& g, const float origo[3], unsigned* keys, unsigned* rank, unsigned* cnt,
                            unsigned* offs, unsigned nkeys, void* temp, size_t temp_bytes, unsigned* max_pop, bool with_max) {
  if (!s.n) return;
  {
    ProfScope prof_(KC_SORT_KEYS, double(s.n));
    B2P_CUDA(cudaMemsetAsync(cnt, 0, (size_t(nkeys) + 2) * sizeof(unsigned), ctx().stream));
    B2P_CUDA(cudaMemsetAsync(max_pop, 0, sizeof(unsigned), ctx().stream));
    k_sort_count<<<(s.n + 511) / 512, 256, 0, ctx().stream>>>(s, g, make_float3(origo[0], origo[1], origo[2]), keys, rank, cnt, nkeys);
    B2P_LAUNCH_CHECK();
  }
  ProfScope prof_(KC_RADIX_SORT, double(s.n));
  B2P_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, cnt, offs, int(nkeys + 2), ctx().stream));
  count_launch(1);
  if (with_max) {
    k_max_count<<<std::min(blocks_for(nkeys), 296u), 256, 0, ctx().stream>>>(cnt, nkeys, max_pop);
    B2P_LAUNCH_CHECK();
  }
}
// steps 3-5.  `cnt` (nkeys + 2 counters, free after the scan) becomes the queue of crowded cells,
// `rank` (free after the scatter) the scratch of k_sort_fix_big.
void launch_sort_scatter_place(const Species& src, const Species& dst, const unsigned* keys, unsigned* rank,
                               const unsigned* offs, unsigned* members, unsigned* cnt, unsigned nkeys, unsigned* max_pop) {
  if (!src.n) return;
  {
    ProfScope prof_(KC_RADIX_SORT, double(src.n));
    k_sort_scatter<<<(src.n + 511) / 512, 256, 0, ctx().stream>>>(keys, rank, offs, members, src.n, nkeys);
    B2P_LAUNCH_CHECK();
    B2P_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned), ctx().stream));
    // cells per block: about half the staging buffer at the container's mean population
    const unsigned mean = std::max(1u, src.n / std::max(1u, nkeys));
    unsigned cells = 256;
    while (cells > 32 && cells * mean > SORT_FIX_STAGE / 2) cells >>= 1;
    k_sort_fix<<<(nkeys + cells - 1) / cells, cells, 0, ctx().stream>>>(offs, members, nkeys, cnt, max_pop);
    B2P_LAUNCH_CHECK();
    k_sort_fix_big<<<592, 256, 0, ctx().stream>>>(offs, members, rank, cnt);
    B2P_LAUNCH_CHECK();
  }
  ProfScope prof_(KC_GATHER, double(src.n));
  k_sort_gather<<<(src.n + 256 * SORT_GATHER_SLOTS - 1) / (256 * SORT_GATHER_SLOTS), 256, 0, ctx().stream>>>(src, dst, members, offs + nkeys);
  B2P_LAUNCH_CHECK();
}

void launch_gather(const Species& src, const Species& dst, const unsigned* perm) {
  ProfScope prof_(KC_GATHER, double(src.n));
  if (!src.n) return;
  k_gather<<<blocks_for(src.n), 256, 0, ctx().stream>>>(src, dst, perm);
  B2P_LAUNCH_CHECK();
}

void launch_make_masks(const Species& s, uint2* masks, const float mins[3], const float maxs[3]) {
  ProfScope prof_(KC_DETECT, double(s.n));
  if (!s.n) return;
  k_make_masks<<<blocks_for(s.n), 256, 0, ctx().stream>>>(s, masks, make_float3(mins[0], mins[1], mins[2]),
                                                         make_float3(maxs[0], maxs[1], maxs[2]));
  B2P_LAUNCH_CHECK();
}

void launch_collect_leavers(const void* jobs, unsigned ncont, unsigned max_words, unsigned long long* list, unsigned* list_count,
                            unsigned list_cap, unsigned* last_alive, unsigned* cont_count) {
  ProfScope prof_(KC_DETECT, double(max_words) * 32.0 * ncont);
  if (!ncont || !max_words) return;
  const unsigned bx = std::min(blocks_for(max_words), 2048u);
  for (unsigned c0 = 0; c0 < ncont; c0 += 65535u) {
    const unsigned nc = std::min(65535u, ncont - c0);
    k_collect_leavers<<<dim3(bx, nc), 256, 0, ctx().stream>>>(static_cast<const CollectJob*>(jobs), c0, list, list_count, list_cap,
                                                              last_alive, cont_count);
    B2P_LAUNCH_CHECK();
  }
}

void launch_gather_outgoing(const unsigned long long* sorted, unsigned total, const void* out_tiles, unsigned* counts) {
  ProfScope prof_(KC_GATHER_OUT, double(total));
  if (!total) return;
  k_gather_outgoing<<<blocks_for(total), 256, 0, ctx().stream>>>(sorted, total, static_cast<const OutTile*>(out_tiles), counts);
  B2P_LAUNCH_CHECK();
}

void launch_last_alive(const unsigned long long* id, unsigned n, unsigned* last_alive) {
  ProfScope prof_(KC_OTHER, double(n));
  if (!n) return;
  k_last_alive<<<blocks_for(n), 256, 0, ctx().stream>>>(id, n, last_alive);
  B2P_LAUNCH_CHECK();
}

void launch_append(const void* jobs, int njobs, unsigned max_count, bool wrap, const float wmin[3], const float wmax[3],
                   const Geom& g, float cfl) {
  ProfScope prof_(KC_APPEND, double(max_count));
  if (!njobs || !max_count) return;
  const unsigned bx = std::min(blocks_for(max_count), 1024u);
  k_append<<<dim3(bx, njobs), 256, 0, ctx().stream>>>(static_cast<const AppendJob*>(jobs), wrap ? 1 : 0,
                                                      make_float3(wmin[0], wmin[1], wmin[2]), make_float3(wmax[0], wmax[1], wmax[2]),
                                                      g, cfl);
  B2P_LAUNCH_CHECK();
}

void launch_fill_dead(unsigned long long* id, unsigned begin, unsigned end) {
  ProfScope prof_(KC_OTHER, 0.0);
  if (end <= begin) return;
  k_fill_dead<<<blocks_for(end - begin), 256, 0, ctx().stream>>>(id, begin, end);
  B2P_LAUNCH_CHECK();
}

void launch_kinetic_energy(const Species& s, double* out) {
  ProfScope prof_(KC_ENERGY, double(s.n));
  if (!s.n) return;
  const unsigned nb = std::min(blocks_for(s.n), unsigned(ctx().sm_count) * 8);
  k_kinetic_energy<<<nb, 256, 0, ctx().stream>>>(s, out);
  B2P_LAUNCH_CHECK();
}

void launch_reflect_at_wall(const Species& s, float* corrJ, const Geom& g, const float origo[3], float cfl, float walloc,
                            float betawall, float gammawall, float charge) {
  ProfScope prof_(KC_OTHER, double(s.n));
  if (!s.n) return;
  k_reflect_at_wall<<<blocks_for(s.n), 256, 0, ctx().stream>>>(s, corrJ, g, make_float3(origo[0], origo[1], origo[2]), cfl, walloc,
                                                              betawall, gammawall, charge);
  B2P_LAUNCH_CHECK();
}

void launch_count_alive(const Species& s, unsigned long long* out) {
  ProfScope prof_(KC_OTHER, double(s.n));
  if (!s.n) return;
  k_count_alive<<<std::min(blocks_for(s.n), unsigned(ctx().sm_count) * 8), 256, 0, ctx().stream>>>(s, out);
  B2P_LAUNCH_CHECK();
}
void launch_snapshot_density(const Species& s, const float mins[3], float inv_stride, int nxt, int nyt, int nzt, float* n_out) {
  ProfScope prof_(KC_OTHER, double(s.n));
  if (!s.n) return;
  k_snapshot_density<<<blocks_for(s.n), 256, 0, ctx().stream>>>(s, make_float3(mins[0], mins[1], mins[2]), inv_stride, nxt, nyt, nzt, n_out);
  B2P_LAUNCH_CHECK();
}

void launch_inject_thermal(const Species& s, const Geom& g, const float mins[3], unsigned ppc, float theta,
                           unsigned long long seed_pos, unsigned long long seed_vel, unsigned long long id_base) {
  ProfScope prof_(KC_OTHER, 0.0);
  const size_t total = size_t(g.N[0]) * g.N[1] * g.N[2] * ppc;
  if (!total) return;
  k_inject_thermal<<<blocks_for(total), 256, 0, ctx().stream>>>(s, g, make_float3(mins[0], mins[1], mins[2]), ppc, theta,
                                                                 seed_pos, seed_vel, id_base);
  B2P_LAUNCH_CHECK();
}

}  // namespace b2p
