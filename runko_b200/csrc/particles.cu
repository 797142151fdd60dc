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
  double* ke;                 // kinetic-energy account of the species (arrivals are added), or nullptr
};

// ParticleContainer::append (pic/particle.h:511-571): AoS -> SoA after the last
// alive particle; with `wrap` the global periodic wrap
// (x<0 ? max : min) + fmodf(x, L) of :534-549.
// Arrivals of a tile whose stayers were deposited by the fused push add their zigzag current
// to that tile's pending J here (scalar atomics: arrivals are ~1% of the particles).
__global__ void __launch_bounds__(256)
k_append(const AppendJob* __restrict__ jobs, const int wrap, const float3 wmin, const float3 wmax, const Geom g, const float cfl) {
  const AppendJob jb = jobs[blockIdx.y];
  B2P_GLOBAL(jb.src); B2P_GLOBAL(jb.Jpend); B2P_GLOBAL(jb.ke); B2P_GLOBAL_SPECIES(jb.dst);
  const float Lx = wmax.x - wmin.x, Ly = wmax.y - wmin.y, Lz = wmax.z - wmin.z;
  double ke = 0.0;
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
    if (jb.ke && st.id != DEAD) {
      const V3 v = { st.vel[0], st.vel[1], st.vel[2] };
      ke += double(sqrtf(1.0f + dot(v, v)) - 1.0f);
    }
  }
  if (jb.ke && ke != 0.0) atomicAdd(jb.ke + ((blockIdx.x * blockDim.x + threadIdx.x) & (KE_SLOTS - 1)), ke);
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

// The same sum for many containers in one launch (blockIdx.y = job; the grid-wide diagnostics of every lap,
// io_average_kinetic_energy): 12 + 8 B per slot, HBM-bound.
constexpr unsigned KE_CHUNK = 16384;      // slots per block: 256 threads x 4 slots (one LDG.128 per stream) x 16 rounds
__global__ void __launch_bounds__(256)
k_kinetic_energy_batch(const EnergyJob* __restrict__ jobs) {
  const EnergyJob jb = jobs[blockIdx.y];
  const Species s = jb.s;
  B2P_GLOBAL_SPECIES(s); B2P_GLOBAL(jb.energy); B2P_GLOBAL(jb.alive);
  const unsigned first = blockIdx.x * KE_CHUNK;
  if (first >= s.n) return;
  const unsigned end = min(first + KE_CHUNK, s.n);
  double acc = 0.0;
  unsigned cnt = 0;
  auto one = [&](const float ux, const float uy, const float uz, const unsigned long long id) {
    const V3 v = { ux, uy, uz };
    const float e = sqrtf(1.0f + dot(v, v)) - 1.0f;
    if (id != DEAD) { acc += double(e); ++cnt; }
  };
  // the streams are 256-byte aligned and `first` is a multiple of 4: four slots per thread and round as 128-bit loads
  const unsigned full = first + ((end - first) & ~3u);
  for (unsigned n = first + 4u * threadIdx.x; n < full; n += 1024u) {
    const float4 ux = *reinterpret_cast<const float4*>(s.ux + n), uy = *reinterpret_cast<const float4*>(s.uy + n),
                 uz = *reinterpret_cast<const float4*>(s.uz + n);
    const ulonglong2 i0 = *reinterpret_cast<const ulonglong2*>(s.id + n), i1 = *reinterpret_cast<const ulonglong2*>(s.id + n + 2);
    one(ux.x, uy.x, uz.x, i0.x); one(ux.y, uy.y, uz.y, i0.y); one(ux.z, uy.z, uz.z, i1.x); one(ux.w, uy.w, uz.w, i1.y);
  }
  for (unsigned n = full + threadIdx.x; n < end; n += 256u) one(s.ux[n], s.uy[n], s.uz[n], s.id[n]);
  for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  __shared__ double sh[8];
  __shared__ unsigned sc[8];
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = acc; sc[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    unsigned long long c = 0;
    for (int q = 0; q < 8; ++q) { t += sh[q]; c += sc[q]; }
    if (jb.energy) atomicAdd(jb.energy, t);
    if (jb.alive) atomicAdd(jb.alive, c);
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

// Bench-only generator for the drifting plasmas of the beam and shock workloads (no reference equivalent): ppc
// particles per cell in the cells [i0, i1) x all j x all k of the tile (pic::Tile::batch_inject_in_x_stripe's cell
// range and i -> j -> k order, pic/tile.c++:235-322), written to the slots [first, first + ppc * cells).  Momentum: a
// Maxwellian of spread sqrt(theta) in the rest frame, flux-weighted flip and boost by Gamma along +-x
// (runko/sample_thermal_distributions.py:58-127 for theta <= 0.2, statistically — not bitwise).
__global__ void __launch_bounds__(256)
k_inject_drifting(const Species s, const unsigned first, const Geom g, const float3 mins, const int i0, const int i1, const unsigned ppc,
                  const float theta, const float Gamma, const float dir, const unsigned long long seed_pos,
                  const unsigned long long seed_vel, const unsigned long long id_base) {
  const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned ncell = unsigned(i1 - i0) * g.N[1] * g.N[2];
  if (p >= ncell * ppc) return;
  const unsigned cell = p % ncell;
  const unsigned k = cell % g.N[2], j = (cell / g.N[2]) % g.N[1], i = unsigned(i0) + cell / (g.N[2] * g.N[1]);
  const unsigned long long gcell = (static_cast<unsigned long long>(i) * g.N[1] + j) * g.N[2] + k;
  const unsigned long long key = (gcell * 64ull + p / ncell) * 0xD6E8FEB86659FD93ull;
  Rng rp{ mix64(seed_pos ^ key) };
  const unsigned n = first + p;
  s.x[n] = mins.x + float(i) + rp.uniform() * 0.999999f;
  s.y[n] = mins.y + float(j) + rp.uniform() * 0.999999f;
  s.z[n] = mins.z + float(k) + rp.uniform() * 0.999999f;
  Rng rv{ mix64(seed_vel ^ key) };
  const float sig = sqrtf(theta);
  const float r1 = sqrtf(-2.0f * logf(rv.uniform())), a1 = 6.2831853f * rv.uniform();
  const float r2 = sqrtf(-2.0f * logf(rv.uniform())), a2 = 6.2831853f * rv.uniform();
  float ux = sig * r1 * cosf(a1);
  const float uy = sig * r1 * sinf(a1), uz = sig * r2 * cosf(a2);
  const float gam = sqrtf(1.0f + ux * ux + uy * uy + uz * uz);
  const float beta = sqrtf(fmaxf(0.0f, 1.0f - 1.0f / (Gamma * Gamma)));
  if (-beta * ux / gam > rv.uniform()) ux = -ux;
  s.ux[n] = dir * Gamma * (ux + beta * gam);
  s.uy[n] = uy;
  s.uz[n] = uz;
  s.id[n] = id_base + p;
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

void launch_make_masks(const Species& s, uint2* masks, const float mins[3], const float maxs[3]) {
  ProfScope prof_(KC_DETECT, double(s.n));
  if (!s.n) return;
  k_make_masks<<<blocks_for(s.n), 256, 0, ctx().stream>>>(s, masks, make_float3(mins[0], mins[1], mins[2]),
                                                         make_float3(maxs[0], maxs[1], maxs[2]));
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

// njobs containers of the DEVICE table `jobs`; max_n = the largest container
void launch_kinetic_energy_batch(const EnergyJob* jobs, int njobs, unsigned max_n, double total_slots) {
  ProfScope prof_(KC_ENERGY, total_slots);
  if (!njobs || !max_n) return;
  const unsigned per_job = (max_n + KE_CHUNK - 1) / KE_CHUNK;
  k_kinetic_energy_batch<<<dim3(per_job, unsigned(njobs)), 256, 0, ctx().stream>>>(jobs);
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

void launch_inject_drifting(const Species& s, unsigned first, const Geom& g, const float mins[3], int i0, int i1, unsigned ppc, float theta,
                            float Gamma, float dir, unsigned long long seed_pos, unsigned long long seed_vel, unsigned long long id_base) {
  ProfScope prof_(KC_OTHER, 0.0);
  const size_t total = size_t(std::max(0, i1 - i0)) * g.N[1] * g.N[2] * ppc;
  if (!total) return;
  k_inject_drifting<<<blocks_for(total), 256, 0, ctx().stream>>>(s, first, g, make_float3(mins[0], mins[1], mins[2]), i0, i1, ppc, theta, Gamma,
                                                                  dir, seed_pos, seed_vel, id_base);
  B2P_LAUNCH_CHECK();
}

}  // namespace b2p
