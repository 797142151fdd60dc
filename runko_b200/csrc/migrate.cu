// migrate.cu — pack_outgoing_particles (pic/tile_communication.c++:68-96 -> ParticleContainer::divide_to_subregions,
// pic/particle.c++:199-348), batched over all containers of a phase and ORDERED BY CONSTRUCTION: no key list and no
// sort.  The reference copies the leavers in container order and then stably sorts them by their 27-way subregion;
// the same buffer is produced here in two sweeps over the leaver masks the push published (one uint2 of ballots per
// 32 slots, pmath.cuh publish_masks):
//   1. k_pack_count  one block per segment of PACK_SEG_WORDS mask words: leavers per subregion of that segment
//                    -> seg[sub][segment]; P = 1 + last slot that stays alive (pic/particle.h:469-488)
//   2. k_pack_scan   one block per tile: exclusive scan of seg over (species, subregion, segment) — which IS the
//                    reference's output order — in place; subregion_particle_ends_ (pic/particle.c++:327-343); tile totals
//   3. k_pack_write  same blocks as 1: every leaver goes to base[sub][segment] + (its rank among the segment's
//                    leavers of that subregion, in slot order) as a 32-byte ParticleState (pic/particle.c++:268-291)
//                    and its slot is marked dead (:304-312).
// The host reads the totals once (between 2 and 3) to size the tiles' buffers: one synchronisation per pack.
#include "particles.cuh"
#include "pmath.cuh"

namespace b2p {

constexpr unsigned PACK_SEG_WORDS = 256;                   // mask words (of 32 slots) per block: 8192 slots

unsigned pack_segments(unsigned n_slots) { return (((n_slots + 31u) / 32u) + PACK_SEG_WORDS - 1u) / PACK_SEG_WORDS; }

// exclusive prefix of `v` over the 256 threads of the block
__device__ __forceinline__ unsigned block_excl_256(const unsigned v, unsigned* total) {
  __shared__ unsigned wsum[8];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= unsigned(o)) incl += t; }
  __syncthreads();                                         // wsum may still be read from a previous call
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  unsigned before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { const unsigned s = wsum[w]; if (unsigned(w) < wid) before += s; all += s; }
  *total = all;
  return before + incl - v;
}

__global__ void __launch_bounds__(256)
k_pack_count(const PackJob* __restrict__ jobs) {
  const PackJob jb = jobs[blockIdx.y];
  if (blockIdx.x >= jb.nseg) return;
  B2P_GLOBAL(jb.masks); B2P_GLOBAL(jb.seg); B2P_GLOBAL(jb.last_alive); B2P_GLOBAL_SPECIES(jb.s);
  __shared__ unsigned hist[27];
  __shared__ unsigned s_last;
  if (threadIdx.x < 27) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_last = 0;
  __syncthreads();
  const unsigned w = blockIdx.x * PACK_SEG_WORDS + threadIdx.x;
  uint2 m = make_uint2(0u, 0u);
  if (w < jb.nwords) m = jb.masks[w];
  unsigned bits = m.x;
  while (bits) {
    const unsigned b = __ffs(bits) - 1;
    bits &= bits - 1;
    const unsigned n = w * 32u + b;
    atomicAdd(&hist[subregion_of(jb.s.x[n], jb.s.y[n], jb.s.z[n], jb.mn, jb.mx)], 1u);
  }
  unsigned last = m.y ? w * 32u + (32u - __clz(m.y)) : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if ((threadIdx.x & 31) == 0 && last) atomicMax(&s_last, last);
  __syncthreads();
  if (threadIdx.x < 27) jb.seg[threadIdx.x * jb.nseg + blockIdx.x] = hist[threadIdx.x];
  if (threadIdx.x == 0 && s_last) atomicMax(jb.last_alive, s_last);
}

// one block per tile: its containers are jobs[first .. first + count)
__global__ void __launch_bounds__(256)
k_pack_scan(const PackJob* __restrict__ jobs, const PackTile* __restrict__ tiles) {
  const PackTile tl = tiles[blockIdx.x];
  unsigned run = 0;                                        // leavers of the tile so far (identical in every thread)
  for (unsigned q = 0; q < tl.count; ++q) {
    const PackJob& jb = jobs[tl.first + q];
    const unsigned total = 27u * jb.nseg;                  // seg is [sub][segment]: scanning it linearly is the reference's order
    if (total == 0 && threadIdx.x < 27) jb.ends[threadIdx.x] = run;   // empty container: every span ends where the previous species ended
    for (unsigned base = 0; base < total; base += 2048u) {
      // 8 consecutive entries per thread
      unsigned v[8], sum = 0;
      const unsigned i0 = base + threadIdx.x * 8u;
#pragma unroll
      for (int r = 0; r < 8; ++r) { v[r] = i0 + r < total ? jb.seg[i0 + r] : 0u; sum += v[r]; }
      unsigned chunk_total;
      unsigned at = run + block_excl_256(sum, &chunk_total);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (i0 + r < total) {
          jb.seg[i0 + r] = at;
          at += v[r];
          // the last segment of a subregion closes it: absolute end offset in the tile's buffer
          if ((i0 + r + 1) % jb.nseg == 0) jb.ends[(i0 + r) / jb.nseg] = at;
        }
      }
      run += chunk_total;
    }
  }
  if (threadIdx.x == 0) *tl.total = run;
}

__global__ void __launch_bounds__(256)
k_pack_write(const PackJob* __restrict__ jobs) {
  const PackJob& jb = jobs[blockIdx.y];
  if (blockIdx.x >= jb.nseg) return;
  __shared__ unsigned short l_off[PACK_SEG_WORDS * 32];    // the segment's leavers, in slot order: slot - first slot of the segment
  __shared__ unsigned char l_sub[PACK_SEG_WORDS * 32];
  __shared__ unsigned s_base[27];
  const unsigned w = blockIdx.x * PACK_SEG_WORDS + threadIdx.x;
  const unsigned seg_first = blockIdx.x * PACK_SEG_WORDS * 32u;
  uint2 m = make_uint2(0u, 0u);
  if (w < jb.nwords) m = jb.masks[w];
  unsigned L;
  unsigned pos = block_excl_256(__popc(m.x), &L);
  if (L == 0) return;
  if (threadIdx.x < 27) s_base[threadIdx.x] = jb.seg[threadIdx.x * jb.nseg + blockIdx.x];
  unsigned bits = m.x;
  while (bits) {
    const unsigned b = __ffs(bits) - 1;
    bits &= bits - 1;
    const unsigned n = w * 32u + b;
    l_off[pos] = static_cast<unsigned short>(n - seg_first);
    l_sub[pos] = static_cast<unsigned char>(subregion_of(jb.s.x[n], jb.s.y[n], jb.s.z[n], jb.mn, jb.mx));
    ++pos;
  }
  __syncthreads();
  // one thread per leaver of the compact list (the first L threads: full warps issue the seven gathers)
  double ke = 0.0;
  for (unsigned e = threadIdx.x; e < L; e += 256u) {
    const unsigned sub = l_sub[e], n = seg_first + l_off[e];
    b2p_particle_state st;
    st.pos[0] = jb.s.x[n]; st.pos[1] = jb.s.y[n]; st.pos[2] = jb.s.z[n];
    st.vel[0] = jb.s.ux[n]; st.vel[1] = jb.s.uy[n]; st.vel[2] = jb.s.uz[n];
    st.id = jb.s.id[n];
    unsigned rank = 0;
    for (unsigned q = 0; q < e; ++q) rank += unsigned(l_sub[q] == sub);
    jb.out[s_base[sub] + rank] = st;
    jb.s.id[n] = DEAD;
    if (jb.ke) {
      const V3 v = { st.vel[0], st.vel[1], st.vel[2] };
      ke -= double(sqrtf(1.0f + dot(v, v)) - 1.0f);
    }
  }
  double* const account = jb.ke;
  B2P_GLOBAL(account);
  if (account && ke != 0.0) atomicAdd(account + (threadIdx.x & (KE_SLOTS - 1)), ke);   // the leavers leave the species' account
}

void launch_pack_count_scan(const PackJob* jobs, unsigned ncont, unsigned max_nseg, const PackTile* tiles, unsigned ntiles,
                            double total_slots) {
  if (!ncont || !max_nseg) return;
  {
    ProfScope prof_(KC_DETECT, total_slots);
    for (unsigned c0 = 0; c0 < ncont; c0 += 65535u) {
      k_pack_count<<<dim3(max_nseg, std::min(65535u, ncont - c0)), 256, 0, ctx().stream>>>(jobs + c0);
      B2P_LAUNCH_CHECK();
    }
  }
  ProfScope prof_(KC_DETECT, 0.0);
  k_pack_scan<<<ntiles, 256, 0, ctx().stream>>>(jobs, tiles);
  B2P_LAUNCH_CHECK();
}

void launch_pack_write(const PackJob* jobs, unsigned ncont, unsigned max_nseg, double total_leavers) {
  ProfScope prof_(KC_GATHER_OUT, total_leavers);
  if (!ncont || !max_nseg) return;
  for (unsigned c0 = 0; c0 < ncont; c0 += 65535u) {
    k_pack_write<<<dim3(max_nseg, std::min(65535u, ncont - c0)), 256, 0, ctx().stream>>>(jobs + c0);
    B2P_LAUNCH_CHECK();
  }
}

}  // namespace b2p
