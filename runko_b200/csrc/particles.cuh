// particles.cuh — launchers of the particle kernels (particles.cu).
#pragma once
#include "common.cuh"

namespace b2p {

struct AppendJobHost {        // mirrors particles.cu::AppendJob
  const b2p_particle_state* src;
  unsigned count;
  unsigned dst_offset;
  Species dst;
  float* Jpend;               // destination tile's pending nodal J, or nullptr
  float3 origo;
  float charge;
  double* ke;                 // kinetic-energy account of the container's species (KE_SLOTS doubles), or nullptr
};


// groups of tiles of one geometry handled by one launch of the small per-tile kernels of the
// particle phase (tables passed by value as kernel arguments)
constexpr int PUSH_GROUP_MAX = 32;
struct NodalBatch { const float* E[PUSH_GROUP_MAX]; const float* B[PUSH_GROUP_MAX]; float4* nod[PUSH_GROUP_MAX]; int n; };
struct EdgeBatch { const float4* Jc[PUSH_GROUP_MAX]; float* J[PUSH_GROUP_MAX]; int n; };
size_t nodal_float4_per_node();   // float4 slots of nodal staging per lattice node (layout: particles.cu, k_nodal_means)
void launch_nodal_means(const NodalBatch& bt, const Geom& g);
void launch_edge_gather(const EdgeBatch& bt, const Geom& g);
void launch_nodal_means(const float* E, const float* B, const Geom& g, float4* nod);
// One container of a batched push launch (push.cu).  The push also publishes the leaver / stayer ballots of
// every warp (masks: one uint2 per 32 slots, rounded up to whole blocks of 256 slots) for
// pack_outgoing_particles; Jc != nullptr: fused push + deposit of the particles that stay inside the tile box.
// Kinetic-energy account of a grid (pic/particle.c++:352-377, the per-lap io_average_kinetic_energy): when a push covers
// every container of a grid it leaves, per species, the sum over the alive particles of sqrt(1 + u.u) - 1 of the
// velocities it stored (fp32 per particle, summed pairwise inside a warp, fp64 across warps) in KE_SLOTS spread
// accumulators; pack_outgoing subtracts the leavers, the particle exchange adds the arrivals, so the account keeps
// equalling the sum over the containers and b2p_grid_energies reads 64 KB per species instead of 20 B per slot.  Any other
// change of a container (injection, upload, reflector wall, ...) voids the account until the next whole-grid push.
// The push keeps it (8 % of its time: 8 M double atomics per launch) only if the energies were read since the previous push.
constexpr int KE_SLOTS = 8192;    // 2^13 accumulators per species: a push launch sends 8 M warp sums, 32 slots measured +90 % push time (L2 same-address atomics)
struct PushJob {
  Species s;
  const float4* nod;          // nodal means of the container's tile (k_nodal_means layout)
  float4* Jc;                 // cell-edge accumulators of the tile (fused deposit), or nullptr
  uint2* masks;               // leaver / stayer ballots, one uint2 per 32 slots
  float3 origo, mn, mx;       // lattice origin, tile box (float(mins/maxs))
  float qm, charge;           // sign(q)/m (pic/particle_boris.h:26-27), signed charge (zigzag)
  double* ke;                 // kinetic-energy account of the species (KE_SLOTS doubles) or nullptr: see KE_SLOTS
};
// The job table travels as a kernel argument (CUDA >= 12.1 allows 32 KB of parameters): no upload, and the
// compiler knows that every pointer in it addresses global memory.
constexpr int PUSH_JOBS_MAX = 64;
struct PushJobs { PushJob job[PUSH_JOBS_MAX]; };
// the first njobs entries: containers of one geometry / pusher / cfl; max_n = the largest container among them
void launch_push_jobs(int pusher, const PushJobs& jobs, int njobs, unsigned max_n, double total_slots, const Geom& g, float cfl,
                      bool fuse);
void launch_deposit(const Species& s, float4* Jc, const Geom& g, const float origo[3], float cfl, float charge);
void launch_edge_gather(const float4* Jc, float* J, const Geom& g);
void launch_sort_keys(const Species& s, const Geom& g, const float origo[3], unsigned* keys, unsigned* idx, unsigned dead_key);
size_t sort_pairs_temp_bytes(unsigned n, int end_bit);
int sort_pairs(void* temp, size_t temp_bytes, unsigned* keys[2], unsigned* vals[2], unsigned n, int end_bit);
void launch_gather(const Species& src, const Species& dst, const unsigned* perm);
// batched counting sort by cell key (sort.cu): one container of a batch
struct SortJob {
  Species src, dst;           // dst: spare storage of the same capacity (swapped with the container afterwards)
  unsigned* keys;             // [src.n] cell key per slot (dead -> nkeys)
  unsigned* rank;             // [src.n] arrival rank inside the cell
  unsigned* members;          // [src.n] slots of every cell, cell by cell
  unsigned* cnt;              // [nkeys + 2] per-key populations (zeroed by the caller)
  unsigned* chunk_sums;       // [sort_scan_chunks(nkeys) + 1] scan scratch (last entry zeroed by the caller)
  unsigned* offs;             // [nkeys + 2] exclusive scan of cnt; offs[nkeys] = alive particles
  unsigned* max_pop;          // largest population of a cell
  float3 origo;
};
constexpr unsigned SORT_RADIX_POP = 4096;    // containers whose last known largest cell exceeds this take the radix sort
unsigned sort_scan_chunks(unsigned nkeys);
void launch_sort_count_scan(const SortJob* jobs, int njobs, unsigned max_n, double total_slots, const Geom& g, unsigned nkeys);
void launch_sort_scatter_place(const SortJob* jobs, int njobs, unsigned max_n, double total_slots, unsigned nkeys);
// batched, ordered pack_outgoing_particles (migrate.cu)
struct PackJob {              // one container
  const uint2* masks;         // leaver / stayer ballots per 32 slots
  unsigned nwords, nseg;      // mask words; segments of the container (pack_segments)
  Species s;
  float3 mn, mx;              // tile box (float(mins/maxs), pic/tile_communication.c++:71-79)
  unsigned* seg;              // [27][nseg]: leavers per (subregion, segment), then their bases in the tile's buffer
  unsigned* last_alive;       // P = 1 + last slot that stays alive (zeroed by the caller)
  unsigned long long* ends;   // [27] subregion_particle_ends_ of this species (absolute offsets in the tile's buffer)
  b2p_particle_state* out;    // the tile's outgoing buffer (filled in for the write pass)
  double* ke;                 // kinetic-energy account of the species (leavers are subtracted), or nullptr
};
struct PackTile { unsigned first, count; unsigned* total; };   // a tile's containers in the job table; its leaver total
unsigned pack_segments(unsigned n_slots);
void launch_pack_count_scan(const PackJob* jobs, unsigned ncont, unsigned max_nseg, const PackTile* tiles, unsigned ntiles, double total_slots);
void launch_pack_write(const PackJob* jobs, unsigned ncont, unsigned max_nseg, double total_leavers);
void launch_make_masks(const Species& s, uint2* masks, const float mins[3], const float maxs[3]);
void launch_last_alive(const unsigned long long* id, unsigned n, unsigned* last_alive);
void launch_append(const void* jobs, int njobs, unsigned max_count, bool wrap, const float wmin[3], const float wmax[3],
                   const Geom& g, float cfl);
void launch_selfcheck_divc(const float* x, unsigned long long n, float c, float* out, float* ref);
void launch_fill_dead(unsigned long long* id, unsigned begin, unsigned end);
void launch_kinetic_energy(const Species& s, double* out);
// one launch for many containers: *energy += sum over alive of (gamma - 1), *alive += number of alive slots (either may be null)
struct EnergyJob { Species s; double* energy; unsigned long long* alive; };
void launch_kinetic_energy_batch(const EnergyJob* jobs, int njobs, unsigned max_n, double total_slots);
void launch_count_alive(const Species& s, unsigned long long* out);   // *out += number of alive slots
// FieldsWriter<3>::pack_tile density (io/snapshots/mpiio_fields.c++:277-316)
void launch_snapshot_density(const Species& s, const float mins[3], float inv_stride, int nxt, int nyt, int nzt, float* n_out);
// ParticleContainer::reflect_at_wall (pic/reflector_wall.c++:126-222); corrJ = nodal correction lattice (3*Ch floats)
void launch_reflect_at_wall(const Species& s, float* corrJ, const Geom& g, const float origo[3], float cfl, float walloc,
                            float betawall, float gammawall, float charge);
void launch_inject_thermal(const Species& s, const Geom& g, const float mins[3], unsigned ppc, float theta,
                           unsigned long long seed_pos, unsigned long long seed_vel, unsigned long long id_base);
void launch_inject_drifting(const Species& s, unsigned first, const Geom& g, const float mins[3], int i0, int i1, unsigned ppc, float theta,
                            float Gamma, float dir, unsigned long long seed_pos, unsigned long long seed_vel, unsigned long long id_base);
}  // namespace b2p
