// host.cuh — host-side objects behind the C-ABI handles (b2p_tile, b2p_grid).
#pragma once
#include <array>
#include <functional>
#include <memory>

#include "common.cuh"
#include "fields.cuh"
#include "particles.cuh"

namespace b2p {

// One particle container (pic::ParticleContainer, pic/particle.h:55-90) on the device.
struct Container {
  DBuf<float> x, y, z, ux, uy, uz;
  DBuf<unsigned long long> id;
  unsigned n = 0;             // size, dead or alive
  double charge = 0, mass = 1;
  bool P_valid = false;       // P = 1 + last alive slot known on the host
  unsigned P = 0;
  DBuf<uint2> masks;          // per 32 slots: {leaving, staying} ballots of the tile box test (particles.cu)
  bool masks_valid = false;   // masks describe the current container contents
  int pop_hint_slot = -1;     // page-locked slot with the largest cell population of the last counting sort (host.cu)
  int radix_sorts_since_probe = 0;
  void touch() { P_valid = false; masks_valid = false; }   // contents changed
  uint2* mask_words();        // sized for the current n (whole blocks of 256 slots)
  Species view() const { return Species{ x.p, y.p, z.p, ux.p, uy.p, uz.p, id.p, n }; }
  size_t capacity() const { return id.cap; }
  void reserve(size_t cap, bool exact = false);   // keeps the first n slots; !exact rounds up to a capacity class
};

struct CommPlan;              // multi-GPU exchange plan (comm.cu)

}  // namespace b2p

struct b2p_tile {
  b2p_config cfg;
  int idx[3];
  double mins[3], maxs[3];
  float origo[3];             // float(mins) - 3  (pic/tile.c++:329-332)
  b2p::Geom g;
  float stencilM[3][3][5];
  b2p::DBuf<float> E, B, Jbuf[2];
  int jcur = 0;
  // fused push+deposit: Jbuf[1 - jcur] holds the current of the particles that stayed in the tile
  // (plus, after the particle exchange, of the arrivals); deposit_current adopts it when the
  // leavers of that same push were removed by pack_outgoing_particles, else it deposits afresh
  bool pendJ_valid = false, pend_packed = false;
  unsigned long long ke_epoch = 0;   // phase_push_particles: detects a tile listed twice in one push
  b2p::DBuf<b2p::FieldPtrs> d_fp;   // 1-entry device tile table for per-tile launches
  bool fp_dirty = true;
  std::vector<b2p::Container> sp;
  unsigned long long tile_tag = 0;
  std::vector<unsigned long long> next_ordinal;
  b2p::DBuf<b2p_particle_state> out_buf;      // subregion_particle_buff_ (pic/tile.h:85)
  std::vector<unsigned long long> out_ends;   // subregion_particle_ends_ (pic/tile.h:84)
  unsigned long long out_count = 0;
  b2p_grid* grid = nullptr;
  int slot = -1;
  bool deferred = false;                      // queued in the pending batch of per-tile calls (host.cu)
  // antenna modes (emf/tile.h:48); lap_coeffs are consumed in order
  struct Antenna { double A[3], wave[3]; int kind; bool has_coeffs; std::vector<std::array<double, 2>> coeffs; size_t next = 0; };
  std::vector<Antenna> antennas;
  // pic-shock boundary pieces
  std::vector<b2p_edge_bc> edge_bcs;          // emf/tile.h:51
  std::vector<b2p_reflector_wall> walls;      // pic/tile.h:71
  b2p::DBuf<float> corrJ;                     // reflector_correction_J_ (pic/tile.h:72), 3*Ch floats
  bool corr_pending = false;                  // reflector_correction_pending_ (pic/tile.h:73)

  float* J() { return Jbuf[jcur].p; }
  b2p::FieldPtrs ptrs() { return b2p::FieldPtrs{ E.p, B.p, J() }; }
  const b2p::FieldPtrs* device_entry();       // uploads when dirty
  size_t lattice_floats() const { return size_t(3) * g.Ch; }
};

struct b2p_grid {
  b2p_config cfg;
  b2p::Geom g;
  std::vector<b2p_tile*> tiles;               // local tiles, in add order
  std::vector<int> slot_of_cid;               // cid -> local slot or -1
  b2p::DBuf<b2p::FieldPtrs> d_tiles;
  b2p::DBuf<int> d_nbr;
  bool table_dirty = true, nbr_dirty = true;
  b2p::CommPlan* comm = nullptr;              // owned; freed in ~b2p_grid (comm.cu)
  int rank = 0, nranks = 1;
  size_t n_boundary = 0;                      // multi-rank grids: tiles[0, n_boundary) have a Moore neighbour on another rank (comm_init orders them first)
  std::vector<int> owner;                     // cid -> rank
  // kinetic-energy account (particles.cuh: KE_SLOTS): [species][KE_SLOTS] doubles, valid between a push of every
  // container of the grid and the next change of a container that the account does not follow
  b2p::DBuf<double> ke_acc;
  bool ke_valid = false;
  bool ke_asked = false;                      // b2p_grid_energies was called since the last whole-grid push: the next push keeps the account

  ~b2p_grid();
  const b2p::FieldPtrs* device_table();
  const int* device_nbr();
  int cid(int i, int j, int k) const { return i + cfg.n_tiles[0] * (j + cfg.n_tiles[1] * k); }
};

namespace b2p {
// phase implementations shared by the per-tile and the batched grid entry points
void phase_push_half_b(const std::vector<b2p_tile*>& tiles, const FieldPtrs* table, int times = 1);
void phase_push_e(const std::vector<b2p_tile*>& tiles, const FieldPtrs* table, bool add_current);
void phase_add_current(const std::vector<b2p_tile*>& tiles, const FieldPtrs* table);
void phase_filter(const std::vector<b2p_tile*>& tiles);
// The first n_first tiles are pushed, then `between` runs on the library stream, then the rest (used to hide the B halo
// exchange under the pushes of the tiles that have no remote neighbour).
void phase_push_particles(const std::vector<b2p_tile*>& tiles, size_t n_first = ~size_t(0), const std::function<void()>& between = nullptr);
void phase_deposit(const std::vector<b2p_tile*>& tiles);
void phase_sort(const std::vector<b2p_tile*>& tiles, bool leave_running = false);
void join_pending_sort();                   // host.cu: makes the library stream wait for a sort left on the worker streams
void phase_pack_outgoing(const std::vector<b2p_tile*>& tiles);
void phase_apply_edge_bcs(const std::vector<b2p_tile*>& tiles, int mode);
void phase_reflect_particles(const std::vector<b2p_tile*>& tiles);
void grid_local_communication(b2p_grid* g, int mode, int part = 0);   // part: fields.cuh launch_halo_fill
void comm_exchange_fields_on_comm_stream(b2p_grid* g, int mode);      // comm.cu: exchange on the plan's own stream, ordered after the library stream
void comm_unpack_halo(b2p_grid* g, int which);                          // the remote-fed halo cells of E / B / J from the staged slabs
void comm_wait_exchange(b2p_grid* g);                                   // library stream waits for that exchange
void flush_deferred();                      // executes the pending batch of per-tile calls (host.cu)
void set_last_error(const std::string& s);
void Scratch_table_upload(const void* src, size_t bytes);   // host -> the shared device table scratch
const void* Scratch_table_ptr();
// multi-GPU plan queries (comm.cu)
bool comm_remote_entry(b2p_grid* g, int slot, int dir_idx, int* entry);
const void* comm_remote_table(b2p_grid* g, int kind);        // kind 0: halo-fill slabs, 1: J-exchange slabs
bool comm_particle_span(b2p_grid* g, int entry, int species, const b2p_particle_state** ptr, unsigned* count);
void comm_particles_consumed(b2p_grid* g);
}  // namespace b2p
