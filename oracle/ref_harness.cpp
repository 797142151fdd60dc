// ref_harness.cpp — a C-ABI around the REFERENCE'S OWN kernel sources.
//
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_kernels.so; see oracle/Makefile.ref).
// This file contains no PIC arithmetic.  It instantiates the reference's
// emf::YeeLattice and pic::ParticleContainer — compiled from the sources where they lie
// under /root/reference/src/runko/{emf,pic}, CPU backend (TYVI_BACKEND_CPU, thrust CPP
// device system) — and restates only the few lines of glue that pic::Tile / emf::Tile put
// between the Python call and those classes (origo = float(mins) - 3, the sort score, the
// subregion dividers, dt = cfl/2 or cfl), each cited below.  It exists to pin the oracle
// restatement (oracle/pic_oracle.cpp) against the reference's real kernels bit for bit:
// tests/test_oracle_vs_reference_build.py.
//
// What is NOT the reference here: corgi (tile grid / MPI transport) and the pybind11
// bindings are not linked (no MPI in this image), so halo / migration transport is
// exercised through YeeLattice::set_*_in_subregion / add_to_J_from_subregion and
// ParticleContainer::{divide_to_subregions, append} directly, driven by this harness in
// corgi's documented Moore order (external/corgi/src/corgi/cellular_automata.h:48-62).
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

#include "runko/emf/edge_bc.h"
#include "runko/emf/stencil_coefficients.h"
#include "runko/emf/yee_lattice.h"
#include "runko/pic/particle.h"
#include "runko/pic/tile.h"   // oracle/ref_shim stand-in: declarations only (see that file)

#include "../include/b200pic.h"

namespace {

thread_local std::string g_err;

struct RefTile {
  b2p_config cfg;
  int idx[3];
  double mins[3], maxs[3];
  emf::YeeLattice lattice;
  std::vector<pic::ParticleContainer> sp;
  std::optional<runko::VecGrid<float>> generated_J;                       // pic/tile.h generated_J_cache_
  thrust::device_vector<runko::ParticleState<float>> out_buf;            // subregion_particle_buff_
  std::vector<std::size_t> out_ends;                                     // subregion_particle_ends_
  emf::StencilCoeffs stencil;
  // pic-shock pieces: the reference's Tile<3>::{register_reflector_wall, reflect_particles,
  // advance_reflector_walls} (compiled from pic/reflector_wall.c++) run on this object; the
  // containers are moved into it for the duration of reflect_particles
  std::unique_ptr<pic::Tile<3>> shock;
  std::vector<emf::edge_bc> edge_bcs;                                                 // emf/tile.h:51

  pic::Tile<3>& shock_tile() {
    if (!shock) {
      shock = std::make_unique<pic::Tile<3>>(
        emf::YeeLatticeCtorArgs{ std::size_t(cfg.n_cells[0]), std::size_t(cfg.n_cells[1]), std::size_t(cfg.n_cells[2]) });
      for (int d = 0; d < 3; ++d) { shock->mins[d] = mins[d]; shock->maxs[d] = maxs[d]; }
      shock->cfl_ = cfg.cfl;
    }
    return *shock;
  }

  RefTile(const b2p_config& c, const int i[3])
      : cfg(c), lattice(emf::YeeLatticeCtorArgs{ std::size_t(c.n_cells[0]), std::size_t(c.n_cells[1]), std::size_t(c.n_cells[2]) }) {
    for (int d = 0; d < 3; ++d) {
      idx[d]  = i[d];
      mins[d] = double(std::size_t(i[d]) * std::size_t(c.n_cells[d]));               // emf/tile.c++:162-170
      maxs[d] = double((std::size_t(i[d]) + 1) * std::size_t(c.n_cells[d]));
    }
    for (int s = 0; s < c.n_species; ++s)
      sp.emplace_back(pic::ParticleContainerArgs{ .N = 0, .charge = c.q[s], .mass = c.m[s] });   // pic/tile.c++:52-71
    for (int a = 0; a < 3; ++a) {                                                     // emf/tile.c++:99-142
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 5; ++q) stencil.axis[a].M[r][q] = c.stencil[a][r][q];
      stencil.axis[a].M[0][0] = 0.0f;
      stencil.axis[a].M[0][0] = stencil.axis[a].alpha();
    }
  }
  std::array<float, 3> origo() const {                                                // pic/tile.c++:329-332
    return { static_cast<float>(mins[0]) - emf::halo_size, static_cast<float>(mins[1]) - emf::halo_size,
             static_cast<float>(mins[2]) - emf::halo_size };
  }
};

template <class MDS>
void fill_from(const MDS& m, const float* src, std::size_t H0, std::size_t H1, std::size_t H2) {
  const std::size_t Ch = H0 * H1 * H2;
  for (std::size_t i = 0; i < H0; ++i)
    for (std::size_t j = 0; j < H1; ++j)
      for (std::size_t k = 0; k < H2; ++k)
        for (std::size_t c = 0; c < 3; ++c) m[i, j, k][c] = src[c * Ch + (i * H1 + j) * H2 + k];
}
template <class MDS>
void read_to(const MDS& m, float* dst, std::size_t H0, std::size_t H1, std::size_t H2) {
  const std::size_t Ch = H0 * H1 * H2;
  for (std::size_t i = 0; i < H0; ++i)
    for (std::size_t j = 0; j < H1; ++j)
      for (std::size_t k = 0; k < H2; ++k)
        for (std::size_t c = 0; c < 3; ++c) dst[c * Ch + (i * H1 + j) * H2 + k] = m[i, j, k][c];
}

}  // namespace

#define REF_TRY try {
#define REF_CATCH                                              \
  }                                                            \
  catch (const std::exception& e) { g_err = e.what(); return 1; } \
  return 0;

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

void* ref_tile_create(const b2p_config* cfg, const int32_t idx[3]) {
  try {
    const int i[3] = { idx[0], idx[1], idx[2] };
    return new RefTile(*cfg, i);
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_tile_destroy(void* t) { delete static_cast<RefTile*>(t); }

// whole haloed lattices, fp32 component-major buf[c*Ch + (i*Hy + j)*Hz + k]
int ref_tile_set_fields(void* tp, const float* E, const float* B, const float* J) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  const auto e = t.lattice.extents_with_halo();
  if (E) fill_from(t.lattice.mds_E(), E, e[0], e[1], e[2]);
  if (B) fill_from(t.lattice.mds_B(), B, e[0], e[1], e[2]);
  if (J) fill_from(t.lattice.mds_J(), J, e[0], e[1], e[2]);
  REF_CATCH
}
int ref_tile_get_fields(void* tp, float* E, float* B, float* J) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  const auto e = t.lattice.extents_with_halo();
  if (E) read_to(t.lattice.mds_E(), E, e[0], e[1], e[2]);
  if (B) read_to(t.lattice.mds_B(), B, e[0], e[1], e[2]);
  if (J) read_to(t.lattice.mds_J(), J, e[0], e[1], e[2]);
  REF_CATCH
}

// raw container upload (dead slots included): a fresh container filled through the
// reference's own add_particles -> append path (pic/particle.h:258-287)
int ref_tile_set_particles(void* tp, int sp, uint64_t n, const float* x, const float* y, const float* z, const float* ux,
                           const float* uy, const float* uz, const uint64_t* id) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  t.sp.at(sp) = pic::ParticleContainer(pic::ParticleContainerArgs{ .N = 0, .charge = t.cfg.q[sp], .mass = t.cfg.m[sp] });
  std::vector<runko::ParticleState<double>> v(n);
  for (uint64_t i = 0; i < n; ++i)
    v[i] = runko::ParticleState<double>{ .pos{ x[i], y[i], z[i] }, .vel{ ux[i], uy[i], uz[i] }, .id = id[i] };
  if (n) {
    // append() looks for the last ALIVE slot from the back; keep trailing dead slots by
    // appending them as a second span after the alive prefix would drop them, so write the
    // whole block through one span into the empty container (all slots are kept).
    t.sp.at(sp).add_particles(v);
  }
  REF_CATCH
}
int ref_tile_container_size(void* tp, int sp, uint64_t* n) {
  REF_TRY
  *n = static_cast<RefTile*>(tp)->sp.at(sp).size();
  REF_CATCH
}
// raw container download (dead slots included)
int ref_tile_get_particles(void* tp, int sp, float* x, float* y, float* z, float* ux, float* uy, float* uz, uint64_t* id) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  const auto& c  = t.sp.at(sp);
  const auto pos = c.pos_mds();
  const auto vel = c.vel_mds();
  const auto ids = c.ids_mds();
  const std::size_t n = c.size();
  for (std::size_t i = 0; i < n; ++i) {
    const auto k = static_cast<runko::index_t>(i);
    x[i] = pos[k][0]; y[i] = pos[k][1]; z[i] = pos[k][2];
    ux[i] = vel[k][0]; uy[i] = vel[k][1]; uz[i] = vel[k][2];
    id[i] = ids[k][];
  }
  REF_CATCH
}

int ref_tile_op(void* tp, const char* name) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  const std::string op(name);
  using vt = emf::YeeLattice::value_type;
  if (op == "push_half_b") {                                                          // emf/tile.c++:359-375
    if (t.cfg.field_propagator == B2P_PROPAGATOR_STENCIL) t.lattice.push_b_stencil(static_cast<vt>(t.cfg.cfl / 2), t.stencil);
    else t.lattice.push_b_fdtd2(static_cast<vt>(t.cfg.cfl / 2));
  } else if (op == "push_e") {                                                        // emf/tile.c++:379-394
    t.lattice.push_e_fdtd2(static_cast<vt>(t.cfg.cfl));
  } else if (op == "add_current") {
    t.lattice.add_current();
  } else if (op == "clear_current") {
    t.lattice.clear_current();
  } else if (op == "filter_current") {                                                // emf/tile.c++:405-426
    if (t.cfg.current_filter == B2P_FILTER_BINOMIAL2) t.lattice.filter_current_binomial2();
    else if (t.cfg.current_filter == B2P_FILTER_BINOMIAL2_UNROLLED) t.lattice.filter_current_binomial2_unrolled();
    else throw std::logic_error("Trying to filter current without specifying `current_filter`!");
  } else if (op == "push_particles") {                                                // pic/tile.c++:326-365
    auto push_impl = [&](const auto& interp) {
      for (auto& c : t.sp) {
        switch (t.cfg.particle_pusher) {
          case B2P_PUSHER_BORIS: c.push_particles_boris(t.cfg.cfl, interp); break;
          case B2P_PUSHER_HIGUERA_CARY: c.push_particles_higuera_cary(t.cfg.cfl, interp); break;
          case B2P_PUSHER_FARADAY: c.push_particles_faraday(t.cfg.cfl, interp); break;
          default: throw std::logic_error("unknown pusher");
        }
      }
    };
    if (t.cfg.field_interpolator == B2P_INTERP_LINEAR_1ST) push_impl(t.lattice.interpolate_EB_linear_1st(t.origo()));
    else push_impl(t.lattice.interpolate_EB_linear_1st_unrolled(t.origo()));
  } else if (op == "deposit_current") {                                               // pic/tile.c++:369-415
    t.lattice.clear_current();
    if (t.cfg.current_depositer == B2P_DEPOSIT_ZIGZAG_1ST) {
      for (const auto& c : t.sp) t.lattice.deposit_current(c.current_zigzag_1st(t.origo(), t.cfg.cfl));
    } else {
      if (!t.generated_J) t.generated_J = runko::VecGrid<float>(t.lattice.extents_with_halo());
      auto& gJ          = t.generated_J.value();
      const auto genJmds = gJ.mds();
      tyvi::mdgrid_work{}.for_each_index(genJmds, [=](const auto idx, const auto tidx) { genJmds[idx][tidx] = 0; }).wait();
      for (const auto& c : t.sp) c.current_zigzag_1st(gJ, t.origo(), t.cfg.cfl);
      t.lattice.deposit_current(gJ);
    }
    if (t.shock && t.shock->reflector_correction_pending_) {                          // pic/tile.c++:411-414
      t.lattice.deposit_current(t.shock->reflector_correction_J_.value());
      t.shock->reflector_correction_pending_ = false;
    }
  } else if (op == "reflect_particles") {                                             // pic/reflector_wall.c++:241-284
    pic::Tile<3>& st = t.shock_tile();
    for (std::size_t i = 0; i < t.sp.size(); ++i) st.particle_buffs_.insert_or_assign(i, std::move(t.sp[i]));
    try { st.reflect_particles(); } catch (...) {
      for (std::size_t i = 0; i < t.sp.size(); ++i) t.sp[i] = std::move(st.particle_buffs_.at(i));
      throw;
    }
    for (std::size_t i = 0; i < t.sp.size(); ++i) t.sp[i] = std::move(st.particle_buffs_.at(i));
    st.particle_buffs_.clear();
  } else if (op == "advance_reflector_walls") {                                       // pic/reflector_wall.c++:286-297
    t.shock_tile().advance_reflector_walls();
  } else if (op == "sort_particles") {                                                // pic/tile.c++:419-438
    const auto m = t.lattice.grid_mapping_with_halo();
    using M      = decltype(m);
    using F      = pic::ParticleContainer::value_type;
    const auto origo_pos = t.origo();
    using Vec3F  = toolbox::Vec3<F>;
    auto score = [=](const F x, const F y, const F z) {
      const auto dx  = Vec3F(x, y, z) - Vec3F(origo_pos);
      const auto idx = dx.template as<typename M::index_type>();
      return m(idx[0], idx[1], idx[2]);
    };
    for (auto& c : t.sp) c.sort(score);
  } else if (op == "pack_outgoing_particles") {                                       // pic/tile_communication.c++:68-96
    using F = pic::ParticleContainer::value_type;
    const auto x_div = std::array{ static_cast<F>(t.mins[0]), static_cast<F>(t.maxs[0]) };
    const auto y_div = std::array{ static_cast<F>(t.mins[1]), static_cast<F>(t.maxs[1]) };
    const auto z_div = std::array{ static_cast<F>(t.mins[2]), static_cast<F>(t.maxs[2]) };
    t.out_ends.assign(27 * t.sp.size(), 0);
    t.out_buf.resize(0);
    for (std::size_t ptype = 0; ptype < t.sp.size(); ++ptype) {
      auto spans = t.sp[ptype].divide_to_subregions(t.out_buf, x_div, y_div, z_div);
      for (const auto& [dir, span] : spans) t.out_ends.at(27 * ptype + dir.neighbor_index()) = std::get<1>(span);
    }
  } else {
    throw std::runtime_error("ref_tile_op: unknown op " + op);
  }
  REF_CATCH
}

// emf::Tile::edge_bc_width (emf/tile.c++:808-827; emf/tile.c++ itself needs corgi and is not
// compiled) + YeeLattice::apply_edge_bc, the reference's own (emf/yee_lattice.c++:263-306)
static void apply_one_edge_bc(RefTile& t, const emf::edge_bc& bc, int mode) {
  const auto d        = bc.direction;
  const auto tile_min = static_cast<emf::edge_bc::value_type>(t.mins[d]);
  const auto tile_max = static_cast<emf::edge_bc::value_type>(t.maxs[d]);
  const auto Nd       = t.lattice.extents_wout_halo()[d];
  std::optional<std::size_t> w;
  if (bc.side == 0) {
    if (bc.position <= tile_min) w = std::nullopt;
    else if (bc.position >= tile_max) w = Nd;
    else w = static_cast<std::size_t>(bc.position - tile_min) + 1;
  } else {
    if (bc.position >= tile_max) w = std::nullopt;
    else if (bc.position <= tile_min) w = Nd;
    else w = Nd - static_cast<std::size_t>(bc.position - tile_min);
  }
  if (w) t.lattice.apply_edge_bc(bc, *w, mode);
}
static emf::edge_bc to_ref(const b2p_edge_bc& b) {
  return emf::edge_bc{ b.direction, b.side, b.position, b.E[0], b.E[1], b.E[2], b.B[0], b.B[1], b.B[2],
                       b.J[0], b.J[1], b.J[2], b.E_components, b.B_components, b.J_components };
}
int ref_tile_register_edge_bc(void* tp, const b2p_edge_bc* bc) {
  REF_TRY
  static_cast<RefTile*>(tp)->edge_bcs.push_back(to_ref(*bc));
  REF_CATCH
}
int ref_tile_apply_edge_bc(void* tp, const b2p_edge_bc* bc, int mode) {
  REF_TRY
  apply_one_edge_bc(*static_cast<RefTile*>(tp), to_ref(*bc), mode);
  REF_CATCH
}
int ref_tile_apply_edge_bcs(void* tp, int mode) {                                     // emf/tile.c++:842-847
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  for (const auto& bc : t.edge_bcs) apply_one_edge_bc(t, bc, mode);
  REF_CATCH
}
int ref_tile_register_reflector_wall(void* tp, const b2p_reflector_wall* w) {
  REF_TRY
  static_cast<RefTile*>(tp)->shock_tile().register_reflector_wall(
    pic::reflector_wall{ .walloc = w->walloc, .betawall = w->betawall, .gammawall = w->gammawall });
  REF_CATCH
}
int ref_tile_reflector_walls(void* tp, b2p_reflector_wall* out, uint64_t cap, uint64_t* n) {
  REF_TRY
  const auto& walls = static_cast<RefTile*>(tp)->shock_tile().reflector_walls_;
  *n = walls.size();
  for (std::size_t q = 0; q < walls.size() && q < cap; ++q) out[q] = b2p_reflector_wall{ walls[q].walloc, walls[q].betawall, walls[q].gammawall };
  REF_CATCH
}

int ref_tile_get_outgoing(void* tp, b2p_particle_state* buf, uint64_t cap, uint64_t* ends, uint64_t* n_out) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  static_assert(sizeof(runko::ParticleState<float>) == sizeof(b2p_particle_state));
  if (n_out) *n_out = t.out_buf.size();
  if (ends) for (std::size_t i = 0; i < t.out_ends.size(); ++i) ends[i] = t.out_ends[i];
  if (buf) {
    if (cap < t.out_buf.size()) throw std::runtime_error("outgoing buffer too small");
    if (!t.out_buf.empty()) std::memcpy(buf, thrust::raw_pointer_cast(t.out_buf.data()), t.out_buf.size() * sizeof(b2p_particle_state));
  }
  REF_CATCH
}

// ParticleContainer::append (pic/particle.h:454-572) of caller-provided AoS spans, with the
// global periodic wrap when wrap != 0 (pic/tile_communication.c++:100-117)
int ref_tile_append(void* tp, int sp, int nspans, const b2p_particle_state* const* spans, const uint64_t* counts, int wrap,
                    const float wmin[3], const float wmax[3]) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  using S = runko::ParticleState<float>;
  std::vector<thrust::device_vector<S>> store(nspans);
  std::vector<std::span<const S>> v;
  for (int i = 0; i < nspans; ++i) {
    store[i].resize(counts[i]);
    if (counts[i]) std::memcpy(thrust::raw_pointer_cast(store[i].data()), spans[i], counts[i] * sizeof(S));
    v.emplace_back(thrust::raw_pointer_cast(store[i].data()), counts[i]);
  }
  if (wrap)
    t.sp.at(sp).append(v, std::array<float, 3>{ wmin[0], wmin[1], wmin[2] }, std::array<float, 3>{ wmax[0], wmax[1], wmax[2] });
  else
    t.sp.at(sp).append(v);
  REF_CATCH
}

// One pairwise halo operation of emf::Tile::local_communication (emf/tile.c++:478-542):
// mode emf_E/B/J: dst.set_X_in_subregion(dir, src); mode emf_J_exchange:
// dst.add_to_J_from_subregion(dir, src).  `dir` is the direction from dst to src.
int ref_tile_halo(void* dstp, void* srcp, const int32_t dir[3], int mode) {
  REF_TRY
  RefTile& d = *static_cast<RefTile*>(dstp);
  RefTile& s = *static_cast<RefTile*>(srcp);
  const emf::YeeLattice::dir_type dd{ dir[0], dir[1], dir[2] };
  switch (mode) {
    case B2P_COMM_EMF_E: d.lattice.set_E_in_subregion(dd, s.lattice); break;
    case B2P_COMM_EMF_B: d.lattice.set_B_in_subregion(dd, s.lattice); break;
    case B2P_COMM_EMF_J: d.lattice.set_J_in_subregion(dd, s.lattice); break;
    case B2P_COMM_EMF_J_EXCHANGE: d.lattice.add_to_J_from_subregion(dd, s.lattice); break;
    default: throw std::logic_error("ref_tile_halo: unsupported mode");
  }
  REF_CATCH
}

int ref_tile_energies(void* tp, double* eB, double* eE, double* kinetic /*n_species*/) {
  REF_TRY
  RefTile& t = *static_cast<RefTile*>(tp);
  if (eB) *eB = t.lattice.total_energy_B();
  if (eE) *eE = t.lattice.total_energy_E();
  if (kinetic) for (std::size_t s = 0; s < t.sp.size(); ++s) kinetic[s] = t.sp[s].total_kinetic_energy();
  REF_CATCH
}

}  // extern "C"
