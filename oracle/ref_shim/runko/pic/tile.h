// TEST-SIDE STAND-IN for the reference's runko/pic/tile.h, found first on the include path when
// oracle/Makefile.ref compiles /root/reference/src/runko/pic/reflector_wall.c++ where it lies.
// The real header pulls in corgi, MPI and pybind11, none of which exist in this image; this
// one declares only the data members and the three methods that reflector_wall.c++ DEFINES
// (Tile<D>::register_reflector_wall / reflect_particles / advance_reflector_walls), with the
// reference's own member names and types (pic/tile.h:55-75, emf/tile.h:40-60, corgi/tile.h
// mins/maxs).  It contains no arithmetic: every line of reflector logic that runs in
// oracle/_ref/libref_kernels.so is compiled from the reference's source file.
#pragma once

#include "runko/emf/yee_lattice.h"
#include "runko/pic/particle.h"
#include "runko/pic/reflector_wall.h"

#include <array>
#include <cstddef>
#include <map>
#include <optional>
#include <vector>

namespace pic {

template<std::size_t D>
class Tile {
public:
  using value_type = float;                      // pic/tile.h: value_type = ParticleContainer::value_type

  std::array<double, 3> mins {}, maxs {};        // corgi::Tile<D>::mins / maxs
  double cfl_ {};                                // emf::Tile<D>::cfl_
  emf::YeeLattice yee_lattice_;                  // emf::Tile<D>::yee_lattice_
  std::map<std::size_t, ParticleContainer> particle_buffs_;

  std::vector<pic::reflector_wall> reflector_walls_ {};
  std::optional<runko::VecGrid<emf::YeeLattice::value_type>> reflector_correction_J_ {};
  bool reflector_correction_pending_ { false };

  explicit Tile(const emf::YeeLatticeCtorArgs a) : yee_lattice_(a) {}

  void register_reflector_wall(pic::reflector_wall wall);
  void reflect_particles();
  void advance_reflector_walls();
};

}  // namespace pic
