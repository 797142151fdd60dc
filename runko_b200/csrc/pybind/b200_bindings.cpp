// b200_bindings.cpp — the pybind11 layer over the C-ABI of libb200pic.so (include/b200pic.h).
//
// This is the compiled half of the drop-in for the reference's `runko_cpp_bindings` / `pycorgi`
// modules (src/runko/bindings/runko_cpp_bindings.c++:14-40, pypic.c++:92-139, pyemf.c++:214-281,
// pytools.c++:23-37, external/corgi/pycorgi/pycorgi.c++:79-126,317-374): one handle class per
// opaque C handle, one method per entry point, C status codes turned into the exception types the
// reference throws (std::logic_error -> B2P_ERR_LOGIC, std::runtime_error otherwise).  The Python
// packages under runko_b200/dropin/ give these handles the reference's module / class names and
// add the host-side logic the reference keeps in C++ above its kernels (Yee-staggered sample
// points of the setters, injection cell order, config parsing rules).
//
// Links only libb200pic.so: no tyvi, no thrust, no MPI, no torch.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/b200pic.h"

namespace py = pybind11;

namespace {

void ck(const int rc) {
  if (rc == B2P_OK) return;
  const std::string msg = b2p_last_error();
  if (rc == B2P_ERR_LOGIC) throw std::logic_error(msg);
  throw std::runtime_error(msg);
}

template <class T>
T pod_from_bytes(const py::bytes& b, const char* what) {
  const std::string s = b;
  if (s.size() != sizeof(T)) throw std::runtime_error(std::string(what) + ": struct size mismatch");
  T v;
  std::memcpy(&v, s.data(), sizeof(T));
  return v;
}

using farr = py::array_t<float, py::array::c_style | py::array::forcecast>;
using darr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using u64arr = py::array_t<uint64_t, py::array::c_style | py::array::forcecast>;

struct TileHandle {
  b2p_tile* h = nullptr;
  b2p_config cfg{};
  TileHandle(const std::array<int32_t, 3> idx, const py::bytes& config) {
    cfg = pod_from_bytes<b2p_config>(config, "b2p_config");
    ck(b2p_tile_create(&cfg, idx.data(), &h));
  }
  TileHandle(const TileHandle&) = delete;
  TileHandle& operator=(const TileHandle&) = delete;
  ~TileHandle() { if (h) b2p_tile_destroy(h); }

  size_t lattice_elems(const bool with_halo) const {
    size_t n = 3;
    for (int d = 0; d < 3; ++d) n *= size_t(cfg.n_cells[d] + (with_halo ? 2 * B2P_HALO : 0));
    return n;
  }
  std::vector<py::ssize_t> lattice_shape(const bool with_halo) const {
    std::vector<py::ssize_t> s{ 3 };
    for (int d = 0; d < 3; ++d) s.push_back(cfg.n_cells[d] + (with_halo ? 2 * B2P_HALO : 0));
    return s;
  }
  const float* field_ptr(const py::object& o, farr& keep, const bool with_halo) const {
    if (o.is_none()) return nullptr;
    keep = farr::ensure(o);
    if (!keep || size_t(keep.size()) != lattice_elems(with_halo)) throw std::runtime_error("Batch field setter returned array with incorrect shape!");
    return keep.data();
  }
};

struct GridHandle {
  b2p_grid* h = nullptr;
  b2p_config cfg{};
  std::vector<py::object> keep;   // tiles added to the grid stay alive with it (pycorgi.c++:333 keep_alive)
  explicit GridHandle(const py::bytes& config) {
    cfg = pod_from_bytes<b2p_config>(config, "b2p_config");
    ck(b2p_grid_create(&cfg, &h));
  }
  GridHandle(const GridHandle&) = delete;
  GridHandle& operator=(const GridHandle&) = delete;
  ~GridHandle() { if (h) b2p_grid_destroy(h); }
};

}  // namespace

PYBIND11_MODULE(_b200pic, m) {
  m.doc() = "pybind11 bindings of libb200pic.so (B200 PIC hot path behind runko's tile API)";
  m.attr("HALO") = B2P_HALO;
  m.attr("MAX_SPECIES") = B2P_MAX_SPECIES;
  m.attr("sizeof_config") = sizeof(b2p_config);
  m.attr("sizeof_edge_bc") = sizeof(b2p_edge_bc);
  m.attr("sizeof_reflector_wall") = sizeof(b2p_reflector_wall);

  m.def("version", [] { return std::string(b2p_version()); });
  m.def("init", [](const int device) { ck(b2p_init(device)); });
  m.def("sync", [] { ck(b2p_sync()); });
  m.def("set_option", [](const std::string& name, const int value) { ck(b2p_set_option(name.c_str(), value)); });
  m.def("gpu_mem_kB", [] { return b2p_gpu_mem_kB(); });                                   // tools/gpu_memory.h:16-28
  m.def("launch_count", [] { return b2p_launch_count(); });
  m.def("nccl_unique_id", [] { char id[128]; ck(b2p_nccl_unique_id(id)); return py::bytes(id, 128); });

  py::class_<TileHandle>(m, "TileHandle")
    .def(py::init<std::array<int32_t, 3>, py::bytes>())
    .def("bounds", [](TileHandle& t) {
      double mn[3], mx[3];
      ck(b2p_tile_bounds(t.h, mn, mx));
      return py::make_tuple(std::vector<double>(mn, mn + 3), std::vector<double>(mx, mx + 3));
    })
    // ---- fields (fp32, component-major, k fastest) ----
    .def("set_fields", [](TileHandle& t, const py::object& E, const py::object& B, const py::object& J, const bool with_halo) {
      farr ke, kb, kj;
      const float* e = t.field_ptr(E, ke, with_halo);
      const float* b = t.field_ptr(B, kb, with_halo);
      const float* j = t.field_ptr(J, kj, with_halo);
      ck(b2p_tile_set_fields(t.h, e, b, j, with_halo ? 1 : 0));
    })
    .def("get_fields", [](TileHandle& t, const bool with_halo) {
      farr E(t.lattice_shape(with_halo)), B(t.lattice_shape(with_halo)), J(t.lattice_shape(with_halo));
      ck(b2p_tile_get_fields(t.h, E.mutable_data(), B.mutable_data(), J.mutable_data(), with_halo ? 1 : 0));
      return py::make_tuple(E, B, J);
    })
    .def("push_half_b", [](TileHandle& t) { ck(b2p_tile_push_half_b(t.h)); })              // pyemf.c++:244
    .def("push_e", [](TileHandle& t) { ck(b2p_tile_push_e(t.h)); })                        // pyemf.c++:245
    .def("add_current", [](TileHandle& t) { ck(b2p_tile_add_current(t.h)); })              // pyemf.c++:246
    .def("filter_current", [](TileHandle& t) { ck(b2p_tile_filter_current(t.h)); })        // pyemf.c++:247
    .def("clear_current", [](TileHandle& t) { ck(b2p_tile_clear_current(t.h)); })
    .def("field_energy", [](TileHandle& t) {
      double b = 0, e = 0;
      ck(b2p_tile_field_energy(t.h, &b, &e));
      return py::make_tuple(b, e);
    })
    // ---- particles ----
    .def("inject", [](TileHandle& t, const int sp, const darr& x, const darr& y, const darr& z, const darr& ux, const darr& uy, const darr& uz) {
      const py::ssize_t n = x.size();
      if (y.size() != n || z.size() != n || ux.size() != n || uy.size() != n || uz.size() != n)
        throw std::runtime_error("pic::Tile::batch_inject_in_x_stripe: batches must have same length.");
      ck(b2p_tile_inject(t.h, sp, uint64_t(n), x.data(), y.data(), z.data(), ux.data(), uy.data(), uz.data()));
    })
    .def("set_particles", [](TileHandle& t, const int sp, const farr& x, const farr& y, const farr& z, const farr& ux, const farr& uy,
                             const farr& uz, const u64arr& id) {
      ck(b2p_tile_set_particles(t.h, sp, uint64_t(id.size()), x.data(), y.data(), z.data(), ux.data(), uy.data(), uz.data(), id.data()));
    })
    .def("container_size", [](TileHandle& t, const int sp) { uint64_t n = 0; ck(b2p_tile_container_size(t.h, sp, &n)); return n; })
    .def("get_particles", [](TileHandle& t, const int sp, const bool alive_only) {
      uint64_t n = 0, mcount = 0;
      ck(b2p_tile_container_size(t.h, sp, &n));
      const py::ssize_t N = py::ssize_t(n);
      farr a[6] = { farr(N), farr(N), farr(N), farr(N), farr(N), farr(N) };
      u64arr id(N);
      ck(b2p_tile_get_particles(t.h, sp, alive_only ? 1 : 0, a[0].mutable_data(), a[1].mutable_data(), a[2].mutable_data(),
                                a[3].mutable_data(), a[4].mutable_data(), a[5].mutable_data(), id.mutable_data(), &mcount));
      py::tuple out(7);
      const py::ssize_t M = py::ssize_t(mcount);
      for (int q = 0; q < 6; ++q) { a[q].resize({ M }); out[q] = a[q]; }
      id.resize({ M });
      out[6] = id;
      return out;
    }, py::arg("sp"), py::arg("alive_only") = true)
    .def("push_particles", [](TileHandle& t) { ck(b2p_tile_push_particles(t.h)); })                    // pypic.c++:126
    .def("pack_outgoing_particles", [](TileHandle& t) { ck(b2p_tile_pack_outgoing_particles(t.h)); })  // pypic.c++:127
    .def("deposit_current", [](TileHandle& t) { ck(b2p_tile_deposit_current(t.h)); })                  // pypic.c++:128
    .def("sort_particles", [](TileHandle& t) { ck(b2p_tile_sort_particles(t.h)); })                    // pypic.c++:129
    .def("kinetic_energy", [](TileHandle& t, const int sp) {
      double e = 0; uint64_t n = 0;
      ck(b2p_tile_kinetic_energy(t.h, sp, &e, &n));
      return py::make_tuple(e, n);
    })
    // ---- pic-shock pieces, antenna ----
    .def("register_edge_bc", [](TileHandle& t, const py::bytes& bc) {
      const b2p_edge_bc v = pod_from_bytes<b2p_edge_bc>(bc, "b2p_edge_bc");
      ck(b2p_tile_register_edge_bc(t.h, &v));
    })
    .def("apply_edge_bcs", [](TileHandle& t, const int mode) { ck(b2p_tile_apply_edge_bcs(t.h, mode)); })
    .def("apply_edge_bc", [](TileHandle& t, const py::bytes& bc, const int mode) {
      const b2p_edge_bc v = pod_from_bytes<b2p_edge_bc>(bc, "b2p_edge_bc");
      ck(b2p_tile_apply_edge_bc(t.h, &v, mode));
    })
    .def("register_reflector_wall", [](TileHandle& t, const py::bytes& w) {
      const b2p_reflector_wall v = pod_from_bytes<b2p_reflector_wall>(w, "b2p_reflector_wall");
      ck(b2p_tile_register_reflector_wall(t.h, &v));
    })
    .def("reflect_particles", [](TileHandle& t) { ck(b2p_tile_reflect_particles(t.h)); })
    .def("advance_reflector_walls", [](TileHandle& t) { ck(b2p_tile_advance_reflector_walls(t.h)); })
    .def("reflector_walls", [](TileHandle& t) {
      b2p_reflector_wall w[16];
      uint64_t n = 0;
      ck(b2p_tile_reflector_walls(t.h, w, 16, &n));
      std::vector<std::array<float, 3>> out;
      for (uint64_t q = 0; q < n && q < 16; ++q) out.push_back({ w[q].walloc, w[q].betawall, w[q].gammawall });
      return out;
    })
    .def("register_antenna", [](TileHandle& t, const std::array<double, 3> A, const std::array<double, 3> wave, const int wave_kind,
                                const py::object& lap_coeffs /* None or (n, 2) float64: re, im */) {
      b2p_antenna_mode mode{};
      for (int d = 0; d < 3; ++d) { mode.A[d] = A[d]; mode.wave[d] = wave[d]; }
      mode.wave_kind = wave_kind;
      darr keep;
      const double one_pair[2] = { 0.0, 0.0 };
      if (!lap_coeffs.is_none()) {
        keep = darr::ensure(lap_coeffs);
        mode.n_lap_coeffs = uint64_t(keep.size() / 2);
        mode.lap_coeffs = keep.size() ? keep.data() : one_pair;   // non-NULL even when empty: NULL means "no lap_coeffs"
      }
      ck(b2p_tile_register_antenna(t.h, &mode));
    })
    .def("deposit_antenna_current", [](TileHandle& t) { ck(b2p_tile_deposit_antenna_current(t.h)); });

  py::class_<GridHandle>(m, "GridHandle")
    .def(py::init<py::bytes>())
    .def("add_tile", [](GridHandle& g, py::object tile_handle) {                                       // pycorgi.c++:333
      TileHandle& t = tile_handle.cast<TileHandle&>();
      ck(b2p_grid_add_tile(g.h, t.h));
      g.keep.push_back(std::move(tile_handle));
    })
    .def("local_communication", [](GridHandle& g, const int mode) { ck(b2p_grid_local_communication(g.h, mode)); })    // corgi.h:1697-1718
    .def("external_communication", [](GridHandle& g, const int mode) { ck(b2p_grid_external_communication(g.h, mode)); })   // corgi.h:1560-1692
    .def("comm_init", [](GridHandle& g, const int rank, const int nranks, const py::bytes& id, const std::vector<int32_t>& owner) {
      const std::string s = id;
      if (s.size() != 128) throw std::runtime_error("NCCL unique id must be 128 bytes");
      ck(b2p_grid_comm_init(g.h, rank, nranks, s.data(), owner.data()));
    })
    .def("push_half_b", [](GridHandle& g) { ck(b2p_grid_push_half_b(g.h)); })
    .def("push_e", [](GridHandle& g) { ck(b2p_grid_push_e(g.h)); })
    .def("add_current", [](GridHandle& g) { ck(b2p_grid_add_current(g.h)); })
    .def("filter_current", [](GridHandle& g) { ck(b2p_grid_filter_current(g.h)); })
    .def("push_particles", [](GridHandle& g) { ck(b2p_grid_push_particles(g.h)); })
    .def("pack_outgoing_particles", [](GridHandle& g) { ck(b2p_grid_pack_outgoing_particles(g.h)); })
    .def("sort_particles", [](GridHandle& g) { ck(b2p_grid_sort_particles(g.h)); })
    .def("deposit_current", [](GridHandle& g) { ck(b2p_grid_deposit_current(g.h)); })
    .def("apply_edge_bcs", [](GridHandle& g, const int mode) { ck(b2p_grid_apply_edge_bcs(g.h, mode)); })
    .def("reflect_particles", [](GridHandle& g) { ck(b2p_grid_reflect_particles(g.h)); })
    .def("advance_reflector_walls", [](GridHandle& g) { ck(b2p_grid_advance_reflector_walls(g.h)); })
    .def("step_pic", [](GridHandle& g, const int64_t lap) { ck(b2p_grid_step_pic(g.h, lap)); })
    .def("step_emf", [](GridHandle& g) { ck(b2p_grid_step_emf(g.h)); })
    .def("energies", [](GridHandle& g) {                                                               // io/pic_average_kinetic_energy.h, io/emf_average_field_energy_density.h
      const int ns = g.cfg.n_species > 0 ? g.cfg.n_species : 1;
      double b = 0, e = 0;
      std::vector<double> kin(size_t(ns), 0.0);
      std::vector<uint64_t> sizes(size_t(ns), 0);
      ck(b2p_grid_energies(g.h, &b, &e, kin.data(), sizes.data()));
      kin.resize(size_t(g.cfg.n_species > 0 ? g.cfg.n_species : 0));
      sizes.resize(kin.size());
      return py::make_tuple(b, e, kin, sizes);
    })
    .def("alive_counts", [](GridHandle& g) {
      std::vector<uint64_t> c(size_t(g.cfg.n_species > 0 ? g.cfg.n_species : 1), 0);
      ck(b2p_grid_alive_counts(g.h, c.data()));
      c.resize(size_t(g.cfg.n_species > 0 ? g.cfg.n_species : 0));
      return c;
    })
    .def("write_fields_snapshot", [](GridHandle& g, const std::string& prefix, const int lap, const int stride, const int nspecies) {
      ck(b2p_grid_write_fields_snapshot(g.h, prefix.c_str(), lap, stride, nspecies));                  // io/snapshots/mpiio_fields.c++:221-400
    });
}
