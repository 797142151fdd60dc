/* b200pic.h — C-ABI of libb200pic.so: a B200 (sm_100a) implementation of runko's
 * per-timestep PIC hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one
 * method of the reference's pybind11-exposed C++ tile API; the reference
 * interface each one stands in for is cited as `file:line` relative to
 * /root/reference.  INTEGRATION.md shows the pybind11-side stubs a runko
 * maintainer would add to `src/runko/bindings/py{emf,pic}.c++` to call these.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary;
 *  - every function returns 0 on success, non-zero on failure; the message is
 *    retrievable with b2p_last_error() (thread-local).  B2P_ERR_LOGIC mirrors
 *    the reference's std::logic_error, everything else std::runtime_error;
 *  - all device work is enqueued on one CUDA stream per process and is
 *    asynchronous; getters and b2p_sync() synchronise (SURVEY.md §8b
 *    "Threading");
 *  - field arrays crossing the boundary are fp32, component-major
 *    `buf[c*N + (i*Ny + j)*Nz + k]` (k fastest), the reference's own layout
 *    (external/tyvi/src/tyvi/mdgrid_buffer.h:61-65);
 *  - particle arrays are SoA fp32 + uint64 ids (src/runko/pic/particle.h:73-78).
 */
#ifndef B200PIC_H
#define B200PIC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2P_MAX_SPECIES 8
#define B2P_HALO 3 /* src/runko/emf/common.h:9 */

enum { B2P_OK = 0, B2P_ERR_RUNTIME = 1, B2P_ERR_LOGIC = 2, B2P_ERR_CUDA = 3 };

/* src/runko/communication_common.h:30-38 */
enum {
  B2P_COMM_EMF_J               = 0,
  B2P_COMM_EMF_E               = 1,
  B2P_COMM_EMF_B               = 2,
  B2P_COMM_PIC_PARTICLE        = 3,
  B2P_COMM_PIC_PARTICLE_EXTRA  = 4,
  B2P_COMM_NUMBER_OF_PARTICLES = 5,
  B2P_COMM_EMF_J_EXCHANGE      = 6
};

enum { B2P_PROPAGATOR_FDTD2 = 0, B2P_PROPAGATOR_STENCIL = 1 };          /* emf/tile.h:27 */
enum { B2P_FILTER_NONE = -1, B2P_FILTER_BINOMIAL2 = 0, B2P_FILTER_BINOMIAL2_UNROLLED = 1 }; /* emf/tile.h:28 */
enum { B2P_PUSHER_NONE = -1, B2P_PUSHER_BORIS = 0, B2P_PUSHER_HIGUERA_CARY = 1, B2P_PUSHER_FARADAY = 2 }; /* pic/tile.h:34 */
enum { B2P_INTERP_LINEAR_1ST = 0, B2P_INTERP_LINEAR_1ST_UNROLLED = 1 }; /* pic/tile.h:35 */
enum { B2P_DEPOSIT_ZIGZAG_1ST = 0, B2P_DEPOSIT_ZIGZAG_1ST_ATOMIC = 1 }; /* pic/tile.h:36 */

/* Flat POD form of the Python config object the reference parses with
 * toolbox::ConfigParser (src/runko/tools/config_parser.c++:14-95); the keys are
 * the ones emf::Tile / pic::Tile read (emf/tile.c++:86-142, pic/tile.c++:28-147). */
typedef struct b2p_config {
  int32_t  n_tiles[3];          /* "n_tiles" */
  int32_t  n_cells[3];          /* "n_cells_per_tile" (each >= 3) */
  double   cfl;                 /* "cfl" */
  int32_t  field_propagator;    /* B2P_PROPAGATOR_* */
  int32_t  current_filter;      /* B2P_FILTER_* */
  /* stencil[a] is emf::StencilAxisCoeffs::M for axis a (emf/stencil_coefficients.h:30-64);
   * entry [a][0][0] (alpha) is ignored on input and recomputed from the others. */
  float    stencil[3][3][5];
  int32_t  n_species;           /* number of contiguous q<i>/m<i> pairs; 0 = emf-only tile */
  double   q[B2P_MAX_SPECIES];  /* "q<i>" signed charge */
  double   m[B2P_MAX_SPECIES];  /* "m<i>" |m/q| */
  int32_t  particle_pusher;     /* B2P_PUSHER_* */
  int32_t  field_interpolator;  /* B2P_INTERP_* */
  int32_t  current_depositer;   /* B2P_DEPOSIT_* */
  uint64_t prealloc_per_species;/* "prealloc_per_species" */
} b2p_config;

/* runko::ParticleState<float> — the 32-byte AoS migration wire format
 * (src/runko/particles_common.h:26-34). */
typedef struct b2p_particle_state {
  float    pos[3];
  float    vel[3];
  uint64_t id;
} b2p_particle_state;

typedef struct b2p_tile b2p_tile;
typedef struct b2p_grid b2p_grid;

/* ---- process-level ------------------------------------------------------ */
const char* b2p_last_error(void);
const char* b2p_version(void);
/* Select the CUDA device for this process (one process per GPU). Fails loudly
 * when no CUDA device is usable — there is no CPU fallback. */
int b2p_init(int device);
int b2p_sync(void);
/* Kernel-variant knob (launch bounds, aggregation strategy; see runko_b200/csrc/common.cuh
 * Tuning).  Never changes results beyond the stated deposit tolerance.  Also read from the
 * environment: B2P_OPTS="name=value,...".  No reference equivalent. */
int b2p_set_option(const char* name, int value);
/* Page-lock / unlock a caller-owned host buffer (cudaHostRegister) so that the getters and
 * setters below move it at full PCIe speed.  Optional; no reference equivalent. */
int b2p_host_register(void* ptr, size_t bytes);
int b2p_host_unregister(void* ptr);
/* tools._get_gpu_mem_kB (src/runko/tools/gpu_memory.h:16-28) */
int64_t b2p_gpu_mem_kB(void);

/* ---- tile life cycle ---------------------------------------------------- */
/* emf::Tile / pic::Tile constructor (emf/tile.c++:82-181, pic/tile.c++:119-147). */
int  b2p_tile_create(const b2p_config* cfg, const int32_t idx[3], b2p_tile** out);
void b2p_tile_destroy(b2p_tile* t);
/* corgi::Tile mins/maxs (emf/tile.c++:162-170), as doubles. */
int  b2p_tile_bounds(const b2p_tile* t, double mins[3], double maxs[3]);

/* ---- fields (emf::Tile) ------------------------------------------------- */
/* YeeLattice::set_EBJ (emf/yee_lattice.h:519-553): upload interior (with_halo=0,
 * arrays of 3*Nx*Ny*Nz floats) or the full haloed lattice (with_halo=1,
 * 3*(Nx+6)(Ny+6)(Nz+6)). Any of E,B,J may be NULL (left untouched). */
int b2p_tile_set_fields(b2p_tile* t, const float* E, const float* B, const float* J, int with_halo);
/* YeeLattice::get_EBJ / get_EBJ_with_halo (emf/yee_lattice.c++:91-168). */
int b2p_tile_get_fields(b2p_tile* t, float* E, float* B, float* J, int with_halo);
int b2p_tile_push_half_b(b2p_tile* t);   /* emf::Tile::push_half_b  emf/tile.c++:359-375 */
int b2p_tile_push_e(b2p_tile* t);        /* emf::Tile::push_e       emf/tile.c++:379-394 */
int b2p_tile_add_current(b2p_tile* t);   /* emf::Tile::add_current  emf/yee_lattice.c++:171-179 */
int b2p_tile_filter_current(b2p_tile* t);/* emf::Tile::filter_current emf/tile.c++:405-426 */
int b2p_tile_clear_current(b2p_tile* t); /* YeeLattice::clear_current emf/yee_lattice.c++:317-321 */
/* YeeLattice::total_energy_{B,E} (emf/yee_lattice.c++:383-428). */
int b2p_tile_field_energy(b2p_tile* t, double* energy_B, double* energy_E);

/* ---- particles (pic::Tile) ---------------------------------------------- */
/* pic::Tile::inject / batch_inject_* after the Python generator has run
 * (pic/tile.c++:207-322): doubles are narrowed to fp32 and ids
 * (tile_tag<<40)|ordinal are assigned here (pic/tile.c++:470-481). */
int b2p_tile_inject(b2p_tile* t, int sp, uint64_t n,
                    const double* x, const double* y, const double* z,
                    const double* ux, const double* uy, const double* uz);
/* Raw container upload including dead slots (id == UINT64_MAX); replaces the
 * container. Test / restart hook; no reference equivalent. */
int b2p_tile_set_particles(b2p_tile* t, int sp, uint64_t n,
                           const float* x, const float* y, const float* z,
                           const float* ux, const float* uy, const float* uz,
                           const uint64_t* id);
/* ParticleContainer::size (dead or alive) — pic/particle.c++:63-67. */
int b2p_tile_container_size(b2p_tile* t, int sp, uint64_t* n);
/* get_positions/get_velocities/get_ids (pic/particle.c++:82-168): alive_only=1
 * returns alive particles in container order; alive_only=0 returns the raw
 * container. Output arrays must hold container_size entries; NULL = skip. */
int b2p_tile_get_particles(b2p_tile* t, int sp, int alive_only,
                           float* x, float* y, float* z,
                           float* ux, float* uy, float* uz,
                           uint64_t* id, uint64_t* n_out);
int b2p_tile_push_particles(b2p_tile* t);          /* pic/tile.c++:326-365 */
int b2p_tile_deposit_current(b2p_tile* t);         /* pic/tile.c++:369-415 */
int b2p_tile_sort_particles(b2p_tile* t);          /* pic/tile.c++:419-438 */
int b2p_tile_pack_outgoing_particles(b2p_tile* t); /* pic/tile_communication.c++:68-96 */
/* The sort score of every slot (pic/tile.c++:430-435, pic/particle.h:607-614). */
int b2p_tile_sort_keys(b2p_tile* t, int sp, uint32_t* keys);
/* Outgoing AoS buffer and the 27*n_species cumulative end offsets
 * (pic/tile.h:84-85). buf may be NULL to query only ends / n_out. */
int b2p_tile_get_outgoing(b2p_tile* t, b2p_particle_state* buf, uint64_t cap,
                          uint64_t* ends, uint64_t* n_out);
/* ParticleContainer::total_kinetic_energy (pic/particle.c++:352-377). */
int b2p_tile_kinetic_energy(b2p_tile* t, int sp, double* energy, uint64_t* container_size);

/* ---- pic-shock boundary pieces (BASELINE configs[3]) ---------- */
/* emf::edge_bc (src/runko/emf/edge_bc.h:25-40): sets field components in the part of the
 * tile left (side 0) / right (side 1) of a global coordinate along `direction`. */
typedef struct b2p_edge_bc {
  uint8_t direction;            /* 0=x, 1=y, 2=z */
  uint8_t side;                 /* 0 = left of position, 1 = right of position */
  float   position;             /* global coordinate */
  float   E[3], B[3], J[3];     /* values written */
  uint8_t E_components, B_components, J_components;   /* bit 0=x, 1=y, 2=z */
} b2p_edge_bc;
/* pic::reflector_wall (src/runko/pic/reflector_wall.h:14-25): conducting piston in the yz-plane. */
typedef struct b2p_reflector_wall {
  float walloc;                 /* wall x in global coordinates */
  float betawall;               /* wall velocity / c */
  float gammawall;              /* wall Lorentz factor */
} b2p_reflector_wall;

/* emf::Tile::register_edge_bc / apply_edge_bcs / apply_edge_bc (emf/tile.c++:808-847,
 * emf/yee_lattice.c++:263-306).  `mode` is B2P_COMM_EMF_{E,B,J}; anything else fails like the
 * reference's std::runtime_error. */
int b2p_tile_register_edge_bc(b2p_tile* t, const b2p_edge_bc* bc);
int b2p_tile_apply_edge_bcs(b2p_tile* t, int mode);
int b2p_tile_apply_edge_bc(b2p_tile* t, const b2p_edge_bc* bc, int mode);
/* pic::Tile::register_reflector_wall / reflect_particles / advance_reflector_walls
 * (pic/reflector_wall.c++:226-297) and ParticleContainer::reflect_at_wall (:126-222): particles
 * behind a registered wall are reflected (or parked = marked dead), the +q / -q correction
 * currents go to a per-tile correction lattice that deposit_current adds to J
 * (pic/tile.c++:411-414). */
int b2p_tile_register_reflector_wall(b2p_tile* t, const b2p_reflector_wall* wall);
int b2p_tile_reflect_particles(b2p_tile* t);
int b2p_tile_advance_reflector_walls(b2p_tile* t);
/* Current state of the registered walls (no reference getter; used by the tests and the injector). */
int b2p_tile_reflector_walls(b2p_tile* t, b2p_reflector_wall* out, uint64_t cap, uint64_t* n);

/* ---- driven-turbulence antenna ---------------------------------- */
/* emf::antenna_mode (src/runko/emf/antenna.h:31-45): one vector-potential Fourier mode,
 * J += Re(cfl * curl(curl(A * phi[lap] * exp(i k.x)))).  `wave_kind` 0: `wave` is the wave vector k;
 * 1: `wave` is the mode number n (k = 2 pi n / L of the global grid).  `lap_coeffs` = n_lap_coeffs
 * complex numbers (re, im pairs) consumed one per deposit in the given order, or NULL for phi = 1. */
typedef struct b2p_antenna_mode {
  double        A[3];
  double        wave[3];
  int32_t       wave_kind;
  uint64_t      n_lap_coeffs;
  const double* lap_coeffs;
} b2p_antenna_mode;
/* emf::Tile::register_antenna / deposit_antenna_current (emf/tile.c++:566-791).  Depositing with an
 * exhausted lap_coeffs list fails with B2P_ERR_LOGIC like the reference's std::logic_error. */
int b2p_tile_register_antenna(b2p_tile* t, const b2p_antenna_mode* mode);
int b2p_tile_deposit_antenna_current(b2p_tile* t);

/* ---- grid (corgi::Grid surface used by runko/simulation.py) -------------- */
int  b2p_grid_create(const b2p_config* cfg, b2p_grid** out);
void b2p_grid_destroy(b2p_grid* g);
/* corgi::Grid::add_tile (external/corgi/src/corgi/corgi.h:434-467): the grid
 * borrows the tile; the caller keeps ownership. */
int  b2p_grid_add_tile(b2p_grid* g, b2p_tile* t);
/* corgi::Grid::local_communication (corgi.h:1697-1718) for all local tiles,
 * mode = B2P_COMM_*. */
int  b2p_grid_local_communication(b2p_grid* g, int mode);
/* The `for tile in local_tiles: tile.<method>()` loops of
 * runko/simulation.py:235-253, batched into one enqueue per phase. */
int b2p_grid_push_half_b(b2p_grid* g);
int b2p_grid_push_e(b2p_grid* g);
int b2p_grid_add_current(b2p_grid* g);
int b2p_grid_filter_current(b2p_grid* g);
int b2p_grid_push_particles(b2p_grid* g);
int b2p_grid_pack_outgoing_particles(b2p_grid* g);
int b2p_grid_sort_particles(b2p_grid* g);
int b2p_grid_deposit_current(b2p_grid* g);
/* grid_apply_edge_bcs / prtcl_reflect_particles / prtcl_advance_reflector_walls over all local
 * tiles (runko/simulation.py lap functions used by projects/pic-shock/pic.py:232-279). */
int b2p_grid_apply_edge_bcs(b2p_grid* g, int mode);
int b2p_grid_reflect_particles(b2p_grid* g);
int b2p_grid_advance_reflector_walls(b2p_grid* g);
/* Alive particles (id != dead) per species over the local tiles.  No reference getter: the reference
 * counts them while writing particle snapshots (io/snapshots/mpiio_particles.c++); used by the
 * conservation checks of bench.py and the tests. */
int b2p_grid_alive_counts(b2p_grid* g, uint64_t* counts);
/* mpiio::FieldsWriter<3>::write (src/runko/io/snapshots/mpiio_fields.c++:221-400, header
 * mpiio_header.h:56-82): "<prefix>/flds_<lap>.bin" = 512-byte "RNKO" v3
 * header + (9 + min(nspecies, 5)) dense fp32 arrays [nz][ny][nx] (x fastest): E, B point-sampled
 * every `stride` cells, J summed over stride^3 blocks, n_s = alive particles per coarse cell.
 * Every rank writes its own tiles with pwrite() into the shared file (POSIX instead of MPI-IO);
 * rank 0 also writes the header.  Readable by runko/mpiio_reader.py. */
int b2p_grid_write_fields_snapshot(b2p_grid* g, const char* prefix, int32_t lap, int32_t stride, int32_t nspecies);
/* One lap of projects/pic-turbulence/pic.py:187-221 (diagnostics/IO excluded);
 * sort when lap % 5 == 0. */
int b2p_grid_step_pic(b2p_grid* g, int64_t lap);
/* One lap of projects/emf-wave/emf.py:48-62. */
int b2p_grid_step_emf(b2p_grid* g);
/* Σ over local tiles of total_energy_{B,E}, Σ kinetic energy and Σ container
 * sizes per species (io/pic_average_kinetic_energy.h:24-155,
 * io/emf_average_field_energy_density.h). */
int b2p_grid_energies(b2p_grid* g, double* energy_B, double* energy_E,
                      double* kinetic /*n_species*/, uint64_t* sizes /*n_species*/);
/* Synthetic uniform thermal plasma generated on the device (bench workloads
 * too large to stage through the host): ppc particles per cell per species,
 * positions cell corner + U[0,1)^3, momenta Maxwellian with spread `delgam`,
 * counter-based RNG seeded by (seed, tile, species). No reference equivalent. */
int b2p_grid_inject_thermal(b2p_grid* g, int ppc, double delgam, uint64_t seed);
/* Synthetic drifting plasma for the beam / shock bench workloads, generated on the device: ppc particles of species
 * sp per cell in the cells pic::Tile::batch_inject_in_x_stripe(sp, pgen, x_left, x_right) would visit
 * (pic/tile.c++:235-322), appended after each container's last alive particle (containers pre-sized by
 * prealloc_per_species take them in their dead slots); momenta: Maxwellian of spread `delgam` boosted by
 * `gamma_drift` along dir_sign * x.  The same seed gives every species the same positions.  No reference equivalent
 * (the reference's drivers generate on the host: projects/pic-shock/pic.py:136-149, beam.py:123-137). */
int b2p_grid_inject_drifting_stripe(b2p_grid* g, int sp, int ppc, double delgam, double gamma_drift, int dir_sign,
                                    double x_left, double x_right, uint64_t seed);
int b2p_grid_set_uniform_B(b2p_grid* g, float bx, float by, float bz);

/* ---- multi-GPU (replaces corgi's MPI transport, corgi.h:1560-1692) ------- */
/* 128-byte NCCL unique id, created on rank 0 and distributed by the caller. */
int b2p_nccl_unique_id(void* id128);
/* owner[cid] = rank owning tile cid = i + Nx*(j + Ny*k) (corgi.h:283-315),
 * i.e. pycorgi Grid.get_mpi_grid. Creates the communicator and the exchange plan. */
int b2p_grid_comm_init(b2p_grid* g, int rank, int nranks, const void* id128, const int32_t* owner);
/* recv_data + send_data + wait_data of one mode (runko/simulation.py:295-319),
 * including the number_of_particles handshake for B2P_COMM_PIC_PARTICLE. */
int b2p_grid_external_communication(b2p_grid* g, int mode);

/* Host-only description of rank `rank`'s exchange plan (no GPU, no NCCL; used by the
 * world_size-2 gloo tests): rows of 7 int64 {peer, my cid, direction index, remote cid,
 * send_key, recv_key, floats of the interior-edge slab}; returns the row count. */
int64_t b2p_plan_describe(const b2p_config* cfg, const int32_t* owner, int rank, int64_t* rows, int64_t cap);

/* ---- timing helpers for bench.py (CUDA events on the library's stream) --- */
int b2p_timer_start(void);
int b2p_timer_stop(float* ms);
/* number of kernels this library has launched since process start */
uint64_t b2p_launch_count(void);
/* bytes this library has copied host->device / device->host since process start */
void b2p_copy_bytes(uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* wall-clock milliseconds the calling host thread has spent blocked in stream synchronisations inside this
 * library since process start (bench.py: host wall time per step minus this = time spent enqueueing) */
void b2p_host_wait_ms(double* ms);
/* Per-kernel-class device timing (CUDA events on the launch stream around every
 * launch of the class).  enable(1) clears and starts, enable(0) stops; report()
 * fills arrays of b2p_profile_num_classes() entries: total ms, launches and
 * processed units (particles or cells) per class. */
int b2p_profile_enable(int on);
int b2p_profile_num_classes(void);
const char* b2p_profile_class_name(int k);
int b2p_profile_report(double* ms, uint64_t* launches, double* units);


/* ---- self-checks of arithmetic building blocks (tests only) -------------- */
/* out[i] = x[i] / c computed by the pushers' constant-divisor division (particles.cu: DivC),
 * ref[i] = x[i] / c by the IEEE division; host pointers. The parity tests require out == ref
 * bit for bit (pic/particle_boris.h:46,53 divide by cfl). */
int b2p_selfcheck_const_division(const float* x, uint64_t n, float c, float* out, float* ref);

#ifdef __cplusplus
}
#endif
#endif /* B200PIC_H */
