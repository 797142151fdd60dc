/* pic_oracle.h — C-ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a single-source, dependency-free CPU
 * restatement of the reference's PIC-step arithmetic (runko @ e302899e) used
 * as the parity checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Nothing under runko_b200/ may import,
 * link or execute it.
 *
 * Parity pin (two independent ones):
 *  1. the reference's OWN kernel sources (emf::YeeLattice, pic::ParticleContainer)
 *     are compiled here from /root/reference into oracle/_ref/libref_kernels.so
 *     (oracle/Makefile.ref, oracle/ref_harness.cpp) and this oracle must match
 *     them BIT FOR BIT on seeded inputs for every kernel and for whole laps on
 *     periodic multi-tile grids: tests/test_oracle_vs_reference_build.py;
 *  2. the reference's own unit tests (tests/py/test_{emf,pic}*.py, 103 cases) run
 *     unmodified against this oracle through tests/refshim:
 *     tests/test_reference_suite.py.
 * The reference ships no numeric golden vectors for this path;
 * tests/golden/*.npz freeze the pinned behaviour for the GPU box.
 */
#ifndef PIC_ORACLE_H
#define PIC_ORACLE_H
#include "../include/b200pic.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_grid orc_grid;

const char* orc_last_error(void);
/* Builds ALL n_tiles tiles of the global periodic grid in this process. */
orc_grid* orc_create(const b2p_config* cfg);
void      orc_destroy(orc_grid* g);
int       orc_num_tiles(const orc_grid* g);
/* tile handle = corgi cid = i + Nx*(j + Ny*k) */
int       orc_tile_cid(const orc_grid* g, int i, int j, int k);

int orc_tile_set_fields(orc_grid* g, int t, const float* E, const float* B, const float* J, int with_halo);
int orc_tile_get_fields(orc_grid* g, int t, float* E, float* B, float* J, int with_halo);
int orc_tile_push_half_b(orc_grid* g, int t);
int orc_tile_push_e(orc_grid* g, int t);
int orc_tile_add_current(orc_grid* g, int t);
int orc_tile_filter_current(orc_grid* g, int t);
int orc_tile_clear_current(orc_grid* g, int t);
int orc_tile_field_energy(orc_grid* g, int t, double* eB, double* eE);

int orc_tile_inject(orc_grid* g, int t, int sp, uint64_t n,
                    const double* x, const double* y, const double* z,
                    const double* ux, const double* uy, const double* uz);
int orc_tile_set_particles(orc_grid* g, int t, int sp, uint64_t n,
                           const float* x, const float* y, const float* z,
                           const float* ux, const float* uy, const float* uz,
                           const uint64_t* id);
int orc_tile_container_size(orc_grid* g, int t, int sp, uint64_t* n);
int orc_tile_get_particles(orc_grid* g, int t, int sp, int alive_only,
                           float* x, float* y, float* z,
                           float* ux, float* uy, float* uz,
                           uint64_t* id, uint64_t* n_out);
int orc_tile_push_particles(orc_grid* g, int t);
int orc_tile_deposit_current(orc_grid* g, int t);
int orc_tile_sort_particles(orc_grid* g, int t);
int orc_tile_pack_outgoing_particles(orc_grid* g, int t);
int orc_tile_sort_keys(orc_grid* g, int t, int sp, uint32_t* keys);
int orc_tile_get_outgoing(orc_grid* g, int t, b2p_particle_state* buf, uint64_t cap,
                          uint64_t* ends, uint64_t* n_out);
int orc_tile_kinetic_energy(orc_grid* g, int t, int sp, double* energy, uint64_t* container_size);
/* Interpolated E,B at n positions (global coordinates) on tile t: out is [n][6]. */
/* pic-shock boundary pieces: emf/tile.c++:808-847 + emf/yee_lattice.c++:263-306; pic/reflector_wall.c++ */
int orc_tile_register_edge_bc(orc_grid* g, int t, const b2p_edge_bc* bc);
int orc_tile_apply_edge_bcs(orc_grid* g, int t, int mode);
int orc_tile_apply_edge_bc(orc_grid* g, int t, const b2p_edge_bc* bc, int mode);
int orc_tile_register_reflector_wall(orc_grid* g, int t, const b2p_reflector_wall* wall);
int orc_tile_reflect_particles(orc_grid* g, int t);
int orc_tile_advance_reflector_walls(orc_grid* g, int t);
int orc_tile_reflector_walls(orc_grid* g, int t, b2p_reflector_wall* out, uint64_t cap, uint64_t* n);

int orc_tile_interpolate(orc_grid* g, int t, uint64_t n, const float* x, const float* y,
                         const float* z, float* out);

int orc_local_communication(orc_grid* g, int mode);
/* all-tile phases, `threads` workers over tiles (one tile per worker at a time:
 * the reference's CPU execution model, serial inside a tile). */
int orc_grid_phase(orc_grid* g, const char* phase, int threads);
int orc_step_pic(orc_grid* g, int64_t lap, int threads);
int orc_step_emf(orc_grid* g, int threads);
/* emf/tile.c++:566-791 */
int orc_tile_register_antenna(orc_grid* g, int t, const b2p_antenna_mode* mode);
int orc_tile_deposit_antenna_current(orc_grid* g, int t);
/* io/snapshots/mpiio_fields.c++:221-400 + mpiio_header.h:56-82 */
int orc_write_fields_snapshot(orc_grid* g, const char* prefix, int32_t lap, int32_t stride, int32_t nspecies);
int orc_energies(orc_grid* g, double* eB, double* eE, double* kinetic, uint64_t* sizes);

#ifdef __cplusplus
}
#endif
#endif
