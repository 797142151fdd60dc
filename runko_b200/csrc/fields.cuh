// fields.cuh — launchers of the Yee-lattice kernels (fields.cu).
#pragma once
#include "common.cuh"

namespace b2p {
void launch_push_b_fdtd2(const FieldPtrs* tiles, int ntiles, const Geom& g, float dt, bool twice = false);   // twice: two half pushes from the same E in one pass
void launch_push_b_stencil(const FieldPtrs* tiles, int ntiles, const Geom& g, float dt, const float M[3][3][5]);
void launch_push_e_fdtd2(const FieldPtrs* tiles, int ntiles, const Geom& g, float dt, bool add_current);
void launch_add_current(const FieldPtrs* tiles, int ntiles, const Geom& g);
// filter_tiles: device array of {const float* src; float* dst;} per tile
void launch_filter(const void* filter_tiles, int ntiles, const Geom& g, bool unrolled);
void launch_zero(float* p, size_t n);
// emf::Tile::deposit_antenna_current (emf/tile.c++:578-777): J += -cfl * curl_backward(curl_forward(vec_pot)) in the
// interior (+ vec_pot itself outside it, as the reference adds the whole scratch lattice); modes passed by value
constexpr int ANTENNA_MAX_MODES = 48;
struct AntennaModes { float A[ANTENNA_MAX_MODES][3]; float K[ANTENNA_MAX_MODES][3]; float W[ANTENNA_MAX_MODES][2]; int n; };
void launch_antenna(float* J, float* vec_pot, float* gen_B, const Geom& g, const AntennaModes& m, const double mins[3],
                    const double maxs[3], float cfl_neg);
// FieldsWriter<3>::pack_tile, E/B/J part (io/snapshots/mpiio_fields.c++:221-275): buf[nf][nzt][nyt][nxt]
void launch_pack_snapshot(const FieldPtrs& f, const Geom& g, int stride, int nxt, int nyt, int nzt, int nf, float* buf);
// YeeLattice::apply_edge_bc (emf/yee_lattice.c++:263-306): masked components of `field` over the box [lo, hi)
struct EdgeBcOp { float* f; int3 lo, hi; unsigned mask; float3 v; };   // one edge BC on one lattice (box of the haloed lattice)
void launch_edge_bc_batch(const EdgeBcOp* ops /*device*/, int nops, size_t max_cells, const Geom& g);   // operations on different lattices
void launch_edge_bc(float* field, const Geom& g, const int lo[3], const int hi[3], unsigned mask, const float v[3]);
// J = J + add over n floats (YeeLattice::deposit_current(VecGrid), emf/yee_lattice.c++:361-375)
void launch_add_lattice(float* J, const float* add, size_t n);
// which: 0=E 1=B 2=J; nbr: device int[ntiles][27] (tile-table index of the neighbour, -1 = remote)
// nbr codes: >= 0 local tile slot; -1 none; <= -2 remote, staged slab remote[-(code+2)] (comm.cu)
// part 0: every halo cell; 1: only those fed by a tile of this rank; 2: only those fed by a remote rank's staged slab
void launch_halo_fill(const FieldPtrs* tiles, const int* nbr, int ntiles, const Geom& g, int which, const SlabDesc* remote, int part = 0);
void launch_J_exchange(const FieldPtrs* tiles, const int* nbr, int ntiles, const Geom& g, const SlabDesc* remote, int tile0 = 0);   // tiles [tile0, tile0 + ntiles) of the table
void launch_field_energy(const FieldPtrs* tiles, int ntiles, const Geom& g, double* out);
}  // namespace b2p
