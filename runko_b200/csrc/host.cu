// host.cu — process context, tile/grid objects and the per-phase orchestration
// (what runko/simulation.py + corgi do around the kernels, batched per phase).
#include "host.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <numeric>

namespace b2p {

// ------------------------------------------------------------------ context --
static Context g_ctx;
static bool g_ctx_ready = false;

static void init_context(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(B2P_ERR_CUDA, std::string("libb200pic: no usable CUDA device (") + cudaGetErrorString(e) +
                                "); this library has no CPU fallback");
  if (device < 0 || device >= count) throw Error(B2P_ERR_CUDA, "libb200pic: device index out of range");
  B2P_CUDA(cudaSetDevice(device));
  g_ctx.device = device;
  B2P_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  B2P_CUDA(cudaDeviceGetAttribute(&g_ctx.sm_count, cudaDevAttrMultiProcessorCount, device));
  B2P_CUDA(cudaEventCreate(&g_ctx.ev0));
  B2P_CUDA(cudaEventCreate(&g_ctx.ev1));
  // keep freed blocks in the stream-ordered pool instead of returning them to the OS
  cudaMemPool_t pool;
  B2P_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long thr = ~0ull;
  B2P_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  g_ctx_ready = true;
}
void init(int device) {
  if (g_ctx_ready) {
    if (device != g_ctx.device) throw Error(B2P_ERR_RUNTIME, "libb200pic: device already selected for this process");
    return;
  }
  init_context(device);
}
Context& ctx() {
  if (!g_ctx_ready) init_context(0);
  return g_ctx;
}
void count_launch(int n) { g_ctx.launches += n; }

static Tuning g_tuning;
static bool set_option(const std::string& name, int value) {
  if (name == "push_minb") g_tuning.push_minb = value;
  else if (name == "deposit_minb") g_tuning.deposit_minb = value;
  else if (name == "deposit_agg") g_tuning.deposit_agg = value;
  else if (name == "agg_min") g_tuning.agg_min = value;
  else if (name == "fuse_deposit") g_tuning.fuse_deposit = value;
  else if (name == "filter_chunk") g_tuning.filter_chunk = value;
  else if (name == "filter_ahead") g_tuning.filter_ahead = value;
  else if (name == "stencil_minb") g_tuning.stencil_minb = value;
  else if (name == "energy_cache") g_tuning.energy_cache = value;
  else if (name == "push_streams") g_tuning.push_streams = value;
  else if (name == "sort_streams") g_tuning.sort_streams = value;
  else if (name == "sort_batch") g_tuning.sort_batch = value;
  else if (name == "comm_overlap") g_tuning.comm_overlap = value;
  else if (name == "filter_pairs") g_tuning.filter_pairs = value;
  else if (name == "push_group") g_tuning.push_group = value;
  else if (name == "push_block") g_tuning.push_block = value;
  else if (name == "sort_overlap") g_tuning.sort_overlap = value;
  else if (name == "sort_counting") g_tuning.sort_counting = value;
  else if (name == "defer_tile_calls") g_tuning.defer_tile_calls = value;
  else return false;
  return true;
}
Tuning& tuning() {
  static bool parsed = false;
  if (!parsed) {
    parsed = true;
    if (const char* e = std::getenv("B2P_OPTS")) {       // "name=value,name=value"
      std::string str(e);
      size_t pos = 0;
      while (pos < str.size()) {
        const size_t end = std::min(str.find(',', pos), str.size());
        const std::string kv = str.substr(pos, end - pos);
        const size_t eq = kv.find('=');
        if (eq != std::string::npos && !set_option(kv.substr(0, eq), std::atoi(kv.c_str() + eq + 1)))
          std::fprintf(stderr, "[b2p] unknown option in B2P_OPTS: %s\n", kv.c_str());
        pos = end + 1;
      }
    }
  }
  return g_tuning;
}

void* dmalloc(size_t bytes) {
  void* p = nullptr;
  B2P_CUDA(cudaMallocAsync(&p, bytes ? bytes : 1, ctx().stream));
  return p;
}
void dfree(void* p) {
  if (p && g_ctx_ready) cudaFreeAsync(p, g_ctx.stream);
}
void timed_stream_sync() {
  const auto t0 = std::chrono::steady_clock::now();
  B2P_CUDA(cudaStreamSynchronize(ctx().stream));
  g_ctx.host_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
static void stream_sync() { timed_stream_sync(); }
template <class T> static void h2d(T* dst, const T* src, size_t n) {
  if (n) { B2P_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx().stream)); g_ctx.h2d_bytes += n * sizeof(T); }
}
template <class T> static void d2h(T* dst, const T* src, size_t n) {
  if (!n) return;
  // a device-to-host copy into pageable memory returns only when the data has landed: the wait for the stream to get there
  // is booked as blocked time, like a synchronisation
  const auto t0 = std::chrono::steady_clock::now();
  B2P_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, ctx().stream));
  g_ctx.host_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  g_ctx.d2h_bytes += n * sizeof(T);
}

// ----------------------------------------------------------------- profiler --
struct ProfRecord { int kc; cudaEvent_t a, b; double units; };
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t take_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; B2P_CUDA(cudaEventCreate(&e)); return e;
}
const char* kernel_class_name(int k) {
  static const char* n[KC_COUNT] = { "nodal_means", "push", "deposit", "sort_count", "sort_place", "sort_gather", "detect_leavers",
                                     "gather_outgoing", "append", "zero", "push_b", "push_e", "add_current", "filter",
                                     "halo_fill", "J_exchange", "energy", "edge_gather", "other", "nccl_exchange" };
  return (k >= 0 && k < KC_COUNT) ? n[k] : "?";
}
ProfScope::ProfScope(KernelClass k, double units) {
  if (!g_prof_on) return;
  ProfRecord r{ int(k), take_event(), take_event(), units };
  cudaEventRecord(r.a, ctx().stream);
  idx = int(g_prof.size());
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_prof[idx].b, ctx().stream);
}

// process-level scratch shared by all tiles (all work is ordered on one stream)
constexpr int MAX_WORKERS = 4;
struct SortBatch {              // scratch of one batch of containers in the counting sort (sort.cu)
  DBuf<unsigned> keys, rank, members;   // one slice per container, sized by its n
  DBuf<unsigned> cnt, offs;             // one slice of nkeys + 2 counters per container
  DBuf<SortJob> table;
  std::vector<Container> spare;         // gather targets (storage swapped with the sorted containers)
};
struct Scratch {
  SortBatch sort_batch[2];        // two batches in flight (alternating worker streams)
  DBuf<float4> nodal_w[MAX_WORKERS];   // per worker stream: nodal field means of the tile being pushed
  DBuf<float4> edges_w[MAX_WORKERS];   // per worker stream: cell-edge current accumulators
  DBuf<float4>& nodal = nodal_w[0];
  DBuf<float4>& edges = edges_w[0];
  DBuf<unsigned> keys_w[MAX_WORKERS][2], vals_w[MAX_WORKERS][2];   // per worker stream: sort keys / permutation
  DBuf<unsigned char> cub_temp_w[MAX_WORKERS];
  DBuf<unsigned> (&keys)[2] = keys_w[0];
  DBuf<unsigned char>& cub_temp = cub_temp_w[0];
  DBuf<unsigned> counters;        // pack: P per container | leaver total per tile | seg[27][nseg] per container
  DBuf<unsigned long long> pack_ends;   // pack: subregion ends per container [27]
  DBuf<unsigned char> table;      // device staging for job / out-tile tables
  DBuf<double> energy;
  DBuf<float> antenna[2];         // vec_pot_buff_ / generated_B_buff_ of the tile depositing its antenna current
  Container spare_w[MAX_WORKERS]; // per worker stream: gather target of the sort (swapped with the container)
};
static Scratch& scratch() { static Scratch* s = new Scratch; return *s; }

// the kinetic-energy account of the species of container `c` of tile `t`, if the grid's account is running
static double* ke_account(const b2p_tile* t, const Container& c) {
  b2p_grid* g = t->grid;
  if (!g || !g->ke_valid) return nullptr;
  return g->ke_acc.p + size_t(&c - t->sp.data()) * KE_SLOTS;
}
// a container of the tile changes in a way the account does not follow
static void ke_void(const b2p_tile* t) { if (t->grid) t->grid->ke_valid = false; }

// Worker streams: the per-tile particle phase (nodal means -> zero -> push per species -> edge
// gather) is round-robined over a few streams so that one tile's small kernels and launch gaps
// hide behind another tile's push.  Forked from / joined to the library stream with events, so
// the library stays stream-ordered for the caller.
struct Workers {
  cudaStream_t s[MAX_WORKERS] = {};
  cudaEvent_t fork = nullptr, fork2 = nullptr, join[MAX_WORKERS] = {};
  bool ready = false;
  int sort_pending = 0;       // worker streams still carry an un-joined sort (join_pending_sort)
  void init() {
    if (ready) return;
    for (int w = 0; w < MAX_WORKERS; ++w) {
      B2P_CUDA(cudaStreamCreateWithFlags(&s[w], cudaStreamNonBlocking));
      B2P_CUDA(cudaEventCreateWithFlags(&join[w], cudaEventDisableTiming));
    }
    B2P_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    B2P_CUDA(cudaEventCreateWithFlags(&fork2, cudaEventDisableTiming));
    ready = true;
  }
};
static Workers& workers() { static Workers w; return w; }
// The sort of b2p_grid_step_pic is left running on the worker streams while the library stream goes on
// with the field phase of the lap (J exchange, filters, B and E updates — none of which touches a particle
// container).  Every C-ABI entry point in this file, and the deposit's fresh-deposit branch, joins it first.
void join_pending_sort() {
  Workers& wk = workers();
  if (!wk.sort_pending) return;
  for (int w = 0; w < wk.sort_pending; ++w) {
    B2P_CUDA(cudaEventRecord(wk.join[w], wk.s[w]));
    B2P_CUDA(cudaStreamWaitEvent(g_ctx.stream, wk.join[w], 0));
  }
  wk.sort_pending = 0;
}
struct StreamScope {          // makes `st` the stream every launcher / ProfScope / DBuf uses
  cudaStream_t saved;
  explicit StreamScope(cudaStream_t st) : saved(g_ctx.stream) { g_ctx.stream = st; }
  ~StreamScope() { g_ctx.stream = saved; }
};

// ---------------------------------------------------------------- container --
void Container::reserve(size_t cap, bool exact) {
  if (cap <= capacity()) return;
  if (!exact)
  // uniform capacity classes (multiples of 256 Ki slots incl. 1/16 slack): containers of a
  // uniform plasma end up with identical capacities, so the sort's spare set never reallocates
  cap = ((cap + cap / 16 + 262143) / 262144) * 262144;
  if (std::getenv("B2P_TRACE")) std::fprintf(stderr, "[b2p trace] container realloc %zu -> %zu slots (n=%u)\n", capacity(), cap, n);
  x.reserve_exact(cap, n); y.reserve_exact(cap, n); z.reserve_exact(cap, n);
  ux.reserve_exact(cap, n); uy.reserve_exact(cap, n); uz.reserve_exact(cap, n);
  id.reserve_exact(cap, n);
}

uint2* Container::mask_words() {
  const size_t words = (size_t(n) + 255) / 256 * 8;
  masks.reserve(std::max<size_t>(words, 8));
  return masks.p;
}

static void swap_storage(Container& a, Container& b) {
  std::swap(a.x.p, b.x.p); std::swap(a.x.cap, b.x.cap);
  std::swap(a.y.p, b.y.p); std::swap(a.y.cap, b.y.cap);
  std::swap(a.z.p, b.z.p); std::swap(a.z.cap, b.z.cap);
  std::swap(a.ux.p, b.ux.p); std::swap(a.ux.cap, b.ux.cap);
  std::swap(a.uy.p, b.uy.p); std::swap(a.uy.cap, b.uy.cap);
  std::swap(a.uz.p, b.uz.p); std::swap(a.uz.cap, b.uz.cap);
  std::swap(a.id.p, b.id.p); std::swap(a.id.cap, b.id.cap);
}

// P = 1 + last alive slot (pic/particle.h:469-488); full pass when not cached
static unsigned find_P(Container& c) {
  if (c.P_valid) return c.P;
  unsigned P = 0;
  if (c.n) {
    Scratch& s = scratch();
    s.counters.reserve(64);
    B2P_CUDA(cudaMemsetAsync(s.counters.p, 0, sizeof(unsigned), ctx().stream));
    launch_last_alive(c.id.p, c.n, s.counters.p);
    d2h(&P, s.counters.p, 1);
    stream_sync();
  }
  c.P = P; c.P_valid = true;
  return P;
}

}  // namespace b2p

using namespace b2p;

// --------------------------------------------------------------------- tile --
const FieldPtrs* b2p_tile::device_entry() {
  d_fp.reserve(1);
  if (fp_dirty) {
    const FieldPtrs f = ptrs();
    h2d(d_fp.p, &f, 1);
    fp_dirty = false;
  }
  return d_fp.p;
}

const FieldPtrs* b2p_grid::device_table() {
  if (table_dirty) {
    std::vector<FieldPtrs> h(tiles.size());
    for (size_t i = 0; i < tiles.size(); ++i) h[i] = tiles[i]->ptrs();
    d_tiles.reserve(std::max<size_t>(1, h.size()));
    h2d(d_tiles.p, h.data(), h.size());
    table_dirty = false;
  }
  return d_tiles.p;
}

static int wrapi(int v, int n) { while (v < 0) v += n; while (v >= n) v -= n; return v; }   // corgi/tile.h:126-136

const int* b2p_grid::device_nbr() {
  if (nbr_dirty) {
    std::vector<int> h(tiles.size() * 27, -1);
    for (size_t t = 0; t < tiles.size(); ++t)
      for (int ir = -1; ir <= 1; ++ir) for (int jr = -1; jr <= 1; ++jr) for (int kr = -1; kr <= 1; ++kr) {
        const int* T = cfg.n_tiles;
        const int c = cid(wrapi(tiles[t]->idx[0] + ir, T[0]), wrapi(tiles[t]->idx[1] + jr, T[1]), wrapi(tiles[t]->idx[2] + kr, T[2]));
        const int di = ((ir + 1) * 3 + (jr + 1)) * 3 + (kr + 1);
        int code = slot_of_cid[c];
        int entry = 0;
        if (code < 0 && comm_remote_entry(this, int(t), di, &entry)) code = -(2 + entry);
        h[t * 27 + di] = code;
      }
    d_nbr.reserve(std::max<size_t>(1, h.size()));
    h2d(d_nbr.p, h.data(), h.size());
    nbr_dirty = false;
  }
  return d_nbr.p;
}

namespace b2p {

// emf/stencil_coefficients.h:43-64
static float stencil_alpha(const float M[3][5]) {
  float sum = 0.0f;
  sum += 3.0f * (M[1][0] + 2.0f * (M[1][1] + M[1][2] + M[1][3] + M[1][4]));
  sum += 5.0f * (M[2][0] + 2.0f * (M[2][1] + M[2][2] + M[2][3] + M[2][4]));
  sum += 1.0f * 2.0f * (M[0][1] + M[0][2] + M[0][3] + M[0][4]);
  return 1.0f - sum;
}

static void validate_config(const b2p_config& c) {
  for (int d = 0; d < 3; ++d) {
    if (c.n_cells[d] < H)
      throw Error(B2P_ERR_RUNTIME, "Yee Lattice extents (" + std::to_string(c.n_cells[0]) + ", " + std::to_string(c.n_cells[1]) +
                                     ", " + std::to_string(c.n_cells[2]) + ") are assumed to be at least halo size: 3");
    if (c.n_tiles[d] < 1) throw Error(B2P_ERR_RUNTIME, "n_tiles must be positive");
  }
  if (c.field_propagator != B2P_PROPAGATOR_FDTD2 && c.field_propagator != B2P_PROPAGATOR_STENCIL)
    throw Error(B2P_ERR_RUNTIME, "not supported field propagator.");
  if (c.n_species < 0 || c.n_species > B2P_MAX_SPECIES) throw Error(B2P_ERR_RUNTIME, "unsupported number of species");
  if (size_t(c.n_tiles[0]) * c.n_tiles[1] * c.n_tiles[2] >= (size_t(1) << 24))
    throw Error(B2P_ERR_RUNTIME, "PIC tile does not support this many tiles.");   // pic/tile.c++:138-140
  const size_t Ch = size_t(c.n_cells[0] + 6) * (c.n_cells[1] + 6) * (c.n_cells[2] + 6);
  if (Ch >= (size_t(1) << 31)) throw Error(B2P_ERR_RUNTIME, "tile lattice too large for 32-bit cell indexing");
}

static Geom make_geom(const b2p_config& c) {
  Geom g;
  g.Ch = 1;
  for (int d = 0; d < 3; ++d) { g.N[d] = c.n_cells[d]; g.Hx[d] = c.n_cells[d] + 2 * H; g.Ch *= unsigned(g.Hx[d]); }
  return g;
}

static b2p_tile* create_tile(const b2p_config& cfg, const int32_t idx[3]) {
  validate_config(cfg);
  for (int d = 0; d < 3; ++d)
    if (idx[d] < 0 || idx[d] >= cfg.n_tiles[d])
      throw Error(B2P_ERR_RUNTIME, "Trying to create tile outside of configured grid.");   // emf/tile.c++:155-157
  ctx();
  std::unique_ptr<b2p_tile> t(new b2p_tile);
  t->cfg = cfg;
  t->g = make_geom(cfg);
  std::memcpy(t->stencilM, cfg.stencil, sizeof(t->stencilM));
  for (int a = 0; a < 3; ++a) t->stencilM[a][0][0] = stencil_alpha(t->stencilM[a]);
  for (int d = 0; d < 3; ++d) {
    t->idx[d] = idx[d];
    t->mins[d] = double(size_t(idx[d]) * size_t(cfg.n_cells[d]));              // emf/tile.c++:162-170
    t->maxs[d] = double((size_t(idx[d]) + 1) * size_t(cfg.n_cells[d]));
    t->origo[d] = static_cast<float>(t->mins[d]) - float(H);
  }
  const size_t nf = t->lattice_floats();
  t->E.reserve(nf); t->B.reserve(nf); t->Jbuf[0].reserve(nf); t->Jbuf[1].reserve(nf);
  launch_zero(t->E.p, nf); launch_zero(t->B.p, nf); launch_zero(t->Jbuf[0].p, nf); launch_zero(t->Jbuf[1].p, nf);
  t->sp.resize(cfg.n_species);
  t->next_ordinal.assign(cfg.n_species, 0);
  for (int s = 0; s < cfg.n_species; ++s) {
    t->sp[s].charge = cfg.q[s]; t->sp[s].mass = cfg.m[s];
    if (cfg.prealloc_per_species) {                                               // pic/particle.c++:39-61
      if (cfg.prealloc_per_species >= (1ull << 32)) throw Error(B2P_ERR_RUNTIME, "prealloc_per_species exceeds uint32 indexing");
      t->sp[s].reserve(cfg.prealloc_per_species);
      t->sp[s].n = unsigned(cfg.prealloc_per_species);
      launch_fill_dead(t->sp[s].id.p, 0, t->sp[s].n);
      t->sp[s].P = 0; t->sp[s].P_valid = true;
    } else {
      t->sp[s].P = 0; t->sp[s].P_valid = true;
    }
  }
  t->tile_tag = (static_cast<unsigned long long>(idx[0]) * cfg.n_tiles[1] + idx[1]) * cfg.n_tiles[2] + idx[2];   // pic/tile.c++:141-143
  return t.release();
}

// ----------------------------------------------------------- field phases --
static float half_dt(const b2p_config& c) { return static_cast<float>(c.cfl / 2); }    // emf/tile.c++:365
static float full_dt(const b2p_config& c) { return static_cast<float>(c.cfl); }        // emf/tile.c++:384

void phase_push_half_b(const std::vector<b2p_tile*>& tiles, const FieldPtrs* table, int times) {
  if (tiles.empty()) return;
  b2p_tile* t0 = tiles[0];
  if (t0->cfg.field_propagator == B2P_PROPAGATOR_STENCIL) {
    for (int q = 0; q < times; ++q) launch_push_b_stencil(table, int(tiles.size()), t0->g, half_dt(t0->cfg), t0->stencilM);
  } else {
    // two back-to-back half pushes (emf.py:50-51) read the same E: one pass, same roundings
    for (int q = 0; q + 1 < times; q += 2) launch_push_b_fdtd2(table, int(tiles.size()), t0->g, half_dt(t0->cfg), true);
    if (times & 1) launch_push_b_fdtd2(table, int(tiles.size()), t0->g, half_dt(t0->cfg), false);
  }
}
void phase_push_e(const std::vector<b2p_tile*>& tiles, const FieldPtrs* table, bool add_current) {
  if (tiles.empty()) return;
  launch_push_e_fdtd2(table, int(tiles.size()), tiles[0]->g, full_dt(tiles[0]->cfg), add_current);
}
void phase_add_current(const std::vector<b2p_tile*>& tiles, const FieldPtrs* table) {
  if (tiles.empty()) return;
  launch_add_current(table, int(tiles.size()), tiles[0]->g);
}

// emf/tile.c++:405-426; out of place into the tile's second J buffer, then swap
void phase_filter(const std::vector<b2p_tile*>& tiles) {
  if (tiles.empty()) return;
  const int cf = tiles[0]->cfg.current_filter;
  if (cf != B2P_FILTER_BINOMIAL2 && cf != B2P_FILTER_BINOMIAL2_UNROLLED)
    throw Error(B2P_ERR_LOGIC, "Trying to filter current without specifying `current_filter`!");
  struct FT { const float* src; float* dst; };
  std::vector<FT> h(tiles.size());
  for (size_t i = 0; i < tiles.size(); ++i) h[i] = FT{ tiles[i]->Jbuf[tiles[i]->jcur].p, tiles[i]->Jbuf[1 - tiles[i]->jcur].p };
  Scratch& s = scratch();
  s.table.reserve(h.size() * sizeof(FT));
  h2d(reinterpret_cast<FT*>(s.table.p), h.data(), h.size());
  launch_filter(s.table.p, int(tiles.size()), tiles[0]->g, cf == B2P_FILTER_BINOMIAL2_UNROLLED);
  for (b2p_tile* t : tiles) {
    t->jcur = 1 - t->jcur;
    t->fp_dirty = true;
    if (t->grid) t->grid->table_dirty = true;
    t->pendJ_valid = false;   // the spare lattice was overwritten
  }
}

// -------------------------------------------------------- particle phases --
static int sign_of(double v) { return (0.0 < v) - (v < 0.0); }   // tools/math.h:181-183

// pic/tile.c++:326-365.  Every push also publishes the leaver/stayer ballots of the pushed
// positions (Container::masks), which pack_outgoing_particles consumes if nothing touched
// the container in between.
void phase_push_particles(const std::vector<b2p_tile*>& tiles, size_t n_first, const std::function<void()>& between) {
  Scratch& s = scratch();
  const bool fuse = tuning().fuse_deposit != 0;
  // Tiles that hold particles, in groups of up to `push_group` tiles of one geometry / pusher / cfl: the
  // kernels of the particle phase (nodal means, clearing the cell-edge scratch, the push of all the
  // group's containers, edge gather) run once per group, as launches large enough to fill the GPU.
  const int gmax = std::max(1, std::min(tuning().push_group, PUSH_GROUP_MAX));
  std::vector<std::vector<b2p_tile*>> groups;
  size_t groups_first = ~size_t(0);                   // groups [0, groups_first) hold the first n_first tiles
  for (size_t ti = 0; ti < tiles.size(); ++ti) {
    b2p_tile* t = tiles[ti];
    t->pendJ_valid = t->pend_packed = false;
    if (ti == n_first) groups_first = groups.size();
    bool any = false;
    for (const Container& c : t->sp) any = any || c.n;
    if (!any) continue;
    bool open = !groups.empty() && int(groups.back().size()) < gmax && groups.size() != groups_first;
    if (open) {
      const b2p_tile* f = groups.back().front();
      size_t nc = t->sp.size();
      for (const b2p_tile* q : groups.back()) nc += q->sp.size();
      open = std::memcmp(&f->g, &t->g, sizeof(Geom)) == 0 && f->cfg.particle_pusher == t->cfg.particle_pusher &&
             f->cfg.cfl == t->cfg.cfl && nc <= size_t(PUSH_JOBS_MAX);
    }
    if (!open) groups.emplace_back();
    groups.back().push_back(t);
  }
  // Kinetic-energy account: this push restarts it when it covers every tile of one grid exactly once
  b2p_grid* kg = nullptr;
  {
    for (b2p_tile* t : tiles) ke_void(t);
    b2p_grid* g0 = tiles.empty() ? nullptr : tiles[0]->grid;
    // ... and only for a caller that reads the energies every lap (the reference's lap does; a bare loop of laps does
    // not pay the 8 % the account costs the push)
    if (g0 && g0->ke_asked && tuning().energy_cache && fuse && tiles.size() == g0->tiles.size()) {
      static unsigned long long epoch = 0;
      ++epoch;
      bool all = true;
      for (b2p_tile* t : tiles) { all = all && t->grid == g0 && t->ke_epoch != epoch; t->ke_epoch = epoch; }
      if (all) { kg = g0; g0->ke_asked = false; }
    }
    if (kg) {
      const size_t nacc = size_t(std::max(1, kg->cfg.n_species)) * KE_SLOTS;
      kg->ke_acc.reserve(nacc);
      B2P_CUDA(cudaMemsetAsync(kg->ke_acc.p, 0, nacc * sizeof(double), ctx().stream));
    }
  }
  if (groups.empty()) { if (kg) kg->ke_valid = true; if (between) between(); return; }
  if (groups_first > groups.size()) groups_first = groups.size();
  const int nw = std::max(1, std::min({ tuning().push_streams, MAX_WORKERS, int(groups.size()) }));
  Workers& wk = workers();
  cudaStream_t main_stream = ctx().stream;
  if (nw > 1) {
    wk.init();
    B2P_CUDA(cudaEventRecord(wk.fork, main_stream));
    for (int w = 0; w < nw; ++w) B2P_CUDA(cudaStreamWaitEvent(wk.s[w], wk.fork, 0));
  }
  size_t gi = 0;
  static PushJobs* jobs = new PushJobs;             // 8.7 KB kernel-argument table, filled per group
  bool between_done = !between;
  for (const std::vector<b2p_tile*>& grp : groups) {
    if (!between_done && gi >= groups_first) {
      // everything after this point needs what `between` enqueues on the library stream
      between();
      between_done = true;
      if (nw > 1) {
        B2P_CUDA(cudaEventRecord(wk.fork2, main_stream));
        for (int w = 0; w < nw; ++w) B2P_CUDA(cudaStreamWaitEvent(wk.s[w], wk.fork2, 0));
      }
    }
    const int w = int(gi++ % size_t(nw));
    StreamScope on(nw > 1 ? wk.s[w] : main_stream);
    const Geom& g = grp.front()->g;
    const size_t nod_stride = (nodal_float4_per_node() * g.Ch + 1) & ~size_t(1), edge_stride = size_t(3) * g.Ch;   // float4 per tile
    s.nodal_w[w].reserve(nod_stride * grp.size());
    NodalBatch nb{};
    EdgeBatch eb{};
    for (size_t q = 0; q < grp.size(); ++q) {
      nb.E[q] = grp[q]->E.p; nb.B[q] = grp[q]->B.p; nb.nod[q] = s.nodal_w[w].p + q * nod_stride;
    }
    nb.n = int(grp.size());
    launch_nodal_means(nb, g);
    if (fuse) {
      s.edges_w[w].reserve(edge_stride * grp.size());
      launch_zero(reinterpret_cast<float*>(s.edges_w[w].p), size_t(4) * edge_stride * grp.size());
    }
    int nj = 0;
    unsigned max_n = 0;
    double slots = 0;
    for (size_t q = 0; q < grp.size(); ++q) {
      b2p_tile* t = grp[q];
      const float mn[3] = { float(t->mins[0]), float(t->mins[1]), float(t->mins[2]) };
      const float mx[3] = { float(t->maxs[0]), float(t->maxs[1]), float(t->maxs[2]) };
      float4* Jc = fuse ? s.edges_w[w].p + q * edge_stride : nullptr;
      for (Container& c : t->sp) {
        if (!c.n) continue;
        const float qm = static_cast<float>(sign_of(c.charge) / c.mass);
        jobs->job[nj++] = PushJob{ c.view(), nb.nod[q], Jc, c.mask_words(), make_float3(t->origo[0], t->origo[1], t->origo[2]),
                                   make_float3(mn[0], mn[1], mn[2]), make_float3(mx[0], mx[1], mx[2]), qm, static_cast<float>(c.charge),
                                   kg ? kg->ke_acc.p + size_t(&c - t->sp.data()) * KE_SLOTS : nullptr };
        max_n = std::max(max_n, c.n);
        slots += c.n;
        c.touch();
        c.masks_valid = true;
      }
      if (fuse) {
        eb.Jc[q] = Jc; eb.J[q] = t->Jbuf[1 - t->jcur].p;
        t->pendJ_valid = true;
      }
    }
    launch_push_jobs(grp.front()->cfg.particle_pusher, *jobs, nj, max_n, slots, g, static_cast<float>(grp.front()->cfg.cfl), fuse);
    if (fuse) {
      eb.n = int(grp.size());
      launch_edge_gather(eb, g);
    }
  }
  if (nw > 1)
    for (int w = 0; w < nw; ++w) {
      B2P_CUDA(cudaEventRecord(wk.join[w], wk.s[w]));
      B2P_CUDA(cudaStreamWaitEvent(main_stream, wk.join[w], 0));
    }
  if (!between_done) between();
  if (kg) kg->ke_valid = true;
}

// pic/tile.c++:369-415.  clear_current + scratch accumulate + `J += scratch`
// collapse to "zero J, accumulate into J" (0 + x == x).
void phase_deposit(const std::vector<b2p_tile*>& tiles) {
  Scratch& s = scratch();
  std::vector<b2p_tile*> fresh;
  for (b2p_tile* t : tiles) {
    if (t->pendJ_valid && t->pend_packed) {
      // the fused push (stayers) and the particle exchange (arrivals) already accumulated this
      // lap's current in the spare lattice: adopt it
      t->jcur = 1 - t->jcur;
      t->fp_dirty = true;
      if (t->grid) t->grid->table_dirty = true;
      t->pendJ_valid = t->pend_packed = false;
    } else {
      fresh.push_back(t);
      t->pendJ_valid = t->pend_packed = false;
    }
  }
  if (!fresh.empty()) {
    // Fresh deposits (tiles whose particles changed after the push: reflector wall, injection, uploads — and tiles
    // that hold no particles at all), in groups of one geometry: one clear of the group's cell-edge scratch, one
    // deposit per non-empty container, one edge gather (= clear_current + `J += generated_J`) per group.
    join_pending_sort();                                            // a fresh deposit reads the containers
    const int gmax = std::max(1, std::min(tuning().push_group, PUSH_GROUP_MAX));
    for (size_t b = 0; b < fresh.size();) {
      const Geom g = fresh[b]->g;
      size_t e = b;
      while (e < fresh.size() && e - b < size_t(gmax) && std::memcmp(&fresh[e]->g, &g, sizeof(Geom)) == 0) ++e;
      const size_t edge_stride = size_t(3) * g.Ch;
      s.edges.reserve(edge_stride * (e - b));
      EdgeBatch eb{};
      launch_zero(reinterpret_cast<float*>(s.edges.p), size_t(4) * edge_stride * (e - b));
      for (size_t q = b; q < e; ++q) {
        b2p_tile* t = fresh[q];
        float4* Jc = s.edges.p + (q - b) * edge_stride;
        for (Container& c : t->sp)
          if (c.n) launch_deposit(c.view(), Jc, t->g, t->origo, static_cast<float>(t->cfg.cfl), static_cast<float>(c.charge));
        eb.Jc[eb.n] = Jc; eb.J[eb.n] = t->J(); ++eb.n;              // a tile without particles gathers zeros: J = 0
      }
      if (eb.n) launch_edge_gather(eb, g);
      b = e;
    }
  }
  for (b2p_tile* t : tiles) {
    if (t->corr_pending) {                                          // pic/tile.c++:411-414
      launch_add_lattice(t->J(), t->corrJ.p, t->lattice_floats());
      t->corr_pending = false;
    }
  }
}

// ------------------------------------------------- pic-shock boundary pieces --
// emf/tile.c++:808-827
static bool edge_bc_width(const b2p_tile* t, const b2p_edge_bc& bc, size_t* width) {
  const int d = bc.direction;
  const float tile_min = static_cast<float>(t->mins[d]);
  const float tile_max = static_cast<float>(t->maxs[d]);
  const size_t Nd = size_t(t->g.N[d]);
  if (bc.side == 0) {
    if (bc.position <= tile_min) return false;
    if (bc.position >= tile_max) { *width = Nd; return true; }
    *width = static_cast<size_t>(bc.position - tile_min) + 1;
    return true;
  }
  if (bc.position >= tile_max) return false;
  if (bc.position <= tile_min) { *width = Nd; return true; }
  *width = Nd - static_cast<size_t>(bc.position - tile_min);
  return true;
}

// emf::Tile::apply_edge_bc (emf/tile.c++:835-840) + YeeLattice::apply_edge_bc (emf/yee_lattice.c++:263-306)
// the lattice, component mask, values and box of one BC on one tile; false: the BC does not reach this tile
static bool edge_bc_op(b2p_tile* t, const b2p_edge_bc& bc, int mode, EdgeBcOp* op) {
  if (bc.direction > 2) throw Error(B2P_ERR_RUNTIME, "edge_bc: direction must be 0, 1 or 2");
  float* field;
  unsigned mask;
  const float* v;
  switch (mode) {
    case B2P_COMM_EMF_E: field = t->E.p; mask = bc.E_components; v = bc.E; break;
    case B2P_COMM_EMF_B: field = t->B.p; mask = bc.B_components; v = bc.B; break;
    case B2P_COMM_EMF_J: field = t->J(); mask = bc.J_components; v = bc.J; break;
    default:
      throw Error(B2P_ERR_RUNTIME, "YeeLattice::apply_edge_bc does not support given communication mode: " + std::to_string(mode));
  }
  size_t width = 0;
  if (!edge_bc_width(t, bc, &width) || width == 0 || !(mask & 7u)) return false;
  const int d = bc.direction;
  const int Nd = t->g.N[d];
  const int w = int(std::min<size_t>(width, size_t(Nd)));
  int lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    if (a != d) { lo[a] = 0; hi[a] = t->g.Hx[a]; }
    else if (bc.side == 0) { lo[a] = 0; hi[a] = H + w; }
    else { lo[a] = H + Nd - w; hi[a] = t->g.Hx[a]; }
  }
  *op = EdgeBcOp{ field, make_int3(lo[0], lo[1], lo[2]), make_int3(hi[0], hi[1], hi[2]), mask, make_float3(v[0], v[1], v[2]) };
  return true;
}

// emf::Tile::apply_edge_bc (emf/tile.c++:835-840) + YeeLattice::apply_edge_bc (emf/yee_lattice.c++:263-306)
static void apply_edge_bc(b2p_tile* t, const b2p_edge_bc& bc, int mode) {
  EdgeBcOp op;
  if (!edge_bc_op(t, bc, mode, &op)) return;
  const int lo[3] = { op.lo.x, op.lo.y, op.lo.z }, hi[3] = { op.hi.x, op.hi.y, op.hi.z };
  const float v[3] = { op.v.x, op.v.y, op.v.z };
  launch_edge_bc(op.f, t->g, lo, hi, op.mask, v);
}

// emf/tile.c++:842-847 for many tiles: a tile's BCs apply in registration order, tiles are independent, so round r
// applies the r-th BC of every tile in one launch
void phase_apply_edge_bcs(const std::vector<b2p_tile*>& tiles, int mode) {
  if (tiles.size() == 1) {
    for (const b2p_edge_bc& bc : tiles[0]->edge_bcs) apply_edge_bc(tiles[0], bc, mode);
    return;
  }
  Scratch& s = scratch();
  size_t rounds = 0;
  for (b2p_tile* t : tiles) rounds = std::max(rounds, t->edge_bcs.size());
  for (size_t r = 0; r < rounds; ++r) {
    // group by geometry (the kernel takes one Geom)
    std::vector<EdgeBcOp> ops;
    size_t max_cells = 0;
    const Geom* g = nullptr;
    auto flush = [&] {
      if (ops.empty()) return;
      s.table.reserve(ops.size() * sizeof(EdgeBcOp));
      h2d(reinterpret_cast<EdgeBcOp*>(s.table.p), ops.data(), ops.size());
      launch_edge_bc_batch(reinterpret_cast<const EdgeBcOp*>(s.table.p), int(ops.size()), max_cells, *g);
      ops.clear(); max_cells = 0;
    };
    for (b2p_tile* t : tiles) {
      if (r >= t->edge_bcs.size()) continue;
      EdgeBcOp op;
      if (!edge_bc_op(t, t->edge_bcs[r], mode, &op)) continue;
      if (g && std::memcmp(g, &t->g, sizeof(Geom)) != 0) flush();
      g = &t->g;
      ops.push_back(op);
      max_cells = std::max(max_cells, size_t(op.hi.x - op.lo.x) * size_t(op.hi.y - op.lo.y) * size_t(op.hi.z - op.lo.z));
    }
    flush();
  }
}

// pic::Tile::reflect_particles (pic/reflector_wall.c++:241-284)
void phase_reflect_particles(const std::vector<b2p_tile*>& tiles) {
  for (b2p_tile* t : tiles) {
    if (t->walls.empty()) continue;
    auto wall_is_in_tile = [&](const b2p_reflector_wall& w) {
      return w.walloc >= float(t->mins[0]) - float(t->cfg.cfl) && w.walloc <= float(t->maxs[0]);
    };
    bool any = false;
    for (const b2p_reflector_wall& w : t->walls) any = any || wall_is_in_tile(w);
    if (!any) continue;
    t->corrJ.reserve(t->lattice_floats());
    t->corr_pending = true;
    launch_zero(t->corrJ.p, t->lattice_floats());
    // the particles change after the push: a current deposited inside the push no longer describes them
    t->pendJ_valid = t->pend_packed = false;
    ke_void(t);
    for (const b2p_reflector_wall& w : t->walls) {
      if (!wall_is_in_tile(w)) continue;
      for (Container& c : t->sp) {
        launch_reflect_at_wall(c.view(), t->corrJ.p, t->g, t->origo, static_cast<float>(t->cfg.cfl), w.walloc, w.betawall,
                               w.gammawall, static_cast<float>(c.charge));
        c.touch();
      }
    }
  }
}

// pic/tile.c++:419-438 + pic/particle.h:575-703: stable sort by cell key, dead last.
// Fast path: counting sort by cell (sort.cu), batched over containers; a container whose largest cell exceeded
// SORT_RADIX_POP at its previous sort (page-locked hint, no host round trip) takes this general radix sort instead.
static void sort_radix(b2p_tile* t, Container& c, int w) {
  Scratch& s = scratch();
  for (int b = 0; b < 2; ++b) { s.keys_w[w][b].reserve(c.n); s.vals_w[w][b].reserve(c.n); }
  // Alive keys are < Ch (particles live inside the haloed lattice), so dead slots are keyed
  // Ch instead of UINT32_MAX and only bits(Ch) key bits are sorted: same stable order,
  // one radix pass fewer.  (Keys >= Ch — positions outside the lattice, undefined
  // behaviour in the reference — are clamped to Ch.)
  int key_bits = 1;
  while ((1ull << key_bits) <= t->g.Ch) ++key_bits;
  launch_sort_keys(c.view(), t->g, t->origo, s.keys_w[w][0].p, s.vals_w[w][0].p, t->g.Ch);
  const size_t tb = sort_pairs_temp_bytes(c.n, key_bits);
  s.cub_temp_w[w].reserve(tb);
  unsigned* k[2] = { s.keys_w[w][0].p, s.keys_w[w][1].p };
  unsigned* v[2] = { s.vals_w[w][0].p, s.vals_w[w][1].p };
  const int sel = sort_pairs(s.cub_temp_w[w].p, tb, k, v, c.n, key_bits);
  Container& spare = s.spare_w[w];
  spare.reserve(c.capacity(), /*exact=*/true);
  spare.n = c.n;
  launch_gather(c.view(), spare.view(), v[sel]);
  swap_storage(c, spare);
  c.touch();
}

// Largest cell population seen by the previous counting sort of a container, copied to page-locked
// host memory without a synchronisation.  It only steers the choice between the counting sort
// (correct for any input, quadratic in the cell population) and the radix sort, so a value that is
// one sort old — or has not landed yet — is as good as a fresh one.
struct PopHints {
  static constexpr unsigned CHUNK = 4096, UNKNOWN = 0xFFFFFFFFu;
  std::vector<unsigned*> chunks;
  unsigned used = 0;
  unsigned* slot(int& idx) {
    if (idx < 0) {
      if (used % CHUNK == 0) {
        unsigned* c = nullptr;
        B2P_CUDA(cudaMallocHost(&c, CHUNK * sizeof(unsigned)));
        for (unsigned q = 0; q < CHUNK; ++q) c[q] = UNKNOWN;
        chunks.push_back(c);
      }
      idx = int(used++);
    }
    return chunks[size_t(idx) / CHUNK] + size_t(idx) % CHUNK;
  }
};
static PopHints& pop_hints() { static PopHints* h = new PopHints; return *h; }

void phase_sort(const std::vector<b2p_tile*>& tiles, bool leave_running) {
  Scratch& s = scratch();
  struct Item { b2p_tile* t; Container* c; };
  std::vector<Item> items, crowded;
  const bool counting = tuning().sort_counting != 0;
  for (b2p_tile* t : tiles)
    for (Container& c : t->sp) {
      if (c.n < 2) continue;
      volatile unsigned* hint = pop_hints().slot(c.pop_hint_slot);
      if (!counting || (*hint != PopHints::UNKNOWN && *hint > SORT_RADIX_POP)) crowded.push_back(Item{ t, &c });
      else items.push_back(Item{ t, &c });
    }
  if (items.empty() && crowded.empty()) return;
  // The sort runs on worker streams — batches alternate between two of them, so that the atomic-latency-bound
  // count of one batch overlaps the bandwidth-bound gather of the other — and b2p_grid_step_pic can leave it
  // running under the field phase of the lap (leave_running); launches are per batch, not per container.
  Workers& wk = workers();
  cudaStream_t main_stream = ctx().stream;
  const int nws = std::max(0, std::min(tuning().sort_streams, 2));
  if (nws) {
    wk.init();
    B2P_CUDA(cudaEventRecord(wk.fork, main_stream));
    for (int w = 0; w < nws; ++w) B2P_CUDA(cudaStreamWaitEvent(wk.s[w], wk.fork, 0));
  }
  {
    const size_t B = size_t(std::max(1, tuning().sort_batch));
    size_t q = 0, nbatch = 0;
    while (q < items.size()) {
      const int w = nws ? int(nbatch % size_t(nws)) : 0;
      ++nbatch;
      StreamScope on(nws ? wk.s[w] : main_stream);
      SortBatch& sb = s.sort_batch[w];
      if (sb.spare.size() < B) sb.spare.resize(B);
      // a batch: up to B containers of one lattice geometry
      const Geom& g = items[q].t->g;
      const unsigned nkeys = g.Ch;
      std::vector<Item> batch;
      while (q < items.size() && batch.size() < B && std::memcmp(&items[q].t->g, &g, sizeof(Geom)) == 0) batch.push_back(items[q++]);
      // per container: nkeys + 2 counters, then the scan's chunk totals + 1 (zeroed together with the counters)
      const size_t ncnt = (size_t(nkeys) + 2 + 63) & ~size_t(63);
      const size_t ncount = ncnt + ((size_t(sort_scan_chunks(nkeys)) + 1 + 63) & ~size_t(63));
      size_t total = 0;
      unsigned max_n = 0;
      std::vector<size_t> off(batch.size());
      for (size_t b = 0; b < batch.size(); ++b) { off[b] = total; total += (size_t(batch[b].c->n) + 63) & ~size_t(63); max_n = std::max(max_n, batch[b].c->n); }
      sb.keys.reserve(total); sb.rank.reserve(total); sb.members.reserve(total);
      sb.cnt.reserve(ncount * batch.size()); sb.offs.reserve(ncount * batch.size());
      sb.table.reserve(B);
      bool first_timer = false;
      auto build = [&](const std::vector<Item>& bt, std::vector<SortJob>& jobs, double& slots) {
        jobs.clear(); slots = 0; max_n = 0;
        for (size_t b = 0; b < bt.size(); ++b) {
          Container& c = *bt[b].c;
          Container& spare = sb.spare[b];
          spare.reserve(c.capacity(), /*exact=*/true);
          spare.n = c.n;
          volatile unsigned* hint = pop_hints().slot(c.pop_hint_slot);
          jobs.push_back(SortJob{ c.view(), spare.view(), sb.keys.p + off[b], sb.rank.p + off[b], sb.members.p + off[b],
                                  sb.cnt.p + ncount * b, sb.cnt.p + ncount * b + ncnt, sb.offs.p + ncount * b, const_cast<unsigned*>(hint),
                                  make_float3(bt[b].t->origo[0], bt[b].t->origo[1], bt[b].t->origo[2]) });
          slots += c.n;
          max_n = std::max(max_n, c.n);
        }
      };
      for (const Item& it : batch) first_timer = first_timer || *pop_hints().slot(it.c->pop_hint_slot) == PopHints::UNKNOWN;
      std::vector<SortJob> jobs;
      double slots = 0;
      build(batch, jobs, slots);
      h2d(sb.table.p, jobs.data(), jobs.size());
      B2P_CUDA(cudaMemsetAsync(sb.cnt.p, 0, ncount * batch.size() * sizeof(unsigned), ctx().stream));
      // the scan leaves every container's largest cell population in its page-locked hint slot (zero-copy store)
      launch_sort_count_scan(sb.table.p, int(jobs.size()), max_n, slots, g, nkeys);
      if (first_timer) {
        // no history for some container of the batch: wait for the populations once and re-route the crowded ones
        B2P_CUDA(cudaStreamSynchronize(ctx().stream));
        std::vector<Item> keep;
        std::vector<size_t> keep_off;
        for (size_t b = 0; b < batch.size(); ++b) {
          if (*pop_hints().slot(batch[b].c->pop_hint_slot) > SORT_RADIX_POP) crowded.push_back(batch[b]);
          else { keep.push_back(batch[b]); keep_off.push_back(b); }
        }
        if (keep.size() != batch.size()) {
          // the scratch slices stay where the count / scan kernels filled them: rebuild the table for the kept jobs only
          std::vector<SortJob> kj;
          for (size_t i = 0; i < keep.size(); ++i) {
            SortJob j = jobs[keep_off[i]];
            Container& spare = sb.spare[i];
            spare.reserve(keep[i].c->capacity(), /*exact=*/true);
            spare.n = keep[i].c->n;
            j.dst = spare.view();
            kj.push_back(j);
          }
          batch.swap(keep);
          jobs.swap(kj);
          slots = 0; max_n = 0;
          for (const Item& it : batch) { slots += it.c->n; max_n = std::max(max_n, it.c->n); }
          if (!jobs.empty()) h2d(sb.table.p, jobs.data(), jobs.size());
        }
      }
      if (!jobs.empty()) {
        launch_sort_scatter_place(sb.table.p, int(jobs.size()), max_n, slots, nkeys);
        for (size_t b = 0; b < batch.size(); ++b) { swap_storage(*batch[b].c, sb.spare[b]); batch[b].c->touch(); }
      }
    }
    StreamScope on_crowded(nws ? wk.s[0] : main_stream);
    for (const Item& it : crowded) {
      sort_radix(it.t, *it.c, 0);
      // let a crowded container return to the counting sort once it has thinned out: probe again every few sorts
      volatile unsigned* hint = pop_hints().slot(it.c->pop_hint_slot);
      if (counting && ++it.c->radix_sorts_since_probe >= 8) { it.c->radix_sorts_since_probe = 0; *hint = PopHints::UNKNOWN; }
    }
  }
  if (nws) {
    wk.sort_pending = nws;
    if (!leave_running || !tuning().sort_overlap) join_pending_sort();
  }
}

// pic/tile_communication.c++:68-96 + pic/particle.c++:199-348, batched over tiles
void phase_pack_outgoing(const std::vector<b2p_tile*>& tiles) {
  if (tiles.empty()) return;
  Scratch& s = scratch();
  for (b2p_tile* t : tiles) { t->out_ends.assign(27 * t->sp.size(), 0); t->out_count = 0; }
  std::vector<Container*> conts;
  std::vector<PackJob> jobs;
  std::vector<PackTile> ptiles;
  std::vector<b2p_tile*> owners;            // tiles that hold at least one container, in ptiles order
  size_t total_slots = 0, seg_total = 0;
  unsigned max_nseg = 0;
  for (b2p_tile* t : tiles) {
    const float mn[3] = { float(t->mins[0]), float(t->mins[1]), float(t->mins[2]) };
    const float mx[3] = { float(t->maxs[0]), float(t->maxs[1]), float(t->maxs[2]) };
    if (t->sp.empty()) continue;
    ptiles.push_back(PackTile{ unsigned(jobs.size()), unsigned(t->sp.size()), nullptr });
    owners.push_back(t);
    for (Container& c : t->sp) {
      uint2* words = c.mask_words();
      if (!c.masks_valid && c.n) {                                     // not pushed since the last change
        launch_make_masks(c.view(), words, mn, mx);
        t->pendJ_valid = false;
      }
      PackJob jb{};
      jb.masks = words;
      jb.nwords = unsigned((size_t(c.n) + 31) / 32);
      jb.nseg = pack_segments(c.n);
      jb.s = c.view();
      jb.mn = make_float3(mn[0], mn[1], mn[2]); jb.mx = make_float3(mx[0], mx[1], mx[2]);
      jb.ke = ke_account(t, c);
      jobs.push_back(jb);
      conts.push_back(&c);
      total_slots += c.n;
      seg_total += size_t(27) * jb.nseg;
      max_nseg = std::max(max_nseg, jb.nseg);
    }
  }
  const size_t nc = conts.size(), nt = ptiles.size();
  if (nc == 0) return;
  // device scratch: per container P and seg[27][nseg]; per tile the leaver total; per container ends[27] (u64)
  s.counters.reserve(nc + nt + seg_total);
  s.pack_ends.reserve(nc * 27);
  unsigned* d_P = s.counters.p;
  unsigned* d_tot = s.counters.p + nc;
  unsigned* d_seg = s.counters.p + nc + nt;
  size_t so = 0;
  for (size_t c = 0; c < nc; ++c) {
    jobs[c].last_alive = d_P + c;
    jobs[c].seg = d_seg + so;
    jobs[c].ends = s.pack_ends.p + 27 * c;
    so += size_t(27) * jobs[c].nseg;
  }
  for (size_t q = 0; q < nt; ++q) ptiles[q].total = d_tot + q;
  s.table.reserve(nc * sizeof(PackJob) + nt * sizeof(PackTile) + 64);
  PackJob* d_jobs = reinterpret_cast<PackJob*>(s.table.p);
  PackTile* d_tiles = reinterpret_cast<PackTile*>(s.table.p + ((nc * sizeof(PackJob) + 15) & ~size_t(15)));
  h2d(d_jobs, jobs.data(), nc);
  h2d(d_tiles, ptiles.data(), nt);
  B2P_CUDA(cudaMemsetAsync(d_P, 0, nc * sizeof(unsigned), ctx().stream));
  if (max_nseg) launch_pack_count_scan(d_jobs, unsigned(nc), max_nseg, d_tiles, unsigned(nt), double(total_slots));
  // the one host round trip of the pack: P per container, leaver totals per tile, subregion ends
  std::vector<unsigned> hc(nc + nt, 0);
  std::vector<unsigned long long> hends(nc * 27, 0);
  d2h(hc.data(), s.counters.p, nc + nt);
  if (max_nseg) d2h(hends.data(), s.pack_ends.p, nc * 27);
  stream_sync();
  unsigned long long total = 0;
  for (size_t c = 0; c < nc; ++c) { conts[c]->P = hc[c]; conts[c]->P_valid = true; }
  size_t c = 0;
  for (size_t q = 0; q < nt; ++q) {
    b2p_tile* t = owners[q];
    const unsigned tile_total = max_nseg ? hc[nc + q] : 0u;
    t->out_count = tile_total;
    total += tile_total;
    t->out_buf.reserve(std::max<size_t>(tile_total, 1));
    for (size_t sp = 0; sp < t->sp.size(); ++sp, ++c) {
      jobs[c].out = t->out_buf.p;
      for (int r = 0; r < 27; ++r) t->out_ends[27 * sp + r] = hends[27 * c + r];   // pic/particle.c++:327-343
    }
  }
  if (total) {
    h2d(d_jobs, jobs.data(), nc);                                    // now with the tiles' buffers
    launch_pack_write(d_jobs, unsigned(nc), max_nseg, double(total));
    for (Container* ct : conts) ct->masks_valid = false;              // leavers are dead now
  }
  for (b2p_tile* t : tiles) t->pend_packed = t->pendJ_valid;
}

// ---------------------------------------------------------- communication --
struct SpanRef { const b2p_particle_state* p; unsigned n; };

static void append_spans(std::vector<Container*>& conts, std::vector<b2p_tile*>& owners, std::vector<std::vector<SpanRef>>& spans,
                         bool wrap, const float wmin[3], const float wmax[3]) {
  // pic/particle.h:454-572 for many containers at once
  std::vector<AppendJobHost> jobs;
  unsigned max_count = 0;
  for (size_t c = 0; c < conts.size(); ++c) {
    if (spans[c].empty()) continue;
    Container& ct = *conts[c];
    const unsigned P = find_P(ct);
    size_t total = 0;
    for (const SpanRef& sr : spans[c]) total += sr.n;
    if (size_t(P) + total >= (size_t(1) << 32)) throw Error(B2P_ERR_RUNTIME, "particle container exceeds uint32 indexing");
    ct.reserve(size_t(P) + total);
    // keep only [0,P): slots beyond are dead and get overwritten or dropped
    ct.n = unsigned(P + total);
    unsigned off = P;
    for (const SpanRef& sr : spans[c]) {
      if (sr.n) {
        b2p_tile* t = owners[c];
        float* jp = (t->pendJ_valid && t->pend_packed) ? t->Jbuf[1 - t->jcur].p : nullptr;   // arrivals join the fused deposit
        jobs.push_back(AppendJobHost{ sr.p, sr.n, off, ct.view(), jp, make_float3(t->origo[0], t->origo[1], t->origo[2]),
                                      static_cast<float>(ct.charge), ke_account(t, ct) });
        max_count = std::max(max_count, sr.n);
      }
      off += sr.n;
    }
    ct.P = ct.n; ct.P_valid = true;   // either the last appended slot is alive, or n == P
    ct.masks_valid = false;
  }
  if (jobs.empty()) return;
  Scratch& s = scratch();
  s.table.reserve(jobs.size() * sizeof(AppendJobHost));
  h2d(reinterpret_cast<AppendJobHost*>(s.table.p), jobs.data(), jobs.size());
  for (size_t b = 0; b < jobs.size(); b += 65535) {
    const int nj = int(std::min<size_t>(65535, jobs.size() - b));
    launch_append(reinterpret_cast<AppendJobHost*>(s.table.p) + b, nj, max_count, wrap, wmin, wmax, owners[0]->g,
                  static_cast<float>(owners[0]->cfg.cfl));
  }
}

// corgi::Grid::local_communication (external/corgi/src/corgi/corgi.h:1697-1718)
// emf_J_exchange for the tiles [first, first + count) of the grid
static void grid_J_exchange(b2p_grid* g, size_t first, size_t count) {
  launch_J_exchange(g->device_table(), g->device_nbr(), int(count), g->g, static_cast<const SlabDesc*>(comm_remote_table(g, 1)), int(first));
}

void grid_local_communication(b2p_grid* g, int mode, int part) {
  const int nt = int(g->tiles.size());
  if (!nt) return;
  switch (mode) {
    case B2P_COMM_EMF_E: launch_halo_fill(g->device_table(), g->device_nbr(), nt, g->g, 0, static_cast<const SlabDesc*>(comm_remote_table(g, 0)), part); return;
    case B2P_COMM_EMF_B: launch_halo_fill(g->device_table(), g->device_nbr(), nt, g->g, 1, static_cast<const SlabDesc*>(comm_remote_table(g, 0)), part); return;
    case B2P_COMM_EMF_J: launch_halo_fill(g->device_table(), g->device_nbr(), nt, g->g, 2, static_cast<const SlabDesc*>(comm_remote_table(g, 0)), part); return;
    case B2P_COMM_EMF_J_EXCHANGE: grid_J_exchange(g, 0, size_t(nt)); return;
    case B2P_COMM_PIC_PARTICLE: break;
    default:
      throw Error(B2P_ERR_LOGIC, "local_communication does not support given communication mode: " + std::to_string(mode));
  }
  // pic/tile_communication.c++:121-195: for every Moore direction (kr, jr, ir order) take
  // the neighbour's span for the inverted direction; then the postlude appends (:100-117)
  std::vector<Container*> conts;
  std::vector<b2p_tile*> owners;
  std::vector<std::vector<SpanRef>> spans;
  const int* T = g->cfg.n_tiles;
  for (b2p_tile* me : g->tiles) {
    const size_t base = conts.size();
    for (Container& c : me->sp) { conts.push_back(&c); owners.push_back(me); spans.emplace_back(); }
    for (int kr = -1; kr <= 1; ++kr) for (int jr = -1; jr <= 1; ++jr) for (int ir = -1; ir <= 1; ++ir) {
      if (!ir && !jr && !kr) continue;
      const int oc = g->cid(wrapi(me->idx[0] + ir, T[0]), wrapi(me->idx[1] + jr, T[1]), wrapi(me->idx[2] + kr, T[2]));
      const int os = g->slot_of_cid[oc];
      if (os < 0) {
        // remote neighbour: spans received by the external exchange (the reference reads the
        // pic::VirtualTile buffer here, pic/tile_communication.c++:166-181)
        int entry = 0;
        if (!comm_remote_entry(g, me->slot, ((ir + 1) * 3 + (jr + 1)) * 3 + (kr + 1), &entry)) continue;
        for (size_t q = 0; q < me->sp.size(); ++q) {
          const b2p_particle_state* ptr = nullptr; unsigned cnt = 0;
          comm_particle_span(g, entry, int(q), &ptr, &cnt);
          spans[base + q].push_back(SpanRef{ ptr, cnt });
        }
        continue;
      }
      b2p_tile* other = g->tiles[os];
      if (other->out_ends.size() != 27 * other->sp.size())
        throw Error(B2P_ERR_LOGIC, "pic_particle communication requires pack_outgoing_particles first");
      const int inv = ((-ir + 1) * 3 + (-jr + 1)) * 3 + (-kr + 1);
      for (size_t q = 0; q < me->sp.size(); ++q) {
        const size_t index = 27 * q + inv;
        const unsigned long long end = other->out_ends[index];
        const unsigned long long begin = index == 0 ? 0 : other->out_ends[index - 1];
        spans[base + q].push_back(SpanRef{ other->out_buf.p + begin, unsigned(end - begin) });
      }
    }
  }
  float wmin[3], wmax[3];
  for (int d = 0; d < 3; ++d) { wmin[d] = 0.0f; wmax[d] = static_cast<float>(double(size_t(T[d]) * size_t(g->cfg.n_cells[d]))); }
  append_spans(conts, owners, spans, true, wmin, wmax);
  for (b2p_tile* t : g->tiles) t->out_ends.clear();
  comm_particles_consumed(g);
}

}  // namespace b2p

// -------------------------------------------------- deferred per-tile calls --
// runko/simulation.py drives a lap as `for tile in tiles: tile.<op>()`.  Tile-local operations of
// the same kind on different tiles are independent, so consecutive per-tile calls of one kind are
// collected and executed as ONE batched phase (one launch over all collected tiles for the field
// kernels; worker streams, one exchange of counters, ... for the particle phases) as soon as any
// other entry point is called — every C-ABI entry flushes first, so callers observe exactly the
// synchronous semantics of the reference.
namespace b2p {
enum DeferKind { DK_NONE, DK_PUSH_HALF_B, DK_PUSH_E, DK_ADD_CURRENT, DK_FILTER, DK_PUSH, DK_PACK, DK_SORT, DK_DEPOSIT };
static int g_defer_kind = DK_NONE;
static std::vector<b2p_tile*> g_defer_tiles;
static DBuf<FieldPtrs>& defer_table() { static DBuf<FieldPtrs>* t = new DBuf<FieldPtrs>; return *t; }

static const FieldPtrs* table_for(const std::vector<b2p_tile*>& tiles) {
  b2p_grid* g = tiles[0]->grid;
  if (g && g->tiles == tiles) return g->device_table();
  if (tiles.size() == 1) return tiles[0]->device_entry();
  std::vector<FieldPtrs> h(tiles.size());
  for (size_t i = 0; i < tiles.size(); ++i) h[i] = tiles[i]->ptrs();
  defer_table().reserve(h.size());
  h2d(defer_table().p, h.data(), h.size());
  stream_sync();                                   // `h` is pageable host memory about to go out of scope
  return defer_table().p;
}

// A batch that fails when it is finally executed inside an entry point that cannot report (destroy,
// b2p_launch_count) leaves its error here; the next reporting entry point (b2p_sync included) returns it.
static int g_sticky_code = 0;
static std::string g_sticky_msg;
void record_sticky_error(int code, const std::string& msg) {
  if (!g_sticky_code) { g_sticky_code = code; g_sticky_msg = msg; }
}
void throw_sticky_error() {
  if (!g_sticky_code) return;
  const int c = g_sticky_code;
  g_sticky_code = 0;
  throw Error(c, "deferred tile call failed: " + g_sticky_msg);
}
static void flush_quietly() {               // for entry points without a status: keep the error for the next one
  try { join_pending_sort(); flush_deferred(); }
  catch (const Error& e) { record_sticky_error(e.code, e.what()); }
  catch (const std::exception& e) { record_sticky_error(B2P_ERR_RUNTIME, e.what()); }
}

void flush_deferred() {
  join_pending_sort();                      // a sort left on the worker streams owns the containers until it is joined
  if (g_defer_kind == DK_NONE) return;
  const int kind = g_defer_kind;
  std::vector<b2p_tile*> tiles;
  tiles.swap(g_defer_tiles);
  g_defer_kind = DK_NONE;
  for (b2p_tile* t : tiles) t->deferred = false;
  switch (kind) {
    case DK_PUSH_HALF_B: phase_push_half_b(tiles, table_for(tiles)); break;
    case DK_PUSH_E: phase_push_e(tiles, table_for(tiles), false); break;
    case DK_ADD_CURRENT: phase_add_current(tiles, table_for(tiles)); break;
    case DK_FILTER: phase_filter(tiles); break;
    case DK_PUSH: phase_push_particles(tiles); break;
    case DK_PACK: phase_pack_outgoing(tiles); break;
    case DK_SORT: phase_sort(tiles, false); break;
    case DK_DEPOSIT: phase_deposit(tiles); break;
    default: break;
  }
}

// Tiles may share a batch only if every parameter a batched phase reads from tiles[0] is the same
// for all of them: geometry, cfl, propagator + stencil coefficients, filter, pusher.
static bool same_batch_config(const b2p_tile* a, const b2p_tile* b) {
  return a->grid == b->grid && std::memcmp(&a->g, &b->g, sizeof(Geom)) == 0 && a->cfg.cfl == b->cfg.cfl &&
         a->cfg.field_propagator == b->cfg.field_propagator && a->cfg.current_filter == b->cfg.current_filter &&
         a->cfg.particle_pusher == b->cfg.particle_pusher && std::memcmp(a->stencilM, b->stencilM, sizeof(a->stencilM)) == 0;
}

static void defer(int kind, b2p_tile* t) {
  ctx();                                    // no usable device: fail at the call, not at the flush
  if (!tuning().defer_tile_calls) {
    flush_deferred();
    g_defer_kind = kind; g_defer_tiles.assign(1, t);
    flush_deferred();
    return;
  }
  if (g_defer_kind != kind || t->deferred || (!g_defer_tiles.empty() && !same_batch_config(g_defer_tiles[0], t)))
    flush_deferred();
  g_defer_kind = kind;
  g_defer_tiles.push_back(t);
  t->deferred = true;
}
}  // namespace b2p

// ==================================================================== C ABI ==
static thread_local std::string g_last_error;
namespace b2p {
void set_last_error(const std::string& s) { g_last_error = s; }
void Scratch_table_upload(const void* src, size_t bytes) {
  Scratch& s = scratch();
  s.table.reserve(bytes);
  B2P_CUDA(cudaMemcpyAsync(s.table.p, src, bytes, cudaMemcpyHostToDevice, ctx().stream));
  g_ctx.h2d_bytes += bytes;
}
const void* Scratch_table_ptr() { return scratch().table.p; }
}  // namespace b2p
#define B2P_TRY try { b2p::join_pending_sort(); b2p::flush_deferred(); b2p::throw_sticky_error();
#define B2P_TRY_DEFER try {
#define B2P_CATCH                                                              \
  }                                                                            \
  catch (const b2p::Error& e) { g_last_error = e.what(); return e.code; }     \
  catch (const std::exception& e) { g_last_error = e.what(); return B2P_ERR_RUNTIME; } \
  return B2P_OK;

static b2p_tile* T(b2p_tile* t) { if (!t) throw Error(B2P_ERR_RUNTIME, "null tile handle"); return t; }
static b2p_grid* G(b2p_grid* g) { if (!g) throw Error(B2P_ERR_RUNTIME, "null grid handle"); return g; }
static Container& C(b2p_tile* t, int sp) {
  if (sp < 0 || sp >= int(T(t)->sp.size())) throw Error(B2P_ERR_RUNTIME, "particle type " + std::to_string(sp) + " is not configured");
  return t->sp[sp];
}
namespace b2p { void init(int device); }

extern "C" {

const char* b2p_last_error(void) { return g_last_error.c_str(); }
const char* b2p_version(void) { return "b200pic 0.1 (sm_100a)"; }
int b2p_init(int device) { B2P_TRY b2p::init(device); B2P_CATCH }
int b2p_sync(void) { B2P_TRY stream_sync(); B2P_CATCH }
int b2p_set_option(const char* name, int value) {
  B2P_TRY
  b2p::tuning();
  if (!name || !b2p::set_option(name, value)) throw Error(B2P_ERR_RUNTIME, std::string("unknown option: ") + (name ? name : "(null)"));
  B2P_CATCH
}
int b2p_host_register(void* ptr, size_t bytes) {
  B2P_TRY
  ctx();
  B2P_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  B2P_CATCH
}
int b2p_host_unregister(void* ptr) { B2P_TRY B2P_CUDA(cudaHostUnregister(ptr)); B2P_CATCH }
int64_t b2p_gpu_mem_kB(void) {
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return -1;
  return int64_t((tot - fr) / 1024);
}

int b2p_tile_create(const b2p_config* cfg, const int32_t idx[3], b2p_tile** out) {
  B2P_TRY
  if (!cfg || !idx || !out) throw Error(B2P_ERR_RUNTIME, "null argument");
  *out = create_tile(*cfg, idx);
  B2P_CATCH
}
void b2p_tile_destroy(b2p_tile* t) {
  if (!t) return;
  b2p::flush_quietly();                     // also joins a sort still running on the worker streams before the buffers are freed
  if (t->grid) {
    b2p_grid* g = t->grid;
    g->ke_valid = false;
    auto it = std::find(g->tiles.begin(), g->tiles.end(), t);
    if (it != g->tiles.end()) {
      g->tiles.erase(it);
      std::fill(g->slot_of_cid.begin(), g->slot_of_cid.end(), -1);
      for (size_t i = 0; i < g->tiles.size(); ++i) {
        g->tiles[i]->slot = int(i);
        g->slot_of_cid[g->cid(g->tiles[i]->idx[0], g->tiles[i]->idx[1], g->tiles[i]->idx[2])] = int(i);
      }
      g->table_dirty = g->nbr_dirty = true;
    }
  }
  delete t;
}
int b2p_tile_bounds(const b2p_tile* t, double mins[3], double maxs[3]) {
  B2P_TRY
  for (int d = 0; d < 3; ++d) { mins[d] = t->mins[d]; maxs[d] = t->maxs[d]; }
  B2P_CATCH
}

static void upload_field(b2p_tile* t, float* dev, const float* host, int with_halo) {
  if (!host) return;
  const Geom& g = t->g;
  if (with_halo) { h2d(dev, host, t->lattice_floats()); stream_sync(); return; }
  std::vector<float> tmp(t->lattice_floats());
  d2h(tmp.data(), dev, tmp.size());
  stream_sync();
  const size_t Ni = size_t(g.N[0]) * g.N[1] * g.N[2];
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < g.N[0]; ++i) for (int j = 0; j < g.N[1]; ++j)
      std::memcpy(&tmp[c * size_t(g.Ch) + (size_t(i + H) * g.Hx[1] + (j + H)) * g.Hx[2] + H],
                  &host[c * Ni + (size_t(i) * g.N[1] + j) * g.N[2]], sizeof(float) * g.N[2]);
  h2d(dev, tmp.data(), tmp.size());
  stream_sync();
}
static void download_field(b2p_tile* t, const float* dev, float* host, int with_halo) {
  if (!host) return;
  const Geom& g = t->g;
  if (with_halo) { d2h(host, dev, t->lattice_floats()); stream_sync(); return; }
  std::vector<float> tmp(t->lattice_floats());
  d2h(tmp.data(), dev, tmp.size());
  stream_sync();
  const size_t Ni = size_t(g.N[0]) * g.N[1] * g.N[2];
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < g.N[0]; ++i) for (int j = 0; j < g.N[1]; ++j)
      std::memcpy(&host[c * Ni + (size_t(i) * g.N[1] + j) * g.N[2]],
                  &tmp[c * size_t(g.Ch) + (size_t(i + H) * g.Hx[1] + (j + H)) * g.Hx[2] + H], sizeof(float) * g.N[2]);
}

int b2p_tile_set_fields(b2p_tile* t, const float* E, const float* B, const float* J, int with_halo) {
  B2P_TRY
  T(t);
  upload_field(t, t->E.p, E, with_halo); upload_field(t, t->B.p, B, with_halo); upload_field(t, t->J(), J, with_halo);
  B2P_CATCH
}
int b2p_tile_get_fields(b2p_tile* t, float* E, float* B, float* J, int with_halo) {
  B2P_TRY
  T(t);
  download_field(t, t->E.p, E, with_halo); download_field(t, t->B.p, B, with_halo); download_field(t, t->J(), J, with_halo);
  B2P_CATCH
}
int b2p_tile_push_half_b(b2p_tile* t) { B2P_TRY_DEFER defer(DK_PUSH_HALF_B, T(t)); B2P_CATCH }
int b2p_tile_push_e(b2p_tile* t) { B2P_TRY_DEFER defer(DK_PUSH_E, T(t)); B2P_CATCH }
int b2p_tile_add_current(b2p_tile* t) { B2P_TRY_DEFER defer(DK_ADD_CURRENT, T(t)); B2P_CATCH }
int b2p_tile_filter_current(b2p_tile* t) {
  B2P_TRY_DEFER
  const int cf = T(t)->cfg.current_filter;                                        // emf/tile.c++:407-411
  if (cf != B2P_FILTER_BINOMIAL2 && cf != B2P_FILTER_BINOMIAL2_UNROLLED)
    throw Error(B2P_ERR_LOGIC, "Trying to filter current without specifying `current_filter`!");
  defer(DK_FILTER, t);
  B2P_CATCH
}
int b2p_tile_clear_current(b2p_tile* t) { B2P_TRY launch_zero(T(t)->J(), t->lattice_floats()); B2P_CATCH }
int b2p_tile_field_energy(b2p_tile* t, double* eB, double* eE) {
  B2P_TRY
  T(t);
  Scratch& s = scratch();
  s.energy.reserve(2);
  launch_field_energy(t->device_entry(), 1, t->g, s.energy.p);
  double h[2];
  d2h(h, s.energy.p, 2);
  stream_sync();
  if (eB) *eB = h[0] / 2.0;                                                       // emf/yee_lattice.c++:402-403
  if (eE) *eE = h[1] / 2.0;
  B2P_CATCH
}

int b2p_tile_inject(b2p_tile* t, int sp, uint64_t n, const double* x, const double* y, const double* z,
                    const double* ux, const double* uy, const double* uz) {
  B2P_TRY
  Container& c = C(t, sp);
  t->pendJ_valid = false;
  ke_void(t);
  // pic/tile.c++:207-217 + pic/particle.h:258-287: narrow, assign ids, append after the last alive slot
  const unsigned P = find_P(c);
  if (size_t(P) + n >= (size_t(1) << 32)) throw Error(B2P_ERR_RUNTIME, "particle container exceeds uint32 indexing");
  c.reserve(size_t(P) + n);
  c.n = unsigned(P + n);
  if (n) {
    std::vector<float> f(n);
    const double* src[6] = { x, y, z, ux, uy, uz };
    float* dst[6] = { c.x.p, c.y.p, c.z.p, c.ux.p, c.uy.p, c.uz.p };
    for (int a = 0; a < 6; ++a) {
      for (uint64_t i = 0; i < n; ++i) f[i] = static_cast<float>(src[a][i]);
      h2d(dst[a] + P, f.data(), n);
      stream_sync();
    }
    std::vector<unsigned long long> ids(n);
    for (uint64_t i = 0; i < n; ++i) {
      const unsigned long long ordinal = t->next_ordinal[sp]++;
      if (ordinal >= (2ull << 40)) throw Error(B2P_ERR_RUNTIME, "PIC tile ran out of particle ids (2^40 per species)!");
      ids[i] = (t->tile_tag << 40) | ordinal;                                     // pic/tile.c++:470-481
    }
    h2d(c.id.p + P, ids.data(), n);
    stream_sync();
  }
  c.P = c.n; c.P_valid = true; c.masks_valid = false;
  B2P_CATCH
}

int b2p_tile_set_particles(b2p_tile* t, int sp, uint64_t n, const float* x, const float* y, const float* z,
                           const float* ux, const float* uy, const float* uz, const uint64_t* id) {
  B2P_TRY
  Container& c = C(t, sp);
  t->pendJ_valid = false;
  ke_void(t);
  if (n >= (1ull << 32)) throw Error(B2P_ERR_RUNTIME, "particle container exceeds uint32 indexing");
  c.n = 0;
  c.reserve(n);
  c.n = unsigned(n);
  h2d(c.x.p, x, n); h2d(c.y.p, y, n); h2d(c.z.p, z, n);
  h2d(c.ux.p, ux, n); h2d(c.uy.p, uy, n); h2d(c.uz.p, uz, n);
  h2d(c.id.p, reinterpret_cast<const unsigned long long*>(id), n);
  stream_sync();
  c.touch();
  B2P_CATCH
}
int b2p_tile_container_size(b2p_tile* t, int sp, uint64_t* n) { B2P_TRY *n = C(t, sp).n; B2P_CATCH }
int b2p_tile_get_particles(b2p_tile* t, int sp, int alive_only, float* x, float* y, float* z,
                           float* ux, float* uy, float* uz, uint64_t* id, uint64_t* n_out) {
  B2P_TRY
  Container& c = C(t, sp);
  const size_t n = c.n;
  float* dst[6] = { x, y, z, ux, uy, uz };
  const float* src[6] = { c.x.p, c.y.p, c.z.p, c.ux.p, c.uy.p, c.uz.p };
  if (!alive_only) {                                                              // raw container: straight copies
    for (int a = 0; a < 6; ++a) if (dst[a]) d2h(dst[a], src[a], n);
    if (id) d2h(reinterpret_cast<unsigned long long*>(id), c.id.p, n);
    stream_sync();
    if (n_out) *n_out = n;
    return B2P_OK;
  }
  std::vector<unsigned long long> hid(n);
  d2h(hid.data(), c.id.p, n);
  std::vector<float> tmp[6];
  for (int a = 0; a < 6; ++a) if (dst[a]) { tmp[a].resize(n); d2h(tmp[a].data(), src[a], n); }
  stream_sync();
  uint64_t m = 0;
  for (size_t i = 0; i < n; ++i) {                                               // pic/particle.c++:82-168
    if (hid[i] == DEAD) continue;
    for (int a = 0; a < 6; ++a) if (dst[a]) dst[a][m] = tmp[a][i];
    if (id) id[m] = hid[i];
    ++m;
  }
  if (n_out) *n_out = m;
  B2P_CATCH
}
int b2p_tile_push_particles(b2p_tile* t) {
  B2P_TRY_DEFER
  const int pp = T(t)->cfg.particle_pusher;
  if (pp != B2P_PUSHER_BORIS && pp != B2P_PUSHER_HIGUERA_CARY && pp != B2P_PUSHER_FARADAY)
    throw Error(B2P_ERR_LOGIC, "pic::Tile::push_particles: unkown particle pusher");
  defer(DK_PUSH, t);
  B2P_CATCH
}
int b2p_tile_deposit_current(b2p_tile* t) { B2P_TRY_DEFER defer(DK_DEPOSIT, T(t)); B2P_CATCH }
int b2p_tile_sort_particles(b2p_tile* t) { B2P_TRY_DEFER defer(DK_SORT, T(t)); B2P_CATCH }
int b2p_tile_pack_outgoing_particles(b2p_tile* t) { B2P_TRY_DEFER defer(DK_PACK, T(t)); B2P_CATCH }
int b2p_tile_sort_keys(b2p_tile* t, int sp, uint32_t* keys) {
  B2P_TRY
  Container& c = C(t, sp);
  if (c.n) {
    Scratch& s = scratch();
    s.keys[0].reserve(c.n);
    launch_sort_keys(c.view(), t->g, t->origo, s.keys[0].p, nullptr, 0xFFFFFFFFu);
    d2h(keys, s.keys[0].p, c.n);
    stream_sync();
  }
  B2P_CATCH
}
int b2p_tile_get_outgoing(b2p_tile* t, b2p_particle_state* buf, uint64_t cap, uint64_t* ends, uint64_t* n_out) {
  B2P_TRY
  T(t);
  if (n_out) *n_out = t->out_count;
  if (ends) for (size_t i = 0; i < t->out_ends.size(); ++i) ends[i] = t->out_ends[i];
  if (buf) {
    if (cap < t->out_count) throw Error(B2P_ERR_RUNTIME, "outgoing buffer too small");
    d2h(buf, t->out_buf.p, t->out_count);
    stream_sync();
  }
  B2P_CATCH
}
int b2p_tile_kinetic_energy(b2p_tile* t, int sp, double* energy, uint64_t* container_size) {
  B2P_TRY
  Container& c = C(t, sp);
  double h = 0;
  if (c.n) {
    Scratch& s = scratch();
    s.energy.reserve(2);
    B2P_CUDA(cudaMemsetAsync(s.energy.p, 0, sizeof(double), ctx().stream));
    launch_kinetic_energy(c.view(), s.energy.p);
    d2h(&h, s.energy.p, 1);
    stream_sync();
  }
  if (energy) *energy = h;
  if (container_size) *container_size = c.n;
  B2P_CATCH
}

// --------------------------------------------------------------------- grid --
int b2p_tile_register_antenna(b2p_tile* t, const b2p_antenna_mode* m) {      // emf/tile.c++:566-576
  B2P_TRY
  if (!m) throw Error(B2P_ERR_RUNTIME, "null antenna_mode");
  if (m->wave_kind != 0 && m->wave_kind != 1) throw Error(B2P_ERR_RUNTIME, "antenna_mode expects k or n to be defined but not both.");
  b2p_tile::Antenna a;
  for (int d = 0; d < 3; ++d) { a.A[d] = m->A[d]; a.wave[d] = m->wave[d]; }
  a.kind = m->wave_kind;
  a.has_coeffs = m->lap_coeffs != nullptr;
  for (uint64_t q = 0; a.has_coeffs && q < m->n_lap_coeffs; ++q) a.coeffs.push_back({ m->lap_coeffs[2 * q], m->lap_coeffs[2 * q + 1] });
  T(t)->antennas.push_back(std::move(a));
  B2P_CATCH
}
int b2p_tile_deposit_antenna_current(b2p_tile* t) {                            // emf/tile.c++:578-777
  B2P_TRY
  T(t);
  const size_t nm = t->antennas.size();
  // the modes travel as a kernel argument; more than ANTENNA_MAX_MODES are deposited in several sweeps (J += is additive)
  std::vector<AntennaModes> sweeps((nm + ANTENNA_MAX_MODES - 1) / ANTENNA_MAX_MODES + (nm == 0 ? 1 : 0));
  for (AntennaModes& sw : sweeps) sw.n = 0;
  for (size_t n = 0; n < nm; ++n) {
    b2p_tile::Antenna& a = t->antennas[n];
    AntennaModes& sw = sweeps[n / ANTENNA_MAX_MODES];
    const int q = sw.n++;
    for (int d = 0; d < 3; ++d) {
      double k = a.wave[d];
      if (a.kind == 1) {                                                       // :603-614
        const double L = double(t->cfg.n_tiles[d]) * double(t->cfg.n_cells[d]);
        const double tmp = 2 * 3.141592653589793238462643383279502884 * a.wave[d];
        k = tmp / L;
      }
      sw.A[q][d] = static_cast<float>(a.A[d]);
      sw.K[q][d] = static_cast<float>(k);
    }
    if (a.has_coeffs && a.next >= a.coeffs.size())
      throw Error(B2P_ERR_LOGIC, "Can not deposit antenna current, antenna_mode ran out of lap_coeffs!");
    if (a.has_coeffs) { sw.W[q][0] = static_cast<float>(a.coeffs[a.next][0]); sw.W[q][1] = static_cast<float>(a.coeffs[a.next][1]); ++a.next; }
    else { sw.W[q][0] = 1.0f; sw.W[q][1] = 0.0f; }
  }
  Scratch& s = scratch();
  s.antenna[0].reserve(t->lattice_floats());
  s.antenna[1].reserve(t->lattice_floats());
  for (const AntennaModes& sw : sweeps)
    launch_antenna(t->J(), s.antenna[0].p, s.antenna[1].p, t->g, sw, t->mins, t->maxs, static_cast<float>(-t->cfg.cfl));
  B2P_CATCH
}

int b2p_tile_register_edge_bc(b2p_tile* t, const b2p_edge_bc* bc) {
  B2P_TRY
  if (!bc) throw Error(B2P_ERR_RUNTIME, "null edge_bc");
  T(t)->edge_bcs.push_back(*bc);
  B2P_CATCH
}
int b2p_tile_apply_edge_bcs(b2p_tile* t, int mode) { B2P_TRY phase_apply_edge_bcs({ T(t) }, mode); B2P_CATCH }
int b2p_tile_apply_edge_bc(b2p_tile* t, const b2p_edge_bc* bc, int mode) {
  B2P_TRY
  if (!bc) throw Error(B2P_ERR_RUNTIME, "null edge_bc");
  apply_edge_bc(T(t), *bc, mode);
  B2P_CATCH
}
int b2p_tile_register_reflector_wall(b2p_tile* t, const b2p_reflector_wall* wall) {
  B2P_TRY
  if (!wall) throw Error(B2P_ERR_RUNTIME, "null reflector_wall");
  T(t)->walls.push_back(*wall);
  B2P_CATCH
}
int b2p_tile_reflect_particles(b2p_tile* t) { B2P_TRY phase_reflect_particles({ T(t) }); B2P_CATCH }
int b2p_tile_advance_reflector_walls(b2p_tile* t) {                 // pic/reflector_wall.c++:286-297
  B2P_TRY
  for (b2p_reflector_wall& w : T(t)->walls) w.walloc += w.betawall * float(t->cfg.cfl);
  B2P_CATCH
}
int b2p_tile_reflector_walls(b2p_tile* t, b2p_reflector_wall* out, uint64_t cap, uint64_t* n) {
  B2P_TRY
  if (n) *n = T(t)->walls.size();
  for (size_t q = 0; out && q < t->walls.size() && q < cap; ++q) out[q] = t->walls[q];
  B2P_CATCH
}

int b2p_grid_create(const b2p_config* cfg, b2p_grid** out) {
  B2P_TRY
  if (!cfg || !out) throw Error(B2P_ERR_RUNTIME, "null argument");
  validate_config(*cfg);
  ctx();
  b2p_grid* g = new b2p_grid;
  g->cfg = *cfg;
  g->g = make_geom(*cfg);
  const size_t nt = size_t(cfg->n_tiles[0]) * cfg->n_tiles[1] * cfg->n_tiles[2];
  g->slot_of_cid.assign(nt, -1);
  g->owner.assign(nt, 0);
  *out = g;
  B2P_CATCH
}
void b2p_grid_destroy(b2p_grid* g) {
  b2p::flush_quietly();
  delete g;
}
int b2p_grid_add_tile(b2p_grid* g, b2p_tile* t) {
  B2P_TRY
  G(g); T(t);
  for (int d = 0; d < 3; ++d)
    if (t->cfg.n_tiles[d] != g->cfg.n_tiles[d] || t->cfg.n_cells[d] != g->cfg.n_cells[d])
      throw Error(B2P_ERR_RUNTIME, "tile and grid configurations differ");
  // the batched grid phases read these from the grid's first tile
  if (t->cfg.cfl != g->cfg.cfl || t->cfg.field_propagator != g->cfg.field_propagator || t->cfg.current_filter != g->cfg.current_filter ||
      t->cfg.particle_pusher != g->cfg.particle_pusher || std::memcmp(t->cfg.stencil, g->cfg.stencil, sizeof(g->cfg.stencil)) != 0)
    throw Error(B2P_ERR_RUNTIME, "tile and grid configurations differ (cfl / field_propagator / current_filter / particle_pusher / stencil)");
  const int c = g->cid(t->idx[0], t->idx[1], t->idx[2]);
  if (g->slot_of_cid[c] >= 0) throw Error(B2P_ERR_RUNTIME, "tile already added at this index");
  t->grid = g; t->slot = int(g->tiles.size());
  g->ke_valid = false;
  g->slot_of_cid[c] = t->slot;
  g->tiles.push_back(t);
  g->table_dirty = g->nbr_dirty = true;
  B2P_CATCH
}
int b2p_grid_local_communication(b2p_grid* g, int mode) { B2P_TRY grid_local_communication(G(g), mode); B2P_CATCH }
int b2p_grid_push_half_b(b2p_grid* g) { B2P_TRY phase_push_half_b(G(g)->tiles, g->device_table()); B2P_CATCH }
int b2p_grid_push_e(b2p_grid* g) { B2P_TRY phase_push_e(G(g)->tiles, g->device_table(), false); B2P_CATCH }
int b2p_grid_add_current(b2p_grid* g) { B2P_TRY phase_add_current(G(g)->tiles, g->device_table()); B2P_CATCH }
int b2p_grid_filter_current(b2p_grid* g) { B2P_TRY phase_filter(G(g)->tiles); B2P_CATCH }
int b2p_grid_push_particles(b2p_grid* g) { B2P_TRY phase_push_particles(G(g)->tiles); B2P_CATCH }
int b2p_grid_pack_outgoing_particles(b2p_grid* g) { B2P_TRY phase_pack_outgoing(G(g)->tiles); B2P_CATCH }
int b2p_grid_sort_particles(b2p_grid* g) { B2P_TRY phase_sort(G(g)->tiles, false); B2P_CATCH }
int b2p_grid_deposit_current(b2p_grid* g) { B2P_TRY phase_deposit(G(g)->tiles); B2P_CATCH }
int b2p_grid_apply_edge_bcs(b2p_grid* g, int mode) { B2P_TRY phase_apply_edge_bcs(G(g)->tiles, mode); B2P_CATCH }
int b2p_grid_reflect_particles(b2p_grid* g) { B2P_TRY phase_reflect_particles(G(g)->tiles); B2P_CATCH }
int b2p_grid_advance_reflector_walls(b2p_grid* g) {
  B2P_TRY
  for (b2p_tile* t : G(g)->tiles)
    for (b2p_reflector_wall& w : t->walls) w.walloc += w.betawall * float(t->cfg.cfl);
  B2P_CATCH
}

int b2p_grid_external_communication(b2p_grid* g, int mode);

struct HostTrace {
  bool on = std::getenv("B2P_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what, int64_t lap) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (ms > 5.0) std::fprintf(stderr, "[b2p trace] lap %lld %-22s host %.1f ms\n", (long long)lap, what, ms);
    t0 = t1;
  }
};

// mpiio::FieldsWriter<3>::write (io/snapshots/mpiio_fields.c++:221-400; header mpiio_header.h:56-82).
// POSIX pwrite instead of MPI-IO: every rank places its tiles' rows, rank 0 writes the header.
int b2p_grid_write_fields_snapshot(b2p_grid* g, const char* prefix, int32_t lap, int32_t stride, int32_t nspecies) {
  B2P_TRY
  G(g);
  if (!prefix) throw Error(B2P_ERR_RUNTIME, "null snapshot prefix");
  if (stride < 1) throw Error(B2P_ERR_RUNTIME, "snapshot stride must be >= 1");
  const b2p_config& c = g->cfg;
  const int nsp = std::max(0, std::min(nspecies, 5));                       // mpiio_header.h:26 max_species
  const int nf = 9 + nsp;
  const int nxt = std::max(1, c.n_cells[0] / stride), nyt = std::max(1, c.n_cells[1] / stride), nzt = std::max(1, c.n_cells[2] / stride);
  if (nxt * stride > c.n_cells[0] || nyt * stride > c.n_cells[1] || nzt * stride > c.n_cells[2])
    throw Error(B2P_ERR_RUNTIME, "snapshot stride exceeds the tile mesh");
  const int nx = c.n_tiles[0] * nxt, ny = c.n_tiles[1] * nyt, nz = c.n_tiles[2] * nzt;
  const std::string fn = std::string(prefix) + "/flds_" + std::to_string(lap) + ".bin";
  const int fd = ::open(fn.c_str(), O_CREAT | O_WRONLY, 0644);
  if (fd < 0) throw Error(B2P_ERR_RUNTIME, "cannot open " + fn);
  struct Closer { int fd; ~Closer() { ::close(fd); } } closer{ fd };
  const size_t field_elems = size_t(nx) * ny * nz, tile_elems = size_t(nxt) * nyt * nzt;
  const off_t total = off_t(512) + off_t(nf) * off_t(field_elems) * 4;
  if (::ftruncate(fd, total) != 0) throw Error(B2P_ERR_RUNTIME, "cannot size " + fn);   // MPI_File_set_size
  if (g->rank == 0) {
    char hdr[512];
    std::memset(hdr, 0, sizeof hdr);
    auto put = [&](int off, uint32_t v) { std::memcpy(hdr + off, &v, 4); };
    put(0, 0x524E4B4Fu); put(4, 3u); put(8, 512u); put(12, uint32_t(nf));
    put(16, uint32_t(nx)); put(20, uint32_t(ny)); put(24, uint32_t(nz)); put(28, uint32_t(stride));
    put(32, uint32_t(c.n_tiles[0])); put(36, uint32_t(c.n_tiles[1])); put(40, uint32_t(c.n_tiles[2]));
    put(44, uint32_t(c.n_cells[0])); put(48, uint32_t(c.n_cells[1])); put(52, uint32_t(c.n_cells[2]));
    put(56, uint32_t(lap)); put(60, 4u);
    static const char* names[9] = { "ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz" };
    for (int f = 0; f < 9; ++f) std::memcpy(hdr + 64 + f * 16, names[f], std::strlen(names[f]));
    for (int s = 0; s < nsp; ++s) {
      const std::string nm = "n" + std::to_string(s);
      std::memcpy(hdr + 64 + (9 + s) * 16, nm.data(), std::min<size_t>(nm.size(), 15));
    }
    if (::pwrite(fd, hdr, 512, 0) != 512) throw Error(B2P_ERR_RUNTIME, "short write to " + fn);
  }
  DBuf<float> dbuf;
  dbuf.reserve(size_t(nf) * tile_elems);
  std::vector<float> hbuf(size_t(nf) * tile_elems);
  const float inv_stride = 1.0f / float(stride);
  for (b2p_tile* t : g->tiles) {
    launch_pack_snapshot(t->ptrs(), t->g, stride, nxt, nyt, nzt, nf, dbuf.p);
    const float mn[3] = { float(t->mins[0]), float(t->mins[1]), float(t->mins[2]) };
    const int ndep = std::min<int>(int(t->sp.size()), nsp);
    for (int sp = 0; sp < ndep; ++sp)
      launch_snapshot_density(t->sp[sp].view(), mn, inv_stride, nxt, nyt, nzt, dbuf.p + size_t(9 + sp) * tile_elems);
    d2h(hbuf.data(), dbuf.p, hbuf.size());
    stream_sync();
    for (int f = 0; f < nf; ++f)
      for (int ks = 0; ks < nzt; ++ks)
        for (int js = 0; js < nyt; ++js) {
          const off_t off = off_t(512) + 4 * (off_t(f) * off_t(field_elems) +
                                              (off_t(t->idx[2] * nzt + ks) * ny + off_t(t->idx[1] * nyt + js)) * nx + off_t(t->idx[0]) * nxt);
          const size_t bytes = size_t(nxt) * 4;
          if (::pwrite(fd, hbuf.data() + size_t(f) * tile_elems + (size_t(ks) * nyt + js) * nxt, bytes, off) != ssize_t(bytes))
            throw Error(B2P_ERR_RUNTIME, "short write to " + fn);
        }
  }
  B2P_CATCH
}

// projects/pic-turbulence/pic.py:187-221
int b2p_grid_step_pic(b2p_grid* g, int64_t lap) {
  B2P_TRY
  G(g);
  HostTrace tr;
  const bool multi = g->nranks > 1;
  auto ext = [&](int mode) {
    if (multi) { const int rc = b2p_grid_external_communication(g, mode); if (rc) throw Error(rc, g_last_error); }
  };
  phase_push_half_b(g->tiles, g->device_table());
  if (multi && tuning().comm_overlap) {
    // The B halo exchange flies on the plan's own stream while the tiles that have no remote neighbour are pushed:
    // halo cells fed by local tiles are filled at once, those fed by remote ranks when the exchange has landed,
    // right before the boundary tiles' pushes.
    comm_exchange_fields_on_comm_stream(g, B2P_COMM_EMF_B);
    grid_local_communication(g, B2P_COMM_EMF_B, /*part=*/1);
    std::vector<b2p_tile*> order;
    size_t n_interior = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (b2p_tile* t : g->tiles) {
        bool remote = false;
        for (int di = 0; di < 27 && !remote; ++di) { int e = 0; remote = comm_remote_entry(g, t->slot, di, &e); }
        if (remote == (pass == 1)) order.push_back(t);
        if (pass == 0 && !remote) ++n_interior;
      }
    tr.mark("fields", lap);
    phase_push_particles(order, n_interior, [&] { comm_wait_exchange(g); comm_unpack_halo(g, /*B*/ 1); });
  } else {
    ext(B2P_COMM_EMF_B); grid_local_communication(g, B2P_COMM_EMF_B);
    tr.mark("fields", lap);
    phase_push_particles(g->tiles);         // also publishes the leaver masks pack_outgoing consumes
  }
  tr.mark("push enqueue", lap);
  phase_pack_outgoing(g->tiles);
  tr.mark("pack", lap);
  ext(B2P_COMM_PIC_PARTICLE); grid_local_communication(g, B2P_COMM_PIC_PARTICLE);
  tr.mark("particle comm", lap);
  if (lap % 5 == 0) phase_sort(g->tiles, /*leave_running=*/true);   // joined by the next entry point (or a fresh deposit)
  tr.mark("sort enqueue", lap);
  phase_deposit(g->tiles);
  tr.mark("deposit enqueue", lap);
  if (multi && tuning().comm_overlap >= 2 && g->n_boundary > 0) {
    // Every exchange of the field phase flies on the plan's own stream under work that does not need it.  The grid's
    // tiles are ordered boundary-first (comm.cu: order_boundary_tiles_first), so a phase is: the boundary tiles, whose
    // results the exchange sends; the exchange; the same phase on the interior tiles [nb, nt) and the fill of the
    // locally fed halo cells meanwhile; then the remote slabs go into the boundary tiles' halos (comm_unpack_halo).
    // Which tile is updated first does not change any operand: same bits as the sequential lap.
    const size_t nb = g->n_boundary, nt = g->tiles.size();
    const std::vector<b2p_tile*> bnd(g->tiles.begin(), g->tiles.begin() + nb), inte(g->tiles.begin() + nb, g->tiles.end());
    auto round = [&](int mode, int which, const std::function<void()>& before_fill, const std::function<void()>& after_fill) {
      comm_exchange_fields_on_comm_stream(g, mode);
      if (before_fill) before_fill();                       // interior work the local fill reads
      grid_local_communication(g, mode, /*part=*/1);
      if (after_fill) after_fill();                         // interior work that needs only locally fed halos
      comm_wait_exchange(g);
      comm_unpack_halo(g, which);
    };
    comm_exchange_fields_on_comm_stream(g, B2P_COMM_EMF_J);   // the halos the J exchange adds (and the edges, unused here)
    grid_J_exchange(g, nb, nt - nb);
    comm_wait_exchange(g);
    grid_J_exchange(g, 0, nb);
    if (g->cfg.current_filter >= 0) {
      round(B2P_COMM_EMF_J, 2, nullptr, [&] { phase_filter(inte); });
      phase_filter(bnd);
      round(B2P_COMM_EMF_J, 2, nullptr, [&] { phase_filter(inte); });
      phase_filter(bnd);
      phase_filter(g->tiles);
    } else {
      round(B2P_COMM_EMF_J, 2, nullptr, nullptr);
    }
    phase_push_half_b(bnd, g->device_table());
    round(B2P_COMM_EMF_B, 1, [&] { phase_push_half_b(inte, g->device_table() + nb); }, nullptr);
    phase_push_e(bnd, g->device_table(), true);
    round(B2P_COMM_EMF_E, 0, [&] { phase_push_e(inte, g->device_table() + nb, true); }, nullptr);
    tr.mark("J comm+filter+fields", lap);
    return B2P_OK;
  }
  ext(B2P_COMM_EMF_J); grid_local_communication(g, B2P_COMM_EMF_J_EXCHANGE);
  ext(B2P_COMM_EMF_J); grid_local_communication(g, B2P_COMM_EMF_J);
  if (g->cfg.current_filter >= 0) {
    phase_filter(g->tiles);
    ext(B2P_COMM_EMF_J); grid_local_communication(g, B2P_COMM_EMF_J);
    phase_filter(g->tiles);
    phase_filter(g->tiles);
  }
  phase_push_half_b(g->tiles, g->device_table());
  ext(B2P_COMM_EMF_B); grid_local_communication(g, B2P_COMM_EMF_B);
  phase_push_e(g->tiles, g->device_table(), true);   // push_e + add_current fused (same roundings)
  ext(B2P_COMM_EMF_E); grid_local_communication(g, B2P_COMM_EMF_E);
  tr.mark("J comm+filter+fields", lap);
  B2P_CATCH
}

// projects/emf-wave/emf.py:48-62
int b2p_grid_step_emf(b2p_grid* g) {
  B2P_TRY
  G(g);
  const bool multi = g->nranks > 1;
  auto ext = [&](int mode) {
    if (multi) { const int rc = b2p_grid_external_communication(g, mode); if (rc) throw Error(rc, g_last_error); }
  };
  ext(B2P_COMM_EMF_E); grid_local_communication(g, B2P_COMM_EMF_E);
  phase_push_half_b(g->tiles, g->device_table(), 2);
  ext(B2P_COMM_EMF_B); grid_local_communication(g, B2P_COMM_EMF_B);
  phase_push_e(g->tiles, g->device_table(), false);
  B2P_CATCH
}

// one k_kinetic_energy_batch launch per 256 containers of the grid: per-species sums into energy[q] / alive[q]
static void energy_jobs(b2p_grid* g, double* energy, unsigned long long* alive) {
  const int ns = g->cfg.n_species;
  std::vector<EnergyJob> jobs;
  unsigned max_n = 0;
  double slots = 0;
  auto flush = [&] {
    if (jobs.empty()) return;
    Scratch& s = scratch();
    s.table.reserve(jobs.size() * sizeof(EnergyJob));
    h2d(reinterpret_cast<EnergyJob*>(s.table.p), jobs.data(), jobs.size());
    launch_kinetic_energy_batch(reinterpret_cast<const EnergyJob*>(s.table.p), int(jobs.size()), max_n, slots);
    jobs.clear(); max_n = 0; slots = 0;      // (a pageable source is staged before cudaMemcpyAsync returns)
  };
  for (b2p_tile* t : g->tiles)
    for (int q = 0; q < ns; ++q) {
      Container& c = t->sp[q];
      if (!c.n) continue;
      jobs.push_back(EnergyJob{ c.view(), energy ? energy + q : nullptr, alive ? alive + q : nullptr });
      max_n = std::max(max_n, c.n); slots += c.n;
      if (jobs.size() == 4096) flush();
    }
  flush();
}

int b2p_grid_energies(b2p_grid* g, double* eB, double* eE, double* kinetic, uint64_t* sizes) {
  B2P_TRY
  G(g);
  const int nt = int(g->tiles.size());
  const int ns = g->cfg.n_species;
  Scratch& s = scratch();
  s.energy.reserve(size_t(2) * std::max(nt, 1) + ns + 1);
  double* dk = s.energy.p + 2 * size_t(std::max(nt, 1));
  launch_field_energy(g->device_table(), nt, g->g, s.energy.p);
  B2P_CUDA(cudaMemsetAsync(dk, 0, sizeof(double) * (ns + 1), ctx().stream));
  std::vector<uint64_t> hs(ns, 0);
  const bool account = g->ke_valid && tuning().energy_cache && ns > 0;
  if (kinetic) g->ke_asked = true;
  if (!account) energy_jobs(g, dk, nullptr);
  for (b2p_tile* t : g->tiles)
    for (int q = 0; q < ns; ++q) hs[q] += t->sp[q].n;
  std::vector<double> h(size_t(2) * nt + ns);
  std::vector<double> acc(account ? size_t(ns) * KE_SLOTS : 0);
  d2h(h.data(), s.energy.p, size_t(2) * nt);
  if (account) d2h(acc.data(), g->ke_acc.p, acc.size());       // the running account: 64 KB per species
  else d2h(h.data() + 2 * size_t(nt), dk, ns);
  stream_sync();
  if (account)
    for (int q = 0; q < ns; ++q) {
      double t = 0;
      for (int k = 0; k < KE_SLOTS; ++k) t += acc[size_t(q) * KE_SLOTS + k];
      h[2 * size_t(nt) + q] = t;
    }
  double b = 0, e = 0;
  for (int i = 0; i < nt; ++i) { b += h[2 * i] / 2.0; e += h[2 * i + 1] / 2.0; }
  if (eB) *eB = b;
  if (eE) *eE = e;
  for (int q = 0; q < ns; ++q) { if (kinetic) kinetic[q] = h[2 * size_t(nt) + q]; if (sizes) sizes[q] = hs[q]; }
  B2P_CATCH
}

int b2p_grid_alive_counts(b2p_grid* g, uint64_t* counts) {
  B2P_TRY
  G(g);
  const int ns = g->cfg.n_species;
  Scratch& s = scratch();
  s.energy.reserve(size_t(ns) + 1);
  unsigned long long* dc = reinterpret_cast<unsigned long long*>(s.energy.p);
  B2P_CUDA(cudaMemsetAsync(dc, 0, sizeof(unsigned long long) * ns, ctx().stream));
  energy_jobs(g, nullptr, dc);
  std::vector<unsigned long long> h(ns);
  d2h(h.data(), dc, size_t(ns));
  stream_sync();
  for (int q = 0; q < ns; ++q) counts[q] = h[q];
  B2P_CATCH
}

int b2p_grid_inject_thermal(b2p_grid* g, int ppc, double delgam, uint64_t seed) {
  B2P_TRY
  G(g);
  if (ppc < 0) throw Error(B2P_ERR_RUNTIME, "ppc must be non-negative");
  g->ke_valid = false;
  for (b2p_tile* t : g->tiles) {
    const size_t total = size_t(t->g.N[0]) * t->g.N[1] * t->g.N[2] * size_t(ppc);
    if (total >= (size_t(1) << 32)) throw Error(B2P_ERR_RUNTIME, "particle container exceeds uint32 indexing");
    const float mn[3] = { float(t->mins[0]), float(t->mins[1]), float(t->mins[2]) };
    t->pendJ_valid = false;
    for (size_t q = 0; q < t->sp.size(); ++q) {
      Container& c = t->sp[q];
      if (c.n) throw Error(B2P_ERR_RUNTIME, "inject_thermal requires empty containers");
      c.reserve(total);
      c.n = unsigned(total);
      const unsigned long long sp_seed = seed * 0x9E3779B97F4A7C15ull + t->tile_tag * 0xC2B2AE3D27D4EB4Full;
      launch_inject_thermal(c.view(), t->g, mn, unsigned(ppc), float(delgam), sp_seed, sp_seed ^ (0xA5A5A5A5ull * (q + 1)),
                            (t->tile_tag << 40) | t->next_ordinal[q]);
      t->next_ordinal[q] += total;
      c.P = c.n; c.P_valid = true; c.masks_valid = false;
    }
  }
  B2P_CATCH
}

int b2p_grid_inject_drifting_stripe(b2p_grid* g, int sp, int ppc, double delgam, double gamma_drift, int dir_sign, double x_left,
                                    double x_right, uint64_t seed) {
  B2P_TRY
  G(g);
  if (ppc < 0 || gamma_drift < 1.0 || (dir_sign != 1 && dir_sign != -1)) throw Error(B2P_ERR_RUNTIME, "inject_drifting_stripe: bad arguments");
  g->ke_valid = false;
  for (b2p_tile* t : g->tiles) {
    Container& c = C(t, sp);
    // the cell range of pic::Tile::batch_inject_in_x_stripe (pic/tile.c++:235-262)
    const double xmin = t->mins[0], xmax = t->maxs[0];
    if (x_right <= xmin || x_left >= xmax) continue;
    const int nx = t->g.N[0];
    const double dl = x_left - xmin, dr = x_right - xmin;
    const int i0 = dl <= 0.0 ? 0 : std::min(nx, int(std::floor(dl)));
    const int i1 = dr >= double(nx) ? nx : int(std::ceil(dr));
    if (i0 >= i1) continue;
    const size_t total = size_t(i1 - i0) * t->g.N[1] * t->g.N[2] * size_t(ppc);
    const unsigned P = find_P(c);
    if (size_t(P) + total >= (size_t(1) << 32)) throw Error(B2P_ERR_RUNTIME, "particle container exceeds uint32 indexing");
    c.reserve(size_t(P) + total);
    if (size_t(P) + total > c.n) c.n = unsigned(P + total);      // pre-allocated containers take the batch in their dead slots
    t->pendJ_valid = false;
    const float mn[3] = { float(t->mins[0]), float(t->mins[1]), float(t->mins[2]) };
    const unsigned long long pos_seed = seed * 0x9E3779B97F4A7C15ull + t->tile_tag * 0xC2B2AE3D27D4EB4Full;   // the same positions for every species
    launch_inject_drifting(c.view(), P, t->g, mn, i0, i1, unsigned(ppc), float(delgam), float(gamma_drift), float(dir_sign), pos_seed,
                           pos_seed ^ (0xA5A5A5A5ull * (sp + 1)) ^ (dir_sign > 0 ? 0x5151ull : 0ull), (t->tile_tag << 40) | t->next_ordinal[sp]);
    t->next_ordinal[sp] += total;
    c.P = unsigned(P + total); c.P_valid = true; c.masks_valid = false;
  }
  B2P_CATCH
}

int b2p_grid_set_uniform_B(b2p_grid* g, float bx, float by, float bz) {
  B2P_TRY
  G(g);
  for (b2p_tile* t : g->tiles) {
    std::vector<float> h(t->lattice_floats());
    const float v[3] = { bx, by, bz };
    for (int c = 0; c < 3; ++c) std::fill(h.begin() + size_t(c) * t->g.Ch, h.begin() + size_t(c + 1) * t->g.Ch, v[c]);
    h2d(t->B.p, h.data(), h.size());
    stream_sync();
  }
  B2P_CATCH
}

int b2p_timer_start(void) { B2P_TRY B2P_CUDA(cudaEventRecord(ctx().ev0, ctx().stream)); B2P_CATCH }
int b2p_timer_stop(float* ms) {
  B2P_TRY
  B2P_CUDA(cudaEventRecord(ctx().ev1, ctx().stream));
  B2P_CUDA(cudaEventSynchronize(ctx().ev1));
  B2P_CUDA(cudaEventElapsedTime(ms, ctx().ev0, ctx().ev1));
  B2P_CATCH
}
uint64_t b2p_launch_count(void) {
  b2p::flush_quietly();
  return g_ctx_ready ? ctx().launches : 0;
}
void b2p_copy_bytes(uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (h2d_bytes) *h2d_bytes = g_ctx_ready ? ctx().h2d_bytes : 0;
  if (d2h_bytes) *d2h_bytes = g_ctx_ready ? ctx().d2h_bytes : 0;
}
void b2p_host_wait_ms(double* ms) {
  if (ms) *ms = g_ctx_ready ? ctx().host_wait_ms : 0.0;
}
int b2p_profile_enable(int on) {
  B2P_TRY
  stream_sync();
  for (auto& r : b2p::g_prof) { b2p::g_event_pool.push_back(r.a); b2p::g_event_pool.push_back(r.b); }
  b2p::g_prof.clear();
  b2p::g_prof_on = on != 0;
  B2P_CATCH
}
int b2p_profile_num_classes(void) { return b2p::KC_COUNT; }
const char* b2p_profile_class_name(int k) { return b2p::kernel_class_name(k); }
int b2p_selfcheck_const_division(const float* x, uint64_t n, float c, float* out, float* ref) {
  B2P_TRY
  if (!x || !out || !ref) throw Error(B2P_ERR_RUNTIME, "null argument");
  DBuf<float> dx, dout, dref;
  dx.reserve_exact(n + 1); dout.reserve_exact(n + 1); dref.reserve_exact(n + 1);
  h2d(dx.p, x, n);
  launch_selfcheck_divc(dx.p, n, c, dout.p, dref.p);
  d2h(out, dout.p, n); d2h(ref, dref.p, n);
  stream_sync();
  B2P_CATCH
}
int b2p_profile_report(double* ms, uint64_t* launches, double* units) {
  B2P_TRY
  stream_sync();
  for (int k = 0; k < b2p::KC_COUNT; ++k) { ms[k] = 0; launches[k] = 0; units[k] = 0; }
  for (auto& r : b2p::g_prof) {
    float t = 0;
    B2P_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.kc] += t; launches[r.kc] += 1; units[r.kc] += r.units;
  }
  B2P_CATCH
}

}  // extern "C"
