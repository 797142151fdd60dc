// comm.cu — multi-GPU exchange over NCCL point-to-point (NVLink 5 / NVSwitch).
//
// Replaces corgi's MPI transport (external/corgi/src/corgi/corgi.h:1560-1692:
// Tile::send_data / VirtualTile::recv_data of whole 3- or 6-thick hollow shells per
// neighbouring rank) with packed per-(tile, direction) slabs: one grouped
// ncclSend/ncclRecv pair per peer rank and mode on the library stream.
//
//   field modes  : for a local tile T whose Moore neighbour O in direction d lives on
//                  another rank, O's corresponding_subregion(d) (interior edge, the
//                  source of T's halo fill) and, for emf_J, also O's subregion(-d)
//                  (O's halo, the source of emf_J_exchange) are received into a staging
//                  buffer; grid.local_communication(mode) then reads them exactly where
//                  the reference reads the VirtualTile's hollow grids
//                  (emf/tile.c++:510-534), in the same Moore order.
//   pic_particle : the number_of_particles handshake (pic/tile_communication.c++:47-52)
//                  followed by only the spans the receiver will read
//                  (pic/tile_communication.c++:133-181), 32-byte ParticleState AoS.
//
// Both sides enumerate the (receiver tile cid, receiver direction) pairs of a rank pair
// in the same sorted order, so no tags are needed.  NCCL is dlopen()ed on first use;
// single-rank grids never touch it.
#include "host.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <map>

namespace b2p {

// ---------------------------------------------------------------- NCCL binding --
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl& nccl() {
  static Nccl n;
  if (n.h) return n;
  const char* names[] = { "libnccl.so.2", "libnccl.so" };
  for (const char* nm : names) { n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.h) break; }
  if (!n.h) throw Error(B2P_ERR_RUNTIME, std::string("cannot load NCCL: ") + dlerror());
#define LOAD(field, sym)                                                                   \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, sym));                          \
  if (!n.field) throw Error(B2P_ERR_RUNTIME, std::string("NCCL symbol missing: ") + sym);
  LOAD(GetUniqueId, "ncclGetUniqueId") LOAD(CommInitRank, "ncclCommInitRank") LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(Send, "ncclSend") LOAD(Recv, "ncclRecv") LOAD(GroupStart, "ncclGroupStart") LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  return n;
}
#define B2P_NCCL(expr)                                                                               \
  do {                                                                                               \
    ncclResult_t r__ = (expr);                                                                       \
    if (r__ != ncclSuccess)                                                                          \
      throw Error(B2P_ERR_RUNTIME, std::string(#expr) + ": " + nccl().GetErrorString(r__));          \
  } while (0)

// ----------------------------------------------------------------- the plan --
// One (tile, direction) pair whose Moore neighbour is owned by another rank.
struct PlanEntry {
  int peer;                 // the other rank
  int cid;                  // my tile
  int dir[3];               // direction from my tile to the remote neighbour
  int remote_cid;
  unsigned long long recv_key, send_key;   // (receiver cid << 5) | receiver direction index
  int dims[3];              // slab extents: 3 along axes with dir != 0, N otherwise
  size_t volume() const { return size_t(dims[0]) * dims[1] * dims[2]; }
};

static int dir_index(const int d[3]) { return ((d[0] + 1) * 3 + (d[1] + 1)) * 3 + (d[2] + 1); }
static int wrapc(int v, int n) { while (v < 0) v += n; while (v >= n) v -= n; return v; }

// Pure host logic (also exported for the CPU tests): all (tile, dir) pairs of `rank`
// with a remote neighbour, in Moore order per tile.
std::vector<PlanEntry> build_plan(const b2p_config& cfg, const std::vector<int>& owner, int rank) {
  std::vector<PlanEntry> out;
  const int* T = cfg.n_tiles;
  for (int k = 0; k < T[2]; ++k) for (int j = 0; j < T[1]; ++j) for (int i = 0; i < T[0]; ++i) {
    const int cid = i + T[0] * (j + T[1] * k);
    if (owner[cid] != rank) continue;
    for (int kr = -1; kr <= 1; ++kr) for (int jr = -1; jr <= 1; ++jr) for (int ir = -1; ir <= 1; ++ir) {
      if (!ir && !jr && !kr) continue;
      const int oc = wrapc(i + ir, T[0]) + T[0] * (wrapc(j + jr, T[1]) + T[1] * wrapc(k + kr, T[2]));
      if (owner[oc] == rank) continue;
      PlanEntry e;
      e.peer = owner[oc]; e.cid = cid; e.remote_cid = oc;
      e.dir[0] = ir; e.dir[1] = jr; e.dir[2] = kr;
      const int inv[3] = { -ir, -jr, -kr };
      e.recv_key = (static_cast<unsigned long long>(cid) << 5) | unsigned(dir_index(e.dir));     // I receive as (me, dir)
      e.send_key = (static_cast<unsigned long long>(oc) << 5) | unsigned(dir_index(inv));        // the peer receives as (it, -dir)
      for (int d = 0; d < 3; ++d) e.dims[d] = e.dir[d] ? H : cfg.n_cells[d];
      out.push_back(e);
    }
  }
  return out;
}

struct PeerBuffers {
  int peer = -1;
  std::vector<int> send_order, recv_order;      // indices into CommPlan::entries, sorted by key
  size_t send_floats[2] = { 0, 0 }, recv_floats[2] = { 0, 0 };   // per kind (0 = interior edge, 1 = halo)
  size_t send_off = 0, recv_off = 0;            // offsets of this peer's block in the staging buffers
};

struct CommPlan {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  std::vector<PlanEntry> entries;               // Moore order per tile (this rank's remote pairs)
  std::map<std::pair<int, int>, int> entry_of;  // (tile slot, dir index) -> entry
  std::vector<PeerBuffers> peers;
  DBuf<float> sendbuf, recvbuf;
  DBuf<SlabDesc> d_remote_fill, d_remote_exch;
  DBuf<int> d_slot_of_entry;                    // entry -> local tile slot (k_unpack_halo)
  // Pack tables are persistent: one per (mode, which of a tile's two J buffers is current), rebuilt only when the
  // lattice pointers behind it change (the exchanges of a lap reuse four tables that are uploaded once).
  struct PackTable { std::vector<const float*> sig; DBuf<SlabDesc> d; size_t n = 0; };
  PackTable pack[4];
  std::vector<size_t> recv_slab_off[2];         // per entry, per kind: float offset in recvbuf
  // particles
  DBuf<b2p_particle_state> psend, precv;
  DBuf<unsigned> d_cnt_send, d_cnt_recv;
  std::vector<std::vector<std::pair<size_t, unsigned>>> pspan;   // [entry][species] -> (offset in precv, count)
  bool pspan_valid = false;
  cudaStream_t cstream = nullptr;               // the plan's own stream for exchanges that overlap compute
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  ~CommPlan() {
    if (comm) nccl().CommDestroy(comm);
    if (cstream) cudaStreamDestroy(cstream);
    if (ev_ready) cudaEventDestroy(ev_ready);
    if (ev_done) cudaEventDestroy(ev_done);
  }
};

// ----------------------------------------------------------------- kernels --
// gather a lattice region into a packed comp-major slab
__global__ void __launch_bounds__(256)
k_pack_slabs(const SlabDesc* __restrict__ slabs, const Geom g) {
  const SlabDesc s = slabs[blockIdx.y];
  B2P_GLOBAL(s.base); B2P_GLOBAL(s.field);
  // 32-bit index arithmetic: a slab is at most one face of a tile lattice (checked when the plan is built)
  const unsigned d1 = unsigned(s.dims[1]), d2 = unsigned(s.dims[2]), vol = unsigned(s.dims[0]) * d1 * d2;
  const unsigned Hy = unsigned(g.Hx[1]), Hz = unsigned(g.Hx[2]);
  const unsigned n0 = (unsigned(s.begin[0]) * Hy + unsigned(s.begin[1])) * Hz + unsigned(s.begin[2]);
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < 3u * vol; q += gridDim.x * blockDim.x) {
    const unsigned c = q / vol, r = q - c * vol;
    const unsigned t = r / d2, kk = r - t * d2, ii = t / d1, jj = t - ii * d1;
    s.base[q] = s.field[size_t(c) * g.Ch + n0 + (ii * Hy + jj) * Hz + kk];
  }
}

// the inverse for the halo fill: every staged slab goes to the halo region of its tile that faces the remote neighbour
// (what k_halo_fill does for these cells, driven by the slabs instead of by a sweep over every halo cell of every tile)
__global__ void __launch_bounds__(256)
k_unpack_halo(const FieldPtrs* __restrict__ tiles, const SlabDesc* __restrict__ slabs, const int* __restrict__ slot_of_entry,
              const Geom g, const int which) {
  const SlabDesc s = slabs[blockIdx.y];
  const FieldPtrs f = tiles[slot_of_entry[blockIdx.y]];
  float* __restrict__ dst = which == 0 ? f.E : (which == 1 ? f.B : f.J);
  B2P_GLOBAL(s.base); B2P_GLOBAL(dst);
  const unsigned d1 = unsigned(s.dims[1]), d2 = unsigned(s.dims[2]), vol = unsigned(s.dims[0]) * d1 * d2;
  const unsigned Hy = unsigned(g.Hx[1]), Hz = unsigned(g.Hx[2]);
  const unsigned n0 = (unsigned(s.begin[0]) * Hy + unsigned(s.begin[1])) * Hz + unsigned(s.begin[2]);
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < 3u * vol; q += gridDim.x * blockDim.x) {
    const unsigned c = q / vol, r = q - c * vol;
    const unsigned t = r / d2, kk = r - t * d2, ii = t / d1, jj = t - ii * d1;
    dst[size_t(c) * g.Ch + n0 + (ii * Hy + jj) * Hz + kk] = s.base[q];
  }
}

struct CopyJob { const b2p_particle_state* src; b2p_particle_state* dst; unsigned count; };
__global__ void __launch_bounds__(256)
k_copy_spans(const CopyJob* __restrict__ jobs) {
  const CopyJob jb = jobs[blockIdx.y];
  const uint4* s = reinterpret_cast<const uint4*>(jb.src);
  uint4* d = reinterpret_cast<uint4*>(jb.dst);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * jb.count; i += gridDim.x * blockDim.x) d[i] = s[i];
}

}  // namespace b2p

using namespace b2p;

b2p_grid::~b2p_grid() {
  for (b2p_tile* t : tiles) { t->grid = nullptr; t->slot = -1; }
  delete comm;
}

namespace b2p {

static void sync_stream() { timed_stream_sync(); }

// region of tile lattice that is SENT for entry e: kind 0 = corresponding_subregion(-dir)
// (my interior edge facing the peer), kind 1 = subregion(dir) (my halo facing the peer)
static void send_region_begin(const PlanEntry& e, const b2p_config& cfg, int kind, int begin[3]) {
  for (int d = 0; d < 3; ++d) {
    const int N = cfg.n_cells[d];
    if (kind == 0) begin[d] = e.dir[d] == 0 ? H : (e.dir[d] == 1 ? N : H);          // corr(-dir): dir=+1 -> [N,N+3), dir=-1 -> [3,6)
    else begin[d] = e.dir[d] == 0 ? H : (e.dir[d] == 1 ? H + N : 0);                // subregion(dir)
  }
}

// Tiles with a Moore neighbour on another rank first: a phase can then run on the boundary tiles, start their exchange,
// and run on the interior tiles [n_boundary, n) of the same device table while the slabs travel.
static void order_boundary_tiles_first(b2p_grid* g, int rank) {
  const int* T = g->cfg.n_tiles;
  auto is_boundary = [&](const b2p_tile* t) {
    for (int kr = -1; kr <= 1; ++kr) for (int jr = -1; jr <= 1; ++jr) for (int ir = -1; ir <= 1; ++ir) {
      const int oc = wrapc(t->idx[0] + ir, T[0]) + T[0] * (wrapc(t->idx[1] + jr, T[1]) + T[1] * wrapc(t->idx[2] + kr, T[2]));
      if (g->owner[oc] != rank) return true;
    }
    return false;
  };
  const auto mid = std::stable_partition(g->tiles.begin(), g->tiles.end(), is_boundary);
  g->n_boundary = size_t(mid - g->tiles.begin());
  std::fill(g->slot_of_cid.begin(), g->slot_of_cid.end(), -1);
  for (size_t i = 0; i < g->tiles.size(); ++i) {
    g->tiles[i]->slot = int(i);
    g->slot_of_cid[g->cid(g->tiles[i]->idx[0], g->tiles[i]->idx[1], g->tiles[i]->idx[2])] = int(i);
  }
  g->table_dirty = g->nbr_dirty = true;
}

static void finalize_plan(b2p_grid* g) {
  CommPlan& p = *g->comm;
  order_boundary_tiles_first(g, p.rank);
  p.entries = build_plan(g->cfg, g->owner, p.rank);
  p.entry_of.clear();
  std::map<int, PeerBuffers> byp;
  for (size_t i = 0; i < p.entries.size(); ++i) {
    const PlanEntry& e = p.entries[i];
    const int slot = g->slot_of_cid[e.cid];
    if (slot < 0) throw Error(B2P_ERR_RUNTIME, "comm_init: a tile owned by this rank has not been added to the grid");
    p.entry_of[{ slot, dir_index(e.dir) }] = int(i);
    byp[e.peer].peer = e.peer;
    byp[e.peer].send_order.push_back(int(i));
    byp[e.peer].recv_order.push_back(int(i));
  }
  p.peers.clear();
  size_t soff = 0, roff = 0;
  p.recv_slab_off[0].assign(p.entries.size(), 0);
  p.recv_slab_off[1].assign(p.entries.size(), 0);
  for (auto& kv : byp) {
    PeerBuffers pb = kv.second;
    std::sort(pb.send_order.begin(), pb.send_order.end(), [&](int a, int b) { return p.entries[a].send_key < p.entries[b].send_key; });
    std::sort(pb.recv_order.begin(), pb.recv_order.end(), [&](int a, int b) { return p.entries[a].recv_key < p.entries[b].recv_key; });
    pb.send_off = soff; pb.recv_off = roff;
    for (int kind = 0; kind < 2; ++kind) {
      for (int i : pb.send_order) pb.send_floats[kind] += 3 * p.entries[i].volume();
      size_t o = roff + (kind ? pb.recv_floats[0] : 0);
      for (int i : pb.recv_order) { p.recv_slab_off[kind][i] = o; o += 3 * p.entries[i].volume(); pb.recv_floats[kind] += 3 * p.entries[i].volume(); }
    }
    soff += pb.send_floats[0] + pb.send_floats[1];
    roff += pb.recv_floats[0] + pb.recv_floats[1];
    p.peers.push_back(pb);
  }
  p.sendbuf.reserve(std::max<size_t>(soff, 1));
  p.recvbuf.reserve(std::max<size_t>(roff, 1));
  // device tables of the staged slabs, indexed by entry
  std::vector<SlabDesc> rf(p.entries.size()), rx(p.entries.size());
  for (size_t i = 0; i < p.entries.size(); ++i) {
    for (int kind = 0; kind < 2; ++kind) {
      SlabDesc& s = kind ? rx[i] : rf[i];
      s.base = p.recvbuf.p + p.recv_slab_off[kind][i];
      s.field = nullptr;
      for (int d = 0; d < 3; ++d) {
        // where the slab lands in the receiving lattice (k_unpack_halo): subregion(dir), the halo facing the neighbour
        const int dr = p.entries[i].dir[d];
        s.begin[d] = dr == 0 ? H : (dr == 1 ? H + g->cfg.n_cells[d] : 0);
        s.dims[d] = p.entries[i].dims[d];
      }
    }
  }
  std::vector<int> slots(p.entries.size());
  for (size_t i = 0; i < p.entries.size(); ++i) slots[i] = g->slot_of_cid[p.entries[i].cid];
  p.d_slot_of_entry.reserve(std::max<size_t>(slots.size(), 1));
  if (!slots.empty())
    B2P_CUDA(cudaMemcpyAsync(p.d_slot_of_entry.p, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, ctx().stream));
  p.d_remote_fill.reserve(std::max<size_t>(rf.size(), 1));
  p.d_remote_exch.reserve(std::max<size_t>(rx.size(), 1));
  if (!rf.empty()) {
    B2P_CUDA(cudaMemcpyAsync(p.d_remote_fill.p, rf.data(), rf.size() * sizeof(SlabDesc), cudaMemcpyHostToDevice, ctx().stream));
    B2P_CUDA(cudaMemcpyAsync(p.d_remote_exch.p, rx.data(), rx.size() * sizeof(SlabDesc), cudaMemcpyHostToDevice, ctx().stream));
  }
  sync_stream();
  g->nbr_dirty = true;
}

// exchange the field slabs of one mode
static void exchange_fields(b2p_grid* g, int mode) {
  CommPlan& p = *g->comm;
  const int nk = mode == B2P_COMM_EMF_J ? 2 : 1;
  // which persistent table: E, B, J with the first local tile's jcur (all tiles of a grid flip together)
  const int which = mode == B2P_COMM_EMF_E ? 0 : (mode == B2P_COMM_EMF_B ? 1 : 2 + (g->tiles.empty() ? 0 : g->tiles[0]->jcur));
  CommPlan::PackTable& pt = p.pack[which];
  std::vector<const float*> sig;
  for (const PeerBuffers& pb : p.peers)
    for (int i : pb.send_order) {
      b2p_tile* t = g->tiles[g->slot_of_cid[p.entries[i].cid]];
      sig.push_back(mode == B2P_COMM_EMF_E ? t->E.p : (mode == B2P_COMM_EMF_B ? t->B.p : t->J()));
    }
  if (sig != pt.sig || pt.n == 0) {
    std::vector<SlabDesc> pack;
    for (const PeerBuffers& pb : p.peers) {
      size_t o = pb.send_off;
      for (int kind = 0; kind < nk; ++kind)
        for (int i : pb.send_order) {
          const PlanEntry& e = p.entries[i];
          b2p_tile* t = g->tiles[g->slot_of_cid[e.cid]];
          SlabDesc s;
          s.base = p.sendbuf.p + o;
          s.field = mode == B2P_COMM_EMF_E ? t->E.p : (mode == B2P_COMM_EMF_B ? t->B.p : t->J());
          send_region_begin(e, g->cfg, kind, s.begin);
          for (int d = 0; d < 3; ++d) s.dims[d] = e.dims[d];
          pack.push_back(s);
          o += 3 * e.volume();
        }
    }
    pt.n = pack.size();
    pt.sig.swap(sig);
    if (pt.n) {
      pt.d.reserve(pt.n);
      B2P_CUDA(cudaMemcpyAsync(pt.d.p, pack.data(), pt.n * sizeof(SlabDesc), cudaMemcpyHostToDevice, ctx().stream));
    }
  }
  for (size_t b = 0; b < pt.n; b += 65535) {
    ProfScope prof_(KC_HALO, 0.0);
    const unsigned nb = unsigned(std::min<size_t>(65535, pt.n - b));
    k_pack_slabs<<<dim3(24, nb), 256, 0, ctx().stream>>>(pt.d.p + b, g->g);
    B2P_LAUNCH_CHECK();
  }
  Nccl& n = nccl();
  ProfScope prof_nccl_(KC_NCCL, 0.0);     // send/recv kernels + the wait for the slowest peer
  B2P_NCCL(n.GroupStart());
  for (const PeerBuffers& pb : p.peers) {
    size_t ns = 0, nr = 0;
    for (int kind = 0; kind < nk; ++kind) { ns += pb.send_floats[kind]; nr += pb.recv_floats[kind]; }
    B2P_NCCL(n.Send(p.sendbuf.p + pb.send_off, ns, ncclFloat, pb.peer, p.comm, ctx().stream));
    B2P_NCCL(n.Recv(p.recvbuf.p + pb.recv_off, nr, ncclFloat, pb.peer, p.comm, ctx().stream));
  }
  B2P_NCCL(n.GroupEnd());
}

// number_of_particles handshake + payload (only the spans the receiver reads)
static void exchange_particles(b2p_grid* g) {
  CommPlan& p = *g->comm;
  const int ns = g->cfg.n_species;
  Nccl& n = nccl();
  size_t total_send_cnt = 0, total_recv_cnt = 0;
  for (const PeerBuffers& pb : p.peers) { total_send_cnt += pb.send_order.size() * ns; total_recv_cnt += pb.recv_order.size() * ns; }
  std::vector<unsigned> cs(std::max<size_t>(total_send_cnt, 1)), cr(std::max<size_t>(total_recv_cnt, 1));
  struct Src { const b2p_particle_state* p; unsigned n; };
  std::vector<Src> srcs(total_send_cnt);
  size_t q = 0;
  for (const PeerBuffers& pb : p.peers)
    for (int i : pb.send_order) {
      const PlanEntry& e = p.entries[i];
      b2p_tile* t = g->tiles[g->slot_of_cid[e.cid]];
      if (t->out_ends.size() != size_t(27) * ns) throw Error(B2P_ERR_LOGIC, "pic_particle communication requires pack_outgoing_particles first");
      const int sub = dir_index(e.dir);           // the peer reads my span for the direction from me to it
      for (int s = 0; s < ns; ++s, ++q) {
        const size_t index = 27 * size_t(s) + sub;
        const unsigned long long end = t->out_ends[index], begin = index == 0 ? 0 : t->out_ends[index - 1];
        cs[q] = unsigned(end - begin);
        srcs[q] = Src{ t->out_buf.p + begin, unsigned(end - begin) };
      }
    }
  p.d_cnt_send.reserve(cs.size()); p.d_cnt_recv.reserve(cr.size());
  B2P_CUDA(cudaMemcpyAsync(p.d_cnt_send.p, cs.data(), cs.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx().stream));
  ProfScope* prof_hs_ = new ProfScope(KC_NCCL, 0.0);
  B2P_NCCL(n.GroupStart());
  size_t so = 0, ro = 0;
  for (const PeerBuffers& pb : p.peers) {
    B2P_NCCL(n.Send(p.d_cnt_send.p + so, pb.send_order.size() * ns, ncclUint32, pb.peer, p.comm, ctx().stream));
    B2P_NCCL(n.Recv(p.d_cnt_recv.p + ro, pb.recv_order.size() * ns, ncclUint32, pb.peer, p.comm, ctx().stream));
    so += pb.send_order.size() * ns; ro += pb.recv_order.size() * ns;
  }
  B2P_NCCL(n.GroupEnd());
  delete prof_hs_;
  {
    const auto t0 = std::chrono::steady_clock::now();     // pageable destination: the copy blocks until the stream gets there
    B2P_CUDA(cudaMemcpyAsync(cr.data(), p.d_cnt_recv.p, cr.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx().stream));
    ctx().host_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  sync_stream();
  // payload: contiguous per peer
  size_t send_total = 0, recv_total = 0;
  for (size_t i = 0; i < total_send_cnt; ++i) send_total += cs[i];
  for (size_t i = 0; i < total_recv_cnt; ++i) recv_total += cr[i];
  p.psend.reserve(std::max<size_t>(send_total, 1));
  p.precv.reserve(std::max<size_t>(recv_total, 1));
  std::vector<CopyJob> jobs;
  size_t off = 0;
  unsigned maxc = 0;
  for (size_t i = 0; i < total_send_cnt; ++i) {
    if (srcs[i].n) { jobs.push_back(CopyJob{ srcs[i].p, p.psend.p + off, srcs[i].n }); maxc = std::max(maxc, srcs[i].n); }
    off += srcs[i].n;
  }
  if (!jobs.empty()) {
    Scratch_table_upload(jobs.data(), jobs.size() * sizeof(CopyJob));
    const CopyJob* dj = reinterpret_cast<const CopyJob*>(Scratch_table_ptr());
    for (size_t b = 0; b < jobs.size(); b += 65535) {
      ProfScope prof_(KC_APPEND, 0.0);
      const unsigned nb = unsigned(std::min<size_t>(65535, jobs.size() - b));
      k_copy_spans<<<dim3(std::min((2 * maxc + 255) / 256, 64u), nb), 256, 0, ctx().stream>>>(dj + b);
      B2P_LAUNCH_CHECK();
    }
  }
  ProfScope prof_pl_(KC_NCCL, 0.0);
  B2P_NCCL(n.GroupStart());
  size_t sq = 0, rq = 0, soff = 0, roff = 0;
  p.pspan.assign(p.entries.size(), std::vector<std::pair<size_t, unsigned>>(ns));
  for (const PeerBuffers& pb : p.peers) {
    size_t sbytes = 0, rbytes = 0;
    for (size_t i = 0; i < pb.send_order.size() * ns; ++i) sbytes += size_t(cs[sq + i]);
    size_t o = roff;
    for (size_t i = 0; i < pb.recv_order.size(); ++i)
      for (int s = 0; s < ns; ++s) {
        const unsigned c = cr[rq + i * ns + s];
        p.pspan[pb.recv_order[i]][s] = { o, c };
        o += c; rbytes += c;
      }
    if (sbytes) B2P_NCCL(n.Send(p.psend.p + soff, sbytes * sizeof(b2p_particle_state), ncclChar, pb.peer, p.comm, ctx().stream));
    if (rbytes) B2P_NCCL(n.Recv(p.precv.p + roff, rbytes * sizeof(b2p_particle_state), ncclChar, pb.peer, p.comm, ctx().stream));
    sq += pb.send_order.size() * ns; rq += pb.recv_order.size() * ns;
    soff += sbytes; roff += rbytes;
  }
  B2P_NCCL(n.GroupEnd());
  p.pspan_valid = true;
}

// The exchange of one field mode on the plan's own stream: it starts when the work enqueued so far on the library
// stream is done, and the library stream only waits for it at comm_wait_exchange — kernels enqueued in between run
// concurrently with the pack kernel and the NCCL transfer.
void comm_exchange_fields_on_comm_stream(b2p_grid* g, int mode) {
  CommPlan& p = *g->comm;
  if (!p.cstream) {
    B2P_CUDA(cudaStreamCreateWithFlags(&p.cstream, cudaStreamNonBlocking));
    B2P_CUDA(cudaEventCreateWithFlags(&p.ev_ready, cudaEventDisableTiming));
    B2P_CUDA(cudaEventCreateWithFlags(&p.ev_done, cudaEventDisableTiming));
  }
  Context& c = ctx();
  B2P_CUDA(cudaEventRecord(p.ev_ready, c.stream));
  B2P_CUDA(cudaStreamWaitEvent(p.cstream, p.ev_ready, 0));
  cudaStream_t saved = c.stream;
  c.stream = p.cstream;
  try { exchange_fields(g, mode); } catch (...) { c.stream = saved; throw; }
  c.stream = saved;
  B2P_CUDA(cudaEventRecord(p.ev_done, p.cstream));
}
// the remote-fed halo cells of field `which` (0 E, 1 B, 2 J) from the slabs of the last exchange of that mode
void comm_unpack_halo(b2p_grid* g, int which) {
  CommPlan& p = *g->comm;
  if (p.entries.empty()) return;
  ProfScope prof_(KC_HALO, 0.0);
  for (size_t b = 0; b < p.entries.size(); b += 65535) {
    const unsigned nb = unsigned(std::min<size_t>(65535, p.entries.size() - b));
    k_unpack_halo<<<dim3(24, nb), 256, 0, ctx().stream>>>(g->device_table(), p.d_remote_fill.p + b, p.d_slot_of_entry.p + b, g->g, which);
    B2P_LAUNCH_CHECK();
  }
}
void comm_wait_exchange(b2p_grid* g) {
  CommPlan& p = *g->comm;
  if (p.ev_done) B2P_CUDA(cudaStreamWaitEvent(ctx().stream, p.ev_done, 0));
}

// used by grid_local_communication (host.cu)
bool comm_remote_entry(b2p_grid* g, int slot, int dir_idx, int* entry) {
  if (!g->comm) return false;
  auto it = g->comm->entry_of.find({ slot, dir_idx });
  if (it == g->comm->entry_of.end()) return false;
  *entry = it->second;
  return true;
}
const void* comm_remote_table(b2p_grid* g, int kind) {
  if (!g->comm) return nullptr;
  return kind ? static_cast<const void*>(g->comm->d_remote_exch.p) : static_cast<const void*>(g->comm->d_remote_fill.p);
}
bool comm_particle_span(b2p_grid* g, int entry, int species, const b2p_particle_state** ptr, unsigned* count) {
  CommPlan& p = *g->comm;
  if (!p.pspan_valid) throw Error(B2P_ERR_LOGIC, "pic_particle local communication on a multi-rank grid requires the external exchange first");
  *ptr = p.precv.p + p.pspan[entry][species].first;
  *count = p.pspan[entry][species].second;
  return true;
}
void comm_particles_consumed(b2p_grid* g) { if (g->comm) g->comm->pspan_valid = false; }

}  // namespace b2p

// ==================================================================== C ABI ==
static thread_local std::string g_comm_error;
extern "C" const char* b2p_last_error(void);
namespace b2p { void set_last_error(const std::string& s); }
namespace b2p { void flush_deferred(); }
#define COMM_TRY try { b2p::flush_deferred();
#define COMM_CATCH                                                                 \
  }                                                                                \
  catch (const b2p::Error& e) { b2p::set_last_error(e.what()); return e.code; }    \
  catch (const std::exception& e) { b2p::set_last_error(e.what()); return B2P_ERR_RUNTIME; } \
  return B2P_OK;

extern "C" {

int b2p_nccl_unique_id(void* id128) {
  COMM_TRY
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  B2P_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  COMM_CATCH
}

int b2p_grid_comm_init(b2p_grid* g, int rank, int nranks, const void* id128, const int32_t* owner) {
  COMM_TRY
  if (!g) throw Error(B2P_ERR_RUNTIME, "null grid handle");
  if (owner) g->owner.assign(owner, owner + g->owner.size());
  g->rank = rank; g->nranks = nranks;
  delete g->comm; g->comm = nullptr;
  if (nranks <= 1) { g->nbr_dirty = true; return B2P_OK; }
  if (!id128) throw Error(B2P_ERR_RUNTIME, "comm_init: NCCL unique id required for nranks > 1");
  g->comm = new CommPlan;
  g->comm->rank = rank; g->comm->nranks = nranks;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ctx();
  B2P_NCCL(nccl().CommInitRank(&g->comm->comm, nranks, id, rank));
  finalize_plan(g);
  COMM_CATCH
}

int b2p_grid_external_communication(b2p_grid* g, int mode) {
  COMM_TRY
  if (!g) throw Error(B2P_ERR_RUNTIME, "null grid handle");
  if (g->nranks <= 1 || !g->comm) return B2P_OK;
  switch (mode) {
    case B2P_COMM_EMF_E: case B2P_COMM_EMF_B: case B2P_COMM_EMF_J: exchange_fields(g, mode); break;
    case B2P_COMM_PIC_PARTICLE: exchange_particles(g); break;
    case B2P_COMM_NUMBER_OF_PARTICLES: break;   // folded into the pic_particle exchange
    default: throw Error(B2P_ERR_LOGIC, "external communication does not support given communication mode: " + std::to_string(mode));
  }
  COMM_CATCH
}

// Host-only description of the exchange plan of `rank` (no GPU, no NCCL): used by the
// world_size-2 gloo tests.  Each row is {peer, my cid, dir index, remote cid, send_key, recv_key,
// floats of the interior-edge slab}; returns the number of rows (at most cap are written).
int64_t b2p_plan_describe(const b2p_config* cfg, const int32_t* owner, int rank, int64_t* rows, int64_t cap) {
  try {
    const size_t nt = size_t(cfg->n_tiles[0]) * cfg->n_tiles[1] * cfg->n_tiles[2];
    std::vector<int> own(owner, owner + nt);
    const std::vector<PlanEntry> es = build_plan(*cfg, own, rank);
    for (size_t i = 0; i < es.size() && int64_t(i) < cap; ++i) {
      int64_t* r = rows + 7 * i;
      r[0] = es[i].peer; r[1] = es[i].cid; r[2] = dir_index(es[i].dir); r[3] = es[i].remote_cid;
      r[4] = int64_t(es[i].send_key); r[5] = int64_t(es[i].recv_key); r[6] = int64_t(3 * es[i].volume());
    }
    return int64_t(es.size());
  } catch (...) { return -1; }
}

}  // extern "C"
