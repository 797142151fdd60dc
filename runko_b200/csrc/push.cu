// push.cu — field interpolation + relativistic push + fused zigzag deposit: ONE launch for many containers.
//
// Reference: pic/particle_boris.h:16-62, pic/particle_higuera_cary.h:16-77, pic/particle_faraday.h:38-108,
// emf/yee_lattice_interpolate_linear_1st.h:58-138, pic/particle_current_zigzag_1st.c++:241-338,
// pic/particle.c++:228-262 (leaver test).
//
// blockIdx.y indexes a job table that travels as a kernel argument (PushJobs, 64 containers, 8.7 KB):
// a particle phase is one launch per group of tiles instead of one per container, and no table is
// uploaded.  One thread per slot.
//
// Arithmetic contract: fp32, no FMA contraction (-fmad=false), IEEE div/sqrt, the
// reference's operation order => positions, momenta, cell keys and leaver lists are bit-identical to
// the reference's unfused CPU build.
//
// Measured alternatives that were NOT kept (profiles/r02_hwprobe.txt,
// profiles/r02_ncu_push2_pair_kernel.json): two adjacent slots per thread with all vector algebra on the
// packed fp32x2 pipe, corner reuse and 64-slot aggregation windows (bit-exact; same 740 warp-instructions
// per 32 particles, 128 registers, 184 us against 133 us); nodal box staged in shared memory by
// cp.async.bulk (LDS gathers cost the L1 data pipe what the L1-hit gathers cost: 94 vs 102 cycles per
// warp-gather); TMA bulk reduction of the 48-byte cell records (5.3 vs 6.3 SM-cycles per record).
#include "particles.cuh"
#include "pmath.cuh"

namespace b2p {

struct EB { V3 E, B; };

// emf/yee_lattice_interpolate_linear_1st.h:58-138 on top of the nodal means.
// Node indices fit 32 bits (Ch < 2^31 is checked at tile creation), so all index
// arithmetic is 32-bit; only the four row base addresses are widened.
__device__ __forceinline__ EB interpolate(const float4* __restrict__ nod, const Geom& g, const float3 origo,
                                          const float px, const float py, const float pz) {
  const float lx = px - origo.x, ly = py - origo.y, lz = pz - origo.z;
  const unsigned i = __float2uint_rz(lx), j = __float2uint_rz(ly), k = __float2uint_rz(lz);
  const float dx = lx - float(i), dy = ly - float(j), dz = lz - float(k);
  const unsigned sj = unsigned(g.Hx[2]), si = unsigned(g.Hx[1]) * unsigned(g.Hx[2]);
  const unsigned n = (i * unsigned(g.Hx[1]) + j) * sj + k;
  const float2* __restrict__ nodB = reinterpret_cast<const float2*>(nod + g.Ch);
  const unsigned off[2][2] = { { n, n + sj }, { n + si, n + si + sj } };
  float4 a[2][2][2];
  float2 b[2][2][2];
#pragma unroll
  for (int ic = 0; ic < 2; ++ic)
#pragma unroll
    for (int jc = 0; jc < 2; ++jc)
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        a[ic][jc][kc] = __ldg(nod + off[ic][jc] + kc);
        b[ic][jc][kc] = __ldg(nodB + off[ic][jc] + kc);
      }
  // lerp3D (:29-52): along x, then y, then z — (1-w)*A + w*B per lerp, the same two products and one sum
  // as the reference.  The six components travel as three register pairs {Ex,Ey}, {Ez,Bx}, {By,Bz} — exactly
  // how the LDG.128 / LDG.64 above deliver them — through Blackwell's packed fp32x2 multiply
  // (FMUL2, round-to-nearest per lane): the 84 products take 42 issue slots.  The sums stay scalar FADDs on
  // purpose: ptxas contracts a packed add whose operand is a packed product into FFMA2 even under --fmad=false
  // (checked in SASS), which would change the rounding.
  const float2 wx = make_float2(dx, dx), wy = make_float2(dy, dy), wz = make_float2(dz, dz);
  const float2 ox = make_float2(1.0f - dx, 1.0f - dx), oy = make_float2(1.0f - dy, 1.0f - dy), oz = make_float2(1.0f - dz, 1.0f - dz);
  auto lerp2 = [](const float2 o, const float2 w, const float2 A, const float2 B) {
    const float2 p = __fmul2_rn(o, A), q = __fmul2_rn(w, B);
    return make_float2(__fadd_rn(p.x, q.x), __fadd_rn(p.y, q.y));
  };
#define LERP3P(sel)                                                                              \
  lerp2(oz, wz,                                                                                  \
        lerp2(oy, wy, lerp2(ox, wx, sel(0, 0, 0), sel(1, 0, 0)), lerp2(ox, wx, sel(0, 1, 0), sel(1, 1, 0))), \
        lerp2(oy, wy, lerp2(ox, wx, sel(0, 0, 1), sel(1, 0, 1)), lerp2(ox, wx, sel(0, 1, 1), sel(1, 1, 1))))
#define SEL_XY(i_, j_, k_) make_float2(a[i_][j_][k_].x, a[i_][j_][k_].y)
#define SEL_ZW(i_, j_, k_) make_float2(a[i_][j_][k_].z, a[i_][j_][k_].w)
#define SEL_B(i_, j_, k_) b[i_][j_][k_]
  const float2 exy = LERP3P(SEL_XY), ezbx = LERP3P(SEL_ZW), byz = LERP3P(SEL_B);
#undef SEL_XY
#undef SEL_ZW
#undef SEL_B
#undef LERP3P
  EB eb;
  eb.E.x = exy.x; eb.E.y = exy.y; eb.E.z = ezbx.x;
  eb.B.x = ezbx.y; eb.B.y = byz.x; eb.B.z = byz.y;
  return eb;
}




// `masks` has one uint2 per 32 slots (rounded up to the block).
// FUSE: the zigzag current of every particle that STAYS in the tile box is deposited right here
// from the registers (cell-edge scratch Jc, charge); particles that leave are deposited when they
// arrive in their new tile (k_append), so every tile's J still receives exactly the particles that
// reside in it after migration — the reference's deposit_current, minus one pass over HBM.
template <int PUSHER, int MINB, int FUSE>
__global__ void __launch_bounds__(256, MINB)
k_push(const __grid_constant__ PushJobs jobs, const Geom g, const float cfl, const int agg_min) {
  const PushJob& jb = jobs.job[blockIdx.y];
  if (blockIdx.x * blockDim.x >= jb.s.n) return;
  const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
  // All seven streams are requested before the id is looked at (pinned loads: the compiler must
  // not sink the six value loads below the dead-slot test, which would put two DRAM round trips
  // in series); dead slots hold unspecified but readable values.
  unsigned long long id = DEAD;
  float px = 0.f, py = 0.f, pz = 0.f;
  V3 u = { 0.f, 0.f, 0.f };
  if (n < jb.s.n) {
    id = ld_pinned(jb.s.id + n);
    px = ld_pinned(jb.s.x + n); py = ld_pinned(jb.s.y + n); pz = ld_pinned(jb.s.z + n);
    u.x = ld_pinned(jb.s.ux + n); u.y = ld_pinned(jb.s.uy + n); u.z = ld_pinned(jb.s.uz + n);
  }
  const bool alive = id != DEAD;                                   // :33
  float nx = 0.f, ny = 0.f, nz = 0.f;
  V3 vel = { 0.f, 0.f, 0.f };
  if (alive) {
  const EB eb = interpolate(jb.nod, g, jb.origo, px, py, pz);
  const float qm = jb.qm;
  const DivC div_cfl(cfl);
  if (PUSHER == B2P_PUSHER_BORIS) {                                // pic/particle_boris.h:37-59
    const V3 v0 = cfl * u;
    const V3 E0 = 0.5f * qm * eb.E;
    const V3 u0 = v0 + E0;
    const float ginv = cfl / sqrtf(cfl * cfl + dot(u0, u0));
    const V3 B0 = div_cfl(0.5f * qm * ginv * eb.B);
    const float f = 2.0f / (1.0f + dot(B0, B0));
    const V3 u1 = f * (u0 + cross(u0, B0));
    const V3 u2 = u0 + cross(u1, B0) + E0;
    vel = div_cfl(u2);
    const float ginv2 = cfl / sqrtf(cfl * cfl + dot(u2, u2));
    nx = px + vel.x * ginv2 * cfl; ny = py + vel.y * ginv2 * cfl; nz = pz + vel.z * ginv2 * cfl;
  } else if (PUSHER == B2P_PUSHER_HIGUERA_CARY) {                  // pic/particle_higuera_cary.h:26-75
    const float hqm = 0.5f * qm, cfl2 = cfl * cfl, cinv = 1.0f / cfl, cinv2 = cinv * cinv;
    const V3 v0 = cfl * u;
    const V3 E0 = hqm * eb.E;
    const V3 u0 = v0 + E0;
    const V3 Bt = hqm * eb.B;
    const float u0sq = dot(u0, u0), b2 = dot(Bt, Bt), bdotu = dot(Bt, u0);
    const float gmb = 1.0f + u0sq * cinv2 - b2 * cinv2;
    const float disc = gmb * gmb + 4.0f * (b2 * cinv2 + bdotu * bdotu * cinv2);
    const float ginv = 1.0f / sqrtf(0.5f * (gmb + sqrtf(disc)));
    const float gc = ginv * cinv;
    const V3 B0 = gc * Bt;
    const float f = 2.0f / (1.0f + gc * gc * b2);
    const V3 u1 = f * (u0 + cross(u0, B0));
    const V3 u2 = u0 + cross(u1, B0) + E0;
    const float ginv2 = cfl / sqrtf(cfl2 + dot(u2, u2));
    vel = u2 * cinv;
    nx = px + u2.x * ginv2; ny = py + u2.y * ginv2; nz = pz + u2.z * ginv2;
  } else {                                                         // pic/particle_faraday.h:53-107
    const V3 v0 = cfl * u;
    const float gcfl = sqrtf(cfl * cfl + dot(v0, v0));
    const V3 u0 = v0 + 0.5f * qm * eb.E;
    const float geff_cfl = sqrtf(cfl * cfl + dot(u0, u0));
    const float kappa = 0.5f * qm / geff_cfl;
    const V3 eps = kappa * eb.E;
    const V3 beta = kappa * eb.B;
    const float w0 = gcfl + dot(eps, v0);
    const V3 W = v0 + eps * gcfl + cross(v0, beta) + w0 * eps;
    const float b2 = dot(beta, beta);
    const float f = 1.0f / (1.0f + b2);
    const V3 W_rot = f * (W - cross(beta, W) + dot(beta, W) * beta);
    const float bde = dot(beta, eps);
    const V3 eps_rot = f * (eps - cross(beta, eps) + bde * beta);
    const float D = 1.0f - f * (dot(eps, eps) + bde * bde);
    const V3 u2 = W_rot + eps_rot * (dot(eps, W_rot) / D);
    vel = div_cfl(u2);
    const float ginv2 = cfl / sqrtf(cfl * cfl + dot(u2, u2));
    nx = px + vel.x * ginv2 * cfl; ny = py + vel.y * ginv2 * cfl; nz = pz + vel.z * ginv2 * cfl;
  }
  jb.s.ux[n] = vel.x; jb.s.uy[n] = vel.y; jb.s.uz[n] = vel.z;
  jb.s.x[n] = nx; jb.s.y[n] = ny; jb.s.z[n] = nz;
  }
  if (jb.ke) {
    // kinetic-energy account (particles.cuh: KE_SLOTS): dead slots carry vel = 0, i.e. sqrt(1) - 1 = 0
    float e = sqrtf(1.0f + dot(vel, vel)) - 1.0f;                   // pic/particle.c++:352-377
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e != 0.0f) {
      const unsigned gw = (blockIdx.y * gridDim.x + blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);   // warp of the launch
      atomicAdd(jb.ke + ((gw * 2654435761u) >> 19), double(e));                                            // scattered over the 2^13 slots
    }
  }
  const bool inside = inside_box(nx, ny, nz, jb.mn, jb.mx);
  publish_masks(alive, inside, n, jb.masks);
  if (FUSE) {
    const bool stays = alive && inside;
    Zigzag z;
    if (stays) z = zigzag_split(V3{ nx, ny, nz }, vel, jb.origo, cfl, jb.charge, g);
    deposit_split<(FUSE > 1)>(stays, z, jb.Jc, agg_min);
  }
}

// One launch for the first `njobs` containers of the table (one geometry / pusher / cfl).
void launch_push_jobs(int pusher, const PushJobs& jobs, int njobs, unsigned max_n, double total_slots, const Geom& g, float cfl,
                      bool fuse) {
  ProfScope prof_(KC_PUSH, total_slots);
  if (!njobs || !max_n) return;
  if (njobs > PUSH_JOBS_MAX) throw Error(B2P_ERR_LOGIC, "launch_push_jobs: job table overflow");
  const int f = fuse ? (tuning().deposit_agg ? 2 : 1) : 0;
  const int agg_min = tuning().agg_min;
  const unsigned bs = tuning().push_block == 128 ? 128u : 256u;
  const dim3 grid((max_n + bs - 1) / bs, unsigned(njobs));
  const int minb = tuning().push_minb;
#define PUSH_LAUNCH(P, M, F) k_push<P, M, F><<<grid, bs, 0, ctx().stream>>>(jobs, g, cfl, agg_min)
#define PUSH_CASE(P)                                                                                   \
  case P:                                                                                              \
    if (f == 2) { if (minb >= 6) PUSH_LAUNCH(P, 6, 2); else if (minb >= 5) PUSH_LAUNCH(P, 5, 2); else PUSH_LAUNCH(P, 4, 2); } \
    else if (f == 1) { if (minb >= 6) PUSH_LAUNCH(P, 6, 1); else PUSH_LAUNCH(P, 4, 1); }              \
    else { if (minb >= 6) PUSH_LAUNCH(P, 6, 0); else PUSH_LAUNCH(P, 5, 0); }                          \
    break;
  switch (pusher) {
    PUSH_CASE(B2P_PUSHER_BORIS)
    PUSH_CASE(B2P_PUSHER_HIGUERA_CARY)
    PUSH_CASE(B2P_PUSHER_FARADAY)
    default: throw Error(B2P_ERR_LOGIC, "pic::Tile::push_particles: unkown particle pusher");
  }
#undef PUSH_CASE
#undef PUSH_LAUNCH
  B2P_LAUNCH_CHECK();
}

}  // namespace b2p
