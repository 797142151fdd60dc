// push.cu — field interpolation + relativistic push + fused zigzag deposit, two slots per thread.
//
// Reference: pic/particle_boris.h:16-62, pic/particle_higuera_cary.h:16-77, pic/particle_faraday.h:38-108,
// emf/yee_lattice_interpolate_linear_1st.h:58-138, pic/particle_current_zigzag_1st.c++:241-338,
// pic/particle.c++:228-262 (leaver test).
//
// Shape of the kernel (what the B200 measurements of profiles/r02_hwprobe.txt ask for):
//   * one launch pushes MANY containers: blockIdx.y indexes a device job table (PushJob), so a
//     particle phase is a handful of launches instead of one per container;
//   * a thread owns the ADJACENT slots 2p, 2p+1 of a container: the seven streams arrive through
//     LDG.64 / LDG.128 and leave through STG.64, and all per-particle vector algebra (pusher and
//     zigzag split) runs on Blackwell's packed fp32x2 pipe with lane .x = slot 2p, lane .y = slot
//     2p+1 — half the issue slots of the scalar kernel for the same IEEE operations;
//   * in a cell-sorted container the two slots usually sit in the same cell: the eight nodal
//     records gathered for slot 2p are reused for slot 2p+1 (no second gather), and the pair's
//     deposit records are merged in registers before the warp-level run aggregation, which now
//     spans 64 slots per warp instead of 32 — fewer gathers and fewer RED.128 per particle
//     through the L1 data pipe, the unit that bounded the one-slot-per-thread kernel.
//
// Arithmetic contract: every sum/product below is the reference's, rounded once,
// in the reference's order.  Packed products are FMUL2; packed sums are FFMA2(a, 1, b) — the fma
// rounds the exact a*1 + b once, i.e. gives the bits of the add — with the 1 a kernel parameter,
// because ptxas contracts a packed add with a packed product feeding it even under -fmad=false.
#include "particles.cuh"
#include "pmath.cuh"

namespace b2p {

namespace {

struct P3 { float2 x, y, z; };   // a 3-vector for the thread's two particles

struct Kc {                      // packed constants
  float2 one, mone, zero, half, c1;   // (1,1) and (-1,-1) opaque to the compiler; (0,0), (.5,.5), (1,1) literal
};

#define MUL2(a, b) __fmul2_rn((a), (b))
// ADD2(a, b): `a` must never be a compile-time constant — ptxas would fold literal * one into a plain packed add and
// then contract that add with a product feeding `b` (seen as a 1-ulp mismatch of the Higuera-Cary gamma).
#define ADD2(a, b) __ffma2_rn((a), K.one, (b))     // a + b
#define SUB2(a, b) __ffma2_rn((b), K.mone, (a))    // a - b   ((-b) + a, one rounding)

__device__ __forceinline__ float2 splat(const float v) { return make_float2(v, v); }

__device__ __forceinline__ P3 scale(const float2 s, const P3& a) { return { MUL2(a.x, s), MUL2(a.y, s), MUL2(a.z, s) }; }
__device__ __forceinline__ P3 add(const P3& a, const P3& b, const Kc& K) { return { ADD2(a.x, b.x), ADD2(a.y, b.y), ADD2(a.z, b.z) }; }
__device__ __forceinline__ P3 sub(const P3& a, const P3& b, const Kc& K) { return { SUB2(a.x, b.x), SUB2(a.y, b.y), SUB2(a.z, b.z) }; }
// tools/vector.h:248-255: accumulates from 0, left to right
__device__ __forceinline__ float2 dot(const P3& a, const P3& b, const Kc& K) {
  float2 r = ADD2(MUL2(a.x, b.x), K.zero);
  r = ADD2(r, MUL2(a.y, b.y));
  return ADD2(r, MUL2(a.z, b.z));
}
// tools/vector.h:281-289: { a.y*b.z - a.z*b.y, -a.x*b.z + a.z*b.x, a.x*b.y - a.y*b.x }; (-p) + q == q - p bit for bit
__device__ __forceinline__ P3 cross(const P3& a, const P3& b, const Kc& K) {
  return { SUB2(MUL2(a.y, b.z), MUL2(a.z, b.y)), SUB2(MUL2(a.z, b.x), MUL2(a.x, b.z)), SUB2(MUL2(a.x, b.y), MUL2(a.y, b.x)) };
}
__device__ __forceinline__ float2 sqrt2(const float2 v) { return make_float2(sqrtf(v.x), sqrtf(v.y)); }
__device__ __forceinline__ float2 div2(const float2 a, const float2 b) { return make_float2(a.x / b.x, a.y / b.y); }

// pmath.cuh DivC for a pair: the exact constant division's fast path on the packed pipe, the range
// test and the plain-division fallback per particle.
struct DivC2 {
  float c, hi;
  float2 rc, negc;
  __device__ __forceinline__ explicit DivC2(const float c_) : c(c_) {
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(c_));
    const float r1 = fmaf(r0, fmaf(-c_, r0, 1.0f), r0);
    rc = splat(r1);
    negc = splat(-c_);
    hi = (fabsf(c_) >= 0x1p-20f && fabsf(c_) <= 0x1p20f) ? 0x1p100f : -1.0f;
  }
  __device__ __forceinline__ float2 fast(const float2 x) const {
    const float2 q = MUL2(x, rc);
    return __ffma2_rn(__ffma2_rn(negc, q, x), rc, q);
  }
  __device__ __forceinline__ P3 operator()(const P3& v) const {
    P3 r = { fast(v.x), fast(v.y), fast(v.z) };
    const float loA = fminf(fminf(fabsf(v.x.x), fabsf(v.y.x)), fabsf(v.z.x)), mxA = fmaxf(fmaxf(fabsf(v.x.x), fabsf(v.y.x)), fabsf(v.z.x));
    const float loB = fminf(fminf(fabsf(v.x.y), fabsf(v.y.y)), fabsf(v.z.y)), mxB = fmaxf(fmaxf(fabsf(v.x.y), fabsf(v.y.y)), fabsf(v.z.y));
    if (!(loA >= 0x1p-100f && mxA <= hi)) { r.x.x = v.x.x / c; r.y.x = v.y.x / c; r.z.x = v.z.x / c; }
    if (!(loB >= 0x1p-100f && mxB <= hi)) { r.x.y = v.x.y / c; r.y.y = v.y.y / c; r.z.y = v.z.y / c; }
    return r;
  }
};

// ---- interpolation (emf/yee_lattice_interpolate_linear_1st.h:58-138 on top of the nodal means) ----
struct Corners { float4 a[8]; float2 b[8]; };   // corner q = ic*4 + jc*2 + kc

__device__ __forceinline__ void load_corners(Corners& c, const float4* __restrict__ nodA, const float2* __restrict__ nodB,
                                             const unsigned n, const unsigned sj, const unsigned si) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const unsigned o = n + ((q >> 2) & 1) * si + ((q >> 1) & 1) * sj + (q & 1);
    c.a[q] = __ldg(nodA + o);
    c.b[q] = __ldg(nodB + o);
  }
}

// lerp3D (:29-52): along x, then y, then z; (1-w)*A + w*B per lerp.  The six components travel as the register
// pairs {Ex,Ey}, {Ez,Bx}, {By,Bz} the LDG.128 / LDG.64 deliver.
struct EBq { float2 exy, ezbx, byz; };
__device__ __forceinline__ EBq lerp_corners(const Corners& c, const float dx, const float dy, const float dz, const Kc& K) {
  const float2 wx = splat(dx), wy = splat(dy), wz = splat(dz);
  const float2 ox = splat(1.0f - dx), oy = splat(1.0f - dy), oz = splat(1.0f - dz);
  auto lerp2 = [&K](const float2 o, const float2 w, const float2 A, const float2 B) {
    const float2 p = MUL2(o, A), q = MUL2(w, B);
    return ADD2(p, q);
  };
#define B2P_LERP3(sel)                                                                                                   \
  lerp2(oz, wz, lerp2(oy, wy, lerp2(ox, wx, sel(0), sel(4)), lerp2(ox, wx, sel(2), sel(6))),                            \
        lerp2(oy, wy, lerp2(ox, wx, sel(1), sel(5)), lerp2(ox, wx, sel(3), sel(7))))
#define SEL_XY(q) make_float2(c.a[q].x, c.a[q].y)
#define SEL_ZW(q) make_float2(c.a[q].z, c.a[q].w)
#define SEL_B(q) c.b[q]
  EBq r;
  r.exy = B2P_LERP3(SEL_XY);
  r.ezbx = B2P_LERP3(SEL_ZW);
  r.byz = B2P_LERP3(SEL_B);
#undef SEL_XY
#undef SEL_ZW
#undef SEL_B
#undef B2P_LERP3
  return r;
}

// ---- the three pushers on a pair; pos and u are updated in place (u -> stored velocity) ----
template <int PUSHER>
__device__ __forceinline__ void push_pair(const P3& E, const P3& B, P3& pos, P3& u, const float cfl, const float qm, const Kc& K) {
  const float2 c2 = splat(cfl), cc2 = splat(cfl * cfl), hq2 = splat(0.5f * qm);
  const DivC2 div_cfl(cfl);
  if (PUSHER == B2P_PUSHER_BORIS) {                                // pic/particle_boris.h:37-59
    const P3 v0 = scale(c2, u);
    const P3 E0 = scale(hq2, E);
    const P3 u0 = add(v0, E0, K);
    const float2 ginv = div2(c2, sqrt2(ADD2(dot(u0, u0, K), cc2)));
    const P3 B0 = div_cfl(scale(MUL2(hq2, ginv), B));
    const float2 f = div2(splat(2.0f), ADD2(dot(B0, B0, K), K.c1));
    const P3 u1 = scale(f, add(u0, cross(u0, B0, K), K));
    const P3 u2 = add(add(u0, cross(u1, B0, K), K), E0, K);
    u = div_cfl(u2);
    const float2 ginv2 = div2(c2, sqrt2(ADD2(dot(u2, u2, K), cc2)));
    pos.x = ADD2(pos.x, MUL2(MUL2(u.x, ginv2), c2));
    pos.y = ADD2(pos.y, MUL2(MUL2(u.y, ginv2), c2));
    pos.z = ADD2(pos.z, MUL2(MUL2(u.z, ginv2), c2));
  } else if (PUSHER == B2P_PUSHER_HIGUERA_CARY) {                  // pic/particle_higuera_cary.h:26-75
    const float cinv = 1.0f / cfl;
    const float2 ci2 = splat(cinv), cinv2 = splat(cinv * cinv);
    const P3 v0 = scale(c2, u);
    const P3 E0 = scale(hq2, E);
    const P3 u0 = add(v0, E0, K);
    const P3 Bt = scale(hq2, B);
    const float2 u0sq = dot(u0, u0, K), b2 = dot(Bt, Bt, K), bdotu = dot(Bt, u0, K);
    const float2 gmb = SUB2(ADD2(MUL2(u0sq, cinv2), K.c1), MUL2(b2, cinv2));
    const float2 disc = ADD2(MUL2(gmb, gmb), MUL2(splat(4.0f), ADD2(MUL2(b2, cinv2), MUL2(MUL2(bdotu, bdotu), cinv2))));
    const float2 ginv = div2(K.c1, sqrt2(MUL2(K.half, ADD2(gmb, sqrt2(disc)))));
    const float2 gc = MUL2(ginv, ci2);
    const P3 B0 = scale(gc, Bt);
    const float2 f = div2(splat(2.0f), ADD2(MUL2(MUL2(gc, gc), b2), K.c1));
    const P3 u1 = scale(f, add(u0, cross(u0, B0, K), K));
    const P3 u2 = add(add(u0, cross(u1, B0, K), K), E0, K);
    const float2 ginv2 = div2(c2, sqrt2(ADD2(dot(u2, u2, K), cc2)));
    u = scale(ci2, u2);
    pos.x = ADD2(pos.x, MUL2(u2.x, ginv2));
    pos.y = ADD2(pos.y, MUL2(u2.y, ginv2));
    pos.z = ADD2(pos.z, MUL2(u2.z, ginv2));
  } else {                                                         // pic/particle_faraday.h:53-107
    const P3 v0 = scale(c2, u);
    const float2 gcfl = sqrt2(ADD2(dot(v0, v0, K), cc2));
    const P3 u0 = add(v0, scale(hq2, E), K);
    const float2 geff_cfl = sqrt2(ADD2(dot(u0, u0, K), cc2));
    const float2 kappa = div2(hq2, geff_cfl);
    const P3 eps = scale(kappa, E);
    const P3 beta = scale(kappa, B);
    const float2 w0 = ADD2(gcfl, dot(eps, v0, K));
    const P3 W = add(add(add(v0, scale(gcfl, eps), K), cross(v0, beta, K), K), scale(w0, eps), K);
    const float2 b2 = dot(beta, beta, K);
    const float2 f = div2(K.c1, ADD2(b2, K.c1));
    const P3 W_rot = scale(f, add(sub(W, cross(beta, W, K), K), scale(dot(beta, W, K), beta), K));
    const float2 bde = dot(beta, eps, K);
    const P3 eps_rot = scale(f, add(sub(eps, cross(beta, eps, K), K), scale(bde, beta), K));
    const float2 D = SUB2(K.c1, MUL2(f, ADD2(dot(eps, eps, K), MUL2(bde, bde))));
    const P3 u2 = add(W_rot, scale(div2(dot(eps, W_rot, K), D), eps_rot), K);
    u = div_cfl(u2);
    const float2 ginv2 = div2(c2, sqrt2(ADD2(dot(u2, u2, K), cc2)));
    pos.x = ADD2(pos.x, MUL2(MUL2(u.x, ginv2), c2));
    pos.y = ADD2(pos.y, MUL2(MUL2(u.y, ginv2), c2));
    pos.z = ADD2(pos.z, MUL2(MUL2(u.z, ginv2), c2));
  }
}

// ---- zigzag split of a pair (pic/particle_current_zigzag_1st.c++:241-336; scalar twin: pmath.cuh zigzag_split) ----
struct Rec { unsigned n; float4 ex, ey, ez; };   // the 12 edge currents of one segment in cell n (layout: pmath.cuh)

__device__ __forceinline__ float relay1(const float f1, const float f2, const float mid) {
  const float lo = (f1 < f2 ? f1 : f2) + 1.0f;
  const float b1 = f1 > f2 ? f1 : f2;
  const float b = b1 > mid ? b1 : mid;
  return lo < b ? lo : b;
}

__device__ __forceinline__ void zigzag_pair(const P3& pos, const P3& u, const float3 origo, const float cfl, const float charge,
                                            const Geom& g, const Kc& K, Rec& a1, Rec& a2, Rec& b1, Rec& b2) {
  const float2 g2 = ADD2(dot(u, u, K), K.c1);
  const float2 invgam = div2(K.c1, sqrt2(g2));
  const P3 x2 = { SUB2(pos.x, splat(origo.x)), SUB2(pos.y, splat(origo.y)), SUB2(pos.z, splat(origo.z)) };
  const float2 t = MUL2(splat(cfl), invgam);
  const P3 x1 = sub(x2, scale(t, u), K);
  const P3 fi1 = { make_float2(floorf(x1.x.x), floorf(x1.x.y)), make_float2(floorf(x1.y.x), floorf(x1.y.y)), make_float2(floorf(x1.z.x), floorf(x1.z.y)) };
  const P3 fi2 = { make_float2(floorf(x2.x.x), floorf(x2.x.y)), make_float2(floorf(x2.y.x), floorf(x2.y.y)), make_float2(floorf(x2.z.x), floorf(x2.z.y)) };
  const P3 mid = scale(K.half, add(x1, x2, K));
  const P3 xr = { make_float2(relay1(fi1.x.x, fi2.x.x, mid.x.x), relay1(fi1.x.y, fi2.x.y, mid.x.y)),
                  make_float2(relay1(fi1.y.x, fi2.y.x, mid.y.x), relay1(fi1.y.y, fi2.y.y, mid.y.y)),
                  make_float2(relay1(fi1.z.x, fi2.z.x, mid.z.x), relay1(fi1.z.y, fi2.z.y, mid.z.y)) };
  const float2 q2 = splat(charge);
  const P3 F1 = scale(q2, sub(xr, x1, K));
  const P3 F2 = scale(q2, sub(x2, xr, K));
  const P3 W1 = sub(scale(K.half, add(x1, xr, K)), fi1, K);
  const P3 W2 = sub(scale(K.half, add(x2, xr, K)), fi2, K);
  const unsigned Hy = unsigned(g.Hx[1]), Hz = unsigned(g.Hx[2]);
  a1.n = (__float2uint_rz(fi1.x.x) * Hy + __float2uint_rz(fi1.y.x)) * Hz + __float2uint_rz(fi1.z.x);
  b1.n = (__float2uint_rz(fi1.x.y) * Hy + __float2uint_rz(fi1.y.y)) * Hz + __float2uint_rz(fi1.z.y);
  a2.n = (__float2uint_rz(fi2.x.x) * Hy + __float2uint_rz(fi2.y.x)) * Hz + __float2uint_rz(fi2.z.x);
  b2.n = (__float2uint_rz(fi2.x.y) * Hy + __float2uint_rz(fi2.y.y)) * Hz + __float2uint_rz(fi2.z.y);
  // F.c * wa * wb in the reference's association (F.c * wa) * wb, for the four (wa, wb) corner weights of each component
#define B2P_EDGES(F, W, ra, rb)                                                                    \
  {                                                                                                \
    const P3 o = { SUB2(K.c1, W.x), SUB2(K.c1, W.y), SUB2(K.c1, W.z) };                            \
    float2 t0, t1, e0, e1, e2, e3;                                                                 \
    t0 = MUL2(F.x, o.y); t1 = MUL2(F.x, W.y);                                                      \
    e0 = MUL2(t0, o.z); e1 = MUL2(t1, o.z); e2 = MUL2(t0, W.z); e3 = MUL2(t1, W.z);                \
    ra.ex = make_float4(e0.x, e1.x, e2.x, e3.x); rb.ex = make_float4(e0.y, e1.y, e2.y, e3.y);      \
    t0 = MUL2(F.y, o.x); t1 = MUL2(F.y, W.x);                                                      \
    e0 = MUL2(t0, o.z); e1 = MUL2(t1, o.z); e2 = MUL2(t0, W.z); e3 = MUL2(t1, W.z);                \
    ra.ey = make_float4(e0.x, e1.x, e2.x, e3.x); rb.ey = make_float4(e0.y, e1.y, e2.y, e3.y);      \
    t0 = MUL2(F.z, o.x); t1 = MUL2(F.z, W.x);                                                      \
    e0 = MUL2(t0, o.y); e1 = MUL2(t1, o.y); e2 = MUL2(t0, W.y); e3 = MUL2(t1, W.y);                \
    ra.ez = make_float4(e0.x, e1.x, e2.x, e3.x); rb.ez = make_float4(e0.y, e1.y, e2.y, e3.y);      \
  }
  B2P_EDGES(F1, W1, a1, b1)
  B2P_EDGES(F2, W2, a2, b2)
#undef B2P_EDGES
}

__device__ __forceinline__ void rec_add(Rec& d, const Rec& s) {
  d.ex.x += s.ex.x; d.ex.y += s.ex.y; d.ex.z += s.ex.z; d.ex.w += s.ex.w;
  d.ey.x += s.ey.x; d.ey.y += s.ey.y; d.ey.z += s.ey.z; d.ey.w += s.ey.w;
  d.ez.x += s.ez.x; d.ez.y += s.ez.y; d.ez.z += s.ez.z; d.ez.w += s.ez.w;
}

constexpr unsigned NOCELL = 0xFFFFFFFFu;

// bits of a (low 16) to the even positions, bits of b (low 16) to the odd ones
__device__ __forceinline__ unsigned spread16(unsigned x) {
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__device__ __forceinline__ unsigned interleave16(const unsigned a, const unsigned b) { return spread16(a & 0xFFFFu) | (spread16(b & 0xFFFFu) << 1); }

}  // namespace

constexpr int PUSH2_THREADS = 256;
constexpr int PUSH2_CHUNKS = 2;                                    // 512-slot chunks per block
constexpr unsigned PUSH2_BLOCK_SLOTS = 2u * PUSH2_THREADS * PUSH2_CHUNKS;

// FUSE 0: push only; 1: stayers' zigzag current deposited here (plain per-lane REDs); 2: ... with warp run aggregation
template <int PUSHER, int FUSE>
__global__ void __launch_bounds__(PUSH2_THREADS, 2)
k_push2(const __grid_constant__ PushJobs jobs, const Geom g, const float cfl, const float one, const int agg_min) {
  const PushJob& jb = jobs.job[blockIdx.y];
  const unsigned n_total = jb.s.n;
  const unsigned block_first = blockIdx.x * PUSH2_BLOCK_SLOTS;
  if (block_first >= n_total) return;
  Kc K;
  K.one = splat(one); K.mone = splat(-one); K.zero = splat(0.0f); K.half = splat(0.5f); K.c1 = splat(1.0f);
  const unsigned lane = threadIdx.x & 31;
  const unsigned sj = unsigned(g.Hx[2]), si = unsigned(g.Hx[1]) * unsigned(g.Hx[2]);
  const float4* __restrict__ nodA = jb.nod;
  const float2* __restrict__ nodB = reinterpret_cast<const float2*>(jb.nod + g.Ch);

#pragma unroll 1
  for (int ch = 0; ch < PUSH2_CHUNKS; ++ch) {
    const unsigned s0 = block_first + unsigned(ch) * (2u * PUSH2_THREADS) + 2u * threadIdx.x;   // my even slot
    if (s0 - 2u * lane >= n_total) break;                          // the whole warp is past the end (warp-uniform)
    // ---- load: all seven streams before the ids are looked at; dead slots hold unspecified but readable values
    //      (capacities are even, so the odd slot of the last pair is allocated even when it is not part of the container)
    ulonglong2 id2 = make_ulonglong2(DEAD, DEAD);
    P3 pos = { K.zero, K.zero, K.zero }, u = { K.zero, K.zero, K.zero };
    if (s0 < n_total) {
      const unsigned p = s0 >> 1;
      asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(id2.x), "=l"(id2.y) : "l"(reinterpret_cast<const ulonglong2*>(jb.s.id) + p));
#define B2P_LD2(dst, ptr) asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(dst.x), "=f"(dst.y) : "l"(reinterpret_cast<const float2*>(ptr) + p))
      B2P_LD2(pos.x, jb.s.x); B2P_LD2(pos.y, jb.s.y); B2P_LD2(pos.z, jb.s.z);
      B2P_LD2(u.x, jb.s.ux); B2P_LD2(u.y, jb.s.uy); B2P_LD2(u.z, jb.s.uz);
#undef B2P_LD2
    }
    const bool aliveA = id2.x != DEAD;                             // pic/particle_boris.h:33
    const bool aliveB = id2.y != DEAD && s0 + 1u < n_total;
    bool insideA = false, insideB = false;
    if (aliveA || aliveB) {
      // a dead partner computes on a copy of the alive particle (same cell: no gather of its own; results dropped)
      if (!aliveA) { pos.x.x = pos.x.y; pos.y.x = pos.y.y; pos.z.x = pos.z.y; u.x.x = u.x.y; u.y.x = u.y.y; u.z.x = u.z.y; }
      if (!aliveB) { pos.x.y = pos.x.x; pos.y.y = pos.y.x; pos.z.y = pos.z.x; u.x.y = u.x.x; u.y.y = u.y.x; u.z.y = u.z.x; }
      // ---- interpolate (p = pos - origo; (i,j,k) = trunc(p))
      const float2 lx = SUB2(pos.x, splat(jb.origo.x)), ly = SUB2(pos.y, splat(jb.origo.y)), lz = SUB2(pos.z, splat(jb.origo.z));
      const unsigned iA = __float2uint_rz(lx.x), jA = __float2uint_rz(ly.x), kA = __float2uint_rz(lz.x);
      const unsigned iB = __float2uint_rz(lx.y), jB = __float2uint_rz(ly.y), kB = __float2uint_rz(lz.y);
      const unsigned nA = (iA * unsigned(g.Hx[1]) + jA) * sj + kA, nB = (iB * unsigned(g.Hx[1]) + jB) * sj + kB;
      Corners c;
      load_corners(c, nodA, nodB, nA, sj, si);
      const EBq fa = lerp_corners(c, lx.x - float(iA), ly.x - float(jA), lz.x - float(kA), K);
      if (nB != nA) load_corners(c, nodA, nodB, nB, sj, si);
      const EBq fb = lerp_corners(c, lx.y - float(iB), ly.y - float(jB), lz.y - float(kB), K);
      const P3 E = { make_float2(fa.exy.x, fb.exy.x), make_float2(fa.exy.y, fb.exy.y), make_float2(fa.ezbx.x, fb.ezbx.x) };
      const P3 B = { make_float2(fa.ezbx.y, fb.ezbx.y), make_float2(fa.byz.x, fb.byz.x), make_float2(fa.byz.y, fb.byz.y) };
      // ---- push
      push_pair<PUSHER>(E, B, pos, u, cfl, jb.qm, K);
      // ---- store (dead slots are left untouched)
      const unsigned p = s0 >> 1;
      if (aliveA && aliveB) {
        reinterpret_cast<float2*>(jb.s.ux)[p] = u.x; reinterpret_cast<float2*>(jb.s.uy)[p] = u.y; reinterpret_cast<float2*>(jb.s.uz)[p] = u.z;
        reinterpret_cast<float2*>(jb.s.x)[p] = pos.x; reinterpret_cast<float2*>(jb.s.y)[p] = pos.y; reinterpret_cast<float2*>(jb.s.z)[p] = pos.z;
      } else if (aliveA) {
        jb.s.ux[s0] = u.x.x; jb.s.uy[s0] = u.y.x; jb.s.uz[s0] = u.z.x;
        jb.s.x[s0] = pos.x.x; jb.s.y[s0] = pos.y.x; jb.s.z[s0] = pos.z.x;
      } else {
        jb.s.ux[s0 + 1] = u.x.y; jb.s.uy[s0 + 1] = u.y.y; jb.s.uz[s0 + 1] = u.z.y;
        jb.s.x[s0 + 1] = pos.x.y; jb.s.y[s0 + 1] = pos.y.y; jb.s.z[s0 + 1] = pos.z.y;
      }
      insideA = inside_box(pos.x.x, pos.y.x, pos.z.x, jb.mn, jb.mx);
      insideB = inside_box(pos.x.y, pos.y.y, pos.z.y, jb.mn, jb.mx);
    }
    // ---- leaver / stayer ballots of the warp's 64 slots: words 2w (lanes 0-15) and 2w+1 (lanes 16-31)
    {
      const unsigned lA = __ballot_sync(0xffffffffu, aliveA && !insideA), lB = __ballot_sync(0xffffffffu, aliveB && !insideB);
      const unsigned sA = __ballot_sync(0xffffffffu, aliveA && insideA), sB = __ballot_sync(0xffffffffu, aliveB && insideB);
      if (lane == 0) {
        const uint4 w = make_uint4(interleave16(lA, lB), interleave16(sA, sB), interleave16(lA >> 16, lB >> 16), interleave16(sA >> 16, sB >> 16));
        *reinterpret_cast<uint4*>(jb.masks + (s0 >> 5)) = w;       // s0 is a multiple of 64 here: 16-byte aligned
      }
    }
    // ---- fused deposit of the particles that stay in the tile box
    if (FUSE) {
      const bool stA = aliveA && insideA, stB = aliveB && insideB;
      Rec a1, a2, b1, b2;
      a1.n = a2.n = b1.n = b2.n = NOCELL;
      if (stA || stB) {
        zigzag_pair(pos, u, jb.origo, cfl, jb.charge, g, K, a1, a2, b1, b2);
        if (!stA) a1.n = a2.n = NOCELL;
        if (!stB) b1.n = b2.n = NOCELL;
        // both segments in one cell (about half of a thermal plasma): one record
        if (a1.n == a2.n && stA) { rec_add(a2, a1); a1.n = NOCELL; }
        if (b1.n == b2.n && stB) { rec_add(b2, b1); b1.n = NOCELL; }
        // the pair's records that share a cell (neighbouring slots of a sorted container): one record
        if (b2.n != NOCELL && b2.n == a2.n) { rec_add(a2, b2); b2.n = NOCELL; }
        if (b1.n != NOCELL && b1.n == a1.n) { rec_add(a1, b1); b1.n = NOCELL; }
      }
      if (__any_sync(0xffffffffu, a2.n != NOCELL)) reduce_runs_and_red<(FUSE > 1)>(a2.n, a2.ex, a2.ey, a2.ez, jb.Jc, agg_min);
      if (__any_sync(0xffffffffu, b2.n != NOCELL)) reduce_runs_and_red<(FUSE > 1)>(b2.n, b2.ex, b2.ey, b2.ez, jb.Jc, agg_min);
      if (__any_sync(0xffffffffu, a1.n != NOCELL)) reduce_runs_and_red<(FUSE > 1)>(a1.n, a1.ex, a1.ey, a1.ez, jb.Jc, agg_min);
      if (__any_sync(0xffffffffu, b1.n != NOCELL)) reduce_runs_and_red<(FUSE > 1)>(b1.n, b1.ex, b1.ey, b1.ez, jb.Jc, agg_min);
    }
  }
}

#undef MUL2
#undef ADD2
#undef SUB2

// One launch for the first `njobs` containers of the table (one geometry / pusher / cfl).
void launch_push_jobs(int pusher, const PushJobs& jobs, int njobs, unsigned max_n, double total_slots, const Geom& g, float cfl,
                      bool fuse) {
  ProfScope prof_(KC_PUSH, total_slots);
  if (!njobs || !max_n) return;
  if (njobs > PUSH_JOBS_MAX) throw Error(B2P_ERR_LOGIC, "launch_push_jobs: job table overflow");
  const int f = fuse ? (tuning().deposit_agg ? 2 : 1) : 0;
  const int agg_min = tuning().agg_min;
  const unsigned bx = (max_n + PUSH2_BLOCK_SLOTS - 1) / PUSH2_BLOCK_SLOTS;
  {
    const dim3 grid(bx, unsigned(njobs));
#define PUSH2_LAUNCH(P, F) k_push2<P, F><<<grid, PUSH2_THREADS, 0, ctx().stream>>>(jobs, g, cfl, 1.0f, agg_min)
#define PUSH2_CASE(P)                                                                              \
  case P:                                                                                          \
    if (f == 2) PUSH2_LAUNCH(P, 2); else if (f == 1) PUSH2_LAUNCH(P, 1); else PUSH2_LAUNCH(P, 0);  \
    break;
    switch (pusher) {
      PUSH2_CASE(B2P_PUSHER_BORIS)
      PUSH2_CASE(B2P_PUSHER_HIGUERA_CARY)
      PUSH2_CASE(B2P_PUSHER_FARADAY)
      default: throw Error(B2P_ERR_LOGIC, "pic::Tile::push_particles: unkown particle pusher");
    }
#undef PUSH2_CASE
#undef PUSH2_LAUNCH
    B2P_LAUNCH_CHECK();
  }
}

}  // namespace b2p
