// fields.cu — Yee-lattice kernels: FDTD2 / extended-stencil B push, E push,
// add_current, binomial current filter, halo fill, J exchange, field energy.
// All kernels are batched over a device-resident tile table (blockIdx.z spans
// tiles), read/write fp32 component-major lattices with k fastest, and use
// no FMA contraction (-fmad=false) so results are bit-identical to the
// reference's unfused CPU arithmetic.
#include "fields.cuh"

#include <algorithm>

namespace b2p {

// thread <-> cell mapping shared by the interior sweeps: x->k, y->j, z->(tile,i)
// grid = (k blocks, j blocks * planes, tiles): blockIdx.y carries (plane, j block)
#define INTERIOR_CELL_OR_RETURN() INTERIOR_CELL_OR_RETURN_AT(0)
#define INTERIOR_CELL_OR_RETURN_AT(TILE0)                                      \
  const int jblocks = (g.N[1] + int(blockDim.y) - 1) / int(blockDim.y);        \
  const int k = blockIdx.x * blockDim.x + threadIdx.x;                         \
  const int i = blockIdx.y / jblocks;                                          \
  const int j = (blockIdx.y - i * jblocks) * blockDim.y + threadIdx.y;         \
  const int tile = int(blockIdx.z) + (TILE0);                                  \
  if (k >= g.N[2] || j >= g.N[1]) return;                                      \
  const size_t sj = g.Hx[2], si = size_t(g.Hx[1]) * g.Hx[2];                   \
  const size_t n = (size_t(i + H) * g.Hx[1] + (j + H)) * g.Hx[2] + (k + H);    \
  const FieldPtrs f = tiles[tile];                                             \
  const size_t Ch = g.Ch;

// emf/yee_lattice_fdtd2.c++:43-57 — the reference's three per-component sweeps
// fused into one pass (each B component depends on E only, so fusing does not
// change any operand): 36 B/cell algorithmic.
// TWICE: two consecutive half pushes from the same E (projects/emf-wave/emf.py:50-51) in one pass — B1 = B + dt*curl E,
// B2 = B1 + dt*curl E, the same two roundings as two calls, for one read of E and one read + write of B.
template <bool TWICE>
__global__ void __launch_bounds__(256)
k_push_b_fdtd2(const FieldPtrs* __restrict__ tiles, const Geom g, const float dt) {
  INTERIOR_CELL_OR_RETURN();
  const float* __restrict__ Ex = f.E;
  const float* __restrict__ Ey = f.E + Ch;
  const float* __restrict__ Ez = f.E + 2 * Ch;
  const float ex = Ex[n], ey = Ey[n], ez = Ez[n];
  const float DkEy = Ey[n + 1] - ey;
  const float DjEz = Ez[n + sj] - ez;
  const float DiEz = Ez[n + si] - ez;
  const float DkEx = Ex[n + 1] - ex;
  const float DjEx = Ex[n + sj] - ex;
  const float DiEy = Ey[n + si] - ey;
  const float cx = dt * (DkEy - DjEz), cy = dt * (DiEz - DkEx), cz = dt * (DjEx - DiEy);
  float bx = f.B[n] + cx, by = f.B[Ch + n] + cy, bz = f.B[2 * Ch + n] + cz;
  if (TWICE) { bx = bx + cx; by = by + cy; bz = bz + cz; }
  f.B[n] = bx; f.B[Ch + n] = by; f.B[2 * Ch + n] = bz;
}

// emf/yee_lattice_fdtd2.c++:96-110, optionally followed by add_current
// (emf/yee_lattice.c++:176-178): E1 = E + dt*curl; E2 = E1 - J, same two roundings.
template <bool ADD_CURRENT>
__global__ void __launch_bounds__(256)
k_push_e_fdtd2(const FieldPtrs* __restrict__ tiles, const Geom g, const float dt) {
  INTERIOR_CELL_OR_RETURN();
  const float* __restrict__ Bx = f.B;
  const float* __restrict__ By = f.B + Ch;
  const float* __restrict__ Bz = f.B + 2 * Ch;
  const float bx = Bx[n], by = By[n], bz = Bz[n];
  const float DkBy = By[n - 1] - by;
  const float DjBz = Bz[n - sj] - bz;
  const float DiBz = Bz[n - si] - bz;
  const float DkBx = Bx[n - 1] - bx;
  const float DjBx = Bx[n - sj] - bx;
  const float DiBy = By[n - si] - by;
  float e0 = f.E[n] + dt * (DkBy - DjBz);
  float e1 = f.E[Ch + n] + dt * (DiBz - DkBx);
  float e2 = f.E[2 * Ch + n] + dt * (DjBx - DiBy);
  if (ADD_CURRENT) {
    e0 = e0 - f.J[n];
    e1 = e1 - f.J[Ch + n];
    e2 = e2 - f.J[2 * Ch + n];
  }
  f.E[n] = e0; f.E[Ch + n] = e1; f.E[2 * Ch + n] = e2;
}

// emf/yee_lattice.c++:171-179
__global__ void __launch_bounds__(256)
k_add_current(const FieldPtrs* __restrict__ tiles, const Geom g) {
  INTERIOR_CELL_OR_RETURN();
  f.E[n] = f.E[n] - f.J[n];
  f.E[Ch + n] = f.E[Ch + n] - f.J[Ch + n];
  f.E[2 * Ch + n] = f.E[2 * Ch + n] - f.J[2 * Ch + n];
  (void)sj; (void)si;
}

struct StencilM { float M[3][3][5]; };   // axis, row, col (emf/stencil_coefficients.h:30-64)

// One extended-stencil derivative, the 15-term sum of
// emf/yee_lattice_stencil.c++:55-85 in source order (left-associated).
__device__ __forceinline__ float stencil_deriv(const float* __restrict__ F, const long c, const float (&M)[3][5],
                                               const long sa, const long s1, const long s2) {
  float acc = 0.0f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const long hi = c + (r + 1) * sa, lo = c - r * sa;
    const float t0 = M[r][0] * (F[hi] - F[lo]);
    const float t1 = M[r][1] * ((F[hi + s1] + F[hi - s1]) - (F[lo + s1] + F[lo - s1]));
    const float t2 = M[r][2] * ((F[hi + s2] + F[hi - s2]) - (F[lo + s2] + F[lo - s2]));
    const float t3 = M[r][3] * ((F[hi + 2 * s1] + F[hi - 2 * s1]) - (F[lo + 2 * s1] + F[lo - 2 * s1]));
    const float t4 = M[r][4] * ((F[hi + 2 * s2] + F[hi - 2 * s2]) - (F[lo + 2 * s2] + F[lo - 2 * s2]));
    acc = (r == 0) ? t0 : acc + t0;
    acc = acc + t1; acc = acc + t2; acc = acc + t3; acc = acc + t4;
  }
  return acc;
}

// emf/yee_lattice_stencil.c++:18-295.  One B component per block (blockIdx.z = 3 * tile + component, like the
// reference's three per-component sweeps): two 54-point derivatives per thread instead of six keep the kernel at 64
// registers and twice the resident warps of the all-components version (107 registers, 23 % occupancy, issue and L1
// both under 50 %: it waited on its own loads).  Same operands and order => same bits.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_push_b_stencil(const FieldPtrs* __restrict__ tiles, const Geom g, const float dt, const StencilM c) {
  const int jblocks = (g.N[1] + int(blockDim.y) - 1) / int(blockDim.y);
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y / jblocks;
  const int j = (blockIdx.y - i * jblocks) * blockDim.y + threadIdx.y;
  const int tile = blockIdx.z / 3, comp = blockIdx.z - 3 * tile;
  if (k >= g.N[2] || j >= g.N[1]) return;
  const long sj = g.Hx[2], si = long(g.Hx[1]) * g.Hx[2];
  const long m = (long(i + H) * g.Hx[1] + (j + H)) * g.Hx[2] + (k + H);
  const FieldPtrs f = tiles[tile];
  const size_t Ch = g.Ch;
  const float* __restrict__ Ex = f.E;
  const float* __restrict__ Ey = f.E + Ch;
  const float* __restrict__ Ez = f.E + 2 * Ch;
  const long s[3] = { si, sj, 1 };
  // D*_a uses axis[a] with perp1=(a+1)%3, perp2=(a+2)%3
  float* __restrict__ B = f.B + size_t(comp) * Ch;
  if (comp == 0) {
    const float DzEy = stencil_deriv(Ey, m, c.M[2], s[2], s[0], s[1]);
    const float DyEz = stencil_deriv(Ez, m, c.M[1], s[1], s[2], s[0]);
    B[m] = B[m] + dt * (DzEy - DyEz);
  } else if (comp == 1) {
    const float DxEz = stencil_deriv(Ez, m, c.M[0], s[0], s[1], s[2]);
    const float DzEx = stencil_deriv(Ex, m, c.M[2], s[2], s[0], s[1]);
    B[m] = B[m] + dt * (DxEz - DzEx);
  } else {
    const float DyEx = stencil_deriv(Ex, m, c.M[1], s[1], s[2], s[0]);
    const float DxEy = stencil_deriv(Ey, m, c.M[0], s[0], s[1], s[2]);
    B[m] = B[m] + dt * (DyEx - DxEy);
  }
}

// ---- edge boundary condition (emf/yee_lattice.c++:263-306) ---------------------
// writes the masked components of one field over the box [lo, hi) of the haloed lattice
__global__ void __launch_bounds__(256)
k_edge_bc(float* __restrict__ f, const Geom g, const int3 lo, const int3 hi, const unsigned mask, const float3 v) {
  const size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t ny = size_t(hi.y - lo.y), nz = size_t(hi.z - lo.z);
  if (q >= size_t(hi.x - lo.x) * ny * nz) return;
  const size_t i = lo.x + q / (ny * nz), j = lo.y + (q / nz) % ny, k = lo.z + q % nz;
  const size_t n = (i * g.Hx[1] + j) * g.Hx[2] + k;
  if (mask & 1u) f[n] = v.x;
  if (mask & 2u) f[size_t(g.Ch) + n] = v.y;
  if (mask & 4u) f[2 * size_t(g.Ch) + n] = v.z;
}
// the same for a table of (lattice, box, mask, value) operations on DIFFERENT lattices: blockIdx.y = operation
__global__ void __launch_bounds__(256)
k_edge_bc_batch(const EdgeBcOp* __restrict__ ops, const Geom g) {
  const EdgeBcOp op = ops[blockIdx.y];
  const size_t ny = size_t(op.hi.y - op.lo.y), nz = size_t(op.hi.z - op.lo.z);
  const size_t total = size_t(op.hi.x - op.lo.x) * ny * nz;
  for (size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x; q < total; q += size_t(gridDim.x) * blockDim.x) {
    const size_t i = op.lo.x + q / (ny * nz), j = op.lo.y + (q / nz) % ny, k = op.lo.z + q % nz;
    const size_t n = (i * g.Hx[1] + j) * g.Hx[2] + k;
    if (op.mask & 1u) op.f[n] = op.v.x;
    if (op.mask & 2u) op.f[size_t(g.Ch) + n] = op.v.y;
    if (op.mask & 4u) op.f[2 * size_t(g.Ch) + n] = op.v.z;
  }
}
// J = J + add over whole haloed lattices (emf/yee_lattice.c++:361-375)
__global__ void __launch_bounds__(256)
k_add_lattice(float* __restrict__ J, const float* __restrict__ add, const size_t n) {
  const size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q < n) J[q] = J[q] + add[q];
}

// ---- antenna (emf/tile.c++:578-777) ---------------------------------------------
// vec_pot over the whole haloed lattice: sum over modes of A * Re(w * exp(i k.x)) at the Yee-staggered points;
// coordinates and the phase k.x in fp64 like the reference (global_coordinate_map, emf/tile.h:207-229), the
// phase then narrowed to fp32 for cosf / sinf.
__global__ void __launch_bounds__(256)
k_antenna_vecpot(float* __restrict__ vp, const Geom g, const AntennaModes m, const double3 mins, const double3 L) {
  const size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= size_t(g.Ch)) return;
  const int kk = int(q % g.Hx[2]), jj = int((q / g.Hx[2]) % g.Hx[1]), ii = int(q / (size_t(g.Hx[1]) * g.Hx[2]));
  const double i = double(ii) - H, j = double(jj) - H, k = double(kk) - H;
  auto gc = [&](const double a, const double b, const double c, double (&o)[3]) {
    o[0] = mins.x + (a / double(g.N[0])) * L.x;
    o[1] = mins.y + (b / double(g.N[1])) * L.y;
    o[2] = mins.z + (c / double(g.N[2])) * L.z;
  };
  double loc[3][3];
  gc(i + 0.5, j, k, loc[0]);
  gc(i, j + 0.5, k, loc[1]);
  gc(i, j, k + 0.5, loc[2]);
  float acc[3] = { 0.0f, 0.0f, 0.0f };
  for (int n = 0; n < m.n; ++n) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double dotv = 0.0;
      dotv = dotv + loc[c][0] * double(m.K[n][0]); dotv = dotv + loc[c][1] * double(m.K[n][1]); dotv = dotv + loc[c][2] * double(m.K[n][2]);
      const float phi = float(dotv);
      const float re = cosf(phi), im = sinf(phi);
      acc[c] = acc[c] + m.A[n][c] * (m.W[n][0] * re - m.W[n][1] * im);
    }
  }
  vp[q] = acc[0]; vp[size_t(g.Ch) + q] = acc[1]; vp[2 * size_t(g.Ch) + q] = acc[2];
}
// out = coeff * curl(X) with forward (DIR = +1) or backward (DIR = -1) neighbours over the box [lo, hi)
// (the `curl` lambda of emf/tile.c++:714-737)
template <int DIR>
__global__ void __launch_bounds__(256)
k_antenna_curl(float* __restrict__ out, const float* __restrict__ X, const Geom g, const int3 lo, const int3 hi, const float coeff) {
  const size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t ny = size_t(hi.y - lo.y), nz = size_t(hi.z - lo.z);
  if (q >= size_t(hi.x - lo.x) * ny * nz) return;
  const size_t i = lo.x + q / (ny * nz), j = lo.y + (q / nz) % ny, k = lo.z + q % nz;
  const long sj = g.Hx[2], si = long(g.Hx[1]) * g.Hx[2];
  const long n = long((i * g.Hx[1] + j) * g.Hx[2] + k);
  const size_t Ch = g.Ch;
  const float* X0 = X; const float* X1 = X + Ch; const float* X2 = X + 2 * Ch;
  const long di = DIR * si, dj = DIR * sj, dk = DIR;
  { const float Dk = X1[n + dk] - X1[n], Dj = X2[n + dj] - X2[n]; out[n] = coeff * (Dj - Dk); }
  { const float Di = X2[n + di] - X2[n], Dk = X0[n + dk] - X0[n]; out[Ch + n] = coeff * (Dk - Di); }
  { const float Dj = X0[n + dj] - X0[n], Di = X1[n + di] - X1[n]; out[2 * Ch + n] = coeff * (Di - Dj); }
}

// ---- field snapshot packing (io/snapshots/mpiio_fields.c++:221-275) -------------
// One thread per coarse cell of the tile: E, B sampled at interior index (ix,iy,iz)*stride, J summed
// over the stride^3 block in the reference's loop order (kk outermost), density slots zeroed.
// buf[f][iz][iy][ix], f = ex,ey,ez,bx,by,bz,jx,jy,jz,n0...
__global__ void __launch_bounds__(256)
k_pack_snapshot(const FieldPtrs f, const Geom g, const int stride, const int nxt, const int nyt, const int nzt, const int nf,
                float* __restrict__ buf) {
  const size_t o = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t te = size_t(nxt) * nyt * nzt;
  if (o >= te) return;
  const int ix = int(o % nxt), iy = int((o / nxt) % nyt), iz = int(o / (size_t(nxt) * nyt));
  const size_t Ch = g.Ch;
  auto lin = [&](const int i, const int j, const int k) { return (size_t(i + H) * g.Hx[1] + (j + H)) * g.Hx[2] + (k + H); };
  const size_t l = lin(ix * stride, iy * stride, iz * stride);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    buf[(0 + d) * te + o] = f.E[d * Ch + l];
    buf[(3 + d) * te + o] = f.B[d * Ch + l];
  }
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  for (int kk = 0; kk < stride; ++kk)
    for (int jj = 0; jj < stride; ++jj)
      for (int ii = 0; ii < stride; ++ii) {
        const size_t m = lin(ix * stride + ii, iy * stride + jj, iz * stride + kk);
        sx += f.J[m]; sy += f.J[Ch + m]; sz += f.J[2 * Ch + m];
      }
  buf[6 * te + o] = sx; buf[7 * te + o] = sy; buf[8 * te + o] = sz;
  for (int s = 9; s < nf; ++s) buf[s * te + o] = 0.0f;
}

// ---- binomial current filter --------------------------------------------------
// Both variants write every cell of dst: the region [1,H-1)^3 gets the filtered
// value, the outermost layer gets 0 (binomial2: the reference move-assigns a
// value-initialised grid, ..._binomial2.c++:43,75) or the old J (binomial2_unrolled
// filters in place, :139-156).
struct FilterTile { const float* src; float* dst; };

// One thread per (j,k) column marching along i with the three contributing planes kept in
// registers (27 values, or three separable partial sums): 9 loads per output instead of 27,
// every load coalesced along k.  grid = ((Hy*Hz)/256, i chunks, tiles*3 components).
template <bool UNROLLED>
__global__ void __launch_bounds__(256)
k_filter_binomial2(const FilterTile* __restrict__ tiles, const Geom g, const int chunk, const int ahead) {
  const int HyHz = g.Hx[1] * g.Hx[2];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= HyHz) return;
  const int j = q / g.Hx[2], k = q - j * g.Hx[2];
  const int tile = blockIdx.z / 3, c = blockIdx.z - 3 * tile;
  const int ibeg = blockIdx.y * chunk, iend = min(ibeg + chunk, g.Hx[0]);
  const float* __restrict__ J = tiles[tile].src + size_t(c) * g.Ch;
  float* __restrict__ out = tiles[tile].dst + size_t(c) * g.Ch;
  const int Hz = g.Hx[2];
  if (j == 0 || j == g.Hx[1] - 1 || k == 0 || k == Hz - 1) {      // outermost layer: 0 / unchanged
    for (int i = ibeg; i < iend; ++i) { const size_t n = size_t(i) * HyHz + q; out[n] = UNROLLED ? J[n] : 0.0f; }
    return;
  }
  int i = ibeg;
  if (i == 0) { out[q] = UNROLLED ? J[q] : 0.0f; i = 1; }
  if (UNROLLED) {
    // separable z, y, x passes (..._binomial2.c++:118-156): t2(plane) = y-pass of the z-pass;
    // same products and the same left-associated 3-term sums per pass
    auto t2_of = [&](const int ii) {
      const float* p = J + size_t(ii) * HyHz + q;
      float t1[3];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const float* r = p + (b - 1) * Hz;
        t1[b] = __fmaf_rn(0.25f, r[1], __fmaf_rn(0.5f, r[0], 0.25f * r[-1]));
      }
      return __fmaf_rn(0.25f, t1[2], __fmaf_rn(0.5f, t1[1], 0.25f * t1[0]));
    };
    float a0 = t2_of(i - 1), a1 = t2_of(i);
    for (; i < iend; ++i) {
      const size_t n = size_t(i) * HyHz + q;
      if (i == g.Hx[0] - 1) { out[n] = J[n]; break; }
      if (i + 1 + ahead < g.Hx[0]) asm volatile("prefetch.global.L2 [%0];" ::"l"(J + size_t(i + 1 + ahead) * HyHz + q));
      const float a2 = t2_of(i + 1);
      out[n] = __fmaf_rn(0.25f, a2, __fmaf_rn(0.5f, a1, 0.25f * a0));
      a0 = a1; a1 = a2;
    }
  } else {
    // 27-point sum accumulated from 0 in index_space order, a slowest (:58-73); the weights are powers of two, so the
    // fma below equals the reference's product-then-add bit for bit (see k_filter_binomial2_pairs)
    float P[3][9];
    // three row pointers (j-1, j, j+1) that advance one plane per step: the nine loads of a plane
    // then use immediate offsets -1, 0, +1 instead of nine 64-bit address computations
    const float* r0 = J + size_t(i - 1) * HyHz + q - Hz;
    const float* r1 = r0 + Hz;
    const float* r2 = r1 + Hz;
    auto load = [&](float (&dst)[9]) {
      dst[0] = r0[-1]; dst[1] = r0[0]; dst[2] = r0[1];
      dst[3] = r1[-1]; dst[4] = r1[0]; dst[5] = r1[1];
      dst[6] = r2[-1]; dst[7] = r2[0]; dst[8] = r2[1];
      r0 += HyHz; r1 += HyHz; r2 += HyHz;
    };
    load(P[0]);
    load(P[1]);
    float* o = out + size_t(i) * HyHz + q;
    const int last = g.Hx[0] - 1;
    // one output: planes A (i-1), B (i) are in registers, C receives plane i+1.  Called with the three
    // register planes in rotating roles, so no plane is ever copied.
    auto step = [&](const float (&A)[9], const float (&B)[9], float (&C)[9]) -> bool {
      if (i >= iend) return false;
      if (i == last) { *o = 0.0f; return false; }
      load(C);
      float acc = 0.0f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const float w = ((a == 1) ? 2.f : 1.f) * ((b == 1) ? 2.f : 1.f) * ((d == 1) ? 2.f : 1.f) / 64.f;
            acc = __fmaf_rn(w, a == 0 ? A[b * 3 + d] : a == 1 ? B[b * 3 + d] : C[b * 3 + d], acc);
          }
      *o = acc;
      ++i;
      o += HyHz;
      return true;
    };
    while (step(P[0], P[1], P[2]) && step(P[1], P[2], P[0]) && step(P[2], P[0], P[1])) {}
  }
}

// binomial2 (the 27-point variant) with TWO k-adjacent outputs per thread, k = 1 + 2 kp and k + 1 (lattices with an
// even Hz; the outermost layer is written by k_filter_binomial2_shell).  Every row of a plane arrives as two aligned
// LDG.64 (cells k-1..k+2), and the 27 terms of the pair are 27 packed FFMA2.
// Why an fma is allowed here although the parity contract forbids contraction: every weight is a power of two
// (1, 2, 4, 8 / 64), so the product w * x is exact and fma(w, x, acc) = rn(w * x + acc) = rn(rn(w * x) + acc), the
// reference's product-then-add, bit for bit (the only exception: a product that is subnormal, |x| < 2^-120).
// Same terms, same index_space order (a slowest) from 0 => bit-identical to the scalar kernel.
__global__ void __launch_bounds__(256)
k_filter_binomial2_pairs(const FilterTile* __restrict__ tiles, const Geom g, const int chunk, const int ahead) {
  const int Hz = g.Hx[2], HyHz = g.Hx[1] * Hz, npair = (Hz - 2) / 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (g.Hx[1] - 2) * npair) return;
  const int j = 1 + q / npair, k = 1 + 2 * (q - (j - 1) * npair);
  const int tile = blockIdx.z / 3, c = blockIdx.z - 3 * tile;
  const int ibeg = max(1, int(blockIdx.y) * chunk), iend = min(int(blockIdx.y + 1) * chunk, g.Hx[0] - 1);
  if (ibeg >= iend) return;
  const float* __restrict__ J = tiles[tile].src + size_t(c) * g.Ch;
  float* __restrict__ out = tiles[tile].dst + size_t(c) * g.Ch;
  // a plane: rows j-1, j, j+1, each as the pairs (k-1,k), (k,k+1), (k+1,k+2) = the operands of the two outputs for d = -1, 0, +1
  float2 P[3][9];
  const float* const jend = J + g.Ch;
  const float* r0 = J + size_t(ibeg - 1) * HyHz + size_t(j - 1) * Hz + (k - 1);     // k - 1 is even and Hz is even: 8-byte aligned
  auto load = [&](float2 (&dst)[9]) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const float2 lo = *reinterpret_cast<const float2*>(r0 + b * Hz), hi = *reinterpret_cast<const float2*>(r0 + b * Hz + 2);
      dst[b * 3 + 0] = lo;
      dst[b * 3 + 1] = make_float2(lo.y, hi.x);
      dst[b * 3 + 2] = hi;
    }
    // the own row of the plane `ahead` steps further is requested into L2 now (no register is held for it): the
    // kernel waits on load latency, not on bandwidth, and an L2 hit is a third of a DRAM round trip
    if (ahead > 0) {
      const float* pf = r0 + Hz + size_t(ahead) * HyHz;
      if (pf < jend) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
    }
    r0 += HyHz;
  };
  load(P[0]);
  load(P[1]);
  int i = ibeg;
  float* o = out + size_t(i) * HyHz + size_t(j) * Hz + k;
  auto step = [&](const float2 (&A)[9], const float2 (&B)[9], float2 (&C)[9]) -> bool {
    if (i >= iend) return false;
    load(C);
    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float w = ((a == 1) ? 2.f : 1.f) * ((b == 1) ? 2.f : 1.f) * ((d == 1) ? 2.f : 1.f) / 64.f;
          const float2 x = a == 0 ? A[b * 3 + d] : a == 1 ? B[b * 3 + d] : C[b * 3 + d];
          acc = __ffma2_rn(make_float2(w, w), x, acc);
        }
    o[0] = acc.x; o[1] = acc.y;
    ++i;
    o += HyHz;
    return true;
  };
  while (step(P[0], P[1], P[2]) && step(P[1], P[2], P[0]) && step(P[2], P[0], P[1])) {}
}
// the outermost layer of dst for the pair kernel: 0 (binomial2 move-assigns a value-initialised grid)
__global__ void __launch_bounds__(256)
k_filter_binomial2_shell(const FilterTile* __restrict__ tiles, const Geom g) {
  const int Hx = g.Hx[0], Hy = g.Hx[1], Hz = g.Hx[2];
  const int nA = 2 * Hy * Hz, nB = (Hx - 2) * 2 * Hz, nC = (Hx - 2) * (Hy - 2) * 2;
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nA + nB + nC) return;
  int i, j, k;
  if (q < nA) { i = q < Hy * Hz ? 0 : Hx - 1; q %= Hy * Hz; j = q / Hz; k = q - j * Hz; }
  else if (q < nA + nB) { q -= nA; i = 1 + q / (2 * Hz); q %= 2 * Hz; j = q < Hz ? 0 : Hy - 1; k = q % Hz; }
  else { q -= nA + nB; i = 1 + q / ((Hy - 2) * 2); q %= (Hy - 2) * 2; j = 1 + q / 2; k = (q & 1) ? Hz - 1 : 0; }
  const int tile = blockIdx.z / 3, c = blockIdx.z - 3 * tile;
  tiles[tile].dst[size_t(c) * g.Ch + (size_t(i) * Hy + j) * Hz + k] = 0.0f;
}

// ---- halo fill / J exchange (corgi local_communication, Moore order) ----------
// Per axis, region of direction d (emf/yee_lattice.h:493-517, 580-602):
//   subregion(d):               d=-1 [0,3)   d=0 [3,3+N)   d=+1 [3+N,6+N)
//   corresponding_subregion(d): d=-1 [N,N+3) d=0 [3,3+N)   d=+1 [3,6)
__device__ __forceinline__ int dir_of_halo(int a, int N) { return a < H ? -1 : (a >= H + N ? 1 : 0); }

// me.subregion(dir) <- neighbour(dir).corresponding_subregion(dir) for all 26
// directions of every tile (emf/yee_lattice.c++:206-239). One thread per lattice
// cell and component; interior cells exit. `which` selects E/B/J.
__global__ void __launch_bounds__(256)
k_halo_fill(const FieldPtrs* __restrict__ tiles, const int* __restrict__ nbr, const Geom g, const int which,
            const SlabDesc* __restrict__ remote, const int part /* 0: all halo cells; 1: those fed by a local tile; 2: by a remote rank */) {
  // Only halo cells get a thread.  They are enumerated as three groups (k fastest in each):
  //   A: the six full (j,k) planes with i in the halo        6 * Hy * Hz
  //   B: for interior i, the six halo rows in j              Nx * 6 * Hz
  //   C: for interior i and j, the six halo cells in k       Nx * Ny * 6
  const int Hy = g.Hx[1], Hz = g.Hx[2];
  const int nA = 2 * H * Hy * Hz, nB = g.N[0] * 2 * H * Hz, nC = g.N[0] * g.N[1] * 2 * H;
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int tile = blockIdx.z;
  if (q >= nA + nB + nC) return;
  int i, j, k;
  if (q < nA) {
    const int p = q / (Hy * Hz);
    q -= p * Hy * Hz;
    i = p < H ? p : g.N[0] + p;                     // 0,1,2, N+3,N+4,N+5
    j = q / Hz;
    k = q - j * Hz;
  } else if (q < nA + nB) {
    q -= nA;
    const int ii = q / (2 * H * Hz);
    q -= ii * 2 * H * Hz;
    const int p = q / Hz;
    i = ii + H;
    j = p < H ? p : g.N[1] + p;
    k = q - p * Hz;
  } else {
    q -= nA + nB;
    const int ii = q / (g.N[1] * 2 * H);
    q -= ii * g.N[1] * 2 * H;
    const int jj = q / (2 * H);
    const int p = q - jj * 2 * H;
    i = ii + H;
    j = jj + H;
    k = p < H ? p : g.N[2] + p;
  }
  const int di = dir_of_halo(i, g.N[0]), dj = dir_of_halo(j, g.N[1]), dk = dir_of_halo(k, g.N[2]);
  const int o = nbr[tile * 27 + ((di + 1) * 3 + (dj + 1)) * 3 + (dk + 1)];
  if (o == -1) return;
  if ((part == 1 && o < -1) || (part == 2 && o >= 0)) return;
  const size_t n = (size_t(i) * g.Hx[1] + j) * g.Hx[2] + k;
  if (o < -1) {
    // remote neighbour: its corresponding_subregion(d) was staged by the external exchange
    // (emf/yee_lattice.c++:431-472 reads the VirtualTile's hollow grid here)
    const SlabDesc s = remote[-(o + 2)];
    const int a[3] = { i, j, k }, dr[3] = { di, dj, dk };
    int r[3];
#pragma unroll
    for (int q2 = 0; q2 < 3; ++q2) r[q2] = dr[q2] == 0 ? a[q2] - H : (dr[q2] == 1 ? a[q2] - (H + g.N[q2]) : a[q2]);
    const size_t vol = size_t(s.dims[0]) * s.dims[1] * s.dims[2];
    const size_t m = (size_t(r[0]) * s.dims[1] + r[1]) * s.dims[2] + r[2];
    float* dstr = which == 0 ? tiles[tile].E : (which == 1 ? tiles[tile].B : tiles[tile].J);
#pragma unroll
    for (int c = 0; c < 3; ++c) dstr[size_t(c) * g.Ch + n] = s.base[size_t(c) * vol + m];
    return;
  }
  // my halo index a in subregion(d) maps to a - d*N in the neighbour (same formula for d=-1,0,+1)
  const int si_ = i - di * g.N[0], sj_ = j - dj * g.N[1], sk_ = k - dk * g.N[2];
  const size_t m = (size_t(si_) * g.Hx[1] + sj_) * g.Hx[2] + sk_;
  const FieldPtrs me = tiles[tile], ot = tiles[o];
  float* dst = which == 0 ? me.E : (which == 1 ? me.B : me.J);
  const float* src = which == 0 ? ot.E : (which == 1 ? ot.B : ot.J);
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[size_t(c) * g.Ch + n] = src[size_t(c) * g.Ch + m];
}

// me.corresponding_subregion(-dir) += neighbour(dir).subregion(-dir) for the 26
// directions in Moore order kr->jr->ir (emf/yee_lattice.c++:249-261,
// corgi cellular_automata.h:48-62), so every cell accumulates its up-to-7
// contributions in the reference's order. One thread per interior cell/component.
__global__ void __launch_bounds__(256)
k_J_exchange(const FieldPtrs* __restrict__ tiles, const int* __restrict__ nbr, const Geom g,
             const SlabDesc* __restrict__ remote, const int tile0 /* the launch covers tiles [tile0, tile0 + gridDim.z) */) {
  INTERIOR_CELL_OR_RETURN_AT(tile0);
  (void)sj; (void)si;
  const int a[3] = { i + H, j + H, k + H };
  // along one axis, cell a (haloed index) lies in corresponding_subregion(-d) for
  //   d=+1: a in [N, N+3)  -> neighbour cell a - N  (its lower halo)
  //   d=-1: a in [3, 6)    -> neighbour cell a + N  (its upper halo)
  //   d= 0: always         -> neighbour cell a
  bool lo[3], hi[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { lo[d] = a[d] < 2 * H; hi[d] = a[d] >= g.N[d]; }
  if (!(lo[0] | hi[0] | lo[1] | hi[1] | lo[2] | hi[2])) return;
  float acc[3] = { f.J[n], f.J[Ch + n], f.J[2 * Ch + n] };
  // Moore order (kr slowest), restricted per axis to the directions this cell can receive from:
  // -1 only in the lower band, +1 only in the upper band — the other directions of the 26 contribute
  // nothing, and skipping them up front leaves a face cell with one iteration instead of 26
  for (int kr = lo[2] ? -1 : 0; kr <= (hi[2] ? 1 : 0); ++kr)
    for (int jr = lo[1] ? -1 : 0; jr <= (hi[1] ? 1 : 0); ++jr)
      for (int ir = lo[0] ? -1 : 0; ir <= (hi[0] ? 1 : 0); ++ir) {
        if (ir == 0 && jr == 0 && kr == 0) continue;
        const int dr[3] = { ir, jr, kr };
        int s[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) s[d] = a[d] - dr[d] * g.N[d];
        const int o = nbr[tile * 27 + ((ir + 1) * 3 + (jr + 1)) * 3 + (kr + 1)];
        if (o == -1) continue;
        if (o < -1) {
          // remote neighbour: its halo subregion(-dir) was staged by the external exchange
          // (emf/yee_lattice.c++:498-524)
          const SlabDesc sl = remote[-(o + 2)];
          int r[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) r[d] = dr[d] == 1 ? a[d] - g.N[d] : a[d] - H;
          const size_t vol = size_t(sl.dims[0]) * sl.dims[1] * sl.dims[2];
          const size_t mm = (size_t(r[0]) * sl.dims[1] + r[1]) * sl.dims[2] + r[2];
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c] = acc[c] + sl.base[size_t(c) * vol + mm];
          continue;
        }
        const float* oJ = tiles[o].J;
        const size_t m = (size_t(s[0]) * g.Hx[1] + s[1]) * g.Hx[2] + s[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = acc[c] + oJ[size_t(c) * Ch + m];
      }
  f.J[n] = acc[0]; f.J[Ch + n] = acc[1]; f.J[2 * Ch + n] = acc[2];
}

// emf/yee_lattice.c++:383-428 — Σ(Fx²+Fy²+Fz²) over the interior, accumulated in
// double (the reference's serial fp32 sum is reproduced only to tolerance).
__global__ void __launch_bounds__(256)
k_field_energy(const FieldPtrs* __restrict__ tiles, const Geom g, double* __restrict__ out /*[ntiles][2]*/) {
  B2P_GLOBAL(tiles[blockIdx.y].E); B2P_GLOBAL(tiles[blockIdx.y].B);
  const int tile = blockIdx.y;
  const size_t Ni = size_t(g.N[0]) * g.N[1] * g.N[2];
  double sB = 0, sE = 0;
  const FieldPtrs f = tiles[tile];
  for (size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x; q < Ni; q += size_t(gridDim.x) * blockDim.x) {
    const int k = q % g.N[2];
    const int j = (q / g.N[2]) % g.N[1];
    const int i = q / (size_t(g.N[2]) * g.N[1]);
    const size_t n = (size_t(i + H) * g.Hx[1] + (j + H)) * g.Hx[2] + (k + H);
    const float bx = f.B[n], by = f.B[g.Ch + n], bz = f.B[2 * size_t(g.Ch) + n];
    const float ex = f.E[n], ey = f.E[g.Ch + n], ez = f.E[2 * size_t(g.Ch) + n];
    sB += double(bx * bx + by * by + bz * bz);
    sE += double(ex * ex + ey * ey + ez * ez);
  }
  for (int o = 16; o > 0; o >>= 1) { sB += __shfl_xor_sync(0xffffffffu, sB, o); sE += __shfl_xor_sync(0xffffffffu, sE, o); }
  __shared__ double shB[8], shE[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { shB[w] = sB; shE[w] = sE; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0, e = 0;
    for (int q = 0; q < int(blockDim.x >> 5); ++q) { b += shB[q]; e += shE[q]; }
    atomicAdd(&out[2 * tile], b);
    atomicAdd(&out[2 * tile + 1], e);
  }
}

// ------------------------------------------------------------------ launchers --
static dim3 cell_block() { return dim3(32, 8, 1); }
// grid.z (<= 65535) spans tiles
constexpr int MAX_TILES_PER_LAUNCH = 65535;
static dim3 interior_grid(const Geom& g, int ntiles) {
  return dim3((g.N[2] + 31) / 32, unsigned((g.N[1] + 7) / 8) * g.N[0], unsigned(ntiles));
}
static void check_tiles(int ntiles) {
  if (ntiles > MAX_TILES_PER_LAUNCH) throw Error(B2P_ERR_RUNTIME, "more than 65535 local tiles per GPU are not supported");
}

void launch_push_b_fdtd2(const FieldPtrs* tiles, int ntiles, const Geom& g, float dt, bool twice) {
  ProfScope prof_(KC_PUSH_B, double(ntiles) * g.N[0] * g.N[1] * g.N[2] * (twice ? 2.0 : 1.0));
  if (!ntiles) return;
  check_tiles(ntiles);
  if (twice) k_push_b_fdtd2<true><<<interior_grid(g, ntiles), cell_block(), 0, ctx().stream>>>(tiles, g, dt);
  else k_push_b_fdtd2<false><<<interior_grid(g, ntiles), cell_block(), 0, ctx().stream>>>(tiles, g, dt);
  B2P_LAUNCH_CHECK();
}
void launch_push_b_stencil(const FieldPtrs* tiles, int ntiles, const Geom& g, float dt, const float M[3][3][5]) {
  ProfScope prof_(KC_PUSH_B, double(ntiles) * g.N[0] * g.N[1] * g.N[2]);
  if (!ntiles) return;
  check_tiles(ntiles);
  StencilM c;
  for (int a = 0; a < 3; ++a) for (int r = 0; r < 3; ++r) for (int q = 0; q < 5; ++q) c.M[a][r][q] = M[a][r][q];
  if (ntiles * 3 > MAX_TILES_PER_LAUNCH) throw Error(B2P_ERR_RUNTIME, "more than 21845 local tiles per GPU are not supported");
  dim3 grid = interior_grid(g, ntiles);
  grid.z *= 3;
  switch (tuning().stencil_minb) {
    case 2: k_push_b_stencil<2><<<grid, cell_block(), 0, ctx().stream>>>(tiles, g, dt, c); break;
    case 3: k_push_b_stencil<3><<<grid, cell_block(), 0, ctx().stream>>>(tiles, g, dt, c); break;
    case 4: k_push_b_stencil<4><<<grid, cell_block(), 0, ctx().stream>>>(tiles, g, dt, c); break;
    default: k_push_b_stencil<5><<<grid, cell_block(), 0, ctx().stream>>>(tiles, g, dt, c); break;
  }
  B2P_LAUNCH_CHECK();
}
void launch_push_e_fdtd2(const FieldPtrs* tiles, int ntiles, const Geom& g, float dt, bool add_current) {
  ProfScope prof_(KC_PUSH_E, double(ntiles) * g.N[0] * g.N[1] * g.N[2]);
  if (!ntiles) return;
  check_tiles(ntiles);
  if (add_current) k_push_e_fdtd2<true><<<interior_grid(g, ntiles), cell_block(), 0, ctx().stream>>>(tiles, g, dt);
  else k_push_e_fdtd2<false><<<interior_grid(g, ntiles), cell_block(), 0, ctx().stream>>>(tiles, g, dt);
  B2P_LAUNCH_CHECK();
}
void launch_add_current(const FieldPtrs* tiles, int ntiles, const Geom& g) {
  ProfScope prof_(KC_ADD_CURRENT, double(ntiles) * g.N[0] * g.N[1] * g.N[2]);
  if (!ntiles) return;
  check_tiles(ntiles);
  k_add_current<<<interior_grid(g, ntiles), cell_block(), 0, ctx().stream>>>(tiles, g);
  B2P_LAUNCH_CHECK();
}
void launch_filter(const void* filter_tiles, int ntiles, const Geom& g, bool unrolled) {
  ProfScope prof_(KC_FILTER, double(ntiles) * g.Ch);
  if (!ntiles) return;
  check_tiles(ntiles);
  if (ntiles * 3 > MAX_TILES_PER_LAUNCH) throw Error(B2P_ERR_RUNTIME, "more than 21845 local tiles per GPU are not supported");
  const int chunk = tuning().filter_chunk > 0 ? tuning().filter_chunk : 35;
  const dim3 grid((g.Hx[1] * g.Hx[2] + 255) / 256, (g.Hx[0] + chunk - 1) / chunk, unsigned(ntiles) * 3);
  const FilterTile* ft = static_cast<const FilterTile*>(filter_tiles);
  if (unrolled) k_filter_binomial2<true><<<grid, 256, 0, ctx().stream>>>(ft, g, chunk, tuning().filter_ahead > 0 ? tuning().filter_ahead : 1 << 20);
  else if ((g.Hx[2] & 1) == 0 && tuning().filter_pairs) {
    // even Hz: two outputs per thread on the packed fp32x2 pipe + the zeroed outermost layer
    const int npair = (g.Hx[1] - 2) * ((g.Hx[2] - 2) / 2);
    k_filter_binomial2_pairs<<<dim3((npair + 255) / 256, grid.y, grid.z), 256, 0, ctx().stream>>>(ft, g, chunk, tuning().filter_ahead);
    B2P_LAUNCH_CHECK();
    const int nshell = 2 * g.Hx[1] * g.Hx[2] + (g.Hx[0] - 2) * 2 * g.Hx[2] + (g.Hx[0] - 2) * (g.Hx[1] - 2) * 2;
    k_filter_binomial2_shell<<<dim3((nshell + 255) / 256, 1, grid.z), 256, 0, ctx().stream>>>(ft, g);
  } else k_filter_binomial2<false><<<grid, 256, 0, ctx().stream>>>(ft, g, chunk, 0);
  B2P_LAUNCH_CHECK();
}
void launch_edge_bc(float* field, const Geom& g, const int lo[3], const int hi[3], unsigned mask, const float v[3]) {
  ProfScope prof_(KC_OTHER, 0.0);
  const size_t total = size_t(hi[0] - lo[0]) * size_t(hi[1] - lo[1]) * size_t(hi[2] - lo[2]);
  if (!total || !(mask & 7u)) return;
  k_edge_bc<<<unsigned((total + 255) / 256), 256, 0, ctx().stream>>>(field, g, make_int3(lo[0], lo[1], lo[2]),
                                                                    make_int3(hi[0], hi[1], hi[2]), mask,
                                                                    make_float3(v[0], v[1], v[2]));
  B2P_LAUNCH_CHECK();
}
void launch_edge_bc_batch(const EdgeBcOp* ops, int nops, size_t max_cells, const Geom& g) {
  ProfScope prof_(KC_OTHER, 0.0);
  if (!nops || !max_cells) return;
  const unsigned bx = unsigned(std::min<size_t>((max_cells + 255) / 256, 1024));
  for (int o = 0; o < nops; o += 65535) {
    k_edge_bc_batch<<<dim3(bx, unsigned(std::min(65535, nops - o))), 256, 0, ctx().stream>>>(ops + o, g);
    B2P_LAUNCH_CHECK();
  }
}
void launch_add_lattice(float* J, const float* add, size_t n) {
  ProfScope prof_(KC_ADD_CURRENT, double(n));
  if (!n) return;
  k_add_lattice<<<unsigned((n + 255) / 256), 256, 0, ctx().stream>>>(J, add, n);
  B2P_LAUNCH_CHECK();
}
void launch_antenna(float* J, float* vec_pot, float* gen_B, const Geom& g, const AntennaModes& m, const double mins[3],
                    const double maxs[3], float cfl_neg) {
  ProfScope prof_(KC_OTHER, 0.0);
  const double3 mn = make_double3(mins[0], mins[1], mins[2]);
  const double3 L = make_double3(maxs[0] - mins[0], maxs[1] - mins[1], maxs[2] - mins[2]);
  k_antenna_vecpot<<<unsigned((g.Ch + 255) / 256), 256, 0, ctx().stream>>>(vec_pot, g, m, mn, L);
  B2P_LAUNCH_CHECK();
  const int3 lo1 = make_int3(H - 1, H - 1, H - 1), hi1 = make_int3(H + g.N[0] + 1, H + g.N[1] + 1, H + g.N[2] + 1);
  const size_t n1 = size_t(g.N[0] + 2) * (g.N[1] + 2) * (g.N[2] + 2);
  k_antenna_curl<+1><<<unsigned((n1 + 255) / 256), 256, 0, ctx().stream>>>(gen_B, vec_pot, g, lo1, hi1, 1.0f);
  B2P_LAUNCH_CHECK();
  const int3 lo2 = make_int3(H, H, H), hi2 = make_int3(H + g.N[0], H + g.N[1], H + g.N[2]);
  const size_t n2 = size_t(g.N[0]) * g.N[1] * g.N[2];
  k_antenna_curl<-1><<<unsigned((n2 + 255) / 256), 256, 0, ctx().stream>>>(vec_pot, gen_B, g, lo2, hi2, cfl_neg);
  B2P_LAUNCH_CHECK();
  k_add_lattice<<<unsigned((size_t(3) * g.Ch + 255) / 256), 256, 0, ctx().stream>>>(J, vec_pot, size_t(3) * g.Ch);
  B2P_LAUNCH_CHECK();
}
void launch_pack_snapshot(const FieldPtrs& f, const Geom& g, int stride, int nxt, int nyt, int nzt, int nf, float* buf) {
  ProfScope prof_(KC_OTHER, 0.0);
  const size_t te = size_t(nxt) * nyt * nzt;
  k_pack_snapshot<<<unsigned((te + 255) / 256), 256, 0, ctx().stream>>>(f, g, stride, nxt, nyt, nzt, nf, buf);
  B2P_LAUNCH_CHECK();
}
void launch_zero(float* p, size_t n) {
  ProfScope prof_(KC_ZERO, double(n));
  if (!n) return;
  B2P_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), ctx().stream));   // +0.0f is all-zero bits
  count_launch();
}
void launch_halo_fill(const FieldPtrs* tiles, const int* nbr, int ntiles, const Geom& g, int which, const SlabDesc* remote, int part) {
  ProfScope prof_(KC_HALO, double(ntiles) * g.Ch);
  if (!ntiles) return;
  check_tiles(ntiles);
  const int nhalo = 2 * H * g.Hx[1] * g.Hx[2] + g.N[0] * 2 * H * g.Hx[2] + g.N[0] * g.N[1] * 2 * H;
  k_halo_fill<<<dim3((nhalo + 255) / 256, 1, unsigned(ntiles)), 256, 0, ctx().stream>>>(tiles, nbr, g, which, remote, part);
  B2P_LAUNCH_CHECK();
}
void launch_J_exchange(const FieldPtrs* tiles, const int* nbr, int ntiles, const Geom& g, const SlabDesc* remote, int tile0) {
  ProfScope prof_(KC_J_EXCHANGE, double(ntiles) * g.N[0] * g.N[1] * g.N[2]);
  if (!ntiles) return;
  check_tiles(ntiles);
  k_J_exchange<<<interior_grid(g, ntiles), cell_block(), 0, ctx().stream>>>(tiles, nbr, g, remote, tile0);
  B2P_LAUNCH_CHECK();
}
void launch_field_energy(const FieldPtrs* tiles, int ntiles, const Geom& g, double* out) {
  ProfScope prof_(KC_ENERGY, double(ntiles) * g.N[0] * g.N[1] * g.N[2]);
  if (!ntiles) return;
  check_tiles(ntiles);
  B2P_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 2 * ntiles, ctx().stream));
  const size_t Ni = size_t(g.N[0]) * g.N[1] * g.N[2];
  const unsigned bx = unsigned(std::min<size_t>((Ni + 255) / 256, 64));
  k_field_energy<<<dim3(bx, ntiles), 256, 0, ctx().stream>>>(tiles, g, out);
  B2P_LAUNCH_CHECK();
}

}  // namespace b2p
