// pic_oracle.cpp — CPU oracle for the PIC-step hot path.  TEST INFRASTRUCTURE ONLY
// (see pic_oracle.h).  Each function restates one reference function with the
// reference's fp32 operation order; citations are `file:line` under
// /root/reference/src/runko unless another root is given.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
// -ffp-contract=off is part of the contract: no a*b+c fusion anywhere, so that
// IEEE add/mul/div/sqrt give the same bits as the CUDA kernels built -fmad=false.
#include "pic_oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

thread_local std::string g_err;
constexpr int H = B2P_HALO;                       // emf/common.h:9
constexpr uint64_t DEAD = ~uint64_t(0);           // particles_common.h:23

struct Span { const b2p_particle_state* p; size_t n; };

struct Container {                                 // pic/particle.h:55-90
  std::vector<float> x, y, z, ux, uy, uz;
  std::vector<uint64_t> id;
  double charge = 0, mass = 1;
  size_t size() const { return x.size(); }
  void resize(size_t n) {
    x.resize(n); y.resize(n); z.resize(n); ux.resize(n); uy.resize(n); uz.resize(n); id.resize(n);
  }
};

struct Tile {
  int idx[3];
  double mins[3], maxs[3];                         // corgi::Tile::mins/maxs (emf/tile.c++:162-170)
  int N[3], Hx[3];
  size_t Ch;
  std::vector<float> E, B, J;                      // 3*Ch each, buf[c*Ch + lin(i,j,k)]
  std::vector<float> genJ;                         // generated_J_cache_ (pic/tile.c++:386-389)
  std::vector<float> t1, t2;                       // unrolled-filter cache
  std::vector<Container> sp;
  uint64_t tile_tag = 0;
  std::vector<uint64_t> next_ordinal;
  std::vector<b2p_particle_state> out_buf;         // subregion_particle_buff_
  std::vector<size_t> out_ends;                    // subregion_particle_ends_
  std::vector<std::vector<Span>> incoming;         // incoming_subregion_particles_
  struct Antenna { double A[3], wave[3]; int kind; bool has_coeffs; std::vector<std::array<double, 2>> coeffs; size_t next = 0; };
  std::vector<Antenna> antennas;                   // emf/tile.h:48
  std::vector<float> vec_pot, gen_B;               // vec_pot_buff_, generated_B_buff_ (emf/tile.h:68-69)
  std::vector<b2p_edge_bc> edge_bcs;               // emf/tile.h:51
  std::vector<b2p_reflector_wall> walls;           // pic/tile.h:71
  std::vector<float> corrJ;                        // reflector_correction_J_ (pic/tile.h:72)
  bool corr_pending = false;                       // reflector_correction_pending_ (pic/tile.h:73)
  size_t lin(size_t i, size_t j, size_t k) const { return (i * Hx[1] + j) * Hx[2] + k; }
};

}  // namespace

struct orc_grid {
  b2p_config cfg;
  float stencilM[3][3][5];
  std::vector<Tile> tiles;                         // index = cid
  double gmins[3], gmaxs[3];
  int threads = 1;   // workers used by orc_local_communication (set by the step drivers)
};

namespace {

// emf/stencil_coefficients.h:43-64
float stencil_alpha(const float M[3][5]) {
  float sum = 0.0f;
  sum += 3.0f * (M[1][0] + 2.0f * (M[1][1] + M[1][2] + M[1][3] + M[1][4]));
  sum += 5.0f * (M[2][0] + 2.0f * (M[2][1] + M[2][2] + M[2][3] + M[2][4]));
  sum += 1.0f * 2.0f * (M[0][1] + M[0][2] + M[0][3] + M[0][4]);
  return 1.0f - sum;
}

// ---------------------------------------------------------------- fields --
// emf/yee_lattice_fdtd2.c++:20-58
void push_b_fdtd2(Tile& t, float dt) {
  const size_t Ch = t.Ch;
  float* Bx = &t.B[0]; float* By = &t.B[Ch]; float* Bz = &t.B[2 * Ch];
  const float* Ex = &t.E[0]; const float* Ey = &t.E[Ch]; const float* Ez = &t.E[2 * Ch];
  const size_t sj = t.Hx[2], si = size_t(t.Hx[1]) * t.Hx[2];
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DkEy = Ey[n + 1] - Ey[n];
        const float DjEz = Ez[n + sj] - Ez[n];
        Bx[n] = Bx[n] + dt * (DkEy - DjEz);
      }
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DiEz = Ez[n + si] - Ez[n];
        const float DkEx = Ex[n + 1] - Ex[n];
        By[n] = By[n] + dt * (DiEz - DkEx);
      }
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DjEx = Ex[n + sj] - Ex[n];
        const float DiEy = Ey[n + si] - Ey[n];
        Bz[n] = Bz[n] + dt * (DjEx - DiEy);
      }
}

// emf/yee_lattice_fdtd2.c++:69-111
void push_e_fdtd2(Tile& t, float dt) {
  const size_t Ch = t.Ch;
  float* Ex = &t.E[0]; float* Ey = &t.E[Ch]; float* Ez = &t.E[2 * Ch];
  const float* Bx = &t.B[0]; const float* By = &t.B[Ch]; const float* Bz = &t.B[2 * Ch];
  const size_t sj = t.Hx[2], si = size_t(t.Hx[1]) * t.Hx[2];
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DkBy = By[n - 1] - By[n];
        const float DjBz = Bz[n - sj] - Bz[n];
        Ex[n] = Ex[n] + dt * (DkBy - DjBz);
      }
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DiBz = Bz[n - si] - Bz[n];
        const float DkBx = Bx[n - 1] - Bx[n];
        Ey[n] = Ey[n] + dt * (DiBz - DkBx);
      }
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DjBx = Bx[n - sj] - Bx[n];
        const float DiBy = By[n - si] - By[n];
        Ez[n] = Ez[n] + dt * (DjBx - DiBy);
      }
}

// One extended-stencil derivative D*_a F at lattice point n: the 15-term sum of
// emf/yee_lattice_stencil.c++:55-85 (and its five siblings), in source order.
// sa = stride of the axial direction, s1/s2 = strides of perp1=(a+1)%3, perp2=(a+2)%3.
float stencil_deriv(const float* F, size_t n, const float M[3][5],
                    ptrdiff_t sa, ptrdiff_t s1, ptrdiff_t s2) {
  const ptrdiff_t c = ptrdiff_t(n);
  float acc = 0.0f;
  for (int r = 0; r < 3; ++r) {
    const ptrdiff_t hi = c + (r + 1) * sa;   // a+1, a+2, a+3
    const ptrdiff_t lo = c - r * sa;         // a,   a-1, a-2
    const float t0 = M[r][0] * (F[hi] - F[lo]);
    const float t1 = M[r][1] * ((F[hi + s1] + F[hi - s1]) - (F[lo + s1] + F[lo - s1]));
    const float t2 = M[r][2] * ((F[hi + s2] + F[hi - s2]) - (F[lo + s2] + F[lo - s2]));
    const float t3 = M[r][3] * ((F[hi + 2 * s1] + F[hi - 2 * s1]) - (F[lo + 2 * s1] + F[lo - 2 * s1]));
    const float t4 = M[r][4] * ((F[hi + 2 * s2] + F[hi - 2 * s2]) - (F[lo + 2 * s2] + F[lo - 2 * s2]));
    if (r == 0) acc = t0; else acc = acc + t0;   // the source has no leading "0 +"
    acc = acc + t1;
    acc = acc + t2;
    acc = acc + t3;
    acc = acc + t4;
  }
  return acc;
}

// emf/yee_lattice_stencil.c++:18-295
void push_b_stencil(Tile& t, float dt, const float M[3][3][5]) {
  const size_t Ch = t.Ch;
  float* Bx = &t.B[0]; float* By = &t.B[Ch]; float* Bz = &t.B[2 * Ch];
  const float* Ex = &t.E[0]; const float* Ey = &t.E[Ch]; const float* Ez = &t.E[2 * Ch];
  const ptrdiff_t s[3] = { ptrdiff_t(t.Hx[1]) * t.Hx[2], ptrdiff_t(t.Hx[2]), 1 };
  auto D = [&](const float* F, size_t n, int a) {
    return stencil_deriv(F, n, M[a], s[a], s[(a + 1) % 3], s[(a + 2) % 3]);
  };
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DzEy = D(Ey, n, 2), DyEz = D(Ez, n, 1);
        Bx[n] = Bx[n] + dt * (DzEy - DyEz);
      }
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DxEz = D(Ez, n, 0), DzEx = D(Ex, n, 2);
        By[n] = By[n] + dt * (DxEz - DzEx);
      }
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float DyEx = D(Ex, n, 1), DxEy = D(Ey, n, 0);
        Bz[n] = Bz[n] + dt * (DyEx - DxEy);
      }
}

// emf/yee_lattice.c++:171-179
void add_current(Tile& t) {
  for (int c = 0; c < 3; ++c)
    for (int i = H; i < H + t.N[0]; ++i)
      for (int j = H; j < H + t.N[1]; ++j)
        for (int k = H; k < H + t.N[2]; ++k) {
          const size_t n = c * t.Ch + t.lin(i, j, k);
          t.E[n] = t.E[n] - t.J[n];
        }
}

// emf/yee_lattice_current_filter_binomial2.c++:23-76.  The result replaces J_
// with a freshly value-initialised grid, so the outermost layer becomes 0.
void filter_binomial2(Tile& t) {
  static const float C3[3][3][3] = {
    { { 1.f / 64.f, 2.f / 64.f, 1.f / 64.f }, { 2.f / 64.f, 4.f / 64.f, 2.f / 64.f }, { 1.f / 64.f, 2.f / 64.f, 1.f / 64.f } },
    { { 2.f / 64.f, 4.f / 64.f, 2.f / 64.f }, { 4.f / 64.f, 8.f / 64.f, 4.f / 64.f }, { 2.f / 64.f, 4.f / 64.f, 2.f / 64.f } },
    { { 1.f / 64.f, 2.f / 64.f, 1.f / 64.f }, { 2.f / 64.f, 4.f / 64.f, 2.f / 64.f }, { 1.f / 64.f, 2.f / 64.f, 1.f / 64.f } } };
  std::vector<float> out(3 * t.Ch, 0.0f);
  for (int c = 0; c < 3; ++c) {
    const float* Jc = &t.J[c * t.Ch];
    float* oc = &out[c * t.Ch];
    for (int i = 1; i < t.Hx[0] - 1; ++i)
      for (int j = 1; j < t.Hx[1] - 1; ++j)
        for (int k = 1; k < t.Hx[2] - 1; ++k) {
          float acc = 0.0f;
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
              for (int d = 0; d < 3; ++d)
                acc = acc + C3[a][b][d] * Jc[t.lin(i - 1 + a, j - 1 + b, k - 1 + d)];
          oc[t.lin(i, j, k)] = acc;
        }
  }
  t.J.swap(out);
}

// emf/yee_lattice_current_filter_binomial2.c++:88-157 (in place; outer layer untouched)
void filter_binomial2_unrolled(Tile& t) {
  const float B1[3] = { 0.25f, 0.5f, 0.25f };
  const int Nx = t.Hx[0], Ny = t.Hx[1], Nz = t.Hx[2];
  const size_t n1 = size_t(Nx) * Ny * (Nz - 2), n2 = size_t(Nx) * (Ny - 2) * (Nz - 2);
  t.t1.resize(3 * n1); t.t2.resize(3 * n2);
  for (int c = 0; c < 3; ++c) {
    const float* Jc = &t.J[c * t.Ch];
    float* a1 = &t.t1[c * n1];
    for (int i = 0; i < Nx; ++i) for (int j = 0; j < Ny; ++j) for (int k = 0; k < Nz - 2; ++k)
      a1[(size_t(i) * Ny + j) * (Nz - 2) + k] =
        B1[0] * Jc[t.lin(i, j, k)] + B1[1] * Jc[t.lin(i, j, k + 1)] + B1[2] * Jc[t.lin(i, j, k + 2)];
    float* a2 = &t.t2[c * n2];
    for (int i = 0; i < Nx; ++i) for (int j = 0; j < Ny - 2; ++j) for (int k = 0; k < Nz - 2; ++k)
      a2[(size_t(i) * (Ny - 2) + j) * (Nz - 2) + k] =
        B1[0] * a1[(size_t(i) * Ny + j) * (Nz - 2) + k] + B1[1] * a1[(size_t(i) * Ny + j + 1) * (Nz - 2) + k] +
        B1[2] * a1[(size_t(i) * Ny + j + 2) * (Nz - 2) + k];
    float* Jw = &t.J[c * t.Ch];
    for (int i = 0; i < Nx - 2; ++i) for (int j = 0; j < Ny - 2; ++j) for (int k = 0; k < Nz - 2; ++k)
      Jw[t.lin(i + 1, j + 1, k + 1)] =
        B1[0] * a2[(size_t(i) * (Ny - 2) + j) * (Nz - 2) + k] + B1[1] * a2[(size_t(i + 1) * (Ny - 2) + j) * (Nz - 2) + k] +
        B1[2] * a2[(size_t(i + 2) * (Ny - 2) + j) * (Nz - 2) + k];
  }
}

// emf/yee_lattice.c++:383-428: thrust::reduce of fp32 values (sequential on the
// CPP device system), then / 2.0 in double.
double field_energy(const Tile& t, const std::vector<float>& F) {
  float acc = 0.0f;
  for (int i = H; i < H + t.N[0]; ++i)
    for (int j = H; j < H + t.N[1]; ++j)
      for (int k = H; k < H + t.N[2]; ++k) {
        const size_t n = t.lin(i, j, k);
        const float a = F[n], b = F[t.Ch + n], c = F[2 * t.Ch + n];
        acc = acc + (a * a + b * b + c * c);
      }
  return double(acc) / 2.0;
}

// --------------------------------------------------------- interpolation --
struct EB { float E[3], B[3]; };

inline float lerp1(float x, float A, float B) { return (1 - x) * A + x * B; }   // ..._linear_1st.h:38-40
inline float lerp3D(const float d[3], const float m[2][2][2]) {                  // ..._linear_1st.h:29-52
  const float c00 = lerp1(d[0], m[0][0][0], m[1][0][0]);
  const float c01 = lerp1(d[0], m[0][0][1], m[1][0][1]);
  const float c10 = lerp1(d[0], m[0][1][0], m[1][1][0]);
  const float c11 = lerp1(d[0], m[0][1][1], m[1][1][1]);
  const float c0 = lerp1(d[1], c00, c10);
  const float c1 = lerp1(d[1], c01, c11);
  return lerp1(d[2], c0, c1);
}
inline float mean2(float a, float b) { return (a + b) / 2.0f; }                  // :87-90
inline float mean4(float a, float b, float c, float d) { return (a + (b + (c + d))) / 4.0f; }  // right fold

// emf/yee_lattice_interpolate_linear_1st.h:58-138 (the _unrolled variant :141-367
// performs the same operations per corner and is bit-identical).
EB interpolate(const Tile& t, const float origo[3], float px, float py, float pz) {
  const float pl[3] = { px - origo[0], py - origo[1], pz - origo[2] };
  const uint32_t i = uint32_t(pl[0]), j = uint32_t(pl[1]), k = uint32_t(pl[2]);
  const float* Ex = &t.E[0]; const float* Ey = &t.E[t.Ch]; const float* Ez = &t.E[2 * t.Ch];
  const float* Bx = &t.B[0]; const float* By = &t.B[t.Ch]; const float* Bz = &t.B[2 * t.Ch];
  float mEx[2][2][2], mEy[2][2][2], mEz[2][2][2], mBx[2][2][2], mBy[2][2][2], mBz[2][2][2];
  for (uint32_t ic = 0; ic < 2; ++ic) for (uint32_t jc = 0; jc < 2; ++jc) for (uint32_t kc = 0; kc < 2; ++kc) {
    const size_t ii = i + ic, jj = j + jc, kk = k + kc;
    mEx[ic][jc][kc] = mean2(Ex[t.lin(ii - 1, jj, kk)], Ex[t.lin(ii, jj, kk)]);
    mEy[ic][jc][kc] = mean2(Ey[t.lin(ii, jj - 1, kk)], Ey[t.lin(ii, jj, kk)]);
    mEz[ic][jc][kc] = mean2(Ez[t.lin(ii, jj, kk - 1)], Ez[t.lin(ii, jj, kk)]);
    mBx[ic][jc][kc] = mean4(Bx[t.lin(ii, jj, kk)], Bx[t.lin(ii, jj - 1, kk)], Bx[t.lin(ii, jj, kk - 1)], Bx[t.lin(ii, jj - 1, kk - 1)]);
    mBy[ic][jc][kc] = mean4(By[t.lin(ii, jj, kk)], By[t.lin(ii - 1, jj, kk)], By[t.lin(ii, jj, kk - 1)], By[t.lin(ii - 1, jj, kk - 1)]);
    mBz[ic][jc][kc] = mean4(Bz[t.lin(ii, jj, kk)], Bz[t.lin(ii - 1, jj, kk)], Bz[t.lin(ii, jj - 1, kk)], Bz[t.lin(ii - 1, jj - 1, kk)]);
  }
  const float d[3] = { pl[0] - float(i), pl[1] - float(j), pl[2] - float(k) };
  EB eb;
  eb.E[0] = lerp3D(d, mEx); eb.E[1] = lerp3D(d, mEy); eb.E[2] = lerp3D(d, mEz);
  eb.B[0] = lerp3D(d, mBx); eb.B[1] = lerp3D(d, mBy); eb.B[2] = lerp3D(d, mBz);
  return eb;
}

// tools/vector.h:248-289
struct V3 { float v[3]; float operator[](int i) const { return v[i]; } float& operator[](int i) { return v[i]; } };
inline V3 operator*(const V3& a, float s) { return { { a[0] * s, a[1] * s, a[2] * s } }; }
inline V3 operator*(float s, const V3& a) { return a * s; }
inline V3 operator/(const V3& a, float s) { return { { a[0] / s, a[1] / s, a[2] / s } }; }
inline V3 operator+(const V3& a, const V3& b) { return { { a[0] + b[0], a[1] + b[1], a[2] + b[2] } }; }
inline V3 operator-(const V3& a, const V3& b) { return { { a[0] - b[0], a[1] - b[1], a[2] - b[2] } }; }
inline float dot(const V3& a, const V3& b) { float r = 0; for (int i = 0; i < 3; ++i) r += a[i] * b[i]; return r; }
inline V3 cross(const V3& a, const V3& b) {
  V3 r;
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = -a[0] * b[2] + a[2] * b[0];
  r[2] = a[0] * b[1] - a[1] * b[0];
  return r;
}

void tile_origo(const Tile& t, float origo[3]) {               // pic/tile.c++:329-332
  for (int d = 0; d < 3; ++d) origo[d] = static_cast<float>(t.mins[d]) - float(H);
}

int sign_of(double v) { return (0.0 < v) - (v < 0.0); }        // tools/math.h:181-183

// pic/particle_boris.h:16-62
void push_boris(const Tile& t, Container& c, double cfl_d) {
  float origo[3]; tile_origo(t, origo);
  const float cfl = static_cast<float>(cfl_d);
  const float qm = static_cast<float>(sign_of(c.charge) / c.mass);
  for (size_t n = 0; n < c.size(); ++n) {
    if (c.id[n] == DEAD) continue;
    const EB eb = interpolate(t, origo, c.x[n], c.y[n], c.z[n]);
    const V3 E{ { eb.E[0], eb.E[1], eb.E[2] } }, B{ { eb.B[0], eb.B[1], eb.B[2] } };
    const V3 v0 = cfl * V3{ { c.ux[n], c.uy[n], c.uz[n] } };
    const V3 E0 = 0.5f * qm * E;
    const V3 u0 = v0 + E0;
    const float ginv = cfl / std::sqrt(cfl * cfl + dot(u0, u0));
    const V3 B0 = 0.5f * qm * ginv * B / cfl;
    const float f = 2.0f / (1.0f + dot(B0, B0));
    const V3 u1 = f * (u0 + cross(u0, B0));
    const V3 u2 = u0 + cross(u1, B0) + E0;
    c.ux[n] = u2[0] / cfl; c.uy[n] = u2[1] / cfl; c.uz[n] = u2[2] / cfl;
    const float ginv2 = cfl / std::sqrt(cfl * cfl + dot(u2, u2));
    c.x[n] = c.x[n] + c.ux[n] * ginv2 * cfl;
    c.y[n] = c.y[n] + c.uy[n] * ginv2 * cfl;
    c.z[n] = c.z[n] + c.uz[n] * ginv2 * cfl;
  }
}

// pic/particle_higuera_cary.h:16-77
void push_higuera_cary(const Tile& t, Container& c, double cfl_d) {
  float origo[3]; tile_origo(t, origo);
  const float cfl = static_cast<float>(cfl_d);
  const float qm = static_cast<float>(sign_of(c.charge) / c.mass);
  const float hqm = 0.5f * qm;
  const float cfl2 = cfl * cfl;
  const float cinv = 1.0f / cfl;
  const float cinv2 = cinv * cinv;
  for (size_t n = 0; n < c.size(); ++n) {
    if (c.id[n] == DEAD) continue;
    const EB eb = interpolate(t, origo, c.x[n], c.y[n], c.z[n]);
    const V3 E{ { eb.E[0], eb.E[1], eb.E[2] } }, B{ { eb.B[0], eb.B[1], eb.B[2] } };
    const V3 v0 = cfl * V3{ { c.ux[n], c.uy[n], c.uz[n] } };
    const V3 E0 = hqm * E;
    const V3 u0 = v0 + E0;
    const V3 Bt = hqm * B;
    const float u0sq = dot(u0, u0);
    const float b2 = dot(Bt, Bt);
    const float bdotu = dot(Bt, u0);
    const float gmb = 1.0f + u0sq * cinv2 - b2 * cinv2;
    const float disc = gmb * gmb + 4.0f * (b2 * cinv2 + bdotu * bdotu * cinv2);
    const float ginv = 1.0f / std::sqrt(0.5f * (gmb + std::sqrt(disc)));
    const float gc = ginv * cinv;
    const V3 B0 = gc * Bt;
    const float f = 2.0f / (1.0f + gc * gc * b2);
    const V3 u1 = f * (u0 + cross(u0, B0));
    const V3 u2 = u0 + cross(u1, B0) + E0;
    const float ginv2 = cfl / std::sqrt(cfl2 + dot(u2, u2));
    c.ux[n] = u2[0] * cinv; c.x[n] += u2[0] * ginv2;
    c.uy[n] = u2[1] * cinv; c.y[n] += u2[1] * ginv2;
    c.uz[n] = u2[2] * cinv; c.z[n] += u2[2] * ginv2;
  }
}

// pic/particle_faraday.h:38-108
void push_faraday(const Tile& t, Container& c, double cfl_d) {
  float origo[3]; tile_origo(t, origo);
  const float cfl = static_cast<float>(cfl_d);
  const float qm = static_cast<float>(sign_of(c.charge) / c.mass);
  for (size_t n = 0; n < c.size(); ++n) {
    if (c.id[n] == DEAD) continue;
    const EB eb = interpolate(t, origo, c.x[n], c.y[n], c.z[n]);
    const V3 E{ { eb.E[0], eb.E[1], eb.E[2] } }, B{ { eb.B[0], eb.B[1], eb.B[2] } };
    const V3 v0 = cfl * V3{ { c.ux[n], c.uy[n], c.uz[n] } };
    const float gcfl = std::sqrt(cfl * cfl + dot(v0, v0));
    const V3 u0 = v0 + 0.5f * qm * E;
    const float geff_cfl = std::sqrt(cfl * cfl + dot(u0, u0));
    const float kappa = 0.5f * qm / geff_cfl;
    const V3 eps = kappa * E;
    const V3 beta = kappa * B;
    const float w0 = gcfl + dot(eps, v0);
    const V3 W = v0 + eps * gcfl + cross(v0, beta) + w0 * eps;
    const float b2 = dot(beta, beta);
    const float f = 1.0f / (1.0f + b2);
    const V3 W_rot = f * (W - cross(beta, W) + dot(beta, W) * beta);
    const float bde = dot(beta, eps);
    const V3 eps_rot = f * (eps - cross(beta, eps) + bde * beta);
    const float D = 1.0f - f * (dot(eps, eps) + bde * bde);
    const V3 u2 = W_rot + eps_rot * (dot(eps, W_rot) / D);
    c.ux[n] = u2[0] / cfl; c.uy[n] = u2[1] / cfl; c.uz[n] = u2[2] / cfl;
    const float ginv2 = cfl / std::sqrt(cfl * cfl + dot(u2, u2));
    c.x[n] = c.x[n] + c.ux[n] * ginv2 * cfl;
    c.y[n] = c.y[n] + c.uy[n] * ginv2 * cfl;
    c.z[n] = c.z[n] + c.uz[n] * ginv2 * cfl;
  }
}

// ---------------------------------------------------------------- deposit --
struct Contribution { uint32_t i, j, k; float J[3]; };

// The 14 (node, current) pairs of one particle: pic/particle_current_zigzag_1st.c++:79-178
// (sort-reduce variant) == :241-336 (atomic variant), same order.
inline void zigzag_contributions(const Container& c, size_t n, const float origo[3], float cfl,
                                 float charge, Contribution out[14]) {
  const V3 u{ { c.ux[n], c.uy[n], c.uz[n] } };
  const float invgam = 1.0f / std::sqrt(1.0f + dot(u, u));
  const V3 x2 = V3{ { c.x[n], c.y[n], c.z[n] } } - V3{ { origo[0], origo[1], origo[2] } };
  const V3 x1 = x2 - cfl * invgam * u;
  const V3 fi1{ { std::floor(x1[0]), std::floor(x1[1]), std::floor(x1[2]) } };
  const V3 fi2{ { std::floor(x2[0]), std::floor(x2[1]), std::floor(x2[2]) } };
  auto relay = [&](int j) {
    const float a = (fi1[j] < fi2[j] ? fi1[j] : fi2[j]) + 1.0f;     // sstd::min, tools/math.h:91-95
    const float b1 = fi1[j] > fi2[j] ? fi1[j] : fi2[j];             // sstd::max
    const float b2 = 0.5f * (x1[j] + x2[j]);
    const float b = b1 > b2 ? b1 : b2;
    return a < b ? a : b;
  };
  const V3 xr{ { relay(0), relay(1), relay(2) } };
  const V3 F1 = charge * (xr - x1);
  const V3 F2 = charge * (x2 - xr);
  const uint32_t i1[3] = { uint32_t(fi1[0]), uint32_t(fi1[1]), uint32_t(fi1[2]) };
  const uint32_t i2[3] = { uint32_t(fi2[0]), uint32_t(fi2[1]), uint32_t(fi2[2]) };
  const V3 W1 = 0.5f * (x1 + xr) - fi1;
  const V3 W2 = 0.5f * (x2 + xr) - fi2;
  const float Fx1 = F1[0], Fy1 = F1[1], Fz1 = F1[2], Fx2 = F2[0], Fy2 = F2[1], Fz2 = F2[2];
  const float Wx1 = W1[0], Wy1 = W1[1], Wz1 = W1[2], Wx2 = W2[0], Wy2 = W2[1], Wz2 = W2[2];
  const float one = 1.0f;
  auto put = [&](int s, const uint32_t b[3], int di, int dj, int dk, float jx, float jy, float jz) {
    out[s] = { b[0] + di, b[1] + dj, b[2] + dk, { jx, jy, jz } };
  };
  put(0, i1, 0, 0, 0, Fx1 * (one - Wy1) * (one - Wz1), Fy1 * (one - Wx1) * (one - Wz1), Fz1 * (one - Wx1) * (one - Wy1));
  put(1, i2, 0, 0, 0, Fx2 * (one - Wy2) * (one - Wz2), Fy2 * (one - Wx2) * (one - Wz2), Fz2 * (one - Wx2) * (one - Wy2));
  put(2, i1, 1, 0, 0, 0, Fy1 * Wx1 * (one - Wz1), Fz1 * Wx1 * (one - Wy1));
  put(3, i2, 1, 0, 0, 0, Fy2 * Wx2 * (one - Wz2), Fz2 * Wx2 * (one - Wy2));
  put(4, i1, 0, 1, 0, Fx1 * Wy1 * (one - Wz1), 0, Fz1 * (one - Wx1) * Wy1);
  put(5, i2, 0, 1, 0, Fx2 * Wy2 * (one - Wz2), 0, Fz2 * (one - Wx2) * Wy2);
  put(6, i1, 0, 0, 1, Fx1 * (one - Wy1) * Wz1, Fy1 * (one - Wx1) * Wz1, 0);
  put(7, i2, 0, 0, 1, Fx2 * (one - Wy2) * Wz2, Fy2 * (one - Wx2) * Wz2, 0);
  put(8, i1, 0, 1, 1, Fx1 * Wy1 * Wz1, 0, 0);
  put(9, i2, 0, 1, 1, Fx2 * Wy2 * Wz2, 0, 0);
  put(10, i1, 1, 0, 1, 0, Fy1 * Wx1 * Wz1, 0);
  put(11, i2, 1, 0, 1, 0, Fy2 * Wx2 * Wz2, 0);
  put(12, i1, 1, 1, 0, 0, 0, Fz1 * Wx1 * Wy1);
  put(13, i2, 1, 1, 0, 0, 0, Fz2 * Wx2 * Wy2);
}

// pic/tile.c++:369-415
void deposit_current(Tile& t, const b2p_config& cfg) {
  std::fill(t.J.begin(), t.J.end(), 0.0f);                       // clear_current, emf/yee_lattice.c++:317-321
  float origo[3]; tile_origo(t, origo);
  const float cfl = static_cast<float>(cfg.cfl);
  Contribution cb[14];
  if (cfg.current_depositer == B2P_DEPOSIT_ZIGZAG_1ST_ATOMIC) {
    t.genJ.assign(3 * t.Ch, 0.0f);                               // pic/tile.c++:393-398
    for (const Container& c : t.sp) {                            // particle_current_zigzag_1st.c++:226-339
      const float charge = static_cast<float>(c.charge);
      for (size_t n = 0; n < c.size(); ++n) {
        if (c.id[n] == DEAD) continue;
        zigzag_contributions(c, n, origo, cfl, charge, cb);
        for (int s = 0; s < 14; ++s) {
          const size_t l = t.lin(cb[s].i, cb[s].j, cb[s].k);
          for (int d = 0; d < 3; ++d) t.genJ[d * t.Ch + l] += cb[s].J[d];   // sstd::atomic_add, serial order
        }
      }
    }
    for (size_t n = 0; n < 3 * t.Ch; ++n) t.J[n] = t.J[n] + t.genJ[n];      // emf/yee_lattice.c++:361-375
  } else {
    // particle_current_zigzag_1st.c++:30-223: 14N pairs, stable sort by (i,j,k),
    // sequential reduce_by_key, then J[loc] += reduced (emf/yee_lattice.c++:331-350).
    for (const Container& c : t.sp) {
      const float charge = static_cast<float>(c.charge);
      std::vector<Contribution> all;
      all.reserve(14 * c.size());
      for (size_t n = 0; n < c.size(); ++n) {
        if (c.id[n] == DEAD) continue;   // the reference stores zero currents for dead slots
        zigzag_contributions(c, n, origo, cfl, charge, cb);
        all.insert(all.end(), cb, cb + 14);
      }
      std::stable_sort(all.begin(), all.end(), [](const Contribution& a, const Contribution& b) {
        return (a.i < b.i) || (a.i == b.i && a.j < b.j) || (a.i == b.i && a.j == b.j && a.k < b.k);
      });
      size_t p = 0;
      while (p < all.size()) {
        size_t e = p + 1;
        float acc[3] = { all[p].J[0], all[p].J[1], all[p].J[2] };
        while (e < all.size() && all[e].i == all[p].i && all[e].j == all[p].j && all[e].k == all[p].k) {
          for (int d = 0; d < 3; ++d) acc[d] = acc[d] + all[e].J[d];
          ++e;
        }
        const size_t l = t.lin(all[p].i, all[p].j, all[p].k);
        for (int d = 0; d < 3; ++d) t.J[d * t.Ch + l] = t.J[d * t.Ch + l] + acc[d];
        p = e;
      }
    }
  }
  if (t.corr_pending) {                                            // pic/tile.c++:411-414
    for (size_t n = 0; n < 3 * t.Ch; ++n) t.J[n] = t.J[n] + t.corrJ[n];   // emf/yee_lattice.c++:361-375
    t.corr_pending = false;
  }
}

// ------------------------------------------------- pic-shock boundary pieces --
// emf/tile.c++:808-827: how many interior cells of the tile the edge region covers
bool edge_bc_width(const Tile& t, const b2p_edge_bc& bc, size_t* width) {
  const int d = bc.direction;
  const float tile_min = static_cast<float>(t.mins[d]);
  const float tile_max = static_cast<float>(t.maxs[d]);
  const size_t Nd = size_t(t.N[d]);
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

// emf/yee_lattice.c++:263-306
int apply_edge_bc(Tile& t, const b2p_edge_bc& bc, int mode) {
  size_t width = 0;
  if (!edge_bc_width(t, bc, &width)) return 0;                     // emf/tile.c++:835-840
  if (width == 0) return 0;
  const int d = bc.direction;
  const size_t Nd = size_t(t.N[d]);
  const size_t w = std::min(width, Nd);
  size_t lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    if (a != d) { lo[a] = 0; hi[a] = size_t(t.Hx[a]); }
    else if (bc.side == 0) { lo[a] = 0; hi[a] = size_t(H) + w; }
    else { lo[a] = size_t(H) + Nd - w; hi[a] = size_t(t.Hx[a]); }
  }
  std::vector<float>* f;
  uint8_t mask;
  const float* v;
  switch (mode) {
    case B2P_COMM_EMF_E: f = &t.E; mask = bc.E_components; v = bc.E; break;
    case B2P_COMM_EMF_B: f = &t.B; mask = bc.B_components; v = bc.B; break;
    case B2P_COMM_EMF_J: f = &t.J; mask = bc.J_components; v = bc.J; break;
    default:
      g_err = "YeeLattice::apply_edge_bc does not support given communication mode: " + std::to_string(mode);
      return B2P_ERR_RUNTIME;
  }
  for (size_t i = lo[0]; i < hi[0]; ++i)
    for (size_t j = lo[1]; j < hi[1]; ++j)
      for (size_t k = lo[2]; k < hi[2]; ++k)
        for (int c = 0; c < 3; ++c)
          if (mask & (1u << c)) (*f)[c * t.Ch + t.lin(i, j, k)] = v[c];
  return 0;
}

// pic/reflector_wall.c++:35-118: atomic zigzag deposit of ONE sub-trajectory x1 -> x2 (lattice-local
// coordinates): 14 nodes x 3 components, explicit zeros included, in source order.
void zigzag_deposit_single(Tile& t, std::vector<float>& J, const V3 x1, const V3 x2, const float charge) {
  const V3 fi1{ { std::floor(x1[0]), std::floor(x1[1]), std::floor(x1[2]) } };
  const V3 fi2{ { std::floor(x2[0]), std::floor(x2[1]), std::floor(x2[2]) } };
  auto relay = [&](int j) {
    const float a = (fi1[j] < fi2[j] ? fi1[j] : fi2[j]) + 1.0f;
    const float m = fi1[j] > fi2[j] ? fi1[j] : fi2[j];
    const float h = 0.5f * (x1[j] + x2[j]);
    const float b = m > h ? m : h;
    return a < b ? a : b;
  };
  const V3 xr{ { relay(0), relay(1), relay(2) } };
  const V3 F1 = charge * (xr - x1);
  const V3 F2 = charge * (x2 - xr);
  const uint32_t i1[3] = { uint32_t(fi1[0]), uint32_t(fi1[1]), uint32_t(fi1[2]) };
  const uint32_t i2[3] = { uint32_t(fi2[0]), uint32_t(fi2[1]), uint32_t(fi2[2]) };
  const V3 W1 = 0.5f * (x1 + xr) - fi1;
  const V3 W2 = 0.5f * (x2 + xr) - fi2;
  const float Fx1 = F1[0], Fy1 = F1[1], Fz1 = F1[2], Fx2 = F2[0], Fy2 = F2[1], Fz2 = F2[2];
  const float Wx1 = W1[0], Wy1 = W1[1], Wz1 = W1[2], Wx2 = W2[0], Wy2 = W2[1], Wz2 = W2[2];
  const float one = 1.0f;
  auto store = [&](const uint32_t b[3], int di, int dj, int dk, float jx, float jy, float jz) {
    const size_t l = t.lin(b[0] + di, b[1] + dj, b[2] + dk);
    J[0 * t.Ch + l] += jx; J[1 * t.Ch + l] += jy; J[2 * t.Ch + l] += jz;
  };
  store(i1, 0, 0, 0, Fx1 * (one - Wy1) * (one - Wz1), Fy1 * (one - Wx1) * (one - Wz1), Fz1 * (one - Wx1) * (one - Wy1));
  store(i2, 0, 0, 0, Fx2 * (one - Wy2) * (one - Wz2), Fy2 * (one - Wx2) * (one - Wz2), Fz2 * (one - Wx2) * (one - Wy2));
  store(i1, 1, 0, 0, 0, Fy1 * Wx1 * (one - Wz1), Fz1 * Wx1 * (one - Wy1));
  store(i2, 1, 0, 0, 0, Fy2 * Wx2 * (one - Wz2), Fz2 * Wx2 * (one - Wy2));
  store(i1, 0, 1, 0, Fx1 * Wy1 * (one - Wz1), 0, Fz1 * (one - Wx1) * Wy1);
  store(i2, 0, 1, 0, Fx2 * Wy2 * (one - Wz2), 0, Fz2 * (one - Wx2) * Wy2);
  store(i1, 0, 0, 1, Fx1 * (one - Wy1) * Wz1, Fy1 * (one - Wx1) * Wz1, 0);
  store(i2, 0, 0, 1, Fx2 * (one - Wy2) * Wz2, Fy2 * (one - Wx2) * Wz2, 0);
  store(i1, 0, 1, 1, Fx1 * Wy1 * Wz1, 0, 0);
  store(i2, 0, 1, 1, Fx2 * Wy2 * Wz2, 0, 0);
  store(i1, 1, 0, 1, 0, Fy1 * Wx1 * Wz1, 0);
  store(i2, 1, 0, 1, 0, Fy2 * Wx2 * Wz2, 0);
  store(i1, 1, 1, 0, 0, 0, Fz1 * Wx1 * Wy1);
  store(i2, 1, 1, 0, 0, 0, Fz2 * Wx2 * Wy2);
}

// ParticleContainer::reflect_at_wall, pic/reflector_wall.c++:126-222 (branch-free float masks)
void reflect_at_wall(Tile& t, Container& cont, const b2p_reflector_wall& wall, const float origo_[3], const double cfl) {
  const float EPS = 1e-10f;                                          // :23
  const float walloc = wall.walloc, betawall = wall.betawall, gammawall = wall.gammawall;
  const float charge = static_cast<float>(cont.charge);
  const float c = static_cast<float>(cfl);
  const float walloc0 = walloc - betawall * c;                       // previous wall location
  const V3 origo{ { origo_[0], origo_[1], origo_[2] } };
  for (size_t n = 0; n < cont.size(); ++n) {
    if (cont.id[n] == DEAD) continue;
    const V3 pos_1{ { cont.x[n], cont.y[n], cont.z[n] } };
    const V3 u{ { cont.ux[n], cont.uy[n], cont.uz[n] } };
    const float gam = std::sqrt(1.0f + dot(u, u));
    const float invgam = 1.0f / gam;
    const V3 pos_0 = pos_1 - c * invgam * u;
    const float mask_skip = (pos_1[0] >= walloc) ? 1.0f : 0.0f;
    const float mask_close = (walloc0 - pos_0[0] <= c) ? 1.0f : 0.0f;
    const float denom = betawall * c - c * u[0] * invgam;
    const float dt = std::fabs((pos_0[0] - walloc0) / (denom + EPS));
    const float mask_crossed = (dt <= 1.0f) ? 1.0f : 0.0f;
    const float mask_refl = (1.0f - mask_skip) * mask_close * mask_crossed;
    const float mask_park = (1.0f - mask_skip) - mask_refl;
    const V3 pos_col = pos_0 + c * dt * invgam * u;
    const float ux_new = gammawall * gammawall * gam * (2.0f * betawall - u[0] * invgam * (1.0f + betawall * betawall));
    const V3 u_new{ { ux_new, u[1], u[2] } };
    const float gam_new = std::sqrt(1.0f + dot(u_new, u_new));
    const float invgam_new = 1.0f / gam_new;
    const float ratio = std::fabs((pos_1[0] - pos_col[0]) / (pos_1[0] - pos_0[0] + EPS));
    const float dt_refl = 1.0f < ratio ? 1.0f : ratio;               // sstd::min(1.0f, ratio)
    const V3 pos_refl = pos_col + c * dt_refl * invgam_new * u_new;
    const V3 p1l = pos_1 - origo;
    const V3 dep_fwd_from = p1l + mask_refl * (pos_0 - pos_1);
    const V3 dep_fwd_to = p1l + mask_refl * (pos_col - pos_1);
    zigzag_deposit_single(t, t.corrJ, dep_fwd_from, dep_fwd_to, mask_refl * charge);
    const V3 x1_deposit = pos_refl - c * invgam_new * u_new;
    const V3 dep_rev_from = p1l + mask_refl * (x1_deposit - pos_1);
    const V3 dep_rev_to = p1l + mask_refl * (pos_col - pos_1);
    zigzag_deposit_single(t, t.corrJ, dep_rev_from, dep_rev_to, mask_refl * (-charge));
    cont.x[n] = (1.0f - mask_refl) * pos_1[0] + mask_refl * pos_refl[0];
    cont.y[n] = (1.0f - mask_refl) * pos_1[1] + mask_refl * pos_refl[1];
    cont.z[n] = (1.0f - mask_refl) * pos_1[2] + mask_refl * pos_refl[2];
    cont.ux[n] = (1.0f - mask_refl) * u[0] + mask_refl * ux_new;
    if (mask_park > 0.5f) cont.id[n] = DEAD;
  }
}

// pic::Tile::reflect_particles, pic/reflector_wall.c++:241-284
void reflect_particles(Tile& t, const b2p_config& cfg) {
  if (t.walls.empty()) return;
  auto wall_is_in_tile = [&](const b2p_reflector_wall& w) {
    return w.walloc >= float(t.mins[0]) - float(cfg.cfl) && w.walloc <= float(t.maxs[0]);
  };
  bool any = false;
  for (const b2p_reflector_wall& w : t.walls) any = any || wall_is_in_tile(w);
  if (!any) return;
  t.corr_pending = true;
  t.corrJ.assign(3 * t.Ch, 0.0f);
  float origo[3]; tile_origo(t, origo);
  for (const b2p_reflector_wall& w : t.walls) {
    if (!wall_is_in_tile(w)) continue;
    for (Container& c : t.sp) reflect_at_wall(t, c, w, origo, cfg.cfl);
  }
}

// ------------------------------------------------------------------- sort --
// pic/tile.c++:419-438 + pic/particle.h:575-703
void sort_keys(const Tile& t, const Container& c, std::vector<uint32_t>& keys) {
  float origo[3]; tile_origo(t, origo);
  keys.resize(c.size());
  for (size_t n = 0; n < c.size(); ++n) {
    if (c.id[n] == DEAD) { keys[n] = std::numeric_limits<uint32_t>::max(); continue; }
    const uint32_t i = uint32_t(c.x[n] - origo[0]);
    const uint32_t j = uint32_t(c.y[n] - origo[1]);
    const uint32_t k = uint32_t(c.z[n] - origo[2]);
    keys[n] = (i * uint32_t(t.Hx[1]) + j) * uint32_t(t.Hx[2]) + k;   // layout_right mapping, uint32 index_type
  }
}

void sort_particles(const Tile& t, Container& c) {
  std::vector<uint32_t> keys;
  sort_keys(t, c, keys);
  std::vector<uint32_t> perm(c.size());
  std::iota(perm.begin(), perm.end(), 0u);
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  auto gatherf = [&](std::vector<float>& v) {
    std::vector<float> tmp(v);
    for (size_t n = 0; n < v.size(); ++n) v[n] = tmp[perm[n]];
  };
  gatherf(c.x); gatherf(c.y); gatherf(c.z); gatherf(c.ux); gatherf(c.uy); gatherf(c.uz);
  std::vector<uint64_t> tmp(c.id);
  for (size_t n = 0; n < c.id.size(); ++n) c.id[n] = tmp[perm[n]];
}

// -------------------------------------------------------------- migration --
// communication_common.h:137-147: ((dx+1)*3 + (dy+1))*3 + (dz+1)
inline int subregion_index(int i, int j, int k) { return ((i + 1) * 3 + (j + 1)) * 3 + (k + 1); }

// pic/particle.c++:199-348 via pic/tile_communication.c++:68-96
void pack_outgoing(Tile& t) {
  const float xd[2] = { float(t.mins[0]), float(t.maxs[0]) };
  const float yd[2] = { float(t.mins[1]), float(t.maxs[1]) };
  const float zd[2] = { float(t.mins[2]), float(t.maxs[2]) };
  auto sub = [&](float x, float y, float z) {
    const int i = int(x >= xd[0]) - int(x < xd[1]);
    const int j = int(y >= yd[0]) - int(y < yd[1]);
    const int k = int(z >= zd[0]) - int(z < zd[1]);
    return subregion_index(i, j, k);
  };
  t.out_ends.assign(27 * t.sp.size(), 0);
  t.out_buf.clear();
  for (size_t s = 0; s < t.sp.size(); ++s) {
    Container& c = t.sp[s];
    const size_t prev = t.out_buf.size();
    for (size_t n = 0; n < c.size(); ++n)
      if (c.id[n] != DEAD && sub(c.x[n], c.y[n], c.z[n]) != 13)
        t.out_buf.push_back({ { c.x[n], c.y[n], c.z[n] }, { c.ux[n], c.uy[n], c.uz[n] }, c.id[n] });
    for (size_t n = 0; n < c.size(); ++n)                       // marks every slot whose position leaves
      if (sub(c.x[n], c.y[n], c.z[n]) != 13) c.id[n] = DEAD;
    std::stable_sort(t.out_buf.begin() + prev, t.out_buf.end(),
                     [&](const b2p_particle_state& a, const b2p_particle_state& b) {
                       return sub(a.pos[0], a.pos[1], a.pos[2]) < sub(b.pos[0], b.pos[1], b.pos[2]);
                     });
    size_t counts[27] = { 0 };
    for (size_t n = prev; n < t.out_buf.size(); ++n)
      counts[sub(t.out_buf[n].pos[0], t.out_buf[n].pos[1], t.out_buf[n].pos[2])]++;
    size_t next = prev;
    for (int r = 0; r < 27; ++r) {
      if (r != 13) next += counts[r];
      t.out_ends[27 * s + r] = next;
    }
  }
}

// pic/particle.h:454-572
void append_spans(Container& c, const std::vector<Span>& spans, const float* wmin, const float* wmax) {
  if (spans.empty()) return;
  size_t P = c.size();
  while (P > 0 && c.id[P - 1] == DEAD) --P;                       // find_if over reversed ordinals
  size_t total = 0;
  for (const Span& s : spans) total += s.n;
  c.resize(P + total);
  size_t off = P;
  for (const Span& s : spans) {
    for (size_t i = 0; i < s.n; ++i) {
      const b2p_particle_state& st = s.p[i];
      const size_t j = off + i;
      if (wmin && wmax) {
        const float Lx = wmax[0] - wmin[0], Ly = wmax[1] - wmin[1], Lz = wmax[2] - wmin[2];
        c.x[j] = (st.pos[0] < 0 ? wmax[0] : wmin[0]) + std::fmod(st.pos[0], Lx);
        c.y[j] = (st.pos[1] < 0 ? wmax[1] : wmin[1]) + std::fmod(st.pos[1], Ly);
        c.z[j] = (st.pos[2] < 0 ? wmax[2] : wmin[2]) + std::fmod(st.pos[2], Lz);
      } else {
        c.x[j] = st.pos[0]; c.y[j] = st.pos[1]; c.z[j] = st.pos[2];
      }
      c.ux[j] = st.vel[0]; c.uy[j] = st.vel[1]; c.uz[j] = st.vel[2];
      c.id[j] = st.id;
    }
    off += s.n;
  }
}

// pic/particle.c++:352-377
double kinetic_energy(const Container& c) {
  double acc = 0.0;
  for (size_t n = 0; n < c.size(); ++n) {
    const bool alive = c.id[n] != DEAD;
    const V3 v{ { c.ux[n], c.uy[n], c.uz[n] } };
    const float e = std::sqrt(1.0f + dot(v, v)) - 1.0f;
    acc += double(alive) * static_cast<double>(e);
  }
  return acc;
}

// ------------------------------------------------------------ halo regions --
struct Range { int b, e; };
// emf/yee_lattice.h:493-517
Range subregion(const Tile& t, int d, int dir) {
  switch (dir) {
    case -1: return { 0, H };
    case 0: return { H, H + t.N[d] };
    default: return { H + t.N[d], 2 * H + t.N[d] };
  }
}
// emf/yee_lattice.h:580-602 (switch on the inverted direction)
Range corresponding_subregion(const Tile& t, int d, int dir) {
  switch (-dir) {
    case -1: return { H, 2 * H };
    case 0: return { H, H + t.N[d] };
    default: return { t.N[d], H + t.N[d] };
  }
}

// emf/yee_lattice.c++:206-239: me.subregion(dir) <- other.corresponding_subregion(dir)
void set_in_subregion(Tile& me, std::vector<float>& mine, const Tile& other, const std::vector<float>& theirs,
                      const int dir[3]) {
  Range a[3], b[3];
  for (int d = 0; d < 3; ++d) { a[d] = subregion(me, d, dir[d]); b[d] = corresponding_subregion(other, d, dir[d]); }
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < a[0].e - a[0].b; ++i)
      for (int j = 0; j < a[1].e - a[1].b; ++j)
        for (int k = 0; k < a[2].e - a[2].b; ++k)
          mine[c * me.Ch + me.lin(a[0].b + i, a[1].b + j, a[2].b + k)] =
            theirs[c * other.Ch + other.lin(b[0].b + i, b[1].b + j, b[2].b + k)];
}

// emf/yee_lattice.c++:249-261: me.corresponding_subregion(-dir) += other.subregion(-dir)
void add_to_J_from_subregion(Tile& me, const Tile& other, const int dir[3]) {
  Range a[3], b[3];
  for (int d = 0; d < 3; ++d) { a[d] = corresponding_subregion(me, d, -dir[d]); b[d] = subregion(other, d, -dir[d]); }
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < a[0].e - a[0].b; ++i)
      for (int j = 0; j < a[1].e - a[1].b; ++j)
        for (int k = 0; k < a[2].e - a[2].b; ++k) {
          const size_t m = c * me.Ch + me.lin(a[0].b + i, a[1].b + j, a[2].b + k);
          me.J[m] = me.J[m] + other.J[c * other.Ch + other.lin(b[0].b + i, b[1].b + j, b[2].b + k)];
        }
}

int wrap(int v, int n) { while (v < 0) v += n; while (v >= n) v -= n; return v; }   // corgi/tile.h:126-136

template <class F>
void parallel_tiles(orc_grid* g, int threads, F f) {
  const int nt = int(g->tiles.size());
  if (threads <= 1 || nt <= 1) { for (int t = 0; t < nt; ++t) f(g->tiles[t]); return; }
  std::atomic<int> next{ 0 };
  std::vector<std::thread> pool;
  const int nw = std::min(threads, nt);
  for (int w = 0; w < nw; ++w)
    pool.emplace_back([&] { for (int t; (t = next.fetch_add(1)) < nt;) f(g->tiles[t]); });
  for (auto& th : pool) th.join();
}

void tile_push_half_b(orc_grid* g, Tile& t) {                     // emf/tile.c++:359-375
  const float dt = static_cast<float>(g->cfg.cfl / 2);
  if (g->cfg.field_propagator == B2P_PROPAGATOR_STENCIL) push_b_stencil(t, dt, g->stencilM);
  else push_b_fdtd2(t, dt);
}
void tile_push_e(orc_grid* g, Tile& t) { push_e_fdtd2(t, static_cast<float>(g->cfg.cfl)); }   // emf/tile.c++:379-394
int tile_filter(orc_grid* g, Tile& t) {                           // emf/tile.c++:405-426
  if (g->cfg.current_filter == B2P_FILTER_BINOMIAL2) filter_binomial2(t);
  else if (g->cfg.current_filter == B2P_FILTER_BINOMIAL2_UNROLLED) filter_binomial2_unrolled(t);
  else { g_err = "Trying to filter current without specifying `current_filter`!"; return B2P_ERR_LOGIC; }
  return 0;
}
void tile_push_particles(orc_grid* g, Tile& t) {                  // pic/tile.c++:326-365
  for (Container& c : t.sp) {
    switch (g->cfg.particle_pusher) {
      case B2P_PUSHER_BORIS: push_boris(t, c, g->cfg.cfl); break;
      case B2P_PUSHER_HIGUERA_CARY: push_higuera_cary(t, c, g->cfg.cfl); break;
      default: push_faraday(t, c, g->cfg.cfl); break;
    }
  }
}

Tile* get_tile(orc_grid* g, int t) {
  if (!g || t < 0 || t >= int(g->tiles.size())) { g_err = "bad tile handle"; return nullptr; }
  return &g->tiles[t];
}
Container* get_container(orc_grid* g, int t, int sp) {
  Tile* tl = get_tile(g, t);
  if (!tl) return nullptr;
  if (sp < 0 || sp >= int(tl->sp.size())) { g_err = "particle type not configured"; return nullptr; }
  return &tl->sp[sp];
}

}  // namespace

// ================================================================== C ABI ==
extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

orc_grid* orc_create(const b2p_config* cfg) {
  for (int d = 0; d < 3; ++d) {
    if (cfg->n_cells[d] < H) { g_err = "Yee Lattice extents are assumed to be at least halo size"; return nullptr; }
    if (cfg->n_tiles[d] < 1) { g_err = "n_tiles must be positive"; return nullptr; }
  }
  auto* g = new orc_grid;
  g->cfg = *cfg;
  std::memcpy(g->stencilM, cfg->stencil, sizeof(g->stencilM));
  for (int a = 0; a < 3; ++a) g->stencilM[a][0][0] = stencil_alpha(g->stencilM[a]);
  const int Tx = cfg->n_tiles[0], Ty = cfg->n_tiles[1], Tz = cfg->n_tiles[2];
  g->tiles.resize(size_t(Tx) * Ty * Tz);
  for (int d = 0; d < 3; ++d) { g->gmins[d] = 0.0; g->gmaxs[d] = double(cfg->n_tiles[d]) * cfg->n_cells[d]; }
  for (int k = 0; k < Tz; ++k) for (int j = 0; j < Ty; ++j) for (int i = 0; i < Tx; ++i) {
    Tile& t = g->tiles[i + Tx * (j + Ty * k)];
    t.idx[0] = i; t.idx[1] = j; t.idx[2] = k;
    t.Ch = 1;
    for (int d = 0; d < 3; ++d) {
      t.N[d] = cfg->n_cells[d]; t.Hx[d] = t.N[d] + 2 * H; t.Ch *= size_t(t.Hx[d]);
      t.mins[d] = double(size_t(t.idx[d]) * size_t(t.N[d]));
      t.maxs[d] = double((size_t(t.idx[d]) + 1) * size_t(t.N[d]));
    }
    t.E.assign(3 * t.Ch, 0.0f); t.B.assign(3 * t.Ch, 0.0f); t.J.assign(3 * t.Ch, 0.0f);
    t.sp.resize(cfg->n_species);
    t.next_ordinal.assign(cfg->n_species, 0);
    t.incoming.resize(cfg->n_species);
    for (int s = 0; s < cfg->n_species; ++s) {
      t.sp[s].charge = cfg->q[s]; t.sp[s].mass = cfg->m[s];
      if (cfg->prealloc_per_species) {                             // pic/particle.c++:39-61
        t.sp[s].resize(cfg->prealloc_per_species);
        std::fill(t.sp[s].id.begin(), t.sp[s].id.end(), DEAD);
      }
    }
    t.tile_tag = (uint64_t(i) * Ty + j) * Tz + k;                  // pic/tile.c++:141-143 (layout_right)
  }
  return g;
}

void orc_destroy(orc_grid* g) { delete g; }
int orc_num_tiles(const orc_grid* g) { return int(g->tiles.size()); }
int orc_tile_cid(const orc_grid* g, int i, int j, int k) {
  return i + g->cfg.n_tiles[0] * (j + g->cfg.n_tiles[1] * k);
}

static void copy_field(const Tile& t, std::vector<float>& dst, const float* src, int with_halo) {
  if (!src) return;
  if (with_halo) { std::memcpy(dst.data(), src, 3 * t.Ch * sizeof(float)); return; }
  const size_t Ni = size_t(t.N[0]) * t.N[1] * t.N[2];
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < t.N[0]; ++i) for (int j = 0; j < t.N[1]; ++j) for (int k = 0; k < t.N[2]; ++k)
      dst[c * t.Ch + t.lin(i + H, j + H, k + H)] = src[c * Ni + (size_t(i) * t.N[1] + j) * t.N[2] + k];
}
static void read_field(const Tile& t, const std::vector<float>& src, float* dst, int with_halo) {
  if (!dst) return;
  if (with_halo) { std::memcpy(dst, src.data(), 3 * t.Ch * sizeof(float)); return; }
  const size_t Ni = size_t(t.N[0]) * t.N[1] * t.N[2];
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < t.N[0]; ++i) for (int j = 0; j < t.N[1]; ++j) for (int k = 0; k < t.N[2]; ++k)
      dst[c * Ni + (size_t(i) * t.N[1] + j) * t.N[2] + k] = src[c * t.Ch + t.lin(i + H, j + H, k + H)];
}

int orc_tile_set_fields(orc_grid* g, int t, const float* E, const float* B, const float* J, int with_halo) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  copy_field(*tl, tl->E, E, with_halo); copy_field(*tl, tl->B, B, with_halo); copy_field(*tl, tl->J, J, with_halo);
  return 0;
}
int orc_tile_get_fields(orc_grid* g, int t, float* E, float* B, float* J, int with_halo) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  read_field(*tl, tl->E, E, with_halo); read_field(*tl, tl->B, B, with_halo); read_field(*tl, tl->J, J, with_halo);
  return 0;
}
int orc_tile_push_half_b(orc_grid* g, int t) { Tile* tl = get_tile(g, t); if (!tl) return 1; tile_push_half_b(g, *tl); return 0; }
int orc_tile_push_e(orc_grid* g, int t) { Tile* tl = get_tile(g, t); if (!tl) return 1; tile_push_e(g, *tl); return 0; }
int orc_tile_add_current(orc_grid* g, int t) { Tile* tl = get_tile(g, t); if (!tl) return 1; add_current(*tl); return 0; }
int orc_tile_filter_current(orc_grid* g, int t) { Tile* tl = get_tile(g, t); if (!tl) return 1; return tile_filter(g, *tl); }
int orc_tile_clear_current(orc_grid* g, int t) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  std::fill(tl->J.begin(), tl->J.end(), 0.0f); return 0;
}
int orc_tile_field_energy(orc_grid* g, int t, double* eB, double* eE) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  if (eB) *eB = field_energy(*tl, tl->B);
  if (eE) *eE = field_energy(*tl, tl->E);
  return 0;
}

int orc_tile_inject(orc_grid* g, int t, int sp, uint64_t n, const double* x, const double* y, const double* z,
                    const double* ux, const double* uy, const double* uz) {
  Tile* tl = get_tile(g, t); Container* c = get_container(g, t, sp); if (!c) return 1;
  // pic/tile.c++:207-217,304-321 (ids) + pic/particle.h:258-287 (narrowing + append, no wrap)
  std::vector<b2p_particle_state> st(n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t ordinal = tl->next_ordinal[sp]++;
    st[i] = { { float(x[i]), float(y[i]), float(z[i]) }, { float(ux[i]), float(uy[i]), float(uz[i]) },
              (tl->tile_tag << 40) | ordinal };
  }
  // NB: an empty batch still runs append's find-last-alive trim (pic/particle.h:469-509)
  append_spans(*c, { Span{ st.data(), size_t(n) } }, nullptr, nullptr);
  return 0;
}

int orc_tile_set_particles(orc_grid* g, int t, int sp, uint64_t n, const float* x, const float* y, const float* z,
                           const float* ux, const float* uy, const float* uz, const uint64_t* id) {
  Container* c = get_container(g, t, sp); if (!c) return 1;
  c->resize(n);
  if (n) {
    std::memcpy(c->x.data(), x, n * 4); std::memcpy(c->y.data(), y, n * 4); std::memcpy(c->z.data(), z, n * 4);
    std::memcpy(c->ux.data(), ux, n * 4); std::memcpy(c->uy.data(), uy, n * 4); std::memcpy(c->uz.data(), uz, n * 4);
    std::memcpy(c->id.data(), id, n * 8);
  }
  return 0;
}
int orc_tile_container_size(orc_grid* g, int t, int sp, uint64_t* n) {
  Container* c = get_container(g, t, sp); if (!c) return 1; *n = c->size(); return 0;
}
int orc_tile_get_particles(orc_grid* g, int t, int sp, int alive_only, float* x, float* y, float* z,
                           float* ux, float* uy, float* uz, uint64_t* id, uint64_t* n_out) {
  Container* c = get_container(g, t, sp); if (!c) return 1;
  uint64_t m = 0;
  for (size_t n = 0; n < c->size(); ++n) {                        // pic/particle.c++:82-168
    if (alive_only && c->id[n] == DEAD) continue;
    if (x) x[m] = c->x[n]; if (y) y[m] = c->y[n]; if (z) z[m] = c->z[n];
    if (ux) ux[m] = c->ux[n]; if (uy) uy[m] = c->uy[n]; if (uz) uz[m] = c->uz[n];
    if (id) id[m] = c->id[n];
    ++m;
  }
  if (n_out) *n_out = m;
  return 0;
}
int orc_tile_push_particles(orc_grid* g, int t) { Tile* tl = get_tile(g, t); if (!tl) return 1; tile_push_particles(g, *tl); return 0; }
int orc_tile_deposit_current(orc_grid* g, int t) { Tile* tl = get_tile(g, t); if (!tl) return 1; deposit_current(*tl, g->cfg); return 0; }
int orc_tile_sort_particles(orc_grid* g, int t) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  for (Container& c : tl->sp) sort_particles(*tl, c);
  return 0;
}
int orc_tile_pack_outgoing_particles(orc_grid* g, int t) { Tile* tl = get_tile(g, t); if (!tl) return 1; pack_outgoing(*tl); return 0; }
int orc_tile_sort_keys(orc_grid* g, int t, int sp, uint32_t* keys) {
  Tile* tl = get_tile(g, t); Container* c = get_container(g, t, sp); if (!c) return 1;
  std::vector<uint32_t> k; sort_keys(*tl, *c, k);
  std::memcpy(keys, k.data(), k.size() * 4);
  return 0;
}
int orc_tile_get_outgoing(orc_grid* g, int t, b2p_particle_state* buf, uint64_t cap, uint64_t* ends, uint64_t* n_out) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  if (n_out) *n_out = tl->out_buf.size();
  if (ends) for (size_t i = 0; i < tl->out_ends.size(); ++i) ends[i] = tl->out_ends[i];
  if (buf) {
    if (cap < tl->out_buf.size()) { g_err = "outgoing buffer too small"; return 1; }
    std::memcpy(buf, tl->out_buf.data(), tl->out_buf.size() * sizeof(b2p_particle_state));
  }
  return 0;
}
int orc_tile_kinetic_energy(orc_grid* g, int t, int sp, double* energy, uint64_t* container_size) {
  Container* c = get_container(g, t, sp); if (!c) return 1;
  if (energy) *energy = kinetic_energy(*c);
  if (container_size) *container_size = c->size();
  return 0;
}
int orc_tile_register_antenna(orc_grid* g, int t, const b2p_antenna_mode* m) {   // emf/tile.c++:566-576
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  Tile::Antenna a;
  for (int d = 0; d < 3; ++d) { a.A[d] = m->A[d]; a.wave[d] = m->wave[d]; }
  a.kind = m->wave_kind;
  a.has_coeffs = m->lap_coeffs != nullptr;
  for (uint64_t q = 0; a.has_coeffs && q < m->n_lap_coeffs; ++q) a.coeffs.push_back({ m->lap_coeffs[2 * q], m->lap_coeffs[2 * q + 1] });
  tl->antennas.push_back(std::move(a));
  return 0;
}

// emf::Tile::deposit_antenna_current, emf/tile.c++:578-777
int orc_tile_deposit_antenna_current(orc_grid* g, int t) {
  Tile* tp = get_tile(g, t); if (!tp) return 1;
  Tile& tl = *tp;
  const size_t nm = tl.antennas.size();
  std::vector<std::array<float, 3>> A(nm), K(nm);
  std::vector<std::array<float, 2>> W(nm);
  for (size_t n = 0; n < nm; ++n) {
    Tile::Antenna& a = tl.antennas[n];
    double k[3];
    if (a.kind == 0) { for (int d = 0; d < 3; ++d) k[d] = a.wave[d]; }
    else {                                                            // :603-614
      for (int d = 0; d < 3; ++d) {
        const double L = g->gmaxs[d] - g->gmins[d];
        const double tmp = 2 * 3.141592653589793238462643383279502884 * a.wave[d];
        k[d] = tmp / L;
      }
    }
    for (int d = 0; d < 3; ++d) { A[n][d] = static_cast<float>(a.A[d]); K[n][d] = static_cast<float>(k[d]); }
    if (a.has_coeffs && a.next >= a.coeffs.size()) {
      g_err = "Can not deposit antenna current, antenna_mode ran out of lap_coeffs!";
      return B2P_ERR_LOGIC;
    } else if (a.has_coeffs) {
      W[n] = { static_cast<float>(a.coeffs[a.next][0]), static_cast<float>(a.coeffs[a.next][1]) };
      ++a.next;
    } else {
      W[n] = { 1.0f, 0.0f };
    }
  }
  tl.vec_pot.assign(3 * tl.Ch, 0.0f);
  tl.gen_B.assign(3 * tl.Ch, 0.0f);
  const double Lt[3] = { tl.maxs[0] - tl.mins[0], tl.maxs[1] - tl.mins[1], tl.maxs[2] - tl.mins[2] };
  auto gcmap = [&](double i, double j, double k, double out[3]) {     // emf/tile.h:207-229
    out[0] = tl.mins[0] + (i / double(tl.N[0])) * Lt[0];
    out[1] = tl.mins[1] + (j / double(tl.N[1])) * Lt[1];
    out[2] = tl.mins[2] + (k / double(tl.N[2])) * Lt[2];
  };
  for (int ii = 0; ii < tl.Hx[0]; ++ii)
    for (int jj = 0; jj < tl.Hx[1]; ++jj)
      for (int kk = 0; kk < tl.Hx[2]; ++kk) {
        const double i = double(ii) - H, j = double(jj) - H, k = double(kk) - H;
        double loc[3][3];
        gcmap(i + 0.5, j, k, loc[0]);
        gcmap(i, j + 0.5, k, loc[1]);
        gcmap(i, j, k + 0.5, loc[2]);
        const size_t l = tl.lin(ii, jj, kk);
        for (size_t n = 0; n < nm; ++n) {
          for (int c = 0; c < 3; ++c) {
            double dotv = 0;                                          // toolbox::dot, from 0 left to right
            for (int d = 0; d < 3; ++d) dotv += loc[c][d] * double(K[n][d]);
            const float phi = static_cast<float>(dotv);
            const float re = ::cosf(phi), im = ::sinf(phi);
            tl.vec_pot[c * tl.Ch + l] = tl.vec_pot[c * tl.Ch + l] + A[n][c] * (W[n][0] * re - W[n][1] * im);
          }
        }
      }
  // B = curl(vec_pot) on the interior plus a one-deep shell (:700-744)
  auto X = [&](const std::vector<float>& f, int c, int i, int j, int k) { return f[c * tl.Ch + tl.lin(i, j, k)]; };
  for (int i = H - 1; i < H + tl.N[0] + 1; ++i)
    for (int j = H - 1; j < H + tl.N[1] + 1; ++j)
      for (int k = H - 1; k < H + tl.N[2] + 1; ++k) {
        const std::vector<float>& P = tl.vec_pot;
        const size_t l = tl.lin(i, j, k);
        { const float Dk = X(P, 1, i, j, k + 1) - X(P, 1, i, j, k), Dj = X(P, 2, i, j + 1, k) - X(P, 2, i, j, k);
          tl.gen_B[0 * tl.Ch + l] = 1.0f * (Dj - Dk); }
        { const float Di = X(P, 2, i + 1, j, k) - X(P, 2, i, j, k), Dk = X(P, 0, i, j, k + 1) - X(P, 0, i, j, k);
          tl.gen_B[1 * tl.Ch + l] = 1.0f * (Dk - Di); }
        { const float Dj = X(P, 0, i, j + 1, k) - X(P, 0, i, j, k), Di = X(P, 1, i + 1, j, k) - X(P, 1, i, j, k);
          tl.gen_B[2 * tl.Ch + l] = 1.0f * (Di - Dj); }
      }
  // curl(B) in the interior, written over vec_pot with coefficient -cfl and backward neighbours (:747-771)
  const float coeff = static_cast<float>(-g->cfg.cfl);
  for (int i = H; i < H + tl.N[0]; ++i)
    for (int j = H; j < H + tl.N[1]; ++j)
      for (int k = H; k < H + tl.N[2]; ++k) {
        const std::vector<float>& Bf = tl.gen_B;
        const size_t l = tl.lin(i, j, k);
        float o0, o1, o2;
        { const float Dk = X(Bf, 1, i, j, k - 1) - X(Bf, 1, i, j, k), Dj = X(Bf, 2, i, j - 1, k) - X(Bf, 2, i, j, k); o0 = coeff * (Dj - Dk); }
        { const float Di = X(Bf, 2, i - 1, j, k) - X(Bf, 2, i, j, k), Dk = X(Bf, 0, i, j, k - 1) - X(Bf, 0, i, j, k); o1 = coeff * (Dk - Di); }
        { const float Dj = X(Bf, 0, i, j - 1, k) - X(Bf, 0, i, j, k), Di = X(Bf, 1, i - 1, j, k) - X(Bf, 1, i, j, k); o2 = coeff * (Di - Dj); }
        tl.vec_pot[0 * tl.Ch + l] = o0; tl.vec_pot[1 * tl.Ch + l] = o1; tl.vec_pot[2 * tl.Ch + l] = o2;
      }
  // YeeLattice::deposit_current(vec_pot) over the WHOLE haloed lattice (:773; emf/yee_lattice.c++:361-375): outside the
  // interior vec_pot still holds the potential itself, exactly as in the reference
  for (size_t n = 0; n < 3 * tl.Ch; ++n) tl.J[n] = tl.J[n] + tl.vec_pot[n];
  return 0;
}

int orc_tile_register_edge_bc(orc_grid* g, int t, const b2p_edge_bc* bc) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  tl->edge_bcs.push_back(*bc);                                       // emf/tile.c++:829-833
  return 0;
}
int orc_tile_apply_edge_bc(orc_grid* g, int t, const b2p_edge_bc* bc, int mode) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  return apply_edge_bc(*tl, *bc, mode);
}
int orc_tile_apply_edge_bcs(orc_grid* g, int t, int mode) {         // emf/tile.c++:842-847
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  for (const b2p_edge_bc& bc : tl->edge_bcs) { const int rc = apply_edge_bc(*tl, bc, mode); if (rc) return rc; }
  return 0;
}
int orc_tile_register_reflector_wall(orc_grid* g, int t, const b2p_reflector_wall* wall) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  tl->walls.push_back(*wall);                                        // pic/reflector_wall.c++:226-233
  return 0;
}
int orc_tile_reflect_particles(orc_grid* g, int t) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  reflect_particles(*tl, g->cfg);
  return 0;
}
int orc_tile_advance_reflector_walls(orc_grid* g, int t) {          // pic/reflector_wall.c++:286-297
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  for (b2p_reflector_wall& w : tl->walls) w.walloc += w.betawall * float(g->cfg.cfl);
  return 0;
}
int orc_tile_reflector_walls(orc_grid* g, int t, b2p_reflector_wall* out, uint64_t cap, uint64_t* n) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  *n = tl->walls.size();
  for (size_t q = 0; q < tl->walls.size() && q < cap; ++q) out[q] = tl->walls[q];
  return 0;
}

int orc_tile_interpolate(orc_grid* g, int t, uint64_t n, const float* x, const float* y, const float* z, float* out) {
  Tile* tl = get_tile(g, t); if (!tl) return 1;
  float origo[3]; tile_origo(*tl, origo);
  for (uint64_t i = 0; i < n; ++i) {
    const EB eb = interpolate(*tl, origo, x[i], y[i], z[i]);
    for (int d = 0; d < 3; ++d) { out[6 * i + d] = eb.E[d]; out[6 * i + 3 + d] = eb.B[d]; }
  }
  return 0;
}

// external/corgi/src/corgi/corgi.h:1697-1718 with emf/tile.c++:478-542 and
// pic/tile_communication.c++:100-195. Moore order: cellular_automata.h:48-62.
int orc_local_communication(orc_grid* g, int mode) {
  switch (mode) {
    case B2P_COMM_EMF_E: case B2P_COMM_EMF_B: case B2P_COMM_EMF_J: case B2P_COMM_EMF_J_EXCHANGE: case B2P_COMM_PIC_PARTICLE: break;
    default: g_err = "local_communication does not support given communication mode"; return B2P_ERR_LOGIC;
  }
  if (mode == B2P_COMM_PIC_PARTICLE)
    for (const Tile& t : g->tiles)
      if (t.out_ends.size() != 27 * t.sp.size()) { g_err = "pack_outgoing_particles not called"; return 1; }
  const int* T = g->cfg.n_tiles;
  // corgi loops directions outermost and tiles innermost; no tile reads anything another
  // tile writes within one mode, so tiles can be processed independently (here: one
  // worker per tile, the reference's one-rank-per-core execution model) as long as each
  // tile keeps the Moore order kr -> jr -> ir of its own 26 exchanges.
  parallel_tiles(g, g->threads, [&](Tile& me) {
    for (int kr = -1; kr <= 1; ++kr) for (int jr = -1; jr <= 1; ++jr) for (int ir = -1; ir <= 1; ++ir) {
      if (ir == 0 && jr == 0 && kr == 0) continue;
      const int dir[3] = { ir, jr, kr };
      const int oi = wrap(me.idx[0] + ir, T[0]), oj = wrap(me.idx[1] + jr, T[1]), ok = wrap(me.idx[2] + kr, T[2]);
      Tile& other = g->tiles[oi + T[0] * (oj + T[1] * ok)];
      switch (mode) {
        case B2P_COMM_EMF_E: set_in_subregion(me, me.E, other, other.E, dir); break;
        case B2P_COMM_EMF_B: set_in_subregion(me, me.B, other, other.B, dir); break;
        case B2P_COMM_EMF_J: set_in_subregion(me, me.J, other, other.J, dir); break;
        case B2P_COMM_EMF_J_EXCHANGE: add_to_J_from_subregion(me, other, dir); break;
        default: {
          const int inv = subregion_index(-ir, -jr, -kr);
          for (size_t s = 0; s < me.sp.size(); ++s) {
            const size_t index = 27 * s + inv;
            const size_t end = other.out_ends[index];
            const size_t begin = index == 0 ? 0 : other.out_ends[index - 1];
            me.incoming[s].push_back({ other.out_buf.data() + begin, end - begin });
          }
        }
      }
    }
  });
  if (mode == B2P_COMM_PIC_PARTICLE) {                             // postlude, tile_communication.c++:100-117
    float wmin[3], wmax[3];
    for (int d = 0; d < 3; ++d) { wmin[d] = float(g->gmins[d]); wmax[d] = float(g->gmaxs[d]); }
    parallel_tiles(g, g->threads, [&](Tile& me) {
      for (size_t s = 0; s < me.sp.size(); ++s) { append_spans(me.sp[s], me.incoming[s], wmin, wmax); me.incoming[s].clear(); }
    });
    for (Tile& me : g->tiles) me.out_ends.clear();
  }
  return 0;
}

int orc_grid_phase(orc_grid* g, const char* phase, int threads) {
  const std::string p(phase);
  int rc = 0;
  if (p == "push_half_b") parallel_tiles(g, threads, [&](Tile& t) { tile_push_half_b(g, t); });
  else if (p == "push_e") parallel_tiles(g, threads, [&](Tile& t) { tile_push_e(g, t); });
  else if (p == "add_current") parallel_tiles(g, threads, [&](Tile& t) { add_current(t); });
  else if (p == "filter_current") {
    if (g->cfg.current_filter < 0) { g_err = "Trying to filter current without specifying `current_filter`!"; return B2P_ERR_LOGIC; }
    parallel_tiles(g, threads, [&](Tile& t) { tile_filter(g, t); });
  }
  else if (p == "push_particles") parallel_tiles(g, threads, [&](Tile& t) { tile_push_particles(g, t); });
  else if (p == "pack_outgoing_particles") parallel_tiles(g, threads, [&](Tile& t) { pack_outgoing(t); });
  else if (p == "sort_particles") parallel_tiles(g, threads, [&](Tile& t) { for (Container& c : t.sp) sort_particles(t, c); });
  else if (p == "deposit_current") parallel_tiles(g, threads, [&](Tile& t) { deposit_current(t, g->cfg); });
  else if (p == "reflect_particles") parallel_tiles(g, threads, [&](Tile& t) { reflect_particles(t, g->cfg); });
  else if (p == "advance_reflector_walls")
    parallel_tiles(g, threads, [&](Tile& t) { for (b2p_reflector_wall& w : t.walls) w.walloc += w.betawall * float(g->cfg.cfl); });
  else if (p == "apply_edge_bcs_E" || p == "apply_edge_bcs_B" || p == "apply_edge_bcs_J") {
    const int mode = p.back() == 'E' ? B2P_COMM_EMF_E : p.back() == 'B' ? B2P_COMM_EMF_B : B2P_COMM_EMF_J;
    parallel_tiles(g, threads, [&](Tile& t) { for (const b2p_edge_bc& bc : t.edge_bcs) apply_edge_bc(t, bc, mode); });
  }
  else { g_err = "unknown phase " + p; rc = 1; }
  return rc;
}

// projects/pic-turbulence/pic.py:187-221 (IO/diagnostics excluded; comm_external is a
// no-op in a single process)
int orc_step_pic(orc_grid* g, int64_t lap, int threads) {
  int rc = 0;
  g->threads = threads;
#define DO(x) do { rc = (x); if (rc) return rc; } while (0)
  DO(orc_grid_phase(g, "push_half_b", threads));
  DO(orc_local_communication(g, B2P_COMM_EMF_B));
  DO(orc_grid_phase(g, "push_particles", threads));
  DO(orc_grid_phase(g, "pack_outgoing_particles", threads));
  DO(orc_local_communication(g, B2P_COMM_PIC_PARTICLE));
  if (lap % 5 == 0) DO(orc_grid_phase(g, "sort_particles", threads));
  DO(orc_grid_phase(g, "deposit_current", threads));
  DO(orc_local_communication(g, B2P_COMM_EMF_J_EXCHANGE));
  DO(orc_local_communication(g, B2P_COMM_EMF_J));
  if (g->cfg.current_filter >= 0) {
    DO(orc_grid_phase(g, "filter_current", threads));
    DO(orc_local_communication(g, B2P_COMM_EMF_J));
    DO(orc_grid_phase(g, "filter_current", threads));
    DO(orc_grid_phase(g, "filter_current", threads));
  }
  DO(orc_grid_phase(g, "push_half_b", threads));
  DO(orc_local_communication(g, B2P_COMM_EMF_B));
  DO(orc_grid_phase(g, "push_e", threads));
  DO(orc_grid_phase(g, "add_current", threads));
  DO(orc_local_communication(g, B2P_COMM_EMF_E));
  return 0;
}

// projects/emf-wave/emf.py:48-62
int orc_step_emf(orc_grid* g, int threads) {
  int rc = 0;
  g->threads = threads;
  DO(orc_local_communication(g, B2P_COMM_EMF_E));
  DO(orc_grid_phase(g, "push_half_b", threads));
  DO(orc_grid_phase(g, "push_half_b", threads));
  DO(orc_local_communication(g, B2P_COMM_EMF_B));
  DO(orc_grid_phase(g, "push_e", threads));
#undef DO
  return 0;
}

// mpiio::write_header (io/snapshots/mpiio_fields.c++:22-78; layout mpiio_header.h:56-82)
static void snapshot_header(char buf[512], int32_t nx, int32_t ny, int32_t nz, int32_t stride, const b2p_config& c, int32_t lap,
                            int32_t num_fields) {
  std::memset(buf, 0, 512);
  auto put = [&](int off, uint32_t v) { std::memcpy(buf + off, &v, 4); };
  put(0, 0x524E4B4Fu); put(4, 3u); put(8, 512u); put(12, uint32_t(num_fields));
  put(16, uint32_t(nx)); put(20, uint32_t(ny)); put(24, uint32_t(nz)); put(28, uint32_t(stride));
  put(32, uint32_t(c.n_tiles[0])); put(36, uint32_t(c.n_tiles[1])); put(40, uint32_t(c.n_tiles[2]));
  put(44, uint32_t(c.n_cells[0])); put(48, uint32_t(c.n_cells[1])); put(52, uint32_t(c.n_cells[2]));
  put(56, uint32_t(lap)); put(60, 4u);
  static const char* names[9] = { "ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz" };
  for (int f = 0; f < 9; ++f) std::memcpy(buf + 64 + f * 16, names[f], std::strlen(names[f]));
  for (int s = 0; s < num_fields - 9; ++s) {
    const std::string nm = "n" + std::to_string(s);
    std::memcpy(buf + 64 + (9 + s) * 16, nm.data(), std::min<size_t>(nm.size(), 15));
  }
}

// FieldsWriter<3>::write: pack_tile (:221-321) + write_payload_ (:345-400), plain stdio instead of MPI-IO
int orc_write_fields_snapshot(orc_grid* g, const char* prefix, int32_t lap, int32_t stride, int32_t nspecies) {
  if (stride < 1) { g_err = "snapshot stride must be >= 1"; return B2P_ERR_RUNTIME; }
  const b2p_config& c = g->cfg;
  const int nsp = std::min(nspecies, 5);                                   // mpiio_header.h:26
  const int nf = 9 + std::max(nsp, 0);
  const int nxt = std::max(1, c.n_cells[0] / stride), nyt = std::max(1, c.n_cells[1] / stride), nzt = std::max(1, c.n_cells[2] / stride);
  const int nx = c.n_tiles[0] * nxt, ny = c.n_tiles[1] * nyt, nz = c.n_tiles[2] * nzt;
  const std::string fn = std::string(prefix) + "/flds_" + std::to_string(lap) + ".bin";
  FILE* fh = std::fopen(fn.c_str(), "wb");
  if (!fh) { g_err = "cannot open " + fn; return B2P_ERR_RUNTIME; }
  char hdr[512];
  snapshot_header(hdr, nx, ny, nz, stride, c, lap, nf);
  std::fwrite(hdr, 1, 512, fh);
  const size_t field_elems = size_t(nx) * ny * nz, tile_elems = size_t(nxt) * nyt * nzt;
  std::vector<float> zeros(field_elems, 0.0f);
  for (int f = 0; f < nf; ++f) std::fwrite(zeros.data(), 4, field_elems, fh);   // MPI_File_set_size: zero-filled
  std::vector<float> buf(size_t(nf) * tile_elems);
  for (const Tile& t : g->tiles) {
    std::fill(buf.begin(), buf.end(), 0.0f);
    for (int iz = 0; iz < nzt; ++iz)
      for (int iy = 0; iy < nyt; ++iy)
        for (int ix = 0; ix < nxt; ++ix) {
          const size_t o = (size_t(iz) * nyt + iy) * nxt + ix;
          const size_t l = t.lin(size_t(ix) * stride + H, size_t(iy) * stride + H, size_t(iz) * stride + H);
          for (int d = 0; d < 3; ++d) {
            buf[(0 + d) * tile_elems + o] = t.E[d * t.Ch + l];
            buf[(3 + d) * tile_elems + o] = t.B[d * t.Ch + l];
          }
          float sj[3] = { 0.0f, 0.0f, 0.0f };
          for (int kk = 0; kk < stride; ++kk)
            for (int jj = 0; jj < stride; ++jj)
              for (int ii = 0; ii < stride; ++ii) {
                const size_t m = t.lin(size_t(ix) * stride + ii + H, size_t(iy) * stride + jj + H, size_t(iz) * stride + kk + H);
                for (int d = 0; d < 3; ++d) sj[d] += t.J[d * t.Ch + m];
              }
          for (int d = 0; d < 3; ++d) buf[(6 + d) * tile_elems + o] = sj[d];
        }
    const float mx = float(t.mins[0]), my = float(t.mins[1]), mz = float(t.mins[2]);
    const float inv_stride = 1.0f / float(stride);
    const int ndep = std::min<int>(int(t.sp.size()), nsp);
    for (int s = 0; s < ndep; ++s) {
      const Container& pc = t.sp[s];
      for (size_t n = 0; n < pc.size(); ++n) {
        if (pc.id[n] == DEAD) continue;
        const size_t ci = size_t(std::floor((pc.x[n] - mx) * inv_stride));
        const size_t cj = size_t(std::floor((pc.y[n] - my) * inv_stride));
        const size_t ck = size_t(std::floor((pc.z[n] - mz) * inv_stride));
        if (ci >= size_t(nxt) || cj >= size_t(nyt) || ck >= size_t(nzt)) continue;   // outside the coarse tile: UB in the reference
        buf[(9 + s) * tile_elems + (ck * nyt + cj) * nxt + ci] += 1.0f;
      }
    }
    for (int f = 0; f < nf; ++f)
      for (int ks = 0; ks < nzt; ++ks)
        for (int js = 0; js < nyt; ++js) {
          const size_t off = 512 + 4 * (size_t(f) * field_elems +
                                        (size_t(t.idx[2] * nzt + ks) * ny + size_t(t.idx[1] * nyt + js)) * nx + size_t(t.idx[0]) * nxt);
          std::fseek(fh, long(off), SEEK_SET);
          std::fwrite(buf.data() + size_t(f) * tile_elems + (size_t(ks) * nyt + js) * nxt, 4, size_t(nxt), fh);
        }
  }
  std::fclose(fh);
  return 0;
}

int orc_energies(orc_grid* g, double* eB, double* eE, double* kinetic, uint64_t* sizes) {
  double b = 0, e = 0;
  for (const Tile& t : g->tiles) { b += field_energy(t, t.B); e += field_energy(t, t.E); }
  if (eB) *eB = b;
  if (eE) *eE = e;
  for (int s = 0; s < g->cfg.n_species; ++s) {
    double k = 0; uint64_t n = 0;
    for (const Tile& t : g->tiles) { k += kinetic_energy(t.sp[s]); n += t.sp[s].size(); }
    if (kinetic) kinetic[s] = k;
    if (sizes) sizes[s] = n;
  }
  return 0;
}

}  // extern "C"
