// Host-side setup algebra of the B200 MGMC path (no CUDA in this header).
//
// The reference builds every coarse operator as an Eigen sparse triple product R A R^T
// (LinearOperator::coarsen, linear_operator/linear_operator.cc:10-23) and keeps dense n x m
// matrices for the low-rank smoother correction (SORSmoother ctor, smoother/sor_smoother.cc:17-38).
// Here the operators stay matrix-free: on a structured lattice with constant kappa the Galerkin
// product of a stencil with the full-weighting transfer {1/2,1,1/2}^(x)2 is again a stencil, exactly
//   A_c(I, I+D) = sum_{p,q in {-1,0,1}^2} w(p) w(q) a_{s = 2I+p}(2D + q - p),
// constant in the interior and (for the squared operator only) different on the first / last
// interior line.  A level therefore stores 9 position classes x 25 coefficients.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mgmc {

// position class along one axis: 0 = first interior line, 2 = last interior line, 1 = in between
inline int pos_class(int i, int n) { return i == 1 ? 0 : (i == n - 1 ? 2 : 1); }

struct StencilSet {
  double a[9][25];  // a[cx + 3 cy][(dj + 2) * 5 + (di + 2)]
  int radius = 1;
  int ncolours = 2;
  bool uniform = true;  // all classes identical
  double at(int cls, int di, int dj) const { return a[cls][(dj + 2) * 5 + (di + 2)]; }
};

inline void classify(StencilSet &s) {
  s.radius = 0;
  bool corners = false;
  s.uniform = true;
  for (int c = 0; c < 9; ++c) {
    const int cx = c % 3, cy = c / 3;
    for (int dj = -2; dj <= 2; ++dj)
      for (int di = -2; di <= 2; ++di) {
        const double v = s.at(c, di, dj);
        if (v != 0.0) {
          s.radius = std::max(s.radius, std::max(std::abs(di), std::abs(dj)));
          if (di != 0 && dj != 0) corners = true;
        }
        // entries that point across the boundary from the first / last interior line multiply the
        // zero ghost lines: they are irrelevant for the comparison between classes
        const bool relevant = (cx != 0 || di >= 0) && (cx != 2 || di <= 0) && (cy != 0 || dj >= 0) && (cy != 2 || dj <= 0);
        const double ref = s.at(4, di, dj);
        if (relevant && std::fabs(v - ref) > 1e-14 * std::fabs(s.at(4, 0, 0))) s.uniform = false;
      }
  }
  s.ncolours = (s.radius >= 2) ? 9 : (corners ? 4 : 2);
}

// ShiftedLaplaceFEMOperator (linear_operator/shiftedlaplace_fem_operator.cc:9-150) with a constant correlation length on a
// dim-dimensional lattice (dim = 2, 3): multilinear elements, 2-point Gauss rule per direction.  Every interior vertex lies in 2^dim
// cells, so the operator is one uniform 3^dim-point stencil: the cell with lower corner v + c, c in {-1, 0}^dim, has v as its corner
// alpha = -c and couples it to the corners beta, i.e. to the vertices v + c + beta:
//   a(c + beta) += sum_q (kappa^2 phi_alpha phi_beta + sum_d h_d^-2 d_d phi_alpha d_d phi_beta)(xhat_q) w_q * cell volume   (:127-135)
// out[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] (dim = 2: the plane dk = 0)
inline void fem_stencil(int dim, const int *n, double Lambda, double *out27) {
  std::memset(out27, 0, sizeof(double) * 27);
  double hinv2[3] = {0, 0, 0}, vol = 1.0;
  for (int d = 0; d < dim; ++d) {
    const double h = 1. / double(n[d]);
    hinv2[d] = 1. / (h * h);
    vol *= h;
  }
  const double kappa_sq = 1.0 / std::pow(Lambda, 2);
  const double gp[2] = {0.5 * (1.0 - 1.0 / std::sqrt(3.0)), 0.5 * (1.0 + 1.0 / std::sqrt(3.0))};
  const int ncorner = 1 << dim;
  const double wq = 1.0 / ncorner;  // product of the 1d weights 1/2
  for (int cell = 0; cell < ncorner; ++cell) {   // bit d set: c_d = -1, i.e. alpha_d = 1
    for (int beta = 0; beta < ncorner; ++beta) {
      double entry = 0.0;
      for (int q = 0; q < ncorner; ++q) {
        double pa = 1.0, pb = 1.0, grad = 0.0;
        double fa[3], fb[3];
        for (int d = 0; d < dim; ++d) {
          const double x = gp[(q >> d) & 1];
          fa[d] = ((cell >> d) & 1) ? x : 1.0 - x;
          fb[d] = ((beta >> d) & 1) ? x : 1.0 - x;
          pa *= fa[d];
          pb *= fb[d];
        }
        for (int d = 0; d < dim; ++d) {
          double ga = ((cell >> d) & 1) ? 1.0 : -1.0, gb = ((beta >> d) & 1) ? 1.0 : -1.0;
          for (int e = 0; e < dim; ++e)
            if (e != d) {
              ga *= fa[e];
              gb *= fb[e];
            }
          grad += hinv2[d] * ga * gb;
        }
        entry += (kappa_sq * pa * pb + grad) * wq;
      }
      int sh[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d) sh[d] = -((cell >> d) & 1) + ((beta >> d) & 1);
      out27[(sh[2] + 1) * 9 + (sh[1] + 1) * 3 + (sh[0] + 1)] += entry * vol;
    }
  }
}

// fine-level stencils
//  pde 0: ShiftedLaplaceFDOperator (shiftedlaplace_fd_operator.cc:33-56): h^d (kappa^2 + sum 2/h_d^2), -h^d/h_d^2
//  pde 1: SquaredShiftedLaplaceFDOperator (squared_shiftedlaplace_fd_operator.cc:40-93)
//  pde 2: ShiftedLaplaceFEMOperator (shiftedlaplace_fem_operator.cc:9-150), constant correlation length
inline StencilSet fine_stencil(int pde, int nx, int ny, double Lambda) {
  StencilSet s;
  std::memset(s.a, 0, sizeof(s.a));
  const double hx = 1.0 / double(nx), hy = 1.0 / double(ny);
  const double hinv2x = 1.0 / (hx * hx), hinv2y = 1.0 / (hy * hy);
  const double vol = hx * hy;
  const double kappa_sq = 1.0 / std::pow(Lambda, 2);
  for (int cy = 0; cy < 3; ++cy)
    for (int cx = 0; cx < 3; ++cx) {
      double *a = s.a[cx + 3 * cy];
      auto A = [&](int di, int dj) -> double & { return a[(dj + 2) * 5 + (di + 2)]; };
      if (pde == 2) {
        // shiftedlaplace_fem: uniform 9-point stencil (all classes alike: entries towards the boundary multiply zeros)
        const int n2[2] = {nx, ny};
        double a27[27];
        fem_stencil(2, n2, Lambda, a27);
        for (int dj = -1; dj <= 1; ++dj)
          for (int di = -1; di <= 1; ++di) A(di, dj) = a27[9 + (dj + 1) * 3 + (di + 1)];
      } else if (pde == 0) {
        double diagonal = vol * kappa_sq;
        diagonal += 2. * vol * hinv2x;
        diagonal += 2. * vol * hinv2y;
        A(0, 0) = diagonal;
        A(-1, 0) = A(+1, 0) = -vol * hinv2x;
        A(0, -1) = A(0, +1) = -vol * hinv2y;
      } else {
        const double lap00 = -2 * (hinv2x + hinv2y), lap10 = hinv2x, lap01 = hinv2y;
        double ss[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        ss[0][0] = 6 * (hinv2x * hinv2x + hinv2y * hinv2y) + 8 * hinv2x * hinv2y;
        ss[1][0] = -4 * hinv2x * (hinv2x + hinv2y);
        ss[0][1] = -4 * hinv2y * (hinv2x + hinv2y);
        ss[2][0] = hinv2x * hinv2x;
        ss[0][2] = hinv2y * hinv2y;
        ss[1][1] = 2 * hinv2x * hinv2y;
        double diagonal = (kappa_sq * kappa_sq - 2. * kappa_sq * lap00 + ss[0][0]) * vol;
        for (int j = -2; j <= 2; ++j)
          for (int k = -2; k <= 2; ++k) {
            if ((std::abs(j) + std::abs(k) > 2) || (j == 0 && k == 0)) continue;
            double e = ss[std::abs(j)][std::abs(k)];
            if (std::abs(j) + std::abs(k) == 1) e += -2. * kappa_sq * ((j != 0) ? lap10 : lap01);
            A(j, k) = e * vol;
          }
        // missing +-1 neighbour => add the +-2 coefficient to the diagonal (:83-91)
        if (cx == 0) diagonal += ss[2][0] * vol;
        if (cx == 2 || nx == 2) diagonal += ss[2][0] * vol;
        if (cy == 0) diagonal += ss[0][2] * vol;
        if (cy == 2 || ny == 2) diagonal += ss[0][2] * vol;
        A(0, 0) = diagonal;
      }
    }
  classify(s);
  return s;
}

// Galerkin coarsening of a stencil set (nxf, nyf = fine cells)
inline StencilSet coarsen_stencil(const StencilSet &f, int nxf, int nyf) {
  StencilSet c;
  std::memset(c.a, 0, sizeof(c.a));
  const int nxc = nxf / 2, nyc = nyf / 2;
  const double w1[3] = {0.5, 1.0, 0.5};
  for (int cy = 0; cy < 3; ++cy)
    for (int cx = 0; cx < 3; ++cx) {
      double *a = c.a[cx + 3 * cy];
      for (int Dy = -2; Dy <= 2; ++Dy)
        for (int Dx = -2; Dx <= 2; ++Dx) {
          // representative coarse vertex of the class for THIS entry: all interior vertices share one
          // stencil, so pick one whose target (I + Dx, J + Dy) lies inside the lattice
          const int I = (cx == 0) ? 1 : (cx == 2 ? nxc - 1 : (Dx < 0 ? nxc - 2 : 2));
          const int J = (cy == 0) ? 1 : (cy == 2 ? nyc - 1 : (Dy < 0 ? nyc - 2 : 2));
          if (I < 1 || I > nxc - 1 || J < 1 || J > nyc - 1 || pos_class(I, nxc) != cx || pos_class(J, nyc) != cy) continue;
          if (I + Dx < 1 || I + Dx > nxc - 1 || J + Dy < 1 || J + Dy > nyc - 1) continue;
          double acc = 0.0;
          for (int py = -1; py <= 1; ++py)
            for (int px = -1; px <= 1; ++px) {
              const int si = 2 * I + px, sj = 2 * J + py;
              const int fc = pos_class(si, nxf) + 3 * pos_class(sj, nyf);
              for (int qy = -1; qy <= 1; ++qy)
                for (int qx = -1; qx <= 1; ++qx) {
                  const int ox = 2 * Dx + qx - px, oy = 2 * Dy + qy - py;
                  if (std::abs(ox) > 2 || std::abs(oy) > 2) continue;
                  acc += w1[px + 1] * w1[py + 1] * w1[qx + 1] * w1[qy + 1] * f.at(fc, ox, oy);
                }
            }
          a[(Dy + 2) * 5 + (Dx + 2)] = acc;
        }
    }
  // classes without a representative (tiny lattices) copy the interior class / class 0
  for (int k = 0; k < 9; ++k) {
    bool zero = true;
    for (int e = 0; e < 25; ++e) zero = zero && (c.a[k][e] == 0.0);
    if (zero) std::memcpy(c.a[k], c.a[4][12] != 0.0 ? c.a[4] : c.a[0], sizeof(c.a[k]));
  }
  classify(c);
  return c;
}

// ---- 3d lattices (Lattice3d, lattice/lattice3d.hh:43-270): shiftedlaplace_fd with a constant correlation length ----
// 7-point fine operator (shiftedlaplace_fd_operator.cc:33-56 with dim = 3: diagonal h^3 kappa^2 + sum_d 2 h^3 / h_d^2 accumulated
// in the reference's order, off-diagonals -h^3 / h_d^2).  Entries that point to a boundary vertex multiply the zero
// boundary planes of the device layout, so one stencil serves every vertex.
inline void fine_stencil3(int nx, int ny, int nz, double Lambda, double *a27) {
  std::memset(a27, 0, sizeof(double) * 27);
  const int n[3] = {nx, ny, nz};
  double hinv2[3], vol = 1.0;
  for (int d = 0; d < 3; ++d) {
    const double h = 1. / double(n[d]);
    hinv2[d] = 1. / (h * h);
    vol *= h;
  }
  double diagonal = vol * (1.0 / std::pow(Lambda, 2));
  const int step[3] = {1, 3, 9};
  for (int d = 0; d < 3; ++d) {
    a27[13 - step[d]] = a27[13 + step[d]] = -vol * hinv2[d];
    diagonal += 2. * vol * hinv2[d];
  }
  a27[13] = diagonal;
}

// Galerkin product R A R^T (linear_operator.cc:12-15) of a uniform radius-1 stencil with the trilinear full weighting
// {1/2, 1, 1/2}^(x)3 (intergrid_operator_linear.cc:8-30): again a uniform radius-1 (27-point) stencil,
//   A_c(D) = sum_{p, q in {-1,0,1}^3} w(p) w(q) a(2 D + q - p)
// (every fine vertex 2 I + p of an interior coarse vertex I is interior, so no position classes arise).
inline void coarsen_stencil3(const double *f27, double *c27) {
  const double w1[3] = {0.5, 1.0, 0.5};
  for (int Dz = -1; Dz <= 1; ++Dz)
    for (int Dy = -1; Dy <= 1; ++Dy)
      for (int Dx = -1; Dx <= 1; ++Dx) {
        double acc = 0.0;
        for (int pz = -1; pz <= 1; ++pz)
          for (int py = -1; py <= 1; ++py)
            for (int px = -1; px <= 1; ++px)
              for (int qz = -1; qz <= 1; ++qz)
                for (int qy = -1; qy <= 1; ++qy)
                  for (int qx = -1; qx <= 1; ++qx) {
                    const int ox = 2 * Dx + qx - px, oy = 2 * Dy + qy - py, oz = 2 * Dz + qz - pz;
                    if (std::abs(ox) > 1 || std::abs(oy) > 1 || std::abs(oz) > 1) continue;
                    acc += w1[px + 1] * w1[py + 1] * w1[pz + 1] * w1[qx + 1] * w1[qy + 1] * w1[qz + 1] * f27[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)];
                  }
        c27[(Dz + 1) * 9 + (Dy + 1) * 3 + (Dx + 1)] = acc;
      }
}

// colours of a 3d radius-1 stencil: 2 (7-point: (i + j + k) & 1) or 8 ((i & 1) + 2 (j & 1) + 4 (k & 1))
inline int colours_stencil3(const double *a27) {
  for (int dk = -1; dk <= 1; ++dk)
    for (int dj = -1; dj <= 1; ++dj)
      for (int di = -1; di <= 1; ++di)
        if (std::abs(di) + std::abs(dj) + std::abs(dk) > 1 && a27[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] != 0.0) return 8;
  return 2;
}

// one entry of the sparse n x m matrix B (or W): Euclidean vertex (i, j), column, value
struct SEntry {
  int i, j, col;
  double val;
};

// B_c = R B (linear_operator.cc:19)
inline std::vector<SEntry> coarsen_B(const std::vector<SEntry> &B, int nxf, int nyf) {
  const int nxc = nxf / 2, nyc = nyf / 2;
  std::map<std::pair<int, long long>, double> acc;  // (col, J * 2^20 + I)
  for (const SEntry &e : B) {
    for (int J = (e.j - 1 + 1) / 2; J <= (e.j + 1) / 2; ++J)
      for (int I = (e.i - 1 + 1) / 2; I <= (e.i + 1) / 2; ++I) {
        if (I < 1 || I > nxc - 1 || J < 1 || J > nyc - 1) continue;
        const int di = std::abs(e.i - 2 * I), dj = std::abs(e.j - 2 * J);
        if (di > 1 || dj > 1) continue;
        acc[{e.col, (long long)J * (1ll << 20) + I}] += (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * e.val;
      }
  }
  std::vector<SEntry> out;
  for (auto &kv : acc) out.push_back({int(kv.first.second % (1ll << 20)), int(kv.first.second >> 20), kv.first.first, kv.second});
  return out;
}

// 3d lattices: an entry of B / W stores the vertex (i, j, k) as (i, J) with J = k (ny + 1) + j, the row of the stacked planes
// (csrc/lattice3d.cuh).  B_c = R B with the trilinear full weighting (linear_operator.cc:19, intergrid_operator_linear.cc:8-30)
inline std::vector<SEntry> coarsen_B3(const std::vector<SEntry> &B, int nxf, int nyf, int nzf) {
  const int nxc = nxf / 2, nyc = nyf / 2, nzc = nzf / 2;
  std::map<std::pair<int, long long>, double> acc;  // (col, (K (nyc + 1) + J) * 2^20 + I)
  for (const SEntry &e : B) {
    const int ek = e.j / (nyf + 1), ej = e.j % (nyf + 1);
    for (int K = (ek - 1 + 1) / 2; K <= (ek + 1) / 2; ++K)
      for (int J = (ej - 1 + 1) / 2; J <= (ej + 1) / 2; ++J)
        for (int I = (e.i - 1 + 1) / 2; I <= (e.i + 1) / 2; ++I) {
          if (I < 1 || I > nxc - 1 || J < 1 || J > nyc - 1 || K < 1 || K > nzc - 1) continue;
          const int di = std::abs(e.i - 2 * I), dj = std::abs(ej - 2 * J), dk = std::abs(ek - 2 * K);
          if (di > 1 || dj > 1 || dk > 1) continue;
          acc[{e.col, (long long)(K * (nyc + 1) + J) * (1ll << 20) + I}] += (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0) * e.val;
        }
  }
  std::vector<SEntry> out;
  for (auto &kv : acc) out.push_back({int(kv.first.second % (1ll << 20)), int(kv.first.second >> 20), kv.first.first, kv.second});
  return out;
}

inline int site_colour3(int nc, int i, int j, int k) { return (nc == 2) ? ((i + j + k) & 1) : ((i & 1) + 2 * (j & 1) + 4 * (k & 1)); }

inline int site_colour(int nc, int i, int j) {
  if (nc == 2) return (i + j) & 1;
  if (nc == 4) return (i & 1) + 2 * (j & 1);
  return (i % 3) + 3 * (j % 3);
}

struct HostLevel {
  int nx = 0, ny = 0;
  int nz = 0;             // > 0: 3d lattice (Lattice3d, lattice/lattice3d.hh); the operator is st3, `st` only carries radius / ncolours
  double st3[27] = {0};   // 3d: uniform radius-1 stencil, entry (di, dj, dk) at [(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)]
  bool d3() const { return nz > 0; }
  StencilSet st;
  std::vector<SEntry> B;  // sorted by column
  // Variable coefficients (a correlation length that depends on x: PeriodicCorrelationLengthModel,
  // correlationlength_model.hh:83-113): one radius-1 stencil PER VERTEX, vc[k * np + j * (nx + 1) + i] with
  // k = (dj + 1) * 3 + (di + 1), np = (nx + 1) * (ny + 1).  Entries of boundary vertices and entries that point to a
  // boundary vertex are zero (the reference's matrices hold interior vertices only).  Empty: `st` is the operator.
  std::vector<double> vc;
  bool varcoef() const { return !vc.empty(); }
  long long ndof() const { return (long long)(nx - 1) * (ny - 1) * (nz > 0 ? nz - 1 : 1); }
  // matrix entry A[(i, j), (i + di, j + dj)] of an interior vertex (i, j)
  double coef(int i, int j, int di, int dj) const {
    if (vc.empty()) return st.at(pos_class(i, nx) + 3 * pos_class(j, ny), di, dj);
    if (di < -1 || di > 1 || dj < -1 || dj > 1) return 0.0;
    return vc[(size_t)((dj + 1) * 3 + (di + 1)) * (size_t)(nx + 1) * (ny + 1) + (size_t)j * (nx + 1) + i];
  }
};

// radius / number of colours of a per-vertex operator, as the oracle's scan of the matrix rows: which offsets occur at all
// (`st` of such a level only carries this classification)
inline void classify_varcoef(HostLevel &L) {
  const size_t np = (size_t)(L.nx + 1) * (L.ny + 1);
  std::memset(L.st.a, 0, sizeof(L.st.a));
  for (int k = 0; k < 9; ++k) {
    double big = 0.0;
    for (size_t v = 0; v < np; ++v) big = std::max(big, std::fabs(L.vc[(size_t)k * np + v]));
    for (int cls = 0; cls < 9; ++cls) L.st.a[cls][(k / 3 + 1) * 5 + (k % 3 + 1)] = big;
  }
  classify(L.st);
}

// ShiftedLaplaceFDOperator with kappa^2 given per interior vertex (lexicographic, shiftedlaplace_fd_operator.cc:33-56:
// diagonal h^d kappa^2(x) + sum_d 2 h^d / h_d^2 accumulated in the reference's order, off-diagonals -h^d / h_d^2 towards
// interior neighbours only)
inline void fine_varcoef(HostLevel &L, const double *kappa_sq) {
  const int nx = L.nx, ny = L.ny;
  const size_t np = (size_t)(nx + 1) * (ny + 1);
  L.vc.assign(9 * np, 0.0);
  const double hx = 1.0 / double(nx), hy = 1.0 / double(ny);
  const double hinv2x = 1.0 / (hx * hx), hinv2y = 1.0 / (hy * hy);
  const double vol = hx * hy;
  auto A = [&](int i, int j, int di, int dj) -> double & { return L.vc[(size_t)((dj + 1) * 3 + (di + 1)) * np + (size_t)j * (nx + 1) + i]; };
  for (int j = 1; j < ny; ++j)
    for (int i = 1; i < nx; ++i) {
      double diagonal = vol * kappa_sq[(size_t)(j - 1) * (nx - 1) + (i - 1)];
      diagonal += 2. * vol * hinv2x;
      diagonal += 2. * vol * hinv2y;
      A(i, j, 0, 0) = diagonal;
      if (i > 1) A(i, j, -1, 0) = -vol * hinv2x;
      if (i < nx - 1) A(i, j, +1, 0) = -vol * hinv2x;
      if (j > 1) A(i, j, 0, -1) = -vol * hinv2y;
      if (j < ny - 1) A(i, j, 0, +1) = -vol * hinv2y;
    }
  classify_varcoef(L);
}

// Galerkin product A_c = R A R^T (LinearOperator::coarsen, linear_operator.cc:12-15) of a per-vertex radius-1 operator
// with the un-normalised full-weighting R = {1/2, 1, 1/2} x {1/2, 1, 1/2}: again a per-vertex radius-1 (9-point) operator
inline void coarsen_varcoef(const HostLevel &f, HostLevel &c) {
  const int nxc = c.nx, nyc = c.ny;
  const size_t npc = (size_t)(nxc + 1) * (nyc + 1);
  c.vc.assign(9 * npc, 0.0);
  const double w1[3] = {0.5, 1.0, 0.5};
  for (int J = 1; J < nyc; ++J)
    for (int I = 1; I < nxc; ++I) {
      // t = (row (I, J) of R) A on the 5 x 5 fine window around (2 I, 2 J)
      double t[5][5];
      for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b) t[a][b] = 0.0;
      for (int py = -1; py <= 1; ++py)
        for (int px = -1; px <= 1; ++px) {
          const int pi = 2 * I + px, pj = 2 * J + py;  // always an interior fine vertex
          const double w = w1[px + 1] * w1[py + 1];
          for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di) t[py + dj + 2][px + di + 2] += w * f.coef(pi, pj, di, dj);
        }
      for (int Dy = -1; Dy <= 1; ++Dy)
        for (int Dx = -1; Dx <= 1; ++Dx) {
          if (I + Dx < 1 || I + Dx > nxc - 1 || J + Dy < 1 || J + Dy > nyc - 1) continue;
          double acc = 0.0;
          for (int qy = -1; qy <= 1; ++qy)
            for (int qx = -1; qx <= 1; ++qx) {
              const int oy = 2 * Dy + qy, ox = 2 * Dx + qx;
              if (ox < -2 || ox > 2 || oy < -2 || oy > 2) continue;
              acc += w1[qx + 1] * w1[qy + 1] * t[oy + 2][ox + 2];
            }
          c.vc[(size_t)((Dy + 1) * 3 + (Dx + 1)) * npc + (size_t)J * (nxc + 1) + I] = acc;
        }
    }
  classify_varcoef(c);
}

// general m x m inverse (Gauss-Jordan, partial pivoting); replaces Eigen's .inverse() (sor_smoother.cc:29,35)
inline std::vector<double> invert_dense(std::vector<double> A, int n) {
  std::vector<double> I((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) I[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)p * n + c])) p = r;
    if (A[(size_t)p * n + c] == 0.0) throw std::runtime_error("singular low-rank capacitance matrix");
    if (p != c)
      for (int k = 0; k < n; ++k) {
        std::swap(A[(size_t)p * n + k], A[(size_t)c * n + k]);
        std::swap(I[(size_t)p * n + k], I[(size_t)c * n + k]);
      }
    const double inv = 1.0 / A[(size_t)c * n + c];
    for (int k = 0; k < n; ++k) {
      A[(size_t)c * n + k] *= inv;
      I[(size_t)c * n + k] *= inv;
    }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = A[(size_t)r * n + c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) {
        A[(size_t)r * n + k] -= f * A[(size_t)c * n + k];
        I[(size_t)r * n + k] -= f * I[(size_t)c * n + k];
      }
    }
  }
  return I;
}

// Low-rank smoother data of one level and one sweep direction:
//   W = M_0^{-1} B with M_0 = D/omega + (strictly earlier part of A_0 in the colour ordering),
//   G = B^T W, K = (Sigma + G)^{-1}                                  (sor_smoother.cc:17-38)
// W is computed column by column as one colour sweep from x = 0 on a window around supp(B_k):
// in the colour ordering information travels at most (ncolours - 1) * radius sites per sweep.
struct LowRankDir {
  std::vector<SEntry> W;
  std::vector<double> G, K;     // m x m row-major
  std::vector<double> Mneg, Ms;  // -K and I - K G: d = Ms s + Mneg (B^T x)
};

inline void lowrank_finish(LowRankDir &out, const std::vector<double> &Sigma);
inline LowRankDir lowrank_setup3(const HostLevel &L, const std::vector<double> &Sigma, double omega, bool forward);

inline LowRankDir lowrank_setup(const HostLevel &L, const std::vector<double> &Sigma, double omega, bool forward) {
  if (L.d3()) return lowrank_setup3(L, Sigma, omega, forward);
  const int m = (int)Sigma.size();
  LowRankDir out;
  out.G.assign((size_t)m * m, 0.0);
  const int nc = L.st.ncolours, reach = (nc - 1) * L.st.radius, rad = L.st.radius;
  std::vector<std::vector<SEntry>> cols(m);
  for (const SEntry &e : L.B) cols[e.col].push_back(e);
  // W_k on its window (dense: a global measurement has the whole lattice as window)
  struct Win {
    int ilo = 0, jlo = 0, wx = 0, wy = 0;
    std::vector<double> v;
    double at(int i, int j) const {
      if (i < ilo || i >= ilo + wx || j < jlo || j >= jlo + wy) return 0.0;
      return v[(size_t)(j - jlo) * wx + (i - ilo)];
    }
  };
  std::vector<Win> Wwin(m);
  for (int k = 0; k < m; ++k) {
    if (cols[k].empty()) continue;
    int ilo = 1 << 30, ihi = -1, jlo = 1 << 30, jhi = -1;
    for (const SEntry &e : cols[k]) {
      ilo = std::min(ilo, e.i);
      ihi = std::max(ihi, e.i);
      jlo = std::min(jlo, e.j);
      jhi = std::max(jhi, e.j);
    }
    ilo = std::max(1, ilo - reach);
    jlo = std::max(1, jlo - reach);
    ihi = std::min(L.nx - 1, ihi + reach);
    jhi = std::min(L.ny - 1, jhi + reach);
    const int wx = ihi - ilo + 1, wy = jhi - jlo + 1;
    std::vector<double> xw((size_t)wx * wy, 0.0), bw((size_t)wx * wy, 0.0);
    for (const SEntry &e : cols[k]) bw[(size_t)(e.j - jlo) * wx + (e.i - ilo)] += e.val;
    for (int cc = 0; cc < nc; ++cc) {
      const int colour = forward ? cc : nc - 1 - cc;
      for (int j = jlo; j <= jhi; ++j)
        for (int i = ilo; i <= ihi; ++i) {
          if (site_colour(nc, i, j) != colour) continue;
          double s = 0.0;
          for (int dj = -rad; dj <= rad; ++dj)
            for (int di = -rad; di <= rad; ++di) {
              const int ii = i + di, jj = j + dj;
              if (ii < ilo || ii > ihi || jj < jlo || jj > jhi) continue;
              s += L.coef(i, j, di, dj) * xw[(size_t)(jj - jlo) * wx + (ii - ilo)];
            }
          double &xc = xw[(size_t)(j - jlo) * wx + (i - ilo)];
          xc += omega * (bw[(size_t)(j - jlo) * wx + (i - ilo)] - s) / L.coef(i, j, 0, 0);
        }
    }
    for (int j = jlo; j <= jhi; ++j)
      for (int i = ilo; i <= ihi; ++i) {
        const double v = xw[(size_t)(j - jlo) * wx + (i - ilo)];
        if (v != 0.0) out.W.push_back({i, j, k, v});
      }
    Wwin[k].ilo = ilo;
    Wwin[k].jlo = jlo;
    Wwin[k].wx = wx;
    Wwin[k].wy = wy;
    Wwin[k].v.swap(xw);
  }
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      double g = 0.0;
      if (!Wwin[b].v.empty())
        for (const SEntry &e : cols[a]) {
          const double w = Wwin[b].at(e.i, e.j);
          if (w != 0.0) g += e.val * w;
        }
      out.G[(size_t)a * m + b] = g;
    }
  lowrank_finish(out, Sigma);
  return out;
}

// K = (Sigma + G)^{-1} and the two m x m matrices of the fix-up, from G = B^T W
inline void lowrank_finish(LowRankDir &out, const std::vector<double> &Sigma) {
  const int m = (int)Sigma.size();
  std::vector<double> S((size_t)m * m, 0.0);
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) S[(size_t)a * m + b] = out.G[(size_t)a * m + b] + ((a == b) ? Sigma[a] : 0.0);
  out.K = invert_dense(S, m);
  out.Mneg.assign((size_t)m * m, 0.0);
  out.Ms.assign((size_t)m * m, 0.0);
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      double kg = 0.0;
      for (int c = 0; c < m; ++c) kg += out.K[(size_t)a * m + c] * out.G[(size_t)c * m + b];
      out.Mneg[(size_t)a * m + b] = -out.K[(size_t)a * m + b];
      out.Ms[(size_t)a * m + b] = ((a == b) ? 1.0 : 0.0) - kg;
    }
}

// 3d twin of lowrank_setup (uniform radius-1 stencil st3, 2 / 8 colours): W_k by one colour sweep from x = 0 on a 3d window
inline LowRankDir lowrank_setup3(const HostLevel &L, const std::vector<double> &Sigma, double omega, bool forward) {
  const int m = (int)Sigma.size();
  LowRankDir out;
  out.G.assign((size_t)m * m, 0.0);
  const int nc = L.st.ncolours, reach = nc - 1, pr = L.ny + 1;
  std::vector<std::vector<SEntry>> cols(m);
  for (const SEntry &e : L.B) cols[e.col].push_back(e);
  struct Win {
    int lo[3] = {0, 0, 0}, w[3] = {0, 0, 0};
    std::vector<double> v;
    double at(int i, int j, int k) const {
      if (i < lo[0] || i >= lo[0] + w[0] || j < lo[1] || j >= lo[1] + w[1] || k < lo[2] || k >= lo[2] + w[2]) return 0.0;
      return v[((size_t)(k - lo[2]) * w[1] + (j - lo[1])) * w[0] + (i - lo[0])];
    }
  };
  std::vector<Win> Wwin(m);
  const int n[3] = {L.nx, L.ny, L.nz};
  for (int q = 0; q < m; ++q) {
    if (cols[q].empty()) continue;
    int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
    for (const SEntry &e : cols[q]) {
      const int p[3] = {e.i, e.j % pr, e.j / pr};
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], p[d]);
        hi[d] = std::max(hi[d], p[d]);
      }
    }
    Win &W = Wwin[q];
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::max(1, lo[d] - reach);
      hi[d] = std::min(n[d] - 1, hi[d] + reach);
      W.lo[d] = lo[d];
      W.w[d] = hi[d] - lo[d] + 1;
    }
    const size_t nw = (size_t)W.w[0] * W.w[1] * W.w[2];
    std::vector<double> xw(nw, 0.0), bw(nw, 0.0);
    auto idx = [&](int i, int j, int k) { return ((size_t)(k - lo[2]) * W.w[1] + (j - lo[1])) * W.w[0] + (i - lo[0]); };
    for (const SEntry &e : cols[q]) bw[idx(e.i, e.j % pr, e.j / pr)] += e.val;
    for (int cc = 0; cc < nc; ++cc) {
      const int colour = forward ? cc : nc - 1 - cc;
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i) {
            if (site_colour3(nc, i, j, k) != colour) continue;
            double s = 0.0;
            for (int dk = -1; dk <= 1; ++dk)
              for (int dj = -1; dj <= 1; ++dj)
                for (int di = -1; di <= 1; ++di) {
                  const int ii = i + di, jj = j + dj, kk = k + dk;
                  if (ii < lo[0] || ii > hi[0] || jj < lo[1] || jj > hi[1] || kk < lo[2] || kk > hi[2]) continue;
                  s += L.st3[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] * xw[idx(ii, jj, kk)];
                }
            xw[idx(i, j, k)] += omega * (bw[idx(i, j, k)] - s) / L.st3[13];
          }
    }
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          const double v = xw[idx(i, j, k)];
          if (v != 0.0) out.W.push_back({i, k * pr + j, q, v});
        }
    W.v.swap(xw);
  }
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      double g = 0.0;
      if (!Wwin[b].v.empty())
        for (const SEntry &e : cols[a]) {
          const double w = Wwin[b].at(e.i, e.j % pr, e.j / pr);
          if (w != 0.0) g += e.val * w;
        }
      out.G[(size_t)a * m + b] = g;
    }
  lowrank_finish(out, Sigma);
  return out;
}

// true if the measurements do not interact on this level in either sweep direction: B^T W is diagonal and every
// site of supp(W_j) keeps a distance > `margin` from the bounding box of supp(B_k), k != j.  Then the Woodbury
// fix-up of measurement k only needs x near supp(B_k) (in-kernel fix-up without exchange, row strips).
inline bool lowrank_is_diagonal(const HostLevel &L, const std::vector<double> &Sigma, double omega, int margin = 8) {
  const int m = (int)Sigma.size();
  std::vector<int> bb((size_t)m * 4);
  for (int k = 0; k < m; ++k) {
    bb[4 * k] = bb[4 * k + 2] = 1 << 30;
    bb[4 * k + 1] = bb[4 * k + 3] = -(1 << 30);
  }
  for (const SEntry &e : L.B) {
    bb[4 * e.col] = std::min(bb[4 * e.col], e.i);
    bb[4 * e.col + 1] = std::max(bb[4 * e.col + 1], e.i);
    bb[4 * e.col + 2] = std::min(bb[4 * e.col + 2], e.j);
    bb[4 * e.col + 3] = std::max(bb[4 * e.col + 3], e.j);
  }
  for (int dir = 0; dir < 2; ++dir) {
    const LowRankDir h = lowrank_setup(L, Sigma, omega, dir == 0);
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b)
        if (a != b && (h.Mneg[(size_t)a * m + b] != 0.0 || h.Ms[(size_t)a * m + b] != 0.0)) return false;
    for (const SEntry &e : h.W)
      for (int k = 0; k < m; ++k)
        if (k != e.col && e.i >= bb[4 * k] - margin && e.i <= bb[4 * k + 1] + margin && e.j >= bb[4 * k + 2] - margin && e.j <= bb[4 * k + 3] + margin) return false;
  }
  return true;
}

// dense matrix of the coarsest level, A_0 + B Sigma^{-1} B^T (cholesky_sampler.cc:25-38), its lower
// Cholesky factor padded with identity to Np = multiple of 32, and the inverses of the diagonal blocks
struct CoarseFactor {
  int N = 0, Np = 0;
  std::vector<double> L;     // Np x Np row-major lower Cholesky factor (kept for inspection)
  std::vector<double> T;     // L^{-1}, Np x Np row-major (lower triangular)
  std::vector<double> TT;    // L^{-T}, Np x Np row-major (upper triangular)
};

inline CoarseFactor coarse_factor(const HostLevel &Lv, const std::vector<double> &Sigma) {
  CoarseFactor cf;
  const int w = Lv.nx - 1, h = Lv.ny - 1;
  const int N = (int)Lv.ndof(), Np = ((N + 31) / 32) * 32;
  cf.N = N;
  cf.Np = Np;
  std::vector<double> &A = cf.L;
  A.assign((size_t)Np * Np, 0.0);
  if (Lv.d3()) {
    const int d = Lv.nz - 1;
    for (int k = 1; k <= d; ++k)
      for (int j = 1; j <= h; ++j)
        for (int i = 1; i <= w; ++i) {
          const int row = ((k - 1) * h + (j - 1)) * w + (i - 1);
          for (int dk = -1; dk <= 1; ++dk)
            for (int dj = -1; dj <= 1; ++dj)
              for (int di = -1; di <= 1; ++di) {
                const int ii = i + di, jj = j + dj, kk = k + dk;
                if (ii < 1 || ii > w || jj < 1 || jj > h || kk < 1 || kk > d) continue;
                const double v = Lv.st3[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)];
                if (v != 0.0) A[(size_t)row * Np + ((kk - 1) * h + (jj - 1)) * w + (ii - 1)] = v;
              }
        }
  }
  for (int j = 1; j <= h && !Lv.d3(); ++j)
    for (int i = 1; i <= w; ++i) {
      const int row = (j - 1) * w + (i - 1);
      for (int dj = -2; dj <= 2; ++dj)
        for (int di = -2; di <= 2; ++di) {
          const int ii = i + di, jj = j + dj;
          if (ii < 1 || ii > w || jj < 1 || jj > h) continue;
          const double v = Lv.coef(i, j, di, dj);
          if (v != 0.0) A[(size_t)row * Np + (jj - 1) * w + (ii - 1)] = v;
        }
    }
  const int m = (int)Sigma.size();
  std::vector<std::vector<SEntry>> cols(m);
  for (const SEntry &e : Lv.B) cols[e.col].push_back(e);
  for (int k = 0; k < m; ++k)
    for (const SEntry &p : cols[k])
      for (const SEntry &q : cols[k]) {
        // (3d: j is the row of the stacked planes, k_z (ny + 1) + j)
        const int pr = Lv.ny + 1;
        const int rp = Lv.d3() ? ((p.j / pr - 1) * h + (p.j % pr - 1)) * w + (p.i - 1) : (p.j - 1) * w + (p.i - 1);
        const int rq = Lv.d3() ? ((q.j / pr - 1) * h + (q.j % pr - 1)) * w + (q.i - 1) : (q.j - 1) * w + (q.i - 1);
        A[(size_t)rp * Np + rq] += p.val * q.val / Sigma[k];
      }
  for (int r = N; r < Np; ++r) A[(size_t)r * Np + r] = 1.0;
  // in-place lower Cholesky
  for (int j = 0; j < Np; ++j) {
    double d = A[(size_t)j * Np + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * Np + k] * A[(size_t)j * Np + k];
    if (!(d > 0.0)) throw std::runtime_error("coarse matrix is not positive definite");
    d = std::sqrt(d);
    A[(size_t)j * Np + j] = d;
    for (int i = j + 1; i < Np; ++i) {
      double s = A[(size_t)i * Np + j];
      const double *ai = &A[(size_t)i * Np], *aj = &A[(size_t)j * Np];
      for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[(size_t)i * Np + j] = s / d;
    }
    for (int k = j + 1; k < Np; ++k) A[(size_t)j * Np + k] = 0.0;
  }
  // T = L^{-1} column by column (forward substitution); TT = T^T
  cf.T.assign((size_t)Np * Np, 0.0);
  cf.TT.assign((size_t)Np * Np, 0.0);
  std::vector<double> col(Np);
  for (int c = 0; c < Np; ++c) {
    for (int r = c; r < Np; ++r) {
      double s = (r == c) ? 1.0 : 0.0;
      const double *lr = &A[(size_t)r * Np];
      for (int k = c; k < r; ++k) s -= lr[k] * col[k];
      col[r] = s / lr[r];
    }
    for (int r = c; r < Np; ++r) {
      cf.T[(size_t)r * Np + c] = col[r];
      cf.TT[(size_t)c * Np + r] = col[r];
    }
  }
  return cf;
}

}  // namespace mgmc
