// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_core.hh).  Implementation of the CPU restatement.
#include "oracle_core.hh"

#include <atomic>
#include <thread>

namespace orc {

// ---------------------------------------------------------------------------------------------
// shiftedlaplace_fd_operator.cc:9-56 :  h^d ( kappa^2 + sum_d 2/h_d^2 ) on the diagonal,
// -h^d/h_d^2 off-diagonal, rows truncated at the Dirichlet boundary
// ---------------------------------------------------------------------------------------------
LinearOperator make_shiftedlaplace_fd(const Lattice &lattice, const KappaModel &km) {
  const int dim = lattice.dim;
  const long nrow = lattice.Nvertex();
  double hinv2[3], cell_volume = 1.0;
  for (int d = 0; d < dim; ++d) {
    const double h = 1. / double(lattice.n[d]);
    hinv2[d] = 1. / (h * h);
    cell_volume *= h;
  }
  std::vector<Triplet> t;
  t.reserve((1 + 2 * dim) * nrow);
  for (long ell = 0; ell < nrow; ++ell) {
    double x[3];
    lattice.vertex_coordinates(ell, x);
    double diagonal = cell_volume * km.kappa_sq(x, dim);
    for (int d = 0; d < dim; ++d) {
      for (int j = 0; j < 2; ++j) {
        int shift[3] = {0, 0, 0};
        shift[d] = 2 * j - 1;
        long es;
        if (lattice.shifted_vertex_is_internal(ell, shift, es)) t.push_back({ell, es, -cell_volume * hinv2[d]});
      }
      diagonal += 2. * cell_volume * hinv2[d];
    }
    t.push_back({ell, ell, diagonal});
  }
  LinearOperator op;
  op.lattice = lattice;
  op.A = CSR::from_triplets(nrow, nrow, t);
  return op;
}

// ---------------------------------------------------------------------------------------------
// squared_shiftedlaplace_fd_operator.cc:9-96 (2d only): 13-point diamond, boundary rule
// "missing +-1 neighbour => add the +-2 coefficient to the diagonal"
// ---------------------------------------------------------------------------------------------
LinearOperator make_squared_shiftedlaplace_fd(const Lattice &lattice, const KappaModel &km) {
  if (lattice.dim != 2) throw std::runtime_error("SquaredShiftedLaplaceFDOperator only implemented for d=2");
  const long nrow = lattice.Nvertex();
  double h[2], hinv2[2], cell_volume = 1.0;
  for (int d = 0; d < 2; ++d) {
    h[d] = 1. / double(lattice.n[d]);
    hinv2[d] = 1. / (h[d] * h[d]);
    cell_volume *= h[d];
  }
  double sl[2][2] = {{0, 0}, {0, 0}};
  sl[0][0] = -2 * (hinv2[0] + hinv2[1]);
  sl[1][0] = hinv2[0];
  sl[0][1] = hinv2[1];
  double ss[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  ss[0][0] = 6 * (hinv2[0] * hinv2[0] + hinv2[1] * hinv2[1]) + 8 * hinv2[0] * hinv2[1];
  ss[1][0] = -4 * hinv2[0] * (hinv2[0] + hinv2[1]);
  ss[0][1] = -4 * hinv2[1] * (hinv2[0] + hinv2[1]);
  ss[2][0] = hinv2[0] * hinv2[0];
  ss[0][2] = hinv2[1] * hinv2[1];
  ss[1][1] = 2 * hinv2[0] * hinv2[1];
  std::vector<Triplet> t;
  t.reserve(13 * nrow);
  for (long ell = 0; ell < nrow; ++ell) {
    double x[3];
    lattice.vertex_coordinates(ell, x);
    const double alpha_b = km.kappa_sq(x, 2);
    double diagonal = (alpha_b * alpha_b - 2. * alpha_b * sl[0][0] + ss[0][0]) * cell_volume;
    for (int j = -2; j <= 2; ++j)
      for (int k = -2; k <= 2; ++k) {
        if ((std::abs(j) + std::abs(k) > 2) || ((j == 0) && (k == 0))) continue;
        int shift[3] = {j, k, 0};
        long es;
        if (lattice.shifted_vertex_is_internal(ell, shift, es)) {
          double e = ss[std::abs(j)][std::abs(k)];
          if (std::abs(j) + std::abs(k) == 1) e += -2. * alpha_b * sl[std::abs(j)][std::abs(k)];
          t.push_back({ell, es, e * cell_volume});
        } else if (std::abs(j) + std::abs(k) == 1) {
          diagonal += ss[2 * std::abs(j)][2 * std::abs(k)] * cell_volume;
        }
      }
    t.push_back({ell, ell, diagonal});
  }
  LinearOperator op;
  op.lattice = lattice;
  op.A = CSR::from_triplets(nrow, nrow, t);
  return op;
}

// ---------------------------------------------------------------------------------------------
// shiftedlaplace_fem_operator.cc:9-187: multilinear (Q1) FEM with 2-point Gauss quadrature
// ---------------------------------------------------------------------------------------------
static double fem_phi(const std::vector<int> &alpha, const std::vector<double> &xhat) {  // :155-164
  double p = 1.0;
  for (size_t j = 0; j < alpha.size(); ++j) p *= (alpha[j] == 0) ? (1.0 - xhat[j]) : xhat[j];
  return p;
}
static std::vector<double> fem_grad_phi(const std::vector<int> &alpha, const std::vector<double> &xhat) {  // :167-187
  const size_t dim = alpha.size();
  std::vector<double> g(dim);
  for (size_t k = 0; k < dim; ++k) {
    double v = 1.0;
    for (size_t j = 0; j < dim; ++j) {
      if (j == k)
        v *= (alpha[j] == 0) ? -1.0 : +1.0;
      else
        v *= (alpha[j] == 0) ? (1.0 - xhat[j]) : xhat[j];
    }
    g[k] = v;
  }
  return g;
}

LinearOperator make_shiftedlaplace_fem(const Lattice &lattice, const KappaModel &km) {
  const int dim = lattice.dim;
  const long nrow = lattice.Nvertex();
  double h[3], hinv2[3], cell_volume = 1.0;
  for (int d = 0; d < dim; ++d) {
    h[d] = 1. / double(lattice.n[d]);
    hinv2[d] = 1. / (h[d] * h[d]);
    cell_volume *= h[d];
  }
  // STEP 1 (:41-63): sparsity pattern = all 3^d shifts that stay interior
  std::vector<Triplet> t;
  std::vector<std::vector<int>> shifts = cartesian_product(std::vector<int>{-1, 0, +1}, dim);
  for (long er = 0; er < nrow; ++er)
    for (auto &s : shifts) {
      int sh[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d) sh[d] = s[d];
      long ec;
      if (lattice.shifted_vertex_is_internal(er, sh, ec)) t.push_back({er, ec, 0.0});
    }
  LinearOperator op;
  op.lattice = lattice;
  op.A = CSR::from_triplets(nrow, nrow, t);
  // STEP 2 (:65-138): assembly
  GaussLegendreQuadrature quad(dim, 1);
  std::vector<std::vector<int>> basis_idx = cartesian_product(std::vector<int>{0, 1}, dim);
  std::vector<double> phi_phi, gradphi_gradphi;
  for (auto &alpha : basis_idx)
    for (auto &beta : basis_idx)
      for (size_t j = 0; j < quad.points.size(); ++j) {
        const std::vector<double> &xhat = quad.points[j];
        phi_phi.push_back(fem_phi(alpha, xhat) * fem_phi(beta, xhat));
        std::vector<double> ga = fem_grad_phi(alpha, xhat), gb = fem_grad_phi(beta, xhat);
        double s = 0.0;
        for (int d = 0; d < dim; ++d) s += ga[d] * hinv2[d] * gb[d];
        gradphi_gradphi.push_back(s);
      }
  const long ncell = lattice.Ncell();
  for (long cell = 0; cell < ncell; ++cell) {
    size_t count = 0;
    int cc[3];
    lattice.cell_l2e(cell, cc);
    for (auto &alpha : basis_idx)
      for (auto &beta : basis_idx) {
        long er, ec;
        int a3[3] = {0, 0, 0}, b3[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d) {
          a3[d] = alpha[d];
          b3[d] = beta[d];
        }
        if (lattice.corner_is_internal_vertex(cell, a3, er) && lattice.corner_is_internal_vertex(cell, b3, ec)) {
          double local = 0.0;
          for (size_t j = 0; j < quad.points.size(); ++j) {
            double x[3];
            for (int d = 0; d < dim; ++d) x[d] = h[d] * (quad.points[j][d] + double(cc[d]));
            local += (km.kappa_sq(x, dim) * phi_phi[count] + gradphi_gradphi[count]) * quad.weights[j];
            count++;
          }
          op.A.coeff_ref(er, ec) += local * cell_volume;
        } else {
          count += quad.points.size();
        }
      }
  }
  return op;
}

// sampler/test_sampler.hh:30-67
LinearOperator make_test_operator_1d(bool lowrank) {
  int n8[1] = {8};
  Lattice lat(1, n8);
  const long nrow = lat.Nvertex();
  std::vector<Triplet> t;
  for (long i = 0; i < nrow; ++i) {
    t.push_back({i, i, +6.0});
    if (i > 0) t.push_back({i, (i - 1 + nrow) % nrow, -1.0});
    if (i < nrow - 1) t.push_back({i, (i + 1 + nrow) % nrow, -1.0});
  }
  LinearOperator op;
  op.lattice = lat;
  op.A = CSR::from_triplets(nrow, nrow, t);
  if (lowrank) {
    op.m_lowrank = 2;
    std::vector<Triplet> b = {{3, 0, 10.0}, {4, 1, 10.0}};
    op.set_B(CSR::from_triplets(nrow, 2, b));
    op.Sigma = {4.2, 9.3};
  }
  return op;
}

// ---------------------------------------------------------------------------------------------
// measured_operator.cc
// ---------------------------------------------------------------------------------------------
static double V_sphere(double radius, int dim) {  // :52-66
  if (dim == 0) return 1.0;
  if (dim == 1) return 2. * radius;
  return 2. * M_PI / double(dim) * radius * radius * V_sphere(radius, dim - 2);
}

void measurement_vector(const Lattice &lattice, const double *x0, double radius, std::vector<long> &idx, std::vector<double> &val) {
  idx.clear();
  val.clear();
  const int dim = lattice.dim;
  const long nv = lattice.Nvertex();
  if (radius < 1.E-12) {  // :74-91 nearest vertex (first minimum in lexicographic order, strict <)
    double d_min = double(dim);
    long ell_min = 0;
    for (long ell = 0; ell < nv; ++ell) {
      double x[3], d2 = 0.0;
      lattice.vertex_coordinates(ell, x);
      for (int d = 0; d < dim; ++d) d2 += (x[d] - x0[d]) * (x[d] - x0[d]);
      const double dist = std::sqrt(d2);
      if (dist < d_min) {
        d_min = dist;
        ell_min = ell;
      }
    }
    idx.push_back(ell_min);
    val.push_back(1.0);
    return;
  }
  // :92-168 ball average against the bilinear hat functions
  double h[3];
  const double cell_volume = lattice.cell_volume();
  const double normalisation = 1. / V_sphere(radius, dim);
  for (int d = 0; d < dim; ++d) h[d] = 1. / double(lattice.n[d]);
  GaussLegendreQuadrature quad(dim, 1);
  std::vector<std::vector<int>> basis_idx = cartesian_product(std::vector<int>{0, 1}, dim);
  std::vector<double> dense(nv, 0.0);
  std::vector<char> used(nv, 0);
  const long ncell = lattice.Ncell();
  for (long cell = 0; cell < ncell; ++cell) {
    int cc[3];
    lattice.cell_l2e(cell, cc);
    bool overlap = false;
    double cmin[3] = {2.0, 2.0, 2.0}, cmax[3] = {-1.0, -1.0, -1.0};
    for (auto &om : basis_idx) {
      double d2 = 0.0;
      for (int d = 0; d < dim; ++d) {
        const double xc = h[d] * double(cc[d] + om[d]);
        cmin[d] = std::min(cmin[d], xc);
        cmax[d] = std::max(cmax[d], xc);
        d2 += (xc - x0[d]) * (xc - x0[d]);
      }
      overlap = overlap || (std::sqrt(d2) < radius);
    }
    bool centre_in_cell = true;
    for (int d = 0; d < dim; ++d) centre_in_cell = centre_in_cell && (cmin[d] <= x0[d]) && (x0[d] <= cmax[d]);
    overlap = overlap || centre_in_cell;
    if (!overlap) continue;
    for (auto &alpha : basis_idx) {
      long ell;
      int a3[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d) a3[d] = alpha[d];
      if (lattice.corner_is_internal_vertex(cell, a3, ell)) {
        double local = 0.0;
        for (size_t j = 0; j < quad.points.size(); ++j) {
          const std::vector<double> &xhat = quad.points[j];
          double d2 = 0.0;
          for (int d = 0; d < dim; ++d) {
            const double x = h[d] * (xhat[d] + double(cc[d]));
            d2 += (x - x0[d]) * (x - x0[d]);
          }
          const double xi = std::sqrt(d2) / radius;
          if (xi < 1.0) {
            double phihat = 1.0;  // f_meas(xi) = 1 (measured_operator.hh:69)
            for (int d = 0; d < dim; ++d) phihat *= (alpha[d] == 0) ? (1.0 - xhat[d]) : xhat[d];
            local += phihat * quad.weights[j] * cell_volume * normalisation;
          }
        }
        dense[ell] += local;  // coeffRef creates the entry even if local == 0
        used[ell] = 1;
      }
    }
  }
  for (long ell = 0; ell < nv; ++ell)
    if (used[ell]) {
      idx.push_back(ell);
      val.push_back(dense[ell]);
    }
}

LinearOperator make_measured_operator(const LinearOperator &base, int n_meas, const double *locations, const double *variance_scaled,
                                      double radius, bool measure_global, double variance_global) {
  LinearOperator op;
  op.lattice = base.lattice;
  op.A = base.A;
  op.m_lowrank = n_meas + (measure_global ? 1 : 0);
  const long nrow = base.lattice.Nvertex();
  const int dim = base.lattice.dim;
  op.Sigma.assign(op.m_lowrank, 0.0);
  for (int k = 0; k < n_meas; ++k) op.Sigma[k] = variance_scaled[k];
  std::vector<Triplet> t;
  for (int k = 0; k < n_meas; ++k) {
    std::vector<long> idx;
    std::vector<double> val;
    measurement_vector(base.lattice, locations + (long)k * dim, radius, idx, val);
    for (size_t q = 0; q < idx.size(); ++q) t.push_back({idx[q], (long)k, val[q]});
  }
  if (measure_global) {  // :31-46
    const double cell_volume = base.lattice.cell_volume();
    for (long ell = 0; ell < nrow; ++ell) t.push_back({ell, (long)n_meas, cell_volume});
    op.Sigma[n_meas] = variance_global;
  }
  if (op.m_lowrank > 0) op.set_B(CSR::from_triplets(nrow, op.m_lowrank, t));
  return op;
}

// ---------------------------------------------------------------------------------------------
// orderings
// ---------------------------------------------------------------------------------------------
// 3d (radius-1 operators only): 2 colours for the 7-point operator, 8 for operators that couple diagonal neighbours
static int colour_count_3d(const Lattice &lat, const CSR &A) {
  const long w = lat.n[0] - 1, h = lat.n[1] - 1;
  bool diag_coupling = false;
  for (long r = 0; r < A.rows; ++r) {
    const long i = r % w, j = (r / w) % h, k = r / (w * h);
    for (long q = A.rowptr[r]; q < A.rowptr[r + 1]; ++q) {
      if (A.val[q] == 0.0) continue;
      const long c = A.col[q];
      const int di = (int)std::labs(c % w - i), dj = (int)std::labs((c / w) % h - j), dk = (int)std::labs(c / (w * h) - k);
      if (std::max(di, std::max(dj, dk)) > 1) throw std::runtime_error("colour ordering in 3d: radius-1 operators only");
      if ((di > 0) + (dj > 0) + (dk > 0) > 1) diag_coupling = true;
    }
  }
  return diag_coupling ? 8 : 2;
}

int colour_count_2d(const Lattice &lat, const CSR &A) {
  if (lat.dim == 3) return colour_count_3d(lat, A);
  if (lat.dim != 2) throw std::runtime_error("colour ordering implemented for 2d and 3d lattices only");
  const long w = lat.n[0] - 1;
  int radius = 0;
  bool diag_coupling = false;
  for (long r = 0; r < A.rows; ++r) {
    const long i = r % w, j = r / w;
    for (long k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) {
      if (A.val[k] == 0.0) continue;
      const long ci = A.col[k] % w, cj = A.col[k] / w;
      const int di = (int)std::labs(ci - i), dj = (int)std::labs(cj - j);
      radius = std::max(radius, std::max(di, dj));
      if (di > 0 && dj > 0) diag_coupling = true;
    }
  }
  if (radius >= 2) return 9;
  return diag_coupling ? 4 : 2;
}

std::vector<long> make_order(const Lattice &lat, const CSR &A, int ordering) {
  const long n = lat.Nvertex();
  std::vector<long> order(n);
  for (long k = 0; k < n; ++k) order[k] = k;
  if (ordering == ORDER_LEX) return order;
  const int nc = colour_count_2d(lat, A);
  std::vector<int> colour(n);
  for (long ell = 0; ell < n; ++ell) {
    int idx[3];
    lat.vertex_l2e(ell, idx);
    const int i = idx[0], j = idx[1];
    if (lat.dim == 3)  // the B200 path's 3d colourings (csrc/lattice3d.cuh)
      colour[ell] = (nc == 2) ? ((i + j + idx[2]) & 1) : ((i & 1) + 2 * (j & 1) + 4 * (idx[2] & 1));
    else if (nc == 2)
      colour[ell] = (i + j) & 1;
    else if (nc == 4)
      colour[ell] = (i & 1) + 2 * (j & 1);
    else
      colour[ell] = (i % 3) + 3 * (j % 3);
  }
  std::stable_sort(order.begin(), order.end(), [&](long a, long b) { return colour[a] < colour[b]; });
  return order;
}

// ---------------------------------------------------------------------------------------------
// SORSmoother
// ---------------------------------------------------------------------------------------------
// SET-UP helper: independent pieces of the low-rank smoother data (one per measurement) on the host's threads.
// Nothing on the sampling / smoothing path uses it.
template <class F>
static void setup_parallel_for(int n, F &&body) {
  const int nt = (int)std::max(1u, std::min((unsigned)n, std::thread::hardware_concurrency()));
  if (nt <= 1) {
    for (int k = 0; k < n; ++k) body(k);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&] {
      for (int k = next++; k < n; k = next++) body(k);
    });
  for (auto &t : th) t.join();
}

SORSmoother::SORSmoother(const LinearOperator *op_, double omega_, int nsmooth_, Direction dir_, const std::vector<long> &order_)
    : op(op_), omega(omega_), nsmooth(nsmooth_), direction(dir_), order(order_), diag(op_->A.diagonal()) {
  const int m = op->m_lowrank;
  if (m > 0) {
    // sor_smoother.cc:17-38.  (L + D/omega)^{-1} b equals one forward SOR sweep on A_0 started
    // from x = 0 with right-hand side b (and (L^T + D/omega)^{-1} b one backward sweep); this is
    // the triangular solve of the reference expressed for an arbitrary visiting order.
    const long n = op->ndof();
    // (SET-UP only is threaded -- the m columns are independent; the sampling / smoothing path below stays on one
    //  thread like the reference: one sequential RNG stream, lexicographic Gauss-Seidel)
    std::vector<Vec> W(m, Vec(n, 0.0));
    setup_parallel_for(m, [&](int k) {
      Vec col(n, 0.0);
      for (long p = op->BT.rowptr[k]; p < op->BT.rowptr[k + 1]; ++p) col[op->BT.col[p]] = op->BT.val[p];
      sweep_once(col.data(), W[k].data());
    });
    std::vector<double> S(m * m, 0.0);
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b) {
        double s = (a == b) ? op->Sigma[a] : 0.0;
        for (long p = op->BT.rowptr[a]; p < op->BT.rowptr[a + 1]; ++p) s += op->BT.val[p] * W[b][op->BT.col[p]];
        S[a * m + b] = s;
      }
    std::vector<double> Sinv = dense_inverse(S, m);
    B_bar.assign(m, Vec(n, 0.0));
    setup_parallel_for(m, [&](int b) {
      for (int a = 0; a < m; ++a) {
        const double s = Sinv[a * m + b];
        if (s == 0.0) continue;
        const Vec &w = W[a];
        Vec &o = B_bar[b];
        for (long i = 0; i < n; ++i) o[i] += w[i] * s;
      }
    });
  }
}

void SORSmoother::sweep_once(const double *b, double *x) const {
  const CSR &A = op->A;
  const long nrow = A.rows;
  for (long e_ = 0; e_ < nrow; ++e_) {
    const long ell = order[(direction == forward) ? e_ : nrow - 1 - e_];
    double residual = 0.0;
    for (long k = A.rowptr[ell]; k < A.rowptr[ell + 1]; ++k) residual += A.val[k] * x[A.col[k]];
    x[ell] += omega * (b[ell] - residual) / diag[ell];
  }
}

void SORSmoother::apply_sparse(const double *b, double *x) const {
  for (int k = 0; k < nsmooth; ++k) sweep_once(b, x);  // sor_smoother.cc:64 (inner nsmooth loop: k^2 quirk)
}

void SORSmoother::apply(const double *b, double *x) const {
  const int m = op->m_lowrank;
  const long n = op->ndof();
  for (int k = 0; k < nsmooth; ++k) {
    apply_sparse(b, x);
    if (m > 0) {
      Vec BT_x(m);
      op->BT.matvec(x, BT_x.data());
      for (int c = 0; c < m; ++c) {
        const double s = BT_x[c];
        const Vec &bb = B_bar[c];
        for (long i = 0; i < n; ++i) x[i] -= bb[i] * s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SORSampler (sor_sampler.cc)
// ---------------------------------------------------------------------------------------------
SORSampler::SORSampler(const LinearOperator *op_, NoiseSource noise_, double omega_, int nsmooth_, Direction dir_,
                       const std::vector<long> &order, int level_)
    : op(op_), noise(noise_), omega(omega_), direction(dir_), nsmooth(nsmooth_), level(level_),
      smoother(op_, omega_, 1, dir_, order), c_rhs(op_->ndof()), xi(op_->m_lowrank) {
  const Vec diag = op->A.diagonal();
  sqrt_precision_diag.resize(diag.size());
  for (size_t e = 0; e < diag.size(); ++e) sqrt_precision_diag[e] = std::sqrt(diag[e] * (2. - omega) / omega);
  Sigma_inv_sqrt.resize(op->m_lowrank);
  for (int k = 0; k < op->m_lowrank; ++k) Sigma_inv_sqrt[k] = std::sqrt(1.0 / op->Sigma[k]);
}

void SORSampler::apply(const double *f, double *x) const {
  const long n = op->ndof();
  const int m = op->m_lowrank;
  for (int k = 0; k < nsmooth; ++k) {
    if (!noise.philox()) {
      for (long ell = 0; ell < n; ++ell) c_rhs[ell] = sqrt_precision_diag[ell] * noise.dist(*noise.engine) + f[ell];
      if (m > 0)
        for (int q = 0; q < m; ++q) xi[q] = noise.dist(*noise.engine);
    } else {
      PhiloxCtx &px = *noise.px;
      if (op->lattice.dim != 2 && op->lattice.dim != 3) throw std::runtime_error("philox noise implemented for 2d and 3d lattices only");
      const uint32_t c1 = ((uint32_t)level << 24) | (px.sweep_counter[level]++ & 0xFFFFFFu);
      const int w = op->lattice.n[0] - 1;
      const uint32_t G = (uint32_t)(op->lattice.n[0] / 4 + 1);
      // 3d: the B200 path stacks the planes k = 0 .. nz in the row direction, row = k (ny + 1) + j (csrc/lattice3d.cuh)
      const bool d3 = (op->lattice.dim == 3);
      const long h = d3 ? op->lattice.n[1] - 1 : 0;
      for (long ell = 0; ell < n; ++ell) {
        const uint32_t i = (uint32_t)(ell % w) + 1;
        const uint32_t j = d3 ? (uint32_t)(((ell / w) / h + 1) * (op->lattice.n[1] + 1) + (ell / w) % h + 1) : (uint32_t)(ell / w) + 1;
        double z0, z1;
        Philox::normal_pair(px.seed, ((j * G + (i >> 2)) << 1) | (i & 1u), c1, px.sample, px.chain, z0, z1);
        c_rhs[ell] = sqrt_precision_diag[ell] * ((i & 2u) ? z1 : z0) + f[ell];
      }
      for (int q = 0; q < m; ++q) {
        double z0, z1;
        Philox::normal_pair(px.seed, 0x80000000u | ((uint32_t)q >> 1), c1, px.sample, px.chain, z0, z1);
        xi[q] = (q & 1) ? z1 : z0;
      }
    }
    if (m > 0) {  // sor_sampler.cc:48-56 : c += B Sigma^{-1/2} xi
      for (long r = 0; r < n; ++r)
        for (long q = op->B.rowptr[r]; q < op->B.rowptr[r + 1]; ++q) c_rhs[r] += op->B.val[q] * Sigma_inv_sqrt[op->B.col[q]] * xi[op->B.col[q]];
    }
    smoother.apply(c_rhs.data(), x);
  }
}

// ---------------------------------------------------------------------------------------------
// Cholesky sampler / solver
// ---------------------------------------------------------------------------------------------
CholeskySampler::CholeskySampler(const LinearOperator *op_, NoiseSource noise_, int level_)
    : op(op_), noise(noise_), level(level_), n(op_->ndof()), L(op_->precision()), xi(n), g(n) {
  dense_cholesky(L, n);  // cholesky_sampler.cc:25-38 folds B Sigma^{-1} B^T into the matrix
}

void CholeskySampler::apply(const double *f, double *x) const {
  if (!noise.philox()) {
    for (long ell = 0; ell < n; ++ell) xi[ell] = noise.dist(*noise.engine);
  } else {
    PhiloxCtx &px = *noise.px;
    const uint32_t c1 = ((uint32_t)level << 24) | (px.sweep_counter[level]++ & 0xFFFFFFu);
    for (long ell = 0; ell < n; ++ell) {
      double z0, z1;
      Philox::normal_pair(px.seed, 0x40000000u | ((uint32_t)ell >> 1), c1, px.sample, px.chain, z0, z1);
      xi[ell] = (ell & 1) ? z1 : z0;
    }
  }
  const double *gp;
  if (rhs_fixed) {
    gp = g_rhs.data();
  } else {
    dense_solveL(L, n, f, g.data());
    gp = g.data();
  }
  Vec rhs(n);
  for (long i = 0; i < n; ++i) rhs[i] = xi[i] + gp[i];
  dense_solveLT(L, n, rhs.data(), x);
}

void CholeskySampler::fix_rhs(const double *f) {
  g_rhs.resize(n);
  dense_solveL(L, n, f, g_rhs.data());
  rhs_fixed = true;
}

CholeskySolver::CholeskySolver(const LinearOperator *op_) : op(op_), n(op_->ndof()), L(op_->A.to_dense()) {
  dense_cholesky(L, n);
  const int m = op->m_lowrank;
  if (m > 0) {  // cholesky_solver.cc:13-26
    std::vector<Vec> W(m, Vec(n));
    Vec col(n), y(n);
    for (int k = 0; k < m; ++k) {
      std::fill(col.begin(), col.end(), 0.0);
      for (long p = op->BT.rowptr[k]; p < op->BT.rowptr[k + 1]; ++p) col[op->BT.col[p]] = op->BT.val[p];
      dense_solveL(L, n, col.data(), y.data());
      dense_solveLT(L, n, y.data(), W[k].data());
    }
    std::vector<double> S(m * m, 0.0);
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b) {
        double s = (a == b) ? op->Sigma[a] : 0.0;
        for (long p = op->BT.rowptr[a]; p < op->BT.rowptr[a + 1]; ++p) s += op->BT.val[p] * W[b][op->BT.col[p]];
        S[a * m + b] = s;
      }
    std::vector<double> Sinv = dense_inverse(S, m);
    B_bar.assign(m, Vec(n, 0.0));
    for (int b = 0; b < m; ++b)
      for (int a = 0; a < m; ++a)
        for (long i = 0; i < n; ++i) B_bar[b][i] += W[a][i] * Sinv[a * m + b];
  }
}

void CholeskySolver::apply(const double *b, double *x) const {  // cholesky_solver.cc:30-41
  Vec y(n), z(n);
  dense_solveL(L, n, b, z.data());
  dense_solveLT(L, n, z.data(), y.data());
  const int m = op->m_lowrank;
  if (m > 0) {
    Vec BTy(m);
    op->BT.matvec(y.data(), BTy.data());
    for (long i = 0; i < n; ++i) x[i] = y[i];
    for (int c = 0; c < m; ++c)
      for (long i = 0; i < n; ++i) x[i] -= B_bar[c][i] * BTy[c];
  } else {
    for (long i = 0; i < n; ++i) x[i] = y[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Hierarchy (shared part of multigridmc_sampler.cc:76-98 / multigrid_preconditioner.cc:47-69)
// ---------------------------------------------------------------------------------------------
Hierarchy::Hierarchy(const std::shared_ptr<LinearOperator> &fine, int nlevel, int ordering) {
  std::shared_ptr<LinearOperator> lin_op = fine;
  for (int level = 0; level < nlevel; ++level) {
    ops.push_back(lin_op);
    orders.push_back(make_order(lin_op->lattice, lin_op->A, ordering));
    if (level < nlevel - 1) {
      intergrids.emplace_back(lin_op->lattice);
      lin_op = std::make_shared<LinearOperator>(lin_op->coarsen(intergrids.back()));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MultigridMCSampler
// ---------------------------------------------------------------------------------------------
MultigridMCSampler::MultigridMCSampler(const std::shared_ptr<Hierarchy> &H_, std::mt19937_64 *engine_, const MultigridParameters &p,
                                       bool use_philox, uint64_t philox_seed)
    : H(H_), params(p), engine(engine_) {
  if ((int)H->ops.size() < p.nlevel) throw std::runtime_error("hierarchy has fewer levels than params.nlevel");
  if (use_philox) {
    px = std::make_shared<PhiloxCtx>();
    px->seed = philox_seed;
    px->sweep_counter.assign(p.nlevel, 0);
  }
  NoiseSource ns;
  ns.engine = engine;
  ns.px = px.get();
  for (int level = 0; level < p.nlevel; ++level) {
    const LinearOperator *op = H->ops[level].get();
    const long n = op->ndof();
    x_ell.emplace_back(n, 0.0);
    f_ell.emplace_back(n, 0.0);
    r_ell.emplace_back(n, 0.0);
    LevelSamplers ls;
    if (level < p.nlevel - 1) {  // the reference also builds (unused) samplers on the coarsest level (:86-89)
      if (p.smoother == 0) {
        ls.sor_pre = std::make_shared<SORSampler>(op, ns, p.omega, p.npresmooth, forward, H->orders[level], level);
        ls.sor_post = std::make_shared<SORSampler>(op, ns, p.omega, p.npostsmooth, backward, H->orders[level], level);
      } else {
        ls.ssor_pre = std::make_shared<SSORSampler>(op, ns, p.omega, p.npresmooth, H->orders[level], level);
        ls.ssor_post = std::make_shared<SSORSampler>(op, ns, p.omega, p.npostsmooth, H->orders[level], level);
      }
    }
    samplers.push_back(ls);
  }
  const int lc = p.nlevel - 1;
  if (p.coarse_solver == 1)
    coarse_cholesky = std::make_shared<CholeskySampler>(H->ops[lc].get(), ns, lc);
  else
    coarse_ssor = std::make_shared<SSORSampler>(H->ops[lc].get(), ns, p.omega, p.ncoarsesmooth, H->orders[lc], lc);
}

void MultigridMCSampler::sample(int level) const {
  if (level == params.nlevel - 1) {
    if (coarse_cholesky)
      coarse_cholesky->apply(f_ell[level].data(), x_ell[level].data());
    else
      coarse_ssor->apply(f_ell[level].data(), x_ell[level].data());
    return;
  }
  const int cycle_ = (level > 0) ? params.cycle : 1;
  const LevelSamplers &ls = samplers[level];
  const long n = (long)x_ell[level].size();
  for (int j = 0; j < cycle_; ++j) {
    if (ls.sor_pre) ls.sor_pre->apply(f_ell[level].data(), x_ell[level].data());
    else ls.ssor_pre->apply(f_ell[level].data(), x_ell[level].data());
    H->ops[level]->apply(x_ell[level].data(), r_ell[level].data());
    for (long i = 0; i < n; ++i) r_ell[level][i] = f_ell[level][i] - r_ell[level][i];
    H->intergrids[level].restrict(r_ell[level].data(), f_ell[level + 1].data());
    std::fill(x_ell[level + 1].begin(), x_ell[level + 1].end(), 0.0);
    sample(level + 1);
    H->intergrids[level].prolongate_add(params.coarse_scaling, x_ell[level + 1].data(), x_ell[level].data());
    if (ls.sor_post) ls.sor_post->apply(f_ell[level].data(), x_ell[level].data());
    else ls.ssor_post->apply(f_ell[level].data(), x_ell[level].data());
  }
}

void MultigridMCSampler::apply(const double *f, double *x) const {
  const long n = (long)x_ell[0].size();
  std::copy(f, f + n, f_ell[0].begin());
  std::copy(x, x + n, x_ell[0].begin());
  if (px) std::fill(px->sweep_counter.begin(), px->sweep_counter.end(), 0u);
  sample(0);
  std::copy(x_ell[0].begin(), x_ell[0].end(), x);
  if (px) px->sample++;
}

// ---------------------------------------------------------------------------------------------
// MultigridPreconditioner
// ---------------------------------------------------------------------------------------------
MultigridPreconditioner::MultigridPreconditioner(const std::shared_ptr<Hierarchy> &H_, const MultigridParameters &p) : H(H_), params(p) {
  if ((int)H->ops.size() < p.nlevel) throw std::runtime_error("hierarchy has fewer levels than params.nlevel");
  for (int level = 0; level < p.nlevel; ++level) {
    const LinearOperator *op = H->ops[level].get();
    const long n = op->ndof();
    x_ell.emplace_back(n, 0.0);
    b_ell.emplace_back(n, 0.0);
    r_ell.emplace_back(n, 0.0);
    LevelSmoothers ls;
    if (level < p.nlevel - 1) {
      if (p.smoother == 0) {
        ls.sor_pre = std::make_shared<SORSmoother>(op, p.omega, p.npresmooth, forward, H->orders[level]);
        ls.sor_post = std::make_shared<SORSmoother>(op, p.omega, p.npostsmooth, backward, H->orders[level]);
      } else {
        ls.ssor_pre = std::make_shared<SSORSmoother>(op, p.omega, p.npresmooth, H->orders[level]);
        ls.ssor_post = std::make_shared<SSORSmoother>(op, p.omega, p.npostsmooth, H->orders[level]);
      }
    }
    smoothers.push_back(ls);
  }
  coarse_solver = std::make_shared<CholeskySolver>(H->ops[p.nlevel - 1].get());  // always Cholesky (:41-45)
}

void MultigridPreconditioner::solve(int level) {
  std::fill(x_ell[level].begin(), x_ell[level].end(), 0.0);
  if (level == params.nlevel - 1) {
    coarse_solver->apply(b_ell[level].data(), x_ell[level].data());
    return;
  }
  const int cycle_ = (level > 0) ? params.cycle : 1;
  const LevelSmoothers &ls = smoothers[level];
  const long n = (long)x_ell[level].size();
  for (int j = 0; j < cycle_; ++j) {
    if (ls.sor_pre) ls.sor_pre->apply(b_ell[level].data(), x_ell[level].data());
    else ls.ssor_pre->apply(b_ell[level].data(), x_ell[level].data());
    H->ops[level]->apply(x_ell[level].data(), r_ell[level].data());
    for (long i = 0; i < n; ++i) r_ell[level][i] = b_ell[level][i] - r_ell[level][i];
    H->intergrids[level].restrict(r_ell[level].data(), b_ell[level + 1].data());
    solve(level + 1);
    H->intergrids[level].prolongate_add(params.coarse_scaling, x_ell[level + 1].data(), x_ell[level].data());
    if (ls.sor_post) ls.sor_post->apply(b_ell[level].data(), x_ell[level].data());
    else ls.ssor_post->apply(b_ell[level].data(), x_ell[level].data());
  }
}

void MultigridPreconditioner::apply(const double *b, double *x) {
  std::copy(b, b + b_ell[0].size(), b_ell[0].begin());
  solve(0);
  std::copy(x_ell[0].begin(), x_ell[0].end(), x);
}

// ---------------------------------------------------------------------------------------------
// LoopSolver (loop_solver.cc:9-53)
// ---------------------------------------------------------------------------------------------
LoopSolverResult loop_solve(const LinearOperator &op, MultigridPreconditioner &prec, double rtol, double atol, int maxiter, int verbose,
                            const double *b, double *x) {
  const long n = op.ndof();
  LoopSolverResult res;
  double r0 = 0.0;
  for (long i = 0; i < n; ++i) r0 += b[i] * b[i];
  r0 = std::sqrt(r0);
  res.r0_nrm = r0;
  if (verbose >= 2) printf("Initial residual ||r_0|| =  %12.4f\n", r0);
  std::fill(x, x + n, 0.0);
  Vec r(n), Pr(n);
  double r_nrm = 0.0, rold = r0;
  if (verbose >= 2) printf("%5s   %8s   %12s   %6s\n", "iter", "||r||", "||r||/||r_0||", "rho");
  res.niter = maxiter;
  for (int k = 0; k < maxiter; ++k) {
    op.apply(x, r.data());
    double s = 0.0;
    for (long i = 0; i < n; ++i) {
      r[i] -= b[i];
      s += r[i] * r[i];
    }
    r_nrm = std::sqrt(s);
    res.history.push_back(r_nrm);
    if (verbose >= 2) printf("%5d   %8.3e   %12.3e   %6.3f\n", k, r_nrm, r_nrm / r0, r_nrm / rold);
    if ((r_nrm / r0 < rtol) && (r_nrm < atol)) {
      res.niter = k;
      res.converged = true;
      break;
    }
    rold = r_nrm;
    prec.apply(r.data(), Pr.data());
    for (long i = 0; i < n; ++i) x[i] -= Pr[i];
  }
  if (verbose >= 1) {
    if (res.converged)
      printf("Solver converged after %5d iterations\n||r|| = %8.3e, ||r||/||r_0|| = %8.3e\n", res.niter, r_nrm, r_nrm / r0);
    else
      printf("Solver failed to converge after %5d iterations\n", maxiter);
  }
  return res;
}

// ---------------------------------------------------------------------------------------------
// statistics.cc:4-79 specialised to a scalar series: running averages S_k of q_t q_{t-k} over
// t >= k, C(k) = S_k - avg^2, tau_int = 1 + 2 sum_{k=1}^{K-1} (1 - k/K) C(k)/C(0)
// ---------------------------------------------------------------------------------------------
double tau_int_scalar(const double *q, long n, int k_max) {
  if (n < 2) return 1.0;
  const int K = (int)std::min<long>(k_max, n);
  double avg = 0.0;
  for (long t = 0; t < n; ++t) avg += (q[t] - avg) / (1.0 * (t + 1));
  std::vector<double> S(K, 0.0);
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    long cnt = 0;
    for (long t = k; t < n; ++t) {
      cnt++;
      s += (q[t] * q[t - k] - s) / (1.0 * cnt);
    }
    S[k] = s;
  }
  const double variance = S[0] - avg * avg;
  double tau = 1.0;
  for (int k = 1; k < K; ++k) tau += 2 * (1. - k / (1.0 * K)) * (S[k] - avg * avg) / variance;
  return tau;
}

}  // namespace orc
