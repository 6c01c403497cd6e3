// CholeskySolver / CholeskySampler / their factories over the C ABI (needs a GPU).  Follows the reference's own tests for
// these classes: the Cholesky solver solves A x = b exactly (tests/test_solver.hh:93-113, with and without the
// low-rank measurement term), the Cholesky sampler reproduces the posterior mean (tests/test_sampler.hh:215-283,
// nsamples scaled down).  Prints the numbers tests/test_host_drivers.py asserts on.
#include "mgmc_host.hh"

int main() {
  const unsigned int n = 32;
  auto lattice = std::make_shared<Lattice2d>(n, n);
  ConstantCorrelationLengthModelParameters cp;
  cp.Lambda = 0.3;
  auto clm = std::make_shared<ConstantCorrelationLengthModel>(cp);
  auto prior = std::make_shared<ShiftedLaplaceFDOperator>(lattice, clm);
  MeasurementParameters mp;
  mp.dim = 2;
  mp.n = 3;
  mp.radius = 0.05;
  mp.variance_scaling = 1.0;
  mp.measurement_locations = {Eigen::VectorXd({0.25, 0.25}), Eigen::VectorXd({0.25, 0.75}), Eigen::VectorXd({0.625, 0.5})};
  mp.mean = Eigen::VectorXd({1.0, -0.5, 0.25});
  mp.variance = Eigen::VectorXd({0.1, 0.05, 0.2});
  mp.sample_location = Eigen::VectorXd({0.5, 0.5});
  auto posterior = std::make_shared<MeasuredOperator>(prior, mp);
  std::mt19937_64 rng(31841287);
  std::normal_distribution<double> normal(0.0, 1.0);
  const unsigned int ndof = prior->get_ndof();
  int k = 0;
  for (std::shared_ptr<LinearOperator> op : {std::static_pointer_cast<LinearOperator>(prior), std::static_pointer_cast<LinearOperator>(posterior)}) {
    Eigen::VectorXd x_exact(ndof), b(ndof), x(ndof), Ax(ndof);
    for (unsigned int i = 0; i < ndof; ++i) x_exact[i] = normal(rng);
    op->apply(x_exact, b);
    CholeskySolverFactory factory;
    std::shared_ptr<LinearSolver> solver = factory.get(op);
    solver->apply(b, x);
    op->apply(x, Ax);
    printf("solver %d: |x - x_exact| / |x_exact| = %.3e   |A x - b| / |b| = %.3e\n", k, (x - x_exact).norm() / x_exact.norm(), (Ax - b).norm() / b.norm());
    // sampler: mean of the samples -> A^{-1} f
    Eigen::VectorXd f(ndof), mean(ndof), s(ndof);
    for (unsigned int i = 0; i < ndof; ++i) f[i] = normal(rng);
    solver->apply(f, x);
    DenseCholeskySamplerFactory sfactory(rng);
    std::shared_ptr<Sampler> sampler = sfactory.get(op);
    const int nsamples = 20000;
    mean.setZero();
    double var_mid = 0.0;
    for (int q = 0; q < nsamples; ++q) {
      sampler->apply(f, s);
      mean += s;
      var_mid += (s[ndof / 2] - x[ndof / 2]) * (s[ndof / 2] - x[ndof / 2]);
    }
    mean *= 1.0 / nsamples;
    // exact variance of the entry: (A^{-1})_{ii} = e_i^T A^{-1} e_i
    Eigen::VectorXd e(ndof), col(ndof);
    e.setZero();
    e[ndof / 2] = 1.0;
    solver->apply(e, col);
    printf("sampler %d: |mean - A^-1 f| / |A^-1 f| = %.3e   var / exact = %.4f   (nsamples = %d)\n", k, (mean - x).norm() / x.norm(), var_mid / nsamples / col[ndof / 2], nsamples);
    ++k;
  }
  return 0;
}
