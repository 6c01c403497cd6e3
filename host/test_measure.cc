// Prints the measurement functional MeasuredOperator::measurement_vector (measured_operator.cc:69-170) of the host layer for a few
// fixed points on 2d and 3d lattices (closest vertex and ball averages), entry by entry -- tests/test_host_drivers.py compares the
// output with the oracle's restatement.  No CUDA calls.
#include "mgmc_host.hh"

static void dump(const char *tag, const std::shared_ptr<Lattice> &lattice, const std::vector<double> &x0, double radius) {
  ConstantCorrelationLengthModelParameters cp;
  cp.Lambda = 0.2;
  auto clm = std::make_shared<ConstantCorrelationLengthModel>(cp);
  auto prior = std::make_shared<ShiftedLaplaceFDOperator>(lattice, clm, 0);
  MeasurementParameters mp;
  mp.dim = lattice->dim();
  mp.n = 0;
  mp.radius = radius;
  MeasuredOperator op(prior, mp);
  Eigen::VectorXd p((long)x0.size());
  for (size_t d = 0; d < x0.size(); ++d) p[(long)d] = x0[d];
  const Eigen::SparseVector<double> r = op.measurement_vector(p, radius);
  for (auto &e : r.entries()) printf("mv %s %ld %.17g\n", tag, (long)e.first, e.second);
}

int main() {
  auto l2 = std::make_shared<Lattice2d>(16, 12);
  auto l3 = std::make_shared<Lattice3d>(8, 12, 10);
  dump("2d_point", l2, {0.37, 0.62}, 0.0);
  dump("2d_ball", l2, {0.37, 0.62}, 0.15);
  dump("2d_edge", l2, {0.97, 0.02}, 0.1);
  dump("3d_point", l3, {0.37, 0.62, 0.48}, 0.0);
  dump("3d_corner_point", l3, {0.999, 0.001, 1.0}, 0.0);
  dump("3d_ball", l3, {0.37, 0.62, 0.48}, 0.2);
  dump("3d_edge", l3, {0.95, 0.05, 0.5}, 0.15);
  return 0;
}
