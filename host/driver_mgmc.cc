// driver_mgmc CONFIGURATIONFILE -- MGMC / SSOR sampling experiments on the B200 (reference:
// src/driver_mgmc.cc).  Same parameter files, same experiments (time series + timing, convergence of
// the first moments from x = 0, posterior mean / variance fields), same stdout lines and output files.
// The hot loops run device-resident (Sampler::sample_series): the QoI z = sample_vector . x is
// evaluated on the GPU after every cycle instead of copying x to the host.
#include <chrono>

#include "driver_common.hh"

namespace {

Eigen::VectorXd measurement_targets(const MeasurementParameters &mp) {
  Eigen::VectorXd y(mp.n + mp.measure_global);
  for (unsigned int k = 0; k < mp.n; ++k) y[k] = mp.mean[k];
  if (mp.measure_global) y[mp.n] = mp.mean_global;
  return y;
}

// driver_mgmc.cc:40-107
void measure_sampling_time(std::shared_ptr<Sampler> sampler, const SamplingParameters &sp, const MeasurementParameters &mp, const std::string label,
                           const std::string filename) {
  const std::shared_ptr<LinearOperator> op = sampler->get_linear_operator();
  const unsigned int ndof = op->get_ndof();
  Eigen::VectorXd xbar(ndof);
  xbar.setZero();
  const Eigen::VectorXd y = measurement_targets(mp);
  const Eigen::VectorXd mean_x_exact = op->mean(xbar, y);
  const auto measured = std::make_shared<MeasuredOperator>(op, mp);
  const Eigen::SparseVector<double> sample_vector = measured->measurement_vector(mp.sample_location, mp.radius);
  Eigen::VectorXd x(ndof), f(ndof);
  x.setZero();
  op->apply(mean_x_exact, f);
  sampler->fix_rhs(f);
  std::vector<double> warm(sp.nwarmup), data(sp.nsamples);
  sampler->sample_series(f, x, sample_vector, warm);
  const auto t_start = std::chrono::high_resolution_clock::now();
  sampler->sample_series(f, x, sample_vector, data);
  const auto t_finish = std::chrono::high_resolution_clock::now();
  const double t_elapsed = std::chrono::duration<double, std::milli>(t_finish - t_start).count() / (1.0 * sp.nsamples);
  printf("  %12s time per sample = %12.4f ms\n", label.c_str(), t_elapsed);
  std::ofstream out(filename);
  for (double z : data) out << z << std::endl;
  out.close();
  Statistics stat(label, 20);
  for (double z : data) stat.record_sample(z);
  double x_avg = 0.0, xsq_avg = 0.0;
  for (unsigned int k = 0; k < sp.nsamples; ++k) {
    x_avg += (data[k] - x_avg) / (k + 1.0);
    xsq_avg += (data[k] * data[k] - xsq_avg) / (k + 1.0);
  }
  const double variance = xsq_avg - x_avg * x_avg;
  const double x_error = sqrt(variance / sp.nsamples);
  double mean_exact, variance_exact;
  measured->observed_mean_and_variance(xbar, y, sample_vector, mean_exact, variance_exact);
  printf("  %12s mean     = %12.4e +/- %12.4e [ignoring IACT]\n", label.c_str(), x_avg, x_error);
  printf("  %12s mean     = %12.4e\n", "exact", mean_exact);
  printf("  %12s variance = %12.4e\n", label.c_str(), variance);
  printf("  %12s variance = %12.4e\n", "exact", variance_exact);
  if (sp.nsamples > 40) printf("  %12s tau_int  = %12.4f\n", label.c_str(), stat.tau_int());
  printf("\n");
}

// driver_mgmc.cc:118-171
void posterior_statistics(std::shared_ptr<MultigridMCSampler> sampler, const SamplingParameters &sp, const MeasurementParameters &mp) {
  const std::shared_ptr<LinearOperator> op = sampler->get_linear_operator();
  const unsigned int ndof = op->get_ndof();
  Eigen::VectorXd xbar(ndof);
  xbar.setZero();
  const Eigen::VectorXd y = measurement_targets(mp);
  const Eigen::VectorXd mean_x_exact = op->mean(xbar, y);
  Eigen::VectorXd x(ndof), f(ndof), mean(ndof), second(ndof), variance(ndof);
  x.setZero();
  op->apply(mean_x_exact, f);
  const Eigen::SparseVector<double> none(ndof);
  std::vector<double> warm(sp.nwarmup);
  sampler->sample_series(f, x, none, warm);
  sampler->sample_moments(f, x, sp.nsamples, mean, second);
  for (unsigned int ell = 0; ell < ndof; ++ell) variance[ell] = second[ell] - mean[ell] * mean[ell];
  VTKWriter2d vtk("posterior.vtk", op->get_lattice(), 1);
  vtk.add_state(mean_x_exact, "x_exact_mean");
  vtk.add_state(mean, "x_mean");
  vtk.add_state(variance, "x_variance");
  vtk.write();
}

// driver_mgmc.cc:188-314
void measure_convergence(std::shared_ptr<Sampler> sampler, const SamplingParameters &sp, const MeasurementParameters &mp, const std::string filename) {
  const std::shared_ptr<LinearOperator> op = sampler->get_linear_operator();
  const unsigned int ndof = op->get_ndof();
  Eigen::VectorXd xbar(ndof);
  xbar.setZero();
  const Eigen::VectorXd y = measurement_targets(mp);
  const Eigen::VectorXd mean_x_exact = op->mean(xbar, y);
  const auto measured = std::make_shared<MeasuredOperator>(op, mp);
  const Eigen::SparseVector<double> sample_vector = measured->measurement_vector(mp.sample_location, mp.radius);
  Eigen::VectorXd x(ndof), f(ndof);
  op->apply(mean_x_exact, f);
  sampler->fix_rhs(f);
  const unsigned int nsteps = sp.nstepsconvergence, nsamples = sp.nsamplesconvergence;
  std::vector<double> m1(nsteps + 1, 0.0), m2(nsteps + 1, 0.0), m3(nsteps + 1, 0.0), m4(nsteps + 1, 0.0), z(nsteps);
  for (unsigned int k = 0; k < nsamples; ++k) {
    x.setZero();
    sampler->sample_series(f, x, sample_vector, z);  // z[j-1] = observation after j steps from x = 0
    for (unsigned int j = 1; j <= nsteps; ++j) {
      const double v = z[j - 1];
      m1[j] += (v - m1[j]) / (k + 1.0);
      m2[j] += (v * v - m2[j]) / (k + 1.0);
      m3[j] += (v * v * v - m3[j]) / (k + 1.0);
      m4[j] += (v * v * v * v - m4[j]) / (k + 1.0);
    }
  }
  double mean_exact, variance_exact;
  measured->observed_mean_and_variance(xbar, y, sample_vector, mean_exact, variance_exact);
  std::vector<double> diff[2], err[2];
  for (unsigned int j = 0; j <= nsteps; ++j) {
    diff[0].push_back(fabs(m1[j] - mean_exact));
    diff[1].push_back(fabs(m2[j] - m1[j] * m1[j] - variance_exact));
    const double sigma_sq = nsamples / (nsamples - 1.) * (m2[j] - m1[j] * m1[j]);
    const double mu4 = m4[j] - 4 * m1[j] * m3[j] + 6 * pow(m1[j], 2) * m2[j] - 3 * pow(m1[j], 4);
    err[0].push_back(sqrt(sigma_sq / nsamples));
    err[1].push_back(sqrt((mu4 - (nsamples - 3.) / (nsamples - 1.) * sigma_sq * sigma_sq) / nsamples));
  }
  std::ofstream out(filename);
  char buffer[256];
  for (int q = 0; q < 2; ++q) {
    out << (q == 0 ? "**** q_k = |E[z^k] - E[z]| **** " : "**** q_k = |Var[z^k] - Var[z]| **** ") << std::endl;
    const char *label = (q == 0) ? "mean" : "variance";
    snprintf(buffer, sizeof(buffer), "  %12s   %3s : %12s %35s %35s\n", "", "k", "q_k", "q_k/q_0", "q_k/q_{k-1}");
    out << buffer;
    for (unsigned int j = 0; j <= nsteps; ++j) {
      const double d = diff[q][j], e = err[q][j], d0 = diff[q][0];
      snprintf(buffer, sizeof(buffer), "  %12s   %3d : %12.8f +/- %12.8f       %12.8f +/- %12.8f      ", label, j, d, e, d / d0, e / d0);
      out << buffer;
      if (j > 0) {
        const double dp = diff[q][j - 1], ep = err[q][j - 1];
        snprintf(buffer, sizeof(buffer), " %12.8f +/- %12.8f \n", d / dp, d / dp * sqrt(pow(e / d, 2) + pow(ep / dp, 2)));
      } else {
        snprintf(buffer, sizeof(buffer), " %12s\n", "---");
      }
      out << buffer;
    }
    out << std::endl;
  }
}

}  // namespace

int main(int argc, char *argv[]) {
  if (argc != 2) {
    std::cout << "Usage: " << argv[0] << " CONFIGURATIONFILE" << std::endl;
    exit(-1);
  }
  const auto t_start = std::chrono::high_resolution_clock::now();
  ProblemSetup s = setup_problem(argv[1], true);
  std::mt19937_64 rng(5418513);  // driver_mgmc.cc:448 (here: the source of the Philox keys)
  if (s.general.do_cholesky)
    std::cout << "NOTE: do_cholesky is ignored: a Cholesky factorisation of the fine-level matrix is not part of the device path" << std::endl;
  if (s.general.do_ssor) {
    std::cout << "**** SSOR ****" << std::endl;
    std::shared_ptr<Sampler> ssor = std::make_shared<SSORSampler>(s.linear_operator, rng, s.smoother.omega, s.smoother.nsmooth);
    measure_sampling_time(ssor, s.sampling, s.measurements, "SSOR", "timeseries_ssor.txt");
    measure_convergence(ssor, s.sampling, s.measurements, "convergence_ssor.txt");
  }
  if (s.general.do_multigridmc) {
    std::cout << "**** Multigrid MC ****" << std::endl;
    auto mgmc = std::make_shared<MultigridMCSampler>(s.linear_operator, rng, s.multigrid, s.cholesky);
    measure_sampling_time(mgmc, s.sampling, s.measurements, "MultigridMC", "timeseries_multigridmc.txt");
    measure_convergence(mgmc, s.sampling, s.measurements, "convergence_multigridmc.txt");  // the cfg flag is ignored (driver_mgmc.cc:510)
    if (s.general.save_posterior_statistics) posterior_statistics(mgmc, s.sampling, s.measurements);
  }
  const auto t_finish = std::chrono::high_resolution_clock::now();
  const long total = std::chrono::duration_cast<std::chrono::seconds>(t_finish - t_start).count();
  printf("total runtime = %ld s [ %ld h %ld m %ld s ]\n", total, total / 3600, (total / 60) % 60, total % 60);
  return 0;
}
