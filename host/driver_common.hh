// Shared set-up of the two drivers: parameter files -> lattice -> prior / posterior operator, the same
// sequence and error messages as the reference's main() functions (driver_mgmc.cc:319-446,
// driver_mg.cc:28-151), built from the device-backed classes of mgmc_host.hh.
#ifndef MGMC_DRIVER_COMMON_HH
#define MGMC_DRIVER_COMMON_HH
#include "mgmc_host.hh"

struct ProblemSetup {
  GeneralParameters general;
  LatticeParameters lattice_p;
  CholeskyParameters cholesky;
  SmootherParameters smoother;
  IterativeSolverParameters iterative_solver;
  MultigridParameters multigrid;
  SamplingParameters sampling;
  PriorParameters prior;
  ConstantCorrelationLengthModelParameters constant_clm;
  PeriodicCorrelationLengthModelParameters periodic_clm;
  MeasurementParameters measurements;
  std::shared_ptr<Lattice> lattice;
  std::shared_ptr<LinearOperator> prior_operator, linear_operator;
  std::shared_ptr<MeasuredOperator> posterior_operator;
};

inline void die(const std::string &msg) {
  std::cout << msg << std::endl;
  exit(-1);
}

inline ProblemSetup setup_problem(const std::string &filename, bool need_sampling) {
  ProblemSetup s;
  std::cout << "Reading parameters from file '" << filename << "'" << std::endl;
  s.general.read_from_file(filename);
  s.lattice_p.read_from_file(filename);
  if (need_sampling) s.cholesky.read_from_file(filename);
  s.smoother.read_from_file(filename);
  s.multigrid.read_from_file(filename);
  if (need_sampling) s.sampling.read_from_file(filename);
  else s.iterative_solver.read_from_file(filename);
  s.prior.read_from_file(filename);
  s.constant_clm.read_from_file(filename);
  s.periodic_clm.read_from_file(filename);
  s.measurements.read_from_file(filename);
  if (s.measurements.dim != s.general.dim) die("ERROR: dimension of measurement locations differs from problem dimension");
  std::cout << "B200 device path: matrix-free stencils, multicolour sweeps, Philox noise (libmgmc_b200)." << std::endl << std::endl;

  if (s.general.dim == 2) s.lattice = std::make_shared<Lattice2d>(s.lattice_p.nx, s.lattice_p.ny);
  else if (s.general.dim == 3) s.lattice = std::make_shared<Lattice3d>(s.lattice_p.nx, s.lattice_p.ny, s.lattice_p.nz);
  else die("ERROR: Invalid dimension : " + std::to_string(s.general.dim));

  std::shared_ptr<CorrelationLengthModel> clm;
  if (s.prior.correlationlength_model == "constant") clm = std::make_shared<ConstantCorrelationLengthModel>(s.constant_clm);
  else if (s.prior.correlationlength_model == "periodic") clm = std::make_shared<PeriodicCorrelationLengthModel>(s.periodic_clm);
  else die("Error: invalid correlationlengthmodel '" + s.prior.correlationlength_model + "'");

  if (s.prior.pde_model == "shiftedlaplace_fd") s.prior_operator = std::make_shared<ShiftedLaplaceFDOperator>(s.lattice, clm, 1);
  else if (s.prior.pde_model == "squared_shiftedlaplace_fd") s.prior_operator = std::make_shared<SquaredShiftedLaplaceFDOperator>(s.lattice, clm, 1);
  else if (s.prior.pde_model == "shiftedlaplace_fem") s.prior_operator = std::make_shared<ShiftedLaplaceFEMOperator>(s.lattice, clm, 1);
  else die("Error: invalid prior '" + s.prior.pde_model + "'");

  s.posterior_operator = std::make_shared<MeasuredOperator>(s.prior_operator, s.measurements);
  if (s.general.operator_name == "prior") s.linear_operator = s.prior_operator;
  else if (s.general.operator_name == "posterior") s.linear_operator = s.posterior_operator;
  else die("ERROR: invalid operator : " + s.general.operator_name);
  return s;
}
#endif
