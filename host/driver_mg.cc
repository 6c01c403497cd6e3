// driver_mg CONFIGURATIONFILE -- deterministic multigrid solve on the B200 (reference: src/driver_mg.cc).
// Same parameter file, same right-hand side (std::mt19937_64(1482817) normals in lexicographic
// order, driver_mg.cc:165-172), same LoopSolver output and solution.vtk.
#include <chrono>

#include "driver_common.hh"

int main(int argc, char *argv[]) {
  if (argc != 2) {
    std::cout << "Usage: " << argv[0] << " CONFIGURATIONFILE" << std::endl;
    exit(-1);
  }
  std::cout << std::endl << "+------------------+" << std::endl << "! Multigrid solver !" << std::endl << "+------------------+" << std::endl << std::endl;
  ProblemSetup s = setup_problem(argv[1], false);
  std::shared_ptr<Preconditioner> prec = std::make_shared<MultigridPreconditioner>(s.linear_operator, s.multigrid);
  std::cout << std::endl;
  LoopSolver solver(s.linear_operator, prec, s.iterative_solver);
  const unsigned int ndof = s.linear_operator->get_ndof();
  Eigen::VectorXd x(ndof), b(ndof);
  std::mt19937_64 rng(1482817);
  std::normal_distribution<double> normal_dist(0.0, 1.0);
  for (unsigned int ell = 0; ell < s.lattice->Nvertex; ++ell) b[ell] = normal_dist(rng);
  const auto t0 = std::chrono::high_resolution_clock::now();
  solver.apply(b, x);
  const auto t1 = std::chrono::high_resolution_clock::now();
  printf("solve time = %12.4f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
  VTKWriter2d vtk("solution.vtk", s.lattice, 1);
  vtk.add_state(x, "numerical");
  vtk.write();
  return 0;
}
