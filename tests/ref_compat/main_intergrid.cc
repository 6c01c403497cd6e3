// The reference's own src/intergrid/test_intergrid.hh against the drop-in host layer.  Without a device only the cases that need
// none run: TestCoarsenOperator2d / 3d (coarsening the FEM operator with constant coefficients gives the FEM operator of the coarse
// lattice, test_intergrid.hh:172-207) -- they exercise the library's host-side Galerkin stencil algebra through
// LinearOperator::coarsen / get_sparse.  argv[1] = substring filter.
#include "test_intergrid.hh"
int main(int argc, char *argv[]) { return ::testing::run_all_tests(argc > 1 ? argv[1] : ""); }
