// Runs the reference's own quadrature test file (src/auxilliary/test_quadrature.hh, compiled in place from a temporary copy) against
// the drop-in host layer's GaussLegendreQuadrature.
#include "test_quadrature.hh"
int main() { return ::testing::run_all_tests(); }
