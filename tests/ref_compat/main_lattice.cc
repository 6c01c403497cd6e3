// Runs the reference's own lattice test file (src/lattice/test_lattice.hh, compiled in place from a temporary copy -- never part of
// this repository) against the drop-in host layer.
#include "test_lattice.hh"
int main() { return ::testing::run_all_tests(); }
