// Prints what the parameter classes read from a configuration file (used by tests/test_host_drivers.py
// to check the libconfig-subset reader against the reference's file format).  No CUDA calls.
#include "mgmc_host.hh"

int main(int argc, char *argv[]) {
  if (argc != 2) return 2;
  const std::string fn(argv[1]);
  GeneralParameters g;
  LatticeParameters l;
  CholeskyParameters c;
  SmootherParameters s;
  IterativeSolverParameters it;
  MultigridParameters mg;
  SamplingParameters sp;
  PriorParameters pr;
  ConstantCorrelationLengthModelParameters cc;
  PeriodicCorrelationLengthModelParameters pc;
  MeasurementParameters m;
  g.read_from_file(fn);
  l.read_from_file(fn);
  c.read_from_file(fn);
  s.read_from_file(fn);
  it.read_from_file(fn);
  mg.read_from_file(fn);
  sp.read_from_file(fn);
  pr.read_from_file(fn);
  cc.read_from_file(fn);
  pc.read_from_file(fn);
  m.read_from_file(fn);
  printf("dim=%d operator=%s do_ssor=%d do_multigridmc=%d\n", g.dim, g.operator_name.c_str(), (int)g.do_ssor, (int)g.do_multigridmc);
  printf("lattice=%u,%u,%u\n", l.nx, l.ny, l.nz);
  printf("smoother omega=%.17g nsmooth=%u\n", s.omega, s.nsmooth);
  printf("solver rtol=%.17g atol=%.17g maxiter=%u verbose=%d\n", it.rtol, it.atol, it.maxiter, it.verbose);
  printf("multigrid nlevel=%u smoother=%s coarse=%s pre=%u post=%u ncoarse=%u cycle=%u scaling=%.17g omega=%.17g\n", mg.nlevel, mg.smoother.c_str(),
         mg.coarse_solver.c_str(), mg.npresmooth, mg.npostsmooth, mg.ncoarsesmooth, mg.cycle, mg.coarse_scaling, mg.omega);
  printf("sampling nsamples=%u nwarmup=%u nsteps=%u nconv=%u\n", sp.nsamples, sp.nwarmup, sp.nstepsconvergence, sp.nsamplesconvergence);
  printf("prior pde=%s clm=%s Lambda=%.17g Lmin=%.17g Lmax=%.17g\n", pr.pde_model.c_str(), pr.correlationlength_model.c_str(), cc.Lambda, pc.Lambda_min, pc.Lambda_max);
  printf("measurements n=%u dim=%d radius=%.17g vscale=%.17g sample=%.17g,%.17g global=%d\n", m.n, m.dim, m.radius, m.variance_scaling, m.sample_location[0],
         m.sample_location[1], (int)m.measure_global);
  for (unsigned int k = 0; k < m.n; ++k) printf("meas %u %.17g %.17g %.17g %.17g\n", k, m.measurement_locations[k][0], m.measurement_locations[k][1], m.mean[k], m.variance[k]);
  // lattice known answers (test_lattice.hh)
  Lattice2d lat(4, 5);
  printf("lattice2d Nvertex=%u Ncell=%u fine_vertex_idx(7)=%u\n", lat.Nvertex, lat.Ncell, lat.fine_vertex_idx(7));
  return 0;
}
