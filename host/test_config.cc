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
  // test_lattice.hh:30-101 (1d lattice, n = 6)
  Lattice1d lat1(6);
  printf("lattice1d Nvertex=%u Ncell=%u cell5=%d cell(3)=%u cellshifts=%u,%u,%u,%u vertex4=%d vertex(3)=%u shifts=%u,%u,%u,%u fine=%u,%u,%u info='%s'\n", lat1.Nvertex, lat1.Ncell,
         lat1.cellidx_linear2euclidean(5)[0], lat1.cellidx_euclidean2linear(Eigen::VectorXi({3})), lat1.shift_cellidx(3, Eigen::VectorXi({1})),
         lat1.shift_cellidx(3, Eigen::VectorXi({-1})), lat1.shift_cellidx(4, Eigen::VectorXi({1})), lat1.shift_cellidx(4, Eigen::VectorXi({-1})),
         lat1.vertexidx_linear2euclidean(4)[0], lat1.vertexidx_euclidean2linear(Eigen::VectorXi({3})), lat1.shift_vertexidx(3, Eigen::VectorXi({1})),
         lat1.shift_vertexidx(3, Eigen::VectorXi({-1})), lat1.shift_vertexidx(4, Eigen::VectorXi({1})), lat1.shift_vertexidx(4, Eigen::VectorXi({-1})),
         lat1.fine_vertex_idx(3), lat1.fine_vertex_idx(0), lat1.fine_vertex_idx(2), lat1.get_info().c_str());
  // test_lattice.hh:171-242 (3d lattice 4 x 5 x 6)
  Lattice3d lat3(4, 5, 6);
  const Eigen::VectorXi c53 = lat3.cellidx_linear2euclidean(53), v23 = lat3.vertexidx_linear2euclidean(23);
  printf("lattice3d Nvertex=%u Ncell=%u cell53=%d,%d,%d cell(1,3,2)=%u vertex23=%d,%d,%d vertex(3,4,2)=%u\n", lat3.Nvertex, lat3.Ncell, c53[0], c53[1], c53[2],
         lat3.cellidx_euclidean2linear(Eigen::VectorXi({1, 3, 2})), v23[0], v23[1], v23[2], lat3.vertexidx_euclidean2linear(Eigen::VectorXi({3, 4, 2})));
  printf("lattice3d cellshifts(59)=%u,%u,%u,%u,%u,%u\n", Lattice3d(4, 5, 6).shift_cellidx(59, Eigen::VectorXi({0, 1, 0})), Lattice3d(4, 5, 6).shift_cellidx(59, Eigen::VectorXi({0, -1, 0})),
         Lattice3d(4, 5, 6).shift_cellidx(59, Eigen::VectorXi({1, 0, 0})), Lattice3d(4, 5, 6).shift_cellidx(59, Eigen::VectorXi({-1, 0, 0})),
         Lattice3d(4, 5, 6).shift_cellidx(59, Eigen::VectorXi({0, 0, 1})), Lattice3d(4, 5, 6).shift_cellidx(59, Eigen::VectorXi({0, 0, -1})));
  printf("lattice3d shifts(23)=%u,%u,%u,%u,%u,%u fine_vertex_idx(23)=%u\n", lat3.shift_vertexidx(23, Eigen::VectorXi({0, 1, 0})),
         lat3.shift_vertexidx(23, Eigen::VectorXi({0, -1, 0})), lat3.shift_vertexidx(23, Eigen::VectorXi({1, 0, 0})), lat3.shift_vertexidx(23, Eigen::VectorXi({-1, 0, 0})),
         lat3.shift_vertexidx(23, Eigen::VectorXi({0, 0, 1})), lat3.shift_vertexidx(23, Eigen::VectorXi({0, 0, -1})), lat3.fine_vertex_idx(23));
  {
    const Eigen::VectorXd xc = lat3.vertex_coordinates(23);
    const std::shared_ptr<Lattice> coarse = Lattice3d(8, 4, 6).get_coarse_lattice();
    const Eigen::VectorXi cs = coarse->shape();
    printf("lattice3d coords(23)=%.4f,%.4f,%.4f coarse(8,4,6)=%d,%d,%d info='%s'\n", xc[0], xc[1], xc[2], cs[0], cs[1], cs[2], lat3.get_info().c_str());
  }
  return 0;
}
