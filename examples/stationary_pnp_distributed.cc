// stationary_pnp_distributed -- the reference's parallel run (mpirun -np N dune_pnp <cfg>: MPIHelper, grid->loadBalance(),
// PDELab's non-overlapping backends; /root/reference/src/pnp_solver_main.cc:70-116, bin/dune_pnp.py:19-41) on N B200s, one
// process per GPU, WITHOUT a Python launcher around the library: any process launcher that sets RANK / WORLD_SIZE /
// LOCAL_RANK works (torchrun, mpirun with a wrapper, a shell loop).
//   1. every rank solves the reference flow (PB Newton -> interpolate(BCExtension) -> PNP Newton) on the mesh refined
//      `coarse` times -- small, unpartitioned;
//   2. the Gmsh mesh is partitioned natively (pnp_partition_build), every rank refines its part `levels` times, the coarse
//      solution is carried to the fine mesh (nested iteration);
//   3. monolithic PNP Newton on the distributed fine mesh: BiCGSTAB + distributed multigrid, NCCL halo exchange.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "pnp_b200/drivers.hh"

using namespace Dune::PNPB200;

static int env_int(const char* name, int dflt) { const char* e = std::getenv(name); return e ? std::atoi(e) : dflt; }

int main(int argc, char** argv) {
  if (argc < 6) { std::printf("usage: %s <config.cfg> <mesh.msh> <levels> <coarse level> <rendezvous file>\n", argv[0]); return 1; }
  const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1), local = env_int("LOCAL_RANK", rank);
  const int levels = std::stoi(argv[3]), coarse = std::stoi(argv[4]);
  try {
    typedef ISTLBackend_NOVLP_BCGS_AMG LS;
    std::vector<double> fx, fy, uc;
    { // 1. coarse solve, on every rank
      Grid g0(local);
      g0.readConfigFile(argv[1]); g0.readGmsh(argv[2]);
      g0.globalRefine(coarse); g0.finalize();
      LS pbls(g0, 2, 20000, 0), ls(g0, 2, 20000, 0);
      pnp_solver_set_option(g0.ctx(), pbls.handle(), "amg_geometric", 0);
      pnp_solver_set_option(g0.ctx(), ls.handle(), "amg_geometric", 0);
      Vector u(g0, 3);
      const pnp_newton_result r = stationary_pnp_from_pb(g0, u, pbls, ls, PNP_JAC_ANALYTIC);
      if (rank == 0) std::printf("coarse level %d: %ld vertices, PNP Newton %d iterations, defect %.3e\n", coarse, g0.size(), r.iterations, r.defect);
      fx = g0.coordinates(0); fy = g0.coordinates(1); uc = u.get();
    }
    // 2. decomposition and nested iteration
    Grid grid(local);
    grid.readConfigFile(argv[1]); grid.readGmsh(argv[2]);
    grid.initCommunication(rank, world, argv[5]);
    const int replica = coarse < levels ? coarse : levels - 1;
    const int start = grid.loadBalance(levels, replica, coarse, 3, &fx, &fy, &uc);
    Vector u(grid, 3, start, true);
    // 3. Newton on the fine mesh
    GridOperator<PnpOperator> go(grid);
    LS ls(grid, 2, 20000, 0);
    Newton<GridOperator<PnpOperator>, LS> newton(go, u, ls);
    configure_newton(grid, newton);
    newton.setJacobianMode(PNP_JAC_ANALYTIC);
    newton.apply();
    const pnp_newton_result& r = newton.result();
    std::printf("rank %d of %d: %ld owned + %ld ghost vertices | PNP Newton %d iterations, defect %.6e -> %.6e, %d linear iterations, %.3f s\n",
                rank, world, grid.ownedSize(), grid.size() - grid.ownedSize(), r.iterations, r.first_defect, r.defect, r.linear_iterations,
                r.seconds_total);
  } catch (const Exception& e) {
    std::printf("Dune reported error: %s\n", e.what());
    return 1;
  }
  return 0;
}
