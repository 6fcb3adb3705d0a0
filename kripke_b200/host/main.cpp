// kripke.exe for the B200: same command line and stdout blocks as the reference driver
// (src/kripke.cpp:174-525) -- input echo, decomposition, per-iteration particle counts, timers with
// the TIMER_NAMES/TIMER_DATA lines, figures of merit -- over the sm_100a hot path.
#include "Kripke.h"

#include <cstring>

int main(int argc, char **argv) {
  Kripke::Core::Comm::init(&argc, &argv);
  Kripke::Core::Comm comm;
  const int myid = (int)comm.rank();
  const int num_tasks = (int)comm.size();

  if (myid == 0) {
    printf("\nKripke hot path on NVIDIA B200 (sm_100a)  --  version %s\n", KRIPKE_VERSION);
    printf("A from-scratch implementation of LLNL/Kripke's source iteration; interface-compatible, not a port.\n\n");
    char name[128] = "";
    int sms = 0, maj = 0, min = 0;
    size_t free_b = 0, total_b = 0;
    KB200_CALL(kb200_device_info(name, sizeof(name), &sms, &maj, &min, &free_b, &total_b));
    printf("Compilation Options:\n");
    printf("  Architecture:           B200 (hand-written CUDA, sm_100a, C ABI v%d)\n", kb200_abi_version());
    printf("  Device:                 %s, %d SMs, cc %d.%d, %.1f GB HBM\n", name, sms, maj, min, total_b / 1e9);
    printf("  CHAI Enabled:           No\n  CUDA Enabled:           Yes\n  MPI Enabled:            No (NCCL, %d rank%s)\n",
           num_tasks, num_tasks == 1 ? "" : "s");
    printf("  OpenMP Enabled:         No\n  Caliper Enabled:        No\n");
  }

  InputVariables vars;
  if (parseCommandLine(argc, argv, vars)) {
    Kripke::Core::Comm::finalize();
    return 1;
  }
  if (vars.checkValues()) return 1;

  if (myid == 0) {
    printf("\nInput Parameters\n================\n\n  Problem Size:\n");
    printf("    Zones:                 %d x %d x %d  (%d total)\n", vars.nx, vars.ny, vars.nz, vars.nx * vars.ny * vars.nz);
    printf("    Groups:                %d\n", vars.num_groups);
    printf("    Legendre Order:        %d\n", vars.legendre_order);
    printf("    Quadrature Set:        ");
    if (vars.quad_num_polar == 0) printf("Dummy S2 with %d points\n", vars.num_directions);
    else printf("Gauss-Legendre, %d polar, %d azimuthal (%d points)\n", vars.quad_num_polar, vars.quad_num_azimuthal, vars.num_directions);
    printf("\n  Physical Properties:\n");
    printf("    Total X-Sec:           sigt=[%lf, %lf, %lf]\n", vars.sigt[0], vars.sigt[1], vars.sigt[2]);
    printf("    Scattering X-Sec:      sigs=[%lf, %lf, %lf]\n", vars.sigs[0], vars.sigs[1], vars.sigs[2]);
    printf("\n  Solver Options:\n    Number iterations:     %d\n", vars.niter);
    printf("\n  MPI Decomposition Options:\n");
    printf("    Total MPI tasks:       %d\n", num_tasks);
    printf("    Spatial decomp:        %d x %d x %d MPI tasks\n", vars.npx, vars.npy, vars.npz);
    printf("    Block solve method:    %s\n", vars.parallel_method == PMETHOD_SWEEP ? "Sweep" : "Block Jacobi");
    printf("\n  Per-Task Options:\n");
    printf("    DirSets/Directions:    %d sets, %d directions/set\n", vars.num_dirsets, vars.num_directions / vars.num_dirsets);
    printf("    GroupSet/Groups:       %d sets, %d groups/set\n", vars.num_groupsets, vars.num_groups / vars.num_groupsets);
    printf("    Zone Sets:             %d x %d x %d\n", vars.num_zonesets_dim[0], vars.num_zonesets_dim[1], vars.num_zonesets_dim[2]);
    printf("    Architecture:          %s\n", Kripke::archToString(vars.al_v.arch_v).c_str());
    printf("    Data Layout:           %s\n", Kripke::layoutToString(vars.al_v.layout_v).c_str());
  }

  {
    Kripke::Core::DataStore data_store;
    Kripke::generateProblem(data_store, vars);
    Kripke::SteadyStateSolver(data_store, vars.niter, vars.parallel_method == PMETHOD_BJ);

    auto &timing = data_store.getVariable<Kripke::Timing>("timing");
    timing.print();

    // figures of merit (src/kripke.cpp:495-516)
    size_t num_unknowns = data_store.getVariable<Kripke::Core::Set>("Set/Group").globalSize() *
                          data_store.getVariable<Kripke::Core::Set>("Set/Direction").globalSize() *
                          data_store.getVariable<Kripke::Core::Set>("Set/Zone").globalSize();
    size_t num_iter = timing.getCount("SweepSolver");
    double iter_time = timing.getTotal("Solve") / num_iter;
    double grind_time = iter_time / num_unknowns;
    double throughput = num_unknowns / iter_time;
    double sweep_eff = 100.0 * timing.getTotal("SweepSubdomain") / timing.getTotal("SweepSolver");
    if (myid == 0) {
      printf("\nFigures of Merit\n================\n\n");
      printf("  Throughput:         %e [unknowns/(second/iteration)]\n", throughput);
      printf("  Grind time :        %e [(seconds/iteration)/unknowns]\n", grind_time);
      printf("  Sweep efficiency :  %4.5lf [100.0 * SweepSubdomain time / SweepSolver time]\n", sweep_eff);
      printf("  Number of unknowns: %lu\n", (unsigned long)num_unknowns);
    }
    KB200_CALL(kb200_device_sync());
  }
  Kripke::Core::Comm::finalize();
  if (myid == 0) printf("\nEND\n");
  return 0;
}
