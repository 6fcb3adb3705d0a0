/* TEST INFRASTRUCTURE -- CPU oracle for the Kripke source-iteration hot path.
 *
 * A plain-C restatement of the reference's algorithm (LLNL/Kripke, /root/reference).  It exists
 * ONLY as the checker: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * it; the product (kripke_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity pinning: every function below is checked element-wise (bit-exact except the population
 * reduction) against the unmodified reference compiled by `make -C oracle ref`
 * (tests/test_oracle_vs_reference.py, skipped where /root/reference is absent) and against the
 * committed golden vectors in tests/golden/ that the reference itself produced
 * (tools/make_golden.py).
 */
#ifndef KRIPKE_ORACLE_H
#define KRIPKE_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout enum values follow src/Kripke/ArchLayout.h:92-101 */
enum { KO_DGZ = 0, KO_DZG, KO_GDZ, KO_GZD, KO_ZDG, KO_ZGD };

typedef struct {
  int nx, ny, nz;          /* --zones */
  int num_directions;      /* --quad N or P*A */
  int num_groups;          /* --groups */
  int legendre_order;      /* --legendre */
  int quad_num_polar;      /* --quad P:A (0 = dummy S2) */
  int quad_num_azimuthal;
  int layout;              /* --layout */
  int npx, npy, npz;       /* --procs: folded into the zone sets */
  int num_dirsets;         /* --dset */
  int num_groupsets;       /* --gset */
  int num_zonesets_dim[3]; /* --zset */
  double sigt[3];          /* --sigt */
  double sigs[3];          /* --sigs */
  int num_material_subsamples;
} ko_input;

typedef struct ko_problem ko_problem;

void ko_default_input(ko_input *in);             /* src/Kripke/InputVariables.cpp:17-47 */
int ko_check_input(const ko_input *in);          /* src/Kripke/InputVariables.cpp:52-109 ; 0 = ok */
ko_problem *ko_create(const ko_input *in);       /* src/Kripke/Generate.cpp:24-94 */
void ko_destroy(ko_problem *p);

/* kernels (whole-problem entry points, like Kripke::Kernel::*) */
void ko_zero(ko_problem *p, const char *field);  /* Kernel::kConst(field, 0.0) */
void ko_ltimes(ko_problem *p);                   /* src/Kripke/Kernel/LTimes.cpp:79-105 */
void ko_lplustimes(ko_problem *p);               /* src/Kripke/Kernel/LPlusTimes.cpp:68-92 */
void ko_scattering(ko_problem *p);               /* src/Kripke/Kernel/Scattering.cpp:112-164 */
void ko_source(ko_problem *p);                   /* src/Kripke/Kernel/Source.cpp:83-115 */
void ko_sweep_subdomain(ko_problem *p, int sdom);/* src/Kripke/Kernel/SweepSubdomain.cpp:115-124 */
void ko_sweep_solver(ko_problem *p, int block_jacobi); /* src/Kripke/SweepSolver.cpp:23-87 */
double ko_population(ko_problem *p);             /* src/Kripke/Kernel/Population.cpp:74-101 */
/* src/Kripke/SteadyStateSolver.cpp:24-115 ; particles[niter] receives the per-iteration count */
void ko_solve(ko_problem *p, int niter, int block_jacobi, double *particles);

/* introspection */
int ko_num_subdomains(const ko_problem *p);      /* SPACE_PQR */
/* order in which SweepSolver visits the subdomains (src/Kripke/SweepSolver.cpp:52-80) */
void ko_sweep_order(const ko_problem *p, int *order);
void ko_adjacency(const ko_problem *p, int sdom, int upwind[3], int downwind[3]);
/* number of chunks of a field, and pointer/size of a chunk (chunks in work-list order).
 * Field names are the reference's DataStore names ("psi", "phi", "data/sigs", "quadrature/w" ...).
 * elsize receives 8 (double) or 4 (int). Returns 0 on success. */
int ko_field_chunks(const ko_problem *p, const char *field);
int ko_field_chunk(const ko_problem *p, const char *field, int chunk, void **ptr, size_t *n, int *elsize);
/* sizes: which = "Gs","Ds","ni","nj","nk","M","L1","G","D","Z","SP","SQ","Sx","Sy","Sz" */
int ko_dim(const ko_problem *p, const char *which);
/* element offset inside a chunk for psi/rhs (d,g,z), phi/phi_out (nm,g,z), planes, sigt, sigs */
size_t ko_offset_flux(const ko_problem *p, int d, int g, int z);
size_t ko_offset_moment(const ko_problem *p, int nm, int g, int z);

#ifdef __cplusplus
}
#endif
#endif
