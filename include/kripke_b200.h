/*
 * kripke_b200.h -- C ABI of the B200 (sm_100a) implementation of Kripke's source-iteration hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point
 * names the reference interface it replaces (paths relative to the LLNL/Kripke source tree).
 * The C++ host layer in kripke_b200/host/ (namespace Kripke, same signatures as the reference's
 * Kripke::Kernel::* / SweepSolver / ParallelComm / DataStore) is the only in-tree caller; a
 * maintainer of the reference would bind exactly these symbols.
 *
 * Conventions
 *  - every function returns 0 on success, otherwise a non-zero code; kb200_last_error() gives the
 *    message (thread-local).  The C++ wrapper turns non-zero into KRIPKE_ABORT semantics.
 *  - all pointers named d_* / inside descriptors are DEVICE pointers of the current device,
 *    h_* are host pointers.  Fields are never allocated behind the caller's back; the library keeps
 *    a few private scratch buffers that live until kb200_finalize: the device copies of the
 *    descriptors, the per-zone material fractions of kb200_scattering, and -- only for the
 *    nestings whose moments are not zone-fastest -- two transposed copies of a moments chunk.
 *    kb200_free keeps blocks up to 64 MB in a pool for the next kb200_alloc of the same size; while streams created
 *    with kb200_stream_create exist it synchronises the device first (like cudaFree), otherwise the library stream orders
 *    the old and the new owner.
 *  - `layout` is the reference's LayoutV value: 0=DGZ 1=DZG 2=GDZ 3=GZD 4=ZDG 5=ZGD
 *    (src/Kripke/ArchLayout.h:92-101); storage orders follow src/Kripke/VarTypes.h:73-101.
 *  - `stream` is a cudaStream_t passed as void* (NULL = the library's default stream).
 *  - there is NO CPU fallback: without a visible sm_100 device kb200_init fails.
 */
#ifndef KRIPKE_B200_H
#define KRIPKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB200_ABI_VERSION 1
#define KB200_MAX_DIRSETS 64 /* direction-set chunks fused into one LTimes / LPlusTimes launch */

typedef void *kb200_stream_t;
typedef void *kb200_event_t;

/* ---- runtime: replaces Core::Comm::init/finalize (src/Kripke/Core/Comm.h:26-60) and the
 *      allocation in Core::FieldStorage (src/Kripke/Core/Field.h:61-104) ------------------- */
int kb200_abi_version(void);
const char *kb200_last_error(void);
int kb200_device_count(int *count);
int kb200_init(int device);           /* bind this process to one GPU, create the default stream */
int kb200_finalize(void);
int kb200_device_info(char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor,
                      size_t *free_bytes, size_t *total_bytes);
int kb200_alloc(size_t bytes, void **d_ptr);      /* Field.h:69  new ElementType[sdom_size] */
int kb200_free(void *d_ptr);                      /* Field.h:98-104 */
int kb200_pool_trim(void);                        /* return the pooled free blocks (at most 4 GB) to the driver */
int kb200_alloc_host(size_t bytes, void **h_ptr); /* pinned host memory for mirrors / e2e staging */
int kb200_free_host(void *h_ptr);
int kb200_upload(void *d_dst, const void *h_src, size_t bytes, kb200_stream_t stream);
int kb200_download(void *h_dst, const void *d_src, size_t bytes, kb200_stream_t stream);
int kb200_copy(void *d_dst, const void *d_src, size_t bytes, kb200_stream_t stream); /* Kernel::kCopy, Kernel.h:60-81; ParallelComm.cpp:149-161 */
int kb200_fill_f64(double *d_ptr, double value, size_t n, kb200_stream_t stream);    /* Kernel::kConst, Kernel.h:37-55 */
int kb200_memset(void *d_ptr, int byte_value, size_t bytes, kb200_stream_t stream);  /* kConst on fields of any element size */
int kb200_stream_create(kb200_stream_t *stream);
int kb200_stream_destroy(kb200_stream_t stream);
int kb200_stream_sync(kb200_stream_t stream);
int kb200_device_sync(void);
int kb200_event_create(kb200_event_t *ev);
int kb200_event_destroy(kb200_event_t ev);
int kb200_event_record(kb200_event_t ev, kb200_stream_t stream);
int kb200_event_sync(kb200_event_t ev);
int kb200_event_elapsed_ms(kb200_event_t start, kb200_event_t stop, float *ms);
int kb200_stream_wait_event(kb200_stream_t stream, kb200_event_t ev);
/* number of kernels this library has launched since the last reset (bench.py "gpu_launches") */
int kb200_launch_count(uint64_t *count, int reset);
/* 1: every kernel keeps multiplies and adds separate and sums in the reference's order, so all
 * fields are bit-identical to the reference's CPU path (test mode, ~2x fp64 instructions);
 * 0 (default, or env KB200_EXACT unset): products are contracted to DFMA (differences ~1e-16 relative). */
int kb200_set_exact(int on);

/* ---- LTimes: Kripke::Kernel::LTimes (src/Kripke/Kernel/LTimes.cpp:79-105, body :54-65) ----
 * phi(nm,g,z) (+)= sum_q sum_d ell_q(nm,d) * psi_q(d,g,z) for the `nsets` direction-set chunks
 * q (ascending) that share one phi chunk (same group set and zone set).  ell is stored [d][nm]
 * in every layout.  accumulate=0 overwrites phi (folds the caller's kConst(phi,0),
 * SteadyStateSolver.cpp:53), accumulate=1 is the reference's "+=".
 */
typedef struct {
  int layout, M, Ds, Gs, Zs, nsets, accumulate;
  const double *ell[KB200_MAX_DIRSETS]; /* [Ds][M] each */
  const double *psi[KB200_MAX_DIRSETS]; /* [Ds x Gs x Zs] in `layout` order */
  double *phi;                          /* [M x Gs x Zs] in `layout` order */
} kb200_ltimes_desc;
int kb200_ltimes(const kb200_ltimes_desc *h_descs, int n, kb200_stream_t stream);
/* launches of the producer/consumer-group LTimes kernel (kb200_moments_slab.cu: M = 25, zone- or group-fastest columns) so far */
unsigned long long kb200_ltimes_slab_launches(void);

/* ---- LPlusTimes: Kripke::Kernel::LPlusTimes (src/Kripke/Kernel/LPlusTimes.cpp:68-92, body :49-60)
 * rhs_q(d,g,z) (+)= sum_nm ell_plus_q(d,nm) * phi_out(nm,g,z) for each direction-set chunk q.
 */
typedef struct {
  int layout, M, Ds, Gs, Zs, nsets, accumulate;
  const double *ell_plus[KB200_MAX_DIRSETS]; /* [Ds][M] each */
  const double *phi_out;                     /* [M x Gs x Zs] */
  double *rhs[KB200_MAX_DIRSETS];            /* [Ds x Gs x Zs] each */
} kb200_lplustimes_desc;
int kb200_lplustimes(const kb200_lplustimes_desc *h_descs, int n, kb200_stream_t stream);
/* launches so far of the producer/consumer-group kernel as a plain product (M >= 36, DGZ/DZG, outputs a multiple of 32) */
unsigned long long kb200_lplustimes_slab_launches(void);

/* ---- Scattering: Kripke::Kernel::scattering (src/Kripke/Kernel/Scattering.cpp:112-164, body :73-99)
 * For one zone set R and one destination group set: phi_out(nm,g,z) (+)= sum over the `nsrc`
 * source group-set chunks (ascending) and gp of sigs_z(n(nm), g+glower_dst, gp+glower_src, z) *
 * phi_src(nm,gp,z), sigs_z = sum_{mix in zone z} sigs(mat(mix),n,G,GP) * fraction(mix).
 * sigs is the global [3][L+1][G][G] table in `layout` order (VarTypes.h:44).
 */
typedef struct {
  int layout, M, L1, G, Gs, Zs, nsrc, glower_dst, accumulate;
  int glower_src[KB200_MAX_DIRSETS];
  const double *phi_src[KB200_MAX_DIRSETS]; /* [M x Gs x Zs] each */
  double *phi_out;                          /* [M x Gs x Zs] */
  const double *sigs;                       /* [3 x L1 x G x G] */
  const int *moment_to_legendre;            /* [M] */
  const int *zone_to_mixelem;               /* [Zs] */
  const int *zone_to_num_mixelem;           /* [Zs] */
  const int *mixelem_to_material;           /* [num_mixelem] */
  const double *mixelem_to_fraction;        /* [num_mixelem] */
} kb200_scattering_desc;
int kb200_scattering(const kb200_scattering_desc *h_descs, int n, kb200_stream_t stream);
/* The same with Kernel::source (Kernel/Source.cpp:59-75) folded into the epilogue of the tensor-core kernel: moment 0 also
 * receives strength * (volume fraction of material 0 in the zone).  *folded = 1: done, kb200_source must not follow;
 * *folded = 0: the kernel in use cannot fold (bit-exact mode, odd shapes) and only the scattering was done. */
int kb200_scattering_source(const kb200_scattering_desc *h_descs, int n, double strength, int *folded, kb200_stream_t stream);
/* What a scattering call would do with this descriptor list, without a device: 0 = the one-read kernel takes it (descriptor
 * groups sharing their source chunks, destination groups per CTA, sibling CTAs per tile sequence), -1 = declined. */
int kb200_scattering_plan(const kb200_scattering_desc *h_descs, int n, int *ngroups, int *outputs_per_cta, int *siblings);
/* kernel family that served the last scattering call: "slab", "mma", "transposed+slab", "transposed+mma" or "dfma" */
const char *kb200_last_scattering_kernel(void);

/* ---- Source: Kripke::Kernel::source (src/Kripke/Kernel/Source.cpp:83-115, body :59-75) ----
 * phi_out(0,g,zone(mix)) += strength * fraction(mix) for every mixelem of material 0.
 */
typedef struct {
  int layout, M, Gs, Zs, num_mixelem;
  double strength;
  const int *mixelem_to_zone;
  const int *mixelem_to_material;
  const double *mixelem_to_fraction;
  double *phi_out;
} kb200_source_desc;
int kb200_source(const kb200_source_desc *h_descs, int n, kb200_stream_t stream);

/* ---- Population: Kripke::Kernel::population (src/Kripke/Kernel/Population.cpp:74-101, body :49-63)
 * *d_result (+)= sum over the n chunks of sum_{d,g,z} (w(d)*psi(d,g,z))*volume(z).  d_result is a
 * device double; d_scratch must hold kb200_population_scratch_doubles() doubles.
 */
typedef struct {
  int layout, Ds, Gs, Zs;
  const double *w;      /* [Ds] */
  const double *volume; /* [Zs] */
  const double *psi;    /* [Ds x Gs x Zs] */
} kb200_population_desc;
size_t kb200_population_scratch_doubles(void);
int kb200_population(const kb200_population_desc *h_descs, int n, double *d_scratch, double *d_result,
                     kb200_stream_t stream);

/* ---- Sweep: Kripke::Kernel::sweepSubdomain (src/Kripke/Kernel/SweepSubdomain.cpp:115-124,
 *      body :34-109) and its batched form used by Kripke::SweepSolver (src/Kripke/SweepSolver.cpp:52-80).
 * Diamond-difference upwind sweep of one subdomain.  On entry i/j/k_plane hold the incoming
 * (upwind) face fluxes, on exit the outgoing ones (in place, like the reference).  If
 * inflow_zero[dim] is set the incoming face of that dimension is vacuum and the plane is not
 * read (folds SweepSolver.cpp:64-72's kConst).  If out_plane[dim] is non-NULL the outgoing face is
 * ALSO written there (the downwind subdomain's plane chunk, possibly peer-GPU memory), which
 * folds ParallelComm::postSends' copy (src/Kripke/ParallelComm.cpp:149-161).
 * All n subdomains of one call must be mutually independent (no upwind relation among them).
 */
typedef struct {
  int layout, Ds, Gs, ni, nj, nk;
  int id, jd, kd;               /* sweep direction (+1/-1) of direction 0 of the set */
  const double *xcos, *ycos, *zcos; /* [Ds] */
  const double *dx, *dy, *dz;       /* [ni], [nj], [nk] */
  const double *sigt;               /* sigt_zonal chunk [Gs x Zs] */
  const double *rhs;                /* [Ds x Gs x Zs] */
  double *psi;                      /* [Ds x Gs x Zs] */
  double *i_plane, *j_plane, *k_plane;
  int inflow_zero[3];
  double *out_plane[3];
} kb200_sweep_desc;
int kb200_sweep(const kb200_sweep_desc *h_descs, int n, kb200_stream_t stream);
/* Same sweep, and -- where the kernel in use supports it -- Kernel::population's sum over the n swept
 * subdomains is accumulated while psi is still in registers (saves re-reading psi,
 * 8 bytes per unknown).  h_w[i] / h_volume[i] are the DEVICE pointers of quadrature/w [Ds] and volume
 * [Zs] of subdomain i.  On return *count doubles have been written to d_partials (capacity >= count is
 * checked); their sum in index order is sum_i sum_{d,g,z} (w*psi)*volume.  *count == 0 means the fused
 * path did not apply and only the sweep was done: the caller then uses kb200_population. */
int kb200_sweep_population(const kb200_sweep_desc *h_descs, int n, const double *const *h_w, const double *const *h_volume,
                           double *d_partials, int capacity, int *count, kb200_stream_t stream);
/* Same again with a hint: h_vol_const[i] > 0 promises that EVERY entry of h_volume[i] equals that value (Kripke's
 * generator only produces uniform meshes, src/Kripke/Generate/Space.cpp:110-136), so the fused sum needs no volume loads;
 * h_vol_const[i] == 0 (or h_vol_const == NULL) reads the array. */
int kb200_sweep_population_uniform(const kb200_sweep_desc *h_descs, int n, const double *const *h_w, const double *const *h_volume,
                                   const double *h_vol_const, double *d_partials, int capacity, int *count, kb200_stream_t stream);
/* name of the kernel family that served the most recent sweep call ("irow", "irow+population", "pencil", "pencil+population", "zline", "elem", "tile";
 * "none" before the first call): the fast paths cover a subset of shapes, and a benchmark line must say which one ran */
const char *kb200_last_sweep_kernel(void);
/* *d_result = sum of d_partials[0..n) in index order (fixed-order tree, deterministic) */
int kb200_population_reduce(const double *d_partials, int n, double *d_result, kb200_stream_t stream);
/* *h_value = the common value of d_v[0..n) if all n doubles are equal and positive, else 0 (synchronises the stream; the host
 * layer uses it to recognise a uniform zone-volume field for kb200_sweep_population_uniform without downloading the field) */
int kb200_uniform_positive_value(const double *d_v, size_t n, double *h_value, kb200_stream_t stream);

/* ---- layout transform (remaining nestings via transform) --------------------------
 * Re-orders a 3-index field (a,b,c extents in canonical <Direction|Moment, Group, Zone> order)
 * from src_layout to dst_layout storage. */
int kb200_layout_transform(int src_layout, int dst_layout, int n_dm, int n_g, int n_z,
                           const double *d_src, double *d_dst, kb200_stream_t stream);

/* ---- exchange: replaces ParallelComm::postSends/postRecvs/testRecieves (src/Kripke/ParallelComm.cpp:61-251)
 *      and Core::Comm::allReduceSumDouble (src/Kripke/Core/Comm.h:161-166).  NCCL point-to-point over
 *      NVLink; one process per GPU.  The 128-byte unique id is produced on rank 0 and distributed by
 *      the launcher (torch.distributed / a file), like MPI_Init's out-of-band wire-up. */
int kb200_comm_unique_id(void *id128);
int kb200_comm_init(int rank, int nranks, const void *id128);
int kb200_comm_destroy(void);
int kb200_comm_rank(int *rank, int *nranks);
int kb200_comm_group_start(void);
int kb200_comm_group_end(void);
int kb200_comm_send(const double *d_buf, size_t count, int peer, kb200_stream_t stream); /* MPI_Isend, ParallelComm.cpp:176 */
int kb200_comm_recv(double *d_buf, size_t count, int peer, kb200_stream_t stream);       /* MPI_Irecv, ParallelComm.cpp:106 */
int kb200_comm_allreduce_sum_f64(double *d_buf, size_t count, kb200_stream_t stream);    /* Comm.h:161-166 */
int kb200_comm_allreduce_sum_i64(long long *d_buf, size_t count, kb200_stream_t stream); /* Comm.h:121-130 */

int kb200_comm_allgather(const void *h_send, size_t bytes_per_rank, void *h_recv);       /* setup: Comm.h / Set.cpp all-gathers */
int kb200_comm_barrier(kb200_stream_t stream);                                           /* stream-ordered, no host sync */

/* ---- exchange over NVLink peer memory (CUDA IPC), the path SweepSolver uses between GPUs of one node -------------
 * The receiver exports its plane chunks (and one array of flags) once; the sender maps them and passes the mapped
 * pointer as kb200_sweep_desc.out_plane, so the sweep kernel itself stores the outgoing faces into the downwind
 * subdomain's plane chunk on the other GPU (ParallelComm.cpp:99-107 "receive straight into the plane chunk", :170-178).
 * kb200_p2p_signal raises flags in peer memory after all earlier work of the stream, kb200_p2p_wait holds later work of
 * the stream until local flags have reached a value (MPI_Testany, :222-230) -- device side, no host synchronisation. */
int kb200_ipc_export(const void *d_ptr, void *handle64);        /* d_ptr: base of a kb200_alloc block */
int kb200_ipc_open(const void *handle64, void **d_peer_ptr);
int kb200_ipc_close(void *d_peer_ptr);
int kb200_p2p_signal(unsigned *const *h_flags, int n, unsigned value, kb200_stream_t stream);
int kb200_p2p_wait(const unsigned *const *h_flags, int n, unsigned value, kb200_stream_t stream);

/* ---- problem generator on the device: the zone loops of Kripke::generateProblem that take seconds on the host at 128^3
 *      zones (src/Kripke/Generate/Space.cpp:189-363).  Bit-identical to the host generator.
 * kb200_generate_mix_count: material volume fractions of every zone by ns^3 sub-samples (d_frac [Zs][3]), materials per
 *   zone (d_z2n = zone_to_num_mixelem), their exclusive prefix sum (d_z2m = zone_to_mixelem), the number of mixelems and
 *   the three material volumes of this zone set (host outputs; synchronises the stream).
 * kb200_generate_mix_fill: mixelem_to_zone / _material / _fraction in zone order, materials ascending inside a zone.
 * kb200_generate_sigt: sigt_zonal(g,z) = sum over the zone's mixelems of fraction * sigt[material], stored in `layout`
 *   order ([g][z] for DGZ, GDZ, GZD; [z][g] otherwise). */
typedef struct {
  int ni, nj, nk;                          /* zones of this zone set */
  int i0, j0, k0;                          /* global index of its first zone */
  double x_min, y_min, z_min, dx, dy, dz;  /* mesh origin and (uniform) zone widths */
  int ns;                                  /* sub-samples per axis (InputVariables::num_material_subsamples) */
} kb200_zoneset_desc;
int kb200_generate_mix_count(const kb200_zoneset_desc *zs, double *d_frac, int *d_z2n, int *d_z2m, int *h_num_mixelem,
                             double *h_material_volume3, kb200_stream_t stream);
int kb200_generate_mix_fill(int Zs, const double *d_frac, const int *d_z2m, int *d_m2z, int *d_m2m, double *d_m2f,
                            kb200_stream_t stream);
int kb200_generate_sigt(int layout, int Gs, int Zs, const int *d_z2m, const int *d_z2n, const int *d_m2m, const double *d_m2f,
                        const double *h_sigt3, double *d_sigt, kb200_stream_t stream);
int kb200_device_bound(void);  /* 1 once kb200_init has bound this process to a GPU */

/* ---- micro-benchmarks used for the roofline denominators (bench.py / tools) ------------------ */
int kb200_peak_fp64_gflops(int use_dmma, int iters, double *gflops); /* DFMA / DMMA issue peak */
int kb200_peak_copy_gbs(size_t bytes, int iters, double *gbs);       /* STREAM-style copy */

#ifdef __cplusplus
}
#endif
#endif /* KRIPKE_B200_H */
