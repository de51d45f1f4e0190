/*
 * emc.h -- C-ABI of the B200-native ensemble Monte Carlo hot path.
 *
 * This is the drop-in boundary for the particle loop of
 * edsolibet/Monte_Carlo_Carrier_Transport.  The reference has no FFI layer: its
 * boundary is the Python function surface of main_.py / scatter_.py /
 * poisson_.py, so every entry point below names the reference function (file:line
 * into the reference repository) whose work it replaces.  The reference-side
 * binding (a ctypes stub) is shown in INTEGRATION.md and shipped as
 * monte_carlo_carrier_transport_b200/abi.py.
 *
 * Conventions
 *   - plain C: pointers, sizes and POD structs only; no C++ or torch types.
 *   - every function returns an int status: 0 = ok, < 0 = EMC_ERR_* (never throws);
 *     emc_last_error() returns the message of the last failure.
 *   - "device pointer" arguments point into memory owned by the CALLER (torch
 *     tensors in the Python host); the library never allocates particle-sized
 *     memory.  All launches are asynchronous on the caller's stream
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream) unless a
 *     function returns a value to the HOST, in which case it synchronises that
 *     stream before returning.
 *   - one EmcCtx per (process, device); not thread-safe without external locking
 *     (the reference is single-threaded).
 *   - particle state is structure-of-arrays float64: kx ky kz [1/m], x y z [m],
 *     dtau [s] (time left of the current free flight, main_.py:337,391) and an
 *     int32 valley (0 = Gamma, 1 = L, 2 = X).  A particle the reference would have
 *     raised on (scatter_.py:45,47,1079, NaN propagation) gets
 *     valley = -(1 + EMC_FAULT_*) and is frozen; faults are counted in EmcObs.
 */
#ifndef EMC_H
#define EMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMC_ABI_VERSION 1
#define EMC_MAX_MECH 16

/* status codes */
enum {
    EMC_OK = 0,
    EMC_ERR_INVALID = -1,      /* bad argument */
    EMC_ERR_CUDA = -2,         /* a CUDA runtime call failed */
    EMC_ERR_NO_TABLES = -3,    /* emc_upload_tables has not been called */
    EMC_ERR_DIVERGED = -4,     /* SOR update >= 1e3, the reference's "Too large du" (poisson_.py:78-80) */
    EMC_ERR_UNSUPPORTED = -5
};

/* polar-angle samplers (scatter_.py:795-889) */
enum { EMC_SAMPLER_SELF = 0, EMC_SAMPLER_ISO = 1, EMC_SAMPLER_POP_ABS = 2,
       EMC_SAMPLER_POP_EMS = 3, EMC_SAMPLER_ION = 4 };

/* per-particle fault codes (valley = -(1+code)) */
enum { EMC_FAULT_NAN = 0,        /* scatter_.py:45  "Nan energy" */
       EMC_FAULT_ZERO_E = 1,     /* scatter_.py:47  zero energy */
       EMC_FAULT_NO_MECH = 2,    /* no column with cum > r2 (cannot happen: last column is 1) */
       EMC_FAULT_DOMAIN = 3,     /* NaN final state (arcsin/arccos/sqrt domain), the reference propagates NaN */
       EMC_FAULT_NEG_ENERGY = 4, /* E0 + dE < 0 -> sqrt(negative) at scatter_.py:1071 */
       EMC_FAULT_ZERO_K = 5,     /* scatter_.py:1079 exact-zero component */
       EMC_FAULT_STREAM = 6,     /* replay stream exhausted */
       EMC_FAULT_OUTSIDE = 7 };  /* init_.where_am_i raised "Point is outside all layers" (init_.py:46,56) */

/* charge assignment schemes */
enum { EMC_DEPOSIT_NGP = 0,      /* poisson_.py:30-38 (inclusive box test, face hits count twice) */
       EMC_DEPOSIT_CIC = 1 };    /* the commented weight at poisson_.py:34, 2^-32 fixed point */

/* Poisson sweep orders */
enum { EMC_SOR_LEXICOGRAPHIC = 0, /* poisson_.py:74 order, executed as hyperplane wavefronts: bit-identical iterates */
       EMC_SOR_RED_BLACK = 1,     /* same fixed point, parallel colouring */
       EMC_SOR_MULTIGRID = 2 };   /* same equation and fixed point, geometric multigrid V(2,2) cycles (emc_poisson_mg):
                                     `tol` is tested on the TRUE residual / |e| (= the Gauss-Seidel update the reference's
                                     rule looks at, poisson_.py:76,82), max_sweeps counts V-cycles, omega0 is unused */

/* field sources of the flight kernel */
enum { EMC_FIELD_CONSTANT = 0,   /* dev.elec_field, main_.py:367,373,388 */
       EMC_FIELD_GROUPS = 1,     /* particle i uses field_groups[(pid_offset+i)/group_size] (velocity-field sweeps) */
       EMC_FIELD_MESH = 2,       /* nearest-cell gather from the Poisson field (poisson_.py:88-96), V/m */
       EMC_FIELD_MESH_PACKED = 3,   /* same field, one 32-byte record per cell written by emc_field_packed:
                                    `ex` points to it, ey = ez = NULL */
       EMC_FIELD_MESH_XZ = 4 };  /* ny = 1 meshes: one 16-byte (ex, ez) record per INTERIOR cell, unpadded (nx, nz),
                                    written by emc_field_xz (ey is identically -0 there, poisson_.py:95 between two
                                    halo zeros): `ex` points to it, ey = ez = NULL.  Identical trajectories. */

/* Physical / numerical parameters: a snapshot of init_.Parameters, init_.GaAs and
 * init_.Device (init_.py:58-157). */
typedef struct EmcParams {
    double dt;            /* init_.py:60  */
    double gamma;         /* init_.py:83  GaAs.tot_scat (self-scattering ceiling) */
    double hbar;          /* init_.py:68  */
    double q;             /* init_.py:67  */
    double mass[3];       /* init_.py:87  */
    double nonparab[3];   /* init_.py:89  */
    double tot_dim[3];    /* init_.py:147 */
    double efield[3];     /* init_.py:133, V/cm */
    double e_phonon;      /* init_.py:79  */
    double ion_eb;        /* scatter_.py:868-869  q/(2*eps_0*b), b = (1/(2*dope))**(1/3) */
    double kT;            /* init_.py:64 */
    double mean_energy;   /* init_.py:145 (in units of kT) */
    /* mesh (init_.py:135-150) */
    int32_t seg[3];       /* cells per axis */
    int32_t z_cyclic;     /* 0 = the reference's specular z walls (main_.py:140-154); 1 = extension for bulk
                             runs (BASELINE configs[1]): z wraps like x and y (main_.py:125-138) */
    double dl[3];         /* cell size */
    double rho_background;/* init_.py:189-191: -dope*prod(dl)*1e6 per cell */
    double rho_increment; /* poisson_.py:36: dope*vol*1e6/num_carr per super-particle */
    double eps_s;         /* GaAs.eps_0, the permittivity used at poisson_.py:76 */
    double surf_val;      /* Dirichlet value of the z-min face, poisson_.py:51-52 */
} EmcParams;

/* One init_.Layer (init_.py:15-34) with the per-material numbers the particle loop reads through
 * `mat[mat_i]` / `scat_tables[mat_i]` (main_.py:366-396): a multi-layer device stacks layers
 * along z and init_.where_am_i (init_.py:36-56) picks the layer of a particle from its z. */
#define EMC_MAX_LAYERS 4
typedef struct EmcLayer {
    double lz;            /* Layer.lz: thickness along z [m] */
    double mass[3];       /* matl.mass */
    double nonparab[3];   /* matl.nonparab */
    double e_phonon;      /* matl.E_phonon */
    double ion_eb;        /* scatter_.py:868-869 with matl.dope */
    double init_weight;   /* share of the initial ensemble placed in this layer by emc_init_ensemble
                             (extension; all weights 0 = the reference's uniform z, init_.py:200) */
} EmcLayer;

/* Ensemble observables of one timestep (main_.py:394-403), as raw sums so that
 * ranks can be combined with an integer/double all-reduce. */
typedef struct EmcObs {
    double sum_z, sum_vz, sum_e;   /* main_.py:398-400 numerators */
    double sum_vz2, sum_e2;        /* for the standard errors */
    int64_t n_valley[3];           /* main_.py:401-403 */
    int64_t n_fault;               /* frozen particles */
    int64_t n_segments;            /* carrier_drift calls: the "flight+scatter steps" of the headline metric */
    int64_t n_events;              /* scatter_.scatter calls */
} EmcObs;

typedef struct EmcCtx EmcCtx;

/* ---- lifetime ------------------------------------------------------------------------- */
int emc_abi_version(void);
/* Creates a context on CUDA device `device`.  Fails (EMC_ERR_CUDA) when no device exists:
 * there is no CPU fallback. */
int emc_create(const EmcParams *params, int device, EmcCtx **out);
void emc_destroy(EmcCtx *ctx);
/* message of the last failure on ctx (or of the last failed emc_create when ctx == NULL) */
const char *emc_last_error(const EmcCtx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t emc_launch_count(const EmcCtx *ctx);

/* ---- tables: scatter_.calc_scatter_table (scatter_.py:954-984) + scatter_mech (:892-951) -- */
/* HOST pointers.  ek: energy grid (n_energy), must be np.linspace(0, e_max, n_energy)
 * (init_.py:160-166).  cum[v]: row-major (n_energy, n_mech[v]) cumulative table of valley v.
 * dE/trans/sampler: [3][EMC_MAX_MECH] mechanism metadata.  The tables are copied to the
 * device (2.7 MB, L2-resident). */
int emc_upload_tables(EmcCtx *ctx, const double *ek, int32_t n_energy,
                      const double *cum0, const double *cum1, const double *cum2,
                      const int32_t n_mech[3], const double *dE, const int32_t *trans,
                      const int32_t *sampler);

/* Multi-layer devices (dev.layers, init_.py:127-130; main_.py:366,372,377,387,393 look the layer
 * up with where_am_i before every drift, scatter and observable).  emc_set_layers replaces the
 * single material of EmcParams by n_layers (1..EMC_MAX_LAYERS) stacked layers; the tables of each
 * layer (scat_tables[mat_i], main_.py:66) are then uploaded with emc_upload_tables_layer, same
 * arguments as emc_upload_tables plus the layer index; all layers share one energy grid and the
 * per-valley column counts.  sum(lz) should equal EmcParams.tot_dim[2]; a particle whose z falls
 * outside every layer (where_am_i raises) is frozen with EMC_FAULT_OUTSIDE. */
int emc_set_layers(EmcCtx *ctx, const EmcLayer *layers, int32_t n_layers);
int emc_upload_tables_layer(EmcCtx *ctx, int32_t layer, const double *ek, int32_t n_energy,
                            const double *cum0, const double *cum1, const double *cum2,
                            const int32_t n_mech[3], const double *dE, const int32_t *trans,
                            const int32_t *sampler);
/* When are the observables of emc_flight_scatter / emc_flight_scatter_replay taken?
 *   0 (default): after the timestep, as main_.py:394-403 does;
 *   1: BEFORE it -- the sums and valley counts describe the ensemble as the call found it (i.e. the
 *      result of the previous timestep), n_segments / n_events still count the call's own work.
 * In mode 1 the reduction runs in the kernel's load phase with all 32 lanes of a warp busy instead
 * of at particle retirement (17 of 32 on average): a time loop gets the same rows one launch later
 * and reads the last one with emc_observables. */
int emc_set_obs_phase(EmcCtx *ctx, int32_t before_step);
/* Composition of the inverse-trigonometric round trips of scatter_.py:1061-1062 and the polar
 * samplers (:811-872).  0 (default; EMC_EXACT_LIBM=1 in the environment makes 1 the default): the
 * algebraic equivalents cos(acos c) = c, sin(acos c) = sqrt((1-c)(1+c)), cos(2 atan t) = (1-t^2)/(1+t^2);
 * 1: acos / asin / atan / sin / cos composed exactly like the reference (about twice the fp64 work;
 * 6x fewer particles beyond 1e-9 in the ill-conditioned corner). */
int emc_set_exact_libm(EmcCtx *ctx, int32_t on);
/* Writes a one-line description of the run-time knobs that select kernels on this context
 * ("exact_libm=0 zonly=1 field_ahead=1 waves=1 ...", NUL-terminated, at most cap bytes) into buf
 * (HOST): what bench.py records next to its numbers.  Returns the length written. */
int emc_describe(const EmcCtx *ctx, char *buf, int32_t cap);
/* Material (layer) used by the single-function batches emc_calc_energy / emc_scatter: the `mat` /
 * `mat_i` argument of scatter_.calc_energy (scatter_.py:19) and scatter_.scatter (:986).  Default 0. */
int emc_select_layer(EmcCtx *ctx, int32_t layer);
/* Background charge per z-cell for emc_mesh_to_rho (extension for doping profiles: the reference
 * has one uniform background, init_.py:189-191).  HOST pointer to nz doubles; NULL restores the
 * uniform EmcParams.rho_background. */
int emc_set_background_profile(EmcCtx *ctx, const double *background_z, int32_t nz);

/* constant field (V/cm) used by EMC_FIELD_CONSTANT; replaces EmcParams.efield */
int emc_set_efield(EmcCtx *ctx, const double efield[3]);
/* EMC_FIELD_GROUPS table: HOST pointer to n_groups x 3 doubles (V/cm) */
int emc_set_field_groups(EmcCtx *ctx, const double *efields, int32_t n_groups, int64_t group_size);

/* ---- ensemble initialisation: init_.init_coords (init_.py:194-234) + main_.py:53,337 ------ */
/* Device pointers.  Philox4x32-10 keyed by `seed`, counter (draw/2, 0xFFFFFFFF, pid). */
int emc_init_ensemble(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y,
                      double *z, double *dtau, int32_t *valley, int64_t n, uint64_t seed,
                      uint64_t pid_offset, void *stream);

/* ---- photo-excited carriers: init_.init_photoex (init_.py:244-280) + main_.py:345-351 ---------- */
/* Draws n_candidates carriers -- depth z ~ Exp(z_scale) (the reference: min(1/alpha, slab)),
 * x, y ~ Normal(0, xy_std) (laser_std), mono-energetic `energy` [eV] with Normal(0, 2 pi) angles,
 * valley 0, first free flight (-1/gamma_first) log u (the reference: 1E15) -- DROPS the ones
 * beyond the slab (init_.py:247-248) and writes the survivors, in candidate order, to the
 * pointers given (the free tail of the caller's state arrays, room for `capacity` >= n_candidates
 * particles).  *n_accepted (HOST) receives how many were written; synchronises the stream.
 * Philox keyed by (seed, candidate_offset + candidate): independent of launch shape. */
int emc_init_photoex(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y,
                     double *z, double *dtau, int32_t *valley, int64_t capacity,
                     int64_t n_candidates, double energy, double z_scale, double xy_std,
                     double gamma_first, uint64_t seed, uint64_t candidate_offset,
                     int64_t *n_accepted, void *stream);

/* ---- the particle loop of one timestep: main_.py:360-397 --------------------------------- */
/* free_flight (main_.py:114) + carrier_drift/cyclic_boundary/specular_refl (:125-213) +
 * scatter_.scatter (scatter_.py:986-1099), fused, one call per timestep for the whole
 * ensemble.  Random numbers: in-kernel Philox4x32-10, key = seed, counter =
 * (draw/2, step, pid_offset+i).  field_mode selects EMC_FIELD_*; ex/ey/ez are the padded
 * (nx+2,ny+2,nz+2) field arrays of emc_field (only for EMC_FIELD_MESH, else NULL).
 * obs (device pointer to one EmcObs, may be NULL): if given, the observables of
 * main_.py:394-403 are reduced in the same kernel and written there.
 * mesh (device int64[nx*ny*nz], may be NULL): if given, the end-of-step NGP charge
 * assignment (poisson_.py:30-38) is fused into the kernel (atomically ADDED to mesh).
 * At most 2^32 - 65 particles per call (EMC_ERR_UNSUPPORTED beyond: the kernel indexes the slice with
 * 32 bits); larger ensembles are sliced by the caller, pid_offset keeps the random streams global. */
int emc_flight_scatter(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y,
                       double *z, double *dtau, int32_t *valley, int64_t n,
                       int32_t field_mode, const double *ex, const double *ey, const double *ez,
                       uint64_t seed, uint64_t pid_offset, uint32_t step,
                       EmcObs *obs, int64_t *mesh, void *stream);

/* n_steps consecutive timesteps (first_step, first_step + 1, ...) enqueued by ONE call: the time
 * loop of main_.py:339 without a host round trip per step (a small ensemble -- the reference's
 * 10^4 electrons take 12 us of GPU time per timestep -- is otherwise bound by the caller's own
 * per-call overhead).  obs_rows: device array of n_steps EmcObs (or NULL), row s written by
 * timestep first_step + s according to emc_set_obs_phase.
 * With obs_rows == NULL (no per-step observables wanted: burn-in phases, sweeps that sample every
 * m-th timestep) and the lean case -- constant or group field with Ex = Ey = +0, one layer, default
 * libm composition, the reference's table layout -- all n_steps (2 .. 65535) timesteps run in ONE
 * persistent launch in which a particle stays in its warp's queues from timestep to timestep
 * (no reload / write-back per timestep); the final state is bit-identical to n_steps single launches.
 * EMC_FUSED_STEPS=0 in the environment keeps the launch loop.
 * With obs_rows != NULL in the before-step observable phase (emc_set_obs_phase(ctx, 1)) and the same lean case,
 * the timesteps run as timestep chains: up to 1024 timesteps and their rows per persistent launch, the blocks
 * taking (slice, timestep) work items from a device-side queue -- a timestep of a slice waits only for the
 * previous timestep of THAT slice.  State and rows are bit-identical to n_steps single launches.  One chain
 * per context at a time (the queue lives in the context, like the observable partials: callers that drive
 * one context from several streams serialise these calls).  EMC_CHAIN_STEPS=0 keeps the launch loop. */
int emc_flight_scatter_steps(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x,
                             double *y, double *z, double *dtau, int32_t *valley, int64_t n,
                             int32_t field_mode, const double *ex, const double *ey,
                             const double *ez, uint64_t seed, uint64_t pid_offset,
                             uint32_t first_step, int32_t n_steps, EmcObs *obs_rows, void *stream);

/* Self-check of the fused launch above: EMC_ERR_CUDA if any warp ever stopped with particles left in
 * its queues (cannot happen by construction; synchronises the stream). */
int emc_flight_steps_check(EmcCtx *ctx, void *stream);

/* Replay mode: the same loop, but every np.random.uniform() of the reference is read from a
 * per-particle sequential stream: uniforms is a device (n, stride) row-major array, cursor a
 * device int32[n] read position that the call advances. */
int emc_flight_scatter_replay(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x,
                              double *y, double *z, double *dtau, int32_t *valley, int64_t n,
                              int32_t field_mode, const double *ex, const double *ey,
                              const double *ez, uint64_t pid_offset,
                              const double *uniforms, int64_t stride, int32_t *cursor,
                              EmcObs *obs, int64_t *mesh, void *stream);

/* main_.py:394-403 without stepping.  out is a HOST pointer; synchronises the stream. */
int emc_observables(EmcCtx *ctx, const double *kx, const double *ky, const double *kz,
                    const double *z, const int32_t *valley, int64_t n, EmcObs *out, void *stream);

/* The same reduction, only ENQUEUED: out_dev is a DEVICE pointer to one EmcObs, nothing is copied
 * to the host and the stream is not synchronised (pipelines that read the row later, e.g. the
 * chunked host-array time loop). */
int emc_observables_async(EmcCtx *ctx, const double *kx, const double *ky, const double *kz,
                          const double *z, const int32_t *valley, int64_t n, EmcObs *out_dev,
                          void *stream);

/* ---- single-function batches (the reference's per-particle signatures, n particles) -------- */
/* scatter_.calc_energy (scatter_.py:19-54): e_np[n], idx[n], fault[n] (-1 = ok) */
int emc_calc_energy(EmcCtx *ctx, const double *kx, const double *ky, const double *kz,
                    const int32_t *valley, int64_t n, double *e_np, int32_t *idx,
                    int32_t *fault, void *stream);
/* main_.free_flight (main_.py:114-123): out[i] = (-1/g0)*log(r[i]) */
int emc_free_flight(EmcCtx *ctx, double g0, const double *r, int64_t n, double *out, void *stream);
/* main_.carrier_drift (main_.py:179-213) with per-particle field (n x 3, V/cm), flight time
 * and effective mass (the reference passes mat.mass[valley]); state updated in place */
int emc_carrier_drift(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y,
                      double *z, const double *mass, const double *efield3, const double *dt2,
                      int64_t n, void *stream);
/* scatter_.scatter (scatter_.py:986-1099) with explicit uniforms u (n x 3: r2, azimuth,
 * polar).  k and valley updated in place unless the particle faults; mech[n], n_draws[n]
 * (2 for self-scattering, else 3) and fault[n] (-1 = ok) are outputs. */
int emc_scatter(EmcCtx *ctx, double *kx, double *ky, double *kz, int32_t *valley,
                const double *u3, int64_t n, int32_t *mech, int32_t *n_draws, int32_t *fault,
                void *stream);

/* Self-test of the guard-free sqrt / division / log the kernels use (csrc/emc_device.cuh): n
 * Philox-drawn operand sets are compared BIT FOR BIT with CUDA's correctly rounded sqrt(), `/`
 * and its log().  mismatches (HOST): counts for division, sqrt, log; all must be 0. */
int emc_selftest_math(EmcCtx *ctx, int64_t n, uint64_t seed, int64_t mismatches[3], void *stream);

/* Self-test of emc_mesh_to_rho's closed-form replay of the reference's sequential accumulation
 * (`rho[idx] += charge` once per particle, poisson_.py:36): n Philox-drawn (background, increment,
 * count <= max_count) triples, result compared BIT FOR BIT with the literal loop. *mismatches (HOST)
 * must be 0. */
int emc_selftest_accumulate(EmcCtx *ctx, int64_t n, uint64_t seed, int32_t max_count,
                            int64_t *mismatches, void *stream);

/* ---- layout helpers for callers that hold the reference's (N,6) array (main_.py:51-52) ----- */
int emc_aos_to_soa(EmcCtx *ctx, const double *coords6, int64_t n, double *kx, double *ky,
                   double *kz, double *x, double *y, double *z, void *stream);
int emc_soa_to_aos(EmcCtx *ctx, const double *kx, const double *ky, const double *kz,
                   const double *x, const double *y, const double *z, int64_t n,
                   double *coords6, void *stream);

/* ---- charge assignment: poisson_.get_near (poisson_.py:30-44; twin main_.py:242-261) ------- */
/* ADDS to mesh (device int64[nx*ny*nz], z fastest like rho.reshape(tot_seg), poisson_.py:42):
 * NGP adds 1 per hit cell, CIC adds llrint(w * 2^32).  Integer, hence bit-exact and
 * independent of the order particles are visited in (and of the number of ranks). */
int emc_deposit(EmcCtx *ctx, const double *x, const double *y, const double *z, int64_t n,
                int64_t *mesh, int32_t scheme, void *stream);
/* NGP on meshes too large for one block's shared memory (> 24576 cells): which kernel emc_deposit uses.
 *   1 direct: one 64-bit L2 reduction per particle straight into the mesh;
 *   2 tiled:  the cell ids are counting-sorted by mesh tile (4096 cells), every tile is accumulated in a
 *             shared-memory partial mesh and committed with one 64-bit add per touched cell;
 *   0 auto (default): the measured winner (profiles/r2_deposit_ab.txt).
 * Both are integer and give bit-identical meshes. */
int emc_set_deposit_mode(EmcCtx *ctx, int32_t mode);
/* Multi-GPU: sum the charge meshes of `world` ranks over NVLink peer memory in ONE kernel
 * (reduce-scatter + all-gather): this rank sums its 1/world slice over all ranks' input meshes
 * with peer loads and stores the totals into every rank's output mesh.  in_ptrs / out_ptrs:
 * HOST arrays of `world` peer-mapped device pointers (e.g. torch symmetric memory), each an
 * int64 mesh of n_cells rounded up to an even count, 16-byte aligned; index = rank.  The caller
 * orders the launch between two cross-rank barriers (inputs complete before, outputs read
 * after); the kernel itself never waits on another GPU.  Integer sums: bit-exact. */
#define EMC_MAX_PEERS 16
int emc_mesh_allreduce_p2p(EmcCtx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs,
                           int32_t world, int32_t rank, int64_t n_cells, void *stream);
/* The same exchange FUSED with rho (poisson_.py:27,36 per cell, as emc_mesh_to_rho): this rank sums its
 * slice of the counts, turns the totals of that slice into rho and stores totals AND rho into every rank's
 * buffers -- the sequential accumulation of poisson_.py:36 is then done once per cell in the whole job instead
 * of once per cell on every rank.  rho_ptrs: `world` peer-mapped padded (nx+2,ny+2,nz+2) fp64 arrays whose halo
 * is already zero; interior cells only are written.  Same barriers as emc_mesh_allreduce_p2p; rho is
 * bit-identical to emc_mesh_to_rho of the totals. */
int emc_mesh_allreduce_rho_p2p(EmcCtx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs,
                               double *const *rho_ptrs, int32_t world, int32_t rank, int64_t n_cells,
                               int32_t scheme, void *stream);

/* rho_pad (device, (nx+2,ny+2,nz+2), zero halo as np.pad at poisson_.py:53).  NGP: background
 * then `count` sequential += increment (poisson_.py:27,36: bit-identical to the reference's
 * accumulation).  CIC: background + increment * mesh/2^32. */
int emc_mesh_to_rho(EmcCtx *ctx, const int64_t *mesh, int32_t scheme, double *rho_pad,
                    void *stream);

/* The same, writing the interior cells only: for a rho_pad whose halo is already zero (any buffer a
 * previous emc_mesh_to_rho filled).  A time loop calls the full version once, then this one: a
 * third of the cells on an ny = 1 mesh. */
int emc_mesh_to_rho_interior(EmcCtx *ctx, const int64_t *mesh, int32_t scheme, double *rho_pad,
                             void *stream);

/* ---- Poisson: poisson_.py:47-84 (twin main_.py:283-315) ------------------------------------ */
/* u_pad: device (nx+2,ny+2,nz+2), holds the initial guess INCLUDING boundary values
 * (poisson_.py:47-52) and receives the solution.  Iterates until RMS(update) <= tol or
 * max_sweeps.  sweeps_out / err_out (HOST, may be NULL) receive the sweep count and the
 * last RMS update; err_hist (HOST, max_sweeps doubles, may be NULL) the whole history.
 * Synchronises the stream -- unless all three output pointers are NULL: then the solve is
 * only enqueued (the time loop does not need the counts) and emc_poisson_last reports later. */
int emc_poisson_sor(EmcCtx *ctx, double *u_pad, const double *rho_pad, double omega0, double tol,
                    int32_t max_sweeps, int32_t order, int32_t *sweeps_out, double *err_out,
                    double *err_hist, void *stream);
/* sweep count / last RMS update of the most recent emc_poisson_sor on this context, and divergence
 * (EMC_ERR_DIVERGED) of ANY solve since the previous emc_poisson_last (sticky device flag: solves that
 * were only enqueued are not lost); synchronises the stream */
int emc_poisson_last(EmcCtx *ctx, int32_t *sweeps_out, double *err_out, void *stream);
/* Geometric multigrid on the SAME discrete problem (poisson_.py:55-60,76: 7-point stencil, zero halo,
 * z-min face at its Dirichlet value as found in u_pad): cell-centred coarsening along the strongly
 * coupled axes (never an odd or < 4 cell axis, e.g. the reference's ny = 1), red-black Gauss-Seidel
 * V(nu1, nu2) cycles, residual averaging / linear prolongation, one cooperative launch for the whole
 * solve.  Stops when the RMS of the true residual  dl0^2 rho/eps - A u  (volts) <= tol_residual or after
 * max_cycles.  cycles_out / res_out / res_hist (HOST, may be NULL; res_hist: max_cycles doubles, RMS
 * residual after every cycle): if all three are NULL the solve is only enqueued (emc_poisson_last then
 * reports cycles and RMS residual).  EMC_ERR_UNSUPPORTED for a mesh that cannot be coarsened at all. */
int emc_poisson_mg(EmcCtx *ctx, double *u_pad, const double *rho_pad, double tol_residual,
                   int32_t max_cycles, int32_t nu1, int32_t nu2, int32_t *cycles_out, double *res_out,
                   double *res_hist, void *stream);
/* RMS and maximum of the TRUE residual  dl0^2 rho/eps - [neighbour sum - 2(1 + xy + xz) u]  (volts) of any
 * potential over the interior cells -- what the reference's stopping rule (RMS of the update) does not
 * bound.  rms_out / max_out: HOST (synchronises the stream if either is given); out_dev: DEVICE pointer
 * to two doubles (rms, max) or NULL, written in stream order. */
int emc_poisson_residual(EmcCtx *ctx, const double *u_pad, const double *rho_pad, double *rms_out,
                         double *max_out, double *out_dev, void *stream);
/* E = -central difference on interior cells (poisson_.py:88-96); padded outputs, V/m */
int emc_field(EmcCtx *ctx, const double *u_pad, double *ex, double *ey, double *ez, void *stream);
/* The same field as one record (ex, ey, ez, 0) of four doubles per padded cell, already in the
 * V/cm carrier_drift expects (main_.py:196: value * 1e-2, the identical multiplication the
 * gather of EMC_FIELD_MESH applies): two 128-bit loads per flight segment instead of three
 * scattered 64-bit ones.  e4: device, 4 * (nx+2)(ny+2)(nz+2) doubles, 32-byte aligned. */
int emc_field_packed(EmcCtx *ctx, const double *u_pad, double *e4, void *stream);
/* ny = 1 meshes: the field as (ex, ez) in V/cm per interior cell, e2: device, 2 * nx * nz doubles, 16-byte
 * aligned (EMC_FIELD_MESH_XZ).  Half the gathered bytes of emc_field_packed and no halo records. */
int emc_field_xz(EmcCtx *ctx, const double *u_pad, double *e2, void *stream);
/* interior records only (the halo records of e4 are already zero: any buffer a previous
 * emc_field_packed filled) */
int emc_field_packed_interior(EmcCtx *ctx, const double *u_pad, double *e4, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EMC_H */
