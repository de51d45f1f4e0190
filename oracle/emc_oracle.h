/*
 * emc_oracle.h -- CPU oracle for the ensemble-Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product package (monte_carlo_carrier_transport_b200/) never links, imports
 * or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's algorithm in the reference's
 * own operation order (file:line citations into /root/reference are on every
 * function in emc_oracle.c).  It is pinned against golden vectors produced by
 * executing the unmodified reference (tests/golden/generate_golden.py).
 */
#ifndef EMC_ORACLE_H
#define EMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_MECH 16

/* polar-angle sampler ids (scatter_.py:795-889) */
enum { ORC_SAMPLER_SELF = 0, ORC_SAMPLER_ISO = 1, ORC_SAMPLER_POP_ABS = 2,
       ORC_SAMPLER_POP_EMS = 3, ORC_SAMPLER_ION = 4 };

/* fault codes; a faulted particle gets valley = -(1+code) and is frozen */
enum { ORC_FAULT_NAN = 0, ORC_FAULT_ZERO_E = 1, ORC_FAULT_NO_MECH = 2,
       ORC_FAULT_DOMAIN = 3, ORC_FAULT_NEG_ENERGY = 4, ORC_FAULT_ZERO_K = 5,
       ORC_FAULT_STREAM = 6, ORC_FAULT_OUTSIDE = 7 /* where_am_i raised, init_.py:46,56 */ };

#define ORC_MAX_LAYERS 4

typedef struct OrcParams {
    double dt;            /* init_.py:60  */
    double gamma;         /* init_.py:83  GaAs.tot_scat */
    double hbar;          /* init_.py:68  */
    double q;             /* init_.py:67  */
    double mass[3];       /* init_.py:87  */
    double nonparab[3];   /* init_.py:89  */
    double tot_dim[3];    /* init_.py:147 */
    double efield[3];     /* init_.py:133, V/cm */
    double e_phonon;      /* init_.py:79  */
    double ion_eb;        /* scatter_.py:868-869: q/(2*eps_0*b), b=(1/(2*dope))**(1/3) */
    int32_t n_energy;     /* main_.py:45 */
    int32_t n_mech[3];    /* scatter_.py:909-946: 10/12/12 */
    const double *ek;     /* energy grid, n_energy */
    const double *cum[3]; /* row-major (n_energy, n_mech[v]) cumulative table */
    double  mech_dE[3][ORC_MAX_MECH];
    int32_t mech_trans[3][ORC_MAX_MECH];
    int32_t mech_sampler[3][ORC_MAX_MECH];
    /* field source of carrier_drift (extensions beyond the reference's constant dev.elec_field):
       0 constant `efield`; 1 per-group constant: particle pid uses field_groups[pid/group_size];
       2 nearest-cell gather from a padded (nx+2,ny+2,nz+2) mesh field in V/m (poisson_.py:88-96) */
    int32_t field_mode;
    int32_t n_groups;
    int64_t group_size;
    const double *field_groups;
    int32_t seg[3];
    int32_t z_cyclic;     /* extension: 1 = z wraps like x and y instead of the specular walls */
    double dl[3];
    const double *ex, *ey, *ez;
    double kT, mean_energy;   /* init_.py:64,145 (ensemble initialisation) */
    /* multi-layer devices (dev.layers, init_.py:127-130): n_layers > 1 makes the loop look the
       layer up with where_am_i (init_.py:36-56) before every drift / scatter / observable
       (main_.py:366,372,377,387,393) and read mass, nonparab, e_phonon, ion_eb, tables and
       mechanism metadata from layer_mat[layer]; everything else comes from this struct */
    int32_t n_layers;
    double layer_lz[ORC_MAX_LAYERS];
    const struct OrcParams *layer_mat[ORC_MAX_LAYERS];
    double layer_cw[ORC_MAX_LAYERS];   /* cumulative shares of the initial ensemble (extension); all 0 = uniform z */
} OrcParams;

/* init_.where_am_i (init_.py:36-56): layer index or -1 where the reference raises */
int    orc_where_am_i(const OrcParams *p, double dist);

typedef struct OrcObs {
    double sum_z, sum_vz, sum_e;        /* main_.py:394-400 */
    double sum_vz2, sum_e2;
    int64_t n_valley[3];                /* main_.py:401-403 */
    int64_t n_fault;
    int64_t n_segments;                 /* carrier_drift calls  (SURVEY 8d metric i) */
    int64_t n_events;                   /* scatter_.scatter calls */
} OrcObs;

/* random-stream source for one run */
typedef struct OrcStream {
    int32_t mode;              /* 0 = replay buffer, 1 = Philox4x32-10 */
    const double *uniforms;    /* replay: (n, stride) row-major */
    int64_t stride;
    int32_t *cursor;           /* replay: per-particle read position, updated */
    uint64_t seed;             /* philox key */
    uint64_t pid_offset;       /* global id of particle 0 (rank sharding) */
    uint32_t step;             /* philox counter word 1 */
} OrcStream;

void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_philox_uniform(uint64_t seed, uint64_t pid, uint32_t step, uint32_t draw);

int    orc_calc_energy(const OrcParams *p, const double k[3], int val, double *e_np, int32_t *idx);
double orc_free_flight(double g0, double r1);
void   orc_carrier_drift(const OrcParams *p, double coord[6], const double efield[3], double dt2, double mass);
/* one scatter event with explicit uniforms (r2, r_az, r_pol); returns fault code or -1;
   *n_draws receives how many uniforms were consumed (2 for self-scattering, else 3) */
int    orc_scatter(const OrcParams *p, double coord[6], int32_t *val, const double u[3],
                   int32_t *mech_out, int32_t *n_draws);

/* one synchronous-ensemble timestep over n particles (main_.py:360-397) */
void   orc_timestep(const OrcParams *p, int64_t n, double *kx, double *ky, double *kz,
                    double *x, double *y, double *z, double *dtau, int32_t *valley,
                    OrcStream *s, OrcObs *obs, int n_threads);

/* init_.init_coords (init_.py:194-234) + valley 0 (main_.py:53) + first free flights
   (main_.py:337) with the Philox draw layout documented in emc_oracle.c */
void   orc_init_ensemble(const OrcParams *p, int64_t n, double *kx, double *ky, double *kz,
                         double *x, double *y, double *z, double *dtau, int32_t *valley,
                         uint64_t seed, uint64_t pid_offset);

/* init_.init_photoex (init_.py:244-280) + main_.py:349-351; returns the number written */
int64_t orc_init_photoex(const OrcParams *p, int64_t n_cand, double energy, double z_scale,
                         double xy_std, double gamma_first, uint64_t seed, uint64_t cand_offset,
                         double *kx, double *ky, double *kz, double *x, double *y, double *z,
                         double *dtau, int32_t *valley);

/* observables only (main_.py:394-403) */
void   orc_observables(const OrcParams *p, int64_t n, const double *kx, const double *ky,
                       const double *kz, const double *z, const int32_t *valley, OrcObs *obs);

/* NGP charge assignment (poisson_.py:30-44): counts[c] += 1 for every cell whose centre is
   within dl/2 of the particle on all three axes (inclusive: face hits count twice) */
void   orc_deposit_ngp(int64_t n, const double *x, const double *y, const double *z,
                       const int32_t seg[3], const double dl[3], int64_t *counts);
/* CIC weights in 2^-32 fixed point (extension; commented line poisson_.py:34) */
void   orc_deposit_cic(int64_t n, const double *x, const double *y, const double *z,
                       const int32_t seg[3], const double dl[3], int64_t *mesh_fx);
/* rho[c] = background, then `counts[c]` sequential `+= increment` (poisson_.py:27,36) */
void   orc_counts_to_rho(int64_t n_cells, const int64_t *counts, double background,
                         double increment, double *rho);

/* CIC fixed point -> rho: background + increment * mesh/2^32 */
int64_t orc_div_const_mismatches(double b, int64_t n, uint64_t seed, int e_lo, int e_hi);
void   orc_fx_to_rho(int64_t n_cells, const int64_t *mesh_fx, double background,
                     double increment, double *rho);

/* lexicographic Gauss-Seidel/SOR (poisson_.py:55-84) on padded arrays (nx+2,ny+2,nz+2),
   C order (z fastest).  Returns sweeps done; err_hist gets RMS update per sweep. */
int    orc_poisson_sor(const int32_t seg[3], const double dl[3], double eps_s, double *u_pad,
                       const double *rho_pad, double omega0, double tol, int max_sweeps,
                       double *err_hist);
/* E = -central difference on interior cells (poisson_.py:88-96) */
void   orc_field(const int32_t seg[3], const double dl[3], const double *u_pad,
                 double *ex, double *ey, double *ez);

#ifdef __cplusplus
}
#endif
#endif
