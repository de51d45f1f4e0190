// emc_host.h -- host-side context and launcher prototypes shared by the .cu files.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string>
#include "emc_device.cuh"

namespace emc {

#ifndef EMC_FLIGHT_THREADS
#define EMC_FLIGHT_THREADS 512       // tuning knob (-DEMC_FLIGHT_THREADS=...), multiple of 32
#endif
constexpr int FLIGHT_THREADS = EMC_FLIGHT_THREADS;
constexpr int SMEM_MESH_BYTES = 96 * 1024;     // NGP counts privatised in shared memory up to this size

struct ObsPartial {
    double d[5];
    long long c[6];
};

// Control block of a timestep chain (flight_scatter_kernel<..., CHAIN = true>, launch_flight_chain): K timesteps
// WITH their observable rows in one persistent launch.  The work items are (slice, timestep) pairs in a FIFO in
// device memory: a block pops one, runs the one-timestep body on that slice, and pushes (slice, timestep + 1) when
// it has written the slice back -- no grid-wide barrier between timesteps (slices are independent: a timestep
// only needs ITS slice of the previous one), so a slow SM never holds the others up and there is no launch gap.
// ready[i % CHAIN_RING]: (i + 1) << 32 | timestep << 16 | slice, written with release semantics into an EMPTY slot
// (0; the reader empties it again); the first `slices` items (timestep 0 of every slice) are implicit.  Zeroed by a
// memset before every chain launch.  The protocol under random schedules: tests/test_chain_protocol_model.py.
constexpr int CHAIN_RING = 4096;          // >> the number of items in flight (<= slices <= max_grid)
constexpr int CHAIN_MAX_STEPS = 1024;     // timesteps per chain launch (longer loops are split)
struct ChainCtl {
    unsigned int popped, pushed;
    unsigned int pad[30];
    unsigned long long ready[CHAIN_RING];
    unsigned int step_done[CHAIN_MAX_STEPS];      // observable partials of a timestep written so far
};

// kernel arguments of the timestep kernel (device pointers are the caller's)
struct FlightArgs {
    double *kx, *ky, *kz, *x, *y, *z, *dtau;
    int32_t *valley;
    int64_t n;
    // random stream
    const double *uniforms;      // replay
    int64_t stride;
    int32_t *cursor;
    uint64_t seed, pid_offset;   // philox
    uint32_t step;
    // EMC_FIELD_MESH
    const double *ex, *ey, *ez;
    // fused epilogues
    EmcObs *obs;
    ObsPartial *partials;
    unsigned int *ticket;
    int64_t *mesh;
    int32_t smem_cells;
    int32_t field_ahead;         // EMC_FIELD_MESH_PACKED: stage the next refill batch's field records (launch_flight decides)
    int32_t obs_before;          // emc_set_obs_phase: reduce the observables of the state as loaded
    int32_t bulk;                // refill batches arrive by bulk async copies (launch_flight decides: aligned rows)
    const MechMeta *mech_layers; // multi-layer: device array of n_layers mechanism-metadata blocks
    // timestep chain (CHAIN kernels): obs = the first of chain_steps rows, partials = [chain_steps][chain_slices]
    ChainCtl *chain;
    int32_t chain_steps, chain_slices;
};

struct Ctx {
    EmcParams params;
    DevConst dc;
    MechMeta mech;                // layer 0 (== mech_layers[0]); kernel parameter of single-layer launches
    MechMeta mech_layers[EMC_MAX_LAYERS];
    MechMeta *d_mech_layers = nullptr;
    EmcLayer layers[EMC_MAX_LAYERS];
    bool layer_tables[EMC_MAX_LAYERS] = {false, false, false, false};
    int obs_before = 0;           // emc_set_obs_phase
    bool no_field_ahead = false;  // EMC_FIELD_AHEAD=0 in the environment (tuning)
    bool no_bulk_refill = false;  // EMC_BULK_REFILL=0 in the environment (tuning)
    bool bulk_mesh = false;       // EMC_BULK_REFILL=2: bulk refill tiles in the mesh-field modes too (A/B)
    bool no_fused_steps = false;  // EMC_FUSED_STEPS=0: emc_flight_scatter_steps always launches once per timestep
    bool no_lean = false;         // EMC_LEAN=0 in the environment: always the general flight kernel (A/B, tests)
    bool self_last = false;       // every uploaded table: last column = the one self-scattering mechanism, dE 0, same valley
    bool zonly_constant = false;  // dc.efield has Ex = Ey = +0
    bool zonly_groups = false;    // every field group has Ex = Ey = +0
    int batch_layer = 0;          // emc_select_layer: material of the single-function batches
    double *d_cum_layers[EMC_MAX_LAYERS][3] = {};
    double *d_thr_layers[EMC_MAX_LAYERS][3] = {};   // self-scattering thresholds (Mat::thr)
    double *d_background_z = nullptr;   // emc_set_background_profile
    int background_nz = 0;
    int device = 0;
    int sm_count = 0;
    int max_grid = 0;             // upper bound of any persistent grid (partials capacity)
    int waves = 1;                // flight kernel: grid = resident blocks * waves (EMC_WAVES)
    bool have_tables = false;
    bool exact_libm = false;      // EMC_EXACT_LIBM=1: compose acos/asin/atan exactly like the reference
    double *d_groups = nullptr;
    ObsPartial *partials = nullptr;
    unsigned int *ticket = nullptr;
    EmcObs *d_obs = nullptr;      // scratch for emc_observables
    // Poisson scratch
    double *d_sor_scalars = nullptr;   // [0] err, [1] omega, ...
    double *d_err_hist = nullptr;
    int err_hist_cap = 0;
    double *d_sor_partials = nullptr;
    int32_t *d_sor_flags = nullptr;
    void *mg = nullptr;           // multigrid buffers (emc_multigrid.cu), created on first use
    int deposit_mode = 0;         // emc_set_deposit_mode: 0 auto, 1 direct atomics, 2 tiled (large NGP meshes)
    void *d_dep_scratch = nullptr;   // tiled deposit: cell ids, sorted cell ids, tile histogram
    size_t dep_scratch_bytes = 0;
    // the state rows as one 2-D tensor for the refill tiles (launch_flight; cached per base / length / pitch)
    alignas(64) CUtensorMap state_tmap = {};
    const char *tmap_base = nullptr;
    int64_t tmap_n = 0;
    ptrdiff_t tmap_pitch = 0;
    bool no_tensor_map = false;   // EMC_REFILL_TENSOR_MAP=0 in the environment (A/B, tests)
    // timestep chain: control block and per-timestep observable partials (grown on demand)
    ChainCtl *d_chain = nullptr;
    ObsPartial *chain_partials = nullptr;
    size_t chain_partials_cap = 0;    // in ObsPartial units
    bool no_chain = false;        // EMC_CHAIN_STEPS=0 in the environment: one launch per timestep (A/B, tests)
    int64_t launches = 0;
    std::string last_error;
};

// emc_flight.cu
cudaError_t launch_flight(Ctx *ctx, FlightArgs &A, bool replay, int field_mode, cudaStream_t st);
cudaError_t launch_flight_steps(Ctx *ctx, FlightArgs &A, int field_mode, int n_steps, cudaStream_t st);
// n_steps timesteps with their observable rows (A.obs[0 .. n_steps)) in one persistent launch; cudaErrorNotSupported
// = the chain conditions do not hold for this call (the caller launches once per timestep instead)
cudaError_t launch_flight_chain(Ctx *ctx, FlightArgs &A, int field_mode, int n_steps, cudaStream_t st);
cudaError_t launch_observables(Ctx *ctx, const double *kx, const double *ky, const double *kz, const double *z,
                               const int32_t *valley, int64_t n, EmcObs *out_dev, cudaStream_t st);
cudaError_t launch_init_ensemble(Ctx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                 double *dtau, int32_t *valley, int64_t n, uint64_t seed, uint64_t pid_offset,
                                 cudaStream_t st);
cudaError_t launch_init_photoex(Ctx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                double *dtau, int32_t *valley, int64_t n_cand, double energy, double z_scale,
                                double xy_std, double gamma_first, uint64_t seed, uint64_t cand_offset,
                                int64_t *n_accepted_host, cudaStream_t st);
cudaError_t launch_calc_energy(Ctx *ctx, const double *kx, const double *ky, const double *kz,
                               const int32_t *valley, int64_t n, double *e_np, int32_t *idx, int32_t *fault,
                               cudaStream_t st);
cudaError_t launch_free_flight(Ctx *ctx, double g0, const double *r, int64_t n, double *out, cudaStream_t st);
cudaError_t launch_carrier_drift(Ctx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                 const double *mass, const double *ef3, const double *dt2, int64_t n,
                                 cudaStream_t st);
cudaError_t launch_scatter(Ctx *ctx, double *kx, double *ky, double *kz, int32_t *valley, const double *u3,
                           int64_t n, int32_t *mech, int32_t *n_draws, int32_t *fault, cudaStream_t st);
cudaError_t launch_aos_to_soa(Ctx *ctx, const double *c6, int64_t n, double *kx, double *ky, double *kz, double *x,
                              double *y, double *z, cudaStream_t st);
cudaError_t launch_soa_to_aos(Ctx *ctx, const double *kx, const double *ky, const double *kz, const double *x,
                              const double *y, const double *z, int64_t n, double *c6, cudaStream_t st);

cudaError_t launch_selftest_math(Ctx *ctx, int64_t n, uint64_t seed, unsigned long long *counts_dev, cudaStream_t st);

// emc_mesh.cu
cudaError_t launch_deposit(Ctx *ctx, const double *x, const double *y, const double *z, int64_t n, int64_t *mesh,
                           int scheme, cudaStream_t st);
cudaError_t launch_selftest_accumulate(Ctx *ctx, int64_t n, uint64_t seed, int max_count, unsigned long long *bad_dev,
                                       cudaStream_t st);
cudaError_t launch_mesh_to_rho(Ctx *ctx, const int64_t *mesh, int scheme, double *rho_pad, cudaStream_t st,
                               bool interior_only = false);
cudaError_t launch_field(Ctx *ctx, const double *u_pad, double *ex, double *ey, double *ez, cudaStream_t st);
// returns 0 / EMC_ERR_*; fills sweeps/err; synchronises
int run_poisson_sor(Ctx *ctx, double *u_pad, const double *rho_pad, double omega0, double tol, int max_sweeps,
                    int order, int32_t *sweeps_out, double *err_out, double *err_hist, cudaStream_t st);

cudaError_t launch_mesh_allreduce_p2p(Ctx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs, int world,
                                      int rank, int64_t n_cells, cudaStream_t st);
cudaError_t launch_mesh_allreduce_rho_p2p(Ctx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs,
                                          double *const *rho_ptrs, int world, int rank, int64_t n_cells, int scheme,
                                          cudaStream_t st);
cudaError_t launch_field_packed(Ctx *ctx, const double *u_pad, double *e4, cudaStream_t st, bool interior_only = false);
cudaError_t launch_field_xz(Ctx *ctx, const double *u_pad, double *e2, cudaStream_t st);
int poisson_last(Ctx *ctx, int32_t *sweeps_out, double *err_out, cudaStream_t st);

// emc_multigrid.cu
int run_poisson_mg(Ctx *ctx, double *u_pad, const double *rho_pad, double tol_residual, int max_cycles, int nu1, int nu2,
                   int32_t *cycles_out, double *res_out, double *res_hist, cudaStream_t st);
int run_poisson_residual(Ctx *ctx, const double *u_pad, const double *rho_pad, double *rms_out, double *max_out,
                         double *out_dev, cudaStream_t st);
void mg_release(Ctx *ctx);

}  // namespace emc
