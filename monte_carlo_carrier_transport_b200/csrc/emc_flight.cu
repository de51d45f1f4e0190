// emc_flight.cu -- the fused free-flight / drift / scatter kernel of one EMC timestep
// (main_.py:360-397), its observables (main_.py:394-403), the ensemble initialiser
// (init_.py:194-234) and the single-function batch kernels behind the reference's
// per-particle signatures.  sm_100a; compiled with -fmad=false (see emc_device.cuh).
#include "emc_host.h"

namespace emc {

// ------------------------------------------------------------------------------------------
// block-level reduction of the observables
// ------------------------------------------------------------------------------------------
struct ObsAcc {
    double d[5];       // sum_z, sum_vz, sum_e, sum_vz2, sum_e2
    long long c[6];    // n_valley[3], n_fault, n_segments, n_events
};

__device__ __forceinline__ void obs_particle(const DevConst &P, ObsAcc &a, double kx, double ky, double kz,
                                             double z, int val)
{
    if (val < 0) { a.c[3]++; return; }
    // main_.py:105,109,394-397
    const double E = energy_parabolic(P, kx, ky, kz, val);
    const double enp = energy_nonparabolic(P, E, val);
    const double vz = kz * P.hbar / P.mass[val];
    a.d[0] += z;
    a.d[1] += vz;
    a.d[2] += enp;
    a.d[3] += vz * vz;
    a.d[4] += enp * enp;
    a.c[val]++;
}

// Deterministic two-level reduction: warp shuffles, then shared memory, one partial per block;
// the last block to finish (ticket) folds the partials in a fixed order, so the result does
// not depend on block scheduling.
__device__ void obs_block_reduce_and_finish(ObsAcc &a, ObsPartial *partials, unsigned int *ticket, EmcObs *out)
{
    __shared__ ObsPartial warp_part[32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k) a.d[k] += __shfl_down_sync(0xffffffffu, a.d[k], off);
#pragma unroll
        for (int k = 0; k < 6; ++k) a.c[k] += __shfl_down_sync(0xffffffffu, a.c[k], off);
    }
    if (lane == 0) {
        for (int k = 0; k < 5; ++k) warp_part[warp].d[k] = a.d[k];
        for (int k = 0; k < 6; ++k) warp_part[warp].c[k] = a.c[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ObsPartial t = warp_part[0];
        for (int w = 1; w < nwarp; ++w) {
            for (int k = 0; k < 5; ++k) t.d[k] += warp_part[w].d[k];
            for (int k = 0; k < 6; ++k) t.c[k] += warp_part[w].c[k];
        }
        partials[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed-order fold: thread t owns value t (11 values), walks the blocks in index order
    if (threadIdx.x < 11) {
        const int k = threadIdx.x;
        if (k < 5) {
            double s = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[b].d[k]);
            (&out->sum_z)[k] = s;
        } else {
            long long s = 0;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[b].c[k - 5]);
            out->n_valley[0 + (k - 5)] = s;      // n_valley[3], n_fault, n_segments, n_events are contiguous
        }
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// ------------------------------------------------------------------------------------------
// NGP charge assignment of one particle (poisson_.py:30-38); `add(cell)` commits one count
// ------------------------------------------------------------------------------------------
template <class Add>
__device__ __forceinline__ void deposit_ngp_particle(const DevConst &P, double x, double y, double z, Add add)
{
    int hx[3], hy[3], hz[3];
    const int nx = axis_hits(x, P.seg[0], P.dl[0], hx);
    const int ny = axis_hits(y, P.seg[1], P.dl[1], hy);
    const int nz = axis_hits(z, P.seg[2], P.dl[2], hz);
    for (int a = 0; a < nx; ++a)
        for (int b = 0; b < ny; ++b)
            for (int c = 0; c < nz; ++c)
                add(((int64_t)hx[a] * P.seg[1] + hy[b]) * P.seg[2] + hz[c]);
}

// ------------------------------------------------------------------------------------------
// the timestep kernel
// ------------------------------------------------------------------------------------------
template <int FIELD>
__device__ __forceinline__ void field_at(const DevConst &P, const FlightArgs &A, int64_t i, double x, double y,
                                         double z, double &efx, double &efy, double &efz)
{
    if (FIELD == EMC_FIELD_CONSTANT) {
        efx = P.efield[0]; efy = P.efield[1]; efz = P.efield[2];
    } else if (FIELD == EMC_FIELD_GROUPS) {
        int64_t g = (int64_t)((A.pid_offset + (uint64_t)i) / (uint64_t)P.group_size);
        if (g >= P.n_groups) g = P.n_groups - 1;
        efx = __ldg(P.field_groups + 3 * g);
        efy = __ldg(P.field_groups + 3 * g + 1);
        efz = __ldg(P.field_groups + 3 * g + 2);
    } else {
        // nearest-cell gather from the padded (nx+2,ny+2,nz+2) field of emc_field, V/m -> V/cm
        const int ci = axis_cell(x, P.seg[0], P.dl[0]) + 1;
        const int cj = axis_cell(y, P.seg[1], P.dl[1]) + 1;
        const int ck = axis_cell(z, P.seg[2], P.dl[2]) + 1;
        const int64_t c = ((int64_t)ci * (P.seg[1] + 2) + cj) * (P.seg[2] + 2) + ck;
        efx = __ldg(A.ex + c) * 1e-2;
        efy = __ldg(A.ey + c) * 1e-2;
        efz = __ldg(A.ez + c) * 1e-2;
    }
}

// main_.py:360-391 for one particle.  Returns with the state updated; counts segments/events.
template <class Stream, int FIELD, bool EXACT>
__device__ __forceinline__ void step_particle(const DevConst &P, const MechMeta &M, const FlightArgs &A, int64_t i,
                                              double &kx, double &ky, double &kz, double &x, double &y, double &z,
                                              double &dtau, int &val, Stream &rng, long long &n_seg,
                                              long long &n_ev)
{
    if (val < 0) return;                          // frozen after a fault
    const double dt = P.dt;
    double dte = dtau;                            // :361
    double efx, efy, efz;
    field_at<FIELD>(P, A, i, x, y, z, efx, efy, efz);
    if (dte >= dt) {                              // :363
        carrier_drift(P, kx, ky, kz, x, y, z, efx, efy, efz, dt, P.mass[val]);        // :367
        n_seg++;
    } else {
        carrier_drift(P, kx, ky, kz, x, y, z, efx, efy, efz, dte, P.mass[val]);       // :373
        n_seg++;
        while (dte < dt) {                        // :375
            const double dte2 = dte;              // :376
            int f = scatter_event<EXACT>(P, M, kx, ky, kz, val, rng);                 // :378
            if (rng.overflow()) f = EMC_FAULT_STREAM;
            if (f >= 0) { val = -(1 + f); dtau = dte; return; }
            n_ev++;
            const double dt3 = P.neg_inv_gamma * log(rng.next());                    // :379
            if (rng.overflow()) { val = -(1 + EMC_FAULT_STREAM); dtau = dte; return; }
            const double dtp = dt - dte2;         // :381
            const double dt2 = (dt3 <= dtp) ? dt3 : dtp;                              // :382-385
            if (FIELD == EMC_FIELD_MESH) field_at<FIELD>(P, A, i, x, y, z, efx, efy, efz);
            carrier_drift(P, kx, ky, kz, x, y, z, efx, efy, efz, dt2, P.mass[val]);  // :388
            n_seg++;
            dte = dte2 + dt3;                     // :389
        }
    }
    dte -= dt;                                    // :390
    dtau = dte;                                   // :391
}

template <class Stream, int FIELD, bool EXACT>
__global__ void __launch_bounds__(FLIGHT_THREADS)
flight_scatter_kernel(const __grid_constant__ DevConst P, const __grid_constant__ MechMeta Mg,
                      const __grid_constant__ FlightArgs A)
{
    __shared__ MechMeta M;
    extern __shared__ unsigned int smesh[];       // privatised NGP counts (when A.smem_cells > 0)
    {
        const int nw = sizeof(MechMeta) / 4;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&Mg);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&M);
        for (int t = threadIdx.x; t < nw; t += blockDim.x) dst[t] = src[t];
        for (int t = threadIdx.x; t < A.smem_cells; t += blockDim.x) smesh[t] = 0u;
    }
    __syncthreads();

    ObsAcc acc;
#pragma unroll
    for (int k = 0; k < 5; ++k) acc.d[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc.c[k] = 0;

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += stride) {
        double kx = A.kx[i], ky = A.ky[i], kz = A.kz[i];
        double x = A.x[i], y = A.y[i], z = A.z[i];
        double dtau = A.dtau[i];
        int val = A.valley[i];
        Stream rng;
        if constexpr (Stream::is_replay) rng.init(A.uniforms, A.stride, i, A.cursor[i]);
        else rng.init(A.seed, A.pid_offset + (uint64_t)i, A.step);

        step_particle<Stream, FIELD, EXACT>(P, M, A, i, kx, ky, kz, x, y, z, dtau, val, rng, acc.c[4], acc.c[5]);

        A.kx[i] = kx; A.ky[i] = ky; A.kz[i] = kz;
        A.x[i] = x; A.y[i] = y; A.z[i] = z;
        A.dtau[i] = dtau;
        A.valley[i] = val;
        if constexpr (Stream::is_replay) A.cursor[i] = rng.pos;

        if (A.obs) obs_particle(P, acc, kx, ky, kz, z, val);
        if (A.mesh) {
            if (A.smem_cells > 0)
                deposit_ngp_particle(P, x, y, z, [&](int64_t c) { atomicAdd(&smesh[c], 1u); });
            else
                deposit_ngp_particle(P, x, y, z, [&](int64_t c) {
                    atomicAdd(reinterpret_cast<unsigned long long *>(A.mesh) + c, 1ull); });
        }
    }
    if (A.mesh && A.smem_cells > 0) {
        __syncthreads();
        for (int t = threadIdx.x; t < A.smem_cells; t += blockDim.x) {
            const unsigned int v = smesh[t];
            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(A.mesh) + t, (unsigned long long)v);
        }
    }
    if (A.obs) obs_block_reduce_and_finish(acc, A.partials, A.ticket, A.obs);
}

template <class Stream, int FIELD>
static cudaError_t launch_flight_t(const Ctx *ctx, const FlightArgs &A, int grid, size_t smem, cudaStream_t st)
{
    if (ctx->exact_libm)
        flight_scatter_kernel<Stream, FIELD, true><<<grid, FLIGHT_THREADS, smem, st>>>(ctx->dc, ctx->mech, A);
    else
        flight_scatter_kernel<Stream, FIELD, false><<<grid, FLIGHT_THREADS, smem, st>>>(ctx->dc, ctx->mech, A);
    return cudaGetLastError();
}

cudaError_t launch_flight(Ctx *ctx, FlightArgs &A, bool replay, int field_mode, cudaStream_t st)
{
    const int64_t blocks_needed = (A.n + FLIGHT_THREADS - 1) / FLIGHT_THREADS;
    int grid = (int)(blocks_needed < (int64_t)ctx->max_grid ? blocks_needed : ctx->max_grid);
    if (grid < 1) grid = 1;
    size_t smem = 0;
    A.smem_cells = 0;
    if (A.mesh) {
        const int64_t cells = (int64_t)ctx->dc.seg[0] * ctx->dc.seg[1] * ctx->dc.seg[2];
        if (cells * 4 <= SMEM_MESH_BYTES) { A.smem_cells = (int)cells; smem = (size_t)cells * 4; }
    }
    A.partials = ctx->partials;
    A.ticket = ctx->ticket;
    ctx->launches++;
#define EMC_DISPATCH(STREAM)                                                                     \
    switch (field_mode) {                                                                        \
    case EMC_FIELD_CONSTANT: return launch_flight_t<STREAM, EMC_FIELD_CONSTANT>(ctx, A, grid, smem, st); \
    case EMC_FIELD_GROUPS:   return launch_flight_t<STREAM, EMC_FIELD_GROUPS>(ctx, A, grid, smem, st);   \
    case EMC_FIELD_MESH:     return launch_flight_t<STREAM, EMC_FIELD_MESH>(ctx, A, grid, smem, st);     \
    default: return cudaErrorInvalidValue;                                                       \
    }
    if (replay) { EMC_DISPATCH(ReplayStream) } else { EMC_DISPATCH(PhiloxStream) }
#undef EMC_DISPATCH
}

// ------------------------------------------------------------------------------------------
// observables only (main_.py:394-403)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FLIGHT_THREADS)
observables_kernel(const __grid_constant__ DevConst P, const double *kx, const double *ky, const double *kz,
                   const double *z, const int32_t *valley, int64_t n, ObsPartial *partials,
                   unsigned int *ticket, EmcObs *out)
{
    ObsAcc acc;
    for (int k = 0; k < 5; ++k) acc.d[k] = 0.0;
    for (int k = 0; k < 6; ++k) acc.c[k] = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        obs_particle(P, acc, kx[i], ky[i], kz[i], z[i], valley[i]);
    obs_block_reduce_and_finish(acc, partials, ticket, out);
}

cudaError_t launch_observables(Ctx *ctx, const double *kx, const double *ky, const double *kz, const double *z,
                               const int32_t *valley, int64_t n, EmcObs *out_dev, cudaStream_t st)
{
    const int64_t blocks_needed = (n + FLIGHT_THREADS - 1) / FLIGHT_THREADS;
    int grid = (int)(blocks_needed < (int64_t)ctx->max_grid ? blocks_needed : ctx->max_grid);
    if (grid < 1) grid = 1;
    ctx->launches++;
    observables_kernel<<<grid, FLIGHT_THREADS, 0, st>>>(ctx->dc, kx, ky, kz, z, valley, n, ctx->partials,
                                                        ctx->ticket, out_dev);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// ensemble initialisation: init_.init_coords (init_.py:194-234), valley = 0 (main_.py:53),
// first free flight (main_.py:337).  Draw order per particle (Philox step word 0xFFFFFFFF):
//   0,1,2  x, y, z = u * box                                  (init_.py:198-200)
//   3..6   three standard normals (Box-Muller pairs (3,4), (5,6); the 4th is discarded);
//          E = loc + scale * sqrt(n1^2+n2^2+n3^2), loc = mean*kT, scale = 1e-7
//          (scipy maxwell.rvs == chi(3), init_.py:207)
//   7,8    alpha, beta = 2*pi * (Box-Muller pair)             (init_.py:216-217: Normal(0, 2*pi))
//   9      dtau = (-1/G0) * log(u)                            (main_.py:337)
// Box-Muller: rad = sqrt(-2 log(1-u_a)), n = rad*cos(2 pi u_b), rad*sin(2 pi u_b).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void box_muller(double ua, double ub, double &n0, double &n1)
{
    const double rad = sqrt(-2 * log(1 - ua));
    double s, c;
    sincos(2 * 3.14159265358979323846 * ub, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

__global__ void __launch_bounds__(256)
init_ensemble_kernel(const __grid_constant__ DevConst P, double kT, double mean_energy, double *kx, double *ky,
                     double *kz, double *x, double *y, double *z, double *dtau, int32_t *valley, int64_t n,
                     uint64_t seed, uint64_t pid_offset)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        PhiloxStream rng;
        rng.init(seed, pid_offset + (uint64_t)i, 0xFFFFFFFFu);
        const double px = P.tot_dim[0] * rng.next();
        const double py = P.tot_dim[1] * rng.next();
        const double pz = P.tot_dim[2] * rng.next();
        double n1, n2, n3, n4;
        { const double a = rng.next(), b = rng.next(); box_muller(a, b, n1, n2); }
        { const double a = rng.next(), b = rng.next(); box_muller(a, b, n3, n4); }
        const double e = mean_energy * kT + 1E-7 * sqrt(n1 * n1 + n2 * n2 + n3 * n3);
        const double k = sqrt(2 * P.mass[0] * P.q * e / (P.hbar * P.hbar));      // init_.py:215
        double na, nb;
        { const double a = rng.next(), b = rng.next(); box_muller(a, b, na, nb); }
        const double alpha = 2 * 3.14159265358979323846 * na, beta = 2 * 3.14159265358979323846 * nb;
        double sa, ca, sb, cb;
        sincos(alpha, &sa, &ca);
        sincos(beta, &sb, &cb);
        kx[i] = k * ca * sb;                       // init_.py:218-220
        ky[i] = k * sa * sb;
        kz[i] = k * cb;
        x[i] = px; y[i] = py; z[i] = pz;
        dtau[i] = P.neg_inv_gamma * log(rng.next());
        valley[i] = 0;
    }
}

cudaError_t launch_init_ensemble(Ctx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                 double *dtau, int32_t *valley, int64_t n, uint64_t seed, uint64_t pid_offset,
                                 cudaStream_t st)
{
    const int64_t blocks_needed = (n + 255) / 256;
    int grid = (int)(blocks_needed < (int64_t)ctx->max_grid ? blocks_needed : ctx->max_grid);
    if (grid < 1) grid = 1;
    ctx->launches++;
    init_ensemble_kernel<<<grid, 256, 0, st>>>(ctx->dc, ctx->params.kT, ctx->params.mean_energy, kx, ky, kz, x, y,
                                               z, dtau, valley, n, seed, pid_offset);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// single-function batches
// ------------------------------------------------------------------------------------------
__global__ void calc_energy_kernel(const __grid_constant__ DevConst P, const double *kx, const double *ky,
                                   const double *kz, const int32_t *valley, int64_t n, double *e_np,
                                   int32_t *idx, int32_t *fault)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e = 0.0;
    int32_t j = 0;
    const int f = calc_energy(P, kx[i], ky[i], kz[i], valley[i], e, j);
    e_np[i] = e;
    idx[i] = j;
    fault[i] = f;
}

__global__ void free_flight_kernel(double g0, const double *r, int64_t n, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (-1 / g0) * log(r[i]);              // main_.py:122
}

__global__ void carrier_drift_kernel(const __grid_constant__ DevConst P, double *kx, double *ky, double *kz,
                                     double *x, double *y, double *z, const double *mass,
                                     const double *ef3, const double *dt2, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = kx[i], b = ky[i], c = kz[i], px = x[i], py = y[i], pz = z[i];
    carrier_drift(P, a, b, c, px, py, pz, ef3[3 * i], ef3[3 * i + 1], ef3[3 * i + 2], dt2[i], mass[i]);
    kx[i] = a; ky[i] = b; kz[i] = c;
    x[i] = px; y[i] = py; z[i] = pz;
}

template <bool EXACT>
__global__ void scatter_kernel(const __grid_constant__ DevConst P, const __grid_constant__ MechMeta Mg, double *kx,
                               double *ky, double *kz, int32_t *valley, const double *u3, int64_t n,
                               int32_t *mech_out, int32_t *n_draws, int32_t *fault)
{
    __shared__ MechMeta M;
    {
        const int nw = sizeof(MechMeta) / 4;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&Mg);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&M);
        for (int t = threadIdx.x; t < nw; t += blockDim.x) dst[t] = src[t];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = kx[i], b = ky[i], c = kz[i];
    int val = valley[i];
    ReplayStream rng;
    rng.init(u3, 3, i, 0);
    int mech = -1;
    const int f = scatter_event<EXACT>(P, M, a, b, c, val, rng, &mech);
    kx[i] = a; ky[i] = b; kz[i] = c;
    valley[i] = val;
    mech_out[i] = mech;
    n_draws[i] = rng.pos;
    fault[i] = f;
}

cudaError_t launch_calc_energy(Ctx *ctx, const double *kx, const double *ky, const double *kz,
                               const int32_t *valley, int64_t n, double *e_np, int32_t *idx, int32_t *fault,
                               cudaStream_t st)
{
    ctx->launches++;
    calc_energy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->dc, kx, ky, kz, valley, n, e_np, idx, fault);
    return cudaGetLastError();
}

cudaError_t launch_free_flight(Ctx *ctx, double g0, const double *r, int64_t n, double *out, cudaStream_t st)
{
    ctx->launches++;
    free_flight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g0, r, n, out);
    return cudaGetLastError();
}

cudaError_t launch_carrier_drift(Ctx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                 const double *mass, const double *ef3, const double *dt2, int64_t n,
                                 cudaStream_t st)
{
    ctx->launches++;
    carrier_drift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->dc, kx, ky, kz, x, y, z, mass, ef3,
                                                                     dt2, n);
    return cudaGetLastError();
}

cudaError_t launch_scatter(Ctx *ctx, double *kx, double *ky, double *kz, int32_t *valley, const double *u3,
                           int64_t n, int32_t *mech, int32_t *n_draws, int32_t *fault, cudaStream_t st)
{
    ctx->launches++;
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (ctx->exact_libm)
        scatter_kernel<true><<<grid, 128, 0, st>>>(ctx->dc, ctx->mech, kx, ky, kz, valley, u3, n, mech, n_draws, fault);
    else
        scatter_kernel<false><<<grid, 128, 0, st>>>(ctx->dc, ctx->mech, kx, ky, kz, valley, u3, n, mech, n_draws, fault);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// (N,6) array-of-structures <-> structure-of-arrays (main_.py:51-52 layout)
// ------------------------------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const double2 *__restrict__ c6, int64_t n, double *kx, double *ky, double *kz,
                                  double *x, double *y, double *z)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 a = c6[3 * i], b = c6[3 * i + 1], c = c6[3 * i + 2];
        kx[i] = a.x; ky[i] = a.y; kz[i] = b.x;
        x[i] = b.y; y[i] = c.x; z[i] = c.y;
    }
}

__global__ void soa_to_aos_kernel(const double *__restrict__ kx, const double *__restrict__ ky,
                                  const double *__restrict__ kz, const double *__restrict__ x,
                                  const double *__restrict__ y, const double *__restrict__ z, int64_t n,
                                  double2 *c6)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        c6[3 * i] = make_double2(kx[i], ky[i]);
        c6[3 * i + 1] = make_double2(kz[i], x[i]);
        c6[3 * i + 2] = make_double2(y[i], z[i]);
    }
}

cudaError_t launch_aos_to_soa(Ctx *ctx, const double *c6, int64_t n, double *kx, double *ky, double *kz, double *x,
                              double *y, double *z, cudaStream_t st)
{
    const int64_t blocks_needed = (n + 255) / 256;
    int grid = (int)(blocks_needed < (int64_t)ctx->max_grid * 4 ? blocks_needed : (int64_t)ctx->max_grid * 4);
    if (grid < 1) grid = 1;
    ctx->launches++;
    aos_to_soa_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const double2 *>(c6), n, kx, ky, kz, x, y, z);
    return cudaGetLastError();
}

cudaError_t launch_soa_to_aos(Ctx *ctx, const double *kx, const double *ky, const double *kz, const double *x,
                              const double *y, const double *z, int64_t n, double *c6, cudaStream_t st)
{
    const int64_t blocks_needed = (n + 255) / 256;
    int grid = (int)(blocks_needed < (int64_t)ctx->max_grid * 4 ? blocks_needed : (int64_t)ctx->max_grid * 4);
    if (grid < 1) grid = 1;
    ctx->launches++;
    soa_to_aos_kernel<<<grid, 256, 0, st>>>(kx, ky, kz, x, y, z, n, reinterpret_cast<double2 *>(c6));
    return cudaGetLastError();
}

}  // namespace emc
