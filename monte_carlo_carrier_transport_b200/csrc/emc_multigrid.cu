// emc_multigrid.cu -- geometric multigrid for the Poisson problem of poisson_.py:47-84 (twin
// main_.py:283-315): the SAME discrete equation the reference relaxes,
//     u(i+1) + u(i-1) + xy (u(j+1) + u(j-1)) + xz (u(k+1) + u(k-1)) - 2 (1 + xy + xz) u = dl0^2 rho / eps
// (poisson_.py:55-60,76), the same Dirichlet data (zero halo, z-min face = surf_val, :47-52), hence the
// same fixed point as the reference's SOR -- reached in a handful of V-cycles where SOR's stopping
// rule (RMS of the UPDATE <= tol, :82) fires ~1e6 sweeps early on a 1024 x 1 x 1024 mesh.
//
// Components (all hand-written, one cooperative launch for the whole solve):
//   * cell-centred coarsening by 2 per axis, only along the strongly coupled axes of a level
//     (semi-coarsening: c_a >= 0.3 max c, c_a = coefficient of axis a; the z spacing of the
//     BASELINE mesh is 3.3x finer than x, so the first coarsening is z-only), never along an axis
//     with an odd cell count or fewer than 4 cells (the reference's ny = 1);
//   * red-black Gauss-Seidel smoother, V(nu1, nu2);
//   * restriction of the residual = average over the children (the cell-centred full weighting),
//     prolongation = (bi/tri)linear, weights 3/4 - 1/4 per coarsened axis;
//   * coarse operators by rediscretisation; the Dirichlet node of the reference sits one FINE
//     spacing outside the first cell centre, i.e. at d = 1/2 + 2^-(m+1) coarse spacings after m
//     coarsenings: the ghost value that makes the linear interpolant vanish there,
//     u_ghost = -(1-d)/d u_first, is folded into the diagonal of the cells next to the boundary;
//   * levels of <= MG_BLOCK_CELLS cells are solved by ONE block in shared memory between two grid
//     barriers (a grid barrier costs ~2 us, a V-cycle has ~10 per grid level);
//   * stopping rule: RMS of the TRUE residual (in the equation's own units, volts) <= tol_residual,
//     reduced in a fixed order (deterministic).
// Measured on B200 (profiles/r2_multigrid.txt): convergence factor ~0.05 per V(2,2) cycle.
#include <cooperative_groups.h>
#include <cmath>
#include <cstdio>
#include <vector>
#include "emc_host.h"

namespace cg = cooperative_groups;

namespace emc {

constexpr int MG_MAX_LEVELS = 24;
constexpr int MG_BLOCK_CELLS = 1024;       // levels at most this big live in one block's shared memory
constexpr int MG_THREADS = 256;

struct MgLevel {
    int n[3];                // cells per axis
    int co[3];               // 1: this axis is halved to get the next (coarser) level
    double c[3];             // coefficients of the axis neighbours (level 0: 1, xy, xz)
    double g[3];             // ghost factor of the axis: u_ghost = g * u_first (0 on uncoarsened axes)
    double *u, *f;           // padded (n+2)^3 arrays; level 0: u = the caller's u_pad, f unused (from rho)
    long long pad_cells;
};

struct MgPlan {
    int n_levels;
    int first_block_level;   // levels >= this are solved by block 0 in shared memory (n_levels if none)
    MgLevel lev[MG_MAX_LEVELS];
    const double *rho_pad;   // level 0 right-hand side: f = rhs_scale * rho
    double rhs_scale;        // dl0^2 / eps_s
    double tol_residual;
    int max_cycles, nu1, nu2, coarse_sweeps;
};

__device__ __forceinline__ long long mg_idx(const MgLevel &L, int i, int j, int k)
{
    return ((long long)i * (L.n[1] + 2) + j) * (L.n[2] + 2) + k;
}

// diagonal of cell (i, j, k) (1-based): -2 sum c + ghost terms next to the boundary
__device__ __forceinline__ double mg_diag(const MgLevel &L, int i, int j, int k)
{
    double d = -2.0 * (L.c[0] + L.c[1] + L.c[2]);
    d += L.c[0] * L.g[0] * (double)((i == 1) + (i == L.n[0]));
    d += L.c[1] * L.g[1] * (double)((j == 1) + (j == L.n[1]));
    d += L.c[2] * L.g[2] * (double)((k == 1) + (k == L.n[2]));
    return d;
}

__device__ __forceinline__ double mg_rhs(const MgPlan &P, const MgLevel &L, int lvl, const double *f, long long c)
{
    return lvl == 0 ? P.rhs_scale * P.rho_pad[c] : f[c];
}

// residual r = f - A u at one cell
__device__ __forceinline__ double mg_residual(const MgPlan &P, const MgLevel &L, int lvl, const double *u, const double *f,
                                              int i, int j, int k)
{
    const long long sy = L.n[2] + 2, sx = (long long)(L.n[1] + 2) * sy;
    const long long c = mg_idx(L, i, j, k);
    double s = L.c[0] * (u[c + sx] + u[c - sx]);
    s += L.c[1] * (u[c + sy] + u[c - sy]);
    s += L.c[2] * (u[c + 1] + u[c - 1]);
    return mg_rhs(P, L, lvl, f, c) - s - mg_diag(L, i, j, k) * u[c];
}

// one colour of a red-black Gauss-Seidel sweep; threads tid, tid + nthr, ... over the cells of the colour
__device__ void mg_smooth_colour(const MgPlan &P, int lvl, double *u, const double *f, int colour, long long tid, long long nthr)
{
    const MgLevel &L = P.lev[lvl];
    const int nx = L.n[0], ny = L.n[1], nz = L.n[2];
    const int hz = (nz + 1) >> 1;
    const long long work = (long long)nx * ny * hz;
    const long long sy = nz + 2, sx = (long long)(ny + 2) * sy;
    for (long long t = tid; t < work; t += nthr) {
        const uint32_t t32 = (uint32_t)t, q = t32 / (uint32_t)hz;
        const int kk = (int)(t32 - q * (uint32_t)hz);
        const int i0 = (int)(q / (uint32_t)ny);
        const int j = 1 + (int)(q - (uint32_t)i0 * (uint32_t)ny);
        const int i = 1 + i0;
        const int k = 1 + 2 * kk + (((i + j + 1) & 1) != colour ? 1 : 0);
        if (k > nz) continue;
        const long long c = i * sx + j * sy + k;
        double s = L.c[0] * (u[c + sx] + u[c - sx]);
        s += L.c[1] * (u[c + sy] + u[c - sy]);
        s += L.c[2] * (u[c + 1] + u[c - 1]);
        u[c] = (mg_rhs(P, L, lvl, f, c) - s) / mg_diag(L, i, j, k);
    }
}

// f_coarse = average of the children's residuals, u_coarse = 0
__device__ void mg_restrict(const MgPlan &P, int lvl, const double *u, const double *f, double *uc, double *fc,
                            long long tid, long long nthr)
{
    const MgLevel &L = P.lev[lvl];
    const MgLevel &C = P.lev[lvl + 1];
    const long long work = (long long)C.n[0] * C.n[1] * C.n[2];
    const int rx = L.co[0] ? 2 : 1, ry = L.co[1] ? 2 : 1, rz = L.co[2] ? 2 : 1;
    const double w = 1.0 / (double)(rx * ry * rz);
    for (long long t = tid; t < work; t += nthr) {
        const uint32_t t32 = (uint32_t)t, q = t32 / (uint32_t)C.n[2];
        const int K = 1 + (int)(t32 - q * (uint32_t)C.n[2]);
        const int I = 1 + (int)(q / (uint32_t)C.n[1]);
        const int J = 1 + (int)(q - (uint32_t)(I - 1) * (uint32_t)C.n[1]);
        double s = 0.0;
        for (int a = 0; a < rx; ++a)
            for (int b = 0; b < ry; ++b)
                for (int d = 0; d < rz; ++d)
                    s += mg_residual(P, L, lvl, u, f, rx * (I - 1) + 1 + a, ry * (J - 1) + 1 + b, rz * (K - 1) + 1 + d);
        const long long cc = mg_idx(C, I, J, K);
        fc[cc] = w * s;
        uc[cc] = 0.0;
    }
}

// value of the coarse correction at (I, J, K) where an index may be 0 or n+1 (ghost = g * first)
__device__ __forceinline__ double mg_coarse_at(const MgLevel &C, const double *e, int I, int J, int K)
{
    double sgn = 1.0;
    if (I < 1) { I = 1; sgn *= C.g[0]; } else if (I > C.n[0]) { I = C.n[0]; sgn *= C.g[0]; }
    if (J < 1) { J = 1; sgn *= C.g[1]; } else if (J > C.n[1]) { J = C.n[1]; sgn *= C.g[1]; }
    if (K < 1) { K = 1; sgn *= C.g[2]; } else if (K > C.n[2]) { K = C.n[2]; sgn *= C.g[2]; }
    return sgn * e[mg_idx(C, I, J, K)];
}

// u_fine += linear interpolation of the coarse correction (3/4 - 1/4 per coarsened axis)
__device__ void mg_prolong(const MgPlan &P, int lvl, double *u, const double *ec, long long tid, long long nthr)
{
    const MgLevel &L = P.lev[lvl];
    const MgLevel &C = P.lev[lvl + 1];
    const long long work = (long long)L.n[0] * L.n[1] * L.n[2];
    for (long long t = tid; t < work; t += nthr) {
        const uint32_t t32 = (uint32_t)t, q = t32 / (uint32_t)L.n[2];
        const int k = 1 + (int)(t32 - q * (uint32_t)L.n[2]);
        const int i = 1 + (int)(q / (uint32_t)L.n[1]);
        const int j = 1 + (int)(q - (uint32_t)(i - 1) * (uint32_t)L.n[1]);
        // parent and the neighbour on the child's side, per axis
        int I0 = i, I1 = i, J0 = j, J1 = j, K0 = k, K1 = k;
        double wx = 0.0, wy = 0.0, wz = 0.0;                 // weight of the far neighbour
        if (L.co[0]) { I0 = (i + 1) >> 1; I1 = (i & 1) ? I0 - 1 : I0 + 1; wx = 0.25; }
        if (L.co[1]) { J0 = (j + 1) >> 1; J1 = (j & 1) ? J0 - 1 : J0 + 1; wy = 0.25; }
        if (L.co[2]) { K0 = (k + 1) >> 1; K1 = (k & 1) ? K0 - 1 : K0 + 1; wz = 0.25; }
        double e = 0.0;
        for (int a = 0; a < (L.co[0] ? 2 : 1); ++a)
            for (int b = 0; b < (L.co[1] ? 2 : 1); ++b)
                for (int d = 0; d < (L.co[2] ? 2 : 1); ++d) {
                    const double w = (a ? wx : 1.0 - wx) * (b ? wy : 1.0 - wy) * (d ? wz : 1.0 - wz);
                    e += w * mg_coarse_at(C, ec, a ? I1 : I0, b ? J1 : J0, d ? K1 : K0);
                }
        u[mg_idx(L, i, j, k)] += e;
    }
}

// the levels that fit one block: a V-cycle from `first` down to the coarsest and back, all in shared
// memory (padded arrays), __syncthreads between the steps
__device__ void mg_block_levels(const MgPlan &P, double *smem)
{
    const int first = P.first_block_level;
    // carve the shared arrays: u and f of every block level
    double *su[MG_MAX_LEVELS], *sf[MG_MAX_LEVELS];
    {
        double *p = smem;
        for (int l = first; l < P.n_levels; ++l) { su[l] = p; p += P.lev[l].pad_cells; sf[l] = p; p += P.lev[l].pad_cells; }
    }
    const long long tid = threadIdx.x, nthr = blockDim.x;
    // load f of the first block level (written by the grid-level restriction), zero everything else
    for (int l = first; l < P.n_levels; ++l)
        for (long long t = tid; t < P.lev[l].pad_cells; t += nthr) {
            su[l][t] = 0.0;
            sf[l][t] = (l == first) ? P.lev[l].f[t] : 0.0;
        }
    __syncthreads();
    for (int l = first; l < P.n_levels; ++l) {
        const bool coarsest = (l == P.n_levels - 1);
        const int sweeps = coarsest ? P.coarse_sweeps : P.nu1;
        for (int s = 0; s < sweeps; ++s)
            for (int colour = 0; colour < 2; ++colour) {
                mg_smooth_colour(P, l, su[l], sf[l], colour, tid, nthr);
                __syncthreads();
            }
        if (!coarsest) {
            mg_restrict(P, l, su[l], sf[l], su[l + 1], sf[l + 1], tid, nthr);
            __syncthreads();
        }
    }
    for (int l = P.n_levels - 2; l >= first; --l) {
        mg_prolong(P, l, su[l], su[l + 1], tid, nthr);
        __syncthreads();
        for (int s = 0; s < P.nu2; ++s)
            for (int colour = 0; colour < 2; ++colour) {
                mg_smooth_colour(P, l, su[l], sf[l], colour, tid, nthr);
                __syncthreads();
            }
    }
    for (long long t = tid; t < P.lev[first].pad_cells; t += nthr) P.lev[first].u[t] = su[first][t];
}

__device__ __forceinline__ double mg_block_sum(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += scratch[w];
        scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// RMS and maximum of the true residual on level 0; every block ends up with the same numbers
__device__ void mg_residual_norm(const MgPlan &P, cg::grid_group &grid, double *partials, double *scratch, double &rms,
                                 double &rmax)
{
    const MgLevel &L = P.lev[0];
    const long long work = (long long)L.n[0] * L.n[1] * L.n[2];
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    double ss = 0.0, mx = 0.0;
    for (long long t = gtid; t < work; t += nthr) {
        const uint32_t t32 = (uint32_t)t, q = t32 / (uint32_t)L.n[2];
        const int k = 1 + (int)(t32 - q * (uint32_t)L.n[2]);
        const int i = 1 + (int)(q / (uint32_t)L.n[1]);
        const int j = 1 + (int)(q - (uint32_t)(i - 1) * (uint32_t)L.n[1]);
        const double r = mg_residual(P, L, 0, L.u, nullptr, i, j, k);
        ss += r * r;
        mx = fmax(mx, fabs(r));
    }
    const double bs = mg_block_sum(ss, scratch);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
    __shared__ double wmax[32];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmax(m, wmax[w]);
        partials[2 * blockIdx.x] = bs;
        partials[2 * blockIdx.x + 1] = m;
    }
    grid.sync();
    double v = 0.0, m = 0.0;
    if (threadIdx.x < 32) {
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) {
            v += __ldcg(&partials[2 * b]);
            m = fmax(m, __ldcg(&partials[2 * b + 1]));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v += __shfl_down_sync(0xffffffffu, v, off);
            m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
        }
        if (threadIdx.x == 0) { scratch[32] = v; scratch[31] = m; }
    }
    __syncthreads();
    rms = sqrt(scratch[32] / (double)work);
    rmax = scratch[31];
    __syncthreads();
    grid.sync();                 // partials may be overwritten by the next call
}

__global__ void __launch_bounds__(MG_THREADS)
poisson_mg_kernel(const __grid_constant__ MgPlan P, double *partials, double *scalars, int32_t *flags, double *res_hist)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double mg_smem[];
    __shared__ double scratch[33];
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const int LB = P.first_block_level;          // grid levels: 0 .. LB-1
    double rms = 0.0, rmax = 0.0;
    mg_residual_norm(P, grid, partials, scratch, rms, rmax);
    const double rms0 = rms;
    int cycle = 0;
    bool bad = !(rms == rms);
    while (cycle < P.max_cycles && rms > P.tol_residual && !bad) {
        for (int l = 0; l < LB; ++l) {
            const MgLevel &L = P.lev[l];
            const bool coarsest = (l == P.n_levels - 1);
            const int sweeps = coarsest ? P.coarse_sweeps : P.nu1;
            for (int s = 0; s < sweeps; ++s)
                for (int colour = 0; colour < 2; ++colour) {
                    mg_smooth_colour(P, l, L.u, L.f, colour, gtid, nthr);
                    grid.sync();
                }
            if (!coarsest) {
                mg_restrict(P, l, L.u, L.f, P.lev[l + 1].u, P.lev[l + 1].f, gtid, nthr);
                grid.sync();
            }
        }
        if (LB < P.n_levels) {
            if (blockIdx.x == 0) mg_block_levels(P, mg_smem);
            grid.sync();
        }
        for (int l = (LB < P.n_levels ? LB - 1 : P.n_levels - 2); l >= 0; --l) {
            const MgLevel &L = P.lev[l];
            mg_prolong(P, l, L.u, P.lev[l + 1].u, gtid, nthr);
            grid.sync();
            for (int s = 0; s < P.nu2; ++s)
                for (int colour = 0; colour < 2; ++colour) {
                    mg_smooth_colour(P, l, L.u, L.f, colour, gtid, nthr);
                    grid.sync();
                }
        }
        mg_residual_norm(P, grid, partials, scratch, rms, rmax);
        if (gtid == 0 && res_hist) res_hist[cycle] = rms;
        ++cycle;
        bad = !(rms == rms) || rms > 1e6 * (rms0 + 1.0);
    }
    if (gtid == 0) {
        scalars[0] = rms;          // RMS true residual
        scalars[1] = rmax;         // max |residual|
        scalars[2] = rms0;         // before the first cycle
        flags[0] = cycle;
        flags[1] = bad ? 1 : 0;
        if (bad) flags[2] = 1;
    }
}

// ------------------------------------------------------------------------------------------
// host: level plan, buffers, launch
// ------------------------------------------------------------------------------------------
struct MgBuffers {
    int seg[3] = {0, 0, 0};
    std::vector<double *> u, f;        // per coarse level (index 1..)
    double *partials = nullptr;
    double *res_hist = nullptr;
    int res_hist_cap = 0;
};

static MgBuffers *mg_buffers(Ctx *ctx)
{
    if (!ctx->mg) ctx->mg = new MgBuffers();
    return static_cast<MgBuffers *>(ctx->mg);
}

void mg_release(Ctx *ctx)
{
    if (!ctx->mg) return;
    MgBuffers *B = static_cast<MgBuffers *>(ctx->mg);
    for (double *p : B->u) if (p) cudaFree(p);
    for (double *p : B->f) if (p) cudaFree(p);
    if (B->partials) cudaFree(B->partials);
    if (B->res_hist) cudaFree(B->res_hist);
    delete B;
    ctx->mg = nullptr;
}

// ghost factor after m coarsenings of an axis: the reference's Dirichlet node is one FINE spacing
// outside the first fine cell centre = d = 1/2 + 2^-(m+1) spacings of the level
static double mg_ghost(int m)
{
    if (m == 0) return 0.0;
    const double d = 0.5 + std::ldexp(1.0, -(m + 1));
    return -(1.0 - d) / d;
}

static int mg_build_plan(Ctx *ctx, MgPlan &P, double *u_pad, const double *rho_pad)
{
    const double *dl = ctx->dc.dl;
    int n[3] = {ctx->dc.seg[0], ctx->dc.seg[1], ctx->dc.seg[2]};
    double c[3] = {1.0, (dl[0] / dl[1]) * (dl[0] / dl[1]), (dl[0] / dl[2]) * (dl[0] / dl[2])};   // poisson_.py:56-57
    int m[3] = {0, 0, 0};
    P.n_levels = 0;
    while (P.n_levels < MG_MAX_LEVELS) {
        MgLevel &L = P.lev[P.n_levels++];
        for (int a = 0; a < 3; ++a) { L.n[a] = n[a]; L.c[a] = c[a]; L.g[a] = mg_ghost(m[a]); L.co[a] = 0; }
        L.pad_cells = (long long)(n[0] + 2) * (n[1] + 2) * (n[2] + 2);
        L.u = L.f = nullptr;
        double cmax = 0.0;
        bool any = false;
        for (int a = 0; a < 3; ++a)
            if (n[a] % 2 == 0 && n[a] >= 4) { any = true; cmax = std::fmax(cmax, c[a]); }
        if (!any || P.n_levels == MG_MAX_LEVELS) break;
        for (int a = 0; a < 3; ++a)
            if (n[a] % 2 == 0 && n[a] >= 4 && c[a] >= 0.3 * cmax) { L.co[a] = 1; n[a] /= 2; c[a] /= 4.0; m[a] += 1; }
    }
    // block levels: the tail of the hierarchy whose levels all fit MG_BLOCK_CELLS cells (and shared memory)
    P.first_block_level = P.n_levels;
    size_t smem = 0;
    for (int l = P.n_levels - 1; l >= 1; --l) {
        const long long cells = (long long)P.lev[l].n[0] * P.lev[l].n[1] * P.lev[l].n[2];
        const size_t need = smem + 2 * sizeof(double) * (size_t)P.lev[l].pad_cells;
        if (cells > MG_BLOCK_CELLS || need > 160 * 1024) break;
        smem = need;
        P.first_block_level = l;
    }
    MgBuffers *B = mg_buffers(ctx);
    if (B->seg[0] != ctx->dc.seg[0] || B->seg[1] != ctx->dc.seg[1] || B->seg[2] != ctx->dc.seg[2] ||
        (int)B->u.size() != P.n_levels) {
        for (double *p : B->u) if (p) cudaFree(p);
        for (double *p : B->f) if (p) cudaFree(p);
        B->u.assign(P.n_levels, nullptr);
        B->f.assign(P.n_levels, nullptr);
        for (int l = 1; l < P.n_levels; ++l) {
            const size_t bytes = sizeof(double) * (size_t)P.lev[l].pad_cells;
            if (cudaMalloc(&B->u[l], bytes) != cudaSuccess || cudaMalloc(&B->f[l], bytes) != cudaSuccess) return EMC_ERR_CUDA;
            cudaMemset(B->u[l], 0, bytes);      // the halo of the coarse levels stays zero for ever
            cudaMemset(B->f[l], 0, bytes);
        }
        for (int a = 0; a < 3; ++a) B->seg[a] = ctx->dc.seg[a];
    }
    if (!B->partials && cudaMalloc(&B->partials, sizeof(double) * 2 * 4096) != cudaSuccess) return EMC_ERR_CUDA;
    for (int l = 1; l < P.n_levels; ++l) { P.lev[l].u = B->u[l]; P.lev[l].f = B->f[l]; }
    P.lev[0].u = u_pad;
    P.lev[0].f = nullptr;
    P.rho_pad = rho_pad;
    P.rhs_scale = dl[0] * dl[0] / ctx->params.eps_s;     // poisson_.py:76
    return (int)smem;
}

int run_poisson_mg(Ctx *ctx, double *u_pad, const double *rho_pad, double tol_residual, int max_cycles, int nu1, int nu2,
                   int32_t *cycles_out, double *res_out, double *res_hist, cudaStream_t st)
{
    if (max_cycles < 1 || nu1 < 0 || nu2 < 0 || nu1 + nu2 < 1) { ctx->last_error = "multigrid: max_cycles >= 1, nu1 + nu2 >= 1"; return EMC_ERR_INVALID; }
    MgPlan P;
    const int smem = mg_build_plan(ctx, P, u_pad, rho_pad);
    if (smem < 0) { ctx->last_error = "multigrid: cudaMalloc of the coarse levels failed"; return EMC_ERR_CUDA; }
    if (P.n_levels < 2) { ctx->last_error = "multigrid: this mesh cannot be coarsened (every axis odd or < 4 cells): use the SOR orders"; return EMC_ERR_UNSUPPORTED; }
    P.tol_residual = tol_residual;
    P.max_cycles = max_cycles;
    P.nu1 = nu1; P.nu2 = nu2;
    P.coarse_sweeps = 32;
    MgBuffers *B = mg_buffers(ctx);
    cudaError_t e;
    if (B->res_hist_cap < max_cycles) {
        if (B->res_hist) cudaFree(B->res_hist);
        B->res_hist = nullptr;
        if ((e = cudaMalloc(&B->res_hist, sizeof(double) * (size_t)max_cycles)) != cudaSuccess) goto cuda_fail;
        B->res_hist_cap = max_cycles;
    }
    if ((e = cudaMemsetAsync(ctx->d_sor_flags, 0, 2 * sizeof(int32_t), st)) != cudaSuccess) goto cuda_fail;
    {
        static bool optin[16];
        if (smem > 48 * 1024 && !optin[ctx->device & 15]) {
            if ((e = cudaFuncSetAttribute(poisson_mg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) goto cuda_fail;
            optin[ctx->device & 15] = true;
        }
        int per_sm = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, poisson_mg_kernel, MG_THREADS, (size_t)smem)) != cudaSuccess) goto cuda_fail;
        if (per_sm < 1) { ctx->last_error = "multigrid: kernel does not fit an SM"; return EMC_ERR_CUDA; }
        if (per_sm > 2) per_sm = 2;
        const long long cells = (long long)P.lev[0].n[0] * P.lev[0].n[1] * P.lev[0].n[2];
        int grid = ctx->sm_count * per_sm;
        const long long want = (cells / 2 + MG_THREADS - 1) / MG_THREADS;
        if (want < grid) grid = (int)want;
        if (grid < 1) grid = 1;
        if (grid > 4096) grid = 4096;
        double *pt = B->partials, *sc = ctx->d_sor_scalars, *rh = B->res_hist;
        int32_t *fl = ctx->d_sor_flags;
        void *args[] = {(void *)&P, (void *)&pt, (void *)&sc, (void *)&fl, (void *)&rh};
        ctx->launches++;
        e = cudaLaunchCooperativeKernel((void *)poisson_mg_kernel, dim3(grid), dim3(MG_THREADS), args, (size_t)smem, st);
        if (e != cudaSuccess) goto cuda_fail;
    }
    if (cycles_out || res_out || res_hist) {
        double scal[3];
        int32_t fl[2];
        if ((e = cudaMemcpyAsync(scal, ctx->d_sor_scalars, sizeof(scal), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(fl, ctx->d_sor_flags, sizeof(fl), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
            (e = cudaStreamSynchronize(st)) != cudaSuccess)
            goto cuda_fail;
        if (cycles_out) *cycles_out = fl[0];
        if (res_out) *res_out = scal[0];
        if (res_hist && fl[0] > 0 &&
            (e = cudaMemcpy(res_hist, B->res_hist, sizeof(double) * (size_t)fl[0], cudaMemcpyDeviceToHost)) != cudaSuccess)
            goto cuda_fail;
        if (fl[1]) { ctx->last_error = "multigrid diverged (NaN or growing residual)"; return EMC_ERR_DIVERGED; }
    }
    return EMC_OK;
cuda_fail:
    ctx->last_error = std::string("CUDA: ") + cudaGetErrorString(e);
    return EMC_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------
// true residual of ANY potential (whatever solver produced it): RMS and max of
//   f - A u = dl0^2 rho / eps - [sum of neighbours - 2 (1 + xy + xz) u]          (volts)
// over the interior cells, in a fixed reduction order
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
poisson_residual_kernel(const __grid_constant__ MgPlan P, double *partials, unsigned int *ticket, double *out)
{
    __shared__ double scratch[33];
    __shared__ double wmax[32];
    __shared__ bool is_last;
    const MgLevel &L = P.lev[0];
    const long long work = (long long)L.n[0] * L.n[1] * L.n[2];
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    double ss = 0.0, mx = 0.0;
    for (long long t = gtid; t < work; t += nthr) {
        const uint32_t t32 = (uint32_t)t, q = t32 / (uint32_t)L.n[2];
        const int k = 1 + (int)(t32 - q * (uint32_t)L.n[2]);
        const int i = 1 + (int)(q / (uint32_t)L.n[1]);
        const int j = 1 + (int)(q - (uint32_t)(i - 1) * (uint32_t)L.n[1]);
        const double r = mg_residual(P, L, 0, L.u, nullptr, i, j, k);
        ss += r * r;
        mx = fmax(mx, fabs(r));
    }
    const double bs = mg_block_sum(ss, scratch);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmax(m, wmax[w]);
        partials[2 * blockIdx.x] = bs;
        partials[2 * blockIdx.x + 1] = m;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        double s = 0.0, m = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) { s += __ldcg(&partials[2 * b]); m = fmax(m, __ldcg(&partials[2 * b + 1])); }
        out[0] = sqrt(s / (double)work);
        out[1] = m;
        *ticket = 0u;
    }
}

int run_poisson_residual(Ctx *ctx, const double *u_pad, const double *rho_pad, double *rms_out, double *max_out,
                         double *out_dev, cudaStream_t st)
{
    MgPlan P;
    P.n_levels = 1;
    P.first_block_level = 1;
    MgLevel &L = P.lev[0];
    const double *dl = ctx->dc.dl;
    const double c[3] = {1.0, (dl[0] / dl[1]) * (dl[0] / dl[1]), (dl[0] / dl[2]) * (dl[0] / dl[2])};
    for (int a = 0; a < 3; ++a) { L.n[a] = ctx->dc.seg[a]; L.c[a] = c[a]; L.g[a] = 0.0; L.co[a] = 0; }
    L.pad_cells = (long long)(L.n[0] + 2) * (L.n[1] + 2) * (L.n[2] + 2);
    L.u = const_cast<double *>(u_pad);
    L.f = nullptr;
    P.rho_pad = rho_pad;
    P.rhs_scale = dl[0] * dl[0] / ctx->params.eps_s;
    MgBuffers *B = mg_buffers(ctx);
    cudaError_t e;
    if (!B->partials && (e = cudaMalloc(&B->partials, sizeof(double) * 2 * 4096)) != cudaSuccess) goto cuda_fail;
    {
        const long long cells = (long long)L.n[0] * L.n[1] * L.n[2];
        int grid = (int)((cells + 255) / 256);
        if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
        if (grid < 1) grid = 1;
        double *dst = out_dev ? out_dev : ctx->d_sor_scalars + 4;
        ctx->launches++;
        poisson_residual_kernel<<<grid, 256, 0, st>>>(P, B->partials, ctx->ticket, dst);
        if ((e = cudaGetLastError()) != cudaSuccess) goto cuda_fail;
        if (rms_out || max_out) {
            double h[2];
            if ((e = cudaMemcpyAsync(h, dst, sizeof(h), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                (e = cudaStreamSynchronize(st)) != cudaSuccess)
                goto cuda_fail;
            if (rms_out) *rms_out = h[0];
            if (max_out) *max_out = h[1];
        }
    }
    return EMC_OK;
cuda_fail:
    ctx->last_error = std::string("CUDA: ") + cudaGetErrorString(e);
    return EMC_ERR_CUDA;
}

}  // namespace emc
