// emc_mesh.cu -- charge assignment (poisson_.py:30-44), rho assembly (poisson_.py:27,36,42,53),
// Poisson SOR (poisson_.py:47-84) and field (poisson_.py:88-96) on the device-resident mesh.
// Mesh layout everywhere: C order of (nx, ny, nz) / padded (nx+2, ny+2, nz+2), z fastest, like
// rho.reshape(dev.tot_seg) in the reference.  Compiled with -fmad=false.
#include <cooperative_groups.h>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "emc_host.h"

namespace cg = cooperative_groups;

namespace emc {

// ------------------------------------------------------------------------------------------
// charge assignment
// ------------------------------------------------------------------------------------------
// cloud-in-cell weights on one axis: the two nearest centres, w = 1 - |centre - r|/dl
// (the commented line poisson_.py:34)
__device__ __forceinline__ void axis_cic(double r, double dl, int j[2], double w[2])
{
    const double t = r / dl - 0.5;
    const double f = floor(t);
    const double w1 = t - f;
    j[0] = (int)f;
    j[1] = j[0] + 1;
    w[0] = 1.0 - w1;
    w[1] = w1;
}

template <bool SMEM>
__global__ void __launch_bounds__(256)
deposit_ngp_kernel(const __grid_constant__ DevConst P, const double *__restrict__ x, const double *__restrict__ y,
                   const double *__restrict__ z, int64_t n, unsigned long long *mesh, int cells)
{
    extern __shared__ unsigned int scount[];
    if (SMEM) {
        for (int t = threadIdx.x; t < cells; t += blockDim.x) scount[t] = 0u;
        __syncthreads();
    }
    // four particles per thread and pass: all twelve position loads are issued before the first
    // is consumed (the kernel is latency bound: ~100 instructions and one atomic per 24 bytes)
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        double px[U], py[U], pz[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) { px[u] = x[i]; py[u] = y[i]; pz[u] = z[i]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * stride >= n) break;
            int hx[3], hy[3], hz[3];
            const int nx = axis_hits(px[u], P.seg[0], P.dl[0], P.inv_dl[0], hx);
            const int ny = axis_hits(py[u], P.seg[1], P.dl[1], P.inv_dl[1], hy);
            const int nz = axis_hits(pz[u], P.seg[2], P.dl[2], P.inv_dl[2], hz);
            for (int a = 0; a < nx; ++a)
                for (int b = 0; b < ny; ++b)
                    for (int c = 0; c < nz; ++c) {
                        const int64_t cell = ((int64_t)hx[a] * P.seg[1] + hy[b]) * P.seg[2] + hz[c];
                        if (SMEM) atomicAdd(&scount[cell], 1u);
                        else atomicAdd(mesh + cell, 1ull);
                    }
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int t = threadIdx.x; t < cells; t += blockDim.x) {
            const unsigned int v = scount[t];
            if (v) atomicAdd(mesh + t, (unsigned long long)v);
        }
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(256)
deposit_cic_kernel(const __grid_constant__ DevConst P, const double *__restrict__ x, const double *__restrict__ y,
                   const double *__restrict__ z, int64_t n, unsigned long long *mesh, int cells)
{
    extern __shared__ unsigned long long sfx[];
    if (SMEM) {
        for (int t = threadIdx.x; t < cells; t += blockDim.x) sfx[t] = 0ull;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int jx[2], jy[2], jz[2];
        double wx[2], wy[2], wz[2];
        axis_cic(x[i], P.dl[0], jx, wx);
        axis_cic(y[i], P.dl[1], jy, wy);
        axis_cic(z[i], P.dl[2], jz, wz);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (jx[a] < 0 || jx[a] >= P.seg[0] || jy[b] < 0 || jy[b] >= P.seg[1] || jz[c] < 0 ||
                        jz[c] >= P.seg[2])
                        continue;
                    const double w = wx[a] * wy[b] * wz[c];
                    const unsigned long long q = (unsigned long long)__double2ll_rn(w * 4294967296.0);
                    const int64_t cell = ((int64_t)jx[a] * P.seg[1] + jy[b]) * P.seg[2] + jz[c];
                    if (SMEM) atomicAdd(&sfx[cell], q);
                    else atomicAdd(mesh + cell, q);
                }
    }
    if (SMEM) {
        __syncthreads();
        for (int t = threadIdx.x; t < cells; t += blockDim.x) {
            const unsigned long long v = sfx[t];
            if (v) atomicAdd(mesh + t, v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// NGP on a mesh that does not fit one block's shared memory: the tiled variant of the north-star
// sketch ("sort by cell, per-tile partial meshes in shared memory, 64-bit commit").
//
// The mesh is cut into tiles of DEP_TILE_CELLS consecutive cells (16 KB of 32-bit counts).  Three
// passes, all integer, hence bit-identical to the direct kernel and order independent:
//   1. cells:   per particle the cell of its FIRST hit -> cell_of[n] (4 B) and a per-tile histogram
//               (shared-memory partials, one global add per block and tile); the additional cells of
//               a particle ON a face (poisson_.py:31 inclusive test; ~1e-9 of random positions) go
//               straight to the mesh;
//   2. scatter: after an exclusive scan of the histogram every block reserves, per tile, a range of
//               the tile's segment (one global atomic per block and tile) and writes its cell ids there:
//               a counting sort of the 4-byte cell ids by tile -- the positions are not moved;
//   3. tiles:   one block per tile portion accumulates its cell ids into a shared-memory partial mesh
//               and commits every touched cell with ONE 64-bit add.
// Traffic 24 + 4 + 4 + 4 + 4 = 40 B per particle against 24 B + one L2 atomic per particle for the
// direct kernel.  Measured (profiles/r2_deposit_ab.txt): see launch_deposit.
// ------------------------------------------------------------------------------------------
constexpr int DEP_TILE_CELLS = 4096;
constexpr int DEP_MAX_TILES = 4096;          // shared-memory histogram of the first two passes (16 KB)

__global__ void __launch_bounds__(256)
deposit_tiled_cells_kernel(const __grid_constant__ DevConst P, const double *__restrict__ x, const double *__restrict__ y,
                           const double *__restrict__ z, int64_t n, unsigned long long *mesh, int32_t *cell_of,
                           unsigned int *tile_count, int n_tiles)
{
    extern __shared__ unsigned int hist[];
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) hist[t] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int hx[3], hy[3], hz[3];
        const int nx = axis_hits(x[i], P.seg[0], P.dl[0], P.inv_dl[0], hx);
        const int ny = axis_hits(y[i], P.seg[1], P.dl[1], P.inv_dl[1], hy);
        const int nz = axis_hits(z[i], P.seg[2], P.dl[2], P.inv_dl[2], hz);
        int32_t first = -1;
        for (int a = 0; a < nx; ++a)
            for (int b = 0; b < ny; ++b)
                for (int c = 0; c < nz; ++c) {
                    const int64_t cell = ((int64_t)hx[a] * P.seg[1] + hy[b]) * P.seg[2] + hz[c];
                    if (first < 0) first = (int32_t)cell;
                    else atomicAdd(mesh + cell, 1ull);            // a face / edge / corner hit: rare
                }
        cell_of[i] = first;                                       // -1: outside the mesh (counts nowhere)
        if (first >= 0) atomicAdd(&hist[first / DEP_TILE_CELLS], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x)
        if (hist[t]) atomicAdd(&tile_count[t], hist[t]);
}

// exclusive scan of the tile histogram (one block); cursor[t] = start[t] for the scatter pass
__global__ void deposit_tiled_scan_kernel(const unsigned int *tile_count, unsigned int *tile_start, unsigned int *cursor,
                                          int n_tiles)
{
    __shared__ unsigned int part[1024];
    const int per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(n_tiles, lo + per);
    unsigned int s = 0;
    for (int t = lo; t < hi; ++t) s += tile_count[t];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) { const unsigned int v = part[t]; part[t] = run; run += v; }
    }
    __syncthreads();
    unsigned int run = part[threadIdx.x];
    for (int t = lo; t < hi; ++t) { tile_start[t] = run; cursor[t] = run; run += tile_count[t]; }
    if (hi == n_tiles && lo < hi) tile_start[n_tiles] = run;
    if (n_tiles == 0 && threadIdx.x == 0) tile_start[0] = 0;
}

__global__ void __launch_bounds__(256)
deposit_tiled_scatter_kernel(const int32_t *__restrict__ cell_of, int64_t n, unsigned int *cursor, int32_t *sorted,
                             int n_tiles, int64_t chunk)
{
    extern __shared__ unsigned int sh[];
    unsigned int *hist = sh, *base = sh + n_tiles;
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(n, lo + chunk);
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) hist[t] = 0u;
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int32_t c = cell_of[i];
        if (c >= 0) atomicAdd(&hist[c / DEP_TILE_CELLS], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        base[t] = hist[t] ? atomicAdd(&cursor[t], hist[t]) : 0u;   // this block's range of tile t
        hist[t] = 0u;
    }
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int32_t c = cell_of[i];
        if (c >= 0) {
            const int t = c / DEP_TILE_CELLS;
            sorted[base[t] + atomicAdd(&hist[t], 1u)] = c;
        }
    }
}

__global__ void __launch_bounds__(256)
deposit_tiled_tiles_kernel(const int32_t *__restrict__ sorted, const unsigned int *__restrict__ tile_start,
                           unsigned long long *mesh, int n_tiles, int parts, int64_t cells)
{
    __shared__ unsigned int cnt[DEP_TILE_CELLS];
    const int tile = blockIdx.x / parts, part = blockIdx.x % parts;
    for (int t = threadIdx.x; t < DEP_TILE_CELLS; t += blockDim.x) cnt[t] = 0u;
    __syncthreads();
    const unsigned int a = tile_start[tile], b = tile_start[tile + 1];
    const unsigned int len = b - a, per = (len + parts - 1) / parts;
    const unsigned int lo = a + part * per, hi = min(b, lo + per);
    const int32_t c0 = tile * DEP_TILE_CELLS;
    for (unsigned int e = lo + threadIdx.x; e < hi; e += blockDim.x) atomicAdd(&cnt[sorted[e] - c0], 1u);
    __syncthreads();
    for (int t = threadIdx.x; t < DEP_TILE_CELLS; t += blockDim.x) {
        const unsigned int v = cnt[t];
        if (v && (int64_t)c0 + t < cells) atomicAdd(mesh + c0 + t, (unsigned long long)v);
    }
}

static cudaError_t launch_deposit_tiled(Ctx *ctx, const double *x, const double *y, const double *z, int64_t n,
                                        unsigned long long *m, int64_t cells, cudaStream_t st)
{
    const int n_tiles = (int)((cells + DEP_TILE_CELLS - 1) / DEP_TILE_CELLS);
    cudaError_t e;
    // scratch: cell_of[n], sorted[n] (int32), tile_count / tile_start / cursor
    const size_t need = sizeof(int32_t) * 2 * (size_t)n + sizeof(unsigned int) * (3 * (size_t)n_tiles + 4);
    if (ctx->dep_scratch_bytes < need) {
        if (ctx->d_dep_scratch) cudaFree(ctx->d_dep_scratch);
        ctx->d_dep_scratch = nullptr;
        ctx->dep_scratch_bytes = 0;
        if ((e = cudaMalloc(&ctx->d_dep_scratch, need)) != cudaSuccess) return e;
        ctx->dep_scratch_bytes = need;
    }
    int32_t *cell_of = reinterpret_cast<int32_t *>(ctx->d_dep_scratch);
    int32_t *sorted = cell_of + n;
    unsigned int *tile_count = reinterpret_cast<unsigned int *>(sorted + n);
    unsigned int *tile_start = tile_count + n_tiles, *cursor = tile_start + n_tiles + 1;
    if ((e = cudaMemsetAsync(tile_count, 0, sizeof(unsigned int) * (size_t)n_tiles, st)) != cudaSuccess) return e;
    int grid = (int)((n + 255) / 256 < (int64_t)ctx->sm_count * 8 ? (n + 255) / 256 : (int64_t)ctx->sm_count * 8);
    if (grid < 1) grid = 1;
    ctx->launches += 4;
    deposit_tiled_cells_kernel<<<grid, 256, sizeof(unsigned int) * (size_t)n_tiles, st>>>(ctx->dc, x, y, z, n, m, cell_of,
                                                                                        tile_count, n_tiles);
    deposit_tiled_scan_kernel<<<1, 1024, 0, st>>>(tile_count, tile_start, cursor, n_tiles);
    const int64_t chunk = (n + grid - 1) / grid;
    deposit_tiled_scatter_kernel<<<grid, 256, sizeof(unsigned int) * 2 * (size_t)n_tiles, st>>>(cell_of, n, cursor, sorted,
                                                                                            n_tiles, chunk);
    // enough blocks per tile to fill the machine
    int parts = (ctx->sm_count * 4 + n_tiles - 1) / n_tiles;
    if (parts < 1) parts = 1;
    deposit_tiled_tiles_kernel<<<n_tiles * parts, 256, 0, st>>>(sorted, tile_start, m, n_tiles, parts, cells);
    return cudaGetLastError();
}

// NGP on a mesh that does not fit one block's shared memory.  A/B on B200, 1024 x 1 x 1024 mesh
// (profiles/r2_deposit_ab.txt): the direct kernel -- one 64-bit L2 reduction per particle -- against the
// tiled pipeline above; emc_set_deposit_mode picks (0 = auto: the measured winner, 1 = direct, 2 = tiled).
static cudaError_t launch_deposit_large(Ctx *ctx, const double *x, const double *y, const double *z, int64_t n,
                                        unsigned long long *m, int64_t cells, int grid, cudaStream_t st)
{
    const int n_tiles = (int)((cells + DEP_TILE_CELLS - 1) / DEP_TILE_CELLS);
    const bool tiled_ok = n_tiles <= DEP_MAX_TILES && n < ((int64_t)1 << 31);
    const bool tiled = ctx->deposit_mode == 2 && tiled_ok;
    if (tiled) {
        ctx->launches--;
        return launch_deposit_tiled(ctx, x, y, z, n, m, cells, st);
    }
    deposit_ngp_kernel<false><<<grid, 256, 0, st>>>(ctx->dc, x, y, z, n, m, (int)cells);
    return cudaGetLastError();
}

cudaError_t launch_deposit(Ctx *ctx, const double *x, const double *y, const double *z, int64_t n, int64_t *mesh,
                           int scheme, cudaStream_t st)
{
    const int64_t cells = (int64_t)ctx->dc.seg[0] * ctx->dc.seg[1] * ctx->dc.seg[2];
    const int64_t blocks_needed = (n + 255) / 256;
    int grid = (int)(blocks_needed < (int64_t)ctx->max_grid ? blocks_needed : ctx->max_grid);
    if (grid < 1) grid = 1;
    unsigned long long *m = reinterpret_cast<unsigned long long *>(mesh);
    ctx->launches++;
    // dynamic shared memory above 48 KB needs the per-kernel opt-in (once per kernel and device)
    static bool optin_ngp[16], optin_cic[16];
    const int dev = ctx->device & 15;
    cudaError_t e;
    if (scheme == EMC_DEPOSIT_NGP) {
        if (cells * 4 <= SMEM_MESH_BYTES) {
            if (cells * 4 > 48 * 1024 && !optin_ngp[dev]) {
                if ((e = cudaFuncSetAttribute(deposit_ngp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SMEM_MESH_BYTES)) != cudaSuccess) return e;
                optin_ngp[dev] = true;
            }
            deposit_ngp_kernel<true><<<grid, 256, (size_t)cells * 4, st>>>(ctx->dc, x, y, z, n, m, (int)cells);
        } else {
            return launch_deposit_large(ctx, x, y, z, n, m, cells, grid, st);
        }
    } else {
        if (cells * 8 <= SMEM_MESH_BYTES) {
            if (cells * 8 > 48 * 1024 && !optin_cic[dev]) {
                if ((e = cudaFuncSetAttribute(deposit_cic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SMEM_MESH_BYTES)) != cudaSuccess) return e;
                optin_cic[dev] = true;
            }
            deposit_cic_kernel<true><<<grid, 256, (size_t)cells * 8, st>>>(ctx->dc, x, y, z, n, m, (int)cells);
        } else {
            deposit_cic_kernel<false><<<grid, 256, 0, st>>>(ctx->dc, x, y, z, n, m, (int)cells);
        }
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// charge-mesh all-reduce over NVLink peer memory (one kernel: reduce-scatter + all-gather)
//
// Every rank owns one slice of the mesh: it SUMS that slice over all ranks' input meshes with
// peer loads and writes the total into every rank's output mesh with peer stores.  Inputs are
// complete before the launch and outputs are read after it (the caller brackets the launch with
// a symmetric-memory barrier), so the kernel itself never waits on another GPU.  int64 sums:
// bit-exact and independent of rank count and order.
// ------------------------------------------------------------------------------------------
struct PeerMeshes {
    const longlong2 *in[EMC_MAX_PEERS];
    longlong2 *out[EMC_MAX_PEERS];
};

__global__ void __launch_bounds__(256)
mesh_allreduce_p2p_kernel(const __grid_constant__ PeerMeshes M, int world, int64_t first2, int64_t count2)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count2; t += stride) {
        const int64_t c = first2 + t;
        longlong2 s = make_longlong2(0, 0);
        for (int p = 0; p < world; ++p) {
            const longlong2 v = M.in[p][c];
            s.x += v.x;
            s.y += v.y;
        }
        for (int p = 0; p < world; ++p) M.out[p][c] = s;
    }
}

cudaError_t launch_mesh_allreduce_p2p(Ctx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs, int world,
                                      int rank, int64_t n_cells, cudaStream_t st)
{
    PeerMeshes M;
    for (int p = 0; p < world; ++p) {
        M.in[p] = reinterpret_cast<const longlong2 *>(in_ptrs[p]);
        M.out[p] = reinterpret_cast<longlong2 *>(out_ptrs[p]);
    }
    const int64_t pairs = (n_cells + 1) / 2;                 // meshes are padded to an even cell count
    const int64_t per = (pairs + world - 1) / world;
    const int64_t first = (int64_t)rank * per;
    const int64_t count = first >= pairs ? 0 : (first + per > pairs ? pairs - first : per);
    if (count == 0) return cudaSuccess;
    int grid = (int)((count + 255) / 256);
    if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
    ctx->launches++;
    mesh_allreduce_p2p_kernel<<<grid, 256, 0, st>>>(M, world, first, count);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// r after m sequential `r = r + inc` in round-to-nearest-even double arithmetic -- the
// reference's accumulation `rho[idx] += charge` once per particle (poisson_.py:36) -- WITHOUT m
// dependent additions.  (A cell of the 1-D diode mesh holds 10^4 super-particles: the literal loop
// was 40-80 us of one serial DADD chain per thread, a quarter of that configuration's timestep.)
//
// While r stays strictly inside one binade [2^e, 2^(e+1)) (in magnitude) every value is a
// multiple of u = 2^(e-52), so RN(r + inc) = r + d with d = inc rounded to the nearest multiple
// of u -- the same d for every step of the binade, as long as inc/u is not exactly half-way
// between two integers (then the tie rule looks at the parity of r: taken literally).  So the
// steps of a binade collapse into one exact multiply-add (k*d and r + k*d are multiples of u below
// 2^(e+1): representable), with literal single steps only within two steps of a binade boundary,
// at zero / subnormals, and for ties.  d == 0 means r has stopped moving.  Exactly the literal
// result for every input; emc_selftest_accumulate checks it against the literal loop.
// ------------------------------------------------------------------------------------------
__device__ double add_repeated(double r, double inc, long long m)
{
    if (inc < 0) return -add_repeated(-r, -inc, m);          // RN is symmetric (one level: -inc > 0)
    if (!(inc > 0)) {                                        // zero or NaN increment
        for (long long t = 0; t < m && t < 2; ++t) r = r + inc;
        return r;
    }
    while (m > 0) {
        const double a = fabs(r);
        bool literal = !(a >= 4.450147717014403e-308) || !(a < 4.49e307);     // zero, subnormal fringe, huge, NaN
        long long k = 0;
        double d = 0.0;
        if (!literal) {
            const int e = (int)((__double2hiint(a) >> 20) & 0x7ff) - 1023;    // a in [2^e, 2^(e+1))
            const double u = __hiloint2double((e - 52 + 1023) << 20, 0);       // 2^(e-52), normal: e >= -1021
            const double lo = __hiloint2double((e + 1023) << 20, 0);           // 2^e
            const double qf = inc * __hiloint2double((52 - e + 1023) << 20, 0);   // inc / u, exact unless it overflows
            if (!(qf < 4503599627370496.0)) literal = true;                    // inc >= 2^52 u: one step leaves the binade
            else {
                const double q = rint(qf);
                if (fabs(qf - q) == 0.5) literal = true;                       // tie: depends on the parity of r
                else {
                    d = q * u;
                    if (d == 0.0) return r;                                    // inc < u/2: r never changes again
                    // room to the boundary the walk approaches, keeping two steps of margin
                    const double room = r > 0 ? (2 * lo - r) - 2 * d : (a - lo) - 2 * d;
                    if (!(room > d)) literal = true;
                    else {
                        const double kf = floor(room / d) - 1;
                        k = kf > 9.0e18 ? m : (long long)kf;
                        if (k > m) k = m;
                        if (k < 1) literal = true;
                    }
                }
            }
        }
        if (literal) { r = r + inc; --m; }
        else { r = r + (double)k * d; m -= k; }
    }
    return r;
}

__global__ void selftest_accumulate_kernel(int64_t n, uint64_t seed, int max_count, unsigned long long *bad)
{
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    unsigned long long mism = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 r0 = philox4x32_10(make_uint4(0u, 0xacc0acc0u, (uint32_t)i, (uint32_t)(i >> 32)), key);
        const uint4 r1 = philox4x32_10(make_uint4(1u, 0xacc0acc0u, (uint32_t)i, (uint32_t)(i >> 32)), key);
        // background of either sign over 24 binades, increment over 40 binades, sometimes with few
        // mantissa bits (ties, exact sums) and sometimes exactly cancelling the background
        double bg = ldexp(1.0 + u53(r0.x, r0.y), (int)(r0.z % 24u) - 12);
        if (r0.w & 1u) bg = -bg;
        if ((r0.w & 0xf0u) == 0) bg = 0.0;
        double inc = ldexp(1.0 + u53(r1.x, r1.y), (int)(r1.z % 40u) - 30);
        if (r1.w & 2u) inc = ldexp((double)(1u + (r1.x & 7u)), (int)(r1.z % 40u) - 32);     // short mantissa
        if ((r1.w & 0x300u) == 0 && bg < 0) inc = -bg / (double)(1 + (r1.y % 50u));         // passes through zero
        if (r1.w & 4u) inc = -inc;
        const long long m = (long long)(u53(r0.y, r1.y) * u53(r0.x, r1.w) * max_count);
        double lit = bg;
        for (long long t = 0; t < m; ++t) lit = lit + inc;
        const double fast = add_repeated(bg, inc, m);
        mism += __double_as_longlong(lit) != __double_as_longlong(fast);
    }
    if (mism) atomicAdd(bad, mism);
}

cudaError_t launch_selftest_accumulate(Ctx *ctx, int64_t n, uint64_t seed, int max_count, unsigned long long *bad_dev,
                                       cudaStream_t st)
{
    ctx->launches++;
    selftest_accumulate_kernel<<<ctx->sm_count * 8, 128, 0, st>>>(n, seed, max_count, bad_dev);
    return cudaGetLastError();
}

// rho_pad: zero halo (np.pad, poisson_.py:53); interior = background + charge
__global__ void mesh_to_rho_kernel(const __grid_constant__ DevConst P, const long long *__restrict__ mesh, int scheme,
                                   double background, const double *__restrict__ background_z, double increment,
                                   double *rho_pad, int interior_only)
{
    const int nx = P.seg[0], ny = P.seg[1], nz = P.seg[2];
    // (32-bit index arithmetic: a 64-bit division costs ~80 instructions, and there are three per
    // cell; emc_create refuses meshes of 2^31 padded cells or more)
    int i, j, k;
    int64_t c;
    if (interior_only) {
        // the halo is already zero (a buffer this context has written before): one thread per
        // interior cell, a third of the padded cells on an ny = 1 mesh
        const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= (int64_t)nx * ny * nz) return;
        const uint32_t t32 = (uint32_t)t, q32 = t32 / (uint32_t)nz;
        k = 1 + (int)(t32 - q32 * (uint32_t)nz);
        i = 1 + (int)(q32 / (uint32_t)ny);
        j = 1 + (int)(q32 - (uint32_t)(i - 1) * (uint32_t)ny);
        c = ((int64_t)i * (ny + 2) + j) * (nz + 2) + k;
    } else {
        const int64_t total = (int64_t)(nx + 2) * (ny + 2) * (nz + 2);
        c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (c >= total) return;
        const uint32_t c32 = (uint32_t)c, q32 = c32 / (uint32_t)(nz + 2);
        k = (int)(c32 - q32 * (uint32_t)(nz + 2));
        i = (int)(q32 / (uint32_t)(ny + 2));
        j = (int)(q32 - (uint32_t)i * (uint32_t)(ny + 2));
        if (i < 1 || i > nx || j < 1 || j > ny || k < 1 || k > nz) { rho_pad[c] = 0.0; return; }
    }
    const long long m = mesh[((int64_t)(i - 1) * ny + (j - 1)) * nz + (k - 1)];
    if (background_z) background = background_z[k - 1];      // doping profile (emc_set_background_profile)
    double r = background;                                   // poisson_.py:27
    if (scheme == EMC_DEPOSIT_NGP) {
        r = add_repeated(r, increment, m);                   // poisson_.py:36, one += per hit, in sequence
    } else {
        r = background + increment * ((double)m / 4294967296.0);
    }
    rho_pad[c] = r;
}

cudaError_t launch_mesh_to_rho(Ctx *ctx, const int64_t *mesh, int scheme, double *rho_pad, cudaStream_t st,
                               bool interior_only)
{
    const int64_t total = interior_only ? (int64_t)ctx->dc.seg[0] * ctx->dc.seg[1] * ctx->dc.seg[2]
                                        : (int64_t)(ctx->dc.seg[0] + 2) * (ctx->dc.seg[1] + 2) * (ctx->dc.seg[2] + 2);
    ctx->launches++;
    mesh_to_rho_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
        ctx->dc, reinterpret_cast<const long long *>(mesh), scheme, ctx->params.rho_background,
        ctx->d_background_z, ctx->params.rho_increment, rho_pad, interior_only ? 1 : 0);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// The charge all-reduce FUSED with rho (multi-GPU coupled timestep): reduce-scatter of the integer
// counts over peer memory, rho of this rank's 1/world slice of the cells from the totals -- the
// reference's sequential accumulation (add_repeated), which costs more the more particles a cell
// holds and which every rank used to repeat for the WHOLE mesh (0.095 ms of a 1.0 ms timestep at
// 8 GPUs) -- and all-gather of both, totals and rho, with peer stores.  One launch between the same
// two cross-rank barriers as mesh_allreduce_p2p_kernel; same integers, same add_repeated on them:
// rho is bit-identical to emc_mesh_to_rho of the totals on every rank.
// ------------------------------------------------------------------------------------------
struct PeerRho {
    const long long *in[EMC_MAX_PEERS];
    long long *out[EMC_MAX_PEERS];
    double *rho[EMC_MAX_PEERS];              // padded (nx+2, ny+2, nz+2) arrays, halo already zero
};

// one thread per cell: the accumulation of poisson_.py:36 is a serial chain per cell, so the kernel's time is
// one such chain plus the peer round trips -- more threads, not more work per thread
__global__ void __launch_bounds__(128)
mesh_allreduce_rho_p2p_kernel(const __grid_constant__ DevConst P, const __grid_constant__ PeerRho M, int world,
                              int64_t first, int64_t count, int scheme, double background,
                              const double *__restrict__ background_z, double increment)
{
    const uint32_t ny = (uint32_t)P.seg[1], nz = (uint32_t)P.seg[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const int64_t cell = first + t;
        long long m = 0;
        for (int p = 0; p < world; ++p) m += M.in[p][cell];
        for (int p = 0; p < world; ++p) M.out[p][cell] = m;
        const uint32_t c32 = (uint32_t)cell, q32 = c32 / nz;
        const uint32_t k = c32 - q32 * nz, i = q32 / ny, j = q32 - i * ny;
        const int64_t idx = ((int64_t)(i + 1) * (ny + 2) + (j + 1)) * (nz + 2) + (k + 1);
        const double bg = background_z ? background_z[k] : background;     // poisson_.py:27
        const double r = scheme == EMC_DEPOSIT_NGP ? add_repeated(bg, increment, m)   // poisson_.py:36
                                                   : bg + increment * ((double)m / 4294967296.0);
        for (int p = 0; p < world; ++p) M.rho[p][idx] = r;
    }
}

cudaError_t launch_mesh_allreduce_rho_p2p(Ctx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs,
                                          double *const *rho_ptrs, int world, int rank, int64_t n_cells, int scheme,
                                          cudaStream_t st)
{
    PeerRho M;
    for (int p = 0; p < world; ++p) {
        M.in[p] = reinterpret_cast<const long long *>(in_ptrs[p]);
        M.out[p] = reinterpret_cast<long long *>(out_ptrs[p]);
        M.rho[p] = rho_ptrs[p];
    }
    // slices of whole 128-byte lines of the count meshes
    const int64_t per = (((n_cells + world - 1) / world) + 15) & ~(int64_t)15;
    const int64_t first = (int64_t)rank * per;
    const int64_t count = first >= n_cells ? 0 : (first + per > n_cells ? n_cells - first : per);
    if (count == 0) return cudaSuccess;
    int grid = (int)((count + 127) / 128);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    ctx->launches++;
    mesh_allreduce_rho_p2p_kernel<<<grid, 128, 0, st>>>(ctx->dc, M, world, first, count, scheme,
                                                        ctx->params.rho_background, ctx->d_background_z,
                                                        ctx->params.rho_increment);
    return cudaGetLastError();
}

// E = -central difference on interior cells (poisson_.py:88-96), zero halo
__global__ void field_kernel(const __grid_constant__ DevConst P, const double *__restrict__ u, double *ex,
                             double *ey, double *ez)
{
    const int nx = P.seg[0], ny = P.seg[1], nz = P.seg[2];
    const int64_t sy = nz + 2, sx = (int64_t)(ny + 2) * (nz + 2);
    const int64_t total = (int64_t)(nx + 2) * sx;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total) return;
    const uint32_t c32 = (uint32_t)c, q32 = c32 / (uint32_t)(nz + 2);
    const int k = (int)(c32 - q32 * (uint32_t)(nz + 2));
    const int i = (int)(q32 / (uint32_t)(ny + 2));
    const int j = (int)(q32 - (uint32_t)i * (uint32_t)(ny + 2));
    if (i < 1 || i > nx || j < 1 || j > ny || k < 1 || k > nz) {
        ex[c] = 0.0; ey[c] = 0.0; ez[c] = 0.0;
        return;
    }
    ex[c] = -0.5 * (u[c + sx] - u[c - sx]) / P.dl[0];      // :94
    ey[c] = -0.5 * (u[c + sy] - u[c - sy]) / P.dl[1];      // :95
    ez[c] = -0.5 * (u[c + 1] - u[c - 1]) / P.dl[2];        // :96
}

// the same differences, packed (ex, ey, ez, 0) per cell and pre-scaled V/m -> V/cm
__global__ void field_packed_kernel(const __grid_constant__ DevConst P, const double *__restrict__ u, double4 *e4,
                                    int interior_only)
{
    const int nx = P.seg[0], ny = P.seg[1], nz = P.seg[2];
    const int64_t sy = nz + 2, sx = (int64_t)(ny + 2) * (nz + 2);
    int i, j, k;
    int64_t c;
    if (interior_only) {                          // halo records already zero: interior cells only
        const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= (int64_t)nx * ny * nz) return;
        const uint32_t t32 = (uint32_t)t, q32 = t32 / (uint32_t)nz;
        k = 1 + (int)(t32 - q32 * (uint32_t)nz);
        i = 1 + (int)(q32 / (uint32_t)ny);
        j = 1 + (int)(q32 - (uint32_t)(i - 1) * (uint32_t)ny);
        c = (int64_t)i * sx + (int64_t)j * sy + k;
    } else {
        const int64_t total = (int64_t)(nx + 2) * sx;
        c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (c >= total) return;
        const uint32_t c32 = (uint32_t)c, q32 = c32 / (uint32_t)(nz + 2);
        k = (int)(c32 - q32 * (uint32_t)(nz + 2));
        i = (int)(q32 / (uint32_t)(ny + 2));
        j = (int)(q32 - (uint32_t)i * (uint32_t)(ny + 2));
    }
    double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
    if (!(i < 1 || i > nx || j < 1 || j > ny || k < 1 || k > nz)) {
        v.x = (-0.5 * (u[c + sx] - u[c - sx]) / P.dl[0]) * 1e-2;
        v.y = (-0.5 * (u[c + sy] - u[c - sy]) / P.dl[1]) * 1e-2;
        v.z = (-0.5 * (u[c + 1] - u[c - 1]) / P.dl[2]) * 1e-2;
    }
    e4[c] = v;
}

cudaError_t launch_field_packed(Ctx *ctx, const double *u_pad, double *e4, cudaStream_t st, bool interior_only)
{
    const int64_t total = interior_only ? (int64_t)ctx->dc.seg[0] * ctx->dc.seg[1] * ctx->dc.seg[2]
                                        : (int64_t)(ctx->dc.seg[0] + 2) * (ctx->dc.seg[1] + 2) * (ctx->dc.seg[2] + 2);
    ctx->launches++;
    field_packed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->dc, u_pad, reinterpret_cast<double4 *>(e4),
                                                                        interior_only ? 1 : 0);
    return cudaGetLastError();
}

// ny = 1 meshes: (ex, ez) per INTERIOR cell, unpadded (nx, nz), V/cm -- 16 bytes per gather instead of 32,
// a 16 MB array instead of 101 MB allocated / 34 MB touched on 1024 x 1 x 1024 (EMC_FIELD_MESH_XZ)
__global__ void field_xz_kernel(const __grid_constant__ DevConst P, const double *__restrict__ u, double2 *e2)
{
    const int nx = P.seg[0], nz = P.seg[2];
    const int64_t sy = nz + 2, sx = (int64_t)3 * (nz + 2);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nx * nz) return;
    const uint32_t t32 = (uint32_t)t, i0 = t32 / (uint32_t)nz;
    const int k = 1 + (int)(t32 - i0 * (uint32_t)nz), i = 1 + (int)i0;
    const int64_t c = (int64_t)i * sx + sy + k;
    e2[t] = make_double2((-0.5 * (u[c + sx] - u[c - sx]) / P.dl[0]) * 1e-2, (-0.5 * (u[c + 1] - u[c - 1]) / P.dl[2]) * 1e-2);
}

cudaError_t launch_field_xz(Ctx *ctx, const double *u_pad, double *e2, cudaStream_t st)
{
    const int64_t total = (int64_t)ctx->dc.seg[0] * ctx->dc.seg[2];
    ctx->launches++;
    field_xz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->dc, u_pad, reinterpret_cast<double2 *>(e2));
    return cudaGetLastError();
}

cudaError_t launch_field(Ctx *ctx, const double *u_pad, double *ex, double *ey, double *ez, cudaStream_t st)
{
    const int64_t total = (int64_t)(ctx->dc.seg[0] + 2) * (ctx->dc.seg[1] + 2) * (ctx->dc.seg[2] + 2);
    ctx->launches++;
    field_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->dc, u_pad, ex, ey, ez);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Poisson SOR
// ------------------------------------------------------------------------------------------
struct SorConst {
    int nx, ny, nz;
    double xy, xz, e;       // poisson_.py:56-57,61
    double dl0sq;           // dl[0]*dl[0]     (poisson_.py:76)
    double eps_s;
    double pch2_quarter;    // 0.25*(p*p), p = 0.5*sum(cos(pi/n)) (poisson_.py:68,81)
    double omega0, tol;
    int max_sweeps;
};

// one Gauss-Seidel/SOR cell update, poisson_.py:58-60,76-77.  Returns u_new - u_old.
__device__ __forceinline__ double sor_update(const SorConst &S, double *u, const double *rho, int64_t c, int64_t sx,
                                             int64_t sy, double relx, bool &diverged)
{
    const double uc = u[c];
    double sum = S.e * uc;
    sum += u[c + sx];
    sum += u[c - sx];
    sum += S.xy * u[c + sy];
    sum += S.xy * u[c - sy];
    sum += S.xz * u[c + 1];
    sum += S.xz * u[c - 1];
    const double du = (sum - S.dl0sq * rho[c] / S.eps_s) * relx / S.e;
    const double un = uc - du;
    u[c] = un;
    if (fabs(du) >= 1e3) diverged = true;       // poisson_.py:78-80
    return un - uc;
}

__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < nwarp; ++w) t += scratch[w];
        scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// Lexicographic order (poisson_.py:74: k outer, j, i inner) executed as hyperplane wavefronts
// h = i+j+k: a cell's (i-1, j-1, k-1) neighbours lie on plane h-1 (already updated, as in the
// sequential sweep) and its (+1) neighbours on plane h+1 (still old), and cells of one plane
// are mutually independent -- so every iterate is bit-identical to the sequential loop.
// One CTA; the whole solve (all sweeps) is one launch.  SMEM: u and rho live in shared memory.
template <bool SMEM>
__global__ void __launch_bounds__(1024)
sor_lex_kernel(const __grid_constant__ SorConst S, double *u_g, const double *rho_g, double *err_hist,
               double *scalars, int32_t *flags)
{
    extern __shared__ double sm[];
    __shared__ double scratch[33];
    __shared__ int s_div;
    const int nx = S.nx, ny = S.ny, nz = S.nz;
    const int64_t sy = nz + 2, sx = (int64_t)(ny + 2) * (nz + 2);
    const int64_t total = (int64_t)(nx + 2) * sx;
    double *u = u_g;
    const double *rho = rho_g;
    if (SMEM) {
        double *su = sm, *srho = sm + total;
        for (int64_t t = threadIdx.x; t < total; t += blockDim.x) { su[t] = u_g[t]; srho[t] = rho_g[t]; }
        u = su;
        rho = srho;
    }
    if (threadIdx.x == 0) s_div = 0;
    __syncthreads();
    double relx = S.omega0;          // poisson_.py:66
    double err = 1;
    int num = 0;
    bool diverged = false;
    while (err > S.tol && num < S.max_sweeps) {      // poisson_.py:69
        double acc = 0.0;
        for (int h = 3; h <= nx + ny + nz; ++h) {
            const int ilo = max(1, h - ny - nz), ihi = min(nx, h - 2);
            const int pairs = (ihi - ilo + 1) * ny;
            for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
                const int i = ilo + p / ny, j = 1 + p % ny, k = h - i - j;
                if (k >= 1 && k <= nz) {
                    const double d = sor_update(S, u, rho, i * sx + j * sy + k, sx, sy, relx, diverged);
                    acc += d * d;
                }
            }
            __syncthreads();
        }
        relx = 1 / (1 - S.pch2_quarter * relx);      // poisson_.py:81
        const double ssq = block_sum(acc, scratch);
        err = sqrt(ssq / (double)total);             // poisson_.py:82 (RMS over the padded array)
        if (threadIdx.x == 0) err_hist[num] = err;
        num += 1;
        if (diverged) s_div = 1;
        __syncthreads();
        if (s_div) break;
    }
    if (SMEM) {
        for (int64_t t = threadIdx.x; t < total; t += blockDim.x) u_g[t] = u[t];
    }
    if (threadIdx.x == 0) {
        scalars[0] = err;
        scalars[1] = relx;
        flags[0] = num;
        flags[1] = s_div;
        if (s_div) flags[2] = 1;
    }
}

// Red-black ordering: colour = (i+j+k) & 1.  Cooperative persistent grid, two grid syncs per
// sweep; omega is updated once per sweep and the stopping rule is the reference's RMS update.
__global__ void __launch_bounds__(256)
sor_rb_kernel(const __grid_constant__ SorConst S, double *u, const double *__restrict__ rho, double *err_hist,
              double *scalars, int32_t *flags, double *partials)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double scratch[33];
    const int nx = S.nx, ny = S.ny, nz = S.nz;
    const int64_t sy = nz + 2, sx = (int64_t)(ny + 2) * (nz + 2);
    const int64_t total = (int64_t)(nx + 2) * sx;
    const int hz = (nz + 1) >> 1;                       // k slots per (i,j) column and colour
    const int64_t work = (int64_t)nx * ny * hz;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double relx = S.omega0;
    double err = 1;
    int num = 0;
    bool diverged = false;
    int any_div = 0;
    while (err > S.tol && num < S.max_sweeps) {
        double acc = 0.0;
        for (int colour = 0; colour < 2; ++colour) {
            for (int64_t t = gtid; t < work; t += gstride) {
                const uint32_t t32 = (uint32_t)t, q32 = t32 / (uint32_t)hz;
                const int kk = (int)(t32 - q32 * (uint32_t)hz);
                const int i0 = (int)(q32 / (uint32_t)ny);
                const int j = 1 + (int)(q32 - (uint32_t)i0 * (uint32_t)ny);
                const int i = 1 + i0;
                const int k = 1 + 2 * kk + (((i + j + 1) & 1) != colour ? 1 : 0);
                if (k <= nz) {
                    const double d = sor_update(S, u, rho, i * sx + j * sy + k, sx, sy, relx, diverged);
                    acc += d * d;
                }
            }
            if (colour == 0) grid.sync();
        }
        relx = 1 / (1 - S.pch2_quarter * relx);
        const double bs = block_sum(acc, scratch);
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = bs;
            if (diverged) { atomicExch(&flags[1], 1); atomicExch(&flags[2], 1); }
        }
        grid.sync();
        // every block folds the partials in the same fixed order -> identical err everywhere
        double v = 0.0;
        if (threadIdx.x < 32) {
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) v += __ldcg(&partials[b]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (threadIdx.x == 0) {
                scratch[32] = v;
                scratch[31] = (double)__ldcg(&flags[1]);
            }
        }
        __syncthreads();
        err = sqrt(scratch[32] / (double)total);
        any_div = scratch[31] != 0.0;
        __syncthreads();
        if (gtid == 0) err_hist[num] = err;
        num += 1;
        if (any_div) break;
    }
    if (gtid == 0) {
        scalars[0] = err;
        scalars[1] = relx;
        flags[0] = num;
    }
}

int run_poisson_sor(Ctx *ctx, double *u_pad, const double *rho_pad, double omega0, double tol, int max_sweeps,
                    int order, int32_t *sweeps_out, double *err_out, double *err_hist, cudaStream_t st)
{
    const int nx = ctx->dc.seg[0], ny = ctx->dc.seg[1], nz = ctx->dc.seg[2];
    const double *dl = ctx->dc.dl;
    SorConst S;
    S.nx = nx; S.ny = ny; S.nz = nz;
    S.xy = (dl[0] / dl[1]) * (dl[0] / dl[1]);            // poisson_.py:56 (x/y)**2
    S.xz = (dl[0] / dl[2]) * (dl[0] / dl[2]);
    S.e = -2 * (1 + S.xy + S.xz);
    S.dl0sq = dl[0] * dl[0];
    S.eps_s = ctx->params.eps_s;
    const double pch = 0.5 * std::cos(M_PI / nx) + 0.5 * std::cos(M_PI / ny) + 0.5 * std::cos(M_PI / nz);
    S.pch2_quarter = 0.25 * (pch * pch);
    S.omega0 = omega0; S.tol = tol; S.max_sweeps = max_sweeps;
    if (max_sweeps < 1) { ctx->last_error = "max_sweeps < 1"; return EMC_ERR_INVALID; }

    cudaError_t e;
    if (ctx->err_hist_cap < max_sweeps) {
        if (ctx->d_err_hist) cudaFree(ctx->d_err_hist);
        ctx->d_err_hist = nullptr;
        if ((e = cudaMalloc(&ctx->d_err_hist, sizeof(double) * (size_t)max_sweeps)) != cudaSuccess) goto cuda_fail;
        ctx->err_hist_cap = max_sweeps;
    }
    // flags: [0] sweeps, [1] diverged (this solve), [2] diverged (sticky until emc_poisson_last reads it)
    if ((e = cudaMemsetAsync(ctx->d_sor_flags, 0, 2 * sizeof(int32_t), st)) != cudaSuccess) goto cuda_fail;
    {
        const int64_t total = (int64_t)(nx + 2) * (ny + 2) * (nz + 2);
        ctx->launches++;
        if (order == EMC_SOR_LEXICOGRAPHIC) {
            const size_t need = (size_t)total * 2 * sizeof(double);
            if (need <= 200 * 1024) {
                if ((e = cudaFuncSetAttribute(sor_lex_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)need)) != cudaSuccess) goto cuda_fail;
                sor_lex_kernel<true><<<1, 1024, need, st>>>(S, u_pad, rho_pad, ctx->d_err_hist, ctx->d_sor_scalars,
                                                           ctx->d_sor_flags);
            } else {
                sor_lex_kernel<false><<<1, 1024, 0, st>>>(S, u_pad, rho_pad, ctx->d_err_hist, ctx->d_sor_scalars,
                                                          ctx->d_sor_flags);
            }
            e = cudaGetLastError();
        } else if (order == EMC_SOR_RED_BLACK) {
            int per_sm = 0;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sor_rb_kernel, 256, 0)) != cudaSuccess)
                goto cuda_fail;
            {   // persistent blocks per SM of the red-black solver; EMC_SOR_BLOCKS_PER_SM tunes it
                const int occ = per_sm;
                if (per_sm > 4) per_sm = 4;
                const char *env = std::getenv("EMC_SOR_BLOCKS_PER_SM");
                if (env && std::atoi(env) > 0) per_sm = std::atoi(env) < occ ? std::atoi(env) : occ;
            }
            const int64_t work = (int64_t)nx * ny * ((nz + 1) / 2);
            int grid = ctx->sm_count * per_sm;
            const int64_t want = (work + 255) / 256;
            if (want < grid) grid = (int)want;
            if (grid < 1) grid = 1;
            if (grid > 4096) grid = 4096;        // d_sor_partials capacity
            double *eh = ctx->d_err_hist, *sc = ctx->d_sor_scalars, *pt = ctx->d_sor_partials;
            int32_t *fl = ctx->d_sor_flags;
            void *args[] = {(void *)&S, (void *)&u_pad, (void *)&rho_pad, (void *)&eh, (void *)&sc, (void *)&fl,
                            (void *)&pt};
            e = cudaLaunchCooperativeKernel((void *)sor_rb_kernel, dim3(grid), dim3(256), args, 0, st);
        } else if (order == EMC_SOR_MULTIGRID) {
            // tol keeps the meaning of poisson_.py:63,82 -- the RMS of one Gauss-Seidel UPDATE -- but is
            // evaluated on the true residual: an update is residual / |e| (poisson_.py:76)
            ctx->launches--;
            const double ae = std::fabs(S.e);
            int32_t cyc = 0;
            double res = 0.0;
            const bool want = sweeps_out || err_out || err_hist;
            const int r = run_poisson_mg(ctx, u_pad, rho_pad, tol * ae, max_sweeps, 2, 2, want ? &cyc : nullptr,
                                         want ? &res : nullptr, err_hist, st);
            if (r != EMC_OK) return r;
            if (sweeps_out) *sweeps_out = cyc;
            if (err_out) *err_out = res / ae;
            if (err_hist) for (int k = 0; k < cyc; ++k) err_hist[k] /= ae;
            return EMC_OK;
        } else {
            ctx->last_error = "unknown SOR order";
            return EMC_ERR_INVALID;
        }
        if (e != cudaSuccess) goto cuda_fail;
    }
    if (sweeps_out || err_out || err_hist) {      // otherwise asynchronous: query with emc_poisson_last
        double scal[2];
        int32_t fl[2];
        if ((e = cudaMemcpyAsync(scal, ctx->d_sor_scalars, sizeof(scal), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            goto cuda_fail;
        if ((e = cudaMemcpyAsync(fl, ctx->d_sor_flags, sizeof(fl), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            goto cuda_fail;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) goto cuda_fail;
        if (sweeps_out) *sweeps_out = fl[0];
        if (err_out) *err_out = scal[0];
        if (err_hist && fl[0] > 0) {
            if ((e = cudaMemcpy(err_hist, ctx->d_err_hist, sizeof(double) * (size_t)fl[0], cudaMemcpyDeviceToHost)) !=
                cudaSuccess)
                goto cuda_fail;
        }
        if (fl[1]) {
            ctx->last_error = "Too large du (poisson_.py:78-80)";
            return EMC_ERR_DIVERGED;
        }
    }
    return EMC_OK;
cuda_fail:
    ctx->last_error = std::string("CUDA: ") + cudaGetErrorString(e);
    return EMC_ERR_CUDA;
}

int poisson_last(Ctx *ctx, int32_t *sweeps_out, double *err_out, cudaStream_t st)
{
    double scal[2];
    int32_t fl[3];
    cudaError_t e;
    if ((e = cudaMemcpyAsync(scal, ctx->d_sor_scalars, sizeof(scal), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(fl, ctx->d_sor_flags, sizeof(fl), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemsetAsync(ctx->d_sor_flags + 2, 0, sizeof(int32_t), st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess) {
        ctx->last_error = std::string("CUDA: ") + cudaGetErrorString(e);
        return EMC_ERR_CUDA;
    }
    if (sweeps_out) *sweeps_out = fl[0];
    if (err_out) *err_out = scal[0];
    if (fl[1] || fl[2]) {       // this solve, or any enqueued-only solve since the last query
        ctx->last_error = "Too large du (poisson_.py:78-80)";
        return EMC_ERR_DIVERGED;
    }
    return EMC_OK;
}

}  // namespace emc
