// emc_abi.cu -- the extern "C" boundary declared in include/emc.h.
// Argument validation, context lifetime, table upload; every compute entry point is a thin
// wrapper over a kernel launcher in emc_flight.cu / emc_mesh.cu.  No exception crosses the ABI.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "emc_host.h"

using namespace emc;

struct EmcCtx {
    Ctx c;
};

static std::string g_create_error;

#define CTX_OR_FAIL(ctx)                                       \
    if (!(ctx)) return EMC_ERR_INVALID;                        \
    Ctx *c = &(ctx)->c;                                        \
    {                                                          \
        cudaError_t e__ = cudaSetDevice(c->device);            \
        if (e__ != cudaSuccess) return fail_cuda(c, e__, "cudaSetDevice"); \
    }

// +0 exactly (not -0, not tiny): what carrier_drift<ZONLY> may drop
static bool is_plus_zero(double v) { return v == 0.0 && !std::signbit(v); }

static void update_zonly_constant(Ctx *c)
{
    const char *env = std::getenv("EMC_ZONLY");
    const bool allow = !(env && std::atoi(env) == 0);
    c->zonly_constant = allow && is_plus_zero(c->dc.efield[0]) && is_plus_zero(c->dc.efield[1]);
}

static int fail(Ctx *c, int code, const char *msg)
{
    c->last_error = msg;
    return code;
}

static int fail_cuda(Ctx *c, cudaError_t e, const char *where)
{
    c->last_error = std::string(where) + ": " + cudaGetErrorString(e);
    return EMC_ERR_CUDA;
}

// init_.where_am_i (init_.py:36-56) on the host, the reference's own arithmetic
static int where_am_i_host(const EmcLayer *layers, int n, double dist)
{
    if (dist < 0) return -1;
    for (int l = 0; l < n; ++l) {
        if (dist <= layers[l].lz) return l;
        dist -= layers[l].lz;
    }
    return -1;
}

// largest double d >= 0 with where_am_i(d) in [0, l]: the loop is monotone in d, so bisection over
// the (ordered) bit patterns of the non-negative doubles finds it exactly
static double layer_threshold(const EmcLayer *layers, int n, int l)
{
    auto inside = [&](double d) { const int w = where_am_i_host(layers, n, d); return w >= 0 && w <= l; };
    uint64_t lo = 0, hi = 0x7ff0000000000000ULL;        // +0 is inside (layer 0), +inf is not
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        double d;
        std::memcpy(&d, &mid, sizeof(d));
        if (inside(d)) lo = mid; else hi = mid;
    }
    double d;
    std::memcpy(&d, &lo, sizeof(d));
    return d;
}

// single-precision constants of select_self_thr's index guess (needs the energy grid: called again at table upload)
static void fill_guess(Ctx *c)
{
    DevConst &d = c->dc;
    for (int l = 0; l < d.n_layers; ++l) {
        Mat &m = d.mat[l];
        for (int v = 0; v < 3; ++v) {
            m.g_c1[v] = (float)(d.hbar2 / m.den[v]);
            m.g_4a[v] = (float)m.four_a[v];
            m.g_k[v] = d.ek_step > 0 ? (float)(1.0 / (m.two_a[v] * d.ek_step)) : 0.0f;
        }
    }
}

static void fill_devconst(Ctx *c)
{
    const EmcParams &p = c->params;
    DevConst &d = c->dc;
    d.dt = p.dt;
    d.neg_inv_gamma = -1 / p.gamma;             // main_.py:122 (-1/G0)
    d.hbar = p.hbar;
    d.q = p.q;
    d.half_q = 0.5 * p.q;
    d.hbar2 = p.hbar * p.hbar;
    for (int l = 0; l < c->dc.n_layers; ++l) {
        Mat &m = d.mat[l];
        const EmcLayer &L = c->layers[l];
        for (int v = 0; v < 3; ++v) {
            m.mass[v] = L.mass[v];
            m.den[v] = 2 * L.mass[v] * p.q;     // (2*m)*q
            m.four_a[v] = 4 * L.nonparab[v];
            m.two_a[v] = 2 * L.nonparab[v];
            m.r_den[v] = 1 / m.den[v];
            m.r_two_a[v] = 1 / m.two_a[v];
            m.r_mass[v] = 1 / L.mass[v];
        }
        m.e_phonon = L.e_phonon;
        m.ion_eb = L.ion_eb;
        d.layer_lz[l] = L.lz;
        d.layer_thr[l] = layer_threshold(c->layers, c->dc.n_layers, l);
    }
    {   // cumulative shares of the initial ensemble (all weights 0: uniform z like init_.py:200)
        double tot = 0.0, run = 0.0;
        for (int l = 0; l < c->dc.n_layers; ++l) tot += c->layers[l].init_weight;
        for (int l = 0; l < EMC_MAX_LAYERS; ++l) d.layer_cw[l] = 0.0;
        if (tot > 0)
            for (int l = 0; l < c->dc.n_layers; ++l) {
                run += c->layers[l].init_weight;
                d.layer_cw[l] = (l == c->dc.n_layers - 1) ? 1.0 : run / tot;
            }
    }
    for (int v = 0; v < 3; ++v) {
        d.inv_dl[v] = 1 / p.dl[v];
        d.tot_dim[v] = p.tot_dim[v];
        d.efield[v] = p.efield[v];
        d.seg[v] = p.seg[v];
        d.dl[v] = p.dl[v];
    }
    d.r_hbar = 1 / p.hbar;
    d.r_hbar2 = 1 / d.hbar2;
    d.z_cyclic = p.z_cyclic != 0;
    fill_guess(c);
}

// the single layer EmcParams describes (init_.py:127: layers = [layer0])
static void default_layer(Ctx *c)
{
    const EmcParams &p = c->params;
    EmcLayer &L = c->layers[0];
    std::memset(&L, 0, sizeof(L));
    L.lz = p.tot_dim[2];
    for (int v = 0; v < 3; ++v) { L.mass[v] = p.mass[v]; L.nonparab[v] = p.nonparab[v]; }
    L.e_phonon = p.e_phonon;
    L.ion_eb = p.ion_eb;
    c->dc.n_layers = 1;
}

extern "C" int emc_abi_version(void) { return EMC_ABI_VERSION; }

extern "C" int emc_create(const EmcParams *params, int device, EmcCtx **out)
{
    if (!params || !out) { g_create_error = "null argument"; return EMC_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return EMC_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return EMC_ERR_INVALID; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return EMC_ERR_CUDA;
    }
    for (int a = 0; a < 3; ++a)
        if (params->seg[a] < 1 || !(params->dl[a] > 0) || !(params->tot_dim[a] > 0) || !(params->mass[a] > 0)) {
            g_create_error = "seg/dl/tot_dim/mass must be positive";
            return EMC_ERR_INVALID;
        }
    if (!(params->gamma > 0) || !(params->dt > 0)) { g_create_error = "gamma and dt must be positive"; return EMC_ERR_INVALID; }
    if ((double)(params->seg[0] + 2) * (double)(params->seg[1] + 2) * (double)(params->seg[2] + 2) >= 2147483648.0) {
        g_create_error = "mesh too large: fewer than 2^31 padded cells are supported";
        return EMC_ERR_UNSUPPORTED;
    }
    EmcCtx *h = new (std::nothrow) EmcCtx();
    if (!h) { g_create_error = "out of host memory"; return EMC_ERR_INVALID; }
    Ctx *c = &h->c;
    c->params = *params;
    c->device = device;
    std::memset(&c->dc, 0, sizeof(c->dc));
    std::memset(&c->mech, 0, sizeof(c->mech));
    std::memset(c->mech_layers, 0, sizeof(c->mech_layers));
    default_layer(c);
    fill_devconst(c);
    update_zonly_constant(c);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) goto fail_cuda_create;
    c->sm_count = prop.multiProcessorCount;
    {
        int per_sm = 8;
        const char *wv = std::getenv("EMC_WAVES");
        if (wv && std::atoi(wv) > 0) c->waves = std::atoi(wv);
        const char *env = std::getenv("EMC_BLOCKS_PER_SM");
        if (env && std::atoi(env) > 0) per_sm = std::atoi(env);
        c->max_grid = c->sm_count * per_sm;
        const char *fa = std::getenv("EMC_FIELD_AHEAD");
        c->no_field_ahead = fa && std::atoi(fa) == 0;
        const char *fs = std::getenv("EMC_FUSED_STEPS");
        c->no_fused_steps = fs && std::atoi(fs) == 0;
        const char *cs = std::getenv("EMC_CHAIN_STEPS");
        c->no_chain = cs && std::atoi(cs) == 0;
        const char *tm = std::getenv("EMC_REFILL_TENSOR_MAP");
        c->no_tensor_map = tm && std::atoi(tm) == 0;
        const char *ln = std::getenv("EMC_LEAN");
        c->no_lean = ln && std::atoi(ln) == 0;
        const char *br = std::getenv("EMC_BULK_REFILL");
        c->no_bulk_refill = br && std::atoi(br) == 0;
        c->bulk_mesh = br && std::atoi(br) == 2;
        const char *ex = std::getenv("EMC_EXACT_LIBM");
        c->exact_libm = ex && std::atoi(ex) != 0;
    }
    if ((e = cudaMalloc(&c->partials, sizeof(ObsPartial) * (size_t)c->max_grid)) != cudaSuccess) goto fail_cuda_create;
    if ((e = cudaMalloc(&c->ticket, sizeof(unsigned int))) != cudaSuccess) goto fail_cuda_create;
    if ((e = cudaMemset(c->ticket, 0, sizeof(unsigned int))) != cudaSuccess) goto fail_cuda_create;
    if ((e = cudaMalloc(&c->d_obs, sizeof(EmcObs))) != cudaSuccess) goto fail_cuda_create;
    if ((e = cudaMalloc(&c->d_sor_scalars, sizeof(double) * 8)) != cudaSuccess) goto fail_cuda_create;
    if ((e = cudaMalloc(&c->d_sor_flags, sizeof(int32_t) * 4)) != cudaSuccess) goto fail_cuda_create;
    if ((e = cudaMemset(c->d_sor_flags, 0, sizeof(int32_t) * 4)) != cudaSuccess) goto fail_cuda_create;
    if ((e = cudaMalloc(&c->d_sor_partials, sizeof(double) * 4096)) != cudaSuccess) goto fail_cuda_create;
    *out = h;
    return EMC_OK;
fail_cuda_create:
    g_create_error = std::string("CUDA: ") + cudaGetErrorString(e);
    emc_destroy(h);
    return EMC_ERR_CUDA;
}

extern "C" void emc_destroy(EmcCtx *ctx)
{
    if (!ctx) return;
    Ctx *c = &ctx->c;
    cudaSetDevice(c->device);
    for (int l = 0; l < EMC_MAX_LAYERS; ++l)
        for (int v = 0; v < 3; ++v) {
            if (c->d_cum_layers[l][v]) cudaFree(c->d_cum_layers[l][v]);
            if (c->d_thr_layers[l][v]) cudaFree(c->d_thr_layers[l][v]);
        }
    if (c->d_mech_layers) cudaFree(c->d_mech_layers);
    if (c->d_background_z) cudaFree(c->d_background_z);
    if (c->d_groups) cudaFree(c->d_groups);
    if (c->partials) cudaFree(c->partials);
    if (c->ticket) cudaFree(c->ticket);
    if (c->d_chain) cudaFree(c->d_chain);
    if (c->chain_partials) cudaFree(c->chain_partials);
    if (c->d_obs) cudaFree(c->d_obs);
    if (c->d_sor_scalars) cudaFree(c->d_sor_scalars);
    if (c->d_sor_flags) cudaFree(c->d_sor_flags);
    if (c->d_sor_partials) cudaFree(c->d_sor_partials);
    if (c->d_err_hist) cudaFree(c->d_err_hist);
    mg_release(c);
    if (c->d_dep_scratch) cudaFree(c->d_dep_scratch);
    delete ctx;
}

extern "C" const char *emc_last_error(const EmcCtx *ctx)
{
    return ctx ? ctx->c.last_error.c_str() : g_create_error.c_str();
}

extern "C" int64_t emc_launch_count(const EmcCtx *ctx) { return ctx ? ctx->c.launches : 0; }

static int upload_tables_layer(Ctx *c, int layer, const double *ek, int32_t n_energy, const double *cum0,
                               const double *cum1, const double *cum2, const int32_t n_mech[3],
                               const double *dE, const int32_t *trans, const int32_t *sampler)
{
    if (layer < 0 || layer >= c->dc.n_layers) return fail(c, EMC_ERR_INVALID, "upload_tables: layer out of range");
    if (!ek || !cum0 || !cum1 || !cum2 || !n_mech || !dE || !trans || !sampler || n_energy < 3)
        return fail(c, EMC_ERR_INVALID, "emc_upload_tables: null argument or n_energy < 3");
    // the kernel regenerates the grid as i*step (np.linspace with start 0): verify that is what we were given
    const double step = ek[1] - ek[0];
    if (!(step > 0) || ek[0] != 0.0) return fail(c, EMC_ERR_UNSUPPORTED, "energy grid must start at 0 and increase");
    for (int32_t i = 0; i < n_energy - 1; ++i)
        if (ek[i] != (double)i * step)
            return fail(c, EMC_ERR_UNSUPPORTED, "energy grid is not np.linspace(0, e_max, n) (init_.py:160-166)");
    const double *cum[3] = {cum0, cum1, cum2};
    for (int v = 0; v < 3; ++v) {
        if (n_mech[v] < 1 || n_mech[v] > EMC_MAX_MECH) return fail(c, EMC_ERR_INVALID, "n_mech out of range");
        for (int j = 0; j < n_mech[v]; ++j) {
            const int t = trans[v * EMC_MAX_MECH + j], s = sampler[v * EMC_MAX_MECH + j];
            if (t < 0 || t > 2 || s < 0 || s > EMC_SAMPLER_ION)
                return fail(c, EMC_ERR_INVALID, "mechanism metadata out of range");
        }
    }
    // all layers share the energy grid and the column counts (one DevConst.n_mech / ek_step)
    for (int l = 0; l < c->dc.n_layers; ++l) {
        if (l == layer || !c->layer_tables[l]) continue;
        if (c->dc.n_energy != n_energy || c->dc.ek_step != step || c->dc.e_max != ek[n_energy - 1] ||
            c->dc.n_mech[0] != n_mech[0] || c->dc.n_mech[1] != n_mech[1] || c->dc.n_mech[2] != n_mech[2])
            return fail(c, EMC_ERR_UNSUPPORTED, "all layers must share the energy grid and the per-valley column counts");
    }
    // fast mechanism selection needs rows of <= 12 columns that are non-decreasing and NaN-free
    // (choose_mech in emc_device.cuh); anything else is scanned literally
    int literal = 0;
    for (int v = 0; v < 3 && !literal; ++v) {
        if (n_mech[v] > 12) literal = 1;
        for (int64_t i = 0; i < n_energy && !literal; ++i) {
            const double *row = cum[v] + i * n_mech[v];
            for (int j = 0; j < n_mech[v]; ++j)
                if (std::isnan(row[j]) || (j > 0 && row[j] < row[j - 1])) { literal = 1; break; }
        }
    }
    bool others = false;
    for (int l = 0; l < c->dc.n_layers; ++l) others = others || (l != layer && c->layer_tables[l]);
    if (others && c->dc.literal_scan != literal)
        return fail(c, EMC_ERR_UNSUPPORTED, "all layers must use the same table layout (padded or literal rows)");
    c->dc.literal_scan = literal;
    for (int v = 0; v < 3; ++v) {
        // device rows: padded to 12 doubles with +inf (32-B aligned, no column-count guards in
        // choose_mech), or the caller's layout when the rows have to be scanned literally
        const int stride = literal ? n_mech[v] : 12;
        c->dc.row_stride[v] = stride;
        std::vector<double> staged;
        const double *src = cum[v];
        if (!literal) {
            staged.assign((size_t)n_energy * 12, HUGE_VAL);
            for (int64_t i = 0; i < n_energy; ++i)
                std::memcpy(&staged[(size_t)i * 12], cum[v] + i * n_mech[v], sizeof(double) * (size_t)n_mech[v]);
            src = staged.data();
        }
        const size_t bytes = sizeof(double) * (size_t)n_energy * (size_t)stride;
        double *&dst = c->d_cum_layers[layer][v];
        if (dst) { cudaFree(dst); dst = nullptr; }
        cudaError_t e = cudaMalloc(&dst, bytes + sizeof(double) * EMC_MAX_MECH);   // + one padding row
        if (e != cudaSuccess) return fail_cuda(c, e, "cudaMalloc(table)");
        e = cudaMemset(dst, 0, bytes + sizeof(double) * EMC_MAX_MECH);
        if (e != cudaSuccess) return fail_cuda(c, e, "cudaMemset(table)");
        e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail_cuda(c, e, "cudaMemcpy(table)");
        c->dc.mat[layer].cum[v] = dst;
        c->dc.n_mech[v] = n_mech[v];
        {   // self-scattering thresholds (Mat::thr, select_self_spec<.., THR>): the second-to-last column of
            // every row and the smallest last column of the valley; one padding entry like the rows
            std::vector<double> thr((size_t)n_energy + 1, 0.0);
            double min_last = HUGE_VAL;
            for (int64_t i = 0; i < n_energy; ++i) {
                const double *row = cum[v] + i * n_mech[v];
                thr[(size_t)i] = n_mech[v] > 1 ? row[n_mech[v] - 2] : -HUGE_VAL;
                if (!(row[n_mech[v] - 1] >= min_last)) min_last = row[n_mech[v] - 1];
            }
            double *&dthr = c->d_thr_layers[layer][v];
            if (dthr) { cudaFree(dthr); dthr = nullptr; }
            e = cudaMalloc(&dthr, sizeof(double) * thr.size());
            if (e != cudaSuccess) return fail_cuda(c, e, "cudaMalloc(thresholds)");
            e = cudaMemcpy(dthr, thr.data(), sizeof(double) * thr.size(), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return fail_cuda(c, e, "cudaMemcpy(thresholds)");
            c->dc.mat[layer].thr[v] = dthr;
            c->dc.mat[layer].min_last[v] = min_last;
        }
        for (int j = 0; j < EMC_MAX_MECH; ++j) {
            c->mech_layers[layer].dE[v][j] = dE[v * EMC_MAX_MECH + j];
            c->mech_layers[layer].trans[v][j] = trans[v * EMC_MAX_MECH + j];
            c->mech_layers[layer].sampler[v][j] = sampler[v * EMC_MAX_MECH + j];
        }
    }
    if (layer == 0) c->mech = c->mech_layers[0];
    if (!c->d_mech_layers) {
        cudaError_t e = cudaMalloc(&c->d_mech_layers, sizeof(c->mech_layers));
        if (e != cudaSuccess) return fail_cuda(c, e, "cudaMalloc(mech_layers)");
    }
    {
        cudaError_t e = cudaMemcpy(c->d_mech_layers, c->mech_layers, sizeof(c->mech_layers), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail_cuda(c, e, "cudaMemcpy(mech_layers)");
    }
    c->dc.n_energy = n_energy;
    c->dc.n_energy_m2 = (double)(n_energy - 2);
    c->dc.ek_step = step;
    c->dc.inv_ek_step = 1 / step;
    c->dc.e_max = ek[n_energy - 1];
    fill_guess(c);
    c->layer_tables[layer] = true;
    {   // the lean flight kernel's assumption (select_self_spec<LEAN>): in every valley of every layer the
        // last column is the ONLY self-scattering mechanism, changes neither energy nor valley
        bool ok = true;
        for (int l = 0; l < c->dc.n_layers && ok; ++l) {
            if (!c->layer_tables[l]) continue;
            const MechMeta &mm = c->mech_layers[l];
            for (int v = 0; v < 3 && ok; ++v)
                for (int j = 0; j < c->dc.n_mech[v]; ++j) {
                    const bool last = j == c->dc.n_mech[v] - 1;
                    if ((mm.sampler[v][j] == EMC_SAMPLER_SELF) != last) ok = false;
                    if (last && (mm.dE[v][j] != 0.0 || mm.trans[v][j] != v)) ok = false;
                }
        }
        c->self_last = ok;
    }
    c->have_tables = true;
    for (int l = 0; l < c->dc.n_layers; ++l) c->have_tables = c->have_tables && c->layer_tables[l];
    return EMC_OK;
}

extern "C" int emc_upload_tables(EmcCtx *ctx, const double *ek, int32_t n_energy, const double *cum0,
                                 const double *cum1, const double *cum2, const int32_t n_mech[3],
                                 const double *dE, const int32_t *trans, const int32_t *sampler)
{
    CTX_OR_FAIL(ctx);
    return upload_tables_layer(c, 0, ek, n_energy, cum0, cum1, cum2, n_mech, dE, trans, sampler);
}

extern "C" int emc_upload_tables_layer(EmcCtx *ctx, int32_t layer, const double *ek, int32_t n_energy,
                                       const double *cum0, const double *cum1, const double *cum2,
                                       const int32_t n_mech[3], const double *dE, const int32_t *trans,
                                       const int32_t *sampler)
{
    CTX_OR_FAIL(ctx);
    return upload_tables_layer(c, layer, ek, n_energy, cum0, cum1, cum2, n_mech, dE, trans, sampler);
}

extern "C" int emc_set_layers(EmcCtx *ctx, const EmcLayer *layers, int32_t n_layers)
{
    CTX_OR_FAIL(ctx);
    if (!layers || n_layers < 1 || n_layers > EMC_MAX_LAYERS)
        return fail(c, EMC_ERR_INVALID, "emc_set_layers: 1..EMC_MAX_LAYERS layers");
    for (int l = 0; l < n_layers; ++l) {
        if (!(layers[l].lz > 0) || !(layers[l].init_weight >= 0))
            return fail(c, EMC_ERR_INVALID, "emc_set_layers: lz must be positive, init_weight non-negative");
        for (int v = 0; v < 3; ++v)
            if (!(layers[l].mass[v] > 0) || !(layers[l].nonparab[v] > 0))
                return fail(c, EMC_ERR_INVALID, "emc_set_layers: mass and nonparab must be positive");
    }
    for (int l = 0; l < n_layers; ++l) c->layers[l] = layers[l];
    c->dc.n_layers = n_layers;
    // tables have to be uploaded again for every layer
    for (int l = 0; l < EMC_MAX_LAYERS; ++l) c->layer_tables[l] = false;
    c->have_tables = false;
    c->batch_layer = 0;
    fill_devconst(c);
    return EMC_OK;
}

extern "C" int emc_set_obs_phase(EmcCtx *ctx, int32_t before_step)
{
    CTX_OR_FAIL(ctx);
    c->obs_before = before_step != 0;
    return EMC_OK;
}

extern "C" int emc_set_exact_libm(EmcCtx *ctx, int32_t on)
{
    CTX_OR_FAIL(ctx);
    c->exact_libm = on != 0;
    return EMC_OK;
}

extern "C" int emc_describe(const EmcCtx *ctx, char *buf, int32_t cap)
{
    if (!ctx || !buf || cap < 1) return EMC_ERR_INVALID;
    const Ctx *c = &ctx->c;
    const char *sor = std::getenv("EMC_SOR_BLOCKS_PER_SM");
    const int n = std::snprintf(buf, (size_t)cap,
                                "exact_libm=%d zonly_constant=%d zonly_groups=%d field_ahead=%d bulk_refill=%d lean=%d waves=%d max_grid=%d "
                                "obs_before=%d n_layers=%d literal_scan=%d sor_blocks_per_sm=%s flight_threads=%d refill_tensor_map=%d chain_steps=%d",
                                (int)c->exact_libm, (int)c->zonly_constant, (int)c->zonly_groups,
                                (int)!c->no_field_ahead, (int)!c->no_bulk_refill, (int)(!c->no_lean && c->self_last), c->waves, c->max_grid, c->obs_before, c->dc.n_layers,
                                c->dc.literal_scan, sor ? sor : "auto", FLIGHT_THREADS, (int)(!c->no_tensor_map && c->tmap_base != nullptr), (int)!c->no_chain);
    return n < cap ? n : cap - 1;
}

extern "C" int emc_select_layer(EmcCtx *ctx, int32_t layer)
{
    CTX_OR_FAIL(ctx);
    if (layer < 0 || layer >= c->dc.n_layers) return fail(c, EMC_ERR_INVALID, "emc_select_layer: layer out of range");
    c->batch_layer = layer;
    return EMC_OK;
}

extern "C" int emc_set_background_profile(EmcCtx *ctx, const double *background_z, int32_t nz)
{
    CTX_OR_FAIL(ctx);
    if (c->d_background_z) { cudaFree(c->d_background_z); c->d_background_z = nullptr; c->background_nz = 0; }
    if (!background_z) return EMC_OK;
    if (nz != c->params.seg[2]) return fail(c, EMC_ERR_INVALID, "emc_set_background_profile: nz must equal seg[2]");
    cudaError_t e = cudaMalloc(&c->d_background_z, sizeof(double) * (size_t)nz);
    if (e != cudaSuccess) return fail_cuda(c, e, "cudaMalloc(background)");
    e = cudaMemcpy(c->d_background_z, background_z, sizeof(double) * (size_t)nz, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail_cuda(c, e, "cudaMemcpy(background)");
    c->background_nz = nz;
    return EMC_OK;
}

extern "C" int emc_set_efield(EmcCtx *ctx, const double efield[3])
{
    CTX_OR_FAIL(ctx);
    if (!efield) return fail(c, EMC_ERR_INVALID, "emc_set_efield: null");
    for (int a = 0; a < 3; ++a) c->params.efield[a] = c->dc.efield[a] = efield[a];
    update_zonly_constant(c);
    return EMC_OK;
}

extern "C" int emc_set_field_groups(EmcCtx *ctx, const double *efields, int32_t n_groups, int64_t group_size)
{
    CTX_OR_FAIL(ctx);
    if (!efields || n_groups < 1 || group_size < 1) return fail(c, EMC_ERR_INVALID, "emc_set_field_groups: bad argument");
    if (group_size >= (int64_t)1 << 31) return fail(c, EMC_ERR_UNSUPPORTED, "emc_set_field_groups: group_size must be < 2^31");
    if (c->d_groups) { cudaFree(c->d_groups); c->d_groups = nullptr; }
    cudaError_t e = cudaMalloc(&c->d_groups, sizeof(double) * 3 * (size_t)n_groups);
    if (e != cudaSuccess) return fail_cuda(c, e, "cudaMalloc(groups)");
    e = cudaMemcpy(c->d_groups, efields, sizeof(double) * 3 * (size_t)n_groups, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail_cuda(c, e, "cudaMemcpy(groups)");
    c->dc.field_groups = c->d_groups;
    c->dc.n_groups = n_groups;
    c->dc.group_size = group_size;
    {
        const char *env = std::getenv("EMC_ZONLY");
        bool z = !(env && std::atoi(env) == 0);
        for (int g = 0; g < n_groups && z; ++g) z = is_plus_zero(efields[3 * g]) && is_plus_zero(efields[3 * g + 1]);
        c->zonly_groups = z;
    }
    return EMC_OK;
}

static int check_state(Ctx *c, const void *kx, const void *ky, const void *kz, const void *x, const void *y,
                       const void *z, const void *dtau, const void *valley, int64_t n)
{
    if (n < 0) return fail(c, EMC_ERR_INVALID, "n < 0");
    if (n > 0 && (!kx || !ky || !kz || !x || !y || !z || !dtau || !valley))
        return fail(c, EMC_ERR_INVALID, "null particle-state pointer");
    return EMC_OK;
}

extern "C" int emc_init_ensemble(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                 double *dtau, int32_t *valley, int64_t n, uint64_t seed, uint64_t pid_offset,
                                 void *stream)
{
    CTX_OR_FAIL(ctx);
    int r = check_state(c, kx, ky, kz, x, y, z, dtau, valley, n);
    if (r) return r;
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_init_ensemble(c, kx, ky, kz, x, y, z, dtau, valley, n, seed, pid_offset,
                                         (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "init_ensemble launch");
}

extern "C" int emc_init_photoex(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                double *dtau, int32_t *valley, int64_t capacity, int64_t n_candidates,
                                double energy, double z_scale, double xy_std, double gamma_first, uint64_t seed,
                                uint64_t candidate_offset, int64_t *n_accepted, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!n_accepted) return fail(c, EMC_ERR_INVALID, "emc_init_photoex: n_accepted is null");
    *n_accepted = 0;
    int r = check_state(c, kx, ky, kz, x, y, z, dtau, valley, n_candidates);
    if (r) return r;
    if (n_candidates == 0) return EMC_OK;
    if (capacity < n_candidates) return fail(c, EMC_ERR_INVALID, "emc_init_photoex: capacity < n_candidates");
    if (!(energy > 0) || !(z_scale > 0) || !(xy_std >= 0) || !(gamma_first > 0))
        return fail(c, EMC_ERR_INVALID, "emc_init_photoex: energy, z_scale, gamma_first must be positive");
    cudaError_t e = launch_init_photoex(c, kx, ky, kz, x, y, z, dtau, valley, n_candidates, energy, z_scale, xy_std,
                                        gamma_first, seed, candidate_offset, n_accepted, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "init_photoex");
}

static int flight_common(Ctx *c, FlightArgs &A, bool replay, int32_t field_mode, const double *ex, const double *ey,
                         const double *ez, void *stream)
{
    if (!c->have_tables) return fail(c, EMC_ERR_NO_TABLES, "emc_upload_tables has not been called");
    if (field_mode == EMC_FIELD_MESH && (!ex || !ey || !ez)) return fail(c, EMC_ERR_INVALID, "EMC_FIELD_MESH needs ex, ey, ez");
    if (field_mode == EMC_FIELD_GROUPS && !c->dc.field_groups)
        return fail(c, EMC_ERR_INVALID, "EMC_FIELD_GROUPS needs emc_set_field_groups");
    if (field_mode == EMC_FIELD_MESH_PACKED && (!ex || ((uintptr_t)ex & 31) != 0))
        return fail(c, EMC_ERR_INVALID, "EMC_FIELD_MESH_PACKED needs a 32-byte aligned packed field in ex");
    if (field_mode == EMC_FIELD_MESH_XZ && (!ex || ((uintptr_t)ex & 15) != 0 || c->dc.seg[1] != 1))
        return fail(c, EMC_ERR_INVALID, "EMC_FIELD_MESH_XZ needs ny = 1 and a 16-byte aligned (ex, ez) record array in ex");
    if (field_mode < 0 || field_mode > EMC_FIELD_MESH_XZ) return fail(c, EMC_ERR_INVALID, "unknown field_mode");
    if (A.n >= ((int64_t)1 << 32) - 64)
        return fail(c, EMC_ERR_UNSUPPORTED, "at most 2^32 - 65 particles per call: slice the ensemble (pid_offset keeps the random streams global)");
    if (c->dc.n_layers > 1 && A.mesh)
        return fail(c, EMC_ERR_UNSUPPORTED, "multi-layer devices: no fused deposit (use emc_deposit)");
    A.ex = ex; A.ey = ey; A.ez = ez;
    A.obs_before = c->obs_before;
    c->dc.zonly = (field_mode == EMC_FIELD_CONSTANT && c->zonly_constant) ||
                  (field_mode == EMC_FIELD_GROUPS && c->zonly_groups);
    if (A.n == 0) return EMC_OK;
    cudaError_t e = launch_flight(c, A, replay, field_mode, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "flight_scatter launch");
}

extern "C" int emc_flight_scatter(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                  double *dtau, int32_t *valley, int64_t n, int32_t field_mode, const double *ex,
                                  const double *ey, const double *ez, uint64_t seed, uint64_t pid_offset,
                                  uint32_t step, EmcObs *obs, int64_t *mesh, void *stream)
{
    CTX_OR_FAIL(ctx);
    int r = check_state(c, kx, ky, kz, x, y, z, dtau, valley, n);
    if (r) return r;
    FlightArgs A;
    std::memset(&A, 0, sizeof(A));
    A.kx = kx; A.ky = ky; A.kz = kz; A.x = x; A.y = y; A.z = z; A.dtau = dtau; A.valley = valley; A.n = n;
    A.seed = seed; A.pid_offset = pid_offset; A.step = step;
    A.obs = obs; A.mesh = mesh;
    return flight_common(c, A, false, field_mode, ex, ey, ez, stream);
}

extern "C" int emc_flight_scatter_steps(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y,
                                        double *z, double *dtau, int32_t *valley, int64_t n, int32_t field_mode,
                                        const double *ex, const double *ey, const double *ez, uint64_t seed,
                                        uint64_t pid_offset, uint32_t first_step, int32_t n_steps, EmcObs *obs_rows,
                                        void *stream)
{
    CTX_OR_FAIL(ctx);
    if (n_steps < 0) return fail(c, EMC_ERR_INVALID, "emc_flight_scatter_steps: n_steps < 0");
    int r = check_state(c, kx, ky, kz, x, y, z, dtau, valley, n);
    if (r) return r;
    // no per-step observables wanted and the lean conditions hold: all n_steps timesteps in ONE
    // persistent launch (flight_steps_kernel), bit-identical to the loop below
    c->dc.zonly = (field_mode == EMC_FIELD_CONSTANT && c->zonly_constant) || (field_mode == EMC_FIELD_GROUPS && c->zonly_groups);
    if (!obs_rows && n_steps >= 2 && n_steps <= 65535 && n > 0 && !c->no_fused_steps && c->have_tables &&
        (field_mode == EMC_FIELD_CONSTANT || (field_mode == EMC_FIELD_GROUPS && c->dc.field_groups)) && c->dc.zonly &&
        c->dc.n_layers == 1 && !c->exact_libm && !c->dc.literal_scan && c->self_last && n < ((int64_t)1 << 32) - 64) {
        FlightArgs A;
        std::memset(&A, 0, sizeof(A));
        A.kx = kx; A.ky = ky; A.kz = kz; A.x = x; A.y = y; A.z = z; A.dtau = dtau; A.valley = valley; A.n = n;
        A.seed = seed; A.pid_offset = pid_offset; A.step = first_step;
        cudaError_t e = launch_flight_steps(c, A, field_mode, n_steps, (cudaStream_t)stream);
        return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "flight_steps launch");
    }
    int32_t s = 0;
    // per-step observable rows in the before-step phase and the lean conditions hold: timestep chains, i.e. up to
    // CHAIN_MAX_STEPS timesteps with their rows per persistent launch (launch_flight_chain), bit-identical -- state
    // and rows -- to the loop below
    while (obs_rows && c->obs_before && c->have_tables && n > 0 && n_steps - s >= 2) {
        const int32_t k = n_steps - s < CHAIN_MAX_STEPS ? n_steps - s : CHAIN_MAX_STEPS;
        FlightArgs A;
        std::memset(&A, 0, sizeof(A));
        A.kx = kx; A.ky = ky; A.kz = kz; A.x = x; A.y = y; A.z = z; A.dtau = dtau; A.valley = valley; A.n = n;
        A.seed = seed; A.pid_offset = pid_offset; A.step = first_step + (uint32_t)s;
        A.obs = obs_rows + s;
        const cudaError_t e = launch_flight_chain(c, A, field_mode, k, (cudaStream_t)stream);
        if (e == cudaErrorNotSupported) break;              // nothing enqueued: one launch per timestep
        if (e != cudaSuccess) return fail_cuda(c, e, "flight chain launch");
        s += k;
    }
    for (; s < n_steps; ++s) {
        FlightArgs A;
        std::memset(&A, 0, sizeof(A));
        A.kx = kx; A.ky = ky; A.kz = kz; A.x = x; A.y = y; A.z = z; A.dtau = dtau; A.valley = valley; A.n = n;
        A.seed = seed; A.pid_offset = pid_offset; A.step = first_step + (uint32_t)s;
        A.obs = obs_rows ? obs_rows + s : nullptr;
        if ((r = flight_common(c, A, false, field_mode, ex, ey, ez, stream)) != EMC_OK) return r;
    }
    return EMC_OK;
}

extern "C" int emc_flight_scatter_replay(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y,
                                         double *z, double *dtau, int32_t *valley, int64_t n, int32_t field_mode,
                                         const double *ex, const double *ey, const double *ez, uint64_t pid_offset,
                                         const double *uniforms, int64_t stride, int32_t *cursor, EmcObs *obs,
                                         int64_t *mesh, void *stream)
{
    CTX_OR_FAIL(ctx);
    int r = check_state(c, kx, ky, kz, x, y, z, dtau, valley, n);
    if (r) return r;
    if (n > 0 && (!uniforms || !cursor || stride < 1 || stride > 0x7fffffff))
        return fail(c, EMC_ERR_INVALID, "replay needs uniforms, cursor and 1 <= stride < 2^31");
    FlightArgs A;
    std::memset(&A, 0, sizeof(A));
    A.kx = kx; A.ky = ky; A.kz = kz; A.x = x; A.y = y; A.z = z; A.dtau = dtau; A.valley = valley; A.n = n;
    A.uniforms = uniforms; A.stride = stride; A.cursor = cursor; A.pid_offset = pid_offset;
    A.obs = obs; A.mesh = mesh;
    return flight_common(c, A, true, field_mode, ex, ey, ez, stream);
}

extern "C" int emc_flight_steps_check(EmcCtx *ctx, void *stream)
{
    CTX_OR_FAIL(ctx);
    int32_t flag = 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(&flag, c->d_sor_flags + 3, sizeof(flag), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail_cuda(c, e, "flight_steps_check");
    return flag ? fail(c, EMC_ERR_CUDA, "flight_steps_kernel: a warp stopped with particles left in its queues") : EMC_OK;
}

extern "C" int emc_observables(EmcCtx *ctx, const double *kx, const double *ky, const double *kz, const double *z,
                               const int32_t *valley, int64_t n, EmcObs *out, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!out || n < 0 || (n > 0 && (!kx || !ky || !kz || !z || !valley)))
        return fail(c, EMC_ERR_INVALID, "emc_observables: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = launch_observables(c, kx, ky, kz, z, valley, n, c->d_obs, st);
    if (e != cudaSuccess) return fail_cuda(c, e, "observables launch");
    if ((e = cudaMemcpyAsync(out, c->d_obs, sizeof(EmcObs), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        return fail_cuda(c, e, "observables copy");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda(c, e, "observables sync");
    return EMC_OK;
}

extern "C" int emc_observables_async(EmcCtx *ctx, const double *kx, const double *ky, const double *kz,
                                     const double *z, const int32_t *valley, int64_t n, EmcObs *out_dev, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!out_dev || n < 0 || (n > 0 && (!kx || !ky || !kz || !z || !valley)))
        return fail(c, EMC_ERR_INVALID, "emc_observables_async: bad argument");
    cudaError_t e = launch_observables(c, kx, ky, kz, z, valley, n, out_dev, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "observables launch");
}

extern "C" int emc_calc_energy(EmcCtx *ctx, const double *kx, const double *ky, const double *kz,
                               const int32_t *valley, int64_t n, double *e_np, int32_t *idx, int32_t *fault,
                               void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!c->have_tables) return fail(c, EMC_ERR_NO_TABLES, "emc_upload_tables has not been called");
    if (n < 0 || (n > 0 && (!kx || !ky || !kz || !valley || !e_np || !idx || !fault)))
        return fail(c, EMC_ERR_INVALID, "emc_calc_energy: bad argument");
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_calc_energy(c, kx, ky, kz, valley, n, e_np, idx, fault, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "calc_energy launch");
}

extern "C" int emc_free_flight(EmcCtx *ctx, double g0, const double *r, int64_t n, double *out, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (n < 0 || (n > 0 && (!r || !out))) return fail(c, EMC_ERR_INVALID, "emc_free_flight: bad argument");
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_free_flight(c, g0, r, n, out, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "free_flight launch");
}

extern "C" int emc_carrier_drift(EmcCtx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                 const double *mass, const double *efield3, const double *dt2, int64_t n,
                                 void *stream)
{
    CTX_OR_FAIL(ctx);
    if (n < 0 || (n > 0 && (!kx || !ky || !kz || !x || !y || !z || !mass || !efield3 || !dt2)))
        return fail(c, EMC_ERR_INVALID, "emc_carrier_drift: bad argument");
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_carrier_drift(c, kx, ky, kz, x, y, z, mass, efield3, dt2, n, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "carrier_drift launch");
}

extern "C" int emc_scatter(EmcCtx *ctx, double *kx, double *ky, double *kz, int32_t *valley, const double *u3,
                           int64_t n, int32_t *mech, int32_t *n_draws, int32_t *fault, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!c->have_tables) return fail(c, EMC_ERR_NO_TABLES, "emc_upload_tables has not been called");
    if (n < 0 || (n > 0 && (!kx || !ky || !kz || !valley || !u3 || !mech || !n_draws || !fault)))
        return fail(c, EMC_ERR_INVALID, "emc_scatter: bad argument");
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_scatter(c, kx, ky, kz, valley, u3, n, mech, n_draws, fault, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "scatter launch");
}

extern "C" int emc_selftest_math(EmcCtx *ctx, int64_t n, uint64_t seed, int64_t mismatches[3], void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!mismatches || n < 0) return fail(c, EMC_ERR_INVALID, "emc_selftest_math: bad argument");
    unsigned long long *d = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMalloc(&d, 3 * sizeof(unsigned long long));
    if (e != cudaSuccess) return fail_cuda(c, e, "cudaMalloc(selftest)");
    unsigned long long h[3] = {0, 0, 0};
    if ((e = cudaMemsetAsync(d, 0, sizeof(h), st)) == cudaSuccess &&
        (e = launch_selftest_math(c, n, seed, d, st)) == cudaSuccess &&
        (e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st)) == cudaSuccess)
        e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(c, e, "selftest_math");
    for (int k = 0; k < 3; ++k) mismatches[k] = (int64_t)h[k];
    return EMC_OK;
}

extern "C" int emc_selftest_accumulate(EmcCtx *ctx, int64_t n, uint64_t seed, int32_t max_count, int64_t *mismatches,
                                       void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!mismatches || n < 0 || max_count < 1) return fail(c, EMC_ERR_INVALID, "emc_selftest_accumulate: bad argument");
    unsigned long long *d = nullptr, h = 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMalloc(&d, sizeof(unsigned long long));
    if (e != cudaSuccess) return fail_cuda(c, e, "cudaMalloc(selftest)");
    if ((e = cudaMemsetAsync(d, 0, sizeof(h), st)) == cudaSuccess &&
        (e = launch_selftest_accumulate(c, n, seed, max_count, d, st)) == cudaSuccess &&
        (e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, st)) == cudaSuccess)
        e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(c, e, "selftest_accumulate");
    *mismatches = (int64_t)h;
    return EMC_OK;
}

extern "C" int emc_aos_to_soa(EmcCtx *ctx, const double *coords6, int64_t n, double *kx, double *ky, double *kz,
                              double *x, double *y, double *z, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (n < 0 || (n > 0 && (!coords6 || !kx || !ky || !kz || !x || !y || !z)))
        return fail(c, EMC_ERR_INVALID, "emc_aos_to_soa: bad argument");
    if (((uintptr_t)coords6 & 15) != 0) return fail(c, EMC_ERR_INVALID, "coords6 must be 16-byte aligned");
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_aos_to_soa(c, coords6, n, kx, ky, kz, x, y, z, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "aos_to_soa launch");
}

extern "C" int emc_soa_to_aos(EmcCtx *ctx, const double *kx, const double *ky, const double *kz, const double *x,
                              const double *y, const double *z, int64_t n, double *coords6, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (n < 0 || (n > 0 && (!coords6 || !kx || !ky || !kz || !x || !y || !z)))
        return fail(c, EMC_ERR_INVALID, "emc_soa_to_aos: bad argument");
    if (((uintptr_t)coords6 & 15) != 0) return fail(c, EMC_ERR_INVALID, "coords6 must be 16-byte aligned");
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_soa_to_aos(c, kx, ky, kz, x, y, z, n, coords6, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "soa_to_aos launch");
}

extern "C" int emc_deposit(EmcCtx *ctx, const double *x, const double *y, const double *z, int64_t n, int64_t *mesh,
                           int32_t scheme, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (n < 0 || !mesh || (n > 0 && (!x || !y || !z))) return fail(c, EMC_ERR_INVALID, "emc_deposit: bad argument");
    if (scheme != EMC_DEPOSIT_NGP && scheme != EMC_DEPOSIT_CIC) return fail(c, EMC_ERR_INVALID, "unknown deposit scheme");
    if (n == 0) return EMC_OK;
    cudaError_t e = launch_deposit(c, x, y, z, n, mesh, scheme, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "deposit launch");
}

extern "C" int emc_set_deposit_mode(EmcCtx *ctx, int32_t mode)
{
    CTX_OR_FAIL(ctx);
    if (mode < 0 || mode > 2) return fail(c, EMC_ERR_INVALID, "emc_set_deposit_mode: 0 auto, 1 direct, 2 tiled");
    c->deposit_mode = mode;
    return EMC_OK;
}

extern "C" int emc_mesh_to_rho(EmcCtx *ctx, const int64_t *mesh, int32_t scheme, double *rho_pad, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!mesh || !rho_pad) return fail(c, EMC_ERR_INVALID, "emc_mesh_to_rho: null");
    if (scheme != EMC_DEPOSIT_NGP && scheme != EMC_DEPOSIT_CIC) return fail(c, EMC_ERR_INVALID, "unknown deposit scheme");
    cudaError_t e = launch_mesh_to_rho(c, mesh, scheme, rho_pad, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "mesh_to_rho launch");
}

extern "C" int emc_mesh_to_rho_interior(EmcCtx *ctx, const int64_t *mesh, int32_t scheme, double *rho_pad, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!mesh || !rho_pad) return fail(c, EMC_ERR_INVALID, "emc_mesh_to_rho_interior: null");
    if (scheme != EMC_DEPOSIT_NGP && scheme != EMC_DEPOSIT_CIC) return fail(c, EMC_ERR_INVALID, "unknown deposit scheme");
    cudaError_t e = launch_mesh_to_rho(c, mesh, scheme, rho_pad, (cudaStream_t)stream, true);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "mesh_to_rho launch");
}

extern "C" int emc_field_packed_interior(EmcCtx *ctx, const double *u_pad, double *e4, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!u_pad || !e4 || ((uintptr_t)e4 & 31) != 0) return fail(c, EMC_ERR_INVALID, "emc_field_packed_interior: null or unaligned");
    cudaError_t e = launch_field_packed(c, u_pad, e4, (cudaStream_t)stream, true);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "field_packed launch");
}

extern "C" int emc_poisson_sor(EmcCtx *ctx, double *u_pad, const double *rho_pad, double omega0, double tol,
                               int32_t max_sweeps, int32_t order, int32_t *sweeps_out, double *err_out,
                               double *err_hist, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!u_pad || !rho_pad) return fail(c, EMC_ERR_INVALID, "emc_poisson_sor: null");
    return run_poisson_sor(c, u_pad, rho_pad, omega0, tol, max_sweeps, order, sweeps_out, err_out, err_hist,
                           (cudaStream_t)stream);
}

extern "C" int emc_poisson_mg(EmcCtx *ctx, double *u_pad, const double *rho_pad, double tol_residual, int32_t max_cycles,
                              int32_t nu1, int32_t nu2, int32_t *cycles_out, double *res_out, double *res_hist,
                              void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!u_pad || !rho_pad) return fail(c, EMC_ERR_INVALID, "emc_poisson_mg: null");
    return run_poisson_mg(c, u_pad, rho_pad, tol_residual, max_cycles, nu1, nu2, cycles_out, res_out, res_hist,
                          (cudaStream_t)stream);
}

extern "C" int emc_poisson_residual(EmcCtx *ctx, const double *u_pad, const double *rho_pad, double *rms_out,
                                    double *max_out, double *out_dev, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!u_pad || !rho_pad || (!rms_out && !max_out && !out_dev)) return fail(c, EMC_ERR_INVALID, "emc_poisson_residual: null");
    return run_poisson_residual(c, u_pad, rho_pad, rms_out, max_out, out_dev, (cudaStream_t)stream);
}

extern "C" int emc_mesh_allreduce_p2p(EmcCtx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs,
                                      int32_t world, int32_t rank, int64_t n_cells, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!in_ptrs || !out_ptrs || world < 1 || world > EMC_MAX_PEERS || rank < 0 || rank >= world || n_cells < 1)
        return fail(c, EMC_ERR_INVALID, "emc_mesh_allreduce_p2p: bad argument");
    for (int p = 0; p < world; ++p)
        if (!in_ptrs[p] || !out_ptrs[p] || ((uintptr_t)in_ptrs[p] & 15) || ((uintptr_t)out_ptrs[p] & 15))
            return fail(c, EMC_ERR_INVALID, "emc_mesh_allreduce_p2p: null or unaligned peer pointer");
    cudaError_t e = launch_mesh_allreduce_p2p(c, in_ptrs, out_ptrs, world, rank, n_cells, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "mesh_allreduce_p2p launch");
}

extern "C" int emc_mesh_allreduce_rho_p2p(EmcCtx *ctx, const int64_t *const *in_ptrs, int64_t *const *out_ptrs,
                                          double *const *rho_ptrs, int32_t world, int32_t rank, int64_t n_cells,
                                          int32_t scheme, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!in_ptrs || !out_ptrs || !rho_ptrs || world < 1 || world > EMC_MAX_PEERS || rank < 0 || rank >= world ||
        (scheme != EMC_DEPOSIT_NGP && scheme != EMC_DEPOSIT_CIC))
        return fail(c, EMC_ERR_INVALID, "emc_mesh_allreduce_rho_p2p: bad argument");
    if (n_cells != (int64_t)c->dc.seg[0] * c->dc.seg[1] * c->dc.seg[2])
        return fail(c, EMC_ERR_INVALID, "emc_mesh_allreduce_rho_p2p: n_cells is not this context's mesh");
    for (int p = 0; p < world; ++p)
        if (!in_ptrs[p] || !out_ptrs[p] || !rho_ptrs[p] || ((uintptr_t)in_ptrs[p] & 15) || ((uintptr_t)out_ptrs[p] & 15) ||
            ((uintptr_t)rho_ptrs[p] & 7))
            return fail(c, EMC_ERR_INVALID, "emc_mesh_allreduce_rho_p2p: null or unaligned peer pointer");
    cudaError_t e = launch_mesh_allreduce_rho_p2p(c, in_ptrs, out_ptrs, rho_ptrs, world, rank, n_cells, scheme,
                                                  (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "mesh_allreduce_rho_p2p launch");
}

extern "C" int emc_field_packed(EmcCtx *ctx, const double *u_pad, double *e4, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!u_pad || !e4 || ((uintptr_t)e4 & 31) != 0) return fail(c, EMC_ERR_INVALID, "emc_field_packed: null or unaligned");
    cudaError_t e = launch_field_packed(c, u_pad, e4, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "field_packed launch");
}

extern "C" int emc_field_xz(EmcCtx *ctx, const double *u_pad, double *e2, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!u_pad || !e2 || ((uintptr_t)e2 & 15) != 0) return fail(c, EMC_ERR_INVALID, "emc_field_xz: null or unaligned");
    if (c->dc.seg[1] != 1) return fail(c, EMC_ERR_UNSUPPORTED, "emc_field_xz: only for meshes with ny = 1");
    cudaError_t e = launch_field_xz(c, u_pad, e2, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "field_xz launch");
}

extern "C" int emc_poisson_last(EmcCtx *ctx, int32_t *sweeps_out, double *err_out, void *stream)
{
    CTX_OR_FAIL(ctx);
    return poisson_last(c, sweeps_out, err_out, (cudaStream_t)stream);
}

extern "C" int emc_field(EmcCtx *ctx, const double *u_pad, double *ex, double *ey, double *ez, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!u_pad || !ex || !ey || !ez) return fail(c, EMC_ERR_INVALID, "emc_field: null");
    cudaError_t e = launch_field(c, u_pad, ex, ey, ez, (cudaStream_t)stream);
    return e == cudaSuccess ? EMC_OK : fail_cuda(c, e, "field launch");
}
