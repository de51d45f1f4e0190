/*
 * emc_oracle.c -- CPU oracle (plain C restatement of the reference algorithm).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see emc_oracle.h).  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC emc_oracle.c -lm
 * -ffp-contract=off matters: the reference is numpy, which never fuses a
 * multiply with an add, and the restatement keeps its operation ORDER so that
 * every step built only from IEEE + - * / sqrt is bit-identical to numpy.
 *
 * All citations are file:line into /root/reference.
 */
#include "emc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 constants).  Shared  */
/* SPEC with the CUDA path, separate implementation.                   */
/* ------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Draw `draw` of particle `pid` in timestep `step`:
 *   counter = (draw/2, step, pid_lo, pid_hi), key = (seed_lo, seed_hi);
 *   words (2h, 2h+1), h = draw%2, give a 53-bit uniform in [0,1) the way
 *   numpy's legacy generator does: (a>>5)*2^26 + (b>>6), times 2^-53.      */
double orc_philox_uniform(uint64_t seed, uint64_t pid, uint32_t step, uint32_t draw)
{
    uint32_t ctr[4] = { draw >> 1, step, (uint32_t)pid, (uint32_t)(pid >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t w[4];
    orc_philox4x32_10(ctr, key, w);
    uint32_t a = w[(draw & 1u) * 2u] >> 5, b = w[(draw & 1u) * 2u + 1u] >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

/* per-particle stream cursor */
typedef struct {
    const OrcStream *s;
    int64_t i;          /* local particle index */
    uint32_t draw;      /* philox: draw index within the timestep */
    int overflow;
} Cursor;

static double next_uniform(Cursor *c)
{
    const OrcStream *s = c->s;
    if (s->mode == 0) {
        int32_t pos = s->cursor[c->i];
        if (pos >= s->stride) { c->overflow = 1; return 0.5; }
        s->cursor[c->i] = pos + 1;
        return s->uniforms[c->i * s->stride + pos];
    }
    return orc_philox_uniform(s->seed, s->pid_offset + (uint64_t)c->i, s->step, c->draw++);
}

/* ------------------------------------------------------------------ */
/* calc_energy: scatter_.py:19-54 (twin main_.py:92-112)               */
/* ------------------------------------------------------------------ */
int orc_calc_energy(const OrcParams *p, const double k[3], int val, double *e_np, int32_t *idx)
{
    /* scatter_.py:44  E = sum(np.square(k))*hbar**2/(2*m*q); Python's sum() starts at 0 */
    double s = 0.0 + k[0] * k[0];
    s = s + k[1] * k[1];
    s = s + k[2] * k[2];
    double hbar2 = p->hbar * p->hbar;            /* (Parameters.hbar)**2 */
    double den = 2 * p->mass[val] * p->q;        /* (2*m)*q */
    double E = s * hbar2 / den;
    if (isnan(E)) return ORC_FAULT_NAN;          /* scatter_.py:45 */
    if (E == 0.0) return ORC_FAULT_ZERO_E;       /* scatter_.py:47 */
    double a = p->nonparab[val];
    double enp = (-1 + sqrt(1 + 4 * a * E)) / (2 * a);   /* scatter_.py:51 */
    /* scatter_.py:52  (dfEk - E_np).abs().idxmin(): first index of the minimum.
       The grid is monotone, so only the neighbours of the bracketing cell can win;
       scan them in ascending order with a strict '<' to keep the first minimum. */
    int n = p->n_energy;
    double step = p->ek[1] - p->ek[0];
    double t = enp / step;
    long j0 = (t < 0) ? 0 : (t > (double)(n - 1) ? (n - 1) : (long)t);
    long lo = j0 - 2 < 0 ? 0 : j0 - 2;
    long hi = j0 + 2 > n - 1 ? n - 1 : j0 + 2;
    long best = lo;
    double bd = fabs(p->ek[lo] - enp);
    for (long j = lo + 1; j <= hi; ++j) {
        double d = fabs(p->ek[j] - enp);
        if (d < bd) { bd = d; best = j; }
    }
    *e_np = enp;
    *idx = (int32_t)best;
    return -1;
}

/* free_flight: main_.py:114-123  (-1/G0)*np.log(r1) */
double orc_free_flight(double g0, double r1)
{
    return (-1 / g0) * log(r1);
}

/* carrier_drift + cyclic_boundary + specular_refl: main_.py:179-213,125-154 */
void orc_carrier_drift(const OrcParams *p, double coord[6], const double ef[3], double dt2, double mass)
{
    double k_[3], r_[3];
    for (int a = 0; a < 3; ++a) {
        /* main_.py:196  k - q*dt2*elec_field*1E2/hbar */
        k_[a] = coord[a] - p->q * dt2 * ef[a] * 1E2 / p->hbar;
        /* main_.py:197  r + (hbar*k - 0.5*q*elec_field*1e2*dt2)*(dt2/mass) */
        r_[a] = coord[3 + a] + (p->hbar * coord[a] - 0.5 * p->q * ef[a] * 1e2 * dt2) * (dt2 / mass);
    }
    /* main_.py:134  r[:2] = r[:2] - dim*(r>=dim) + dim*(r<=0)  (single wrap) */
    for (int a = 0; a < 2; ++a) {
        double d = p->tot_dim[a];
        r_[a] = r_[a] - d * (double)(r_[a] >= d) + d * (double)(r_[a] <= 0);
    }
    /* main_.py:151-153  reflect until inside (NaN guard added: the reference would spin) */
    double L = p->tot_dim[2];
    int guard = 0;
    /* extension for bulk runs: z wraps like x and y */
    if (p->z_cyclic) r_[2] = r_[2] - L * (double)(r_[2] >= L) + L * (double)(r_[2] <= 0);
    while (!p->z_cyclic && !(r_[2] >= 0 && r_[2] <= L) && !isnan(r_[2]) && guard++ < 1000000) {
        int out = (r_[2] > L) || (r_[2] < 0);
        int in  = (r_[2] > 0) && (r_[2] < L);
        k_[2] = -1 * out * k_[2] + in * k_[2];
        int ge = r_[2] >= L, le = r_[2] <= 0;
        r_[2] = 2 * L * ge + -1 * (ge || le) * r_[2] + in * r_[2];
    }
    for (int a = 0; a < 3; ++a) { coord[a] = k_[a]; coord[3 + a] = r_[a]; }
}

/* polar-angle samplers, scatter_.py:795-889.  Return the ANGLE (rad). */
static double sample_polar(const OrcParams *p, int sampler, double Ek, double r)
{
    switch (sampler) {
    case ORC_SAMPLER_ISO:      /* scatter_.py:811 */
        return acos(1 - 2 * r);
    case ORC_SAMPLER_POP_ABS: {/* scatter_.py:828-831 */
        double Ep = p->e_phonon;
        double d = sqrt(Ek) - sqrt(Ek + Ep);
        double eta = 2 * sqrt(Ek * (Ek + Ep)) / (d * d);
        return acos(((1 + eta) - pow(1 + 2 * eta, r)) / eta);
    }
    case ORC_SAMPLER_POP_EMS: {/* scatter_.py:848-851 */
        double Ep = p->e_phonon;
        double d = sqrt(Ek) - sqrt(Ek - Ep);
        double eta = 2 * sqrt(Ek * (Ek - Ep)) / (d * d);
        return acos(((1 + eta) - pow(1 + 2 * eta, r)) / eta);
    }
    case ORC_SAMPLER_ION: {    /* scatter_.py:868-872 */
        double alpha = 2 * atan(p->ion_eb / (Ek + 1E-6));
        return acos(1 - (1 - cos(alpha)) / (1 - r * (1 - 0.5 * (1 - cos(alpha)))));
    }
    default:                   /* scatter_.py:889 self-scattering: 0 */
        return 0.0;
    }
}

/* choose_mech: scatter_.py:1029-1034 -- first column with cum > r2 (strict) */
static int choose_mech(const OrcParams *p, int val, int32_t idx, double r2)
{
    int n = p->n_mech[val];
    const double *row = p->cum[val] + (int64_t)idx * n;
    for (int j = 0; j < n; ++j)
        if (row[j] > r2) return j;
    return -1;
}

/* scatter_angles: scatter_.py:1059-1089, given the three uniforms of the event.
   Fault policy (the reference raises/propagates NaN; see DESIGN.md "faults"): any
   fault leaves coord/val untouched and returns the code. */
static int scatter_core(const OrcParams *p, double coord[6], int32_t *val, int mech,
                        double e_np, double r_az, double r_pol)
{
    int v = *val;
    double kx = coord[0], ky = coord[1], kz = coord[2];
    double s = 0.0 + kx * kx; s = s + ky * ky; s = s + kz * kz;
    double k0 = sqrt(s);                          /* :1060 */
    double p0 = acos(kz / k0);                    /* :1061 */
    double a0 = asin(kx / (k0 * sin(p0)));        /* :1062 */
    double az = 2 * M_PI * r_az;                  /* :1068 */
    double E0 = e_np;                             /* :1069 */
    int sampler = p->mech_sampler[v][mech];
    double polar = sample_polar(p, sampler, E0, r_pol);   /* :1070 */
    double dE = p->mech_dE[v][mech];
    if (E0 + dE < 0) return ORC_FAULT_NEG_ENERGY;
    double kp = sqrt(2 * p->mass[v] * p->q * (E0 + dE) / (p->hbar * p->hbar));   /* :1071 */
    double kxr = kp * sin(polar) * cos(az);       /* :1072 */
    double kyr = kp * sin(polar) * sin(az);       /* :1073 */
    double kzr = kp * cos(polar);                 /* :1074 */
    double kxp = kxr * cos(a0) * cos(p0) - kyr * sin(a0) + kzr * cos(p0) * sin(a0);   /* :1075 */
    double kyp = kxr * sin(a0) * cos(p0) + kyr * cos(a0) + kzr * sin(p0) * sin(a0);   /* :1076 */
    double kzp = -kxr * sin(p0) + kzr * cos(p0);                                      /* :1077 */
    if (isnan(kxp) || isnan(kyp) || isnan(kzp)) return ORC_FAULT_DOMAIN;
    if (kxp == 0 || kyp == 0 || kzp == 0) return ORC_FAULT_ZERO_K;                    /* :1079 */
    coord[0] = kxp; coord[1] = kyp; coord[2] = kzp;                                   /* :1088 */
    *val = p->mech_trans[v][mech];                                                    /* :1098 */
    return -1;
}

int orc_scatter(const OrcParams *p, double coord[6], int32_t *val, const double u[3],
                int32_t *mech_out, int32_t *n_draws)
{
    double e_np; int32_t idx;
    *n_draws = 0;
    int f = orc_calc_energy(p, coord, *val, &e_np, &idx);      /* scatter_.py:1092 */
    if (f >= 0) return f;
    int mech = choose_mech(p, *val, idx, u[0]);
    *n_draws = 1;
    if (mech < 0) return ORC_FAULT_NO_MECH;
    if (mech_out) *mech_out = mech;
    int self = p->mech_sampler[*val][mech] == ORC_SAMPLER_SELF;
    *n_draws = self ? 2 : 3;
    return scatter_core(p, coord, val, mech, e_np, u[1], u[2]);
}

/* scatter with draws pulled from the particle's stream, in the reference's order:
   r2 (:1029), azimuth (:1068), polar (:811/830/850/871; none for self :889) */
static int scatter_stream(const OrcParams *p, double coord[6], int32_t *val, Cursor *c)
{
    double e_np; int32_t idx;
    int f = orc_calc_energy(p, coord, *val, &e_np, &idx);
    if (f >= 0) return f;
    /* Philox layout: every event starts on draw index 4e of the (particle, timestep) stream:
       block 2e = (r2, azimuth), block 2e+1 = (polar, flight).  A self-scattering event draws no
       polar angle and its azimuth has no effect (sin(polar) = 0), so the flight that follows
       re-reads the azimuth draw 4e+1 and block 2e+1 is never evaluated. */
    if (c->s->mode == 1) c->draw = (c->draw + 3u) & ~3u;
    double r2 = next_uniform(c);
    int mech = choose_mech(p, *val, idx, r2);
    if (mech < 0) return ORC_FAULT_NO_MECH;
    double r_az = next_uniform(c);
    double r_pol = 0.0;
    if (p->mech_sampler[*val][mech] != ORC_SAMPLER_SELF) r_pol = next_uniform(c);
    else if (c->s->mode == 1) c->draw -= 1u;      /* the caller's free-flight draw is 4e+1 again */
    return scatter_core(p, coord, val, mech, e_np, r_az, r_pol);
}

/* field seen by particle i at position r (V/cm) */
static void field_at(const OrcParams *p, const OrcStream *s, int64_t i, const double coord[6], double ef[3])
{
    if (p->field_mode == 1) {
        int64_t g = (int64_t)((s->pid_offset + (uint64_t)i) / (uint64_t)p->group_size);
        if (g >= p->n_groups) g = p->n_groups - 1;
        for (int a = 0; a < 3; ++a) ef[a] = p->field_groups[3 * g + a];
    } else if (p->field_mode == 2) {
        int c[3];
        for (int a = 0; a < 3; ++a) {
            int j = (int)floor(coord[3 + a] / p->dl[a]);
            if (j < 0) j = 0;
            if (j > p->seg[a] - 1) j = p->seg[a] - 1;
            c[a] = j + 1;
        }
        int64_t idx = ((int64_t)c[0] * (p->seg[1] + 2) + c[1]) * (p->seg[2] + 2) + c[2];
        ef[0] = p->ex[idx] * 1e-2; ef[1] = p->ey[idx] * 1e-2; ef[2] = p->ez[idx] * 1e-2;
    } else {
        for (int a = 0; a < 3; ++a) ef[a] = p->efield[a];
    }
}

static void obs_zero(OrcObs *o) { memset(o, 0, sizeof(*o)); }

static void obs_add(OrcObs *a, const OrcObs *b)
{
    a->sum_z += b->sum_z; a->sum_vz += b->sum_vz; a->sum_e += b->sum_e;
    a->sum_vz2 += b->sum_vz2; a->sum_e2 += b->sum_e2;
    for (int v = 0; v < 3; ++v) a->n_valley[v] += b->n_valley[v];
    a->n_fault += b->n_fault; a->n_segments += b->n_segments; a->n_events += b->n_events;
}

/* where_am_i: init_.py:36-56 */
int orc_where_am_i(const OrcParams *p, double dist)
{
    if (dist < 0) return -1;                      /* :45-46 */
    for (int l = 0; l < p->n_layers; ++l) {       /* :48 */
        if (dist <= p->layer_lz[l]) return l;     /* :50 */
        dist -= p->layer_lz[l];                   /* :54 */
    }
    return -1;                                    /* :56 */
}

/* material of layer l (single-layer runs: the struct itself) */
static const OrcParams *mat_of(const OrcParams *p, int l)
{
    return (p->n_layers > 1 && l >= 0) ? p->layer_mat[l] : p;
}

/* main_.py:393-397 for one particle */
static void obs_particle(const OrcParams *pg, const double coord[6], int val, OrcObs *o)
{
    if (val < 0) { o->n_fault++; return; }
    const OrcParams *p = pg;
    if (pg->n_layers > 1) p = mat_of(pg, orc_where_am_i(pg, coord[5]));      /* :393 */
    double E = (0.0 + coord[0] * coord[0]); E = E + coord[1] * coord[1]; E = E + coord[2] * coord[2];
    E = E * (p->hbar * p->hbar) / (2 * p->mass[val] * p->q);           /* main_.py:105 */
    double a = p->nonparab[val];
    double enp = (-1 + sqrt(1 + 4 * a * E)) / (2 * a);                 /* main_.py:109 */
    double vz = coord[2] * p->hbar / p->mass[val];                     /* main_.py:395 */
    o->sum_e += enp; o->sum_e2 += enp * enp;
    o->sum_vz += vz; o->sum_vz2 += vz * vz;
    o->sum_z += coord[5];
    o->n_valley[val]++;
}

/* main_.py:360-397: one timestep for one particle */
static void step_particle(const OrcParams *p, double coord[6], double *dtau, int32_t *valley,
                          Cursor *c, OrcObs *o)
{
    if (*valley < 0) return;                      /* frozen after a fault */
    double dt = p->dt;
    double dte = *dtau;                           /* :361 */
    double ef[3];
    const int ML = p->n_layers > 1;
    const OrcParams *pm = p;                      /* mat[mat_i] / scat_tables[mat_i] */
    if (ML) {                                     /* :366 / :372 */
        int l = orc_where_am_i(p, coord[5]);
        if (l < 0) { *valley = -(1 + ORC_FAULT_OUTSIDE); return; }
        pm = mat_of(p, l);
    }
    field_at(p, c->s, c->i, coord, ef);
    if (dte >= dt) {                              /* :363 */
        orc_carrier_drift(p, coord, ef, dt, pm->mass[*valley]);     /* :367 */
        o->n_segments++;
    } else {
        double dt2 = dte;                         /* :370 */
        orc_carrier_drift(p, coord, ef, dt2, pm->mass[*valley]);    /* :373 */
        o->n_segments++;
        while (dte < dt) {                        /* :375 */
            double dte2 = dte;                    /* :376 */
            if (ML) {                             /* :377 (and :387: a scatter does not move the particle) */
                int l = orc_where_am_i(p, coord[5]);
                if (l < 0) { *valley = -(1 + ORC_FAULT_OUTSIDE); *dtau = dte; return; }
                pm = mat_of(p, l);
            }
            int f = scatter_stream(pm, coord, valley, c);                 /* :378 */
            if (c->overflow) f = ORC_FAULT_STREAM;
            if (f >= 0) { *valley = -(1 + f); *dtau = dte; return; }
            o->n_events++;
            double dt3 = orc_free_flight(p->gamma, next_uniform(c));      /* :379 */
            if (c->overflow) { *valley = -(1 + ORC_FAULT_STREAM); *dtau = dte; return; }
            double dtp = dt - dte2;               /* :381 */
            dt2 = (dt3 <= dtp) ? dt3 : dtp;       /* :382-385 */
            if (p->field_mode == 2) field_at(p, c->s, c->i, coord, ef);
            orc_carrier_drift(p, coord, ef, dt2, pm->mass[*valley]);/* :388 */
            o->n_segments++;
            dte = dte2 + dt3;                     /* :389 */
        }
    }
    dte -= dt;                                    /* :390 */
    *dtau = dte;                                  /* :391 */
}

void orc_timestep(const OrcParams *p, int64_t n, double *kx, double *ky, double *kz,
                  double *x, double *y, double *z, double *dtau, int32_t *valley,
                  OrcStream *s, OrcObs *obs, int n_threads)
{
    OrcObs total; obs_zero(&total);
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
    #pragma omp parallel num_threads(n_threads)
#endif
    {
        OrcObs local; obs_zero(&local);
#ifdef _OPENMP
        #pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < n; ++i) {
            double coord[6] = { kx[i], ky[i], kz[i], x[i], y[i], z[i] };
            Cursor c = { s, i, 0u, 0 };
            step_particle(p, coord, &dtau[i], &valley[i], &c, &local);
            kx[i] = coord[0]; ky[i] = coord[1]; kz[i] = coord[2];
            x[i] = coord[3]; y[i] = coord[4]; z[i] = coord[5];
            if (p->n_layers > 1 && valley[i] >= 0 && orc_where_am_i(p, coord[5]) < 0)
                valley[i] = -(1 + ORC_FAULT_OUTSIDE);              /* :393 would raise */
            obs_particle(p, coord, valley[i], &local);
        }
#ifdef _OPENMP
        #pragma omp critical
#endif
        obs_add(&total, &local);
    }
    if (obs) *obs = total;
}

/* ------------------------------------------------------------------ */
/* init_coords: init_.py:194-234; valley 0 main_.py:53; dtau main_.py:337
 * The reference draws from numpy's global MT19937 stream (all x, all y, all z,
 * then all energies, then two normals per particle), which a counter-based
 * per-particle generator cannot reproduce draw for draw; the SAME DISTRIBUTIONS are
 * sampled per particle from Philox (step word 0xFFFFFFFF), draws in this order:
 *   0,1,2  x, y, z = box * u                               (init_.py:198-200)
 *   3..6   n1,n2 = BoxMuller(u3,u4); n3,_ = BoxMuller(u5,u6);
 *          E = mean*kT + 1e-7*sqrt(n1^2+n2^2+n3^2)        (scipy maxwell == chi(3), init_.py:207)
 *   7,8    alpha, beta = 2*pi*BoxMuller(u7,u8)             (Normal(0, 2*pi), init_.py:216-217)
 *   9      dtau = (-1/G0)*log(u9)                          (main_.py:337)
 * BoxMuller(a,b) = sqrt(-2 log(1-a)) * (cos 2 pi b, sin 2 pi b).                  */
/* ------------------------------------------------------------------ */
static void box_muller(double ua, double ub, double *n0, double *n1)
{
    double rad = sqrt(-2 * log(1 - ua));
    *n0 = rad * cos(2 * M_PI * ub);
    *n1 = rad * sin(2 * M_PI * ub);
}

void orc_init_ensemble(const OrcParams *p, int64_t n, double *kx, double *ky, double *kz,
                       double *x, double *y, double *z, double *dtau, int32_t *valley,
                       uint64_t seed, uint64_t pid_offset)
{
#ifdef _OPENMP
    #pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < n; ++i) {
        double u[10];
        for (uint32_t d = 0; d < 10; ++d)
            u[d] = orc_philox_uniform(seed, pid_offset + (uint64_t)i, 0xFFFFFFFFu, d);
        x[i] = p->tot_dim[0] * u[0];
        y[i] = p->tot_dim[1] * u[1];
        z[i] = p->tot_dim[2] * u[2];
        const OrcParams *pm = p;
        if (p->n_layers > 1) {
            int layer = 0;
            if (p->layer_cw[p->n_layers - 1] > 0) {
                /* extension (doping profiles): piecewise-uniform z by inverse CDF of u[2] */
                double c0 = 0.0, z0 = 0.0;
                while (layer < p->n_layers - 1 && u[2] >= p->layer_cw[layer]) {
                    c0 = p->layer_cw[layer];
                    z0 += p->layer_lz[layer];
                    ++layer;
                }
                z[i] = z0 + p->layer_lz[layer] * ((u[2] - c0) / (p->layer_cw[layer] - c0));
            }
            layer = orc_where_am_i(p, z[i]);      /* init_.py:214 */
            if (layer < 0) layer = p->n_layers - 1;
            pm = mat_of(p, layer);
        }
        double n1, n2, n3, n4, na, nb;
        box_muller(u[3], u[4], &n1, &n2);
        box_muller(u[5], u[6], &n3, &n4);
        double e = p->mean_energy * p->kT + 1E-7 * sqrt(n1 * n1 + n2 * n2 + n3 * n3);
        double k = sqrt(2 * pm->mass[0] * p->q * e / (p->hbar * p->hbar));   /* init_.py:215 */
        box_muller(u[7], u[8], &na, &nb);
        double alpha = 2 * M_PI * na, beta = 2 * M_PI * nb;
        kx[i] = k * cos(alpha) * sin(beta);       /* init_.py:218-220 */
        ky[i] = k * sin(alpha) * sin(beta);
        kz[i] = k * cos(beta);
        dtau[i] = (-1 / p->gamma) * log(u[9]);
        valley[i] = 0;
    }
}

/* init_.init_photoex (init_.py:244-280) + main_.py:349-351.  Candidates drawn beyond the slab
 * are dropped (init_.py:247-248), survivors appended in candidate order.  Draws of candidate c
 * (Philox step word 0xFFFFFFFE, particle word cand_offset + c):
 *   0 z = -scale*log(1-u);  1,2 x,y = std*BoxMuller;  3,4 alpha,beta = 2 pi*BoxMuller;
 *   5 dtau = (-1/gamma_first)*log(u).   Returns the number of carriers written.          */
int64_t orc_init_photoex(const OrcParams *p, int64_t n_cand, double energy, double z_scale,
                         double xy_std, double gamma_first, uint64_t seed, uint64_t cand_offset,
                         double *kx, double *ky, double *kz, double *x, double *y, double *z,
                         double *dtau, int32_t *valley)
{
    int64_t o = 0;
    for (int64_t c = 0; c < n_cand; ++c) {
        double u[6];
        for (uint32_t d = 0; d < 6; ++d)
            u[d] = orc_philox_uniform(seed, cand_offset + (uint64_t)c, 0xFFFFFFFEu, d);
        double pz = -z_scale * log(1 - u[0]);                 /* init_.py:246 */
        if (!(pz < p->tot_dim[2])) continue;                  /* :247-248 */
        double n0, n1, na, nb;
        box_muller(u[1], u[2], &n0, &n1);
        box_muller(u[3], u[4], &na, &nb);
        const OrcParams *pm = p;
        if (p->n_layers > 1) {
            int layer = orc_where_am_i(p, pz);                /* :262 */
            if (layer < 0) layer = p->n_layers - 1;
            pm = mat_of(p, layer);
        }
        double k = sqrt(2 * pm->mass[0] * p->q * energy / (p->hbar * p->hbar));   /* :263 */
        double alpha = 2 * M_PI * na, beta = 2 * M_PI * nb;
        kx[o] = k * cos(alpha) * sin(beta);                   /* :266-268 */
        ky[o] = k * sin(alpha) * sin(beta);
        kz[o] = k * cos(beta);
        x[o] = xy_std * n0;
        y[o] = xy_std * n1;
        z[o] = pz;
        dtau[o] = (-1 / gamma_first) * log(u[5]);             /* main_.py:351 */
        valley[o] = 0;                                        /* main_.py:349 */
        ++o;
    }
    return o;
}

void orc_observables(const OrcParams *p, int64_t n, const double *kx, const double *ky,
                     const double *kz, const double *z, const int32_t *valley, OrcObs *obs)
{
    obs_zero(obs);
    for (int64_t i = 0; i < n; ++i) {
        double coord[6] = { kx[i], ky[i], kz[i], 0, 0, z[i] };
        int val = valley[i];
        if (p->n_layers > 1 && val >= 0 && orc_where_am_i(p, z[i]) < 0) val = -(1 + ORC_FAULT_OUTSIDE);
        obs_particle(p, coord, val, obs);
    }
}

/* ------------------------------------------------------------------ */
/* charge assignment: poisson_.py:20,30-44                             */
/* ------------------------------------------------------------------ */
/* grid centre j on an axis: np.arange(1, 2n+1, 2)*dl/2  (poisson_.py:20) */
static inline double centre(int j, double dl) { return (double)(2 * j + 1) * dl / 2; }

/* cells j on one axis with |centre_j - r| <= dl/2  (poisson_.py:31); at most 2 */
static int axis_hits(double r, int n, double dl, int out[3])
{
    int cnt = 0;
    double t = r / dl;
    if (!(t > -2.0 && t < (double)n + 2.0)) return 0;
    int j0 = (int)floor(t);
    for (int j = j0 - 1; j <= j0 + 1; ++j) {
        if (j < 0 || j >= n) continue;
        if (fabs(centre(j, dl) - r) <= dl / 2) { if (cnt < 3) out[cnt++] = j; }
    }
    return cnt;
}

void orc_deposit_ngp(int64_t n, const double *x, const double *y, const double *z,
                     const int32_t seg[3], const double dl[3], int64_t *counts)
{
    for (int64_t i = 0; i < n; ++i) {
        int hx[3], hy[3], hz[3];
        int nx = axis_hits(x[i], seg[0], dl[0], hx);
        int ny = axis_hits(y[i], seg[1], dl[1], hy);
        int nz = axis_hits(z[i], seg[2], dl[2], hz);
        for (int a = 0; a < nx; ++a)
            for (int b = 0; b < ny; ++b)
                for (int c = 0; c < nz; ++c)   /* itertools.product order: z fastest (poisson_.py:26,42) */
                    counts[((int64_t)hx[a] * seg[1] + hy[b]) * seg[2] + hz[c]] += 1;
    }
}

/* cloud-in-cell, weight = prod_axis (1 - |centre - r|/dl) over the two nearest centres
   (the reference's commented-out line poisson_.py:34), quantised to 2^-32 so that
   the sum is integer and order independent.  Out-of-mesh neighbours are dropped. */
static int axis_cic(double r, int n, double dl, int j[2], double w[2])
{
    double t = r / dl - 0.5;
    double f = floor(t);
    int j0 = (int)f;
    double w1 = t - f;
    j[0] = j0; j[1] = j0 + 1; w[0] = 1.0 - w1; w[1] = w1;
    (void)n;
    return 2;
}

void orc_deposit_cic(int64_t n, const double *x, const double *y, const double *z,
                     const int32_t seg[3], const double dl[3], int64_t *mesh_fx)
{
    const double SCALE = 4294967296.0;
    for (int64_t i = 0; i < n; ++i) {
        int jx[2], jy[2], jz[2]; double wx[2], wy[2], wz[2];
        axis_cic(x[i], seg[0], dl[0], jx, wx);
        axis_cic(y[i], seg[1], dl[1], jy, wy);
        axis_cic(z[i], seg[2], dl[2], jz, wz);
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
            if (jx[a] < 0 || jx[a] >= seg[0] || jy[b] < 0 || jy[b] >= seg[1] ||
                jz[c] < 0 || jz[c] >= seg[2]) continue;
            double w = wx[a] * wy[b] * wz[c];
            mesh_fx[((int64_t)jx[a] * seg[1] + jy[b]) * seg[2] + jz[c]] += (int64_t)llrint(w * SCALE);
        }
    }
}

void orc_counts_to_rho(int64_t n_cells, const int64_t *counts, double background,
                       double increment, double *rho)
{
    for (int64_t c = 0; c < n_cells; ++c) {
        double r = background;                       /* poisson_.py:27 */
        for (int64_t k = 0; k < counts[c]; ++k) r += increment;   /* poisson_.py:36 */
        rho[c] = r;
    }
}

/* Checker for the CUDA path's division by run constants (csrc/emc_device.cuh div_const): the
   same five operations in C (fma from libm) against the IEEE quotient, on `n` random
   dividends a = m * 2^e, m in [1,2), e uniform in [e_lo, e_hi].  Returns the mismatches. */
int64_t orc_div_const_mismatches(double b, int64_t n, uint64_t seed, int e_lo, int e_hi)
{
    const double rb = 1 / b;
    int64_t bad = 0;
    uint64_t s = seed * 6364136223846793005ull + 1442695040888963407ull;
    for (int64_t i = 0; i < n; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        double m = 1.0 + (double)(s >> 11) * (1.0 / 9007199254740992.0);
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        int e = e_lo + (int)((s >> 33) % (uint64_t)(e_hi - e_lo + 1));
        double a = ldexp((s >> 20) & 1 ? -m : m, e);
        double q = a * rb;
        double r = fma(-q, b, a);
        q = fma(r, rb, q);
        r = fma(-q, b, a);
        q = fma(r, rb, q);
        if (q != a / b) ++bad;
    }
    return bad;
}

void orc_fx_to_rho(int64_t n_cells, const int64_t *mesh_fx, double background,
                   double increment, double *rho)
{
    for (int64_t c = 0; c < n_cells; ++c)
        rho[c] = background + increment * ((double)mesh_fx[c] / 4294967296.0);
}

/* ------------------------------------------------------------------ */
/* Poisson SOR: poisson_.py:55-84                                      */
/* ------------------------------------------------------------------ */
/* numpy's pairwise summation (add.reduce over a contiguous double array) */
static double np_pairwise_sum(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

int orc_poisson_sor(const int32_t seg[3], const double dl[3], double eps_s, double *u,
                    const double *rho, double omega0, double tol, int max_sweeps,
                    double *err_hist)
{
    const int nx = seg[0], ny = seg[1], nz = seg[2];
    const int64_t sy = nz + 2, sx = (int64_t)(ny + 2) * (nz + 2);
    const int64_t total = (int64_t)(nx + 2) * sx;
    /* poisson_.py:56-57,61 */
    const double xy = (dl[0] / dl[1]) * (dl[0] / dl[1]);
    const double xz = (dl[0] / dl[2]) * (dl[0] / dl[2]);
    const double e = -2 * (1 + xy + xz);
    /* poisson_.py:68 */
    const double pch = 0.5 * cos(M_PI / nx) + 0.5 * cos(M_PI / ny) + 0.5 * cos(M_PI / nz);
    double relx = omega0;                              /* poisson_.py:66 */
    double err = 1;
    int num = 0;
    double *prev = (double *)malloc(sizeof(double) * total);
    double *sq = (double *)malloc(sizeof(double) * total);
    while (err > tol && num < max_sweeps) {            /* poisson_.py:69 */
        memcpy(prev, u, sizeof(double) * total);       /* :73 */
        for (int k = 1; k <= nz; ++k)                  /* :74 product(k, j, i): i fastest */
            for (int j = 1; j <= ny; ++j)
                for (int i = 1; i <= nx; ++i) {
                    int64_t c = i * sx + j * sy + k;
                    /* :58-60 adj list; np.sum over the 7-element list is a plain left-to-right
                       sum (numpy's pairwise kernel is sequential below 8 elements) */
                    double sum = e * u[c];
                    sum += u[c + sx];
                    sum += u[c - sx];
                    sum += xy * u[c + sy];
                    sum += xy * u[c - sy];
                    sum += xz * u[c + 1];
                    sum += xz * u[c - 1];
                    /* :76 */
                    double du = (sum - dl[0] * dl[0] * rho[c] / eps_s) * relx / e;
                    u[c] -= du;                        /* :77 */
                    if (fabs(du) >= 1e3) { free(prev); free(sq); return -(num + 1); }   /* :78-80 */
                }
        relx = 1 / (1 - 0.25 * (pch * pch) * relx);    /* :81 */
        for (int64_t c = 0; c < total; ++c) { double d = u[c] - prev[c]; sq[c] = d * d; }
        err = sqrt(np_pairwise_sum(sq, total) / (double)total);   /* :82 */
        if (err_hist) err_hist[num] = err;
        num += 1;
    }
    free(prev); free(sq);
    return num;
}

void orc_field(const int32_t seg[3], const double dl[3], const double *u,
               double *ex, double *ey, double *ez)
{
    const int nx = seg[0], ny = seg[1], nz = seg[2];
    const int64_t sy = nz + 2, sx = (int64_t)(ny + 2) * (nz + 2);
    for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                int64_t c = i * sx + j * sy + k;
                ex[c] = -0.5 * (u[c + sx] - u[c - sx]) / dl[0];   /* poisson_.py:94 */
                ey[c] = -0.5 * (u[c + sy] - u[c - sy]) / dl[1];   /* :95 */
                ez[c] = -0.5 * (u[c + 1] - u[c - 1]) / dl[2];     /* :96 */
            }
}
