// emc_flight.cu -- the fused free-flight / drift / scatter kernel of one EMC timestep
// (main_.py:360-397), its observables (main_.py:394-403), the ensemble initialiser
// (init_.py:194-234) and the single-function batch kernels behind the reference's
// per-particle signatures.  sm_100a; compiled with -fmad=false (see emc_device.cuh).
#include <cstddef>
#include <type_traits>
#include "emc_host.h"

namespace emc {

// ------------------------------------------------------------------------------------------
// block-level reduction of the observables
// ------------------------------------------------------------------------------------------
// per-lane accumulators: doubles for the sums, 32-bit counters (a lane sees far fewer than 2^32
// particles per launch); widened to 64 bit in the block reduction
struct ObsAcc {
    double d[5];        // sum_z, sum_vz, sum_e, sum_vz2, sum_e2
    unsigned int c[6];  // n_valley[3], n_fault, n_segments, n_events
    __device__ __forceinline__ void zero(double *)
    {
#pragma unroll
        for (int k = 0; k < 5; ++k) d[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) c[k] = 0u;
    }
    template <int K> __device__ __forceinline__ void add(double v) { d[K] += v; }
    template <int K> __device__ __forceinline__ double sum() const { return d[K]; }
    __device__ __forceinline__ void count_valley(int val) { c[0] += (val == 0); c[1] += (val == 1); c[2] += (val == 2); }
    __device__ __forceinline__ void count_fault() { c[3]++; }
    __device__ __forceinline__ void count_segment() { c[4]++; }
    __device__ __forceinline__ void derive_segments() { c[4] = c[0] + c[1] + c[2] + c[5]; }
    __device__ __forceinline__ void count_event() { c[5]++; }
    __device__ __forceinline__ long long counter(int k) const { return (long long)c[k]; }
};

// The same accumulators with a smaller register footprint (EMC_OBS_SMEM, for block shapes that
// must fit 96 registers): the five double sums live in a per-thread column of shared memory
// ([k][thread]: conflict-free), the three valley counts and the fault count are packed into two
// words (16 bits each: a lane sees fewer than 65 536 particles per launch, checked by the launcher).
struct ObsAccShared {
    double *s;          // this thread's column: s[k * FLIGHT_THREADS]
    unsigned int v01, v2f, seg, ev;
    __device__ __forceinline__ void zero(double *column)
    {
        s = column;
#pragma unroll
        for (int k = 0; k < 5; ++k) s[k * FLIGHT_THREADS] = 0.0;
        v01 = v2f = seg = ev = 0u;
    }
    template <int K> __device__ __forceinline__ void add(double v) { s[K * FLIGHT_THREADS] += v; }
    template <int K> __device__ __forceinline__ double sum() const { return s[K * FLIGHT_THREADS]; }
    __device__ __forceinline__ void count_valley(int val)
    {
        v01 += (val == 0) + ((val == 1) << 16);
        v2f += (val == 2);
    }
    __device__ __forceinline__ void count_fault() { v2f += 1u << 16; }
    __device__ __forceinline__ void count_segment() { seg++; }
    __device__ __forceinline__ void derive_segments() { seg = (v01 & 0xffffu) + (v01 >> 16) + (v2f & 0xffffu) + ev; }
    __device__ __forceinline__ void count_event() { ev++; }
    __device__ __forceinline__ long long counter(int k) const
    {
        const unsigned int w[6] = {v01 & 0xffffu, v01 >> 16, v2f & 0xffffu, v2f >> 16, seg, ev};
        return (long long)w[k];
    }
};

template <class Acc>
__device__ __forceinline__ void obs_particle(const DevConst &P, const Mat &MT, Acc &a, double kx, double ky,
                                             double kz, double z, int val)
{
    if (val < 0) { a.count_fault(); return; }
    // main_.py:105,109,394-397 (MT = mat[mat_i] of the final position, main_.py:393)
    const double E = energy_parabolic(P, MT, kx, ky, kz, val);
    const double enp = energy_nonparabolic(MT, E, val);
    const double vz = div_const(kz * P.hbar, MT.mass[val], MT.r_mass[val]);
    a.template add<0>(z);
    a.template add<1>(vz);
    a.template add<2>(enp);
    a.template add<3>(vz * vz);
    a.template add<4>(enp * enp);
    a.count_valley(val);
}

// Deterministic two-level reduction: warp shuffles, then shared memory, one partial per block;
// the last block to finish (ticket) folds the partials in a fixed order, so the result does
// not depend on block scheduling.
// (bidx / nblk: the block's slot and the number of slots -- blockIdx.x / gridDim.x, or the slice of a timestep
// chain's work item and the number of slices; a chain's per-timestep tickets are zeroed by the host, not here)
template <class Acc>
__device__ void obs_block_reduce_and_finish(const Acc &acc, ObsPartial *partials, unsigned int *ticket, EmcObs *out,
                                            const unsigned int bidx, const unsigned int nblk, const bool reset_ticket)
{
    __shared__ ObsPartial warp_part[32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    double d[5];
    long long c[6];
#pragma unroll
    d[0] = acc.template sum<0>(); d[1] = acc.template sum<1>(); d[2] = acc.template sum<2>();
    d[3] = acc.template sum<3>(); d[4] = acc.template sum<4>();
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = acc.counter(k);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k) d[k] += __shfl_down_sync(0xffffffffu, d[k], off);
#pragma unroll
        for (int k = 0; k < 6; ++k) c[k] += __shfl_down_sync(0xffffffffu, c[k], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) warp_part[warp].d[k] = d[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) warp_part[warp].c[k] = c[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ObsPartial t = warp_part[0];
        for (int w = 1; w < nwarp; ++w) {
            for (int k = 0; k < 5; ++k) t.d[k] += warp_part[w].d[k];
            for (int k = 0; k < 6; ++k) t.c[k] += warp_part[w].c[k];
        }
        partials[bidx] = t;
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        is_last = (done == nblk - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed-order fold: thread k owns value k (11 values) and walks the blocks in index order
    if (threadIdx.x < 11) {
        const int k = threadIdx.x;
        if (k < 5) {
            double s = 0.0;
            for (unsigned int b = 0; b < nblk; ++b) s += __ldcg(&partials[b].d[k]);
            (&out->sum_z)[k] = s;
        } else {
            long long s = 0;
            for (unsigned int b = 0; b < nblk; ++b) s += __ldcg(&partials[b].c[k - 5]);
            (&out->n_valley[0])[k - 5] = s;      // n_valley[3], n_fault, n_segments, n_events are contiguous
        }
    }
    if (threadIdx.x == 0 && reset_ticket) *ticket = 0u;
}

// ------------------------------------------------------------------------------------------
// NGP charge assignment of one particle (poisson_.py:30-38); `add(cell)` commits one count
// ------------------------------------------------------------------------------------------
template <class Add>
__device__ __forceinline__ void deposit_ngp_particle(const DevConst &P, double x, double y, double z, Add add)
{
    int hx[3], hy[3], hz[3];
    const int nx = axis_hits(x, P.seg[0], P.dl[0], P.inv_dl[0], hx);
    const int ny = axis_hits(y, P.seg[1], P.dl[1], P.inv_dl[1], hy);
    const int nz = axis_hits(z, P.seg[2], P.dl[2], P.inv_dl[2], hz);
    for (int a = 0; a < nx; ++a)
        for (int b = 0; b < ny; ++b)
            for (int c = 0; c < nz; ++c)
                add(((int64_t)hx[a] * P.seg[1] + hy[b]) * P.seg[2] + hz[c]);
}

#ifdef EMC_WARP_END_TIMES
// Debug builds only (profiles/warp_end_times.py): when every warp of the last flight_scatter_kernel launch entered and
// left its loop (%globaltimer, ns).  Not part of the C-ABI of the product library.
__device__ unsigned long long g_warp_t0[8192], g_warp_t1[8192];
}  // namespace emc
extern "C" int emc_debug_warp_times(unsigned long long *t0, unsigned long long *t1, int n)
{
    if (n > 8192) n = 8192;
    if (cudaMemcpyFromSymbol(t0, emc::g_warp_t0, sizeof(unsigned long long) * n) != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(t1, emc::g_warp_t1, sizeof(unsigned long long) * n) != cudaSuccess) return 1;
    return 0;
}
namespace emc {
#endif

// ------------------------------------------------------------------------------------------
// the timestep kernel
//
// Work decomposition.  Events per particle-timestep are Poisson(Gamma*dt = 0.8): 45 % of the
// particles need no scatter event, 36 % one, the tail is unbounded.  One thread per particle
// therefore runs the (fp64-heavy, ~1400 instruction) event body with a third of the lanes
// active (measured 10/32, profiles/r1a_baseline_flight_ncu.txt).  Instead each WARP owns a
// run of 32-particle batches of the ensemble (a contiguous slice, or every W-th batch: ILV below) and two
// small queues in shared memory, and every pass of its
// loop runs ONE of three phases with (nearly) all 32 lanes busy:
//   refill   load the next 32 particles (fully coalesced 256-B rows per state array), fly the
//            first segment, retire the particles whose free flight outlasts the timestep and
//            push the others into the EVENT queue, compacted with ballot/popc;
//   select   pop 32 particles from the event queue; energy, table row, mechanism choice.
//            Self-scattering (70-85 % of events) is finished on the spot with the short
//            closed form (scatter_self); particles that drew a REAL mechanism are parked in the
//            SCATTER queue with their mechanism, energy and azimuth uniform;
//   scatter  pop 32 particles from the scatter queue: polar-angle sampler, sincos, rotation.
// (The LEAN instantiation's select pass decides "self-scattering or not" from three 8-byte thresholds
// around a single-precision guess of the table index and parks EVERYTHING else -- real mechanisms,
// ambiguous draws, faults -- with E_np, r2 and the azimuth uniform; its scatter pass computes the exact
// index, loads the row and runs the literal event: select_self_thr / scatter_parked in emc_device.cuh.)
// select and scatter share the tail (free-flight draw, flight segment, retirement or push
// back).  So the divergent branches -- event count, self vs. real mechanism -- are turned into
// queue traffic instead of idle lanes.  A particle's trajectory depends only on its own state
// and its own (counter-based) random stream, so the regrouping changes no result.
// ------------------------------------------------------------------------------------------
#ifndef EMC_FLIGHT_MIN_BLOCKS
// resident blocks per SM the register allocator must allow.  Measured on configs[1] (2.4e7
// particles): 512 threads x 1 block 1.47 ms, 256 x 2 1.51 ms, 128 x 4 1.54 ms (all 126 registers,
// 16 warps/SM); 192 x 3 (96 registers, 18 warps, spills) 2.22 ms; 256 x 3 (80 registers) 1.72 ms.
#define EMC_FLIGHT_MIN_BLOCKS 1
#endif
#ifndef EMC_SPEC_SELECT
// 1: free-running select passes run select_self_spec (all dependency chains of a self-scattering
// event issued before the first decision, emc_device.cuh) instead of event_select + scatter_self
#define EMC_SPEC_SELECT 1
#endif
#ifndef EMC_STAGE_REFILL
// 1: cp.async (LDGSTS) the next refill batch into shared memory while the current one is
// processed, so the refill phase never waits on HBM.  Measured with the L2 prefetch one batch
// further ahead: 1.387 -> 1.370 ms per 2.4e7-particle timestep in round r1h; since the event
// body lost 15 % of its instructions (r1n) the staging's own 5 % of issue slots cost more than
// the wait it hides: 1.177 ms staged vs 1.160 ms with plain loads behind the L2 prefetch.
#define EMC_STAGE_REFILL 0
#endif
#ifndef EMC_FIELD_AHEAD
// 1: EMC_FIELD_MESH_PACKED on meshes too large for L1: a refill pass loads the POSITIONS of the
// following batch early and, after its own flight, gathers that batch's 32-byte field records with
// cp.async into per-lane shared-memory slots, so that the next refill does not wait for two
// dependent memory round trips (state, then field).  Measured, 1024 x 1 x 1024 mesh, 1.25e7
// particles (profiles/r1s, r1t): flight+deposit 0.883 -> 0.843 ms.  Variants that lost: the same
// with prefetch hints instead of staging (0.914 ms), staging the event passes' gathers too (no
// gain: they are ~500 instructions ahead of their use anyway), and any of it on a mesh whose
// field fits L1 (diode, 1024 cells: pure overhead) -- hence the run-time switch FlightArgs.field_ahead.
#define EMC_FIELD_AHEAD 1
#endif
#ifndef EMC_OBS_SMEM
// 1: the flight kernel keeps its observable sums in shared memory (ObsAccShared): ten registers
// fewer, for block shapes that have to fit 96 registers (EMC_FLIGHT_THREADS=640)
#define EMC_OBS_SMEM 0
#endif
#if EMC_OBS_SMEM
typedef ObsAccShared FlightObsAcc;
#else
typedef ObsAcc FlightObsAcc;
#endif
// (Tried: leaving x, y, z, dtau of a popped particle in its queue slot until the flight segment --
// eight registers fewer across the event math -- measured slower, 1.152 vs 1.134 ms.)
#ifndef EMC_PREFETCH_BATCHES
// refill batches to prefetch ahead into L2 (0 = off).  Measured: 0 -> 1.475 ms, 1 -> 1.409 ms,
// 2 -> 1.411 ms, 4 -> 1.423 ms per 2.4e7-particle timestep.
#define EMC_PREFETCH_BATCHES 2
#endif
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
#ifndef EMC_BULK_REFILL
// 1: the next refill batch (32 particles: 256 B of each of the seven fp64 rows + 128 B of valleys) is
// brought into a per-warp shared-memory tile by EIGHT bulk async copies issued by one lane
// (cp.async.bulk.shared.global, SASS UBLKCP, completion on an mbarrier) right after the previous
// batch has been read out of the tile -- a whole pass (thousands of cycles) before it is needed.
// The refill phase then reads shared memory instead of waiting for HBM: in profiles/r1w the first
// use of the refill loads was 10.7 % of all stall samples (long scoreboard) although the rows had
// been prefetched into L2, and the per-lane loads plus the prefetch address arithmetic were ~55
// instructions per batch.  Needs 16-byte aligned rows (checked by the launcher, FlightArgs.bulk);
// the ragged last batch of a slice and unaligned callers take the per-lane loads below.
#define EMC_BULK_REFILL 1
#endif
#ifndef EMC_REFILL_TILES
// tiles per warp: 1 = the batch after the current one is in flight; 2 = two batches ahead (the mbarrier wait is
// 3.6 % of the stall samples in profiles/r2x).  Measured: 2 tiles 0.949 ms, 1 tile 0.945 ms (profiles/r2z_ab_*):
// the wait is the try_wait's own latency, not late data; the second tile only costs L1.
#define EMC_REFILL_TILES 1
#endif
#ifndef EMC_REFILL_TENSOR_MAP
// 1: when the seven state rows are the rows of ONE 2-D array (equal, 16-byte aligned pitch -- engine.Ensemble's
// (7, N) tensor), the seven 256-byte row slices of a batch arrive by ONE tensor copy (cp.async.bulk.tensor.2d,
// SASS UTMALDG, box 32 x 7 of a CUtensorMap that launch_flight encodes) instead of seven bulk copies with their
// uniform-datapath address arithmetic: 49 -> ~15 single-lane instructions per refill (profiles/r2ag)
#define EMC_REFILL_TENSOR_MAP 1
#endif
struct alignas(128) RefillTile {             // tensor-copy destinations are 128-byte aligned
    double row[7][32];                       // kx ky kz x y z dtau of the next batch
    int32_t val[32];
    unsigned long long bar;                  // mbarrier: one arrival (the issuing lane) + 1920 bytes
    uint32_t loaded;                         // written by every lane once its loads of the tile have completed (see the refill pass)
};
static_assert(sizeof(RefillTile) % 16 == 0, "RefillTile must keep the next tile 16-byte aligned");
constexpr uint32_t REFILL_TILE_BYTES = 7 * 256 + 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp, chosen by the hardware: code under `if (elect_one())` compiles to
// straight uniform-datapath instructions (an `if (lane == 0)` makes ptxas wrap every bulk copy in a
// uniformisation loop)
__device__ __forceinline__ bool elect_one()
{
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    return leader != 0;
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one lane: arm the tile's mbarrier with the byte count and issue the eight row copies of the batch at `j`
__device__ __forceinline__ void refill_issue(RefillTile &T, const FlightArgs &A, uint32_t j, const CUtensorMap *tmap)
{
    const uint32_t bar = smem_u32(&T.bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(REFILL_TILE_BYTES) : "memory");
    if (EMC_REFILL_TENSOR_MAP && A.bulk == 2) {               // warp-uniform (kernel argument)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(&T.row[0][0])), "l"(tmap), "r"((int32_t)j), "r"(0), "r"(bar) : "memory");
        bulk_g2s(smem_u32(&T.val[0]), A.valley + j, 128u, bar);
        return;
    }
    const uint32_t t0 = smem_u32(&T.row[0][0]);
#pragma unroll
    for (int r = 0; r < 7; ++r) bulk_g2s(t0 + r * 256, (&A.kx)[r] + j, 256u, bar);
    bulk_g2s(smem_u32(&T.val[0]), A.valley + j, 128u, bar);
}

__device__ __forceinline__ void refill_wait(RefillTile &T, uint32_t parity)
{
    const uint32_t bar = smem_u32(&T.bar);
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra WAIT_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int QCAP = 64;                    // a phase runs only if its pushes fit: <= 32 + 32
// A queue entry is four 16-byte pairs (one 128-bit shared-memory access each, consecutive slots:
// conflict-free) instead of seven 8-byte and three 4-byte fields: a pop is 4 loads, a push 4 stores.
struct WarpQueue {
    double2 k01[QCAP];                       // kx, ky
    double2 k2x[QCAP];                       // kz, x
    double2 yz[QCAP];                        // y, z
    double2 tw[QCAP];                        // dte, { slice offset (< 2^29) | valley << 29 , random-stream state }
    int32_t grp[QCAP];                       // EMC_FIELD_GROUPS only: field group of the particle
};
struct WarpQueues {
    WarpQueue ev;                            // particles waiting for their next event
    WarpQueue sc;                            // mid-event particles that drew a real mechanism ...
    double2 sc_er[QCAP];                     // ... with E_np (scatter_.py:1069) and the azimuth uniform
    int32_t sc_mech[QCAP];
#if EMC_STAGE_REFILL
    double stage[7][32];                     // next refill batch, landed here by cp.async (LDGSTS)
    int32_t stage_val[32];
#endif
};
#ifndef EMC_SELF_THRESHOLD
// 1: the lean kernels decide self-scattering from the 8-byte threshold table (Mat::thr,
// select_self_spec<.., THR>) and the scatter pass re-runs a parked event from its counters
// 2: ... from the thresholds of the three rows around a single-precision GUESS of the index
// (select_self_thr): no exact index and no load at the end of a dependency chain in the select pass
#define EMC_SELF_THRESHOLD 2
#endif
// the queues of the threshold form: a parked particle carries nothing but itself (the scatter pass runs the
// event again from its counters) ...
struct WarpQueuesLean {
    WarpQueue ev;
    WarpQueue sc;
};
#ifndef EMC_PARK_FIELDS
// 1: ... or, in the one-step kernel, E_np, r2 and the azimuth uniform of the select pass (ParkFields), so that the
// scatter pass neither evaluates Philox block A nor the energy chain a second time (~150 of its ~1170
// instructions; the scatter pass was 13.6 % of the stall samples in profiles/r2ag)
#define EMC_PARK_FIELDS 1
#endif
struct WarpQueuesPark {
    WarpQueue ev;
    WarpQueue sc;
    double2 sc_er[QCAP];                     // E_np (NaN: energy fault), azimuth uniform
    double sc_r2[QCAP];
};
// (not in the multi-layer kernel: its six more live registers cost that kernel 20 %, 0.697 vs 0.580 ms for the
// flight of configs[2], profiles/r2ap_*)
template <bool LEAN, bool ML> struct FlightQueues {
    typedef WarpQueues type;
};
#if EMC_SELF_THRESHOLD && !EMC_STAGE_REFILL
template <> struct FlightQueues<true, true> {
    typedef WarpQueuesLean type;
};
template <> struct FlightQueues<true, false> {
#if EMC_SELF_THRESHOLD == 2 && EMC_PARK_FIELDS
    typedef WarpQueuesPark type;
#else
    typedef WarpQueuesLean type;
#endif
};
#endif
// One whole event (threshold form of the lean kernels): the scatter pass of the parked particles
// (0.1 passes per batch) -- real mechanisms and the cold exits of the select pass (faults, r2 at or
// above the valley's smallest last column) -- runs scatter_event -- the general, non-speculative
// event: same arithmetic, same counters, same fault order -- from the particle's event counter.
// A CALL keeps its registers (the whole table row, sincos, the samplers) and its ~1500 instructions
// out of the hot loop's allocation and instruction footprint.
struct EventOut {
    double kx, ky, kz, lg;      // k' and log(flight uniform) of the free flight that follows
    int32_t val, f;             // valley after the event; fault code or -1
};
#ifndef EMC_COLD_CALL
// 1: event_cold is a real CALL (smaller loop, fewer registers in the loop's own allocation: the variant
// for more than 16 warps per SM); 0: inlined.  Measured at 512 threads: call 0.995 ms, inlined 0.959 ms
// (the ABI marshalling and the convergence barrier around the call cost 2.5 % more instructions).
#define EMC_COLD_CALL 0
#endif
#if EMC_COLD_CALL
#define EMC_COLD_INLINE __noinline__
#else
#define EMC_COLD_INLINE __forceinline__
#endif
__device__ EMC_COLD_INLINE EventOut event_cold(const DevConst *P, const Mat *MT, const MechMeta *M, double kx, double ky,
                                            double kz, int32_t val, uint64_t seed, uint64_t pid, uint32_t step, uint32_t ev)
{
    PhiloxStream rng;
    rng.init(seed, pid, step, ev);
    EventOut o;
    int v = val;
    o.f = scatter_event<false>(*P, *MT, *M, kx, ky, kz, v, rng);
    o.lg = fast_log(rng.free_flight());
    o.kx = kx; o.ky = ky; o.kz = kz; o.val = v;
    return o;
}
// EMC_FIELD_MESH_PACKED with field_ahead only (kept out of WarpQueues: shared memory the other modes
// do not need would shrink their L1 -- the table-row gathers live on its hit rate: 2 KB per warp
// more cost the configs[1] kernel 3.4 % in r1t)
struct FieldStage {
    double2 rs_a[32], rs_b[32];              // (ex, ey), (ez, 0) of each lane's cell in the NEXT refill batch
};

#if EMC_STAGE_REFILL
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc));
}
// each lane stages ITS OWN particle of the batch: slot [lane] is written and later read by the
// same thread only, so the per-thread cp.async group semantics are all the ordering needed
__device__ __forceinline__ void stage_batch(WarpQueues &WQ, const FlightArgs &A, int64_t j, int64_t wend, int lane)
{
    if (j < wend) {
        cp_async8(&WQ.stage[0][lane], A.kx + j); cp_async8(&WQ.stage[1][lane], A.ky + j);
        cp_async8(&WQ.stage[2][lane], A.kz + j); cp_async8(&WQ.stage[3][lane], A.x + j);
        cp_async8(&WQ.stage[4][lane], A.y + j); cp_async8(&WQ.stage[5][lane], A.z + j);
        cp_async8(&WQ.stage[6][lane], A.dtau + j); cp_async4(&WQ.stage_val[lane], A.valley + j);
    }
    asm volatile("cp.async.commit_group;");
}
#endif

struct Particle {
    double kx, ky, kz, x, y, z, dte;
    int32_t val;
    uint32_t off;
    int32_t grp;
};

// padded-mesh index of the cell that holds the particle (nearest-cell gather, EMC_FIELD_MESH*)
__device__ __forceinline__ int64_t mesh_cell(const DevConst &P, const Particle &p)
{
    const int ci = axis_cell(p.x, P.seg[0], P.dl[0], P.inv_dl[0]) + 1;
    // a single-cell axis (the reference's ny = 1 meshes): the clamp sends every y to cell 0
    const int cj = P.seg[1] == 1 ? 1 : axis_cell(p.y, P.seg[1], P.dl[1], P.inv_dl[1]) + 1;
    const int ck = axis_cell(p.z, P.seg[2], P.dl[2], P.inv_dl[2]) + 1;
    return ((int64_t)ci * (P.seg[1] + 2) + cj) * (P.seg[2] + 2) + ck;
}

// EMC_FIELD_MESH_XZ (ny = 1 meshes): index of the particle's cell in the UNPADDED (nx, nz) record array
__device__ __forceinline__ int64_t mesh_cell_xz(const DevConst &P, const Particle &p)
{
    const int ci = axis_cell(p.x, P.seg[0], P.dl[0], P.inv_dl[0]);
    const int ck = axis_cell(p.z, P.seg[2], P.dl[2], P.inv_dl[2]);
    return (int64_t)ci * P.seg[2] + ck;
}
template <int FIELD>
__device__ __forceinline__ int64_t mesh_cell_of(const DevConst &P, const Particle &p)
{
    return FIELD == EMC_FIELD_MESH_XZ ? mesh_cell_xz(P, p) : mesh_cell(P, p);
}

// The field a flight segment sees depends only on where the segment STARTS, and a scatter event
// does not move the particle: so the event passes know the cell as soon as a particle is popped
// and pull its field record towards L1 ~500 instructions before the segment needs it.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc));
}

template <int FIELD>
__device__ __forceinline__ void field_prefetch(const FlightArgs &A, int64_t c)
{
    if (FIELD == EMC_FIELD_MESH_XZ) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.ex + 2 * c));
    } else if (FIELD == EMC_FIELD_MESH_PACKED) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.ex + 4 * c));
    } else if (FIELD == EMC_FIELD_MESH) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.ex + c));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.ey + c));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.ez + c));
    }
}

// each lane stages ITS OWN record and reads it back itself: the per-thread cp.async group
// semantics are all the ordering needed
template <int FIELD>
__device__ __forceinline__ void field_stage(const FlightArgs &A, int64_t c, FieldStage &FS, int lane)
{
    if (FIELD == EMC_FIELD_MESH_XZ) {             // one 16-byte (ex, ez) record
        cp_async16(&FS.rs_a[lane], reinterpret_cast<const double2 *>(A.ex) + c);
    } else {
        const double2 *rec = reinterpret_cast<const double2 *>(A.ex) + 2 * c;
        cp_async16(&FS.rs_a[lane], rec);
        cp_async16(&FS.rs_b[lane], rec + 1);
    }
    asm volatile("cp.async.commit_group;");
}

template <int FIELD, bool ZON = false>
__device__ __forceinline__ void field_at(const DevConst &P, const FlightArgs &A, const Particle &p, int64_t c,
                                         double &efx, double &efy, double &efz, FieldStage &FS, int lane,
                                         bool staged = false)
{
    if (FIELD == EMC_FIELD_CONSTANT) {
        efx = P.efield[0]; efy = P.efield[1]; efz = P.efield[2];
    } else if (FIELD == EMC_FIELD_GROUPS) {
        efz = __ldg(P.field_groups + 3 * p.grp + 2);
        if (ZON || P.zonly) { efx = 0.0; efy = 0.0; }
        else { efx = __ldg(P.field_groups + 3 * p.grp); efy = __ldg(P.field_groups + 3 * p.grp + 1); }
    } else {
        // nearest-cell gather from the padded (nx+2,ny+2,nz+2) field of emc_field, V/m -> V/cm
        if (FIELD == EMC_FIELD_MESH_XZ) {         // one 16-B (ex, ez) record per interior cell, V/cm; ey is -0 on ny = 1
            const double2 a = staged ? FS.rs_a[lane] : __ldg(reinterpret_cast<const double2 *>(A.ex) + c);
            efx = a.x; efy = 0.0; efz = a.y;
        } else if (FIELD == EMC_FIELD_MESH_PACKED) {     // one 32-B record per cell, already V/cm
            double2 a, b;
            if (staged) {                         // landed: the caller waited for the cp.async group
                a = FS.rs_a[lane]; b = FS.rs_b[lane];
            } else {
                const double2 *rec = reinterpret_cast<const double2 *>(A.ex) + 2 * c;
                a = __ldg(rec); b = __ldg(rec + 1);
            }
            efx = a.x; efy = a.y; efz = b.x;
        } else {
            efx = __ldg(A.ex + c) * 1e-2;
            efy = __ldg(A.ey + c) * 1e-2;
            efz = __ldg(A.ez + c) * 1e-2;
        }
    }
}

// retire a particle: write its state back, add it to the observables and the charge mesh
template <bool REPLAY, bool DEPOSIT, bool ML>
__device__ __forceinline__ void retire(const DevConst &P, const FlightArgs &A, uint32_t i, const Particle &p,
                                       uint32_t cursor, FlightObsAcc &acc, unsigned int *smesh)
{
    A.kx[i] = p.kx; A.ky[i] = p.ky; A.kz[i] = p.kz;
    A.x[i] = p.x; A.y[i] = p.y; A.z[i] = p.z;
    A.dtau[i] = p.dte;
    int val = p.val, layer = 0;
    if (ML && val >= 0) {                         // main_.py:393: the layer of the final position
        layer = where_am_i(P, p.z);
        if (layer < 0) { val = -(1 + EMC_FAULT_OUTSIDE); layer = 0; }
    }
    A.valley[i] = val;
    if (REPLAY) A.cursor[i] = (int32_t)cursor;
    if (A.obs && !A.obs_before) obs_particle(P, P.mat[layer], acc, p.kx, p.ky, p.kz, p.z, val);
    if (DEPOSIT) {
        if (A.smem_cells > 0)
            deposit_ngp_particle(P, p.x, p.y, p.z, [&](int64_t c) { atomicAdd(&smesh[c], 1u); });
        else
            deposit_ngp_particle(P, p.x, p.y, p.z, [&](int64_t c) {
                atomicAdd(reinterpret_cast<unsigned long long *>(A.mesh) + c, 1ull); });
    }
}

template <int FIELD>
__device__ __forceinline__ void queue_push(WarpQueue &Q, int slot, const Particle &p, uint32_t s0)
{
    Q.k01[slot] = make_double2(p.kx, p.ky);
    Q.k2x[slot] = make_double2(p.kz, p.x);
    Q.yz[slot] = make_double2(p.y, p.z);
    // only live particles (valley 0..2) are queued
    Q.tw[slot] = make_double2(p.dte, __hiloint2double((int)s0, (int)(p.off | ((uint32_t)p.val << 29))));
    if (FIELD == EMC_FIELD_GROUPS) Q.grp[slot] = p.grp;
}

#ifndef EMC_INTERLEAVE_BATCHES
// 1: the lean constant / group-field kernels give a warp every W-th batch of the ensemble instead of a contiguous
// run (see ILV in flight_scatter_kernel): 0.911 -> 0.889 ms on configs[1], profiles/r2au_ab_*.
// 2: the lean mesh-field kernels, too -- measured 0.7 % slower (0.647 vs 0.6425 ms on 1024 x 1 x 1024,
// profiles/r2aw_ab_*): their plain loads behind an L2 prefetch like the contiguous run better.
#define EMC_INTERLEAVE_BATCHES 1
#endif
#ifndef EMC_INTERLEAVE_STEPS_KERNEL
#define EMC_INTERLEAVE_STEPS_KERNEL 0
#endif
#ifndef EMC_LANE_IN_REGISTER
#define EMC_LANE_IN_REGISTER 1
#endif
#ifndef EMC_EARLY_GROUP_FIELD
// 1 (lean kernel, EMC_FIELD_GROUPS): the particle's field value is loaded as soon as its group is known
// (refill / queue pop) instead of at the head of the flight segment, where its L1/L2 latency was 2.6 %
// of the stall samples (profiles/r2f).  Measured in r2o: no gain (0.991 vs 0.985 ms); measured again on the
// final kernel, whose flight no longer hides behind a long tail: -0.7 %, and -1.8 % together with
// EMC_PARK_UNCOND (0.912 vs 0.929 ms, profiles/r2ar_ab_*): on.
#define EMC_EARLY_GROUP_FIELD 1
#endif
#ifndef EMC_LEAN_KERNEL
#define EMC_LEAN_KERNEL 1
#endif
#ifndef EMC_REFILL_FLIES
// (2: in the mesh-field modes, too -- measured: no gain there, 0.631 vs 0.629 ms on 1024 x 1 x 1024, profiles/r2ae_ab_*:
// the refill of those modes waits for its field record, not for arithmetic)
#define EMC_REFILL_FLIES 1
#endif
#ifndef EMC_PUSH_FIRST
#define EMC_PUSH_FIRST 1
#endif
#ifndef EMC_PARK_UNCOND
// 1: the (predicated) push into the scatter queue is issued in every pass instead of behind a branch on the phase
// (one ballot, two POPCs and five predicated-off stores more in the refill and scatter passes, one BSSY / BRA /
// BSYNC region less in every pass: -0.7 %, profiles/r2ar_ab_*)
#define EMC_PARK_UNCOND 1
#endif
#ifndef EMC_PREDICATED_TAIL
// 1: the write-back of a finished particle and the queue pushes are predicated stores instead of
// branches around them.  The loop tail ran at 130 stall samples per instruction against 43 in the
// select block (profiles/r1w source page): BSSY / BRA / BSYNC triples around a handful of stores,
// each with its branch-resolution bubble, on a kernel that is bound by per-warp latency.
#define EMC_PREDICATED_TAIL 1
#endif
// the same push as predicated shared-memory stores (no branch around them); `on` is the lane's predicate
template <int FIELD>
__device__ __forceinline__ void queue_push_if(WarpQueue &Q, int slot, const Particle &p, uint32_t s0, bool on)
{
    const uint32_t a = smem_u32(&Q.k01[0]) + (uint32_t)slot * 16u;
    const double w = __hiloint2double((int)s0, (int)(p.off | ((uint32_t)p.val << 29)));
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t"
                 "@q st.shared.v2.f64 [%0], {%1, %2};\n\t"
                 "@q st.shared.v2.f64 [%0+1024], {%3, %4};\n\t"
                 "@q st.shared.v2.f64 [%0+2048], {%5, %6};\n\t"
                 "@q st.shared.v2.f64 [%0+3072], {%7, %8};\n\t}"
                 ::"r"(a), "d"(p.kx), "d"(p.ky), "d"(p.kz), "d"(p.x), "d"(p.y), "d"(p.z), "d"(p.dte), "d"(w),
                 "r"((uint32_t)on) : "memory");
    if (FIELD == EMC_FIELD_GROUPS) {
        const uint32_t g = smem_u32(&Q.grp[0]) + (uint32_t)slot * 4u;
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}"
                     ::"r"(g), "r"(p.grp), "r"((uint32_t)on) : "memory");
    }
}
static_assert(offsetof(WarpQueue, k2x) == 1024 && offsetof(WarpQueue, yz) == 2048 && offsetof(WarpQueue, tw) == 3072,
              "queue_push_if hard-codes the WarpQueue layout");

// write-back of a finished particle as eight predicated global stores
__device__ __forceinline__ void retire_stores_if(const FlightArgs &A, uint32_t i, const Particle &p, int val, bool on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %16, 0;\n\t"
                 "@q st.global.f64 [%0], %8;\n\t@q st.global.f64 [%1], %9;\n\t@q st.global.f64 [%2], %10;\n\t"
                 "@q st.global.f64 [%3], %11;\n\t@q st.global.f64 [%4], %12;\n\t@q st.global.f64 [%5], %13;\n\t"
                 "@q st.global.f64 [%6], %14;\n\t@q st.global.u32 [%7], %15;\n\t}"
                 ::"l"(A.kx + i), "l"(A.ky + i), "l"(A.kz + i), "l"(A.x + i), "l"(A.y + i), "l"(A.z + i), "l"(A.dtau + i),
                 "l"(A.valley + i), "d"(p.kx), "d"(p.ky), "d"(p.kz), "d"(p.x), "d"(p.y), "d"(p.z), "d"(p.dte), "r"(val),
                 "r"((uint32_t)on) : "memory");
}

template <int FIELD>
__device__ __forceinline__ void queue_pop(const WarpQueue &Q, int slot, Particle &p, uint32_t &s0)
{
    const double2 a = Q.k01[slot], b = Q.k2x[slot], c = Q.yz[slot], d = Q.tw[slot];
    p.kx = a.x; p.ky = a.y; p.kz = b.x; p.x = b.y; p.y = c.x; p.z = c.y; p.dte = d.x;
    const uint32_t w = (uint32_t)__double2loint(d.y);
    p.off = w & 0x1fffffffu;
    p.val = (int32_t)(w >> 29);
    s0 = (uint32_t)__double2hiint(d.y);
    p.grp = (FIELD == EMC_FIELD_GROUPS) ? Q.grp[slot] : 0;
}

// ML = multi-layer device (emc_set_layers): the layer of a particle is looked up from its z with
// where_am_i before every drift, scatter and observable (main_.py:366,372,377,387,393) and selects
// the material constants, the table and the mechanism metadata.  ML = false compiles to the
// single-material code (layer 0 everywhere).
#ifdef EMC_FLIGHT_MAXNREG
#define EMC_FLIGHT_BOUNDS __maxnreg__(EMC_FLIGHT_MAXNREG)     // tuning: explicit register cap instead of launch bounds
#else
#define EMC_FLIGHT_BOUNDS __launch_bounds__(FLIGHT_THREADS, EMC_FLIGHT_MIN_BLOCKS)
#endif
// LEAN: the instantiation for the common production launch -- free-running Philox, default libm
// composition, no fused deposit, observables (if any) in the before-step phase, padded tables whose last
// column is self-scattering, and for the constant / group field modes Ex = Ey = +0 -- with every
// alternative of those run-time switches compiled OUT.  The kernel is sensitive to the size of its
// loop (three copies of the flight code cost 50 %, profiles/ab_r2g): 270 instructions and 9 registers
// fewer are 2.4 % (profiles/ab_r2h).  launch_flight picks it whenever the conditions hold.
// CHAIN (lean constant / group-field kernels): the block does not run ONE timestep of slice blockIdx.x but pops
// (slice, timestep) work items from the FIFO of a timestep chain (ChainCtl, emc_host.h) until all
// A.chain_steps x A.chain_slices of them are done -- K timesteps WITH their observable rows in one persistent
// launch.  Same slices, same arithmetic, same Philox counters and the same fixed-order fold of the observable
// partials (by slice) as K launches: bit-identical state and rows.
template <class Stream, int FIELD, bool EXACT, bool DEPOSIT, bool ML, bool LEAN = false, bool CHAIN = false>
__global__ void EMC_FLIGHT_BOUNDS
flight_scatter_kernel(const __grid_constant__ DevConst P, const __grid_constant__ MechMeta Mg,
                      const __grid_constant__ FlightArgs A, const __grid_constant__ CUtensorMap tmap)
{
    constexpr bool REPLAY = Stream::is_replay;
    static_assert(!CHAIN || (LEAN && !ML && (FIELD == EMC_FIELD_CONSTANT || FIELD == EMC_FIELD_GROUPS)),
                  "timestep chains: lean single-layer constant / group-field kernels only");
    constexpr bool ZON = LEAN && (FIELD == EMC_FIELD_CONSTANT || FIELD == EMC_FIELD_GROUPS);   // zonly known at compile time
    static_assert(!LEAN || (!REPLAY && !EXACT && !DEPOSIT), "LEAN is the Philox / default-libm / no-deposit kernel");
    constexpr bool THR = LEAN && EMC_SELF_THRESHOLD && !EMC_STAGE_REFILL;
    using WarpQueues = typename FlightQueues<LEAN, ML>::type;
    constexpr bool PARKF = THR && EMC_SELF_THRESHOLD == 2 && EMC_PARK_FIELDS && !ML;
    __shared__ MechMeta Ms[ML ? EMC_MAX_LAYERS : 1];
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    WarpQueues *queues = reinterpret_cast<WarpQueues *>(dyn_smem);
    constexpr bool AHEAD_OK = EMC_FIELD_AHEAD && (FIELD == EMC_FIELD_MESH_PACKED || FIELD == EMC_FIELD_MESH_XZ);
    const bool AHEAD = AHEAD_OK && A.field_ahead;                 // warp-uniform (kernel argument)
    const size_t fs_bytes = AHEAD ? sizeof(FieldStage) : 0;
    FieldStage *fstages = reinterpret_cast<FieldStage *>(dyn_smem + sizeof(WarpQueues) * (FLIGHT_THREADS / 32));
    constexpr size_t OBS_BYTES = EMC_OBS_SMEM ? sizeof(double) * 5 * FLIGHT_THREADS : 0;
    const bool BULK = EMC_BULK_REFILL && A.bulk;                  // warp-uniform (kernel argument)
    const size_t rt_bytes = BULK ? sizeof(RefillTile) * EMC_REFILL_TILES : 0;
    RefillTile *rtiles = reinterpret_cast<RefillTile *>(dyn_smem + (sizeof(WarpQueues) + fs_bytes) * (FLIGHT_THREADS / 32));
    double *obs_columns = reinterpret_cast<double *>(dyn_smem + (sizeof(WarpQueues) + fs_bytes + rt_bytes) * (FLIGHT_THREADS / 32));
    unsigned int *smesh = reinterpret_cast<unsigned int *>(dyn_smem + (sizeof(WarpQueues) + fs_bytes + rt_bytes) * (FLIGHT_THREADS / 32) + OBS_BYTES);
    {
        const int nw = sizeof(MechMeta) / 4;
        uint32_t *dst = reinterpret_cast<uint32_t *>(&Ms[0]);
        if (ML) {                                 // per-layer metadata lives in device memory
            const uint32_t *src = reinterpret_cast<const uint32_t *>(A.mech_layers);
            for (int t = threadIdx.x; t < nw * P.n_layers; t += blockDim.x) dst[t] = __ldg(src + t);
        } else {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&Mg);
            for (int t = threadIdx.x; t < nw; t += blockDim.x) dst[t] = src[t];
        }
        for (int t = threadIdx.x; t < A.smem_cells; t += blockDim.x) smesh[t] = 0u;
    }
    __syncthreads();

    uint32_t rt_count = 0;                                        // bulk batches consumed so far (mbarrier phase: across work items)
    [[maybe_unused]] bool chain_first = true;
    [[maybe_unused]] __shared__ unsigned long long chain_item_s;
    for (;;) {                                                    // CHAIN: one work item per pass; otherwise a single pass
    unsigned int bid = blockIdx.x, nblk = gridDim.x;              // the block's slice and the number of slices
    uint32_t step_now = A.step;
    EmcObs *obs_out = A.obs;
    ObsPartial *obs_parts = A.partials;
    unsigned int *obs_ticket = A.ticket;
    [[maybe_unused]] uint32_t chain_j = 0;
    if constexpr (CHAIN) {
        if (threadIdx.x == 0) {
            ChainCtl *C = A.chain;
            const unsigned int total = (unsigned int)A.chain_steps * (unsigned int)A.chain_slices;
            const unsigned int T = atomicAdd(&C->popped, 1u);
            unsigned long long item;
            if (T >= total) {
                item = ~0ull;                                     // every item has an owner: done
            } else if (T < (unsigned int)A.chain_slices) {
                item = (unsigned long long)T;                     // timestep 0 of slice T: the state as the caller left it
            } else {
                // item T is pushed by the block that finishes the timestep before it on that slice: a block with a
                // smaller ticket, and a ticket is only ever held by a block that is running -- no deadlock, whatever
                // part of the grid is resident.  The slot's tag (T + 1) tells this item from the ones that use the
                // slot CHAIN_RING items earlier or later; at most `slices` items are in flight at a time (<< CHAIN_RING).
                const unsigned long long *slot = &C->ready[T % CHAIN_RING];
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(item) : "l"(slot) : "memory");
                    if ((unsigned int)(item >> 32) == T + 1u) break;
                    __nanosleep(64);
                }
                // the slot is free again (a pusher waits for an empty slot: tests/test_chain_protocol_model.py)
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slot), "l"(0ull) : "memory");
                item &= 0xffffffffull;
                __threadfence();
            }
            chain_item_s = item;
        }
        __syncthreads();
        const unsigned long long item = chain_item_s;
        if (item == ~0ull) break;
        // the slice was written by another SM's ordinary stores and is read here by the copy engine (async proxy)
        asm volatile("fence.proxy.async;" ::: "memory");
        bid = (unsigned int)__shfl_sync(0xffffffffu, (int)(item & 0xffffu), 0);
        chain_j = (uint32_t)__shfl_sync(0xffffffffu, (int)((item >> 16) & 0xffffu), 0);
        nblk = (unsigned int)A.chain_slices;
        step_now = A.step + chain_j;
        obs_out = A.obs + chain_j;
        obs_parts = A.partials + (size_t)chain_j * nblk;
        obs_ticket = &A.chain->step_done[chain_j];
    }

    FlightObsAcc acc;
    acc.zero(obs_columns + threadIdx.x);

    const unsigned FULL = 0xffffffffu;
    // the warp index through a full-mask shuffle from lane 0: the compiler then KNOWS it is warp-uniform
    // and keeps everything derived from it (slice bounds, queue and tile addresses, the operands of the
    // bulk copies) in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
#if EMC_LANE_IN_REGISTER
    // lane and the lower-lanes mask read ONCE from their special registers and pinned in two general
    // registers (volatile asm is not rematerialised): the compiler otherwise re-reads SR_TID.X in every
    // pass of the loop -- an S2R (~20 cycles) at the head of each pass, 1.3 % of the stall samples in r2f
    int lane;
    unsigned lt_mask;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
#else
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
#endif
    WarpQueues &WQ = queues[warp];
    FieldStage &FS = fstages[AHEAD ? warp : 0];                   // only dereferenced when AHEAD
    const int64_t total_warps = (int64_t)nblk * (FLIGHT_THREADS / 32);
    const int64_t gw = (int64_t)bid * (FLIGHT_THREADS / 32) + warp;
    // ILV (lean constant / group-field kernels): a warp's batches are interleaved with the other warps' -- batch
    // gw, gw + W, gw + 2 W ... of the ensemble (W = all warps of the grid) -- instead of one contiguous run.  With
    // field groups a contiguous run lies inside ONE group, and the groups differ in work (a 50 kV/cm group has
    // several times the real scattering of a 0.1 kV/cm one): the blocks of configs[1] finished 3.3 % apart (ncu
    // r2z: sm__cycles_active min / max).  Interleaved, every warp samples every group.  Batches stay 32 consecutive
    // particles (the same 256-byte row slices), and at any moment the grid works on one contiguous window of the
    // ensemble.  A particle's offset in its warp's run (`next` + lane, p.off) -> its index: w0 + (off / 32) * 32 W + off % 32.
    constexpr bool ILV = EMC_INTERLEAVE_BATCHES && (ZON || (EMC_INTERLEAVE_BATCHES > 1 && LEAN));   // 2: the lean mesh-field kernels, too (A/B)
    int64_t wstart, wend;
    uint32_t wlen_;
    if (ILV) {
        const int64_t nb = (A.n + 31) >> 5;                       // batches of the ensemble, the last one may be ragged
        const int64_t mine = gw < nb ? (nb - gw + total_warps - 1) / total_warps : 0;
        wstart = mine ? gw * 32 : 0;
        wlen_ = (uint32_t)(mine * 32);
        if (mine && gw + (mine - 1) * total_warps == nb - 1) wlen_ -= (uint32_t)(nb * 32 - A.n);
        wend = 0;                                                 // (only the disabled staging variant reads it)
    } else {
        const int64_t slice = (((A.n + total_warps - 1) / total_warps) + 31) & ~(int64_t)31;
        wstart = gw * slice;
        if (wstart > A.n) wstart = A.n;
        wend = (wstart + slice < A.n) ? wstart + slice : A.n;
        wlen_ = (uint32_t)(wend - wstart);
    }
    const uint32_t Wp = ILV ? (uint32_t)total_warps : 1u;         // stride of the warp's batches, in batches
    // loop state in 32 bits (a warp's run is shorter than 2^29 particles: the queues pack the
    // valley above the offset): `next` and the length of the run relative to wstart
    uint32_t next = 0;
    const uint32_t wlen = wlen_;
    const uint32_t w0 = (uint32_t)wstart;
    RefillTile *RTW = rtiles + (BULK ? warp * EMC_REFILL_TILES : 0);      // only dereferenced when BULK
    // the full batches of the slice are issued in order, EMC_REFILL_TILES ahead, into the tiles in turn:
    // batch b uses tile b % TILES and the mbarrier phase (b / TILES) & 1
    if (BULK) {
        if ((!CHAIN || chain_first) && elect_one()) {
#pragma unroll
            for (int t = 0; t < EMC_REFILL_TILES; ++t)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&RTW[t].bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        chain_first = false;
        __syncwarp();
        if (elect_one()) {
#pragma unroll
            for (int t = 0; t < EMC_REFILL_TILES; ++t)
                if (wlen >= 32u * (t + 1)) refill_issue(RTW[t], A, w0 + 32u * t * Wp, &tmap);
        }
    }
#if EMC_STAGE_REFILL
    stage_batch(WQ, A, wstart + next + lane, wend, lane);
#endif
#ifdef EMC_WARP_END_TIMES
    if (LEAN && lane == 0 && gw < 8192) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_warp_t0[gw] = t;
    }
#endif
    int qn = 0, qsc = 0;                          // lengths of the event / scatter queues (warp-uniform)
    // EMC_FIELD_MESH_PACKED: a refill pass gathers the field record of the FOLLOWING batch (whose
    // positions it loads early and turns into a cell after its own flight), so that a refill never
    // waits for two dependent memory round trips (state, then field)

    bool field_staged = false;                    // the batch at `next` has its field record in the rs slots
    const double dt = P.dt;

    // EMC_FIELD_GROUPS: group index / remainder of the slice's next particle, advanced without
    // a 64-bit division per particle
    // (group_size < 2^31 and n_groups < 2^31: checked by emc_set_field_groups)
    int32_t grp_next = 0;
    uint32_t grp_rem = 0;
    const uint32_t gsz = (uint32_t)P.group_size;
    [[maybe_unused]] uint32_t grp_dq = 0, grp_dr = 32;            // one batch further on: groups / remainder to add
    if (FIELD == EMC_FIELD_GROUPS) {
        const uint64_t g0 = A.pid_offset + (uint64_t)wstart;
        const uint64_t gq = g0 / (uint64_t)P.group_size;
        grp_next = (int32_t)(gq < (uint64_t)P.n_groups ? gq : (uint64_t)P.n_groups);
        grp_rem = (uint32_t)(g0 % (uint64_t)P.group_size);
        if (ILV) {                                                // the warp's next batch is 32 W particles further on
            const uint64_t delta = 32ull * (uint64_t)total_warps;
            const uint64_t dq = delta / (uint64_t)P.group_size;
            grp_dq = (uint32_t)(dq < 0x7fffffffull ? dq : 0x7fffffffull);
            grp_dr = (uint32_t)(delta % (uint64_t)P.group_size);
        }
    }

    // The phases share ONE copy of the event tail, the flight segment, the retirement and the
    // queue pushes (the kernel has to stay inside the 32 KB instruction cache).
    enum { PH_REFILL = 0, PH_SELECT = 1, PH_SCATTER = 2, PH_DONE = 3 };
    // warp-uniform phase choice: a full scatter or select pass if one is possible, else
    // refill, else drain whatever is left.  A pass pops only what its pushes can hold.
    int n_sc = 0, n_sel = 0;
    auto choose_phase = [&]() -> int {
        const int room_ev = QCAP - qn, room_sc = QCAP - qsc;
        n_sc = min(32, min(qsc, room_ev));                        // scatter pushes back into ev
        n_sel = min(32, min(qn, room_sc));                        // select may push all into sc
        if (n_sc == 32) return PH_SCATTER;
        if (n_sel == 32) return PH_SELECT;
        if (qn < 32 && next < wlen) return PH_REFILL;
        if (n_sel > 0 && n_sel >= n_sc) return PH_SELECT;
        if (n_sc > 0) return PH_SCATTER;
        return PH_DONE;
    };
#ifndef EMC_ROTATE_LOOP
// 1: the phase of the NEXT pass is chosen at the end of a pass, so that the back edge can jump straight
// into that phase's block (one far taken branch per pass instead of two).  Measured: no gain (0.974 vs
// 0.969 ms, profiles/r2v_ab_*); off.
#define EMC_ROTATE_LOOP 0
#endif
#if EMC_ROTATE_LOOP
    int phase = choose_phase();
    while (phase != PH_DONE) {
#else
    while (true) {
        const int phase = choose_phase();
        if (phase == PH_DONE) break;
#endif
        Particle p;
        Stream rng;
        uint32_t s0 = 0;
        // particle index inside the launch in 32 bits (launch_flight refuses n >= 2^32 - 64): every
        // state access is then ONE IMAD.WIDE.U32 off the row pointer in the constant bank instead of a
        // 64-bit add per row
        uint32_t i = 0;
        int mech = 0;
        double e_np = 0.0, r_az = 0.0;
        ParkFields pf = {0.0, 0.0, 0.0};                          // PARKF: what a parked particle carries to its scatter pass
        bool act, pend = false, fly = false, park = false, tail = false;
        constexpr bool MESHF = (FIELD == EMC_FIELD_MESH || FIELD == EMC_FIELD_MESH_PACKED || FIELD == EMC_FIELD_MESH_XZ);
        constexpr bool SPEC = EMC_SPEC_SELECT && !REPLAY && !EXACT;
        double lg = 0.0;                                          // SPEC: log(flight uniform) of a finished event
        constexpr bool EARLY_GF = EMC_EARLY_GROUP_FIELD && ZON && FIELD == EMC_FIELD_GROUPS;
        double gfz = 0.0;                                         // EARLY_GF: Ez of the particle's field group
        int64_t cell = 0;                                         // mesh cell of the coming flight segment
        int layer = 0;                                            // ML: where_am_i of the particle's z
        double dt_seg = 0.0;
        double nx_ = 0.0, ny_ = 0.0, nz_ = 0.0;                   // AHEAD: position of the lane's particle in the next batch
        bool ahead = false;
        // one flight segment of the lane's particle (main_.py:367 / :373 / :388)
        auto flight_segment = [&](bool staged) {
            double efx, efy, efz;
            if (EARLY_GF) { efx = 0.0; efy = 0.0; efz = gfz; }
            else field_at<FIELD, ZON>(P, A, p, cell, efx, efy, efz, FS, lane, staged);
            const Mat &MT = P.mat[ML ? layer : 0];
            const double dtm = div_const(dt_seg, MT.mass[p.val], MT.r_mass[p.val]);
            if (FIELD == EMC_FIELD_MESH_XZ)
                carrier_drift<2>(P, p.kx, p.ky, p.kz, p.x, p.y, p.z, efx, efy, efz, dt_seg, dtm);
            else if ((FIELD == EMC_FIELD_CONSTANT || FIELD == EMC_FIELD_GROUPS) && (ZON || P.zonly))
                carrier_drift<1>(P, p.kx, p.ky, p.kz, p.x, p.y, p.z, efx, efy, efz, dt_seg, dtm);
            else
                carrier_drift<0>(P, p.kx, p.ky, p.kz, p.x, p.y, p.z, efx, efy, efz, dt_seg, dtm);
            // LEAN: every live particle of the ensemble as loaded flies one first segment and every event is
            // followed by one more (a Philox stream never overflows): n_segments = n_valley[0..2] of the
            // before-step observables + n_events, added up after the loop -- one loop-carried register less
            if (!LEAN) acc.count_segment();
        };
        // REFILL_FLIES (lean kernel, constant / group fields): the refill pass has its own copy of the flight
        // segment, in ONE basic block with the observables of the loaded state -- the two dependency chains
        // (|k|^2 -> E -> E_np -> sums, 30 serial fp64 operations; free-flight time -> drift -> boundary test) are
        // independent, but as long as the flight lived in the shared tail they ran one after the other, each at one
        // instruction per 8 cycles (profiles/r2x: the refill block 75, the tail 176 stall samples per instruction).
        constexpr bool REFILL_FLIES = EMC_REFILL_FLIES && LEAN && (EMC_REFILL_FLIES > 1 || !MESHF);
        if (phase == PH_REFILL) {
            // ---- refill phase: main_.py:361-373 for the next 32 particles -----------------------
            i = w0 + next * Wp + lane;
            act = next + lane < wlen;
            if (AHEAD && next + lane + 32 < wlen) {
                ahead = true;
                nx_ = A.x[i + 32u * Wp]; ny_ = A.y[i + 32u * Wp]; nz_ = A.z[i + 32u * Wp];
            }
#if EMC_STAGE_REFILL
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
            if (BULK && next + 32u <= wlen) {
                // a full batch: it sits in its tile (issued EMC_REFILL_TILES refills ago); act is true for every lane
                RefillTile &RT = RTW[EMC_REFILL_TILES == 1 ? 0 : rt_count % EMC_REFILL_TILES];
                refill_wait(RT, (rt_count / EMC_REFILL_TILES) & 1u);
                ++rt_count;
                p.kx = RT.row[0][lane]; p.ky = RT.row[1][lane]; p.kz = RT.row[2][lane];
                p.x = RT.row[3][lane]; p.y = RT.row[4][lane]; p.z = RT.row[5][lane];
                p.dte = RT.row[6][lane];
                p.val = RT.val[lane];
                // The batch EMC_REFILL_TILES further on, if it is a full one, too, goes into this tile -- once every
                // lane's loads have COMPLETED.  A shared-memory load that has been issued has not yet read the tile, and a
                // warp barrier orders instructions, not the copy engine: without the store below, one copy in ~10^9
                // particle-timesteps landed before a lane's load had executed and gave that lane a row of the NEXT batch
                // (run-to-run differences of a few trajectories, found with profiles/kernel_ab.py's checksum once a
                // larger scatter queue had moved the tiles).  The store's operand is a fold of all eight loaded
                // registers, so it cannot issue before they have arrived; the copy is issued behind it and the barrier.
                // (Issuing at the end of the refill block or of the pass instead is safe, too, but the shorter lead costs
                // 4 % and 7 %: 0.958 / 0.984 vs 0.92 ms, profiles/r2am_*.)
                {
                    const uint32_t fold = (uint32_t)__double2loint(p.kx) ^ (uint32_t)__double2loint(p.ky) ^ (uint32_t)__double2loint(p.kz) ^
                                          (uint32_t)__double2loint(p.x) ^ (uint32_t)__double2loint(p.y) ^ (uint32_t)__double2loint(p.z) ^
                                          (uint32_t)__double2loint(p.dte) ^ (uint32_t)p.val;
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(&RT.loaded)), "r"(fold) : "memory");
                }
                __syncwarp();
                if (next + 32u * (EMC_REFILL_TILES + 1) <= wlen && elect_one()) refill_issue(RT, A, w0 + (next + 32u * EMC_REFILL_TILES) * Wp, &tmap);
            } else if (act) {
                if constexpr (CHAIN) {
                    // the slice was written by another SM a moment ago: read it where that SM's stores went (L2), whatever
                    // this SM's L1 still holds of an earlier timestep of the same slice
                    p.kx = __ldcg(A.kx + i); p.ky = __ldcg(A.ky + i); p.kz = __ldcg(A.kz + i);
                    p.x = __ldcg(A.x + i); p.y = __ldcg(A.y + i); p.z = __ldcg(A.z + i);
                    p.dte = __ldcg(A.dtau + i);
                    p.val = __ldcg(A.valley + i);
                } else {
                p.kx = A.kx[i]; p.ky = A.ky[i]; p.kz = A.kz[i];
                p.x = A.x[i]; p.y = A.y[i]; p.z = A.z[i];
                p.dte = A.dtau[i];
                p.val = A.valley[i];
                }
            }
            if (act) {
                p.off = next + lane;
                p.grp = 0;
                if (REPLAY) s0 = (uint32_t)A.cursor[i];
                if (FIELD == EMC_FIELD_GROUPS) {
                    int32_t g = grp_next;
                    uint32_t r = grp_rem + lane;
                    while (r >= gsz) { r -= gsz; ++g; }
                    p.grp = g >= P.n_groups ? P.n_groups - 1 : g;
                    if (EARLY_GF) gfz = __ldg(P.field_groups + 3 * p.grp + 2);
                }
                if (MESHF && !(AHEAD && field_staged)) {     // else: staged by the previous refill
                    cell = mesh_cell_of<FIELD>(P, p);
                    if (AHEAD) field_stage<FIELD>(A, cell, FS, lane);
                    else field_prefetch<FIELD>(A, cell);
                }
                if (ML && p.val >= 0) {                           // :366 / :372
                    layer = where_am_i(P, p.z);
                    if (layer < 0) { p.val = -(1 + EMC_FAULT_OUTSIDE); layer = 0; }
                }
                // obs_before: the observables of the ensemble AS LOADED (= the result of the previous
                // timestep) are reduced here, where all 32 lanes are busy, instead of at retirement
                // (17 of 32 lanes on average); a time loop reads the same rows one launch later
                if (REFILL_FLIES) {
                    // (the sums are accumulated whether or not this launch reports them: a branch on A.obs would
                    // split the block; the lean kernel never reduces at retirement)
                    if (p.val >= 0) {
                        const double dte = p.dte;                 // :361
                        if (dte >= dt) { dt_seg = dt; p.dte = dte - dt; }      // :363-367, :390-391
                        else { dt_seg = dte; pend = true; }       // :370-373
                        obs_particle(P, P.mat[ML ? layer : 0], acc, p.kx, p.ky, p.kz, p.z, p.val);
                        if (AHEAD_OK && AHEAD) asm volatile("cp.async.wait_group 0;" ::: "memory");   // the staged field record
                        flight_segment(AHEAD_OK && AHEAD);
                    } else {
                        acc.count_fault();
                    }
                } else {
                if (A.obs && A.obs_before) obs_particle(P, P.mat[ML ? layer : 0], acc, p.kx, p.ky, p.kz, p.z, p.val);
                if (p.val >= 0) {                                 // else frozen after a fault
                    const double dte = p.dte;                     // :361
                    fly = true;
                    if (dte >= dt) { dt_seg = dt; p.dte = dte - dt; }      // :363-367, :390-391
                    else { dt_seg = dte; pend = true; }           // :370-373
                }
                }
            }
            next += 32;
#if EMC_STAGE_REFILL
            stage_batch(WQ, A, wstart + next + lane, wend, lane);
#endif
#if EMC_PREFETCH_BATCHES > 0
            {   // pull the state rows of a later refill batch into L2 while this one is processed:
                // a batch is 256 B of each of the seven fp64 rows (two 128-B lines) and 128 B of the
                // valley row, so ONE prefetch instruction covers it -- lane l < 14 takes line l&1 of
                // row l>>1, lane 14 the valley line (the row pointers are consecutive kernel
                // parameters, indexed from the constant bank)
                const uint32_t o0 = next + (uint32_t)(EMC_PREFETCH_BATCHES - 1) * 32u;
                const int64_t j0 = wstart + (int64_t)o0 * Wp;
                if (!BULK && o0 < wlen && lane < 15 && (!(lane & 1) || lane == 14 || o0 + 16 < wlen)) {
                    const char *base = reinterpret_cast<const char *>((&A.kx)[lane >> 1]);
                    const int64_t off = lane < 14 ? j0 * 8 + (lane & 1) * 128 : j0 * 4;
                    prefetch_l2(base + off);
                }
            }
#endif
            if (FIELD == EMC_FIELD_GROUPS) {
                if (ILV) {
                    // (saturating: past the last group every lane clamps to it anyway)
                    grp_next = (int32_t)min((uint32_t)grp_next + grp_dq, 0x7fffffffu);
                    grp_rem += grp_dr;                            // < 2 gsz <= 2^32
                    if (grp_rem >= gsz) { grp_rem -= gsz; if (grp_next < 0x7fffffff) ++grp_next; }
                } else {
                    grp_rem += 32;
                    while (grp_rem >= gsz) { grp_rem -= gsz; ++grp_next; }
                }
            }
        } else {
            // ---- an event, main_.py:375-389 / scatter_.py:986-1099, in one or two passes ---------
            const bool sel = (phase == PH_SELECT);
            WarpQueue &Qp = sel ? WQ.ev : WQ.sc;
            const int cnt = sel ? n_sel : n_sc;
            int &qlen = sel ? qn : qsc;
            act = lane < cnt;
            qlen -= cnt;
            const int slot = qlen + lane;
            if (act) {
                queue_pop<FIELD>(Qp, slot, p, s0);
                if (EARLY_GF) gfz = __ldg(P.field_groups + 3 * p.grp + 2);
                if constexpr (!THR) {
                    if (!sel) {
                        const double2 er = WQ.sc_er[slot];
                        mech = WQ.sc_mech[slot]; e_np = er.x; r_az = er.y;
                    }
                }
                if constexpr (PARKF) {
                    if (!sel) {
                        const double2 er = WQ.sc_er[slot];
                        pf.e_np = er.x; pf.r_az = er.y; pf.r2 = WQ.sc_r2[slot];
                    }
                }
            }
            __syncwarp();
            if (act) {
                i = ILV ? w0 + (p.off & ~31u) * Wp + (p.off & 31u) : w0 + p.off;
                if (MESHF) { cell = mesh_cell_of<FIELD>(P, p); field_prefetch<FIELD>(A, cell); }
                if constexpr (REPLAY) rng.init(A.uniforms, A.stride, i, (int32_t)s0);
                else rng.init(A.seed, A.pid_offset + (uint64_t)i, step_now, s0);
                int v = p.val, f;
                if (ML) layer = where_am_i(P, p.z);               // :377 (and :387: a scatter does not move)
                const Mat &MT = P.mat[ML ? (layer < 0 ? 0 : layer) : 0];
                const MechMeta &M = Ms[ML ? (layer < 0 ? 0 : layer) : 0];
                if (ML && layer < 0) {
                    f = EMC_FAULT_OUTSIDE;
                    layer = 0;
                } else if (sel && SPEC && THR && EMC_SELF_THRESHOLD == 2) {
                    // finished here (a certain, clean self-scattering event) or parked for the literal path
                    f = -1;
                    // (Tried: the next free flight and its flight segment inside this block as well, speculatively on
                    // the self-scattering k' and committed with selects -- 0.938 vs 0.918 ms, profiles/r2ac_ab_*: the
                    // segment depends on k', so it only lengthens the block's longest chain.)
                    if (select_self_thr<Stream, PARKF>(P, MT, p.kx, p.ky, p.kz, v, rng, lg, pf)) tail = true;
                    else { park = true; s0 = rng.state(); }
                } else if (sel && SPEC) {
                    bool is_self;
                    f = select_self_spec<Stream, LEAN, THR>(P, MT, M, p.kx, p.ky, p.kz, v, rng, e_np, r_az, mech, is_self, lg);
                    // THR: a cold exit (fault, r2 >= min_last) is parked like a real mechanism -- the scatter
                    // pass runs the whole event literally, whatever it turns out to be
                    if (THR && f == EMC_SELECT_COLD) { f = -1; is_self = false; }
                    if (f < 0) {
                        if (is_self) tail = true;
                        else { park = true; s0 = rng.state(); }
                    }
                } else if (sel) {
                    // select pass: energy, table row, r2 / azimuth uniforms, mechanism (:1092, :1029-1034)
                    f = event_select(P, MT, p.kx, p.ky, p.kz, v, rng, e_np, r_az, mech);
                    if (f < 0) {
                        if (M.sampler[v][mech] == EMC_SAMPLER_SELF) {
                            rng.draw_polar(false);                // no polar draw (:889); fetches the flight uniform
                            f = scatter_self<EXACT>(P, MT, M, p.kx, p.ky, p.kz, v, mech, e_np);
                            tail = true;
                        } else {
                            park = true;                          // real mechanism: finish in a scatter pass
                            s0 = rng.state();
                        }
                    }
                } else if (THR) {
                    // scatter pass of the threshold form: the parked event again from its counters, now
                    // with the whole table row (a real mechanism: polar sampler, rotation :1059-1089)
                    if constexpr (PARKF) {
                        // ... or, with the select pass's E_np, r2 and azimuth uniform, the rest of it (scatter_parked)
                        f = scatter_parked<Stream>(P, MT, M, p.kx, p.ky, p.kz, v, rng, pf);
                        lg = fast_log(rng.free_flight());
                    } else if constexpr (THR) {
                        const EventOut o = event_cold(&P, &MT, &M, p.kx, p.ky, p.kz, v, A.seed, A.pid_offset + (uint64_t)i, step_now, s0);
                        f = o.f; lg = o.lg;
                        if (f < 0) { p.kx = o.kx; p.ky = o.ky; p.kz = o.kz; v = o.val; }
                    }
                    tail = true;
                } else {
                    // scatter pass: polar sampler, rotation (:1059-1089)
                    const double r_pol = rng.draw_polar(true);
                    if (SPEC) lg = fast_log(rng.free_flight());   // independent of the rotation below
                    f = scatter_core<EXACT>(P, MT, M, p.kx, p.ky, p.kz, v, mech, e_np, r_az, r_pol);
                    tail = true;
                }
                if (rng.overflow()) { f = EMC_FAULT_STREAM; park = false; }
                if (f >= 0) { p.val = -(1 + f); tail = false; }  // frozen; dtau keeps dte
                else p.val = v;
            }
            if (tail) {
                // the event happened: draw the next free flight and fly (main_.py:379-389)
                acc.count_event();
                const double dte2 = p.dte;                        // :376
                const double dt3 = P.neg_inv_gamma * (SPEC ? lg : fast_log(rng.free_flight()));    // :379
                if (rng.overflow()) {
                    p.val = -(1 + EMC_FAULT_STREAM);
                } else {
                    const double dtp = dt - dte2;                 // :381
                    dt_seg = (dt3 <= dtp) ? dt3 : dtp;            // :382-385
                    fly = true;
                    const double dte = dte2 + dt3;                // :389
                    if (dte < dt) { p.dte = dte; pend = true; }   // :375
                    else p.dte = dte - dt;                        // :390-391
                }
                rng.end_event();
                s0 = rng.state();
            } else if (act && !park && REPLAY) {
                s0 = rng.state();                                 // faulted: cursor as far as it was read
            }
        }
        if (AHEAD_OK && AHEAD && phase == PH_REFILL) asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (fly) flight_segment(AHEAD_OK && AHEAD && phase == PH_REFILL);   // :367 / :373 / :388
        if (phase == PH_REFILL && AHEAD) {
            // the rs slots of this batch have been read (own lane, program order): gather the next one
            if (ahead) {
                Particle q;
                q.x = nx_; q.y = ny_; q.z = nz_;
                field_stage<FIELD>(A, mesh_cell_of<FIELD>(P, q), FS, lane);
            }
            field_staged = next < wlen;           // `next` already points at that batch
        }
#if EMC_PREDICATED_TAIL
        // queue pushes FIRST, write-back last (EMC_PUSH_FIRST): the ballot that followed the eight scattered
        // global stores waited ~50 cycles for them to leave the load/store queue (its destination register
        // was still an address operand of one of them, profiles/r2x), and the loop's back edge another ~75 for
        // the shared-memory stores; in this order the global stores drain behind the branch
        auto push_queues = [&]() {
            const unsigned pm = __ballot_sync(FULL, pend);
            queue_push_if<FIELD>(WQ.ev, qn + __popc(pm & lt_mask), p, s0, pend);
            qn += __popc(pm);
            if (EMC_PARK_UNCOND || phase == PH_SELECT) {
                const unsigned km = __ballot_sync(FULL, park);
                const int slot = qsc + __popc(km & lt_mask);
                queue_push_if<FIELD>(WQ.sc, slot, p, s0, park);
                if constexpr (!THR)
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.shared.u32 [%0], %1;\n\t"
                                 "@q st.shared.v2.f64 [%2], {%3, %4};\n\t}"
                                 ::"r"(smem_u32(&WQ.sc_mech[0]) + (uint32_t)slot * 4u), "r"(mech),
                                 "r"(smem_u32(&WQ.sc_er[0]) + (uint32_t)slot * 16u), "d"(e_np), "d"(r_az), "r"((uint32_t)park) : "memory");
                if constexpr (PARKF)
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.shared.f64 [%0], %1;\n\t"
                                 "@q st.shared.v2.f64 [%2], {%3, %4};\n\t}"
                                 ::"r"(smem_u32(&WQ.sc_r2[0]) + (uint32_t)slot * 8u), "d"(pf.r2),
                                 "r"(smem_u32(&WQ.sc_er[0]) + (uint32_t)slot * 16u), "d"(pf.e_np), "d"(pf.r_az), "r"((uint32_t)park) : "memory");
                qsc += __popc(km);
            }
        };
        if (EMC_PUSH_FIRST) push_queues();
        {
            const bool done = act && !pend && !park;
            int val = p.val, lay = 0;
            if (ML && val >= 0) {                     // main_.py:393: the layer of the final position
                lay = where_am_i(P, p.z);
                if (lay < 0) { val = -(1 + EMC_FAULT_OUTSIDE); lay = 0; }
            }
            retire_stores_if(A, i, p, val, done);
            if (REPLAY && done) A.cursor[i] = (int32_t)s0;
            if (!LEAN && A.obs && !A.obs_before) {    // warp-uniform: the default phase reduces at load time
                if (done) obs_particle(P, P.mat[lay], acc, p.kx, p.ky, p.kz, p.z, val);
            }
            if (DEPOSIT && done) {
                if (A.smem_cells > 0)
                    deposit_ngp_particle(P, p.x, p.y, p.z, [&](int64_t c) { atomicAdd(&smesh[c], 1u); });
                else
                    deposit_ngp_particle(P, p.x, p.y, p.z, [&](int64_t c) {
                        atomicAdd(reinterpret_cast<unsigned long long *>(A.mesh) + c, 1ull); });
            }
        }
        if (!EMC_PUSH_FIRST) push_queues();
#else
        if (act && !pend && !park) retire<REPLAY, DEPOSIT, ML>(P, A, i, p, s0, acc, smesh);
        const unsigned pm = __ballot_sync(FULL, pend);
        if (pend) queue_push<FIELD>(WQ.ev, qn + __popc(pm & lt_mask), p, s0);
        qn += __popc(pm);
        if (phase == PH_SELECT) {
            const unsigned km = __ballot_sync(FULL, park);
            if (park) {
                const int slot = qsc + __popc(km & lt_mask);
                queue_push<FIELD>(WQ.sc, slot, p, s0);
                if constexpr (!THR) { WQ.sc_mech[slot] = mech; WQ.sc_er[slot] = make_double2(e_np, r_az); }
                if constexpr (PARKF) { WQ.sc_r2[slot] = pf.r2; WQ.sc_er[slot] = make_double2(pf.e_np, pf.r_az); }
            }
            qsc += __popc(km);
        }
#endif
        __syncwarp();
#if EMC_ROTATE_LOOP
        phase = choose_phase();
#endif
        }

    if (DEPOSIT && A.smem_cells > 0) {
        __syncthreads();
        for (int t = threadIdx.x; t < A.smem_cells; t += blockDim.x) {
            const unsigned int v = smesh[t];
            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(A.mesh) + t, (unsigned long long)v);
        }
    }
#ifdef EMC_WARP_END_TIMES
    if (LEAN && lane == 0 && gw < 8192) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_warp_t1[gw] = t;
    }
#endif
    if (LEAN) acc.derive_segments();
    if (obs_out) obs_block_reduce_and_finish(acc, obs_parts, obs_ticket, obs_out, bid, nblk, !CHAIN);
    if constexpr (!CHAIN) {
        break;
    } else {
        // The slice is written back: hand its next timestep to whichever block asks next.  Every thread's state
        // stores precede the barriers of the reduction above, thread 0's fence makes them visible at GPU scope
        // before the item (release), and the popping thread acquires it.
        if (threadIdx.x == 0 && chain_j + 1u < (uint32_t)A.chain_steps) {
            ChainCtl *C = A.chain;
            __threadfence();
            const unsigned int Pn = (unsigned int)A.chain_slices + atomicAdd(&C->pushed, 1u);
            const unsigned long long v = ((unsigned long long)(Pn + 1u) << 32) | ((unsigned long long)(chain_j + 1u) << 16) | bid;
            unsigned long long *slot = &C->ready[Pn % CHAIN_RING];
            // wait for an EMPTY slot (its reader of CHAIN_RING items ago empties it): never taken in practice, but it makes
            // the ring safe for every schedule instead of for every plausible one
            for (;;) {
                unsigned long long cur;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(slot) : "memory");
                if (cur == 0ull) break;
                __nanosleep(64);
            }
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(slot), "l"(v) : "memory");
        }
    }
    }   // work items
}

template <class Stream, int FIELD, bool EXACT, bool DEPOSIT, bool ML, bool LEAN = false, bool CHAIN = false>
static cudaError_t launch_flight_k(const Ctx *ctx, const FlightArgs &A, int grid, size_t smem, cudaStream_t st)
{
    auto k = flight_scatter_kernel<Stream, FIELD, EXACT, DEPOSIT, ML, LEAN, CHAIN>;
    // launch_flight sized `smem` for the general queues; the lean kernels' are smaller or larger
    smem = smem - sizeof(WarpQueues) * (FLIGHT_THREADS / 32) + sizeof(typename FlightQueues<LEAN, ML>::type) * (FLIGHT_THREADS / 32);
    // persistent grid: exactly the blocks that are resident at once (one wave), unless the
    // ensemble is too small to give every block a full tile.  The attribute / occupancy queries
    // are cached per kernel variant and device (they cost tens of microseconds per launch).
    static size_t cached_smem[16];
    static int cached_per_sm[16];
    const int dev = ctx->device & 15;
    cudaError_t e;
    if (cached_per_sm[dev] == 0 || cached_smem[dev] != smem) {
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        int q = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k, FLIGHT_THREADS, smem)) != cudaSuccess) return e;
        cached_per_sm[dev] = q < 1 ? 1 : q;
        cached_smem[dev] = smem;
    }
    const int per_sm = cached_per_sm[dev];
    const int resident = ctx->sm_count * per_sm * ctx->waves;
    if (grid > resident) grid = resident;
    if constexpr (CHAIN) {
        // the slices of a chain are the blocks of the one-timestep launch; all of them resident (a block waits for items
        // that running blocks produce)
        if (ctx->waves != 1 || grid > ctx->sm_count * per_sm) return cudaErrorNotSupported;
        FlightArgs B = A;
        B.chain_slices = grid;
        k<<<grid, FLIGHT_THREADS, smem, st>>>(ctx->dc, ctx->mech, B, ctx->state_tmap);
        return cudaGetLastError();
    }
    k<<<grid, FLIGHT_THREADS, smem, st>>>(ctx->dc, ctx->mech, A, ctx->state_tmap);
    return cudaGetLastError();
}

template <class Stream, int FIELD>
static cudaError_t launch_flight_t(const Ctx *ctx, const FlightArgs &A, int grid, size_t smem, cudaStream_t st)
{
    if constexpr (!Stream::is_replay) {
        // the lean instantiation (see flight_scatter_kernel) whenever its compile-time assumptions hold
        const bool zon_ok = (FIELD != EMC_FIELD_CONSTANT && FIELD != EMC_FIELD_GROUPS) || ctx->dc.zonly;
        if constexpr (FIELD == EMC_FIELD_CONSTANT || FIELD == EMC_FIELD_GROUPS) {
            if (A.chain)      // launch_flight_chain has checked the lean conditions
                return launch_flight_k<Stream, FIELD, false, false, false, true, true>(ctx, A, grid, smem, st);
        }
        if (EMC_LEAN_KERNEL && !ctx->no_lean && !ctx->exact_libm && !A.mesh && !(A.obs && !A.obs_before) &&
            !ctx->dc.literal_scan && ctx->self_last && zon_ok)
            return ctx->dc.n_layers > 1 ? launch_flight_k<Stream, FIELD, false, false, true, true>(ctx, A, grid, smem, st)
                                        : launch_flight_k<Stream, FIELD, false, false, false, true>(ctx, A, grid, smem, st);
    }
    // multi-layer devices: separate deposit kernel (checked by the caller)
    if (ctx->dc.n_layers > 1)
        return ctx->exact_libm ? launch_flight_k<Stream, FIELD, true, false, true>(ctx, A, grid, smem, st)
                               : launch_flight_k<Stream, FIELD, false, false, true>(ctx, A, grid, smem, st);
    if (ctx->exact_libm)
        return A.mesh ? launch_flight_k<Stream, FIELD, true, true, false>(ctx, A, grid, smem, st)
                      : launch_flight_k<Stream, FIELD, true, false, false>(ctx, A, grid, smem, st);
    return A.mesh ? launch_flight_k<Stream, FIELD, false, true, false>(ctx, A, grid, smem, st)
                  : launch_flight_k<Stream, FIELD, false, false, false>(ctx, A, grid, smem, st);
}

// The CUtensorMap of the state rows as one (7, n) fp64 array with a box of 32 particles x 7 rows, when the
// caller's seven pointers ARE such an array (equal pitch, a multiple of 16 bytes; 16-byte aligned base).  Encoded
// by the driver (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: no link-time dependency on libcuda) and
// cached while base, length and pitch stay the same.  false = use the per-row bulk copies.
static bool state_tensor_map(Ctx *ctx, const FlightArgs &A)
{
    const char *base = reinterpret_cast<const char *>(A.kx);
    const ptrdiff_t pitch = reinterpret_cast<const char *>(A.ky) - base;
    if (pitch <= 0 || (pitch & 15) || ((uintptr_t)base & 15) || A.n < 32 || A.n > 0x7fffffff || pitch < A.n * 8) return false;
    for (int r = 1; r < 7; ++r)
        if (reinterpret_cast<const char *>((&A.kx)[r]) - base != pitch * r) return false;
    if (ctx->tmap_base == base && ctx->tmap_n == A.n && ctx->tmap_pitch == pitch) return true;
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn encode = []() -> EncodeFn {         // looked up once (thread-safe static initialisation)
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeFn>(fn);
        (void)cudaGetLastError();
        return nullptr;
    }();
    if (!encode) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)A.n, 7};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {32, 7}, estr[2] = {1, 1};
    ctx->tmap_base = nullptr;
    if (encode(&ctx->state_tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<char *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    ctx->tmap_base = base; ctx->tmap_n = A.n; ctx->tmap_pitch = pitch;
    return true;
}

cudaError_t launch_flight(Ctx *ctx, FlightArgs &A, bool replay, int field_mode, cudaStream_t st)
{
    if (A.n >= ((int64_t)1 << 32) - 64) return cudaErrorInvalidValue;   // 32-bit particle index inside a launch (checked by the ABI)
    // one warp needs at least a few refills of work; small ensembles use fewer blocks
    const int64_t blocks_needed = (A.n + FLIGHT_THREADS - 1) / FLIGHT_THREADS;
    int grid = (int)(blocks_needed < (int64_t)ctx->max_grid ? blocks_needed : ctx->max_grid);
    if (grid < 1) grid = 1;
    size_t smem = sizeof(WarpQueues) * (FLIGHT_THREADS / 32);
    // refill-ahead field staging only where the field does not live in L1 anyway (EMC_FIELD_AHEAD)
    const int64_t field_cells = (int64_t)ctx->dc.seg[0] * ctx->dc.seg[1] * ctx->dc.seg[2];
    A.field_ahead = EMC_FIELD_AHEAD && (field_mode == EMC_FIELD_MESH_PACKED || field_mode == EMC_FIELD_MESH_XZ) &&
                    field_cells >= 8192 && !ctx->no_field_ahead;
    if (A.field_ahead) smem += sizeof(FieldStage) * (FLIGHT_THREADS / 32);
    {   // bulk-async refill: every state row 16-byte aligned (the copies are 256-B slices at multiples of 32 particles)
        uintptr_t bits = (uintptr_t)A.valley;
        for (int r = 0; r < 7; ++r) bits |= (uintptr_t)(&A.kx)[r];
        // (not for the mesh-field modes: the tiles' 31 KB of shared memory come out of the L1 that the
        // field gathers live on -- measured 0.745 ms with, 0.721 ms without on 1024 x 1 x 1024, profiles/ab_r2e)
        const bool mesh_field = field_mode == EMC_FIELD_MESH || field_mode == EMC_FIELD_MESH_PACKED || field_mode == EMC_FIELD_MESH_XZ;
        A.bulk = EMC_BULK_REFILL && (bits & 15) == 0 && !ctx->no_bulk_refill && !replay && (!mesh_field || ctx->bulk_mesh);
        if (A.bulk && EMC_REFILL_TENSOR_MAP && !ctx->no_tensor_map && state_tensor_map(ctx, A)) A.bulk = 2;
        if (A.bulk) smem += sizeof(RefillTile) * EMC_REFILL_TILES * (FLIGHT_THREADS / 32);
    }
    if (EMC_OBS_SMEM) {
        smem += sizeof(double) * 5 * FLIGHT_THREADS;
        // ObsAccShared packs the valley / fault counts into 16 bits per lane and launch
        const int64_t warps = (int64_t)(grid < ctx->sm_count ? grid : ctx->sm_count) * (FLIGHT_THREADS / 32);   // one wave at least
        if ((A.n + warps - 1) / warps / 32 + 1 > 65535) return cudaErrorInvalidValue;
    }
    A.smem_cells = 0;
    if (A.mesh) {
        const int64_t cells = (int64_t)ctx->dc.seg[0] * ctx->dc.seg[1] * ctx->dc.seg[2];
        // privatise the counts in shared memory only while two blocks per SM still fit beside the queues
        if (cells * 4 <= 16 * 1024) { A.smem_cells = (int)cells; smem += (size_t)cells * 4; }
    }
    if (!A.chain) {
        A.partials = ctx->partials;
        A.ticket = ctx->ticket;
    }
    A.mech_layers = ctx->d_mech_layers;
    ctx->launches++;
#define EMC_DISPATCH(STREAM)                                                                     \
    switch (field_mode) {                                                                        \
    case EMC_FIELD_CONSTANT: return launch_flight_t<STREAM, EMC_FIELD_CONSTANT>(ctx, A, grid, smem, st); \
    case EMC_FIELD_GROUPS:   return launch_flight_t<STREAM, EMC_FIELD_GROUPS>(ctx, A, grid, smem, st);   \
    case EMC_FIELD_MESH:     return launch_flight_t<STREAM, EMC_FIELD_MESH>(ctx, A, grid, smem, st);     \
    case EMC_FIELD_MESH_PACKED: return launch_flight_t<STREAM, EMC_FIELD_MESH_PACKED>(ctx, A, grid, smem, st); \
    case EMC_FIELD_MESH_XZ:  return launch_flight_t<STREAM, EMC_FIELD_MESH_XZ>(ctx, A, grid, smem, st); \
    default: return cudaErrorInvalidValue;                                                       \
    }
#ifdef EMC_PROBE_ONLY
    // static-analysis builds (profiles/sass_static.py): only the configs[1] instantiation, compiled in seconds
    if (replay || field_mode != EMC_FIELD_GROUPS || ctx->dc.n_layers > 1 || ctx->exact_libm || A.mesh) return cudaErrorInvalidValue;
    return launch_flight_k<PhiloxStream, EMC_FIELD_GROUPS, false, false, false, EMC_LEAN_KERNEL != 0>(ctx, A, grid, smem, st);
#else
    if (replay) { EMC_DISPATCH(ReplayStream) } else { EMC_DISPATCH(PhiloxStream) }
#endif
#undef EMC_DISPATCH
}

// ------------------------------------------------------------------------------------------
// Timestep chain: n_steps timesteps WITH their observable rows in one persistent launch of the lean one-timestep
// kernel (CHAIN instantiation, ChainCtl in emc_host.h).  What it removes is the time an SM idles per launch: the
// launch ramp, the spread of the blocks' finishing times (ncu r2z: the average SM is busy 97.3 % of a 0.92-ms
// launch, i.e. 25 us idle; ~30 us of a 0.18-ms launch over a 2^22-particle chunk of the host-array loop), the
// gap between launches.  The state makes the same HBM round trip per timestep as with n_steps launches.
// cudaErrorNotSupported: the conditions do not hold (nothing has been enqueued); the caller loops over launches.
// ------------------------------------------------------------------------------------------
cudaError_t launch_flight_chain(Ctx *ctx, FlightArgs &A, int field_mode, int n_steps, cudaStream_t st)
{
    const bool zonly = (field_mode == EMC_FIELD_CONSTANT && ctx->zonly_constant) || (field_mode == EMC_FIELD_GROUPS && ctx->zonly_groups);
    if (!EMC_LEAN_KERNEL || ctx->no_chain || ctx->no_lean || ctx->exact_libm || ctx->dc.literal_scan || !ctx->self_last ||
        !zonly || ctx->dc.n_layers != 1 || ctx->waves != 1 || !ctx->obs_before || !A.obs || A.mesh ||
        (field_mode != EMC_FIELD_CONSTANT && field_mode != EMC_FIELD_GROUPS) ||
        (field_mode == EMC_FIELD_GROUPS && !ctx->dc.field_groups) || n_steps < 2 || n_steps > CHAIN_MAX_STEPS ||
        A.n < 1 || A.n >= ((int64_t)1 << 32) - 64 || ctx->max_grid > 65535 || ctx->max_grid >= CHAIN_RING)
        return cudaErrorNotSupported;
    cudaError_t e;
    if (!ctx->d_chain && (e = cudaMalloc(&ctx->d_chain, sizeof(ChainCtl))) != cudaSuccess) return e;
    const size_t need = (size_t)n_steps * (size_t)ctx->max_grid;
    if (ctx->chain_partials_cap < need) {
        // (cudaFree synchronises the device: an earlier chain may still be using the old buffer; growth is rare)
        if (ctx->chain_partials && (e = cudaFree(ctx->chain_partials)) != cudaSuccess) return e;
        ctx->chain_partials = nullptr;
        ctx->chain_partials_cap = 0;
        if ((e = cudaMalloc(&ctx->chain_partials, sizeof(ObsPartial) * need)) != cudaSuccess) return e;
        ctx->chain_partials_cap = need;
    }
    if ((e = cudaMemsetAsync(ctx->d_chain, 0, sizeof(ChainCtl), st)) != cudaSuccess) return e;
    A.chain = ctx->d_chain;
    A.chain_steps = n_steps;
    A.chain_slices = 0;                   // launch_flight_k: the grid it launches
    A.partials = ctx->chain_partials;
    A.ticket = nullptr;
    A.obs_before = 1;
    ctx->dc.zonly = zonly;
    return launch_flight(ctx, A, false, field_mode, st);
}

// ------------------------------------------------------------------------------------------
// K timesteps in ONE persistent launch: for uncoupled runs that do not need
// the observables of every timestep -- burn-in phases, the velocity-field sweep that samples every
// tenth timestep -- a particle stays in its warp's queues from timestep to timestep instead of being
// written back and reloaded: no refill and no write-back per timestep (25 % of the one-step kernel's
// stall samples).  A third queue, NX, holds the particles that have finished a timestep; an ADVANCE
// pass pops 32 of them (all lanes busy), starts their next timestep (main_.py:361-373 on the carried
// dtau) or, after the last one, writes them back.  Same per-particle arithmetic and the same Philox
// counters (key = seed, counter = (block, first_step + ts, particle)) as n_steps launches of
// flight_scatter_kernel: bit-identical final state (tests/test_gpu_parity.py).  Restricted to the
// lean case: free-running Philox, default libm composition, one layer, constant or group field
// with Ex = Ey = +0, padded tables with self-scattering last; everything else takes the launch loop.
// The parked (real-mechanism) particles carry no extra fields: the scatter pass re-derives energy,
// mechanism and azimuth from the same counters (scatter_event), 0.1 passes per batch.
// ------------------------------------------------------------------------------------------
struct StepQueues {
    WarpQueue ev;      // waiting for the next event of their current timestep
    WarpQueue sc;      // drew a real mechanism
    WarpQueue nx;      // finished a timestep: next timestep / write-back after the last one
};

template <int FIELD>
__global__ void __launch_bounds__(FLIGHT_THREADS, 1)
flight_steps_kernel(const __grid_constant__ DevConst P, const __grid_constant__ MechMeta Mg,
                    const __grid_constant__ FlightArgs A, const int n_steps, int32_t *stuck_flag)
{
    static_assert(FIELD == EMC_FIELD_CONSTANT || FIELD == EMC_FIELD_GROUPS, "uncoupled field modes only");
    __shared__ MechMeta Ms;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    StepQueues *queues = reinterpret_cast<StepQueues *>(dyn_smem);
    {
        const int nw = sizeof(MechMeta) / 4;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&Mg);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&Ms);
        for (int t = threadIdx.x; t < nw; t += blockDim.x) dst[t] = src[t];
    }
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
    int lane;
    unsigned lt_mask;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    StepQueues &WQ = queues[warp];
    const int64_t total_warps = (int64_t)gridDim.x * (FLIGHT_THREADS / 32);
    const int64_t gw = (int64_t)blockIdx.x * (FLIGHT_THREADS / 32) + warp;
    // interleaved batches, as in flight_scatter_kernel (ILV): batch gw, gw + W, gw + 2 W ... of the ensemble, so that
    // every warp samples every field group.  Measured SLOWER here (0.848 vs 0.832 ms per timestep at K = 10,
    // profiles/r2av_ab_*: over K timesteps the groups' differences average out less than the strided plain loads
    // and write-backs cost); off.
    constexpr bool ILV = EMC_INTERLEAVE_STEPS_KERNEL != 0;
    int64_t wstart;
    uint32_t wlen_;
    if (ILV) {
        const int64_t nb = (A.n + 31) >> 5;
        const int64_t mine = gw < nb ? (nb - gw + total_warps - 1) / total_warps : 0;
        wstart = mine ? gw * 32 : 0;
        wlen_ = (uint32_t)(mine * 32);
        if (mine && gw + (mine - 1) * total_warps == nb - 1) wlen_ -= (uint32_t)(nb * 32 - A.n);
    } else {
        const int64_t slice = (((A.n + total_warps - 1) / total_warps) + 31) & ~(int64_t)31;
        wstart = gw * slice;
        if (wstart > A.n) wstart = A.n;
        const int64_t wend = (wstart + slice < A.n) ? wstart + slice : A.n;
        wlen_ = (uint32_t)(wend - wstart);
    }
    const uint32_t Wp = ILV ? (uint32_t)total_warps : 1u;
    uint32_t next = 0;
    const uint32_t wlen = wlen_, w0 = (uint32_t)wstart;
    int qn = 0, qsc = 0, qnx = 0;
    const double dt = P.dt;
    int32_t grp_next = 0;
    uint32_t grp_rem = 0;
    [[maybe_unused]] uint32_t grp_dq = 0, grp_dr = 32;
    const uint32_t gsz = (uint32_t)P.group_size;
    if (FIELD == EMC_FIELD_GROUPS) {
        const uint64_t g0 = A.pid_offset + (uint64_t)wstart;
        const uint64_t gq = g0 / (uint64_t)P.group_size;
        grp_next = (int32_t)(gq < (uint64_t)P.n_groups ? gq : (uint64_t)P.n_groups);
        grp_rem = (uint32_t)(g0 % (uint64_t)P.group_size);
        if (ILV) {
            const uint64_t delta = 32ull * (uint64_t)total_warps;
            const uint64_t dq = delta / (uint64_t)P.group_size;
            grp_dq = (uint32_t)(dq < 0x7fffffffull ? dq : 0x7fffffffull);
            grp_dr = (uint32_t)(delta % (uint64_t)P.group_size);
        }
    }
    const Mat &MT = P.mat[0];
    enum { PH_REFILL = 0, PH_SELECT = 1, PH_SCATTER = 2, PH_ADVANCE = 3 };
    while (true) {
        // a pass pops only what every queue it may push into can hold; a particle goes to exactly one
        const int room_ev = QCAP - qn, room_sc = QCAP - qsc, room_nx = QCAP - qnx;
        const int n_sc = min(min(32, qsc), min(room_ev, room_nx));
        const int n_sel = min(min(32, qn), min(room_sc, room_nx));
        const int n_adv = min(min(32, qnx), room_ev);            // NX gets back at most what was popped
        const bool can_refill = next < wlen && room_ev >= 32 && room_nx >= 32;
        int phase;
        if (n_sc == 32) phase = PH_SCATTER;
        else if (n_sel == 32) phase = PH_SELECT;
        else if (n_adv == 32) phase = PH_ADVANCE;
        else if (can_refill) phase = PH_REFILL;
        else if (n_sc > 0 && n_sc >= n_sel && n_sc >= n_adv) phase = PH_SCATTER;
        else if (n_sel > 0 && n_sel >= n_adv) phase = PH_SELECT;
        else if (n_adv > 0) phase = PH_ADVANCE;
        else {
            if ((qn | qsc | qnx) != 0 && lane == 0) atomicExch(stuck_flag, 1);     // cannot happen (see the analysis in DESIGN.md)
            break;
        }
        Particle p;
        uint32_t s0 = 0;                      // events so far in the timestep (low 16 bits) | timestep index << 16
        uint32_t i = 0;
        bool act, pend = false, fly = false, park = false, store = false;
        double dt_seg = 0.0;
        if (phase == PH_REFILL || phase == PH_ADVANCE) {
            if (phase == PH_REFILL) {
                i = w0 + next * Wp + lane;
                act = next + lane < wlen;
                if (act) {
                    p.kx = A.kx[i]; p.ky = A.ky[i]; p.kz = A.kz[i];
                    p.x = A.x[i]; p.y = A.y[i]; p.z = A.z[i];
                    p.dte = A.dtau[i];
                    p.val = A.valley[i];
                    p.off = next + lane;
                    p.grp = 0;
                    if (FIELD == EMC_FIELD_GROUPS) {
                        int32_t g = grp_next;
                        uint32_t r = grp_rem + lane;
                        while (r >= gsz) { r -= gsz; ++g; }
                        p.grp = g >= P.n_groups ? P.n_groups - 1 : g;
                    }
                    if (p.val < 0) act = false;               // frozen on entry: stays as it is in memory
                }
                next += 32;
#if EMC_PREFETCH_BATCHES > 0
                {
                    const uint32_t o0 = next + (uint32_t)(EMC_PREFETCH_BATCHES - 1) * 32u;
                    const int64_t j0 = wstart + (int64_t)o0 * Wp;
                    if (o0 < wlen && lane < 15 && (!(lane & 1) || lane == 14 || o0 + 16 < wlen)) {
                        const char *base = reinterpret_cast<const char *>((&A.kx)[lane >> 1]);
                        const int64_t off = lane < 14 ? j0 * 8 + (lane & 1) * 128 : j0 * 4;
                        prefetch_l2(base + off);
                    }
                }
#endif
                if (FIELD == EMC_FIELD_GROUPS) {
                    if (ILV) {
                        grp_next = (int32_t)min((uint32_t)grp_next + grp_dq, 0x7fffffffu);
                        grp_rem += grp_dr;
                        if (grp_rem >= gsz) { grp_rem -= gsz; if (grp_next < 0x7fffffff) ++grp_next; }
                    } else {
                        grp_rem += 32;
                        while (grp_rem >= gsz) { grp_rem -= gsz; ++grp_next; }
                    }
                }
            } else {
                act = lane < n_adv;
                qnx -= n_adv;
                if (act) queue_pop<FIELD>(WQ.nx, qnx + lane, p, s0);
                __syncwarp();
                if (act) {
                    i = ILV ? w0 + (p.off & ~31u) * Wp + (p.off & 31u) : w0 + p.off;
                    const uint32_t ts = s0 >> 16;
                    if (ts + 1u >= (uint32_t)n_steps) { store = true; act = false; }   // the last timestep is done: write back
                    else s0 = (ts + 1u) << 16;
                }
            }
            if (act) {                                        // main_.py:361-373 on the carried dtau
                const double dte = p.dte;
                fly = true;
                if (dte >= dt) { dt_seg = dt; p.dte = dte - dt; }
                else { dt_seg = dte; pend = true; }
            }
        } else {
            const bool sel = (phase == PH_SELECT);
            WarpQueue &Qp = sel ? WQ.ev : WQ.sc;
            const int cnt = sel ? n_sel : n_sc;
            int &qlen = sel ? qn : qsc;
            act = lane < cnt;
            qlen -= cnt;
            if (act) queue_pop<FIELD>(Qp, qlen + lane, p, s0);
            __syncwarp();
            bool tail = false;
            double lg = 0.0;
            PhiloxStreamSteps rng;
            if (act) {
                i = ILV ? w0 + (p.off & ~31u) * Wp + (p.off & 31u) : w0 + p.off;
                rng.init(A.seed, A.pid_offset + (uint64_t)i, A.step + (s0 >> 16), s0 & 0xffffu);
                int v = p.val, f, mech = 0;
                double e_np = 0.0, r_az = 0.0;
                if (sel && EMC_SELF_THRESHOLD == 2) {
                    f = -1;
                    ParkFields unused;
                    if (select_self_thr<PhiloxStreamSteps>(P, MT, p.kx, p.ky, p.kz, v, rng, lg, unused)) tail = true;
                    else park = true;                         // the event index in s0 is unchanged: the scatter pass redoes the event
                } else if (sel) {
                    bool is_self;
                    f = select_self_spec<PhiloxStreamSteps, true, EMC_SELF_THRESHOLD != 0>(P, MT, Ms, p.kx, p.ky, p.kz, v, rng, e_np, r_az, mech, is_self, lg);
                    // a cold exit (fault, r2 >= min_last) is parked like a real mechanism: the scatter pass
                    // runs the whole event literally, whatever it turns out to be
                    if (EMC_SELF_THRESHOLD && f == EMC_SELECT_COLD) { f = -1; is_self = false; }
                    if (f < 0) {
                        if (is_self) tail = true;
                        else park = true;                     // the event index in s0 is unchanged: the scatter pass redoes the event
                    }
                } else {
                    const EventOut o = event_cold(&P, &MT, &Ms, p.kx, p.ky, p.kz, v, A.seed, A.pid_offset + (uint64_t)i,
                                                  A.step + (s0 >> 16), s0 & 0xffffu);
                    f = o.f; lg = o.lg;
                    if (f < 0) { p.kx = o.kx; p.ky = o.ky; p.kz = o.kz; v = o.val; }
                    tail = f < 0;
                }
                // a faulted particle is written back at once, frozen (dtau keeps dte): only live valleys
                // (0..2) fit the queues' packed word
                if (f >= 0) { p.val = -(1 + f); tail = false; park = false; store = true; act = false; }
                else p.val = v;
            }
            if (tail) {
                const double dte2 = p.dte;                        // main_.py:376
                const double dt3 = P.neg_inv_gamma * lg;          // :379
                const double dtp = dt - dte2;                     // :381
                dt_seg = (dt3 <= dtp) ? dt3 : dtp;                // :382-385
                fly = true;
                const double dte = dte2 + dt3;                    // :389
                if (dte < dt) { p.dte = dte; pend = true; }       // :375
                else p.dte = dte - dt;                            // :390-391
                rng.end_event();
                if (rng.state() > 0xffffu) {                      // the queues keep 16 bits of the event counter
                    p.val = -(1 + EMC_FAULT_STREAM); fly = false; pend = false; store = true; act = false;
                } else {
                    s0 = (s0 & 0xffff0000u) | rng.state();
                }
            }
        }
        if (fly) {
            double efz;
            if (FIELD == EMC_FIELD_CONSTANT) efz = P.efield[2];
            else efz = __ldg(P.field_groups + 3 * p.grp + 2);
            const double dtm = div_const(dt_seg, MT.mass[p.val], MT.r_mass[p.val]);
            carrier_drift<1>(P, p.kx, p.ky, p.kz, p.x, p.y, p.z, 0.0, 0.0, efz, dt_seg, dtm);
        }
        retire_stores_if(A, i, p, p.val, store);
        const bool done = act && !pend && !park;                  // finished the timestep: next-step queue
        const unsigned pm = __ballot_sync(FULL, pend);
        queue_push_if<FIELD>(WQ.ev, qn + __popc(pm & lt_mask), p, s0, pend);
        qn += __popc(pm);
        const unsigned dm = __ballot_sync(FULL, done);
        queue_push_if<FIELD>(WQ.nx, qnx + __popc(dm & lt_mask), p, s0, done);
        qnx += __popc(dm);
        if (phase == PH_SELECT) {
            const unsigned km = __ballot_sync(FULL, park);
            queue_push_if<FIELD>(WQ.sc, qsc + __popc(km & lt_mask), p, s0, park);
            qsc += __popc(km);
        }
        __syncwarp();
    }
}

cudaError_t launch_flight_steps(Ctx *ctx, FlightArgs &A, int field_mode, int n_steps, cudaStream_t st)
{
    const int64_t blocks_needed = (A.n + FLIGHT_THREADS - 1) / FLIGHT_THREADS;
    int grid = (int)(blocks_needed < (int64_t)ctx->sm_count ? blocks_needed : ctx->sm_count);
    if (grid < 1) grid = 1;
    const size_t smem = sizeof(StepQueues) * (FLIGHT_THREADS / 32);
    static bool optin[16][2];
    const int dev = ctx->device & 15;
    cudaError_t e;
    ctx->launches++;
    if (field_mode == EMC_FIELD_CONSTANT) {
        auto k = flight_steps_kernel<EMC_FIELD_CONSTANT>;
        if (!optin[dev][0]) {
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            optin[dev][0] = true;
        }
        k<<<grid, FLIGHT_THREADS, smem, st>>>(ctx->dc, ctx->mech, A, n_steps, ctx->d_sor_flags + 3);
    } else {
        auto k = flight_steps_kernel<EMC_FIELD_GROUPS>;
        if (!optin[dev][1]) {
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            optin[dev][1] = true;
        }
        k<<<grid, FLIGHT_THREADS, smem, st>>>(ctx->dc, ctx->mech, A, n_steps, ctx->d_sor_flags + 3);
    }
    return cudaGetLastError();
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
    acc.zero(nullptr);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // four particles per iteration: their 20 loads are in flight together (the kernel streams
    // 36 B per particle and is bound by memory-level parallelism otherwise)
    constexpr int U = 4;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        double a[U], b[U], c[U], zz[U];
        int v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            const bool ok = i < n;
            a[u] = ok ? kx[i] : 0.0; b[u] = ok ? ky[i] : 0.0; c[u] = ok ? kz[i] : 0.0; zz[u] = ok ? z[i] : 0.0;
            v[u] = ok ? valley[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * stride >= n) break;
            int val = v[u], layer = 0;
            if (P.n_layers > 1 && val >= 0) {     // main_.py:393
                layer = where_am_i(P, zz[u]);
                if (layer < 0) { val = -(1 + EMC_FAULT_OUTSIDE); layer = 0; }
            }
            obs_particle(P, P.mat[layer], acc, a[u], b[u], c[u], zz[u], val);
        }
    }
    obs_block_reduce_and_finish(acc, partials, ticket, out, blockIdx.x, gridDim.x, true);
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
        double pz;
        int layer = 0;
        {
            const double u = rng.next();
            if (P.n_layers > 1 && P.layer_cw[P.n_layers - 1] > 0) {
                // extension (doping profiles): piecewise-uniform z, layer l holds the share
                // cw[l] - cw[l-1] of the ensemble; inverse CDF of the one uniform
                double c0 = 0.0, z0 = 0.0;
                while (layer < P.n_layers - 1 && u >= P.layer_cw[layer]) {
                    c0 = P.layer_cw[layer];
                    z0 += P.layer_lz[layer];
                    ++layer;
                }
                pz = z0 + P.layer_lz[layer] * ((u - c0) / (P.layer_cw[layer] - c0));
            } else {
                pz = P.tot_dim[2] * u;
            }
            if (P.n_layers > 1) {                 // init_.py:214: the material of the layer at z
                layer = where_am_i(P, pz);
                if (layer < 0) layer = P.n_layers - 1;
            }
        }
        double n1, n2, n3, n4;
        { const double a = rng.next(), b = rng.next(); box_muller(a, b, n1, n2); }
        { const double a = rng.next(), b = rng.next(); box_muller(a, b, n3, n4); }
        const double e = mean_energy * kT + 1E-7 * sqrt(n1 * n1 + n2 * n2 + n3 * n3);
        const double k = sqrt(2 * P.mat[layer].mass[0] * P.q * e / (P.hbar * P.hbar));      // init_.py:215
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
// photo-excited carriers: init_.init_photoex (init_.py:244-280) + main_.py:349-351 (valley 0,
// first free flight from free_flight(1E15)).  Of n_cand candidates the ones drawn beyond the slab
// are dropped (init_.py:247-248) and the survivors are appended in candidate order (deterministic
// compaction: per-block counts, one-block scan, ordered write).  Draws of candidate c (Philox step
// word 0xFFFFFFFE, particle word = cand_offset + c):
//   0      z = -scale * log(1 - u)                      np.random.exponential(scale)   (init_.py:246)
//   1,2    x, y = std * BoxMuller(u1, u2)               np.random.normal(0, laser_std) (:249,:251)
//   3,4    alpha, beta = 2 pi * BoxMuller(u3, u4)       Normal(0, 2 pi)                (:264-265)
//   5      dtau = (-1/G_first) * log(u5)                free_flight(1E15)              (main_.py:351)
// ------------------------------------------------------------------------------------------
struct PhotoexArgs {
    double *kx, *ky, *kz, *x, *y, *z, *dtau;
    int32_t *valley;
    int64_t n_cand;
    double energy, z_scale, xy_std, neg_inv_gamma_first;
    uint64_t seed, cand_offset;
    int64_t *block_counts;       // [gridDim.x + 1]: counts, then (after the scan) exclusive offsets; last = total
};

__device__ __forceinline__ double photoex_depth(const PhotoexArgs &Q, int64_t c)
{
    PhiloxStream rng;
    rng.init(Q.seed, Q.cand_offset + (uint64_t)c, 0xFFFFFFFEu);
    return -Q.z_scale * log(1 - rng.next());
}

constexpr int PHOTOEX_THREADS = 256;
constexpr int PHOTOEX_PER_BLOCK = 4096;

__global__ void __launch_bounds__(PHOTOEX_THREADS)
photoex_count_kernel(const __grid_constant__ DevConst P, const __grid_constant__ PhotoexArgs Q)
{
    __shared__ int warp_cnt[PHOTOEX_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * PHOTOEX_PER_BLOCK;
    int cnt = 0;
    for (int t = threadIdx.x; t < PHOTOEX_PER_BLOCK; t += PHOTOEX_THREADS) {
        const int64_t c = base + t;
        if (c < Q.n_cand) cnt += photoex_depth(Q, c) < P.tot_dim[2];          // init_.py:247
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, off);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < PHOTOEX_THREADS / 32; ++w) s += warp_cnt[w];
        Q.block_counts[blockIdx.x] = s;
    }
}

__global__ void photoex_scan_kernel(int64_t *block_counts, int n_blocks)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int b = 0; b < n_blocks; ++b) {
            const int64_t v = block_counts[b];
            block_counts[b] = run;
            run += v;
        }
        block_counts[n_blocks] = run;
    }
}

__global__ void __launch_bounds__(PHOTOEX_THREADS)
photoex_write_kernel(const __grid_constant__ DevConst P, const __grid_constant__ PhotoexArgs Q)
{
    __shared__ int warp_off[PHOTOEX_THREADS / 32];
    __shared__ int chunk_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * PHOTOEX_PER_BLOCK;
    if (threadIdx.x == 0) chunk_base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < PHOTOEX_PER_BLOCK; t0 += PHOTOEX_THREADS) {
        const int64_t c = base + t0 + threadIdx.x;
        double pz = 0.0;
        bool keep = false;
        if (c < Q.n_cand) {
            pz = photoex_depth(Q, c);
            keep = pz < P.tot_dim[2];
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_off[warp] = __popc(m);
        __syncthreads();
        int before = chunk_base;
        for (int w = 0; w < warp; ++w) before += warp_off[w];
        if (keep) {
            const int64_t o = Q.block_counts[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
            PhiloxStream rng;
            rng.init(Q.seed, Q.cand_offset + (uint64_t)c, 0xFFFFFFFEu);
            rng.next();                                           // draw 0: the depth above
            double n0, n1, na, nb;
            { const double a = rng.next(), b = rng.next(); box_muller(a, b, n0, n1); }
            { const double a = rng.next(), b = rng.next(); box_muller(a, b, na, nb); }
            int layer = 0;
            if (P.n_layers > 1) { layer = where_am_i(P, pz); if (layer < 0) layer = P.n_layers - 1; }   // init_.py:262
            const double k = sqrt(2 * P.mat[layer].mass[0] * P.q * Q.energy / (P.hbar * P.hbar));     // :263
            const double alpha = 2 * 3.14159265358979323846 * na, beta = 2 * 3.14159265358979323846 * nb;
            double sa, ca, sb, cb;
            sincos(alpha, &sa, &ca);
            sincos(beta, &sb, &cb);
            Q.kx[o] = k * ca * sb;                                // :266-268
            Q.ky[o] = k * sa * sb;
            Q.kz[o] = k * cb;
            Q.x[o] = Q.xy_std * n0;
            Q.y[o] = Q.xy_std * n1;
            Q.z[o] = pz;
            Q.dtau[o] = Q.neg_inv_gamma_first * log(rng.next());  // main_.py:351
            Q.valley[o] = 0;                                      // main_.py:349
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int w = 0; w < PHOTOEX_THREADS / 32; ++w) s += warp_off[w];
            chunk_base += s;
        }
        __syncthreads();
    }
}

cudaError_t launch_init_photoex(Ctx *ctx, double *kx, double *ky, double *kz, double *x, double *y, double *z,
                                double *dtau, int32_t *valley, int64_t n_cand, double energy, double z_scale,
                                double xy_std, double gamma_first, uint64_t seed, uint64_t cand_offset,
                                int64_t *n_accepted_host, cudaStream_t st)
{
    const int n_blocks = (int)((n_cand + PHOTOEX_PER_BLOCK - 1) / PHOTOEX_PER_BLOCK);
    cudaError_t e;
    int64_t *counts = nullptr;
    if ((e = cudaMallocAsync(&counts, sizeof(int64_t) * (size_t)(n_blocks + 1), st)) != cudaSuccess) return e;
    PhotoexArgs Q;
    Q.kx = kx; Q.ky = ky; Q.kz = kz; Q.x = x; Q.y = y; Q.z = z; Q.dtau = dtau; Q.valley = valley;
    Q.n_cand = n_cand; Q.energy = energy; Q.z_scale = z_scale; Q.xy_std = xy_std;
    Q.neg_inv_gamma_first = -1 / gamma_first;
    Q.seed = seed; Q.cand_offset = cand_offset; Q.block_counts = counts;
    ctx->launches += 3;
    photoex_count_kernel<<<n_blocks, PHOTOEX_THREADS, 0, st>>>(ctx->dc, Q);
    photoex_scan_kernel<<<1, 32, 0, st>>>(counts, n_blocks);
    photoex_write_kernel<<<n_blocks, PHOTOEX_THREADS, 0, st>>>(ctx->dc, Q);
    if ((e = cudaGetLastError()) == cudaSuccess)
        e = cudaMemcpyAsync(n_accepted_host, counts + n_blocks, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(counts, st);
    return e;
}

// ------------------------------------------------------------------------------------------
// single-function batches
// ------------------------------------------------------------------------------------------
__global__ void calc_energy_kernel(const __grid_constant__ DevConst P, int layer, const double *kx, const double *ky,
                                   const double *kz, const int32_t *valley, int64_t n, double *e_np,
                                   int32_t *idx, int32_t *fault)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e = 0.0;
    int32_t j = 0;
    const int f = calc_energy(P, P.mat[layer], kx[i], ky[i], kz[i], valley[i], e, j);
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
    carrier_drift(P, a, b, c, px, py, pz, ef3[3 * i], ef3[3 * i + 1], ef3[3 * i + 2], dt2[i], dt2[i] / mass[i]);
    kx[i] = a; ky[i] = b; kz[i] = c;
    x[i] = px; y[i] = py; z[i] = pz;
}

template <bool EXACT>
__global__ void scatter_kernel(const __grid_constant__ DevConst P, int layer, const __grid_constant__ MechMeta Mg, double *kx,
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
    const int f = scatter_event<EXACT>(P, P.mat[layer], M, a, b, c, val, rng, &mech);
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
    calc_energy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->dc, ctx->batch_layer, kx, ky, kz, valley, n,
                                                                    e_np, idx, fault);
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
    const int L = ctx->batch_layer;       // emc_select_layer: the mat_i argument of scatter_.scatter
    if (ctx->exact_libm)
        scatter_kernel<true><<<grid, 128, 0, st>>>(ctx->dc, L, ctx->mech_layers[L], kx, ky, kz, valley, u3, n, mech,
                                                   n_draws, fault);
    else
        scatter_kernel<false><<<grid, 128, 0, st>>>(ctx->dc, L, ctx->mech_layers[L], kx, ky, kz, valley, u3, n, mech,
                                                    n_draws, fault);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// self-test of the branch-free sqrt / division / log (emc_device.cuh) against CUDA's own
// correctly rounded sqrt(), `/` and its log(): bit-for-bit, on operands drawn from Philox over
// the ranges the kernels use and well beyond.  counts[0..2] = mismatches of div, sqrt, log.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double selftest_operand(uint32_t a, uint32_t b, int e_lo, int e_span, bool neg)
{
    // mantissa from 52 random bits, exponent uniform in [e_lo, e_lo + e_span), random sign if asked
    const uint64_t mant = (((uint64_t)a << 32) | b) & 0x000fffffffffffffULL;
    const int ex = e_lo + (int)((a >> 20) % (uint32_t)e_span);
    const uint64_t sign = (neg && (b & 1u)) ? 0x8000000000000000ULL : 0ULL;
    return __longlong_as_double((long long)(sign | ((uint64_t)(ex + 1023) << 52) | mant));
}

__global__ void selftest_math_kernel(int64_t n, uint64_t seed, unsigned long long *counts)
{
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    unsigned long long bad_div = 0, bad_sqrt = 0, bad_log = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 r0 = philox4x32_10(make_uint4(0u, 0x5e1f7e57u, (uint32_t)i, (uint32_t)(i >> 32)), key);
        const uint4 r1 = philox4x32_10(make_uint4(1u, 0x5e1f7e57u, (uint32_t)i, (uint32_t)(i >> 32)), key);
        // division: wide range, and the kernel's own shape kz / |k|, kx / (|k| sin)
        {
            const double a = selftest_operand(r0.x, r0.y, -80, 160, true);
            const double b = selftest_operand(r0.z, r0.w, -80, 160, true);
            bad_div += __double_as_longlong(fast_div(a, b)) != __double_as_longlong(a / b);
            const double k0 = selftest_operand(r1.x, r1.y, 20, 16, false);
            const double kz = k0 * (2 * u53(r1.z, r1.w) - 1);
            bad_div += __double_as_longlong(fast_div(kz, k0)) != __double_as_longlong(kz / k0);
            bad_div += __double_as_longlong(fast_div(0.0, k0)) != __double_as_longlong(0.0 / k0);
        }
        // sqrt: wide range, |k|^2, 1 + 4 alpha E, (1 - c)(1 + c) down to the last ulp below 1
        {
            const double a = selftest_operand(r0.x, r0.z, -200, 400, false);
            bad_sqrt += __double_as_longlong(fast_sqrt(a)) != __double_as_longlong(sqrt(a));
            const double c = 2 * u53(r1.x, r1.w) - 1;
            const double w = (1 - c) * (1 + c);
            bad_sqrt += __double_as_longlong(fast_sqrt(w)) != __double_as_longlong(sqrt(w));
            const double cc = 1 - ldexp(u53(r1.y, r1.z), -(int)(r0.w % 50u));          // c -> 1
            const double ww = (1 - cc) * (1 + cc);
            bad_sqrt += __double_as_longlong(fast_sqrt(ww)) != __double_as_longlong(sqrt(ww));
            const double e = 1 + 4 * 0.61 * (2 * u53(r0.y, r1.y));
            bad_sqrt += __double_as_longlong(fast_sqrt(e)) != __double_as_longlong(sqrt(e));
        }
        // log: the uniforms of free_flight and general positive normal numbers
        {
            const double u = u53(r0.x, r1.x);
            if (u > 0) bad_log += __double_as_longlong(fast_log(u)) != __double_as_longlong(log(u));
            const double v = ldexp(u53(r0.z, r1.z), -(int)(r0.y % 53u));                 // small uniforms
            if (v > 0) bad_log += __double_as_longlong(fast_log(v)) != __double_as_longlong(log(v));
            const double a = selftest_operand(r0.w, r1.w, -300, 600, false);
            bad_log += __double_as_longlong(fast_log(a)) != __double_as_longlong(log(a));
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        // specials: zero and negative arguments
        bad_sqrt += __double_as_longlong(fast_sqrt(0.0)) != __double_as_longlong(0.0);
        bad_sqrt += !isnan(fast_sqrt(-1.0));
        bad_log += !(fast_log(0.0) < -1e308);
        bad_div += !isnan(fast_div(1.0, 0.0));
    }
    if (bad_div) atomicAdd(&counts[0], bad_div);
    if (bad_sqrt) atomicAdd(&counts[1], bad_sqrt);
    if (bad_log) atomicAdd(&counts[2], bad_log);
}

cudaError_t launch_selftest_math(Ctx *ctx, int64_t n, uint64_t seed, unsigned long long *counts_dev, cudaStream_t st)
{
    ctx->launches++;
    selftest_math_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(n, seed, counts_dev);
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
