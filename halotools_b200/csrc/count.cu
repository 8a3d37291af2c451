// K2 — the cell-pair kernel family (sm_100a).  Every kernel is a persistent grid of
// warps pulling sample1 tiles from an atomic queue and running walk_tile<V>() with a
// variant functor V:
//   Fast3    npairs_3d, <= 16 monotone bins: strict-IEEE f64 distance, integer-pipe range
//            test on the raw bit pattern, in-range pairs pushed as 32-bit monotone keys to a
//            per-lane shared-memory queue and binned later at full lane occupancy; tiles
//            that saw a key equal to an edge key are re-evaluated with exact 64-bit compares.
//   Gen3     npairs_3d literal top-down scan (any bins)          npairs_3d_engine.pyx:171-182
//   GenXYZ   npairs_xy_z literal nested scans                    npairs_xy_z_engine.pyx:178-194
//   GenSMU   npairs_s_mu (differential histogram)                npairs_s_mu_engine.pyx:196-229
//   Marked3  marked_npairs_3d, 17 weight functions               marked_npairs_3d_engine.pyx:204-216
//   DSigma   mean_delta_sigma per-object accumulators (2-D)      mean_delta_sigma_engine.pyx:162-180
#include <type_traits>
#include <vector>
#include <cstring>
#include "walk.cuh"
#include "count.cuh"
#include "kernel.cuh"

// ------------------------------------------------------------------ tile list
// One thread per fine COLUMN of sample1 (fixed slow-dimension fine indices; a contiguous run of the sorted arrays):
// the part of the column that lies in reference cells [first_cell1, last_cell1) is cut greedily into tiles of up to
// G.tile consecutive points that straddle at most G.maxspan reference cells and cover at most G.maxfine fine cells
// along the fast dimension.
// Tile descriptor: {first sorted index, column | (points << 24)}.
__device__ __forceinline__ uint32_t htb_column_tiles(const uint32_t *__restrict__ off1, const WalkGeom &G, int64_t col,
                                                     int64_t first_cell1, int64_t last_cell1, uint2 *out, uint32_t base)
{
    const int F = G.dim - 1;
    int64_t rem = col, rid = 0, mul = 1;
    // reference cell id of (slow dims of this column, fast index 0): last dimension fastest
    int fsl[2] = {0, 0};
    for (int d = F - 1; d >= 0; --d) { fsl[d] = (int)(rem % G.nf1[d]); rem /= G.nf1[d]; }
    for (int d = 0; d < F; ++d) rid = rid * G.nd1[d] + fsl[d] / G.m1[d];
    (void)mul;
    rid *= G.nd1[F];
    int64_t zlo64 = first_cell1 - rid, zhi64 = last_cell1 - rid;
    const int zlo = (int)(zlo64 < 0 ? 0 : (zlo64 > G.nd1[F] ? G.nd1[F] : zlo64));
    const int zhi = (int)(zhi64 < 0 ? 0 : (zhi64 > G.nd1[F] ? G.nd1[F] : zhi64));
    if (zlo >= zhi) return 0;
    const uint32_t *o = off1 + col * G.nf1[F];
    const int mf = G.m1[F];
    uint32_t pos = o[zlo * mf];
    const uint32_t end = o[zhi * mf];
    uint32_t n = 0;
    int f = zlo * mf;
    while (pos < end) {
        while (o[f + 1] <= pos) ++f;                          // fine cell (fast dim) of the tile's first point
        const int lim = min(min(zhi, f / mf + G.maxspan) * mf, f + G.maxfine);
        const uint32_t e = min(pos + (uint32_t)G.tile, o[lim]);
        if (out) out[base + n] = make_uint2(pos, (uint32_t)col | ((e - pos) << 24));
        ++n;
        pos = e;
    }
    return n;
}

// The same cut by ONE WARP per column, for meshes with few columns and many fine cells along the fast dimension
// (the bench's DD launch: 144 columns of 384 cells): a thread walking its column cell by cell waits a full L2 round
// trip per cell.  Here the warp keeps a window of the column's offsets in shared memory (coalesced loads), every lane
// runs the same walk on broadcast shared-memory reads, lane 0 writes.  The window always covers
// [f, f + G.maxfine + 1]: everything one tile needs.
#define HTB_TL_WIN 1024
#define HTB_TL_WARPS 4
__device__ __forceinline__ uint32_t htb_column_tiles_warp(const uint32_t *__restrict__ off1, const WalkGeom &G, int64_t col,
                                                          int64_t first_cell1, int64_t last_cell1, uint2 *out, uint32_t base,
                                                          uint32_t *so, int lane)
{
    const int F = G.dim - 1;
    int64_t rem = col, rid = 0;
    int fsl[2] = {0, 0};
    for (int d = F - 1; d >= 0; --d) { fsl[d] = (int)(rem % G.nf1[d]); rem /= G.nf1[d]; }
    for (int d = 0; d < F; ++d) rid = rid * G.nd1[d] + fsl[d] / G.m1[d];
    rid *= G.nd1[F];
    int64_t zlo64 = first_cell1 - rid, zhi64 = last_cell1 - rid;
    const int zlo = (int)(zlo64 < 0 ? 0 : (zlo64 > G.nd1[F] ? G.nd1[F] : zlo64));
    const int zhi = (int)(zhi64 < 0 ? 0 : (zhi64 > G.nd1[F] ? G.nd1[F] : zhi64));
    if (zlo >= zhi) return 0;
    const uint32_t *o = off1 + col * G.nf1[F];
    const int mf = G.m1[F], last = zhi * mf;              // offsets o[zlo * mf .. last] are the ones that matter
    int wb = zlo * mf;                                    // the window holds o[wb .. min(wb + WIN - 1, last)]
    auto load = [&](int from) {
        __syncwarp();
        wb = from;
        for (int i = lane; i < HTB_TL_WIN && wb + i <= last; i += 32) so[i] = o[wb + i];
        __syncwarp();
    };
    load(wb);
    uint32_t pos = so[0];
    const uint32_t end = o[last];
    uint32_t n = 0;
    int f = zlo * mf;
    while (pos < end) {
        while (true) {                                      // fine cell (fast dim) of the tile's first point
            if (f + 1 - wb >= HTB_TL_WIN) load(f);
            if (so[f + 1 - wb] > pos) break;
            ++f;
        }
        const int lim = min(min(zhi, f / mf + G.maxspan) * mf, f + G.maxfine);
        if (lim - wb >= HTB_TL_WIN) load(f);
        const uint32_t e = min(pos + (uint32_t)G.tile, so[lim - wb]);
        if (out && lane == 0) out[base + n] = make_uint2(pos, (uint32_t)col | ((e - pos) << 24));
        ++n;
        pos = e;
    }
    return n;
}

template <bool FILL>
__global__ void __launch_bounds__(32 * HTB_TL_WARPS)
k_tiles_warp(const uint32_t *__restrict__ off1, WalkGeom G, int64_t ncol, int64_t first_cell1, int64_t last_cell1,
             const long long *__restrict__ range, uint32_t *__restrict__ ntile, const uint32_t *__restrict__ tbase,
             uint2 *__restrict__ tiles)
{
    __shared__ uint32_t win[HTB_TL_WARPS][HTB_TL_WIN];
    if (range) { first_cell1 = range[0]; last_cell1 = range[1]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t s = (int64_t)blockIdx.x * HTB_TL_WARPS + warp; s < ncol; s += (int64_t)gridDim.x * HTB_TL_WARPS) {
        if (FILL) {
            if (ntile[s]) (void)htb_column_tiles_warp(off1, G, s, first_cell1, last_cell1, tiles, tbase[s], win[warp], lane);
        } else {
            const uint32_t n = htb_column_tiles_warp(off1, G, s, first_cell1, last_cell1, nullptr, 0u, win[warp], lane);
            if (lane == 0) ntile[s] = n;
        }
    }
}

__global__ void k_seg_tiles(const uint32_t *__restrict__ off1, WalkGeom G, int64_t ncol,
                            int64_t first_cell1, int64_t last_cell1, const long long *__restrict__ range,
                            uint32_t *__restrict__ ntile)
{
    if (range) { first_cell1 = range[0]; last_cell1 = range[1]; }      // this rank's shard, cut on the device
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < ncol; s += (int64_t)gridDim.x * blockDim.x)
        ntile[s] = htb_column_tiles(off1, G, s, first_cell1, last_cell1, nullptr, 0u);
}

__global__ void k_fill_tiles(const uint32_t *__restrict__ off1, WalkGeom G, int64_t ncol,
                             int64_t first_cell1, int64_t last_cell1, const long long *__restrict__ range,
                             const uint32_t *__restrict__ ntile, const uint32_t *__restrict__ tbase,
                             uint2 *__restrict__ tiles)
{
    if (range) { first_cell1 = range[0]; last_cell1 = range[1]; }
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < ncol; s += (int64_t)gridDim.x * blockDim.x)
        if (ntile[s]) (void)htb_column_tiles(off1, G, s, first_cell1, last_cell1, tiles, tbase[s]);
}

// pairs the reference loop nest visits, per reference cell1 (W_ref), and - for the multi-GPU cut -
// the pairs this engine expects to evaluate: in symmetric mode a zero-shift neighbour cell is evaluated only from
// the cell with the smaller id (half of the own cell), wrapped neighbours from both sides.
// `wtab` (optional): per neighbour-cell offset inside the window, the fraction of its pairs that survives the walker's
// pruning (pairs within the search length): the corner cells of a window cost far less than its face cells, and the
// x-layers a rank's range touches at the periodic boundary are corner-heavy (an unweighted cut gave rank 0 of 8 7 %
// fewer evaluated pairs than the others).
__global__ void k_wref(const uint32_t *__restrict__ rc1, const uint32_t *__restrict__ rc2, WalkGeom G,
                       int64_t ncell1, double *__restrict__ work, double *__restrict__ balance, const float *__restrict__ wtab)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell1; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t n1 = rc1[c];
        double w = 0.0, wb = 0.0;
        if (n1) {
            int a[3] = {0, 0, 0};
            int64_t rem = c;
            for (int d = G.dim - 1; d >= 0; --d) { a[d] = (int)(rem % G.nd1[d]); rem /= G.nd1[d]; }
            int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
            for (int d = 0; d < G.dim; ++d) { lo[d] = a[d] * G.per[d] - G.cover[d]; hi[d] = (a[d] + 1) * G.per[d] + G.cover[d]; }
            unsigned long long s = 0;
            double s2 = 0.0, sw = 0.0;                 // s2: twice the symmetric-mode pairs; both weighted by wtab
            const int wy_n = hi[1] - lo[1], wz_n = G.dim == 3 ? hi[2] - lo[2] : 1;
            for (int ux = lo[0]; ux < hi[0]; ++ux) {
                const int kx = floor_div(ux, G.nd2[0]), wx = ux - kx * G.nd2[0];
                for (int uy = lo[1]; uy < hi[1]; ++uy) {
                    const int ky = floor_div(uy, G.nd2[1]), wy = uy - ky * G.nd2[1];
                    if (G.dim == 2) {
                        const int64_t c2 = (int64_t)wx * G.nd2[1] + wy;
                        const unsigned long long n2 = rc2[c2];
                        const double f = wtab ? (double)wtab[(ux - lo[0]) * wy_n + (uy - lo[1])] : 1.0;
                        s += n2;
                        sw += f * (double)n2;
                        s2 += f * (double)(((kx | ky) != 0 || c2 > c) ? 2 * n2 : (c2 == c ? n2 : 0));
                        continue;
                    }
                    for (int uz = lo[2]; uz < hi[2]; ++uz) {
                        const int kz = floor_div(uz, G.nd2[2]), wz = uz - kz * G.nd2[2];
                        const int64_t c2 = ((int64_t)wx * G.nd2[1] + wy) * G.nd2[2] + wz;
                        const unsigned long long n2 = rc2[c2];
                        const double f = wtab ? (double)wtab[((ux - lo[0]) * wy_n + (uy - lo[1])) * wz_n + (uz - lo[2])] : 1.0;
                        s += n2;
                        sw += f * (double)n2;
                        s2 += f * (double)(((kx | ky | kz) != 0 || c2 > c) ? 2 * n2 : (c2 == c ? n2 : 0));
                    }
                }
            }
            w = (double)n1 * (double)s;
            wb = G.sym ? (double)n1 * s2 : 2.0 * (double)n1 * sw;
        }
        work[c] = w;
        if (balance) balance[c] = wb;
    }
}

// Multi-GPU shard of a call (htb_set_shard): rank r of `world` takes the contiguous run of reference mesh1 cells of
// [first, last) whose exclusive cumulative predicted work lies in [total * r / world, total * (r + 1) / world) - the
// reference's contiguous cell ranges (mesh_helpers.py:183-221) with the cut points moved so that every rank gets
// the same number of pair evaluations instead of the same number of cells.  One block; range_out = {first, last}.
// 256 threads, not 1024: at 39 registers a 1024-thread block needs 40 K registers of ONE SM, which no SM has free while a
// persistent count kernel of another stream is resident (two 32 K-register blocks per SM) - the set-up of the second and
// third count of a sharded statistic then waited for the first count to END (timeline in DESIGN section 5).
#define HTB_SR_THREADS 256
__global__ void __launch_bounds__(HTB_SR_THREADS) k_shard_range(const double *__restrict__ work, long long first, long long last,
                                                       int rank, int world, long long *__restrict__ range_out)
{
    __shared__ double part[HTB_SR_THREADS];
    __shared__ long long below[2];
    const int t = threadIdx.x;
    const long long n = last > first ? last - first : 0;
    const long long per = (n + HTB_SR_THREADS - 1) / HTB_SR_THREADS;
    const long long a = first + min(n, per * t), b = first + min(n, per * (t + 1));
    double s = 0.0;
    for (long long c = a; c < b; ++c) s += work[c];
    part[t] = s;
    if (t < 2) below[t] = 0;
    __syncthreads();
    // exclusive scan of the partial sums (the values are integers below 2^53: the sums are exact, so
    // the order of the additions does not matter)
    for (int o = 1; o < HTB_SR_THREADS; o <<= 1) {
        const double v = t >= o ? part[t - o] : 0.0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const double total = part[HTB_SR_THREADS - 1];
    double cum = part[t] - s;
    const double lo = total * (double)rank / (double)world, hi = total * (double)(rank + 1) / (double)world;
    long long nlo = 0, nhi = 0;
    for (long long c = a; c < b; ++c) {
        nlo += cum < lo;              // cells that belong to lower ranks
        nhi += cum < hi;              // ... to this rank or lower ranks
        cum += work[c];
    }
    if (nlo) atomicAdd((unsigned long long *)&below[0], (unsigned long long)nlo);
    if (nhi) atomicAdd((unsigned long long *)&below[1], (unsigned long long)nhi);
    __syncthreads();
    if (t == 0) {
        range_out[0] = first + below[0];
        range_out[1] = rank + 1 >= world ? last : first + below[1];
    }
}


// ------------------------------------------------------------------ Fast3
// npairs_3d with <= 16 monotone bins.  Per evaluated pair the hot loop spends, besides the 8 strict
// f64 operations of the reference (npairs_3d_engine.pyx:173-176), five integer instructions:
//   key  = (bits(dsq) >> 26) - (bits(top edge) >> 26)      one LEA.HI; 32-bit, wraps harmlessly
//   ctop += key >> 31                                      one LEA.HI: pairs certainly inside the top edge
//   if (key <= F[second edge from the top]) push key       ISETP + predicated STS + pointer bump
// plus two 3-input minimum trackers per TWO pairs that catch the cases 32 bits cannot decide: a key
// EQUAL to the top-edge key (umin == 0) and a separation so small (or exactly zero) that the key
// would wrap (hmin < Hwin).  The trackers are tested together with the queue-full test once per
// group of 16 pairs; when one fires the group's pushes are rolled back and the group is re-evaluated
// with exact 64-bit compares.  Queued keys (12 % of the pairs for log-spaced bins) are binned later at
// full lane occupancy by a compaction cascade: flush1() applies one more edge and keeps the survivors
// at the bottom of the queue, deep() runs the remaining edges only when enough survivors have piled
// up.  The cumulative count of an edge is the number of keys that survived it — the reference's
// top-down scan (npairs_3d_engine.pyx:178-182).  A queued key equal to an edge key marks the tile
// dirty: its counts are discarded and the tile is re-evaluated exactly.
#ifndef QCAP
#define QCAP 72               // queue slots per lane
#endif
#ifndef QSURV
#define QSURV 24              // run the deep cascade when a lane holds more survivors than this
#endif
#ifndef FAST3_WARPS
#define FAST3_WARPS 8
#define FAST3_MINBLOCKS 2
#endif
#ifndef FAST3_PPL
#define FAST3_PPL 2
#endif
#define QGROUP 16             // pairs per lane between two queue checks

__device__ __forceinline__ void lds_f64x2_tok(uint32_t addr, uint32_t tok, double &a, double &b)
{
    // not volatile: the scheduler may hoist these loads over the queue stores.  `tok` changes with every
    // staged chunk, so loads of different chunks are never merged.
    asm("ld.shared.v2.f64 {%0, %1}, [%2]; // %3" : "=d"(a), "=d"(b) : "r"(addr), "r"(tok));
}
__device__ __forceinline__ void sts_u32_nc(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}

template <int PPL_, int KIND = 0>      // KIND 0: npairs_3d;  1: npairs_xy_z with one or two pi edges
struct Fast3T {
    static constexpr int DIM = 3, NPAY = 0, PPL = PPL_, WARPS = FAST3_WARPS, MINBLOCKS = FAST3_MINBLOCKS;
    static constexpr bool TMA = true;
    static constexpr int GJ = QGROUP / PPL;     // sample2 points per group
    static constexpr int TOP = HTB_NBF - 1;
    static constexpr int QC = KIND == 1 ? QCAP - 16 : QCAP;       // (rp, pi): 16 rows go to the lower-pi-edge counters
    static constexpr uint32_t QFULL = 128u * (QC - QGROUP - PPL);     // flush when a lane's queue is longer than this
    static constexpr bool HAS_SELF = true;
    typedef Fast3Params Params;
    const Params &P;
    uint32_t qbase;             // shared-space address of this lane's queue column
    uint32_t qs;                // end of the survivor region = start of the keys pushed since flush1()
    uint32_t qptr;              // next free slot (qbase + 128 * entries)
    uint32_t qsave;             // qptr at the start of the current group (roll-back point)
    int lane;
    uint32_t idx[PPL];          // sorted positions of this lane's points (coordinates are re-read when the shift changes)
    unsigned valmask;
    double xs[PPL], ys[PPL], zs[PPL];   // this lane's points, periodic shift applied
    double sentinel;
    unsigned c[HTB_NBF];
    unsigned ctop, csave;
    unsigned zself;             // (rp, pi): exact-zero separations met by the exact path (self pairs, duplicates)
    unsigned umin;
    int hmin;
    int hzmin;                  // (rp, pi): smallest dz^2 high word since the last check
    uint32_t c0base;            // (rp, pi): shared-space address of this lane's 16 counters of pairs that also lie
                                // inside the LOWER pi edge (rare by construction, always decided by the exact path)
    unsigned wt_now;
    unsigned long long tot0;
    bool exact, dirty, always_exact;
    unsigned long long tot;

    static size_t scratch_bytes(const Params &) { return sizeof(uint32_t) * QCAP * 32; }
    __device__ __forceinline__ void tile_weight(unsigned wt) { wt_now = wt; }
    __device__ __forceinline__ void force_exact() { exact = true; }

    __device__ __forceinline__ Fast3T(const Params &p, void *scratch, int ln, const WalkArrays &A) : P(p), lane(ln)
    {
        qbase = smem_u32(scratch) + 4u * (uint32_t)ln;
        qs = qptr = qsave = qbase;
        tot = 0; ctop = csave = 0; zself = 0; umin = 0xffffffffu; hmin = 0x7fffffff; hzmin = 0x7fffffff; wt_now = 1; tot0 = 0;
        c0base = qbase + 128u * QC;
        if (KIND == 1) { for (int k = 0; k < HTB_NBF; ++k) sts_u32(c0base + 128u * k, 0u); }
        exact = dirty = false;
        valmask = 0; sentinel = 0.0;
        // a point outside [0, period] can be arbitrarily far away: the 32-bit keys could wrap
        always_exact = ((A.flags1[0] | A.flags2[0]) & 1u) != 0u;
#pragma unroll
        for (int s = 0; s < HTB_NBF; ++s) c[s] = 0;
#pragma unroll
        for (int q = 0; q < PPL; ++q) { idx[q] = 0; xs[q] = ys[q] = zs[q] = 0.0; }
    }
    __device__ __forceinline__ void tile_begin(const double (&p)[PPL][3], const bool (&val)[PPL], const uint32_t (&id)[PPL],
                                               const WalkArrays &)
    {
        valmask = 0;
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            idx[q] = id[q];
            if (val[q]) valmask |= 1u << q; else sentinel = p[q][0];
        }
        exact = always_exact; dirty = false;
    }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &A)
    {
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            const double bx = ((valmask >> q) & 1u) ? A.c1[0][idx[q]] : sentinel;
            xs[q] = bx - sh[0];
            ys[q] = A.c1[1][idx[q]] - sh[1];
            zs[q] = A.c1[2][idx[q]] - sh[2];
        }
    }
    // Compaction pass over this lane's keys in [r, end): keys <= Fk are kept (moved down to w), a key equal to Fin
    // (the edge that admitted the keys) marks the tile dirty.  Four independent loads per trip hide the latency.
    __device__ __forceinline__ uint32_t compact(uint32_t r, uint32_t end, uint32_t w, int Fk, int Fin)
    {
        bool eq = false;
        for (; r + 512u <= end; r += 512u) {
            const int k0 = (int)lds_u32(r), k1 = (int)lds_u32(r + 128u), k2 = (int)lds_u32(r + 256u), k3 = (int)lds_u32(r + 384u);
            eq |= (k0 == Fin) | (k1 == Fin) | (k2 == Fin) | (k3 == Fin);
            if (k0 <= Fk) { sts_u32_nc(w, (uint32_t)k0); w += 128u; }
            if (k1 <= Fk) { sts_u32_nc(w, (uint32_t)k1); w += 128u; }
            if (k2 <= Fk) { sts_u32_nc(w, (uint32_t)k2); w += 128u; }
            if (k3 <= Fk) { sts_u32_nc(w, (uint32_t)k3); w += 128u; }
        }
        for (; r != end; r += 128u) {
            const int k0 = (int)lds_u32(r);
            eq |= (k0 == Fin);
            if (k0 <= Fk) { sts_u32_nc(w, (uint32_t)k0); w += 128u; }
        }
        dirty |= eq;
        return w;
    }
    // one cascade level: the keys in [qbase, end) were admitted by F[S + 1]; keep those <= F[S]
    // (S == -1: nothing left to apply, only look for a key equal to F[0])
    template <int S>
    __device__ __forceinline__ void cascade(uint32_t end)
    {
        if (!__any_sync(HTB_FULL, end != qbase)) return;
        if constexpr (S >= 0) {
            const uint32_t w = compact(qbase, end, qbase, P.F[S], P.F[S + 1]);
            c[S] += (w - qbase) >> 7;
            cascade<S - 1>(w);
        } else {
            (void)compact(qbase, end, qbase, (int)0x80000000, P.F[0]);
        }
    }
    __device__ __forceinline__ void deep()
    {
        cascade<TOP - 3>(qs);
        qs = qptr = qsave = qbase;
    }
    // bin the keys pushed since the last call against one more edge; survivors stay queued
    __device__ __forceinline__ void flush1(bool force_deep)
    {
        const uint32_t w = compact(qs, qptr, qs, P.F[TOP - 2], P.F[TOP - 1]);
        c[TOP - 1] += (qptr - qs) >> 7;
        c[TOP - 2] += (w - qs) >> 7;
        qs = qptr = qsave = w;
        if (force_deep || __any_sync(HTB_FULL, w > qbase + 128u * QSURV)) deep();
    }
    __device__ __forceinline__ void key_of(double x1s, double y1s, double z1s, double xj, double yj, double zj, int &key, int &hi)
    {
        const double dx = x1s - xj, dy = y1s - yj, dz = z1s - zj;
        if (KIND == 0) {
            const double dsq = dx * dx + dy * dy + dz * dz;
            hi = __double2hiint(dsq);
            // (bits >> 26) + nbias, written as the high word of a left shift so that it maps to one LEA.HI
            key = (int)(__funnelshift_l((unsigned)__double2loint(dsq), (unsigned)hi, 6) + (unsigned)P.nbias);
        } else {
            // npairs_xy_z_engine.pyx:180-183: dxy_sq = dx*dx + dy*dy, dz_sq = dz*dz
            const double dxy_sq = dx * dx + dy * dy;
            const double dz_sq = dz * dz;
            hi = __double2hiint(dxy_sq);
            hzmin = min(hzmin, __double2hiint(dz_sq));
            const int k = (int)(__funnelshift_l((unsigned)__double2loint(dxy_sq), (unsigned)hi, 6) + (unsigned)P.nbias);
            key = (dz_sq <= P.pi_top_sq) ? k : 0x7fffffff;       // exact f64 compare: outside the top pi edge = out of range
        }
    }
    __device__ __forceinline__ void push(int key)
    {
        ctop += (unsigned)key >> 31;
        if (key <= P.F[TOP - 1]) { sts_u32_nc(qptr, (uint32_t)key); qptr += 128u; }
    }
    // one point of this lane against two staged points
    __device__ __forceinline__ void pair2(int q, double xa, double ya, double za, double xb, double yb, double zb)
    {
        int ka, kb, ha, hb;
        key_of(xs[q], ys[q], zs[q], xa, ya, za, ka, ha);
        key_of(xs[q], ys[q], zs[q], xb, yb, zb, kb, hb);
        umin = min(umin, min((unsigned)ka, (unsigned)kb));
        hmin = min(hmin, min(ha, hb));
        push(ka);
        push(kb);
    }
    __device__ __forceinline__ void pair_fast(int q, double xj, double yj, double zj)
    {
        int k, h;
        key_of(xs[q], ys[q], zs[q], xj, yj, zj, k, h);
        umin = min(umin, (unsigned)k);
        hmin = min(hmin, h);
        push(k);
    }
    __device__ __forceinline__ double edge(int s) const { return __longlong_as_double((long long)P.E[s]); }
    // c_s += (d <= e_s) for four levels: one DSETP and one predicated integer add each (the C form comes back from
    // ptxas as add / predicated move / move per level)
    static __device__ __forceinline__ void count4(unsigned &c0, unsigned &c1, unsigned &c2, unsigned &c3, double e0, double e1,
                                                  double e2, double e3, double d)
    {
        asm("{\n\t.reg .pred p;\n\t"
            "setp.le.f64 p, %8, %4;\n\t@p add.u32 %0, %0, 1;\n\t"
            "setp.le.f64 p, %8, %5;\n\t@p add.u32 %1, %1, 1;\n\t"
            "setp.le.f64 p, %8, %6;\n\t@p add.u32 %2, %2, 1;\n\t"
            "setp.le.f64 p, %8, %7;\n\t@p add.u32 %3, %3, 1;\n\t}"
            : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3) : "d"(e0), "d"(e1), "d"(e2), "d"(e3), "d"(d));
    }
    __device__ __forceinline__ void count_levels(double d)
    {
        static_assert(HTB_NBF == 16, "four groups of four levels");
#pragma unroll
        for (int s = 0; s < HTB_NBF; s += 4) count4(c[s], c[s + 1], c[s + 2], c[s + 3], edge(s), edge(s + 1), edge(s + 2), edge(s + 3), d);
    }
    __device__ __forceinline__ void pair_exact(int q, double xj, double yj, double zj)
    {
        const double dx = xs[q] - xj, dy = ys[q] - yj, dz = zs[q] - zj;
        if (KIND == 0) {
            // f64 compares (DSETP, FP64 pipe) order the non-negative values here exactly like 64-bit integer compares of
            // the bit patterns (two ISETP each on the half-rate ALU pipe)
            const double dsq = dx * dx + dy * dy + dz * dz;
            if (dsq <= edge(HTB_NBF - 1)) count_levels(dsq);
        } else {
            const double dxy_sq = dx * dx + dy * dy;
            const double dz_sq = dz * dz;
            const unsigned long long b = (unsigned long long)__double_as_longlong(dxy_sq);
            // an exact zero (a self pair of the tile's own range in symmetric mode, a duplicate) lies inside every rp edge
            // (the fast path requires bins >= 0) and inside both pi edges: counted once here, added to every counter at the
            // end of the tile - the per-edge loops below ran with ONE active lane for each of them (5 % of config 3)
            if ((b | (unsigned long long)__double_as_longlong(dz_sq)) == 0ULL) { zself += 1u; return; }
            if (dxy_sq <= edge(HTB_NBF - 1) && dz_sq <= P.pi_top_sq) {
                count_levels(dxy_sq);
                if (P.counts0 && (unsigned long long)__double_as_longlong(dz_sq) <= P.Epi0) {
                    // also inside the lower pi edge
#pragma unroll 1
                    for (int s = 0; s < HTB_NBF; ++s)
                        if (b <= P.E[s]) sts_u32(c0base + 128u * s, lds_u32(c0base + 128u * s) + 1u);
                }
            }
        }
    }
    __device__ __forceinline__ void exact_range(uint32_t stage, int j0, int j1)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH;
#pragma unroll 1
        for (int j = j0; j < j1; ++j) {
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j);
#pragma unroll
            for (int q = 0; q < PPL; ++q) pair_exact(q, xj, yj, zj);
        }
    }
    // after every group of <= QGROUP + PPL pairs per lane: staged entries [j0, j1) since the last check
    __device__ __forceinline__ void check(uint32_t stage, int j0, int j1)
    {
        const bool undecided = (umin == 0u) | (hmin < P.Hwin) | (KIND == 1 && hzmin <= P.Hz0);
        const bool full = qptr > qbase + QFULL;
        if (__any_sync(HTB_FULL, undecided | full)) {
            if (__any_sync(HTB_FULL, undecided)) {
                // some pair of this group cannot be decided from its 32-bit key (or may lie inside the lower pi
                // edge): take the whole group back
                qptr = qsave; ctop = csave;
                exact_range(stage, j0, j1);
                umin = 0xffffffffu; hmin = 0x7fffffff; hzmin = 0x7fffffff;
            }
            if (__any_sync(HTB_FULL, qptr > qbase + QFULL)) flush1(false);
        }
        qsave = qptr; csave = ctop;
    }
    // own: the tile's own index range in symmetric mode.  Every group of it holds a self pair, i.e. would be pushed, taken
    // back and re-evaluated exactly - and its pairs are close, nearly all of them in range, where counting against the
    // edges directly is cheaper than the queue (measured: sending them through the queue costs +14 % on config 4).
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t tok, bool own)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH;
        if (exact | own) { exact_range(stage, lo, hi); return; }
        int j = lo;
        if ((j & 1) && j < hi) {
            // odd leading entry: evaluated alone, checked together with the group that follows
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j);
#pragma unroll
            for (int q = 0; q < PPL; ++q) pair_fast(q, xj, yj, zj);
            ++j;
        }
        int j0 = lo;
        if (j + GJ <= hi) {
            // register double buffering: the loads of the next two staged points are issued before the
            // queue stores of the current ones (the hardware keeps shared loads behind earlier stores)
            double xa, xb, ya, yb, za, zb;
            lds_f64x2_tok(bx + 8 * j, tok, xa, xb);
            lds_f64x2_tok(by + 8 * j, tok, ya, yb);
            lds_f64x2_tok(bz + 8 * j, tok, za, zb);
#pragma unroll 1
            for (; j + GJ <= hi; j += GJ) {
#pragma unroll
                for (int u = 0; u < GJ; u += 2) {
                    double xc, xd, yc, yd, zc, zd;
                    lds_f64x2_tok(bx + 8 * (j + u + 2), tok, xc, xd);     // may run past hi: harmless, never used
                    lds_f64x2_tok(by + 8 * (j + u + 2), tok, yc, yd);
                    lds_f64x2_tok(bz + 8 * (j + u + 2), tok, zc, zd);
#pragma unroll
                    for (int q = 0; q < PPL; ++q) pair2(q, xa, ya, za, xb, yb, zb);
                    xa = xc; xb = xd; ya = yc; yb = yd; za = zc; zb = zd;
                }
                check(stage, j0, j + GJ);
                j0 = j + GJ;
            }
        }
        if (j0 < hi) {
#pragma unroll 1
            for (; j < hi; ++j) {
                const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j);
#pragma unroll
                for (int q = 0; q < PPL; ++q) pair_fast(q, xj, yj, zj);
            }
            check(stage, j0, hi);
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&)[PPL], int pass, unsigned wt)
    {
        if (!exact) {
            flush1(true);
            c[TOP] += ctop;
            ctop = csave = 0;
            if (__any_sync(HTB_FULL, dirty) && pass == 0) {
                // a queued key collided with an edge key: throw the tile's counts away and redo it exactly
#pragma unroll
                for (int s = 0; s < HTB_NBF; ++s) c[s] = 0;
                if (KIND == 1) { for (int k = 0; k < HTB_NBF; ++k) sts_u32(c0base + 128u * k, 0u); }
                dirty = false; exact = true;
                zself = 0;
                return true;
            }
        }
        if (KIND == 1) {
            // exact zeros met by the exact path: inside every rp edge and inside the lower pi edge
#pragma unroll
            for (int s = 0; s < HTB_NBF; ++s) c[s] += zself;
            if (zself) { for (int k = 0; k < HTB_NBF; ++k) sts_u32(c0base + 128u * k, lds_u32(c0base + 128u * k) + zself); }
            zself = 0;
        }
#pragma unroll
        for (int s = 0; s < HTB_NBF; ++s) {
            const unsigned r = __reduce_add_sync(HTB_FULL, c[s]);
            if (lane == s) tot += (unsigned long long)wt * (unsigned long long)r;
            c[s] = 0;
        }
        if (KIND == 1) {
#pragma unroll 1
            for (int s = 0; s < HTB_NBF; ++s) {
                const unsigned r = __reduce_add_sync(HTB_FULL, lds_u32(c0base + 128u * s));
                if (lane == s) tot0 += (unsigned long long)wt * (unsigned long long)r;
                sts_u32(c0base + 128u * s, 0u);
            }
        }
        return false;
    }
    __device__ __forceinline__ void kernel_end()
    {
        const int k = lane - (HTB_NBF - P.nb);
        if (lane < HTB_NBF && k >= 0 && tot) atomicAdd(P.counts + k, tot);
        if (KIND == 1 && P.counts0 && lane < HTB_NBF && k >= 0 && tot0) atomicAdd(P.counts0 + k, tot0);
    }
};
typedef Fast3T<FAST3_PPL, 0> Fast3;
typedef Fast3T<FAST3_PPL, 1> FastXYZ;

// ------------------------------------------------------------------ generic integer-count variants
// per-warp u32 histogram in shared memory (scratch), flushed to the global u64 histogram per tile
template <int KIND>   // 0: 3-D r   1: (rp, pi)   2: (s, mu) differential
struct GenCount {
    static constexpr int DIM = 3, NPAY = 0, PPL = 2, WARPS = 8, MINBLOCKS = 2;
    static constexpr bool TMA = true;
    typedef GenParams Params;
    const Params &P;
    uint32_t *hist;
    int lane;
    double x0, y0, z0, x1, y1, z1;
    double xs0, ys0, zs0, xs1, ys1, zs1;
    bool v0, v1;

    static size_t scratch_bytes(const Params &p) { return sizeof(uint32_t) * (size_t)((p.nhist + 3) & ~3); }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
        xs0 = x0 - sh[0]; ys0 = y0 - sh[1]; zs0 = z0 - sh[2];
        xs1 = x1 - sh[0]; ys1 = y1 - sh[1]; zs1 = z1 - sh[2];
    }

    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() {}
    __device__ GenCount(const Params &p, void *scratch, int ln, const WalkArrays &) : P(p), hist((uint32_t *)scratch), lane(ln)
    {
        for (int k = lane; k < P.nhist; k += 32) hist[k] = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void tile_begin(const double (&p)[2][3], const bool (&val)[2], const uint32_t (&)[2],
                                               const WalkArrays &)
    {
        x0 = p[0][0]; y0 = p[0][1]; z0 = p[0][2];
        x1 = p[1][0]; y1 = p[1][1]; z1 = p[1][2];
        v0 = val[0]; v1 = val[1];
    }
    __device__ __forceinline__ void pair(bool valid, double xs, double ys, double zs, double xj, double yj, double zj)
    {
        const double dx = xs - xj, dy = ys - yj, dz = zs - zj;
        if (KIND == 0) {
            const double dsq = dx * dx + dy * dy + dz * dz;
            bool alive = valid;
            for (int k = P.n0 - 1; k >= 0; --k) {
                alive = alive && (dsq <= P.e0[k]);
                const unsigned b = __ballot_sync(HTB_FULL, alive);
                if (!b) break;
                if (lane == 0) hist[k] += __popc(b);
            }
        } else if (KIND == 1) {
            const double dxy_sq = dx * dx + dy * dy;
            const double dz_sq = dz * dz;
            bool ak = valid;
            for (int k = P.n0 - 1; k >= 0; --k) {
                ak = ak && (dxy_sq <= P.e0[k]);
                if (!__any_sync(HTB_FULL, ak)) break;
                bool ag = ak;
                for (int g = P.n1 - 1; g >= 0; --g) {
                    ag = ag && (dz_sq <= P.e1[g]);
                    const unsigned b = __ballot_sync(HTB_FULL, ag);
                    if (!b) break;
                    if (lane == 0) hist[k * P.n1 + g] += __popc(b);
                }
            }
        } else {
            const double dxy_sq = dx * dx + dy * dy;
            const double dz_sq = dz * dz;
            const double sqr_s = dz_sq + dxy_sq;
            if (valid && !(sqr_s > P.max0)) {
                double sqr_mu = 0.0;
                if (sqr_s > 0.0) sqr_mu = dxy_sq / sqr_s;
                if (!(sqr_mu > P.max1)) {
                    int k = P.n0 - 2;
                    while (k != -1) { if (sqr_s > P.e0[k]) break; --k; }
                    int g = P.n1 - 2;
                    while (g != -1) { if (sqr_mu > P.e1[g]) break; --g; }
                    atomicAdd(&hist[(k + 1) * P.n1 + (g + 1)], 1u);
                }
            }
        }
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH;
        for (int j = lo; j < hi; ++j) {
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j);
            pair(v0, xs0, ys0, zs0, xj, yj, zj);
            pair(v1, xs1, ys1, zs1, xj, yj, zj);
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&)[2], int, unsigned wt)
    {
        __syncwarp();
        for (int k = lane; k < P.nhist; k += 32) {
            const uint32_t h = hist[k];
            if (h) { atomicAdd(P.counts + k, (unsigned long long)wt * h); hist[k] = 0; }
        }
        __syncwarp();
        return false;
    }
    __device__ __forceinline__ void kernel_end() {}
};

// ------------------------------------------------------------------ Marked3
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HTB_FULL, v, o);
    return v;
}

struct Marked3 {
    static constexpr int DIM = 3, NPAY = HTB_MAX_NW, PPL = 2, WARPS = 8, MINBLOCKS = 2;
    static constexpr bool TMA = true;
    typedef GenParams Params;
    const Params &P;
    double *hist;
    int lane;
    double x0, y0, z0, x1, y1, z1;
    double xs0, ys0, zs0, xs1, ys1, zs1;
    double wa[HTB_MAX_NW], wb[HTB_MAX_NW];
    bool v0, v1;

    static size_t scratch_bytes(const Params &p) { return sizeof(double) * (size_t)((p.nhist + 1) & ~1); }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
        xs0 = x0 - sh[0]; ys0 = y0 - sh[1]; zs0 = z0 - sh[2];
        xs1 = x1 - sh[0]; ys1 = y1 - sh[1]; zs1 = z1 - sh[2];
    }

    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() {}
    __device__ Marked3(const Params &p, void *scratch, int ln, const WalkArrays &) : P(p), hist((double *)scratch), lane(ln)
    {
        for (int k = lane; k < P.nhist; k += 32) hist[k] = 0.0;
        __syncwarp();
    }
    __device__ __forceinline__ void tile_begin(const double (&p)[2][3], const bool (&val)[2], const uint32_t (&idx)[2],
                                               const WalkArrays &A)
    {
        x0 = p[0][0]; y0 = p[0][1]; z0 = p[0][2];
        x1 = p[1][0]; y1 = p[1][1]; z1 = p[1][2];
        v0 = val[0]; v1 = val[1];
        const uint32_t i0 = idx[0], i1 = idx[1];
#pragma unroll
        for (int k = 0; k < HTB_MAX_NW; ++k) {
            wa[k] = (k < A.nw) ? A.pay1[(size_t)i0 * A.nw + k] : 0.0;
            wb[k] = (k < A.nw) ? A.pay1[(size_t)i1 * A.nw + k] : 0.0;
        }
    }
    __device__ __forceinline__ void pair(bool valid, const double *w1, double xs, double ys, double zs,
                                         double xj, double yj, double zj, uint32_t w2)
    {
        const double dx = xs - xj, dy = ys - yj, dz = zs - zj;
        const double dsq = dx * dx + dy * dy + dz * dz;
        bool alive = valid && (dsq <= P.e0[P.n0 - 1]);
        if (!__any_sync(HTB_FULL, alive)) return;
        double w2l[HTB_MAX_NW];
#pragma unroll
        for (int k = 0; k < HTB_MAX_NW; ++k) w2l[k] = (k < P.nw) ? lds_f64(w2 + 8 * k) : 0.0;
        const double w = alive ? htb_pair_weight(P.wfunc, w1, w2l) : 0.0;
        for (int k = P.n0 - 1; k >= 0; --k) {
            alive = alive && (dsq <= P.e0[k]);
            if (!__any_sync(HTB_FULL, alive)) break;
            const double s = warp_sum(alive ? w : 0.0);
            if (lane == 0) hist[k] += s;
        }
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH, bw = stage + 24 * HTB_CH;
        for (int j = lo; j < hi; ++j) {
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j);
            pair(v0, wa, xs0, ys0, zs0, xj, yj, zj, bw + 8 * j * P.nw);
            pair(v1, wb, xs1, ys1, zs1, xj, yj, zj, bw + 8 * j * P.nw);
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&)[2], int, unsigned)
    {
        __syncwarp();
        for (int k = lane; k < P.nhist; k += 32) {
            const double h = hist[k];
            if (h != 0.0) { atomicAdd(P.fcounts + k, h); hist[k] = 0.0; }
        }
        __syncwarp();
        return false;
    }
    __device__ __forceinline__ void kernel_end() {}
};

// ------------------------------------------------------------------ DSigma (2-D, per-object accumulators)
// scratch layout per warp: acc[slot][64], slot 0 = mass inside rp[0]; slots 1..nbin = mass in bin b
// (differential); slots nbin+1..2*nbin = sum m*(1 - ln(rp[b+1]^2 / d^2)) over pairs in bin b.
struct DSigma {
    static constexpr int DIM = 2, NPAY = 1, PPL = 2, WARPS = 4, MINBLOCKS = 2;
    static constexpr bool TMA = true;
    typedef GenParams Params;
    const Params &P;
    double *acc;
    int lane;
    double x0, y0, x1, y1;
    double xs0, ys0, xs1, ys1;
    bool v0, v1;

    static size_t scratch_bytes(const Params &p) { return sizeof(double) * 64 * (size_t)(2 * (p.n0 - 1) + 1); }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
        xs0 = x0 - sh[0]; ys0 = y0 - sh[1]; xs1 = x1 - sh[0]; ys1 = y1 - sh[1];
    }

    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() {}
    __device__ DSigma(const Params &p, void *scratch, int ln, const WalkArrays &) : P(p), acc((double *)scratch), lane(ln) {}
    __device__ __forceinline__ void tile_begin(const double (&p)[2][3], const bool (&val)[2], const uint32_t (&)[2],
                                               const WalkArrays &)
    {
        x0 = p[0][0]; y0 = p[0][1]; x1 = p[1][0]; y1 = p[1][1];
        v0 = val[0]; v1 = val[1];
        const int nslot = 2 * (P.n0 - 1) + 1;
        for (int s = 0; s < nslot; ++s) { acc[s * 64 + lane] = 0.0; acc[s * 64 + 32 + lane] = 0.0; }
        __syncwarp();
    }
    __device__ __forceinline__ void pair(bool valid, int col, double xs, double ys, double xj, double yj, double mj)
    {
        const double dx = xs - xj, dy = ys - yj;
        const double dxy_sq = dx * dx + dy * dy;
        const int nbin = P.n0 - 1;
        if (valid && dxy_sq <= P.e0[nbin]) {
            int k = nbin - 1;
            while (k >= 0 && dxy_sq <= P.e0[k]) --k;      // pair lies in bin k (rp[k] < d <= rp[k+1]); k == -1: inside rp[0]
            acc[(k + 1) * 64 + col] += mj;
            if (k >= 0) acc[(nbin + 1 + k) * 64 + col] += mj * (1 - log(P.e0[k + 1] / dxy_sq));
        }
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bm = stage + 16 * HTB_CH;
        for (int j = lo; j < hi; ++j) {
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), mj = lds_f64(bm + 8 * j);
            pair(v0, lane, xs0, ys0, xj, yj, mj);
            pair(v1, 32 + lane, xs1, ys1, xj, yj, mj);
        }
    }
    __device__ __forceinline__ void finish(bool valid, int col, uint32_t isorted, const WalkArrays &A)
    {
        if (!valid) return;
        const int nbin = P.n0 - 1;
        const int64_t row = P.perm1 ? (int64_t)P.perm1[isorted] : (int64_t)isorted;
        double inside = acc[col];
        for (int k = 0; k < nbin; ++k) {
            // sum_{pairs inside rp[k]} m * 2 * dlog[k]  -  sum_{pairs in bin k} m (1 - ln(rp[k+1]^2/d^2))
            const double ds = inside * 2 * P.e1[k] - acc[(nbin + 1 + k) * 64 + col];
            // (the rows start at zero; slices of one tile add their shares)
            atomicAdd(&P.fcounts[row * nbin + k], ds / (3.14159265358979323846 * (P.e0[k + 1] - P.e0[k])));
            inside += acc[(k + 1) * 64 + col];
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &A, const uint32_t (&idx)[2], int, unsigned)
    {
        __syncwarp();
        finish(v0, lane, idx[0], A);
        finish(v1, 32 + lane, idx[1], A);
        __syncwarp();
        return false;
    }
    __device__ __forceinline__ void kernel_end() {}
};

// ------------------------------------------------------------------ DSigmaU (uniform particle mass)
// With one mass m for every particle the per-pair logarithm disappears:
//   sum_j m (1 - ln(R^2 / d_j^2)) = m [ n (1 - ln R^2) + ln prod_j d_j^2 ]
// and the product is carried as (product of mantissas in [1,2), integer sum of exponents), renormalised
// before it can overflow.  Per in-bin pair: 5 f64 ops for d^2 + 1 DMUL, no division, no log.
// scratch per warp: cnt[slot][64] u32 (slot 0 = inside rp[0], slot b+1 = bin b), es[bin][64] i32,
// pm[bin][64] f64.
struct DSigmaU {
    static constexpr int DIM = 2, NPAY = 0, PPL = 2, WARPS = 4, MINBLOCKS = 3;
    static constexpr bool TMA = true;
    typedef GenParams Params;
    const Params &P;
    double *pm;
    int *es;
    unsigned *cnt;
    int lane, nbin;
    double x0, y0, x1, y1;
    double xs0, ys0, xs1, ys1;
    bool v0, v1;
    int since;
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
        xs0 = x0 - sh[0]; ys0 = y0 - sh[1]; xs1 = x1 - sh[0]; ys1 = y1 - sh[1];
    }

    static size_t scratch_bytes(const Params &p)
    {
        const size_t nbin = (size_t)(p.n0 - 1);
        return 64 * (8 * nbin + 4 * nbin + 4 * (nbin + 1));
    }
    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() {}
    __device__ DSigmaU(const Params &p, void *scratch, int ln, const WalkArrays &) : P(p), lane(ln)
    {
        nbin = P.n0 - 1;
        pm = (double *)scratch;
        es = (int *)(pm + 64 * nbin);
        cnt = (unsigned *)(es + 64 * nbin);
        since = 0;
    }
    __device__ __forceinline__ void tile_begin(const double (&p)[2][3], const bool (&val)[2], const uint32_t (&)[2],
                                               const WalkArrays &)
    {
        x0 = p[0][0]; y0 = p[0][1]; x1 = p[1][0]; y1 = p[1][1];
        v0 = val[0]; v1 = val[1];
        for (int s = 0; s < nbin; ++s) {
            pm[s * 64 + lane] = 1.0; pm[s * 64 + 32 + lane] = 1.0;
            es[s * 64 + lane] = 0; es[s * 64 + 32 + lane] = 0;
        }
        for (int s = 0; s <= nbin; ++s) { cnt[s * 64 + lane] = 0; cnt[s * 64 + 32 + lane] = 0; }
        since = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void pair(bool valid, int col, double xs, double ys, double xj, double yj)
    {
        const double dx = xs - xj, dy = ys - yj;
        const double dxy_sq = dx * dx + dy * dy;
        if (valid && dxy_sq <= P.e0[nbin]) {
            int k = nbin - 1;
            while (k >= 0 && dxy_sq <= P.e0[k]) --k;
            cnt[(k + 1) * 64 + col] += 1u;
            if (k >= 0) {
                const int hi = __double2hiint(dxy_sq);
                es[k * 64 + col] += (hi >> 20) - 1023;
                const double mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(dxy_sq));
                pm[k * 64 + col] *= mant;
            }
        }
    }
    __device__ __forceinline__ void renorm(int col)
    {
        for (int b = 0; b < nbin; ++b) {
            const double v = pm[b * 64 + col];
            const int hi = __double2hiint(v);
            es[b * 64 + col] += (hi >> 20) - 1023;
            pm[b * 64 + col] = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(v));
        }
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH;
        // every multiply grows a product by < 2: renormalise before 2^1023 can be reached
        if (since + (hi - lo) > 900) { renorm(lane); renorm(32 + lane); since = 0; }
        since += hi - lo;
        for (int j = lo; j < hi; ++j) {
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j);
            pair(v0, lane, xs0, ys0, xj, yj);
            pair(v1, 32 + lane, xs1, ys1, xj, yj);
        }
    }
    __device__ __forceinline__ void finish(bool valid, int col, uint32_t isorted)
    {
        if (!valid) return;
        const double m = P.max0;
        const int64_t row = P.perm1 ? (int64_t)P.perm1[isorted] : (int64_t)isorted;
        double inside = (double)cnt[col];
        for (int k = 0; k < nbin; ++k) {
            const double n = (double)cnt[(k + 1) * 64 + col];
            // sum over the bin's pairs of ln(d^2 / rp[k+1]^2), exponents and mantissas kept apart
            const double r2 = P.e0[k + 1];
            const int rhi = __double2hiint(r2);
            const int re = (rhi >> 20) - 1023;
            const double rm = __hiloint2double((rhi & 0x000fffff) | 0x3ff00000, __double2loint(r2));
            const double v = pm[k * 64 + col];
            const double lnratio = 0.6931471805599453 * ((double)es[k * 64 + col] - n * (double)re) + (log(v) - n * log(rm));
            const double t = m * (n + lnratio);
            const double ds = m * inside * 2 * P.e1[k] - t;
            // (the rows start at zero; slices of one tile add their shares)
            atomicAdd(&P.fcounts[row * nbin + k], ds / (3.14159265358979323846 * (P.e0[k + 1] - P.e0[k])));
            inside += n;
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&idx)[2], int, unsigned)
    {
        __syncwarp();
        finish(v0, lane, idx[0]);
        finish(v1, 32 + lane, idx[1]);
        __syncwarp();
        return false;
    }
    __device__ __forceinline__ void kernel_end() {}
};

// ------------------------------------------------------------------ MarkedQ (marked_npairs_3d, weight = w1 * w2)
// marked_npairs_3d_engine.pyx:204-216 with marking function 1 (mweights, marking_functions.pyx:14-24; id 0, the
// custom hook, is the same product).  The distance arithmetic, the 32-bit relative keys, the trackers and the
// compaction cascade are those of Fast3; what differs is that every pair carries a weight:
//   * the weights w2_j of all pairs certainly inside the top edge are summed per point in a register - one predicated
//     DADD per pair, the reference's `counts[k] += weight` for the top bin;
//   * a pair that may lie inside the second edge pushes an 8-byte entry {key, sorted index of j | point slot << 31}
//     (one STS.64; round 1 queued {key, w2_j} as 16 bytes, which cost five instructions per push, 168 registers
//     and half of the queue depth).  The weight is fetched when the entry DROPS OUT of the cascade (once per entry,
//     12 % of the pairs, at full lane occupancy) from the sorted weight array, which the TMA stage has just pulled
//     through L2;
//   * a key that drops out at level S adds w1_i * w2_j to the DIFFERENTIAL sum of level S; the finishing kernel forms
//     the cumulative sums.  Ambiguous keys: group roll-back / exact tile redo as in Fast3.
#ifndef MQ_QCAP
#define MQ_QCAP 48            // queue entries (8 bytes) per lane
#endif
#ifndef MQ_QSURV
#define MQ_QSURV 14           // run the deep cascade when a lane holds more survivors than this
#endif
#ifndef MQ_WARPS
#define MQ_WARPS 4
#define MQ_MINBLOCKS 3
#endif
#define MQ_GROUP 16           // pairs per lane between two queue checks

__device__ __forceinline__ void sts_kj(uint32_t addr, int key, uint32_t j)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(key), "r"(j));
}
__device__ __forceinline__ void lds_kj(uint32_t addr, int &key, uint32_t &j)
{
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(key), "=r"(j) : "r"(addr));
}
// if (key < 0) W += w, as ONE f64 instruction plus two integer ones: W = fma(w, m, W) with m = 1.0 or 0.0 built from the
// sign of the key (w * 1.0 is exact, so the result is the correctly rounded W + w; w * 0.0 leaves W alone for finite w).
// A predicated DADD comes back from ptxas as DADD + two FSEL.
__device__ __forceinline__ void add_if_neg(double &W, int key, double w)
{
    W = __fma_rn(w, __hiloint2double((key >> 31) & 0x3ff00000, 0), W);
}

struct MarkedQ {
    static constexpr int DIM = 3, NPAY = 1, PPL = 2, WARPS = MQ_WARPS, MINBLOCKS = MQ_MINBLOCKS;
    static constexpr bool TMA = true, WANTS_BASE = true, HAS_SELF = true;
    static constexpr int GJ = MQ_GROUP / PPL;
    static constexpr int TOP = HTB_NBF - 1;
    static constexpr uint32_t QFULL = 256u * (MQ_QCAP - MQ_GROUP - PPL);
    typedef Fast3Params Params;
    const Params &P;
    int lane;
    const double *w2g;          // sorted weights of sample2 (global memory)
    uint32_t qbase, qs, qptr, qsave;
    uint32_t jg;                // sorted index of staged slot 0 of the current chunk
    double xs[PPL], ys[PPL], zs[PPL], x[PPL], y[PPL], z[PPL], w1[PPL];
    double Wtop[PPL], Wsave[PPL];   // sum of w2 over the pairs certainly inside the top edge
    double accD[HTB_NBF];           // differential weighted sums per level (this tile), w1 already applied
    double Xall;                    // all weights added outside Wtop (exact path, exact zeros of the own range)
    unsigned umin;
    int hmin;
    bool exact, dirty, always_exact;
    double tot;

    static size_t scratch_bytes(const Params &) { return 8 * 32 * MQ_QCAP; }
    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() { exact = true; }
    __device__ __forceinline__ void chunk_base(uint32_t j0) { jg = j0; }

    __device__ __forceinline__ MarkedQ(const Params &p, void *scratch, int ln, const WalkArrays &A) : P(p), lane(ln)
    {
        w2g = A.pay2;
        jg = 0;
        qbase = smem_u32(scratch) + 8u * (uint32_t)ln;
        qs = qptr = qsave = qbase;
        tot = 0.0; Xall = 0.0; umin = 0xffffffffu; hmin = 0x7fffffff;
        exact = dirty = false;
        always_exact = ((A.flags1[0] | A.flags2[0]) & 1u) != 0u;
#pragma unroll
        for (int s = 0; s < HTB_NBF; ++s) accD[s] = 0.0;
#pragma unroll
        for (int q = 0; q < PPL; ++q) { x[q] = y[q] = z[q] = xs[q] = ys[q] = zs[q] = w1[q] = Wtop[q] = Wsave[q] = 0.0; }
    }
    __device__ __forceinline__ void tile_begin(const double (&p)[PPL][3], const bool (&val)[PPL], const uint32_t (&id)[PPL],
                                               const WalkArrays &A)
    {
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            x[q] = p[q][0]; y[q] = p[q][1]; z[q] = p[q][2];
            w1[q] = val[q] ? A.pay1[id[q]] : 0.0;
            Wtop[q] = Wsave[q] = 0.0;
        }
        exact = always_exact; dirty = false;
    }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
#pragma unroll
        for (int q = 0; q < PPL; ++q) { xs[q] = x[q] - sh[0]; ys[q] = y[q] - sh[1]; zs[q] = z[q] - sh[2]; }
    }
    // w1 of the entry's point times w2 of its sample2 point
    __device__ __forceinline__ double weight_of(uint32_t j) const
    {
        const double wj = __ldg(w2g + (j & 0x7fffffffu));
        return ((j >> 31) ? w1[1] : w1[0]) * wj;
    }
    // Compaction pass over this lane's entries in [r, end): keys <= Fk are kept (moved down to w), the weights of the
    // others are summed (they drop out at this level); a key equal to Fin (the edge that admitted it) = dirty.
    // Four independent entries per trip hide the shared- and global-memory latencies.
    __device__ __forceinline__ uint32_t compact(uint32_t r, uint32_t end, uint32_t w, int Fk, int Fin, double &dropped)
    {
        bool eq = false;
        double s0 = 0.0, s1 = 0.0;
        for (; r + 1024u <= end; r += 1024u) {
            int k0, k1, k2, k3;
            uint32_t j0, j1, j2, j3;
            lds_kj(r, k0, j0); lds_kj(r + 256u, k1, j1); lds_kj(r + 512u, k2, j2); lds_kj(r + 768u, k3, j3);
            eq |= (k0 == Fin) | (k1 == Fin) | (k2 == Fin) | (k3 == Fin);
            const double a0 = k0 <= Fk ? 0.0 : weight_of(j0), a1 = k1 <= Fk ? 0.0 : weight_of(j1);
            const double a2 = k2 <= Fk ? 0.0 : weight_of(j2), a3 = k3 <= Fk ? 0.0 : weight_of(j3);
            if (k0 <= Fk) { sts_kj(w, k0, j0); w += 256u; }
            if (k1 <= Fk) { sts_kj(w, k1, j1); w += 256u; }
            if (k2 <= Fk) { sts_kj(w, k2, j2); w += 256u; }
            if (k3 <= Fk) { sts_kj(w, k3, j3); w += 256u; }
            s0 += a0 + a1;
            s1 += a2 + a3;
        }
        for (; r != end; r += 256u) {
            int k0;
            uint32_t j0;
            lds_kj(r, k0, j0);
            eq |= (k0 == Fin);
            if (k0 <= Fk) { sts_kj(w, k0, j0); w += 256u; }
            else s0 += weight_of(j0);
        }
        dirty |= eq;
        dropped = s0 + s1;
        return w;
    }
    template <int S>
    __device__ __forceinline__ void cascade(uint32_t end)
    {
        if (!__any_sync(HTB_FULL, end != qbase)) return;
        if constexpr (S >= 1) {
            // keys in [qbase, end) were admitted by F[S]; those above F[S - 1] have level S
            double d;
            const uint32_t w = compact(qbase, end, qbase, P.F[S - 1], P.F[S], d);
            accD[S] += d;
            cascade<S - 1>(w);
        } else {
            // level 0: everything left has level 0
            double d;
            (void)compact(qbase, end, qbase, (int)0x80000000, P.F[0], d);
            accD[0] += d;
        }
    }
    __device__ __forceinline__ void deep()
    {
        cascade<TOP - 2>(qs);
        qs = qptr = qsave = qbase;
    }
    // the entries pushed since the last call were admitted by F[TOP - 1]: those above F[TOP - 2] have level TOP - 1;
    // the survivors stay queued
    __device__ __forceinline__ void flush1(bool force_deep)
    {
        double d;
        const uint32_t w = compact(qs, qptr, qs, P.F[TOP - 2], P.F[TOP - 1], d);
        accD[TOP - 1] += d;
        qs = qptr = qsave = w;
        if (force_deep || __any_sync(HTB_FULL, w > qbase + 256u * MQ_QSURV)) deep();
    }
    __device__ __forceinline__ void key_of(int q, double xj, double yj, double zj, int &key, int &hi)
    {
        const double dx = xs[q] - xj, dy = ys[q] - yj, dz = zs[q] - zj;
        const double dsq = dx * dx + dy * dy + dz * dz;
        hi = __double2hiint(dsq);
        key = (int)(__funnelshift_l((unsigned)__double2loint(dsq), (unsigned)hi, 6) + (unsigned)P.nbias);
    }
    __device__ __forceinline__ void push(int q, int key, uint32_t jtag, double wj)
    {
        add_if_neg(Wtop[q], key, wj);
        if (key <= P.F[TOP - 1]) { sts_kj(qptr, key, jtag); qptr += 256u; }
    }
    // one point of this lane against two staged points (three-input minimum trackers: one VIMNMX3 each per two pairs)
    __device__ __forceinline__ void pair2(int q, uint32_t tag, uint32_t ja, double xa, double ya, double za, double wa,
                                          double xb, double yb, double zb, double wb)
    {
        int ka, kb, ha, hb;
        key_of(q, xa, ya, za, ka, ha);
        key_of(q, xb, yb, zb, kb, hb);
        umin = min(umin, min((unsigned)ka, (unsigned)kb));
        hmin = min(hmin, min(ha, hb));
        push(q, ka, ja | tag, wa);
        push(q, kb, (ja + 1u) | tag, wb);
    }
    __device__ __forceinline__ void pair_fast(int q, uint32_t jtag, double xj, double yj, double zj, double wj)
    {
        const double dx = xs[q] - xj, dy = ys[q] - yj, dz = zs[q] - zj;
        const double dsq = dx * dx + dy * dy + dz * dz;
        const int hi = __double2hiint(dsq);
        const int key = (int)(__funnelshift_l((unsigned)__double2loint(dsq), (unsigned)hi, 6) + (unsigned)P.nbias);
        umin = min(umin, (unsigned)key);
        hmin = min(hmin, hi);
        add_if_neg(Wtop[q], key, wj);
        if (key <= P.F[TOP - 1]) { sts_kj(qptr, key, jtag); qptr += 256u; }
    }
    // Five levels of the exact scan: if (dsq <= e_s && no lower level took the pair) a_s += w.  Per level one DSETP (the
    // f64 compare orders non-negative values exactly like the 64-bit integer compare of their bit patterns, and runs on
    // the FP64 pipe instead of two ISETP on the half-rate ALU pipe), the predicate as a 1.0 / 0.0 multiplier (one SEL), one
    // DFMA (w * 1.0 is exact, w * 0.0 leaves the sum alone) and one predicate OR; the C version - and a predicated
    // add.f64 as well - comes back from ptxas as an unconditional DADD and two FSEL, with two ISETP for the compare.
    static __device__ __forceinline__ void levels5(double &a0, double &a1, double &a2, double &a3, double &a4, double e0, double e1,
                                                   double e2, double e3, double e4, double d, double w, unsigned &below)
    {
        asm("{\n\t.reg .pred ad, bel;\n\t.reg .b32 h, z;\n\t.reg .f64 m;\n\t"
            "mov.b32 z, 0;\n\tsetp.ne.u32 bel, %5, 0;\n\t"
            "setp.le.and.f64 ad, %11, %6, !bel;\n\tselp.b32 h, 0x3ff00000, 0, ad;\n\tmov.b64 m, {z, h};\n\tfma.rn.f64 %0, %12, m, %0;\n\tor.pred bel, bel, ad;\n\t"
            "setp.le.and.f64 ad, %11, %7, !bel;\n\tselp.b32 h, 0x3ff00000, 0, ad;\n\tmov.b64 m, {z, h};\n\tfma.rn.f64 %1, %12, m, %1;\n\tor.pred bel, bel, ad;\n\t"
            "setp.le.and.f64 ad, %11, %8, !bel;\n\tselp.b32 h, 0x3ff00000, 0, ad;\n\tmov.b64 m, {z, h};\n\tfma.rn.f64 %2, %12, m, %2;\n\tor.pred bel, bel, ad;\n\t"
            "setp.le.and.f64 ad, %11, %9, !bel;\n\tselp.b32 h, 0x3ff00000, 0, ad;\n\tmov.b64 m, {z, h};\n\tfma.rn.f64 %3, %12, m, %3;\n\tor.pred bel, bel, ad;\n\t"
            "setp.le.and.f64 ad, %11, %10, !bel;\n\tselp.b32 h, 0x3ff00000, 0, ad;\n\tmov.b64 m, {z, h};\n\tfma.rn.f64 %4, %12, m, %4;\n\tor.pred bel, bel, ad;\n\t"
            "selp.u32 %5, 1, 0, bel;\n\t}"
            : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+r"(below)
            : "d"(e0), "d"(e1), "d"(e2), "d"(e3), "d"(e4), "d"(d), "d"(w));
    }
    __device__ __forceinline__ double edge(int s) const { return __longlong_as_double((long long)(s == HTB_NBF - 1 ? P.E_top : P.E[s])); }
    __device__ __forceinline__ void pair_exact(int q, double xj, double yj, double zj, double wj)
    {
        const double dx = xs[q] - xj, dy = ys[q] - yj, dz = zs[q] - zj;
        const double dsq = dx * dx + dy * dy + dz * dz;
        if (dsq <= edge(HTB_NBF - 1)) {
            const double w = w1[q] * wj;
            Xall += w;
            // the weight goes to the lowest level whose edge holds the pair (the edges ascend)
            unsigned below = 0;
            static_assert(HTB_NBF - 1 == 15, "three groups of five levels");
            levels5(accD[0], accD[1], accD[2], accD[3], accD[4], edge(0), edge(1), edge(2), edge(3), edge(4), dsq, w, below);
            levels5(accD[5], accD[6], accD[7], accD[8], accD[9], edge(5), edge(6), edge(7), edge(8), edge(9), dsq, w, below);
            levels5(accD[10], accD[11], accD[12], accD[13], accD[14], edge(10), edge(11), edge(12), edge(13), edge(14), dsq, w, below);
        }
    }
    __device__ __forceinline__ void exact_range(uint32_t stage, int j0, int j1)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH, bw = stage + 24 * HTB_CH;
#pragma unroll 1
        for (int j = j0; j < j1; ++j) {
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j), wj = lds_f64(bw + 8 * j);
#pragma unroll
            for (int q = 0; q < PPL; ++q) pair_exact(q, xj, yj, zj, wj);
        }
    }
    __device__ __forceinline__ void check(uint32_t stage, int j0, int j1)
    {
        const bool undecided = (umin == 0u) | (hmin < P.Hwin);
        const bool full = qptr > qbase + QFULL;
        if (__any_sync(HTB_FULL, undecided | full)) {
            if (__any_sync(HTB_FULL, undecided)) {
                qptr = qsave; Wtop[0] = Wsave[0]; Wtop[1] = Wsave[1];
                exact_range(stage, j0, j1);
                umin = 0xffffffffu; hmin = 0x7fffffff;
            }
            if (__any_sync(HTB_FULL, qptr > qbase + QFULL)) flush1(false);
        }
        qsave = qptr; Wsave[0] = Wtop[0]; Wsave[1] = Wtop[1];
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t tok, bool own)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH, bw = stage + 24 * HTB_CH;
        if (exact | own) { exact_range(stage, lo, hi); return; }    // own index range: see Fast3T::chunk
        int j = lo;
        if ((j & 1) && j < hi) {
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j), wj = lds_f64(bw + 8 * j);
            pair_fast(0, jg + (uint32_t)j, xj, yj, zj, wj);
            pair_fast(1, (jg + (uint32_t)j) | 0x80000000u, xj, yj, zj, wj);
            ++j;
        }
        int j0 = lo;
        if (j + GJ <= hi) {
            double xa, xb, ya, yb, za, zb, wa, wb;
            lds_f64x2_tok(bx + 8 * j, tok, xa, xb);
            lds_f64x2_tok(by + 8 * j, tok, ya, yb);
            lds_f64x2_tok(bz + 8 * j, tok, za, zb);
            lds_f64x2_tok(bw + 8 * j, tok, wa, wb);
#pragma unroll 1
            for (; j + GJ <= hi; j += GJ) {
#pragma unroll
                for (int u = 0; u < GJ; u += 2) {
                    double xc, xd, yc, yd, zc, zd, wc, wd;
                    lds_f64x2_tok(bx + 8 * (j + u + 2), tok, xc, xd);     // may run past hi: harmless, never used
                    lds_f64x2_tok(by + 8 * (j + u + 2), tok, yc, yd);
                    lds_f64x2_tok(bz + 8 * (j + u + 2), tok, zc, zd);
                    lds_f64x2_tok(bw + 8 * (j + u + 2), tok, wc, wd);
                    const uint32_t ja = jg + (uint32_t)(j + u);
                    pair2(0, 0u, ja, xa, ya, za, wa, xb, yb, zb, wb);
                    pair2(1, 0x80000000u, ja, xa, ya, za, wa, xb, yb, zb, wb);
                    xa = xc; xb = xd; ya = yc; yb = yd; za = zc; zb = zd; wa = wc; wb = wd;
                }
                check(stage, j0, j + GJ);
                j0 = j + GJ;
            }
        }
        if (j0 < hi) {
#pragma unroll 1
            for (; j < hi; ++j) {
                const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = lds_f64(bz + 8 * j), wj = lds_f64(bw + 8 * j);
                pair_fast(0, jg + (uint32_t)j, xj, yj, zj, wj);
                pair_fast(1, (jg + (uint32_t)j) | 0x80000000u, xj, yj, zj, wj);
            }
            check(stage, j0, hi);
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&)[PPL], int pass, unsigned wt)
    {
        if (!exact) {
            flush1(true);
            if (__any_sync(HTB_FULL, dirty) && pass == 0) {
#pragma unroll
                for (int s = 0; s < HTB_NBF; ++s) accD[s] = 0.0;
                Wtop[0] = Wtop[1] = Wsave[0] = Wsave[1] = 0.0; Xall = 0.0;
                dirty = false; exact = true;
                return true;
            }
        }
        const double all_in = warp_sum(w1[0] * Wtop[0] + w1[1] * Wtop[1] + Xall);
        if (lane == HTB_NBF) tot += (double)wt * all_in;
#pragma unroll
        for (int s = 0; s < HTB_NBF; ++s) {
            const double r = warp_sum(accD[s]);
            if (lane == s) tot += (double)wt * r;
            accD[s] = 0.0;
        }
        Wtop[0] = Wtop[1] = Wsave[0] = Wsave[1] = 0.0; Xall = 0.0;
        return false;
    }
    __device__ __forceinline__ void kernel_end()
    {
        if (lane <= HTB_NBF && tot != 0.0) atomicAdd(P.fsums + lane, tot);
    }
};

// ------------------------------------------------------------------ DSigmaQ (uniform particle mass, fast path)
// mean_delta_sigma_engine.pyx:162-180 for one particle mass m.  Per galaxy and annulus the engine needs the number of
// particles n and sum ln d^2 (see DSigmaU); here every lane owns ONE galaxy and
//   * a pair certainly inside the TOP annulus (F1 < key < 0, the same 32-bit relative keys as Fast3) is folded
//     immediately into register accumulators: mantissa product (one predicated DMUL), exponent sum, count;
//   * every other pair that may be in range (key <= F1) pushes its full 64-bit d^2 to the lane's queue; the queue
//     is drained at full lane occupancy by a top-down compaction cascade with exact 64-bit compares, each pass
//     folding the keys of one annulus into that annulus' accumulators (mantissa products in registers, exponent
//     sums and counts in shared memory);
//   * a key equal to the top-edge key, or a separation too small for the 32-bit key, rolls the group back and
//     pushes the group's in-range pairs after exact compares (the cascade decides everything exactly).
// Integer decisions are therefore exact; the float sums differ from the reference by summation order only.
#ifndef DSQ_QCAP
#define DSQ_QCAP 44           // queue slots (64-bit) per lane
#endif
#ifndef DSQ_QSURV
#define DSQ_QSURV 10          // run the deep cascade when a lane holds more survivors than this
#endif
#ifndef DSQ_WARPS
#define DSQ_WARPS 4
#define DSQ_MINBLOCKS 3
#endif
#ifndef DSQ_GROUP
#define DSQ_GROUP 16          // pairs per lane between two queue checks
#endif

__device__ __forceinline__ void sts_u64_nc(uint32_t addr, unsigned long long v)
{
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v));
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

struct DSigmaQ {
    static constexpr int DIM = 2, NPAY = 0, PPL = 1, WARPS = DSQ_WARPS, MINBLOCKS = DSQ_MINBLOCKS;
    static constexpr bool TMA = true;
    static constexpr int TOP = HTB_NBF - 1;
    static constexpr uint32_t QFULL = 256u * (DSQ_QCAP - DSQ_GROUP - 1);
    typedef DSQParams Params;
    const Params &P;
    int lane, pad;
    uint32_t qbase, qs, qptr, qsave;
    int *ex;                    // [HTB_NBF][32] exponent sums (unbiased) per annulus slot
    unsigned *nn;               // [HTB_NBF][32] counts per slot (slot pad = inside rp[0])
    double M[HTB_NBF];          // mantissa products per slot, kept in [1, 2)
    double x, y, xs, ys;
    double Mtop, Msave;         // top annulus: running mantissa product since the last fold
    int Etop, Esave;            // ... sum of raw exponent fields
    unsigned Ntop, Nsave;       // ... count
    unsigned umin;
    int hmin;
    int groups;
    bool valid, always_exact;

    static size_t scratch_bytes(const Params &) { return 8 * 32 * DSQ_QCAP + 2 * 4 * 32 * HTB_NBF; }

    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() {}
    __device__ __forceinline__ DSigmaQ(const Params &p, void *scratch, int ln, const WalkArrays &A) : P(p), lane(ln)
    {
        pad = HTB_NBF - P.nrp;
        qbase = smem_u32(scratch) + 8u * (uint32_t)ln;
        qs = qptr = qsave = qbase;
        ex = (int *)((unsigned char *)scratch + 8 * 32 * DSQ_QCAP) + ln;
        nn = (unsigned *)((unsigned char *)scratch + 8 * 32 * DSQ_QCAP + 4 * 32 * HTB_NBF) + ln;
        x = y = xs = ys = 0.0;
        Mtop = Msave = 1.0; Etop = Esave = 0; Ntop = Nsave = 0;
        umin = 0xffffffffu; hmin = 0x7fffffff; groups = 0; valid = false;
        always_exact = ((A.flags1[0] | A.flags2[0]) & 1u) != 0u;
#pragma unroll
        for (int s = 0; s < HTB_NBF; ++s) M[s] = 1.0;
    }
    __device__ __forceinline__ void tile_begin(const double (&p)[1][3], const bool (&val)[1], const uint32_t (&)[1],
                                               const WalkArrays &)
    {
        x = p[0][0]; y = p[0][1]; valid = val[0];
#pragma unroll
        for (int s = 0; s < HTB_NBF; ++s) { M[s] = 1.0; ex[32 * s] = 0; nn[32 * s] = 0u; }
        Mtop = Msave = 1.0; Etop = Esave = 0; Ntop = Nsave = 0;
        groups = 0;
    }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
        xs = x - sh[0]; ys = y - sh[1];
    }
    static __device__ __forceinline__ double mant_of(unsigned long long b)
    {
        return __longlong_as_double((long long)((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL));
    }
    // fold the running top-annulus product into slot TOP (keeps every product in [1, 2): no overflow)
    __device__ __forceinline__ void fold_top()
    {
        const double v = M[TOP] * Mtop;
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        ex[32 * TOP] += Etop - 1023 * (int)Ntop + (int)(b >> 52) - 1023;
        nn[32 * TOP] += Ntop;
        M[TOP] = mant_of(b);
        Mtop = Msave = 1.0; Etop = Esave = 0; Ntop = Nsave = 0;
    }
    // One cascade pass per annulus slot S: the keys in [begin, end) satisfy d^2 <= E[S].  Those above E[S - 1] belong
    // to slot S and are folded into its accumulators; the others are compacted down for the next pass.
    template <int S>
    __device__ __forceinline__ void cascade(uint32_t begin, uint32_t end)
    {
        if (!__any_sync(HTB_FULL, end != begin)) return;
        if (S <= pad) {
            // everything left lies inside the lowest edge: count only
            nn[32 * (S < 0 ? 0 : S)] += (end - begin) >> 8;
            return;
        }
        if constexpr (S > 0) {
            const unsigned long long lower = P.E[S - 1];
            double m = 1.0;
            int e = 0;
            uint32_t w = begin;
            for (uint32_t r = begin; r != end; r += 256u) {
                const unsigned long long b = lds_u64(r);
                if (b > lower) { m *= mant_of(b); e += (int)(b >> 52); }
                else { sts_u64_nc(w, b); w += 256u; }
            }
            const unsigned folded = (end - w) >> 8;
            const double v = M[S] * m;
            const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
            ex[32 * S] += e - 1023 * (int)folded + (int)(vb >> 52) - 1023;
            nn[32 * S] += folded;
            M[S] = mant_of(vb);
            cascade<S - 1>(begin, w);
        }
    }
    __device__ __forceinline__ void deep()
    {
        cascade<TOP - 2>(qbase, qs);
        qs = qptr = qsave = qbase;
    }
    // Drain the keys pushed since the last call through the passes of the two outermost annuli (the top one only
    // receives the rare keys the 32-bit test could not place); the rest stays queued for deep().
    __device__ __forceinline__ void flush1(bool force_deep)
    {
        const unsigned long long lower = P.E[TOP - 1], lower2 = P.E[TOP - 2];
        const bool two = pad < TOP - 1;             // slot TOP - 1 is an annulus (not the inside bucket)
        double m1 = 1.0;
        int e1 = 0;
        unsigned n1 = 0;
        uint32_t w = qs;
        for (uint32_t r = qs; r != qptr; r += 256u) {
            const unsigned long long b = lds_u64(r);
            if (b > lower) { Mtop *= mant_of(b); Etop += (int)(b >> 52); Ntop += 1u; }
            else if (two && b > lower2) { m1 *= mant_of(b); e1 += (int)(b >> 52); n1 += 1u; }
            else { sts_u64_nc(w, b); w += 256u; }
        }
        fold_top();
        if (two) {
            const double v = M[TOP - 1] * m1;
            const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
            ex[32 * (TOP - 1)] += e1 - 1023 * (int)n1 + (int)(vb >> 52) - 1023;
            nn[32 * (TOP - 1)] += n1;
            M[TOP - 1] = mant_of(vb);
        }
        qs = qptr = qsave = w;
        if (force_deep || __any_sync(HTB_FULL, w > qbase + 256u * DSQ_QSURV)) {
            if (two) deep();
            else { nn[32 * (TOP - 1)] += (qs - qbase) >> 8; qs = qptr = qsave = qbase; }   // one annulus: the rest is inside rp[0]
        }
    }
    __device__ __forceinline__ void pair2(double xa, double ya, double xb, double yb)
    {
        const double dxa = xs - xa, dya = ys - ya, dxb = xs - xb, dyb = ys - yb;
        const double da = dxa * dxa + dya * dya, db = dxb * dxb + dyb * dyb;
        const int ha = __double2hiint(da), hb = __double2hiint(db);
        const int ka = (int)(__funnelshift_l((unsigned)__double2loint(da), (unsigned)ha, 6) + (unsigned)P.nbias);
        const int kb = (int)(__funnelshift_l((unsigned)__double2loint(db), (unsigned)hb, 6) + (unsigned)P.nbias);
        umin = min(umin, min((unsigned)ka, (unsigned)kb));
        hmin = min(hmin, min(ha, hb));
        one(da, ha, ka);
        one(db, hb, kb);
    }
    __device__ __forceinline__ void one(double d, int h, int k)
    {
        if ((unsigned)(k - P.F1 - 1) < P.Tspan) {
            // certainly inside the top annulus
            Mtop *= __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(d));
            Etop += (int)((unsigned)h >> 20);
            Ntop += 1u;
        }
        if (k <= P.F1) { sts_u64_nc(qptr, (unsigned long long)__double_as_longlong(d)); qptr += 256u; }
    }
    __device__ __forceinline__ void pair_fast(double xj, double yj)
    {
        const double dx = xs - xj, dy = ys - yj;
        const double d = dx * dx + dy * dy;
        const int h = __double2hiint(d);
        const int k = (int)(__funnelshift_l((unsigned)__double2loint(d), (unsigned)h, 6) + (unsigned)P.nbias);
        umin = min(umin, (unsigned)k);
        hmin = min(hmin, h);
        one(d, h, k);
    }
    // exact evaluation of staged entries [j0, j1): every pair inside the top edge goes to the queue
    __device__ __forceinline__ void exact_range(uint32_t stage, int j0, int j1)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH;
#pragma unroll 1
        for (int j = j0; j < j1; ++j) {
            const double dx = xs - lds_f64(bx + 8 * j), dy = ys - lds_f64(by + 8 * j);
            const double d = dx * dx + dy * dy;
            const unsigned long long b = (unsigned long long)__double_as_longlong(d);
            if (b <= P.E[TOP]) { sts_u64_nc(qptr, b); qptr += 256u; }
            if (__any_sync(HTB_FULL, qptr > qbase + 256u * (DSQ_QCAP - 2))) flush1(false);
        }
    }
    __device__ __forceinline__ void check(uint32_t stage, int j0, int j1)
    {
        const bool undecided = (umin == 0u) | (hmin < P.Hwin);
        const bool full = qptr > qbase + QFULL;
        ++groups;
        if (__any_sync(HTB_FULL, undecided | full | ((groups & 31) == 0))) {
            if (__any_sync(HTB_FULL, undecided)) {
                qptr = qsave; Mtop = Msave; Etop = Esave; Ntop = Nsave;
                exact_range(stage, j0, j1);
                umin = 0xffffffffu; hmin = 0x7fffffff;
            }
            if (__any_sync(HTB_FULL, qptr > qbase + QFULL)) flush1(false);
            else if ((groups & 31) == 0) fold_top();      // bounds the running product (< 2^512) and exponent sum
        }
        qsave = qptr; Msave = Mtop; Esave = Etop; Nsave = Ntop;
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t tok)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH;
        if (always_exact) { exact_range(stage, lo, hi); qsave = qptr; return; }
        int j = lo;
        if ((j & 1) && j < hi) { pair_fast(lds_f64(bx + 8 * j), lds_f64(by + 8 * j)); ++j; }
        int j0 = lo;
        if (j + DSQ_GROUP <= hi) {
            double xa, xb, ya, yb;
            lds_f64x2_tok(bx + 8 * j, tok, xa, xb);
            lds_f64x2_tok(by + 8 * j, tok, ya, yb);
#pragma unroll 1
            for (; j + DSQ_GROUP <= hi; j += DSQ_GROUP) {
#pragma unroll
                for (int u = 0; u < DSQ_GROUP; u += 2) {
                    double xc, xd, yc, yd;
                    lds_f64x2_tok(bx + 8 * (j + u + 2), tok, xc, xd);     // may run past hi: harmless, never used
                    lds_f64x2_tok(by + 8 * (j + u + 2), tok, yc, yd);
                    pair2(xa, ya, xb, yb);
                    xa = xc; xb = xd; ya = yc; yb = yd;
                }
                check(stage, j0, j + DSQ_GROUP);
                j0 = j + DSQ_GROUP;
            }
        }
        if (j0 < hi) {
#pragma unroll 1
            for (; j < hi; ++j) pair_fast(lds_f64(bx + 8 * j), lds_f64(by + 8 * j));
            check(stage, j0, hi);
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&idx)[1], int, unsigned)
    {
        flush1(true);
        if (valid) {
            const int nbin = P.nrp - 1;
            const double m = P.mass;
            const int64_t row = P.perm1 ? (int64_t)P.perm1[idx[0]] : (int64_t)idx[0];
            double inside = (double)nn[32 * pad];
#pragma unroll
            for (int s = 1; s < HTB_NBF; ++s) {
                const int k = s - 1 - pad;            // annulus between edges k and k + 1
                if (k < 0) continue;
                const double n = (double)nn[32 * s];
                // sum over the annulus of ln(d^2 / rp[k+1]^2): exponents and mantissas kept apart
                const double r2 = P.e0[k + 1];
                const int rhi = __double2hiint(r2);
                const int re = (rhi >> 20) - 1023;
                const double rm = __hiloint2double((rhi & 0x000fffff) | 0x3ff00000, __double2loint(r2));
                const double lnratio = 0.6931471805599453 * ((double)ex[32 * s] - n * (double)re) + (log(M[s]) - n * log(rm));
                const double t = m * (n + lnratio);
                const double ds = m * inside * 2 * P.e1[k] - t;
                atomicAdd(&P.out[row * nbin + k], ds / (3.14159265358979323846 * (P.e0[k + 1] - P.e0[k])));
                inside += n;
            }
        }
        return false;
    }
    __device__ __forceinline__ void kernel_end() {}
};

// ------------------------------------------------------------------ DSigmaR (uniform particle mass, cell-resolved)
// mean_delta_sigma_engine.pyx:162-180 for one particle mass.  As in DSigmaQ every lane owns one galaxy and needs,
// per annulus, the number of particles and sum ln d^2 (carried as a product).  Most particles of config-5-like
// inputs lie in range (65 % of the evaluated pairs) and the annuli are wide compared with a fine cell of the
// particle mesh, so the annulus of a pair is decided PER (galaxy, CELL) instead of per pair: from the galaxy's
// distance range to the cell's box the lane knows the slots (annuli) its pairs with that cell can fall into;
//   * at most NE = 1, 2 or 3 edges inside the range (the warp takes the largest NE of its lanes): per pair 5 f64
//     ops for d^2, one unconditional DMUL into the cell product and, per edge, one exact 64-bit integer compare, a
//     select, a DMUL into the CUMULATIVE product of the pairs below that edge and a predicated count - no queue, no
//     key, no cascade.  At the end of the cell the lane banks the cumulative products: slot s + i gets
//     C_i as numerator and C_(i-1) as denominator, and the product of an annulus is recovered at the end of the
//     tile as numerator / denominator (in logarithms);
//   * more edges inside some lane's range, or a pair that may be (nearly) coincident: the whole warp takes the
//     exact per-pair scan for that cell (slot search bounded by the lane's range).
// Slots: s = number of edges e with e < d^2 (0: inside rp[0], 1..nrp-1: annulus s-1, nrp: outside); every decision
// is the reference's exact `dxy_sq <= rp^2` on the same f64 value, so only the summation order differs.
#ifndef DSR_WARPS
#define DSR_WARPS 4
#define DSR_MINBLOCKS 4
#endif
#ifndef DSR_CH
#define DSR_CH 128          // particles per staged chunk
#endif
struct DSigmaR {
    static constexpr int DIM = 2, NPAY = 0, PPL = 1, WARPS = DSR_WARPS, MINBLOCKS = DSR_MINBLOCKS;
    static constexpr bool TMA = true, PER_CELL = true;
    static constexpr int CH = DSR_CH;
    static constexpr int NS = HTB_NBF + 4;
    typedef DSRParams Params;
    const Params &P;
    int lane;
    uint32_t *extra;
    // per-lane accumulators per slot, indexed with the lane's own slot numbers: they live in LOCAL memory (L1
    // resident, touched only at cell boundaries and by the exact path), which keeps shared memory for the staging
    // ring and lets 12-16 warps per SM hide the f64 latencies
    double Num[NS], Den[NS];    // mantissas in [1, 2) of the numerator / denominator products
    int eNum[NS], eDen[NS];     // ... their binary exponents
    unsigned cnt[NS];           // pair counts
    uint32_t es;                // shared-space address of the warp's copy of the squared edges (+inf padded)
    double x, y, xs, ys;
    bool valid, always_exact;
    int mode;                   // warp-uniform: 1..3 = edges per lane handled in registers, 0 = exact scan
    int uex;                    // warp-uniform: edges the exact scan has to test per pair (the widest lane range)
    int slo, shi;
    unsigned long long eb0, eb1, eb2;
    double Pall, C0, C1, C2;
    int Eall, E0, E1, E2;
    unsigned n0, n1, n2, npart;
    int since;                  // factors multiplied into the running products since the last renormalisation

    static size_t scratch_bytes(const Params &) { return 8 * HTB_SPAN_CAP + 8 * (2 * HTB_NBF + 4); }
    __device__ __forceinline__ uint32_t *span_extra() { return extra; }
    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() {}
    __device__ __forceinline__ DSigmaR(const Params &p, void *scratch, int ln, const WalkArrays &A) : P(p), lane(ln)
    {
        unsigned char *b = (unsigned char *)scratch;
        extra = (uint32_t *)b; b += 8 * HTB_SPAN_CAP;
        double *ed = (double *)b; b += 8 * (2 * HTB_NBF + 4);
        es = smem_u32(ed);
        for (int k = ln; k < 2 * HTB_NBF + 4; k += 32) ed[k] = k < HTB_NBF ? P.Ed[k] : __longlong_as_double(0x7ff0000000000000LL);
        x = y = xs = ys = 0.0; valid = false; mode = 1; uex = 0; slo = shi = 0; eb0 = eb1 = eb2 = 0;
        Pall = C0 = C1 = C2 = 1.0; Eall = E0 = E1 = E2 = 0; n0 = n1 = n2 = npart = 0; since = 0;
        always_exact = ((A.flags1[0] | A.flags2[0]) & 1u) != 0u;
        __syncwarp();
    }
    __device__ __forceinline__ void tile_begin(const double (&p)[1][3], const bool (&val)[1], const uint32_t (&)[1],
                                               const WalkArrays &)
    {
        // unused lanes of a partial tile shadow lane 0 (always valid): nothing they meet is out of the ordinary
        const double x0 = __shfl_sync(HTB_FULL, p[0][0], 0), y0 = __shfl_sync(HTB_FULL, p[0][1], 0);
        valid = val[0];
        x = valid ? p[0][0] : x0;
        y = valid ? p[0][1] : y0;
#pragma unroll
        for (int s = 0; s < NS; ++s) { Num[s] = 1.0; Den[s] = 1.0; eNum[s] = 0; eDen[s] = 0; cnt[s] = 0u; }
    }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
        xs = x - sh[0]; ys = y - sh[1];
    }
    static __device__ __forceinline__ void renorm(double &m, int &e)
    {
        const int h = __double2hiint(m);
        e += (h >> 20) - 1023;
        m = __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(m));
    }
    __device__ __forceinline__ void cell_begin(double bx0, double bx1, double by0, double by1)
    {
        npart = 0;
        if (always_exact) { slo = 0; shi = P.nrp; mode = 0; uex = P.nrp; return; }
        const double gx = fmax(0.0, fmax(bx0 - xs, xs - bx1)), gy = fmax(0.0, fmax(by0 - ys, ys - by1));
        const double fx = fmax(xs - bx0, bx1 - xs), fy = fmax(ys - by0, by1 - ys);
        const double dmin2 = (gx * gx + gy * gy) * (1.0 - 1e-12), dmax2 = (fx * fx + fy * fy) * (1.0 + 1e-12);
        int a = 0, b = 0;
#pragma unroll
        for (int k = 0; k < HTB_NBF; ++k) {
            const double e = P.Ed[k];
            a += (e < dmin2) ? 1 : 0;
            b += (e < dmax2) ? 1 : 0;
        }
        slo = a; shi = b;
        const int u = (dmin2 < P.tiny2) ? 99 : b - a;           // edges inside this lane's range
        const int umax = __reduce_max_sync(HTB_FULL, u);
        mode = umax <= 1 ? 1 : (umax <= 3 ? umax : 0);
        if (mode == 0) { uex = __reduce_max_sync(HTB_FULL, b - a); return; }
        // this lane's edges: pairs with bits(d^2) <= eb_i are inside edge slo + i; edges at or beyond the range's
        // upper end hold every pair of the cell (+inf: the compare is always true)
        const unsigned long long inf = 0x7ff0000000000000ULL;
        eb0 = u >= 1 ? lds_u64(es + 8u * (uint32_t)a) : inf;
        eb1 = u >= 2 ? lds_u64(es + 8u * (uint32_t)(a + 1)) : inf;
        eb2 = u >= 3 ? lds_u64(es + 8u * (uint32_t)(a + 2)) : inf;
        Pall = C0 = C1 = C2 = 1.0; Eall = E0 = E1 = E2 = 0; n0 = n1 = n2 = 0; since = 0;
    }
    // if (bits(d) <= edge) { C *= d; n += 1; } as one 64-bit compare, a predicated DMUL and a predicated IADD
    // (left to the compiler this becomes an unconditional DMUL, two selects and two integer instructions)
    static __device__ __forceinline__ void below(double &C, unsigned &n, double d, unsigned long long b, unsigned long long edge)
    {
        // the compare runs on the FP64 pipe (one DSETP) instead of two ISETP on the half-rate ALU pipe; for the values here
        // (d >= +0, edges finite or +inf) it orders exactly like the 64-bit integer compare of the bit patterns
        asm("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %2, %3;\n\t@p mul.rn.f64 %0, %0, %2;\n\t@p add.u32 %1, %1, 1;\n\t}"
            : "+d"(C), "+r"(n) : "d"(d), "d"(__longlong_as_double((long long)edge)));
        (void)b;
    }
    template <int NE>
    __device__ __forceinline__ void pair_fast(double xj, double yj)
    {
        const double dx = xs - xj, dy = ys - yj;
        const double d = dx * dx + dy * dy;
        const unsigned long long b = (unsigned long long)__double_as_longlong(d);
        Pall *= d;
        below(C0, n0, d, b, eb0);
        if (NE >= 2) below(C1, n1, d, b, eb1);
        if (NE >= 3) below(C2, n2, d, b, eb2);
    }
    template <int NE>
    __device__ __forceinline__ void renorm_all()
    {
        renorm(Pall, Eall);
        renorm(C0, E0);
        if (NE >= 2) renorm(C1, E1);
        if (NE >= 3) renorm(C2, E2);
    }
    template <int NE>
    __device__ __forceinline__ void chunk_fast(uint32_t bx, uint32_t by, int lo, int hi, uint32_t tok)
    {
        // The running products are renormalised (exponent pulled out) every P.renorm pairs instead of after every group
        // of 8: the host chose P.renorm so that that many factors of any separation the kernel can meet (between tiny2
        // and 8 L^2) keep a product that started in [1, 2) inside the double range with room to spare.
        int j = lo;
        if ((j & 1) && j < hi) { pair_fast<NE>(lds_f64(bx + 8 * j), lds_f64(by + 8 * j)); ++j; ++since; }
#pragma unroll 1
        for (; j + 8 <= hi; j += 8) {
            double xa[8], ya[8];
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                lds_f64x2_tok(bx + 8 * (j + u), tok, xa[u], xa[u + 1]);
                lds_f64x2_tok(by + 8 * (j + u), tok, ya[u], ya[u + 1]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) pair_fast<NE>(xa[u], ya[u]);
            since += 8;
            if (since >= P.renorm) { renorm_all<NE>(); since = 0; }
        }
#pragma unroll 1
        for (; j < hi; ++j) {
            pair_fast<NE>(lds_f64(bx + 8 * j), lds_f64(by + 8 * j));
            if (++since >= P.renorm) { renorm_all<NE>(); since = 0; }
        }
    }
    // exact path: the slot of a pair is slo + the number of this lane's edges below d^2; every lane tests the same
    // number of edges (uex, the widest range of the warp - edges beyond a lane's own range cannot be below d^2, and
    // the table is padded with +inf), so the loop is uniform and its loads are independent
    __device__ __forceinline__ void pair_exact(double xj, double yj)
    {
        const double dx = xs - xj, dy = ys - yj;
        const double d = dx * dx + dy * dy;
        const unsigned long long b = (unsigned long long)__double_as_longlong(d);
        int s = slo;
        const uint32_t e0 = es + 8u * (uint32_t)slo;
#pragma unroll 4
        for (int k = 0; k < uex; ++k) s += (b > lds_u64(e0 + 8u * (uint32_t)k)) ? 1 : 0;
        if (s < P.nrp) {
            cnt[s] += 1u;
            if (s >= 1) {
                Num[s] *= __longlong_as_double((long long)((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL));
                eNum[s] += (int)(b >> 52) - 1023;
            }
        }
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t tok)
    {
        const uint32_t bx = stage, by = stage + 8 * CH;
        npart += (unsigned)(hi - lo);
        if (mode == 1) chunk_fast<1>(bx, by, lo, hi, tok);
        else if (mode == 2) chunk_fast<2>(bx, by, lo, hi, tok);
        else if (mode == 3) chunk_fast<3>(bx, by, lo, hi, tok);
        else {
#pragma unroll 1
            for (int j = lo; j < hi; ++j) pair_exact(lds_f64(bx + 8 * j), lds_f64(by + 8 * j));
            // the exact pairs multiplied mantissas in [1, 2) into Num (<= CH of them): pull the exponents out
#pragma unroll 1
            for (int s = max(slo, 1); s <= min(shi, P.nrp - 1); ++s) renorm(Num[s], eNum[s]);
        }
    }
    __device__ __forceinline__ void bank(double *arr, int *earr, int s, double m, int e)
    {
        double v = arr[s] * m;
        int ee = earr[s] + e;
        renorm(v, ee);
        arr[s] = v; earr[s] = ee;
    }
    __device__ __forceinline__ void cell_end()
    {
        if (mode == 0) return;
        // (bank() multiplies a banked mantissa in [1, 2) by the running product and renormalises: the running product
        // holds fewer than P.renorm factors, which the host's bound covers)
        // cumulative products C_0 <= C_1 <= C_2 <= Pall (as sets of pairs): slot slo + i holds C_i / C_(i-1)
        cnt[slo] += n0;
        bank(Num, eNum, slo, C0, E0);
        double pm = C0; int pe = E0; unsigned pn = n0;
        if (mode >= 2) {
            cnt[slo + 1] += n1 - pn;
            bank(Num, eNum, slo + 1, C1, E1);
            bank(Den, eDen, slo + 1, pm, pe);
            pm = C1; pe = E1; pn = n1;
        }
        if (mode >= 3) {
            cnt[slo + 2] += n2 - pn;
            bank(Num, eNum, slo + 2, C2, E2);
            bank(Den, eDen, slo + 2, pm, pe);
            pm = C2; pe = E2; pn = n2;
        }
        cnt[slo + mode] += npart - pn;
        bank(Num, eNum, slo + mode, Pall, Eall);
        bank(Den, eDen, slo + mode, pm, pe);
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&idx)[1], int, unsigned)
    {
        if (valid) {
            const int nbin = P.nrp - 1;
            const double m = P.mass;
            const int64_t row = P.perm1 ? (int64_t)P.perm1[idx[0]] : (int64_t)idx[0];
            double inside = (double)cnt[0];
#pragma unroll 1
            for (int k = 0; k < nbin; ++k) {
                const int s = k + 1;
                const double n = (double)cnt[s];
                // sum over the annulus of ln(d^2 / rp[k+1]^2): exponents and mantissas kept apart
                const double r2 = P.e0[k + 1];
                const int rhi = __double2hiint(r2);
                const int re = (rhi >> 20) - 1023;
                const double rm = __hiloint2double((rhi & 0x000fffff) | 0x3ff00000, __double2loint(r2));
                const double ex = (double)(eNum[s] - eDen[s]);
                const double lm = log(Num[s]) - log(Den[s]);
                const double lnratio = 0.6931471805599453 * (ex - n * (double)re) + (lm - n * log(rm));
                const double t = m * (n + lnratio);
                const double ds = m * inside * 2 * P.e1[k] - t;
                atomicAdd(&P.out[row * nbin + k], ds / (3.14159265358979323846 * (P.e0[k + 1] - P.e0[k])));
                inside += n;
            }
        }
        return false;
    }
    __device__ __forceinline__ void kernel_end() {}
};

// ------------------------------------------------------------------ host launchers

int htb_fast3_ppl() { return FAST3_PPL; }
int htb_launch_fast3(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const Fast3Params &P, int *l)
{ return launch_count<Fast3>(st, G, A, P, l); }
int htb_launch_markedq(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const Fast3Params &P, int *l)
{ return launch_count<MarkedQ>(st, G, A, P, l); }
int htb_launch_fastxyz(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const Fast3Params &P, int *l)
{ return launch_count<FastXYZ>(st, G, A, P, l); }
int htb_launch_dsq(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const DSQParams &P, int *l)
{ return launch_count<DSigmaQ>(st, G, A, P, l); }
int htb_launch_dsr(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const DSRParams &P, int *l)
{ return launch_count<DSigmaR>(st, G, A, P, l); }
int htb_launch_gen(cudaStream_t st, int kind, const WalkGeom &G, const WalkArrays &A, const GenParams &P, int *l)
{
    switch (kind) {
    case 0: return launch_count<GenCount<0>>(st, G, A, P, l);
    case 1: return launch_count<GenCount<1>>(st, G, A, P, l);
    case 2: return launch_count<GenCount<2>>(st, G, A, P, l);
    case 3: return launch_count<Marked3>(st, G, A, P, l);
    case 4: return launch_count<DSigma>(st, G, A, P, l);
    case 5: return launch_count<DSigmaU>(st, G, A, P, l);
    }
    htb_set_error("unknown kernel kind %d", kind);
    return 1;
}

// x-layer windows of a rank's cell range: mesh1 layers [first / (ny nz), (last - 1) / (ny nz)] and the mesh2 layers
// their neighbour windows reach (npairs_3d_engine.pyx:117-123); xwin = {first layer, layers} (mesh1), {..} (mesh2)
__global__ void k_shard_windows(const long long *__restrict__ range, WalkGeom G, int *__restrict__ xwin)
{
    if (threadIdx.x || blockIdx.x) return;
    long long per0 = 1;
    for (int d = 1; d < G.dim; ++d) per0 *= G.nd1[d];
    const long long first = range[0], last = range[1];
    if (last <= first) { xwin[0] = 0; xwin[1] = 0; xwin[2] = 0; xwin[3] = 0; return; }
    const int lo = (int)(first / per0), hi = (int)((last - 1) / per0);
    xwin[0] = lo; xwin[1] = hi - lo + 1;
    xwin[2] = lo * G.per[0] - G.cover[0];
    xwin[3] = (hi - lo + 1) * G.per[0] + 2 * G.cover[0];
}

int htb_shard_windows(cudaStream_t st, const long long *range_dev, const WalkGeom &G, int *xwin_dev, int *launches)
{
    k_shard_windows<<<1, 32, 0, st>>>(range_dev, G, xwin_dev);
    if (launches) *launches += 1;
    HTB_CUDA(cudaGetLastError());
    return 0;
}

int htb_shard_range(cudaStream_t st, const double *work_dev, int64_t first_cell1, int64_t last_cell1,
                    int rank, int world, long long *range_dev, int *launches)
{
    k_shard_range<<<1, HTB_SR_THREADS, 0, st>>>(work_dev, (long long)first_cell1, (long long)last_cell1, rank, world, range_dev);
    if (launches) *launches += 1;
    HTB_CUDA(cudaGetLastError());
    return 0;
}

int htb_build_tiles(cudaStream_t st, Workspace &ws, const WalkGeom &G, const SortedSample &s1,
                    int64_t first_cell1, int64_t last_cell1, const long long *range_dev,
                    uint2 **tiles_out, uint32_t **ntiles_dev_out,
                    int64_t *max_tiles_out, int *launches)
{
    const int F = G.dim - 1;
    int64_t nseg = 1;                                   // fine columns of sample1
    for (int d = 0; d < F; ++d) nseg *= G.nf1[d];
    if (nseg >= (1 << 24)) { htb_set_error("too many fine columns in the sample1 mesh (%lld)", (long long)nseg); return 1; }
    uint32_t *ntile = nullptr, *tbase = nullptr, *total = nullptr;
    if (ws.alloc((void **)&ntile, sizeof(uint32_t) * (size_t)(nseg + 1))) return 1;
    if (ws.alloc((void **)&tbase, sizeof(uint32_t) * (size_t)(nseg + 1))) return 1;
    if (ws.alloc((void **)&total, sizeof(uint32_t) * 4)) return 1;
    // every tile is either full or ends at a reference-cell boundary of its column
    const int64_t max_tiles = s1.n / G.tile + nseg * G.nd1[F] + 1;
    uint2 *tiles = nullptr;
    if (ws.alloc((void **)&tiles, sizeof(uint2) * (size_t)max_tiles)) return 1;
    // few columns with many fine cells each: one warp per column (see k_tiles_warp); many columns: one thread each
    const bool by_warp = nseg < 148 * 16 * 32 && G.nf1[F] >= 64 && G.maxfine + 2 < HTB_TL_WIN && !getenv("HTB_TILES_BY_THREAD");
    if (by_warp) {
        int blocks = (int)((nseg + HTB_TL_WARPS - 1) / HTB_TL_WARPS);
        if (blocks > 148 * 4) blocks = 148 * 4;
        k_tiles_warp<false><<<blocks, 32 * HTB_TL_WARPS, 0, st>>>(s1.off, G, nseg, first_cell1, last_cell1, range_dev, ntile, nullptr, nullptr);
        if (launches) *launches += 1;
        if (htb_exclusive_scan_u32(st, ws, ntile, tbase, nseg, total, launches)) return 1;
        k_tiles_warp<true><<<blocks, 32 * HTB_TL_WARPS, 0, st>>>(s1.off, G, nseg, first_cell1, last_cell1, range_dev, ntile, tbase, tiles);
        if (launches) *launches += 1;
    } else {
        int blocks = (int)((nseg + 127) / 128);
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_seg_tiles<<<blocks, 128, 0, st>>>(s1.off, G, nseg, first_cell1, last_cell1, range_dev, ntile);
        if (launches) *launches += 1;
        if (htb_exclusive_scan_u32(st, ws, ntile, tbase, nseg, total, launches)) return 1;
        k_fill_tiles<<<blocks, 128, 0, st>>>(s1.off, G, nseg, first_cell1, last_cell1, range_dev, ntile, tbase, tiles);
        if (launches) *launches += 1;
    }
    HTB_CUDA(cudaGetLastError());
    *tiles_out = tiles;
    *ntiles_dev_out = total;
    *max_tiles_out = max_tiles;
    return 0;
}

int htb_reference_work(cudaStream_t st, Workspace &ws, const WalkGeom &G, const SortedSample &s1,
                       const SortedSample &s2, double **work_dev_out, double **balance_dev_out, int64_t *ncell1_out,
                       int *launches, bool presort)
{
    int64_t nc1 = 1, nc2 = 1;
    for (int d = 0; d < G.dim; ++d) { nc1 *= G.nd1[d]; nc2 *= G.nd2[d]; }
    uint32_t *rc1 = nullptr, *rc2 = nullptr;
    double *work = nullptr, *balance = nullptr;
    if (balance_dev_out && ws.alloc((void **)&balance, sizeof(double) * (size_t)nc1)) return 1;
    if (ws.alloc((void **)&rc1, sizeof(uint32_t) * (size_t)nc1)) return 1;
    if (ws.alloc((void **)&rc2, sizeof(uint32_t) * (size_t)nc2)) return 1;
    if (ws.alloc((void **)&work, sizeof(double) * (size_t)nc1)) return 1;
    HTB_CUDA(cudaMemsetAsync(rc1, 0, sizeof(uint32_t) * (size_t)nc1, st));
    HTB_CUDA(cudaMemsetAsync(rc2, 0, sizeof(uint32_t) * (size_t)nc2, st));
    if (presort) {
        if (htb_ref_cell_counts_pre(st, s1, rc1, launches)) return 1;
        if (htb_ref_cell_counts_pre(st, s2, rc2, launches)) return 1;
    } else {
        if (htb_ref_cell_counts(st, s1, rc1, launches)) return 1;
        if (htb_ref_cell_counts(st, s2, rc2, launches)) return 1;
    }
    // pruning weights of the window's cells (only for the multi-GPU cut): Monte-Carlo fraction of the pairs between a
    // mesh1 cell and the mesh2 cell at each window offset that lie within the search length, + a floor for the per-cell
    // overheads; a fixed sample set, recomputed only when the geometry changes
    float *wtab_dev = nullptr;
    if (balance) {
        static thread_local std::vector<float> tab;
        static thread_local double key[16] = {0};
        double k[16] = {(double)G.dim, (double)G.sphere, 0};
        int wn[3] = {1, 1, 1};
        size_t ntab = 1;
        for (int d = 0; d < G.dim; ++d) {
            wn[d] = G.per[d] + 2 * G.cover[d];
            ntab *= (size_t)wn[d];
            k[2 + 4 * d] = G.per[d]; k[3 + 4 * d] = G.cover[d]; k[4 + 4 * d] = G.period[d] / G.nd2[d]; k[5 + 4 * d] = G.reach[d];
        }
        if (ntab <= 4096) {
            if (tab.size() != ntab || memcmp(k, key, sizeof(k)) != 0) {
                tab.assign(ntab, 0.f);
                const int NS = 2048;
                unsigned long long rng = 0x9E3779B97F4A7C15ULL;
                auto u01 = [&]() { rng = rng * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(rng >> 11) * (1.0 / 9007199254740992.0); };
                std::vector<double> p1((size_t)NS * 3), p2((size_t)NS * 3);
                for (auto &v : p1) v = u01();
                for (auto &v : p2) v = u01();
                for (size_t t = 0; t < ntab; ++t) {
                    int o[3] = {0, 0, 0};
                    size_t rem = t;
                    for (int d = G.dim - 1; d >= 0; --d) { o[d] = (int)(rem % wn[d]); rem /= wn[d]; }
                    int hit = 0;
                    for (int i = 0; i < NS; ++i) {
                        double q = 0.0, qz = 0.0;
                        for (int d = 0; d < G.dim; ++d) {
                            const double cs2 = G.period[d] / G.nd2[d];
                            const double a = p1[(size_t)i * 3 + d] * (G.per[d] * cs2);                  // in the mesh1 cell
                            const double b = ((double)(o[d] - G.cover[d]) + p2[(size_t)i * 3 + d]) * cs2;  // in the mesh2 cell at this offset
                            const double r = G.reach[d] > 0 ? (a - b) / G.reach[d] : 0.0;
                            if (!G.sphere && d == G.dim - 1) qz = r * r; else q += r * r;
                        }
                        hit += (q <= 1.0 && qz <= 1.0) ? 1 : 0;
                    }
                    tab[t] = 0.03f + (float)hit / (float)NS;
                }
                memcpy(key, k, sizeof(k));
            }
            if (ws.alloc((void **)&wtab_dev, sizeof(float) * ntab)) return 1;
            HTB_CUDA(cudaMemcpyAsync(wtab_dev, tab.data(), sizeof(float) * ntab, cudaMemcpyHostToDevice, st));
        }
    }
    int blocks = (int)((nc1 + 127) / 128);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_wref<<<blocks, 128, 0, st>>>(rc1, rc2, G, nc1, work, balance, wtab_dev);
    if (launches) *launches += 1;
    HTB_CUDA(cudaGetLastError());
    *work_dev_out = work;
    if (balance_dev_out) *balance_dev_out = balance;
    *ncell1_out = nc1;
    return 0;
}
