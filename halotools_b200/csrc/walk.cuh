// K2 skeleton — the cell-pair "walker" shared by every pair-counting kernel.
//
// One warp owns one TILE of <= 64 sample1 points (2 per lane, held in registers) that
// lie in one fine column and one reference mesh1 cell.  For that tile the warp
//   1. rebuilds the reference's neighbour window (npairs_3d_engine.pyx:113-153) in
//      units of the refined sample2 grid, keeping the reference's per-cell periodic
//      shift (-L / 0 / +L decided by the UNWRAPPED reference cell index),
//   2. prunes columns / z-runs whose cells are provably farther than the search
//      radius from the tile's bounding box (count-preserving),
//   3. turns every surviving (column, wrap piece) into ONE contiguous span of the
//      z-fastest sorted sample2 arrays,
//   4. streams the spans through a per-warp shared-memory ring filled by 1-D TMA
//      bulk copies (cp.async.bulk + mbarrier complete_tx), and
//   5. hands each staged chunk to the variant's pair functor.
// Variants differ only in the functor (bins, weights, accumulators).
#pragma once
#include "htb_internal.cuh"

#define HTB_WARPS 8                 // warps per block of the counting kernels (overridable per variant)
#ifndef HTB_CH
#define HTB_CH 64                   // sample2 points per staged chunk
#endif
#ifndef HTB_NSTAGE
#define HTB_NSTAGE 2
#endif
#ifndef HTB_SPAN_CAP
#define HTB_SPAN_CAP 96
#endif
#define HTB_FULL 0xffffffffu
#define HTB_REDO_SPLIT 16           // an exact re-evaluation of a tile slice is cut into this many finer slices (<= 32)
#ifndef HTB_ITEMS_PER_WARP
#define HTB_ITEMS_PER_WARP 32       // tiles are sliced until every resident warp has about this many work items
#endif

struct WalkGeom {
    int dim;
    int pbc;
    int sphere;                     // 1: fast-dim reach shrinks with slow-dim distance (3-D r / 2-D rp); 0: cylinder (rp, pi)
    int nocull;
    int sym;                        // sample1 IS sample2 (same sorted arrays): count each zero-shift unordered pair once, weight 2
    int nd1[3], nd2[3], per[3], cover[3];
    int m1[3], m2[3], nf1[3], nf2[3];
    double period[3], h2[3], slop[3], reach[3];
    double r2slow;                  // (max separation over the slow dims)^2, with safety margin
    double sentinel;                // x coordinate given to the unused lanes of a partial tile (never in range)
    double cs1f;                    // reference mesh1 cell size along the fast dimension (as handed to the mesh sort)
    int tile;                       // sample1 points per tile (32 * points per lane of the kernel variant)
    int maxspan;                    // reference mesh1 cells along the fast dimension one tile may straddle (>= 1)
    int maxfine;                    // fine mesh1 cells along the fast dimension one tile may cover (bounds its extent)
    int maxslices;                  // a tile's sample2 columns may be cut into up to this many independent work items
    int items_per_warp;             // ... until every resident warp has about this many work items
    int early_exit;                 // idle warps do not wait for the redo queue to drain (HTB_FLAG_EARLY_EXIT)
    int tail_eighths;               // the last (resident warps x this / 8) tiles are cut into maxslices slices (0: off)
};

struct WalkArrays {
    const double *c1[3];            // sorted sample1 coords
    const uint32_t *off1;
    const double *c2[3];            // sorted sample2 coords
    const uint32_t *off2;
    const double *pay1;             // sorted sample1 payload rows (n1, nw) or null
    const double *pay2;             // sorted sample2 payload rows (n2, nw) or null
    int nw;
    const uint32_t *flags1, *flags2;
    const uint2 *tiles;             // {first sorted index, segment id}
    const uint32_t *ntiles_dev;     // [1] number of tiles (device resident; no host sync)
    unsigned int *tile_counter;
    unsigned long long *pairs_evaluated;
    unsigned int *tiles_redone;
    uint2 *redo_ent;                // exact re-evaluations waiting for a warp: {tile, weight pass << 24 | fine slice}
    unsigned int *redo_ctr;         // [0] reserved, [1] published, [2] taken, [3] ordinary items completed
    unsigned int redo_cap;          // entries redo_ent can hold (0: re-evaluate in place)
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
    return ok;      // data dependence for loads that must not be scheduled above the wait
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// explicit shared-space loads (the staging pointers reach the functors as generic pointers otherwise)
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double &a, double &b)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(HTB_FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(HTB_FULL, v, o));
    return v;
}
__device__ __forceinline__ int floor_div(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

// Per-warp shared-memory context: HTB_NSTAGE stage buffers (DIM coordinate rows of HTB_CH doubles + payload rows of
// HTB_CH * NPAY doubles), one mbarrier per stage, the span list.  Plain scalars only (no arrays indexed at run time,
// which would live in local memory).
template <int DIM, int NPAY, int CH = HTB_CH>
struct WarpSmem {
    double *stage0;                 // generic pointer to stage 0
    uint32_t stage0_s;              // its shared-space address
    uint32_t bar0;                  // shared-space address of mbarrier 0 (16 bytes apart)
    uint32_t *span;                 // HTB_SPAN_CAP * 3 u32: {jb, je, code}
    // (+ 2: the register double-buffering of the fast kernels reads one pair of doubles past the last row of a stage;
    // without the pad that read lands in the next stage, which a TMA copy may be filling - harmless, the value is never
    // used, but racecheck reports it)
    static __host__ __device__ constexpr int stage_doubles() { return CH * (DIM + NPAY) + 2; }
    static __host__ __device__ constexpr size_t bytes()
    {
        return sizeof(double) * HTB_NSTAGE * stage_doubles() + 16 * HTB_NSTAGE + sizeof(uint32_t) * 3 * HTB_SPAN_CAP;
    }
    __device__ __forceinline__ double *stage(int s) const { return stage0 + s * stage_doubles(); }
    __device__ __forceinline__ uint32_t stage_s(int s) const { return stage0_s + (uint32_t)(s * stage_doubles() * 8); }
    __device__ __forceinline__ uint32_t bar(int s) const { return bar0 + 16u * (uint32_t)s; }
};

// sample2 points per staged chunk of a variant: V::CH if it declares one, else HTB_CH
template <class V, class = void> struct HtbChunk { static constexpr int value = HTB_CH; };
template <class V> struct HtbChunk<V, decltype((void)V::CH)> { static constexpr int value = V::CH; };

// variants that want to know when a chunk is the tile's OWN index range in symmetric mode (it holds the self pairs,
// whose exact zeros a 32-bit relative key cannot represent): V::HAS_SELF and chunk(stage, lo, hi, tok, own)
template <class V, class = void> struct HtbHasSelf { static constexpr bool value = false; };
template <class V> struct HtbHasSelf<V, decltype((void)V::HAS_SELF)> { static constexpr bool value = V::HAS_SELF; };

// variants that need the sorted index of staged slot 0 (to name sample2 points in a queue): V::WANTS_BASE and chunk_base()
template <class V, class = void> struct HtbWantsBase { static constexpr bool value = false; };
template <class V> struct HtbWantsBase<V, decltype((void)V::WANTS_BASE)> { static constexpr bool value = V::WANTS_BASE; };

struct TileInfo {
    int cnt;                // valid points in the tile
    uint32_t start;         // first sorted index
    bool v0, v1;            // validity of this lane's two points
};

// The walker.  V must provide:
//   static constexpr int DIM, NPAY, PPL; static constexpr bool TMA;
//   __device__ void set_shift(const double (&sh)[3], const WalkArrays &A) — the periodic shift of the following chunks
//        (the variant keeps its points' shifted coordinates, x1 - shift first as in npairs_3d_engine.pyx:167)
//   __device__ void chunk(uint32_t stage_smem_addr, int lo, int hi, uint32_t tok) — evaluate pairs between the lane's
//        PPL points and staged sample2 entries [lo, hi); tok is a value that depends on the stage's mbarrier wait
//        (an input for loads that may be scheduled freely).
template <class V>
__device__ __forceinline__ void walk_tile(V &v, const WalkGeom &G, const WalkArrays &A,
                                          WarpSmem<V::DIM, V::NPAY> &S, uint32_t &gchunk,
                                          const double (&blo)[3], const double (&bhi)[3],
                                          const int (&fs)[3] /* tile's fine/ref indices: slow dims fine idx, fast dim FIRST ref cell */,
                                          const int nref /* reference cells along the fast dimension the tile straddles */,
                                          unsigned long long &pairs, int tile_cnt,
                                          int wt_pass /* 0: every span; 1: symmetric mode, weight-1 spans; 2: weight-2 spans */,
                                          uint32_t ts, uint32_t te /* the tile's own sorted index range */,
                                          const int slice, const int nslices /* this work item's share of the columns */)
{
    constexpr int DIM = V::DIM;
    constexpr int F = DIM - 1;                 // fast dimension
    const int lane = threadIdx.x & 31;
    const bool cull = !G.nocull && !((A.flags1[0] | A.flags2[0]) & 1u);

    // reference cell1 index per dim and the window in fine-2 units
    int a[3], wlo[3], wn[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        a[d] = (d == F) ? fs[d] : fs[d] / G.m1[d];
        wlo[d] = (a[d] * G.per[d] - G.cover[d]) * G.m2[d];
        // A tile that straddles nref reference cells along the fast dimension walks the union of their windows:
        // every cell of the union keeps the shift of its own unwrapped index, the cells a point would not have
        // visited in the reference are farther than the search length from it (cover cells away), and the host
        // only allows nref > 1 when the union cannot reach the same cell twice (maxspan).
        wn[d] = (G.per[d] * (d == F ? nref : 1) + 2 * G.cover[d]) * G.m2[d];
    }
    const int ncol_all = (DIM == 3) ? wn[0] * wn[1] : wn[0];
    const int col0 = (int)((long long)ncol_all * slice / nslices);          // this slice: columns [col0, ncol)
    const int ncol = (int)((long long)ncol_all * (slice + 1) / nslices);
    const int kfmin = floor_div(wlo[F], G.nf2[F]);
    const int kfmax = floor_div(wlo[F] + wn[F] - 1, G.nf2[F]);

    int nspan = 0;

    uint32_t cur_code = 0xffffffffu;          // shift code the variant currently holds (none yet for this tile pass)
    auto consume = [&]() {
        // ---- stream the span list through the staging ring
        int si = 0, sc = 0;                    // issue / compute span cursors
        uint32_t ji = 0, jc = 0;               // aligned start of the next chunk
        if (nspan > 0) { ji = jc = S.span[0] & ~1u; }
        int inflight = 0;
        while (sc < nspan) {
            while (si < nspan && inflight < HTB_NSTAGE) {
                const uint32_t je = S.span[3 * si + 1];
                const uint32_t jend = (je + 1u) & ~1u;
                const uint32_t cnt = min((uint32_t)HTB_CH, jend - ji);
                const int stg = (int)((gchunk + (uint32_t)inflight) % HTB_NSTAGE);
                if (V::TMA) {
                    if (lane == 0) {
                        const uint32_t dst = S.stage_s(stg), bar = S.bar(stg);
                        const uint32_t bytes = cnt * 8u * (DIM + A.nw * (V::NPAY > 0 ? 1 : 0));
                        mbar_expect_tx(bar, bytes);
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            tma_bulk_g2s(dst + (uint32_t)(d * HTB_CH * 8), A.c2[d] + ji, cnt * 8u, bar);
                        if (V::NPAY > 0)
                            tma_bulk_g2s(dst + (uint32_t)(DIM * HTB_CH * 8), A.pay2 + (size_t)ji * A.nw, cnt * 8u * A.nw, bar);
                    }
                } else {
                    double *dst = S.stage(stg);
                    for (uint32_t q = lane; q < cnt; q += 32) {
#pragma unroll
                        for (int d = 0; d < DIM; ++d) dst[d * HTB_CH + q] = A.c2[d][ji + q];
                    }
                    if (V::NPAY > 0)
                        for (uint32_t q = lane; q < cnt * A.nw; q += 32) dst[DIM * HTB_CH + q] = A.pay2[(size_t)ji * A.nw + q];
                }
                ++inflight;
                ji += HTB_CH;
                if (ji >= je) { ++si; if (si < nspan) ji = S.span[3 * si] & ~1u; }
            }
            // ---- current chunk
            const uint32_t jb = S.span[3 * sc], je = S.span[3 * sc + 1], fullcode = S.span[3 * sc + 2];
            const uint32_t code = fullcode & 0xffu;      // bit 8: the span is the tile's own index range (symmetric mode)
            const int stg = (int)(gchunk % HTB_NSTAGE);
            if (code != cur_code) {
                double sh[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    const int k = (int)((code >> (2 * d)) & 3u) - 1;
                    sh[d] = (double)(k * G.pbc) * G.period[d];
                }
                v.set_shift(sh, A);
                cur_code = code;
            }
            uint32_t tok = jc;
            if (V::TMA) tok += mbar_wait(S.bar(stg), (gchunk / HTB_NSTAGE) & 1u);
            else __syncwarp();
            const int lo = (int)(max(jb, jc) - jc);
            const int hi = (int)(min(je, jc + HTB_CH) - jc);
            if constexpr (HtbWantsBase<V>::value) v.chunk_base(jc);
            if constexpr (HtbHasSelf<V>::value) v.chunk(S.stage_s(stg), lo, hi, tok, (fullcode & 0x100u) != 0u);
            else v.chunk(S.stage_s(stg), lo, hi, tok);
            pairs += (unsigned long long)(hi - lo) * (unsigned)tile_cnt;
            __syncwarp();
            ++gchunk;
            --inflight;
            jc += HTB_CH;
            if (jc >= je) { ++sc; if (sc < nspan) jc = S.span[3 * sc] & ~1u; }
        }
        nspan = 0;
    };

    // generate spans (lane-parallel over columns) until the list is nearly full, then stream them
    int base = col0, kf = kfmin;
    bool more = ncol > col0;
    while (more) {
        while (true) {
            if (base >= ncol) { more = false; break; }
            const int col = base + lane;
            const bool valid = col < ncol;
            int U[3] = {0, 0, 0};
            if (DIM == 3) { U[0] = wlo[0] + col / wn[1]; U[1] = wlo[1] + col % wn[1]; }
            else { U[0] = wlo[0] + col; }
            double d2 = 0.0;
            uint32_t code = 0;
            int64_t slowlin = 0;
#pragma unroll
            for (int d = 0; d < F; ++d) {
                const int k = floor_div(U[d], G.nf2[d]);
                const int w = U[d] - k * G.nf2[d];
                const int kc = k < -1 ? -1 : (k > 1 ? 1 : k);   // the reference only distinguishes <0 / inside / >= ndivs2
                const double elo = (double)w * G.h2[d] + (double)(kc * G.pbc) * G.period[d] - G.slop[d];
                const double ehi = elo + G.h2[d] + 2.0 * G.slop[d];
                const double gap = fmax(0.0, fmax(elo - bhi[d], blo[d] - ehi));
                d2 += gap * gap;
                code |= (uint32_t)(kc + 1) << (2 * d);
                slowlin = slowlin * G.nf2[d] + w;
            }
            const bool keep = valid && (!cull || d2 <= G.r2slow);
            double reach = 0.0;
            if (cull) {
                reach = G.sphere ? sqrt(fmax(G.r2slow - d2, 0.0)) : G.reach[F];
                reach += G.slop[F];
            }
            {
                const int k = kf;
                const int kc = k < -1 ? -1 : (k > 1 ? 1 : k);
                int plo = max(wlo[F], k * G.nf2[F]);
                int phi = min(wlo[F] + wn[F], (k + 1) * G.nf2[F]) - 1;
                bool has = keep && plo <= phi;
                if (has && cull) {
                    // effective coordinate of fine cell U in this piece: (U - k*nf2)*h2 + kc*pbc*L
                    const double offc = -(double)k * G.period[F] + (double)(kc * G.pbc) * G.period[F];
                    const double qlo = (blo[F] - reach - offc) / G.h2[F];
                    const double qhi = (bhi[F] + reach - offc) / G.h2[F];
                    plo = (int)fmax(floor(qlo), (double)plo);
                    phi = (int)fmin(floor(qhi), (double)phi);
                    has = plo <= phi;
                }
                uint32_t jb = 0, je = 0;
                bool own = false;
                if (has) {
                    const int64_t cbase = slowlin * G.nf2[F] - (int64_t)k * G.nf2[F];
                    jb = A.off2[cbase + plo];
                    je = A.off2[cbase + phi + 1];
                    if (wt_pass != 0) {
                        // Symmetric auto-correlation: a pair evaluated with zero shift in every dimension
                        // has bit-identical dsq in both directions (IEEE subtraction is antisymmetric), so
                        // it is evaluated once, from the point with the smaller sorted index, and counted
                        // twice.  The tile's own index range is evaluated in full with weight 1 (it holds
                        // the self pairs); wrapped spans are evaluated from both sides as the reference does.
                        const uint32_t fullcode = code | ((uint32_t)(kc + 1) << (2 * F));
                        const bool zero = fullcode == (DIM == 3 ? 21u : 5u);
                        if (zero) {
                            if (wt_pass == 1) { jb = max(jb, ts); je = min(je, te); own = true; }
                            else { jb = max(jb, te); }
                        } else if (wt_pass == 2) {
                            je = jb;
                        }
                    }
                    has = je > jb;
                }
                const uint32_t bal = __ballot_sync(HTB_FULL, has);
                if (has) {
                    const int pos = nspan + __popc(bal & ((1u << lane) - 1u));
                    S.span[3 * pos] = jb;
                    S.span[3 * pos + 1] = je;
                    S.span[3 * pos + 2] = code | ((uint32_t)(kc + 1) << (2 * F)) | (own ? 0x100u : 0u);
                }
                nspan += __popc(bal);
            }
            if (++kf > kfmax) { kf = kfmin; base += 32; }
            if (nspan + 32 > HTB_SPAN_CAP) break;
        }
        __syncwarp();
        consume();
        __syncwarp();
    }
}

// ------------------------------------------------------------------ cell-resolved walker
// walk_tile() for variants that take warp-wide decisions PER FINE CELL of sample2 (V::PER_CELL): the spans are the
// same runs of fine cells along the fast dimension, streamed by the same TMA ring, but the compute side cuts every
// run at the cell boundaries (one coalesced load of the run's offsets, then shuffles) and brackets the staged
// points of each non-empty cell with
//   v.cell_begin(xlo, xhi, ylo, yhi)    the cell's bounding box in sample2's own (unshifted) coordinates, slop included
//   v.chunk(stage, lo, hi, tok) ...     as in walk_tile
//   v.cell_end()
// 2-D only (DIM == 2), no symmetric mode.  V::span_extra() gives 2 * HTB_SPAN_CAP extra u32 of per-warp scratch.
template <class V>
__device__ __forceinline__ void walk_tile_cells(V &v, const WalkGeom &G, const WalkArrays &A,
                                                WarpSmem<V::DIM, V::NPAY, HtbChunk<V>::value> &S, uint32_t &gchunk,
                                                const double (&blo)[3], const double (&bhi)[3],
                                                const int (&fs)[3], const int nref,
                                                unsigned long long &pairs, int tile_cnt,
                                                const int slice, const int nslices)
{
    static_assert(V::DIM == 2, "cell-resolved walker: 2-D meshes only");
    constexpr int DIM = 2;
    constexpr int F = 1;
    constexpr int CH = HtbChunk<V>::value;
    const int lane = threadIdx.x & 31;
    const bool cull = !G.nocull && !((A.flags1[0] | A.flags2[0]) & 1u);
    uint32_t *extra = v.span_extra();          // per span: {first fine cell of the run (unwrapped, biased), column | ncells << 16}

    int a[3], wlo[3], wn[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        a[d] = (d == F) ? fs[d] : fs[d] / G.m1[d];
        wlo[d] = (a[d] * G.per[d] - G.cover[d]) * G.m2[d];
        wn[d] = (G.per[d] * (d == F ? nref : 1) + 2 * G.cover[d]) * G.m2[d];
    }
    const int ncol_all = wn[0];
    const int col0 = (int)((long long)ncol_all * slice / nslices);
    const int ncol = (int)((long long)ncol_all * (slice + 1) / nslices);
    const int kfmin = floor_div(wlo[F], G.nf2[F]);
    const int kfmax = floor_div(wlo[F] + wn[F] - 1, G.nf2[F]);

    int nspan = 0;
    uint32_t cur_code = 0xffffffffu;
    auto consume = [&]() {
        int si = 0, sc = 0;
        uint32_t ji = 0, jc = 0;
        if (nspan > 0) { ji = jc = S.span[0] & ~1u; }
        int inflight = 0;
        // state of the run being computed
        int run = -1;                          // span index the cell cursor belongs to
        int cell = 0, ncell = 0;               // cell cursor inside the run, cells in the run
        int cbatch = 0;                        // first cell of the offsets held in `bnd`
        uint32_t bnd = 0;                      // lane l: end offset of cell cbatch + l of the run
        uint32_t cell_end = 0;
        int U0 = 0, ufirst = 0;
        int64_t cbase = 0;
        bool began = false;
        while (sc < nspan) {
            while (si < nspan && inflight < HTB_NSTAGE) {
                const uint32_t je = S.span[3 * si + 1];
                const uint32_t jend = (je + 1u) & ~1u;
                const uint32_t cnt = min((uint32_t)CH, jend - ji);
                const int stg = (int)((gchunk + (uint32_t)inflight) % HTB_NSTAGE);
                if (lane == 0) {
                    const uint32_t dst = S.stage_s(stg), bar = S.bar(stg);
                    mbar_expect_tx(bar, cnt * 8u * DIM);
#pragma unroll
                    for (int d = 0; d < DIM; ++d)
                        tma_bulk_g2s(dst + (uint32_t)(d * CH * 8), A.c2[d] + ji, cnt * 8u, bar);
                }
                ++inflight;
                ji += CH;
                if (ji >= je) { ++si; if (si < nspan) ji = S.span[3 * si] & ~1u; }
            }
            const uint32_t jb = S.span[3 * sc], je = S.span[3 * sc + 1], code = S.span[3 * sc + 2];
            const int stg = (int)(gchunk % HTB_NSTAGE);
            if (code != cur_code) {
                double sh[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    const int k = (int)((code >> (2 * d)) & 3u) - 1;
                    sh[d] = (double)(k * G.pbc) * G.period[d];
                }
                v.set_shift(sh, A);
                cur_code = code;
            }
            if (run != sc) {
                // a new run: its cells and their offsets
                run = sc;
                const uint32_t e0 = extra[2 * sc], e1 = extra[2 * sc + 1];
                ufirst = (int)e0 - (1 << 30);                 // unwrapped fine index (fast dim) of the run's first cell
                U0 = wlo[0] + (int)(e1 & 0xffffu);            // unwrapped fine column
                ncell = (int)(e1 >> 16);
                const int k0 = floor_div(U0, G.nf2[0]);
                const int kf = floor_div(ufirst, G.nf2[F]);
                cbase = (int64_t)(U0 - k0 * G.nf2[0]) * G.nf2[F] - (int64_t)kf * G.nf2[F];
                cell = 0; cbatch = 0;
                bnd = A.off2[cbase + ufirst + 1 + min(lane, ncell - 1)];
                cell_end = __shfl_sync(HTB_FULL, bnd, 0);
                began = false;
            }
            uint32_t tok = jc;
            tok += mbar_wait(S.bar(stg), (gchunk / HTB_NSTAGE) & 1u);
            uint32_t pos = max(jb, jc);
            const uint32_t end = min(je, jc + CH);
            pairs += (unsigned long long)(end - pos) * (unsigned)tile_cnt;
            while (pos < end) {
                while (cell_end <= pos) {
                    // the cell is exhausted: close it, move to the next one of the run
                    if (began) { v.cell_end(); began = false; }
                    ++cell;
                    if (cell - cbatch >= 32) {
                        cbatch = cell;
                        bnd = A.off2[cbase + ufirst + 1 + min(cbatch + lane, ncell - 1)];
                    }
                    cell_end = __shfl_sync(HTB_FULL, bnd, cell - cbatch);
                }
                if (!began) {
                    const int u = ufirst + cell;
                    const int kf = floor_div(u, G.nf2[F]), k0 = floor_div(U0, G.nf2[0]);
                    const double x0 = (double)(U0 - k0 * G.nf2[0]) * G.h2[0], y0 = (double)(u - kf * G.nf2[F]) * G.h2[F];
                    v.cell_begin(x0 - G.slop[0], x0 + G.h2[0] + G.slop[0], y0 - G.slop[F], y0 + G.h2[F] + G.slop[F]);
                    began = true;
                }
                const uint32_t seg = min(end, cell_end);
                v.chunk(S.stage_s(stg), (int)(pos - jc), (int)(seg - jc), tok);
                pos = seg;
            }
            __syncwarp();
            ++gchunk;
            --inflight;
            jc += CH;
            if (jc >= je) {
                if (began) { v.cell_end(); began = false; }
                ++sc;
                if (sc < nspan) jc = S.span[3 * sc] & ~1u;
            }
        }
        nspan = 0;
    };

    int base = col0, kf = kfmin;
    bool more = ncol > col0;
    while (more) {
        while (true) {
            if (base >= ncol) { more = false; break; }
            const int col = base + lane;
            const bool valid = col < ncol;
            const int U = wlo[0] + col;
            double d2 = 0.0;
            uint32_t code = 0;
            int64_t slowlin = 0;
            {
                const int k = floor_div(U, G.nf2[0]);
                const int w = U - k * G.nf2[0];
                const int kc = k < -1 ? -1 : (k > 1 ? 1 : k);
                const double elo = (double)w * G.h2[0] + (double)(kc * G.pbc) * G.period[0] - G.slop[0];
                const double ehi = elo + G.h2[0] + 2.0 * G.slop[0];
                const double gap = fmax(0.0, fmax(elo - bhi[0], blo[0] - ehi));
                d2 = gap * gap;
                code = (uint32_t)(kc + 1);
                slowlin = w;
            }
            const bool keep = valid && (!cull || d2 <= G.r2slow);
            double reach = 0.0;
            if (cull) reach = sqrt(fmax(G.r2slow - d2, 0.0)) + G.slop[F];
            {
                const int k = kf;
                const int kc = k < -1 ? -1 : (k > 1 ? 1 : k);
                int plo = max(wlo[F], k * G.nf2[F]);
                int phi = min(wlo[F] + wn[F], (k + 1) * G.nf2[F]) - 1;
                bool has = keep && plo <= phi;
                if (has && cull) {
                    const double offc = -(double)k * G.period[F] + (double)(kc * G.pbc) * G.period[F];
                    const double qlo = (blo[F] - reach - offc) / G.h2[F];
                    const double qhi = (bhi[F] + reach - offc) / G.h2[F];
                    plo = (int)fmax(floor(qlo), (double)plo);
                    phi = (int)fmin(floor(qhi), (double)phi);
                    has = plo <= phi;
                }
                uint32_t jb = 0, je = 0;
                if (has) {
                    const int64_t cb = slowlin * G.nf2[F] - (int64_t)k * G.nf2[F];
                    jb = A.off2[cb + plo];
                    je = A.off2[cb + phi + 1];
                    has = je > jb;
                }
                const uint32_t bal = __ballot_sync(HTB_FULL, has);
                if (has) {
                    const int pos = nspan + __popc(bal & ((1u << lane) - 1u));
                    S.span[3 * pos] = jb;
                    S.span[3 * pos + 1] = je;
                    S.span[3 * pos + 2] = code | ((uint32_t)(kc + 1) << (2 * F));
                    extra[2 * pos] = (uint32_t)(plo + (1 << 30));
                    extra[2 * pos + 1] = (uint32_t)col | ((uint32_t)(phi - plo + 1) << 16);
                }
                nspan += __popc(bal);
            }
            if (++kf > kfmax) { kf = kfmin; base += 32; }
            if (nspan + 32 > HTB_SPAN_CAP) break;
        }
        __syncwarp();
        consume();
        __syncwarp();
    }
}
