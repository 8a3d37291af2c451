// K2e — BinQ: pair counts on ANY monotone set of edges (no limit of 16), one or two bin axes.
//   KIND 0  3-D r          npairs_3d with more than 16 rbins            npairs_3d_engine.pyx:171-182
//   KIND 1  (rp, pi)       npairs_xy_z with any number of pi edges      npairs_xy_z_engine.pyx:178-194
//   KIND 2  (s, mu)        npairs_s_mu                                  npairs_s_mu_engine.pyx:196-229
//   KIND 3  2-D rp         weighted_npairs_xy                           weighted_npairs_xy_engine.pyx:150-175
//   MODE 0  integer counts (shared-memory u32 histogram per warp)
//   MODE 1  weighted sums  marked_npairs_xy_z / marked_npairs_3d with general marks / weighted_npairs_xy
//                          marked_npairs_xy_z_engine.pyx:209-225, marked_npairs_3d_engine.pyx:204-216
//   MODE 2  per-object     npairs_per_object_3d                         npairs_per_object_3d_engine.pyx:190-207
//   MODE 4  per-object weighted rows (weight = sample2's w2[0]), input order
//                          weighted_npairs_per_object_xy_engine.pyx:150-185
//   MODE 5  MODE 4's lane-private rows, reduced over the warp at the end of the tile into ONE row (few bins: no
//           shared-memory atomics in the replay)   weighted_npairs_xy_engine.pyx:150-175
//   MODE 6  MODE 3 with the ROLES OF THE SAMPLES EXCHANGED (the lanes hold sample2 points, sample1 is staged): the
//           periodic shift is applied to the STAGED coordinate, so that the separation is the reference's own
//           (x1 - shift) - x2, bit for bit (pass B of the jackknife counters)
//   MODE 3  per-object weighted sums folded by the point's jackknife tag (payload rows {weight, tag})
//                          npairs_jackknife_3d_engine.pyx:213-233, npairs_jackknife_xy_z_engine.pyx:222-246
// The hot loop only DECIDES whether a pair can be inside the top edge(s): the reference's strict f64 separation
// (5-8 operations), one integer compare of its high word per bin axis against the high word of the top squared
// edge (a conservative superset: the high word of a non-negative double is monotone), and one predicated integer
// instruction that records the pair as a BIT of a per-lane 64-bit mask (one bit per staged sample2 point and lane
// point) - no shared-memory queue, no stores.  At the end of every staged chunk the set bits are replayed - by ALL
// lanes in equal shares (the set bits of a quarter chunk are first laid out as one list in shared memory): the
// separation is recomputed with the same arithmetic, located among the edges by a table lookup on its exponent and
// leading mantissa bits (first edge that can be >= the value; exact 64-bit compares of the raw bit patterns - order
// preserving for non-negative doubles - only where an edge shares the value's table cell), and the pair (or its weight) is added to the DIFFERENTIAL
// histogram cell (lowest satisfied edge per axis).  The host turns the differential histogram into the reference's
// cumulative counts (for monotone edges the reference's top-down scans stop exactly at the lowest satisfied edge).
#include "kernel.cuh"

__device__ __forceinline__ void bq_lds_f64x2_tok(uint32_t addr, uint32_t tok, double &a, double &b)
{
    // not volatile: free to be scheduled early; `tok` changes with every staged chunk (no merging across chunks)
    asm("ld.shared.v2.f64 {%0, %1}, [%2]; // %3" : "=d"(a), "=d"(b) : "r"(addr), "r"(tok));
}
__device__ __forceinline__ unsigned long long bq_lds_u64(uint32_t addr)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

template <int KIND, int MODE>
struct BinQ {
    static constexpr int DIM = KIND == 3 ? 2 : 3, NPAY = MODE == 1 ? HTB_MAX_NW : ((MODE == 3 || MODE == 6) ? 2 : (MODE >= 4 ? 1 : 0)), PPL = 2,
                         WARPS = MODE >= 3 ? 4 : 8,            // per-point f64 rows: smaller blocks, so that wide rows still fit
                         MINBLOCKS = MODE >= 3 ? 4 : 2;        // 16 warps per SM either way (128 registers per thread)
    static constexpr bool TMA = true;
    static constexpr bool JK = MODE == 3 || MODE == 6, REV = MODE == 6;
    typedef BinQParams Params;
    const Params &P;
    double *shist;              // JK, wide rows: one shared-memory histogram for tiles whose points share a tag
    bool wide, uni;             // JK: rows in global memory / the current tile's points share the tag `utag`
    int utag;
    double rs0, rs1, rs2;       // REV: the (warp-uniform) periodic shift of the current span, NEGATED (x1 - shift_A = x1 + shift_B)
    uint32_t *hist;             // MODE 0: per-warp differential histogram (n0 * n1 u32); MODE 2: 64 rows of `rstride` u32
    double *fhist;              // MODE 1: per-warp differential float sums (n0 * n1); MODE 3: 64 rows of `rstride` f64
    uint32_t e_s;               // shared-space address: raw bits of the squared edges (n0 then n1), u64
    uint32_t lut_s[2];          // shared-space addresses of the two lookup tables (u8)
    int lane;
    int rstride;                // MODE 2: u32 per point row (odd: no bank conflicts between lanes)
    unsigned vmask;             // MODE 2: validity of this lane's points
    double x0, y0, z0, x1, y1, z1;
    double xs0, ys0, zs0, xs1, ys1, zs1;
    double wa[MODE == 1 ? HTB_MAX_NW : 1], wb[MODE == 1 ? HTB_MAX_NW : 1];
    int tag[2];                 // MODE 3: jackknife tags of this lane's points
    // balanced replay: the tile's 64 points {x - shift, y - shift, z - shift, weight(s)} and the list of the recorded
    // pairs of a 16-point quarter chunk live in shared memory, so that ANY lane can replay ANY pair
    static constexpr int PT_BYTES = MODE == 1 ? 64 : 32;         // MODE 1 keeps the point's (up to) 5 marks next to it
    static constexpr int BAL_BYTES = 64 * PT_BYTES + 2 * 1024;
    uint32_t pts_s, list_s;
    // MODE 1 with few cells and MODE 5: every lane sums the pairs IT replays into its own row (no shared-memory f64
    // atomics: all lanes hitting a handful of cells serialise), the rows are reduced over the warp at the end of the tile
    static constexpr int PRIV_MAX = 48;
    bool priv;

    static __host__ __device__ size_t lut_bytes(const Params &p) { return (((size_t)p.T[0] + 7) & ~(size_t)7) + (((size_t)p.T[1] + 7) & ~(size_t)7); }
    static size_t scratch_bytes(const Params &p)
    {
        const size_t ne = (size_t)p.n0 + p.n1, nh = (size_t)p.n0 * p.n1;
        size_t acc;
        if (MODE == 0) acc = 4 * ((nh + 3) & ~(size_t)3);
        else if (MODE == 1) acc = nh <= PRIV_MAX ? 8 * 32 * (nh | 1) : 8 * nh;
        else if (MODE == 2) acc = 4 * ((64 * (size_t)(p.n0 | 1) + 3) & ~(size_t)3);
        else if (MODE == 5) acc = 8 * 32 * (nh | 1);
        else if (JK && nh > HTB_JK_SHARED_CELLS) acc = 8 * nh;            // rows in global memory (P.grows) + one histogram
        else acc = 8 * 64 * (nh | 1);
        return BAL_BYTES + 8 * ne + lut_bytes(p) + acc;
    }
    __device__ BinQ(const Params &p, void *scratch, int ln, const WalkArrays &) : P(p), lane(ln)
    {
        const int ne = P.n0 + P.n1;
        // edges and lookup tables are one contiguous block of 8-byte words on the device
        const int nl = (int)(lut_bytes(P) >> 3);
        pts_s = smem_u32(scratch);                  // 16-byte aligned (LDS.128)
        list_s = pts_s + 64 * PT_BYTES;
        unsigned long long *e = (unsigned long long *)((unsigned char *)scratch + BAL_BYTES);
        hist = (uint32_t *)(e + ne + nl);
        fhist = (double *)(e + ne + nl);
        e_s = smem_u32(e);
        lut_s[0] = e_s + 8u * (uint32_t)ne;
        lut_s[1] = lut_s[0] + (uint32_t)((P.T[0] + 7) & ~7);
        rstride = (MODE >= 3 || MODE == 1) ? ((P.n0 * P.n1) | 1) : (P.n0 | 1);
        wide = JK && P.n0 * P.n1 > HTB_JK_SHARED_CELLS;
        uni = false; utag = 0;
        shist = fhist;
        if (wide) {
            // Wide point rows (rp_pi_tpcf_jackknife): tiles whose points share one tag - nearly all of them, the
            // sub-volumes being spatial - sum into ONE shared-memory histogram; only mixed tiles use per-point rows,
            // which then live in this warp's slab of global memory.
            const unsigned gw = blockIdx.x * (unsigned)WARPS + (threadIdx.x >> 5);
            if (gw >= P.grows_warps) __trap();
            fhist = P.grows + (size_t)gw * 64u * (size_t)rstride;
            for (int k = lane; k < P.n0 * P.n1; k += 32) shist[k] = 0.0;
        }
        priv = MODE == 5 || (MODE == 1 && P.n0 * P.n1 <= PRIV_MAX);
        vmask = 0;
        for (int k = lane; k < ne + nl; k += 32) e[k] = P.edges[k];
        if (MODE == 0) { for (int k = lane; k < P.n0 * P.n1; k += 32) hist[k] = 0; }
        else if (MODE == 1) { for (int k = lane; k < (priv ? 32 * rstride : P.n0 * P.n1); k += 32) fhist[k] = 0.0; }
        else if (MODE == 2) { for (int k = lane; k < 64 * rstride; k += 32) hist[k] = 0; }
        else if (!wide) { for (int k = lane; k < (MODE == 5 ? 32 : 64) * rstride; k += 32) fhist[k] = 0.0; }
        tag[0] = tag[1] = 0;
        wa[0] = wb[0] = 0.0;
        x0 = y0 = z0 = x1 = y1 = z1 = 0.0;
        xs0 = ys0 = zs0 = xs1 = ys1 = zs1 = 0.0;
        __syncwarp();
    }
    __device__ __forceinline__ void tile_weight(unsigned) {}
    __device__ __forceinline__ void force_exact() {}
    __device__ __forceinline__ void tile_begin(const double (&p)[2][3], const bool (&val)[2], const uint32_t (&idx)[2], const WalkArrays &A)
    {
        // unused lanes of a partial tile carry the far sentinel in x: never inside the top edge
        x0 = p[0][0]; y0 = p[0][1]; z0 = p[0][2];
        x1 = p[1][0]; y1 = p[1][1]; z1 = p[1][2];
        vmask = (val[0] ? 1u : 0u) | (val[1] ? 2u : 0u);
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < HTB_MAX_NW; ++k) {
                wa[k] = (A.pay1 && k < A.nw) ? A.pay1[(size_t)idx[0] * A.nw + k] : 0.0;
                wb[k] = (A.pay1 && k < A.nw) ? A.pay1[(size_t)idx[1] * A.nw + k] : 0.0;
            }
        }
        if (JK) {
            wa[0] = A.pay1[(size_t)idx[0] * 2]; wb[0] = A.pay1[(size_t)idx[1] * 2];
            tag[0] = (int)A.pay1[(size_t)idx[0] * 2 + 1]; tag[1] = (int)A.pay1[(size_t)idx[1] * 2 + 1];
            utag = __shfl_sync(HTB_FULL, tag[0], 0);               // lane 0's first point is always valid
            uni = __all_sync(HTB_FULL, (!val[0] || tag[0] == utag) && (!val[1] || tag[1] == utag));
        }
    }
    __device__ __forceinline__ void set_shift(const double (&sh)[3], const WalkArrays &)
    {
        // npairs_3d_engine.pyx:167: the periodic shift is applied to the sample1 coordinate first
        if (REV) {
            // exchanged roles: the lanes keep their raw coordinates, the shift goes to the staged sample1 point
            xs0 = x0; ys0 = y0; zs0 = z0; xs1 = x1; ys1 = y1; zs1 = z1;
            rs0 = sh[0]; rs1 = sh[1]; rs2 = sh[2];
        } else {
            xs0 = x0 - sh[0]; ys0 = y0 - sh[1];
            xs1 = x1 - sh[0]; ys1 = y1 - sh[1];
            if (DIM == 3) { zs0 = z0 - sh[2]; zs1 = z1 - sh[2]; }
        }
        {
            __syncwarp();
            const uint32_t a = pts_s + (uint32_t)PT_BYTES * (uint32_t)lane, b = pts_s + (uint32_t)PT_BYTES * (uint32_t)(32 + lane);
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(xs0), "d"(ys0) : "memory");
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + 16u), "d"(zs0), "d"(wa[0]) : "memory");
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(b), "d"(xs1), "d"(ys1) : "memory");
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(b + 16u), "d"(zs1), "d"(wb[0]) : "memory");
            if (MODE == 1) {
#pragma unroll
                for (int k = 1; k < HTB_MAX_NW; ++k) {
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a + 24u + 8u * (uint32_t)k), "d"(wa[MODE == 1 ? k : 0]) : "memory");
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(b + 24u + 8u * (uint32_t)k), "d"(wb[MODE == 1 ? k : 0]) : "memory");
                }
            }
            __syncwarp();
        }
    }
    // the two separations the bins are defined on, in the reference's evaluation order
    __device__ __forceinline__ void seps(double xs, double ys, double zs, double xj, double yj, double zj, double &a, double &b)
    {
        // REV: this pass's shift is minus the reference's (the window is seen from the sample2 point), so the
        // reference's x1tmp = x1 - shift (npairs_jackknife_3d_engine.pyx:213-215) is xj + rs; dx only enters squared
        const double dx = REV ? (xj + rs0) - xs : xs - xj, dy = REV ? (yj + rs1) - ys : ys - yj;
        if (KIND == 3) {
            a = dx * dx + dy * dy;                      // weighted_npairs_xy_engine.pyx:163-165
            b = 0.0;
            return;
        }
        const double dz = REV ? (zj + rs2) - zs : zs - zj;
        if (KIND == 0) {
            a = dx * dx + dy * dy + dz * dz;            // npairs_3d_engine.pyx:176
            b = 0.0;
        } else if (KIND == 1) {
            a = dx * dx + dy * dy;                      // npairs_xy_z_engine.pyx:180-183
            b = dz * dz;
        } else {
            b = dx * dx + dy * dy;                      // npairs_s_mu_engine.pyx:196-200: dxy_sq, then sqr_s = dz_sq + dxy_sq
            const double dz_sq = dz * dz;
            a = dz_sq + b;
        }
    }
    __device__ __forceinline__ unsigned maybe(double xs, double ys, double zs, double xj, double yj, double zj)
    {
        double a, b;
        seps(xs, ys, zs, xj, yj, zj, a, b);
        if (KIND == 1) return (__double2hiint(a) <= P.H0 && __double2hiint(b) <= P.H1) ? 1u : 0u;
        return (__double2hiint(a) <= P.H0) ? 1u : 0u;
    }
    // first index i in [0, n] whose edge is >= the value with raw bits `bits` (n: above every edge): one table lookup on
    // the value's key; exact 64-bit compares only when an edge shares the value's table cell (the entry's low bit)
    __device__ __forceinline__ int locate(int axis, int first, int n, unsigned long long bits)
    {
        // key = bits >> S with S >= 44: a 32-bit shift of the high word.  Table entry 0 serves every key below the
        // smallest non-zero edge's, entry T - 1 every key above the top edge's (or a negative NaN: a huge unsigned key)
        const unsigned key = (unsigned)(bits >> 32) >> (P.S[axis] - 32);
        const int t = min((int)key - (int)P.kmin[axis] + 1, P.T[axis] - 1);            // keys are below 2^20
        const uint32_t la = lut_s[axis] + (uint32_t)max(t, 0);
        unsigned v, v2;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(la));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v2) : "r"(la + (t < P.T[axis] - 1 ? 1u : 0u)));
        int i = (int)(v >> 1);
        // edges inside the value's table cell: the next entry's index minus this one (0 unless the entry is flagged).
        // One edge (the usual case of a flagged cell) is settled by one unconditional compare; more take the loop.
        const int cnt = (v & 1u) ? (int)(v2 >> 1) - i : 0;
        const uint32_t eb = e_s + 8u * (uint32_t)first;
        const bool adv = (cnt >= 1) & (i < n) & (bq_lds_u64(eb + 8u * (uint32_t)min(i, n - 1)) < bits);
        i += adv ? 1 : 0;
        if (cnt >= 2 && adv) {
            while (i < n && bq_lds_u64(eb + 8u * (uint32_t)i) < bits) ++i;
        }
        return i;
    }
    // one replayed pair: recompute, locate; returns the histogram cell or -1
    __device__ __forceinline__ int replay_one(bool act, double xs, double ys, double zs, double xj, double yj, double zj)
    {
        double a, b;
        seps(xs, ys, zs, xj, yj, zj, a, b);
        const int i0 = locate(0, 0, P.n0, (unsigned long long)__double_as_longlong(a));
        int i1 = 0;
        if (KIND == 1) {
            i1 = locate(1, P.n0, P.n1, (unsigned long long)__double_as_longlong(b));
        } else if (KIND == 2) {
            // npairs_s_mu_engine.pyx:205-211
            const double sqr_mu = (a > 0.0) ? b / a : 0.0;
            i1 = locate(1, P.n0, P.n1, (unsigned long long)__double_as_longlong(sqr_mu));
        }
        return (act && i0 < P.n0 && i1 < P.n1) ? i0 * P.n1 + i1 : -1;
    }
    __device__ __forceinline__ double weight_of(const double *w1, uint32_t w2addr)
    {
        if (P.wfunc < 0) return lds_f64(w2addr);        // weighted_npairs_xy_engine.pyx:167: the weight is w2[j] alone
        double w2l[HTB_MAX_NW];
#pragma unroll
        for (int k = 0; k < HTB_MAX_NW; ++k) w2l[k] = (k < P.nw) ? lds_f64(w2addr + 8 * k) : 0.0;
        return htb_pair_weight(P.wfunc, w1, w2l);
    }
    // Balanced replay of one quarter chunk (16 staged points): the lanes' recorded pairs (a0 / a1: 16-bit masks of this
    // lane's two points) are laid out as ONE list in shared memory (warp prefix sum of the counts, entry = tile slot
    // << 4 | staged point), then every lane replays an equal share of the list - whatever lanes the pairs came from.
    // (A lane's pairs are all in range or all out of range of a staged run far more often than not: replaying them
    // in place keeps about 30 % of the lanes busy.)
    __device__ __forceinline__ void replay_quarter(uint32_t stage, unsigned a0, unsigned a1, int jbase)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH, bw = stage + 8 * DIM * HTB_CH;
        const int c = __popc(a0) + __popc(a1);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(HTB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const int total = __shfl_sync(HTB_FULL, inc, 31);
        uint32_t w = list_s + 2u * (uint32_t)(inc - c);
        for (unsigned m = a0; m; m &= m - 1u) {
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(w), "h"((unsigned short)((lane << 4) | (__ffs((int)m) - 1))) : "memory");
            w += 2u;
        }
        for (unsigned m = a1; m; m &= m - 1u) {
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(w), "h"((unsigned short)(((32 + lane) << 4) | (__ffs((int)m) - 1))) : "memory");
            w += 2u;
        }
        __syncwarp();
        const int per = (total + 31) >> 5;                       // every lane replays `per` consecutive entries
#pragma unroll 1
        for (int k = 0; k < per; ++k) {
            const int g = lane * per + k;
            const bool act = g < total;
            unsigned short e;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(e) : "r"(list_s + 2u * (uint32_t)min(g, total - 1)));
            const int slot = (int)(e >> 4), j = jbase + (int)(e & 15u);
            double px, py, pz, pw;
            lds_f64x2(pts_s + (uint32_t)PT_BYTES * (uint32_t)slot, px, py);
            lds_f64x2(pts_s + (uint32_t)PT_BYTES * (uint32_t)slot + 16u, pz, pw);
            const double xj = lds_f64(bx + 8 * j), yj = lds_f64(by + 8 * j), zj = DIM == 3 ? lds_f64(bz + 8 * j) : 0.0;
            const int h = replay_one(act, px, py, pz, xj, yj, zj);
            if (h >= 0) {
                if (MODE == 0) atomicAdd(hist + h, 1u);
                else if (MODE == 1) {
                    double w1l[HTB_MAX_NW];
                    w1l[0] = pw;
#pragma unroll
                    for (int q = 1; q < HTB_MAX_NW; ++q) w1l[q] = (q < P.nw) ? lds_f64(pts_s + (uint32_t)PT_BYTES * (uint32_t)slot + 24u + 8u * (uint32_t)q) : 0.0;
                    const double w = weight_of(w1l, bw + 8 * j * P.nw);
                    if (priv) fhist[lane * rstride + h] += w;
                    else atomicAdd(fhist + h, w);
                }
                else if (MODE == 2) atomicAdd(hist + slot * rstride + h, 1u);
                else if (JK) {
                    const double w = REV ? lds_f64(bw + 16 * j) * pw : pw * lds_f64(bw + 16 * j);                     // jweight's w1 * w2
                    // one tag for the whole tile (the rule: sub-volumes are spatial): one histogram serves every point -
                    // shared atomics for wide tables, else the replaying lane's private row (summed over the warp at the
                    // end of the tile); mixed tiles keep a row per point
                    if (!uni) atomicAdd(fhist + slot * rstride + h, w);
                    else if (wide) atomicAdd(shist + h, w);
                    else fhist[lane * rstride + h] += w;
                }
                else if (MODE == 5) fhist[lane * rstride + h] += lds_f64(bw + 8 * j);                     // the weight is w2[j]
                else atomicAdd(fhist + slot * rstride + h, lds_f64(bw + 8 * j));
            }
        }
        __syncwarp();                                            // the list is rewritten by the next quarter
    }
    __device__ __forceinline__ void chunk(uint32_t stage, int lo, int hi, uint32_t tok)
    {
        const uint32_t bx = stage, by = stage + 8 * HTB_CH, bz = stage + 16 * HTB_CH;
        // groups of 4 staged points, aligned; entries outside [lo, hi) are evaluated on whatever the stage holds
        // and masked out afterwards
        const int jb = lo & ~3, je = (hi + 3) & ~3;
        uint32_t m[2][2] = {{0u, 0u}, {0u, 0u}};
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int b0 = max(jb, 32 * w), b1 = min(je, 32 * w + 32);
            uint32_t a0 = 0u, a1 = 0u;
#pragma unroll 2
            for (int j = b0; j < b1; j += 4) {
                double xa, xb, xc, xd, ya, yb, yc, yd, za = 0.0, zb = 0.0, zc = 0.0, zd = 0.0;
                bq_lds_f64x2_tok(bx + 8 * j, tok, xa, xb);
                bq_lds_f64x2_tok(by + 8 * j, tok, ya, yb);
                if (DIM == 3) bq_lds_f64x2_tok(bz + 8 * j, tok, za, zb);
                bq_lds_f64x2_tok(bx + 8 * j + 16, tok, xc, xd);
                bq_lds_f64x2_tok(by + 8 * j + 16, tok, yc, yd);
                if (DIM == 3) bq_lds_f64x2_tok(bz + 8 * j + 16, tok, zc, zd);
                unsigned n0 = maybe(xs0, ys0, zs0, xa, ya, za);
                unsigned n1 = maybe(xs1, ys1, zs1, xa, ya, za);
                n0 |= maybe(xs0, ys0, zs0, xb, yb, zb) << 1;
                n1 |= maybe(xs1, ys1, zs1, xb, yb, zb) << 1;
                n0 |= maybe(xs0, ys0, zs0, xc, yc, zc) << 2;
                n1 |= maybe(xs1, ys1, zs1, xc, yc, zc) << 2;
                n0 |= maybe(xs0, ys0, zs0, xd, yd, zd) << 3;
                n1 |= maybe(xs1, ys1, zs1, xd, yd, zd) << 3;
                a0 |= n0 << (j & 31);
                a1 |= n1 << (j & 31);
            }
            m[0][w] = a0; m[1][w] = a1;
        }
        // the two 32-point halves of the chunk are replayed one after the other (32-bit masks)
        const unsigned long long range = (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
        {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const unsigned rw = (unsigned)(range >> (32 * w));
                const unsigned f0 = m[0][w] & rw, f1 = m[1][w] & rw;
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                    const unsigned a0 = (f0 >> (16 * hq)) & 0xffffu, a1 = (f1 >> (16 * hq)) & 0xffffu;
                    if (__any_sync(HTB_FULL, (a0 | a1) != 0u)) replay_quarter(stage, a0, a1, 32 * w + 16 * hq);
                }
            }
        }
    }
    __device__ __forceinline__ bool tile_end(const WalkArrays &, const uint32_t (&idx)[2], int, unsigned wt)
    {
        __syncwarp();
        const int nh = P.n0 * P.n1;
        if (MODE == 0) {
            for (int k = lane; k < nh; k += 32) {
                const uint32_t h = hist[k];
                if (h) { atomicAdd(P.counts + k, (unsigned long long)wt * h); hist[k] = 0; }
            }
        } else if (MODE == 1 && !priv) {
            for (int k = lane; k < nh; k += 32) {
                const double h = fhist[k];
                if (h != 0.0) { atomicAdd(P.fcounts + k, wt == 2u ? h + h : h); fhist[k] = 0.0; }
            }
        } else if (MODE == 1 || MODE == 5) {
            // one row for the whole call: sum the lanes' rows (differential cells), one atomic per cell and tile
            for (int k = 0; k < nh; ++k) {
                double *r = fhist + lane * rstride;
                double v = r[k];
                r[k] = 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HTB_FULL, v, o);
                if (lane == 0 && v != 0.0) atomicAdd(P.fcounts + k, wt == 2u ? v + v : v);
            }
        } else if (MODE == 4) {
            // per object: cumulative over the edges, rows in input order (weighted_npairs_per_object_xy_engine.pyx:175-185)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (!((vmask >> q) & 1u)) continue;
                double *row = P.fcounts + (size_t)P.perm1[idx[q]] * (size_t)nh;
                double *r = fhist + (q * 32 + lane) * rstride;
                double cum = 0.0;
                for (int k = 0; k < nh; ++k) {
                    cum += r[k];
                    r[k] = 0.0;
                    if (cum != 0.0) atomicAdd(row + k, cum);
                }
            }
        } else if (JK && uni) {
            double *dst = P.fcounts + (size_t)utag * (size_t)nh;
            if (wide) {
                for (int k = lane; k < nh; k += 32) {
                    const double x = shist[k];
                    if (x != 0.0) { atomicAdd(dst + k, x); shist[k] = 0.0; }
                }
            } else {
                double *r = fhist + lane * rstride;
                for (int k = 0; k < nh; ++k) {
                    double x = r[k];
                    r[k] = 0.0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(HTB_FULL, x, o);
                    if (lane == 0 && x != 0.0) atomicAdd(dst + k, x);
                }
            }
        } else if (JK) {
            // fold the rows of this lane's points into the table row of their jackknife tag (differential cells)
            // (sub-volumes are spatial: the points of a tile nearly always share one tag - then the rows are summed
            // over the warp first and one lane adds them)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const bool v = ((vmask >> q) & 1u) != 0u;
                const int tref = __shfl_sync(HTB_FULL, tag[q], q ? 31 : 0);      // slot 32 q lives in lane 0 (q = 0) / 31 (q = 1)
                double *r = fhist + (q * 32 + lane) * rstride;
                if (__all_sync(HTB_FULL, !v || tag[q] == tref)) {
                    if (!__any_sync(HTB_FULL, v)) continue;
                    double *dst = P.fcounts + (size_t)tref * (size_t)nh;
                    for (int k = 0; k < nh; ++k) {
                        double x = v ? r[k] : 0.0;
                        r[k] = 0.0;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(HTB_FULL, x, o);
                        if (lane == 0 && x != 0.0) atomicAdd(dst + k, x);
                    }
                } else if (v) {
                    double *dst = P.fcounts + (size_t)tag[q] * (size_t)nh;
                    for (int k = 0; k < nh; ++k) {
                        const double x = r[k];
                        if (x != 0.0) { atomicAdd(dst + k, x); r[k] = 0.0; }
                    }
                }
            }
        } else {
            // per object: cumulative over the edges (npairs_per_object_3d_engine.pyx:190-207), rows in input order;
            // several work items (column slices) may add to the same row
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (!((vmask >> q) & 1u)) continue;
                unsigned long long *row = P.rows + (size_t)P.perm1[idx[q]] * (size_t)P.n0;
                uint32_t *r = hist + (q * 32 + lane) * rstride;
                unsigned long long cum = 0ull;
                for (int k = 0; k < P.n0; ++k) {
                    cum += r[k];
                    r[k] = 0u;
                    if (cum) atomicAdd(row + k, cum);
                }
            }
        }
        __syncwarp();
        return false;
    }
    __device__ __forceinline__ void kernel_end() {}
};

int htb_launch_binq(cudaStream_t st, int kind, int mode, const WalkGeom &G, const WalkArrays &A, const BinQParams &P, int *l)
{
    switch (mode * 4 + kind) {
    case 0: return launch_count<BinQ<0, 0>>(st, G, A, P, l);
    case 1: return launch_count<BinQ<1, 0>>(st, G, A, P, l);
    case 2: return launch_count<BinQ<2, 0>>(st, G, A, P, l);
    case 4: return launch_count<BinQ<0, 1>>(st, G, A, P, l);
    case 5: return launch_count<BinQ<1, 1>>(st, G, A, P, l);
    case 7: return launch_count<BinQ<3, 1>>(st, G, A, P, l);
    case 8: return launch_count<BinQ<0, 2>>(st, G, A, P, l);
    case 12: return launch_count<BinQ<0, 3>>(st, G, A, P, l);
    case 13: return launch_count<BinQ<1, 3>>(st, G, A, P, l);
    case 24: return launch_count<BinQ<0, 6>>(st, G, A, P, l);
    case 25: return launch_count<BinQ<1, 6>>(st, G, A, P, l);
    case 19: return launch_count<BinQ<3, 4>>(st, G, A, P, l);
    case 23: return launch_count<BinQ<3, 5>>(st, G, A, P, l);
    }
    htb_set_error("unknown BinQ kind %d / mode %d", kind, mode);
    return 1;
}
