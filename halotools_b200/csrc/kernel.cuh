// The counting-kernel skeleton k_count<V> (persistent warps pulling work items, redo queue) and its host launcher,
// shared by the translation units that define variant functors (count.cu, binq.cu).
#pragma once
#include <type_traits>
#include "walk.cuh"
#include "count.cuh"

// ------------------------------------------------------------------ kernel skeleton
template <class V, class = void> struct HtbPerCell { static constexpr bool value = false; };
template <class V> struct HtbPerCell<V, std::void_t<decltype(V::PER_CELL)>> { static constexpr bool value = V::PER_CELL; };

__device__ __forceinline__ unsigned long long htb_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <class V>
__global__ void __launch_bounds__(V::WARPS * 32, V::MINBLOCKS)
k_count(const __grid_constant__ WalkGeom G, const __grid_constant__ WalkArrays A,
        const __grid_constant__ typename V::Params P, const int scratch_bytes_per_warp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int DIM = V::DIM;
    constexpr int F = DIM - 1;
    constexpr int PPL = V::PPL;                 // sample1 points per lane; a tile holds 32 * PPL points
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    typedef WarpSmem<DIM, V::NPAY, HtbChunk<V>::value> WS;
    unsigned char *mine = smem_raw + (size_t)warp * (WS::bytes() + (size_t)scratch_bytes_per_warp);
    WS S;
    S.stage0 = (double *)mine;
    S.stage0_s = smem_u32(mine);
    unsigned char *bars = mine + sizeof(double) * HTB_NSTAGE * WS::stage_doubles();
    S.bar0 = smem_u32(bars);
    S.span = (uint32_t *)(bars + 16 * HTB_NSTAGE);
    void *scratch = mine + WS::bytes();
    if (V::TMA) {
        if (lane == 0) {
            for (int s = 0; s < HTB_NSTAGE; ++s) mbar_init(S.bar(s), 1);
            mbar_fence_init();
        }
    }
    __syncwarp();

    V v(P, scratch, lane, A);
    uint32_t gchunk = 0;
    unsigned long long pairs = 0;
    unsigned int redone = 0;
    const int ntiles = A.ntiles_dev[0];
    // Work items: when there are too few tiles to keep every resident warp busy to the end (small samples, one
    // rank's shard of a multi-GPU count, clustered sample1), every tile is cut into K SLICES of its sample2
    // columns; slices of one tile are independent work items (counts and per-object sums are additive).
    // Two phases: the first nA tiles are whole work items, the LAST nB tiles (G.tail_eighths / 8 tiles per resident warp)
    // are cut into Kmax slices each, so that the warps run out of work within a fraction of a tile of each other instead
    // of up to a whole tile (one tile of the bench's RR count is 2 ms of one warp; a rank of an 8-GPU run has four tiles
    // per warp).  With fewer than two tiles per resident warp every tile is sliced (KA) as far as it pays.
    const int Kmax = G.sym ? min(G.maxslices, 8) : G.maxslices;
    const long long W = (long long)gridDim.x * V::WARPS;
    int KA = 1, KB = 1, nB = 0;
    if (ntiles >= 2 * W && Kmax > 1 && G.tail_eighths > 0) {
        KB = Kmax;
        nB = (int)min((long long)ntiles, W * G.tail_eighths / 8);
    } else {
        const long long target = W * G.items_per_warp;
        if (ntiles > 0 && ntiles < target) KA = (int)min((long long)Kmax, (target + ntiles - 1) / ntiles);
        if (KA < 1) KA = 1;
        KB = KA;
    }
    const int nA = ntiles - nB;
    const int itemsA = nA * KA;
    const int nitems = itemsA + nB * KB;

    // HTB_FLAG_EARLY_EXIT (several count kernels of a statistic in flight on different streams): a launch whose items
    // do not divide evenly over the resident warps ends with a round in which most warps idle - e.g. 4.2 tiles per warp
    // on one rank of 8 take the time of 5.  The surplus blocks give their SM slots away AT ONCE instead: with `rounds`
    // = ceil(items / warps), ceil(items / (rounds - 1/4)) warps finish in the same number of rounds, and the blocks beyond
    // them retire before they start, so the next kernel's blocks (and its set-up kernels) run beside this one for its whole
    // length rather than squeezing into its tail.
    if (G.early_exit && nitems > 0 && V::WARPS * (long long)gridDim.x > 0) {
        const long long Wt = (long long)gridDim.x * V::WARPS;
        const long long rounds = (nitems + Wt - 1) / Wt;
        const long long needed = (4LL * nitems + (4 * rounds - 1) - 1) / (4 * rounds - 1);
        const long long blocks_needed = (needed + V::WARPS - 1) / V::WARPS;
        if ((long long)blockIdx.x >= blocks_needed) return;
    }

    // Device-side time stamps of the launch's own execution (first warp in, last warp out; words 40..43 of the counter
    // block, zeroed by the host): with several count kernels of a statistic in flight on different streams a CUDA-event
    // bracket around this launch also measures the time its blocks WAIT for the other kernels' blocks to retire.
    if (lane == 0) atomicMax((unsigned long long *)(A.tile_counter + 40), ~htb_globaltimer());

    // One work item: slice `slice` of `nsl` of tile t.  redo_sub < 0: the normal evaluation (both weight passes in
    // symmetric mode).  A fast kernel that finds it cannot decide a tile from its 32-bit keys asks for an exact
    // re-evaluation (tile_end returns true): that re-evaluation is many times slower per pair, so it is not done
    // in place but cut into HTB_REDO_SPLIT finer slices that go to a second queue served by every warp that runs
    // out of ordinary items - otherwise a single late redo is the tail of the whole launch.
    // redo_sub >= 0: such an exact re-evaluation of weight pass redo_sub.
    bool main_done = false;
    while (true) {
        // ---- next work item: the ordinary queue first, then the redo queue (entries published by any warp; done
        // when every ordinary item has completed and the queue is drained: reserved == published == taken)
        int t = 0, slice = 0, nsl = KA, redo_sub = -1;
        if (!main_done) {
            if (lane == 0) t = (int)atomicAdd(A.tile_counter, 1u);
            t = __shfl_sync(HTB_FULL, t, 0);
            if (t >= nitems) main_done = true;
            else if (t < itemsA) { slice = t % KA; t /= KA; }
            else { const int u = t - itemsA; slice = u % KB; t = nA + u / KB; nsl = KB; }
        }
        if (main_done) {
            if (A.redo_cap < HTB_REDO_SPLIT) break;
            unsigned got = 0xffffffffu;
            int quit = 0;
            if (lane == 0) {
                volatile unsigned *c = A.redo_ctr;
                unsigned backoff = 500;
                while (true) {
                    const unsigned done = c[3];               // read first: a finished item has published its entries
                    __threadfence();
                    const unsigned reserved = c[0], published = c[1], taken = c[2];
                    if (taken < published) {
                        if (atomicCAS(A.redo_ctr + 2, taken, taken + 1u) == taken) { got = taken; break; }
                        continue;
                    }
                    // nothing to take: finished only if no ordinary item is still running (it could publish more)
                    // and every reservation has been published
                    if (done >= (unsigned)nitems && published == reserved) { quit = 1; break; }
                    // HTB_FLAG_EARLY_EXIT: leave as soon as there is nothing to take, so that the block can retire and the
                    // blocks of the next kernel (waiting on another stream) fill this launch's tail.  Every published entry
                    // is still served: its publisher comes through this loop after its own items.
                    if (G.early_exit) { quit = 1; break; }
                    __nanosleep(backoff);
                    if (backoff < 8000) backoff *= 2;
                }
            }
            quit = __shfl_sync(HTB_FULL, quit, 0);
            if (quit) break;
            got = __shfl_sync(HTB_FULL, got, 0);
            __threadfence();
            const uint2 e = A.redo_ent[got];
            t = (int)e.x; slice = (int)(e.y & 0xffffffu); nsl = (t < nA ? KA : KB) * HTB_REDO_SPLIT; redo_sub = (int)(e.y >> 24);
        }
        const uint2 td = A.tiles[t];
        const uint32_t start = td.x;
        const int cnt = (int)(td.y >> 24);
        int64_t slow = (int64_t)(td.y & 0xffffffu);
        int fs[3] = {0, 0, 0};
#pragma unroll
        for (int d = F - 1; d >= 0; --d) { fs[d] = (int)(slow % G.nf1[d]); slow /= G.nf1[d]; }
        // this lane's points (index clamped for the bounding box; unused lanes get a far sentinel)
        bool val[PPL];
        uint32_t idx[PPL];
        double p[PPL][3], blo[3] = {0, 0, 0}, bhi[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            // odd point slots run backwards, so a lane holds points from both ends of the tile (sorted along the
            // fast dimension): its share of in-range pairs, hence its queue length, follows the warp average
            const int slot = (q & 1) ? 32 * q + 31 - lane : 32 * q + lane;
            val[q] = slot < cnt;
            idx[q] = start + (uint32_t)min(slot, cnt - 1);
            p[q][0] = p[q][1] = p[q][2] = 0.0;
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double lo = 0.0, hi = 0.0;
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                p[q][d] = A.c1[d][idx[q]];
                lo = q ? fmin(lo, p[q][d]) : p[q][d];
                hi = q ? fmax(hi, p[q][d]) : p[q][d];
            }
            blo[d] = warp_min(lo);
            bhi[d] = warp_max(hi);
        }
#pragma unroll
        for (int q = 0; q < PPL; ++q) if (!val[q]) p[q][0] = G.sentinel;
        // reference cells (fast dimension) of the tile's first and last point: the digitisation of the mesh sort
        fs[F] = htb_ref_digitize(__shfl_sync(HTB_FULL, p[0][F], 0), G.cs1f, G.nd1[F]);
        const int nref = htb_ref_digitize(__shfl_sync(HTB_FULL, p[PPL - 1][F], (PPL & 1) ? 31 : 0), G.cs1f, G.nd1[F]) - fs[F] + 1;
        const int nsub = G.sym ? 2 : 1;
#pragma unroll 1
        for (int sub = (redo_sub < 0 ? 0 : redo_sub); sub < (redo_sub < 0 ? nsub : redo_sub + 1); ++sub) {
            v.tile_begin(p, val, idx, A);
            v.tile_weight(sub == 1 ? 2u : 1u);
            if (redo_sub >= 0) v.force_exact();
#pragma unroll 1
            for (int pass = (redo_sub < 0 ? 0 : 1); pass < 2; ++pass) {
                if constexpr (HtbPerCell<V>::value)
                    walk_tile_cells<V>(v, G, A, S, gchunk, blo, bhi, fs, nref, pairs, cnt, slice, nsl);
                else
                    walk_tile<V>(v, G, A, S, gchunk, blo, bhi, fs, nref, pairs, cnt, G.sym ? sub + 1 : 0, start, start + (uint32_t)cnt,
                                 slice, nsl);
                const bool redo = v.tile_end(A, idx, pass, sub == 1 ? 2u : 1u);
                if (!redo) break;
                ++redone;
                // hand the exact re-evaluation to the redo queue (if it has room), else do it here
                unsigned pos = 0xffffffffu;
                if (A.redo_cap >= HTB_REDO_SPLIT) {
                    if (lane == 0) {
                        unsigned old = *(volatile unsigned *)(A.redo_ctr + 0);                 // reserve (never past the end)
                        while (old + HTB_REDO_SPLIT <= A.redo_cap) {
                            const unsigned seen = atomicCAS(A.redo_ctr + 0, old, old + HTB_REDO_SPLIT);
                            if (seen == old) { pos = old; break; }
                            old = seen;
                        }
                    }
                    pos = __shfl_sync(HTB_FULL, pos, 0);
                }
                if (pos == 0xffffffffu) continue;                                              // in place (pass 1)
                if (lane < HTB_REDO_SPLIT)
                    A.redo_ent[pos + lane] = make_uint2((unsigned)t, ((unsigned)sub << 24) | (unsigned)(slice * HTB_REDO_SPLIT + lane));
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(A.redo_ctr + 1, (unsigned)HTB_REDO_SPLIT);            // publish
                break;
            }
        }
        __syncwarp();
        if (redo_sub < 0 && lane == 0) { __threadfence(); atomicAdd(A.redo_ctr + 3, 1u); }     // ordinary items completed
    }
    v.kernel_end();
    if (lane == 0) {
        if (pairs) atomicAdd(A.pairs_evaluated, pairs);
        if (redone) atomicAdd(A.tiles_redone, redone);
        atomicMax((unsigned long long *)(A.tile_counter + 42), htb_globaltimer());
    }
}

// ------------------------------------------------------------------ pair weights of the marked counters
__device__ __forceinline__ double htb_pair_weight(int id, const double *w1, const double *w2)
{
    // marking_functions.pyx:14-217 (id 8 keeps the reference's '+', :106), custom_marking_func.pyx:12-16
    double d;
    switch (id) {
    case 0: case 1: return w1[0] * w2[0];
    case 2:  return w1[0] + w2[0];
    case 3:  return (w1[0] == w2[0]) ? w1[1] * w2[1] : 0.0;
    case 4:  return (w1[0] != w2[0]) ? w1[1] * w2[1] : 0.0;
    case 5:  return (w2[0] > w1[0]) ? w1[1] * w2[1] : 0.0;
    case 6:  return (w2[0] < w1[0]) ? w1[1] * w2[1] : 0.0;
    case 7:  return (w2[0] > (w1[0] + w1[1])) ? w2[1] : 0.0;
    case 8:  return (w2[0] < (w1[0] + w1[1])) ? w2[1] : 0.0;
    case 9:  return (fabs(w1[0] - w2[0]) < w1[1]) ? w2[1] : 0.0;
    case 10: return (fabs(w1[0] - w2[0]) > w1[1]) ? w2[1] : 0.0;
    case 11: return (w2[0] > w1[0] * w1[1]) ? w2[1] : 0.0;
    case 12: return w1[0] * w2[0] * (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]);
    case 13: d = (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]); return w1[0] * w2[0] * d * d;
    case 14: return w1[0] * w2[0] * (w1[1] * w2[1] + w1[2] * w2[2]);
    case 15: d = (w1[1] * w2[1] + w1[2] * w2[2]); return w1[0] * w2[0] * d * d;
    case 16: d = (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]);
             return (w1[4] == w2[4]) ? w1[0] * w2[0] * d * d : 0.0;
    case 17: d = (w1[1] * w2[1] + w1[2] * w2[2] + w1[3] * w2[3]);
             return (w1[4] != w2[4]) ? w1[0] * w2[0] * d * d : 0.0;
    default: return 0.0;
    }
}

// ------------------------------------------------------------------ host launcher
template <class V>
static int launch_count(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const typename V::Params &P,
                        int *launches)
{
    typedef WarpSmem<V::DIM, V::NPAY, HtbChunk<V>::value> WS;
    const size_t scratch = (V::scratch_bytes(P) + 15) & ~(size_t)15;
    const size_t smem = V::WARPS * (WS::bytes() + scratch);
    if (smem > 200 * 1024) {
        htb_set_error("too many bins for the shared-memory accumulators (%zu bytes of shared memory per block needed)", smem);
        return 1;
    }
    // the function attribute and the occupancy query cost ~10 us each: done once per (variant, device, shared-memory size)
    struct Cached { size_t smem; int sms, per_sm; };
    static Cached cache[64] = {};
    int dev = 0;
    HTB_CUDA(cudaGetDevice(&dev));
    Cached &cc = cache[dev & 63];
    if (cc.per_sm == 0 || cc.smem != smem) {
        int sms = 0, per_sm = 0;
        HTB_CUDA(cudaFuncSetAttribute(k_count<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        HTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_count<V>, V::WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        cc.smem = smem; cc.sms = sms; cc.per_sm = per_sm;
    }
    const int sms = cc.sms, per_sm = cc.per_sm;
    k_count<V><<<sms * per_sm, V::WARPS * 32, smem, st>>>(G, A, P, (int)scratch);
    if (launches) *launches += 1;
    HTB_CUDA(cudaGetLastError());
    return 0;
}
