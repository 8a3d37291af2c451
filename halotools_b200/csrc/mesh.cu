// K1 — spatial index on the GPU: reference cell assignment + counting sort into
// contiguous SoA float64 per (refined) cell.
//
// Replaces RectangularMesh.__init__ (digitize -> argsort -> searchsorted,
// /root/reference/halotools/mock_observables/pair_counters/rectangular_mesh.py:118-222)
// by: k_assign (fine cell id of every point + points per cell through fire-and-forget atomics: RED, no round trip), an
// exclusive scan of the per-cell counts (one block for small meshes, three passes otherwise; it zeroes the counters on the
// way, which then serve as fill counters), and k_scatter (position = first position of the cell + arrival rank from one
// returning atomic on the cell's fill counter; it also writes the pad entries).  O(N) HBM traffic: read 24 B + write 4 B
// per point in k_assign, read 28 B + write 24 B (+ payload) in k_scatter.  1e8 particles (2-D): 5.9 ms = 0.73 ms assign
// (FP64 bound until the fmod-based floor division was replaced by an exact floor from a reciprocal estimate and one fma,
// htb_ref_digitize) + 4.7 ms scatter (bound by the 8-byte stores into random 32-byte sectors: 7.3 GB of DRAM traffic for
// 3.6 GB of algorithmic bytes; profiles/r02_k1_ncu.txt).  Round 1: 10.4 ms with the arrival rank taken and stored in
// k_assign.  A first two-pass variant (block-local partition into <= 256 buckets of consecutive cells, then an
// L2-resident scatter per bucket) measured 10.7 ms.
#include <cstdlib>
#include "htb_internal.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

template <int DIM>
__global__ void __launch_bounds__(256)
k_assign(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
         int64_t stride, int64_t n, FineGrid g, uint32_t *__restrict__ cell,
         uint32_t *__restrict__ count, uint32_t *__restrict__ flags)
{
    const double *src[3] = {x, y, z};
    uint32_t bad = 0;
    double rc[DIM], mc[DIM];                          // the two divisions per dimension, once per thread
#pragma unroll
    for (int d = 0; d < DIM; ++d) { rc[d] = 1.0 / g.cs[d]; mc[d] = (double)g.m[d] / g.cs[d]; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t cid = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            const double p = src[d][i * stride];
            const int r = htb_ref_digitize(p, g.cs[d], rc[d], g.nd[d]);
            int f = r;
            if (g.m[d] > 1) {
                int sub = (int)floor((p - (double)r * g.cs[d]) * mc[d]);
                sub = sub < 0 ? 0 : (sub >= g.m[d] ? g.m[d] - 1 : sub);
                f = r * g.m[d] + sub;
            }
            if (!(p >= 0.0 && p <= g.period[d])) bad = 1;
            cid = cid * (uint32_t)g.nf[d] + (uint32_t)f;
        }
        cell[i] = cid;
        atomicAdd(&count[cid], 1u);                   // result unused: a RED, no round trip
    }
    if (bad) atomicOr(flags, 1u);
}

#define PART_THREADS 256
#define PART_ITEMS 8
#define PART_CHUNK (PART_THREADS * PART_ITEMS)

// every point to its cell: pos = first position of the cell + arrival rank (one returning atomic on the cell's fill counter)
template <int DIM>
__global__ void __launch_bounds__(PART_THREADS)
k_scatter(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z, int64_t stride, int64_t n,
          const uint32_t *__restrict__ cell, const uint32_t *__restrict__ off, uint32_t *__restrict__ fill,
          double *__restrict__ ox, double *__restrict__ oy, double *__restrict__ oz,
          uint32_t *__restrict__ perm, const double *__restrict__ w, double *__restrict__ ow, int nw,
          int64_t npad, double pad_value)
{
    // entries [n, npad) get a far-away sentinel, so that staging reads past a span end are harmless
    if (blockIdx.x == 0 && n + threadIdx.x < npad) {
        ox[n + threadIdx.x] = pad_value;
        oy[n + threadIdx.x] = pad_value;
        if (DIM == 3) oz[n + threadIdx.x] = pad_value;
    }
    // the PART_ITEMS points of a thread go through the dependent chain (cell id -> first position -> returning atomic ->
    // stores) side by side, phase by phase, so that a thread has PART_ITEMS gathers / atomics in flight instead of one
    const int64_t nchunk = (n + PART_CHUNK - 1) / PART_CHUNK;
    for (int64_t c = blockIdx.x; c < nchunk; c += gridDim.x) {
        const int64_t i0 = c * PART_CHUNK + threadIdx.x;
        uint32_t cid[PART_ITEMS], pos[PART_ITEMS];
        double px[PART_ITEMS], py[PART_ITEMS], pz[PART_ITEMS];
        bool ok[PART_ITEMS];
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k) {
            const int64_t i = i0 + (int64_t)k * PART_THREADS;
            ok[k] = i < n;
            cid[k] = ok[k] ? cell[i] : 0u;
            px[k] = ok[k] ? x[i * stride] : 0.0;
            py[k] = ok[k] ? y[i * stride] : 0.0;
            pz[k] = (DIM == 3 && ok[k]) ? z[i * stride] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k) {
            const uint32_t first = ok[k] ? off[cid[k]] : 0u, next = ok[k] ? off[cid[k] + 1] : 0u;
            ok[k] = next != first;                        // an emptied cell: outside this rank's window (htb_sort_finish)
            pos[k] = first;
        }
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k)
            if (ok[k]) pos[k] += atomicAdd(&fill[cid[k]], 1u);
#pragma unroll
        for (int k = 0; k < PART_ITEMS; ++k) {
            if (!ok[k]) continue;
            ox[pos[k]] = px[k];
            oy[pos[k]] = py[k];
            if (DIM == 3) oz[pos[k]] = pz[k];
            if (perm) perm[pos[k]] = (uint32_t)(i0 + (int64_t)k * PART_THREADS);
            if (ow) {
                const int64_t i = i0 + (int64_t)k * PART_THREADS;
                for (int q = 0; q < nw; ++q) ow[(int64_t)pos[k] * nw + q] = w[i * nw + q];
            }
        }
    }
}

// pad entries [n, npad) with a far-away sentinel so that staging reads past a span end are harmless
__global__ void k_pad(double *a, double *b, double *c, int64_t n, int64_t npad, double value)
{
    int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) {
        a[i] = value;
        b[i] = value;
        if (c) c[i] = value;
    }
}

// ------------------------------------------------------------------ exclusive scan
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_partial(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ bsum)
{
    __shared__ uint32_t red[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += red[w];
        bsum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of the block sums in place, total to *total.  256 threads: single-block kernels must fit
// beside the resident blocks of a persistent count kernel of another stream (see k_shard_range)
#define SCAN_ONE 256
__global__ void __launch_bounds__(SCAN_ONE)
k_scan_bsums(uint32_t *__restrict__ bsum, int64_t nb, uint32_t *__restrict__ total)
{
    __shared__ uint32_t sh[SCAN_ONE];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += SCAN_ONE) {
        int64_t i = base + threadIdx.x;
        uint32_t v = (i < nb) ? bsum[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_ONE; o <<= 1) {
            uint32_t t = (threadIdx.x >= (unsigned)o) ? sh[threadIdx.x - o] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        const uint32_t incl = sh[threadIdx.x];
        const uint32_t c = carry;
        if (i < nb) bsum[i] = c + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_ONE - 1) carry = c + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_final(uint32_t *__restrict__ in, int64_t n, const uint32_t *__restrict__ bsum,
             uint32_t *__restrict__ out, int zero_in)
{
    // block-local exclusive scan of SCAN_BLOCK items laid out [k][thread] is awkward; use a
    // thread-contiguous layout instead: thread t owns items [t*ITEMS, (t+1)*ITEMS).
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    if (zero_in) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) in[base + k] = 0u;
    }
    // warp inclusive scan of s
    uint32_t inc = s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += wsum[w];
    uint32_t run = bsum[blockIdx.x] + wbase + inc - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void k_set_u32(uint32_t *dst, const uint32_t *src) { *dst = *src; }

// one launch instead of three for small inputs (a 1e5-point call spends its time between kernels, not in them): one
// block (256 threads, see k_scan_bsums) walks the array in chunks of 256 x SCAN_SMALL_ITEMS, warp-shuffle scan inside a
// chunk, running carry across
#define SCAN_SMALL_ITEMS 16
#define SCAN_SMALL_THREADS 256
#define SCAN_SMALL_MAX (1 << 15)
__global__ void __launch_bounds__(SCAN_SMALL_THREADS)
k_scan_small(uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ out, uint32_t *__restrict__ total, int zero_in)
{
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int64_t base = 0; base < n; base += SCAN_SMALL_THREADS * SCAN_SMALL_ITEMS) {
        const int64_t i0 = base + (int64_t)threadIdx.x * SCAN_SMALL_ITEMS;
        uint32_t v[SCAN_SMALL_ITEMS], s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_SMALL_ITEMS; ++k) { v[k] = (i0 + k < n) ? in[i0 + k] : 0u; s += v[k]; }
        if (zero_in) {
#pragma unroll
            for (int k = 0; k < SCAN_SMALL_ITEMS; ++k) if (i0 + k < n) in[i0 + k] = 0u;
        }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = lane < SCAN_SMALL_THREADS / 32 ? wsum[lane] : 0u, winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            wsum[lane] = winc - w;                       // exclusive warp bases
            if (lane == 31) carry_s = winc;              // chunk total
        }
        __syncthreads();
        uint32_t run = carry + wsum[wid] + inc - s;
#pragma unroll
        for (int k = 0; k < SCAN_SMALL_ITEMS; ++k) {
            if (i0 + k < n) out[i0 + k] = run;
            run += v[k];
        }
        carry += carry_s;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// zero_in: the input is zeroed once it has been read (the sort reuses the scanned counters as fill counters)
int htb_exclusive_scan_u32(cudaStream_t st, Workspace &ws, uint32_t *in, uint32_t *out,
                           int64_t n, uint32_t *total_dev, int *launches, bool zero_in)
{
    if (n <= 0) {
        if (total_dev) HTB_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), st));
        return 0;
    }
    if (n <= SCAN_SMALL_MAX) {
        k_scan_small<<<1, SCAN_SMALL_THREADS, 0, st>>>(in, n, out, total_dev, zero_in ? 1 : 0);
        if (launches) *launches += 1;
        HTB_CUDA(cudaGetLastError());
        return 0;
    }
    const int64_t nblk = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    uint32_t *bsum = nullptr;
    if (ws.alloc((void **)&bsum, sizeof(uint32_t) * (size_t)nblk)) return 1;
    k_scan_partial<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(in, n, bsum);
    k_scan_bsums<<<1, SCAN_ONE, 0, st>>>(bsum, nblk, total_dev);
    k_scan_final<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(in, n, bsum, out, zero_in ? 1 : 0);
    if (launches) *launches += 3;
    HTB_CUDA(cudaGetLastError());
    return 0;
}

static int grid_for(int64_t n, int threads)
{
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// fine cells whose reference index along dimension 0 is outside the window {first layer (may be negative), layers}
// lose their points: a rank that counts only some x-layers of mesh1 cells needs the points of its own window only
__global__ void k_mask_counts(uint32_t *__restrict__ count, FineGrid g, const int *__restrict__ xwin)
{
    const int lo = xwin[0], cnt = xwin[1];
    if (cnt >= g.nd[0]) return;
    int64_t per0 = 1;
    for (int d = 1; d < g.dim; ++d) per0 *= g.nf[d];
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < g.ncells; c += (int64_t)gridDim.x * blockDim.x) {
        const int r0 = (int)(c / per0) / g.m[0];
        int rel = (r0 - lo) % g.nd[0];
        if (rel < 0) rel += g.nd[0];
        if (rel >= cnt) count[c] = 0u;
    }
}

// first half of the sort: allocations, fine cell id and arrival rank of every point, points per fine cell
int htb_sort_begin(cudaStream_t st, Workspace &ws, const FineGrid &g,
                   const double *const *cd, int64_t stride, int64_t n,
                   const double *w_dev, int nw, bool keep_perm, SortedSample &out, int *launches)
{
    out.n = n;
    out.g = g;
    out.nw = w_dev ? nw : 0;
    out.npad = ((n + 1) & ~(int64_t)1) + 2;
    for (int d = 0; d < g.dim; ++d)
        if (ws.alloc((void **)&out.c[d], sizeof(double) * (size_t)out.npad)) return 1;
    if (ws.alloc((void **)&out.off, sizeof(uint32_t) * (size_t)(g.ncells + 2))) return 1;
    if (keep_perm && ws.alloc((void **)&out.perm, sizeof(uint32_t) * (size_t)(n > 0 ? n : 1))) return 1;
    if (w_dev && ws.alloc((void **)&out.w, sizeof(double) * (size_t)((n > 0 ? n : 1) * nw + 2))) return 1;
    if (ws.alloc((void **)&out.cell, sizeof(uint32_t) * (size_t)(n > 0 ? n : 1))) return 1;
    // the flags word sits behind the counters: one memset for both
    if (ws.alloc((void **)&out.count, sizeof(uint32_t) * (size_t)(g.ncells + 2 + 4))) return 1;
    out.flags = out.count + g.ncells + 2;
    HTB_CUDA(cudaMemsetAsync(out.count, 0, sizeof(uint32_t) * (size_t)(g.ncells + 2 + 4), st));
    if (n > 0) {
        const int blocks = grid_for(n, 256);
        if (g.dim == 3)
            k_assign<3><<<blocks, 256, 0, st>>>(cd[0], cd[1], cd[2], stride, n, g, out.cell, out.count, out.flags);
        else
            k_assign<2><<<blocks, 256, 0, st>>>(cd[0], cd[1], nullptr, stride, n, g, out.cell, out.count, out.flags);
        if (launches) *launches += 1;
    }
    HTB_CUDA(cudaGetLastError());
    return 0;
}

// second half: (optional window) -> offsets -> scatter into SoA order
int htb_sort_finish(cudaStream_t st, Workspace &ws, const double *const *cd, int64_t stride,
                    const double *w_dev, int nw, double pad_value, const int *xwin_dev, SortedSample &out, int *launches)
{
    const FineGrid &g = out.g;
    const int64_t n = out.n;
    if (xwin_dev) {
        k_mask_counts<<<grid_for(g.ncells, 256), 256, 0, st>>>(out.count, g, xwin_dev);
        if (launches) *launches += 1;
    }
    // off[0..ncells] : exclusive scan over ncells+1 entries (the extra entry is zero) gives off[ncells] = points kept
    // (the per-cell counts, once scanned into `off`, are zeroed by the scan: the array then serves as the cells' fill counters)
    if (htb_exclusive_scan_u32(st, ws, out.count, out.off, g.ncells + 1, nullptr, launches, true)) return 1;
    if (n > 0) {
        const int64_t nchunk = (n + PART_CHUNK - 1) / PART_CHUNK;
        const int blocks = (int)(nchunk < 148 * 8 ? nchunk : 148 * 8);
        if (g.dim == 3)
            k_scatter<3><<<blocks, PART_THREADS, 0, st>>>(cd[0], cd[1], cd[2], stride, n, out.cell, out.off, out.count,
                                                          out.c[0], out.c[1], out.c[2], out.perm, w_dev, out.w, nw, out.npad, pad_value);
        else
            k_scatter<2><<<blocks, PART_THREADS, 0, st>>>(cd[0], cd[1], nullptr, stride, n, out.cell, out.off, out.count,
                                                          out.c[0], out.c[1], nullptr, out.perm, w_dev, out.w, nw, out.npad, pad_value);
        if (launches) *launches += 1;
    }
    if (n <= 0) {
        k_pad<<<1, 32, 0, st>>>(out.c[0], out.c[1], g.dim == 3 ? out.c[2] : nullptr, n, out.npad, pad_value);
        if (launches) *launches += 1;
    }
    HTB_CUDA(cudaGetLastError());
    return 0;
}

int htb_sort_sample(cudaStream_t st, Workspace &ws, const FineGrid &g,
                    const double *const *cd, int64_t stride, int64_t n,
                    const double *w_dev, int nw, bool keep_perm, double pad_value, SortedSample &out, int *launches)
{
    if (htb_sort_begin(st, ws, g, cd, stride, n, w_dev, nw, keep_perm, out, launches)) return 1;
    return htb_sort_finish(st, ws, cd, stride, w_dev, nw, pad_value, nullptr, out, launches);
}

// ------------------------------------------------------------------ reference cell ids only
template <int DIM>
__global__ void k_ref_ids(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                          int64_t stride, int64_t n, double c0, double c1, double c2, int n0, int n1, int n2,
                          int64_t *__restrict__ ids)
{
    const double *src[3] = {x, y, z};
    const double cs[3] = {c0, c1, c2};
    const int nd[3] = {n0, n1, n2};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t cid = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) cid = cid * nd[d] + htb_ref_digitize(src[d][i * stride], cs[d], nd[d]);
        ids[i] = cid;
    }
}

int htb_ref_cell_ids(cudaStream_t st, int dim, const double *const *cd, int64_t stride, int64_t n,
                     const double *cell_size, const int *ndivs, int64_t *ids_dev, int *launches)
{
    if (n <= 0) return 0;
    const int blocks = grid_for(n, 256);
    if (dim == 3)
        k_ref_ids<3><<<blocks, 256, 0, st>>>(cd[0], cd[1], cd[2], stride, n, cell_size[0], cell_size[1], cell_size[2],
                                            ndivs[0], ndivs[1], ndivs[2], ids_dev);
    else
        k_ref_ids<2><<<blocks, 256, 0, st>>>(cd[0], cd[1], nullptr, stride, n, cell_size[0], cell_size[1], 1.0,
                                            ndivs[0], ndivs[1], 1, ids_dev);
    if (launches) *launches += 1;
    HTB_CUDA(cudaGetLastError());
    return 0;
}

// per reference-cell point counts from the fine offsets
__global__ void k_ref_counts(const uint32_t *__restrict__ off, FineGrid g, uint32_t *__restrict__ counts)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < g.ncells; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = off[c + 1] - off[c];
        if (!k) continue;
        int64_t rem = c;
        int f[3] = {0, 0, 0};
        for (int d = g.dim - 1; d >= 0; --d) { f[d] = (int)(rem % g.nf[d]); rem /= g.nf[d]; }
        int64_t rid = 0;
        for (int d = 0; d < g.dim; ++d) rid = rid * g.nd[d] + f[d] / g.m[d];
        atomicAdd(&counts[rid], k);
    }
}

// ... from the per-fine-cell point counts (before the offsets exist)
__global__ void k_ref_counts_c(const uint32_t *__restrict__ count, FineGrid g, uint32_t *__restrict__ counts)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < g.ncells; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = count[c];
        if (!k) continue;
        int64_t rem = c;
        int f[3] = {0, 0, 0};
        for (int d = g.dim - 1; d >= 0; --d) { f[d] = (int)(rem % g.nf[d]); rem /= g.nf[d]; }
        int64_t rid = 0;
        for (int d = 0; d < g.dim; ++d) rid = rid * g.nd[d] + f[d] / g.m[d];
        atomicAdd(&counts[rid], k);
    }
}

int htb_ref_cell_counts_pre(cudaStream_t st, const SortedSample &s, uint32_t *counts_dev, int *launches)
{
    const int blocks = grid_for(s.g.ncells, 256);
    k_ref_counts_c<<<blocks, 256, 0, st>>>(s.count, s.g, counts_dev);
    if (launches) *launches += 1;
    HTB_CUDA(cudaGetLastError());
    return 0;
}

int htb_ref_cell_counts(cudaStream_t st, const SortedSample &s, uint32_t *counts_dev, int *launches)
{
    const int blocks = grid_for(s.g.ncells, 256);
    k_ref_counts<<<blocks, 256, 0, st>>>(s.off, s.g, counts_dev);
    if (launches) *launches += 1;
    HTB_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ column extrema of a device-resident sample
// The front-ends' bounds check (mock_observables_helpers.py:25-71) for samples that already live on the GPU: one
// streaming pass over the (n, cols) rows - HBM bound, 8 * cols bytes per point.  part[b] = {min[3], max[3], nan}.
#define HTB_MM_BLOCK 256
__global__ void __launch_bounds__(HTB_MM_BLOCK)
k_minmax(const double *__restrict__ base, int64_t n, int64_t stride, int cols, double *__restrict__ part)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool bad = false;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += 4 * step) {
        double v[4][3];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t rr = r + u * step;
            ok[u] = rr < n;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[u][c] = (ok[u] && c < cols) ? base[rr * stride + c] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (ok[u] && c < cols) {
                    lo[c] = fmin(lo[c], v[u][c]);
                    hi[c] = fmax(hi[c], v[u][c]);
                    bad |= (v[u][c] != v[u][c]);
                }
            }
    }
    __shared__ double s[HTB_MM_BLOCK / 32][7];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    const bool anybad = __any_sync(0xffffffffu, bad);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        for (int c = 0; c < 3; ++c) { s[w][c] = lo[c]; s[w][3 + c] = hi[c]; }
        s[w][6] = anybad ? 1.0 : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double out[7];
        for (int k = 0; k < 7; ++k) out[k] = s[0][k];
        for (int i = 1; i < HTB_MM_BLOCK / 32; ++i) {
            for (int c = 0; c < 3; ++c) { out[c] = fmin(out[c], s[i][c]); out[3 + c] = fmax(out[3 + c], s[i][3 + c]); }
            out[6] = fmax(out[6], s[i][6]);
        }
        for (int k = 0; k < 7; ++k) part[(size_t)blockIdx.x * 7 + k] = out[k];
    }
}

int htb_device_minmax_launch(cudaStream_t st, const double *base_dev, int64_t n, int64_t stride, int cols,
                             double *part_dev, int blocks)
{
    k_minmax<<<blocks, HTB_MM_BLOCK, 0, st>>>(base_dev, n, stride, cols, part_dev);
    HTB_CUDA(cudaGetLastError());
    return 0;
}
