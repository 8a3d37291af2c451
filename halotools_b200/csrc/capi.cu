// C ABI of libhalotools_b200.so (declared in include/halotools_b200.h).
// Host-side orchestration of one engine call: H2D -> K1 mesh sort of both samples ->
// tile list -> K2 counting kernel -> D2H of the (tiny) result.  Everything is issued on
// one CUDA stream with stream-ordered allocations; the only host sync is the final one.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>          // header-only NVTX 3: ranges cost a null-pointer test when no tool is attached

#include "count.cuh"

// ------------------------------------------------------------------ errors / globals
static thread_local std::string g_err;
static thread_local cudaStream_t g_user_stream = nullptr;
static thread_local bool g_have_user_stream = false;
static cudaStream_t g_lib_stream[64] = {nullptr};

void htb_set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

extern "C" const char *htb_last_error(void) { return g_err.c_str(); }
extern "C" int htb_abi_version(void) { return HTB_ABI_VERSION; }
extern "C" int htb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int htb_set_device(int device)
{
    HTB_CUDA(cudaSetDevice(device));
    return 0;
}
extern "C" int htb_set_stream(void *s)
{
    g_user_stream = (cudaStream_t)s;
    g_have_user_stream = (s != nullptr);
    return 0;
}

// Multi-GPU shard of the calling thread's engine calls (htb_set_shard): rank r of `world` ranks.
static thread_local int g_shard_rank = 0, g_shard_world = 1;
extern "C" int htb_set_shard(int rank, int world)
{
    if (world < 1 || rank < 0 || rank >= world) { htb_set_error("htb_set_shard: need 0 <= rank < world"); return 1; }
    g_shard_rank = rank;
    g_shard_world = world;
    return 0;
}

// Upload cache (htb_cache_begin / htb_cache_end): between the two calls the coordinate arrays a thread's engine
// calls bring to the device stay there, keyed by (host pointers, stride, count), so the DD, DR and RR counts of one
// tpcf() call move every sample across PCIe once (the reference re-gathers every sample in every call,
// tpcf.py:76-113,164-205).  The caller promises not to modify the arrays in between.
struct UploadEnt {
    const double *src[3];
    int cnt;
    int64_t stride, n;
    const double *dev[3];
    int64_t dstride;
    void *blocks[3];
    int nblocks;
    int device;
    cudaStream_t st;
};
static thread_local bool g_cache_on = false;
static thread_local std::vector<UploadEnt> *g_cache = nullptr;

static void sort_cache_clear();
static void prep_cache_clear();

extern "C" int htb_cache_begin(void)
{
    if (!g_cache) g_cache = new std::vector<UploadEnt>();
    g_cache_on = true;
    return 0;
}
extern "C" int htb_cache_end(void)
{
    g_cache_on = false;
    sort_cache_clear();
    prep_cache_clear();
    if (g_cache) {
        for (auto &e : *g_cache)
            for (int k = 0; k < e.nblocks; ++k) cudaFreeAsync(e.blocks[k], e.st);
        g_cache->clear();
    }
    return 0;
}

// Sorted-sample cache (same scope as the upload cache): inside one statistic the counting sort of a sample on a given
// fine grid is done ONCE - tpcf's randoms are sample2 of the DR count and both samples of the RR count on the same
// grid (the reference builds a new RectangularDoubleMesh, i.e. two argsorts, in every npairs call, tpcf.py:76-113,164-205).
// Entries own their device memory (their own Workspace on the creating stream); a hit from another stream waits for the
// entry's `ready` event and is remembered, so that htb_cache_end() frees the blocks behind every user.
struct SortEnt {
    const double *src[3];
    int64_t stride, n;
    const double *w;
    int nw;
    bool perm;
    double pad;
    FineGrid g;
    SortedSample s;
    Workspace ws;
    cudaEvent_t ready;
    std::vector<cudaStream_t> users;
};
static thread_local std::vector<SortEnt *> *g_sort_cache = nullptr;

// The same for one rank of a SHARDED call, where the sort of the two samples, the rank's cell range and its windows
// belong together: keyed by everything that determines them.  It lets a statistic run the set-up of all its counts
// first (HTB_FLAG_PREPARE) and launch the count kernels afterwards - see DeviceStatistic.
struct PrepKey {
    const double *d1[3], *d2[3], *dw1, *dw2;
    int64_t ds1, ds2, n1, n2, first, last;
    htb_mesh_geom geom;
    int m1[3], m2[3];
    int nw, perm1, sym, rank, world, window;
    double pad;
};
struct PrepEnt {
    PrepKey key;
    SortedSample s1, s2;
    long long *range_dev;
    double *work_dev;
    int64_t nc1;
    Workspace ws;
    cudaEvent_t ready;
    std::vector<cudaStream_t> users;
};
static thread_local std::vector<PrepEnt *> *g_prep_cache = nullptr;

static void sort_cache_clear()
{
    if (!g_sort_cache) return;
    for (SortEnt *e : *g_sort_cache) {
        for (cudaStream_t u : e->users) {
            if (u == e->ws.st) continue;
            cudaEvent_t ev;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
                cudaEventRecord(ev, u);
                cudaStreamWaitEvent(e->ws.st, ev, 0);
                cudaEventDestroy(ev);
            }
        }
        e->ws.release();
        if (e->ready) cudaEventDestroy(e->ready);
        delete e;
    }
    g_sort_cache->clear();
}

static void prep_cache_clear()
{
    if (!g_prep_cache) return;
    for (PrepEnt *e : *g_prep_cache) {
        for (cudaStream_t u : e->users) {
            if (u == e->ws.st) continue;
            cudaEvent_t ev;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
                cudaEventRecord(ev, u);
                cudaStreamWaitEvent(e->ws.st, ev, 0);
                cudaEventDestroy(ev);
            }
        }
        e->ws.release();
        if (e->ready) cudaEventDestroy(e->ready);
        delete e;
    }
    g_prep_cache->clear();
}

static bool g_pool_ready[64] = {false};

static int get_stream(cudaStream_t *out)
{
    int dev = 0;
    HTB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { htb_set_error("device index %d out of range", dev); return 1; }
    if (!g_pool_ready[dev]) {
        // keep freed blocks cached in the stream-ordered pool between calls (the default is to hand them
        // back to the driver at every synchronisation, which costs far more than the kernels)
        cudaMemPool_t pool;
        HTB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t thresh = UINT64_MAX;
        HTB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        g_pool_ready[dev] = true;
    }
    if (g_have_user_stream) { *out = g_user_stream; return 0; }
    if (!g_lib_stream[dev]) HTB_CUDA(cudaStreamCreateWithFlags(&g_lib_stream[dev], cudaStreamNonBlocking));
    *out = g_lib_stream[dev];
    return 0;
}

int Workspace::alloc(void **p, size_t bytes)
{
    if (nptrs >= 256) { htb_set_error("workspace pointer table full"); return 1; }
    if (bytes == 0) bytes = 16;
    bytes = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMallocAsync(p, bytes, st);
    if (e != cudaSuccess) {
        htb_set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return 1;
    }
    ptrs[nptrs++] = *p;
    return 0;
}
void Workspace::release()
{
    for (int i = nptrs - 1; i >= 0; --i) cudaFreeAsync(ptrs[i], st);
    nptrs = 0;
}

#define HTB_GUARD_BEGIN try {
#define HTB_GUARD_END                                                          \
    } catch (const std::exception &e) { htb_set_error("exception: %s", e.what()); return 1; }

// ------------------------------------------------------------------ refinement heuristics
static void env_triplet(const char *name, int *m, int dim)
{
    const char *e = getenv(name);
    if (!e) return;
    int a = 0, b = 0, c = 0;
    int k = sscanf(e, "%d,%d,%d", &a, &b, &c);
    int v[3] = {a, k > 1 ? b : a, k > 2 ? c : (k > 1 ? b : a)};
    for (int d = 0; d < dim; ++d) if (v[d] >= 1 && v[d] <= 64) m[d] = v[d];
}

static int64_t cell_budget(int64_t n)
{
    int64_t b = 2 * n;
    if (b < (1 << 16)) b = 1 << 16;
    if (b > (1 << 25)) b = 1 << 25;
    return b;
}

static int clampi(double v, int lo, int hi)
{
    if (!(v > lo)) return lo;
    if (v > hi) return hi;
    return (int)v;
}

static void choose_refinement(const htb_mesh_geom *g, int64_t n1, int64_t n2, int tile, int *m1, int *m2, int mode = 0)
{
    const int dim = g->ndim, F = dim - 1, S = dim - 1;
    double vol = 1.0;
    for (int d = 0; d < dim; ++d) vol *= g->period[d];
    const double dens1 = (double)(n1 > 0 ? n1 : 1) / vol, dens2 = (double)(n2 > 0 ? n2 : 1) / vol;
    // sample1: a tile should be roughly a cube / square
    const double side = pow((double)tile / dens1, 1.0 / dim);
    for (int d = 0; d < dim; ++d) {
        if (d == F) m1[d] = clampi(floor(4.0 * g->cell1_size[d] / side + 0.5), 1, 16);
        else m1[d] = clampi(floor(g->cell1_size[d] / side + 0.5), 1, 8);
    }
    // sample2: columns fine enough to prune, coarse enough that a span holds >= ~128 points
    double colarea = 1.0;
    for (int d = 0; d < S; ++d) colarea *= g->cell2_size[d];
    const double span = dens2 * colarea * 1.6 * g->search[F];
    const double scale = pow(span / 128.0, 1.0 / (S > 0 ? S : 1));
    for (int d = 0; d < S; ++d) {
        const int cap = clampi(floor(4.0 * g->cell2_size[d] / g->search[d] + 0.5), 1, 8);
        m2[d] = clampi(floor(scale), 1, cap);
    }
    // along the fast dimension the spans are cut at fine-cell resolution (half a cell of wasted evaluations at either
    // end): cells of a 32nd of the search length, as long as a cell still holds ~5 points on average - below that the
    // sort of the larger mesh costs what the tighter spans save (configs 1, 3, 4 stay near the round-1 value of an 8th;
    // the bench's RR launch evaluates 11 % fewer pairs, 61.0 -> 55.8 ms; profiles/r02_refine_sweep.txt).
    {
        double fz = 32.0;
        if (const char *e = getenv("HTB_FZ")) { const double v = atof(e); if (v >= 1.0 && v <= 64.0) fz = v; }
        double colfine = 1.0;
        for (int d = 0; d < S; ++d) colfine *= g->cell2_size[d] / m2[d];
        const int cap = clampi(floor(dens2 * colfine * g->cell2_size[F] / 5.0), 8, 64);
        m2[F] = clampi(floor(fz * g->cell2_size[F] / g->search[F] + 0.5), 1, std::min(cap, (int)fz));
    }
    if (mode == 1) {
        // cell-resolved kernels (DSigmaR) decide per fine cell of sample2: cells small against the bins (a
        // sixteenth of the search length) but holding >= ~256 points, so the per-cell work is amortised
        const double side_n = pow(256.0 / dens2, 1.0 / dim);
        for (int d = 0; d < dim; ++d) {
            const double side = std::max(g->search[d] / 16.0, side_n);
            m2[d] = clampi(floor(g->cell2_size[d] / side + 0.5), 1, 16);
        }
    }
    // respect the cell budgets
    for (int which = 0; which < 2; ++which) {
        int *m = which ? m2 : m1;
        const int32_t *nd = which ? g->ndivs2 : g->ndivs1;
        const int64_t budget = cell_budget(which ? n2 : n1);
        while (true) {
            int64_t nc = 1;
            for (int d = 0; d < dim; ++d) nc *= (int64_t)nd[d] * m[d];
            if (nc <= budget) break;
            if (m[F] > 1) { m[F] = (m[F] + 1) / 2; continue; }
            int big = 0;
            for (int d = 0; d < S; ++d) if (m[d] > m[big]) big = d;
            if (m[big] > 1) { m[big] -= 1; continue; }
            break;
        }
    }
    env_triplet("HTB_M1", m1, dim);
    env_triplet("HTB_M2", m2, dim);
}

static FineGrid make_grid(const htb_mesh_geom *g, int which, const int *m)
{
    FineGrid f{};
    f.dim = g->ndim;
    f.ncells = 1;
    for (int d = 0; d < g->ndim; ++d) {
        f.nd[d] = which ? g->ndivs2[d] : g->ndivs1[d];
        f.m[d] = m[d];
        f.nf[d] = f.nd[d] * m[d];
        f.cs[d] = which ? g->cell2_size[d] : g->cell1_size[d];
        f.h[d] = f.cs[d] / m[d];
        f.period[d] = g->period[d];
        f.ncells *= f.nf[d];
    }
    for (int d = g->ndim; d < 3; ++d) { f.nd[d] = 1; f.m[d] = 1; f.nf[d] = 1; f.cs[d] = 1; f.h[d] = 1; f.period[d] = 1; }
    return f;
}

// ------------------------------------------------------------------ host -> device staging of pageable arrays
// cudaMemcpyAsync from pageable memory goes through the driver's single bounce buffer (about 2 GB/s measured on
// the B200 boxes).  Large pageable inputs are therefore packed by a few host threads into a ring of pinned
// chunks (de-interleaving strided columns on the way) and copied chunk by chunk on per-thread streams.
#define HTB_STAGE_THREADS 16
#define HTB_STAGE_CHUNK (256 * 1024)          // elements per column and chunk
struct StagePool {
    double *pinned[HTB_STAGE_THREADS][2] = {{nullptr}};
    cudaEvent_t ev[HTB_STAGE_THREADS][2] = {{nullptr}};
    cudaEvent_t done[HTB_STAGE_THREADS] = {nullptr};
    cudaEvent_t dst_ready = nullptr;          // the destination blocks are ours in the order of the engine's stream
    cudaStream_t st[HTB_STAGE_THREADS] = {nullptr};
    int cols = 0;
    bool ready = false;
};
static StagePool g_stage[64];

static int stage_pool_init(int dev, int cols)
{
    StagePool &sp = g_stage[dev];
    if (sp.ready && sp.cols >= cols) return 0;
    for (int t = 0; t < HTB_STAGE_THREADS; ++t)
        for (int b = 0; b < 2; ++b) {
            if (sp.pinned[t][b]) cudaFreeHost(sp.pinned[t][b]);
            HTB_CUDA(cudaHostAlloc((void **)&sp.pinned[t][b], sizeof(double) * (size_t)HTB_STAGE_CHUNK * cols, cudaHostAllocDefault));
            if (!sp.ev[t][b]) HTB_CUDA(cudaEventCreateWithFlags(&sp.ev[t][b], cudaEventDisableTiming));
        }
    for (int t = 0; t < HTB_STAGE_THREADS; ++t) {
        if (!sp.st[t]) HTB_CUDA(cudaStreamCreateWithFlags(&sp.st[t], cudaStreamNonBlocking));
        if (!sp.done[t]) HTB_CUDA(cudaEventCreateWithFlags(&sp.done[t], cudaEventDisableTiming));
    }
    if (!sp.dst_ready) HTB_CUDA(cudaEventCreateWithFlags(&sp.dst_ready, cudaEventDisableTiming));
    sp.cols = cols;
    sp.ready = true;
    return 0;
}

static bool host_pointer_is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// copy cnt strided host columns of n elements into cnt contiguous device arrays; `st` waits for the copies
static int staged_upload(cudaStream_t st, const double *const *src, int cnt, int64_t stride, int64_t n, double *const *dst)
{
    int dev = 0;
    HTB_CUDA(cudaGetDevice(&dev));
    if (stage_pool_init(dev, cnt)) return 1;
    StagePool &sp = g_stage[dev];
    const int64_t nchunk = (n + HTB_STAGE_CHUNK - 1) / HTB_STAGE_CHUNK;
    int nthreads = (int)std::min<int64_t>(HTB_STAGE_THREADS, nchunk);
    unsigned hw = std::thread::hardware_concurrency();
    if (hw > 0 && g_shard_world > 1) hw = std::max(2u, hw / (unsigned)g_shard_world);     // one process per GPU shares the host's cores
    if (hw > 0 && (unsigned)nthreads > hw) nthreads = (int)hw;
    // the destination blocks come from the stream-ordered pool of `st`: they may still be in use by work enqueued
    // earlier on `st` (asynchronous calls), so the copy streams start behind everything `st` holds now
    HTB_CUDA(cudaEventRecord(sp.dst_ready, st));
    std::atomic<int> failed(0);
    auto work = [&](int t) {
        if (cudaSetDevice(dev) != cudaSuccess) { failed = 1; return; }
        if (cudaStreamWaitEvent(sp.st[t], sp.dst_ready, 0) != cudaSuccess) { failed = 1; return; }
        int k = 0;
        for (int64_t c = t; c < nchunk; c += nthreads, ++k) {
            const int b = k & 1;
            // (a never-recorded event is complete; a buffer used by an earlier upload of this call may still be in flight)
            if (cudaEventSynchronize(sp.ev[t][b]) != cudaSuccess) { failed = 1; return; }
            const int64_t i0 = c * HTB_STAGE_CHUNK;
            const int64_t len = std::min<int64_t>(HTB_STAGE_CHUNK, n - i0);
            double *buf = sp.pinned[t][b];
            for (int col = 0; col < cnt; ++col) {
                const double *s = src[col] + i0 * stride;
                double *d = buf + (size_t)col * HTB_STAGE_CHUNK;
                if (stride == 1) memcpy(d, s, sizeof(double) * (size_t)len);
                else for (int64_t i = 0; i < len; ++i) d[i] = s[i * stride];
            }
            for (int col = 0; col < cnt; ++col)
                if (cudaMemcpyAsync(dst[col] + i0, buf + (size_t)col * HTB_STAGE_CHUNK, sizeof(double) * (size_t)len,
                                    cudaMemcpyHostToDevice, sp.st[t]) != cudaSuccess) { failed = 1; return; }
            if (cudaEventRecord(sp.ev[t][b], sp.st[t]) != cudaSuccess) { failed = 1; return; }
        }
        if (cudaEventRecord(sp.done[t], sp.st[t]) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (failed) { htb_set_error("staged host->device upload failed: %s", cudaGetErrorString(cudaGetLastError())); return 1; }
    for (int t = 0; t < nthreads; ++t) HTB_CUDA(cudaStreamWaitEvent(st, sp.done[t], 0));
    // the pinned chunks are reused by the next call: make sure this call's copies are done before it returns
    // (every engine call ends with a stream synchronisation of `st`, which now depends on them)
    return 0;
}

// device -> host for large outputs into pageable memory (per-object tables): the same ring of pinned chunks, filled by
// per-thread streams that wait for `st`, emptied by the host threads; returns when `dst` is complete
static int staged_download(cudaStream_t st, const double *src_dev, double *dst, int64_t n)
{
    int dev = 0;
    HTB_CUDA(cudaGetDevice(&dev));
    if (stage_pool_init(dev, 1)) return 1;
    StagePool &sp = g_stage[dev];
    cudaEvent_t ready;
    HTB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    HTB_CUDA(cudaEventRecord(ready, st));
    const int64_t nchunk = (n + HTB_STAGE_CHUNK - 1) / HTB_STAGE_CHUNK;
    int nthreads = (int)std::min<int64_t>(HTB_STAGE_THREADS, nchunk);
    unsigned hw = std::thread::hardware_concurrency();
    if (hw > 0 && g_shard_world > 1) hw = std::max(2u, hw / (unsigned)g_shard_world);     // one process per GPU shares the host's cores
    if (hw > 0 && (unsigned)nthreads > hw) nthreads = (int)hw;
    std::atomic<int> failed(0);
    auto work = [&](int t) {
        if (cudaSetDevice(dev) != cudaSuccess) { failed = 1; return; }
        if (cudaStreamWaitEvent(sp.st[t], ready, 0) != cudaSuccess) { failed = 1; return; }
        int64_t prev = -1;
        int k = 0;
        auto drain = [&](int64_t c, int b) {
            if (cudaEventSynchronize(sp.ev[t][b]) != cudaSuccess) { failed = 1; return; }
            const int64_t i0 = c * HTB_STAGE_CHUNK;
            const int64_t len = std::min<int64_t>(HTB_STAGE_CHUNK, n - i0);
            memcpy(dst + i0, sp.pinned[t][b], sizeof(double) * (size_t)len);
        };
        for (int64_t c = t; c < nchunk; c += nthreads, ++k) {
            const int b = k & 1;
            const int64_t i0 = c * HTB_STAGE_CHUNK;
            const int64_t len = std::min<int64_t>(HTB_STAGE_CHUNK, n - i0);
            // buffer b was drained two chunks ago (or never used by this call); any earlier use has been synchronised
            if (cudaMemcpyAsync(sp.pinned[t][b], src_dev + i0, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, sp.st[t]) != cudaSuccess ||
                cudaEventRecord(sp.ev[t][b], sp.st[t]) != cudaSuccess) { failed = 1; return; }
            if (prev >= 0) drain(prev, b ^ 1);
            if (failed) return;
            prev = c;
        }
        if (prev >= 0) drain(prev, (k - 1) & 1);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    cudaEventDestroy(ready);
    if (failed) { htb_set_error("staged device->host download failed: %s", cudaGetErrorString(cudaGetLastError())); return 1; }
    return 0;
}

// copy an output table to the caller: large pageable destinations go through the pinned ring
static int download(cudaStream_t st, const double *src_dev, double *dst, int64_t n, uint32_t flags)
{
    (void)flags;
    if (n <= 0) return 0;
    if (n >= 8 * (int64_t)HTB_STAGE_CHUNK && host_pointer_is_pageable(dst) && !getenv("HTB_NO_STAGED_UPLOAD"))
        return staged_download(st, src_dev, dst, n);
    HTB_CUDA(cudaMemcpyAsync(dst, src_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    return 0;
}

// column-wise minimum and maximum of `cols` strided host columns (threads; one pass over memory)
// rows [a, b) of a contiguous (n, 3) matrix: 24 doubles = 8 rows per trip, lane j belongs to column j % 3; the inner loop
// has a fixed trip count and no branches, so the host compiler vectorises it (min / max / a NaN detector that needs no
// compare: v - v is 0 for finite v).  1.4 x the row-by-row loop with SSE2, 2 x with AVX2 (measured per core).
#define HTB_MM_SCAN3_BODY                                                                                         \
    double lo[24], hi[24], nanw[24];                                                                              \
    for (int j = 0; j < 24; ++j) { lo[j] = INFINITY; hi[j] = -INFINITY; nanw[j] = 0.0; }                         \
    const double *p = base + a * 3;                                                                               \
    const int64_t nblk = (b - a) / 8;                                                                             \
    for (int64_t k = 0; k < nblk; ++k, p += 24) {                                                                 \
        for (int j = 0; j < 24; ++j) {                                                                            \
            const double v = p[j];                                                                                \
            lo[j] = v < lo[j] ? v : lo[j];                                                                        \
            hi[j] = v > hi[j] ? v : hi[j];                                                                        \
            nanw[j] += v - v;                                                                                     \
        }                                                                                                         \
    }                                                                                                             \
    bool bad = false;                                                                                             \
    for (int c = 0; c < 3; ++c) { l[c] = INFINITY; h[c] = -INFINITY; }                                            \
    for (int j = 0; j < 24; ++j) {                                                                                \
        if (lo[j] < l[j % 3]) l[j % 3] = lo[j];                                                                   \
        if (hi[j] > h[j % 3]) h[j % 3] = hi[j];                                                                   \
        bad |= !(nanw[j] == 0.0);                     /* a NaN (or an infinity: then re-checked below) */          \
    }                                                                                                             \
    for (int64_t i = a + nblk * 8; i < b; ++i)                                                                    \
        for (int c = 0; c < 3; ++c) {                                                                             \
            const double v = base[i * 3 + c];                                                                     \
            if (v < l[c]) l[c] = v;                                                                               \
            if (v > h[c]) h[c] = v;                                                                               \
            bad |= (v != v);                                                                                      \
        }                                                                                                         \
    *maybe_nan = bad;
static void mm_scan3_generic(const double *base, int64_t a, int64_t b, double *l, double *h, bool *maybe_nan) { HTB_MM_SCAN3_BODY }
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static void mm_scan3_avx2(const double *base, int64_t a, int64_t b, double *l, double *h, bool *maybe_nan) { HTB_MM_SCAN3_BODY }
#endif

extern "C" int htb_host_minmax(const double *base, int64_t n, int64_t stride, int32_t cols, double *min_out, double *max_out)
{
    HTB_GUARD_BEGIN
    if (!base || cols < 1 || cols > 8 || !min_out || !max_out || stride < cols) { htb_set_error("htb_host_minmax: bad arguments"); return 1; }
    for (int c = 0; c < cols; ++c) { min_out[c] = INFINITY; max_out[c] = -INFINITY; }
    if (n <= 0) return 0;
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<int64_t>(hw ? hw : 1, std::min<int64_t>(16, (n + 65535) / 65536));
    std::vector<double> lo((size_t)nt * 8, INFINITY), hi((size_t)nt * 8, -INFINITY), nan((size_t)nt, 0.0);
    bool avx2 = false;
#if defined(__x86_64__) && defined(__GNUC__)
    avx2 = __builtin_cpu_supports("avx2");
#endif
    auto work = [&](int t) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        double l[8], h[8];
        bool bad = false;
        for (int c = 0; c < cols; ++c) { l[c] = INFINITY; h[c] = -INFINITY; }
        bool fast = cols == 3 && stride == 3;
        if (fast) {
            bool maybe = false;
#if defined(__x86_64__) && defined(__GNUC__)
            if (avx2) mm_scan3_avx2(base, a, b, l, h, &maybe); else
#endif
            mm_scan3_generic(base, a, b, l, h, &maybe);
            // the detector also fires on infinities (inf - inf): only then look for real NaNs, row by row
            if (maybe) fast = false;
        }
        for (int64_t i = fast ? b : a; i < b; ++i) {
            const double *row = base + i * stride;
            for (int c = 0; c < cols; ++c) {
                const double v = row[c];
                if (v < l[c]) l[c] = v;
                if (v > h[c]) h[c] = v;
                bad |= (v != v);
            }
        }
        for (int c = 0; c < cols; ++c) { lo[(size_t)t * 8 + c] = l[c]; hi[(size_t)t * 8 + c] = h[c]; }
        nan[(size_t)t] = bad ? 1.0 : 0.0;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    bool anynan = false;
    for (int t = 0; t < nt; ++t) {
        anynan |= nan[(size_t)t] != 0.0;
        for (int c = 0; c < cols; ++c) {
            if (lo[(size_t)t * 8 + c] < min_out[c]) min_out[c] = lo[(size_t)t * 8 + c];
            if (hi[(size_t)t * 8 + c] > max_out[c]) max_out[c] = hi[(size_t)t * 8 + c];
        }
    }
    if (anynan) for (int c = 0; c < cols; ++c) { min_out[c] = NAN; max_out[c] = NAN; }   // comparisons with NaN fail, as in numpy
    return 0;
    HTB_GUARD_END
}

// the same for a DEVICE-resident matrix (torch CUDA tensors handed to the front-ends): one HBM-bound pass on the GPU
extern "C" int htb_device_minmax(const double *base_dev, int64_t n, int64_t stride, int32_t cols, double *min_out, double *max_out)
{
    HTB_GUARD_BEGIN
    if (!base_dev || cols < 1 || cols > 3 || !min_out || !max_out || stride < cols) { htb_set_error("htb_device_minmax: bad arguments"); return 1; }
    for (int c = 0; c < cols; ++c) { min_out[c] = INFINITY; max_out[c] = -INFINITY; }
    if (n <= 0) return 0;
    cudaStream_t st;
    if (get_stream(&st)) return 1;
    int dev = 0, sms = 0;
    HTB_CUDA(cudaGetDevice(&dev));
    HTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = (int)std::min<int64_t>((int64_t)sms * 8, (n + 1023) / 1024);
    double *part = nullptr;
    HTB_CUDA(cudaMallocAsync((void **)&part, sizeof(double) * 7 * (size_t)blocks, st));
    if (htb_device_minmax_launch(st, base_dev, n, stride, cols, part, blocks)) return 1;
    std::vector<double> h((size_t)blocks * 7);
    HTB_CUDA(cudaMemcpyAsync(h.data(), part, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    HTB_CUDA(cudaFreeAsync(part, st));
    HTB_CUDA(cudaStreamSynchronize(st));
    bool anynan = false;
    for (int b = 0; b < blocks; ++b) {
        anynan |= h[(size_t)b * 7 + 6] != 0.0;
        for (int c = 0; c < cols; ++c) {
            if (h[(size_t)b * 7 + c] < min_out[c]) min_out[c] = h[(size_t)b * 7 + c];
            if (h[(size_t)b * 7 + 3 + c] > max_out[c]) max_out[c] = h[(size_t)b * 7 + 3 + c];
        }
    }
    if (anynan) for (int c = 0; c < cols; ++c) { min_out[c] = NAN; max_out[c] = NAN; }
    return 0;
    HTB_GUARD_END
}

// ------------------------------------------------------------------ one engine call
// sum of the per-cell predicted work over the cells [first, last) (or the device-side shard range): W_ref of a call
__global__ void __launch_bounds__(256) k_sum_range(const double *__restrict__ work, long long ncells, long long first, long long last,
                                                   const long long *__restrict__ range, double *__restrict__ out)
{
    __shared__ double sh[256];
    if (range) { first = range[0]; last = range[1]; }
    if (first < 0) first = 0;
    if (last > ncells) last = ncells;
    double s = 0.0;
    for (long long c = first + threadIdx.x; c < last; c += 256) s += work[c];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// CUDA events of the per-call timing (htb_stats): created once per thread and device, reused by every call
// (five cudaEventCreate/Destroy per call were part of the 0.8 ms fixed cost of a small call)
static thread_local cudaEvent_t g_call_ev[64][5] = {{nullptr}};
// asynchronous calls (HTB_FLAG_DEVICE_OUTPUT) bracket their counting kernel with a pair of events from a small ring;
// htb_async_count_times() reads the elapsed times once the caller has synchronised (bench: the kernel's live duration
// inside the timed region)
#define HTB_ASYNC_RING 16
static thread_local cudaEvent_t g_async_ev[HTB_ASYNC_RING][2] = {{nullptr}};
static thread_local int g_async_n = 0;
// ... and the kernel's own device-side time stamps {~first warp in, last warp out} (globaltimer, ns; words 40..43 of the
// call's counter block), copied to a pinned ring right behind the kernel: htb_async_kernel_spans()
static thread_local unsigned long long (*g_async_ts)[2] = nullptr;

struct Call {
    cudaStream_t st = nullptr;
    Workspace ws;
    cudaEvent_t *ev = nullptr;            // this thread's five timing events on the current device (null: asynchronous call)
    int launches = 0;
    WalkGeom G{};
    WalkArrays A{};
    SortedSample s1, s2;
    int m1[3] = {1, 1, 1}, m2[3] = {1, 1, 1};
    unsigned int *ctr = nullptr;          // [0] tile counter, [1] tiles redone, [2..3] pairs evaluated (u64)
    int64_t max_tiles = 0;
    int64_t first_cell = 0, last_cell = 0;  // this call's range of reference mesh1 cells
    long long *range_dev = nullptr;         // device {first, last}: this rank's shard of that range (htb_set_shard)
    double *work_dev = nullptr;             // predicted work per reference mesh1 cell (computed once per call)
    int64_t nc1 = 0;
    uint32_t flags = 0;
    bool async = false;                     // HTB_FLAG_DEVICE_OUTPUT: results stay on the device, no host synchronisation
    bool prepared = false;                  // HTB_FLAG_PREPARE: setup() stopped after the mesh sorts (now in the caches)

    // NVTX ranges of the host side of a call (enqueue order = stream order): htb:h2d / htb:mesh_sort / htb:count /
    // htb:finalize, visible on the timeline of nsys / ncu next to the kernels they enqueue
    int nvtx_open = 0;
    void nvtx_next(const char *name)
    {
        if (nvtx_open) { nvtxRangePop(); nvtx_open = 0; }
        if (name) { nvtxRangePushA(name); nvtx_open = 1; }
    }
    ~Call() { nvtx_next(nullptr); ws.release(); }
    int begin(uint32_t fl = 0)
    {
        nvtx_next("htb:h2d");
        if (get_stream(&st)) return 1;
        ws.st = st;
        async = (fl & HTB_FLAG_DEVICE_OUTPUT) != 0;
        if (!async) {
            int dev = 0;
            HTB_CUDA(cudaGetDevice(&dev));
            ev = g_call_ev[dev];
            for (int i = 0; i < 5; ++i) if (!ev[i]) HTB_CUDA(cudaEventCreate(&ev[i]));
            HTB_CUDA(cudaEventRecord(ev[0], st));
        }
        return 0;
    }
    int mark(int i)
    {
        static const char *const names[5] = {"htb:h2d", "htb:mesh_sort", "htb:count", "htb:finalize", nullptr};
        nvtx_next(names[i]);
        if (ev) HTB_CUDA(cudaEventRecord(ev[i], st));
        else if (async && (i == 2 || i == 3)) {
            cudaEvent_t *pair = g_async_ev[g_async_n % HTB_ASYNC_RING];
            for (int k = 0; k < 2; ++k) if (!pair[k]) HTB_CUDA(cudaEventCreate(&pair[k]));
            HTB_CUDA(cudaEventRecord(pair[i - 2], st));
            if (i == 3) {
                if (!g_async_ts) HTB_CUDA(cudaMallocHost((void **)&g_async_ts, sizeof(unsigned long long) * 2 * HTB_ASYNC_RING));
                unsigned long long *slot = g_async_ts[g_async_n % HTB_ASYNC_RING];
                slot[0] = slot[1] = 0;
                if (ctr) HTB_CUDA(cudaMemcpyAsync(slot, ctr + 40, 16, cudaMemcpyDeviceToHost, st));
                ++g_async_n;
            }
        }
        return 0;
    }
    // bring `cnt` arrays of n elements (common element stride) to the device; returns device pointers + stride
    // *stable (optional): the device pointers keep naming this sample until htb_cache_end() (device-resident input inside
    // a cache scope, or a host sample held by the upload cache) - the condition for the sorted-sample cache
    int stage_coords(const double *const *src, int cnt, int64_t stride, int64_t n, const double **dst, int64_t *dstride,
                     bool cacheable = false, bool *stable = nullptr)
    {
        if (stable) *stable = false;
        if (flags & HTB_FLAG_DEVICE_INPUT) {
            for (int k = 0; k < cnt; ++k) dst[k] = src[k];
            *dstride = stride;
            if (stable) *stable = g_cache_on && n > 0;
            return 0;
        }
        if (n <= 0) { for (int k = 0; k < cnt; ++k) dst[k] = nullptr; *dstride = 1; return 0; }
        UploadEnt ent{};
        const bool caching = cacheable && g_cache_on && g_cache;
        int dev = 0;
        HTB_CUDA(cudaGetDevice(&dev));
        // blocks taken for the cache belong to nobody until done() hands them over: free them on every failure path
        struct BlockGuard {
            UploadEnt &e; cudaStream_t st; bool armed;
            ~BlockGuard() { if (armed) for (int k = 0; k < e.nblocks; ++k) cudaFreeAsync(e.blocks[k], st); }
        } guard{ent, st, true};
        if (caching) {
            for (auto &e : *g_cache) {
                bool hit = e.cnt == cnt && e.stride == stride && e.n == n && e.st == st && e.device == dev;
                for (int k = 0; k < cnt && hit; ++k) hit = e.src[k] == src[k];
                if (hit) { for (int k = 0; k < cnt; ++k) dst[k] = e.dev[k]; *dstride = e.dstride; if (stable) *stable = true; return 0; }
            }
            ent.cnt = cnt; ent.stride = stride; ent.n = n; ent.st = st; ent.device = dev;
            for (int k = 0; k < cnt; ++k) ent.src[k] = src[k];
        }
        // device blocks: owned by the call's workspace, or by the upload cache when it is on
        auto dev_alloc = [&](void **p, size_t bytes) -> int {
            if (!caching) return ws.alloc(p, bytes);
            bytes = (bytes + 255) & ~(size_t)255;
            cudaError_t e = cudaMallocAsync(p, bytes, st);
            if (e != cudaSuccess) { htb_set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e)); return 1; }
            ent.blocks[ent.nblocks++] = *p;
            return 0;
        };
        auto done = [&]() -> int {
            if (caching) {
                for (int k = 0; k < cnt; ++k) ent.dev[k] = dst[k];
                ent.dstride = *dstride;
                g_cache->push_back(ent);
                if (stable) *stable = true;
            }
            guard.armed = false;
            return 0;
        };
        if (n >= 4 * HTB_STAGE_CHUNK && host_pointer_is_pageable(src[0]) && !getenv("HTB_NO_STAGED_UPLOAD")) {
            double *bufs[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < cnt; ++k) {
                if (dev_alloc((void **)&bufs[k], sizeof(double) * (size_t)n)) return 1;
                dst[k] = bufs[k];
            }
            *dstride = 1;
            if (staged_upload(st, src, cnt, stride, n, bufs)) return 1;
            return done();
        }
        // columns of one row-major (n, stride) host matrix (e.g. x, y of an (N, 3) sample): ship the whole
        // block with ONE contiguous copy and keep the stride on the device
        bool interleaved = cnt > 1 && stride >= cnt;
        int64_t maxoff = 0;
        for (int k = 1; k < cnt && interleaved; ++k) {
            const int64_t off = src[k] - src[0];
            interleaved = off > 0 && off < stride;
            if (off > maxoff) maxoff = off;
        }
        if (interleaved) {
            double *buf = nullptr;
            const size_t len = (size_t)(n - 1) * (size_t)stride + (size_t)maxoff + 1;
            if (dev_alloc((void **)&buf, sizeof(double) * len)) return 1;
            HTB_CUDA(cudaMemcpyAsync(buf, src[0], sizeof(double) * len, cudaMemcpyHostToDevice, st));
            for (int k = 0; k < cnt; ++k) dst[k] = buf + (src[k] - src[0]);
            *dstride = stride;
            return done();
        }
        for (int k = 0; k < cnt; ++k) {
            double *buf = nullptr;
            if (dev_alloc((void **)&buf, sizeof(double) * (size_t)n)) return 1;
            if (stride == 1)
                HTB_CUDA(cudaMemcpyAsync(buf, src[k], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
            else
                HTB_CUDA(cudaMemcpy2DAsync(buf, sizeof(double), src[k], sizeof(double) * (size_t)stride, sizeof(double),
                                           (size_t)n, cudaMemcpyHostToDevice, st));
            dst[k] = buf;
        }
        *dstride = 1;
        return done();
    }
    int stage_rows(const double *src, int64_t n, int nw, const double **dst)
    {
        if (!src) { *dst = nullptr; return 0; }
        if (flags & HTB_FLAG_DEVICE_INPUT) { *dst = src; return 0; }
        double *buf = nullptr;
        if (ws.alloc((void **)&buf, sizeof(double) * (size_t)(n > 0 ? n : 1) * nw)) return 1;
        if (n * nw >= 4 * HTB_STAGE_CHUNK && host_pointer_is_pageable(src) && !getenv("HTB_NO_STAGED_UPLOAD")) {
            const double *one[1] = {src};
            double *dstp[1] = {buf};
            if (staged_upload(st, one, 1, 1, n * nw, dstp)) return 1;
        } else if (n > 0) HTB_CUDA(cudaMemcpyAsync(buf, src, sizeof(double) * (size_t)n * nw, cudaMemcpyHostToDevice, st));
        *dst = buf;
        return 0;
    }
    // counting sort of one sample, served from the sorted-sample cache when the sample is `stable` (see stage_coords)
    int sorted(const FineGrid &g, const double *const *d, int64_t ds, int64_t n, const double *dw, int nw, bool perm,
               double pad, bool stable, SortedSample &out)
    {
        if (!stable || !g_cache_on || (dw && !(flags & HTB_FLAG_DEVICE_INPUT)) || getenv("HTB_NO_SORT_CACHE"))
            return htb_sort_sample(st, ws, g, d, ds, n, dw, nw, perm, pad, out, &launches);
        if (!g_sort_cache) g_sort_cache = new std::vector<SortEnt *>();
        for (SortEnt *e : *g_sort_cache) {
            bool hit = e->stride == ds && e->n == n && e->w == dw && e->nw == (dw ? nw : 0) && e->perm == perm && e->pad == pad &&
                       e->g.dim == g.dim;
            for (int k = 0; k < g.dim && hit; ++k)
                hit = e->src[k] == d[k] && e->g.nd[k] == g.nd[k] && e->g.m[k] == g.m[k] && e->g.cs[k] == g.cs[k] &&
                      e->g.period[k] == g.period[k];
            if (!hit || !e->ready) continue;
            if (e->ws.st != st) {
                HTB_CUDA(cudaStreamWaitEvent(st, e->ready, 0));
                if (std::find(e->users.begin(), e->users.end(), st) == e->users.end()) e->users.push_back(st);
            }
            out = e->s;
            return 0;
        }
        SortEnt *e = new SortEnt();
        for (int k = 0; k < 3; ++k) e->src[k] = k < g.dim ? d[k] : nullptr;
        e->stride = ds; e->n = n; e->w = dw; e->nw = dw ? nw : 0; e->perm = perm; e->pad = pad; e->g = g;
        e->ws.st = st;
        e->ready = nullptr;
        g_sort_cache->push_back(e);                       // (owned by the cache from here on: freed by htb_cache_end)
        if (htb_sort_sample(st, e->ws, g, d, ds, n, dw, nw, perm, pad, e->s, &launches)) return 1;
        HTB_CUDA(cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming));
        HTB_CUDA(cudaEventRecord(e->ready, st));
        out = e->s;
        return 0;
    }
    // full set-up: upload, sort both samples, tile list, walker geometry
    int setup(const htb_mesh_geom *g, int sphere, bool allow_sym,
              const double *const *c1, int64_t stride1, int64_t n1, const double *w1,
              const double *const *c2, int64_t stride2, int64_t n2, const double *w2, int nw,
              bool perm1, int64_t first_cell1, int64_t last_cell1, uint32_t fl,
              int tile = HTB_TILE, double sentinel = 1.0e150, int refine_mode = 0)
    {
        flags = fl;
        first_cell = first_cell1;
        last_cell = last_cell1;
        const int dim = g->ndim;
        if (dim != 2 && dim != 3) { htb_set_error("mesh->ndim must be 2 or 3"); return 1; }
        if (n1 < 0 || n2 < 0 || n1 > 2000000000LL || n2 > 2000000000LL) { htb_set_error("sample sizes must be in [0, 2e9]"); return 1; }
        for (int d = 0; d < dim; ++d) {
            if (g->ndivs1[d] < 1 || g->ndivs2[d] < g->ndivs1[d] || g->ndivs2[d] % g->ndivs1[d] != 0 || g->cover[d] < 0 ||
                !(g->period[d] > 0) || !(g->cell1_size[d] > 0) || !(g->cell2_size[d] > 0) || !(g->search[d] >= 0)) {
                htb_set_error("inconsistent mesh geometry in dimension %d", d);
                return 1;
            }
        }
        // ---- inputs
        const double *d1[3] = {nullptr, nullptr, nullptr}, *d2[3] = {nullptr, nullptr, nullptr};
        int64_t ds1 = 1, ds2 = 1;
        const double *dw1 = nullptr, *dw2 = nullptr;
        bool same = (n1 == n2 && stride1 == stride2);
        for (int d = 0; d < dim && same; ++d) same = (c1[d] == c2[d]);
        bool stable1 = false, stable2 = false;
        if (stage_coords(c1, dim, stride1, n1, d1, &ds1, (fl & HTB_FLAG_CACHE_SAMPLE1) != 0, &stable1)) return 1;
        if (same) { for (int d = 0; d < dim; ++d) d2[d] = d1[d]; ds2 = ds1; stable2 = stable1; }
        else if (stage_coords(c2, dim, stride2, n2, d2, &ds2, (fl & HTB_FLAG_CACHE_SAMPLE2) != 0, &stable2)) return 1;
        if (stage_rows(w1, n1, nw, &dw1)) return 1;
        if (w2 == w1 && same) dw2 = dw1;
        else if (stage_rows(w2, n2, nw, &dw2)) return 1;
        if (mark(1)) return 1;
        // ---- K1
        choose_refinement(g, n1, n2, tile, m1, m2, refine_mode);
        // Symmetric auto-correlation (count each zero-shift unordered pair once, weight 2) needs the two
        // samples to be the SAME sorted arrays and the reference window to be symmetric (mesh1 == mesh2).
        bool sym = allow_sym && same && !(fl & HTB_FLAG_NO_SYM) && !getenv("HTB_NO_SYM");
        for (int d = 0; d < dim && sym; ++d) sym = (g->ndivs1[d] == g->ndivs2[d]);
        {
            // A partial cell range must return what the reference engine returns for that cell1_tuple: pairs (i, j)
            // with i inside the range and j anywhere.  The symmetric shortcut attributes an unordered pair to the
            // point with the smaller sorted index, so its per-range counts are only right when the caller sums a
            // partition of the cells (device-side shards, HTB_FLAG_PARTITION_SUM).
            int64_t nc1all = 1;
            for (int d = 0; d < dim; ++d) nc1all *= g->ndivs1[d];
            const bool partial = first_cell1 > 0 || last_cell1 < nc1all;
            if (partial && g_shard_world == 1 && !(fl & HTB_FLAG_PARTITION_SUM)) sym = false;
        }
        if (sym) for (int d = 0; d < dim; ++d) m1[d] = m2[d];
        const FineGrid g1 = make_grid(g, 0, m1), g2 = make_grid(g, 1, m2);
        G.sym = sym ? 1 : 0;
        G.tile = tile;
        G.sentinel = sentinel;
        {
            // a tile may straddle K reference cells along the fast dimension if the union of their windows
            // (K * per + 2 * cover mesh2 cells) cannot reach the same mesh2 cell twice
            const int Fd = dim - 1;
            const int per = g->ndivs2[Fd] / g->ndivs1[Fd];
            int K = (g->ndivs2[Fd] - 2 * g->cover[Fd]) / per;
            if (K > g->ndivs1[Fd]) K = g->ndivs1[Fd];
            // HTB_FLAG_NO_CULL promises exactly the reference's cell windows: no union windows then
            if (K < 1 || (fl & HTB_FLAG_NO_CULL) || getenv("HTB_NO_STRADDLE")) K = 1;
            G.maxspan = K;
            G.maxfine = m1[Fd];           // a tile is never longer than one reference cell
            G.cs1f = g->cell1_size[Fd];
            // slicing a tile only pays if a slice still holds enough pair evaluations to amortise the per-item
            // latencies (tile fetch, point loads, first TMA): about 2e5 pairs per item
            double est = (double)tile * (double)(n2 > 0 ? n2 : 1);
            for (int d = 0; d < dim; ++d) est *= std::min(1.0, 2.6 * g->search[d] / g->period[d]);
            G.maxslices = clampi(est / 2.0e5, 1, 16);
            if (const char *e = getenv("HTB_MAXSLICES")) { const int v = atoi(e); if (v >= 1 && v <= 64) G.maxslices = v; }
            G.early_exit = ((fl & HTB_FLAG_EARLY_EXIT) || getenv("HTB_EARLY_EXIT")) ? 1 : 0;
            G.tail_eighths = 0;          // measured (profiles/r02_tail_sweep.txt): slices cost more than the idle tail they remove
            if (const char *e = getenv("HTB_TAIL_EIGHTHS")) { const int v = atoi(e); if (v >= 0 && v <= 512) G.tail_eighths = v; }
            G.items_per_warp = HTB_ITEMS_PER_WARP;
            if (const char *e = getenv("HTB_ITEMS_PER_WARP")) { const int v = atoi(e); if (v >= 1 && v <= 1024) G.items_per_warp = v; }
        }
        // ---- walker geometry
        G.dim = dim;
        G.pbc = g->pbc ? 1 : 0;
        G.sphere = sphere;
        G.nocull = (fl & HTB_FLAG_NO_CULL) ? 1 : 0;
        double rslow = 0.0;
        for (int d = 0; d < 3; ++d) {
            const bool on = d < dim;
            G.nd1[d] = on ? g->ndivs1[d] : 1;
            G.nd2[d] = on ? g->ndivs2[d] : 1;
            G.per[d] = on ? g->ndivs2[d] / g->ndivs1[d] : 1;
            G.cover[d] = on ? g->cover[d] : 0;
            G.m1[d] = on ? m1[d] : 1;
            G.m2[d] = on ? m2[d] : 1;
            G.nf1[d] = G.nd1[d] * G.m1[d];
            G.nf2[d] = G.nd2[d] * G.m2[d];
            G.period[d] = on ? g->period[d] : 1.0;
            G.h2[d] = on ? g2.h[d] : 1.0;
            G.slop[d] = on ? 1e-9 * g->period[d] : 0.0;
            G.reach[d] = on ? g->search[d] : 0.0;
            if (on && d < dim - 1 && g->search[d] > rslow) rslow = g->search[d];
        }
        if (sphere && g->search[dim - 1] > rslow) rslow = g->search[dim - 1];
        G.r2slow = rslow * rslow * (1.0 + 1e-9);
        // ---- K1: counting sort of both samples
        if (g_shard_world > 1) {
            // One rank of a sharded call: cell ids and per-cell counts of BOTH samples first; from them this rank's
            // share of [first_cell1, last_cell1), cut by predicted work on the device (no host sync); then only the
            // points inside the x-layers that share can touch are put in order (the rest is never moved).
            const bool window = !getenv("HTB_NO_WINDOW_SORT");
            const bool cacheable = g_cache_on && stable1 && stable2 && !getenv("HTB_NO_SORT_CACHE") &&
                                   ((!dw1 && !dw2) || (flags & HTB_FLAG_DEVICE_INPUT));
            PrepKey key;
            memset(&key, 0, sizeof(key));
            for (int d = 0; d < dim; ++d) { key.d1[d] = d1[d]; key.d2[d] = d2[d]; key.m1[d] = m1[d]; key.m2[d] = m2[d]; }
            key.dw1 = dw1; key.dw2 = dw2; key.ds1 = ds1; key.ds2 = ds2; key.n1 = n1; key.n2 = n2;
            key.first = first_cell1; key.last = last_cell1; key.geom = *g; key.geom.reserved = 0;
            key.nw = nw; key.perm1 = perm1 ? 1 : 0; key.sym = sym ? 1 : 0; key.rank = g_shard_rank; key.world = g_shard_world;
            key.window = window ? 1 : 0; key.pad = sentinel;
            PrepEnt *pe = nullptr;
            if (cacheable) {
                if (!g_prep_cache) g_prep_cache = new std::vector<PrepEnt *>();
                for (PrepEnt *e : *g_prep_cache) if (e->ready && memcmp(&e->key, &key, sizeof(key)) == 0) { pe = e; break; }
            }
            if (pe) {
                if (pe->ws.st != st) {
                    HTB_CUDA(cudaStreamWaitEvent(st, pe->ready, 0));
                    if (std::find(pe->users.begin(), pe->users.end(), st) == pe->users.end()) pe->users.push_back(st);
                }
                s1 = pe->s1; s2 = pe->s2; range_dev = pe->range_dev; work_dev = pe->work_dev; nc1 = pe->nc1;
            } else {
                PrepEnt *ne = nullptr;
                if (cacheable) {
                    ne = new PrepEnt();
                    ne->key = key; ne->ws.st = st; ne->ready = nullptr; ne->range_dev = nullptr; ne->work_dev = nullptr; ne->nc1 = 0;
                    g_prep_cache->push_back(ne);             // (owned by the cache from here on: freed by htb_cache_end)
                }
                Workspace &w = ne ? ne->ws : ws;
                if (!sym && htb_sort_begin(st, w, g1, d1, ds1, n1, dw1, nw, perm1, s1, &launches)) return 1;
                if (htb_sort_begin(st, w, g2, d2, ds2, n2, dw2, nw, sym ? perm1 : false, s2, &launches)) return 1;
                double *balance_dev = nullptr;
                if (htb_reference_work(st, w, G, sym ? s2 : s1, s2, &work_dev, &balance_dev, &nc1, &launches, true)) return 1;
                if (w.alloc((void **)&range_dev, 2 * sizeof(long long))) return 1;
                const int64_t lo = first_cell1 < 0 ? 0 : first_cell1, hi = last_cell1 > nc1 ? nc1 : last_cell1;
                if (htb_shard_range(st, balance_dev, lo, hi, g_shard_rank, g_shard_world, range_dev, &launches)) return 1;
                int *xwin = nullptr;
                if (w.alloc((void **)&xwin, 4 * sizeof(int))) return 1;
                if (htb_shard_windows(st, range_dev, G, xwin, &launches)) return 1;
                if (!sym && htb_sort_finish(st, w, d1, ds1, dw1, nw, sentinel, window ? xwin : nullptr, s1, &launches)) return 1;
                if (htb_sort_finish(st, w, d2, ds2, dw2, nw, sentinel, window ? xwin + 2 : nullptr, s2, &launches)) return 1;
                if (sym) s1 = s2;
                if (ne) {
                    ne->s1 = s1; ne->s2 = s2; ne->range_dev = range_dev; ne->work_dev = work_dev; ne->nc1 = nc1;
                    HTB_CUDA(cudaEventCreateWithFlags(&ne->ready, cudaEventDisableTiming));
                    HTB_CUDA(cudaEventRecord(ne->ready, st));
                }
            }
        } else if (sym) {
            if (sorted(g2, d2, ds2, n2, dw2, nw, perm1, sentinel, stable2, s2)) return 1;
            s1 = s2;
        } else {
            if (sorted(g1, d1, ds1, n1, dw1, nw, perm1, sentinel, stable1, s1)) return 1;
            if (sorted(g2, d2, ds2, n2, dw2, nw, false, sentinel, stable2, s2)) return 1;
        }
        if (fl & HTB_FLAG_PREPARE) { prepared = true; return 0; }
        // ---- tiles + counters
        uint2 *tiles = nullptr;
        uint32_t *ntiles_dev = nullptr;
        if (htb_build_tiles(st, ws, G, s1, first_cell1, last_cell1, range_dev, &tiles, &ntiles_dev, &max_tiles, &launches)) return 1;
        // three cache lines: the work-item counter every warp bumps, the statistics, the redo-queue counters the idle warps
        // poll at the end of the launch (on one line the pollers slowed the item fetches of the warps still working)
        if (ws.alloc((void **)&ctr, 512)) return 1;
        HTB_CUDA(cudaMemsetAsync(ctr, 0, 512, st));
        for (int d = 0; d < 3; ++d) { A.c1[d] = s1.c[d]; A.c2[d] = s2.c[d]; }
        A.off1 = s1.off; A.off2 = s2.off;
        A.pay1 = s1.w; A.pay2 = s2.w; A.nw = nw;
        A.flags1 = s1.flags; A.flags2 = s2.flags;
        A.tiles = tiles; A.ntiles_dev = ntiles_dev;
        A.tile_counter = ctr;
        A.tiles_redone = ctr + 32;
        A.pairs_evaluated = (unsigned long long *)(ctr + 34);
        A.redo_ctr = ctr + 64;
        A.redo_cap = getenv("HTB_REDO_INPLACE") ? 0u : (1u << 16);
        A.redo_ent = nullptr;
        if (A.redo_cap && ws.alloc((void **)&A.redo_ent, sizeof(uint2) * (size_t)A.redo_cap)) return 1;
        if (mark(2)) return 1;
        return 0;
    }
    bool count_marked = false;
    // end of the counting kernels (output copies that follow are not part of ms_count)
    int mark_count_end()
    {
        if (!count_marked) { if (mark(3)) return 1; count_marked = true; }
        return 0;
    }
    int finish(htb_stats *stats, int path)
    {
        if (async) {
            // results stay on the device; the caller synchronises once for the whole statistic (no stats)
            if (stats) memset(stats, 0, sizeof(*stats));
            return 0;
        }
        if (mark_count_end()) return 1;
        struct { unsigned int h[4]; uint32_t ntiles; uint32_t pad; double wref; } rb = {{0, 0, 0, 0}, 0, 0, 0.0};   // h[0] tiles redone, h[2..3] pairs evaluated (u64)
        double *wsum = nullptr;
        if (stats) {
            // W_ref of this call's (this rank's) cell range, summed on the device: one 8-byte read-back instead of the
            // per-cell work vector
            if (!work_dev && htb_reference_work(st, ws, G, s1, s2, &work_dev, nullptr, &nc1, &launches)) return 1;
            if (ws.alloc((void **)&wsum, sizeof(double))) return 1;
            k_sum_range<<<1, 256, 0, st>>>(work_dev, (long long)nc1, (long long)first_cell, (long long)last_cell, range_dev, wsum);
            launches += 1;
            HTB_CUDA(cudaGetLastError());
            HTB_CUDA(cudaMemcpyAsync(&rb.wref, wsum, sizeof(double), cudaMemcpyDeviceToHost, st));
            HTB_CUDA(cudaMemcpyAsync(rb.h, ctr + 32, sizeof(rb.h), cudaMemcpyDeviceToHost, st));
            HTB_CUDA(cudaMemcpyAsync(&rb.ntiles, A.ntiles_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        if (mark(4)) return 1;
        HTB_CUDA(cudaStreamSynchronize(st));
        if (stats) {
            memset(stats, 0, sizeof(*stats));
            unsigned long long pe;
            memcpy(&pe, &rb.h[2], sizeof(pe));
            stats->pairs_evaluated = (double)pe;
            stats->pairs_reference = rb.wref;
            cudaEventElapsedTime(&stats->ms_h2d, ev[0], ev[1]);
            cudaEventElapsedTime(&stats->ms_mesh, ev[1], ev[2]);
            cudaEventElapsedTime(&stats->ms_count, ev[2], ev[3]);
            cudaEventElapsedTime(&stats->ms_total, ev[0], ev[4]);
            stats->kernel_launches = launches;
            stats->tiles = (int32_t)rb.ntiles;
            stats->tiles_redone = (int32_t)rb.h[0];
            for (int d = 0; d < 3; ++d) { stats->refine1[d] = m1[d]; stats->refine2[d] = m2[d]; }
            stats->path = path;
        }
        return 0;
    }
    // hand n 8-byte result words to the caller: device -> host, or device -> device for an asynchronous call
    int deliver(void *out, const void *src_dev, size_t n)
    {
        if (n == 0) return 0;
        HTB_CUDA(cudaMemcpyAsync(out, src_dev, 8 * n, async ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        return 0;
    }
    // device buffer a finishing kernel writes the caller's n result words to: the caller's own (device) array for an
    // asynchronous call, else a workspace block that out_fetch() copies to the host array
    int out_buffer(void *user_out, size_t n, void **dev)
    {
        if (async) { *dev = user_out; return 0; }
        return ws.alloc(dev, 8 * (n ? n : 1));
    }
    int out_fetch(void *user_out, const void *dev, size_t n)
    {
        if (async || n == 0) return 0;
        HTB_CUDA(cudaMemcpyAsync(user_out, dev, 8 * n, cudaMemcpyDeviceToHost, st));
        return 0;
    }
};

// ------------------------------------------------------------------ finishing kernels (tiny, one block)
// differential histogram (lowest satisfied edge per axis) -> the reference's cumulative counts: inclusive prefix sums
// along the second axis, then along the first (npairs_s_mu_engine.pyx:232-234; additions only, so float sums lose nothing)
template <class T>
__global__ void __launch_bounds__(128) k_prefix2d(const T *__restrict__ diff, int n0, int n1, T *__restrict__ out)
{
    __shared__ T sh[4096];
    const int n = n0 * n1, t = threadIdx.x;
    for (int i = t; i < n; i += 128) sh[i] = diff[i];
    __syncthreads();
    for (int k = t; k < n0; k += 128) { T run = 0; for (int g = 0; g < n1; ++g) { run += sh[k * n1 + g]; sh[k * n1 + g] = run; } }
    __syncthreads();
    for (int g = t; g < n1; g += 128) { T run = 0; for (int k = 0; k < n0; ++k) { run += sh[k * n1 + g]; sh[k * n1 + g] = run; } }
    __syncthreads();
    for (int i = t; i < n; i += 128) out[i] = sh[i];
}
// FastXYZ: device layout {top pi column [nrp], lower pi column [nrp]} -> row-major (nrp, npi)
__global__ void k_xyz_columns(const unsigned long long *__restrict__ cols, int nrp, int npi, unsigned long long *__restrict__ out)
{
    const int k = threadIdx.x;
    if (k >= nrp) return;
    out[k * npi + (npi - 1)] = cols[k];
    if (npi == 2) out[k * npi] = cols[nrp + k];
}
// MarkedQ: differential level sums [HTB_NBF] + the sum over all in-range pairs [1] -> cumulative sums [nb]
// (marked_npairs_3d_engine.pyx:212-216: a pair of level s counts for every edge >= s)
__global__ void k_markedq_cumulate(const double *__restrict__ h, int nb, double *__restrict__ out)
{
    if (threadIdx.x) return;
    const int pad = HTB_NBF - nb;
    double run = 0.0;
    for (int s = 0; s < HTB_NBF; ++s) {
        run += h[s];
        if (s >= pad && s < HTB_NBF - 1) out[s - pad] = run;
    }
    out[nb - 1] = h[HTB_NBF];
}

static bool finite_all(const double *v, int n)
{
    for (int i = 0; i < n; ++i) if (!std::isfinite(v[i])) return false;
    return true;
}

static unsigned long long dbits(double v)
{
    unsigned long long b;
    memcpy(&b, &v, sizeof(b));
    return b;
}

static int upload(Call &c, const void *host, size_t bytes, void **dev)
{
    if (c.ws.alloc(dev, bytes)) return 1;
    HTB_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, c.st));
    return 0;
}


// Parameters of the fast queue kernels (Fast3 / FastXYZ) for the squared edges rsq[0..nb), or false if they are
// not eligible.  The 32-bit keys are (bits(dsq) >> 26) relative to the top squared edge, so every separation a
// tile can meet and every non-zero edge must lie within 2^+-31 of it:
//   above: points inside the box are at most 2L apart per dimension after the periodic shift and the sentinel
//          of unused lanes / pad entries sits at 8L  ->  dsq < 256 Lmax^2;
//   below: smaller separations (and exact zeros) are caught per group by the kernel (Hwin) and evaluated
//          exactly; edges below the window send the call to the generic kernel.
static bool fast3_params(const htb_mesh_geom *mesh, const double *bins, const double *rsq, int nb, uint32_t flags,
                         Fast3Params *out, double *lmax_out)
{
    bool fast = !(flags & HTB_FLAG_GENERIC) && nb >= 1 && nb <= HTB_NBF && finite_all(bins, nb);
    for (int k = 0; k + 1 < nb && fast; ++k) fast = bins[k] >= 0.0 && rsq[k] <= rsq[k + 1];
    if (fast) fast = bins[nb - 1] >= 0.0 && rsq[nb - 1] > 1e-290 && rsq[nb - 1] < 1e290;
    double lmax = 0.0;
    for (int d = 0; d < mesh->ndim; ++d) if (mesh->period[d] > lmax) lmax = mesh->period[d];
    if (fast) fast = std::isfinite(lmax) && lmax > 0.0 && lmax < 1e140;
    *lmax_out = lmax;
    if (!fast) return false;
    Fast3Params fp{};
    const long long Kt = (long long)(dbits(rsq[nb - 1]) >> 26);
    const long long Kmax = (long long)(dbits(256.0 * lmax * lmax) >> 26);
    const long long lim = 31LL << 26;
    if (Kmax - Kt >= lim) return false;
    fp.nb = nb;
    fp.nbias = (int)(unsigned)(0ULL - (unsigned long long)Kt);
    fp.Hwin = (int)(dbits(rsq[nb - 1]) >> 32) - (31 << 20);
    fp.Hz0 = -1;
    const int pad = HTB_NBF - nb;
    for (int s = 0; s < HTB_NBF; ++s) {
        fp.F[s] = INT32_MIN; fp.E[s] = 0;
        if (s < pad) continue;
        const double e = rsq[s - pad];
        fp.E[s] = dbits(e);
        if (e == 0.0) continue;                      // only exact zeros satisfy it: always decided exactly
        const long long d = (long long)(dbits(e) >> 26) - Kt;
        if (d <= -lim) return false;
        fp.F[s] = (int)d;
    }
    fp.E_top = dbits(rsq[nb - 1]);
    *out = fp;
    return true;
}

// BinQ (binq.cu) serves any set of finite, non-decreasing squared edges on one or two bin axes: the differential
// histogram it returns (lowest satisfied edge per axis) becomes the reference's cumulative counts by inclusive
// prefix sums (for monotone edges the reference's top-down scans stop exactly at the lowest satisfied edge,
// npairs_3d_engine.pyx:178-182, npairs_xy_z_engine.pyx:186-194, npairs_s_mu_engine.pyx:218-234).
static bool binq_ok(const double *e0, int n0, const double *e1, int n1, uint32_t flags)
{
    if ((flags & HTB_FLAG_GENERIC) || getenv("HTB_NO_BINQ")) return false;
    if (n0 < 1 || n1 < 1 || n0 > 127 || n1 > 127 || (long long)n0 * n1 > 4096) return false;
    if (!finite_all(e0, n0) || (e1 && !finite_all(e1, n1))) return false;
    for (int k = 0; k + 1 < n0; ++k) if (!(e0[k] <= e0[k + 1])) return false;
    for (int k = 0; e1 && k + 1 < n1; ++k) if (!(e1[k] <= e1[k + 1])) return false;
    return e0[0] >= 0.0 && (!e1 || e1[0] >= 0.0);
}

// The reference's top-down scans stop at the first failing edge (npairs_3d_engine.pyx:178-182), so edge k admits a
// pair iff the pair satisfies edges k..n-1: the counts are those of the suffix minima of the squared edges, which are
// non-decreasing by construction.
static void suffix_min(std::vector<double> &e, int first, int n)
{
    for (int k = n - 2; k >= 0; --k) if (e[(size_t)first + k + 1] < e[(size_t)first + k]) e[(size_t)first + k] = e[(size_t)first + k + 1];
}

// edges + lookup tables of a BinQ launch (uploaded with the call's workspace)
static int binq_prepare(Call &c, const double *e0, int n0, const double *e1, int n1, BinQParams *out)
{
    std::vector<unsigned long long> eb((size_t)n0 + n1, 0ULL);
    for (int k = 0; k < n0; ++k) eb[k] = dbits(e0[k] + 0.0);
    for (int k = 0; e1 && k < n1; ++k) eb[(size_t)n0 + k] = dbits(e1[k] + 0.0);
    // Lookup tables: key = bits >> S keeps the exponent and up to 8 mantissa bits.  lut[key - kmin] = (i << 1) | a with i
    // the first edge whose key is >= key (every earlier edge is certainly below the value) and a = 1 if some edge has
    // exactly this key (only then the kernel needs exact compares: the value and that edge share a table cell).
    // kmin is the key of the smallest non-zero edge; values with a smaller key lie below every non-zero edge (index =
    // number of zero edges, or 0 for an exact zero).  S grows until the table has at most 2048 entries.
    BinQParams bp{};
    std::vector<unsigned char> lut[2];
    for (int a = 0; a < 2; ++a) {
        const int n = a ? n1 : n0;
        const unsigned long long *b = eb.data() + (a ? n0 : 0);
        int z = 0;
        while (z < n - 1 && b[z] == 0ULL) ++z;
        int S = 44;
        while (((b[n - 1] >> S) - (b[z] >> S) + 1ULL) > 2040ULL) ++S;
        const unsigned long long kmin = b[z] >> S;
        const int Tin = (int)((b[n - 1] >> S) - kmin + 1ULL);
        // two guard cells: entry 0 = every key below kmin (the value lies below every non-zero edge: edge `z`, or - if
        // there are zero edges - possibly edge 0: compare), entry Tin + 1 = every key above the top edge's (index n)
        const int T = Tin + 2;
        lut[a].assign(((size_t)T + 7) & ~(size_t)7, (unsigned char)(n << 1));
        lut[a][0] = (b[z] == 0ULL) ? (unsigned char)((n << 1)) : (unsigned char)(z > 0 ? 1 : 0);   // z > 0: start at 0, compare
        int i = 0;
        for (int t = 0; t < Tin; ++t) {
            while (i < n && (b[i] >> S) < kmin + (unsigned long long)t) ++i;
            const bool amb = i < n && (b[i] >> S) == kmin + (unsigned long long)t;
            lut[a][(size_t)t + 1] = (unsigned char)((i << 1) | (amb ? 1 : 0));
        }
        lut[a][(size_t)Tin + 1] = (unsigned char)(n << 1);
        if (b[z] == 0ULL) lut[a][1] = 1;               // all edges zero: key 0 holds them all (start at 0, compare)
        bp.S[a] = S; bp.T[a] = T; bp.kmin[a] = (unsigned)kmin; bp.nzero[a] = z;
    }
    const size_t ne = eb.size();
    eb.resize(ne + (lut[0].size() + lut[1].size()) / 8);
    memcpy((unsigned char *)(eb.data() + ne), lut[0].data(), lut[0].size());
    memcpy((unsigned char *)(eb.data() + ne) + lut[0].size(), lut[1].data(), lut[1].size());
    void *edev = nullptr;
    if (upload(c, eb.data(), sizeof(unsigned long long) * eb.size(), &edev)) return 1;
    bp.n0 = n0; bp.n1 = n1;
    bp.H0 = (int)(eb[(size_t)n0 - 1] >> 32);
    bp.H1 = e1 ? (int)(eb[(size_t)n0 + n1 - 1] >> 32) : 0;
    bp.edges = (const unsigned long long *)edev;
    *out = bp;
    return 0;
}

static int run_binq(Call &c, int kind, const double *e0, int n0, const double *e1, int n1, int64_t *counts_out, htb_stats *stats)
{
    BinQParams bp{};
    if (binq_prepare(c, e0, n0, e1, n1, &bp)) return 1;
    const int nh = n0 * n1;
    unsigned long long *counts_dev = nullptr, *cum = nullptr;
    if (c.ws.alloc((void **)&counts_dev, sizeof(unsigned long long) * (size_t)nh)) return 1;
    HTB_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(unsigned long long) * (size_t)nh, c.st));
    bp.counts = counts_dev;
    if (htb_launch_binq(c.st, kind, 0, c.G, c.A, bp, &c.launches)) return 1;
    if (c.mark_count_end()) return 1;
    if (c.out_buffer(counts_out, (size_t)nh, (void **)&cum)) return 1;
    k_prefix2d<unsigned long long><<<1, 128, 0, c.st>>>(counts_dev, n0, n1, cum);
    c.launches += 1;
    HTB_CUDA(cudaGetLastError());
    if (c.out_fetch(counts_out, cum, (size_t)nh)) return 1;
    return c.finish(stats, 3);
}

// weighted sums (MODE 1): differential float histogram -> cumulative sums.  The differential cells are summed in
// double precision; the reference adds every pair's weight to each of its cumulative cells in turn
// (marked_npairs_xy_z_engine.pyx:217-225), so the two agree to rounding (1e-16 relative per term).
static int run_binq_weighted(Call &c, int kind, int nw, int wfunc, const double *e0, int n0, const double *e1, int n1,
                             double *counts_out, htb_stats *stats, int mode = 1)
{
    BinQParams bp{};
    if (binq_prepare(c, e0, n0, e1, n1, &bp)) return 1;
    const int nh = n0 * n1;
    double *sums_dev = nullptr, *cum = nullptr;
    if (c.ws.alloc((void **)&sums_dev, sizeof(double) * (size_t)nh)) return 1;
    HTB_CUDA(cudaMemsetAsync(sums_dev, 0, sizeof(double) * (size_t)nh, c.st));
    bp.fcounts = sums_dev;
    bp.nw = nw; bp.wfunc = wfunc;
    if (htb_launch_binq(c.st, kind, mode, c.G, c.A, bp, &c.launches)) return 1;
    if (c.mark_count_end()) return 1;
    if (c.out_buffer(counts_out, (size_t)nh, (void **)&cum)) return 1;
    k_prefix2d<double><<<1, 128, 0, c.st>>>(sums_dev, n0, n1, cum);
    c.launches += 1;
    HTB_CUDA(cudaGetLastError());
    if (c.out_fetch(counts_out, cum, (size_t)nh)) return 1;
    return c.finish(stats, 3);
}

// ------------------------------------------------------------------ npairs_3d
extern "C" int htb_npairs_3d_engine(const htb_mesh_geom *mesh,
                                    const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                    const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                    const double *rbins, int32_t nb, int64_t first_cell1, int64_t last_cell1,
                                    int64_t *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rbins || !counts_out || nb < 1) { htb_set_error("htb_npairs_3d_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_npairs_3d_engine needs a 3-d mesh"); return 1; }
    std::vector<double> rsq((size_t)nb);
    for (int k = 0; k < nb; ++k) rsq[k] = rbins[k] * rbins[k];
    double lmax = 0.0;
    Fast3Params fp{};
    bool fast = fast3_params(mesh, rbins, rsq.data(), nb, flags, &fp, &lmax);
    Call c;
    if (c.begin(flags)) return 1;
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    if (c.setup(mesh, 1, true, c1, stride1, n1, nullptr, c2, stride2, n2, nullptr, 0, false, first_cell1, last_cell1, flags,
                fast ? 32 * htb_fast3_ppl() : HTB_TILE, fast ? 8.0 * lmax : 1.0e150)) return 1;
    if (c.prepared) return 0;
    unsigned long long *counts_dev = nullptr;
    if (c.ws.alloc((void **)&counts_dev, sizeof(unsigned long long) * (size_t)nb)) return 1;
    HTB_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(unsigned long long) * (size_t)nb, c.st));
    if (fast) {
        fp.counts = counts_dev;
        // symmetric counts run two weight passes per work item: half as many, larger items pay off when a rank's shard
        // leaves few tiles (8-GPU RR shard of the bench: 9.05 -> 8.55 ms; no change on one GPU)
        if (c.G.sym && !getenv("HTB_ITEMS_PER_WARP")) c.G.items_per_warp = 16;
        if (htb_launch_fast3(c.st, c.G, c.A, fp, &c.launches)) return 1;
    } else if (suffix_min(rsq, 0, nb), binq_ok(rsq.data(), nb, nullptr, 1, flags)) {
        return run_binq(c, 0, rsq.data(), nb, nullptr, 1, counts_out, stats);
    } else {
        for (int k = 0; k < nb; ++k) rsq[k] = rbins[k] * rbins[k];
        GenParams gp{};
        gp.n0 = nb; gp.n1 = 1; gp.nhist = nb;
        void *e0 = nullptr;
        if (upload(c, rsq.data(), sizeof(double) * (size_t)nb, &e0)) return 1;
        gp.e0 = (const double *)e0;
        gp.counts = counts_dev;
        if (htb_launch_gen(c.st, 0, c.G, c.A, gp, &c.launches)) return 1;
    }
    if (c.mark_count_end()) return 1;
    if (c.deliver(counts_out, counts_dev, (size_t)nb)) return 1;
    return c.finish(stats, fast ? 1 : 0);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ npairs_xy_z
extern "C" int htb_npairs_xy_z_engine(const htb_mesh_geom *mesh,
                                      const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                      const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                      const double *rp_bins, int32_t nrp, const double *pi_bins, int32_t npi,
                                      int64_t first_cell1, int64_t last_cell1,
                                      int64_t *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rp_bins || !pi_bins || !counts_out || nrp < 1 || npi < 1) { htb_set_error("htb_npairs_xy_z_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_npairs_xy_z_engine needs a 3-d mesh"); return 1; }
    std::vector<double> e((size_t)nrp + npi);
    for (int k = 0; k < nrp; ++k) e[k] = rp_bins[k] * rp_bins[k];
    for (int k = 0; k < npi; ++k) e[nrp + k] = pi_bins[k] * pi_bins[k];
    // Fast path (FastXYZ): the rp edges as in npairs_3d, one pi edge, or two when the lower one is so small that
    // pairs inside it are rare (wp's pi_bins = [0, pi_max]): those are found by the kernel and counted exactly.
    double lmax = 0.0;
    Fast3Params fp{};
    bool fast = (npi == 1 || npi == 2) && finite_all(pi_bins, npi) && pi_bins[npi - 1] >= 0.0 &&
                fast3_params(mesh, rp_bins, e.data(), nrp, flags, &fp, &lmax);
    if (fast && npi == 2) fast = pi_bins[0] >= 0.0 && e[nrp] <= 1e-8 * e[nrp + 1];
    const int nh = nrp * npi;
    Call c;
    if (c.begin(flags)) return 1;
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    if (c.setup(mesh, 0, true, c1, stride1, n1, nullptr, c2, stride2, n2, nullptr, 0, false, first_cell1, last_cell1, flags,
                fast ? 32 * htb_fast3_ppl() : HTB_TILE, fast ? 8.0 * lmax : 1.0e150)) return 1;
    if (c.prepared) return 0;
    unsigned long long *counts_dev = nullptr;
    if (c.ws.alloc((void **)&counts_dev, sizeof(unsigned long long) * (size_t)nh)) return 1;
    HTB_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(unsigned long long) * (size_t)nh, c.st));
    if (fast) {
        // device layout: column of the top pi edge first, then the lower pi edge column; interleaved on the host
        fp.pi_top_sq = e[nrp + npi - 1];
        fp.counts = counts_dev;
        if (npi == 2) {
            fp.Epi0 = dbits(e[nrp]);
            fp.Hz0 = (int)(dbits(e[nrp]) >> 32);
            fp.counts0 = counts_dev + nrp;
        }
        if (htb_launch_fastxyz(c.st, c.G, c.A, fp, &c.launches)) return 1;
        if (c.mark_count_end()) return 1;
        unsigned long long *res = nullptr;
        if (c.out_buffer(counts_out, (size_t)nh, (void **)&res)) return 1;
        k_xyz_columns<<<1, 32, 0, c.st>>>(counts_dev, nrp, npi, res);        // nrp <= HTB_NBF
        c.launches += 1;
        HTB_CUDA(cudaGetLastError());
        if (c.out_fetch(counts_out, res, (size_t)nh)) return 1;
        return c.finish(stats, 1);
    }
    {
        std::vector<double> em(e);
        suffix_min(em, 0, nrp);
        suffix_min(em, nrp, npi);
        if (binq_ok(em.data(), nrp, em.data() + nrp, npi, flags))
            return run_binq(c, 1, em.data(), nrp, em.data() + nrp, npi, counts_out, stats);
    }
    void *edev = nullptr;
    if (upload(c, e.data(), sizeof(double) * e.size(), &edev)) return 1;
    GenParams gp{};
    gp.n0 = nrp; gp.n1 = npi; gp.nhist = nh;
    gp.e0 = (const double *)edev; gp.e1 = (const double *)edev + nrp;
    gp.counts = counts_dev;
    if (htb_launch_gen(c.st, 1, c.G, c.A, gp, &c.launches)) return 1;
    if (c.mark_count_end()) return 1;
    if (c.deliver(counts_out, counts_dev, (size_t)nh)) return 1;
    return c.finish(stats, 0);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ npairs_s_mu
extern "C" int htb_npairs_s_mu_engine(const htb_mesh_geom *mesh,
                                      const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                      const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                      const double *s_bins, int32_t ns, const double *mu_bins, int32_t nmu,
                                      int64_t first_cell1, int64_t last_cell1,
                                      int64_t *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !s_bins || !mu_bins || !counts_out || ns < 1 || nmu < 1) { htb_set_error("htb_npairs_s_mu_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_npairs_s_mu_engine needs a 3-d mesh"); return 1; }
    Call c;
    if (c.begin(flags)) return 1;
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    if (c.setup(mesh, 1, true, c1, stride1, n1, nullptr, c2, stride2, n2, nullptr, 0, false, first_cell1, last_cell1, flags)) return 1;
    if (c.prepared) return 0;
    std::vector<double> e((size_t)ns + nmu);
    double m0 = -INFINITY, m1 = -INFINITY;
    for (int k = 0; k < ns; ++k) { e[k] = s_bins[k] * s_bins[k]; if (e[k] > m0) m0 = e[k]; }
    for (int k = 0; k < nmu; ++k) { e[ns + k] = mu_bins[k] * mu_bins[k]; if (e[ns + k] > m1) m1 = e[ns + k]; }
    if (binq_ok(e.data(), ns, e.data() + ns, nmu, flags))
        return run_binq(c, 2, e.data(), ns, e.data() + ns, nmu, counts_out, stats);
    void *edev = nullptr;
    if (upload(c, e.data(), sizeof(double) * e.size(), &edev)) return 1;
    const int nh = ns * nmu;
    unsigned long long *counts_dev = nullptr;
    if (c.ws.alloc((void **)&counts_dev, sizeof(unsigned long long) * (size_t)nh)) return 1;
    HTB_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(unsigned long long) * (size_t)nh, c.st));
    GenParams gp{};
    gp.n0 = ns; gp.n1 = nmu; gp.nhist = nh;
    gp.e0 = (const double *)edev; gp.e1 = (const double *)edev + ns;
    gp.max0 = m0; gp.max1 = m1;
    gp.counts = counts_dev;
    if (htb_launch_gen(c.st, 2, c.G, c.A, gp, &c.launches)) return 1;
    if (c.mark_count_end()) return 1;
    if (nh > 4096) { htb_set_error("htb_npairs_s_mu_engine: at most 4096 (s, mu) cells"); return 1; }
    // 2-D inclusive prefix sums (npairs_s_mu_engine.pyx:232-234) over the tiny histogram
    unsigned long long *cum = nullptr;
    if (c.out_buffer(counts_out, (size_t)nh, (void **)&cum)) return 1;
    k_prefix2d<unsigned long long><<<1, 128, 0, c.st>>>(counts_dev, ns, nmu, cum);
    c.launches += 1;
    HTB_CUDA(cudaGetLastError());
    if (c.out_fetch(counts_out, cum, (size_t)nh)) return 1;
    return c.finish(stats, 0);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ marked_npairs_3d
extern "C" int htb_marked_npairs_3d_engine(const htb_mesh_geom *mesh,
                                           const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                           const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                           const double *w1, const double *w2, int32_t nw, int32_t weight_func_id,
                                           const double *rbins, int32_t nb, int64_t first_cell1, int64_t last_cell1,
                                           double *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rbins || !counts_out || nb < 1 || !w1 || !w2) { htb_set_error("htb_marked_npairs_3d_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_marked_npairs_3d_engine needs a 3-d mesh"); return 1; }
    if (nw < 1 || nw > HTB_MAX_NW) { htb_set_error("weights per point must be in [1, %d]", HTB_MAX_NW); return 1; }
    if (weight_func_id < 0 || weight_func_id > 17) { htb_set_error("marking function does not exist, id=%d", weight_func_id); return 1; }
    std::vector<double> rsq((size_t)nb);
    for (int k = 0; k < nb; ++k) rsq[k] = rbins[k] * rbins[k];
    // Fast path (MarkedQ): one weight per point and the product marking function (ids 0 and 1), edges as in
    // npairs_3d.  With the same sample AND the same weights on both sides the product is symmetric, so the
    // symmetric auto-correlation mode applies.
    double lmax = 0.0;
    Fast3Params fp{};
    const bool fast = nw == 1 && (weight_func_id == 0 || weight_func_id == 1) &&
                      fast3_params(mesh, rbins, rsq.data(), nb, flags, &fp, &lmax);
    Call c;
    if (c.begin(flags)) return 1;
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    if (c.setup(mesh, 1, fast && w1 == w2, c1, stride1, n1, w1, c2, stride2, n2, w2, nw, false, first_cell1, last_cell1, flags,
                HTB_TILE, fast ? 8.0 * lmax : 1.0e150)) return 1;
    if (c.prepared) return 0;
    if (fast) {
        double *fs = nullptr;
        if (c.ws.alloc((void **)&fs, sizeof(double) * (HTB_NBF + 1))) return 1;
        HTB_CUDA(cudaMemsetAsync(fs, 0, sizeof(double) * (HTB_NBF + 1), c.st));
        fp.fsums = fs;
        if (htb_launch_markedq(c.st, c.G, c.A, fp, &c.launches)) return 1;
        if (c.mark_count_end()) return 1;
        double *res = nullptr;
        if (c.out_buffer(counts_out, (size_t)nb, (void **)&res)) return 1;
        k_markedq_cumulate<<<1, 32, 0, c.st>>>(fs, nb, res);
        c.launches += 1;
        HTB_CUDA(cudaGetLastError());
        if (c.out_fetch(counts_out, res, (size_t)nb)) return 1;
        return c.finish(stats, 1);
    }
    {
        std::vector<double> em(rsq);
        suffix_min(em, 0, nb);
        if (binq_ok(em.data(), nb, nullptr, 1, flags))
            return run_binq_weighted(c, 0, nw, weight_func_id, em.data(), nb, nullptr, 1, counts_out, stats);
    }
    void *edev = nullptr;
    if (upload(c, rsq.data(), sizeof(double) * (size_t)nb, &edev)) return 1;
    double *counts_dev = nullptr;
    if (c.ws.alloc((void **)&counts_dev, sizeof(double) * (size_t)nb)) return 1;
    HTB_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(double) * (size_t)nb, c.st));
    GenParams gp{};
    gp.n0 = nb; gp.n1 = 1; gp.nhist = nb; gp.nw = nw; gp.wfunc = weight_func_id;
    gp.e0 = (const double *)edev;
    gp.fcounts = counts_dev;
    if (htb_launch_gen(c.st, 3, c.G, c.A, gp, &c.launches)) return 1;
    if (c.mark_count_end()) return 1;
    if (c.deliver(counts_out, counts_dev, (size_t)nb)) return 1;
    return c.finish(stats, 0);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ mean_delta_sigma
// ------------------------------------------------------------------ marked_npairs_xy_z
extern "C" int htb_marked_npairs_xy_z_engine(const htb_mesh_geom *mesh,
                                             const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                             const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                             const double *w1, const double *w2, int32_t nw, int32_t weight_func_id,
                                             const double *rp_bins, int32_t nrp, const double *pi_bins, int32_t npi,
                                             int64_t first_cell1, int64_t last_cell1,
                                             double *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rp_bins || !pi_bins || !counts_out || nrp < 1 || npi < 1 || !w1 || !w2) { htb_set_error("htb_marked_npairs_xy_z_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_marked_npairs_xy_z_engine needs a 3-d mesh"); return 1; }
    if (nw < 1 || nw > HTB_MAX_NW) { htb_set_error("weights per point must be in [1, %d]", HTB_MAX_NW); return 1; }
    if (weight_func_id < 0 || weight_func_id > 17) { htb_set_error("marking function does not exist, id=%d", weight_func_id); return 1; }
    std::vector<double> e((size_t)nrp + npi);
    for (int k = 0; k < nrp; ++k) e[k] = rp_bins[k] * rp_bins[k];
    for (int k = 0; k < npi; ++k) e[nrp + k] = pi_bins[k] * pi_bins[k];
    suffix_min(e, 0, nrp);
    suffix_min(e, nrp, npi);
    if (!binq_ok(e.data(), nrp, e.data() + nrp, npi, flags & ~HTB_FLAG_GENERIC)) {
        htb_set_error("htb_marked_npairs_xy_z_engine: bins must be finite, at most 255 per axis and 4096 cells");
        return 1;
    }
    Call c;
    if (c.begin(flags)) return 1;
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    if (c.setup(mesh, 0, false, c1, stride1, n1, w1, c2, stride2, n2, w2, nw, false, first_cell1, last_cell1, flags)) return 1;
    if (c.prepared) return 0;
    return run_binq_weighted(c, 1, nw, weight_func_id, e.data(), nrp, e.data() + nrp, npi, counts_out, stats);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ weighted_npairs_xy (2-D mesh)
extern "C" int htb_weighted_npairs_xy_engine(const htb_mesh_geom *mesh,
                                             const double *x1, const double *y1, int64_t stride1, int64_t n1,
                                             const double *x2, const double *y2, int64_t stride2, int64_t n2,
                                             const double *w2, const double *rp_bins, int32_t nrp,
                                             int64_t first_cell1, int64_t last_cell1,
                                             double *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rp_bins || !counts_out || nrp < 1 || !w2) { htb_set_error("htb_weighted_npairs_xy_engine: bad arguments"); return 1; }
    if (mesh->ndim != 2) { htb_set_error("htb_weighted_npairs_xy_engine needs a 2-d mesh"); return 1; }
    std::vector<double> e((size_t)nrp);
    for (int k = 0; k < nrp; ++k) e[k] = rp_bins[k] * rp_bins[k];
    suffix_min(e, 0, nrp);
    if (!binq_ok(e.data(), nrp, nullptr, 1, flags & ~HTB_FLAG_GENERIC)) {
        htb_set_error("htb_weighted_npairs_xy_engine: bins must be finite and at most 255");
        return 1;
    }
    Call c;
    if (c.begin(flags)) return 1;
    const double *c1[3] = {x1, y1, nullptr}, *c2[3] = {x2, y2, nullptr};
    if (c.setup(mesh, 1, false, c1, stride1, n1, nullptr, c2, stride2, n2, w2, 1, false, first_cell1, last_cell1, flags)) return 1;
    if (c.prepared) return 0;
    // few bins: lane-private rows + a warp reduction per tile (MODE 5) instead of shared-memory atomics on a handful of cells
    return run_binq_weighted(c, 3, 1, -1, e.data(), nrp, nullptr, 1, counts_out, stats, nrp <= 48 ? 5 : 1);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ weighted_npairs_per_object_xy (2-D mesh)
extern "C" int htb_weighted_npairs_per_object_xy_engine(const htb_mesh_geom *mesh,
                                                        const double *x1, const double *y1, int64_t stride1, int64_t n1,
                                                        const double *x2, const double *y2, int64_t stride2, int64_t n2,
                                                        const double *w2, const double *rp_bins, int32_t nrp,
                                                        int64_t first_cell1, int64_t last_cell1,
                                                        double *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rp_bins || !counts_out || nrp < 1 || !w2) { htb_set_error("htb_weighted_npairs_per_object_xy_engine: bad arguments"); return 1; }
    if (mesh->ndim != 2) { htb_set_error("htb_weighted_npairs_per_object_xy_engine needs a 2-d mesh"); return 1; }
    if (nrp > 48) { htb_set_error("htb_weighted_npairs_per_object_xy_engine: at most 48 rp_bins (per-point shared-memory rows)"); return 1; }
    std::vector<double> e((size_t)nrp);
    for (int k = 0; k < nrp; ++k) e[k] = rp_bins[k] * rp_bins[k];
    suffix_min(e, 0, nrp);
    if (!binq_ok(e.data(), nrp, nullptr, 1, flags & ~HTB_FLAG_GENERIC)) { htb_set_error("htb_weighted_npairs_per_object_xy_engine: rp_bins must be finite"); return 1; }
    Call c;
    if (flags & HTB_FLAG_DEVICE_OUTPUT) { htb_set_error("per-object tables are returned to host memory (HTB_FLAG_DEVICE_OUTPUT is not supported here)"); return 1; }
    if (c.begin()) return 1;
    const double *c1[3] = {x1, y1, nullptr}, *c2[3] = {x2, y2, nullptr};
    if (c.setup(mesh, 1, false, c1, stride1, n1, nullptr, c2, stride2, n2, w2, 1, true, first_cell1, last_cell1, flags)) return 1;
    if (c.prepared) return 0;
    BinQParams bp{};
    if (binq_prepare(c, e.data(), nrp, nullptr, 1, &bp)) return 1;
    const size_t nout = (size_t)(n1 > 0 ? n1 : 1) * (size_t)nrp;
    double *rows = nullptr;
    if (c.ws.alloc((void **)&rows, sizeof(double) * nout)) return 1;
    HTB_CUDA(cudaMemsetAsync(rows, 0, sizeof(double) * nout, c.st));
    bp.fcounts = rows;
    bp.perm1 = c.s1.perm;
    bp.nw = 1; bp.wfunc = -1;
    c.G.maxslices = 1;
    if (htb_launch_binq(c.st, 3, 4, c.G, c.A, bp, &c.launches)) return 1;
    if (c.mark_count_end()) return 1;
    if (download(c.st, rows, counts_out, n1 * (int64_t)nrp, flags)) return 1;
    return c.finish(stats, 3);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ npairs_per_object_3d
extern "C" int htb_npairs_per_object_3d_engine(const htb_mesh_geom *mesh,
                                               const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                               const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                               const double *rbins, int32_t nb, int64_t first_cell1, int64_t last_cell1,
                                               int64_t *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rbins || !counts_out || nb < 1) { htb_set_error("htb_npairs_per_object_3d_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_npairs_per_object_3d_engine needs a 3-d mesh"); return 1; }
    if (nb > 64) { htb_set_error("htb_npairs_per_object_3d_engine: at most 64 rbins (per-point shared-memory rows)"); return 1; }
    std::vector<double> e((size_t)nb);
    for (int k = 0; k < nb; ++k) e[k] = rbins[k] * rbins[k];
    suffix_min(e, 0, nb);
    if (!binq_ok(e.data(), nb, nullptr, 1, flags & ~HTB_FLAG_GENERIC)) { htb_set_error("htb_npairs_per_object_3d_engine: rbins must be finite"); return 1; }
    Call c;
    if (flags & HTB_FLAG_DEVICE_OUTPUT) { htb_set_error("per-object tables are returned to host memory (HTB_FLAG_DEVICE_OUTPUT is not supported here)"); return 1; }
    if (c.begin()) return 1;
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    if (c.setup(mesh, 1, false, c1, stride1, n1, nullptr, c2, stride2, n2, nullptr, 0, true, first_cell1, last_cell1, flags)) return 1;
    if (c.prepared) return 0;
    BinQParams bp{};
    if (binq_prepare(c, e.data(), nb, nullptr, 1, &bp)) return 1;
    const size_t nout = (size_t)(n1 > 0 ? n1 : 1) * (size_t)nb;
    unsigned long long *rows = nullptr;
    if (c.ws.alloc((void **)&rows, sizeof(unsigned long long) * nout)) return 1;
    HTB_CUDA(cudaMemsetAsync(rows, 0, sizeof(unsigned long long) * nout, c.st));
    bp.rows = rows;
    bp.perm1 = c.s1.perm;
    c.G.maxslices = 1;             // every row is written by one work item (no column slices: they would each add 64 x nb atomics)
    if (htb_launch_binq(c.st, 0, 2, c.G, c.A, bp, &c.launches)) return 1;
    if (c.mark_count_end()) return 1;
    if (download(c.st, (const double *)rows, (double *)counts_out, n1 * (int64_t)nb, flags)) return 1;      // 8-byte words
    return c.finish(stats, 3);
    HTB_GUARD_END
}

// ------------------------------------------------------------------ npairs_jackknife_3d / npairs_jackknife_xy_z
// The reference adds jweight(s, j1, j2, w1, w2) to counts[s] for EVERY sample s and every pair
// (npairs_jackknife_3d_engine.pyx:227-233): w1*w2 if neither point carries tag s, half of it if exactly one does, 0 if
// both do, and w1*w2 for s = 0 (:283-289).  With A[s] = the weighted pair sums of the sample1 points tagged s and B[s]
// the same for the sample2 points, counts[0] = T = sum_s A[s] and counts[s] = T - (A[s] + B[s]) / 2 - O(1) work per
// pair instead of O(N_samples).  A comes from one pass of the per-object BinQ kernel (rows folded by tag); B from a
// second pass with the roles of the samples exchanged, on mesh1's grid for both samples (cells >= the search length,
// cover 1: every pair within the search length is visited once, with the image the reference uses).  In pass B the
// periodic shift is applied to the STAGED sample1 coordinate (BinQ MODE 6): x1 + shift_B = x1 - shift_A exactly, so both
// passes evaluate the reference's own (x1 - shift) - x2 and every pair lands in the same bin in A and in B.
// host twin of k_prefix2d (the jackknife tables are combined on the host)
template <class T>
static void prefix2d(const T *diff, int n0, int n1, T *out)
{
    for (int k = 0; k < n0; ++k) { T run = 0; for (int g = 0; g < n1; ++g) { run += diff[(size_t)k * n1 + g]; out[(size_t)k * n1 + g] = run; } }
    for (int g = 0; g < n1; ++g) { T run = 0; for (int k = 0; k < n0; ++k) { run += out[(size_t)k * n1 + g]; out[(size_t)k * n1 + g] = run; } }
}

static int jackknife_pass(const htb_mesh_geom *mesh, int kind, bool swapped,
                          const double *const *ca, int64_t sa, int64_t na, const double *pa,
                          const double *const *cb, int64_t sb, int64_t nb_, const double *pb,
                          const double *e0, int n0, const double *e1, int n1, int32_t nsamples,
                          int64_t first, int64_t last, std::vector<double> &table, uint32_t flags, htb_stats *stats)
{
    Call c;
    if (c.begin()) return 1;
    if (c.setup(mesh, kind == 0 ? 1 : 0, false, ca, sa, na, pa, cb, sb, nb_, pb, 2, false, first, last, flags)) return 1;
    if (c.prepared) return 0;
    BinQParams bp{};
    if (binq_prepare(c, e0, n0, e1, n1, &bp)) return 1;
    const size_t nh = (size_t)n0 * n1, nt = (size_t)(nsamples + 1) * nh;
    double *tab = nullptr;
    if (c.ws.alloc((void **)&tab, sizeof(double) * nt)) return 1;
    HTB_CUDA(cudaMemsetAsync(tab, 0, sizeof(double) * nt, c.st));
    bp.fcounts = tab;
    bp.nw = 2; bp.wfunc = 1;
    if (nh > HTB_JK_SHARED_CELLS) {
        // wide rows (rp_pi_tpcf_jackknife): the warps' 64 point rows live in global memory (L2 resident per warp)
        int dev = 0, sms = 0;
        HTB_CUDA(cudaGetDevice(&dev));
        HTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        bp.grows_warps = (unsigned)sms * 8u * 4u;              // <= 8 resident blocks of 4 warps per SM
        const size_t gbytes = sizeof(double) * (size_t)bp.grows_warps * 64 * (nh | 1);
        if (c.ws.alloc((void **)&bp.grows, gbytes)) return 1;
        HTB_CUDA(cudaMemsetAsync(bp.grows, 0, gbytes, c.st));
    }
    if (htb_launch_binq(c.st, kind, swapped ? 6 : 3, c.G, c.A, bp, &c.launches)) return 1;
    table.assign(nt, 0.0);
    HTB_CUDA(cudaMemcpyAsync(table.data(), tab, sizeof(double) * nt, cudaMemcpyDeviceToHost, c.st));
    return c.finish(stats, 3);
}

static int run_jackknife(const htb_mesh_geom *mesh, int kind,
                         const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                         const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                         const double *w1, const double *w2, const int64_t *jtags1, const int64_t *jtags2, int32_t nsamples,
                         const double *e0in, int n0, const double *e1in, int n1e,
                         int64_t first_cell1, int64_t last_cell1, double *counts_out, uint32_t flags, htb_stats *stats)
{
    if (flags & (HTB_FLAG_DEVICE_INPUT | HTB_FLAG_DEVICE_OUTPUT)) { htb_set_error("jackknife engines take and return host arrays"); return 1; }
    if (nsamples < 1 || nsamples > 100000) { htb_set_error("N_samples must be in [1, 100000]"); return 1; }
    std::vector<double> e((size_t)n0 + n1e);
    for (int k = 0; k < n0; ++k) e[k] = e0in[k] * e0in[k];
    for (int k = 0; k < n1e; ++k) e[(size_t)n0 + k] = e1in[k] * e1in[k];
    suffix_min(e, 0, n0);
    if (n1e > 1 || kind == 1) suffix_min(e, n0, n1e);
    const double *e1 = kind == 1 ? e.data() + n0 : nullptr;
    const int nn1 = kind == 1 ? n1e : 1;
    if (!binq_ok(e.data(), n0, e1, nn1, flags & ~HTB_FLAG_GENERIC)) {
        htb_set_error("jackknife engines: bins must be finite, at most 127 per axis and 4096 cells");
        return 1;
    }
    // payload rows {weight, tag}
    std::vector<double> p1((size_t)(n1 > 0 ? n1 : 1) * 2), p2((size_t)(n2 > 0 ? n2 : 1) * 2);
    for (int64_t i = 0; i < n1; ++i) {
        if (jtags1[i] < 0 || jtags1[i] > nsamples) { htb_set_error("jtags1 out of [0, N_samples]"); return 1; }
        p1[(size_t)i * 2] = w1[i]; p1[(size_t)i * 2 + 1] = (double)jtags1[i];
    }
    for (int64_t i = 0; i < n2; ++i) {
        if (jtags2[i] < 0 || jtags2[i] > nsamples) { htb_set_error("jtags2 out of [0, N_samples]"); return 1; }
        p2[(size_t)i * 2] = w2[i]; p2[(size_t)i * 2 + 1] = (double)jtags2[i];
    }
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    std::vector<double> A, B;
    // pass A: the reference's own geometry and cell range
    if (jackknife_pass(mesh, kind, false, c1, stride1, n1, p1.data(), c2, stride2, n2, p2.data(),
                       e.data(), n0, e1, nn1, nsamples, first_cell1, last_cell1, A, flags, stats)) return 1;
    // pass B: roles exchanged, both samples on mesh1's grid.  A caller-chosen cell range selects sample1 points
    // (the reference's worker split): the sample1 points outside it get weight 0 here.
    htb_mesh_geom gb = *mesh;
    int64_t nc1 = 1;
    for (int d = 0; d < 3; ++d) {
        gb.ndivs2[d] = gb.ndivs1[d];
        gb.cell2_size[d] = gb.cell1_size[d];
        gb.cover[d] = (int)ceil(gb.search[d] / gb.cell1_size[d]);
        if (gb.cover[d] < 1) gb.cover[d] = 1;
        nc1 *= mesh->ndivs1[d];
    }
    if (first_cell1 > 0 || last_cell1 < nc1) {
        std::vector<int64_t> ids((size_t)(n1 > 0 ? n1 : 1));
        if (n1 > 0 && htb_mesh_cell_ids(3, x1, y1, z1, stride1, n1, mesh->cell1_size, mesh->ndivs1, ids.data(), 0)) return 1;
        for (int64_t i = 0; i < n1; ++i) if (ids[(size_t)i] < first_cell1 || ids[(size_t)i] >= last_cell1) p1[(size_t)i * 2] = 0.0;
    }
    htb_stats sb;
    if (jackknife_pass(&gb, kind, true, c2, stride2, n2, p2.data(), c1, stride1, n1, p1.data(),
                       e.data(), n0, e1, nn1, nsamples, 0, nc1, B, flags, stats ? &sb : nullptr)) return 1;
    if (stats) {
        stats->pairs_evaluated += sb.pairs_evaluated;
        stats->ms_h2d += sb.ms_h2d; stats->ms_mesh += sb.ms_mesh; stats->ms_count += sb.ms_count; stats->ms_total += sb.ms_total;
        stats->kernel_launches += sb.kernel_launches;
    }
    // differential -> cumulative per sample, then counts[s] = T - (A[s] + B[s]) / 2
    const size_t nh = (size_t)n0 * nn1;
    std::vector<double> T(nh, 0.0), cum(nh), row(nh);
    for (int s = 0; s <= nsamples; ++s) for (size_t k = 0; k < nh; ++k) T[k] += A[(size_t)s * nh + k];
    prefix2d<double>(T.data(), n0, nn1, cum.data());
    for (size_t k = 0; k < nh; ++k) counts_out[k] = cum[k];
    for (int s = 1; s <= nsamples; ++s) {
        for (size_t k = 0; k < nh; ++k) row[k] = A[(size_t)s * nh + k] + B[(size_t)s * nh + k];
        std::vector<double> rc(nh);
        prefix2d<double>(row.data(), n0, nn1, rc.data());
        for (size_t k = 0; k < nh; ++k) counts_out[(size_t)s * nh + k] = cum[k] - 0.5 * rc[k];
    }
    return 0;
}

extern "C" int htb_npairs_jackknife_3d_engine(const htb_mesh_geom *mesh,
                                              const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                              const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                              const double *w1, const double *w2, const int64_t *jtags1, const int64_t *jtags2,
                                              int32_t n_samples, const double *rbins, int32_t nb,
                                              int64_t first_cell1, int64_t last_cell1,
                                              double *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rbins || !counts_out || nb < 1 || !w1 || !w2 || !jtags1 || !jtags2) { htb_set_error("htb_npairs_jackknife_3d_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_npairs_jackknife_3d_engine needs a 3-d mesh"); return 1; }
    return run_jackknife(mesh, 0, x1, y1, z1, stride1, n1, x2, y2, z2, stride2, n2, w1, w2, jtags1, jtags2, n_samples,
                         rbins, nb, nullptr, 0, first_cell1, last_cell1, counts_out, flags, stats);
    HTB_GUARD_END
}

extern "C" int htb_npairs_jackknife_xy_z_engine(const htb_mesh_geom *mesh,
                                                const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                                const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                                const double *w1, const double *w2, const int64_t *jtags1, const int64_t *jtags2,
                                                int32_t n_samples, const double *rp_bins, int32_t nrp, const double *pi_bins, int32_t npi,
                                                int64_t first_cell1, int64_t last_cell1,
                                                double *counts_out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rp_bins || !pi_bins || !counts_out || nrp < 1 || npi < 1 || !w1 || !w2 || !jtags1 || !jtags2) { htb_set_error("htb_npairs_jackknife_xy_z_engine: bad arguments"); return 1; }
    if (mesh->ndim != 3) { htb_set_error("htb_npairs_jackknife_xy_z_engine needs a 3-d mesh"); return 1; }
    return run_jackknife(mesh, 1, x1, y1, z1, stride1, n1, x2, y2, z2, stride2, n2, w1, w2, jtags1, jtags2, n_samples,
                         rp_bins, nrp, pi_bins, npi, first_cell1, last_cell1, counts_out, flags, stats);
    HTB_GUARD_END
}

// Column sums of the (n, nbin) per-object rows in a fixed order (HTB_FLAG_COLUMN_SUM): block b sums rows
// b, b + gridDim.x, ... per column, a second launch adds the per-block partial sums.
__global__ void __launch_bounds__(256) k_colsum_partial(const double *__restrict__ rows, long long n, int nbin,
                                                        double *__restrict__ partial)
{
    __shared__ double sh[256];
    const long long total = n * nbin;
    // thread t of block b owns column (t % nbin) for rows b * rpb + t / nbin + i * rows_per_pass
    const int rows_per_block = 256 / nbin;             // nbin <= 256 checked on the host
    const int r = threadIdx.x / nbin, k = threadIdx.x % nbin;
    double s = 0.0;
    if (r < rows_per_block)
        for (long long row = (long long)blockIdx.x * rows_per_block + r; row < n; row += (long long)gridDim.x * rows_per_block)
            s += rows[row * nbin + k];
    (void)total;
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < nbin) {
        double t = 0.0;
        for (int q = 0; q < rows_per_block; ++q) t += sh[q * nbin + threadIdx.x];
        partial[(size_t)blockIdx.x * nbin + threadIdx.x] = t;
    }
}
__global__ void k_colsum_final(const double *__restrict__ partial, int nblocks, int nbin, double *__restrict__ out)
{
    const int k = threadIdx.x;
    if (k >= nbin) return;
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += partial[(size_t)b * nbin + k];
    out[k] = t;
}

extern "C" int htb_mean_delta_sigma_engine(const htb_mesh_geom *mesh,
                                           const double *x1, const double *y1, int64_t stride1, int64_t n1,
                                           const double *x2, const double *y2, int64_t stride2, const double *m2, int64_t n2,
                                           const double *rp_bins, int32_t nrp, int64_t first_cell1, int64_t last_cell1,
                                           double *out, uint32_t flags, htb_stats *stats)
{
    HTB_GUARD_BEGIN
    if (!mesh || !rp_bins || !out || nrp < 2 || !m2) { htb_set_error("htb_mean_delta_sigma_engine: bad arguments"); return 1; }
    if (mesh->ndim != 2) { htb_set_error("htb_mean_delta_sigma_engine needs a 2-d mesh"); return 1; }
    Call c;
    if (c.begin(flags)) return 1;
    const double *c1[3] = {x1, y1, nullptr}, *c2[3] = {x2, y2, nullptr};
    // HTB_FLAG_UNIFORM_MASS: every particle has the mass m2[0] (the reference's scalar
    // ``effective_particle_masses``): the mass array is neither uploaded nor sorted.
    const bool uniform = (flags & HTB_FLAG_UNIFORM_MASS) != 0;
    double mass = 0.0;
    if (uniform && n2 > 0) {
        if (flags & HTB_FLAG_DEVICE_INPUT) HTB_CUDA(cudaMemcpy(&mass, m2, sizeof(double), cudaMemcpyDeviceToHost));
        else mass = m2[0];
    }
    // fast path (DSigmaQ): uniform mass, <= HTB_NBF monotone edges whose 32-bit relative keys cannot wrap
    // (same window argument as htb_npairs_3d_engine)
    std::vector<double> rsq((size_t)nrp);
    for (int k = 0; k < nrp; ++k) rsq[k] = rp_bins[k] * rp_bins[k];
    bool fast = uniform && !(flags & HTB_FLAG_GENERIC) && nrp <= HTB_NBF && finite_all(rp_bins, nrp) && std::isfinite(mass);
    for (int k = 0; k + 1 < nrp && fast; ++k) fast = rp_bins[k] >= 0.0 && rsq[k] <= rsq[k + 1];
    if (fast) fast = rsq[nrp - 1] > 1e-290 && rsq[nrp - 1] < 1e290;
    double lmax = mesh->period[0] > mesh->period[1] ? mesh->period[0] : mesh->period[1];
    if (fast) fast = std::isfinite(lmax) && lmax > 0.0 && lmax < 1e140;
    DSQParams qp{};
    if (fast) {
        const long long Kt = (long long)(dbits(rsq[nrp - 1]) >> 26);
        const long long Kmax = (long long)(dbits(256.0 * lmax * lmax) >> 26);
        const long long lim = 31LL << 26;
        if (Kmax - Kt >= lim) fast = false;
        qp.nrp = nrp;
        qp.nbias = (int)(unsigned)(0ULL - (unsigned long long)Kt);
        qp.Hwin = (int)(dbits(rsq[nrp - 1]) >> 32) - (31 << 20);
        const int pad = HTB_NBF - nrp;
        for (int s = 0; s < HTB_NBF; ++s) qp.E[s] = s < pad ? 0ULL : dbits(rsq[s - pad]);
        // lower edge of the top annulus; an edge below the key window (or zero) makes every in-range pair
        // take the queue (Tspan = 0 keeps the in-register top annulus off only when F1 == 0)
        long long d = (long long)(dbits(rsq[nrp - 2]) >> 26) - Kt;
        if (rsq[nrp - 2] == 0.0 || d <= -lim) d = -lim + 1;
        qp.F1 = (int)d;
        qp.Tspan = d < 0 ? (unsigned)(-d - 1) : 0u;
        qp.mass = mass;
    }
    // cell-resolved fast path (DSigmaR): additionally needs a positive lowest edge and squared separations / edges
    // within the exponent window its running products are renormalised for
    bool ring = fast && nrp >= 2 && rsq[0] >= 1e-290 && !getenv("HTB_NO_DSR");
    if (ring) {
        const double et = rsq[nrp - 1];
        ring = et >= ldexp(1.0, -38) && et <= ldexp(1.0, 58) && 8.0 * lmax * lmax <= ldexp(et, 40);
        for (int d = 0; d < 2 && ring; ++d) ring = (int64_t)mesh->ndivs2[d] * 16 < 65536;
    }
    if (c.setup(mesh, 1, false, c1, stride1, n1, nullptr, c2, stride2, n2, uniform ? nullptr : m2, uniform ? 0 : 1, true,
                first_cell1, last_cell1, flags, fast ? 32 : HTB_TILE, fast ? 8.0 * lmax : 1.0e150, ring ? 1 : 0)) return 1;
    if (c.prepared) return 0;
    const int nbin = nrp - 1;
    std::vector<double> e((size_t)nrp + nbin);
    for (int k = 0; k < nrp; ++k) e[k] = rp_bins[k] * rp_bins[k];
    for (int k = 0; k < nbin; ++k) e[nrp + k] = log(rp_bins[k + 1] / rp_bins[k]);
    void *edev = nullptr;
    if (upload(c, e.data(), sizeof(double) * e.size(), &edev)) return 1;
    double *out_dev = nullptr;
    const size_t nout = (size_t)(n1 > 0 ? n1 : 1) * nbin;
    if (c.ws.alloc((void **)&out_dev, sizeof(double) * nout)) return 1;
    HTB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * nout, c.st));
    if (ring) {
        DSRParams rp{};
        rp.nrp = nrp;
        for (int k = 0; k < HTB_NBF; ++k) rp.Ed[k] = k < nrp ? rsq[k] : INFINITY;
        rp.tiny2 = ldexp(rsq[nrp - 1], -24);
        {
            // factors lie in [tiny2, 8 L^2]: K of them move a product that started in [1, 2) by at most K * maxlog binades;
            // keep that below 900 of the 1022 available (the banked mantissa adds one more)
            const double maxlog = std::max(fabs(log2(8.0 * lmax * lmax)), fabs(log2(rp.tiny2))) + 1.0;
            int K = (int)(900.0 / maxlog);
            K = K < 8 ? 8 : (K > 64 ? 64 : K);
            rp.renorm = K & ~7;
            if (const char *e = getenv("HTB_DSR_RENORM")) { const int v = atoi(e); if (v >= 8 && v <= K) rp.renorm = v & ~7; }
        }
        rp.mass = mass;
        rp.e0 = (const double *)edev; rp.e1 = (const double *)edev + nrp;
        rp.out = out_dev;
        rp.perm1 = c.s1.perm;
        if (htb_launch_dsr(c.st, c.G, c.A, rp, &c.launches)) return 1;
    } else if (fast) {
        qp.e0 = (const double *)edev; qp.e1 = (const double *)edev + nrp;
        qp.out = out_dev;
        qp.perm1 = c.s1.perm;
        if (htb_launch_dsq(c.st, c.G, c.A, qp, &c.launches)) return 1;
    } else {
        GenParams gp{};
        gp.n0 = nrp; gp.n1 = nbin; gp.nhist = 0; gp.nw = 1;
        gp.e0 = (const double *)edev; gp.e1 = (const double *)edev + nrp;
        gp.fcounts = out_dev;
        gp.perm1 = c.s1.perm;
        gp.max0 = mass;
        if (htb_launch_gen(c.st, uniform ? 5 : 4, c.G, c.A, gp, &c.launches)) return 1;
    }
    if (flags & HTB_FLAG_COLUMN_SUM) {
        if (nbin > 256) { htb_set_error("HTB_FLAG_COLUMN_SUM supports at most 256 bins"); return 1; }
        const int nblocks = 592;
        double *partial = nullptr, *sums = nullptr;
        if (c.ws.alloc((void **)&partial, sizeof(double) * (size_t)nblocks * nbin)) return 1;
        if (c.ws.alloc((void **)&sums, sizeof(double) * (size_t)nbin)) return 1;
        k_colsum_partial<<<nblocks, 256, 0, c.st>>>(out_dev, (long long)(n1 > 0 ? n1 : 0), nbin, partial);
        k_colsum_final<<<1, 256, 0, c.st>>>(partial, nblocks, nbin, sums);
        c.launches += 2;
        HTB_CUDA(cudaGetLastError());
        if (c.mark_count_end()) return 1;
        if (c.deliver(out, sums, (size_t)nbin)) return 1;
    } else if (n1 > 0) {
        if (c.mark_count_end()) return 1;
        if (c.async) { if (c.deliver(out, out_dev, (size_t)n1 * nbin)) return 1; }
        else if (download(c.st, out_dev, out, n1 * (int64_t)nbin, flags)) return 1;
    }
    return c.finish(stats, ring ? 2 : (fast ? 1 : 0));
    HTB_GUARD_END
}

// ------------------------------------------------------------------ K3: two-point estimators on device-resident counts
// tpcf_estimators.py:14-119 (_TP_estimator, _TP_estimator_crossx), the np.diff of tpcf.py:76-113 / rp_pi_tpcf.py:296-330
// and wp's 2 * xi * pi_max (wp.py:219-221), evaluated by one block over the (tiny) cumulative count tables the
// engines left on the device - after the ranks' tables were summed by an all-reduce on the same stream.  Every
// expression is written as numpy evaluates the reference's (one rounding per operation, -fmad=false), so the result
// is the host formula's bit for bit.
struct TpOperand {
    const long long *cum;      // cumulative int64 counts [n0 * n1], or null
    const double *diff;        // differential float counts (analytic randoms) [nout], used when cum is null
};
__device__ __forceinline__ long long tp_diff(const long long *c, int k, int g, int n1)
{
    if (n1 == 1) return c[k + 1] - c[k];
    return (c[(k + 1) * n1 + g + 1] - c[k * n1 + g + 1]) - (c[(k + 1) * n1 + g] - c[k * n1 + g]);
}
__global__ void k_tp_estimator(int n0, int n1, int estimator, int cross, TpOperand DD, TpOperand D1R, TpOperand D2R, TpOperand RR,
                               double inv_f1, double inv_f2, double wp_pi_max, double *__restrict__ xi, int *__restrict__ flag)
{
    const int ng = n1 > 1 ? n1 - 1 : 1;
    const int nout = (n0 - 1) * ng;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        const int k = i / ng, g = i % ng;
        // operands as numpy sees them: int64 arrays become float64 in a true division; int64 * int64 stays int64 (wraps)
        long long idd = 0, id1 = 0, id2 = 0, irr = 0;
        double dd = 0.0, d1 = 0.0, d2 = 0.0, rr = 0.0;
        const bool hdd = DD.cum || DD.diff, hd1 = D1R.cum || D1R.diff, hd2 = D2R.cum || D2R.diff, hrr = RR.cum || RR.diff;
        if (DD.cum) { idd = tp_diff(DD.cum, k, g, n1); dd = (double)idd; } else if (DD.diff) dd = DD.diff[i];
        if (D1R.cum) { id1 = tp_diff(D1R.cum, k, g, n1); d1 = (double)id1; } else if (D1R.diff) d1 = D1R.diff[i];
        if (D2R.cum) { id2 = tp_diff(D2R.cum, k, g, n1); d2 = (double)id2; } else if (D2R.diff) d2 = D2R.diff[i];
        if (RR.cum) { irr = tp_diff(RR.cum, k, g, n1); rr = (double)irr; } else if (RR.diff) rr = RR.diff[i];
        (void)hdd;
        // _test_for_zero_division (tpcf_estimators.py:165-183)
        if (estimator != 3 && hrr && rr == 0.0) atomicOr(flag, 1);
        if (estimator == 3 && ((hd1 && d1 == 0.0) || (cross && hd2 && d2 == 0.0))) atomicOr(flag, 2);
        double v;
        if (estimator == 0) v = inv_f1 * (dd / rr) - 1.0;                                   // Natural
        else if (estimator == 1) v = inv_f1 * (dd / d1) - 1.0;                              // Davis-Peebles
        else if (estimator == 2) v = inv_f1 * (dd / rr) - inv_f2 * (d1 / rr);               // Hewett
        else if (estimator == 3) {                                                          // Hamilton
            // (DD * RR) / (DR * DR): products of two int64 arrays are int64 products in numpy
            const double num = (DD.cum && RR.cum) ? (double)(long long)((unsigned long long)idd * (unsigned long long)irr) : dd * rr;
            double den;
            if (cross) den = (D1R.cum && D2R.cum) ? (double)(long long)((unsigned long long)id1 * (unsigned long long)id2) : d1 * d2;
            else den = D1R.cum ? (double)(long long)((unsigned long long)id1 * (unsigned long long)id1) : d1 * d1;
            v = num / den - 1.0;
        } else if (!cross) v = inv_f1 * (dd / rr) - inv_f2 * (2.0 * d1 / rr) + 1.0;        // Landy-Szalay
        else v = inv_f1 * (dd / rr) - inv_f2 * (d1 / rr) - inv_f2 * (d2 / rr) + 1.0;
        if (wp_pi_max > 0.0) v = 2.0 * v * wp_pi_max;
        xi[i] = v;
    }
}

extern "C" int htb_tp_estimator(int32_t n0, int32_t n1, int32_t estimator, int32_t cross,
                                const int64_t *DD_cum, const int64_t *D1R_cum, const int64_t *D2R_cum, const int64_t *RR_cum,
                                const double *D1R_diff, const double *D2R_diff, const double *RR_diff,
                                double inv_factor1, double inv_factor2, double wp_pi_max,
                                double *xi_out, int32_t *flag_out)
{
    HTB_GUARD_BEGIN
    if (n0 < 2 || n1 < 1 || (long long)n0 * n1 > 65536 || estimator < 0 || estimator > 4 || !xi_out || !flag_out) {
        htb_set_error("htb_tp_estimator: bad arguments");
        return 1;
    }
    if (wp_pi_max > 0.0 && n1 != 2) { htb_set_error("htb_tp_estimator: the wp integration needs exactly two pi edges"); return 1; }
    if (cross && (estimator == 1 || estimator == 2)) { htb_set_error("this estimator is not supported for cross-correlations"); return 1; }
    cudaStream_t st;
    if (get_stream(&st)) return 1;
    TpOperand a{(const long long *)DD_cum, nullptr}, b{(const long long *)D1R_cum, D1R_diff}, c{(const long long *)D2R_cum, D2R_diff},
              d{(const long long *)RR_cum, RR_diff};
    k_tp_estimator<<<1, 256, 0, st>>>(n0, n1, estimator, cross, a, b, c, d, inv_factor1, inv_factor2, wp_pi_max, xi_out, flag_out);
    HTB_CUDA(cudaGetLastError());
    return 0;
    HTB_GUARD_END
}

// Host -> device copy of `count` doubles on the thread's stream (asynchronous): pinned memory goes out with one
// cudaMemcpyAsync, large pageable arrays through the ring of pinned chunks filled by host threads (staged_upload) - the
// same path the engines use for their inputs, exposed so that a statistic can bring every sample to the device ONCE
// (or a rank its 1/world share, completed by an all-gather over NVLink) before it chains several engine calls.
extern "C" int htb_upload_f64(const double *host_src, int64_t count, double *dev_dst)
{
    HTB_GUARD_BEGIN
    if (count < 0 || (count > 0 && (!host_src || !dev_dst))) { htb_set_error("htb_upload_f64: bad arguments"); return 1; }
    if (count == 0) return 0;
    cudaStream_t st;
    if (get_stream(&st)) return 1;
    if (count >= 4 * (int64_t)HTB_STAGE_CHUNK && host_pointer_is_pageable(host_src) && !getenv("HTB_NO_STAGED_UPLOAD")) {
        const double *src[1] = {host_src};
        double *dst[1] = {dev_dst};
        return staged_upload(st, src, 1, 1, count, dst);
    }
    HTB_CUDA(cudaMemcpyAsync(dev_dst, host_src, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, st));
    return 0;
    HTB_GUARD_END
}

// the stream this thread's engine calls are issued on (cudaStream_t as void *), and a host wait for it: the one
// synchronisation of a statistic built from asynchronous (HTB_FLAG_DEVICE_OUTPUT) engine calls
extern "C" int htb_get_stream(void **stream_out)
{
    HTB_GUARD_BEGIN
    cudaStream_t st;
    if (!stream_out) { htb_set_error("htb_get_stream: null argument"); return 1; }
    if (get_stream(&st)) return 1;
    *stream_out = (void *)st;
    return 0;
    HTB_GUARD_END
}
// elapsed times (ms) of the counting kernels of this thread's asynchronous calls since the previous query, oldest
// first (at most HTB_ASYNC_RING are kept); the caller must have synchronised the stream
extern "C" int htb_async_count_times(float *ms_out, int32_t max_out, int32_t *n_out)
{
    HTB_GUARD_BEGIN
    if (!ms_out || !n_out || max_out < 0) { htb_set_error("htb_async_count_times: bad arguments"); return 1; }
    const int have = g_async_n < HTB_ASYNC_RING ? g_async_n : HTB_ASYNC_RING;
    const int n = have < max_out ? have : max_out;
    for (int k = 0; k < n; ++k) {
        cudaEvent_t *pair = g_async_ev[(g_async_n - n + k) % HTB_ASYNC_RING];
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pair[0], pair[1]) != cudaSuccess) { cudaGetLastError(); ms = -1.f; }
        ms_out[k] = ms;
    }
    *n_out = n;
    g_async_n = 0;
    return 0;
    HTB_GUARD_END
}
// the same launches by their own device-side stamps (first warp in -> last warp out, ms): what the kernel took once its
// blocks ran, without the time it waited behind the kernels of other streams.  Does not reset the ring: call it BEFORE
// htb_async_count_times().  -1 where a call launched no count kernel.
extern "C" int htb_async_kernel_spans(float *ms_out, int32_t max_out, int32_t *n_out)
{
    HTB_GUARD_BEGIN
    if (!ms_out || !n_out || max_out < 0) { htb_set_error("htb_async_kernel_spans: bad arguments"); return 1; }
    const int have = g_async_n < HTB_ASYNC_RING ? g_async_n : HTB_ASYNC_RING;
    const int n = have < max_out ? have : max_out;
    for (int k = 0; k < n; ++k) {
        const unsigned long long *slot = g_async_ts ? g_async_ts[(g_async_n - n + k) % HTB_ASYNC_RING] : nullptr;
        ms_out[k] = (slot && slot[0] && slot[1] && slot[1] > ~slot[0]) ? (float)((double)(slot[1] - ~slot[0]) * 1e-6) : -1.f;
    }
    *n_out = n;
    return 0;
    HTB_GUARD_END
}
// ... and the raw stamps {first warp in, last warp out} (globaltimer, ns; 0, 0 where no kernel ran), two words per call:
// launches of one statistic that run side by side on several streams are measured by the union of their spans
extern "C" int htb_async_kernel_stamps(uint64_t *ns_out, int32_t max_out, int32_t *n_out)
{
    HTB_GUARD_BEGIN
    if (!ns_out || !n_out || max_out < 0) { htb_set_error("htb_async_kernel_stamps: bad arguments"); return 1; }
    const int have = g_async_n < HTB_ASYNC_RING ? g_async_n : HTB_ASYNC_RING;
    const int n = have < max_out ? have : max_out;
    for (int k = 0; k < n; ++k) {
        const unsigned long long *slot = g_async_ts ? g_async_ts[(g_async_n - n + k) % HTB_ASYNC_RING] : nullptr;
        const bool ok = slot && slot[0] && slot[1];
        ns_out[2 * k] = ok ? (uint64_t)~slot[0] : 0;
        ns_out[2 * k + 1] = ok ? (uint64_t)slot[1] : 0;
    }
    *n_out = n;
    return 0;
    HTB_GUARD_END
}
extern "C" int htb_stream_synchronize(void)
{
    HTB_GUARD_BEGIN
    cudaStream_t st;
    if (get_stream(&st)) return 1;
    HTB_CUDA(cudaStreamSynchronize(st));
    return 0;
    HTB_GUARD_END
}

// ------------------------------------------------------------------ 8f-4: the input step on the device
// catalog_analysis_helpers.py:108-265 (return_xyz_formatted_array) and :268-327 (apply_zspace_distortion) for samples that
// already live in HBM (a mock generated or kept on the device between likelihood evaluations): one elementwise pass that
// writes the (N, 3) row-major sample the pair counters take, so no coordinate crosses PCIe.  Every expression is
// evaluated operation by operation as numpy does (np.mod = fmod with the sign fix of npy_divmod; the distortion
// ((1 + z) * v) / 100 / E(z)), so the device sample is bit-identical to the host function's.
__device__ __forceinline__ double htb_np_mod(double a, double b)
{
    // numpy's float remainder (npy_divmod): the result takes the sign of the divisor
    double mod = fmod(a, b);
    if (b == 0.0) return mod;                        // nan, as numpy
    if (mod != 0.0) { if ((b < 0.0) != (mod < 0.0)) mod += b; }
    else mod = copysign(0.0, b);
    return mod;
}
__global__ void __launch_bounds__(256) k_xyz_formatted(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                                                       long long n, double px, double py, double pz,
                                                       const double *__restrict__ vel, int dist_dim, double one_plus_z, double efunc,
                                                       double *__restrict__ out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double p[3] = {htb_np_mod(x[i], px), htb_np_mod(y[i], py), htb_np_mod(z[i], pz)};
        if (dist_dim >= 0) {
            const double L = dist_dim == 0 ? px : (dist_dim == 1 ? py : pz);
            const double d = one_plus_z * vel[i] / 100.0 / efunc;      // spatial_distortion, :139
            double v = p[dist_dim] + d;
            if (!isinf(L)) v = htb_np_mod(v, L);                         // enforce_periodicity_of_box, model_helpers.py:164-169
            p[dist_dim] = v;
        }
        out[3 * i] = p[0]; out[3 * i + 1] = p[1]; out[3 * i + 2] = p[2];
    }
}
__global__ void __launch_bounds__(256) k_zspace(const double *__restrict__ pos, const double *__restrict__ vpec, long long n,
                                                double efunc, double scale_factor, double Lbox, int wrap, double *__restrict__ out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double err = vpec[i] / 100.0 / efunc / scale_factor;     // pos_err, :322
        double v = pos[i] + err;
        if (wrap) v = htb_np_mod(v, Lbox);
        out[i] = v;
    }
}

extern "C" int htb_return_xyz_formatted_array(const double *x, const double *y, const double *z, int64_t n, const double *period3,
                                              const double *velocity, int32_t distortion_dim, double redshift, double efunc,
                                              double *pos_out)
{
    HTB_GUARD_BEGIN
    if (!period3 || !pos_out || n < 0 || (n > 0 && (!x || !y || !z)) || distortion_dim > 2 || (distortion_dim >= 0 && !velocity)) {
        htb_set_error("htb_return_xyz_formatted_array: bad arguments");
        return 1;
    }
    if (n == 0) return 0;
    cudaStream_t st;
    if (get_stream(&st)) return 1;
    const int blocks = (int)std::min<int64_t>(148 * 8, (n + 255) / 256);
    k_xyz_formatted<<<blocks, 256, 0, st>>>(x, y, z, (long long)n, period3[0], period3[1], period3[2], velocity,
                                            distortion_dim < 0 ? -1 : distortion_dim, 1.0 + redshift, efunc, pos_out);
    HTB_CUDA(cudaGetLastError());
    return 0;
    HTB_GUARD_END
}

extern "C" int htb_apply_zspace_distortion(const double *true_pos, const double *peculiar_velocity, int64_t n,
                                           double redshift, double efunc, double Lbox, int32_t wrap, double *zspace_out)
{
    HTB_GUARD_BEGIN
    if (!zspace_out || n < 0 || (n > 0 && (!true_pos || !peculiar_velocity))) { htb_set_error("htb_apply_zspace_distortion: bad arguments"); return 1; }
    if (n == 0) return 0;
    cudaStream_t st;
    if (get_stream(&st)) return 1;
    const int blocks = (int)std::min<int64_t>(148 * 8, (n + 255) / 256);
    k_zspace<<<blocks, 256, 0, st>>>(true_pos, peculiar_velocity, (long long)n, efunc, 1.0 / (1.0 + redshift), Lbox, wrap ? 1 : 0, zspace_out);
    HTB_CUDA(cudaGetLastError());
    return 0;
    HTB_GUARD_END
}

// ------------------------------------------------------------------ mesh-only entry points
extern "C" int htb_mesh_cell_ids(int32_t ndim, const double *x, const double *y, const double *z, int64_t stride, int64_t n,
                                 const double *cell_size, const int32_t *ndivs, int64_t *ids_out, uint32_t flags)
{
    HTB_GUARD_BEGIN
    if ((ndim != 2 && ndim != 3) || !cell_size || !ndivs || !ids_out) { htb_set_error("htb_mesh_cell_ids: bad arguments"); return 1; }
    Call c;
    if (c.begin()) return 1;
    c.flags = flags;
    const double *src[3] = {x, y, z}, *dev[3] = {nullptr, nullptr, nullptr};
    int64_t ds = 1;
    if (c.stage_coords(src, ndim, stride, n, dev, &ds)) return 1;
    int64_t *ids = nullptr;
    if (c.ws.alloc((void **)&ids, sizeof(int64_t) * (size_t)(n > 0 ? n : 1))) return 1;
    int nd[3] = {ndivs[0], ndivs[1], ndim == 3 ? ndivs[2] : 1};
    if (htb_ref_cell_ids(c.st, ndim, dev, ds, n, cell_size, nd, ids, &c.launches)) return 1;
    if (n > 0) HTB_CUDA(cudaMemcpyAsync(ids_out, ids, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, c.st));
    HTB_CUDA(cudaStreamSynchronize(c.st));
    return 0;
    HTB_GUARD_END
}

extern "C" int htb_mesh_cell_id_indices(int32_t ndim, const double *x, const double *y, const double *z, int64_t stride,
                                        int64_t n, const double *cell_size, const int32_t *ndivs,
                                        int64_t *out, uint32_t flags)
{
    HTB_GUARD_BEGIN
    if ((ndim != 2 && ndim != 3) || !cell_size || !ndivs || !out) { htb_set_error("htb_mesh_cell_id_indices: bad arguments"); return 1; }
    Call c;
    if (c.begin()) return 1;
    c.flags = flags;
    const double *src[3] = {x, y, z}, *dev[3] = {nullptr, nullptr, nullptr};
    int64_t ds = 1;
    if (c.stage_coords(src, ndim, stride, n, dev, &ds)) return 1;
    FineGrid f{};
    f.dim = ndim;
    f.ncells = 1;
    for (int d = 0; d < 3; ++d) {
        const bool on = d < ndim;
        f.nd[d] = on ? ndivs[d] : 1; f.m[d] = 1; f.nf[d] = f.nd[d];
        f.cs[d] = on ? cell_size[d] : 1.0; f.h[d] = f.cs[d];
        f.period[d] = f.cs[d] * f.nd[d];
        f.ncells *= f.nf[d];
    }
    SortedSample s;
    if (htb_sort_sample(c.st, c.ws, f, dev, ds, n, nullptr, 0, false, 1.0e150, s, &c.launches)) return 1;
    std::vector<uint32_t> off((size_t)f.ncells + 1);
    HTB_CUDA(cudaMemcpyAsync(off.data(), s.off, sizeof(uint32_t) * off.size(), cudaMemcpyDeviceToHost, c.st));
    HTB_CUDA(cudaStreamSynchronize(c.st));
    for (size_t i = 0; i < off.size(); ++i) out[i] = (int64_t)off[i];
    return 0;
    HTB_GUARD_END
}

extern "C" int htb_cell1_work(const htb_mesh_geom *mesh,
                              const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                              const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                              double *work_out, uint32_t flags)
{
    HTB_GUARD_BEGIN
    if (!mesh || !work_out) { htb_set_error("htb_cell1_work: bad arguments"); return 1; }
    Call c;
    if (c.begin()) return 1;
    const double *c1[3] = {x1, y1, z1}, *c2[3] = {x2, y2, z2};
    if (c.setup(mesh, 1, false, c1, stride1, n1, nullptr, c2, stride2, n2, nullptr, 0, false, 0, 0, flags)) return 1;
    if (c.prepared) return 0;
    double *work_dev = nullptr;
    int64_t nc1 = 0;
    if (htb_reference_work(c.st, c.ws, c.G, c.s1, c.s2, &work_dev, nullptr, &nc1, &c.launches)) return 1;
    HTB_CUDA(cudaMemcpyAsync(work_out, work_dev, sizeof(double) * (size_t)nc1, cudaMemcpyDeviceToHost, c.st));
    HTB_CUDA(cudaStreamSynchronize(c.st));
    return 0;
    HTB_GUARD_END
}

// ------------------------------------------------------------------ FP64 issue-rate microbenchmark
// 8 independent dependent-chains of alternating DADD / DMUL per thread, no FMA (compiled with
// -fmad=false): measures the non-FMA FP64 instruction-lane rate the roofline is quoted against.
__global__ void __launch_bounds__(256) k_fp64_rate(double *out, int iters, double a, double b)
{
    double v0 = a + threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v0 = v0 + a; v1 = v1 + a; v2 = v2 + a; v3 = v3 + a; v4 = v4 + a; v5 = v5 + a; v6 = v6 + a; v7 = v7 + a;
            v0 = v0 * b; v1 = v1 * b; v2 = v2 * b; v3 = v3 * b; v4 = v4 * b; v5 = v5 * b; v6 = v6 * b; v7 = v7 * b;
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

extern "C" int htb_measure_fp64_rate(double *ops_per_second_out, double *sm_clock_mhz_out)
{
    HTB_GUARD_BEGIN
    cudaStream_t st;
    if (get_stream(&st)) return 1;
    int dev = 0, sms = 0, clk = 0;
    HTB_CUDA(cudaGetDevice(&dev));
    HTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    HTB_CUDA(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double *buf = nullptr;
    HTB_CUDA(cudaMalloc((void **)&buf, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    HTB_CUDA(cudaEventCreate(&e0));
    HTB_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        HTB_CUDA(cudaEventRecord(e0, st));
        k_fp64_rate<<<blocks, threads, 0, st>>>(buf, iters, 1.0000001, 0.9999999);
        HTB_CUDA(cudaEventRecord(e1, st));
        HTB_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * threads * (double)iters * 8.0 * 16.0;
        const double rate = ops / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    if (ops_per_second_out) *ops_per_second_out = best;
    if (sm_clock_mhz_out) *sm_clock_mhz_out = clk / 1000.0;
    return 0;
    HTB_GUARD_END
}
