// Internal declarations shared by the translation units of libhalotools_b200.so.
// sm_100a only; compiled with -fmad=false so every f64 expression is evaluated
// exactly as written (no FMA contraction) — required for bit-exact integer counts
// against the reference's scalar SSE2 arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/halotools_b200.h"

#define HTB_TILE 64            // default sample1 points per tile (one warp, 2 points per lane)
#define HTB_MAX_DIM 3

struct HtbError {
    std::string msg;
};
void htb_set_error(const char *fmt, ...);

#define HTB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            htb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                  \
                          cudaGetErrorString(e__));                                      \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

// One sample binned on a refinement of a reference mesh (RectangularMesh): every
// reference cell is split into m[d] sub-cells per dimension; the LAST dimension is
// the fastest in the cell id, as in the reference (rectangular_mesh.py:224-225).
struct FineGrid {
    int dim;
    int nd[HTB_MAX_DIM];        // reference num_divs
    int m[HTB_MAX_DIM];         // sub-divisions per reference cell
    int nf[HTB_MAX_DIM];        // nd * m
    double cs[HTB_MAX_DIM];     // reference cell size (python float handed down by the host layer)
    double h[HTB_MAX_DIM];      // cs / m
    double period[HTB_MAX_DIM];
    int64_t ncells;             // prod nf
};

struct SortedSample {
    int64_t n = 0;
    int64_t npad = 0;           // allocation length of each coordinate array (even, >= n + 2)
    FineGrid g{};
    double *c[HTB_MAX_DIM] = {nullptr, nullptr, nullptr};   // SoA coordinates, sorted by fine cell
    double *w = nullptr;        // row-major (n, nw) payload in sorted order (weights / masses) or null
    int nw = 0;
    uint32_t *off = nullptr;    // [ncells + 1] first sorted position of each fine cell
    uint32_t *perm = nullptr;   // [n] sorted position -> input index
    uint32_t *flags = nullptr;  // [1] bit0: some point lies outside [0, period] in some dimension
    uint32_t *cell = nullptr;   // scratch [n]: fine cell id per input point
    uint32_t *count = nullptr;  // scratch [ncells + 2]: points per fine cell (after the sort: the cells' fill counters)
};

struct Workspace;   // stream-ordered allocations of one engine call (freed at the end)

// ---- mesh.cu
int htb_sort_sample(cudaStream_t st, Workspace &ws, const FineGrid &g,
                    const double *const *coords_dev, int64_t stride, int64_t n,
                    const double *w_dev, int nw, bool keep_perm, double pad_value, SortedSample &out, int *launches);
// the same sort in two halves, so that a rank of a sharded call can look at the cell counts of both samples,
// settle its cell range and drop the points outside its window (xwin_dev: device {first reference x-layer, layers})
int htb_sort_begin(cudaStream_t st, Workspace &ws, const FineGrid &g,
                   const double *const *coords_dev, int64_t stride, int64_t n,
                   const double *w_dev, int nw, bool keep_perm, SortedSample &out, int *launches);
int htb_sort_finish(cudaStream_t st, Workspace &ws, const double *const *coords_dev, int64_t stride,
                    const double *w_dev, int nw, double pad_value, const int *xwin_dev, SortedSample &out, int *launches);
int htb_ref_cell_counts_pre(cudaStream_t st, const SortedSample &s, uint32_t *counts_dev /* [prod nd] zeroed */, int *launches);
int htb_exclusive_scan_u32(cudaStream_t st, Workspace &ws, uint32_t *in, uint32_t *out,
                           int64_t n, uint32_t *total_dev, int *launches, bool zero_in = false);
int htb_ref_cell_ids(cudaStream_t st, int dim, const double *const *coords_dev, int64_t stride, int64_t n,
                     const double *cell_size, const int *ndivs, int64_t *ids_dev, int *launches);
int htb_ref_cell_counts(cudaStream_t st, const SortedSample &s, uint32_t *counts_dev /* [prod nd] zeroed */,
                        int *launches);

int htb_device_minmax_launch(cudaStream_t st, const double *base_dev, int64_t n, int64_t stride, int cols,
                             double *part_dev /* [blocks][7] */, int blocks);

// ---- workspace (capi.cu)
struct Workspace {
    cudaStream_t st = nullptr;
    void *ptrs[256];
    int nptrs = 0;
    int alloc(void **p, size_t bytes);
    void release();
};

// numpy's float64 floor-division followed by the reference's clip (rectangular_mesh.py:19-22).  numpy's npy_divmod is
// fmod based: mod = fmod(p, c) is the EXACT remainder, div = (p - mod) / c lands within an ulp of an integer and is
// snapped to it, so for finite p and c > 0 the result is floor(p / c) of the exact quotient, never of the rounded one.
// That integer is found here without fmod's bit-serial loop and without a division: an estimate from the rounded
// reciprocal is at most one off, and the sign of fma(-q, c, p) - one rounding of the exact p - q c - is exact.
// rc = 1.0 / c, hoisted by callers that bin many points with one cell size
__host__ __device__ inline int htb_ref_digitize(double p, double c, double rc, int ndivs)
{
    if (!(fabs(p) < INFINITY)) return 0;                // nan, +-inf: rejected by the bounds check; numpy yields garbage here
    double fl = floor(p * rc);
    if (fl >= (double)ndivs + 1.0) return ndivs - 1;   // the exact quotient is >= ndivs as well
    if (fma(-fl, c, p) < 0.0) fl -= 1.0;
    else if (fma(-(fl + 1.0), c, p) >= 0.0) fl += 1.0;
    // astype(int) then np.where(ip >= num_divs, num_divs - 1, ip); negatives are undefined
    // behaviour in the reference (they index before the first cell) — clamp to cell 0.
    if (!(fl > -1.0)) return 0;
    if (fl >= (double)ndivs) return ndivs - 1;
    return (int)fl;
}

__host__ __device__ inline int htb_ref_digitize(double p, double c, int ndivs) { return htb_ref_digitize(p, c, 1.0 / c, ndivs); }
