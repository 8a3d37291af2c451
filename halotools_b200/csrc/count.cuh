// Parameter blocks of the counting kernels (passed by value as __grid_constant__).
#pragma once
#include "walk.cuh"

#define HTB_NBF 16          // max bins of the fast npairs_3d kernel (one register counter per bin)
#define HTB_MAX_NW 5        // max weights per point (marked_npairs_3d.py:281-326)

struct Fast3Params {
    int nb;                          // number of rbins (<= HTB_NBF); slot s holds bin s - (HTB_NBF - nb)
    int nbias;                       // -(bits(top squared edge) >> 26), low 32 bits: key = (bits(dsq) >> 26) + nbias
    int Hwin;                        // pairs whose dsq high word is below this are re-evaluated exactly (key would wrap)
    int F[HTB_NBF];                  // keys of the squared edges relative to the top edge (pads / zero edges = INT_MIN)
    unsigned long long E[HTB_NBF];   // raw bit patterns of the squared edges (pads = 0)
    unsigned long long E_top;
    unsigned long long *counts;      // [nb] global accumulators
    // (rp, pi) variant only: the keys are those of dx^2 + dy^2; a pair takes part if dz^2 <= pi_top_sq
    double pi_top_sq;                // squared top pi edge
    unsigned long long Epi0;         // raw bits of the squared lower pi edge (two pi edges only)
    int Hz0;                         // dz^2 high words <= this send the group to the exact path (-1: one pi edge)
    unsigned long long *counts0;     // [nb] accumulators of the lower pi edge column, or null
    // marked variant (MarkedQ, weight function w1 * w2): [HTB_NBF + 1] float sums: slot s = pairs whose lowest
    // satisfied edge is slot s (differential), slot HTB_NBF = all pairs inside the top edge
    double *fsums;
};

// mean_delta_sigma fast path (one mass for all particles, <= HTB_NBF rp edges)
struct DSQParams {
    int nrp;                         // number of rp edges; edge k lives in slot k + (HTB_NBF - nrp)
    int nbias;                       // key = (bits(dsq) >> 26) + nbias, relative to the top squared edge
    int Hwin;                        // dsq high words below this: key would wrap -> decided exactly
    int F1;                          // key of the second edge from the top (lower edge of the top bin)
    unsigned Tspan;                  // -F1 - 1 (0 disables the in-register top bin)
    unsigned long long E[HTB_NBF];   // raw bit patterns of the squared edges (pads = 0)
    const double *e0, *e1;           // device: squared edges, ln(rp[k+1]/rp[k])
    double mass;
    double *out;                     // (n1, nrp - 1) rows in input order
    const uint32_t *perm1;           // sorted position -> input row
};

// mean_delta_sigma, cell-resolved fast path (DSigmaR): one mass for all particles, 2 <= nrp <= HTB_NBF edges
struct DSRParams {
    int nrp;                         // number of rp edges
    double Ed[HTB_NBF];              // squared edges, ascending; entries >= nrp hold +inf
    double tiny2;                    // a cell that may hold a pair closer than this (squared) takes the exact path
    int renorm;                      // running products are renormalised after this many factors (a multiple of 8)
    double mass;
    const double *e0, *e1;           // device: squared edges, ln(rp[k+1]/rp[k])
    double *out;                     // (n1, nrp - 1) rows in input order
    const uint32_t *perm1;           // sorted position -> input row
};

struct GenParams {
    int n0, n1;                      // number of edges along the first / second bin axis
    int nhist;                       // histogram length
    int nw, wfunc;                   // marked: weights per point, weight_func_id
    const double *e0, *e1;           // device arrays: squared edges (DSigma: e1 = ln(rp[k+1]/rp[k]))
    double max0, max1;               // s_mu: max squared edges
    unsigned long long *counts;      // integer histogram (global)
    double *fcounts;                 // float histogram / per-object output (global)
    const uint32_t *perm1;           // DSigma: sorted position -> input row
};

#define HTB_JK_SHARED_CELLS 48
// BinQ (binq.cu): integer counts on any monotone edges, one or two bin axes, differential histogram
struct BinQParams {
    int n0, n1;                      // edges along the first (r / rp / s) and second (none: 1 / pi / mu) bin axis
    int H0, H1;                      // high words of the top squared edges: the hot loop's conservative range test
    const unsigned long long *edges; // device: raw bits of the squared edges, n0 of axis 0 then n1 of axis 1, ascending;
                                     // followed by the two lookup tables (T[0], T[1] bytes, each padded to 8)
    unsigned long long *counts;      // device: [n0 * n1] differential histogram (lowest satisfied edge per axis)
    // per axis: lut[(bits >> S) - kmin] = (first edge whose key (bits >> S) is >= that of the value) << 1 | (an edge
    // has exactly this key: exact compares needed); T entries
    int S[2], T[2];
    unsigned kmin[2];
    int nzero[2];                    // per axis: number of leading zero edges (index of a positive value below the table)
    // weighted modes (marked_npairs_xy_z, marked_npairs_3d with general marks, weighted_npairs_xy)
    int nw, wfunc;                   // weights per point; weight_func_id, or -1: the weight is sample2's w2[0] alone
    double *fcounts;                 // device: [n0 * n1] differential float sums
    // jackknife modes with more than HTB_JK_SHARED_CELLS cells per point row: the rows of every warp live in global
    // memory (64 rows of (n0 * n1) | 1 doubles per warp, zero-filled by the host)
    double *grows;
    unsigned grows_warps;            // warps the allocation serves
    // per-object mode (npairs_per_object_3d)
    unsigned long long *rows;        // device: (n1_points, n0) cumulative counts in INPUT order
    const uint32_t *perm1;           // sorted position -> input row
};

int htb_fast3_ppl();            // sample1 points per lane of the fast kernel (its tiles hold 32x that)
int htb_launch_fast3(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const Fast3Params &P, int *launches);
int htb_launch_markedq(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const Fast3Params &P, int *launches);
int htb_launch_fastxyz(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const Fast3Params &P, int *launches);
int htb_launch_dsq(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const DSQParams &P, int *launches);
int htb_launch_dsr(cudaStream_t st, const WalkGeom &G, const WalkArrays &A, const DSRParams &P, int *launches);
int htb_launch_binq(cudaStream_t st, int kind /* 0 r, 1 (rp, pi), 2 (s, mu), 3 2-D rp */, int mode /* 0 counts, 1 weighted, 2 per object */, const WalkGeom &G, const WalkArrays &A, const BinQParams &P, int *launches);
int htb_launch_gen(cudaStream_t st, int kind, const WalkGeom &G, const WalkArrays &A, const GenParams &P, int *launches);
int htb_build_tiles(cudaStream_t st, Workspace &ws, const WalkGeom &G, const SortedSample &s1,
                    int64_t first_cell1, int64_t last_cell1, const long long *range_dev /* device {first, last} or null */,
                    uint2 **tiles_out, uint32_t **ntiles_dev_out,
                    int64_t *max_tiles_out, int *launches);
int htb_shard_range(cudaStream_t st, const double *work_dev, int64_t first_cell1, int64_t last_cell1,
                    int rank, int world, long long *range_dev, int *launches);
int htb_reference_work(cudaStream_t st, Workspace &ws, const WalkGeom &G, const SortedSample &s1,
                       const SortedSample &s2, double **work_dev_out, double **balance_dev_out /* may be null */,
                       int64_t *ncell1_out, int *launches, bool presort = false /* samples only went through htb_sort_begin */);
int htb_shard_windows(cudaStream_t st, const long long *range_dev, const WalkGeom &G, int *xwin_dev /* [4] */, int *launches);
