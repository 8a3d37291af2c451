/*
 * halotools_b200 — C ABI of the B200-native pair-counting engine.
 *
 * The reference (astropy/halotools) has no C ABI: its native boundary is the
 * Cython engine call
 *     engine(double_mesh, x1, y1, z1, x2, y2, z2, bins..., (first_cell1, last_cell1)) -> counts
 * (/root/reference/halotools/mock_observables/pair_counters/cpairs/npairs_3d_engine.pyx:17).
 * Every htb_*_engine entry point below replaces exactly one such engine and takes
 * the same information as plain pointers and sizes:
 *     double_mesh  -> htb_mesh_geom   (the scalars the engines read from the mesh object,
 *                                      npairs_3d_engine.pyx:47-96; the per-point cell
 *                                      assignment / sort is done on the GPU)
 *     x1in..z2in   -> base pointers + element stride + count (UNSORTED, as the engine gets them)
 *     cell1_tuple  -> first_cell1, last_cell1 (reference mesh1 cell ids, half-open)
 *     return value -> caller-allocated output array.
 * Pointers are HOST pointers unless HTB_FLAG_DEVICE_INPUT is set, in which case the
 * coordinate / weight arrays are device pointers on the current CUDA device (outputs
 * and bins are always host, unless HTB_FLAG_DEVICE_OUTPUT is set).  All functions return 0 on success, non-zero on error
 * (htb_last_error() gives the message).  No torch types appear here.
 *
 * Threading: device, stream, shard and upload-cache settings are per calling thread; every engine call is
 * synchronous (it returns when its outputs are in host memory) unless HTB_FLAG_DEVICE_OUTPUT is set.  Calls from several threads are fine on DIFFERENT
 * devices; on one device they must be serialised by the caller (the pinned staging ring for pageable inputs and
 * large outputs is shared per device) - the reference's own parallelism (multiprocessing over cell ranges) maps
 * to one process per GPU (htb_set_shard) instead.
 */
#ifndef HALOTOOLS_B200_H
#define HALOTOOLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTB_ABI_VERSION 1

/* flags */
#define HTB_FLAG_DEVICE_INPUT 1u   /* coordinate/weight pointers are device pointers          */
#define HTB_FLAG_NO_CULL      2u   /* visit exactly the reference's cell windows (no pruning) */
#define HTB_FLAG_GENERIC      4u   /* force the generic (literal top-down scan) kernels       */
#define HTB_FLAG_NO_TMA       8u   /* stage sample2 tiles with ld.global/st.shared instead of cp.async.bulk */
#define HTB_FLAG_UNIFORM_MASS 32u  /* mean_delta_sigma: all particle masses equal m2[0] (scalar effective_particle_masses) */
#define HTB_FLAG_COLUMN_SUM   64u  /* mean_delta_sigma: delta_sigma_out is f64[nrp-1], the sums of the per-object rows over this call's galaxies (per_object=False needs nothing else) */
#define HTB_FLAG_CACHE_SAMPLE1 128u /* sample1's host coordinate arrays take part in the upload cache (htb_cache_begin) */
#define HTB_FLAG_CACHE_SAMPLE2 256u /* ... sample2's */
#define HTB_FLAG_DEVICE_OUTPUT 1024u /* the output array is a DEVICE pointer on the current device and the call is ASYNCHRONOUS: everything is
                                       enqueued on the thread's stream and the function returns without a host synchronisation (stats are not
                                       filled).  Host input arrays must stay alive and unmodified until the caller synchronises
                                       (htb_stream_synchronize).  Supported by the counters a statistic chains: npairs_3d, npairs_xy_z,
                                       npairs_s_mu, marked_npairs_3d, marked_npairs_xy_z, weighted_npairs_xy, mean_delta_sigma.      */
#define HTB_FLAG_EARLY_EXIT 2048u  /* warps of the counting kernel that run out of work retire at once instead of waiting to help with late exact
                                       re-evaluations: lets the next kernel, enqueued on ANOTHER stream, fill this launch's tail (results unchanged) */
#define HTB_FLAG_PREPARE 4096u     /* run only the set-up of the call (upload, mesh sorts, multi-GPU cut) inside an htb_cache_begin/end scope and
                                       return: the same call without the flag then finds its sorted samples in the caches.  A statistic prepares
                                       all its counts first and launches the count kernels afterwards, so that no mesh sort waits behind a
                                       persistent count kernel that holds every SM */
#define HTB_FLAG_NO_SYM       16u  /* auto-correlations: evaluate (i,j) and (j,i) separately, as the reference does */
#define HTB_FLAG_PARTITION_SUM 512u /* [first_cell1, last_cell1) is one part of a partition of the mesh1 cells whose results the
                                       caller SUMS (multi-GPU shards): auto-correlations may then keep the symmetric shortcut
                                       (each unordered zero-shift pair evaluated once, from the point with the smaller sorted
                                       index, and counted twice), whose per-part counts differ from the reference's per-range
                                       counts although their sum over the partition is identical.  Without the flag a partial
                                       range returns exactly what the reference engine returns for that cell1_tuple. */

/* Scalars of RectangularDoubleMesh / RectangularDoubleMesh2D
 * (/root/reference/halotools/mock_observables/pair_counters/rectangular_mesh.py:228-374,
 *  rectangular_mesh_2d.py:147-250).  For ndim == 2 only entries [0], [1] are read.  */
typedef struct htb_mesh_geom {
    int32_t ndim;            /* 3, or 2 for the surface-density mesh                        */
    int32_t pbc;             /* double_mesh._PBCs                                            */
    int32_t ndivs1[3];       /* mesh1.num_{x,y,z}divs                                        */
    int32_t ndivs2[3];       /* mesh2.num_{x,y,z}divs  (a multiple of ndivs1)                */
    int32_t cover[3];        /* ceil(search_length / mesh2 cell size), engine.pyx:74-79      */
    int32_t reserved;
    double  period[3];       /* {x,y,z}period                                                */
    double  cell1_size[3];   /* mesh1.{x,y,z}cell_size                                       */
    double  cell2_size[3];   /* mesh2.{x,y,z}cell_size                                       */
    double  search[3];       /* search_{x,y,z}length                                         */
} htb_mesh_geom;

/* Work / timing counters filled by every engine call (all optional: pass NULL). */
typedef struct htb_stats {
    double   pairs_evaluated;   /* W_gpu: (i, j) pairs whose separation the kernel computed  */
    double   pairs_reference;   /* W_ref: pairs the reference loop nest would visit          */
    float    ms_h2d;            /* host->device copies                                       */
    float    ms_mesh;           /* cell assignment + counting sort of both samples           */
    float    ms_count;          /* the pair-counting kernel(s)                               */
    float    ms_total;          /* whole call, CUDA-event timed on the engine's stream       */
    int32_t  kernel_launches;   /* kernels launched by this call                             */
    int32_t  tiles;             /* sample1 tiles processed                                   */
    int32_t  tiles_redone;      /* tiles re-evaluated by the exact path (edge-ambiguous key) */
    int32_t  refine1[3];        /* sub-divisions of each reference mesh1 cell                */
    int32_t  refine2[3];        /* sub-divisions of each reference mesh2 cell                */
    int32_t  path;              /* 0 generic, 1 fast queue kernel, 2 cell-resolved, 3 BinQ   */
} htb_stats;

const char *htb_last_error(void);
int  htb_abi_version(void);
/* number of visible CUDA devices (0 if none / driver missing) */
int  htb_device_count(void);
/* select the CUDA device used by subsequent calls from this thread */
int  htb_set_device(int device);
/* use an existing CUDA stream (cudaStream_t as void*) for subsequent calls; NULL = library stream */
int  htb_set_stream(void *cuda_stream);

/* Upload cache: between htb_cache_begin() and htb_cache_end() the host coordinate arrays this thread's engine calls
 * bring to the device stay resident, keyed by (pointers, stride, count): the DD, DR and RR counts of one tpcf() move
 * every sample across PCIe once (the reference re-gathers every sample in every engine call, tpcf.py:76-113,164-205).
 * Only samples flagged HTB_FLAG_CACHE_SAMPLE1/2 take part; the caller must not modify or free those arrays in
 * between.  htb_cache_end() frees the device copies.                                                          */
int  htb_cache_begin(void);
int  htb_cache_end(void);

/* Multi-GPU sharding (one process per GPU): after htb_set_shard(rank, world) every engine call of this thread
 * processes only rank's share of the reference mesh1 cells in [first_cell1, last_cell1) - contiguous cell ranges as
 * in the reference's _cell1_parallelization_indices (mesh_helpers.py:183-221), with the cut points placed on the
 * device so that every rank gets the same predicted work (pairs the reference would visit) instead of the same
 * number of cells.  The caller sums the outputs of the ranks (npairs_3d.py:145).  world = 1 switches it off.  */
int  htb_set_shard(int rank, int world);

/* npairs_3d_engine.pyx:17 — counts[k] = #{(i,j): dx^2+dy^2+dz^2 <= rbins[k]^2}, int64[nb]. */
int htb_npairs_3d_engine(const htb_mesh_geom *mesh,
                         const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                         const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                         const double *rbins, int32_t nb,
                         int64_t first_cell1, int64_t last_cell1,
                         int64_t *counts_out, uint32_t flags, htb_stats *stats);

/* npairs_xy_z_engine.pyx:17 — counts[k,g] = #{dx^2+dy^2 <= rp[k]^2 and dz^2 <= pi[g]^2}, int64[nrp*npi]. */
int htb_npairs_xy_z_engine(const htb_mesh_geom *mesh,
                           const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                           const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                           const double *rp_bins, int32_t nrp, const double *pi_bins, int32_t npi,
                           int64_t first_cell1, int64_t last_cell1,
                           int64_t *counts_out, uint32_t flags, htb_stats *stats);

/* npairs_s_mu_engine.pyx:18 — mu_bins are the sin(theta_los) edges the reference front-end
 * passes down (npairs_s_mu.py:174-175); output int64[ns*nmu], 2-D cumulative.               */
int htb_npairs_s_mu_engine(const htb_mesh_geom *mesh,
                           const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                           const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                           const double *s_bins, int32_t ns, const double *mu_bins, int32_t nmu,
                           int64_t first_cell1, int64_t last_cell1,
                           int64_t *counts_out, uint32_t flags, htb_stats *stats);

/* marked_npairs_3d_engine.pyx:22 — counts[k] = sum f_id(w1_i, w2_j) over dsq <= rbins[k]^2, f64[nb].
 * w1, w2: row-major (n, nw) weights in the SAME (unsorted) order as the coordinates.          */
int htb_marked_npairs_3d_engine(const htb_mesh_geom *mesh,
                                const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                const double *w1, const double *w2, int32_t nw, int32_t weight_func_id,
                                const double *rbins, int32_t nb,
                                int64_t first_cell1, int64_t last_cell1,
                                double *counts_out, uint32_t flags, htb_stats *stats);

/* mean_delta_sigma_engine.pyx:19 — per-object excess surface density, f64[n1*(nrp-1)], rows in
 * the INPUT order of sample1 (the engine un-sorts at exit, :184-185); rows of galaxies whose
 * mesh1 cell is outside [first_cell1, last_cell1) are zero, as in each reference worker.      */
int htb_mean_delta_sigma_engine(const htb_mesh_geom *mesh,
                                const double *x1, const double *y1, int64_t stride1, int64_t n1,
                                const double *x2, const double *y2, int64_t stride2, const double *m2, int64_t n2,
                                const double *rp_bins, int32_t nrp,
                                int64_t first_cell1, int64_t last_cell1,
                                double *delta_sigma_out, uint32_t flags, htb_stats *stats);

/* marked_npairs_xy_z_engine.pyx:21 (marked_cpairs/) - counts[k,g] = sum f_id(w1_i, w2_j) over pairs with
 * dx^2+dy^2 <= rp[k]^2 and dz^2 <= pi[g]^2, f64[nrp*npi] (:209-225).  Weights as in the marked 3-D engine.  */
int htb_marked_npairs_xy_z_engine(const htb_mesh_geom *mesh,
                                  const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                  const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                  const double *w1, const double *w2, int32_t nw, int32_t weight_func_id,
                                  const double *rp_bins, int32_t nrp, const double *pi_bins, int32_t npi,
                                  int64_t first_cell1, int64_t last_cell1,
                                  double *counts_out, uint32_t flags, htb_stats *stats);

/* npairs_per_object_3d_engine.pyx:17 (cpairs/) - counts[i,k] = #{j: dsq_ij <= rbins[k]^2} for every sample1 point,
 * int64[n1*nb], rows in the INPUT order of sample1 (the engine un-sorts at exit, :209-213); rows of points whose
 * mesh1 cell is outside [first_cell1, last_cell1) are zero, as in each reference worker.  nb <= 64.
 * (npairs_projected_engine.pyx:17 needs no entry point of its own: it is htb_npairs_xy_z_engine with the single
 * pi edge pi_max, :184-189.)                                                                                       */
int htb_npairs_per_object_3d_engine(const htb_mesh_geom *mesh,
                                    const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                    const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                    const double *rbins, int32_t nb,
                                    int64_t first_cell1, int64_t last_cell1,
                                    int64_t *counts_out, uint32_t flags, htb_stats *stats);

/* weighted_npairs_xy_engine.pyx:17 (surface_density/engines/) - 2-D mesh (mesh->ndim == 2):
 * counts[k] = sum w2_j over pairs with dx^2+dy^2 <= rp[k]^2, f64[nrp] (:150-175).             */
int htb_weighted_npairs_xy_engine(const htb_mesh_geom *mesh,
                                  const double *x1, const double *y1, int64_t stride1, int64_t n1,
                                  const double *x2, const double *y2, int64_t stride2, int64_t n2,
                                  const double *w2, const double *rp_bins, int32_t nrp,
                                  int64_t first_cell1, int64_t last_cell1,
                                  double *counts_out, uint32_t flags, htb_stats *stats);

/* weighted_npairs_per_object_xy_engine.pyx:17 (surface_density/engines/) - 2-D mesh: counts[i,k] = sum w2_j over the
 * sample2 points with dx^2+dy^2 <= rp[k]^2 of sample1 point i, f64[n1*nrp], rows in the INPUT order of sample1
 * (:150-190; the counter under total_mass_enclosed_per_cylinder, mass_in_cylinders.py:222).  nrp <= 48.            */
int htb_weighted_npairs_per_object_xy_engine(const htb_mesh_geom *mesh,
                                             const double *x1, const double *y1, int64_t stride1, int64_t n1,
                                             const double *x2, const double *y2, int64_t stride2, int64_t n2,
                                             const double *w2, const double *rp_bins, int32_t nrp,
                                             int64_t first_cell1, int64_t last_cell1,
                                             double *counts_out, uint32_t flags, htb_stats *stats);

/* npairs_jackknife_3d_engine.pyx:20 (cpairs/) - counts[s,k] = sum over pairs with dsq <= rbins[k]^2 of
 * jweight(s, jtag1_i, jtag2_j, w1_i, w2_j) (:237-291): s = 0 is the full sample, s >= 1 leaves sub-volume s out.
 * w1, w2: one weight per point; jtags1, jtags2: int64 tags in [1, N_samples]; output f64[(N_samples+1)*nb].
 * HOST arrays only.                                                                                            */
int htb_npairs_jackknife_3d_engine(const htb_mesh_geom *mesh,
                                   const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                   const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                   const double *w1, const double *w2, const int64_t *jtags1, const int64_t *jtags2,
                                   int32_t n_samples, const double *rbins, int32_t nb,
                                   int64_t first_cell1, int64_t last_cell1,
                                   double *counts_out, uint32_t flags, htb_stats *stats);

/* npairs_jackknife_xy_z_engine.pyx:20 (cpairs/) - the same on (rp, pi) bins, f64[(N_samples+1)*nrp*npi] (:222-246);
 * more than 48 cells per point row (rp_pi_tpcf_jackknife) keep the rows in global memory.                       */
int htb_npairs_jackknife_xy_z_engine(const htb_mesh_geom *mesh,
                                     const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                                     const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                                     const double *w1, const double *w2, const int64_t *jtags1, const int64_t *jtags2,
                                     int32_t n_samples, const double *rp_bins, int32_t nrp, const double *pi_bins, int32_t npi,
                                     int64_t first_cell1, int64_t last_cell1,
                                     double *counts_out, uint32_t flags, htb_stats *stats);

/* RectangularMesh cell assignment alone (rectangular_mesh.py:19-22,211-225): writes the
 * reference cell id of every point (int64[n]) — used by the mesh parity tests.               */
int htb_mesh_cell_ids(int32_t ndim, const double *x, const double *y, const double *z, int64_t stride, int64_t n,
                      const double *cell_size, const int32_t *ndivs, int64_t *cell_ids_out, uint32_t flags);

/* RectangularMesh.cell_id_indices of a sample (int64[ncells+1]) computed by the GPU counting sort. */
int htb_mesh_cell_id_indices(int32_t ndim, const double *x, const double *y, const double *z, int64_t stride, int64_t n,
                             const double *cell_size, const int32_t *ndivs, int64_t *cell_id_indices_out, uint32_t flags);

/* Predicted work (pairs the reference would visit) per reference mesh1 cell, f64[ncells1]:
 * the partitioner input for sharding mesh1 cells over ranks (mesh_helpers.py:183-221).        */
int htb_cell1_work(const htb_mesh_geom *mesh,
                   const double *x1, const double *y1, const double *z1, int64_t stride1, int64_t n1,
                   const double *x2, const double *y2, const double *z2, int64_t stride2, int64_t n2,
                   double *work_out, uint32_t flags);

/* Host helper for the front-ends' bounds checks (mock_observables_helpers.py:25-71 enforce_sample_respects_pbcs):
 * minimum and maximum of `cols` adjacent strided columns of a row-major host matrix, one threaded pass.
 * NaNs anywhere make every result NaN (so that comparisons fail as they do in numpy).            */
int htb_host_minmax(const double *base, int64_t n, int64_t stride, int32_t cols, double *min_out, double *max_out);

/* The same extrema for a DEVICE-resident row-major matrix (cols <= 3; `base_dev` is a device pointer): one HBM-bound
 * pass on the GPU, on the library's current stream; results land in host memory.                  */
int htb_device_minmax(const double *base_dev, int64_t n, int64_t stride, int32_t cols, double *min_out, double *max_out);

/* K3 - the two-point estimators over DEVICE-resident cumulative count tables (tpcf_estimators.py:14-119 _TP_estimator /
 * _TP_estimator_crossx, including the np.diff calls of tpcf.py:76-113 and rp_pi_tpcf.py:296-330 and wp's 2 * xi * pi_max,
 * wp.py:219-221): one tiny kernel on the thread's stream, so DD / DR / RR counted with HTB_FLAG_DEVICE_OUTPUT, summed over
 * the ranks on the device and combined into xi need ONE host synchronisation per statistic.
 *   n0, n1            edges along the first / second bin axis of every table (n1 = 1: 3-d r bins)
 *   estimator         0 Natural, 1 Davis-Peebles, 2 Hewett, 3 Hamilton, 4 Landy-Szalay;  cross: _TP_estimator_crossx
 *   *_cum             DEVICE int64[n0 * n1] cumulative counts as the engines return them, or NULL
 *   *_diff            DEVICE f64[(n0 - 1) * max(n1 - 1, 1)] differential counts (analytic randoms), read when *_cum is NULL
 *   inv_factor1/2     1 / (ND1 ND2 / (NR1 NR2)), 1 / (ND1 NR2 / (NR1 NR2)) as the reference forms them (for Davis-Peebles
 *                     inv_factor1 = 1 / (ND1 ND2 / (ND1 NR2)))
 *   wp_pi_max         > 0: the output is 2 * xi[:, 0] * pi_max (n1 must be 2)
 *   xi_out            DEVICE f64[(n0 - 1) * max(n1 - 1, 1)]
 *   flag_out          DEVICE int32, OR-ed: bit 0 some RR bin is zero, bit 1 some DR bin is zero (_test_for_zero_division,
 *                     tpcf_estimators.py:165-183; the caller raises the reference's ValueError)                           */
int htb_tp_estimator(int32_t n0, int32_t n1, int32_t estimator, int32_t cross,
                     const int64_t *DD_cum, const int64_t *D1R_cum, const int64_t *D2R_cum, const int64_t *RR_cum,
                     const double *D1R_diff, const double *D2R_diff, const double *RR_diff,
                     double inv_factor1, double inv_factor2, double wp_pi_max,
                     double *xi_out, int32_t *flag_out);

/* Asynchronous host -> device copy of `count` doubles on the thread's stream (pinned: one copy; large pageable arrays:
 * the threaded pinned-chunk ring the engines use).  host_src must stay alive until the stream is synchronised.     */
int htb_upload_f64(const double *host_src, int64_t count, double *dev_dst);

/* The CUDA stream (cudaStream_t as void*) this thread's calls are issued on, and a host wait for it. */
int htb_get_stream(void **stream_out);
int htb_stream_synchronize(void);
/* CUDA-event durations (ms) of the counting kernels of this thread's HTB_FLAG_DEVICE_OUTPUT calls since the previous
 * query, oldest first (a ring of 16); call after htb_stream_synchronize().                                    */
int htb_async_count_times(float *ms_out, int32_t max_out, int32_t *n_out);
/* The same launches by the kernels' own device-side time stamps (first warp in -> last warp out, ms): without the time a
 * launch waits for the blocks of kernels on other streams to retire.  Does not reset the ring: query it BEFORE
 * htb_async_count_times().  -1 where a call launched no counting kernel.                                        */
int htb_async_kernel_spans(float *ms_out, int32_t max_out, int32_t *n_out);
/* ... and the raw stamps, two words per call {first warp in, last warp out} in ns of the device's globaltimer (0, 0: no
 * kernel): counts of one statistic that run side by side on several streams are measured by the union of their spans. */
int htb_async_kernel_stamps(uint64_t *ns_out, int32_t max_out, int32_t *n_out);

/* The input step before the path, for samples that live in HBM: return_xyz_formatted_array
 * (catalog_analysis_helpers.py:108-265) and apply_zspace_distortion (:268-327) as one elementwise kernel each, enqueued on
 * the thread's stream (asynchronous).  ALL pointers except period3 are DEVICE pointers.
 *   x, y, z         f64[n];  period3: host f64[3] (inf = no wrap)
 *   velocity        f64[n] or NULL;  distortion_dim: 0 / 1 / 2 = the coordinate that receives (1 + z) v / 100 / E(z), -1 none
 *   efunc           E(z) = H(z) / H0 of the caller's cosmology (host scalar)
 *   pos_out         f64[n * 3] row-major: the (Npts, 3) sample the pair counters take                                   */
int htb_return_xyz_formatted_array(const double *x, const double *y, const double *z, int64_t n, const double *period3,
                                   const double *velocity, int32_t distortion_dim, double redshift, double efunc,
                                   double *pos_out);
/* zspace = true_pos + v_pec / 100 / E(z) / a, wrapped into [0, Lbox) when wrap != 0; f64[n] device arrays.              */
int htb_apply_zspace_distortion(const double *true_pos, const double *peculiar_velocity, int64_t n,
                                double redshift, double efunc, double Lbox, int32_t wrap, double *zspace_out);

/* Measured FP64 non-FMA issue rate (DADD/DMUL instr-lanes per second) of the current device. */
int htb_measure_fp64_rate(double *ops_per_second_out, double *sm_clock_mhz_out);

#ifdef __cplusplus
}
#endif
#endif /* HALOTOOLS_B200_H */
