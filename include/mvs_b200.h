/*
 * mvs_b200.h -- C ABI of the B200-native registration-and-fusion engine.
 *
 * Drop-in boundary for the hot path of multiview-stitcher (reference commit
 * 629f72d; paths below are relative to src/multiview_stitcher/).  Plain C:
 * pointers and sizes only, no torch / C++ types.  All `d_` / "device" pointers
 * are CUDA device pointers on the current device; `stream` is a cudaStream_t
 * passed as void* (NULL = default stream).  Every entry point returns 0 on
 * success and a negative mvs_status otherwise; the message of the last failure
 * on the calling thread is available from mvs_last_error().  Entry points are
 * thread-safe (the reference calls its hooks from dask worker threads,
 * registration.py:2680-2692) as long as distinct plans/workspaces are used per
 * thread; work is enqueued on `stream` and is asynchronous unless stated.
 *
 * The reference has no FFI of its own for this path (it is pure Python calling
 * scipy / scikit-image); each entry point cites the Python interface whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 */
#ifndef MVS_B200_H
#define MVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVS_ABI_VERSION 1

typedef enum {
  MVS_OK = 0,
  MVS_ERR_INVALID = -1,     /* bad argument                                  */
  MVS_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail)   */
  MVS_ERR_UNSUPPORTED = -3, /* valid request the engine cannot serve         */
  MVS_ERR_NO_DEVICE = -4    /* no sm_100 device visible                      */
} mvs_status;

typedef enum { MVS_U8 = 0, MVS_U16 = 1, MVS_F32 = 2 } mvs_dtype;

/* fusion_func selector: fusion/_core.py:61-94, :42-58, :97-131 */
typedef enum {
  MVS_FUSE_WAVG = 0, /* weighted_average_fusion with blending weights        */
  MVS_FUSE_MAX = 1,  /* max_fusion                                            */
  MVS_FUSE_MEAN = 2  /* simple_average_fusion                                 */
} mvs_fusion_mode;

const char* mvs_last_error(void);
int mvs_abi_version(void);
/* Fills name (<= cap bytes), SM count, compute capability; MVS_ERR_NO_DEVICE
 * when no CUDA device is usable. */
int mvs_device_info(char* name, int cap, int* sm_count, int* cc_major, int* cc_minor);
/* sizeof(mvs_view_xform), sizeof(mvs_chunk) as compiled (binding self-check). */
int mvs_struct_sizes(int* view_xform_bytes, int* chunk_bytes);

/* ------------------------------------------------------------------------
 * (ii) fused affine resample + blending weights + weighted accumulation
 *
 * Replaces, for a batch of output chunks in ONE launch, what the reference
 * does per chunk in fuse_np (fusion/_core.py:1513-1733):
 *   transform_sim -> scipy.ndimage.affine_transform(mode="constant", cval=NaN)
 *       (transformation.py:15-148)              matrix/offset below
 *   get_blending_weights (weights.py:391-511)   wmatrix/woffset + table below
 *   mask + normalize_weights (_core.py:1647-1649, weights.py:325-345)
 *   fusion_func (_core.py:42-131), trim (:1687-1711), nan_to_num + cast (:1713)
 *
 * All index triples are (z, y, x); 2-D problems use z extent 1.
 * ---------------------------------------------------------------------- */

/* One (chunk, view) pairing: what one transform_sim call sees. */
typedef struct {
  const void* data;   /* device ptr to element [0,0,0] of the view window    */
  int32_t dtype;      /* mvs_dtype of the view                                */
  int32_t shape[3];   /* window extent                                        */
  int64_t stride[3];  /* element strides of the window                        */
  /* sample position in window pixels = matrix * (chunk px + halo) + offset,
   * exactly the matrix/offset transform_sim hands to scipy (already rounded
   * to 10 decimals and snapped, transformation.py:72-83); row-major 3x3.     */
  double matrix[9];
  double offset[3];
  /* same for the 5^ndim blending-support table (weights.py:465-481)          */
  double wmatrix[9];
  double woffset[3];
  int32_t table;      /* index of the view's table in `tables`                */
  int32_t reserved;
} mvs_view_xform;

/* One output chunk (after trimming). */
typedef struct {
  void* out;          /* device ptr to the chunk's first voxel, or NULL       */
  int32_t out_dtype;  /* mvs_dtype; fused float32 is NaN->0 then C-cast       */
  int32_t shape[3];   /* trimmed chunk extent                                 */
  int64_t stride[3];  /* element strides of `out`, `acc_num`, `acc_den`       */
  int32_t halo[3];    /* trim_overlap_in_pixels: sample index = voxel + halo  */
  int32_t first_xform;/* this chunk's pairings: xforms[first .. first+n)      */
  int32_t n_xforms;   /* in view order (order fixes the float32 sums)         */
  /* optional partial accumulators (multi-GPU partial mode, WAVG only):
   * acc_num = sum_i v_i*b_i, acc_den = sum_i b_i with UN-normalised blending
   * weights (overwritten, not accumulated); `out` may then be NULL.         */
  float* acc_num;
  float* acc_den;
} mvs_chunk;

typedef struct mvs_fuse_plan mvs_fuse_plan;

/* Uploads the work list (host arrays; device pointers inside) and builds the
 * block schedule.  tables: n_tables * 125 floats (5x5x5, z-major; 2-D tables
 * occupy the first 25 entries of their slot), host pointer.
 * ndim 2|3; order 0|1 (interpolation_order, _core.py:797). */
int mvs_fuse_plan_create(mvs_fuse_plan** plan, const mvs_chunk* chunks, int n_chunks,
                         const mvs_view_xform* xforms, int n_xforms,
                         const float* tables, int n_tables, int ndim, int order,
                         int fusion_mode, void* stream);
/* Enqueues the fused kernel for every chunk of the plan. */
int mvs_fuse_plan_run(mvs_fuse_plan* plan, void* stream);
/* Same for chunks [first_chunk, first_chunk + n_chunks) only (order of
 * mvs_fuse_plan_create) -- lets a host pipeline fuse a band of chunks as soon
 * as its tiles have arrived and download it while the next band is fused.
 * A plan must not be run concurrently on two streams. */
int mvs_fuse_plan_run_chunks(mvs_fuse_plan* plan, int first_chunk, int n_chunks, void* stream);
/* Number of kernel launches one mvs_fuse_plan_run issues / blocks scheduled. */
int mvs_fuse_plan_info(const mvs_fuse_plan* plan, int* launches, int64_t* blocks,
                       int64_t* out_voxels);
int mvs_fuse_plan_destroy(mvs_fuse_plan* plan);

/* Finalises partial accumulators: out = cast(nan_to_num(num / (den==0?1:den))).
 * Used after the NCCL sum of (acc_num, acc_den) across ranks. */
int mvs_fuse_finalize(const float* acc_num, const float* acc_den, void* out,
                      int out_dtype, int64_t n, void* stream);

/* Same for a list of boxes (HOST array of mvs_chunk): box i divides its packed,
 * C-ordered accumulators acc_num / acc_den (shape[0]*shape[1]*shape[2] floats
 * each) into the strided output window `out` (element strides `stride`).  The
 * owner of a border chunk calls this after summing the partial sums its
 * neighbours sent (distributed.fuse_tile_partitioned). */
int mvs_fuse_finalize_boxes(const mvs_chunk* boxes, int n_boxes, void* stream);

/* ------------------------------------------------------------------------
 * Post-resample stage on (V, *chunk) float32 stacks (NaN = outside): the
 * arithmetic behind the reference's fusion_func / weights_func hooks
 * (docs/extension_api_fusion.md; fusion/_core.py:1653-1685) and the multi-pass
 * half of content-weighted fusion.  A stack is V contiguous volumes of
 * shape[0]*shape[1]*shape[2] voxels on the device.
 * ---------------------------------------------------------------------- */

/* transform_sim of n_views views onto one chunk grid (fusion/_core.py:1621-1632)
 * and, if d_weights != NULL, the un-normalised blending weights masked by
 * validity (weights.py:391-511, _core.py:1636-1648).  xforms / tables are HOST
 * arrays (device pointers inside); shape / halo as in mvs_chunk. */
int mvs_resample_views(const mvs_view_xform* xforms, int n_views, const float* tables,
                       int n_tables, const int32_t shape[3], const int32_t halo[3], int ndim,
                       int order, float* d_views, float* d_weights, void* stream);

/* weights.normalize_weights (weights.py:325-345), in place. */
int mvs_normalize_weights(float* d_weights, int V, int64_t N, void* stream);

/* scipy.ndimage.gaussian_filter(mode="reflect") of `batch` volumes with the
 * symmetric kernel weights[0..radius] (HOST float64, weights[j] = weight at
 * distance j; scipy's own normalised kernel for sigma, truncate 4). */
int mvs_gaussian_filter(const float* d_in, float* d_out, int batch, const int32_t shape[3],
                        int ndim, const double* weights, int radius, void* stream);

/* weights.content_based (weights.py:22-74): views whose normalised blending
 * weight is < 1e-7 are ignored, W = G2 ~* (v - G1 ~* v)^2 with NaN-normalised
 * Gaussians (weights.py:293-322), then normalised over V.  w1/r1, w2/r2: the two
 * kernels as in mvs_gaussian_filter. */
int mvs_content_based(const float* d_views, const float* d_blending, int V,
                      const int32_t shape[3], int ndim, const double* w1, int r1,
                      const double* w2, int r2, float* d_out_weights, void* stream);

/* weights.content_based_dct (weights.py:77-290): per view and block of block[0..2] voxels
 * (<= 32 per axis; z = 1 in 2-D) the Shannon entropy of the orthonormal DCT-II coefficients
 * with L1 frequency index < r_o, normalised by the block's L2 norm and scaled by 2 / r_o^2
 * (r_o < 0: every coefficient, L1-mean normalisation, weights.py:232-243), raised to
 * `exponent`; blocks with < 20 % valid voxels score 0, NaNs are filled with the block's
 * minimum.  Scores are normalised over the views, interpolated to voxel resolution
 * (order 1, mode "nearest") and normalised again.  d_out_weights: (V, *shape) float32. */
int mvs_content_based_dct(const float* d_views, int V, const int32_t shape[3], int ndim,
                          const int32_t block[3], float r_o, float exponent, float* d_out_weights,
                          void* stream);

/* scipy.ndimage.convolve(input, kernel, mode, cval) of one float32 volume with a small
 * dense kernel (HOST float32, odd extents <= 15; z extent 1 in 2-D): mode 0 = "mirror",
 * 1 = "constant".  float32 accumulation.  d_in != d_out. */
int mvs_convolve(const float* d_in, float* d_out, const int32_t shape[3], const float* kernel,
                 const int32_t kshape[3], int mode, float cval, void* stream);

/* fusion.mv_deconv.multi_view_deconvolution (fusion/mv_deconv.py:251-500) as a fusion_func on
 * (V, *shape) float32 stacks: psi0 = clip(nansum(views * weights)), then n_iterations of the
 * sequential per-view Richardson-Lucy update (forward convolution with kernels1[v], mode
 * mirror; ratio gated by the view's coverage and blending weight; back-projection with the
 * compound kernel kernels2[v], mode constant 1; optional Tikhonov regularisation; clamp to
 * min_value), optional erosion of the union coverage mask by erosion_px face-neighbour
 * steps.  kernels1 / kernels2: HOST arrays of V kernels of kshape (odd, <= 15) -- the PSFs
 * and compound kernels are tiny and built on the host exactly like the reference (:173-245,
 * :373-415).  d_out: float32 volume. */
int mvs_mv_deconvolution(const float* d_views, const float* d_weights, int V, const int32_t shape[3],
                         int ndim, const float* kernels1, const float* kernels2,
                         const int32_t kshape[3], int n_iterations, float lambda_reg, float min_value,
                         int erosion_px, float* d_out, void* stream);

/* fusion_func on stacks: MVS_FUSE_WAVG = weighted_average_fusion(views, blending,
 * fusion_weights or NULL) (_core.py:61-94), MVS_FUSE_MAX (_core.py:42-58),
 * MVS_FUSE_MEAN (_core.py:97-131).  d_out: float32 volume (NaN where the
 * reference yields NaN). */
int mvs_fuse_stack(const float* d_views, const float* d_blending, const float* d_fusion_weights,
                   int V, int64_t N, int fusion_mode, float* d_out, void* stream);

/* fused[trim:-trim] -> nan_to_num -> astype(out_dtype) (_core.py:1687-1713);
 * out_stride in elements. */
int mvs_trim_cast(const float* d_in, const int32_t shape[3], const int32_t trim[3], void* d_out,
                  int out_dtype, const int64_t out_stride[3], void* stream);

/* ------------------------------------------------------------------------
 * (i) batched 2-D/3-D phase-correlation pairwise registration
 *
 * Replaces registration.phase_correlation_registration
 * (registration.py:353-565) and the scikit-image / scipy calls inside it:
 *   rescale_intensity (:382-389), two phase_cross_correlation calls with
 *   normalization "phase" / None and upsample_factor (:410-431), candidate
 *   resampling with scipy.ndimage.affine_transform order 1 (:494-500), mask
 *   statistics (:501-528), structural_similarity (:535-548) and
 *   scipy.stats.spearmanr (:109-111, :551-553).
 * A plan serves pairs of ONE crop shape (z, y, x; z = 1 in 2-D); the stages
 * are batched over pairs / candidates.  The small data-dependent decisions in
 * between (candidate expansion :461-477, the 10 % overlap rule :503, window
 * size :535-536, list semantics :530-533, argmax :558) stay on the host.
 * Outputs are written to HOST arrays; each stage synchronises `stream`.
 * ---------------------------------------------------------------------- */
typedef struct mvs_pc_plan mvs_pc_plan;

/* upsample_factor: 10 (2-D) / 2 (3-D) are the reference defaults (:410-411). */
int mvs_pc_plan_create(mvs_pc_plan** plan, int ndim, const int32_t shape[3], int max_pairs,
                       int upsample_factor);
int mvs_pc_plan_destroy(mvs_pc_plan* plan);
/* region = ceil(1.5 * upsample) samples per axis of the upsampled DFT. */
int mvs_pc_plan_info(const mvs_pc_plan* plan, int* region, int64_t* voxels,
                     int* launches_per_correlate);
/* Test hook: copies the complex work buffer `which` (0: plain cross-power spectrum
 * P left by mvs_pc_correlate, 1: packed spectrum after the inverse passes that
 * store) of one pair to host[2 * voxels]. */
int mvs_pc_debug_copy(const mvs_pc_plan* plan, int which, int pair, float* host);

/* Stage A: per-image statistics and rescale_intensity to [0,1] (NaN kept).
 * fixed/moving: n device pointers to contiguous float32 crops of the plan's
 * shape (NaN = outside).  stats_host[2n][9] (image 2i = fixed i, 2i+1 =
 * moving i): nanmin, nanmax, NaN count, bbox lo z,y,x, bbox hi z,y,x (inclusive)
 * of the non-NaN voxels.  nanmin == nanmax is the caller's constant-image
 * guard (registration.py:1504-1530). */
int mvs_pc_load_pairs(mvs_pc_plan* plan, int n, const float* const* fixed,
                      const float* const* moving, double* stats_host, void* stream);

/* Stage B: forward FFT, normalised / plain cross-power spectrum, inverse FFT,
 * integer peaks and upsampled-DFT samples for the n loaded pairs.
 * peaks_host[n][2][3]: wrapped integer peak (z,y,x); slot 0 = normalization
 * None, slot 1 = "phase".  updft_host[n][2][region^ndim][2]: complex128
 * cross-correlation samples around each peak, C order (z,y,x); sample index
 * argmax |.| minus fix(region/2), divided by upsample, refines the peak. */
int mvs_pc_correlate(mvs_pc_plan* plan, int n, int32_t* peaks_host, double* updft_host,
                     void* stream);

/* Stage C: for each candidate translation t (z,y,x; fixed px -> moving px) of
 * pair cand_pair[i]: stats_host[i][8] = count(mask), count(~isnan(im1t)),
 * bbox lo z,y,x and hi z,y,x (inclusive) of ~isnan(im1t). */
int mvs_pc_candidate_stats(mvs_pc_plan* plan, int n_cand, const int32_t* cand_pair,
                           const double* cand_t, int64_t* stats_host, void* stream);

/* Stage D: SSIM of nan_to_num(im0) vs nan_to_num(im1t) on slices[i] =
 * lo z,y,x, hi z,y,x (exclusive) with an odd window win[i] in 3..7; all
 * candidates of one call share the window size (the kernels are specialised
 * on it; the host batches by window).
 * out_host[i][2] = mean SSIM, nanmax(im1t[slices]) (NaN if all-NaN). */
int mvs_pc_candidate_ssim(mvs_pc_plan* plan, int n_cand, const int32_t* cand_pair,
                          const double* cand_t, const int32_t* slices, const int32_t* win,
                          double* out_host, void* stream);

/* Stage E: Spearman rank correlation (average ranks for ties) of im0[mask] vs
 * (im1t[mask] - 1 in float32, registration.py:551-553) for n (pair, t) items;
 * n_mask[i] = count(mask) from stage C.  rho_host[i] is NaN when fewer than two
 * samples or a constant input. */
int mvs_pc_spearman_batch(mvs_pc_plan* plan, int n, const int32_t* pairs, const double* ts,
                          const int64_t* n_mask, double* rho_host, void* stream);

/* ------------------------------------------------------------------------
 * Host <-> device movement of PAGEABLE host arrays: what the reference's hooks hand
 * over are plain numpy arrays (view slices, fusion/_core.py:1579-1587; the zarr
 * region a fused block is written to, :2130-2150).  `rows` rows of `width` bytes
 * are cut into 1 MiB pieces that rotate through a ring of pinned slots per direction: a small
 * pool of threads copies between the user's array and the slots (downloads with
 * cache-bypassing stores) while the calling thread alone enqueues the DMAs on `stream` and
 * waits for their events, so uploads, downloads and the host copies all overlap.
 * mvs_copy_h2d_2d returns once h_src has been staged (the device copy is ordered on
 * `stream`); mvs_copy_d2h_2d returns once h_dst is filled (work on `stream` enqueued
 * before the call is waited for).  Pitches in bytes.
 * ---------------------------------------------------------------------- */
int mvs_copy_h2d_2d(void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch,
                    size_t width, size_t rows, void* stream);
int mvs_copy_d2h_2d(void* h_dst, size_t h_pitch, const void* d_src, size_t d_pitch,
                    size_t width, size_t rows, void* stream);
/* n contiguous host arrays -> n device buffers as one pipelined transfer (the crops a
 * pairwise_reg_func batch is handed, registration.py:1926-1941). */
int mvs_copy_h2d_many(int n, void* const* d_dst, const void* const* h_src, const size_t* bytes,
                      void* stream);
/* The same for `planes` planes `*_plane` bytes apart (3-D windows of tiles and fused blocks). */
int mvs_copy_h2d_3d(void* d_dst, size_t d_pitch, size_t d_plane, const void* h_src, size_t h_pitch,
                    size_t h_plane, size_t width, size_t rows, size_t planes, void* stream);
int mvs_copy_d2h_3d(void* h_dst, size_t h_pitch, size_t h_plane, const void* d_src, size_t d_pitch,
                    size_t d_plane, size_t width, size_t rows, size_t planes, void* stream);

/* Chunk files of a Zarr directory store (output side of fuse(output_zarr_url=...),
 * fusion/_core.py:1160-1168, :2130-2150; input tiles likewise): n independent files written
 * from (write != 0) or read into host buffers by a pool of threads. */
int mvs_io_files(const char* const* paths, void* const* bufs, const size_t* sizes, int n, int write);

/* Zarr v2 chunk encode / decode (what zarr-python does under `da.to_zarr(region=...)`,
 * fusion/_core.py:2130-2150, and under `write_sim_to_ome_zarr`, ngff_utils.py:1353-1362; the
 * reference's own statement of the encoding is VirtualOMEZarr.read_chunk / _pad_edge_chunk,
 * ngff_utils.py:372-395, :425-436): a chunk is the C-order bytes of a chunk[0] x chunk[1] x
 * chunk[2] box, edge chunks padded with the fill value 0.  mvs_chunks_pack gathers a dense
 * level (element strides, x contiguous; 2-D uses z extent / chunk 1) into chunk-major order --
 * chunk (iz, iy, ix) at ((iz*gy + iy)*gx + ix) * prod(chunk) items, g = ceil(shape / chunk) --
 * and mvs_chunks_unpack scatters it back (bytes beyond the dense extent are dropped).
 * mvs_chunks_store writes n consecutive packed chunks of chunk_bytes each to paths[i]
 * (device -> ring of pinned chunk slots on `stream` -> files written by the copy pool while the
 * next DMAs run; returns when every file is written); mvs_chunks_load reads them back (a missing
 * file reads as zeros = the fill value; a short file is an error) and returns when all chunks
 * are resident. */
int mvs_chunks_pack(const void* d_dense, int item_size, const int32_t shape[3],
                    const int64_t stride[3], const int32_t chunk[3], void* d_packed, void* stream);
int mvs_chunks_unpack(const void* d_packed, int item_size, const int32_t shape[3],
                      const int64_t stride[3], const int32_t chunk[3], void* d_dense, void* stream);
int mvs_chunks_store(const void* d_packed, size_t chunk_bytes, int n, const char* const* paths,
                     void* stream);
int mvs_chunks_load(void* d_packed, size_t chunk_bytes, int n, const char* const* paths,
                    void* stream);

/* ------------------------------------------------------------------------
 * Pair preparation (register_pair_of_msims, registration.py:1732-1968): the
 * views are mean-binned (`sim.coarsen(binning, boundary="trim").mean()
 * .astype(dtype)`, :1732-1743), cropped to the overlap box plus one pixel
 * (:1765-1779; a strided window, no kernel) and resampled onto the fixed
 * view's pixel grid (sims_to_intrinsic_coord_system, :280-350 =
 * mvs_resample_views with cval NaN).
 * ---------------------------------------------------------------------- */

/* d_out (C-contiguous, shape[d] / bin[d] per axis, same dtype) = window means of
 * the strided volume d_in; integer means are float64 means truncated.  float32
 * windows skip NaNs when skip_nan != 0 (xarray's `.mean()`), else propagate them
 * (`np.mean(...).astype(dtype)`: the level-to-level step of the output pyramid,
 * ngff_utils.py:1284-1285, msi_utils.py:21-22, 49-60).  All triples are (z, y, x);
 * 2-D uses z extent / bin 1. */
int mvs_bin_mean(const void* d_in, int dtype, const int32_t shape[3],
                 const int64_t stride[3], const int32_t bin[3], int skip_nan, void* d_out,
                 void* stream);

/* ------------------------------------------------------------------------
 * Synthetic tiles (benchmark / test inputs).  Integer-only
 * value-noise ground truth sampled at integer global coordinates
 * origin + index, so overlapping tiles agree exactly.
 * ---------------------------------------------------------------------- */
int mvs_synth_tile(void* d_out, int dtype, const int32_t shape[3],
                   const int64_t stride[3], const int64_t origin[3], uint32_t seed,
                   void* stream);

/* Band-limited analytic ground truth for SUB-PIXEL tile positions: the tile is
 * f(origin + index) for f(p) = base + sum_k a_k sin(2 pi w_k.p + phi_k), clamped to
 * [0, 1) and scaled by out_scale (float32: 1, uint16: 4095), plus per-tile hash noise
 * of amplitude noise_amp.  The caller tabulates the per-axis factors in float64:
 * d_e{z,y,x}[k * n_axis + i] = exp(2 pi i w_k[axis] (origin[axis] + i)) as float2 and
 * d_coef[k] = a_k exp(i phi_k)  (device pointers; n_terms <= 128). */
int mvs_synth_field(void* d_out, int dtype, const int32_t shape[3], const int64_t stride[3],
                    const float* d_ez, const float* d_ey, const float* d_ex,
                    const float* d_coef, int n_terms, float base, float out_scale,
                    float noise_amp, uint32_t noise_seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVS_B200_H */
