/* log_b200_raster.h -- C ABI of the B200-native differentiable Gaussian-splatting rasteriser.
 *
 * This is the drop-in boundary for the hot path of zju3dv/LoG.  Plain pointers and sizes only; no torch types.
 * All pointers named *_d are DEVICE pointers (fp32 / int32, contiguous, 16-byte aligned); `stream` is a
 * cudaStream_t passed as void*.  Every function returns 0 on success, a positive cudaError_t value if the CUDA
 * runtime reported one, or a negative LGR_E_* code.  Nothing here ever falls back to the CPU.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   lgr_compute_radius     LoG/cuda/compute_radius_kernel.cu:107-183 (`compute_radius`, bound at :185-187) and the
 *                          fork's `rasterizer.compute_radius(xyz, scaling, rotation)` (LoG/model/level_of_gaussian.py:59)
 *   lgr_forward_project +
 *   lgr_forward_render     forward of `GaussianRasterizer.__call__` of diff_gaussian_rasterization[_wodilate]
 *                          (call sites LoG/render/renderer.py:153,190; LoG/model/level_of_gaussian.py:211)
 *   lgr_backward           its autograd backward (triggered at LoG/utils/trainer.py:158)
 * The Python binding a LoG maintainer uses is log_b200/rasterizer.py (ctypes); see INTEGRATION.md.
 */
#ifndef LOG_B200_RASTER_H
#define LOG_B200_RASTER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGR_ABI_VERSION 16
#define LGR_TILE 16

/* low-pass filter on the 2D covariance */
#define LGR_FILTER_ADD 0  /* stock 3DGS: cov_xx += 0.3, cov_yy += 0.3                                          */
#define LGR_FILTER_MAX 1  /* LoG fork ("wodilate"): cov_xx = max(cov_xx, 0.3) -- compute_radius_kernel.cu:100-103 */
#define LGR_FILTER_NONE 2 /* fork with use_filter=False (renderer.py:151-152)                                    */

#define LGR_E_BADARG (-1)
#define LGR_E_CAPACITY (-2) /* instance buffers smaller than the D the project stage reported */
#define LGR_E_UNSUPPORTED (-3)

/* Per-view constants.  Mirrors GaussianRasterizationSettings (kwargs at LoG/render/renderer.py:63-76).
 * viewmatrix/projmatrix/campos/bg stay on the device exactly as LoG hands them over (no host read-back). */
typedef struct lgr_view {
  int32_t image_height, image_width;
  float tanfovx, tanfovy;
  float scale_modifier;
  int32_t sh_degree;    /* active SH degree 0..3 (ignored when colors_precomp is given) */
  int32_t sh_coeffs;    /* K: coefficients per Gaussian stored in `shs` (N,K,3) */
  int32_t filter_mode;  /* LGR_FILTER_* */
  int32_t want_aux;     /* 1: also produce point_id_pixel / point_weight_pixel / point_weight (fork 5-tuple) */
  int32_t tile_row_begin, tile_row_end; /* this call renders tile rows [begin,end); 0,0 = all (multi-GPU shard) */
  /* Multi-GPU band mode (0 = off).  With num_owners = R > 0 the projection also compacts, without atomics, the ids of
   * the Gaussians that reach the rendered tile band: CTA b (256 consecutive ids) stores its ids at
   * band_ids_d[256 b ...) and their number in band_blk_d[b]; a scan then writes the exclusive prefix of those counts to
   * band_blk_d[B + b] (B = ceil(N/256) CTAs, B+1 prefix entries) and the per-owner totals to band_count_d[o], owner
   * o = id / owner_chunk with owner_chunk = LGR_OWNER_CHUNK(N, R) (a multiple of 256, so no CTA straddles owners).
   * Scatter and the per-Gaussian backward walk only those lists (ascending ids = grouped by owner), and splat records
   * of Gaussians outside the band are not written. */
  int32_t num_owners;
  int32_t raw_params;    /* 0: scales/opacities/rotations/colors_precomp are activated values (the reference call, default).
                            1: they are LoG's raw parameters and the activations of
                            LoG/model/activation.py:36-44 are fused into the projection and its backward:
                            scale = exp(raw), opacity = sigmoid(raw), rotation = raw / max(|raw|, 1e-12),
                            colour = C0 * raw + 0.5 (SH2RGB, sh_utils.py:72-73); gradients are w.r.t. the raw values.
                            Colour sources with raw_params: colors_precomp alone (DC colour), or colors_precomp (raw DC)
                            TOGETHER with shs = LoG's "rest" coefficients (N, sh_coeffs, 3), sh_coeffs >= (sh_degree+1)^2-1:
                            then LoG's colour activation (activation.py:27-34) is fused -- SH2RGB(dc) +
                            eval_sh_wobase(normalize(mean - campos), shs, sh_degree), NOT clamped at 0, direction
                            detached (no colour gradient into the mean).  Not available in band mode. */
  int32_t* band_ids_d;   /* (256 B) int32, or NULL */
  int32_t* band_blk_d;   /* (2 B + 1) int32 */
  int32_t* band_count_d; /* (num_owners) int32 */
  int32_t* band_rows_d;  /* (N) int32: dense packed-row -> id map, written by lgr_forward_render */
  float* band_dsplat_d;  /* (N,12) or NULL: the backward's dsplat_d; lgr_forward_render zeroes the rows of listed
                            Gaussians so that the caller need not zero-fill all N rows.  Also honoured with num_owners = 0:
                            the rows of all Gaussians with radius > 0 are zeroed (the only rows lgr_backward reads) */
  int32_t* tile_rank_d;  /* (rows,4) int32 or NULL.  When set, the counting pass (lgr_forward_project / lgr_shard_recv_bin)
                            takes the tile slots of every splat that covers <= 4 tiles with RETURNING atomics and stores
                            the 4 ranks here (row = Gaussian index); lgr_forward_render then places those instances at
                            tile_start + rank without touching an atomic again.  Splats covering more tiles are counted
                            in a second per-tile counter and still take their slots in lgr_forward_render.  NULL: every
                            slot is taken in lgr_forward_render (two atomic passes over the instances). */
  const int64_t* gather_index_d; /* (n) int64 or NULL (the gather of LoG/model/level_of_gaussian.py:262-296
                            fused): the n rows of the call are rows gather_index_d[0..n) of the input TABLES (means3D,
                            scales, rotations, opacities, colors_precomp, shs may hold any number of rows >= max index + 1);
                            every output -- splat, radii, point_weight, point_count, all gradients -- is compact, row i
                            belonging to table row gather_index_d[i], which is exactly the gradient layout LoG's
                            SparseOptimizer consumes (sparse_optimizer.py:163-196).  Not available in band mode. */
  const int32_t* pid_map_d; /* (n) int32 or NULL: when set, point_id_pixel holds pid_map_d[row] instead of the row index of the
                            winning splat (shard mode renders received ROWS; the map turns them into global Gaussian ids) */
  uint8_t* contrib_d;    /* (instances) uint8 or NULL, indexed like sorted_ids_d.  When set, the forward blend records per list
                            entry which of the tile's eight 8x4 sub-tiles had a CONTRIBUTING pixel for that splat (bit w =
                            sub-tile w), and the backward sweep walks exactly those (sub-tile, splat) pairs instead of
                            re-testing the conservative boxes: the ~20 % of box hits that contribute nothing are never
                            evaluated again.  Pass the same view (and buffer) to the forward and to the backward, and give
                            the backward the forward's n_contrib_d as last_contrib_d. */
  const int32_t* last_contrib_d; /* (H,W) int32 or NULL: the n_contrib_d output of lgr_forward_render (per pixel: list index + 1 of
                            its last contributor).  Read by lgr_backward / lgr_blend_backward together with contrib_d: a pixel
                            is finished once the sweep has passed its last contributor (both must be set, or neither). */
  const int32_t* region_count_d; /* (num_regions) int32 or NULL.  Shard mode: the rows of the call are num_regions regions of
                            region_cap rows each (rows = num_regions * region_cap) and only the first region_count_d[s] rows
                            of region s are in use (lgr_shard_layout: the counts live in the exchange buffer at off_count).
                            When set, lgr_shard_recv_bin[_aux] and lgr_forward_render[_device_sized] visit the used rows
                            only -- unused rows are neither read nor written (their radii are stale) -- so their cost
                            follows the rows a rank received, not the size of the exchange buffer. */
  int64_t region_cap;    /* rows per region (with region_count_d) */
  int32_t num_regions;   /* 1 .. LGR_SHARD_MAX_RANKS (with region_count_d) */
  int32_t reserved0;
  const float* cov3D_precomp_d; /* (N,6) or NULL: the stock API's cov3D_precomp (upper triangle xx xy xz yy yz zz of the world-space
                            covariance, diff_gaussian_rasterization's layout).  When set, lgr_forward_project / lgr_backward
                            take the covariance from here (scale_modifier is NOT applied, as in the stock rasteriser),
                            scales_d / rotations_d / dscales_d / drotations_d may be NULL, and lgr_backward writes
                            dL/dcov3D into dcov3D_d.  Not available with raw_params or in band mode. */
  float* dcov3D_d;       /* (N,6): gradient w.r.t. cov3D_precomp_d (off-diagonal entries receive the sum of the two symmetric
                            partials, like the stock backward); required by lgr_backward when cov3D_precomp_d is set */
  const float* viewmatrix_d; /* (4,4) world_view_transform, stored transposed (LoG/dataset/base.py:40-46) */
  const float* projmatrix_d; /* (4,4) full_proj_transform, same convention */
  const float* campos_d;     /* (3,) */
  const float* bg_d;         /* (3,) */
} lgr_view;

/* Sizes of the buffers the caller must provide. */
#define LGR_SPLAT_FLOATS 12 /* per-Gaussian projected record: 3 x float4 */
#define LGR_GRAD_FLOATS 12  /* per-Gaussian 2D-gradient accumulator: 3 x float4 */
#define LGR_TILE_SCRATCH_INTS 33 /* per tile: one counter per 128-byte line (32 ints) + one slot of the long-tile list */
#define LGR_META_INTS 8     /* meta_d: [0]=D binned instances [1]=longest tile list [2..3]=D by the stock
                               radius-square rule (lo,hi 32 bits) [4]=#Gaussians with radius>0
                               [5]=#tiles whose list exceeds the small shared-memory sort
                               [6]=overflow flags of lgr_forward_render_device_sized (0 = the outputs are valid) */

int lgr_abi_version(void);

/* radii_d[i] = 3*sqrt(lambda_max) of the projected Gaussian, 0 if culled.  Semantics of
 * compute_radius_cuda (compute_radius_kernel.cu:107-156): NDC cull +-1.3, no near cull, max(.,0.3) filter. */
int lgr_compute_radius(int64_t n, const float* means3D_d, const float* scales_d, const float* rotations_d,
                       const float* projmatrix_d, const float* viewmatrix_d, float focal_x, float focal_y,
                       float tan_fovx, float tan_fovy, float* radii_d, void* stream);

/* visible_d[i] = 1 when point i lies in front of the near plane (view-space z > 0.2), else 0: the stock module's
 * GaussianRasterizer.markVisible(positions) (diff_gaussian_rasterization; not called by LoG, part of the class LoG
 * instantiates at LoG/render/renderer.py:77).  viewmatrix_d as in lgr_view. */
int lgr_mark_visible(int64_t n, const float* means3D_d, const float* viewmatrix_d, uint8_t* visible_d, void* stream);

/* Stage 1 of the forward: per-Gaussian projection + EWA covariance + colour, tile counting, tile scan.
 *   in : means3D (N,3) opacities (N) scales (N,3) rotations (N,4); colors_precomp (N,3) XOR shs (N,K,3)
 *   out: splat_d (N,12) radii_d (N) int32; clamped_d (N) uint8 (SH only, may be NULL with colors_precomp);
 *        tile_start_d (tiles+1) int32 exclusive scan of per-tile counts (tiles = gx * rows rendered);
 *        tile_cursor_d (LGR_TILE_SCRATCH_INTS*tiles) int32 scratch; meta_d (LGR_META_INTS) int32.
 * The caller reads meta_d (one 32-byte D2H) to size the instance buffers for lgr_forward_render. */
int lgr_forward_project(const lgr_view* view, int64_t n, const float* means3D_d, const float* opacities_d,
                        const float* scales_d, const float* rotations_d, const float* colors_precomp_d,
                        const float* shs_d, float* splat_d, int32_t* radii_d, uint8_t* clamped_d,
                        int32_t* tile_start_d, int32_t* tile_cursor_d, int32_t* meta_d, void* stream);

/* Stage 2 of the forward: bin (Gaussian,tile) instances, per-tile (depth,index) radix sort, front-to-back blend.
 *   num_instances / max_tile_len / num_long_tiles : the values read from meta_d[0], meta_d[1], meta_d[5]
 *   scratch: inst_key_d, inst_val_d (num_instances) uint32; inst_tmp_d (2*num_instances) uint32, only needed when
 *            max_tile_len exceeds the shared-memory sort capacity (lgr_sort_smem_capacity()), else may be NULL
 *   out: sorted_ids_d (num_instances) int32 (kept for backward); image_d (3,H,W); final_T_d (H,W);
 *        n_contrib_d (H,W) int32; when view->want_aux: point_id_pixel_d (H,W) int32, point_weight_pixel_d (H,W),
 *        point_weight_d (N) -- must be zero-filled by the caller;
 *        point_count_d (N) int32 or NULL -- zero-filled by the caller; receives, per Gaussian, the number of pixels whose
 *        point_id_pixel is that Gaussian (the histogram LoG builds with torch.unique, renderer.py:156-159). */
int lgr_forward_render(const lgr_view* view, int64_t n, int64_t num_instances, int32_t max_tile_len,
                       int32_t num_long_tiles, const float* splat_d, const int32_t* radii_d, const int32_t* tile_start_d,
                       int32_t* tile_cursor_d, uint32_t* inst_key_d, uint32_t* inst_val_d, uint32_t* inst_tmp_d,
                       int32_t* sorted_ids_d, float* image_d, float* final_T_d, int32_t* n_contrib_d,
                       int32_t* point_id_pixel_d, float* point_weight_pixel_d, float* point_weight_d,
                       int32_t* point_count_d, void* stream);
int32_t lgr_sort_smem_capacity(void);

/* Stage 2 without the host read of meta_d ("device-sized"): the same work as lgr_forward_render, but every launch shape is
 * independent of D / the longest list / the number of long tiles -- the kernels read them from meta_d on the device -- so the
 * whole forward needs no host synchronisation and can be captured in a CUDA graph.  The caller sizes inst_key_d /
 * inst_val_d / sorted_ids_d for `instance_capacity` instances (e.g. 1.25 x the D of the previous view).  If the view needs
 * more (meta_d[0] > instance_capacity) or holds a tile list longer than lgr_sort_smem_capacity(), nothing is written out of
 * bounds, meta_d[6] is set to a non-zero value (bit 0: capacity, bit 1: list too long) and the outputs of this view are
 * INVALID: the caller checks meta_d[6] when it next synchronises and redoes the view through lgr_forward_render.
 * tile_start_d is mutable here (emptied on overflow). */
int lgr_forward_render_device_sized(const lgr_view* view, int64_t n, int64_t instance_capacity, int32_t* meta_d,
                                    const float* splat_d, const int32_t* radii_d, int32_t* tile_start_d,
                                    int32_t* tile_cursor_d, uint32_t* inst_key_d, uint32_t* inst_val_d,
                                    int32_t* sorted_ids_d, float* image_d, float* final_T_d, int32_t* n_contrib_d,
                                    int32_t* point_id_pixel_d, float* point_weight_pixel_d, float* point_weight_d,
                                    int32_t* point_count_d, void* stream);

/* Backward: per-tile gradient sweep (front to back, re-using the rendered image_d of the forward for the colour
 * behind each splat), then per-Gaussian projection backward.
 *   image_d (3,H,W): the forward's output, unmodified.  dsplat_d (N,12) scratch, zero-filled by the caller.
 *   out (each written for every Gaussian; culled ones get 0): dmeans3D (N,3) dmeans2D (N,3; d/d(ndc x,y), z = 0)
 *        dopacities (N) dscales (N,3) drotations (N,4) and dcolors (N,3) XOR dshs (N,K,3). */
int lgr_backward(const lgr_view* view, int64_t n, int64_t num_instances, const float* means3D_d,
                 const float* opacities_d, const float* scales_d, const float* rotations_d,
                 const float* colors_precomp_d, const float* shs_d, const float* splat_d, const int32_t* radii_d,
                 const uint8_t* clamped_d, const int32_t* tile_start_d, const int32_t* sorted_ids_d,
                 const float* image_d, const float* dL_dimage_d, float* dsplat_d,
                 float* dmeans3D_d, float* dmeans2D_d, float* dopacities_d, float* dscales_d, float* drotations_d,
                 float* dcolors_d, float* dshs_d, float* grad_rows_d, void* const* peer_stage_d, int32_t my_rank,
                 int64_t num_rows, void* stream);

/* Band mode only (view->num_owners > 0, precomputed colours): when grad_rows_d != NULL lgr_backward writes, instead of
 * the dense d*_d outputs (which may then be NULL), one packed row of LGR_ROW_FLOATS floats per listed Gaussian, rows
 * grouped by owner in list order:  [dmeans3D 0..2 | dmeans2D 3..5 | dopacity 6 | dscales 7..9 | drotations 10..13 |
 * dcolors 14..16 | id (int bits) 17 | radius 18 | 0].  lgr_grad_scatter_add adds received rows whose id lies in [lo,hi)
 * into a dense shard of (hi-lo) x LGR_ROW_FLOATS floats (row id-lo; slot 18 takes the maximum).  In band mode radii_d is
 * only valid for Gaussians that can reach the band (0 elsewhere): the owner's radius is shard[:, 18]. */
/* Fused exchange (band mode): when peer_stage_d != NULL it is a DEVICE array of num_owners pointers, entry o being owner
 * rank o's staging buffer mapped into this process (NVLink peer memory, e.g. torch symmetric memory).  lgr_backward then
 * stores every packed row straight into its owner's buffer instead of grad_rows_d:
 *     stage layout: LGR_STAGE_HEADER_FLOATS floats of header (int32 counts[source rank]) followed by
 *                   num_owners regions of owner_chunk rows; source rank s fills region s from its start.
 * After a cross-rank barrier the owner calls lgr_grad_scatter_add_staged on its own buffer. */
#define LGR_STAGE_HEADER_FLOATS 64
#define LGR_ROW_FLOATS 20
#define LGR_OWNER_CHUNK(n, r) ((((n) + (r) - 1) / (r) + 255) / 256 * 256)
int lgr_grad_scatter_add(int64_t num_rows, const float* rows_d, int64_t lo, int64_t hi, float* shard_d, void* stream);
int lgr_grad_scatter_add_staged(const float* stage_d, int32_t num_sources, int64_t owner_chunk, int64_t lo, int64_t hi,
                                float* shard_d, void* stream);

/* ---- Multi-GPU shard mode (BASELINE configs 4 and 5): Gaussians sharded over the ranks, tile-row bands owned
 * by ranks, projected splat records pushed to the band owners over NVLink peer memory, 2D gradients returned the same
 * way.  The reference has no multi-GPU path; log_b200/sharded.py:SplatExchange is the host side.
 *
 * Every rank allocates one exchange buffer of the same size in peer-mapped memory (e.g. torch symmetric memory);
 * peer_base_d is a DEVICE array of num_ranks pointers, entry r = rank r's buffer as mapped into this process.  All
 * offsets are in floats from the buffer start and must be multiples of 4 (16-byte alignment):
 *   off_count : int32[num_ranks]      rows received from each source rank
 *   off_splat : float[num_ranks*cap][12]  received splat records, region s (rows s*cap ..) from source rank s
 *   off_radii : int32[num_ranks*cap]   off_gid : int32[num_ranks*cap] (global Gaussian index of the row)
 *   off_dsplat: float[num_ranks*cap][12]  RETURNED 2D gradients, region o from band owner o, in pushed row order
 *   off_weight: uint32[num_ranks*cap] / off_pcount: int32[num_ranks*cap]  RETURNED point_weight bits / point counts
 * cap = LGR_OWNER_CHUNK(N, num_ranks) rows per (source, owner) pair; rank r owns Gaussians [r*cap, min(N,(r+1)*cap)).
 * Band owner o renders the tile rows tile_row_partition(H, num_ranks)[o] (the first gy % R bands have one more row). */
typedef struct lgr_shard_layout {
  int32_t num_ranks, my_rank;
  int64_t cap;
  int64_t off_count, off_splat, off_radii, off_gid, off_dsplat, off_weight, off_pcount;
} lgr_shard_layout;
#define LGR_SHARD_MAX_RANKS 32
/* int32 scratch of lgr_shard_send, kept until lgr_shard_gather of the same step: 2*R*B + R with B = ceil(max(n,1)/256) */
#define LGR_SHARD_SEND_INTS(n_local, r) (2 * (int64_t)(r) * ((((n_local) > 0 ? (n_local) : 1) + 255) / 256) + (r))

/* Source rank, after lgr_forward_project of its own n_local Gaussians with a FULL-IMAGE view (splat_d, radii_d): assign
 * rows without atomics and push records / radii / global ids (gid_base + i) into the band owners' buffers, and the row
 * counts into their headers.  A cross-rank barrier must follow before owners call lgr_shard_recv_bin. */
int lgr_shard_send(const lgr_view* view, const lgr_shard_layout* layout, int64_t n_local, int64_t gid_base,
                   const float* splat_d, const int32_t* radii_d, int32_t* send_scratch_d, void* const* peer_base_d,
                   void* stream);

/* Band owner: view->tile_row_begin/end = its band, and the view MUST carry the layout's region map (region_count_d =
 * (int32*)(exchange_d + off_count), region_cap = cap, num_regions = num_ranks; LGR_E_BADARG otherwise): this call and the
 * render that follows visit the used rows only, unused rows keep whatever an earlier step left in them.  Counts tiles of
 * the received rows, zeroes the dsplat_d rows (num_ranks*cap, 12) of used slots, then scans: tile_start_d / tile_cursor_d /
 * meta_d exactly as lgr_forward_project leaves them.  Continue with the SAME view: lgr_forward_render(view, n = num_ranks*cap, ..., splat_d =
 * exchange_d + off_splat, radii_d = exchange_d + off_radii, ...): point_id_pixel then holds ROW indices (map them
 * through the gid array), point_weight_d / point_count_d are per row. */
int lgr_shard_recv_bin(const lgr_view* view, const lgr_shard_layout* layout, float* exchange_d, float* dsplat_d,
                       int32_t* tile_start_d, int32_t* tile_cursor_d, int32_t* meta_d, void* stream);

/* The per-tile gradient sweep alone (first half of lgr_backward): accumulates into dsplat_d (n,12). */
int lgr_blend_backward(const lgr_view* view, int64_t n, int64_t num_instances, const float* splat_d,
                       const int32_t* tile_start_d, const int32_t* sorted_ids_d, const float* image_d,
                       const float* dL_dimage_d, float* dsplat_d, void* stream);

/* Band owner: send per-row data (rows_d: (num_ranks*cap, row_floats) fp32/int32, row_floats = 12 or 1) back to the
 * ranks that pushed the rows, into their buffers at dst_offset_floats (off_dsplat / off_weight / off_pcount).
 * total_rows: sum of the received counts (sizes the grid only).  A cross-rank barrier must follow. */
int lgr_shard_return_rows(const lgr_shard_layout* layout, const float* exchange_d, int64_t total_rows, const void* rows_d,
                          int32_t row_floats, int64_t dst_offset_floats, void* const* peer_base_d, void* stream);

/* Source rank: sum what the band owners returned into dense arrays of the local shard: dsplat_local_d (n_local,12) --
 * feed it to lgr_backward(view, n_local, num_instances = 0, ...) for the per-Gaussian backward --, and optionally
 * point_weight_d (n_local) fp32 (max over bands) / point_count_d (n_local) int32 (sum over bands). */
int lgr_shard_gather(const lgr_view* view, const lgr_shard_layout* layout, int64_t n_local, const float* splat_d,
                     const int32_t* radii_d, const int32_t* send_scratch_d, const float* exchange_d,
                     float* dsplat_local_d, float* point_weight_d, int32_t* point_count_d, void* stream);

/* The same three steps with fewer launches and no host-side sizes (what log_b200/sharded.py uses):
 *  - lgr_shard_recv_bin_aux also zeroes the used rows of the per-row aux accumulators the blend writes (point_weight_rows_d
 *    fp32, point_count_rows_d int32, each num_ranks*cap; either may be NULL), instead of two full-size memsets per step;
 *  - lgr_shard_return_packed sends everything back in ONE launch of fixed size: the 12-float gradient row with the aux
 *    values in its unused floats 9 (max alpha*T, fp32 bits) and 10 (winner-pixel count, int32 bits);
 *  - lgr_shard_gather_packed reads the aux values from there. */
int lgr_shard_recv_bin_aux(const lgr_view* view, const lgr_shard_layout* layout, float* exchange_d, float* dsplat_d,
                           int32_t* tile_start_d, int32_t* tile_cursor_d, int32_t* meta_d, float* point_weight_rows_d,
                           int32_t* point_count_rows_d, void* stream);
int lgr_shard_return_packed(const lgr_shard_layout* layout, const float* exchange_d, const float* dsplat_rows_d,
                            const float* point_weight_rows_d, const int32_t* point_count_rows_d, void* const* peer_base_d,
                            void* stream);
int lgr_shard_gather_packed(const lgr_view* view, const lgr_shard_layout* layout, int64_t n_local, const float* splat_d,
                            const int32_t* radii_d, const int32_t* send_scratch_d, const float* exchange_d,
                            float* dsplat_local_d, float* point_weight_d, int32_t* point_count_d, void* stream);

/* ---- Level-of-Gaussian tree traversal ----------------------------------------------------------
 * Replaces TensorTree.traverse / _query_tree_torch (LoG/model/tensor_tree.py:132-186) and the per-level
 * model.compute_radius calls inside it (LoG/model/level_of_gaussian.py:64-93: gather + exp / F.normalize activations
 * + compute_radius_cuda) by level-synchronous kernels with no host synchronisation between levels.
 *   node_index_d (num_points) int32: row of tree_d holding the children of a point, -1 = leaf
 *   tree_d (num_nodes, max_child) int32: child point ids, -1 = empty slot
 *   xyz_d (num_points,3); scaling_raw_d (num_points,3) RAW (scale = exp(raw), the 'exp' activation of activation.py:7);
 *   rotation_raw_d (num_points,4) RAW (normalised as F.normalize does); matrices / focal / tan_fov as lgr_compute_radius
 *   root_index_d (num_roots) int64: the visible roots, in the order the reference passes them
 *   max_depth: the reference's max_depth argument (child levels descended: min(max_level, max_depth))
 * Output: index_out_d (capacity num_points) int64 = exactly the reference's index_concat (same elements, same order:
 * kept roots, then the kept nodes of level 1, 2, ... in parent/child-slot order, then the nodes cut off at the depth
 * limit); *count_out_d their number.  scratch_d: LGR_TREE_SCRATCH_INTS(num_points, S) int32 with
 * S = max(num_roots, num_nodes * max_child). */
typedef struct lgr_tree {
  int64_t num_points, num_nodes;
  int32_t max_child, max_level;
  const int32_t* node_index_d;
  const int32_t* tree_d;
} lgr_tree;
#define LGR_TREE_SCRATCH_INTS(p, s) (8 + 2 * (int64_t)(p) + (int64_t)(s) + 2 * (((int64_t)(s) + 255) / 256) + 2 + ((int64_t)(s) + 3) / 4)
int lgr_tree_traverse(const lgr_tree* tree, const float* xyz_d, const float* scaling_raw_d, const float* rotation_raw_d,
                      const float* projmatrix_d, const float* viewmatrix_d, float focal_x, float focal_y, float tan_fovx,
                      float tan_fovy, const int64_t* root_index_d, int64_t num_roots, float min_resolution_pixel,
                      int32_t max_depth, int32_t* scratch_d, int64_t* index_out_d, int64_t* count_out_d, void* stream);

/* Sorted compaction of the non-zero entries of point_count_d: ids_out_d / counts_out_d (capacity min(N, H*W)) receive the
 * ids in ascending order and their pixel counts, *num_out_d their number.  Equals
 * torch.unique(point_id_pixel, sorted=True, return_counts=True) with the -1 entry dropped (renderer.py:156-159).
 * scratch_d: 2*ceil(N/1024)+1 int32. */
int lgr_point_compact(int64_t n, const int32_t* point_count_d, int32_t* scratch_d, int32_t* ids_out_d,
                      int32_t* counts_out_d, int32_t* num_out_d, void* stream);

/* Fused sparse Adam step: replaces SparseOptimizer.step's gather / _single_tensor_adam / scatter
 * (LoG/model/sparse_optimizer.py:41-78, 163-196) for one parameter.  For every k < rows and c < row_floats, with
 * i = index_d[k]:   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  [vmax = max(vmax, v)] ;
 *                   param -= lr / (1 - b1^step) * m / (sqrt(v or vmax) / sqrt(1 - b2^step) + eps)
 * in place on param_d / exp_avg_d / exp_avg_sq_d [/ max_exp_avg_sq_d, NULL = no amsgrad] (N, row_floats);
 * grad_d is the compact (rows, row_floats) gradient of the gathered rows; index_d int64 (rows), unique. */
int lgr_sparse_adam(int64_t rows, int32_t row_floats, const int64_t* index_d, const float* grad_d, float* param_d,
                    float* exp_avg_d, float* exp_avg_sq_d, float* max_exp_avg_sq_d, int64_t step, double lr, double beta1,
                    double beta2, double eps, void* stream);

/* Diagnostics (not on the data path): per-kernel CUDA-event timing on the launching stream.
 * lgr_profile_enable(1) starts recording; lgr_profile_collect() synchronises the recorded events, writes the summed
 * milliseconds and the launch counts per kernel id (LGR_PROFILE_KERNELS entries) and resets the counters. */
#define LGR_PROFILE_KERNELS 12
int lgr_profile_enable(int on);
int lgr_profile_collect(double* ms_out, int32_t* launches_out, int32_t capacity);
const char* lgr_profile_kernel_name(int kernel_id);

/* Multi-GPU helper (tile-sharded ranks): out[i] += in[i] for the per-Gaussian gradient exchange is done with NCCL
 * by the host side (log_b200/sharded.py); no entry point is needed here for it. */

#ifdef __cplusplus
}
#endif
#endif /* LOG_B200_RASTER_H */
