/*
 * mccnn_b200.h -- C ABI of libmccnn_b200.so: the B200 (sm_100a) stereo-matching hot path of
 * Jackie-Chou/MC-CNN-python, behind plain pointers and sizes.
 *
 * The reference has no FFI layer: its boundary is the module-level Python functions of
 * src/process_functional.py ("pf") that src/match.py:132-175 calls.  Every entry point below
 * replaces the BODY of one of those functions; the Python mirror in
 * mc-cnn-python_b200/process_functional.py keeps the reference names and argument order and
 * binds these symbols with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; float32 everywhere;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing,
 *     keeps no global state except a thread-local error string, and returns 0 on success or a
 *     negative mccnn_status; no C++ exception crosses the ABI;
 *   - images      [H][W]          (the reference's [H,W,1] squeezed)
 *   - features    [H][W][C]       (C = 64)
 *   - disparity   [H][W]          float32 (the reference stores WTA indices as float32, pf:243)
 *   - cost volume "HWD": [H][W][Dp], disparity fastest, Dp = mccnn_dpitch(D) = D rounded up to a
 *     multiple of 4; the pad cells are never read as data.  The reference's logical [D,H,W]
 *     array is the permuted view; mccnn_dhw_to_hwd / mccnn_hwd_to_dhw convert to and from the
 *     reference's physical [D][H][W] order.
 */
#ifndef MCCNN_B200_H
#define MCCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum mccnn_status {
    MCCNN_OK = 0,
    MCCNN_ERR_ARG = -1,      /* invalid argument (message in mccnn_last_error) */
    MCCNN_ERR_CUDA = -2,     /* a CUDA call or launch failed */
    MCCNN_ERR_UNSUPPORTED = -3
};

/* Thread-local message of the last failing call ("" if none). */
const char *mccnn_last_error(void);
/* ABI version (bumped on any signature change). */
int mccnn_abi_version(void);
/* Physical disparity pitch of the HWD layout. */
int mccnn_dpitch(int D);
/* Number of kernels launched by this library in this process so far (for bench.py's gpu_launches). */
unsigned long long mccnn_launch_count(void);

/* [D][H][W] (reference order)  <->  [H][W][Dp] (this library's order). */
int mccnn_dhw_to_hwd(const float *dhw, float *hwd, int D, int H, int W, void *stream);
int mccnn_hwd_to_dhw(const float *hwd, float *dhw, int D, int H, int W, void *stream);

/* ---- a1/a2  model.NET forward + pf:15 compute_features (model.py:40-64, :90-125; pf:20-25) ----
 * img [H][W] normalised image with `pad` implicit zero pixels per side (compute_features pads by
 * (patch-1)/2 = num_layers, pf:20-25; NET itself applies VALID convolutions, pad = 0).
 * weights_host / biases_host: HOST arrays of num_layers DEVICE pointers; weights HWIO
 * ([3][3][1][64], then [3][3][64][64]), biases [64].
 * out [H+2pad-2n][W+2pad-2n][64], L2-normalised over channels (model.py:64).
 * scratch: mccnn_features_scratch_bytes(H, W, pad, num_layers) bytes. */
size_t mccnn_features_scratch_bytes(int H, int W, int pad, int num_layers);
int mccnn_features(const float *img, int H, int W, int pad, int num_layers,
                   const float *const *weights_host, const float *const *biases_host,
                   float *out, void *scratch, void *stream);
/* The same with the network's constant part done once (the reference builds its graph once per process and runs it
 * per image, pf:30-47): mccnn_features_prepare writes the hi/lo tf32 split of the weights of layers 2..n into
 * `prepared` (mccnn_features_weights_bytes(num_layers) bytes, 32-byte aligned); mccnn_features_prepared uses it
 * instead of splitting the weights on every call.  Same results bit for bit.
 * Layers 2..n run on the tensor cores with FP16 hi/lo operands (three products, float32 accumulation: float32 accuracy);
 * a layer whose input or weights exceed fp16's range runs with TF32 hi/lo operands instead (device-side decision). */
size_t mccnn_features_weights_bytes(int num_layers);
int mccnn_features_prepare(int num_layers, const float *const *weights_host, void *prepared, void *stream);
int mccnn_features_prepared(const float *img, int H, int W, int pad, int num_layers,
                            const float *const *weights_host, const float *const *biases_host,
                            const void *prepared, float *out, void *scratch, void *stream);

/* ---- a3  pf:78 compute_cost_volume ----
 * fl, fr [H][W][C]; L, R: HWD volumes.  Requires C == 64, 1 <= D <= 512 (a limit the reference does not have: the
 * band of one tile and the SGM register tiling are sized for it; more disparities go through the slab entry points
 * below), W >= D + 2. */
int mccnn_cost_volume(const float *fl, const float *fr, float *L, float *R,
                      int H, int W, int C, int D, void *stream);

/* ---- a4  pf:571 compute_cross_region, arm-length form ----
 * arms [H][W][4] u8 = {up, down, left, right}; count [H][W] i32 = |U(h,w)|. */
int mccnn_cross_arms(const float *img, uint8_t *arms, int32_t *count, int H, int W,
                     float intensity_threshold, int distance_threshold, void *stream);
/* sum[0] (device, 8 bytes) = sum over the image of up + down: the host-side choice between the two bit-identical
 * separable aggregation schedules reads this one number back (process_functional.cbca_auto_mode). */
int mccnn_arms_vertical_sum(const uint8_t *arms, int H, int W, unsigned long long *sum, void *stream);
/* The reference's explicit list: region [H][W][(2*dist)^2][2] i32 padded with -1 (pf:638-655). */
int mccnn_cross_region_list(const uint8_t *arms, int32_t *region, int H, int W,
                            int distance_threshold, void *stream);

/* ---- a5  pf:117 cost_volume_aggregation, one volume ----
 * `iters` rounds of region mean; in is left untouched, out receives the result, scratch is one
 * more HWD volume.
 * mode MCCNN_CBCA_SEPARABLE (default): row sums re-used down each column (<= 54 additions per cell at
 *      distance_threshold 14); equals the reference up to float32 re-association of the sum (~1e-7 relative).
 *      The rounds of a call are chained: one row pass, then per further round ONE kernel (k_cbca_colrow_g) that forms a
 *      round's column sums + division in shared memory and the next round's row sums from them, then one column pass:
 *      8 instead of 16 B per cell per round through HBM.  `scratch` is needed for any iters >= 1.
 * mode MCCNN_CBCA_SEPARABLE_TWO_PASS: the same sums in the same order as two streaming passes per round (rows into
 *      `scratch`, columns into `out`).  Bit-identical to MCCNN_CBCA_SEPARABLE; faster than it on piece-wise constant
 *      images whose vertical arms all sit at the distance limit, slower on natural ones; also the cross-check of the
 *      chained kernel.
 * mode MCCNN_CBCA_EXACT: one float32 running sum over the whole region in the reference's enumeration
 *      order (pf:149-163), bit-identical to the reference (<= 729 additions per cell); `scratch` for iters >= 2.
 * distance_threshold (1..255) must be the value the arms were built with (it sizes the chained kernel's tile). */
enum mccnn_cbca_mode { MCCNN_CBCA_SEPARABLE = 0, MCCNN_CBCA_EXACT = 1, MCCNN_CBCA_SEPARABLE_TWO_PASS = 2 };
int mccnn_cbca(const float *in, float *out, float *scratch, const uint8_t *arms,
               const int32_t *count, int D, int H, int W, int iters, int distance_threshold, int mode,
               void *stream);

/* ---- a6  pf:476 semi_global_matching, one in-place pass over one volume ----
 * (rh, rw) in {(0,1),(0,-1),(-1,0),(1,0)}.  P1/P2/Q1/Q2/tauD arrive as doubles and are rounded to
 * float32 exactly where the reference rounds them (pf:504-505, :538-541).
 * flags_scratch: mccnn_sgm_scratch_bytes(H, W, D) bytes.  D <= 512 (MCCNN_ERR_UNSUPPORTED above). */
size_t mccnn_sgm_scratch_bytes(int H, int W, int D);
int mccnn_sgm_pass(float *vol, const float *img_left, const float *img_right, void *flags_scratch,
                   int D, int H, int W, int rh, int rw,
                   double sgm_P1, double sgm_P2, double sgm_Q1, double sgm_Q2, double sgm_D,
                   int is_left, void *stream);
/* ---- a7  pf:187 SGM_average for one volume: the four chained in-place passes (pf:194-210). */
int mccnn_sgm_average(float *vol, const float *img_left, const float *img_right, void *flags_scratch,
                      int D, int H, int W,
                      double sgm_P1, double sgm_P2, double sgm_Q1, double sgm_Q2, double sgm_D,
                      double sgm_V, int is_left, void *stream);

/* Both volumes of a pair in shared launches (the two volumes are independent until WTA); either
 * volume pointer may be NULL. */
int mccnn_sgm_average_pair(float *vol_left, float *vol_right, const float *img_left, const float *img_right,
                           void *flags_scratch, int D, int H, int W,
                           double sgm_P1, double sgm_P2, double sgm_Q1, double sgm_Q2, double sgm_D,
                           double sgm_V, void *stream);

/* ---- a8  pf:239 disparity_prediction, one volume: first minimum over d, stored as float32. */
int mccnn_wta(const float *vol, float *disp, int D, int H, int W, void *stream);

/* ---- a5 + a8 fused: match.py:155-160 runs disparity_prediction (pf:239) on the volume cost_volume_aggregation (pf:117)
 * has just produced.  Same as mccnn_cbca (separable modes only) followed by mccnn_wta on `out`, but the closing column
 * pass of the aggregation takes the first minimum itself, so the volume is not read again: disp [H][W] f32 is identical
 * to mccnn_wta's (first minimum; -1 where no finite cost exists).  keys: scratch of H*W 8-byte words.  store_volume = 0
 * additionally skips writing the aggregated volume (`out` then holds intermediate row sums; match.py never reads the
 * right volume after its WTA, match.py:160-166); `out` and `scratch` are needed as work space either way. */
int mccnn_cbca_wta(const float *in, float *out, float *scratch, const uint8_t *arms, const int32_t *count,
                   int D, int H, int W, int iters, int distance_threshold, int mode, int store_volume,
                   void *keys, float *disp, void *stream);

/* ---- a9  pf:279 interpolation.  labels [H][W] i32 scratch/output (0 match, 1 mismatch, 2 occlusion). */
int mccnn_lr_interp(const float *disp_left, const float *disp_right, float *out, int32_t *labels,
                    int H, int W, int ndisp, void *stream);

/* ---- a10 pf:381 subpixel_enhance (vol is the HWD left volume). */
int mccnn_subpixel(const float *disp, const float *vol, float *out, int D, int H, int W, void *stream);

/* ---- a11 pf:403 median_filter: border-clipped fh x fw window, np.median semantics; fh*fw <= 121. */
int mccnn_median(const float *in, float *out, int H, int W, int fh, int fw, void *stream);

/* ---- a12 pf:424 bilateral_filter.  table [fh][fw] = float32 weights of pf:433-436; fh*fw <= 121. */
int mccnn_bilateral(const float *img, const float *in, float *out, const float *table,
                    int H, int W, int fh, int fw, float blur_threshold, void *stream);

/* ---- One big pair partitioned over GPUs by disparity slab (the reference has no such mode: its
 * only parallelism is disjoint pair windows, match.py:26-28).  A slab [d_base, d_base + d_count) of a volume is its
 * own HWD volume of d_count disparities; d_base is a multiple of 4.  Cost volume, CBCA and WTA are independent
 * per disparity plane and run on slabs; SGM needs every disparity of a pixel and runs on row slabs (horizontal
 * passes) and column slabs (vertical passes) after a re-partition (mccnn_copy3d + all-to-all). */

/* a3 restricted to disparities [d_base, d_base + d_count) of an ndisp = D problem (pf:78-113). */
int mccnn_cost_volume_slab(const float *fl, const float *fr, float *L, float *R,
                           int H, int W, int C, int D, int d_base, int d_count, void *stream);

/* a5 on a disparity slab whose result is re-partitioned into row slabs right after: `iters` rounds of the default
 * mode, the closing column pass storing row h into dst[r] for row_bounds[r] <= h <
 * row_bounds[r+1] -- a row slab [rows_r][W][4 * g_total] of the owner (peer memory), at granule offset g_offset.
 * `out` holds the intermediate rounds.  row_bounds (nparts + 1) and dst (nparts device pointers) are HOST arrays. */
int mccnn_cbca_to(const float *in, float *out, float *scratch, const uint8_t *arms, const int32_t *count,
                  int D, int H, int W, int iters, int distance_threshold, int nparts, const int *row_bounds,
                  float *const *dst, int g_offset, int g_total, void *stream);

/* a6/a7: two of the four chained passes (pf:194-208).  which = 0: (0,1) then (0,-1) on a ROW slab -- volumes
 * [H][W][Dp] and images hold the slab's H rows, w_base = 0, w_count = W.  which = 1: (-1,0) then (1,0) on a COLUMN
 * slab -- volumes [H][w_count][Dp] hold image columns [w_base, w_base + w_count), images are whole [H][W].
 * flags_scratch: mccnn_sgm_scratch_bytes(H, W, D). */
int mccnn_sgm_passes_slab(float *vol_left, float *vol_right, const float *img_left, const float *img_right,
                          void *flags_scratch, int D, int H, int W, int w_base, int w_count, int which,
                          double sgm_P1, double sgm_P2, double sgm_Q1, double sgm_Q2, double sgm_D,
                          double sgm_V, void *stream);

/* The same with the pair's second pass storing every cell straight into the buffer of the rank that needs it in
 * the next layout (peer memory over NVLink) instead of in place: which = 0 sends columns [bounds[r], bounds[r+1])
 * to dst_*[r], a column slab [H_total][bounds[r+1]-bounds[r]][Dp] whose row h_base + h receives this slab's row h;
 * which = 1 sends granules [bounds[r], bounds[r+1]) to dst_*[r], a disparity slab [H][W][4*(bounds[r+1]-bounds[r])].
 * bounds (nparts + 1 ints) and the dst tables (nparts device pointers each) are HOST arrays; nparts <= 8. */
int mccnn_sgm_passes_slab_to(float *vol_left, float *vol_right, const float *img_left, const float *img_right,
                             void *flags_scratch, int D, int H, int W, int w_base, int w_count, int which,
                             double sgm_P1, double sgm_P2, double sgm_Q1, double sgm_Q2, double sgm_D,
                             double sgm_V, int nparts, const int *bounds, float *const *dst_left,
                             float *const *dst_right, int h_base, void *stream);

/* a8 on a slab of D disparities starting at d_base: disp = the pair's disparity of the slab's first minimum,
 * minval = its cost.  mccnn_wta_combine takes the all-gathered arrays (slab s at element s * slab_stride) and keeps, per pixel, the
 * first strict minimum in slab order (the lowest disparity wins ties, pf:247-252). */
int mccnn_wta_slab(const float *vol, float *disp, float *minval, int D, int H, int W, int d_base, void *stream);
int mccnn_wta_combine(const float *minvals, const float *disps, float *out, int nslabs, long long slab_stride,
                      int H, int W, void *stream);

/* a10 on a slab: triple [3][H][W] = (C[d-1], C[d], C[d+1]) of pf:392-394 where this slab owns the cell, else 0
 * (a sum over the slabs restores the cells exactly); mccnn_subpixel_triple applies pf:395-398 to the summed triple. */
int mccnn_subpixel_gather(const float *disp, const float *vol, float *triple, int D, int H, int W, int d_base,
                          int ndisp, void *stream);
int mccnn_subpixel_triple(const float *disp, const float *triple, float *out, int ndisp, int H, int W, void *stream);

/* Strided block copy [n0][n1][n2_granules x 16 bytes]; strides in 16-byte granules.  Packs / unpacks the blocks a
 * re-partition exchanges (rows <-> columns <-> disparity slabs of an HWD volume). */
int mccnn_copy3d(const float *src, float *dst, long long n0, long long n1, int n2_granules,
                 long long src_stride0, long long src_stride1, long long dst_stride0, long long dst_stride1,
                 void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MCCNN_B200_H */
