/* sbb_textline.h -- C ABI of the B200-native tiled segmentation hot path.
 *
 * The reference (qurator-spk/sbb_textline_detection) has no FFI of its own: its hot path is the
 * Python method textline_detector.do_prediction (qurator/sbb_textline_detector/main.py:225-380)
 * calling keras Model.predict (main.py:287-288, 373-374) on a model returned by
 * start_new_session_and_model (main.py:216-223).  This header is the boundary cut beneath those
 * two methods; sbb_textline_detection_b200/detector.py binds it with ctypes.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 (SBB_OK) or a negative sbb_status;
 *     the message of the last failure on the calling thread is sbb_last_error().
 *   - images are uint8, HWC, 3 channels in BGR order (cv2.imread, main.py:197), row stride given
 *     in bytes; label maps are uint8 HW (class id per pixel; the reference replicates it to 3
 *     channels with np.repeat at main.py:292 -- callers that need that do it lazily on the host).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the model's own stream).  Calls are
 *     asynchronous with respect to the host only when every buffer is a device pointer
 *     (SBB_MEM_DEVICE); with host buffers the call returns after the results are in place.
 *   - a model handle is thread-compatible: one host thread inside a call per handle at a time (the Python
 *     binding holds a per-handle lock).  On the device the calls of one handle are ordered by the library
 *     itself, also when the caller alternates streams (they share the handle's workspace).  Several handles
 *     (one per GPU) may be driven from different threads.
 *   - every entry point leaves the calling thread's current CUDA context / device as it found it.
 */
#ifndef SBB_TEXTLINE_H_
#define SBB_TEXTLINE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBB_ABI_VERSION 1

typedef enum sbb_status {
  SBB_OK = 0,
  SBB_ERR_INVALID = -1,     /* bad argument / malformed weight blob            */
  SBB_ERR_CUDA = -2,        /* CUDA runtime/driver error (message has detail)  */
  SBB_ERR_UNSUPPORTED = -3, /* not an sm_100 device, tile size not /32, ...     */
  SBB_ERR_NOMEM = -4
} sbb_status;

/* Arithmetic of the conv stack (all modes accumulate in fp32 on the tensor cores).
 *   SBB_PREC_FP16X3: every operand is carried as an fp16 (hi, lo) pair and each product is
 *     hi*hi + hi*lo + lo*hi (3 tcgen05.mma per K step) -- fp32-grade results; this is the mode
 *     that meets the reference tolerance (1e-3 on logits, label IoU >= 0.999) and the default.
 *   SBB_PREC_FP16:   single fp16 operands, 1 MMA per K step; 3x less tensor work but NOT within
 *     the reference tolerance on the synthetic models; opt-in only.            */
typedef enum sbb_precision { SBB_PREC_FP16X3 = 0, SBB_PREC_FP16 = 1 } sbb_precision;

/* SBB_BACKEND_TCGEN05: TMA + tcgen05 implicit-GEMM kernels (the product path).
 * SBB_BACKEND_SIMT:    plain CUDA-core kernels over the same layer plan; slow, kept as an
 *                      on-device cross-check for tests.  Never selected implicitly.            */
typedef enum sbb_backend { SBB_BACKEND_TCGEN05 = 0, SBB_BACKEND_SIMT = 1 } sbb_backend;

typedef enum sbb_memkind { SBB_MEM_HOST = 0, SBB_MEM_DEVICE = 1 } sbb_memkind;

typedef struct sbb_model sbb_model; /* opaque; replaces the (keras model, tf session) pair of main.py:216-223 */

typedef struct sbb_model_desc {
  int32_t tile_h, tile_w;      /* model input size == model.layers[-1].output_shape[1:3] (main.py:227-228) */
  int32_t n_classes;           /* == output_shape[3] (main.py:229); must match the blob; <= 8              */
  int32_t precision;           /* sbb_precision                                                             */
  int32_t backend;             /* sbb_backend                                                               */
  int32_t device;              /* CUDA device ordinal                                                       */
  int32_t max_batch;           /* tiles processed per pass (workspace is sized for it); 0 = default (48)    */
  int32_t reserved;
  const void* weights;         /* host pointer to an SBBW0001 blob (sbb_textline_detection_b200/weights.py) */
  size_t weights_nbytes;
} sbb_model_desc;

int sbb_abi_version(void);
const char* sbb_last_error(void);

/* Replaces start_new_session_and_model (main.py:216-223): uploads the BN-folded weights, builds
 * the layer plan, TMA descriptors and workspace for `max_batch` tiles.                          */
int sbb_model_create(const sbb_model_desc* desc, sbb_model** out);
/* Replaces session.close(); del model; gc.collect(); K.clear_session() (main.py:428-436 etc.). */
void sbb_model_destroy(sbb_model* m);

/* model.layers[-1].output_shape[1:4] (main.py:227-229). */
int sbb_model_shape(const sbb_model* m, int32_t* tile_h, int32_t* tile_w, int32_t* n_classes);

/* do_prediction(patches=True, img, model) (main.py:231-366) in one call: /255, tile grid with
 * margin, per-tile forward, argmax, 9-case margin crop and last-writer-wins stitch.
 *   bgr:     uint8 [H][W][3], row stride `row_stride` bytes (>= 3*W)
 *   margin:  -1 => int(0.1 * tile_w) (main.py:233); otherwise the overlap margin in pixels
 *   labels:  uint8 [H][W], row stride `out_row_stride` bytes; every pixel is written (pixels no
 *            tile owns get 0, as in the reference's zero-initialised prediction_true)
 * Requires H >= tile_h and W >= tile_w (the reference wraps around with negative slices there). */
int sbb_predict_page_tiled(sbb_model* m, const uint8_t* bgr, int32_t H, int32_t W, int64_t row_stride,
                           int32_t margin, uint8_t* labels, int64_t out_row_stride,
                           int32_t memkind, void* stream);

/* Throughput form of sbb_predict_page_tiled for many pages of the SAME size: n_pages pages stacked vertically in one
 * buffer (page p = rows [p*H, (p+1)*H) of bgr_stack / labels_stack, common row strides) run through the network as
 * one batch, so the per-launch costs are paid once for all of them (when max_batch covers their tiles; otherwise
 * the batch is simply split).  Every page is tiled and stitched exactly as by sbb_predict_page_tiled. */
int sbb_predict_pages_stacked(sbb_model* m, const uint8_t* bgr_stack, int32_t n_pages, int32_t H, int32_t W,
                              int64_t row_stride, int32_t margin, uint8_t* labels_stack, int64_t out_row_stride,
                              int32_t memkind, void* stream);

/* model.predict(x) + np.argmax(axis=3) for a batch of tiles (main.py:287-290); the parity hook.
 *   tiles:  float32 [n][tile_h][tile_w][3] (already /255, BGR)
 *   labels: uint8 [n][tile_h][tile_w] or NULL
 *   probs:  float32 [n][tile_h][tile_w][n_classes] softmax output or NULL
 *   logits: float32 [n][tile_h][tile_w][n_classes] pre-softmax (after the final BN) or NULL   */
int sbb_predict_tiles(sbb_model* m, const float* tiles, int32_t n, uint8_t* labels, float* probs,
                      float* logits, int32_t memkind, void* stream);

/* do_prediction(patches=False, ...) core (main.py:373-377): one forward of a uint8 BGR image that
 * is already at tile size; the cv2.INTER_NEAREST resizes on both sides stay with the caller.   */
int sbb_predict_full(sbb_model* m, const uint8_t* bgr_tile, uint8_t* labels, int32_t memkind, void* stream);

/* Per-layer precision plan for SBB_PREC_FP16X3 handles.  The launches named in `layers` (comma separated layer
 * names as sbb_model_layer_time reports them; "" resets) read only the hi plane of their input activations: 2 MMA
 * units per K step instead of 3, no A_lo operand traffic; weights keep both planes.  Which layers tolerate this
 * within the 1e-3 logit tolerance is a property of the WEIGHTS (the randomly initialised synthetic models tolerate
 * none); sbb_textline_detection_b200/precision.py measures each layer's contribution and picks the plan. */
int sbb_model_set_precision_plan(sbb_model* m, const char* layers);

/* Page-geometry cache of a handle: tile / owner tables and the decoder work lists of the last few (H, W, margin)
 * geometries stay resident (every page of a run has its own border crop, main.py:2061 -> 2072); a hit costs
 * no upload and no synchronisation.  Slots: SBB_GEOM_CACHE (default 8).  Counters since sbb_model_create. */
int sbb_model_geom_cache_stats(const sbb_model* m, int64_t* hits, int64_t* misses);

/* Latency mode for ONE page on several GPUs: tiles [tile_first, tile_first + tile_count) of
 * the grid sbb_compute_tile_grid describes (reference loop order, main.py:259-260), device buffers only.
 * Every page pixel is owned by exactly one tile (main.py:294-364), so ranks that run disjoint tile ranges
 * with the SAME labels buffer -- one rank's memory, opened on the others with sbb_peer_open -- stitch the
 * page by the head epilogue's own stores over NVLink, without a collective.  keep_labels != 0: do not clear
 * the label map first (the owner clears it once before the ranks start). */
int sbb_predict_page_tile_range(sbb_model* m, const uint8_t* bgr_hwc, int32_t H, int32_t W, int64_t row_stride,
                                int32_t margin, uint8_t* labels_hw, int64_t out_row_stride,
                                int32_t tile_first, int32_t tile_count, int32_t keep_labels, void* stream);

/* Multi-GPU init: one process per GPU, pages shard page-per-GPU, the only exchange is ONE
 * broadcast of the frozen weights at init over NCCL (NVLink / NVSwitch).  NCCL is bound at run time (libnccl.so.2;
 * SBB_NCCL_LIB overrides the name), so linking this library does not require it.
 *   sbb_nccl_unique_id / sbb_nccl_comm_create / sbb_nccl_comm_destroy: ncclGetUniqueId / ncclCommInitRank /
 *     ncclCommDestroy for callers that do not bring their own communicator (the 128-byte id travels from rank 0
 *     to the others by whatever the launcher offers: a file, a TCP store, MPI).
 *   sbb_model_broadcast: every rank passes a host buffer of the same nbytes; on return each holds rank `root`'s
 *     blob and calls sbb_model_create on it.  `comm` is an ncclComm_t (void*), `stream` a cudaStream_t or NULL. */
int sbb_nccl_unique_id(uint8_t id[128]);
int sbb_nccl_comm_create(const uint8_t id[128], int32_t n_ranks, int32_t rank, int32_t device, void** comm);
int sbb_nccl_comm_destroy(void* comm);
int sbb_model_broadcast(void* blob, size_t nbytes, int32_t root, void* comm, int32_t device, void* stream);

/* Peer-visible device memory (CUDA IPC, one process per GPU): alloc returns the pointer and a 64-byte handle
 * to send to the other ranks; open maps another rank's buffer on `device` (peer access is enabled lazily). */
int sbb_peer_alloc(int32_t device, size_t nbytes, void** ptr, uint8_t handle[64]);
int sbb_peer_open(int32_t device, const uint8_t handle[64], void** ptr);
int sbb_peer_close(void* ptr);
int sbb_peer_free(void* ptr);

/* Host-only (no GPU touched): the tile grid and stitch ownership of do_prediction(patches=True)
 * (main.py:233-364).  tile_org receives {x0, y0, i, j} per tile in the reference's loop order (i outer,
 * j inner; capacity tile_cap tiles) and owner_x[W] / owner_y[H] the tile column / row whose write
 * survives at each page coordinate (-1: no tile writes there).  Any output pointer may be NULL. */
int sbb_compute_tile_grid(int32_t H, int32_t W, int32_t tile_h, int32_t tile_w, int32_t margin,
                          int32_t* nxf, int32_t* nyf, int32_t* tile_org, int32_t tile_cap,
                          int16_t* owner_x, int16_t* owner_y);

/* Host-only planner introspection (tests): the decoder launches of a page call skip the work outside the
 * region the stitch keeps (the margin the 9-case crop of main.py:294-364 discards).  For decoder block
 * `level` (1..5; 5 = tile resolution; its output grid is (tile_h >> (5-level)) x (tile_w >> (5-level)), one
 * launch over the half-resolution grid per output parity) of an H x W page this returns the M-tile shape
 * bw x bh the planner uses (full_grid_shapes != 0: the shape chosen for the whole grid) and the work items
 * {variant, page tile, X0, Y0}: variant = output parity 2*py + px (merged == 1: one item covers all four
 * parities of its low-res pixels, variant 0 -- the fused head; merged == 2: one item covers both COLUMN parities,
 * variant = row parity py -- dec4), (X0, Y0) the item's origin on the half-resolution grid.
 * items may be NULL (count only); at most item_cap items are written.                                */
int sbb_plan_decoder_tiles(int32_t H, int32_t W, int32_t tile_h, int32_t tile_w, int32_t margin, int32_t level,
                           int32_t merged, int32_t full_grid_shapes, int32_t* bw, int32_t* bh, int32_t* items,
                           int32_t item_cap, int32_t* count);

/* Introspection for tests (no GPU): the work list of a chained launch (SBB_CHAIN=1: a bottleneck's expand conv, variant 0
 * with nA N tiles, and the next block's reduce conv, variant 1 with nB, over the same flat 128-pixel M tiles of `px`
 * pixels) for a persistent grid of R CTA pairs.  items: {variant | n_tile << 8, first pixel} per work item, two
 * consecutive items (M tiles 2j, 2j+1) per pair position.  items may be NULL to query the count. */
int sbb_plan_chain_list(int64_t px, int32_t nA, int32_t nB, int32_t R, int32_t* items, int32_t item_cap, int32_t* count);

/* ---- byte-image operations around the models.  Model-independent; uint8 HWC
 * images with row strides in bytes; `stream` NULL = the legacy default stream; with SBB_MEM_HOST
 * buffers the call returns after the result is in place.  Bit-identical to OpenCV.            */

/* cv2.resize(src, (ow, oh), interpolation=cv2.INTER_NEAREST) -- resize_image, main.py:112-113
 * (get_image_and_scales :214, the no-patch path :371 and :378).  1 <= C <= 4.                 */
int sbb_resize_nearest_u8(const uint8_t* src, int32_t H, int32_t W, int32_t C, int64_t src_stride,
                          uint8_t* dst, int32_t oh, int32_t ow, int64_t dst_stride,
                          int32_t memkind, int32_t device, void* stream);

/* otsu_copy (main.py:178-194): cv2.threshold(THRESH_BINARY + THRESH_OTSU) of channel 0 of a C-channel
 * image, result (0 / 255) written to all THREE channels of dst [H][W][3] (the reference's quirk,
 * :191-193).  *threshold (optional, may be NULL) receives Otsu's threshold.                    */
int sbb_otsu_copy_u8(const uint8_t* src, int32_t H, int32_t W, int32_t C, int64_t src_stride,
                     uint8_t* dst, int64_t dst_stride, int32_t* threshold,
                     int32_t memkind, int32_t device, void* stream);

/* cv2.erode (op 0) / cv2.dilate (op 1) with the 5x5 ones kernel the reference uses everywhere
 * (self.kernel, main.py:57) and `iterations` (main.py:397: dilate x6; :2074-2075: erode x3, dilate x4),
 * per channel, OpenCV's default border (the border never wins).  src and dst must not overlap. */
int sbb_morph5x5_u8(const uint8_t* src, int32_t H, int32_t W, int32_t C, int64_t src_stride,
                    uint8_t* dst, int64_t dst_stride, int32_t op, int32_t iterations,
                    int32_t memkind, int32_t device, void* stream);

/* Row profiles of a binary mask under rotation: the inner loop of the deskew search
 * (return_deskew_slope main.py:1601-1718 -> rotate_image :159-163 + img_rotated[img_rotated!=0]=1 +
 * .sum(axis=1) in get_standard_deviation_of_summed_textline_patch_along_width :1545-1546).
 * The mask [h][w] (non-zero = set) sits at row oy, column ox of an S x S zero image; for each of the n
 * inverse affine maps (6 doubles, dst -> src, as cv2.warpAffine derives them from the 2x3 matrix)
 * profiles[k][y] = number of non-zero pixels in row y of
 * cv2.warpAffine(padded.astype(float64), M_k, (S, S), flags=INTER_CUBIC, borderMode=BORDER_REPLICATE).
 * Bit-identical to OpenCV.  inv_affine is always a host pointer; memkind applies to mask and profiles. */
int sbb_rotate_rowsum_u8(const uint8_t* mask, int32_t h, int32_t w, int64_t stride, int32_t S, int32_t oy,
                         int32_t ox, const double* inv_affine, int32_t n, int32_t* profiles,
                         int32_t memkind, int32_t device, void* stream);

/* Introspection used by tests and bench.py. */
int sbb_model_num_activations(const sbb_model* m);
/* name/shape of activation i (per tile): h, w, c.  Names follow the oracle's taps. */
int sbb_model_activation_info(const sbb_model* m, int32_t i, const char** name, int32_t* h, int32_t* w, int32_t* c);
/* Copy activation i of tile `tile` from the LAST forward to host float32 [h][w][c]. */
int sbb_model_read_activation(sbb_model* m, int32_t i, int32_t tile, float* out_hwc);
/* Kernel launches issued by the last predict call (claim for bench.py's gpu_launches). */
int64_t sbb_model_last_launch_count(const sbb_model* m);
/* Device timing of the following forwards.  enable == 1: a cudaEvent pair around every launch (each pair adds
 * a few microseconds of gap -- use the per-layer figures for SHARES); sbb_model_layer_time returns name/ms/flops of
 * entry i (i < num_layers) afterwards.  enable == 2: three events per forward only (start | first decoder launch |
 * end), no synchronisation: the encoder / decoder split of UNDISTURBED back-to-back forwards; sbb_model_part_times sums
 * the (at most 64) forwards recorded since the last read with reset != 0.  enable == 0: off. */
int sbb_model_set_profiling(sbb_model* m, int32_t enable);
int sbb_model_part_times(sbb_model* m, float* encoder_ms, float* decoder_ms, int32_t* forwards, int32_t reset);
int sbb_model_num_layers(const sbb_model* m);
int sbb_model_layer_time(const sbb_model* m, int32_t i, const char** name, float* ms, double* flops);

#ifdef __cplusplus
}
#endif
#endif /* SBB_TEXTLINE_H_ */
