// Host side of the C ABI (include/sbb_textline.h): weight-blob parsing, the layer planner that
// turns the ResNet50-U-Net into implicit-GEMM launches, TMA descriptor
// encoding, the per-batch forward and the page tiler/stitcher of do_prediction (main.py:225-380).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sbb_textline.h"
#include "conv_gemm_tc.cuh"
#include "conv_gemm_pair.cuh"
#include "kernels_aux.cuh"
#include "plan.h"

using namespace sbb;

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return fail(SBB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------ blob
struct Rec {
  std::string name;
  int kh, kw, cin, cout;
  const float* w;  // OHWI
  const float* b;
};

static int parse_blob(const void* data, size_t n, int* n_classes, int* tile_h, int* tile_w, std::vector<Rec>* recs) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  if (n < 24 || memcmp(p, "SBBW0001", 8) != 0) return fail(SBB_ERR_INVALID, "weight blob: bad magic");
  uint32_t hdr[4];
  memcpy(hdr, p + 8, 16);
  *n_classes = (int)hdr[0];
  *tile_h = (int)hdr[2];  // 0: the blob does not record the input size it was converted for
  *tile_w = (int)hdr[3];
  size_t off = 24;
  for (uint32_t i = 0; i < hdr[1]; ++i) {
    if (off + 56 > n) return fail(SBB_ERR_INVALID, "weight blob: truncated header of record %u", i);
    Rec r;
    char nm[33];
    memcpy(nm, p + off, 32);
    nm[32] = 0;
    r.name = nm;
    uint32_t d[4];
    uint64_t nw;
    memcpy(d, p + off + 32, 16);
    memcpy(&nw, p + off + 48, 8);
    off += 56;
    r.kh = d[0]; r.kw = d[1]; r.cin = d[2]; r.cout = d[3];
    if (off + 4 * nw + 4 * (size_t)r.cout > n) return fail(SBB_ERR_INVALID, "weight blob: truncated record %s", nm);
    if (r.kh > 0 && nw != (uint64_t)r.kh * r.kw * r.cin * r.cout)
      return fail(SBB_ERR_INVALID, "weight blob: record %s has %llu weights", nm, (unsigned long long)nw);
    r.w = reinterpret_cast<const float*>(p + off);
    off += 4 * nw;
    r.b = reinterpret_cast<const float*>(p + off);
    off += 4 * (size_t)r.cout;
    recs->push_back(r);
  }
  if (off != n) return fail(SBB_ERR_INVALID, "weight blob: %zu trailing bytes", n - off);
  return SBB_OK;
}

// ------------------------------------------------------------------------------------------ tiler
// Tile grid + separable owner tables of do_prediction(patches=True) (main.py:233-364).
//   keep-x of tile column i: [0, mw-m) if i==0; else [m, mw) if i==nxf-1; else [m, mw-m)   (the
//   reference's if/elif chain tests i==0 first), same for rows.  Writes overwrite with i outer,
//   j inner, so pixel (y, x) ends up owned by the LARGEST i whose keep-x contains x and the
//   largest j whose keep-y contains y.  Unowned positions get -1 (label stays 0).
extern "C" int sbb_compute_tile_grid(int32_t H, int32_t W, int32_t tile_h, int32_t tile_w, int32_t margin,
                                     int32_t* nxf_out, int32_t* nyf_out, int32_t* tile_org, int32_t tile_cap,
                                     int16_t* owner_x, int16_t* owner_y) {
  if (margin < 0) margin = (int)(0.1 * tile_w);
  const int wm = tile_w - 2 * margin, hm = tile_h - 2 * margin;
  if (H < tile_h || W < tile_w)
    return fail(SBB_ERR_INVALID, "page %dx%d smaller than the %dx%d tile (the reference wraps around here)", H, W,
                tile_h, tile_w);
  if (wm <= 0 || hm <= 0) return fail(SBB_ERR_INVALID, "margin %d leaves no tile interior", margin);
  const int nxf = (W + wm - 1) / wm, nyf = (H + hm - 1) / hm;
  if (nxf > 32000 || nyf > 32000) return fail(SBB_ERR_INVALID, "too many tiles");
  *nxf_out = nxf;
  *nyf_out = nyf;
  std::vector<int> x0(nxf), y0(nyf);
  for (int i = 0; i < nxf; ++i) { x0[i] = i * wm; if (x0[i] + tile_w > W) x0[i] = W - tile_w; }
  for (int j = 0; j < nyf; ++j) { y0[j] = j * hm; if (y0[j] + tile_h > H) y0[j] = H - tile_h; }
  if (tile_org) {
    if (tile_cap < nxf * nyf) return fail(SBB_ERR_INVALID, "tile_org capacity %d < %d tiles", tile_cap, nxf * nyf);
    int t = 0;
    for (int i = 0; i < nxf; ++i)
      for (int j = 0; j < nyf; ++j, ++t) {
        tile_org[4 * t + 0] = x0[i]; tile_org[4 * t + 1] = y0[j];
        tile_org[4 * t + 2] = i; tile_org[4 * t + 3] = j;
      }
  }
  if (owner_x) {
    for (int x = 0; x < W; ++x) owner_x[x] = -1;
    for (int i = 0; i < nxf; ++i) {
      const int a = (i == 0) ? 0 : margin;
      const int b = (i == 0) ? tile_w - margin : (i == nxf - 1 ? tile_w : tile_w - margin);
      for (int x = x0[i] + a; x < x0[i] + b; ++x) owner_x[x] = (int16_t)i;
    }
  }
  if (owner_y) {
    for (int y = 0; y < H; ++y) owner_y[y] = -1;
    for (int j = 0; j < nyf; ++j) {
      const int a = (j == 0) ? 0 : margin;
      const int b = (j == 0) ? tile_h - margin : (j == nyf - 1 ? tile_h : tile_h - margin);
      for (int y = y0[j] + a; y < y0[j] + b; ++y) owner_y[y] = (int16_t)j;
    }
  }
  return SBB_OK;
}

// ------------------------------------------------------------------------------------------ model
struct Tensor {
  __half* d = nullptr;
  int H = 0, W = 0, C = 0, planes = 1;
  int64_t pix() const { return (int64_t)planes * C; }
  int lo_off() const { return planes == 2 ? C : 0; }
};

struct WSrc {  // where a segment's weights come from
  int rec, ky, kx, cin0;
  bool packed_row;  // packed input-image segment (kSegPacked): the chunk is 8 consecutive input pixels x
                    // {c0h c1h c2h 1 | c0l c1l c2l 0}; pixel j carries tap (ky, j) for j < kw; cin0 = first
                    // of the 3 input channels inside the record; `kx` = pixel whose constant-1 slot
                    // carries the bias (-1: none)
  uint32_t tapmask = 0;  // != 0: SUM of the taps with bit (ky*kw+kx) set (merged upsample taps)
  bool identity = false; // residual segment: weight block W[o][c] = (o % BN == c)
  // merged-parity head (dec5): output column o = parity * (Cout/4) + channel, parity = 2*py + px.
  //   up-sampled segment: tapmask4[parity] = the taps that read this segment's low-res pixel (0: none)
  //   packed input row  : `ky` = row r of the 4x4 input window of a low-res pixel, pixel j of the chunk is
  //                       its column; parity (py, px) has tap (r - py, j - px) there; `kx` = bias pixel
  bool merged = false;
  uint32_t tapmask4[4] = {0, 0, 0, 0};
  // merged COLUMN parities of a decoder block with few output channels (dec4): output column
  // o = px * (Cout/2) + channel for the launch's fixed row parity; tapmask4[px] as above (n_par = 2)
  int n_par = 4;
};

enum OpKind { OP_STEM_PAD, OP_POOL, OP_CONV };

struct Rect { int x0, y0, x1, y1; };  // inclusive pixel bounds; empty when x1 < x0 or y1 < y0

struct WorkList {  // device work list of a decoder launch for one batch of a page geometry
  uint64_t serial = 0; int t0 = -1, nb = 0;
  int4* d = nullptr; size_t cap = 0; int count = 0;
};

// One cached page geometry (H, W, margin): every page of a production run has its own border crop
// (main.py:2061 -> 2072), so the tile/owner tables and the decoder work lists of the last few geometries stay
// resident -- a hit touches nothing; a miss fills the least recently used slot through pinned staging, without
// a host synchronisation.
struct Geom {
  int H = 0, W = 0, margin = -2, n_pages = 1;   // n_pages same-size pages stacked vertically (H = ONE page's height)
  uint64_t serial = 0, last_use = 0;   // serial 0: slot never filled
  int nxf = 0, nyf = 0;
  int32_t* d_tile_org = nullptr; size_t cap_tiles = 0;
  int16_t* d_owner_x = nullptr; size_t cap_x = 0;
  int16_t* d_owner_y = nullptr; size_t cap_y = 0;
  std::vector<Rect> keep;              // per page tile: bounding box of the pixels it owns (tile coordinates)
};

struct Op {
  OpKind kind;
  std::string name;
  std::vector<ConvParams> variants;  // OP_CONV: 1, or the 4 output-parity classes of a decoder block
  ConvParams* d_variants = nullptr;
  __half* pool_out = nullptr;        // OP_POOL
  bool head = false;
  bool flat = false;      // logical grid is the flattened N*H*W pixel list
  int64_t per_img_px = 0; // flat: pixels per image
  int GW = 0, GH = 0;     // per-image logical grid of one variant
  int BN = 0;
  int dec_level = 0;      // decoder block 1..5 (0: not a decoder launch)
  bool pair = false;      // runs on conv_gemm_pair_kernel (CTA pairs, cta_group::2 MMAs)
  bool resb = false;      // ... with the whole (<= 9 chunk) weight matrix resident in shared memory (N = 64 launches)
  bool chain = false;     // variant 0 = an expand conv (2c), variant 1 = the next block's reduce conv (2a), ONE launch ordered by
                          // per-M-tile completion counters (LaunchArgs::chain_flags): the reduce conv reads its input from L2
  uint32_t* chain_flags = nullptr;
  int chain_flags_n = 0;
  int sub_stage = 0;      // ResNet stage (2..5) this launch belongs to (0: none)
  int sub_parts = 1;      // SBB_SUBBATCH: the stage runs over the batch in this many parts
  int sub_align = 1;      // parts are multiples of this many images (the largest images-per-tile of the stage)
  std::vector<WorkList> lists;
  double flops_per_img = 0.0;  // algorithmic FLOPs (all variants)
  float ms = 0.0f;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct ActInfo {
  std::string name;
  Tensor t;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct sbb_model {
  int tile_h, tile_w, n_classes, precision, backend, device, NB;
  int planes;
  int win_chunks = 4;
  int wide_n = 1;
  int res_in_mma = 1;
  int debug = 0;
  uint32_t* role_buf = nullptr;  // SBB_DEBUG bit 16: [num_sms][16] wait-cycle counters of the last launch
  int num_sms = 0;
  cudaStream_t own_stream = nullptr;
  EncodeTiledFn encode = nullptr;
  std::vector<void*> allocs;
  std::vector<Op> ops;
  std::vector<ActInfo> acts;
  Tensor f1;
  __half* xp = nullptr;  // packed, zero-bordered input tiles [NB][PH][pitch][8] (kernels_aux.cuh: stem_pad_kernel)
  int PH = 0, pitch = 0;
  // stem
  float *bn1_scale = nullptr, *bn1_shift = nullptr;
  // head constants
  float *w_cls = nullptr, *b_cls = nullptr;
  // page-mode scratch
  std::vector<Geom> geoms;            // LRU of page geometries (SBB_GEOM_CACHE slots, default 8)
  Geom full_geom;                     // the single whole-image "tile" of sbb_predict_full (filled at create)
  const Geom* cur = nullptr;          // geometry of the call in progress
  uint64_t use_clock = 0;
  uint8_t* stage = nullptr; size_t stage_cap = 0, stage_used = 0;  // pinned staging arena for table / work-list uploads
  cudaEvent_t stage_ev = nullptr; bool stage_pending = false;       // recorded after the last upload out of the arena
  cudaEvent_t chain_ev = nullptr; cudaStream_t last_stream = nullptr; bool chain_pending = false;
  uint8_t* d_page = nullptr; size_t page_cap = 0;
  uint8_t* d_labels = nullptr; size_t labels_cap = 0;
  float* d_tiles = nullptr; float* d_probs = nullptr; float* d_logits = nullptr; uint8_t* d_tlabels = nullptr;
  int last_nb = 0;
  // page-mode work lists: region of each tile the stitch keeps, per batch image
  uint64_t geom_serial = 0;           // last serial handed to a geometry slot
  int64_t geom_hits = 0, geom_misses = 0;
  int crop = 1;                       // SBB_CROP=0 disables the margin crop of decoder work
  int dec_rect = 1;                   // SBB_DEC_RECT=0: decoder tile shapes from the full grid (choose_rect) only
  int dec5_merged = 1;                // SBB_DEC5_MERGED=0: dec5 as four output-parity variants of N = 32
  int img_boxes = 1;                  // SBB_IMG_BOXES=0: encoder M tiles never span images (choose_rect)
  int pair_mode = 2;                  // SBB_PAIR: 0 never; 2 every N = 128 launch with >= pair_min_chunks K chunks runs as CTA
                                      // pairs; 1 only the multi-tap ones (3x3 convs, decoder blocks, head)
  int pair_min_chunks = 4;            // SBB_PAIR_MIN_CHUNKS
  int pair64 = 2;                     // SBB_PAIR64: 0 the N = 64 launches (conv1, stage-2 2a / 2b) stay on the single-CTA
                                      // kernel; 1 pairs; 2 pairs + the stem's A x [B_hi; B_lo] as one N = 128 MMA
  int direct_store = 0;               // SBB_DIRECT_STORE=1 (experiment build -DSBB_X_DIRECT_STORE only): the CTA-pair kernel's
                                      // epilogue warps write their rows with st.global instead of TMA stores -- slower
  int chain = 0;                      // SBB_CHAIN=1: an identity block's expand conv and the next block's reduce conv as ONE
                                      // flag-ordered launch (stages 3-5).  Bit-identical, but no gain: see build_plan
  int pair_resb = 0;                  // SBB_PAIR_RESB=1: N = 64 pair launches keep their (<= 9 chunk) weight matrix resident in
                                      // shared memory -- measured without a gain (profiles/r02x_resident_b_abab.txt), off
  int pair_head = 1;                  // SBB_PAIR_HEAD=0: the fused head (dec5) stays on the single-CTA kernel
  int sub_parts[6] = {1, 1, 1, 1, 1, 1};  // SBB_SUBBATCH="4:2,3:4": ResNet stage -> parts (see forward)
  int dec4_merged = 1;                // SBB_DEC4_MERGED=0: dec4 as four output-parity variants of N = 64 (single-CTA kernel)
  int64_t launches = 0;
  bool profiling = false;
  bool part_profiling = false;        // sbb_model_set_profiling(m, 2): three events per forward (start | first decoder launch | end)
  static constexpr int kPartSlots = 64;                // forwards that can be pending between two reads
  cudaEvent_t part_ev[kPartSlots][3] = {};             // created by sbb_model_set_profiling(m, 2)
  int part_count = 0;                                  // forwards recorded since the last read
  cudaStream_t part_stream = nullptr;
  size_t bytes_allocated = 0;
};

static int dev_alloc(sbb_model* m, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) return fail(SBB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  m->allocs.push_back(*p);
  m->bytes_allocated += bytes;
  return SBB_OK;
}
#define TRY(expr) do { int rc__ = (expr); if (rc__ != SBB_OK) return rc__; } while (0)

template <typename T>
static int ensure(sbb_model* m, T** p, size_t* cap, size_t need) {
  if (*cap >= need) return SBB_OK;
  T* np_ = nullptr;
  TRY(dev_alloc(m, (void**)&np_, need * sizeof(T)));  // old buffer is released at destroy
  *p = np_;
  *cap = need;
  return SBB_OK;
}

// Entry points run on the model's (or the caller's) device and leave the calling thread's current CUDA
// context as they found it: a one-process-per-GPU host (torch) must not find its current device changed
// underneath it, and a thread that had NO context must not get a stray primary context on device 0 from a
// restoring cudaSetDevice(0).  Driver entry points come from the statically linked runtime (no -lcuda).
struct DeviceScope {
  typedef CUresult (*GetFn)(CUcontext*);
  typedef CUresult (*SetFn)(CUcontext);
  CUcontext saved = nullptr;
  bool restore = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceScope(int device) {
    static GetFn get = nullptr;
    if (!get) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        get = reinterpret_cast<GetFn>(fn);
    }
    if (get && get(&saved) == CUDA_SUCCESS) restore = true;
    err = cudaSetDevice(device);
  }
  ~DeviceScope() {
    static SetFn set = nullptr;
    if (!restore) return;
    if (!set) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuCtxSetCurrent", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        set = reinterpret_cast<SetFn>(fn);
    }
    if (set) set(saved);
  }
};
#define ENTER_DEVICE(dev)        \
  DeviceScope dev_scope__(dev);  \
  CU_TRY(dev_scope__.err)

// Pinned staging arena for small host->device uploads (tile / owner tables, decoder work lists): the copies
// are asynchronous and the host buffer must outlive them, so it belongs to the handle.  Space is handed out
// bump-style; when the arena is full (or another stream takes over) the event recorded after the last
// upload is waited for -- normally long complete -- and the arena starts over.
static int stage_alloc(sbb_model* m, size_t bytes, cudaStream_t st, void** out) {
  bytes = (bytes + 255) & ~size_t(255);
  if (!m->stage_ev) CU_TRY(cudaEventCreateWithFlags(&m->stage_ev, cudaEventDisableTiming));
  if (m->stage_used + bytes > m->stage_cap) {
    if (m->stage_pending) { CU_TRY(cudaEventSynchronize(m->stage_ev)); m->stage_pending = false; }
    m->stage_used = 0;
    if (bytes > m->stage_cap) {
      if (m->stage) CU_TRY(cudaFreeHost(m->stage));
      m->stage = nullptr;
      m->stage_cap = std::max(bytes, (size_t)4 << 20);
      CU_TRY(cudaMallocHost((void**)&m->stage, m->stage_cap));
    }
  }
  *out = m->stage + m->stage_used;
  m->stage_used += bytes;
  (void)st;
  return SBB_OK;
}
static int stage_upload(sbb_model* m, void* dst, const void* staged, size_t bytes, cudaStream_t st) {
  CU_TRY(cudaMemcpyAsync(dst, staged, bytes, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaEventRecord(m->stage_ev, st));
  m->stage_pending = true;
  return SBB_OK;
}

// Calls on one handle share its workspace, so they are ordered on the device even when the caller alternates
// streams: every call records an event at its end and a call on ANOTHER stream waits for it first.
static int chain_begin(sbb_model* m, cudaStream_t st) {
  if (!m->chain_ev) CU_TRY(cudaEventCreateWithFlags(&m->chain_ev, cudaEventDisableTiming));
  if (m->chain_pending && st != m->last_stream) CU_TRY(cudaStreamWaitEvent(st, m->chain_ev, 0));
  return SBB_OK;
}
static int chain_end(sbb_model* m, cudaStream_t st) {
  CU_TRY(cudaEventRecord(m->chain_ev, st));
  m->chain_pending = true;
  m->last_stream = st;
  return SBB_OK;
}

static int alloc_tensor(sbb_model* m, Tensor* t, int H, int W, int C) {
  t->H = H; t->W = W; t->C = C; t->planes = m->planes;
  size_t bytes = (size_t)m->NB * H * W * t->pix() * sizeof(__half);
  TRY(dev_alloc(m, (void**)&t->d, bytes));
  return SBB_OK;
}

static RawView flat_view(const sbb_model* m, const Tensor& t) {
  RawView v{};
  v.base = t.d;
  v.W = m->NB * t.H * t.W; v.H = 1; v.N = 1;
  v.sW = t.pix(); v.sH = (int64_t)v.W * t.pix(); v.sN = v.sH;
  v.lo_off = t.lo_off();
  return v;
}
static RawView full_view(const sbb_model* m, const Tensor& t) {
  RawView v{};
  v.base = t.d;
  v.W = t.W; v.H = t.H; v.N = m->NB;
  v.sW = t.pix(); v.sH = (int64_t)t.W * t.pix(); v.sN = (int64_t)t.H * t.W * t.pix();
  v.lo_off = t.lo_off();
  return v;
}
// every second pixel starting at (qy, qx)
static RawView sub2_view(const sbb_model* m, const Tensor& t, int qy, int qx) {
  RawView v{};
  v.base = t.d + ((int64_t)qy * t.W + qx) * t.pix();
  v.W = (t.W - qx + 1) / 2; v.H = (t.H - qy + 1) / 2; v.N = m->NB;
  v.sW = 2 * t.pix(); v.sH = 2 * (int64_t)t.W * t.pix(); v.sN = (int64_t)t.H * t.W * t.pix();
  v.lo_off = t.lo_off();
  return v;
}

static int encode_view(sbb_model* m, CUtensorMap* map, const RawView& v, int chan_extent, int BW, int BH, int BI = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)chan_extent, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)kChunk, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BI};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = m->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)v.base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SBB_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d (dims %llu %llu %llu %llu box %u %u)", (int)r,
                (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], box[1], box[2]);
  return SBB_OK;
}
// {32 ch, BW, BH, 1} boxes with the 64-byte swizzle: the epilogue's staging slices (TMA store of the output).
static int encode_slice_view(sbb_model* m, CUtensorMap* map, const RawView& v, int chan_extent, int BW, int BH, int BI = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)chan_extent, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BI};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = m->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)v.base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SBB_ERR_CUDA, "cuTensorMapEncodeTiled(slice) failed: %d (dims %llu %llu %llu %llu box %d %d)", (int)r,
                (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], BW, BH);
  return SBB_OK;
}
static int encode_wmat(sbb_model* m, CUtensorMap* map, const __half* w, int rows, int Ktot, int BN) {
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kChunk, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = m->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SBB_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  return SBB_OK;
}

// Best BW x BH (<= 128 pixels) rectangle for a GW x GH grid: maximise useful rows per 128-row MMA.
static void choose_rect(int GW, int GH, int* BW, int* BH) {
  double best = -1.0;
  for (int bw = 1; bw <= std::min(GW, 128); ++bw) {
    int bh = std::min(GH, 128 / bw);
    if (bh > 256) bh = 256;
    double tiles = (double)((GW + bw - 1) / bw) * ((GH + bh - 1) / bh);
    double eff = (double)GW * GH / (tiles * 128.0);
    if (eff > best + 1e-9 || (std::fabs(eff - best) <= 1e-9 && bw > *BW)) { best = eff; *BW = bw; *BH = bh; }
  }
}

// M tile of an encoder conv on a GW x GH map over `nimg` images: BW x BH pixels of BI consecutive images
// (one 4-D TMA box), BW*BH*BI <= 128.  Fewest tiles wins -- a 28x28 map fills 128 rows with {4, 4, 8 images}
// (294 instead of 336 tiles per 48 images), a 14x14 map 126 rows with {14, 3, 3 images}.
static void choose_box(int GW, int GH, int nimg, int* BW, int* BH, int* BI) {
  auto tiles_of = [&](int bw, int bh, int bi) {
    return (long)((GW + bw - 1) / bw) * ((GH + bh - 1) / bh) * ((nimg + bi - 1) / bi);
  };
  long fewest = -1;
  for (int bw = std::min(GW, 128); bw >= 1; --bw)
    for (int bh = std::min(GH, 128 / bw); bh >= 1; --bh)
      for (int bi = std::min(nimg, 128 / (bw * bh)); bi >= 1; --bi) {
        const long t = tiles_of(bw, bh, bi);
        if (fewest < 0 || t < fewest) fewest = t;
      }
  // among the shapes within 2 % of the fewest tiles: fewest images per box (a box that spans images reads as
  // many DRAM streams at once), then the fuller, then the wider tile
  bool have = false;
  *BW = *BH = *BI = 1;
  for (int bi = 1; bi <= std::min(nimg, 128) && !have; ++bi)
    for (int bw = std::min(GW, 128 / bi); bw >= 1; --bw)
      for (int bh = std::min(GH, 128 / (bw * bi)); bh >= 1; --bh) {
        if (tiles_of(bw, bh, bi) * 100 > fewest * 102) continue;
        if (!have || bw * bh > *BW * *BH) { *BW = bw; *BH = bh; *BI = bi; have = true; }
      }
}

// Region of a decoder block's output (level 1..5; level 5 = tile resolution) that is needed to
// produce the level-5 pixels inside `keep`: every block reads its low-res input at [X-1, X+1].
static Rect level_rect(Rect r, int level, int TH, int TW) {
  int H = TH, W = TW;
  for (int l = 5; l > level; --l) {
    if (r.x1 < r.x0 || r.y1 < r.y0) return r;
    H /= 2; W /= 2;
    r.x0 = std::max(0, (r.x0 >> 1) - 1); r.y0 = std::max(0, (r.y0 >> 1) - 1);
    r.x1 = std::min(W - 1, (r.x1 >> 1) + 1); r.y1 = std::min(H - 1, (r.y1 >> 1) + 1);
  }
  return r;
}

// Work items of ONE image of a decoder launch: the M tiles (bw x bh low-res pixels) that intersect the
// region `r` of the block's OUTPUT (hi-res coordinates of that level), for each output-parity variant
// (py, px) -- or, for the merged-parity head (py < 0: one item yields all four output pixels of a
// low-res pixel), once.  Tiles are anchored at the first needed pixel, not at the image origin, so that
// no tile straddles the region's edge needlessly; the variants and N tiles of one M tile are adjacent
// so that they share their input tiles in L2.  items == nullptr: count only.
static int enumerate_items(Rect r, int bw, int bh, int n_tiles_n, const int (*parity)[2], int n_variants, int img,
                           std::vector<int4>* items) {
  if (r.x1 < r.x0 || r.y1 < r.y0) return 0;
  int X0[4], Y0[4], nx[4], ny[4], mx = 0, my = 0, count = 0;
  for (int v = 0; v < n_variants; ++v) {
    const int py = parity[v][0], px = parity[v][1];
    int X1, Y1;
    if (py < 0) {  // merged parity: every low-res pixel with at least one needed output pixel
      X0[v] = r.x0 >> 1; Y0[v] = r.y0 >> 1; X1 = r.x1 >> 1; Y1 = r.y1 >> 1;
    } else if (px < 0) {  // merged COLUMN parities, row parity py: columns as above, rows Y with r.y0 <= 2Y+py <= r.y1
      X0[v] = r.x0 >> 1; X1 = r.x1 >> 1;
      Y0[v] = (r.y0 - py + 1) >> 1; Y1 = (r.y1 - py) < 0 ? -1 : (r.y1 - py) >> 1;
    } else {       // the low-res pixels X with r.x0 <= 2X+px <= r.x1
      X0[v] = (r.x0 - px + 1) >> 1; Y0[v] = (r.y0 - py + 1) >> 1;
      X1 = (r.x1 - px) < 0 ? -1 : (r.x1 - px) >> 1; Y1 = (r.y1 - py) < 0 ? -1 : (r.y1 - py) >> 1;
    }
    nx[v] = X1 < X0[v] ? 0 : (X1 - X0[v]) / bw + 1;
    ny[v] = Y1 < Y0[v] ? 0 : (Y1 - Y0[v]) / bh + 1;
    mx = std::max(mx, nx[v]); my = std::max(my, ny[v]);
  }
  for (int ty = 0; ty < my; ++ty)
    for (int tx = 0; tx < mx; ++tx)
      for (int v = 0; v < n_variants; ++v) {
        if (tx >= nx[v] || ty >= ny[v]) continue;
        count += n_tiles_n;
        if (items)
          for (int nt = 0; nt < n_tiles_n; ++nt)
            items->push_back(make_int4(v | (nt << 8), img, X0[v] + tx * bw, Y0[v] + ty * bh));
      }
  return count;
}

// Per page tile: bounding box (tile coordinates) of the pixels it owns after the stitch.
static std::vector<Rect> keep_boxes(const std::vector<int32_t>& org, const std::vector<int16_t>& ox,
                                    const std::vector<int16_t>& oy, int ntiles, int tile_h, int tile_w) {
  std::vector<Rect> keep(ntiles, Rect{0, 0, -1, -1});
  for (int t = 0; t < ntiles; ++t) {
    const int x0 = org[4 * t], y0 = org[4 * t + 1], i = org[4 * t + 2], j = org[4 * t + 3];
    Rect r{tile_w, tile_h, -1, -1};
    for (int x = 0; x < tile_w; ++x)
      if (ox[x0 + x] == i) { r.x0 = std::min(r.x0, x); r.x1 = std::max(r.x1, x); }
    for (int y = 0; y < tile_h; ++y)
      if (oy[y0 + y] == j) { r.y0 = std::min(r.y0, y); r.y1 = std::max(r.y1, y); }
    keep[t] = r;
  }
  return keep;
}

// Tile shape of a decoder launch.  Page calls skip the work outside the region the stitch keeps and
// anchor their tiles at that region, so the shape that fills the FULL grid best (choose_rect) is not
// the one that covers the kept regions with the fewest 128-row MMA tiles (a 112x1 tile needs two tiles
// per row for the 180 kept columns of dec5, a 60x2 tile three per TWO rows).  Count the work items of a
// nominal page -- the reference's margin rule on BASELINE config 2's 2800x2000 scaled to the tile size --
// plus one uncropped image (sbb_predict_tiles / sbb_predict_full) for every shape; cheapest wins, ties go
// to the better filled, then the wider tile.  Any shape is correct; this only picks the fastest.
static void choose_rect_dec(int tile_h, int tile_w, int level, int GW, int GH, int merged, int* BW, int* BH) {
  const int H = tile_h * 2800 / 448, W = tile_w * 2000 / 448;
  int nxf = 0, nyf = 0;
  if (sbb_compute_tile_grid(H, W, tile_h, tile_w, -1, &nxf, &nyf, nullptr, 0, nullptr, nullptr) != SBB_OK) {
    choose_rect(GW, GH, BW, BH);
    return;
  }
  const int ntiles = nxf * nyf;
  std::vector<int32_t> org(4 * (size_t)ntiles);
  std::vector<int16_t> ox(W), oy(H);
  sbb_compute_tile_grid(H, W, tile_h, tile_w, -1, &nxf, &nyf, org.data(), ntiles, ox.data(), oy.data());
  std::vector<Rect> need = keep_boxes(org, ox, oy, ntiles, tile_h, tile_w);
  for (Rect& r : need) r = level_rect(r, level, tile_h, tile_w);
  need.push_back(Rect{0, 0, 2 * GW - 1, 2 * GH - 1});
  const int par4[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}}, par1[1][2] = {{-1, -1}}, par2[2][2] = {{0, -2}, {1, -2}};
  const int (*par)[2] = merged == 1 ? par1 : (merged == 2 ? par2 : par4);
  const int n_par = merged == 1 ? 1 : (merged == 2 ? 2 : 4);
  long best = -1;
  for (int bw = std::min(GW, 128); bw >= std::min(GW, 4); --bw) {
    const int bh = std::min(GH, 128 / bw);
    long cost = 0;
    for (const Rect& r : need) cost += enumerate_items(r, bw, bh, 1, par, n_par, 0, nullptr);
    // strict '<' on (cost, -fill): bw descends, so among equals the wider tile is kept
    if (best < 0 || cost < best || (cost == best && bw * bh > *BW * *BH)) { best = cost; *BW = bw; *BH = bh; }
  }
}

extern "C" int sbb_plan_decoder_tiles(int32_t H, int32_t W, int32_t tile_h, int32_t tile_w, int32_t margin,
                                      int32_t level, int32_t merged, int32_t full_grid_shapes, int32_t* bw,
                                      int32_t* bh, int32_t* items, int32_t item_cap, int32_t* count) {
  if (level < 1 || level > 5 || !bw || !bh || !count) return fail(SBB_ERR_INVALID, "bad argument");
  if (tile_h <= 0 || tile_w <= 0 || tile_h % 32 || tile_w % 32) return fail(SBB_ERR_INVALID, "bad tile size");
  int nxf = 0, nyf = 0;
  TRY(sbb_compute_tile_grid(H, W, tile_h, tile_w, margin, &nxf, &nyf, nullptr, 0, nullptr, nullptr));
  const int ntiles = nxf * nyf;
  std::vector<int32_t> org(4 * (size_t)ntiles);
  std::vector<int16_t> ox(W), oy(H);
  TRY(sbb_compute_tile_grid(H, W, tile_h, tile_w, margin, &nxf, &nyf, org.data(), ntiles, ox.data(), oy.data()));
  const std::vector<Rect> keep = keep_boxes(org, ox, oy, ntiles, tile_h, tile_w);
  const int GW = tile_w >> (6 - level), GH = tile_h >> (6 - level);  // the launch's (half-resolution) grid
  if (merged < 0 || merged > 2) return fail(SBB_ERR_INVALID, "merged: 0 four parities, 1 all merged (head), 2 column parities merged (dec4)");
  if (full_grid_shapes) choose_rect(GW, GH, bw, bh);
  else choose_rect_dec(tile_h, tile_w, level, GW, GH, merged, bw, bh);
  const int par4[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}}, par1[1][2] = {{-1, -1}}, par2[2][2] = {{0, -2}, {1, -2}};
  const int (*par)[2] = merged == 1 ? par1 : (merged == 2 ? par2 : par4);
  const int n_par = merged == 1 ? 1 : (merged == 2 ? 2 : 4);
  std::vector<int4> v;
  for (int t = 0; t < ntiles; ++t)
    enumerate_items(level_rect(keep[t], level, tile_h, tile_w), *bw, *bh, 1, par, n_par, t, &v);
  *count = (int32_t)v.size();
  for (size_t i = 0; items && i < v.size() && (int32_t)i < item_cap; ++i) {
    items[4 * i + 0] = v[i].x & 255; items[4 * i + 1] = v[i].y; items[4 * i + 2] = v[i].z; items[4 * i + 3] = v[i].w;
  }
  return SBB_OK;
}

struct SegSpec {
  RawView view;
  int chan_extent;  // innermost extent of the view's tensor (planes * C)
  int dx, dy, c0, nchunks;
  WSrc w;
  int flags = 0;
};

struct ConvSpec {
  std::string name;
  std::vector<SegSpec> segs;
  std::vector<int> bias_recs;
  bool flat;
  int GW, GH;            // per-image logical grid (flat: GW = pixels per image, GH = 1)
  int BW = 0, BH = 0;    // M tile shape (0: choose_box / choose_rect on the full grid)
  int Cout;
  bool relu;
  __half* out; int64_t oN, oH, oW; int out_lo_off;
  __half* out2 = nullptr;   // merged column parities: columns [Cout/2, Cout) go to this (px = 1) sub-view, channels from 0
  int bias_mod = 0;         // != 0: bias[o] = record bias[o % bias_mod]
  const __half* res; int64_t rN, rH, rW; int res_lo_off;
  bool head;
  double flops_per_img;
};

// Adds one implicit GEMM as a new launch, or (append) as a further variant of the last launch.
static int build_conv(sbb_model* m, const std::vector<Rec>& recs, const ConvSpec& cs, bool append = false) {
  if (!append) {
    Op fresh;
    fresh.kind = OP_CONV;
    fresh.name = cs.name;
    fresh.head = cs.head;
    fresh.flat = cs.flat;
    fresh.per_img_px = (int64_t)cs.GW * cs.GH;
    fresh.GW = cs.GW; fresh.GH = cs.GH;
    fresh.BN = cs.Cout >= 128 ? 128 : cs.Cout;
    m->ops.push_back(fresh);
  }
  Op& op = m->ops.back();
  if (append && (op.GW != cs.GW || op.GH != cs.GH || op.head != cs.head || op.flat != cs.flat))
    return fail(SBB_ERR_INVALID, "%s: variant geometry mismatch", cs.name.c_str());
  op.flops_per_img += cs.flops_per_img;
  op.variants.emplace_back();
  ConvParams& p = op.variants.back();
  memset(&p, 0, sizeof p);
  p.planes = m->planes;
  p.Cout = cs.Cout;
  if (op.BN != 128 && op.BN != 64 && op.BN != 32) return fail(SBB_ERR_UNSUPPORTED, "%s: Cout %d", cs.name.c_str(), cs.Cout);
  p.n_tiles_n = cs.Cout / op.BN;
  p.BI = 1;
  if (cs.flat) { p.BW = 128; p.BH = 1; }
  else if (cs.BW > 0) { p.BW = cs.BW; p.BH = cs.BH; }
  else if (m->img_boxes && !cs.head && cs.res == nullptr && m->backend == SBB_BACKEND_TCGEN05)
    choose_box(cs.GW, cs.GH, m->NB, &p.BW, &p.BH, &p.BI);
  else choose_rect(cs.GW, cs.GH, &p.BW, &p.BH);
  // views: dedupe by (base, strides)
  std::vector<RawView> views;
  std::vector<int> view_ext;
  int total_chunks = 0;
  if ((int)cs.segs.size() > kMaxSegs) return fail(SBB_ERR_INVALID, "%s: too many segments", cs.name.c_str());
  for (size_t s = 0; s < cs.segs.size(); ++s) {
    const SegSpec& ss = cs.segs[s];
    int vi = -1;
    for (size_t k = 0; k < views.size(); ++k)
      if (views[k].base == ss.view.base && views[k].sW == ss.view.sW && views[k].sH == ss.view.sH) vi = (int)k;
    if (vi < 0) {
      if ((int)views.size() == kMaxViews) return fail(SBB_ERR_INVALID, "%s: too many views", cs.name.c_str());
      views.push_back(ss.view);
      view_ext.push_back(ss.chan_extent);
      vi = (int)views.size() - 1;
    }
    p.segs[s].view = (int16_t)vi;
    p.segs[s].dx = (int16_t)ss.dx; p.segs[s].dy = (int16_t)ss.dy;
    p.segs[s].c0 = (int16_t)ss.c0; p.segs[s].nchunks = (int16_t)ss.nchunks;
    p.segs[s].flags = (int16_t)ss.flags;
    total_chunks += ss.nchunks;
  }
  p.n_segs = (int)cs.segs.size();
  p.n_views = (int)views.size();
  p.total_chunks = total_chunks;
  p.Ktot = total_chunks * kChunk;
  for (size_t k = 0; k < views.size(); ++k) {
    p.views[k] = views[k];
    if (m->backend == SBB_BACKEND_TCGEN05) TRY(encode_view(m, &p.tmapA[k], views[k], view_ext[k], p.BW, p.BH, p.BI));
  }
  // weight matrix [planes*Cout][Ktot]
  {
    const int K = p.Ktot, Co = cs.Cout;
    std::vector<__half> w((size_t)m->planes * Co * K, __float2half(0.0f));
    int kbase = 0;
    for (const SegSpec& ss : cs.segs) {
      const Rec& r = recs[ss.w.identity ? 0 : ss.w.rec];
      const int nch = ss.nchunks * kChunk;
      for (int o = 0; o < Co; ++o)
        for (int c = 0; c < nch; ++c) {
          double val = 0.0;
          bool lo_slot = false;  // packed chunks: slot that multiplies the LO half of the activation
          if (ss.w.identity) {
            val = ((o % op.BN) == c) ? 1.0 : 0.0;
          } else if (ss.w.merged) {
            const int par = o / (Co / ss.w.n_par), oc = o % (Co / ss.w.n_par), py = par >> 1, px = par & 1;  // (py, px): head only
            if (ss.w.packed_row) {
              const int j = c / 8, slot = c % 8, ch = slot % 4;
              const int ky = ss.w.ky - py, kx = j - px;
              lo_slot = slot >= 4;
              if (ch < 3 && ky >= 0 && ky < r.kh && kx >= 0 && kx < r.kw)
                val = r.w[(((size_t)oc * r.kh + ky) * r.kw + kx) * r.cin + ss.w.cin0 + ch];
              else if (ch == 3 && !lo_slot && j == ss.w.kx) val = r.b[oc];
            } else if (ss.w.cin0 + c < r.cin) {
              for (int t = 0; t < r.kh * r.kw; ++t)
                if (ss.w.tapmask4[par] >> t & 1) val += (double)r.w[((size_t)oc * r.kh * r.kw + t) * r.cin + ss.w.cin0 + c];
            }
          } else if (ss.w.packed_row) {
            const int px = c / 8, slot = c % 8, ch = slot % 4;
            lo_slot = slot >= 4;
            if (px < r.kw && ch < 3) val = r.w[(((size_t)o * r.kh + ss.w.ky) * r.kw + px) * r.cin + ss.w.cin0 + ch];
            else if (ch == 3 && !lo_slot && px == ss.w.kx) val = r.b[o];
          } else if (ss.w.cin0 + c < r.cin) {
            if (ss.w.tapmask) {
              for (int t = 0; t < r.kh * r.kw; ++t)
                if (ss.w.tapmask >> t & 1) val += (double)r.w[((size_t)o * r.kh * r.kw + t) * r.cin + ss.w.cin0 + c];
            } else {
              val = r.w[(((size_t)o * r.kh + ss.w.ky) * r.kw + ss.w.kx) * r.cin + ss.w.cin0 + c];
            }
          }
          const __half hi = __float2half_rn((float)val);
          w[(size_t)o * K + kbase + c] = hi;
          // a_lo * w_lo is dropped everywhere (below fp32 resolution): lo slots get no lo weight
          if (m->planes == 2 && !lo_slot)
            w[(size_t)(Co + o) * K + kbase + c] = __float2half_rn((float)(val - (double)__half2float(hi)));
        }
      kbase += nch;
    }
    __half* dw;
    TRY(dev_alloc(m, (void**)&dw, w.size() * sizeof(__half)));
    CU_TRY(cudaMemcpy(dw, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
    p.wmat = dw;
    if (m->backend == SBB_BACKEND_TCGEN05) {
      TRY(encode_wmat(m, &p.tmapB, dw, m->planes * Co, K, op.BN));
      TRY(encode_wmat(m, &p.tmapBh, dw, m->planes * Co, K, op.BN / 2));
    }
    std::vector<float> b(Co, 0.0f);
    for (int ri : cs.bias_recs)
      for (int o = 0; o < Co; ++o) b[o] += recs[ri].b[cs.bias_mod ? o % cs.bias_mod : o];
    float* db;
    TRY(dev_alloc(m, (void**)&db, Co * sizeof(float)));
    CU_TRY(cudaMemcpy(db, b.data(), Co * sizeof(float), cudaMemcpyHostToDevice));
    p.bias = db;
  }
  p.out = cs.out; p.oN = cs.oN; p.oH = cs.oH; p.oW = cs.oW; p.out_lo_off = cs.out_lo_off;
#ifdef SBB_X_DIRECT_STORE
  p.out2 = cs.out2;
#endif
  p.res = cs.res; p.rN = cs.rN; p.rH = cs.rH; p.rW = cs.rW; p.res_lo_off = cs.res_lo_off;
  p.relu = cs.relu ? 1 : 0;
  if (m->backend == SBB_BACKEND_TCGEN05 && !cs.head) {
    // the epilogue stores (and fetches the residual) through TMA: same grid geometry as the launch
    auto grid_view = [&](const __half* base, int64_t sW, int64_t sH, int64_t sN, int lo) {
      RawView v{};
      v.base = base; v.lo_off = lo; v.sW = sW;
      if (cs.flat) { v.W = m->NB * cs.GW; v.H = 1; v.N = 1; v.sH = v.sN = (int64_t)v.W * sW; }
      else { v.W = cs.GW; v.H = cs.GH; v.N = m->NB; v.sH = sH; v.sN = sN; }
      return v;
    };
    const int out_c = cs.out2 ? cs.Cout / 2 : cs.Cout;   // channels of the tensor the store lands in
    TRY(encode_slice_view(m, &p.tmapOut, grid_view(cs.out, cs.oW, cs.oH, cs.oN, cs.out_lo_off), m->planes * out_c,
                          p.BW, p.BH, p.BI));
    if (cs.out2)
      TRY(encode_slice_view(m, &p.tmapOut2, grid_view(cs.out2, cs.oW, cs.oH, cs.oN, cs.out_lo_off), m->planes * out_c,
                            p.BW, p.BH, p.BI));
  }
  // a window's K steps are dealt round-robin to kNCH accumulator chains (conv_gemm_tc.cuh), so a window of
  // win_chunks * kNCH chunks keeps the per-accumulator chain length (the truncation error) unchanged
  p.win_chunks = m->win_chunks * (op.BN == 128 ? 1 : (op.BN == 64 ? 2 : 4));
  p.wide_n = m->wide_n;
  {
    // every TMEM window must feed all accumulator chains of the tile (conv_gemm_tc.cuh: kNCH), or the
    // epilogue would add an uninitialised chain
    const int nch = op.BN == 128 ? 1 : (op.BN == 64 ? 2 : 4);
    int in_win = 0, ks = 0;
    bool ok = true;
    for (const SegSpec& ss : cs.segs)
      for (int c = 0; c < ss.nchunks; ++c) {
        ks += seg_ksteps(ss.flags);
        if (++in_win == p.win_chunks) { ok = ok && ks >= nch; in_win = 0; ks = 0; }
      }
    if (in_win > 0) ok = ok && ks >= nch;
    if (!ok) return fail(SBB_ERR_UNSUPPORTED, "%s: a TMEM window with fewer than %d K steps", cs.name.c_str(), nch);
    // the MMA issuer's compile-time chain assignment (conv_gemm_tc.cuh, kLean) needs every chunk to carry a
    // multiple of kNCH K steps
    if (nch <= 2)
      for (const SegSpec& ss : cs.segs)
        if (seg_ksteps(ss.flags) % nch) return fail(SBB_ERR_UNSUPPORTED, "%s: %d K steps per chunk with %d accumulator chains", cs.name.c_str(), seg_ksteps(ss.flags), nch);
  }
  return SBB_OK;
}

static int find_rec(const std::vector<Rec>& recs, const std::string& name) {
  for (size_t i = 0; i < recs.size(); ++i)
    if (recs[i].name == name) return (int)i;
  return -1;
}

static void set_out_flat(ConvSpec* cs, const Tensor& t) {
  cs->out = t.d; cs->oW = t.pix(); cs->oH = 0; cs->oN = 0; cs->out_lo_off = t.lo_off();
}
static void set_out_full(ConvSpec* cs, const Tensor& t) {
  cs->out = t.d; cs->oW = t.pix(); cs->oH = (int64_t)t.W * t.pix(); cs->oN = (int64_t)t.H * t.W * t.pix();
  cs->out_lo_off = t.lo_off();
}

static int build_plan(sbb_model* m, const std::vector<Rec>& recs) {
  const int TH = m->tile_h, TW = m->tile_w;
  const int H1 = (TH + 6 - 7) / 2 + 1, W1 = (TW + 6 - 7) / 2 + 1;
  const int H2 = (H1 - 3) / 2 + 1, W2 = (W1 - 3) / 2 + 1;
  auto rec = [&](const std::string& n) { return find_rec(recs, n); };
  auto need = [&](const std::string& n, int kh, int kw, int cin, int cout) -> int {
    int i = find_rec(recs, n);
    if (i < 0) return fail(SBB_ERR_INVALID, "weight blob: missing record %s", n.c_str());
    const Rec& r = recs[i];
    if (r.kh != kh || r.kw != kw || r.cin != cin || r.cout != cout)
      return fail(SBB_ERR_INVALID, "weight blob: record %s is %dx%dx%d->%d, expected %dx%dx%d->%d", n.c_str(), r.kh,
                  r.kw, r.cin, r.cout, kh, kw, cin, cout);
    return SBB_OK;
  };

  // ---- stem: gather+pad -> conv1 (raw, = skip f1) -> bn_conv1+ReLU+maxpool
  TRY(need("conv1", 7, 7, 3, 64));
  TRY(need("bn_conv1", 0, 0, 0, 64));
  m->PH = TH + 6;
  m->pitch = TW + 16;  // 3 + TW + 3 pixels of image, rest slack for the 8-pixel TMA windows
  TRY(dev_alloc(m, (void**)&m->xp, ((size_t)m->NB * m->PH * m->pitch + 64) * 8 * sizeof(__half)));
  CU_TRY(cudaMemset(m->xp, 0, ((size_t)m->NB * m->PH * m->pitch + 64) * 8 * sizeof(__half)));
  {
    Op op; op.kind = OP_STEM_PAD; op.name = "stem_pad";
    m->ops.push_back(op);
  }
  // window view onto the packed image: element (k, X, Y, n) = xp[n][row0 + sy*Y][col0 + sx*X + k/8][k%8]
  auto xp_view = [&](int row0, int col0, int sy, int sx, int GW, int GH) {
    RawView v{};
    v.base = m->xp + ((int64_t)row0 * m->pitch + col0) * 8;
    v.W = GW; v.H = GH; v.N = m->NB;
    v.sW = (int64_t)sx * 8; v.sH = (int64_t)sy * m->pitch * 8; v.sN = (int64_t)m->PH * m->pitch * 8;
    v.lo_off = 0;
    return v;
  };
  Tensor f1;
  TRY(alloc_tensor(m, &f1, H1, W1, 64));
  m->f1 = f1;
  {
    // ZeroPadding2D(3) + Conv2D 7x7 stride 2: output (oy, ox), tap row ky reads the 7 padded pixels
    // (2*oy + ky, 2*ox .. 2*ox + 6): one packed 8-pixel window per tap row, 7 segments.
    ConvSpec cs{};
    cs.name = "conv1"; cs.flat = false; cs.GW = W1; cs.GH = H1; cs.Cout = 64; cs.relu = false;
    for (int ky = 0; ky < 7; ++ky) {
      SegSpec s{};
      s.view = xp_view(ky, 0, 2, 2, W1, H1); s.chan_extent = 64; s.c0 = 0; s.nchunks = 1;
      s.w = WSrc{rec("conv1"), ky, -1, 0, true};
      s.flags = kSegPacked;  // 7 pixels x 8 halves = 56 -> all 4 K steps
      cs.segs.push_back(s);
    }
    cs.bias_recs = {rec("conv1")};
    set_out_full(&cs, f1);
    cs.flops_per_img = 2.0 * H1 * W1 * 147 * 64;
    TRY(build_conv(m, recs, cs));
  }
  m->acts.push_back({"conv1", f1});
  Tensor p1;
  TRY(alloc_tensor(m, &p1, H2, W2, 64));
  {
    const Rec& r = recs[rec("bn_conv1")];
    TRY(dev_alloc(m, (void**)&m->bn1_scale, 64 * sizeof(float)));
    TRY(dev_alloc(m, (void**)&m->bn1_shift, 64 * sizeof(float)));
    CU_TRY(cudaMemcpy(m->bn1_scale, r.w, 64 * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(m->bn1_shift, r.b, 64 * sizeof(float), cudaMemcpyHostToDevice));
    Op op; op.kind = OP_POOL; op.name = "bn_relu_maxpool";
    op.pool_out = p1.d;
    m->ops.push_back(op);
  }
  m->acts.push_back({"pool1", p1});

  // ---- encoder stages
  struct StageDef { int stage; const char* blocks; int f1, f2, f3, stride; };
  const StageDef stages[4] = {{2, "abc", 64, 64, 256, 1}, {3, "abcd", 128, 128, 512, 2},
                              {4, "abcdef", 256, 256, 1024, 2}, {5, "abc", 512, 512, 2048, 2}};
  Tensor x = p1;
  Tensor feats[6];
  for (const StageDef& sd : stages) {
    const int Hs = sd.stride == 2 ? (x.H - 1) / 2 + 1 : x.H;
    const int Ws = sd.stride == 2 ? (x.W - 1) / 2 + 1 : x.W;
    for (const char* b = sd.blocks; *b; ++b) {
      const bool first = (*b == 'a');
      const int st = first ? sd.stride : 1;
      char base[64];
      snprintf(base, sizeof base, "res%d%c_branch", sd.stage, *b);
      const std::string n2a = std::string(base) + "2a", n2b = std::string(base) + "2b", n2c = std::string(base) + "2c",
                        n1 = std::string(base) + "1";
      TRY(need(n2a, 1, 1, x.C, sd.f1));
      TRY(need(n2b, 3, 3, sd.f1, sd.f2));
      TRY(need(n2c, 1, 1, sd.f2, sd.f3));
      Tensor h1, h2, xo;
      TRY(alloc_tensor(m, &h1, Hs, Ws, sd.f1));
      TRY(alloc_tensor(m, &h2, Hs, Ws, sd.f2));
      TRY(alloc_tensor(m, &xo, Hs, Ws, sd.f3));
      {  // 2a: 1x1 (stride st) + BN + ReLU
        ConvSpec cs{};
        cs.name = n2a; cs.Cout = sd.f1; cs.relu = true;
        SegSpec s{};
        s.chan_extent = (int)x.pix(); s.c0 = 0; s.nchunks = x.C / kChunk; s.w = WSrc{rec(n2a), 0, 0, 0, false};
        if (st == 1) { cs.flat = true; cs.GW = Hs * Ws; cs.GH = 1; s.view = flat_view(m, x); set_out_flat(&cs, h1); }
        else { cs.flat = false; cs.GW = Ws; cs.GH = Hs; s.view = sub2_view(m, x, 0, 0); set_out_full(&cs, h1); }
        cs.segs.push_back(s);
        cs.bias_recs = {rec(n2a)};
        cs.flops_per_img = 2.0 * Hs * Ws * x.C * sd.f1;
        TRY(build_conv(m, recs, cs));
      }
      {  // 2b: 3x3 'same' + BN + ReLU
        ConvSpec cs{};
        cs.name = n2b; cs.Cout = sd.f2; cs.relu = true; cs.flat = false; cs.GW = Ws; cs.GH = Hs;
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            SegSpec s{};
            s.view = full_view(m, h1); s.chan_extent = (int)h1.pix(); s.dx = kx - 1; s.dy = ky - 1;
            s.c0 = 0; s.nchunks = sd.f1 / kChunk; s.w = WSrc{rec(n2b), ky, kx, 0, false};
            cs.segs.push_back(s);
          }
        cs.bias_recs = {rec(n2b)};
        set_out_full(&cs, h2);
        cs.flops_per_img = 2.0 * Hs * Ws * 9 * sd.f1 * sd.f2;
        TRY(build_conv(m, recs, cs));
      }
      {  // 2c: 1x1 + BN (+ projection shortcut K-concatenated | + identity residual) + ReLU
        ConvSpec cs{};
        cs.name = n2c; cs.Cout = sd.f3; cs.relu = true;
        SegSpec s{};
        s.chan_extent = (int)h2.pix(); s.c0 = 0; s.nchunks = sd.f2 / kChunk; s.w = WSrc{rec(n2c), 0, 0, 0, false};
        cs.bias_recs = {rec(n2c)};
        cs.flops_per_img = 2.0 * Hs * Ws * sd.f2 * sd.f3;
        if (first) {
          TRY(need(n1, 1, 1, x.C, sd.f3));
          SegSpec sc{};
          sc.chan_extent = (int)x.pix(); sc.c0 = 0; sc.nchunks = x.C / kChunk; sc.w = WSrc{rec(n1), 0, 0, 0, false};
          cs.bias_recs.push_back(rec(n1));
          cs.flops_per_img += 2.0 * Hs * Ws * x.C * sd.f3;
          if (st == 1) {
            cs.flat = true; cs.GW = Hs * Ws; cs.GH = 1;
            s.view = flat_view(m, h2); sc.view = flat_view(m, x); set_out_flat(&cs, xo);
          } else {
            cs.flat = false; cs.GW = Ws; cs.GH = Hs;
            s.view = full_view(m, h2); sc.view = sub2_view(m, x, 0, 0); set_out_full(&cs, xo);
          }
          cs.segs.push_back(s);
          cs.segs.push_back(sc);
        } else {
          cs.flat = true; cs.GW = Hs * Ws; cs.GH = 1;
          s.view = flat_view(m, h2); set_out_flat(&cs, xo);
          cs.segs.push_back(s);
          if (m->res_in_mma) {
            // identity shortcut: + x as one more K segment (the 128 channels of the N tile against an
            // identity block), prefetched by the operand pipeline
            SegSpec sr{};
            sr.view = flat_view(m, x); sr.chan_extent = (int)x.pix(); sr.c0 = 0;
            sr.nchunks = std::min(sd.f3, 128) / kChunk; sr.flags = kSegNtile;
            sr.w = WSrc{-1, 0, 0, 0, false};
            sr.w.identity = true;
            cs.segs.push_back(sr);
          } else {
            cs.res = x.d; cs.rW = x.pix(); cs.rH = 0; cs.rN = 0; cs.res_lo_off = x.lo_off();
          }
        }
        TRY(build_conv(m, recs, cs));
      }
      x = xo;
      char an[32];
      snprintf(an, sizeof an, "res%d%c", sd.stage, *b);
      m->acts.push_back({an, xo});
    }
    feats[sd.stage] = x;
  }

  // ---- decoder
  auto conv1x1 = [&](const std::string& name, const Tensor& in, Tensor* out, int cout) -> int {
    TRY(need(name, 1, 1, in.C, cout));
    TRY(alloc_tensor(m, out, in.H, in.W, cout));
    ConvSpec cs{};
    cs.name = name; cs.Cout = cout; cs.relu = true; cs.flat = true; cs.GW = in.H * in.W; cs.GH = 1;
    SegSpec s{};
    s.view = flat_view(m, in); s.chan_extent = (int)in.pix(); s.c0 = 0; s.nchunks = in.C / kChunk;
    s.w = WSrc{rec(name), 0, 0, 0, false};
    cs.segs.push_back(s);
    cs.bias_recs = {rec(name)};
    set_out_flat(&cs, *out);
    cs.flops_per_img = 2.0 * in.H * in.W * in.C * cout;
    return build_conv(m, recs, cs);
  };
  Tensor v5, v4;
  TRY(conv1x1("dec_v5", feats[5], &v5, 512));
  m->acts.push_back({"dec_v5", v5});
  TRY(conv1x1("dec_v4", feats[4], &v4, 512));
  m->acts.push_back({"dec_v4", v4});

  // UpSampling2D(2) + concatenate([up, skip]) + ZeroPadding2D(1) + Conv3x3 'valid' + BN + ReLU,
  // ONE launch whose four variants are the output parity classes (py, px): output pixel (2Y+py, 2X+px); tap (ky, kx) reads
  //   up  [Y + floor((py+ky-1)/2), X + floor((px+kx-1)/2)]
  //   skip[2Y + py+ky-1 - shift,   2X + px+kx-1 - shift]     (shift = 1 for one_side_pad'ed f2)
  auto fdiv2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };
  auto decoder = [&](const std::string& name, int level, const Tensor& up, const Tensor* skip, int skip_shift,
                     int skip_c, Tensor* out, int cout, bool head) -> int {
    const int Cin = up.C + skip_c;
    TRY(need(name, 3, 3, Cin, cout));
    const int Ho = 2 * up.H, Wo = 2 * up.W;
    if (!head) TRY(alloc_tensor(m, out, Ho, Wo, cout));
    const bool merged = head && m->dec5_merged;
    // a block with 64 output channels (dec4) fills only half of an N = 128 tile: its two COLUMN parities are
    // merged into one N = 2 x 64 GEMM per row parity (CTA-pair kernel only): 6 up-sampled + 12 skip taps = 24
    // chunks for two parities instead of 2 x 17, every A tile feeding twice the columns
    const bool merged_px = !head && cout == 64 && m->dec4_merged && m->pair_mode != 0 && m->backend == SBB_BACKEND_TCGEN05 &&
                           m->planes == 2 && m->wide_n && (m->debug & ~16) == 0 && skip != nullptr;
    int BWd = 0, BHd = 0;
    if (m->dec_rect) choose_rect_dec(TH, TW, level, up.W, up.H, merged ? 1 : (merged_px ? 2 : 0), &BWd, &BHd);
    if (merged_px) {
      for (int py = 0; py < 2; ++py) {
        ConvSpec cs{};
        cs.name = name; cs.Cout = 2 * cout; cs.relu = true; cs.flat = false; cs.GW = up.W; cs.GH = up.H; cs.head = false;
        cs.BW = BWd; cs.BH = BHd;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            uint32_t mask[2] = {0, 0};
            for (int px = 0; px < 2; ++px)
              for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx)
                  if (fdiv2(py + ky - 1) == dy && fdiv2(px + kx - 1) == dx) mask[px] |= 1u << (ky * 3 + kx);
            if (!mask[0] && !mask[1]) continue;
            SegSpec s{};
            s.view = full_view(m, up); s.chan_extent = (int)up.pix();
            s.dy = dy; s.dx = dx; s.c0 = 0; s.nchunks = up.C / kChunk;
            s.w = WSrc{rec(name), 0, 0, 0, false};
            s.w.merged = true; s.w.n_par = 2; s.w.tapmask4[0] = mask[0]; s.w.tapmask4[1] = mask[1];
            cs.segs.push_back(s);
          }
        for (int ky = 0; ky < 3; ++ky) {
          const int dyv = py + ky - 1 - skip_shift;
          const int qy = ((dyv % 2) + 2) % 2;
          for (int dxv = -1 - skip_shift; dxv <= 2 - skip_shift; ++dxv) {   // px + kx - 1 - shift over both px
            const int qx = ((dxv % 2) + 2) % 2;
            SegSpec k{};
            k.view = sub2_view(m, *skip, qy, qx); k.chan_extent = (int)skip->pix();
            k.dy = (dyv - qy) / 2; k.dx = (dxv - qx) / 2;
            k.c0 = 0; k.nchunks = skip->C / kChunk;
            k.w = WSrc{rec(name), 0, 0, up.C, false};
            k.w.merged = true; k.w.n_par = 2;
            for (int px = 0; px < 2; ++px) {
              const int kx = dxv + 1 + skip_shift - px;
              if (kx >= 0 && kx < 3) k.w.tapmask4[px] = 1u << (ky * 3 + kx);
            }
            cs.segs.push_back(k);
          }
        }
        cs.bias_recs = {rec(name)};
        cs.bias_mod = cout;
        cs.out = out->d + ((int64_t)py * Wo + 0) * out->pix();
        cs.out2 = out->d + ((int64_t)py * Wo + 1) * out->pix();
        cs.oW = 2 * out->pix(); cs.oH = 2 * (int64_t)Wo * out->pix(); cs.oN = (int64_t)Ho * Wo * out->pix();
        cs.out_lo_off = out->lo_off();
        cs.flops_per_img = 2 * (2.0 * up.H * up.W * 9 * Cin * cout);
        TRY(build_conv(m, recs, cs, /*append=*/py > 0));
        m->ops.back().variants.back().head_py = py;
        m->ops.back().variants.back().head_px = -2;
        m->ops.back().dec_level = level;
      }
      return SBB_OK;
    }
    if (merged) {
      // dec5 with the four output parities MERGED into one N = 4*32 GEMM over the low-res grid: item
      // (Y, X) yields output pixels (2Y+py, 2X+px); the up-sampled input is read through the 9 low-res
      // taps (dy, dx) -- each parity has weights on its 2x2 of them, zeros elsewhere -- and the raw-input
      // skip through the 4 rows of the pixel's 4x4 hi-res window (one packed 4-pixel window per row).
      // Per 512 output pixels that is 9 + 4 operand tiles instead of 4 * (4 + 3): the launch was bound by
      // operand rows, not by MMA issue, so the zero blocks are free.
      ConvSpec cs{};
      cs.name = name; cs.Cout = 4 * cout; cs.relu = true; cs.flat = false; cs.GW = up.W; cs.GH = up.H; cs.head = true;
      cs.BW = BWd; cs.BH = BHd;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          SegSpec s{};
          s.view = full_view(m, up); s.chan_extent = (int)up.pix();
          s.dy = dy; s.dx = dx; s.c0 = 0; s.nchunks = up.C / kChunk;
          s.w = WSrc{rec(name), 0, 0, 0, false};
          s.w.merged = true;
          for (int par = 0; par < 4; ++par)
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx)
                if (fdiv2((par >> 1) + ky - 1) == dy && fdiv2((par & 1) + kx - 1) == dx) s.w.tapmask4[par] |= 1u << (ky * 3 + kx);
          cs.segs.push_back(s);
        }
      // hi-res row 2Y-1+r, columns 2X-1 .. 2X+2 = padded-image pixels (2Y+r+2, 2X+2 .. 2X+5): 4 pixels x 8
      // halves = 32 -> 2 K steps.  The bias rides on the constant-1 channel of window pixel (1, 1), which
      // every parity's 3x3 covers.
      for (int r = 0; r < 4; ++r) {
        SegSpec k{};
        k.view = xp_view(r + 2, 2, 2, 2, up.W, up.H); k.chan_extent = 64;
        k.c0 = 0; k.nchunks = 1;
        k.w = WSrc{rec(name), r, r == 1 ? 1 : -1, up.C, true};
        k.w.merged = true;
        k.flags = kSegPacked | (2 << 4);
        cs.segs.push_back(k);
      }
      cs.flops_per_img = 4 * (2.0 * up.H * up.W * 9 * Cin * cout + 2.0 * up.H * up.W * 32 * m->n_classes);
      TRY(build_conv(m, recs, cs));
      m->ops.back().variants.back().head_py = -1;
      m->ops.back().variants.back().head_px = -1;
      m->ops.back().dec_level = level;
      return SBB_OK;
    }
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ConvSpec cs{};
        cs.name = name; cs.Cout = cout; cs.relu = true; cs.flat = false; cs.GW = up.W; cs.GH = up.H; cs.head = head;
        cs.BW = BWd; cs.BH = BHd;
        // up path: nearest 2x upsampling makes several taps read the SAME low-res pixel, so their
        // weights are pre-summed (sub-pixel identity): 4 merged taps instead of 9 per parity class.
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            uint32_t mask = 0;
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx)
                if (fdiv2(py + ky - 1) == dy && fdiv2(px + kx - 1) == dx) mask |= 1u << (ky * 3 + kx);
            if (!mask) continue;
            SegSpec s{};
            s.view = full_view(m, up); s.chan_extent = (int)up.pix();
            s.dy = dy; s.dx = dx; s.c0 = 0; s.nchunks = up.C / kChunk;
            s.w = WSrc{rec(name), 0, 0, 0, false, mask};
            cs.segs.push_back(s);
          }
        for (int ky = 0; ky < 3 && skip; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const int dyv = py + ky - 1 - skip_shift, dxv = px + kx - 1 - skip_shift;
            const int qy = ((dyv % 2) + 2) % 2, qx = ((dxv % 2) + 2) % 2;
            SegSpec k{};
            k.view = sub2_view(m, *skip, qy, qx); k.chan_extent = (int)skip->pix();
            k.dy = (dyv - qy) / 2; k.dx = (dxv - qx) / 2;
            k.c0 = 0; k.nchunks = skip->C / kChunk; k.w = WSrc{rec(name), ky, kx, up.C, false};
            cs.segs.push_back(k);
          }
        cs.bias_recs = {rec(name)};
        if (head) {
          // 'inp' skip of the last block: 3x3 taps over the 3 raw input channels at output pixel
          // (2Y+py, 2X+px) = padded-image pixels (2Y+py+ky+2, 2X+px+2 .. +4): one packed window per tap
          // row (3 pixels x 8 halves = 24 -> 2 K steps).  The bias rides on the centre pixel's
          // constant-1 channel, so the epilogue adds nothing.
          for (int ky = 0; ky < 3; ++ky) {
            SegSpec k{};
            k.view = xp_view(py + ky + 2, px + 2, 2, 2, up.W, up.H); k.chan_extent = 64;
            k.c0 = 0; k.nchunks = 1;
            k.w = WSrc{rec(name), ky, ky == 1 ? 1 : -1, up.C, true};
            k.flags = kSegPacked | (2 << 4);
            cs.segs.push_back(k);
          }
          cs.bias_recs.clear();
        }
        if (!head) {
          cs.out = out->d + ((int64_t)py * Wo + px) * out->pix();
          cs.oW = 2 * out->pix(); cs.oH = 2 * (int64_t)Wo * out->pix(); cs.oN = (int64_t)Ho * Wo * out->pix();
          cs.out_lo_off = out->lo_off();
        }
        cs.flops_per_img = 2.0 * up.H * up.W * 9 * Cin * cout;
        if (head) cs.flops_per_img += 2.0 * up.H * up.W * 32 * m->n_classes;
        TRY(build_conv(m, recs, cs, /*append=*/py + px > 0));
        m->ops.back().variants.back().head_py = py;
        m->ops.back().variants.back().head_px = px;
        m->ops.back().dec_level = level;
      }
    return SBB_OK;
  };
  Tensor d1, d2, d3, d4, none;
  TRY(decoder("dec1", 1, v5, &v4, 0, v4.C, &d1, 512, false));
  m->acts.push_back({"dec1", d1});
  TRY(decoder("dec2", 2, d1, &feats[3], 0, feats[3].C, &d2, 256, false));
  m->acts.push_back({"dec2", d2});
  TRY(decoder("dec3", 3, d2, &feats[2], 1, feats[2].C, &d3, 128, false));
  m->acts.push_back({"dec3", d3});
  TRY(decoder("dec4", 4, d3, &f1, 0, f1.C, &d4, 64, false));
  m->acts.push_back({"dec4", d4});
  if (d4.H != H1 || d4.W != W1 || 2 * d4.H != TH) return fail(SBB_ERR_UNSUPPORTED, "tile geometry mismatch");
  TRY(decoder("dec5", 5, d4, nullptr, 0, 3, &none, 32, true));

  // ---- head constants: classifier (+ folded BN)
  {
    TRY(need("cls", 1, 1, 32, m->n_classes));
    const Rec& rc = recs[rec("cls")];
    std::vector<float> wc(32 * 8, 0.0f), bc(8, 0.0f);
    for (int c = 0; c < m->n_classes; ++c) {
      bc[c] = rc.b[c];
      for (int j = 0; j < 32; ++j) wc[j * 8 + c] = rc.w[(size_t)c * 32 + j];
    }
    TRY(dev_alloc(m, (void**)&m->w_cls, wc.size() * 4));
    TRY(dev_alloc(m, (void**)&m->b_cls, bc.size() * 4));
    CU_TRY(cudaMemcpy(m->w_cls, wc.data(), wc.size() * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(m->b_cls, bc.data(), bc.size() * 4, cudaMemcpyHostToDevice));
  }
  // ---- sub-batched stages (SBB_SUBBATCH): tag the launches of each ResNet stage
  for (int stage = 2; stage <= 5; ++stage) {
    int align = 1;
    char pre[8];
    snprintf(pre, sizeof pre, "res%d", stage);
    for (Op& op : m->ops)
      if (op.kind == OP_CONV && op.name.compare(0, 4, pre) == 0) align = std::max(align, op.variants[0].BI);
    for (Op& op : m->ops)
      if (op.kind == OP_CONV && op.name.compare(0, 4, pre) == 0) {
        op.sub_stage = stage; op.sub_parts = m->sub_parts[stage]; op.sub_align = align;
      }
  }
  // ---- chained launches: an identity block's expand conv (2c) and the next block's reduce conv (2a) are both 1x1 convs
  // over the same flat 128-pixel M tiles, and the second reads exactly what the first writes -- 150..600 MB that make
  // the round trip through HBM between two launches.  As two variants of ONE work-list launch whose reduce-conv items
  // trail their M tile's expand-conv items by a couple of waves (ordered by per-M-tile counters, conv_gemm_pair.cuh) the
  // tensor is read back from L2 and a launch disappears.  Measured (profiles/r02ai_chain_abab.txt, r02aj_*): results
  // bit-identical, page time unchanged (8.64 / 8.71 vs 8.68 / 8.69 ms) -- the reduce conv waits for operands just as
  // long when they come from L2 (unshared A lines arrive at ~28 B/clk/SM from either source), and the release fence
  // that publishes a tile's stores costs the epilogue ~15 % of the launch.  Off by default (SBB_CHAIN=1).
  {
    bool sub = false;
    for (int st = 2; st <= 5; ++st) sub = sub || m->sub_parts[st] > 1;
    const bool can = m->chain && !sub && m->pair_mode >= 2 && m->backend == SBB_BACKEND_TCGEN05 && m->planes == 2 && m->wide_n &&
                     (m->debug & ~16) == 0;
    for (size_t i = 0; can && i + 1 < m->ops.size(); ++i) {
      Op& A = m->ops[i];
      const Op& B = m->ops[i + 1];
      if (A.kind != OP_CONV || B.kind != OP_CONV || A.chain || !A.flat || !B.flat || A.head || B.head) continue;
      if (A.BN != 128 || B.BN != 128 || A.variants.size() != 1 || B.variants.size() != 1) continue;
      if (A.per_img_px != B.per_img_px || A.sub_stage == 0 || A.sub_stage != B.sub_stage) continue;
      const ConvParams& a0 = A.variants[0];
      const ConvParams& b0 = B.variants[0];
      if (b0.n_segs != 1 || b0.views[b0.segs[0].view].base != a0.out || a0.res != nullptr || b0.res != nullptr) continue;
      if (a0.total_chunks < m->pair_min_chunks || b0.total_chunks < m->pair_min_chunks || a0.BI != 1 || b0.BI != 1) continue;
      A.variants.push_back(b0);
      A.name += "+" + B.name;
      A.flops_per_img += B.flops_per_img;
      A.chain = true;
      A.chain_flags_n = (int)((A.per_img_px * m->NB + 127) / 128) + 2;
      TRY(dev_alloc(m, (void**)&A.chain_flags, (size_t)A.chain_flags_n * sizeof(uint32_t)));
      m->ops.erase(m->ops.begin() + i + 1);
    }
  }
  // ---- the static launch descriptions (incl. TMA descriptors) live in device memory
  for (Op& op : m->ops) {
    if (op.kind != OP_CONV) continue;
    for (const ConvParams& v : op.variants)
      if (v.BW != op.variants[0].BW || v.BH != op.variants[0].BH || v.BI != op.variants[0].BI ||
          (!op.chain && v.n_tiles_n != op.variants[0].n_tiles_n))
        return fail(SBB_ERR_INVALID, "%s: variants disagree on the tile shape", op.name.c_str());
    {
      // CTA pairs: every N = 128 launch with at least pair_min_chunks K chunks -- the 3x3 convs, decoder blocks and the
      // merged-parity head gain 15-25 %, the 1x1 convs of stages 3-5 5-15 % (profiles/r02s_pair_scope_sweep.txt; before
      // the TMEM hand-over lost its GPU-scope membar the 1x1 convs LOST in pair mode, r02e); the 2-3 chunk expand convs
      // of stage 2 sit at the HBM roofline and stay single-CTA.  Packed (hi, lo)-interleaved views are handled for
      // the head's input-skip rows only (2 K steps per chunk).
      bool ok = m->pair_mode != 0 && m->backend == SBB_BACKEND_TCGEN05 && m->planes == 2 && m->wide_n &&
                (op.BN == 128 || (op.BN == 64 && m->pair64 && !op.head)) &&
                (m->debug & ~16) == 0 && (!op.head || (op.variants.size() == 1 && op.variants[0].head_py < 0 && m->pair_head));
      for (const ConvParams& v : op.variants) {
        ok = ok && v.total_chunks >= m->pair_min_chunks && v.res == nullptr && (m->pair_mode >= 2 || v.n_segs >= 4);
        for (int sgi = 0; sgi < v.n_segs; ++sgi)   // packed views: the head's input-skip rows and the stem (N = 64)
          ok = ok && (op.head || op.BN == 64 || !(v.segs[sgi].flags & kSegPacked)) &&
               (op.head || seg_ksteps(v.segs[sgi].flags) == 4);
      }
      op.pair = ok;
      if (op.chain && !ok) return fail(SBB_ERR_UNSUPPORTED, "%s: a chained launch needs the CTA-pair kernel", op.name.c_str());
      op.resb = op.pair && op.BN == 64 && m->pair_resb && op.variants.size() == 1 && op.variants[0].n_tiles_n == 1 &&
                op.variants[0].total_chunks <= PairCfg<false, 64, true>::kResBChunks && op.dec_level == 0;
      if (op.pair && op.BN == 64 && m->pair64 >= 2) {   // all-packed N = 64 launch (the stem): one wide MMA per K step
        bool all_packed = true;
        for (const ConvParams& v : op.variants)
          for (int sgi = 0; sgi < v.n_segs; ++sgi) all_packed = all_packed && (v.segs[sgi].flags & kSegPacked);
        if (all_packed)
          for (ConvParams& v : op.variants) v.wide_n = 2;
      }
      for (const ConvParams& v : op.variants)
        if (!op.head && v.head_px == -2 && !op.pair)
          return fail(SBB_ERR_UNSUPPORTED, "%s: merged column parities need the CTA-pair kernel (SBB_PAIR_MIN_CHUNKS too high?)",
                      op.name.c_str());
    }
    TRY(dev_alloc(m, (void**)&op.d_variants, op.variants.size() * sizeof(ConvParams)));
    CU_TRY(cudaMemcpy(op.d_variants, op.variants.data(), op.variants.size() * sizeof(ConvParams), cudaMemcpyHostToDevice));
  }
  return SBB_OK;
}

// ------------------------------------------------------------------------------------------ launch
template <int BN, bool SPLIT, bool HEAD>
static int launch_tc(sbb_model* m, const LaunchArgs& a, cudaStream_t st) {
  using Cfg = TcCfg<BN, SPLIT, HEAD>;
  static bool configured[16] = {false};
  auto kern = conv_gemm_tc_kernel<BN, SPLIT, HEAD>;
  if (!configured[m->device & 15]) {
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured[m->device & 15] = true;
  }
  if (a.total_work <= 0) return SBB_OK;
  const int grid = std::min(a.total_work, m->num_sms);
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(a);
  CU_TRY(cudaGetLastError());
  m->launches++;
  return SBB_OK;
}

template <bool HEAD, int BN, bool RESB = false>
static int launch_pair(sbb_model* m, const LaunchArgs& a, cudaStream_t st) {
  using Cfg = PairCfg<HEAD, BN, RESB>;
  static int max_clusters[16] = {0};
  int& mc = max_clusters[m->device & 15];
  auto kern = conv_gemm_pair_kernel<HEAD, BN, RESB>;
  if (mc == 0) {
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    // the persistent loop strides by the number of clusters: launch no more than can be resident at once
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(m->num_sms & ~1u); cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = m->num_sms / 2;
    }
    mc = std::min(n, m->num_sms / 2);
    if (m->debug & 16) fprintf(stderr, "[pair] %d CTA pairs resident on %d SMs\n", mc, m->num_sms);
  }
  if (a.total_work <= 0) return SBB_OK;
  const int n_pairs = a.worklist != nullptr ? a.total_work / 2 : ((a.total_work / a.n_tiles_n + 1) / 2) * a.n_tiles_n;
  const int clusters = std::min(n_pairs, mc);
  kern<<<2 * clusters, Cfg::kThreads, Cfg::kSmemBytes, st>>>(a);   // __cluster_dims__(2, 1, 1)
  CU_TRY(cudaGetLastError());
  m->launches++;
  return SBB_OK;
}

// SBB_DEBUG bit 16: where the three single-thread roles and the epilogue of the last launch waited.
static int report_role_cycles(sbb_model* m, const Op& op, int grid, cudaStream_t st, bool pair = false) {
  std::vector<uint32_t> h((size_t)grid * 16);
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaMemcpy(h.data(), m->role_buf, h.size() * 4, cudaMemcpyDeviceToHost));
  double s[16] = {0};
  int live = 0;
  for (int c = 0; c < grid; ++c) {
    live += h[(size_t)c * 16 + 5] != 0;
    for (int k = 0; k < 16; ++k) s[k] += h[(size_t)c * 16 + k];
  }
  if (pair) {   // CTA-pair kernel (experiment build -DSBB_X_ROLES): `grid` is an upper bound; only the leader of a pair issues MMAs
    grid = live > 0 ? live : 1;
    s[1] *= 2; s[2] *= 2; s[8] *= 2;
  }
  const double tot = s[5] > 0 ? s[5] : 1;
  fprintf(stderr, "[roles] %-16s grid %3d items/cta %6.1f cycles/item %7.0f | producer waits stage %4.1f%% | mma waits operands "
          "%4.1f%% tmem %4.1f%% issue %4.1f%% | epilogue waits window %4.1f%% store handoff %4.1f%% | chain flag %4.1f%%\n",
          op.name.c_str(), grid, s[6] / grid, s[6] > 0 ? s[5] / s[6] : 0.0, 100 * s[0] / tot, 100 * s[1] / tot, 100 * s[2] / tot,
          100 * s[8] / tot, 100 * s[3] / tot, 100 * s[7] / tot, 100 * s[9] / tot);
  return SBB_OK;
}

// Work list of a decoder launch over batch images [t0, t0+nb): per image only the M tiles that
// intersect the needed region (the whole grid when `crop` is false).
static bool serial_live(const sbb_model* m, uint64_t serial) {
  if (serial == 0) return true;  // the uncropped full-grid lists never go stale
  for (const Geom& g : m->geoms)
    if (g.serial == serial) return true;
  return false;
}

static int get_worklist(sbb_model* m, Op& op, int t0, int nb, bool crop, cudaStream_t st, const int4** d_list,
                        int* count) {
  const uint64_t serial = crop ? m->cur->serial : 0;
  WorkList* wl = nullptr;
  for (WorkList& c : op.lists)
    if (c.serial == serial && c.t0 == t0 && c.nb == nb) { *d_list = c.d; *count = c.count; return SBB_OK; }
  for (WorkList& c : op.lists)
    if (!serial_live(m, c.serial)) wl = &c;  // list of an evicted page geometry: reuse its buffer
  const ConvParams& p0 = op.variants[0];
  // anchored tiles never outnumber the origin-anchored grid along an axis by more than one
  const int tiles_x = (op.GW + p0.BW - 1) / p0.BW + 1, tiles_y = (op.GH + p0.BH - 1) / p0.BH + 1;
  const size_t cap = (size_t)m->NB * tiles_x * tiles_y * p0.n_tiles_n * op.variants.size() + 64;
  if (!wl) {
    op.lists.emplace_back();
    wl = &op.lists.back();
    TRY(dev_alloc(m, (void**)&wl->d, cap * sizeof(int4)));
    wl->cap = cap;
  }
  int parity[4][2];
  for (size_t v = 0; v < op.variants.size(); ++v) { parity[v][0] = op.variants[v].head_py; parity[v][1] = op.variants[v].head_px; }
  std::vector<int4> items;
  items.reserve(cap);
  for (int b = 0; b < nb; ++b) {
    Rect r{0, 0, 2 * op.GW - 1, 2 * op.GH - 1};
    if (crop) r = level_rect(m->cur->keep[t0 + b], op.dec_level, m->tile_h, m->tile_w);
    enumerate_items(r, p0.BW, p0.BH, p0.n_tiles_n, parity, (int)op.variants.size(), b, &items);
  }
  if (op.pair) {
    // CTA pairs take items 2q and 2q+1, which must share the variant and the N tile (one weight tile, one MMA
    // stream): pair each item with the next one of the same (variant, N tile); a group's odd item gets a partner
    // on out-of-range coordinates (TMA zero-fills its loads and clips its stores).  Pairs keep the order of
    // their first items, so neighbouring pairs still share input tiles in L2.
    std::vector<int4> paired;
    paired.reserve(items.size() + 64);
    std::vector<char> used(items.size(), 0);
    for (size_t i = 0; i < items.size(); ++i) {
      if (used[i]) continue;
      used[i] = 1;
      paired.push_back(items[i]);
      size_t j = i + 1;
      while (j < items.size() && (used[j] || items[j].x != items[i].x)) ++j;
      if (j < items.size()) { used[j] = 1; paired.push_back(items[j]); }
      else paired.push_back(make_int4(items[i].x, items[i].y, 30000, 30000));
    }
    items.swap(paired);
  }
  if (items.size() > wl->cap) return fail(SBB_ERR_INVALID, "%s: work list overflow", op.name.c_str());
  if (!items.empty()) {
    void* h = nullptr;
    TRY(stage_alloc(m, items.size() * sizeof(int4), st, &h));
    memcpy(h, items.data(), items.size() * sizeof(int4));
    TRY(stage_upload(m, wl->d, h, items.size() * sizeof(int4), st));
  }
  wl->serial = serial; wl->t0 = t0; wl->nb = nb; wl->count = (int)items.size();
  *d_list = wl->d; *count = wl->count;
  return SBB_OK;
}

// Work list of a chained launch over nb images.  M-tile pair j = flat pixels [256 j, 256 j + 256); the expand conv has
// one pair item per (j, N tile), the reduce conv likewise.  The persistent grid of R clusters takes list positions
// round-robin, so the list is built in ROUNDS of R pair items of ONE kind: every cluster then runs the same sequence
// of item kinds (a reduce-conv item takes 2-3x as long as an expand-conv one; mixed rounds leave the launch waiting
// for the clusters that drew more of them -- measured +10 %).  A round of reduce-conv items is emitted once R of them
// have had their M tiles' expand-conv items in the list for at least three rounds: consumers never wait in steady
// state, and the tensor in flight (3-4 rounds, < 70 MB) stays inside L2.
static void chain_items(int64_t px, int nA, int nB, int R, std::vector<int4>* items) {
  const int Mt = (int)((px + 127) / 128), Mp = (Mt + 1) / 2;
  const int64_t totA = (int64_t)Mp * nA, totB = (int64_t)Mp * nB;
  auto emit = [&](int variant, int n_tiles, int64_t k) {   // k-th pair item of a kind: M pair k / n_tiles, N tile k % n_tiles
    const int j = (int)(k / n_tiles), nt = (int)(k % n_tiles);
    for (int r = 0; r < 2; ++r) items->push_back(make_int4(variant | (nt << 8), 0, (2 * j + r) * 128, 0));
  };
  int64_t a_done = 0, b_done = 0;
  while (a_done < totA || b_done < totB) {
    if (a_done < totA) {
      const int64_t n = std::min<int64_t>(R, totA - a_done);
      for (int64_t k = 0; k < n; ++k) emit(0, nA, a_done + k);
      a_done += n;
    }
    // reduce-conv items whose M pair was completely listed at least three rounds ago
    const int64_t old_a = a_done < totA ? std::max<int64_t>(0, a_done - 3 * R) : totA;
    const int64_t ready = std::min(totB, (old_a / nA) * nB);
    if (a_done >= totA) {
      for (; b_done < totB; ++b_done) emit(1, nB, b_done);
    } else if (ready - b_done >= R) {
      for (int64_t k = 0; k < R; ++k) emit(1, nB, b_done + k);
      b_done += R;
    }
  }
}

// Introspection for tests (no GPU): the chained work list for `px` flat pixels, nA / nB N tiles of the two convs and a
// persistent grid of R clusters; items as {variant | n_tile << 8, x0} pairs.
extern "C" int sbb_plan_chain_list(int64_t px, int32_t nA, int32_t nB, int32_t R, int32_t* items, int32_t item_cap, int32_t* count) {
  if (px <= 0 || nA <= 0 || nB <= 0 || R <= 0 || !count) return fail(SBB_ERR_INVALID, "bad argument");
  std::vector<int4> v;
  chain_items(px, nA, nB, R, &v);
  *count = (int32_t)v.size();
  if (items) {
    if ((int)v.size() > item_cap) return fail(SBB_ERR_INVALID, "item buffer too small: %d > %d", (int)v.size(), item_cap);
    for (size_t i = 0; i < v.size(); ++i) { items[2 * i] = v[i].x; items[2 * i + 1] = v[i].z; }
  }
  return SBB_OK;
}

static int get_chain_list(sbb_model* m, Op& op, int nb, cudaStream_t st, const int4** d_list, int* count) {
  for (WorkList& c : op.lists)
    if (c.nb == nb) { *d_list = c.d; *count = c.count; return SBB_OK; }
  const int nA = op.variants[0].n_tiles_n, nB = op.variants[1].n_tiles_n;
  auto tiles_of = [&](int n) { return (int)((op.per_img_px * n + 127) / 128); };
  const size_t cap = (size_t)((tiles_of(m->NB) + 1) / 2) * 2 * (nA + nB) + 64;
  WorkList* wl = nullptr;
  if (op.lists.size() >= 4) wl = &op.lists[nb % 4];   // a handful of batch sizes occur (full batches + a page's last one)
  else {
    op.lists.emplace_back();
    wl = &op.lists.back();
    TRY(dev_alloc(m, (void**)&wl->d, cap * sizeof(int4)));
    wl->cap = cap;
  }
  std::vector<int4> items;
  items.reserve(cap);
  chain_items(op.per_img_px * nb, nA, nB, std::max(1, m->num_sms / 2) /* clusters of the persistent grid (launch_pair) */, &items);
  if (items.size() > wl->cap) return fail(SBB_ERR_INVALID, "%s: chained work list overflow", op.name.c_str());
  void* h = nullptr;
  TRY(stage_alloc(m, items.size() * sizeof(int4), st, &h));
  memcpy(h, items.data(), items.size() * sizeof(int4));
  TRY(stage_upload(m, wl->d, h, items.size() * sizeof(int4), st));
  wl->serial = 0; wl->t0 = -2; wl->nb = nb; wl->count = (int)items.size();
  *d_list = wl->d; *count = wl->count;
  return SBB_OK;
}

// img0: first batch image of this launch (sub-batched stages; 0 otherwise), nb: images it covers
static int launch_conv(sbb_model* m, Op& op, int t0, int nb, bool crop, const HeadParams* hp, cudaStream_t st, int img0 = 0) {
  const ConvParams& p0 = op.variants[0];
  LaunchArgs a{};
  a.variants = op.d_variants;
  a.n_variants = (int)op.variants.size();
  if (op.flat) { a.GW = (int)(op.per_img_px * nb); a.GH = 1; a.NIMG = 1; a.x_off = (int)(op.per_img_px * img0); }
  else { a.GW = op.GW; a.GH = op.GH; a.NIMG = nb; a.img0 = img0; }
  a.tiles_x = (a.GW + p0.BW - 1) / p0.BW;
  a.tiles_y = (a.GH + p0.BH - 1) / p0.BH;
  a.BI = p0.BI;
  a.total_work = a.tiles_x * a.tiles_y * ((a.NIMG + p0.BI - 1) / p0.BI) * p0.n_tiles_n;
  a.head = *hp;
  a.debug = m->debug;
#ifdef SBB_X_DIRECT_STORE
  a.direct_store = (op.pair && m->direct_store) ? 1 : 0;
#endif
  a.BW = p0.BW; a.BH = p0.BH; a.n_tiles_n = p0.n_tiles_n; a.has_res = p0.res != nullptr;
  if (m->backend == SBB_BACKEND_SIMT) {
    const int64_t M = (int64_t)a.GW * a.GH * a.NIMG;
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)(p0.Cout / 32));
    for (int v = 0; v < a.n_variants; ++v) {
      if (op.head) conv_simt_kernel<true><<<grid, 128, 0, st>>>(op.d_variants + v, a);
      else conv_simt_kernel<false><<<grid, 128, 0, st>>>(op.d_variants + v, a);
      CU_TRY(cudaGetLastError());
      m->launches++;
    }
    return SBB_OK;
  }
  if (op.dec_level > 0) TRY(get_worklist(m, op, t0, nb, crop, st, &a.worklist, &a.total_work));
  if (op.chain) {
    TRY(get_chain_list(m, op, nb, st, &a.worklist, &a.total_work));
    CU_TRY(cudaMemsetAsync(op.chain_flags, 0, (size_t)op.chain_flags_n * sizeof(uint32_t), st));
    a.chain_flags = op.chain_flags;
    a.chain_need = 2 * op.variants[0].n_tiles_n;   // two epilogue groups (store threads) per CTA and N tile
  }
  const bool roles = (m->debug & 16) != 0;
  if (roles) {
    if (!m->role_buf) TRY(dev_alloc(m, (void**)&m->role_buf, (size_t)m->num_sms * 16 * 4));
    CU_TRY(cudaMemsetAsync(m->role_buf, 0, (size_t)m->num_sms * 16 * 4, st));
    a.role_cycles = m->role_buf;
  }
  const bool split = m->planes == 2;
  int rc = SBB_ERR_UNSUPPORTED;
  if (op.pair) {
    rc = op.head ? launch_pair<true, 128>(m, a, st)
                 : (op.BN == 64 ? (op.resb ? launch_pair<false, 64, true>(m, a, st) : launch_pair<false, 64>(m, a, st))
                                : launch_pair<false, 128>(m, a, st));
#ifdef SBB_X_ROLES
    if (rc == SBB_OK && roles && a.total_work > 0) rc = report_role_cycles(m, op, m->num_sms, st, /*pair=*/true);
#endif
    return rc;
  }
  if (op.head && op.BN == 128) rc = split ? launch_tc<128, true, true>(m, a, st) : launch_tc<128, false, true>(m, a, st);
  else if (op.head) rc = split ? launch_tc<32, true, true>(m, a, st) : launch_tc<32, false, true>(m, a, st);
  else switch (op.BN) {
    case 128: rc = split ? launch_tc<128, true, false>(m, a, st) : launch_tc<128, false, false>(m, a, st); break;
    case 64: rc = split ? launch_tc<64, true, false>(m, a, st) : launch_tc<64, false, false>(m, a, st); break;
    case 32: rc = split ? launch_tc<32, true, false>(m, a, st) : launch_tc<32, false, false>(m, a, st); break;
    default: return fail(SBB_ERR_UNSUPPORTED, "BN %d", op.BN);
  }
  if (rc == SBB_OK && roles && a.total_work > 0) rc = report_role_cycles(m, op, std::min(a.total_work, m->num_sms), st);
  return rc;
}

// One forward over nb tiles (page tiles [t0, t0+nb) when `crop`: decoder work outside the region
// the stitch keeps is skipped).  `hp` carries the input source and the output sinks.
static int forward(sbb_model* m, int t0, int nb, bool crop, const HeadParams& hp, cudaStream_t st) {
  const int H1 = m->f1.H, W1 = m->f1.W;
  cudaEvent_t* pev = m->part_profiling ? m->part_ev[m->part_count % sbb_model::kPartSlots] : nullptr;
  if (pev) { CU_TRY(cudaEventRecord(pev[0], st)); m->part_stream = st; }
  for (size_t oi = 0; oi < m->ops.size(); ++oi) {
    Op& op = m->ops[oi];
    if (pev && op.name == "dec_v5") CU_TRY(cudaEventRecord(pev[1], st));
    // A ResNet stage listed in SBB_SUBBATCH runs its launches over the batch in parts: all launches of the stage for
    // the first images, then for the next ones -- the stage's working set (block input/output + the 64..512-channel
    // intermediates) then fits the 126 MB L2 and the 1x1 convs stop being HBM-bound.  Parts are multiples of the
    // stage's images-per-tile, so no M tile straddles two parts.
    if (op.kind == OP_CONV && op.sub_parts > 1 && nb >= 2 * op.sub_align) {
      size_t oe = oi;
      while (oe < m->ops.size() && m->ops[oe].kind == OP_CONV && m->ops[oe].sub_stage == op.sub_stage) ++oe;
      const int parts = op.sub_parts;
      int per = (nb + parts - 1) / parts;
      per = (per + op.sub_align - 1) / op.sub_align * op.sub_align;
      if (m->profiling) CU_TRY(cudaEventRecord(op.ev0, st));
      for (int i0 = 0; i0 < nb; i0 += per)
        for (size_t k = oi; k < oe; ++k) TRY(launch_conv(m, m->ops[k], t0, std::min(per, nb - i0), crop, &hp, st, i0));
      if (m->profiling) {  // the stage's time is booked on its first launch, the others read 0
        CU_TRY(cudaEventRecord(op.ev1, st));
        for (size_t k = oi + 1; k < oe; ++k) { CU_TRY(cudaEventRecord(m->ops[k].ev0, st)); CU_TRY(cudaEventRecord(m->ops[k].ev1, st)); }
      }
      oi = oe - 1;
      continue;
    }
    if (m->profiling) CU_TRY(cudaEventRecord(op.ev0, st));
    switch (op.kind) {
      case OP_STEM_PAD: {
        StemParams s{};
        s.page = hp.page; s.page_row_stride = hp.page_row_stride; s.tiles = hp.tiles; s.tile_org = hp.tile_org;
        s.mode = hp.mode; s.TH = m->tile_h; s.TW = m->tile_w; s.nimg = nb;
        s.PH = m->PH; s.pitch = m->pitch; s.xp = m->xp;
        const int64_t total = (int64_t)nb * m->PH * m->pitch;
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)m->num_sms * 16);
        stem_pad_kernel<<<blocks, 256, 0, st>>>(s);
        CU_TRY(cudaGetLastError());
        m->launches++;
        break;
      }
      case OP_POOL: {
        PoolParams q{};
        q.in = m->f1.d; q.out = op.pool_out; q.scale = m->bn1_scale; q.shift = m->bn1_shift;
        q.nimg = nb; q.H1 = H1; q.W1 = W1; q.H2 = (H1 - 3) / 2 + 1; q.W2 = (W1 - 3) / 2 + 1; q.planes = m->planes;
        const int64_t total = (int64_t)nb * q.H2 * q.W2 * 8;
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)m->num_sms * 16);
        stem_bn_relu_maxpool_kernel<<<blocks, 256, 0, st>>>(q);
        CU_TRY(cudaGetLastError());
        m->launches++;
        break;
      }
      case OP_CONV:
        TRY(launch_conv(m, op, t0, nb, crop, &hp, st));
        break;
    }
    if (m->profiling) CU_TRY(cudaEventRecord(op.ev1, st));
  }
  if (pev) { CU_TRY(cudaEventRecord(pev[2], st)); m->part_count++; }
  m->last_nb = nb;
  return SBB_OK;
}

static int finish_profiling(sbb_model* m, cudaStream_t st) {
  if (!m->profiling) return SBB_OK;   // (the encoder / decoder split is read out later, without a sync per forward)
  CU_TRY(cudaStreamSynchronize(st));
  for (Op& op : m->ops) {
    float ms = 0.0f;
    CU_TRY(cudaEventElapsedTime(&ms, op.ev0, op.ev1));
    op.ms += ms;
  }
  return SBB_OK;
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" int sbb_abi_version(void) { return SBB_ABI_VERSION; }
extern "C" const char* sbb_last_error(void) { return g_err.c_str(); }

extern "C" void sbb_model_destroy(sbb_model* m) {
  if (!m) return;
  DeviceScope scope(m->device);
  cudaDeviceSynchronize();
  for (Op& op : m->ops) {
    if (op.ev0) cudaEventDestroy(op.ev0);
    if (op.ev1) cudaEventDestroy(op.ev1);
  }
  for (void* p : m->allocs) cudaFree(p);
  for (auto& slot : m->part_ev)
    for (cudaEvent_t e : slot) if (e) cudaEventDestroy(e);
  if (m->stage) cudaFreeHost(m->stage);
  if (m->stage_ev) cudaEventDestroy(m->stage_ev);
  if (m->chain_ev) cudaEventDestroy(m->chain_ev);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
}

extern "C" int sbb_model_create(const sbb_model_desc* d, sbb_model** out) {
  if (!d || !out) return fail(SBB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->tile_h <= 0 || d->tile_w <= 0 || d->tile_h % 32 || d->tile_w % 32)
    return fail(SBB_ERR_UNSUPPORTED, "tile size %dx%d must be a positive multiple of 32", d->tile_h, d->tile_w);
  if (d->n_classes < 1 || d->n_classes > 8) return fail(SBB_ERR_UNSUPPORTED, "n_classes %d not in [1,8]", d->n_classes);
  if (d->precision != SBB_PREC_FP16X3 && d->precision != SBB_PREC_FP16) return fail(SBB_ERR_INVALID, "bad precision");
  if (d->backend != SBB_BACKEND_TCGEN05 && d->backend != SBB_BACKEND_SIMT) return fail(SBB_ERR_INVALID, "bad backend");
  if (!d->weights) return fail(SBB_ERR_INVALID, "null weights");
  int nc = 0, blob_th = 0, blob_tw = 0;
  std::vector<Rec> recs;
  TRY(parse_blob(d->weights, d->weights_nbytes, &nc, &blob_th, &blob_tw, &recs));
  if (nc != d->n_classes) return fail(SBB_ERR_INVALID, "blob has %d classes, desc says %d", nc, d->n_classes);
  if ((blob_th || blob_tw) && (blob_th != d->tile_h || blob_tw != d->tile_w))
    return fail(SBB_ERR_INVALID, "blob was converted from a %dx%d model, desc says %dx%d: the tile grid and margins would "
                "differ from the reference's (main.py:227-233)", blob_th, blob_tw, d->tile_h, d->tile_w);
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (d->device < 0 || d->device >= ndev) return fail(SBB_ERR_INVALID, "device %d of %d", d->device, ndev);
  ENTER_DEVICE(d->device);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, d->device));
  if (d->backend == SBB_BACKEND_TCGEN05 && prop.major != 10)
    return fail(SBB_ERR_UNSUPPORTED, "tcgen05 backend needs an sm_100 GPU, device %d is sm_%d%d", d->device, prop.major,
                prop.minor);
  std::unique_ptr<sbb_model> m(new sbb_model());
  m->tile_h = d->tile_h; m->tile_w = d->tile_w; m->n_classes = d->n_classes;
  m->precision = d->precision; m->backend = d->backend; m->device = d->device;
  m->NB = d->max_batch > 0 ? d->max_batch : 48;
  m->planes = d->precision == SBB_PREC_FP16X3 ? 2 : 1;
  m->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("SBB_WIN_CHUNKS")) m->win_chunks = std::max(1, atoi(e));  // tuning knobs
  if (const char* e = getenv("SBB_WIDE_N")) m->wide_n = atoi(e) != 0;
  if (const char* e = getenv("SBB_CROP")) m->crop = atoi(e) != 0;
  if (const char* e = getenv("SBB_DEC_RECT")) m->dec_rect = atoi(e) != 0;
  if (const char* e = getenv("SBB_DEC5_MERGED")) m->dec5_merged = atoi(e) != 0;
  if (const char* e = getenv("SBB_IMG_BOXES")) m->img_boxes = atoi(e) != 0;
  if (const char* e = getenv("SBB_PAIR")) m->pair_mode = atoi(e);
  if (const char* e = getenv("SBB_PAIR_MIN_CHUNKS")) m->pair_min_chunks = std::max(1, atoi(e));
  if (const char* e = getenv("SBB_PAIR_HEAD")) m->pair_head = atoi(e) != 0;
  if (const char* e = getenv("SBB_PAIR64")) m->pair64 = atoi(e);
  if (const char* e = getenv("SBB_PAIR_RESB")) m->pair_resb = atoi(e) != 0;
  if (const char* e = getenv("SBB_CHAIN")) m->chain = atoi(e) != 0;
  if (const char* e = getenv("SBB_DIRECT_STORE")) m->direct_store = atoi(e) != 0;
  if (const char* e = getenv("SBB_DEC4_MERGED")) m->dec4_merged = atoi(e) != 0;
  if (const char* e = getenv("SBB_SUBBATCH")) {
    for (const char* p = e; *p;) {
      int stage = 0, parts = 1;
      if (sscanf(p, "%d:%d", &stage, &parts) == 2 && stage >= 2 && stage <= 5 && parts >= 1) m->sub_parts[stage] = parts;
      const char* q = strchr(p, ',');
      p = q ? q + 1 : p + strlen(p);
    }
  }
  if (const char* e = getenv("SBB_RES_IN_MMA")) m->res_in_mma = atoi(e) != 0;
  if (const char* e = getenv("SBB_DEBUG")) m->debug = atoi(e);
  if (d->backend == SBB_BACKEND_TCGEN05) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(SBB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    m->encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  CU_TRY(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  int rc = build_plan(m.get(), recs);
  if (rc != SBB_OK) { sbb_model_destroy(m.release()); return rc; }
  {
    int slots = 8;
    if (const char* e = getenv("SBB_GEOM_CACHE")) slots = std::max(1, atoi(e));
    m->geoms.resize(slots);
    // sbb_predict_full: one "tile" that owns every pixel of a tile-sized image
    Geom& g = m->full_geom;
    const int H = m->tile_h, W = m->tile_w;
    std::vector<int32_t> org = {0, 0, 0, 0};
    std::vector<int16_t> ox(W, 0), oy(H, 0);
    rc = ensure(m.get(), &g.d_tile_org, &g.cap_tiles, 4);
    if (rc == SBB_OK) rc = ensure(m.get(), &g.d_owner_x, &g.cap_x, (size_t)W);
    if (rc == SBB_OK) rc = ensure(m.get(), &g.d_owner_y, &g.cap_y, (size_t)H);
    if (rc != SBB_OK) { sbb_model_destroy(m.release()); return rc; }
    if (cudaMemcpy(g.d_tile_org, org.data(), 16, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(g.d_owner_x, ox.data(), ox.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(g.d_owner_y, oy.data(), oy.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
      sbb_model_destroy(m.release());
      return fail(SBB_ERR_CUDA, "uploading the whole-image owner tables failed");
    }
    g.H = H; g.W = W; g.margin = 0; g.nxf = g.nyf = 1;
    g.keep.assign(1, Rect{0, 0, W - 1, H - 1});
  }
  for (Op& op : m->ops) {
    if (cudaEventCreate(&op.ev0) != cudaSuccess || cudaEventCreate(&op.ev1) != cudaSuccess) {
      sbb_model_destroy(m.release());
      return fail(SBB_ERR_CUDA, "cudaEventCreate failed");
    }
  }
  CU_TRY(cudaDeviceSynchronize());
  *out = m.release();
  return SBB_OK;
}

extern "C" int sbb_model_shape(const sbb_model* m, int32_t* th, int32_t* tw, int32_t* nc) {
  if (!m) return fail(SBB_ERR_INVALID, "null model");
  if (th) *th = m->tile_h;
  if (tw) *tw = m->tile_w;
  if (nc) *nc = m->n_classes;
  return SBB_OK;
}

// Fills geometry slot `g` for an H x W page: tile origins, owner tables (device, through the pinned arena)
// and per tile the box of pixels the stitch keeps (host).  No host synchronisation.
static int fill_geom(sbb_model* m, Geom* g, int H, int W, int margin, int n_pages, cudaStream_t st) {
  int nxf = 0, nyf = 0;
  TRY(sbb_compute_tile_grid(H, W, m->tile_h, m->tile_w, margin, &nxf, &nyf, nullptr, 0, nullptr, nullptr));
  const int ntiles1 = nxf * nyf, ntiles = ntiles1 * n_pages;
  std::vector<int32_t> org(4 * (size_t)ntiles);
  std::vector<int16_t> ox(W), oy((size_t)H * n_pages);
  TRY(sbb_compute_tile_grid(H, W, m->tile_h, m->tile_w, margin, &nxf, &nyf, org.data(), ntiles1, ox.data(), oy.data()));
  // stacked pages: page p occupies rows [p*H, (p+1)*H) of one tall buffer; its tiles are page 0's shifted down,
  // the row-owner table repeats (tile indices (i, j) are per page)
  for (int p = 1; p < n_pages; ++p) {
    for (int t = 0; t < ntiles1; ++t) {
      int32_t* d = &org[4 * ((size_t)p * ntiles1 + t)];
      const int32_t* s0 = &org[4 * (size_t)t];
      d[0] = s0[0]; d[1] = s0[1] + p * H; d[2] = s0[2]; d[3] = s0[3];
    }
    std::copy(oy.begin(), oy.begin() + H, oy.begin() + (size_t)p * H);
  }
  TRY(ensure(m, &g->d_tile_org, &g->cap_tiles, 4 * (size_t)ntiles));
  TRY(ensure(m, &g->d_owner_x, &g->cap_x, (size_t)W));
  TRY(ensure(m, &g->d_owner_y, &g->cap_y, (size_t)H * n_pages));
  // one staged block at a time (alloc -> fill -> upload), so that an arena wrap never finds a block that is
  // filled but not yet submitted
  auto put = [&](void* dst, const void* src, size_t bytes) -> int {
    void* h = nullptr;
    TRY(stage_alloc(m, bytes, st, &h));
    memcpy(h, src, bytes);
    return stage_upload(m, dst, h, bytes, st);
  };
  TRY(put(g->d_tile_org, org.data(), org.size() * 4));
  TRY(put(g->d_owner_x, ox.data(), ox.size() * 2));
  TRY(put(g->d_owner_y, oy.data(), oy.size() * 2));
  g->keep = keep_boxes(org, ox, oy, ntiles, m->tile_h, m->tile_w);
  g->H = H; g->W = W; g->margin = margin; g->n_pages = n_pages; g->nxf = nxf; g->nyf = nyf;
  g->serial = ++m->geom_serial;
  return SBB_OK;
}

// LRU lookup of the page geometry (H, W, margin): a hit touches neither the device nor the stream.
static int get_geom(sbb_model* m, int H, int W, int margin, int n_pages, cudaStream_t st, const Geom** out) {
  Geom* lru = nullptr;
  for (Geom& g : m->geoms) {
    if (g.serial != 0 && g.H == H && g.W == W && g.margin == margin && g.n_pages == n_pages) {
      g.last_use = ++m->use_clock;
      m->geom_hits++;
      *out = &g;
      return SBB_OK;
    }
    if (!lru || g.last_use < lru->last_use) lru = &g;
  }
  m->geom_misses++;
  TRY(fill_geom(m, lru, H, W, margin, n_pages, st));
  lru->last_use = ++m->use_clock;
  *out = lru;
  return SBB_OK;
}

// tiles [tile_first, tile_first + tile_count) of the page grid (tile_count < 0: all); keep_labels: do not clear
// the label map first (several ranks stitch disjoint tile ranges into ONE map, see sbb_predict_page_tile_range)
// n_pages > 1: that many same-size pages stacked vertically in `bgr` / `labels` (H = one page's height)
static int predict_page_impl(sbb_model* m, const uint8_t* bgr, int32_t H, int32_t W, int64_t row_stride,
                             int32_t margin, uint8_t* labels, int64_t out_row_stride, int32_t tile_first,
                             int32_t tile_count, bool keep_labels, int32_t memkind, void* stream, int32_t n_pages = 1) {
  if (!m || !bgr || !labels) return fail(SBB_ERR_INVALID, "null argument");
  if (n_pages < 1 || (int64_t)n_pages * H > 32000000) return fail(SBB_ERR_INVALID, "bad page count");
  if (row_stride < 3 * (int64_t)W || out_row_stride < W) return fail(SBB_ERR_INVALID, "row stride too small");
  ENTER_DEVICE(m->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : m->own_stream;
  int nxf = 0, nyf = 0;
  TRY(sbb_compute_tile_grid(H, W, m->tile_h, m->tile_w, margin, &nxf, &nyf, nullptr, 0, nullptr, nullptr));
  const int ntiles = nxf * nyf * n_pages;
  const int HS = H * n_pages;   // rows of the stacked buffers
  if (tile_count < 0) { tile_first = 0; tile_count = ntiles; }
  if (tile_first < 0 || tile_first + tile_count > ntiles)
    return fail(SBB_ERR_INVALID, "tile range [%d, %d) outside the %d tiles of the page", tile_first, tile_first + tile_count, ntiles);
  TRY(chain_begin(m, st));
  const Geom* g = nullptr;
  TRY(get_geom(m, H, W, margin, n_pages, st, &g));
  m->cur = g;
  const uint8_t* d_in = bgr;
  uint8_t* d_out = labels;
  int64_t in_stride = row_stride, o_stride = out_row_stride;
  if (memkind == SBB_MEM_HOST) {
    TRY(ensure(m, &m->d_page, &m->page_cap, (size_t)HS * W * 3));
    TRY(ensure(m, &m->d_labels, &m->labels_cap, (size_t)HS * W));
    CU_TRY(cudaMemcpy2DAsync(m->d_page, (size_t)W * 3, bgr, (size_t)row_stride, (size_t)W * 3, HS, cudaMemcpyHostToDevice, st));
    d_in = m->d_page; d_out = m->d_labels; in_stride = (int64_t)W * 3; o_stride = W;
  }
  if (!keep_labels) CU_TRY(cudaMemset2DAsync(d_out, (size_t)o_stride, 0, (size_t)W, HS, st));
  m->launches = 0;
  for (Op& op : m->ops) op.ms = 0.0f;
  const int t_end = tile_first + tile_count;
  for (int t0 = tile_first; t0 < t_end; t0 += m->NB) {
    const int nb = std::min(m->NB, t_end - t0);
    HeadParams hp{};
    hp.page = d_in; hp.page_row_stride = in_stride; hp.tile_org = g->d_tile_org + 4 * (size_t)t0;
    hp.owner_x = g->d_owner_x; hp.owner_y = g->d_owner_y;
    hp.labels = d_out; hp.labels_row_stride = o_stride;
    hp.w_cls = m->w_cls; hp.b_cls = m->b_cls;
    hp.n_classes = m->n_classes; hp.TH = m->tile_h; hp.TW = m->tile_w; hp.mode = 0;
    TRY(forward(m, t0, nb, m->crop != 0, hp, st));
    TRY(finish_profiling(m, st));
  }
  if (memkind == SBB_MEM_HOST)
    CU_TRY(cudaMemcpy2DAsync(labels, (size_t)out_row_stride, m->d_labels, (size_t)W, (size_t)W, HS, cudaMemcpyDeviceToHost, st));
  TRY(chain_end(m, st));
  if (memkind == SBB_MEM_HOST) CU_TRY(cudaStreamSynchronize(st));
  return SBB_OK;
}

extern "C" int sbb_predict_page_tiled(sbb_model* m, const uint8_t* bgr, int32_t H, int32_t W, int64_t row_stride,
                                      int32_t margin, uint8_t* labels, int64_t out_row_stride, int32_t memkind,
                                      void* stream) {
  return predict_page_impl(m, bgr, H, W, row_stride, margin, labels, out_row_stride, 0, -1, false, memkind, stream);
}

// Throughput form for many pages: n_pages pages of the SAME size stacked vertically in one buffer (page p = rows
// [p*H, (p+1)*H) of `bgr_stack` / `labels_stack`) go through the network as ONE batch -- per-launch costs (pipeline
// fill, drain, wave tails: ~15 us on each of the 58 launches) are paid once for all of them when the handle's
// max_batch covers their tiles.  Each page is tiled and stitched exactly as by sbb_predict_page_tiled.
extern "C" int sbb_predict_pages_stacked(sbb_model* m, const uint8_t* bgr_stack, int32_t n_pages, int32_t H, int32_t W,
                                         int64_t row_stride, int32_t margin, uint8_t* labels_stack, int64_t out_row_stride,
                                         int32_t memkind, void* stream) {
  return predict_page_impl(m, bgr_stack, H, W, row_stride, margin, labels_stack, out_row_stride, 0, -1, false, memkind,
                           stream, n_pages);
}

extern "C" int sbb_predict_page_tile_range(sbb_model* m, const uint8_t* bgr, int32_t H, int32_t W, int64_t row_stride,
                                           int32_t margin, uint8_t* labels, int64_t out_row_stride, int32_t tile_first,
                                           int32_t tile_count, int32_t keep_labels, void* stream) {
  if (tile_count < 0) return fail(SBB_ERR_INVALID, "negative tile count");
  return predict_page_impl(m, bgr, H, W, row_stride, margin, labels, out_row_stride, tile_first, tile_count,
                           keep_labels != 0, SBB_MEM_DEVICE, stream);
}

// ---- peer-visible device buffers (CUDA IPC): the label map one rank owns and every rank's head epilogue
// stores into over NVLink.  Plain cudaMalloc memory (IPC handles name whole allocations).
extern "C" int sbb_peer_alloc(int32_t device, size_t nbytes, void** ptr, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle || nbytes == 0) return fail(SBB_ERR_INVALID, "bad argument");
  ENTER_DEVICE(device);
  CU_TRY(cudaMalloc(ptr, nbytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return fail(SBB_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
  memcpy(handle, &h, 64);
  return SBB_OK;
}
extern "C" int sbb_peer_open(int32_t device, const uint8_t handle[64], void** ptr) {
  if (!ptr || !handle) return fail(SBB_ERR_INVALID, "bad argument");
  ENTER_DEVICE(device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SBB_OK;
}
extern "C" int sbb_peer_close(void* ptr) {
  if (ptr) CU_TRY(cudaIpcCloseMemHandle(ptr));
  return SBB_OK;
}
extern "C" int sbb_peer_free(void* ptr) {
  if (ptr) CU_TRY(cudaFree(ptr));
  return SBB_OK;
}

extern "C" int sbb_predict_tiles(sbb_model* m, const float* tiles, int32_t n, uint8_t* labels, float* probs,
                                 float* logits, int32_t memkind, void* stream) {
  if (!m || !tiles || n <= 0) return fail(SBB_ERR_INVALID, "bad argument");
  ENTER_DEVICE(m->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : m->own_stream;
  TRY(chain_begin(m, st));
  const size_t px = (size_t)m->tile_h * m->tile_w;
  const int C = m->n_classes;
  if (memkind == SBB_MEM_HOST) {
    if (!m->d_tiles) {
      TRY(dev_alloc(m, (void**)&m->d_tiles, (size_t)m->NB * px * 3 * 4));
      TRY(dev_alloc(m, (void**)&m->d_probs, (size_t)m->NB * px * C * 4));
      TRY(dev_alloc(m, (void**)&m->d_logits, (size_t)m->NB * px * C * 4));
      TRY(dev_alloc(m, (void**)&m->d_tlabels, (size_t)m->NB * px));
    }
  }
  m->launches = 0;
  for (Op& op : m->ops) op.ms = 0.0f;
  for (int t0 = 0; t0 < n; t0 += m->NB) {
    const int nb = std::min(m->NB, n - t0);
    HeadParams hp{};
    hp.mode = 1;
    hp.w_cls = m->w_cls; hp.b_cls = m->b_cls;
    hp.n_classes = C; hp.TH = m->tile_h; hp.TW = m->tile_w;
    if (memkind == SBB_MEM_HOST) {
      CU_TRY(cudaMemcpyAsync(m->d_tiles, tiles + (size_t)t0 * px * 3, (size_t)nb * px * 3 * 4, cudaMemcpyHostToDevice, st));
      hp.tiles = m->d_tiles;
      hp.labels = labels ? m->d_tlabels : nullptr;
      hp.probs = probs ? m->d_probs : nullptr;
      hp.logits = logits ? m->d_logits : nullptr;
    } else {
      hp.tiles = tiles + (size_t)t0 * px * 3;
      hp.labels = labels ? labels + (size_t)t0 * px : nullptr;
      hp.probs = probs ? probs + (size_t)t0 * px * C : nullptr;
      hp.logits = logits ? logits + (size_t)t0 * px * C : nullptr;
    }
    hp.labels_row_stride = m->tile_w;
    TRY(forward(m, 0, nb, false, hp, st));
    TRY(finish_profiling(m, st));
    if (memkind == SBB_MEM_HOST) {
      if (labels) CU_TRY(cudaMemcpyAsync(labels + (size_t)t0 * px, m->d_tlabels, (size_t)nb * px, cudaMemcpyDeviceToHost, st));
      if (probs) CU_TRY(cudaMemcpyAsync(probs + (size_t)t0 * px * C, m->d_probs, (size_t)nb * px * C * 4, cudaMemcpyDeviceToHost, st));
      if (logits) CU_TRY(cudaMemcpyAsync(logits + (size_t)t0 * px * C, m->d_logits, (size_t)nb * px * C * 4, cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaStreamSynchronize(st));
    }
  }
  TRY(chain_end(m, st));
  return SBB_OK;
}

extern "C" int sbb_predict_full(sbb_model* m, const uint8_t* bgr_tile, uint8_t* labels, int32_t memkind, void* stream) {
  if (!m) return fail(SBB_ERR_INVALID, "null model");
  // a tile-sized page: the tiler would make 2x2 clamped tiles; run ONE tile whose owner tables cover
  // everything instead (do_prediction(patches=False) has no margin crop).  Those tables were uploaded once
  // at sbb_model_create (full_geom).
  if (!bgr_tile || !labels) return fail(SBB_ERR_INVALID, "null argument");
  ENTER_DEVICE(m->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : m->own_stream;
  const int H = m->tile_h, W = m->tile_w;
  TRY(chain_begin(m, st));
  const Geom& g = m->full_geom;
  m->cur = &g;
  const uint8_t* d_in = bgr_tile;
  uint8_t* d_out = labels;
  if (memkind == SBB_MEM_HOST) {
    TRY(ensure(m, &m->d_page, &m->page_cap, (size_t)H * W * 3));
    TRY(ensure(m, &m->d_labels, &m->labels_cap, (size_t)H * W));
    CU_TRY(cudaMemcpyAsync(m->d_page, bgr_tile, (size_t)H * W * 3, cudaMemcpyHostToDevice, st));
    d_in = m->d_page; d_out = m->d_labels;
  }
  m->launches = 0;
  for (Op& op : m->ops) op.ms = 0.0f;
  HeadParams hp{};
  hp.page = d_in; hp.page_row_stride = (int64_t)W * 3; hp.tile_org = g.d_tile_org;
  hp.owner_x = g.d_owner_x; hp.owner_y = g.d_owner_y;
  hp.labels = d_out; hp.labels_row_stride = W;
  hp.w_cls = m->w_cls; hp.b_cls = m->b_cls;
  hp.n_classes = m->n_classes; hp.TH = H; hp.TW = W; hp.mode = 0;
  TRY(forward(m, 0, 1, false, hp, st));
  TRY(finish_profiling(m, st));
  if (memkind == SBB_MEM_HOST) CU_TRY(cudaMemcpyAsync(labels, m->d_labels, (size_t)H * W, cudaMemcpyDeviceToHost, st));
  TRY(chain_end(m, st));
  if (memkind == SBB_MEM_HOST) CU_TRY(cudaStreamSynchronize(st));
  return SBB_OK;
}

extern "C" int sbb_model_num_activations(const sbb_model* m) { return m ? (int)m->acts.size() : 0; }
extern "C" int sbb_model_activation_info(const sbb_model* m, int32_t i, const char** name, int32_t* h, int32_t* w,
                                         int32_t* c) {
  if (!m || i < 0 || i >= (int)m->acts.size()) return fail(SBB_ERR_INVALID, "bad activation index");
  const ActInfo& a = m->acts[i];
  if (name) *name = a.name.c_str();
  if (h) *h = a.t.H;
  if (w) *w = a.t.W;
  if (c) *c = a.t.C;
  return SBB_OK;
}
extern "C" int sbb_model_read_activation(sbb_model* m, int32_t i, int32_t tile, float* out) {
  if (!m || i < 0 || i >= (int)m->acts.size() || !out) return fail(SBB_ERR_INVALID, "bad argument");
  if (tile < 0 || tile >= m->last_nb) return fail(SBB_ERR_INVALID, "tile %d not in the last batch of %d", tile, m->last_nb);
  ENTER_DEVICE(m->device);
  CU_TRY(cudaDeviceSynchronize());
  const Tensor& t = m->acts[i].t;
  const size_t npx = (size_t)t.H * t.W;
  std::vector<__half> buf(npx * t.pix());
  CU_TRY(cudaMemcpy(buf.data(), t.d + (size_t)tile * npx * t.pix(), buf.size() * sizeof(__half), cudaMemcpyDeviceToHost));
  for (size_t p = 0; p < npx; ++p)
    for (int c = 0; c < t.C; ++c) {
      float v = __half2float(buf[p * t.pix() + c]);
      if (t.planes == 2) v += __half2float(buf[p * t.pix() + t.C + c]);
      out[p * t.C + c] = v;
    }
  return SBB_OK;
}
// Precision plan (fp16x3 handles only): the launches named in `layers` (comma separated layer names as
// sbb_model_layer_time reports them; "" = none) read only the hi plane of their input activations -- A_lo is
// neither loaded nor multiplied, 2 MMA units per K step instead of 3 -- every other launch keeps the full
// hi/lo scheme.  Whether a layer tolerates that depends on the weights: precision.plan_layers measures it.
extern "C" int sbb_model_set_precision_plan(sbb_model* m, const char* layers) {
  if (!m || !layers) return fail(SBB_ERR_INVALID, "null argument");
  if (m->planes != 2 || m->backend != SBB_BACKEND_TCGEN05) return fail(SBB_ERR_UNSUPPORTED, "precision plans apply to the fp16x3 tcgen05 path");
  ENTER_DEVICE(m->device);
  std::vector<std::string> names;
  for (const char* p = layers; *p;) {
    const char* q = strchr(p, ',');
    std::string n = q ? std::string(p, q) : std::string(p);
    if (!n.empty()) names.push_back(n);
    p = q ? q + 1 : p + strlen(p);
  }
  for (const std::string& n : names) {
    bool found = false;
    for (const Op& op : m->ops) found = found || (op.kind == OP_CONV && op.name == n);
    if (!found) return fail(SBB_ERR_INVALID, "precision plan: no conv launch named %s", n.c_str());
  }
  CU_TRY(cudaDeviceSynchronize());   // no launch in flight reads the descriptions while they change
  for (Op& op : m->ops) {
    if (op.kind != OP_CONV) continue;
    const int v = std::find(names.begin(), names.end(), op.name) != names.end() ? 1 : 0;
    bool changed = false;
    for (ConvParams& p : op.variants) { changed = changed || p.a_hi_only != v; p.a_hi_only = v; }
    if (changed)
      CU_TRY(cudaMemcpy(op.d_variants, op.variants.data(), op.variants.size() * sizeof(ConvParams), cudaMemcpyHostToDevice));
  }
  return SBB_OK;
}

extern "C" int sbb_model_geom_cache_stats(const sbb_model* m, int64_t* hits, int64_t* misses) {
  if (!m) return fail(SBB_ERR_INVALID, "null model");
  if (hits) *hits = m->geom_hits;
  if (misses) *misses = m->geom_misses;
  return SBB_OK;
}
extern "C" int64_t sbb_model_last_launch_count(const sbb_model* m) { return m ? m->launches : 0; }
extern "C" int sbb_model_part_times(sbb_model* m, float* encoder_ms, float* decoder_ms, int32_t* forwards, int32_t reset) {
  if (!m) return fail(SBB_ERR_INVALID, "null model");
  ENTER_DEVICE(m->device);
  if (m->part_stream) CU_TRY(cudaStreamSynchronize(m->part_stream));
  const int n = std::min(m->part_count, (int)sbb_model::kPartSlots);   // older forwards were overwritten
  float enc = 0.0f, dec = 0.0f;
  for (int i = 0; i < n; ++i) {
    float a = 0.0f, b = 0.0f;
    CU_TRY(cudaEventElapsedTime(&a, m->part_ev[i][0], m->part_ev[i][1]));
    CU_TRY(cudaEventElapsedTime(&b, m->part_ev[i][1], m->part_ev[i][2]));
    enc += a; dec += b;
  }
  if (encoder_ms) *encoder_ms = enc;
  if (decoder_ms) *decoder_ms = dec;
  if (forwards) *forwards = n;
  if (reset) m->part_count = 0;
  return SBB_OK;
}
extern "C" int sbb_model_set_profiling(sbb_model* m, int32_t enable) {
  if (!m) return fail(SBB_ERR_INVALID, "null model");
  if (enable == 2) {   // coarse: encoder / decoder split from three events per forward (no per-launch gaps)
    ENTER_DEVICE(m->device);
    for (auto& slot : m->part_ev)
      for (cudaEvent_t& e : slot)
        if (!e) CU_TRY(cudaEventCreate(&e));
    m->part_profiling = true; m->profiling = false;
    m->part_count = 0;
    return SBB_OK;
  }
  m->part_profiling = false;
  m->profiling = enable != 0;
  return SBB_OK;
}
extern "C" int sbb_model_num_layers(const sbb_model* m) { return m ? (int)m->ops.size() : 0; }
extern "C" int sbb_model_layer_time(const sbb_model* m, int32_t i, const char** name, float* ms, double* flops) {
  if (!m || i < 0 || i >= (int)m->ops.size()) return fail(SBB_ERR_INVALID, "bad layer index");
  const Op& op = m->ops[i];
  if (name) *name = op.name.c_str();
  if (ms) *ms = op.ms;
  if (flops) *flops = op.flops_per_img;
  return SBB_OK;
}

// ------------------------------------------------------------------------------------------ pre/post byte ops
// the cv2 calls around the three models (prepost.cuh).  Model-independent;
// `stream` NULL = the legacy default stream.  SBB_MEM_HOST buffers are staged through stream-ordered
// device allocations and the call returns with the result in place.
#include "prepost.cuh"

namespace {
// The stream-ordered allocator trims its pool back to the OS at every synchronisation unless told otherwise
// (release threshold 0): every byte-op call then re-created its temporaries with physical allocations (3-7 ms per
// otsu / resize call under the page dispatcher instead of microseconds).  Keep freed memory in the pool.
void keep_pool_memory() {
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[dev] = true;
}

struct Scratch {  // stream-ordered temporaries, released when the call returns
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t s) : st(s) { keep_pool_memory(); }
  ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  int get(void** p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, st);
    if (e != cudaSuccess) return fail(SBB_ERR_NOMEM, "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
    ptrs.push_back(*p);
    return SBB_OK;
  }
};
int pp_blocks(int64_t total) { return (int)std::min<int64_t>((total + 255) / 256, 148 * 16); }

// Brings a host image to the device (or passes a device image through).  Returns pointer + row stride.
int pp_stage_in(Scratch& sc, const uint8_t* src, int H, int64_t row_bytes, int64_t stride, int memkind,
                const uint8_t** d, int64_t* d_stride) {
  if (memkind == SBB_MEM_DEVICE) { *d = src; *d_stride = stride; return SBB_OK; }
  void* p = nullptr;
  TRY(sc.get(&p, (size_t)H * row_bytes));
  CU_TRY(cudaMemcpy2DAsync(p, (size_t)row_bytes, src, (size_t)stride, (size_t)row_bytes, H, cudaMemcpyHostToDevice, sc.st));
  *d = (const uint8_t*)p; *d_stride = row_bytes;
  return SBB_OK;
}
int pp_stage_out(Scratch& sc, uint8_t* dst, int H, int64_t row_bytes, int64_t stride, int memkind, uint8_t** d,
                 int64_t* d_stride) {
  if (memkind == SBB_MEM_DEVICE) { *d = dst; *d_stride = stride; return SBB_OK; }
  void* p = nullptr;
  TRY(sc.get(&p, (size_t)H * row_bytes));
  *d = (uint8_t*)p; *d_stride = row_bytes;
  return SBB_OK;
}
int pp_finish(Scratch& sc, uint8_t* dst, const uint8_t* d, int H, int64_t row_bytes, int64_t stride, int memkind) {
  if (memkind == SBB_MEM_DEVICE) return SBB_OK;
  CU_TRY(cudaMemcpy2DAsync(dst, (size_t)stride, d, (size_t)row_bytes, (size_t)row_bytes, H, cudaMemcpyDeviceToHost, sc.st));
  CU_TRY(cudaStreamSynchronize(sc.st));
  return SBB_OK;
}
}  // namespace

extern "C" int sbb_resize_nearest_u8(const uint8_t* src, int32_t H, int32_t W, int32_t C, int64_t src_stride,
                                     uint8_t* dst, int32_t oh, int32_t ow, int64_t dst_stride, int32_t memkind,
                                     int32_t device, void* stream) {
  if (!src || !dst || H <= 0 || W <= 0 || oh <= 0 || ow <= 0 || C < 1 || C > 4) return fail(SBB_ERR_INVALID, "bad argument");
  if (src_stride < (int64_t)W * C || dst_stride < (int64_t)ow * C) return fail(SBB_ERR_INVALID, "row stride too small");
  ENTER_DEVICE(device);
  Scratch sc((cudaStream_t)stream);
  // OpenCV resizeNN: ifx = 1 / (dsize.width / (double)ssize.width); x_ofs[x] = min(floor(x * ifx), ssize.width - 1)
  const double ifx = 1.0 / ((double)ow / (double)W), ify = 1.0 / ((double)oh / (double)H);
  // the index tables depend on the four sizes only; a pageable upload per call would synchronise the stream, so the
  // last few geometries stay on the device (the pipeline alternates between a handful: page -> tile, tile -> page)
  struct TabEntry { int dev, H, W, oh, ow; void* d; uint64_t use; };
  static std::vector<TabEntry> cache;
  static std::mutex cache_mu;
  static uint64_t clock_ = 0;
  void* d_tab = nullptr;
  {
    std::lock_guard<std::mutex> lk(cache_mu);
    for (TabEntry& e : cache)
      if (e.dev == device && e.H == H && e.W == W && e.oh == oh && e.ow == ow) { d_tab = e.d; e.use = ++clock_; }
    if (!d_tab) {
      std::vector<int32_t> tab((size_t)oh + ow);
      for (int y = 0; y < oh; ++y) tab[y] = std::min((int)std::floor(y * ify), H - 1);
      for (int x = 0; x < ow; ++x) tab[oh + x] = std::min((int)std::floor(x * ifx), W - 1);
      if (cache.size() >= 32) {   // evict the least recently used table (synchronously: nothing may still read it)
        size_t lru = 0;
        for (size_t i = 1; i < cache.size(); ++i) if (cache[i].use < cache[lru].use) lru = i;
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(cache[lru].d);
        cache.erase(cache.begin() + lru);
      }
      CU_TRY(cudaMalloc(&d_tab, tab.size() * 4));
      CU_TRY(cudaMemcpy(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
      cache.push_back(TabEntry{device, H, W, oh, ow, d_tab, ++clock_});
    }
  }
  const uint8_t* d_src; uint8_t* d_dst; int64_t ss, ds;
  TRY(pp_stage_in(sc, src, H, (int64_t)W * C, src_stride, memkind, &d_src, &ss));
  TRY(pp_stage_out(sc, dst, oh, (int64_t)ow * C, dst_stride, memkind, &d_dst, &ds));
  resize_nearest_u8_kernel<<<pp_blocks((int64_t)oh * ow), 256, 0, sc.st>>>(d_src, ss, C, d_dst, ds, oh, ow, (const int32_t*)d_tab,
                                                                           (const int32_t*)d_tab + oh);
  CU_TRY(cudaGetLastError());
  return pp_finish(sc, dst, d_dst, oh, (int64_t)ow * C, dst_stride, memkind);
}

extern "C" int sbb_otsu_copy_u8(const uint8_t* src, int32_t H, int32_t W, int32_t C, int64_t src_stride, uint8_t* dst,
                                int64_t dst_stride, int32_t* threshold, int32_t memkind, int32_t device, void* stream) {
  if (!src || !dst || H <= 0 || W <= 0 || C < 1 || C > 4) return fail(SBB_ERR_INVALID, "bad argument");
  if (src_stride < (int64_t)W * C || dst_stride < (int64_t)W * 3) return fail(SBB_ERR_INVALID, "row stride too small");
  ENTER_DEVICE(device);
  Scratch sc((cudaStream_t)stream);
  void* d_hist = nullptr;
  TRY(sc.get(&d_hist, 257 * 4));
  CU_TRY(cudaMemsetAsync(d_hist, 0, 257 * 4, sc.st));
  const uint8_t* d_src; uint8_t* d_dst; int64_t ss, ds;
  TRY(pp_stage_in(sc, src, H, (int64_t)W * C, src_stride, memkind, &d_src, &ss));
  TRY(pp_stage_out(sc, dst, H, (int64_t)W * 3, dst_stride, memkind, &d_dst, &ds));
  int* d_thr = (int*)d_hist + 256;
  hist_ch0_kernel<<<pp_blocks((int64_t)H * W), 256, 0, sc.st>>>(d_src, ss, H, W, C, (unsigned int*)d_hist);
  otsu_threshold_kernel<<<1, 32, 0, sc.st>>>((const unsigned int*)d_hist, (int64_t)H * W, d_thr);
  otsu_apply_kernel<<<pp_blocks((int64_t)H * W), 256, 0, sc.st>>>(d_src, ss, C, d_dst, ds, H, W, d_thr);
  CU_TRY(cudaGetLastError());
  if (threshold) {  // optional: costs a synchronisation
    CU_TRY(cudaMemcpyAsync(threshold, d_thr, 4, cudaMemcpyDeviceToHost, sc.st));
    CU_TRY(cudaStreamSynchronize(sc.st));
  }
  return pp_finish(sc, dst, d_dst, H, (int64_t)W * 3, dst_stride, memkind);
}

extern "C" int sbb_morph5x5_u8(const uint8_t* src, int32_t H, int32_t W, int32_t C, int64_t src_stride, uint8_t* dst,
                               int64_t dst_stride, int32_t op, int32_t iterations, int32_t memkind, int32_t device,
                               void* stream) {
  if (!src || !dst || H <= 0 || W <= 0 || C < 1 || C > 4 || iterations < 1 || (op != 0 && op != 1))
    return fail(SBB_ERR_INVALID, "bad argument");
  if (src_stride < (int64_t)W * C || dst_stride < (int64_t)W * C) return fail(SBB_ERR_INVALID, "row stride too small");
  ENTER_DEVICE(device);
  Scratch sc((cudaStream_t)stream);
  const int r = 2 * iterations;  // n iterations of the 5x5 rectangle == one (4n+1)^2 rectangle (as OpenCV does itself)
  const uint8_t* d_src; uint8_t* d_dst; int64_t ss, ds;
  TRY(pp_stage_in(sc, src, H, (int64_t)W * C, src_stride, memkind, &d_src, &ss));
  TRY(pp_stage_out(sc, dst, H, (int64_t)W * C, dst_stride, memkind, &d_dst, &ds));
  void* tmp = nullptr;
  TRY(sc.get(&tmp, (size_t)H * W * C));
  const int blocks = pp_blocks((int64_t)H * W * C);
  if (op == 1) {
    morph_pass_kernel<true><<<blocks, 256, 0, sc.st>>>(d_src, ss, (uint8_t*)tmp, (int64_t)W * C, H, W, C, r, 0);
    morph_pass_kernel<true><<<blocks, 256, 0, sc.st>>>((const uint8_t*)tmp, (int64_t)W * C, d_dst, ds, H, W, C, r, 1);
  } else {
    morph_pass_kernel<false><<<blocks, 256, 0, sc.st>>>(d_src, ss, (uint8_t*)tmp, (int64_t)W * C, H, W, C, r, 0);
    morph_pass_kernel<false><<<blocks, 256, 0, sc.st>>>((const uint8_t*)tmp, (int64_t)W * C, d_dst, ds, H, W, C, r, 1);
  }
  CU_TRY(cudaGetLastError());
  return pp_finish(sc, dst, d_dst, H, (int64_t)W * C, dst_stride, memkind);
}

extern "C" int sbb_rotate_rowsum_u8(const uint8_t* mask, int32_t h, int32_t w, int64_t stride, int32_t S, int32_t oy,
                                    int32_t ox, const double* inv_affine, int32_t n, int32_t* profiles, int32_t memkind,
                                    int32_t device, void* stream) {
  if (!mask || !inv_affine || !profiles || h <= 0 || w <= 0 || S <= 0 || n <= 0 || n > 65535)
    return fail(SBB_ERR_INVALID, "bad argument");
  if (stride < w) return fail(SBB_ERR_INVALID, "row stride too small");
  if (oy < 0 || ox < 0 || oy + h > S || ox + w > S) return fail(SBB_ERR_INVALID, "mask does not fit the padded square");
  ENTER_DEVICE(device);
  Scratch sc((cudaStream_t)stream);
  // OpenCV's interpolateCubic (A = -0.75) at the 32 phases of INTER_TAB_SIZE, in float like initInterTab1D
  CubicTab tab;
  for (int i = 0; i < 32; ++i) {
    const float A = -0.75f, x = (float)i * (1.0f / 32);
    volatile float c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    volatile float c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
    volatile float c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    volatile float c3 = 1.f - c0 - c1 - c2;
    tab.c[i][0] = c0; tab.c[i][1] = c1; tab.c[i][2] = c2; tab.c[i][3] = c3;
  }
  void* d_m = nullptr;
  TRY(sc.get(&d_m, (size_t)n * 6 * sizeof(double)));
  CU_TRY(cudaMemcpyAsync(d_m, inv_affine, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, sc.st));
  const uint8_t* d_mask; int64_t ms;
  TRY(pp_stage_in(sc, mask, h, w, stride, memkind, &d_mask, &ms));
  int32_t* d_prof = profiles;
  if (memkind != SBB_MEM_DEVICE) {
    void* p = nullptr;
    TRY(sc.get(&p, (size_t)n * S * 4));
    d_prof = (int32_t*)p;
  }
  rotate_rowsum_kernel<<<dim3(S, n), 256, 0, sc.st>>>(d_mask, ms, h, w, S, oy, ox, (const double*)d_m, tab, d_prof);
  CU_TRY(cudaGetLastError());
  if (memkind != SBB_MEM_DEVICE)
    CU_TRY(cudaMemcpyAsync(profiles, d_prof, (size_t)n * S * 4, cudaMemcpyDeviceToHost, sc.st));
  CU_TRY(cudaStreamSynchronize(sc.st));  // inv_affine is a caller-owned host buffer: do not return while it is in flight
  return SBB_OK;
}

// ------------------------------------------------------------------------------------------ NCCL weight broadcast
// one process per GPU, the frozen weights are broadcast ONCE at init over NCCL
// (NVLink / NVSwitch) and every rank then builds its own handle from the same blob; no data-path collective.
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy a host like PyTorch already loaded is reused), so
// the library itself has no link-time dependency on it.
namespace {
struct NcclId { char bytes[128]; };  // ncclUniqueId, passed by value to ncclCommInitRank
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
int nccl_api(NcclApi** out) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("SBB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(api.lib, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(api.lib, "ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
      api.Broadcast = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(api.lib, "ncclBroadcast"));
      api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast)
    return fail(SBB_ERR_UNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2 failed; set SBB_NCCL_LIB): %s", dlerror());
  *out = &api;
  return SBB_OK;
}
#define NCCL_TRY(api, expr)                                                                                         \
  do {                                                                                                              \
    int r__ = (expr);                                                                                               \
    if (r__ != 0) return fail(SBB_ERR_CUDA, "%s failed: %s", #expr, (api)->GetErrorString ? (api)->GetErrorString(r__) : "?"); \
  } while (0)
}  // namespace

extern "C" int sbb_nccl_unique_id(uint8_t id[128]) {
  if (!id) return fail(SBB_ERR_INVALID, "null argument");
  NcclApi* api = nullptr;
  TRY(nccl_api(&api));
  NcclId u;
  NCCL_TRY(api, api->GetUniqueId(&u));
  memcpy(id, u.bytes, 128);
  return SBB_OK;
}

extern "C" int sbb_nccl_comm_create(const uint8_t id[128], int32_t n_ranks, int32_t rank, int32_t device, void** comm) {
  if (!id || !comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(SBB_ERR_INVALID, "bad argument");
  NcclApi* api = nullptr;
  TRY(nccl_api(&api));
  ENTER_DEVICE(device);
  NcclId u;
  memcpy(u.bytes, id, 128);
  NCCL_TRY(api, api->CommInitRank(comm, n_ranks, u, rank));
  return SBB_OK;
}

extern "C" int sbb_nccl_comm_destroy(void* comm) {
  if (!comm) return SBB_OK;
  NcclApi* api = nullptr;
  TRY(nccl_api(&api));
  NCCL_TRY(api, api->CommDestroy(comm));
  return SBB_OK;
}

// Broadcast of the packed weight blob from rank `root` to all ranks of `comm` (an ncclComm_t -- made by
// sbb_nccl_comm_create or handed over by the host application), staged through a device buffer on `device`.
// Every rank passes a host buffer of the SAME nbytes (the root's holds the blob, the others receive it) and then
// calls sbb_model_create on it.  Blocking.
extern "C" int sbb_model_broadcast(void* blob, size_t nbytes, int32_t root, void* comm, int32_t device, void* stream) {
  if (!blob || nbytes == 0 || !comm) return fail(SBB_ERR_INVALID, "bad argument");
  NcclApi* api = nullptr;
  TRY(nccl_api(&api));
  ENTER_DEVICE(device);
  cudaStream_t st = (cudaStream_t)stream;
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, nbytes);
  if (e != cudaSuccess) return fail(SBB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", nbytes, cudaGetErrorString(e));
  int rc = SBB_OK;
  // every rank uploads its buffer (only the root's content matters): no rank id is needed here
  if (cudaMemcpyAsync(d, blob, nbytes, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = fail(SBB_ERR_CUDA, "H2D of the blob failed");
  if (rc == SBB_OK) {
    const int r = api->Broadcast(d, d, nbytes, /*ncclUint8*/ 1, root, comm, st);
    if (r != 0) rc = fail(SBB_ERR_CUDA, "ncclBroadcast failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
  }
  if (rc == SBB_OK && cudaMemcpyAsync(blob, d, nbytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = fail(SBB_ERR_CUDA, "D2H of the blob failed");
  if (cudaStreamSynchronize(st) != cudaSuccess && rc == SBB_OK) rc = fail(SBB_ERR_CUDA, "broadcast stream failed: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
  return rc;
}
