// libav1p.so - host side: weight blob loader, op-list executor, cascade scheduler and the C ABI
// declared in include/av1p.h.  All device work is hand-written sm_100a code from the .cuh files in
// this directory; there is deliberately no CPU or library fallback - without a Blackwell device
// every entry point fails with AV1P_ENODEV / AV1P_ECUDA.
#include <climits>
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/av1p.h"
#include "aux_kernels.cuh"
#include "blob_format.h"
#include "fc_tcgen05.cuh"
#include "conv_res_tcgen05.cuh"
#include "stem_tc.cuh"
#include "stem_tma.cuh"
#include "gen_kernels.cuh"
#include "train_kernels.cuh"

using namespace av1p;

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return fail(AV1P_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

extern "C" const char* av1p_last_error(void) { return g_err.c_str(); }
extern "C" int av1p_version(void) { return 100; }

// ------------------------------------------------------------------------------ device context
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DeviceCtx {
  bool ok = false;
  int sms = 0;
  int grid_sms = 0;               // SMs a persistent grid may occupy (= sms unless av1p_set_option("grid_sms") / AV1P_GRID_SMS lowers it:
                                  // two cascades on two streams, each on half of the SMs, run side by side instead of back to back)
  EncodeTiledFn encode = nullptr;
  bool fc_pair = true;            // FC layers on CTA pairs (tcgen05 cta_group::2); AV1P_FC_PAIR=0 selects the single-CTA kernel
  bool stem_tma = true;           // frame input: TMA-staged stem (stem_tma.cuh) when the frame geometry allows a tensor map;
                                  // AV1P_STEM_TMA=0 keeps the per-thread gather kernel (stem_tc.cuh, INT_PIX)
  bool speculate = true;          // small batches: run all four stages on every block side by side (av1p_cascade_predict); AV1P_SPECULATE=0
  bool pdl = true;                // programmatic dependent launch between the kernels of an op program (AV1P_PDL=0: plain stream order)
  int cr_resid_epi = 1;           // layer1 residual convs add the identity branch in the epilogue: 1 = in place in the staging sets,
                                  // 2 = from per-thread global loads (conv_res_tcgen05.cuh, resid_epi); AV1P_CR_RESID_EPI=0: identity
                                  // MMAs through the operand ring
  bool fc_resid_epi = false;      // AV1P_FC_RESID_EPI=1: residual FC layers add the identity branch in the epilogue (aux ring)
                                  // instead of on the tensor core (FC_W_IDENT schedule entries).  Measured slower (layer2.1.conv2
                                  // 859 vs 740 us on 518 k rows): the aux ring costs two of the six operand-ring stages.
  int* watchdog_host = nullptr;   // mapped pinned memory: survives a kernel trap
  int* watchdog_dev = nullptr;
};
// One context per device ordinal: the >48 KB dynamic shared-memory opt-in (cudaFuncSetAttribute) and the SM count are per
// device, so a process that builds pipelines on cuda:0 and then on cuda:1 initialises each device on its first use.
constexpr int MAX_DEVICES = 64;
DeviceCtx g_ctxs[MAX_DEVICES];
std::mutex g_ctx_mu;

inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 0;
  return dev;
}
inline DeviceCtx& cur_ctx() { return g_ctxs[current_device()]; }
#define g_ctx (cur_ctx())

int ensure_ctx() {
  int dev = 0, cc_major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(AV1P_ENODEV, "no CUDA device: %s", cudaGetErrorString(e));
  if (dev < 0 || dev >= MAX_DEVICES) return fail(AV1P_ENODEV, "device ordinal %d outside [0, %d)", dev, MAX_DEVICES);
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceCtx& c = g_ctxs[dev];
  if (c.ok) return AV1P_OK;
  CUDA_TRY(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major != 10) return fail(AV1P_ENODEV, "libav1p is built for sm_100a only; device has compute capability %d.x", cc_major);
  CUDA_TRY(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
  c.grid_sms = c.sms;
  if (const char* e = getenv("AV1P_GRID_SMS")) {
    const int v = atoi(e);
    if (v >= 2 && v <= c.sms) c.grid_sms = v & ~1;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(AV1P_ECUDA, "cuTensorMapEncodeTiled not available");
  c.encode = reinterpret_cast<EncodeTiledFn>(fn);
  // function attributes are per device: set them on THIS device
  CUDA_TRY(cudaFuncSetAttribute(fc_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(fc_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM_BYTES));
  if (const char* e = getenv("AV1P_FC_PAIR")) c.fc_pair = atoi(e) != 0;
  if (const char* e = getenv("AV1P_FC_RESID_EPI")) c.fc_resid_epi = atoi(e) != 0;
  if (const char* e = getenv("AV1P_PDL")) c.pdl = atoi(e) != 0;
  if (const char* e = getenv("AV1P_SPECULATE")) c.speculate = atoi(e) != 0;
  if (const char* e = getenv("AV1P_CR_RESID_EPI")) c.cr_resid_epi = std::max(0, std::min(2, atoi(e)));
  {
    int n_k = 0;
    const ConvResKernel* ks = conv_res_all_kernels(&n_k);
    for (int i = 0; i < n_k; ++i) CUDA_TRY(cudaFuncSetAttribute(ks[i], cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_BYTES));
  }
  CUDA_TRY(cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(stem_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(stem_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_BYTES));
  if (const char* e = getenv("AV1P_STEM_TMA")) c.stem_tma = atoi(e) != 0;
  CUDA_TRY(cudaFuncSetAttribute(extract_blocks_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, EX_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(extract_blocks_tma_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, EX_SMEM_BYTES));
  CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c.watchdog_host), 2 * sizeof(int), cudaHostAllocMapped));
  c.watchdog_host[0] = 0;      // [0] watchdog tag of a pipeline barrier that never completed
  c.watchdog_host[1] = 0;      // [1] a frame sample above 2048 reached the integer-pixel stem (see stem_tc.cuh, INT_PIX)
  CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c.watchdog_dev), c.watchdog_host, 0));
  c.ok = true;
  return AV1P_OK;
}

// 2-D fp16 tensor map, row-major [rows][cols].  Operand tiles: box {64 cols, box_rows}, 128-byte swizzle
// (one swizzle atom per row = the K-major layout tcgen05.mma reads).  Epilogue staging tiles (TMA stores):
// box {32 cols, 128 rows}, 64-byte swizzle.
int make_map_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems,
                uint32_t box_rows, uint32_t box_cols = 64, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (ld_elems * 2) % 16)
    return fail(AV1P_EINVAL, "tensor map base/stride not 16-byte aligned");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_ctx.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AV1P_ECUDA, "cuTensorMapEncodeTiled failed with %d", int(r));
  return AV1P_OK;
}
// Activation buffer in the tiled layout (act_off): `cols` columns (multiple of 64), `rows` rows (padded to 128)
// = a plain 2-D tensor [rows_padded * cols / 64][64]; tile (mt, kb) starts at tensor row (mt * cols/64 + kb) * 128.
int make_act_map(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, bool store) {
  if (cols == 0 || cols % 64) return fail(AV1P_EINVAL, "activation buffer width %llu is not a multiple of 64", (unsigned long long)cols);
  const uint64_t rows_padded = (rows + FC_TILE_M - 1) / FC_TILE_M * FC_TILE_M;
  return store ? make_map_2d(map, base, 64, rows_padded * (cols / 64), 64, FC_TILE_M, EPI_CHUNK, CU_TENSOR_MAP_SWIZZLE_64B)
               : make_map_2d(map, base, 64, rows_padded * (cols / 64), 64, FC_TILE_M);
}

// Residual connection of an FC layer (FC_EPI_ADD_RELU): the identity branch is accumulated on the tensor
// core.  Appends, to every N tile's schedule, one FC_W_IDENT entry per 64-wide K block of the residual
// buffer that falls inside the tile (hi plane through source 2, lo plane through source 3).  The caller
// has pointed a_map[2] / a_map[3] at the residual planes.
int add_residual_entries(FcParams& f, bool has_lo) {
  if (f.block_n % FC_TILE_K) return fail(AV1P_EINVAL, "residual FC layer needs block_n to be a multiple of 64");
  if (f.row_scale) return fail(AV1P_EINVAL, "residual FC layer cannot use a row scale");
  const float s = 1.0f / f.acc_scale;
  if (!(s >= 1.0f && s <= 32768.0f)) return fail(AV1P_EINVAL, "residual FC layer: weight scale %g outside [1, 2^15]", double(s));
  std::vector<uint16_t> src, w;
  int begin[FC_MAX_NT + 1];
  begin[0] = 0;
  const int per_tile = f.block_n / FC_TILE_K;
  const int step = f.pair_mode ? 2 : 1;             // regular entries come in (hi, lo) pairs in split precision
  if (f.pair_mode && !has_lo) return fail(AV1P_EINVAL, "split-precision residual FC layer needs the residual's lo plane");
  for (int t = 0; t < f.n_tiles; ++t) {
    // Residual K blocks carry almost no tensor work, so a run of them would expose one producer<->MMA ring round trip
    // each; interleave them with the regular entries (after the first one, which initialises the accumulator).
    int next_res = 0;
    auto push_res = [&]() {
      const uint16_t kb = uint16_t(f.out_col0 / FC_TILE_K + t * per_tile + next_res++);
      src.push_back(uint16_t((2u << 14) | kb));
      w.push_back(FC_W_IDENT);
      if (has_lo) {
        src.push_back(uint16_t((3u << 14) | kb));
        w.push_back(FC_W_IDENT);
      }
    };
    for (int e = f.kb_begin[t]; e < f.kb_begin[t + 1]; e += step) {
      for (int j = 0; j < step && e + j < f.kb_begin[t + 1]; ++j) {
        if ((f.kb_src[e + j] >> 14) >= 2) return fail(AV1P_EINVAL, "residual FC layer uses more than one activation source");
        src.push_back(f.kb_src[e + j]);
        w.push_back(f.kb_w[e + j]);
      }
      if (next_res < per_tile) push_res();
    }
    while (next_res < per_tile) push_res();
    begin[t + 1] = int(src.size());
  }
  if (src.size() > size_t(FC_MAX_KB)) return fail(AV1P_EINVAL, "residual FC layer: %zu schedule entries exceed %d", src.size(), FC_MAX_KB);
  for (size_t i = 0; i < src.size(); ++i) {
    f.kb_src[i] = src[i];
    f.kb_w[i] = w[i];
  }
  for (int t = 0; t <= f.n_tiles; ++t) f.kb_begin[t] = begin[t];
  return AV1P_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace

// Runtime switches of the CURRENT device's context (kernel-development / scheduling experiments; defaults are the product
// configuration).  "grid_sms": SMs a persistent kernel's grid may occupy (even, 2 .. SM count; 0 restores the SM count).
extern "C" int av1p_set_option(const char* name, int32_t value) {
  if (!name) return fail(AV1P_EINVAL, "null option name");
  if (int rc = ensure_ctx()) return rc;
  DeviceCtx& c = cur_ctx();
  if (!strcmp(name, "grid_sms")) {
    if (value == 0) value = c.sms;
    if (value < 2 || value > c.sms) return fail(AV1P_EINVAL, "grid_sms %d outside [2, %d]", value, c.sms);
    c.grid_sms = value & ~1;
    return AV1P_OK;
  }
  if (!strcmp(name, "fc_pair")) {
    c.fc_pair = value != 0;
    return AV1P_OK;
  }
  if (!strcmp(name, "stem_tma")) {          // read at every launch
    c.stem_tma = value != 0;
    return AV1P_OK;
  }
  if (!strcmp(name, "pdl")) {               // read at every launch
    c.pdl = value != 0;
    return AV1P_OK;
  }
  if (!strcmp(name, "speculate")) {         // read at every cascade call
    c.speculate = value != 0;
    return AV1P_OK;
  }
  if (!strcmp(name, "cr_resid_epi")) {      // read when a stage is planned (and by av1p_conv_res_forward)
    if (value < 0 || value > 2) return fail(AV1P_EINVAL, "cr_resid_epi %d outside 0..2", value);
    c.cr_resid_epi = value;
    return AV1P_OK;
  }
  return fail(AV1P_EINVAL, "unknown option '%s'", name);
}
extern "C" int av1p_get_option(const char* name) {
  if (!name || ensure_ctx()) return -1;
  DeviceCtx& c = cur_ctx();
  if (!strcmp(name, "grid_sms")) return c.grid_sms;
  if (!strcmp(name, "sms")) return c.sms;
  if (!strcmp(name, "fc_pair")) return c.fc_pair ? 1 : 0;
  if (!strcmp(name, "stem_tma")) return c.stem_tma ? 1 : 0;
  if (!strcmp(name, "pdl")) return c.pdl ? 1 : 0;
  if (!strcmp(name, "speculate")) return c.speculate ? 1 : 0;
  if (!strcmp(name, "cr_resid_epi")) return c.cr_resid_epi;
  return -1;
}

extern "C" int av1p_debug_watchdog(void) { return g_ctx.watchdog_host ? *g_ctx.watchdog_host : 0; }

// Launch of an op-program kernel.  These kernels call pdl_wait() before they touch anything a predecessor produced, so they
// may be launched with programmatic stream serialization: their prologue (barrier init, TMEM allocation, weight loads)
// overlaps the previous kernel's tail instead of following its last CTA (ptx_sm100.cuh).  `cluster` > 1: CTA clusters.
template <typename... KArgs, typename... Args>
static cudaError_t launch_op(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, unsigned cluster,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (g_ctx.pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// One FC layer: `rows` is the host-side upper bound on block rows (sizes the grid).  CTA-pair variant: clusters of two
// CTAs, each pair takes two M tiles of an item.
static int launch_fc(const FcParams& f, int rows, cudaStream_t st) {
  const int m_tiles = ceil_div(rows, FC_TILE_M);
  if (g_ctx.fc_pair && f.block_n % 16 == 0) {
    const int pairs = std::min(g_ctx.grid_sms / 2, ceil_div(m_tiles, 2) * f.n_tiles);
    CUDA_TRY(launch_op(fc_tcgen05_kernel<true>, unsigned(2 * std::max(pairs, 1)), FC_THREADS, FC_SMEM_BYTES, st, 2, f));
  } else {
    const int grid = std::min(g_ctx.grid_sms, m_tiles * f.n_tiles);
    CUDA_TRY(launch_op(fc_tcgen05_kernel<false>, unsigned(grid), FC_THREADS, FC_SMEM_BYTES, st, 1, f));
  }
  return AV1P_OK;
}

// ------------------------------------------------------------------------------ per-launch profiler
// Optional CUDA-event bracket around every kernel launch (bench.py's roofline leg).  Events are
// recorded on the launching stream; nothing synchronises until av1p_profile_end.
namespace {
enum ProfClass { PROF_STEM = 0, PROF_FC = 1, PROF_SAM = 2, PROF_FGVC = 3, PROF_ROUTE = 4, PROF_FINALIZE = 5, PROF_SE = 6, PROF_CONV = 7, PROF_CLASSES = 8 };
struct ProfRec { int cls; cudaEvent_t a, b; };
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
thread_local Profiler g_prof;
// NVTX range per kernel class around every launch (a no-op unless a tool such as ncu / nsys is attached): SURVEY.md section 5.
const char* const PROF_NAMES[PROF_CLASSES] = {"av1p:stem", "av1p:fc_tcgen05", "av1p:spatial_attention", "av1p:fgvc_tail",
                                              "av1p:route", "av1p:finalize_labels", "av1p:squeeze_excite", "av1p:conv_res_tcgen05"};
struct ProfScope {
  cudaStream_t st;
  bool active;
  ProfRec r;
  ProfScope(int cls, cudaStream_t s) : st(s), active(g_prof.on) {
    nvtxRangePushA(PROF_NAMES[cls]);
    if (!active) return;
    r.cls = cls;
    r.a = g_prof.get();
    r.b = g_prof.get();
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    nvtxRangePop();
    if (!active) return;
    cudaEventRecord(r.b, st);
    g_prof.recs.push_back(r);
  }
};
}  // namespace

extern "C" int av1p_profile_begin(void) {
  g_prof.on = true;
  return AV1P_OK;
}
extern "C" int av1p_profile_end(float* ms_by_class, int32_t* launches_by_class) {
  g_prof.on = false;
  for (int i = 0; i < PROF_CLASSES; ++i) {
    if (ms_by_class) ms_by_class[i] = 0.f;
    if (launches_by_class) launches_by_class[i] = 0;
  }
  for (ProfRec& r : g_prof.recs) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    if (ms_by_class) ms_by_class[r.cls] += ms;
    if (launches_by_class) launches_by_class[r.cls] += 1;
    g_prof.pool.push_back(r.a);
    g_prof.pool.push_back(r.b);
  }
  g_prof.recs.clear();
  return AV1P_OK;
}

// Per-launch variant: device time and class of every bracketed launch in issue order (tools/profile_ops.py pairs them
// with the packed op names).  Writes at most `cap` entries, *n_out = number of launches recorded.
extern "C" int av1p_profile_end_launches(float* ms, int32_t* cls, int32_t cap, int32_t* n_out) {
  g_prof.on = false;
  int n = 0;
  for (ProfRec& r : g_prof.recs) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
    if (n < cap) {
      if (ms) ms[n] = t;
      if (cls) cls[n] = r.cls;
    }
    ++n;
    g_prof.pool.push_back(r.a);
    g_prof.pool.push_back(r.b);
  }
  g_prof.recs.clear();
  if (n_out) *n_out = n;
  return AV1P_OK;
}

// Strided host->device copy of the luma planes only (2/3 of a 4:2:0 frame): one cudaMemcpy2DAsync.
extern "C" int av1p_upload_luma(const uint16_t* frames_host, int32_t n_frames, int32_t width, int32_t height,
                                int64_t frame_stride, uint16_t* luma_dev, void* stream) {
  if (!frames_host || !luma_dev || n_frames <= 0 || width <= 0 || height <= 0 || frame_stride < int64_t(width) * height)
    return fail(AV1P_EINVAL, "bad argument");
  const size_t row = size_t(width) * height * 2;
  CUDA_TRY(cudaMemcpy2DAsync(luma_dev, row, frames_host, size_t(frame_stride) * 2, row, size_t(n_frames),
                             cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return AV1P_OK;
}

// ------------------------------------------------------------------------------ model
struct av1p_model {
  std::vector<uint8_t> host;        // header + op table (host copy for planning)
  uint8_t* dev = nullptr;           // whole blob on the device
  Av1pBlobHeader hdr;
  std::vector<Av1pBlobOp> ops;
  std::vector<uint32_t> buf_cols;
  int device = 0;                   // ordinal of the device holding `dev`
  int block = 16;                   // luma block size the program was packed for (8 / 16 / 32 / 64)
};

extern "C" int av1p_model_create(const void* blob, size_t bytes, av1p_model** out) {
  if (!blob || !out) return fail(AV1P_EINVAL, "null argument");
  if (bytes < sizeof(Av1pBlobHeader)) return fail(AV1P_EINVAL, "blob too small");
  Av1pBlobHeader h;
  memcpy(&h, blob, sizeof h);
  if (h.magic != AV1P_BLOB_MAGIC || h.version != AV1P_BLOB_VERSION)
    return fail(AV1P_EINVAL, "bad blob magic/version (%08x, %u)", h.magic, h.version);
  if (h.total_bytes != bytes || h.ops_off + uint64_t(h.n_ops) * sizeof(Av1pBlobOp) > bytes ||
      h.bufs_off + uint64_t(h.n_bufs) * 4 > bytes)
    return fail(AV1P_EINVAL, "blob table out of range");
  if (int rc = ensure_ctx()) return rc;
  av1p_model* m = new (std::nothrow) av1p_model();
  if (!m) return fail(AV1P_ENOMEM, "host allocation failed");
  m->hdr = h;
  m->device = current_device();
  m->block = h.block_size ? int(h.block_size) : 16;
  if (m->block != 8 && m->block != 16 && m->block != 32 && m->block != 64) {
    delete m;
    return fail(AV1P_EINVAL, "blob packed for block size %d (8 / 16 / 32 / 64 are supported)", int(h.block_size));
  }
  m->ops.resize(h.n_ops);
  memcpy(m->ops.data(), static_cast<const uint8_t*>(blob) + h.ops_off, h.n_ops * sizeof(Av1pBlobOp));
  m->buf_cols.resize(h.n_bufs);
  memcpy(m->buf_cols.data(), static_cast<const uint8_t*>(blob) + h.bufs_off, h.n_bufs * 4);
  for (const Av1pBlobOp& op : m->ops) {
    auto bad_buf = [&](int b) { return b >= int(h.n_bufs); };
    if (bad_buf(op.src[0]) || bad_buf(op.src[1]) || bad_buf(op.src[2]) || bad_buf(op.src[3]) || bad_buf(op.aux) ||
        bad_buf(op.aux_lo) || bad_buf(op.out) || bad_buf(op.out_lo)) {
      delete m;
      return fail(AV1P_EINVAL, "op references a buffer that does not exist");
    }
    if (op.type == AV1P_OP_FC) {
      if (op.n_tiles < 1 || op.n_tiles > FC_MAX_NT || op.block_n < 32 || op.block_n > FC_MAX_N || op.block_n % 32 ||
          op.n_kb_total < 1 || op.n_kb_total > AV1P_BLOB_MAX_KB || op.tail_n > FC_TAIL_MAX || op.n_w_chunks < 1 ||
          op.out_col0 < 0 || op.out_col0 % FC_TILE_K ||
          op.w_off + uint64_t(op.n_w_chunks) * op.block_n * 128 > bytes) {
        delete m;
        return fail(AV1P_EINVAL, "malformed FC op");
      }
      // the device kernels walk kb_begin[t] .. kb_begin[t + 1] and index kb_src / kb_w with those values: a truncated or
      // corrupt blob must be rejected here, not turn into out-of-bounds schedule reads and arbitrary TMA coordinates
      bool sched_ok = op.kb_begin[0] == 0 && op.kb_begin[op.n_tiles] == op.n_kb_total;
      for (int t = 0; t < op.n_tiles && sched_ok; ++t)
        sched_ok = op.kb_begin[t] >= 0 && op.kb_begin[t] <= op.kb_begin[t + 1] && op.kb_begin[t + 1] <= op.n_kb_total &&
                   (!op.pair_mode || ((op.kb_begin[t] | op.kb_begin[t + 1]) & 1) == 0);
      for (int i = 0; i < op.n_kb_total && sched_ok; ++i) sched_ok = int(op.kb_w[i]) < op.n_w_chunks && (op.kb_src[i] >> 14) < FC_MAX_SRC;
      if (!sched_ok) {
        delete m;
        return fail(AV1P_EINVAL, "malformed FC op: K-block schedule is not monotonic / in range");
      }
    }
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&m->dev), bytes);
  if (e == cudaSuccess) e = cudaMemcpy(m->dev, blob, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (m->dev) cudaFree(m->dev);
    delete m;
    return fail(AV1P_ECUDA, "blob upload failed: %s", cudaGetErrorString(e));
  }
  *out = m;
  return AV1P_OK;
}
extern "C" void av1p_model_destroy(av1p_model* m) {
  if (!m) return;
  if (m->dev) cudaFree(m->dev);
  delete m;
}
extern "C" int av1p_model_num_outputs(const av1p_model* m) { return m ? int(m->hdr.n_out) : 0; }

// ------------------------------------------------------------------------------ stage (plan)
namespace {

struct PlannedOp {
  int type = 0;
  FcParams fc;          // AV1P_OP_FC
  ConvResParams cr;     // AV1P_OP_CONV_RES
  StemParams stem;      // AV1P_OP_STEM
  const __half* stem_w_int = nullptr;   // integer-pixel weight set (frames input) and its scale
  float stem_scale_int = 0.f;
  const __half* stem_w_raw = nullptr;   // raw-word weight set of the TMA-staged frame kernel (stem_tma.cuh) and its scale
  float stem_scale_raw = 0.f;
  const __half* src = nullptr;   // SAM / FGVC
  const __half* src_lo = nullptr;
  __half* dst = nullptr;         // SE
  __half* dst_lo = nullptr;
  int se_c = 0, se_npos = 0;
  int ld = 0;
  float f0 = 0.f, f1 = 0.f;
  const float* w = nullptr;
  StemGenParams sg;              // AV1P_OP_STEM_GEN
  int gen_kb_in = 0, gen_kb_out = 0;   // AV1P_OP_SE_GEN / AV1P_OP_SAM_POOL: 64-column blocks per row of source / destination
};

// Activation region shared by every stage of a cascade: buffer i has max-over-models cols.
constexpr int SAM_PART_SLOTS = 8;      // se4.fc2: 2 N tiles x 2 epilogue groups x 2 column halves
struct ActLayout {
  std::vector<uint32_t> cols;
  std::vector<size_t> off;
  size_t row_scale_off = 0, sam_part_off = 0, bytes = 0;
  int cap = 0;
};

ActLayout make_act_layout(const av1p_model* const* models, int n_models, int capacity) {
  ActLayout L;
  L.cap = ceil_div(std::max(capacity, 1), FC_TILE_M) * FC_TILE_M;
  size_t nb = 0;
  for (int i = 0; i < n_models; ++i) nb = std::max(nb, models[i]->buf_cols.size());
  L.cols.assign(nb, 0);
  for (int i = 0; i < n_models; ++i)
    for (size_t b = 0; b < models[i]->buf_cols.size(); ++b) L.cols[b] = std::max(L.cols[b], models[i]->buf_cols[b]);
  // Buffers whose live ranges (first .. last op that touches them) never overlap in any of the models share memory: the
  // three 1024-column layer1 buffers are dead once layer2 runs, so the 512 / 256-column buffers of layers 2..4 live inside
  // them (21.8 -> ~13 KB per block row).  Two buffers touched by the same op interfere by construction (inclusive
  // ranges), so no kernel ever reads and writes one piece of memory.  AV1P_ALIAS_BUFFERS=0 gives every buffer its own memory.
  std::vector<size_t> size(nb);
  for (size_t b = 0; b < nb; ++b) size[b] = align_up(size_t(L.cap) * L.cols[b] * 2, 1024);
  std::vector<std::vector<char>> clash(nb, std::vector<char>(nb, 0));
  const char* alias_env = getenv("AV1P_ALIAS_BUFFERS");
  const bool alias = !(alias_env && atoi(alias_env) == 0);
  for (int i = 0; i < n_models; ++i) {
    std::vector<int> lo(nb, INT_MAX), hi(nb, -1);
    int k = 0;
    for (const Av1pBlobOp& op : models[i]->ops) {
      const int ids[8] = {op.src[0], op.src[1], op.src[2], op.src[3], op.aux, op.aux_lo, op.out, op.out_lo};
      for (int id : ids)
        if (id >= 0 && size_t(id) < nb) {
          lo[id] = std::min(lo[id], k);
          hi[id] = std::max(hi[id], k);
        }
      ++k;
    }
    for (size_t a = 0; a < nb; ++a)
      for (size_t b = 0; b < nb; ++b)
        if (a != b && hi[a] >= 0 && hi[b] >= 0 && lo[a] <= hi[b] && lo[b] <= hi[a]) clash[a][b] = 1;
  }
  std::vector<size_t> order(nb);
  for (size_t b = 0; b < nb; ++b) order[b] = b;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return size[a] > size[b]; });
  L.off.assign(nb, 0);
  std::vector<char> placed(nb, 0);
  size_t o = 0;
  for (size_t b : order) {
    // lowest offset at which [off, off + size) avoids every placed buffer this one clashes with
    size_t off = 0;
    for (bool moved = true; moved;) {
      moved = false;
      for (size_t a = 0; a < nb; ++a)
        if (placed[a] && (!alias || clash[a][b]) && off < L.off[a] + size[a] && L.off[a] < off + size[b]) {
          off = L.off[a] + size[a];
          moved = true;
        }
    }
    L.off[b] = off;
    placed[b] = 1;
    o = std::max(o, off + size[b]);
  }
  L.row_scale_off = o;
  L.sam_part_off = o + align_up(size_t(L.cap) * 4, 1024);
  o += align_up(size_t(L.cap) * 4, 1024);
  o += align_up(size_t(L.cap) * SAM_PART_SLOTS * 2 * 4, 1024);        // spatial-attention partials (FcParams::sam_part)
  L.bytes = o;
  return L;
}

}  // namespace

struct av1p_stage {
  const av1p_model* model = nullptr;
  int cap = 0;
  std::vector<PlannedOp> ops;
  float* row_scale = nullptr;
  float* sam_part = nullptr;
  float* own_logits = nullptr;   // unused by the cascade (it passes its own logits buffers)
  int* range_flag = nullptr;     // frame input: word the stem sets when a sample above 2048 is met (cascade plans point it
                                 // into their workspace, av1p_cascade_buffer(c, 8); standalone stages leave it null)
  float* features_out = nullptr; // FGVC models: optional fp32 [rows][512] L2-normalised features (av1p_stage_set_features_out)
};

namespace {

int plan_stage(const av1p_model* m, const ActLayout& L, uint8_t* act_base, av1p_stage* s) {
  s->model = m;
  s->cap = L.cap;
  s->row_scale = reinterpret_cast<float*>(act_base + L.row_scale_off);
  s->sam_part = reinterpret_cast<float*>(act_base + L.sam_part_off);
  auto buf = [&](int id) -> __half* { return id < 0 ? nullptr : reinterpret_cast<__half*>(act_base + L.off[id]); };
  auto at = [&](uint64_t off) -> const uint8_t* { return off ? m->dev + off : nullptr; };
  for (const Av1pBlobOp& op : m->ops) {
    PlannedOp P;
    P.type = op.type;
    switch (op.type) {
      case AV1P_OP_STEM: {
        memset(&P.stem, 0, sizeof P.stem);
        P.stem.w = reinterpret_cast<const __half*>(at(op.w_off));
        P.stem.acc_scale = op.f0;
        P.stem_w_int = P.stem.w ? P.stem.w + 2 * 128 * 64 : nullptr;
        P.stem_scale_int = op.f1;
        P.stem_w_raw = P.stem.w ? P.stem.w + 4 * 128 * 64 : nullptr;
        P.stem_scale_raw = ldexpf(1.0f, op.tail_n);
        if (op.tail_n < -16 || op.tail_n > 24) return fail(AV1P_EINVAL, "stem op: raw weight-set exponent %d out of range", op.tail_n);
        if (!(op.f1 > 0.f)) return fail(AV1P_EINVAL, "stem op without the integer-pixel weight set");
        P.stem.err_flag = g_ctx.watchdog_dev;
        P.stem.b = reinterpret_cast<const float*>(at(op.bias_off));
        P.stem.out = buf(op.out);
        P.stem.out_lo = buf(op.out_lo);
        if (!P.stem.w || !P.stem.b || !P.stem.out || L.cols[op.out] != 1024) return fail(AV1P_EINVAL, "malformed stem op");
        break;
      }
      case AV1P_OP_FC: {
        FcParams& f = P.fc;
        memset(&f, 0, sizeof f);
        if (op.src[0] < 0 || (op.out < 0 && op.epi != FC_EPI_HEAD)) return fail(AV1P_EINVAL, "FC op without source/out");
        for (int i = 0; i < FC_MAX_SRC; ++i) {
          const int sb = op.src[i] >= 0 ? op.src[i] : op.src[0];
          if (int rc = make_act_map(&f.a_map[i], buf(sb), L.cols[sb], L.cap, false)) return rc;
          f.src_kb[i] = int(L.cols[sb] / 64);
        }
        if (int rc = make_map_2d(&f.w_map, at(op.w_off), 64, uint64_t(op.n_w_chunks) * op.block_n, 64, op.block_n)) return rc;
        if (int rc = make_map_2d(&f.w_half_map, at(op.w_off), 64, uint64_t(op.n_w_chunks) * op.block_n, 64, op.block_n / 2)) return rc;
        f.n_tiles = op.n_tiles;
        f.block_n = op.block_n;
        f.epi = op.epi;
        f.bias = reinterpret_cast<const float*>(at(op.bias_off));
        f.row_scale = (op.use_row_scale & 1) ? s->row_scale : nullptr;
        f.aux_row_scale = (op.use_row_scale & 2) ? s->row_scale : nullptr;
        if (op.use_row_scale & 4) {        // leave the spatial-attention partials of this layer's output behind
          if (op.n_tiles * 4 != SAM_PART_SLOTS || (op.block_n / EPI_CHUNK) % 2 || op.n_tiles * op.block_n != 512)
            return fail(AV1P_EINVAL, "spatial-attention partials need a 512-wide output in 2 N tiles of 256");
          f.sam_part = s->sam_part;
        }
        f.acc_scale = op.f0;
        f.pair_mode = op.pair_mode;
        if (op.pair_mode) {
          for (int i = 0; i <= op.n_tiles; ++i)
            if (op.kb_begin[i] & 1) return fail(AV1P_EINVAL, "pair-mode schedule with an odd entry count");
        }
        f.aux = buf(op.aux);
        f.aux_lo = buf(op.aux_lo);
        f.aux_kb = op.aux >= 0 ? int(L.cols[op.aux] / 64) : 0;
        f.out = buf(op.out);
        f.out_lo = buf(op.out_lo);
        f.out_kb = op.out >= 0 ? int(L.cols[op.out] / 64) : 0;
        f.out_col0 = op.out_col0;
        if ((op.aux_lo >= 0 && L.cols[op.aux_lo] != L.cols[op.aux]) || (op.out_lo >= 0 && L.cols[op.out_lo] != L.cols[op.out]))
          return fail(AV1P_EINVAL, "hi/lo buffers differ in width");
        f.tail_w = reinterpret_cast<const float*>(at(op.tail_w_off));
        f.tail_b = reinterpret_cast<const float*>(at(op.tail_b_off));
        f.tail_n = op.tail_n;
        f.err_flag = g_ctx.watchdog_dev;
        if ((op.epi == FC_EPI_ADD_RELU || op.epi == FC_EPI_GATE || op.epi == FC_EPI_ADD) && !f.aux) return fail(AV1P_EINVAL, "FC op needs aux");
        if (op.epi == FC_EPI_HEAD && (!f.tail_w || !f.tail_b || op.n_tiles != 1 || op.tail_n < 1))
          return fail(AV1P_EINVAL, "malformed head op");
        if (op.epi != FC_EPI_HEAD && op.out_col0 + op.n_tiles * op.block_n > f.out_kb * 64) return fail(AV1P_EINVAL, "FC output wider than its buffer");
        if (op.out_col0 && (op.epi == FC_EPI_GATE || op.epi == FC_EPI_ADD || op.epi == FC_EPI_HEAD || (op.use_row_scale & 4)))
          return fail(AV1P_EINVAL, "an FC op with a column offset cannot be a gate / adapter / head layer");
        for (int i = 0; i <= op.n_tiles; ++i) f.kb_begin[i] = op.kb_begin[i];
        if (f.kb_begin[0] != 0 || f.kb_begin[op.n_tiles] != op.n_kb_total) return fail(AV1P_EINVAL, "bad K-block schedule");
        for (int i = 0; i < op.n_kb_total; ++i) {
          f.kb_src[i] = op.kb_src[i];
          f.kb_w[i] = op.kb_w[i];
          const int sb = op.src[op.kb_src[i] >> 14];
          if (sb < 0 || uint32_t(op.kb_src[i] & 0x3FFF) * 64 + 64 > L.cols[sb]) return fail(AV1P_EINVAL, "K block outside its source buffer");
          if (int(op.kb_w[i]) >= op.n_w_chunks) return fail(AV1P_EINVAL, "weight chunk index out of range");
        }
        if (op.out >= 0) {
          if (int rc = make_act_map(&f.out_map[0], buf(op.out), L.cols[op.out], L.cap, true)) return rc;
          if (op.out_lo >= 0)
            if (int rc = make_act_map(&f.out_map[1], buf(op.out_lo), L.cols[op.out_lo], L.cap, true)) return rc;
        }
        if (op.epi == FC_EPI_GATE || op.epi == FC_EPI_ADD) {
          if (op.n_tiles * op.block_n > int(L.cols[op.aux])) return fail(AV1P_EINVAL, "gate input narrower than the FC output");
          if (int rc = make_act_map(&f.aux_map[0], buf(op.aux), L.cols[op.aux], L.cap, true)) return rc;
          if (op.aux_lo >= 0)
            if (int rc = make_act_map(&f.aux_map[1], buf(op.aux_lo), L.cols[op.aux_lo], L.cap, true)) return rc;
        }
        if (op.epi == FC_EPI_ADD_RELU && g_ctx.fc_resid_epi) {
          // residual added by the epilogue from the aux ring
          if (op.n_tiles * op.block_n > int(L.cols[op.aux])) return fail(AV1P_EINVAL, "residual narrower than the FC output");
          if (f.row_scale) return fail(AV1P_EINVAL, "residual FC layer cannot use a row scale");
          f.aux_epi = 1;
          if (int rc = make_act_map(&f.aux_map[0], buf(op.aux), L.cols[op.aux], L.cap, true)) return rc;
          if (op.aux_lo >= 0)
            if (int rc = make_act_map(&f.aux_map[1], buf(op.aux_lo), L.cols[op.aux_lo], L.cap, true)) return rc;
        } else if (op.epi == FC_EPI_ADD_RELU) {
          if (op.out_col0 + op.n_tiles * op.block_n > int(L.cols[op.aux])) return fail(AV1P_EINVAL, "residual narrower than the FC output");
          if (int rc = make_act_map(&f.a_map[2], buf(op.aux), L.cols[op.aux], L.cap, false)) return rc;
          f.src_kb[2] = f.src_kb[3] = int(L.cols[op.aux] / 64);
          if (op.aux_lo >= 0)
            if (int rc = make_act_map(&f.a_map[3], buf(op.aux_lo), L.cols[op.aux_lo], L.cap, false)) return rc;
          if (int rc = add_residual_entries(f, op.aux_lo >= 0)) return rc;
        }
        break;
      }
      case AV1P_OP_CONV_RES: {
        ConvResParams& f = P.cr;
        memset(&f, 0, sizeof f);
        const bool split = op.pair_mode != 0;
        if (op.src[0] < 0 || op.out < 0 || L.cols[op.src[0]] != 1024 || L.cols[op.out] != 1024 || (split && (op.src[1] < 0 || op.out_lo < 0)) ||
            !op.w_off || !op.bias_off || op.w_off + uint64_t(split ? 2 : 1) * CR_W_PLANE_BYTES > m->hdr.total_bytes)
          return fail(AV1P_EINVAL, "malformed resident-conv op");
        if (int rc = make_act_map(&f.a_map[0], buf(op.src[0]), 1024, L.cap, false)) return rc;
        if (int rc = make_act_map(&f.a_map[1], buf(split ? op.src[1] : op.src[0]), 1024, L.cap, false)) return rc;
        if (int rc = make_map_2d(&f.w_map, at(op.w_off), 64, uint64_t(split ? 2 : 1) * 9 * 64, 64, 64)) return rc;
        f.split = split ? 1 : 0;
        f.epi = op.epi;
        f.bias = reinterpret_cast<const float*>(at(op.bias_off));
        f.acc_scale = op.f0;
        f.out = buf(op.out);
        f.out_lo = buf(op.out_lo);
        f.out_kb = 16;
        f.err_flag = g_ctx.watchdog_dev;
        if (op.epi != FC_EPI_RELU && op.epi != FC_EPI_ADD_RELU && op.epi != FC_EPI_LINEAR) return fail(AV1P_EINVAL, "resident-conv epilogue %d", op.epi);
        if (!(1.0f / op.f0 >= 1.0f && 1.0f / op.f0 <= 32768.0f)) return fail(AV1P_EINVAL, "resident-conv weight scale outside [1, 2^15]");
        if (int rc = make_act_map(&f.out_map[0], buf(op.out), 1024, L.cap, true)) return rc;
        if (op.out_lo >= 0)
          if (int rc = make_act_map(&f.out_map[1], buf(op.out_lo), 1024, L.cap, true)) return rc;
        if (op.epi == FC_EPI_ADD_RELU) {
          if (op.aux < 0 || L.cols[op.aux] != 1024 || (op.aux_lo >= 0 && L.cols[op.aux_lo] != 1024))
            return fail(AV1P_EINVAL, "resident-conv op needs a 1024-wide residual");
          if (int rc = make_act_map(&f.a_map[2], buf(op.aux), 1024, L.cap, false)) return rc;
          if (int rc = make_act_map(&f.a_map[3], buf(op.aux_lo >= 0 ? op.aux_lo : op.aux), 1024, L.cap, false)) return rc;
          f.has_aux_lo = op.aux_lo >= 0 ? 1 : 0;
          if (g_ctx.cr_resid_epi) {
            f.resid_epi = g_ctx.cr_resid_epi;
            f.aux = buf(op.aux);
            f.aux_lo = buf(op.aux_lo);
            f.aux_kb = 16;
            if (int rc = make_act_map(&f.res_map[0], buf(op.aux), 1024, L.cap, true)) return rc;
            if (int rc = make_act_map(&f.res_map[1], buf(op.aux_lo >= 0 ? op.aux_lo : op.aux), 1024, L.cap, true)) return rc;
          }
        }
        if (!conv_res_build_schedule(f)) return fail(AV1P_EINVAL, "resident-conv schedule does not fit its tables");
        break;
      }
      case AV1P_OP_SAM:
      case AV1P_OP_FGVC_TAIL: {
        P.se_npos = (op.type == AV1P_OP_SAM && op.tail_n == 1) ? SAM_PART_SLOTS : 0;     // SAM from partials
        P.src = buf(op.src[0]);
        P.src_lo = buf(op.src[1]);
        if (!P.src || L.cols[op.src[0]] != 512 || (op.src[1] >= 0 && L.cols[op.src[1]] != 512)) return fail(AV1P_EINVAL, "SAM/FGVC op needs a 512-wide source");
        P.ld = 512;
        P.f0 = op.f0;
        P.f1 = op.f1;
        P.w = reinterpret_cast<const float*>(at(op.w_off));
        if (op.type == AV1P_OP_FGVC_TAIL && !P.w) return fail(AV1P_EINVAL, "FGVC tail without weights");
        break;
      }
      case AV1P_OP_SE: {
        P.src = buf(op.src[0]);
        P.src_lo = buf(op.src[1]);
        P.dst = buf(op.out);
        P.dst_lo = buf(op.out_lo);
        P.se_c = op.block_n;
        P.se_npos = op.n_tiles;
        P.w = reinterpret_cast<const float*>(at(op.w_off));
        const bool shape_ok = (P.se_c == 64 && P.se_npos == 16) || (P.se_c == 128 && P.se_npos == 4) || (P.se_c == 256 && P.se_npos == 1);
        if (!P.src || !P.dst || !P.w || !shape_ok || int(L.cols[op.src[0]]) != P.se_c * P.se_npos ||
            int(L.cols[op.out]) != P.se_c * P.se_npos || (op.src[1] >= 0) != (op.out_lo >= 0))
          return fail(AV1P_EINVAL, "malformed SE op");
        break;
      }
      case AV1P_OP_STEM_GEN: {
        memset(&P.sg, 0, sizeof P.sg);
        const int g1 = op.n_tiles / 4;
        P.sg.bs = op.n_tiles;
        P.sg.w = reinterpret_cast<const float*>(at(op.w_off));
        P.sg.b = reinterpret_cast<const float*>(at(op.bias_off));
        P.sg.out = buf(op.out);
        P.sg.out_lo = buf(op.out_lo);
        P.sg.out_kb = op.out >= 0 ? int(L.cols[op.out] / 64) : 0;
        if (op.n_tiles != m->block || !P.sg.w || !P.sg.b || !P.sg.out || P.sg.out_kb < g1 * g1)
          return fail(AV1P_EINVAL, "malformed generic stem op");
        break;
      }
      case AV1P_OP_SE_GEN: {
        P.src = buf(op.src[0]);
        P.src_lo = buf(op.src[1]);
        P.dst = buf(op.out);
        P.dst_lo = buf(op.out_lo);
        P.se_c = op.block_n;
        P.se_npos = op.n_tiles;
        P.w = reinterpret_cast<const float*>(at(op.w_off));
        const bool c_ok = P.se_c == 64 || P.se_c == 128 || P.se_c == 256 || P.se_c == 512;
        if (!P.src || !P.dst || !P.w || !c_ok || P.se_npos < 1 || L.cols[op.src[0]] != L.cols[op.out] ||
            int(L.cols[op.src[0]]) < P.se_c * P.se_npos || (op.src[1] >= 0) != (op.out_lo >= 0))
          return fail(AV1P_EINVAL, "malformed generic SE op");
        P.gen_kb_in = int(L.cols[op.src[0]] / 64);
        break;
      }
      case AV1P_OP_SAM_POOL: {
        P.src = buf(op.src[0]);
        P.src_lo = buf(op.src[1]);
        P.dst = buf(op.out);
        P.dst_lo = buf(op.out_lo);
        P.se_npos = op.n_tiles;            // map side g
        P.w = reinterpret_cast<const float*>(at(op.w_off));
        if (!P.src || !P.dst || !P.w || P.se_npos < 1 || P.se_npos > 2 || int(L.cols[op.src[0]]) < P.se_npos * P.se_npos * 512 ||
            L.cols[op.out] < 512 || (op.src[1] >= 0) != (op.out_lo >= 0))
          return fail(AV1P_EINVAL, "malformed attention / pooling op");
        P.gen_kb_in = int(L.cols[op.src[0]] / 64);
        P.gen_kb_out = int(L.cols[op.out] / 64);
        break;
      }
      default:
        return fail(AV1P_EINVAL, "unknown op type %d", op.type);
    }
    s->ops.push_back(P);
  }
  return AV1P_OK;
}

int convert_input(const av1p_input* in, StemInput* si, int block = 16) {
  memset(si, 0, sizeof *si);
  if (!in) return fail(AV1P_EINVAL, "null input");
  si->kind = in->kind;
  if (in->kind == 0) {
    if (!in->frames_dev || in->width <= 0 || in->height <= 0 || in->pitch < in->width)
      return fail(AV1P_EINVAL, "bad frame geometry");
    si->frames = in->frames_dev;
    si->frame_stride = in->frame_stride;
    si->width = in->width;
    si->height = in->height;
    si->pitch = in->pitch;
    si->n_frames = std::max(in->n_frames, 1);
    si->blocks_x = ceil_div(in->width, block);
    si->blocks_per_frame = si->blocks_x * ceil_div(in->height, block);
    si->inv_bx = ~0ULL / (unsigned long long)si->blocks_x + 1ULL;              // unused when the divisor is 1
    si->inv_bpf = ~0ULL / (unsigned long long)si->blocks_per_frame + 1ULL;
  } else if (in->kind == 1) {
    if (!in->images_dev) return fail(AV1P_EINVAL, "null image tensor");
    si->images = in->images_dev;
    si->blocks_per_frame = 1;
    si->blocks_x = 1;
  } else {
    return fail(AV1P_EINVAL, "unknown input kind %d", in->kind);
  }
  return AV1P_OK;
}

int run_stage(av1p_stage* s, const StemInput& si, const int32_t* idx, const int32_t* n_dev, int n, float* logits,
              cudaStream_t st) {
  if (n <= 0) return AV1P_OK;
  if (n > s->cap) return fail(AV1P_EINVAL, "n=%d exceeds the stage capacity %d", n, s->cap);
  if (s->model && s->model->device != current_device())
    return fail(AV1P_EINVAL, "stage was planned on device %d but device %d is current", s->model->device, current_device());
  int op_index = 0;
  for (PlannedOp& P : s->ops) {
    switch (P.type) {
      case AV1P_OP_STEM: {
        StemParams sp = P.stem;
        sp.in = si;
        sp.idx = idx;
        sp.n_dev = n_dev;
        sp.n = n;
        sp.range_flag = s->range_flag;
        const int grid = std::min(ceil_div(n, ST_BLOCKS), g_ctx.grid_sms);
        ProfScope ps(PROF_STEM, st);
        // frames whose geometry a tensor map can describe (16-byte aligned base, row and frame strides): TMA-staged kernel
        if (si.kind == 0 && g_ctx.stem_tma && (reinterpret_cast<uintptr_t>(si.frames) & 15u) == 0 && si.pitch % 8 == 0 &&
            (si.n_frames == 1 || si.frame_stride % 8 == 0)) {
          StemTmaParams tp;
          tp.s = sp;
          tp.s.w = P.stem_w_raw;
          tp.s.acc_scale = P.stem_scale_raw;
          {
            const char* dbg = getenv("AV1P_STEM_DEBUG");
            tp.debug = dbg ? atoi(dbg) : 0;
          }
          cuuint64_t dims[3] = {cuuint64_t(si.width), cuuint64_t(si.height), cuuint64_t(si.n_frames)};
          cuuint64_t strides[2] = {cuuint64_t(si.pitch) * 2,
                                   cuuint64_t(si.n_frames == 1 ? (long long)si.pitch * si.height : si.frame_stride) * 2};
          // unrouted input whose block rows hold a multiple of four blocks: one 64 x 16 box per tile (see stem_tma.cuh, WIDE)
          // (opt-in, AV1P_STEM_WIDE=1: fewer and larger TMA requests, but its swizzled source addressing costs the builders more
          // instructions - same box, 518 k rows: 1061 / 1106 us vs 1051 / 1008 us for per-block boxes, 1180 / 1158 us for the
          // per-thread gather kernel)
          const char* wide_env = getenv("AV1P_STEM_WIDE");
          const bool wide = idx == nullptr && si.blocks_x % ST_BLOCKS == 0 && wide_env && atoi(wide_env) != 0;
          cuuint32_t box[3] = {wide ? 64u : 16u, 16, 1};
          cuuint32_t estr[3] = {1, 1, 1};
          CUresult r = g_ctx.encode(&tp.map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(si.frames), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) return fail(AV1P_ECUDA, "cuTensorMapEncodeTiled (stem frames %dx%dx%d) failed: %d", si.width, si.height, si.n_frames, int(r));
          if (wide) CUDA_TRY(launch_op(stem_tma_kernel<true>, unsigned(grid), SM_THREADS, SM_SMEM_BYTES, st, 1, tp));
          else CUDA_TRY(launch_op(stem_tma_kernel<false>, unsigned(grid), SM_THREADS, SM_SMEM_BYTES, st, 1, tp));
        } else if (si.kind == 0) {
          // frames: integer pixel plane + the / 1023 weight set (see stem_tc.cuh, INT_PIX)
          sp.w = P.stem_w_int;
          sp.acc_scale = P.stem_scale_int;
          CUDA_TRY(launch_op(stem_tc_kernel<true>, unsigned(grid), ST_THREADS, ST_SMEM_BYTES, st, 1, sp));
        } else {
          CUDA_TRY(launch_op(stem_tc_kernel<false>, unsigned(grid), ST_THREADS, ST_SMEM_BYTES, st, 1, sp));
        }
        break;
      }
      case AV1P_OP_FC: {
        P.fc.n_rows_dev = n_dev;
        P.fc.n_rows = n;
        P.fc.logits = logits;
        ProfScope ps(PROF_FC, st);
        if (int rc = launch_fc(P.fc, n, st)) return rc;
        break;
      }
      case AV1P_OP_CONV_RES: {
        P.cr.n_rows_dev = n_dev;
        P.cr.n_rows = n;
        const int grid = 2 * std::max(1, std::min(g_ctx.grid_sms / 2, ceil_div(n, FC_TILE_M)));      // (M tile, half) items, even grid
        ProfScope ps(PROF_CONV, st);
        CUDA_TRY(launch_op(conv_res_kernel_for(P.cr), unsigned(grid), CR_THREADS, CR_SMEM_BYTES, st, 1, P.cr));
        break;
      }
      case AV1P_OP_SAM: {
        ProfScope ps(PROF_SAM, st);
        if (P.se_npos > 0) {
          CUDA_TRY(launch_op(sam_finish_kernel, unsigned(std::min(ceil_div(n, 256), g_ctx.sms * 8)), 256, 0, st, 1,
                             (const float*)s->sam_part, P.se_npos, n_dev, n, P.f0, P.f1, s->row_scale));
        } else {
          const int grid = std::min(ceil_div(n, 8), g_ctx.sms * 8);
          CUDA_TRY(launch_op(sam_gate_kernel, unsigned(grid), 256, 0, st, 1, (const __half*)P.src, (const __half*)P.src_lo, P.ld, n_dev, n,
                             P.f0, P.f1, s->row_scale));
        }
        break;
      }
      case AV1P_OP_SE: {
        ProfScope ps(PROF_SE, st);
        const size_t smem = size_t(2) * (P.se_c / 16) * P.se_c * sizeof(float);
        // rows per warp iteration (template R): sharing one pass over the shared-memory weights between two rows measured
        // slower (fewer warps in flight), so every shape runs with R = 1
        if (P.se_c == 64)
          CUDA_TRY(launch_op(se_kernel<64, 16, 1>, unsigned(std::min(ceil_div(n, 8), g_ctx.sms * 8)), 256, smem, st, 1, (const __half*)P.src,
                             (const __half*)P.src_lo, P.dst, P.dst_lo, n_dev, n, (const float*)P.w));
        else if (P.se_c == 128)
          CUDA_TRY(launch_op(se_kernel<128, 4, 1>, unsigned(std::min(ceil_div(n, 8), g_ctx.sms * 8)), 256, smem, st, 1, (const __half*)P.src,
                             (const __half*)P.src_lo, P.dst, P.dst_lo, n_dev, n, (const float*)P.w));
        else
          CUDA_TRY(launch_op(se_kernel<256, 1, 1>, unsigned(std::min(ceil_div(n, 8), g_ctx.sms * 8)), 256, smem, st, 1, (const __half*)P.src,
                             (const __half*)P.src_lo, P.dst, P.dst_lo, n_dev, n, (const float*)P.w));
        break;
      }
      case AV1P_OP_STEM_GEN: {
        StemGenParams sp = P.sg;
        sp.in = si;
        sp.idx = idx;
        sp.n_dev = n_dev;
        sp.n = n;
        const size_t smem = size_t(sp.bs + 6) * (sp.bs + 7) * sizeof(float);
        ProfScope ps(PROF_STEM, st);
        stem_generic_kernel<<<std::min(n, g_ctx.sms * 8), SG_THREADS, smem, st>>>(sp);
        break;
      }
      case AV1P_OP_SE_GEN: {
        ProfScope ps(PROF_SE, st);
        se_generic_kernel<<<std::min(n, g_ctx.sms * 8), SEG_THREADS, 0, st>>>(P.src, P.src_lo, P.dst, P.dst_lo, n_dev, n, P.se_c, P.se_npos,
                                                                             P.gen_kb_in, P.w);
        break;
      }
      case AV1P_OP_SAM_POOL: {
        ProfScope ps(PROF_SAM, st);
        sam_pool_kernel<<<std::min(n, g_ctx.sms * 16), SP_THREADS, 0, st>>>(P.src, P.src_lo, P.dst, P.dst_lo, n_dev, n, P.se_npos, P.gen_kb_in,
                                                                           P.gen_kb_out, P.w);
        break;
      }
      case AV1P_OP_FGVC_TAIL: {
        const int grid = std::min(ceil_div(n, 8), g_ctx.sms * 8);
        ProfScope ps(PROF_FGVC, st);
        CUDA_TRY(launch_op(fgvc_tail_kernel, unsigned(grid), 256, 0, st, 1, (const __half*)P.src, (const __half*)P.src_lo, P.ld, n_dev, n,
                           (const float*)P.w, P.f0, logits, s->features_out));
        break;
      }
    }
    // a launch that fails (bad configuration, missing shared-memory opt-in ...) is reported at the op that caused it
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess)
      return fail(AV1P_ECUDA, "launch of op %d (type %d) failed: %s", op_index, P.type, cudaGetErrorString(e));
    ++op_index;
  }
  return AV1P_OK;
}

}  // namespace

extern "C" size_t av1p_stage_workspace_bytes(const av1p_model* m, int32_t capacity_rows) {
  if (!m || capacity_rows <= 0) return 0;
  return make_act_layout(&m, 1, capacity_rows).bytes + 1024;
}

extern "C" int av1p_stage_create(const av1p_model* m, int32_t capacity_rows, void* ws, size_t ws_bytes, av1p_stage** out) {
  if (!m || !ws || !out || capacity_rows <= 0) return fail(AV1P_EINVAL, "bad argument");
  if (int rc = ensure_ctx()) return rc;
  ActLayout L = make_act_layout(&m, 1, capacity_rows);
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 1024));
  if (size_t(base - static_cast<uint8_t*>(ws)) + L.bytes > ws_bytes)
    return fail(AV1P_ENOMEM, "workspace too small: need %zu bytes", L.bytes + 1024);
  av1p_stage* s = new (std::nothrow) av1p_stage();
  if (!s) return fail(AV1P_ENOMEM, "host allocation failed");
  if (int rc = plan_stage(m, L, base, s)) {
    delete s;
    return rc;
  }
  *out = s;
  return AV1P_OK;
}
extern "C" void av1p_stage_destroy(av1p_stage* s) { delete s; }
extern "C" int av1p_stage_launches_per_forward(const av1p_stage* s) { return s ? int(s->ops.size()) : 0; }

// FGVCModel.forward(x, return_features=True) (scripts/006_train_stage3_ab_fgvc.py:277-296): the next forwards of this stage
// also write the L2-normalised 512-wide features (fp32 [rows][512], caller-owned); nullptr switches it off.
extern "C" int av1p_stage_set_features_out(av1p_stage* s, float* features_dev) {
  if (!s) return fail(AV1P_EINVAL, "null stage");
  bool has_tail = false;
  for (const PlannedOp& P : s->ops) has_tail |= P.type == AV1P_OP_FGVC_TAIL;
  if (features_dev && !has_tail) return fail(AV1P_EINVAL, "this model has no FGVC tail: no features to return");
  s->features_out = features_dev;
  return AV1P_OK;
}

extern "C" int av1p_stage_forward(av1p_stage* s, const av1p_input* in, const int32_t* idx_dev, const int32_t* n_dev,
                                  int32_t n, float* logits_dev, void* stream) {
  if (!s || !logits_dev) return fail(AV1P_EINVAL, "null argument");
  StemInput si;
  if (int rc = convert_input(in, &si, s->model ? s->model->block : 16)) return rc;
  return run_stage(s, si, idx_dev, n_dev, n, logits_dev, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------ routing ABI
namespace {
struct RouteScratch {
  int tile_counts[2 * ROUTE_MAX_TILES];
  unsigned int ticket;
  int pad[63];
};

int launch_route(RouteParams rp, int n, cudaStream_t st) {
  const int tiles = ceil_div(std::max(n, 1), ROUTE_TILE);
  if (tiles > ROUTE_MAX_TILES) return fail(AV1P_EINVAL, "too many rows for one routing call (%d)", n);
  {
    ProfScope ps(PROF_ROUTE, st);
    route_count_kernel<<<tiles, ROUTE_THREADS, 0, st>>>(rp);
  }
  {
    ProfScope ps(PROF_ROUTE, st);
    route_scatter_kernel<<<tiles, ROUTE_THREADS, 0, st>>>(rp);
  }
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}
}  // namespace

extern "C" size_t av1p_route_scratch_bytes(void) { return sizeof(RouteScratch); }

extern "C" int av1p_route_stage1(const float* logits, const int32_t* n_dev, int32_t n, float thr, int32_t* idx,
                                 int32_t* count, uint8_t* l8, int64_t* l64, void* scratch, void* stream) {
  if (!logits || !idx || !count || !scratch || n < 0) return fail(AV1P_EINVAL, "bad argument");
  if (n == 0) return AV1P_OK;
  RouteScratch* sc = static_cast<RouteScratch*>(scratch);
  RouteParams rp{};
  rp.kind = 0;
  rp.logits = logits;
  rp.n_dev = n_dev;
  rp.n = n;
  rp.thr = thr;
  rp.tile_counts = sc->tile_counts;
  rp.ticket = &sc->ticket;
  rp.out_idx0 = idx;
  rp.out_idx1 = idx;   // never written for kind 0
  rp.counts = count;   // counts[1] is written too: caller provides int32[2]
  rp.labels_u8 = l8;
  rp.labels_i64 = reinterpret_cast<long long*>(l64);
  return launch_route(rp, n, static_cast<cudaStream_t>(stream));
}

extern "C" int av1p_route_stage2(const float* logits3, const int32_t* idx_in, const int32_t* n_dev, int32_t n,
                                 int32_t* idx_rect, int32_t* idx_ab, int32_t* counts2, uint8_t* l8, int64_t* l64,
                                 void* scratch, void* stream) {
  if (!logits3 || !idx_in || !idx_rect || !idx_ab || !counts2 || !scratch || n < 0) return fail(AV1P_EINVAL, "bad argument");
  if (n == 0) return AV1P_OK;
  RouteScratch* sc = static_cast<RouteScratch*>(scratch);
  RouteParams rp{};
  rp.kind = 1;
  rp.logits = logits3;
  rp.src = idx_in;
  rp.n_dev = n_dev;
  rp.n = n;
  rp.tile_counts = sc->tile_counts;
  rp.ticket = &sc->ticket;
  rp.out_idx0 = idx_rect;
  rp.out_idx1 = idx_ab;
  rp.counts = counts2;
  rp.labels_u8 = l8;
  rp.labels_i64 = reinterpret_cast<long long*>(l64);
  return launch_route(rp, n, static_cast<cudaStream_t>(stream));
}

static int launch_finalize(const float* logits, int32_t k, int32_t base, const int32_t* idx, const int32_t* n_dev, int32_t n,
                           uint8_t* l8, int64_t* l64, int use_softmax, void* stream) {
  if (!logits || !idx || k < 1 || k > FINALIZE_MAX_K || n < 0) return fail(AV1P_EINVAL, "bad argument");
  if (n == 0) return AV1P_OK;
  if (int rc = ensure_ctx()) return rc;
  const int grid = std::min(ceil_div(n, 256), g_ctx.sms * 8);
  {
    ProfScope ps(PROF_FINALIZE, static_cast<cudaStream_t>(stream));
    finalize_labels_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, k, base, idx, n_dev, n, l8, reinterpret_cast<long long*>(l64), use_softmax);
  }
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}
extern "C" int av1p_finalize_labels(const float* logits, int32_t k, int32_t base, const int32_t* idx,
                                    const int32_t* n_dev, int32_t n, uint8_t* l8, int64_t* l64, void* stream) {
  return launch_finalize(logits, k, base, idx, n_dev, n, l8, l64, 1, stream);
}
extern "C" int av1p_finalize_labels_argmax(const float* logits, int32_t k, int32_t base, const int32_t* idx,
                                           const int32_t* n_dev, int32_t n, uint8_t* l8, int64_t* l64, void* stream) {
  return launch_finalize(logits, k, base, idx, n_dev, n, l8, l64, 0, stream);
}

extern "C" int av1p_threshold_sweep(const float* logits, const uint8_t* labels, int32_t n, const double* thresholds_host,
                                    int32_t n_thr, float* probs, uint64_t* counts, void* stream) {
  if (!logits || !labels || !thresholds_host || !counts || n < 0 || n_thr < 1 || n_thr > SWEEP_MAX_T)
    return fail(AV1P_EINVAL, "bad argument (1..%d thresholds per call)", SWEEP_MAX_T);
  if (int rc = ensure_ctx()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemsetAsync(counts, 0, size_t(n_thr) * 4 * sizeof(uint64_t), st));
  if (n == 0) return AV1P_OK;
  SweepParams sp{};
  sp.logits = logits;
  sp.labels = labels;
  sp.n = n;
  sp.n_thr = n_thr;
  for (int i = 0; i < n_thr; ++i) sp.thr[i] = thresholds_host[i];
  sp.probs = probs;
  sp.counts = reinterpret_cast<unsigned long long*>(counts);
  const int grid = std::min(ceil_div(n, 256), g_ctx.sms * 8);
  threshold_sweep_kernel<<<grid, 256, 0, st>>>(sp);
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}

// ------------------------------------------------------------------------------ ensemble voting ABI
extern "C" int av1p_ensemble_vote(const float* logits, int32_t n_models, int32_t n, int32_t k, int32_t mode, const float* weights,
                                  int64_t* pred, float* conf, float* mean_probs, float* std_probs, float* agreement,
                                  float* all_probs, void* stream) {
  if (n < 0 || n_models < 1 || n_models > ENS_MAX_M || k < 1 || k > ENS_MAX_K || mode < 0 || mode > 2 || (mode == 2 && !weights))
    return fail(AV1P_EINVAL, "bad argument (1..%d models, 1..%d classes, mode 0/1/2, weights for mode 2)", ENS_MAX_M, ENS_MAX_K);
  if (int rc = ensure_ctx()) return rc;
  if (n == 0) return AV1P_OK;             // an empty batch has no buffers to check
  if (!logits || !pred) return fail(AV1P_EINVAL, "null logits / predictions");
  EnsembleParams ep{};
  ep.logits = logits;
  ep.weights = weights;
  ep.n_models = n_models;
  ep.n = n;
  ep.k = k;
  ep.mode = mode;
  ep.pred = reinterpret_cast<long long*>(pred);
  ep.conf = conf;
  ep.mean_probs = mean_probs;
  ep.std_probs = std_probs;
  ep.agreement = agreement;
  ep.all_probs = all_probs;
  const int grid = std::min(ceil_div(n, 256), g_ctx.sms * 8);
  ensemble_vote_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ep);
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}

// ------------------------------------------------------------------------------ extraction ABI
template <typename OUT>
static int launch_extract(const uint16_t* y, int w, int h, int pitch, int bs, OUT* out, void* stream, int n_frames = 1,
                          long long frame_stride = 0) {
  if (!y || !out || w <= 0 || h <= 0 || pitch < w || n_frames < 1 || (n_frames > 1 && frame_stride < (long long)pitch * (h - 1) + w))
    return fail(AV1P_EINVAL, "bad frame geometry");
  if (bs != 8 && bs != 16 && bs != 32 && bs != 64) return fail(AV1P_EINVAL, "unsupported block size %d", bs);
  if (int rc = ensure_ctx()) return rc;
  const int bx = ceil_div(w, bs), by = ceil_div(h, bs);
  if (frame_stride == 0) frame_stride = (long long)pitch * h;
  // AV1P_EXTRACT_TMA=1: the TMA-staged kernel (needs the tensor map's alignment rules: 16-byte base and strides).  Both are
  // device paths; the plain vectorised-load kernel is the default because it measured faster (32 4K frames in one launch:
  // 3.94 vs 3.58 TB/s for float output, 4.60 vs 4.16 TB/s for uint16 - the TMA version pays a CTA barrier per box and the
  // kernel is bound by its scattered 32-byte stores either way).  Read per call so that a test can exercise both.
  const char* tma_env = getenv("AV1P_EXTRACT_TMA");
  const bool use_tma = tma_env && atoi(tma_env) != 0;
  if (use_tma && (reinterpret_cast<uintptr_t>(y) & 15u) == 0 && pitch % 8 == 0 && (n_frames == 1 || frame_stride % 8 == 0) && w >= 1) {
    CUtensorMap map;
    cuuint64_t dims[3] = {cuuint64_t(w), cuuint64_t(h), cuuint64_t(n_frames)};
    cuuint64_t strides[2] = {cuuint64_t(pitch) * 2, cuuint64_t(n_frames == 1 ? (long long)pitch * h : frame_stride) * 2};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_ctx.encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(y), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
      const int tpb = 64 / bs;
      const long long tiles = (long long)n_frames * ceil_div(by, tpb) * ceil_div(bx, tpb);
      const int grid = int(std::min<long long>(tiles, (long long)g_ctx.sms * 4));
      extract_blocks_tma_kernel<OUT><<<grid, EX_THREADS, EX_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(map, bs, bx, by, n_frames, out);
      CUDA_TRY(cudaGetLastError());
      return AV1P_OK;
    }
  }
  const long long chunks = (long long)bx * bs / 8 * by * bs * n_frames;
  const int grid = int(std::min<long long>((chunks + 255) / 256, (long long)g_ctx.sms * 16));
  extract_blocks_kernel<OUT><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, w, h, pitch, bs, bx, by, out, n_frames, frame_stride);
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}
extern "C" int av1p_extract_u16(const uint16_t* y, int32_t w, int32_t h, int32_t pitch, int32_t bs, uint16_t* out, void* stream) {
  return launch_extract<uint16_t>(y, w, h, pitch, bs, out, stream);
}
extern "C" int av1p_extract_norm_u16(const uint16_t* y, int32_t w, int32_t h, int32_t pitch, int32_t bs, float* out, void* stream) {
  return launch_extract<float>(y, w, h, pitch, bs, out, stream);
}
extern "C" int av1p_extract_frames_u16(const uint16_t* frames, int32_t n_frames, int64_t frame_stride, int32_t w, int32_t h,
                                       int32_t pitch, int32_t bs, uint16_t* out, void* stream) {
  return launch_extract<uint16_t>(frames, w, h, pitch, bs, out, stream, n_frames, frame_stride);
}
extern "C" int av1p_extract_frames_norm_u16(const uint16_t* frames, int32_t n_frames, int64_t frame_stride, int32_t w, int32_t h,
                                            int32_t pitch, int32_t bs, float* out, void* stream) {
  return launch_extract<float>(frames, w, h, pitch, bs, out, stream, n_frames, frame_stride);
}

// ------------------------------------------------------------------------------ cascade
struct av1p_cascade {
  av1p_stage stage[4];
  int cap = 0;
  float* logits[4] = {nullptr, nullptr, nullptr, nullptr};
  int32_t* idx2 = nullptr;
  int32_t* idx_rect = nullptr;
  int32_t* idx_ab = nullptr;
  int32_t* counts = nullptr;     // [0]=n2, [1]=scratch, [2]=nR, [3]=nA
  RouteScratch* scratch = nullptr;
  // The two stage-3 specialists are independent (008:105-125): with a second activation region the AB network runs on a
  // side stream next to the RECT network, so one stage's last (partial) wave and launch gaps overlap the other's work.
  bool overlap3 = false;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // Speculative small-batch path (av1p_cascade_predict, n <= spec_cap): stages 2 / RECT / AB get their own small activation
  // regions and streams and run on EVERY block next to stage 1; their logits are compacted by the routing lists afterwards.
  int spec_cap = 0;
  av1p_stage spec_stage[3];      // stage 2, RECT, AB planned on the speculative regions
  float* spec_logits[3] = {nullptr, nullptr, nullptr};   // full (uncompacted) logits [spec_cap][3 | 2 | 4]
  cudaStream_t spec_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t spec_fork = nullptr, spec_join[3] = {nullptr, nullptr, nullptr};
  ~av1p_cascade() {
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
    if (spec_fork) cudaEventDestroy(spec_fork);
    for (int i = 0; i < 3; ++i) {
      if (spec_join[i]) cudaEventDestroy(spec_join[i]);
      if (spec_stream[i]) cudaStreamDestroy(spec_stream[i]);
    }
  }
};

namespace {
constexpr int SPEC_MAX_BLOCKS = 4096;      // 32 M tiles per stage: four stages side by side still fit one wave of 148 SMs
inline int spec_max_blocks() {             // AV1P_SPEC_MAX overrides (experiments; read when a cascade is sized / created)
  const char* e = getenv("AV1P_SPEC_MAX");
  const int v = e ? atoi(e) : SPEC_MAX_BLOCKS;
  return std::max(1, std::min(v, 1 << 16));
}
struct CascadeLayout {
  ActLayout act;
  size_t act2_off;               // second activation region (0 = none: the specialists run back to back)
  size_t logits_off[4], idx_off[3], counts_off, scratch_off, bytes;
  ActLayout spec;                // activation layout of one speculative region (capacity spec.cap rows)
  size_t spec_off[3], spec_logits_off[3];
};
// second region only while it stays small next to 180 GB of HBM (AV1P_STAGE3_OVERLAP=0 disables)
inline bool stage3_overlap_wanted(size_t act_bytes) {
  const char* e = getenv("AV1P_STAGE3_OVERLAP");
  if (e && atoi(e) == 0) return false;
  return act_bytes <= (size_t(24) << 30);
}
CascadeLayout make_cascade_layout(const av1p_model* const models[4], int capacity) {
  CascadeLayout C;
  C.act = make_act_layout(models, 4, capacity);
  size_t o = C.act.bytes;
  C.act2_off = 0;
  if (stage3_overlap_wanted(C.act.bytes)) {
    C.act2_off = align_up(o, 1024);
    o = C.act2_off + C.act.bytes;
  }
  const int outs[4] = {1, 3, 2, 4};
  for (int i = 0; i < 4; ++i) {
    C.logits_off[i] = o;
    o += align_up(size_t(C.act.cap) * outs[i] * 4, 1024);
  }
  for (int i = 0; i < 3; ++i) {
    C.idx_off[i] = o;
    o += align_up(size_t(C.act.cap) * 4, 1024);
  }
  // speculative small-batch regions (12.4 KB per row: 3 x 51 MB at the full 4,096 rows)
  C.spec = make_act_layout(models, 4, std::min(capacity, spec_max_blocks()));
  for (int i = 0; i < 3; ++i) {
    C.spec_off[i] = align_up(o, 1024);
    o = C.spec_off[i] + C.spec.bytes;
    C.spec_logits_off[i] = o;
    o += align_up(size_t(C.spec.cap) * outs[i + 1] * 4, 1024);
  }
  C.counts_off = o;
  o += 1024;
  C.scratch_off = o;
  o += align_up(sizeof(RouteScratch), 1024);
  C.bytes = o;
  return C;
}
}  // namespace

extern "C" size_t av1p_cascade_workspace_bytes(const av1p_model* const models[4], int32_t capacity) {
  if (!models || capacity <= 0) return 0;
  for (int i = 0; i < 4; ++i)
    if (!models[i]) return 0;
  return make_cascade_layout(models, capacity).bytes + 1024;
}

extern "C" int av1p_cascade_create(const av1p_model* const models[4], int32_t capacity, void* ws, size_t ws_bytes,
                                   av1p_cascade** out) {
  if (!models || !ws || !out || capacity <= 0) return fail(AV1P_EINVAL, "bad argument");
  const int outs[4] = {1, 3, 2, 4};
  for (int i = 0; i < 4; ++i) {
    if (!models[i]) return fail(AV1P_EINVAL, "null model %d", i);
    if (int(models[i]->hdr.n_out) != outs[i])
      return fail(AV1P_EINVAL, "model %d has %u outputs, the cascade needs %d", i, models[i]->hdr.n_out, outs[i]);
    if (models[i]->block != models[0]->block)
      return fail(AV1P_EINVAL, "the models of a cascade must be packed for one block size (%d vs %d)", models[i]->block, models[0]->block);
  }
  if (int rc = ensure_ctx()) return rc;
  CascadeLayout C = make_cascade_layout(models, capacity);
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 1024));
  if (size_t(base - static_cast<uint8_t*>(ws)) + C.bytes > ws_bytes)
    return fail(AV1P_ENOMEM, "workspace too small: need %zu bytes", C.bytes + 1024);
  av1p_cascade* c = new (std::nothrow) av1p_cascade();
  if (!c) return fail(AV1P_ENOMEM, "host allocation failed");
  c->cap = C.act.cap;
  c->overlap3 = C.act2_off != 0;
  for (int i = 0; i < 4; ++i) {
    uint8_t* region = (i == 3 && c->overlap3) ? base + C.act2_off : base;
    if (int rc = plan_stage(models[i], C.act, region, &c->stage[i])) {
      delete c;
      return rc;
    }
    c->logits[i] = reinterpret_cast<float*>(base + C.logits_off[i]);
    c->stage[i].range_flag = reinterpret_cast<int*>(base + C.counts_off) + 8;
  }
  if (c->overlap3) {
    if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      delete c;
      return fail(AV1P_ECUDA, "side stream / events for the stage-3 overlap could not be created");
    }
  }
  c->spec_cap = C.spec.cap;
  for (int i = 0; i < 3; ++i) {
    if (int rc = plan_stage(models[i + 1], C.spec, base + C.spec_off[i], &c->spec_stage[i])) {
      delete c;
      return rc;
    }
    c->spec_stage[i].range_flag = reinterpret_cast<int*>(base + C.counts_off) + 8;
    c->spec_logits[i] = reinterpret_cast<float*>(base + C.spec_logits_off[i]);
    if (cudaStreamCreateWithFlags(&c->spec_stream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->spec_join[i], cudaEventDisableTiming) != cudaSuccess) {
      delete c;
      return fail(AV1P_ECUDA, "streams / events of the speculative small-batch path could not be created");
    }
  }
  if (cudaEventCreateWithFlags(&c->spec_fork, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return fail(AV1P_ECUDA, "fork event of the speculative small-batch path could not be created");
  }
  c->idx2 = reinterpret_cast<int32_t*>(base + C.idx_off[0]);
  c->idx_rect = reinterpret_cast<int32_t*>(base + C.idx_off[1]);
  c->idx_ab = reinterpret_cast<int32_t*>(base + C.idx_off[2]);
  c->counts = reinterpret_cast<int32_t*>(base + C.counts_off);
  c->scratch = reinterpret_cast<RouteScratch*>(base + C.scratch_off);
  cudaError_t e = cudaMemset(base + C.counts_off, 0, C.bytes - C.counts_off);
  if (e != cudaSuccess) {
    delete c;
    return fail(AV1P_ECUDA, "workspace init failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  return AV1P_OK;
}
extern "C" void av1p_cascade_destroy(av1p_cascade* c) { delete c; }

extern "C" int av1p_cascade_launches_per_predict(const av1p_cascade* c) {
  if (!c) return 0;
  int n = 0;
  for (int i = 0; i < 4; ++i) n += int(c->stage[i].ops.size());
  return n + 2 /*route1*/ + 2 /*route2*/ + 2 /*finalize x2*/;
}

extern "C" const void* av1p_cascade_buffer(const av1p_cascade* c, int32_t which) {
  if (!c) return nullptr;
  switch (which) {
    case 0: case 1: case 2: case 3: return c->logits[which];
    case 4: return c->idx2;
    case 5: return c->idx_rect;
    case 6: return c->idx_ab;
    case 7: return c->counts;
    case 8: return c->counts + 8;      // int32: 1 once a frame sample above 2048 was met (the caller zeroes / reads it in stream order)
  }
  return nullptr;
}

// HierarchicalPipelineV6.predict (008:69-127) without its three host synchronisations: every
// downstream kernel reads its row count from the device counters the routing kernels wrote.
extern "C" int av1p_cascade_predict(av1p_cascade* c, const av1p_input* in, int32_t n_blocks, float thr, uint8_t* l8,
                                    int64_t* l64, void* stream) {
  if (!c) return fail(AV1P_EINVAL, "null cascade");
  if (n_blocks < 0 || n_blocks > c->cap) return fail(AV1P_EINVAL, "n_blocks=%d outside [0, %d]", n_blocks, c->cap);
  if (n_blocks == 0) return AV1P_OK;
  StemInput si;
  if (int rc = convert_input(in, &si, c->stage[0].model->block)) return rc;
  if (in->kind == 0 && (long long)si.blocks_per_frame * in->n_frames < n_blocks)
    return fail(AV1P_EINVAL, "n_blocks exceeds the blocks in the given frames");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* n2 = c->counts + 0;
  int32_t* n_rect = c->counts + 2;
  int32_t* n_ab = c->counts + 3;
  if (g_ctx.speculate && n_blocks <= c->spec_cap) {
    // Small batch (what the reference's evaluate_pipeline feeds: 256 blocks per call, 008:278-284): a stage forward on a
    // few M tiles occupies a handful of SMs for ~0.4 ms of dependent launches, and the routed cascade runs four of them
    // one after the other.  Here stages 2 / RECT / AB run on EVERY block on three side streams next to stage 1 (their own
    // activation regions), then the usual routing kernels run on logits compacted by the index lists they produce.  A row's
    // logits do not depend on which other rows share its batch (the kernels are row-independent: bitwise, see the block
    // independence test), so logits, index lists, counts and labels are exactly those of the routed path.
    CUDA_TRY(cudaEventRecord(c->spec_fork, st));
    int rc = AV1P_OK;
    for (int i = 0; i < 3 && !rc; ++i) {
      CUDA_TRY(cudaStreamWaitEvent(c->spec_stream[i], c->spec_fork, 0));
      rc = run_stage(&c->spec_stage[i], si, nullptr, nullptr, n_blocks, c->spec_logits[i], c->spec_stream[i]);
    }
    if (!rc) rc = run_stage(&c->stage[0], si, nullptr, nullptr, n_blocks, c->logits[0], st);
    for (int i = 0; i < 3; ++i) {      // always join, also after an error
      cudaEventRecord(c->spec_join[i], c->spec_stream[i]);
      cudaStreamWaitEvent(st, c->spec_join[i], 0);
    }
    if (rc) return rc;
    auto gather = [&](const float* src, const int32_t* idx, const int32_t* n_dev, int k, float* dst) -> int {
      gather_rows_kernel<<<std::min(ceil_div(n_blocks * k, 256), g_ctx.sms * 4), 256, 0, st>>>(src, idx, n_dev, n_blocks, k, dst);
      CUDA_TRY(cudaGetLastError());
      return AV1P_OK;
    };
    if ((rc = av1p_route_stage1(c->logits[0], nullptr, n_blocks, thr, c->idx2, n2, l8, l64, c->scratch, st))) return rc;
    if ((rc = gather(c->spec_logits[0], c->idx2, n2, 3, c->logits[1]))) return rc;
    if ((rc = av1p_route_stage2(c->logits[1], c->idx2, n2, n_blocks, c->idx_rect, c->idx_ab, n_rect, l8, l64, c->scratch, st))) return rc;
    if ((rc = gather(c->spec_logits[1], c->idx_rect, n_rect, 2, c->logits[2]))) return rc;
    if ((rc = av1p_finalize_labels(c->logits[2], 2, 2, c->idx_rect, n_rect, n_blocks, l8, l64, st))) return rc;
    if ((rc = gather(c->spec_logits[2], c->idx_ab, n_ab, 4, c->logits[3]))) return rc;
    return av1p_finalize_labels(c->logits[3], 4, 4, c->idx_ab, n_ab, n_blocks, l8, l64, st);
  }
  // Stage 1 on every block, then threshold + compaction (008:76-85)
  if (int rc = run_stage(&c->stage[0], si, nullptr, nullptr, n_blocks, c->logits[0], st)) return rc;
  if (int rc = av1p_route_stage1(c->logits[0], nullptr, n_blocks, thr, c->idx2, n2, l8, l64, c->scratch, st)) return rc;
  // Stage 2 on the routed blocks, argmax + 3-way partition (008:91-103, 115-116)
  if (int rc = run_stage(&c->stage[1], si, c->idx2, n2, n_blocks, c->logits[1], st)) return rc;
  if (int rc = av1p_route_stage2(c->logits[1], c->idx2, n2, n_blocks, c->idx_rect, c->idx_ab, n_rect, l8, l64,
                                 c->scratch, st))
    return rc;
  // Stage 3 specialists (008:105-125): independent networks on disjoint blocks - AB on the side stream when it has its own
  // activation region (fork / join with events, which also works under stream capture)
  cudaStream_t st_ab = st;
  if (c->overlap3) {
    CUDA_TRY(cudaEventRecord(c->ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    st_ab = c->side;
  }
  int rc = run_stage(&c->stage[2], si, c->idx_rect, n_rect, n_blocks, c->logits[2], st);
  if (!rc) rc = av1p_finalize_labels(c->logits[2], 2, 2, c->idx_rect, n_rect, n_blocks, l8, l64, st);
  if (!rc) rc = run_stage(&c->stage[3], si, c->idx_ab, n_ab, n_blocks, c->logits[3], st_ab);
  if (!rc) rc = av1p_finalize_labels(c->logits[3], 4, 4, c->idx_ab, n_ab, n_blocks, l8, l64, st_ab);
  if (c->overlap3) {        // always join, also after an error, so that the side stream never outlives the call's ordering
    cudaEventRecord(c->ev_join, c->side);
    cudaStreamWaitEvent(st, c->ev_join, 0);
  }
  return rc;
}

// ------------------------------------------------------------------------------ flatten cascade
// Stage 1 -> 7-way Stage2FlatModel (reference scripts/008b_run_pipeline_flatten_eval.py:177-229): blocks whose
// stage-1 probability reaches the threshold get label 1 + argmax of the 7 flat logits, the others label 0.
struct av1p_flat_cascade {
  av1p_stage stage[2];
  int cap = 0;
  float* logits[2] = {nullptr, nullptr};
  int32_t* idx2 = nullptr;
  int32_t* counts = nullptr;
  RouteScratch* scratch = nullptr;
  // speculative small-batch path, as in av1p_cascade: the flat model on every block on a side stream next to stage 1
  int spec_cap = 0;
  av1p_stage spec_stage;
  float* spec_logits = nullptr;
  cudaStream_t spec_stream = nullptr;
  cudaEvent_t spec_fork = nullptr, spec_join = nullptr;
  ~av1p_flat_cascade() {
    if (spec_fork) cudaEventDestroy(spec_fork);
    if (spec_join) cudaEventDestroy(spec_join);
    if (spec_stream) cudaStreamDestroy(spec_stream);
  }
};

namespace {
struct FlatLayout {
  ActLayout act;
  size_t logits_off[2], idx_off, counts_off, scratch_off, bytes;
  ActLayout spec;
  size_t spec_off, spec_logits_off;
};
FlatLayout make_flat_layout(const av1p_model* const models[2], int capacity) {
  FlatLayout C;
  C.act = make_act_layout(models, 2, capacity);
  size_t o = C.act.bytes;
  const int outs[2] = {1, 7};
  for (int i = 0; i < 2; ++i) {
    C.logits_off[i] = o;
    o += align_up(size_t(C.act.cap) * outs[i] * 4, 1024);
  }
  C.idx_off = o;
  o += align_up(size_t(C.act.cap) * 4, 1024);
  C.spec = make_act_layout(models, 2, std::min(capacity, spec_max_blocks()));
  C.spec_off = align_up(o, 1024);
  o = C.spec_off + C.spec.bytes;
  C.spec_logits_off = o;
  o += align_up(size_t(C.spec.cap) * 7 * 4, 1024);
  C.counts_off = o;
  o += 1024;
  C.scratch_off = o;
  o += align_up(sizeof(RouteScratch), 1024);
  C.bytes = o;
  return C;
}
}  // namespace

extern "C" size_t av1p_flat_cascade_workspace_bytes(const av1p_model* const models[2], int32_t capacity) {
  if (!models || !models[0] || !models[1] || capacity <= 0) return 0;
  return make_flat_layout(models, capacity).bytes + 1024;
}

extern "C" int av1p_flat_cascade_create(const av1p_model* const models[2], int32_t capacity, void* ws, size_t ws_bytes,
                                        av1p_flat_cascade** out) {
  if (!models || !models[0] || !models[1] || !ws || !out || capacity <= 0) return fail(AV1P_EINVAL, "bad argument");
  if (models[0]->hdr.n_out != 1 || models[1]->hdr.n_out != 7)
    return fail(AV1P_EINVAL, "the flatten cascade needs a 1-output stage-1 model and a 7-output flat model");
  if (models[0]->block != models[1]->block) return fail(AV1P_EINVAL, "the models of a cascade must be packed for one block size");
  if (int rc = ensure_ctx()) return rc;
  FlatLayout C = make_flat_layout(models, capacity);
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 1024));
  if (size_t(base - static_cast<uint8_t*>(ws)) + C.bytes > ws_bytes)
    return fail(AV1P_ENOMEM, "workspace too small: need %zu bytes", C.bytes + 1024);
  av1p_flat_cascade* c = new (std::nothrow) av1p_flat_cascade();
  if (!c) return fail(AV1P_ENOMEM, "host allocation failed");
  c->cap = C.act.cap;
  for (int i = 0; i < 2; ++i) {
    if (int rc = plan_stage(models[i], C.act, base, &c->stage[i])) {
      delete c;
      return rc;
    }
    c->logits[i] = reinterpret_cast<float*>(base + C.logits_off[i]);
    c->stage[i].range_flag = reinterpret_cast<int*>(base + C.counts_off) + 8;
  }
  c->spec_cap = C.spec.cap;
  if (int rc = plan_stage(models[1], C.spec, base + C.spec_off, &c->spec_stage)) {
    delete c;
    return rc;
  }
  c->spec_stage.range_flag = reinterpret_cast<int*>(base + C.counts_off) + 8;
  c->spec_logits = reinterpret_cast<float*>(base + C.spec_logits_off);
  if (cudaStreamCreateWithFlags(&c->spec_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->spec_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->spec_join, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return fail(AV1P_ECUDA, "stream / events of the speculative small-batch path could not be created");
  }
  c->idx2 = reinterpret_cast<int32_t*>(base + C.idx_off);
  c->counts = reinterpret_cast<int32_t*>(base + C.counts_off);
  c->scratch = reinterpret_cast<RouteScratch*>(base + C.scratch_off);
  cudaError_t e = cudaMemset(base + C.counts_off, 0, C.bytes - C.counts_off);
  if (e != cudaSuccess) {
    delete c;
    return fail(AV1P_ECUDA, "workspace init failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  return AV1P_OK;
}
extern "C" void av1p_flat_cascade_destroy(av1p_flat_cascade* c) { delete c; }

extern "C" const void* av1p_flat_cascade_buffer(const av1p_flat_cascade* c, int32_t which) {
  if (!c) return nullptr;
  switch (which) {
    case 0: case 1: return c->logits[which];
    case 2: return c->idx2;
    case 3: return c->counts;
    case 4: return c->counts + 8;      // input-range flag, see av1p_cascade_buffer(c, 8)
  }
  return nullptr;
}

extern "C" int av1p_flat_cascade_predict(av1p_flat_cascade* c, const av1p_input* in, int32_t n_blocks, float thr, uint8_t* l8,
                                         int64_t* l64, void* stream) {
  if (!c) return fail(AV1P_EINVAL, "null cascade");
  if (n_blocks < 0 || n_blocks > c->cap) return fail(AV1P_EINVAL, "n_blocks=%d outside [0, %d]", n_blocks, c->cap);
  if (n_blocks == 0) return AV1P_OK;
  StemInput si;
  if (int rc = convert_input(in, &si, c->stage[0].model->block)) return rc;
  if (in->kind == 0 && (long long)si.blocks_per_frame * in->n_frames < n_blocks)
    return fail(AV1P_EINVAL, "n_blocks exceeds the blocks in the given frames");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* n2 = c->counts;
  if (g_ctx.speculate && n_blocks <= c->spec_cap) {
    // small batch: the flat model runs on every block on the side stream while stage 1 runs here (see av1p_cascade_predict)
    CUDA_TRY(cudaEventRecord(c->spec_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(c->spec_stream, c->spec_fork, 0));
    int rc = run_stage(&c->spec_stage, si, nullptr, nullptr, n_blocks, c->spec_logits, c->spec_stream);
    if (!rc) rc = run_stage(&c->stage[0], si, nullptr, nullptr, n_blocks, c->logits[0], st);
    cudaEventRecord(c->spec_join, c->spec_stream);
    cudaStreamWaitEvent(st, c->spec_join, 0);
    if (rc) return rc;
    if ((rc = av1p_route_stage1(c->logits[0], nullptr, n_blocks, thr, c->idx2, n2, l8, l64, c->scratch, st))) return rc;
    gather_rows_kernel<<<std::min(ceil_div(n_blocks * 7, 256), g_ctx.sms * 4), 256, 0, st>>>(c->spec_logits, c->idx2, n2, n_blocks, 7, c->logits[1]);
    CUDA_TRY(cudaGetLastError());
    return av1p_finalize_labels_argmax(c->logits[1], 7, 1, c->idx2, n2, n_blocks, l8, l64, st);
  }
  if (int rc = run_stage(&c->stage[0], si, nullptr, nullptr, n_blocks, c->logits[0], st)) return rc;
  if (int rc = av1p_route_stage1(c->logits[0], nullptr, n_blocks, thr, c->idx2, n2, l8, l64, c->scratch, st)) return rc;
  if (int rc = run_stage(&c->stage[1], si, c->idx2, n2, n_blocks, c->logits[1], st)) return rc;
  return av1p_finalize_labels_argmax(c->logits[1], 7, 1, c->idx2, n2, n_blocks, l8, l64, st);
}

// ------------------------------------------------------------------------------ FC test hook
extern "C" int av1p_fc_forward(const av1p_fc_desc* d, void* stream) {
  if (!d || !d->a_dev[0] || !d->w_dev || !d->kb_begin || !d->kb_src || !d->kb_w) return fail(AV1P_EINVAL, "null argument");
  if (int rc = ensure_ctx()) return rc;
  if (d->n_tiles < 1 || d->n_tiles > FC_MAX_NT || d->block_n < 32 || d->block_n > FC_MAX_N || d->block_n % 32 ||
      d->n_kb_total < 1 || d->n_kb_total > FC_MAX_KB || d->rows < 1 || d->n_w_chunks < 1)
    return fail(AV1P_EINVAL, "bad FC shape");
  FcParams f;
  memset(&f, 0, sizeof f);
  for (int i = 0; i < FC_MAX_SRC; ++i) {
    const int j = d->a_dev[i] ? i : 0;
    if (int rc = make_act_map(&f.a_map[i], d->a_dev[j], uint64_t(d->a_cols[j]), uint64_t(d->rows), false)) return rc;
    f.src_kb[i] = d->a_cols[j] / 64;
  }
  if (int rc = make_map_2d(&f.w_map, d->w_dev, 64, uint64_t(d->n_w_chunks) * d->block_n, 64, d->block_n)) return rc;
  if (int rc = make_map_2d(&f.w_half_map, d->w_dev, 64, uint64_t(d->n_w_chunks) * d->block_n, 64, d->block_n / 2)) return rc;
  f.n_rows_dev = d->n_dev;
  f.n_rows = d->rows;
  f.n_tiles = d->n_tiles;
  f.block_n = d->block_n;
  f.epi = d->epi;
  f.bias = d->bias_dev;
  f.row_scale = d->row_scale_dev;
  f.acc_scale = d->acc_scale;
  f.pair_mode = d->pair_mode;
  f.aux = static_cast<const __half*>(d->aux_dev);
  f.aux_lo = static_cast<const __half*>(d->aux_lo_dev);
  f.aux_kb = d->aux_ld / 64;
  f.out = static_cast<__half*>(d->out_dev);
  f.out_lo = static_cast<__half*>(d->out_lo_dev);
  f.out_kb = d->out_ld / 64;
  f.tail_w = d->tail_w_dev;
  f.tail_b = d->tail_b_dev;
  f.logits = d->logits_dev;
  f.tail_n = d->tail_n;
  f.err_flag = g_ctx.watchdog_dev;
  for (int i = 0; i <= d->n_tiles; ++i) f.kb_begin[i] = d->kb_begin[i];
  for (int i = 0; i < d->n_kb_total; ++i) {
    f.kb_src[i] = d->kb_src[i];
    f.kb_w[i] = d->kb_w[i];
  }
  if (d->epi != FC_EPI_HEAD) {
    if (!d->out_dev || d->out_ld < d->n_tiles * d->block_n) return fail(AV1P_EINVAL, "FC output missing or too narrow");
    if (int rc = make_act_map(&f.out_map[0], d->out_dev, uint64_t(d->out_ld), uint64_t(d->rows), true)) return rc;
    if (d->out_lo_dev)
      if (int rc = make_act_map(&f.out_map[1], d->out_lo_dev, uint64_t(d->out_ld), uint64_t(d->rows), true)) return rc;
  }
  if (d->epi == FC_EPI_GATE || d->epi == FC_EPI_ADD) {
    if (!d->aux_dev || d->aux_ld < d->n_tiles * d->block_n) return fail(AV1P_EINVAL, "gate input missing or too narrow");
    if (int rc = make_act_map(&f.aux_map[0], d->aux_dev, uint64_t(d->aux_ld), uint64_t(d->rows), true)) return rc;
    if (d->aux_lo_dev)
      if (int rc = make_act_map(&f.aux_map[1], d->aux_lo_dev, uint64_t(d->aux_ld), uint64_t(d->rows), true)) return rc;
  }
  if (d->epi == FC_EPI_ADD_RELU && g_ctx.fc_resid_epi) {
    if (!d->aux_dev || d->aux_ld < d->n_tiles * d->block_n) return fail(AV1P_EINVAL, "residual missing or too narrow");
    if (f.row_scale) return fail(AV1P_EINVAL, "residual FC layer cannot use a row scale");
    f.aux_epi = 1;
    if (int rc = make_act_map(&f.aux_map[0], d->aux_dev, uint64_t(d->aux_ld), uint64_t(d->rows), true)) return rc;
    if (d->aux_lo_dev)
      if (int rc = make_act_map(&f.aux_map[1], d->aux_lo_dev, uint64_t(d->aux_ld), uint64_t(d->rows), true)) return rc;
  } else if (d->epi == FC_EPI_ADD_RELU) {
    if (!d->aux_dev || d->aux_ld < d->n_tiles * d->block_n) return fail(AV1P_EINVAL, "residual missing or too narrow");
    if (int rc = make_act_map(&f.a_map[2], d->aux_dev, uint64_t(d->aux_ld), uint64_t(d->rows), false)) return rc;
    f.src_kb[2] = f.src_kb[3] = d->aux_ld / 64;
    if (d->aux_lo_dev)
      if (int rc = make_act_map(&f.a_map[3], d->aux_lo_dev, uint64_t(d->aux_ld), uint64_t(d->rows), false)) return rc;
    if (int rc = add_residual_entries(f, d->aux_lo_dev != nullptr)) return rc;
  }
  if (int rc = launch_fc(f, d->rows, static_cast<cudaStream_t>(stream))) return rc;
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}

// ------------------------------------------------------------------------------ resident-conv test hook
extern "C" int av1p_conv_res_forward(const av1p_conv_res_desc* d, void* stream) {
  if (!d || !d->x_dev || !d->w_dev || !d->out_dev || !d->bias_dev || d->rows < 1) return fail(AV1P_EINVAL, "null argument");
  if (d->split && (!d->x_lo_dev || !d->out_lo_dev)) return fail(AV1P_EINVAL, "split precision needs the lo planes");
  if (d->epi < 0 || d->epi > 2 || (d->epi == 2 && !d->aux_dev)) return fail(AV1P_EINVAL, "bad epilogue");
  if (int rc = ensure_ctx()) return rc;
  ConvResParams f;
  memset(&f, 0, sizeof f);
  if (int rc = make_act_map(&f.a_map[0], d->x_dev, 1024, uint64_t(d->rows), false)) return rc;
  if (int rc = make_act_map(&f.a_map[1], d->split ? d->x_lo_dev : d->x_dev, 1024, uint64_t(d->rows), false)) return rc;
  if (int rc = make_map_2d(&f.w_map, d->w_dev, 64, uint64_t(d->split ? 2 : 1) * 9 * 64, 64, 64)) return rc;
  if (int rc = make_act_map(&f.out_map[0], d->out_dev, 1024, uint64_t(d->rows), true)) return rc;
  if (d->out_lo_dev)
    if (int rc = make_act_map(&f.out_map[1], d->out_lo_dev, 1024, uint64_t(d->rows), true)) return rc;
  if (d->epi == FC_EPI_ADD_RELU) {
    if (int rc = make_act_map(&f.a_map[2], d->aux_dev, 1024, uint64_t(d->rows), false)) return rc;
    if (int rc = make_act_map(&f.a_map[3], d->aux_lo_dev ? d->aux_lo_dev : d->aux_dev, 1024, uint64_t(d->rows), false)) return rc;
    f.has_aux_lo = d->aux_lo_dev ? 1 : 0;
    if (g_ctx.cr_resid_epi) {
      f.resid_epi = g_ctx.cr_resid_epi;
      f.aux = static_cast<const __half*>(d->aux_dev);
      f.aux_lo = static_cast<const __half*>(d->aux_lo_dev);
      f.aux_kb = 16;
      if (int rc = make_act_map(&f.res_map[0], d->aux_dev, 1024, uint64_t(d->rows), true)) return rc;
      if (int rc = make_act_map(&f.res_map[1], d->aux_lo_dev ? d->aux_lo_dev : d->aux_dev, 1024, uint64_t(d->rows), true)) return rc;
    }
  }
  if (!(1.0f / d->acc_scale >= 1.0f && 1.0f / d->acc_scale <= 32768.0f)) return fail(AV1P_EINVAL, "weight scale outside [1, 2^15]");
  f.n_rows_dev = d->n_dev;
  f.n_rows = d->rows;
  f.split = d->split ? 1 : 0;
  f.epi = d->epi;
  f.bias = d->bias_dev;
  f.acc_scale = d->acc_scale;
  f.out = static_cast<__half*>(d->out_dev);
  f.out_lo = static_cast<__half*>(d->out_lo_dev);
  f.out_kb = 16;
  f.err_flag = g_ctx.watchdog_dev;
  if (!conv_res_build_schedule(f)) return fail(AV1P_EINVAL, "resident-conv schedule does not fit its tables");
  if (const char* dbg = getenv("AV1P_CR_DEBUG")) f.debug = atoi(dbg);   // kernel-development switch of this test hook only
  const int grid = 2 * std::max(1, std::min(g_ctx.grid_sms / 2, ceil_div(d->rows, FC_TILE_M)));
  conv_res_kernel_for(f)<<<grid, CR_THREADS, CR_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(f);
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}

// ------------------------------------------------------------------------------ training step (BASELINE configs[4])
// The two non-convolution pieces of the data-parallel Stage-1 training step as single launches (train_kernels.cuh).
extern "C" int av1p_focal_loss_binary(const float* logits_dev, const int64_t* targets_dev, int32_t n, float alpha, float gamma,
                                      float* loss_dev, float* dlogits_dev, void* stream) {
  if (!logits_dev || !targets_dev || !loss_dev) return fail(AV1P_EINVAL, "null argument");
  if (n <= 0) return fail(AV1P_EINVAL, "focal loss over %d logits", n);
  if (!(gamma >= 0.f) || !(alpha >= 0.f && alpha <= 1.f)) return fail(AV1P_EINVAL, "focal loss: alpha %g / gamma %g out of range", double(alpha), double(gamma));
  if (int rc = ensure_ctx()) return rc;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 targets");
  focal_loss_binary_kernel<<<1, FOCAL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      logits_dev, reinterpret_cast<const long long*>(targets_dev), n, alpha, gamma, loss_dev, dlogits_dev);
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}

extern "C" int av1p_adamw_flat(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, double lr,
                               double beta1, double beta2, double eps, double weight_decay, double grad_scale, int32_t* step_dev,
                               int32_t advance_step, void* stream) {
  if (!param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || !step_dev) return fail(AV1P_EINVAL, "null argument");
  if (n <= 0) return fail(AV1P_EINVAL, "AdamW over %lld parameters", (long long)n);
  {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(param_dev) & 15u;
    if ((a0 & 3u) || (reinterpret_cast<uintptr_t>(grad_dev) & 15u) != a0 || (reinterpret_cast<uintptr_t>(exp_avg_dev) & 15u) != a0 ||
        (reinterpret_cast<uintptr_t>(exp_avg_sq_dev) & 15u) != a0)
      return fail(AV1P_EINVAL, "AdamW buffers must be 4-byte aligned and share their address modulo 16");
  }
  if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0) || !(lr >= 0.0) || !(weight_decay >= 0.0))
    return fail(AV1P_EINVAL, "AdamW hyper-parameters out of range");
  if (int rc = ensure_ctx()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (advance_step) {
    train_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
    CUDA_TRY(cudaGetLastError());
  }
  AdamWArgs a{param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, (long long)n, lr, beta1, beta2, float(1.0 - lr * weight_decay),
              float(1.0 - beta1), float(beta2), float(1.0 - beta2), float(eps), float(grad_scale), step_dev};
  // HBM-bound streaming pass: a whole number of waves, at most eight resident CTAs of 256 threads per SM
  const long long want = ((n >> 2) + TRAIN_THREADS - 1) / TRAIN_THREADS;
  const int grid = int(std::max<long long>(1, std::min<long long>(want, (long long)g_ctx.sms * 8)));
  adamw_flat_kernel<<<grid, TRAIN_THREADS, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}

// Reduce-scatter + AdamW + all-gather of the data-parallel step in one kernel over peer memory (train_kernels.cuh).
extern "C" int av1p_dp_adamw_fused(const float* const* grad_ptrs, float* const* param_ptrs, int32_t* const* flag_ptrs, int32_t rank,
                                   int32_t world, int64_t n, int64_t shard, float* exp_avg_shard_dev, float* exp_avg_sq_shard_dev,
                                   double lr, double beta1, double beta2, double eps, double weight_decay, int64_t skip_lo,
                                   int64_t skip_hi, int32_t* step_dev, int32_t epoch, int32_t* err_dev, void* stream) {
  if (!grad_ptrs || !param_ptrs || !flag_ptrs || !exp_avg_shard_dev || !exp_avg_sq_shard_dev || !step_dev || !err_dev)
    return fail(AV1P_EINVAL, "null argument");
  if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return fail(AV1P_EINVAL, "rank %d / world %d (at most %d ranks)", rank, world, DP_MAX_WORLD);
  if (n <= 0 || shard <= 0 || shard % 4 || shard * world < n) return fail(AV1P_EINVAL, "shard of %lld elements does not cover %lld over %d ranks (multiple of 4 required)", (long long)shard, (long long)n, world);
  if (skip_lo > skip_hi) return fail(AV1P_EINVAL, "bad skip range");
  if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0) || !(lr >= 0.0) || !(weight_decay >= 0.0))
    return fail(AV1P_EINVAL, "AdamW hyper-parameters out of range");
  if (int rc = ensure_ctx()) return rc;
  DpAdamWArgs a = {};
  for (int p = 0; p < world; ++p) {
    if (!grad_ptrs[p] || !param_ptrs[p] || !flag_ptrs[p]) return fail(AV1P_EINVAL, "null peer pointer for rank %d", p);
    if ((reinterpret_cast<uintptr_t>(grad_ptrs[p]) | reinterpret_cast<uintptr_t>(param_ptrs[p])) & 15u)
      return fail(AV1P_EINVAL, "peer buffers must be 16-byte aligned");
    a.grad[p] = grad_ptrs[p];
    a.param[p] = param_ptrs[p];
    a.flags[p] = flag_ptrs[p];
  }
  if ((reinterpret_cast<uintptr_t>(exp_avg_shard_dev) | reinterpret_cast<uintptr_t>(exp_avg_sq_shard_dev)) & 15u)
    return fail(AV1P_EINVAL, "moment shards must be 16-byte aligned");
  a.m = exp_avg_shard_dev;
  a.v = exp_avg_sq_shard_dev;
  a.n = n;
  a.shard = shard;
  a.skip_lo = skip_lo;
  a.skip_hi = skip_hi;
  a.rank = rank;
  a.world = world;
  a.epoch = epoch;
  a.h = AdamWArgs{nullptr, nullptr, nullptr, nullptr, 0, lr, beta1, beta2, float(1.0 - lr * weight_decay), float(1.0 - beta1),
                  float(beta2), float(1.0 - beta2), float(eps), float(1.0 / world), step_dev};
  a.err = err_dev;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  train_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
  CUDA_TRY(cudaGetLastError());
  // every CTA must be resident (phase 0 spins in all of them): at most four CTAs of 256 threads per SM
  const long long want = ((std::min<long long>(shard, n) >> 2) + TRAIN_THREADS - 1) / TRAIN_THREADS;
  const int grid = int(std::max<long long>(1, std::min<long long>(want, (long long)g_ctx.sms * 4)));
  dp_adamw_fused_kernel<<<grid, TRAIN_THREADS, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  return AV1P_OK;
}
extern "C" int av1p_dp_flag_words(void) { return DP_FLAG_WORDS; }
// Kernels of the CURRENT device may dereference memory of device `peer_device` afterwards (NVLink / NVSwitch peer mapping).
extern "C" int av1p_enable_peer_access(int32_t peer_device) {
  int me = 0, can = 0;
  CUDA_TRY(cudaGetDevice(&me));
  if (peer_device == me) return AV1P_OK;
  CUDA_TRY(cudaDeviceCanAccessPeer(&can, me, peer_device));
  if (!can) return fail(AV1P_ENODEV, "device %d cannot access device %d as a peer", me, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();                      // clear the sticky-free error state
    return AV1P_OK;
  }
  if (e != cudaSuccess) return fail(AV1P_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
  return AV1P_OK;
}

// CUDA IPC plumbing for the peer buffers of av1p_dp_adamw_fused.  Export: handle of the cudaMalloc block that holds
// `dev_ptr` + the pointer's offset inside it.  Import (with the IMPORTING rank's device current, so that the mapping is
// made for that device's kernels; peer access is enabled as needed): base address of the block in this process.
extern "C" int av1p_ipc_export(const void* dev_ptr, uint8_t handle_out[64], int64_t* offset_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!dev_ptr || !handle_out || !offset_out) return fail(AV1P_EINVAL, "null argument");
  CUdeviceptr base = 0;
  size_t size = 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(AV1P_ECUDA, "cuMemGetAddressRange not available");
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  const CUresult r = reinterpret_cast<RangeFn>(fn)(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr));
  if (r != CUDA_SUCCESS) return fail(AV1P_ECUDA, "cuMemGetAddressRange failed with %d", int(r));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle_out, &h, sizeof h);
  *offset_out = int64_t(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
  return AV1P_OK;
}
extern "C" int av1p_ipc_import(const uint8_t handle[64], void** base_out) {
  if (!handle || !base_out) return fail(AV1P_EINVAL, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CUDA_TRY(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return AV1P_OK;
}
extern "C" int av1p_ipc_close(void* base) {
  if (!base) return AV1P_OK;
  CUDA_TRY(cudaIpcCloseMemHandle(base));
  return AV1P_OK;
}
