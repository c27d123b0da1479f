// Block-Toeplitz fully-connected layer on tcgen05 tensor cores.
//
// Every conv / linear layer of the v6 backbone (reference pesquisa_v6/v6_pipeline/models.py:104-126
// and torchvision BasicBlock) is executed, at 16x16 input, as
//     out[m, n] = epilogue( sum_k  act[m, k] * W[n, k] )            m = block index in the batch
// where act is the previous layer's output stored row-major per block ([position][channel] flattened)
// and W is the convolution unrolled over the (tiny) spatial grid.  W is block-sparse: an output
// position only sees the input positions under its 3x3 window, so for every N tile the packer
// lists the 64-wide K blocks that are not identically zero and only those are multiplied.
//
// Kernel shape (persistent, warp specialised, one CTA per SM):
//   warp 0      TMA producer : A tile [128 rows x 64 K] + W tile [block_n x 64 K] per K block
//   warp 1      MMA issuer   : 4 x tcgen05.mma (M128, N=block_n, K16) per K block, fp32 acc in TMEM
//   warps 2..17 epilogue     : tcgen05.ld -> bias / ReLU / gate -> fp16 hi/lo -> swizzled smem staging tile
//   warps 18,19 store        : one TMA bulk store per staged [128 rows x 32 cols] tile (hi and lo planes), one warp per set
// with a 4-deep smem ring (full/empty mbarriers), a 2-deep TMEM accumulator ring and a hi/lo pair of
// staging tiles (full/free mbarriers).
//
// Why the staging: a TMEM lane is an output row, so an epilogue warp holds 32 different rows; storing
// them straight to global touches 32 cache lines per instruction and the L1 wavefront pipe, not the
// tensor pipe, bounded the kernel (ncu, round 1: l1tex lsu wavefronts 43-67 % vs tensor 38-47 %).
// Residual connections never pass through the epilogue either: the identity branch is accumulated
// on the tensor core as  x . (S I)  (S = the layer's power-of-two weight scale), K block by K block,
// from the same TMA ring - see FC_W_IDENT.
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"

namespace av1p {

constexpr int FC_TILE_M = 128;
constexpr int FC_TILE_K = 64;          // one 128-byte swizzle atom of fp16
constexpr int FC_MAX_N = 256;          // per-tile N (UMMA N limit)
constexpr int FC_STAGES = 4;
constexpr int FC_A_BYTES = FC_TILE_M * FC_TILE_K * 2;        // 16 KB
constexpr int FC_W_BYTES = FC_MAX_N * FC_TILE_K * 2;         // 32 KB
constexpr int FC_STAGE_BYTES = FC_A_BYTES + FC_W_BYTES;      // 48 KB
// CTA-pair variant (cta_group::2): each CTA of the pair loads its own A tile and HALF of the W tile per K block, so a
// stage is 32 KB and six of them fit the same ring area; the W half the MMA needs from the other CTA never crosses
// this SM's L2->SM port (62 -> 42 B/clk/SM at full tensor rate).
constexpr int FC2_STAGES = 6;
constexpr int FC2_W_BYTES = FC_W_BYTES / 2;                  // 16 KB
constexpr int FC2_STAGE_BYTES = FC_A_BYTES + FC2_W_BYTES;    // 32 KB
static_assert(FC2_STAGES * FC2_STAGE_BYTES == FC_STAGES * FC_STAGE_BYTES, "both variants share one shared-memory layout");
constexpr int FC_MAX_NT = 8;           // N tiles per layer
constexpr int FC_MAX_KB = 192;         // scheduled K blocks per layer (sum over N tiles), incl. the residual entries added at plan time
constexpr int FC_MAX_SRC = 4;          // activation sources per layer
constexpr int FC_TAIL_MAX = 8;         // outputs of the in-epilogue final linear (7 for the flatten head)
constexpr int FC_EPI_WARPS = 16;         // four per TMEM lane quadrant: 8 of every 32 staged columns each
constexpr int FC_STORE_WARPS = 2;                            // one per staging set
constexpr int FC_THREADS = 64 + 32 * FC_EPI_WARPS + 32 * FC_STORE_WARPS;   // producer, MMA, epilogue, store
constexpr int FC_STORE_WARP = 2 + FC_EPI_WARPS;
constexpr int EPI_CHUNK = 32;                                // output columns per staging tile
constexpr int EPI_UNIT_BYTES = FC_TILE_M * EPI_CHUNK * 2;    // 8 KB: [128 rows][64 B], SWIZZLE_64B
constexpr int EPI_IDENT_BYTES = 512;                         // 16 x 16 fp16 scaled identity, no swizzle
constexpr int EPI_SETS = 2;                                  // staging is double buffered: chunk g uses set g & 1
constexpr int EPI_STAGING_BYTES = EPI_SETS * 2 * EPI_UNIT_BYTES;   // {hi, lo} x 2 sets = 32 KB
constexpr int FC_OFF_STAGING = FC_STAGES * FC_STAGE_BYTES;
// Layers whose epilogue needs a second activation operand - gate layers (FC_EPI_GATE: out = aux * sigmoid(acc), the SE
// excitation) and residual layers (FC_EPI_ADD_RELU with aux_epi) - run on a shortened operand ring and the upper part of
// the ring area becomes a ring of AUX tiles: the [128 rows x 32 cols] hi / lo pieces of that operand, loaded by the
// producer warp with TMA in the staging tiles' SWIZZLE_64B layout.  (For the residual layers this measured SLOWER than
// residual K blocks on the tensor core - the MMA pipeline misses the two operand-ring stages - so aux_epi is off by default.)  (Reading aux straight from global touched 32 cache lines per load instruction - one per
// accumulator row - and made se4.fc2 the slowest FC layer at 5 % tensor activity.)
constexpr int GATE_AUX_SETS = 4;
constexpr int GATE_AUX_SET_BYTES = 2 * EPI_UNIT_BYTES;       // hi + lo tile
constexpr int FC_OFF_GATE_AUX = 128 * 1024;
constexpr int FC_GATE_STAGES = 2;                            // 2 x 48 KB (single CTA) ...
constexpr int FC2_GATE_STAGES = 4;                           // ... 4 x 32 KB (CTA pair) stay below FC_OFF_GATE_AUX
static_assert(FC_GATE_STAGES * FC_STAGE_BYTES <= FC_OFF_GATE_AUX && FC2_GATE_STAGES * FC2_STAGE_BYTES <= FC_OFF_GATE_AUX &&
              FC_OFF_GATE_AUX + GATE_AUX_SETS * GATE_AUX_SET_BYTES <= FC_OFF_STAGING, "gate aux ring must fit the ring area");
constexpr int FC_OFF_TAIL = FC_OFF_STAGING;                  // head layers store nothing: their tail weights reuse the staging area
constexpr int FC_OFF_IDENT = FC_OFF_STAGING + EPI_STAGING_BYTES;
constexpr int FC_OFF_BARS = FC_OFF_IDENT + EPI_IDENT_BYTES;
constexpr int FC_SMEM_BYTES = FC_OFF_BARS + 256 + 1024 /*align*/;
static_assert(FC_TAIL_MAX * FC_MAX_N * 4 <= EPI_STAGING_BYTES, "tail weights must fit the staging area");
constexpr uint16_t FC_W_IDENT = 0xFFFFu;   // schedule entry: A tile is a residual K block, B is the scaled identity

// Precision.  Operands are fp16, accumulation is fp32 in TMEM.  In split mode ("fp16x3") every
// activation x is stored as hi = fp16(x), lo = fp16(x - hi), every weight as hi/lo likewise, and the
// schedule lists three products per K block: (x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi) - about 22
// significant bits, i.e. fp32-grade logits, on the fp16 tensor pipe.  Weights are pre-scaled by a power
// of two per layer (acc_scale undoes it) so that the lo parts stay in fp16's normal range.
enum FcEpilogue : int {
  FC_EPI_LINEAR = 0,    // out = s*acc + b
  FC_EPI_RELU = 1,      // out = relu(s*acc + b)
  FC_EPI_ADD_RELU = 2,  // out = relu(s*acc + b + residual): acc already holds S * residual (FC_W_IDENT schedule entries; default),
                        // or, with aux_epi = 1, the residual arrives through the aux ring and is added here (A/B knob)
  FC_EPI_GATE = 3,      // out = aux * sigmoid(acc)             (SE excitation)
  FC_EPI_HEAD = 4,      // h = relu(s*acc + b); logits = h . tail_w^T + tail_b   (fp32, no fp16 store)
  FC_EPI_ADD = 5,       // out = s*acc + b + r * aux, no ReLU (adapter up-projection + skip, models.py:292-310); aux arrives
                        // through the aux ring; r = aux_row_scale[row] (the spatial-attention scalar) or 1
};

struct FcParams {
  CUtensorMap a_map[FC_MAX_SRC]; // activation sources (tiled layout, see act_off): 2-D [rows*KB][64] fp16, box {64, 128}, SWIZZLE_128B
  CUtensorMap w_map;           // packed weights, 2-D [n_kb_total*block_n][64] fp16, box {64, block_n}
  CUtensorMap w_half_map;      // same tensor, box {64, block_n / 2} (CTA-pair variant)
  CUtensorMap out_map[2];      // output hi / lo planes (tiled layout): 2-D [rows*KB][64] fp16, box {32, 128}, SWIZZLE_64B
  CUtensorMap aux_map[2];      // FC_EPI_GATE: gate input hi / lo planes, same box as out_map (loaded, not stored)
  int src_kb[FC_MAX_SRC];      // 64-column blocks per row of each source buffer
  const int* n_rows_dev;       // device-side row count (nullptr -> n_rows)
  int n_rows;
  int n_tiles;                 // number of N tiles
  int block_n;                 // tile N (multiple of 32, <= 256)
  int epi;
  const float* bias;           // [n_tiles*block_n] or nullptr
  const float* row_scale;      // [rows] or nullptr (spatial-attention scalar folded into the next linear)
  const float* aux_row_scale;  // FC_EPI_ADD: [rows] scale of the aux operand, or nullptr
  float* sam_part;             // optional [rows][n_tiles * 4][2]: per-thread partial (sum, max) of this layer's OUTPUT over the
                               // columns a thread handles - the spatial-attention statistics (models.py:56-61) of se4's
                               // output without a second pass over it (sam_finish_kernel reduces the partials per row)
  float acc_scale;             // power of two undoing the weight pre-scale
  const __half* aux;           // gate input rows (FC_EPI_GATE only)
  const __half* aux_lo;        // split mode: low part of aux (nullptr otherwise)
  int aux_kb;                  // 64-column blocks per row of the aux buffer
  int aux_epi;                 // FC_EPI_ADD_RELU: 1 = the epilogue adds the residual from the aux ring (no FC_W_IDENT entries)
  __half* out;
  __half* out_lo;              // split mode: low part of the output (nullptr otherwise)
  int out_kb;                  // 64-column blocks per row of the output buffer
  int out_col0;                // first output column of this op (multiple of 64; bias / schedule are op-local, buffers are not)
  const float* tail_w;         // [tail_n][block_n]
  const float* tail_b;         // [tail_n]
  float* logits;               // [rows][tail_n]
  int tail_n;
  int pair_mode;               // split precision: schedule entries come in pairs e = (x_hi, w_hi), e+1 = (x_lo, w_lo)
                               // of one K block; the MMA warp issues hi*hi, hi*lo, lo*hi (each tile is loaded once)
  int* err_flag;
  int kb_begin[FC_MAX_NT + 1]; // schedule range of each N tile
  uint16_t kb_src[FC_MAX_KB];  // bits 14..15: activation source, bits 0..13: K offset / 64
  uint16_t kb_w[FC_MAX_KB];    // weight chunk ([block_n x 64] tile) index, or FC_W_IDENT
};

// SE / attention gates: ex2.approx + rcp.approx (relative error ~2e-7 + 1e-7, far inside the split-precision budget);
// the accurate expf + IEEE division cost ~25 instructions per element and were ~40 % of the gate layers' time.
// Routing decisions (aux_kernels.cuh) keep the accurate form.
// (the clamp keeps 1 + e^-x finite: __fdividef(1, inf) is NaN, not 0)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-fmaxf(x, -80.0f))); }

// ------------------------------------------------------------------------------------------------
// Epilogue staging: two sets of two [128 rows x 32 cols] fp16 tiles (hi plane, lo plane) in the SWIZZLE_64B layout
// a TMA store expects; chunk g of a CTA's output stream uses set g & 1.
//   full[s] : the 8 epilogue warps of group s -> store warp s ("both tiles of set s are written and fenced")
//   free_[s]: store warp -> epilogue warps ("the bulk stores have finished reading set s")
//   loaded[s]: (in-place residual, conv_res kernel only) TMA transaction barrier: the store warp has loaded the RESIDUAL of the
//              set's next chunk into the set; the epilogue adds it and overwrites it with the output (see epi_tile_store)
struct EpiStage {
  uint8_t* staging;  // [EPI_SETS][2 planes][EPI_UNIT_BYTES]
  uint64_t* full;    // [EPI_SETS], count FC_EPI_WARPS / 2
  uint64_t* free_;   // [EPI_SETS], count 1
  uint64_t* loaded;  // [EPI_SETS] or nullptr
  __device__ __forceinline__ uint8_t* unit(uint32_t set, int plane) const {
    return staging + (set * 2u + uint32_t(plane)) * EPI_UNIT_BYTES;
  }
};
__device__ __forceinline__ void epi_stage_init(EpiStage& es, uint8_t* staging, uint64_t* full, uint64_t* free_) {
  es.staging = staging;
  es.full = full;
  es.free_ = free_;
  es.loaded = nullptr;
}

// 8 consecutive fp16 of row r: 16-byte chunk `part` (0..3) of the 64-byte row
__device__ __forceinline__ void stage_store8(uint8_t* unit, int r_local, int part, const uint4& a) {
  const int sw = (r_local >> 1) & 3;                     // Swizzle<2,4,3>: chunk index ^= address bits [7,9)
  *reinterpret_cast<uint4*>(unit + r_local * (EPI_CHUNK * 2) + ((part ^ sw) << 4)) = a;
}

__device__ __forceinline__ uint4 stage_load8(const uint8_t* unit, int r_local, int part) {
  const int sw = (r_local >> 1) & 3;
  return *reinterpret_cast<const uint4*>(unit + r_local * (EPI_CHUNK * 2) + ((part ^ sw) << 4));
}

// Gate input ring (FC_EPI_GATE): set ga % GATE_AUX_SETS holds the hi (+ lo) aux tile of the ga-th output chunk of this CTA.
//   full[s] : TMA transaction barrier (producer warp posts the expectation);  empty[s]: the 16 epilogue warps hand it back
struct GateAux {
  const uint8_t* tiles;
  uint64_t* full;
  uint64_t* empty;
  uint32_t ga;
};

// 16 x 16 identity scaled by `s` in the no-swizzle K-major core-matrix layout (LBO 128 B, SBO 256 B)
__device__ __forceinline__ void write_ident_tile(uint8_t* ident, float s, int tid, int nthreads) {
  for (int i = tid; i < EPI_IDENT_BYTES / 2; i += nthreads) {
    // element index i -> (16-byte unit u, element e); unit u = (n & 7) + 8 * (k >> 3) + 16 * (n >> 3)
    const int u = i >> 3, e = i & 7;
    const int n = (u & 7) + ((u >> 4) << 3), k = (((u >> 3) & 1) << 3) + e;
    reinterpret_cast<__half*>(ident)[i] = __float2half_rn(n == k ? s : 0.f);
  }
}
__device__ __forceinline__ uint64_t ident_desc(uint32_t ident_addr) { return umma_desc_nosw(ident_addr, 128, 256); }

// Store warps: drain `n_chunks` staged chunks of one accumulator (columns col0 .., rows row0 ..).  There is one store
// warp per staging set (`my_set`); each walks the whole chunk sequence and handles the chunks that land in its set:
// wait until the 16 epilogue warps have filled it, issue the bulk stores (lane 0: bulk async-groups are per-thread
// state, so one fixed lane owns them), wait until the TMA engine has READ the tiles and hand the set straight back.
// A set therefore returns to the epilogue as soon as its own stores have left shared memory, independently of the
// other set's progress.  (With a single store warp the hand-back of set s waited for the NEXT chunk to be staged and
// issued - wait_group.read cannot be polled - which serialised epilogue and stores at ~1.3 us per chunk, the pace of
// every store-heavy layer.)
// In-place residual (res_hi != nullptr): before the epilogue may touch the set, the store warp loads the RESIDUAL tile of the
// same chunk into it (the previous stores of this set have left shared memory - wait_group.read above - so the set is free);
// the epilogue reads its residual values from exactly the 16-byte pieces it then overwrites with the output.
__device__ __forceinline__ void epi_store_chunks(const EpiStage& es, uint32_t& g, const CUtensorMap* map_hi,
                                                 const CUtensorMap* map_lo, bool has_lo, int col0, int n_chunks, int mt,
                                                 int out_kb, int* err_flag, uint32_t my_set, const CUtensorMap* res_hi = nullptr,
                                                 const CUtensorMap* res_lo = nullptr) {
  const bool leader = (threadIdx.x & 31) == 0;
  for (int c = 0; c < n_chunks; ++c, ++g) {
    const uint32_t set = g & 1u;
    if (set != my_set) continue;
    if (res_hi != nullptr && leader) {
      const int col = col0 + c * EPI_CHUNK;
      const int trow = (mt * out_kb + (col >> 6)) * FC_TILE_M;      // the residual buffer has the output's width and layout
      mbar_arrive_expect_tx(&es.loaded[set], res_lo ? 2u * EPI_UNIT_BYTES : uint32_t(EPI_UNIT_BYTES));
      tma_load_2d(es.unit(set, 0), res_hi, &es.loaded[set], col & 63, trow);
      if (res_lo) tma_load_2d(es.unit(set, 1), res_lo, &es.loaded[set], col & 63, trow);
    }
    mbar_wait(&es.full[set], (g >> 1) & 1u, err_flag, 900 + int(set));
    if (leader) {
      const int col = col0 + c * EPI_CHUNK;
      const int trow = (mt * out_kb + (col >> 6)) * FC_TILE_M;      // tiled layout: tile (mt, col / 64)
      tma_store_2d(map_hi, es.unit(set, 0), col & 63, trow);
      if (has_lo) tma_store_2d(map_lo, es.unit(set, 1), col & 63, trow);
      tma_store_commit();
      tma_store_wait_read<0>();
      mbar_arrive(&es.free_[set]);
    }
    __syncwarp();
  }
}
// After the last chunk of the kernel: nothing hands the final set back (nobody waits for it) but the stores must
// have left shared memory before the CTA exits.
__device__ __forceinline__ void epi_store_drain() {
  if ((threadIdx.x & 31) == 0) tma_store_wait_all<0>();
  __syncwarp();
}

// Epilogue of one accumulator tile [128 rows x block_n fp32 columns at TMEM column t_col]: warps 2..17, four
// warps per TMEM lane quadrant, all on the same 32-column chunk (8 columns each).  Waits on `full`, applies
// scale / bias / gate / ReLU, writes fp16 hi (+lo) either into the staging tiles (whole M tiles) or straight to
// global (the last, partial M tile: rows past n_rows stay untouched), arrives on `empty` as soon as the last
// TMEM read has completed.  The TMEM read of chunk c+1 is in flight while chunk c is converted and staged.
// P is FcParams or any struct with the same epilogue members.
// Hand an accumulator back to the MMA issuer: a local arrive, or (peer CTA of a pair) an arrive on the leader's barrier.
__device__ __forceinline__ void acc_release(uint64_t* empty, uint32_t empty_cluster_addr) {
  if (empty_cluster_addr) mbar_arrive_cluster(empty_cluster_addr);
  else mbar_arrive(empty);
}

template <typename P>
__device__ __forceinline__ int epi_debug(const P& p) {
  if constexpr (requires { p.debug; }) return p.debug; else return 0;
}
template <typename P>
__device__ __forceinline__ void epi_tile_store(const P& p, const EpiStage& es, uint32_t& g, int n_rows, int mt, int col0,
                                               int block_n, uint32_t t_col, uint64_t* full, uint32_t full_phase,
                                               uint64_t* empty, int warp, int lane, int tag, uint32_t empty_remote = 0u,
                                               GateAux* gx = nullptr) {
  // The 16 epilogue warps form two groups of eight (two warps per TMEM lane quadrant).  Group k owns staging set k and
  // handles the chunks whose running index has parity k, 16 columns per thread (two 8-column halves), so two chunks are
  // in flight per CTA and a group only ever synchronises with its own store warp.  (All 16 warps sharing every
  // 32-column chunk left each warp 8 columns of work per barrier round trip: ~1.3 us per chunk, latency-bound.)
  const int quad = warp & 3;              // TMEM lane quadrant this warp may read
  const uint32_t grp = uint32_t(warp - 2) >> 3;
  const int sub = ((warp - 2) >> 2) & 1;  // which 16 columns of the chunk
  const int r_local = quad * 32 + lane;
  const int row = mt * FC_TILE_M + r_local;
  const bool row_ok = row < n_rows;
  const bool skip_out = (epi_debug(p) & 2) != 0;
  const bool staged = (mt + 1) * FC_TILE_M <= n_rows && !skip_out;
  const bool gate = p.epi == FC_EPI_GATE && gx != nullptr;
  const bool resid = (p.epi == FC_EPI_ADD_RELU || p.epi == FC_EPI_ADD) && gx != nullptr;      // residual through the aux ring
  const bool aux_on = gate || resid;
  const bool relu = p.epi == FC_EPI_RELU || p.epi == FC_EPI_ADD_RELU;
  const bool has_lo = p.out_lo != nullptr;
  bool inplace = false;             // residual added here (conv_res kernel, resid_epi): 1 = it arrives in the staging set itself,
  bool direct = false;              // 2 = every thread loads its 2 x 16 bytes per plane straight from global memory (L2-prefetched)
  if constexpr (requires { p.resid_epi; }) {
    inplace = p.resid_epi != 0 && p.epi == FC_EPI_ADD_RELU;
    direct = inplace && p.resid_epi == 2;
  }
  const float rs = ((p.row_scale && row_ok) ? p.row_scale[row] : 1.0f) * p.acc_scale;
  float ars = 1.0f;                 // scale of the aux operand (FC_EPI_ADD after spatial attention)
  if constexpr (requires { p.aux_row_scale; }) ars = (p.aux_row_scale && row_ok) ? p.aux_row_scale[row] : 1.0f;
  const int n_chunks = block_n / EPI_CHUNK;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  const uint32_t g0 = g;
  const int c_first = int((g0 & 1u) ^ grp);                  // first chunk of this tile with (g0 + c) & 1 == grp
  // direct residual: this thread's residual values of its first chunk are requested BEFORE the wait for the accumulator, so
  // that their (L2) latency overlaps the time the epilogue spends waiting for the tensor pipe anyway
  uint4 pre_q[2] = {zero4, zero4}, pre_l[2] = {zero4, zero4};
  if (direct && row_ok && c_first < n_chunks) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t o = act_off(row, col0 + c_first * EPI_CHUNK + sub * 16 + h * 8, p.aux_kb);
      pre_q[h] = __ldg(reinterpret_cast<const uint4*>(p.aux + o));
      if (p.aux_lo) pre_l[h] = __ldg(reinterpret_cast<const uint4*>(p.aux_lo + o));
    }
  }
  mbar_wait(full, full_phase, p.err_flag, tag);
  tc_fence_after_sync();
  if (epi_debug(p) & 16) {        // development switch: hand the accumulator straight back
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) acc_release(empty, empty_remote);
    return;
  }
  const uint32_t ga0 = aux_on ? gx->ga : 0u;
  float ssum = 0.f, smax = -INFINITY;         // sam_part partials of this thread (columns of its chunks in this tile)
  const uint32_t t_addr = t_col + (uint32_t(quad * 32) << 16) + uint32_t(sub * 16);
  uint32_t v[8];
  bool released = false;
  if (c_first < n_chunks) tmem_ld_32x8(t_addr + uint32_t(c_first * EPI_CHUNK), v);
  for (int c = c_first; c < n_chunks; c += 2) {
    uint4 ax[2] = {zero4, zero4}, axl[2] = {zero4, zero4};
    uint32_t aset = 0;
    if (aux_on) {
      // this chunk's gate input from the aux ring (both halves).  The set is handed back to the producer only after the
      // values have been USED (below): an mbarrier arrive is not ordered behind shared-memory loads still queued in the
      // LSU (here behind the partial tile's direct global stores), and the refill is an async-proxy write - released
      // right after issuing the loads, a late load read the NEXT chunk's tile (seen as run-to-run differences in the
      // partial last M tile once the gate math became fast).
      const uint32_t ga = ga0 + uint32_t(c);
      aset = ga % GATE_AUX_SETS;
      mbar_wait(&gx->full[aset], (ga / GATE_AUX_SETS) & 1u, p.err_flag, tag + 20);
      const uint8_t* u = gx->tiles + aset * GATE_AUX_SET_BYTES;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ax[h] = stage_load8(u, r_local, sub * 2 + h);
        if (p.aux_lo) axl[h] = stage_load8(u + EPI_UNIT_BYTES, r_local, sub * 2 + h);
      }
    }
    uint4 rq[2] = {zero4, zero4}, rql[2] = {zero4, zero4};      // in-place residual: this thread's 2 x 8 values (hi, lo)
    if (inplace) {
      if (direct && c == c_first) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          rq[h] = pre_q[h];
          rql[h] = pre_l[h];
        }
      } else if (staged && !direct) {
        // use u = (g0 + c) >> 1 of set `grp`: its residual tile has landed (which also means the set's previous stores are done)
        mbar_wait(&es.loaded[grp], ((g0 + uint32_t(c)) >> 1) & 1u, p.err_flag, tag + 30);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          rq[h] = stage_load8(es.unit(grp, 0), r_local, sub * 2 + h);
          if (p.aux_lo) rql[h] = stage_load8(es.unit(grp, 1), r_local, sub * 2 + h);
        }
      } else if (row_ok) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const size_t o = act_off(row, col0 + c * EPI_CHUNK + sub * 16 + h * 8, p.aux_kb);
          rq[h] = *reinterpret_cast<const uint4*>(p.aux + o);
          if (p.aux_lo) rql[h] = *reinterpret_cast<const uint4*>(p.aux_lo + o);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = col0 + c * EPI_CHUNK + sub * 16 + h * 8;
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (p.bias) {
        int bcol = col;                                       // the bias vector is local to the op
        if constexpr (requires { p.out_col0; }) bcol -= p.out_col0;
        b0 = __ldg(reinterpret_cast<const float4*>(p.bias + bcol));
        b1 = __ldg(reinterpret_cast<const float4*>(p.bias + bcol) + 1);
      }
      tmem_ld_wait();
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[i]) * rs;
      if (h == 0) {
        tmem_ld_32x8(t_addr + uint32_t(c * EPI_CHUNK + 8), v);             // second half in flight
      } else if (c + 2 < n_chunks) {
        tmem_ld_32x8(t_addr + uint32_t((c + 2) * EPI_CHUNK), v);           // this group's next chunk in flight
      } else {
        // every TMEM read of this accumulator by this warp is complete -> hand it back to the MMA warp
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) acc_release(empty, empty_remote);
        released = true;
      }
      f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
      f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      if (aux_on) {
        const __half2* hh = reinterpret_cast<const __half2*>(&ax[h]);
        const __half2* hl = reinterpret_cast<const __half2*>(&axl[h]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __half22float2(hh[i]), al = __half22float2(hl[i]);
          if (gate) {
            f[2 * i] = (a.x + al.x) * fast_sigmoid(f[2 * i]);
            f[2 * i + 1] = (a.y + al.y) * fast_sigmoid(f[2 * i + 1]);
          } else {
            f[2 * i] = fmaf(ars, a.x + al.x, f[2 * i]);
            f[2 * i + 1] = fmaf(ars, a.y + al.y, f[2 * i + 1]);
          }
        }
        if (h == 1) {             // both halves' aux registers have been consumed: the loads are complete
          __syncwarp();
          if (lane == 0) mbar_arrive(&gx->empty[aset]);
        }
      }
      if (inplace) {
        const __half2* hh = reinterpret_cast<const __half2*>(&rq[h]);
        const __half2* hl = reinterpret_cast<const __half2*>(&rql[h]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __half22float2(hh[i]), al = __half22float2(hl[i]);
          f[2 * i] += a.x + al.x;
          f[2 * i + 1] += a.y + al.y;
        }
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      if constexpr (requires { p.sam_part; }) {
        if (p.sam_part) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            ssum += f[i];
            smax = fmaxf(smax, f[i]);
          }
        }
      }
      __align__(16) __half2 hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hi[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        const float2 hf = __half22float2(hi[i]);
        lo[i] = __floats2half2_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
      }
      const uint4 hq = *reinterpret_cast<const uint4*>(hi);
      const uint4 lq = *reinterpret_cast<const uint4*>(lo);
      if (staged) {
        // use u = (g0 + c) >> 1 of set `grp`: the first use needs no wait (parity trick), use u waits for the (u-1)-th hand-back
        if (h == 0 && (!inplace || direct)) mbar_wait(&es.free_[grp], (((g0 + uint32_t(c)) >> 1) & 1u) ^ 1u, p.err_flag, tag + 10);
        stage_store8(es.unit(grp, 0), r_local, sub * 2 + h, hq);
        if (has_lo) stage_store8(es.unit(grp, 1), r_local, sub * 2 + h, lq);
        if (h == 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&es.full[grp]);
        }
      } else if (row_ok && !skip_out) {
        const size_t o = act_off(row, col, p.out_kb);
        *reinterpret_cast<uint4*>(p.out + o) = hq;
        if (has_lo) *reinterpret_cast<uint4*>(p.out_lo + o) = lq;
      }
    }
  }
  if (!released) {                // a group without a chunk in this tile (odd chunk counts) still signs off
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) acc_release(empty, empty_remote);
  }
  if constexpr (requires { p.sam_part; }) {
    if (p.sam_part && row_ok && c_first < n_chunks) {
      // slot (N tile, group, sub): every (row, slot) is written by exactly one thread
      const int slots = p.n_tiles * 4;
      float* d = p.sam_part + (size_t(row) * slots + size_t((col0 / block_n) * 4 + int(grp) * 2 + sub)) * 2;   // out_col0 = 0 here
      d[0] = ssum;
      d[1] = smax;
    }
  }
  if (staged) g = g0 + uint32_t(n_chunks);
  if (aux_on) gx->ga = ga0 + uint32_t(n_chunks);
}

// Head epilogue (FC_EPI_HEAD): h = relu(s*acc + b), logits = h . tail_w^T + tail_b in the thread that owns the row.
// The first four epilogue warps take every column; the others only release the accumulator.
__device__ __forceinline__ void epi_tile_head(const FcParams& p, int n_rows, int mt, int block_n, uint32_t t_col,
                                              uint64_t* full, uint32_t full_phase, uint64_t* empty, const float* tail_w_s,
                                              int warp, int lane, int tag, uint32_t empty_remote = 0u) {
  const int quad = warp & 3;
  const int half = (warp - 2) >> 2;
  const int row = mt * FC_TILE_M + quad * 32 + lane;
  const bool row_ok = row < n_rows;
  const float rs = ((p.row_scale && row_ok) ? p.row_scale[row] : 1.0f) * p.acc_scale;
  mbar_wait(full, full_phase, p.err_flag, tag);
  tc_fence_after_sync();
  if (half == 0) {
    const uint32_t t_addr = t_col + (uint32_t(quad * 32) << 16);
    float tail[FC_TAIL_MAX];
#pragma unroll
    for (int j = 0; j < FC_TAIL_MAX; ++j) tail[j] = 0.f;
    for (int c = 0; c < block_n; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(t_addr + uint32_t(c), v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * rs;
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(b4 + i);
          f[4 * i + 0] += b.x;
          f[4 * i + 1] += b.y;
          f[4 * i + 2] += b.z;
          f[4 * i + 3] += b.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
#pragma unroll
      for (int j = 0; j < FC_TAIL_MAX; ++j) {
        if (j < p.tail_n) {
          const float* w = tail_w_s + j * block_n + c;
          float t = tail[j];
#pragma unroll
          for (int i = 0; i < 32; ++i) t = fmaf(f[i], w[i], t);
          tail[j] = t;
        }
      }
    }
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < FC_TAIL_MAX; ++j)
        if (j < p.tail_n) p.logits[size_t(row) * p.tail_n + j] = tail[j] + p.tail_b[j];
    }
  }
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) acc_release(empty, empty_remote);
}

// 16 x 16 scaled identity for the CTA-pair variant: the pair MMA takes B rows 0..7 from the leader and rows 8..15 from
// the peer, each at the descriptor's start address, so CTA `rank` stores rows 8*rank .. 8*rank+7 as its first row group.
__device__ __forceinline__ void write_ident_tile_pair(uint8_t* ident, float s, uint32_t rank, int tid, int nthreads) {
  for (int i = tid; i < EPI_IDENT_BYTES / 2; i += nthreads) {
    const int u = i >> 3, e = i & 7;
    const int n_local = (u & 7) + ((u >> 4) << 3), k = (((u >> 3) & 1) << 3) + e;
    reinterpret_cast<__half*>(ident)[i] = __float2half_rn((n_local < 8 && int(rank) * 8 + n_local == k) ? s : 0.f);
  }
}

// PAIR = false: one CTA per SM, tcgen05.mma.cta_group::1 (M = 128).
// PAIR = true : clusters of two CTAs (launch with cluster dimension 2); CTA rank r of a cluster works on M tile 2*mp + r
//               of its item (mp, nt); the leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256) for both.
template <bool PAIR>
__global__ void __launch_bounds__(FC_THREADS, 1) fc_tcgen05_kernel(const __grid_constant__ FcParams p) {
  constexpr int STAGES = PAIR ? FC2_STAGES : FC_STAGES;
  constexpr int STAGE_BYTES = PAIR ? FC2_STAGE_BYTES : FC_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + FC_OFF_BARS);   // [STAGES] (pair: the leader's are used)
  uint64_t* empty_bar = full_bar + FC2_STAGES;                            // [STAGES]
  uint64_t* acc_full = empty_bar + FC2_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2] (pair: the leader's collect both CTAs' epilogue warps)
  uint64_t* stg_full = acc_empty + 2;            // [2]
  uint64_t* stg_free = stg_full + 2;             // [2]
  uint64_t* aux_full = stg_free + 2;             // [GATE_AUX_SETS] gate layers only
  uint64_t* aux_empty = aux_full + GATE_AUX_SETS; // [GATE_AUX_SETS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_empty + GATE_AUX_SETS);
  static_assert((2 * FC2_STAGES + 8 + 2 * GATE_AUX_SETS) * 8 + 4 <= 256, "barrier area");
  float* tail_w_s = reinterpret_cast<float*>(smem + FC_OFF_TAIL);
  EpiStage es;
  epi_stage_init(es, smem + FC_OFF_STAGING, stg_full, stg_free);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0u;
  const bool gate = p.epi == FC_EPI_GATE || p.epi == FC_EPI_ADD || (p.epi == FC_EPI_ADD_RELU && p.aux_epi);    // layers that use the aux ring
  const int stages = gate ? (PAIR ? FC2_GATE_STAGES : FC_GATE_STAGES) : STAGES;   // operand ring depth of this layer
  const int worker = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);        // index of this CTA (pair) among the workers
  const int n_workers = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < FC_MAX_SRC; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(PAIR ? &p.w_half_map : &p.w_map);
    if (p.out) tma_prefetch_desc(&p.out_map[0]);
    if (p.out_lo) tma_prefetch_desc(&p.out_map[1]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], PAIR ? 2 * FC_EPI_WARPS : FC_EPI_WARPS);
      mbar_init(&stg_full[s], FC_EPI_WARPS / 2);
      mbar_init(&stg_free[s], 1);
    }
    for (int s = 0; s < GATE_AUX_SETS; ++s) {
      mbar_init(&aux_full[s], 1);
      mbar_init(&aux_empty[s], FC_EPI_WARPS / 2);
    }
    if (gate) {
      tma_prefetch_desc(&p.aux_map[0]);
      if (p.aux_lo) tma_prefetch_desc(&p.aux_map[1]);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  if (p.epi == FC_EPI_HEAD) {
    for (int i = threadIdx.x; i < p.tail_n * p.block_n; i += FC_THREADS) tail_w_s[i] = p.tail_w[i];
  }
  if constexpr (PAIR) write_ident_tile_pair(smem + FC_OFF_IDENT, 1.0f / p.acc_scale, rank, threadIdx.x, FC_THREADS);
  else write_ident_tile(smem + FC_OFF_IDENT, 1.0f / p.acc_scale, threadIdx.x, FC_THREADS);
  fence_proxy_async_smem();       // identity tile: generic stores, read by tcgen05.mma
  tc_fence_before_sync();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // everything above overlapped the previous kernel's tail (programmatic dependent launch); its results - the activations
  // and the device-side row count - are visible from here on
  pdl_wait();
  const int n_rows = p.n_rows_dev ? *p.n_rows_dev : p.n_rows;
  const int m_tiles = (n_rows + FC_TILE_M - 1) / FC_TILE_M;
  const int m_groups = PAIR ? (m_tiles + 1) / 2 : m_tiles;                  // M tiles (pairs of M tiles) to process
  const int n_items = m_groups * p.n_tiles;
  auto item_mt = [&](int item) { return PAIR ? 2 * (item / p.n_tiles) + int(rank) : item / p.n_tiles; };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, one lane issues)
    int stage = 0;
    uint32_t phase = 0;
    uint32_t ga = 0;
    const uint32_t w_rows = PAIR ? uint32_t(p.block_n) / 2u : uint32_t(p.block_n);
    const uint32_t tx_bytes = FC_A_BYTES + w_rows * FC_TILE_K * 2;
    for (int item = worker; item < n_items; item += n_workers) {
      const int mt = item_mt(item);
      const int nt = item % p.n_tiles;
      for (int kb = p.kb_begin[nt]; kb < p.kb_begin[nt + 1]; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag, 100 + stage);
        const uint32_t e = p.kb_src[kb];
        const uint32_t wi = p.kb_w[kb];
        uint8_t* a_dst = smem + stage * STAGE_BYTES;
        if (elect_one_sync()) {
          const uint32_t bytes = wi == FC_W_IDENT ? uint32_t(FC_A_BYTES) : tx_bytes;
          const int a_row = (mt * p.src_kb[e >> 14] + int(e & 0x3FFFu)) * FC_TILE_M;
          if constexpr (PAIR) {
            // the leader's barrier counts both CTAs' bytes; only the leader posts the expectation
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * bytes);
            tma_load_2d_pair(a_dst, &p.a_map[e >> 14], &full_bar[stage], 0, a_row);
            if (wi != FC_W_IDENT)
              tma_load_2d_pair(a_dst + FC_A_BYTES, &p.w_half_map, &full_bar[stage], 0, int(wi) * p.block_n + int(rank * w_rows));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            tma_load_2d(a_dst, &p.a_map[e >> 14], &full_bar[stage], 0, a_row);
            if (wi != FC_W_IDENT) tma_load_2d(a_dst + FC_A_BYTES, &p.w_map, &full_bar[stage], 0, int(wi) * p.block_n);
          }
        }
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (gate) {
        // gate input of this item's output chunks, in the order the epilogue consumes them (local barriers: every CTA
        // of a pair feeds its own epilogue)
        const int n_chunks = p.block_n / EPI_CHUNK;
        for (int c = 0; c < n_chunks; ++c, ++ga) {
          const uint32_t set = ga % GATE_AUX_SETS;
          mbar_wait(&aux_empty[set], ((ga / GATE_AUX_SETS) & 1u) ^ 1u, p.err_flag, 150 + int(set));
          if (elect_one_sync()) {
            const int col = p.out_col0 + nt * p.block_n + c * EPI_CHUNK;
            const int trow = (mt * p.aux_kb + (col >> 6)) * FC_TILE_M;
            uint8_t* dst = smem + FC_OFF_GATE_AUX + set * GATE_AUX_SET_BYTES;
            mbar_arrive_expect_tx(&aux_full[set], p.aux_lo ? 2u * EPI_UNIT_BYTES : uint32_t(EPI_UNIT_BYTES));
            tma_load_2d(dst, &p.aux_map[0], &aux_full[set], col & 63, trow);
            if (p.aux_lo) tma_load_2d(dst + EPI_UNIT_BYTES, &p.aux_map[1], &aux_full[set], col & 63, trow);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one lane issues; pair: leader only)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = PAIR ? umma_idesc_f16_pair(uint32_t(p.block_n)) : umma_idesc_f16(uint32_t(p.block_n));
      const uint32_t idesc_id = PAIR ? umma_idesc_f16_pair(16u) : umma_idesc_f16(16u);
      const uint64_t id_desc = ident_desc(base + FC_OFF_IDENT);
      const uint32_t a_lo0 = umma_desc_lo_sw128(base);
      auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc0) {
        if constexpr (PAIR) umma_f16_ss_lo_pair(d, a_lo, b_lo, idesc, acc0);
        else umma_f16_ss_lo(d, a_lo, b_lo, idesc, acc0);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      for (int item = worker; item < n_items; item += n_workers) {
        const int nt = item % p.n_tiles;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1u, p.err_flag, 200 + acc);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * FC_MAX_N);
        const int kb0 = p.kb_begin[nt], kb1 = p.kb_begin[nt + 1];
        uint32_t prev_a = 0, prev_w = 0;
        int prev_stage = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 300 + stage);
          tc_fence_after_sync();
          // low descriptor words of this slot's A and W tiles (+2 per K step of 16 elements)
          const uint32_t a_lo = a_lo0 + uint32_t(stage) * (STAGE_BYTES >> 4);
          const uint32_t w_lo = a_lo + (FC_A_BYTES >> 4);
          const bool ident = p.kb_w[kb] == FC_W_IDENT;
          const bool first_of_pair = p.pair_mode && !ident && ((kb - kb0) & 1) == 0;
          if (elect_one_sync()) {
            if (ident) {
              // residual K block: acc[:, c0 .. c0+63] += A . (S I), 16 columns per instruction.  These entries
              // close a tile's schedule, so the accumulator already holds data.
              const uint32_t c0 = (uint32_t(p.kb_src[kb]) & 0x3FFFu) * FC_TILE_K - uint32_t(p.out_col0 + nt * p.block_n);
#pragma unroll
              for (int j = 0; j < FC_TILE_K / 16; ++j) {
                const uint64_t a_desc = (uint64_t(0x40004040u) << 32) | uint64_t(a_lo + 2 * j);
                if constexpr (PAIR) umma_f16_ss_pair(d_tmem + c0 + j * 16, a_desc, id_desc, idesc_id, 1u);
                else umma_f16_ss(d_tmem + c0 + j * 16, a_desc, id_desc, idesc_id, 1u);
              }
              commit(&empty_bar[stage]);
            } else if (!p.pair_mode) {
              mma(d_tmem, a_lo, w_lo, kb > kb0 ? 1u : 0u);
              mma(d_tmem, a_lo + 2, w_lo + 2, 1u);
              mma(d_tmem, a_lo + 4, w_lo + 4, 1u);
              mma(d_tmem, a_lo + 6, w_lo + 6, 1u);
              commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
            } else if (first_of_pair) {
              // (x_hi, w_hi): the slot stays live until the cross products of the next entry are done
              mma(d_tmem, a_lo, w_lo, kb > kb0 ? 1u : 0u);
              mma(d_tmem, a_lo + 2, w_lo + 2, 1u);
              mma(d_tmem, a_lo + 4, w_lo + 4, 1u);
              mma(d_tmem, a_lo + 6, w_lo + 6, 1u);
            } else {
              // this slot holds (x_lo, w_lo): issue x_hi * w_lo and x_lo * w_hi
#pragma unroll
              for (int k = 0; k < FC_TILE_K / 16; ++k) mma(d_tmem, prev_a + 2 * k, w_lo + 2 * k, 1u);
#pragma unroll
              for (int k = 0; k < FC_TILE_K / 16; ++k) mma(d_tmem, a_lo + 2 * k, prev_w + 2 * k, 1u);
              commit(&empty_bar[prev_stage]);
              commit(&empty_bar[stage]);
            }
            if (kb + 1 == kb1) commit(&acc_full[acc]);   // accumulator complete -> epilogue (of both CTAs)
          }
          __syncwarp();
          if (first_of_pair) {
            prev_a = a_lo;
            prev_w = w_lo;
            prev_stage = stage;
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp < FC_STORE_WARP) {
    // ------------------------------------------------------------ epilogue (warps 2..17)
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t g = 0;
    GateAux gx{smem + FC_OFF_GATE_AUX, aux_full, aux_empty, 0u};
    // pair: the peer's epilogue warps hand accumulators back on the leader's barriers
    const uint32_t rem0 = (PAIR && !leader) ? mapa_shared(smem_u32(&acc_empty[0]), 0u) : 0u;
    const uint32_t rem1 = (PAIR && !leader) ? mapa_shared(smem_u32(&acc_empty[1]), 0u) : 0u;
    for (int item = worker; item < n_items; item += n_workers) {
      const int mt = item_mt(item);
      const int nt = item % p.n_tiles;
      const uint32_t t_col = tmem_base + uint32_t(acc * FC_MAX_N);
      const uint32_t rem = acc ? rem1 : rem0;
      if (p.epi == FC_EPI_HEAD)
        epi_tile_head(p, n_rows, mt, p.block_n, t_col, &acc_full[acc], acc_phase, &acc_empty[acc], tail_w_s, warp, lane, 400 + acc, rem);
      else
        epi_tile_store(p, es, g, n_rows, mt, p.out_col0 + nt * p.block_n, p.block_n, t_col, &acc_full[acc], acc_phase, &acc_empty[acc],
                       warp, lane, 400 + acc, rem, gate ? &gx : nullptr);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ store warp: staged tiles -> global (TMA)
    if (p.epi != FC_EPI_HEAD) {
      uint32_t g = 0;
      for (int item = worker; item < n_items; item += n_workers) {
        const int mt = item_mt(item);
        const int nt = item % p.n_tiles;
        if ((mt + 1) * FC_TILE_M > n_rows) continue;      // partial tile: the epilogue stores it directly
        epi_store_chunks(es, g, &p.out_map[0], &p.out_map[1], p.out_lo != nullptr, p.out_col0 + nt * p.block_n, p.block_n / EPI_CHUNK,
                         mt, p.out_kb, p.err_flag, uint32_t(warp - FC_STORE_WARP));
      }
      epi_store_drain();
    }
  }

  tc_fence_before_sync();
  if constexpr (PAIR) cluster_sync_all();     // nobody may exit while the other CTA can still signal its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace av1p
