// layer1 convolutions (3x3, stride 1, 64 -> 64 channels on the 4x4 map) with SMEM-resident weights.
//
// The generic block-Toeplitz FC kernel (fc_tcgen05.cuh) streams a weight tile next to every activation
// tile; for layer1 that makes it L2->SM bandwidth bound (ncu: 9.7 TB/s of TMA fill, tensor pipe 40-48 %).
// A 3x3 conv has only nine distinct [64 co x 64 ci] weight matrices, 72 KB in fp16 (144 KB as hi + lo
// planes), so this kernel keeps them in shared memory for its whole life and streams activations only:
//
//   out[block, (oy,ox), co] = epi( sum_{(iy,ix) in 3x3 window of (oy,ox)} act[block, (iy,ix), :] . W[ky,kx][co, :] )
//       ky = iy - oy + 1, kx = ix - ox + 1     (reference: torchvision BasicBlock conv3x3, models.py:110)
//
// Work decomposition per 128-block M tile: two halves (output rows {0,1}, then {2,3}); a half needs
// three input rows = 12 activation tiles [128 x 64] per plane, each loaded once per half and multiplied
// into every output position that sees it.  One accumulator = one output row = 4 positions x 64
// channels = 256 TMEM columns (rows 0,1,2,3 -> column halves 0,1,0,1), handed to the epilogue position by
// position (CR_SLOTS below) so that the stores of a tile's tail overlap the MMAs of the next tile's head.
// The three horizontally adjacent output positions fed by one input tile use taps kx = 2,1,0 - the
// resident weights are stored in that order per ky, so they are ONE tcgen05.mma with N = 192
// (N = 128 at the left/right edge): B = rows [(2-kx_first)*64, ...) of the ky stack.
//
// Split precision (fp16x3): hi and lo activation tiles are separate ring slots, each released as soon
// as its own products are issued (x_hi: x_hi.w_hi + x_hi.w_lo; x_lo: x_lo.w_hi).
//
// Residual (BasicBlock identity branch): after the last conv product of an output row, the residual's
// four position tiles (hi, lo) come through the same ring and are accumulated as  x . (S I)  with a
// 16 x 16 scaled identity as the B operand - the epilogue never loads an aux row.  Output rows leave
// through the staging tiles + TMA bulk stores of fc_tcgen05.cuh.
//
// Shared memory: 3 x 16 KB ring (tiles prefetched into L2) + 144 KB weights + 32 KB staging + identity + barriers.
#pragma once
#include <cuda.h>
#include "fc_tcgen05.cuh"

namespace av1p {

constexpr int CR_STAGES = 3;                                      // shallow ring: tiles are prefetched into L2 (CR_PREFETCH ahead)
constexpr int CR_PREFETCH = 12;
constexpr int CR_A_BYTES = FC_TILE_M * FC_TILE_K * 2;             // 16 KB
constexpr int CR_W_TILE_BYTES = 64 * 64 * 2;                      // one tap, one plane: 8 KB
constexpr int CR_W_PLANE_BYTES = 9 * CR_W_TILE_BYTES;             // 72 KB
constexpr int CR_W_BYTES = 2 * CR_W_PLANE_BYTES;                  // hi + lo planes
constexpr int CR_OFF_W = CR_STAGES * CR_A_BYTES;
constexpr int CR_OFF_STAGING = CR_OFF_W + CR_W_BYTES;
constexpr int CR_OFF_IDENT = CR_OFF_STAGING + EPI_STAGING_BYTES;
constexpr int CR_OFF_BARS = CR_OFF_IDENT + EPI_IDENT_BYTES;
constexpr int CR_SMEM_BYTES = CR_OFF_BARS + 256 + 1024 /*align*/;
constexpr int CR_THREADS = FC_THREADS;
static_assert(CR_SMEM_BYTES <= 232448, "conv_res shared memory exceeds the 227 KB opt-in limit");
static_assert(FC_SMEM_BYTES <= 232448, "fc shared memory exceeds the 227 KB opt-in limit");

// The per-M-tile work list is static.  The producer walks `ring[i]` = which [128 x 64] activation tile to load
// i-th (map << 4 | position), compiled on the host by conv_res_build_schedule; the MMA issuer runs the same order as
// fully unrolled straight-line code (cr_issue_mtile): with one thread issuing, every instruction between two
// tcgen05.mma is tensor-pipe idle time, and an interpreted schedule cost ~350 cycles per group of four MMAs
// (measured: the skeleton without any MMA / memory traffic ran at 60 % of the full kernel's time).
constexpr int CR_MAX_RING = 80;
// Accumulators: the 512 TMEM columns are eight 64-column POSITION slots; output position p = oy * 4 + ox of an M tile
// lives in slot p & 7 (rows 0/2 -> slots 0..3, rows 1/3 -> slots 4..7).  The issuer commits every position as soon as
// its last input tile has been multiplied in and waits for a slot only right before the first product into it; the
// epilogue drains positions in the order below (= completion order: during the last input row the positions of rows 2
// and 3 complete in lock-step), so the next M tile starts while the tail of this one is still being stored.  (With two
// whole-row accumulators both rows of a half completed together and the issuer idled for two row epilogues per tile.)
constexpr int CR_SLOTS = 8;
__device__ __constant__ const uint8_t CR_DRAIN[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 9, 13, 10, 14, 11, 15};
static_assert((2 * CR_STAGES + 2 * CR_SLOTS + 4 + 1 + 2) * 8 + 4 <= 256, "barrier area");

struct ConvResParams {
  CUtensorMap a_map[4];        // x_hi, x_lo, residual hi, residual lo (tiled layout): 2-D [rows*16][64] fp16, box {64, 128}, SWIZZLE_128B
  CUtensorMap w_map;           // resident weights: 2-D [planes*9*64][64] fp16, box {64, 64}; tile (plane, ky, 2-kx)
  CUtensorMap out_map[2];      // output hi, lo (tiled layout): 2-D [rows*16][64] fp16, box {32, 128}, SWIZZLE_64B
  CUtensorMap res_map[2];      // resid_epi: residual hi, lo with the output's box (loaded into the staging sets)
  const int* n_rows_dev;
  int n_rows;
  int split;                   // 1: hi/lo planes, three products; 0: single fp16 product
  int has_aux_lo;              // residual has a lo plane
  int resid_epi;               // 1: the residual is added in the epilogue - the store warps TMA-load its [128 x 32] tiles into the
                               // staging set the epilogue is about to fill (in place), instead of 16 extra ring tiles and 64
                               // identity MMAs per half M tile (a ring tile that feeds four tiny MMAs exposes a ring round trip);
                               // 2: added in the epilogue from per-thread 16-byte global loads issued before the wait for the
                               // accumulator (L2-prefetched).  259 k rows, same box: 760 us (1), 785 us (2), 883 us (0); plain conv 522 us
  int debug;                   // test hook only (AV1P_CR_DEBUG): 1 skip MMA issue, 2 skip staging/stores, 4 skip TMA loads, 8 skip L2 prefetch, 16 skip the epilogue,
                               // 32 skip the L2 prefetch of the in-place residual tiles
  int n_ring;
  int ring_begin[3];           // ring entries of half h: [ring_begin[h], ring_begin[h + 1])
  uint8_t ring[CR_MAX_RING];
  // epilogue members (names shared with FcParams, see epi_tile_store)
  int epi;
  const float* bias;           // [1024]
  const float* row_scale;      // always nullptr here
  float acc_scale;
  const __half* aux;           // resid_epi: residual planes (read directly for the partial last M tile)
  const __half* aux_lo;
  int aux_kb;
  __half* out;
  __half* out_lo;
  int out_kb;                  // 16
  int* err_flag;
};

// One group = four K=16 instructions of one ring tile against a run of n64 resident tap tiles.
__device__ __forceinline__ void cr_group(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t n64, uint32_t acc0) {
  const uint32_t idesc = 0x08000010u + (n64 << 20);          // umma_idesc_f16(64 * n64)
  umma_f16_ss_lo(d_tmem, a_lo, b_lo, idesc, acc0);
  umma_f16_ss_lo(d_tmem, a_lo + 2, b_lo + 2, idesc, 1u);
  umma_f16_ss_lo(d_tmem, a_lo + 4, b_lo + 4, idesc, 1u);
  umma_f16_ss_lo(d_tmem, a_lo + 6, b_lo + 6, idesc, 1u);
}

struct CrPipe {
  uint64_t* full_bar;
  uint64_t* empty_bar;
  uint64_t* acc_full;
  uint64_t* acc_empty;
  int stage;
  uint32_t phase;
  uint32_t acc_phase;      // bit s: parity of position slot s (0..7)
};

// MMA issue for one work item = half H of an M tile (output rows 2H, 2H + 1), straight-line (every loop below has
// compile-time bounds and unrolls completely).
template <bool SPLIT, bool RESID, bool AUX_LO, int H>
__device__ __forceinline__ void cr_issue_half(CrPipe& q, uint32_t base, uint32_t tmem_base, uint64_t id_desc, int* err_flag,
                                              bool skip_mma) {
  constexpr int PLANES = SPLIT ? 2 : 1;
  constexpr int AUX_PLANES = AUX_LO ? 2 : 1;
  const uint32_t w_lo0 = umma_desc_lo_sw128(base + CR_OFF_W);
  const uint32_t idesc_id = umma_idesc_f16(16u);
  auto wait_tile = [&]() -> uint32_t {
    mbar_wait(&q.full_bar[q.stage], q.phase, err_flag, 300 + q.stage);
    return umma_desc_lo_sw128(base + q.stage * CR_A_BYTES);
  };
  auto next_stage = [&]() {
    if (++q.stage == CR_STAGES) {
      q.stage = 0;
      q.phase ^= 1u;
    }
  };
  {
    constexpr int h = H;
#pragma unroll
    for (int iyi = 0; iyi < 3; ++iyi) {
      const int iy = h + iyi;
#pragma unroll
      for (int xi = 0; xi < 4; ++xi) {
        const int ix = xi == 0 ? 1 : (xi == 1 ? 0 : xi);        // order 1, 0, 2, 3: a fresh output row starts with one N = 192 group
        const int ox0 = ix > 0 ? ix - 1 : 0;
        const int ox1 = ix < 3 ? ix + 1 : 3;
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) {
          const uint32_t a_lo = wait_tile();
          if (pl == 0 && (xi == 0 || xi == 2)) {
            // first touch of output positions (tile ix = 1 starts positions 0..2 of a row, tile ix = 2 position 3): the
            // epilogue must have drained the previous tenant of each 64-column position slot
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              const int oy = 2 * h + o;
              const int first_iy = oy < 2 ? 0 : oy - 1;
              if (iy != first_iy) continue;
#pragma unroll
              for (int ox = (xi == 0 ? 0 : 3); ox < (xi == 0 ? 3 : 4); ++ox) {
                const int slot = o * 4 + ox;
                mbar_wait(&q.acc_empty[slot], ((q.acc_phase >> slot) & 1u) ^ 1u, err_flag, 200 + slot);
              }
            }
          }
          tc_fence_after_sync();
          if (elect_one_sync()) {
            if (!skip_mma) {
#pragma unroll
              for (int o = 0; o < 2; ++o) {
                const int oy = 2 * h + o;
                const int ky = iy - oy + 1;
                if (ky < 0 || ky > 2) continue;
                const int first_iy = oy < 2 ? 0 : oy - 1;
                const bool fresh = (iy == first_iy) && pl == 0;      // positions may still hold the previous tile's sums
                const uint32_t d_row = tmem_base + uint32_t(o * 256);
                // taps for ox..oe are kx = ix-ox+1 .. (descending) = tiles (2 - kx_first).. of the ky stack
                auto run = [&](int ox, int oe, uint32_t acc0) {
                  const uint32_t tile = uint32_t(ky * 3 + (2 - (ix - ox + 1)));
                  const uint32_t b_hi = w_lo0 + tile * (CR_W_TILE_BYTES >> 4);
                  cr_group(d_row + uint32_t(ox * 64), a_lo, b_hi, uint32_t(oe - ox + 1), acc0);
                  if (SPLIT && pl == 0)
                    cr_group(d_row + uint32_t(ox * 64), a_lo, b_hi + (CR_W_PLANE_BYTES >> 4), uint32_t(oe - ox + 1), 1u);
                };
                if (!fresh || xi == 1 || xi == 3) {
                  run(ox0, ox1, 1u);
                } else if (xi == 0) {
                  run(ox0, ox1, 0u);          // ix = 1: positions 0..2, all fresh
                } else {
                  run(1, 2, 1u);              // ix = 2: positions 1, 2 were started by ix = 1 ...
                  run(3, 3, 0u);              // ... position 3 is new
                }
              }
            }
            umma_commit(&q.empty_bar[q.stage]);     // frees the ring slot once these MMAs have read it
            if (pl == PLANES - 1 && xi > 0) {
              // positions of the output rows whose last input row this is, complete once tile ix has been multiplied in:
              // ix = 0 (xi 1) -> position 0, ix = 2 -> position 1, ix = 3 -> positions 2 and 3 (with a residual branch
              // position 3 still waits for its residual tile, which closes the row below)
#pragma unroll
              for (int o = 0; o < 2; ++o) {
                const int oy = 2 * h + o;
                if (iy != (oy < 3 ? oy + 1 : 3)) continue;
                umma_commit(&q.acc_full[o * 4 + xi - 1]);
                if (!RESID && xi == 3) umma_commit(&q.acc_full[o * 4 + 3]);
              }
            }
          }
          __syncwarp();
          if (pl == PLANES - 1 && xi > 0) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              const int oy = 2 * h + o;
              if (iy != (oy < 3 ? oy + 1 : 3)) continue;
              q.acc_phase ^= 1u << (o * 4 + xi - 1);
              if (!RESID && xi == 3) q.acc_phase ^= 1u << (o * 4 + 3);
            }
          }
          next_stage();
        }
        // Output rows whose last input row is iy: their residual tile of position ox = xi follows immediately, so the
        // (almost MMA-free) residual ring tiles interleave with conv tiles instead of exposing a ring round trip each.
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int oy = 2 * h + o;
          if (!RESID || iy != (oy < 3 ? oy + 1 : 3)) continue;
          const int ox = xi;
#pragma unroll
          for (int pl = 0; pl < AUX_PLANES; ++pl) {
            const uint32_t a_lo = wait_tile();
            tc_fence_after_sync();
            if (elect_one_sync()) {
              const uint32_t d_tmem = tmem_base + uint32_t(o * 256 + ox * 64);
              if (!skip_mma) {
#pragma unroll
                for (int j = 0; j < FC_TILE_K / 16; ++j) {
                  const uint64_t a_desc = (uint64_t(0x40004040u) << 32) | uint64_t(a_lo + 2 * j);
                  umma_f16_ss(d_tmem + j * 16, a_desc, id_desc, idesc_id, 1u);
                }
              }
              umma_commit(&q.empty_bar[q.stage]);
              if (ox == 3 && pl == AUX_PLANES - 1) umma_commit(&q.acc_full[o * 4 + 3]);
            }
            __syncwarp();
            if (ox == 3 && pl == AUX_PLANES - 1) q.acc_phase ^= 1u << (o * 4 + 3);
            next_stage();
          }
        }
      }
    }
  }
}

template <bool SPLIT, bool RESID, bool AUX_LO>
__global__ void __launch_bounds__(CR_THREADS, 1) conv_res_tcgen05_kernel(const __grid_constant__ ConvResParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CR_OFF_BARS);
  uint64_t* empty_bar = full_bar + CR_STAGES;
  uint64_t* acc_full = empty_bar + CR_STAGES;   // [CR_SLOTS]
  uint64_t* acc_empty = acc_full + CR_SLOTS;    // [CR_SLOTS]
  uint64_t* stg_full = acc_empty + CR_SLOTS;    // [2]
  uint64_t* stg_free = stg_full + 2;            // [2]
  uint64_t* w_bar = stg_free + 2;
  uint64_t* stg_loaded = w_bar + 1;             // [2] in-place residual tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_loaded + 2);
  EpiStage es;
  epi_stage_init(es, smem + CR_OFF_STAGING, stg_full, stg_free);
  es.loaded = stg_loaded;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int planes = SPLIT ? 2 : 1;
  pdl_launch_dependents();
  constexpr bool residual = RESID;
  // Work item = (M tile, half): the grid is even, CTA b always takes half b & 1 of the M tiles b >> 1, b >> 1 + grid / 2, ...
  // so that (a) the wave quantisation of a launch is in half tiles, (b) few-tile launches spread over twice as many SMs and
  // (c) the two CTAs that need input rows 1 and 2 of an M tile read them at the same time (one DRAM read, one L2 hit).
  const int half = int(blockIdx.x & 1u);
  const int mt0 = int(blockIdx.x >> 1);
  const int mt_step = int(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.a_map[1]);
    tma_prefetch_desc(&p.w_map);
    tma_prefetch_desc(&p.out_map[0]);
    if (p.out_lo) tma_prefetch_desc(&p.out_map[1]);
    if (residual || p.resid_epi) {
      tma_prefetch_desc(&p.a_map[2]);
      tma_prefetch_desc(&p.a_map[3]);
    }
    for (int s = 0; s < CR_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < CR_SLOTS; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], FC_EPI_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&stg_full[s], FC_EPI_WARPS / 2);
      mbar_init(&stg_free[s], 1);
      mbar_init(&stg_loaded[s], 1);
    }
    mbar_init(w_bar, 1);
    if (p.resid_epi) {
      tma_prefetch_desc(&p.res_map[0]);
      tma_prefetch_desc(&p.res_map[1]);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  write_ident_tile(smem + CR_OFF_IDENT, 1.0f / p.acc_scale, threadIdx.x, CR_THREADS);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // resident weights first: planes x 9 tiles of 8 KB on one barrier.  They are constant data, so the load is issued
    // before pdl_wait() and overlaps the previous kernel's tail (the MMA warp always waits for it, work or no work).
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, uint32_t(planes * CR_W_PLANE_BYTES));
      for (int t = 0; t < planes * 9; ++t)
        tma_load_2d(smem + CR_OFF_W + t * CR_W_TILE_BYTES, &p.w_map, w_bar, 0, t * 64);
    }
    __syncwarp();
  }
  pdl_wait();      // the previous kernel's activations and the device-side row count are visible from here on
  const int n_rows = p.n_rows_dev ? *p.n_rows_dev : p.n_rows;
  const int m_tiles = (n_rows + FC_TILE_M - 1) / FC_TILE_M;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, one lane issues)
    if (mt0 < m_tiles) {
      int stage = 0;
      uint32_t phase = 0;
      // L2 prefetch cursor: runs CR_PREFETCH ring tiles ahead of the loads (into the next M tile of this CTA)
      const int rb = p.ring_begin[half], re = p.ring_begin[half + 1];
      int pf_mt = mt0, pf_i = rb;
      auto prefetch_next = [&]() {
        if (pf_mt >= m_tiles) return;
        const uint32_t e = p.ring[pf_i];
        if (!(p.debug & 8) && elect_one_sync()) tma_prefetch_l2_2d(&p.a_map[e >> 4], 0, (pf_mt * 16 + int(e & 15u)) * FC_TILE_M);
        __syncwarp();
        if (++pf_i == re) {
          pf_i = rb;
          pf_mt += mt_step;
        }
      };
      for (int i = 0; i < CR_PREFETCH; ++i) prefetch_next();
      for (int mt = mt0; mt < m_tiles; mt += mt_step) {
        for (int i = rb; i < re; ++i) {
          const uint32_t e = p.ring[i];
          prefetch_next();
          mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag, 100 + stage);
          if (elect_one_sync()) {
            if (p.debug & 4) {
              mbar_arrive(&full_bar[stage]);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], CR_A_BYTES);
              tma_load_2d(smem + stage * CR_A_BYTES, &p.a_map[e >> 4], &full_bar[stage], 0, (mt * 16 + int(e & 15u)) * FC_TILE_M);
            }
          }
          __syncwarp();
          if (++stage == CR_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform, one lane issues)
    mbar_wait(w_bar, 0u, p.err_flag, 500);      // also when this CTA has no work: it must not exit under its own TMA load
    if (mt0 < m_tiles) {
      CrPipe q{full_bar, empty_bar, acc_full, acc_empty, 0, 0u, 0u};
      const uint64_t id_desc = ident_desc(base + CR_OFF_IDENT);
      tc_fence_after_sync();
      if (half == 0) {
        for (int mt = mt0; mt < m_tiles; mt += mt_step)
          cr_issue_half<SPLIT, RESID, AUX_LO, 0>(q, base, tmem_base, id_desc, p.err_flag, (p.debug & 1) != 0);
      } else {
        for (int mt = mt0; mt < m_tiles; mt += mt_step)
          cr_issue_half<SPLIT, RESID, AUX_LO, 1>(q, base, tmem_base, id_desc, p.err_flag, (p.debug & 1) != 0);
      }
    }
  } else if (warp < FC_STORE_WARP) {
    // ------------------------------------------------------------ epilogue (warps 2..9): output rows 0..3 -> slots 0,1,0,1
    uint32_t acc_phase = 0u;      // bit s: parity of position slot s
    uint32_t g = 0;
    for (int mt = mt0; mt < m_tiles; mt += mt_step) {
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int pos = CR_DRAIN[8 * half + i];
        const int slot = pos & 7;
        epi_tile_store(p, es, g, n_rows, mt, pos * 64, 64, tmem_base + uint32_t(slot * 64), &acc_full[slot],
                       (acc_phase >> slot) & 1u, &acc_empty[slot], warp, lane, 400 + slot);
        acc_phase ^= 1u << slot;
      }
    }
  } else {
    // ------------------------------------------------------------ store warp
    uint32_t g = 0;
    for (int mt = mt0; mt < m_tiles; mt += mt_step) {
      if ((mt + 1) * FC_TILE_M > n_rows || (p.debug & 18)) continue;
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        if (p.resid_epi && !(p.debug & 32) && warp == FC_STORE_WARP && lane == 0) {
          // pull the residual of the NEXT output position into L2 (two chunk periods ahead of its in-place loads)
          const int nmt = i < 7 ? mt : mt + mt_step;
          if ((nmt + 1) * FC_TILE_M <= n_rows) {
            const int npos = int(CR_DRAIN[8 * half + ((i + 1) & 7)]);
            tma_prefetch_l2_2d(&p.a_map[2], 0, (nmt * 16 + npos) * FC_TILE_M);
            if (p.aux_lo) tma_prefetch_l2_2d(&p.a_map[3], 0, (nmt * 16 + npos) * FC_TILE_M);
          }
        }
        epi_store_chunks(es, g, &p.out_map[0], &p.out_map[1], p.out_lo != nullptr, int(CR_DRAIN[8 * half + i]) * 64, 64 / EPI_CHUNK, mt, 16,
                         p.err_flag, uint32_t(warp - FC_STORE_WARP), p.resid_epi == 1 ? &p.res_map[0] : nullptr,
                         (p.resid_epi == 1 && p.aux_lo) ? &p.res_map[1] : nullptr);
      }
    }
    epi_store_drain();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// Host side: the producer's per-M-tile load order (must match cr_issue_mtile): two halves (output rows {0,1},
// {2,3}); per half the three input rows it needs; per input row the positions ix = 1, 0, 2, 3; hi plane then lo
// plane; while an output row's last input row streams by, its residual tile of position xi follows x tile xi.
inline bool conv_res_build_schedule(ConvResParams& f) {
  const int planes = f.split ? 2 : 1;
  const bool residual = f.epi == FC_EPI_ADD_RELU && !f.resid_epi;
  const int aux_planes = f.has_aux_lo ? 2 : 1;
  static const int ix_order[4] = {1, 0, 2, 3};
  int n = 0;
  for (int h = 0; h < 2; ++h) {
    f.ring_begin[h] = n;
    for (int iy = h; iy < h + 3; ++iy) {
      for (int xi = 0; xi < 4; ++xi) {
        for (int pl = 0; pl < planes; ++pl) {
          if (n >= CR_MAX_RING) return false;
          f.ring[n++] = uint8_t((pl << 4) | (iy * 4 + ix_order[xi]));
        }
        for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
          if (!residual || iy != (oy < 3 ? oy + 1 : 3)) continue;
          for (int pl = 0; pl < aux_planes; ++pl) {
            if (n >= CR_MAX_RING) return false;
            f.ring[n++] = uint8_t(((2 + pl) << 4) | (oy * 4 + xi));
          }
        }
      }
    }
  }
  f.n_ring = n;
  f.ring_begin[2] = n;
  return true;
}

// Kernel variant for the given shape (split precision / residual branch / residual lo plane).
typedef void (*ConvResKernel)(const ConvResParams);
inline ConvResKernel conv_res_kernel_for(const ConvResParams& f) {
  const bool resid = f.epi == FC_EPI_ADD_RELU && !f.resid_epi;      // residual on the tensor core (identity MMAs)
  if (f.split) {
    if (!resid) return conv_res_tcgen05_kernel<true, false, false>;
    return f.has_aux_lo ? conv_res_tcgen05_kernel<true, true, true> : conv_res_tcgen05_kernel<true, true, false>;
  }
  if (!resid) return conv_res_tcgen05_kernel<false, false, false>;
  return f.has_aux_lo ? conv_res_tcgen05_kernel<false, true, true> : conv_res_tcgen05_kernel<false, true, false>;
}
inline const ConvResKernel* conv_res_all_kernels(int* n) {
  static const ConvResKernel all[6] = {conv_res_tcgen05_kernel<true, false, false>, conv_res_tcgen05_kernel<true, true, true>,
                                       conv_res_tcgen05_kernel<true, true, false>, conv_res_tcgen05_kernel<false, false, false>,
                                       conv_res_tcgen05_kernel<false, true, true>, conv_res_tcgen05_kernel<false, true, false>};
  *n = 6;
  return all;
}

}  // namespace av1p
