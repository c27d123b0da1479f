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
// channels = 256 TMEM columns; the two accumulators alternate exactly like the FC kernel's double
// buffer (output rows 0,1,2,3 -> slots 0,1,0,1), so output row r's epilogue overlaps the MMAs of r+1.
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
// Shared memory: 4 x 16 KB ring + 144 KB weights + 16 KB staging + identity + barriers = 225.8 KB.
#pragma once
#include <cuda.h>
#include "fc_tcgen05.cuh"

namespace av1p {

constexpr int CR_STAGES = 4;
constexpr int CR_A_BYTES = FC_TILE_M * FC_TILE_K * 2;             // 16 KB
constexpr int CR_W_TILE_BYTES = 64 * 64 * 2;                      // one tap, one plane: 8 KB
constexpr int CR_W_PLANE_BYTES = 9 * CR_W_TILE_BYTES;             // 72 KB
constexpr int CR_W_BYTES = 2 * CR_W_PLANE_BYTES;                  // hi + lo planes
constexpr int CR_OFF_W = CR_STAGES * CR_A_BYTES;
constexpr int CR_OFF_STAGING = CR_OFF_W + CR_W_BYTES;
constexpr int CR_OFF_IDENT = CR_OFF_STAGING + 2 * EPI_UNIT_BYTES;
constexpr int CR_OFF_BARS = CR_OFF_IDENT + EPI_IDENT_BYTES;
constexpr int CR_SMEM_BYTES = CR_OFF_BARS + 256 + 1024 /*align*/;
constexpr int CR_THREADS = FC_THREADS;
static_assert(CR_SMEM_BYTES <= 232448, "conv_res shared memory exceeds the 227 KB opt-in limit");
static_assert(FC_SMEM_BYTES <= 232448, "fc shared memory exceeds the 227 KB opt-in limit");

struct ConvResParams {
  CUtensorMap a_map[2];        // x_hi, x_lo: 2-D [rows][1024] fp16, box {64, 128}, SWIZZLE_128B
  CUtensorMap aux_map[2];      // residual hi, lo (same geometry); used when epi == FC_EPI_ADD_RELU
  CUtensorMap w_map;           // resident weights: 2-D [planes*9*64][64] fp16, box {64, 64}; tile (plane, ky, 2-kx)
  CUtensorMap out_map[2];      // output hi, lo: 2-D [rows][1024] fp16, box {32, 128}, SWIZZLE_64B
  const int* n_rows_dev;
  int n_rows;
  int split;                   // 1: hi/lo planes, three products; 0: single fp16 product
  int has_aux_lo;              // residual has a lo plane
  // epilogue members (names shared with FcParams, see epi_tile_store)
  int epi;
  const float* bias;           // [1024]
  const float* row_scale;      // always nullptr here
  float acc_scale;
  const __half* aux;           // unused (no gate epilogue here)
  const __half* aux_lo;
  int aux_ld;
  __half* out;
  __half* out_lo;
  int out_ld;
  int* err_flag;
};

__global__ void __launch_bounds__(CR_THREADS, 1) conv_res_tcgen05_kernel(const __grid_constant__ ConvResParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_base = base + CR_OFF_W;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CR_OFF_BARS);
  uint64_t* empty_bar = full_bar + CR_STAGES;
  uint64_t* acc_full = empty_bar + CR_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint64_t* stg_full = acc_empty + 2;           // [2]
  uint64_t* stg_free = stg_full + 2;            // [2]
  uint64_t* w_bar = stg_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  EpiStage es;
  es.unit[0] = smem + CR_OFF_STAGING;
  es.unit[1] = smem + CR_OFF_STAGING + EPI_UNIT_BYTES;
  es.full = stg_full;
  es.free_ = stg_free;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_rows = p.n_rows_dev ? *p.n_rows_dev : p.n_rows;
  const int m_tiles = (n_rows + FC_TILE_M - 1) / FC_TILE_M;
  const int planes = p.split ? 2 : 1;
  const bool residual = p.epi == FC_EPI_ADD_RELU;
  const int aux_planes = p.has_aux_lo ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.a_map[1]);
    tma_prefetch_desc(&p.w_map);
    tma_prefetch_desc(&p.out_map[0]);
    if (p.out_lo) tma_prefetch_desc(&p.out_map[1]);
    if (residual) {
      tma_prefetch_desc(&p.aux_map[0]);
      tma_prefetch_desc(&p.aux_map[1]);
    }
    for (int s = 0; s < CR_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], FC_EPI_WARPS);
      mbar_init(&stg_full[s], FC_EPI_WARPS);
      mbar_init(&stg_free[s], 1);
    }
    mbar_init(w_bar, 1);
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

  // Both the producer and the MMA issuer walk the same per-M-tile sequence of ring tiles:
  //   for h in {0,1}: for iy in h..h+2: { for ix in 0..3: x tile (iy,ix) hi[, lo] ;
  //                                       for every output row oy of the half whose last input row is iy:
  //                                           (residual only) for ox in 0..3: aux tile (oy,ox) hi[, lo] }
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, one lane issues)
    if (blockIdx.x < m_tiles) {
      // resident weights first: planes x 9 tiles of 8 KB on one barrier
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(w_bar, uint32_t(planes * CR_W_PLANE_BYTES));
        for (int t = 0; t < planes * 9; ++t)
          tma_load_2d(smem + CR_OFF_W + t * CR_W_TILE_BYTES, &p.w_map, w_bar, 0, t * 64);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](const CUtensorMap* map, int pos, int mt) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag, 100 + stage);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], CR_A_BYTES);
          tma_load_2d(smem + stage * CR_A_BYTES, map, &full_bar[stage], pos * FC_TILE_K, mt * FC_TILE_M);
        }
        __syncwarp();
        if (++stage == CR_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      };
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        for (int h = 0; h < 2; ++h) {
          for (int iy = h; iy < h + 3; ++iy) {
            for (int ix = 0; ix < 4; ++ix)
              for (int pl = 0; pl < planes; ++pl) load(&p.a_map[pl], iy * 4 + ix, mt);
            if (residual) {
              for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
                if (iy != (oy < 3 ? oy + 1 : 3)) continue;
                for (int ox = 0; ox < 4; ++ox)
                  for (int pl = 0; pl < aux_planes; ++pl) load(&p.aux_map[pl], oy * 4 + ox, mt);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one lane issues)
    if (blockIdx.x < m_tiles) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_phase = 0u;               // bit s: parity of accumulator slot s
      const uint32_t idesc_id = umma_idesc_f16(16u);
      const uint64_t id_desc = ident_desc(base + CR_OFF_IDENT);
      auto acquire = [&]() -> uint32_t {     // next ring tile: wait until it has landed, return its smem address
        mbar_wait(&full_bar[stage], phase, p.err_flag, 300 + stage);
        tc_fence_after_sync();
        return base + stage * CR_A_BYTES;
      };
      auto advance = [&]() {
        if (++stage == CR_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      };
      mbar_wait(w_bar, 0u, p.err_flag, 500);
      tc_fence_after_sync();
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        for (int h = 0; h < 2; ++h) {
          uint32_t init_mask = 0u;               // bit 4*slot + ox: that output position's accumulator holds data
          uint32_t acquired = 0u;                // bit slot
          for (int iy = h; iy < h + 3; ++iy) {
            for (int ix = 0; ix < 4; ++ix) {
              const int ox0 = ix > 0 ? ix - 1 : 0;
              const int ox1 = ix < 3 ? ix + 1 : 3;
              for (int pl = 0; pl < planes; ++pl) {
                const uint32_t a_addr = acquire();
                // accumulator slots this tile feeds must have been drained by the epilogue
                for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
                  const int ky = iy - oy + 1;
                  const int slot = oy & 1;
                  if (ky < 0 || ky > 2 || ((acquired >> slot) & 1u)) continue;
                  mbar_wait(&acc_empty[slot], ((acc_phase >> slot) & 1u) ^ 1u, p.err_flag, 200 + slot);
                  tc_fence_after_sync();
                  acquired |= 1u << slot;
                }
                if (elect_one_sync()) {
                  for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
                    const int ky = iy - oy + 1;
                    if (ky < 0 || ky > 2) continue;
                    const int slot = oy & 1;
                    // the lo plane always accumulates (its hi twin ran first); the hi plane starts positions
                    // whose accumulator is still empty
                    const uint32_t im = pl ? 0xFu : ((init_mask >> (4 * slot)) & 0xFu);
                    // runs of output positions with the same accumulate state -> one MMA group each
                    int ox = ox0;
                    while (ox <= ox1) {
                      const uint32_t st = (im >> ox) & 1u;
                      int oe = ox;
                      while (oe + 1 <= ox1 && ((im >> (oe + 1)) & 1u) == st) ++oe;
                      const uint32_t n = uint32_t(oe - ox + 1) * 64u;
                      const uint32_t idesc = umma_idesc_f16(n);
                      const uint32_t d_tmem = tmem_base + uint32_t(slot * 256 + ox * 64);
                      // taps for ox..oe are kx = ix-ox+1 .. ix-oe+1 (descending) = rows (2-kx_first)*64.. of the ky stack
                      const uint32_t w_hi = w_base + uint32_t(ky * 3 + (2 - (ix - ox + 1))) * CR_W_TILE_BYTES;
                      const uint32_t w_lo = w_hi + CR_W_PLANE_BYTES;
#pragma unroll
                      for (int k = 0; k < FC_TILE_K / 16; ++k)
                        umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(w_hi + k * 32), idesc,
                                    (st || k > 0) ? 1u : 0u);
                      if (p.split && pl == 0) {
#pragma unroll
                        for (int k = 0; k < FC_TILE_K / 16; ++k)
                          umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(w_lo + k * 32), idesc, 1u);
                      }
                      ox = oe + 1;
                    }
                  }
                  umma_commit(&empty_bar[stage]);   // frees the ring slot once these MMAs have read it
                }
                __syncwarp();
                if (pl == 0) {
                  for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
                    const int ky = iy - oy + 1;
                    if (ky >= 0 && ky <= 2) init_mask |= ((1u << (ox1 + 1)) - (1u << ox0)) << (4 * (oy & 1));
                  }
                }
                advance();
              }
            }
            // output row complete once its last input row (oy + 1, clamped) has been consumed
            for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
              if (iy != (oy < 3 ? oy + 1 : 3)) continue;
              const int slot = oy & 1;
              if (residual) {
                for (int ox = 0; ox < 4; ++ox) {
                  for (int pl = 0; pl < aux_planes; ++pl) {
                    const uint32_t a_addr = acquire();
                    if (elect_one_sync()) {
                      const uint32_t d_tmem = tmem_base + uint32_t(slot * 256 + ox * 64);
#pragma unroll
                      for (int j = 0; j < FC_TILE_K / 16; ++j)
                        umma_f16_ss(d_tmem + j * 16, umma_desc_sw128(a_addr + j * 32), id_desc, idesc_id, 1u);
                      umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    advance();
                  }
                }
              }
              if (elect_one_sync()) umma_commit(&acc_full[slot]);
              __syncwarp();
              acc_phase ^= 1u << slot;
            }
          }
        }
      }
    }
  } else if (warp < FC_STORE_WARP) {
    // ------------------------------------------------------------ epilogue (warps 2..9): output rows 0..3 -> slots 0,1,0,1
    uint32_t acc_phase = 0u;
    uint32_t g = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      for (int oy = 0; oy < 4; ++oy) {
        const int slot = oy & 1;
        epi_tile_store(p, es, g, n_rows, mt, oy * 256, 256, tmem_base + uint32_t(slot * 256), &acc_full[slot], acc_phase,
                       &acc_empty[slot], warp, lane, 400 + slot);
        if (slot) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ store warp
    uint32_t g = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      if ((mt + 1) * FC_TILE_M > n_rows) continue;
      epi_store_chunks(es, g, &p.out_map[0], &p.out_map[1], p.out_lo != nullptr, 0, 1024 / EPI_CHUNK, mt * FC_TILE_M,
                       p.err_flag);
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace av1p
