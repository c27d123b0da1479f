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

// The per-M-tile work list is static, so the host compiles it once (conv_res_build_schedule) and the device
// roles only interpret it: `ring[i]` = which [128 x 64] activation tile the producer loads i-th (map << 4 | position),
// `tab[j]` = one MMA group (four K=16 instructions on the current ring tile):
//   bits  0..13  B operand: smem offset / 16 from the aligned base (a run of 64-row tap tiles of the resident weights)
//   bits 14..22  accumulator column (slot * 256 + ox * 64)
//   bit  23      identity group: 4 x (N = 16) instructions spreading the tile over 64 columns (residual branch)
//   bit  24      accumulate flag of the first instruction (0 = the group starts these output positions)
//   bits 25..26  N / 64 - 1
//   bit  27      first group of a ring tile (wait for the tile)     bit 28  last group of a ring tile (release it)
//   bit  29/30   output row in accumulator slot 0/1 is complete after this group
//   bit  31      first use of the accumulator slot in this half (wait until the epilogue has drained it)
constexpr int CR_MAX_GROUPS = 176;
constexpr int CR_MAX_RING = 80;
constexpr uint32_t CR_G_IDENT = 1u << 23, CR_G_ACC = 1u << 24, CR_G_FIRST = 1u << 27, CR_G_LAST = 1u << 28,
                   CR_G_DONE0 = 1u << 29, CR_G_DONE1 = 1u << 30, CR_G_NEED_ACC = 1u << 31;

struct ConvResParams {
  CUtensorMap a_map[4];        // x_hi, x_lo, residual hi, residual lo: 2-D [rows][1024] fp16, box {64, 128}, SWIZZLE_128B
  CUtensorMap w_map;           // resident weights: 2-D [planes*9*64][64] fp16, box {64, 64}; tile (plane, ky, 2-kx)
  CUtensorMap out_map[2];      // output hi, lo: 2-D [rows][1024] fp16, box {32, 128}, SWIZZLE_64B
  const int* n_rows_dev;
  int n_rows;
  int split;                   // 1: hi/lo planes, three products; 0: single fp16 product
  int has_aux_lo;              // residual has a lo plane
  int n_ring, n_groups;
  uint8_t ring[CR_MAX_RING];
  uint32_t tab[CR_MAX_GROUPS];
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
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CR_OFF_BARS);
  uint64_t* empty_bar = full_bar + CR_STAGES;
  uint64_t* acc_full = empty_bar + CR_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint64_t* stg_full = acc_empty + 2;           // [2]
  uint64_t* stg_free = stg_full + 2;            // [2]
  uint64_t* w_bar = stg_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  EpiStage es;
  epi_stage_init(es, smem + CR_OFF_STAGING, stg_full, stg_free);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_rows = p.n_rows_dev ? *p.n_rows_dev : p.n_rows;
  const int m_tiles = (n_rows + FC_TILE_M - 1) / FC_TILE_M;
  const int planes = p.split ? 2 : 1;
  const bool residual = p.epi == FC_EPI_ADD_RELU;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.a_map[1]);
    tma_prefetch_desc(&p.w_map);
    tma_prefetch_desc(&p.out_map[0]);
    if (p.out_lo) tma_prefetch_desc(&p.out_map[1]);
    if (residual) {
      tma_prefetch_desc(&p.a_map[2]);
      tma_prefetch_desc(&p.a_map[3]);
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
      // L2 prefetch cursor: runs CR_PREFETCH ring tiles ahead of the loads (into the next M tile of this CTA)
      int pf_mt = blockIdx.x, pf_i = 0;
      auto prefetch_next = [&]() {
        if (pf_mt >= m_tiles) return;
        const uint32_t e = p.ring[pf_i];
        if (elect_one_sync()) tma_prefetch_l2_2d(&p.a_map[e >> 4], int(e & 15u) * FC_TILE_K, pf_mt * FC_TILE_M);
        __syncwarp();
        if (++pf_i == p.n_ring) {
          pf_i = 0;
          pf_mt += gridDim.x;
        }
      };
      for (int i = 0; i < CR_PREFETCH; ++i) prefetch_next();
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        for (int i = 0; i < p.n_ring; ++i) {
          const uint32_t e = p.ring[i];
          prefetch_next();
          mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag, 100 + stage);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[stage], CR_A_BYTES);
            tma_load_2d(smem + stage * CR_A_BYTES, &p.a_map[e >> 4], &full_bar[stage], int(e & 15u) * FC_TILE_K, mt * FC_TILE_M);
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
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one lane issues)
    if (blockIdx.x < m_tiles) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_phase = 0u;               // bit s: parity of accumulator slot s
      const uint32_t idesc_id = umma_idesc_f16(16u);
      const uint64_t id_desc = ident_desc(base + CR_OFF_IDENT);
      mbar_wait(w_bar, 0u, p.err_flag, 500);
      tc_fence_after_sync();
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        uint32_t a_addr = 0;
        for (int gi = 0; gi < p.n_groups; ++gi) {
          const uint32_t w = p.tab[gi];
          const uint32_t d_col = (w >> 14) & 0x1FFu;
          const uint32_t slot = d_col >> 8;
          if (w & CR_G_FIRST) {
            mbar_wait(&full_bar[stage], phase, p.err_flag, 300 + stage);
            a_addr = base + stage * CR_A_BYTES;
          }
          if (w & CR_G_NEED_ACC) mbar_wait(&acc_empty[slot], ((acc_phase >> slot) & 1u) ^ 1u, p.err_flag, 200 + slot);
          tc_fence_after_sync();
          if (elect_one_sync()) {
            const uint32_t d_tmem = tmem_base + d_col;
            if (w & CR_G_IDENT) {
#pragma unroll
              for (int j = 0; j < FC_TILE_K / 16; ++j)
                umma_f16_ss(d_tmem + j * 16, umma_desc_sw128(a_addr + j * 32), id_desc, idesc_id, 1u);
            } else {
              const uint32_t b_addr = base + ((w & 0x3FFFu) << 4);
              const uint32_t idesc = umma_idesc_f16((((w >> 25) & 3u) + 1u) * 64u);
              const uint32_t acc0 = (w >> 24) & 1u;
#pragma unroll
              for (int k = 0; k < FC_TILE_K / 16; ++k)
                umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                            k > 0 ? 1u : acc0);
            }
            if (w & CR_G_LAST) umma_commit(&empty_bar[stage]);   // frees the ring slot once these MMAs have read it
            if (w & CR_G_DONE0) umma_commit(&acc_full[0]);
            if (w & CR_G_DONE1) umma_commit(&acc_full[1]);
          }
          __syncwarp();
          if (w & CR_G_LAST) {
            if (++stage == CR_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          acc_phase ^= (w >> 29) & 3u;
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
    epi_store_drain();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// Host side: compile the static per-M-tile schedule (see ConvResParams) for the given precision / residual shape.
// Order: two halves (output rows {0,1}, {2,3}); per half the three input rows it needs; per input row the
// positions ix = 1, 0, 2, 3 (ix = 1 first so that a fresh output row starts with one N = 192 group); hi plane then
// lo plane; after the last input row of an output row its residual tiles; groups of one ring tile are contiguous.
inline bool conv_res_build_schedule(ConvResParams& f) {
  const int planes = f.split ? 2 : 1;
  const bool residual = f.epi == FC_EPI_ADD_RELU;
  const int aux_planes = f.has_aux_lo ? 2 : 1;
  int n_ring = 0, n_groups = 0;
  auto push_group = [&](uint32_t w) -> bool {
    if (n_groups >= CR_MAX_GROUPS) return false;
    f.tab[n_groups++] = w;
    return true;
  };
  static const int ix_order[4] = {1, 0, 2, 3};
  for (int h = 0; h < 2; ++h) {
    uint32_t init_mask[2] = {0u, 0u};
    bool touched[2] = {false, false};
    for (int iy = h; iy < h + 3; ++iy) {
      for (int xi = 0; xi < 4; ++xi) {
        const int ix = ix_order[xi];
        const int ox0 = ix > 0 ? ix - 1 : 0, ox1 = ix < 3 ? ix + 1 : 3;
        for (int pl = 0; pl < planes; ++pl) {
          if (n_ring >= CR_MAX_RING) return false;
          f.ring[n_ring++] = uint8_t((pl << 4) | (iy * 4 + ix));
          const int g_first = n_groups;
          for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
            const int ky = iy - oy + 1;
            if (ky < 0 || ky > 2) continue;
            const int slot = oy & 1;
            uint32_t need = 0u;
            if (!touched[slot]) {
              need = CR_G_NEED_ACC;
              touched[slot] = true;
            }
            const uint32_t im = pl ? 0xFu : init_mask[slot];
            int ox = ox0;
            while (ox <= ox1) {
              const uint32_t st = (im >> ox) & 1u;
              int oe = ox;
              while (oe + 1 <= ox1 && ((im >> (oe + 1)) & 1u) == st) ++oe;
              const uint32_t n64 = uint32_t(oe - ox + 1);
              const uint32_t d_col = uint32_t(slot * 256 + ox * 64);
              // taps for ox..oe are kx = ix-ox+1 .. ix-oe+1 (descending) = tiles (2-kx_first).. of the ky stack
              const uint32_t tile = uint32_t(ky * 3 + (2 - (ix - ox + 1)));
              const uint32_t w_hi = (uint32_t(CR_OFF_W) + tile * CR_W_TILE_BYTES) >> 4;
              const uint32_t w_lo = (uint32_t(CR_OFF_W) + CR_W_PLANE_BYTES + tile * CR_W_TILE_BYTES) >> 4;
              const uint32_t common = (d_col << 14) | ((n64 - 1u) << 25);
              if (!push_group(w_hi | common | (st ? CR_G_ACC : 0u) | need)) return false;
              need = 0u;
              if (f.split && pl == 0)
                if (!push_group(w_lo | common | CR_G_ACC)) return false;
              ox = oe + 1;
            }
            if (pl == 0) init_mask[slot] |= (1u << (ox1 + 1)) - (1u << ox0);
          }
          f.tab[g_first] |= CR_G_FIRST;
          f.tab[n_groups - 1] |= CR_G_LAST;
        }
      }
      for (int oy = 2 * h; oy < 2 * h + 2; ++oy) {
        if (iy != (oy < 3 ? oy + 1 : 3)) continue;
        const int slot = oy & 1;
        if (residual) {
          for (int ox = 0; ox < 4; ++ox) {
            for (int pl = 0; pl < aux_planes; ++pl) {
              if (n_ring >= CR_MAX_RING) return false;
              f.ring[n_ring++] = uint8_t(((2 + pl) << 4) | (oy * 4 + ox));
              if (!push_group((uint32_t(slot * 256 + ox * 64) << 14) | CR_G_IDENT | CR_G_ACC | CR_G_FIRST | CR_G_LAST)) return false;
            }
          }
        }
        f.tab[n_groups - 1] |= slot ? CR_G_DONE1 : CR_G_DONE0;
      }
    }
  }
  f.n_ring = n_ring;
  f.n_groups = n_groups;
  return true;
}

}  // namespace av1p
