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
//   warps 2..9  epilogue     : tcgen05.ld -> bias / residual / ReLU / gate -> fp16 rows to global
//                              (two warps per TMEM lane quadrant, each takes half of the tile's columns)
// with a 4-deep smem ring (full/empty mbarriers) and a 2-deep TMEM accumulator ring.
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
constexpr int FC_MAX_NT = 8;           // N tiles per layer
constexpr int FC_MAX_KB = 128;         // scheduled K blocks per layer (sum over N tiles)
constexpr int FC_MAX_SRC = 4;          // activation sources per layer
constexpr int FC_TAIL_MAX = 4;         // outputs of the in-epilogue final linear
constexpr int FC_EPI_WARPS = 8;
constexpr int FC_THREADS = 64 + 32 * FC_EPI_WARPS;
constexpr int FC_SMEM_BYTES = FC_STAGES * FC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                              FC_TAIL_MAX * FC_MAX_N * 4 /*tail weights*/;

// Precision.  Operands are fp16, accumulation is fp32 in TMEM.  In split mode ("fp16x3") every
// activation x is stored as hi = fp16(x), lo = fp16(x - hi), every weight as hi/lo likewise, and the
// schedule lists three products per K block: (x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi) - about 22
// significant bits, i.e. fp32-grade logits, on the fp16 tensor pipe.  Weights are pre-scaled by a power
// of two per layer (acc_scale undoes it) so that the lo parts stay in fp16's normal range.
enum FcEpilogue : int {
  FC_EPI_LINEAR = 0,    // out = s*acc + b
  FC_EPI_RELU = 1,      // out = relu(s*acc + b)
  FC_EPI_ADD_RELU = 2,  // out = relu(acc + b + aux)            (identity residual)
  FC_EPI_GATE = 3,      // out = aux * sigmoid(acc)             (SE excitation)
  FC_EPI_HEAD = 4,      // h = relu(s*acc + b); logits = h . tail_w^T + tail_b   (fp32, no fp16 store)
};

struct FcParams {
  CUtensorMap a_map[FC_MAX_SRC]; // activation sources, 2-D [rows][K] fp16, box {64, 128}, SWIZZLE_128B
  CUtensorMap w_map;           // packed weights, 2-D [n_kb_total*block_n][64] fp16, box {64, block_n}
  const int* n_rows_dev;       // device-side row count (nullptr -> n_rows)
  int n_rows;
  int n_tiles;                 // number of N tiles
  int block_n;                 // tile N (multiple of 32, <= 256)
  int epi;
  const float* bias;           // [n_tiles*block_n] or nullptr
  const float* row_scale;      // [rows] or nullptr (spatial-attention scalar folded into the next linear)
  float acc_scale;             // power of two undoing the weight pre-scale
  const __half* aux;           // residual / gate input rows
  const __half* aux_lo;        // split mode: low part of aux (nullptr otherwise)
  int aux_ld;
  __half* out;
  __half* out_lo;              // split mode: low part of the output (nullptr otherwise)
  int out_ld;
  const float* tail_w;         // [tail_n][block_n]
  const float* tail_b;         // [tail_n]
  float* logits;               // [rows][tail_n]
  int tail_n;
  int pair_mode;               // split precision: schedule entries come in pairs e = (x_hi, w_hi), e+1 = (x_lo, w_lo)
                               // of one K block; the MMA warp issues hi*hi, hi*lo, lo*hi (each tile is loaded once)
  int* err_flag;
  int kb_begin[FC_MAX_NT + 1]; // schedule range of each N tile
  uint16_t kb_src[FC_MAX_KB];  // bits 14..15: activation source, bits 0..13: K offset / 64
  uint16_t kb_w[FC_MAX_KB];    // weight chunk ([block_n x 64] tile) index
};

__device__ __forceinline__ float fast_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// 32 consecutive fp16 of a row (64 B) as four 16-byte loads
struct Half32 {
  uint4 q[4];
};
__device__ __forceinline__ void ld_half32(Half32& d, const __half* p) {
  const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) d.q[i] = __ldg(s + i);
}
__device__ __forceinline__ void zero_half32(Half32& d) {
#pragma unroll
  for (int i = 0; i < 4; ++i) d.q[i] = make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void add_half32(float (&f)[32], const Half32& d) {
  const __half2* h = reinterpret_cast<const __half2*>(d.q);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 v = __half22float2(h[i]);
    f[2 * i] += v.x;
    f[2 * i + 1] += v.y;
  }
}

// Epilogue of one accumulator tile [128 rows x block_n fp32 columns at TMEM column t_col]: warps 2..9,
// two warps per TMEM lane quadrant, each takes half of the tile's columns.  Waits on `full`, applies
// bias / residual / ReLU / gate / in-thread final linear, stores fp16 hi (+lo) rows, then arrives on `empty`.
// P is FcParams or any struct with the same epilogue members.
template <typename P>
__device__ __forceinline__ void fc_epilogue_tile(const P& p, int n_rows, int mt, int col0, int block_n, uint32_t t_col,
                                                 uint64_t* full, uint32_t full_phase, uint64_t* empty,
                                                 const float* tail_w_s, int warp, int lane, int tag) {
  const int ew = warp - 2;
  const int quad = warp & 3;              // TMEM lane quadrant this warp may read
  const int half = ew >> 2;               // which half of the tile's columns
  // the in-thread head needs whole rows: there the first four warps take every column
  const bool whole = (p.epi == FC_EPI_HEAD) || (block_n < 64);
  const int c_begin = whole ? 0 : half * (block_n / 2);
  const int c_end = whole ? (half == 0 ? block_n : 0) : c_begin + block_n / 2;
  const bool needs_aux = (p.epi == FC_EPI_ADD_RELU || p.epi == FC_EPI_GATE);
  const int row = mt * FC_TILE_M + quad * 32 + lane;
  const bool row_ok = row < n_rows;
  const float rs = ((p.row_scale && row_ok) ? p.row_scale[row] : 1.0f) * p.acc_scale;
  // prefetch the first residual / gate chunk while the MMAs of this tile are still running
  Half32 ax, axl;
  zero_half32(ax);
  zero_half32(axl);
  if (needs_aux && row_ok && c_begin < c_end) {
    ld_half32(ax, p.aux + size_t(row) * p.aux_ld + col0 + c_begin);
    if (p.aux_lo) ld_half32(axl, p.aux_lo + size_t(row) * p.aux_ld + col0 + c_begin);
  }
  mbar_wait(full, full_phase, p.err_flag, tag);
  tc_fence_after_sync();
  const uint32_t t_addr = t_col + (uint32_t(quad * 32) << 16);
  float tail[FC_TAIL_MAX] = {0.f, 0.f, 0.f, 0.f};
  for (int c = c_begin; c < c_end; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(t_addr + uint32_t(c), v);
    // next chunk's residual / gate input goes in flight before this chunk is consumed
    Half32 nx, nxl;
    zero_half32(nx);
    zero_half32(nxl);
    if (needs_aux && row_ok && c + 32 < c_end) {
      ld_half32(nx, p.aux + size_t(row) * p.aux_ld + col0 + c + 32);
      if (p.aux_lo) ld_half32(nxl, p.aux_lo + size_t(row) * p.aux_ld + col0 + c + 32);
    }
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * rs;
    if (p.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b4 + i);
        f[4 * i + 0] += b.x;
        f[4 * i + 1] += b.y;
        f[4 * i + 2] += b.z;
        f[4 * i + 3] += b.w;
      }
    }
    if (p.epi == FC_EPI_ADD_RELU) {
      add_half32(f, ax);
      add_half32(f, axl);
    } else if (p.epi == FC_EPI_GATE) {
      float g[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = 0.f;
      add_half32(g, ax);
      add_half32(g, axl);
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = g[i] * fast_sigmoid(f[i]);
    }
    ax = nx;
    axl = nxl;
    if (p.epi == FC_EPI_RELU || p.epi == FC_EPI_ADD_RELU || p.epi == FC_EPI_HEAD) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (p.epi == FC_EPI_HEAD) {
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
    } else if (row_ok) {
      __align__(16) __half h[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) h[i] = __float2half_rn(f[i]);
      uint4* o4 = reinterpret_cast<uint4*>(p.out + size_t(row) * p.out_ld + col0 + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) o4[i] = reinterpret_cast<const uint4*>(h)[i];
      if (p.out_lo) {
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i] = __float2half_rn(f[i] - __half2float(h[i]));
        uint4* l4 = reinterpret_cast<uint4*>(p.out_lo + size_t(row) * p.out_ld + col0 + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) l4[i] = reinterpret_cast<const uint4*>(h)[i];
      }
    }
  }
  if (p.epi == FC_EPI_HEAD && row_ok && c_begin < c_end) {
#pragma unroll
    for (int j = 0; j < FC_TAIL_MAX; ++j)
      if (j < p.tail_n) p.logits[size_t(row) * p.tail_n + j] = tail[j] + p.tail_b[j];
  }
  // all TMEM reads of this accumulator are complete (tmem_ld_wait above) -> hand it back
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(empty);
}

__global__ void __launch_bounds__(FC_THREADS, 1) fc_tcgen05_kernel(const __grid_constant__ FcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + FC_STAGES * FC_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + FC_STAGES;
  uint64_t* acc_full = empty_bar + FC_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* tail_w_s = reinterpret_cast<float*>(smem + FC_STAGES * FC_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_rows = p.n_rows_dev ? *p.n_rows_dev : p.n_rows;
  const int m_tiles = (n_rows + FC_TILE_M - 1) / FC_TILE_M;
  const int n_items = m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < FC_MAX_SRC; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.w_map);
    for (int s = 0; s < FC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], FC_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (p.epi == FC_EPI_HEAD) {
    for (int i = threadIdx.x; i < p.tail_n * p.block_n; i += FC_THREADS) tail_w_s[i] = p.tail_w[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = FC_A_BYTES + p.block_n * FC_TILE_K * 2;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / p.n_tiles;
        const int nt = item - mt * p.n_tiles;
        for (int kb = p.kb_begin[nt]; kb < p.kb_begin[nt + 1]; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag, 100 + stage);
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          const uint32_t e = p.kb_src[kb];
          uint8_t* a_dst = smem + stage * FC_STAGE_BYTES;
          tma_load_2d(a_dst, &p.a_map[e >> 14], &full_bar[stage], int(e & 0x3FFFu) * FC_TILE_K, mt * FC_TILE_M);
          tma_load_2d(a_dst + FC_A_BYTES, &p.w_map, &full_bar[stage], 0, int(p.kb_w[kb]) * p.block_n);
          if (++stage == FC_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_f16(uint32_t(p.block_n));
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
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
          const uint32_t a_addr = base + stage * FC_STAGE_BYTES;
          const uint32_t w_addr = a_addr + FC_A_BYTES;
          if (!p.pair_mode) {
#pragma unroll
            for (int k = 0; k < FC_TILE_K / 16; ++k)
              umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc,
                          (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
          } else if (((kb - kb0) & 1) == 0) {
            // (x_hi, w_hi): the slot stays live until the cross products of the next entry are done
#pragma unroll
            for (int k = 0; k < FC_TILE_K / 16; ++k)
              umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc,
                          (kb > kb0 || k > 0) ? 1u : 0u);
            prev_a = a_addr;
            prev_w = w_addr;
            prev_stage = stage;
          } else {
            // this slot holds (x_lo, w_lo): issue x_hi * w_lo and x_lo * w_hi
#pragma unroll
            for (int k = 0; k < FC_TILE_K / 16; ++k)
              umma_f16_ss(d_tmem, umma_desc_sw128(prev_a + k * 32), umma_desc_sw128(w_addr + k * 32), idesc, 1u);
#pragma unroll
            for (int k = 0; k < FC_TILE_K / 16; ++k)
              umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(prev_w + k * 32), idesc, 1u);
            umma_commit(&empty_bar[prev_stage]);
            umma_commit(&empty_bar[stage]);
          }
          if (++stage == FC_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&acc_full[acc]);        // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = item / p.n_tiles;
      const int nt = item - mt * p.n_tiles;
      fc_epilogue_tile(p, n_rows, mt, nt * p.block_n, p.block_n, tmem_base + uint32_t(acc * FC_MAX_N), &acc_full[acc],
                       acc_phase, &acc_empty[acc], tail_w_s, warp, lane, 400 + acc);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace av1p
