// Stem on the tensor cores: block gather (straight from the planar 10-bit frame, or from a float
// block tensor) -> /1023 -> conv1 7x7 s2 p3 (BN folded) -> ReLU -> maxpool 3x3 s2 p1 -> fp16 hi/lo
// activation rows [block][4x4 positions][64 channels].
//
// Reference semantics: pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py:353-457 (tiling, zero
// pad bottom/right, row-major block order), pesquisa_v6/v6_pipeline/data_hub.py:70-77 (float32(u16) /
// 1023.0, true division) and models.py:105-108 (conv1 / bn1 / relu / maxpool).
//
// GEMM view, per tile of 4 blocks:   D[channel, (block, position)] = W[channel, tap] . patch[(block, position), tap]
//   M operand : folded conv1 weights, 64 channels x 64 K (49 taps + zero pad), stacked twice to fill M = 128,
//               fp16 hi and lo, resident in shared memory for the whole (persistent) kernel;
//   N operand : im2col rows (4 blocks x 64 conv positions = 256) x 64 K, fp16 hi and lo, built by the producer
//               warps directly in the SWIZZLE_128B K-major layout.  K is laid out as ky * 8 + kx (the kernel row
//               padded to eight taps, weight 0 for kx = 7 and for k >= 56), so one 16-byte K chunk of an im2col row
//               is eight CONSECUTIVE pixels of one row of the pixel tile: the tile is kept as fp16 hi / lo planes
//               (split once per pixel) and a chunk is four 32-bit shared loads per plane, no per-tap arithmetic;
//   products  : W_hi.P_hi + W_hi.P_lo + W_lo.P_hi (split precision), fp32 accumulators in TMEM (256 cols x 2).
//   INT_PIX   : frame input (kind 0).  A 10-bit sample is an integer that fp16 holds exactly, so the pixel tile keeps the
//               raw integers in ONE plane and the /1023 moves into a second weight set (w / 1023 folded in float64 by the
//               packer, split hi/lo): two products W_hi.P + W_lo.P instead of three and half the im2col traffic.
//               (Exact for samples <= 2048; a sample above that is rounded to fp16's 11 bits - such a value is outside
//               the 10-bit format the reference's reader warns about, 005:198-204.  The kernel raises a flag in the
//               plan's workspace and the Python frame entry points that synchronise turn it into an error.)
// Because the channels are the accumulator rows (TMEM lanes), one epilogue thread owns a channel and sees all
// 64 conv positions of a block in its columns: bias, ReLU and the 3x3/s2 max-pool run in registers, and a warp
// stores 32 consecutive channels (64 contiguous bytes) per pooled position.
#pragma once
#include "ptx_sm100.cuh"

namespace av1p {

constexpr int ST_BLOCKS = 4;                     // blocks per tile
constexpr int ST_N = ST_BLOCKS * 64;             // im2col rows per tile
constexpr int ST_STAGES = 2;
constexpr int ST_W_BYTES = 128 * 128;            // one weight plane (hi or lo): 128 rows x 128 B
constexpr int ST_P_BYTES = ST_N * 128;           // one patch plane: 256 rows x 128 B
constexpr int ST_STAGE_BYTES = 2 * ST_P_BYTES;   // hi + lo
constexpr int ST_TILE_H = 22, ST_TILE_W = 24;    // 16x16 block + 3-pixel zero halo (width padded)
constexpr int ST_PIX_PLANE = ST_BLOCKS * ST_TILE_H * ST_TILE_W;   // fp16 elements per plane
constexpr int ST_PIX_BYTES = 2 * 2 * ST_PIX_PLANE * 2;           // two buffers x (hi + lo) planes
constexpr int ST_PRODUCERS = 256;                // 8 warps (2 per scheduler) so the im2col LDS latency overlaps
constexpr int ST_EPI_WARPS = 8;                  // two per TMEM lane quadrant: one block of the pair each
constexpr int ST_THREADS = ST_PRODUCERS + 32 + 32 * ST_EPI_WARPS;   // producers, MMA warp, epilogue warps
constexpr int ST_SMEM_BYTES = 1024 + 2 * ST_W_BYTES + ST_STAGES * ST_STAGE_BYTES + ST_PIX_BYTES + 256;

struct StemInput {
  // kind 0: planar YUV 4:2:0 10-bit LE frames resident in HBM.  Block id g -> frame g / blocks_per_frame,
  //         grid row (g % bpf) / blocks_x, grid col (g % bpf) % blocks_x.
  // kind 1: float32 blocks [n][16*16] (the tensor HierarchicalPipelineV6.predict receives).
  int kind;
  const uint16_t* frames;
  long long frame_stride;   // elements between consecutive frames (Y + U + V)
  int width, height, pitch; // luma geometry, pitch in elements
  int n_frames;
  int blocks_x, blocks_per_frame;
  unsigned long long inv_bx, inv_bpf;   // floor(2^64 / d) + 1: __umul64hi(g, inv) == g / d for every 32-bit g (d > 1)
  const float* images;
};

struct StemParams {
  StemInput in;
  const int* idx;           // optional gather list: row r processes block id idx[r]
  const int* n_dev;         // device-side row count (nullptr -> n)
  int n;
  const __half* w;          // [2][128][64] folded conv1 weights x 2^s: hi plane then lo plane, K = ky*8+kx (kx = 7 and k >= 56 zero)
                            // (INT_PIX: the integer-pixel set, weights / 1023)
  const float* b;           // [64] folded bias
  float acc_scale;          // 2^-s
  __half* out;              // [rows][1024] in the tiled activation layout (act_off, 16 blocks per row)
  __half* out_lo;           // split precision: fp16(x - fp16(x)), nullptr otherwise
  int* err_flag;
  int* range_flag;          // optional: set to 1 when a frame sample above 2048 is met (INT_PIX)
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool INT_PIX>
__global__ void __launch_bounds__(ST_THREADS, 1) stem_tc_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t st_smem_raw[];
  const uint32_t base = (smem_u32(st_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = st_smem_raw + (base - smem_u32(st_smem_raw));
  uint8_t* w_hi = smem;                                   // [128][128 B] swizzled
  uint8_t* w_lo = smem + ST_W_BYTES;
  uint8_t* stages = smem + 2 * ST_W_BYTES;                // [ST_STAGES][hi | lo]
  // pixel tiles: [2 buffers][hi, lo][4 blocks][22][24] fp16; hi = fp16(x), lo = fp16(x - hi), 3-pixel zero halo
  __half* pix_hi = reinterpret_cast<__half*>(stages + ST_STAGES * ST_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(pix_hi) + ST_PIX_BYTES);
  uint64_t* empty_bar = full_bar + ST_STAGES;
  uint64_t* acc_full = empty_bar + ST_STAGES;             // [2]
  uint64_t* acc_empty = acc_full + 2;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // ---- one-time setup: weights into swizzled smem, zero halo, barriers, TMEM
  for (int q = threadIdx.x; q < 2 * 128 * 8; q += ST_THREADS) {
    const int plane = q >> 10, row = (q >> 3) & 127, c = q & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t(plane) * 128 + row) * 64 + c * 8));
    *reinterpret_cast<uint4*>((plane ? w_lo : w_hi) + row * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < ST_PIX_BYTES / 4; i += ST_THREADS) reinterpret_cast<uint32_t*>(pix_hi)[i] = 0u;   // halo stays zero
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(&full_bar[s], ST_PRODUCERS / 32);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], ST_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == ST_PRODUCERS / 32) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();       // the weight tiles were written with generic stores, tcgen05.mma reads them
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // frames / gather list / row count of the previous kernels are visible from here on; the output buffer is free
  const int n = p.n_dev ? *p.n_dev : p.n;
  const int tiles = (n + ST_BLOCKS - 1) / ST_BLOCKS;

  if (warp < ST_PRODUCERS / 32) {
    // ------------------------------------------------------------ producers: gather + im2col
    const int tid = threadIdx.x;                     // 0..ST_PRODUCERS-1
    const int c = tid & 7;                           // 16-byte K chunk this thread always writes = kernel row ky (7: zero pad)
    // pixel gather: 4 blocks x 16 rows x 4 quarter-rows = 256 work items of 4 samples, one per producer thread.
    // The raw samples of tile t+2 are in flight (registers) while tile t+1 is split into the spare pixel buffer and
    // tile t's im2col is built from the current one: one barrier per tile.
    const int pb = tid >> 6, ppy = (tid >> 2) & 15, ppx0 = (tid & 3) * 4;
    // raw[] holds either two packed pairs of 16-bit samples (mode 1, converted when the tile is consumed, so that
    // nothing waits on the load here) or four ready float bit patterns (mode 0).
    auto gather = [&](int tile, uint32_t (&raw)[4], int& mode) {
      mode = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) raw[j] = 0u;
      const int r = tile * ST_BLOCKS + pb;
      if (tile >= tiles || r >= n) return;
      const int g = p.idx ? __ldg(p.idx + r) : r;
      if (p.in.kind == 0) {
        // two divisions per thread per tile as multiply-high by the host's reciprocals (exact, see StemInput)
        const unsigned ug = unsigned(g);
        const int f = p.in.blocks_per_frame == 1 ? int(ug) : int(__umul64hi((unsigned long long)ug, p.in.inv_bpf));
        const int gb = g - f * p.in.blocks_per_frame;
        const int by = p.in.blocks_x == 1 ? gb : int(__umul64hi((unsigned long long)unsigned(gb), p.in.inv_bx));
        const int bx = gb - by * p.in.blocks_x;
        const int y = by * 16 + ppy, x0 = bx * 16 + ppx0;
        if (y < p.in.height) {
          const uint16_t* src = p.in.frames + size_t(f) * p.in.frame_stride + size_t(y) * p.in.pitch + x0;
          if (x0 + 3 < p.in.width && ((reinterpret_cast<uintptr_t>(src) & 7u) == 0)) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(src));
            raw[0] = q.x; raw[1] = q.y;
            mode = 1;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (x0 + j < p.in.width)
                raw[j] = __float_as_uint(INT_PIX ? float(__ldg(src + j)) : __fdiv_rn(float(__ldg(src + j)), 1023.0f));
          }
        }
      } else {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.in.images + size_t(g) * 256 + ppy * 16 + ppx0));
        raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
      }
    };
    // split the four samples into the fp16 hi / lo planes of pixel buffer `buf`
    auto write_pix = [&](int buf, const uint32_t (&raw)[4], int mode) {
      const int o = buf * 2 * ST_PIX_PLANE + (pb * ST_TILE_H + ppy + 3) * ST_TILE_W + ppx0 + 3;
      float x[4];
      if (mode && INT_PIX) {
        x[0] = float(raw[0] & 0xFFFFu);
        x[1] = float(raw[0] >> 16);
        x[2] = float(raw[1] & 0xFFFFu);
        x[3] = float(raw[1] >> 16);
      } else if (mode) {
        x[0] = __fdiv_rn(float(raw[0] & 0xFFFFu), 1023.0f);
        x[1] = __fdiv_rn(float(raw[0] >> 16), 1023.0f);
        x[2] = __fdiv_rn(float(raw[1] & 0xFFFFu), 1023.0f);
        x[3] = __fdiv_rn(float(raw[1] >> 16), 1023.0f);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = __uint_as_float(raw[j]);
      }
      if (INT_PIX && fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])) > 2048.0f && p.range_flag) {
        // outside the 10-bit format (the reference's reader warns above 1023, 005:198-204, and passes the value on): the
        // single integer plane is exact only up to 2048, so the caller is told (a word in the plan's workspace, see
        // av1p_cascade_buffer) instead of silently diverging from predict(images)
        *p.range_flag = 1;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {                   // o is odd: scalar fp16 stores
        const __half h = __float2half_rn(x[j]);
        pix_hi[o + j] = h;
        if (!INT_PIX) pix_hi[o + ST_PIX_PLANE + j] = __float2half_rn(x[j] - __half2float(h));
      }
    };
    int stage = 0, cur = 0;
    uint32_t phase = 0;
    uint32_t nxt[4];
    int nxt_mode;
    gather(blockIdx.x, nxt, nxt_mode);
    write_pix(0, nxt, nxt_mode);
    gather(blockIdx.x + gridDim.x, nxt, nxt_mode);
    named_bar_sync(1, ST_PRODUCERS);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      write_pix(cur ^ 1, nxt, nxt_mode);               // tile t+1 -> spare buffer
      gather(tile + 2 * gridDim.x, nxt, nxt_mode);     // tile t+2 in flight
      mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag, 500 + stage);
      uint8_t* s_hi = stages + stage * ST_STAGE_BYTES;
      uint8_t* s_lo = s_hi + ST_P_BYTES;
      const __half* ph_base = pix_hi + cur * 2 * ST_PIX_PLANE;
#pragma unroll
      for (int i = 0; i < (ST_N * 8) / ST_PRODUCERS; ++i) {
        const int row = (tid >> 3) + (ST_PRODUCERS / 8) * i;   // im2col row = block * 64 + conv position
        const int b = row >> 6, pos = row & 63;
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
        if (c < 7) {
          // kernel row c of conv position (py, px): tile row 2*py + c, tile columns 2*px .. 2*px + 7 (4-byte aligned)
          const int o = (b * ST_TILE_H + 2 * (pos >> 3) + c) * ST_TILE_W + 2 * (pos & 7);
          const uint32_t* ph = reinterpret_cast<const uint32_t*>(ph_base + o);
          hi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          if (!INT_PIX) {
            const uint32_t* pl = reinterpret_cast<const uint32_t*>(ph_base + ST_PIX_PLANE + o);
            lo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          }
        }
        const int dst = row * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(s_hi + dst) = hi;
        if (!INT_PIX) *reinterpret_cast<uint4*>(s_lo + dst) = lo;
      }
      fence_proxy_async_smem();                       // generic-proxy stores -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
      named_bar_sync(1, ST_PRODUCERS);                // tile t+1's pixels are complete; nobody still reads buffer `cur`
      cur ^= 1;
      if (++stage == ST_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == ST_PRODUCERS / 32) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one lane issues)
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_f16(ST_N);
    const uint32_t a_hi = umma_desc_lo_sw128(smem_u32(w_hi)), a_lo = umma_desc_lo_sw128(smem_u32(w_lo));
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1u, p.err_flag, 600 + acc);
      mbar_wait(&full_bar[stage], phase, p.err_flag, 700 + stage);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t d_tmem = tmem_base + uint32_t(acc * ST_N);
        const uint32_t b_hi = umma_desc_lo_sw128(smem_u32(stages + stage * ST_STAGE_BYTES)), b_lo = b_hi + (ST_P_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss_lo(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, k > 0 ? 1u : 0u);
        if (!INT_PIX) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss_lo(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss_lo(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
        umma_commit(&empty_bar[stage]);
        umma_commit(&acc_full[acc]);
      }
      __syncwarp();
      if (++stage == ST_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: bias, ReLU, max-pool, hi/lo stores
    const int quad = warp & 3;
    const int ch = (quad & 1) * 32 + lane;            // accumulator row -> channel (rows 64..127 repeat 0..63)
    const int blk0 = (quad >> 1) * 2;                 // rows 0..63 take blocks 0,1 of the tile, rows 64..127 blocks 2,3
    const int bi = (warp - (ST_PRODUCERS / 32 + 1)) >> 2;   // which block of that pair this warp handles
    const float bias = p.b[ch];
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      mbar_wait(&acc_full[acc], acc_phase, p.err_flag, 800 + acc);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * ST_N);
      {
        const int blk = blk0 + bi;
        const int r = tile * ST_BLOCKS + blk;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(t_addr + uint32_t(blk * 64), v0);        // conv rows 0..3
        tmem_ld_32x32(t_addr + uint32_t(blk * 64 + 32), v1);   // conv rows 4..7
        tmem_ld_wait();
        if (r < n) {
        // max-pool the raw accumulators first: relu(s * a + b) is monotonic in a (s = 2^-k > 0), so
        // max_window relu(s * a + b) = relu(s * max_window a + b) - 16 affine + ReLU evaluations instead of 64, and the
        // 3x3 window is separable (row maxima, then column maxima)
        float rm[8][4];
#pragma unroll
        for (int y = 0; y < 8; ++y) {
#pragma unroll
          for (int qx = 0; qx < 4; ++qx) {
            auto at = [&](int x) { return __uint_as_float(y < 4 ? v0[y * 8 + x] : v1[(y - 4) * 8 + x]); };
            float m = fmaxf(at(2 * qx), at(2 * qx + 1));
            if (qx > 0) m = fmaxf(m, at(2 * qx - 1));
            rm[y][qx] = m;
          }
        }
        // tiled activation layout (act_off): position q is the 64-column block q of row r
        const size_t obase = act_off(r, ch, 16);
        __half* o = p.out + obase;
        __half* ol = p.out_lo ? p.out_lo + obase : nullptr;
#pragma unroll
        for (int qy = 0; qy < 4; ++qy) {
#pragma unroll
          for (int qx = 0; qx < 4; ++qx) {
            float a = fmaxf(rm[2 * qy][qx], rm[2 * qy + 1][qx]);
            if (qy > 0) a = fmaxf(a, rm[2 * qy - 1][qx]);
            const float m = fmaxf(fmaf(a, p.acc_scale, bias), 0.f);
            const __half h = __float2half_rn(m);
            o[size_t(qy * 4 + qx) << 13] = h;
            if (ol) ol[size_t(qy * 4 + qx) << 13] = __float2half_rn(m - __half2float(h));
          }
        }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == ST_PRODUCERS / 32) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace av1p
