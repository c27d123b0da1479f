// Stem for FRAME input with TMA staging: the 16x16 luma blocks are fetched from the planar 10-bit frame by the TMA
// engine (one 16 x 16 box per block, straight into shared memory; out-of-bounds rows / columns arrive as zeros, which is
// the reference's bottom / right padding, 005:380-383), and the im2col operand is built from those RAW tiles without any
// conversion.  Same GEMM and epilogue as stem_tc.cuh (channels are the accumulator rows, 4 blocks = 256 im2col rows per
// tile, bias + ReLU + 3x3/s2 max-pool in registers), see there for the reference semantics.
//
// No conversion: the bit pattern of a 10-bit sample v, read as fp16, IS the number v * 2^-24 (fp16 subnormals for v < 1024,
// the first normal binade for 1024..2047 - both with ulp 2^-24; tcgen05.mma does not flush subnormal operands,
// tests/test_gpu_fc_kernel.py).  So the raw uint16 words are a valid fp16 operand plane and the factor 2^24 / 1023 lives in
// the weights ("raw" weight set of the stem op: w / 1023 * 2^s, hi + lo planes, acc_scale = 2^(24 - s)).  A sample of
// 2048 or more is outside the linear range (and outside the 10-bit format, 005:198-204): the builders test the words they
// copy and convert such samples explicitly (fp16(v) * 2^-24, i.e. the integer-pixel kernel's 11-bit rounding), raising
// the plan's range flag above 2048 exactly like that kernel does.
//
// K layout: k = ky * 8 + kxx, kxx = kx + 1 (kxx = 0 and ky = 7 carry zero weights), i.e. the 16-byte K chunk `ky` of conv
// position (py, px) is the eight consecutive samples x = 2 px - 4 .. 2 px + 3 of row y = 2 py + ky - 3 of the block: four
// 4-byte-aligned words of the raw tile (words outside the block are the conv's zero padding and are masked).
//
// Roles (18 warps): warp 0 issues the TMA loads (lanes 0..3: one block each) into an 8-deep ring of raw tiles; warps 2..9
// copy raw words into the SWIZZLE_128B operand tiles (4 stages); warp 1 issues the MMAs; warps 10..17 run the epilogue.
// Every hand-over is an mbarrier, so loading, im2col, MMA and epilogue of four different tiles overlap (the integer-pixel
// kernel stem_tc.cuh runs gather + split + im2col in one warp group with a CTA-wide named barrier per tile).
#pragma once
#include <cuda.h>
#include "stem_tc.cuh"

namespace av1p {

constexpr int SM_OP_STAGES = 4;                      // operand (im2col) tiles: one plane of 256 rows x 128 B
constexpr int SM_RAW_STAGES = 8;                     // raw pixel tiles: 4 blocks x 16 x 16 samples
constexpr int SM_RAW_BYTES = ST_BLOCKS * 512;
constexpr int SM_BUILD_WARPS = 8;
constexpr int SM_THREADS = 64 + 32 * SM_BUILD_WARPS + 32 * ST_EPI_WARPS;
constexpr int SM_OFF_OPS = 2 * ST_W_BYTES;
constexpr int SM_OFF_RAW = SM_OFF_OPS + SM_OP_STAGES * ST_P_BYTES;
constexpr int SM_OFF_BARS = SM_OFF_RAW + SM_RAW_STAGES * SM_RAW_BYTES;
constexpr int SM_SMEM_BYTES = 1024 + SM_OFF_BARS + 256;
static_assert((2 * SM_RAW_STAGES + 2 * SM_OP_STAGES + 4) * 8 + 4 <= 256, "barrier area");
static_assert(SM_SMEM_BYTES <= 232448, "stem_tma shared memory exceeds the 227 KB opt-in limit");

struct StemTmaParams {
  CUtensorMap map;          // frames as uint16 [n_frames][height][width] (strides: frame_stride, pitch).  WIDE = false: box
                            // {16, 16, 1} (one block), no swizzle; WIDE = true: box {64, 16, 1} (the four horizontally adjacent
                            // blocks of a tile, 128-byte rows), SWIZZLE_128B
  StemParams s;             // s.in.kind == 0; s.w = the raw weight set [2][128][64]; s.acc_scale = 2^(24 - s)
  int debug;                // development switch (AV1P_STEM_DEBUG): 1 skip the MMAs, 2 skip the epilogue's math and stores,
                            // 4 skip the im2col copy, 8 skip the TMA loads - isolates what paces the kernel
};

// Slow path of a builder thread (kept out of line so that none of it is if-converted into the copy loop): re-reads the
// eight chunks the thread has just stored, converts every sample >= 2048 like the integer-pixel kernel does - fp16(v),
// i.e. rounded to 11 bits, times the operand's 2^-24 (exponent field - 24; fp16(v) >= 2^11 keeps the result normal) - and
// zeroes the chunks of blocks past the end of the list.  Returns non-zero when a sample ABOVE 2048 was met (range flag).
__device__ __noinline__ uint32_t stem_tma_fix_tile(uint8_t* dst, int n_blk) {
  uint32_t bad = 0u;
  for (int i = 0; i < 8; ++i) {
    const int b = i >> 1, h = i & 1;
    uint4* slot = reinterpret_cast<uint4*>(dst + (b * 64 + h * 32) * 128);
    uint4 v = *slot;
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int m = 0; m < 4; ++m) {
      if (b >= n_blk) {
        w[m] = 0u;
        continue;
      }
      bad |= __vcmpgtu2(w[m], 0x08000800u);
      uint32_t hw[2] = {w[m] & 0xFFFFu, w[m] >> 16};
      for (int j = 0; j < 2; ++j)
        if (hw[j] >= 2048u) {
          const uint32_t bits = __half_as_ushort(__ushort2half_rn(static_cast<unsigned short>(hw[j])));
          hw[j] = bits >= 0x7C00u ? 0x7C00u : bits - 0x6000u;
        }
      w[m] = hw[0] | (hw[1] << 16);
    }
    *slot = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return bad;
}

// WIDE (opt-in): unrouted input (no gather list) whose block rows hold a multiple of four blocks - the four blocks of a tile
// are neighbours in the frame, so ONE box of 16 rows x 128 bytes fetches them (a 16-sample box row is a 32-byte request;
// with everything but the loads switched off the per-block boxes take 350 us for 518 k blocks, the wide boxes 221 us).  The
// loads are not what paces the kernel though, and the swizzled source addressing costs the builders two more instructions
// per word: measured equal or slower than the per-block variant, which is the default.
template <bool WIDE>
__global__ void __launch_bounds__(SM_THREADS, 1) stem_tma_kernel(const __grid_constant__ StemTmaParams q) {
  const StemParams& p = q.s;
  extern __shared__ uint8_t sm_smem_raw[];
  const uint32_t base = (smem_u32(sm_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = sm_smem_raw + (base - smem_u32(sm_smem_raw));
  uint8_t* w_hi = smem;                                   // [128][128 B] swizzled
  uint8_t* w_lo = smem + ST_W_BYTES;
  uint8_t* ops = smem + SM_OFF_OPS;
  uint8_t* raw = smem + SM_OFF_RAW;
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem + SM_OFF_BARS);
  uint64_t* raw_empty = raw_full + SM_RAW_STAGES;
  uint64_t* op_full = raw_empty + SM_RAW_STAGES;
  uint64_t* op_empty = op_full + SM_OP_STAGES;
  uint64_t* acc_full = op_empty + SM_OP_STAGES;           // [2]
  uint64_t* acc_empty = acc_full + 2;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // ---- one-time setup (overlaps the previous kernel's tail): weights into swizzled smem, barriers, TMEM
  for (int i = threadIdx.x; i < 2 * 128 * 8; i += SM_THREADS) {
    const int plane = i >> 10, row = (i >> 3) & 127, c = i & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t(plane) * 128 + row) * 64 + c * 8));
    *reinterpret_cast<uint4*>((plane ? w_lo : w_hi) + row * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&q.map);
    for (int s = 0; s < SM_RAW_STAGES; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], SM_BUILD_WARPS);
    }
    for (int s = 0; s < SM_OP_STAGES; ++s) {
      mbar_init(&op_full[s], SM_BUILD_WARPS);
      mbar_init(&op_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], ST_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA issuer: lanes 0..3 fetch one block each (WIDE: lane 0 all four)
    int rs = 0;
    uint32_t rphase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      mbar_wait(&raw_empty[rs], rphase ^ 1u, p.err_flag, 100 + rs);
      if constexpr (WIDE) {
        if (lane == 0) {
          if (q.debug & 8) {
            mbar_arrive(&raw_full[rs]);
          } else {
            const unsigned ug = unsigned(tile * ST_BLOCKS);
            const int f = int(__umul64hi((unsigned long long)ug, p.in.inv_bpf));
            const int gb = int(ug) - f * p.in.blocks_per_frame;
            const int by = int(__umul64hi((unsigned long long)unsigned(gb), p.in.inv_bx));
            const int bx = gb - by * p.in.blocks_x;
            mbar_arrive_expect_tx(&raw_full[rs], uint32_t(SM_RAW_BYTES));      // bytes past the frame edge arrive as zeros and count
            tma_load_3d(raw + rs * SM_RAW_BYTES, &q.map, &raw_full[rs], bx * 16, by * 16, f);
          }
        }
        __syncwarp();
        if (++rs == SM_RAW_STAGES) {
          rs = 0;
          rphase ^= 1u;
        }
        continue;
      }
      const int r = tile * ST_BLOCKS + lane;
      const bool valid = lane < ST_BLOCKS && r < n;
      const unsigned n_valid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
      if (lane == 0) {
        if (q.debug & 8) mbar_arrive(&raw_full[rs]);
        else mbar_arrive_expect_tx(&raw_full[rs], 512u * n_valid);
      }
      __syncwarp();
      if (valid && !(q.debug & 8)) {
        const int g = p.idx ? __ldg(p.idx + r) : r;
        const unsigned ug = unsigned(g);
        const int f = p.in.blocks_per_frame == 1 ? int(ug) : int(__umul64hi((unsigned long long)ug, p.in.inv_bpf));
        const int gb = g - f * p.in.blocks_per_frame;
        const int by = p.in.blocks_x == 1 ? gb : int(__umul64hi((unsigned long long)unsigned(gb), p.in.inv_bx));
        const int bx = gb - by * p.in.blocks_x;
        tma_load_3d(raw + rs * SM_RAW_BYTES + lane * 512, &q.map, &raw_full[rs], bx * 16, by * 16, f);
      }
      __syncwarp();
      if (++rs == SM_RAW_STAGES) {
        rs = 0;
        rphase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one lane issues)
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_f16(ST_N);
    const uint32_t a_hi = umma_desc_lo_sw128(smem_u32(w_hi)), a_lo = umma_desc_lo_sw128(smem_u32(w_lo));
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1u, p.err_flag, 600 + acc);
      mbar_wait(&op_full[stage], phase, p.err_flag, 700 + stage);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t d_tmem = tmem_base + uint32_t(acc * ST_N);
        const uint32_t b_lo = umma_desc_lo_sw128(smem_u32(ops + stage * ST_P_BYTES));
        if (!(q.debug & 1)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss_lo(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, 1u);
        }
        umma_commit(&op_empty[stage]);
        umma_commit(&acc_full[acc]);
      }
      __syncwarp();
      if (++stage == SM_OP_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp < 2 + SM_BUILD_WARPS) {
    // ------------------------------------------------------------ builders: raw tile -> im2col operand tile
    // A warp covers 8 conv columns x 4 kernel rows: the 16-byte chunks it stores (chunk index ky ^ px of eight consecutive
    // operand rows) and its 32 word loads are free of bank conflicts - per-block tiles (32-byte rows): bank = 8 ky + px +
    // const with four CONSECUTIVE kernel rows; WIDE tiles (128-byte rows, SWIZZLE_128B): the chunk index is XORed with
    // y & 7, and four kernel rows TWO apart land in four different chunk pairs.
    const int tid = threadIdx.x - 64;                // 0..255
    const int wv = tid >> 5;
    const int c = WIDE ? 2 * (lane >> 3) + (wv & 1) : (lane >> 3) + 4 * (wv & 1);   // K chunk = kernel row ky (7: zero weights)
    const int px = lane & 7;                         // conv column of every row this thread writes (eight consecutive lanes = eight
                                                     // columns of one kernel row: a 128-bit store phase hits eight different chunks)
    const int pyb = wv >> 1;                         // conv rows pyb and pyb + 4
    // the chunk's four words are samples 2 px - 4 + 2 m, + 1 of the block row (m = 0..3): word px - 2 + m of the raw row.
    // Source of (block b, half h, word m):  per-block tiles  src + 512 b + off[h][m];
    //                                       WIDE tiles        src + off[h][m] + ((32 b) ^ xsw[h][m])   (swizzled chunk index)
    uint32_t off[2][4], xsw[2][4], msk[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int y = 2 * (pyb + 4 * h) + c - 3;
      const int yc = min(max(y, 0), 15);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int wi = px - 2 + m;
        const int wc = min(max(wi, 0), 7);
        msk[h][m] = (wi >= 0 && wi < 8 && c < 7 && y >= 0 && y < 16) ? 0xFFFFFFFFu : 0u;
        off[h][m] = WIDE ? uint32_t(yc * 128 + (wc & 3) * 4) : uint32_t(yc * 32 + wc * 4);
        xsw[h][m] = uint32_t(((wc >> 2) ^ (yc & 7)) << 4);
      }
    }
    const uint32_t dst_c = uint32_t((c ^ px) << 4);  // swizzled 16-byte chunk inside the 128-byte operand row (row & 7 == px)
    uint32_t bad = 0u;
    int rs = 0, os = 0;
    uint32_t rphase = 0, ophase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      mbar_wait(&raw_full[rs], rphase, p.err_flag, 200 + rs);
      mbar_wait(&op_empty[os], ophase ^ 1u, p.err_flag, 300 + os);
      const uint32_t src = smem_u32(raw + rs * SM_RAW_BYTES);
      uint8_t* dst = ops + os * ST_P_BYTES + (pyb * 8 + px) * 128 + dst_c;
      uint32_t seen = 0u;                             // OR of every word this thread copied
      if (!(q.debug & 4)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int b = i >> 1, h = i & 1;
          uint4 v;
          v.x = lds_u32(WIDE ? src + off[h][0] + (uint32_t(b * 32) ^ xsw[h][0]) : src + uint32_t(b * 512) + off[h][0]) & msk[h][0];
          v.y = lds_u32(WIDE ? src + off[h][1] + (uint32_t(b * 32) ^ xsw[h][1]) : src + uint32_t(b * 512) + off[h][1]) & msk[h][1];
          v.z = lds_u32(WIDE ? src + off[h][2] + (uint32_t(b * 32) ^ xsw[h][2]) : src + uint32_t(b * 512) + off[h][2]) & msk[h][2];
          v.w = lds_u32(WIDE ? src + off[h][3] + (uint32_t(b * 32) ^ xsw[h][3]) : src + uint32_t(b * 512) + off[h][3]) & msk[h][3];
          seen |= v.x | v.y;
          seen |= v.z | v.w;
          *reinterpret_cast<uint4*>(dst + (b * 64 + h * 32) * 128) = v;
        }
        // Rare, out of line: a sample >= 2048 (outside the range in which the bit pattern is linear - and outside the 10-bit
        // format), or the last tile of a list that is not a multiple of four blocks (its missing blocks were not loaded).
        const int n_blk = n - tile * ST_BLOCKS;
        if ((seen & 0xF800F800u) || n_blk < ST_BLOCKS) bad |= stem_tma_fix_tile(dst, n_blk);
      }
      fence_proxy_async_smem();                       // generic-proxy stores -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&op_full[os]);
        mbar_arrive(&raw_empty[rs]);                  // every load of this warp has been consumed by the stores above
      }
      if (++rs == SM_RAW_STAGES) {
        rs = 0;
        rphase ^= 1u;
      }
      if (++os == SM_OP_STAGES) {
        os = 0;
        ophase ^= 1u;
      }
    }
    if (bad && p.range_flag) *p.range_flag = 1;
  } else {
    // ------------------------------------------------------------ epilogue: bias, ReLU, max-pool, hi/lo stores
    const int quad = warp & 3;
    const int ch = (quad & 1) * 32 + lane;            // accumulator row -> channel (rows 64..127 repeat 0..63)
    const int blk0 = (quad >> 1) * 2;                 // rows 0..63 take blocks 0,1 of the tile, rows 64..127 blocks 2,3
    const int bi = (warp - (2 + SM_BUILD_WARPS)) >> 2;   // which block of that pair this warp handles
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
        if (r < n && !(q.debug & 2)) {
          // max-pool the raw accumulators first (relu(s * a + b) is monotonic in a), separable 3x3 window
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
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace av1p
