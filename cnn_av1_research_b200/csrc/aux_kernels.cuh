// Small CUDA-core kernels around the tensor-core layers: spatial-attention gate, FGVC cosine
// tail, standalone block extraction / normalisation, and the inter-stage routing (threshold /
// argmax + stable stream compaction + label scatter).
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"

namespace av1p {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------
// SpatialAttention at 1x1 spatial (reference models.py:56-61): only the centre tap of the 7x7
// kernel touches data, so the module reduces to one scalar per block,
//   g = sigmoid(w_avg * mean_c(x) + w_max * max_c(x)),
// which the next linear layer applies as a row scale.  One warp per row of 512 fp16.
// (`off` = act_off(row, col, kb): 16 consecutive columns never straddle a 64-column block)
__device__ __forceinline__ void load_row16(const __half* __restrict__ hi, const __half* __restrict__ lo, size_t off,
                                           float (&x)[16]) {
  // 16 consecutive values of a row as fp32 (hi + lo in split precision)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(hi + off) + j);
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      x[j * 8 + 2 * i] = f.x;
      x[j * 8 + 2 * i + 1] = f.y;
    }
  }
  if (lo) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(lo + off) + j);
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        x[j * 8 + 2 * i] += f.x;
        x[j * 8 + 2 * i + 1] += f.y;
      }
    }
  }
}

// Spatial attention from the partial statistics the producing FC layer left behind (FcParams::sam_part): per row the
// sum and the maximum over `slots` partials, then the same scalar as sam_gate_kernel.
__global__ void __launch_bounds__(256) sam_finish_kernel(const float* __restrict__ part, int slots, const int* n_dev, int n,
                                                         float w_avg, float w_max, float* __restrict__ row_scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int rows = n_dev ? *n_dev : n;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const float2* q = reinterpret_cast<const float2*>(part) + size_t(r) * slots;
    float s = 0.f, m = -INFINITY;
    for (int i = 0; i < slots; ++i) {
      const float2 v = q[i];
      s += v.x;
      m = fmaxf(m, v.y);
    }
    row_scale[r] = 1.0f / (1.0f + expf(-(w_avg * (s * (1.0f / 512.0f)) + w_max * m)));
  }
}

__global__ void __launch_bounds__(256) sam_gate_kernel(const __half* __restrict__ x, const __half* __restrict__ x_lo,
                                                       int ld, const int* n_dev, int n, float w_avg, float w_max,
                                                       float* __restrict__ row_scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int rows = n_dev ? *n_dev : n;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_per_grid) {
    float v[16];
    load_row16(x, x_lo, act_off(r, lane * 16, ld >> 6), v);
    float s = 0.f, m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      s += v[i];
      m = fmaxf(m, v[i]);
    }
    s = warp_sum(s);
    m = warp_max(m);
    if (lane == 0) row_scale[r] = 1.0f / (1.0f + expf(-(w_avg * (s * (1.0f / 512.0f)) + w_max * m)));
  }
}

// ------------------------------------------------------------------------------------------
// Squeeze-and-excitation block (reference models.py:24-43) as one memory-bound pass:
//   s = sigmoid(W2 . relu(W1 . mean_positions(x)));   out[pos][c] = x[pos][c] * s[c]
// One warp per block row ([NPOS][C] fp16, hi + optional lo); the two tiny weight matrices sit in
// shared memory as fp32 ([H][C] and W2 transposed to [H][C], H = C/16).  Lane l owns the 8 channels of
// channel group l % (C/8) at NPOS*C/256 positions, so every global access is a 16-byte vector and a
// warp touches 512 contiguous bytes per instruction.  Used for se1..se3 (C = 64, 128, 256); se4's
// 2 x 64 KB of weights do not fit this scheme and stay on the tensor-core path.
// Sum the H per-lane partial vectors of a warp (reduce-scatter butterfly), apply ReLU and broadcast: hid[j] in every lane.
template <int H>
__device__ __forceinline__ void se_hidden(float (&part)[H], float (&hid)[H], int lane) {
  // after the step with offset o, a lane keeps the units whose index bit matches its lane bit
  constexpr int STEPS = H == 16 ? 4 : (H == 8 ? 3 : (H == 4 ? 2 : 1));
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int o = 16 >> st;
    if (st < STEPS) {
      const int half_w = H >> (st + 1);
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int j = 0; j < H / 2; ++j) {
        if (j < half_w) {
          const float keep = upper ? part[j + half_w] : part[j];
          const float send = upper ? part[j] : part[j + half_w];
          part[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
    } else {
      part[0] += __shfl_xor_sync(0xffffffffu, part[0], o);
    }
  }
  const float mine = fmaxf(part[0], 0.f);
  // unit j lives in the lane whose first STEPS bits (16, 8, ...) spell j, other bits zero
#pragma unroll
  for (int j = 0; j < H; ++j) {
    int src_lane = 0;
#pragma unroll
    for (int st = 0; st < STEPS; ++st)
      if (j & (H >> (st + 1))) src_lane |= 16 >> st;
    hid[j] = __shfl_sync(0xffffffffu, mine, src_lane);
  }
}

template <int C, int NPOS, int R>
__global__ void __launch_bounds__(256) se_kernel(const __half* __restrict__ x, const __half* __restrict__ x_lo,
                                                 __half* __restrict__ out, __half* __restrict__ out_lo,
                                                 const int* n_dev, int n, const float* __restrict__ w /*[2][H][C]*/) {
  // R = block rows a warp processes together: every shared-memory weight word is loaded once per R rows (the C = 256
  // instance was bound by its weight LDS wavefronts, not by HBM).
  constexpr int H = C / 16;
  constexpr int L = C * NPOS;
  constexpr int GROUPS = C / 8;           // channel groups of 8
  constexpr int CPL = L / 256;            // 16-byte chunks per lane
  static_assert(L % 256 == 0 && GROUPS <= 32, "unsupported SE shape");
  // smem copy of [W1 ; W2^T], each row permuted to [half][group][4] so that the lanes of a warp (one channel
  // group each) read consecutive 16-byte words: channel g*8 + h*4 + k  ->  h*(C/2) + g*4 + k
  extern __shared__ __align__(16) float se_w[];
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 2 * H * C; i += blockDim.x) {      // constant weights: overlaps the previous kernel's tail
    const int row = i / C, ch = i - row * C;
    se_w[row * C + ((ch >> 2) & 1) * (C / 2) + (ch >> 3) * 4 + (ch & 3)] = w[i];
  }
  __syncthreads();
  pdl_wait();
  const float* w1 = se_w;
  const float* w2t = se_w + H * C;
  const int wg = (threadIdx.x & 31) % GROUPS * 4;      // this lane's word offset inside a half row
  const int rows = n_dev ? *n_dev : n;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R; r0 < rows; r0 += warps_per_grid * R) {
    float v[R][CPL][8];
    float mean[R][8];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int r = r0 + q;
#pragma unroll
      for (int k = 0; k < 8; ++k) mean[q][k] = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[q][i][k] = 0.f;
        if (r < rows) {
          const size_t off = act_off(r, (lane + 32 * i) * 8, L / 64);
          const uint4 qh = __ldg(reinterpret_cast<const uint4*>(x + off));
          const __half2* h = reinterpret_cast<const __half2*>(&qh);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            v[q][i][2 * k] = f.x;
            v[q][i][2 * k + 1] = f.y;
          }
          if (x_lo) {
            const uint4 ql = __ldg(reinterpret_cast<const uint4*>(x_lo + off));
            const __half2* hl = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __half22float2(hl[k]);
              v[q][i][2 * k] += f.x;
              v[q][i][2 * k + 1] += f.y;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) mean[q][k] += v[q][i][k];
      }
      // finish the spatial mean across the lanes that hold the same channel group
#pragma unroll
      for (int o = GROUPS; o < 32; o <<= 1)
#pragma unroll
        for (int k = 0; k < 8; ++k) mean[q][k] += __shfl_xor_sync(0xffffffffu, mean[q][k], o);
#pragma unroll
      for (int k = 0; k < 8; ++k) mean[q][k] *= (1.0f / NPOS);
    }
    // hidden = relu(W1 . mean): one leader lane per channel group contributes a partial for every hidden unit; per
    // row the H partial vectors are summed with a reduce-scatter butterfly (H + H/2 + ... shuffles instead of 5 per
    // unit) and broadcast back.
    float part[R][H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (lane < GROUPS) {
        a = *reinterpret_cast<const float4*>(w1 + j * C + wg);
        b = *reinterpret_cast<const float4*>(w1 + j * C + C / 2 + wg);
      }
#pragma unroll
      for (int q = 0; q < R; ++q)
        part[q][j] = mean[q][0] * a.x + mean[q][1] * a.y + mean[q][2] * a.z + mean[q][3] * a.w + mean[q][4] * b.x +
                     mean[q][5] * b.y + mean[q][6] * b.z + mean[q][7] * b.w;
    }
    float hid[R][H];
#pragma unroll
    for (int q = 0; q < R; ++q) se_hidden<H>(part[q], hid[q], lane);
    float sc[R][8];
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int k = 0; k < 8; ++k) sc[q][k] = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(w2t + j * C + wg);
      const float4 b = *reinterpret_cast<const float4*>(w2t + j * C + C / 2 + wg);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const float hj = hid[q][j];
        sc[q][0] = fmaf(a.x, hj, sc[q][0]); sc[q][1] = fmaf(a.y, hj, sc[q][1]);
        sc[q][2] = fmaf(a.z, hj, sc[q][2]); sc[q][3] = fmaf(a.w, hj, sc[q][3]);
        sc[q][4] = fmaf(b.x, hj, sc[q][4]); sc[q][5] = fmaf(b.y, hj, sc[q][5]);
        sc[q][6] = fmaf(b.z, hj, sc[q][6]); sc[q][7] = fmaf(b.w, hj, sc[q][7]);
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int r = r0 + q;
      if (r >= rows) continue;
#pragma unroll
      for (int k = 0; k < 8; ++k) sc[q][k] = 1.0f / (1.0f + expf(-sc[q][k]));
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const size_t off = act_off(r, (lane + 32 * i) * 8, L / 64);
        __align__(16) __half2 hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = v[q][i][2 * k] * sc[q][2 * k], b = v[q][i][2 * k + 1] * sc[q][2 * k + 1];
          hi[k] = __floats2half2_rn(a, b);
          const float2 hf = __half22float2(hi[k]);
          lo[k] = __floats2half2_rn(a - hf.x, b - hf.y);
        }
        *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(hi);
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + off) = *reinterpret_cast<const uint4*>(lo);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// FGVC tail (reference scripts/006_train_stage3_ab_fgvc.py:235-241, 290-293):
//   f = h / max(||h||_2, 1e-12);  logits = 20 * f . What^T,  What = row-normalised classifier weight
// (normalised once at pack time).  One warp per row of 512 fp16.
__global__ void __launch_bounds__(256) fgvc_tail_kernel(const __half* __restrict__ h, const __half* __restrict__ h_lo,
                                                        int ld, const int* n_dev, int n,
                                                        const float* __restrict__ what, float scale,
                                                        float* __restrict__ logits, float* __restrict__ features) {
  pdl_launch_dependents();
  pdl_wait();
  const int rows = n_dev ? *n_dev : n;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_per_grid) {
    float x[16];
    load_row16(h, h_lo, act_off(r, lane * 16, ld >> 6), x);
    float ss = 0.f, d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 16; ++i) ss = fmaf(x[i], x[i], ss);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4* w4 = reinterpret_cast<const float4*>(what + c * 512 + lane * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = __ldg(w4 + i);
        d[c] = fmaf(x[4 * i], w.x, d[c]);
        d[c] = fmaf(x[4 * i + 1], w.y, d[c]);
        d[c] = fmaf(x[4 * i + 2], w.z, d[c]);
        d[c] = fmaf(x[4 * i + 3], w.w, d[c]);
      }
    }
    ss = warp_sum(ss);
#pragma unroll
    for (int c = 0; c < 4; ++c) d[c] = warp_sum(d[c]);
    if (features) {
      // FGVCModel.forward(x, return_features=True) (006...fgvc.py:290, 294-296): the L2-normalised features, fp32 [rows][512]
      const float invn = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      float4* f4 = reinterpret_cast<float4*>(features + size_t(r) * 512 + lane * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        f4[i] = make_float4(x[4 * i] * invn, x[4 * i + 1] * invn, x[4 * i + 2] * invn, x[4 * i + 3] * invn);
    }
    if (lane == 0) {
      const float inv = scale / fmaxf(sqrtf(ss), 1e-12f);
      *reinterpret_cast<float4*>(logits + size_t(r) * 4) = make_float4(d[0] * inv, d[1] * inv, d[2] * inv, d[3] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Standalone extraction (reference 005:353-457 + data_hub.py:70-77).  Each thread moves 8
// consecutive luma samples (one 16-byte load) of one block row; blocks are emitted in row-major
// grid order, out-of-frame samples are zero.  OUT = uint16_t: raw tiles; OUT = float: tiles / 1023.
template <typename OUT>
__global__ void __launch_bounds__(256) extract_blocks_kernel(const uint16_t* __restrict__ y0, int width, int height,
                                                             int pitch, int bs, int blocks_x, int blocks_y,
                                                             OUT* __restrict__ out0, int n_frames, long long frame_stride) {
  // n_frames luma planes `frame_stride` samples apart (planar YUV 4:2:0 sequences: 005:166-172), one launch; frame f's
  // tiles follow frame f-1's in `out`
  const int chunks_x = blocks_x * bs / 8;                       // 8-sample chunks per padded row
  const long long per_frame = (long long)chunks_x * blocks_y * bs;
  const long long total = per_frame * n_frames;
  for (long long tt = blockIdx.x * (long long)blockDim.x + threadIdx.x; tt < total;
       tt += (long long)gridDim.x * blockDim.x) {
    const long long f = tt / per_frame, t = tt - f * per_frame;
    const uint16_t* __restrict__ y = y0 + f * frame_stride;
    OUT* __restrict__ out = out0 + f * ((long long)blocks_x * blocks_y * bs * bs);
    int py, x0;                                                 // padded-frame row, first sample of the 8-sample chunk
    if constexpr (sizeof(OUT) == 2) {
      // uint16 output: as many bytes read as written - work items in INPUT order (coalesced row reads; measured
      // 4.59 TB/s for 32 4K frames vs 3.57 TB/s in output order)
      py = int(t / chunks_x);
      x0 = int(t % chunks_x) * 8;
    } else {
      // float output: twice as many bytes written as read - work items in OUTPUT order (tile, row, chunk), so a warp
      // writes whole tiles contiguously (3.97 vs 3.53 TB/s)
      const int cpr = bs / 8;                                   // chunks per tile row
      const long long q = t / cpr;
      const long long blk = q / bs;
      py = int(blk / blocks_x) * bs + int(q % bs);
      x0 = int(blk % blocks_x) * bs + int(t % cpr) * 8;
    }
    uint16_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (py < height) {
      const uint16_t* src = y + size_t(py) * pitch + x0;
      if (x0 + 7 < width && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
        *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(src));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (x0 + j < width) v[j] = __ldg(src + j);
      }
    }
    const int by = py / bs, r = py - by * bs;
    const int bx = x0 / bs, c = x0 - bx * bs;
    OUT* dst = out + ((size_t(by) * blocks_x + bx) * bs + r) * bs + c;
    if constexpr (sizeof(OUT) == 2) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    } else {
      float4 a, b;
      a.x = __fdiv_rn(float(v[0]), 1023.0f); a.y = __fdiv_rn(float(v[1]), 1023.0f);
      a.z = __fdiv_rn(float(v[2]), 1023.0f); a.w = __fdiv_rn(float(v[3]), 1023.0f);
      b.x = __fdiv_rn(float(v[4]), 1023.0f); b.y = __fdiv_rn(float(v[5]), 1023.0f);
      b.z = __fdiv_rn(float(v[6]), 1023.0f); b.w = __fdiv_rn(float(v[7]), 1023.0f);
      reinterpret_cast<float4*>(dst)[0] = a;
      reinterpret_cast<float4*>(dst)[1] = b;
    }
  }
}

// Same operation with TMA staging (used whenever the plane pitch / frame stride are 16-byte multiples): the luma planes
// are a 3-D tensor [frames][H][W] of uint16; a CTA walks [64 rows x 64 samples] boxes (= (64/bs)^2 tiles), each brought
// into shared memory by ONE cp.async.bulk.tensor load in the SWIZZLE_128B layout (out-of-frame samples arrive as zeros:
// the reference's right / bottom zero padding comes for free), EX_STAGES boxes in flight per CTA.  Threads read 16-byte
// chunks conflict-free and write every tile as one contiguous piece of the output.
constexpr int EX_STAGES = 6;
constexpr int EX_THREADS = 256;
constexpr int EX_BOX_BYTES = 64 * 128;
constexpr int EX_SMEM_BYTES = EX_STAGES * EX_BOX_BYTES + 1024;
template <typename OUT>
__global__ void __launch_bounds__(EX_THREADS) extract_blocks_tma_kernel(const __grid_constant__ CUtensorMap map, int bs, int blocks_x,
                                                                        int blocks_y, int n_frames, OUT* __restrict__ out0) {
  extern __shared__ uint8_t ex_smem_raw[];
  const uint32_t base = (smem_u32(ex_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = ex_smem_raw + (base - smem_u32(ex_smem_raw));
  __shared__ uint64_t full[EX_STAGES];
  const int tpb = 64 / bs;                                          // tiles per box side
  const int tiles_x = (blocks_x + tpb - 1) / tpb, tiles_y = (blocks_y + tpb - 1) / tpb;
  const long long total = (long long)n_frames * tiles_y * tiles_x;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map);
    for (int s = 0; s < EX_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](long long tile, int s) {
    const int tx = int(tile % tiles_x);
    const long long q = tile / tiles_x;
    mbar_arrive_expect_tx(&full[s], uint32_t(EX_BOX_BYTES));
    tma_load_3d(smem + s * EX_BOX_BYTES, &map, &full[s], tx * 64, int(q % tiles_y) * 64, int(q / tiles_y));
  };
  if (threadIdx.x == 0)
    for (int s = 0; s < EX_STAGES; ++s) {
      const long long tile = blockIdx.x + (long long)s * gridDim.x;
      if (tile < total) issue(tile, s);
    }
  const int cpr = bs / 8;                                          // 16-byte chunks per tile row
  long long it = 0;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int s = int(it % EX_STAGES);
    mbar_wait(&full[s], uint32_t(it / EX_STAGES) & 1u, nullptr, 0);
    const uint8_t* box = smem + s * EX_BOX_BYTES;
    const int tx = int(tile % tiles_x);
    const long long q = tile / tiles_x;
    const int ty = int(q % tiles_y);
    const long long f = q / tiles_y;
    OUT* __restrict__ out = out0 + f * ((long long)blocks_x * blocks_y * bs * bs);
#pragma unroll 2
    for (int i = threadIdx.x; i < 512; i += EX_THREADS) {
      // output order inside the box: tile b (row-major), row r, chunk c
      const int c = i % cpr, r = (i / cpr) % bs, b = i / (cpr * bs);
      const int bry = b / tpb, bcx = b - bry * tpb;
      const int by = ty * tpb + bry, bx = tx * tpb + bcx;
      const int row = bry * bs + r;                                // box row
      const int chunk16 = bcx * cpr + c;                           // 16-byte chunk of the 128-byte box row
      const uint4 raw = *reinterpret_cast<const uint4*>(box + row * 128 + ((chunk16 ^ (row & 7)) << 4));
      if (bx < blocks_x && by < blocks_y) {
        OUT* dst = out + ((size_t(by) * blocks_x + bx) * bs + r) * bs + c * 8;
        if constexpr (sizeof(OUT) == 2) {
          *reinterpret_cast<uint4*>(dst) = raw;
        } else {
          const uint16_t* v = reinterpret_cast<const uint16_t*>(&raw);
          float4 lo4, hi4;
          lo4.x = __fdiv_rn(float(v[0]), 1023.0f); lo4.y = __fdiv_rn(float(v[1]), 1023.0f);
          lo4.z = __fdiv_rn(float(v[2]), 1023.0f); lo4.w = __fdiv_rn(float(v[3]), 1023.0f);
          hi4.x = __fdiv_rn(float(v[4]), 1023.0f); hi4.y = __fdiv_rn(float(v[5]), 1023.0f);
          hi4.z = __fdiv_rn(float(v[6]), 1023.0f); hi4.w = __fdiv_rn(float(v[7]), 1023.0f);
          reinterpret_cast<float4*>(dst)[0] = lo4;
          reinterpret_cast<float4*>(dst)[1] = hi4;
        }
      }
    }
    __syncthreads();                    // every thread has USED its shared-memory loads: the stage may be refilled
    if (threadIdx.x == 0) {
      const long long next = tile + (long long)EX_STAGES * gridDim.x;
      if (next < total) issue(next, s);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Routing.  Reference scripts/008_run_pipeline_eval_v6.py:76-125.
//
// A routing step classifies n rows into up to two "keep" classes and writes, for each class, the
// ascending list of original block ids (the order torch's nonzero / boolean-mask indexing gives),
// its length (device counter, read by the next stage's kernels - no host sync), and labels.
//   kind 0 (stage 1): row i is block i.  keep0 <=> sigmoid(logit) >= thr.  labels[i] = 0 for all i.
//   kind 1 (stage 2): row i is block src[i].  cls = argmax softmax(logits[i, 0:3]) (first max wins);
//                     cls 0 -> labels = 1 (SPLIT); keep0 <=> cls == 1 (RECT); keep1 <=> cls == 2 (AB).
// Two passes over a tile decomposition: count (the last CTA to finish scans the tile counts), then
// scatter with an in-tile scan.  No spinning on other CTAs anywhere.
constexpr int ROUTE_THREADS = 256;
constexpr int ROUTE_ITEMS = 4;
constexpr int ROUTE_TILE = ROUTE_THREADS * ROUTE_ITEMS;
constexpr int ROUTE_MAX_TILES = 8192;   // capacity 8 Mi rows

struct RouteParams {
  int kind;
  const float* logits;
  const int* src;          // kind 1: block id of each row
  const int* n_dev;
  int n;
  float thr;
  int* tile_counts;        // [2][ROUTE_MAX_TILES] scratch; holds exclusive offsets after pass 1
  unsigned int* ticket;    // zero-initialised once; reset by the last CTA
  int* out_idx0;
  int* out_idx1;
  int* counts;             // [2] device counters
  uint8_t* labels_u8;      // optional
  long long* labels_i64;   // optional
};

// exp(x - max) as torch's CPU softmax evaluates it.  Logits that differ by a few ulps give
// arguments of ~1e-7; there the result must be the correctly rounded 1 - d (so that a 1-ulp gap is
// not collapsed into a tie, or is collapsed exactly when fp32 rounding collapses it) - expf's 2-ulp
// error bound is not enough, 1 - d is exact to half an ulp for d < 2^-12.
__device__ __forceinline__ float softmax_exp(float x, float m) {
  const float d = m - x;
  return d < 2.44140625e-4f ? 1.0f - d : expf(-d);
}

__device__ __forceinline__ int route_class(const RouteParams& p, int i) {
  // returns 0 = not routed, 1 = keep0, 2 = keep1, 3 = SPLIT (stage 2 only)
  if (p.kind == 0) {
    const float x = p.logits[i];
    const float prob = 1.0f / (1.0f + expf(-x));            // torch.sigmoid in fp32
    return prob >= p.thr ? 1 : 0;
  }
  const float a = p.logits[3 * i], b = p.logits[3 * i + 1], c = p.logits[3 * i + 2];
  const float m = fmaxf(a, fmaxf(b, c));
  const float ea = softmax_exp(a, m), eb = softmax_exp(b, m), ec = softmax_exp(c, m);
  const float s = (ea + eb) + ec;
  const float pa = __fdiv_rn(ea, s), pb = __fdiv_rn(eb, s), pc = __fdiv_rn(ec, s);
  int cls = 0;
  float best = pa;
  if (pb > best) { best = pb; cls = 1; }
  if (pc > best) { best = pc; cls = 2; }
  return cls == 0 ? 3 : cls;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot /*[8]*/, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < ROUTE_THREADS / 32; ++w) {
    const int t = warp_tot[w];
    if (w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(ROUTE_THREADS) route_count_kernel(const RouteParams p) {
  __shared__ int warp_tot[ROUTE_THREADS / 32];
  __shared__ bool is_last;
  const int n = p.n_dev ? *p.n_dev : p.n;
  const int tiles = gridDim.x;
  const int i0 = blockIdx.x * ROUTE_TILE + threadIdx.x * ROUTE_ITEMS;
  int c0 = 0, c1 = 0;
#pragma unroll
  for (int j = 0; j < ROUTE_ITEMS; ++j) {
    const int i = i0 + j;
    if (i < n) {
      const int cls = route_class(p, i);
      c0 += cls == 1;
      c1 += cls == 2;
    }
  }
  int t0, t1;
  block_exclusive_scan(c0, warp_tot, t0);
  block_exclusive_scan(c1, warp_tot, t1);
  if (threadIdx.x == 0) {
    p.tile_counts[blockIdx.x] = t0;
    p.tile_counts[ROUTE_MAX_TILES + blockIdx.x] = t1;
    __threadfence();
    const unsigned int done = atomicAdd(p.ticket, 1u);
    is_last = (done == unsigned(tiles - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last CTA: exclusive scan of the tile counts (tiles <= ROUTE_MAX_TILES), one class at a time
  for (int k = 0; k < 2; ++k) {
    int* tc = p.tile_counts + k * ROUTE_MAX_TILES;
    int carry = 0;
    for (int basei = 0; basei < tiles; basei += ROUTE_THREADS) {
      const int i = basei + threadIdx.x;
      const int v = i < tiles ? __ldcg(tc + i) : 0;
      int tot;
      const int ex = block_exclusive_scan(v, warp_tot, tot);
      if (i < tiles) tc[i] = carry + ex;
      carry += tot;
    }
    if (threadIdx.x == 0) p.counts[k] = carry;
  }
  if (threadIdx.x == 0) *p.ticket = 0u;
}

__global__ void __launch_bounds__(ROUTE_THREADS) route_scatter_kernel(const RouteParams p) {
  __shared__ int warp_tot[ROUTE_THREADS / 32];
  const int n = p.n_dev ? *p.n_dev : p.n;
  const int i0 = blockIdx.x * ROUTE_TILE + threadIdx.x * ROUTE_ITEMS;
  if (blockIdx.x * ROUTE_TILE >= n) return;
  int cls[ROUTE_ITEMS];
  int c0 = 0, c1 = 0;
#pragma unroll
  for (int j = 0; j < ROUTE_ITEMS; ++j) {
    const int i = i0 + j;
    cls[j] = i < n ? route_class(p, i) : 0;
    c0 += cls[j] == 1;
    c1 += cls[j] == 2;
  }
  int t;
  int r0 = p.tile_counts[blockIdx.x] + block_exclusive_scan(c0, warp_tot, t);
  int r1 = p.tile_counts[ROUTE_MAX_TILES + blockIdx.x] + block_exclusive_scan(c1, warp_tot, t);
#pragma unroll
  for (int j = 0; j < ROUTE_ITEMS; ++j) {
    const int i = i0 + j;
    if (i >= n) break;
    const int g = p.kind == 0 ? i : p.src[i];
    if (cls[j] == 1) p.out_idx0[r0++] = g;
    if (cls[j] == 2) p.out_idx1[r1++] = g;
    if (p.kind == 0) {
      if (p.labels_u8) p.labels_u8[g] = 0;
      if (p.labels_i64) p.labels_i64[g] = 0;
    } else if (cls[j] == 3) {
      if (p.labels_u8) p.labels_u8[g] = 1;
      if (p.labels_i64) p.labels_i64[g] = 1;
    }
  }
}

// Label scatter: labels[idx[i]] = base + argmax_j logits[i, j] (first max wins), k <= 8 classes.
//   use_softmax = 1: argmax over softmax probabilities exactly as torch evaluates them (stage 3, 008:108-125);
//   use_softmax = 0: argmax over the raw logits (flatten cascade, 008b:213 `stage2_logits.argmax(dim=1)`).
constexpr int FINALIZE_MAX_K = 8;
__global__ void __launch_bounds__(256) finalize_labels_kernel(const float* __restrict__ logits, int k, int base,
                                                              const int* __restrict__ idx, const int* n_dev, int n,
                                                              uint8_t* labels_u8, long long* labels_i64, int use_softmax) {
  const int rows = n_dev ? *n_dev : n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
    float x[FINALIZE_MAX_K];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < FINALIZE_MAX_K; ++j) {
      x[j] = j < k ? logits[size_t(i) * k + j] : -INFINITY;
      m = fmaxf(m, x[j]);
    }
    int cls = 0;
    if (use_softmax) {
      float e[FINALIZE_MAX_K], s = 0.f;
#pragma unroll
      for (int j = 0; j < FINALIZE_MAX_K; ++j) {
        e[j] = j < k ? softmax_exp(x[j], m) : 0.f;
        if (j < k) s += e[j];
      }
      float best = __fdiv_rn(e[0], s);
#pragma unroll
      for (int j = 1; j < FINALIZE_MAX_K; ++j) {
        const float pj = __fdiv_rn(e[j], s);
        if (j < k && pj > best) { best = pj; cls = j; }
      }
    } else {
      float best = x[0];
#pragma unroll
      for (int j = 1; j < FINALIZE_MAX_K; ++j)
        if (j < k && x[j] > best) { best = x[j]; cls = j; }
    }
    const int g = idx[i];
    if (labels_u8) labels_u8[g] = uint8_t(base + cls);
    if (labels_i64) labels_i64[g] = base + cls;
  }
}

// Rows of `k` floats picked by an index list with a device-side count: dst[j] = src[idx[j]], j < *n_dev.  The speculative
// small-batch cascade (av1p.cu) runs every stage on every block and then compacts the logits of the routed blocks with this,
// so that the routing kernels and every cascade output (logits, index lists, counts) are those of the routed path.
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                          const int* n_dev, int n, int k, float* __restrict__ dst) {
  const int rows = n_dev ? *n_dev : n;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < rows * k; t += gridDim.x * blockDim.x) {
    const int j = t / k, c = t - j * k;
    dst[t] = src[size_t(idx[j]) * k + c];
  }
}

// ------------------------------------------------------------------------------------------
// Stage-1 threshold sweep (reference scripts/007_optimize_thresholds.py:24-71): prob = sigmoid(logit) in fp32,
// pred_t = prob >= thr[t], and for every threshold the confusion counts against the binary stage-1 label:
// counts[t][0..3] = {tn, fp, fn, tp}.  One pass over the logits, warp-aggregated integer atomics -> exact.
// Optionally writes the probabilities (what 007 collects on the host).  The reference compares a float32 numpy
// array with np.float64 thresholds (np.arange, 007:153): under NumPy >= 2 that comparison runs in float64, so the
// probability is widened, not the threshold narrowed.
constexpr int SWEEP_MAX_T = 32;
struct SweepParams {
  const float* logits;
  const uint8_t* labels;     // 0 / 1
  int n;
  int n_thr;
  double thr[SWEEP_MAX_T];
  float* probs;              // optional
  unsigned long long* counts;   // [n_thr][4], zeroed by the caller
};
__global__ void __launch_bounds__(256) threshold_sweep_kernel(const SweepParams p) {
  __shared__ unsigned int sm[SWEEP_MAX_T * 4];
  for (int i = threadIdx.x; i < p.n_thr * 4; i += blockDim.x) sm[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    const bool ok = i < p.n;
    float prob = 0.f;
    int lab = 0;
    if (ok) {
      prob = 1.0f / (1.0f + expf(-p.logits[i]));            // torch.sigmoid in fp32
      lab = p.labels[i] != 0;
      if (p.probs) p.probs[i] = prob;
    }
    for (int t = 0; t < p.n_thr; ++t) {
      const int cls = ok ? lab * 2 + (double(prob) >= p.thr[t] ? 1 : 0) : -1;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const unsigned int votes = __ballot_sync(0xffffffffu, cls == c);
        if (lane == 0 && votes) atomicAdd(&sm[t * 4 + c], __popc(votes));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.n_thr * 4; i += blockDim.x)
    if (sm[i]) atomicAdd(&p.counts[i], (unsigned long long)sm[i]);
}

// ------------------------------------------------------------------------------------------------
// Ensemble voting over the logits of several Stage-3-AB models (reference pesquisa_v6/v6_pipeline/ensemble.py):
//   mode 0  hard voting  (:57-79)  : per-model argmax (first maximum), majority class; torch.unique sorts the votes and
//                                    counts.argmax() takes the first maximum, so ties go to the SMALLEST class id;
//                                    confidence = majority count / n_models
//   mode 1  soft voting  (:50-55)  : softmax per model (fp32), mean over models, argmax / max of the mean
//   mode 2  weighted soft (:165-183): sum_m w_m * softmax_m with the caller's normalised weights
// Optional outputs of predict_with_uncertainty (:83-116): mean / unbiased std of the probabilities over the models, the
// share of models whose argmax equals the prediction, and all probabilities.  One thread per block row; k <= 8, M <= 8.
constexpr int ENS_MAX_K = 8;
constexpr int ENS_MAX_M = 8;
struct EnsembleParams {
  const float* logits;      // [M][n][k]
  const float* weights;     // [M] (mode 2)
  int n_models, n, k, mode;
  long long* pred;          // [n]
  float* conf;              // [n] or nullptr
  float* mean_probs;        // [n][k] or nullptr
  float* std_probs;         // [n][k] or nullptr
  float* agreement;         // [n] or nullptr
  float* all_probs;         // [M][n][k] or nullptr
};
__global__ void __launch_bounds__(256) ensemble_vote_kernel(const EnsembleParams p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    float acc[ENS_MAX_K];
    int votes[ENS_MAX_K], am[ENS_MAX_M];
#pragma unroll
    for (int c = 0; c < ENS_MAX_K; ++c) {
      acc[c] = 0.f;
      votes[c] = 0;
    }
    float prob[ENS_MAX_M][ENS_MAX_K];
    for (int m = 0; m < p.n_models; ++m) {
      const float* l = p.logits + (size_t(m) * p.n + i) * p.k;
      float mx = l[0];
      int best = 0;
      for (int c = 1; c < p.k; ++c)
        if (l[c] > mx) {             // strict: the first maximum wins, as torch.argmax
          mx = l[c];
          best = c;
        }
      am[m] = best;
      votes[best] += 1;
      float e[ENS_MAX_K], sum = 0.f;
      for (int c = 0; c < p.k; ++c) {
        e[c] = expf(l[c] - mx);
        sum += e[c];
      }
      const float w = p.mode == 2 ? p.weights[m] : 1.0f;
      for (int c = 0; c < p.k; ++c) {
        const float q = e[c] / sum;
        prob[m][c] = q;
        acc[c] += p.mode == 2 ? q * w : q;
        if (p.all_probs) p.all_probs[(size_t(m) * p.n + i) * p.k + c] = q;
      }
    }
    if (p.mode != 2)
      for (int c = 0; c < p.k; ++c) acc[c] = acc[c] / float(p.n_models);      // torch .mean(dim=0)
    int pred = 0;
    float conf = 0.f;
    if (p.mode == 0) {
      int top = votes[0];
      for (int c = 1; c < p.k; ++c)
        if (votes[c] > top) {
          top = votes[c];
          pred = c;
        }
      conf = float(top) / float(p.n_models);
    } else {
      conf = acc[0];
      for (int c = 1; c < p.k; ++c)
        if (acc[c] > conf) {
          conf = acc[c];
          pred = c;
        }
    }
    p.pred[i] = pred;
    if (p.conf) p.conf[i] = conf;
    if (p.mean_probs)
      for (int c = 0; c < p.k; ++c) p.mean_probs[size_t(i) * p.k + c] = acc[c];
    if (p.std_probs) {
      for (int c = 0; c < p.k; ++c) {
        float s2 = 0.f;
        for (int m = 0; m < p.n_models; ++m) {
          const float d = prob[m][c] - acc[c];
          s2 += d * d;
        }
        p.std_probs[size_t(i) * p.k + c] = p.n_models > 1 ? sqrtf(s2 / float(p.n_models - 1)) : nanf("");
      }
    }
    if (p.agreement) {
      int same = 0;
      for (int m = 0; m < p.n_models; ++m) same += am[m] == pred;
      p.agreement[i] = float(same) / float(p.n_models);
    }
  }
}

}  // namespace av1p
