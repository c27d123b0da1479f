// Kernels of the generic block-size path (blocks of 8, 32 or 64 luma samples; the 16x16 path has its own
// specialised stem / squeeze-excite / attention kernels).  The reference's networks accept any block size
// (adaptive pooling, pesquisa_v6/v6_pipeline/models.py:64-126) and its dataset tools cut 8 / 16 / 32 / 64
// (pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py:32, pesquisa_v6/scripts/001_prepare_v6_dataset.py:198).
// For a b x b block the backbone's maps are  conv1: b/2, layer1: g1 = b/4, layer2..4: ceil-halved down to 1.
// Every conv / linear layer still runs on the tcgen05 block-Toeplitz FC kernel (fc_tcgen05.cuh); the three
// kernels here cover what is not a matrix product:
//   stem_generic_kernel   gather (planar frame + index list, or float blocks) -> /1023 -> conv1 7x7 s2 p3 + folded
//                         BN -> ReLU -> maxpool 3x3 s2 p1 -> fp16 hi/lo rows [g1*g1 positions][64 channels] (fp32 math)
//   se_generic_kernel     squeeze-excite (models.py:24-43) for any (channels, positions)
//   sam_pool_kernel       SpatialAttention (models.py:46-61: 7x7 conv over the [mean_c, max_c] map) + global average
//                         pool (models.py:122-124) -> 512 features per block, fp16 hi/lo
#pragma once
#include <cuda.h>
#include "aux_kernels.cuh"
#include "stem_tc.cuh"

namespace av1p {

// ------------------------------------------------------------------------------------------------
// One CTA per block (grid-stride).  The block plus a 3-sample zero halo sits in shared memory as fp32 (already divided
// by 1023 with IEEE division, data_hub.py:70-77); thread t owns channel t % 64 - its 49 folded weights live in
// registers - and pooled positions t / 64, t / 64 + 4, ...  For a pooled position it evaluates the (up to) nine conv
// positions of its 3x3 / stride-2 window and keeps the maximum (max-pool pads with -inf: windows are clipped).
constexpr int SG_THREADS = 256;
struct StemGenParams {
  StemInput in;             // kind 0: frames (block geometry for block size `bs`), kind 1: float blocks [n][bs*bs]
  const int* idx;
  const int* n_dev;
  int n;
  int bs;                   // block size: 8, 32 or 64 (16 works too and is used by the parity tests of this kernel)
  const float* w;           // [64][49] folded conv1 weights
  const float* b;           // [64] folded bias
  __half* out;              // [rows][g1*g1*64] tiled activation layout
  __half* out_lo;
  int out_kb;               // 64-column blocks per row of the output buffer (>= g1*g1)
};

__global__ void __launch_bounds__(SG_THREADS) stem_generic_kernel(const StemGenParams p) {
  extern __shared__ float sg_pix[];                 // [(bs + 6)][(bs + 6) + 1]
  const int n = p.n_dev ? *p.n_dev : p.n;
  const int bs = p.bs, tw = bs + 6, pitch = tw + 1;
  const int conv_n = bs >> 1;                       // conv1 output side
  const int g1 = bs >> 2;                           // pooled side
  const int npos = g1 * g1;
  const int ch = threadIdx.x & 63;
  float w[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) w[i] = __ldg(p.w + ch * 49 + i);
  const float bias = __ldg(p.b + ch);
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    const int g = p.idx ? __ldg(p.idx + r) : r;
    int y0 = 0, x0 = 0;
    const uint16_t* frame = nullptr;
    if (p.in.kind == 0) {
      const int f = g / p.in.blocks_per_frame, gb = g - f * p.in.blocks_per_frame;
      const int by = gb / p.in.blocks_x;
      y0 = by * bs;
      x0 = (gb - by * p.in.blocks_x) * bs;
      frame = p.in.frames + size_t(f) * p.in.frame_stride;
    }
    __syncthreads();                                // previous block's readers are done
    for (int i = threadIdx.x; i < tw * tw; i += SG_THREADS) {
      const int ty = i / tw, tx = i - ty * tw;
      const int y = ty - 3, x = tx - 3;
      float v = 0.f;
      if (y >= 0 && y < bs && x >= 0 && x < bs) {
        if (p.in.kind == 0) {
          const int fy = y0 + y, fx = x0 + x;
          if (fy < p.in.height && fx < p.in.width)     // zero padding bottom / right (005:380-383)
            v = __fdiv_rn(float(__ldg(frame + size_t(fy) * p.in.pitch + fx)), 1023.0f);
        } else {
          v = __ldg(p.in.images + size_t(g) * bs * bs + y * bs + x);
        }
      }
      sg_pix[ty * pitch + tx] = v;
    }
    __syncthreads();
    for (int q = threadIdx.x >> 6; q < npos; q += SG_THREADS / 64) {
      const int qy = q / g1, qx = q - qy * g1;
      float best = -INFINITY;
      for (int dy = -1; dy <= 1; ++dy) {
        const int cy = 2 * qy + dy;
        if (cy < 0 || cy >= conv_n) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int cx = 2 * qx + dx;
          if (cx < 0 || cx >= conv_n) continue;
          // conv position (cy, cx): input window rows 2cy-3 .. 2cy+3 = tile rows 2cy .. 2cy+6
          const float* px = sg_pix + (2 * cy) * pitch + 2 * cx;
          float acc = 0.f;
#pragma unroll
          for (int ky = 0; ky < 7; ++ky)
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) acc = fmaf(px[ky * pitch + kx], w[ky * 7 + kx], acc);
          best = fmaxf(best, acc);
        }
      }
      const float m = fmaxf(best + bias, 0.f);         // relu(max(conv) + b) = max(relu(conv + b)): both are monotonic
      const __half h = __float2half_rn(m);
      const size_t o = act_off(r, q * 64 + ch, p.out_kb);
      p.out[o] = h;
      if (p.out_lo) p.out_lo[o] = __float2half_rn(m - __half2float(h));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Squeeze-excite for any shape: one CTA per block row of [npos][C] values (hi + optional lo planes).
//   s = sigmoid(W2 . relu(W1 . mean_positions(x)));  out[pos][c] = x[pos][c] * s[c]         (models.py:39-43)
// w = fp32 [2][C/16][C]: W1 and W2 transposed (same packing as OP_SE).  C in {64, 128, 256, 512}.
constexpr int SEG_THREADS = 256;
__global__ void __launch_bounds__(SEG_THREADS) se_generic_kernel(const __half* __restrict__ x, const __half* __restrict__ x_lo,
                                                                 __half* __restrict__ out, __half* __restrict__ out_lo,
                                                                 const int* n_dev, int n, int C, int npos, int kb,
                                                                 const float* __restrict__ w) {
  __shared__ float s_part[512];
  __shared__ float s_mean[512];
  __shared__ float s_hid[32];
  __shared__ float s_gate[512];
  const int rows = n_dev ? *n_dev : n;
  const int H = C >> 4;
  const int total = npos * C;       // kb = 64-column blocks per row of the (equally wide) source and destination buffers
  // channel sums in a FIXED order (bitwise run-to-run determinism): `per` threads share a channel, each sums every per-th
  // position into its own slot, the slots are added in index order
  const int per = (C < SEG_THREADS) ? SEG_THREADS / C : 1;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    __syncthreads();
    for (int cbase = 0; cbase < C; cbase += SEG_THREADS) {
      const int c = cbase + (C < SEG_THREADS ? threadIdx.x % C : threadIdx.x);
      const int slot = (C < SEG_THREADS) ? threadIdx.x / C : 0;
      float acc = 0.f;
      for (int pos = slot; pos < npos; pos += per) {
        const size_t o = act_off(r, pos * C + c, kb);
        acc += __half2float(x[o]) + (x_lo ? __half2float(x_lo[o]) : 0.f);
      }
      s_part[slot * C + c] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += SEG_THREADS) {
      float acc = 0.f;
      for (int j = 0; j < per; ++j) acc += s_part[j * C + c];
      s_mean[c] = acc * (1.0f / float(npos));
    }
    __syncthreads();
    // hidden units: warp j computes unit j, j + 8, ...
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < H; j += SEG_THREADS / 32) {
      float acc = 0.f;
      for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(w + j * C + c), s_mean[c], acc);
      acc = warp_sum(acc);
      if (lane == 0) s_hid[j] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += SEG_THREADS) {
      float acc = 0.f;
      for (int j = 0; j < H; ++j) acc = fmaf(__ldg(w + (H + j) * C + c), s_hid[j], acc);
      s_gate[c] = 1.0f / (1.0f + expf(-acc));
    }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += SEG_THREADS) {
      const size_t o = act_off(r, e, kb);
      const float v = (__half2float(x[o]) + (x_lo ? __half2float(x_lo[o]) : 0.f)) * s_gate[e % C];
      const __half h = __float2half_rn(v);
      out[o] = h;
      if (out_lo) out_lo[o] = __float2half_rn(v - __half2float(h));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Spatial attention + global average pool on a g x g map of 512 channels (g = 1 or 2: block sizes up to 64):
//   a[p] = sigmoid( sum_q  k_avg[q - p] * mean_c x[q] + k_max[q - p] * max_c x[q] )      7x7 kernel, zero padding 3
//   feat[c] = mean_p ( x[p][c] * a[p] )                                                   (models.py:56-61, 122-124)
// One CTA (128 threads, 4 channels each) per block row.  kern = fp32 [2][7][7].
constexpr int SP_THREADS = 128;
__global__ void __launch_bounds__(SP_THREADS) sam_pool_kernel(const __half* __restrict__ x, const __half* __restrict__ x_lo,
                                                              __half* __restrict__ out, __half* __restrict__ out_lo,
                                                              const int* n_dev, int n, int g, int kb_in, int kb_out,
                                                              const float* __restrict__ kern) {
  __shared__ float s_sum[4][SP_THREADS / 32];
  __shared__ float s_max[4][SP_THREADS / 32];
  __shared__ float s_att[4];
  const int rows = n_dev ? *n_dev : n;
  const int npos = g * g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    float v[4][4];
    __syncthreads();
#pragma unroll
    for (int pq = 0; pq < 4; ++pq) {
      if (pq >= npos) continue;
      float s = 0.f, m = -INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t o = act_off(r, pq * 512 + threadIdx.x * 4 + i, kb_in);
        const float f = __half2float(x[o]) + (x_lo ? __half2float(x_lo[o]) : 0.f);
        v[pq][i] = f;
        s += f;
        m = fmaxf(m, f);
      }
      s = warp_sum(s);
      m = warp_max(m);
      if (lane == 0) {
        s_sum[pq][warp] = s;
        s_max[pq][warp] = m;
      }
    }
    __syncthreads();
    if (threadIdx.x < npos) {
      const int py = threadIdx.x / g, px = threadIdx.x - py * g;
      float a = 0.f;
      for (int q = 0; q < npos; ++q) {
        const int qy = q / g, qx = q - qy * g;
        const int ky = qy - py + 3, kx = qx - px + 3;          // tap that connects input q to output p
        if (ky < 0 || ky > 6 || kx < 0 || kx > 6) continue;
        float s = 0.f, m = -INFINITY;
        for (int wv = 0; wv < SP_THREADS / 32; ++wv) {
          s += s_sum[q][wv];
          m = fmaxf(m, s_max[q][wv]);
        }
        a = fmaf(__ldg(kern + ky * 7 + kx), s * (1.0f / 512.0f), a);
        a = fmaf(__ldg(kern + 49 + ky * 7 + kx), m, a);
      }
      s_att[threadIdx.x] = 1.0f / (1.0f + expf(-a));
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f = 0.f;
#pragma unroll
      for (int pq = 0; pq < 4; ++pq)
        if (pq < npos) f = fmaf(v[pq][i], s_att[pq], f);
      f *= 1.0f / float(npos);
      const __half h = __float2half_rn(f);
      const size_t o = act_off(r, threadIdx.x * 4 + i, kb_out);
      out[o] = h;
      if (out_lo) out_lo[o] = __float2half_rn(f - __half2float(h));
    }
  }
}

}  // namespace av1p
