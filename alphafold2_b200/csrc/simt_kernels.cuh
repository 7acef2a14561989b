// Bandwidth-bound helper kernels (no tensor-core work): row LayerNorm with optional fused pair-bias
// projection, channel-major -> token-major LayerNorm*gate, pair-mask counts, rotary embedding.
// All are coalesced / vectorised and sized as grid-stride loops over 148 x k CTAs.
#pragma once
#include "common.cuh"

namespace af2 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// y = LayerNorm(x) * gamma + beta  -> bf16 [T, d]          (nn.LayerNorm, eps inside sqrt)
// optionally  bias[h][(t / n_inner) * pitch + t % n_inner] = <x_raw[t, :], Wb[h, :]>   (bf16)
//   (edges_to_attn_bias of alphafold2.py:214-217,245-247 acts on the RAW, un-normalised pair tensor)
// One warp per row; the row lives in registers (d <= 1024, d % 4 == 0).
// ------------------------------------------------------------------------------------------------
struct LnParams {
  const float* x;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;        // may be nullptr (bias only)
  long long T;
  int d;
  float eps;
  const float* wb;         // [H, d] or nullptr
  __nv_bfloat16* bias_out; // [H][bias_hs]
  int heads;
  long long bias_hs;       // elements between heads
  int n_inner, pitch;      // token t -> (t / n_inner) * pitch + t % n_inner
};

template <int MAXC>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long warp_global = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * warps_per_block;
  const int nchunk = p.d >> 2;   // float4 chunks per row
  for (long long t = warp_global; t < p.T; t += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(p.x + t * p.d);
    float4 v[MAXC];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int idx = lane + 32 * c;
      if (idx < nchunk) {
        v[c] = __ldg(xr + idx);
        sum += v[c].x + v[c].y + v[c].z + v[c].w;
      } else {
        v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = warp_sum(sum) / p.d;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int idx = lane + 32 * c;
      if (idx < nchunk) {
        const float a = v[c].x - mean, b = v[c].y - mean, cc = v[c].z - mean, dd = v[c].w - mean;
        sq += a * a + b * b + cc * cc + dd * dd;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / p.d + p.eps);
    if (p.y) {
      uint2* yr = reinterpret_cast<uint2*>(p.y + t * p.d);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int idx = lane + 32 * c;
        if (idx < nchunk) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + idx);
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta) + idx);
          const float o0 = (v[c].x - mean) * rstd * g.x + b.x;
          const float o1 = (v[c].y - mean) * rstd * g.y + b.y;
          const float o2 = (v[c].z - mean) * rstd * g.z + b.z;
          const float o3 = (v[c].w - mean) * rstd * g.w + b.w;
          yr[idx] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
      }
    }
    if (p.wb) {
      const long long off = (t / p.n_inner) * p.pitch + (t % p.n_inner);
      // 8 heads at a time: per-lane partial dot products, then a recursive-halving butterfly (9 shuffles for 8
      // sums instead of 40): after the xor-16/8/4 steps lane l holds head ((l>>2)&7)'s partial, xor-2/1 finish it.
      for (int h0 = 0; h0 < p.heads; h0 += 8) {
        float acc[8];
#pragma unroll
        for (int hh = 0; hh < 8; ++hh) {
          acc[hh] = 0.f;
          if (h0 + hh < p.heads) {
            const float4* wr = reinterpret_cast<const float4*>(p.wb + static_cast<long long>(h0 + hh) * p.d);
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
              const int idx = lane + 32 * c;
              if (idx < nchunk) {
                const float4 w = __ldg(wr + idx);
                acc[hh] += v[c].x * w.x + v[c].y * w.y + v[c].z * w.z + v[c].w * w.w;
              }
            }
          }
        }
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        float w4[4], w2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float mine = b4 ? acc[4 + i] : acc[i];
          const float other = b4 ? acc[i] : acc[4 + i];
          w4[i] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float mine = b3 ? w4[2 + i] : w4[i];
          const float other = b3 ? w4[i] : w4[2 + i];
          w2[i] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
        }
        float w1 = (b2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? w2[0] : w2[1], 4);
        w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
        w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
        const int hsel = h0 + ((lane >> 2) & 7);
        if ((lane & 3) == 0 && hsel < p.heads) p.bias_out[hsel * p.bias_hs + off] = __float2bfloat16(w1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Channel-major fp32 contraction output  ->  token-major bf16 operand of the following Linear.
//   src[c][row * pitch + j]   (c < d channels, row < rows, j < n)    token = row * n + j
//   mode 0 (triangle multiply tail, alphafold2.py:315-316):  y = LayerNorm_c(src) * gamma + beta, * gate[token, c]
//   mode 1 (outer mean tail,      alphafold2.py:347-349):    y = src * scale[token]
// Block = 256 threads handles 32 consecutive j of one row: coalesced reads along j, smem transpose,
// coalesced bf16 writes along c.
// ------------------------------------------------------------------------------------------------
struct ChanLnParams {
  const float* src;
  long long chan_stride;
  int pitch, rows, n, d;
  int mode;
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* gate;   // [tokens, d]
  const float* scale;          // [tokens]
  float scale_const;           // used when scale == nullptr (mode 1)
  float eps;
  __nv_bfloat16* y;            // [tokens, d]
};

__global__ void __launch_bounds__(256) chan_to_token_kernel(const ChanLnParams p) {
  extern __shared__ float tile[];   // [d][33]
  const int jt = blockIdx.x, row = blockIdx.y;
  const int j0 = jt * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool j_ok = (j0 + lane) < p.n;
  const float* src = p.src + static_cast<long long>(row) * p.pitch + j0 + lane;
  for (int c = warp; c < p.d; c += 8)
    tile[c * 33 + lane] = j_ok ? __ldg(src + c * p.chan_stride) : 0.f;
  __syncthreads();
  for (int tk = warp; tk < 32; tk += 8) {
    if (j0 + tk >= p.n) break;
    const long long token = static_cast<long long>(row) * p.n + j0 + tk;
    if (p.mode == 0) {
      float sum = 0.f;
      for (int c = lane; c < p.d; c += 32) sum += tile[c * 33 + tk];
      const float mean = warp_sum(sum) / p.d;
      float sq = 0.f;
      for (int c = lane; c < p.d; c += 32) {
        const float a = tile[c * 33 + tk] - mean;
        sq += a * a;
      }
      const float rstd = rsqrtf(warp_sum(sq) / p.d + p.eps);
      for (int c = lane; c < p.d; c += 32) {
        const float g = __bfloat162float(p.gate[token * p.d + c]);
        const float o = ((tile[c * 33 + tk] - mean) * rstd * __ldg(p.gamma + c) + __ldg(p.beta + c)) * g;
        p.y[token * p.d + c] = __float2bfloat16(o);
      }
    } else {
      const float sc = p.scale ? __ldg(p.scale + token) : p.scale_const;
      for (int c = lane; c < p.d; c += 32) p.y[token * p.d + c] = __float2bfloat16(tile[c * 33 + tk] * sc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pair bias of axial / triangle attention (alphafold2.py:214-217, 245-247):
//   bias[h][(t / n_inner) * pitch + t % n_inner] = <x_raw[t, :], Wb[h, :]>      bf16, H <= 8 heads per pass, d <= 256
// Eight lanes share a token (4 tokens per warp instruction): 128-byte coalesced row segments, w_edge staged in shared
// memory, 3-level shuffle reductions.  Streams x once (HBM bound).
// ------------------------------------------------------------------------------------------------
struct PairBiasParams {
  const float* x;
  long long T;
  int d;
  const float* wb;          // [H, d]
  __nv_bfloat16* bias_out;  // [H][bias_hs]
  int heads;
  long long bias_hs;
  int n_inner, pitch;
};

__global__ void __launch_bounds__(256) pair_bias_kernel(const PairBiasParams p) {
  extern __shared__ float wsm[];                       // [heads][d]
  for (int i = threadIdx.x; i < p.heads * p.d; i += blockDim.x) wsm[i] = __ldg(p.wb + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 7, rg = lane >> 3;
  const int nj = p.d >> 5;                             // float4 chunks per lane
  const long long warp_global = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long t0 = warp_global * 4; t0 < p.T; t0 += nwarps * 4) {
    const long long t = t0 + rg;
    const bool live = t < p.T;
    float4 v[8];
    const float4* xr = reinterpret_cast<const float4*>(p.x + (live ? t : 0) * p.d);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (live && j < nj) ? __ldg(xr + j * 8 + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    float acc[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      acc[h] = 0.f;
      if (h < p.heads) {
        const float4* wr = reinterpret_cast<const float4*>(wsm + h * p.d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nj) {
            const float4 w = wr[j * 8 + sub];
            acc[h] += v[j].x * w.x + v[j].y * w.y + v[j].z * w.z + v[j].w * w.w;
          }
        }
      }
    }
    // reduce over the 8 lanes of the token: after xor-4 / 2 / 1 halving, lane `sub` holds head `sub`
    float r4[4], r2[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool hi = sub & 4;
      const float mine = hi ? acc[4 + k] : acc[k], other = hi ? acc[k] : acc[4 + k];
      r4[k] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const bool hi = sub & 2;
      const float mine = hi ? r4[2 + k] : r4[k], other = hi ? r4[k] : r4[2 + k];
      r2[k] = mine + __shfl_xor_sync(0xffffffffu, other, 2);
    }
    const bool hi1 = sub & 1;
    const float res = (hi1 ? r2[1] : r2[0]) + __shfl_xor_sync(0xffffffffu, hi1 ? r2[0] : r2[1], 1);
    // lane sub now holds head index: bit2 = sub&4 -> +4, bit1 -> +2, bit0 -> +1  == sub
    if (live && sub < p.heads) {
      const long long off = (t / p.n_inner) * p.pitch + (t % p.n_inner);
      p.bias_out[sub * p.bias_hs + off] = __float2bfloat16(res);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Channel-major fp32 contraction output -> token-major bf16 operand, tile version (replaces chan_to_token_kernel when the
// token grid is dense, pitch == n):  a CTA (512 threads) owns 64 consecutive tokens x all d channels:
//   phase 1: 256-byte coalesced row segments src[c][t0 .. t0+63] -> shared tile [d][64] (16 independent 16-byte loads per thread)
//   phase 2: thread = (token, 32-channel slice): LayerNorm over channels (mode 0) or scale (mode 1), gate, 64-byte bf16 stores
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) chan_to_token_tile_kernel(const ChanLnParams p, long long T) {
  extern __shared__ float tile[];                      // [d][64] followed by [8][64][2] partial moments
  const long long t0 = static_cast<long long>(blockIdx.x) * 64;
  const int tid = threadIdx.x;
  // ---- phase 1 ----
  {
    const int q = tid & 15;                            // float4 index inside the 64-token segment
    const bool ok = (t0 + q * 4) < T;                  // T % 4 == 0 (pitch multiple of 4)
    for (int c = tid >> 4; c < p.d; c += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = __ldg(reinterpret_cast<const float4*>(p.src + static_cast<long long>(c) * p.chan_stride + t0) + q);
      reinterpret_cast<float4*>(tile + c * 64)[q] = v;
    }
  }
  __syncthreads();
  // ---- phase 2 ----
  const int tok = tid & 63, slice = tid >> 6;          // 8 slices of d/8 channels
  const int cpt = p.d >> 3;                            // channels per thread (<= 32)
  const long long token = t0 + tok;
  float* part = tile + p.d * 64;
  float mean = 0.f, rstd = 1.f;
  if (p.mode == 0) {
    float s1 = 0.f;
    for (int i = 0; i < cpt; ++i) s1 += tile[(slice * cpt + i) * 64 + tok];
    part[(slice * 64 + tok) * 2] = s1;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[(k * 64 + tok) * 2];
    mean = tot / p.d;
    float s2 = 0.f;
    for (int i = 0; i < cpt; ++i) {
      const float a = tile[(slice * cpt + i) * 64 + tok] - mean;
      s2 += a * a;
    }
    part[(slice * 64 + tok) * 2 + 1] = s2;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) var += part[(k * 64 + tok) * 2 + 1];
    rstd = rsqrtf(var / p.d + p.eps);
  }
  if (token < T) {
    const int c0 = slice * cpt;
    const float sc = (p.mode == 1) ? (p.scale ? __ldg(p.scale + token) : p.scale_const) : 1.f;
    for (int i = 0; i < cpt; i += 8) {
      float o[8];
      if (p.mode == 0) {
        const uint4 gq = __ldg(reinterpret_cast<const uint4*>(p.gate + token * p.d + c0 + i));
        const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = c0 + i + k;
          const float g = (k & 1) ? bf16hi_to_f32(gw[k >> 1]) : bf16lo_to_f32(gw[k >> 1]);
          o[k] = ((tile[c * 64 + tok] - mean) * rstd * __ldg(p.gamma + c) + __ldg(p.beta + c)) * g;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = tile[(c0 + i + k) * 64 + tok] * sc;
      }
      *reinterpret_cast<uint4*>(p.y + token * p.d + c0 + i) =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// OuterMean normaliser (quirk Q3, alphafold2.py:345-347):
//   scale[b][i][j] = 1 / (S * (sum_s mask[b,s,i] * mask[b,s,j] + eps))      (fp32, like the reference)
// ------------------------------------------------------------------------------------------------
__global__ void outer_scale_kernel(const uint8_t* __restrict__ mask, float* __restrict__ scale, int B, int S, int N,
                                   float eps) {
  const long long total = static_cast<long long>(B) * N * N;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = idx % N;
    const int i = (idx / N) % N;
    const int b = idx / (static_cast<long long>(N) * N);
    const uint8_t* mb = mask + static_cast<long long>(b) * S * N;
    int cnt = 0;
    for (int s = 0; s < S; ++s) cnt += (mb[s * N + i] != 0) & (mb[s * N + j] != 0);
    scale[idx] = 1.0f / (static_cast<float>(S) * (static_cast<float>(cnt) + eps));
  }
}

// same for a band of pair rows [row0, row0 + rows) (sharded outer mean): scale[(i - row0) * N + j]
__global__ void outer_scale_rows_kernel(const uint8_t* __restrict__ mask, float* __restrict__ scale, int row0, int rows,
                                        int S, int N, float eps) {
  const long long total = static_cast<long long>(rows) * N;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = idx % N;
    const int i = row0 + static_cast<int>(idx / N);
    int cnt = 0;
    for (int s = 0; s < S; ++s) cnt += (mask[s * N + i] != 0) & (mask[s * N + j] != 0);
    scale[idx] = 1.0f / (static_cast<float>(S) * (static_cast<float>(cnt) + eps));
  }
}

// bool mask -> float 0/1 row scale
__global__ void mask_to_float_kernel(const uint8_t* __restrict__ mask, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = mask[i] ? 1.0f : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// apply_rotary_pos_emb (rotary.py:9-20): x [b, h, n, dh] fp32; sin/cos [bs, n, rot] (bs = 1 or b).
// Interleaved pairs (x0, x1) -> (x0 c0 - x1 s0, x1 c1 + x0 s1); channels >= rot pass through.
// ------------------------------------------------------------------------------------------------
__global__ void rotary_kernel(const float* __restrict__ x, const float* __restrict__ sn, const float* __restrict__ cs,
                              float* __restrict__ y, int b, int h, int n, int dh, int rot, int sincos_batch) {
  const long long pairs = static_cast<long long>(b) * h * n * (dh / 2);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < pairs;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pr = idx % (dh / 2);
    const long long row = idx / (dh / 2);      // (b, h, n) flattened
    const int ni = row % n;
    const int bi = row / (static_cast<long long>(n) * h);
    const float2 v = reinterpret_cast<const float2*>(x)[idx];
    float2 o = v;
    if (2 * pr + 1 < rot) {
      const long long so = (static_cast<long long>(sincos_batch > 1 ? bi : 0) * n + ni) * rot + 2 * pr;
      const float s0 = sn[so], s1 = sn[so + 1], c0 = cs[so], c1 = cs[so + 1];
      // separate roundings (no FMA contraction) so the result is bit-identical to x*cos + rot(x)*sin
      o.x = __fadd_rn(__fmul_rn(v.x, c0), __fmul_rn(-v.y, s0));
      o.y = __fadd_rn(__fmul_rn(v.y, c1), __fmul_rn(v.x, s1));
    }
    reinterpret_cast<float2*>(y)[idx] = o;
  }
}

}  // namespace af2
