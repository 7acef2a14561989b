// Channel-major fp32 contraction output -> token-major bf16 GEMM operand, persistent TMA-pipelined version.
// Follows the triangle / outer-product contractions (alphafold2.py:340-351 to_out_norm * out_gate, :381-395 mean scaling):
//   mode 0: y[t][c] = (LayerNorm_c(src[c][t]) * gamma + beta) * gate[t][c]
//   mode 1: y[t][c] = src[c][t] * scale[t]
// One CTA per SM walks tiles of 32 tokens x D channels:
//   warp 0      TMA producer: the fp32 tile [D][32] (128-byte rows) and the gate tile (D/64 boxes [32 tok][64 ch], 128B swizzle)
//               arrive through two independent 4-deep rings, so the loads of the next tiles are always in flight (the
//               tile-per-CTA version reached 2.4 TB/s: every CTA ran load -> sync -> compute -> store back to back)
//   warp 1      store warp: TMA-stores a finished tile and hands the gate ring slot back
//   2 groups    of D/32 warps, alternating tiles: thread = (token, 32-channel slice); the slice is read from shared memory
//               conflict-free (lane = token), LayerNorm moments are combined through shared memory, and the result
//               overwrites the gate values it was computed from (same swizzled place), which is the TMA store source
// HBM bytes per token: 4 D (src) + 2 D (gate) + 2 D (y).
#pragma once
#include "common.cuh"

namespace af2 {

struct Chan2TokParams {
  long long T;                 // tokens
  int mode;                    // 0: LayerNorm * gate, 1: scale
  const float* gamma;
  const float* beta;
  const float* scale;          // [tokens] or nullptr (mode 1)
  float scale_const;
  float eps;
};

constexpr int C2T_STAGES = 4;
constexpr int C2T_TOK = 32;

template <int D>
struct Chan2TokSmem {
  static constexpr int SL = D / 32;                        // warps per consumer group
  static constexpr int X_BYTES = D * C2T_TOK * 4;
  static constexpr int G_BYTES = C2T_TOK * D * 2;          // D/64 boxes of 4 KB
  static constexpr int X_OFF = 0;
  static constexpr int G_OFF = C2T_STAGES * X_BYTES;
  static constexpr int AFF_OFF = G_OFF + C2T_STAGES * G_BYTES;     // gamma[D], beta[D]
  static constexpr int PART_OFF = AFF_OFF + 2 * D * 4;            // [2 groups][2 moments][SL][32]
  static constexpr int BAR_OFF = PART_OFF + 2 * 2 * SL * 32 * 4;
  static constexpr int TOTAL = BAR_OFF + 5 * C2T_STAGES * 8;
  static constexpr int THREADS = 64 + 2 * SL * 32;
};

// tmX: fp32 2-D (token [T], channel [D]), box (32, D), no swizzle.  tmG / tmY: bf16 2-D (channel [D], token [T]), box (64, 32), SW128.
template <int D>
__global__ void __launch_bounds__(Chan2TokSmem<D>::THREADS, 1)
chan_to_token_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                         const __grid_constant__ CUtensorMap tmY, const __grid_constant__ Chan2TokParams p) {
  using L = Chan2TokSmem<D>;
  constexpr int SL = L::SL;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* xempty = xfull + C2T_STAGES;       // count SL: every consumer warp has its slice in registers
  uint64_t* gfull = xempty + C2T_STAGES;
  uint64_t* gempty = gfull + C2T_STAGES;       // count 1: the store warp has drained the slot
  uint64_t* ostaged = gempty + C2T_STAGES;     // count SL: every consumer warp has written its outputs into the slot
  float* aff = reinterpret_cast<float*>(smem + L::AFF_OFF);
  float* part = reinterpret_cast<float*>(smem + L::PART_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (p.T + C2T_TOK - 1) / C2T_TOK;
  const long long my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmG); prefetch_tmap(&tmY);
    for (int s = 0; s < C2T_STAGES; ++s) {
      mbar_init(&xfull[s], 1);
      mbar_init(&xempty[s], SL);
      mbar_init(&gfull[s], 1);
      mbar_init(&gempty[s], 1);
      mbar_init(&ostaged[s], SL);
    }
    fence_barrier_init();
  }
  if (p.mode == 0)
    for (int i = threadIdx.x; i < D; i += L::THREADS) { aff[i] = p.gamma[i]; aff[D + i] = p.beta[i]; }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (long long i = 0; i < my_tiles; ++i) {
        const int s = static_cast<int>(i % C2T_STAGES);
        const uint32_t ph = static_cast<uint32_t>((i / C2T_STAGES) & 1);
        const int t0 = static_cast<int>((blockIdx.x + i * gridDim.x) * C2T_TOK);
        mbar_wait(&xempty[s], ph ^ 1);
        mbar_arrive_expect_tx(&xfull[s], L::X_BYTES);
        tma_load_2d(smem + L::X_OFF + s * L::X_BYTES, &tmX, &xfull[s], t0, 0);
        if (p.mode == 0) {
          mbar_wait(&gempty[s], ph ^ 1);
          mbar_arrive_expect_tx(&gfull[s], L::G_BYTES);
#pragma unroll
          for (int q = 0; q < D / 64; ++q) tma_load_2d(smem + L::G_OFF + s * L::G_BYTES + q * 4096, &tmG, &gfull[s], q * 64, t0);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ store warp ===================================
    if (lane == 0) {
      for (long long i = 0; i < my_tiles; ++i) {
        const int s = static_cast<int>(i % C2T_STAGES);
        const uint32_t ph = static_cast<uint32_t>((i / C2T_STAGES) & 1);
        const int t0 = static_cast<int>((blockIdx.x + i * gridDim.x) * C2T_TOK);
        mbar_wait(&ostaged[s], ph);
#pragma unroll
        for (int q = 0; q < D / 64; ++q) tma_store_2d(&tmY, smem + L::G_OFF + s * L::G_BYTES + q * 4096, q * 64, t0);   // tokens >= T are clipped
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&gempty[s]);
      }
    }
  } else {
    // ================================ consumer groups ==============================
    const int grp = (warp - 2) / SL, slice = (warp - 2) % SL;
    const int c0 = slice * 32;
    float* p1 = part + (grp * 2 + 0) * SL * 32;
    float* p2 = part + (grp * 2 + 1) * SL * 32;
    for (long long i = grp; i < my_tiles; i += 2) {
      const int s = static_cast<int>(i % C2T_STAGES);
      const uint32_t ph = static_cast<uint32_t>((i / C2T_STAGES) & 1);
      const long long token = (blockIdx.x + i * gridDim.x) * C2T_TOK + lane;
      const float* xs = reinterpret_cast<const float*>(smem + L::X_OFF + s * L::X_BYTES);
      mbar_wait(&xfull[s], ph);
      float x[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = xs[(c0 + k) * C2T_TOK + lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&xempty[s]);
      uint8_t* gs = smem + L::G_OFF + s * L::G_BYTES + (slice >> 1) * 4096 + lane * 128;
      const uint32_t ch0 = static_cast<uint32_t>((slice & 1) * 4), sw = static_cast<uint32_t>(lane & 7);
      if (p.mode == 0) {
        float s1 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) s1 += x[k];
        p1[slice * 32 + lane] = s1;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(SL * 32) : "memory");
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < SL; ++k) tot += p1[k * 32 + lane];
        const float mean = tot * (1.0f / D);
        float s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float a = x[k] - mean;
          s2 += a * a;
        }
        p2[slice * 32 + lane] = s2;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(SL * 32) : "memory");
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < SL; ++k) var += p2[k * 32 + lane];
        const float rstd = rsqrtf(var * (1.0f / D) + p.eps);
        mbar_wait(&gfull[s], ph);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4* gp = reinterpret_cast<uint4*>(gs + (((ch0 + j) ^ sw) << 4));
          const uint4 gq = *gp;
          const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
          const float4 ga = *reinterpret_cast<const float4*>(aff + c0 + j * 8), gb = *reinterpret_cast<const float4*>(aff + c0 + j * 8 + 4);
          const float4 ba = *reinterpret_cast<const float4*>(aff + D + c0 + j * 8), bb = *reinterpret_cast<const float4*>(aff + D + c0 + j * 8 + 4);
          const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
          const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float g = (k & 1) ? bf16hi_to_f32(gw[k >> 1]) : bf16lo_to_f32(gw[k >> 1]);
            o[k] = ((x[j * 8 + k] - mean) * rstd * gm[k] + bt[k]) * g;
          }
          *gp = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
      } else {
        const float sc = p.scale ? ((token < p.T) ? __ldg(p.scale + token) : 0.f) : p.scale_const;
        mbar_wait(&gempty[s], ph ^ 1);          // no gate load implies the slot is free: wait for the store warp ourselves
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4* gp = reinterpret_cast<uint4*>(gs + (((ch0 + j) ^ sw) << 4));
          *gp = make_uint4(pack_bf16x2(x[j * 8 + 0] * sc, x[j * 8 + 1] * sc), pack_bf16x2(x[j * 8 + 2] * sc, x[j * 8 + 3] * sc),
                           pack_bf16x2(x[j * 8 + 4] * sc, x[j * 8 + 5] * sc), pack_bf16x2(x[j * 8 + 6] * sc, x[j * 8 + 7] * sc));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ostaged[s]);
    }
  }
}

}  // namespace af2
