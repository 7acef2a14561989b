// Generic batched bf16 GEMM on tcgen05 tensor cores with TMA-staged operands and a programmable
// per-column-tile epilogue.  One kernel serves every dense contraction of the Evoformer block:
//
//   K-major mode  : C[b][m][n] = sum_k A[b][m][k] * B[b][n][k]      (Linear layers: B = weight [out,in];
//                                                                    triangle "outgoing" per channel)
//   MN-major mode : C[b][m][n] = sum_k A[b][k][m] * B[b][k][n]      (triangle "ingoing", outer-mean:
//                                                                    channel-major operands, k = row axis)
//
// Tiling: BM = 128 rows (one TMEM lane per row), BN in {64,128,256} accumulator columns, BK = 64.
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> global).  Persistent over
// tiles; the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Operands use the 128B swizzle written by TMA and read back by the UMMA descriptors.
#pragma once
#include "common.cuh"

namespace af2 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

enum EpiMode : int { EPI_STORE_BF16 = 0, EPI_GATED_BF16 = 1, EPI_RESID_F32 = 2, EPI_STORE_F32 = 3 };
enum EpiAct : int { ACT_NONE = 0, ACT_SIGMOID = 1, ACT_GELU = 2 };
enum EpiLayout : int { LAYOUT_TOKEN = 0, LAYOUT_CHANNEL = 1 };

// Epilogue program of one BN-wide accumulator column tile.
struct NTile {
  int mode;        // EpiMode
  int act;         // EpiAct (applied to the value for STORE, to the gate half for GATED)
  int layout;      // EpiLayout
  int col0;        // first output column (token-major) / first output channel (channel-major)
  int ncols;       // number of valid OUTPUT columns of this tile
  int use_rowscale;
  void* out;
  const float* bias;   // [BN] accumulator-column bias of this tile, or nullptr
  long long ld;        // token-major: row stride (elements); channel-major: channel stride (elements)
};

// Every column tile runs the same epilogue program `tile`; tile nt produces output columns
// [nt*W, nt*W + W) with W = BN (BN/2 for EPI_GATED, whose weight rows are packed per tile as
// [value rows of the tile | gate rows of the tile]), clipped to out_cols.
struct GemmParams {
  int M, N, K, batch;          // N = accumulator columns = rows of the B operand
  int num_ntiles;
  int out_cols;                // valid output columns (N, or N/2 for EPI_GATED)
  const float* rowscale;       // [batch*M] multiplier per row (mask), or nullptr
  const float* resid;          // EPI_RESID_F32: fp32 [M, ld_resid]
  long long ld_resid;
  long long out_batch_stride;  // elements
  int cm_inner, cm_pitch;      // channel-major: row r -> (r / cm_inner) * cm_pitch + r % cm_inner
  NTile tile;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;       // barriers + 1 KB alignment slack
};

template <int BN, int STAGES, bool MN_MAJOR>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams p) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = p.num_ntiles;
  const int total_tiles = p.batch * m_tiles * n_tiles;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        const int mt = (tile / n_tiles) % m_tiles;
        const int b = tile / (n_tiles * m_tiles);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if constexpr (!MN_MAJOR) {
            tma_load_3d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, mt * GEMM_BM, b);
            tma_load_3d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, nt * BN, b);
          } else {
            // operand stored [k][mn]: 64-wide mn boxes, each 64 k-rows x 128 B = 8 KB
#pragma unroll
            for (int h = 0; h < GEMM_BM / 64; ++h)
              tma_load_3d(sa + h * 8192, &tmA, &full_bar[stage], mt * GEMM_BM + h * 64, kb * GEMM_BK, b);
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_3d(sb + h * 8192, &tmB, &full_bar[stage], nt * BN + h * 64, kb * GEMM_BK, b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, MN_MAJOR ? 1 : 0, MN_MAJOR ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t adesc, bdesc;
            if constexpr (!MN_MAJOR) {
              // rows of 128 B (64 bf16 of K); 8-row swizzle atoms 1024 B apart; K step = 32 B
              adesc = umma_smem_desc(sa + k * 32, 16, 1024, SWZ_128);
              bdesc = umma_smem_desc(sb + k * 32, 16, 1024, SWZ_128);
            } else {
              // k-rows of 128 B (64 bf16 of M/N); 8 k-rows = 1024 B (SBO); next 64-wide mn group 8 KB (LBO)
              adesc = umma_smem_desc(sa + k * 2048, 8192, 1024, SWZ_128);
              bdesc = umma_smem_desc(sb + k * 2048, 8192, 1024, SWZ_128);
            }
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // smem slot free once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % n_tiles;
      const int mt = (tile / n_tiles) % m_tiles;
      const int b = tile / (n_tiles * m_tiles);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      NTile t = p.tile;
      {
        const int w = (t.mode == EPI_GATED_BF16) ? BN / 2 : BN;
        t.col0 = nt * w;
        t.ncols = min(w, p.out_cols - nt * w);
        if (t.bias) t.bias += nt * BN;
      }
      const int row = mt * GEMM_BM + row_in_tile;
      const bool row_ok = row < p.M;
      const long long grow = static_cast<long long>(b) * p.M + row;   // row index over the whole batch
      float rs = 1.0f;
      if (t.use_rowscale && row_ok) rs = __ldg(p.rowscale + grow);
      long long cm_off = 0;
      if (t.layout == LAYOUT_CHANNEL) cm_off = static_cast<long long>(row / p.cm_inner) * p.cm_pitch + row % p.cm_inner;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);

      if (t.mode == EPI_GATED_BF16) {
        constexpr int HALF = BN / 2;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(t.out) + b * p.out_batch_stride;
#pragma unroll 1
        for (int c0 = 0; c0 < HALF; c0 += 32) {
          uint32_t u[32], g[32];
          tmem_ld32(t_acc + c0, u);
          tmem_ld32(t_acc + HALF + c0, g);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float uu = __uint_as_float(u[j]), gg = __uint_as_float(g[j]);
            if (t.bias) {
              uu += __ldg(t.bias + c0 + j);
              gg += __ldg(t.bias + HALF + c0 + j);
            }
            const float a = (t.act == ACT_GELU) ? gelu_erf(gg) : sigmoidf_fast(gg);
            v[j] = uu * a * rs;
          }
          if (row_ok) {
            if (t.layout == LAYOUT_TOKEN) {
              __nv_bfloat16* dst = out + static_cast<long long>(row) * t.ld + t.col0 + c0;
              if (c0 + 32 <= t.ncols) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 pk = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                        pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                  *reinterpret_cast<uint4*>(dst + j) = pk;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (c0 + j < t.ncols) dst[j] = __float2bfloat16(v[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < t.ncols) out[static_cast<long long>(t.col0 + c0 + j) * t.ld + cm_off] = __float2bfloat16(v[j]);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (c0 >= t.ncols) break;           // warp-uniform
          uint32_t u[32];
          tmem_ld32(t_acc + c0, u);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(u[j]);
            if (t.bias) x += __ldg(t.bias + c0 + j);
            if (t.act == ACT_SIGMOID) x = sigmoidf_fast(x);
            else if (t.act == ACT_GELU) x = gelu_erf(x);
            v[j] = x * rs;
          }
          if (!row_ok) continue;
          const bool full = (c0 + 32 <= t.ncols);
          if (t.mode == EPI_STORE_BF16) {
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(t.out) + b * p.out_batch_stride;
            if (t.layout == LAYOUT_TOKEN) {
              __nv_bfloat16* dst = out + static_cast<long long>(row) * t.ld + t.col0 + c0;
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 pk = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                        pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                  *reinterpret_cast<uint4*>(dst + j) = pk;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (c0 + j < t.ncols) dst[j] = __float2bfloat16(v[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < t.ncols) out[static_cast<long long>(t.col0 + c0 + j) * t.ld + cm_off] = __float2bfloat16(v[j]);
            }
          } else {
            float* out = reinterpret_cast<float*>(t.out) + b * p.out_batch_stride;
            float* dst = out + static_cast<long long>(row) * t.ld + t.col0 + c0;
            if (t.mode == EPI_RESID_F32) {
              const float* rsd = p.resid + static_cast<long long>(row) * p.ld_resid + t.col0 + c0;
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 r4 = *reinterpret_cast<const float4*>(rsd + j);
                  *reinterpret_cast<float4*>(dst + j) =
                      make_float4(v[j] + r4.x, v[j + 1] + r4.y, v[j + 2] + r4.z, v[j + 3] + r4.w);
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (c0 + j < t.ncols) dst[j] = v[j] + rsd[j];
              }
            } else {
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
                for (int j = 0; j < 32; ++j)
                  if (c0 + j < t.ncols) dst[j] = v[j];
              }
            }
          }
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace af2
