// Axial self-attention, second generation: split-KV inside the CTA + P through tensor memory.
// Same contract as attention_tc.cuh (Attention.forward alphafold2.py:125-190 under the AxialAttention folds :228-253,
// shared additive pair bias, two-sided mask semantics of quirk Q1, sigmoid gate in the epilogue), different execution:
//
//   * TWO softmax groups (4 warps each, ONE THREAD PER QUERY ROW) work concurrently on alternating key blocks of the
//     CTA's block stream: group g owns every block with (global block index & 1) == g, the S accumulator buffer g, the
//     K/V/bias smem stage g and its own O accumulator.  Each group keeps its own running (max, sum); the two partial
//     results of a work item are merged in the epilogue  O = (O_A 2^(mA-m) + O_B 2^(mB-m)) / (lA 2^(mA-m) + lB 2^(mB-m)).
//     There is no cross-warp exchange inside a block (the first generation split a row between two warps and met at a
//     named barrier), and while one group is in its exp/sum phase the other is loading / storing tensor memory.
//   * a row is processed in two passes over tensor memory (pass 1: maximum, pass 2: exp2 / sum / pack), 32 columns at
//     a time, so a full 128-key row never has to live in registers;
//   * P goes back INTO the S buffer as packed bf16 (tcgen05.st) and the PV product reads it as the A operand from tensor
//     memory (tcgen05.mma with A in TMEM): no P tile in shared memory, no generic->async proxy fence per block;
//   * the pair bias is added BY THE TENSOR CORE: S = Q K^T + I_128 x Bias, eight extra K = 16 steps whose A operand is a
//     constant identity tile in shared memory and whose B operand is the TMA-staged bias tile [query][key] read MN-major.
//     The softmax threads (the bottleneck; the tensor pipe idles ~90 % of the time here) no longer load / unpack / add
//     128 x 128 bias values per block.
//
// TMEM (512 columns): S/P buffer 0 [0,128) | S/P buffer 1 [128,256) | O of group 0 [256, 256+DH) | O of group 1 [320, 320+DH)
// Warps (352 threads): 0 TMA producer | 1 MMA issuer | 2..5 softmax group 0 | 6..9 softmax group 1 | 10 key/query mask
#pragma once
#include "attention_tc.cuh"

namespace af2 {

constexpr int ATTN2_THREADS = 352;   // TMA warp, MMA warp, 2 x 4 softmax warps, key-mask warp

template <int DH>
struct Attn2Smem {
  static constexpr int Q_BYTES = 128 * DH * 2;
  static constexpr int K_BYTES = 128 * DH * 2;
  static constexpr int V_BYTES = 128 * DH * 2;
  static constexpr int BIAS_BYTES = 128 * 128 * 2;      // two 64-key boxes of [128 q rows x 128 B]
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES + BIAS_BYTES;
  static constexpr int Q_OFF = 0;
  static constexpr int G_OFF = Q_BYTES;                 // [2] sigmoid-gate tiles [128 q x DH]
  static constexpr int STAGE_OFF = 3 * Q_BYTES;
  static constexpr int IDENT_OFF = STAGE_OFF + 2 * STAGE_BYTES;   // identity [128 x 128] bf16, K-major SW128 (two 64-column halves)
  static constexpr int BAR_OFF = IDENT_OFF + 32768;
  static constexpr int KB_OFF = BAR_OFF + 256;          // float key term (0 / -inf) [2][128]
  static constexpr int ML_OFF = KB_OFF + 2 * 128 * 4;   // float (max, sum) of the two groups [2][128][2]
  static constexpr int QV_OFF = ML_OFF + 2 * 128 * 2 * 4;   // query-mask bytes [2][128]
  static constexpr int KF_OFF = QV_OFF + 2 * 128;       // [2] per-stage flag: some key of the block is masked / padding
  static constexpr int TOTAL = KF_OFF + 16;
};

template <int DH>
__global__ void __launch_bounds__(ATTN2_THREADS, 1)
attention2_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmBias,
                     const __grid_constant__ CUtensorMap tmG, const __grid_constant__ AttnParams p) {
  using L = Attn2Smem<DH>;
  constexpr uint32_t SWZ = (DH == 64) ? SWZ_128 : SWZ_64;
  constexpr uint32_t ROWB = DH * 2;                 // bytes per Q/K/V row
  constexpr uint32_t SBO = 8 * ROWB;                // 8-row swizzle atom
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* q_full = bars + 0;     // Q tile landed
  uint64_t* q_empty = bars + 1;    // Q tile consumed by the item's last QK^T
  uint64_t* g_full = bars + 2;     // [2] gate tile landed (slot it & 1)
  uint64_t* g_empty = bars + 4;    // [2] gate tile consumed by the epilogue (8 warps)
  uint64_t* kv_full = bars + 6;    // [2] K, V, bias of a block landed in stage
  uint64_t* kv_empty = bars + 8;   // [2] ... and consumed (PV retired)
  uint64_t* s_full = bars + 10;    // [2] S of the group's block complete
  uint64_t* p_full = bars + 12;    // [2] P written back to tensor memory (4 warps)
  uint64_t* pv_done = bars + 14;   // [2] PV of the group's block retired
  uint64_t* kb_full = bars + 16;   // [2] key-mask terms of a block staged
  uint64_t* kb_empty = bars + 18;  // [2] ... and consumed by the group's 4 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* keyb = reinterpret_cast<float*>(smem + L::KB_OFF);   // [2][128]
  float* mlbuf = reinterpret_cast<float*>(smem + L::ML_OFF);  // [2][128][2]
  uint8_t* qvbuf = smem + L::QV_OFF;                          // [2][128]
  uint32_t* kflag = reinterpret_cast<uint32_t*>(smem + L::KF_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkv = (p.n + 127) / 128;
  const int nqb = nkv;
  const int total_items = nqb * p.heads * p.nbatch;
  const int my_items = (total_items > static_cast<int>(blockIdx.x))
                           ? (total_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  // item id = ((h * nqb + qb) * nbatch + b'): neighbours in time differ only in b' and share the bias tile
  auto decode = [&](int it, int& qb_, int& h_, int& b_) {
    const int id = blockIdx.x + it * gridDim.x;
    b_ = id % p.nbatch;
    qb_ = (id / p.nbatch) % nqb;
    h_ = id / (p.nbatch * nqb);
  };
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t S_COL = 0, O_COL = 256;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmG);
    if (p.has_bias) prefetch_tmap(&tmBias);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&g_full[s], 1);
      mbar_init(&g_empty[s], 8);
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&pv_done[s], 1);
      mbar_init(&kb_full[s], 1);
      mbar_init(&kb_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (p.has_bias) {
    // identity tile: element (r, c) of half c/64 at r*128 + (((c%64)/8) ^ (r&7))*16 + (c%8)*2
    uint4* id4 = reinterpret_cast<uint4*>(smem + L::IDENT_OFF);
    for (int i = threadIdx.x; i < 2048; i += ATTN2_THREADS) id4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (threadIdx.x < 128) {
      const uint32_t r = threadIdx.x, c = threadIdx.x;
      *reinterpret_cast<__nv_bfloat16*>(smem + L::IDENT_OFF + (c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2) =
          __float2bfloat16(1.0f);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (int it = 0; it < my_items; ++it) {
        int qb, h, b;
        decode(it, qb, h, b);
        const int gs = it & 1;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, L::Q_BYTES);
        tma_load_4d(smem + L::Q_OFF, &tmQ, q_full, 0, qb * 128, h, b);
        mbar_wait(&g_empty[gs], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&g_full[gs], L::Q_BYTES);
        tma_load_4d(smem + L::G_OFF + gs * L::Q_BYTES, &tmG, &g_full[gs], 0, qb * 128, h, b);
        for (int j = 0; j < nkv; ++j) {
          const int g = it * nkv + j;
          const int st = g & 1;
          mbar_wait(&kv_empty[st], ((g >> 1) & 1) ^ 1);
          uint8_t* sk = smem + L::STAGE_OFF + st * L::STAGE_BYTES;
          uint8_t* sv = sk + L::K_BYTES;
          uint8_t* sbias = sv + L::V_BYTES;
          mbar_arrive_expect_tx(&kv_full[st], L::K_BYTES + L::V_BYTES + (p.has_bias ? L::BIAS_BYTES : 0));
          tma_load_4d(sk, &tmK, &kv_full[st], 0, j * 128, h, b);
          tma_load_4d(sv, &tmV, &kv_full[st], 0, j * 128, h, b);
          if (p.has_bias) {
            tma_load_3d(sbias, &tmBias, &kv_full[st], j * 128, qb * 128, h);
            tma_load_3d(sbias + 16384, &tmBias, &kv_full[st], j * 128 + 64, qb * 128, h);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S = Q K^T, both K-major
    constexpr uint32_t idesc_b = umma_idesc_bf16(128, 128, 0, 1);   // S += I Bias: identity K-major, bias tile MN-major
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);    // O = P V, P from tensor memory, V MN-major
    const int total_blocks = my_items * nkv;
    auto issue_s = [&](int g) {
      const int it = g / nkv, j = g - it * nkv;
      const int st = g & 1;
      if (j == 0) mbar_wait(q_full, it & 1);
      mbar_wait(&kv_full[st], (g >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sq = smem_u32(smem + L::Q_OFF);
        const uint32_t sk = smem_u32(smem + L::STAGE_OFF + st * L::STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint64_t ad = umma_smem_desc(sq + k * 32, 16, SBO, SWZ);
          const uint64_t bd = umma_smem_desc(sk + k * 32, 16, SBO, SWZ);
          umma_bf16(tmem_base + S_COL + st * 128, ad, bd, idesc_s, k != 0 ? 1u : 0u);
        }
        if (p.has_bias) {
          const uint32_t si = smem_u32(smem + L::IDENT_OFF);
          const uint32_t sbz = sk + L::K_BYTES + L::V_BYTES;      // bias tile: [128 query rows x 128 B] x two 64-key boxes
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // A: identity columns 16k..16k+15 (K-major, 64-column halves 16 KB apart)
            const uint64_t ad = umma_smem_desc(si + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, SWZ_128);
            // B: bias rows (= K index) 16k..16k+15, keys contiguous (MN-major): 8-row atoms 1024 B apart, 64-key boxes 16 KB apart
            const uint64_t bd = umma_smem_desc(sbz + k * 2048, 16384, 1024, SWZ_128);
            umma_bf16(tmem_base + S_COL + st * 128, ad, bd, idesc_b, 1u);
          }
        }
        umma_commit(&s_full[st]);
        if (j == nkv - 1) umma_commit(q_empty);          // Q buffer reusable once the item's last QK^T retires
      }
      __syncwarp();
    };
    // S(g) overwrites the buffer that held P(g-2): it is issued after PV(g-2), and the tensor core runs in issue order
    if (total_blocks > 0) issue_s(0);
    if (total_blocks > 1) issue_s(1);
    for (int g = 0; g < total_blocks; ++g) {
      const int st = g & 1;
      const int j = g % nkv;
      mbar_wait(&p_full[st], (g >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sv = smem_u32(smem + L::STAGE_OFF + st * L::STAGE_BYTES + L::K_BYTES);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A = P [128 rows x 16 keys] = 8 packed columns of the S/P buffer; B = V [key][dh], MN-major: 16 keys = 2 atoms
          const uint64_t bd = umma_smem_desc(sv + k * 2 * SBO, 16, SBO, SWZ);
          umma_bf16_ts(tmem_base + O_COL + st * 64, tmem_base + S_COL + st * 128 + k * 8, bd, idesc_o, (j >= 2 || k != 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&pv_done[st]);
      }
      __syncwarp();
      if (g + 2 < total_blocks) issue_s(g + 2);
    }
  } else if (warp == 10) {
    // ================================ key-mask warp ================================
    const float NEG_INF = -__int_as_float(0x7f800000);
    for (int it = 0; it < my_items; ++it) {
      int qb, h, b;
      decode(it, qb, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int g = it * nkv + j;
        const int st = g & 1;
        float kb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kidx = j * 128 + lane * 4 + i;
          kb[i] = NEG_INF;
          if (kidx < p.n) kb[i] = (!p.mask || p.mask[b * p.mask_sb + kidx * p.mask_si]) ? 0.f : NEG_INF;
        }
        uint32_t qv = 0x01010101u;
        if (j == 0 && p.mask) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int qi = qb * 128 + lane * 4 + i;
            if (qi < p.n && p.mask[b * p.mask_sb + qi * p.mask_si] == 0) qv &= ~(0xffu << (8 * i));
          }
        }
        const bool some_masked = __any_sync(0xffffffffu, (kb[0] != 0.f) | (kb[1] != 0.f) | (kb[2] != 0.f) | (kb[3] != 0.f));
        mbar_wait(&kb_empty[st], ((g >> 1) & 1) ^ 1);
        if (lane == 0) kflag[st] = some_masked ? 1u : 0u;
        *reinterpret_cast<float4*>(keyb + st * 128 + lane * 4) = make_float4(kb[0], kb[1], kb[2], kb[3]);
        if (j == 0) *reinterpret_cast<uint32_t*>(qvbuf + (it & 1) * 128 + lane * 4) = qv;
        __syncwarp();
        if (lane == 0) mbar_arrive(&kb_full[st]);
      }
    }
  } else {
    // ================================ softmax groups + epilogue (warps 2..9) ==============
    const int q = warp & 3;               // TMEM lane quarter
    const int grp = (warp - 2) >> 2;      // softmax group = S buffer = smem stage = parity of the blocks it owns
    const int r = q * 32 + lane;          // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const float NEG_INF = -__int_as_float(0x7f800000);
    const uint32_t s_buf = tmem_base + S_COL + grp * 128 + lane_sel;
    const uint32_t o_mine = tmem_base + O_COL + grp * 64 + lane_sel;
    const float* kbs = keyb + grp * 128;
    uint32_t nb = 0;                      // blocks this group has processed so far (phase counter of its barriers)

    // finish the logits of 32 keys [c4*32, c4*32+32) of this thread's row (log2 domain; the pair bias is already in S):
    // key term of partially masked blocks, and the uniform row of a masked query
    auto fix_logits = [&](int c4, int j, bool keys_masked, bool q_valid, uint32_t (&u)[32]) {
      if (keys_masked) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 k4 = *reinterpret_cast<const float4*>(kbs + c4 * 32 + c * 4);
          u[c * 4 + 0] = __float_as_uint(__uint_as_float(u[c * 4 + 0]) + k4.x);
          u[c * 4 + 1] = __float_as_uint(__uint_as_float(u[c * 4 + 1]) + k4.y);
          u[c * 4 + 2] = __float_as_uint(__uint_as_float(u[c * 4 + 2]) + k4.z);
          u[c * 4 + 3] = __float_as_uint(__uint_as_float(u[c * 4 + 3]) + k4.w);
        }
      }
      if (!q_valid) {                        // rare: masked query row -> uniform over the n real keys
#pragma unroll
        for (int k = 0; k < 32; ++k) u[k] = __float_as_uint((j * 128 + c4 * 32 + k) < p.n ? 0.f : NEG_INF);
      }
    };
    auto chunk_max = [&](const uint32_t (&u)[32]) {
      float m0 = fmaxf(__uint_as_float(u[0]), __uint_as_float(u[1])), m1 = fmaxf(__uint_as_float(u[2]), __uint_as_float(u[3]));
#pragma unroll
      for (int k = 4; k < 32; k += 4) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(u[k]), __uint_as_float(u[k + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(u[k + 2]), __uint_as_float(u[k + 3])));
      }
      return fmaxf(m0, m1);
    };

    for (int it = 0; it < my_items; ++it) {
      int qb, h, b;
      decode(it, qb, h, b);
      const int qi = qb * 128 + r;
      const bool q_in = qi < p.n;
      float m_run = NEG_INF, l_run = 0.f;
      bool q_valid = true, first = true;
      // j runs over the key blocks of this item whose global block index it*nkv + j has parity grp
      for (int j = ((it * nkv) & 1) ^ grp; j < nkv; j += 2) {
        mbar_wait(&kb_full[grp], nb & 1);           // key terms (and the item's query-mask bytes) staged
        if (first) q_valid = qvbuf[(it & 1) * 128 + r] != 0;
        const bool keys_masked = kflag[grp] != 0u;
        mbar_wait(&s_full[grp], nb & 1);
        tc_fence_after();
        // ---- pass 1: row maximum; the next 32-column chunk is in flight while one is reduced ----
        uint32_t ua[32], ub[32];
        float mx;
        tmem_ld32(s_buf, ua);
        tmem_ld_wait();
        tmem_ld32(s_buf + 32, ub);
        fix_logits(0, j, keys_masked, q_valid, ua);
        mx = chunk_max(ua);
        tmem_ld_wait();
        tmem_ld32(s_buf + 64, ua);
        fix_logits(1, j, keys_masked, q_valid, ub);
        mx = fmaxf(mx, chunk_max(ub));
        tmem_ld_wait();
        tmem_ld32(s_buf + 96, ub);
        fix_logits(2, j, keys_masked, q_valid, ua);
        mx = fmaxf(mx, chunk_max(ua));
        tmem_ld_wait();
        tmem_ld32(s_buf, ua);                              // first chunk of pass 2
        fix_logits(3, j, keys_masked, q_valid, ub);
        mx = fmaxf(mx, chunk_max(ub));
        float m_new = fmaxf(m_run, mx);
        // lazy rescaling (see attention_tc.cuh): keep a stale maximum while the new one exceeds it by <= 2^8
        bool rescale = true;
        if (!first) {
          rescale = __any_sync(0xffffffffu, m_new > m_run + 8.0f);
          if (!rescale) m_new = m_run;
        }
        const float m_use = (m_new == NEG_INF) ? 0.f : m_new;
        const float corr = fast_exp2(m_run - m_use);     // m_run = -inf -> 0
        if (!first) {
          mbar_wait(&pv_done[grp], (nb - 1) & 1);        // the group's previous P V has retired: O may be rescaled
          tc_fence_after();
          if (rescale) {
#pragma unroll
            for (int cb = 0; cb < DH; cb += 32) {
              uint32_t o[32];
              tmem_ld32(o_mine + cb, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st32(o_mine + cb, o);
            }
          }
        }
        // ---- pass 2: exp2, row sum, P -> tensor memory (packed bf16 pairs over the S columns already consumed) ----
        float ls0 = 0.f, ls1 = 0.f;
        auto emit = [&](int c4, uint32_t (&u)[32]) {
          fix_logits(c4, j, keys_masked, q_valid, u);
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const float e0 = fast_exp2(__uint_as_float(u[k]) - m_use), e1 = fast_exp2(__uint_as_float(u[k + 1]) - m_use);
            ls0 += e0; ls1 += e1;
            pk[k >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st16(s_buf + c4 * 16, pk);
        };
        tmem_ld_wait();
        tmem_ld32(s_buf + 32, ub);
        emit(0, ua);
        tmem_ld_wait();
        tmem_ld32(s_buf + 64, ua);
        emit(1, ub);
        tmem_ld_wait();
        tmem_ld32(s_buf + 96, ub);
        emit(2, ua);
        tmem_ld_wait();
        emit(3, ub);
        l_run = l_run * corr + (ls0 + ls1);
        m_run = m_new;
        __syncwarp();
        if (lane == 0) mbar_arrive(&kb_empty[grp]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[grp]);
        first = false;
        ++nb;
      }
      // ---- merge the two groups' partial results and write the gated output ----
      mlbuf[(grp * 128 + r) * 2] = m_run;
      mlbuf[(grp * 128 + r) * 2 + 1] = l_run;
      if (!first) {
        mbar_wait(&pv_done[grp], (nb - 1) & 1);          // this group's last P V of the item has retired
        tc_fence_after();
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // both groups: (m, l) published, both O complete
      const float m_o = mlbuf[((grp ^ 1) * 128 + r) * 2], l_o = mlbuf[((grp ^ 1) * 128 + r) * 2 + 1];
      const float m_all = fmaxf(m_run, m_o);
      const float m_ref = (m_all == NEG_INF) ? 0.f : m_all;
      const float f_me = (l_run > 0.f) ? fast_exp2(m_run - m_ref) : 0.f;
      const float f_ot = (l_o > 0.f) ? fast_exp2(m_o - m_ref) : 0.f;
      const float inv_l = 1.0f / (l_run * f_me + l_o * f_ot);
      const float fa = (grp == 0 ? f_me : f_ot) * inv_l, fb = (grp == 0 ? f_ot : f_me) * inv_l;   // factors of O_A / O_B
      const bool use_a = (grp == 0 ? l_run : l_o) > 0.f, use_b = (grp == 0 ? l_o : l_run) > 0.f;
      mbar_wait(&g_full[it & 1], (it >> 1) & 1);
      {
        // this thread writes DH/2 output columns [grp*DH/2, +DH/2) of its row from BOTH partial accumulators
        constexpr int HC = DH / 2;
        const uint32_t oa = tmem_base + O_COL + lane_sel + grp * HC, ob = tmem_base + O_COL + 64 + lane_sel + grp * HC;
        const long long tok = b * p.tok_sb + static_cast<long long>(qi) * p.tok_si;
        __nv_bfloat16* op = p.out + tok * p.ld_out + h * DH + grp * HC;
        const uint8_t* gt = smem + L::G_OFF + (it & 1) * L::Q_BYTES;
#pragma unroll
        for (int cb = 0; cb < HC; cb += 16) {
          uint32_t va[16], vb[16];
          tmem_ld16(oa + cb, va);
          tmem_ld16(ob + cb, vb);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int chunk = (grp * HC + cb) / 8 + i;        // 16-byte chunk of the gate row (8 bf16)
            const uint32_t goff = (DH == 64) ? swz128_off(r, chunk) : (r * 64u + ((static_cast<uint32_t>(chunk) ^ ((r >> 1) & 3u)) << 4));
            const uint4 gq = *reinterpret_cast<const uint4*>(gt + goff);
            const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
            uint32_t ow[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int e = 8 * i + 2 * t;
              const float a0 = (use_a ? __uint_as_float(va[e]) * fa : 0.f) + (use_b ? __uint_as_float(vb[e]) * fb : 0.f);
              const float a1 = (use_a ? __uint_as_float(va[e + 1]) * fa : 0.f) + (use_b ? __uint_as_float(vb[e + 1]) * fb : 0.f);
              ow[t] = pack_bf16x2(a0 * bf16lo_to_f32(gw[t]), a1 * bf16hi_to_f32(gw[t]));
            }
            if (q_in) *reinterpret_cast<uint4*>(op + cb + 8 * i) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_empty[it & 1]);
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");     // nobody still reads an O accumulator when the next item's P V starts
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace af2
