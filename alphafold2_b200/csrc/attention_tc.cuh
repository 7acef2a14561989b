// Axial self-attention with shared additive pair bias and the reference's two-sided mask, on tcgen05.
// Replaces Attention.forward (alphafold2.py:125-190) as driven by AxialAttention (alphafold2.py:219-255).
//
// Persistent CTAs; a work item = one (folded batch element b', head h, block of 128 queries).  Items are ordered
// (h, query block, b') and every CTA owns a CONTIGUOUS range of them, so its items share the (h, query block) pair-bias
// tiles: for n <= 256 those [128 q x n k] bf16 tiles stay RESIDENT in shared memory and are reloaded only when the range
// crosses into the next (h, query block) (at most twice per CTA at C2).  That removes 64 of the 160 KB that every item
// used to pull from L2 -- the kernel is bound by the L2 -> SM fill rate, not by the tensor pipe.  Keys/values are
// streamed in blocks of 128 through a TMA pipeline (2 stages with a bias, 4 without; for n > 256 the bias tile of the
// block travels with its K/V stage); the producer runs ahead across work items.
//   S_j = Q K_j^T + I B_j   tcgen05.mma  (M=128, N=128, K=DH, then K=128 against a constant identity tile): the pair bias
//                           tile B_j [query][key] is the MN-major B operand, so the TENSOR CORE adds the bias and the
//                           softmax threads never load / unpack / add it  -> TMEM (double buffered, 2 x 128 cols)
//   softmax warps (8)       two threads per query row (each owns 64 of the 128 keys of the block):
//                           TMEM -> regs, (+ key mask), online max/sum (row max exchanged through smem),
//                           P_j -> packed bf16 written back over the S columns it came from (tcgen05.st); O is
//                           rescaled in TMEM only when the running max moves by more than 2^8
//   O  += P_j V_j           tcgen05.mma with A = P_j read FROM TENSOR MEMORY (M=128, N=DH, K=128), V consumed MN-major
//                           straight from its [key][dh] layout.  No P tile in shared memory, no proxy fence per block,
//                           and P is double buffered with S, so the softmax never waits for the previous P V.
//   epilogue                O / l * sigmoid-gate -> bf16, written IN PLACE over the gate tile in shared memory and sent
//                           to [token, h*DH + e] by one TMA store per 32 rows (a thread-per-row STG.128 pattern
//                           kept the LSU busy for ~1.5k cycles per item: 32 lines per instruction)
// Logits are produced directly in the log2 domain: the host folds dim_head^-0.5 * log2(e) into to_q and
// log2(e) into edges_to_attn_bias, so the softmax is exp2(v - max) with no per-element scaling.
//
// Mask semantics (quirk Q1): logits where !(mask[q] & mask[k]) are REPLACED by -FLT_MAX.  For an unmasked
// query that gives probability exactly 0 on masked keys; for a masked query every logit is equal, i.e. a
// uniform distribution over all n keys (masked ones included).  Keys >= n (tile padding) never count.
#pragma once
#include "common.cuh"

namespace af2 {

struct AttnParams {
  int n;              // sequence length along the attended axis
  int heads;
  int nbatch;         // folded batch B'
  int has_bias;
  const uint8_t* mask;        // nullptr or bool mask; element (b', i) at mask[b'*mask_sb + i*mask_si]
  long long mask_sb, mask_si;
  const __nv_bfloat16* gate;  // sigmoid(gating) [token, heads*DH]; token(b', i) = b'*tok_sb + i*tok_si
  __nv_bfloat16* out;         // [token, heads*DH]
  long long tok_sb, tok_si;
  long long ld_gate, ld_out;
};

constexpr int ATTN_THREADS = 384;   // Q/G/K/bias TMA warp, MMA warp, 8 softmax warps, key-mask warp, V TMA warp

template <int DH>
struct AttnSmem {
  static constexpr int Q_BYTES = 128 * DH * 2;
  static constexpr int K_BYTES = 128 * DH * 2;
  static constexpr int V_BYTES = 128 * DH * 2;
  static constexpr int BIAS_BYTES = 128 * 128 * 2;      // two 64-key boxes of [128 q rows x 128 B]
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES + BIAS_BYTES;
  // Identity operand of the bias MMA.  K step k needs A_k[r][e] = (r == 16k + e): only the two 8-row groups 2k, 2k+1 are
  // non-zero and they look the same for every k, so ONE strip of 30 row groups (32-byte swizzle atoms of 8 rows x 16
  // columns, 256 B each) -- 14 zero groups, the two diagonal groups, 14 zero groups -- serves all eight steps: step k
  // starts its descriptor (14 - 2k) groups into the strip.
  static constexpr int IDENT_BYTES = 32 * 256;
  static constexpr int Q_OFF = 0;                       // [2] Q tiles (slot it & 1)
  static constexpr int G_OFF = 2 * Q_BYTES;             // [2] sigmoid-gate tiles [128 q x DH] (same layout as Q)
  // K/V (+ bias) region of 2 * STAGE_BYTES, carved at run time:
  //   resident bias (n <= 256): [K V] x 2 stages | bias tiles of key blocks 0, 1
  //   streamed bias (n > 256) : [K V bias] x 2 stages
  //   no bias                 : [K V] x 4 stages
  static constexpr int KV_BYTES = K_BYTES + V_BYTES;
  static constexpr int STAGE_OFF = 4 * Q_BYTES;
  static constexpr int IDENT_OFF = STAGE_OFF + 2 * STAGE_BYTES;
  static constexpr int BAR_OFF = IDENT_OFF + IDENT_BYTES;
  static constexpr int KB_OFF = BAR_OFF + 320;          // float key term (0 / -inf) [2][128]
  static constexpr int MX_OFF = KB_OFF + 2 * 128 * 4;   // float row-max / row-sum exchange [2][128]
  static constexpr int L_OFF = MX_OFF + 2 * 128 * 4;    // float row-sum exchange [2][128]
  static constexpr int QV_OFF = L_OFF + 2 * 128 * 4;    // query-mask bytes [2][128]
  static constexpr int KF_OFF = QV_OFF + 2 * 128;       // [2] per-stage flag: some key of the block is masked / padding
  static constexpr int TOTAL = KF_OFF + 16 + 1024;
};

// tmQ/tmK/tmV: 4-D maps over the projection buffer, dims (e [DH], i [n], h [heads], b' [nbatch]),
// box (DH, 128, 1, 1), swizzle = DH*2 bytes.  tmBias: 3-D (k [npad], q [n], h), box (64, 128, 1), SW128.
template <int DH>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmBias,
                    const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmO,
                    const __grid_constant__ AttnParams p) {
  using L = AttnSmem<DH>;
  constexpr uint32_t SWZ = (DH == 64) ? SWZ_128 : SWZ_64;
  constexpr uint32_t ROWB = DH * 2;                 // bytes per Q/K/V row
  constexpr uint32_t SBO = 8 * ROWB;                // 8-row swizzle atom
  // 1024-byte aligned by declaration (128B-swizzle atoms); keeping the array symbol (no integer round-up of the pointer)
  // lets the compiler prove the shared address space and emit LDS/STS instead of generic LD/ST
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* q_full = bars + 0;     // [2] Q tile landed (slot it & 1)
  uint64_t* g_full = bars + 2;     // [2] gate tile landed (slot it & 1)
  uint64_t* q_empty = bars + 4;    // [2] Q tile consumed by the item's last QK^T
  uint64_t* k_full = bars + 6;     // [4] K (+ streamed bias) of a block landed
  uint64_t* k_empty = bars + 10;   // [4] ... and consumed by the block's S MMAs
  uint64_t* v_full = bars + 14;    // [4] V of a block landed
  uint64_t* v_empty = bars + 18;   // [4] ... and consumed by the block's P V
  uint64_t* s_full = bars + 22;    // [2]
  uint64_t* p_full = bars + 24;    // [2] P of the block is in its S buffer
  uint64_t* pv_done = bars + 26;
  uint64_t* kb_full = bars + 27;   // [2] key-mask terms of a block staged
  uint64_t* kb_empty = bars + 29;  // [2] ... and consumed by the 8 softmax warps
  uint64_t* g_empty = bars + 31;   // [2] gate / output tile drained by the TMA stores of the 4 row quarters
  uint64_t* bias_full = bars + 33; // resident bias tiles of the current (h, query block) landed
  uint64_t* bias_empty = bars + 34;// ... and consumed by the last S MMA of that (h, query block)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 35);
  float* keyb = reinterpret_cast<float*>(smem + L::KB_OFF);   // [2][128]
  float* mxbuf = reinterpret_cast<float*>(smem + L::MX_OFF);  // [2][128]
  float* lbuf = reinterpret_cast<float*>(smem + L::L_OFF);    // [2][128]
  uint8_t* qvbuf = smem + L::QV_OFF;                          // [2][128]
  uint32_t* kflag = reinterpret_cast<uint32_t*>(smem + L::KF_OFF);   // [2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkv = (p.n + 127) / 128;
  const int nqb = nkv;
  const int total_items = nqb * p.heads * p.nbatch;
  // item id = ((h * nqb + qb) * nbatch + b'); this CTA owns ids [item0, item0 + my_items)
  const int item0 = static_cast<int>(static_cast<long long>(total_items) * blockIdx.x / gridDim.x);
  const int my_items = static_cast<int>(static_cast<long long>(total_items) * (blockIdx.x + 1) / gridDim.x) - item0;
  auto decode = [&](int it, int& qb_, int& h_, int& b_) {
    const int id = item0 + it;
    b_ = id % p.nbatch;
    qb_ = (id / p.nbatch) % nqb;
    h_ = id / (p.nbatch * nqb);
  };
  auto combo_of = [&](int it) { return (item0 + it) / p.nbatch; };       // (h, query block) index: selects the bias tiles
  const bool resident = p.has_bias && nkv <= 2;
  const int nst = p.has_bias ? 2 : 4;                                    // K/V pipeline depth
  const int stage_stride = (p.has_bias && !resident) ? L::STAGE_BYTES : L::KV_BYTES;
  const int bias_res_off = L::STAGE_OFF + 2 * L::KV_BYTES;               // resident tiles: + j * BIAS_BYTES
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t S_COL = 0, O_COL = 256;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmG);
    if (p.has_bias) prefetch_tmap(&tmBias);
    prefetch_tmap(&tmO);
    mbar_init(bias_full, 1);
    mbar_init(bias_empty, 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&g_full[s], 1);
      mbar_init(&g_empty[s], 4);
      mbar_init(&s_full[s], 1);
      mbar_init(&kb_full[s], 1);
      mbar_init(&kb_empty[s], 8);
    }
    mbar_init(&p_full[0], 8);
    mbar_init(&p_full[1], 8);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (p.has_bias) {
    uint4* id4 = reinterpret_cast<uint4*>(smem + L::IDENT_OFF);
    for (int i = threadIdx.x; i < L::IDENT_BYTES / 16; i += ATTN_THREADS) id4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (threadIdx.x < 16) {
      // diagonal element i of the 16 x 16 block: row group 14 + i / 8, row r = i % 8, column i (16-byte chunk i / 8,
      // XOR-ed with bit 7 of the byte address = r >> 2 by the 32-byte swizzle)
      const uint32_t i = threadIdx.x, r = i & 7, c = i >> 3;
      *reinterpret_cast<__nv_bfloat16*>(smem + L::IDENT_OFF + (14 + c) * 256 + r * 32 + ((c ^ (r >> 2)) << 4) + (i & 7) * 2) =
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
      int prev_combo = -1, nc = 0;
      for (int it = 0; it < my_items; ++it) {
        int qb, h, b;
        decode(it, qb, h, b);
        const int gs = it & 1;
        if (resident && combo_of(it) != prev_combo) {
          prev_combo = combo_of(it);
          mbar_wait(bias_empty, (nc & 1) ^ 1);
          mbar_arrive_expect_tx(bias_full, nkv * L::BIAS_BYTES);
          for (int j = 0; j < nkv; ++j) {
            uint8_t* sb = smem + bias_res_off + j * L::BIAS_BYTES;
            tma_load_3d(sb, &tmBias, bias_full, j * 128, qb * 128, h);
            tma_load_3d(sb + 16384, &tmBias, bias_full, j * 128 + 64, qb * 128, h);
          }
          ++nc;
        }
        mbar_wait(&q_empty[gs], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[gs], L::Q_BYTES);
        tma_load_4d(smem + L::Q_OFF + gs * L::Q_BYTES, &tmQ, &q_full[gs], 0, qb * 128, h, b);
        mbar_wait(&g_empty[gs], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&g_full[gs], L::Q_BYTES);
        tma_load_4d(smem + L::G_OFF + gs * L::Q_BYTES, &tmG, &g_full[gs], 0, qb * 128, h, b);
        const bool stream_bias = p.has_bias && !resident;
        for (int j = 0; j < nkv; ++j) {
          const int g = it * nkv + j;
          const int kst = g % nst;
          // K is dead as soon as the block's S MMAs retire -- a whole softmax earlier than V -- so its stage refills early
          mbar_wait(&k_empty[kst], ((g / nst) & 1) ^ 1);
          uint8_t* sk = smem + L::STAGE_OFF + kst * stage_stride;
          mbar_arrive_expect_tx(&k_full[kst], L::K_BYTES + (stream_bias ? L::BIAS_BYTES : 0));
          tma_load_4d(sk, &tmK, &k_full[kst], 0, j * 128, h, b);
          if (stream_bias) {
            uint8_t* sbias = sk + L::KV_BYTES;
            tma_load_3d(sbias, &tmBias, &k_full[kst], j * 128, qb * 128, h);
            tma_load_3d(sbias + 16384, &tmBias, &k_full[kst], j * 128 + 64, qb * 128, h);
          }
        }
      }
    }
  } else if (warp == 11) {
    // ================================ V producer ==================================
    if (lane == 0) {
      for (int it = 0; it < my_items; ++it) {
        int qb, h, b;
        decode(it, qb, h, b);
        for (int j = 0; j < nkv; ++j) {
          const int g = it * nkv + j;
          const int kst = g % nst;
          mbar_wait(&v_empty[kst], ((g / nst) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[kst], L::V_BYTES);
          tma_load_4d(smem + L::STAGE_OFF + kst * stage_stride + L::K_BYTES, &tmV, &v_full[kst], 0, j * 128, h, b);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S = Q K^T, both K-major
    constexpr uint32_t idesc_b = umma_idesc_bf16(128, 128, 0, 1);   // S += I B: identity K-major, bias tile MN-major
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);    // O = P V, P from tensor memory, V MN-major
    const int total_blocks = my_items * nkv;
    int prev_combo = -1, nc = 0;
    auto issue_s = [&](int g) {
      const int it = g / nkv, j = g - it * nkv;
      const int st = g & 1, kst = g % nst;
      if (j == 0) {
        mbar_wait(&q_full[it & 1], (it >> 1) & 1);
        if (resident && combo_of(it) != prev_combo) {     // this CTA's range entered the next (h, query block): new bias tiles
          prev_combo = combo_of(it);
          mbar_wait(bias_full, nc & 1);
          ++nc;
        }
      }
      mbar_wait(&k_full[kst], (g / nst) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sq = smem_u32(smem + L::Q_OFF + (it & 1) * L::Q_BYTES);
        const uint32_t sk = smem_u32(smem + L::STAGE_OFF + kst * stage_stride);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint64_t ad = umma_smem_desc(sq + k * 32, 16, SBO, SWZ);
          const uint64_t bd = umma_smem_desc(sk + k * 32, 16, SBO, SWZ);
          umma_bf16(tmem_base + S_COL + st * 128, ad, bd, idesc_s, k != 0 ? 1u : 0u);
        }
        if (p.has_bias) {
          const uint32_t si = smem_u32(smem + L::IDENT_OFF);
          // bias tile of key block j: [128 query rows x 128 B] x two 64-key boxes
          const uint32_t sbz = resident ? smem_u32(smem + bias_res_off + j * L::BIAS_BYTES) : sk + L::KV_BYTES;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // A: identity columns 16k .. 16k+15 = the diagonal strip entered (14 - 2k) row groups in (see AttnSmem)
            const uint64_t ad = umma_smem_desc(si + (14 - 2 * k) * 256, 16, 256, SWZ_32);
            // B: bias rows (the K index) 16k .. 16k+15 with the keys contiguous (MN-major): 8-row atoms 1024 B apart,
            // the two 64-key boxes 16 KB apart
            const uint64_t bd = umma_smem_desc(sbz + k * 2048, 16384, 1024, SWZ_128);
            umma_bf16(tmem_base + S_COL + st * 128, ad, bd, idesc_b, 1u);
          }
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[kst]);
        if (j == nkv - 1) {
          umma_commit(&q_empty[it & 1]);                   // Q slot reusable once the item's last QK^T retires
          if (resident && (it + 1 == my_items || combo_of(it + 1) != prev_combo)) umma_commit(bias_empty);
        }
      }
      __syncwarp();
    };
    // can S(g) be issued without blocking?  (Q / bias of a new item and the K/V stage have landed; the S buffer itself is
    // free by construction: S(g) is issued after P V(g-2), and the tensor core executes in issue order)
    auto s_ready = [&](int g) {
      const int it = g / nkv, j = g - it * nkv;
      if (j == 0) {
        if (!mbar_test(&q_full[it & 1], (it >> 1) & 1)) return false;
        if (resident && combo_of(it) != prev_combo && !mbar_test(bias_full, nc & 1)) return false;
      }
      return mbar_test(&k_full[g % nst], (g / nst) & 1);
    };
    if (total_blocks > 0) issue_s(0);
    for (int g = 0; g < total_blocks; ++g) {
      const int st = g & 1;
      // S(g+1) goes out as early as its operands allow, but P V(g) never queues behind a K/V load that is still in flight
      bool s_issued = g + 1 >= total_blocks;
      while (!mbar_test(&p_full[st], (g >> 1) & 1)) {
        if (!s_issued && s_ready(g + 1)) { issue_s(g + 1); s_issued = true; }
      }
      mbar_wait(&v_full[g % nst], (g / nst) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sv = smem_u32(smem + L::STAGE_OFF + (g % nst) * stage_stride + L::K_BYTES);
        const bool first = (g % nkv) == 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A = P [128 rows x 16 keys] = 8 packed columns of the S / P buffer; B = V [key][dh], MN-major: 16 keys = 2 atoms
          const uint64_t bd = umma_smem_desc(sv + k * 2 * SBO, 16, SBO, SWZ);
          umma_bf16_ts(tmem_base + O_COL, tmem_base + S_COL + st * 128 + k * 8, bd, idesc_o, (!first || k != 0) ? 1u : 0u);
        }
        umma_commit(&v_empty[g % nst]);
        umma_commit(pv_done);
      }
      __syncwarp();
      if (!s_issued) issue_s(g + 1);
    }
  } else if (warp == 10) {
    // ================================ key-mask warp ================================
    // stages, one block ahead of the softmax warps, the additive key term of every 128-key block:
    // 0 = usable key, -inf = masked key or tile padding beyond n (hides the mask's global-load latency)
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
    // ================================ softmax + epilogue (warps 2..9) ==============
    const int q = warp & 3;               // TMEM lane quarter
    const int hk = (warp - 2) >> 2;       // which 64-key half of every 128-key block this thread owns
    const int r = q * 32 + lane;          // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const float NEG_INF = -__int_as_float(0x7f800000);
    constexpr bool O_OWNER_ALL = (DH == 64);     // DH=64: each half owns 32 O columns; DH=32: half 0 owns all 32
    const bool o_owner = O_OWNER_ALL || hk == 0;
    const uint32_t o_col = O_COL + (O_OWNER_ALL ? hk * 32 : 0);

    for (int it = 0; it < my_items; ++it) {
    int qb, h, b;
    decode(it, qb, h, b);
    bool q_valid = true;

    float m_run = NEG_INF, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int g = it * nkv + j;
      const int st = g & 1;
      mbar_wait(&kb_full[st], (g >> 1) & 1);      // key terms (and, at j == 0, query-mask bytes) staged by the key-mask warp
      if (j == 0) q_valid = qvbuf[(it & 1) * 128 + r] != 0;
      mbar_wait(&s_full[st], (g >> 1) & 1);
      tc_fence_after();

      float s[64];
      {
        uint32_t u0[32], u1[32];
        tmem_ld32(tmem_base + S_COL + st * 128 + hk * 64 + lane_sel, u0);
        tmem_ld32(tmem_base + S_COL + st * 128 + hk * 64 + 32 + lane_sel, u1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = __uint_as_float(u0[i]);
          s[32 + i] = __uint_as_float(u1[i]);
        }
      }

      // logits (log2 domain) = s (pair bias included by the tensor core) + keyterm
      const float* kbs = keyb + st * 128 + hk * 64;
      const bool keys_masked = kflag[st] != 0u;      // block-uniform: most blocks have every key usable
      float mx0 = NEG_INF, mx1 = NEG_INF;
      if (keys_masked) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 k0 = *reinterpret_cast<const float4*>(kbs + c * 8);
          const float4 k1 = *reinterpret_cast<const float4*>(kbs + c * 8 + 4);
          s[c * 8 + 0] += k0.x; s[c * 8 + 1] += k0.y; s[c * 8 + 2] += k0.z; s[c * 8 + 3] += k0.w;
          s[c * 8 + 4] += k1.x; s[c * 8 + 5] += k1.y; s[c * 8 + 6] += k1.z; s[c * 8 + 7] += k1.w;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&kb_empty[st]);
      if (!q_valid) {                        // rare: masked query row -> uniform over the n real keys
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] = (j * 128 + hk * 64 + k) < p.n ? 0.f : NEG_INF;
      }
#pragma unroll
      for (int k = 0; k < 64; k += 2) {
        mx0 = fmaxf(mx0, s[k]);
        mx1 = fmaxf(mx1, s[k + 1]);
      }
      mxbuf[hk * 128 + r] = fmaxf(mx0, mx1);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");     // only the two warps sharing these 32 rows
      float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), mxbuf[(hk ^ 1) * 128 + r]));
      // Lazy rescaling: the running maximum only has to be SOME upper bound up to a factor that neither overflows bf16 P nor
      // the fp32 sums.  While the block maximum exceeds it by <= 2^8 (log2 domain) we keep the stale one and skip the
      // TMEM round trip of O; the decision is made per warp (both warps sharing these rows see identical values).
      bool rescale = true;
      if (j > 0) {
        rescale = __any_sync(0xffffffffu, m_new > m_run + 8.0f);
        if (!rescale) m_new = m_run;
      }
      const float m_use = (m_new == NEG_INF) ? 0.f : m_new;
      const float corr = fast_exp2(m_run - m_use);     // m_run = -inf -> 0
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int k = 0; k < 64; k += 2) {
        const float e0 = fast_exp2(s[k] - m_use);
        const float e1 = fast_exp2(s[k + 1] - m_use);
        s[k] = e0; s[k + 1] = e1;
        ls0 += e0; ls1 += e1;
      }
      l_run = l_run * corr + (ls0 + ls1);
      m_run = m_new;

      // the previous P V must have retired before O is rescaled (rare: the running maximum moved by more than 2^8)
      if (j > 0 && rescale) {
        mbar_wait(pv_done, (g - 1) & 1);
        tc_fence_after();
        if (o_owner) {
          uint32_t o[32];
          tmem_ld32(tmem_base + o_col + lane_sel, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
          tmem_st32(tmem_base + o_col + lane_sel, o);
        }
      }
      // P_j -> tensor memory: this thread's 64 keys become 32 packed columns [hk*32, hk*32+32) of the S buffer.  Those
      // columns held logits of the hk = 0 half, which its owner loaded before the row-max barrier above.
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack_bf16x2(s[2 * c], s[2 * c + 1]);
        tmem_st32(tmem_base + S_COL + st * 128 + hk * 32 + lane_sel, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
      // the previous item's output tile has had a whole key block of time to drain: hand its gate slot back to the producer
      if (j == 0 && it > 0 && hk == 0 && lane == 0) {
        tma_store_wait_read<0>();
        mbar_arrive(&g_empty[(it - 1) & 1]);
      }
    }

    // ---- epilogue: O / l * gate -> out ----
    lbuf[hk * 128 + r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    const float inv_l = 1.0f / (l_run + lbuf[(hk ^ 1) * 128 + r]);
    mbar_wait(pv_done, (it * nkv + nkv - 1) & 1);
    tc_fence_after();
    mbar_wait(&g_full[it & 1], (it >> 1) & 1);
    uint8_t* gt = smem + L::G_OFF + (it & 1) * L::Q_BYTES;
    if (o_owner) {
      uint32_t o[32];
      tmem_ld32(tmem_base + o_col + lane_sel, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // gate tile rows are DH*2 bytes with the TMA swizzle of Q (128B for DH = 64, 64B for DH = 32); the gated output
        // replaces the gate values it was computed from, in the same (swizzled) place
        const uint32_t goff = (DH == 64) ? swz128_off(r, hk * 4 + i) : (r * 64u + ((static_cast<uint32_t>(i) ^ ((r >> 1) & 3u)) << 4));
        const uint4 gq = *reinterpret_cast<const uint4*>(gt + goff);
        const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
        uint32_t ow[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float a = __uint_as_float(o[8 * i + 2 * t]) * inv_l * bf16lo_to_f32(gw[t]);
          const float bb = __uint_as_float(o[8 * i + 2 * t + 1]) * inv_l * bf16hi_to_f32(gw[t]);
          ow[t] = pack_bf16x2(a, bb);
        }
        *reinterpret_cast<uint4*>(gt + goff) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
    tc_fence_before();            // O has been read: order it before the next item's first P V
    fence_proxy_async_smem();
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");     // both column halves of these 32 rows are in the tile
    if (hk == 0 && lane == 0) {
      // rows beyond n are clipped by the tensor map
      tma_store_4d(&tmO, gt + q * 32 * ROWB, 0, qb * 128 + q * 32, h, b);
      tma_store_commit();
    }
    }  // work items
    if (hk == 0 && lane == 0) tma_store_wait_read<0>();   // shared memory must outlive the last output store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace af2
