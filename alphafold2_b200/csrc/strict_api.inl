// Host side of the STRICT precision mode (included at the end of api.cu).  Every module is the same sequence of sub-ops as
// the default path, but activations stay fp32 between kernels and every GEMM runs on split-bf16 operands (3 tensor-core
// passes, GemmParams::nseg): results match the reference's fp32 path inside rtol 1e-3 / atol 1e-4 (tests/test_gpu_strict.py).
namespace {

inline int a8(long long v) { return (int)align_up(v, 8); }

// C[b] = A[b] B[b]^T (K-major) or A[b]^T-style MN-major contraction on split operands; epilogue as configured by the caller
GemmCall split_call(const void* A, long long lda, long long a_half, long long a_batch, const void* Bm, long long ldb,
                    long long b_half, long long b_batch, int M, int N, int K, int batch, bool mn_major) {
  GemmCall c;
  memset(&c, 0, sizeof(c));
  c.A = A; c.lda = lda; c.a_half = a_half; c.a_batch = a_batch;
  c.Bm = Bm; c.ldb = ldb; c.b_half = b_half; c.b_batch = b_batch;
  c.M = M; c.N = N; c.K = K; c.batch = batch; c.mn_major = mn_major; c.bn = pick_bn(N); c.nseg = SPL_NSEG;
  return c;
}

// out[T][N] fp32 = A_split[T][2][Pk] x W_split[N][2][Pk]^T + bias
int strict_linear_f32(const __nv_bfloat16* A, const void* W, const float* bias, float* out, long long T, int N, int K, cudaStream_t s) {
  const int Pk = a8(K);
  GemmCall c = split_call(A, (long long)SPL * Pk, Pk, 0, W, (long long)SPL * Pk, Pk, 0, (int)T, N, K, 1, false);
  c.mode = EPI_STORE_F32; c.layout = LAYOUT_TOKEN; c.out = out; c.ld_out = N; c.bias = bias;
  return launch_gemm(c, s);
}
// x[T][N] fp32 += A_split[T][2][Pk] x W_split[N][2][Pk]^T + bias   (output projections with residual)
int strict_linear_resid(const __nv_bfloat16* A, const void* W, const float* bias, float* x, long long T, int N, int K, cudaStream_t s) {
  const int Pk = a8(K);
  GemmCall c = split_call(A, (long long)SPL * Pk, Pk, 0, W, (long long)SPL * Pk, Pk, 0, (int)T, N, K, 1, false);
  c.mode = EPI_RESID_F32; c.out = x; c.ld_out = N; c.bias = bias; c.resid = x; c.ld_resid = N;
  return launch_gemm(c, s);
}

int strict_ln_split(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y, long long T, int d, cudaStream_t s) {
  if (T <= 0) return AF2_OK;
  const long long need = (T + 7) / 8, cap = (long long)sm_count() * 16;
  ProfScope ps(s, KC_LAYERNORM, 0.0, (double)T * d * 8.0);
  strict_ln_split_kernel<<<(int)(need < cap ? need : cap), 256, 0, s>>>(x, gamma, beta, y, T, d, a8(d), 1e-5f);
  CUDA_OK(cudaGetLastError());
  return AF2_OK;
}

int strict_chan_to_token(const StrictC2TParams& p, cudaStream_t s) {
  const size_t smem = (size_t)p.d * 33 * sizeof(float);
  static size_t configured[MAX_DEVICES] = {0};
  if (smem > 48 * 1024 && smem > configured[cur_dev()]) {
    CUDA_OK(cudaFuncSetAttribute(strict_chan_to_token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[cur_dev()] = smem;
  }
  dim3 grid((p.n + 31) / 32, p.rows);
  ProfScope ps(s, KC_CHAN2TOK, 0.0, (double)p.rows * p.n * p.d * 8.0);
  strict_chan_to_token_kernel<<<grid, 256, smem, s>>>(p);
  CUDA_OK(cudaGetLastError());
  return AF2_OK;
}

int strict_tok2chan(const float* src, long long ld, int val_off, int gate_off, const unsigned char* mask, __nv_bfloat16* out, int C,
                    int rows, int inner, long long cs, long long hs, long long rs, cudaStream_t s) {
  if (rows <= 0 || inner <= 0) return AF2_OK;
  dim3 grid((unsigned)(rows * ((inner + 31) / 32)), (unsigned)((C + 31) / 32));
  ProfScope ps(s, KC_MISC, 0.0, 0.0);
  strict_tok2chan_split_kernel<<<grid, 256, 0, s>>>(src, ld, val_off, gate_off, mask, out, C, inner, cs, hs, rs);
  CUDA_OK(cudaGetLastError());
  return AF2_OK;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------ utilities (tests)
// y split [rows][SPL][align8(K)] of x fp32 [rows][K]
int af2_split_bf16(const float* x, void* y, long long rows, int K, af2_stream_t stream) {
  return strict_ln_split(x, nullptr, nullptr, static_cast<__nv_bfloat16*>(y), rows, K, static_cast<cudaStream_t>(stream));
}

// C[b][m][n] fp32 = sum_k A[b][m][k] B[b][n][k] on split operands A [batch][M][SPL][P], B [batch][N][SPL][P], P = align8(K)
int af2_gemm_split_f32(const void* A, const void* Bm, float* C, long long ldc, int M, int N, int K, int batch, af2_stream_t stream) {
  const int P = a8(K);
  GemmCall c = split_call(A, (long long)SPL * P, P, (long long)M * SPL * P, Bm, (long long)SPL * P, P, (long long)N * SPL * P, M, N, K, batch, false);
  c.mode = EPI_STORE_F32; c.layout = LAYOUT_TOKEN; c.out = C; c.ld_out = ldc; c.out_batch = (long long)M * ldc;
  return launch_gemm(c, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------ FeedForward
long long af2_feed_forward_strict_workspace(long long tokens, int d, int hidden) {
  return align_up(tokens * SPL * a8(d) * 2, 256) + align_up(tokens * 2 * hidden * 4, 256) + align_up(tokens * SPL * a8(hidden) * 2, 256) + 1024;
}

int af2_feed_forward_strict(const af2_ff_weights_strict* w, float* x, long long tokens, int d, int hidden, void* workspace,
                            long long workspace_bytes, af2_stream_t stream) {
  NvtxRange nvtx_("af2_feed_forward_strict");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!w || !x) return fail(AF2_ERR_BAD_ARG, "feed_forward_strict: null argument");
  if (d % 4 || hidden % 4) return fail(AF2_ERR_BAD_ARG, "feed_forward_strict: d=%d and hidden=%d must be multiples of 4", d, hidden);
  if (tokens > 0x7fffffffLL) return fail(AF2_ERR_BAD_ARG, "feed_forward_strict: too many tokens");
  Arena ar(workspace, workspace_bytes);
  __nv_bfloat16* xs = ar.take<__nv_bfloat16>(tokens * SPL * a8(d));
  float* h = ar.take<float>(tokens * 2 * hidden);
  __nv_bfloat16* hs = ar.take<__nv_bfloat16>(tokens * SPL * a8(hidden));
  if (!ar.ok) return fail(AF2_ERR_WORKSPACE, "feed_forward_strict: workspace too small");
  AF2_TRY(strict_ln_split(x, w->ln_gamma, w->ln_beta, xs, tokens, d, s));
  AF2_TRY(strict_linear_f32(xs, w->w1, w->b1, h, tokens, 2 * hidden, d, s));
  {
    ProfScope ps(s, KC_MISC, 0.0, 0.0);
    strict_geglu_split_kernel<<<ew_grid(tokens * a8(hidden)), 256, 0, s>>>(h, hs, tokens, hidden, a8(hidden));
    CUDA_OK(cudaGetLastError());
  }
  return strict_linear_resid(hs, w->w2, w->b2, x, tokens, d, hidden, s);
}

// ------------------------------------------------------------------------------------------------ AxialAttention
static long long strict_attn_chunk(int nb, int heads, int n) {      // folded batch elements per pass: logits <= ~1 GiB
  const long long per = (long long)heads * n * align_up(n, 4) * 4;
  long long c = (1LL << 30) / (per > 0 ? per : 1);
  if (c < 1) c = 1;
  return c < nb ? c : nb;
}

long long af2_axial_attention_strict_workspace(int B, int h, int wdim, int d, int heads, int dim_head, int row_attn) {
  const long long T = (long long)B * h * wdim, I = (long long)heads * dim_head;
  const int n = row_attn ? wdim : h, nb = row_attn ? h : wdim;
  const long long ch = strict_attn_chunk(nb, heads, n);
  const int Pd = a8(dim_head), Pn = a8(n);
  return align_up(T * SPL * a8(d) * 2, 256) + align_up(T * 4 * I * 4, 256) + align_up(T * SPL * a8(I) * 2, 256) +
         align_up((long long)heads * n * n * 4, 256) +
         2 * align_up(ch * heads * n * SPL * Pd * 2, 256) + align_up(ch * heads * dim_head * SPL * Pn * 2, 256) +
         align_up(ch * heads * n * align_up(n, 4) * 4, 256) + align_up(ch * heads * n * SPL * Pn * 2, 256) +
         align_up(ch * heads * n * (long long)dim_head * 4, 256) + 4096;
}

int af2_axial_attention_strict(const af2_attn_weights_strict* w, float* x, const float* edges, const unsigned char* mask, int B,
                               int h, int wdim, int d, int heads, int dim_head, int row_attn, int flags, void* workspace,
                               long long workspace_bytes, af2_stream_t stream) {
  NvtxRange nvtx_("af2_axial_attention_strict");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!w || !x) return fail(AF2_ERR_BAD_ARG, "axial_attention_strict: null argument");
  if (dim_head % 4 || d % 4) return fail(AF2_ERR_BAD_ARG, "axial_attention_strict: dim %d / dim_head %d must be multiples of 4", d, dim_head);
  const long long T = (long long)B * h * wdim, I = (long long)heads * dim_head;
  if (T > 0x7fffffffLL) return fail(AF2_ERR_BAD_ARG, "axial_attention_strict: too many tokens");
  const int n = row_attn ? wdim : h, nb = row_attn ? h : wdim;
  const long long ch = strict_attn_chunk(nb, heads, n);
  const int Pd = a8(dim_head), Pn = a8(n), Pi = a8(I);
  const long long lds = align_up(n, 4);
  const bool has_bias = edges != nullptr && w->w_edge != nullptr;
  Arena ar(workspace, workspace_bytes);
  __nv_bfloat16* xs = ar.take<__nv_bfloat16>(T * SPL * a8(d));
  float* p1 = ar.take<float>(T * 4 * I);
  __nv_bfloat16* og = ar.take<__nv_bfloat16>(T * SPL * Pi);
  float* bias = ar.take<float>((long long)heads * n * n);
  __nv_bfloat16* Q = ar.take<__nv_bfloat16>(ch * heads * n * SPL * Pd);
  __nv_bfloat16* K = ar.take<__nv_bfloat16>(ch * heads * n * SPL * Pd);
  __nv_bfloat16* Vt = ar.take<__nv_bfloat16>(ch * heads * dim_head * SPL * Pn);
  float* S = ar.take<float>(ch * heads * n * lds);
  __nv_bfloat16* P = ar.take<__nv_bfloat16>(ch * heads * n * SPL * Pn);
  float* O = ar.take<float>(ch * heads * n * (long long)dim_head);
  if (!ar.ok) return fail(AF2_ERR_WORKSPACE, "axial_attention_strict: workspace too small");

  AF2_TRY(strict_ln_split(x, w->ln_gamma, w->ln_beta, xs, T, d, s));
  AF2_TRY(strict_linear_f32(xs, w->w_qkvg, w->b_qkvg, p1, T, (int)(4 * I), d, s));
  const long long tok_sb = row_attn ? wdim : 1, tok_si = row_attn ? 1 : wdim;
  for (int b = 0; b < B; ++b) {
    const long long t0 = (long long)b * h * wdim;
    if (flags & 1) {  // tied queries (alphafold2.py:142-151): q <- mean over the folded batch, in place in the fp32 projection
      ProfScope ps(s, KC_MISC, 0.0, 0.0);
      tie_queries_kernel<float><<<ew_grid((long long)n * I), 256, 0, s>>>(p1 + t0 * 4 * I, 4 * I, (int)I, n, nb, tok_sb, tok_si);
      CUDA_OK(cudaGetLastError());
    }
    if (has_bias) {   // raw (un-normalised) edges of this batch element, alphafold2.py:245-247 (quirks Q4 / Q5)
      const long long Te = (long long)n * n;
      ProfScope ps(s, KC_LAYERNORM, 0.0, (double)Te * d * 4);
      const long long need = (Te + 7) / 8, cap = (long long)sm_count() * 16;
      strict_pair_bias_kernel<<<(int)(need < cap ? need : cap), 256, 0, s>>>(edges + (long long)b * Te * d, w->w_edge, bias, Te, d, heads);
      CUDA_OK(cudaGetLastError());
    }
    for (long long b0 = 0; b0 < nb; b0 += ch) {
      const int nbc = (int)((nb - b0) < ch ? (nb - b0) : ch);
      const int bh = nbc * heads;
      {
        ProfScope ps(s, KC_MISC, 0.0, 0.0);
        strict_qkv_split_kernel<<<ew_grid((long long)bh * n * Pd), 256, 0, s>>>(p1 + t0 * 4 * I, Q, K, Vt, (int)b0, nbc, heads, n, dim_head, Pd, Pn, tok_sb, tok_si);
        CUDA_OK(cudaGetLastError());
      }
      {   // logits S = Q K^T per (b', head)
        GemmCall c = split_call(Q, (long long)SPL * Pd, Pd, (long long)n * SPL * Pd, K, (long long)SPL * Pd, Pd, (long long)n * SPL * Pd, n, n, dim_head, bh, false);
        c.mode = EPI_STORE_F32; c.layout = LAYOUT_TOKEN; c.out = S; c.ld_out = lds; c.out_batch = (long long)n * lds;
        AF2_TRY(launch_gemm(c, s));
      }
      {
        ProfScope ps(s, KC_ATTENTION, 0.0, 0.0);
        const long long rows = (long long)bh * n, need = (rows + 7) / 8, cap = (long long)sm_count() * 16;
        strict_softmax_split_kernel<<<(int)(need < cap ? need : cap), 256, 0, s>>>(S, lds, has_bias ? bias : nullptr, mask ? mask + t0 : nullptr, P,
                                                                               (int)b0, nbc, heads, n, Pn, tok_sb, tok_si);
        CUDA_OK(cudaGetLastError());
      }
      {   // O = P V per (b', head): A = P [n][keys], B = V^T [dh][keys]
        GemmCall c = split_call(P, (long long)SPL * Pn, Pn, (long long)n * SPL * Pn, Vt, (long long)SPL * Pn, Pn, (long long)dim_head * SPL * Pn, n, dim_head, n, bh, false);
        c.mode = EPI_STORE_F32; c.layout = LAYOUT_TOKEN; c.out = O; c.ld_out = dim_head; c.out_batch = (long long)n * dim_head;
        AF2_TRY(launch_gemm(c, s));
      }
      {
        ProfScope ps(s, KC_MISC, 0.0, 0.0);
        strict_gate_split_kernel<<<ew_grid((long long)nbc * n * Pi), 256, 0, s>>>(O, dim_head, p1 + t0 * 4 * I, og + t0 * SPL * Pi, (int)b0, nbc, heads, n, dim_head, Pi, tok_sb, tok_si);
        CUDA_OK(cudaGetLastError());
      }
    }
  }
  return strict_linear_resid(og, w->w_out, w->b_out, x, T, d, (int)I, s);
}

// ------------------------------------------------------------------------------------------------ TriangleMultiplicativeModule
long long af2_triangle_multiply_strict_workspace(int B, int N, int d) {
  const long long T = (long long)B * N * N;
  const int P8 = a8(N);
  return align_up(T * SPL * a8(d) * 2, 256) + align_up(T * 5 * d * 4, 256) + 2 * align_up((long long)d * N * SPL * P8 * 2, 256) +
         align_up((long long)d * N * align_up(N, 4) * 4, 256) + align_up((long long)N * N * SPL * a8(d) * 2, 256) + 2048;
}

int af2_triangle_multiply_strict(const af2_trimul_weights_strict* w, float* x, const unsigned char* mask, int B, int N, int d,
                                 int ingoing, void* workspace, long long workspace_bytes, af2_stream_t stream) {
  NvtxRange nvtx_("af2_triangle_multiply_strict");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!w || !x) return fail(AF2_ERR_BAD_ARG, "triangle_multiply_strict: null argument");
  if (d % 4) return fail(AF2_ERR_BAD_ARG, "triangle_multiply_strict: dim %d must be a multiple of 4", d);
  const long long T = (long long)B * N * N, Tb = (long long)N * N;
  if (T > 0x7fffffffLL) return fail(AF2_ERR_BAD_ARG, "triangle_multiply_strict: too many tokens");
  const int P8 = a8(N), np4 = (int)align_up(N, 4), Pd = a8(d);
  const long long cs = (long long)N * SPL * P8;      // channel stride of the split operands
  Arena ar(workspace, workspace_bytes);
  __nv_bfloat16* xs = ar.take<__nv_bfloat16>(T * SPL * Pd);
  float* p5 = ar.take<float>(T * 5 * d);             // left | right | left_gate | right_gate | out_gate (pre-activation)
  __nv_bfloat16* Lc = ar.take<__nv_bfloat16>(d * cs);
  __nv_bfloat16* Rc = ar.take<__nv_bfloat16>(d * cs);
  float* Oc = ar.take<float>((long long)d * N * np4);
  __nv_bfloat16* tn = ar.take<__nv_bfloat16>(Tb * SPL * Pd);
  if (!ar.ok) return fail(AF2_ERR_WORKSPACE, "triangle_multiply_strict: workspace too small");
  AF2_TRY(strict_ln_split(x, w->ln_gamma, w->ln_beta, xs, T, d, s));
  AF2_TRY(strict_linear_f32(xs, w->w5, w->b5, p5, T, 5 * d, d, s));
  for (int b = 0; b < B; ++b) {
    const float* pb = p5 + (long long)b * Tb * 5 * d;
    const unsigned char* mb = mask ? mask + (long long)b * Tb : nullptr;
    // pad columns of the operand planes are read by TMA as K / MN padding only when P8 != N: keep them zero
    if (P8 != N) {
      CUDA_OK(cudaMemsetAsync(Lc, 0, (size_t)d * cs * 2, s));
      CUDA_OK(cudaMemsetAsync(Rc, 0, (size_t)d * cs * 2, s));
    }
    GemmCall c;
    if (!ingoing) {   // O_c = L_c R_c^T (alphafold2.py:285): K-major operands [c][i][SPL][P8]
      AF2_TRY(strict_tok2chan(pb, 5LL * d, 0, 2 * d, mb, Lc, d, N, N, cs, P8, (long long)SPL * P8, s));
      AF2_TRY(strict_tok2chan(pb, 5LL * d, d, 3 * d, mb, Rc, d, N, N, cs, P8, (long long)SPL * P8, s));
      c = split_call(Lc, (long long)SPL * P8, P8, cs, Rc, (long long)SPL * P8, P8, cs, N, N, N, d, false);
    } else {          // O_c[i][j] = sum_k R_c[k][i] L_c[k][j] (alphafold2.py:287, quirk Q6): MN-major operands [c][SPL][k][P8]
      AF2_TRY(strict_tok2chan(pb, 5LL * d, 0, 2 * d, mb, Lc, d, N, N, cs, (long long)N * P8, P8, s));
      AF2_TRY(strict_tok2chan(pb, 5LL * d, d, 3 * d, mb, Rc, d, N, N, cs, (long long)N * P8, P8, s));
      c = split_call(Rc, P8, (long long)N * P8, cs, Lc, P8, (long long)N * P8, cs, N, N, N, d, true);
    }
    c.mode = EPI_STORE_F32; c.layout = LAYOUT_TOKEN; c.out = Oc; c.ld_out = np4; c.out_batch = (long long)N * np4;
    AF2_TRY(launch_gemm(c, s));
    StrictC2TParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.src = Oc; cp.chan_stride = (long long)N * np4; cp.pitch = np4; cp.rows = N; cp.n = N; cp.d = d; cp.P = Pd; cp.mode = 0;
    cp.gamma = w->on_gamma; cp.beta = w->on_beta; cp.gate_src = pb; cp.gate_ld = 5LL * d; cp.gate_off = 4 * d; cp.eps = 1e-5f; cp.y = tn;
    AF2_TRY(strict_chan_to_token(cp, s));
    AF2_TRY(strict_linear_resid(tn, w->w_out, w->b_out, x + (long long)b * Tb * d, Tb, d, d, s));
  }
  return AF2_OK;
}

// ------------------------------------------------------------------------------------------------ OuterMean
long long af2_outer_mean_strict_workspace(int B, int S, int N, int d) {
  const long long Tm = (long long)B * S * N;
  const int P8 = a8(N);
  return align_up(Tm * SPL * a8(d) * 2, 256) + align_up(Tm * 2 * d * 4, 256) + align_up((long long)2 * d * SPL * S * P8 * 2, 256) +
         align_up((long long)d * N * align_up(N, 4) * 4, 256) + align_up((long long)N * N * SPL * a8(d) * 2, 256) +
         align_up((long long)N * N * 4, 256) + align_up((long long)((S + 31) / 32) * N * 4, 256) + 2048;
}

int af2_outer_mean_strict(const af2_outer_weights_strict* w, float* x, const float* m, const unsigned char* msa_mask, int B, int S,
                          int N, int d, float eps, void* workspace, long long workspace_bytes, af2_stream_t stream) {
  NvtxRange nvtx_("af2_outer_mean_strict");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!w || !x || !m) return fail(AF2_ERR_BAD_ARG, "outer_mean_strict: null argument");
  if (d % 4) return fail(AF2_ERR_BAD_ARG, "outer_mean_strict: dim %d must be a multiple of 4", d);
  const long long Tm = (long long)B * S * N, Tmb = (long long)S * N, Txb = (long long)N * N;
  if (Tm > 0x7fffffffLL || (long long)B * Txb > 0x7fffffffLL) return fail(AF2_ERR_BAD_ARG, "outer_mean_strict: too many tokens");
  const int P8 = a8(N), np4 = (int)align_up(N, 4), Pd = a8(d);
  const long long cs = (long long)SPL * S * P8;      // channel stride: [c][SPL][S][P8]
  Arena ar(workspace, workspace_bytes);
  __nv_bfloat16* ms = ar.take<__nv_bfloat16>(Tm * SPL * Pd);
  float* p2 = ar.take<float>(Tm * 2 * d);             // left | right
  __nv_bfloat16* LRc = ar.take<__nv_bfloat16>(2LL * d * cs);
  float* Oc = ar.take<float>((long long)d * N * np4);
  __nv_bfloat16* tn = ar.take<__nv_bfloat16>(Txb * SPL * Pd);
  float* scale = ar.take<float>(Txb);
  uint32_t* mwords = ar.take<uint32_t>((long long)((S + 31) / 32) * N);
  if (!ar.ok) return fail(AF2_ERR_WORKSPACE, "outer_mean_strict: workspace too small");
  AF2_TRY(strict_ln_split(m, w->ln_gamma, w->ln_beta, ms, Tm, d, s));
  AF2_TRY(strict_linear_f32(ms, w->w_lr, w->b_lr, p2, Tm, 2 * d, d, s));
  for (int b = 0; b < B; ++b) {
    const float* pb = p2 + (long long)b * Tmb * 2 * d;
    const unsigned char* mb = msa_mask ? msa_mask + (long long)b * Tmb : nullptr;
    if (P8 != N) CUDA_OK(cudaMemsetAsync(LRc, 0, (size_t)2 * d * cs * 2, s));
    // channels [0, d) = left, [d, 2d) = right; rows = MSA row s, columns = residue
    AF2_TRY(strict_tok2chan(pb, 2LL * d, 0, -1, mb, LRc, 2 * d, S, N, cs, (long long)S * P8, P8, s));
    GemmCall c = split_call(LRc, P8, (long long)S * P8, cs, LRc + (long long)d * cs, P8, (long long)S * P8, cs, N, N, S, d, true);
    c.mode = EPI_STORE_F32; c.layout = LAYOUT_TOKEN; c.out = Oc; c.ld_out = np4; c.out_batch = (long long)N * np4;
    AF2_TRY(launch_gemm(c, s));
    if (mb) AF2_TRY(launch_outer_scale(mb, scale, mwords, 0, N, S, N, eps, s));
    StrictC2TParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.src = Oc; cp.chan_stride = (long long)N * np4; cp.pitch = np4; cp.rows = N; cp.n = N; cp.d = d; cp.P = Pd; cp.mode = 1;
    cp.scale = mb ? scale : nullptr; cp.scale_const = 1.0f / (float)S; cp.y = tn;
    AF2_TRY(strict_chan_to_token(cp, s));
    AF2_TRY(strict_linear_resid(tn, w->w_out, w->b_out, x + (long long)b * Txb * d, Txb, d, d, s));
  }
  return AF2_OK;
}

}  // extern "C"

// =================================================================================================
// Pre- / post-trunk glue (glue_kernels.cuh)
// =================================================================================================
extern "C" {

long long af2_embed_pair_init_workspace(int B, int n, int d) { return align_up((long long)B * n * d * 4, 256) + align_up((long long)B * n * 2 * d * 4, 256) + 512; }

// alphafold2.py:676-726: x [B][n][n][d], m [B][S][n][d] (fp32) from token ids.  seq_embed [B][n][d], msa_embed [B][S][n][d],
// seq_index [n] (int64) are optional (NULL).
int af2_embed_pair_init(const long long* seq, const long long* msa, const float* token_emb, int vocab, const float* seq_embed,
                        const float* msa_embed, const float* w_pair, const float* b_pair, const float* pos_emb, int max_rel_dist,
                        const long long* seq_index, float* x, float* m, int B, int S, int n, int d, void* workspace,
                        long long workspace_bytes, af2_stream_t stream) {
  NvtxRange nvtx_("af2_embed_pair_init");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!seq || !token_emb || !w_pair || !b_pair || !pos_emb || !x) return fail(AF2_ERR_BAD_ARG, "embed_pair_init: null argument");
  if (d % 4 || d > 1024) return fail(AF2_ERR_BAD_ARG, "embed_pair_init: dim %d must be a multiple of 4 and <= 1024", d);
  Arena ar(workspace, workspace_bytes);
  float* e = ar.take<float>((long long)B * n * d);
  float* lr = ar.take<float>((long long)B * n * 2 * d);
  if (!ar.ok) return fail(AF2_ERR_WORKSPACE, "embed_pair_init: workspace too small");
  {
    ProfScope ps(s, KC_MISC, 0.0, 0.0);
    glue_seq_kernel<<<B * n, 256, d * sizeof(float), s>>>(seq, token_emb, seq_embed, w_pair, b_pair, e, lr, d, vocab);
    CUDA_OK(cudaGetLastError());
  }
  if (msa && m) {
    const long long tokens = (long long)B * S * n;
    ProfScope ps(s, KC_MISC, 0.0, (double)tokens * d * 8);
    glue_msa_init_kernel<<<ew_grid(tokens * (d / 4)), 256, 0, s>>>(msa, token_emb, msa_embed, e, m, tokens, S, n, d, vocab);
    CUDA_OK(cudaGetLastError());
  }
  {
    ProfScope ps(s, KC_MISC, 0.0, (double)B * n * n * d * 4);
    glue_pair_init_kernel<<<ew_grid((long long)B * n * n * (d / 4)), 256, 0, s>>>(lr, pos_emb, seq_index, x, B, n, d, max_rel_dist);
    CUDA_OK(cudaGetLastError());
  }
  return AF2_OK;
}

// alphafold2.py:821-823: out [B][n][n][buckets] = Linear(LayerNorm((x + x^T) / 2))
int af2_distogram_head(const float* x, const float* gamma, const float* beta, const float* w, const float* bias, float* out, int B,
                       int n, int d, int buckets, af2_stream_t stream) {
  NvtxRange nvtx_("af2_distogram_head");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!x || !gamma || !beta || !w || !bias || !out) return fail(AF2_ERR_BAD_ARG, "distogram_head: null argument");
  if (d % 128 || d > 512) return fail(AF2_ERR_BAD_ARG, "distogram_head: dim %d must be a multiple of 128 and <= 512", d);
  const size_t smem = ((size_t)buckets * d + buckets) * sizeof(float);
  if (smem > 200 * 1024) return fail(AF2_ERR_BAD_ARG, "distogram_head: %d buckets x dim %d does not fit in shared memory", buckets, d);
  const long long T = (long long)B * n * n;
  const long long need = (T + 7) / 8;
  const int per_sm = (int)(200 * 1024 / (smem + 1024));
  const long long cap = (long long)sm_count() * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
  const int grid = (int)(need < cap ? need : cap);
  ProfScope ps(s, KC_MISC, 0.0, (double)T * d * 8 + (double)T * buckets * 4);
#define AF2_DISTO(V)                                                                                                  \
  {                                                                                                                   \
    static size_t configured[MAX_DEVICES] = {0};                                                                      \
    if (smem > 48 * 1024 && smem > configured[cur_dev()]) {                                                           \
      CUDA_OK(cudaFuncSetAttribute(glue_distogram_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      configured[cur_dev()] = smem;                                                                                   \
    }                                                                                                                 \
    glue_distogram_kernel<V><<<grid, 256, smem, s>>>(x, gamma, beta, w, bias, out, B, n, buckets, 1e-5f);             \
  }
  switch (d / 32) {
    case 4: AF2_DISTO(4) break;
    case 8: AF2_DISTO(8) break;
    case 12: AF2_DISTO(12) break;
    default: AF2_DISTO(16) break;
  }
#undef AF2_DISTO
  CUDA_OK(cudaGetLastError());
  return AF2_OK;
}

}  // extern "C"

// =================================================================================================
// L2 residency of the fp32 pair stream (experiment knob AF2_L2_PERSIST): an access-policy window on `stream`
// marks [ptr, ptr + bytes) as persisting in L2 for every kernel launched on it afterwards.  ptr == NULL clears the window.
// =================================================================================================
extern "C" int af2_l2_persist(const void* ptr, long long bytes, float hit_ratio, af2_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int dev = cur_dev();
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (!ptr || bytes <= 0) {
    attr.accessPolicyWindow.base_ptr = nullptr;
    attr.accessPolicyWindow.num_bytes = 0;
    attr.accessPolicyWindow.hitRatio = 0.f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    CUDA_OK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
    CUDA_OK(cudaCtxResetPersistingL2Cache());
    return AF2_OK;
  }
  int max_persist = 0, max_window = 0;
  CUDA_OK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  CUDA_OK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  if (max_persist <= 0 || max_window <= 0) return fail(AF2_ERR_CUDA, "l2_persist: device reports no persisting L2 (%d / %d)", max_persist, max_window);
  static long long limit_set[MAX_DEVICES] = {0};
  const long long want = bytes < (long long)max_persist ? bytes : (long long)max_persist;
  if (limit_set[dev] < want) {
    CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want));
    limit_set[dev] = want;
  }
  attr.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  attr.accessPolicyWindow.num_bytes = (size_t)(bytes < (long long)max_window ? bytes : (long long)max_window);
  // if the window is larger than the set-aside, only a matching fraction of it can persist without thrashing
  float hr = hit_ratio > 0.f ? hit_ratio : 1.0f;
  const float fit = (float)want / (float)attr.accessPolicyWindow.num_bytes;
  if (hr > fit) hr = fit;
  attr.accessPolicyWindow.hitRatio = hr;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  CUDA_OK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
  return AF2_OK;
}
