// Host launcher of proj_tc_kernel (included inside api.cu's anonymous namespace, after make_tmap / ProfScope).

struct ProjOut {
  int ntiles;        // accumulator tiles (256 columns) of the segment inside w_cat
  int kind;          // EpiKind
  int out_cols;      // valid output columns
  void* out;
  long long ld;      // token layout: row stride (elements); channel layout: channel stride (elements)
};
struct ProjCall {
  int a_mode;
  const float* x; long long T; int d; long long src_cs;
  const float* gamma; const float* beta;
  const void* gate_cm; long long gate_cs;
  const float* scale; float scale_const;
  const float* wb; void* bias_out; int heads; long long bias_hs; int n_inner, pitch;
  const void* w_cat;
  const void* w_ext;                      // bf16 [tiles * 256][16]: columns 0 / 1 = hi / lo split of the fp32 bias, rest 0
  int w_rows;                             // rows of w_cat (0: every segment padded to whole 256-row tiles)
  const unsigned char* rowmask;
  float* resid; long long ld_resid;       // EK_RESID_F32 segment: residual source (== out for in-place)
  int nseg; ProjOut seg[PROJ_MAX_SEG];
};

// 2: CTA-pair kernel (default); 1: single-CTA variant; 0: fused kernel disabled (legacy LayerNorm + GEMM launches)
int g_proj_ctas = 2;

// can this LN -> Linear cluster run on the fused kernel?
bool proj_dim_ok(int d) { return d % 64 == 0 && d >= 128 && d <= 256; }
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int CTAS, int AMODE, int KINDS>
int launch_proj_inst(const CUtensorMap& tb, const CUtensorMap* tc, const CUtensorMap& tr, const CUtensorMap& tx, ProjParams& p,
                     double flops, double bytes, cudaStream_t s) {
  using L = ProjSmem<CTAS>;
  static bool configured = false;
  static int max_clusters = 0;
  auto kern = proj_tc_kernel<CTAS, AMODE, KINDS>;
  if (!configured) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    if (CTAS == 2) {
      cudaLaunchConfig_t qc;
      memset(&qc, 0, sizeof(qc));
      qc.gridDim = dim3(sm_count(), 1, 1); qc.blockDim = dim3(PROJ_THREADS, 1, 1); qc.dynamicSmemBytes = L::TOTAL;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kern, &qc) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = sm_count() / 2; }
      max_clusters = nc < sm_count() / 2 ? nc : sm_count() / 2;
    } else {
      max_clusters = sm_count();
    }
    configured = true;
  }
  const int m_units = (p.m_tiles + CTAS - 1) / CTAS;
  if (m_units <= 0) return AF2_OK;
  // column split: an item costs max(MMA time of its tiles, A-producer time) because A is double buffered; the producer
  // costs about as much as PROD_TILES column tiles, so splitting columns only pays when there are too few row units
  const double PROD_TILES = 4.0;
  int best = 1; double best_cost = 1e30;
  for (int ns = 1; ns <= p.n_tiles_total && ns <= 8; ++ns) {
    const long long items = (long long)m_units * ns;
    const long long waves = (items + max_clusters - 1) / max_clusters;
    const double tiles = (double)((p.n_tiles_total + ns - 1) / ns);
    const double cost = (double)waves * (tiles > PROD_TILES ? tiles : PROD_TILES) + 0.5 * PROD_TILES;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ns; }
  }
  p.nsplit = best;
  const long long items = (long long)m_units * p.nsplit;
  const int clusters = (int)(items < max_clusters ? items : max_clusters);
  ProfScope ps(s, KC_GEMM_LINEAR, flops, bytes);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * CTAS, 1, 1); cfg.blockDim = dim3(PROJ_THREADS, 1, 1); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tb, tc[0], tc[1], tc[2], tr, tx, p));
  return AF2_OK;
}

int launch_proj(const ProjCall& c, cudaStream_t s) {
  if (c.T <= 0) return AF2_OK;
  if (!proj_dim_ok(c.d)) return fail(AF2_ERR_BAD_ARG, "proj: dim %d unsupported by the fused kernel", c.d);
  const int ctas = g_proj_ctas == 1 ? 1 : 2;
  ProjParams p;
  memset(&p, 0, sizeof(p));
  p.x = c.x; p.T = c.T; p.d = c.d; p.inv_d = 1.0f / (float)c.d; p.src_cs = c.src_cs; p.gamma = c.gamma; p.beta = c.beta; p.eps = 1e-5f;
  p.gate_cm = static_cast<const __nv_bfloat16*>(c.gate_cm); p.gate_cs = c.gate_cs; p.scale = c.scale; p.scale_const = c.scale_const;
  p.wb = c.wb; p.bias_out = static_cast<__nv_bfloat16*>(c.bias_out); p.heads = c.heads; p.bias_hs = c.bias_hs;
  p.n_inner = c.n_inner > 0 ? c.n_inner : 1; p.pitch = c.pitch > 0 ? c.pitch : 1;
  p.rowmask = c.rowmask; p.nseg = c.nseg;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("AF2_PROJ_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  if (!c.w_ext) return fail(AF2_ERR_BAD_ARG, "proj: bias block (w_ext) missing");
  p.m_tiles = (int)((c.T + 127) / 128);
  CUtensorMap tc[3], tr, tb, tx;
  int tile0 = 0;
  bool has_resid = false;
  double flops = 0, bytes = (double)c.T * c.d * 4;
  for (int i = 0; i < c.nseg; ++i) {
    const ProjOut& o = c.seg[i];
    p.seg[i].tile0 = tile0; p.seg[i].ntiles = o.ntiles; p.seg[i].kind = o.kind; p.seg[i].out_cols = o.out_cols; p.seg[i].map = i;
    tile0 += o.ntiles;
    const bool f32 = (o.kind == EK_RESID_F32);
    const bool chan = (o.kind == EK_STORE_CH || o.kind == EK_STORE_CH_SIG || o.kind == EK_GATED_CH_SIG);
    const int es = f32 ? 4 : 2;
    if (!aligned16(o.out) || (o.ld * es) % 16) return fail(AF2_ERR_BAD_ARG, "proj: output %d not TMA-describable", i);
    const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (!chan) {
      unsigned long long dc[3] = {(unsigned long long)o.out_cols, (unsigned long long)c.T, 1ull};
      unsigned long long sc[2] = {(unsigned long long)o.ld * es, (unsigned long long)o.ld * c.T * es};
      unsigned bc[3] = {(unsigned)(f32 ? 16 : 32), 128, 1};          // 64-byte rows: 8 KB staging buffers
      AF2_TRY(make_tmap(&tc[i], o.out, 3, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_64B, dt));
      if (f32) {
        if (!aligned16(c.resid) || (c.ld_resid * 4) % 16) return fail(AF2_ERR_BAD_ARG, "proj: residual not TMA-describable");
        unsigned long long sr[2] = {(unsigned long long)c.ld_resid * 4, (unsigned long long)c.ld_resid * c.T * 4};
        AF2_TRY(make_tmap(&tr, c.resid, 3, dc, sr, bc, CU_TENSOR_MAP_SWIZZLE_64B, dt));
        has_resid = true;
      }
    } else {
      unsigned long long dc[3] = {(unsigned long long)c.T, (unsigned long long)o.out_cols, 1ull};
      unsigned long long sc[2] = {(unsigned long long)o.ld * 2, (unsigned long long)o.ld * o.out_cols * 2};
      unsigned bc[3] = {64, 32, 1};
      AF2_TRY(make_tmap(&tc[i], o.out, 3, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_128B, dt));
    }
    const int W = (o.kind == EK_GATED_TOK_GELU || o.kind == EK_GATED_CH_SIG) ? 2 : 1;
    flops += 2.0 * c.T * (double)o.out_cols * W * c.d;
    bytes += (double)c.T * o.out_cols * es * (f32 ? 2 : 1);
  }
  for (int i = c.nseg; i < 3; ++i) tc[i] = tc[0];
  if (!has_resid) tr = tc[0];
  p.n_tiles_total = tile0;

  bytes += (double)tile0 * 256 * c.d * 2;
  if (c.a_mode == 1) bytes += (double)c.T * c.d * 2;
  {
    unsigned long long db[2] = {(unsigned long long)c.d, (unsigned long long)(c.w_rows > 0 ? c.w_rows : tile0 * 256)};
    unsigned long long sb[1] = {(unsigned long long)c.d * 2};
    unsigned bb[2] = {64, (unsigned)(256 / ctas)};
    AF2_TRY(make_tmap(&tb, c.w_cat, 2, db, sb, bb, CU_TENSOR_MAP_SWIZZLE_128B));
    unsigned long long dx[2] = {16ull, (unsigned long long)tile0 * 256};
    unsigned long long sx[1] = {32ull};
    unsigned bx[2] = {16, (unsigned)(256 / ctas)};
    AF2_TRY(make_tmap(&tx, c.w_ext, 2, dx, sx, bx, CU_TENSOR_MAP_SWIZZLE_32B));
  }
  int kinds = 0;
  for (int i = 0; i < c.nseg; ++i) kinds |= KBIT(c.seg[i].kind);
#define AF2_PROJ_DISPATCH(AM, KS)                                                               \
  if (c.a_mode == AM && kinds == KS) {                                                          \
    if (ctas == 2) return launch_proj_inst<2, AM, KS>(tb, tc, tr, tx, p, flops, bytes, s);      \
    return launch_proj_inst<1, AM, KS>(tb, tc, tr, tx, p, flops, bytes, s);                     \
  }
  AF2_PROJ_DISPATCH(0, PK_ATTN)
  AF2_PROJ_DISPATCH(0, PK_TRI)
  AF2_PROJ_DISPATCH(0, PK_TRI_CH)
  AF2_PROJ_DISPATCH(0, PK_FF)
  AF2_PROJ_DISPATCH(0, PK_OUTER)
  AF2_PROJ_DISPATCH(1, PK_TAIL)
  AF2_PROJ_DISPATCH(2, PK_TAIL)
#undef AF2_PROJ_DISPATCH
  return fail(AF2_ERR_BAD_ARG, "proj: no kernel instantiation for producer mode %d / epilogue set 0x%x", c.a_mode, kinds);
}
