// Host launcher of proj_tc_kernel (included inside api.cu's anonymous namespace, after make_tmap / ProfScope).

struct ProjOut {
  int ntiles;        // accumulator tiles (256 columns) of the segment inside w_cat
  int kind;          // EpiKind
  int out_cols;      // valid output columns
  void* out;
  long long ld;      // token layout: row stride (elements); channel layout: channel stride (elements)
};
struct ProjCall {
  const float* x; long long T; int d;
  const void* w_cat;
  const void* w_ext;                      // bf16 [tiles * 256][16]: columns 0 / 1 = hi / lo split of the fp32 bias, rest 0
  const unsigned char* rowmask;
  int nseg; ProjOut seg[PROJ_MAX_SEG];
};

// 2: CTA-pair kernel (default); 1: single-CTA variant; 0: fused kernel disabled (legacy LayerNorm + GEMM launches)
int g_proj_ctas = 2;
int g_proj_wide = 1;              // 1: 64-column epilogue steps where the segment width allows (AF2_PROJ_WIDE)
int g_proj_balance = 1;           // 1: equal (row unit, column tile) ranges per cluster; 0: round-robin items (AF2_PROJ_BALANCE)
long long* g_proj_trace = nullptr; // device buffer of 1024 stamps when AF2_PROJ_TRACE=1 (debug only)
int g_proj_l2pf = 0;              // 1: producer warps prefetch the next item's rows into L2 (AF2_PROJ_L2PF); measured -0.5..-1 % when off
                                 // (two same-box A/Bs, profiles/r02_ab_*.log): the prefetched lines are evicted by the kernel's own output stream
double g_proj_prod_tiles = 4.0;   // cost of producing one A tile in units of one 256-column MMA tile (AF2_PROJ_PRODTILES)

// can this LN -> Linear cluster run on the fused kernel?
bool proj_dim_ok(int d) { return d % 64 == 0 && d >= 128 && d <= 256; }

template <int CTAS, int KINDS>
int launch_proj_inst(const CUtensorMap& tb, const CUtensorMap* tc, const CUtensorMap& tx, ProjParams& p,
                     double flops, double bytes, cudaStream_t s) {
  using L = ProjSmem<CTAS>;
  static bool configured_dev[MAX_DEVICES] = {false};
  static int max_clusters_dev[MAX_DEVICES] = {0};
  bool& configured = configured_dev[cur_dev()];
  int& max_clusters = max_clusters_dev[cur_dev()];
  auto kern = proj_tc_kernel<CTAS, KINDS>;
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
  const double PROD_TILES = g_proj_prod_tiles;
  int best = 1; double best_cost = 1e30;
  for (int ns = 1; ns <= p.n_tiles_total && ns <= 8; ++ns) {
    const long long items = (long long)m_units * ns;
    const long long waves = (items + max_clusters - 1) / max_clusters;
    const double tiles = (double)((p.n_tiles_total + ns - 1) / ns);
    const double cost = (double)waves * (tiles > PROD_TILES ? tiles : PROD_TILES) + 0.5 * PROD_TILES;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ns; }
  }
  p.nsplit = best;
  p.balance = g_proj_balance;
  p.trace = g_proj_trace;
  p.l2_prefetch = g_proj_l2pf;
  p.x_evict_last = x_hint(p.T, p.d);
  const long long items = p.balance ? (long long)m_units * p.n_tiles_total : (long long)m_units * p.nsplit;
  const int clusters = (int)(items < max_clusters ? items : max_clusters);
  ProfScope ps(s, KC_GEMM_LINEAR, flops, bytes);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * CTAS, 1, 1); cfg.blockDim = dim3(PROJ_THREADS, 1, 1); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 2 : 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tb, tc[0], tc[1], tc[2], tx, p));
  return AF2_OK;
}

int launch_proj(const ProjCall& c, cudaStream_t s) {
  if (c.T <= 0) return AF2_OK;
  if (!proj_dim_ok(c.d)) return fail(AF2_ERR_BAD_ARG, "proj: dim %d unsupported by the fused kernel", c.d);
  if (!c.w_ext) return fail(AF2_ERR_BAD_ARG, "proj: bias block (w_ext) missing");
  const int ctas = g_proj_ctas == 1 ? 1 : 2;
  ProjParams p;
  memset(&p, 0, sizeof(p));
  p.x = c.x; p.T = c.T; p.d = c.d; p.inv_d = 1.0f / (float)c.d; p.eps = 1e-5f;
  p.rowmask = c.rowmask; p.nseg = c.nseg;
  p.m_tiles = (int)((c.T + 127) / 128);
  CUtensorMap tc[3], tb, tx;
  int tile0 = 0, kinds = 0;
  double flops = 0, bytes = (double)c.T * c.d * 4;
  // one staging discipline per launch (the 64-column blocks and the 32-column chunks share the warps' 4 KB buffers with
  // different reuse rules): wide only when every segment's width is a multiple of 64
  bool all_wide = g_proj_wide != 0;
  for (int i = 0; i < c.nseg; ++i) all_wide = all_wide && (c.seg[i].out_cols % 64) == 0;
  for (int i = 0; i < c.nseg; ++i) {
    const ProjOut& o = c.seg[i];
    p.seg[i].tile0 = tile0; p.seg[i].ntiles = o.ntiles; p.seg[i].kind = o.kind; p.seg[i].out_cols = o.out_cols; p.seg[i].map = i;
    tile0 += o.ntiles;
    kinds |= KBIT(o.kind);
    const bool chan = (o.kind == EK_STORE_CH || o.kind == EK_STORE_CH_SIG || o.kind == EK_GATED_CH_SIG);
    if (!aligned16(o.out) || (o.ld * 2) % 16) return fail(AF2_ERR_BAD_ARG, "proj: output %d not TMA-describable", i);
    // every epilogue warp stores from its private 4 KB staging block: 64 columns x 32 rows per store when the segment's
    // width allows it (128-byte rows, 128B swizzle; channel-major: 64 channels x 32 tokens, 64B swizzle), else 32 x 32
    const bool wide = all_wide;
    p.seg[i].wide = wide ? 1 : 0;
    if (!chan) {
      unsigned long long dc[3] = {(unsigned long long)o.out_cols, (unsigned long long)c.T, 1ull};
      unsigned long long sc[2] = {(unsigned long long)o.ld * 2, (unsigned long long)o.ld * c.T * 2};
      unsigned bc[3] = {wide ? 64u : 32u, 32, 1};
      AF2_TRY(make_tmap(&tc[i], o.out, 3, dc, sc, bc, wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
    } else {
      unsigned long long dc[3] = {(unsigned long long)c.T, (unsigned long long)o.out_cols, 1ull};
      unsigned long long sc[2] = {(unsigned long long)o.ld * 2, (unsigned long long)o.ld * o.out_cols * 2};
      unsigned bc[3] = {32, wide ? 64u : 32u, 1};
      AF2_TRY(make_tmap(&tc[i], o.out, 3, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_64B));
    }
    const int W = (o.kind == EK_GATED_TOK_GELU || o.kind == EK_GATED_CH_SIG) ? 2 : 1;
    flops += 2.0 * c.T * (double)o.out_cols * W * c.d;
    bytes += (double)c.T * o.out_cols * 2;
  }
  for (int i = c.nseg; i < 3; ++i) tc[i] = tc[0];
  p.n_tiles_total = tile0;
  bytes += (double)tile0 * 256 * c.d * 2;
  {
    unsigned long long db[2] = {(unsigned long long)c.d, (unsigned long long)tile0 * 256};
    unsigned long long sb[1] = {(unsigned long long)c.d * 2};
    unsigned bb[2] = {64, (unsigned)(256 / ctas)};
    AF2_TRY(make_tmap(&tb, c.w_cat, 2, db, sb, bb, CU_TENSOR_MAP_SWIZZLE_128B));
    unsigned long long dx[2] = {16ull, (unsigned long long)tile0 * 256};
    unsigned long long sx[1] = {32ull};
    unsigned bx[2] = {16, (unsigned)(256 / ctas)};
    AF2_TRY(make_tmap(&tx, c.w_ext, 2, dx, sx, bx, CU_TENSOR_MAP_SWIZZLE_32B));
  }
#define AF2_PROJ_DISPATCH(KS)                                                              \
  if (kinds == KS) {                                                                       \
    if (ctas == 2) return launch_proj_inst<2, KS>(tb, tc, tx, p, flops, bytes, s);         \
    return launch_proj_inst<1, KS>(tb, tc, tx, p, flops, bytes, s);                        \
  }
  AF2_PROJ_DISPATCH(PK_ATTN)
  AF2_PROJ_DISPATCH(PK_TRI)
  AF2_PROJ_DISPATCH(PK_FF)
  AF2_PROJ_DISPATCH(PK_OUTER)
#undef AF2_PROJ_DISPATCH
  return fail(AF2_ERR_BAD_ARG, "proj: no kernel instantiation for epilogue set 0x%x", kinds);
}
