// Host side of the paired-CTA tensor-core kernel (k1c_pair.cuh).  Included by k1b_tc.cu after TcState / TcWorkspace.

struct PairLayerImg {
  PairChunk* chunks = nullptr;
  uint32_t* slots = nullptr;
  uint8_t* wimg = nullptr;
  float* bias = nullptr;
  int n_chunks = 0;
  PairLayerParams prm;
};

struct TcPairState {
  PairLayerImg layers[kMaxLayers];
  bool built = false;
};

static void pair_free(TcPairState* s) {
  if (!s) return;
  for (int l = 0; l < kMaxLayers; ++l) {
    if (s->layers[l].wimg) cudaFree(s->layers[l].wimg);
    if (s->layers[l].chunks) cudaFree(s->layers[l].chunks);
    if (s->layers[l].slots) cudaFree(s->layers[l].slots);
    if (s->layers[l].bias) cudaFree(s->layers[l].bias);
  }
  delete s;
}

// Which models the paired kernel takes: merged factored cells, every layer the same H in {256, 512}, a Dense top
// (its rows ride in the S1 tile of the last layer), ranks <= 256.
static bool pair_supported(const ModelDesc& md, const char** why) {
  if (md.n_out < 1 || md.n_out > 128) { *why = "paired kernel needs a Dense top with 1..128 outputs"; return false; }
  const int H = md.layers[0].units;
  if (H != 256 && H != 512) { *why = "paired kernel needs units in {256, 512}"; return false; }
  if (md.layers[0].d_in > 64) { *why = "layer-0 input_dim above 64"; return false; }
  for (int l = 0; l < md.n_layers; ++l) {
    const LayerDesc& L = md.layers[l];
    if (L.units != H) { *why = "paired kernel needs the same units in every layer"; return false; }
    if (L.n_blocks != 2 || L.blocks[0].left == nullptr || L.blocks[1].left == nullptr) { *why = "paired kernel needs merged factored cells"; return false; }
    if (L.blocks[0].rank > 256 || L.blocks[1].rank > 256) { *why = "ranks above 256"; return false; }
  }
  return true;
}

// Per-layer plan + chunk lists of both ranks.  chunks[r] is rank r's stream in consumption order (S1 then S2 of one step).
static bool pair_layer_plan(const ModelDesc& md, int l, PairLayerParams& p, std::vector<PairChunk> chunks[2], const char** why) {
  const LayerDesc& L = md.layers[l];
  const Block& bw = L.blocks[0];
  const Block& bu = L.blocks[1];
  const bool last = l == md.n_layers - 1;
  const int H = L.units;
  p = PairLayerParams{};
  p.H = H;
  p.ru = bu.rank;
  p.ru_pad = round_up(bu.rank, 16);
  p.kin = l == 0 ? round_up(L.d_in, 16) : round_up(bw.rank, 16);
  p.x_rows = last ? 0 : md.layers[l + 1].blocks[0].rank;
  p.x_pad = round_up(p.x_rows, 16);
  p.n_dense = last ? md.n_out : 0;
  const int xr = last ? p.n_dense : p.x_rows;
  const int xkind = last ? PK_DENSE : PK_X;
  if (p.ru <= 128 && xr <= 128) {
    p.n_mt = 1;
    p.half_kind[0][0] = PK_U; p.half_r0[0][0] = 0;
    p.half_kind[0][1] = xkind; p.half_r0[0][1] = 0;
  } else {
    p.n_mt = 2;
    p.half_kind[0][0] = PK_U; p.half_r0[0][0] = 0;
    p.half_kind[0][1] = p.ru > 128 ? PK_U : PK_NONE; p.half_r0[0][1] = 128;
    p.half_kind[1][0] = xkind; p.half_r0[1][0] = 0;
    p.half_kind[1][1] = xr > 128 ? xkind : PK_NONE; p.half_r0[1][1] = 128;
  }
  const int nubc = H / 256;
  for (int r = 0; r < 2; ++r) {
    chunks[r].clear();
    for (int mt = 0; mt < p.n_mt; ++mt) {
      const int kind = p.half_kind[mt][r], r0 = p.half_r0[mt][r];
      int rows = 0;
      if (kind == PK_U) rows = round_up(imin(128, p.ru - r0), 8);
      else if (kind == PK_X) rows = round_up(imin(128, p.x_rows - r0), 8);
      else if (kind == PK_DENSE) rows = round_up(imin(128, p.n_dense - r0), 8);
      for (int k0 = 0; k0 < H; k0 += 64) chunks[r].push_back(PairChunk{0, (int16_t)kind, (int16_t)rows, 64, (int16_t)r0, (int16_t)k0, 0, 0, 0});
    }
    p.n_s1 = (int)chunks[r].size();
    for_pair_s2(nubc, p.kin, p.ru_pad, [&](int gate, int ub, int late, int k0, int kc) {
      const int kind = late ? PK_S2_LATE : (l == 0 ? PK_S2_EARLY0 : PK_S2_EARLY);
      chunks[r].push_back(PairChunk{0, (int16_t)kind, 128, (int16_t)kc, 0, (int16_t)k0, (int16_t)gate, (int16_t)(r * (H / 2) + ub * 128), 0});
    });
    p.n_s2 = (int)chunks[r].size() - p.n_s1;
  }
  p.in_stages = l == 0 ? 3 : 2;
  p.w_slots = kPairMaxSlots;
  while (p.w_slots > 3 && pair_plan(p).total > kSmemCap) --p.w_slots;
  if (pair_plan(p).total > kSmemCap && p.in_stages > 1) {
    p.in_stages = 1;
    p.w_slots = kPairMaxSlots;
    while (p.w_slots > 3 && pair_plan(p).total > kSmemCap) --p.w_slots;
  }
  if (pair_plan(p).total > kSmemCap) { *why = "paired kernel: activation buffers do not fit shared memory next to a weight ring"; return false; }
  return true;
}

template <int NUBC>
static int pair_launch(const PairPipeParams& pp, uint32_t smem_bytes, bool cooperative, cudaStream_t stream) {
  SVD_CUDA_TRY(cudaFuncSetAttribute(lstm_tc_pair_kernel<NUBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pp.n_layers * pp.n_tiles));
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cooperative ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tc_pair_kernel<NUBC>, pp);
  if (e != cudaSuccess && cooperative) {
    // co-residency is what makes the inter-layer waits safe; with grid <= SM count and one CTA per SM a plain cluster launch is
    // co-resident as well (layer 0's pairs have the lowest block indices and are placed first)
    (void)cudaGetLastError();
    cfg.numAttrs = 0;
    e = cudaLaunchKernelEx(&cfg, lstm_tc_pair_kernel<NUBC>, pp);
  }
  SVD_CUDA_TRY(e);
  return 0;
}

static int run_tc_pair(const ModelDesc& md, TcPairState** state, bool weights_dirty, const ForwardArgs& a, cudaStream_t stream, int n_sm,
                       TcWorkspace* ws, int* launches) {
  const char* why = "";
  int nl = 0;
  if (*state == nullptr) {
    *state = new TcPairState();
    weights_dirty = true;
  }
  TcPairState* st = *state;
  const int L = md.n_layers;
  if (weights_dirty || !st->built) {
    for (int l = 0; l < L; ++l) {
      PairLayerImg& li = st->layers[l];
      PairLayerParams p;
      std::vector<PairChunk> chunks[2];
      SVD_REQUIRE(pair_layer_plan(md, l, p, chunks, &why), "tensor-core engine: %s", why);
      SVD_REQUIRE(chunks[0].size() == chunks[1].size(), "paired kernel: rank streams differ in length");
      // image: rank 0's chunks then rank 1's; slot table entry = (offset >> 8) | (bytes >> 8) << 16
      std::vector<PairChunk> all;
      std::vector<uint32_t> slots;
      uint32_t off = 0;
      for (int r = 0; r < 2; ++r)
        for (PairChunk c : chunks[r]) {
          uint32_t bytes = 0;
          if (c.kind != PK_NONE) bytes = (uint32_t)c.rows * (uint32_t)c.kc * 2u;
          c.byte_off = off;
          slots.push_back((off >> 8) | ((bytes >> 8) << 16));
          off += bytes;
          all.push_back(c);
        }
      SVD_REQUIRE((off >> 8) <= 0xFFFFu, "paired kernel: weight image too large");
      if (li.wimg || li.chunks || li.slots || li.bias) SVD_CUDA_TRY(cudaStreamSynchronize(stream));   // a previous forward may still read them
      if (li.wimg) cudaFree(li.wimg);
      if (li.bias) cudaFree(li.bias);
      if (li.chunks) cudaFree(li.chunks);
      if (li.slots) cudaFree(li.slots);
      li.wimg = nullptr; li.bias = nullptr; li.chunks = nullptr; li.slots = nullptr;
      SVD_CUDA_TRY(cudaMalloc(&li.wimg, off ? off : 256));
      SVD_CUDA_TRY(cudaMalloc(&li.bias, sizeof(float) * 4 * p.H));
      SVD_CUDA_TRY(cudaMalloc(&li.chunks, sizeof(PairChunk) * all.size()));
      SVD_CUDA_TRY(cudaMalloc(&li.slots, sizeof(uint32_t) * slots.size()));
      SVD_CUDA_TRY(cudaMemcpy(li.chunks, all.data(), sizeof(PairChunk) * all.size(), cudaMemcpyHostToDevice));
      SVD_CUDA_TRY(cudaMemcpy(li.slots, slots.data(), sizeof(uint32_t) * slots.size(), cudaMemcpyHostToDevice));
      li.n_chunks = (int)all.size();
      const LayerDesc& Ld = md.layers[l];
      const Block bw_next = l + 1 < L ? md.layers[l + 1].blocks[0] : Block{};
      pack_pair_kernel<<<(unsigned)all.size(), 256, 0, stream>>>(li.chunks, Ld.blocks[0], Ld.blocks[1], bw_next, p.H, Ld.d_in, md.dense_kernel,
                                                                 p.n_dense, md.n_out, reinterpret_cast<__half*>(li.wimg));
      pack_bias_kernel<<<(4 * p.H + 255) / 256, 256, 0, stream>>>(Ld.bias, p.H, li.bias);
      nl += 2;
      p.wimg = li.wimg;
      p.slots = li.slots;
      p.bias = li.bias;
      li.prm = p;
    }
    st->built = true;
    SVD_CUDA_TRY(cudaGetLastError());
  }
  const int B = a.B, T = a.T;
  const int n_tiles = (B + 127) / 128;
  const bool pipe = 2 * L * n_tiles <= n_sm;
  // workspaces
  if (ws->used && ws->last_stream != stream) SVD_CUDA_TRY(cudaStreamSynchronize(ws->last_stream));
  ws->used = true;
  ws->last_stream = stream;
  const int Dpad = st->layers[0].prm.kin;
  const size_t xbytes = (size_t)2 * n_tiles * T * act_tile_bytes(Dpad, 64);
  if (ws->xseq_bytes < xbytes) {
    if (ws->xseq) {
      SVD_CUDA_TRY(cudaStreamSynchronize(stream));
      cudaFree(ws->xseq);
    }
    SVD_CUDA_TRY(cudaMalloc(&ws->xseq, xbytes));
    ws->xseq_bytes = xbytes;
  }
  for (int l = 0; l + 1 < L; ++l) {
    const size_t hb = (size_t)2 * n_tiles * T * act_tile_bytes(st->layers[l].prm.x_pad, 64);
    if (ws->seq_bytes[l] < hb) {
      if (ws->seq[l]) {
        SVD_CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(ws->seq[l]);
      }
      SVD_CUDA_TRY(cudaMalloc(&ws->seq[l], hb));
      ws->seq_bytes[l] = hb;
    }
  }
  {
    const size_t need = (size_t)L * n_tiles;
    if (ws->progress_elems < need) {
      if (ws->progress) {
        SVD_CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(ws->progress);
      }
      SVD_CUDA_TRY(cudaMalloc(&ws->progress, sizeof(int) * need));
      ws->progress_elems = need;
    }
    if (pipe) SVD_CUDA_TRY(cudaMemsetAsync(ws->progress, 0, sizeof(int) * need, stream));
  }
  {
    const int D = md.input_dim;
    const int TT = D <= 16 ? 8 : (D <= 32 ? 4 : 2);
    pack_x_kernel<<<dim3((T + TT - 1) / TT, 2 * n_tiles), 256, sizeof(float) * 64 * TT * D, stream>>>(a.x, B, T, D, Dpad, TT, 64, ws->xseq);
    ++nl;
  }
  PairPipeParams pp{};
  const char* stage_env = getenv("SVDLSTM_PAIR_STAGE");
  pp.stage = stage_env ? atoi(stage_env) : 0;
  pp.off = getenv("SVDLSTM_PAIR_OFF") ? atoi(getenv("SVDLSTM_PAIR_OFF")) : 0;
  static long long* tl_buf = nullptr;
  const bool want_tl = getenv("SVDLSTM_TC_TIMELINE") != nullptr;
  if (want_tl && !tl_buf) SVD_CUDA_TRY(cudaMalloc(&tl_buf, sizeof(long long) * 64 * kMaxLayers));
  if (want_tl) SVD_CUDA_TRY(cudaMemsetAsync(tl_buf, 0, sizeof(long long) * 64 * kMaxLayers, stream));
  pp.tl = want_tl ? tl_buf : nullptr;
  uint32_t smem_max = 0;
  const int nubc = md.layers[0].units / 256;
  for (int l = 0; l < L; ++l) {
    PairLayerParams p = st->layers[l].prm;
    p.T = T;
    p.B = B;
    p.in_seq = l == 0 ? ws->xseq : ws->seq[l - 1];
    p.out_seq = l + 1 < L ? ws->seq[l] : nullptr;
    p.y = a.y;
    p.dense_bias = md.dense_bias;
    p.prog_in = (pipe && l > 0) ? ws->progress + (size_t)(l - 1) * n_tiles : nullptr;
    p.prog_out = (pipe && l + 1 < L) ? ws->progress + (size_t)l * n_tiles : nullptr;
    if (l > 0) {
      const PairLayerParams& q = st->layers[l - 1].prm;
      p.in_contrib = 0;
      for (int mt = 0; mt < q.n_mt; ++mt)
        for (int r = 0; r < 2; ++r) p.in_contrib += q.half_kind[mt][r] == PK_X ? 1 : 0;
    }
    const uint32_t smem_bytes = pair_plan(p).total;
    if (pipe) {
      pp.layer[l] = p;
      smem_max = smem_bytes > smem_max ? smem_bytes : smem_max;
      continue;
    }
    PairPipeParams one{};
    one.tl = nullptr;
    one.stage = pp.stage;
    one.off = pp.off;
    one.layer[0] = p;
    one.n_layers = 1;
    one.n_tiles = n_tiles;
    int lrc = nubc == 1 ? pair_launch<1>(one, smem_bytes, false, stream) : pair_launch<2>(one, smem_bytes, false, stream);
    if (lrc != 0) return lrc;
    ++nl;
  }
  if (pipe) {
    pp.n_layers = L;
    pp.n_tiles = n_tiles;
    int lrc = nubc == 1 ? pair_launch<1>(pp, smem_max, true, stream) : pair_launch<2>(pp, smem_max, true, stream);
    if (lrc != 0) return lrc;
    ++nl;
  }
  SVD_CUDA_TRY(cudaGetLastError());
  *launches = nl;
  if (want_tl && pipe) {   // debug build: stamps of step 40, tile 0 (ns relative to the MMA warp's step start)
    static long long host[64 * kMaxLayers];
    SVD_CUDA_TRY(cudaStreamSynchronize(stream));
    SVD_CUDA_TRY(cudaMemcpy(host, tl_buf, sizeof(host), cudaMemcpyDeviceToHost));
    for (int l = 0; l < L; ++l) {
      const long long* r = host + l * 64;
      fprintf(stderr, "[pair timeline] layer %d MMA: h_wait_start 0 | s1_issued %lld | s1_commit %lld | in_ready %lld | early0 %lld | T_READY %lld | late0+commit %lld | early1 %lld | late1+commit %lld | slot-wait total %lld\n",
              l, r[1] - r[0], r[2] - r[0], r[3] - r[0], r[4] - r[0], r[8] - r[0], r[5] - r[0], r[6] - r[0], r[7] - r[0], r[9]);
      fprintf(stderr, "[pair timeline] layer %d MMA warp, step 40: %lld slots, %lld cycles in slot hand-over (commit + full wait), %lld cycles issuing MMAs\n", l, r[12], r[10], r[11]);
      for (int c = 0; c < 2; ++c) {
        const long long* e = r + 16 + 16 * c;
        fprintf(stderr, "[pair timeline] layer %d EPI cta%d: S1_FULL %lld | T arrived %lld | i seen %lld | g seen %lld | f seen %lld | o seen %lld | H arrived %lld\n", l, c,
                e[0] - r[0], e[1] - r[0], e[2] - r[0], e[3] - r[0], e[4] - r[0], e[5] - r[0], e[6] - r[0]);
      }
    }
  }
  if (getenv("SVDLSTM_PAIR_DEBUG")) {   // bring-up aid: report the first wait that timed out (the launch aborts itself instead of hanging)
    unsigned int dbg[4 * 32] = {};
    SVD_CUDA_TRY(cudaStreamSynchronize(stream));
    SVD_CUDA_TRY(cudaMemcpyFromSymbol(dbg, g_pair_dbg, sizeof(dbg)));
    unsigned int ab = 0, nrec = 0;
    SVD_CUDA_TRY(cudaMemcpyFromSymbol(&ab, g_pair_abort, sizeof(ab)));
    SVD_CUDA_TRY(cudaMemcpyFromSymbol(&nrec, g_pair_nrec, sizeof(nrec)));
    if (ab) {
      const unsigned int zero = 0;
      SVD_CUDA_TRY(cudaMemcpyToSymbol(g_pair_abort, &zero, sizeof(zero)));
      SVD_CUDA_TRY(cudaMemcpyToSymbol(g_pair_nrec, &zero, sizeof(zero)));
      char msg[2048];
      int n = snprintf(msg, sizeof(msg), "paired kernel: %u stuck waits (site@step block/thread parity):", nrec);
      for (unsigned int i = 0; i < nrec && i < 32u && n < (int)sizeof(msg) - 64; ++i)
        n += snprintf(msg + n, sizeof(msg) - n, " %u@%u b%u/t%u p%u;", dbg[4 * i] & 0xFFu, dbg[4 * i] >> 8, dbg[4 * i + 1], dbg[4 * i + 2], dbg[4 * i + 3]);
      set_error("%s", msg);
      return -1;
    }
  }
  return 0;
}
