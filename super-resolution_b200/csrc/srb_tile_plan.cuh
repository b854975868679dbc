// srb_tile_plan.cuh -- host side of the fused tile kernel: planning (which samples are regular, the
// (frame, tap) entry lists per sub-pixel phase, the border band, the table-driven residual pass),
// device upload, tensor maps and launches.  plan_tile_model() is pure host code (no CUDA calls) and
// is what srb_plan() exposes for the CPU tests.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "srb_common.cuh"
#include "srb_kernels_band.cuh"
#include "srb_kernels_tile.cuh"
#include "srb_kernels_tilez.cuh"
#include "srb_kernels_regtile.cuh"

namespace srb {

// ================================================================================================
// host side
// ================================================================================================
// Everything the tile kernel needs to know about a model, computed on the host without touching CUDA.
struct TilePlan {
  bool supported = false;
  std::string why;         // why the tile kernel does not cover this model
  bool frac = false;       // some shift is fractional: bilinear forward / transpose taps
  int KH = 0;              // PSF half width
  double u[9], v[9];       // psf[i][j] = u[i] * v[j]
  std::vector<TEntry> entries;   // (frame, tap) entries, grouped by sub-pixel phase
  std::vector<int> phase_begin;  // [s*s + 1]
  int qoff_min_r = 0, qoff_max_r = 0, qoff_min_c = 0, qoff_max_c = 0;
  BandGeom band{};         // special LR samples
  BandGeom reach{};        // HR pixels the special samples can reach
  bool has_band = false;
  std::vector<TFast> fast[2];    // table-driven residual pass, tile height 32 / 64 (empty: not applicable)
  std::vector<long long> fast_y[2];  // observation offsets of entries 1 .. fast_E-1, [e-1][items]
  int fast_E = 0;                // entries per sub-pixel phase in the table-driven pass (1, 2 or 4)
  // Z layout (k_tile_z): 0 the model does not qualify; 1 integer shifts, PSF 3x3 .. 9x9 and exactly one
  // frame on every sub-pixel phase; 2 the same with some phases empty (a frame shard; HOLES variant)
  int zlayout = 0;
  // transposed Z layout (k_tile_zt): integer shifts, PSF 3x3 .. 9x9, at most one DISTINCT shift per sub-pixel
  // phase and the same number zt_n of frames on every non-empty phase (frames with the same shift are averaged
  // at upload: cfg4 has 2 per phase, cfg5 4)
  bool zt = false;
  int zt_n = 0;
  // zt_n > 1: the groups of frames with equal shifts = the non-empty sub-pixel phases: their phase index and one
  // frame of each (the band kernels walk groups instead of frames, srb_kernels_band.cuh)
  std::vector<int> group_phase, group_frame;
  int max_shift = 0;   // largest |shift| over the frames, HR pixels, rounded up (+1 for the bilinear partner)
};

struct TileState {
  TilePlan plan;
  bool supported = false;
  bool frac = false;
  int KH = 0;
  double u[9], v[9];
  TEntry* d_entries = nullptr;
  int* d_phase_begin = nullptr;
  int num_entries = 0;
  int qoff_min_r = 0, qoff_max_r = 0, qoff_min_c = 0, qoff_max_c = 0;
  BandGeom band{};       // special LR samples
  BandGeom reach{};      // HR pixels the special samples can reach
  bool has_band = false;
  double* d_pooled = nullptr;  // [N][Ct][band.count()]
  TFast* d_fast[2] = {nullptr, nullptr};  // tile height 32 / 64; NULL when the model does not qualify
  long long* d_fast_y[2] = {nullptr, nullptr};
  double* d_yz = nullptr;  // observations on the HR grid [Ct][H][W] (k_tile_z; SRB_ZLAYOUT=1), else NULL
  bool yz_valid = false;   // d_yz matches the observations currently in d_y
  bool yz_holes = false;   // some sub-pixel phases have no frame (NaN in d_yz; k_tile_z<.., true>)
  double* d_yzt = nullptr; // observations in the transposed, padded Z layout [Ct][cols_p][rows_p] (k_tile_zt)
  int yzt_rows = 0, yzt_cols = 0;
  double* d_yvar = nullptr;      // [Ct] constant part of the merged data cost (zt_n > 1), else NULL
  double* d_yvar_part = nullptr; // per-block partial sums of k_build_yzt (+ k_band_merge)
  size_t yvar_stride = 0, yvar_band_offset = 0;   // slots per channel in d_yvar_part; first slot of the band sums
  int band_groups = 0;           // > 0: the border-band kernels run on merged groups of frames
  int* d_group_phase = nullptr;  // [band_groups]
  int* d_group_frame = nullptr;  // [band_groups]
  double* d_yband = nullptr;     // [band_groups][Ct][band.count()] group means of the band samples
  double* d_stage = nullptr;     // [2 * kStageBlocks] first-stage sums of the cost reduction
  static constexpr int kStageBlocks = 64;
  bool tma_ok = false;
  int tile_h = 32;         // SRB_TILE_H=32|64 overrides (tuning knob; 32 measured faster at cfg3)
  void* encode = nullptr;  // cuTensorMapEncodeTiled
  std::string why;         // why the tile kernel does not cover this model
};

inline TileState*& tile_state(srb_ctx* c) { return reinterpret_cast<TileState*&>(c->fused); }
inline const TileState* tile_state(const srb_ctx* c) { return reinterpret_cast<const TileState*>(c->fused); }
inline bool fused_supported(const srb_ctx* c) {
  const TileState* st = tile_state(c);
  return st && st->supported;
}

inline void fused_teardown(srb_ctx* c) {
  TileState*& st = tile_state(c);
  if (!st) return;
  if (st->d_entries) cudaFree(st->d_entries);
  if (st->d_phase_begin) cudaFree(st->d_phase_begin);
  if (st->d_pooled) cudaFree(st->d_pooled);
  for (TFast* f : st->d_fast)
    if (f) cudaFree(f);
  for (long long* f : st->d_fast_y)
    if (f) cudaFree(f);
  if (st->d_yz) cudaFree(st->d_yz);
  if (st->d_yzt) cudaFree(st->d_yzt);
  if (st->d_yvar) cudaFree(st->d_yvar);
  if (st->d_yvar_part) cudaFree(st->d_yvar_part);
  if (st->d_group_phase) cudaFree(st->d_group_phase);
  if (st->d_group_frame) cudaFree(st->d_group_frame);
  if (st->d_yband) cudaFree(st->d_yband);
  if (st->d_stage) cudaFree(st->d_stage);
  delete st;
  st = nullptr;
}

// Rank-1 factorisation psf = u v^T (true for blur_module.cpp:20-22's outer-product Gaussian).
inline bool factor_separable(const std::vector<double>& psf, int K, double* u, double* v) {
  int bi = 0, bj = 0;
  double best = 0.0;
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j)
      if (std::fabs(psf[i * K + j]) > best) best = std::fabs(psf[i * K + j]), bi = i, bj = j;
  if (!(best > 0.0)) return false;
  const double pivot = psf[bi * K + bj];
  for (int i = 0; i < K; ++i) u[i] = psf[i * K + bj];
  for (int j = 0; j < K; ++j) v[j] = psf[bi * K + j] / pivot;
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j)
      if (std::fabs(psf[i * K + j] - u[i] * v[j]) > 8.0 * 2.220446049250313e-16 * best) return false;
  return true;
}

inline int pymod(int a, int b) {
  int m = a % b;
  return m < 0 ? m + b : m;
}
inline int pydiv(int a, int b) { return (a - pymod(a, b)) / b; }

// Is LR sample q (one dimension; HR size L, half PSF width hk, scale s) "special" for a frame whose
// quantised forward / transpose warps are n32 / t32 (1/32 px)?  Regular means: commuting the PSF
// with the shift changes neither the LR prediction nor the back-projected gradient, and every
// transpose tap lands inside the image or in the Z halo (PSF half width) of a border tile.
inline bool sample_is_special(int q, int L, int hk, int s, int n32, int t32) {
  const int n = n32 >> 5, fa = (n32 & 31) ? 1 : 0;
  const int nt = t32 >> 5, fb = (t32 & 31) ? 1 : 0;
  const int p0 = s * q;
  auto in = [L](int p) { return p >= 0 && p < L; };
  for (int b = 0; b <= fb; ++b)
    if (p0 - nt - b < -hk || p0 - nt - b >= L + hk) return true;  // beyond the Z halo of the border tiles
  for (int i = -hk; i <= hk; ++i) {
    if (in(p0 + i)) continue;
    for (int a = 0; a <= fa; ++a)
      if (in(p0 + i + n + a)) return true;   // forward: window tap clipped before, not after, the shift
    for (int b = 0; b <= fb; ++b)
      if (in(p0 + i - nt - b)) return true;  // transpose: G outside the image reaches a pixel inside
  }
  return false;
}

// Pure host planning.  warp_fwd / warp_tr are the quantised forward / transpose warps per frame.
inline void plan_tile_model(const Geometry& G, const std::vector<double>& psf_h, const std::vector<WarpQ>& warp_fwd,
                            const std::vector<WarpQ>& warp_tr, bool warps_uniform, bool warps_integer, TilePlan* st) {
  if (G.K > 9) { st->why = "PSF larger than 9x9"; return; }
  if (G.s > FT_MAX_SCALE) { st->why = "downsampling scale larger than 8"; return; }
  if (!warps_uniform) { st->why = "a shift sits on a fixed-point rounding boundary"; return; }
  if (!factor_separable(psf_h, G.K, st->u, st->v)) { st->why = "PSF is not separable (rank 1)"; return; }
  st->KH = G.hk;
  st->frac = !warps_integer;
  const int s = G.s, hk = G.hk;
  const int FR = st->frac ? 1 : 0;

  // ---- band of special samples (frame independent: the union over frames) -----------------------
  int lo[2] = {0, 0}, hi[2] = {G.h, G.w};
  int max_shift = 0;
  for (int dim = 0; dim < 2; ++dim) {
    const int L = dim == 0 ? G.H : G.W, l = dim == 0 ? G.h : G.w;
    for (int k = 0; k < G.N; ++k) {
      const int n32 = dim == 0 ? warp_fwd[k].nY : warp_fwd[k].nX;
      const int t32 = dim == 0 ? warp_tr[k].nY : warp_tr[k].nX;
      max_shift = std::max(max_shift, std::max(std::abs(n32 >> 5), std::abs(t32 >> 5)) + 1);
      for (int q = 0; q < l; ++q) {
        if (!sample_is_special(q, L, hk, s, n32, t32)) continue;
        if (2 * q < l) lo[dim] = std::max(lo[dim], q + 1);
        else hi[dim] = std::min(hi[dim], q);
      }
    }
    if (lo[dim] >= hi[dim]) { st->why = "image too small for the shifts (every sample is a border sample)"; return; }
  }
  st->max_shift = max_shift;
  st->band = BandGeom{G.h, G.w, lo[0], hi[0], lo[1], hi[1]};
  st->has_band = st->band.count() > 0;
  {
    // HR pixels reachable from the band: PSF half width + the largest shift + 1 around its samples
    const int m = hk + max_shift + 1;
    BandGeom R{G.H, G.W, 0, G.H, 0, G.W};
    if (lo[0] > 0) R.lo_r = std::min(G.H, s * (lo[0] - 1) + m + 1);
    if (hi[0] < G.h) R.hi_r = std::max(0, s * hi[0] - m);
    if (lo[1] > 0) R.lo_c = std::min(G.W, s * (lo[1] - 1) + m + 1);
    if (hi[1] < G.w) R.hi_c = std::max(0, s * hi[1] - m);
    if (R.lo_r >= R.hi_r || R.lo_c >= R.hi_c) { st->why = "image too small for the shifts"; return; }
    st->reach = R;
  }

  // ---- phase lists of the regular samples -------------------------------------------------------
  const int HB = hk + FR;
  const int BP = (FT_W + 2 * HB) | 1;
  std::vector<std::vector<TEntry>> lists((size_t)s * s);
  std::vector<int> first_frame((size_t)s * s, -1);   // a frame of every non-empty phase
  const long long hw = (long long)G.h * G.w;
  bool first = true;
  for (int k = 0; k < G.N; ++k) {
    const int nY = warp_fwd[k].nY, nX = warp_fwd[k].nX;
    const int tY = warp_tr[k].nY, tX = warp_tr[k].nX;
    const int n_r = nY >> 5, n_c = nX >> 5, fy = nY & 31, fx = nX & 31;
    const int t_r = tY >> 5, t_c = tX >> 5, ty = tY & 31, tx = tX & 31;
    for (int a = 0; a <= (ty ? 1 : 0); ++a)
      for (int b = 0; b <= (tx ? 1 : 0); ++b)
        for (int pr = 0; pr < s; ++pr)
          for (int pc = 0; pc < s; ++pc) {
            if (pymod(pr + t_r + a, s) != 0 || pymod(pc + t_c + b, s) != 0) continue;
            const int qoff_r = pydiv(pr + t_r + a, s), qoff_c = pydiv(pc + t_c + b, s);
            const int dr = t_r + n_r + a, dc = t_c + n_c + b;
            // the Bx samples (and their bilinear partners) must stay inside the Bx halo
            if (dr < -FR || dr > 0 || dc < -FR || dc > 0) {
              st->why = "forward and transpose warps of a frame quantise too far apart";
              return;
            }
            if (std::abs(qoff_r) > 30000 || std::abs(qoff_c) > 30000) { st->why = "shift too large"; return; }
            TEntry e;
            e.yoff = (long long)k * G.Ct * hw + (long long)qoff_r * G.w + qoff_c;
            e.bxoff = (HB - hk + dr) * BP + (HB - hk + dc);
            e.qoff = (qoff_r & 0xffff) | (qoff_c << 16);
            const double wy = a ? ty / 32.0 : (32 - ty) / 32.0, wx = b ? tx / 32.0 : (32 - tx) / 32.0;
            e.wT = wy * wx;
            e.fy = (short)fy;
            e.fx = (short)fx;
            e.owner = (a == 0 && b == 0) ? 1 : 0;
            lists[(size_t)pr * s + pc].push_back(e);
            if (first_frame[(size_t)pr * s + pc] < 0) first_frame[(size_t)pr * s + pc] = k;
            if (first) {
              st->qoff_min_r = st->qoff_max_r = qoff_r;
              st->qoff_min_c = st->qoff_max_c = qoff_c;
              first = false;
            }
            st->qoff_min_r = std::min(st->qoff_min_r, qoff_r); st->qoff_max_r = std::max(st->qoff_max_r, qoff_r);
            st->qoff_min_c = std::min(st->qoff_min_c, qoff_c); st->qoff_max_c = std::max(st->qoff_max_c, qoff_c);
          }
  }
  std::vector<TEntry>& flat = st->entries;
  std::vector<int>& begin = st->phase_begin;
  begin.assign((size_t)s * s + 1, 0);
  for (size_t ph = 0; ph < lists.size(); ++ph) {
    begin[ph] = (int)flat.size();
    flat.insert(flat.end(), lists[ph].begin(), lists[ph].end());
  }
  begin[(size_t)s * s] = (int)flat.size();
  if ((int)flat.size() > FT_MAX_ENTRIES) { st->why = "too many (frame, tap) entries for shared memory"; return; }

  // ---- table-driven residual pass (TFast): integer shifts, the same number E in {1, 2, 4} of frames
  //      on every sub-pixel phase, s | tile size ----------------------------------------------------
  {
    const int E = (int)lists[0].size();
    bool balanced = !st->frac && (32 % s == 0) && (E == 1 || E == 2 || E == 4);
    for (size_t ph = 0; balanced && ph < lists.size(); ++ph) {
      balanced = (int)lists[ph].size() == E;
      for (const TEntry& e : lists[ph]) balanced = balanced && e.bxoff == lists[ph][0].bxoff;
    }
    for (int v = 0; balanced && v < 2; ++v) {
      const int TH = v == 0 ? 32 : 64;
      const int ZW = FT_W + 2 * hk, ZP = ZW | 1;
      std::vector<TFast> tab;
      std::vector<std::vector<long long>> extra((size_t)std::max(E - 1, 0));
      auto item = [&](int r, int c) {  // Z-region pixel (r, c): tile-relative HR position (r-hk, c-hk)
        const int pr = r - hk, pc = c - hk;
        const int dmr = pydiv(pr, s), dmc = pydiv(pc, s);
        const std::vector<TEntry>& le = lists[(size_t)(pr - dmr * s) * s + (pc - dmc * s)];
        TFast f;
        f.yrel = le[0].yoff + (long long)dmr * G.w + dmc;
        f.bxo = r * BP + c + le[0].bxoff;
        f.zo = r * ZP + c;
        tab.push_back(f);
        for (int e = 1; e < E; ++e) extra[(size_t)e - 1].push_back(le[(size_t)e].yoff + (long long)dmr * G.w + dmc);
      };
      for (int blk = 0; blk < TH / 32; ++blk)          // pass A: id = blk*(FT_W*s) + rho*FT_W + cm
        for (int rho = 0; rho < s; ++rho)
          for (int cm = 0; cm < FT_W; ++cm) item(hk + blk * 32 + rho, hk + cm);
      for (int rr = 0; rr < 2 * hk; ++rr)             // ring: top + bottom halo rows, full width
        for (int c = 0; c < ZW; ++c) item(rr < hk ? rr : rr + TH, c);
      for (int rm = 0; rm < TH; ++rm)                 // ring: left / right halo columns
        for (int hc = 0; hc < 2 * hk; ++hc) item(hk + rm, hc < hk ? hc : hc + FT_W);
      st->fast[v] = tab;
      st->fast_y[v].clear();
      for (const auto& ex : extra) st->fast_y[v].insert(st->fast_y[v].end(), ex.begin(), ex.end());
      st->fast_E = E;
    }
  }

  // ---- Z layout: every HR pixel receives at most one regular sample -----------------------------
  if (!st->frac && hk >= 1 && hk <= 4) {  // (an empty frame shard qualifies too: every position is a hole)
    bool at_most_one = true, all_one = true;
    for (size_t ph = 0; ph < lists.size(); ++ph) {
      at_most_one = at_most_one && lists[ph].size() <= 1;
      all_one = all_one && lists[ph].size() == 1;
    }
    st->zlayout = flat.empty() ? 0 : (all_one && !st->fast[0].empty()) ? 1 : (at_most_one && !all_one) ? 2 : 0;
    // k_tile_zt: one distinct shift per phase, the same number of frames on every non-empty phase
    bool mergeable = true;
    int n = 0;
    for (size_t ph = 0; ph < lists.size(); ++ph) {
      if (lists[ph].empty()) continue;
      if (n == 0) n = (int)lists[ph].size();
      mergeable = mergeable && (int)lists[ph].size() == n;
      for (const TEntry& e : lists[ph])
        mergeable = mergeable && e.qoff == lists[ph][0].qoff && e.bxoff == lists[ph][0].bxoff;
    }
    st->zt = mergeable;
    st->zt_n = mergeable ? std::max(n, 1) : 0;
    if (mergeable && n > 1)
      for (size_t ph = 0; ph < lists.size(); ++ph)
        if (!lists[ph].empty()) {
          st->group_phase.push_back((int)ph);
          st->group_frame.push_back(first_frame[ph]);
        }
  }

  st->supported = true;
}

// Plans the model and uploads the tables.  A model the tile kernel does not cover leaves
// supported == false (the reference-order kernels run instead) and is not an error.
inline srb_status fused_setup(srb_ctx* c) {
  TileState* st = new TileState();
  tile_state(c) = st;
  const Geometry& G = c->g;
  TilePlan& plan = st->plan;
  plan_tile_model(G, c->psf_h, c->warp_fwd, c->warp_tr, c->warps_uniform, c->warps_integer, &plan);
  st->why = plan.why;
  if (!plan.supported) return SRB_OK;
  st->frac = plan.frac;
  st->KH = plan.KH;
  for (int i = 0; i < 9; ++i) st->u[i] = plan.u[i], st->v[i] = plan.v[i];
  st->num_entries = (int)plan.entries.size();
  st->qoff_min_r = plan.qoff_min_r; st->qoff_max_r = plan.qoff_max_r;
  st->qoff_min_c = plan.qoff_min_c; st->qoff_max_c = plan.qoff_max_c;
  st->band = plan.band;
  st->reach = plan.reach;
  st->has_band = plan.has_band;
  if (cudaMalloc((void**)&st->d_entries, (plan.entries.size() + 1) * sizeof(TEntry)) != cudaSuccess ||
      cudaMalloc((void**)&st->d_phase_begin, plan.phase_begin.size() * sizeof(int)) != cudaSuccess)
    return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (tile kernel tables)");
  SRB_CUDA_CHECK(c, cudaMemcpy(st->d_entries, plan.entries.data(), plan.entries.size() * sizeof(TEntry), cudaMemcpyHostToDevice));
  SRB_CUDA_CHECK(c, cudaMemcpy(st->d_phase_begin, plan.phase_begin.data(), plan.phase_begin.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (st->has_band) {
    const size_t n = (size_t)G.N * G.Ct * (size_t)st->band.count();
    if (cudaMalloc((void**)&st->d_pooled, n * sizeof(double)) != cudaSuccess)
      return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (border band residuals)");
  }
  for (int v = 0; v < 2; ++v) {
    if (plan.fast[v].empty()) continue;
    if (cudaMalloc((void**)&st->d_fast[v], (plan.fast[v].size() + 1) * sizeof(TFast)) != cudaSuccess)
      return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (tile kernel tables)");
    SRB_CUDA_CHECK(c, cudaMemcpy(st->d_fast[v], plan.fast[v].data(), plan.fast[v].size() * sizeof(TFast), cudaMemcpyHostToDevice));
    if (!plan.fast_y[v].empty()) {
      if (cudaMalloc((void**)&st->d_fast_y[v], plan.fast_y[v].size() * sizeof(long long)) != cudaSuccess)
        return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (tile kernel tables)");
      SRB_CUDA_CHECK(c, cudaMemcpy(st->d_fast_y[v], plan.fast_y[v].data(), plan.fast_y[v].size() * sizeof(long long), cudaMemcpyHostToDevice));
    }
  }
  // ---- TMA: the tensor-map encoder comes from the driver through the runtime ---------------------
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess && fn != nullptr)
      st->encode = fn;
    else
      (void)cudaGetLastError();
    st->tma_ok = st->encode != nullptr && (G.W % 2 == 0);  // global strides must be multiples of 16 B
  }
  if (const char* e = getenv("SRB_TILE_H")) st->tile_h = atoi(e) == 64 ? 64 : 32;
  // Z layout (k_tile_z): integer shifts, one frame per sub-pixel phase, PSF of 3x3 .. 9x9, TMA.
  // Models where some phases have no frame at all -- the frame shards of a multi-GPU run -- take the
  // HOLES variant (NaN in yz).  SRB_ZLAYOUT=0 keeps k_tile everywhere, 1 only the all-phases form (A/B).
  {
    const char* e = getenv("SRB_ZLAYOUT");
    const int mode = e == nullptr ? 2 : atoi(e);
    const bool take = plan.zlayout == 1 ? mode != 0 : plan.zlayout == 2 ? mode >= 2 : false;
    // SRB_ZT=0 keeps the round-1 row-major Z layout (k_tile_z) for A/B runs; default: k_tile_zt
    const char* zt_env = getenv("SRB_ZT");
    const bool use_zt = plan.zt && mode != 0 && (zt_env == nullptr || atoi(zt_env) != 0) && st->tma_ok && st->tile_h == 32;
    if (use_zt) {
      st->yzt_rows = zt_rows_padded(G.H, plan.KH);
      st->yzt_cols = zt_cols_padded(G.W, plan.KH);
      if (cudaMalloc((void**)&st->d_yzt, (size_t)G.Ct * st->yzt_rows * st->yzt_cols * sizeof(double)) != cudaSuccess)
        return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (observations in transposed Z layout)");
      if (plan.zt_n > 1) {
        const size_t per_channel = (size_t)((st->yzt_rows + 255) / 256) * st->yzt_cols;
        st->yvar_band_offset = per_channel;
        st->yvar_stride = per_channel;
        if (st->has_band) {  // the border band walks the merged groups too
          const int ng = (int)plan.group_phase.size();
          const size_t bcnt = (size_t)st->band.count();
          st->band_groups = ng;
          st->yvar_stride += (size_t)ng * ((bcnt + 255) / 256);
          if (cudaMalloc((void**)&st->d_group_phase, ng * sizeof(int)) != cudaSuccess ||
              cudaMalloc((void**)&st->d_group_frame, ng * sizeof(int)) != cudaSuccess ||
              cudaMalloc((void**)&st->d_yband, (size_t)ng * G.Ct * bcnt * sizeof(double)) != cudaSuccess)
            return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (merged border band)");
          SRB_CUDA_CHECK(c, cudaMemcpy(st->d_group_phase, plan.group_phase.data(), ng * sizeof(int), cudaMemcpyHostToDevice));
          SRB_CUDA_CHECK(c, cudaMemcpy(st->d_group_frame, plan.group_frame.data(), ng * sizeof(int), cudaMemcpyHostToDevice));
        }
        if (cudaMalloc((void**)&st->d_yvar, (size_t)G.Ct * sizeof(double)) != cudaSuccess ||
            cudaMalloc((void**)&st->d_yvar_part, (size_t)G.Ct * st->yvar_stride * sizeof(double)) != cudaSuccess)
          return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (merged-frame cost constants)");
      }
    } else if (take && st->tma_ok && st->tile_h == 32) {
      if (cudaMalloc((void**)&st->d_yz, (size_t)G.Ct * G.H * G.W * sizeof(double)) != cudaSuccess)
        return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (observations in Z layout)");
      st->yz_holes = plan.zlayout == 2;
    }
  }
  if (cudaMalloc((void**)&st->d_stage, 2 * TileState::kStageBlocks * sizeof(double)) != cudaSuccess)
    return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (cost reduction stage)");
  st->supported = true;
  return SRB_OK;
}

// After new observations were stored in c->d_y (stream-ordered): refresh their Z-layout copy.
inline srb_status fused_observations_changed(srb_ctx* c) {
  TileState* st = tile_state(c);
  if (!st || !st->supported) return SRB_OK;
  const Geometry& G = c->g;
  if (st->d_yzt) {
    const int KH = st->KH;
    const dim3 grid((unsigned)((st->yzt_rows + 255) / 256), (unsigned)st->yzt_cols, (unsigned)G.Ct);
    k_build_yzt<<<grid, 256, 0, c->stream>>>(G.h, G.w, G.s, st->yzt_rows, st->yzt_cols, (KH + 1) & ~1, KH,
                                             st->band.lo_r, st->band.hi_r, st->band.lo_c, st->band.hi_c,
                                             st->d_entries, st->d_phase_begin, c->d_y, st->d_yzt, st->d_yvar_part,
                                             st->yvar_stride);
    c->timing.kernel_launches += 1;
    if (st->d_yvar) {
      if (st->band_groups > 0) {
        const dim3 bg((unsigned)((st->band.count() + 255) / 256), (unsigned)(st->band_groups * G.Ct));
        k_band_merge<TEntry><<<bg, 256, 0, c->stream>>>(st->band, G.Ct, G.w, st->d_entries, st->d_phase_begin, st->d_group_phase,
                                                        c->d_y, st->d_yband, st->d_yvar_part, st->yvar_stride, st->yvar_band_offset);
        c->timing.kernel_launches += 1;
      }
      k_reduce_yvar<<<G.Ct, 1024, 0, c->stream>>>(st->d_yvar_part, st->yvar_stride, (double)G.s * G.s, st->d_yvar);
      c->timing.kernel_launches += 1;
    }
    SRB_CUDA_CHECK(c, cudaGetLastError());
    st->yz_valid = true;
    return SRB_OK;
  }
  if (!st->d_yz) return SRB_OK;
  const dim3 grid((unsigned)((G.W + 255) / 256), (unsigned)G.H, (unsigned)G.Ct);
  k_build_yz<<<grid, 256, 0, c->stream>>>(G.H, G.W, G.h, G.w, G.s, st->d_entries, st->d_phase_begin, c->d_y, st->d_yz);
  c->timing.kernel_launches += 1;
  SRB_CUDA_CHECK(c, cudaGetLastError());
  st->yz_valid = true;
  return SRB_OK;
}

typedef CUresult (*srb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

// 3-D tensor map over [planes][H][W] doubles with a (box_w x box_h x 1) box, zero fill outside.
inline bool make_plane_map(const TileState* st, CUtensorMap* map, const double* base, int W, int H, int planes,
                           int box_w, int box_h) {
  if (((size_t)base & 15) != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * H * 8};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = ((srb_encode_tiled_fn)st->encode)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int KH, bool FRAC, int TH, int FE>
inline srb_status tile_launch(srb_ctx* c, TileParams& P, int unit_end) {
  using D = TileDims<KH, FRAC, TH>;
  const dim3 grid((P.W + FT_W - 1) / FT_W, unit_end - P.unit_begin, 1);
  const TileState* st = tile_state(c);
  CUtensorMap mx, mw;
  memset(&mx, 0, sizeof mx);
  memset(&mw, 0, sizeof mw);
  P.use_tma = 0;
  if (st->tma_ok) {
    bool ok = make_plane_map(st, &mx, P.x, P.W, P.H, P.Ca, D::XW, D::XH);
    if (ok && P.reg_fused) ok = make_plane_map(st, &mw, P.wts, P.W, P.H, P.Ca, D::WW, D::WH);
    P.use_tma = ok ? 1 : 0;
  }
  static const size_t smem_pad = getenv("SRB_SMEM_PAD") ? (size_t)atoi(getenv("SRB_SMEM_PAD")) : 0;  // occupancy experiments
  const size_t smem = D::smem_bytes(P.num_entries) + smem_pad;
  static size_t attr_set[64] = {};
  if (c->device >= 64 || attr_set[c->device] < smem) {
    SRB_CUDA_CHECK(c, cudaFuncSetAttribute(k_tile<KH, FRAC, TH, FE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (c->device < 64) attr_set[c->device] = smem;
  }
  if (c->profiling) cudaEventRecord(c->ev[4], c->stream);
  k_tile<KH, FRAC, TH, FE><<<grid, D::NT, smem, c->stream>>>(P, mx, mw);
  if (c->profiling) cudaEventRecord(c->ev[5], c->stream);
  return SRB_OK;
}

// k_tile_z launch (Z layout); returns SRB_ERR_STATE without launching when a tensor map cannot be made
// (the caller then launches k_tile).
template <int KH, bool HOLES>
inline srb_status tile_launch_z(srb_ctx* c, TileParams& P, int unit_end) {
  using D = TileDims<KH, false, 32>;
  constexpr int HYC = (KH + 1) & ~1;
  const dim3 grid((P.W + FT_W - 1) / FT_W, unit_end - P.unit_begin, 1);
  const TileState* st = tile_state(c);
  CUtensorMap mx, mw, my;
  memset(&mx, 0, sizeof mx);
  memset(&mw, 0, sizeof mw);
  memset(&my, 0, sizeof my);
  bool ok = make_plane_map(st, &mx, P.x, P.W, P.H, P.Ca, D::XW, D::XH);
  if (ok && P.reg_fused) ok = make_plane_map(st, &mw, P.wts, P.W, P.H, P.Ca, D::WW, D::WH);
  if (ok) ok = make_plane_map(st, &my, P.yz, P.W, P.H, P.Ct, FT_W + 2 * HYC, 32 + 2 * KH);
  if (!ok) return SRB_ERR_STATE;
  P.use_tma = 1;
  const size_t smem = D::smem_bytes(P.num_entries);
  static size_t attr_set[64] = {};
  if (c->device >= 64 || attr_set[c->device] < smem) {
    SRB_CUDA_CHECK(c, cudaFuncSetAttribute(k_tile_z<KH, HOLES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (c->device < 64) attr_set[c->device] = smem;
  }
  if (c->profiling) cudaEventRecord(c->ev[4], c->stream);
  k_tile_z<KH, HOLES><<<grid, D::NT, smem, c->stream>>>(P, mx, mw, my);
  if (c->profiling) cudaEventRecord(c->ev[5], c->stream);
  return SRB_OK;
}

// k_tile_zt launch (transposed Z layout); SRB_ERR_STATE without launching when a tensor map cannot be made.
template <int KH>
inline srb_status tile_launch_zt(srb_ctx* c, TileParams& P, int unit_end) {
  using Z = ZtDims<KH>;
  using D = typename Z::D;
  const dim3 grid((P.W + FT_W - 1) / FT_W, unit_end - P.unit_begin, 1);
  const TileState* st = tile_state(c);
  CUtensorMap mx, mw, my;
  memset(&mx, 0, sizeof mx);
  memset(&mw, 0, sizeof mw);
  memset(&my, 0, sizeof my);
  bool ok = make_plane_map(st, &mx, P.x, P.W, P.H, P.Ca, D::XW, D::XH);
  if (ok && P.reg_fused) ok = make_plane_map(st, &mw, P.wts, P.W, P.H, P.Ca, D::WW, D::WH);
  // [planes][cols_p][rows_p]: the inner (contiguous) dimension is the HR row
  if (ok) ok = make_plane_map(st, &my, st->d_yzt, st->yzt_rows, st->yzt_cols, P.Ct, Z::YR, Z::YC);
  if (!ok) return SRB_ERR_STATE;
  P.use_tma = 1;
  static bool attr_set[64] = {};
  if (c->device >= 64 || !attr_set[c->device]) {
    SRB_CUDA_CHECK(c, cudaFuncSetAttribute(k_tile_zt<KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Z::SMEM_BYTES));
    if (c->device < 64) attr_set[c->device] = true;
  }
  if (c->profiling) cudaEventRecord(c->ev[4], c->stream);
  k_tile_zt<KH><<<grid, Z::NT, Z::SMEM_BYTES, c->stream>>>(P, mx, mw, my);
  if (c->profiling) cudaEventRecord(c->ev[5], c->stream);
  return SRB_OK;
}

// Tile height the current model runs with, and the number of (channel, tile row) units.
inline int tile_height(const srb_ctx* c) {
  const TileState* st = tile_state(c);
  return (c->g.H >= 256 && c->g.W >= 256) ? st->tile_h : 32;
}
inline int tile_rows_per_channel(const srb_ctx* c) {
  const int TH = tile_height(c);
  return (c->g.H + TH - 1) / TH;
}

// Regularizers the fused path evaluates itself: 2-D TV inside the tile kernel, BTV (R <= 4) and 3-D TV by the
// tiled kernels of srb_kernels_regtile.cuh right behind it.  Anything else runs the reference-order kernels
// after the tile kernel (eval_core).
inline bool fused_reg_covered(const srb_ctx* c) {
  if (c->reg_kind == SRB_REG_TV) return true;
  if (tile_height(c) != 32) return false;
  return c->reg_kind == SRB_REG_TV3D || (c->reg_kind == SRB_REG_BTV && c->btv_R >= 1 && c->btv_R <= 4);
}

// BTV / 3-D TV term of units [unit_begin, unit_end): adds into the gradient rows the tile kernel has written.
inline srb_status reg_tile_launch(srb_ctx* c, const double* d_x, double* d_g, bool do_reg, int unit_begin, int unit_end,
                                  double* part_reg, bool* reg_done) {
  if (!do_reg || c->reg_kind == SRB_REG_TV || !fused_reg_covered(c)) return SRB_OK;
  const Geometry& G = c->g;
  RegTileParams R;
  R.H = G.H; R.W = G.W; R.Ca = c->Ca();
  R.row0 = c->reg_row0; R.row1 = c->reg_row1;
  R.unit_begin = unit_begin; R.tile_rows = tile_rows_per_channel(c);
  R.x = d_x; R.w = c->d_w; R.g = d_g;
  R.two_lambda = 2.0 * c->lambda;
  for (int i = 0; i < 9; ++i) R.decay[i] = i < (int)c->decay_h.size() ? c->decay_h[i] : 0.0;
  R.part_reg = part_reg;
  const dim3 grid((G.W + FT_W - 1) / FT_W, unit_end - unit_begin, 1);
  if (c->reg_kind == SRB_REG_TV3D) {
    k_tv3d_tile<<<grid, 256, 0, c->stream>>>(R);
  } else {
    switch (c->btv_R) {
      case 1: k_btv_tile<1><<<grid, 256, 0, c->stream>>>(R); break;
      case 2: k_btv_tile<2><<<grid, 256, 0, c->stream>>>(R); break;
      case 3: k_btv_tile<3><<<grid, 256, 0, c->stream>>>(R); break;
      default: k_btv_tile<4><<<grid, 256, 0, c->stream>>>(R); break;
    }
  }
  c->timing.kernel_launches += 1;
  *reg_done = true;
  return SRB_OK;
}

struct TileLayout {  // cost-partial slots of one evaluation
  size_t nblocks, nband;
  dim3 bgrid;
};
inline TileLayout tile_layout(const srb_ctx* c) {
  const TileState* st = tile_state(c);
  const Geometry& G = c->g;
  TileLayout L;
  L.nblocks = (size_t)((G.W + FT_W - 1) / FT_W) * tile_rows_per_channel(c) * c->Ca();
  const long long bcount = st->has_band ? st->band.count() : 0;
  const int band_frames = (st->band_groups > 0 && st->yz_valid) ? st->band_groups : G.N;  // merged groups or frames
  L.bgrid = dim3((unsigned)((bcount + 255) / 256), (unsigned)(band_frames * c->Ca()));
  L.nband = st->has_band ? (size_t)L.bgrid.x * L.bgrid.y : 0;
  return L;
}

// Data term (+ 2-D TV term when fused) of the (channel, tile row) units [unit_begin, unit_end) of
// the active channel range: writes their gradient rows and their cost partial sums.
inline srb_status fused_eval_units(srb_ctx* c, const double* d_x, double* d_g, bool do_reg, int unit_begin,
                                   int unit_end, bool* reg_done) {
  const TileState* st = tile_state(c);
  const Geometry& G = c->g;
  const int Ca = c->Ca();
  TileParams P;
  P.H = G.H; P.W = G.W; P.h = G.h; P.w = G.w; P.s = G.s; P.Ct = G.Ct; P.c0 = c->c0; P.Ca = Ca;
  P.sshift = -1;
  for (int b = 0; b < 4; ++b)
    if ((1 << b) == G.s) P.sshift = b;
  P.x = d_x; P.y = c->d_y; P.g = d_g; P.wts = c->d_w;
  P.entries = st->d_entries; P.phase_begin = st->d_phase_begin; P.num_entries = st->num_entries;
  P.qoff_min_r = st->qoff_min_r; P.qoff_max_r = st->qoff_max_r;
  P.qoff_min_c = st->qoff_min_c; P.qoff_max_c = st->qoff_max_c;
  P.lo_r = st->band.lo_r; P.hi_r = st->band.hi_r; P.lo_c = st->band.lo_c; P.hi_c = st->band.hi_c;
  for (int i = 0; i < 9; ++i) P.u[i] = i < G.K ? st->u[i] : 0.0, P.v[i] = i < G.K ? st->v[i] : 0.0;
  P.s2 = (double)G.s * G.s;
  P.two_s2 = 2.0 * P.s2;
  P.two_lambda = 2.0 * c->lambda;
  P.reg_fused = (do_reg && c->reg_kind == SRB_REG_TV) ? 1 : 0;
  P.row0 = c->reg_row0; P.row1 = c->reg_row1;
  *reg_done = P.reg_fused != 0;
  const int TH = tile_height(c);
  P.tile_rows = tile_rows_per_channel(c);
  static const bool no_table = getenv("SRB_NO_TABLE") != nullptr;  // A/B: generic residual pass everywhere
  P.fast = no_table ? nullptr : st->d_fast[TH == 64 ? 1 : 0];
  P.fast_y = st->d_fast_y[TH == 64 ? 1 : 0];
  P.fast_E = st->plan.fast_E;
  P.fast_items = (int)st->plan.fast[TH == 64 ? 1 : 0].size();
  P.yz = st->yz_valid ? st->d_yz : nullptr;
  P.unit_begin = unit_begin;
  const TileLayout L = tile_layout(c);
  const size_t need = 2 * L.nblocks + L.nband;
  if (need > c->partial_capacity) {
    if (c->d_partial) cudaFree(c->d_partial);
    c->d_partial = nullptr;
    c->partial_capacity = 0;
    if (cudaMalloc((void**)&c->d_partial, need * sizeof(double)) != cudaSuccess)
      return c->fail(SRB_ERR_NOMEM, "cudaMalloc failed (cost partials)");
    c->partial_capacity = need;
  }
  // layout: [data partials of the tiles][data partials of the band][reg partials of the tiles]
  P.part_data = c->d_partial;
  P.part_reg = c->d_partial + L.nblocks + L.nband;
  srb_status rc = SRB_OK;
  // kernel instantiation: PSF half width x fractional x tile height, and for integer shifts the number of
  // frames per sub-pixel phase the table-driven residual pass is specialised for (1, 2 or 4)
  const int fe = (!st->frac && P.fast != nullptr && (P.fast_E == 2 || P.fast_E == 4)) ? P.fast_E : 1;
  if (P.fast_E != fe) P.fast = nullptr;  // a kernel only ever sees the table it is specialised for
  P.yvar = nullptr;
  if (st->d_yzt != nullptr && st->yz_valid && TH == 32) {  // transposed Z layout: every tile, one code path
    srb_status zr = SRB_ERR_STATE;
    TileParams PZ = P;
    if (st->plan.zt_n > 1) {  // n frames of the same shift merged per phase: n ||A x - mean||^2 + constant
      PZ.s2 = P.s2 * st->plan.zt_n;
      PZ.two_s2 = P.two_s2 * st->plan.zt_n;
      PZ.yvar = st->d_yvar;
    }
    switch (st->KH) {
      case 1: zr = tile_launch_zt<1>(c, PZ, unit_end); break;
      case 2: zr = tile_launch_zt<2>(c, PZ, unit_end); break;
      case 3: zr = tile_launch_zt<3>(c, PZ, unit_end); break;
      case 4: zr = tile_launch_zt<4>(c, PZ, unit_end); break;
      default: break;
    }
    if (zr == SRB_OK) {
      c->timing.kernel_launches += 1;
      return reg_tile_launch(c, d_x, d_g, do_reg, unit_begin, unit_end, P.part_reg, reg_done);
    }
    if (zr != SRB_ERR_STATE) return zr;
    if (st->plan.zt_n > 1)   // the merged observations (and the merged border band) have no other kernel
      return c->fail(SRB_ERR_STATE, "tile kernel: the tensor maps of the transposed Z layout could not be built");
  }
  if (P.yz != nullptr && TH == 32 && fe == 1 && !st->frac) {  // row-major Z layout (SRB_ZT=0)
    srb_status zr = SRB_ERR_STATE;
    switch (st->KH * 2 + (st->yz_holes ? 1 : 0)) {
      case 2: zr = tile_launch_z<1, false>(c, P, unit_end); break;
      case 3: zr = tile_launch_z<1, true>(c, P, unit_end); break;
      case 4: zr = tile_launch_z<2, false>(c, P, unit_end); break;
      case 5: zr = tile_launch_z<2, true>(c, P, unit_end); break;
      case 6: zr = tile_launch_z<3, false>(c, P, unit_end); break;
      case 7: zr = tile_launch_z<3, true>(c, P, unit_end); break;
      case 8: zr = tile_launch_z<4, false>(c, P, unit_end); break;
      case 9: zr = tile_launch_z<4, true>(c, P, unit_end); break;
      default: break;
    }
    if (zr == SRB_OK) {
      c->timing.kernel_launches += 1;
      return reg_tile_launch(c, d_x, d_g, do_reg, unit_begin, unit_end, P.part_reg, reg_done);
    }
    if (zr != SRB_ERR_STATE) return zr;
  }
  const int key = ((st->KH * 2 + (st->frac ? 1 : 0)) * 2 + (TH == 64 ? 1 : 0)) * 3 + (fe == 1 ? 0 : fe == 2 ? 1 : 2);
  switch (key) {
#define SRB_TILE_CASE_FE(KH_, FR_, TH_, FE_, IDX_) \
    case (((KH_) * 2 + (FR_)) * 2 + ((TH_) == 64 ? 1 : 0)) * 3 + (IDX_): rc = tile_launch<KH_, (FR_) != 0, TH_, FE_>(c, P, unit_end); break;
#define SRB_TILE_CASE(KH_)                                                                          \
    SRB_TILE_CASE_FE(KH_, 0, 32, 1, 0) SRB_TILE_CASE_FE(KH_, 0, 32, 2, 1) SRB_TILE_CASE_FE(KH_, 0, 32, 4, 2) \
    SRB_TILE_CASE_FE(KH_, 0, 64, 1, 0) SRB_TILE_CASE_FE(KH_, 0, 64, 2, 1) SRB_TILE_CASE_FE(KH_, 0, 64, 4, 2) \
    SRB_TILE_CASE_FE(KH_, 1, 32, 1, 0) SRB_TILE_CASE_FE(KH_, 1, 64, 1, 0)
    SRB_TILE_CASE(0) SRB_TILE_CASE(1) SRB_TILE_CASE(2) SRB_TILE_CASE(3) SRB_TILE_CASE(4)
#undef SRB_TILE_CASE
#undef SRB_TILE_CASE_FE
    default: return c->fail(SRB_ERR_STATE, "tile kernel: unsupported PSF size");
  }
  if (rc != SRB_OK) return rc;
  c->timing.kernel_launches += 1;
  return reg_tile_launch(c, d_x, d_g, do_reg, unit_begin, unit_end, P.part_reg, reg_done);
}

// The border band (exact, reference order) of the whole image, or -- rows != NULL -- of one device's gradient rows.
inline srb_status fused_band(srb_ctx* c, const double* d_x, double* d_g, const BandRows* rows) {
  const TileState* st = tile_state(c);
  if (!st->has_band) return SRB_OK;
  const Geometry& G = c->g;
  const int Ca = c->Ca();
  const TileLayout L = tile_layout(c);
  GenericParams GP;
  GP.H = G.H; GP.W = G.W; GP.h = G.h; GP.w = G.w; GP.s = G.s; GP.K = G.K; GP.hk = G.hk;
  GP.N = G.N; GP.Ca = Ca; GP.Ct = G.Ct; GP.c0 = c->c0;
  GP.src_r = c->d_src_r; GP.src_c = c->d_src_c; GP.psf = c->d_psf;
  GP.rowY = c->d_rowY_fwd; GP.nX = c->d_nX_fwd;
  BandGroups M{0, 1, nullptr, nullptr};
  if (st->band_groups > 0 && st->yz_valid) M = BandGroups{st->band_groups, st->plan.zt_n, st->d_group_frame, st->d_yband};
  int sshift = -1;
  for (int b = 0; b < 5; ++b)
    if ((1 << b) == G.s) sshift = b;
  const BandRows whole{0, 0, Ca - 1, G.H, 0, 0};
  const BandRows RW = rows ? *rows : whole;
  k_band_forward<<<L.bgrid, 256, 0, c->stream>>>(GP, st->band, M, RW, d_x, c->d_y, st->d_pooled, c->d_partial + L.nblocks);
  c->timing.kernel_launches += 1;
  if (d_g) {
    GP.rowY = c->d_rowY_tr; GP.nX = c->d_nX_tr;
    const dim3 rgrid((unsigned)((st->reach.count() + 255) / 256), (unsigned)Ca);
    k_band_adjoint<<<rgrid, 256, 0, c->stream>>>(GP, st->band, st->reach, M, RW, sshift, st->d_pooled, d_g);
    c->timing.kernel_launches += 1;
  }
  return SRB_OK;
}

// HR rows of x around a gradient row that an evaluation reads: the PSF twice (forward and adjoint pass), one row for
// TV / R for BTV, and -- for models with a border band, whose samples are evaluated through the full warp -- the
// largest shift twice.
inline int stencil_halo_rows(const srb_ctx* c) {
  const TileState* st = tile_state(c);
  int reg = 1;
  if (c->reg_kind == SRB_REG_BTV && c->lambda > 0.0) reg = c->btv_R;
  int halo = 2 * c->g.hk + reg;
  if (st && st->has_band) halo += 2 * st->plan.max_shift + 1;
  return halo;
}

// The border band restricted to the gradient rows of units [u0, u1) (row-band partition): cost slots accumulate.
inline srb_status fused_band_units(srb_ctx* c, const double* d_x, double* d_g, int u0, int u1) {
  const TileState* st = tile_state(c);
  if (!st->has_band || u1 <= u0) return SRB_OK;
  const int tr = tile_rows_per_channel(c), TH = tile_height(c), H = c->g.H;
  BandRows RW;
  RW.ch_first = u0 / tr;
  RW.row_first = std::min(H, (u0 - RW.ch_first * tr) * TH);
  RW.ch_last = (u1 - 1) / tr;
  RW.row_last = std::min(H, (u1 - RW.ch_last * tr) * TH);
  RW.pad = c->g.hk + st->plan.max_shift + 1;
  RW.accumulate = 1;
  return fused_band(c, d_x, d_g, &RW);
}

// After every unit has been evaluated: the border band and the cost.
// Leaves the data cost in d_cost[0], the fused regularization cost in d_cost[1] and their sum in
// d_cost[2] (and *tail).  run_band = false: the band has been evaluated already (fused_band_units).
inline srb_status fused_eval_finish(srb_ctx* c, const double* d_x, double* d_g, double* tail, bool run_band = true) {
  const TileState* st = tile_state(c);
  const TileLayout L = tile_layout(c);
  if (run_band) {
    srb_status bs = fused_band(c, d_x, d_g, nullptr);
    if (bs != SRB_OK) return bs;
  }
  const size_t nd = L.nblocks + L.nband;
  if (nd + L.nblocks >= 32768) {  // large tile counts: a parallel first stage, then the closing block
    constexpr int SB = TileState::kStageBlocks;
    k_stage_partials<<<SB, 1024, 0, c->stream>>>(c->d_partial, nd, c->d_partial + nd, L.nblocks, st->d_stage);
    k_finish_partials<<<1, 1024, 0, c->stream>>>(st->d_stage, SB, st->d_stage + SB, SB, c->d_cost, tail);
    c->timing.kernel_launches += 2;
  } else {
    k_finish_partials<<<1, 1024, 0, c->stream>>>(c->d_partial, nd, c->d_partial + nd, L.nblocks, c->d_cost, tail);
    c->timing.kernel_launches += 1;
  }
  return SRB_OK;
}

inline srb_status fused_eval(srb_ctx* c, const double* d_x, double* d_g, bool do_reg, double* tail,
                             bool* reg_done) {
  const int units = tile_rows_per_channel(c) * c->Ca();
  srb_status st = fused_eval_units(c, d_x, d_g, do_reg, 0, units, reg_done);
  if (st != SRB_OK) return st;
  return fused_eval_finish(c, d_x, d_g, tail);
}


}  // namespace srb
