// api.cu - the C-ABI entry points that chain the kernels (include/ast_frontend.h).
#include <cmath>

#include "common.cuh"

#include <cstring>
#include <string>
#include <vector>

namespace ast {

// ---- diagnostic profiler ---------------------------------------------------------------------
struct ProfileRecord {
  const char* name;
  cudaEvent_t begin, end;
};
static bool g_profile_on = false;
static std::vector<ProfileRecord> g_profile;

bool profile_on() { return g_profile_on; }

void profile_mark(const char* name, cudaStream_t st, bool begin) {
  if (begin) {
    ProfileRecord r;
    r.name = name;
    cudaEventCreate(&r.begin);
    cudaEventCreate(&r.end);
    cudaEventRecord(r.begin, st);
    g_profile.push_back(r);
  } else {
    for (size_t i = g_profile.size(); i-- > 0;)
      if (g_profile[i].name == name) {
        cudaEventRecord(g_profile[i].end, st);
        break;
      }
  }
}

static int g_overlap_streams = 1;  // AST_OVERLAP=0: plain launch order STFT, decimator, CQT (no programmatic overlap of the STFT)
void set_overlap_streams(int on) { g_overlap_streams = on; }
static int g_stft_second = 1;   // AST_FEATURE_ORDER=dcs: round 2's earlier order decimator -> CQT -> STFT
void set_stft_second(int on) { g_stft_second = on; }

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Workspace {
  float2* stats_table;  // batch * 2 * 597 (mean, rstd): the CQT epilogue's table
  float4* stat4;        // batch * kStat4Stride: the STFT kernel's per-bin (mean_re, rstd_re, mean_im, rstd_im)
  float* octaves;       // batch * cqt_ws_clip_stride floats
  int* dec_flags;       // the decimator's per-tile completion counters
  size_t used;
};

static size_t stats_table_bytes(int batch) {
  const size_t rows = (size_t)(batch > 0 ? batch : 1);
  return align_up(sizeof(float2) * 2 * kFTotal * rows) + align_up(sizeof(float4) * kStat4Stride * rows);
}
static size_t octave_bytes(int batch, long long max_samples) {
  return align_up(sizeof(float) * (size_t)cqt_ws_clip_stride(max_samples) * (size_t)batch);
}

// the decimator's completion counters + one more int: the STFT's finished-CTA counter of the chained feature call
static size_t flag_bytes(int batch, long long max_samples) { return align_up(decimator_flag_bytes(batch, max_samples) + sizeof(int)); }
static int tail_counter_index(int batch, long long max_samples) { return (int)((decimator_flag_bytes(batch, max_samples) + 3) / 4); }

static int carve(void* ws, size_t ws_bytes, int batch, long long max_samples, Workspace* w) {
  const size_t need = stats_table_bytes(batch) + octave_bytes(batch, max_samples) + flag_bytes(batch, max_samples);
  if (!ws || ws_bytes < need)
    return fail(AST_ERR_WORKSPACE, "workspace of %zu bytes is too small, need %zu (ast_workspace_bytes)", ws_bytes, need);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(AST_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
  char* p = static_cast<char*>(ws);
  w->stats_table = reinterpret_cast<float2*>(p);
  w->stat4 = reinterpret_cast<float4*>(p + align_up(sizeof(float2) * 2 * kFTotal * (size_t)(batch > 0 ? batch : 1)));
  w->octaves = reinterpret_cast<float*>(p + stats_table_bytes(batch));
  w->dec_flags = reinterpret_cast<int*>(p + stats_table_bytes(batch) + octave_bytes(batch, max_samples));
  w->used = need;
  return AST_OK;
}

// ---- statistics pass: per-tile partial moments instead of features (stft.cu / cqt_tc.cu statistics mode) ------------
struct StatsScratch {
  float2* part_stft;   // batch * stft_tiles_max * 2 * 513
  float* part_n;       // batch * stft_tiles_max
  float2* part_cqt;    // batch * 7 * cqt_tiles * 4 * 24
  double* clip_stats;  // batch * 4 * 597
};
static int stats_stft_tiles_max(long long max_samples) {
  // the STFT kernel's statistics mode works in fixed tiles of 64 frame pairs = 128 frames (stft.cu)
  return (int)((num_frames(max_samples) + 127) / 128);
}
static int stats_cqt_tiles(long long max_samples) { return (num_frames(max_samples) + 127) / 128; }
static size_t stats_scratch_bytes(int batch, long long max_samples, StatsScratch* s, char* base) {
  const size_t b = (size_t)(batch > 0 ? batch : 1);
  const size_t n0 = align_up(sizeof(float2) * b * stats_stft_tiles_max(max_samples) * 2 * kFStft);
  const size_t n1 = align_up(sizeof(float) * b * stats_stft_tiles_max(max_samples));
  const size_t n2 = align_up(sizeof(float2) * b * kOctaves * stats_cqt_tiles(max_samples) * 4 * kCqtCols);
  const size_t n3 = align_up(sizeof(double) * 4 * kFTotal * b);
  if (s) {
    s->part_stft = reinterpret_cast<float2*>(base);
    s->part_n = reinterpret_cast<float*>(base + n0);
    s->part_cqt = reinterpret_cast<float2*>(base + n0 + n1);
    s->clip_stats = reinterpret_cast<double*>(base + n0 + n1 + n2);
  }
  return n0 + n1 + n2 + n3;
}

static int check_wave(const ast_plan* plan, const float* wave, int batch, long long max_samples, long long wave_stride) {
  if (!plan) return fail(AST_ERR_INVALID_ARG, "plan is null");
  if (batch < 0) return fail(AST_ERR_INVALID_ARG, "negative batch");
  if (batch > 65535) return fail(AST_ERR_INVALID_ARG, "batch %d exceeds 65535 clips per call", batch);
  if (batch > 0 && !wave) return fail(AST_ERR_INVALID_ARG, "wave is null");
  if (wave_stride < max_samples) return fail(AST_ERR_INVALID_ARG, "wave_stride %lld < max_samples %lld", wave_stride, max_samples);
  // torch.stft(center=True, pad_mode="reflect") raises when the pad (512) is not smaller than the input
  if (max_samples <= kNfft / 2)
    return fail(AST_ERR_TOO_SHORT, "clips of %lld samples are too short: reflect padding needs more than %d", max_samples,
                kNfft / 2);
  return AST_OK;
}

static OutSpec make_out(const ast_plan* plan, float* out, int layout, int dim1, int f_row, int f_off) {
  OutSpec o;
  o.out = out;
  o.layout = layout;
  o.dim1 = dim1;
  o.f_row = f_row;
  o.f_off = f_off;
  o.window = plan->cfg.window_size;
  o.step = plan->cfg.window_size - plan->cfg.overlap_frames;
  o.stats = nullptr;
  o.stats_clip_stride = 0;
  o.f_stats = kFTotal;
  o.stats_off = 0;
  o.cqt_part = nullptr;
  return o;
}
}  // namespace ast

using namespace ast;

extern "C" {

size_t ast_workspace_bytes(const ast_plan* plan, int32_t batch, int64_t max_samples) {
  (void)plan;
  if (batch < 0 || max_samples < 0) return 0;
  return stats_table_bytes(batch) + octave_bytes(batch, max_samples) + flag_bytes(batch, max_samples);
}

size_t ast_stats_workspace_bytes(const ast_plan* plan, int32_t batch, int64_t max_samples) {
  if (batch < 0 || max_samples < 0) return 0;
  // the feature call's scratch + per-tile partial moments + per-clip moments; the features themselves are never stored.
  // ast_stats_accumulate_features needs only the per-clip moments (the last term) and accepts this figure.
  return ast_workspace_bytes(plan, batch, max_samples) + stats_scratch_bytes(batch, max_samples, nullptr, nullptr);
}

int ast_stft_forward(const ast_plan* plan, const float* wave, const int32_t* lengths, int32_t batch, int64_t max_samples,
                     int64_t wave_stride, float* out, int32_t t_out, void* stream) {
  int rc = check_wave(plan, wave, batch, max_samples, wave_stride);
  if (rc != AST_OK) return rc;
  if (!out || t_out < 0) return fail(AST_ERR_INVALID_ARG, "ast_stft_forward: bad output");
  OutSpec o = make_out(plan, out, AST_LAYOUT_FLAT, t_out, kFStft, 0);
  return launch_stft(plan, wave, lengths, batch, max_samples, wave_stride, o, (cudaStream_t)stream);
}

int ast_cqt_forward(const ast_plan* plan, const float* wave, const int32_t* lengths, int32_t batch, int64_t max_samples,
                    int64_t wave_stride, void* workspace, size_t workspace_bytes, float* out, int32_t t_out, void* stream) {
  int rc = check_wave(plan, wave, batch, max_samples, wave_stride);
  if (rc != AST_OK) return rc;
  if (!out || t_out < 0) return fail(AST_ERR_INVALID_ARG, "ast_cqt_forward: bad output");
  Workspace w;
  rc = carve(workspace, workspace_bytes, batch, max_samples, &w);
  if (rc != AST_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long ws_stride = cqt_ws_clip_stride(max_samples);
  rc = launch_decimate_cascade(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, w.dec_flags, st);
  if (rc != AST_OK) return rc;
  OutSpec o = make_out(plan, out, AST_LAYOUT_FLAT, t_out, kFCqt, 0);  // 84-wide rows, CQT bin 0 at column 0
  return launch_cqt(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, use_tc_decimator() ? w.dec_flags : nullptr, o, st);
}

int ast_features_forward(const ast_plan* plan, const float* wave, const int32_t* lengths, int32_t batch,
                         int64_t max_samples, int64_t wave_stride, const float* mean, const float* std_,
                         int32_t stats_per_clip, float eps, void* workspace, size_t workspace_bytes, float* out,
                         int32_t dim1, int32_t layout, int32_t* n_sections, void* stream) {
  int rc = check_wave(plan, wave, batch, max_samples, wave_stride);
  if (rc != AST_OK) return rc;
  if (!out || dim1 < 0) return fail(AST_ERR_INVALID_ARG, "ast_features_forward: bad output");
  if (layout != AST_LAYOUT_FLAT && layout != AST_LAYOUT_SECTIONS) return fail(AST_ERR_INVALID_ARG, "unknown layout %d", layout);
  if ((mean == nullptr) != (std_ == nullptr)) return fail(AST_ERR_INVALID_ARG, "mean and std must both be given or both be null");
  if (layout == AST_LAYOUT_SECTIONS && lengths == nullptr &&
      num_sections(num_frames(max_samples), plan->cfg.window_size, plan->cfg.overlap_frames) == 0)
    return fail(AST_ERR_NO_SECTIONS, "%d frames give no section (need >= %d frames)", num_frames(max_samples),
                (plan->cfg.window_size + 1) / 2);
  Workspace w;
  rc = carve(workspace, workspace_bytes, batch, max_samples, &w);
  if (rc != AST_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const float2* table = mean ? w.stats_table : nullptr;
  const int n_stats = mean ? 2 * kFTotal * (stats_per_clip ? batch : 1) : 0;
  // default path: ONE small kernel prepares the statistics table, the section counts and the decimator's zeroed
  // completion counters, and the decimator is launched as its programmatic dependent
  const bool chained = g_overlap_streams && !profile_on() && use_tc_decimator() && use_tc_cqt() && batch > 0;
  if (chained) {
    rc = launch_features_prologue(mean, std_, eps, n_stats, w.stats_table, w.stat4, lengths, batch, max_samples, layout, dim1,
                                  plan->cfg.window_size, plan->cfg.overlap_frames, n_sections, w.dec_flags,
                                  tail_counter_index(batch, max_samples) + 1, st);
    if (rc != AST_OK) return rc;
  } else {
    if (mean) {
      rc = launch_prep_stats(mean, std_, eps, n_stats, w.stats_table, w.stat4, st);
      if (rc != AST_OK) return rc;
    }
    if (n_sections) {
      rc = launch_count_sections(lengths, batch, max_samples, layout, dim1, plan->cfg.window_size, plan->cfg.overlap_frames,
                                 n_sections, st);
      if (rc != AST_OK) return rc;
    }
  }
  OutSpec o = make_out(plan, out, layout, dim1, kFTotal, 0);
  o.stats = table;
  o.stats_clip_stride = stats_per_clip ? 2 * kFTotal : 0;
  o.stats_off = 0;
  OutSpec oq = o;
  oq.f_off = kFStft;
  oq.stats_off = kFStft;
  const long long ws_stride = cqt_ws_clip_stride(max_samples);
  const float4* stat4 = mean ? w.stat4 : nullptr;
  const int stat4_stride = stats_per_clip ? kStat4Stride : 0;
  if (!g_overlap_streams || profile_on()) {
    rc = launch_stft(plan, wave, lengths, batch, max_samples, wave_stride, o, st, 0, false, nullptr, stat4, stat4_stride);
    if (rc != AST_OK) return rc;
    rc = launch_decimate_cascade(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, w.dec_flags, st);
    if (rc != AST_OK) return rc;
    return launch_cqt(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, use_tc_decimator() ? w.dec_flags : nullptr, oq, st);
  }
  // One stream, three programmatic dependent launches behind the prologue ("Launch structure").
  // Default order: decimator cascade -> STFT -> CQT projection (below); AST_FEATURE_ORDER=dcs keeps the earlier order
  // decimator -> CQT projection -> STFT (further below).  When the call is not chained, the memset inside
  // launch_decimate_cascade is an ordinary stream operation, so nothing starts before the previous call has finished.
  rc = launch_decimate_cascade(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, w.dec_flags, st,
                               /*flags_zeroed=*/chained);
  if (rc != AST_OK) return rc;
  if (chained && g_stft_second) {
    // decimator -> STFT -> CQT projection.  The STFT waits for nothing (it reads the waveform and writes its own
    // columns): its small CTAs take the SMs over as the decimator's persistent CTAs retire (the 32 chain CTAs last), and
    // the projection's persistent CTAs (a whole SM each) start as the STFT's last wave drains, drawing their tiles from a
    // queue so that they end together.  (With the default build the STFT does not co-reside with the decimator: that
    // needs the full shared-memory carve-out and a 96-register decimator, was measured and is slower - DESIGN.md section 5.)
    // The STFT's last CTA waits for its programmatic primary (the decimator), every CQT CTA for its primary (the STFT)
    // before it exits: "the call's last kernel is complete" still means the whole call is.
    rc = launch_stft(plan, wave, lengths, batch, max_samples, wave_stride, o, st, 0, /*pdl=*/true,
                     reinterpret_cast<unsigned int*>(w.dec_flags + tail_counter_index(batch, max_samples)), stat4, stat4_stride);
    if (rc != AST_OK) return rc;
    return launch_cqt(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, w.dec_flags, oq, st,
                      /*tile_queue=*/true);
  }
  rc = launch_cqt(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride,
                  use_tc_decimator() ? w.dec_flags : nullptr, oq, st);
  if (rc != AST_OK) return rc;
  // the STFT never waits for the CQT projection it is a programmatic dependent of, so it could finish first; its last
  // CTA to finish then waits for that grid (tail counter), so that "the call's last kernel is complete" means the whole
  // call is complete - which a following programmatic dependent (the next call's prologue, the iSTFT) relies on
  return launch_stft(plan, wave, lengths, batch, max_samples, wave_stride, o, st, 0, /*pdl=*/chained,
                     chained ? reinterpret_cast<unsigned int*>(w.dec_flags + tail_counter_index(batch, max_samples)) : nullptr,
                     stat4, stat4_stride);
}

int ast_istft_forward(const ast_plan* plan, const float* spec, int32_t batch, int32_t dim1, int32_t f_in, int32_t layout,
                      int32_t overlap_frames, int32_t original_size, float* wave_out, int64_t out_stride, void* stream) {
  if (!plan) return fail(AST_ERR_INVALID_ARG, "plan is null");
  if (batch < 0 || batch > 65535 || dim1 <= 0 || (batch > 0 && !spec))
    return fail(AST_ERR_INVALID_ARG, "ast_istft_forward: bad argument");
  if (f_in < kFStft) return fail(AST_ERR_SHAPE, "spectrogram rows have %d columns, need at least %d", f_in, kFStft);
  int n_frames, window = plan->cfg.window_size;
  if (layout == AST_LAYOUT_FLAT) {
    n_frames = dim1;
  } else if (layout == AST_LAYOUT_SECTIONS) {
    if (overlap_frames < 0 || 2 * overlap_frames > window)
      return fail(AST_ERR_INVALID_ARG, "need 0 <= 2 * overlap <= window (got %d / %d)", overlap_frames, window);
    n_frames = (window - overlap_frames) * (dim1 - 1) + window;
  } else {
    return fail(AST_ERR_INVALID_ARG, "unknown layout %d", layout);
  }
  if (original_size > 0 && original_size < n_frames) n_frames = original_size;
  if (out_stride < (long long)kHop * (n_frames - 1)) return fail(AST_ERR_INVALID_ARG, "out_stride too small");
  if (batch > 0 && n_frames > 1 && !wave_out) return fail(AST_ERR_INVALID_ARG, "wave_out is null");
  return launch_istft(plan, spec, batch, dim1, f_in, layout, window, overlap_frames, n_frames, wave_out, out_stride,
                      (cudaStream_t)stream);
}

int ast_stats_accumulate_features(const ast_plan* plan, const float* feats, const int32_t* n_frames,
                                  const int32_t* group_ids, int32_t batch, int32_t t_dim, int32_t f_dim, void* workspace,
                                  size_t workspace_bytes, int32_t n_groups, double* acc, double* counts, void* stream) {
  if (!plan || !feats || !acc || !counts || batch < 0 || batch > 65535 || t_dim <= 0 || f_dim <= 0 || n_groups <= 0)
    return fail(AST_ERR_INVALID_ARG, "ast_stats_accumulate_features: bad argument");
  const size_t need = sizeof(double) * 4 * (size_t)f_dim * (size_t)batch;
  if (!workspace || workspace_bytes < need) return fail(AST_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
  double* clip_stats = static_cast<double*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_clip_stats(feats, n_frames, batch, t_dim, f_dim, clip_stats, st);
  if (rc != AST_OK) return rc;
  return launch_stats_accumulate(clip_stats, group_ids, batch, f_dim, n_groups, acc, counts, st);
}

int ast_stats_accumulate(const ast_plan* plan, const float* wave, const int32_t* lengths, const int32_t* group_ids,
                         int32_t batch, int64_t max_samples, int64_t wave_stride, void* workspace, size_t workspace_bytes,
                         int32_t n_groups, double* acc, double* counts, void* stream) {
  int rc = check_wave(plan, wave, batch, max_samples, wave_stride);
  if (rc != AST_OK) return rc;
  if (!acc || !counts || n_groups <= 0) return fail(AST_ERR_INVALID_ARG, "ast_stats_accumulate: bad argument");
  const size_t need = ast_stats_workspace_bytes(plan, batch, max_samples);
  if (!workspace || workspace_bytes < need) return fail(AST_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
  if (batch == 0) return AST_OK;
  const size_t base = ast_workspace_bytes(plan, batch, max_samples);
  Workspace w;
  rc = carve(workspace, base, batch, max_samples, &w);
  if (rc != AST_OK) return rc;
  StatsScratch sc;
  stats_scratch_bytes(batch, max_samples, &sc, static_cast<char*>(workspace) + base);
  cudaStream_t st = (cudaStream_t)stream;
  // The raw (un-normalised) features of compute_stats (compute_separated_stats.py:22-28) are never stored: the STFT and
  // the CQT projection run in statistics mode and leave per-tile (mean, M2) partials (~ 100 KB per clip against the
  // 4.1 MB of a flat feature tensor), which one small kernel merges per clip in frame order.
  const int t_dim = num_frames(max_samples);
  OutSpec o = make_out(plan, nullptr, AST_LAYOUT_FLAT, t_dim, kFTotal, 0);
  const int stft_tiles = stft_tiles_per_clip(plan, batch, t_dim, /*stats_mode=*/true);
  if (stft_tiles > stats_stft_tiles_max(max_samples)) return fail(AST_ERR_WORKSPACE, "internal: statistics tile bound exceeded");
  // same order as the feature call: decimator -> STFT (statistics mode; it waits for nothing: it reads the waveform and
  // writes its own partials) -> CQT projection; AST_STATS_ORDER_DCS keeps the earlier order with the STFT last.  The
  // finalise kernel is an ordinary launch: it starts when everything before it on the stream has completed.
  const long long ws_stride = cqt_ws_clip_stride(max_samples);
  rc = launch_decimate_cascade(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, w.dec_flags, st);
  if (rc != AST_OK) return rc;
  OutSpec oq = o;
  oq.f_off = kFStft;
  oq.stats_off = kFStft;
  oq.cqt_part = sc.part_cqt;
  const bool chained = g_overlap_streams && !profile_on() && use_tc_decimator();
  if (chained && g_stft_second && !getenv("AST_STATS_ORDER_DCS")) {
    // as in the feature call: decimator -> STFT (statistics mode) -> CQT projection drawing its tiles from the queue
    rc = launch_stft(plan, wave, lengths, batch, max_samples, wave_stride, o, st, 0, /*pdl=*/true, nullptr, nullptr, 0,
                     sc.part_stft, sc.part_n);
    if (rc != AST_OK) return rc;
    rc = launch_cqt(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride, w.dec_flags, oq, st,
                    /*tile_queue=*/true);
    if (rc != AST_OK) return rc;
  } else {
    rc = launch_cqt(plan, wave, lengths, batch, max_samples, wave_stride, w.octaves, ws_stride,
                    use_tc_decimator() ? w.dec_flags : nullptr, oq, st);
    if (rc != AST_OK) return rc;
    rc = launch_stft(plan, wave, lengths, batch, max_samples, wave_stride, o, st, 0, /*pdl=*/chained, nullptr, nullptr, 0,
                     sc.part_stft, sc.part_n);
    if (rc != AST_OK) return rc;
  }
  rc = launch_stats_finalize_clips(sc.part_stft, sc.part_n, stft_tiles, sc.part_cqt, stats_cqt_tiles(max_samples), lengths,
                                   max_samples, batch, sc.clip_stats, st);
  if (rc != AST_OK) return rc;
  return launch_stats_accumulate(sc.clip_stats, group_ids, batch, kFTotal, n_groups, acc, counts, st);
}

int ast_profile_enable(int32_t on) {
  for (auto& r : g_profile) {
    cudaEventDestroy(r.begin);
    cudaEventDestroy(r.end);
  }
  g_profile.clear();
  g_profile_on = on != 0;
  return AST_OK;
}

int ast_profile_collect(char* names, float* total_ms, int32_t* launches, int32_t capacity, int32_t* n_kernels) {
  if (!names || !total_ms || !launches || !n_kernels || capacity <= 0)
    return fail(AST_ERR_INVALID_ARG, "ast_profile_collect: bad argument");
  std::vector<std::string> keys;
  for (auto& r : g_profile) {
    AST_CUDA_TRY(cudaEventSynchronize(r.end));
    float ms = 0.f;
    AST_CUDA_TRY(cudaEventElapsedTime(&ms, r.begin, r.end));
    size_t k = 0;
    for (; k < keys.size(); ++k)
      if (keys[k] == r.name) break;
    if (k == keys.size()) {
      if ((int)k >= capacity) continue;
      keys.push_back(r.name);
      total_ms[k] = 0.f;
      launches[k] = 0;
      std::strncpy(names + 32 * k, r.name, 31);
      names[32 * k + 31] = 0;
    }
    total_ms[k] += ms;
    launches[k] += 1;
    cudaEventDestroy(r.begin);
    cudaEventDestroy(r.end);
  }
  g_profile.clear();
  *n_kernels = (int32_t)keys.size();
  return AST_OK;
}

}  // extern "C"
