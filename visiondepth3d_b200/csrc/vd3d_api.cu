// vd3d_api.cu -- host side of libvd3d.so: context, workspaces, the per-frame kernel
// sequence of render_sbs_3d / pixel_shift_cuda (core/render_3d.py:561-712,1227-1419),
// and the C ABI declared in include/vd3d.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "dibr_launch.h"

using namespace vd3d;

namespace {

std::string g_create_error;

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

constexpr int kSlots = 16;   // staging slots: two depth batches of up to kMaxDepthBatch frames in flight
constexpr int kClones = 2;   // depth engine instances (own activation buffers / stream), alternating per batch
constexpr int kMaxDepthBatch = 8;
constexpr int kJobs = 5;  // pct, subj(norm), quantile(d0), subj(d0), subj(shaped)
constexpr size_t kJobWords = 4096 + 4 * 4096 + 4 * 64 + 64 + 64;  // + count (padded)
constexpr size_t kBarWords = 64;

}  // namespace

struct vd3d_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  std::string err;
  uint64_t launches = 0;
  int use_graphs = 1;
  int stats_only = 0;  // advance temporal state only (exact multi-GPU sharding: SURVEY 8(e))
  int exact = 0;       // 1: one-kernel-per-op path with correctly rounded transcendentals (dibr_kernels.cu);
                       // 0: persistent stats kernel + fused render kernel (dibr_fast.cu)
  int stats_blocks = 0;
  // bumped whenever a device resource that captured graphs bake in moves or changes content (ensure() reallocations,
  // linspace axes, INTER_AREA tables, DOF kernel bank); run_frame_slot drops stale graphs
  uint64_t res_epoch = 0, fg_epoch = 0;
  int fast_dbg = 0;           // env VD3D_FAST_DEBUG: triage bits of the fast path (1 exact pow, 2 exact k_shift)
  int prof_depth_frames = 0;  // frames covered by the stage-2 (depth) samples since the last collect
  uint64_t dclone_wver = 0;
  unsigned* bar = nullptr;  // grid barrier counter of k_stats (inside jobwords: zeroed by begin_frame)

  DevState* st = nullptr;
  FrameScalars* fs = nullptr;
  vd3d_frame_info* info_pinned = nullptr;
  FrameScalars* fs_pinned = nullptr;
  DevState* st_pinned = nullptr;
  uint32_t* jobwords = nullptr;  // kJobs * kJobWords, zeroed every frame
  SelTarget* tgs = nullptr;      // kJobs * 4
  JobMem jm[kJobs];

  // linspace tables
  Buf xs, ys;
  int xs_n = 0, ys_n = 0;
  // planes
  Buf tdf, dn0, dn1, rgb_s, frameB, d, shift, e2, eyeL, eyeR, eyeL2, eyeR2;
  Buf in_frame[kSlots], in_depth[kSlots], out_dev[kSlots], in_rgbf, in_depthf;
  Buf dof_kern;
  // cv2 ResizeArea_ tables of the current non-integer eye fit (device) + the sizes they were built for
  // ([0]: the frame path, whose CUDA graphs bake the pointers in; [1]: the stage entry points vd3d_fit_eye ...)
  Buf area_tab[2];
  int at_key[2][4] = {};
  int at_t[2] = {0, 0};
  int tdf_w = 0, tdf_h = 0;
  int frame_parity = 0;
  cudaEvent_t ev_h2d[kSlots] = {}, ev_done[kSlots] = {}, ev_d2h[kSlots] = {};
  // optional per-stage device timing (bench.py roofline): event pairs on ctx->stream
  int prof = 0;
  std::vector<cudaEvent_t> prof_ev[4];  // stage -> [start, stop, start, stop, ...]
  std::vector<cudaEvent_t> prof_pool;
  // CUDA graphs of the per-frame kernel sequence, keyed by (staging slot, depth ping-pong parity)
  struct FrameGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t n_ctx = 0, n_depth = 0;
  } fg[kSlots][2];
  vd3d_render_params fg_rp;
  int fg_h = 0, fg_w = 0, fg_dch = 0, fg_warm = 0;
  void* fg_depth = nullptr;
  // depth stage of the depth+stereo clip: two engine clones (shared weights) on two streams so the
  // depth forwards of consecutive frames overlap each other and the DIBR kernels of the previous frame
  vd3d_depth* dclone[kClones] = {};
  vd3d_depth* dclone_parent = nullptr;
  cudaStream_t s_depth[kClones] = {};
  cudaEvent_t ev_depth[kClones] = {};
  struct DepthGraph {  // one graph per (engine instance, frames in the batch)
    cudaGraphExec_t exec = nullptr;
    uint64_t n = 0;
  } dg[kClones][kMaxDepthBatch + 1];
  int dg_warm = 0, dg_h = 0, dg_w = 0;
  int depth_batch = 4;  // frames per depth forward in vd3d_render_clip_depth (env VD3D_DEPTH_BATCH, 1..8)
  // dof kernel cache
  double dof_sigma_cached = -1.0;
  int dof_nlevels = 0, dof_ksize[8] = {0}, dof_koff[8] = {0}, dof_halo = 0;
};

extern "C" {
uint64_t vd3d_depth_weights_version(vd3d_depth* e);
static void drop_graphs(vd3d_ctx* ctx);
static void drop_depth_graphs(vd3d_ctx* ctx);
int vd3d_depth_clone(vd3d_depth* e, void* cuda_stream, vd3d_depth** out);
void vd3d_depth_destroy(vd3d_depth* e);
}

namespace {

#define CK(call)                                                                    \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      char _b[512];                                                                 \
      snprintf(_b, sizeof _b, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      ctx->err = _b;                                                                \
      return VD3D_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

cudaEvent_t prof_event(vd3d_ctx* ctx) {
  cudaEvent_t e;
  if (!ctx->prof_pool.empty()) {
    e = ctx->prof_pool.back();
    ctx->prof_pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
struct ProfScope {  // records start now, stop at destruction
  vd3d_ctx* c;
  int stage;
  ProfScope(vd3d_ctx* ctx, int st) : c(ctx), stage(st) {
    if (c->prof) {
      cudaEvent_t e = prof_event(c);
      cudaEventRecord(e, c->stream);
      c->prof_ev[stage].push_back(e);
    }
  }
  ~ProfScope() {
    if (c->prof) {
      cudaEvent_t e = prof_event(c);
      cudaEventRecord(e, c->stream);
      c->prof_ev[stage].push_back(e);
    }
  }
};

int fail(vd3d_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int ensure(vd3d_ctx* ctx, Buf& b, size_t bytes) {
  if (b.cap >= bytes) return VD3D_OK;
  if (b.p) {
    CK(cudaDeviceSynchronize());  // frames still in flight (other streams, graph replays) may read the old block
    CK(cudaFree(b.p));
  }
  b.p = nullptr;
  b.cap = 0;
  ctx->res_epoch++;
  CK(cudaMalloc(&b.p, bytes));
  b.cap = bytes;
  return VD3D_OK;
}

// torch.linspace(-1, 1, n) fp32: step=(end-start)/(n-1); i<n/2: fma(step,i,start) else fma(-step,n-1-i,end)
void linspace32(float start, float end, int n, std::vector<float>& out) {
  out.resize(n);
  if (n == 1) {
    out[0] = start;
    return;
  }
  float step = (end - start) / (float)(n - 1);
  for (int i = 0; i < n; ++i)
    out[i] = (i < n / 2) ? fmaf(step, (float)i, start) : fmaf(-step, (float)(n - 1 - i), end);
}

int ensure_axes(vd3d_ctx* ctx, int W, int H) {
  std::vector<float> v;
  if (ctx->xs_n != W || ctx->ys_n != H) CK(cudaDeviceSynchronize());  // queued frames may still read the old axes
  if (ctx->xs_n != W) {
    int r = ensure(ctx, ctx->xs, sizeof(float) * W);
    if (r) return r;
    linspace32(-1.f, 1.f, W, v);
    CK(cudaMemcpyAsync(ctx->xs.p, v.data(), sizeof(float) * W, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->xs_n = W;
    ctx->res_epoch++;
  }
  if (ctx->ys_n != H) {
    int r = ensure(ctx, ctx->ys, sizeof(float) * H);
    if (r) return r;
    linspace32(-1.f, 1.f, H, v);
    CK(cudaMemcpyAsync(ctx->ys.p, v.data(), sizeof(float) * H, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ys_n = H;
    ctx->res_epoch++;
  }
  return VD3D_OK;
}

// torch.quantile rank arithmetic: rank = fp32(q) * (n-1) in fp32
void quantile_rank(double q, long long n, uint32_t& lo, uint32_t& hi, float& w) {
  float rank = (float)q * (float)(n - 1);
  float fl = floorf(rank), ce = ceilf(rank);
  lo = (uint32_t)fl;
  hi = (uint32_t)ce;
  w = rank - fl;
}

SelJob make_job(const JobMem& m, const float* data, int W, int x0, int x1, int y0, int y1, int masked, int nt,
                int rank_from_count, bool hist64) {
  SelJob j;
  j.data = data;
  j.W = W;
  j.x0 = x0;
  j.x1 = x1;
  j.y0 = y0;
  j.y1 = y1;
  j.masked = masked;
  j.ntargets = nt;
  j.rank_from_count = rank_from_count;
  j.hist1 = m.hist1;
  j.hist2 = m.hist2;
  j.hist3 = m.hist3;
  j.hist64 = hist64 ? m.hist64 : nullptr;
  j.tg = m.tg;
  j.count = m.count;
  return j;
}

int sel_blocks(int rows) {
  int b = 148 * 4;
  return rows < b ? rows : b;
}

// ---------------------------------------------------------------------------
// pixel_shift core: from a depth plane [sh,sw] (already normalised) and an RGB source
// to the two u8 eyes.  Shared by vd3d_pixel_shift and vd3d_render_frame.
// ---------------------------------------------------------------------------
struct CoreIn {
  const float* depth;  // f32 [sh, sw]
  int sh, sw;
  const uint8_t* src_u8;  // identity path
  int src_pitch, cx0, cy0;
  const float* src_f32;  // RGB planes [3,H,W]
  int W, H;
  vd3d_shift_params p;
  int grade;
  float sat, con, bri;
  uint8_t* left;
  uint8_t* right;
};

int run_core(vd3d_ctx* ctx, const CoreIn& in) {
  cudaStream_t s = ctx->stream;
  const int W = in.W, H = in.H;
  int r;
  if ((r = ensure_axes(ctx, W, H))) return r;
  if ((r = ensure(ctx, ctx->d, sizeof(float) * (size_t)W * H))) return r;
  if ((r = ensure(ctx, ctx->shift, sizeof(float) * (size_t)W * H))) return r;
  const float* xs = (const float*)ctx->xs.p;
  const float* ys = (const float*)ctx->ys.p;
  float* d = (float*)ctx->d.p;
  float* shift = (float*)ctx->shift.p;

  launch_d0(in.depth, in.sh, in.sw, d, H, W, xs, ys, 0.08f, s);
  ctx->launches += 1;
  // quantiles (depth_stretch_lo/hi) over the full map + subject estimate on the centre crop
  uint32_t l0, l1, h0, h1;
  float wlo, whi;
  quantile_rank(in.p.depth_stretch_lo, (long long)W * H, l0, l1, wlo);
  quantile_rank(in.p.depth_stretch_hi, (long long)W * H, h0, h1, whi);
  launch_set_ranks(ctx->jm[2].tg, l0, l1, h0, h1, s);
  SelJob qj = make_job(ctx->jm[2], d, W, 0, W, 0, H, 0, 4, 0, false);
  SelJob s0 = make_job(ctx->jm[3], d, W, W / 5, W * 4 / 5, H / 5, H * 4 / 5, 1, 1, 1, true);
  launch_select(qj, &s0, sel_blocks(H), s);
  launch_fin_d0(qj, s0, wlo, whi, ctx->fs, s);
  launch_shape(d, W * H, ctx->fs, (float)in.p.depth_pop_mid, (float)in.p.depth_pop_gamma, s);
  SelJob s1 = make_job(ctx->jm[4], d, W, W / 5, W * 4 / 5, H / 5, H * 4 / 5, 1, 1, 1, true);
  launch_select(s1, nullptr, sel_blocks(H * 4 / 5 - H / 5), s);
  ShiftArgs sa;
  sa.p = in.p;
  sa.W = W;
  sa.H = H;
  launch_fin_shape(s1, sa, ctx->st, ctx->fs, s);
  if (ctx->stats_only) {  // every temporal state update of the frame has happened by now
    ctx->launches += 1 + 6 + 1 + 1 + 6 + 1;
    CK(cudaGetLastError());
    return VD3D_OK;
  }
  launch_shift(d, shift, H, W, ctx->fs, in.p.enable_edge_masking ? 1 : 0, (float)in.p.feather_strength, s);
  ctx->launches += 1 + 6 + 1 + 1 + 6 + 1 + 1;
  int feather = in.p.enable_feathering ? 1 : 0;
  if (feather) {
    if (in.p.blur_ksize < 1 || in.p.blur_ksize > 63) return fail(ctx, VD3D_ERR_UNSUPPORTED, "blur_ksize must be in [1,63]");
    if ((r = ensure(ctx, ctx->e2, sizeof(float2) * (size_t)W * H))) return r;
    launch_warp_edges(d, shift, (float2*)ctx->e2.p, H, W, xs, ys, (float)in.p.feather_strength, s);
    ctx->launches += 1;
  }
  ComposeArgs ca;
  ca.src_u8 = in.src_u8;
  ca.src_pitch = in.src_pitch;
  ca.cx0 = in.cx0;
  ca.cy0 = in.cy0;
  ca.src_f32 = in.src_f32;
  ca.shift = shift;
  ca.e2 = (const float2*)ctx->e2.p;
  ca.xs = xs;
  ca.ys = ys;
  ca.H = H;
  ca.W = W;
  ca.k = in.p.blur_ksize;
  ca.feather = feather;
  ca.grade = in.grade;
  ca.sat = in.sat;
  ca.con = in.con;
  ca.bri = in.bri;
  ca.left = in.left;
  ca.right = in.right;
  {
    ProfScope ps(ctx, 1);
    launch_compose(ca, s);
  }
  ctx->launches += 1;
  CK(cudaGetLastError());
  return VD3D_OK;
}

// ---------------------------------------------------------------------------
// fast path (dibr_fast.cu): k_stats -> k_shift<fast> -> k_render
// ---------------------------------------------------------------------------
struct FastLoop {  // loop-level inputs of k_stats (null for the pixel_shift_cuda entry)
  const IngestArgs* ia;
  float4* rgbx_s;
  float4* rgbx;
  float* dn;
  const float* dn_prev;
  const LoopArgs* la;
};
struct FastPost {  // fused bars + sharpen + fit + pack (fuse == 0: write the eyes)
  int fuse = 0;
  int sharpen = 0;
  float kc = 0.f, ke = 0.f;
  uint8_t* out = nullptr;
  int out_w = 0, per_eye_w = 0;
  const float4* src_rgbx = nullptr;
};

int run_core_fast(vd3d_ctx* ctx, const CoreIn& in, const FastLoop* lp, const FastPost& post) {
  cudaStream_t s = ctx->stream;
  const int W = in.W, H = in.H;
  int r;
  if ((r = ensure_axes(ctx, W, H))) return r;
  if ((r = ensure(ctx, ctx->d, sizeof(float) * (size_t)W * H))) return r;
  if ((r = ensure(ctx, ctx->shift, sizeof(float) * (size_t)W * H))) return r;
  const float* xs = (const float*)ctx->xs.p;
  const float* ys = (const float*)ctx->ys.p;
  float* d = (float*)ctx->d.p;
  float* shift = (float*)ctx->shift.p;
  int feather = in.p.enable_feathering ? 1 : 0;
  if (feather && (in.p.blur_ksize < 1 || in.p.blur_ksize > 63))
    return fail(ctx, VD3D_ERR_UNSUPPORTED, "blur_ksize must be in [1,63]");

  StatsArgs sa;
  memset(&sa, 0, sizeof sa);
  sa.loop = lp ? 1 : 0;
  if (lp) {
    sa.ia = *lp->ia;
    sa.rgbx_s = lp->rgbx_s;
    sa.rgbx = lp->rgbx;
    sa.dn = lp->dn;
    sa.dn_prev = lp->dn_prev;
    sa.la = *lp->la;
    quantile_rank(0.02, (long long)lp->ia->tw * lp->ia->th, sa.pct_rank[0], sa.pct_rank[1], sa.pct_wlo);
    quantile_rank(0.98, (long long)lp->ia->tw * lp->ia->th, sa.pct_rank[2], sa.pct_rank[3], sa.pct_whi);
  }
  sa.core_depth = in.depth;
  sa.sh = in.sh;
  sa.sw = in.sw;
  sa.d = d;
  sa.H = H;
  sa.W = W;
  sa.xs = xs;
  sa.ys = ys;
  sa.sa.p = in.p;
  sa.sa.W = W;
  sa.sa.H = H;
  quantile_rank(in.p.depth_stretch_lo, (long long)W * H, sa.q_rank[0], sa.q_rank[1], sa.q_wlo);
  quantile_rank(in.p.depth_stretch_hi, (long long)W * H, sa.q_rank[2], sa.q_rank[3], sa.q_whi);
  for (int j = 0; j < kJobs; ++j) sa.jm[j] = ctx->jm[j];
  sa.st = ctx->st;
  sa.fs = ctx->fs;
  sa.bar = ctx->bar;
  sa.dbg = ctx->fast_dbg;
  CK(launch_stats(sa, ctx->stats_blocks, s));
  ctx->launches += 1;
  if (ctx->stats_only) return VD3D_OK;
  if (ctx->fast_dbg & 2)
    launch_shift(d, shift, H, W, ctx->fs, in.p.enable_edge_masking ? 1 : 0, (float)in.p.feather_strength, s);
  else
    launch_shift_fast(d, shift, H, W, ctx->fs, in.p.enable_edge_masking ? 1 : 0, (float)in.p.feather_strength, s);
  ctx->launches += 1;

  ComposeArgs ca;
  memset(&ca, 0, sizeof ca);
  ca.src_u8 = in.src_u8;
  ca.src_pitch = in.src_pitch;
  ca.cx0 = in.cx0;
  ca.cy0 = in.cy0;
  ca.src_f32 = in.src_f32;
  ca.shift = shift;
  ca.xs = xs;
  ca.ys = ys;
  ca.H = H;
  ca.W = W;
  ca.k = in.p.blur_ksize;
  ca.feather = feather;
  ca.grade = in.grade;
  ca.sat = in.sat;
  ca.con = in.con;
  ca.bri = in.bri;
  ca.left = in.left;
  ca.right = in.right;
  if (render_supports(feather, in.p.blur_ksize)) {
    RenderArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.c = ca;
    ra.src_rgbx = post.src_rgbx;
    ra.d = d;
    ra.feather_strength = (float)in.p.feather_strength;
    ra.fuse = post.fuse;
    ra.fs = ctx->fs;
    ra.sharpen = post.sharpen;
    ra.kc = post.kc;
    ra.ke = post.ke;
    ra.out = post.out;
    ra.out_w = post.out_w;
    ra.per_eye_w = post.per_eye_w;
    {
      ProfScope ps(ctx, 1);
      CK(launch_render(ra, s));
    }
    ctx->launches += 1;
  } else {
    // even / large box sizes: the generic kernels of the exact path (callers never request fuse for these)
    if (post.src_rgbx) return fail(ctx, VD3D_ERR_STATE, "generic compose needs planar RGB");
    if (feather) {
      if ((r = ensure(ctx, ctx->e2, sizeof(float2) * (size_t)W * H))) return r;
      launch_warp_edges(d, shift, (float2*)ctx->e2.p, H, W, xs, ys, (float)in.p.feather_strength, s);
      ctx->launches += 1;
    }
    ca.e2 = (const float2*)ctx->e2.p;
    {
      ProfScope ps(ctx, 1);
      launch_compose(ca, s);
    }
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return VD3D_OK;
}

int begin_frame(vd3d_ctx* ctx) {
  CK(cudaMemsetAsync(ctx->jobwords, 0, sizeof(uint32_t) * (kJobs * kJobWords + kBarWords), ctx->stream));
  CK(cudaMemsetAsync(ctx->fs, 0, sizeof(FrameScalars), ctx->stream));
  return VD3D_OK;
}

int fetch_info(vd3d_ctx* ctx, vd3d_frame_info* info) {
  CK(cudaMemcpyAsync(ctx->fs_pinned, ctx->fs, sizeof(FrameScalars), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->st_pinned, ctx->st, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const FrameScalars& f = *ctx->fs_pinned;
  const DevState& t = *ctx->st_pinned;
  info->pct_lo = t.pct_lo;
  info->pct_hi = t.pct_hi;
  info->subj_raw = f.subj_raw;
  info->stretch_lo = f.st_lo;
  info->stretch_hi = f.st_hi;
  info->subj_shaped = f.subj;
  info->subj_norm = f.subj_norm;
  info->dyn_scale = f.dyn;
  info->fg = f.fg;
  info->mg = f.mg;
  info->bg = f.bg;
  info->zero_parallax_offset = f.zpo;
  info->focal_depth = f.focal;
  info->motion_metric = f.motion;
  info->stable_zero = f.stable_zero;
  info->bar_width = f.bar_width;
  info->bar_side = f.bar_side;
  return VD3D_OK;
}

// torchvision _get_gaussian_kernel2d for the DOF levels (core/render_3d.py:798-806)
int ensure_dof_kernels(vd3d_ctx* ctx, double max_sigma, int num_levels) {
  if (ctx->dof_sigma_cached == max_sigma && ctx->dof_nlevels == num_levels) return VD3D_OK;
  std::vector<float> sig;
  linspace32(0.f, (float)max_sigma, num_levels, sig);
  std::vector<float> all;
  int halo = 0;
  for (int l = 0; l < num_levels; ++l) {
    float s = sig[l];
    int ks = 1;
    if ((double)s != 0.0) ks = (int)(2 * ceil(2 * (double)s) + 1);
    if (ks > 17) return fail(ctx, VD3D_ERR_UNSUPPORTED, "dof_strength too large (Gaussian ksize > 17)");
    ctx->dof_ksize[l] = ks;
    ctx->dof_koff[l] = (int)all.size();
    if (ks > 1) {
      std::vector<float> x, pdf(ks);
      float half = (float)((ks - 1) * 0.5);
      linspace32(-half, half, ks, x);
      float sum = 0.f;
      for (int i = 0; i < ks; ++i) {
        float q = x[i] / (float)(double)s;
        float e = -0.5f * (q * q);
        pdf[i] = (float)exp((double)e);
      }
      for (int i = 0; i < ks; ++i) sum += pdf[i];
      for (int i = 0; i < ks; ++i) pdf[i] = pdf[i] / sum;
      for (int y = 0; y < ks; ++y)
        for (int xx = 0; xx < ks; ++xx) all.push_back(pdf[y] * pdf[xx]);
      if (ks / 2 > halo) halo = ks / 2;
    }
  }
  if (all.empty()) all.push_back(1.f);
  CK(cudaDeviceSynchronize());
  int r = ensure(ctx, ctx->dof_kern, sizeof(float) * all.size());
  if (r) return r;
  CK(cudaMemcpyAsync(ctx->dof_kern.p, all.data(), sizeof(float) * all.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->res_epoch++;
  ctx->dof_sigma_cached = max_sigma;
  ctx->dof_nlevels = num_levels;
  ctx->dof_halo = halo;
  return VD3D_OK;
}

void sharpen_coeffs(double factor, float& kc, float& ke) {
  // apply_sharpening: fp32 kernel [[0,-1,0],[-1,5+f,-1],[0,-1,0]] / sum (core/render_3d.py:719-728)
  float c = (float)(5.0 + factor);
  float ks = c - 4.0f;
  kc = c;
  ke = -1.0f;
  if (ks != 0.f) {
    kc = c / ks;
    ke = -1.0f / ks;
  }
}

struct FitPlan {
  int fit_x0, fit_y0, fit_w, fit_h, sx, sy;
  // non-integer INTER_AREA shrink: device tables (null otherwise), see PostArgs
  const int *xofs = nullptr, *xcnt = nullptr, *yofs = nullptr, *ycnt = nullptr;
  const float *xal = nullptr, *yal = nullptr;
  int area_t = 0;
  int lin = 0;  // enlarged axis: cv2's fixed-point bilinear emulation of INTER_AREA (opt-in, see plan_fit)
};

// cv2's computeResizeAreaTab (imgproc/src/resize.cpp, opencv 4.13 as installed with the reference): geometry in double,
// one fp32 weight per (destination, source) pair; the sources of one destination index are consecutive.
void area_axis_tab(int ssize, int dsize, std::vector<int>& ofs, std::vector<int>& cnt, std::vector<std::vector<float>>& al) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  ofs.assign(dsize, 0);
  cnt.assign(dsize, 0);
  al.assign(dsize, {});
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = fmin(scale, (double)ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    int first = -1;
    auto push = [&](int si, float a) {
      if (first < 0) first = si;
      al[dx].push_back(a);
    };
    if (sx1 - fsx1 > 1e-3) push(sx1 - 1, (float)((sx1 - fsx1) / cell));
    for (int sx = sx1; sx < sx2; ++sx) push(sx, (float)(1.0 / cell));
    if (fsx2 - sx2 > 1e-3) push(sx2, (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell));
    ofs[dx] = first < 0 ? 0 : first;
    cnt[dx] = (int)al[dx].size();
  }
}

// cv2 resize(): the "area_mode" coefficients of the bilinear scheme INTER_AREA falls back to when an axis grows
// (generic path of imgproc/src/resize.cpp): source index + two weights, saturate_cast<short>(c * 2048)
void area_linear_tab(int ssize, int dsize, std::vector<int>& ofs, std::vector<int>& a01) {
  const double inv = (double)dsize / (double)ssize, scale = 1.0 / inv;
  ofs.assign(dsize, 0);
  a01.assign((size_t)2 * dsize, 0);
  for (int dx = 0; dx < dsize; ++dx) {
    int sx = (int)floor(dx * scale);
    float fx = (float)((dx + 1) - (sx + 1) * inv);
    fx = fx <= 0 ? 0.f : fx - floorf(fx);
    if (sx < 0) {
      fx = 0.f;
      sx = 0;
    }
    if (sx >= ssize - 1) {
      fx = 0.f;
      sx = ssize - 1;
    }
    const float c0 = 1.f - fx;
    ofs[dx] = sx;
    a01[2 * dx] = (int)lrintf(c0 * 2048.f);
    a01[2 * dx + 1] = (int)lrintf(fx * 2048.f);
  }
}

// build (or reuse) the device tables for a W x H -> nw x nh non-integer INTER_AREA shrink
int ensure_area_tabs(vd3d_ctx* ctx, int which, int W, int H, int nw, int nh, FitPlan& f, bool lin = false) {
  Buf& tab = ctx->area_tab[which];
  int* key = ctx->at_key[which];
  const bool have = key[0] == W && key[1] == H && key[2] == nw && key[3] == nh && tab.p;
  if (!have) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) return fail(ctx, VD3D_ERR_STATE, "INTER_AREA tables missing during capture");
    std::vector<int> xo, xc, yo, yc;
    std::vector<std::vector<float>> xa, ya;
    if (lin) {
      std::vector<int> ax, ay;
      area_linear_tab(W, nw, xo, ax);
      area_linear_tab(H, nh, yo, ay);
      xc.assign(nw, 2);
      yc.assign(nh, 2);
      xa.resize(nw);
      ya.resize(nh);
      for (int i = 0; i < nw; ++i) xa[i] = {(float)ax[2 * i], (float)ax[2 * i + 1]};
      for (int i = 0; i < nh; ++i) ya[i] = {(float)ay[2 * i], (float)ay[2 * i + 1]};
    } else {
      area_axis_tab(W, nw, xo, xc, xa);
      area_axis_tab(H, nh, yo, yc, ya);
    }
    int T = lin ? 2 : 1;
    for (int c : xc) T = c > T ? c : T;
    for (int c : yc) T = c > T ? c : T;
    if (T > 64) return fail(ctx, VD3D_ERR_UNSUPPORTED, "INTER_AREA shrink factor too large");
    const size_t nint = (size_t)2 * nw + (size_t)2 * nh, nflt = ((size_t)nw + nh) * T;
    std::vector<unsigned char> host(nint * 4 + nflt * 4);
    int* ip = (int*)host.data();
    float* fp = (float*)(host.data() + nint * 4);
    memcpy(ip, xo.data(), (size_t)nw * 4);
    memcpy(ip + nw, xc.data(), (size_t)nw * 4);
    memcpy(ip + 2 * nw, yo.data(), (size_t)nh * 4);
    memcpy(ip + 2 * nw + nh, yc.data(), (size_t)nh * 4);
    for (int i = 0; i < nw; ++i)
      for (int k = 0; k < T; ++k) fp[(size_t)i * T + k] = k < xc[i] ? xa[i][k] : 0.f;
    for (int i = 0; i < nh; ++i)
      for (int k = 0; k < T; ++k) fp[((size_t)nw + i) * T + k] = k < yc[i] ? ya[i][k] : 0.f;
    // frames already queued may still read the previous tables
    CK(cudaStreamSynchronize(ctx->stream));
    int r = ensure(ctx, tab, host.size());
    if (r) return r;
    CK(cudaMemcpy(tab.p, host.data(), host.size(), cudaMemcpyHostToDevice));
    key[0] = W;
    key[1] = H;
    key[2] = nw;
    key[3] = nh;
    ctx->at_t[which] = T;
    ctx->res_epoch++;
  }
  const int* ip = (const int*)tab.p;
  f.xofs = ip;
  f.xcnt = ip + nw;
  f.yofs = ip + 2 * nw;
  f.ycnt = ip + 2 * nw + nh;
  f.xal = (const float*)(ip + 2 * nw + 2 * nh);
  f.yal = f.xal + (size_t)nw * ctx->at_t[which];
  f.area_t = ctx->at_t[which];
  f.lin = lin ? 1 : 0;
  return VD3D_OK;
}

// eye fit: cv2.resize INTER_AREA (Half-SBS) or pad_to_aspect_ratio (core/render_3d.py:101-131,1409-1417)
int plan_fit(vd3d_ctx* ctx, int fmt, int W, int H, int pw, int ph, FitPlan& f, int tab_slot = 0) {
  int nw, nh;
  if (fmt == VD3D_FMT_HALF_SBS) {
    nw = pw;
    nh = ph;
    f.fit_x0 = f.fit_y0 = 0;
  } else {
    double ta = (double)pw / (double)ph;
    double ca = (double)W / (double)H;
    if (ca > ta) {
      nw = pw;
      nh = (int)(pw / ca);
    } else {
      nh = ph;
      nw = (int)(ca * ph);
    }
    f.fit_x0 = (pw - nw) / 2;
    f.fit_y0 = (ph - nh) / 2;
  }
  if (nw <= 0 || nh <= 0) return fail(ctx, VD3D_ERR_ARG, "degenerate eye size");
  f.fit_w = nw;
  f.fit_h = nh;
  if (W % nw == 0 && H % nh == 0) {  // identity or cv2's integer "area fast" path
    f.sx = W / nw;
    f.sy = H / nh;
    return VD3D_OK;
  }
  if (nw > W || nh > H) {
    // cv2 switches INTER_AREA to a fixed-point bilinear scheme as soon as one axis grows (e.g. the hard-coded
    // 1920x1080 Full-SBS eyes for sources below 1080p, core/render_3d.py:1121).  VD3D_FIT_ENLARGE=0 rejects these.
    static int enlarge = -1;
    if (enlarge < 0) {
      const char* v = getenv("VD3D_FIT_ENLARGE");
      enlarge = v ? atoi(v) : 1;
    }
    if (!enlarge) return fail(ctx, VD3D_ERR_UNSUPPORTED, "eye fit would enlarge the eye (INTER_AREA upscaling)");
    f.sx = f.sy = 0;
    // the key of the cached tables does not encode the mode: W x H -> nw x nh is either a shrink or an enlargement
    return ensure_area_tabs(ctx, tab_slot, W, H, nw, nh, f, true);
  }
  f.sx = f.sy = 0;
  return ensure_area_tabs(ctx, tab_slot, W, H, nw, nh, f);
}

void set_fit(PostArgs& pa, const FitPlan& fp) {
  pa.fit_x0 = fp.fit_x0;
  pa.fit_y0 = fp.fit_y0;
  pa.fit_w = fp.fit_w;
  pa.fit_h = fp.fit_h;
  pa.sx = fp.sx;
  pa.sy = fp.sy;
  pa.inv_area = fp.sx > 0 ? (float)(1.0 / (double)(fp.sx * fp.sy)) : 1.f;
  pa.xofs = fp.xofs;
  pa.xcnt = fp.xcnt;
  pa.xal = fp.xal;
  pa.yofs = fp.yofs;
  pa.ycnt = fp.ycnt;
  pa.yal = fp.yal;
  pa.area_t = fp.area_t;
  pa.lin = fp.lin;
}

int copy_in(vd3d_ctx* ctx, Buf& b, const void* src, size_t bytes, int mem, cudaStream_t s, const void** dev) {
  if (mem == VD3D_MEM_DEVICE) {
    *dev = src;
    return VD3D_OK;
  }
  int r = ensure(ctx, b, bytes);
  if (r) return r;
  CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
  *dev = b.p;
  return VD3D_OK;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int vd3d_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(vd3d_shift_params);
    case 1: return (int)sizeof(vd3d_render_params);
    case 2: return (int)sizeof(vd3d_size_plan);
    case 3: return (int)sizeof(vd3d_frame_info);
    default: return -1;
  }
}

int vd3d_create(int device, vd3d_ctx** out) {
  if (!out) return VD3D_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " (libvd3d has no CPU fallback)";
    return VD3D_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    g_create_error = "bad device index";
    return VD3D_ERR_ARG;
  }
  vd3d_ctx* ctx = new vd3d_ctx();
  ctx->device = device;
  auto bail = [&](const char* what, cudaError_t er) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(er);
    delete ctx;
    return VD3D_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  for (int i = 0; i < kSlots; ++i) {
    cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming);
  }
  if ((e = cudaMalloc(&ctx->st, sizeof(DevState))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->fs, sizeof(FrameScalars))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->jobwords, sizeof(uint32_t) * (kJobs * kJobWords + kBarWords))) != cudaSuccess)
    return bail("cudaMalloc", e);
  ctx->bar = ctx->jobwords + kJobs * kJobWords;
  if ((e = cudaMalloc(&ctx->tgs, sizeof(SelTarget) * kJobs * 4)) != cudaSuccess) return bail("cudaMalloc", e);
  cudaMemset(ctx->st, 0, sizeof(DevState));
  cudaMemset(ctx->fs, 0, sizeof(FrameScalars));
  cudaMemset(ctx->jobwords, 0, sizeof(uint32_t) * (kJobs * kJobWords + kBarWords));
  cudaMemset(ctx->tgs, 0, sizeof(SelTarget) * kJobs * 4);
  for (int j = 0; j < kJobs; ++j) {
    uint32_t* b = ctx->jobwords + (size_t)j * kJobWords;
    ctx->jm[j].hist1 = b;
    ctx->jm[j].hist2 = b + 4096;
    ctx->jm[j].hist3 = b + 4096 + 4 * 4096;
    ctx->jm[j].hist64 = b + 4096 + 4 * 4096 + 4 * 64;
    ctx->jm[j].count = b + 4096 + 4 * 4096 + 4 * 64 + 64;
    ctx->jm[j].tg = ctx->tgs + j * 4;
  }
  if ((e = cudaMallocHost(&ctx->fs_pinned, sizeof(FrameScalars))) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = cudaMallocHost(&ctx->st_pinned, sizeof(DevState))) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = init_kernel_attributes()) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
  if ((e = stats_grid(device, &ctx->stats_blocks)) != cudaSuccess) return bail("k_stats occupancy", e);
  if (const char* v = getenv("VD3D_STATS_BLOCKS")) {  // tuning: fewer CTAs leave SMs to the depth kernels of the next batch
    int n = atoi(v);
    if (n >= 8 && n < ctx->stats_blocks) ctx->stats_blocks = n;
  }
  {
    const char* v = getenv("VD3D_EXACT");
    ctx->exact = (v && atoi(v)) ? 1 : 0;
    if ((v = getenv("VD3D_DEPTH_BATCH"))) ctx->depth_batch = atoi(v);
    if ((v = getenv("VD3D_FAST_DEBUG"))) ctx->fast_dbg = atoi(v);
    if (ctx->depth_batch < 1) ctx->depth_batch = 1;
    if (ctx->depth_batch > kMaxDepthBatch) ctx->depth_batch = kMaxDepthBatch;
  }
  *out = ctx;
  return VD3D_OK;
}

void vd3d_destroy(vd3d_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  Buf* bufs[] = {&ctx->xs,    &ctx->ys,    &ctx->tdf,   &ctx->dn0,   &ctx->dn1,      &ctx->rgb_s,
                 &ctx->frameB, &ctx->d,     &ctx->shift, &ctx->e2,    &ctx->eyeL,     &ctx->eyeR,
                 &ctx->eyeL2, &ctx->eyeR2, &ctx->in_rgbf, &ctx->in_depthf, &ctx->dof_kern,
                 &ctx->area_tab[0], &ctx->area_tab[1]};
  for (Buf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (int i = 0; i < kSlots; ++i) {
    if (ctx->in_frame[i].p) cudaFree(ctx->in_frame[i].p);
    if (ctx->in_depth[i].p) cudaFree(ctx->in_depth[i].p);
    if (ctx->out_dev[i].p) cudaFree(ctx->out_dev[i].p);
  }
  drop_graphs(ctx);
  drop_depth_graphs(ctx);
  for (int i = 0; i < kClones; ++i) {
    if (ctx->dclone[i]) vd3d_depth_destroy(ctx->dclone[i]);
    if (ctx->s_depth[i]) cudaStreamDestroy(ctx->s_depth[i]);
    if (ctx->ev_depth[i]) cudaEventDestroy(ctx->ev_depth[i]);
  }
  cudaFree(ctx->st);
  cudaFree(ctx->fs);
  cudaFree(ctx->jobwords);
  cudaFree(ctx->tgs);
  cudaFreeHost(ctx->fs_pinned);
  cudaFreeHost(ctx->st_pinned);
  for (int i = 0; i < kSlots; ++i) {
    cudaEventDestroy(ctx->ev_h2d[i]);
    cudaEventDestroy(ctx->ev_done[i]);
    cudaEventDestroy(ctx->ev_d2h[i]);
  }
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->s_h2d);
  cudaStreamDestroy(ctx->s_d2h);
  delete ctx;
}

const char* vd3d_last_error(vd3d_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int vd3d_reset_state(vd3d_ctx* ctx, uint32_t which) {
  if (!ctx) return VD3D_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  DevState h;
  CK(cudaMemcpy(&h, ctx->st, sizeof h, cudaMemcpyDeviceToHost));
  if (which & VD3D_STATE_GLOBAL) {
    h.pct_lo = h.pct_hi = 0.f;
    h.pct_init = 0;
    h.conv_val = 0.0;
    h.conv_init = 0;
    h.fw_prev = 0.0;
    h.fw_count = 0;
    h.bar_prev = 0;
  }
  if (which & VD3D_STATE_CLIP) {
    h.tdf_init = 0;
    h.sm_fg = h.sm_mg = h.sm_bg = 0.0;
    h.sm_init = 0;
    h.focal = 0.0;
    h.focal_alpha = 0.15;
    h.focal_init = 0;
    h.have_prev_depth = 0;
    ctx->frame_parity = 0;
  }
  CK(cudaMemcpy(ctx->st, &h, sizeof h, cudaMemcpyHostToDevice));
  return VD3D_OK;
}

void* vd3d_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void vd3d_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
void* vd3d_stream(vd3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int vd3d_sync(vd3d_ctx* ctx) {
  if (!ctx) return VD3D_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(ctx->s_d2h));
  return VD3D_OK;
}
uint64_t vd3d_launch_count(vd3d_ctx* ctx) { return ctx ? ctx->launches : 0; }
int vd3d_profile(vd3d_ctx* ctx, int enable) {
  if (!ctx) return VD3D_ERR_ARG;
  ctx->prof = enable;
  return VD3D_OK;
}
int vd3d_profile_collect(vd3d_ctx* ctx, int stage, double* total_ms, int* count) {
  if (!ctx || stage < 0 || stage >= 4 || !total_ms || !count) return VD3D_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  double t = 0;
  int n = 0;
  auto& v = ctx->prof_ev[stage];
  for (size_t i = 0; i + 1 < v.size(); i += 2) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
    t += ms;
    ++n;
  }
  for (cudaEvent_t e : v) ctx->prof_pool.push_back(e);
  v.clear();
  *total_ms = t;
  *count = n;
  if (stage == 2) {  // depth samples cover whole batches: report frames so that total / count is per frame
    if (ctx->prof_depth_frames > 0) *count = ctx->prof_depth_frames;
    ctx->prof_depth_frames = 0;
  }
  return VD3D_OK;
}
// 1: the one-kernel-per-op path with correctly rounded transcendentals (bit-for-bit with oracle/dibr.py);
// 0 (default; env VD3D_EXACT=1 flips the default): persistent stats kernel + fused render kernel
int vd3d_set_exact(vd3d_ctx* ctx, int enable) {
  if (!ctx) return VD3D_ERR_ARG;
  if (ctx->exact != (enable ? 1 : 0)) {
    cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    ctx->exact = enable ? 1 : 0;
  }
  return VD3D_OK;
}
int vd3d_get_exact(vd3d_ctx* ctx) { return ctx ? ctx->exact : -1; }
// 1 while frame graphs are enabled (0 after vd3d_set_graphs(0) or after a failed capture fell back to eager launches)
int vd3d_graphs_active(vd3d_ctx* ctx) { return ctx ? ctx->use_graphs : -1; }
int vd3d_set_graphs(vd3d_ctx* ctx, int enable) {
  if (!ctx) return VD3D_ERR_ARG;
  ctx->use_graphs = enable;
  if (!enable) drop_graphs(ctx);
  return VD3D_OK;
}

// ---------------------------------------------------------------------------
int vd3d_pixel_shift(vd3d_ctx* ctx, const float* rgb, const float* depth, int in_h, int in_w, int width,
                     int height, const vd3d_shift_params* p, uint8_t* left_bgr, uint8_t* right_bgr, float* shift,
                     int mem, vd3d_frame_info* info) {
  if (!ctx || !rgb || !depth || !p || !left_bgr || !right_bgr) return fail(ctx, VD3D_ERR_ARG, "null argument");
  if (in_h < 2 || in_w < 2 || width < 2 || height < 2) return fail(ctx, VD3D_ERR_ARG, "image too small");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int W = width, H = height;
  int r;
  const bool fastp = !ctx->exact && render_supports(p->enable_feathering ? 1 : 0, p->blur_ksize);
  if ((r = begin_frame(ctx))) return r;
  if (!fastp) launch_set_shifts(ctx->fs, p->fg_shift, p->mg_shift, p->bg_shift, s);  // k_stats does it itself
  ctx->launches += 2;
  const void *rgb_d, *depth_d;
  if ((r = copy_in(ctx, ctx->in_rgbf, rgb, sizeof(float) * 3 * (size_t)in_h * in_w, mem, s, &rgb_d))) return r;
  if ((r = copy_in(ctx, ctx->in_depthf, depth, sizeof(float) * (size_t)in_h * in_w, mem, s, &depth_d))) return r;
  const float* src_f32 = (const float*)rgb_d;
  if (in_h != H || in_w != W) {
    if ((r = ensure(ctx, ctx->frameB, sizeof(float) * 3 * (size_t)W * H))) return r;
    launch_resize_planar((const float*)rgb_d, 3, in_h, in_w, (float*)ctx->frameB.p, H, W, s);
    ctx->launches += 1;
    src_f32 = (const float*)ctx->frameB.p;
  }
  uint8_t *l_d = left_bgr, *r_d = right_bgr;
  size_t eye_bytes = (size_t)W * H * 3;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->eyeL, eye_bytes))) return r;
    if ((r = ensure(ctx, ctx->eyeR, eye_bytes))) return r;
    l_d = (uint8_t*)ctx->eyeL.p;
    r_d = (uint8_t*)ctx->eyeR.p;
  }
  CoreIn ci;
  ci.depth = (const float*)depth_d;
  ci.sh = in_h;
  ci.sw = in_w;
  ci.src_u8 = nullptr;
  ci.src_pitch = 0;
  ci.cx0 = ci.cy0 = 0;
  ci.src_f32 = src_f32;
  ci.W = W;
  ci.H = H;
  ci.p = *p;
  ci.grade = 0;
  ci.sat = 1.f;
  ci.con = 1.f;
  ci.bri = 0.f;
  ci.left = l_d;
  ci.right = r_d;
  if (fastp) {
    FastPost post;
    if ((r = run_core_fast(ctx, ci, nullptr, post))) return r;
  } else {
    if ((r = run_core(ctx, ci))) return r;
  }
  if (mem == VD3D_MEM_HOST) {
    CK(cudaMemcpyAsync(left_bgr, l_d, eye_bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(right_bgr, r_d, eye_bytes, cudaMemcpyDeviceToHost, s));
  }
  if (shift)
    CK(cudaMemcpyAsync(shift, ctx->shift.p, sizeof(float) * (size_t)W * H,
                       mem == VD3D_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
  if (info) {
    if ((r = fetch_info(ctx, info))) return r;
  } else {
    CK(cudaStreamSynchronize(s));
  }
  return VD3D_OK;
}

int vd3d_plan_sizes(int src_w, int src_h, const vd3d_render_params* rp, vd3d_size_plan* o) {
  if (!rp || !o || src_w < 2 || src_h < 2) return VD3D_ERR_ARG;
  double target_ratio = rp->aspect_ratio;
  int cw = src_w, ch = src_h, cx0 = 0, cy0 = 0;
  double cur = (double)src_w / (double)src_h;
  if (fabs(cur - target_ratio) > 0.01) {
    if (cur > target_ratio) {
      cw = (int)(src_h * target_ratio);
      cx0 = (src_w - cw) / 2;
    } else {
      ch = (int)(src_w / target_ratio);
      cy0 = (src_h - ch) / 2;
    }
  }
  int rw, rh, pw, ph, outw, outh, tew, teh;
  int fmt = rp->output_format;
  if (rp->preserve_original_aspect) {
    // original_video_* default to the FIRST frame tensor's size (pre-crop) (core/render_3d.py:1087-1092)
    int ow = src_w, oh = src_h;
    if (rp->original_video_width > 0 && rp->original_video_height > 0) {
      ow = rp->original_video_width;
      oh = rp->original_video_height;
    }
    rw = ow;
    rh = oh;
    if (fmt == VD3D_FMT_FULL_SBS) {
      pw = rw; ph = rh; outw = rw * 2; outh = rh;
    } else if (fmt == VD3D_FMT_HALF_SBS) {
      pw = rw / 2; ph = rh; outw = rw; outh = rh;
    } else if (fmt == VD3D_FMT_VR) {  // core/render_3d.py:1104-1108
      pw = 1440; ph = 1600; outw = 2880; outh = 1600;
    } else {
      pw = rw; ph = rh; outw = rw * 2; outh = rh;
    }
    tew = pw;
    teh = ph;
  } else {
    rh = rp->output_height;
    rw = (int)(rh * target_ratio);
    if (rw % 2 != 0) rw += 1;
    if (fmt == VD3D_FMT_FULL_SBS) {
      pw = 1920; ph = 1080; outw = 3840; outh = 1080;
    } else if (fmt == VD3D_FMT_HALF_SBS) {
      pw = rw / 2; ph = rh; outw = rw; outh = rh;
    } else if (fmt == VD3D_FMT_VR) {  // core/render_3d.py:1129-1133
      pw = 1440; ph = 1600; outw = 2880; outh = 1600;
    } else {
      pw = rw; ph = rh; outw = rw * 2; outh = rh;
    }
    tew = pw;
    teh = (int)(pw / target_ratio);
    if (teh % 2 != 0) teh += 1;
  }
  o->crop_x0 = cx0;
  o->crop_y0 = cy0;
  o->crop_w = cw;
  o->crop_h = ch;
  o->target_eye_w = tew;
  o->target_eye_h = teh;
  o->resized_width = rw;
  o->resized_height = rh;
  o->per_eye_w = pw;
  o->per_eye_h = ph;
  o->out_width = outw;
  o->out_height = outh;
  return VD3D_OK;
}

// the vd3d_shift_params render_sbs_3d hands to pixel_shift_cuda (core/render_3d.py:1284-1331)
static vd3d_shift_params loop_shift_params(const vd3d_render_params* rp) {
  vd3d_shift_params sp;
  memset(&sp, 0, sizeof sp);
  sp.fg_shift = sp.mg_shift = sp.bg_shift = 0.0;  // taken from FrameScalars (set by k_fin_norm)
  sp.blur_ksize = rp->blur_ksize;
  sp.feather_strength = rp->feather_strength;
  sp.max_pixel_shift_percent = rp->max_pixel_shift_percent;
  sp.parallax_balance = 0.8;  // never forwarded by render_sbs_3d (core/render_3d.py:1284-1331)
  sp.zero_parallax_strength = rp->zero_parallax_strength;
  sp.use_subject_tracking = rp->use_subject_tracking;
  sp.enable_floating_window = rp->use_floating_window;
  sp.enable_feathering = rp->enable_feathering;
  sp.enable_edge_masking = rp->enable_edge_masking;
  sp.convergence_strength = rp->convergence_strength;
  sp.enable_dynamic_convergence = rp->enable_dynamic_convergence;
  sp.depth_pop_gamma = 0.85;  // hard-coded at the call site (1299-1305)
  sp.depth_pop_mid = 0.50;
  sp.depth_stretch_lo = 0.05;
  sp.depth_stretch_hi = 0.95;
  sp.fg_pop_multiplier = 1.20;
  sp.bg_push_multiplier = 1.10;
  sp.subject_lock_strength = 1.00;
  return sp;
}

// fast-path core of one loop iteration: k_stats (ingest .. scalar trackers) -> k_shift -> k_render.  *fused tells the
// caller that bars + sharpen + eye fit + pack already happened inside k_render (out_d is complete).
static int enqueue_core_fast(vd3d_ctx* ctx, const uint8_t* frame_d, const uint8_t* depth_d, int depth_channels,
                             int src_h, int src_w, const vd3d_render_params* rp, const vd3d_size_plan& pl,
                             uint8_t* out_d, bool ident, bool* fused) {
  int r;
  const int tw = pl.target_eye_w, th = pl.target_eye_h;
  const int W = pl.resized_width, H = pl.resized_height;
  const size_t tpx = (size_t)tw * th;
  IngestArgs ia;
  memset(&ia, 0, sizeof ia);
  ia.frame = frame_d;
  ia.depth = depth_d;
  ia.depth_ch = depth_channels;
  ia.src_w = src_w;
  ia.src_h = src_h;
  ia.cx0 = pl.crop_x0;
  ia.cy0 = pl.crop_y0;
  ia.cw = pl.crop_w;
  ia.ch = pl.crop_h;
  ia.tw = tw;
  ia.th = th;
  ia.tdf = (float*)ctx->tdf.p;
  ia.alpha = 0.5f;
  ia.one_minus_alpha = (float)(1 - 0.5);
  ia.st = ctx->st;
  FastLoop lp;
  memset(&lp, 0, sizeof lp);
  lp.ia = &ia;
  FastPost post;
  if (!ident) {
    if ((r = ensure(ctx, ctx->rgb_s, sizeof(float4) * tpx))) return r;
    lp.rgbx_s = (float4*)ctx->rgb_s.p;
    post.src_rgbx = lp.rgbx_s;
    if (W != tw || H != th) {
      if ((r = ensure(ctx, ctx->frameB, sizeof(float4) * (size_t)W * H))) return r;
      lp.rgbx = (float4*)ctx->frameB.p;
      post.src_rgbx = lp.rgbx;
    }
  }
  lp.dn = (float*)(ctx->frame_parity ? ctx->dn1.p : ctx->dn0.p);
  lp.dn_prev = (const float*)(ctx->frame_parity ? ctx->dn0.p : ctx->dn1.p);
  LoopArgs la;
  la.fg = rp->fg_shift;
  la.mg = rp->mg_shift;
  la.bg = rp->bg_shift;
  la.ipd = rp->ipd_factor;
  la.resized_width = W;
  la.use_floating_window = rp->use_floating_window;
  la.use_subject_tracking = rp->use_subject_tracking;
  la.crop_count = (long long)(th * 3 / 4 - th / 4) * (long long)(tw * 3 / 4 - tw / 4);
  la.npix = (long long)tpx;
  la.dyn_min = (float)0.90;
  la.dyn_span = (float)(1.15 - 0.90);
  lp.la = &la;

  const size_t eye_bytes = (size_t)W * H * 3;
  CoreIn ci;
  memset(&ci, 0, sizeof ci);
  ci.depth = lp.dn;
  ci.sh = th;
  ci.sw = tw;
  ci.src_u8 = ident ? frame_d : nullptr;
  ci.src_pitch = src_w;
  ci.cx0 = pl.crop_x0;
  ci.cy0 = pl.crop_y0;
  ci.W = W;
  ci.H = H;
  ci.p = loop_shift_params(rp);
  const bool dof = rp->dof_strength > 0.0;
  ci.grade = dof ? 0 : 1;
  ci.sat = (float)rp->color_saturation;
  ci.con = (float)rp->color_contrast;
  ci.bri = (float)rp->color_brightness;

  // can k_render finish the frame?  SBS formats whose eye fit is the identity or cv2's integer 2:1 horizontal INTER_AREA
  *fused = false;
  if (!ctx->stats_only && !dof && (rp->output_format == VD3D_FMT_HALF_SBS || rp->output_format == VD3D_FMT_FULL_SBS)) {
    FitPlan fp;
    if ((r = plan_fit(ctx, rp->output_format, W, H, pl.per_eye_w, pl.per_eye_h, fp))) return r;
    const bool plain = !fp.xal && !fp.lin && fp.fit_x0 == 0 && fp.fit_y0 == 0 && fp.fit_w == pl.per_eye_w &&
                       fp.fit_h == pl.per_eye_h && fp.sy == 1 && pl.per_eye_h == H;
    if (plain && fp.sx == 1 && pl.per_eye_w == W) post.fuse = 1;
    if (plain && fp.sx == 2 && pl.per_eye_w * 2 == W) post.fuse = 2;
  }
  if (post.fuse) {
    post.sharpen = 1;
    sharpen_coeffs(rp->sharpness_factor, post.kc, post.ke);
    post.out = out_d;
    post.out_w = 2 * pl.per_eye_w;
    post.per_eye_w = pl.per_eye_w;
    *fused = true;
  } else if (!ctx->stats_only) {
    if ((r = ensure(ctx, ctx->eyeL, eye_bytes))) return r;
    if ((r = ensure(ctx, ctx->eyeR, eye_bytes))) return r;
    ci.left = (uint8_t*)ctx->eyeL.p;
    ci.right = (uint8_t*)ctx->eyeR.p;
  }
  return run_core_fast(ctx, ci, &lp, post);
}

// enqueue one loop iteration on ctx->stream; inputs/outputs are DEVICE pointers
static int enqueue_frame(vd3d_ctx* ctx, const uint8_t* frame_d, const uint8_t* depth_d, int depth_channels,
                         int src_h, int src_w, const vd3d_render_params* rp, const vd3d_size_plan& pl,
                         uint8_t* out_d) {
  cudaStream_t s = ctx->stream;
  int r;
  ProfScope frame_scope(ctx, 0);
  const int tw = pl.target_eye_w, th = pl.target_eye_h;
  const int W = pl.resized_width, H = pl.resized_height;
  if (tw < 8 || th < 8 || W < 8 || H < 8) return fail(ctx, VD3D_ERR_ARG, "frame too small");
  size_t tpx = (size_t)tw * th;
  if (ctx->tdf_w != tw || ctx->tdf_h != th) {
    if ((r = ensure(ctx, ctx->tdf, sizeof(float) * tpx))) return r;
    if ((r = ensure(ctx, ctx->dn0, sizeof(float) * tpx))) return r;
    if ((r = ensure(ctx, ctx->dn1, sizeof(float) * tpx))) return r;
    ctx->tdf_w = tw;
    ctx->tdf_h = th;
    if ((r = vd3d_reset_state(ctx, VD3D_STATE_CLIP))) return r;
  }
  bool ident = (tw == pl.crop_w && th == pl.crop_h && W == tw && H == th);
  bool need_rgb_s = !ident;
  if ((r = begin_frame(ctx))) return r;
  ctx->launches += 2;

  const bool fastp = !ctx->exact && render_supports(rp->enable_feathering ? 1 : 0, rp->blur_ksize);
  bool fused = false;
  if (fastp) {
    if ((r = enqueue_core_fast(ctx, frame_d, depth_d, depth_channels, src_h, src_w, rp, pl, out_d, ident, &fused)))
      return r;
    if (ctx->stats_only || fused) {
      ctx->frame_parity ^= 1;
      return VD3D_OK;
    }
  } else {
  // ---- ingest + TemporalDepthFilter
  IngestArgs ia;
  ia.frame = frame_d;
  ia.depth = depth_d;
  ia.depth_ch = depth_channels;
  ia.src_w = src_w;
  ia.src_h = src_h;
  ia.cx0 = pl.crop_x0;
  ia.cy0 = pl.crop_y0;
  ia.cw = pl.crop_w;
  ia.ch = pl.crop_h;
  ia.tw = tw;
  ia.th = th;
  ia.tdf = (float*)ctx->tdf.p;
  ia.rgb_s = nullptr;
  if (need_rgb_s) {
    if ((r = ensure(ctx, ctx->rgb_s, sizeof(float) * 3 * tpx))) return r;
    ia.rgb_s = (float*)ctx->rgb_s.p;
  }
  ia.alpha = 0.5f;
  ia.one_minus_alpha = (float)(1 - 0.5);
  ia.st = ctx->st;
  launch_ingest(ia, s);
  // ---- DepthPercentileEMA(p_lo=.02, p_hi=.98, alpha=.92)
  uint32_t l0, l1, h0, h1;
  float wlo, whi;
  quantile_rank(0.02, (long long)tpx, l0, l1, wlo);
  quantile_rank(0.98, (long long)tpx, h0, h1, whi);
  launch_set_ranks(ctx->jm[0].tg, l0, l1, h0, h1, s);
  SelJob pj = make_job(ctx->jm[0], (const float*)ctx->tdf.p, tw, 0, tw, 0, th, 0, 4, 0, false);
  launch_select(pj, nullptr, sel_blocks(th), s);
  launch_fin_pct(pj, wlo, whi, 0.92f, (float)(1 - 0.92), ctx->st, ctx->fs, s);
  float* dn = (float*)(ctx->frame_parity ? ctx->dn1.p : ctx->dn0.p);
  const float* dn_prev = (const float*)(ctx->frame_parity ? ctx->dn0.p : ctx->dn1.p);
  launch_normalize((const float*)ctx->tdf.p, dn, dn_prev, th, tw, ctx->st, ctx->fs, s);
  SelJob sn = make_job(ctx->jm[1], dn, tw, tw / 5, tw * 4 / 5, th / 5, th * 4 / 5, 1, 1, 1, true);
  launch_select(sn, nullptr, sel_blocks(th * 4 / 5 - th / 5), s);
  LoopArgs la;
  la.fg = rp->fg_shift;
  la.mg = rp->mg_shift;
  la.bg = rp->bg_shift;
  la.ipd = rp->ipd_factor;
  la.resized_width = W;
  la.use_floating_window = rp->use_floating_window;
  la.use_subject_tracking = rp->use_subject_tracking;
  la.crop_count = (long long)(th * 3 / 4 - th / 4) * (long long)(tw * 3 / 4 - tw / 4);
  la.npix = (long long)tpx;
  la.dyn_min = (float)0.90;
  la.dyn_span = (float)(1.15 - 0.90);
  launch_fin_norm(sn, la, ctx->st, ctx->fs, s);
  ctx->launches += 1 + 1 + 6 + 1 + 1 + 6 + 1;

  // ---- pixel_shift_cuda
  size_t eye_bytes = (size_t)W * H * 3;
  if ((r = ensure(ctx, ctx->eyeL, eye_bytes))) return r;
  if ((r = ensure(ctx, ctx->eyeR, eye_bytes))) return r;
  CoreIn ci;
  ci.depth = dn;
  ci.sh = th;
  ci.sw = tw;
  ci.src_u8 = nullptr;
  ci.src_pitch = src_w;
  ci.cx0 = pl.crop_x0;
  ci.cy0 = pl.crop_y0;
  ci.src_f32 = nullptr;
  if (ident) {
    ci.src_u8 = frame_d;
  } else {
    const float* base = (const float*)ctx->rgb_s.p;
    if (W == tw && H == th) {
      ci.src_f32 = base;
    } else {
      if ((r = ensure(ctx, ctx->frameB, sizeof(float) * 3 * (size_t)W * H))) return r;
      launch_resize_planar(base, 3, th, tw, (float*)ctx->frameB.p, H, W, s);
      ctx->launches += 1;
      ci.src_f32 = (const float*)ctx->frameB.p;
    }
  }
  ci.W = W;
  ci.H = H;
  ci.p = loop_shift_params(rp);
  bool dof = rp->dof_strength > 0.0;
  ci.grade = dof ? 0 : 1;
  ci.sat = (float)rp->color_saturation;
  ci.con = (float)rp->color_contrast;
  ci.bri = (float)rp->color_brightness;
  ci.left = (uint8_t*)ctx->eyeL.p;
  ci.right = (uint8_t*)ctx->eyeR.p;
  if ((r = run_core(ctx, ci))) return r;

  if (ctx->stats_only) {
    ctx->frame_parity ^= 1;
    return VD3D_OK;
  }
  }  // exact path
  const size_t eye_bytes = (size_t)W * H * 3;
  const bool dof = rp->dof_strength > 0.0;
  float* dn = (float*)(ctx->frame_parity ? ctx->dn1.p : ctx->dn0.p);
  const uint8_t* eye_l = (const uint8_t*)ctx->eyeL.p;
  const uint8_t* eye_r = (const uint8_t*)ctx->eyeR.p;
  if (dof) {
    if ((r = ensure_dof_kernels(ctx, rp->dof_strength, 5))) return r;
    if ((r = ensure(ctx, ctx->eyeL2, eye_bytes))) return r;
    if ((r = ensure(ctx, ctx->eyeR2, eye_bytes))) return r;
    DofArgs da;
    da.src_l = eye_l;
    da.src_r = eye_r;
    da.dst_l = (uint8_t*)ctx->eyeL2.p;
    da.dst_r = (uint8_t*)ctx->eyeR2.p;
    da.H = H;
    da.W = W;
    da.depth = dn;
    da.dh = th;
    da.dw = tw;
    da.focal = 0.f;
    da.fs = ctx->fs;  // focal comes from the FocalDepthTracker state on the device
    da.focus_w = (float)(0.35 + 1e-6);
    da.idx_max = (float)(5 - 1 - 1e-6);
    da.nlevels = 5;
    for (int i = 0; i < 8; ++i) {
      da.ksize[i] = ctx->dof_ksize[i];
      da.koff[i] = ctx->dof_koff[i];
    }
    da.kern = (const float*)ctx->dof_kern.p;
    da.halo = ctx->dof_halo;
    da.sat = (float)rp->color_saturation;
    da.con = (float)rp->color_contrast;
    da.bri = (float)rp->color_brightness;
    launch_dof(da, 2, s);
    ctx->launches += 1;
    eye_l = da.dst_l;
    eye_r = da.dst_r;
  }
  // ---- bars + sharpen + fit + pack
  FitPlan fp;
  if ((r = plan_fit(ctx, rp->output_format, W, H, pl.per_eye_w, pl.per_eye_h, fp))) return r;
  PostArgs pa;
  pa.left = eye_l;
  pa.right = eye_r;
  pa.H = H;
  pa.W = W;
  pa.fs = ctx->fs;
  pa.sharpen = 1;
  sharpen_coeffs(rp->sharpness_factor, pa.kc, pa.ke);
  pa.fmt = rp->output_format;
  pa.per_eye_w = pl.per_eye_w;
  pa.per_eye_h = pl.per_eye_h;
  set_fit(pa, fp);
  pa.out = out_d;
  if (rp->output_format == VD3D_FMT_ANAGLYPH || rp->output_format == VD3D_FMT_INTERLACED) {
    pa.out_w = pl.per_eye_w;
    pa.out_h = pl.per_eye_h;
  } else {
    pa.out_w = 2 * pl.per_eye_w;  // hstack of the fitted eyes
    pa.out_h = pl.per_eye_h;
  }
  launch_post(pa, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  ctx->frame_parity ^= 1;
  return VD3D_OK;
}

extern "C" uint64_t vd3d_depth_launch_count(vd3d_depth* e);
extern "C" void vd3d_depth_add_launches(vd3d_depth* e, uint64_t n);

extern "C" int vd3d_depth_infer_batch_device(vd3d_depth* e, int B, const uint8_t* const* frames_bgr_dev, int h, int w,
                                             uint8_t* const* depth_u8_dev, float* const* depth_f32_dev, int invert);

static void drop_depth_graphs(vd3d_ctx* ctx) {
  for (int i = 0; i < kClones; ++i)
    for (int j = 0; j <= kMaxDepthBatch; ++j)
      if (ctx->dg[i][j].exec) {
        cudaGraphExecDestroy(ctx->dg[i][j].exec);
        ctx->dg[i][j].exec = nullptr;
      }
  ctx->dg_warm = 0;
}

static int ensure_depth_clones(vd3d_ctx* ctx, vd3d_depth* parent) {
  if (ctx->dclone_parent == parent && ctx->dclone[0] && ctx->dclone_wver == vd3d_depth_weights_version(parent))
    return VD3D_OK;
  CK(cudaDeviceSynchronize());
  drop_depth_graphs(ctx);
  drop_graphs(ctx);
  ctx->dclone_wver = vd3d_depth_weights_version(parent);
  for (int i = 0; i < kClones; ++i) {
    if (ctx->dclone[i]) vd3d_depth_destroy(ctx->dclone[i]);
    ctx->dclone[i] = nullptr;
    if (!ctx->s_depth[i]) CK(cudaStreamCreateWithFlags(&ctx->s_depth[i], cudaStreamNonBlocking));
    if (!ctx->ev_depth[i]) CK(cudaEventCreateWithFlags(&ctx->ev_depth[i], cudaEventDisableTiming));
    if (vd3d_depth_clone(parent, ctx->s_depth[i], &ctx->dclone[i]) != VD3D_OK)
      return fail(ctx, VD3D_ERR_ARG, "vd3d_depth_clone failed");
  }
  ctx->dclone_parent = parent;
  return VD3D_OK;
}

// depth inference of one batch (staging slots slot0 .. slot0+nb-1) on engine instance c and its stream; one batched
// forward for the nb frames, graph-replayed once the configuration has been seen a few times
static int run_depth_group(vd3d_ctx* ctx, vd3d_depth* parent, int c, int slot0, int nb, int src_h, int src_w) {
  vd3d_depth* e = ctx->dclone[c];
  const uint8_t* f_d[kMaxDepthBatch];
  uint8_t* d_d[kMaxDepthBatch];
  for (int j = 0; j < nb; ++j) {
    f_d[j] = (const uint8_t*)ctx->in_frame[slot0 + j].p;
    d_d[j] = (uint8_t*)ctx->in_depth[slot0 + j].p;
  }
  if (ctx->dg_h != src_h || ctx->dg_w != src_w) {
    drop_depth_graphs(ctx);
    ctx->dg_h = src_h;
    ctx->dg_w = src_w;
  }
  auto eager = [&]() -> int {
    uint64_t l0 = vd3d_depth_launch_count(e);
    int r = vd3d_depth_infer_batch_device(e, nb, f_d, src_h, src_w, d_d, nullptr, 0);
    if (r) ctx->err = std::string("depth engine: ") + vd3d_depth_last_error(e);
    vd3d_depth_add_launches(parent, vd3d_depth_launch_count(e) - l0);
    return r;
  };
  if (!ctx->use_graphs || ctx->dg_warm < 2 * kClones) {  // workspaces of both instances are allocated eagerly first
    ctx->dg_warm++;
    return eager();
  }
  vd3d_ctx::DepthGraph& g = ctx->dg[c][nb];
  if (!g.exec) {
    uint64_t l0 = vd3d_depth_launch_count(e);
    if (cudaStreamBeginCapture(ctx->s_depth[c], cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      fprintf(stderr, "vd3d: CUDA graph capture unavailable (%s); continuing with eager launches\n",
              cudaGetErrorString(cudaGetLastError()));
      ctx->use_graphs = 0;
      return eager();
    }
    int r = vd3d_depth_infer_batch_device(e, nb, f_d, src_h, src_w, d_d, nullptr, 0);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(ctx->s_depth[c], &graph);
    uint64_t n = vd3d_depth_launch_count(e) - l0;
    if (r != VD3D_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
      fprintf(stderr, "vd3d: CUDA graph capture failed (r=%d, %s); continuing with eager launches\n", r,
              cudaGetErrorString(cudaGetLastError()));
      if (graph) cudaGraphDestroy(graph);
      g.exec = nullptr;
      ctx->use_graphs = 0;
      return eager();
    }
    cudaGraphDestroy(graph);
    g.n = n;
  }
  CK(cudaGraphLaunch(g.exec, ctx->s_depth[c]));
  vd3d_depth_add_launches(parent, g.n);
  return VD3D_OK;
}

static void drop_graphs(vd3d_ctx* ctx) {
  for (int i = 0; i < kSlots; ++i)
    for (int j = 0; j < 2; ++j)
      if (ctx->fg[i][j].exec) {
        cudaGraphExecDestroy(ctx->fg[i][j].exec);
        ctx->fg[i][j].exec = nullptr;
      }
  ctx->fg_warm = 0;
}

// One frame on the staging buffers of slot b: [depth inference ->] render_sbs_3d loop body.
// After two eager frames of an unchanged configuration the launch sequence (~200 kernels) is captured
// once per (slot, parity) into a CUDA graph and replayed; all scalars it depends on live on the device.
static int run_frame_slot(vd3d_ctx* ctx, vd3d_depth* depth, int b, int depth_channels, int src_h, int src_w,
                          const vd3d_render_params* rp, const vd3d_size_plan& pl) {
  const uint8_t* f_d = (const uint8_t*)ctx->in_frame[b].p;
  uint8_t* d_d = (uint8_t*)ctx->in_depth[b].p;
  uint8_t* o_d = (uint8_t*)ctx->out_dev[b].p;
  auto eager = [&]() -> int {
    int r;
    if (depth) {
      ProfScope ps(ctx, 2);
      if ((r = vd3d_depth_infer_device(depth, f_d, src_h, src_w, d_d, nullptr, 0))) {
        ctx->err = std::string("depth engine: ") + vd3d_depth_last_error(depth);
        return r;
      }
    }
    return enqueue_frame(ctx, f_d, d_d, depth ? 1 : depth_channels, src_h, src_w, rp, pl, o_d);
  };
  bool same = ctx->fg_h == src_h && ctx->fg_w == src_w && ctx->fg_dch == depth_channels && ctx->fg_depth == depth &&
              memcmp(&ctx->fg_rp, rp, sizeof *rp) == 0;
  if (ctx->fg_epoch != ctx->res_epoch) {  // another entry point moved / rewrote something the graphs bake in
    drop_graphs(ctx);
    ctx->fg_epoch = ctx->res_epoch;
  }
  if (!same) {
    drop_graphs(ctx);
    ctx->fg_h = src_h;
    ctx->fg_w = src_w;
    ctx->fg_dch = depth_channels;
    ctx->fg_depth = depth;
    ctx->fg_rp = *rp;
  }
  if (!ctx->use_graphs || ctx->prof || ctx->fg_warm < 3) {
    ctx->fg_warm++;
    return eager();
  }
  const int par = ctx->frame_parity;
  vd3d_ctx::FrameGraph& g = ctx->fg[b][par];
  if (!g.exec) {
    uint64_t l0 = ctx->launches, d0 = depth ? vd3d_depth_launch_count(depth) : 0;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      fprintf(stderr, "vd3d: CUDA graph capture unavailable (%s); continuing with eager launches\n",
              cudaGetErrorString(cudaGetLastError()));
      ctx->use_graphs = 0;
      return eager();
    }
    int r = eager();
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->frame_parity = par;  // capture does not execute
    uint64_t nl = ctx->launches - l0, nd = depth ? vd3d_depth_launch_count(depth) - d0 : 0;
    ctx->launches = l0;
    if (depth) vd3d_depth_add_launches(depth, (uint64_t)0 - nd);
    if (r != VD3D_OK || ce != cudaSuccess || !graph ||
        cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
      fprintf(stderr, "vd3d: CUDA graph capture failed (r=%d, %s); continuing with eager launches\n", r,
              cudaGetErrorString(cudaGetLastError()));
      if (graph) cudaGraphDestroy(graph);
      g.exec = nullptr;
      ctx->use_graphs = 0;  // fall back to eager launches for the rest of this ctx
      return eager();
    }
    cudaGraphDestroy(graph);
    g.n_ctx = nl;
    g.n_depth = nd;
  }
  CK(cudaGraphLaunch(g.exec, ctx->stream));
  ctx->launches += g.n_ctx;
  if (depth) vd3d_depth_add_launches(depth, g.n_depth);
  ctx->frame_parity ^= 1;
  return VD3D_OK;
}

// the packed frame format_3d_output produces (core/render_3d.py:837-860): SBS / VR = hstack of the two fitted eyes,
// i.e. 2 * per_eye_w wide -- for an odd preserve-aspect Half-SBS width that is one less than `out_width`, the size the
// reference opens its writer with (1099-1103)
static int packed_w(const vd3d_render_params* rp, const vd3d_size_plan& pl) {
  if (rp->output_format == VD3D_FMT_ANAGLYPH || rp->output_format == VD3D_FMT_INTERLACED) return pl.per_eye_w;
  return 2 * pl.per_eye_w;
}
static size_t out_bytes(const vd3d_render_params* rp, const vd3d_size_plan& pl) {
  return (size_t)packed_w(rp, pl) * pl.per_eye_h * 3;
}

int vd3d_render_frame(vd3d_ctx* ctx, const uint8_t* frame_bgr, const uint8_t* depth, int depth_channels,
                      int src_h, int src_w, const vd3d_render_params* rp, uint8_t* out_bgr, int mem,
                      vd3d_frame_info* info) {
  if (!ctx || !frame_bgr || !depth || !rp || !out_bgr) return fail(ctx, VD3D_ERR_ARG, "null argument");
  if (depth_channels != 1 && depth_channels != 3) return fail(ctx, VD3D_ERR_ARG, "depth_channels must be 1 or 3");
  CK(cudaSetDevice(ctx->device));
  vd3d_size_plan pl;
  int r = vd3d_plan_sizes(src_w, src_h, rp, &pl);
  if (r) return fail(ctx, r, "unsupported output format / sizes");
  cudaStream_t s = ctx->stream;
  const void *f_d, *d_d;
  size_t fb = (size_t)src_w * src_h * 3, db = (size_t)src_w * src_h * depth_channels;
  if ((r = copy_in(ctx, ctx->in_frame[0], frame_bgr, fb, mem, s, &f_d))) return r;
  if ((r = copy_in(ctx, ctx->in_depth[0], depth, db, mem, s, &d_d))) return r;
  size_t ob = out_bytes(rp, pl);
  uint8_t* o_d = out_bgr;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->out_dev[0], ob))) return r;
    o_d = (uint8_t*)ctx->out_dev[0].p;
  }
  if ((r = enqueue_frame(ctx, (const uint8_t*)f_d, (const uint8_t*)d_d, depth_channels, src_h, src_w, rp, pl, o_d)))
    return r;
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(out_bgr, o_d, ob, cudaMemcpyDeviceToHost, s));
  if (info) return fetch_info(ctx, info);
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

int vd3d_render_clip(vd3d_ctx* ctx, int n, const uint8_t* const* frames, const uint8_t* const* depths,
                     int depth_channels, int src_h, int src_w, const vd3d_render_params* rp,
                     uint8_t* const* outs, int mem, vd3d_frame_info* infos) {
  if (!ctx || !frames || !depths || !rp || !outs || n < 0) return fail(ctx, VD3D_ERR_ARG, "null argument");
  if (depth_channels != 1 && depth_channels != 3) return fail(ctx, VD3D_ERR_ARG, "depth_channels must be 1 or 3");
  CK(cudaSetDevice(ctx->device));
  vd3d_size_plan pl;
  int r = vd3d_plan_sizes(src_w, src_h, rp, &pl);
  if (r) return fail(ctx, r, "unsupported output format / sizes");
  size_t fb = (size_t)src_w * src_h * 3, db = (size_t)src_w * src_h * depth_channels;
  size_t ob = out_bytes(rp, pl);
  for (int b = 0; b < kSlots; ++b) {
    if ((r = ensure(ctx, ctx->in_frame[b], fb))) return r;
    if ((r = ensure(ctx, ctx->in_depth[b], db))) return r;
    if ((r = ensure(ctx, ctx->out_dev[b], ob))) return r;
  }
  for (int i = 0; i < n; ++i) {
    int b = i % kSlots;
    if (mem == VD3D_MEM_DEVICE) {
      // fixed staging addresses keep the frame graph replayable; D2D copies are ~1 % of a frame
      if (i >= kSlots) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0));
      CK(cudaMemcpyAsync(ctx->in_frame[b].p, frames[i], fb, cudaMemcpyDeviceToDevice, ctx->stream));
      CK(cudaMemcpyAsync(ctx->in_depth[b].p, depths[i], db, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      // software pipeline over three streams: H2D(i+1) | kernels(i) | D2H(i-1)
      if (i >= kSlots) CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_done[b], 0));
      CK(cudaMemcpyAsync(ctx->in_frame[b].p, frames[i], fb, cudaMemcpyHostToDevice, ctx->s_h2d));
      CK(cudaMemcpyAsync(ctx->in_depth[b].p, depths[i], db, cudaMemcpyHostToDevice, ctx->s_h2d));
      CK(cudaEventRecord(ctx->ev_h2d[b], ctx->s_h2d));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
      if (i >= kSlots) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0));
    }
    if ((r = run_frame_slot(ctx, nullptr, b, depth_channels, src_h, src_w, rp, pl))) return r;
    if (infos && (r = fetch_info(ctx, &infos[i]))) return r;
    CK(cudaEventRecord(ctx->ev_done[b], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[b], 0));
    CK(cudaMemcpyAsync(outs[i], ctx->out_dev[b].p, ob,
                       mem == VD3D_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->s_d2h));
    CK(cudaEventRecord(ctx->ev_d2h[b], ctx->s_d2h));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(ctx->s_d2h));
  return VD3D_OK;
}

int vd3d_render_clip_depth(vd3d_ctx* ctx, vd3d_depth* depth, int n, const uint8_t* const* frames, int src_h,
                           int src_w, const vd3d_render_params* rp, uint8_t* const* outs, int mem) {
  if (!ctx || !depth || !frames || !rp || !outs || n < 0) return fail(ctx, VD3D_ERR_ARG, "null argument");
  CK(cudaSetDevice(ctx->device));
  vd3d_size_plan pl;
  int r = vd3d_plan_sizes(src_w, src_h, rp, &pl);
  if (r) return fail(ctx, r, "unsupported output format / sizes");
  size_t fb = (size_t)src_w * src_h * 3, db = (size_t)src_w * src_h;
  size_t ob = out_bytes(rp, pl);
  for (int b = 0; b < kSlots; ++b) {
    if ((r = ensure(ctx, ctx->in_frame[b], fb))) return r;
    if ((r = ensure(ctx, ctx->in_depth[b], db))) return r;
    if ((r = ensure(ctx, ctx->out_dev[b], ob))) return r;
  }
  // Frames travel in groups of `depth_batch`: one batched depth forward per group on one of two engine instances
  // (group g+1's forward overlaps the neck / head tail of group g and the DIBR kernels of group g's frames), then the
  // DIBR loop body frame by frame on the main stream (sequential temporal state), D2H behind it.
  const bool serial = ctx->prof != 0;  // stage timing wants everything on one stream
  const int B = ctx->depth_batch;
  if (!serial && (r = ensure_depth_clones(ctx, depth))) return r;
  const cudaMemcpyKind kin = mem == VD3D_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind kout = mem == VD3D_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  int group = 0;
  for (int i0 = 0; i0 < n; i0 += B, ++group) {
    const int nb = (n - i0) < B ? (n - i0) : B;
    const int c = group % kClones;
    const int slot0 = c * kMaxDepthBatch;
    const bool reuse = group >= kClones;  // these slots have been used before in this call
    // ---- stage the frames of the group (a slot is free once the DIBR pass that read it has finished) ----
    for (int j = 0; j < nb; ++j) {
      const int sl = slot0 + j;
      if (reuse) CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_done[sl], 0));
      CK(cudaMemcpyAsync(ctx->in_frame[sl].p, frames[i0 + j], fb, kin, ctx->s_h2d));
      CK(cudaEventRecord(ctx->ev_h2d[sl], ctx->s_h2d));
    }
    if (serial) {
      for (int j = 0; j < nb; ++j) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[slot0 + j], 0));
      {
        ProfScope ps(ctx, 2);  // one sample per batch: vd3d_profile_collect divides by the frames it covered
        const uint8_t* f_d[kMaxDepthBatch];
        uint8_t* d_d[kMaxDepthBatch];
        for (int j = 0; j < nb; ++j) {
          f_d[j] = (const uint8_t*)ctx->in_frame[slot0 + j].p;
          d_d[j] = (uint8_t*)ctx->in_depth[slot0 + j].p;
        }
        if ((r = vd3d_depth_infer_batch_device(depth, nb, f_d, src_h, src_w, d_d, nullptr, 0))) {
          ctx->err = std::string("depth engine: ") + vd3d_depth_last_error(depth);
          return r;
        }
        ctx->prof_depth_frames += nb;
      }
    } else {
      cudaStream_t sd = ctx->s_depth[c];
      for (int j = 0; j < nb; ++j) {
        CK(cudaStreamWaitEvent(sd, ctx->ev_h2d[slot0 + j], 0));
        if (reuse) CK(cudaStreamWaitEvent(sd, ctx->ev_done[slot0 + j], 0));  // in_depth[slot] consumed
      }
      if ((r = run_depth_group(ctx, depth, c, slot0, nb, src_h, src_w))) return r;
      CK(cudaEventRecord(ctx->ev_depth[c], sd));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_depth[c], 0));
    }
    // ---- DIBR loop body of the group's frames on the main stream ----
    for (int j = 0; j < nb; ++j) {
      const int sl = slot0 + j;
      if (reuse) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[sl], 0));  // out_dev[slot] drained
      if ((r = run_frame_slot(ctx, nullptr, sl, 1, src_h, src_w, rp, pl))) return r;
      CK(cudaEventRecord(ctx->ev_done[sl], ctx->stream));
      CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[sl], 0));
      CK(cudaMemcpyAsync(outs[i0 + j], ctx->out_dev[sl].p, ob, kout, ctx->s_d2h));
      CK(cudaEventRecord(ctx->ev_d2h[sl], ctx->s_d2h));
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaStreamSynchronize(ctx->s_d2h));
  if (!serial)
    for (int b = 0; b < kClones; ++b) CK(cudaStreamSynchronize(ctx->s_depth[b]));
  return VD3D_OK;
}

int vd3d_set_depth_batch(vd3d_ctx* ctx, int frames) {
  if (!ctx || frames < 1 || frames > kMaxDepthBatch) return fail(ctx, VD3D_ERR_ARG, "depth batch must be in [1, 8]");
  if (frames != ctx->depth_batch) {
    CK(cudaDeviceSynchronize());
    ctx->depth_batch = frames;
  }
  return VD3D_OK;
}
int vd3d_get_depth_batch(vd3d_ctx* ctx) { return ctx ? ctx->depth_batch : -1; }

// ---- exact frame sharding: advance / export / import the temporal state ----------
// One loop iteration without rendering: updates TemporalDepthFilter, DepthPercentileEMA, ShiftSmoother,
// FocalDepthTracker, ConvergenceEMA, FloatingBarEaser and FloatingWindowTracker exactly as
// vd3d_render_frame would (same kernels up to the shift map), at ~1/3 of a frame's DIBR cost.
int vd3d_advance_state(vd3d_ctx* ctx, const uint8_t* frame_bgr, const uint8_t* depth, int depth_channels, int src_h,
                       int src_w, const vd3d_render_params* rp, int mem) {
  if (!ctx || !frame_bgr || !depth || !rp) return fail(ctx, VD3D_ERR_ARG, "null argument");
  if (depth_channels != 1 && depth_channels != 3) return fail(ctx, VD3D_ERR_ARG, "depth_channels must be 1 or 3");
  CK(cudaSetDevice(ctx->device));
  vd3d_size_plan pl;
  int r = vd3d_plan_sizes(src_w, src_h, rp, &pl);
  if (r) return fail(ctx, r, "unsupported output format / sizes");
  cudaStream_t s = ctx->stream;
  const void *f_d, *d_d;
  size_t fb = (size_t)src_w * src_h * 3, db = (size_t)src_w * src_h * depth_channels;
  if ((r = copy_in(ctx, ctx->in_frame[0], frame_bgr, fb, mem, s, &f_d))) return r;
  if ((r = copy_in(ctx, ctx->in_depth[0], depth, db, mem, s, &d_d))) return r;
  ctx->stats_only = 1;
  r = enqueue_frame(ctx, (const uint8_t*)f_d, (const uint8_t*)d_d, depth_channels, src_h, src_w, rp, pl, nullptr);
  ctx->stats_only = 0;
  if (r) return r;
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

struct StateHeader {
  uint32_t magic, tw, th, reserved;
};
size_t vd3d_state_bytes(vd3d_ctx* ctx) {
  if (!ctx) return 0;
  return sizeof(StateHeader) + sizeof(DevState) + 2 * sizeof(float) * (size_t)ctx->tdf_w * ctx->tdf_h;
}
// blob = header | DevState | tdf plane | previous normalised-depth plane   (host or device memory)
int vd3d_export_state(vd3d_ctx* ctx, void* dst, size_t cap, int mem) {
  if (!ctx || !dst) return VD3D_ERR_ARG;
  size_t need = vd3d_state_bytes(ctx);
  if (cap < need) return fail(ctx, VD3D_ERR_ARG, "state buffer too small");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  cudaMemcpyKind kd = mem == VD3D_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  cudaMemcpyKind kh = mem == VD3D_MEM_HOST ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice;
  StateHeader h = {0x56443344u, (uint32_t)ctx->tdf_w, (uint32_t)ctx->tdf_h, 0};
  uint8_t* p = (uint8_t*)dst;
  CK(cudaMemcpyAsync(p, &h, sizeof h, kh, s));
  CK(cudaStreamSynchronize(s));  // h is a stack object
  p += sizeof h;
  CK(cudaMemcpyAsync(p, ctx->st, sizeof(DevState), kd, s));
  p += sizeof(DevState);
  size_t plane = sizeof(float) * (size_t)ctx->tdf_w * ctx->tdf_h;
  if (plane) {
    CK(cudaMemcpyAsync(p, ctx->tdf.p, plane, kd, s));
    p += plane;
    const void* prev = ctx->frame_parity ? ctx->dn0.p : ctx->dn1.p;  // what the next frame reads as prev_depth
    CK(cudaMemcpyAsync(p, prev, plane, kd, s));
  }
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}
int vd3d_import_state(vd3d_ctx* ctx, const void* src, size_t bytes, int mem) {
  if (!ctx || !src || bytes < sizeof(StateHeader) + sizeof(DevState)) return fail(ctx, VD3D_ERR_ARG, "bad state blob");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  cudaMemcpyKind kd = mem == VD3D_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  cudaMemcpyKind kh = mem == VD3D_MEM_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost;
  StateHeader h;
  CK(cudaMemcpyAsync(&h, src, sizeof h, kh, s));
  CK(cudaStreamSynchronize(s));
  if (h.magic != 0x56443344u) return fail(ctx, VD3D_ERR_ARG, "bad state blob magic");
  size_t plane = sizeof(float) * (size_t)h.tw * h.th;
  if (bytes < sizeof h + sizeof(DevState) + 2 * plane) return fail(ctx, VD3D_ERR_ARG, "state blob truncated");
  int r;
  if (plane) {
    if ((r = ensure(ctx, ctx->tdf, plane)) || (r = ensure(ctx, ctx->dn0, plane)) || (r = ensure(ctx, ctx->dn1, plane)))
      return r;
  }
  ctx->tdf_w = (int)h.tw;
  ctx->tdf_h = (int)h.th;
  const uint8_t* p = (const uint8_t*)src + sizeof h;
  CK(cudaMemcpyAsync(ctx->st, p, sizeof(DevState), kd, s));
  p += sizeof(DevState);
  if (plane) {
    CK(cudaMemcpyAsync(ctx->tdf.p, p, plane, kd, s));
    p += plane;
    ctx->frame_parity = 0;  // next frame writes dn0 and reads dn1 as prev_depth
    CK(cudaMemcpyAsync(ctx->dn1.p, p, plane, kd, s));
  }
  CK(cudaStreamSynchronize(s));
  drop_graphs(ctx);
  return VD3D_OK;
}

// forget the engine clones made for `depth` (call before destroying that engine)
int vd3d_release_depth(vd3d_ctx* ctx, vd3d_depth* depth) {
  if (!ctx) return VD3D_ERR_ARG;
  if (ctx->dclone_parent == depth || !depth) {
    cudaStreamSynchronize(ctx->stream);
    drop_depth_graphs(ctx);
    drop_graphs(ctx);
    for (int i = 0; i < kClones; ++i) {
      if (ctx->s_depth[i]) cudaStreamSynchronize(ctx->s_depth[i]);
      if (ctx->dclone[i]) vd3d_depth_destroy(ctx->dclone[i]);
      ctx->dclone[i] = nullptr;
    }
    ctx->dclone_parent = nullptr;
    ctx->fg_depth = nullptr;
  }
  return VD3D_OK;
}

// Validate a render configuration without launching anything: sizes, eye-fit mode (builds the INTER_AREA tables it
// will need), DOF kernel bank.  render_sbs_3d calls it before it creates the output file, so that an unsupported
// configuration fails with a message instead of leaving an empty video behind.
int vd3d_check_config(vd3d_ctx* ctx, int src_h, int src_w, const vd3d_render_params* rp) {
  if (!ctx || !rp) return fail(ctx, VD3D_ERR_ARG, "null argument");
  CK(cudaSetDevice(ctx->device));
  vd3d_size_plan pl;
  int r = vd3d_plan_sizes(src_w, src_h, rp, &pl);
  if (r) return fail(ctx, r, "unsupported output format / sizes");
  if (pl.target_eye_w < 8 || pl.target_eye_h < 8 || pl.resized_width < 8 || pl.resized_height < 8)
    return fail(ctx, VD3D_ERR_ARG, "frame too small");
  if (rp->enable_feathering && (rp->blur_ksize < 1 || rp->blur_ksize > 63))
    return fail(ctx, VD3D_ERR_UNSUPPORTED, "blur_ksize must be in [1,63]");
  FitPlan fp;
  if ((r = plan_fit(ctx, rp->output_format, pl.resized_width, pl.resized_height, pl.per_eye_w, pl.per_eye_h, fp)))
    return r;
  if (rp->dof_strength > 0.0 && (r = ensure_dof_kernels(ctx, rp->dof_strength, 5))) return r;
  return VD3D_OK;
}

// cv2.resize(u8 [h,w,ch], (ow,oh), INTER_CUBIC): the resize of the depth writer (core/render_depth.py:1917, 193; ch = 1)
// and of run_esrgan's resize chain (core/merged_pipeline.py:262-266; ch = 3)
int vd3d_resize_cubic(vd3d_ctx* ctx, const uint8_t* src, int h, int w, int ch, uint8_t* dst, int oh, int ow, int mem) {
  if (!ctx || !src || !dst || h < 1 || w < 1 || oh < 1 || ow < 1 || ch < 1 || ch > 4) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const void* s_d;
  int r;
  if ((r = copy_in(ctx, ctx->eyeL, src, (size_t)h * w * ch, mem, s, &s_d))) return r;
  uint8_t* o_d = dst;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->out_dev[0], (size_t)oh * ow * ch))) return r;
    o_d = (uint8_t*)ctx->out_dev[0].p;
  }
  if (h == oh && w == ow)
    CK(cudaMemcpyAsync(o_d, s_d, (size_t)h * w * ch, cudaMemcpyDeviceToDevice, s));  // cv2.resize copies on equal sizes
  else
    launch_resize_cubic_u8((const uint8_t*)s_d, h, w, ch, o_d, oh, ow, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(dst, o_d, (size_t)oh * ow * ch, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}
int vd3d_resize_cubic_u8(vd3d_ctx* ctx, const uint8_t* src, int h, int w, uint8_t* dst, int oh, int ow, int mem) {
  return vd3d_resize_cubic(ctx, src, h, w, 1, dst, oh, ow, mem);
}

// cv2.addWeighted(a, alpha, b, 1 - alpha... any beta, 0) on n bytes (blend_images, core/merged_pipeline.py:233-238)
int vd3d_add_weighted(vd3d_ctx* ctx, const uint8_t* a, double alpha, const uint8_t* b, double beta, size_t n, uint8_t* dst,
                      int mem) {
  if (!ctx || !a || !b || !dst || !n) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const void *a_d, *b_d;
  int r;
  if ((r = copy_in(ctx, ctx->eyeL, a, n, mem, s, &a_d)) || (r = copy_in(ctx, ctx->eyeR, b, n, mem, s, &b_d))) return r;
  uint8_t* o_d = dst;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->out_dev[0], n))) return r;
    o_d = (uint8_t*)ctx->out_dev[0].p;
  }
  launch_add_weighted((const uint8_t*)a_d, (float)alpha, (const uint8_t*)b_d, (float)beta, o_d, n, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(dst, o_d, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

// apply_color_grade (core/render_3d.py:734-767) on f32 RGB planes [3,h,w] in 0..1
int vd3d_color_grade(vd3d_ctx* ctx, const float* rgb, int h, int w, double sat, double con, double bri, float* out,
                     int mem) {
  if (!ctx || !rgb || !out || h < 1 || w < 1) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t bytes = sizeof(float) * 3 * (size_t)h * w;
  const void* r_d;
  int r;
  if ((r = copy_in(ctx, ctx->in_rgbf, rgb, bytes, mem, s, &r_d))) return r;
  float* o_d = out;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->frameB, bytes))) return r;
    o_d = (float*)ctx->frameB.p;
  }
  launch_grade_f32((const float*)r_d, o_d, h * w, (float)sat, (float)con, (float)bri, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(out, o_d, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

int vd3d_sharpen(vd3d_ctx* ctx, const uint8_t* src, int h, int w, double factor, uint8_t* dst, int mem) {
  if (!ctx || !src || !dst || h < 2 || w < 2) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t bytes = (size_t)h * w * 3;
  const void* s_d;
  int r;
  if ((r = copy_in(ctx, ctx->eyeL, src, bytes, mem, s, &s_d))) return r;
  uint8_t* o_d = dst;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->out_dev[0], bytes))) return r;
    o_d = (uint8_t*)ctx->out_dev[0].p;
  }
  PostArgs pa;
  memset(&pa, 0, sizeof pa);
  pa.left = (const uint8_t*)s_d;
  pa.right = (const uint8_t*)s_d;
  pa.H = h;
  pa.W = w;
  pa.fs = nullptr;
  pa.sharpen = 1;
  sharpen_coeffs(factor, pa.kc, pa.ke);
  pa.fmt = VD3D_FMT_INTERLACED;  // single-eye pass-through layout
  pa.per_eye_w = w;
  pa.per_eye_h = h;
  pa.fit_w = w;
  pa.fit_h = h;
  pa.sx = pa.sy = 1;
  pa.inv_area = 1.f;
  pa.out = o_d;
  pa.out_w = w;
  pa.out_h = h;
  launch_post(pa, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(dst, o_d, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

// heal_missing_pixels (core/render_3d.py:431-459) on f32 RGB planes [3,h,w]; edge_mask [h,w] or null
int vd3d_heal(vd3d_ctx* ctx, const float* warped, const float* original, const float* edge_mask, int h, int w,
              double heal_strength, float* out, int mem) {
  if (!ctx || !warped || !original || !out || h < 1 || w < 1) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t bytes = sizeof(float) * 3 * (size_t)h * w;
  const void *w_d, *o_d, *e_d = nullptr;
  int r;
  if ((r = copy_in(ctx, ctx->in_rgbf, warped, bytes, mem, s, &w_d))) return r;
  if ((r = copy_in(ctx, ctx->frameB, original, bytes, mem, s, &o_d))) return r;
  if (edge_mask && (r = copy_in(ctx, ctx->in_depthf, edge_mask, bytes / 3, mem, s, &e_d))) return r;
  float* out_d = out;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->rgb_s, bytes))) return r;
    out_d = (float*)ctx->rgb_s.p;
  }
  launch_heal((const float*)w_d, (const float*)o_d, (const float*)e_d, out_d, h, w, (float)heal_strength, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(out, out_d, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

// format_3d_output / generate_anaglyph_3d (core/render_3d.py:837-883) on two same-size u8 BGR eyes
int vd3d_pack(vd3d_ctx* ctx, const uint8_t* left, const uint8_t* right, int h, int w, int fmt, uint8_t* dst, int mem) {
  if (!ctx || !left || !right || !dst || h < 1 || w < 1) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  if (fmt < VD3D_FMT_HALF_SBS || fmt > VD3D_FMT_VR) return fail(ctx, VD3D_ERR_UNSUPPORTED, "format");
  // VR: format_3d_output resizes each eye to 1440x1600 (846-849); the render loop hands it eyes that already have that
  // size (pad_to_aspect_ratio, 1415-1417), where cv2.resize is the identity.  Other sizes (INTER_LINEAR) are off-path.
  if (fmt == VD3D_FMT_VR && (w != 1440 || h != 1600))
    return fail(ctx, VD3D_ERR_UNSUPPORTED, "VR pack expects 1440x1600 eyes (pad_to_aspect_ratio output)");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t bytes = (size_t)h * w * 3;
  bool sbs = (fmt == VD3D_FMT_HALF_SBS || fmt == VD3D_FMT_FULL_SBS || fmt == VD3D_FMT_VR);
  size_t obytes = sbs ? bytes * 2 : bytes;
  const void *l_d, *r_d;
  int r;
  if ((r = copy_in(ctx, ctx->eyeL, left, bytes, mem, s, &l_d))) return r;
  if ((r = copy_in(ctx, ctx->eyeR, right, bytes, mem, s, &r_d))) return r;
  uint8_t* o_d = dst;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->out_dev[0], obytes))) return r;
    o_d = (uint8_t*)ctx->out_dev[0].p;
  }
  PostArgs pa;
  memset(&pa, 0, sizeof pa);
  pa.left = (const uint8_t*)l_d;
  pa.right = (const uint8_t*)r_d;
  pa.H = h;
  pa.W = w;
  pa.fmt = fmt;
  pa.per_eye_w = w;
  pa.per_eye_h = h;
  pa.fit_w = w;
  pa.fit_h = h;
  pa.sx = pa.sy = 1;
  pa.inv_area = 1.f;
  pa.out = o_d;
  pa.out_w = sbs ? 2 * w : w;
  pa.out_h = h;
  launch_post(pa, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(dst, o_d, obytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

// eye fit on one u8 BGR image: keep_aspect != 0 -> pad_to_aspect_ratio(image, target_w, target_h) with a black canvas
// (core/render_3d.py:101-131); 0 -> cv2.resize(image, (target_w, target_h), interpolation=cv2.INTER_AREA) (1413-1414).
// INTER_AREA shrinking only (identity, integer "area fast", or cv2's general ResizeArea_ tables).
int vd3d_fit_eye(vd3d_ctx* ctx, const uint8_t* src, int h, int w, int target_w, int target_h, int keep_aspect,
                 uint8_t* dst, int mem) {
  if (!ctx || !src || !dst || h < 1 || w < 1 || target_w < 1 || target_h < 1) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t bytes = (size_t)h * w * 3, obytes = (size_t)target_h * target_w * 3;
  const void* s_d;
  int r;
  if ((r = copy_in(ctx, ctx->eyeL, src, bytes, mem, s, &s_d))) return r;
  uint8_t* o_d = dst;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->out_dev[0], obytes))) return r;
    o_d = (uint8_t*)ctx->out_dev[0].p;
  }
  FitPlan fp;
  if ((r = plan_fit(ctx, keep_aspect ? VD3D_FMT_FULL_SBS : VD3D_FMT_HALF_SBS, w, h, target_w, target_h, fp, 1))) return r;
  PostArgs pa;
  memset(&pa, 0, sizeof pa);
  pa.left = (const uint8_t*)s_d;
  pa.right = (const uint8_t*)s_d;
  pa.H = h;
  pa.W = w;
  pa.fmt = VD3D_FMT_INTERLACED;  // single-eye pass-through layout
  pa.per_eye_w = target_w;
  pa.per_eye_h = target_h;
  set_fit(pa, fp);
  pa.out = o_d;
  pa.out_w = target_w;
  pa.out_h = target_h;
  launch_post(pa, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(dst, o_d, obytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

// host-only test hook: the cv2 computeResizeAreaTab restatement used by the fractional INTER_AREA fit.
// ofs/cnt [dsize], alpha [dsize * cap] (zero padded); returns the largest tap count, or a negative error.
int vd3d_area_table(int ssize, int dsize, int* ofs, int* cnt, float* alpha, int cap) {
  if (ssize < 1 || dsize < 1 || dsize > ssize || !ofs || !cnt || !alpha || cap < 1) return VD3D_ERR_ARG;
  std::vector<int> o, c;
  std::vector<std::vector<float>> a;
  area_axis_tab(ssize, dsize, o, c, a);
  int T = 0;
  for (int i = 0; i < dsize; ++i) {
    if (c[i] > cap) return VD3D_ERR_ARG;
    T = c[i] > T ? c[i] : T;
    ofs[i] = o[i];
    cnt[i] = c[i];
    for (int k = 0; k < cap; ++k) alpha[(size_t)i * cap + k] = k < c[i] ? a[i][k] : 0.f;
  }
  return T;
}

// host-only test hook: the fixed-point bilinear tables of cv2's INTER_AREA emulation for an enlarged axis
int vd3d_area_linear_table(int ssize, int dsize, int* ofs, int* a01) {
  if (ssize < 1 || dsize < 1 || !ofs || !a01) return VD3D_ERR_ARG;
  std::vector<int> o, a;
  area_linear_tab(ssize, dsize, o, a);
  memcpy(ofs, o.data(), (size_t)dsize * sizeof(int));
  memcpy(a01, a.data(), (size_t)2 * dsize * sizeof(int));
  return VD3D_OK;
}

int vd3d_dof_grade(vd3d_ctx* ctx, const uint8_t* eye_bgr, int h, int w, const float* depth01, int dh, int dw,
                   double focal, double max_sigma, double sat, double con, double bri, uint8_t* dst, int mem) {
  if (!ctx || !eye_bgr || !dst || h < 2 || w < 2) return fail(ctx, VD3D_ERR_ARG, "bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t bytes = (size_t)h * w * 3;
  const void *s_d, *dp_d = nullptr;
  int r;
  if ((r = copy_in(ctx, ctx->eyeL, eye_bgr, bytes, mem, s, &s_d))) return r;
  bool dof = max_sigma > 0.0;
  if (dof) {
    if (!depth01) return fail(ctx, VD3D_ERR_ARG, "depth required for DOF");
    if ((r = copy_in(ctx, ctx->in_depthf, depth01, sizeof(float) * (size_t)dh * dw, mem, s, &dp_d))) return r;
    if ((r = ensure_dof_kernels(ctx, max_sigma, 5))) return r;
  }
  uint8_t* o_d = dst;
  if (mem == VD3D_MEM_HOST) {
    if ((r = ensure(ctx, ctx->out_dev[0], bytes))) return r;
    o_d = (uint8_t*)ctx->out_dev[0].p;
  }
  DofArgs da;
  memset(&da, 0, sizeof da);
  da.src_l = da.src_r = (const uint8_t*)s_d;
  da.dst_l = da.dst_r = o_d;
  da.H = h;
  da.W = w;
  da.depth = (const float*)dp_d;
  da.dh = dh;
  da.dw = dw;
  da.focal = (float)focal;
  da.focus_w = (float)(0.35 + 1e-6);
  da.idx_max = (float)(5 - 1 - 1e-6);
  da.nlevels = dof ? 5 : 1;
  if (dof) {
    for (int i = 0; i < 8; ++i) {
      da.ksize[i] = ctx->dof_ksize[i];
      da.koff[i] = ctx->dof_koff[i];
    }
    da.kern = (const float*)ctx->dof_kern.p;
    da.halo = ctx->dof_halo;
  }
  da.sat = (float)sat;
  da.con = (float)con;
  da.bri = (float)bri;
  launch_dof(da, 1, s);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) CK(cudaMemcpyAsync(dst, o_d, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return VD3D_OK;
}

}  // extern "C"
