// depth_engine.cu -- Depth-Anything-V2 (DINOv2 ViT + DPT neck/head) forward as a fixed
// launch sequence of tcgen05 GEMMs and small fused kernels.  Host side: named device
// tensors (pushed by the Python loader in HF state_dict terms), TMA tensor maps, buffers.
//
// Reference call path: core/render_depth.py:1106-1119 -> transformers pipeline
// ("depth-estimation") -> DepthAnythingForDepthEstimation.forward (transformers 5.5.0,
// models/depth_anything/modeling_depth_anything.py; backbone models/dinov2/modeling_dinov2.py).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/vd3d.h"
#include "depth_launch.h"
#include "umma_gemm.cuh"

using namespace vd3d;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

bool load_encode() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return false;
  g_encode = (EncodeTiledFn)fn;
  return true;
}

struct DTensor {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct vd3d_depth {
  vd3d_depth_config cfg;
  cudaStream_t stream = nullptr;
  std::string err;
  std::map<std::string, DTensor> w;    // weights by name
  std::map<std::string, DTensor> buf;  // activations by name
  uint64_t launches = 0;
  int ph = 0, pw = 0, ntok = 0, npad = 0;
  bool planned = false;
  bool flash = true;  // fused attention kernel (VD3D_FLASH=0 selects the 3-kernel path)
  // the per-image neck / head tails of a batch run on side streams (forked after the transformer, joined at the end)
  cudaStream_t cur = nullptr;          // stream the helpers launch on (null: `stream`)
  cudaStream_t aux[8] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[8] = {};
  bool owns_weights = true;  // clones share the weight tensors of their parent
  uint64_t weights_version = 0;  // bumped whenever a weight tensor moves (clones / captured graphs hold raw pointers)
  // optional device timing of one GEMM class (the fc1 launches) for the roofline report
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;
  long long prof_rows = 0;  // rows (tokens of all images of the batch) summed over the timed fc1 launches
  // tuning aid (vd3d_depth_profile(e, 2)): device time of every launch class of the forward, eager mode only
  struct Span {
    const char* tag;
    cudaEvent_t e0, e1;
  };
  std::vector<Span> spans;
  int span_begin(const char* tag, cudaStream_t s) {
    if (!prof_spans) return -1;
    Span sp{tag, nullptr, nullptr};
    cudaEventCreate(&sp.e0);
    cudaEventCreate(&sp.e1);
    cudaEventRecord(sp.e0, s);
    spans.push_back(sp);
    return (int)spans.size() - 1;
  }
  void span_end(int i, cudaStream_t s) {
    if (i >= 0) cudaEventRecord(spans[i].e1, s);
  }
  bool prof_spans = false;
};

namespace {

#define DCK(call)                                                                   \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      char _b[512];                                                                 \
      snprintf(_b, sizeof _b, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      e->err = _b;                                                                  \
      return VD3D_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

int dfail(vd3d_depth* e, int code, const std::string& msg) {
  e->err = msg;
  return code;
}

int get_buf(vd3d_depth* e, const std::string& name, size_t bytes, void** out, bool zero = false) {
  DTensor& t = e->buf[name];
  if (t.bytes < bytes) {
    if (t.p) DCK(cudaFree(t.p));
    DCK(cudaMalloc(&t.p, bytes));
    t.bytes = bytes;
    zero = true;
  }
  if (zero) DCK(cudaMemsetAsync(t.p, 0, t.bytes, e->cur ? e->cur : e->stream));
  *out = t.p;
  return VD3D_OK;
}

template <typename T>
int W(vd3d_depth* e, const std::string& name, const T** out, size_t min_count = 0) {
  auto it = e->w.find(name);
  if (it == e->w.end()) return dfail(e, VD3D_ERR_STATE, "missing weight tensor: " + name);
  if (min_count && it->second.bytes < min_count * sizeof(T))
    return dfail(e, VD3D_ERR_STATE, "weight tensor too small: " + name);
  *out = (const T*)it->second.p;
  return VD3D_OK;
}

// f16 3-D tensor map: dims (d0 inner, d1, d2), strides in ELEMENTS for d1 and d2, box (64, b1, b2)
int make_map(vd3d_depth* e, CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
             uint64_t s2, uint32_t b1, uint32_t b2) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * 2, s2 * 2};
  cuuint32_t box[3] = {64, b1, b2};
  cuuint32_t es[3] = {1, 1, 1};
  if ((strides[0] % 16) || (strides[1] % 16) || ((uintptr_t)ptr % 16))
    return dfail(e, VD3D_ERR_ARG, "tensor map: pointer / strides must be 16-byte aligned");
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(64,%u,%u)",
             (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)strides[0], (unsigned long long)strides[1], b1, b2);
    return dfail(e, VD3D_ERR_CUDA, b);
  }
  return VD3D_OK;
}

GemmArgs base_args(int M, int N, int K, int epi) {
  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.M = M;
  g.N = N;
  g.K = K;
  g.epi = epi;
  g.ldc = N;
  return g;
}

int pick_bn(int N) { return N >= 128 ? 128 : (N >= 64 ? 64 : 32); }

// B tensor-map box rows for a tile code (see launch_gemm): the CTA-pair kernels stage half the tile's N per CTA
int box_rows(int bn) { return bn == 256 ? 128 : (bn == -128 ? 64 : (bn == 1128 ? 128 : bn)); }

// VD3D_GEMM_2CTA=1: route large plain GEMMs through the cta_group::2 kernel (256-row tiles)
int pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* v = getenv("VD3D_GEMM_2CTA");
    mode = v ? atoi(v) : 0;
  }
  return mode;
}
// Kernel choice of the token GEMMs (VD3D_GEMM_POLICY=0: everything on the 128x128 single-CTA kernel).
//  * fc2 (fp32 residual epilogue, K = 4D) and the QKV GEMM of the Large model (K = 1024) run on the CTA-pair kernel, 256x256
//    tiles: the 128x128 kernel is bound by the L2->smem operand feed (~20 TB/s device-wide at 64 FLOP/B), the pair tile
//    halves the bytes per FLOP.  In situ, 4 frames per forward: fc2 63 -> 54 us (Base), 95 -> 82 (Large); QKV-Large 60 -> 57.
//    fc1 stays: its bias + GELU epilogue is ALU bound and the pair kernel has half the epilogue warps per SM (54 -> 72 us).
//    The pair kernel's fp32 accumulation is not bit-identical to the single-CTA kernel's, so this choice depends on the
//    model and the per-image token count only, never on the batch size: a frame's depth is the same inferred alone or in
//    a batch (tests/test_depth_gpu.py::test_infer_batch_equals_single_frames; exact sharding == one GPU).
//  * proj (K = D: two tiles per CTA, 1.6 us of MMAs per tile) of a batched forward runs one CTA per SM with 16 epilogue
//    warps, each prefetching its slice of the residual stream while the MMAs run (41 -> 38.6 us; same MMA shape as the
//    default kernel: bit-identical).
int gemm_policy() {
  static int mode = -1;
  if (mode < 0) {
    const char* v = getenv("VD3D_GEMM_POLICY");
    mode = v ? atoi(v) : 1;
  }
  return mode;
}
int pick_bn_gemm(const vd3d_depth* e, int M, int N, int K, int epi) {
  const int pm = pair_mode();
  if (pm == 3) {  // tuning: every large batched GEMM on the pair kernel
    if (M >= 4096 && N >= 768 && (long)N * K >= 1536L * 768) return 256;
    return pick_bn(N);
  }
  if (pm && M >= 512) {
    if (N >= 1536) return 256;
    if (N >= 512) return -128;
  }
  if (gemm_policy() && e->ntok >= 1024 && (N % 256) == 0) {
    if (epi == EPI_RESID_LS && K >= 1536) return 256;
    if (epi == EPI_QKV && K >= 1024) return 256;
  }
  if (gemm_policy() && epi == EPI_RESID_LS && M >= 4096 && N >= 128) return 1128;
  return pick_bn(N);
}
int pick_bn_conv(int M, int N) {  // VD3D_GEMM_2CTA=2: implicit-GEMM convs too (two pixel tiles per pair)
  if (pair_mode() >= 2 && M >= 2048) {
    if (N >= 256) return 256;
    if (N >= 128) return -128;
  }
  return pick_bn(N);
}

// plain GEMM: A [M, K] (lda), B [N, K] (ldb)
int gemm(vd3d_depth* e, const __half* A, int lda, const __half* B, int ldb, GemmArgs g, int bn = 0) {
  if (!bn) bn = pick_bn_gemm(e, g.M, g.N, g.K, g.epi);  // (64-wide tiles for sub-wave grids were measured slower: L2->smem fill bound)
  CUtensorMap ma, mb;
  int r;
  if ((r = make_map(e, &ma, A, g.K, g.M, 1, lda, (uint64_t)lda * g.M, 128, 1))) return r;
  if ((r = make_map(e, &mb, B, g.K, g.N, 1, ldb, (uint64_t)ldb * g.N, box_rows(bn), 1))) return r;
  cudaError_t ce = launch_gemm(bn, ma, mb, g, (g.M + 127) / 128, 1, e->cur ? e->cur : e->stream);
  if (ce != cudaSuccess) return dfail(e, VD3D_ERR_CUDA, std::string("gemm launch: ") + cudaGetErrorString(ce));
  e->launches++;
  return VD3D_OK;
}

// batched GEMM over blockIdx.z: A [batch][M, K] (lda, sa), B [batch][N, K] (ldb, sb)
int gemm_batched(vd3d_depth* e, const __half* A, int lda, uint64_t sa, const __half* B, int ldb, uint64_t sb,
                 int batch, GemmArgs g, int bn) {
  CUtensorMap ma, mb;
  int r;
  if ((r = make_map(e, &ma, A, g.K, g.M, batch, lda, sa, 128, 1))) return r;
  if ((r = make_map(e, &mb, B, g.K, g.N, batch, ldb, sb, box_rows(bn), 1))) return r;
  cudaError_t ce = launch_gemm(bn, ma, mb, g, (g.M + 127) / 128, batch, e->cur ? e->cur : e->stream);
  if (ce != cudaSuccess) return dfail(e, VD3D_ERR_CUDA, std::string("gemm launch: ") + cudaGetErrorString(ce));
  e->launches++;
  return VD3D_OK;
}

void pick_tile(int W, int H, int& tw, int& th) {
  const int cand[4][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}};
  long best = -1;
  for (auto& c : cand) {
    long cover = (long)((W + c[0] - 1) / c[0]) * c[0] * (long)((H + c[1] - 1) / c[1]) * c[1];
    if (best < 0 || cover < best) {
      best = cover;
      tw = c[0];
      th = c[1];
    }
  }
}

// conv over an NHWC f16 map [H, W, cin] (3x3 pad 1 when k3, else 1x1); weights [N, taps*cin]
int conv(vd3d_depth* e, const __half* in, int H, int W, int cin, const __half* wt, bool k3, GemmArgs g, int bn = 0) {
  if (cin % 64) return dfail(e, VD3D_ERR_ARG, "conv: cin must be a multiple of 64");
  if (!bn) bn = pick_bn_conv(H * W, g.N);
  int tw, th;
  pick_tile(W, H, tw, th);
  g.conv = k3 ? 1 : 2;
  g.cin = cin;
  g.imgW = W;
  g.imgH = H;
  g.tw = tw;
  g.th = th;
  g.M = H * W;
  g.K = (k3 ? 9 : 1) * cin;
  CUtensorMap ma, mb;
  int r;
  if ((r = make_map(e, &ma, in, cin, W, H, cin, (uint64_t)cin * W, tw, th))) return r;
  if ((r = make_map(e, &mb, wt, g.K, g.N, 1, g.K, (uint64_t)g.K * g.N, box_rows(bn), 1))) return r;
  int m_tiles = ((W + tw - 1) / tw) * ((H + th - 1) / th);
  cudaError_t ce = launch_gemm(bn, ma, mb, g, m_tiles, 1, e->cur ? e->cur : e->stream);
  if (ce != cudaSuccess) return dfail(e, VD3D_ERR_CUDA, std::string("conv launch: ") + cudaGetErrorString(ce));
  e->launches++;
  return VD3D_OK;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

extern "C" {

int vd3d_depth_create(const vd3d_depth_config* cfg, void* stream, vd3d_depth** out) {
  if (!cfg || !out) return VD3D_ERR_ARG;
  *out = nullptr;
  if (!load_encode()) return VD3D_ERR_CUDA;
  if (cfg->hidden % 128 || cfg->hidden > 1024 || cfg->hidden / cfg->heads != 64 || cfg->fusion % 64 ||
      cfg->image_h % 14 || cfg->image_w % 14)
    return VD3D_ERR_ARG;
  vd3d_depth* e = new vd3d_depth();
  e->cfg = *cfg;
  e->stream = (cudaStream_t)stream;
  e->ph = cfg->image_h / 14;
  e->pw = cfg->image_w / 14;
  e->ntok = e->ph * e->pw + 1;
  e->npad = round_up(e->ntok, 128);
  if (const char* f = getenv("VD3D_FLASH")) e->flash = atoi(f) != 0;
  if (e->ntok > 3072) {
    delete e;
    return VD3D_ERR_UNSUPPORTED;
  }
  *out = e;
  return VD3D_OK;
}

// device timing of the fc1 GEMM launches (k_umma_gemm<128,3>, M=tokens, N=4D, K=D): bench.py roofline
int vd3d_depth_profile(vd3d_depth* e, int enable) {
  if (!e) return VD3D_ERR_ARG;
  e->prof = enable == 1;
  e->prof_spans = enable == 2;
  return VD3D_OK;
}
// "tag total_ms launches" lines of the spans recorded since the last call (vd3d_depth_profile(e, 2))
int vd3d_depth_profile_spans(vd3d_depth* e, char* out, size_t cap) {
  if (!e || !out || !cap) return VD3D_ERR_ARG;
  DCK(cudaDeviceSynchronize());
  std::vector<std::pair<std::string, std::pair<double, int>>> acc;
  for (auto& sp : e->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) != cudaSuccess) ms = 0.f;
    cudaEventDestroy(sp.e0);
    cudaEventDestroy(sp.e1);
    size_t i = 0;
    for (; i < acc.size(); ++i)
      if (acc[i].first == sp.tag) break;
    if (i == acc.size()) acc.push_back({sp.tag, {0.0, 0}});
    acc[i].second.first += ms;
    acc[i].second.second++;
  }
  e->spans.clear();
  std::string txt;
  char line[128];
  for (auto& a : acc) {
    snprintf(line, sizeof line, "%s %.4f %d\n", a.first.c_str(), a.second.first, a.second.second);
    txt += line;
  }
  snprintf(out, cap, "%s", txt.c_str());
  return VD3D_OK;
}
int vd3d_depth_profile_collect(vd3d_depth* e, double* total_ms, int* count, double* gflop_per_launch) {
  if (!e || !total_ms || !count) return VD3D_ERR_ARG;
  DCK(cudaStreamSynchronize(e->stream));
  double t = 0;
  int n = 0;
  for (size_t i = 0; i + 1 < e->prof_ev.size(); i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->prof_ev[i], e->prof_ev[i + 1]) == cudaSuccess) {
      t += ms;
      ++n;
    }
  }
  for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
  e->prof_ev.clear();
  *total_ms = t;
  *count = n;
  if (gflop_per_launch)  // average over the timed launches (batches of different sizes have different row counts)
    *gflop_per_launch = 2.0 * (n ? (double)e->prof_rows / n : (double)e->ntok) * (4.0 * e->cfg.hidden) * e->cfg.hidden / 1e9;
  e->prof_rows = 0;
  return VD3D_OK;
}

// A second engine instance on another stream that shares the parent's weights (own activation
// buffers): lets two frames' depth forwards overlap on the GPU (vd3d_render_clip_depth).
int vd3d_depth_clone(vd3d_depth* src, void* stream, vd3d_depth** out) {
  if (!src || !out) return VD3D_ERR_ARG;
  vd3d_depth* e = new vd3d_depth();
  e->cfg = src->cfg;
  e->stream = (cudaStream_t)stream;
  e->w = src->w;
  e->owns_weights = false;
  e->ph = src->ph;
  e->pw = src->pw;
  e->ntok = src->ntok;
  e->npad = src->npad;
  e->flash = src->flash;
  *out = e;
  return VD3D_OK;
}

void vd3d_depth_destroy(vd3d_depth* e) {
  if (!e) return;
  cudaStreamSynchronize(e->stream);
  for (int i = 0; i < 8; ++i) {
    if (e->aux[i]) {
      cudaStreamSynchronize(e->aux[i]);
      cudaStreamDestroy(e->aux[i]);
    }
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  for (auto& kv : e->w)
    if (e->owns_weights && kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : e->buf)
    if (kv.second.p) cudaFree(kv.second.p);
  delete e;
}

const char* vd3d_depth_last_error(vd3d_depth* e) { return e ? e->err.c_str() : "null depth engine"; }
uint64_t vd3d_depth_launch_count(vd3d_depth* e) { return e ? e->launches : 0; }
void vd3d_depth_add_launches(vd3d_depth* e, uint64_t n) {
  if (e) e->launches += n;  // graph replays account for the launches they contain
}

int vd3d_depth_set_tensor(vd3d_depth* e, const char* name, const void* host_data, size_t bytes) {
  if (!e || !name || !host_data || !bytes) return VD3D_ERR_ARG;
  if (!e->owns_weights) return dfail(e, VD3D_ERR_STATE, "weights of an engine clone belong to its parent");
  DTensor& t = e->w[name];
  const size_t cap = (bytes + 255) / 256 * 256;
  if (t.p && (t.bytes + 255) / 256 * 256 == cap) {
    // same footprint: overwrite in place so that clones and captured graphs (raw pointers) stay valid; work already
    // enqueued on any stream finishes first
    DCK(cudaDeviceSynchronize());
    t.bytes = bytes;
    DCK(cudaMemcpy(t.p, host_data, bytes, cudaMemcpyHostToDevice));
    return VD3D_OK;
  }
  if (t.p) {
    DCK(cudaDeviceSynchronize());
    cudaFree(t.p);
  }
  t.p = nullptr;
  e->weights_version++;  // the tensor moves: vd3d_ctx rebuilds its clones and drops its graphs (vd3d_api.cu)
  DCK(cudaMalloc(&t.p, cap));
  t.bytes = bytes;
  DCK(cudaMemcpy(t.p, host_data, bytes, cudaMemcpyHostToDevice));
  return VD3D_OK;
}
uint64_t vd3d_depth_weights_version(vd3d_depth* e) { return e ? e->weights_version : 0; }

int vd3d_depth_get_buffer(vd3d_depth* e, const char* name, void* host_out, size_t bytes) {
  if (!e || !name || !host_out) return VD3D_ERR_ARG;
  auto it = e->buf.find(name);
  if (it == e->buf.end()) return dfail(e, VD3D_ERR_ARG, std::string("no such buffer: ") + name);
  if (bytes > it->second.bytes) return dfail(e, VD3D_ERR_ARG, "buffer smaller than requested");
  DCK(cudaStreamSynchronize(e->stream));
  DCK(cudaMemcpy(host_out, it->second.p, bytes, cudaMemcpyDeviceToHost));
  return VD3D_OK;
}

// unit-test hook: C[M,N] f32 = A[M,K] f16 x B[N,K]^T f16 through the tcgen05 kernel
int vd3d_gemm_f16(vd3d_depth* e, const void* A, const void* B, int M, int N, int K, float* C_host, int bn) {
  if (!e || !A || !B || !C_host || (K % 8)) return VD3D_ERR_ARG;
  void *da, *db, *dc;
  int r;
  if ((r = get_buf(e, "t.a", (size_t)M * K * 2, &da))) return r;
  if ((r = get_buf(e, "t.b", (size_t)N * K * 2, &db))) return r;
  if ((r = get_buf(e, "t.c", (size_t)M * N * 4, &dc, true))) return r;
  DCK(cudaMemcpyAsync(da, A, (size_t)M * K * 2, cudaMemcpyHostToDevice, e->stream));
  DCK(cudaMemcpyAsync(db, B, (size_t)N * K * 2, cudaMemcpyHostToDevice, e->stream));
  GemmArgs g = base_args(M, N, K, EPI_F32);
  g.out_f32 = (float*)dc;
  if ((r = gemm(e, (const __half*)da, K, (const __half*)db, K, g, bn))) return r;
  DCK(cudaMemcpyAsync(C_host, dc, (size_t)M * N * 4, cudaMemcpyDeviceToHost, e->stream));
  DCK(cudaStreamSynchronize(e->stream));
  return VD3D_OK;
}

// tuning hook: time one GEMM shape (EPI_F16 epilogue, optional GELU) on device-resident operands with an explicit
// kernel variant (launch_gemm_variant) and debug mode (GemmArgs::dbg); returns the average launch time in ms
int vd3d_gemm_bench(vd3d_depth* e, int M, int N, int K, int variant, int dbg, int act, int iters, float* ms_out) {
  if (!e || !ms_out || (K % 8) || iters < 1) return VD3D_ERR_ARG;
  void *da, *db, *dc;
  int r;
  if ((r = get_buf(e, "t.ba", (size_t)M * K * 2, &da))) return r;
  if ((r = get_buf(e, "t.bb", (size_t)N * K * 2, &db))) return r;
  if ((r = get_buf(e, "t.bc", (size_t)M * N * 2, &dc))) return r;
  DCK(cudaMemsetAsync(da, 0x2c, (size_t)M * K * 2, e->stream));  // 0x2c2c = 0.0652 in f16
  DCK(cudaMemsetAsync(db, 0x2c, (size_t)N * K * 2, e->stream));
  GemmArgs g = base_args(M, N, K, EPI_F16);
  g.out_f16 = (__half*)dc;
  g.act = act & 0xff;
  g.dbg = dbg;
  if (act & 0x100) {  // the proj / fc2 epilogue: fp32 residual stream read-modify-write with bias and LayerScale
    void *dx, *dv;
    if ((r = get_buf(e, "t.bx", (size_t)M * N * 4, &dx))) return r;
    if ((r = get_buf(e, "t.bv", (size_t)N * 4, &dv))) return r;  // zeros: bias = ls = 0 keeps x finite over the iterations
    g.epi = EPI_RESID_LS;
    g.act = 0;
    g.out_f32 = (float*)dx;
    g.bias = (const float*)dv;
    g.ls = (const float*)dv;
    g.ldc = N;
  }
  CUtensorMap ma, mb;
  if ((r = make_map(e, &ma, da, K, M, 1, K, (uint64_t)K * M, 128, 1))) return r;
  if ((r = make_map(e, &mb, db, K, N, 1, K, (uint64_t)K * N, variant >= 20 ? 64 : 128, 1))) return r;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto done = [&](int rc) {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
  };
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)
    return done(dfail(e, VD3D_ERR_CUDA, "gemm bench: event creation failed"));
  cudaError_t ce = cudaSuccess;
  for (int i = 0; i < 3 + iters && ce == cudaSuccess; ++i) {
    if (i == 3) ce = cudaEventRecord(e0, e->stream);
    if (ce == cudaSuccess) ce = launch_gemm_variant(variant, ma, mb, g, (M + 127) / 128, 1, e->stream);
  }
  if (ce == cudaSuccess) ce = cudaEventRecord(e1, e->stream);
  if (ce == cudaSuccess) ce = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e0, e1);
  if (ce != cudaSuccess) return done(dfail(e, VD3D_ERR_CUDA, std::string("gemm bench: ") + cudaGetErrorString(ce)));
  done(0);
  *ms_out = ms / iters;
  return VD3D_OK;
}

// unit-test hook: 3x3 (or 1x1) conv on an NHWC f16 map through the implicit-GEMM path
int vd3d_conv_f16(vd3d_depth* e, const void* in_nhwc, int H, int W, int cin, const void* wt, int cout, int k3,
                  const float* bias, int relu, float* out_host) {
  if (!e || !in_nhwc || !wt || !out_host) return VD3D_ERR_ARG;
  int taps = k3 ? 9 : 1;
  void *di, *dw, *dob, *dbias = nullptr;
  int r;
  if ((r = get_buf(e, "t.ci", (size_t)H * W * cin * 2, &di))) return r;
  if ((r = get_buf(e, "t.cw", (size_t)cout * taps * cin * 2, &dw))) return r;
  if ((r = get_buf(e, "t.co", (size_t)H * W * cout * 2, &dob, true))) return r;
  DCK(cudaMemcpyAsync(di, in_nhwc, (size_t)H * W * cin * 2, cudaMemcpyHostToDevice, e->stream));
  DCK(cudaMemcpyAsync(dw, wt, (size_t)cout * taps * cin * 2, cudaMemcpyHostToDevice, e->stream));
  if (bias) {
    if ((r = get_buf(e, "t.cb", (size_t)cout * 4, &dbias))) return r;
    DCK(cudaMemcpyAsync(dbias, bias, (size_t)cout * 4, cudaMemcpyHostToDevice, e->stream));
  }
  GemmArgs g = base_args(H * W, cout, taps * cin, EPI_F16);
  g.out_f16 = (__half*)dob;
  g.bias = (const float*)dbias;
  g.act = relu ? 2 : 0;
  if ((r = conv(e, (const __half*)di, H, W, cin, (const __half*)dw, k3 != 0, g))) return r;
  std::vector<__half> tmp((size_t)H * W * cout);
  DCK(cudaMemcpyAsync(tmp.data(), dob, tmp.size() * 2, cudaMemcpyDeviceToHost, e->stream));
  DCK(cudaStreamSynchronize(e->stream));
  for (size_t i = 0; i < tmp.size(); ++i) out_host[i] = __half2float(tmp[i]);
  return VD3D_OK;
}

}  // extern "C"

namespace {

constexpr int kMaxBatch = 8;

// B images through the network in one pass.  The token-wise stages (LayerNorm, QKV / proj / fc1 / fc2 GEMMs) see one
// matrix of (B-1)*NP + NT rows -- image b owns rows [b*NP, b*NP + NT), NP = NT rounded up to the 128-row tile -- so the
// GEMMs run at 3-4x the rows of one frame (the reference batches too: core/render_depth.py:1113-1119); attention runs
// with grid.y = B * heads; patch embedding, taps, neck and head stay per image (their maps are per-image 2-D tensors).
// px_dev / depth_dev: DEVICE pointers, f32 [3, image_h, image_w] in, f32 [image_h, image_w] out.
int forward_core(vd3d_depth* e, int B, const float* const* px_dev, float* const* depth_dev) {
  if (B < 1 || B > kMaxBatch) return dfail(e, VD3D_ERR_ARG, "batch must be in [1, 8]");
  e->cur = nullptr;  // (an earlier call may have failed inside its forked section)
  const vd3d_depth_config& c = e->cfg;
  cudaStream_t s = e->stream;
  const int D = c.hidden, L = c.layers, Hh = c.heads, F = c.fusion;
  const int ph = e->ph, pw = e->pw, NT = e->ntok, NP = e->npad, NPATCH = ph * pw;
  const int IH = c.image_h, IW = c.image_w;
  const int KPE = 592;
  int r;
  char nm[96];
  // ---- buffers ----
  void *x, *xn, *q, *k, *vt, *S, *P, *attn, *hb, *ape;
  const int MT = (B - 1) * NP + NT;  // rows of the stacked token matrix
  if (B > 1 && !e->flash) return dfail(e, VD3D_ERR_UNSUPPORTED, "batched forward needs the fused attention kernel");
  if ((r = get_buf(e, "x", (size_t)B * NP * D * 4, &x))) return r;
  if ((r = get_buf(e, "xn", (size_t)B * NP * D * 2, &xn))) return r;
  if ((r = get_buf(e, "q", (size_t)B * Hh * NP * 64 * 2, &q))) return r;
  if ((r = get_buf(e, "k", (size_t)B * Hh * NP * 64 * 2, &k))) return r;
  if ((r = get_buf(e, "vt", (size_t)B * Hh * 64 * NP * 2, &vt))) return r;
  S = P = nullptr;
  if (!e->flash) {
    if ((r = get_buf(e, "S", (size_t)Hh * NP * NP * 4, &S))) return r;
    if ((r = get_buf(e, "P", (size_t)Hh * NP * NP * 2, &P))) return r;
  }
  if ((r = get_buf(e, "attn", (size_t)B * NP * D * 2, &attn))) return r;
  if ((r = get_buf(e, "h", (size_t)B * NP * 4 * D * 2, &hb))) return r;
  if ((r = get_buf(e, "ape", (size_t)NPATCH * KPE * 2, &ape))) return r;

  // ---- patch embedding + CLS + position embeddings ----
  const int sp_embed = e->span_begin("embed", s);
  const __half* pe_w;
  const float *pe_b, *cls, *pos;
  if ((r = W(e, "pe.w", &pe_w, (size_t)D * KPE)) || (r = W(e, "pe.b", &pe_b, D)) || (r = W(e, "cls", &cls, D)) ||
      (r = W(e, "pos", &pos, (size_t)NT * D)))
    return r;
  for (int b = 0; b < B; ++b) {
    float* xb = (float*)x + (size_t)b * NP * D;
    launch_patch_im2col(px_dev[b], IH, IW, ph, pw, (__half*)ape, KPE, s);
    GemmArgs g = base_args(NPATCH, D, KPE, EPI_PATCH);
    g.out_f32 = xb;
    g.bias = pe_b;
    g.pos = pos;
    g.ldc = D;
    if ((r = gemm(e, (const __half*)ape, KPE, pe_w, KPE, g))) return r;
    launch_set_cls(xb, cls, pos, D, s);
    e->launches += 2;
    // rows [NT, NP) of every image but the last take part in the stacked GEMMs: keep them at zero so that they stay
    // finite from frame to frame (they are never read as keys / values: the attention tensor maps end at NT)
    if (b + 1 < B) DCK(cudaMemsetAsync(xb + (size_t)NT * D, 0, (size_t)(NP - NT) * D * 4, s));
  }

  e->span_end(sp_embed, s);
  // ---- transformer blocks ----
  int tap_idx = 0;
  for (int l = 0; l < L; ++l) {
    const float *g1, *b1, *g2, *b2, *bqkv, *bo, *ls1, *bf1, *bf2, *ls2;
    const __half *wqkv, *wo, *wf1, *wf2;
    auto nmf = [&](const char* suffix) {
      snprintf(nm, sizeof nm, "l%d.%s", l, suffix);
      return std::string(nm);
    };
    if ((r = W(e, nmf("ln1.g"), &g1, D)) || (r = W(e, nmf("ln1.b"), &b1, D)) || (r = W(e, nmf("ln2.g"), &g2, D)) ||
        (r = W(e, nmf("ln2.b"), &b2, D)) || (r = W(e, nmf("qkv.w"), &wqkv, (size_t)3 * D * D)) ||
        (r = W(e, nmf("qkv.b"), &bqkv, 3 * D)) || (r = W(e, nmf("proj.w"), &wo, (size_t)D * D)) ||
        (r = W(e, nmf("proj.b"), &bo, D)) || (r = W(e, nmf("ls1"), &ls1, D)) ||
        (r = W(e, nmf("fc1.w"), &wf1, (size_t)4 * D * D)) || (r = W(e, nmf("fc1.b"), &bf1, 4 * D)) ||
        (r = W(e, nmf("fc2.w"), &wf2, (size_t)4 * D * D)) || (r = W(e, nmf("fc2.b"), &bf2, D)) ||
        (r = W(e, nmf("ls2"), &ls2, D)))
      return r;
    int sp_ = e->span_begin("ln", s);
    launch_layernorm((const float*)x, MT, D, g1, b1, (__half*)xn, 0, s);
    e->span_end(sp_, s);
    sp_ = e->span_begin("qkv", s);
    {
      GemmArgs g = base_args(MT, 3 * D, D, EPI_QKV);
      g.bias = bqkv;
      g.q = (__half*)q;
      g.k = (__half*)k;
      g.vt = (__half*)vt;
      g.heads = Hh;
      g.npad = NP;
      g.dmodel = D;
      g.qscale = 0.125f;  // 1/sqrt(64), exact in f16
      if ((r = gemm(e, (const __half*)xn, D, wqkv, D, g))) return r;
    }
    e->span_end(sp_, s);
    sp_ = e->span_begin("attn", s);
    if (e->flash) {
      // fused tcgen05 attention: scores stay in TMEM, probabilities in shared memory
      CUtensorMap mq, mk, mv;
      if ((r = make_map(e, &mq, q, 64, NT, B * Hh, 64, (uint64_t)NP * 64, 128, 1))) return r;
      if ((r = make_map(e, &mk, k, 64, NT, B * Hh, 64, (uint64_t)NP * 64, 128, 1))) return r;
      if ((r = make_map(e, &mv, vt, NT, 64, B * Hh, NP, (uint64_t)64 * NP, 64, 1))) return r;
      cudaError_t ce = launch_attention(mq, mk, mv, NT, D, (__half*)attn, Hh, B, NP, s);
      if (ce != cudaSuccess) return dfail(e, VD3D_ERR_CUDA, std::string("attention launch: ") + cudaGetErrorString(ce));
      e->launches++;
    } else {
      {  // scores = (q / sqrt(d)) k^T, per head
        GemmArgs g = base_args(NT, NT, 64, EPI_F32);
        g.out_f32 = (float*)S;
        g.ldc = NP;
        g.out_batch_stride = (long long)NP * NP;
        if ((r = gemm_batched(e, (const __half*)q, 64, (uint64_t)NP * 64, (const __half*)k, 64, (uint64_t)NP * 64, Hh,
                              g, 128)))
          return r;
      }
      launch_softmax((const float*)S, (__half*)P, NT, Hh, NT, NP, s);
      e->launches++;
      {  // context = P v, written head-interleaved into attn [NT, D]
        GemmArgs g = base_args(NT, 64, NT, EPI_F16);
        g.out_f16 = (__half*)attn;
        g.ldc = D;
        g.out_batch_stride = 64;
        if ((r = gemm_batched(e, (const __half*)P, NP, (uint64_t)NP * NP, (const __half*)vt, NP, (uint64_t)64 * NP, Hh,
                              g, 64)))
          return r;
      }
    }
    e->span_end(sp_, s);
    sp_ = e->span_begin("proj", s);
    {
      GemmArgs g = base_args(MT, D, D, EPI_RESID_LS);
      g.out_f32 = (float*)x;
      g.bias = bo;
      g.ls = ls1;
      g.ldc = D;
      if ((r = gemm(e, (const __half*)attn, D, wo, D, g))) return r;
    }
    e->span_end(sp_, s);
    sp_ = e->span_begin("ln", s);
    launch_layernorm((const float*)x, MT, D, g2, b2, (__half*)xn, 0, s);
    e->span_end(sp_, s);
    sp_ = e->span_begin("fc1", s);
    {
      GemmArgs g = base_args(MT, 4 * D, D, EPI_F16);
      g.out_f16 = (__half*)hb;
      g.bias = bf1;
      g.act = 1;
      g.ldc = 4 * D;
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (e->prof) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        e->prof_rows += MT;
      }
      if ((r = gemm(e, (const __half*)xn, D, wf1, D, g))) return r;
      if (e->prof) {
        cudaEventRecord(e1, s);
        e->prof_ev.push_back(e0);
        e->prof_ev.push_back(e1);
      }
    }
    e->span_end(sp_, s);
    sp_ = e->span_begin("fc2", s);
    {
      GemmArgs g = base_args(MT, D, 4 * D, EPI_RESID_LS);
      g.out_f32 = (float*)x;
      g.bias = bf2;
      g.ls = ls2;
      g.ldc = D;
      if ((r = gemm(e, (const __half*)hb, 4 * D, wf2, 4 * D, g))) return r;
    }
    e->span_end(sp_, s);
    e->launches += 2;
    if (tap_idx < 4 && c.taps[tap_idx] == l + 1) {
      // backbone output: final LayerNorm applied (apply_layernorm=True), CLS dropped by the neck
      const float *ng, *nb;
      if ((r = W(e, "norm.g", &ng, D)) || (r = W(e, "norm.b", &nb, D))) return r;
      for (int b = 0; b < B; ++b) {
        void* tp;
        snprintf(nm, sizeof nm, "tap%d.%d", tap_idx, b);
        if ((r = get_buf(e, nm, (size_t)round_up(NPATCH, 128) * D * 2, &tp))) return r;
        launch_layernorm((const float*)x + (size_t)b * NP * D, NPATCH, D, ng, nb, (__half*)tp, 1, s);
        e->launches++;
      }
      ++tap_idx;
    }
  }
  if (tap_idx != 4) return dfail(e, VD3D_ERR_ARG, "taps must be increasing layer indices <= layers");

  // neck + head: independent per image.  With more than one image each tail runs on its own side stream (forked here,
  // joined below) with its own activation buffers, so that the many small launches of the tails (grids of 5..77 CTAs on
  // the coarse maps) overlap each other instead of leaving most SMs idle; captured into the same CUDA graph.
  const bool fork = B > 1;
  cudaStream_t s_main = s;
  const int sp_tail = e->span_begin("tail(all images, forked)", s_main);
  if (fork) {
    if (!e->ev_fork) DCK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    DCK(cudaEventRecord(e->ev_fork, s_main));
  }
  for (int b = 0; b < B; ++b) {
  if (fork) {
    if (!e->aux[b]) DCK(cudaStreamCreateWithFlags(&e->aux[b], cudaStreamNonBlocking));
    if (!e->ev_join[b]) DCK(cudaEventCreateWithFlags(&e->ev_join[b], cudaEventDisableTiming));
    s = e->aux[b];
    e->cur = s;
    DCK(cudaStreamWaitEvent(s, e->ev_fork, 0));
  }
  auto bname = [&](const char* base) {  // per-image activation buffers when the tails run concurrently
    std::string n(base);
    if (fork) n += "#" + std::to_string(b);
    return n;
  };
  void* depth_d = depth_dev[b];
  // ---- neck: reassemble + 3x3 conv to the fusion width ----
  int fh[4], fw[4];
  void* feat[4];
  for (int i = 0; i < 4; ++i) {
    const int C = c.neck[i], CP = round_up(C, 64);
    const __half *pw_, *cw;
    const float* pb;
    snprintf(nm, sizeof nm, "r%d.proj.w", i);
    if ((r = W(e, nm, &pw_, (size_t)CP * D))) return r;
    snprintf(nm, sizeof nm, "r%d.proj.b", i);
    if ((r = W(e, nm, &pb, CP))) return r;
    snprintf(nm, sizeof nm, "tap%d.%d", i, b);
    void *tp = e->buf[nm].p, *rp, *rs = nullptr;
    snprintf(nm, sizeof nm, "r%d.p", i);
    if ((r = get_buf(e, bname(nm), (size_t)round_up(NPATCH, 128) * CP * 2, &rp))) return r;
    {
      GemmArgs g = base_args(NPATCH, CP, D, EPI_F16);
      g.out_f16 = (__half*)rp;
      g.bias = pb;
      g.ldc = CP;
      if ((r = gemm(e, (const __half*)tp, D, pw_, D, g))) return r;
    }
    if (i < 2) {  // ConvTranspose2d(kernel=stride=4 / 2)
      const int kk = (i == 0) ? 4 : 2;
      const __half* uw;
      const float* ub;
      snprintf(nm, sizeof nm, "r%d.up.w", i);
      if ((r = W(e, nm, &uw, (size_t)kk * kk * CP * CP))) return r;
      snprintf(nm, sizeof nm, "r%d.up.b", i);
      if ((r = W(e, nm, &ub, (size_t)kk * kk * CP))) return r;
      fh[i] = ph * kk;
      fw[i] = pw * kk;
      snprintf(nm, sizeof nm, "r%d.s", i);
      if ((r = get_buf(e, bname(nm), (size_t)fh[i] * fw[i] * CP * 2, &rs))) return r;
      GemmArgs g = base_args(NPATCH, kk * kk * CP, CP, EPI_CONVT);
      g.out_f16 = (__half*)rs;
      g.bias = ub;
      g.ct_k = kk;
      g.ct_cout = CP;
      g.ct_w = pw;
      if ((r = gemm(e, (const __half*)rp, CP, uw, CP, g))) return r;
    } else if (i == 2) {
      fh[i] = ph;
      fw[i] = pw;
      rs = rp;
    } else {  // Conv2d(3x3, stride 2, pad 1)
      const __half* dw;
      const float* db;
      if ((r = W(e, "r3.down.w", &dw, (size_t)CP * 9 * CP)) || (r = W(e, "r3.down.b", &db, CP))) return r;
      fh[i] = (ph - 1) / 2 + 1;
      fw[i] = (pw - 1) / 2 + 1;
      void* col;
      if ((r = get_buf(e, bname("r3.col"), (size_t)round_up(fh[i] * fw[i], 128) * 9 * CP * 2, &col))) return r;
      launch_im2col_s2((const __half*)rp, ph, pw, CP, CP, (__half*)col, fh[i], fw[i], s);
      e->launches++;
      if ((r = get_buf(e, bname("r3.s"), (size_t)round_up(fh[i] * fw[i], 128) * CP * 2, &rs))) return r;
      GemmArgs g = base_args(fh[i] * fw[i], CP, 9 * CP, EPI_F16);
      g.out_f16 = (__half*)rs;
      g.bias = db;
      g.ldc = CP;
      if ((r = gemm(e, (const __half*)col, 9 * CP, dw, 9 * CP, g))) return r;
    }
    snprintf(nm, sizeof nm, "n%d.conv.w", i);
    if ((r = W(e, nm, &cw, (size_t)F * 9 * CP))) return r;
    snprintf(nm, sizeof nm, "f%d", i);
    if ((r = get_buf(e, bname(nm), (size_t)fh[i] * fw[i] * F * 2, &feat[i]))) return r;
    GemmArgs g = base_args(0, F, 0, EPI_F16);
    g.out_f16 = (__half*)feat[i];
    g.ldc = F;
    if ((r = conv(e, (const __half*)rs, fh[i], fw[i], CP, cw, true, g))) return r;
  }

  // ---- fusion stage (coarsest first) ----
  void* fused = nullptr;
  int ch = 0, cw_ = 0;
  for (int j = 0; j < 4; ++j) {
    const int fi = 3 - j;
    const int Hc = fh[fi], Wc = fw[fi];
    const size_t n = (size_t)Hc * Wc * F;
    void *t_relu, *t_mid, *hraw, *hrelu, *yraw, *up, *prj;
    if ((r = get_buf(e, bname("fu.relu"), n * 2, &t_relu)) || (r = get_buf(e, bname("fu.mid"), n * 2, &t_mid)) ||
        (r = get_buf(e, bname("fu.h"), n * 2, &hraw)) || (r = get_buf(e, bname("fu.hrelu"), n * 2, &hrelu)) ||
        (r = get_buf(e, bname("fu.y"), n * 2, &yraw)))
      return r;
    auto cv = [&](const char* unit, const char* which, const __half** wv, const float** bv) -> int {
      snprintf(nm, sizeof nm, "f%d.%s.%s.w", j, unit, which);
      int rr = W(e, nm, wv, (size_t)F * 9 * F);
      if (rr) return rr;
      snprintf(nm, sizeof nm, "f%d.%s.%s.b", j, unit, which);
      return W(e, nm, bv, F);
    };
    const __half *w1, *w2;
    const float *bb1, *bb2;
    const __half* hin_raw;
    if (j == 0) {
      hin_raw = (const __half*)feat[fi];
      launch_relu_f16(hin_raw, (__half*)hrelu, n, s);
      e->launches++;
    } else {
      // h = fused + residual_layer1(feat)
      if ((r = cv("rl1", "c1", &w1, &bb1)) || (r = cv("rl1", "c2", &w2, &bb2))) return r;
      launch_relu_f16((const __half*)feat[fi], (__half*)t_relu, n, s);
      e->launches++;
      GemmArgs g1 = base_args(0, F, 0, EPI_F16);
      g1.out_f16 = (__half*)t_mid;
      g1.bias = bb1;
      g1.act = 2;
      g1.ldc = F;
      if ((r = conv(e, (const __half*)t_relu, Hc, Wc, F, w1, true, g1))) return r;
      // y = conv2(mid) + b2 + feat ; h = y + fused  (two residual adds: second via res on a 1x1-free pass)
      GemmArgs g2 = base_args(0, F, 0, EPI_F16);
      g2.out_f16 = (__half*)yraw;
      g2.bias = bb2;
      g2.res_f16 = (const __half*)feat[fi];
      g2.ldc = F;
      if ((r = conv(e, (const __half*)t_mid, Hc, Wc, F, w2, true, g2))) return r;
      // h = fused + y ; relu(h) feeds residual_layer2's first conv
      launch_add_relu_f16((const __half*)fused, (const __half*)yraw, (__half*)hraw, (__half*)hrelu, n, s);
      e->launches++;
      hin_raw = (const __half*)hraw;
    }
    // residual_layer2
    if ((r = cv("rl2", "c1", &w1, &bb1)) || (r = cv("rl2", "c2", &w2, &bb2))) return r;
    {
      GemmArgs g1 = base_args(0, F, 0, EPI_F16);
      g1.out_f16 = (__half*)t_mid;
      g1.bias = bb1;
      g1.act = 2;
      g1.ldc = F;
      if ((r = conv(e, (const __half*)hrelu, Hc, Wc, F, w1, true, g1))) return r;
      GemmArgs g2 = base_args(0, F, 0, EPI_F16);
      g2.out_f16 = (__half*)yraw;
      g2.bias = bb2;
      g2.res_f16 = hin_raw;
      g2.ldc = F;
      if ((r = conv(e, (const __half*)t_mid, Hc, Wc, F, w2, true, g2))) return r;
    }
    // upsample (to the next feature's size, or x2 at the end), then 1x1 projection
    int OH = (j < 3) ? fh[fi - 1] : Hc * 2, OW = (j < 3) ? fw[fi - 1] : Wc * 2;
    snprintf(nm, sizeof nm, "fu.up%d", j);
    if ((r = get_buf(e, bname(nm), (size_t)OH * OW * F * 2, &up))) return r;
    launch_upsample_ac((const __half*)yraw, Hc, Wc, F, (__half*)up, OH, OW, s);
    e->launches++;
    const __half* pwt;
    const float* pbs;
    snprintf(nm, sizeof nm, "f%d.proj.w", j);
    if ((r = W(e, nm, &pwt, (size_t)F * F))) return r;
    snprintf(nm, sizeof nm, "f%d.proj.b", j);
    if ((r = W(e, nm, &pbs, F))) return r;
    snprintf(nm, sizeof nm, "fused%d", j);
    if ((r = get_buf(e, bname(nm), (size_t)round_up(OH * OW, 128) * F * 2, &prj))) return r;
    GemmArgs g = base_args(OH * OW, F, F, EPI_F16);
    g.out_f16 = (__half*)prj;
    g.bias = pbs;
    g.ldc = F;
    if ((r = gemm(e, (const __half*)up, F, pwt, F, g))) return r;
    fused = prj;
    ch = OH;
    cw_ = OW;
  }

  // ---- head ----
  {
    const int F2 = round_up(F / 2, 64);
    const __half *w1, *w2;
    const float *b1h, *b2h, *w3, *b3;
    if ((r = W(e, "h.c1.w", &w1, (size_t)F2 * 9 * F)) || (r = W(e, "h.c1.b", &b1h, F2)) ||
        (r = W(e, "h.c2.w", &w2, (size_t)32 * 9 * F2)) || (r = W(e, "h.c2.b", &b2h, 32)) ||
        (r = W(e, "h.c3.w", &w3, 32)) || (r = W(e, "h.c3.b", &b3, 1)))
      return r;
    void *h1, *h1u;
    if ((r = get_buf(e, bname("h1"), (size_t)ch * cw_ * F2 * 2, &h1)) || (r = get_buf(e, bname("h1u"), (size_t)IH * IW * F2 * 2, &h1u)))
      return r;
    GemmArgs g1 = base_args(0, F2, 0, EPI_F16);
    g1.out_f16 = (__half*)h1;
    g1.bias = b1h;
    g1.ldc = F2;
    if ((r = conv(e, (const __half*)fused, ch, cw_, F, w1, true, g1))) return r;
    launch_upsample_ac((const __half*)h1, ch, cw_, F2, (__half*)h1u, IH, IW, s);
    e->launches++;
    GemmArgs g2 = base_args(0, 32, 0, EPI_HEAD);
    g2.out_f32 = (float*)depth_d;
    g2.bias = b2h;
    g2.w3 = w3;
    g2.b3p = b3;
    if ((r = conv(e, (const __half*)h1u, IH, IW, F2, w2, true, g2, 32))) return r;
  }
  if (fork) DCK(cudaEventRecord(e->ev_join[b], s));
  }  // images
  e->cur = nullptr;
  if (fork)
    for (int b = 0; b < B; ++b) DCK(cudaStreamWaitEvent(s_main, e->ev_join[b], 0));
  e->span_end(sp_tail, s_main);
  DCK(cudaGetLastError());
  return VD3D_OK;
}

}  // namespace

extern "C" {

// pixel_values: f32 [3, image_h, image_w] (already resized + normalised); depth_out f32 [image_h, image_w]
int vd3d_depth_forward(vd3d_depth* e, const float* pixel_values, float* depth_out, int mem) {
  if (!e || !pixel_values || !depth_out) return VD3D_ERR_ARG;
  const int IH = e->cfg.image_h, IW = e->cfg.image_w;
  cudaStream_t s = e->stream;
  int r;
  void *px_d = (void*)pixel_values, *depth_d = depth_out;
  if (mem == VD3D_MEM_HOST) {
    if ((r = get_buf(e, "px", (size_t)3 * IH * IW * 4, &px_d))) return r;
    if ((r = get_buf(e, "depth", (size_t)IH * IW * 4, &depth_d))) return r;
    DCK(cudaMemcpyAsync(px_d, pixel_values, (size_t)3 * IH * IW * 4, cudaMemcpyHostToDevice, s));
  }
  const float* pxs[1] = {(const float*)px_d};
  float* ds[1] = {(float*)depth_d};
  if ((r = forward_core(e, 1, pxs, ds))) return r;
  if (mem == VD3D_MEM_HOST) {
    DCK(cudaMemcpyAsync(depth_out, depth_d, (size_t)IH * IW * 4, cudaMemcpyDeviceToHost, s));
    DCK(cudaStreamSynchronize(s));
  }
  return VD3D_OK;
}

// B frames (BGR u8 [h,w,3], DEVICE) -> DPT processor each -> ONE batched forward -> bicubic back + min-max u8 each.
// depth_u8_dev[b] / depth_f32_dev[b] (either array may be null) receive the results; enqueues on the engine stream.
int vd3d_depth_infer_batch_device(vd3d_depth* e, int B, const uint8_t* const* frames_bgr_dev, int h, int w,
                                  uint8_t* const* depth_u8_dev, float* const* depth_f32_dev, int invert) {
  if (!e || !frames_bgr_dev || B < 1 || B > kMaxBatch || h < 16 || w < 16) return VD3D_ERR_ARG;
  const int IH = e->cfg.image_h, IW = e->cfg.image_w;
  cudaStream_t s = e->stream;
  void *tmp, *rgb, *mm, *up;
  int r;
  char nm[32];
  const float* pxs[kMaxBatch];
  float* dds[kMaxBatch];
  if ((r = get_buf(e, "pp.tmp", (size_t)h * IW * 3, &tmp)) || (r = get_buf(e, "pp.rgb", (size_t)IH * IW * 3, &rgb)) ||
      (r = get_buf(e, "post.up", (size_t)h * w * 4, &up)) || (r = get_buf(e, "post.mm", 64, &mm)))
    return r;
  for (int b = 0; b < B; ++b) {
    void *px, *dd;
    snprintf(nm, sizeof nm, b ? "px.%d" : "px", b);
    if ((r = get_buf(e, nm, (size_t)3 * IH * IW * 4, &px))) return r;
    snprintf(nm, sizeof nm, b ? "depth.%d" : "depth", b);
    if ((r = get_buf(e, nm, (size_t)IH * IW * 4, &dd))) return r;
    launch_preprocess(frames_bgr_dev[b], h, w, (uint8_t*)tmp, (uint8_t*)rgb, (float*)px, IH, IW, s);
    e->launches += 3;
    pxs[b] = (const float*)px;
    dds[b] = (float*)dd;
  }
  if ((r = forward_core(e, B, pxs, dds))) return r;
  for (int b = 0; b < B; ++b) {
    float* upt = (depth_f32_dev && depth_f32_dev[b]) ? depth_f32_dev[b] : (float*)up;
    uint8_t* u8 = depth_u8_dev ? depth_u8_dev[b] : nullptr;
    launch_depth_post(dds[b], IH, IW, upt, h, w, (unsigned*)mm, u8, invert, s);
    e->launches += u8 ? 3 : 2;
  }
  DCK(cudaGetLastError());
  return VD3D_OK;
}

// host-buffer form of the batch (pipe protocol: the reference hands the whole list to the HF pipeline)
int vd3d_depth_infer_batch(vd3d_depth* e, int B, const uint8_t* const* frames_bgr, int h, int w, float* const* depth_f32,
                           uint8_t* const* depth_u8, int invert) {
  if (!e || !frames_bgr || B < 1 || B > kMaxBatch) return VD3D_ERR_ARG;
  int r;
  char nm[32];
  const uint8_t* fd[kMaxBatch];
  uint8_t* u8d[kMaxBatch];
  float* f32d[kMaxBatch];
  for (int b = 0; b < B; ++b) {
    void *a, *c, *d;
    snprintf(nm, sizeof nm, "io.frame.%d", b);
    if ((r = get_buf(e, nm, (size_t)h * w * 3, &a))) return r;
    snprintf(nm, sizeof nm, "io.u8.%d", b);
    if ((r = get_buf(e, nm, (size_t)h * w, &c))) return r;
    snprintf(nm, sizeof nm, "io.f32.%d", b);
    if ((r = get_buf(e, nm, (size_t)h * w * 4, &d))) return r;
    DCK(cudaMemcpyAsync(a, frames_bgr[b], (size_t)h * w * 3, cudaMemcpyHostToDevice, e->stream));
    fd[b] = (const uint8_t*)a;
    u8d[b] = (uint8_t*)c;
    f32d[b] = (float*)d;
  }
  if ((r = vd3d_depth_infer_batch_device(e, B, fd, h, w, u8d, f32d, invert))) return r;
  for (int b = 0; b < B; ++b) {
    if (depth_f32 && depth_f32[b])
      DCK(cudaMemcpyAsync(depth_f32[b], f32d[b], (size_t)h * w * 4, cudaMemcpyDeviceToHost, e->stream));
    if (depth_u8 && depth_u8[b])
      DCK(cudaMemcpyAsync(depth_u8[b], u8d[b], (size_t)h * w, cudaMemcpyDeviceToHost, e->stream));
  }
  DCK(cudaStreamSynchronize(e->stream));
  return VD3D_OK;
}

// ---------------------------------------------------------------------------
// Real-ESRGAN upscale stage (core/merged_pipeline.py:240-267; the network is the ONNX export of xinntao/Real-ESRGAN's
// SRVGGNetCompact: conv3x3(3->64)+PReLU, num_conv x [conv3x3(64->64)+PReLU], conv3x3(64->48), PixelShuffle(4), + nearest
// x4 of the input).  Every conv is the tcgen05 implicit GEMM of the depth neck (f16 NHWC activations, fp32 accumulate);
// the last one keeps fp32.  The engine object is the depth engine's container (named weights + buffers).
int vd3d_sr_create(void* stream, vd3d_depth** out) {
  if (!out) return VD3D_ERR_ARG;
  *out = nullptr;
  if (!load_encode()) return VD3D_ERR_CUDA;
  vd3d_depth* e = new vd3d_depth();
  memset(&e->cfg, 0, sizeof e->cfg);
  e->stream = (cudaStream_t)stream;
  *out = e;
  return VD3D_OK;
}

// frame BGR u8 [h,w,3] -> BGR u8 [4h,4w,3] (preprocess_esr -> network -> postprocess_esr).  Weights: "sr.c{i}.w" f16
// [Cout, 9*64] (tap-major, input channels padded to 64), "sr.c{i}.b" f32 [Cout], "sr.a{i}" f32 [64] (PReLU slopes),
// i = 0 .. num_conv+1.
int vd3d_sr_forward(vd3d_depth* e, const uint8_t* frame_bgr, int h, int w, int num_conv, uint8_t* out_bgr, int mem) {
  if (!e || !frame_bgr || !out_bgr || h < 8 || w < 8 || num_conv < 1 || num_conv > 64) return VD3D_ERR_ARG;
  cudaStream_t s = e->stream;
  const size_t npix = (size_t)h * w;
  void *fd = (void*)frame_bgr, *od = out_bgr, *x0, *x1, *cvo;
  int r;
  char nm[32];
  if (mem == VD3D_MEM_HOST) {
    if ((r = get_buf(e, "sr.in", npix * 3, &fd)) || (r = get_buf(e, "sr.outu8", npix * 48, &od))) return r;
    DCK(cudaMemcpyAsync(fd, frame_bgr, npix * 3, cudaMemcpyHostToDevice, s));
  }
  if ((r = get_buf(e, "sr.x0", npix * 64 * 2, &x0)) || (r = get_buf(e, "sr.x1", npix * 64 * 2, &x1)) ||
      (r = get_buf(e, "sr.cv", npix * 48 * 4, &cvo)))
    return r;
  launch_sr_in((const uint8_t*)fd, (__half*)x0, (int)npix, s);
  e->launches++;
  __half *cur = (__half*)x0, *nxt = (__half*)x1;
  for (int i = 0; i <= num_conv + 1; ++i) {
    const bool last = (i == num_conv + 1);
    const int cout = last ? 48 : 64;
    const __half* wt;
    const float *bias, *slope = nullptr;
    snprintf(nm, sizeof nm, "sr.c%d.w", i);
    if ((r = W(e, nm, &wt, (size_t)cout * 9 * 64))) return r;
    snprintf(nm, sizeof nm, "sr.c%d.b", i);
    if ((r = W(e, nm, &bias, cout))) return r;
    if (!last) {
      snprintf(nm, sizeof nm, "sr.a%d", i);
      if ((r = W(e, nm, &slope, 64))) return r;
    }
    GemmArgs g = base_args(0, cout, 0, last ? EPI_F32 : EPI_F16);
    g.bias = bias;
    if (last) {
      g.out_f32 = (float*)cvo;
      g.ldc = 48;
    } else {
      g.out_f16 = nxt;
      g.ldc = 64;
      g.act = 3;
      g.ls = slope;
    }
    if ((r = conv(e, cur, h, w, 64, wt, true, g, last ? 32 : 64))) return r;
    if (!last) {
      __half* t = cur;
      cur = nxt;
      nxt = t;
    }
  }
  launch_sr_out((const float*)cvo, 48, (const uint8_t*)fd, (uint8_t*)od, h, w, s);
  e->launches++;
  DCK(cudaGetLastError());
  if (mem == VD3D_MEM_HOST) {
    DCK(cudaMemcpyAsync(out_bgr, od, npix * 48, cudaMemcpyDeviceToHost, s));
    DCK(cudaStreamSynchronize(s));
  }
  return VD3D_OK;
}

// frame (BGR u8 [h,w,3], DEVICE) -> processor -> forward -> bicubic back to (h,w) -> min-max u8 [h,w]
// (hf pipeline + convert_depth_to_grayscale, core/render_depth.py:1113-1119,605-611,1907-1917).
// Enqueues on the engine stream without synchronising; depth_f32_or_null receives the resized
// float depth ("predicted_depth" of the pipe protocol).
int vd3d_depth_infer_device(vd3d_depth* e, const uint8_t* frame_bgr_dev, int h, int w, uint8_t* depth_u8_dev,
                            float* depth_f32_or_null, int invert) {
  if (!e || !frame_bgr_dev || h < 16 || w < 16) return VD3D_ERR_ARG;
  const int IH = e->cfg.image_h, IW = e->cfg.image_w;
  cudaStream_t s = e->stream;
  void *tmp, *rgb, *px, *up, *mm, *dd;
  int r;
  if ((r = get_buf(e, "pp.tmp", (size_t)h * IW * 3, &tmp)) || (r = get_buf(e, "pp.rgb", (size_t)IH * IW * 3, &rgb)) ||
      (r = get_buf(e, "px", (size_t)3 * IH * IW * 4, &px)) || (r = get_buf(e, "depth", (size_t)IH * IW * 4, &dd)) ||
      (r = get_buf(e, "post.up", (size_t)h * w * 4, &up)) || (r = get_buf(e, "post.mm", 64, &mm)))
    return r;
  launch_preprocess(frame_bgr_dev, h, w, (uint8_t*)tmp, (uint8_t*)rgb, (float*)px, IH, IW, s);
  e->launches += 3;
  if ((r = vd3d_depth_forward(e, (const float*)px, (float*)dd, VD3D_MEM_DEVICE))) return r;
  float* upt = depth_f32_or_null ? depth_f32_or_null : (float*)up;
  launch_depth_post((const float*)dd, IH, IW, upt, h, w, (unsigned*)mm, depth_u8_dev, invert, s);
  e->launches += depth_u8_dev ? 3 : 2;
  DCK(cudaGetLastError());
  return VD3D_OK;
}

// host-buffer form (pipe protocol / tests): frame BGR u8 [h,w,3] -> depth f32 [h,w] and/or u8 [h,w]
int vd3d_depth_infer(vd3d_depth* e, const uint8_t* frame_bgr, int h, int w, float* depth_f32, uint8_t* depth_u8,
                     int invert) {
  if (!e || !frame_bgr) return VD3D_ERR_ARG;
  void *fd, *u8d, *f32d;
  int r;
  if ((r = get_buf(e, "io.frame", (size_t)h * w * 3, &fd)) || (r = get_buf(e, "io.u8", (size_t)h * w, &u8d)) ||
      (r = get_buf(e, "io.f32", (size_t)h * w * 4, &f32d)))
    return r;
  DCK(cudaMemcpyAsync(fd, frame_bgr, (size_t)h * w * 3, cudaMemcpyHostToDevice, e->stream));
  if ((r = vd3d_depth_infer_device(e, (const uint8_t*)fd, h, w, (uint8_t*)u8d, (float*)f32d, invert))) return r;
  if (depth_f32) DCK(cudaMemcpyAsync(depth_f32, f32d, (size_t)h * w * 4, cudaMemcpyDeviceToHost, e->stream));
  if (depth_u8) DCK(cudaMemcpyAsync(depth_u8, u8d, (size_t)h * w, cudaMemcpyDeviceToHost, e->stream));
  DCK(cudaStreamSynchronize(e->stream));
  return VD3D_OK;
}

}  // extern "C"
