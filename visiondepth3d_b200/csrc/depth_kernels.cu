// depth_kernels.cu -- non-GEMM kernels of the Depth-Anything-V2 forward and the
// instantiations / launchers of the tcgen05 GEMM (umma_gemm.cuh).
// Reference: transformers 5.5.0 models/dinov2/modeling_dinov2.py and
// models/depth_anything/modeling_depth_anything.py as called by
// core/render_depth.py:1106-1119 (hf_batch_safe_pipe).
#include "depth_launch.h"
#include "umma_gemm.cuh"

namespace vd3d {

// ---------------------------------------------------------------------------
// LayerNorm (eps 1e-6) over the fp32 residual stream -> f16 GEMM operand.  One warp per row.
// row_off / out_off: skip the CLS row when producing the backbone feature maps.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ x, int rows, int D,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   __half* __restrict__ out, int row_off, float eps) {
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)(row + row_off) * D;
  float v[32];  // D <= 1024
  float s = 0.f;
  int n = D / 32;
  for (int i = 0; i < n; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i];
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / (float)D;
  float q = 0.f;
  for (int i = 0; i < n; ++i) {
    float d = v[i] - mean;
    q += d * d;
  }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  float rstd = rsqrtf(q / (float)D + eps);
  __half* orow = out + (size_t)row * D;
  for (int i = 0; i < n; ++i) {
    int c = lane + 32 * i;
    orow[c] = __float2half_rn((v[i] - mean) * rstd * gamma[c] + beta[c]);
  }
}

// softmax over keys (scores already scaled: q was multiplied by 1/sqrt(d)); one warp per row
__global__ void __launch_bounds__(256) k_softmax(const float* __restrict__ S, __half* __restrict__ P, int rows_per_head,
                                                 int heads, int ncols, int ld) {
  int gr = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (gr >= rows_per_head * heads) return;
  int h = gr / rows_per_head, r = gr % rows_per_head;
  const float* s = S + ((size_t)h * ld + r) * ld;
  __half* p = P + ((size_t)h * ld + r) * ld;
  float v[96];  // ncols <= 3072
  int n = (ncols + 31) / 32;
  float mx = -INFINITY;
  for (int i = 0; i < n; ++i) {
    int c = lane + 32 * i;
    v[i] = (c < ncols) ? s[c] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int i = 0; i < n; ++i) {
    v[i] = (lane + 32 * i < ncols) ? expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float inv = 1.0f / sum;
  for (int i = 0; i < n; ++i) {
    int c = lane + 32 * i;
    if (c < ncols) p[c] = __float2half_rn(v[i] * inv);
  }
}

// patch embedding im2col: pixel_values f32 [3, IH, IW] -> A f16 [ph*pw, kpad], k = c*196 + dy*14 + dx
__global__ void __launch_bounds__(256) k_patch_im2col(const float* __restrict__ px, int IH, int IW, int ph, int pw,
                                                      __half* __restrict__ A, int kpad) {
  int idx = blockIdx.x * 256 + threadIdx.x;
  int total = ph * pw * 588;
  if (idx >= total) return;
  int k = idx % 588, t = idx / 588;
  int c = k / 196, rr = k % 196, dy = rr / 14, dx = rr % 14;
  int py = t / pw, pxx = t % pw;
  float v = px[((size_t)c * IH + (py * 14 + dy)) * IW + (pxx * 14 + dx)];
  A[(size_t)t * kpad + k] = __float2half_rn(v);
}

// x[0, :] = cls + pos[0, :]
__global__ void k_set_cls(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int D) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i < D) x[i] = cls[i] + pos[i];
}

// 3x3 stride-2 pad-1 im2col on NHWC f16 (reassemble factor 0.5): out [(OH*OW), 9*C]
__global__ void __launch_bounds__(256) k_im2col_s2(const __half* __restrict__ in, int H, int W, int C, int ldc,
                                                   __half* __restrict__ out, int OH, int OW) {
  size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  size_t total = (size_t)OH * OW * 9 * C;
  if (idx >= total) return;
  int c = idx % C;
  size_t r = idx / C;
  int tap = r % 9;
  size_t o = r / 9;
  int ox = o % OW, oy = o / OW;
  int iy = oy * 2 - 1 + tap / 3, ix = ox * 2 - 1 + tap % 3;
  __half v = __float2half_rn(0.f);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = in[((size_t)iy * W + ix) * ldc + c];
  out[idx] = v;
}

// bilinear upsample, align_corners=True, NHWC f16 (fusion stages and the head)
__global__ void __launch_bounds__(256) k_upsample_ac(const __half* __restrict__ in, int H, int W, int C,
                                                     __half* __restrict__ out, int OH, int OW) {
  size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;  // one thread per (pixel, 8 channels)
  int c8 = C / 8;
  size_t total = (size_t)OH * OW * c8;
  if (idx >= total) return;
  int cg = idx % c8;
  size_t o = idx / c8;
  int ox = o % OW, oy = o / OW;
  float sy = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  float sx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  float fy = sy * oy, fx = sx * ox;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  float ly = fy - y0, lx = fx - x0;
  const uint4* p00 = (const uint4*)(in + ((size_t)y0 * W + x0) * C) + cg;
  const uint4* p01 = (const uint4*)(in + ((size_t)y0 * W + x1) * C) + cg;
  const uint4* p10 = (const uint4*)(in + ((size_t)y1 * W + x0) * C) + cg;
  const uint4* p11 = (const uint4*)(in + ((size_t)y1 * W + x1) * C) + cg;
  uint4 a = *p00, b = *p01, c = *p10, d = *p11, r;
  const __half2* ha = (const __half2*)&a;
  const __half2* hb = (const __half2*)&b;
  const __half2* hc = (const __half2*)&c;
  const __half2* hd = (const __half2*)&d;
  __half2* hr = (__half2*)&r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 fa = __half22float2(ha[j]), fb = __half22float2(hb[j]), fc = __half22float2(hc[j]),
           fd = __half22float2(hd[j]);
    float t0x = fa.x + lx * (fb.x - fa.x), t0y = fa.y + lx * (fb.y - fa.y);
    float t1x = fc.x + lx * (fd.x - fc.x), t1y = fc.y + lx * (fd.y - fc.y);
    hr[j] = __floats2half2_rn(t0x + ly * (t1x - t0x), t0y + ly * (t1y - t0y));
  }
  ((uint4*)(out + o * C))[cg] = r;
}

// elementwise relu copy (f16)
__global__ void __launch_bounds__(256) k_relu_f16(const __half* __restrict__ in, __half* __restrict__ out, size_t n8) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n8) return;
  uint4 a = ((const uint4*)in)[i];
  __half2* h = (__half2*)&a;
  const __half2 z = __floats2half2_rn(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __hmax2(h[j], z);
  ((uint4*)out)[i] = a;
}

// sum = a + b ; sum_relu = relu(sum)   (fusion: hidden_state + residual_layer1(residual))
__global__ void __launch_bounds__(256) k_add_relu_f16(const __half* __restrict__ a, const __half* __restrict__ b,
                                                      __half* __restrict__ sum, __half* __restrict__ sum_relu,
                                                      size_t n8) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n8) return;
  uint4 ua = ((const uint4*)a)[i], ub = ((const uint4*)b)[i], us, ur;
  const __half2* ha = (const __half2*)&ua;
  const __half2* hb = (const __half2*)&ub;
  __half2* hs = (__half2*)&us;
  __half2* hr = (__half2*)&ur;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 fa = __half22float2(ha[j]), fb = __half22float2(hb[j]);
    float sx = fa.x + fb.x, sy = fa.y + fb.y;
    hs[j] = __floats2half2_rn(sx, sy);
    hr[j] = __floats2half2_rn(fmaxf(sx, 0.f), fmaxf(sy, 0.f));
  }
  ((uint4*)sum)[i] = us;
  ((uint4*)sum_relu)[i] = ur;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
void launch_add_relu_f16(const __half* a, const __half* b, __half* sum, __half* sum_relu, size_t n, cudaStream_t s) {
  size_t n8 = n / 8;
  k_add_relu_f16<<<(unsigned)((n8 + 255) / 256), 256, 0, s>>>(a, b, sum, sum_relu, n8);
}
void launch_layernorm(const float* x, int rows, int D, const float* g, const float* b, __half* out, int row_off,
                      cudaStream_t s) {
  k_layernorm<<<(rows + 7) / 8, 256, 0, s>>>(x, rows, D, g, b, out, row_off, 1e-6f);
}
void launch_softmax(const float* S, __half* P, int rows, int heads, int ncols, int ld, cudaStream_t s) {
  k_softmax<<<(rows * heads + 7) / 8, 256, 0, s>>>(S, P, rows, heads, ncols, ld);
}
void launch_patch_im2col(const float* px, int IH, int IW, int ph, int pw, __half* A, int kpad, cudaStream_t s) {
  int total = ph * pw * 588;
  k_patch_im2col<<<(total + 255) / 256, 256, 0, s>>>(px, IH, IW, ph, pw, A, kpad);
}
void launch_set_cls(float* x, const float* cls, const float* pos, int D, cudaStream_t s) {
  k_set_cls<<<(D + 255) / 256, 256, 0, s>>>(x, cls, pos, D);
}
void launch_im2col_s2(const __half* in, int H, int W, int C, int ldc, __half* out, int OH, int OW, cudaStream_t s) {
  size_t total = (size_t)OH * OW * 9 * C;
  k_im2col_s2<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, H, W, C, ldc, out, OH, OW);
}
void launch_upsample_ac(const __half* in, int H, int W, int C, __half* out, int OH, int OW, cudaStream_t s) {
  size_t total = (size_t)OH * OW * (C / 8);
  k_upsample_ac<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, H, W, C, out, OH, OW);
}
void launch_relu_f16(const __half* in, __half* out, size_t n, cudaStream_t s) {
  size_t n8 = n / 8;
  k_relu_f16<<<(unsigned)((n8 + 255) / 256), 256, 0, s>>>(in, out, n8);
}

template <int BN, int STAGES>
static cudaError_t launch_gemm_t(const CUtensorMap& a, const CUtensorMap& b, const GemmArgs& g, dim3 grid,
                                 cudaStream_t s) {
  static bool attr_set = false;
  constexpr int smem = GemmSmem<BN, STAGES>::kTotal;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_umma_gemm<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_umma_gemm<BN, STAGES><<<grid, kGemmThreads, smem, s>>>(a, b, g);
  return cudaGetLastError();
}

cudaError_t launch_gemm(int bn, const CUtensorMap& a, const CUtensorMap& b, const GemmArgs& g, int m_tiles,
                        int batch, cudaStream_t s) {
  dim3 grid((g.N + bn - 1) / bn, m_tiles, batch);
  switch (bn) {
    case 128: return launch_gemm_t<128, 3>(a, b, g, grid, s);
    case 64: return launch_gemm_t<64, 4>(a, b, g, grid, s);
    case 32: return launch_gemm_t<32, 4>(a, b, g, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace vd3d
