// umma_gemm.cuh -- hand-written sm_100a GEMM on the 5th-gen tensor cores.
//
//   D[M,N] (fp32 in TMEM) = A[M,K] (f16, K-major) x B[N,K]^T (f16, K-major)
//
// * operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) into a STAGES-deep
//   shared-memory ring guarded by mbarriers;
// * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//   accumulating in tensor memory; tcgen05.commit releases ring slots / signals the epilogue;
// * four epilogue warps read the accumulator with tcgen05.ld (32 lanes x 32 columns per
//   warp) and apply the fused epilogue (bias, GELU, LayerScale+residual, QKV head split with
//   V transposed, pixel-shuffle for ConvTranspose, ReLU / residual for convs, DPT head).
// * A can also be an NHWC activation read through a 3-D tensor map (C, W, H): the K loop
//   then walks the 3x3 taps and channel blocks (implicit GEMM); out-of-image taps are
//   zero-filled by TMA, so no im2col buffer and no padding copies exist.
//
// Used for the Depth-Anything-V2 forward (the dense contraction of the hot path):
// transformers' DepthAnythingForDepthEstimation as called from core/render_depth.py:1106-1119.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vd3d {

enum Epi : int {
  EPI_F16 = 0,        // out_f16[m, n] = act(acc + bias) (+ res_f16), optional second relu copy
  EPI_F32 = 1,        // out_f32[m, n] = acc (+bias)                       (attention scores)
  EPI_RESID_LS = 2,   // x_f32[m, n] += ls[n] * (acc + bias[n])            (proj / fc2)
  EPI_QKV = 3,        // split heads: q (x scale), k -> [h][m][64]; v -> vT [h][64][m]
  EPI_PATCH = 4,      // x_f32[m+1, n] = acc + bias[n] + pos[m+1, n]       (patch embedding)
  EPI_CONVT = 5,      // ConvTranspose2d(k=s): pixel-shuffle scatter into NHWC f16
  EPI_HEAD = 6,       // depth[m] = relu(sum_n relu(acc+bias)[n] * w3[n] + b3)
};

struct GemmArgs {
  int M, N, K;            // logical sizes (K = total reduction length)
  int nt, mt, nz;         // tile counts along N, M and batch (filled by launch_gemm)
  int epi;
  int act;                // 0 none, 1 GELU(erf), 2 ReLU, 3 PReLU (slopes in ls)
  // conv mode (implicit GEMM over a (C, W, H) activation map)
  int conv;               // 0 = plain GEMM, 1 = 3x3 pad 1, 2 = 1x1 over the same map
  int cin;                // channels per tap (multiple of 64)
  int imgW, imgH;         // output == input spatial size (stride 1)
  int tw, th;             // pixel tile: tw * th == 128
  // outputs / epilogue operands
  __half* out_f16;
  __half* out2_f16;       // optional relu(out) copy (pre-activation consumers)
  const __half* res_f16;  // optional residual added before the activation
  float* out_f32;
  const float* bias;      // [N] or null
  const float* ls;        // LayerScale lambda [N]
  const float* pos;       // position embeddings [(M+1), N]
  int ldc;                // leading dimension of out (elements)
  long long out_batch_stride;  // per blockIdx.z
  // EPI_QKV
  __half* q;
  __half* k;
  __half* vt;
  int heads, npad, dmodel;
  float qscale;
  // EPI_CONVT
  int ct_k, ct_cout, ct_w;  // kernel(=stride), Cout, input width (tokens per row)
  // EPI_HEAD
  const float* w3;
  const float* b3p;
  // tuning hook (vd3d_gemm_bench), low 3 bits: 0 = normal, 1 = skip the MMAs (operand-feed rate), 2 = skip the TMA
  // loads (MMA rate), 3 = prologue + teardown only, 4 = epilogue only hands the accumulator back, 5 = epilogue
  // reads TMEM but stores nothing; bit 3 (8): poll barriers with test_wait instead of try_wait
  int dbg;
};

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// non-suspending variant (tuning hook): polls with test_wait instead of the potentially-suspending try_wait
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_dbg(uint32_t bar, uint32_t parity, int spin) {
  if (spin)
    mbar_wait_spin(bar, parity);
  else
    mbar_wait(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}

// K-major, 128B-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(=1, unused for swizzled K-major)<<16 | SBO(1024B>>4)<<32 | version 1<<46 | SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=f16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) /*c_format f32*/ | (0u << 7) /*a f16*/ | (0u << 10) /*b f16*/ | (0u << 15) | (0u << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 fp32 columns -> 32 registers per thread (thread = lane = accumulator row)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 64 fp32 columns -> 64 registers per thread
__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// registers -> TMEM, 32 lanes x 32 fp32 columns (used to rescale the attention output accumulator)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 16-column variants (keep register pressure low where a whole 32/64-column chunk is already live)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x) taken from the low mantissa bits of x + 1.5 * 2^23,
// degree-4 minimax of 2^f on [-0.5, 0.5] (relative error 2.7e-6 in fp32), n added to the exponent field.  Valid for
// -125 <= x < 2^21 (clamped below).  The attention softmax runs a third of its exponentials here: ex2.approx issues at
// a quarter of a warp per SM sub-partition clock on B200 and was the kernel's bound (ncu: XU pipe 71 % of elapsed).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.009570102207362652f, 0.05591785907745361f);
  p = fmaf(f, p, 0.240247443318367f);
  p = fmaf(f, p, 0.6931217908859253f);
  p = fmaf(f, p, 0.9999992847442627f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), exact-erf form as in HF's "gelu".  erfc(z) for z >= 0 from Abramowitz-Stegun
// 7.1.26 (|error| <= 1.5e-7): erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1 / (1 + p z), and
// GELU(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2)  (no cancellation for x < 0).  ~14 FP32 ops + 2 MUFU instead of
// the ~45 of erff(): the fc1 epilogue was ALU bound on erff (tools/gemm_sweep.py).  Output is stored as f16.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float e = ex2_fast(x * x * -0.72134752044448170368f);  // exp(-x^2 / 2)
  return fmaf(-0.5f * ax * p, e, fmaxf(x, 0.f));
}

// 256-bit global accesses (sm_100: LDG/STG.E.256): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void ldg_v8(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void ldg_nc_v8(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 16 fp32 -> 16 f16 (32 bytes) starting at a[base]
__device__ __forceinline__ void pack16(const float (&a)[32], int base, bool relu, uint32_t (&u)[8]) {
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float x = a[base + 2 * t], y = a[base + 2 * t + 1];
    if (relu) {
      x = fmaxf(x, 0.f);
      y = fmaxf(y, 0.f);
    }
    __half2 h = __floats2half2_rn(x, y);
    u[t] = *(uint32_t*)&h;
  }
}

}  // namespace umma

// proj / fc2 epilogue (EPI_RESID_LS) split in two so that the read of the fp32 residual stream -- which does not depend on
// the MMAs -- can be issued BEFORE the accumulator is complete: issued after tmem_full, its ~1.5 us of L2 latency per
// chunk was exposed twice per tile (proj, K = 768, spends 1.6 us per tile in the MMAs).
__device__ __forceinline__ bool resid_split_ok(const GemmArgs& g) {
  return g.epi == EPI_RESID_LS && (g.N & 31) == 0 && (g.ldc & 7) == 0 && g.bias != nullptr && g.conv == 0;
}
__device__ __forceinline__ void resid_load(const GemmArgs& g, const int m, const int n0, uint32_t (&x)[32]) {
  const float* src = g.out_f32 + (size_t)m * g.ldc + n0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(x[8 * j]), "=r"(x[8 * j + 1]), "=r"(x[8 * j + 2]), "=r"(x[8 * j + 3]), "=r"(x[8 * j + 4]),
                   "=r"(x[8 * j + 5]), "=r"(x[8 * j + 6]), "=r"(x[8 * j + 7])
                 : "l"(src + 8 * j));
}
// x[m, n0 .. n0+31] = x + ls * (acc + bias), x already in registers
__device__ __forceinline__ void resid_apply_store(const GemmArgs& g, const uint32_t (&v)[32], const int m, const int n0,
                                                  uint32_t (&x)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b4 = __ldg((const float4*)(g.bias + n0) + j);
    const float4 l4 = __ldg((const float4*)(g.ls + n0) + j);
    x[4 * j + 0] = __float_as_uint(fmaf(l4.x, __uint_as_float(v[4 * j + 0]) + b4.x, __uint_as_float(x[4 * j + 0])));
    x[4 * j + 1] = __float_as_uint(fmaf(l4.y, __uint_as_float(v[4 * j + 1]) + b4.y, __uint_as_float(x[4 * j + 1])));
    x[4 * j + 2] = __float_as_uint(fmaf(l4.z, __uint_as_float(v[4 * j + 2]) + b4.z, __uint_as_float(x[4 * j + 2])));
    x[4 * j + 3] = __float_as_uint(fmaf(l4.w, __uint_as_float(v[4 * j + 3]) + b4.w, __uint_as_float(x[4 * j + 3])));
  }
  float* dst = g.out_f32 + (size_t)m * g.ldc + n0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * j), "r"(x[8 * j]),
                 "r"(x[8 * j + 1]), "r"(x[8 * j + 2]), "r"(x[8 * j + 3]), "r"(x[8 * j + 4]), "r"(x[8 * j + 5]),
                 "r"(x[8 * j + 6]), "r"(x[8 * j + 7])
                 : "memory");
}

// Fused epilogue of one 32-column chunk of one accumulator row (thread == row): v = raw fp32 accumulator bits,
// m = logical output row, n0 = first output column.  Shared by the 1-CTA and the CTA-pair kernels; every
// register array is indexed with compile-time constants only so the chunk stays in registers.
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmArgs& g, const uint32_t (&v)[32], const int m, const int z,
                                                    const int n0, const bool row_ok, float& head_acc) {
  if (!row_ok || n0 >= g.N) return;
  const int nvalid = min(32, g.N - n0);
  const bool full = (nvalid == 32);
  float a[32];
  if (g.bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b4 = __ldg((const float4*)(g.bias + n0) + j);
        a[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b4.x;
        a[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
        a[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
        a[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]) + ((j < nvalid) ? g.bias[n0 + j] : 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]);
  }
  switch (g.epi) {
    case EPI_F16: {
      const size_t o = (size_t)z * g.out_batch_stride + (size_t)m * g.ldc + n0;
      const bool vec = full && ((o & 15) == 0);  // 32-byte aligned: 256-bit accesses
      if (g.res_f16) {
        if (vec) {
          uint32_t rv[2][8];
          umma::ldg_nc_v8(g.res_f16 + o, rv[0]);
          umma::ldg_nc_v8(g.res_f16 + o + 16, rv[1]);
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              float2 f = __half22float2(*(const __half2*)&rv[j][t]);
              a[16 * j + 2 * t] += f.x;
              a[16 * j + 2 * t + 1] += f.y;
            }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) a[j] += __half2float(g.res_f16[o + j]);
        }
      }
      if (g.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = umma::gelu_erf(a[j]);
      } else if (g.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = fmaxf(a[j], 0.f);
      } else if (g.act == 3) {  // PReLU, per-channel slopes in g.ls (SRVGGNetCompact)
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) a[j] = a[j] > 0.f ? a[j] : a[j] * __ldg(g.ls + n0 + j);
      }
      if (vec) {
        uint32_t u[8];
        umma::pack16(a, 0, false, u);
        umma::stg_v8(g.out_f16 + o, u);
        umma::pack16(a, 16, false, u);
        umma::stg_v8(g.out_f16 + o + 16, u);
        if (g.out2_f16) {
          umma::pack16(a, 0, true, u);
          umma::stg_v8(g.out2_f16 + o, u);
          umma::pack16(a, 16, true, u);
          umma::stg_v8(g.out2_f16 + o + 16, u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) {
            g.out_f16[o + j] = __float2half_rn(a[j]);
            if (g.out2_f16) g.out2_f16[o + j] = __float2half_rn(fmaxf(a[j], 0.f));
          }
      }
    } break;
    case EPI_F32: {
      float* dst = g.out_f32 + (size_t)z * g.out_batch_stride + (size_t)m * g.ldc + n0;
      if (full && ((((size_t)m * g.ldc + n0) & 3) == 0) && ((g.out_batch_stride & 3) == 0)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ((float4*)dst)[j] = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = a[j];
      }
    } break;
    case EPI_RESID_LS: {
      // read-modify-write of the fp32 residual stream: issue all loads before the first store
      // (a load/store-per-element loop serialises on possible aliasing: ~800 cycles per element);
      // 256-bit accesses: every instruction moves whole 32-byte sectors
      float* dst = g.out_f32 + (size_t)m * g.ldc + n0;
      if (full && ((((size_t)m * g.ldc + n0) & 7) == 0)) {
        uint32_t x8[4][8];
        float4 l4[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) umma::ldg_v8(dst + 8 * j, x8[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) l4[j] = __ldg((const float4*)(g.ls + n0) + j);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const float4 l = l4[2 * j + t];
            x8[j][4 * t + 0] = __float_as_uint(fmaf(l.x, a[8 * j + 4 * t + 0], __uint_as_float(x8[j][4 * t + 0])));
            x8[j][4 * t + 1] = __float_as_uint(fmaf(l.y, a[8 * j + 4 * t + 1], __uint_as_float(x8[j][4 * t + 1])));
            x8[j][4 * t + 2] = __float_as_uint(fmaf(l.z, a[8 * j + 4 * t + 2], __uint_as_float(x8[j][4 * t + 2])));
            x8[j][4 * t + 3] = __float_as_uint(fmaf(l.w, a[8 * j + 4 * t + 3], __uint_as_float(x8[j][4 * t + 3])));
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) umma::stg_v8(dst + 8 * j, x8[j]);
      } else {
        float xv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) xv[j] = (j < nvalid) ? dst[j] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = xv[j] + g.ls[n0 + j] * a[j];
      }
    } break;
    case EPI_QKV: {
      // n0 is a multiple of 32 and head_dim is 64: the 32 columns lie in one (which, head)
      const int which = n0 / g.dmodel;
      const int rem = n0 - which * g.dmodel;
      const int h = rem >> 6, d0 = rem & 63;
      // batched forward: row m = image * npad + token; q, k are [image][head][npad][64], vT [image][head][64][npad]
      const int img = m / g.npad, tok = m - img * g.npad;
      const size_t zh = (size_t)img * g.heads + h;
      if (which < 2) {
        __half* dst = (which == 0 ? g.q : g.k) + (zh * g.npad + tok) * 64 + d0;
        const float sc = which == 0 ? g.qscale : 1.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] *= sc;
        uint32_t u[8];
        umma::pack16(a, 0, false, u);
        umma::stg_v8(dst, u);
        umma::pack16(a, 16, false, u);
        umma::stg_v8(dst + 16, u);
      } else {
        __half* dst = g.vt + (zh * 64 + d0) * g.npad + tok;  // transposed: lanes -> consecutive tokens
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[(size_t)j * g.npad] = __float2half_rn(a[j]);
      }
    } break;
    case EPI_PATCH: {
      float* dst = g.out_f32 + (size_t)(m + 1) * g.ldc + n0;
      const float* pe = g.pos + (size_t)(m + 1) * g.ldc + n0;
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pv[j] = (j < nvalid) ? __ldg(pe + j) : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) dst[j] = a[j] + pv[j];
    } break;
    case EPI_CONVT: {
      // n = (dy*k + dx)*Cout + co ; m = y*ct_w + x ; out NHWC [(H*k), (W*k), Cout]
      const int tap = n0 / g.ct_cout, co = n0 % g.ct_cout;
      const int dy = tap / g.ct_k, dx = tap % g.ct_k;
      const int y = m / g.ct_w, x = m % g.ct_w;
      size_t o = ((size_t)(y * g.ct_k + dy) * (g.ct_w * g.ct_k) + (x * g.ct_k + dx)) * g.ct_cout + co;
      if (full && ((o & 15) == 0)) {
        uint32_t u[8];
        umma::pack16(a, 0, false, u);
        umma::stg_v8(g.out_f16 + o, u);
        umma::pack16(a, 16, false, u);
        umma::stg_v8(g.out_f16 + o + 16, u);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) g.out_f16[o + j] = __float2half_rn(a[j]);
      }
    } break;
    case EPI_HEAD: {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) head_acc += fmaxf(a[j], 0.f) * __ldg(g.w3 + n0 + j);
    } break;
  }
}

constexpr int kGemmThreads = 320;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue (2 per lane quarter)
constexpr int kBK = 64;            // 64 f16 = 128 B = one swizzle row

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = 128 * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kTotal = STAGES * kStage + 1024 /*align*/ + 256 /*barriers*/;
};

// EW: epilogue warps (8: two per TMEM lane quarter, two CTAs per SM; 16: four per quarter for the one-CTA-per-SM launches
// of the residual GEMMs, where every warp owns ONE chunk of a tile and prefetches its slice of the residual stream)
template <int BN, int STAGES, int EW = 8>
__global__ void __launch_bounds__(64 + 32 * EW, EW == 8 ? 0 : 1)
k_umma_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * S::kStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2] accumulator stage complete (MMA -> epilogue)
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2] accumulator stage drained (epilogue -> MMA)
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // persistent: this CTA walks tiles blockIdx.x, +gridDim.x, ...; n fastest so co-resident CTAs share A in L2
  const int total_tiles = g.nt * g.mt * g.nz;
  const int nkb = (g.K + kBK - 1) / kBK;
  const int dmode = g.dbg & 7, spin = g.dbg & 8;

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA);
    umma::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&full[s]), 1);
      umma::mbar_init(umma::smem_u32(&empty[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(umma::smem_u32(&tmem_full[i]), 1);
      umma::mbar_init(umma::smem_u32(&tmem_empty[i]), EW);
    }
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(umma::smem_u32(tmem_slot), 2 * BN);  // two accumulator stages
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& n_blk, int& m_blk, int& z, int& px0, int& py0) {
    n_blk = tile % g.nt;
    int rem = tile / g.nt;
    m_blk = rem % g.mt;
    z = rem / g.mt;
    px0 = py0 = 0;
    if (g.conv) {  // pixel tile origin for conv mode
      int tiles_x = (g.imgW + g.tw - 1) / g.tw;
      py0 = (m_blk / tiles_x) * g.th;
      px0 = (m_blk % tiles_x) * g.tw;
    }
  };

  if (dmode == 3) {
    // tuning: prologue + teardown only
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int kit = 0;  // k-block counter across tiles (ring position)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int n_blk, m_blk, z, px0, py0;
      decode(tile, n_blk, m_blk, z, px0, py0);
      for (int kb = 0; kb < nkb; ++kb, ++kit) {
        const int s = kit % STAGES;
        const uint32_t ph = (kit / STAGES) & 1;
        umma::mbar_wait_dbg(umma::smem_u32(&empty[s]), ph ^ 1, spin);
        const uint32_t fb = umma::smem_u32(&full[s]);
        if (dmode == 2) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
          continue;
        }
        umma::mbar_expect_tx(fb, S::kStage);
        const uint32_t sa = umma::smem_u32(smem + s * S::kStage);
        const uint32_t sb = sa + S::kABytes;
        if (g.conv == 0) {
          umma::tma_load_3d(sa, &tmA, fb, kb * kBK, m_blk * 128, z);
        } else {
          const int cblocks = g.cin / kBK;
          const int tap = kb / cblocks, cb = kb % cblocks;
          int dx = 0, dy = 0;
          if (g.conv == 1) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          umma::tma_load_3d(sa, &tmA, fb, cb * kBK, px0 + dx, py0 + dy);
        }
        umma::tma_load_3d(sb, &tmB, fb, kb * kBK, n_blk * BN, z);
      }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma::make_idesc(BN);
      int kit = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;  // accumulator stage
      umma::mbar_wait_dbg(umma::smem_u32(&tmem_empty[as]), ((it >> 1) & 1) ^ 1, spin);
      umma::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < nkb; ++kb, ++kit) {
        const int s = kit % STAGES;
        const uint32_t ph = (kit / STAGES) & 1;
        umma::mbar_wait_dbg(umma::smem_u32(&full[s]), ph, spin);
        umma::tc_fence_after();
        const uint32_t sa = umma::smem_u32(smem + s * S::kStage);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t da = umma::make_desc(sa);
        const uint64_t db = umma::make_desc(sb);
        if (dmode != 1) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16-byte units
            umma::mma_f16(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
        }
        umma::umma_commit(umma::smem_u32(&empty[s]));  // frees this ring slot when the MMAs retire
      }
      umma::umma_commit(umma::smem_u32(&tmem_full[as]));  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quarter, each taking half of the tile's 32-column chunks.  Every register
    // array below is indexed with compile-time constants only (fully unrolled, predicated) so the 32
    // accumulator values of a chunk stay in registers.
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which share of the chunks (0 .. EW/4 - 1)
    constexpr int kChunks = BN / 32;
    constexpr int kStride = EW / 4;
    const bool split = EW == 16 && resid_split_ok(g) && dmode == 0;  // (8-warp kernels: 76 registers, two CTAs per SM)
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    int n_blk, m_blk, z, px0, py0;
    decode(tile, n_blk, m_blk, z, px0, py0);
    const int as = it & 1;
    const int r = q * 32 + lane;  // accumulator row inside the tile
    int m;                        // logical output row
    bool row_ok;
    if (g.conv) {
      int ly = r / g.tw, lx = r % g.tw;
      int y = py0 + ly, x = px0 + lx;
      row_ok = (y < g.imgH) && (x < g.imgW);
      m = y * g.imgW + x;
    } else {
      m = m_blk * 128 + r;
      row_ok = m < g.M;
    }
    uint32_t xr[32];  // residual-stream slice of this thread's next chunk, in flight while the MMAs run
    if (split && row_ok && half < kChunks) resid_load(g, m, n_blk * BN + half * 32, xr);
    umma::mbar_wait_dbg(umma::smem_u32(&tmem_full[as]), (it >> 1) & 1, spin);
    umma::tc_fence_after();
    const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
    float head_acc = 0.f;
#pragma unroll 1
    for (int ci = half; ci < kChunks; ci += kStride) {
      if (dmode == 4) break;
      const int c0 = ci * 32;
      uint32_t v[32];
      umma::tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      const int n0 = n_blk * BN + c0;
      if (dmode == 5) {
        if (v[0] == 0x7fc12345u && v[17] == 0x12345u) g.out_f32[0] = 1.f;  // keep the load alive
        continue;
      }
      if (split) {
        if (row_ok) {
          resid_apply_store(g, v, m, n0, xr);
          if (ci + kStride < kChunks) resid_load(g, m, n0 + kStride * 32, xr);
        }
        continue;
      }
      gemm_epilogue_chunk(g, v, m, z, n0, row_ok, head_acc);
    }
    // DPT head: N == 32 is a single chunk, owned by the half-0 warps
    if (g.epi == EPI_HEAD && half == 0 && row_ok && n_blk == 0) g.out_f32[m] = fmaxf(head_acc + g.b3p[0], 0.f);
    // this accumulator stage may be overwritten by the MMA warp
    umma::tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&tmem_empty[as])) : "memory");
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace vd3d
