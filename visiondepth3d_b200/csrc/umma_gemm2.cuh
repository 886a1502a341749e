// umma_gemm2.cuh -- CTA-pair variant of k_umma_gemm (tcgen05.mma.cta_group::2).
//
// Two CTAs of one cluster (the two SMs of a TPC) compute one 256 x BN output tile: each CTA stages ITS 128
// rows of A and ITS BN/2 rows of B (TMA, 128B swizzle) and owns the 128 x BN fp32 accumulator of its rows in
// its own tensor memory; the leader CTA (cluster rank 0) issues M=256 MMAs that read both CTAs' shared memory.
// Per 1 M MACs a CTA pulls 16 KB (BN=256) / 24 KB (BN=128) through L2->smem instead of the 32 KB of the
// 128 x 128 single-CTA tile -- the single-CTA kernel is L2->smem fill bound.
//
// Barrier protocol (cutlass's 2-SM pipeline shape, restated):
//   full[s]        leader's barrier, count 1: the leader's producer arrives with expect_tx of BOTH CTAs' bytes,
//                  both producers' TMA loads complete_tx on it (cp.async.bulk.tensor ... cta_group::2);
//   empty[s]       one per CTA, count 1: the leader's tcgen05.commit multicasts the arrive to both CTAs;
//   tmem_full[a]   one per CTA, count 1: multicast commit after the last k-block of a tile;
//   tmem_empty[a]  leader's barrier, count 16: 8 epilogue warps of each CTA arrive (remote arrive from rank 1).
// Same GemmArgs / epilogues as k_umma_gemm (gemm_epilogue_chunk); g.mt counts 128-row tiles, pairs take two.
#pragma once
#include "umma_gemm.cuh"

namespace vd3d {
namespace umma {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire (the arrivals come from the peer CTA / the peer's async proxy)
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cl_spin(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cl_dbg(uint32_t bar, uint32_t parity, int spin) {
  if (spin)
    mbar_wait_cl_spin(bar, parity);
  else
    mbar_wait_cl(bar, parity);
}
// TMA load into THIS CTA's shared memory whose completion bytes are signalled on a barrier of the pair's leader
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(dst),
      "l"((uint64_t)tm), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far retire) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// kind::f16 instruction descriptor with explicit M (256 for the CTA pair)
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace umma

template <int BN, int STAGES>  // BN = N extent of the pair's tile; each CTA stages BN/2 rows of B
struct Gemm2Smem {
  static constexpr int kABytes = 128 * kBK * 2;
  static constexpr int kBBytes = (BN / 2) * kBK * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kTotal = STAGES * kStage + 1024 /*align*/ + 256 /*barriers*/;
};

// Epilogue warps per CTA (a multiple of 4: TMEM lane quarters).  16 were tried for the bias + GELU epilogue of fc1 (one CTA
// per SM here against two of the single-CTA kernel): the 96-register cap of 576 threads spills the generic epilogue and
// every shape got slower (fc2 45.6 -> 51.2 us at M = 10123, fc1-Large 54.9 -> 57.5); 8 it is.
constexpr int kGemm2EpiWarps = 8;
constexpr int kGemm2Threads = 64 + 32 * kGemm2EpiWarps;

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemm2Threads, 1)
k_umma_gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using S = Gemm2Smem<BN, STAGES>;
  static_assert(2 * BN <= 512, "two accumulator stages must fit tensor memory");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // same offset in both CTAs
  uint64_t* bars = (uint64_t*)(smem + STAGES * S::kStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = umma::cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int mtp = (g.mt + 1) >> 1;  // pair tiles along M
  const int total_tiles = g.nt * mtp * g.nz;
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
      umma::mbar_init(umma::smem_u32(&tmem_empty[i]), 2 * kGemm2EpiWarps);
    }
    umma::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) umma::tmem_alloc_2sm(umma::smem_u32(tmem_slot), 2 * BN);
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / complete_tx
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (n block, this CTA's 128-row block, batch, pixel-tile origin)
  auto decode = [&](int tile, int& n_blk, int& m_blk, int& z, int& px0, int& py0) {
    n_blk = tile % g.nt;
    int rem = tile / g.nt;
    m_blk = (rem % mtp) * 2 + (int)rank;
    z = rem / mtp;
    px0 = py0 = 0;
    if (g.conv) {
      int tiles_x = (g.imgW + g.tw - 1) / g.tw;
      py0 = (m_blk / tiles_x) * g.th;  // past the image for the odd tail tile: TMA zero-fills, rows are masked
      px0 = (m_blk % tiles_x) * g.tw;
    }
  };

  if (dmode == 3) {
    // tuning: prologue + teardown only
  } else if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int kit = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int n_blk, m_blk, z, px0, py0;
        decode(tile, n_blk, m_blk, z, px0, py0);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          umma::mbar_wait_cl_dbg(umma::smem_u32(&empty[s]), ph ^ 1, spin);
          const uint32_t fb_local = umma::smem_u32(&full[s]);
          if (dmode == 2) {
            if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb_local) : "memory");
            continue;
          }
          if (rank == 0) umma::mbar_expect_tx(fb_local, 2 * S::kStage);
          const uint32_t fb = umma::map_to_rank(fb_local, 0);
          const uint32_t sa = umma::smem_u32(smem + s * S::kStage);
          const uint32_t sb = sa + S::kABytes;
          if (g.conv == 0) {
            umma::tma_load_3d_2sm(sa, &tmA, fb, kb * kBK, m_blk * 128, z);
          } else {
            const int cblocks = g.cin / kBK;
            const int tap = kb / cblocks, cb = kb % cblocks;
            int dx = 0, dy = 0;
            if (g.conv == 1) {
              dy = tap / 3 - 1;
              dx = tap % 3 - 1;
            }
            umma::tma_load_3d_2sm(sa, &tmA, fb, cb * kBK, px0 + dx, py0 + dy);
          }
          umma::tma_load_3d_2sm(sb, &tmB, fb, kb * kBK, n_blk * BN + (int)rank * (BN / 2), z);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma::make_idesc_mn(256, BN);
      int kit = 0, it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
        const int as = it & 1;
        umma::mbar_wait_cl_dbg(umma::smem_u32(&tmem_empty[as]), ((it >> 1) & 1) ^ 1, spin);
        umma::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          umma::mbar_wait_cl_dbg(umma::smem_u32(&full[s]), ph, spin);
          umma::tc_fence_after();
          const uint32_t sa = umma::smem_u32(smem + s * S::kStage);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t da = umma::make_desc(sa);
          const uint64_t db = umma::make_desc(sb);
          if (dmode != 1) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma::mma_f16_2sm(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma::umma_commit_2sm(umma::smem_u32(&empty[s]), 3);  // frees the slot in both CTAs
        }
        umma::umma_commit_2sm(umma::smem_u32(&tmem_full[as]), 3);  // accumulators complete in both CTAs
      }
    }
  } else {
    // ===================== epilogue (warps 2.., both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // which share of the tile's 32-column chunks (0 .. kGemm2EpiWarps/4 - 1)
    constexpr int kChunks = BN / 32;
    constexpr int kStride = kGemm2EpiWarps / 4;
    const bool split = resid_split_ok(g) && dmode == 0;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
      int n_blk, m_blk, z, px0, py0;
      decode(tile, n_blk, m_blk, z, px0, py0);
      const int as = it & 1;
      const int r = q * 32 + lane;
      int m;
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
      uint32_t xr[32];  // residual-stream slice of the next chunk, in flight while the MMAs run (see resid_load)
      if (split && row_ok && half < kChunks) resid_load(g, m, n_blk * BN + half * 32, xr);
      umma::mbar_wait_cl_dbg(umma::smem_u32(&tmem_full[as]), (it >> 1) & 1, spin);
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
      if (g.epi == EPI_HEAD && half == 0 && row_ok && n_blk == 0) g.out_f32[m] = fmaxf(head_acc + g.b3p[0], 0.f);
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive_cluster(umma::map_to_rank(umma::smem_u32(&tmem_empty[as]), 0));
    }
  }

  // neither CTA may exit (or free tensor memory) while the other still signals it / reads its shared memory
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();
  if (warp == 1) umma::tmem_dealloc_2sm(tmem_base, 2 * BN);
}

}  // namespace vd3d
