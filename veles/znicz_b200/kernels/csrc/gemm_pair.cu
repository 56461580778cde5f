// Dense GEMM for large shapes, Blackwell 2-CTA form (sm_100a):
//
//   D[M, N] = act(alpha * sum_k A[m, k] * B[n, k] + bias[n])        A, B bf16 K-major, D bf16 / fp32
//
// * a CLUSTER OF TWO CTAs (one TPC) owns a 256 x BN tile: `tcgen05.mma.cta_group::2` with
//   UMMA_M = 256, each CTA stages its own 128 rows of A and only HALF of the B tile (BN / 2 rows)
//   - the tensor core reads both halves across the pair, so B's shared-memory and L2 traffic per
//   SM is halved against two independent 128-row tiles;
// * TMA loads `.cta_group::2`: both CTAs' copies complete on the LEADER's mbarrier (one wait for
//   the MMA issuer), a deep ring (6 x 32 KB stages at BN = 256);
// * PERSISTENT: 74 clusters loop over the tiles (grouped raster: 8 M-tiles x all N-tiles per
//   group, so a wave re-reads A and B from L2), the accumulator is DOUBLE BUFFERED in TMEM
//   (2 x BN fp32 columns = all 512 columns at BN = 256): the epilogue of tile t (TMEM -> registers
//   -> bias / activation -> swizzled smem -> TMA store) overlaps the MMAs of tile t + 1;
// * warp roles: 0 = TMA producer, 1 = MMA issuer (leader CTA only), 2 = TMEM allocation,
//   4..7 = epilogue (lane l of warp 4 + w owns accumulator row 32 w + l).
//
// Used for the 8192^3-class GEMMs and the large fully-connected layers (perf shape of the
// reference: /root/reference/tests/unit/test_all2all.py:98-127); skinny / small shapes stay on
// gemm_umma.cu. The single-CTA kernel reached 46 % of cuBLAS on 8192^3 (profiles/roofline_r1.md).
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace zn {

using namespace umma;

namespace pr {

constexpr int BM = 128;                 // rows per CTA; the pair's UMMA_M is 256
constexpr int BK = 64;                  // bf16 elements = 128 bytes = one SW128 row
constexpr int A_BYTES = BM * 128;
constexpr int EPI_BUF = 128 * 128;      // 128 rows x 128 bytes per staging buffer
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit: "same variable in CTA 0"

// A operand: dense [M][K] tile, or the implicit-GEMM operand of a convolution fetched by TMA in
// im2col mode (one [128 pixels x 64 channels] box per k-block): fprop = im2col(x), dgrad
// (unit stride) = windows of err_out for the flipped filter.
enum { AM_DENSE = 0, AM_IM2COL_FPROP = 1, AM_IM2COL_DGRAD = 2,
       AM_DENSE_MN = 3 };          // A stored [K][M] (M contiguous): err^T of the weight-gradient GEMMs
// B operand: K-major tile, MN-major tile ([K][N], N contiguous), or - conv wgrad - the im2col
// operand as MN-major blocks fetched by TMA im2col ([64 pixels][64 channels of one tap])
enum { BM_K = 0, BM_MN = 1, BM_IM2COL_MN = 2 };

struct ConvP { int H, W, OH, OW, KY, KX, SY, SX, PT, PL, inner; };

struct Params {
  int M, N, K;
  const float* bias; int act; float alpha;
  int tiles_m, tiles_n;
  int splits, kb_per_split;       // split-K over the reduction (fp32 partials [splits][M][N])
  ConvP g;
  const __nv_bfloat16* dmul; int dact; long long ldo;   // dgrad: out *= f'(dmul[row][col])
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of both CTAs: 2 x 128 rows] * B[smem of both CTAs: 2 x BN/2 rows]
__device__ __forceinline__ void mma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all MMAs issued so far -> one arrival on `bar` in EVERY CTA of the mask
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// tile -> this CTA's shared memory, completion bytes -> the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// im2col-mode load (see umma.cuh::tma_load_im2col_4d), 2-CTA form: bytes complete on CTA 0's barrier
__device__ __forceinline__ void tma_load_im2col_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                    int c, int w, int h, int n, uint16_t off_w,
                                                    uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read(int keep) {
  if (keep == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

// grouped raster: GM consecutive M-tiles share a sweep over all N-tiles
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
  constexpr int GM = 8;
  const int per_group = GM * tiles_n;
  const int g = t / per_group, r = t - g * per_group;
  const int m0 = g * GM;
  const int gm = min(GM, tiles_m - m0);
  tn = r / gm;
  tm = m0 + (r - tn * gm);
}

template <int BN, bool OUT_F32, int AMODE = AM_DENSE, int BMODE = BM_K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
gemm_pair_k(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_c, const Params p) {
  constexpr int B_BYTES = (BN / 2) * 128;
  constexpr int STAGE = A_BYTES + B_BYTES;
  constexpr int NST = (BN == 256) ? 6 : 8;
  constexpr int EPI_COLS = OUT_F32 ? 32 : 64;          // 128-byte staging rows
  constexpr int NCHUNK = BN / EPI_COLS;
  constexpr bool A_MN = (AMODE == AM_DENSE_MN);
  constexpr bool B_MN = (BMODE != BM_K);
  constexpr uint32_t IDESC = make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
  constexpr uint32_t TMEM_COLS = 2 * BN;               // double-buffered accumulator (256 or 512)

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NST];
  __shared__ __align__(8) uint64_t empty_bar[NST];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* tiles = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* epi = tiles + (size_t)NST * STAGE;          // 2 x EPI_BUF, 1024-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int splits = p.splits > 1 ? p.splits : 1;
  const int num_tiles = p.tiles_m * p.tiles_n * splits;      // work items (tile, k-split)
  const int total_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1u); mbar_init(&empty_bar[s], 1u); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1u); mbar_init(&tmem_empty_bar[a], 8u); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_c);
  }
  if (warp == 2) { tmem_alloc2(&tmem_base_smem, TMEM_COLS); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();            // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = cid; t < num_tiles; t += nclusters) {
        int tm, tn;
        tile_coords(t / splits, p.tiles_m, p.tiles_n, tm, tn);
        const int split = t % splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = splits > 1 ? min(total_kb, kb0 + p.kb_per_split) : total_kb;
        const int m0 = tm * 2 * BM + (int)rank * BM;
        const int n0 = tn * BN + (int)rank * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % NST; const uint32_t ph = (it / NST) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = tiles + (size_t)s * STAGE;
          // the leader's barrier expects the bytes of BOTH CTAs; the peer's copies may complete
          // before this expect_tx is posted (the tx-count just goes negative for a moment)
          uint32_t tx = 2u * STAGE;
          int nblk = BN / 128;                 // 64-wide B blocks this CTA loads (im2col B only)
          if (BMODE == BM_IM2COL_MN) {
            // blocks past the last reduction-weight index are not fetched: both CTAs' counts
            const int base = tn * BN;
            int cnt[2];
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
              const int left = p.N - (base + r2 * (BN / 2));
              cnt[r2] = left <= 0 ? 0 : min(BN / 128, (left + 63) / 64);
            }
            nblk = cnt[rank];
            tx = 2u * A_BYTES + (uint32_t)(cnt[0] + cnt[1]) * 8192u;
          }
          if (leader) mbar_arrive_expect_tx(&full_bar[s], tx);
          if (AMODE == AM_DENSE) {
            tma_load_2d_2sm(sa, &tmap_a, &full_bar[s], kb * BK, m0);
          } else if (AMODE == AM_DENSE_MN) {
            // A stored [K][M]: two [64 k][64 m] boxes for this CTA's 128 rows
            tma_load_2d_2sm(sa, &tmap_a, &full_bar[s], m0, kb * BK);
            tma_load_2d_2sm(sa + 8192, &tmap_a, &full_bar[s], m0 + 64, kb * BK);
          } else {
            // this CTA's 128 pixels start at m0; k-block = 64 channels of one filter tap
            const ConvP& g = p.g;
            const int k0 = kb * BK;
            const int tap = k0 / g.inner, c0 = k0 - tap * g.inner;
            const int ky = tap / g.KX, kx = tap - ky * g.KX;
            int w0, h0, n, ow, oh;
            if (AMODE == AM_IM2COL_FPROP) {
              const int q = m0 % g.OW; const int t2 = m0 / g.OW;
              const int pr = t2 % g.OH; n = t2 / g.OH;
              w0 = q * g.SX - g.PL; h0 = pr * g.SY - g.PT; ow = kx; oh = ky;
            } else {
              const int ix = m0 % g.W; const int t2 = m0 / g.W;
              const int iy = t2 % g.H; n = t2 / g.H;
              w0 = ix - (g.KX - 1 - g.PL); h0 = iy - (g.KY - 1 - g.PT);
              ow = g.KX - 1 - kx; oh = g.KY - 1 - ky;
            }
            tma_load_im2col_2sm(sa, &tmap_a, &full_bar[s], c0, w0, h0, n, (uint16_t)ow, (uint16_t)oh);
          }
          if (BMODE == BM_K) {
            tma_load_2d_2sm(sa + A_BYTES, &tmap_b, &full_bar[s], kb * BK, n0);
          } else if (BMODE == BM_MN) {
            // B stored [K][N] (N contiguous): 64 x 64 boxes, one per 64 columns of this CTA's half
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
              tma_load_2d_2sm(sa + A_BYTES + j * 8192, &tmap_b, &full_bar[s], n0 + j * 64, kb * BK);
          } else {
            // conv wgrad: reduction = pixels kb * 64 .., N = (tap, channel); one im2col box of
            // [64 pixels][64 channels] per 64-wide N block
            const ConvP& g = p.g;
            const int pix0 = kb * BK;
            const int q = pix0 % g.OW; const int t2 = pix0 / g.OW;
            const int pr = t2 % g.OH; const int n = t2 / g.OH;
            const int w0 = q * g.SX - g.PL, h0 = pr * g.SY - g.PT;
            for (int j = 0; j < nblk; ++j) {
              const int kidx0 = n0 + j * 64;
              const int tap = kidx0 / g.inner, c0 = kidx0 - tap * g.inner;
              tma_load_im2col_2sm(sa + A_BYTES + j * 8192, &tmap_b, &full_bar[s], c0, w0, h0, n,
                                  (uint16_t)(tap % g.KX), (uint16_t)(tap / g.KX));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int t = cid; t < num_tiles; t += nclusters, ++tcount) {
        const uint32_t acc = tcount & 1, acc_ph = (tcount >> 1) & 1;
        const int split = t % splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = splits > 1 ? min(total_kb, kb0 + p.kb_per_split) : total_kb;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);      // both CTAs' epilogues drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % NST; const uint32_t ph = (it / NST) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024)
                                     : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024)
                                     : make_smem_desc(sb + k * 32, 16, 1024);
            mma_f16_2sm(d_tmem, da, db, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          mma_commit_2sm(&empty_bar[s], 3);              // frees the stage in both CTAs
        }
        mma_commit_2sm(&tmem_full_bar[acc], 3);          // accumulator ready in both CTAs
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): TMEM -> regs -> smem -> TMA store ============
    const int ew = warp - 4;                              // TMEM lanes 32 ew .. 32 ew + 31
    const int et = threadIdx.x - 128;                     // 0..127 = row of the CTA's sub-tile
    uint32_t tcount = 0, chunk_no = 0;
    for (int t = cid; t < num_tiles; t += nclusters, ++tcount) {
      int tm, tn;
      tile_coords(t / splits, p.tiles_m, p.tiles_n, tm, tn);
      const int split = t % splits;
      const int m0 = tm * 2 * BM + (int)rank * BM;
      const int n0 = tn * BN;
      // split-K partial s lives at rows [s * M, (s + 1) * M) of the output map: a CTA whose 128
      // rows lie entirely past M must not store (it would land in the next partial)
      const bool store_ok = m0 < p.M;
      const int row0 = split * p.M + m0;
      const uint32_t acc = tcount & 1, acc_ph = (tcount >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c, ++chunk_no) {
        uint8_t* buf = epi + (chunk_no & 1) * EPI_BUF;
        // the TMA store that read this buffer two chunks ago must be done with it
        if (et == 0) tma_store_wait_read(1);
        named_bar_sync(1, 128);
        uint32_t r[32];
        uint8_t* rowp = buf + et * 128;
        const int sw = et & 7;
        if (OUT_F32) {
          tmem_ld_32x32(trow + c * 32, r);
          tmem_ld_wait();
          const int nb = n0 + c * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {                   // 8 chunks of 4 floats = 16 bytes
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nb + q * 4 + j;
              float x = __uint_as_float(r[q * 4 + j]) * p.alpha;
              if (p.bias != nullptr && n < p.N) x += __ldg(p.bias + n);
              v[j] = act_fwd5_fast(p.act, x);
            }
            *reinterpret_cast<float4*>(rowp + ((q ^ sw) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
          }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld_32x32(trow + c * 64 + h * 32, r);
            tmem_ld_wait();
            const int nb = n0 + c * 64 + h * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {                 // 4 chunks of 8 bf16 = 16 bytes
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int n = nb + q * 8 + j;
                float x = __uint_as_float(r[q * 8 + j]) * p.alpha;
                if (p.bias != nullptr && n < p.N) x += __ldg(p.bias + n);
                v[j] = act_fwd5_fast(p.act, x);
              }
              if (p.dmul != nullptr) {
                // err_input *= f'(x): x has the layout of the output (row pitch ldo)
                const int row = m0 + et;
                const int n8 = nb + q * 8;
                if (row < p.M && n8 + 8 <= p.N) {
                  float xv[8];
                  ld8(p.dmul + (long long)row * p.ldo + n8, xv);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] *= act_deriv(p.dact, 0.f, xv[j]);
                } else if (row < p.M) {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (n8 + j < p.N)
                      v[j] *= act_deriv(p.dact, 0.f, __bfloat162float(p.dmul[(long long)row * p.ldo + n8 + j]));
                }
              }
              st8(reinterpret_cast<__nv_bfloat16*>(rowp + (((h * 4 + q) ^ sw) << 4)), v);
            }
          }
        }
        if (c == NCHUNK - 1) {
          // last TMEM read of this accumulator: hand the buffer back to the MMA issuer (CTA 0)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[acc], 0);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (et == 0) {
          if (store_ok) tma_store_2d(&tmap_c, buf, n0 + c * EPI_COLS, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();            // the pair retires together: no remote arrival / smem read outlives a CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}
// 2-D map, inner dim contiguous, row pitch `ld` elements, box {box_inner, box_rows}, SWIZZLE_128B
static int make_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int esize, long long inner,
                    long long outer, long long ld, int box_inner, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * esize)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int BN, bool OUT_F32, int AMODE = AM_DENSE, int BMODE = BM_K>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Params& p,
                  cudaStream_t st) {
  constexpr int NST = (BN == 256) ? 6 : 8;
  constexpr int smem = NST * (A_BYTES + (BN / 2) * 128) + 2 * EPI_BUF + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_k<BN, OUT_F32, AMODE, BMODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int sms = 148;
  {
    static int cached = 0;
    if (!cached) {
      int dev = 0; cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
      if (cached < 2) cached = 2;
    }
    sms = cached;
  }
  const int tiles = p.tiles_m * p.tiles_n * (p.splits > 1 ? p.splits : 1);
  int clusters = sms / 2;
  if (clusters > tiles) clusters = tiles;
  gemm_pair_k<BN, OUT_F32, AMODE, BMODE><<<dim3(2 * clusters), dim3(256), smem, st>>>(ta, tb, tc, p);
  return (int)cudaGetLastError();
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// im2col-mode map over NHWC bf16 [N][H][W][C]: 128 window positions x 64 channels per load
static int make_map_im2col(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int lower_w,
                           int lower_h, int upper_w, int upper_h, int stride_w, int stride_h,
                           int pixels = BM) {
  static EncodeIm2colFn enc = nullptr;
  if (!enc) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &q, cudaEnableDefault, &qr) != cudaSuccess || !q)
      return -1;
    enc = reinterpret_cast<EncodeIm2colFn>(q);
  }
  const int v[6] = {lower_w, lower_h, upper_w, upper_h, 0, 0};
  for (int i = 0; i < 4; ++i) if (v[i] < -128 || v[i] > 127) return -7;
  if (stride_w > 8 || stride_h > 8) return -7;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1u, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower,
                   upper, 64u, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
static bool conv_pair_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ZNICZ_CONV_PAIR"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on != 0;
}

}  // namespace pr

// Convolution forward as an implicit GEMM on the 2-CTA persistent kernel: A = im2col(x) by TMA
// (im2col mode), B = w_lp [F][ldw] (K = (tap, c) contiguous). out[pix][f] = act(. + bias[f]).
// Returns 0 when launched, non-zero when the geometry is not served here (caller falls back).
int launch_conv_fprop_pair(const void* x, const void* w_lp, long long ldw, const float* bias, void* out,
                           int N, int H, int W, int C, int OH, int OW, int F, int KY, int KX, int SY,
                           int SX, int PT, int PL, int act, cudaStream_t st) {
  using namespace pr;
  if (!conv_pair_enabled()) return -8;
  const int M = N * OH * OW, K = KY * KX * C;
  if ((C % 64) || F < 128 || (F % 8) || M < 1024 || (ldw % 8)) return -6;
  if (((uintptr_t)x & 15) || ((uintptr_t)w_lp & 15) || ((uintptr_t)out & 15)) return -3;
  const int bn = (F % 256 == 0) ? 256 : 128;
  CUtensorMap ta, tb, tc;
  int r = make_map_im2col(&ta, x, N, H, W, C, -PL, -PT, (OW - 1) * SX + 1 - PL - W,
                          (OH - 1) * SY + 1 - PT - H, SX, SY);
  if (r) return r;
  r = make_map(&tb, w_lp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, F, ldw, 64, bn / 2);
  if (r) return r;
  r = make_map(&tc, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, F, M, F, 64, BM);
  if (r) return r;
  Params p{};
  p.M = M; p.N = F; p.K = K; p.bias = bias; p.act = act; p.alpha = 1.f;
  p.tiles_m = (M + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (F + bn - 1) / bn;
  p.g = ConvP{H, W, OH, OW, KY, KX, SY, SX, PT, PL, C};
  p.ldo = F;
  p.splits = 1; p.kb_per_split = (K + BK - 1) / BK;
  return bn == 256 ? launch<256, false, AM_IM2COL_FPROP, BM_K>(ta, tb, tc, p, st)
                   : launch<128, false, AM_IM2COL_FPROP, BM_K>(ta, tb, tc, p, st);
}

// Unit-stride conv dgrad on the same kernel: A = windows of err_out [N][OH][OW][F] for the flipped
// filter (TMA im2col), B = wd_lp [(tap, f)][ldc] (N = input channels contiguous),
// err_in[ipix][c] = alpha * (.) (* f'(dmul)).
int launch_conv_dgrad_pair(const void* err_out, const void* wd_lp, long long ldc, void* err_in, int N,
                           int H, int W, int C, int OH, int OW, int F, int KY, int KX, int PT, int PL,
                           float alpha, const void* dmul, int dact, cudaStream_t st) {
  using namespace pr;
  if (!conv_pair_enabled()) return -8;
  const int M = N * H * W, K = KY * KX * F;
  if ((F % 64) || C < 128 || (C % 8) || M < 1024 || (ldc % 8)) return -6;
  if (((uintptr_t)err_out & 15) || ((uintptr_t)wd_lp & 15) || ((uintptr_t)err_in & 15)) return -3;
  const int bn = (C % 256 == 0) ? 256 : 128;
  CUtensorMap ta, tb, tc;
  const int lw = -(KX - 1 - PL), lh = -(KY - 1 - PT);
  int r = make_map_im2col(&ta, err_out, N, OH, OW, F, lw, lh, W + lw - OW, H + lh - OH, 1, 1);
  if (r) return r;
  // B stored [K][ldc]: inner dim = N (channels), boxes of 64 x 64
  r = make_map(&tb, wd_lp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ldc, K, ldc, 64, 64);
  if (r) return r;
  r = make_map(&tc, err_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, M, C, 64, BM);
  if (r) return r;
  Params p{};
  p.M = M; p.N = C; p.K = K; p.bias = nullptr; p.act = 0; p.alpha = alpha;
  p.tiles_m = (M + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (C + bn - 1) / bn;
  p.g = ConvP{H, W, OH, OW, KY, KX, 1, 1, PT, PL, F};
  p.dmul = (const __nv_bfloat16*)dmul; p.dact = dmul ? dact : 0; p.ldo = C;
  p.splits = 1; p.kb_per_split = (K + BK - 1) / BK;
  return bn == 256 ? launch<256, false, AM_IM2COL_DGRAD, BM_MN>(ta, tb, tc, p, st)
                   : launch<128, false, AM_IM2COL_DGRAD, BM_MN>(ta, tb, tc, p, st);
}

bool conv_pair_on() { return pr::conv_pair_enabled(); }

// Conv weight gradient on the 2-CTA kernel: partials[s][f][kidx] (fp32) = sum over the pixels of
// split s of err_out[pix][f] * im2col(x)[pix][kidx]. A = err_out^T (stored [pix][F]: MN-major),
// B = im2col(x) as MN-major [64 pixels][64 channels] boxes (TMA im2col); the output is already in
// the weights' [F][Kw] layout - no transposed store. F % 128 == 0, C % 64 == 0.
int launch_conv_wgrad_pair(const void* err_out, const void* x, float* partials, int splits, int N, int H,
                           int W, int C, int OH, int OW, int F, int KY, int KX, int SY, int SX, int PT,
                           int PL, cudaStream_t st) {
  using namespace pr;
  if (!conv_pair_enabled()) return -8;
  const int Kw = KY * KX * C, P = N * OH * OW;
  if ((C % 64) || (F % 128) || F < 128 || Kw < 128 || P < 4096) return -6;
  if (((uintptr_t)err_out & 15) || ((uintptr_t)x & 15) || ((uintptr_t)partials & 15)) return -3;
  const int bn = Kw >= 256 ? 256 : 128;
  const int total_kb = (P + BK - 1) / BK;
  if (splits < 1) splits = 1;
  const int kbs = (total_kb + splits - 1) / splits;
  if ((long long)(splits - 1) * kbs >= total_kb) return -6;      // an empty split: not here
  CUtensorMap ta, tb, tc;
  int r = make_map(&ta, err_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, F, P, F, 64, 64);
  if (r) return r;
  r = make_map_im2col(&tb, x, N, H, W, C, -PL, -PT, (OW - 1) * SX + 1 - PL - W,
                      (OH - 1) * SY + 1 - PT - H, SX, SY, 64);
  if (r) return r;
  r = make_map(&tc, partials, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Kw, (long long)splits * F, Kw, 32, BM);
  if (r) return r;
  Params p{};
  p.M = F; p.N = Kw; p.K = P; p.bias = nullptr; p.act = 0; p.alpha = 1.f;
  p.tiles_m = (F + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (Kw + bn - 1) / bn;
  p.splits = splits; p.kb_per_split = kbs;
  p.g = ConvP{H, W, OH, OW, KY, KX, SY, SX, PT, PL, C};
  p.ldo = Kw;
  return bn == 256 ? launch<256, true, AM_DENSE_MN, BM_IM2COL_MN>(ta, tb, tc, p, st)
                   : launch<128, true, AM_DENSE_MN, BM_IM2COL_MN>(ta, tb, tc, p, st);
}
// splits that give the 74 clusters of the pair kernel ~1 work item each (0: shape not served)
int pair_wgrad_splits(int Kw, int F, int P, int max_splits) {
  using namespace pr;
  if (!conv_pair_enabled() || (Kw % 64) || (F % 128) || F < 128 || Kw < 128 || P < 4096) return 0;
  const int bn = Kw >= 256 ? 256 : 128;
  const long long tiles = (long long)((F + 255) / 256) * ((Kw + bn - 1) / bn);
  int s = (int)((74 + tiles - 1) / tiles);
  const int total_kb = (P + BK - 1) / BK;
  if (s > total_kb / 8) s = total_kb / 8;
  if (s > max_splits) s = max_splits;
  return s < 1 ? 1 : s;
}

// Fully-connected weight gradient: out[n_out][n_in] (fp32) = err^T . x with err stored
// [batch][n_out] and x stored [batch][n_in] (both MN-major operands, K = batch).
int launch_fc_wgrad_pair(const void* err, long long lde, const void* x, long long ldx, float* out,
                         long long ldo, int n_out, int n_in, int batch, cudaStream_t st) {
  using namespace pr;
  if (!conv_pair_enabled()) return -8;
  if (n_out < 256 || n_in < 256 || (lde % 8) || (ldx % 8) || (ldo % 4) || batch < 64) return -6;
  if (((uintptr_t)err & 15) || ((uintptr_t)x & 15) || ((uintptr_t)out & 15)) return -3;
  const int bn = 256;
  CUtensorMap ta, tb, tc;
  int r = make_map(&ta, err, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_out, batch, lde, 64, 64);
  if (r) return r;
  r = make_map(&tb, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, n_in, batch, ldx, 64, 64);
  if (r) return r;
  r = make_map(&tc, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, n_in, n_out, ldo, 32, BM);
  if (r) return r;
  Params p{};
  p.M = n_out; p.N = n_in; p.K = batch; p.alpha = 1.f;
  p.tiles_m = (n_out + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (n_in + bn - 1) / bn;
  p.splits = 1; p.kb_per_split = (batch + BK - 1) / BK;
  p.ldo = ldo;
  return launch<256, true, AM_DENSE_MN, BM_MN>(ta, tb, tc, p, st);
}

// a [M][lda] bf16, b [N][ldb] bf16 (both K contiguous), out [M][ldo] bf16 or fp32.
// Returns 0, or non-zero when the shape / alignment is not served by this kernel (nothing launched).
int launch_gemm_pair(const void* a, long long lda, const void* b, long long ldb, void* out, int out_f32,
                     long long ldo, int M, int N, int K, const float* bias, int act, float alpha,
                     cudaStream_t st) {
  using namespace pr;
  const int esz = out_f32 ? 4 : 2;
  if ((lda % 8) || (ldb % 8) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) return -3;
  if (((ldo * esz) % 16) || ((uintptr_t)out & 15)) return -3;
  if (M < 256 || N < 128 || K < 64) return -6;                // skinny / tiny: gemm_umma.cu
  const int bn = (N % 256 == 0 || N >= 1024) ? 256 : 128;
  CUtensorMap ta, tb, tc;
  int r = make_map(&ta, a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, M, lda, 64, BM);
  if (r) return r;
  r = make_map(&tb, b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, N, ldb, 64, bn / 2);
  if (r) return r;
  r = make_map(&tc, out, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, esz,
               N, M, ldo, out_f32 ? 32 : 64, BM);
  if (r) return r;
  Params p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.act = act; p.alpha = alpha;
  p.tiles_m = (M + 2 * BM - 1) / (2 * BM);
  p.tiles_n = (N + bn - 1) / bn;
  p.splits = 1; p.kb_per_split = (K + BK - 1) / BK;
  if (bn == 256) return out_f32 ? launch<256, true>(ta, tb, tc, p, st) : launch<256, false>(ta, tb, tc, p, st);
  return out_f32 ? launch<128, true>(ta, tb, tc, p, st) : launch<128, false>(ta, tb, tc, p, st);
}

}  // namespace zn
