// Persistent LSTM sequence kernels (sm_100a): the whole time loop of an LSTM layer in ONE launch.
//
//   forward   z_t = [x_t | h_{t-1}] . W^T + b,  (i, f, g, o) = act(z_t),  c_t = i g + f c_{t-1},
//             h_t = o * A tanh(B c_t)                         (/root/reference/lstm.py:75-143)
//   backward  dh_t = err_t + dz_{t+1} . W_h,  cell derivative -> dz_t,  dc_{t-1} = dc_t f
//
// The per-time-step path is 2-3 launches per step (gemm + cell kernel, ~4 us each: the recurrence
// is latency bound). Here a CLUSTER of H / 32 CTAs owns R batch rows (R = 32 / 64 / 128, chosen so
// that the clusters spread over the SMs) for all T steps:
//
// forward
// * CTA c owns hidden units [32 c, 32 c + 32): the 4 x 32 rows of W of their gates stay RESIDENT
//   in shared memory for the whole sequence (96 KB at I = 128, H = 256), loaded once by TMA;
// * per step ONE tcgen05 accumulation D[128 x 128] in TMEM (double buffered): the x part (K = I,
//   no recurrence) is issued a step ahead, the h part (K = H) the moment h_{t-1} has arrived;
// * the cell math runs in the epilogue straight from TMEM (thread = batch row x 16 units), c_t
//   lives in registers across the steps;
// * the cluster exchanges h_t through L2: bf16 stores into the [x | h] operand of step t + 1 ->
//   fence.proxy.async -> barrier.cluster (release / acquire) -> TMA loads. (Measured
//   alternatives, tools/lstm_probe.py: pushing h into the 8 peers with st.shared::cluster costs
//   64 KB of DSMEM stores per CTA and step at ~21 B/clk - 3x slower than the 64 KB TMA load.)
// * critical path first: h is stored and announced BEFORE the state for the backward pass is
//   written; that state (gates as bf16, c as fp32) uses a lane-major private layout in which every
//   warp-wide store is one contiguous 512-byte piece (row-major 64-byte pieces cost 5.9 of 9.2 us
//   per step in the first version).
// backward
// * dz_t of the CTA's own units goes from the epilogue straight into its own A operand (swizzled
//   smem) and is multiplied by the resident W_h rows of those gates: P_c[R x H] = dz_t[:, own] .
//   W_h[own, :] (K = 128, N = H, one accumulation per step);
// * the partial sums are reduce-scattered through L2 as bf16 in the same lane-major layout
//   (coalesced on both sides) under barrier.cluster; CTA d adds the H / 32 pieces of its units in
//   fp32 -> dh_{t-1}; dc lives in registers.
// Launch-count: 2 (+ the W_h^T permutation) instead of ~5 T.
#include "common.cuh"
#include "umma.cuh"

namespace zn {

using namespace umma;

namespace lp {

constexpr int BM = 128;
constexpr int U = 32;                   // hidden units per CTA
constexpr int UH = 16;                  // ... per epilogue thread (two warps share a TMEM lane quarter)
constexpr int NG = 4 * U;               // gate columns per CTA
constexpr int THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int SVEC = 12;                // 16-byte vectors per thread and step in the private state layout

__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
// generic-proxy global stores of this thread become visible to later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(uint4 u, float* v) {
  v[0] = bf_lo(u.x); v[1] = bf_hi(u.x); v[2] = bf_lo(u.y); v[3] = bf_hi(u.y);
  v[4] = bf_lo(u.z); v[5] = bf_hi(u.z); v[6] = bf_lo(u.w); v[7] = bf_hi(u.w);
}
// byte offset of 16-byte chunk `ch` of row `r` in a [128 rows][128 B] SWIZZLE_128B K-major block
__device__ __forceinline__ uint32_t sw128(int r, int ch) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4));
}

// Private state layout (forward -> backward), 16-byte vectors:
//   ((t * n_ctas + cta) * 8 + epilogue warp) * SVEC + v) * 32 + lane
// v = 0,1: input gate of the thread's 16 units as bf16, 2,3 forget, 4,5 memory maker, 6,7 output gate,
// 8..11: cell state fp32. Every warp-wide access is one contiguous 512-byte piece.
__device__ __forceinline__ size_t state_idx(int t, int n_ctas, int cta, int ew, int v, int lane) {
  return ((((size_t)t * n_ctas + cta) * 8 + ew) * SVEC + v) * 32 + lane;
}

struct FwdP {
  int T, B, I, H, R;          // R: batch rows per cluster
  const float* bias;          // [4H] or null
  uint4* state;               // private layout, see state_idx
  __nv_bfloat16* xh;          // [T + 1][B][I + H]
  long long* dbg;             // optional clock64 stamps [T][8] of CTA 0 (tools/lstm_probe.py)
};

__global__ void __launch_bounds__(THREADS, 1)
lstm_fwd_k(const __grid_constant__ CUtensorMap t_xh, const __grid_constant__ CUtensorMap t_w, FwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KX = p.I / 64, KH = p.H / 64;
  uint8_t* sW = sm;                                   // (KX + KH) blocks x [NG rows][128 B]
  uint8_t* sX = sW + (size_t)(KX + KH) * NG * 128;    // 2 stages x KX blocks x [128 rows][128 B]
  uint8_t* sH = sX + (size_t)2 * KX * BM * 128;       // KH blocks x [128 rows][128 B]: h_{t-1}
  __shared__ __align__(8) uint64_t w_full, x_full[2], x_empty[2], h_full, acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[NG];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (int)cl_rank();
  const int csize = p.H / U;
  const int T = p.T, B = p.B, I = p.I, H = p.H, R = p.R;
  const int row0 = (blockIdx.x / csize) * R;           // first batch row of this cluster
  const uint32_t tile_bytes = (uint32_t)(R * 128);     // one k-block of R rows

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1);
    mbar_init(&h_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1);
      mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], (uint32_t)(2 * (R / 32)));
    }
    fence_barrier_init();
    tma_prefetch_desc(&t_xh);
    tma_prefetch_desc(&t_w);
  }
  if (threadIdx.x < NG) {
    const int g = threadIdx.x / U, j = threadIdx.x % U;
    s_bias[threadIdx.x] = p.bias ? p.bias[g * H + c * U + j] : 0.f;
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, 2 * NG); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t IDESC = make_idesc_bf16(BM, NG, 0, 0);

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_full, (uint32_t)((KX + KH) * NG * 128));
      for (int kb = 0; kb < KX + KH; ++kb)
        for (int g = 0; g < 4; ++g)
          tma_load_2d(sW + (size_t)kb * NG * 128 + g * U * 128, &t_w, &w_full, kb * 64, g * H + c * U);
      for (int t0 = 0; t0 < 2 && t0 < T; ++t0) {
        mbar_arrive_expect_tx(&x_full[t0], KX * tile_bytes);
        for (int kb = 0; kb < KX; ++kb)
          tma_load_2d(sX + (size_t)(t0 * KX + kb) * BM * 128, &t_xh, &x_full[t0], kb * 64, t0 * B + row0);
      }
    }
    for (int t = 1; t < T; ++t) {
      cl_arrive(); cl_wait();                         // every CTA's slice of h_{t-1} is in xh[t]
      if (lane == 0) {
        fence_proxy_async_all();
        mbar_arrive_expect_tx(&h_full, KH * tile_bytes);
        for (int kb = 0; kb < KH; ++kb)
          tma_load_2d(sH + (size_t)kb * BM * 128, &t_xh, &h_full, I + kb * 64, t * B + row0);
        if (t + 1 < T) {
          const int s = (t + 1) & 1;                  // stage of x_{t+1}; its previous user was x_{t-1}
          mbar_wait(&x_empty[s], (((t + 1) >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&x_full[s], KX * tile_bytes);
          for (int kb = 0; kb < KX; ++kb)
            tma_load_2d(sX + (size_t)(s * KX + kb) * BM * 128, &t_xh, &x_full[s], kb * 64, (t + 1) * B + row0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    auto x_part = [&](int t) {
      const int s = t & 1;
      mbar_wait(&x_full[s], (t >> 1) & 1);
      if (t >= 2) mbar_wait(&acc_empty[s], ((t >> 1) - 1) & 1);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(s * NG);
      for (int kb = 0; kb < KX; ++kb) {
        const uint32_t sa = smem_u32(sX + (size_t)(s * KX + kb) * BM * 128);
        const uint32_t sb = smem_u32(sW + (size_t)kb * NG * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16(d, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), IDESC,
                  (kb > 0 || k > 0) ? 1u : 0u);
      }
      mma_commit(&x_empty[s]);
    };
    if (lane == 0) {
      mbar_wait(&w_full, 0);
      x_part(0);
      mma_commit(&acc_full[0]);                       // h_{-1} = 0: step 0 is the x part alone
      if (T > 1) x_part(1);
    }
    __syncwarp();
    for (int t = 1; t < T; ++t) {
      cl_arrive(); cl_wait();
      if (lane == 0) {
        mbar_wait(&h_full, (t - 1) & 1);
        if (p.dbg && blockIdx.x == 0) p.dbg[t * 8 + 2] = clock64();
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)((t & 1) * NG);
        for (int kb = 0; kb < KH; ++kb) {
          const uint32_t sa = smem_u32(sH + (size_t)kb * BM * 128);
          const uint32_t sb = smem_u32(sW + (size_t)(KX + kb) * NG * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16(d, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), IDESC, 1u);
        }
        mma_commit(&acc_full[t & 1]);
        if (t + 1 < T) x_part(t + 1);                 // no recurrence in it: runs under the epilogue
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue: cell math
    const int ew = warp - 2;
    const int q = warp & 3, half = ew >> 2;
    const int r = q * 32 + lane;
    const int b = row0 + r;
    const bool active = q * 32 < R;                    // warp-uniform
    const bool valid = active && b < B;
    const int j0 = c * U + half * UH;                  // first hidden unit of this thread
    const int n_ctas = gridDim.x;
    float cst[UH];
#pragma unroll
    for (int j = 0; j < UH; ++j) cst[j] = 0.f;
    for (int t = 0; t < T; ++t) {
      float gi[UH], gf[UH], gg[UH], go[UH], hv[UH];
      if (active) {
        mbar_wait(&acc_full[t & 1], (t >> 1) & 1);
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[t * 8 + 3] = clock64();
        tc_fence_after();
        uint32_t z[4][UH];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t & 1) * NG + half * UH);
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld_32x16(ta + g * U, z[g]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[t & 1]);
#pragma unroll
        for (int j = 0; j < UH; ++j) {
          const int n = half * UH + j;
          gi[j] = sigm_fast(__uint_as_float(z[0][j]) + s_bias[n]);
          gf[j] = sigm_fast(__uint_as_float(z[1][j]) + s_bias[U + n]);
          gg[j] = 1.7159f * tanh_fast(0.6666f * (__uint_as_float(z[2][j]) + s_bias[2 * U + n]));
          go[j] = sigm_fast(__uint_as_float(z[3][j]) + s_bias[3 * U + n]);
          cst[j] = fmaf(gi[j], gg[j], gf[j] * cst[j]);
          hv[j] = go[j] * (1.7159f * tanh_fast(0.6666f * cst[j]));
        }
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[t * 8 + 4] = clock64();
        // critical path first: h_t into the operand of step t + 1 (also the wgrad operand)
        if (valid) {
          uint4* np = reinterpret_cast<uint4*>(p.xh + ((size_t)(t + 1) * B + b) * (I + H) + I + j0);
          np[0] = pack8(hv);
          np[1] = pack8(hv + 8);
        }
      }
      if (t < T - 1) {
        fence_proxy_async_all();
        if (t > 0) cl_wait();                         // close phase t - 1 before arriving on phase t
        cl_arrive();
      }
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[t * 8 + 5] = clock64();
      if (active) {
        // off the critical path: state for the backward pass, one 512-byte piece per warp store
        uint4* sp = p.state + state_idx(t, n_ctas, blockIdx.x, ew, 0, lane);
        sp[0 * 32] = pack8(gi); sp[1 * 32] = pack8(gi + 8);
        sp[2 * 32] = pack8(gf); sp[3 * 32] = pack8(gf + 8);
        sp[4 * 32] = pack8(gg); sp[5 * 32] = pack8(gg + 8);
        sp[6 * 32] = pack8(go); sp[7 * 32] = pack8(go + 8);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          sp[(8 + v) * 32] = make_uint4(__float_as_uint(cst[4 * v]), __float_as_uint(cst[4 * v + 1]),
                                        __float_as_uint(cst[4 * v + 2]), __float_as_uint(cst[4 * v + 3]));
      }
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 128) p.dbg[t * 8 + 6] = clock64();
    }
    if (T > 1) cl_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * NG); }
}

struct BwdP {
  int T, B, H, R;
  const __nv_bfloat16* err;   // dL/dh from above: [B][T][H] (seq) or [B][H] (last step only)
  long long lde;              // row pitch of err (elements)
  int seq;
  const uint4* state;         // private layout written by lstm_fwd_k
  uint4* part;                // [2][n_ctas][8 warps][H / 16 vectors][32 lanes]: bf16 partial sums
  __nv_bfloat16* dz;          // [T][B][4H]
};

__global__ void __launch_bounds__(THREADS, 1)
lstm_bwd_k(const __grid_constant__ CUtensorMap t_whp, BwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, T = p.T, B = p.B, R = p.R;
  const int csize = H / U;
  uint8_t* sB = sm;                                   // 2 k-blocks x [H rows][128 B]: W_h rows of own gates
  uint8_t* sA = sB + (size_t)2 * H * 128;             // 2 k-blocks x [128 rows][128 B]: own dz_t
  __shared__ __align__(8) uint64_t b_full, a_ready, acc_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (int)cl_rank();
  const int row0 = (blockIdx.x / csize) * R;
  const uint32_t tcols = H <= 64 ? 64u : (H <= 128 ? 128u : 256u);
  const int NV = H / 16;                              // partial vectors (8 bf16) per thread: H / 2 columns
  const int n_ctas = gridDim.x;
  const size_t part_buf = (size_t)n_ctas * 8 * NV * 32;

  if (threadIdx.x == 0) {
    mbar_init(&b_full, 1);
    mbar_init(&a_ready, (uint32_t)(2 * (R / 32)));
    mbar_init(&acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&t_whp);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, tcols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(BM, H, 0, 0);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&b_full, (uint32_t)(2 * H * 128));
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sB + (size_t)kb * H * 128, &t_whp, &b_full, c * NG + kb * 64, 0);
    }
    for (int k = 0; k < T - 1; ++k) { cl_arrive(); cl_wait(); }
  } else if (warp == 1) {
    if (lane == 0) mbar_wait(&b_full, 0);
    __syncwarp();
    for (int k = 0; k < T - 1; ++k) {
      if (lane == 0) {
        mbar_wait(&a_ready, k & 1);                    // dz_s of the own units is in sA
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t sa = smem_u32(sA + (size_t)kb * BM * 128);
          const uint32_t sb = smem_u32(sB + (size_t)kb * H * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_f16(tmem_base, make_smem_desc(sa + kk * 32, 16, 1024), make_smem_desc(sb + kk * 32, 16, 1024),
                    idesc, (kb > 0 || kk > 0) ? 1u : 0u);
        }
        mma_commit(&acc_full);
      }
      __syncwarp();
      cl_arrive(); cl_wait();
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3, half = ew >> 2;
    const int r = q * 32 + lane;
    const int b = row0 + r;
    const bool active = q * 32 < R;
    const bool valid = active && b < B;
    const int j0 = c * U + half * UH;
    // where the partial sums of this thread's 16 units sit inside a source CTA's lane-major buffer
    const int n0 = j0;                                // first column (unit) wanted
    const int src_half = n0 / (H / 2);
    const int src_v = (n0 % (H / 2)) / 8;
    const int src_ew = src_half * 4 + ((q + 2) & 3);
    const int cta0 = (blockIdx.x / csize) * csize;    // first CTA of this cluster
    float dcs[UH];
#pragma unroll
    for (int j = 0; j < UH; ++j) dcs[j] = 0.f;
    for (int s = T - 1; s >= 0; --s) {
      const int k = T - 1 - s;
      uint4 pk[4][2];
      if (active) {
        // -------- A phase: dh_s -> dz_s of the own units
        float gi[UH], gf[UH], gg[UH], go[UH], cc[UH], cp[UH], dh[UH];
        {
          const uint4* sp = p.state + state_idx(s, n_ctas, blockIdx.x, ew, 0, lane);
          const uint4* pp = p.state + state_idx(s > 0 ? s - 1 : 0, n_ctas, blockIdx.x, ew, 8, lane);
          const uint4 a0 = sp[0], a1 = sp[32], a2 = sp[64], a3 = sp[96], a4 = sp[128], a5 = sp[160];
          const uint4 a6 = sp[192], a7 = sp[224];
          unpack8(a0, gi); unpack8(a1, gi + 8); unpack8(a2, gf); unpack8(a3, gf + 8);
          unpack8(a4, gg); unpack8(a5, gg + 8); unpack8(a6, go); unpack8(a7, go + 8);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 cv = sp[(8 + v) * 32];
            cc[4 * v] = __uint_as_float(cv.x); cc[4 * v + 1] = __uint_as_float(cv.y);
            cc[4 * v + 2] = __uint_as_float(cv.z); cc[4 * v + 3] = __uint_as_float(cv.w);
            const uint4 pv = s > 0 ? pp[v * 32] : make_uint4(0u, 0u, 0u, 0u);
            cp[4 * v] = __uint_as_float(pv.x); cp[4 * v + 1] = __uint_as_float(pv.y);
            cp[4 * v + 2] = __uint_as_float(pv.z); cp[4 * v + 3] = __uint_as_float(pv.w);
          }
        }
#pragma unroll
        for (int j = 0; j < UH; ++j) dh[j] = 0.f;
        if (p.seq || s == T - 1) {
          const uint4* e4 = reinterpret_cast<const uint4*>(p.err + (size_t)(valid ? b : 0) * p.lde +
                                                           (p.seq ? (size_t)s * H : 0) + j0);
          unpack8(e4[0], dh);
          unpack8(e4[1], dh + 8);
        }
        if (k > 0) {
          cl_wait();                                   // phase k - 1: every CTA's partial sums are out
          const uint4* pb = p.part + (size_t)((k - 1) & 1) * part_buf;
          for (int src = 0; src < csize; ++src) {
            const uint4* pv = pb + ((size_t)((cta0 + src) * 8 + src_ew) * NV + src_v) * 32 + lane;
            float t0[8], t1[8];
            unpack8(__ldcg(pv), t0);
            unpack8(__ldcg(pv + 32), t1);
#pragma unroll
            for (int j = 0; j < 8; ++j) { dh[j] += t0[j]; dh[8 + j] += t1[j]; }
          }
        }
        float zz[4][UH];
#pragma unroll
        for (int j = 0; j < UH; ++j) {
          const float tc = 1.7159f * tanh_fast(0.6666f * cc[j]);
          const float dtc = fmaf(tc * tc, -0.388484177f, 1.14381894f);        // d(A tanh(B c)) / dc
          const float dc = fmaf(dh[j] * go[j], dtc, dcs[j]);
          zz[0][j] = dc * gg[j] * gi[j] * (1.f - gi[j]);
          zz[1][j] = dc * cp[j] * gf[j] * (1.f - gf[j]);
          zz[2][j] = dc * gi[j] * fmaf(gg[j] * gg[j], -0.388484177f, 1.14381894f);
          zz[3][j] = dh[j] * tc * go[j] * (1.f - go[j]);
          dcs[j] = dc * gf[j];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) { pk[g][0] = pack8(zz[g]); pk[g][1] = pack8(zz[g] + 8); }
        if (s > 0) {
          // own A operand: k index = gate * 32 + local unit (the order W_h^T was permuted to)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint8_t* blk = sA + (size_t)(g >> 1) * BM * 128;
            const int ch = (g & 1) * 4 + half * 2;
            *reinterpret_cast<uint4*>(blk + sw128(r, ch)) = pk[g][0];
            *reinterpret_cast<uint4*>(blk + sw128(r, ch + 1)) = pk[g][1];
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready);
        }
        if (valid) {
          __nv_bfloat16* dr = p.dz + ((size_t)s * B + b) * 4 * H + j0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            reinterpret_cast<uint4*>(dr + (size_t)g * H)[0] = pk[g][0];
            reinterpret_cast<uint4*>(dr + (size_t)g * H)[1] = pk[g][1];
          }
        }
      } else if (k > 0) {
        cl_wait();
      }
      if (s == 0) break;
      // -------- B phase: P = dz_s[:, own] . W_h[own, :] -> lane-major bf16 pieces for their owners
      if (active) {
        mbar_wait(&acc_full, k & 1);
        tc_fence_after();
        uint4* pw = p.part + (size_t)(k & 1) * part_buf + ((size_t)(blockIdx.x * 8 + ew) * NV) * 32 + lane;
        for (int v0 = 0; v0 < NV; v0 += 4) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (H / 2) + v0 * 8), v);
          tmem_ld_wait();
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[8 * w4]), __uint_as_float(v[8 * w4 + 1]));
            o.y = pack_bf16(__uint_as_float(v[8 * w4 + 2]), __uint_as_float(v[8 * w4 + 3]));
            o.z = pack_bf16(__uint_as_float(v[8 * w4 + 4]), __uint_as_float(v[8 * w4 + 5]));
            o.w = pack_bf16(__uint_as_float(v[8 * w4 + 6]), __uint_as_float(v[8 * w4 + 7]));
            pw[(size_t)(v0 + w4) * 32] = o;
          }
        }
        tc_fence_before();
      }
      cl_arrive();                                     // phase k: this thread's partial sums are out
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tcols); }
}

// W_h^T in the k order of the backward kernel: dst[n][c * 128 + g * 32 + j] = w[g * H + c * 32 + j][col0 + n]
__global__ void permute_wh_k(const __nv_bfloat16* __restrict__ w, long long ldw, int col0, int H,
                             __nv_bfloat16* __restrict__ dst) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;          // k0: permuted k index (multiple of 32)
  const int c = k0 / NG, g = (k0 % NG) / U;
  const int row = g * H + c * U;                                  // first source row of this k run
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    tile[i][threadIdx.x] = w[(size_t)(row + i) * ldw + col0 + n0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    dst[(size_t)(n0 + i) * 4 * H + k0 + threadIdx.x] = tile[threadIdx.x][i];
}

// private state layout -> the unit's public gates [T][B][4H] / cells [T][B][H] (per-step fallback)
__global__ void lstm_unpack_state_k(const uint4* __restrict__ state, float* __restrict__ gates,
                                    float* __restrict__ cells, int T, int B, int H, int R, int n_ctas) {
  const size_t total = (size_t)T * n_ctas * 8 * SVEC * 32;
  const int csize = H / U;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int lane = (int)(i % 32);
    const int v = (int)((i / 32) % SVEC);
    const int ew = (int)((i / (32 * SVEC)) % 8);
    const int cta = (int)((i / (32 * SVEC * 8)) % n_ctas);
    const int t = (int)(i / ((size_t)32 * SVEC * 8 * n_ctas));
    const int q = (ew + 2) & 3, half = ew >> 2;
    const int r = q * 32 + lane;
    const int b = (cta / csize) * R + r;
    if (r >= R || b >= B) continue;
    const int j = (cta % csize) * U + half * UH;
    const uint4 val = state[i];
    if (v < 8) {
      float f[8];
      unpack8(val, f);
      float* dst = gates + ((size_t)t * B + b) * 4 * H + (size_t)(v >> 1) * H + j + 8 * (v & 1);
      reinterpret_cast<float4*>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(f[4], f[5], f[6], f[7]);
    } else {
      float* dst = cells + ((size_t)t * B + b) * H + j + 4 * (v - 8);
      *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(val.x), __uint_as_float(val.y),
                                                    __uint_as_float(val.z), __uint_as_float(val.w));
    }
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
static int make_map(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <typename... KArgs, typename... Args>
static int launch_cluster(void (*kernel)(KArgs...), int blocks, int csize, size_t smem, cudaStream_t st,
                          Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace lp

static long long g_lstm_persist_launches = 0;
long long lstm_persist_launches() { return g_lstm_persist_launches; }

static bool lstm_shape_ok(int T, int B, int I, int H) {
  const int cs = H / lp::U;
  return I >= 64 && I % 64 == 0 && H >= 64 && H % 64 == 0 && cs <= 8 && T >= 1 && B >= 1;
}
// batch rows per cluster: the smallest of 128 / 64 / 32 that keeps the grid within ~one wave
// (fewer rows per CTA = less epilogue math, TMA bytes and state traffic per step and SM)
int lstm_persist_rows(int B, int H) {
  const int cs = H / lp::U;
  const char* e = getenv("ZNICZ_LSTM_ROWS");
  if (e && (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 128)) return atoi(e);
  for (int r = 32; r < 128; r *= 2)
    if (((B + r - 1) / r) * cs <= 128) return r;
  return 128;
}
static int lstm_ctas(int B, int H) {
  const int R = lstm_persist_rows(B, H);
  return ((B + R - 1) / R) * (H / lp::U);
}
// 16-byte elements of the private state buffer the forward kernel needs (0: shape not supported)
long long lstm_persist_state_elems(int T, int B, int I, int H) {
  if (!lstm_shape_ok(T, B, I, H)) return 0;
  return (long long)T * lstm_ctas(B, H) * 8 * lp::SVEC * 32;
}
// 16-byte elements of the backward kernel's partial-sum exchange buffer
long long lstm_persist_part_elems(int B, int H) { return 2LL * lstm_ctas(B, H) * 8 * (H / 16) * 32; }

// 0 = launched; < 0 = shape outside the kernel's range (the caller keeps the per-step path)
int launch_lstm_fwd_persist(void* xh, const void* w_lp, long long ldw, const float* bias, void* state,
                            int T, int B, int I, int H, long long* dbg, cudaStream_t st) {
  using namespace lp;
  if (!lstm_shape_ok(T, B, I, H)) return -3;
  const int KX = I / 64, KH = H / 64;
  const size_t smem = (size_t)(KX + KH) * NG * 128 + (size_t)2 * KX * BM * 128 + (size_t)KH * BM * 128 + 1024;
  if (smem > 227 * 1024 - 2048) return -4;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(lstm_fwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048) != cudaSuccess)
      return -5;
    attr_set = true;
  }
  CUtensorMap t_xh, t_w;
  const int R = lstm_persist_rows(B, H);
  int r = make_map(&t_xh, xh, I + H, (long long)(T + 1) * B, I + H, R);
  if (r) return -6;
  r = make_map(&t_w, w_lp, I + H, 4LL * H, ldw, U);
  if (r) return -6;
  FwdP p{T, B, I, H, R, bias, (uint4*)state, (__nv_bfloat16*)xh, dbg};
  const int csize = H / U;
  const int clusters = (B + R - 1) / R;
  r = launch_cluster(lstm_fwd_k, clusters * csize, csize, smem, st, t_xh, t_w, p);
  if (r == 0) ++g_lstm_persist_launches;
  return r == 0 ? 0 : -7;
}

int launch_lstm_bwd_persist(const void* err, long long lde, int seq, const void* state, void* part, void* dz,
                            const void* w_lp, long long ldw, void* whp, int T, int B, int I, int H,
                            cudaStream_t st) {
  using namespace lp;
  if (!lstm_shape_ok(T, B, I, H)) return -3;
  const int csize = H / U;
  const size_t smem = (size_t)2 * H * 128 + (size_t)2 * BM * 128 + 1024;
  if (smem > 227 * 1024 - 2048) return -4;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(lstm_bwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048) != cudaSuccess)
      return -5;
    attr_set = true;
  }
  // W_h^T [H][4H], columns permuted to (CTA, gate, unit) order, from the bf16 weight shadow
  {
    dim3 grid(H / 32, 4 * H / 32), block(32, 8);
    permute_wh_k<<<grid, block, 0, st>>>((const __nv_bfloat16*)w_lp, ldw, I, H, (__nv_bfloat16*)whp);
  }
  CUtensorMap t_whp;
  int r = make_map(&t_whp, whp, 4LL * H, H, 4LL * H, H);
  if (r) return -6;
  const int R = lstm_persist_rows(B, H);
  BwdP p{T, B, H, R, (const __nv_bfloat16*)err, lde, seq, (const uint4*)state, (uint4*)part, (__nv_bfloat16*)dz};
  const int clusters = (B + R - 1) / R;
  r = launch_cluster(lstm_bwd_k, clusters * csize, csize, smem, st, t_whp, p);
  if (r == 0) ++g_lstm_persist_launches;
  return r == 0 ? 0 : -7;
}

void launch_lstm_unpack_state(const void* state, float* gates, float* cells, int T, int B, int H,
                              cudaStream_t st) {
  using namespace lp;
  lstm_unpack_state_k<<<296, 256, 0, st>>>((const uint4*)state, gates, cells, T, B, H, lstm_persist_rows(B, H),
                                           lstm_ctas(B, H));
}

}  // namespace zn
