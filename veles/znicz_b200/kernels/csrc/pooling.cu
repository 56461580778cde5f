// Pooling forward/backward + depooling for NHWC tensors (ceil-mode partial windows).
// Parity: /root/reference/cuda/pooling.cu (max :22, avg :77, stochastic :151, pool+depool
// :248), gradient_descent_pooling.cu (:24,:56), depooling.cu:4. One thread per (pixel, channel)
// with channels fastest => coalesced. Backward kernels are *gather* formulations (each input
// element inspects the windows that cover it): no memset, no atomics, deterministic.
#include "common.cuh"

namespace zn {

enum PoolMode { POOL_MAX = 0, POOL_MAXABS = 1, POOL_AVG = 2, POOL_STOCH = 3, POOL_STOCH_ABS = 4,
                POOL_STOCH_DEPOOL = 5, POOL_STOCH_ABS_DEPOOL = 6 };

struct PoolGeom { int N, H, W, C, OH, OW, KY, KX, SY, SX; };

template <typename T>
__global__ void pool_forward_k(const T* __restrict__ in, T* __restrict__ out, int* __restrict__ offs,
                               PoolGeom g, int mode, const int* __restrict__ rng) {
  long long total = (long long)g.N * g.OH * g.OW * g.C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % g.C); long long t = i / g.C;
  int ox = (int)(t % g.OW); t /= g.OW; int oy = (int)(t % g.OH); int n = (int)(t / g.OH);
  int y1 = oy * g.SY, x1 = ox * g.SX;
  int y2 = min(y1 + g.KY, g.H), x2 = min(x1 + g.KX, g.W);
  const T* base = in + (size_t)n * g.H * g.W * g.C + c;
  if (mode == POOL_AVG) {
    float s = 0.f;
    for (int y = y1; y < y2; ++y)
      for (int x = x1; x < x2; ++x) s += ldf(base + ((size_t)y * g.W + x) * g.C);
    stf(out + i, s / (float)((y2 - y1) * (x2 - x1)));
    return;
  }
  bool use_abs = (mode == POOL_MAXABS || mode == POOL_STOCH_ABS || mode == POOL_STOCH_ABS_DEPOOL);
  int best_off = 0; float best_v = 0.f;
  if (mode == POOL_MAX || mode == POOL_MAXABS) {
    float best_key = -3.0e38f;
    for (int y = y1; y < y2; ++y)
      for (int x = x1; x < x2; ++x) {
        size_t o = ((size_t)y * g.W + x) * g.C;
        float v = ldf(base + o);
        float key = use_abs ? fabsf(v) : v;
        if (key > best_key) { best_key = key; best_v = v; best_off = (int)(o); }
      }
  } else {
    // stochastic: P(index) ~ max(v,0) (or |v|); 16-bit random per output element
    uint32_t rnd = hash_u32((uint32_t)rng[0], (uint32_t)rng[1], (uint64_t)i) >> 16;
    float vsum = 0.f; int cnt = (y2 - y1) * (x2 - x1);
    for (int y = y1; y < y2; ++y)
      for (int x = x1; x < x2; ++x) {
        float v = ldf(base + ((size_t)y * g.W + x) * g.C);
        vsum += use_abs ? fabsf(v) : fmaxf(v, 0.f);
      }
    bool found = false;
    if (vsum == 0.f) {
      int pick = (int)(((unsigned long long)rnd * (unsigned)cnt) >> 16);
      int k = 0;
      for (int y = y1; y < y2 && !found; ++y)
        for (int x = x1; x < x2; ++x, ++k)
          if (k == pick) { size_t o = ((size_t)y * g.W + x) * g.C; best_off = (int)o; best_v = ldf(base + o); found = true; break; }
    } else {
      float pos = (float)rnd * vsum / 65536.f, acc = 0.f;
      for (int y = y1; y < y2 && !found; ++y)
        for (int x = x1; x < x2; ++x) {
          size_t o = ((size_t)y * g.W + x) * g.C;
          float v = ldf(base + o);
          acc += use_abs ? fabsf(v) : fmaxf(v, 0.f);
          if (pos <= acc) { best_off = (int)o; best_v = v; found = true; break; }
        }
      if (!found) {  // rounding at the very end of the window
        size_t o = ((size_t)(y2 - 1) * g.W + (x2 - 1)) * g.C; best_off = (int)o; best_v = ldf(base + o);
      }
    }
  }
  int flat = (int)((size_t)n * g.H * g.W * g.C + c) + best_off;
  offs[i] = flat;
  if (mode == POOL_STOCH_DEPOOL || mode == POOL_STOCH_ABS_DEPOOL) {
    // in-place pool+depool (non-overlapping windows): keep the winner, zero the rest
    T* wbase = const_cast<T*>(base);
    for (int y = y1; y < y2; ++y)
      for (int x = x1; x < x2; ++x) {
        size_t o = ((size_t)y * g.W + x) * g.C;
        if ((int)o != best_off) stf(wbase + o, 0.f);
      }
  } else {
    stf(out + i, best_v);
  }
}

// err_in[n,y,x,c] = sum over windows covering (y,x) of err_out[window] * [offs[window]==self]
template <typename T>
__global__ void pool_backward_max_k(const T* __restrict__ err_out, const int* __restrict__ offs,
                                    T* __restrict__ err_in, PoolGeom g) {
  long long total = (long long)g.N * g.H * g.W * g.C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % g.C); long long t = i / g.C;
  int x = (int)(t % g.W); t /= g.W; int y = (int)(t % g.H); int n = (int)(t / g.H);
  int oy_hi = min(y / g.SY, g.OH - 1), ox_hi = min(x / g.SX, g.OW - 1);
  int oy_lo = max(0, (y - g.KY + g.SY) / g.SY), ox_lo = max(0, (x - g.KX + g.SX) / g.SX);
  if (y - g.KY + 1 <= 0) oy_lo = 0;
  if (x - g.KX + 1 <= 0) ox_lo = 0;
  float s = 0.f;
  for (int oy = oy_lo; oy <= oy_hi; ++oy)
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      size_t o = (((size_t)n * g.OH + oy) * g.OW + ox) * g.C + c;
      if (offs[o] == (int)i) s += ldf(err_out + o);
    }
  stf(err_in + i, s);
}

template <typename T>
__global__ void pool_backward_avg_k(const T* __restrict__ err_out, T* __restrict__ err_in, PoolGeom g) {
  long long total = (long long)g.N * g.H * g.W * g.C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % g.C); long long t = i / g.C;
  int x = (int)(t % g.W); t /= g.W; int y = (int)(t % g.H); int n = (int)(t / g.H);
  int oy_hi = min(y / g.SY, g.OH - 1), ox_hi = min(x / g.SX, g.OW - 1);
  int oy_lo = (y - g.KY + 1 <= 0) ? 0 : (y - g.KY + g.SY) / g.SY;
  int ox_lo = (x - g.KX + 1 <= 0) ? 0 : (x - g.KX + g.SX) / g.SX;
  float s = 0.f;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    int hy = min(oy * g.SY + g.KY, g.H) - oy * g.SY;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      int hx = min(ox * g.SX + g.KX, g.W) - ox * g.SX;
      s += ldf(err_out + (((size_t)n * g.OH + oy) * g.OW + ox) * g.C + c) / (float)(hy * hx);
    }
  }
  stf(err_in + i, s);
}

void launch_pool_forward(const void* in, void* out, int* offs, int N, int H, int W, int C, int OH, int OW,
                         int KY, int KX, int SY, int SX, int mode, const int* rng, bool bf16,
                         cudaStream_t st) {
  PoolGeom g{N, H, W, C, OH, OW, KY, KX, SY, SX};
  long long total = (long long)N * OH * OW * C;
  int grid = cdiv(total, 256);
  if (bf16) pool_forward_k<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, offs, g, mode, rng);
  else pool_forward_k<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, offs, g, mode, rng);
}
void launch_pool_backward(const void* err_out, const int* offs, void* err_in, int N, int H, int W, int C,
                          int OH, int OW, int KY, int KX, int SY, int SX, int is_avg, bool bf16,
                          cudaStream_t st) {
  PoolGeom g{N, H, W, C, OH, OW, KY, KX, SY, SX};
  long long total = (long long)N * H * W * C;
  int grid = cdiv(total, 256);
  if (is_avg) {
    if (bf16) pool_backward_avg_k<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)err_out, (__nv_bfloat16*)err_in, g);
    else pool_backward_avg_k<float><<<grid, 256, 0, st>>>((const float*)err_out, (float*)err_in, g);
  } else {
    if (bf16) pool_backward_max_k<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)err_out, offs, (__nv_bfloat16*)err_in, g);
    else pool_backward_max_k<float><<<grid, 256, 0, st>>>((const float*)err_out, offs, (float*)err_in, g);
  }
}

// ------------------------------------------------------------------------------------ LRN
// y_i = x_i * (k + alpha * sum_{j in win(i)} x_j^2)^-beta  (/root/reference/cuda/normalization.cu)
template <typename T>
__global__ void lrn_forward_k(const T* __restrict__ x, T* __restrict__ y, long long pixels, int C,
                              int half, float alpha, float beta, float k) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * C) return;
  int c = (int)(i % C);
  const T* px = x + (i - c);
  int lo = max(0, c - half), hi = min(C - 1, c + half);
  float s = 0.f;
  for (int j = lo; j <= hi; ++j) { float v = ldf(px + j); s += v * v; }
  s = k + alpha * s;
  stf(y + i, ldf(px + c) * __powf(s, -beta));
}
// eh_i = ey_i s_i^-b - 2ab x_i sum_{j in win(i)} ey_j x_j s_j^(-b-1)
template <typename T>
__global__ void lrn_backward_k(const T* __restrict__ ey, const T* __restrict__ x, T* __restrict__ eh,
                               long long pixels, int C, int half, float alpha, float beta, float k) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * C) return;
  int c = (int)(i % C);
  const T* px = x + (i - c);
  const T* pe = ey + (i - c);
  int lo = max(0, c - half), hi = min(C - 1, c + half);
  // sliding subsums for the (up to) 2*half+1 neighbours: recompute (C and n are small)
  float xi = ldf(px + c);
  float acc = 0.f, si = 0.f;
  for (int j = lo; j <= hi; ++j) {
    int jlo = max(0, j - half), jhi = min(C - 1, j + half);
    float s = 0.f;
    for (int q = jlo; q <= jhi; ++q) { float v = ldf(px + q); s += v * v; }
    s = k + alpha * s;
    if (j == c) si = s;
    acc += ldf(pe + j) * ldf(px + j) * __powf(s, -beta - 1.f);
  }
  stf(eh + i, ldf(pe + c) * __powf(si, -beta) - 2.f * alpha * beta * xi * acc);
}

void launch_lrn_forward(const void* x, void* y, long long pixels, int C, int n, float alpha, float beta,
                        float k, bool bf16, cudaStream_t st) {
  int grid = cdiv(pixels * C, 256);
  if (bf16) lrn_forward_k<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, pixels, C, n / 2, alpha, beta, k);
  else lrn_forward_k<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, pixels, C, n / 2, alpha, beta, k);
}
void launch_lrn_backward(const void* ey, const void* x, void* eh, long long pixels, int C, int n,
                         float alpha, float beta, float k, bool bf16, cudaStream_t st) {
  int grid = cdiv(pixels * C, 256);
  if (bf16) lrn_backward_k<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)ey, (const __nv_bfloat16*)x, (__nv_bfloat16*)eh, pixels, C, n / 2, alpha, beta, k);
  else lrn_backward_k<float><<<grid, 256, 0, st>>>((const float*)ey, (const float*)x, (float*)eh, pixels, C, n / 2, alpha, beta, k);
}

}  // namespace zn
