// Kohonen self-organising map kernels — the reference has only OpenCL kernels for this op
// (/root/reference/ocl/kohonen.cl:19,59,139,161,182: calculate_distances, calculate_argmin,
// set_total, compute_gravity, apply_gradient); this is a CUDA design for sm_100a:
//   som_winners_k : one CTA per sample, warp-shuffle argmin of ||w_n - x||^2, winners histogram
//   som_update_k  : one thread per (neuron, feature); sums gravity * (x - w) over the batch and
//                   applies the step (batch semantics identical to the numpy oracle)
#include "common.cuh"

namespace zn {

__global__ void som_winners_k(const float* __restrict__ x, const float* __restrict__ w,
                              int* __restrict__ argmins, int* __restrict__ winners, int neurons,
                              int len, int count_winners) {
  pdl_entry();
  const int s = blockIdx.x;
  const float* xs = x + (size_t)s * len;
  float best = 3.0e38f; int bi = 0x7fffffff;
  for (int n = threadIdx.x; n < neurons; n += blockDim.x) {
    const float* wn = w + (size_t)n * len;
    float d = 0.f;
    for (int k = 0; k < len; ++k) { float t = wn[k] - xs[k]; d = fmaf(t, t, d); }
    if (d < best || (d == best && n < bi)) { best = d; bi = n; }
  }
  __shared__ float sb[32]; __shared__ int si[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { sb[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? sb[lane] : 3.0e38f; bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      argmins[s] = bi;
      if (count_winners) atomicAdd(winners + bi, 1);
    }
  }
}

__global__ void som_update_k(const float* __restrict__ x, float* __restrict__ w,
                             const float* __restrict__ coords, const int* __restrict__ argmins,
                             int batch, int neurons, int len, float sigma, float gmult) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= neurons * len) return;
  const int n = i / len, k = i - n * len;
  const float cx = coords[2 * n], cy = coords[2 * n + 1];
  const float inv = -1.f / (2.f * sigma * sigma);
  const float wv = w[i];
  float g = 0.f;
  for (int s = 0; s < batch; ++s) {
    const int win = argmins[s];
    const float dx = cx - coords[2 * win], dy = cy - coords[2 * win + 1];
    g += __expf((dx * dx + dy * dy) * inv) * (x[(size_t)s * len + k] - wv);
  }
  w[i] = wv + g * gmult;
}

void launch_som_winners(const float* x, const float* w, int* argmins, int* winners, int batch,
                        int neurons, int len, int count_winners, cudaStream_t st) {
  int threads = neurons >= 256 ? 256 : ((neurons + 31) / 32) * 32;
  launch_k(som_winners_k, batch, threads, 0, st, x, w, argmins, winners, neurons, len, count_winners);
}
void launch_som_update(const float* x, float* w, const float* coords, const int* argmins, int batch,
                       int neurons, int len, float sigma, float gmult, cudaStream_t st) {
  int total = neurons * len;
  launch_k(som_update_k, (total + 255) / 256, 256, 0, st, x, w, coords, argmins, batch, neurons, len, sigma, gmult);
}

}  // namespace zn
