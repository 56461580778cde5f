// sm_100a executor of the native runtime: launches the training engine's own kernels
// (csrc/gemm_simt.cu, pooling.cu, elementwise.cu, softmax_eval.cu) through their C++ launchers —
// no libtorch, no cuBLAS. Weights are uploaded once and stay resident in HBM.
#include "znicz_native.h"

#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

namespace zn {
void launch_gemm_simt(const void*, bool, long long, int, const void*, bool, long long, int, void*, bool,
                      long long, int, int, int, int, const float*, int, float, float, int, long long,
                      cudaStream_t);
void launch_conv_fprop_simt_raw(const void*, bool, const float*, long long, int, const float*, void*, bool,
                                const int*, int, cudaStream_t);
void launch_pool_forward(const void*, void*, int*, int, int, int, int, int, int, int, int, int, int, int,
                         const int*, bool, int, cudaStream_t);
void launch_lrn_forward(const void*, void*, long long, int, int, float, float, float, bool, cudaStream_t);
void launch_act_forward(const void*, void*, long long, int, float, bool, cudaStream_t);
void launch_softmax_rows(const void*, bool, float*, int*, int, int, cudaStream_t);
void launch_crop_nhwc(const void*, void*, int, int, int, int, int, int, int, int, int, bool, cudaStream_t);
}  // namespace zn

namespace znicz {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_)); } while (0)

struct Engine::CudaState {
  std::vector<float*> w, b;
  float* buf[2] = {nullptr, nullptr};
  int* ibuf = nullptr;
  size_t cap = 0, icap = 0;
  cudaStream_t st = nullptr;
  ~CudaState() {
    for (auto p : w) cudaFree(p);
    for (auto p : b) cudaFree(p);
    cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(ibuf);
    if (st) cudaStreamDestroy(st);
  }
};

bool Engine::cuda_available() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

std::vector<float> Engine::run_cuda(const float* input, const Shape4& in) {
  if (!cuda_available()) throw std::runtime_error("no CUDA device available");
  if (!cuda_) {
    cuda_.reset(new CudaState());
    CK(cudaStreamCreate(&cuda_->st));
    for (auto& u : units_) {
      float *w = nullptr, *b = nullptr;
      if (!u.weights.data.empty()) { CK(cudaMalloc(&w, u.weights.data.size() * 4)); CK(cudaMemcpy(w, u.weights.data.data(), u.weights.data.size() * 4, cudaMemcpyHostToDevice)); }
      if (u.include_bias && !u.bias.data.empty()) { CK(cudaMalloc(&b, u.bias.data.size() * 4)); CK(cudaMemcpy(b, u.bias.data.data(), u.bias.data.size() * 4, cudaMemcpyHostToDevice)); }
      cuda_->w.push_back(w); cuda_->b.push_back(b);
    }
  }
  // size the ping-pong activation buffers
  size_t need = (size_t)in.size(); Shape4 s = in; size_t ineed = 0;
  for (auto& u : units_) { s = out_shape(u, s); need = std::max(need, (size_t)s.size()); if (u.kind == "pool" && u.pool_mode != 2) ineed = std::max(ineed, (size_t)s.size()); if (u.softmax) ineed = std::max(ineed, (size_t)s.n); }
  if (need > cuda_->cap) { cudaFree(cuda_->buf[0]); cudaFree(cuda_->buf[1]); CK(cudaMalloc(&cuda_->buf[0], need * 4)); CK(cudaMalloc(&cuda_->buf[1], need * 4)); cuda_->cap = need; }
  if (ineed > cuda_->icap) { cudaFree(cuda_->ibuf); CK(cudaMalloc(&cuda_->ibuf, ineed * 4)); cuda_->icap = ineed; }
  cudaStream_t st = cuda_->st;
  CK(cudaMemcpyAsync(cuda_->buf[0], input, (size_t)in.size() * 4, cudaMemcpyHostToDevice, st));
  int cur = 0; s = in;
  for (size_t i = 0; i < units_.size(); ++i) {
    auto& u = units_[i];
    Shape4 o = out_shape(u, s);
    float* x = cuda_->buf[cur]; float* y = cuda_->buf[cur ^ 1];
    if (u.kind == "all2all") {
      int K = (int)u.weights.shape[1], N = o.c;
      zn::launch_gemm_simt(x, false, K, 0, cuda_->w[i], false, K, 1, y, false, N, 0, s.n, N, K, cuda_->b[i],
                           u.softmax ? 0 : (int)u.act, 1.f, 0.f, 1, 0, st);
      if (u.softmax) { zn::launch_softmax_rows(y, false, x, cuda_->ibuf, s.n, N, st); std::swap(x, y); cur ^= 1; }
    } else if (u.kind == "conv") {
      int g[13] = {s.n, s.h, s.w, s.c, o.h, o.w, o.c, u.ky, u.kx, u.sy, u.sx, u.pad[1], u.pad[0]};
      zn::launch_conv_fprop_simt_raw(x, false, cuda_->w[i], u.weights.shape[1], 0, cuda_->b[i], y, false, g, (int)u.act, st);
    } else if (u.kind == "pool") {
      zn::launch_pool_forward(x, y, cuda_->ibuf, s.n, s.h, s.w, s.c, o.h, o.w, u.ky, u.kx, u.sy, u.sx, u.pool_mode, nullptr, false, 0, st);
    } else if (u.kind == "lrn") {
      zn::launch_lrn_forward(x, y, s.size() / s.c, s.c, u.n, u.alpha, u.beta, u.k, false, st);
    } else if (u.kind == "act") {
      zn::launch_act_forward(x, y, s.size(), u.act_code, u.factor, false, st);
    } else if (u.kind == "cutter") {
      zn::launch_crop_nhwc(x, y, s.n, s.h, s.w, s.c, o.h, o.w, u.pad[1], u.pad[0], 0, false, st);
    } else {
      CK(cudaMemcpyAsync(y, x, (size_t)s.size() * 4, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaGetLastError());
    cur ^= 1; s = o;
  }
  std::vector<float> out((size_t)s.size());
  CK(cudaMemcpyAsync(out.data(), cuda_->buf[cur], out.size() * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return out;
}

}  // namespace znicz
