// CPU-only build of the forward runtime (Android / hosts without the CUDA toolkit): the two
// CUDA entry points of Engine resolve to "not available", everything else is engine.cc.
// Selected with -DZNICZ_WITH_CUDA=OFF (CMake) — the counterpart of the reference's ndk-build
// target (/root/reference/libZnicz/android/Android.mk.in).
#include "znicz_native.h"

#include <stdexcept>

namespace znicz {

bool Engine::cuda_available() { return false; }
long long Engine::cuda_tensor_core_launches() const { return 0; }

std::vector<float> Engine::run_cuda(const float*, const Shape4&) {
  throw std::runtime_error("znicz_native was built without CUDA support");
}

}  // namespace znicz
