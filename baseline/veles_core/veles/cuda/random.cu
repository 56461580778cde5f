// Core include of the (absent) Veles platform, written for the reference-arm shim:
// xorshift128+ step used by the reference's dropout kernel (cuda/dropout.cu:16).
#ifndef _VELES_RANDOM_CU_
#define _VELES_RANDOM_CU_

__device__ __forceinline__ void xorshift128plus(ulonglong2 &state, ulong &output) {
  ulong x = state.x;
  ulong const y = state.y;
  state.x = y;
  x ^= x << 23;
  state.y = x ^ y ^ (x >> 17) ^ (y >> 26);
  output = state.y + y;
}

#endif  // _VELES_RANDOM_CU_
