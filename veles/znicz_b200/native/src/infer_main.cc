// znicz_infer: run a package on a raw float32 input file.
//   znicz_infer <package.zip|dir> <input.f32> <n> <h> <w> <c> [--cuda]
#include "znicz_native.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

int main(int argc, char** argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: %s package input.f32 n h w c [--cuda]\n", argv[0]); return 2; }
  try {
    znicz::Engine eng(argv[1]);
    znicz::Shape4 in{std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6])};
    std::vector<float> x((size_t)in.size());
    std::ifstream f(argv[2], std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(x.data()), x.size() * 4)) { std::fprintf(stderr, "short input\n"); return 3; }
    bool cuda = argc > 7 && !std::strcmp(argv[7], "--cuda");
    auto y = cuda ? eng.run_cuda(x.data(), in) : eng.run_cpu(x.data(), in);
    znicz::Shape4 o = eng.infer(in);
    std::printf("workflow %s units %zu output %dx%dx%dx%d\n", eng.workflow_name().c_str(), eng.num_units(), o.n, o.h, o.w, o.c);
    for (size_t i = 0; i < y.size() && i < 32; ++i) std::printf("%g ", y[i]);
    std::printf("\n");
  } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
