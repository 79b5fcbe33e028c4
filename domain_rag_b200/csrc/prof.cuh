// Optional per-launch CUDA-event timing of the heavy kernels (GEMM, attention), used by bench.py to
// measure the dominant kernel's average launch duration live, on the launching stream.
#pragma once
#include <cuda_runtime.h>

namespace drag {

enum ProfClass : int { PROF_GEMM = 0, PROF_ATTENTION = 1, PROF_NUM = 2 };

bool prof_enabled();
// Records an event on `st` and returns a slot id (or -1 when disabled); call prof_end with it.
int prof_begin(int cls, double work, cudaStream_t st);
void prof_end(int slot, cudaStream_t st);

}  // namespace drag
