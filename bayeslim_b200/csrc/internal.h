// Error plumbing shared by the translation units of libb200rime.so
#pragma once
#include <cuda_runtime.h>

namespace b200rime {
// records msg (thread-local) and returns a nonzero status
int set_error(const char* msg);
// cudaGetLastError() after a launch; 0 if clean
int check_launch(const char* what);
}  // namespace b200rime
