// Error plumbing shared by the translation units of libb200rime.so
#pragma once
#include <cuda_runtime.h>

namespace b200rime {
// records msg (thread-local) and returns a nonzero status
int set_error(const char* msg);
// cudaGetLastError() after a launch; 0 if clean
int check_launch(const char* what);
// per-call-site "first use on the current device" latch: function attributes such as the
// dynamic shared-memory limit are per device, so a process that drives several GPUs must set
// them once on each
struct DeviceOnce {
    unsigned long long seen[2] = {0ull, 0ull};
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 128) return true;
        unsigned long long& w = seen[dev >> 6];
        const unsigned long long bit = 1ull << (dev & 63);
        if (w & bit) return false;
        w |= bit;
        return true;
    }
};
}  // namespace b200rime
