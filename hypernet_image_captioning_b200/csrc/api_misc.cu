// Library-wide bookkeeping exported through the C-ABI.
#include "common.cuh"

unsigned long long caphn_launch_counter = 0;

extern "C" {

// Number of CUDA kernels launched by this library since it was loaded (host-side counter, not thread safe).
int caphn_launch_count(unsigned long long* out) {
    if (!out) return CAPHN_EINVAL;
    *out = caphn_launch_counter;
    return CAPHN_OK;
}

// Compile-time target of the library, for smoke tests: returns 100 for sm_100a builds.
int caphn_build_arch(int* out) {
    if (!out) return CAPHN_EINVAL;
    *out = 100;
    return CAPHN_OK;
}

}  // extern "C"
