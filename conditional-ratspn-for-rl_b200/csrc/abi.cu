// Library identification for the C ABI.
#include <stdio.h>

#include "common.cuh"

#ifndef SPN_B200_VERSION
#define SPN_B200_VERSION "0.1.0"
#endif

extern "C" int spn_version(char* buf, int buflen) {
    if (buf && buflen > 0) snprintf(buf, (size_t)buflen, "libspn_b200 %s (sm_100a, CUDA %d)", SPN_B200_VERSION, CUDART_VERSION);
    return 100;
}
