// Library identity and error strings of the codd_b200 C ABI.
#include "common.cuh"

#define CODD_B200_VERSION 100  // major*10000 + minor*100 + patch -> 0.1.0

extern "C" int codd_version(void) { return CODD_B200_VERSION; }

extern "C" const char* codd_error_string(int code) {
    if (code == 0) return "success";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
        case CODD_E_BADARG: return "codd: null pointer or non-positive dimension";
        case CODD_E_SHAPE: return "codd: dimensions violate a documented constraint";
        case CODD_E_UNSUPPORTED: return "codd: kernel geometry not instantiated";
        case CODD_E_ALIGN: return "codd: pointer or stride not 16-byte aligned";
        default: return "codd: unknown error";
    }
}
