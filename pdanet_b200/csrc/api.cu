// Library-level entry points of libpdab.so (version / error text).
#include "common.cuh"

extern "C" const char *pdab_version(void) { return "pdab 0.1 (sm_100a)"; }

extern "C" const char *pdab_error_string(int code) {
    if (code == 0) return "success";
    if (code == PDAB_EINVAL) return "pdab: invalid argument (null pointer or negative size)";
    if (code == PDAB_EUNSUPPORTED) return "pdab: size outside the range the sm_100a kernels cover";
    if (code <= -1000) return cudaGetErrorString((cudaError_t)(-(code + 1000)));
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pdab: unknown error";
}
