// Library identification and status strings (safe without a GPU).
#include "common.cuh"

extern "C" {

const char *ub200_version(void) { return "unet_b200 0.1 (sm_100a)"; }
int ub200_abi_version(void) { return 1; }

const char *ub200_status_string(int status) {
    switch (status) {
        case UB200_OK: return "ok";
        case UB200_E_BADARG: return "bad argument";
        case UB200_E_UNSUPPORTED: return "unsupported shape / alignment";
        case UB200_E_NODEVICE: return "no sm_100 device or driver entry point";
        default: return status > 0 ? cudaGetErrorName((cudaError_t)status) : "unknown status";
    }
}

}  // extern "C"
