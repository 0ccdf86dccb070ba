// mie_abi.cu — version / error / device entry points of the C ABI (include/mie.h).
#include "mie_common.cuh"
#include "window.cuh"

namespace mie {
unsigned g_kernel_policy = MIE_POLICY_DEFAULT;
}

extern "C" {

int mie_abi_version(void) { return MIE_ABI_VERSION; }

int mie_set_kernel_policy(unsigned mask) {
    if (mask & ~(unsigned)MIE_POLICY_ALL) return MIE_E_UNSUPPORTED;
    mie::g_kernel_policy = mask;
    return MIE_OK;
}

unsigned mie_get_kernel_policy(void) { return mie::g_kernel_policy; }

const char* mie_error_string(int code) {
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    if (code <= MIE_E_NCCL_BASE) {
        switch (MIE_E_NCCL_BASE - code) {   // ncclResult_t
            case 1: return "NCCL: unhandled CUDA error";
            case 2: return "NCCL: unhandled system error";
            case 3: return "NCCL: internal error";
            case 4: return "NCCL: invalid argument";
            case 5: return "NCCL: invalid usage";
            case 6: return "NCCL: remote error";
            case 7: return "NCCL: operation in progress";
            default: return "NCCL: error";
        }
    }
    switch (code) {
        case MIE_OK: return "ok";
        case MIE_E_NULL: return "null pointer argument";
        case MIE_E_DTYPE: return "unsupported dtype or dtype combination";
        case MIE_E_SHAPE: return "invalid or unsupported shape";
        case MIE_E_STRIDE: return "stride smaller than the row or plane";
        case MIE_E_GRID: return "grid_size entries must be positive";
        case MIE_E_PAD: return "cannot compute tiles on the image according to the given grid size";
        case MIE_E_KERNEL: return "kernel size must be odd, positive and within the supported maximum";
        case MIE_E_BORDER: return "unknown border mode, or padding not smaller than the image";
        case MIE_E_WORKSPACE: return "workspace too small";
        case MIE_E_RANGE: return "value range must satisfy hi > lo";
        case MIE_E_UNSUPPORTED: return "request not implemented by this build";
        case MIE_E_ALIGN: return "workspace must be 256-byte aligned";
        default: return "unknown error";
    }
}

int mie_value_range_mode(int dtype, float lo, float hi) {
    if (dtype < MIE_U8 || dtype > MIE_F32) return -1;
    mie::WinCvt cv;
    return mie::range_mode(dtype, lo, hi, &cv);
}

int mie_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int v = 0;
    if (sm_count) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        *sm_count = v;
    }
    if (cc_major) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) return (int)e;
        *cc_major = v;
    }
    if (cc_minor) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev);
        if (e != cudaSuccess) return (int)e;
        *cc_minor = v;
    }
    return MIE_OK;
}

}  // extern "C"
