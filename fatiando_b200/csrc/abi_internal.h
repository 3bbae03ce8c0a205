// abi_internal.h -- what the translation units behind the C ABI share.
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace fb200 {

// records the message fb200_last_error() returns (thread-local) and hands back `code`
int set_error(int code, const std::string &msg);

struct DeviceScope {
    int prev = -1;
    bool ok = false;
    explicit DeviceScope(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceScope()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace fb200

#define FB_CK(call)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fb200::set_error(FB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
