// texir_host.h -- host-side helpers shared by the translation units of libtexir_b200.so (defined in
// texir_kernels.cu): thread-local error text, the launch counter behind texir_launch_count(), device lookup.
#pragma once
#include <string>

namespace texir_host {
int fail(const std::string& msg, int code = 1);  // records the text texir_last_error() returns; returns code
int finish_launch(const char* what);             // cudaGetLastError() check + launch counter; 0 = ok
int device_of(const void* p);                    // device that owns a device pointer (current device otherwise)
struct ScopedDevice {                            // cudaSetDevice for the scope of one entry point
    int prev = -1;
    explicit ScopedDevice(int dev);
    ~ScopedDevice();
};
}  // namespace texir_host
