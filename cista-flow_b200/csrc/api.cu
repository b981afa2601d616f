// Library-level entry points: version, error string, device check.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace cf {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static std::atomic<const char *> g_last_kernel{""};   // process-wide: autograd runs backward kernels on its own thread

void count_launch(const char *kernel) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g_last_kernel.store(kernel, std::memory_order_relaxed);
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct DevInfo {
    int status = 1;  // 1 = not probed yet
    int sms = 0;
};
static DevInfo g_dev[64];

static int probe(int dev) {
    DevInfo &d = g_dev[dev];
    if (d.status != 1) return d.status;
    int major = 0, minor = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        set_error("cudaDeviceGetAttribute failed on device %d", dev);
        return CF_ERR_CUDA;  // not cached: may be transient
    }
    d.sms = sms;
    if (major != 10) {
        set_error("device %d is sm_%d%d; libcistaflow is built for sm_100a (B200) only and has no fallback",
                  dev, major, minor);
        d.status = CF_ERR_ARCH;
    } else {
        d.status = CF_OK;
    }
    return d.status;
}

int check_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        set_error("no usable CUDA device (cudaGetDevice failed); libcistaflow has no CPU fallback");
        return CF_ERR_CUDA;
    }
    return probe(dev);
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (g_dev[dev].status == 1) probe(dev);
    return g_dev[dev].sms > 0 ? g_dev[dev].sms : 148;
}

}  // namespace cf

extern "C" {

int cf_version(void) { return CISTAFLOW_VERSION; }

const char *cf_last_error(void) { return cf::g_err; }

int cf_device_check(void) { return cf::check_device(); }

int64_t cf_launch_count(void) { return cf::g_launches.load(std::memory_order_relaxed); }

const char *cf_last_kernel(void) { return cf::g_last_kernel.load(std::memory_order_relaxed); }

}  // extern "C"
