// abi.cu -- error plumbing and build constants of the C-ABI (include/mccnn_b200.h).
//
// The reference aborts the process on any CUDA error (gpuErrchk -> exit, tf_ops/cuda_kernel_utils.h:16-24) and
// reports shape errors through TensorFlow's OP_REQUIRES.  Here every entry point returns an int status (0 = ok,
// 1 = invalid argument, 2 = CUDA error) and mc_last_error() returns the message of the calling thread's last
// failure; the Python layer turns them into ValueError / RuntimeError.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace mc {

static thread_local char g_lastError[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_kernelLaunches{0};

void count_launches(int n) { g_kernelLaunches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    count_launches(1);  // entry points launching more than one kernel add the rest with count_launches()
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) return 0;
    cudaGetLastError();  // clear the sticky launch error so that later calls are not blamed for it
    set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    return 2;
}

}  // namespace mc

extern "C" const char* mc_last_error(void) { return mc::g_lastError; }

// get_block_size() appended to MCConvModule.py at build time (tf_ops/genCompileScript.py:20,46-47)
extern "C" int mc_get_block_size(void) { return MC_BLOCK_MLP; }

extern "C" int mc_abi_version(void) { return 1; }

// number of kernels this library has launched so far in the process (bench.py reports it as gpu_launches)
extern "C" long long mc_launch_count(void) { return mc::g_kernelLaunches.load(); }

// SM count / compute capability of the current device; lets the host refuse to run on anything but sm_100.
extern "C" int mc_device_info(int* smCount, int* ccMajor, int* ccMinor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        mc::set_error("mc_device_info: %s", cudaGetErrorString(e));
        return 2;
    }
    if (smCount) *smCount = p.multiProcessorCount;
    if (ccMajor) *ccMajor = p.major;
    if (ccMinor) *ccMinor = p.minor;
    return 0;
}
