// A stand-in for libcuda.so.1 -- TEST INFRASTRUCTURE ONLY (see cuda_shim.h).
//
// The CPU-only test tier has no CUDA driver, so the host half of the C ABI (csrc/sb_api.cpp:
// staging buffers, argument blocks, kernel selection, status routing, history strides) could only
// run on the GPU box.  This library implements the ~35 driver entry points sb_api.cpp resolves
// with dlsym: "device" memory is host memory, streams are synchronous, and cuLaunchKernel hands
// the argument block to the host emulation of the same kernel (tests/emu/emu_main.cpp, built from
// the device sources for the same problem; its path is read from $SB_FAKE_EMU_LIB when a module
// is loaded).  tests/test_host_logic.py puts a directory holding this library first on
// LD_LIBRARY_PATH of a child process; the product never loads it.
#include <cuda.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {
struct FakeFunc { std::string name; void (*fn)(const void*) = nullptr; };
struct FakeModule {
    void* lib = nullptr;
    int group_size = 1;
    int hist_stride = 0;
    FakeFunc funcs[16];
    int n_funcs = 0;
};
int g_launches = 0;
const char* emu_symbol(const char* kernel) {
    static const char* const map[][2] = {
        {"sb_forward", "emu_forward"}, {"sb_forward_sens", "emu_forward_sens"},
        {"sb_tables", "emu_tables"}, {"sb_backward", "emu_backward"},
        {"sb_backward_fund", "emu_backward_fund"}, {"sb_eval", "emu_eval"}, {nullptr, nullptr}};
    for (int i = 0; map[i][0]; ++i)
        if (!strcmp(map[i][0], kernel)) return map[i][1];
    return nullptr;
}
}  // namespace

extern "C" {
int fake_cuda_launches() { return g_launches; }

CUresult cuInit(unsigned) { return CUDA_SUCCESS; }
CUresult cuGetErrorString(CUresult, const char** s) { *s = "fake driver error"; return CUDA_SUCCESS; }
CUresult cuDeviceGetCount(int* n) { *n = 1; return CUDA_SUCCESS; }
CUresult cuDeviceGet(CUdevice* d, int i) { *d = i; return i == 0 ? CUDA_SUCCESS : CUDA_ERROR_INVALID_DEVICE; }
CUresult cuDeviceGetAttribute(int* v, CUdevice_attribute a, CUdevice) {
    *v = (a == CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT) ? 148 : 0;
    return CUDA_SUCCESS;
}
CUresult cuDevicePrimaryCtxRetain(CUcontext* c, CUdevice) { *c = (CUcontext)0x1; return CUDA_SUCCESS; }
CUresult cuDevicePrimaryCtxRelease(CUdevice) { return CUDA_SUCCESS; }
CUresult cuCtxPushCurrent(CUcontext) { return CUDA_SUCCESS; }
CUresult cuCtxPopCurrent(CUcontext* c) { if (c) *c = (CUcontext)0x1; return CUDA_SUCCESS; }
CUresult cuCtxSynchronize(void) { return CUDA_SUCCESS; }

CUresult cuModuleLoadData(CUmodule* mod, const void* image) {
    if (!image || memcmp(image, "\x7f" "ELF", 4) != 0) return CUDA_ERROR_INVALID_IMAGE;
    const char* path = getenv("SB_FAKE_EMU_LIB");
    if (!path) return CUDA_ERROR_FILE_NOT_FOUND;
    FakeModule* m = new FakeModule();
    m->lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!m->lib) { fprintf(stderr, "fake_cuda: %s\n", dlerror()); delete m; return CUDA_ERROR_FILE_NOT_FOUND; }
    if (auto f = (int (*)())dlsym(m->lib, "emu_hist_stride")) m->hist_stride = f();
    *mod = (CUmodule)m;
    return CUDA_SUCCESS;
}
CUresult cuModuleUnload(CUmodule mod) { delete (FakeModule*)mod; return CUDA_SUCCESS; }
CUresult cuModuleGetFunction(CUfunction* f, CUmodule mod, const char* name) {
    FakeModule* m = (FakeModule*)mod;
    FakeFunc& ff = m->funcs[m->n_funcs];
    ff.name = name;
    ff.fn = nullptr;
    if (!strcmp(name, "sb_backward_flat")) {
        // the flat build of the backward kernel computes what sb_backward computes
        // (tests/test_gpu_parity.py), so it is a no-op here
    } else {
        const char* sym = emu_symbol(name);
        ff.fn = sym ? (void (*)(const void*))dlsym(m->lib, sym) : nullptr;
        if (!ff.fn) return CUDA_ERROR_NOT_FOUND;
    }
    ++m->n_funcs;
    *f = (CUfunction)&ff;
    return CUDA_SUCCESS;
}
CUresult cuModuleGetGlobal(CUdeviceptr* p, size_t* bytes, CUmodule mod, const char* name) {
    FakeModule* m = (FakeModule*)mod;
    int* v = !strcmp(name, "sb_group_size") ? &m->group_size
             : (!strcmp(name, "sb_hist_stride") && m->hist_stride) ? &m->hist_stride : nullptr;
    if (!v) return CUDA_ERROR_NOT_FOUND;
    *p = (CUdeviceptr)v; *bytes = sizeof(int);
    return CUDA_SUCCESS;
}
CUresult cuFuncGetAttribute(int* v, CUfunction_attribute a, CUfunction) {
    *v = (a == CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK) ? 32 : (a == CU_FUNC_ATTRIBUTE_NUM_REGS) ? 255 : 0;
    return CUDA_SUCCESS;
}
CUresult cuFuncSetAttribute(CUfunction, CUfunction_attribute, int) { return CUDA_SUCCESS; }
CUresult cuOccupancyMaxActiveBlocksPerMultiprocessor(int* n, CUfunction, int, size_t) { *n = 8; return CUDA_SUCCESS; }

CUresult cuMemAlloc(CUdeviceptr* p, size_t bytes) {
    void* q = malloc(bytes ? bytes : 1);
    if (!q) return CUDA_ERROR_OUT_OF_MEMORY;
    memset(q, 0xA5, bytes);          // uninitialised device memory is not zero
    *p = (CUdeviceptr)q;
    return CUDA_SUCCESS;
}
CUresult cuMemFree(CUdeviceptr p) { free((void*)p); return CUDA_SUCCESS; }
CUresult cuMemGetInfo(size_t* free_b, size_t* total_b) {
    *free_b = (size_t)8 << 30; *total_b = (size_t)16 << 30;
    return CUDA_SUCCESS;
}
CUresult cuMemHostAlloc(void** p, size_t bytes, unsigned) { *p = malloc(bytes ? bytes : 1); return *p ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY; }
CUresult cuMemFreeHost(void* p) { free(p); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoH(void* d, CUdeviceptr s, size_t n) { memcpy(d, (const void*)s, n); return CUDA_SUCCESS; }
CUresult cuMemcpyHtoDAsync(CUdeviceptr d, const void* s, size_t n, CUstream) { memcpy((void*)d, s, n); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoHAsync(void* d, CUdeviceptr s, size_t n, CUstream) { memcpy(d, (const void*)s, n); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoDAsync(CUdeviceptr d, CUdeviceptr s, size_t n, CUstream) { memmove((void*)d, (const void*)s, n); return CUDA_SUCCESS; }
CUresult cuMemsetD8Async(CUdeviceptr d, unsigned char v, size_t n, CUstream) { memset((void*)d, v, n); return CUDA_SUCCESS; }
CUresult cuMemsetD32Async(CUdeviceptr d, unsigned v, size_t n, CUstream) {
    unsigned* p = (unsigned*)d;
    for (size_t i = 0; i < n; ++i) p[i] = v;
    return CUDA_SUCCESS;
}
CUresult cuStreamCreate(CUstream* s, unsigned) { *s = (CUstream)0x2; return CUDA_SUCCESS; }
CUresult cuStreamDestroy(CUstream) { return CUDA_SUCCESS; }
CUresult cuStreamSynchronize(CUstream) { return CUDA_SUCCESS; }
CUresult cuStreamWaitEvent(CUstream, CUevent, unsigned) { return CUDA_SUCCESS; }
CUresult cuEventCreate(CUevent* e, unsigned) { *e = (CUevent)0x3; return CUDA_SUCCESS; }
CUresult cuEventDestroy(CUevent) { return CUDA_SUCCESS; }
CUresult cuEventRecord(CUevent, CUstream) { return CUDA_SUCCESS; }
CUresult cuEventSynchronize(CUevent) { return CUDA_SUCCESS; }
CUresult cuEventElapsedTime(float* ms, CUevent, CUevent) { *ms = 0.125f; return CUDA_SUCCESS; }

CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned, unsigned, unsigned bx, unsigned, unsigned,
                        unsigned, CUstream, void** params, void**) {
    FakeFunc* ff = (FakeFunc*)f;
    if (gx == 0 || bx == 0) return CUDA_ERROR_INVALID_VALUE;
    ++g_launches;
    if (ff->fn) ff->fn(params[0]);
    return CUDA_SUCCESS;
}
}  // extern "C"
