// Minimal run-time binding to NCCL (dlopen of libnccl.so.2 — the copy torch already loaded when the host is a
// torch.distributed process, the system one otherwise), so the library has no link-time NCCL dependency and
// single-GPU users never touch it.  Only the six entry points the grid engine needs are bound.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace dfg {

struct NcclUniqueId {
    char internal[128];
};
typedef struct ncclComm* NcclComm;

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;  // single-process multi-GPU handles
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;

    bool load() {
        if (ok) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        GetUniqueId = (int (*)(NcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
        CommInitAll = (int (*)(NcclComm*, int, const int*))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllReduce");
        GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
        ok = GetUniqueId && CommInitRank && CommDestroy && AllReduce && GetErrorString;
        return ok;
    }
};

constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace dfg
