// NCCL, bound at run time.  The library is not linked against libnccl: the few entry points the solver needs are looked up
// with dlopen / dlsym the first time a communicator is created, so that (a) a single-GPU user needs no NCCL at all and
// (b) inside a Python process the instance torch already loaded (same soname) is the one that is used.
// Only the stable part of the NCCL 2.x ABI is declared here (nccl.h: ncclComm_t, ncclUniqueId, ncclAllReduce, ...).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>

namespace argus_nccl {

typedef struct ncclComm *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0 };
enum { ncclSum = 0, ncclMax = 2 };          // ncclRedOp_t
enum { ncclFloat64 = 8 };                    // ncclDataType_t

struct Api {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

inline Api &api()
{
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
    if (!a.handle) { fprintf(stderr, "argus_b200: libnccl.so.2 not found (%s)\n", dlerror()); return a; }
#define ARGUS_NCCL_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
    ARGUS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    ARGUS_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    ARGUS_NCCL_SYM(CommInitAll, "ncclCommInitAll");
    ARGUS_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    ARGUS_NCCL_SYM(AllReduce, "ncclAllReduce");
    ARGUS_NCCL_SYM(GroupStart, "ncclGroupStart");
    ARGUS_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    ARGUS_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef ARGUS_NCCL_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd;
    if (!a.ok) fprintf(stderr, "argus_b200: libnccl is missing an entry point\n");
    return a;
}

}  // namespace argus_nccl
