// NCCL communicator owned by the library (one process per GPU; tensor parallelism over NVLink 5 / NVSwitch).
// NCCL is resolved at run time with dlopen("libnccl.so.2"): when torch has already loaded its bundled copy the same
// handle is returned, otherwise the system library is used; the library itself has no link-time NCCL dependency.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace fl {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;

    void load() {
        if (lib) return;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        FL_CHECK(lib != nullptr, -4, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
        auto sym = [&](const char* n) {
            void* p = dlsym(lib, n);
            FL_CHECK(p != nullptr, -4, std::string("NCCL symbol missing: ") + n);
            return p;
        };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    }
    void check(ncclResult_t r, const char* what) {
        if (r != ncclSuccess) throw Error(-4, std::string(what) + ": " + (GetErrorString ? GetErrorString(r) : "NCCL error"));
    }
};

extern NcclApi g_nccl;

// Peer-mapped exchange area of the persistent decode kernel's in-kernel all-reduce (CUDA IPC; NVLink P2P stores).
// One fixed-layout 4 MiB buffer per process; every rank maps every peer's buffer.
struct PeerComm {
    static constexpr size_t kBytes = 4u << 20;
    static constexpr int kMaxH = 16384, kMaxTp = 8;
    static constexpr size_t kPartOff = 0;                                   // [2][8][kMaxH] f32 = 1 MiB
    static constexpr size_t kFlagOff = 1u << 20;                            // [8][gridDim.x + 1] u32
    static constexpr size_t kAmaxOff = (1u << 20) + (64u << 10);            // [8][2] f32
    static constexpr size_t kErrOff = (1u << 20) + (68u << 10);             // int
    static constexpr size_t kLogitsOff = (1u << 20) + (128u << 10);         // [Vfull] f32
    static constexpr size_t kMaxVocab = (kBytes - kLogitsOff) / 4;
    void* local = nullptr;
    void* peer[kMaxTp] = {};
    bool ready = false;
    unsigned int ar_epoch = 0;      // exchanges performed so far (identical on every rank: same launch sequence)
};
extern PeerComm g_peer;

}  // namespace fl
