// nccl_dyn.h -- NCCL bound at run time (dlopen), not at link time.
//
// libsbr_b200.so must not carry a DT_NEEDED on libnccl: the host process may already hold (or later load) its own
// NCCL -- e.g. PyTorch ships libnccl.so.2 2.28 while the system library is 2.27, and whichever copy is mapped first
// wins the soname for the whole process.  The synchronous multi-GPU exchange (sync_engine.cu) therefore resolves
// the handful of entry points it needs on first use: the copy already mapped in the process if there is one
// (RTLD_NOLOAD), else the system libnccl.so.2.  Single-GPU use never touches NCCL.
#pragma once
#include <nccl.h>  // types and enums only

#include <string>

namespace sbr {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GetVersion)(int*);
};

// nullptr (and *why set) when no usable libnccl.so.2 can be found; thread-safe, resolved once
const NcclApi* nccl_api(std::string* why);

}  // namespace sbr
