// NCCL entry points resolved with dlopen at first use, so that libncm_sd_gpu.so has no link-time
// dependency on NCCL and binds to the libnccl.so.2 the host process already loaded (torch's), if any.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

struct NcclApi {
  bool ok = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *)                                                                    = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                                             = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)              = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)           = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                                        = nullptr;
  const char *(*GetErrorString)(ncclResult_t)                                                                    = nullptr;
};

NcclApi &nccl_api();
