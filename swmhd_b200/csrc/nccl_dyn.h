// nccl_dyn.h — libnccl.so.2 bound at run time (dlopen on first use).
// The halo exchange of multi-GPU contexts uses ncclSend/ncclRecv; a single-GPU user needs no NCCL, and
// the library must load on a machine without it, so there is no link-time dependency.  In a process
// that already holds a libnccl.so.2 (PyTorch's bundled copy) dlopen returns that one.
#pragma once
#include <nccl.h>      // types and prototypes only

namespace swmhd {

struct NcclApi {
    decltype(&ncclGetVersion) GetVersion;
    decltype(&ncclGetUniqueId) GetUniqueId;
    decltype(&ncclCommInitRank) CommInitRank;
    decltype(&ncclCommInitAll) CommInitAll;
    decltype(&ncclCommDestroy) CommDestroy;
    decltype(&ncclGetErrorString) GetErrorString;
    decltype(&ncclSend) Send;
    decltype(&ncclRecv) Recv;
    decltype(&ncclAllReduce) AllReduce;
    decltype(&ncclGroupStart) GroupStart;
    decltype(&ncclGroupEnd) GroupEnd;
};

const NcclApi *nccl_api();          // nullptr when libnccl.so.2 cannot be loaded
const char *nccl_load_error();

} // namespace swmhd
