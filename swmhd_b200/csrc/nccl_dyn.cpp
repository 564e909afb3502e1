// nccl_dyn.cpp — see nccl_dyn.h
#include "nccl_dyn.h"
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include <string>

namespace swmhd {
namespace {
NcclApi g_api;
bool g_ok = false;
std::string g_err = "not loaded";
std::once_flag g_once;

void load() {
    const char *names[] = {getenv("SWMHD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
        g_err = dlerror();
    }
    if (!h) return;
    bool ok = true;
    auto sym = [&](const char *name) -> void * {
        void *p = dlsym(h, name);
        if (!p) { ok = false; g_err = std::string("missing symbol ") + name; }
        return p;
    };
#define BIND(field, name) g_api.field = reinterpret_cast<decltype(g_api.field)>(sym(name))
    BIND(GetVersion, "ncclGetVersion");
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommInitAll, "ncclCommInitAll");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(GetErrorString, "ncclGetErrorString");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(AllReduce, "ncclAllReduce");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
#undef BIND
    g_ok = ok;
    if (ok) g_err.clear();
}
} // namespace

const NcclApi *nccl_api() {
    std::call_once(g_once, load);
    return g_ok ? &g_api : nullptr;
}
const char *nccl_load_error() { return g_err.c_str(); }

} // namespace swmhd
