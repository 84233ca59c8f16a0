// gwz_nccl.cpp -- run-time binding of the handful of NCCL calls the library needs.
#include "gwz_nccl.h"

#include <dlfcn.h>
#include <mutex>
#include <string>

namespace gwz {

static NcclApi g_api;
static bool g_ok = false;
static std::string g_why;
static std::once_flag g_once;

static void load() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        g_why = std::string("cannot dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
        return;
    }
    struct { const char* name; void** slot; } syms[] = {
        {"ncclGetUniqueId", (void**)&g_api.GetUniqueId},   {"ncclCommInitRank", (void**)&g_api.CommInitRank},
        {"ncclCommInitAll", (void**)&g_api.CommInitAll},   {"ncclCommDestroy", (void**)&g_api.CommDestroy},
        {"ncclAllReduce", (void**)&g_api.AllReduce},       {"ncclGroupStart", (void**)&g_api.GroupStart},
        {"ncclGroupEnd", (void**)&g_api.GroupEnd},         {"ncclGetErrorString", (void**)&g_api.GetErrorString},
        {"ncclGetVersion", (void**)&g_api.GetVersion},
    };
    for (auto& s : syms) {
        *s.slot = dlsym(h, s.name);
        if (!*s.slot) {
            g_why = std::string("libnccl lacks symbol ") + s.name;
            return;
        }
    }
    g_ok = true;
}

const NcclApi* nccl_api(const char** why) {
    std::call_once(g_once, load);
    if (!g_ok) {
        if (why) *why = g_why.c_str();
        return nullptr;
    }
    return &g_api;
}

}  // namespace gwz
