#include "nccl_dl.h"

#include <dlfcn.h>

#include <cstdio>
#include <mutex>
#include <stdexcept>
#include <string>

namespace gmg {

namespace {
template <typename F>
void bind(void* lib, const char* name, F& fn) {
    fn = reinterpret_cast<F>(dlsym(lib, name));
    if (!fn) throw std::runtime_error(std::string("libnccl.so.2 lacks ") + name);
}
}  // namespace

const NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    static std::string error;
    std::call_once(once, [] {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            error = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror();
            return;
        }
        try {
            bind(lib, "ncclGetUniqueId", api.GetUniqueId);
            bind(lib, "ncclCommInitRank", api.CommInitRank);
            bind(lib, "ncclCommDestroy", api.CommDestroy);
            bind(lib, "ncclGroupStart", api.GroupStart);
            bind(lib, "ncclGroupEnd", api.GroupEnd);
            bind(lib, "ncclSend", api.Send);
            bind(lib, "ncclRecv", api.Recv);
            bind(lib, "ncclAllReduce", api.AllReduce);
            bind(lib, "ncclAllGather", api.AllGather);
            bind(lib, "ncclBroadcast", api.Broadcast);
            bind(lib, "ncclGetErrorString", api.GetErrorString);
        } catch (const std::exception& e) {
            error = e.what();
        }
    });
    if (!error.empty()) throw std::runtime_error(error);
    return api;
}

void nccl_check(ncclResult_t r, const char* what, const char* file, int line) {
    if (r != ncclSuccess) {
        char buf[512];
        std::snprintf(buf, sizeof buf, "NCCL error %d (%s) at %s:%d: %s", (int)r, nccl().GetErrorString(r), file, line, what);
        throw std::runtime_error(buf);
    }
}

}  // namespace gmg
