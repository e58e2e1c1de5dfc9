// nccl_dl.cpp -- NCCL bound at run time (dlopen), so the library has no link-time dependency on
// it: batched runs never touch NCCL, only the row-sharded large problem does (one all-reduce of
// [lower(J^T J), J^T r] per Jacobian refresh and one scalar all-reduce per pass).
// Inside a PyTorch process dlopen("libnccl.so.2") resolves to the NCCL torch already loaded.
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <string>

#include "nccl_dl.h"

namespace mirb200 {

void set_error(const std::string& msg);

namespace {
struct UniqueId { char internal[128]; };
using GetUniqueId_t = int (*)(UniqueId*);
using CommInitRank_t = int (*)(void**, int, UniqueId, int);
using CommDestroy_t = int (*)(void*);
using AllReduce_t = int (*)(const void*, void*, size_t, int, int, void*, void*);
using GetErrorString_t = const char* (*)(int);

struct Api {
    void* handle = nullptr;
    GetUniqueId_t getUniqueId = nullptr;
    CommInitRank_t commInitRank = nullptr;
    CommDestroy_t commDestroy = nullptr;
    AllReduce_t allReduce = nullptr;
    GetErrorString_t errorString = nullptr;
    bool ok = false;
    std::string why;
};

Api& api()
{
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
        if (!a.handle) { const char* e = dlerror(); a.why = std::string("dlopen(libnccl.so.2) failed: ") + (e ? e : "?"); return; }
        a.getUniqueId = (GetUniqueId_t)dlsym(a.handle, "ncclGetUniqueId");
        a.commInitRank = (CommInitRank_t)dlsym(a.handle, "ncclCommInitRank");
        a.commDestroy = (CommDestroy_t)dlsym(a.handle, "ncclCommDestroy");
        a.allReduce = (AllReduce_t)dlsym(a.handle, "ncclAllReduce");
        a.errorString = (GetErrorString_t)dlsym(a.handle, "ncclGetErrorString");
        a.ok = a.getUniqueId && a.commInitRank && a.commDestroy && a.allReduce;
        if (!a.ok) a.why = "libnccl is missing one of ncclGetUniqueId/ncclCommInitRank/ncclCommDestroy/ncclAllReduce";
    });
    return a;
}

int fail(const char* what, int rc)
{
    Api& a = api();
    set_error(std::string("mir_optim_b200: ") + what + ": " + (a.errorString ? a.errorString(rc) : "NCCL error ") + " (" + std::to_string(rc) + ")");
    return MIR_B200_ENCCL;
}
}  // namespace

int nccl_available()
{
    Api& a = api();
    if (!a.ok) { set_error("mir_optim_b200: NCCL unavailable: " + a.why); return MIR_B200_ENCCL; }
    return MIR_B200_OK;
}

int nccl_allreduce_sum(void* buf, size_t count, bool is_double, void* comm, void* stream)
{
    if (int rc = nccl_available()) return rc;
    const int r = api().allReduce(buf, buf, count, is_double ? 8 /* ncclFloat64 */ : 7 /* ncclFloat32 */, 0 /* ncclSum */, comm, stream);
    return r ? fail("ncclAllReduce", r) : MIR_B200_OK;
}

}  // namespace mirb200

using namespace mirb200;

extern "C" {

int mir_b200_nccl_unique_id(void* id128)
{
    if (!id128) { set_error("mir_optim_b200: null id buffer"); return MIR_B200_EINVAL; }
    if (int rc = nccl_available()) return rc;
    UniqueId id;
    const int r = api().getUniqueId(&id);
    if (r) return fail("ncclGetUniqueId", r);
    std::memcpy(id128, &id, sizeof id);
    return MIR_B200_OK;
}

int mir_b200_nccl_comm_init(void** comm, int nranks, const void* id128, int rank)
{
    if (!comm || !id128) { set_error("mir_optim_b200: null argument"); return MIR_B200_EINVAL; }
    if (int rc = nccl_available()) return rc;
    UniqueId id;
    std::memcpy(&id, id128, sizeof id);
    const int r = api().commInitRank(comm, nranks, id, rank);
    return r ? fail("ncclCommInitRank", r) : MIR_B200_OK;
}

int mir_b200_nccl_comm_destroy(void* comm)
{
    if (!comm) return MIR_B200_OK;
    if (int rc = nccl_available()) return rc;
    const int r = api().commDestroy(comm);
    return r ? fail("ncclCommDestroy", r) : MIR_B200_OK;
}

}  // extern "C"
