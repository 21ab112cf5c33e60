// comm.cu — see comm.cuh.
#include "comm.cuh"
#include <dlfcn.h>
#include <mutex>

namespace elp {

namespace {
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
};
NcclApi g_api;
std::mutex g_api_mutex;
// One communicator per HOST THREAD: a process-per-GPU caller (torchrun) has one thread that owns one rank; the
// single-process multi-GPU solve (abi.cu: device pool) runs one worker thread per GPU, each with its own rank.
thread_local Comm g_comm;

void load_api() {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (g_api.lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.lib) break;
    }
    ELP_REQUIRE(g_api.lib, "cannot dlopen libnccl.so.2: %s", dlerror());
#define ELP_SYM(field, name)                                                    \
    *(void**)(&g_api.field) = dlsym(g_api.lib, name);                           \
    ELP_REQUIRE(g_api.field, "NCCL symbol %s not found", name)
    ELP_SYM(GetUniqueId, "ncclGetUniqueId");
    ELP_SYM(CommInitRank, "ncclCommInitRank");
    ELP_SYM(CommDestroy, "ncclCommDestroy");
    ELP_SYM(AllReduce, "ncclAllReduce");
    ELP_SYM(AllGather, "ncclAllGather");
    ELP_SYM(Send, "ncclSend");
    ELP_SYM(Recv, "ncclRecv");
    ELP_SYM(GroupStart, "ncclGroupStart");
    ELP_SYM(GroupEnd, "ncclGroupEnd");
    ELP_SYM(GetErrorString, "ncclGetErrorString");
#undef ELP_SYM
}

#define ELP_NCCL(call)                                                                         \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            throw elp::Error(elp::format("%s failed: %s", #call, g_api.GetErrorString(r__))); \
    } while (0)
}  // namespace

Comm& comm() { return g_comm; }

void comm_unique_id(void* id128) {
    load_api();
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    ELP_NCCL(g_api.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
}

void comm_init(int nranks, int rank, const void* id128) {
    load_api();
    ELP_REQUIRE(!g_comm.active, "communicator already initialised");
    ELP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ELP_NCCL(g_api.CommInitRank(&g_comm.comm, nranks, id, rank));
    g_comm.nranks = nranks;
    g_comm.rank = rank;
    g_comm.active = true;
}

void comm_destroy() {
    if (!g_comm.active) return;
    g_api.CommDestroy(g_comm.comm);
    g_comm = Comm{};
}

void comm_allreduce_sum(double* buf, size_t count, cudaStream_t stream) {
    if (!g_comm.active || count == 0) return;
    ELP_NCCL(g_api.AllReduce(buf, buf, count, ncclFloat64, ncclSum, g_comm.comm, stream));
}

void comm_allgather(double* buf, size_t count, cudaStream_t stream) {
    if (!g_comm.active || count == 0) return;
    ELP_NCCL(g_api.AllGather(buf + (size_t)g_comm.rank * count, buf, count, ncclFloat64, g_comm.comm, stream));
}

void comm_allgather_bytes(void* buf, size_t bytes, cudaStream_t stream) {
    if (!g_comm.active || bytes == 0) return;
    ELP_NCCL(g_api.AllGather((char*)buf + (size_t)g_comm.rank * bytes, buf, bytes, ncclInt8, g_comm.comm, stream));
}

void comm_group_start() { if (g_comm.active) ELP_NCCL(g_api.GroupStart()); }
void comm_group_end() { if (g_comm.active) ELP_NCCL(g_api.GroupEnd()); }

void comm_send(const void* buf, size_t bytes, int peer, cudaStream_t stream) {
    if (!g_comm.active || bytes == 0) return;
    ELP_NCCL(g_api.Send(buf, bytes, ncclInt8, peer, g_comm.comm, stream));
}

void comm_recv(void* buf, size_t bytes, int peer, cudaStream_t stream) {
    if (!g_comm.active || bytes == 0) return;
    ELP_NCCL(g_api.Recv(buf, bytes, ncclInt8, peer, g_comm.comm, stream));
}

void comm_allreduce_max(double* buf, size_t count, cudaStream_t stream) {
    if (!g_comm.active || count == 0) return;
    ELP_NCCL(g_api.AllReduce(buf, buf, count, ncclFloat64, ncclMax, g_comm.comm, stream));
}

}  // namespace elp
