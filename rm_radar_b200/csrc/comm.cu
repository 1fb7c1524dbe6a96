#include "comm.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../../include/rm_radar_b200.h"

namespace rmr {

namespace {

// the handful of NCCL entry points the exchange needs, with the types of nccl.h (ncclResult_t = int, ncclUniqueId =
// 128 bytes, ncclFloat = 7 in ncclDataType_t)
struct NcclUniqueId { char internal[kUniqueIdBytes]; };
using GetUniqueIdFn = int (*)(NcclUniqueId*);
using CommInitRankFn = int (*)(void**, int, NcclUniqueId, int);
using CommDestroyFn = int (*)(void*);
using CommAbortFn = int (*)(void*);
using AllGatherFn = int (*)(const void*, void*, size_t, int, void*, cudaStream_t);
using GetErrorStringFn = const char* (*)(int);
constexpr int kNcclFloat = 7;

struct Nccl {
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    CommAbortFn comm_abort = nullptr;
    AllGatherFn all_gather = nullptr;
    GetErrorStringFn error_string = nullptr;
};

const Nccl& nccl() {
    static Nccl api;
    static std::once_flag once;
    static std::string error;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { error = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return; }
        api.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
        api.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
        api.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
        api.comm_abort = reinterpret_cast<CommAbortFn>(dlsym(h, "ncclCommAbort"));
        api.all_gather = reinterpret_cast<AllGatherFn>(dlsym(h, "ncclAllGather"));
        api.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
        if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_gather || !api.error_string)
            error = "libnccl.so.2 lacks an expected entry point";
    });
    if (!error.empty()) throw std::runtime_error(error);
    return api;
}

void nccl_check(int r, const char* what) {
    if (r != 0) throw std::runtime_error(std::string(what) + ": " + nccl().error_string(r));
}

}  // namespace

void comm_unique_id(uint8_t out[kUniqueIdBytes]) {
    NcclUniqueId id;
    nccl_check(nccl().get_unique_id(&id), "ncclGetUniqueId");
    std::memcpy(out, id.internal, kUniqueIdBytes);
}

void Comm::pack(const rmr_robot* robots, int n, int max_robots, float* block) {
    std::memset(block, 0, sizeof(float) * kRecordFloats * max_robots);
    n = std::min(n, max_robots);
    for (int i = 0; i < n; ++i) {
        const rmr_robot& r = robots[i];
        float* o = block + static_cast<size_t>(i) * kRecordFloats;
        o[0] = 1.f;
        o[1] = r.is_detected ? static_cast<float>(r.label) : -1.f;
        o[2] = r.confidence;
        o[3] = r.is_located ? 1.f : 0.f;
        if (r.is_located) { o[4] = r.location[0]; o[5] = r.location[1]; o[6] = r.location[2]; }
        o[7] = r.rect[2] * r.rect[3];
    }
}

Comm::Comm(const uint8_t id[kUniqueIdBytes], int rank, int world, int device, int max_robots)
    : rank_(rank), world_(world), device_(device), max_robots_(max_robots) {
    if (world < 1 || rank < 0 || rank >= world || max_robots < 1) throw std::invalid_argument("bad communicator geometry");
    RMR_CUDA(cudaSetDevice(device_));
    NcclUniqueId uid;
    std::memcpy(uid.internal, id, kUniqueIdBytes);
    nccl_check(nccl().comm_init_rank(&comm_, world, uid, rank), "ncclCommInitRank");
    RMR_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    RMR_CUDA(cudaEventCreateWithFlags(&ready_, cudaEventDisableTiming));
    RMR_CUDA(cudaEventCreateWithFlags(&done_, cudaEventDisableTiming));
    const size_t block = sizeof(float) * kRecordFloats * max_robots;
    RMR_CUDA(cudaMallocHost(&pinned_in_, block));
    RMR_CUDA(cudaMallocHost(&pinned_out_, block * world));
    RMR_CUDA(cudaMalloc(&dev_in_, block));
    RMR_CUDA(cudaMalloc(&dev_out_, block * world));
}

// Orderly shutdown, called by every rank at the same point of the program: waits for the exchange in flight, then
// ncclCommDestroy.  A communicator that is merely dropped (process exit, exception) is aborted instead, which never
// waits for a peer that may already be gone.
void Comm::close() {
    if (!comm_) return;
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    nccl().comm_destroy(comm_);
    comm_ = nullptr;
}

Comm::~Comm() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    if (comm_) {
        if (nccl().comm_abort) nccl().comm_abort(comm_);
        else nccl().comm_destroy(comm_);
    }
    cudaFreeHost(pinned_in_); cudaFreeHost(pinned_out_); cudaFree(dev_in_); cudaFree(dev_out_);
    if (ready_) cudaEventDestroy(ready_);
    if (done_) cudaEventDestroy(done_);
    if (stream_) cudaStreamDestroy(stream_);
}

void Comm::publish(const rmr_robot* robots, int n, cudaStream_t after) {
    RMR_CUDA(cudaSetDevice(device_));
    if (pending_) RMR_CUDA(cudaEventSynchronize(done_));   // the previous block has left the pinned buffers
    pack(robots, n, max_robots_, pinned_in_);
    if (after) {
        RMR_CUDA(cudaEventRecord(ready_, after));
        RMR_CUDA(cudaStreamWaitEvent(stream_, ready_, 0));
    }
    const size_t count = static_cast<size_t>(kRecordFloats) * max_robots_;
    RMR_CUDA(cudaMemcpyAsync(dev_in_, pinned_in_, sizeof(float) * count, cudaMemcpyHostToDevice, stream_));
    nccl_check(nccl().all_gather(dev_in_, dev_out_, count, kNcclFloat, comm_, stream_), "ncclAllGather");
    RMR_CUDA(cudaMemcpyAsync(pinned_out_, dev_out_, sizeof(float) * count * world_, cudaMemcpyDeviceToHost, stream_));
    RMR_CUDA(cudaEventRecord(done_, stream_));
    pending_ = true;
}

void Comm::collect(float* out) {
    if (!pending_) throw std::invalid_argument("Comm::collect without publish");
    RMR_CUDA(cudaEventSynchronize(done_));
    std::memcpy(out, pinned_out_, sizeof(float) * kRecordFloats * max_robots_ * world_);
}

}  // namespace rmr
