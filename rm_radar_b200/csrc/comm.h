// Multi-GPU exchange inside the library: one NCCL all-gather of the fixed-size block of world-frame robot records per
// step, issued on the communicator's own stream (the reference has no distributed code).
// Record layout (8 floats per robot, max_robots rows per rank) = rm_radar_b200/dist.py:
//   [valid, label (-1 = undetected), confidence, is_located, x, y, z (metres, world), rect area]
// NCCL is resolved at run time (dlopen libnccl.so.2): single-GPU users carry no dependency, and inside a torch
// process the library instance torch already loaded is the one that gets used.
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

struct rmr_robot;   // include/rm_radar_b200.h

namespace rmr {

constexpr int kRecordFloats = 8;
constexpr int kUniqueIdBytes = 128;

void comm_unique_id(uint8_t out[kUniqueIdBytes]);

class Comm {
public:
    Comm(const uint8_t id[kUniqueIdBytes], int rank, int world, int device, int max_robots);
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    // pack `n` robot records of this rank, upload, all-gather: everything is enqueued on the communicator's stream and
    // the call returns at once; `after` (may be null) is a stream whose work so far the exchange must follow
    void publish(const rmr_robot* robots, int n, cudaStream_t after);
    // wait for the last publish and copy out [world][max_robots][8] floats
    void collect(float* out);
    // collective shutdown (every rank, same program point); the destructor of an unclosed communicator aborts instead
    void close();
    int world() const { return world_; }
    int rank() const { return rank_; }
    int max_robots() const { return max_robots_; }
    cudaStream_t stream() const { return stream_; }
    // host-side packing of one record block (also what the gloo CPU test checks against dist.pack_records)
    static void pack(const rmr_robot* robots, int n, int max_robots, float* block);

private:
    int rank_, world_, device_, max_robots_;
    void* comm_ = nullptr;                 // ncclComm_t
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ready_ = nullptr, done_ = nullptr;
    float *pinned_in_ = nullptr, *pinned_out_ = nullptr, *dev_in_ = nullptr, *dev_out_ = nullptr;
    bool pending_ = false;
};

}  // namespace rmr
