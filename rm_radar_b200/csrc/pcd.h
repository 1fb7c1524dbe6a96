// PCD v0.7 file images -> xyz float32 in device memory (see pcd.cu).
#pragma once
#include <cstddef>

#include "common.cuh"

namespace rmr {

struct PcdHeader {
    long n_points = -1;
    bool binary = false;
    size_t body_offset = 0;
    int column[3] = {-1, -1, -1};   // ascii: index of the x / y / z column on a line
    int offset[3] = {0, 0, 0};      // binary: byte offset of x / y / z inside a record
    int n_columns = 0, record_bytes = 0;
};

// host: the dozen header lines (FIELDS / SIZE / TYPE / COUNT / WIDTH / HEIGHT / POINTS / DATA)
PcdHeader pcd_parse_header(const void* file, size_t size);

class PcdParser {
public:
    PcdParser() = default;
    ~PcdParser();
    PcdParser(const PcdParser&) = delete;
    PcdParser& operator=(const PcdParser&) = delete;
    // uploads the body and parses it into dev_xyz ([n][3] float32, packed); returns the point count
    int parse(const void* file, size_t size, float* dev_xyz, int capacity_points, cudaStream_t s);

private:
    void reserve(size_t bytes, long lines);
    unsigned char *dev_bytes_ = nullptr, *pinned_bytes_ = nullptr;
    int *block_counts_ = nullptr, *block_offsets_ = nullptr, *total_ = nullptr, *pinned_total_ = nullptr;
    long* line_start_ = nullptr;
    size_t cap_bytes_ = 0;
    long cap_lines_ = 0;
};

}  // namespace rmr
