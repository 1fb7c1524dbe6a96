// PCD v0.7 ingestion on the device.  The reference loads its clouds with
// pcl::io::loadPCDFile (/root/reference/samples/main.cpp:42-72): ASCII parsing on one CPU core for the
// frame clouds, a binary read for background.pcd.  Here the file image is uploaded as it is and parsed by
// kernels: the header (a dozen short lines) is read on the host, the body becomes xyz float32 in HBM —
// newline flags -> block scan -> line starts -> one thread per point parses its three fields.  Binary
// bodies are a strided view of the same upload.  Number parsing follows the oracle's semantics
// (oracle/locate_oracle.py: read_pcd — decimal text -> double -> float32).
#include "pcd.h"

#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace rmr {

namespace {

__constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

__device__ __forceinline__ bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\r'; }

// newline census: one thread per byte, per-block totals for the scan
__global__ void __launch_bounds__(256) pcd_count_kernel(const unsigned char* __restrict__ body, long n,
                                                        int* __restrict__ block_counts) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool nl = i < n && body[i] == '\n';
    const int c = __syncthreads_count(nl);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of the block totals by one block (block count is small: bytes / 256)
__global__ void __launch_bounds__(1024) pcd_scan_kernel(const int* __restrict__ counts, int* __restrict__ offsets,
                                                        int nblocks, int* __restrict__ total) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int prefix = carry + (warp ? warp_sums[warp - 1] : 0) + incl - v;
        if (i < nblocks) offsets[i] = prefix;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// line k+1 starts right after the k-th newline; line 0 starts at byte 0
__global__ void __launch_bounds__(256) pcd_line_start_kernel(const unsigned char* __restrict__ body, long n,
                                                             const int* __restrict__ block_offsets,
                                                             long* __restrict__ line_start, int max_lines) {
    __shared__ int warp_tot[8];
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool nl = i < n && body[i] == '\n';
    const unsigned ballot = __ballot_sync(0xffffffffu, nl);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    if (i == 0) line_start[0] = 0;
    if (!nl) return;
    int rank = block_offsets[blockIdx.x] + __popc(ballot & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) rank += warp_tot[w];
    if (rank + 1 < max_lines) line_start[rank + 1] = i + 1;
}

// decimal text -> double (exact for <= 19 significant digits and |exp10| <= 22, i.e. every value a LiDAR
// driver writes) -> float32: the oracle's conversion chain
__device__ float parse_number(const unsigned char* p, const unsigned char* end, const unsigned char** next) {
    while (p < end && is_space(*p)) ++p;
    const unsigned char* tok = p;
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) { neg = (*p == '-'); ++p; }
    if (p < end && (*p == 'n' || *p == 'N' || *p == 'i' || *p == 'I')) {
        const bool is_nan = (*p == 'n' || *p == 'N');
        while (p < end && !is_space(*p) && *p != '\n') ++p;
        *next = p;
        const float v = is_nan ? __int_as_float(0x7fc00000) : __int_as_float(0x7f800000);
        return neg ? -v : v;
    }
    unsigned long long mant = 0;
    int digits = 0, exp10 = 0;
    bool any = false;
    while (p < end && *p >= '0' && *p <= '9') {
        any = true;
        if (digits < 19) { mant = mant * 10ull + (*p - '0'); if (mant) ++digits; }
        else ++exp10;
        ++p;
    }
    if (p < end && *p == '.') {
        ++p;
        while (p < end && *p >= '0' && *p <= '9') {
            any = true;
            if (digits < 19) { mant = mant * 10ull + (*p - '0'); if (mant) ++digits; --exp10; }
            ++p;
        }
    }
    if (any && p < end && (*p == 'e' || *p == 'E')) {
        const unsigned char* q = p + 1;
        bool eneg = false;
        if (q < end && (*q == '-' || *q == '+')) { eneg = (*q == '-'); ++q; }
        if (q < end && *q >= '0' && *q <= '9') {
            int e = 0;
            while (q < end && *q >= '0' && *q <= '9') { if (e < 10000) e = e * 10 + (*q - '0'); ++q; }
            exp10 += eneg ? -e : e;
            p = q;
        }
    }
    if (!any) {   // not a number: skip the token, report NaN
        p = tok;
        while (p < end && !is_space(*p) && *p != '\n') ++p;
        *next = p;
        return __int_as_float(0x7fc00000);
    }
    *next = p;
    double v = static_cast<double>(mant);
    if (exp10 > 0) v = exp10 <= 22 ? v * kPow10[exp10] : v * pow(10.0, static_cast<double>(exp10));
    else if (exp10 < 0) v = -exp10 <= 22 ? v / kPow10[-exp10] : v / pow(10.0, static_cast<double>(-exp10));
    const float f = static_cast<float>(v);
    return neg ? -f : f;
}

// one thread per point: walk the line's whitespace-separated fields, keep the x / y / z columns
__global__ void __launch_bounds__(128) pcd_parse_kernel(const unsigned char* __restrict__ body, long n,
                                                        const long* __restrict__ line_start, int n_points, int fx,
                                                        int fy, int fz, int n_fields, float* __restrict__ xyz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const unsigned char* p = body + line_start[i];
    const unsigned char* end = body + n;
    float out[3] = {0.f, 0.f, 0.f};
    for (int f = 0; f < n_fields; ++f) {
        const unsigned char* next = p;
        const float v = parse_number(p, end, &next);
        if (f == fx) out[0] = v;
        if (f == fy) out[1] = v;
        if (f == fz) out[2] = v;
        p = next;
        if (p >= end || *p == '\n') break;
    }
    xyz[3 * i + 0] = out[0];
    xyz[3 * i + 1] = out[1];
    xyz[3 * i + 2] = out[2];
}

// binary body: x / y / z are float32 at byte offsets inside a fixed-size record
__global__ void __launch_bounds__(256) pcd_gather_kernel(const unsigned char* __restrict__ body, int n_points,
                                                         int record, int ox, int oy, int oz, float* __restrict__ xyz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const unsigned char* r = body + static_cast<size_t>(i) * record;
    float v[3];
    memcpy(&v[0], r + ox, 4);
    memcpy(&v[1], r + oy, 4);
    memcpy(&v[2], r + oz, 4);
    xyz[3 * i + 0] = v[0];
    xyz[3 * i + 1] = v[1];
    xyz[3 * i + 2] = v[2];
}

}  // namespace

PcdHeader pcd_parse_header(const void* file, size_t size) {
    PcdHeader h;
    const char* p = static_cast<const char*>(file);
    size_t pos = 0;
    std::vector<std::string> fields, types;
    std::vector<int> sizes, counts;
    long width = -1, height = 1;
    while (pos < size) {
        size_t e = pos;
        while (e < size && p[e] != '\n') ++e;
        if (e >= size) throw std::invalid_argument("PCD: header is not terminated by a DATA line");
        std::string line(p + pos, e - pos);
        pos = e + 1;
        std::istringstream ss(line);
        std::string key;
        ss >> key;
        if (key.empty() || key[0] == '#') continue;
        std::string tok;
        if (key == "FIELDS") { while (ss >> tok) fields.push_back(tok); }
        else if (key == "SIZE") { int v; while (ss >> v) sizes.push_back(v); }
        else if (key == "TYPE") { while (ss >> tok) types.push_back(tok); }
        else if (key == "COUNT") { int v; while (ss >> v) counts.push_back(v); }
        else if (key == "WIDTH") ss >> width;
        else if (key == "HEIGHT") ss >> height;
        else if (key == "POINTS") ss >> h.n_points;
        else if (key == "DATA") {
            ss >> tok;
            if (tok == "ascii") h.binary = false;
            else if (tok == "binary") h.binary = true;
            else throw std::invalid_argument("PCD: DATA " + tok + " is not supported (ascii and binary are)");
            h.body_offset = pos;
            break;
        }
    }
    if (h.body_offset == 0) throw std::invalid_argument("PCD: no DATA line");
    if (h.n_points < 0) h.n_points = width >= 0 ? width * height : -1;
    if (h.n_points < 0) throw std::invalid_argument("PCD: neither POINTS nor WIDTH given");
    if (fields.empty()) throw std::invalid_argument("PCD: no FIELDS line");
    if (sizes.size() != fields.size()) sizes.assign(fields.size(), 4);
    if (types.size() != fields.size()) types.assign(fields.size(), "F");
    if (counts.size() != fields.size()) counts.assign(fields.size(), 1);
    // SIZE in {1, 2, 4, 8} and COUNT >= 1, as in the PCD v0.7 grammar: a negative SIZE would put a field offset
    // before the record (device reads in front of the upload buffer), a zero record would divide by zero
    for (size_t i = 0; i < fields.size(); ++i) {
        if (sizes[i] != 1 && sizes[i] != 2 && sizes[i] != 4 && sizes[i] != 8)
            throw std::invalid_argument("PCD: SIZE must be 1, 2, 4 or 8");
        if (counts[i] < 1 || counts[i] > 65536) throw std::invalid_argument("PCD: COUNT must be at least 1");
    }
    if (h.n_points > (1L << 31)) throw std::invalid_argument("PCD: POINTS out of range");
    int column = 0, offset = 0;
    for (size_t i = 0; i < fields.size(); ++i) {
        const int which = fields[i] == "x" ? 0 : fields[i] == "y" ? 1 : fields[i] == "z" ? 2 : -1;
        if (which >= 0) {
            if (h.binary && (types[i] != "F" || sizes[i] != 4))
                throw std::invalid_argument("PCD: binary x/y/z must be float32");
            h.column[which] = column;
            h.offset[which] = offset;
        }
        column += counts[i];
        offset += sizes[i] * counts[i];
    }
    h.n_columns = column;
    h.record_bytes = offset;
    for (int k = 0; k < 3; ++k)
        if (h.column[k] < 0) throw std::invalid_argument("PCD: FIELDS must contain x, y and z");
    if (h.record_bytes <= 0) throw std::invalid_argument("PCD: empty record");
    // by division: POINTS * record_bytes may not wrap
    if (h.binary && static_cast<size_t>(h.n_points) > (size - h.body_offset) / static_cast<size_t>(h.record_bytes))
        throw std::invalid_argument("PCD: binary body is shorter than POINTS records");
    return h;
}

PcdParser::~PcdParser() {
    cudaFree(dev_bytes_); cudaFree(block_counts_); cudaFree(block_offsets_); cudaFree(line_start_); cudaFree(total_);
    cudaFreeHost(pinned_bytes_); cudaFreeHost(pinned_total_);
}

void PcdParser::reserve(size_t bytes, long lines) {
    if (bytes > cap_bytes_) {
        cudaFree(dev_bytes_); cudaFreeHost(pinned_bytes_); cudaFree(block_counts_); cudaFree(block_offsets_);
        cap_bytes_ = bytes + bytes / 4 + 4096;
        RMR_CUDA(cudaMalloc(&dev_bytes_, cap_bytes_));
        RMR_CUDA(cudaMallocHost(&pinned_bytes_, cap_bytes_));
        const size_t nb = (cap_bytes_ + 255) / 256;
        RMR_CUDA(cudaMalloc(&block_counts_, sizeof(int) * nb));
        RMR_CUDA(cudaMalloc(&block_offsets_, sizeof(int) * nb));
    }
    if (lines + 2 > cap_lines_) {
        cudaFree(line_start_);
        cap_lines_ = lines + lines / 4 + 1024;
        RMR_CUDA(cudaMalloc(&line_start_, sizeof(long) * cap_lines_));
    }
    if (!total_) {
        RMR_CUDA(cudaMalloc(&total_, sizeof(int)));
        RMR_CUDA(cudaMallocHost(&pinned_total_, sizeof(int)));
    }
}

int PcdParser::parse(const void* file, size_t size, float* dev_xyz, int capacity_points, cudaStream_t s) {
    if (!file || size == 0) throw std::invalid_argument("PCD: empty file image");
    const PcdHeader h = pcd_parse_header(file, size);
    if (h.n_points > capacity_points) throw std::invalid_argument("PCD: more points than the locator's max_points");
    if (h.n_points == 0) return 0;
    const size_t body = size - h.body_offset;
    reserve(body, h.n_points);
    RMR_CUDA(cudaStreamSynchronize(s));   // the pinned staging buffer of the previous file is free again
    std::memcpy(pinned_bytes_, static_cast<const char*>(file) + h.body_offset, body);
    RMR_CUDA(cudaMemcpyAsync(dev_bytes_, pinned_bytes_, body, cudaMemcpyHostToDevice, s));
    const int n = static_cast<int>(h.n_points);
    if (h.binary) {
        pcd_gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(dev_bytes_, n, h.record_bytes, h.offset[0], h.offset[1],
                                                          h.offset[2], dev_xyz);
    } else {
        const int nb = static_cast<int>((body + 255) / 256);
        pcd_count_kernel<<<nb, 256, 0, s>>>(dev_bytes_, static_cast<long>(body), block_counts_);
        pcd_scan_kernel<<<1, 1024, 0, s>>>(block_counts_, block_offsets_, nb, total_);
        pcd_line_start_kernel<<<nb, 256, 0, s>>>(dev_bytes_, static_cast<long>(body), block_offsets_, line_start_,
                                                 static_cast<int>(cap_lines_));
        RMR_CUDA(cudaMemcpyAsync(pinned_total_, total_, sizeof(int), cudaMemcpyDeviceToHost, s));
        RMR_CUDA(cudaStreamSynchronize(s));
        // n_points lines need n_points - 1 newlines (the last line may lack its terminator)
        if (*pinned_total_ + 1 < n) throw std::invalid_argument("PCD: ascii body has fewer lines than POINTS");
        pcd_parse_kernel<<<(n + 127) / 128, 128, 0, s>>>(dev_bytes_, static_cast<long>(body), line_start_, n,
                                                         h.column[0], h.column[1], h.column[2], h.n_columns, dev_xyz);
    }
    RMR_CUDA(cudaGetLastError());
    return n;
}

}  // namespace rmr
