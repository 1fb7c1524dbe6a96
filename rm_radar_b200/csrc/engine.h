// ONNX -> `.rmeng` plan builder and the reference's engine-path resolution (engine.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rmr {

// the bytes of the `.rmeng` file for this graph at this network input size
std::vector<uint8_t> compile_onnx(const std::string& onnx_path, int in_h, int in_w);
void build_engine(const std::string& onnx_path, const std::string& engine_path, int in_h, int in_w);

// `<x>.rmeng`, `<x>.engine` or `<x>.onnx` -> path of `<x>.rmeng`, built from the sibling `<x>.onnx` when absent
// (/root/reference/src/detect/detector.cpp:74-99).  std::invalid_argument when neither file exists.
std::string resolve_engine(const std::string& path, int in_h, int in_w);

}  // namespace rmr
