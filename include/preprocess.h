// preprocess.h -- drop-in for /root/reference/include/preprocess.h (lines 20-23).
#pragma once
#include <string>

namespace Preprocess {

// replaces preprocess_raw  (src/preprocess.cpp:76-141): RAW u16 -> 512x512 8-bit PNG + size sidecar JSON.
// Needs MedicalSeg::initialize_engine first (the resample runs on the GPU; a stage-only handle is
// created on demand when no engine was initialised).
bool preprocess_raw(const std::string& raw_path, const std::string& png_path, const std::string& json_path, int w, int h);

}  // namespace Preprocess
