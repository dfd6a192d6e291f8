// process.h -- drop-in for /root/reference/include/process.h (lines 29-30).
// TensorRTContext / get_thread_local_context (lines 13-26) are implementation details of the
// reference's TensorRT path and are not part of what src/main.cpp calls.
#ifndef PROCESS_H
#define PROCESS_H

#include <string>

namespace MedicalSeg {

// replaces process_single_image  (src/process.cpp:188-262): RAW 16-bit slice -> artefacts in output_dir
bool process_single_image(const std::string& raw_path, int width, int height, const std::string& output_dir);

}  // namespace MedicalSeg

#endif  // PROCESS_H
