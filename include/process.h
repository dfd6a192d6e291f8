// process.h -- drop-in for /root/reference/include/process.h (lines 29-30).
// TensorRTContext / get_thread_local_context (lines 13-26) are implementation details of the
// reference's TensorRT path and are not part of what src/main.cpp calls.
#ifndef PROCESS_H
#define PROCESS_H

#include <string>
#include <vector>

namespace MedicalSeg {

// replaces process_single_image  (src/process.cpp:188-262): RAW 16-bit slice -> artefacts in output_dir
bool process_single_image(const std::string& raw_path, int width, int height, const std::string& output_dir);

// Extensions (not in the reference): the same work for many files at once, batched on the GPU with file reads and
// artefact writes overlapped.  Output of every file is byte-identical to process_single_image.
//   process_image_batch : the loop body of src/main.cpp:148-164 for a list; returns the number of successes, ok[i] per file
//   process_directory   : find_16bit_images + that loop (src/main.cpp:28-48,134-168); prints the reference's summary
int process_image_batch(const std::vector<std::string>& raw_paths, int width, int height, const std::vector<std::string>& output_dirs,
                        std::vector<bool>* ok = nullptr);
bool process_directory(const std::string& input_dir, int width, int height, const std::string& output_dir, bool recursive,
                       int* success_count = nullptr, int* fail_count = nullptr);

}  // namespace MedicalSeg

#endif  // PROCESS_H
