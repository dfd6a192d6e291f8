// initialize.h -- drop-in for /root/reference/include/initialize.h (lines 12-21).
//
// Same namespace and signatures the reference's src/main.cpp uses (src/main.cpp:2,15,88).  The
// reference header also leaks TensorRT types (get_engine(), g_runtime, g_engine; lines 15,24-25);
// main.cpp never touches them, so they are gone.  The first argument is now the path of a JSON
// config or a MSEGW001 weight blob instead of a serialized TensorRT engine.
#ifndef INITIALIZE_H
#define INITIALIZE_H

#include <fstream>
#include <memory>
#include <string>

struct ms_handle;

namespace MedicalSeg {

// replaces initialize_engine(trt_cache_path, log_dir)  (src/initialize.cpp:26-76)
bool initialize_engine(const std::string& engine_path, const std::string& log_dir);

// src/initialize.cpp:84-91
std::ofstream& get_log_file();
std::string get_log_path();

// the C-ABI handle behind the facade (nullptr before initialize_engine / after cleanup_resources)
ms_handle* get_handle();

}  // namespace MedicalSeg

#endif  // INITIALIZE_H
