// postprocess.h -- the reference has no header for this stage: src/postprocess.cpp is textually
// #included at src/process.cpp:9.  OpenCV's C++ SDK is not a dependency here, so cv::Mat becomes a
// plain view.
#pragma once
#include <cstdint>
#include <vector>

namespace MedicalSeg {

struct MaskView {             // stands in for a CV_8UC1 cv::Mat
    const uint8_t* data;
    int rows, cols;
};

// replaces cv::Mat postprocess_mask(const cv::Mat&)  (src/postprocess.cpp:47-79); returns rows*cols bytes
std::vector<uint8_t> postprocess_mask(const MaskView& src);

// replaces cv::Mat mask_to_image(const cv::Mat&)  (src/process.cpp:178-185)
std::vector<uint8_t> mask_to_image(const MaskView& mask);

}  // namespace MedicalSeg
