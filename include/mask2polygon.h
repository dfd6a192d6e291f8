// mask2polygon.h -- drop-in for /root/reference/include/mask2polygon.h (lines 8-22) with cv::Mat /
// cv::Point / nlohmann::json replaced by plain types (OpenCV's C++ SDK and nlohmann are not
// dependencies of this library).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace Mask2Polygon {

struct Point {                // stands in for cv::Point
    int x, y;
    bool operator==(const Point& o) const { return x == o.x && y == o.y; }
};
struct MaskView {             // stands in for a CV_8UC1 cv::Mat
    const uint8_t* data;
    int rows, cols;
};
struct SizeInfo {             // what the reference reads out of the sidecar JSON (src/mask2polygon.cpp:146-160)
    std::string filename;
    int original_width = 0, original_height = 0, scaled_width = 0, scaled_height = 0;
};

// replaces load_size_json  (src/mask2polygon.cpp:16-24) + the key lookup at :146-160
SizeInfo load_size_json(const std::string& json_path, const std::string& base_name);

// replaces extract_contours  (src/mask2polygon.cpp:29-36): threshold(127) + findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
std::vector<std::vector<Point>> extract_contours(const MaskView& mask);

// replaces map_contour_points  (src/mask2polygon.cpp:41-63)
std::vector<std::vector<Point>> map_contour_points(const std::vector<std::vector<Point>>& contours, double scale_x, double scale_y);

// replaces generate_json  (src/mask2polygon.cpp:68-109); byte-identical output
void generate_json(const std::vector<std::vector<Point>>& contours, const std::string& json_path, const std::string& base_name,
                   int original_width, int original_height);

// replaces create_overlay_image  (src/mask2polygon.cpp:114-129); `gray` is the normalised 8-bit image
void create_overlay_image(const std::vector<std::vector<Point>>& contours, const MaskView& gray, const std::string& overlay_path);

// replaces process_single_mask  (src/mask2polygon.cpp:134-222): sidecar lookup, mask PNG read + size check, contours,
// overlay on `original_png` (if not empty), coordinate mapping, <output_dir>/<base_name>.json.  Like the reference it
// reports failures on std::cerr and returns nothing (src/mask2polygon.cpp:219-221).
void process_single_mask(const std::string& mask_path, const std::string& output_dir, const std::string& json_path,
                         const std::string& original_png, const std::string& base_name);

}  // namespace Mask2Polygon
