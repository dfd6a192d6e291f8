// gen_json.cpp -- ORACLE helper (test infrastructure).  Emits, with the *reference's own vendored*
// nlohmann::json 3.12.0 (/root/reference/include/nlohmann/json.hpp, compiled in place, never copied),
// the two documents the reference writes, so tests can pin the product's hand-written formatter
// byte for byte:
//   mode "labelme": the statements of Mask2Polygon::generate_json  (src/mask2polygon.cpp:74-108)
//   mode "sidecar": the statements of Preprocess::preprocess_raw   (src/preprocess.cpp:126-134)
// stdin (labelme): base_name W H n_contours, then per contour: n x0 y0 x1 y1 ...
// stdin (sidecar): filename w h outW outH
// Built only where /root/reference exists (oracle/ref_json/Makefile -> oracle/_ref/gen_json).
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>
#include "nlohmann/json.hpp"
using json = nlohmann::json;

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "labelme";
    if (mode == "sidecar") {
        std::string name; int w, h, ow, oh;
        std::cin >> name >> w >> h >> ow >> oh;
        json j;
        j[name] = {{"original_width", w}, {"original_height", h}, {"scaled_width", ow}, {"scaled_height", oh}};
        std::cout << j << std::endl;
        return 0;
    }
    std::string base; int W, H, nc;
    std::cin >> base >> W >> H >> nc;
    json j;
    j["version"] = "1.0.2.812";
    j["imagePath"] = base + ".raw";
    j["imageData"] = nullptr;
    j["flags"] = json::object();
    j["shapes"] = json::array();
    for (int c = 0; c < nc; ++c) {
        int n; std::cin >> n;
        json shape;
        shape["label"] = 1;
        shape["labelIndex"] = 0;
        json points;
        for (int i = 0; i < n; ++i) { int x, y; std::cin >> x >> y; points.push_back({x, y}); }
        shape["points"] = points;
        shape["shape_type"] = "polygon";
        shape["description"] = "";
        shape["mask"] = nullptr;
        shape["group_id"] = nullptr;
        shape["flags"] = json::object();
        j["shapes"].push_back(shape);
    }
    j["imageWidth"] = W;
    j["imageHeight"] = H;
    std::cout << std::setw(4) << j << std::endl;
    return 0;
}
