// facade_driver.cpp -- test program for the reference-shaped C++ stage API (include/*.h over libmedseg_b200.so).
// usage: facade_driver <weights.msegw> <slice.raw> <w> <h> <out_dir>
// 1. MedicalSeg::initialize_engine / process_single_image / cleanup_resources      (what src/main.cpp calls)
// 2. the stages one by one, the way src/process.cpp:211-242 chains them through files:
//    Preprocess::preprocess_raw -> (mask PNG written by step 1) -> Mask2Polygon::process_single_mask
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <fstream>
#include <iostream>
#include <string>
#include "cleanup.h"
#include "initialize.h"
#include "mask2polygon.h"
#include "postprocess.h"
#include "preprocess.h"
#include "process.h"

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const std::string blob = argv[1], raw = argv[2], out = argv[5];
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]);
    if (!MedicalSeg::initialize_engine(blob, out + "/log")) return 3;
    MedicalSeg::get_log_file() << "driver: engine ready" << std::endl;
    if (!MedicalSeg::process_single_image(raw, w, h, out + "/a")) return 4;
    // stage by stage, through files
    const std::string base = "slice";
    if (!Preprocess::preprocess_raw(raw, out + "/b/" + base + "_normalized.png", out + "/b/" + base + "_original_sizes.json", w, h)) return 5;
    Mask2Polygon::process_single_mask(out + "/a/" + base + "_mask.png", out + "/b", out + "/b/" + base + "_original_sizes.json",
                                      out + "/b/" + base + "_normalized.png", base);
    // in-memory stage calls
    std::vector<uint8_t> m(64 * 64, 0);
    for (int y = 8; y < 56; ++y)
        for (int x = 8; x < 56; ++x) m[y * 64 + x] = 2;
    m[30 * 64 + 30] = 0;   // a one-pixel hole: filled by postprocess
    const std::vector<uint8_t> pp = MedicalSeg::postprocess_mask(MedicalSeg::MaskView{m.data(), 64, 64});
    const std::vector<uint8_t> vis = MedicalSeg::mask_to_image(MedicalSeg::MaskView{pp.data(), 64, 64});
    const auto cs = Mask2Polygon::extract_contours(Mask2Polygon::MaskView{vis.data(), 64, 64});
    std::printf("inmem: hole=%d contours=%zu pts=%zu first=(%d,%d)\n", (int)pp[30 * 64 + 30], cs.size(), cs.empty() ? 0 : cs[0].size(),
                cs.empty() ? -1 : cs[0][0].x, cs.empty() ? -1 : cs[0][0].y);
    // the batched extensions: a directory holding the same slice twice -> same JSON as the single call
    {
        const std::string in_dir = out + "/dir_in";
        std::system(("mkdir -p '" + in_dir + "' && cp '" + raw + "' '" + in_dir + "/s0.raw' && cp '" + raw + "' '" + in_dir + "/s1.RAW'").c_str());
        int good = -1, bad = -1;
        const bool okd = MedicalSeg::process_directory(in_dir, w, h, out + "/dir_out", false, &good, &bad);
        std::vector<bool> flags;
        const int n = MedicalSeg::process_image_batch({in_dir + "/s0.raw", in_dir + "/missing.raw"}, w, h, {out + "/list0", out + "/list1"}, &flags);
        std::printf("dir: ok=%d good=%d bad=%d list=%d flags=%d%d\n", (int)okd, good, bad, n, (int)flags[0], (int)flags[1]);
    }
    std::printf("log: %s\n", MedicalSeg::get_log_path().c_str());
    MedicalSeg::cleanup_resources();
    return 0;
}
