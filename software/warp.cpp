// software/warp, headless: coarse-to-fine warp of two .tri stacks onto each other's image, one
// direction per level, writing "<tri>.warp" (software/warp/main.cpp of the reference).
//
//   warp -ia a.ppm -ib b.ppm -ta a.ppm.tri -tb b.ppm.tri [-native] [-frames F] [-device D]
#include <tpose/tpose>
#include <tpose/triangulation>
#include <tpose/io>
#include <tpose/schedule>

#include <chrono>

#include "image_io.h"

int main(int argc, char* args[])
{
    const std::string ia = option(argc, args, "ia"), ib = option(argc, args, "ib");
    const std::string ta = option(argc, args, "ta"), tb = option(argc, args, "tb");
    if (ia.empty()) { std::cout << "Please specify a first input image with -ia." << std::endl; return 0; }
    if (ib.empty()) { std::cout << "Please specify a second input image with -ib." << std::endl; return 0; }
    if (ta.empty()) { std::cout << "Please specify a first input triangulation with -ta." << std::endl; return 0; }
    if (tb.empty()) { std::cout << "Please specify a second input triangulation with -tb." << std::endl; return 0; }
    Image A, B;
    if (!load_image(ia, A)) { std::cout << "Failed to load image A." << std::endl; return 0; }
    if (!load_image(ib, B)) { std::cout << "Failed to load image B." << std::endl; return 0; }

    if (A.w != B.w || A.h != B.h) { std::cout << "Images don't have the same dimension" << std::endl; return 0; }   // warp/main.cpp:51-54

    const bool native = has_flag(argc, args, "native");
    // the window is sized from image A for both views (warp/main.cpp:64)
    const int rw = native ? A.w : (int)(A.w / 1.5), rh = native ? A.h : (int)(A.h / 1.5);
    const long frames = atol(option(argc, args, "frames", "-1").c_str());

    tpose::init(atoi(option(argc, args, "device", "0").c_str()));
    tpose::texture(TPOSE_IMAGE_A, A.rgba.data(), A.w, A.h, rw, rh);              // warp/main.cpp:118-119
    tpose::texture(TPOSE_IMAGE_B, B.rgba.data(), B.w, B.h, rw, rh);
    const std::string ss = option(argc, args, "stepscale");                      // extension, see tpose_cuda_set_step_scale
    if (!ss.empty()) tpose::check(tpose_cuda_set_step_scale(tpose::context, ss == "auto" ? (450.0f / rh) * (450.0f / rh) : (float)atof(ss.c_str())), "stepscale");

    const auto t0 = std::chrono::steady_clock::now();
    const auto rep = tpose::schedule::warp(ta, tb, frames);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "frames " << rep.frames << ", levels " << rep.levels << ", " << sec << " s, "
              << rep.frames / sec << " steps/s" << std::endl;

    tpose::quit();
    return 0;
}
