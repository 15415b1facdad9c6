// software/triangulate, headless: grows a minimum-energy triangulation of one image and appends the
// levels of the export list to "<image>.tri" (software/triangulate/main.cpp of the reference; the
// window, the key handlers and the mode-2 display are gone, the schedule is tpose/schedule).
//
//   triangulate -i image.ppm [-native] [-max N] [-frames F] [-device D] [-t start.tri]
//     -native   rasterise at the image's own size instead of the reference's 1/1.5 window
//     -max N    extend the export list beyond 1000 in steps of 500 up to N triangles
//     -frames F stop after F optimisation steps
//     -stepscale S|auto  (extension) scale shift.cs's step constant; auto = (450/raster_h)^2
//     -t file   (extension) start from the finest level of an existing .tri stack instead of the
//               two-triangle mesh, and export only levels finer than it.  The reference's step
//               constant (shift.cs:45) is tuned to its 800x450 window: on larger rasters the
//               descent over a handful of huge triangles does not settle, so a
//               large image is started from a stack grown on a reduced copy of itself -- mesh
//               coordinates are resolution-independent.
#include <tpose/tpose>
#include <tpose/triangulation>
#include <tpose/io>
#include <tpose/schedule>

#include <chrono>

#include "image_io.h"

int main(int argc, char* args[])
{
    const std::string path = option(argc, args, "i");
    if (path.empty()) { std::cout << "Please specify an input image with -i." << std::endl; return 0; }
    Image img;
    if (!load_image(path, img)) { std::cout << "Failed to load image." << std::endl; return 0; }

    std::vector<int> exportlist = tpose::schedule::reference_exportlist();      // main.cpp:19-31
    const int upto = atoi(option(argc, args, "max", "1000").c_str());
    for (int n = 1500; n <= upto; n += 500) exportlist.insert(exportlist.begin(), n);
    const long frames = atol(option(argc, args, "frames", "-1").c_str());

    const bool native = has_flag(argc, args, "native");
    const int rw = native ? img.w : (int)(img.w / 1.5), rh = native ? img.h : (int)(img.h / 1.5);   // main.cpp:53
    tpose::RATIO = (float)img.w / (float)img.h;                                  // main.cpp:54

    tpose::init(atoi(option(argc, args, "device", "0").c_str()));               // main.cpp:78
    tpose::texture(TPOSE_IMAGE_A, img.rgba.data(), img.w, img.h, rw, rh);        // main.cpp:74
    const std::string ss = option(argc, args, "stepscale");                      // extension, see tpose_cuda_set_step_scale
    if (!ss.empty()) tpose::check(tpose_cuda_set_step_scale(tpose::context, ss == "auto" ? (450.0f / rh) * (450.0f / rh) : (float)atof(ss.c_str())), "stepscale");

    tpose::triangulation tr;                                                     // main.cpp:105-106
    const std::string start = option(argc, args, "t");
    if (!start.empty()) {
        const float ratio = tpose::RATIO;
        while (tpose::io::read(&tr, start)) {}                                   // the last record read stays in tr
        tpose::RATIO = ratio;                                                    // io::read overwrote it (io.hpp:81)
        // io::read leaves `colors` at this level's NT (io.hpp:139-140); the schedule reads MAXT-capacity
        // colour rows back into it as the mesh grows (triangulate/main.cpp:221-224)
        tr.colors.resize(tpose::triangulation::MAXT, glm::ivec4(0, 0, 0, 0));
        while (!exportlist.empty() && exportlist.back() <= tr.NT) exportlist.pop_back();
    }
    tpose::upload(&tr, false);
    std::cout << "Number of Triangles: " << tr.NT << std::endl;
    tpose::computecolors();                                                      // main.cpp:176

    const auto t0 = std::chrono::steady_clock::now();
    const auto rep = tpose::schedule::triangulate(tr, exportlist, path + ".tri", frames);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "frames " << rep.frames << " (host body on " << rep.host_frames << "), splits " << rep.splits
              << ", flips " << rep.flips << ", prunes " << rep.prunes << ", collapses " << rep.collapses << ", idle host frames "
              << rep.idle_host_frames << ", exports " << rep.exports << ", " << tr.NT << " triangles, " << sec << " s ("
              << rep.s_device << " in tpose_cuda_run, " << rep.s_host << " in the host body), " << rep.frames / sec << " steps/s" << std::endl;
    std::cout << "runs " << rep.runs << ": " << rep.frames_asked << " frames asked, " << rep.frames_run << " run, " << rep.frames << " kept; "
              << rep.resyncs << " runs cut short by a host edit; device-timed " << rep.ms_device_events / 1000.0 << " s; host body: "
              << rep.s_fetch << " s recorded frame -> mesh, " << rep.s_edits << " per-frame edits, " << rep.s_converged << " converged frames, "
              << rep.s_resync << " uploads after an edit" << std::endl;

    tpose::quit();
    return 0;
}
