// software/view, headless: draws every level of a .tri stack as flat-coloured triangles (what the
// reference's viewer shows in its window, view/main.cpp:104-121) and writes one PPM per level.
//
//   view -t mesh.tri [-o prefix] [-height 600] [-s 0.0] [-n] [-device D]
//     -n         wireframe on top (the viewer's `n` key, view/main.cpp:58-69)
//     -height H  raster height; the width is RATIO*H like the reference's window (view/main.cpp:44)
//     -s S       interpolation between points (0) and originpoints (1), the viewer's animated `s`
#include <tpose/tpose>
#include <tpose/triangulation>
#include <tpose/io>

#include <cstdio>

#include "image_io.h"

int main(int argc, char* args[])
{
    const std::string tri = option(argc, args, "t");
    if (tri.empty()) { std::cout << "Please specify a triangulation with -t." << std::endl; return 0; }
    const std::string prefix = option(argc, args, "o", tri.c_str());
    const int H = atoi(option(argc, args, "height", "600").c_str());
    const float s = (float)atof(option(argc, args, "s", "0").c_str());

    tpose::init(atoi(option(argc, args, "device", "0").c_str()));
    tpose::triangulation tr;
    int level = 0;
    while (tpose::io::read(&tr, tri)) {                                          // view/main.cpp:37,84
        const int W = (int)(tpose::RATIO * H);                                   // view/main.cpp:44
        tpose::upload(&tr);                                                      // colours into tcolaccbuf
        std::vector<unsigned char> rgba((size_t)W * H * 4);
        tpose::check((has_flag(argc, args, "n") ? tpose_cuda_render_wire : tpose_cuda_render)(tpose::context, W, H, &tr.originpoints[0].x, s, rgba.data()), "render");
        const std::string out = prefix + "." + std::to_string(level++) + ".ppm";
        FILE* f = fopen(out.c_str(), "wb");
        if (!f) { std::cout << "cannot write " << out << std::endl; return 1; }
        fprintf(f, "P6\n%d %d\n255\n", W, H);
        for (size_t p = 0; p < (size_t)W * H; p++) fwrite(&rgba[4 * p], 1, 3, f);
        fclose(f);
        std::cout << "level " << level - 1 << ": " << tr.NT << " triangles -> " << out << std::endl;
    }
    tpose::quit();
    return 0;
}
