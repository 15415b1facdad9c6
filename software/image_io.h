// image_io.h -- minimal image input for the ported programs.  The reference loads PNG/JPG through
// SDL_image (software/triangulate/main.cpp:42); neither SDL nor libpng exists in this environment,
// so the programs here read binary PPM (P6, RGB) or PAM (P7, RGB_ALPHA) -- `python -c "from PIL
// import Image; Image.open('x.png').convert('RGB').save('x.ppm')"` converts anything else.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

struct Image {
    int w = 0, h = 0;
    std::vector<unsigned char> rgba;   // row 0 = top
};

inline bool load_image(const std::string& path, Image& img)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[3] = {0, 0, 0};
    if (fscanf(f, "%2s", magic) != 1) { fclose(f); return false; }
    int depth = 0, maxval = 0;
    auto skip_comments = [&] {
        int c;
        while ((c = fgetc(f)) != EOF) {
            if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} }
            else if (c != ' ' && c != '\n' && c != '\r' && c != '\t') { ungetc(c, f); break; }
        }
    };
    if (!strcmp(magic, "P6")) {
        skip_comments(); if (fscanf(f, "%d", &img.w) != 1) { fclose(f); return false; }
        skip_comments(); if (fscanf(f, "%d", &img.h) != 1) { fclose(f); return false; }
        skip_comments(); if (fscanf(f, "%d", &maxval) != 1) { fclose(f); return false; }
        fgetc(f);
        depth = 3;
    } else if (!strcmp(magic, "P7")) {
        char key[64];
        while (fscanf(f, "%63s", key) == 1) {
            if (!strcmp(key, "ENDHDR")) { fgetc(f); break; }
            if (!strcmp(key, "WIDTH")) { if (fscanf(f, "%d", &img.w) != 1) break; }
            else if (!strcmp(key, "HEIGHT")) { if (fscanf(f, "%d", &img.h) != 1) break; }
            else if (!strcmp(key, "DEPTH")) { if (fscanf(f, "%d", &depth) != 1) break; }
            else if (!strcmp(key, "MAXVAL")) { if (fscanf(f, "%d", &maxval) != 1) break; }
            else { char rest[64]; if (fscanf(f, "%63s", rest) != 1) break; }
        }
    } else { fclose(f); return false; }
    if (img.w <= 0 || img.h <= 0 || maxval != 255 || (depth != 3 && depth != 4)) { fclose(f); return false; }
    std::vector<unsigned char> raw((size_t)img.w * img.h * depth);
    const bool ok = fread(raw.data(), 1, raw.size(), f) == raw.size();
    fclose(f);
    if (!ok) return false;
    img.rgba.resize((size_t)img.w * img.h * 4);
    for (size_t p = 0; p < (size_t)img.w * img.h; p++) {
        img.rgba[4 * p + 0] = raw[depth * p + 0];
        img.rgba[4 * p + 1] = raw[depth * p + 1];
        img.rgba[4 * p + 2] = raw[depth * p + 2];
        img.rgba[4 * p + 3] = depth == 4 ? raw[depth * p + 3] : 255;
    }
    return true;
}

// "-key value" options like TinyEngine's parse::get (software/triangulate/main.cpp:35)
inline std::string option(int argc, char** argv, const char* key, const char* fallback = "")
{
    for (int i = 1; i + 1 < argc; i++)
        if (argv[i][0] == '-' && !strcmp(argv[i] + 1, key)) return argv[i + 1];
    return fallback;
}
inline bool has_flag(int argc, char** argv, const char* key)
{
    for (int i = 1; i < argc; i++)
        if (argv[i][0] == '-' && !strcmp(argv[i] + 1, key)) return true;
    return false;
}
