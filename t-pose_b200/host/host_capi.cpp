// host_capi.cpp -- C entry points over the C++ host layer (include/tpose/*), so that the Python
// test-suite and bench can drive the product's own tpose::triangulation / tpose::io /
// tpose::schedule code.  Mirrors, name for name, the wrapper the tests put around the reference's
// unmodified headers (oracle/ref_shim/ref_capi.cpp: ref_* there, tph_* here).
#include <tpose/tpose>
#include <tpose/triangulation>
#include <tpose/io>
#include <tpose/schedule>
#include <tpose/detail/edgeview.h>
#include <tpose/multiview>
#include <tpose/pairs>

#include <sstream>

using tpose::triangulation;

namespace {
std::ostringstream g_sink;
struct Quiet {
    std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(g_sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); g_sink.str(""); }
};
}  // namespace

extern "C" {

void tph_set_ratio(float r) { tpose::RATIO = r; }
float tph_get_ratio() { return tpose::RATIO; }

void* tph_tri_new() { return new triangulation(); }
void tph_tri_free(void* h) { delete (triangulation*)h; }
void tph_tri_counts(void* h, int* nt, int* np) { auto* t = (triangulation*)h; *nt = t->NT; *np = t->NP; }
void tph_tri_sizes(void* h, int* ntri, int* nhe, int* npts, int* nopts)
{
    auto* t = (triangulation*)h;
    *ntri = (int)t->triangles.size(); *nhe = (int)t->halfedges.size();
    *npts = (int)t->points.size(); *nopts = (int)t->originpoints.size();
}
void tph_tri_get(void* h, int* tri4, int* he, int* col4, float* pts, float* opts)
{
    auto* t = (triangulation*)h;
    if (tri4) std::memcpy(tri4, t->triangles.data(), sizeof(glm::ivec4) * t->triangles.size());
    if (he) std::memcpy(he, t->halfedges.data(), sizeof(int) * t->halfedges.size());
    if (col4) std::memcpy(col4, t->colors.data(), sizeof(glm::ivec4) * t->triangles.size());
    if (pts) std::memcpy(pts, t->points.data(), sizeof(glm::vec2) * t->points.size());
    if (opts) std::memcpy(opts, t->originpoints.data(), sizeof(glm::vec2) * t->originpoints.size());
}
void tph_tri_set(void* h, int nt, const int* tri4, const int* he, const int* col4, int np, const float* pts, const float* opts)
{
    auto* t = (triangulation*)h;
    t->NT = nt; t->NP = np;
    t->triangles.resize(nt); t->halfedges.resize(3 * nt);
    std::memcpy(t->triangles.data(), tri4, sizeof(glm::ivec4) * nt);
    std::memcpy(t->halfedges.data(), he, sizeof(int) * 3 * nt);
    if (col4) std::memcpy(t->colors.data(), col4, sizeof(glm::ivec4) * nt);
    t->points.resize(np);
    std::memcpy(t->points.data(), pts, sizeof(glm::vec2) * np);
    if (opts) { t->originpoints.resize(np); std::memcpy(t->originpoints.data(), opts, sizeof(glm::vec2) * np); }
}
void tph_tri_set_points(void* h, const float* pts)
{
    auto* t = (triangulation*)h;
    std::memcpy(t->points.data(), pts, sizeof(glm::vec2) * t->points.size());
}
void tph_tri_origin_from_points(void* h) { auto* t = (triangulation*)h; t->originpoints = t->points; }

float tph_tri_angle(void* h, int ha) { return ((triangulation*)h)->angle(ha); }
float tph_tri_hlength(void* h, int ha) { return ((triangulation*)h)->hlength(ha); }
int tph_tri_boundary(void* h, int t) { return ((triangulation*)h)->boundary(t); }
int tph_point_boundary(float x, float y) { return triangulation::boundary(glm::vec2(x, y)) ? 1 : 0; }
int tph_tri_eraset(void* h, int t, int adjust) { return ((triangulation*)h)->eraset(t, adjust != 0); }
int tph_tri_erasep(void* h, int p) { return ((triangulation*)h)->erasep(p); }
int tph_tri_prune(void* h, int t) { return ((triangulation*)h)->prune(t); }
int tph_tri_flip(void* h, int ha, float minangle) { return ((triangulation*)h)->flip(ha, minangle); }
int tph_tri_flip_default(void* h, int ha) { return ((triangulation*)h)->flip(ha); }
int tph_tri_collapse(void* h, int ha) { return ((triangulation*)h)->collapse(ha); }
int tph_tri_split(void* h, int t) { return ((triangulation*)h)->split(t); }
int tph_tri_optimize(void* h) { return ((triangulation*)h)->optimize(); }
void tph_tri_warp(void* h, float* pts, int n)
{
    std::vector<glm::vec2> v(n);
    std::memcpy(v.data(), pts, sizeof(glm::vec2) * n);
    ((triangulation*)h)->warp(v);
    std::memcpy(pts, v.data(), sizeof(glm::vec2) * n);
}
void tph_tri_reversewarp(void* h, float* pts, int n)
{
    std::vector<glm::vec2> v(n);
    std::memcpy(v.data(), pts, sizeof(glm::vec2) * n);
    ((triangulation*)h)->reversewarp(v);
    std::memcpy(pts, v.data(), sizeof(glm::vec2) * n);
}

// tpose/utility
void tph_barycentric(float px, float py, const int* t4, const float* verts, int nv, float* out3)
{
    std::vector<glm::vec2> v(nv);
    std::memcpy(v.data(), verts, sizeof(glm::vec2) * nv);
    glm::vec3 s = tpose::barycentric(glm::vec2(px, py), glm::ivec4(t4[0], t4[1], t4[2], t4[3]), v);
    out3[0] = s.x; out3[1] = s.y; out3[2] = s.z;
}
int tph_intriangle(float px, float py, const int* t4, const float* verts, int nv)
{
    std::vector<glm::vec2> v(nv);
    std::memcpy(v.data(), verts, sizeof(glm::vec2) * nv);
    return tpose::intriangle(glm::vec2(px, py), glm::ivec4(t4[0], t4[1], t4[2], t4[3]), v) ? 1 : 0;
}
int tph_realroots(const double* a7, double* out6)
{
    std::vector<double> a(a7, a7 + 7);
    std::vector<double> r = tpose::realroots(a);
    for (size_t i = 0; i < r.size(); i++) out6[i] = r[i];
    return (int)r.size();
}
double tph_horner(double x, const double* a, int n) { return tpose::horner(x, std::vector<double>(a, a + n)); }

// tpose/io
void tph_io_write(void* h, const char* path) { Quiet q; tpose::io::write((triangulation*)h, path); ((triangulation*)h)->out.flush(); }
int tph_io_read(void* h, const char* path, int dowarp) { Quiet q; return tpose::io::read((triangulation*)h, path, dowarp != 0) ? 1 : 0; }
void tph_io_close(void* h) { auto* t = (triangulation*)h; if (t->out.is_open()) t->out.close(); if (t->in.is_open()) t->in.close(); }
int tph_io_readmatches(const char* path, float* a, float* b, int cap)
{
    Quiet q;
    std::vector<glm::vec2> A, B;
    if (!tpose::io::readmatches(path, A, B)) return -1;
    const int n = (int)std::min<size_t>(A.size(), cap);
    for (int i = 0; i < n; i++) { a[2 * i] = A[i].x; a[2 * i + 1] = A[i].y; b[2 * i] = B[i].x; b[2 * i + 1] = B[i].y; }
    return (int)A.size();
}

// error reductions over host mirrors (no device needed: the mirrors are plain arrays)
static void ensure_mirrors()
{
    if (tpose::terr) return;
    tpose::terr = new int[triangulation::MAXT]();
    tpose::perr = new int[triangulation::MAXT]();
    tpose::cn = new int[triangulation::MAXT]();
    tpose::col = new glm::ivec4[triangulation::MAXT]();
}
void tph_set_terr(const int* v, int n) { ensure_mirrors(); std::memcpy(tpose::terr, v, sizeof(int) * n); }
void tph_set_cn(const int* v, int n) { ensure_mirrors(); std::memcpy(tpose::cn, v, sizeof(int) * n); }
float tph_geterr(void* h) { return tpose::geterr((triangulation*)h); }
float tph_gettoterr(void* h) { return tpose::gettoterr((triangulation*)h); }
int tph_maxerrid(void* h) { return tpose::maxerrid((triangulation*)h); }
float tph_get_toterr() { return tpose::toterr; }
void tph_set_toterr(float v) { tpose::toterr = v; }
float tph_get_relerr() { return tpose::relerr; }

// ---- device-backed part --------------------------------------------------------------------------

int tph_init(int device)
{
    if (tpose::context) return 0;
    if (tpose::terr) { delete[] tpose::terr; delete[] tpose::perr; delete[] tpose::cn; delete[] tpose::col; tpose::terr = nullptr; }
    tpose::init(device);
    return 0;
}
void tph_quit() { if (tpose::context) tpose::quit(); }
void tph_texture(int slot, const unsigned char* rgba, int w, int h, int rw, int rh) { tpose::texture(slot, rgba, w, h, rw, rh); }
void tph_upload(void* h, int uploadcolor) { tpose::upload((triangulation*)h, uploadcolor != 0); }

// tpose/detail/edgeview.h on its own (host only).  Outputs sized by the caller: edges 6*nt, tri_edges 4*nt,
// corner_off np+1, corners 3*nt.  Returns the number of unique edges, or -1 for a bad vertex index.
int tph_edgeview(const int* tri4, int nt, int np, int* edges, int* tri_edges, int* corner_off, int* corners)
{
    tpose::edgeview::view v; std::vector<std::uint64_t> keys;
    if (!tpose::edgeview::build(tri4, nt, np, v, keys)) return -1;
    std::copy(v.edges.begin(), v.edges.end(), edges);
    std::copy(v.tri_edges.begin(), v.tri_edges.end(), tri_edges);
    std::copy(v.corner_off.begin(), v.corner_off.end(), corner_off);
    std::copy(v.corners.begin(), v.corners.end(), corners);
    return v.ne;
}

// The two acos / sqrt-free pre-filters of the per-frame loops (tpose/schedule, host only)
int tph_very_obtuse(void* h, int ha) { return tpose::schedule::detail::very_obtuse(*(triangulation*)h, ha) ? 1 : 0; }
int tph_collapse_candidate(void* h, int ta) { return tpose::schedule::detail::collapse_candidate(*(triangulation*)h, ta); }

// The flip set of a converged frame (tpose/schedule, host only) from the energies in tpose::terr.
// Returns the number of selected halfedges; he_out / energy_out sized 3*NT by the caller.
int tph_flip_set(void* h, int* he_out, float* energy_out)
{
    ensure_mirrors();
    const auto fs = tpose::schedule::detail::flip_set(*(triangulation*)h);
    for (size_t i = 0; i < fs.size(); i++) { he_out[i] = fs[i].first; energy_out[i] = fs[i].second; }
    return (int)fs.size();
}

// The per-frame edit loops of tpose/schedule on their own (host only).  mask32 == nullptr: the plain
// loops of the reference; otherwise the sparse walk over the flagged triangles + `dirty` (64-bit words,
// in/out).  counts: flips, prunes, collapses.  Returns 1 if the mesh must be uploaded again.
int tph_frame_edits(void* h, const unsigned* mask32, int nwords, unsigned long long* dirty_words, int ndirty_words, long* counts3)
{
    auto* t = (triangulation*)h;
    tpose::schedule::report rep;
    tpose::schedule::detail::trimask cand, dirty;
    dirty.reset(t->NT);
    for (int i = 0; i < ndirty_words && i < (int)dirty.w.size(); i++) dirty.w[(size_t)i] = dirty_words[i];
    bool updated;
    if (mask32) {
        cand.reset(t->NT);
        for (int i = 0; i < nwords && (size_t)(i >> 1) < cand.w.size(); i++) cand.w[(size_t)i >> 1] |= (uint64_t)mask32[i] << (32 * (i & 1));
        updated = tpose::schedule::detail::frame_edits(*t, rep, &cand, dirty);
    } else {
        updated = tpose::schedule::detail::frame_edits(*t, rep, nullptr, dirty);
    }
    for (int i = 0; i < ndirty_words; i++) dirty_words[i] = (size_t)i < dirty.w.size() ? dirty.w[(size_t)i] : 0ull;
    counts3[0] = rep.flips; counts3[1] = rep.prunes; counts3[2] = rep.collapses;
    return updated ? 1 : 0;
}

// report: frames, host_frames, runs, splits, exports, levels, finished
void tph_triangulate(void* h, const int* exportlist, int nexport, const char* trifile, long max_frames, double tol, long* rep7)
{
    auto* t = (triangulation*)h;
    Quiet q;
    tpose::upload(t, false);
    tpose::computecolors();
    std::vector<int> ex(exportlist, exportlist + nexport);
    const auto r = tpose::schedule::triangulate(*t, ex, trifile, max_frames, tol, false);
    t->out.flush();
    rep7[0] = r.frames; rep7[1] = r.host_frames; rep7[2] = r.runs; rep7[3] = r.splits; rep7[4] = r.exports; rep7[5] = r.levels; rep7[6] = r.finished;
}
void tph_warp(const char* triA, const char* triB, long max_frames, double tol, long* rep7)
{
    Quiet q;
    const auto r = tpose::schedule::warp(triA, triB, max_frames, tol, false);
    rep7[0] = r.frames; rep7[1] = r.host_frames; rep7[2] = r.runs; rep7[3] = r.splits; rep7[4] = r.exports; rep7[5] = r.levels; rep7[6] = r.finished;
}


// ---- tpose/pairs: BASELINE configs 4 / 5 in one process per GPU ----------------------------------------
// A set of independent view pairs on one device.  tph_init(device) first (device-side warp / reversewarp).
struct PairSet {
    int device = 0;
    std::vector<std::unique_ptr<tpose::pairs::job>> jobs;
    ~PairSet() { for (auto& j : jobs) if (j->ctx) tpose_cuda_destroy(j->ctx); }
};
void* tph_pairs_new(int device) { auto* s = new PairSet(); s->device = device; return s; }
void tph_pairs_free(void* h) { delete (PairSet*)h; }
// adds a pair: both views' pixels (w x h RGBA8, rasterised at rw x rh) and the two .tri stacks; returns its index or < 0
int tph_pairs_add(void* h, const unsigned char* rgbaA, const unsigned char* rgbaB, int w, int hgt, int rw, int rh,
                  const char* triA, const char* triB, int max_triangles, int max_points)
{
    auto* s = (PairSet*)h;
    auto j = std::make_unique<tpose::pairs::job>();
    if (tpose_cuda_create(s->device, max_triangles, max_points, nullptr, &j->ctx) != TPOSE_OK) return -1;
    if (tpose_cuda_set_image(j->ctx, TPOSE_IMAGE_A, rgbaA, w, hgt, rw, rh) != TPOSE_OK) return -2;
    if (tpose_cuda_set_image(j->ctx, TPOSE_IMAGE_B, rgbaB, w, hgt, rw, rh) != TPOSE_OK) return -3;
    j->triA = triA; j->triB = triB;
    Quiet q;
    tpose::pairs::open(*j);
    s->jobs.push_back(std::move(j));
    return (int)s->jobs.size() - 1;
}
// warp level loop of all pairs, then F + 3-D points for all of them; frames / levels: one entry per pair
void tph_pairs_run(void* h, double tol, long level_cap, int with_mview, long* frames, long* levels)
{
    auto* s = (PairSet*)h;
    std::vector<tpose::pairs::job*> js;
    for (auto& j : s->jobs) js.push_back(j.get());
    Quiet q;
    tpose::pairs::warp_batch(js, tol, level_cap, false);
    if (with_mview) tpose::pairs::fundamental_and_points(js, s->device);
    for (size_t i = 0; i < js.size(); i++) { if (frames) frames[i] = js[i]->frames; if (levels) levels[i] = js[i]->levels; }
}
// wall seconds of the last tph_pairs_run: [batch_run, level switches on the host, F + DLT, runs, frames asked, frames run]
void tph_pairs_timing(double* out6)
{
    const auto& t = tpose::pairs::last_timing;
    out6[0] = t.batch_run; out6[1] = t.level_switch; out6[2] = t.mview; out6[3] = (double)t.runs; out6[4] = (double)t.frames_asked; out6[5] = (double)t.frames_run;
}
// sizes of pair i's results: [NP of A, NP of B, matches]
void tph_pairs_sizes(void* h, int i, int* out3)
{
    auto& j = *((PairSet*)h)->jobs[(size_t)i];
    out3[0] = (int)j.trA.points.size(); out3[1] = (int)j.trB.points.size(); out3[2] = (int)j.X.size();
}
// pair i's results: finest level's vertex buffers (points | originpoints of A, then of B), F, the 3-D vertices
void tph_pairs_result(void* h, int i, float* ptsA, float* optsA, float* ptsB, float* optsB, float* F9, float* X4)
{
    auto& j = *((PairSet*)h)->jobs[(size_t)i];
    if (ptsA) std::memcpy(ptsA, j.trA.points.data(), sizeof(glm::vec2) * j.trA.points.size());
    if (optsA) std::memcpy(optsA, j.trA.originpoints.data(), sizeof(glm::vec2) * j.trA.originpoints.size());
    if (ptsB) std::memcpy(ptsB, j.trB.points.data(), sizeof(glm::vec2) * j.trB.points.size());
    if (optsB) std::memcpy(optsB, j.trB.originpoints.data(), sizeof(glm::vec2) * j.trB.originpoints.size());
    if (F9) std::memcpy(F9, j.F.m, sizeof j.F.m);
    if (X4 && !j.X.empty()) std::memcpy(X4, j.X.data(), sizeof(glm::vec4) * j.X.size());
}

// ---- tpose/multiview ---------------------------------------------------------------------------------

void tph_mview_device(int d) { tpose::mview::device = d; }
void tph_camera(float* K9) { auto K = tpose::mview::Camera(); std::memcpy(K9, K.m, sizeof K.m); }
void tph_normalize(float* pts, int n, float* H9)
{
    std::vector<glm::vec2> v(n);
    std::memcpy(v.data(), pts, sizeof(glm::vec2) * n);
    auto H = tpose::mview::normalize(v);
    std::memcpy(pts, v.data(), sizeof(glm::vec2) * n);
    std::memcpy(H9, H.m, sizeof H.m);
}
static std::vector<glm::vec2> vec2s(const float* p, int n)
{
    std::vector<glm::vec2> v(n);
    if (n) std::memcpy(v.data(), p, sizeof(glm::vec2) * n);
    return v;
}
void tph_f8point(const float* A, int na, const float* B, int nb, float* F9)
{
    Quiet q;
    auto F = tpose::mview::F_8Point(vec2s(A, na), vec2s(B, nb));
    std::memcpy(F9, F.m, sizeof F.m);
}
void tph_fsampson(const float* A, int na, const float* B, int nb, const float* w, float* F9)
{
    auto F = w ? tpose::mview::F_Sampson(vec2s(A, na), vec2s(B, nb), std::vector<float>(w, w + na))
               : tpose::mview::F_Sampson(vec2s(A, na), vec2s(B, nb));
    std::memcpy(F9, F.m, sizeof F.m);
}
// tpose::mview::select_matches over plain arrays (a level of A / B and of the other direction's stacks)
void tph_select_matches(const float* originA, const float* pointsA, const float* worigA, int npa, const float* pointsB,
                        const float* originB, const float* worigB, int npb, float ratio, double max_dist, float* X, float* Y,
                        float* W, int* n_out)
{
    struct Lvl { std::vector<glm::vec2> points, originpoints; };
    Lvl a, wa, b, wb;
    a.originpoints = vec2s(originA, npa); a.points = vec2s(pointsA, npa); wa.originpoints = vec2s(worigA, npa);
    b.originpoints = vec2s(originB, npb); b.points = vec2s(pointsB, npb); wb.originpoints = vec2s(worigB, npb);
    std::vector<glm::vec2> mx, my; std::vector<float> w;
    tpose::mview::select_matches(a, wa, b, wb, ratio, mx, my, w, max_dist);
    *n_out = (int)mx.size();
    if (!mx.empty()) {
        std::memcpy(X, mx.data(), sizeof(glm::vec2) * mx.size());
        std::memcpy(Y, my.data(), sizeof(glm::vec2) * my.size());
        std::memcpy(W, w.data(), sizeof(float) * w.size());
    }
}
void tph_triangulate_points(const float* F9, const float* K9, const float* A, const float* B, int n, float* X4)
{
    tpose::mview::Matrix3f F, K;
    std::memcpy(F.m, F9, sizeof F.m); std::memcpy(K.m, K9, sizeof K.m);
    auto X = tpose::mview::triangulate(F, K, vec2s(A, n), vec2s(B, n));
    std::memcpy(X4, X.data(), sizeof(glm::vec4) * n);
}
void tph_correct_match(const float* F9, float* a2, float* b2)
{
    tpose::mview::Matrix3f F; std::memcpy(F.m, F9, sizeof F.m);
    glm::vec2 a(a2[0], a2[1]), b(b2[0], b2[1]);
    tpose::mview::triangulate(F, a, b);
    a2[0] = a.x; a2[1] = a.y; b2[0] = b.x; b2[1] = b.y;
}
void tph_hdlt(const float* PA12, const float* PB12, const float* a2, const float* b2, float* X4)
{
    tpose::mview::Matrix34f PA, PB; std::memcpy(PA.m, PA12, sizeof PA.m); std::memcpy(PB.m, PB12, sizeof PB.m);
    tpose::mview::Vector3f xa, xb; xa(0) = a2[0]; xa(1) = a2[1]; xa(2) = 1; xb(0) = b2[0]; xb(1) = b2[1]; xb(2) = 1;
    auto X = tpose::mview::HDLT(PA, PB, xa, xb);
    std::memcpy(X4, X.m, sizeof X.m);
}
void tph_getpose(const float* E9, float* R1, float* R2, float* t3)
{
    tpose::mview::Matrix3f E; std::memcpy(E.m, E9, sizeof E.m);
    auto P = tpose::mview::GetPose(E);
    std::memcpy(R1, P.R1.m, 36); std::memcpy(R2, P.R2.m, 36); std::memcpy(t3, P.t.m, 12);
}
void tph_epole(const float* F9, int right, float* e2)
{
    tpose::mview::Matrix3f F; std::memcpy(F.m, F9, sizeof F.m);
    auto e = tpose::mview::epole(F, right != 0); e2[0] = e.x; e2[1] = e.y;
}
// the shared host/device math, callable without a GPU (CPU tests pin it against the numpy oracle)
int tph_sm_correct_match(const double* F9, double* a2, double* b2) { return tpose::sm::correct_match(F9, a2, b2); }
void tph_sm_hdlt(const double* PA12, const double* PB12, const double* a2, const double* b2, double* X4) { tpose::sm::hdlt(PA12, PB12, a2, b2, X4); }
void tph_sm_get_pose(const double* E9, double* R1, double* R2, double* t3) { tpose::sm::get_pose(E9, R1, R2, t3); }
void tph_sm_svd3(const double* M9, double* U9, double* S3, double* V9) { tpose::sm::svd3_proper(M9, U9, S3, V9); }
void tph_sm_rank2(double* F9) { tpose::sm::rank2_project(F9); }
int tph_sm_realroots(const double* a7, double* out6) { return tpose::sm::realroots_sorted(a7, out6); }
void tph_sm_jacobi9(double* A81, double* V81) { tpose::sm::jacobi_eigen<9>(A81, V81); }

}  // extern "C"
