// ref_capi.cpp -- C wrappers around the reference's OWN headers (source/tpose.hpp, utility.hpp,
// triangulation.hpp, io.hpp), which are included unmodified from /root/reference through the
// symlinks the Makefile creates under oracle/_ref/include/tpose/.  The resulting
// oracle/_ref/libtpose_ref.so is the authority for topology edits, barycentric warps, geterr /
// maxerrid and .tri bytes in the parity tests.  TEST INFRASTRUCTURE ONLY.
#include <glm/glm.hpp>
#include <TinyEngine/TinyEngine>
#include <sstream>

#include <tpose/tpose>
#include <tpose/triangulation>
#include <tpose/io>

// Eigenvalues of a real upper-Hessenberg-or-general matrix by the shifted QR iteration
// (reduction to Hessenberg form, then Francis double-shift steps; real eigenvalues come out with
// an exactly zero imaginary part, which is what utility.hpp:128 tests for).
namespace Eigen {
VectorXcd tpose_shim_eigenvalues(const MatrixXd& Ain)
{
    const int n = Ain.r;
    std::vector<std::vector<double>> H(n, std::vector<double>(n));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) H[i][j] = Ain(i, j);
    // Hessenberg reduction by stabilised elementary similarity transformations
    for (int m = 1; m < n - 1; m++) {
        double x = 0.0; int i = m;
        for (int j = m; j < n; j++) if (std::fabs(H[j][m - 1]) > std::fabs(x)) { x = H[j][m - 1]; i = j; }
        if (i != m) {
            for (int j = m - 1; j < n; j++) std::swap(H[i][j], H[m][j]);
            for (int j = 0; j < n; j++) std::swap(H[j][i], H[j][m]);
        }
        if (x != 0.0) {
            for (i = m + 1; i < n; i++) {
                double y = H[i][m - 1];
                if (y != 0.0) {
                    y /= x; H[i][m - 1] = y;
                    for (int j = m; j < n; j++) H[i][j] -= y * H[m][j];
                    for (int j = 0; j < n; j++) H[j][m] += y * H[j][i];
                }
            }
        }
    }
    for (int i = 2; i < n; i++) for (int j = 0; j < i - 1; j++) H[i][j] = 0.0;
    std::vector<double> wr(n, 0.0), wi(n, 0.0);
    double anorm = 0.0;
    for (int i = 0; i < n; i++) for (int j = std::max(i - 1, 0); j < n; j++) anorm += std::fabs(H[i][j]);
    int nn = n - 1; double t = 0.0;
    while (nn >= 0) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; l--) {
                double s = std::fabs(H[l - 1][l - 1]) + std::fabs(H[l][l]);
                if (s == 0.0) s = anorm;
                if (std::fabs(H[l][l - 1]) + s == s) { H[l][l - 1] = 0.0; break; }
            }
            double x = H[nn][nn];
            if (l == nn) { wr[nn] = x + t; wi[nn--] = 0.0; }
            else {
                double y = H[nn - 1][nn - 1], w = H[nn][nn - 1] * H[nn - 1][nn];
                if (l == nn - 1) {
                    double p = 0.5 * (y - x), q = p * p + w, z = std::sqrt(std::fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + (p >= 0.0 ? std::fabs(z) : -std::fabs(z));
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = -(wi[nn] = z);
                    }
                    nn -= 2;
                } else {
                    if (its == 60) std::abort();
                    if (its == 10 || its == 20) {
                        t += x;
                        for (int i = 0; i <= nn; i++) H[i][i] -= x;
                        double s = std::fabs(H[nn][nn - 1]) + std::fabs(H[nn - 1][nn - 2]);
                        y = x = 0.75 * s; w = -0.4375 * s * s;
                    }
                    ++its;
                    int m; double p = 0, q = 0, r = 0, z;
                    for (m = nn - 2; m >= l; m--) {
                        z = H[m][m];
                        double rr = x - z, ss = y - z;
                        p = (rr * ss - w) / H[m + 1][m] + H[m][m + 1];
                        q = H[m + 1][m + 1] - z - rr - ss;
                        r = H[m + 2][m + 1];
                        double s = std::fabs(p) + std::fabs(q) + std::fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        double u = std::fabs(H[m][m - 1]) * (std::fabs(q) + std::fabs(r));
                        double v = std::fabs(p) * (std::fabs(H[m - 1][m - 1]) + std::fabs(z) + std::fabs(H[m + 1][m + 1]));
                        if (u + v == v) break;
                    }
                    for (int i = m + 2; i <= nn; i++) { H[i][i - 2] = 0.0; if (i != m + 2) H[i][i - 3] = 0.0; }
                    for (int k = m; k <= nn - 1; k++) {
                        if (k != m) {
                            p = H[k][k - 1]; q = H[k + 1][k - 1]; r = 0.0;
                            if (k != nn - 1) r = H[k + 2][k - 1];
                            if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        double s = std::sqrt(p * p + q * q + r * r);
                        if (p < 0) s = -s;
                        if (s != 0.0) {
                            if (k == m) { if (l != m) H[k][k - 1] = -H[k][k - 1]; }
                            else H[k][k - 1] = -s * x;
                            p += s; x = p / s; y = q / s; z = r / s; q /= p; r /= p;
                            for (int j = k; j <= nn; j++) {
                                p = H[k][j] + q * H[k + 1][j];
                                if (k != nn - 1) { p += r * H[k + 2][j]; H[k + 2][j] -= p * z; }
                                H[k + 1][j] -= p * y; H[k][j] -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; i++) {
                                p = x * H[i][k] + y * H[i][k + 1];
                                if (k != nn - 1) { p += z * H[i][k + 2]; H[i][k + 2] -= p * r; }
                                H[i][k + 1] -= p * q; H[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    VectorXcd out(n, 1);
    for (int i = 0; i < n; i++) out(i) = std::complex<double>(wr[i], wi[i]);
    return out;
}
}  // namespace Eigen

using tpose::triangulation;

extern "C" {

void ref_set_ratio(float r) { tpose::RATIO = r; }
float ref_get_ratio() { return tpose::RATIO; }
void ref_init() { if (!tpose::terr) tpose::init(); }

void* ref_tri_new() { return new triangulation(); }
void ref_tri_free(void* h) { delete (triangulation*)h; }
void ref_tri_counts(void* h, int* nt, int* np) { auto* t = (triangulation*)h; *nt = t->NT; *np = t->NP; }
// sizes of the containers themselves (NT/NP are separate members the reference keeps by hand)
void ref_tri_sizes(void* h, int* ntri, int* nhe, int* npts, int* nopts)
{
    auto* t = (triangulation*)h;
    *ntri = (int)t->triangles.size(); *nhe = (int)t->halfedges.size();
    *npts = (int)t->points.size(); *nopts = (int)t->originpoints.size();
}
void ref_tri_get(void* h, int* tri4, int* he, int* col4, float* pts, float* opts)
{
    auto* t = (triangulation*)h;
    if (tri4) std::memcpy(tri4, t->triangles.data(), sizeof(glm::ivec4) * t->triangles.size());
    if (he) std::memcpy(he, t->halfedges.data(), sizeof(int) * t->halfedges.size());
    if (col4) std::memcpy(col4, t->colors.data(), sizeof(glm::ivec4) * t->triangles.size());
    if (pts) std::memcpy(pts, t->points.data(), sizeof(glm::vec2) * t->points.size());
    if (opts) std::memcpy(opts, t->originpoints.data(), sizeof(glm::vec2) * t->originpoints.size());
}
void ref_tri_set(void* h, int nt, const int* tri4, const int* he, const int* col4, int np, const float* pts, const float* opts)
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
void ref_tri_set_points(void* h, const float* pts)
{
    auto* t = (triangulation*)h;
    std::memcpy(t->points.data(), pts, sizeof(glm::vec2) * t->points.size());
}
void ref_tri_origin_from_points(void* h) { auto* t = (triangulation*)h; t->originpoints = t->points; }

float ref_tri_angle(void* h, int ha) { return ((triangulation*)h)->angle(ha); }
float ref_tri_hlength(void* h, int ha) { return ((triangulation*)h)->hlength(ha); }
int ref_tri_boundary(void* h, int t) { return ((triangulation*)h)->boundary(t); }
int ref_point_boundary(float x, float y) { return triangulation::boundary(glm::vec2(x, y)) ? 1 : 0; }
int ref_tri_eraset(void* h, int t, int adjust) { return ((triangulation*)h)->eraset(t, adjust != 0); }
int ref_tri_erasep(void* h, int p) { return ((triangulation*)h)->erasep(p); }
int ref_tri_prune(void* h, int t) { return ((triangulation*)h)->prune(t); }
int ref_tri_flip(void* h, int ha, float minangle) { return ((triangulation*)h)->flip(ha, minangle); }
int ref_tri_flip_default(void* h, int ha) { return ((triangulation*)h)->flip(ha); }
int ref_tri_collapse(void* h, int ha) { return ((triangulation*)h)->collapse(ha); }
int ref_tri_split(void* h, int t) { return ((triangulation*)h)->split(t); }
int ref_tri_optimize(void* h) { return ((triangulation*)h)->optimize(); }
void ref_tri_warp(void* h, float* pts, int n)
{
    std::vector<glm::vec2> v(n);
    std::memcpy(v.data(), pts, sizeof(glm::vec2) * n);
    ((triangulation*)h)->warp(v);
    std::memcpy(pts, v.data(), sizeof(glm::vec2) * n);
}
void ref_tri_reversewarp(void* h, float* pts, int n)
{
    std::vector<glm::vec2> v(n);
    std::memcpy(v.data(), pts, sizeof(glm::vec2) * n);
    ((triangulation*)h)->reversewarp(v);
    std::memcpy(pts, v.data(), sizeof(glm::vec2) * n);
}

// source/utility.hpp
void ref_barycentric(float px, float py, const int* t4, const float* verts, int nv, float* out3)
{
    std::vector<glm::vec2> v(nv);
    std::memcpy(v.data(), verts, sizeof(glm::vec2) * nv);
    glm::vec3 s = tpose::barycentric(glm::vec2(px, py), glm::ivec4(t4[0], t4[1], t4[2], t4[3]), v);
    out3[0] = s.x; out3[1] = s.y; out3[2] = s.z;
}
int ref_intriangle(float px, float py, const int* t4, const float* verts, int nv)
{
    std::vector<glm::vec2> v(nv);
    std::memcpy(v.data(), verts, sizeof(glm::vec2) * nv);
    return tpose::intriangle(glm::vec2(px, py), glm::ivec4(t4[0], t4[1], t4[2], t4[3]), v) ? 1 : 0;
}
int ref_realroots(const double* a7, double* out6)
{
    std::vector<double> a(a7, a7 + 7);
    std::vector<double> r = tpose::realroots(a);
    for (size_t i = 0; i < r.size(); i++) out6[i] = r[i];
    return (int)r.size();
}
double ref_horner(double x, const double* a, int n) { return tpose::horner(x, std::vector<double>(a, a + n)); }

// source/io.hpp (messages on stdout are silenced)
static std::ostringstream g_sink;
struct Quiet { std::streambuf* old; Quiet() : old(std::cout.rdbuf(g_sink.rdbuf())) {} ~Quiet() { std::cout.rdbuf(old); g_sink.str(""); } };
void ref_io_write(void* h, const char* path) { Quiet q; tpose::io::write((triangulation*)h, path); ((triangulation*)h)->out.flush(); }
int ref_io_read(void* h, const char* path, int dowarp) { Quiet q; return tpose::io::read((triangulation*)h, path, dowarp != 0) ? 1 : 0; }
void ref_io_close(void* h) { auto* t = (triangulation*)h; if (t->out.is_open()) t->out.close(); if (t->in.is_open()) t->in.close(); }
int ref_io_readmatches(const char* path, float* a, float* b, int cap)
{
    Quiet q;
    std::vector<glm::vec2> A, B;
    if (!tpose::io::readmatches(path, A, B)) return -1;
    const int n = (int)std::min<size_t>(A.size(), cap);
    for (int i = 0; i < n; i++) { a[2 * i] = A[i].x; a[2 * i + 1] = A[i].y; b[2 * i] = B[i].x; b[2 * i + 1] = B[i].y; }
    return (int)A.size();
}

// source/triangulation.hpp:578-719: globals, upload, error reductions
void ref_set_terr(const int* v, int n) { ref_init(); std::memcpy(tpose::terr, v, sizeof(int) * n); }
void ref_set_cn(const int* v, int n) { ref_init(); std::memcpy(tpose::cn, v, sizeof(int) * n); }
float ref_geterr(void* h) { return tpose::geterr((triangulation*)h); }
float ref_gettoterr(void* h) { return tpose::gettoterr((triangulation*)h); }
int ref_maxerrid(void* h) { return tpose::maxerrid((triangulation*)h); }
float ref_get_toterr() { return tpose::toterr; }
void ref_set_toterr(float v) { tpose::toterr = v; }
float ref_get_relerr() { return tpose::relerr; }
// upload(tr, uploadcolor): returns what went into the point / index / colour SSBOs
void ref_upload(void* h, int uploadcolor, float* pts_out, int* tri_out, int* col13_out)
{
    ref_init();
    auto* t = (triangulation*)h;
    tpose::upload(t, uploadcolor != 0);
    if (pts_out) tpose::pointbuf->retrieve(t->points.size(), (glm::vec2*)pts_out);
    if (tri_out) tpose::trianglebuf->retrieve(t->triangles.size(), (glm::ivec4*)tri_out);
    if (col13_out && uploadcolor) tpose::tcolaccbuf->retrieve(13 * t->NT, (glm::ivec4*)col13_out);
}

}  // extern "C"
