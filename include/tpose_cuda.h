/*
 * tpose_cuda.h -- C ABI of the B200 (sm_100a) implementation of t-pose's
 * data-parallel hot path.  Plain pointers and sizes only; every entry point
 * returns 0 on success or a negative tpose_cuda_status, never exits the
 * process.  One context = one device + one stream; a context is not thread
 * safe, distinct contexts are independent.  Host buffers are owned by the
 * caller, device buffers by the context.
 *
 * The reference has no FFI: its "device interface" is the set of TinyEngine
 * Buffer / Shader / Compute calls made by four lambdas in each program.  Each
 * entry point below names the reference lines it replaces (paths relative to
 * the reference repository root).
 *
 * Buffer layouts are the reference's SSBO layouts (software/triangulate/
 * main.cpp:84-101): row k*NT+t of cn/ca/ten is variant k (0 = base, 1+4j+s =
 * vertex j moved by (+dp,0),(-dp,0),(0,+dp),(0,-dp) for s = 0..3) of triangle t.
 */
#ifndef TPOSE_CUDA_H
#define TPOSE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tpose_cuda_ctx tpose_cuda_ctx;

typedef enum {
    TPOSE_OK = 0,
    TPOSE_ERR_ARG = -1,       /* bad argument (sizes, slot, null pointer) */
    TPOSE_ERR_CUDA = -2,      /* a CUDA runtime call failed; see last_error */
    TPOSE_ERR_CAPACITY = -3,  /* NT / NP / image exceed what the context was created for */
    TPOSE_ERR_STATE = -4,     /* call order: no image / no mesh uploaded yet */
    TPOSE_ERR_NODEVICE = -5   /* no usable CUDA device: there is NO CPU fallback */
} tpose_cuda_status;

enum { TPOSE_TRIANGULATE = 0, TPOSE_WARP = 1 };          /* shader flavour */
enum { TPOSE_IMAGE_A = 0, TPOSE_IMAGE_B = 1 };           /* texture slot   */

/* trigger bits reported by tpose_cuda_run */
enum {
    TPOSE_TRIG_CONVERGED = 1u,  /* |relerr| < rel_tol (triangulation.hpp:653-674) */
    TPOSE_TRIG_MAXSTEPS  = 2u,  /* max_steps frames done, nothing fired */
    TPOSE_TRIG_TOPOLOGY  = 4u   /* a per-frame topology edit MAY apply (triangulate/main.cpp:316-346) */
};

/* The eight SSBO binding points of the reference (software/triangulate/main.cpp:84-101), for
 * tpose_cuda_read / tpose_cuda_write == Buffer::retrieve / Buffer::fill on that buffer. */
enum {
    TPOSE_BUF_POINTS = 0,   /* pointbuf    vec2   NP      */
    TPOSE_BUF_INDEX = 1,    /* trianglebuf ivec4  NT      */
    TPOSE_BUF_COLACC = 2,   /* tcolaccbuf  ivec4  13*NT   */
    TPOSE_BUF_COLNUM = 3,   /* tcolnumbuf  int    13*NT   */
    TPOSE_BUF_TENERGY = 4,  /* tenergybuf  int    13*NT   */
    TPOSE_BUF_PENERGY = 5,  /* penergybuf  int    (dead in the reference: lambda = 0; reads as 0) */
    TPOSE_BUF_GRADIENT = 6, /* pgradbuf    ivec2  NP      */
    TPOSE_BUF_NRING = 7     /* tnringbuf   int    (never read by the reference; reads as 0) */
};

/* Limits of the packed row-prefix table: */
#define TPOSE_MAX_WIDTH  4095
#define TPOSE_MAX_HEIGHT 8192
/* The reference allocates 2<<18 elements per SSBO and indexes 13*NT of them
 * (source/triangulation.hpp:97,592-601): NT <= 40329. */
#define TPOSE_REF_MAXT (2 << 18)

/* ---- lifetime ---------------------------------------------------------- */

/* Replaces tpose::init() (source/triangulation.hpp:592-608): allocates the
 * device buffers.  `stream` is a cudaStream_t to run on (NULL = the context
 * creates its own non-blocking stream). */
int tpose_cuda_create(int device, int max_triangles, int max_points, void *stream,
                      tpose_cuda_ctx **out);
/* Replaces tpose::quit() (source/triangulation.hpp:610-626). */
int tpose_cuda_destroy(tpose_cuda_ctx *ctx);
/* Message of the last failing call on this context ("" if none). NULL ctx:
 * message of the last failing tpose_cuda_create. */
const char *tpose_cuda_last_error(const tpose_cuda_ctx *ctx);
/* Library version / build string (arch, compiler). */
const char *tpose_cuda_version(void);

/* ---- inputs ------------------------------------------------------------ */

/* Replaces `Texture tex(IMG)` (software/triangulate/main.cpp:74,
 * software/warp/main.cpp:118-119).  rgba: h*w*4 bytes, row 0 = top row.
 * raster_w x raster_h is the window the reference would open
 * (Tiny::window(w/1.5, h/1.5), triangulate/main.cpp:53); pass w,h for the
 * native raster.  Copies the image to the device and builds the row-prefix
 * moment table the step kernels integrate over. */
int tpose_cuda_set_image(tpose_cuda_ctx *ctx, int slot, const uint8_t *rgba_host,
                         int w, int h, int raster_w, int raster_h);
/* Same, source already on the device (same layout). */
int tpose_cuda_set_image_device(tpose_cuda_ctx *ctx, int slot, const uint8_t *rgba_dev,
                                int w, int h, int raster_w, int raster_h);

/* Replaces tpose::upload(tr, uploadcolor) (source/triangulation.hpp:628-643):
 * pts_xy = NP vec2, tri_xyzw = NT ivec4 (w ignored).  colors_rgba (NT ivec4)
 * non-NULL == uploadcolor: replicated into all 13 variant rows of `ca`.
 * ratio = tpose::RATIO (the RATIO uniform of every shader). */
int tpose_cuda_upload(tpose_cuda_ctx *ctx, const float *pts_xy, int np,
                      const int32_t *tri_xyzw, int nt,
                      const int32_t *colors_rgba, float ratio);
/* pointbuf->fill(points) alone (positions changed, topology did not). */
int tpose_cuda_upload_points(tpose_cuda_ctx *ctx, const float *pts_xy, int np);

/* ---- the three lambdas, literally --------------------------------------- */

/* computecolors() (triangulate/main.cpp:121-130) / doreset() (warp/main.cpp:
 * 140-151): draw of 13*NT instances in mode 0.  Triangulate flavour: cn and
 * ca.rgb of every variant; warp flavour: cn only, ca keeps the upload. */
int tpose_cuda_colors(tpose_cuda_ctx *ctx, int slot, int flavour);
/* doenergy() (triangulate/main.cpp:132-141, warp/main.cpp:153-164): draw in
 * mode 1.  ten of every variant from the stored cn/ca; zeroes gr. */
int tpose_cuda_energy(tpose_cuda_ctx *ctx, int slot, int flavour);
/* doshift() (triangulate/main.cpp:143-155, warp/main.cpp:166-178):
 * gradient.cs then shift.cs. */
int tpose_cuda_shift(tpose_cuda_ctx *ctx, int flavour);

/* ---- fused path --------------------------------------------------------- */

/* `nsteps` frames at frozen topology with no host round trip:
 *   triangulate: { mode 0, mode 1, gradient, shift } at the current points
 *   warp:        { doreset, doenergy, doshift }
 * which is the per-frame device work of triangulate/main.cpp:196-197 +
 * :180-184 and warp/main.cpp:220-222.  After the call cn/ca/ten hold the
 * values of the LAST evaluated positions and the points are the shifted ones,
 * i.e. exactly what the reference reads back at main.cpp:201-204. */
int tpose_cuda_step(tpose_cuda_ctx *ctx, int slot, int flavour, int nsteps);

/* One whole frame of the reference's loop in one call and one synchronisation
 * (triangulate/main.cpp:196-204, warp/main.cpp:220-229): pointbuf->fill(pts_in_xy) if non-NULL, the
 * fused step, then the frame's read-backs into whichever of ten13 / cn13 / ca13 / pts_out_xy are
 * non-NULL (sizes as tpose_cuda_retrieve).  pts_in_xy may be reused when the call returns. */
int tpose_cuda_frame(tpose_cuda_ctx *ctx, int slot, int flavour, const float *pts_in_xy,
                     int32_t *ten13, int32_t *cn13, int32_t *ca13_rgba, float *pts_out_xy);

/* Extension (default 1 = the reference): multiplies the step constant of shift.cs:45 (5e-5
 * triangulate, 3e-5 warp).  The integer gradient is a sum over pixels, so the reference's constant
 * -- tuned to its 800x450 window -- takes steps that grow with the pixel count: at 1920x1080 a
 * two-triangle mesh never settles and a loaded 1000-triangle mesh tears itself apart (DESIGN.md
 * section 6).  (450 / raster_h)^2 restores the reference's dynamics on a larger raster. */
int tpose_cuda_set_step_scale(tpose_cuda_ctx *ctx, float scale);

/* Up to `max_steps` frames, stopping after the first one whose geterr value
 * (float32 running sum of the base energies, relative change against the
 * previous sum -- source/triangulation.hpp:653-674) is < rel_tol.  The running
 * `toterr` lives in the context (it is a global in the reference, starts at 1
 * and persists across levels); tpose_cuda_set_toterr overrides it.
 * steps_done / trigger_flags / last_err (|relerr| of the last frame) may be NULL. */
int tpose_cuda_run(tpose_cuda_ctx *ctx, int slot, int flavour, int max_steps, double rel_tol,
                   int *steps_done, uint32_t *trigger_flags, float *last_err);
int tpose_cuda_set_toterr(tpose_cuda_ctx *ctx, float toterr);
int tpose_cuda_get_toterr(tpose_cuda_ctx *ctx, float *toterr);

/* The topology the HOST currently holds (tr.triangles / tr.halfedges).  The reference edits its
 * host mesh every frame (prune, flips of angles > 0.8*pi, collapse of edges <= 0.01 --
 * triangulate/main.cpp:316-346) and uploads it only when prune/collapse/split changed something,
 * so the host mesh can differ from the uploaded one by pending flips.  With a watch topology set,
 * tpose_cuda_run also stops (TPOSE_TRIG_TOPOLOGY) after the first frame at whose end one of those
 * edits could apply -- a conservative superset, decided on the device from the shifted points --
 * so the caller runs the reference's per-frame host code only on frames where it can do
 * something.  tri_xyzw == NULL switches the watch off.  halfedges: 3*nt twins or -1. */
int tpose_cuda_watch(tpose_cuda_ctx *ctx, const int32_t *tri_xyzw, const int32_t *halfedges, int nt);

/* Per-frame history of tpose_cuda_run (frames > 0 switches it on for runs of up to `frames` steps,
 * 0 off).  The reference looks at every frame on the host: it reads the points back and runs its
 * three edit loops (triangulate/main.cpp:199-346).  With the history on, a run records for each of
 * its frames the points the frame left, the base energies it evaluated, toterr after its geterr
 * and the watch mask at its points, and returns them with its verdict in one synchronisation, so
 * the caller can replay its per-frame host code over a whole run afterwards.  Because the 0.8*pi
 * flips never change what is uploaded, such a run no longer stops for them: TPOSE_TRIG_TOPOLOGY
 * then means a prune or collapse may apply (conservatively, against the watch topology as given --
 * the caller, whose own flips re-shape triangles during the replay, must check those itself). */
int tpose_cuda_set_history(tpose_cuda_ctx *ctx, int frames);
/* Frame `frame` (0 .. steps_done-1) of the last tpose_cuda_run; any output may be NULL.
 * pts_xy: NP vec2; mask_words: as tpose_cuda_watch_mask; ten_base: NT int32; toterr: 1 float. */
int tpose_cuda_history_frame(tpose_cuda_ctx *ctx, int frame, float *pts_xy, uint32_t *mask_words, int nwords,
                             int32_t *ten_base, float *toterr);
/* Puts pgradbuf back to what frame `frame` of the last run left (device-side copy).  A caller that
 * resumes from a recorded frame needs it: triangle.vs:93-98 zeroes the gradient only of vertices
 * some triangle references, so a vertex the caller's edit leaves without triangles keeps -- and
 * shift.cs keeps applying -- the entry of exactly that frame. */
int tpose_cuda_history_restore_gradient(tpose_cuda_ctx *ctx, int frame);

/* After a tpose_cuda_run that stopped with TPOSE_TRIG_TOPOLOGY: one bit per triangle of the watch
 * topology (bit t&31 of word t>>5), set where the device's conservative test fired on the frame
 * the run stopped on -- a superset of the triangles on which prune / the 0.8*pi flip / collapse can
 * do anything at the points that frame left.  Every other triangle can be skipped by the host's
 * per-frame loops without changing their result.  nwords >= (NT+31)/32. */
int tpose_cuda_watch_mask(tpose_cuda_ctx *ctx, uint32_t *mask_words, int nwords);

/* A few instances evaluated on their own, at the context's current points: entry i = (v0, v1, v2, k) -- three
 * vertex indices and a variant number k (0 = the triangle as it is, 1 + 4j + s = corner j moved by +-dp like
 * triangle.vs:66-78; dp from the uploaded NT) -- yields what the full triangulate-flavour evaluation would put
 * in row k of a triangle with those vertices: cn (fragment count), ca (colour sums), ten (energy against the
 * instance's own mean colour, triangle.fs:27-43).  Of the context's SSBOs only pgradbuf is touched, and only
 * with reset_gradient != 0: the mode-1 draw this call stands in for zeroes it for every vertex the uploaded
 * mesh references (triangle.vs:93-98), which is visible later if an edit leaves such a vertex unreferenced.  This is the
 * device half of the flip set of a converged frame (software/triangulate/main.cpp:287-306 re-rasterises all
 * 13*NT instances twice to read a few base energies).  Needs the table path.  Outputs may be NULL. */
int tpose_cuda_eval_list(tpose_cuda_ctx *ctx, int slot, const int32_t *entries_v0v1v2k, int n, int reset_gradient,
                         int32_t *ten_out, int32_t *cn_out, int32_t *ca_out_rgba);

/* ---- batches of independent jobs --------------------------------------------
 * n contexts on the same device advance in lock step, one launch per frame for all of them
 * (blockIdx.y = job).  Jobs are independent problems (own image, mesh, buffers); results are
 * identical to running each context alone.  Work is issued on ctxs[0]'s stream. */
int tpose_cuda_batch_step(tpose_cuda_ctx *const *ctxs, int n, const int *slots, int flavour, int nsteps);
int tpose_cuda_batch_run(tpose_cuda_ctx *const *ctxs, int n, const int *slots, int flavour, int max_steps,
                         double rel_tol, int *steps_done, uint32_t *trigger_flags);

/* ---- read back ---------------------------------------------------------- */

/* Buffer::retrieve of tenergybuf / tcolnumbuf / tcolaccbuf / pointbuf
 * (triangulate/main.cpp:201-204,221).  Any pointer may be NULL.
 * ten13, cn13: 13*NT int32; ca13_rgba: 13*NT ivec4; pts_xy: NP vec2. */
int tpose_cuda_retrieve(tpose_cuda_ctx *ctx, int32_t *ten13, int32_t *cn13,
                        int32_t *ca13_rgba, float *pts_xy);
/* Only the base rows (what geterr / maxerrid / the flip set read): NT each. */
int tpose_cuda_retrieve_base(tpose_cuda_ctx *ctx, int32_t *ten, int32_t *cn, int32_t *ca_rgba);
/* pgradbuf: NP ivec2 (debug / parity). */
int tpose_cuda_retrieve_gradient(tpose_cuda_ctx *ctx, int32_t *gr_xy);

/* Buffer::retrieve(count, out) / Buffer::fill(count, in) on one binding point, elements
 * [first, first+count) in that buffer's element type. */
int tpose_cuda_read(tpose_cuda_ctx *ctx, int buffer, size_t first, size_t count, void *out);
int tpose_cuda_write(tpose_cuda_ctx *ctx, int buffer, size_t first, size_t count, const void *in);

/* Per-pixel triangle-ID raster of the NT base instances at the current
 * points, raster_h*raster_w int32, -1 where nothing is drawn; where instances
 * overlap the highest index (last drawn) wins.  This is what mode 2 of
 * triangle.fs displays (triangulate/main.cpp:157-166). */
int tpose_cuda_raster_ids(tpose_cuda_ctx *ctx, int slot, int32_t *ids_host);

/* software/view without a window (view/main.cpp:104-121, view/shader/triangle.{vs,fs}): the NT
 * triangles drawn flat in the colours of tcolaccbuf's base rows (what tpose_cuda_upload put there),
 * on black, at corners mix(points, other_pts, s) -- other_pts_xy (NP vec2, e.g. originpoints) may be
 * NULL for s = 0.  Same raster conventions as the step; rgba_host: raster_h*raster_w*4 bytes. */
int tpose_cuda_render(tpose_cuda_ctx *ctx, int raster_w, int raster_h, const float *other_pts_xy, float s,
                      uint8_t *rgba_host);
/* The same picture with the viewer's wireframe on top (its `n` key; view/main.cpp:58-69,112-121 and
 * view/shader/linestrip.{vs,fs}): the three sides of every triangle as one-pixel white lines.  Line fragments
 * follow the written-down convention of csrc/raster.cuh (GL leaves them to the driver). */
int tpose_cuda_render_wire(tpose_cuda_ctx *ctx, int raster_w, int raster_h, const float *other_pts_xy, float s,
                           uint8_t *rgba_host);

/* Device time in milliseconds of the most recent step/run call on this
 * context, measured with CUDA events on the context's stream. */
int tpose_cuda_last_elapsed_ms(tpose_cuda_ctx *ctx, float *ms);
/* Per-kernel timing (off by default: it puts an event between the two kernels of a step).  When on, tpose_cuda_last_kernel_ms gives the device time of the two
 * kernels of the most recent tpose_cuda_step(.., nsteps = 1): the edge integrals (k_edges) and the
 * per-variant evaluation + vertex update (k_corners).  TPOSE_ERR_STATE after any other call. */
int tpose_cuda_set_kernel_timing(tpose_cuda_ctx *ctx, int on);
int tpose_cuda_last_kernel_ms(tpose_cuda_ctx *ctx, float *edges_ms, float *corners_ms);
/* Which kernels integrate the edges: 0 = gathers from a row-prefix table (k_edges, the default), 1 = TMA-staged
 * image tiles (k_setup + k_tiles; environment TPOSE_CUDA_PATH=tiles when the context is created).  On the
 * tile path edges_ms above is k_tiles alone and tpose_cuda_last_setup_ms is the k_setup in front of it
 * (snapped end points, tile ranges, zeroed records); both produce identical records. */
int tpose_cuda_kernel_path(const tpose_cuda_ctx *ctx);
int tpose_cuda_last_setup_ms(tpose_cuda_ctx *ctx, float *setup_ms);
/* Block until the context's stream is idle. */
int tpose_cuda_sync(tpose_cuda_ctx *ctx);
/* Number of kernel launches issued by this context so far. */
int64_t tpose_cuda_launch_count(const tpose_cuda_ctx *ctx);

/* ---- multiview (source/multiview.hpp), batched; host pointers in and out -----------------
 * These three need no context: pass the CUDA device ordinal.  Errors: tpose_cuda_mview_last_error. */

/* mview::F_8Point (multiview.hpp:130-183) for nbatch independent match lists.  A_xy / B_xy: the
 * lists concatenated (vec2 each); offsets[nbatch+1], offsets[0] = 0; F9_out: nbatch x 9 floats,
 * row-major, scaled so that F(2,2) = 1.  An empty list yields the identity, like the reference. */
int tpose_cuda_f8point(int device, const float *A_xy, const float *B_xy, const int *offsets, int nbatch,
                       float *F9_out);
/* mview::F_Sampson (multiview.hpp:187-242; :244-299 when `weights` -- one float per match, same
 * concatenation -- is non-NULL): the 8-point estimate, then `iterations` (the reference: 100)
 * solves re-weighted by the Sampson denominators of the current estimate.  Same outputs as above. */
int tpose_cuda_fsampson(int device, const float *A_xy, const float *B_xy, const float *weights,
                        const int *offsets, int nbatch, int iterations, float *F9_out);
/* Match selection and weighting by two-way warp inconsistency -- the step the reference keeps as commented-out
 * code in tests/compute_fundamental_mat/model.h:1-82.  For every vertex i of view A, then of view B:
 * d_i = |warp_origin[i] - points[i]| (the other direction's .warp record against this one); the vertex becomes a
 * match with weight 1/d_i unless d_i > max_dist (the reference: 0.0000002), an end point lies on the domain
 * boundary or outside the central half of the domain.  A's matches run originpoints -> points, B's
 * points -> originpoints, both mapped (x, y) -> ((0.5f/RATIO) x, (-0.5f/RATIO) y).  X_xy / Y_xy / weights: room
 * for npa + npb entries; *n_out: matches kept (index order, A's first); dist (may be NULL): all npa + npb d_i.
 * The outputs feed tpose_cuda_fsampson(.., weights, ..). */
int tpose_cuda_select_matches(int device, const float *originA_xy, const float *pointsA_xy, const float *warp_originA_xy, int npa,
                              const float *pointsB_xy, const float *originB_xy, const float *warp_originB_xy, int npb,
                              float ratio, double max_dist, float *X_xy, float *Y_xy, float *weights, float *dist, int *n_out);
/* mview::triangulate(F, K, A, B) (multiview.hpp:529-627): per match the Hartley-Sturm optimal
 * correction (:415-527; skipped if optimal_correction == 0), then HDLT (:370-380) against
 * PA = K[I|0], PB = K[R2|-t] from E = K^T F K (check = 3).  F9 / K9: one 3x3 per list.
 * X4_out: vec4 per match, already divided by w.  A_corr / B_corr (may be NULL): corrected matches. */
int tpose_cuda_triangulate(int device, const float *F9, const float *K9, const float *A_xy, const float *B_xy,
                           const int *offsets, int nbatch, int optimal_correction, float *X4_out,
                           float *A_corr, float *B_corr);
/* mview::HDLT alone (multiview.hpp:370-380) for n matches with caller-supplied 3x4 projections
 * (row-major float[12]).  X4_out is NOT divided by w; sign: largest component positive. */
int tpose_cuda_hdlt(int device, const float *PA12, const float *PB12, const float *A_xy, const float *B_xy,
                    int n, float *X4_out);
/* geterr (source/triangulation.hpp:653-674) over `n` int32 energies on the device: newerr = their float32 sum in index
 * order, relerr = |(*toterr_inout - newerr) / *toterr_inout|, *toterr_inout = newerr, *below_tol_out = relerr < rel_tol
 * (float against a double literal, as `geterr(&tr) < 1E-4` compares).  This is the routine every frame of
 * tpose_cuda_run ends with (csrc/eval.cuh), exposed on its own.  route: -1 as a run would choose for this n,
 * 0 the literal chain of additions by one warp, 1 the block-parallel routine (geterr_block) -- same bits.
 * relerr_out / below_tol_out may be NULL.  Errors: tpose_cuda_mview_last_error(). */
int tpose_cuda_geterr(int device, const int32_t *energies, int n, float *toterr_inout, float *relerr_out,
                      int *below_tol_out, double rel_tol, int route);

/* triangulation::warp / reversewarp (source/triangulation.hpp:465-520): every point of pts_xy that
 * is not on the domain boundary (|x| >= ratio or |y| >= 1) is located in the first triangle, in
 * index order, of the embedding from_xy that strictly contains it, and replaced by the point with
 * the same barycentric coordinates in the embedding to_xy (both nv vec2; tri_xyzw indexes them).
 * Same arithmetic as the host path (GLM's 3x3 inverse, float32, no FMA): identical bits.
 * No context needed; errors through tpose_cuda_mview_last_error. */
int tpose_cuda_carry(int device, const int32_t *tri_xyzw, int nt, const float *from_xy, const float *to_xy, int nv,
                     float *pts_xy, int npts, float ratio);
const char *tpose_cuda_mview_last_error(void);
/* device milliseconds of the kernel(s) of the most recent f8point / triangulate call */
float tpose_cuda_mview_last_ms(void);

/* Page-locked host memory for the buffers handed to set_image / upload / retrieve
 * (optional: pageable pointers work, pinned ones copy faster). */
int tpose_cuda_host_alloc(size_t bytes, void **out);
int tpose_cuda_host_free(void *p);
/* Writes `bytes` of device scratch owned by the context (benchmark hygiene: evicts L2). */
int tpose_cuda_flush_l2(tpose_cuda_ctx *ctx, size_t bytes);

/* ---- software/rgbd: plane-fit meshing core (rgbd/main.cpp:244-349) ----------------------------------------
 * For every triangle of a triangulation (mesh coordinates pts_xy): the depth points whose colour-image pixel
 * (pixels_xy) lies inside the triangle's pixel-space bounding box and strictly inside the triangle (tpose::
 * intriangle after TB) give a plane -- mean of points_xyzw.xyz, normalised mean of normals_xyzw.xyz, summed in
 * index order like the reference's std::set loop; the triangle's three corners (T * p) are un-projected onto
 * it and every vertex becomes the average of the corners its triangles produced (verts_xyzw, np vec4; a vertex
 * with no such triangle is 0/0 = NaN, as coded).  T9 / TB9: the reference's mat3 T and inverse(T)
 * (rgbd/main.cpp:94-99), column-major.  The library piolib the reference leans on is unreleased; its helpers
 * are restated as a pinhole model: project(p) = (fx p.x/p.z + ppx, fy p.y/p.z + ppy), unproject = the ray
 * through the pixel intersected with the plane.  ring (np floats) / tri_counts (nt ints, points gathered per
 * triangle) may be NULL.  The reference's brute-force host loop is O(nt * npoints); here one block per
 * triangle scans the points. */
int tpose_cuda_rgbd_planefit(int device, const int32_t *tri_xyzw, int nt, const float *pts_xy, int np,
                             const float *T9, const float *TB9, const float *pixels_xy, const float *points_xyzw,
                             const float *normals_xyzw, int npoints, const float *intr_fx_fy_ppx_ppy,
                             float *verts_xyzw, float *ring, int32_t *tri_counts);

#ifdef __cplusplus
}
#endif
#endif /* TPOSE_CUDA_H */
