/* painty_b200 — C ABI of the B200-native paint-renderer hot path (imprint + Kubelka-Munk compose).
 *
 * painty has no plugin / FFI layer: its boundary is the C++ template API in namespace painty
 * (SURVEY.md §8b). This header is the flat surface the C++ façade in include/painty/ forwards to,
 * and that any other host (ctypes, cgo, JNI) can bind. Plain pointers and sizes only.
 *
 * Conventions
 *   - Every call returns 0 on success, non-zero on failure; pb_last_error() (thread local) explains.
 *     The C++ façade rethrows the exception type the reference would have thrown.
 *   - "host AoS f64" = the reference's own boundary layout: Mat<vec3> = rows*cols*3 doubles,
 *     Mat<double> = rows*cols doubles, row-major (painty/image/Mat.hxx:40-41, core/Vec.hxx:35).
 *   - Device storage is SoA planes in HBM: Kr,Kg,Kb,Sr,Sg,Sb,V (+R0r,R0g,R0b,h for a canvas), dense
 *     row-major, element type float (PB_F32) or double (PB_F64 validation mode) per context.
 *   - One context per device; a handle is not thread safe; all work is ordered on the context's stream
 *     (the reference is single-submitter too: GpuTaskQueue = ThreadPool{1}, gpu/GpuTaskQueue.hxx:33-46).
 *   - Canvas / PaintLayer constructors take (rows, cols) like the reference.
 *   - There is no CPU fallback: without a CUDA device every entry point fails with an error.
 */
#ifndef PAINTY_B200_H
#define PAINTY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_context pb_context;
typedef struct pb_canvas pb_canvas; /* painty::Canvas<vec3>           renderer/Canvas.hxx:20-198 */
typedef struct pb_layer pb_layer;   /* painty::PaintLayer<vec3>       renderer/PaintLayer.hxx:22-145 */
typedef struct pb_fbrush pb_fbrush; /* painty::FootprintBrush<vec3>   renderer/FootprintBrush.hxx:22-503 */
typedef struct pb_tbrush pb_tbrush; /* painty::TextureBrush<vec3>     renderer/TextureBrush.hxx:19-237 */

enum { PB_F32 = 0, PB_F64 = 1 };

const char* pb_last_error(void);
int pb_version(void);

/* ---- context ------------------------------------------------------------------------------- */
int pb_context_create(int device, int precision, pb_context** out);
int pb_context_destroy(pb_context* ctx);
int pb_context_synchronize(pb_context* ctx);
int pb_context_precision(const pb_context* ctx);
void* pb_context_stream(pb_context* ctx); /* cudaStream_t all work of this context is ordered on */
/* Number of CUDA kernels this context has launched so far (bench.py's gpu_launches). */
int64_t pb_context_launch_count(const pb_context* ctx);

/* ---- host-side scalar calls that stay on the host (bit-exact f64) ---------------------------- */
/* core/KubelkaMunk.hxx:28-83 ComputeReflectance<double,3>. */
int pb_compute_reflectance(const double K[3], const double S[3], const double R0[3], double d, double out[3]);
/* core/KubelkaMunk.hxx:92-124; fails (invalid_argument) unless 0 < Rb < Rw < 1 per channel. */
int pb_compute_scattering_absorption(const double Rb[3], const double Rw[3], double K[3], double S[3]);
/* mixer/src/PaintMixer.cxx:536-548 PaintMixer::mixed. */
int pb_paint_mixed(const double K1[3], const double S1[3], double v1, const double K2[3], const double S2[3], double v2,
                   double K[3], double S[3]);
/* mixer/src/PaintMixer.cxx:327-355 PaintMixer::mixSinglePaint; base = n*3 doubles each. */
int pb_paint_mix_single(int n, const double* baseK, const double* baseS, int n_weights, const double* weights, double K[3],
                        double S[3]);
/* Stroke -> imprint expansion (host, f64).
 * mode 0: library form FootprintBrush::paintStroke (FootprintBrush.hxx:251-267) with p_pre = path[0]
 *         for the first segment (the reference indexes path[-1] there: UB);
 * mode 1: GUI form DigitalCanvas::mouseMoveEvent (apps/painty_gui/DigitalCanvas.cxx:107-123), i.e. the
 *         imprints produced while the points arrive one by one, control = (p[n-3], p[n-2], p[n-1], p[n-1]).
 * path = n*2 doubles (x,y). Writes up to `capacity` imprints into cx/cy/theta and returns the full
 * count in *n_imprints (call with capacity 0 to size). theta = atan2(dir.y, dir.x). */
int pb_expand_stroke(int mode, int n, const double* path_xy, int64_t capacity, double* cx, double* cy, double* theta,
                     int64_t* n_imprints);

/* The same for a list of strokes (stroke s = vertices [first_vertex[s], first_vertex[s] + n_vertices[s]) of path_xy): imprints are
 * written back to back, first_imprint[s] / n_imprints[s] (either may be NULL) receive every stroke's range and *total the
 * overall count. Call with capacity 0 to size. */
int pb_expand_stroke_batch(int mode, int64_t n_strokes, const int64_t* first_vertex, const int32_t* n_vertices, const double* path_xy,
                           int64_t capacity, double* cx, double* cy, double* theta, int64_t* first_imprint, int64_t* n_imprints,
                           int64_t* total);

/* Host-side dataflow planner used by the stroke batches (no device needed; exposed for testing and for hosts that
 * want to inspect the schedule). Strokes are given in submission order by two inclusive pixel rectangles
 * (x0,y0,x1,y1): `box` = what the stroke modifies, `allowed` = what it may read or refresh (for a texture stroke pass
 * the same rectangle twice). Two strokes must keep their order iff the box of one meets the allowed rectangle of the
 * other (at 64 px tile granularity). Writes the predecessor lists as CSR: offsets[n+1] and up to `capacity` entries of
 * preds; *n_preds receives the total count. Waiting for the listed predecessors is sufficient (ordering is
 * transitive through them). */
int pb_plan_dependencies(int rows, int cols, int64_t n, const int32_t* box, const int32_t* allowed, int64_t* offsets,
                         int64_t capacity, int32_t* preds, int64_t* n_preds);

/* The footprint brush's dataflow graph at SEGMENT granularity, as pb_fbrush_stroke_batch builds it (host only;
 * exposed for testing and schedule inspection). Stroke s owns the imprints [first[s], first[s]+count[s]) of cx/cy,
 * has a footprint of side[s] cells and the given radius. A stroke is cut into segments of seg_len[s] consecutive
 * imprints (segment_length, enlarged for very long strokes; 0 = whole strokes). Outputs: seg_first[n+1] (first
 * global segment of each stroke), seg_len[n], and a CSR list per global segment — seg_off[segments+1] (up to
 * seg_capacity+1 entries written) into pred_stroke/pred_need (up to pred_capacity entries, total in *n_preds):
 * the segment may start once stroke pred_stroke[i] (always an earlier stroke) has completed pred_need[i] segments.
 * single (may be NULL): non-zero keeps that stroke in one segment (what the multi-GPU path does for strokes that stage
 * neighbour rows). Call once with zero capacities to size the outputs (seg_first[n] and *n_preds). */
int pb_plan_segments(int rows, int cols, int64_t n, const int64_t* first, const int64_t* count, const int32_t* side,
                     const double* radius, const unsigned char* single, const double* cx, const double* cy, int segment_length,
                     int use_snapshot,
                     int32_t* seg_first, int32_t* seg_len, int64_t seg_capacity, int32_t* seg_off, int64_t pred_capacity,
                     int32_t* pred_stroke, int32_t* pred_need, int64_t* n_preds);

/* Claim order of the device stroke queues (host only; exposed for testing). The strokes are list-scheduled with a
 * cost model: stroke s runs on slot pool pool[s] (its GPU) in launch run[s] of that pool and costs cost[s] per
 * imprint; pool p has runs_per_pool[p] launches with slots[...] concurrent strokes each (flattened pool-major).
 * order[n] receives the global claim sequence: a topological order of the segment-level dependency graph (every
 * stroke after all strokes any of its segments waits for) that keeps the runs of a pool in sequence. *makespan (may be
 * NULL) receives the model's completion time of the batch in units of cost. */
int pb_plan_claim_order(int rows, int cols, int64_t n, const int64_t* first, const int64_t* count, const int32_t* side,
                        const double* radius, const unsigned char* single, const double* cx, const double* cy,
                        int segment_length, int use_snapshot, const int32_t* pool, const int32_t* run, const double* cost,
                        int n_pools, const int32_t* runs_per_pool, const int32_t* slots, int32_t* order, double* makespan);

/* Test hook (host only): the dirty-map words (4 pixels each) the imprint kernel reads for one snapshot ring pass
 * (FootprintBrush.hxx:278-319), enumerated by the same code the device runs. box = (tlx, tly, brx, bry), the
 * footprint box whose open interior is excluded; allowed = (ax0, ay0, ax1, ay1), the clipped allowed box, inclusive.
 * prev_box / prev_allowed (both NULL or both set) = geometry of the previous imprint: the pass then only enumerates what
 * entered the ring since. The dirty map is flat (byte index = row * pitch + column): writes up to `capacity`
 * (row, flat word index) pairs and the total count. */
int pb_ring_rects(const int32_t box[4], const int32_t allowed[4], const int32_t* prev_box, const int32_t* prev_allowed, int pitch,
                  int64_t capacity, int32_t* rows, int32_t* words, int64_t* n_words);

/* Test hook (host only): the canvas pixels whose rotated + rounded position (FootprintBrush.hxx:88-114) is pickup-map
 * cell (mx[i], my[i]) for an imprint at (cx, cy, theta) of a footprint with half side `half_side`, evaluated by the
 * code the device runs. mode 0: the exact f64 test; mode 1: the single-precision test with undecided band `eps`
 * (n_hits[i] = -1 where it defers to the exact test). phase = -1, or the border phase 0..3. px/py hold 2 entries
 * per cell in the reference's row-major order. */
int pb_imprint_hits(double cx, double cy, double theta, int half_side, int rows, int cols, int64_t n_cells, const int32_t* mx,
                    const int32_t* my, int mode, double eps, int phase, int32_t* n_hits, int32_t* px, int32_t* py);

/* ---- PaintLayer ------------------------------------------------------------------------------ */
int pb_layer_create(pb_context* ctx, int rows, int cols, pb_layer** out);
int pb_layer_destroy(pb_layer* l);
int pb_layer_rows(const pb_layer* l);
int pb_layer_cols(const pb_layer* l);
int pb_layer_clear(pb_layer* l);                                                    /* PaintLayer.hxx:68-74 */
int pb_layer_upload(pb_layer* l, const double* K, const double* S, const double* V); /* host AoS f64 -> SoA */
int pb_layer_download(pb_layer* l, double* K, double* S, double* V);                 /* any may be NULL */
int pb_layer_copy(const pb_layer* src, pb_layer* dst);                              /* PaintLayer.hxx:103-111 */
/* PaintLayer::composeOnto (PaintLayer.hxx:81-96), R0 host AoS f64 in/out. */
int pb_layer_compose_onto(pb_layer* l, double* R0);
/* Renderer::compose(layer, R0) (Renderer.hxx:26-41): out = KM(layer over R0), host AoS f64. */
int pb_layer_compose(pb_layer* l, const double* R0, double* out);

/* ---- Canvas ---------------------------------------------------------------------------------- */
int pb_canvas_create(pb_context* ctx, int rows, int cols, pb_canvas** out);
/* Row band [row_begin,row_end) (plus `halo` rows on both sides where they exist) of a rows x cols canvas;
 * coordinates passed to brushes stay global. Used for multi-GPU sharding. */
int pb_canvas_create_band(pb_context* ctx, int rows, int cols, int row_begin, int row_end, int halo, pb_canvas** out);
int pb_canvas_destroy(pb_canvas* c);
int pb_canvas_rows(const pb_canvas* c);
int pb_canvas_cols(const pb_canvas* c);
int pb_canvas_clear(pb_canvas* c);                            /* Canvas.hxx:37-58 */
int pb_canvas_set_background(pb_canvas* c, const double* R0); /* Canvas.hxx:80-87 (clear + copy) */
int pb_canvas_dry(pb_canvas* c);                              /* Canvas.hxx:105-121, fused compose+zero+h */
int pb_canvas_upload_layer(pb_canvas* c, const double* K, const double* S, const double* V);
/* Download stored rows (all rows of a full canvas; band+halo rows of a band canvas). Any may be NULL. */
int pb_canvas_download(pb_canvas* c, double* K, double* S, double* V, double* R0, double* h);
/* Renderer::compose(canvas) (Renderer.hxx:48-53) into host AoS f64 (stored rows). */
int pb_canvas_compose(pb_canvas* c, double* out);
/* Same, result stays on the device: 3 planes (r,g,b) of stored_rows*cols elements of the context's
 * element type, written to caller-provided device memory (plane p at d_out + p*plane_stride elements). */
int pb_canvas_compose_device(pb_canvas* c, void* d_out, int64_t plane_stride);
/* Compose only the owned band rows [row_begin,row_end) into d_out (3 planes, band_rows*cols each). */
int pb_canvas_compose_band_device(pb_canvas* c, void* d_out, int64_t plane_stride);
/* Display / IO epilogue fused into the compose kernel: compose -> ColorConverter::rgb2srgb
 * (core/Color.hxx:198-206) -> quantise, so only 4 (or 3 / 6) bytes per pixel cross PCIe instead of 24.
 *   pb_canvas_compose_qrgb32: the GUI path (apps/painty_gui/DigitalCanvas.cxx:164-177): qRgb(uint8(r*255), ...) =
 *     0xffRRGGBB per pixel, truncating cast; out = rows*cols uint32 (host).
 *   pb_canvas_compose_bgr: the io::imSave path up to the encoder (io/src/ImageIO.cxx:158-191): convertTo with
 *     scale 0xff (bits = 8) or 0xffff (bits = 16) = round-half-even + saturate, channel order BGR interleaved;
 *     out = rows*cols*3 uint8 / uint16 (host). srgb = 0 skips the sRGB conversion (convertTo_sRGB = false). */
int pb_canvas_compose_qrgb32(pb_canvas* c, uint32_t* out);
int pb_canvas_compose_bgr(pb_canvas* c, int bits, int srgb, void* out);
/* The sbr planner's canvas read-back prep on the device (painty/sbr/src/PictureTargetSbrPainter.cxx:334-341):
 * ScaledMat(convertColor(getLinearRgbImage(), rgb_2_CIELab), out_rows, out_cols) = compose -> ColorConverter::rgb2lab
 * (core/Color.hxx:248-252, D65) -> cv::resize(INTER_LANCZOS4) (image/Mat.hxx:141-147) evaluated like OpenCV does for a
 * CV_64FC3 image (separable 8-tap passes, float weights, double sums). out = out_rows*out_cols*3 doubles (host AoS).
 * Only the down-scaled Lab image crosses PCIe. Full canvases only. */
int pb_canvas_compose_lab_scaled(pb_canvas* c, int out_rows, int out_cols, double* out);
/* Test hook (host only): the tap offsets (dst entries: first source index + 3) and weights (dst*8 floats) of one axis of
 * that resize. */
int pb_lanczos4_taps(int src, int dst, int32_t* offsets, float* weights);
/* Renderer::render(canvas) (renderer/Renderer.hxx:60-156): compose + directional-light relighting (Beckmann /
 * Cook-Torrance, the wet layer's thickness as height field, 5-tap BORDER_REFLECT normal) fused in one kernel; host
 * AoS f64 out, clamped to [0,1] like the reference. Full canvases only (the stencil needs the neighbouring rows). */
int pb_canvas_render(pb_canvas* c, double* out);
/* Device plane pointers (element type of the context): 0-2 K, 3-5 S, 6 V, 7-9 R0, 10 h. The pointers are writable, so
 * handing them out counts as a modification of the wet layer (a footprint brush then treats its snapshot bookkeeping as
 * out of date, exactly as after clear / dry / upload). A host that keeps the pointers and writes through them LATER must
 * call pb_canvas_mark_modified before the next brush operation. */
int pb_canvas_device_planes(pb_canvas* c, void* planes[11], int64_t* elems_per_plane);
int pb_canvas_mark_modified(pb_canvas* c);
int pb_canvas_stored_rows(const pb_canvas* c, int* first_row, int* n_rows);
/* Canvas::getPaintLayer(): a non-owning pb_layer aliasing the canvas' wet planes (destroy with pb_layer_destroy;
 * it must not outlive the canvas). */
int pb_canvas_paint_layer(pb_canvas* c, pb_layer** out);
/* Write R0 (getR0()) and/or h (get_h()) from host AoS f64; either may be NULL. Does not clear the wet layer. */
int pb_canvas_upload_substrate(pb_canvas* c, const double* R0, const double* h);

/* ---- raw streaming compose on caller-owned device SoA planes (roofline bench, stacked layers) --- */
/* R = KM(K,S,V over R0) for n pixels; pointers are device pointers of the context's element type.
 * R may alias R0 (in-place composeOnto). */
int pb_km_compose_planes(pb_context* ctx, int64_t n, const void* const K[3], const void* const S[3], const void* V,
                         const void* const R0[3], void* const R[3]);
/* L stacked layers bottom-up in one pass, R kept in registers: layer l planes at index l. */
int pb_km_compose_stacked_planes(pb_context* ctx, int64_t n, int n_layers, const void* const* K /*3*L*/,
                                 const void* const* S /*3*L*/, const void* const* V /*L*/, const void* const R0[3],
                                 void* const R[3]);

/* ---- FootprintBrush -------------------------------------------------------------------------- */
int pb_fbrush_create(pb_context* ctx, pb_fbrush** out);
int pb_fbrush_destroy(pb_fbrush* b);
/* FootprintBrush::setRadius (FootprintBrush.hxx:46-63). The host (painty's own io::imRead + ScaledMat +
 * PaddedMat) supplies the padded footprint: side*side doubles. Acts only when |radius - current| >= 0.5
 * (then the pickup map is reallocated and zeroed); *acted says which. Pass footprint == NULL to ask
 * whether a call would act (*acted) without changing anything. */
int pb_fbrush_set_radius(pb_fbrush* b, double radius, int side, const double* footprint, int* acted);
int pb_fbrush_dip(pb_fbrush* b, const double K[3], const double S[3]); /* :150-154 clean + set paint */
int pb_fbrush_clean(pb_fbrush* b);                                     /* :160-166 */
int pb_fbrush_set_pickup_rate(pb_fbrush* b, double rate);
int pb_fbrush_set_deposition_rate(pb_fbrush* b, double rate);
double pb_fbrush_get_pickup_rate(const pb_fbrush* b);
double pb_fbrush_get_deposition_rate(const pb_fbrush* b);
int pb_fbrush_set_use_snapshot(pb_fbrush* b, int use);
int pb_fbrush_get_use_snapshot(const pb_fbrush* b);
int pb_fbrush_size_map(const pb_fbrush* b);
/* getPickupMap (:174) as host AoS f64 (size_map^2 pixels). */
int pb_fbrush_pickup_map(pb_fbrush* b, double* K, double* S, double* V);
/* Same as a non-owning pb_layer view (valid until the next setRadius that acts). */
int pb_fbrush_pickup_layer(pb_fbrush* b, pb_layer** out);
/* public updateSnapshot(canvas) (:168-172): full copy canvas wet layer -> snapshot. */
int pb_fbrush_update_snapshot(pb_fbrush* b, pb_canvas* c);
int pb_fbrush_snapshot_download(pb_fbrush* b, double* K, double* S, double* V);
/* FootprintBrush::imprint (:73-143): n imprints applied in order, one dependent chain on the device. */
int pb_fbrush_imprint_batch(pb_fbrush* b, pb_canvas* c, int64_t n, const double* cx, const double* cy,
                            const double* theta);
/* A batch of strokes in submission order, each = dip(K,S) -> setRadius(radius) -> its imprints
 * (SbrRenderThread.cxx:68-72). Footprints for every distinct ceil(radius) must have been registered with
 * pb_fbrush_register_footprint. Independent strokes run concurrently; conflicting ones keep their order. */
typedef struct pb_stroke {
  double radius;
  double K[3], S[3];
  int64_t first_imprint; /* into cx/cy/theta */
  int64_t n_imprints;
} pb_stroke;
int pb_fbrush_register_footprint(pb_fbrush* b, double radius, int side, const double* footprint);
int pb_fbrush_stroke_batch(pb_fbrush* b, pb_canvas* c, int64_t n_strokes, const pb_stroke* strokes, int64_t n_imprints,
                           const double* cx, const double* cy, const double* theta);
/* The same in two steps. pb_fbrush_plan_stroke_batch does all the host work of a batch — dataflow graph, claim order, stroke
 * records, per-imprint constants — and touches no stream, so the plan of the next batch can be made (also from another host
 * thread) while the device executes the current one; pb_fbrush_run_batch_plan uploads the plan and launches. dist = NULL
 * for a single GPU, or the descriptor pb_fbrush_stroke_batch_dist takes (run inside the same dist_begin / barrier bracket).
 * A plan is made from the brush's state (radius, registered footprints, snapshot switch) at plan time and fails as stale
 * if the radius differs when it is run; it can be run more than once. */
typedef struct pb_batch_plan pb_batch_plan;
struct pb_dist_desc;
int pb_fbrush_plan_stroke_batch(pb_fbrush* b, pb_canvas* c, const struct pb_dist_desc* dist, int64_t n_strokes,
                                const pb_stroke* strokes, int64_t n_imprints, const double* cx, const double* cy,
                                const double* theta, pb_batch_plan** out);
int pb_fbrush_run_batch_plan(pb_fbrush* b, pb_canvas* c, const struct pb_dist_desc* dist, const pb_batch_plan* plan);
int pb_batch_plan_destroy(pb_batch_plan* plan);
/* Host-side figures of the brush's last stroke or imprint batch (no device access):
 * [0] dataflow planning ms (segments + claim order), [1] per-imprint constants ms, [2] strokes planned (all ranks),
 * [3] dataflow segments, [4] wait entries, [5] the planner's model of the batch duration in ms (0 if the queue order
 * was not planned), [6] strokes executed by this rank, [7] kernel launches of this rank, [8] / [9] / [10] resident
 * thread-block clusters (= concurrent strokes, cudaOccupancyMaxActiveClusters) of this rank's launch shapes for footprints
 * of <= 256 / <= 4096 / more active cells (0 where the batch has none), [11] threads per cluster of the largest shape. */
#define PB_BATCH_STATS 12
int pb_fbrush_batch_stats(const pb_fbrush* b, double out[PB_BATCH_STATS]);
int pb_batch_plan_stats(const pb_batch_plan* plan, double out[PB_BATCH_STATS]);

/* Diagnostics: SM cycle stamps (clock64) inside the imprint kernel for the first 256 imprints of the first stroke of each
 * launch, taken by the first and the last thread of the stroke's first CTA: out[imprint][thread 0|1][8] =
 * {imprint start, snapshot ring done, pickup/deposit done, next interaction list built, barrier passed, 0, 0, 0}. */
int pb_fbrush_enable_trace(pb_fbrush* b, int enable);
int pb_fbrush_read_trace(pb_fbrush* b, uint64_t* out /* 256*2*8 */);
/* Stroke-pixel counters since creation: visited = the reference's `counter` (:119), cells passing both bounds
 * checks; active = those with footprint height > 0. `visited` is only maintained while counting is enabled
 * (it costs a pass over all footprint cells). */
int pb_fbrush_enable_visited_count(pb_fbrush* b, int enable);
int pb_fbrush_counters(pb_fbrush* b, uint64_t* visited, uint64_t* active);

/* ---- multi-GPU: one canvas cut into row bands over the GPUs of an NVSwitch node -------------------- */
/* Rank r (one process per GPU) owns rows [r*rows_per_band, min((r+1)*rows_per_band, rows)) as a band canvas
 * (pb_canvas_create_band(..., halo 0)). Compose / dry / texture strokes need no exchange. Footprint strokes are
 * executed entirely by the rank whose band holds their first imprint; where their footprint or snapshot ring
 * leaves the band, the kernel reads and writes the neighbour's HBM directly through NVLink peer mappings, and
 * strokes wait on completion flags of conflicting earlier strokes on any GPU (system-scope acquire/release).
 * The host exchanges the CUDA IPC handles (e.g. torch.distributed.all_gather_object) and fills pb_dist_desc. */
#define PB_IPC_HANDLE_BYTES 64
#define PB_MAX_BANDS 8
typedef struct pb_dist_desc {
  int32_t world, rank, rows_per_band, reserved;
  /* Per rank, the four buffers pb_fbrush_dist_storage returns on that rank: [rank] = own allocation, others =
   * pb_ipc_import'ed. canvas_base / snapshot_base are pixel-RECORD arrays (8 elements per pixel) of the rank's band. */
  void* canvas_base[PB_MAX_BANDS];
  int64_t canvas_stride[PB_MAX_BANDS];   /* unused, 0 */
  void* snapshot_base[PB_MAX_BANDS];
  int64_t snapshot_stride[PB_MAX_BANDS]; /* unused, 0 */
  void* dirty_base[PB_MAX_BANDS];
  void* flags_base[PB_MAX_BANDS];
} pb_dist_desc;
int pb_ipc_export(pb_context* ctx, void* dev_ptr, unsigned char handle[PB_IPC_HANDLE_BYTES]);
int pb_ipc_import(pb_context* ctx, const unsigned char handle[PB_IPC_HANDLE_BYTES], void** dev_ptr);
int pb_ipc_close(pb_context* ctx, void* dev_ptr);
int pb_canvas_storage(pb_canvas* c, void** base, int64_t* plane_stride_bytes);
/* Makes sure the brush's working record copy of the band's wet layer, its snapshot buffer (full copy of the band,
 * FootprintBrush.hxx:281-284), dirty map and flag buffer exist for this canvas and returns their base pointers for
 * export. */
int pb_fbrush_dist_storage(pb_fbrush* b, pb_canvas* c, void** canvas_records, void** snapshot_records, void** dirty_base,
                           void** flags_base);
/* A distributed batch runs on the ranks' record copies of their bands, which peers read and write through NVLink:
 *   pb_fbrush_dist_begin (planes -> records, synchronises the stream)  ->  process-group barrier  ->
 *   pb_fbrush_stroke_batch_dist, any number of times                  ->  context sync + process-group barrier  ->
 *   pb_fbrush_dist_end (records -> planes).
 * pb_fbrush_stroke_batch_dist has the contract of pb_fbrush_stroke_batch; every rank passes the SAME global stroke list.
 * Brush state after the call: radius and paint are those of the last stroke on every rank; the pickup map of the last stroke
 * is only written back on the rank that executed it — continue with a dip (as every stroke of a batch does) before relying
 * on it. */
int pb_fbrush_dist_begin(pb_fbrush* b, pb_canvas* c);
int pb_fbrush_dist_end(pb_fbrush* b, pb_canvas* c);
int pb_fbrush_stroke_batch_dist(pb_fbrush* b, pb_canvas* c, const pb_dist_desc* dist, int64_t n_strokes, const pb_stroke* strokes,
                                int64_t n_imprints, const double* cx, const double* cy, const double* theta);

/* Final assembly of the reflectance image (SURVEY.md §8e "gather of reflectance bands") without a separate collective: every
 * rank that wants the image creates a pb_band_image (3 planes r,g,b of rows*cols elements of the context's type in ONE
 * allocation), exports its base with pb_ipc_export and imports the others'. pb_canvas_compose_gather composes the
 * canvas' own band (Renderer::compose) and stores the rows straight into the n_dst given images — its own and / or peers'
 * over NVLink — at the band's position (gather to a root: n_dst = 1; all-gather: one entry per rank). All images share
 * one plane stride (pb_band_image_device). An image is complete once every rank's call has finished: synchronise the
 * contexts and run a process-group barrier before reading it (pb_band_image_download = host AoS f64, or the device base). */
typedef struct pb_band_image pb_band_image;
int pb_band_image_create(pb_context* ctx, int rows, int cols, pb_band_image** out);
int pb_band_image_destroy(pb_band_image* im);
int pb_band_image_device(pb_band_image* im, void** base, int64_t* plane_stride_bytes);
int pb_band_image_download(pb_band_image* im, double* out);
int pb_canvas_compose_gather(pb_canvas* band_canvas, int n_dst, void* const* dst_base, int64_t plane_stride_bytes);

/* ---- TextureBrush -------------------------------------------------------------------------------- */
/* thickness map: rows*cols host f64 (BrushStrokeSample::getThicknessMap). */
int pb_tbrush_create(pb_context* ctx, int map_rows, int map_cols, const double* thickness_map, pb_tbrush** out);
int pb_tbrush_destroy(pb_tbrush* b);
/* Brush-texture atlas. The sbr renderer picks a thickness texture per stroke from data/textures
 * (renderer/src/TextureBrushDictionary.cxx:25-69; gray, min-max normalised by the loader, :71-79); on the CPU path that is
 * BrushStrokeSample::setThicknessMap (renderer/BrushStrokeSample.hxx:25) before paintStroke. Textures live in HBM next to
 * the constructor's sample map (id 0); pb_tbrush_add_texture returns ids 1, 2, ... in call order. A batch names the
 * texture per stroke (pb_tstroke.texture_id); pb_tbrush_paint_stroke uses the one chosen with pb_tbrush_select_texture. */
int pb_tbrush_add_texture(pb_tbrush* b, int map_rows, int map_cols, const double* thickness_map, int* texture_id);
int pb_tbrush_select_texture(pb_tbrush* b, int texture_id);
int pb_tbrush_texture_count(const pb_tbrush* b);
int pb_tbrush_set_radius(pb_tbrush* b, double radius); /* TextureBrush.hxx:33-41 */
int pb_tbrush_dip(pb_tbrush* b, const double K[3], const double S[3]);
int pb_tbrush_set_thickness_scale(pb_tbrush* b, double scale); /* BrushBase.hxx:24-30 */
/* TextureBrush::enableSmudge (TextureBrush.hxx:207) + renderer/Smudge.hxx. Off by default here (the CPU class
 * defaults to on, sbr_painter's config to off). The smudge windows are (re)created by the next set_radius that acts
 * while smudge is enabled, exactly like the reference (TextureBrush.hxx:36-40). With smudge on, strokes form a serial
 * chain (the windows carry over from stroke to stroke). */
int pb_tbrush_enable_smudge(pb_tbrush* b, int enable);
/* TextureBrush::paintStroke (TextureBrush.hxx:52-205), path = n*2 doubles. */
int pb_tbrush_paint_stroke(pb_tbrush* b, pb_canvas* c, int n, const double* path_xy);
typedef struct pb_tstroke {
  double radius;
  double K[3], S[3];
  double thickness_scale;
  int64_t first_vertex; /* into path_xy (pairs) */
  int32_t n_vertices;
  int32_t texture_id; /* thickness texture of this stroke: 0 = the constructor's sample map, k = pb_tbrush_add_texture's id */
} pb_tstroke;
int pb_tbrush_stroke_batch(pb_tbrush* b, pb_canvas* c, int64_t n_strokes, const pb_tstroke* strokes, int64_t n_vertices,
                           const double* path_xy);
int pb_tbrush_counters(pb_tbrush* b, uint64_t* pixels);


/* ---- TextureBrushDictionary (host only) --------------------------------------------------------------- */
/* renderer/src/TextureBrushDictionary.cxx: the textures of data/textures are named <size>_<lengthClass>_<nn>.png and
 * grouped by (size key, length key), both ascending (:103-118); per size group the average texture height
 * (_avgSizes, :120-137) and per (size, length) group the average texture width (_avgTexLength, :139-164) are kept.
 * pb_texdict_create takes, per texture, its two file-name keys and its rows / cols; entry i of the dictionary is texture
 * i of the caller (e.g. pb_tbrush_add_texture id i + 1). */
typedef struct pb_texdict pb_texdict;
int pb_texdict_create(int n, const int32_t* size_key, const int32_t* length_key, const int32_t* rows, const int32_t* cols,
                      pb_texdict** out);
int pb_texdict_destroy(pb_texdict* d);
/* TextureBrushDictionary::lookup (:25-69) up to the random draw: stroke length = polyline length of path (n*2 doubles),
 * size group = nearest average height to brush_size, length group = nearest average width to the stroke length — with the
 * reference's quirks kept (the running minima start at _avgSizes[0] / _avgTexLength[i0][0] as VALUES, not distances, and
 * the defaults are i0 = 0, i1 = 1; the length loop runs over the number of SIZE groups, clamped here to the group's own
 * length classes where the reference would index out of range). Writes the group indices and up to `capacity` candidate
 * entries; *n_candidates receives their number (an empty group fails like the reference's runtime_error). The reference
 * then draws one candidate with std::random_device (:61-66) — the caller draws and records the pick in its stroke list. */
int pb_texdict_lookup(const pb_texdict* d, int n, const double* path_xy, double brush_size, int32_t* size_group,
                      int32_t* length_group, int capacity, int32_t* candidates, int32_t* n_candidates);

#ifdef __cplusplus
}
#endif
#endif /* PAINTY_B200_H */
