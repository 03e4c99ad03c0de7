/*
 * lst_b200.h -- C ABI of the B200-native tracking hot path of Laser Spot Track 4.
 *
 * Drop-in boundary.  The library replaces what sits behind these methods
 * of /root/reference/src/main/java/laser_spot_track4/Laser_Spot_Track4.java ("LST4"):
 *
 *   private int  analyseSlice(int slice, ImageProcessor slice_proc)            LST4:1302
 *   public static double[] doMatch_coord_res(ImageProcessor, ImageProcessor,
 *                                            int method, boolean subPix, ...)   LST4:2502
 *   public static double doMatch_test(ImageProcessor src, int method)          LST4:2724
 *   private void calcDisplacement()                                            LST4:2318
 *
 * and, beneath them, the JavaCPP/JNI calls
 *   opencv_imgproc.matchTemplate(Mat, Mat, Mat, int)                           LST4:2560, 2763
 *   opencv_core.minMaxLoc(Mat, DoublePointer, DoublePointer, Point, Point, Mat) LST4:2670
 * together with the ImageJ steps that precede them on every window
 *   ImagePlus.crop / ImageConverter.convertToGray32 / GaussianBlur.blurGaussian(ip,2,2,0.02)
 *                                                                              LST4:1346-1347, 1456-1473
 *   ImageProcessor.getBufferedImage (float -> 8-bit)                           LST4:2529-2533
 *
 * Rules of the ABI: plain C, opaque handle, int status returns (0 = OK, negative = error),
 * no exceptions cross the boundary, the caller owns every buffer it passes, the library
 * owns only what lst_create allocates.  One handle is used by one thread at a time;
 * different handles are independent.  There is NO CPU fallback: every entry point that
 * computes needs a CUDA device (sm_100a) and fails with LST_ERR_CUDA without one.
 *
 * Template / mark order everywhere: 0 = spot, 1 = mark1, 2 = mark2, 3 = mark3, 4 = mark4.
 */
#ifndef LST_B200_H
#define LST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LST_ABI_VERSION 3

/* status codes */
#define LST_OK               0
#define LST_ERR_INVALID     -1   /* bad argument */
#define LST_ERR_UNSUPPORTED -2   /* e.g. 16-bit stacks with matchIntensity=false (LST4:2510-2523), tiled / RGB slice files */
#define LST_ERR_CUDA        -3   /* no device / CUDA runtime error (see lst_last_error) */
#define LST_ERR_NOMEM       -4
#define LST_ERR_STATE       -5   /* call order (e.g. track before set_reference) */
#define LST_WOULD_BLOCK      1   /* lst_submit: the queue is full; lst_fetch: nothing finished yet (not an error) */

/* pixel types == ImageProcessor.getBitDepth() of the stack (LST4:172: DOES_8G, DOES_16, DOES_32) */
#define LST_PIX_U8   8
#define LST_PIX_U16 16
#define LST_PIX_F32 32
#define LST_PIX_RGB 24          /* ColorProcessor pixels: one int32 0x00RRGGBB per pixel (DOES_RGB, LST4:172) */

/* float -> 8-bit range convention of FloatProcessor.getBufferedImage (SURVEY.md 8c "known unknown") */
#define LST_RANGE_DISPLAY 0      /* u8 and RGB: [0,255]; u16/f32: crop's own pre-blur min/max (default) */
#define LST_RANGE_CROP    1      /* always the crop's own pre-blur min/max */
#define LST_RANGE_FIXED   2      /* [range_lo, range_hi] */
#define LST_RANGE_FRAME   3      /* the whole slice's pre-blur min/max (what the crop inherits if createProcessor hands the
                                    parent's min/max down): given per slice with lst_set_frame_ranges
                                    (slice_proc.getMin() / getMax(), LST4:1308-1312) or reduced on the device */
#define LST_QUANT_ROUND255 0     /* (int)(v*255/(max-min) + 0.5f) */
#define LST_QUANT_TRUNC256 1     /* (int)(v*256/(max-min)) */

/* frame status == return value of analyseSlice (LST4:1886-1887, 2315) */
#define LST_STATUS_OK   0
#define LST_STATUS_SKIP 1
#define LST_STATUS_STOP 2

typedef struct lst_ctx lst_ctx;

/* The outputs of getUserParameters (LST4:2370-2425) plus stack geometry. */
typedef struct lst_params {
    int32_t struct_size;       /* sizeof(lst_params), for ABI evolution */
    int32_t width, height;     /* slice size (LST4:813-817 checks every slice against it) */
    int32_t pixel_type;        /* LST_PIX_* */
    int32_t method;            /* 0..5 as in the dialog (LST4:2381); 5 = TM_CCOEFF_NORMED is the default (LST4:82) */
    int32_t templ_size;        /* templSize LST4:84 */
    int32_t s_area;            /* sArea LST4:83; 0 => whole-slice search (LST4:1316-1320) */
    int32_t sub_pixel;         /* LST4:89 */
    int32_t match_intensity;   /* LST4:88; 0 on an RGB stack = 3-channel match (channels blurred separately, CV_8UC3 matchTemplate) */
    int32_t range_mode;        /* LST_RANGE_* */
    int32_t quant_mode;        /* LST_QUANT_* */
    int32_t doubling;          /* search-area doubling recovery (LST4:1666-1735, 2058-2128) */
    int32_t auto_skip;         /* alwaysAutoSkip; headless policy: rejected frame => LST_STATUS_SKIP */
    int32_t max_s_area;        /* maxSArea LST4:131 (0 = no limit) */
    int32_t max_batch;         /* frames resolved per device pass (0 = default) */
    int32_t ingest_mode;       /* host slices: 0 = auto (only the search regions cross PCIe: from pinned memory as
                                  strided 3-D DMA copies, from pageable memory -- e.g. Java arrays held through
                                  JNI -- packed into pinned staging by reader threads), 1 = always whole-slice
                                  copies, 2 = search regions gathered by a kernel from mapped host memory,
                                  3 = always packed by reader threads */
    double  mark_dist;         /* markDist LST4:85 */
    double  range_lo, range_hi;/* LST_RANGE_FIXED */
    double  match_threshold;   /* 0 = matchThreshold[method] of LST4:147 ({0.1, 0.1, 0.05, 0.05, 0.2, 0.2}) */
    int32_t ncc_engine;        /* match kernel: 0 = auto (tensor cores when the template fits, else IDP.4A),
                                  1 = IDP.4A (lst_ncc_match), 2 = tensor cores (lst_ncc_tc); identical results */
    int32_t reserved1;
    double  rgb_weights[3];    /* RGB -> grey weights of convertToGray32 (LST4:519-521); all 0 = ImageJ's default 1/3 each */
} lst_params;

/* One template's match in one frame: the double[3] of doMatch_coord_res plus context. */
typedef struct lst_template_match {
    double  x, y;              /* coord_res[0], coord_res[1]: result-map coords (+ sub-pixel) */
    double  score;             /* coord_res[2]: double of the float32 map value */
    int32_t ix, iy;            /* integer peak (minMaxLoc, first maximum in row-major order) */
    int32_t wx, wy, ww, wh;    /* search window used: xStart, yStart, sWX, sWY (LST4:1327-1345) */
    int32_t s_used;            /* sArea or the doubled sArea_new the match was accepted at */
    int32_t accepted;          /* testMatchResult (LST4:1225-1255) */
} lst_template_match;

/* Everything analyseSlice + calcDisplacement produce for one slice ("output record"). */
typedef struct lst_record {
    int64_t frame;             /* caller's frame index */
    int32_t status;            /* LST_STATUS_* */
    int32_t reject;            /* template that rejected the frame, -1 = none */
    lst_template_match tm[5];
    double  dis[10];           /* disX/disY per template after this frame (LST4:100-108) */
    double  dX_pix, dY_pix;    /* LST4:2321-2322 */
    double  x_abs, y_abs, dL;  /* LST4:2346-2364: ResultsTable columns X_abs, Y_abs, dL (LST4:872-875) */
    int32_t passes;            /* device passes this frame's batch needed (diagnostic) */
    int32_t reserved;
} lst_record;

/* What template selection (LST4:486-702) leaves behind. */
typedef struct lst_reference_info {
    int32_t rect[5][4];        /* *_rect: x, y, width, height (LST4:512-515) */
    double  mideal[5];         /* doMatch_test of each template (LST4:566, 2050) */
    lst_record ref_record;     /* calcDisplacement() with all dis = 0 (LST4:702): row 0 of the table */
} lst_reference_info;

typedef struct lst_stats {
    int64_t frames;            /* frames resolved */
    int64_t tasks;             /* (frame, template) windows computed on the device */
    int64_t respeculated;      /* of those, recomputed because the window origin was mis-speculated */
    int64_t passes;            /* device passes */
    int64_t kernel_launches;   /* kernels of this library launched */
    int64_t h2d_bytes, d2h_bytes;
    double  ncc_kernel_ms;     /* CUDA-event time accumulated over lst_ncc_match launches (if timing on) */
    int64_t ncc_kernel_launches;
    double  ncc_algorithmic_macs; /* sum over timed launches of Rw*Rh*Tw*Th */
    int64_t roi_fallback_tasks;/* ROI ingest: windows read straight from pinned host memory (margin exceeded) */
    double  device_wait_ms;    /* host time blocked in stream synchronisation (device busy) */
    double  host_ms;           /* host time inside the tracking entry points, total */
    double  preprocess_ms;     /* CUDA-event time of lst_preprocess (+ lst_minmax), lst_box_sums and lst_epilogue over the */
    double  box_sums_ms;       /* same timed passes as ncc_kernel_ms (if timing on) */
    double  epilogue_ms;
    double  preprocess_px;     /* pixels blurred in the timed passes (38 flops each: the roofline of lst_preprocess*) */
    int64_t preprocess_launches;
} lst_stats;

/* ---- lifecycle ---------------------------------------------------------------------- */
int lst_abi_version(void);
/* Number of usable CUDA devices (0 without a driver/GPU). */
int lst_device_count(void);
/* Replaces the plugin's fields set by getUserParameters; validates like the dialog path does. */
int lst_create(const lst_params *params, int device, lst_ctx **out);
void lst_destroy(lst_ctx *ctx);
/* UTF-8, owned by the library; ctx == NULL returns the last creation error of this thread. */
const char *lst_last_error(const lst_ctx *ctx);

/* ---- template selection: LST4:486-702 ------------------------------------------------ */
/* ref_frame: host pointer, width*height pixels of pixel_type, row-major, no padding (ImageJ slice
 * array).  ref_xy: refX/refY of spot, mark1..mark4 as Java floats (LST4:95-99).  Crops, converts,
 * blurs and quantises the five templates on the device, computes their mideal, latches
 * spotX0/spotY0 (LST4:2356-2360) and resets all displacements to 0. */
int lst_set_reference(lst_ctx *ctx, const void *ref_frame, const float ref_xy[10], lst_reference_info *out);

/* ---- the tracking loop body: analyseSlice + calcDisplacement, LST4:793-902 ----------- */
/* frames: host memory, n_frames slices, consecutive slices frame_stride_bytes apart (0 = tightly
 * packed).  Slices are analysed in order with the reference's frame-to-frame recurrence
 * (LST4:1327-1328, 832-846); out receives n_frames records.  Synchronous.  Pageable or pinned
 * memory both work, and in both cases only the five search regions of each slice (plus a margin) cross
 * PCIe: from pinned memory (lst_alloc_pinned / lst_host_register) as strided 3-D DMA copies on a copy
 * stream while the previous batch computes; from pageable memory (e.g. Java arrays held through JNI) the
 * rows of the regions are packed into pinned staging by reader threads first.  Whole slices are copied
 * only on request (ingest_mode 1), for whole-slice search (sArea = 0) and when LST_RANGE_FRAME has to
 * reduce the slices on the device. */
int lst_track(lst_ctx *ctx, const void *frames, size_t frame_stride_bytes, int64_t first_frame_index,
              int32_t n_frames, lst_record *out);
/* Same, scattered slices (one pointer per slice -- stack.getProcessor(i).getPixels()). */
int lst_track_ptrs(lst_ctx *ctx, const void *const *frame_ptrs, int64_t first_frame_index,
                   int32_t n_frames, lst_record *out);
/* Same, slices already resident in device memory of ctx's device (benchmark "value" leg). */
int lst_track_device(lst_ctx *ctx, const void *dev_frames, size_t frame_stride_bytes,
                     int64_t first_frame_index, int32_t n_frames, lst_record *out);
/* Same, one device pointer per slice (e.g. a ring of decoded slices reused cyclically). */
int lst_track_device_ptrs(lst_ctx *ctx, const void *const *dev_frame_ptrs, int64_t first_frame_index,
                          int32_t n_frames, lst_record *out);

/* LST_RANGE_FRAME: the [min, max] of the slices handed to the NEXT call that takes slices (lst_set_reference*: one pair for
 * the reference slice; lst_track*: n_frames pairs), e.g. slice_proc.getMin() / getMax().  lo_hi = n x {lo, hi}; consumed by
 * that call.  Without it the library reduces every slice on the device, which needs the whole slice there: host slices are
 * then copied whole instead of by search region, and slice files are refused (LST_ERR_UNSUPPORTED). */
int lst_set_frame_ranges(lst_ctx *ctx, const double *lo_hi, int32_t n);

/* Displacement state (disX/disY x5, LST4:100-108): read it, or overwrite it -- the caller's
 * "keep the result" override (LST4:1281-1300) and the chunk hand-off of a sharded run. */
int lst_get_state(const lst_ctx *ctx, double dis[10]);
int lst_set_state(lst_ctx *ctx, const double dis[10]);

/* ---- the same without blocking the caller (submit / fetch / override) ---------------
 * The plugin's loop (LST4:793-902) and its folder monitor (LST4:958-1132, analyseSlice at LST4:1031) decode one slice after
 * another; with these calls the decoder keeps running while earlier slices are on the device.  A worker thread owned by
 * the context tracks the submitted batches in order with lst_track_ptrs.  The slices of a batch must stay valid until its
 * records have been fetched.  While batches are queued or running the context must not be used through the synchronous
 * entry points (LST_ERR_STATE).
 *   lst_submit   queues n_frames slices (pointer list copied); LST_WOULD_BLOCK when `queue_depth` batches are already
 *                waiting or running (default 2: one on the device, one whose ingest overlaps it).
 *   lst_fetch    copies up to max_records finished records, in slice order, into out; *n_out = how many.  wait != 0 blocks
 *                until at least one record is there (or nothing is in flight); LST_WOULD_BLOCK with *n_out = 0 otherwise.
 *   lst_override the caller's verdict on a slice the library rejected (failureQuestionDlg "keep", LST4:1281-1300, or a
 *                state typed in by the user): everything submitted after `frame_index` is dropped (finished or not), the
 *                displacement state becomes `dis` and *resume_from receives the index the caller submits again from
 *                (frame_index + 1).  Blocks until the worker is idle. */
int lst_submit(lst_ctx *ctx, const void *const *frame_ptrs, int64_t first_frame_index, int32_t n_frames);
int lst_fetch(lst_ctx *ctx, lst_record *out, int32_t max_records, int32_t *n_out, int32_t wait);
int lst_override(lst_ctx *ctx, int64_t frame_index, const double dis[10], int64_t *resume_from);
int lst_set_queue_depth(lst_ctx *ctx, int32_t queue_depth);

/* ---- several GPUs in one process (north_star item 4) ----------------------------------- */
/* One lst_ctx, one host worker thread and one stream set per device.  lst_multi_track cuts the slices it
 * is given into contiguous chunks, one per device, and tracks them concurrently; nothing is reduced
 * between GPUs and there is no NCCL.  The only coupling is the recurrence (LST4:1327-1328): chunk r starts
 * from a speculated displacement state (refined by first tracking the last slices of chunk r-1), and
 * after all chunks finish the boundaries are verified in order on the host -- a chunk whose speculated
 * (int)dis was wrong is tracked again from the true state.  Records are identical to lst_track's. */
typedef struct lst_multi lst_multi;
int lst_multi_create(const lst_params *params, const int32_t *devices, int32_t n_devices, lst_multi **out);
void lst_multi_destroy(lst_multi *m);
const char *lst_multi_last_error(const lst_multi *m);
int lst_multi_set_reference(lst_multi *m, const void *ref_frame, const float ref_xy[10], lst_reference_info *out);
int lst_multi_track(lst_multi *m, const void *const *frame_ptrs, int64_t first_frame_index, int32_t n_frames,
                    lst_record *out);
/* the same with slices that already lie in device memory (every pointer readable from every device of m).
 * A device may be listed more than once in lst_multi_create: each entry is a lane with its own host
 * thread, streams and scratch, so the host-side chain walk of one lane overlaps the kernels of another. */
int lst_multi_track_device(lst_multi *m, const void *const *dev_frame_ptrs, int64_t first_frame_index,
                           int32_t n_frames, lst_record *out);
/* sums over the contexts of m */
int lst_multi_get_stats(const lst_multi *m, lst_stats *out);
int lst_multi_reset_stats(lst_multi *m);
int lst_multi_get_state(const lst_multi *m, double dis[10]);
int lst_multi_set_state(lst_multi *m, const double dis[10]);
/* chunks tracked a second time because their speculated start state placed a window differently */
int64_t lst_multi_retracked(const lst_multi *m);

/* ---- slices as files --------------------------------------------------------
 * The reference's virtual stack decodes one file per slice and opens it twice per frame
 * (LST4:808-817, 831).  These entry points take the files themselves: a slice file is mapped once
 * and only the rows of the five search regions are read and sent to the device.  Supported: classic
 * TIFF, either byte order (big-endian samples are swapped on the device), one sample per pixel, 8/16-bit
 * unsigned or 32-bit float, strips that are uncompressed or PackBits / LZW / Deflate compressed (with
 * or without the horizontal predictor; only the strips covering the search regions are decoded);
 * `pages` (nullable) selects a page of a multi-page file or a slice of an ImageJ stack file.  Grey 8/16-bit
 * PNG files (not interlaced) are read as well, and JPEG files -- the camera time-lapse of the EXIF path,
 * LST4:1199-1223: baseline / extended-sequential Huffman, 8-bit, grey (an 8-bit slice) or YCbCr 4:4:4 /
 * 4:2:2 / 4:2:0 (a ColorProcessor slice, LST_PIX_RGB), restart intervals; reconstructed with the IJG
 * library's default integer IDCT, triangle-filter upsampling and colour tables -- what the JDK decoder
 * behind ImageJ's opener computes -- for the MCUs under the search regions only: the reader threads do
 * the entropy decoding, the reconstruction runs on the device (LST_JPEG_DEVICE=0: on the reader threads).  Other files (progressive
 * or arithmetic-coded JPEG, tiled or RGB TIFF / PNG ...): LST_ERR_UNSUPPORTED -- decode them and use
 * lst_track_ptrs. */
typedef struct lst_tiff_info {
    int32_t width, height, bits, sample_format; /* sample_format: 1 unsigned integer, 3 IEEE float; bits 24 = 0x00RRGGBB pixels (int32) */
    int32_t big_endian, rows_per_strip, n_strips;
    int32_t imagej_images;                      /* "images=N" of an ImageJ description, else 0 */
    int32_t compression, predictor;             /* TIFF tags 259 and 317; a JPEG file reports compression 7 */
    int64_t first_strip_offset;
} lst_tiff_info;
/* host only, no GPU needed; err (nullable) receives a message */
int lst_tiff_probe(const char *path, int32_t page, lst_tiff_info *out, char *err, int32_t err_cap);
/* host only: the pixels of the rectangle [x, x+w) x [y, y+h) of a slice file in native byte order,
 * row pitch w * bits/8 (what the reader threads of lst_track_files fetch for one search region) */
int lst_tiff_read_rect(const char *path, int32_t page, int32_t x, int32_t y, int32_t w, int32_t h, void *out,
                       char *err, int32_t err_cap);
int lst_set_reference_file(lst_ctx *ctx, const char *path, int32_t page, const float ref_xy[10],
                           lst_reference_info *out);
int lst_track_files(lst_ctx *ctx, const char *const *paths, const int32_t *pages,
                    int64_t first_frame_index, int32_t n_frames, lst_record *out);

/* ---- output side ------------------------------------------------------------
 * The ResultsTable rows of LST4:864-875 for a whole batch of records in one call (the reference
 * appends one row and rebuilds the plot per frame).  Skipped frames add no row (LST4:845).
 * rows: n_records x 7 doubles {Time, Frame, dX_pix, dY_pix, X_abs, Y_abs, dL}; Time advances by
 * time_step per analysed frame from time0 (the constant-step time base of LST4:2202-2215), Frame
 * is the table's row counter continuing from first_row_number; src_index (nullable) receives the
 * index of the record each row came from.  Returns the number of rows, or a negative status. */
int64_t lst_results_rows(const lst_record *records, int64_t n_records, double time0, double time_step,
                         int64_t first_row_number, double *rows, int64_t *src_index);

/* ---- the primitives, exposed like the reference's public statics ---------------------- */
/* doMatch_coord_res(src, tpl, method, subPix, null) on the 8-bit images that reach OpenCV
 * (LST4:2502-2563, 2665-2722); method 0..5 as in the dialog (LST4:2381).  out[3] = {x, y, score}
 * (the minimum of the map for methods 0 and 1, else the maximum); score_map (nullable) receives
 * the (sh-th+1) x (sw-tw+1) float32 result of matchTemplate. */
int lst_match_single(lst_ctx *ctx, const uint8_t *src8, int32_t sw, int32_t sh,
                     const uint8_t *tpl8, int32_t tw, int32_t th, int32_t method, int32_t sub_pix,
                     double out[3], float *score_map);
/* doMatch_test(src, method): matchTemplate(src, src) 1x1 (LST4:2724-2767). */
int lst_match_test(lst_ctx *ctx, const uint8_t *src8, int32_t sw, int32_t sh, int32_t method, double *out);
/* crop -> [convertToGray32] -> blurGaussian(2,2,0.02) -> 8-bit of one rectangle of a host frame:
 * the image handed to matchTemplate (LST4:1346-1347, 1456-1473, 2529).  out8: w*h bytes; for LST_PIX_RGB with
 * match_intensity = 0 (the 3-channel match, LST4:2525-2535 case 24) 3*w*h bytes: the blurred planes R, G, B in turn. */
int lst_preprocess(lst_ctx *ctx, const void *frame, int32_t x0, int32_t y0, int32_t w, int32_t h, uint8_t *out8);
/* calcDisplacement() as a pure function (LST4:2318-2365); spot0 NULL => first call (latches). */
int lst_calc_displacement(const float ref_xy[10], const double dis[10], double mark_dist,
                          const double spot0[2], double out5[5]);

/* ---- host memory helpers -------------------------------------------------------------- */
int lst_alloc_pinned(size_t bytes, void **out);
int lst_free_pinned(void *p);
int lst_host_register(void *p, size_t bytes);
int lst_host_unregister(void *p);
/* device memory for lst_track_device (benchmarks / callers that decode on the GPU) */
int lst_device_alloc(lst_ctx *ctx, size_t bytes, void **out);
int lst_device_free(lst_ctx *ctx, void *p);
int lst_device_upload(lst_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);

/* ---- diagnostics ---------------------------------------------------------------------- */
int lst_get_stats(const lst_ctx *ctx, lst_stats *out);
int lst_reset_stats(lst_ctx *ctx);
/* Time every lst_ncc_match launch with CUDA events on its stream (for bench.py's roofline). */
int lst_set_kernel_timing(lst_ctx *ctx, int32_t on);
/* Measured peaks of the device, for the roofline denominators: out[0] = int8 dp4a MAC/s
 * (IDP.4A.U8.U8), out[1] = fp32 FFMA/s, out[2] = int8 tensor-core MAC/s (tcgen05.mma kind::i8,
 * M128 N256 K32 from shared memory).  Runs ~ms_per_test per test. */
int lst_measure_alu_peaks(int device, double ms_per_test, double out[3]);

/* ---- test seams: host logic only, no device, no arithmetic of the path ----------------- */
/* Window placement of analyseSlice for template k (LST4:1325-1345): rect + (int)dis - sArea,
 * clamp by shift; out = xStart, yStart, sWX, sWY. */
int lst_host_window_for(const lst_params *p, const int32_t rect[4], double dis_x, double dis_y, int32_t out[4]);
/* Template rectangle from a click point (LST4:506-515). */
int lst_host_template_rect(const lst_params *p, float ref_x, float ref_y, int32_t out[4]);
/* One (frame, template) window to evaluate, and its result: the unit of work of the resolver. */
typedef struct lst_task {
    int32_t frame;             /* index inside the batch */
    int32_t tpl;               /* 0..4 */
    int32_t wx, wy, ww, wh;    /* window */
    int32_t s_used;            /* search area the accept test uses */
    int32_t reserved;
} lst_task;
typedef struct lst_task_result {
    double  x, y, score;
    int32_t ix, iy;
    int32_t accepted;
    int32_t reserved;
} lst_task_result;
typedef int (*lst_compute_fn)(void *user, const lst_task *tasks, int32_t n_tasks, lst_task_result *results);
/* Runs the library's recurrence resolver (speculate window origins for a whole batch, verify the
 * chain on the host, recompute only mis-speculated windows) with `compute` standing in for the
 * device.  Used by the CPU tests to prove the resolver reproduces the sequential loop of
 * LST4:793-846; the product path (lst_track*) drives the same resolver with the CUDA executor. */
int lst_host_resolve(const lst_params *p, const int32_t rect[5][4], const float ref_xy[10],
                     const double spot0[2], double dis_io[10], int64_t first_frame_index, int32_t n_frames,
                     lst_compute_fn compute, void *user, lst_record *out, lst_stats *stats_io);

#ifdef __cplusplus
}
#endif
#endif /* LST_B200_H */
