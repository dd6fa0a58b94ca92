/* esd.h -- C ABI of libesd.so: B200-native scene-detection scoring.
 *
 * The reference (codihuston/eioku) is pure Python and has NO FFI for this path; its
 * scene task is `ModelManager.detect_scenes(video_path, config)`
 * (/root/reference/ml-service/src/services/model_manager.py:715-835), specified to run
 * PySceneDetect's ContentDetector / AdaptiveDetector / HistogramDetector
 * (/root/reference/README.md:56, .kiro/specs/semantic-video-search/design.md:994-1007;
 * producer "scenedetect", ml-service/src/models/responses.py:141-142).  This ABI is
 * therefore defined by what a Python `ctypes` binding needs to implement the
 * SceneDetector plugin surface (process_frame / post_process) and the scene-task
 * schema on top of device-resident uint8 BGR frames.
 *
 * Conventions
 *  - plain C, POD only; every function returns ESD_OK (0) or a negative esd_status,
 *    never throws/aborts; esd_last_error() gives the detail string of the last failure.
 *  - the caller owns frame memory and every output buffer; the library owns scratch,
 *    tables, score arrays, cut lists and the pinned ingest ring.
 *  - one ctx per (device, video stream); not thread-safe; no global state.
 *  - work is enqueued on the caller's CUDA stream (`void* stream` = cudaStream_t);
 *    only esd_read_scores / esd_get_cuts / esd_synchronize / esd_kernel_time wait.
 *  - there is NO CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef ESD_H
#define ESD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESD_ABI_VERSION 4

#if defined(__GNUC__)
#define ESD_API __attribute__((visibility("default")))
#else
#define ESD_API
#endif

typedef struct esd_ctx esd_ctx;

typedef enum esd_status {
    ESD_OK = 0,
    ESD_ERR_INVALID = -1,     /* bad argument / config */
    ESD_ERR_CUDA = -2,        /* a CUDA runtime call failed (see esd_last_error) */
    ESD_ERR_NOMEM = -3,
    ESD_ERR_STATE = -4,       /* call out of order (e.g. non-sequential frame numbers; a push after a half-failed push) */
    ESD_ERR_UNSUPPORTED = -5, /* e.g. destination width > 1024 when resizing, odd-size hash DCT, NV12 without a downscale */
    ESD_ERR_CAPACITY = -6     /* caller buffer or cut list too small */
} esd_status;

/* detector bitmask: which decision passes (and therefore which per-frame features) run */
enum { ESD_DET_CONTENT = 1, ESD_DET_ADAPTIVE = 2, ESD_DET_HIST = 4, ESD_DET_THRESHOLD = 8, ESD_DET_HASH = 16 };
/* ThresholdDetector.Method */
enum { ESD_THRESH_FLOOR = 0, ESD_THRESH_CEILING = 1 };
/* FlashFilter.Mode of PySceneDetect >= 0.6.4; SUPPRESS == the legacy (<= 0.6.3) min_scene_len rule */
enum { ESD_FILTER_MERGE = 0, ESD_FILTER_SUPPRESS = 1 };
/* compute_downscale_factor: W/256.0 (>= 0.6.2) or W//256 (<= 0.6.1) */
enum { ESD_DOWNSCALE_FLOAT = 0, ESD_DOWNSCALE_INT = 1 };
/* work decomposition of the fused kernel (tuning; results are identical) */
enum { ESD_SPLIT_AUTO = 0, ESD_SPLIT_STRIPS = 1, ESD_SPLIT_CHUNKS = 2 };
/* pixel format of the frames handed to the library.  NV12 (what NVDEC and most hardware decoders emit: a Y plane followed
 * by an interleaved half-resolution UV plane) is converted exactly like cv2.cvtColor(COLOR_YUV2BGR_NV12) inside the fused
 * kernel, only for the source pixels the downscale taps read; it needs a downscaling context and even dimensions.
 * I420 (planar YUV 4:2:0 -- what software decoders emit: ffmpeg / PyAV `yuv420p`: Y plane, U plane, V plane) is the same
 * arithmetic, cv2.cvtColor(COLOR_YUV2BGR_I420): the chroma rows the taps touch are interleaved on the device first. */
enum { ESD_FMT_BGR24 = 0, ESD_FMT_NV12 = 1, ESD_FMT_I420 = 2 };

typedef struct esd_config {
    uint32_t struct_size;      /* = sizeof(esd_config) */
    int32_t detectors;         /* ESD_DET_* bitmask, != 0 */
    int32_t src_width;         /* frame geometry handed to esd_push_frames */
    int32_t src_height;
    int32_t dst_width;         /* SceneManager resize target; 0,0 = auto (downscale_mode); == src = no resize */
    int32_t dst_height;
    int32_t downscale_mode;    /* ESD_DOWNSCALE_*; used when dst_* == 0 */
    int32_t src_format;        /* ESD_FMT_*; 0 = BGR24 */

    /* ContentDetector(threshold, min_scene_len, weights, filter_mode); luma_only = weights {0,0,1,0} */
    double content_threshold;
    double content_weights[4]; /* delta_hue, delta_sat, delta_lum, delta_edges (> 0 enables the Canny/dilate edge map) */
    double content_weight_div; /* sum(abs(w)) as the host language computes it; <= 0 -> naive left-to-right */
    int32_t content_min_scene_len;
    int32_t content_filter_mode;

    /* AdaptiveDetector(adaptive_threshold, min_scene_len, window_width, min_content_val, weights) */
    double adaptive_threshold;
    double adaptive_min_content_val;
    double adaptive_weights[4];
    double adaptive_weight_div;
    int32_t adaptive_window_width;
    int32_t adaptive_min_scene_len;

    /* HistogramDetector(threshold, bins, min_scene_len) */
    double hist_threshold;     /* user threshold; the library applies clamp(1 - t, 0, 1) */
    int32_t hist_bins;         /* 1..256 */
    int32_t hist_min_scene_len;

    /* ThresholdDetector(threshold, min_scene_len, fade_bias, add_final_scene, method) -- fades through black */
    double thresh_threshold;   /* compared as int(threshold), like PySceneDetect */
    double thresh_fade_bias;   /* -1..1 */
    int32_t thresh_min_scene_len;
    int32_t thresh_add_final_scene;
    int32_t thresh_method;     /* ESD_THRESH_* */
    int32_t edge_kernel_size;  /* ContentDetector kernel_size for the edge dilation; 0 = PySceneDetect's estimate */

    /* tuning knobs, 0 = auto */
    int32_t rows_per_group;    /* destination rows one CTA keeps on-chip per frame */
    int32_t pipeline_stages;   /* TMA ring depth per CTA */
    int32_t split_mode;        /* ESD_SPLIT_* */
    int32_t ctas_per_sm;
    int32_t rows_per_stage;    /* destination rows staged per pipeline slot (1..4) */
    int32_t reserved1;         /* tuning: consumer lane stride in destination columns (1, 2, 4, 8); 0 = bank-conflict-free choice */
    int64_t max_cuts;          /* per-detector cut capacity (default 65536) */
    int64_t initial_capacity;  /* frames of per-frame score storage to pre-allocate (grows by doubling) */

    /* HashDetector(threshold, size, lowpass, min_scene_len) -- perceptual (DCT) hash, ABI >= 2 */
    double hash_threshold;     /* normalised Hamming distance, default 0.395 */
    int32_t hash_size;         /* hash is size x size bits (default 16; size * size <= 1024) */
    int32_t hash_lowpass;      /* DCT runs on a (size * lowpass)^2 INTER_AREA thumbnail (default 2; size * lowpass <= 64, even) */
    int32_t hash_min_scene_len;
    int32_t reserved2;         /* tuning: upper bound of the fused kernel's grid in CTAs (0 = every SM x occupancy) */
} esd_config;

/* derived geometry, for the caller's roofline accounting and ingest sizing */
typedef struct esd_geometry {
    int32_t dst_width, dst_height;
    int32_t n_touched_rows;      /* distinct source rows the vertical taps read */
    int32_t row_bytes;           /* bytes per source row: src_width * 3 (BGR24) or src_width (NV12) */
    int64_t alg_bytes_per_frame; /* algorithmic HBM bytes per frame: 32-byte sectors of the touched rows that hold a tap */
    int64_t compact_frame_bytes; /* n_touched_rows * row_bytes: one frame in the compact layout = bytes the kernel fetches */
    int32_t lane_stride;         /* destination columns between neighbouring consumer lanes (bank-conflict-free tap loads) */
    int32_t lane_stride_taps;    /* the same for the gathered-taps ingest layout */
} esd_geometry;

ESD_API int esd_abi_version(void);
ESD_API const char* esd_strerror(int status);
/* detail of the last failure on this ctx (or of the last esd_create failure when ctx == NULL) */
ESD_API const char* esd_last_error(const esd_ctx* ctx);
ESD_API int esd_device_count(void);

/* Replaces: detector construction + SceneManager set-up the spec calls for (README.md:56; design.md:994-1007); in
 * the shipped reference, the ffmpeg command line built at ml-service/src/services/model_manager.py:731-745
 * (config.get("threshold") :732 is the only knob read there). */
ESD_API void esd_config_default(esd_config* cfg); /* PySceneDetect defaults; detectors = CONTENT */
ESD_API int esd_create(esd_ctx** out, const esd_config* cfg, int device);
ESD_API void esd_destroy(esd_ctx* ctx);
/* forget all frames, scores, cuts and filter state (start of a new video) */
ESD_API int esd_reset(esd_ctx* ctx);
ESD_API int esd_get_geometry(const esd_ctx* ctx, esd_geometry* out);
/* touched source rows in ascending order (n_touched_rows entries); NV12: Y rows, then UV rows numbered src_height + r,
 * i.e. row indices into a contiguous NV12 frame */
ESD_API int esd_get_touched_rows(const esd_ctx* ctx, int32_t* rows, int32_t cap);

/* Replaces: the per-frame scoring loop -- in the shipped reference the ffmpeg child process started by
 * subprocess.run at model_manager.py:750-755 (`select='gt(scene,T)'` scores every decoded frame); as specified,
 * PySceneDetect's SceneManager loop `cv2.resize -> detector.process_frame`.
 * Score n device-resident frames (uint8 BGR, row pitch `pitch_bytes`, frame stride
 * `frame_stride_bytes`) and run the decision passes.  Frame numbers must be sequential
 * across calls: first_frame_num == (first frame of the first push) + frames pushed so far.
 * Asynchronous on `stream`. */
ESD_API int esd_push_frames(esd_ctx* ctx, const uint8_t* d_bgr, int64_t n, int64_t frame_stride_bytes,
                    int64_t pitch_bytes, int64_t first_frame_num, void* stream);
/* Replaces: the BGR conversion a decoder hand-off performs before scoring -- `cv2.VideoCapture.read()` delivers BGR made by
 * swscale from the decoder's YUV 4:2:0 output (the other tasks' decode loop, model_manager.py:237-263); here the decoder's
 * surface is consumed as it is and converted in the kernel with cv2.cvtColor(COLOR_YUV2BGR_NV12)'s arithmetic.
 * NV12 contexts: Y plane and interleaved UV plane given separately (as a decoder surface has them: the UV plane usually
 * starts at an aligned height); both use `pitch_bytes` and `frame_stride_bytes`.  esd_push_frames on an NV12 context is
 * the contiguous case d_uv = d_y + src_height * pitch_bytes.  This is the device half of SURVEY.md 8f N1 (decode ->
 * device): the scoring chain consumes decoder surfaces directly, 0.97 MB instead of 1.66 MB per 1080p frame. */
ESD_API int esd_push_nv12(esd_ctx* ctx, const uint8_t* d_y, const uint8_t* d_uv, int64_t n, int64_t frame_stride_bytes,
                          int64_t pitch_bytes, int64_t first_frame_num, void* stream);
/* I420 contexts: the three planes of frame 0 given separately (an AVFrame's data[0..2]); frame k's planes lie
 * k * frame_stride_bytes further; Y rows are pitch_y_bytes apart, U and V rows pitch_uv_bytes.  esd_push_frames (device) and
 * esd_ingest_push_host (host) on an I420 context take the contiguous case: Y plane, then src_height / 2 U rows and as many V rows
 * of pitch / 2 -- the (H * 3 / 2, W) array cv2.cvtColor(COLOR_YUV2BGR_I420) takes. */
ESD_API int esd_push_i420(esd_ctx* ctx, const uint8_t* d_y, const uint8_t* d_u, const uint8_t* d_v, int64_t n,
                          int64_t frame_stride_bytes, int64_t pitch_y_bytes, int64_t pitch_uv_bytes, int64_t first_frame_num,
                          void* stream);
/* Same, for frames stored in the compact layout [n][n_touched_rows][row_bytes] (touched rows only,
 * ascending source-row order; NV12 and I420: the touched Y rows, then the touched chroma rows as interleaved UV rows) that the
 * ingest ring produces. */
ESD_API int esd_push_rows(esd_ctx* ctx, const uint8_t* d_rows, int64_t n, int64_t first_frame_num, void* stream);

/* Replaces: one call of PySceneDetect's `SceneDetector.process_frame(frame_num, frame_img)` with a host (numpy) frame --
 * the plugin surface the reference names (README.md:56; design.md:994-1007) as PySceneDetect's own SceneManager loop
 * drives it, frame by frame.  The frame (src_width x src_height BGR, `pitch_bytes` per row; pageable or pinned) is
 * staged through a library-owned pinned buffer, copied to the device, scored and decided on one library-owned stream
 * with a single synchronisation; cuts [from_index, ...) of `detector` come back exactly like esd_get_cuts.  Frame
 * numbers follow the same sequencing rule as esd_push_frames.  Synchronises. */
ESD_API int esd_process_frame_host(esd_ctx* ctx, const uint8_t* h_bgr, int64_t pitch_bytes, int64_t frame_num, int32_t detector,
                                   int64_t from_index, int64_t* cuts, int64_t cap, int64_t* n_written, int64_t* n_total);

/* Replaces: the decoded-frame hand-off (`ffmpeg -i <path>` decode feeding the select filter, model_manager.py:736-745;
 * `cv2.VideoCapture.read()` in the other tasks, e.g. :237-263).  Decode itself stays outside.
 * Host frames -> pinned ring -> cudaMemcpyAsync on a copy stream -> scoring on the ctx's own
 * compute stream.  Only the touched source rows cross PCIe.  h_bgr may be pageable (staged through
 * the ring by the CPU) or pinned (DMA'd directly). */
ESD_API int esd_ingest_open(esd_ctx* ctx, int32_t n_slots, int32_t frames_per_slot);
ESD_API int esd_ingest_push_host(esd_ctx* ctx, const uint8_t* h_bgr, int64_t n, int64_t frame_stride_bytes,
                         int64_t pitch_bytes, int64_t first_frame_num);
ESD_API int esd_ingest_close(esd_ctx* ctx);
/* n_threads > 0: `esd_ingest_push_host` uses n_threads host threads to gather, per touched row, only the two BGR taps
 * every destination column reads (6 * dst_width bytes per row) into the pinned ring, so e.g. 442 KB instead of
 * 1.66 MB per 1080p frame cross PCIe; the host does no arithmetic.  0 (default) = DMA of whole touched rows, no CPU work. */
ESD_API int esd_ingest_set_gather(esd_ctx* ctx, int32_t n_threads);
/* Lifetime rule of esd_ingest_push_host: with PINNED caller memory (and no gather threads) the H2D copies read the caller's
 * frames directly and are still in flight when the call returns -- the frames must stay valid and unmodified until
 * esd_ingest_wait_copied (H2D copies done; scoring may still be running) or esd_synchronize returns.  Pageable frames and
 * the gather path are consumed before the call returns. */
ESD_API int esd_ingest_wait_copied(esd_ctx* ctx);
/* bytes moved host->device by the ingest path since esd_reset */
ESD_API int esd_ingest_stats(const esd_ctx* ctx, int64_t* h2d_bytes, int64_t* h2d_copies);

/* Wait for all enqueued work of this ctx. */
ESD_API int esd_synchronize(esd_ctx* ctx);
/* The finalize/decision tail of a push runs on a library-owned stream so that it overlaps the next
 * batch's fused kernel.  esd_join makes `stream` wait (on the device) for every tail enqueued so far. */
ESD_API int esd_join(esd_ctx* ctx, void* stream);
ESD_API int64_t esd_frames_pushed(const esd_ctx* ctx);

/* Replaces: nothing in the shipped reference (ffmpeg's per-frame scene score is discarded, model_manager.py:762);
 * PySceneDetect's StatsManager metrics in the specified design.
 * Per-frame results for frames [from_frame, from_frame + n) (absolute frame numbers).  Any output
 * pointer may be NULL.  sums3: [n][3] sum|dH|,sum|dS|,sum|dV| (0 for the first frame of the video);
 * content_val / adaptive_val: float64 score with the content / adaptive weights; adaptive_ratio: NaN
 * where the window is incomplete; hist: [n][bins] Y-histogram counts; hist_diff: NaN for the first frame.
 * Synchronises. */
ESD_API int esd_read_scores(esd_ctx* ctx, int64_t from_frame, int64_t n, uint64_t* sums3, double* content_val,
                    double* adaptive_val, double* adaptive_ratio, uint32_t* hist, double* hist_diff);

/* Number of pixels whose dilated edge bit differs from the previous frame (delta_edges = 255 * count / pixels);
 * only when a delta_edges weight is > 0.  Synchronises. */
ESD_API int esd_read_edge_counts(esd_ctx* ctx, int64_t from_frame, int64_t n, uint32_t* counts);

/* Replaces: nothing in the shipped reference (its ffmpeg scene filter has one fixed metric, model_manager.py:731-745);
 * PySceneDetect's HashDetector.process_frame / StatsManager metric `hash_dist [size=N lowpass=M]` in the specified design
 * (README.md:56 names PySceneDetect as the scene detector).
 * HashDetector's per-frame results: `bits` [n][ceil(size*size/32)] (bit i of the row-major size x size hash in word
 * i / 32, bit i % 32; padding bits zero) and `hash_dist` = Hamming distance to the previous frame / size^2 (NaN for the
 * first frame).  Either pointer may be NULL.  The integer stages (BGR2GRAY, INTER_AREA) are bit-exact with OpenCV; the
 * DCT is evaluated in float64, so a bit whose coefficient lies within float32 rounding noise of the median may differ
 * from a given cv2 build (cv2.dct itself differs between its IPP and plain paths).  Synchronises. */
ESD_API int esd_read_hash(esd_ctx* ctx, int64_t from_frame, int64_t n, uint32_t* bits, double* hash_dist);
/* How far each frame's hash is from flipping a bit: min over the size^2 bits of |DCT coefficient - median| (float32).  cv2.dct's
 * float32 output differs between OpenCV builds (IPP / plain) by 1-2 ulp of the coefficients (~1e-6 for this normalisation), so
 * a frame whose margin is below ~4e-6 has bits that a given cv2 build may set differently; every frame above it hashes
 * identically to cv2 (tests/test_gpu_hash.py).  This is the documented tolerance of the one detector that is not bit-exact by
 * construction (BASELINE north_star states none for it): bits with margin > 4e-6 are exact, the others are reported here.
 * Synchronises. */
ESD_API int esd_read_hash_margin(esd_ctx* ctx, int64_t from_frame, int64_t n, float* min_margin);
/* Test hook: the (size*lowpass)^2 uint8 INTER_AREA thumbnail the hash of frame `frame` was computed from; only frames of
 * the most recent push are available.  Synchronises. */
ESD_API int esd_debug_read_hash_input(esd_ctx* ctx, int64_t frame, uint8_t* out, int64_t cap);

/* ThresholdDetector's per-frame metric: average of all B,G,R values of the (downscaled) frame.  Synchronises. */
ESD_API int esd_read_average_rgb(esd_ctx* ctx, int64_t from_frame, int64_t n, double* average_rgb);

/* SceneDetector.post_process(last_frame_num): cuts a detector adds after the last frame (only the
 * ThresholdDetector can: the final fade-out when add_final_scene is set).  Synchronises. */
ESD_API int esd_post_process(esd_ctx* ctx, int32_t detector, int64_t last_frame_num, int64_t* cuts, int64_t cap,
                             int64_t* n_cuts);

/* Replaces: parsing the `pts_time:` lines of ffmpeg's showinfo output into cut timestamps
 * (model_manager.py:762-786); the scene dicts of :775-781 / :809-824 are built from these cuts on the host
 * (eioku_b200/service.py:scenes_to_dicts).
 * Cuts emitted so far by one detector (ESD_DET_*), starting at index `from_index` of its cut list.
 * *n_total receives the total number of cuts the detector has emitted.  Synchronises. */
ESD_API int esd_get_cuts(esd_ctx* ctx, int32_t detector, int64_t from_index, int64_t* cuts, int64_t cap,
                 int64_t* n_written, int64_t* n_total);

/* Stand-alone decision pass over score arrays already on the host (merge step of frame-range
 * sharding: shards return scores, one global pass decides).  Uses the ctx's detector parameters but
 * none of its frame state.  Runs on a library-owned stream with persistent pinned / device scratch (no allocation, no
 * device-wide synchronisation: other contexts of the device keep running).  For ESD_DET_ADAPTIVE `scores` = adaptive_val (ratios are recomputed);
 * ESD_DET_CONTENT: content_val; ESD_DET_HIST: hist_diff (NaN = no previous frame); ESD_DET_THRESHOLD: average_rgb;
 * ESD_DET_HASH: hash_dist (NaN = no previous frame). */
ESD_API int esd_decide_arrays(esd_ctx* ctx, int32_t detector, int64_t first_frame_num, int64_t n,
                      const double* scores, double* adaptive_ratio_out, int64_t* cuts, int64_t cap,
                      int64_t* n_cuts);

/* Same decision pass over scores ALREADY ON THE DEVICE of this ctx (e.g. score slabs peer-copied or all-gathered from the
 * shards of a long video, so the merge never visits the host).  Enqueued on `stream` (the caller's); waits for this pass
 * only (no device-wide synchronisation) and returns the cuts.  d_adaptive_ratio_out: optional device array [n] that
 * receives the recomputed adaptive ratios (ESD_DET_ADAPTIVE). */
ESD_API int esd_decide_device(esd_ctx* ctx, int32_t detector, int64_t first_frame_num, int64_t n, const double* d_scores,
                              double* d_adaptive_ratio_out, int64_t* cuts, int64_t cap, int64_t* n_cuts, void* stream);

/* One per-frame float64 score array of frames [from_frame, from_frame + n), copied device-to-device into `d_dst` on device
 * `dst_device` (peer copy over NVLink when it differs from the ctx's device; < 0 = the ctx's own device).  Asynchronous on
 * `stream` (a stream of the ctx's device), ordered behind every scoring tail enqueued so far. */
enum { ESD_SCORE_CONTENT_VAL = 0, ESD_SCORE_ADAPTIVE_VAL = 1, ESD_SCORE_HIST_DIFF = 2, ESD_SCORE_AVERAGE_RGB = 3,
       ESD_SCORE_HASH_DIST = 4, ESD_SCORE_ADAPTIVE_RATIO = 5 };
ESD_API int esd_copy_scores_device(esd_ctx* ctx, int32_t kind, int64_t from_frame, int64_t n, double* d_dst, int32_t dst_device,
                                   void* stream);

/* Instrumentation: CUDA-event timing of the fused scoring kernel on the launching stream. */
ESD_API int esd_set_timing(esd_ctx* ctx, int32_t enable);
/* total fused-kernel milliseconds and launch count since the last call (synchronises) */
ESD_API int esd_kernel_time(esd_ctx* ctx, double* fused_ms, int64_t* fused_launches);
/* every kernel the library launched on this ctx since esd_create */
ESD_API int64_t esd_kernel_launches(const esd_ctx* ctx);

/* Test hook: packed H | S<<8 | V<<16 of the last pushed frame at detector resolution
 * ([dst_height][dst_width] uint32), i.e. the on-device twin of cvtColor(resize(frame)).  Synchronises. */
ESD_API int esd_debug_read_prev(esd_ctx* ctx, uint32_t* out, int64_t cap_elems);
/* Test hook for the red-zone allocator (csrc/guard_alloc.h; ESD_GUARD=1 puts 256-byte guard zones around every device buffer of
 * the library and verifies them at free time -- this pool's stand-in for compute-sanitizer).  Returns 1 when the guards are
 * active (0: plain allocations); damage != 0 writes one byte past a guarded buffer first, which must abort the process. */
ESD_API int esd_debug_guard_selftest(int32_t damage);

#ifdef __cplusplus
}
#endif
#endif /* ESD_H */
