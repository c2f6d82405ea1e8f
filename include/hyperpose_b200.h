/*
 * hyperpose_b200.h -- C ABI of the B200-native HyperPose inference path.
 *
 * This is the drop-in boundary: the reference has no FFI, its seam is link-time
 * substitution of src/tensorrt.cpp + src/paf.cpp (CMakeLists.txt:28-34,
 * cmake/hyperpose.fake.cmake:6-19).  The C++ classes of include/hyperpose/operator/...
 * are re-implemented as thin wrappers over the functions below
 * (hyperpose_b200/csrc/hyperpose_api/); INTEGRATION.md shows the wiring.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns
 * HP_OK (0) or a negative hp_status; hp_last_error() gives a thread-local message.
 * The C++ wrappers translate statuses to the reference's conventions
 * (error() -> std::exit(-1), src/logging.hpp:31-37; std::logic_error, src/tensorrt.cpp:439-443).
 * There is NO CPU fallback: without a CUDA device every entry point fails with HP_ERR_CUDA.
 */
#ifndef HYPERPOSE_B200_H
#define HYPERPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HP_N_PARTS 18 /* include/hyperpose/utility/human.hpp:10 */
#define HP_N_PAIRS 19 /* include/hyperpose/utility/human.hpp:11 */

typedef enum hp_status {
    HP_OK = 0,
    HP_ERR_ARG = -1,      /* bad argument (rank/shape/null) -- reference: error() exit, paf.cpp:305-306 */
    HP_ERR_CAPACITY = -2, /* an internal or caller-provided capacity was exceeded */
    HP_ERR_UNSUPPORTED = -3,
    HP_ERR_CUDA = -4,     /* CUDA runtime failure / no device */
    HP_ERR_BATCH = -5     /* batch > max_batch -- reference: std::logic_error, tensorrt.cpp:439-443 */
} hp_status;

/* mirrors hyperpose::body_part_t / human_t (include/hyperpose/utility/human.hpp:14-31) */
typedef struct hp_body_part {
    int32_t has_value;
    float x, y, score;
} hp_body_part;
typedef struct hp_human {
    hp_body_part parts[HP_N_PARTS];
    float score;
} hp_human; /* 292 bytes */

/* debug/parity records (src/post_process.hpp:126-131, src/paf.cpp:7-13) */
typedef struct hp_peak {
    int32_t part_id, x, y;
    float score;
    int32_t id;
} hp_peak;
typedef struct hp_connection {
    int32_t cid1, cid2;
    float score;
} hp_connection;

const char* hp_last_error(void);
int hp_device_count(void);
const char* hp_version(void);

/* ------------------------------------------------------------------------------------------
 * PAF parser -- replaces hyperpose::parser::paf (include/hyperpose/operator/parser/paf.hpp:17-93,
 * src/paf.cpp:284-387, src/post_process.hpp:26-205).
 * ---------------------------------------------------------------------------------------- */
typedef struct hp_paf hp_paf;

/* paf::paf(conf_thresh, paf_thresh, resolution_size) (paf.hpp:27). res_w/res_h = -1 selects the
 * reference default (width 4*H, height 4*W of the first [C,H,W] input, paf.cpp:311-315).
 * device: CUDA ordinal. */
int hp_paf_create(hp_paf** out, float conf_thresh, float paf_thresh, int res_w, int res_h, int device);
void hp_paf_destroy(hp_paf* p);
/* paf::set_conf_thresh / set_paf_thresh (paf.hpp:66-70) */
int hp_paf_set_conf_thresh(hp_paf* p, float thresh);
int hp_paf_set_paf_thresh(hp_paf* p, float thresh);
/* internal capacities (the reference is unbounded; exceeding one returns HP_ERR_CAPACITY, the
 * *_host entry points grow and retry automatically). 0 keeps the current value. */
int hp_paf_set_capacity(hp_paf* p, int max_peaks_per_part, int max_candidates_per_limb, int max_humans);

/* paf::process(conf, paf) (paf.hpp:48, paf.cpp:300-375): one frame, HOST tensors
 * conf[c_conf,H,W], paf[c_paf,H,W] float32 row-major.  Writes <= cap humans, *n_out = count. */
int hp_paf_process_host(hp_paf* p, const float* conf, const float* paf, int c_conf, int c_paf, int H, int W,
                        hp_human* out, int cap, int* n_out);
/* N frames in one call (what the stream's parse stage would hand over, stream.hpp:347-385):
 * conf[N,c_conf,H,W], paf[N,c_paf,H,W]; out[N*cap], n_out[N]. */
int hp_paf_process_host_batched(hp_paf* p, const float* conf, const float* paf, int N, int c_conf, int c_paf, int H, int W,
                                hp_human* out, int cap, int* n_out);
/* Device-resident inputs (engine -> parser hand-off without the reference's D2H/H2D,
 * tensorrt.cpp:398-431).  Enqueues on `stream` (a cudaStream_t; NULL = the parser's own stream);
 * results stay on the device until hp_paf_fetch. */
int hp_paf_process_device(hp_paf* p, const float* d_conf, const float* d_paf, int N, int c_conf, int c_paf, int H, int W,
                          void* stream);
/* Copies the humans of the last hp_paf_process_device call to the host (synchronises its stream). */
int hp_paf_fetch(hp_paf* p, hp_human* out, int cap, int* n_out, int N);

/* parity/debug: peak list (scan-ordered, ids as in post_process.hpp:175-192) and per-limb
 * connections (paf.cpp:252-270) of frame `frame` of the last call. */
int hp_paf_debug_peaks(hp_paf* p, int frame, hp_peak* out, int cap, int* n_out);
int hp_paf_debug_connections(hp_paf* p, int frame, int pair_id, hp_connection* out, int cap, int* n_out);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
long long hp_paf_launch_count(const hp_paf* p);

#ifdef __cplusplus
}
#endif
#endif /* HYPERPOSE_B200_H */
