// handoff.h -- device-resident engine -> parser hand-off BEHIND the reference's host-only `feature_map_t`
// (include/hyperpose/stream/stream.hpp:326-385, src/tensorrt.cpp:398-431).
//
// The reference API moves every network output through host memory: tensorrt::inference() returns one host
// `feature_map_t` per image and output, and the caller (operator API loop, or the stream's thread-pool tasks) hands
// them to parser.process() one image at a time.  The drop-in keeps that interface (the headers are unchanged), but the
// bytes do not have to travel twice: when the engine fills the host buffers it also keeps a device-side snapshot of the
// batch and PUBLISHES the host addresses.  A parser that is handed a published address -- same pointers, same shape,
// same sampled contents -- parses the whole batch once from the snapshot (one batched launch sequence instead of N
// H2D copies + N launch sequences) and serves the other N-1 process() calls from the cached records.
// `feature_map_t` exposes its data only through `const T* view() const` (include/hyperpose/utility/data.hpp:40-45), so a
// published buffer cannot be modified through the API; address reuse after a `feature_map_t` died is caught by the
// content fingerprint and by the bounded lifetime of a publication (HANDOFF_RING batches per engine).
// A miss of any kind falls back to the ordinary host path; results are identical either way (same kernels).
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/hyperpose_b200.h"

namespace hpb {
namespace handoff {

constexpr int HANDOFF_RING = 4;     // published batches kept per engine
constexpr int FP_SAMPLES = 48;      // fingerprint: this many floats of each tensor, evenly spread

struct Fingerprint {
    std::array<float, FP_SAMPLES> a, b;
};

// One published batch: device snapshot of the two output tensors + the host addresses they were copied to.
struct Batch {
    std::mutex mu;                   // serialises the lazy batched parse, fingerprint checks and retirement
    bool valid = false;
    int fail_count = 0;              // batched parses that overflowed a parser capacity; >= 2: stop trying on this batch
    int device = 0;
    int N = 0;
    size_t elems_a = 0, elems_b = 0; // floats per frame of tensor a (conf | pif) and b (paf)
    float* d_a = nullptr; float* d_b = nullptr;
    size_t cap_a = 0, cap_b = 0;     // floats allocated
    cudaEvent_t ready = nullptr;     // snapshot copies complete
    std::vector<const float*> host_a, host_b;
    std::vector<Fingerprint> fp;
    // cache of the batched parse: key = parser kind + its parameters
    int cache_kind = 0;              // 0 empty | 1 PAF | 2 PifPaf
    float key_f[2] = { 0, 0 };
    int key_i[2] = { 0, 0 };
    int hcap = 0;
    std::vector<hp_human> humans;    // [N][hcap]
    std::vector<int> counts;         // [N]
};

struct Hit {
    std::shared_ptr<Batch> batch;
    int frame = -1;
};

bool enabled();
void set_enabled(bool on);

// engine side -------------------------------------------------------------------------------------------------------
// Takes the engine's next ring slot (retiring whatever it held), snapshots d_a / d_b (N frames) on `st` and registers
// the host addresses.  `host_stage_a/b` = the bytes just copied to those addresses (fingerprint source).
int publish(std::shared_ptr<Batch>* ring, int* ring_pos, int device, cudaStream_t st, const float* d_a, const float* d_b, int N,
            size_t elems_a, size_t elems_b, float* const* host_a, float* const* host_b, const float* host_stage_a,
            const float* host_stage_b);
// engine teardown: unregister and free every slot of the ring
void retire_ring(std::shared_ptr<Batch>* ring);

// parser side -------------------------------------------------------------------------------------------------------
// Registry lookup by host address of tensor a (then b and the per-frame sizes must match).  No content check yet.
Hit lookup(const float* host_a, const float* host_b, size_t elems_a, size_t elems_b);
// Content check of frame `frame` (caller holds batch->mu): the sampled floats still equal what was published.
bool fingerprint_matches(const Batch& b, int frame);
void sample(const float* a, size_t elems_a, const float* b, size_t elems_b, Fingerprint* out);

void count_hit();
void count_batch_parse();
void count_miss();
void stats(long long* published, long long* hits, long long* batch_parses, long long* misses);

} // namespace handoff
} // namespace hpb
